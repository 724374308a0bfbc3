// Does an L2 prefetch issued ~20 K cycles ahead turn the FftFilter kernel's phase-A loads into L2
// hits?  Every CTA streams fresh 128 KiB segments (HBM); per iteration: [prefetch segment it+1]
// [timed: load segment it with 32 x LDG.64 per thread] [spin 20 K cycles].  Prefetch flavours:
// 0 none, 1 cp.async.bulk.prefetch.L2 (16 x 8 KiB, one per warp), 2 prefetch.global.L2 per 128 B line,
// 3 cp.async.bulk.prefetch.L2 issued by ONE thread (16 sequential).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o l2_prefetch l2_prefetch.cu
#include <cstdio>
#include <cuda_runtime.h>
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("cuda error %s line %d\n", cudaGetErrorString(e), __LINE__); return 1;}}while(0)

template <int MODE>
__global__ void __launch_bounds__(512, 1) k(const float2* __restrict__ in, float2* out, int iters, int spin, long long* cyc) {
    extern __shared__ __align__(128) float2 sm[];
    const int tid = threadIdx.x;
    float2 acc = make_float2(0.f, 0.f);
    long long tl = 0;
    for (int it = 0; it < iters; ++it) {
        const float2* seg = in + ((size_t)it * gridDim.x + blockIdx.x) * 16384;
        const float2* nxt = seg + (size_t)gridDim.x * 16384;
        if (MODE == 1 && (tid & 31) == 0)
            asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(nxt + (tid >> 5) * 1024), "r"(8192) : "memory");
        if (MODE == 2) {
            asm volatile("prefetch.global.L2 [%0];" ::"l"(reinterpret_cast<const char*>(nxt) + tid * 128));
            asm volatile("prefetch.global.L2 [%0];" ::"l"(reinterpret_cast<const char*>(nxt) + (tid + 512) * 128));
        }
        if (MODE == 3 && tid == 0) {
#pragma unroll 1
            for (int c = 0; c < 16; ++c) asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(nxt + c * 1024), "r"(8192) : "memory");
        }
        __syncthreads();
        long long t0 = clock64();
        float2 v[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = __ldcg(seg + tid + 512 * i);
#pragma unroll
        for (int i = 0; i < 32; ++i) { acc.x += v[i].x; acc.y += v[i].y; }
        __syncthreads();
        long long t1 = clock64();
        if (it > 0) tl += t1 - t0;
        while (clock64() - t1 < spin) { }
    }
    if (tid == 0) cyc[blockIdx.x] = tl;
    if (acc.x == 123.456f) out[tid] = acc;
}

int main() {
    const size_t seg = 16384; const int iters = 100;
    float2 *in, *out; long long* cyc;
    CK(cudaMalloc(&in, (size_t)(iters + 1) * 148 * seg * 8)); CK(cudaMemset(in, 0, (size_t)(iters + 1) * 148 * seg * 8));
    CK(cudaMalloc(&out, 4096 * 8)); CK(cudaMalloc(&cyc, 148 * 8));
    const int SMEM = 200 * 1024;
    CK(cudaFuncSetAttribute(k<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
    CK(cudaFuncSetAttribute(k<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
    CK(cudaFuncSetAttribute(k<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
    CK(cudaFuncSetAttribute(k<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
    const char* names[] = {"no prefetch", "bulk prefetch.L2, 1 per warp", "prefetch.global.L2 per line", "bulk prefetch.L2, 16 by one thread"};
    float2* flush; CK(cudaMalloc(&flush, 512 << 20));
    for (int spin : {20000, 5000}) for (int mode = 0; mode < 4; ++mode) {
        CK(cudaMemset(flush, 1, 512 << 20));   // evict the input from L2
        if (mode == 0) k<0><<<148, 512, SMEM>>>(in, out, iters, spin, cyc);
        if (mode == 1) k<1><<<148, 512, SMEM>>>(in, out, iters, spin, cyc);
        if (mode == 2) k<2><<<148, 512, SMEM>>>(in, out, iters, spin, cyc);
        if (mode == 3) k<3><<<148, 512, SMEM>>>(in, out, iters, spin, cyc);
        CK(cudaDeviceSynchronize());
        long long h[148]; CK(cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost));
        double avg = 0; for (int i = 0; i < 148; ++i) avg += h[i]; avg /= 148;
        printf("spin %5d  %-36s load phase %7.0f cycles per 128 KiB segment (%5.1f B/clk/SM)\n", spin, names[mode], avg / (iters - 1), 131072.0 / (avg / (iters - 1)));
    }
    return 0;
}
