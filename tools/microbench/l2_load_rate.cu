// L2 -> SM load rate with the FftFilter kernel's footprint (one 512-thread CTA per SM, 216 KiB of
// shared memory so L1 is ~12 KiB): every CTA re-reads its own 128 KiB segment (L2 hits) or streams
// fresh segments (HBM), with (a) the kernel's pattern: 32 x LDG.64 per thread at stride 512 elements,
// (b) LDG.128 fully coalesced, (c) cp.async.bulk (TMA) 8 KiB chunks into shared memory.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o l2_load_rate l2_load_rate.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("cuda error %s line %d\n", cudaGetErrorString(e), __LINE__); return 1;}}while(0)

template <int MODE>
__global__ void __launch_bounds__(512, 1) load_kernel(const float2* __restrict__ in, float2* out, int iters, long long stride_per_iter, long long* cyc) {
    extern __shared__ __align__(128) float2 sm[];
    __shared__ __align__(8) unsigned long long mbar;
    const int tid = threadIdx.x;
    float2 acc = make_float2(0.f, 0.f);
    const unsigned mb = (unsigned)__cvta_generic_to_shared(&mbar);
    if (MODE == 2 && tid == 0) { asm volatile("mbarrier.init.shared.b64 [%0], 1;" ::"r"(mb)); }
    __syncthreads();
    long long t0 = clock64();
    unsigned phase = 0;
    for (int it = 0; it < iters; ++it) {
        const float2* seg = in + (size_t)blockIdx.x * 16384 + (size_t)it * stride_per_iter;
        if (MODE == 0) {
            float2 v[32];
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = __ldcg(seg + tid + 512 * i);
#pragma unroll
            for (int i = 0; i < 32; ++i) { acc.x += v[i].x; acc.y += v[i].y; }
        } else if (MODE == 1) {
            float4 v[16];
            const float4* s4 = reinterpret_cast<const float4*>(seg);
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] = __ldcg(s4 + tid + 512 * i);
#pragma unroll
            for (int i = 0; i < 16; ++i) { acc.x += v[i].x + v[i].z; acc.y += v[i].y + v[i].w; }
        } else {
            if (tid == 0) {
                asm volatile("mbarrier.arrive.expect_tx.shared.b64 _, [%0], %1;" ::"r"(mb), "r"(131072));
#pragma unroll 1
                for (int c = 0; c < 16; ++c) {
                    asm volatile("cp.async.bulk.shared::cta.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                                 ::"r"((unsigned)__cvta_generic_to_shared(sm + c * 1024)), "l"(seg + c * 1024), "r"(8192), "r"(mb) : "memory");
                }
            }
            unsigned done = 0;
            while (!done) {
                asm volatile("{.reg .pred p; mbarrier.try_wait.parity.shared.b64 p, [%1], %2; selp.u32 %0, 1, 0, p;}" : "=r"(done) : "r"(mb), "r"(phase) : "memory");
            }
            phase ^= 1;
            acc.x += sm[tid].x;
            __syncthreads();
        }
    }
    long long t1 = clock64();
    if (tid == 0) cyc[blockIdx.x] = t1 - t0;
    if (acc.x == 123.456f) out[tid] = acc;
}

int main() {
    const size_t seg = 16384, nseg_hbm = 148 * 200;
    float2 *in, *out; long long* cyc;
    CK(cudaMalloc(&in, nseg_hbm * seg * 8)); CK(cudaMemset(in, 0, nseg_hbm * seg * 8));
    CK(cudaMalloc(&out, 4096 * 8)); CK(cudaMalloc(&cyc, 148 * 8));
    const int SMEM = 200 * 1024;
    CK(cudaFuncSetAttribute(load_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
    CK(cudaFuncSetAttribute(load_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
    CK(cudaFuncSetAttribute(load_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
    const char* names[] = {"LDG.64 stride-512 (kernel pattern)", "LDG.128 coalesced", "cp.async.bulk 16 x 8 KiB -> smem"};
    for (int src = 0; src < 2; ++src) for (int mode = 0; mode < 3; ++mode) {
        const int iters = 200;
        const long long stride = src == 0 ? 0 : 148ll * seg;     // 0: same segment every iteration (L2 hit); else fresh data (HBM)
        for (int rep = 0; rep < 2; ++rep) {
            if (mode == 0) load_kernel<0><<<148, 512, SMEM>>>(in, out, iters, stride, cyc);
            if (mode == 1) load_kernel<1><<<148, 512, SMEM>>>(in, out, iters, stride, cyc);
            if (mode == 2) load_kernel<2><<<148, 512, SMEM>>>(in, out, iters, stride, cyc);
            CK(cudaDeviceSynchronize());
        }
        long long h[148]; CK(cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost));
        double avg = 0; for (int i = 0; i < 148; ++i) avg += h[i]; avg /= 148;
        printf("%-4s %-38s %8.0f cycles per 128 KiB segment per SM = %6.1f B/clk/SM = %6.2f TB/s chip\n", src ? "HBM" : "L2", names[mode],
               avg / iters, 131072.0 / (avg / iters), 131072.0 / (avg / iters) * 148 * 1.965e9 / 1e12);
    }
    return 0;
}
