// Legacy tensor path on sm_100a: issue rate of mma.sync.m16n8k16 (bf16 x bf16 -> f32, SASS HMMA) per SM,
// with and without an ldmatrix.x4 of the A fragment per mma — the numbers a Toeplitz-block FIR on
// mma.sync would be bounded by.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o hmma_rate hmma_rate.cu
#include <cstdio>
#include <cstdint>
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("cuda error %s line %d\n", cudaGetErrorString(e), __LINE__); return 1;}}while(0)

__device__ __forceinline__ void mma16816(float (&c)[4], const unsigned (&a)[4], const unsigned (&b)[2]) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

template <int MODE>   // 0: mma only; 1: ldmatrix.x4 + mma
__global__ void __launch_bounds__(256) k(float* out, int iters, long long* cyc) {
    __shared__ __align__(16) unsigned short sm[8192];
    for (int i = threadIdx.x; i < 8192; i += 256) sm[i] = (unsigned short)(0x3f80 + (i & 7));
    __syncthreads();
    float c[8][4] = {};
    unsigned a[4] = {0x3f803f80u, 0x3f803f80u, 0x3f803f80u, 0x3f803f80u}, b[2] = {0x3f803f80u, 0x3f003f00u};
    const unsigned base = (unsigned)__cvta_generic_to_shared(sm) + (threadIdx.x & 31) * 16 + (threadIdx.x >> 5) * 1024;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            if (MODE == 1)
                asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(a[0]), "=r"(a[1]), "=r"(a[2]), "=r"(a[3]) : "r"(base + ((it + j) & 7) * 16));
            mma16816(c[j], a, b);
        }
    }
    long long t1 = clock64();
    float s = 0; for (int j = 0; j < 8; ++j) for (int q = 0; q < 4; ++q) s += c[j][q];
    out[blockIdx.x * 256 + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

int main() {
    float* out; long long* cyc; CK(cudaMalloc(&out, 148 * 4 * 256 * 4)); CK(cudaMalloc(&cyc, 8));
    const int iters = 4096;
    for (int mode = 0; mode < 2; ++mode) for (int ctas = 1; ctas <= 4; ctas *= 2) {
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        for (int rep = 0; rep < 2; ++rep) {
            cudaEventRecord(e0);
            if (mode == 0) k<0><<<148 * ctas, 256>>>(out, iters, cyc); else k<1><<<148 * ctas, 256>>>(out, iters, cyc);
            cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
        }
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        long long c; CK(cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost));
        const double mmas_per_sm = (double)iters * 8 * 8 * ctas;       // 8 warps per CTA
        printf("%s, %d CTA(s)/SM x 8 warps: %.3f ms, %.3f mma.m16n8k16 per clk per SM (clock64 of CTA 0: %.3f), %.1f dense bf16 TFLOP/s\n",
               mode ? "ldmatrix.x4 + mma" : "mma only", ctas, ms, mmas_per_sm / (ms * 1e-3 * 1.965e9), (double)iters * 64 / c,
               mmas_per_sm * 148 * 4096 / (ms * 1e-3) / 1e12);
    }
    return 0;
}
