// Does the SM overlap FFMA2 (packed FP32) with shared-memory loads?  One loop body = NF FFMA2 on 8 independent
// accumulators (tap from a uniform register, like fir_rtu_kernel) + NL LDS.64 (conflict free, thread-strided
// like the FIR window reads) whose results feed the next iteration's FFMA2.  Reports cycles per iteration per SM
// for FP only / LDS only / both, for several warps-per-SM counts: if "both" ~ max(FP, LDS) the pipes overlap,
// if "both" ~ FP + LDS they serialise.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o ffma2_lds_overlap ffma2_lds_overlap.cu
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ void fma2(u64& acc, u64 hh, u64 x) { asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc) : "l"(hh), "l"(x)); }
struct Taps { float2 hh[64]; };

template <int MODE, int WIDE>   // MODE 0: FP only, 1: LDS only, 2: both ; WIDE: 0 = LDS.64, 1 = LDS.128
__global__ void __launch_bounds__(512) k(const __grid_constant__ Taps t, int iters, u64* out, long long* cyc) {
    extern __shared__ u64 sm[];
    const int tid = threadIdx.x;
    for (int i = tid; i < 82 * 128; i += blockDim.x) sm[i] = i;
    __syncthreads();
    u64 acc[8]; u64 w[8];
#pragma unroll
    for (int r = 0; r < 8; ++r) { acc[r] = 0; w[r] = r + tid; }
    const u64* base = sm + (tid & 127) * 81;
    const u64* base2 = sm + (tid & 127) * 82;          // 16-byte aligned rows for LDS.128 (2-way conflict between lanes l, l+8)
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            if (MODE != 0) {
                if (WIDE) { if ((q & 1) == 0) { u64 a, b; asm volatile("ld.shared.v2.b64 {%0, %1}, [%2];" : "=l"(a), "=l"(b) : "r"((unsigned)__cvta_generic_to_shared(base2 + (((it * 8 + q) & 62))))); w[q] ^= a; w[q + 1] ^= b; } }
                else { u64 v; asm volatile("ld.shared.b64 %0, [%1];" : "=l"(v) : "r"((unsigned)__cvta_generic_to_shared(base + ((it * 8 + q) & 63)))); w[q] ^= v; }
            }
            if (MODE != 1) {
                const u64 hh = *reinterpret_cast<const u64*>(&t.hh[q]);
#pragma unroll
                for (int r = 0; r < 8; ++r) fma2(acc[r], hh, w[(r + q) & 7]);
            }
        }
    }
    long long t1 = clock64();
    u64 s = 0;
#pragma unroll
    for (int r = 0; r < 8; ++r) s += acc[r] + w[r];
    out[blockIdx.x * blockDim.x + tid] = s;
    if (tid == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int MODE, int WIDE>
double run(int warps, int iters) {
    Taps t; for (int i = 0; i < 64; i++) t.hh[i] = make_float2(1.0f + i * 1e-3f, 1.0f + i * 1e-3f);
    u64* out; long long* cyc; cudaMalloc(&out, 148 * 512 * 8); cudaMalloc(&cyc, 148 * 8);
    const size_t smem = 82 * 128 * 8 + 1024;
    cudaFuncSetAttribute(k<MODE, WIDE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    k<MODE, WIDE><<<148, warps * 32, smem>>>(t, iters, out, cyc);
    k<MODE, WIDE><<<148, warps * 32, smem>>>(t, iters, out, cyc);
    cudaDeviceSynchronize();
    long long h[148]; cudaMemcpy(h, cyc, sizeof h, cudaMemcpyDeviceToHost);
    double m = 0; for (int i = 0; i < 148; i++) m += h[i];
    cudaFree(out); cudaFree(cyc);
    return m / 148 / iters;
}

int main() {
    const int iters = 4000;
    printf("cycles per iteration (one iteration per warp = 64 FFMA2 + 8 LDS.64 [or 4 LDS.128]); pipe floors per SM: FFMA2 32 cycles/warp-iter, LDS 16 wavefronts/warp-iter\n");
    for (int warps : {1, 2, 4, 8, 16}) {
        const double fp = run<0, 0>(warps, iters), ld = run<1, 0>(warps, iters), both = run<2, 0>(warps, iters), ldw = run<1, 1>(warps, iters), bothw = run<2, 1>(warps, iters);
        printf("warps/SM %2d: FP only %7.1f  LDS.64 only %7.1f  both %7.1f  (max %7.1f sum %7.1f) | LDS.128 only %7.1f both %7.1f\n", warps, fp, ld, both, fp > ld ? fp : ld, fp + ld, ldw, bothw);
    }
    return 0;
}
