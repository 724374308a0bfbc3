// Do a shared-memory burst and an FP burst overlap across warps of one SM?  16 warps per SM, each
// loops  [32 x LDS.64 column read] -> [DFT32 + twiddle multiply: 516 FP instr] -> [32 x STS.64] ,
// the per-warp structure of the FftFilter kernel's phase B, with (a) all warps starting together,
// (b) the 8 warps of group 1 delayed by half a period, (c) token ping-pong between the groups
// (named barriers), (d) only the FP part, (e) only the shared-memory part.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I../../rustradio_b200/csrc -o phase_overlap phase_overlap.cu
#include <cstdio>
#include "fft_regs2.cuh"
using namespace rrc::fftr;
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("cuda error %s line %d\n", cudaGetErrorString(e), __LINE__); return 1;}}while(0)

template <int MODE>
__global__ void __launch_bounds__(512, 1) k(float2* out, int iters, int delay) {
    extern __shared__ __align__(16) float2 sm[];
    __shared__ float s_one;
    const int tid = threadIdx.x;
    for (int i = tid; i < 32 * 544; i += 512) sm[i] = make_float2(1e-3f * i, 0.f);
    if (tid == 0) s_one = 1.0f;
    __syncthreads();
    const int g = (tid >> 7) & 1;
    float2* col = sm + (tid >> 4) * 544 + (tid & 15);
    const unsigned one_addr = (unsigned)__cvta_generic_to_shared(&s_one);
    if (MODE == 1 && g == 1) { const long long t0 = clock64(); while (clock64() - t0 < delay) { } }
    if (MODE == 2 && g == 1) asm volatile("bar.arrive 1, 512;" ::: "memory");
    float2 v[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = make_float2(tid * 1e-3f, i);
    for (int it = 0; it < iters; ++it) {
        if (MODE != 3) {
#pragma unroll
            for (int i = 0; i < 32; ++i) v[bitrev(i, 5)] = col[i * 17];
        }
        float one;
        if (MODE == 2) asm volatile("bar.sync %1, 512;\n\tld.volatile.shared.f32 %0, [%2];" : "=f"(one) : "r"(1 + g), "r"(one_addr) : "memory");
        else asm volatile("ld.volatile.shared.f32 %0, [%1];" : "=f"(one) : "r"(one_addr) : "memory");
        if (MODE != 4) {
            dit_g<32, +1>(v, one);
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = cmul(v[i], make_float2(0.999f, 0.001f * (i & 7)));
        }
        if (MODE == 2) asm volatile("bar.arrive %0, 512;" ::"r"(2 - g) : "memory");
        if (MODE != 3) {
#pragma unroll
            for (int i = 0; i < 32; ++i) col[i * 17] = v[i];
            __syncwarp();
        }
        if (MODE == 5 && (it & 3) == 3) __syncthreads();
        if (MODE == 6 && (it & 1) == 1) __syncthreads();
    }
    float2 s = make_float2(0.f, 0.f);
#pragma unroll
    for (int i = 0; i < 32; ++i) s = cadd(s, v[i]);
    if (s.x == 123.456f) out[tid] = s;
}

// Packed variant: 8 warps, each thread owns TWO columns (64 points) as C2 pairs; per iteration
// 64 LDS.64 (re-pair and im-pair words), DFT32 + twiddle multiply on pairs, 64 STS.64: the same
// work per CTA as the 16-warp scalar kernel above.
template <int MODE>
__global__ void __launch_bounds__(256, 1) k2(float2* out, int iters) {
    extern __shared__ __align__(16) float2 sm[];
    __shared__ float s_one;
    const int tid = threadIdx.x;
    for (int i = tid; i < 32 * 552; i += 256) sm[i] = make_float2(1e-3f * i, 0.f);
    if (tid == 0) s_one = 1.0f;
    __syncthreads();
    // plane pitch 552 words (16 banks off per plane): the two planes of a half-warp use disjoint banks
    F2* col = reinterpret_cast<F2*>(sm) + (tid >> 3) * 552 + (tid & 7);      // word (row r): re pair at r*17 + c, im pair at r*17 + 8 + c
    const unsigned one_addr = (unsigned)__cvta_generic_to_shared(&s_one);
    C2 v[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = C2{f2(tid * 1e-3f, i), f2(i, 1.f)};
    for (int it = 0; it < iters; ++it) {
        if (MODE != 3) {
#pragma unroll
            for (int i = 0; i < 32; ++i) { v[bitrev(i, 5)].re = col[i * 17]; v[bitrev(i, 5)].im = col[i * 17 + 8]; }
        }
        float one;
        asm volatile("ld.volatile.shared.f32 %0, [%1];" : "=f"(one) : "r"(one_addr) : "memory");
        if (MODE != 4) {
            dit2<32, +1>(v, one);
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = c2_mul(v[i], C2{splat(0.999f), splat(0.001f * (i & 7))});
        }
        if (MODE != 3) {
#pragma unroll
            for (int i = 0; i < 32; ++i) { col[i * 17] = v[i].re; col[i * 17 + 8] = v[i].im; }
            __syncwarp();
        }
        if (MODE == 5 && (it & 3) == 3) __syncthreads();
    }
    C2 s = v[0];
#pragma unroll
    for (int i = 1; i < 32; ++i) s = c2_add(s, v[i]);
    if (f2_lo(s.re) == 123.456f) out[tid] = make_float2(f2_lo(s.re), f2_hi(s.im));
}

int main() {
    float2* out; CK(cudaMalloc(&out, 4096 * 8));
    const int SMEM = 200 * 1024, iters = 2000;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const char* names[] = {"all 16 warps in phase", "group 1 delayed by half a period", "token ping-pong (named barriers)", "FP only", "shared memory only", "free-running + __syncthreads every 4 iterations", "free-running + __syncthreads every 2 iterations"};
    float ms[7];
    for (int mode = 0; mode < 7; ++mode) {
        for (int rep = 0; rep < 2; ++rep) {
            cudaEventRecord(e0);
            switch (mode) {
                case 0: CK(cudaFuncSetAttribute(k<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM)); k<0><<<148, 512, SMEM>>>(out, iters, 0); break;
                case 1: CK(cudaFuncSetAttribute(k<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM)); k<1><<<148, 512, SMEM>>>(out, iters, 1800); break;
                case 2: CK(cudaFuncSetAttribute(k<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM)); k<2><<<148, 512, SMEM>>>(out, iters, 0); break;
                case 3: CK(cudaFuncSetAttribute(k<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM)); k<3><<<148, 512, SMEM>>>(out, iters, 0); break;
                case 4: CK(cudaFuncSetAttribute(k<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM)); k<4><<<148, 512, SMEM>>>(out, iters, 0); break;
                case 5: CK(cudaFuncSetAttribute(k<5>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM)); k<5><<<148, 512, SMEM>>>(out, iters, 0); break;
                case 6: CK(cudaFuncSetAttribute(k<6>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM)); k<6><<<148, 512, SMEM>>>(out, iters, 0); break;
            }
            cudaEventRecord(e1); CK(cudaEventSynchronize(e1)); cudaEventElapsedTime(&ms[mode], e0, e1);
        }
        printf("%-52s %8.3f ms  %7.0f cycles per iteration (16 warps x [32 LDS.64, 516 FP, 32 STS.64])\n", names[mode], ms[mode], ms[mode] * 1e-3 * 1.965e9 / iters);
    }
    const char* names2[] = {"PACKED 8 warps x 64 points, free-running", "", "", "PACKED FP only", "PACKED shared memory only", "PACKED + __syncthreads every 4 iterations"};
    for (int mode : {0, 3, 4, 5}) {
        float t = 0;
        for (int rep = 0; rep < 2; ++rep) {
            cudaEventRecord(e0);
            switch (mode) {
                case 0: CK(cudaFuncSetAttribute(k2<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM)); k2<0><<<148, 256, SMEM>>>(out, iters); break;
                case 3: CK(cudaFuncSetAttribute(k2<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM)); k2<3><<<148, 256, SMEM>>>(out, iters); break;
                case 4: CK(cudaFuncSetAttribute(k2<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM)); k2<4><<<148, 256, SMEM>>>(out, iters); break;
                case 5: CK(cudaFuncSetAttribute(k2<5>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM)); k2<5><<<148, 256, SMEM>>>(out, iters); break;
            }
            cudaEventRecord(e1); CK(cudaEventSynchronize(e1)); cudaEventElapsedTime(&t, e0, e1);
        }
        printf("%-52s %8.3f ms  %7.0f cycles per iteration (8 warps x [64 LDS.64, DFT32+twiddle on pairs, 64 STS.64])\n", names2[mode], t, t * 1e-3 * 1.965e9 / iters);
    }
    return 0;
}
