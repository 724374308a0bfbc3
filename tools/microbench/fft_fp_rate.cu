// FP-only rate of the in-register DFT32 + twiddle-multiply burst (the arithmetic of phases B / B' of
// the FftFilter kernel) with no memory traffic: what fraction of the FP32 issue rate does this
// instruction stream reach with 16, 8 or 4 warps per SM?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I../../rustradio_b200/csrc -o fft_fp_rate fft_fp_rate.cu
#include <cstdio>
#include "fft_regs.cuh"
using namespace rrc::fftr;
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("cuda error %s line %d\n", cudaGetErrorString(e), __LINE__); return 1;}}while(0)

template <int MODE>
__global__ void __launch_bounds__(512, 1) fp_burst(float2* out, const float2* tw, int iters, float onef) {
    extern __shared__ float2 sm[];
    float2 v[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = make_float2(threadIdx.x * 1e-3f + i, 0.5f * i);
    float2 w[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) w[i] = tw[(threadIdx.x + i) & 511];
    float one = onef;
    for (int it = 0; it < iters; ++it) {
        if (MODE == 0) dit_g<32, +1>(v, one); else dit_g<32, -1>(v, one);
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = cmul(v[i], w[i & 7]);
    }
    float2 s = make_float2(0.f, 0.f);
#pragma unroll
    for (int i = 0; i < 32; ++i) s = cadd(s, v[i]);
    if (s.x == 123.456f) out[threadIdx.x] = s;
}

int main() {
    float2 *out, *tw; CK(cudaMalloc(&out, 4096 * 8)); CK(cudaMalloc(&tw, 512 * 8)); CK(cudaMemset(tw, 0, 512 * 8));
    CK(cudaFuncSetAttribute(fp_burst<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 2000;
    const int fp_per_iter = FP_PER_ITER;   // SASS FP instructions per loop iteration (counted with cuobjdump, passed by -D)
    for (int threads : {512, 256, 128}) {
        float ms = 0;
        for (int rep = 0; rep < 2; ++rep) {
            cudaEventRecord(e0);
            fp_burst<0><<<148, threads, 200 * 1024>>>(out, tw, iters, 1.0f);
            cudaEventRecord(e1); CK(cudaEventSynchronize(e1)); cudaEventElapsedTime(&ms, e0, e1);
        }
        const double clk = ms * 1e-3 * 1.965e9;
        const double winstr = (double)iters * fp_per_iter * (threads / 32);
        printf("warps/SM %2d: %.3f ms, %.3f FP warp-instr/clk/SM (peak 4), %.1f cycles per DFT32+twiddle burst per warp\n",
               threads / 32, ms, winstr / clk, clk / iters);
    }
    return 0;
}
