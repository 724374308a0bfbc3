// Per-phase cycle trace of the TMA-staged FftFilter kernel (fftfilt_tma_kernel in fftfilt.cu; same
// phase functions and staging code), every warp stamping clock64() at each boundary for a few blocks.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I../../rustradio_b200/csrc -o fftfilt_tma_trace fftfilt_tma_trace.cu
#include <algorithm>
#include <cstdio>
#include <vector>
#include "fftfilt_core.cuh"
#include "fftfilt_tables.hpp"
using namespace rrc::fftk;
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("cuda error %s line %d\n", cudaGetErrorString(e), __LINE__); return 1;}}while(0)
constexpr int NSTAMP = 16, NTRACE_BLK = 6;
constexpr size_t SMEM = (size_t)(SMEM_ELEMS + 512 + 512 + HRES_ELEMS + 2) * sizeof(float2);

__device__ __forceinline__ long long clk() { long long c; asm volatile("mov.u64 %0, %%clock64;" : "=l"(c) :: "memory"); return c; }
__device__ __forceinline__ void mbar_wait(unsigned mbar, unsigned parity) {
    asm volatile("{\n.reg .pred p;\nMBAR_WAIT:\nmbarrier.try_wait.parity.shared.b64 p, [%0], %1;\n@p bra MBAR_DONE;\nbra MBAR_WAIT;\nMBAR_DONE:\n}" ::"r"(mbar), "r"(parity) : "memory");
}

__global__ void __launch_bounds__(NT, 1)
trace_kernel(const BlockIO io, const float2* __restrict__ Hp, const float2* __restrict__ tw1g, const float2* __restrict__ tw2g,
             long long nblocks, long long* trace, int early) {
    extern __shared__ __align__(16) float2 sm[];
    float2* s_tw2 = sm + SMEM_ELEMS; float2* s_tw1 = s_tw2 + 512; float2* s_hres = s_tw1 + 512;
    const int tid = threadIdx.x;
    s_tw2[tid] = tw2g[tid]; s_tw1[tid] = tw1g[tid];
    load_hres(tid, Hp, s_hres);
    const unsigned mbar = (unsigned)__cvta_generic_to_shared(s_hres + HRES_ELEMS);
    const unsigned sm_a = (unsigned)__cvta_generic_to_shared(sm);
    if (tid == 0) { asm volatile("mbarrier.init.shared.b64 [%0], 1;" ::"r"(mbar) : "memory"); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    __syncthreads();
    auto stage = [&](long long nb) {
        if (stage_linear_bulk_ok(nb, io)) {
            if (tid == 0) {
                const float2* src = io.in + stage_linear_seg0(nb, io);
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                asm volatile("mbarrier.arrive.expect_tx.shared.b64 _, [%0], %1;" ::"r"(mbar), "r"(N * 8) : "memory");
#pragma unroll 1
                for (int c = 0; c < 8; ++c)
                    asm volatile("cp.async.bulk.shared::cta.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                                 ::"r"(sm_a + c * 16384), "l"(src + c * 2048), "r"(16384), "r"(mbar) : "memory");
            }
        } else {
            stage_linear_fallback(tid, nb, io, sm);
            if (tid == 0) asm volatile("mbarrier.arrive.shared.b64 _, [%0];" ::"r"(mbar) : "memory");
        }
    };
    if (blockIdx.x < nblocks) stage(blockIdx.x);
    unsigned parity = 0;
    int it = 0;
    for (long long blk = blockIdx.x; blk < nblocks; blk += gridDim.x, ++it) {
        const bool tr = blockIdx.x == 3 && it >= 2 && it < 2 + NTRACE_BLK && (tid & 31) == 0;
        long long* t = trace + ((size_t)(it - 2) * 16 + (tid >> 5)) * NSTAMP;
        const long long nb = blk + gridDim.x;
        float2 v[32];
        if (tr) t[0] = clk();
        mbar_wait(mbar, parity); parity ^= 1;
        if (tr) t[1] = clk();
        phase_a_linear_load(tid, sm, v);
        if (tr) t[2] = clk();
        phase_a_linear_compute(tid, s_tw1, v);
        if (tr) t[3] = clk();
        __syncthreads();
        if (tr) t[4] = clk();
        phase_a_linear_store(tid, sm, v);
        if (tr) t[5] = clk();
        __syncthreads();
        if (tr) t[6] = clk();
        phase_mid_b(tid, s_tw2, sm); __syncwarp();
        if (tr) t[7] = clk();
        phase_mid_c(tid, Hp, s_hres, sm); __syncwarp();
        if (tr) t[8] = clk();
        phase_mid_bi(tid, s_tw2, sm);
        if (tr) t[9] = clk();
        __syncthreads();
        if (tr) t[10] = clk();
        phase_ai<false, false>(tid, blk, io, s_tw1, sm, NoTurn(), [&]() {
            if (tr) t[11] = clk();
            __syncthreads();
            if (tr) t[12] = clk();
            if (nb < nblocks && !early) stage(nb);
            if (tr) t[13] = clk();
        });
        if (tr) t[14] = clk();
    }
}

int main() {
    const size_t ntaps = 4097; const long long n = 1ll << 26;
    std::vector<float> taps(2 * ntaps, 0.f); for (size_t i = 0; i < ntaps; i++) taps[2*i] = 1.0f / ntaps;
    std::vector<float2> Hp, tw1, tw2; build_tables(taps.data(), ntaps, Hp, tw1, tw2);
    float2 *dH, *d1, *d2, *din, *dout, *dhist; long long* dtrace;
    CK(cudaMalloc(&dH, Hp.size()*8)); CK(cudaMemcpy(dH, Hp.data(), Hp.size()*8, cudaMemcpyHostToDevice));
    CK(cudaMalloc(&d1, 512*8)); CK(cudaMemcpy(d1, tw1.data(), 512*8, cudaMemcpyHostToDevice));
    CK(cudaMalloc(&d2, 512*8)); CK(cudaMemcpy(d2, tw2.data(), 512*8, cudaMemcpyHostToDevice));
    CK(cudaMalloc(&din, n*8)); CK(cudaMemset(din, 0, n*8)); CK(cudaMalloc(&dout, n*8)); CK(cudaMalloc(&dhist, ntaps*8)); CK(cudaMemset(dhist, 0, ntaps*8));
    CK(cudaMalloc(&dtrace, NTRACE_BLK*16*NSTAMP*8)); CK(cudaMemset(dtrace, 0, NTRACE_BLK*16*NSTAMP*8));
    BlockIO io; io.in = din; io.hist = dhist; io.out = dout; io.n_in = n; io.n_out = n; io.T1 = ntaps-1; io.V = N - io.T1; io.T1_total = ntaps-1; io.shift = 0; io.deci = 1; io.skip = 0;
    const long long nblocks = (n + io.V - 1) / io.V;
    CK(cudaFuncSetAttribute(trace_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int rep = 0; rep < 2; rep++) { cudaEventRecord(e0); trace_kernel<<<148, NT, SMEM>>>(io, dH, d1, d2, nblocks, dtrace, 0); cudaEventRecord(e1); CK(cudaEventSynchronize(e1)); }
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    printf("TMA-staged kernel %.3f ms for 2^26 samples (%lld blocks, %.0f cycles/block/SM at 1965 MHz)\n", ms, nblocks, ms*1e-3*1.965e9/(nblocks/148.0));
    std::vector<long long> tr(NTRACE_BLK*16*NSTAMP); CK(cudaMemcpy(tr.data(), dtrace, tr.size()*8, cudaMemcpyDeviceToHost));
    const char* names[] = {"wait mbarrier (TMA landed)", "A: LDS linear", "A: DFT32 + twiddle", "barrier Y", "A: STS padded", "barrier 1", "B", "C", "B'", "barrier 2",
                           "A': LDS", "barrier X", "stage (issue TMA)", "A': tw, IDFT32, STG"};
    for (int b = 0; b < NTRACE_BLK; b++) {
        long long t0 = tr[(b*16)*NSTAMP]; for (int w = 0; w < 16; w++) t0 = std::min(t0, tr[(b*16+w)*NSTAMP]);
        long long tend = 0; for (int w = 0; w < 16; w++) tend = std::max(tend, tr[(b*16+w)*NSTAMP+14]);
        printf("block iter %d: total %lld cycles\n", b + 2, tend - t0);
        for (int p = 0; p < 14; p++) {
            std::vector<long long> d; for (int w = 0; w < 16; w++) d.push_back(tr[(b*16+w)*NSTAMP+p+1] - tr[(b*16+w)*NSTAMP+p]);
            std::sort(d.begin(), d.end());
            printf("   %-28s min %6lld  med %6lld  max %6lld\n", names[p], d[0], d[8], d[15]);
        }
    }
    return 0;
}
