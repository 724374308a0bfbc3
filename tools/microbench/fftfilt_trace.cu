// Per-phase cycle trace of the FftFilter kernel body (same phase functions as fftfilt.cu), to see
// where a 16384-point block's ~28 K cycles go.  Every warp records clock64() at each phase boundary
// for a few blocks; the host prints min/median/max over warps per phase.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I../../rustradio_b200/csrc -o fftfilt_trace fftfilt_trace.cu
#include <algorithm>
#include <cstdio>
#include <vector>
#include "fftfilt_core.cuh"
#include "fftfilt_tables.hpp"
using namespace rrc::fftk;
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("cuda error %s line %d\n", cudaGetErrorString(e), __LINE__); return 1;}}while(0)
constexpr int NSTAMP = 8, NTRACE_BLK = 6;
constexpr size_t SMEM = (size_t)(SMEM_ELEMS + 512 + 512 + HRES_ELEMS) * sizeof(float2);

__global__ void __launch_bounds__(NT, 1)
trace_kernel(const BlockIO io, const float2* __restrict__ Hp, const float2* __restrict__ tw1g, const float2* __restrict__ tw2g,
             long long nblocks, long long* trace) {
    extern __shared__ __align__(16) float2 sm[];
    float2* s_tw2 = sm + SMEM_ELEMS; float2* s_tw1 = s_tw2 + 512; float2* s_hres = s_tw1 + 512;
    const int tid = threadIdx.x;
    s_tw2[tid] = tw2g[tid]; s_tw1[tid] = tw1g[tid];
    load_hres(tid, Hp, s_hres);
    __syncthreads();
    int it = 0;
    for (long long blk = blockIdx.x; blk < nblocks; blk += gridDim.x, ++it) {
        const bool tr = blockIdx.x == 3 && it >= 2 && it < 2 + NTRACE_BLK && (tid & 31) == 0;
        long long* t = trace + ((size_t)(it - 2) * 16 + (tid >> 5)) * NSTAMP;
        if (tr) t[0] = clock64();
        phase_a(tid, blk, io, s_tw1, sm);
        if (tr) t[1] = clock64();
        __syncthreads();
        if (tr) t[2] = clock64();
        phase_mid_b(tid, s_tw2, sm); __syncwarp();
        if (tr) t[3] = clock64();
        phase_mid_c(tid, Hp, s_hres, sm); __syncwarp();
        if (tr) t[4] = clock64();
        phase_mid_bi(tid, s_tw2, sm);
        if (tr) t[5] = clock64();
        __syncthreads();
        if (tr) t[6] = clock64();
        phase_ai<false, false>(tid, blk, io, s_tw1, sm);
        if (tr) t[7] = clock64();
    }
}

// ---- burst-level trace: stamps around every acquire()/release() of every phase ------------------
// MODE 0: plain kernel (gate only: the volatile load of `one`, no token); MODE 1: ping-pong token.
constexpr int NB_STAMP = 40, NB_BLK = 4;
__device__ __forceinline__ long long clk() { long long c; asm volatile("mov.u64 %0, %%clock64;" : "=l"(c) :: "memory"); return c; }
template <int MODE>
struct TraceTurn {
    PingPong pp; long long* t; int* k; bool on;
    __device__ __forceinline__ float acquire() const {
        if (on) t[(*k)++] = clk();
        float one;
        if (MODE == 1) one = pp.acquire();
        else asm volatile("ld.volatile.shared.f32 %0, [%1];" : "=f"(one) : "r"(pp.one_addr) : "memory");
        if (on) t[(*k)++] = clk();
        return one;
    }
    __device__ __forceinline__ void release() const {
        if (on) t[(*k)++] = clk();
        if (MODE == 1) pp.release();
    }
};
template <int MODE>
__global__ void __launch_bounds__(NT, 1)
burst_trace_kernel(const BlockIO io, const float2* __restrict__ Hp, const float2* __restrict__ tw1g, const float2* __restrict__ tw2g,
                   long long nblocks, long long* trace) {
    extern __shared__ __align__(16) float2 sm[];
    float2* s_tw2 = sm + SMEM_ELEMS; float2* s_tw1 = s_tw2 + 512; float2* s_hres = s_tw1 + 512;
    float* s_one = reinterpret_cast<float*>(s_hres + HRES_ELEMS);
    const int tid = threadIdx.x;
    s_tw2[tid] = tw2g[tid]; s_tw1[tid] = tw1g[tid];
    load_hres(tid, Hp, s_hres);
    if (tid == 0) *s_one = 1.0f;
    __syncthreads();
    const PingPong pp{(tid >> 7) & 1, (unsigned)__cvta_generic_to_shared(s_one)};
    if (MODE == 1 && pp.g == 1) pp.release();
    int it = 0;
    for (long long blk = blockIdx.x; blk < nblocks; blk += gridDim.x, ++it) {
        const bool tr = blockIdx.x == 3 && it >= 2 && it < 2 + NB_BLK && (tid & 31) == 0;
        long long* t = trace + ((size_t)(it - 2) * 16 + (tid >> 5)) * NB_STAMP;
        int k = 0;
        const TraceTurn<MODE> turn{pp, t, &k, tr};
        if (tr) t[k++] = clk();
        phase_a(tid, blk, io, s_tw1, sm, turn);
        if (tr) t[k++] = clk();
        __syncthreads();
        if (tr) t[k++] = clk();
        phase_mid(tid, s_tw2, Hp, s_hres, sm, turn);
        if (tr) t[k++] = clk();
        __syncthreads();
        if (tr) t[k++] = clk();
        phase_ai<false, false>(tid, blk, io, s_tw1, sm, turn);
        if (tr) t[k++] = clk();
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
    for (int rep = 0; rep < 2; rep++) { cudaEventRecord(e0); trace_kernel<<<148, NT, SMEM>>>(io, dH, d1, d2, nblocks, dtrace); cudaEventRecord(e1); CK(cudaEventSynchronize(e1)); }
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    printf("kernel %.3f ms for 2^26 samples (%lld blocks, %.0f cycles/block/SM at 1965 MHz)\n", ms, nblocks, ms*1e-3*1.965e9/(nblocks/148.0));
    std::vector<long long> tr(NTRACE_BLK*16*NSTAMP); CK(cudaMemcpy(tr.data(), dtrace, tr.size()*8, cudaMemcpyDeviceToHost));
    const char* names[] = {"A (load,DFT32,tw,STS)", "wait barrier 1", "B", "C", "B'", "wait barrier 2", "A' (LDS,tw,IDFT32,STG)"};
    for (int b = 0; b < NTRACE_BLK; b++) {
        long long t0 = tr[(b*16)*NSTAMP]; for (int w = 0; w < 16; w++) t0 = std::min(t0, tr[(b*16+w)*NSTAMP]);
        long long tend = 0; for (int w = 0; w < 16; w++) tend = std::max(tend, tr[(b*16+w)*NSTAMP+7]);
        printf("block iter %d: total %lld cycles\n", b + 2, tend - t0);
        for (int p = 0; p < 7; p++) {
            std::vector<long long> d; for (int w = 0; w < 16; w++) d.push_back(tr[(b*16+w)*NSTAMP+p+1] - tr[(b*16+w)*NSTAMP+p]);
            std::sort(d.begin(), d.end());
            printf("   %-28s min %6lld  med %6lld  max %6lld\n", names[p], d[0], d[8], d[15]);
        }
        printf("   phase-A start skew over warps: ");
        for (int w = 0; w < 16; w++) printf("%lld ", tr[(b*16+w)*NSTAMP] - t0);
        printf("\n");
    }
    // ---- burst-level traces ----
    long long* dbt; CK(cudaMalloc(&dbt, NB_BLK*16*NB_STAMP*8));
    for (int mode = 0; mode < 2; mode++) {
        CK(cudaMemset(dbt, 0, NB_BLK*16*NB_STAMP*8));
        auto kern = mode == 0 ? burst_trace_kernel<0> : burst_trace_kernel<1>;
        CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM + 16));
        for (int rep = 0; rep < 2; rep++) { cudaEventRecord(e0); kern<<<148, NT, SMEM + 16>>>(io, dH, d1, d2, nblocks, dbt); cudaEventRecord(e1); CK(cudaEventSynchronize(e1)); }
        cudaEventElapsedTime(&ms, e0, e1);
        printf("\n=== burst trace, %s: kernel %.3f ms (%.0f cycles/block/SM)\n", mode ? "PING-PONG" : "plain (gated)", ms, ms*1e-3*1.965e9/(nblocks/148.0));
        std::vector<long long> bt(NB_BLK*16*NB_STAMP); CK(cudaMemcpy(bt.data(), dbt, bt.size()*8, cudaMemcpyDeviceToHost));
        // stamp layout per block: 0 A.begin | 1 acqA- 2 acqA+ 3 relA | 4 A.end 5 bar1 | B: 6 7 8 | C0: 9 10 11 | C1: 12 13 14 | B': 15 16 17 | 18 mid.end 19 bar2 | A': 20 21 22 | 23 end
        const char* bn[] = {"A", "B", "C0", "C1", "B'", "A'"};
        const int acq[] = {1, 6, 9, 12, 15, 20};
        for (int b = 1; b < 3; b++) {
            long long t0 = bt[(b*16)*NB_STAMP]; for (int w = 0; w < 16; w++) t0 = std::min(t0, bt[(b*16+w)*NB_STAMP]);
            printf("block iter %d (times relative to first warp's A.begin)\n", b + 2);
            for (int w = 0; w < 16; w++) {
                const long long* t = &bt[(b*16+w)*NB_STAMP];
                printf(" w%02d g%d begin %6lld |", w, (w >> 2) & 1, t[0] - t0);
                for (int q = 0; q < 6; q++) printf(" %s pre %5lld wait %5lld fp %5lld |", bn[q], t[acq[q]] - t[acq[q] - 1 - (q == 1 || q == 5 ? 0 : 0)], t[acq[q]+1] - t[acq[q]], t[acq[q]+2] - t[acq[q]+1]);
                printf(" bar1 %5lld bar2 %5lld end %6lld\n", t[5] - t[4], t[19] - t[18], t[23] - t0);
            }
        }
    }
    return 0;
}
