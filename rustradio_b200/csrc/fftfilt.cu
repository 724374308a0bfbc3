// fftfilt.cu — FftFilter (FFT block convolution) for Complex<f32> on sm_100a.
//
// Replaces rustradio's RustFftEngine::new / Engine::run / sum_vec and the
// overlap handling of FftFilter::work (src/fft_filter.rs:144-176, 281-287,
// 331-348).  The reference does overlap-ADD with F = 2*nextpow2(ntaps)
// (SURVEY F1); this kernel computes the same linear convolution
//   y[n] = sum_k h[k] x[n-k],  x[n<0] = 0 (src/fft_filter.rs:270)
// by overlap-SAVE with a fixed 16384-point transform held entirely in one
// CTA's shared memory: one HBM read of the input, one HBM write of the
// output, forward FFT, spectrum multiply and inverse FFT fused in between
// (fftfilt_core.cuh).  State carried across calls: the last ntaps-1 inputs.
//
// Roofline: per 16384-point block, V = 16384-(ntaps-1) outputs; bytes
// 8*(N_in + N_out); ~116 FP32 instructions per transformed point (issue bound
// before HBM bound for ntaps > ~2048).
#include <algorithm>
#include <cmath>
#include <complex>
#include <cstdlib>
#include <type_traits>
#include <vector>

#include "common.cuh"
#include "fftfilt_core.cuh"
#include "fftfilt_tables.hpp"
#include "pipeline.cuh"
#include "fftfilt_handle.hpp"
#include "tmem.cuh"

namespace rrc {

using fftk::BlockIO;

constexpr size_t FFTFILT_SMEM = (size_t)(fftk::SMEM_ELEMS + 512 + 512 + fftk::HRES_ELEMS + 2) * sizeof(float2);

// Input prefetch: pulls the input segment of block `nb` into L2 with one 8 KiB bulk prefetch per warp
// (16 x 8 KiB = segment; SASS UBLKPF.L2).
__device__ __forceinline__ void prefetch_segment(const BlockIO& io, long long nb, long long nblocks, int tid) {
    if (io.real) {      // two real segments of N floats, V apart: warps 0-7 the first, 8-15 the second, 8 KiB each
        const int w = tid >> 5;
        const long long seg0 = 2 * nb * (long long)io.V - io.T1 - io.shift + (w >> 3) * (long long)io.V + (long long)(w & 7) * 2048;
        if ((tid & 31) == 0 && nb < nblocks && seg0 >= 0 && seg0 + 2048 <= io.n_in) {
            const unsigned long long a = (reinterpret_cast<unsigned long long>(io.in) + (unsigned long long)seg0 * 4 + 15ull) & ~15ull;
            asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(a), "r"(8192 - 16) : "memory");
        }
        return;
    }
    const long long seg0 = nb * (long long)io.V - io.T1 - io.shift + (long long)(tid >> 5) * 1024;
    if ((tid & 31) == 0 && nb < nblocks && seg0 >= 0 && seg0 + 1024 <= io.n_in) {
        const unsigned long long esz = io.in_u8 ? 2 : 8;         // bytes per input sample
        const unsigned long long a = (reinterpret_cast<unsigned long long>(io.in) + (unsigned long long)seg0 * esz + 15ull) & ~15ull;
        asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(a), "r"((unsigned)(1024 * esz - 16)) : "memory");
    }
}

// tune: bits 0-3 = prefetch placement (0 none, 1 after barrier 1 for the next block, 2 at block start
// for the next block, 3 at block start for the block after next); bits 8.. = CTA start stagger in
// units of 1024 cycles times (blockIdx & 3) (experiment knobs, RRC_FFTFILT_TUNE).
template <bool DECIM, bool ACCUM, bool TWC>
__global__ void __launch_bounds__(fftk::NT, 1)
fftfilt_kernel(const BlockIO io, const float2* __restrict__ Hp, const float2* __restrict__ tw1g,
               const float2* __restrict__ tw2g, long long nblocks, int tune) {
    extern __shared__ __align__(16) float2 sm[];
    float2* s_tw2 = sm + fftk::SMEM_ELEMS;
    float2* s_tw1 = s_tw2 + 512;
    float2* s_hres = s_tw1 + 512;
    const int tid = threadIdx.x;
    s_tw2[tid] = tw2g[tid];
    s_tw1[tid] = tw1g[tid];
    fftk::load_hres(tid, Hp, s_hres);
    const int pf = tune & 15;
    if (io.hist_next && blockIdx.x == gridDim.x - 1) fftk::update_history(io, tid, fftk::NT);
    if (tune >> 8) {
        const long long t0 = clock64(), wait = (long long)(tune >> 8) * 1024 * (blockIdx.x & 3);
        while (clock64() - t0 < wait) { }
    }
    __syncthreads();
    if (pf == 3) prefetch_segment(io, blockIdx.x + gridDim.x, nblocks, tid);
    for (long long blk = blockIdx.x; blk < nblocks; blk += gridDim.x) {
        if (pf == 2) prefetch_segment(io, blk + gridDim.x, nblocks, tid);
        if (pf == 3) prefetch_segment(io, blk + 2 * gridDim.x, nblocks, tid);
        fftk::phase_a(tid, blk, io, s_tw1, sm);
        __syncthreads();
        if (pf == 1) prefetch_segment(io, blk + gridDim.x, nblocks, tid);
        fftk::phase_mid<TWC>(tid, s_tw2, Hp, s_hres, sm); // half-warp local exchanges: __syncwarp only
        __syncthreads();
        fftk::phase_ai<DECIM, ACCUM>(tid, blk, io, s_tw1, sm);
        // no barrier: the next phase_a writes exactly the words this thread just read
    }
}

// Staged-input variant: the next block's input is copied into the exchange buffer by cp.async while
// phase A' of the current block computes and stores (fftfilt_core.cuh, stage_input).
template <bool DECIM, bool ACCUM>
__global__ void __launch_bounds__(fftk::NT, 1)
fftfilt_st_kernel(const BlockIO io, const float2* __restrict__ Hp, const float2* __restrict__ tw1g,
                  const float2* __restrict__ tw2g, long long nblocks, int tune) {
    extern __shared__ __align__(16) float2 sm[];
    float2* s_tw2 = sm + fftk::SMEM_ELEMS;
    float2* s_tw1 = s_tw2 + 512;
    float2* s_hres = s_tw1 + 512;
    const int tid = threadIdx.x;
    s_tw2[tid] = tw2g[tid];
    s_tw1[tid] = tw1g[tid];
    fftk::load_hres(tid, Hp, s_hres);
    const int pf = tune & 15;
    if (blockIdx.x < nblocks) fftk::stage_input(tid, blockIdx.x, io, sm);
    __syncthreads();
    for (long long blk = blockIdx.x; blk < nblocks; blk += gridDim.x) {
        fftk::phase_a_staged(tid, s_tw1, sm);
        __syncthreads();
        if (pf) prefetch_segment(io, blk + pf * gridDim.x, nblocks, tid);
        fftk::phase_mid(tid, s_tw2, Hp, s_hres, sm);
        __syncthreads();
        const long long nb = blk + gridDim.x;
        fftk::phase_ai<DECIM, ACCUM>(tid, blk, io, s_tw1, sm, fftk::NoTurn(), [&]() {
            __syncthreads();                                  // every thread has read the buffer
            if (nb < nblocks) fftk::stage_input(tid, nb, io, sm);
        });
    }
}

// TMA-staged variant (variant 36; fftfilt_core.cuh "linear staging"): the next block's input is
// copied into the exchange buffer by cp.async.bulk (SASS UBLKCP) while phase A' of the current block
// computes and stores; completion is tracked by one mbarrier (one phase per block).
__device__ __forceinline__ void mbar_wait(unsigned mbar, unsigned parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "MBAR_WAIT:\n"
        "mbarrier.try_wait.parity.shared.b64 p, [%0], %1;\n"
        "@p bra MBAR_DONE;\n"
        "bra MBAR_WAIT;\n"
        "MBAR_DONE:\n"
        "}" ::"r"(mbar), "r"(parity) : "memory");
}

template <bool DECIM, bool ACCUM>
__global__ void __launch_bounds__(fftk::NT, 1)
fftfilt_tma_kernel(const BlockIO io, const float2* __restrict__ Hp, const float2* __restrict__ tw1g,
                   const float2* __restrict__ tw2g, long long nblocks, int tune) {
    extern __shared__ __align__(16) float2 sm[];
    float2* s_tw2 = sm + fftk::SMEM_ELEMS;
    float2* s_tw1 = s_tw2 + 512;
    float2* s_hres = s_tw1 + 512;
    const int tid = threadIdx.x;
    s_tw2[tid] = tw2g[tid];
    s_tw1[tid] = tw1g[tid];
    fftk::load_hres(tid, Hp, s_hres);
    const unsigned mbar = (unsigned)__cvta_generic_to_shared(s_hres + fftk::HRES_ELEMS);
    const unsigned sm_a = (unsigned)__cvta_generic_to_shared(sm);
    const int pf = tune & 15;
    if (io.hist_next && blockIdx.x == gridDim.x - 1) fftk::update_history(io, tid, fftk::NT);
    if (tid == 0) {
        asm volatile("mbarrier.init.shared.b64 [%0], 1;" ::"r"(mbar) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    // Every thread calls stage() (the test is CTA uniform); the caller guarantees that nobody still
    // reads the exchange buffer.
    auto stage = [&](long long nb) {
        if (fftk::stage_linear_bulk_ok(nb, io)) {
            if (tid == 0) {
                const float2* src = io.in + fftk::stage_linear_seg0(nb, io);
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // earlier generic accesses to the buffer before the async-proxy writes
                asm volatile("mbarrier.arrive.expect_tx.shared.b64 _, [%0], %1;" ::"r"(mbar), "r"(fftk::N * 8) : "memory");
#pragma unroll 1
                for (int c = 0; c < 8; ++c)
                    asm volatile("cp.async.bulk.shared::cta.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                                 ::"r"(sm_a + c * 16384), "l"(src + c * 2048), "r"(16384), "r"(mbar) : "memory");
            }
        } else {
            fftk::stage_linear_fallback(tid, nb, io, sm);
            if (tid == 0) asm volatile("mbarrier.arrive.shared.b64 _, [%0];" ::"r"(mbar) : "memory");
        }
    };
    if (tune >> 8) {      // start stagger: SMs that start together stay in lock-step and collide on HBM
        const long long t0 = clock64(), wait = (long long)(tune >> 8) * 1024 * (blockIdx.x & 7);
        while (clock64() - t0 < wait) { }
    }
    if (blockIdx.x < nblocks) stage(blockIdx.x);
    unsigned parity = 0;
    for (long long blk = blockIdx.x; blk < nblocks; blk += gridDim.x) {
        const long long nb = blk + gridDim.x;
        float2 v[32];
        mbar_wait(mbar, parity);
        parity ^= 1;
        fftk::phase_a_linear_load(tid, sm, v);
        fftk::phase_a_linear_compute(tid, s_tw1, v);
        __syncthreads();                                      // every thread has read its linear words
        fftk::phase_a_linear_store(tid, sm, v);
        __syncthreads();
        if (pf) prefetch_segment(io, nb, nblocks, tid);       // next block -> L2, so the bulk copy below is an L2 hit
        fftk::phase_mid<false>(tid, s_tw2, Hp, s_hres, sm);
        __syncthreads();
        fftk::phase_ai<DECIM, ACCUM>(tid, blk, io, s_tw1, sm, fftk::NoTurn(), [&]() {
            __syncthreads();                                  // every thread has read the buffer
            if (nb < nblocks) stage(nb);
        });
    }
}

// Variant 40: fftfilt_tma_kernel with the WHOLE spectrum in tensor memory.  Thread tid multiplies rows k2 = l, l + 16 of plane
// k1 in every block: 32 spectrum values = 64 columns of its own TMEM lane, written once at kernel start (tcgen05.st), read at
// the top of each half of phase C (tcgen05.ld.32x32b.x32).  No spectrum rows in shared memory (72 KiB less), no 64 KiB per
// block from L2 through the 28 KiB of L1 that 216 KiB of shared memory leave, 8 LDS.128 per thread and block fewer.
constexpr size_t FFTFILT_TMH_SMEM = (size_t)(fftk::SMEM_ELEMS + 512 + 512 + 4) * sizeof(float2);

__device__ __forceinline__ void phase_mid_c_tm(int tid, unsigned hbase, float2* sm) {
    const int k1 = tid >> 4, l = tid & 15;
#pragma unroll
    for (int half = 0; half < 2; ++half) {
        float2* row = sm + k1 * fftk::PLANE_PITCH + (l + 16 * half) * fftk::ROW_PITCH;
        float2 hh[16];
        tm_ld32_issue(hbase + 32u * half, hh);                  // in flight behind the row loads and the forward DFT16
        float2 v[16];
#pragma unroll
        for (int n3 = 0; n3 < 16; ++n3) v[fftr::bitrev(n3, 4)] = row[n3];
        fftr::dit<16, +1>(v);
        tm_ld32_wait(hh);
        float2 u[16];
#pragma unroll
        for (int k3 = 0; k3 < 16; ++k3) u[fftr::bitrev(k3, 4)] = fftr::cmul(v[k3], hh[k3]);
        fftr::dit<16, -1>(u);
#pragma unroll
        for (int n3 = 0; n3 < 16; ++n3) row[n3] = u[n3];
    }
}

// TWP (variant 41): the 32 powers W_N^{t k1} of phases A and A' also live in the thread's TMEM strip (columns 256..), written once,
// instead of being rebuilt from tw1[tid] with 31 complex multiplications twice per block.  Same values, bit-identical output.
__device__ __forceinline__ void phase_a_compute_tm(unsigned pbase, float2 (&v)[32]) {
    fftr::dit<32, +1>(v);
#pragma unroll
    for (int half = 0; half < 2; ++half) {
        float2 p[16];
        tm_ld32(pbase + 32u * half, p);
#pragma unroll
        for (int j = 0; j < 16; ++j) v[16 * half + j] = fftr::cmul(v[16 * half + j], p[j]);
    }
}
template <bool DECIM, bool ACCUM, class AfterLoad>
__device__ __forceinline__ void phase_ai_tm(int tid, long long blk, const BlockIO& io, unsigned pbase, const float2* sm, AfterLoad after_load) {
    const float2* s = sm + (tid >> 4) * fftk::ROW_PITCH + (tid & 15);
    float2 v[32];
#pragma unroll
    for (int k1 = 0; k1 < 32; ++k1) v[fftr::bitrev(k1, 5)] = s[k1 * fftk::PLANE_PITCH];
    after_load();
#pragma unroll
    for (int half = 0; half < 2; ++half) {
        float2 p[16];
        tm_ld32(pbase + 32u * half, p);
#pragma unroll
        for (int j = 0; j < 16; ++j) v[fftr::bitrev(16 * half + j, 5)] = fftr::cmul_conj(v[fftr::bitrev(16 * half + j, 5)], p[j]);
    }
    fftr::dit<32, -1>(v);
    fftk::store_outputs<DECIM, ACCUM>(tid, blk, io, v);
}

// TWB (variant 42): instead of the phase-A powers, the 32 twiddles W_512^{l k2} of phases B and B' (64 LDS.64 per thread and
// block out of ~330 shared-memory instructions) live in columns 256.. of the strip.
__device__ __forceinline__ void phase_mid_b_tm(int tid, unsigned wbase, float2* sm) {
    const int k1 = tid >> 4, l = tid & 15;
    float2* col = sm + k1 * fftk::PLANE_PITCH + l;
    float2 v[32];
#pragma unroll
    for (int n2 = 0; n2 < 32; ++n2) v[fftr::bitrev(n2, 5)] = col[n2 * fftk::ROW_PITCH];
    fftr::dit<32, +1>(v);
#pragma unroll
    for (int half = 0; half < 2; ++half) {
        float2 w[16];
        tm_ld32(wbase + 32u * half, w);
#pragma unroll
        for (int j = 0; j < 16; ++j) v[16 * half + j] = fftr::cmul(v[16 * half + j], w[j]);
    }
#pragma unroll
    for (int k2 = 0; k2 < 32; ++k2) col[k2 * fftk::ROW_PITCH] = v[k2];
}
__device__ __forceinline__ void phase_mid_bi_tm(int tid, unsigned wbase, float2* sm) {
    const int k1 = tid >> 4, l = tid & 15;
    float2* col = sm + k1 * fftk::PLANE_PITCH + l;
    float2 v[32];
#pragma unroll
    for (int k2 = 0; k2 < 32; ++k2) v[fftr::bitrev(k2, 5)] = col[k2 * fftk::ROW_PITCH];
#pragma unroll
    for (int half = 0; half < 2; ++half) {
        float2 w[16];
        tm_ld32(wbase + 32u * half, w);
#pragma unroll
        for (int j = 0; j < 16; ++j) v[fftr::bitrev(16 * half + j, 5)] = fftr::cmul_conj(v[fftr::bitrev(16 * half + j, 5)], w[j]);
    }
    fftr::dit<32, -1>(v);
#pragma unroll
    for (int n2 = 0; n2 < 32; ++n2) col[n2 * fftk::ROW_PITCH] = v[n2];
}

template <bool DECIM, bool ACCUM, bool TWP, bool TWB = false>
__global__ void __launch_bounds__(fftk::NT, 1)
fftfilt_tmh_kernel(const BlockIO io, const float2* __restrict__ Hp, const float2* __restrict__ tw1g,
                   const float2* __restrict__ tw2g, long long nblocks, int tune) {
    extern __shared__ __align__(16) float2 sm[];
    float2* s_tw2 = sm + fftk::SMEM_ELEMS;
    float2* s_tw1 = s_tw2 + 512;
    unsigned* s_tmem = reinterpret_cast<unsigned*>(s_tw1 + 512 + 2);
    const int tid = threadIdx.x, warp = tid >> 5;
    s_tw2[tid] = tw2g[tid];
    s_tw1[tid] = tw1g[tid];
    const unsigned mbar = (unsigned)__cvta_generic_to_shared(s_tw1 + 512);
    const unsigned sm_a = (unsigned)__cvta_generic_to_shared(sm);
    const int pf = tune & 15;
    if (io.hist_next && blockIdx.x == gridDim.x - 1) fftk::update_history(io, tid, fftk::NT);
    if (warp == 0) tm_alloc_all(s_tmem);
    if (tid == 0) {
        asm volatile("mbarrier.init.shared.b64 [%0], 1;" ::"r"(mbar) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    tm_fence_before();
    __syncthreads();
    tm_fence_after();
    const unsigned tmem = *s_tmem;
    const unsigned hbase = tm_strip64(tmem, warp);
    {   // this thread's two spectrum rows -> its TMEM strip
        const int k1 = tid >> 4, l = tid & 15;
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            const float4* hp = reinterpret_cast<const float4*>(Hp + (size_t)(k1 * 32 + l + 16 * half) * 16);
            float2 x[16];
#pragma unroll
            for (int i = 0; i < 8; ++i) { const float4 h4 = hp[i]; x[2 * i] = make_float2(h4.x, h4.y); x[2 * i + 1] = make_float2(h4.z, h4.w); }
            tm_st32(hbase + 32u * half, x);
        }
        if constexpr (TWB) {
            float2 p[32];
#pragma unroll
            for (int k2 = 0; k2 < 32; ++k2) p[k2] = tw2g[k2 * 16 + (tid & 15)];
            tm_st32(hbase + 256u, *reinterpret_cast<float2(*)[16]>(&p[0]));
            tm_st32(hbase + 288u, *reinterpret_cast<float2(*)[16]>(&p[16]));
        }
        if constexpr (TWP) {
            float2 p[32];
            fftk::powers32(s_tw1[tid], p);
            tm_st32(hbase + 256u, *reinterpret_cast<float2(*)[16]>(&p[0]));
            tm_st32(hbase + 288u, *reinterpret_cast<float2(*)[16]>(&p[16]));
        }
        tm_wait_st();
    }
    auto stage = [&](long long nb) {
        if (fftk::stage_linear_bulk_ok(nb, io)) {
            if (tid == 0) {
                const float2* src = io.in + fftk::stage_linear_seg0(nb, io);
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                asm volatile("mbarrier.arrive.expect_tx.shared.b64 _, [%0], %1;" ::"r"(mbar), "r"(fftk::N * 8) : "memory");
#pragma unroll 1
                for (int c = 0; c < 8; ++c)
                    asm volatile("cp.async.bulk.shared::cta.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                                 ::"r"(sm_a + c * 16384), "l"(src + c * 2048), "r"(16384), "r"(mbar) : "memory");
            }
        } else {
            fftk::stage_linear_fallback(tid, nb, io, sm);
            if (tid == 0) asm volatile("mbarrier.arrive.shared.b64 _, [%0];" ::"r"(mbar) : "memory");
        }
    };
    if (blockIdx.x < nblocks) stage(blockIdx.x);
    unsigned parity = 0;
    for (long long blk = blockIdx.x; blk < nblocks; blk += gridDim.x) {
        const long long nb = blk + gridDim.x;
        float2 v[32];
        mbar_wait(mbar, parity);
        parity ^= 1;
        fftk::phase_a_linear_load(tid, sm, v);
        if constexpr (TWP) phase_a_compute_tm(hbase + 256u, v);
        else fftk::phase_a_linear_compute(tid, s_tw1, v);
        __syncthreads();
        fftk::phase_a_linear_store(tid, sm, v);
        __syncthreads();
        if (pf) prefetch_segment(io, nb, nblocks, tid);
        if constexpr (TWB) phase_mid_b_tm(tid, hbase + 256u, sm); else fftk::phase_mid_b<true>(tid, s_tw2, sm);
        __syncwarp();
        phase_mid_c_tm(tid, hbase, sm);
        __syncwarp();
        if constexpr (TWB) phase_mid_bi_tm(tid, hbase + 256u, sm); else fftk::phase_mid_bi<true>(tid, s_tw2, sm);
        __syncthreads();
        auto after = [&]() {
            __syncthreads();
            if (nb < nblocks) stage(nb);
        };
        if constexpr (TWP) phase_ai_tm<DECIM, ACCUM>(tid, blk, io, hbase + 256u, sm, after);
        else fftk::phase_ai<DECIM, ACCUM>(tid, blk, io, s_tw1, sm, fftk::NoTurn(), after);
    }
    tm_fence_before();
    __syncthreads();
    if (warp == 0) tm_dealloc_all(tmem);
}

// PACKED kernel (variant 37; fftfilt_pk.cuh): the same 16384-point block with FFMA2 / FADD2 / FMUL2 lanes,
// pair-word exchange layouts, TMA-staged input (32 bulk copies of 4 KiB, one per 512-sample row, into the
// plane-pitched landing layout L0) and three CTA barriers per block.
constexpr int PK_NSTAMP = 12, PK_TRACE_BLK = 6;
constexpr size_t FFTFILT_PK_SMEM = (size_t)(fftp::SMEM_WORDS + 512 + 512 + fftp::HRES_WORDS + 2) * sizeof(float2);

// MODE: 0 = plain, 1 = FP-turn ping-pong between two warp groups, 2 = staggered first load bursts (fftfilt_pk.cuh)
template <bool DECIM, bool ACCUM, int MODE>
__global__ void __launch_bounds__(fftk::NT, 1)
fftfilt_pk_kernel(const BlockIO io, const float2* __restrict__ Hq, const float2* __restrict__ tw1g,
                  const float2* __restrict__ tw2pg, long long nblocks, int tune, long long* __restrict__ trace) {
    extern __shared__ __align__(16) float2 sm[];
    float2* s_tw2p = sm + fftp::SMEM_WORDS;
    float2* s_tw1 = s_tw2p + 512;
    float2* s_hres = s_tw1 + 512;
    const int tid = threadIdx.x;
    s_tw2p[tid] = tw2pg[tid];
    s_tw1[tid] = tw1g[tid];
    fftp::load_hres(tid, Hq, s_hres);
    const unsigned mbar = (unsigned)__cvta_generic_to_shared(s_hres + fftp::HRES_WORDS);
    const unsigned sm_a = (unsigned)__cvta_generic_to_shared(sm);
    const int pf = tune & 15;
    if (io.hist_next && blockIdx.x == gridDim.x - 1) fftk::update_history(io, tid, fftk::NT);
    if (tid == 0) {
        asm volatile("mbarrier.init.shared.b64 [%0], 1;" ::"r"(mbar) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    // Every thread calls stage() (the test is CTA uniform); the caller guarantees nobody still reads the buffer.
    auto stage = [&](long long nb) {
        if (fftp::bulk_ok(nb, io)) {
            if (tid == 0) {
                const float2* src = io.in + fftp::seg0_of(nb, io);
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                asm volatile("mbarrier.arrive.expect_tx.shared.b64 _, [%0], %1;" ::"r"(mbar), "r"(fftk::N * 8) : "memory");
#pragma unroll 1
                for (int n1 = 0; n1 < 32; ++n1)
                    asm volatile("cp.async.bulk.shared::cta.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                                 ::"r"(sm_a + n1 * (fftp::PP * 8)), "l"(src + n1 * 512), "r"(4096), "r"(mbar) : "memory");
            }
        } else {
            fftp::stage_fallback(tid, nb, io, sm);
            __syncthreads();                                  // the arrival below must follow every thread's stores
            if (tid == 0) asm volatile("mbarrier.arrive.shared.b64 _, [%0];" ::"r"(mbar) : "memory");
        }
    };
    if (blockIdx.x < nblocks) stage(blockIdx.x);
    unsigned parity = 0;
    // FP-turn groups: warps 0-3 and 8-11 are group 0, 4-7 and 12-15 group 1 (two warps of each per sub-partition)
    using Turn = typename std::conditional<MODE == 1, fftp::PingPong, typename std::conditional<MODE == 2, fftp::Stagger, fftp::NoTurn>::type>::type;
    Turn turn;
    if constexpr (MODE == 1) {
        turn.g = (tid >> 7) & 1;
        if (turn.g == 1) turn.release();                      // group 0 takes the first turn
    }
    if constexpr (MODE == 2) turn.delay = (tid >> 5) * (tune >> 16 ? (tune >> 16) : 128);
    // RRC_FFTFILT_TRACE (debug): every warp of CTA 3 stamps clock64 at the phase boundaries of its blocks 2..7
    int it = 0;
    auto stamp = [&](int i) {
        if (trace && blockIdx.x == 3 && it >= 2 && it < 2 + PK_TRACE_BLK && (tid & 31) == 0)
            trace[((size_t)(it - 2) * 16 + (tid >> 5)) * PK_NSTAMP + i] = clock64();
    };
    for (long long blk = blockIdx.x; blk < nblocks; blk += gridDim.x, ++it) {
        const long long nb = blk + gridDim.x;
        stamp(0);
        mbar_wait(mbar, parity);
        parity ^= 1;
        stamp(1);
        fftp::phase_a(tid, s_tw1, sm, turn);                  // in place pairwise: __syncwarp only
        stamp(2);
        __syncthreads();
        stamp(3);
        if (pf) prefetch_segment(io, nb, nblocks, tid);       // next block -> L2, so the bulk copies are L2 hits
        fftp::phase_b(tid, s_tw2p, sm, turn);
        __syncwarp();
        stamp(4);
        fftp::phase_c(tid, Hq, s_hres, sm, turn);
        __syncwarp();
        stamp(5);
        fftp::phase_bi(tid, s_tw2p, sm, turn);
        stamp(6);
        __syncthreads();
        stamp(7);
        fftp::phase_ai<DECIM, ACCUM>(tid, blk, io, s_tw1, sm, [&]() {
            stamp(8);
            __syncthreads();                                  // every thread has read the buffer
            if (nb < nblocks) stage(nb);
            stamp(9);
        }, turn);
        stamp(10);
    }
}

// Ping-pong variant of fftfilt_kernel (PingPong policy: fftfilt_core.cuh).
template <bool DECIM, bool ACCUM>
__global__ void __launch_bounds__(fftk::NT, 1)
fftfilt_pp_kernel(const BlockIO io, const float2* __restrict__ Hp, const float2* __restrict__ tw1g,
                  const float2* __restrict__ tw2g, long long nblocks, int /*tune*/) {
    extern __shared__ __align__(16) float2 sm[];
    float2* s_tw2 = sm + fftk::SMEM_ELEMS;
    float2* s_tw1 = s_tw2 + 512;
    float2* s_hres = s_tw1 + 512;
    const int tid = threadIdx.x;
    s_tw2[tid] = tw2g[tid];
    s_tw1[tid] = tw1g[tid];
    fftk::load_hres(tid, Hp, s_hres);
    float* s_one = reinterpret_cast<float*>(s_hres + fftk::HRES_ELEMS);
    if (tid == 0) *s_one = 1.0f;
    __syncthreads();
    const fftk::PingPong turn{(tid >> 7) & 1, (unsigned)__cvta_generic_to_shared(s_one)};
    if (turn.g == 1) turn.release();                       // group 0 takes the first turn
    for (long long blk = blockIdx.x; blk < nblocks; blk += gridDim.x) {
        fftk::phase_a(tid, blk, io, s_tw1, sm, turn);
        __syncthreads();
        {
            const long long nb = blk + gridDim.x;
            const long long seg0 = nb * (long long)io.V - io.T1 - io.shift + (long long)(tid >> 5) * 1024;
            if ((tid & 31) == 0 && nb < nblocks && seg0 >= 0 && seg0 + 1024 <= io.n_in) {
                const unsigned long long a = (reinterpret_cast<unsigned long long>(io.in + seg0) + 15ull) & ~15ull;
                asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(a), "r"(8192 - 16) : "memory");
            }
        }
        fftk::phase_mid(tid, s_tw2, Hp, s_hres, sm, turn);
        __syncthreads();
        fftk::phase_ai<DECIM, ACCUM>(tid, blk, io, s_tw1, sm, turn);
    }
}

// 1024-thread x 16-point variant (fftfilt16_core.cuh).  Plane-local exchanges (B<->C<->D) use
// 128-thread named barriers (two k1 planes per barrier), ids 1..8; id 0 is __syncthreads.
constexpr size_t FFTFILT16_SMEM = (size_t)(fftk16::SMEM16_ELEMS + 1024 + 1024 + 64 + fftk16::HRES16_ELEMS) * sizeof(float2);

__device__ __forceinline__ void plane_barrier(int tid) {
    asm volatile("bar.sync %0, 128;" ::"r"(1 + (tid >> 7)) : "memory");
}

template <bool DECIM, bool ACCUM>
__global__ void __launch_bounds__(fftk16::NT16, 1)
fftfilt16_kernel(const BlockIO io, const float2* __restrict__ Hd, const float2* __restrict__ tw1g,
                 const float2* __restrict__ tw2g, const float2* __restrict__ tw3g, long long nblocks) {
    extern __shared__ __align__(16) float2 sm[];
    float2* s_tw1 = sm + fftk16::SMEM16_ELEMS;
    float2* s_tw2 = s_tw1 + 1024;
    float2* s_tw3 = s_tw2 + 1024;
    float2* s_hres = s_tw3 + 64;
    const int tid = threadIdx.x;
    s_tw1[tid] = tw1g[tid];
    s_tw2[tid] = tw2g[tid];
    if (tid < 64) s_tw3[tid] = tw3g[tid];
    fftk16::load_hres(tid, Hd, s_hres);
    __syncthreads();
    for (long long blk = blockIdx.x; blk < nblocks; blk += gridDim.x) {
        fftk16::phase_a(tid, blk, io, s_tw1, sm);
        __syncthreads();
        {   // next block's input -> L2 (one 4 KiB bulk prefetch per warp, 32 x 4 KiB = segment)
            const long long nb = blk + gridDim.x;
            const long long seg0 = nb * (long long)io.V - io.T1 - io.shift + (long long)(tid >> 5) * 512;
            if ((tid & 31) == 0 && nb < nblocks && seg0 >= 0 && seg0 + 512 <= io.n_in) {
                const unsigned long long a = (reinterpret_cast<unsigned long long>(io.in + seg0) + 15ull) & ~15ull;
                asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(a), "r"(4096 - 16) : "memory");
            }
        }
        fftk16::phase_b(tid, s_tw2, sm);
        plane_barrier(tid);
        fftk16::phase_c(tid, s_tw3, sm);
        plane_barrier(tid);
        fftk16::phase_d(tid, Hd, s_hres, sm);
        plane_barrier(tid);
        fftk16::phase_ci(tid, s_tw3, sm);
        plane_barrier(tid);
        fftk16::phase_bi(tid, s_tw2, sm);
        __syncthreads();
        fftk16::phase_ai<DECIM, ACCUM>(tid, blk, io, s_tw1, sm);
    }
}

// hist_next[i] = x[n - T1 + i] over the concatenation (hist_cur ++ in).
__global__ void fftfilt_hist_kernel(const float2* __restrict__ hist_cur, const float2* __restrict__ in,
                                    long long n, int T1, float2* __restrict__ hist_next, int in_u8) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < T1; i += gridDim.x * blockDim.x) {
        const long long s = n - T1 + i;
        hist_next[i] = s >= 0 ? fftr::ld_iq(in, s, in_u8) : hist_cur[s + T1];
    }
}

// Real-stream history: T1 floats.
__global__ void fftfilt_hist_real_kernel(const float* __restrict__ hist_cur, const float* __restrict__ in,
                                         long long n, int T1, float* __restrict__ hist_next) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < T1; i += gridDim.x * blockDim.x) {
        const long long s = n - T1 + i;
        hist_next[i] = s >= 0 ? in[s] : hist_cur[s + T1];
    }
}

}  // namespace rrc

using namespace rrc;


namespace {

size_t ref_fft_size(size_t ntaps) {   // calc_fft_size, src/fft_filter.rs:36-42
    size_t n = 1;
    while (n < ntaps) n <<= 1;
    return 2 * n;
}

// Tap partitioning: filters with more taps than one 16384-point block can hold efficiently
// are split into partitions of PART_TAPS taps (valid fraction >= 50% per partition).
constexpr size_t PART_TAPS = 8193;
constexpr size_t SINGLE_MAX_TAPS = 12289;     // up to here one partition (valid >= 25%) beats two

template <bool DECIM, bool ACCUM>
int launch_part16(rrc_fftfilt* h, const BlockIO& io, const float2* Hd, cudaStream_t st) {
    const long long nblocks = (io.n_in + io.V - 1) / io.V;
    const int grid = (int)std::min<long long>(nblocks, sm_count(h->device));
    auto kern = fftfilt16_kernel<DECIM, ACCUM>;
    RRC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)FFTFILT16_SMEM));
    kern<<<grid, fftk16::NT16, FFTFILT16_SMEM, st>>>(io, Hd, h->tw1_16, h->tw2_16, h->tw3_16, nblocks);
    RRC_CHECK_LAUNCH();
    count_launch();
    return RRC_OK;
}

template <bool DECIM, bool ACCUM>
int launch_part(rrc_fftfilt* h, const BlockIO& io, const float2* Hp, const float2* Hq, cudaStream_t st) {
    long long nblocks = (io.n_in + io.V - 1) / io.V;
    if (io.real) nblocks = (nblocks + 1) / 2;                   // two real blocks per complex transform
    const int grid = (int)std::min<long long>(nblocks, sm_count(h->device));
    if ((h->variant == 37 || h->variant == 38 || h->variant == 39) && !io.real && !io.in_u8) {
        auto pk = h->variant == 38 ? fftfilt_pk_kernel<DECIM, ACCUM, 1> : h->variant == 39 ? fftfilt_pk_kernel<DECIM, ACCUM, 2> : fftfilt_pk_kernel<DECIM, ACCUM, 0>;
        RRC_CUDA(cudaFuncSetAttribute(pk, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)FFTFILT_PK_SMEM));
        static const int tune_pk = [] { const char* e = getenv("RRC_FFTFILT_TUNE"); return e ? (int)strtol(e, nullptr, 0) : 1; }();
        static const bool want_trace = getenv("RRC_FFTFILT_TRACE") != nullptr;
        long long* dtrace = nullptr;
        const size_t trace_n = (size_t)PK_TRACE_BLK * 16 * PK_NSTAMP;
        if (want_trace) { RRC_CUDA(cudaMalloc((void**)&dtrace, trace_n * 8)); RRC_CUDA(cudaMemsetAsync(dtrace, 0, trace_n * 8, st)); }
        pk<<<grid, fftk::NT, FFTFILT_PK_SMEM, st>>>(io, Hq, h->tw1, h->tw2p, nblocks, tune_pk, dtrace);
        RRC_CHECK_LAUNCH();
        count_launch();
        if (want_trace) {                                       // debug only: synchronous dump of the per-phase cycle table
            std::vector<long long> tr(trace_n);
            RRC_CUDA(cudaStreamSynchronize(st));
            RRC_CUDA(cudaMemcpy(tr.data(), dtrace, trace_n * 8, cudaMemcpyDeviceToHost));
            cudaFree(dtrace);
            static const char* names[] = {"wait mbarrier (TMA landed)", "A (L0 read, DIF32, tw, L1 write)", "barrier 1", "B", "C", "B'", "barrier 2",
                                          "A': loads", "barrier 3 + stage", "A': tw, IDFT32, STG"};
            static int dumps = 0;
            if (nblocks > 148 * 8 && dumps++ < 2) {
                for (int b = 0; b < PK_TRACE_BLK; ++b) {
                    long long t0 = tr[(size_t)(b * 16) * PK_NSTAMP], tend = 0;
                    for (int w = 0; w < 16; ++w) { t0 = std::min(t0, tr[(size_t)(b * 16 + w) * PK_NSTAMP]); tend = std::max(tend, tr[(size_t)(b * 16 + w) * PK_NSTAMP + 10]); }
                    fprintf(stderr, "pk trace (variant %d) block iter %d: total %lld cycles\n", h->variant, b + 2, tend - t0);
                    for (int p = 0; p < 10; ++p) {
                        std::vector<long long> d;
                        for (int w = 0; w < 16; ++w) d.push_back(tr[(size_t)(b * 16 + w) * PK_NSTAMP + p + 1] - tr[(size_t)(b * 16 + w) * PK_NSTAMP + p]);
                        std::sort(d.begin(), d.end());
                        fprintf(stderr, "   %-34s min %6lld  med %6lld  max %6lld\n", names[p], d[0], d[8], d[15]);
                    }
                }
            }
        }
        return RRC_OK;
    }
    if ((h->variant == 40 || h->variant == 41 || h->variant == 42) && !io.real && !io.in_u8) {
        auto tk = h->variant == 42 ? fftfilt_tmh_kernel<DECIM, ACCUM, false, true> : h->variant == 41 ? fftfilt_tmh_kernel<DECIM, ACCUM, true> : fftfilt_tmh_kernel<DECIM, ACCUM, false>;
        RRC_CUDA(cudaFuncSetAttribute(tk, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)FFTFILT_TMH_SMEM));
        static const int tune = [] { const char* e = getenv("RRC_FFTFILT_TUNE"); return e ? (int)strtol(e, nullptr, 0) : 1; }();
        tk<<<grid, fftk::NT, FFTFILT_TMH_SMEM, st>>>(io, Hp, h->tw1, h->tw2, nblocks, tune);
        RRC_CHECK_LAUNCH();
        count_launch();
        return RRC_OK;
    }
    auto kern = io.real ? fftfilt_kernel<DECIM, ACCUM, false> : h->variant == 33 ? fftfilt_pp_kernel<DECIM, ACCUM> : h->variant == 34 ? fftfilt_st_kernel<DECIM, ACCUM> : h->variant == 35 ? fftfilt_kernel<DECIM, ACCUM, true> : (h->variant == 36 && !io.in_u8) ? fftfilt_tma_kernel<DECIM, ACCUM> : fftfilt_kernel<DECIM, ACCUM, false>;
    RRC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)FFTFILT_SMEM));
    static const int tune = [] { const char* e = getenv("RRC_FFTFILT_TUNE"); return e ? (int)strtol(e, nullptr, 0) : 1; }();
    kern<<<grid, fftk::NT, FFTFILT_SMEM, st>>>(io, Hp, h->tw1, h->tw2, nblocks, tune);
    RRC_CHECK_LAUNCH();
    count_launch();
    return RRC_OK;
}

int launch(rrc_fftfilt* h, const float* in, size_t n, float* out, size_t n_out, size_t deci, size_t skip, cudaStream_t st) {
    BlockIO io;
    io.in = reinterpret_cast<const float2*>(in);
    io.hist = h->hist_ext ? h->hist_ext : h->hist[h->cur];
    h->hist_ext = nullptr;                                       // one-shot
    const float2* hist_used = io.hist;
    io.out = reinterpret_cast<float2*>(out);
    io.n_in = (long long)n;
    io.n_out = (long long)n_out;
    io.T1_total = h->T1;
    io.deci = (int)deci;
    io.skip = (long long)skip;
    io.in_u8 = h->in_u8;
    io.real = h->real;
    io.epi = h->epi;
    const bool decim = !(deci == 1 && skip == 0);
    // kernels that update the carried history themselves (one launch per run): the 512-thread LDG / TMA kernels
    const bool fused_hist = h->T1 > 0 && (h->real || h->variant == 32 || h->variant == 35 || h->variant == 36 || h->variant == 37 || h->variant == 38 || h->variant == 39 || h->variant == 40 || h->variant == 41 || h->variant == 42);
    long long shift = 0;
    if (h->epi.kind != RRC_EPI_NONE && h->part_T1.size() > 1)
        return fail(RRC_ERR_UNSUPPORTED, "store epilogues need a single tap partition (ntaps <= 12289) on this path");
    for (size_t p = 0; p < h->part_T1.size(); ++p) {
        io.T1 = h->part_T1[p];
        io.V = fftk::N - io.T1;
        io.shift = shift;
        io.hist_next = (fused_hist && p == 0) ? h->hist[h->cur ^ 1] : nullptr;
        int s;
        if (h->variant == 16 && !h->real) {
            if (p == 0) s = decim ? launch_part16<true, false>(h, io, h->part_Hd[p], st) : launch_part16<false, false>(h, io, h->part_Hd[p], st);
            else        s = decim ? launch_part16<true, true>(h, io, h->part_Hd[p], st) : launch_part16<false, true>(h, io, h->part_Hd[p], st);
        } else if (p == 0) s = decim ? launch_part<true, false>(h, io, h->part_Hp[p], h->part_Hq[p], st) : launch_part<false, false>(h, io, h->part_Hp[p], h->part_Hq[p], st);
        else        s = decim ? launch_part<true, true>(h, io, h->part_Hp[p], h->part_Hq[p], st) : launch_part<false, true>(h, io, h->part_Hp[p], h->part_Hq[p], st);
        RRC_TRY(s);
        shift += io.T1 + 1;
    }
    if (fused_hist) {
        h->cur ^= 1;
    } else if (h->T1 > 0 && h->real) {
        fftfilt_hist_real_kernel<<<(h->T1 + 255) / 256, 256, 0, st>>>(reinterpret_cast<const float*>(hist_used), in, (long long)n, h->T1,
                                                                    reinterpret_cast<float*>(h->hist[h->cur ^ 1]));
        RRC_CHECK_LAUNCH();
        count_launch();
        h->cur ^= 1;
    } else if (h->T1 > 0) {
        fftfilt_hist_kernel<<<(h->T1 + 255) / 256, 256, 0, st>>>(hist_used, io.in, (long long)n, h->T1, h->hist[h->cur ^ 1], h->in_u8);
        RRC_CHECK_LAUNCH();
        count_launch();
        h->cur ^= 1;
    }
    return RRC_OK;
}

// reset()/set_history() wrote handle state on `st`: remember it so the host pipelines can wait for it.
int mark_state(rrc_fftfilt* h, cudaStream_t st) {
    if (!h->state_ev) RRC_CUDA(cudaEventCreateWithFlags(&h->state_ev, cudaEventDisableTiming));
    RRC_CUDA(cudaEventRecord(h->state_ev, st));
    h->state_dirty = true;
    return RRC_OK;
}
int pipe_wait_state(rrc_fftfilt* h) {
    if (h->state_dirty) {
        RRC_CUDA(cudaStreamWaitEvent(h->pipe.s_comp, h->state_ev, 0));
        h->state_dirty = false;
    }
    return RRC_OK;
}

}  // namespace

extern "C" {

int rrc_fftfilt_ref_fft_size(size_t ntaps, size_t* fft_size, size_t* nsamples) {
    if (ntaps == 0) return fail(RRC_ERR_INVALID, "FftFilter needs at least one tap (src/fft_filter.rs:146)");
    const size_t f = ref_fft_size(ntaps);
    if (fft_size) *fft_size = f;
    if (nsamples) *nsamples = f - ntaps;
    return RRC_OK;
}

int rrc_fftfilt_plan(size_t ntaps, size_t buffered, size_t in_len, size_t out_free,
                     size_t* blocks, size_t* consume, size_t* buffered_after, size_t* wait_need, int* wait_on_output) {
    if (!blocks || !consume || !buffered_after || !wait_need || !wait_on_output) return fail(RRC_ERR_INVALID, "NULL argument");
    if (ntaps == 0) return fail(RRC_ERR_INVALID, "ntaps must be nonzero");
    const size_t S = ref_fft_size(ntaps) - ntaps;      // nsamples, src/fft_filter.rs:262-263
    if (buffered >= S) return fail(RRC_ERR_INVALID, "buffered %zu >= nsamples %zu", buffered, S);
    size_t nb = 0, used = 0, b = buffered, avail = in_len, space = out_free;
    for (;;) {
        if (S > space) { *wait_need = S; *wait_on_output = 1; break; }            // :293-303
        const size_t add = std::min(avail, S - b);                                 // :306
        b += add; avail -= add; used += add;                                       // :308,314
        if (b < S) { *wait_need = S - b; *wait_on_output = 0; break; }             // :315-327
        ++nb; space -= S; b = 0;                                                   // :331-352
    }
    *blocks = nb; *consume = used; *buffered_after = b;
    return RRC_OK;
}

int rrc_fftfilt_c32_create(int device, const float* taps, size_t ntaps, rrc_fftfilt_t** out) {
    if (!out) return fail(RRC_ERR_INVALID, "out is NULL");
    *out = nullptr;
    if (!taps || ntaps == 0) return fail(RRC_ERR_INVALID, "FftFilter needs at least one tap (src/fft_filter.rs:146)");
    if (ntaps > ((size_t)1 << 24)) return fail(RRC_ERR_UNSUPPORTED, "ntaps %zu > 2^24", ntaps);
    RRC_CUDA(cudaSetDevice(device));
    auto* h = new rrc_fftfilt();
    h->device = device; h->ntaps = ntaps;
    h->taps_host.assign(taps, taps + 2 * ntaps);
    h->T1 = (int)ntaps - 1;
    auto cleanup = [&](int s) { rrc_fftfilt_destroy(h); return s; };
    auto up = [&](float2** d, const std::vector<float2>& v) -> cudaError_t {
        cudaError_t e = cudaMalloc((void**)d, v.size() * sizeof(float2));
        if (e != cudaSuccess) return e;
        return upload_sync(*d, v.data(), v.size() * sizeof(float2));
    };
    cudaError_t e;
    std::vector<float2> Hp, tw1, tw2;
    const size_t part = ntaps <= SINGLE_MAX_TAPS ? ntaps : PART_TAPS;
    for (size_t off = 0; off < ntaps; off += part) {
        const size_t len = std::min(part, ntaps - off);
        fftk::build_tables(taps + 2 * off, len, Hp, tw1, tw2);
        float2* d = nullptr;
        if ((e = up(&d, Hp)) != cudaSuccess) return cleanup(fail(RRC_ERR_CUDA, "FftFilter table upload failed: %s", cudaGetErrorString(e)));
        h->part_Hp.push_back(d);
        h->part_T1.push_back((int)len - 1);
        std::vector<float2> Hq, tw2p;
        fftk::build_tables_pk(taps + 2 * off, len, Hq, tw2p);
        float2* dq = nullptr;
        if ((e = up(&dq, Hq)) != cudaSuccess) return cleanup(fail(RRC_ERR_CUDA, "FftFilter table upload failed: %s", cudaGetErrorString(e)));
        h->part_Hq.push_back(dq);
        if (!h->tw2p && (e = up(&h->tw2p, tw2p)) != cudaSuccess) return cleanup(fail(RRC_ERR_CUDA, "FftFilter table upload failed: %s", cudaGetErrorString(e)));
        std::vector<float2> Hd, t1, t2, t3;
        fftk::build_tables16(taps + 2 * off, len, Hd, t1, t2, t3);
        float2* dd = nullptr;
        if ((e = up(&dd, Hd)) != cudaSuccess) return cleanup(fail(RRC_ERR_CUDA, "FftFilter table upload failed: %s", cudaGetErrorString(e)));
        h->part_Hd.push_back(dd);
        if (!h->tw1_16 && ((e = up(&h->tw1_16, t1)) != cudaSuccess || (e = up(&h->tw2_16, t2)) != cudaSuccess || (e = up(&h->tw3_16, t3)) != cudaSuccess))
            return cleanup(fail(RRC_ERR_CUDA, "FftFilter table upload failed: %s", cudaGetErrorString(e)));
    }
    if (const char* v = getenv("RRC_FFTFILT_VARIANT")) h->variant = atoi(v) == 16 ? 16 : atoi(v) == 33 ? 33 : atoi(v) == 34 ? 34 : atoi(v) == 35 ? 35 : atoi(v) == 36 ? 36 : atoi(v) == 37 ? 37 : atoi(v) == 38 ? 38 : atoi(v) == 39 ? 39 : atoi(v) == 40 ? 40 : atoi(v) == 41 ? 41 : atoi(v) == 42 ? 42 : 32;
    h->Hp = h->part_Hp[0];
    if ((e = up(&h->tw1, tw1)) != cudaSuccess || (e = up(&h->tw2, tw2)) != cudaSuccess)
        return cleanup(fail(RRC_ERR_CUDA, "FftFilter table upload failed: %s", cudaGetErrorString(e)));
    for (int i = 0; i < 2; ++i) {
        const size_t bytes = std::max<size_t>(1, (size_t)h->T1) * sizeof(float2);
        if ((e = cudaMalloc((void**)&h->hist[i], bytes)) != cudaSuccess || (e = zero_sync(h->hist[i], bytes)) != cudaSuccess)
            return cleanup(fail(RRC_ERR_CUDA, "FftFilter history alloc failed: %s", cudaGetErrorString(e)));
    }
    cudaStreamSynchronize(0);      // callers run on non-blocking streams, which do not wait for the default-stream fills
    *out = h;
    return RRC_OK;
}

int rrc_fftfilt_f32_create(int device, const float* taps_f32, size_t ntaps, rrc_fftfilt_t** out) {
    if (!out) return fail(RRC_ERR_INVALID, "out is NULL");
    *out = nullptr;
    if (!taps_f32 || ntaps == 0) return fail(RRC_ERR_INVALID, "FftFilterFloat needs at least one tap (src/fft_filter.rs:146)");
    std::vector<float> ct(2 * ntaps, 0.f);                      // taps -> Complex::new(f, 0.0), src/fft_filter.rs:404
    for (size_t i = 0; i < ntaps; ++i) ct[2 * i] = taps_f32[i];
    RRC_TRY(rrc_fftfilt_c32_create(device, ct.data(), ntaps, out));
    (*out)->real = 1;
    (*out)->variant = 32;
    return RRC_OK;
}

int rrc_fftfilt_destroy(rrc_fftfilt_t* h) {
    if (!h) return RRC_OK;
    cudaSetDevice(h->device);
    for (float2* p : h->part_Hp) cudaFree(p);
    for (float2* p : h->part_Hd) cudaFree(p);
    for (float2* p : h->part_Hq) cudaFree(p);
    cudaFree(h->tw2p);
    cudaFree(h->tw1_16); cudaFree(h->tw2_16); cudaFree(h->tw3_16);
    cudaFree(h->tw1); cudaFree(h->tw2); cudaFree(h->hist[0]); cudaFree(h->hist[1]);
    if (h->state_ev) cudaEventDestroy(h->state_ev);
    fold_destroy(h);
    poly_destroy(h);
    h->pipe.destroy();
    delete h;
    return RRC_OK;
}

int rrc_fftfilt_reset(rrc_fftfilt_t* h, void* stream) {
    if (!h) return fail(RRC_ERR_INVALID, "fftfilt handle is NULL");
    RRC_CUDA(cudaSetDevice(h->device));
    h->hist_ext = nullptr;
    if (h->T1 > 0) {
        RRC_CUDA(cudaMemsetAsync(h->hist[h->cur], 0, (size_t)h->T1 * sizeof(float2), as_stream(stream)));
        RRC_TRY(mark_state(h, as_stream(stream)));
    }
    return RRC_OK;
}

int rrc_fftfilt_set_history(rrc_fftfilt_t* h, const float* hist, size_t n, void* stream) {
    if (!h) return fail(RRC_ERR_INVALID, "fftfilt handle is NULL");
    if (n != (size_t)h->T1) return fail(RRC_ERR_INVALID, "history must be ntaps-1 = %d samples, got %zu", h->T1, n);
    if (n == 0) return RRC_OK;
    if (!hist) return fail(RRC_ERR_INVALID, "hist is NULL");
    RRC_CUDA(cudaSetDevice(h->device));
    h->hist_ext = nullptr;
    // cudaMemcpyDefault: `hist` may be a peer device's memory (IPC-mapped or peer-enabled): the copy then runs over NVLink.
    RRC_CUDA(cudaMemcpyAsync(h->hist[h->cur], hist, n * (h->real ? sizeof(float) : sizeof(float2)), cudaMemcpyDefault, as_stream(stream)));
    RRC_TRY(mark_state(h, as_stream(stream)));
    return RRC_OK;
}

int rrc_fftfilt_set_history_ptr(rrc_fftfilt_t* h, const float* hist, size_t n) {
    if (!h) return fail(RRC_ERR_INVALID, "fftfilt handle is NULL");
    if (n != (size_t)h->T1) return fail(RRC_ERR_INVALID, "history must be ntaps-1 = %d samples, got %zu", h->T1, n);
    if (n && !hist) return fail(RRC_ERR_INVALID, "hist is NULL");
    h->hist_ext = n ? reinterpret_cast<const float2*>(hist) : nullptr;
    return RRC_OK;
}

int rrc_fftfilt_set_epilogue(rrc_fftfilt_t* h, int kind, float re, float im) {
    if (!h) return fail(RRC_ERR_INVALID, "fftfilt handle is NULL");
    if (kind < RRC_EPI_NONE || kind > RRC_EPI_MAG2) return fail(RRC_ERR_INVALID, "unknown epilogue %d", kind);
    if (kind != RRC_EPI_NONE && h->real) return fail(RRC_ERR_INVALID, "store epilogues exist for Complex filters only");
    h->epi.kind = kind; h->epi.re = re; h->epi.im = im;
    return RRC_OK;
}

int rrc_fftfilt_set_input_u8iq(rrc_fftfilt_t* h, int on) {
    if (!h) return fail(RRC_ERR_INVALID, "fftfilt handle is NULL");
    if (on && h->real) return fail(RRC_ERR_INVALID, "u8 I/Q input needs a Complex filter");
    h->in_u8 = on ? 1 : 0;
    return RRC_OK;
}

int rrc_fftfilt_geometry(const rrc_fftfilt_t* h, size_t* fft_size, size_t* valid) {
    if (!h) return fail(RRC_ERR_INVALID, "fftfilt handle is NULL");
    if (fft_size) *fft_size = fftk::N;
    if (valid) *valid = (size_t)(fftk::N - h->part_T1[0]);
    return RRC_OK;
}

int rrc_fftfilt_run(rrc_fftfilt_t* h, const float* in, size_t n, float* out, void* stream) {
    if (!h) return fail(RRC_ERR_INVALID, "fftfilt handle is NULL");
    if (n == 0) return RRC_OK;
    if (!in || !out) return fail(RRC_ERR_INVALID, "in/out is NULL");
    if (h->in_u8 && (reinterpret_cast<uintptr_t>(in) & 1)) return fail(RRC_ERR_INVALID, "u8 I/Q input must be 2-byte aligned");
    RRC_CUDA(cudaSetDevice(h->device));
    return launch(h, in, n, out, n, 1, 0, as_stream(stream));
}

int rrc_fftfilt_decim_run(rrc_fftfilt_t* h, const float* in, size_t n, size_t deci, size_t skip,
                          float* out, size_t* n_out, void* stream) {
    if (!h) return fail(RRC_ERR_INVALID, "fftfilt handle is NULL");
    if (deci == 0) return fail(RRC_ERR_INVALID, "deci must be nonzero");
    const size_t cnt = n > skip ? (n - skip + deci - 1) / deci : 0;
    if (n_out) *n_out = cnt;
    if (n == 0) return RRC_OK;
    if (!in || (!out && cnt)) return fail(RRC_ERR_INVALID, "in/out is NULL");
    if (h->in_u8 && (reinterpret_cast<uintptr_t>(in) & 1)) return fail(RRC_ERR_INVALID, "u8 I/Q input must be 2-byte aligned");
    RRC_CUDA(cudaSetDevice(h->device));
    // deci == 1 && skip == 0 degenerates to the plain path; otherwise the store
    // predicate in phase A' keeps y[skip + k*deci].
    if (deci == 1 && skip == 0) return launch(h, in, n, out, n, 1, 0, as_stream(stream));
    if (h->real) return fail(RRC_ERR_UNSUPPORTED, "fused decimation is not implemented for real (f32) streams");
    // deci == 8: folded spectrum + 8x smaller inverse transform (fftfilt_fold.cu); 65536-point
    // cluster kernel for 12289 < ntaps <= 49153.  RRC_FFTFILT_NO_FOLD=1 forces the store-predicate path.
    // 2 <= deci <= 16: polyphase form (fftfilt_poly.cu: deci forward transforms on the deci-times slower branch streams, one
    // inverse per block).  RRC_FFTFILT_NO_POLY=1 falls through to the fold kernel / the store predicate.
    const bool poly = poly_supported(h, deci) == RRC_OK;
    if (poly || fold_supported(h, deci) == RRC_OK) {
        const float2* hist_used = h->hist_ext ? h->hist_ext : h->hist[h->cur];
        // both kernels read hist_ext / hist[cur] and write the next history themselves (one launch);
        // RRC_ERR_UNSUPPORTED = nothing was launched (no kept output in this call)
        const int s = !cnt ? RRC_ERR_UNSUPPORTED
                    : poly ? poly_launch(h, in, n, out, cnt, deci, skip, as_stream(stream))
                           : fold_launch(h, in, n, out, cnt, skip, as_stream(stream));
        h->hist_ext = nullptr;                                                          // one-shot
        if (s != RRC_OK && s != RRC_ERR_UNSUPPORTED) return s;
        if (s == RRC_ERR_UNSUPPORTED && h->T1 > 0) {
            fftfilt_hist_kernel<<<(h->T1 + 255) / 256, 256, 0, as_stream(stream)>>>(
                hist_used, reinterpret_cast<const float2*>(in), (long long)n, h->T1, h->hist[h->cur ^ 1], h->in_u8);
            RRC_CHECK_LAUNCH();
            count_launch();
        }
        if (h->T1 > 0) h->cur ^= 1;
        return RRC_OK;
    }
    return launch(h, in, n, out, cnt, deci, skip, as_stream(stream));
}

int rrc_fftfilt_run_host(rrc_fftfilt_t* h, const float* in_host, size_t n_in, float* out_host, size_t* n_out) {
    if (!h) return fail(RRC_ERR_INVALID, "fftfilt handle is NULL");
    // Reference count rule: whole blocks of nsamples only (src/fft_filter.rs:315-327).
    const size_t S = ref_fft_size(h->ntaps) - h->ntaps;
    const size_t total = (n_in / S) * S;
    if (n_out) *n_out = total;
    if (total == 0) return RRC_OK;
    if (!in_host || !out_host) return fail(RRC_ERR_INVALID, "in/out is NULL");
    RRC_TRY(h->pipe.init(h->device));
    RRC_TRY(pipe_wait_state(h));
    const size_t chunk = pipe_chunk_samples_for(total);
    const size_t esz = h->real ? sizeof(float) : h->in_u8 ? 2 : sizeof(float2);   // input bytes per sample
    const size_t osz = (h->real || h->epi.kind == RRC_EPI_MAG2) ? sizeof(float) : sizeof(float2);
    RRC_TRY(h->pipe.reserve(std::min(chunk, total) * esz, std::min(chunk, total) * osz));
    int i = 0;
    for (size_t off = 0, n = 0; off < total; off += n, ++i) {
        n = pipe_next_chunk((size_t)i, total - off, chunk);
        RRC_TRY(h->pipe.stage_in(i, reinterpret_cast<const char*>(in_host) + off * esz, n * esz));
        RRC_TRY(launch(h, (const float*)h->pipe.d_in[i & 1], n, (float*)h->pipe.d_out[i & 1], n, 1, 0, h->pipe.s_comp));
        RRC_TRY(h->pipe.drain_out(i, reinterpret_cast<char*>(out_host) + off * osz, n * osz));
    }
    return h->pipe.finish();
}

int rrc_fftfilt_decim_run_host(rrc_fftfilt_t* h, const float* in_host, size_t n_in, size_t deci, float* out_host, size_t* n_out) {
    if (!h) return fail(RRC_ERR_INVALID, "fftfilt handle is NULL");
    if (deci == 0) return fail(RRC_ERR_INVALID, "deci must be nonzero");
    // FftFilter emits whole blocks of nsamples (src/fft_filter.rs:315-327); RationalResampler(1, deci)
    // keeps every deci-th of them starting with the first (src/rational_resampler.rs:181-198).
    const size_t S = ref_fft_size(h->ntaps) - h->ntaps;
    const size_t total = (n_in / S) * S;
    const size_t total_out = (total + deci - 1) / deci;
    if (n_out) *n_out = total_out;
    if (total == 0) return RRC_OK;
    if (!in_host || !out_host) return fail(RRC_ERR_INVALID, "in/out is NULL");
    // real (FftFilterFloat) handles: plain path only (deci 1), f32 elements; the fused decimation is complex only.
    if (h->real && deci != 1) return fail(RRC_ERR_UNSUPPORTED, "fused decimation is not implemented for real (f32) streams");
    if (h->real) return rrc_fftfilt_run_host(h, in_host, n_in, out_host, n_out);
    RRC_TRY(h->pipe.init(h->device));
    RRC_TRY(pipe_wait_state(h));
    const size_t chunk = pipe_chunk_samples_for(total);
    const size_t esz = h->in_u8 ? 2 : sizeof(float2);
    const size_t osz = h->epi.kind == RRC_EPI_MAG2 ? sizeof(float) : sizeof(float2);
    RRC_TRY(h->pipe.reserve(std::min(chunk, total) * esz, (std::min(chunk, total) / deci + 2) * osz));
    int i = 0;
    size_t produced = 0;
    for (size_t off = 0, n = 0; off < total; off += n, ++i) {
        n = pipe_next_chunk((size_t)i, total - off, chunk);
        const size_t skip = (deci - off % deci) % deci;           // first kept output of this chunk
        size_t cnt = 0;
        RRC_TRY(h->pipe.stage_in(i, reinterpret_cast<const char*>(in_host) + off * esz, n * esz));
        RRC_CUDA(cudaSetDevice(h->device));
        RRC_TRY(rrc_fftfilt_decim_run(h, (const float*)h->pipe.d_in[i & 1], n, deci, skip, (float*)h->pipe.d_out[i & 1], &cnt, h->pipe.s_comp));
        RRC_TRY(h->pipe.drain_out(i, reinterpret_cast<char*>(out_host) + produced * osz, cnt * osz));
        produced += cnt;
    }
    if (produced != total_out) return fail(RRC_ERR_STATE, "decim_run_host produced %zu of %zu", produced, total_out);
    return h->pipe.finish();
}

}  // extern "C"
