// fftfilt_poly.cu — FftFilter fused with decimate-by-D as a polyphase overlap-save filter (fftfilt_poly_core.cuh):
// D forward 16384-point transforms (one per polyphase branch, on the D-times slower streams), the sum over the
// branches in TENSOR MEMORY, one inverse transform per block of V = 16384 - (ceil(ntaps / D) - 1) kept outputs.
//
// Replaces FftFilter::work's engine step (rustradio src/fft_filter.rs:172-176,281-287,331-348) followed by
// RationalResampler(1, D) (src/rational_resampler.rs:155-206) — BASELINE config 5 (16385 taps, D = 8).
//
// Work split (template parameter C = CTAs per cluster, 1, 2 or 4):
//   * A block reads D * 16384 consecutive input samples (1 MiB for D = 8); every branch touches every 128-byte line of
//     it, so the region has to stay in L2 for the whole block.  With one block per SM that is 148 MiB of live lines —
//     more than the 126 MB L2 (measured: 15-20 K cycles per branch gather, 310 K cycles per block).  The C CTAs of a
//     cluster therefore share ONE block: CTA c transforms branches c*D/C .. (c+1)*D/C - 1 (D = 8, C = 4: the two
//     branches that live in the same aligned 16 bytes, fetched with ONE 128-bit load per k — half the sectors and half
//     the LSU time of two 64-bit gathers; the second branch waits in a 128 KiB TMEM stash).  A hardware cluster, not a
//     software group: co-scheduled CTAs ask for the same lines at the same time (same code on cooperative groups with a
//     global-memory barrier and all 148 SMs: gather 25 K cycles instead of 10 K; 33 four-CTA clusters = 132 SMs fit).
//   * Every CTA writes its partial sum to an L2 scratch slot (coalesced 128-bit rows; 4 slots per CTA, indexed by the
//     iteration) and arrives on the cluster barrier (release).  The block's FINISHER (rotating: iteration it -> CTA
//     it % C) runs the inverse transform ONE ITERATION LATER, after its forward transforms of the next block: its own
//     next partial sum is then never late, so no CTA ever waits for a finisher (with immediate finishing the barrier
//     chain serialised forward + inverse: measured 166 K cycles per finisher iteration against 61 K).  Over C blocks every
//     CTA does D forward transforms and one inverse: balanced.
//   * HRES (off by default): the spectrum rows k2 = l of the branch in shared memory (one 72 KiB bulk copy per branch,
//     as fftfilt_tma_kernel keeps them resident).  Phase C drops from 8.8 K to 5.4 K cycles — and the gather rises from
//     10 K to 18-28 K: shared memory is taken from L1, and the L1 size bounds the sector-sparse gather's misses in flight.
//   * Rejected by measurement (profiles/r02_c5_poly_*): four LOADER warps (setmaxnreg 104 / 64) gathering the next branch
//     into the TMEM stash while the 16 transform warps compute — the gather's LSU time then lands on the transform warps'
//     shared-memory exchanges (phase B 3.5 K -> 9 K cycles, phase C 8.6 K -> 12 K): same total; an evict-first hint on the
//     gather (the CTAs of a cluster share sectors); prefetching two blocks ahead.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "common.cuh"
#include "fftfilt_handle.hpp"
#include "fftfilt_poly_core.cuh"
#include "tmem.cuh"
#include "fftfilt_tables.hpp"

namespace rrc {

using fftp::PolyIO;

// Shared memory: exchange buffer + twiddles (+ the resident spectrum rows).  What is not shared memory is L1, and the L1
// size bounds the sector-sparse gather's misses in flight: with the 72 KiB of resident rows the gather of a branch pair
// takes 18-28 K cycles instead of 10-13 K — more than the rows save in phase C.  Hence HRES = false by default.
constexpr size_t poly_smem(bool hres) { return (size_t)(fftk::SMEM_ELEMS + 512 + 512 + (hres ? fftk::HRES_ELEMS : 0) + 4) * sizeof(float2); }
constexpr int POLY_SLOTS = 4;                                  // scratch slots per CTA (see the kernel for why four)
constexpr int POLY_NSTAMP = 16, POLY_TRACE_IT = 6;

__device__ __forceinline__ void bar_sync(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }
__device__ __forceinline__ void bar_arrive(int id, int n) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(n) : "memory"); }

// The running sum of a thread: 64 columns of its lane (tcgen05.st completes before the same thread's next tcgen05.ld
// of those columns: wait::st right after the store).
struct TmemAcc {
    unsigned base;
    __device__ __forceinline__ void load(int, int half, float2 (&x)[16]) const { tm_ld32(base + 32u * half, x); }
    __device__ __forceinline__ void store(int, int half, const float2 (&x)[16]) const { tm_st32(base + 32u * half, x); tm_wait_st(); }
};

__device__ __forceinline__ void poly_mbar_wait(unsigned mbar, unsigned parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "POLY_MBAR_WAIT:\n"
        "mbarrier.try_wait.parity.shared.b64 p, [%0], %1;\n"
        "@p bra POLY_MBAR_DONE;\n"
        "bra POLY_MBAR_WAIT;\n"
        "POLY_MBAR_DONE:\n"
        "}" ::"r"(mbar), "r"(parity) : "memory");
}

// Branches r (-> v) and r + 1 (-> the stash) with 32 aligned 128-bit loads per thread, fftp::poly_load_pair with an L2
// evict-first hint: after this gather the region is dead (both branches of this CTA are on chip).
__device__ __forceinline__ void poly_load_pair_dev(int tid, long long blk, int r, const PolyIO& io, float2 (&v)[32], const struct TmemStash& stash,
                                                   unsigned long long pol);

struct TmemStash {
    unsigned base;
    __device__ __forceinline__ void load(int, int half, float2 (&x)[16]) const { tm_ld32(base + 32u * half, x); }
    __device__ __forceinline__ void store(int, int half, const float2 (&x)[16]) const { tm_st32(base + 32u * half, x); }
    __device__ __forceinline__ void store8(int, int quarter, const float2 (&x)[8]) const { tm_st16(base + 16u * quarter, x); }
};

__device__ __forceinline__ void poly_load_pair_dev(int tid, long long blk, int r, const PolyIO& io, float2 (&v)[32], const TmemStash& stash,
                                                   unsigned long long pol) {
    const float2* q = io.b.in + fftp::poly_g(io, blk, r, tid);
    const long long step = 512ll * io.D;
#pragma unroll
    for (int b = 0; b < 4; ++b) {
        float4 w[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            if (pol == 1ull)                                    // tune bit 32: no L1 allocation only
                asm volatile("ld.global.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];"
                             : "=f"(w[e].x), "=f"(w[e].y), "=f"(w[e].z), "=f"(w[e].w) : "l"(q + step * fftr::bitrev(8 * b + e, 5)));
            else
                asm volatile("ld.global.L1::no_allocate.L2::cache_hint.v4.f32 {%0, %1, %2, %3}, [%4], %5;"
                             : "=f"(w[e].x), "=f"(w[e].y), "=f"(w[e].z), "=f"(w[e].w) : "l"(q + step * fftr::bitrev(8 * b + e, 5)), "l"(pol));
        }
        float2 x[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) { v[8 * b + e] = make_float2(w[e].x, w[e].y); x[e] = make_float2(w[e].z, w[e].w); }
        stash.store8(tid, b, x);
    }
}

// Last phase C of a CTA's branches in a cluster: the partial sum goes to this CTA's scratch slot as 16 coalesced
// 128-bit rows per thread, slot[(half*8 + j)*512 + tid].
template <bool HRES>
__device__ __forceinline__ void phase_c_partial(int tid, const float2* Hp, const float2* Hres, const float2* sm, const TmemAcc& acc, bool first, float4* slot) {
    const int k1 = tid >> 4, l = tid & 15;
    const float4* hp1 = reinterpret_cast<const float4*>(Hp + (size_t)(k1 * 32 + l + 16) * 16);
    float4 h1[8];
    if constexpr (HRES) {
#pragma unroll
        for (int i = 0; i < 8; ++i) h1[i] = hp1[i];
    }
    const float4* hres = reinterpret_cast<const float4*>(Hres + tid * fftk::HRES_PITCH);
#pragma unroll
    for (int half = 0; half < 2; ++half) {
        const float2* row = sm + k1 * fftk::PLANE_PITCH + (l + 16 * half) * fftk::ROW_PITCH;
        float2 u[16];
        if constexpr (!HRES) {
            const float4* hp = reinterpret_cast<const float4*>(Hp + (size_t)(k1 * 32 + l + 16 * half) * 16);
#pragma unroll
            for (int i = 0; i < 8; ++i) h1[i] = hp[i];
        }
        fftp::poly_c_row(tid, half, (HRES && half == 0) ? hres : h1, row, acc, first, u);
#pragma unroll
        for (int j = 0; j < 8; ++j) __stcg(slot + (half * 8 + j) * 512 + tid, make_float4(u[2 * j].x, u[2 * j].y, u[2 * j + 1].x, u[2 * j + 1].y));
    }
}

// The finisher's step: sum of the C partial sums of one block -> inverse DFT16 -> rows for B' / A'.  Touches only the
// rows this thread's half-warp owns.
template <int C>
__device__ __forceinline__ void phase_c_finish(int tid, float2* sm, const float4* slots /* [C][POLY_SLOTS][8192], this cluster */, int slot) {
    const int k1 = tid >> 4, l = tid & 15;
#pragma unroll
    for (int half = 0; half < 2; ++half) {
        const int k2 = l + 16 * half;
        float2* row = sm + k1 * fftk::PLANE_PITCH + k2 * fftk::ROW_PITCH;
        float2 u[16];
#pragma unroll
        for (int oc = 0; oc < C; oc += 2) {                     // the rows of two CTAs in flight at a time (C is 2 or 4)
            const float4* s0 = slots + ((size_t)(oc * POLY_SLOTS + slot) * 8192) + (half * 8) * 512 + tid;
            const float4* s1 = s0 + (size_t)POLY_SLOTS * 8192;
            float4 x[8], y[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) { x[j] = __ldcg(s0 + j * 512); y[j] = __ldcg(s1 + j * 512); }
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const float2 a = make_float2(x[j].x + y[j].x, x[j].y + y[j].y), b = make_float2(x[j].z + y[j].z, x[j].w + y[j].w);
                if (oc == 0) { u[2 * j] = a; u[2 * j + 1] = b; }
                else { u[2 * j].x += a.x; u[2 * j].y += a.y; u[2 * j + 1].x += b.x; u[2 * j + 1].y += b.y; }
            }
        }
        float2 v[16];
#pragma unroll
        for (int k3 = 0; k3 < 16; ++k3) v[fftr::bitrev(k3, 4)] = u[k3];
        fftr::dit<16, -1>(v);
#pragma unroll
        for (int n3 = 0; n3 < 16; ++n3) row[n3] = v[n3];
    }
}

template <int C, bool HRES>
__global__ void __launch_bounds__(512, 1)
fftfilt_poly_kernel(const PolyIO io, const float2* __restrict__ Hph, const float2* __restrict__ tw1g,
                    const float2* __restrict__ tw2g, float4* __restrict__ scratch, long long blk0, long long nblocks, int tune,
                    long long* __restrict__ trace) {
    extern __shared__ __align__(16) float2 sm[];
    float2* s_tw2 = sm + fftk::SMEM_ELEMS;
    float2* s_tw1 = s_tw2 + 512;
    float2* s_hres = s_tw1 + 512;
    constexpr int HRES_N = HRES ? fftk::HRES_ELEMS : 0;
    unsigned* s_tmem = reinterpret_cast<unsigned*>(s_hres + HRES_N + 2);
    const unsigned mbar = (unsigned)__cvta_generic_to_shared(s_hres + HRES_N);
    const unsigned hres_a = (unsigned)__cvta_generic_to_shared(s_hres);
    const float2* Hresg = Hph + (size_t)io.D * fftk::N;       // [D][HRES_ELEMS]: rows k2 = l of every branch, shared-memory layout
    unsigned hpar = 0;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    int c = 0;
    if constexpr (C > 1) asm("mov.u32 %0, %%cluster_ctarank;" : "=r"(c));
    s_tw2[tid] = tw2g[tid];
    s_tw1[tid] = tw1g[tid];
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"((unsigned)__cvta_generic_to_shared(s_tmem)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (io.b.hist_next && blockIdx.x == gridDim.x - 1) fftk::update_history(io.b, tid, 512);
    if (tid == 0) {
        asm volatile("mbarrier.init.shared.b64 [%0], 1;" ::"r"(mbar) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    tm_fence_before();
    __syncthreads();
    tm_fence_after();
    const unsigned tmem = *s_tmem;
    const long long cl = blockIdx.x / C, ncl = gridDim.x / C;
    const int PP = io.D / C;                                    // branches per CTA
    const unsigned tbase = tmem + ((unsigned)(32 * (warp & 3)) << 16) + 64u * (warp >> 2);
    const TmemAcc acc{tbase};                                   // columns 0..255: the running sum
    const TmemStash stash{tbase + 256u};                        // columns 256..511: the second branch of a 128-bit gather
    float4* my_slots = scratch + (size_t)cl * C * POLY_SLOTS * 8192;
    int it = 0;
    // RRC_FFTFILT_TRACE (debug): every warp of CTA 1 % C of cluster 3 stamps clock64 at the phase boundaries of its first two
    // branches in iterations 2..7
    auto stamp = [&](int i) {
        if (trace && cl == 3 && c == 1 % C && it >= 2 && it < 2 + POLY_TRACE_IT && lane == 0)
            trace[((size_t)(it - 2) * 16 + warp) * POLY_NSTAMP + i] = clock64();
    };
    // inverse side of block (jt, jblk): B', A', store
    auto inverse = [&](long long jblk) {
        __syncwarp();
        fftk::phase_mid_bi(tid, s_tw2, sm);
        __syncthreads();
        stamp(13);
        fftk::phase_ai<false, false>(tid, jblk, io.b, s_tw1, sm);
    };
    // tune bit 0 (experiment, RRC_FFTFILT_POLY_TUNE): evict-first hint on the gather — measured SLOWER (the CTAs of a group
    // share sectors: the first reader demotes the line the second one still needs)
    unsigned long long pol_first = 0;
    if (tune & 1) asm("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol_first));
    if (tune & 32) pol_first = 1ull;                            // experiment: L1::no_allocate on the gather, default L2 policy
    // Next block's region -> L2 in 2 * PP slices, issued before and after phase B of every branch: while the transforms run
    // the memory system is idle; issued together with the gather, the prefetch traffic queues in front of the demand loads
    // (a gather that misses L2 runs at the SM's 22 B/clk DRAM rate: 24 K cycles instead of 8-10 K).
    auto prefetch = [&](long long blk, int slice) {
        const long long nb = blk + ((tune & 8) ? 2 : 1) * ncl;
        const int piece = (c * PP * 2 + slice) * 16 + warp;                         // D * 32 pieces of 512 samples
        const long long g = fftp::poly_g(io, nb, 0, 0) + (long long)piece * 512;
        if (lane == 0 && nb < nblocks && g >= 0 && g + 512 <= io.b.n_in) {
            const unsigned long long esz = io.b.in_u8 ? 2 : 8;
            const unsigned long long a = (reinterpret_cast<unsigned long long>(io.b.in) + (unsigned long long)g * esz + 15ull) & ~15ull;
            asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(a), "r"((unsigned)(512 * esz - 16)) : "memory");
        }
    };
    long long prev_blk = -1;
    for (long long blk = blk0 + cl; blk < nblocks; blk += ncl, ++it) {          // this launch's blocks: [blk0, nblocks)
        stamp(0);
        bool stashed = false;
        for (int i = 0; i < PP; ++i) {
            const int r = c * PP + i;
            float2 v[32];
            if (stashed) {
                fftp::poly_load_stash(tid, v, stash);
                stashed = false;
            } else if (i + 1 < PP && fftp::poly_pair_ok(io, blk, r)) {
                if (pol_first) poly_load_pair_dev(tid, blk, r, io, v, stash, pol_first);
                else fftp::poly_load_pair(tid, blk, r, io, v, stash);
                tm_wait_st();
                stashed = true;
            } else {
                fftp::poly_load(tid, blk, r, io, v);
            }
            if ((tune & 16) && !(tune & 2)) { prefetch(blk, 2 * i); prefetch(blk, 2 * i + 1); }
            if (i < 2) stamp(1 + 6 * i);
            fftk::phase_a_linear_compute(tid, s_tw1, v);
            if (i < 2) stamp(2 + 6 * i);
            __syncthreads();                                    // phase C / A' of the previous transform has read the buffer
            if (HRES && tid == 0) {                             // ... and the previous branch's spectrum rows: this branch's -> shared memory
                const float2* src = Hresg + (size_t)r * fftk::HRES_ELEMS;
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                asm volatile("mbarrier.arrive.expect_tx.shared.b64 _, [%0], %1;" ::"r"(mbar), "r"(fftk::HRES_ELEMS * 8) : "memory");
#pragma unroll 1
                for (int k = 0; k < 4; ++k)
                    asm volatile("cp.async.bulk.shared::cta.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                                 ::"r"(hres_a + k * (fftk::HRES_ELEMS * 2)), "l"(src + k * (fftk::HRES_ELEMS / 4)), "r"(fftk::HRES_ELEMS * 2), "r"(mbar) : "memory");
            }
            if (i < 2) stamp(3 + 6 * i);
            fftk::phase_a_linear_store(tid, sm, v);
            __syncthreads();
            if (i < 2) stamp(4 + 6 * i);
            if (!(tune & 18)) prefetch(blk, 2 * i);
            const float2* Hp = Hph + (size_t)r * fftk::N;
            fftk::phase_mid_b(tid, s_tw2, sm);
            __syncwarp();
            if (i < 2) stamp(5 + 6 * i);
            if (!(tune & 18)) prefetch(blk, 2 * i + 1);
            if constexpr (HRES) { poly_mbar_wait(mbar, hpar); hpar ^= 1u; }
            if (i < PP - 1) fftp::phase_c_acc<HRES>(tid, Hp, s_hres, sm, acc, i == 0, false);
            else if constexpr (C == 1) fftp::phase_c_acc<HRES>(tid, Hp, s_hres, sm, acc, i == 0, true);
            else phase_c_partial<HRES>(tid, Hp, s_hres, sm, acc, i == 0, my_slots + (size_t)(c * POLY_SLOTS + (it & 3)) * 8192);
            if (i < 2) stamp(6 + 6 * i);
        }
        if constexpr (C == 1) {
            inverse(blk);
        } else {
            // Slot (it & 3) of this CTA was last read by the finisher of iteration it - 4 during ITS iteration it - 3; this CTA
            // passed wait(it - 2) before writing it, i.e. after that finisher's arrive(it - 2).  Hence four slots.
            if (it >= 1) asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");      // phase it - 1: every partial sum of prev_blk is visible
            asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");                 // phase it: this CTA's partial sum of blk is written
            if (it >= 1 && (it - 1) % C == c) {
                phase_c_finish<C>(tid, sm, my_slots, (it - 1) & 3);
                inverse(prev_blk);
            }
        }
        prev_blk = blk;
        stamp(14);
    }
    if constexpr (C > 1) {
        if (it >= 1) asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
        if (it >= 1 && (it - 1) % C == c) {                     // the last block of this cluster
            phase_c_finish<C>(tid, sm, my_slots, (it - 1) & 3);
            inverse(prev_blk);
        }
    }
    tm_fence_before();
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
}

namespace {

int env_int(const char* name, int dflt) {
    const char* v = getenv(name);
    return v ? atoi(v) : dflt;
}

// Co-resident clusters of the C-CTA kernel (cudaOccupancyMaxActiveClusters; B200: 33 four-CTA clusters, 74 two-CTA ones).
template <int C, bool HRES>
int poly_clusters(rrc_fftfilt* h, int* out) {
    int& mc = h->poly.max_clusters[C == 1 ? 0 : C == 2 ? 1 : 2];
    if (mc == 0) {
        if (C > 1) {
            auto kern = fftfilt_poly_kernel<C, HRES>;
            RRC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)poly_smem(HRES)));
            cudaLaunchConfig_t cfg = {};
            cfg.blockDim = dim3(512);
            cfg.dynamicSmemBytes = poly_smem(HRES);
            cudaLaunchAttribute attr[1];
            attr[0].id = cudaLaunchAttributeClusterDimension;
            attr[0].val.clusterDim.x = C; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
            cfg.attrs = attr; cfg.numAttrs = 1;
            cfg.gridDim = dim3(sm_count(h->device) / C * C);
            RRC_CUDA(cudaOccupancyMaxActiveClusters(&mc, kern, &cfg));
            if (mc < 1) return fail(RRC_ERR_CUDA, "no %d-CTA cluster of the polyphase kernel fits on device %d", C, h->device);
        } else {
            mc = sm_count(h->device);
        }
    }
    *out = mc;
    return RRC_OK;
}

// One launch over the blocks [blk0, nblocks) on at most `ncl` clusters, partial sums in `scratch` (ncl * C * POLY_SLOTS slots).
template <int C, bool HRES>
int launch_poly(rrc_fftfilt* h, const PolyIO& io, const float2* Hph, long long blk0, long long nblocks, long long ncl, float4* scratch,
                cudaStream_t st) {
    auto kern = fftfilt_poly_kernel<C, HRES>;
    RRC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)poly_smem(HRES)));
    cudaLaunchConfig_t cfg = {};
    cfg.blockDim = dim3(512);
    cfg.dynamicSmemBytes = poly_smem(HRES);
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = C; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    cfg.gridDim = dim3((unsigned)(ncl * C));
    static const bool trace_env = getenv("RRC_FFTFILT_TRACE") != nullptr;
    const bool want_trace = trace_env && blk0 == 0;
    long long* dtrace = nullptr;
    const size_t trace_n = (size_t)POLY_TRACE_IT * 16 * POLY_NSTAMP;
    if (want_trace) { RRC_CUDA(cudaMalloc((void**)&dtrace, trace_n * 8)); RRC_CUDA(cudaMemsetAsync(dtrace, 0, trace_n * 8, st)); }
    RRC_CUDA(cudaLaunchKernelEx(&cfg, kern, io, Hph, (const float2*)h->tw1, (const float2*)h->tw2,
                                scratch, blk0, nblocks, env_int("RRC_FFTFILT_POLY_TUNE", 0), dtrace));
    count_launch();
    if (want_trace) {                                           // debug only: synchronous dump of the per-phase cycle table
        std::vector<long long> tr(trace_n);
        RRC_CUDA(cudaStreamSynchronize(st));
        RRC_CUDA(cudaMemcpy(tr.data(), dtrace, trace_n * 8, cudaMemcpyDeviceToHost));
        cudaFree(dtrace);
        static const char* names[] = {"br0 load / stash", "br0 A compute", "br0 barrier", "br0 A store + barrier", "br0 B", "br0 C",
                                      "br1 load / stash", "br1 A compute", "br1 barrier", "br1 A store + barrier", "br1 B", "br1 C",
                                      "other branches, barrier, finish, B'", "A' + store"};
        static int dumps = 0;
        if (nblocks - blk0 > ncl * 8 && dumps++ < 1) {
            for (int b = 0; b < POLY_TRACE_IT; ++b) {
                auto at = [&](int w, int i) { return tr[((size_t)b * 16 + w) * POLY_NSTAMP + i]; };
                long long t0 = at(0, 0), tend = 0;
                for (int w = 0; w < 16; ++w) { t0 = std::min(t0, at(w, 0)); tend = std::max(tend, at(w, 14)); }
                fprintf(stderr, "poly trace C=%d iter %d: total %lld cycles\n", C, b + 2, tend - t0);
                for (int p = 0; p < 14; ++p) {
                    std::vector<long long> d;
                    for (int w = 0; w < 16; ++w) { long long a = at(w, p), e = at(w, p + 1); d.push_back(a && e ? e - a : -1); }
                    std::sort(d.begin(), d.end());
                    fprintf(stderr, "   %-28s min %6lld  med %6lld  max %6lld\n", names[p], d[0], d[8], d[15]);
                }
            }
        }
    }
    return RRC_OK;
}

}  // namespace

// Geometry covered: complex streams, 2 <= deci <= 16, ceil(ntaps / deci) - 1 <= 8192 (at least half of every transform
// is kept output).  RRC_FFTFILT_NO_POLY=1 sends deci == 8 back to the fold kernel and the rest to the store predicate.
int poly_supported(const rrc_fftfilt* h, size_t deci) {
    if (h->real || deci < 2 || deci > (size_t)fftp::POLY_MAX_D) return RRC_ERR_UNSUPPORTED;
    if ((h->ntaps + deci - 2) / deci > 8192) return RRC_ERR_UNSUPPORTED;
    if (env_int("RRC_FFTFILT_NO_POLY", 0) != 0) return RRC_ERR_UNSUPPORTED;
    return RRC_OK;
}

int poly_launch(rrc_fftfilt* h, const float* in, size_t n, float* out, size_t n_out, size_t deci, size_t skip, cudaStream_t st) {
    if (n_out == 0) return RRC_ERR_UNSUPPORTED;                  // nothing to launch: the caller updates the history
    const int smod = (int)(skip % deci);
    if (h->poly.D != (int)deci) {                               // tables are per decimation ...
        for (auto& t : h->poly.Hph) { cudaFree(t); t = nullptr; }
        h->poly.D = (int)deci;
    }
    if (!h->poly.Hph[smod]) {                                   // ... and per decimation phase (skip mod deci), built on first use
        std::vector<float2> Hph;
        fftp::build_poly_tables(h->taps_host.data(), h->ntaps, (int)deci, smod, Hph);
        cudaError_t e = cudaMalloc((void**)&h->poly.Hph[smod], Hph.size() * sizeof(float2));
        if (e == cudaSuccess) e = upload_sync(h->poly.Hph[smod], Hph.data(), Hph.size() * sizeof(float2));
        if (e != cudaSuccess) return fail(RRC_ERR_CUDA, "FftFilter polyphase table upload failed: %s", cudaGetErrorString(e));
    }
    PolyIO io;
    io.b.in = reinterpret_cast<const float2*>(in);
    io.b.hist = h->hist_ext ? h->hist_ext : h->hist[h->cur];
    io.b.out = reinterpret_cast<float2*>(out);
    io.b.n_in = (long long)n;
    io.b.n_out = (long long)n_out;
    io.b.T1 = fftp::poly_T1((long long)h->ntaps, (int)deci, smod);
    io.b.V = fftk::N - io.b.T1;
    io.b.T1_total = h->T1;
    io.b.shift = 0; io.b.deci = 1; io.b.skip = 0;
    io.b.in_u8 = h->in_u8;
    io.b.real = 0;
    io.b.hist_next = h->T1 > 0 ? h->hist[h->cur ^ 1] : nullptr;
    io.b.epi = h->epi;
    io.D = (int)deci;
    io.sbase = (long long)skip - smod;
    const long long nblocks = ((long long)n_out + io.b.V - 1) / io.b.V;
    // CTAs per cluster: the largest of 4, 2, 1 that divides deci (RRC_FFTFILT_POLY_C overrides)
    int C = deci % 4 == 0 ? 4 : deci % 2 == 0 ? 2 : 1;
    const int want = env_int("RRC_FFTFILT_POLY_C", 0);
    if ((want == 1 || want == 2 || want == 4) && deci % want == 0) C = want;
    const bool hres = env_int("RRC_FFTFILT_POLY_HRES", 0) != 0;
    const float2* Hph = h->poly.Hph[smod];
    int mc = 0;
    if (C == 4) RRC_TRY((hres ? poly_clusters<4, true> : poly_clusters<4, false>)(h, &mc));
    else if (C == 2) RRC_TRY((hres ? poly_clusters<2, true> : poly_clusters<2, false>)(h, &mc));
    else RRC_TRY((hres ? poly_clusters<1, true> : poly_clusters<1, false>)(h, &mc));
    long long ncl = std::min<long long>(nblocks, mc);
    if (const int cap = env_int("RRC_FFTFILT_POLY_GROUPS", 0)) ncl = std::min<long long>(ncl, cap);   // experiment: fewer resident clusters
    // Experiment (RRC_FFTFILT_POLY_SPLIT=1, off by default): the SMs the four-CTA clusters leave idle (B200: 148 - 4 * 33 = 16)
    // run a SECOND launch of two-CTA clusters on their share of the blocks, on a side stream forked from and joined to the
    // caller's.  Measured on config 5: 8.45 ms for every share between 5 % and 8 % against 8.10 ms without — the second
    // launch slows the first one down by more than it takes off it (profiles/r02_c5_poly_split_rejected.txt).
    long long nclB = 0, nA = nblocks;
    if (C == 4 && !hres && ncl == mc && env_int("RRC_FFTFILT_POLY_SPLIT", 0) != 0) {
        int mc2 = 0;
        RRC_TRY((poly_clusters<2, false>)(h, &mc2));
        nclB = std::min<long long>((sm_count(h->device) - 4 * mc) / 2, mc2);
        const double w = 0.01 * env_int("RRC_FFTFILT_POLY_SPLIT_W", 85);           // per-SM speed of the two-CTA clusters relative to the four-CTA ones
        const long long nB = nclB > 0 ? (long long)((double)nblocks * (2.0 * nclB * w) / (4.0 * mc + 2.0 * nclB * w)) : 0;
        if (nB >= 4 * nclB && nblocks - nB >= 8 * ncl) nA = nblocks - nB; else nclB = 0;
    }
    const size_t slot_f4 = (size_t)POLY_SLOTS * 8192;
    const size_t need = ((size_t)ncl * C + (size_t)nclB * 2) * slot_f4 * sizeof(float4);
    if (C > 1 && h->poly.scratch_bytes < need) {
        cudaFree(h->poly.scratch);
        h->poly.scratch = nullptr; h->poly.scratch_bytes = 0;
        RRC_CUDA(cudaMalloc((void**)&h->poly.scratch, need));
        h->poly.scratch_bytes = need;
    }
    if (nclB > 0) {
        if (!h->poly.side) {
            RRC_CUDA(cudaStreamCreateWithFlags(&h->poly.side, cudaStreamNonBlocking));
            RRC_CUDA(cudaEventCreateWithFlags(&h->poly.ev_fork, cudaEventDisableTiming));
            RRC_CUDA(cudaEventCreateWithFlags(&h->poly.ev_join, cudaEventDisableTiming));
        }
        RRC_CUDA(cudaEventRecord(h->poly.ev_fork, st));
        RRC_CUDA(cudaStreamWaitEvent(h->poly.side, h->poly.ev_fork, 0));
    }
    int rc;
    if (C == 4) rc = hres ? launch_poly<4, true>(h, io, Hph, 0, nA, ncl, h->poly.scratch, st) : launch_poly<4, false>(h, io, Hph, 0, nA, ncl, h->poly.scratch, st);
    else if (C == 2) rc = hres ? launch_poly<2, true>(h, io, Hph, 0, nA, ncl, h->poly.scratch, st) : launch_poly<2, false>(h, io, Hph, 0, nA, ncl, h->poly.scratch, st);
    else rc = hres ? launch_poly<1, true>(h, io, Hph, 0, nA, ncl, h->poly.scratch, st) : launch_poly<1, false>(h, io, Hph, 0, nA, ncl, h->poly.scratch, st);
    if (rc != RRC_OK) return rc;
    if (nclB > 0) {
        PolyIO iob = io;
        iob.b.hist_next = nullptr;                              // the first launch writes the next history
        RRC_TRY((launch_poly<2, false>)(h, iob, Hph, nA, nblocks, nclB, h->poly.scratch + (size_t)ncl * C * slot_f4, h->poly.side));
        RRC_CUDA(cudaEventRecord(h->poly.ev_join, h->poly.side));
        RRC_CUDA(cudaStreamWaitEvent(st, h->poly.ev_join, 0));
    }
    return RRC_OK;
}

void poly_destroy(rrc_fftfilt* h) {
    for (auto& t : h->poly.Hph) cudaFree(t);
    cudaFree(h->poly.scratch);
    if (h->poly.side) { cudaStreamDestroy(h->poly.side); cudaEventDestroy(h->poly.ev_fork); cudaEventDestroy(h->poly.ev_join); }
    h->poly = rrc_poly_tables();
}

}  // namespace rrc
