// fir_tc5.cu — FirFilter (reference src/fir.rs:166-197, Fir::filter_n; block src/fir.rs:492-527) for c32 or f32 samples, real
// taps, decimation 1 and ntaps <= 65 on the 5th-generation tensor cores: tcgen05.mma (kind::f16), the tap operand and
// the accumulators in tensor memory.  Same arithmetic contract as fir_tc.cuh (block-scaled fp16 hi + lo split of
// samples and taps, the three products hi*hi + hi*lo + lo*hi accumulated in FP32), different machinery:
//
//   D[m][n] = sum_k A[m][k] * B[n][k]                        M = 128, N = 64, K = 16 per tcgen05.mma
//     A[m][k] = w'[k - m]       m < 128 outputs of a block-row, k < 192   TMEM, written once per CTA (hi and lo parts)
//     B[n][k] = z[128 n + k]    n < 64 block-rows of 128 samples          shared memory, SWIZZLE_128B K-major
//
//   * A CTA tile is 8192 outputs = 64 block-rows of 128.  The TAPS are the stationary operand: an MMA reads only the
//     64 x 32 B sample operand from shared memory and runs at the tensor pipe's floor (32 cycles; measured 32.4).
//   * The Toeplitz sample operand is never built.  A block-row of 128 staged fp16 samples is two 128-byte swizzle-atom
//     rows kept in two regions (samples 0-63 and 64-127 of every row): k-steps 0-3 read region 0, 4-7 region 1, 8-11
//     region 0 advanced by one row — the descriptor's start address + 128 B.  Measured on B200: with descriptor base
//     offset 0 the 128-byte-swizzle XOR follows the ABSOLUTE shared-memory address bits [7,10), so the advanced
//     operand needs nothing else (base offset 1 shifts the pattern by a row and reads the wrong 16-byte chunks).
//   * The accumulator lane is the output index inside the block-row, its column the block-row: tcgen05.ld.32x32b gives
//     a lane one output of 16 consecutive block-rows and a warp stores 256 contiguous bytes per block-row.
//   * One CTA per SM, warp-specialised: two producer groups of 9 warps stage alternate tiles into two plane sets (a
//     group's next tile is in flight in its registers while it waits and converts), one warp issues the MMAs into two
//     accumulator sets, four warps (one per TMEM lane quarter) drain finished accumulators.  mbarriers: full[g]
//     (288 producer arrivals) -> MMA warp; done[g] (tcgen05.commit) -> epilogue warps and the producer group that
//     reuses plane set g; accfree[g] (128 epilogue arrivals) -> MMA warp; taps (128 arrivals) -> MMA warp, once.
//
// Measured alternatives (profiles/r02_c1_tcgen05_*.txt; config 1, mma.sync kernel 58.2 us): samples as the A operand
// from shared memory, taps as B — 48 cycles per MMA (6 KB of operand reads at 128 B/clk) on the data path the staging
// stores and the global traffic also use: 72.8 us unpipelined at two CTAs per SM, 62.3 us with this pipeline.
#include <cuda_fp16.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "common.cuh"
#include "fir_tc.hpp"

namespace rrc {
namespace {

// NR = block-rows (of 128 outputs) per CTA tile: 64 (8192 outputs) or 32 (4096: half the pipeline fill)
constexpr int TC5_NPW = 9;                          // warps per producer group
template <int NR> struct Tc5Geom {
    static constexpr int BT = 128 * NR;             // outputs per tile
    static constexpr int HROWS = 2 * NR + 1;        // staged half-rows of 64 samples: BT + 64
    static constexpr int REGION = ((NR + 1) * 128 + 1023) / 1024 * 1024;      // NR + 1 rows x 128 B, rounded up to the atom
    static constexpr int PLANE = 2 * REGION;        // one component (re / im), one part (hi / lo)
    static constexpr int NLD = (HROWS + TC5_NPW - 1) / TC5_NPW;               // half-rows (float4 loads) per producer lane
    static constexpr size_t smem(bool cplx) { return 1024 + (size_t)(cplx ? 8 : 4) * PLANE + 4096; }   // 2 sets x (4 | 2) planes
    static constexpr unsigned IDESC = 0x08000010u | ((unsigned)(NR >> 3) << 17);   // kind::f16: D f32, A/B f16 K-major, N = NR, M = 128
};
constexpr int TC5_EPI0 = 2 * TC5_NPW, TC5_MMAW = TC5_EPI0 + 4;
constexpr int TC5_THREADS = (TC5_MMAW + 1) * 32;    // two producer groups, 4 epilogue warps, 1 MMA warp
constexpr int TC5_TAB = 160;                        // fp16x2 words per tap table (even / odd alignment, hi / lo part)
// Shared-memory matrix descriptor, SWIZZLE_128B K-major (cute/arch/mma_sm100_desc.hpp): start address >> 4 at [0,14),
// LBO (unused: the K extent of an MMA stays inside the atom) = 1 at [16,30), SBO = 1024 B >> 4 at [32,46), version 1 at
// [46,48), base offset 0 at [49,52), layout type 2 at [61,64).
constexpr unsigned TC5_DESC_HI = 0x40004040u;
constexpr int TC5_NSTAMP = 8, TC5_TRACE_IT = 6, TC5_TRACE_WORDS = TC5_TRACE_IT * 4 * TC5_NSTAMP + 16;

struct alignas(16) Tc5Params {                      // kernel parameters: the tap words ride in the constant bank, so
    FirTc5Args a;                                   // building the TMEM operand needs no global load (at kernel start the
    alignas(16) unsigned tab[4 * TC5_TAB];          // memory system is saturated by the first tiles: 10 K cycles measured)
};

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ unsigned tc5_desc_lo(unsigned addr) { return ((addr >> 4) & 0x3fffu) | (1u << 16); }

__device__ __forceinline__ void tc5_mma_ts(unsigned d_tmem, unsigned a_tmem, unsigned b_lo, unsigned accumulate, unsigned idesc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .b64 db;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "mov.b64 db, {%2, %5};\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], db, %3, p;\n\t}"
        :: "r"(d_tmem), "r"(a_tmem), "r"(b_lo), "r"(idesc), "r"(accumulate), "r"(TC5_DESC_HI) : "memory");
}
__device__ __forceinline__ void tc5_ld32(unsigned taddr, unsigned (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tc5_st8(unsigned taddr, const unsigned (&r)[8]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
                 :: "r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]) : "memory");
}
__device__ __forceinline__ unsigned tc5_elect() {
    unsigned pred;
    asm volatile("{\n\t.reg .pred px;\n\telect.sync _|px, 0xffffffff;\n\tselp.b32 %0, 1, 0, px;\n\t}" : "=r"(pred));
    return pred;
}
// (a, b) scaled f32 -> fp16x2 hi word and fp16x2 lo word (a in the lower half).
__device__ __forceinline__ void tc5_split2(float a, float b, unsigned& hi, unsigned& lo) {
    const __half2 h = __floats2half2_rn(a, b);
    const float2 hf = __half22float2(h);
    const __half2 l = __floats2half2_rn(a - hf.x, b - hf.y);
    hi = *reinterpret_cast<const unsigned*>(&h);
    lo = *reinterpret_cast<const unsigned*>(&l);
}
__device__ __forceinline__ void mbar_init(unsigned bar, unsigned count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(bar), "r"(count) : "memory"); }
__device__ __forceinline__ void mbar_arrive(unsigned bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(bar) : "memory"); }
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity) {
    unsigned done = 0;
    while (!done) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    }
}
__device__ __forceinline__ float4 ldg_stream(const float4* p) {      // read once: no L1 allocation
    float4 r;
    asm("ld.global.nc.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
    return r;
}

// Tile loads: half-row (pw + 9 u) of the tile, samples 2*lane, 2*lane + 1 of it.  Samples at or past `need` are zeros.
template <bool CPLX> struct Tc5Val { using T = float4; };       // what a producer lane holds of one half-row: two samples
template <> struct Tc5Val<false> { using T = float2; };
template <int NR, bool CPLX>
__device__ __forceinline__ void tc5_load(const FirTc5Args& a, long long tile, int pw, int lane, typename Tc5Val<CPLX>::T (&v)[Tc5Geom<NR>::NLD]) {
    constexpr int TC5_HROWS = Tc5Geom<NR>::HROWS, TC5_NLD = Tc5Geom<NR>::NLD;
    const long long ch = tile / a.tiles_x, tx = tile - ch * a.tiles_x;
    const long long s0 = tx * Tc5Geom<NR>::BT;
    const long long avail = a.need - s0;
    if constexpr (!CPLX) {                                 // f32 stream: samples 2*lane, 2*lane + 1 in v.x, v.y
        const float* in = reinterpret_cast<const float*>(a.in) + ch * a.in_stride + s0;
        const bool fast = avail >= 64ll * TC5_HROWS && (reinterpret_cast<uintptr_t>(in) & 7) == 0;
#pragma unroll
        for (int u = 0; u < TC5_NLD; ++u) {
            const int hr = pw + TC5_NPW * u;
            const long long s = 64ll * hr + 2 * lane;
            v[u] = make_float2(0.f, 0.f);
            if (hr < TC5_HROWS) {
                if (fast) { v[u] = __ldg(reinterpret_cast<const float2*>(in + s)); }
                else { if (s < avail) v[u].x = __ldg(in + s); if (s + 1 < avail) v[u].y = __ldg(in + s + 1); }
            }
        }
    } else {
    const float2* in = a.in + ch * a.in_stride + s0;
    const bool fast = avail >= 64ll * TC5_HROWS && (reinterpret_cast<uintptr_t>(in) & 15) == 0;
    if (fast) {
        const float4* p = reinterpret_cast<const float4*>(in) + lane;
#pragma unroll
        for (int u = 0; u < TC5_NLD; ++u) {
            const int hr = pw + TC5_NPW * u;
            v[u] = hr < TC5_HROWS ? ldg_stream(p + hr * 32) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
    } else {                                               // ragged end of a channel / 8-byte aligned span
#pragma unroll
        for (int u = 0; u < TC5_NLD; ++u) {
            const long long s = 64ll * (pw + TC5_NPW * u) + 2 * lane;
            v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (pw + TC5_NPW * u < TC5_HROWS) {
                if (s < avail) { const float2 p = __ldg(in + s); v[u].x = p.x; v[u].y = p.y; }
                if (s + 1 < avail) { const float2 q = __ldg(in + s + 1); v[u].z = q.x; v[u].w = q.y; }
            }
        }
    }
    }
}

// RRC_FIR_TC5_TRACE (debug): CTA 2 stamps clock64 per role for its tiles 2..7, plus whole-kernel stamps
#define TC5_STAMP(role, k) do { if (trace && blockIdx.x == 2 && j >= 2 && j < 2 + TC5_TRACE_IT && lane == 0) trace[((j - 2) * 4 + (role)) * TC5_NSTAMP + (k)] = clock64(); } while (0)

template <int NR, bool CPLX>
__global__ void __launch_bounds__(TC5_THREADS, 1) fir_tc5_kernel(const __grid_constant__ Tc5Params prm, long long* __restrict__ trace) {
    constexpr int TC5_HROWS = Tc5Geom<NR>::HROWS, TC5_NLD = Tc5Geom<NR>::NLD, TC5_REGION = Tc5Geom<NR>::REGION, TC5_PLANE = Tc5Geom<NR>::PLANE;
    constexpr int FIR_TC5_BT = Tc5Geom<NR>::BT;
    extern __shared__ unsigned char smem_raw[];
    const FirTc5Args& a = prm.a;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    long long* const ktr = (trace && blockIdx.x == 2) ? trace + TC5_TRACE_IT * 4 * TC5_NSTAMP : nullptr;
    if (ktr && tid == 0) ktr[0] = clock64();
    const unsigned raw = smem_u32(smem_raw);
    unsigned char* sm = smem_raw + ((1024u - (raw & 1023u)) & 1023u);      // swizzle atoms need 1024-byte alignment
    unsigned char* s_planes = sm;                                          // [2 sets][4 planes = re hi, re lo, im hi, im lo][2 regions]
    constexpr int NPL = CPLX ? 4 : 2;                                      // planes per set: (re, im) x (hi, lo) | (hi, lo)
    constexpr unsigned DSET = CPLX ? 2u * NR : (unsigned)NR;               // accumulator columns per set
    unsigned long long* s_bar = reinterpret_cast<unsigned long long*>(sm + 2 * NPL * TC5_PLANE);   // full[2], done[2], accfree[2], taps, lead
    unsigned* s_tmem = reinterpret_cast<unsigned*>(s_bar + 8);
    float* s_inv = reinterpret_cast<float*>(s_tmem + 1);                   // [4]: 1 / (tile scale * tap scale) of tile j & 3
    unsigned* s_red = reinterpret_cast<unsigned*>(s_inv + 4);              // [2 groups][3][16]
    unsigned* s_tab = reinterpret_cast<unsigned*>(sm + 2 * NPL * TC5_PLANE + 512);    // the tap words, staged for per-lane indexing (16-byte aligned)
    const unsigned planes_u = smem_u32(s_planes), bar_u = smem_u32(s_bar);
    const unsigned full_u = bar_u, done_u = bar_u + 16, accfree_u = bar_u + 32, taps_u = bar_u + 48, lead_u = bar_u + 56;

    const long long first = blockIdx.x;
    const int njobs = first < a.total_tiles ? (int)((a.total_tiles - first + gridDim.x - 1) / gridDim.x) : 0;
    if (tid == 0) {
        for (unsigned g = 0; g < 2; ++g) { mbar_init(full_u + 8 * g, 32 * TC5_NPW); mbar_init(done_u + 8 * g, 1); mbar_init(accfree_u + 8 * g, 128); }
        mbar_init(taps_u, 128);
        mbar_init(lead_u, 32 * TC5_NPW);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();                        // the only CTA-wide barrier before the tail: the mbarriers exist
    if (ktr && tid == 0) ktr[1] = clock64();

    unsigned tmem = 0;
    if (warp >= TC5_EPI0) {
        // Set-up off the producers' path: the MMA warp allocates tensor memory, the 160 threads of the MMA and epilogue
        // warps meet on named barrier 3, the epilogue warps write the tap operand into TMEM and arrive on `taps`.
        if (warp == TC5_MMAW) {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" :: "r"(smem_u32(s_tmem)) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
        } else {
            // constant bank -> shared memory with warp-uniform addresses (a lane-indexed constant load serialises per
            // distinct address: 50 K cycles when the operand was built straight from the parameters)
            const int w0 = (warp - TC5_EPI0) * TC5_TAB;
#pragma unroll 8
            for (int i = 0; i < TC5_TAB / 4; ++i) {
                const uint4 x = *reinterpret_cast<const uint4*>(&prm.tab[w0 + 4 * i]);
                if (lane == (i & 31)) *reinterpret_cast<uint4*>(&s_tab[w0 + 4 * i]) = x;
            }
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        asm volatile("bar.sync 3, 160;" ::: "memory");
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        tmem = *s_tmem;
        if (ktr && warp == TC5_MMAW && lane == 0) ktr[9] = clock64();
    }
    if (warp >= TC5_EPI0 && warp < TC5_MMAW) {
        // The tap operand: lane m, column c of part p holds (w'[2c - m], w'[2c + 1 - m]) as fp16x2.  Host tables of such
        // words: E[q] = (w'[2q], w'[2q + 1]), O[q] = (w'[2q + 1], w'[2q + 2]), q + 64 in [0, 160): a lane reads 96 in a row.
        const int m = 32 * (warp & 3) + lane;
        const int base = (m & 1) * TC5_TAB - ((m + 1) >> 1) + 64;
        for (int part = 0; part < 2; ++part)
            for (int c8 = 0; c8 < 12; ++c8) {
                unsigned r[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) r[i] = s_tab[base + part * 2 * TC5_TAB + 8 * c8 + i];
                tc5_st8(tmem + ((unsigned)(32 * (warp & 3)) << 16) + 96u * part + 8u * c8, r);
            }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        mbar_arrive(taps_u);
        if (ktr && warp == TC5_EPI0 && lane == 0) ktr[11] = clock64();
    }

    if (warp < TC5_EPI0) {
        // ================= producers: group g stages tiles j = g, g + 2, ... into plane set g
        const int g = warp >= TC5_NPW ? 1 : 0, pw = warp - g * TC5_NPW;
        typename Tc5Val<CPLX>::T v[TC5_NLD];
        if (g == 1) mbar_wait(lead_u, 0u);              // the CTA's first tile is requested first: it heads the pipeline
        if (g < njobs) tc5_load<NR, CPLX>(a, first + (long long)g * gridDim.x, pw, lane, v);
        if (g == 0) mbar_arrive(lead_u);
        if (ktr && warp == 0 && lane == 0) ktr[8] = clock64();
        unsigned char* planes = s_planes + g * NPL * TC5_PLANE;
        unsigned* red = s_red + g * 48;
        for (int j = g, u = 0; j < njobs; j += 2, ++u) {
            TC5_STAMP(g, 0);
            // ---- largest finite magnitude of the tile -> power-of-two scale (group-wide)
            float mx = 0.f;
#pragma unroll
            for (int i = 0; i < TC5_NLD; ++i) {
                mx = fmaxf(mx, fmaxf(fabsf(v[i].x), fabsf(v[i].y)));
                if constexpr (CPLX) mx = fmaxf(mx, fmaxf(fabsf(v[i].z), fabsf(v[i].w)));
            }
            unsigned wmax = __reduce_max_sync(0xffffffffu, __float_as_uint(mx));      // NaN never wins fmaxf; Inf does
            unsigned* rb = red + 16 * (u & 1);
            if (lane == 0) rb[pw] = wmax;
            asm volatile("bar.sync %0, %1;" :: "r"(1 + g), "n"(32 * TC5_NPW) : "memory");
            unsigned ex = 0;
#pragma unroll
            for (int i = 0; i < TC5_NPW; ++i) ex = max(ex, rb[i]);
            ex >>= 23;
            if (ex == 255u) {               // a non-finite sample: scale by the largest finite one (group-uniform branch)
                float m2 = 0.f;
                auto fin = [](float c) { const float q = fabsf(c); return q <= 3.4028234e38f ? q : 0.f; };
#pragma unroll
                for (int i = 0; i < TC5_NLD; ++i) {
                    m2 = fmaxf(m2, fmaxf(fin(v[i].x), fin(v[i].y)));
                    if constexpr (CPLX) m2 = fmaxf(m2, fmaxf(fin(v[i].z), fin(v[i].w)));
                }
                wmax = __reduce_max_sync(0xffffffffu, __float_as_uint(m2));
                if (lane == 0) red[32 + pw] = wmax;
                asm volatile("bar.sync %0, %1;" :: "r"(1 + g), "n"(32 * TC5_NPW) : "memory");
                ex = 0;
#pragma unroll
                for (int i = 0; i < TC5_NPW; ++i) ex = max(ex, red[32 + i]);
                ex >>= 23;
            }
            const bool scaled = ex >= 14u && ex < 255u;
            const float sc = scaled ? __uint_as_float((267u - ex) << 23) : 1.0f;      // 2^(13 - (ex - 127))
            const float isc = scaled ? __uint_as_float((ex - 13u) << 23) : 1.0f;
            TC5_STAMP(g, 1);
            if (u >= 1) mbar_wait(done_u + 8 * g, (unsigned)(u - 1) & 1u);            // the MMAs of tile j - 2 have read plane set g
            TC5_STAMP(g, 2);
            // ---- split into the four swizzled fp16 planes: a warp store is one 128-byte atom row
            {
                const unsigned col = ((unsigned)(lane & 3)) << 2, c8 = (unsigned)lane >> 2;
#pragma unroll
                for (int i = 0; i < TC5_NLD; ++i) {
                    const unsigned hr = (unsigned)(pw + TC5_NPW * i);
                    if (hr < (unsigned)TC5_HROWS) {
                        const unsigned n = hr >> 1;
                        unsigned char* p = planes + (hr & 1u) * TC5_REGION + n * 128u + ((c8 ^ (n & 7u)) << 4) + col;
                        unsigned rh, rl;
                        if constexpr (CPLX) {
                            unsigned ih, il;
                            tc5_split2(v[i].x * sc, v[i].z * sc, rh, rl);
                            tc5_split2(v[i].y * sc, v[i].w * sc, ih, il);
                            *reinterpret_cast<unsigned*>(p + 2 * TC5_PLANE) = ih;
                            *reinterpret_cast<unsigned*>(p + 3 * TC5_PLANE) = il;
                        } else {
                            tc5_split2(v[i].x * sc, v[i].y * sc, rh, rl);
                        }
                        *reinterpret_cast<unsigned*>(p) = rh;
                        *reinterpret_cast<unsigned*>(p + TC5_PLANE) = rl;
                    }
                }
            }
            if (pw == 0 && lane == 0) s_inv[j & 3] = isc * a.tap_inv_scale;
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");              // generic-proxy stores -> visible to the tensor core's reads
            mbar_arrive(full_u + 8 * g);
            TC5_STAMP(g, 3);
            if (j + 2 < njobs) tc5_load<NR, CPLX>(a, first + (long long)(j + 2) * gridDim.x, pw, lane, v);
            TC5_STAMP(g, 4);
        }
        if (ktr && lane == 0 && pw == 0) ktr[2 + g] = clock64();
    } else if (warp < TC5_MMAW) {
        // ================= epilogue: warp q holds outputs 32 q + lane of every block-row
        const int q = warp & 3;
        for (int j = 0; j < njobs; ++j) {
            const int g = j & 1, u = j >> 1;
            if (q == 0) TC5_STAMP(3, 0);
            mbar_wait(done_u + 8 * g, (unsigned)u & 1u);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            if (q == 0) TC5_STAMP(3, 1);
            const float inv = s_inv[j & 3];
            const long long tile = first + (long long)j * gridDim.x;
            const long long ch = tile / a.tiles_x, tx = tile - ch * a.tiles_x;
            const long long o0 = tx * FIR_TC5_BT;
            const long long cnt = a.out_n - o0 - (32 * q + lane);                   // this lane's output of block-row n exists for 128 n < cnt
            const bool fast = a.out_n - o0 >= FIR_TC5_BT;
#pragma unroll 1
            for (int part = 0; part < NR / 16; ++part) {
                unsigned re[16];
                const unsigned taddr = tmem + 256u + DSET * g + ((unsigned)(32 * q) << 16) + 16u * part;
                tc5_ld32(taddr, re);
                if constexpr (CPLX) {
                    unsigned im[16];
                    tc5_ld32(taddr + NR, im);
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                    float2* out = a.out + ch * a.out_stride + o0 + 32 * q + lane;
#pragma unroll
                    for (int r = 0; r < 16; ++r) {
                        const long long o = 128ll * (16 * part + r);
                        if (fast || o < cnt) out[o] = make_float2(__uint_as_float(re[r]) * inv, __uint_as_float(im[r]) * inv);
                    }
                } else {
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                    float* out = reinterpret_cast<float*>(a.out) + ch * a.out_stride + o0 + 32 * q + lane;
#pragma unroll
                    for (int r = 0; r < 16; ++r) {
                        const long long o = 128ll * (16 * part + r);
                        if (fast || o < cnt) out[o] = __uint_as_float(re[r]) * inv;
                    }
                }
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            mbar_arrive(accfree_u + 8 * g);
            if (q == 0) TC5_STAMP(3, 2);
            if (ktr && q == 0 && lane == 0 && j < 2) ktr[6 + j] = clock64();
        }
        if (ktr && q == 0 && lane == 0) ktr[4] = clock64();
    } else {
        // ================= MMA warp: one elected lane issues a tile's 6 * KS MMAs back to back (UTCHMMA), commit -> done[g]
        constexpr unsigned PL16 = TC5_PLANE / 16, RG16 = TC5_REGION / 16;
        mbar_wait(taps_u, 0u);
        for (int j = 0; j < njobs; ++j) {
            const int g = j & 1, u = j >> 1;
            TC5_STAMP(2, 0);
            mbar_wait(full_u + 8 * g, (unsigned)u & 1u);
            TC5_STAMP(2, 1);
            if (u >= 1) mbar_wait(accfree_u + 8 * g, (unsigned)(u - 1) & 1u);          // tile j - 2 has left accumulator set g
            TC5_STAMP(2, 2);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            if (tc5_elect()) {
                const unsigned pb = tc5_desc_lo(planes_u + (unsigned)g * NPL * TC5_PLANE);
                const unsigned d = tmem + 256u + DSET * (unsigned)g;
                constexpr unsigned idesc = Tc5Geom<NR>::IDESC;
#pragma unroll
                for (int s = 0; s < 12; ++s) {
                    if (s < a.KS) {
                        // k-block 0: region 0; 1: region 1; 2: region 0 advanced by one 128-byte row (offsets in 16-byte units)
                        const unsigned kb = (unsigned)s >> 2;
                        const unsigned xb = pb + (kb == 1 ? RG16 : 0u) + (kb == 2 ? 8u : 0u) + ((unsigned)s & 3u) * 2u;
                        const unsigned a_hi = tmem + 8u * s, a_lo = a_hi + 96u;
                        const unsigned acc = s ? 1u : 0u;
                        tc5_mma_ts(d, a_hi, xb, acc, idesc);                       // re: taps hi * samples hi
                        if constexpr (CPLX) tc5_mma_ts(d + NR, a_hi, xb + 2 * PL16, acc, idesc);       // im
                        tc5_mma_ts(d, a_hi, xb + PL16, 1u, idesc);                 // taps hi * samples lo
                        if constexpr (CPLX) tc5_mma_ts(d + NR, a_hi, xb + 3 * PL16, 1u, idesc);
                        tc5_mma_ts(d, a_lo, xb, 1u, idesc);                        // taps lo * samples hi
                        if constexpr (CPLX) tc5_mma_ts(d + NR, a_lo, xb + 2 * PL16, 1u, idesc);
                    }
                }
                asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(done_u + 8 * g) : "memory");
            }
            __syncwarp();
            TC5_STAMP(2, 3);
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == TC5_MMAW) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" :: "r"(tmem) : "memory");
    if (ktr && tid == 0) ktr[5] = clock64();
}

}  // namespace

size_t fir_tc5_tab_words() { return 4 * TC5_TAB; }

// Tap tables of fir_tc5_kernel for the scaled taps' fp16 parts hi[j], lo[j] (bit patterns), j < ntaps <= 65:
// [part hi / lo][alignment even / odd][q + 64] = (p[2q + odd], p[2q + odd + 1]) as fp16x2, zero outside the taps.
void fir_tc5_build_tab(const unsigned short* hi, const unsigned short* lo, size_t ntaps, unsigned* tab) {
    auto at = [&](const unsigned short* p, long long j) -> unsigned { return (j < 0 || j >= (long long)ntaps) ? 0u : p[j]; };
    for (int part = 0; part < 2; ++part)
        for (int odd = 0; odd < 2; ++odd)
            for (int i = 0; i < TC5_TAB; ++i) {
                const long long q = (long long)i - 64;
                const unsigned short* p = part ? lo : hi;
                tab[(part * 2 + odd) * TC5_TAB + i] = at(p, 2 * q + odd) | (at(p, 2 * q + odd + 1) << 16);
            }
}

// Block-rows per CTA tile for a launch of `tiles8192` tiles of 8192 outputs: 32-row tiles (4096 outputs) halve the pipeline
// fill (first tile stored after 10 K instead of 17 K cycles) and cost 4 % in the steady state (measured):
// ahead below about 12 tiles per SM, behind above.  RRC_FIR_TC5_NR = 32 / 64 overrides (experiments).
int fir_tc5_rows(long long tiles8192, int device, bool real_stream) {
    if (real_stream) return 64;                         // f32 streams: 128-row tiles (the same bytes per tile) measured 5-8 % slower
    if (const char* e = getenv("RRC_FIR_TC5_NR")) { const int v = atoi(e); if (v == 32 || v == 64) return v; }
    return tiles8192 < 12ll * sm_count(device) ? 32 : 64;
}

int fir_tc5_launch(int device, const FirTc5Args& a, const unsigned* tab_host, cudaStream_t st) {
    static bool ready[16] = {};
    const int dv = (device < 0 || device >= 16) ? 0 : device;
    if (!ready[dv]) {
        RRC_CUDA(cudaFuncSetAttribute(fir_tc5_kernel<64, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Tc5Geom<64>::smem(true)));
        RRC_CUDA(cudaFuncSetAttribute(fir_tc5_kernel<32, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Tc5Geom<32>::smem(true)));
        RRC_CUDA(cudaFuncSetAttribute(fir_tc5_kernel<64, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Tc5Geom<64>::smem(false)));
        ready[dv] = true;
    }
    Tc5Params prm;
    prm.a = a;
    memcpy(prm.tab, tab_host, sizeof prm.tab);
    static const bool want_trace = getenv("RRC_FIR_TC5_TRACE") != nullptr;
    long long* dtrace = nullptr;
    if (want_trace) { RRC_CUDA(cudaMalloc((void**)&dtrace, TC5_TRACE_WORDS * 8)); RRC_CUDA(cudaMemsetAsync(dtrace, 0, TC5_TRACE_WORDS * 8, st)); }
    const unsigned grid = (unsigned)std::min<long long>(a.total_tiles, sm_count(device));
    if (a.real_stream) {
        if (a.nr == 64) fir_tc5_kernel<64, false><<<grid, TC5_THREADS, Tc5Geom<64>::smem(false), st>>>(prm, dtrace);
        else return fail(RRC_ERR_INVALID, "fir_tc5: %d block-rows per tile (f32 streams: 64)", a.nr);
    } else {
        if (a.nr == 32) fir_tc5_kernel<32, true><<<grid, TC5_THREADS, Tc5Geom<32>::smem(true), st>>>(prm, dtrace);
        else if (a.nr == 64) fir_tc5_kernel<64, true><<<grid, TC5_THREADS, Tc5Geom<64>::smem(true), st>>>(prm, dtrace);
        else return fail(RRC_ERR_INVALID, "fir_tc5: %d block-rows per tile (c32 streams: 32 or 64)", a.nr);
    }
    RRC_CHECK_LAUNCH();
    count_launch();
    if (want_trace) {                                           // debug only: CTA 2's stamps, cycles relative to the earliest one
        std::vector<long long> tr(TC5_TRACE_WORDS);
        RRC_CUDA(cudaStreamSynchronize(st));
        RRC_CUDA(cudaMemcpy(tr.data(), dtrace, TC5_TRACE_WORDS * 8, cudaMemcpyDeviceToHost));
        cudaFree(dtrace);
        static int dumps = 0;
        if (a.total_tiles > 148 * 10 && dumps++ < 2) {
            long long t0 = 0;
            for (long long x : tr) if (x && (!t0 || x < t0)) t0 = x;
            static const char* role[] = {"producers 0: start, scale known, plane set free, staged + arrived, next loads issued",
                                         "producers 1: (same)", "MMA warp: start, planes full, accumulators free, issued + committed",
                                         "epilogue warp 0: start, MMAs done, stored + arrived"};
            for (int j = 0; j < TC5_TRACE_IT; ++j)
                for (int r = 0; r < 4; ++r) {
                    const long long* p = &tr[(size_t)(j * 4 + r) * TC5_NSTAMP];
                    if (!p[0]) continue;
                    fprintf(stderr, "tc5 tile %d %-14.14s", j + 2, role[r]);
                    for (int k = 0; k < TC5_NSTAMP && p[k]; ++k) fprintf(stderr, " %7lld", p[k] - t0);
                    fprintf(stderr, "\n");
                }
            for (int r = 0; r < 4; ++r) fprintf(stderr, "   %s\n", role[r]);
            const long long* kt = &tr[(size_t)TC5_TRACE_IT * 4 * TC5_NSTAMP];
            fprintf(stderr, "tc5 kernel (CTA 2, %lld tiles in all): start %lld, mbarriers ready %lld, first loads issued %lld, TMEM allocated %lld, taps in TMEM %lld, "
                            "first two tiles stored %lld %lld, producers done %lld %lld, epilogue done %lld, exit %lld\n",
                    (long long)a.total_tiles, kt[0] - t0, kt[1] - t0, kt[8] - t0, kt[9] - t0, kt[11] - t0, kt[6] - t0, kt[7] - t0, kt[2] - t0, kt[3] - t0, kt[4] - t0, kt[5] - t0);
        }
    }
    return RRC_OK;
}

}  // namespace rrc
