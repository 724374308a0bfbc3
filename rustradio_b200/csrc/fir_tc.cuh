// fir_tc.cuh — tensor-core form of the real-tap Complex<f32> FIR (configs 1 and 3).
//
// Same result as Fir<Complex>::filter_n (src/fir.rs:166-197) for taps with im == 0:
//     y[o] = sum_{j<T} z[o*D + j] * w[j],   w[j] = taps[T-1-j]
// written as a Toeplitz-block matrix product.  A block-row b owns R = 8*NTILE consecutive outputs:
//     Y[b][n] = sum_k A[b][k] * B[k][n],  A[b][k] = z[b*R*D + k],  B[k][n] = w[k - n*D] (0 outside 0..T-1)
// with K = (R-1)*D + T rounded up to a multiple of 16.  A is never materialised: its rows are windows
// of the staged input, 16-byte row segments at arbitrary 16-byte-aligned addresses, which is exactly
// what `ldmatrix` takes.  The real and imaginary parts of 8 block-rows are the 16 rows of one
// mma.m16n8k16 A tile, so a thread's accumulators are the (re, im) of two consecutive outputs.
//
// Precision: block-scaled fp16x3.  Per tile the samples are multiplied by the power of two that puts the
// tile's largest |re|,|im| in [2^13, 2^14) (taps likewise, once, on the host), then split x = hi + lo with
// both halves fp16: 22 significant bits, |x - hi - lo| <= max(2^-22 |x|, 2^-25).  The product is
// hi*hi + (hi*lo + lo*hi), FP32 accumulation in two separate accumulators; the dropped lo*lo term is
// <= 2^-22 of the product.  That is FP32-class: measured rel-RMS against the f64 convolution 1e-7..4e-7,
// next to the sequential f32 loop's 1e-7..2e-7 (bar 1e-5).  RRC_FIR_NO_TENSOR keeps the FP32 kernels.
// Tiles whose largest magnitude is below 2^-113 are not scaled; non-finite samples are left out of the maximum, so they
// only affect the outputs whose window contains them (tc_tile_exp).
//
// Pipeline: every WARP is an independent worker with its own tile, its own fp16 planes in shared memory and no
// CTA barrier after start-up (24 warps per SM drift out of phase, so one warp's global-load wait overlaps the
// others' split / product / store phases).  A tile's raw f32 samples never touch shared memory: each lane
// loads NLD float4 into registers (all in flight together), the tile maximum is one warp reduction on those
// registers, and only the fp16 planes are stored (8 B/sample of shared-memory traffic).  Measured on the way
// (profiles/r01_c{1,3}_tc_v{2,3}_ncu_summary.txt): CTA-wide tiles with a TMA-staged raw copy + max pass + split
// pass ran at 77 % shared-memory pipe utilisation; CTA-wide register-staged tiles lost 15-25 % of their time
// in the three barriers per tile.  B fragments are loaded once per CTA.
// SASS: HMMA.16816.F32 + LDSM.16.M88.4 (legacy warp-level tensor path: measured 0.46 mma/clk/SM on B200,
// tools/microbench/hmma_rate.cu, i.e. 7.5x the FP32 FMA lane rate).  The conversion costs ~20 instructions per
// sample, so the path pays off when a sample feeds many taps: the planner takes it for ntaps/deci >= 32
// (config 1: 64) and leaves decimating short filters (config 3: 25.5) on the packed-FP32 kernel, which
// measured faster there.
#pragma once
#include <cuda_fp16.h>

#include <cstdint>

#include "fir_common.cuh"
#include "fir_tc.hpp"

namespace rrc {

// NLD = float4 (two samples) a lane holds per tile: 9 -> warp tiles of <= 576 samples, 14 -> <= 896.

__device__ __forceinline__ void ldsm4(unsigned (&r)[4], unsigned addr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
// not volatile: the scheduler may move the products behind the next k-step's ldmatrix
__device__ __forceinline__ void mma_f16(float (&c)[4], const unsigned (&a)[4], unsigned b0, unsigned b1) {
    asm("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// u8 I/Q tile load (src/rtlsdr_decode.rs:35-43 fused, bit-identical to decoding first): the lane's sample pairs
// e = lane + 32*u as one 32-bit word each (16-bit loads when the span is only 2-byte aligned or ragged); samples beyond
// `avail` read as the byte pair (127, 127), which decodes to exactly 0.
template <int NLD>
__device__ __forceinline__ void tc_load_u8(const unsigned short* __restrict__ in8, long long avail, int npairs, int lane,
                                           float4 (&v)[NLD]) {
    const bool al4 = (reinterpret_cast<unsigned long long>(in8) & 3ull) == 0 && avail >= 2ll * npairs;
#pragma unroll
    for (int u = 0; u < NLD; ++u) {
        const int e = lane + u * 32;
        unsigned w = 0x7f7f7f7fu;
        if (e < npairs) {
            if (al4) {
                w = __ldg(reinterpret_cast<const unsigned*>(in8) + e);
            } else {
                const long long s = 2ll * e;
                const unsigned w0 = s < avail ? in8[s] : 0x7f7fu;
                const unsigned w1 = s + 1 < avail ? in8[s + 1] : 0x7f7fu;
                w = w0 | (w1 << 16);
            }
        }
        const float2 p = decode_iq(w & 0xffffu), q = decode_iq(w >> 16);
        v[u] = make_float4(p.x, p.y, q.x, q.y);
    }
}

// Exponent field of the tile's largest |component| (warp-wide).  A non-finite sample must not switch the scaling off
// for its whole tile (the other samples would overflow fp16 unscaled): then the largest FINITE magnitude is used, the
// non-finite sample becomes an fp16 Inf/NaN and only the outputs whose window contains it are non-finite, like the reference's.
template <int NLD>
__device__ __forceinline__ unsigned tc_tile_exp(const float4 (&v)[NLD]) {
    float mx = 0.f;
#pragma unroll
    for (int u = 0; u < NLD; ++u)
        mx = fmaxf(fmaxf(mx, fmaxf(fabsf(v[u].x), fabsf(v[u].y))), fmaxf(fabsf(v[u].z), fabsf(v[u].w)));
    unsigned ex = __reduce_max_sync(0xffffffffu, __float_as_uint(mx)) >> 23;     // NaN never wins fmaxf; Inf does
    if (ex == 255u) {
        float m2 = 0.f;
        auto fin = [](float c) { const float a = fabsf(c); return a <= 3.4028234e38f ? a : 0.f; };
#pragma unroll
        for (int u = 0; u < NLD; ++u)
            m2 = fmaxf(fmaxf(m2, fmaxf(fin(v[u].x), fin(v[u].y))), fmaxf(fin(v[u].z), fin(v[u].w)));
        ex = __reduce_max_sync(0xffffffffu, __float_as_uint(m2)) >> 23;
    }
    return ex;
}

// (a, b) scaled f32 -> fp16x2 hi word and fp16x2 lo word (a in the lower half).
__device__ __forceinline__ void split2(float a, float b, unsigned& hi, unsigned& lo) {
    const __half2 h = __floats2half2_rn(a, b);
    const float2 hf = __half22float2(h);
    const __half2 l = __floats2half2_rn(a - hf.x, b - hf.y);
    hi = *reinterpret_cast<const unsigned*>(&h);
    lo = *reinterpret_cast<const unsigned*>(&l);
}

// Shared-memory layout (bytes):  [bfrag KS*NTILE*512] then per warp WB = [4 planes PL*2 each][ytile (BT + 2) float2, DEMOD]
// Plane p = 2*(im?) + (lo?).  Sample s of the tile lives at element s + (s / RS) * PAD of its plane.
template <int NTILE, bool DEMOD, int NLD>
__global__ void __launch_bounds__(FIR_TC_THREADS, 3) fir_tc_kernel(const FirTcArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int R = 8 * NTILE;
    constexpr int NW = FIR_TC_THREADS / 32;
    const int BT = a.NM * 8 * R;                           // outputs per warp tile
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    uint4* s_b = reinterpret_cast<uint4*>(smem_raw);
    unsigned char* s_planes = smem_raw + (size_t)a.KS * NTILE * 512 + (size_t)warp * a.WB;
    float2* s_y = reinterpret_cast<float2*>(s_planes + (size_t)a.PL * 8);
    for (int i = tid; i < a.KS * NTILE * 32; i += FIR_TC_THREADS) s_b[i] = a.bfrag[i];
    __syncthreads();                                       // the only CTA barrier

    const int mat = lane >> 3, rr = lane & 7;
    const unsigned planes_u32 = (unsigned)__cvta_generic_to_shared(s_planes);
    const unsigned plane_bytes = (unsigned)a.PL * 2u;
    // hi plane of re (mat even) or im (mat odd); the lo plane follows it
    const unsigned hi_base = planes_u32 + (unsigned)(mat & 1) * 2u * plane_bytes;
    const unsigned khalf = (unsigned)(mat >> 1) * 8u;
    const uint4* bp = s_b + lane;
    const int plw = a.PL >> 1;                             // 32-bit words per plane
    unsigned* pl0 = reinterpret_cast<unsigned*>(s_planes);
    const int npairs = a.L >> 1;
    auto seg_off = [&](int ks) -> unsigned {               // byte offset of this lane's 16-byte row segment at k-step ks
        const unsigned k = 16u * (unsigned)ks + khalf;
        return 2u * (k + __umulhi(k, a.magic) * (unsigned)a.PAD);
    };

    const long long nworkers = (long long)gridDim.x * NW;
    for (long long id = (long long)blockIdx.x * NW + warp; id < a.total_tiles; id += nworkers) {
        const long long ch = (long long)((unsigned)id / (unsigned)a.tiles_x);      // total_tiles < 2^31 (checked on the host): 32-bit division
        const long long ob = (id - ch * a.tiles_x) * BT;
        // ---- A. the lane's samples 2*(lane + 32*u), +1 (zero beyond the channel's `need` samples) ----
        float4 v[NLD];
        {
            const float2* __restrict__ in = a.in + ch * a.in_stride + ob * a.deci;
            const long long avail = a.need - ob * a.deci;
            if (a.in_u8) {
                tc_load_u8<NLD>(reinterpret_cast<const unsigned short*>(a.in) + ch * a.in_stride + ob * a.deci, avail, npairs, lane, v);
            } else if (avail >= a.L && (reinterpret_cast<unsigned long long>(in) & 15ull) == 0) {
#pragma unroll
                for (int u = 0; u < NLD; ++u) {
                    const int e = lane + u * 32;
                    v[u] = e < npairs ? __ldg(reinterpret_cast<const float4*>(in) + e) : make_float4(0.f, 0.f, 0.f, 0.f);
                }
            } else {                                       // ragged end of a channel / 8-byte aligned span
#pragma unroll
                for (int u = 0; u < NLD; ++u) {
                    const int e = lane + u * 32;
                    const long long s = 2ll * e;
                    v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (e < npairs) {
                        if (s < avail) { const float2 p = __ldg(in + s); v[u].x = p.x; v[u].y = p.y; }
                        if (s + 1 < avail) { const float2 q = __ldg(in + s + 1); v[u].z = q.x; v[u].w = q.y; }
                    }
                }
            }
        }
        // ---- B. largest magnitude of the tile -> power-of-two scale ----
        const unsigned ex = tc_tile_exp<NLD>(v);
        const bool scaled = ex >= 14u && ex < 255u;
        const float sc = scaled ? __uint_as_float((267u - ex) << 23) : 1.0f;          // 2^(13 - (ex - 127))
        const float isc = scaled ? __uint_as_float((ex - 13u) << 23) : 1.0f;
        const float inv = isc * a.tap_inv_scale;
        // ---- C. split into the four fp16 planes ----
#pragma unroll
        for (int u = 0; u < NLD; ++u) {
            const int e = lane + u * 32;
            if (e < npairs) {
                unsigned rh, rl, ih, il;
                split2(v[u].x * sc, v[u].z * sc, rh, rl);
                split2(v[u].y * sc, v[u].w * sc, ih, il);
                const unsigned s = 2u * (unsigned)e;
                const unsigned w = (s + __umulhi(s, a.magic) * (unsigned)a.PAD) >> 1;
                pl0[w] = rh;
                pl0[w + plw] = rl;
                pl0[w + 2 * plw] = ih;
                pl0[w + 3 * plw] = il;
            }
        }
        __syncwarp();
        // ---- D. boundary output for the demod epilogue, y[ob + BT] (f32 taps, samples as hi + lo) ----
        if constexpr (DEMOD) {
            const __half* p16 = reinterpret_cast<const __half*>(s_planes);
            float re = 0.f, im = 0.f;
            for (int j = lane; j < a.ntaps; j += 32) {
                const unsigned s = (unsigned)(BT * a.deci + j);
                const unsigned e = s + __umulhi(s, a.magic) * (unsigned)a.PAD;
                const float w = __ldg(a.taps_rev + j);
                re = fmaf(__half2float(p16[e]) + __half2float(p16[e + a.PL]), w, re);
                im = fmaf(__half2float(p16[e + 2 * a.PL]) + __half2float(p16[e + 3 * a.PL]), w, im);
            }
#pragma unroll
            for (int d = 16; d; d >>= 1) {
                re += __shfl_xor_sync(0xffffffffu, re, d);
                im += __shfl_xor_sync(0xffffffffu, im, d);
            }
            if (lane == 0) s_y[BT] = make_float2(re * isc, im * isc);
        }
        // ---- E. Toeplitz product, one m-tile (8 block-rows, re and im) at a time, A fragments one k-step ahead ----
        float2* __restrict__ outc = reinterpret_cast<float2*>(a.out) + ch * a.out_stride + ob;
        const long long left_c = a.out_n - ob;
        const bool st16 = (reinterpret_cast<unsigned long long>(outc) & 15ull) == 0;
        for (int mt = 0; mt < a.NM; ++mt) {
            float acc[NTILE][4], cor[NTILE][4];
#pragma unroll
            for (int n = 0; n < NTILE; ++n)
#pragma unroll
                for (int q = 0; q < 4; ++q) { acc[n][q] = 0.f; cor[n][q] = 0.f; }
            const unsigned row_addr = hi_base + 2u * (unsigned)((mt * 8 + rr) * (a.RS + a.PAD));
            unsigned ah[2][4], al[2][4];
            {
                const unsigned so = seg_off(0);
                ldsm4(ah[0], row_addr + so);
                ldsm4(al[0], row_addr + so + plane_bytes);
            }
            auto step = [&](int ks, int cur) {
                if (ks + 1 < a.KS) {
                    const unsigned so = seg_off(ks + 1);
                    ldsm4(ah[cur ^ 1], row_addr + so);
                    ldsm4(al[cur ^ 1], row_addr + so + plane_bytes);
                }
#pragma unroll
                for (int n = 0; n < NTILE; ++n) {
                    const uint4 b = bp[(ks * NTILE + n) * 32];
                    mma_f16(acc[n], ah[cur], b.x, b.y);
                    mma_f16(cor[n], al[cur], b.x, b.y);
                    mma_f16(cor[n], ah[cur], b.z, b.w);
                }
            };
            int ks = 0;
            for (; ks + 1 < a.KS; ks += 2) { step(ks, 0); step(ks + 1, 1); }
            if (ks < a.KS) step(ks, 0);
            // lane holds (re, im) of outputs 2t, 2t+1 of block-row g
            const int g = lane >> 2, t = lane & 3;
#pragma unroll
            for (int n = 0; n < NTILE; ++n) {
                const int o = (mt * 8 + g) * R + n * 8 + 2 * t;
                const float4 y = make_float4((acc[n][0] + cor[n][0]) * inv, (acc[n][2] + cor[n][2]) * inv,
                                             (acc[n][1] + cor[n][1]) * inv, (acc[n][3] + cor[n][3]) * inv);
                if constexpr (DEMOD) {
                    *reinterpret_cast<float4*>(s_y + o) = y;
                } else {
                    // 64 contiguous bytes per block-row and n-tile: whole 128-byte lines between a warp's n-tiles
                    if (st16 && o + 1 < left_c) {
                        *reinterpret_cast<float4*>(outc + o) = y;
                    } else {
                        if (o < left_c) outc[o] = make_float2(y.x, y.y);
                        if (o + 1 < left_c) outc[o + 1] = make_float2(y.z, y.w);
                    }
                }
            }
        }
        // ---- F. demod epilogue ----
        if constexpr (DEMOD) {
            __syncwarp();
            float* __restrict__ out = reinterpret_cast<float*>(a.out) + ch * a.out_stride + ob;
            const long long left = a.out_n - 1 - ob;       // demod outputs from this tile on
            for (int o = lane; o < BT; o += 32)
                if (o < left) out[o] = demod_pair(s_y[o], s_y[o + 1], a.gain);
        }
        __syncwarp();                                      // the warp's planes / ytile are free again
    }
}

// ---- deci 1, 2 and 4 specialisation: every A fragment is loaded from shared memory ONCE per warp tile -------
// (written for deci == 1 below; for deci D in {2, 4} a warp tile has S = 8/D m-tiles, block-row b = j + S*r, so the row
// pitch stays 64 samples and the whole layout is unchanged; the walk position of (m-tile j, k-step ks) is
// q = D*j + 2*ks, even positions only, and a tile gives 64*S outputs from the same 512 + 16*KS staged samples)
// The generic kernel above loads two ldmatrix.x4 (hi, lo) plus one B fragment per three mma: 12 shared-memory
// wavefronts per 3 mma, and ncu showed the L1 data pipe at 80 % with the tensor pipe at 18 % on config 1
// (profiles/r01_c1_tc_v4_ncu_summary.txt).  For deci == 1 (R = 8 outputs per block-row, RS = 8) the Toeplitz rows
// of different m-tiles are the same samples shifted by whole 8-sample units: with block-row b = j + 8*r
// (m-tile j < 8, row r < 8) the A fragment of (m-tile j, k-step ks) is the pair of 8-sample "half fragments"
// H(j + 2*ks), H(j + 2*ks + 1), where H(a) = rows {64*r + 8*a .. +8} of the re/im x hi/lo planes — ONE
// ldmatrix.x4.  Walking p = 0 .. 7 + 2*(KS-1) feeds all 8*KS (m-tile, k-step) pairs; each position loads its
// A operand {H(p), H(p+1)} for the hi and for the lo planes (two ldmatrix.x4 whose destination quads are the
// mma operands as they are — a sliding two-entry window of single half fragments would save half of these
// loads but needs 8 register moves per position to assemble the operands): 32 ldmatrix per 120 mma for
// 64 taps instead of 80 ldmatrix + 40 B-fragment loads.
// The loop is fully unrolled (KS is a template parameter), so every shared-memory offset is an immediate; a B
// fragment is alive for 8 consecutive positions, so a five-entry register window (one LDS.128 per k-step and
// tile) holds all of them for any KS (instantiated up to 20: 7*deci + ntaps <= 320); and the padded plane layout
// (8 fp16 after every 64 samples -> 144-byte row stride) needs no index arithmetic anywhere:
// the lane's sample pair u lands at word lane + 36*u.
// Geometry of fir_tc1_kernel<KS>: staged samples, float4 per lane, 32-bit words per plane (chunks of 64 + 8 fp16),
// bytes of planes / ytile per warp, dynamic shared memory per CTA.
__host__ __device__ constexpr int fir_tc1_L(int KS) { return 512 + 16 * KS; }                // 504 + 16*KS for the product, 512 + ntaps for the demod boundary output
__host__ __device__ constexpr int fir_tc1_nld(int KS) { return (fir_tc1_L(KS) / 2 + 31) / 32; }
__host__ __device__ constexpr int fir_tc1_plw(int KS) { return (fir_tc1_L(KS) + 63) / 64 * 36; }
constexpr int FIR_TC1_YB = (FIR_TC1_BT + 2) * 8;
__host__ __device__ constexpr int fir_tc1_wb(int KS, bool demod) { return 4 * fir_tc1_plw(KS) * 4 + (demod ? FIR_TC1_YB : 0); }
__host__ __device__ constexpr size_t fir_tc1_smem(int KS, bool demod) { return (size_t)KS * 512 + (size_t)(FIR_TC_THREADS / 32) * fir_tc1_wb(KS, demod); }

template <int KS, bool DEMOD, bool U8, int D>
__global__ void __launch_bounds__(FIR_TC_THREADS, 3) fir_tc1_kernel(const FirTc1Args a) {
    static_assert(D == 1 || D == 2 || D == 4 || D == 8, "fir_tc1_kernel: deci 1, 2, 4 or 8");
    constexpr int S = 8 / D;                               // m-tiles per warp tile: block-row b = j + S*r keeps the row pitch at 64 samples
    constexpr int QL = (8 - D) + 2 * (KS - 1);             // last walk position; q = D*j + 2*ks
    constexpr int QS = D == 1 ? 1 : 2;                     // even decimations only visit even positions
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int NW = FIR_TC_THREADS / 32;
    constexpr int BT = 64 * S;                             // outputs per warp tile (512 input samples + halo)
    constexpr int L = fir_tc1_L(KS);                       // staged samples per tile
    constexpr int NP = L / 2;                              // sample pairs
    constexpr int NLD = fir_tc1_nld(KS);
    constexpr int PLW = fir_tc1_plw(KS);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint4* s_b = reinterpret_cast<const uint4*>(smem_raw) + lane;    // B fragments: [KS][32] uint4, this lane's column
    unsigned char* s_planes = smem_raw + (size_t)KS * 512 + (size_t)warp * fir_tc1_wb(KS, DEMOD);
    float2* s_y = reinterpret_cast<float2*>(s_planes + 4 * PLW * 4);
    unsigned* pl0 = reinterpret_cast<unsigned*>(s_planes);
    for (int i = threadIdx.x; i < KS * 32; i += FIR_TC_THREADS) reinterpret_cast<uint4*>(smem_raw)[i] = __ldg(a.bfrag + i);
    __syncthreads();                                       // the only CTA barrier

    // ldmatrix lane address for the A operand of walk position p: matrix (lane >> 3) = {re @p, im @p, re @p+1, im @p+1}
    // of the hi planes (the lo plane of each follows it), row r = lane & 7 at 144 bytes.  Half fragment q sits at byte
    // 16*q + 16*(q >> 3) of its row (8 fp16 of padding after every 64 samples), so lanes of the "@p+1" matrices need
    // 16 bytes more, and 32 when p + 1 crosses a chunk (p % 8 == 7): two lane bases, every other offset an immediate.
    const int mat = lane >> 3;
    const unsigned lane_addr = (unsigned)__cvta_generic_to_shared(s_planes) + (unsigned)(mat & 1) * (2u * PLW * 4u) +
                               (unsigned)(lane & 7) * 144u + (unsigned)(mat >> 1) * 16u;
    const unsigned lane_addr7 = lane_addr + (unsigned)(mat >> 1) * 16u;

    const long long nworkers = (long long)gridDim.x * NW;
    for (long long id = (long long)blockIdx.x * NW + warp; id < a.total_tiles; id += nworkers) {
        const long long ch = (long long)((unsigned)id / (unsigned)a.tiles_x);      // total_tiles < 2^31 (checked on the host): 32-bit division
        const long long ob = (id - ch * a.tiles_x) * BT;
        float4 v[NLD];
        {
            const float2* __restrict__ in = a.in + ch * a.in_stride + ob * D;
            const long long avail = a.need - ob * D;
            if constexpr (U8) {
                tc_load_u8<NLD>(reinterpret_cast<const unsigned short*>(a.in) + ch * a.in_stride + ob * D, avail, NP, lane, v);
            } else if (avail >= L && (reinterpret_cast<unsigned long long>(in) & 15ull) == 0) {
#pragma unroll
                for (int u = 0; u < NLD; ++u) {
                    const int e = lane + u * 32;
                    v[u] = e < NP ? __ldg(reinterpret_cast<const float4*>(in) + e) : make_float4(0.f, 0.f, 0.f, 0.f);
                }
            } else {
#pragma unroll
                for (int u = 0; u < NLD; ++u) {
                    const int e = lane + u * 32;
                    const long long s = 2ll * e;
                    v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (e < NP) {
                        if (s < avail) { const float2 p = __ldg(in + s); v[u].x = p.x; v[u].y = p.y; }
                        if (s + 1 < avail) { const float2 q = __ldg(in + s + 1); v[u].z = q.x; v[u].w = q.y; }
                    }
                }
            }
        }
        {   // the warp's NEXT tile -> L2 (one bulk prefetch), so that its loads are L2 hits one tile from now
            const long long nid = id + nworkers;
            if (!U8 && lane == 0 && nid < a.total_tiles) {
                const long long nch = (long long)((unsigned)nid / (unsigned)a.tiles_x);
                const long long nob = (nid - nch * a.tiles_x) * BT;
                const float2* nin = a.in + nch * a.in_stride + nob * D;
                if (a.need - nob * D >= L && (reinterpret_cast<unsigned long long>(nin) & 15ull) == 0)
                    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(nin), "r"(L * 8) : "memory");
            }
        }
        const unsigned ex = tc_tile_exp<NLD>(v);
        const bool scaled = ex >= 14u && ex < 255u;
        const float sc = scaled ? __uint_as_float((267u - ex) << 23) : 1.0f;
        const float isc = scaled ? __uint_as_float((ex - 13u) << 23) : 1.0f;
        const float inv = isc * a.tap_inv_scale;
#pragma unroll
        for (int u = 0; u < NLD; ++u) {
            if (lane + u * 32 < NP) {
                unsigned rh, rl, ih, il;
                split2(v[u].x * sc, v[u].z * sc, rh, rl);
                split2(v[u].y * sc, v[u].w * sc, ih, il);
                unsigned* w = pl0 + lane + 36 * u;         // samples 64*u + 2*lane, +1 -> chunk u, 72 fp16 per chunk
                w[0] = rh;
                w[PLW] = rl;
                w[2 * PLW] = ih;
                w[3 * PLW] = il;
            }
        }
        __syncwarp();
        if constexpr (DEMOD) {
            const __half* p16 = reinterpret_cast<const __half*>(s_planes);
            float re = 0.f, im = 0.f;
            for (int j = lane; j < a.ntaps; j += 32) {
                const int s = BT * D + j;                  // = 512 + j
                const int e = s + (s >> 6) * 8;
                const float w = __ldg(a.taps_rev + j);
                re = fmaf(__half2float(p16[e]) + __half2float(p16[e + 2 * PLW]), w, re);
                im = fmaf(__half2float(p16[e + 4 * PLW]) + __half2float(p16[e + 6 * PLW]), w, im);
            }
#pragma unroll
            for (int d = 16; d; d >>= 1) {
                re += __shfl_xor_sync(0xffffffffu, re, d);
                im += __shfl_xor_sync(0xffffffffu, im, d);
            }
            if (lane == 0) s_y[BT] = make_float2(re * isc, im * isc);
        }
        // ---- Toeplitz product over the half-fragment walk ----
        float acc[S][4];
#pragma unroll
        for (int j = 0; j < S; ++j)
#pragma unroll
            for (int q = 0; q < 4; ++q) acc[j][q] = 0.f;
        unsigned ah[2][4], al[2][4];                       // A operands (hi, lo) of walk positions p, p + 1
        uint4 bw[5];                                       // B fragments of the k-steps alive at p (<= 4) and the next one, slot ks % 5
        ldsm4(ah[0], lane_addr);
        ldsm4(al[0], lane_addr + PLW * 4);
        bw[0] = s_b[0];
#pragma unroll
        for (int p = 0; p <= QL; p += QS) {
            const int cur = (p / QS) & 1;
            if (p < QL) {
                const int q = p + QS;
                // second-half lanes (half fragment q + 1) cross a 64-sample chunk when (q + 1) % 8 == 0 (odd positions only)
                const unsigned ad = (((q + 1) & 7) == 0 ? lane_addr7 : lane_addr) + 16u * (unsigned)q + 16u * (unsigned)(q >> 3);
                ldsm4(ah[cur ^ 1], ad);
                ldsm4(al[cur ^ 1], ad + PLW * 4);
            }
            if ((p & 1) == 0 && p / 2 + 1 < KS) bw[(p / 2 + 1) % 5] = s_b[(p / 2 + 1) * 32];   // first used at p + 2
            // term by term over the position's (m-tile, k-step) pairs: consecutive mma write different accumulators
#pragma unroll
            for (int ks = 0; ks < KS; ++ks) {
                const int dj = p - 2 * ks;
                if (dj >= 0 && dj % D == 0 && dj / D < S) mma_f16(acc[dj / D], al[cur], bw[ks % 5].x, bw[ks % 5].y);
            }
#pragma unroll
            for (int ks = 0; ks < KS; ++ks) {
                const int dj = p - 2 * ks;
                if (dj >= 0 && dj % D == 0 && dj / D < S) mma_f16(acc[dj / D], ah[cur], bw[ks % 5].z, bw[ks % 5].w);
            }
#pragma unroll
            for (int ks = 0; ks < KS; ++ks) {
                const int dj = p - 2 * ks;
                if (dj >= 0 && dj % D == 0 && dj / D < S) mma_f16(acc[dj / D], ah[cur], bw[ks % 5].x, bw[ks % 5].y);
            }
        }
        // lane (g, t) of m-tile j holds (re, im) of outputs 2t, 2t+1 of block-row j + S*g
        const int g = lane >> 2, t = lane & 3;
        if constexpr (DEMOD) {
#pragma unroll
            for (int j = 0; j < S; ++j)
                *reinterpret_cast<float4*>(s_y + (j + S * g) * 8 + 2 * t) =
                    make_float4(acc[j][0] * inv, acc[j][2] * inv, acc[j][1] * inv, acc[j][3] * inv);
            __syncwarp();
            float* __restrict__ out = reinterpret_cast<float*>(a.out) + ch * a.out_stride + ob;
            const long long left = a.out_n - 1 - ob;
#pragma unroll 4
            for (int o = lane; o < BT; o += 32)
                if (o < left) out[o] = demod_pair(s_y[o], s_y[o + 1], a.gain);
        } else {
            float2* __restrict__ outc = reinterpret_cast<float2*>(a.out) + ch * a.out_stride + ob;
            const long long left_c = a.out_n - ob;
            if ((reinterpret_cast<unsigned long long>(outc) & 15ull) == 0 && left_c >= BT) {
#pragma unroll
                for (int j = 0; j < S; ++j)
                    *reinterpret_cast<float4*>(outc + (j + S * g) * 8 + 2 * t) =
                        make_float4(acc[j][0] * inv, acc[j][2] * inv, acc[j][1] * inv, acc[j][3] * inv);
            } else {
#pragma unroll
                for (int j = 0; j < S; ++j) {
                    const int o = (j + S * g) * 8 + 2 * t;
                    if (o < left_c) outc[o] = make_float2(acc[j][0] * inv, acc[j][2] * inv);
                    if (o + 1 < left_c) outc[o + 1] = make_float2(acc[j][1] * inv, acc[j][3] * inv);
                }
            }
        }
        __syncwarp();
    }
}

// ---- f32 streams (FirFilter<Float>): the same walk with 16 block-rows per m-tile ------------------------------
// Real samples have no re/im pairing, so the 16 rows of an mma A tile are 16 block-rows (b = j + S*r, r < 16) and a
// warp tile gives 128*S outputs from 1024 + 16*KS staged samples; two planes (hi, lo) with the same chunk layout
// (8 fp16 after every 64 samples, 144-byte row pitch).  The ldmatrix matrices of walk position q are
// {rows 0-7 @q, rows 8-15 @q, rows 0-7 @q+1, rows 8-15 @q+1}; a lane's accumulators are outputs 2t, 2t+1 of
// block-rows j + S*g and j + S*(g+8).  B fragments are the c32 kernel's (same taps, same Toeplitz matrix).
__host__ __device__ constexpr int fir_tcf_L(int KS) { return 1024 + 16 * KS; }
__host__ __device__ constexpr int fir_tcf_nld(int KS) { return (fir_tcf_L(KS) / 4 + 31) / 32; }
__host__ __device__ constexpr int fir_tcf_plw(int KS) { return (fir_tcf_L(KS) + 63) / 64 * 36; }
__host__ __device__ constexpr size_t fir_tcf_smem(int KS) { return (size_t)KS * 512 + (size_t)(FIR_TC_THREADS / 32) * 2 * fir_tcf_plw(KS) * 4; }

template <int KS, int D>
__global__ void __launch_bounds__(FIR_TC_THREADS, 3) fir_tcf_kernel(const FirTcfArgs a) {
    static_assert(D == 1 || D == 2 || D == 4 || D == 8, "fir_tcf_kernel: deci 1, 2, 4 or 8");
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int NW = FIR_TC_THREADS / 32;
    constexpr int S = 8 / D;
    constexpr int QL = (8 - D) + 2 * (KS - 1);
    constexpr int QS = D == 1 ? 1 : 2;
    constexpr int BT = 128 * S;                            // outputs per warp tile (1024 input samples + halo)
    constexpr int L = fir_tcf_L(KS);
    constexpr int NQ = L / 4;                              // float4 units
    constexpr int NLD = fir_tcf_nld(KS);
    constexpr int PLW = fir_tcf_plw(KS);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint4* s_b = reinterpret_cast<const uint4*>(smem_raw) + lane;
    unsigned char* s_planes = smem_raw + (size_t)KS * 512 + (size_t)warp * (2 * PLW * 4);
    unsigned* pl0 = reinterpret_cast<unsigned*>(s_planes);
    for (int i = threadIdx.x; i < KS * 32; i += FIR_TC_THREADS) reinterpret_cast<uint4*>(smem_raw)[i] = __ldg(a.bfrag + i);
    __syncthreads();                                       // the only CTA barrier

    const int mat = lane >> 3;
    const unsigned lane_addr = (unsigned)__cvta_generic_to_shared(s_planes) + (unsigned)((lane & 7) + 8 * (mat & 1)) * 144u +
                               (unsigned)(mat >> 1) * 16u;
    const unsigned lane_addr7 = lane_addr + (unsigned)(mat >> 1) * 16u;
    // the lane's four samples 4*lane + 128*u sit in chunk 2*u + (lane >> 4): word 2*lane + 4*(lane >> 4) + 72*u
    unsigned* lane_w = pl0 + 2 * lane + 4 * (lane >> 4);

    const long long nworkers = (long long)gridDim.x * NW;
    for (long long id = (long long)blockIdx.x * NW + warp; id < a.total_tiles; id += nworkers) {
        const long long ch = (long long)((unsigned)id / (unsigned)a.tiles_x);      // total_tiles < 2^31 (checked on the host): 32-bit division
        const long long ob = (id - ch * a.tiles_x) * BT;
        float4 v[NLD];
        {
            const float* __restrict__ in = a.in + ch * a.in_stride + ob * D;
            const long long avail = a.need - ob * D;
            if (avail >= L && (reinterpret_cast<unsigned long long>(in) & 15ull) == 0) {
#pragma unroll
                for (int u = 0; u < NLD; ++u) {
                    const int e = lane + u * 32;
                    v[u] = e < NQ ? __ldg(reinterpret_cast<const float4*>(in) + e) : make_float4(0.f, 0.f, 0.f, 0.f);
                }
            } else {
#pragma unroll
                for (int u = 0; u < NLD; ++u) {
                    const long long s = 4ll * (lane + u * 32);
                    v[u].x = (s < L && s < avail) ? __ldg(in + s) : 0.f;
                    v[u].y = (s + 1 < L && s + 1 < avail) ? __ldg(in + s + 1) : 0.f;
                    v[u].z = (s + 2 < L && s + 2 < avail) ? __ldg(in + s + 2) : 0.f;
                    v[u].w = (s + 3 < L && s + 3 < avail) ? __ldg(in + s + 3) : 0.f;
                }
            }
        }
        const unsigned ex = tc_tile_exp<NLD>(v);
        const bool scaled = ex >= 14u && ex < 255u;
        const float sc = scaled ? __uint_as_float((267u - ex) << 23) : 1.0f;
        const float inv = (scaled ? __uint_as_float((ex - 13u) << 23) : 1.0f) * a.tap_inv_scale;
#pragma unroll
        for (int u = 0; u < NLD; ++u) {
            if (lane + u * 32 < NQ) {
                unsigned h0, l0, h1, l1;
                split2(v[u].x * sc, v[u].y * sc, h0, l0);
                split2(v[u].z * sc, v[u].w * sc, h1, l1);
                unsigned* w = lane_w + 72 * u;
                *reinterpret_cast<uint2*>(w) = make_uint2(h0, h1);
                *reinterpret_cast<uint2*>(w + PLW) = make_uint2(l0, l1);
            }
        }
        __syncwarp();
        float acc[S][4];
#pragma unroll
        for (int j = 0; j < S; ++j)
#pragma unroll
            for (int q = 0; q < 4; ++q) acc[j][q] = 0.f;
        unsigned ah[2][4], al[2][4];
        uint4 bw[5];
        ldsm4(ah[0], lane_addr);
        ldsm4(al[0], lane_addr + PLW * 4);
        bw[0] = s_b[0];
#pragma unroll
        for (int p = 0; p <= QL; p += QS) {
            const int cur = (p / QS) & 1;
            if (p < QL) {
                const int q = p + QS;
                const unsigned ad = (((q + 1) & 7) == 0 ? lane_addr7 : lane_addr) + 16u * (unsigned)q + 16u * (unsigned)(q >> 3);
                ldsm4(ah[cur ^ 1], ad);
                ldsm4(al[cur ^ 1], ad + PLW * 4);
            }
            if ((p & 1) == 0 && p / 2 + 1 < KS) bw[(p / 2 + 1) % 5] = s_b[(p / 2 + 1) * 32];
#pragma unroll
            for (int ks = 0; ks < KS; ++ks) {
                const int dj = p - 2 * ks;
                if (dj >= 0 && dj % D == 0 && dj / D < S) mma_f16(acc[dj / D], al[cur], bw[ks % 5].x, bw[ks % 5].y);
            }
#pragma unroll
            for (int ks = 0; ks < KS; ++ks) {
                const int dj = p - 2 * ks;
                if (dj >= 0 && dj % D == 0 && dj / D < S) mma_f16(acc[dj / D], ah[cur], bw[ks % 5].z, bw[ks % 5].w);
            }
#pragma unroll
            for (int ks = 0; ks < KS; ++ks) {
                const int dj = p - 2 * ks;
                if (dj >= 0 && dj % D == 0 && dj / D < S) mma_f16(acc[dj / D], ah[cur], bw[ks % 5].x, bw[ks % 5].y);
            }
        }
        // lane (g, t) of m-tile j: outputs 2t, 2t+1 of block-rows j + S*g (c0, c1) and j + S*(g + 8) (c2, c3)
        const int g = lane >> 2, t = lane & 3;
        float* __restrict__ outp = a.out + ch * a.out_stride + ob;
        const long long left = a.out_n - ob;
        if ((reinterpret_cast<unsigned long long>(outp) & 7ull) == 0 && left >= BT) {
#pragma unroll
            for (int j = 0; j < S; ++j) {
                const int o = (j + S * g) * 8 + 2 * t;
                *reinterpret_cast<float2*>(outp + o) = make_float2(acc[j][0] * inv, acc[j][1] * inv);
                *reinterpret_cast<float2*>(outp + o + 64 * S) = make_float2(acc[j][2] * inv, acc[j][3] * inv);
            }
        } else {
#pragma unroll
            for (int j = 0; j < S; ++j) {
                const int o = (j + S * g) * 8 + 2 * t;
                if (o < left) outp[o] = acc[j][0] * inv;
                if (o + 1 < left) outp[o + 1] = acc[j][1] * inv;
                if (o + 64 * S < left) outp[o + 64 * S] = acc[j][2] * inv;
                if (o + 64 * S + 1 < left) outp[o + 64 * S + 1] = acc[j][3] * inv;
            }
        }
        __syncwarp();
    }
}

}  // namespace rrc
