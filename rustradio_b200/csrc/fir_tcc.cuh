// fir_tcc.cuh — complex-tap form of the tensor-core walk kernel (fir_tc.cuh): FirFilter<Complex> with complex taps,
// which is what FirFilter::builder().translate() produces (src/fir.rs:416-474: taps pre-rotated, per-output rotator).
//     y = sum z*w,  y_re = sum zr*wr - zi*wi,  y_im = sum zr*wi + zi*wr
// The A tile already holds the re samples in rows 0-7 and the im samples in rows 8-15, so two Toeplitz products with
// the SAME A operands — P = z * Re(w), Q = z * Im(w) — give all four real sums in one lane's accumulators:
// y_re = P[re row] - Q[im row], y_im = Q[re row] + P[im row].  Twice the mma of the real-tap kernel for a filter that
// costs the FP32 path twice the FMAs; same block-scaled fp16x3 operands, same walk, same layout.  The translate
// rotator is the FP32 kernels' exact-phase one (fir_common.cuh).  128 registers: two CTAs per SM.
#pragma once
#include "fir_tc.cuh"

namespace rrc {

template <int KS, bool DEMOD, int D>
__global__ void __launch_bounds__(FIR_TC_THREADS, 2) fir_tcc_kernel(const FirTccArgs a) {
    static_assert(D == 1 || D == 2 || D == 4 || D == 8, "fir_tcc_kernel: deci 1, 2, 4 or 8");
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
    const uint4* s_b = reinterpret_cast<const uint4*>(smem_raw) + lane;    // B fragments: [KS][2 = re, im taps][32] uint4, this lane's column
    unsigned char* s_planes = smem_raw + (size_t)KS * 1024 + (size_t)warp * fir_tc1_wb(KS, DEMOD);
    float2* s_y = reinterpret_cast<float2*>(s_planes + 4 * PLW * 4);
    unsigned* pl0 = reinterpret_cast<unsigned*>(s_planes);
    for (int i = threadIdx.x; i < KS * 64; i += FIR_TC_THREADS) reinterpret_cast<uint4*>(smem_raw)[i] = __ldg(a.bfrag + i);
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
            if (a.in_u8) {                                 // RtlSdrDecode fused into the tile load (bit-identical to decoding first)
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
            if (lane == 0 && nid < a.total_tiles && !a.in_u8) {
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
                const float2 w = __ldg(a.taps_rev_c + j);
                const float xr = __half2float(p16[e]) + __half2float(p16[e + 2 * PLW]);
                const float xi = __half2float(p16[e + 4 * PLW]) + __half2float(p16[e + 6 * PLW]);
                re = fmaf(xr, w.x, fmaf(-xi, w.y, re));
                im = fmaf(xr, w.y, fmaf(xi, w.x, im));
            }
#pragma unroll
            for (int d = 16; d; d >>= 1) {
                re += __shfl_xor_sync(0xffffffffu, re, d);
                im += __shfl_xor_sync(0xffffffffu, im, d);
            }
            if (lane == 0) {
                float2 yb = make_float2(re * isc, im * isc);
                if (a.translate) yb = cmulf(yb, rotator(a.ratio, (unsigned long long)(a.ntaps - 1) + (a.out_base + (unsigned long long)(ob + BT)) * D));
                s_y[BT] = yb;
            }
        }
        // ---- Toeplitz product over the half-fragment walk ----
        float acc[S][4], acd[S][4];                        // products with the taps' real parts / imaginary parts
#pragma unroll
        for (int j = 0; j < S; ++j)
#pragma unroll
            for (int q = 0; q < 4; ++q) { acc[j][q] = 0.f; acd[j][q] = 0.f; }
        unsigned ah[2][4], al[2][4];                       // A operands (hi, lo) of walk positions p, p + 1
        uint4 bw[5], bx[5];                                // B fragments of the real / imaginary tap parts                                       // B fragments of the k-steps alive at p (<= 4) and the next one, slot ks % 5
        ldsm4(ah[0], lane_addr);
        ldsm4(al[0], lane_addr + PLW * 4);
        bw[0] = s_b[0];
        bx[0] = s_b[32];
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
            if ((p & 1) == 0 && p / 2 + 1 < KS) {            // first used at p + 2
                bw[(p / 2 + 1) % 5] = s_b[(p / 2 + 1) * 64];
                bx[(p / 2 + 1) % 5] = s_b[(p / 2 + 1) * 64 + 32];
            }
            // term by term over the position's (m-tile, k-step) pairs: consecutive mma write different accumulators
#pragma unroll
            for (int ks = 0; ks < KS; ++ks) {
                const int dj = p - 2 * ks;
                if (dj >= 0 && dj % D == 0 && dj / D < S) {
                    mma_f16(acc[dj / D], al[cur], bw[ks % 5].x, bw[ks % 5].y);
                    mma_f16(acd[dj / D], al[cur], bx[ks % 5].x, bx[ks % 5].y);
                }
            }
#pragma unroll
            for (int ks = 0; ks < KS; ++ks) {
                const int dj = p - 2 * ks;
                if (dj >= 0 && dj % D == 0 && dj / D < S) {
                    mma_f16(acc[dj / D], ah[cur], bw[ks % 5].z, bw[ks % 5].w);
                    mma_f16(acd[dj / D], ah[cur], bx[ks % 5].z, bx[ks % 5].w);
                }
            }
#pragma unroll
            for (int ks = 0; ks < KS; ++ks) {
                const int dj = p - 2 * ks;
                if (dj >= 0 && dj % D == 0 && dj / D < S) {
                    mma_f16(acc[dj / D], ah[cur], bw[ks % 5].x, bw[ks % 5].y);
                    mma_f16(acd[dj / D], ah[cur], bx[ks % 5].x, bx[ks % 5].y);
                }
            }
        }
        // lane (g, t) of m-tile j: rows g / g + 8 are the re / im samples of block-row j + S*g, so with P = z * Re(w)
        // (acc) and Q = z * Im(w) (acd):  y_re = P[re row] - Q[im row],  y_im = Q[re row] + P[im row]  for outputs 2t, 2t+1
        const int g = lane >> 2, t = lane & 3;
        auto out_pair = [&](int j, int o, float2& y0, float2& y1) {
            y0 = make_float2((acc[j][0] - acd[j][2]) * inv, (acd[j][0] + acc[j][2]) * inv);
            y1 = make_float2((acc[j][1] - acd[j][3]) * inv, (acd[j][1] + acc[j][3]) * inv);
            if (a.translate) {                              // FirFilter translate epilogue (src/fir.rs:453-473), exact-phase rotator
                const unsigned long long k0 = (unsigned long long)(a.ntaps - 1) + (a.out_base + (unsigned long long)(ob + o)) * D;
                y0 = cmulf(y0, rotator(a.ratio, k0));
                y1 = cmulf(y1, rotator(a.ratio, k0 + D));
            }
        };
        if constexpr (DEMOD) {
#pragma unroll
            for (int j = 0; j < S; ++j) {
                const int o = (j + S * g) * 8 + 2 * t;
                float2 y0, y1;
                out_pair(j, o, y0, y1);
                *reinterpret_cast<float4*>(s_y + o) = make_float4(y0.x, y0.y, y1.x, y1.y);
            }
            __syncwarp();
            float* __restrict__ out = reinterpret_cast<float*>(a.out) + ch * a.out_stride + ob;
            const long long left = a.out_n - 1 - ob;
#pragma unroll 4
            for (int o = lane; o < BT; o += 32)
                if (o < left) out[o] = demod_pair(s_y[o], s_y[o + 1], a.gain);
        } else {
            float2* __restrict__ outc = reinterpret_cast<float2*>(a.out) + ch * a.out_stride + ob;
            const long long left_c = a.out_n - ob;
            const bool fast = (reinterpret_cast<unsigned long long>(outc) & 15ull) == 0 && left_c >= BT;
#pragma unroll
            for (int j = 0; j < S; ++j) {
                const int o = (j + S * g) * 8 + 2 * t;
                float2 y0, y1;
                out_pair(j, o, y0, y1);
                if (fast) {
                    *reinterpret_cast<float4*>(outc + o) = make_float4(y0.x, y0.y, y1.x, y1.y);
                } else {
                    if (o < left_c) outc[o] = y0;
                    if (o + 1 < left_c) outc[o + 1] = y1;
                }
            }
        }
        __syncwarp();
    }
}


}  // namespace rrc
