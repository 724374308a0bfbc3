// fir.cu — decimating FIR for Complex<f32> and f32 streams on sm_100a.
//
// Replaces rustradio's Fir<T>::filter / filter_n_inplace (src/fir.rs:166-197)
// and the compute step of FirFilter<T>::work (src/fir.rs:526-531).
//
//   out[i] = sum_{j<T} in[i*D + j] * h'[j],   h'[j] = taps[T-1-j]
//
// Kernel design (polyphase register-blocked FIR):
//   j = q*D + p  =>  out[i] = sum_p sum_q in[(i+q)*D + p] * h'[q*D + p].
//   For one phase p this is a NON-decimated FIR over the sub-sequence
//   x_p[n] = in[n*D + p], so a thread that owns R consecutive outputs can
//   slide a register window over x_p and reuse every tap for R outputs and
//   every input for R taps (R*R MACs per R window loads + R tap loads).
//   A CTA stages the input span of its NT*R outputs in shared memory once
//   (coalesced global reads, HBM traffic = algorithmic bytes), padded by one
//   element per thread segment (S = R*D elements, R even => S+1 odd) so the
//   thread-strided window reads are bank-conflict free.  Taps live in shared
//   memory phase-major and are read as warp-uniform broadcasts.
//   Outputs are transposed through shared memory for coalesced stores; the
//   fused QuadratureDemod epilogue (rtl_fm shape) reads y[i], y[i+1] there.
//
// Roofline: FP32-FMA bound.  MACs per output: 4*T FFMA (complex taps),
// 2*T (complex data, real taps), T (f32).  Bytes: 8*(N_in + N_out) c32.
//
// Filters with ntaps/deci >= 32 (config 1) leave these FP32 kernels for the tensor-core Toeplitz-block products
// in fir_tc.cuh (real taps, c32 and f32 streams) and fir_tcc.cuh (complex taps / translate): block-scaled
// fp16x3, FP32-class accuracy; plan_tc / plan_tc_cplx below pick the kernel, rrc_fir_uses_tensor_cores declares
// it, RRC_FIR_NO_TENSOR keeps the kernels of this file.
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <type_traits>
#include <vector>

#include "common.cuh"
#include "fir_common.cuh"
#include "epilogue.cuh"
#include "fir_tc.hpp"
#include "pipeline.cuh"

namespace rrc {

constexpr int FIR_R = 8;          // outputs per thread (must be even)
constexpr int FIR_MAX_NT = 256;

struct FirArgs {
    const void* in;
    void* out;
    const void* taps;             // poly: phase-major padded; generic: reversed flat
    long long in_stride, out_stride;
    long long need, out_n;
    int ntaps, deci, qpad, nchunks, S, nseg;
    int nbuf;                     // 1 or 2 input-tile buffers (2 = prefetch next tile while computing)
    long long tiles_x, total_tiles;   // tiles per channel, tiles over all channels
    float gain;
    int translate;
    double ratio;                 // freq / samp_rate
    unsigned long long out_base;  // absolute index of out[0] (translate rotator)
    int in_u8;                    // 1: `in` is u8 I/Q pairs, decoded while the tile is staged (c32 filters only)
    Epi epi;                      // store epilogue of the non-demod c32 paths (epilogue.cuh); MAG2 makes `out` an f32 array
};

// R consecutive c32 outputs y[0..R) of one thread at output index gi0 of channel `ch`, through the epilogue.
template <int R>
__device__ __forceinline__ void fir_store_epi(const FirArgs& a, long long ch, long long gi0, const float2 (&y)[R]) {
    if (a.epi.kind == RRC_EPI_MAG2) {
        float* dst = reinterpret_cast<float*>(a.out) + ch * a.out_stride + gi0;
#pragma unroll
        for (int r = 0; r < R; ++r)
            if (gi0 + r < a.out_n) dst[r] = epi_mag2(y[r]);
    } else {
        float2* dst = reinterpret_cast<float2*>(a.out) + ch * a.out_stride + gi0;
#pragma unroll
        for (int r = 0; r < R; ++r)
            if (gi0 + r < a.out_n) dst[r] = epi_c32(y[r], a.epi);
    }
}

__device__ __forceinline__ void mac(float2& acc, float2 h, float2 x) {
    acc.x = fmaf(h.x, x.x, acc.x);
    acc.x = fmaf(-h.y, x.y, acc.x);
    acc.y = fmaf(h.x, x.y, acc.y);
    acc.y = fmaf(h.y, x.x, acc.y);
}
__device__ __forceinline__ void mac(float2& acc, float h, float2 x) {
    acc.x = fmaf(h, x.x, acc.x);
    acc.y = fmaf(h, x.y, acc.y);
}
__device__ __forceinline__ void mac(float& acc, float h, float x) { acc = fmaf(h, x, acc); }
__device__ __forceinline__ void zero(float2& v) { v = make_float2(0.f, 0.f); }
__device__ __forceinline__ void zero(float& v) { v = 0.f; }

// FirFilter translate epilogue: y[i] *= phi_i, phi_i = exp(-j*theta*((T-1) + i*D))
// (src/fir.rs:453-462 closed form of the :464-473 recurrence).
__device__ __forceinline__ float2 apply_translate(const FirArgs& a, float2 y, long long gi) {
    unsigned long long k = (unsigned long long)(a.ntaps - 1) + (a.out_base + (unsigned long long)gi) * (unsigned long long)a.deci;
    return cmulf(y, rotator(a.ratio, k));
}
__device__ __forceinline__ float apply_translate(const FirArgs&, float y, long long) { return y; }

template <int R>
__device__ __forceinline__ void load_taps(const float2* tp, float2 (&h)[R]) {
#pragma unroll
    for (int k = 0; k < R; k += 2) {
        float4 v = *reinterpret_cast<const float4*>(tp + k);
        h[k] = make_float2(v.x, v.y);
        h[k + 1] = make_float2(v.z, v.w);
    }
}
template <int R>
__device__ __forceinline__ void load_taps(const float* tp, float (&h)[R]) {
#pragma unroll
    for (int k = 0; k < R; k += 4) {
        float4 v = *reinterpret_cast<const float4*>(tp + k);
        h[k] = v.x; h[k + 1] = v.y; h[k + 2] = v.z; h[k + 3] = v.w;
    }
}

// ST: sample type (float2 / float); TT: tap type; DCT: compile-time decimation (0 = run time), so
// that every window offset u*deci is an immediate for the common decimations;
// DEMOD: fused conj-multiply + atan2 epilogue (ST must be float2).
// R: outputs per thread = taps per chunk (8; 16 for the real-tap deci-1 case, which halves the
// shared-memory loads per FMA).

// u8 I/Q -> c32 while staging a tile: 16-bit loads, UB in flight per thread, then decode and store.
// Out-of-range elements read as the byte pair (127, 127), which decodes to exactly 0.  Kept out of
// line so that its registers do not count against the c32 kernels that share fir_load_tile.
__device__ __noinline__ void fir_stage_u8(const unsigned short* __restrict__ in8, long long lim, float2* s_tile,
                                          int L, int S, int NT, int t) {
    const int S1 = S + 1;
    int seg = t / S, rem = t - seg * S;
    const int dseg = NT / S, drem = NT - dseg * S;
    constexpr int UB = 8;
    for (int e = t; e < L; e += UB * NT) {
        unsigned int w[UB];
        int idx[UB];
#pragma unroll
        for (int k = 0; k < UB; ++k) {
            const int ee = e + k * NT;
            idx[k] = seg * S1 + rem;
            w[k] = (ee < L && ee < lim) ? in8[ee] : 0x7f7fu;
            seg += dseg; rem += drem;
            if (rem >= S) { rem -= S; ++seg; }
        }
#pragma unroll
        for (int k = 0; k < UB; ++k)
            if (e + k * NT < L) s_tile[idx[k]] = decode_iq(w[k]);
    }
}

// Stage the input span of tile `bx` of channel `ch` into `s_tile` (padded: one pad element after every S).
template <typename ST, int R>
__device__ __forceinline__ void fir_load_tile(const FirArgs& a, ST* s_tile, long long ch, long long bx, int deci, int S,
                                              int NT, int t, int bstride) {
    const int S1 = S + 1;
    const ST* __restrict__ in = reinterpret_cast<const ST*>(a.in) + ch * a.in_stride;
    const int L = a.nseg * S;
    const long long g0 = bx * bstride * deci;
    if constexpr (sizeof(ST) == 8) {
        if (a.in_u8) {
            fir_stage_u8(reinterpret_cast<const unsigned short*>(a.in) + ch * a.in_stride + g0, a.need - g0,
                         reinterpret_cast<float2*>(s_tile), L, S, NT, t);
            asm volatile("cp.async.commit_group;" ::: "memory");
            return;
        }
    }
    if (g0 + L <= a.need) {
        // Interior tile: asynchronous 8/4-byte copies global -> shared (LDGSTS), all in flight
        // at once, so the tile costs one memory latency instead of one per loop iteration.
        const unsigned sbase = (unsigned)__cvta_generic_to_shared(s_tile);
        if (S >= 64) {      // one warp per thread-segment: no div/mod, constant strides
            const int lane = t & 31, nwarp = NT >> 5;
            for (int sg = t >> 5; sg < a.nseg; sg += nwarp) {
                const ST* src = in + g0 + (long long)sg * S;
                const unsigned dst = sbase + (unsigned)(sg * S1) * (unsigned)sizeof(ST);
                for (int e = lane; e < S; e += 32) {
                    if constexpr (sizeof(ST) == 8)
                        asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst + e * 8), "l"(src + e) : "memory");
                    else
                        asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst + e * 4), "l"(src + e) : "memory");
                }
            }
        } else {            // short segments: flat element loop with incremental (segment, offset)
            const ST* gsrc = in + g0;
            int seg = t / S, rem = t - seg * S;
            const int dseg = NT / S, drem = NT - dseg * S;
            for (int e = t; e < L; e += NT) {
                const unsigned dst = sbase + (unsigned)(seg * S1 + rem) * (unsigned)sizeof(ST);
                if constexpr (sizeof(ST) == 8)
                    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst), "l"(gsrc + e) : "memory");
                else
                    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst), "l"(gsrc + e) : "memory");
                seg += dseg; rem += drem;
                if (rem >= S) { rem -= S; ++seg; }
            }
        }
    } else {
        int seg = t / S, rem = t - seg * S;
        const int dseg = NT / S, drem = NT - dseg * S;
        for (int e = t; e < L; e += NT) {
            const long long g = g0 + e;
            ST v; zero(v);
            if (g < a.need) v = in[g];
            s_tile[seg * S1 + rem] = v;
            seg += dseg; rem += drem;
            if (rem >= S) { rem -= S; ++seg; }
        }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
}

template <typename ST, typename TT, int DCT, bool DEMOD, int R>
__global__ void __launch_bounds__(FIR_MAX_NT) fir_poly_kernel(const FirArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int deci = DCT ? DCT : a.deci;
    const int S = DCT ? R * DCT : a.S;
    const int S1 = S + 1;
    const int NT = blockDim.x, t = threadIdx.x;
    const int BT = NT * R;
    const int bstride = DEMOD ? BT - 1 : BT;
    const int ntap_tab = deci * a.qpad;
    TT* s_taps = reinterpret_cast<TT*>(smem_raw);
    ST* s_tile = reinterpret_cast<ST*>(smem_raw + (((size_t)ntap_tab * sizeof(TT) + 15) & ~(size_t)15));

    // One tile per CTA (blockIdx.x = tile in channel, blockIdx.y = channel): a persistent,
    // double-buffered variant of this kernel measured 8-20 % slower (more registers, and the
    // hardware CTA scheduler already overlaps one CTA's tile load with its neighbours' FMAs).
    const long long ch = blockIdx.y;
    const long long ob = (long long)blockIdx.x * bstride;

    fir_load_tile<ST, R>(a, s_tile, ch, blockIdx.x, deci, S, NT, t, bstride);
    {   // taps -> smem (phase-major, zero padded to qpad per phase)
        const TT* __restrict__ gt = reinterpret_cast<const TT*>(a.taps);
        for (int i = t; i < ntap_tab; i += NT) s_taps[i] = gt[i];
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();

    ST acc[R];
#pragma unroll
    for (int r = 0; r < R; ++r) zero(acc[r]);

    const ST* base_t = s_tile + t * S1;
    for (int p = 0; p < deci; ++p) {
        const ST* bp = base_t + p;
        const TT* tp = s_taps + p * a.qpad;
        ST w[2 * R - 1];
#pragma unroll
        for (int u = 0; u < R - 1; ++u) w[u] = bp[u * deci];
        for (int c = 0; c < a.nchunks; ++c) {
#pragma unroll
            for (int u = R - 1; u < 2 * R - 1; ++u) w[u] = bp[u * deci + (u >= R ? 1 : 0)];
            TT h[R];
            load_taps<R>(tp, h);
#pragma unroll
            for (int k = 0; k < R; ++k)
#pragma unroll
                for (int r = 0; r < R; ++r) mac(acc[r], h[k], w[r + k]);
#pragma unroll
            for (int u = 0; u < R - 1; ++u) w[u] = w[u + R];
            bp += S1;
            tp += R;
        }
    }

    if constexpr (!DEMOD) {
        // Each thread owns R consecutive outputs: store them straight from registers (128-bit
        // when the destination is 16-byte aligned); the sectors a warp half-fills are completed
        // by its own next store instruction, so L2 merges them before write-back.
        ST* __restrict__ out = reinterpret_cast<ST*>(a.out) + ch * a.out_stride;
        const long long gi0 = ob + (long long)t * R;
        if (a.translate) {
#pragma unroll
            for (int r = 0; r < R; ++r) acc[r] = apply_translate(a, acc[r], gi0 + r);
        }
        if constexpr (sizeof(ST) == 8) {
            if (a.epi.kind != RRC_EPI_NONE) { fir_store_epi<R>(a, ch, gi0, acc); return; }
        }
        ST* dst = out + gi0;
        if (gi0 + R <= a.out_n && (reinterpret_cast<unsigned long long>(dst) & 15ull) == 0) {
            constexpr int VE = 16 / (int)sizeof(ST);
#pragma unroll
            for (int r = 0; r < R; r += VE) {
                float4 v;
                if constexpr (sizeof(ST) == 8) v = make_float4(acc[r].x, acc[r].y, acc[r + 1].x, acc[r + 1].y);
                else v = make_float4(acc[r], acc[r + 1], acc[r + 2], acc[r + 3]);
                *reinterpret_cast<float4*>(dst + r) = v;
            }
        } else {
#pragma unroll
            for (int r = 0; r < R; ++r)
                if (gi0 + r < a.out_n) dst[r] = acc[r];
        }
    } else {
        __syncthreads();                 // everyone is done with the input tile
        ST* s_out = s_tile;              // reuse: BT + NT entries <= tile size
#pragma unroll
        for (int r = 0; r < R; ++r) {
            ST y = acc[r];
            if (a.translate) y = apply_translate(a, y, ob + (long long)t * R + r);
            s_out[t * (R + 1) + r] = y;
        }
        __syncthreads();
        float* __restrict__ out = reinterpret_cast<float*>(a.out) + ch * a.out_stride;
        float2 ya[R], yb[R];
#pragma unroll
        for (int r = 0; r < R; ++r) {                 // R independent demods per thread: ILP for atan2
            const int o = t + r * NT;
            ya[r] = s_out[o + o / R];
            yb[r] = s_out[(o + 1) + (o + 1) / R];     // o + 1 <= BT - 1 < BT + NT entries
        }
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const int o = t + r * NT;
            const long long gi = ob + o;
            if (o < BT - 1 && gi < a.out_n - 1) out[gi] = demod_pair(ya[r], yb[r], a.gain);
        }
    }
}

// ---- c32 samples, REAL taps: packed FP32 (FFMA2) kernel ----------------------------------------
// low_pass_complex always produces real taps (src/fir.rs:600-603), so acc.(re,im) += h * x.(re,im)
// is ONE Blackwell packed instruction, fma.rn.f32x2 (SASS FFMA2): sample and accumulator are the
// 64-bit register pairs LDS.64 delivers, the tap is stored in shared memory as the pair (h, h).
// FFMA2 retires two FMAs per issue slot (measured: 1.99 warp-instr/clk/SM = 127 lane-FMA/clk/SM,
// tools/microbench/fp32_pipes.cu), so the window loads, tap loads and address arithmetic issue in the
// shadow of the FP32 pipe instead of competing with it; results are bit-identical to FFMA.
//   SPLIT = 2: two threads share one group of R outputs and take alternate polyphase branches
//   (lanes l and l^16), then add their partial sums with one shuffle per output.  The staged input
//   tile is R*deci samples per GROUP, so this doubles the resident warps for large decimations
//   (config 3: 10 -> 20 warps/SM) without more shared memory.
//   Taps per polyphase branch are NOT rounded up to a multiple of R: full chunks of R taps, then one
//   tail chunk of (Q_p mod R) taps (config 3: 26/25 taps per branch instead of 32 -> -20 % FMAs).
typedef unsigned long long u64;
__device__ __forceinline__ void fma2(u64& acc, u64 hh, u64 x) {
    asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc) : "l"(hh), "l"(x));
}
__device__ __forceinline__ u64 add2(u64 a, u64 b) {
    u64 r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ float2 unpk(u64 v) {
    float2 r;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(r.x), "=f"(r.y) : "l"(v));
    return r;
}

// One chunk of K <= R taps: loads the K new window elements and the K tap pairs, K*R FFMA2,
// then slides the window by K.  bp points at the window start of this chunk (pitch S1 per R*deci).
template <int R, int K>
__device__ __forceinline__ void rt_chunk(u64 (&acc)[R], u64 (&w)[2 * R - 1], const u64* bp, const u64* tp, int deci) {
#pragma unroll
    for (int u = R - 1; u < R - 1 + K; ++u) w[u] = bp[u * deci + (u >= R ? 1 : 0)];
    u64 h[K];
    if constexpr (K % 2 == 0) {
#pragma unroll
        for (int k = 0; k < K; k += 2) {
            const ulonglong2 v = *reinterpret_cast<const ulonglong2*>(tp + k);
            h[k] = v.x; h[k + 1] = v.y;
        }
    } else {
#pragma unroll
        for (int k = 0; k < K; ++k) h[k] = tp[k];
    }
#pragma unroll
    for (int k = 0; k < K; ++k)
#pragma unroll
        for (int r = 0; r < R; ++r) fma2(acc[r], h[k], w[r + k]);
#pragma unroll
    for (int u = 0; u < R - 1; ++u) w[u] = w[u + K];
}

// tail chunk of rem in [1, R) taps (rem is CTA-uniform: a short chain of uniform branches)
template <int R, int K>
struct RtTail {
    static __device__ __forceinline__ void run(int rem, u64 (&acc)[R], u64 (&w)[2 * R - 1], const u64* bp, const u64* tp, int deci) {
        if (rem == K) rt_chunk<R, K>(acc, w, bp, tp, deci);
        else RtTail<R, K - 1>::run(rem, acc, w, bp, tp, deci);
    }
};
template <int R>
struct RtTail<R, 0> {
    static __device__ __forceinline__ void run(int, u64 (&)[R], u64 (&)[2 * R - 1], const u64*, const u64*, int) {}
};

template <int DCT, bool DEMOD, int R, int SPLIT>
__global__ void __launch_bounds__(FIR_MAX_NT) fir_rt_kernel(const FirArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int deci = DCT ? DCT : a.deci;
    const int S = DCT ? R * DCT : a.S;
    const int S1 = S + 1;
    const int NT = blockDim.x, tid = threadIdx.x;
    const int G = NT / SPLIT;                                   // output groups (R outputs each) per CTA
    // half-warps keep 16 consecutive groups (conflict-free strided window reads); the partner that
    // shares a group is lane ^ 16.
    const int s = SPLIT == 1 ? 0 : (tid >> 4) & 1;
    const int t = SPLIT == 1 ? tid : ((tid >> 5) << 4) | (tid & 15);
    const int BT = G * R;
    const int bstride = DEMOD ? BT - 1 : BT;
    const int ntap_tab = deci * a.qpad;
    u64* s_taps = reinterpret_cast<u64*>(smem_raw);
    float2* s_tile = reinterpret_cast<float2*>(smem_raw + (((size_t)ntap_tab * sizeof(u64) + 15) & ~(size_t)15));

    const long long ch = blockIdx.y;
    const long long ob = (long long)blockIdx.x * bstride;
    fir_load_tile<float2, R>(a, s_tile, ch, blockIdx.x, deci, S, NT, tid, bstride);
    {   // taps -> smem as (h, h) pairs, phase-major, zero padded to qpad per phase
        const float* __restrict__ gt = reinterpret_cast<const float*>(a.taps);
        for (int i = tid; i < ntap_tab; i += NT) {
            const float hv = gt[i];
            float2 pr = make_float2(hv, hv);
            s_taps[i] = *reinterpret_cast<u64*>(&pr);
        }
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();

    u64 acc[R];
#pragma unroll
    for (int r = 0; r < R; ++r) acc[r] = 0ull;

    const u64* base_t = reinterpret_cast<const u64*>(s_tile) + t * S1;
    for (int p = s; p < deci; p += SPLIT) {
        const int Qp = a.ntaps > p ? (a.ntaps - p + deci - 1) / deci : 0;   // taps of polyphase branch p
        const u64* bp = base_t + p;
        const u64* tp = s_taps + p * a.qpad;
        u64 w[2 * R - 1];
#pragma unroll
        for (int u = 0; u < R - 1; ++u) w[u] = bp[u * deci];
        const int nfull = Qp / R, rem = Qp - nfull * R;
        for (int c = 0; c < nfull; ++c) {
            rt_chunk<R, R>(acc, w, bp, tp, deci);
            bp += S1;
            tp += R;
        }
        RtTail<R, R - 1>::run(rem, acc, w, bp, tp, deci);
    }
    if constexpr (SPLIT == 2) {
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const u64 o = __shfl_xor_sync(0xffffffffu, acc[r], 16);
            acc[r] = add2(acc[r], o);
        }
    }
    // each thread finishes RS of the group's R outputs: r in [r0, r0 + RS)
    constexpr int RS = R / SPLIT;
    const int r0 = s * RS;

    if constexpr (!DEMOD) {
        float2* __restrict__ out = reinterpret_cast<float2*>(a.out) + ch * a.out_stride;
        const long long gi0 = ob + (long long)t * R + r0;
        float2 y[RS];
#pragma unroll
        for (int r = 0; r < RS; ++r) {
            y[r] = unpk(SPLIT == 2 && s ? acc[(RS + r) % R] : acc[r]);
            if (a.translate) y[r] = apply_translate(a, y[r], gi0 + r);
        }
        if (a.epi.kind != RRC_EPI_NONE) { fir_store_epi<RS>(a, ch, gi0, y); return; }
        float2* dst = out + gi0;
        if (gi0 + RS <= a.out_n && (reinterpret_cast<unsigned long long>(dst) & 15ull) == 0) {
#pragma unroll
            for (int r = 0; r < RS; r += 2)
                *reinterpret_cast<float4*>(dst + r) = make_float4(y[r].x, y[r].y, y[r + 1].x, y[r + 1].y);
        } else {
#pragma unroll
            for (int r = 0; r < RS; ++r)
                if (gi0 + r < a.out_n) dst[r] = y[r];
        }
    } else {
        __syncthreads();                 // everyone is done with the input tile
        float2* s_out = s_tile;          // reuse: BT + G entries <= tile size
#pragma unroll
        for (int r = 0; r < RS; ++r) {
            float2 y = unpk(SPLIT == 2 && s ? acc[(RS + r) % R] : acc[r]);
            if (a.translate) y = apply_translate(a, y, ob + (long long)t * R + r0 + r);
            s_out[t * (R + 1) + r0 + r] = y;
        }
        __syncthreads();
        float* __restrict__ out = reinterpret_cast<float*>(a.out) + ch * a.out_stride;
        float2 ya[RS], yb[RS];
#pragma unroll
        for (int r = 0; r < RS; ++r) {                // RS independent demods per thread: ILP for atan2
            const int o = tid + r * NT;
            ya[r] = s_out[o + o / R];
            yb[r] = s_out[(o + 1) + (o + 1) / R];     // o + 1 <= BT - 1 < BT + G entries
        }
#pragma unroll
        for (int r = 0; r < RS; ++r) {
            const int o = tid + r * NT;
            const long long gi = ob + o;
            if (o < BT - 1 && gi < a.out_n - 1) out[gi] = demod_pair(ya[r], yb[r], a.gain);
        }
    }
}

// ---- c32 samples, REAL taps, taps in the UNIFORM datapath (config 3: the rtl_fm channelizer) ----
// fir_rt_kernel spends 65 % of its issue slots on things that are not FFMA2: tap loads from shared memory
// (one LDS.128 per two taps), window-shift moves (the chunk loop is a run-time loop), tail chunks and index
// arithmetic.  Here the filter is a compile-time SHAPE (DCT polyphase branches x QB taps per branch, zero
// padded) and the taps are a KERNEL PARAMETER: (h, h) pairs in the constant bank, so every FFMA2 takes its
// tap from a uniform register (SASS: FFMA2 R, R.F32x2, UR.F32x2, R — one LDCU.128 on the uniform pipe per
// two taps, no vector register, no shared-memory traffic), the two loops are fully unrolled (the sliding
// window is pure register renaming, every shared-memory offset an immediate) and a thread owns all R = 8
// outputs of its group: per tap ONE LDS.64 and eight FFMA2.  2080 FFMA2 per thread against ~330 window
// loads and ~130 uniform loads.  Results are bit-identical to fir_rt_kernel / FFMA (same products, same
// order: branch by branch, tap by tap).
// Tile loader for a compile-time segment length S (c32, cp.async): a warp copies whole segments, every offset
// an immediate, the segment loop a pointer increment — ~1/4 of the instructions of fir_load_tile's generic loop.
template <int S>
__device__ __forceinline__ void fir_load_tile_ct(const FirArgs& a, float2* s_tile, long long ch, long long bx, int deci,
                                                 int NT, int t, int bstride) {
    constexpr int S1 = S + 1;
    const long long g0 = bx * bstride * deci;
    const int L = a.nseg * S;
    if (a.in_u8 || g0 + L > a.need) {           // u8 input / last tile of a channel: the generic loader
        fir_load_tile<float2, FIR_R>(a, s_tile, ch, bx, deci, S, NT, t, bstride);
        return;
    }
    const float2* __restrict__ in = reinterpret_cast<const float2*>(a.in) + ch * a.in_stride + g0;
    const int lane = t & 31, nwarp = NT >> 5;
    const float2* src = in + (long long)(t >> 5) * S + lane;
    unsigned dst = (unsigned)__cvta_generic_to_shared(s_tile) + (unsigned)(((t >> 5) * S1 + lane) * 8);
    for (int sg = t >> 5; sg < a.nseg; sg += nwarp) {
#pragma unroll
        for (int e = 0; e < S; e += 32)
            if (e + 32 <= S || lane < S - e)
                asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst + e * 8), "l"(src + e) : "memory");
        src += (long long)nwarp * S;
        dst += (unsigned)(nwarp * S1 * 8);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
}

template <int NTAP>
struct RtuTaps { float2 hh[NTAP]; };      // hh[p * QB + q] = (h, h), h = taps_rev[q * DCT + p] (0 beyond ntaps)

template <int DCT, int QB, bool DEMOD, int R>
__global__ void __launch_bounds__(128) fir_rtu_kernel(const FirArgs a, const __grid_constant__ RtuTaps<DCT * QB> taps) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int S = R * DCT, S1 = S + 1;
    const int NT = blockDim.x, tid = threadIdx.x;
    const int BT = NT * R;
    const int bstride = DEMOD ? BT - 1 : BT;
    float2* s_tile = reinterpret_cast<float2*>(smem_raw);
    const long long ch = blockIdx.y;
    const long long ob = (long long)blockIdx.x * bstride;
    fir_load_tile_ct<S>(a, s_tile, ch, blockIdx.x, DCT, NT, tid, bstride);
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();

    u64 acc[R];
#pragma unroll
    for (int r = 0; r < R; ++r) acc[r] = 0ull;
    const u64* base = reinterpret_cast<const u64*>(s_tile) + tid * S1;
    // window element U of branch p (sample (r + q) * DCT + p of the thread's span, U = r + q) sits at
    // p + (U / R) * S1 + (U % R) * DCT: one pad word after every R * DCT samples.
    // The branch loop is NOT unrolled: one branch (QB taps x R FFMA2 + the window loads, ~4 KB of code) stays in
    // the instruction cache; fully unrolled over the branches the kernel is 46 KB of straight-line code and
    // stalls on instruction fetch (ncu: no_instruction 1.2 per issue).  p is warp uniform, so the tap of
    // (p, q) is still a uniform-register operand (constant bank, uniform index).
#pragma unroll 1
    for (int p = 0; p < DCT; ++p) {
        u64 w[R - 1 + QB];
        const u64* bp = base + p;
        const float2* tp = taps.hh + p * QB;
#pragma unroll
        for (int u = 0; u < R - 1; ++u) w[u] = bp[(u / R) * S1 + (u % R) * DCT];
#pragma unroll
        for (int q = 0; q < QB; ++q) {
            const int U = q + R - 1;
            w[U] = bp[(U / R) * S1 + (U % R) * DCT];
            const u64 hh = *reinterpret_cast<const u64*>(&tp[q]);
#pragma unroll
            for (int r = 0; r < R; ++r) fma2(acc[r], hh, w[r + q]);
        }
    }

    if constexpr (!DEMOD) {
        float2* __restrict__ out = reinterpret_cast<float2*>(a.out) + ch * a.out_stride;
        const long long gi0 = ob + (long long)tid * R;
        float2 y[R];
#pragma unroll
        for (int r = 0; r < R; ++r) {
            y[r] = unpk(acc[r]);
            if (a.translate) y[r] = apply_translate(a, y[r], gi0 + r);
        }
        if (a.epi.kind != RRC_EPI_NONE) { fir_store_epi<R>(a, ch, gi0, y); return; }
        float2* dst = out + gi0;
        if (gi0 + R <= a.out_n && (reinterpret_cast<unsigned long long>(dst) & 15ull) == 0) {
#pragma unroll
            for (int r = 0; r < R; r += 2)
                *reinterpret_cast<float4*>(dst + r) = make_float4(y[r].x, y[r].y, y[r + 1].x, y[r + 1].y);
        } else {
#pragma unroll
            for (int r = 0; r < R; ++r)
                if (gi0 + r < a.out_n) dst[r] = y[r];
        }
    } else {
        __syncthreads();                 // everyone is done with the input tile
        float2* s_out = s_tile;          // reuse: BT + NT entries <= tile size
#pragma unroll
        for (int r = 0; r < R; ++r) {
            float2 y = unpk(acc[r]);
            if (a.translate) y = apply_translate(a, y, ob + (long long)tid * R + r);
            s_out[tid * (R + 1) + r] = y;
        }
        __syncthreads();
        float* __restrict__ out = reinterpret_cast<float*>(a.out) + ch * a.out_stride;
        float2 ya[R], yb[R];
#pragma unroll
        for (int r = 0; r < R; ++r) {                 // R independent demods per thread: ILP for atan2
            const int o = tid + r * NT;
            ya[r] = s_out[o + o / R];
            yb[r] = s_out[(o + 1) + (o + 1) / R];     // o + 1 <= BT - 1 < BT + NT entries
        }
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const int o = tid + r * NT;
            const long long gi = ob + o;
            if (o < BT - 1 && gi < a.out_n - 1) out[gi] = demod_pair(ya[r], yb[r], a.gain);
        }
    }
}

// Fallback for geometries whose tile does not fit shared memory (very large
// deci*R or tap tables): one thread per output, taps and inputs through L1/L2.
template <typename ST, typename TT>
__global__ void fir_generic_kernel(const FirArgs a) {
    const ST* __restrict__ in = reinterpret_cast<const ST*>(a.in) + (long long)blockIdx.y * a.in_stride;
    ST* __restrict__ out = reinterpret_cast<ST*>(a.out) + (long long)blockIdx.y * a.out_stride;
    const TT* __restrict__ taps = reinterpret_cast<const TT*>(a.taps);
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < a.out_n; i += stride) {
        const ST* x = in + i * a.deci;
        ST acc; zero(acc);
        if constexpr (sizeof(ST) == 8) {
            if (a.in_u8) {
                const unsigned short* x8 = reinterpret_cast<const unsigned short*>(a.in) + (long long)blockIdx.y * a.in_stride + i * a.deci;
                for (int j = 0; j < a.ntaps; ++j) mac(acc, taps[j], decode_iq(x8[j]));
            } else {
                for (int j = 0; j < a.ntaps; ++j) mac(acc, taps[j], x[j]);
            }
        } else {
            for (int j = 0; j < a.ntaps; ++j) mac(acc, taps[j], x[j]);
        }
        if (a.translate) acc = apply_translate(a, acc, i);
        if constexpr (sizeof(ST) == 8) {
            if (a.epi.kind == RRC_EPI_MAG2) { (reinterpret_cast<float*>(a.out) + (long long)blockIdx.y * a.out_stride)[i] = epi_mag2(acc); continue; }
            acc = epi_c32(acc, a.epi);
        }
        out[i] = acc;
    }
}

__global__ void quad_demod_kernel(const float2* __restrict__ in, long long in_stride, long long n_in,
                                  float gain, float* __restrict__ out, long long out_stride) {
    in += (long long)blockIdx.y * in_stride;
    out += (long long)blockIdx.y * out_stride;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t + 1 < n_in; t += stride)
        out[t] = demod_pair(in[t], in[t + 1], gain);
}

}  // namespace rrc

using namespace rrc;

struct rrc_fir {
    int device = 0;
    bool cplx = false;        // samples are Complex<f32>
    bool real_taps = false;   // tap table holds floats (f32 FIR or c32 real-tap fast path)
    unsigned flags = 0;
    size_t ntaps = 0, deci = 1;
    std::vector<float> taps_host;   // caller order, interleaved if cplx
    void* taps_poly = nullptr;
    void* taps_rev = nullptr;
    int qpad = 0, nchunks = 0, nt = 0, R = FIR_R, nbuf = 1;
    int groups = 0, split = 1;   // nt = groups * split threads; split = 2: two threads per group of R outputs (fir_rt_kernel)
    Epi epi;                     // fused store epilogue (rrc_fir_set_epilogue); set => FP32 kernels only
    bool rt = false;             // c32 samples + real taps: packed-FP32 kernel (fir_rt_kernel)
    // uniform-tap unrolled kernel (fir_rtu_kernel): shape (deci, QB taps per branch), CTA size, smem, (h, h) table
    int rtu_qb = 0, rtu_nt = 0, rtu_r = 8;   // rtu_r: outputs per thread (8 or 16)
    size_t rtu_smem = 0;
    std::vector<float2> rtu_hh;
    size_t smem = 0;
    bool use_poly = false;
    bool translate = false;
    double ratio = 0.0;
    unsigned long long out_counter = 0;
    int in_u8 = 0;               // inputs are u8 I/Q pairs (rrc_fir_set_input_u8iq; c32 filters only)
    // tensor-core Toeplitz kernel (fir_tc.cuh): c32 samples (c32 or u8 I/Q input), real taps, no translate
    bool tc = false;
    int tc_ntile = 1, tc_nld = 9, tc_nm = 1, tc_wb = 0, tc_KS = 0, tc_RS = 0, tc_PAD = 0, tc_L = 0, tc_PL = 0;
    unsigned tc_magic = 0;
    bool tc_cplx = false;        // complex taps (translate filters): fir_tcc_kernel
    bool tc1 = false;            // deci 1/2/4, 7*deci + ntaps <= 320: fir_tc1_kernel (A fragments loaded once per warp tile)
    float tc_tap_inv_scale = 1.0f;
    int tc_ctas_per_sm = 0;      // occupancy of the chosen instantiation (persistent grid), filled at first launch
    void* tc_bfrag = nullptr;
    size_t tc_smem = 0;
    std::vector<unsigned> tc5_tab;   // tcgen05 kernel (fir_tc5.cu): tap words of its TMEM operand; non-empty when the filter qualifies
    int tc5_KS = 0;
    long long tc5_min_tiles = 0;     // launches with fewer 8192-output tiles stay on fir_tc1_kernel (pipeline fill, see plan_tc)
    int last_tc5 = -1;               // what the last launch used: rrc_fir_kernel_name reports it
    Pipe pipe;
};

namespace {

size_t tap_elem(const rrc_fir* h) { return h->real_taps ? sizeof(float) : sizeof(float2); }
size_t samp_elem(const rrc_fir* h) { return h->cplx ? sizeof(float2) : sizeof(float); }

unsigned short f16_rn(float f) {           // round-to-nearest-even f32 -> fp16 (|f| < 65504)
    unsigned x;
    memcpy(&x, &f, 4);
    const unsigned short sign = (unsigned short)((x >> 16) & 0x8000u);
    x &= 0x7fffffffu;
    if (x < 0x38800000u)                   // below 2^-14: subnormal, a multiple of 2^-24
        return (unsigned short)(sign | (unsigned short)std::nearbyint(std::fabs((double)f) * 16777216.0));
    unsigned h = (((x >> 23) - 112u) << 10) | ((x & 0x7fffffu) >> 13);
    const unsigned rem = x & 0x1fffu;
    if (rem > 0x1000u || (rem == 0x1000u && (h & 1u))) ++h;
    return (unsigned short)(sign | h);
}
float f16_to_f32(unsigned short b) {
    const int e = (b >> 10) & 0x1f, m = b & 0x3ff;
    const float v = e ? std::ldexp((float)(m | 0x400), e - 25) : std::ldexp((float)m, -24);
    return (b & 0x8000) ? -v : v;
}

// Geometry and B fragments of fir_tc_kernel for the reversed real taps w[0..T).
int plan_tc(rrc_fir* h, const std::vector<float>& w) {
    h->tc = false;
    h->tc1 = false;
    h->tc_cplx = false;
    if (h->tc_bfrag) { cudaFree(h->tc_bfrag); h->tc_bfrag = nullptr; }
    h->tc5_tab.clear();
    const size_t T = h->ntaps, D = h->deci;
    if (!h->real_taps || (h->flags & (RRC_FIR_NO_TENSOR | RRC_FIR_FORCE_GENERIC)) || T < 16 || D > 512) return RRC_OK;
    if (const char* e = getenv("RRC_FIR_TENSOR")) if (atoi(e) == 0) return RRC_OK;
    for (float v : w) if (!std::isfinite(v)) return RRC_OK;
    auto ksteps = [&](int ntile) { return (int)(((size_t)(8 * ntile - 1) * D + T + 15) / 16); };
    // MMAs per output are 3*KS/64 whatever NTILE is; a wider block-row only saves ldmatrix traffic,
    // so widen while the k-range grows by less than 10 %.
    int ntile = 1;
    for (int c : {2, 4}) if (ksteps(c) * 10 <= ksteps(1) * 11) ntile = c;
    if (const char* e = getenv("RRC_FIR_TC_NTILE")) { int v = atoi(e); if (v == 1 || v == 2 || v == 4) ntile = v; }
    const size_t limit_hi = (size_t)max_smem_optin(h->device);
    // A warp tile = NM m-tiles of 8 block-rows; its input span must fit the lanes' registers (64 * NLD samples).
    // The conversion to fp16 planes costs ~20 instructions per staged sample, so the path is taken when a sample
    // feeds many taps (ntaps/deci >= 32) and the halo is at most half of the staged span.
    bool force = false;
    if (const char* e = getenv("RRC_FIR_TENSOR")) force = atoi(e) == 2;
    // deci 1, 2, 4 and <= 20 k-steps: the walk kernel — measured ahead of the FP32 kernels from 16 taps on (tools/fir_sweep.py) (even k-step counts only for deci 2 and 4: fewer instantiations,
    // the extra k-step multiplies zero taps)
    const int ks1 = (D == 1) ? ksteps(1) : ((ksteps(1) + 1) & ~1);
    h->tc1 = (D == 1 || D == 2 || D == 4 || D == 8) && ks1 <= FIR_TC1_MAX_KS;
    if (const char* e = getenv("RRC_FIR_TC1")) if (atoi(e) == 0) h->tc1 = false;
    if (h->tc1) {
        h->tc = true;
        h->tc_ntile = 1; h->tc_KS = ks1;
    }
    if (!h->cplx && !h->tc1) return RRC_OK;               // f32 streams: walk kernel (fir_tcf_kernel) or FP32
    if (!h->tc1 && !force && T < 32 * D) return RRC_OK;   // the generic kernel only pays for ntaps/deci >= 32
    for (int pass = 0; pass < 2 && !h->tc; ++pass)
        for (int nt = ntile; nt >= 1 && !h->tc; nt >>= 1) {
            const int R = 8 * nt, KS = ksteps(nt);
            const size_t RS = (size_t)R * D;
            const int PAD = ((RS / 8) % 2 == 0) ? 8 : 0;
            const size_t halo = std::max<size_t>((size_t)16 * KS > RS ? (size_t)16 * KS - RS : 0, T);
            int nld = (halo * 6 <= (size_t)64 * 9 && 8 * RS + halo <= (size_t)64 * 9) ? 9 : 14;
            if (const char* e = getenv("RRC_FIR_TC_NLD")) { int v = atoi(e); if (v == 9 || v == 14) nld = v; }
            const size_t lmax = (size_t)64 * nld;
            if (RS >= 65536 || 8 * RS + halo > lmax) continue;
            size_t nm = (lmax - halo) / (8 * RS);
            if (!force && halo > nm * 8 * RS) continue;
            if (const char* e = getenv("RRC_FIR_TC_NM")) { int v = atoi(e); if (v >= 1 && (size_t)v < nm) nm = (size_t)v; }
            size_t L = nm * 8 * RS + halo;
            L = (L + 7) & ~(size_t)7;
            const size_t chunks = (L + RS - 1) / RS;
            const size_t PL = (L + chunks * PAD + 7) & ~(size_t)7;
            const size_t wb = PL * 8 + (((nm * 8 * R + 2) * 8 + 15) & ~(size_t)15);      // ytile counted for both epilogues
            const size_t smem = (size_t)KS * nt * 512 + (FIR_TC_THREADS / 32) * wb;
            if (smem > (pass == 0 ? (size_t)74 * 1024 : limit_hi)) continue;
            h->tc = true;
            h->tc_ntile = nt; h->tc_nld = nld; h->tc_nm = (int)nm; h->tc_KS = KS; h->tc_RS = (int)RS; h->tc_PAD = PAD;
            h->tc_L = (int)L; h->tc_PL = (int)PL; h->tc_wb = (int)wb; h->tc_smem = smem;
            h->tc_magic = (unsigned)((0x100000000ull + RS - 1) / RS);
        }
    if (!h->tc) return RRC_OK;
    // B[k][n] = w[k - n*D]; per lane (g = lane >> 2, t = lane & 3): column n = 8*nt + g,
    // register 0 = rows k0 + 2t, +1, register 1 = rows k0 + 2t + 8, +9 (lower k in the lower half).
    const int KS = h->tc_KS;
    ntile = h->tc_ntile;
    // taps scaled by the power of two that puts the largest one in [2^13, 2^14), then split hi + lo in fp16
    float wmax = 0.0f;
    for (float v : w) wmax = std::max(wmax, std::fabs(v));
    int we = 0;
    if (wmax > 0.0f) std::frexp(wmax, &we);               // wmax = m * 2^we, m in [0.5, 1)
    const int shift = wmax > 0.0f ? 14 - we : 0;
    if (shift > 100 || shift < -100) { h->tc = false; return RRC_OK; }
    h->tc_tap_inv_scale = std::ldexp(1.0f, -shift);
    std::vector<unsigned> frag((size_t)KS * ntile * 32 * 4);
    auto Bval = [&](long long k, long long n) -> float {
        const long long j = k - n * (long long)D;
        return (j >= 0 && j < (long long)T) ? std::ldexp(w[(size_t)j], shift) : 0.0f;
    };
    for (int ks = 0; ks < KS; ++ks)
        for (int nt = 0; nt < ntile; ++nt)
            for (int lane = 0; lane < 32; ++lane) {
                const int g = lane >> 2, t = lane & 3;
                const long long n = 8 * nt + g, k0 = 16ll * ks + 2 * t;
                unsigned* f = &frag[(((size_t)ks * ntile + nt) * 32 + lane) * 4];
                for (int r = 0; r < 2; ++r) {
                    unsigned hi = 0, lo = 0;
                    for (int e = 0; e < 2; ++e) {
                        const float v = Bval(k0 + 8 * r + e, n);
                        const unsigned short vh = f16_rn(v);
                        const unsigned short vl = f16_rn(v - f16_to_f32(vh));
                        hi |= (unsigned)vh << (16 * e);
                        lo |= (unsigned)vl << (16 * e);
                    }
                    f[r] = hi;
                    f[2 + r] = lo;
                }
            }
    RRC_CUDA(cudaMalloc(&h->tc_bfrag, frag.size() * sizeof(unsigned)));
    RRC_CUDA(upload_sync(h->tc_bfrag, frag.data(), frag.size() * sizeof(unsigned)));
    // tcgen05 kernel (fir_tc5.cu): c32 samples, deci 1, ntaps <= 65 (k = m + j <= 127 + 64 < 192).  One persistent CTA per SM
    // works through 8192-output tiles behind a three-stage pipeline whose fill costs about two tile times, so it is taken
    // when a launch gives every SM at least 3 tiles of 8192 outputs (measured against fir_tc1_kernel: 13 % ahead at 3.5, 16 % at 6.9,
    // 10 % at config 1's 13.8); below about 12 per SM the tiles are 4096 outputs (fir_tc5_rows).
    // RRC_FIR_TCGEN05: 0 never, 2 always (tests), default by size.
    int want5 = 1;
    if (const char* e = getenv("RRC_FIR_TCGEN05")) want5 = atoi(e);
    if (want5 && D == 1 && T <= 65) {
        std::vector<unsigned short> hi(T), lo(T);
        for (size_t j = 0; j < T; ++j) {
            const float v = Bval((long long)j, 0);
            hi[j] = f16_rn(v);
            lo[j] = f16_rn(v - f16_to_f32(hi[j]));
        }
        h->tc5_tab.assign(fir_tc5_tab_words(), 0u);
        fir_tc5_build_tab(hi.data(), lo.data(), T, h->tc5_tab.data());
        h->tc5_KS = (int)((127 + T + 15) / 16);
        h->tc5_min_tiles = want5 == 2 ? 0 : 3ll * sm_count(h->device);
    }
    return RRC_OK;
}

// Complex taps (what translate() produces): B fragments of Re(w) and Im(w) for fir_tcc_kernel, one common scale.
int plan_tc_cplx(rrc_fir* h, const std::vector<float>& w2) {       // w2: reversed taps, (re, im) interleaved
    h->tc = false;
    h->tc1 = false;
    h->tc_cplx = false;
    if (h->tc_bfrag) { cudaFree(h->tc_bfrag); h->tc_bfrag = nullptr; }
    const size_t T = h->ntaps, D = h->deci;
    if (!h->cplx || (h->flags & (RRC_FIR_NO_TENSOR | RRC_FIR_FORCE_GENERIC)) || T < 16) return RRC_OK;
    bool force = false;
    if (const char* e = getenv("RRC_FIR_TENSOR")) { if (atoi(e) == 0) return RRC_OK; force = atoi(e) == 2; }
    (void)force;
    if (!(D == 1 || D == 2 || D == 4 || D == 8)) return RRC_OK;
    for (float v : w2) if (!std::isfinite(v)) return RRC_OK;
    int KS = (int)((7 * D + T + 15) / 16);
    if (D != 1) KS = (KS + 1) & ~1;
    if (KS > FIR_TC1_MAX_KS) return RRC_OK;
    float wmax = 0.0f;
    for (float v : w2) wmax = std::max(wmax, std::fabs(v));
    int we = 0;
    if (wmax > 0.0f) std::frexp(wmax, &we);
    const int shift = wmax > 0.0f ? 14 - we : 0;
    if (shift > 100 || shift < -100) return RRC_OK;
    std::vector<unsigned> frag((size_t)KS * 2 * 32 * 4);
    for (int ks = 0; ks < KS; ++ks)
        for (int part = 0; part < 2; ++part)
            for (int lane = 0; lane < 32; ++lane) {
                const int g = lane >> 2, t = lane & 3;
                unsigned* f = &frag[(((size_t)ks * 2 + part) * 32 + lane) * 4];
                for (int r = 0; r < 2; ++r) {
                    unsigned hi = 0, lo = 0;
                    for (int e = 0; e < 2; ++e) {
                        const long long j = 16ll * ks + 2 * t + 8 * r + e - (long long)g * (long long)D;
                        const float v = (j >= 0 && j < (long long)T) ? std::ldexp(w2[2 * (size_t)j + part], shift) : 0.0f;
                        const unsigned short vh = f16_rn(v);
                        const unsigned short vl = f16_rn(v - f16_to_f32(vh));
                        hi |= (unsigned)vh << (16 * e);
                        lo |= (unsigned)vl << (16 * e);
                    }
                    f[r] = hi;
                    f[2 + r] = lo;
                }
            }
    RRC_CUDA(cudaMalloc(&h->tc_bfrag, frag.size() * sizeof(unsigned)));
    RRC_CUDA(upload_sync(h->tc_bfrag, frag.data(), frag.size() * sizeof(unsigned)));
    h->tc = h->tc1 = h->tc_cplx = true;
    h->tc_ntile = 1; h->tc_KS = KS;
    h->tc_tap_inv_scale = std::ldexp(1.0f, -shift);
    return RRC_OK;
}

// (Re)build the device tap tables from taps_host and pick the launch geometry.
int upload_taps(rrc_fir* h) {
    const size_t T = h->ntaps, D = h->deci;
    RRC_CUDA(cudaSetDevice(h->device));
    if (h->taps_poly) { cudaFree(h->taps_poly); h->taps_poly = nullptr; }
    if (h->taps_rev) { cudaFree(h->taps_rev); h->taps_rev = nullptr; }

    bool all_real = true;
    if (h->cplx)
        for (size_t k = 0; k < T; ++k)
            if (h->taps_host[2 * k + 1] != 0.0f) { all_real = false; break; }
    h->real_taps = !h->cplx || (all_real && !(h->flags & RRC_FIR_NO_REAL_TAP_FASTPATH));
    const size_t te = h->real_taps ? 1 : 2;   // floats per tap in the device tables

    const size_t Q = (T + D - 1) / D;
    // outputs per thread: 16 for the real-tap, non-decimating, >= 32-tap case (c32 or f32), else 8
    h->R = (h->real_taps && D == 1 && T >= 32) ? 16 : FIR_R;
    h->qpad = (int)((Q + h->R - 1) / h->R * h->R);
    h->nchunks = h->qpad / h->R;

    auto tap_at = [&](size_t j, float* dst) {   // h'[j] = taps[T-1-j]
        const size_t k = T - 1 - j;
        if (h->cplx) {
            dst[0] = h->taps_host[2 * k];
            if (te == 2) dst[1] = h->taps_host[2 * k + 1];
        } else {
            dst[0] = h->taps_host[k];
        }
    };
    std::vector<float> rev(T * te), poly((size_t)D * h->qpad * te, 0.0f);
    for (size_t j = 0; j < T; ++j) {
        tap_at(j, &rev[j * te]);
        const size_t q = j / D, p = j % D;
        tap_at(j, &poly[(p * h->qpad + q) * te]);
    }
    RRC_CUDA(cudaMalloc(&h->taps_rev, rev.size() * sizeof(float)));
    RRC_CUDA(upload_sync(h->taps_rev, rev.data(), rev.size() * sizeof(float)));
    RRC_CUDA(cudaMalloc(&h->taps_poly, poly.size() * sizeof(float)));
    RRC_CUDA(upload_sync(h->taps_poly, poly.data(), poly.size() * sizeof(float)));

    // Geometry: largest CTA whose tile fits; prefer <= 100 KB so two CTAs share an SM.
    h->use_poly = false;
    // Packed-FP32 kernel for decimating real-tap c32 filters.  For deci == 1 (config 1) the FFMA
    // kernel with R = 16 measured 9 % faster (150.6 vs 137.4 Gsps), so it keeps that path
    // (RRC_FIR_FFMA2=1 forces the packed kernel there too, =0 disables it everywhere).
    h->rt = h->cplx && h->real_taps && D >= 2;
    if (const char* e = getenv("RRC_FIR_FFMA2")) h->rt = h->cplx && h->real_taps && atoi(e) != 0;
    // two threads per output group when the staged tile per group is large (R*deci*8 bytes)
    h->split = (h->rt && D >= 4) ? 2 : 1;
    if (const char* e = getenv("RRC_FIR_SPLIT")) { int v = atoi(e); if (h->rt && (v == 1 || (v == 2 && D >= 2))) h->split = v; }
    // Uniform-tap unrolled kernel (fir_rtu_kernel) for the instantiated shapes: c32 samples, real taps,
    // decimation 10 or 5 with at most 26 taps per polyphase branch (config 3 is 255 taps / 10 -> 26).
    h->rtu_qb = 0;
    if (h->cplx && h->real_taps && !(h->flags & RRC_FIR_FORCE_GENERIC) && (D == 10 || D == 5) && Q <= 26) {
        bool on = true;
        if (const char* e = getenv("RRC_FIR_RTU")) on = atoi(e) != 0;
        if (on) {
            const int QB = Q <= 13 ? 13 : 26;
            h->rtu_qb = QB;
            h->rtu_hh.assign((size_t)D * QB, make_float2(0.f, 0.f));
            for (size_t j = 0; j < T; ++j) {
                float v; tap_at(j, &v);
                h->rtu_hh[(j % D) * QB + j / D] = make_float2(v, v);
            }
            h->rtu_r = 8;
            if (const char* e = getenv("RRC_FIR_RTU_R")) { int v = atoi(e); if (v == 8 || v == 16) h->rtu_r = v; }
            h->rtu_nt = h->rtu_r == 16 ? 32 : 64;
            if (const char* e = getenv("RRC_FIR_RTU_NT")) { int v = atoi(e); if (v == 32 || v == 64 || v == 128) h->rtu_nt = v; }
            const size_t S1u = (size_t)h->rtu_r * D + 1;
            const size_t extra = ((size_t)QB + h->rtu_r - 2) / h->rtu_r;         // segments a thread's window reaches past its own
            h->rtu_smem = (((size_t)(h->rtu_nt + extra) * S1u + 1) & ~(size_t)1) * sizeof(float2);
            if (h->rtu_smem > (size_t)max_smem_optin(h->device)) h->rtu_qb = 0;
        }
    }
    if (!(h->flags & RRC_FIR_FORCE_GENERIC) && D <= (1u << 20)) {
        const size_t S1 = (size_t)h->R * D + 1;
        const size_t tap_bytes = ((size_t)D * h->qpad * (h->rt ? sizeof(float2) : tap_elem(h)) + 15) & ~(size_t)15;
        const size_t limit_hi = (size_t)max_smem_optin(h->device);
        // Large decimations make the staged tile big (R*deci samples per group): prefer CTAs of
        // <= 48 KB so that >= 4 of them share an SM and their load / compute / store phases overlap.
        const size_t limits[3] = {48 * 1024, 100 * 1024, limit_hi};
        const int nt_min[3] = {64, 32, 32};
        // Measured on config 1 (R = 16): 64-thread CTAs (1024 outputs) beat 256-thread ones by 9 %
        // because more, smaller CTAs interleave their tile loads with their neighbours' FMAs.
        int nt_max = h->R == 16 ? 64 : FIR_MAX_NT / h->split;
        if (const char* e = getenv("RRC_FIR_NT")) { int v = atoi(e); if (v == 32 || v == 64 || v == 128 || v == 256) nt_max = std::min(v, FIR_MAX_NT / h->split); }
        for (int pass = 0; pass < 3 && !h->use_poly; ++pass) {
            for (int g = nt_max; g >= std::min(nt_min[pass], nt_max); g >>= 1) {
                const size_t tile = (((size_t)(g + h->nchunks) * S1 + 1) & ~(size_t)1) * samp_elem(h);
                if (tap_bytes + tile <= limits[pass]) {
                    h->nbuf = 1;
                    h->groups = g; h->nt = g * h->split; h->smem = tap_bytes + h->nbuf * tile; h->use_poly = true;
                    break;
                }
            }
        }
    }
    if (h->real_taps) RRC_TRY(plan_tc(h, rev));
    else RRC_TRY(plan_tc_cplx(h, rev));
    return RRC_OK;
}

template <typename ST, typename TT, int DCT, bool DEMOD, int R>
int launch_poly_r(const rrc_fir* h, const FirArgs& a, dim3 grid, cudaStream_t st) {
    auto k = fir_poly_kernel<ST, TT, DCT, DEMOD, R>;
    RRC_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->smem));
    k<<<grid, h->nt, h->smem, st>>>(a);
    RRC_CHECK_LAUNCH();
    count_launch();
    return RRC_OK;
}

template <typename ST, typename TT, int DCT, bool DEMOD>
int launch_poly_d(const rrc_fir* h, const FirArgs& a, dim3 grid, cudaStream_t st) {
    if constexpr (DCT == 1 && std::is_same<TT, float>::value) {
        if (h->R == 16) return launch_poly_r<ST, TT, DCT, DEMOD, 16>(h, a, grid, st);
    }
    return launch_poly_r<ST, TT, DCT, DEMOD, FIR_R>(h, a, grid, st);
}

template <int DCT, bool DEMOD, int R, int SPLIT>
int launch_rt_k(const rrc_fir* h, const FirArgs& a, dim3 grid, cudaStream_t st) {
    auto k = fir_rt_kernel<DCT, DEMOD, R, SPLIT>;
    RRC_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->smem));
    k<<<grid, h->nt, h->smem, st>>>(a);
    RRC_CHECK_LAUNCH();
    count_launch();
    return RRC_OK;
}
template <int DCT, bool DEMOD>
int launch_rt_d(const rrc_fir* h, const FirArgs& a, dim3 grid, cudaStream_t st) {
    if constexpr (DCT == 1) {
        return h->R == 16 ? launch_rt_k<1, DEMOD, 16, 1>(h, a, grid, st) : launch_rt_k<1, DEMOD, FIR_R, 1>(h, a, grid, st);
    } else {
        return h->split == 2 ? launch_rt_k<DCT, DEMOD, FIR_R, 2>(h, a, grid, st) : launch_rt_k<DCT, DEMOD, FIR_R, 1>(h, a, grid, st);
    }
}
template <int DCT, int QB, bool DEMOD, int R>
int launch_rtu_r(const rrc_fir* h, const FirArgs& a, dim3 grid, cudaStream_t st) {
    auto k = fir_rtu_kernel<DCT, QB, DEMOD, R>;
    RtuTaps<DCT * QB> tp;
    memcpy(tp.hh, h->rtu_hh.data(), sizeof(tp.hh));
    RRC_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->rtu_smem));
    k<<<grid, h->rtu_nt, h->rtu_smem, st>>>(a, tp);
    RRC_CHECK_LAUNCH();
    count_launch();
    return RRC_OK;
}
template <int DCT, int QB, bool DEMOD>
int launch_rtu_k(const rrc_fir* h, const FirArgs& a, dim3 grid, cudaStream_t st) {
    return h->rtu_r == 16 ? launch_rtu_r<DCT, QB, DEMOD, 16>(h, a, grid, st) : launch_rtu_r<DCT, QB, DEMOD, 8>(h, a, grid, st);
}
template <bool DEMOD>
int launch_rtu(const rrc_fir* h, const FirArgs& a, dim3 grid, cudaStream_t st) {
    if (h->deci == 10) return h->rtu_qb == 13 ? launch_rtu_k<10, 13, DEMOD>(h, a, grid, st) : launch_rtu_k<10, 26, DEMOD>(h, a, grid, st);
    return h->rtu_qb == 13 ? launch_rtu_k<5, 13, DEMOD>(h, a, grid, st) : launch_rtu_k<5, 26, DEMOD>(h, a, grid, st);
}

template <bool DEMOD>
int launch_rt(const rrc_fir* h, const FirArgs& a, dim3 grid, cudaStream_t st) {
    switch (h->deci) {
    case 1: return launch_rt_d<1, DEMOD>(h, a, grid, st);
    case 2: return launch_rt_d<2, DEMOD>(h, a, grid, st);
    case 4: return launch_rt_d<4, DEMOD>(h, a, grid, st);
    case 5: return launch_rt_d<5, DEMOD>(h, a, grid, st);
    case 8: return launch_rt_d<8, DEMOD>(h, a, grid, st);
    case 10: return launch_rt_d<10, DEMOD>(h, a, grid, st);
    default: return launch_rt_d<0, DEMOD>(h, a, grid, st);
    }
}

template <typename ST, typename TT, bool DEMOD>
int launch_poly(const rrc_fir* h, const FirArgs& a, dim3 grid, cudaStream_t st) {
    switch (h->deci) {       // common decimations get immediate window offsets
    case 1: return launch_poly_d<ST, TT, 1, DEMOD>(h, a, grid, st);
    case 2: return launch_poly_d<ST, TT, 2, DEMOD>(h, a, grid, st);
    case 4: return launch_poly_d<ST, TT, 4, DEMOD>(h, a, grid, st);
    case 5: return launch_poly_d<ST, TT, 5, DEMOD>(h, a, grid, st);
    case 8: return launch_poly_d<ST, TT, 8, DEMOD>(h, a, grid, st);
    case 10: return launch_poly_d<ST, TT, 10, DEMOD>(h, a, grid, st);
    default: return launch_poly_d<ST, TT, 0, DEMOD>(h, a, grid, st);
    }
}

int run_impl(rrc_fir* h, const void* in, size_t in_stride, size_t need, void* out, size_t out_stride,
             size_t out_n, size_t nchan, bool demod, float gain, void* stream) {
    if (!h) return fail(RRC_ERR_INVALID, "fir handle is NULL");
    if (out_n == 0 || nchan == 0) return RRC_OK;
    if (!in || !out) return fail(RRC_ERR_INVALID, "in/out is NULL");
    if (need < (out_n - 1) * h->deci + h->ntaps)
        return fail(RRC_ERR_INVALID, "fir: need %zu < (out_n-1)*deci+ntaps = %zu", need, (out_n - 1) * h->deci + h->ntaps);
    if (nchan > 65535) return fail(RRC_ERR_INVALID, "fir: nchan %zu > 65535", nchan);
    if (demod && !h->cplx) return fail(RRC_ERR_INVALID, "fused demod needs a c32 FIR");
    RRC_CUDA(cudaSetDevice(h->device));
    cudaStream_t st = as_stream(stream);

    FirArgs a{};
    a.in = in; a.out = out;
    a.in_stride = (long long)in_stride; a.out_stride = (long long)out_stride;
    a.need = (long long)need; a.out_n = (long long)out_n;
    a.ntaps = (int)h->ntaps; a.deci = (int)h->deci;
    a.qpad = h->qpad; a.nchunks = h->nchunks;
    a.S = h->R * (int)h->deci; a.nseg = h->groups + h->nchunks;
    a.gain = gain;
    a.translate = h->translate ? 1 : 0;
    a.ratio = h->ratio;
    a.out_base = h->out_counter;
    a.in_u8 = h->in_u8;
    a.epi = h->epi;
    if (h->epi.kind != RRC_EPI_NONE && demod) return fail(RRC_ERR_INVALID, "a store epilogue cannot be combined with the fused demod (its gain is the MultiplyConst)");
    if (h->in_u8 && (reinterpret_cast<uintptr_t>(in) & 1)) return fail(RRC_ERR_INVALID, "u8 I/Q input must be 2-byte aligned");

    const bool use_tc = h->tc && (h->tc_cplx || !h->translate) && h->epi.kind == RRC_EPI_NONE;
    if (use_tc && h->tc_cplx) {
        const size_t work = demod ? out_n - 1 : out_n;
        if (work == 0) return RRC_OK;
        FirTccArgs t{};
        t.in = reinterpret_cast<const float2*>(in); t.out = out;
        t.bfrag = reinterpret_cast<const uint4*>(h->tc_bfrag);
        t.taps_rev_c = reinterpret_cast<const float2*>(h->taps_rev);
        t.in_stride = (long long)in_stride; t.out_stride = (long long)out_stride;
        t.need = (long long)need; t.out_n = (long long)out_n;
        t.ntaps = (int)h->ntaps; t.gain = gain; t.tap_inv_scale = h->tc_tap_inv_scale;
        t.translate = h->translate ? 1 : 0; t.ratio = h->ratio; t.out_base = h->out_counter; t.in_u8 = h->in_u8;
        const size_t bt1 = FIR_TC1_BT / h->deci;
        t.tiles_x = (long long)((work + bt1 - 1) / bt1);
        t.total_tiles = t.tiles_x * (long long)nchan;
        if (t.total_tiles > 0x7fffffffll) return fail(RRC_ERR_INVALID, "fir: %lld tiles in one launch (limit 2^31 - 1)", t.total_tiles);
        RRC_TRY(fir_tcc_launch(FirTcGeom{h->device, 1, 0, h->tc_KS, 0, (int)h->deci}, t, demod, st));
    } else if (use_tc && !h->tc5_tab.empty() && !demod && !h->in_u8 &&
               (long long)((out_n + 8191) / 8192) * (long long)nchan >= h->tc5_min_tiles) {
        FirTc5Args t{};
        t.in = reinterpret_cast<const float2*>(in); t.out = reinterpret_cast<float2*>(out);
        t.in_stride = (long long)in_stride; t.out_stride = (long long)out_stride;
        t.need = (long long)need; t.out_n = (long long)out_n;
        t.KS = h->tc5_KS; t.tap_inv_scale = h->tc_tap_inv_scale; t.real_stream = h->cplx ? 0 : 1;
        t.nr = fir_tc5_rows((long long)((out_n + 8191) / 8192) * (long long)nchan, h->device, !h->cplx);
        const size_t bt5 = (size_t)128 * t.nr;
        t.tiles_x = (long long)((out_n + bt5 - 1) / bt5);
        t.total_tiles = t.tiles_x * (long long)nchan;
        RRC_TRY(fir_tc5_launch(h->device, t, h->tc5_tab.data(), st));
        h->last_tc5 = 1;
    } else if (use_tc && !h->cplx) {
        if (demod) return fail(RRC_ERR_INVALID, "fused demod needs a c32 FIR");
        h->last_tc5 = 0;
        FirTcfArgs t{};
        t.in = reinterpret_cast<const float*>(in); t.out = reinterpret_cast<float*>(out);
        t.bfrag = reinterpret_cast<const uint4*>(h->tc_bfrag);
        t.in_stride = (long long)in_stride; t.out_stride = (long long)out_stride;
        t.need = (long long)need; t.out_n = (long long)out_n;
        t.tap_inv_scale = h->tc_tap_inv_scale;
        const size_t btf = FIR_TCF_IN / h->deci;
        t.tiles_x = (long long)((out_n + btf - 1) / btf);
        t.total_tiles = t.tiles_x * (long long)nchan;
        if (t.total_tiles > 0x7fffffffll) return fail(RRC_ERR_INVALID, "fir: %lld tiles in one launch (limit 2^31 - 1)", t.total_tiles);
        RRC_TRY(fir_tcf_launch(FirTcGeom{h->device, 1, 0, h->tc_KS, 0, (int)h->deci}, t, st));
    } else if (use_tc && h->tc1) {
        const size_t work = demod ? out_n - 1 : out_n;
        if (work == 0) return RRC_OK;
        h->last_tc5 = 0;
        FirTc1Args t{};
        t.in = reinterpret_cast<const float2*>(in); t.out = out;
        t.bfrag = reinterpret_cast<const uint4*>(h->tc_bfrag);
        t.taps_rev = reinterpret_cast<const float*>(h->taps_rev);
        t.in_stride = (long long)in_stride; t.out_stride = (long long)out_stride;
        t.need = (long long)need; t.out_n = (long long)out_n;
        t.ntaps = (int)h->ntaps; t.gain = gain; t.tap_inv_scale = h->tc_tap_inv_scale; t.in_u8 = h->in_u8;
        const size_t bt1 = FIR_TC1_BT / h->deci;
        t.tiles_x = (long long)((work + bt1 - 1) / bt1);
        t.total_tiles = t.tiles_x * (long long)nchan;
        if (t.total_tiles > 0x7fffffffll) return fail(RRC_ERR_INVALID, "fir: %lld tiles in one launch (limit 2^31 - 1)", t.total_tiles);
        RRC_TRY(fir_tc1_launch(FirTcGeom{h->device, 1, 0, h->tc_KS, 0, (int)h->deci}, t, demod, st));
    } else if (use_tc) {
        const size_t work = demod ? out_n - 1 : out_n;
        if (work == 0) return RRC_OK;
        FirTcArgs t{};
        t.in = reinterpret_cast<const float2*>(in); t.out = out;
        t.bfrag = reinterpret_cast<const uint4*>(h->tc_bfrag);
        t.taps_rev = reinterpret_cast<const float*>(h->taps_rev);
        t.in_stride = (long long)in_stride; t.out_stride = (long long)out_stride;
        t.need = (long long)need; t.out_n = (long long)out_n;
        t.ntaps = (int)h->ntaps; t.deci = (int)h->deci;
        t.RS = h->tc_RS; t.PAD = h->tc_PAD; t.magic = h->tc_magic; t.KS = h->tc_KS; t.NM = h->tc_nm; t.L = h->tc_L; t.PL = h->tc_PL; t.WB = h->tc_wb;
        t.gain = gain; t.tap_inv_scale = h->tc_tap_inv_scale; t.in_u8 = h->in_u8;
        const size_t bt = (size_t)h->tc_nm * 8 * 8 * h->tc_ntile;
        t.tiles_x = (long long)((work + bt - 1) / bt);
        t.total_tiles = t.tiles_x * (long long)nchan;
        if (t.total_tiles > 0x7fffffffll) return fail(RRC_ERR_INVALID, "fir: %lld tiles in one launch (limit 2^31 - 1)", t.total_tiles);
        RRC_TRY(fir_tc_launch(FirTcGeom{h->device, h->tc_ntile, h->tc_nld, h->tc_KS, h->tc_smem, (int)h->deci}, t, demod, st));
    } else if (h->rtu_qb) {
        const size_t bt = (size_t)h->rtu_nt * h->rtu_r;
        const size_t per = demod ? bt - 1 : bt;
        const size_t work = demod ? (out_n > 1 ? out_n - 1 : 0) : out_n;
        if (work == 0) return RRC_OK;
        a.S = h->rtu_r * (int)h->deci;
        a.nseg = h->rtu_nt + (h->rtu_qb + h->rtu_r - 2) / h->rtu_r;
        a.tiles_x = (long long)((work + per - 1) / per);
        a.total_tiles = a.tiles_x * (long long)nchan;
        a.nbuf = 1;
        dim3 grid((unsigned)a.tiles_x, (unsigned)nchan);
        RRC_TRY(demod ? launch_rtu<true>(h, a, grid, st) : launch_rtu<false>(h, a, grid, st));
    } else if (h->use_poly) {
        a.taps = h->taps_poly;
        const size_t bt = (size_t)h->groups * h->R;
        const size_t per = demod ? bt - 1 : bt;
        const size_t work = demod ? (out_n > 1 ? out_n - 1 : 0) : out_n;
        if (work == 0) return RRC_OK;
        a.tiles_x = (long long)((work + per - 1) / per);
        a.total_tiles = a.tiles_x * (long long)nchan;
        a.nbuf = h->nbuf;
        dim3 grid((unsigned)a.tiles_x, (unsigned)nchan);
        int s;
        if (h->cplx && !h->real_taps)
            s = demod ? launch_poly<float2, float2, true>(h, a, grid, st) : launch_poly<float2, float2, false>(h, a, grid, st);
        else if (h->rt)
            s = demod ? launch_rt<true>(h, a, grid, st) : launch_rt<false>(h, a, grid, st);
        else if (h->cplx)
            s = demod ? launch_poly<float2, float, true>(h, a, grid, st) : launch_poly<float2, float, false>(h, a, grid, st);
        else
            s = launch_poly<float, float, false>(h, a, grid, st);
        RRC_TRY(s);
    } else {
        a.taps = h->taps_rev;
        void* fir_out = out;
        size_t fir_stride = out_stride;
        float2* tmp = nullptr;
        if (demod) {   // unfused fallback: FIR into scratch, then demod
            RRC_CUDA(cudaMallocAsync((void**)&tmp, sizeof(float2) * out_n * nchan, st));
            fir_out = tmp; fir_stride = out_n;
            a.out = tmp; a.out_stride = (long long)out_n;
        }
        unsigned gx = (unsigned)std::min<size_t>((out_n + 255) / 256, (size_t)sm_count(h->device) * 8);
        dim3 grid(gx, (unsigned)nchan);
        if (h->cplx && !h->real_taps) fir_generic_kernel<float2, float2><<<grid, 256, 0, st>>>(a);
        else if (h->cplx) fir_generic_kernel<float2, float><<<grid, 256, 0, st>>>(a);
        else fir_generic_kernel<float, float><<<grid, 256, 0, st>>>(a);
        RRC_CHECK_LAUNCH();
        count_launch();
        if (demod) {
            if (out_n > 1) {
                unsigned dx = (unsigned)std::min<size_t>((out_n + 255) / 256, (size_t)sm_count(h->device) * 8);
                quad_demod_kernel<<<dim3(dx, (unsigned)nchan), 256, 0, st>>>((const float2*)fir_out, (long long)fir_stride,
                                                                           (long long)out_n, gain, (float*)out, (long long)out_stride);
                RRC_CHECK_LAUNCH();
                count_launch();
            }
            RRC_CUDA(cudaFreeAsync(tmp, st));
        }
    }
    if (h->translate) h->out_counter += out_n;
    return RRC_OK;
}

int create_impl(int device, const float* taps, size_t ntaps, size_t deci, unsigned flags, bool cplx, rrc_fir_t** out) {
    if (!out) return fail(RRC_ERR_INVALID, "out is NULL");
    *out = nullptr;
    if (!taps || ntaps == 0) return fail(RRC_ERR_INVALID, "FirFilter needs at least one tap (src/fir.rs:158,372)");
    if (deci == 0) return fail(RRC_ERR_INVALID, "FirFilter deci must be nonzero (src/fir.rs:319)");
    if (ntaps > (1u << 30) || deci > (1u << 30)) return fail(RRC_ERR_INVALID, "ntaps/deci too large");
    auto* h = new rrc_fir();
    h->device = device; h->cplx = cplx; h->flags = flags; h->ntaps = ntaps; h->deci = deci;
    h->taps_host.assign(taps, taps + ntaps * (cplx ? 2 : 1));
    int s = upload_taps(h);
    if (s != RRC_OK) { rrc_fir_destroy(h); return s; }
    *out = h;
    return RRC_OK;
}

}  // namespace

extern "C" {

int rrc_fir_c32_create(int device, const float* taps, size_t ntaps, size_t deci, unsigned flags, rrc_fir_t** out) {
    return create_impl(device, taps, ntaps, deci, flags, true, out);
}
int rrc_fir_f32_create(int device, const float* taps, size_t ntaps, size_t deci, unsigned flags, rrc_fir_t** out) {
    return create_impl(device, taps, ntaps, deci, flags, false, out);
}

int rrc_fir_set_translate(rrc_fir_t* h, float samp_rate, float freq) {
    if (!h) return fail(RRC_ERR_INVALID, "fir handle is NULL");
    if (!h->cplx) return fail(RRC_ERR_INVALID, "translate() exists only on FirFilterBuilder<Complex> (src/fir.rs:476)");
    if (!(samp_rate > 0.0f)) return fail(RRC_ERR_INVALID, "samp_rate must be > 0 (src/fir.rs:435)");
    if (h->translate) return fail(RRC_ERR_STATE, "translate already set");
    if (freq == 0.0f) return RRC_OK;   // src/fir.rs:438-440
    // Pre-rotate taps by the same f32 recurrence as src/fir.rs:441-449.
    const double input_step = 2.0 * M_PI * (double)freq / (double)samp_rate;
    const float sr = (float)std::cos(input_step), si = (float)std::sin(input_step);
    float pr = 1.0f, pi = 0.0f;
    for (size_t k = 0; k < h->ntaps; ++k) {
        float tr = h->taps_host[2 * k], ti = h->taps_host[2 * k + 1];
        h->taps_host[2 * k] = tr * pr - ti * pi;
        h->taps_host[2 * k + 1] = tr * pi + ti * pr;
        float nr = pr * sr - pi * si, ni = pr * si + pi * sr;
        pr = nr; pi = ni;
    }
    h->translate = true;
    h->ratio = (double)freq / (double)samp_rate;
    h->out_counter = 0;
    return upload_taps(h);
}

int rrc_fir_destroy(rrc_fir_t* h) {
    if (!h) return RRC_OK;
    cudaSetDevice(h->device);
    if (h->taps_poly) cudaFree(h->taps_poly);
    if (h->taps_rev) cudaFree(h->taps_rev);
    if (h->tc_bfrag) cudaFree(h->tc_bfrag);
    h->pipe.destroy();
    delete h;
    return RRC_OK;
}
int rrc_fir_ntaps(const rrc_fir_t* h, size_t* n) {
    if (!h || !n) return fail(RRC_ERR_INVALID, "NULL argument");
    *n = h->ntaps;
    return RRC_OK;
}
int rrc_fir_deci(const rrc_fir_t* h, size_t* d) {
    if (!h || !d) return fail(RRC_ERR_INVALID, "NULL argument");
    *d = h->deci;
    return RRC_OK;
}
int rrc_fir_uses_real_taps(const rrc_fir_t* h, int* yes) {
    if (!h || !yes) return fail(RRC_ERR_INVALID, "NULL argument");
    *yes = (h->cplx && h->real_taps) ? 1 : 0;
    return RRC_OK;
}
int rrc_fir_uses_tensor_cores(const rrc_fir_t* h, int* yes) {
    if (!h || !yes) return fail(RRC_ERR_INVALID, "null argument");
    *yes = (h->tc && (h->tc_cplx || !h->translate) && h->epi.kind == RRC_EPI_NONE) ? 1 : 0;
    return RRC_OK;
}
int rrc_fir_set_epilogue(rrc_fir_t* h, int kind, float re, float im) {
    if (!h) return fail(RRC_ERR_INVALID, "fir handle is NULL");
    if (kind < RRC_EPI_NONE || kind > RRC_EPI_MAG2) return fail(RRC_ERR_INVALID, "unknown epilogue %d", kind);
    if (kind != RRC_EPI_NONE && !h->cplx) return fail(RRC_ERR_INVALID, "store epilogues exist for Complex filters only");
    h->epi.kind = kind; h->epi.re = re; h->epi.im = im;
    return RRC_OK;
}
int rrc_fir_kernel_name(const rrc_fir_t* h, char* buf, size_t buflen) {
    if (!h || !buf || !buflen) return fail(RRC_ERR_INVALID, "null argument");
    const bool use_tc = h->tc && (h->tc_cplx || !h->translate) && h->epi.kind == RRC_EPI_NONE;
    char tmp[320];
    if (use_tc && h->tc_cplx) snprintf(tmp, sizeof tmp, "fir_tcc_kernel<KS=%d,D=%zu> (tensor cores, complex taps, fp16x3)", h->tc_KS, h->deci);
    else if (use_tc && !h->cplx && (h->tc5_tab.empty() || h->last_tc5 == 0)) snprintf(tmp, sizeof tmp, "fir_tcf_kernel<KS=%d,D=%zu> (tensor cores, f32 stream, fp16x3)", h->tc_KS, h->deci);
    else if (use_tc && !h->tc5_tab.empty() && !h->in_u8 && h->last_tc5 != 0)
        snprintf(tmp, sizeof tmp, "fir_tc5_kernel<KS=%d> (tcgen05.mma kind::f16 M128 N64 K16, taps and accumulators in TMEM, fp16x3)%s", h->tc5_KS,
                 h->last_tc5 == 1 ? "" : "; launches below 3 tiles of 8192 outputs per SM and the fused demod use the mma.sync kernel (fir_tc1_kernel / fir_tcf_kernel)");
    else if (use_tc && h->tc1) snprintf(tmp, sizeof tmp, "fir_tc1_kernel<KS=%d,D=%zu> (tensor cores, real taps, fp16x3)", h->tc_KS, h->deci);
    else if (use_tc) snprintf(tmp, sizeof tmp, "fir_tc_kernel<NTILE=%d,NLD=%d> (tensor cores, real taps, fp16x3)", h->tc_ntile, h->tc_nld);
    else if (h->rtu_qb) snprintf(tmp, sizeof tmp, "fir_rtu_kernel<D=%zu,QB=%d,R=%d> (FFMA2, taps as uniform-register operands from the kernel parameters)", h->deci, h->rtu_qb, h->rtu_r);
    else if (h->use_poly && h->rt) snprintf(tmp, sizeof tmp, "fir_rt_kernel<D=%zu,R=%d,SPLIT=%d> (FFMA2)", h->deci, h->R, h->split);
    else if (h->use_poly) snprintf(tmp, sizeof tmp, "fir_poly_kernel<D=%zu,R=%d> (FP32 %s taps)", h->deci, h->R, h->real_taps ? "real" : "complex");
    else snprintf(tmp, sizeof tmp, "fir_generic_kernel");
    snprintf(buf, buflen, "%s", tmp);
    return RRC_OK;
}
int rrc_fir_set_input_u8iq(rrc_fir_t* h, int on) {
    if (!h) return fail(RRC_ERR_INVALID, "fir handle is NULL");
    if (on && !h->cplx) return fail(RRC_ERR_INVALID, "u8 I/Q input needs a Complex FIR");
    h->in_u8 = on ? 1 : 0;
    return RRC_OK;
}

int rrc_fir_reset(rrc_fir_t* h) {
    if (!h) return fail(RRC_ERR_INVALID, "fir handle is NULL");
    h->out_counter = 0;
    return RRC_OK;
}

int rrc_fir_plan(size_t ntaps, size_t deci, size_t in_len, size_t out_free,
                 size_t* consume, size_t* need, size_t* out_n, size_t* wait_need, int* wait_on_output) {
    if (!consume || !need || !out_n || !wait_need || !wait_on_output) return fail(RRC_ERR_INVALID, "NULL argument");
    if (ntaps == 0 || deci == 0) return fail(RRC_ERR_INVALID, "ntaps/deci must be nonzero");
    *consume = *need = *out_n = *wait_need = 0;
    *wait_on_output = 0;
    const size_t absolute_minimum = ntaps + deci - 1;                 // src/fir.rs:497
    if (in_len < absolute_minimum) { *wait_need = absolute_minimum; return RRC_OK; }   // :498-500
    size_t n = deci * ((in_len - ntaps + 1) / deci);                  // :502
    if (out_free < 1) { *wait_need = 1; *wait_on_output = 1; return RRC_OK; }          // :511-515
    n = std::min(n, out_free * deci);                                 // :518
    *consume = n;
    *out_n = n / deci;                                                // :525
    *need = n + ntaps - 1;                                            // :507
    return RRC_OK;
}

int rrc_fir_run(rrc_fir_t* h, const void* in, size_t need, void* out, size_t out_n, void* stream) {
    return run_impl(h, in, 0, need, out, 0, out_n, 1, false, 0.f, stream);
}
int rrc_fir_run_batch(rrc_fir_t* h, const void* in, size_t in_stride, size_t need, void* out, size_t out_stride,
                      size_t out_n, size_t nchan, void* stream) {
    return run_impl(h, in, in_stride, need, out, out_stride, out_n, nchan, false, 0.f, stream);
}
int rrc_fir_c32_demod_run_batch(rrc_fir_t* h, const void* in, size_t in_stride, size_t need, float gain,
                                float* out, size_t out_stride, size_t out_n, size_t nchan, void* stream) {
    return run_impl(h, in, in_stride, need, out, out_stride, out_n, nchan, true, gain, stream);
}

int rrc_fir_run_host(rrc_fir_t* h, const void* in_host, size_t n_in, void* out_host, size_t* n_out) {
    if (!h) return fail(RRC_ERR_INVALID, "fir handle is NULL");
    const size_t T = h->ntaps, D = h->deci, es = samp_elem(h);
    const size_t ies = h->in_u8 ? 2 : es;                               // input bytes per sample
    const size_t total = n_in < T + D - 1 ? 0 : (n_in - T + 1) / D;     // src/fir.rs:496-525 to exhaustion
    if (n_out) *n_out = total;
    if (total == 0) return RRC_OK;
    if (!in_host || !out_host) return fail(RRC_ERR_INVALID, "in/out is NULL");
    RRC_TRY(h->pipe.init(h->device));
    const size_t chunk_out = std::max<size_t>(1, pipe_chunk_samples_for(n_in) / D);
    const size_t max_out = std::min(chunk_out, total);
    const size_t oes = h->epi.kind == RRC_EPI_MAG2 ? sizeof(float) : es;   // ComplexToMag2 epilogue: f32 out
    RRC_TRY(h->pipe.reserve(((max_out - 1) * D + T) * ies, max_out * oes));
    int i = 0;
    for (size_t o = 0, no = 0; o < total; o += no, ++i) {
        no = pipe_next_chunk((size_t)i, total - o, chunk_out);
        const size_t need = (no - 1) * D + T;                              // halo = ntaps-1 re-copied per chunk
        RRC_TRY(h->pipe.stage_in(i, (const char*)in_host + o * D * ies, need * ies));
        RRC_TRY(run_impl(h, h->pipe.d_in[i & 1], 0, need, h->pipe.d_out[i & 1], 0, no, 1, false, 0.f, h->pipe.s_comp));
        RRC_TRY(h->pipe.drain_out(i, (char*)out_host + o * oes, no * oes));
    }
    return h->pipe.finish();
}

// Host-buffer form of the fused FirFilter<Complex> -> QuadratureDemod channelizer (config 3 end to end):
// nchan channels of n_in samples each (channel c at in_host + c*n_in samples; u8 I/Q pairs after
// rrc_fir_set_input_u8iq), whole channels staged in groups so that a chunk stays near the pipeline's
// chunk size; every channel yields floor((n_in-ntaps+1)/deci) - 1 floats at out_host + c*out_stride.
int rrc_fir_c32_demod_run_host_batch(rrc_fir_t* h, const void* in_host, size_t n_in, size_t nchan, float gain,
                                     float* out_host, size_t out_stride, size_t* n_out_per_chan) {
    if (!h) return fail(RRC_ERR_INVALID, "fir handle is NULL");
    if (!h->cplx) return fail(RRC_ERR_INVALID, "fused demod needs a c32 FIR");
    const size_t T = h->ntaps, D = h->deci;
    const size_t ies = h->in_u8 ? 2 : sizeof(float2);
    const size_t fir_n = n_in < T + D - 1 ? 0 : (n_in - T + 1) / D;       // src/fir.rs:496-525 to exhaustion
    const size_t per = fir_n ? fir_n - 1 : 0;                             // src/quadrature_demod.rs:71-73: N -> N-1
    if (n_out_per_chan) *n_out_per_chan = per;
    if (per == 0 || nchan == 0) return RRC_OK;
    if (!in_host || !out_host) return fail(RRC_ERR_INVALID, "in/out is NULL");
    if (out_stride < per) return fail(RRC_ERR_INVALID, "out_stride %zu < outputs per channel %zu", out_stride, per);
    RRC_TRY(h->pipe.init(h->device));
    const size_t need = (fir_n - 1) * D + T;
    const size_t group = std::max<size_t>(1, std::min<size_t>(nchan, PIPE_CHUNK_SAMPLES / std::max<size_t>(n_in, 1)));
    RRC_TRY(h->pipe.reserve(group * n_in * ies, group * per * sizeof(float)));
    int i = 0;
    for (size_t c = 0; c < nchan; c += group, ++i) {
        const size_t g = std::min(group, nchan - c);
        RRC_TRY(h->pipe.stage_in(i, (const char*)in_host + c * n_in * ies, g * n_in * ies));
        RRC_TRY(run_impl(h, h->pipe.d_in[i & 1], n_in, need, h->pipe.d_out[i & 1], per, fir_n, g, true, gain, h->pipe.s_comp));
        const int b = i & 1;                      // drain: contiguous when the caller's rows are packed, else row by row
        RRC_CUDA(cudaEventRecord(h->pipe.comp_done[b], h->pipe.s_comp));
        RRC_CUDA(cudaStreamWaitEvent(h->pipe.s_d2h, h->pipe.comp_done[b], 0));
        if (out_stride == per) {
            RRC_CUDA(cudaMemcpyAsync(out_host + c * per, h->pipe.d_out[b], g * per * sizeof(float), cudaMemcpyDeviceToHost, h->pipe.s_d2h));
        } else {
            RRC_CUDA(cudaMemcpy2DAsync(out_host + c * out_stride, out_stride * sizeof(float), h->pipe.d_out[b], per * sizeof(float),
                                       per * sizeof(float), g, cudaMemcpyDeviceToHost, h->pipe.s_d2h));
        }
        RRC_CUDA(cudaEventRecord(h->pipe.d2h_done[b], h->pipe.s_d2h));
    }
    return h->pipe.finish();
}

int rrc_quad_demod_run_host(int device, const float* in_host, size_t n_in, float gain, float* out_host) {
    if (n_in < 2) return RRC_OK;
    if (!in_host || !out_host) return fail(RRC_ERR_INVALID, "in/out is NULL");
    RRC_CUDA(cudaSetDevice(device));
    Pipe pipe;
    RRC_TRY(pipe.init(device));
    const size_t chunk = pipe_chunk_samples_for(n_in);
    int s = pipe.reserve((std::min(chunk, n_in - 1) + 1) * sizeof(float2), std::min(chunk, n_in - 1) * sizeof(float));
    int i = 0;
    for (size_t o = 0; s == RRC_OK && o < n_in - 1; o += chunk, ++i) {
        const size_t no = std::min(chunk, n_in - 1 - o);
        s = pipe.stage_in(i, in_host + 2 * o, (no + 1) * sizeof(float2));
        if (s == RRC_OK) s = rrc_quad_demod_run(device, (const float*)pipe.d_in[i & 1], no + 1, gain, (float*)pipe.d_out[i & 1], pipe.s_comp);
        if (s == RRC_OK) s = pipe.drain_out(i, out_host + o, no * sizeof(float));
    }
    if (s == RRC_OK) s = pipe.finish();
    pipe.destroy();
    return s;
}

int rrc_quad_demod_run_batch(int device, const float* in, size_t in_stride, size_t n_in, float gain,
                             float* out, size_t out_stride, size_t nchan, void* stream) {
    if (n_in < 2 || nchan == 0) return RRC_OK;
    if (!in || !out) return fail(RRC_ERR_INVALID, "in/out is NULL");
    if (nchan > 65535) return fail(RRC_ERR_INVALID, "nchan > 65535");
    RRC_CUDA(cudaSetDevice(device));
    unsigned gx = (unsigned)std::min<size_t>((n_in + 255) / 256, (size_t)sm_count(device) * 16);
    quad_demod_kernel<<<dim3(gx, (unsigned)nchan), 256, 0, as_stream(stream)>>>(
        (const float2*)in, (long long)in_stride, (long long)n_in, gain, out, (long long)out_stride);
    RRC_CHECK_LAUNCH();
    count_launch();
    return RRC_OK;
}
int rrc_quad_demod_run(int device, const float* in, size_t n_in, float gain, float* out, void* stream) {
    return rrc_quad_demod_run_batch(device, in, 0, n_in, gain, out, 0, 1, stream);
}

}  // extern "C"
