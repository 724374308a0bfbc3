// fftfilt_fold_core.cuh — FftFilter fused with decimate-by-8 (RationalResampler(1, 8)), computed
// with a PRUNED inverse transform: only every 8th output of the block convolution is wanted, and
//     y[8m] = (1/N) * sum_{j < N/8} ( sum_{s<8} Y[j + s*N/8] ) * W_{N/8}^{-j m}
// so the spectrum product Y = X*H is FOLDED 8:1 and the inverse FFT is 8x smaller than the
// forward one.  Two geometries share this code (template parameter NC = CTAs per cluster):
//
//   NC = 1 : N = 16384, one CTA per block (filters up to 12289 taps);
//   NC = 4 : N = 65536, one 4-CTA thread-block CLUSTER per block (BASELINE config 5:
//            16385 taps -> 49152 valid outputs per block instead of 8192 per 16384-point block).
//            CTA c computes the spectrum bins k = c (mod 4):
//              X[c + 4q] = FFT_16384( W_65536^{c n} * sum_{j<4} x[n + 16384 j] (-i)^{c j} )[q]
//            (a radix-4 decimation-in-frequency step done while LOADING: each CTA reads all four
//            quarters of the 512 KiB segment, three of the four reads are L2 hits), folds and
//            inverse-transforms its 2048 bins locally, and the four partial 2048-point results
//            are combined through DISTRIBUTED SHARED MEMORY (3 x 4 KiB remote reads per CTA).
//
// Replaces Engine::run + sum_vec (rustradio src/fft_filter.rs:172-176,281-287) followed by
// RationalResampler::work with interp = 1 (src/rational_resampler.rs:155-206).
//
// Index maps (q = local frequency index of the 16384-point transform, as fftfilt_core.cuh):
//   forward: n = n1*512 + n2*16 + n3 -> q = k1 + 32*k2 + 1024*k3   (phases A, B, C)
//   fold   : q = j2 + 2048*s, j2 = k1 + 32*k2 + 1024*b (b = k3 & 1, s = k3 >> 1): thread-local
//   inverse: 2048 = 2*32*32 over (b, k2, k1) -> m2 = n1'*64 + n2'*2 + n3'
//   combine: z[m2 + 2048*qq] = sum_c i^{c*qq} conj(W_{2048*NC}^{c*m2}) u_c[m2],  output m = m2 + 2048*qq
// All functions are __host__ __device__ so tests/emul runs the same index math on the CPU.
#pragma once
#include "fftfilt_core.cuh"

namespace rrc { namespace fftf {

using namespace rrc::fftk;   // N (16384), NT (512), phys(), powers32, BlockIO ...

constexpr int FOLD_D = 8;                 // decimation factor
constexpr int LU = N / FOLD_D;            // 2048: local inverse size (per CTA)

struct FoldIO {
    const float2* in;        // x[0..n_in)
    const float2* hist;      // previous T1_total samples
    float2* out;             // decimated outputs
    long long n_in;
    long long n_out;         // decimated outputs to write (indices >= n_out are dropped)
    int T1eff;               // (ntaps-1) rounded up to a multiple of 8: first valid segment element
    int V;                   // hop = NBIG - T1eff (multiple of 8)
    int T1_total;            // ntaps - 1 = length of hist
    int r;                   // skip mod 8: segment start alignment so kept outputs sit at n = 0 (mod 8)
    long long jbias;         // (skip - r) / 8
    int in_u8 = 0;           // 1: `in` is u8 I/Q pairs (RtlSdrDecode fused into the load)
    float2* hist_next = nullptr;   // non-NULL: the grid's last CTA writes the next call's history here (one launch per run)
    rrc::Epi epi;                  // store epilogue (epilogue.cuh); MAG2 makes `out` an f32 array
};
RRC_HD void fold_store(const FoldIO& io, long long j, float2 z) {
    if (io.epi.kind == RRC_EPI_MAG2) reinterpret_cast<float*>(io.out)[j] = rrc::epi_mag2(z);
    else io.out[j] = rrc::epi_c32(z, io.epi);
}

RRC_HD void update_history(const FoldIO& io, int tid, int nthreads) {
    for (int i = tid; i < io.T1_total; i += nthreads) {
        const long long s = io.n_in - io.T1_total + i;
        io.hist_next[i] = s >= 0 ? ld_iq(io.in, s, io.in_u8) : io.hist[s + io.T1_total];
    }
}

// p[k] = base * w^k, k = 0..31.
RRC_HD void powers32b(float2 w, float2 base, float2 (&p)[32]) {
    float2 wp[5];
    wp[0] = w;
#pragma unroll
    for (int i = 1; i < 5; ++i) wp[i] = csqr(wp[i - 1]);
    p[0] = base;
#pragma unroll
    for (int k = 1; k < 32; ++k) {
        const int low = k & (-k);
        const int rest = k & (k - 1);
        const int b = low == 1 ? 0 : low == 2 ? 1 : low == 4 ? 2 : low == 8 ? 3 : 4;
        p[k] = cmul(p[rest], wp[b]);
    }
}

template <int NC>
RRC_HD long long seg_start(long long blk, const FoldIO& io) { return blk * (long long)io.V - io.T1eff + io.r; }

RRC_HD float2 fetch(const FoldIO& io, long long g) {
    if (g < 0) return g + io.T1_total >= 0 ? io.hist[g + io.T1_total] : make_float2(0.f, 0.f);
    return g < io.n_in ? ld_iq(io.in, g, io.in_u8) : make_float2(0.f, 0.f);
}

// Phase A: load (NC = 4: with the radix-4 DIF step across the four quarters of the segment and the
// W_128^{c n1} input twiddle), DFT32 over n1, twiddle by gc[tid] * W_N^{tid k1}, write smem.
//   tw1[t] = W_16384^t (t < 512); gc[t] = W_65536^{c t} (NC = 4; unused for NC = 1);
//   twc[n1] = W_128^{c n1} (NC = 4).
struct NoHook { RRC_HD void operator()() const {} };
// before_store(): called after the loads and the DFT32, just before the results are written to the
// exchange buffer (the cluster kernel waits there for the other CTAs to finish reading this CTA's
// previous u array, which lives in the same shared memory).
template <int NC, class Hook = NoHook>
RRC_HD void phase_a(int tid, int c, long long blk, const FoldIO& io, const float2* tw1, const float2* gc,
                    const float2* twc, float2* sm, Hook before_store = Hook()) {
    float2 v[32];
    const long long seg0 = seg_start<NC>(blk, io);
    const bool interior = seg0 >= 0 && seg0 + (long long)NC * N <= io.n_in;
    if constexpr (NC == 1) {
        if (interior && io.in_u8) {
            const unsigned short* p = reinterpret_cast<const unsigned short*>(io.in) + seg0 + tid;
            unsigned int w[32];
#pragma unroll
            for (int n1 = 0; n1 < 32; ++n1) w[n1] = p[512 * n1];
#pragma unroll
            for (int n1 = 0; n1 < 32; ++n1) v[bitrev(n1, 5)] = decode_u8iq(w[n1] & 0xffu, w[n1] >> 8);
        } else if (interior) {
            const float2* p = io.in + seg0 + tid;
#pragma unroll
            for (int n1 = 0; n1 < 32; ++n1) v[bitrev(n1, 5)] = p[512 * n1];
        } else {
#pragma unroll
            for (int n1 = 0; n1 < 32; ++n1) v[bitrev(n1, 5)] = fetch(io, seg0 + tid + 512 * n1);
        }
    } else {
        // u = (-i)^c, s2 = (-1)^c
        const float ur = c == 0 ? 1.f : c == 2 ? -1.f : 0.f;
        const float ui = c == 1 ? -1.f : c == 3 ? 1.f : 0.f;
        const float s2 = (c & 1) ? -1.f : 1.f;
        auto combine = [&](int n1, const float2 (&x)[4]) {
            const float2 E = make_float2(fmaf(s2, x[2].x, x[0].x), fmaf(s2, x[2].y, x[0].y));
            const float2 O = make_float2(fmaf(s2, x[3].x, x[1].x), fmaf(s2, x[3].y, x[1].y));
            const float2 a = make_float2(fmaf(-ui, O.y, fmaf(ur, O.x, E.x)), fmaf(ui, O.x, fmaf(ur, O.y, E.y)));
            v[bitrev(n1, 5)] = cmul(a, twc[n1]);
        };
        // Software-pipelined: the 16 loads of batch b+1 (4 values of n1 x 4 quarters) are issued
        // before batch b is consumed, so every thread keeps 16..32 loads in flight (the loads are
        // L2 hits for three of the four CTAs of the cluster, ~300-800 cycles each).  `ld(o)` reads
        // segment element tid + o: a c32 load, or a 16-bit load + RtlSdrDecode for u8 I/Q input.
        auto pipelined = [&](auto ld) {
            float2 xb[2][4][4];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) xb[0][i][j] = ld(512 * i + j * N);
#pragma unroll
            for (int b = 0; b < 8; ++b) {
                if (b < 7) {
#pragma unroll
                    for (int i = 0; i < 4; ++i)
#pragma unroll
                        for (int j = 0; j < 4; ++j) xb[(b + 1) & 1][i][j] = ld(512 * (4 * (b + 1) + i) + j * N);
                }
#pragma unroll
                for (int i = 0; i < 4; ++i) combine(4 * b + i, xb[b & 1][i]);
            }
        };
        if (interior && io.in_u8) {
            const unsigned short* p = reinterpret_cast<const unsigned short*>(io.in) + seg0 + tid;
            pipelined([&](int o) { const unsigned int w = p[o]; return decode_u8iq(w & 0xffu, w >> 8); });
        } else if (interior) {
            const float2* p = io.in + seg0 + tid;
            pipelined([&](int o) { return p[o]; });
        } else {
#pragma unroll
            for (int n1 = 0; n1 < 32; ++n1) {                   // edge blocks: bounds-checked, history / zero fill
                const long long g = seg0 + tid + 512 * n1;
                const float2 x[4] = {fetch(io, g), fetch(io, g + N), fetch(io, g + 2 * N), fetch(io, g + 3 * N)};
                combine(n1, x);
            }
        }
    }
    dit<32, +1>(v);
    float2 p[32];
    if constexpr (NC == 1) powers32(tw1[tid], p);
    else powers32b(tw1[tid], gc[tid], p);
    float2* s = sm + (tid >> 4) * ROW_PITCH + (tid & 15);
    before_store();
#pragma unroll
    for (int k1 = 0; k1 < 32; ++k1) s[k1 * PLANE_PITCH] = cmul(v[k1], p[k1]);
}

// Phase B is fftk::phase_mid_b unchanged.

// Column of plane k1 that holds the folded pair (row k2, n3' = 0/1) between the fold and the
// inverse: skewed by the plane index so that the 64 inverse tasks read conflict free.
RRC_HD int fold_col(int k1, int n3p) { return (2 * k1 + n3p) & 15; }

// Phase C (forward only) + spectrum multiply + 8:1 fold + DFT2 over b.  Results are written back
// into the thread's OWN two rows of its plane (no other thread reads those rows), so the only
// synchronisation needed before the inverse tasks read them is the CTA barrier that follows.
// Hp as fftk (Hp[(k1*32+k2)*16 + k3] = H[c + NC*(k1 + 32 k2 + 1024 k3)] / NBIG); Hres = rows k2 = l resident.
RRC_HD void phase_c_fold(int tid, const float2* Hp, const float2* Hres, float2* sm) {
    const int k1 = tid >> 4, l = tid & 15;
    const float4* hp1 = reinterpret_cast<const float4*>(Hp + (size_t)(k1 * 32 + l + 16) * 16);
    float4 h1[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) h1[i] = hp1[i];
    const float4* hres = reinterpret_cast<const float4*>(Hres + tid * HRES_PITCH);
#pragma unroll
    for (int half = 0; half < 2; ++half) {
        const int k2 = l + 16 * half;
        float2* row = sm + k1 * PLANE_PITCH + k2 * ROW_PITCH;
        float2 v[16];
#pragma unroll
        for (int n3 = 0; n3 < 16; ++n3) v[bitrev(n3, 4)] = row[n3];
        dit<16, +1>(v);                                         // v[k3], natural order
        float2 z0 = make_float2(0.f, 0.f), z1 = make_float2(0.f, 0.f);
#pragma unroll
        for (int k3 = 0; k3 < 16; k3 += 2) {
            const float4 h = half == 0 ? hres[k3 >> 1] : h1[k3 >> 1];
            z0 = cadd(z0, cmul(v[k3], make_float2(h.x, h.y)));          // b = 0
            z1 = cadd(z1, cmul(v[k3 + 1], make_float2(h.z, h.w)));      // b = 1
        }
        row[fold_col(k1, 0)] = cadd(z0, z1);                    // n3' = 0
        row[fold_col(k1, 1)] = csub(z0, z1);                    // n3' = 1
    }
}

// Inverse step 1 (tasks tau < 64: k1 = tau >> 1, n3' = tau & 1): conj W_64^{n3' k2}, IDFT32 over k2,
// conj W_2048^{k1 (2 n2' + n3')}; results returned in registers (v[n2']).
RRC_HD void inv1_load(int tau, const float2* sm, float2 (&v)[32]) {
    const int k1 = tau >> 1, n3p = tau & 1;
    const float2* col = sm + k1 * PLANE_PITCH + fold_col(k1, n3p);
#pragma unroll
    for (int k2 = 0; k2 < 32; ++k2) {
        const float2 x = col[k2 * ROW_PITCH];
        const float2 w = n3p ? make_float2((float)cos64(k2), (float)-sin64(k2)) : make_float2(1.f, 0.f);
        v[bitrev(k2, 5)] = cmul_conj(x, w);
    }
}
// P2 layout (inverse step 1 -> step 2): element (k1, t = 2 n2' + n3') at k1*P2_PITCH + t.
constexpr int P2_PITCH = 66;
constexpr int P2_ELEMS = 32 * P2_PITCH;
RRC_HD void inv1_compute_store(int tau, const float2* tw1, float2 (&v)[32], float2* sm) {
    const int k1 = tau >> 1, n3p = tau & 1;
    dit<32, -1>(v);                                             // v[n2'], natural order
    const float2 g2 = tw1[8 * k1];                              // W_2048^{k1}
    float2 p[32];
    powers32b(csqr(g2), n3p ? g2 : make_float2(1.f, 0.f), p);   // W_2048^{k1 (2 n2' + n3')}
    float2* dst = sm + k1 * P2_PITCH + n3p;
#pragma unroll
    for (int n2 = 0; n2 < 32; ++n2) dst[2 * n2] = cmul_conj(v[n2], p[n2]);
}
// Inverse step 2 (tasks t < 64): IDFT32 over k1 -> n1'; u[m2 = n1'*64 + t].
RRC_HD void inv2_load(int t, const float2* sm, float2 (&v)[32]) {
#pragma unroll
    for (int k1 = 0; k1 < 32; ++k1) v[bitrev(k1, 5)] = sm[k1 * P2_PITCH + t];
}
RRC_HD void inv2_compute_store(int t, float2 (&v)[32], float2* u) {
    dit<32, -1>(v);
#pragma unroll
    for (int n1 = 0; n1 < 32; ++n1) u[n1 * 64 + t] = v[n1];
}

// Combine + store.  CTA d of the cluster owns m2 in [d*LU/NC, (d+1)*LU/NC); thread tid handles
// m2 = d*LU/NC + tid + 512*i.  uc[c] = pointer to CTA c's u array (DSMEM on the device).
//   twm[m2] = W_{2048*NC}^{m2} (NC = 4 only; m2 < 2048).
template <int NC>
RRC_HD void combine_store(int tid, int d, long long blk, const FoldIO& io, const float2* const (&uc)[NC],
                          const float2* twm) {
    // output index of segment element n = 8*m:  j = jb + m
    const long long jb = (blk * (long long)io.V - io.T1eff) / FOLD_D - io.jbias;   // exact: V, T1eff multiples of 8
    const int m_first = io.T1eff / FOLD_D;
    constexpr int PER_CTA = LU / NC;
#pragma unroll
    for (int i = 0; i < PER_CTA / NT; ++i) {
        const int m2 = d * PER_CTA + tid + NT * i;
        if constexpr (NC == 1) {
            const long long j = jb + m2;
            if (m2 >= m_first && j >= 0 && j < io.n_out) fold_store(io, j, uc[0][m2]);
        } else {
            const float2 g = twm[m2];
            const float2 g2 = csqr(g);
            const float2 e0 = uc[0][m2];
            const float2 e1 = cmul_conj(uc[1][m2], g);
            const float2 e2 = cmul_conj(uc[2][m2], g2);
            const float2 e3 = cmul_conj(uc[3][m2], cmul(g2, g));
            const float2 a = cadd(e0, e2), b = csub(e0, e2), cc = cadd(e1, e3), dd = csub(e1, e3);
            float2 z[4];
            z[0] = cadd(a, cc);
            z[1] = make_float2(b.x - dd.y, b.y + dd.x);          // b + i dd
            z[2] = csub(a, cc);
            z[3] = make_float2(b.x + dd.y, b.y - dd.x);          // b - i dd
#pragma unroll
            for (int qq = 0; qq < 4; ++qq) {
                const int m = m2 + LU * qq;
                const long long j = jb + m;
                if (m >= m_first && j >= 0 && j < io.n_out) fold_store(io, j, z[qq]);
            }
        }
    }
}

}}  // namespace rrc::fftf
