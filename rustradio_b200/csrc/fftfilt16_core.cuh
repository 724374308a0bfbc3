// fftfilt16_core.cuh — 16384-point overlap-save block, 1024 threads x 16 points
// (N = 16*16*16*4).  Same transform, boundary and BlockIO as fftfilt_core.cuh (the
// 512-thread x 32-point variant); this variant trades one more shared-memory exchange
// per direction for 32 resident warps (<= 64 registers/thread), so eight warps per
// scheduler interleave their load / FP / store stretches instead of four running in
// lock step (profiles/r01_fftfilt_v5_ncu_summary.txt: FP pipe, shared-memory crossbar
// and global LSU were used almost serially).
//
//   n = n1*1024 + n2*64 + n3*4 + n4,      k = k1 + 16*k2 + 256*k3 + 4096*k4
//   A : tid = t = n mod 1024.  global -> DFT16 (n1) -> * W_N^{t*k1}                    == CTA barrier
//   B : tid = k1*64 + m (m = 4*n3+n4).  DFT16 (n2) -> * W_1024^{m*k2}                  -- plane barrier
//   C : tid = k1*64 + n4*16 + k2.  DFT16 (n3) -> * W_64^{n4*k3}                        -- plane barrier
//   D : tid = k1*64 + q*16 + k2.   k3 in 4q..4q+3: DFT4 (n4) -> * H -> IDFT4           -- plane barrier
//   C', B', A' : conjugate twiddle first, then the inverse DFT16, mirrored.
// A "plane" is one k1 (1024 points, 16 rows k2/n2 x 64 columns m) and is owned by two
// warps; planes are paired so eight 128-thread named barriers cover B<->C<->D.
// Exchange buffer: phys(k1, row, m) = k1*PLANE16 + row*ROW16 + m with ROW16 = 65 (odd): the
// row-contiguous accesses of A/B and the lanes-over-rows accesses of C/D (k2 is the fastest
// lane index there) are both bank-conflict free for 64-bit words, all offsets are immediates,
// and every phase rewrites exactly the words it read.
#pragma once
#include "fftfilt_core.cuh"

namespace rrc { namespace fftk16 {

using namespace rrc::fftr;
using rrc::fftk::BlockIO;
using rrc::fftk::N;

constexpr int NT16 = 1024;
constexpr int ROW16 = 65;
constexpr int PLANE16 = 16 * ROW16;                 // 1040
constexpr int SMEM16_ELEMS = 16 * PLANE16;          // 16640 float2
constexpr int HRES16_ELEMS = 8 * NT16;              // half of the spectrum, [e][tid]

// p[k] = w^k, k = 0..15.
RRC_HD void powers16(float2 w, float2 (&p)[16]) {
    float2 wp[4];
    wp[0] = w;
#pragma unroll
    for (int i = 1; i < 4; ++i) wp[i] = csqr(wp[i - 1]);
    p[0] = make_float2(1.f, 0.f);
#pragma unroll
    for (int k = 1; k < 16; ++k) {
        const int low = k & (-k);
        const int rest = k & (k - 1);
        const int b = low == 1 ? 0 : low == 2 ? 1 : low == 4 ? 2 : 3;
        p[k] = rest == 0 ? wp[b] : cmul(p[rest], wp[b]);
    }
}

// tw1[t] = W_N^t (t < 1024).
RRC_HD void phase_a(int tid, long long blk, const BlockIO& io, const float2* tw1, float2* sm) {
    float2 v[16];
    const long long seg0 = blk * (long long)io.V - io.T1 - io.shift;
    if (seg0 >= 0 && seg0 + N <= io.n_in) {
        const float2* p = io.in + seg0 + tid;
#pragma unroll
        for (int n1 = 0; n1 < 16; ++n1) v[bitrev(n1, 4)] = p[1024 * n1];
    } else {
        const long long g0 = seg0 + tid;
#pragma unroll
        for (int n1 = 0; n1 < 16; ++n1) {
            const long long g = g0 + 1024 * n1;
            float2 x = make_float2(0.f, 0.f);
            if (g < 0) { if (g + io.T1_total >= 0) x = io.hist[g + io.T1_total]; }
            else if (g < io.n_in) x = io.in[g];
            v[bitrev(n1, 4)] = x;
        }
    }
    dit<16, +1>(v);
    float2 p[16];
    powers16(tw1[tid], p);
    float2* s = sm + (tid >> 6) * ROW16 + (tid & 63);           // (k1 = 0, row n2, col m)
#pragma unroll
    for (int k1 = 0; k1 < 16; ++k1) s[k1 * PLANE16] = cmul(v[k1], p[k1]);
}

// tw2[k2*64 + m] = W_1024^{m*k2}.
RRC_HD void phase_b(int tid, const float2* tw2, float2* sm) {
    const int k1 = tid >> 6, m = tid & 63;
    float2* col = sm + k1 * PLANE16 + m;
    const float2* tw = tw2 + m;
    float2 v[16];
#pragma unroll
    for (int n2 = 0; n2 < 16; ++n2) v[bitrev(n2, 4)] = col[n2 * ROW16];
    dit<16, +1>(v);
#pragma unroll
    for (int k2 = 0; k2 < 16; ++k2) col[k2 * ROW16] = cmul(v[k2], tw[k2 * 64]);
}
RRC_HD void phase_bi(int tid, const float2* tw2, float2* sm) {
    const int k1 = tid >> 6, m = tid & 63;
    float2* col = sm + k1 * PLANE16 + m;
    const float2* tw = tw2 + m;
    float2 v[16];
#pragma unroll
    for (int k2 = 0; k2 < 16; ++k2) v[bitrev(k2, 4)] = cmul_conj(col[k2 * ROW16], tw[k2 * 64]);
    dit<16, -1>(v);
#pragma unroll
    for (int n2 = 0; n2 < 16; ++n2) col[n2 * ROW16] = v[n2];
}

// tw3[k3*4 + n4] = W_64^{n4*k3}.  tid = k1*64 + n4*16 + k2 (k2 fastest across lanes).
RRC_HD void phase_c(int tid, const float2* tw3, float2* sm) {
    const int k1 = tid >> 6, n4 = (tid >> 4) & 3, k2 = tid & 15;
    float2* e = sm + k1 * PLANE16 + k2 * ROW16 + n4;            // element m = 4*n3 + n4
    const float2* tw = tw3 + n4;
    float2 v[16];
#pragma unroll
    for (int n3 = 0; n3 < 16; ++n3) v[bitrev(n3, 4)] = e[4 * n3];
    dit<16, +1>(v);
#pragma unroll
    for (int k3 = 0; k3 < 16; ++k3) e[4 * k3] = cmul(v[k3], tw[k3 * 4]);
}
RRC_HD void phase_ci(int tid, const float2* tw3, float2* sm) {
    const int k1 = tid >> 6, n4 = (tid >> 4) & 3, k2 = tid & 15;
    float2* e = sm + k1 * PLANE16 + k2 * ROW16 + n4;
    const float2* tw = tw3 + n4;
    float2 v[16];
#pragma unroll
    for (int k3 = 0; k3 < 16; ++k3) v[bitrev(k3, 4)] = cmul_conj(e[4 * k3], tw[k3 * 4]);
    dit<16, -1>(v);
#pragma unroll
    for (int n3 = 0; n3 < 16; ++n3) e[4 * n3] = v[n3];
}

// Spectrum layout for phase D: thread tid = k1*64 + q*16 + k2 multiplies the 16 values
// e = (k3 - 4q)*4 + k4 of  Hd[tid*16 + e] = H[k1 + 16*k2 + 256*k3 + 4096*k4] / N.
// e < 8 is kept resident in shared memory as Hres[e*1024 + tid]; e >= 8 comes from L2.
RRC_HD void load_hres(int tid, const float2* Hd, float2* Hres) {
    const float4* src = reinterpret_cast<const float4*>(Hd + (size_t)tid * 16);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float4 h = src[i];
        Hres[(2 * i) * NT16 + tid] = make_float2(h.x, h.y);
        Hres[(2 * i + 1) * NT16 + tid] = make_float2(h.z, h.w);
    }
}
RRC_HD void phase_d(int tid, const float2* Hd, const float2* Hres, float2* sm) {
    const int k1 = tid >> 6, q = (tid >> 4) & 3, k2 = tid & 15;
    float2* e = sm + k1 * PLANE16 + k2 * ROW16 + 16 * q;        // m = 4*k3 + n4, k3 = 4q + j
    const float4* hg = reinterpret_cast<const float4*>(Hd + (size_t)tid * 16 + 8);
    float4 h1[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) h1[i] = hg[i];
    const float2* hr = Hres + tid;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        float2 v[4];
#pragma unroll
        for (int n4 = 0; n4 < 4; ++n4) v[bitrev(n4, 2)] = e[4 * j + n4];
        dit<4, +1>(v);                                          // v[k4]
        float2 u[4];
#pragma unroll
        for (int k4 = 0; k4 < 4; ++k4) {
            const int idx = 4 * j + k4;
            float2 h;
            if (idx < 8) h = hr[idx * NT16];
            else {
                const float4 hh = h1[(idx - 8) >> 1];
                h = (idx & 1) ? make_float2(hh.z, hh.w) : make_float2(hh.x, hh.y);
            }
            u[bitrev(k4, 2)] = cmul(v[k4], h);
        }
        dit<4, -1>(u);                                          // u[n4]
#pragma unroll
        for (int n4 = 0; n4 < 4; ++n4) e[4 * j + n4] = u[n4];
    }
}

template <bool DECIM, bool ACCUM>
RRC_HD void phase_ai(int tid, long long blk, const BlockIO& io, const float2* tw1, const float2* sm) {
    float2 p[16];
    powers16(tw1[tid], p);
    const float2* s = sm + (tid >> 6) * ROW16 + (tid & 63);
    float2 v[16];
#pragma unroll
    for (int k1 = 0; k1 < 16; ++k1) v[bitrev(k1, 4)] = cmul_conj(s[k1 * PLANE16], p[k1]);
    dit<16, -1>(v);
    // v[n1] is segment element n = tid + 1024*n1; n >= T1 are valid outputs, index o = o0 + n.
    const long long o0 = blk * (long long)io.V - io.T1;
    const int tq = io.T1 >> 10, tr = io.T1 & 1023;
    if constexpr (!DECIM) {
        float2* q = io.out + o0 + tid;
        if (o0 + N <= io.n_out) {
#pragma unroll
            for (int n1 = 0; n1 < 16; ++n1)
                if (n1 > tq || (n1 == tq && tid >= tr)) q[1024 * n1] = ACCUM ? cadd(q[1024 * n1], v[n1]) : v[n1];
        } else {
#pragma unroll
            for (int n1 = 0; n1 < 16; ++n1)
                if ((n1 > tq || (n1 == tq && tid >= tr)) && o0 + tid + 1024 * n1 < io.n_out)
                    q[1024 * n1] = ACCUM ? cadd(q[1024 * n1], v[n1]) : v[n1];
        }
    } else {
        const long long D = io.deci;
        long long r = o0 + tid - io.skip;
        long long qd = r >= 0 ? r / D : -((-r + D - 1) / D);
        long long m = r - qd * D;
        const long long sq = 1024 / D, sr = 1024 % D;
#pragma unroll
        for (int n1 = 0; n1 < 16; ++n1) {
            if ((n1 > tq || (n1 == tq && tid >= tr)) && m == 0 && qd >= 0 && qd < io.n_out)
                io.out[qd] = ACCUM ? cadd(io.out[qd], v[n1]) : v[n1];
            qd += sq; m += sr;
            if (m >= D) { m -= D; ++qd; }
        }
    }
}

}}  // namespace rrc::fftk16
