// fftfilt_core.cuh — one 16384-point overlap-save block of the FftFilter
// kernel, written as three barrier-separated phases that are pure functions of
// (thread id, shared memory, parameters).  __host__ __device__ so that the
// same code runs under the CPU emulator in tests/emul (index-math check
// without a GPU) and inside the CUDA kernel.
//
// Replaces Engine::run + sum_vec (rustradio src/fft_filter.rs:172-176,281-287):
// IFFT(FFT(x) * H), with 1/N folded into H like :153-162.
//
// Transform: N = 16384 = N1*N2*N3 = 32*32*16, 512 threads x 32 points.
//   n = n1*512 + n2*16 + n3,  k = k1 + 32*k2 + 1024*k3
//   A  : global -> DFT32 over n1 -> * W_N^{t*k1}            (t = n2*16+n3 = tid)
//   MID: tid = k1*16 + l (a half-warp owns one k1 plane of 32x16 points)
//        B : DFT32 over n2 -> * W_512^{n3*k2}      (column n3 = l)
//        C : DFT16 over n3 -> * H[k] -> IDFT16     (rows k2 = l, l+16)
//        B': * conj W_512^{n3*k2} -> IDFT32 over k2 (column n3 = l)
//        B->C->B' exchange data only inside the half-warp: __syncwarp(), no CTA barrier.
//   A' : * conj W_N^{t*k1} -> IDFT32 over k1 -> global      (tid = t)
// Shared-memory exchange buffer: plane k1 (32 rows x 16 columns of float2) at
// k1*544, row pitch 17 (one pad element per row):
//     phys(k1, r, c) = k1*544 + r*17 + c
// Column access (fixed c, lanes = rows, stride 17) and row access (fixed r,
// lanes = columns) are both bank-conflict free for 64-bit words, and every
// per-register offset is a compile-time immediate.  Every phase writes back
// exactly the words it read (in place), so there are only two CTA barriers
// per block: A -> MID and MID -> A'.
#pragma once
#include "fft_regs.cuh"

namespace rrc { namespace fftk {

using namespace rrc::fftr;

constexpr int N = 16384;
constexpr int NT = 512;          // threads per CTA
constexpr int ROW_PITCH = 17;
constexpr int PLANE_PITCH = 32 * ROW_PITCH;        // 544
constexpr int SMEM_ELEMS = 32 * PLANE_PITCH;       // 17408 float2

struct BlockIO {
    const float2* in;        // this call's input samples, x[0..n_in)
    const float2* hist;      // previous T1 samples, hist[i] = x[i - T1]
    float2* out;             // y[0..n_out)
    long long n_in;          // valid input samples
    long long n_out;         // outputs to write
    int T1;                  // taps of THIS partition - 1
    int V;                   // valid outputs per block = N - T1
    int T1_total;            // ntaps - 1 of the whole filter = length of hist
    long long shift;         // input delay of this tap partition (p * partition length); 0 for a single partition
    int deci;                // output decimation (fused RationalResampler(1,deci)); 1 = none
    long long skip;          // first kept filter output index (decimation phase)
};

RRC_HD int phys(int k1, int r, int c) { return k1 * PLANE_PITCH + r * ROW_PITCH + c; }

#if defined(__CUDA_ARCH__)
#define RRC_SYNCWARP() __syncwarp()
#else
#define RRC_SYNCWARP() ((void)0)
#endif

// Powers p[k] = w^k, k = 0..31, depth <= 5 multiplications each.
RRC_HD void powers32(float2 w, float2 (&p)[32]) {
    float2 wp[5];
    wp[0] = w;
#pragma unroll
    for (int i = 1; i < 5; ++i) wp[i] = csqr(wp[i - 1]);
    p[0] = make_float2(1.f, 0.f);
#pragma unroll
    for (int k = 1; k < 32; ++k) {
        const int low = k & (-k);
        const int rest = k & (k - 1);
        const int b = low == 1 ? 0 : low == 2 ? 1 : low == 4 ? 2 : low == 8 ? 3 : 4;
        p[k] = rest == 0 ? wp[b] : cmul(p[rest], wp[b]);
    }
}

// Phase A: load segment of block `blk`, DFT32 over n1, twiddle, write smem.
// tw1[t] = W_N^t = exp(-2 pi i t / N), t < 512.
RRC_HD void phase_a(int tid, long long blk, const BlockIO& io, const float2* tw1, float2* sm) {
    float2 v[32];
    // input index of segment element 0 (tap partition p filters the input delayed by `shift`)
    const long long seg0 = blk * (long long)io.V - io.T1 - io.shift;
    if (seg0 >= 0 && seg0 + N <= io.n_in) {                     // interior block: no bounds checks
        const float2* p = io.in + seg0 + tid;
#pragma unroll
        for (int n1 = 0; n1 < 32; ++n1) v[bitrev(n1, 5)] = p[512 * n1];
    } else {
        const long long g0 = seg0 + tid;
#pragma unroll
        for (int n1 = 0; n1 < 32; ++n1) {
            const long long g = g0 + 512 * n1;
            float2 x = make_float2(0.f, 0.f);
            if (g < 0) { if (g + io.T1_total >= 0) x = io.hist[g + io.T1_total]; }
            else if (g < io.n_in) x = io.in[g];
            v[bitrev(n1, 5)] = x;
        }
    }
    dit<32, +1>(v);                                             // bit-reversed in, natural k1 out
    float2 p[32];
    powers32(tw1[tid], p);
    float2* s = sm + (tid >> 4) * ROW_PITCH + (tid & 15);       // (k1 = 0, r = n2, c = n3)
#pragma unroll
    for (int k1 = 0; k1 < 32; ++k1) s[k1 * PLANE_PITCH] = cmul(v[k1], p[k1]);
}

// Phase MID: B, C, B' on one k1 plane per half-warp.  tw2[k2*16 + n3] = W_512^{n3*k2}.
// Hp[(k1*32 + k2)*16 + k3] = H[k1 + 32*k2 + 1024*k3] / N.
RRC_HD void phase_mid_b(int tid, const float2* tw2, float2* sm) {
    const int k1 = tid >> 4, l = tid & 15;
    float2* col = sm + k1 * PLANE_PITCH + l;                    // (k1, r = 0, c = l)
    const float2* tw = tw2 + l;
    float2 v[32];
#pragma unroll
    for (int n2 = 0; n2 < 32; ++n2) v[bitrev(n2, 5)] = col[n2 * ROW_PITCH];
    dit<32, +1>(v);
#pragma unroll
    for (int k2 = 0; k2 < 32; ++k2) col[k2 * ROW_PITCH] = cmul(v[k2], tw[k2 * 16]);
}
// Spectrum residency: the row each thread multiplies first (k2 = l) lives in shared memory for
// the whole kernel (Hres, 512 rows, pitch HRES_PITCH so the per-thread 128-bit reads are
// conflict free); the second row (k2 = l + 16) is fetched from L2 into registers at the top of
// phase C and consumed ~400 instructions later.  (The whole spectrum, 128 KiB, does not fit
// beside the 136 KiB exchange buffer.)
constexpr int HRES_PITCH = 18;                                  // float2 per resident row (144 B)
constexpr int HRES_ELEMS = NT * HRES_PITCH;

RRC_HD void load_hres(int tid, const float2* Hp, float2* Hres) {
    const int k1 = tid >> 4, l = tid & 15;
    const float4* src = reinterpret_cast<const float4*>(Hp + (size_t)(k1 * 32 + l) * 16);
    float4* dst = reinterpret_cast<float4*>(Hres + tid * HRES_PITCH);
#pragma unroll
    for (int i = 0; i < 8; ++i) dst[i] = src[i];
}

RRC_HD void phase_mid_c(int tid, const float2* Hp, const float2* Hres, float2* sm) {
    const int k1 = tid >> 4, l = tid & 15;
    const float4* hp1 = reinterpret_cast<const float4*>(Hp + (size_t)(k1 * 32 + l + 16) * 16);
    float4 h1[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) h1[i] = hp1[i];
    const float4* hres = reinterpret_cast<const float4*>(Hres + tid * HRES_PITCH);
#pragma unroll
    for (int half = 0; half < 2; ++half) {
        const int k2 = l + 16 * half;
        float2* row = sm + k1 * PLANE_PITCH + k2 * ROW_PITCH;   // (k1, r = k2, c = 0)
        float2 v[16];
#pragma unroll
        for (int n3 = 0; n3 < 16; ++n3) v[bitrev(n3, 4)] = row[n3];
        dit<16, +1>(v);                                         // v[k3], natural order
        float2 u[16];
#pragma unroll
        for (int k3 = 0; k3 < 16; k3 += 2) {
            const float4 h = half == 0 ? hres[k3 >> 1] : h1[k3 >> 1];
            u[bitrev(k3, 4)] = cmul(v[k3], make_float2(h.x, h.y));
            u[bitrev(k3 + 1, 4)] = cmul(v[k3 + 1], make_float2(h.z, h.w));
        }
        dit<16, -1>(u);                                         // u[n3], natural order
#pragma unroll
        for (int n3 = 0; n3 < 16; ++n3) row[n3] = u[n3];
    }
}
RRC_HD void phase_mid_bi(int tid, const float2* tw2, float2* sm) {
    const int k1 = tid >> 4, l = tid & 15;
    float2* col = sm + k1 * PLANE_PITCH + l;
    const float2* tw = tw2 + l;
    float2 v[32];
#pragma unroll
    for (int k2 = 0; k2 < 32; ++k2) v[bitrev(k2, 5)] = cmul_conj(col[k2 * ROW_PITCH], tw[k2 * 16]);
    dit<32, -1>(v);
#pragma unroll
    for (int n2 = 0; n2 < 32; ++n2) col[n2 * ROW_PITCH] = v[n2];
}
RRC_HD void phase_mid(int tid, const float2* tw2, const float2* Hp, const float2* Hres, float2* sm) {
    phase_mid_b(tid, tw2, sm);
    RRC_SYNCWARP();
    phase_mid_c(tid, Hp, Hres, sm);
    RRC_SYNCWARP();
    phase_mid_bi(tid, tw2, sm);
}

// Phase A': tid = t; conj twiddle, IDFT32 over k1 -> n1; store valid outputs.
template <bool DECIM, bool ACCUM>
RRC_HD void phase_ai(int tid, long long blk, const BlockIO& io, const float2* tw1, const float2* sm) {
    float2 p[32];
    powers32(tw1[tid], p);
    const float2* s = sm + (tid >> 4) * ROW_PITCH + (tid & 15);
    float2 v[32];
#pragma unroll
    for (int k1 = 0; k1 < 32; ++k1) v[bitrev(k1, 5)] = cmul_conj(s[k1 * PLANE_PITCH], p[k1]);
    dit<32, -1>(v);
    // v[n1] is segment element n = tid + 512*n1; elements n >= T1 are valid
    // outputs, filter output index o = o0 + n.
    const long long o0 = blk * (long long)io.V - io.T1;
    const int tq = io.T1 >> 9, tr = io.T1 & 511;
    if constexpr (!DECIM) {
        float2* q = io.out + o0 + tid;
        if (o0 + N <= io.n_out) {                               // interior: only the n >= T1 test
#pragma unroll
            for (int n1 = 0; n1 < 32; ++n1)
                if (n1 > tq || (n1 == tq && tid >= tr)) q[512 * n1] = ACCUM ? cadd(q[512 * n1], v[n1]) : v[n1];
        } else {
#pragma unroll
            for (int n1 = 0; n1 < 32; ++n1)
                if ((n1 > tq || (n1 == tq && tid >= tr)) && o0 + tid + 512 * n1 < io.n_out)
                    q[512 * n1] = ACCUM ? cadd(q[512 * n1], v[n1]) : v[n1];
        }
    } else {
        // keep outputs with (o - skip) >= 0 and (o - skip) % deci == 0, at index (o - skip)/deci.
        const long long D = io.deci;
        long long r = o0 + tid - io.skip;                       // for n1 = 0
        long long qd = r >= 0 ? r / D : -((-r + D - 1) / D);    // floor division
        long long m = r - qd * D;                               // in [0, D)
        const long long sq = 512 / D, sm_ = 512 % D;
#pragma unroll
        for (int n1 = 0; n1 < 32; ++n1) {
            if ((n1 > tq || (n1 == tq && tid >= tr)) && m == 0 && qd >= 0 && qd < io.n_out)
                io.out[qd] = ACCUM ? cadd(io.out[qd], v[n1]) : v[n1];
            qd += sq; m += sm_;
            if (m >= D) { m -= D; ++qd; }
        }
    }
}

}}  // namespace rrc::fftk
