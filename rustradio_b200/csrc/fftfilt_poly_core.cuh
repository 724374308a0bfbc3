// fftfilt_poly_core.cuh — FftFilter fused with decimate-by-D (RationalResampler(1, D)) as a POLYPHASE
// overlap-save filter.  Only y[skip + D m] is wanted, and with j = D q + p
//
//     y[skip + D m] = sum_j h[j] x[skip + D m - j] = sum_{p} sum_q h_p[q] u_p[m - q],
//     h_p[q] = h[D q + p]  (about ntaps / D taps),   u_p[k] = x[D k + skip - p]     (D consecutive values of p),
//
// i.e. D filters of ntaps/D taps, each on a D-times slower stream, summed.  In the frequency domain the
// sum moves in front of the inverse transform:  Y = sum_p H_p X_p, so one output block costs D forward
// 16384-point transforms and ONE inverse (the plain kernel with a store predicate pays D forward and D
// inverse for the same outputs, the folded kernel of fftfilt_fold_core.cuh one 4 x 16384 forward per
// 49152 input samples where this needs 8 x 16384 per 114688).  BASELINE config 5 (16385 taps, D = 8):
// 2049-tap phases, V = 14336 kept outputs = 114688 input samples per block, no cluster.
//
// The running sum over p (16384 complex bins = 128 KiB) does not fit beside the 136 KiB exchange buffer
// in shared memory: it lives in TENSOR MEMORY — thread tid owns the 32 bins its phase C produces
// (rows k2 = l, l + 16 of plane k1), 64 consecutive 32-bit columns of its own TMEM lane, read and written
// with tcgen05.ld / tcgen05.st.32x32b (the Acc policy below; the CPU emulator uses a plain array).
//
// Replaces Engine::run + sum_vec (rustradio src/fft_filter.rs:172-176,281-287) followed by
// RationalResampler::work with interp = 1 (src/rational_resampler.rs:155-206).
// All functions are __host__ __device__ so tests/emul runs the same index math on the CPU.
#pragma once
#include "fftfilt_core.cuh"

namespace rrc { namespace fftp {

using namespace rrc::fftk;   // N (16384), NT (512), phys(), powers32, BlockIO, store_outputs ...

constexpr int POLY_MAX_D = 16;

// Branches are numbered by the RESIDUE r of their samples: with s' = skip mod D and p = s' - r (r = 0..D-1, p may be
// negative: the branch then starts with a zero tap) all D branch streams are aligned on the same k,
//     u_r[k] = x[D k + (skip - s') + r],     h_r[q] = h[D q + s' - r]  (0 outside [0, ntaps)),
// so the two branches r, r + 1 (r even) of one k sit in the same aligned 16 bytes: ONE 128-bit load feeds two branches.
struct PolyIO {
    // b.in / b.hist / b.n_in / b.T1_total / b.in_u8 / b.hist_next describe the INPUT stream (update_history);
    // b.out / b.n_out / b.V / b.T1 (= longest branch - 1) / b.epi the decimated OUTPUT stream (store_outputs);
    // b.deci = 1, b.skip = 0, b.shift = 0, b.real = 0.
    BlockIO b;
    int D;                   // decimation = number of polyphase branches
    long long sbase;         // skip - skip mod D: first kept filter output, rounded down to a multiple of D
};

// Longest branch - 1 for decimation phase smod = skip mod D.
RRC_HD int poly_T1(long long ntaps, int D, int smod) { return (int)((ntaps - 1 - smod + D - 1) / D); }

// Input sample index of segment element n of branch r in block blk.
RRC_HD long long poly_g(const PolyIO& io, long long blk, int r, long long n) {
    return (long long)io.D * (blk * (long long)io.b.V - io.b.T1 + n) + io.sbase + r;
}
// CTA-uniform: every element of the segment is a sample of this call's input.
RRC_HD bool poly_interior(const PolyIO& io, long long blk, int r) {
    return poly_g(io, blk, r, 0) >= 0 && poly_g(io, blk, r, N - 1) < io.b.n_in;
}
// CTA-uniform: branches r and r + 1 can be fetched with aligned 128-bit loads.
RRC_HD bool poly_pair_ok(const PolyIO& io, long long blk, int r) {
    return !io.b.in_u8 && (r & 1) == 0 && (io.D & 1) == 0 && poly_interior(io, blk, r) && poly_interior(io, blk, r + 1) &&
           ((reinterpret_cast<unsigned long long>(io.b.in) + 8ull * (unsigned long long)poly_g(io, blk, r, 0)) & 15ull) == 0;
}

// Sample g of the stream with its carried history (g < 0) and zero fill on both sides.
RRC_HD float2 poly_fetch(const PolyIO& io, long long g) {
    if (g < 0) return g + io.b.T1_total >= 0 ? io.b.hist[g + io.b.T1_total] : make_float2(0.f, 0.f);
    return g < io.b.n_in ? ld_iq(io.b.in, g, io.b.in_u8) : make_float2(0.f, 0.f);
}

// The 32 segment elements n = tid + 512 n1 of branch r (stride D samples in memory), bit-reversed
// for the DIT transform of phase A.
RRC_HD void poly_load(int tid, long long blk, int r, const PolyIO& io, float2 (&v)[32]) {
    const long long g0 = poly_g(io, blk, r, tid);
    const long long step = 512ll * io.D;
    if (poly_interior(io, blk, r)) {
        if (io.b.in_u8) {
            const unsigned short* q = reinterpret_cast<const unsigned short*>(io.b.in) + g0;
            unsigned int w[32];
#pragma unroll
            for (int n1 = 0; n1 < 32; ++n1) w[n1] = q[step * n1];
#pragma unroll
            for (int n1 = 0; n1 < 32; ++n1) v[bitrev(n1, 5)] = decode_u8iq(w[n1] & 0xffu, w[n1] >> 8);
        } else {
            const float2* q = io.b.in + g0;
#pragma unroll
            for (int n1 = 0; n1 < 32; ++n1) v[bitrev(n1, 5)] = q[step * n1];
        }
    } else {
#pragma unroll
        for (int n1 = 0; n1 < 32; ++n1) v[bitrev(n1, 5)] = poly_fetch(io, g0 + step * n1);
    }
}

// Branches r (-> v) and r + 1 (-> the stash, in the order phase A wants them) with 32 aligned 128-bit loads per thread,
// four batches of 8 so that at most 32 load registers are in flight beside v (batches of 16 spilled).  Requires poly_pair_ok().
template <class Stash>
RRC_HD void poly_load_pair(int tid, long long blk, int r, const PolyIO& io, float2 (&v)[32], const Stash& stash) {
    const float2* q = io.b.in + poly_g(io, blk, r, tid);
    const long long step = 512ll * io.D;
#pragma unroll
    for (int b = 0; b < 4; ++b) {
        float4 w[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) w[e] = *reinterpret_cast<const float4*>(q + step * bitrev(8 * b + e, 5));
        float2 x[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) { v[8 * b + e] = make_float2(w[e].x, w[e].y); x[e] = make_float2(w[e].z, w[e].w); }
        stash.store8(tid, b, x);
    }
}
template <class Stash>
RRC_HD void poly_load_stash(int tid, float2 (&v)[32], const Stash& stash) {
    stash.load(tid, 0, *reinterpret_cast<float2(*)[16]>(&v[0]));
    stash.load(tid, 1, *reinterpret_cast<float2(*)[16]>(&v[16]));
}

// CPU-side accumulator / stash (the emulator): arrays indexed like the TMEM columns of thread tid.
struct HostAcc {
    float2* a;               // [512][32]
    void load(int tid, int half, float2 (&x)[16]) const { for (int i = 0; i < 16; ++i) x[i] = a[tid * 32 + half * 16 + i]; }
    void store(int tid, int half, const float2 (&x)[16]) const { for (int i = 0; i < 16; ++i) a[tid * 32 + half * 16 + i] = x[i]; }
    void store8(int tid, int quarter, const float2 (&x)[8]) const { for (int i = 0; i < 8; ++i) a[tid * 32 + quarter * 8 + i] = x[i]; }
};

// One row (k2 = l + 16 half) of phase C of one branch: DFT16 over n3 -> x H_r -> running sum over the branches in u.
// h = the 16 spectrum values of this row as 8 float4.
template <class Acc>
RRC_HD void poly_c_row(int tid, int half, const float4* h, const float2* row, const Acc& acc, bool first, float2 (&u)[16]) {
    float2 v[16];
#pragma unroll
    for (int n3 = 0; n3 < 16; ++n3) v[bitrev(n3, 4)] = row[n3];
    dit<16, +1>(v);                                             // v[k3], natural order
    if (!first) acc.load(tid, half, u);
#pragma unroll
    for (int k3 = 0; k3 < 16; k3 += 2) {
        const float4 hh = h[k3 >> 1];
        const float2 a = cmul(v[k3], make_float2(hh.x, hh.y));
        const float2 c = cmul(v[k3 + 1], make_float2(hh.z, hh.w));
        u[k3] = first ? a : cadd(u[k3], a);
        u[k3 + 1] = first ? c : cadd(u[k3 + 1], c);
    }
}

// Phase C of one branch.  The first branch writes the running sum, the others add to it; the LAST one (single-CTA
// geometry) carries on with the inverse DFT16 and puts the rows back into the exchange buffer for B' / A'
// (fftfilt_core.cuh).  Spectrum of THIS branch: Hp[(k1*32 + k2)*16 + k3] = H_r[k1 + 32 k2 + 1024 k3] / N in global
// memory; the rows k2 = l of it also in shared memory (Hres, the padded layout of fftk::load_hres — on the GPU a bulk
// copy per branch puts them there while phases A and B run): the row each thread multiplies first comes from shared
// memory, the second one from L2 into registers at the top of the phase, as in fftk::phase_mid_c.
// HRES = false: both rows from L2, each at the top of its half (no shared-memory rows, fewer live registers).
template <bool HRES = true, class Acc>
RRC_HD void phase_c_acc(int tid, const float2* Hp, const float2* Hres, float2* sm, const Acc& acc, bool first, bool last) {
    const int k1 = tid >> 4, l = tid & 15;
    const float4* hp1 = reinterpret_cast<const float4*>(Hp + (size_t)(k1 * 32 + l + 16) * 16);
    float4 h1[8];
    if constexpr (HRES) {
#pragma unroll
        for (int i = 0; i < 8; ++i) h1[i] = hp1[i];
    }
    const float4* hres = reinterpret_cast<const float4*>(Hres + tid * HRES_PITCH);
#pragma unroll
    for (int half = 0; half < 2; ++half) {
        float2* row = sm + k1 * PLANE_PITCH + (l + 16 * half) * ROW_PITCH;   // (k1, r = k2, c = 0)
        float2 u[16];
        if constexpr (!HRES) {
            const float4* hp = reinterpret_cast<const float4*>(Hp + (size_t)(k1 * 32 + l + 16 * half) * 16);
#pragma unroll
            for (int i = 0; i < 8; ++i) h1[i] = hp[i];
        }
        poly_c_row(tid, half, (HRES && half == 0) ? hres : h1, row, acc, first, u);
        if (!last) {
            acc.store(tid, half, u);
        } else {
            float2 v[16];
#pragma unroll
            for (int k3 = 0; k3 < 16; ++k3) v[bitrev(k3, 4)] = u[k3];
            dit<16, -1>(v);                                     // v[n3], natural order
#pragma unroll
            for (int n3 = 0; n3 < 16; ++n3) row[n3] = v[n3];
        }
    }
}

}}  // namespace rrc::fftp
