// fftfilt_core.cuh — one 16384-point overlap-save block of the FftFilter
// kernel, written as five barrier-separated phases that are pure functions of
// (thread id, shared memory, parameters).  __host__ __device__ so that the
// same code runs under the CPU emulator in tests/emul (index-math check
// without a GPU) and inside the CUDA kernel.
//
// Replaces Engine::run + sum_vec (rustradio src/fft_filter.rs:172-176,281-287):
// IFFT(FFT(x) * H), with 1/N folded into H like :153-162.
//
// Transform: N = 16384 = N1*N2*N3 = 32*32*16, 512 threads x 32 points.
//   n = n1*512 + n2*16 + n3,  k = k1 + 32*k2 + 1024*k3
//   A : global -> DFT32 over n1 -> * W_N^{t*k1}          (t = n2*16+n3 = tid)
//   B : DFT32 over n2 -> * W_512^{n3*k2}                  (tid = k1*16+n3)
//   C : DFT16 over n3 -> * H[k] -> IDFT16 over k3         (rows P = k1*32+k2: tid, tid+512)
//   B': * conj W_512^{n3*k2} -> IDFT32 over k2            (tid = k1*16+n3)
//   A': * conj W_N^{t*k1} -> IDFT32 over k1 -> global     (tid = t)
// Shared-memory exchange buffer: 16384 float2 (128 KiB), one layout for all
// four exchanges: phys(k1|., row r in [0,32), col c in [0,16)) =
//   k1*512 + r*16 + (c ^ (r & 15)).
// Every phase writes back exactly the set of locations it read (in-place), so
// one __syncthreads() per exchange suffices; the XOR swizzle makes both the
// row-wise (B, B') and column-wise (A, A', C) accesses bank-conflict free.
#pragma once
#include "fft_regs.cuh"

namespace rrc { namespace fftk {

using namespace rrc::fftr;

constexpr int N = 16384;
constexpr int NT = 512;          // threads per CTA
constexpr int N1 = 32, N2 = 32, N3 = 16;

struct BlockIO {
    const float2* in;        // this call's input samples, x[0..n_in)
    const float2* hist;      // previous T1 samples, hist[i] = x[i - T1]
    float2* out;             // y[0..n_out)
    long long n_in;          // valid input samples
    long long n_out;         // outputs to write
    int T1;                  // ntaps - 1
    int V;                   // valid outputs per block = N - T1
    int deci;                // output decimation (fused RationalResampler(1,deci)); 1 = none
    long long skip;          // first kept filter output index (decimation phase)
};

RRC_HD int phys(int k1, int r, int c) { return k1 * 512 + r * 16 + (c ^ (r & 15)); }

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
    const long long g0 = blk * (long long)io.V - io.T1 + tid;
#pragma unroll
    for (int n1 = 0; n1 < 32; ++n1) {
        const long long g = g0 + 512 * n1;
        float2 x = make_float2(0.f, 0.f);
        if (g < 0) { if (g + io.T1 >= 0) x = io.hist[g + io.T1]; }
        else if (g < io.n_in) x = io.in[g];
        v[n1] = x;
    }
    dif<32, +1>(v);
    float2 p[32];
    powers32(tw1[tid], p);
    const int n2 = tid >> 4, n3 = tid & 15;
#pragma unroll
    for (int j = 0; j < 32; ++j) {
        const int k1 = bitrev(j, 5);
        sm[phys(k1, n2, n3)] = cmul(v[j], p[k1]);
    }
}

// Phase B: tid = k1*16 + n3; DFT32 over n2; twiddle tw2[k2*16+n3] = W_512^{n3*k2}.
RRC_HD void phase_b(int tid, const float2* tw2, float2* sm) {
    const int k1 = tid >> 4, n3 = tid & 15;
    float2 v[32];
#pragma unroll
    for (int n2 = 0; n2 < 32; ++n2) v[n2] = sm[phys(k1, n2, n3)];
    dif<32, +1>(v);
#pragma unroll
    for (int j = 0; j < 32; ++j) {
        const int k2 = bitrev(j, 5);
        sm[phys(k1, k2, n3)] = cmul(v[j], tw2[k2 * 16 + n3]);
    }
}

// Phase C: rows P = tid and tid + 512 (P = k1*32 + k2); DFT16 over n3,
// multiply by the pre-permuted spectrum Hp[P*16 + j] (j = register position,
// k3 = bitrev4(j)), inverse DFT16 over k3, write back.
RRC_HD void phase_c(int tid, const float2* Hp, float2* sm) {
#pragma unroll
    for (int half = 0; half < 2; ++half) {
        const int P = tid + 512 * half;
        const int k1 = P >> 5, k2 = P & 31;
        float2 v[16];
#pragma unroll
        for (int n3 = 0; n3 < 16; ++n3) v[n3] = sm[phys(k1, k2, n3)];
        dif<16, +1>(v);
        float2 u[16];
        const float4* hp4 = reinterpret_cast<const float4*>(Hp + (size_t)P * 16);
#pragma unroll
        for (int j = 0; j < 16; j += 2) {
            const float4 h = hp4[j >> 1];
            // v[j] holds k3 = bitrev4(j); feed the inverse DIF in natural k3 order.
            u[bitrev(j, 4)] = cmul(v[j], make_float2(h.x, h.y));
            u[bitrev(j + 1, 4)] = cmul(v[j + 1], make_float2(h.z, h.w));
        }
        dif<16, -1>(u);
#pragma unroll
        for (int j = 0; j < 16; ++j) sm[phys(k1, k2, bitrev(j, 4))] = u[j];
    }
}

// Phase B': tid = k1*16 + n3; conj twiddle, IDFT32 over k2 -> n2.
RRC_HD void phase_bi(int tid, const float2* tw2, float2* sm) {
    const int k1 = tid >> 4, n3 = tid & 15;
    float2 v[32];
#pragma unroll
    for (int k2 = 0; k2 < 32; ++k2) v[k2] = cmul_conj(sm[phys(k1, k2, n3)], tw2[k2 * 16 + n3]);
    dif<32, -1>(v);
#pragma unroll
    for (int j = 0; j < 32; ++j) sm[phys(k1, bitrev(j, 5), n3)] = v[j];
}

// Phase A': tid = t; conj twiddle, IDFT32 over k1 -> n1; store valid outputs.
RRC_HD void phase_ai(int tid, long long blk, const BlockIO& io, const float2* tw1, const float2* sm) {
    const int n2 = tid >> 4, n3 = tid & 15;
    float2 p[32];
    powers32(tw1[tid], p);
    float2 v[32];
#pragma unroll
    for (int k1 = 0; k1 < 32; ++k1) v[k1] = cmul_conj(sm[phys(k1, n2, n3)], p[k1]);
    dif<32, -1>(v);
    const long long o0 = blk * (long long)io.V - io.T1;   // output index of segment element 0
#pragma unroll
    for (int j = 0; j < 32; ++j) {
        const int n = tid + 512 * bitrev(j, 5);
        if (n >= io.T1) {
            const long long o = o0 + n;                    // filter output index in this call
            if (io.deci == 1 && io.skip == 0) {
                if (o < io.n_out) io.out[o] = v[j];
            } else {
                const long long r = o - io.skip;
                if (r >= 0 && r % io.deci == 0) {
                    const long long od = r / io.deci;
                    if (od < io.n_out) io.out[od] = v[j];
                }
            }
        }
    }
}

}}  // namespace rrc::fftk
