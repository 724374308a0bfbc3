// fftfilt_tables.hpp — host-side tables for the 16384-point FftFilter kernel:
// the tap spectrum H (f64 FFT, scaled by 1/N like rustradio
// src/fft_filter.rs:151-162, laid out in phase C's row order) and the two
// twiddle tables.  Shared by fftfilt.cu and the CPU emulator (tests/emul).
#pragma once
#include <cmath>
#include <complex>
#include <vector>

#include "fftfilt_core.cuh"
#include "fftfilt16_core.cuh"
#include "fftfilt_fold_core.cuh"
#include "fftfilt_pk.cuh"
#include "fftfilt_poly_core.cuh"

namespace rrc { namespace fftk {

// Double-precision radix-2 FFT, in place.
inline void fft_host(std::vector<std::complex<double>>& a) {
    const size_t n = a.size();
    for (size_t i = 1, j = 0; i < n; ++i) {
        size_t bit = n >> 1;
        for (; j & bit; bit >>= 1) j ^= bit;
        j ^= bit;
        if (i < j) std::swap(a[i], a[j]);
    }
    for (size_t len = 2; len <= n; len <<= 1) {
        const double ang = -2.0 * M_PI / (double)len;
        for (size_t k = 0; k < len / 2; ++k) {
            const std::complex<double> w(std::cos(ang * (double)k), std::sin(ang * (double)k));
            for (size_t i = 0; i < n; i += len) {
                const auto u = a[i + k], v = a[i + k + len / 2] * w;
                a[i + k] = u + v;
                a[i + k + len / 2] = u - v;
            }
        }
    }
}

// taps: ntaps interleaved c32.  Hp[P*16 + k3] = H[k1 + 32*k2 + 1024*k3] / N,
// P = k1*32 + k2.  tw1[t] = W_N^t (t < 512).  tw2[k2*16 + n3] = W_512^{n3*k2}.
inline void build_tables(const float* taps, size_t ntaps, std::vector<float2>& Hp,
                         std::vector<float2>& tw1, std::vector<float2>& tw2) {
    std::vector<std::complex<double>> H(N);
    for (size_t k = 0; k < ntaps; ++k) H[k] = {(double)taps[2 * k], (double)taps[2 * k + 1]};
    fft_host(H);
    Hp.resize(N); tw1.resize(512); tw2.resize(512);
    for (int P = 0; P < 1024; ++P) {
        const int k1 = P >> 5, k2 = P & 31;
        for (int j = 0; j < 16; ++j) {
            const int k = k1 + 32 * k2 + 1024 * j;
            const auto v = H[k] / (double)N;
            Hp[(size_t)P * 16 + j] = make_float2((float)v.real(), (float)v.imag());
        }
    }
    for (int t = 0; t < 512; ++t) {
        const double a = -2.0 * M_PI * (double)t / (double)N;
        tw1[t] = make_float2((float)std::cos(a), (float)std::sin(a));
    }
    for (int k2 = 0; k2 < 32; ++k2)
        for (int n3 = 0; n3 < 16; ++n3) {
            const double a = -2.0 * M_PI * (double)(n3 * k2) / 512.0;
            tw2[k2 * 16 + n3] = make_float2((float)std::cos(a), (float)std::sin(a));
        }
}

// Tables of the packed kernel (fftfilt_pk.cuh), thread tid = k1*16 + l:
//   Hq[tid*32 + k3]      = (Re H[k1 + 32 l + 1024 k3], Re H[k1 + 32 (l+16) + 1024 k3]) / N
//   Hq[tid*32 + 16 + k3] = the imaginary parts
//   tw2p[(2j + c)*16 + n3] = component c of (W_512^{n3 j}, W_512^{n3 (j+16)})
inline void build_tables_pk(const float* taps, size_t ntaps, std::vector<float2>& Hq, std::vector<float2>& tw2p) {
    std::vector<std::complex<double>> H(N);
    for (size_t k = 0; k < ntaps; ++k) H[k] = {(double)taps[2 * k], (double)taps[2 * k + 1]};
    fft_host(H);
    Hq.resize(2 * (size_t)N); tw2p.resize(512);
    for (int tid = 0; tid < 512; ++tid) {
        const int k1 = tid >> 4, l = tid & 15;
        for (int k3 = 0; k3 < 16; ++k3) {
            const auto a = H[k1 + 32 * l + 1024 * k3] / (double)N, b = H[k1 + 32 * (l + 16) + 1024 * k3] / (double)N;
            Hq[(size_t)tid * 32 + k3] = make_float2((float)a.real(), (float)b.real());
            Hq[(size_t)tid * 32 + 16 + k3] = make_float2((float)a.imag(), (float)b.imag());
        }
    }
    for (int j = 0; j < 16; ++j)
        for (int n3 = 0; n3 < 16; ++n3) {
            const double a0 = -2.0 * M_PI * (double)(n3 * j) / 512.0, a1 = -2.0 * M_PI * (double)(n3 * (j + 16)) / 512.0;
            tw2p[(2 * j) * 16 + n3] = make_float2((float)std::cos(a0), (float)std::cos(a1));
            tw2p[(2 * j + 1) * 16 + n3] = make_float2((float)std::sin(a0), (float)std::sin(a1));
        }
}

// Tables of the 1024-thread variant (fftfilt16_core.cuh):
// Hd[tid*16 + j*4 + k4] = H[k1 + 16*k2 + 256*(4q + j) + 4096*k4] / N for tid = k1*64 + q*16 + k2;
// tw1[t] = W_N^t (t < 1024); tw2[k2*64 + m] = W_1024^{m*k2}; tw3[k3*4 + n4] = W_64^{n4*k3}.
inline void build_tables16(const float* taps, size_t ntaps, std::vector<float2>& Hd, std::vector<float2>& tw1,
                           std::vector<float2>& tw2, std::vector<float2>& tw3) {
    std::vector<std::complex<double>> H(N);
    for (size_t k = 0; k < ntaps; ++k) H[k] = {(double)taps[2 * k], (double)taps[2 * k + 1]};
    fft_host(H);
    Hd.resize(N); tw1.resize(1024); tw2.resize(1024); tw3.resize(64);
    for (int tid = 0; tid < 1024; ++tid) {
        const int k1 = tid >> 6, q = (tid >> 4) & 3, k2 = tid & 15;
        for (int j = 0; j < 4; ++j)
            for (int k4 = 0; k4 < 4; ++k4) {
                const int k = k1 + 16 * k2 + 256 * (4 * q + j) + 4096 * k4;
                const auto v = H[k] / (double)N;
                Hd[(size_t)tid * 16 + j * 4 + k4] = make_float2((float)v.real(), (float)v.imag());
            }
    }
    auto w = [](double num, double den) {
        const double a = -2.0 * M_PI * num / den;
        return make_float2((float)std::cos(a), (float)std::sin(a));
    };
    for (int t = 0; t < 1024; ++t) tw1[t] = w(t, N);
    for (int k2 = 0; k2 < 16; ++k2)
        for (int m = 0; m < 64; ++m) tw2[k2 * 64 + m] = w(m * k2, 1024.0);
    for (int k3 = 0; k3 < 16; ++k3)
        for (int n4 = 0; n4 < 4; ++n4) tw3[k3 * 4 + n4] = w(n4 * k3, 64.0);
}

}}  // namespace rrc::fftk

namespace rrc { namespace fftf {

// Tables of the decimate-by-8 fold kernel (fftfilt_fold_core.cuh), NBIG = nc * 16384:
//   Hc[c*16384 + (k1*32 + k2)*16 + k3] = H_NBIG[c + nc*(k1 + 32 k2 + 1024 k3)] / NBIG
//   gc[c*512 + t] = W_NBIG^{c t};  twc[c*32 + n1] = W_NBIG^{512 c n1};  twm[m2] = W_{2048 nc}^{m2}.
inline void build_fold_tables(const float* taps, size_t ntaps, int nc, std::vector<float2>& Hc,
                              std::vector<float2>& gc, std::vector<float2>& twc, std::vector<float2>& twm) {
    const size_t NBIG = (size_t)nc * fftk::N;
    std::vector<std::complex<double>> H(NBIG);
    for (size_t k = 0; k < ntaps; ++k) H[k] = {(double)taps[2 * k], (double)taps[2 * k + 1]};
    fftk::fft_host(H);
    auto w = [](double num, double den) {
        const double a = -2.0 * M_PI * std::fmod(num, den) / den;
        return make_float2((float)std::cos(a), (float)std::sin(a));
    };
    Hc.resize(NBIG); gc.resize((size_t)nc * 512); twc.resize((size_t)nc * 32); twm.resize(LU);
    for (int c = 0; c < nc; ++c) {
        for (int P = 0; P < 1024; ++P) {
            const int k1 = P >> 5, k2 = P & 31;
            for (int k3 = 0; k3 < 16; ++k3) {
                const size_t k = (size_t)c + (size_t)nc * (size_t)(k1 + 32 * k2 + 1024 * k3);
                const auto v = H[k] / (double)NBIG;
                Hc[(size_t)c * fftk::N + (size_t)P * 16 + k3] = make_float2((float)v.real(), (float)v.imag());
            }
        }
        for (int t = 0; t < 512; ++t) gc[(size_t)c * 512 + t] = w((double)c * t, (double)NBIG);
        for (int n1 = 0; n1 < 32; ++n1) twc[(size_t)c * 32 + n1] = w((double)c * 512.0 * n1, (double)NBIG);
    }
    for (int m2 = 0; m2 < LU; ++m2) twm[m2] = w((double)m2, (double)LU * nc);
}

}}  // namespace rrc::fftf

namespace rrc { namespace fftp {

// Spectra of the D polyphase branches h_r[q] = h[D q + smod - r] (fftfilt_poly_core.cuh; smod = skip mod D), each in phase
// C's row order:  Hph[r*16384 + (k1*32 + k2)*16 + k3] = FFT_16384(h_r)[k1 + 32 k2 + 1024 k3] / 16384.
// Followed by the rows k2 = l of every branch in the padded shared-memory layout of fftk::load_hres (one bulk copy per
// branch):  Hph[D*16384 + r*HRES_ELEMS + tid*HRES_PITCH + k3].
inline void build_poly_tables(const float* taps, size_t ntaps, int D, int smod, std::vector<float2>& Hph) {
    Hph.assign((size_t)D * (fftk::N + fftk::HRES_ELEMS), make_float2(0.f, 0.f));
    std::vector<std::complex<double>> H(fftk::N);
    for (int r = 0; r < D; ++r) {
        std::fill(H.begin(), H.end(), std::complex<double>(0.0, 0.0));
        for (long long q = 0; q < fftk::N; ++q) {
            const long long j = (long long)D * q + smod - r;
            if (j >= 0 && j < (long long)ntaps) H[q] = {(double)taps[2 * j], (double)taps[2 * j + 1]};
        }
        fftk::fft_host(H);
        for (int P = 0; P < 1024; ++P) {
            const int k1 = P >> 5, k2 = P & 31;
            for (int k3 = 0; k3 < 16; ++k3) {
                const auto v = H[k1 + 32 * k2 + 1024 * k3] / (double)fftk::N;
                Hph[(size_t)r * fftk::N + (size_t)P * 16 + k3] = make_float2((float)v.real(), (float)v.imag());
            }
        }
        for (int t = 0; t < fftk::NT; ++t)
            fftk::load_hres(t, Hph.data() + (size_t)r * fftk::N, Hph.data() + (size_t)D * fftk::N + (size_t)r * fftk::HRES_ELEMS);
    }
}

}}  // namespace rrc::fftp
