// fftfilt_tables.hpp — host-side tables for the 16384-point FftFilter kernel:
// the tap spectrum H (f64 FFT, scaled by 1/N like rustradio
// src/fft_filter.rs:151-162, laid out in phase C's row order) and the two
// twiddle tables.  Shared by fftfilt.cu and the CPU emulator (tests/emul).
#pragma once
#include <cmath>
#include <complex>
#include <vector>

#include "fftfilt_core.cuh"

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

}}  // namespace rrc::fftk
