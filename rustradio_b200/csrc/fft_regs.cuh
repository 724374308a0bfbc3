// fft_regs.cuh — compile-time-unrolled in-register DFTs (N = 2..32) for the
// shared-memory FFT used by the overlap-save FftFilter kernel.
//
// dif<N, DIR>(v): decimation-in-frequency radix-2 recursion on v[0..N).
//   natural-order input, BIT-REVERSED output: X[k] ends up in v[bitrev<N>(k)].
//   DIR = +1: forward (W = e^{-2 pi i / N}); DIR = -1: inverse (conjugate), unnormalised.
// All twiddles are compile-time constants; multiplications by 1, -i and
// (+-1 +- i)/sqrt(2) are special-cased.
#pragma once
#include <cuda_runtime.h>

namespace rrc { namespace fftr {

#if defined(__CUDACC__)
#define RRC_HD __host__ __device__ __forceinline__
#else
#define RRC_HD inline
#endif

// cos(2*pi*i/64), i = 0..16 (first quadrant in 1/64-turn steps).
RRC_HD constexpr double cos64_q(int i) {
    constexpr double t[17] = {
        1.0,
        0.99518472667219688624, 0.98078528040323044913, 0.95694033573220886494,
        0.92387953251128675613, 0.88192126434835502971, 0.83146961230254523708,
        0.77301045336273696081, 0.70710678118654752440, 0.63439328416364549822,
        0.55557023301960222474, 0.47139673682599764856, 0.38268343236508977173,
        0.29028467725446236764, 0.19509032201612826785, 0.09801714032956060199,
        0.0};
    return t[i];
}
// cos / sin of 2*pi*num/64 for any integer num.
RRC_HD constexpr double cos64(int num) {
    num = ((num % 64) + 64) % 64;
    if (num <= 16) return cos64_q(num);
    if (num <= 32) return -cos64_q(32 - num);
    if (num <= 48) return -cos64_q(num - 32);
    return cos64_q(64 - num);
}
RRC_HD constexpr double sin64(int num) { return cos64(num - 16); }

RRC_HD constexpr int bitrev(int x, int bits) {
    int r = 0;
    for (int i = 0; i < bits; ++i) r |= ((x >> i) & 1) << (bits - 1 - i);
    return r;
}
RRC_HD constexpr int ilog2(int n) { return n <= 1 ? 0 : 1 + ilog2(n >> 1); }

// RtlSdrDecode fused into a kernel's first load (rustradio src/rtlsdr_decode.rs:35-43; SURVEY 8f
// rank 1): one sample is two bytes (I, Q) and decodes to ((I - 127) * 0.008, (Q - 127) * 0.008),
// the subtraction rounded before the multiplication, exactly like the reference block.
RRC_HD float2 decode_u8iq(unsigned int lo, unsigned int hi) {
#if defined(__CUDA_ARCH__)
    return make_float2(__fmul_rn(__fsub_rn((float)lo, 127.0f), 0.008f), __fmul_rn(__fsub_rn((float)hi, 127.0f), 0.008f));
#else
    volatile float a = (float)lo - 127.0f, b = (float)hi - 127.0f;
    return make_float2(a * 0.008f, b * 0.008f);
#endif
}
// Sample g of an input stream that is either Complex<f32> (u8 == 0) or u8 I/Q pairs (2-byte aligned).
RRC_HD float2 ld_iq(const float2* in, long long g, int u8) {
    if (u8) {
        const unsigned int w = reinterpret_cast<const unsigned short*>(in)[g];
        return decode_u8iq(w & 0xffu, w >> 8);
    }
    return in[g];
}

RRC_HD float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
RRC_HD float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
RRC_HD float2 cmul(float2 a, float2 b) {
    return make_float2(fmaf(a.x, b.x, -(a.y * b.y)), fmaf(a.x, b.y, a.y * b.x));
}
RRC_HD float2 cmul_conj(float2 a, float2 b) {   // a * conj(b)
    return make_float2(fmaf(a.x, b.x, a.y * b.y), fmaf(a.y, b.x, -(a.x * b.y)));
}
RRC_HD float2 csqr(float2 a) { return make_float2(fmaf(a.x, a.x, -(a.y * a.y)), 2.0f * a.x * a.y); }

// d * exp(-DIR * 2*pi*i * NUM/DEN), NUM/DEN in [0, 1/2), DEN | 64.
template <int NUM, int DEN, int DIR>
RRC_HD float2 mul_w(float2 d) {
    constexpr int n64 = NUM * (64 / DEN);          // angle in 1/64 turns
    if constexpr (n64 == 0) {
        return d;
    } else if constexpr (n64 == 16) {              // -i (fwd) / +i (inv)
        return DIR > 0 ? make_float2(d.y, -d.x) : make_float2(-d.y, d.x);
    } else if constexpr (n64 == 8) {               // (1 -+ i)/sqrt2
        constexpr float c = (float)0.70710678118654752440;
        return DIR > 0 ? make_float2(c * (d.x + d.y), c * (d.y - d.x))
                       : make_float2(c * (d.x - d.y), c * (d.y + d.x));
    } else if constexpr (n64 == 24) {              // (-1 -+ i)/sqrt2
        constexpr float c = (float)0.70710678118654752440;
        return DIR > 0 ? make_float2(c * (d.y - d.x), -c * (d.x + d.y))
                       : make_float2(-c * (d.x + d.y), c * (d.x - d.y));
    } else {
        constexpr float wr = (float)cos64(n64);
        constexpr float wi = (float)(-DIR * sin64(n64));
        return make_float2(fmaf(d.x, wr, -(d.y * wi)), fmaf(d.x, wi, d.y * wr));
    }
}

template <int N, int DIR, int I>
struct DifLevel {
    static RRC_HD void run(float2* v) {
        const float2 a = v[I], b = v[I + N / 2];
        v[I] = cadd(a, b);
        v[I + N / 2] = mul_w<I, N, DIR>(csub(a, b));
        if constexpr (I + 1 < N / 2) DifLevel<N, DIR, I + 1>::run(v);
    }
};

template <int N, int DIR>
RRC_HD void dif(float2* v) {
    static_assert(N >= 1 && N <= 64 && (N & (N - 1)) == 0, "N must be a power of two <= 64");
    if constexpr (N >= 2) {
        DifLevel<N, DIR, 0>::run(v);
        dif<N / 2, DIR>(v);
        dif<N / 2, DIR>(v + N / 2);
    }
}

// ---- decimation-in-time with FMA-fused butterflies --------------------------------
// bfly<NUM, DEN, DIR>(a, b):  a <- a + w*b,  b <- a - w*b,  w = exp(-DIR*2*pi*i*NUM/DEN).
// General twiddle: 6 FFMA (sum by two FMA chains, difference as 2a - sum) instead of the
// 2 FMUL + 2 FFMA + 4 FADD of the textbook form; +-(1 -+ i)/sqrt2 twiddles: 2 FADD + 4 FFMA.
template <int NUM, int DEN, int DIR>
RRC_HD void bfly(float2& a, float2& b) {
    constexpr int n64 = NUM * (64 / DEN);
    if constexpr (n64 == 0) {
        const float2 s = cadd(a, b), d = csub(a, b);
        a = s; b = d;
    } else if constexpr (n64 == 16) {              // w = -i (fwd): w*b = (b.y, -b.x); +i (inv): (-b.y, b.x)
        const float2 t = DIR > 0 ? make_float2(b.y, -b.x) : make_float2(-b.y, b.x);
        const float2 s = cadd(a, t), d = csub(a, t);
        a = s; b = d;
    } else if constexpr (n64 == 8 || n64 == 24) {
        constexpr float c = (float)0.70710678118654752440;
        // n64 == 8 : w = c(1 - i) fwd, c(1 + i) inv ; n64 == 24: w = c(-1 - i) fwd, c(-1 + i) inv
        // w*b = cr*(p) + i*cr*(q) with p, q = +-b.x +- b.y
        float p, q;
        if constexpr (n64 == 8) {
            p = DIR > 0 ? b.x + b.y : b.x - b.y;
            q = DIR > 0 ? b.y - b.x : b.y + b.x;
        } else {
            p = DIR > 0 ? b.y - b.x : -(b.x + b.y);
            q = DIR > 0 ? -(b.x + b.y) : b.x - b.y;
        }
        const float2 s = make_float2(fmaf(c, p, a.x), fmaf(c, q, a.y));
        const float2 d = make_float2(fmaf(-c, p, a.x), fmaf(-c, q, a.y));
        a = s; b = d;
    } else {
        constexpr float wr = (float)cos64(n64);
        constexpr float wi = (float)(-DIR * sin64(n64));
        float2 s;
        s.x = fmaf(-wi, b.y, fmaf(wr, b.x, a.x));
        s.y = fmaf(wi, b.x, fmaf(wr, b.y, a.y));
        const float2 d = make_float2(fmaf(2.0f, a.x, -s.x), fmaf(2.0f, a.y, -s.y));
        a = s; b = d;
    }
}

template <int N, int DIR, int I>
struct DitLevel {
    static RRC_HD void run(float2* v) {
        bfly<I, N, DIR>(v[I], v[I + N / 2]);
        if constexpr (I + 1 < N / 2) DitLevel<N, DIR, I + 1>::run(v);
    }
};

// dit<N, DIR>(v): BIT-REVERSED input (v[bitrev<N>(n)] = x[n]), natural-order output (v[k] = X[k]).
template <int N, int DIR>
RRC_HD void dit(float2* v) {
    static_assert(N >= 1 && N <= 64 && (N & (N - 1)) == 0, "N must be a power of two <= 64");
    if constexpr (N >= 2) {
        dit<N / 2, DIR>(v);
        dit<N / 2, DIR>(v + N / 2);
        DitLevel<N, DIR, 0>::run(v);
    }
}

// dit_g: dit with the first-level butterflies written as a +- one*b (one == 1.0f at run time, so
// the results are bit-identical to a +- b at the same instruction count).  Every value of the
// transform then depends on `one`; a kernel that obtains `one` from a volatile load placed after a
// barrier thereby keeps ptxas from hoisting the arithmetic above that barrier (it otherwise does:
// register-only FP instructions are freely scheduled across BAR.SYNC).
template <int N, int DIR>
RRC_HD void dit_g(float2* v, float one) {
    static_assert(N >= 2 && N <= 64 && (N & (N - 1)) == 0, "N must be a power of two, 2..64");
    if constexpr (N == 2) {
        const float2 a = v[0], b = v[1];
        v[0] = make_float2(fmaf(b.x, one, a.x), fmaf(b.y, one, a.y));
        v[1] = make_float2(fmaf(-b.x, one, a.x), fmaf(-b.y, one, a.y));
    } else {
        dit_g<N / 2, DIR>(v, one);
        dit_g<N / 2, DIR>(v + N / 2, one);
        DitLevel<N, DIR, 0>::run(v);
    }
}

}}  // namespace rrc::fftr
