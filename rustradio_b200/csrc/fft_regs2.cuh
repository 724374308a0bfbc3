// fft_regs2.cuh — in-register DFTs on PAIRS of complex values (packed FP32, sm_100a FFMA2/FADD2/FMUL2).
//
// F2 holds two floats in one 64-bit register pair; C2 = {F2 re, im} is a pair of complex numbers
// (lane a, lane b) that go through exactly the same operation sequence with the same twiddles — two
// columns of the shared-memory FFT.  One packed instruction does the work of two scalar ones in one
// issue slot (measured: FFMA2 1.98 warp-instr/clk/SM = the same 127 FMA lanes/clk/SM as scalar FFMA),
// and a compile-time twiddle (c, c) is encoded as an FFMA2 immediate.  PTX has no negated-operand
// form for fma.rn.f32x2, so the "2a - s" half of the 6-FMA butterfly is done with scalar FFMA on the
// two halves of the pair (register aliasing, no moves): issue 8 per two butterflies instead of 12,
// FP-pipe work unchanged.  Host build: plain float pairs (CPU emulator).
#pragma once
#include "fft_regs.cuh"

namespace rrc { namespace fftr {

#if defined(__CUDA_ARCH__)
struct F2 { unsigned long long v; };
__device__ __forceinline__ F2 f2(float a, float b) { F2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r.v) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ float f2_lo(F2 x) { float a, b; asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(x.v)); return a; }
__device__ __forceinline__ float f2_hi(F2 x) { float a, b; asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(x.v)); return b; }
__device__ __forceinline__ F2 operator+(F2 x, F2 y) { F2 r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(x.v), "l"(y.v)); return r; }
__device__ __forceinline__ F2 operator-(F2 x, F2 y) { F2 r; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(x.v), "l"(y.v)); return r; }
__device__ __forceinline__ F2 operator*(F2 x, F2 y) { F2 r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(x.v), "l"(y.v)); return r; }
__device__ __forceinline__ F2 fma2(F2 x, F2 y, F2 z) { F2 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r.v) : "l"(x.v), "l"(y.v), "l"(z.v)); return r; }
#else
struct F2 { float a, b; };
inline F2 f2(float a, float b) { return F2{a, b}; }
inline float f2_lo(F2 x) { return x.a; }
inline float f2_hi(F2 x) { return x.b; }
inline F2 operator+(F2 x, F2 y) { return F2{x.a + y.a, x.b + y.b}; }
inline F2 operator-(F2 x, F2 y) { return F2{x.a - y.a, x.b - y.b}; }
inline F2 operator*(F2 x, F2 y) { return F2{x.a * y.a, x.b * y.b}; }
inline F2 fma2(F2 x, F2 y, F2 z) { return F2{fmaf(x.a, y.a, z.a), fmaf(x.b, y.b, z.b)}; }
#endif
RRC_HD F2 splat(float c) { return f2(c, c); }
// x*c + z with a scalar constant c (an FFMA2 immediate when c is a compile-time constant)
RRC_HD F2 fmac(F2 x, float c, F2 z) { return fma2(x, splat(c), z); }
// 2*a - s, lane-wise, with scalar FFMA (negated addend) on the halves of the pair
RRC_HD F2 twice_minus(F2 a, F2 s) { return f2(fmaf(2.0f, f2_lo(a), -f2_lo(s)), fmaf(2.0f, f2_hi(a), -f2_hi(s))); }
// x*y - z and z - x*y style helpers with a negated product/addend, on the halves
RRC_HD F2 fms_halves(F2 x, F2 y, F2 z) {   // x*y - z
    return f2(fmaf(f2_lo(x), f2_lo(y), -f2_lo(z)), fmaf(f2_hi(x), f2_hi(y), -f2_hi(z)));
}

struct C2 { F2 re, im; };
RRC_HD C2 c2_add(C2 a, C2 b) { return C2{a.re + b.re, a.im + b.im}; }
RRC_HD C2 c2_sub(C2 a, C2 b) { return C2{a.re - b.re, a.im - b.im}; }
// (a * w) lane-wise, w a pair of runtime complex values: 5 issue slots (3 packed + 2 scalar), 8 FMA-lane slots
RRC_HD C2 c2_mul(C2 a, C2 w) {
    C2 r;
    r.re = fms_halves(a.re, w.re, a.im * w.im);          // a.re*w.re - a.im*w.im
    r.im = fma2(a.re, w.im, a.im * w.re);
    return r;
}
RRC_HD C2 c2_mul_conj(C2 a, C2 w) {                      // a * conj(w)
    C2 r;
    r.re = fma2(a.re, w.re, a.im * w.im);
    r.im = fms_halves(a.im, w.re, a.re * w.im);          // a.im*w.re - a.re*w.im
    return r;
}
RRC_HD C2 c2_sqr(C2 a) {
    C2 r;
    r.re = fms_halves(a.re, a.re, a.im * a.im);
    const F2 t = a.re * a.im;
    r.im = t + t;
    return r;
}

// bfly2<NUM, DEN, DIR>(a, b):  a <- a + w*b,  b <- a - w*b,  w = exp(-DIR*2*pi*i*NUM/DEN), on pairs.
template <int NUM, int DEN, int DIR>
RRC_HD void bfly2(C2& a, C2& b) {
    constexpr int n64 = NUM * (64 / DEN);
    if constexpr (n64 == 0) {
        const C2 s = c2_add(a, b), d = c2_sub(a, b);
        a = s; b = d;
    } else if constexpr (n64 == 16) {              // w*b = (b.im, -b.re) fwd ; (-b.im, b.re) inv
        C2 s, d;
        if constexpr (DIR > 0) { s.re = a.re + b.im; s.im = a.im - b.re; d.re = a.re - b.im; d.im = a.im + b.re; }
        else                   { s.re = a.re - b.im; s.im = a.im + b.re; d.re = a.re + b.im; d.im = a.im - b.re; }
        a = s; b = d;
    } else if constexpr (n64 == 8 || n64 == 24) {
        constexpr float c = (float)0.70710678118654752440;
        F2 p, q;                                   // w*b = c*p + i*c*q  (signs folded into sp, sq)
        float sp, sq;
        if constexpr (n64 == 8) {
            p = DIR > 0 ? b.re + b.im : b.re - b.im;  sp = 1.f;
            q = DIR > 0 ? b.im - b.re : b.im + b.re;  sq = 1.f;
        } else {
            p = DIR > 0 ? b.im - b.re : b.re + b.im;  sp = DIR > 0 ? 1.f : -1.f;
            q = DIR > 0 ? b.re + b.im : b.re - b.im;  sq = DIR > 0 ? -1.f : 1.f;
        }
        C2 s, d;
        s.re = fmac(p, sp * c, a.re);  s.im = fmac(q, sq * c, a.im);
        d.re = fmac(p, -sp * c, a.re); d.im = fmac(q, -sq * c, a.im);
        a = s; b = d;
    } else {
        constexpr float wr = (float)cos64(n64);
        constexpr float wi = (float)(-DIR * sin64(n64));
        C2 s;
        s.re = fmac(b.im, -wi, fmac(b.re, wr, a.re));
        s.im = fmac(b.re, wi, fmac(b.im, wr, a.im));
        C2 d;
        d.re = twice_minus(a.re, s.re);
        d.im = twice_minus(a.im, s.im);
        a = s; b = d;
    }
}

template <int N, int DIR, int I>
struct DitLevel2 {
    static RRC_HD void run(C2* v) {
        bfly2<I, N, DIR>(v[I], v[I + N / 2]);
        if constexpr (I + 1 < N / 2) DitLevel2<N, DIR, I + 1>::run(v);
    }
};

// dit2<N, DIR>(v, one): BIT-REVERSED input, natural-order output, on pairs.  The first level is written
// a +- one*b (one == 1.0f at run time, same trick and purpose as dit_g).
template <int N, int DIR>
RRC_HD void dit2(C2* v, float one) {
    static_assert(N >= 2 && N <= 64 && (N & (N - 1)) == 0, "N must be a power of two, 2..64");
    if constexpr (N == 2) {
        const C2 a = v[0], b = v[1];
        const F2 o = splat(one), no = splat(-one);
        v[0] = C2{fma2(b.re, o, a.re), fma2(b.im, o, a.im)};
        v[1] = C2{fma2(b.re, no, a.re), fma2(b.im, no, a.im)};
    } else {
        dit2<N / 2, DIR>(v, one);
        dit2<N / 2, DIR>(v + N / 2, one);
        DitLevel2<N, DIR, 0>::run(v);
    }
}

}}  // namespace rrc::fftr
