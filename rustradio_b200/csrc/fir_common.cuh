// fir_common.cuh — device helpers shared by the FP32 FIR kernels (fir.cu) and the tensor-core ones (fir_tc.cu).
#pragma once
#include <cuda_runtime.h>

namespace rrc {

// RtlSdrDecode fused into the tile load (src/rtlsdr_decode.rs:35-43).
__device__ __forceinline__ float2 decode_iq(unsigned int w) {
    return make_float2(__fmul_rn(__fsub_rn((float)(w & 0xffu), 127.0f), 0.008f), __fmul_rn(__fsub_rn((float)(w >> 8), 127.0f), 0.008f));
}

// atan2 with |error| < 1e-6 rad over the whole plane (bar: 1e-4 rad): octant reduction to
// q = min/max in [0,1], odd minimax polynomial of degree 15, then quadrant fix-ups.  About
// half the instructions of libdevice's atan2f and no slow path.  atan2(0, 0) = 0 like libm.
__device__ __forceinline__ float fast_atan2(float y, float x) {
    const float ax = fabsf(x), ay = fabsf(y);
    const float mx = fmaxf(ax, ay), mn = fminf(ax, ay);
    const float q = mx == 0.0f ? 0.0f : __fdividef(mn, mx);
    const float s = q * q;
    float r = -0.0040540580f;
    r = fmaf(r, s, 0.0218612288f);
    r = fmaf(r, s, -0.0559098861f);
    r = fmaf(r, s, 0.0964200441f);
    r = fmaf(r, s, -0.1390853351f);
    r = fmaf(r, s, 0.1994653599f);
    r = fmaf(r, s, -0.3332985605f);
    r = fmaf(r, s, 0.9999993329f);
    r = r * q;
    if (ay > ax) r = 1.57079632679489662f - r;
    if (x < 0.0f) r = 3.14159265358979324f - r;
    return copysignf(r, y);
}

__device__ __forceinline__ float demod_pair(float2 a, float2 b, float gain) {
    // conj(a) * b, then gain * atan2(im, re)  (src/quadrature_demod.rs:71-73,106-108)
    float re = fmaf(a.x, b.x, a.y * b.y);
    float im = fmaf(a.x, b.y, -(a.y * b.x));
    return gain * fast_atan2(im, re);
}

// exp(-j*2*pi*ratio*k) evaluated from the exact f64 angle.
__device__ __forceinline__ float2 rotator(double ratio, unsigned long long k) {
    double r = ratio * (double)k;
    r -= rint(r);
    double s, c;
    sincospi(-2.0 * r, &s, &c);
    return make_float2((float)c, (float)s);
}
__device__ __forceinline__ float2 cmulf(float2 a, float2 b) {
    return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
}  // namespace rrc
