// epilogue.cuh — the sample-wise neighbours of the filters as STORE EPILOGUES: instead of a
// separate block that costs a full HBM round trip, the filter's own store applies
//   MultiplyConst<Complex>::process_sync   y * val            rustradio src/multiply_const.rs:16-23
//   AddConst<Complex>::process_sync        y + val            src/add_const.rs:36-44
//   ComplexToMag2::process_sync            y.norm_sqr()       src/complex_to_mag2.rs:17-20   (output becomes f32)
// with the same separately rounded operations as the stand-alone kernels (elementwise.cu), so
// filter -> neighbour fused equals the two blocks run back to back bit for bit.
#pragma once
#include <cuda_runtime.h>

#ifndef RRC_EPI_NONE
#define RRC_EPI_NONE 0
#define RRC_EPI_MULTIPLY_CONST 1
#define RRC_EPI_ADD_CONST 2
#define RRC_EPI_MAG2 3
#endif

namespace rrc {
struct Epi {
    int kind = RRC_EPI_NONE;
    float re = 0.f, im = 0.f;
};
#if defined(__CUDACC__)
#define RRC_EPI_HD __host__ __device__ __forceinline__
#else
#define RRC_EPI_HD inline
#endif
// kinds 1, 2: Complex -> Complex (num-complex multiplication order, nothing contracted)
RRC_EPI_HD float2 epi_c32(float2 y, const Epi& e) {
#if defined(__CUDA_ARCH__)
    if (e.kind == RRC_EPI_MULTIPLY_CONST)
        return make_float2(__fsub_rn(__fmul_rn(y.x, e.re), __fmul_rn(y.y, e.im)), __fadd_rn(__fmul_rn(y.x, e.im), __fmul_rn(y.y, e.re)));
    if (e.kind == RRC_EPI_ADD_CONST) return make_float2(__fadd_rn(y.x, e.re), __fadd_rn(y.y, e.im));
    return y;
#else
    if (e.kind == RRC_EPI_MULTIPLY_CONST) {
        volatile float a = y.x * e.re, b = y.y * e.im, c = y.x * e.im, d = y.y * e.re;
        return make_float2(a - b, c + d);
    }
    if (e.kind == RRC_EPI_ADD_CONST) return make_float2(y.x + e.re, y.y + e.im);
    return y;
#endif
}
// kind 3: Complex -> f32
RRC_EPI_HD float epi_mag2(float2 y) {
#if defined(__CUDA_ARCH__)
    return __fadd_rn(__fmul_rn(y.x, y.x), __fmul_rn(y.y, y.y));
#else
    volatile float a = y.x * y.x, b = y.y * y.y;
    return a + b;
#endif
}
}  // namespace rrc
