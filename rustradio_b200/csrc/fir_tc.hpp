// fir_tc.hpp — argument structs and host entry points of the tensor-core FIR kernels (fir_tc.cuh / fir_tc.cu).
#pragma once
#include <cuda_runtime.h>

#include <cstddef>

namespace rrc {

struct FirTcArgs {
    const float2* in;
    void* out;
    const uint4* bfrag;        // [KS][NTILE][32] = {b_hi[0], b_hi[1], b_lo[0], b_lo[1]} per lane (scaled taps, fp16x2)
    const float* taps_rev;     // w[j] in f32 (boundary output of the demod epilogue)
    long long in_stride, out_stride, need, out_n;
    long long tiles_x, total_tiles;
    int ntaps, deci;
    int RS;                    // samples per block-row = R * deci (a multiple of 8)
    int PAD;                   // fp16 elements inserted after every RS staged samples (0 or 8): makes the
                               // byte stride between block-rows an odd multiple of 16 -> conflict-free ldmatrix
    unsigned magic;            // ceil(2^32 / RS): s / RS = umulhi(s, magic) for s < 2^16
    int KS;                    // k-steps of 16
    int NM;                    // m-tiles (8 block-rows each) per warp tile
    int L;                     // staged samples per tile (multiple of 8, <= 64 * NLD)
    int PL;                    // fp16 elements per plane (multiple of 8)
    int WB;                    // bytes of shared memory per warp (planes + ytile), multiple of 16
    float gain;
    float tap_inv_scale;       // 1 / (power of two the taps were multiplied by)
    int in_u8;                 // 1: `in` is u8 I/Q pairs (RtlSdrDecode fused into the tile load; in_stride in samples)
};

struct FirTc1Args {
    const float2* in;
    void* out;
    const uint4* bfrag;        // [KS][32], NTILE = 1 layout
    const float* taps_rev;
    long long in_stride, out_stride, need, out_n;
    long long tiles_x, total_tiles;
    int ntaps;
    float gain, tap_inv_scale;
    int in_u8;                 // 1: `in` is u8 I/Q pairs (RtlSdrDecode fused into the tile load; in_stride in samples)
};

struct FirTcfArgs {
    const float* in;
    float* out;
    const uint4* bfrag;        // [KS][32], NTILE = 1 layout
    long long in_stride, out_stride, need, out_n;
    long long tiles_x, total_tiles;
    float tap_inv_scale;
};

struct FirTccArgs {             // fir_tcc_kernel: c32 samples, complex taps (translate filters)
    const float2* in;
    void* out;
    const uint4* bfrag;        // [KS][2 = Re(w), Im(w)][32]
    const float2* taps_rev_c;  // w[j] in f32 (boundary output of the demod epilogue)
    long long in_stride, out_stride, need, out_n;
    long long tiles_x, total_tiles;
    int ntaps;
    float gain, tap_inv_scale;
    int in_u8;                 // 1: `in` is u8 I/Q pairs (in_stride in samples)
    int translate;             // apply the per-output rotator exp(-j*2*pi*ratio*((ntaps-1) + (out_base + i)*deci))
    double ratio;
    unsigned long long out_base;
};

struct FirTc5Args {             // fir_tc5_kernel (tcgen05 / TMEM, fir_tc5.cu): c32 or f32 samples, real taps, deci 1, ntaps <= 65
    const float2* in;          // f32 streams: float arrays behind these pointers (real_stream = 1), strides in samples
    float2* out;
    long long in_stride, out_stride, need, out_n;
    int real_stream;
    long long tiles_x, total_tiles;   // tiles of 128 * nr outputs per channel / in the launch
    int nr;                    // block-rows per tile (fir_tc5_rows)
    int KS;                    // k-steps of 16: ceil((127 + ntaps) / 16) <= 12
    float tap_inv_scale;
};

constexpr int FIR_TC_THREADS = 256;
constexpr int FIR_TC1_BT = 512;            // INPUT samples a warp tile of fir_tc1_kernel advances by: 512/deci outputs
constexpr int FIR_TCF_IN = 1024;           // INPUT samples a warp tile of fir_tcf_kernel (f32 streams) advances by
constexpr int FIR_TC1_MAX_KS = 20;         // deci 1/2/4 kernel: up to 20 k-steps of 16 samples, i.e. 7*deci + ntaps <= 320

// What the launchers need to know about the filter (filled by plan_tc in fir.cu).
struct FirTcGeom {
    int device;
    int ntile, nld, KS;
    size_t smem;               // generic kernel: dynamic shared memory per CTA
    int deci;                  // fir_tc1_kernel: 1, 2 or 4
};
int fir_tc_launch(const FirTcGeom& g, const FirTcArgs& a, bool demod, cudaStream_t st);
int fir_tc1_launch(const FirTcGeom& g, const FirTc1Args& a, bool demod, cudaStream_t st);
int fir_tcf_launch(const FirTcGeom& g, const FirTcfArgs& a, cudaStream_t st);
int fir_tcc_launch(const FirTcGeom& g, const FirTccArgs& a, bool demod, cudaStream_t st);
size_t fir_tc5_tab_words();       // words of the tap table fir_tc5_build_tab fills (kernel parameter of fir_tc5_kernel)
void fir_tc5_build_tab(const unsigned short* hi, const unsigned short* lo, size_t ntaps, unsigned* tab);
int fir_tc5_rows(long long tiles8192, int device, bool real_stream);   // block-rows of 128 outputs per CTA tile: 32 / 64 (c32), 64 (f32)
int fir_tc5_launch(int device, const FirTc5Args& a, const unsigned* tab_host, cudaStream_t st);

}  // namespace rrc
