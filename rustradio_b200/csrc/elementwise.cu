// elementwise.cu — the sample-wise neighbours of the filtering path on sm_100a.
//
//   MultiplyConst<T>   x * val                      src/multiply_const.rs:16-23
//   AddConst<T>        x + val                      src/add_const.rs:36-44
//   ComplexToMag2      re*re + im*im                src/complex_to_mag2.rs:17-20
//   Tee<T>             (x, x)                       src/tee.rs:20-24
//   IqBalance          mean = mean*(1-a) + x*a; x - mean     src/iq_balance.rs:75-80
//
// The maps are HBM-bound streaming kernels (128-bit loads/stores, grid-stride, 4 vectors in flight
// per thread).  Their arithmetic is written with __fmul_rn/__fadd_rn/__fsub_rn so nothing contracts
// into an FMA: results are BIT-IDENTICAL to the reference's separately rounded f32 operations
// (num-complex: (a+bi)(c+di) = (ac - bd) + (ad + bc)i, src/lib.rs `Complex`).
//
// IqBalance is a first-order linear recurrence.  It is evaluated as an affine scan: a run of samples
// maps the incoming mean m to A*m + B (A = (1-a)^len, B complex), and runs compose associatively.
//   pass 1  (iq_tile_kernel<false>)  per 4096-sample tile: B of every thread's 16 samples -> in-tile
//                                    scan -> tile aggregate                          (8 B/sample)
//   pass 2  (iq_carry_kernel)        one CTA: scan of the tile aggregates -> mean entering each tile
//   pass 3  (iq_tile_kernel<true>)   same in-tile scan, then each thread replays its 16 samples
//                                    from the exact incoming mean and writes x - mean   (16 B/sample)
// A call that fits one tile launches pass 3 only.  24 B/sample against the algorithmic 16: the input is
// read twice (a single-pass decoupled look-back scan is the next step).  Same recurrence, different
// association than the reference's sequential f32 loop => parity is a tolerance (1e-5 rel-RMS), not
// bit-exact.
#include <algorithm>
#include <cmath>

#include "common.cuh"

namespace rrc {

enum MapOp : int { MAP_MUL_F32, MAP_MUL_C32, MAP_ADD_F32, MAP_ADD_C32 };

__device__ __forceinline__ float2 cmul_rn(float2 a, float2 b) {
    return make_float2(__fsub_rn(__fmul_rn(a.x, b.x), __fmul_rn(a.y, b.y)), __fadd_rn(__fmul_rn(a.x, b.y), __fmul_rn(a.y, b.x)));
}

template <int OP>
__device__ __forceinline__ float4 map4(float4 v, float2 c) {
    if constexpr (OP == MAP_MUL_F32) return make_float4(__fmul_rn(v.x, c.x), __fmul_rn(v.y, c.x), __fmul_rn(v.z, c.x), __fmul_rn(v.w, c.x));
    if constexpr (OP == MAP_ADD_F32) return make_float4(__fadd_rn(v.x, c.x), __fadd_rn(v.y, c.x), __fadd_rn(v.z, c.x), __fadd_rn(v.w, c.x));
    if constexpr (OP == MAP_ADD_C32) return make_float4(__fadd_rn(v.x, c.x), __fadd_rn(v.y, c.y), __fadd_rn(v.z, c.x), __fadd_rn(v.w, c.y));
    const float2 a = cmul_rn(make_float2(v.x, v.y), c), b = cmul_rn(make_float2(v.z, v.w), c);
    return make_float4(a.x, a.y, b.x, b.y);
}
template <int OP>
__device__ __forceinline__ void map_tail(const float* in, float* out, size_t i, float2 c) {
    if constexpr (OP == MAP_MUL_F32) out[i] = __fmul_rn(in[i], c.x);
    else if constexpr (OP == MAP_ADD_F32) out[i] = __fadd_rn(in[i], c.x);
    else {   // complex ops: i counts floats, always even here
        const float2 x = make_float2(in[i], in[i + 1]);
        const float2 y = OP == MAP_ADD_C32 ? make_float2(__fadd_rn(x.x, c.x), __fadd_rn(x.y, c.y)) : cmul_rn(x, c);
        out[i] = y.x; out[i + 1] = y.y;
    }
}

// nf floats in, nf floats out (complex streams count 2 floats per sample); in/out 16-byte aligned.
template <int OP>
__global__ void __launch_bounds__(256) map_kernel(const float* __restrict__ in, float* __restrict__ out, size_t nf, float2 c) {
    constexpr int U = 4;
    const size_t nv = nf / 4;
    const float4* __restrict__ in4 = reinterpret_cast<const float4*>(in);
    float4* __restrict__ out4 = reinterpret_cast<float4*>(out);
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (; i + (U - 1) * stride < nv; i += U * stride) {
        float4 v[U];
#pragma unroll
        for (int k = 0; k < U; ++k) v[k] = in4[i + k * stride];
#pragma unroll
        for (int k = 0; k < U; ++k) out4[i + k * stride] = map4<OP>(v[k], c);
    }
    for (; i < nv; i += stride) out4[i] = map4<OP>(in4[i], c);
    if (blockIdx.x == 0 && threadIdx.x == 0)
        for (size_t j = nv * 4; j < nf; j += (OP == MAP_MUL_C32 || OP == MAP_ADD_C32) ? 2 : 1) map_tail<OP>(in, out, j, c);
}
// Unaligned pointers (a ring window may start anywhere): scalar form.
template <int OP>
__global__ void __launch_bounds__(256) map_scalar_kernel(const float* __restrict__ in, float* __restrict__ out, size_t nf, float2 c) {
    constexpr size_t step = (OP == MAP_MUL_C32 || OP == MAP_ADD_C32) ? 2 : 1;
    const size_t stride = (size_t)gridDim.x * blockDim.x * step;
    for (size_t i = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * step; i < nf; i += stride) map_tail<OP>(in, out, i, c);
}

// Complex -> |x|^2: 2 x 128-bit loads (4 samples), one 128-bit store.
__global__ void __launch_bounds__(256) mag2_kernel(const float2* __restrict__ in, float* __restrict__ out, size_t n, int aligned) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (aligned) {
        const size_t nv = n / 4;
        const float4* __restrict__ in4 = reinterpret_cast<const float4*>(in);
        float4* __restrict__ out4 = reinterpret_cast<float4*>(out);
        for (size_t v = i; v < nv; v += stride) {
            const float4 a = in4[2 * v], b = in4[2 * v + 1];
            out4[v] = make_float4(__fadd_rn(__fmul_rn(a.x, a.x), __fmul_rn(a.y, a.y)), __fadd_rn(__fmul_rn(a.z, a.z), __fmul_rn(a.w, a.w)),
                                  __fadd_rn(__fmul_rn(b.x, b.x), __fmul_rn(b.y, b.y)), __fadd_rn(__fmul_rn(b.z, b.z), __fmul_rn(b.w, b.w)));
        }
        i += nv * 4;
        if (i < n) { const float2 x = in[i]; out[i] = __fadd_rn(__fmul_rn(x.x, x.x), __fmul_rn(x.y, x.y)); }
    } else {
        for (; i < n; i += stride) { const float2 x = in[i]; out[i] = __fadd_rn(__fmul_rn(x.x, x.x), __fmul_rn(x.y, x.y)); }
    }
}

// Tee: one read, two writes (a second cudaMemcpy would read the input twice).  W = widest word all
// three pointers are aligned to (a ring window may start at any sample); the byte tail is < sizeof(W).
template <typename W>
__global__ void __launch_bounds__(256) tee_kernel(const W* __restrict__ in, W* __restrict__ o1, W* __restrict__ o2, size_t nw, size_t nbytes) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (; i + 3 * stride < nw; i += 4 * stride) {
        W v[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) v[k] = in[i + k * stride];
#pragma unroll
        for (int k = 0; k < 4; ++k) { o1[i + k * stride] = v[k]; o2[i + k * stride] = v[k]; }
    }
    for (; i < nw; i += stride) { const W v = in[i]; o1[i] = v; o2[i] = v; }
    if (blockIdx.x == 0 && threadIdx.x == 0)
        for (size_t b = nw * sizeof(W); b < nbytes; ++b) {
            const unsigned char v = reinterpret_cast<const unsigned char*>(in)[b];
            reinterpret_cast<unsigned char*>(o1)[b] = v;
            reinterpret_cast<unsigned char*>(o2)[b] = v;
        }
}

// ------------------------------------------------------------------ IqBalance -----
constexpr int IQ_L = 16;                 // samples per thread
constexpr int IQ_NT = 256;
constexpr int IQ_TILE = IQ_L * IQ_NT;    // 4096 samples per CTA
constexpr int IQ_PITCH = IQ_L + 1;       // one pad element per thread run: conflict-free 64-bit strided reads

struct Aff { float a; float2 b; };       // m -> a*m + b
// run `p` first, then `q`
__device__ __forceinline__ Aff aff_then(Aff p, Aff q) {
    return Aff{p.a * q.a, make_float2(fmaf(p.b.x, q.a, q.b.x), fmaf(p.b.y, q.a, q.b.y))};
}
__device__ __forceinline__ Aff aff_shfl_up(Aff v, int d) {
    return Aff{__shfl_up_sync(0xffffffffu, v.a, d), make_float2(__shfl_up_sync(0xffffffffu, v.b.x, d), __shfl_up_sync(0xffffffffu, v.b.y, d))};
}
// Inclusive scan over the CTA's threads (thread order); returns the EXCLUSIVE prefix of this thread
// and leaves the CTA total in *total (valid for every thread).  NW = warps per CTA (<= 32).
template <int NW>
__device__ __forceinline__ Aff cta_affine_scan(Aff mine, Aff* s_warp, Aff* total) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    Aff inc = mine;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const Aff up = aff_shfl_up(inc, d);
        if (lane >= d) inc = aff_then(up, inc);
    }
    if (lane == 31) s_warp[wid] = inc;
    __syncthreads();
    if (wid == 0) {
        Aff w = lane < NW ? s_warp[lane] : Aff{1.f, make_float2(0.f, 0.f)};
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const Aff up = aff_shfl_up(w, d);
            if (lane >= d) w = aff_then(up, w);
        }
        if (lane < NW) s_warp[lane] = w;                  // inclusive over warps
    }
    __syncthreads();
    Aff ex = aff_shfl_up(inc, 1);                          // exclusive within the warp
    if (lane == 0) ex = Aff{1.f, make_float2(0.f, 0.f)};
    if (wid > 0) ex = aff_then(s_warp[wid - 1], ex);
    *total = s_warp[NW - 1];
    return ex;
}

// APPLY = false: write the tile's aggregate B (A is the constant (1-a)^4096) to agg[tile].
// APPLY = true : read the mean entering the tile from carry[tile], write x - mean, and the last tile
//                stores the final mean to *mean_out (state for the next call).
template <bool APPLY>
__global__ void __launch_bounds__(IQ_NT) iq_tile_kernel(const float2* __restrict__ in, size_t n, float alpha, float oma,
                                                        float2* __restrict__ agg, const float2* carry,
                                                        float2* __restrict__ out, float2* mean_out) {
    __shared__ float2 s_x[IQ_NT * IQ_PITCH];
    __shared__ Aff s_warp[IQ_NT / 32];
    const int t = threadIdx.x;
    const size_t base = (size_t)blockIdx.x * IQ_TILE;
    const int cnt = (int)min((size_t)IQ_TILE, n - base);                 // valid samples in this tile
#pragma unroll
    for (int k = 0; k < IQ_L; ++k) {
        const int e = t + k * IQ_NT;
        s_x[e + e / IQ_L] = e < cnt ? in[base + e] : make_float2(0.f, 0.f);
    }
    __syncthreads();
    const float2* mx = s_x + t * IQ_PITCH;
    const int mycnt = max(0, min(IQ_L, cnt - t * IQ_L));
    float2 x[IQ_L];
#pragma unroll
    for (int k = 0; k < IQ_L; ++k) x[k] = mx[k];
    Aff mine{1.f, make_float2(0.f, 0.f)};
#pragma unroll
    for (int k = 0; k < IQ_L; ++k)
        if (k < mycnt) {                                                 // mean = mean*(1-a) + x*a
            mine.a *= oma;
            mine.b = make_float2(fmaf(mine.b.x, oma, x[k].x * alpha), fmaf(mine.b.y, oma, x[k].y * alpha));
        }
    Aff total;
    const Aff ex = cta_affine_scan<IQ_NT / 32>(mine, s_warp, &total);
    if constexpr (!APPLY) {
        if (t == 0) agg[blockIdx.x] = total.b;
    } else {
        const float2 m0 = carry[blockIdx.x];
        float2 m = make_float2(fmaf(m0.x, ex.a, ex.b.x), fmaf(m0.y, ex.a, ex.b.y));    // mean entering this thread's run
        __syncthreads();                                                 // every thread holds its x[] in registers
        float2* my = s_x + t * IQ_PITCH;
#pragma unroll
        for (int k = 0; k < IQ_L; ++k)
            if (k < mycnt) {
                m = make_float2(fmaf(m.x, oma, x[k].x * alpha), fmaf(m.y, oma, x[k].y * alpha));
                my[k] = make_float2(x[k].x - m.x, x[k].y - m.y);
            }
        if (base + IQ_TILE >= n && mycnt > 0 && t * IQ_L + mycnt == cnt) *mean_out = m;   // the thread holding the last sample
        __syncthreads();
#pragma unroll
        for (int k = 0; k < IQ_L; ++k) {
            const int e = t + k * IQ_NT;
            if (e < cnt) out[base + e] = s_x[e + e / IQ_L];
        }
    }
}

// carry[i] = mean entering tile i, i <= nagg: carry[0] = *mean_in, carry[i+1] = A*carry[i] + agg[i]
// (A = tile_a = (1-a)^4096); agg holds the nagg = ntiles - 1 full tiles before the last one.
__global__ void __launch_bounds__(1024) iq_carry_kernel(const float2* __restrict__ agg, float2* __restrict__ carry, size_t nagg,
                                                         float tile_a, const float2* mean_in) {
    __shared__ Aff s_warp[32];
    const int t = threadIdx.x;
    const size_t per = (nagg + 1023) / 1024;
    const size_t lo = min(nagg, (size_t)t * per), hi = min(nagg, lo + per);
    Aff mine{1.f, make_float2(0.f, 0.f)};
    for (size_t i = lo; i < hi; ++i) mine = aff_then(mine, Aff{tile_a, agg[i]});
    Aff total;
    const Aff ex = cta_affine_scan<32>(mine, s_warp, &total);
    const float2 m0 = *mean_in;
    float2 m = make_float2(fmaf(m0.x, ex.a, ex.b.x), fmaf(m0.y, ex.a, ex.b.y));
    for (size_t i = lo; i < hi; ++i) {
        carry[i] = m;
        const float2 b = agg[i];
        m = make_float2(fmaf(m.x, tile_a, b.x), fmaf(m.y, tile_a, b.y));
    }
    if (lo < hi && hi == nagg) carry[nagg] = m;
}

inline unsigned stream_grid(int device, size_t items, int per_cta) {
    const size_t want = (items + per_cta - 1) / per_cta;
    return (unsigned)std::max<size_t>(1, std::min<size_t>(want, (size_t)sm_count(device) * 8));
}

template <int OP>
int run_map(int device, const float* in, size_t nf, float2 c, float* out, void* stream) {
    if (nf == 0) return RRC_OK;
    if (!in || !out) return fail(RRC_ERR_INVALID, "in/out is NULL");
    RRC_CUDA(cudaSetDevice(device));
    const bool aligned = ((reinterpret_cast<uintptr_t>(in) | reinterpret_cast<uintptr_t>(out)) & 15) == 0;
    if (aligned) map_kernel<OP><<<stream_grid(device, nf / 4 + 1, 256 * 4), 256, 0, as_stream(stream)>>>(in, out, nf, c);
    else map_scalar_kernel<OP><<<stream_grid(device, nf, 256), 256, 0, as_stream(stream)>>>(in, out, nf, c);
    RRC_CHECK_LAUNCH();
    count_launch();
    return RRC_OK;
}

}  // namespace rrc

using namespace rrc;

struct rrc_iq_balance {
    int device = 0;
    float alpha = 0.f, oma = 1.f, tile_a = 1.f;
    float2* mean = nullptr;          // device: carried mean
    float2* scratch = nullptr;       // agg[ntiles] ++ carry[ntiles]
    size_t scratch_tiles = 0;
};

extern "C" {

int rrc_multiply_const_f32_run(int device, const float* in_dev, size_t n, float val, float* out_dev, void* stream) {
    return run_map<MAP_MUL_F32>(device, in_dev, n, make_float2(val, 0.f), out_dev, stream);
}
int rrc_multiply_const_c32_run(int device, const float* in_dev, size_t n, float val_re, float val_im, float* out_dev, void* stream) {
    return run_map<MAP_MUL_C32>(device, in_dev, 2 * n, make_float2(val_re, val_im), out_dev, stream);
}
int rrc_add_const_f32_run(int device, const float* in_dev, size_t n, float val, float* out_dev, void* stream) {
    return run_map<MAP_ADD_F32>(device, in_dev, n, make_float2(val, 0.f), out_dev, stream);
}
int rrc_add_const_c32_run(int device, const float* in_dev, size_t n, float val_re, float val_im, float* out_dev, void* stream) {
    return run_map<MAP_ADD_C32>(device, in_dev, 2 * n, make_float2(val_re, val_im), out_dev, stream);
}

int rrc_complex_to_mag2_run(int device, const float* in_dev_c32, size_t n, float* out_dev, void* stream) {
    if (n == 0) return RRC_OK;
    if (!in_dev_c32 || !out_dev) return fail(RRC_ERR_INVALID, "in/out is NULL");
    RRC_CUDA(cudaSetDevice(device));
    const int aligned = ((reinterpret_cast<uintptr_t>(in_dev_c32) | reinterpret_cast<uintptr_t>(out_dev)) & 15) == 0;
    mag2_kernel<<<stream_grid(device, aligned ? n / 4 + 1 : n, 256), 256, 0, as_stream(stream)>>>(
        reinterpret_cast<const float2*>(in_dev_c32), out_dev, n, aligned);
    RRC_CHECK_LAUNCH();
    count_launch();
    return RRC_OK;
}

int rrc_tee_run(int device, const void* in_dev, size_t nbytes, void* out1_dev, void* out2_dev, void* stream) {
    if (nbytes == 0) return RRC_OK;
    if (!in_dev || !out1_dev || !out2_dev) return fail(RRC_ERR_INVALID, "in/out is NULL");
    RRC_CUDA(cudaSetDevice(device));
    const uintptr_t bits = reinterpret_cast<uintptr_t>(in_dev) | reinterpret_cast<uintptr_t>(out1_dev) | reinterpret_cast<uintptr_t>(out2_dev);
    cudaStream_t st = as_stream(stream);
    if ((bits & 15) == 0)
        tee_kernel<uint4><<<stream_grid(device, nbytes / 16 + 1, 1024), 256, 0, st>>>((const uint4*)in_dev, (uint4*)out1_dev, (uint4*)out2_dev, nbytes / 16, nbytes);
    else if ((bits & 7) == 0)
        tee_kernel<uint2><<<stream_grid(device, nbytes / 8 + 1, 1024), 256, 0, st>>>((const uint2*)in_dev, (uint2*)out1_dev, (uint2*)out2_dev, nbytes / 8, nbytes);
    else if ((bits & 3) == 0)
        tee_kernel<unsigned int><<<stream_grid(device, nbytes / 4 + 1, 1024), 256, 0, st>>>((const unsigned int*)in_dev, (unsigned int*)out1_dev, (unsigned int*)out2_dev, nbytes / 4, nbytes);
    else
        tee_kernel<unsigned char><<<stream_grid(device, nbytes, 1024), 256, 0, st>>>((const unsigned char*)in_dev, (unsigned char*)out1_dev, (unsigned char*)out2_dev, nbytes, nbytes);
    RRC_CHECK_LAUNCH();
    count_launch();
    return RRC_OK;
}

int rrc_iq_balance_alpha_from_tau(unsigned sample_rate, double tau_seconds, float* alpha) {   // src/iq_balance.rs:41-57
    if (!alpha) return fail(RRC_ERR_INVALID, "alpha is NULL");
    const double fs = (double)std::max(sample_rate, 1u);
    const double tau = (std::isfinite(tau_seconds) && tau_seconds > 0.0) ? tau_seconds : 0.5;
    const double a = 1.0 - std::exp(-1.0 / (tau * fs));
    *alpha = (float)std::min(1.0, std::max(0.0, a));
    return RRC_OK;
}

int rrc_iq_balance_create(int device, float alpha, rrc_iq_balance_t** out) {                  // with_alpha, :62-73
    if (!out) return fail(RRC_ERR_INVALID, "out is NULL");
    *out = nullptr;
    if (alpha != alpha) return fail(RRC_ERR_INVALID, "IqBalance: alpha is NaN");
    auto* h = new rrc_iq_balance();
    h->device = device;
    h->alpha = std::min(1.0f, std::max(0.0f, alpha));
    h->oma = 1.0f - h->alpha;
    h->tile_a = (float)std::pow((double)h->oma, (double)IQ_TILE);
    cudaError_t e = cudaSetDevice(device);
    if (e == cudaSuccess) e = cudaMalloc((void**)&h->mean, sizeof(float2));
    if (e == cudaSuccess) e = zero_sync(h->mean, sizeof(float2));                          // mean: Complex::default()
    if (e == cudaSuccess) e = cudaStreamSynchronize(0);      // callers run on non-blocking streams, which do not wait for this fill
    if (e != cudaSuccess) {
        int s = fail(RRC_ERR_CUDA, "IqBalance create: %s", cudaGetErrorString(e));
        rrc_iq_balance_destroy(h);
        return s;
    }
    *out = h;
    return RRC_OK;
}
int rrc_iq_balance_destroy(rrc_iq_balance_t* h) {
    if (!h) return RRC_OK;
    cudaSetDevice(h->device);
    if (h->mean) cudaFree(h->mean);
    if (h->scratch) cudaFree(h->scratch);
    delete h;
    return RRC_OK;
}
int rrc_iq_balance_reset(rrc_iq_balance_t* h, void* stream) {
    if (!h) return fail(RRC_ERR_INVALID, "iq_balance handle is NULL");
    RRC_CUDA(cudaSetDevice(h->device));
    RRC_CUDA(cudaMemsetAsync(h->mean, 0, sizeof(float2), as_stream(stream)));
    return RRC_OK;
}
int rrc_iq_balance_mean(rrc_iq_balance_t* h, float* mean_re_im, void* stream) {
    if (!h || !mean_re_im) return fail(RRC_ERR_INVALID, "NULL argument");
    RRC_CUDA(cudaSetDevice(h->device));
    RRC_CUDA(cudaMemcpyAsync(mean_re_im, h->mean, sizeof(float2), cudaMemcpyDeviceToHost, as_stream(stream)));
    RRC_CUDA(cudaStreamSynchronize(as_stream(stream)));
    return RRC_OK;
}

int rrc_iq_balance_run(rrc_iq_balance_t* h, const float* in_dev_c32, size_t n, float* out_dev_c32, void* stream) {
    if (!h) return fail(RRC_ERR_INVALID, "iq_balance handle is NULL");
    if (n == 0) return RRC_OK;
    if (!in_dev_c32 || !out_dev_c32) return fail(RRC_ERR_INVALID, "in/out is NULL");
    RRC_CUDA(cudaSetDevice(h->device));
    cudaStream_t st = as_stream(stream);
    const float2* in = reinterpret_cast<const float2*>(in_dev_c32);
    float2* out = reinterpret_cast<float2*>(out_dev_c32);
    const size_t ntiles = (n + IQ_TILE - 1) / IQ_TILE;
    if (ntiles > 0x7fffffffu) return fail(RRC_ERR_INVALID, "IqBalance: n too large for one launch");
    if (ntiles == 1) {
        iq_tile_kernel<true><<<1, IQ_NT, 0, st>>>(in, n, h->alpha, h->oma, nullptr, h->mean, out, h->mean);
        RRC_CHECK_LAUNCH();
        count_launch();
        return RRC_OK;
    }
    if (ntiles > h->scratch_tiles) {
        if (h->scratch) { RRC_CUDA(cudaStreamSynchronize(st)); RRC_CUDA(cudaFree(h->scratch)); h->scratch = nullptr; h->scratch_tiles = 0; }
        RRC_CUDA(cudaMalloc((void**)&h->scratch, 2 * ntiles * sizeof(float2)));
        h->scratch_tiles = ntiles;
    }
    float2* agg = h->scratch;
    float2* carry = h->scratch + h->scratch_tiles;
    iq_tile_kernel<false><<<(unsigned)(ntiles - 1), IQ_NT, 0, st>>>(in, n, h->alpha, h->oma, agg, nullptr, nullptr, nullptr);
    RRC_CHECK_LAUNCH();
    iq_carry_kernel<<<1, 1024, 0, st>>>(agg, carry, ntiles - 1, h->tile_a, h->mean);
    RRC_CHECK_LAUNCH();
    iq_tile_kernel<true><<<(unsigned)ntiles, IQ_NT, 0, st>>>(in, n, h->alpha, h->oma, nullptr, carry, out, h->mean);
    RRC_CHECK_LAUNCH();
    count_launch(3);
    return RRC_OK;
}

}  // extern "C"
