// hilbert.cu — Hilbert transformer on sm_100a.
//
// Replaces rustradio's Hilbert::work compute (src/hilbert.rs:72-128) and the tap design
// fir::hilbert (src/fir.rs:660-680) / WindowType::make_window (src/window.rs:63-185).
//
// Stream semantics of the reference (chunking independent, see hilbert.rs:86-125): the block keeps
// `ntaps` samples of history, initially zeros.  With z = [0]*T ++ x (T = ntaps, odd),
//     out[i] = Complex( z[i + T/2],  sum_{j<T} z[i + j] * h'[j] ),   h'[j] = taps[T-1-j],  i < N
// so N input samples give N outputs, the real part is the input delayed by (T+1)/2 samples and the
// imaginary part is the FIR of the zero-prefixed stream (one sample later than a centred filter).
//
// Kernel: the register-blocked sliding-window FIR of fir.cu specialised to f32 in / c32 out with a
// two-source input (carried history, then this call's samples): a CTA of 128 threads stages the
// span of its 1024 outputs in shared memory (one pad word per 8 so the thread-strided window reads
// are conflict free), every thread slides an 8-wide register window over 8-tap chunks (64 FMA per
// 8 window loads + 2 broadcast tap loads) and writes its 8 (re, im) pairs with four 128-bit stores.
// Bytes 4 + 8 per sample.
//
// Half-band structure: fir::hilbert only sets taps at ODD distances from the centre (src/fir.rs:667-676),
// so h'[j] == 0 unless j = par (mod 2), par = (T/2 + 1) & 1.  hilbert_half_kernel exploits that when it
// holds exactly for the given taps: out[i] = sum_m g[m] * z[i + par + 2m], g[m] = h'[par + 2m] — for the
// outputs of one parity a NON-decimated FIR over every second sample.  A thread owns 16 consecutive
// outputs as two interleaved sets of 8 (even / odd i) that share the tap registers: half the FMAs and
// half the window loads of the dense form (T = 65: 32 instead of 72 FMA per output).
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <vector>

#include "common.cuh"

namespace rrc {

constexpr int HIL_R = 8;
constexpr int HIL_NT = 128;
constexpr int HIL_MAX_TAPS = 8191;

struct HilbertArgs {
    const float* in;        // this call's samples x[0..n)
    const float* hist;      // T carried samples: z[0..T)
    float2* out;            // n outputs
    const float* taps;      // h' (reversed), zero padded to nchunks * 8
    long long n;
    int T, nchunks, mid;    // mid = T / 2
};

__device__ __forceinline__ float hil_z(const HilbertArgs& a, long long m) {
    if (m < a.T) return a.hist[m];
    m -= a.T;
    return m < a.n ? a.in[m] : 0.f;
}

__global__ void __launch_bounds__(HIL_NT) hilbert_kernel(const HilbertArgs a) {
    extern __shared__ __align__(16) float hil_smem[];
    constexpr int R = HIL_R, S1 = R + 1, BT = HIL_NT * R;
    const int t = threadIdx.x;
    const int ntap_tab = a.nchunks * R;
    float* s_taps = hil_smem;
    float* s_tile = hil_smem + ((ntap_tab + 3) & ~3);
    const long long ob = (long long)blockIdx.x * BT;           // first output (= first z index) of the tile
    const int L = (HIL_NT + a.nchunks) * R;                    // staged z elements

    for (int i = t; i < ntap_tab; i += HIL_NT) s_taps[i] = a.taps[i];
    if (ob >= a.T && ob - a.T + L <= a.n) {                    // interior: straight from this call's input
        const float* src = a.in + (ob - a.T);
        for (int e = t; e < L; e += HIL_NT) s_tile[e + e / R] = src[e];
    } else {
        for (int e = t; e < L; e += HIL_NT) s_tile[e + e / R] = hil_z(a, ob + e);
    }
    __syncthreads();

    float acc[R];
#pragma unroll
    for (int r = 0; r < R; ++r) acc[r] = 0.f;
    const float* bp = s_tile + t * S1;
    const float* tp = s_taps;
    float w[2 * R - 1];
#pragma unroll
    for (int u = 0; u < R - 1; ++u) w[u] = bp[u];
    for (int c = 0; c < a.nchunks; ++c) {
#pragma unroll
        for (int u = R - 1; u < 2 * R - 1; ++u) w[u] = bp[u + (u >= R ? 1 : 0)];
        const float4 h0 = *reinterpret_cast<const float4*>(tp), h1 = *reinterpret_cast<const float4*>(tp + 4);
        const float h[R] = {h0.x, h0.y, h0.z, h0.w, h1.x, h1.y, h1.z, h1.w};
#pragma unroll
        for (int k = 0; k < R; ++k)
#pragma unroll
            for (int r = 0; r < R; ++r) acc[r] = fmaf(h[k], w[r + k], acc[r]);
#pragma unroll
        for (int u = 0; u < R - 1; ++u) w[u] = w[u + R];
        bp += S1;
        tp += R;
    }

    const long long gi0 = ob + (long long)t * R;
    float re[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
        const int e = t * R + r + a.mid;                       // z[i + T/2] is inside the staged span
        re[r] = s_tile[e + e / R];
    }
    float2* dst = a.out + gi0;
    if (gi0 + R <= a.n) {                                      // cudaMalloc'd / ring windows: 16-byte aligned when gi0 is even
        if ((reinterpret_cast<unsigned long long>(dst) & 15ull) == 0) {
#pragma unroll
            for (int r = 0; r < R; r += 2)
                *reinterpret_cast<float4*>(dst + r) = make_float4(re[r], acc[r], re[r + 1], acc[r + 1]);
        } else {
#pragma unroll
            for (int r = 0; r < R; ++r) dst[r] = make_float2(re[r], acc[r]);
        }
    } else {
#pragma unroll
        for (int r = 0; r < R; ++r)
            if (gi0 + r < a.n) dst[r] = make_float2(re[r], acc[r]);
    }
}

// Half-band form (see the file header).  Tile element e <-> z[ob + par + e]; padded index e + e/16, so a
// thread's 16-element runs are 17 words apart (conflict free) and every window offset is an immediate:
// set s in {0,1}, window slot u in [0,15): element (t+c)*16 + s + 2u  ->  (t+c)*17 + s + 2u + (u >= 8).
constexpr int HILH_NT = 128;
constexpr int HILH_OUT = 16;                                   // outputs per thread
__global__ void __launch_bounds__(HILH_NT) hilbert_half_kernel(const HilbertArgs a, int par) {
    extern __shared__ __align__(16) float hil_smem[];
    constexpr int R = 8, P = HILH_OUT + 1, BT = HILH_NT * HILH_OUT;
    const int t = threadIdx.x;
    const int ntap_tab = a.nchunks * R;                        // a.nchunks: chunks of 8 taps of g
    float* s_taps = hil_smem;
    float* s_tile = hil_smem + ((ntap_tab + 3) & ~3);
    const long long ob = (long long)blockIdx.x * BT;
    const int L = (HILH_NT + a.nchunks + 1) * HILH_OUT;
    const long long zb = ob + par;                             // z index of tile element 0

    for (int i = t; i < ntap_tab; i += HILH_NT) s_taps[i] = a.taps[i];
    if (zb >= a.T && zb - a.T + L <= a.n) {
        const float* src = a.in + (zb - a.T);
        for (int e = t; e < L; e += HILH_NT) s_tile[e + (e >> 4)] = src[e];
    } else {
        for (int e = t; e < L; e += HILH_NT) s_tile[e + (e >> 4)] = hil_z(a, zb + e);
    }
    __syncthreads();

    float acc0[R], acc1[R];
#pragma unroll
    for (int r = 0; r < R; ++r) { acc0[r] = 0.f; acc1[r] = 0.f; }
    const float* bp = s_tile + t * P;
    const float* tp = s_taps;
    float w0[2 * R - 1], w1[2 * R - 1];
#pragma unroll
    for (int u = 0; u < R - 1; ++u) { w0[u] = bp[2 * u]; w1[u] = bp[2 * u + 1]; }
    for (int c = 0; c < a.nchunks; ++c) {
#pragma unroll
        for (int u = R - 1; u < 2 * R - 1; ++u) {
            w0[u] = bp[2 * u + (u >= R ? 1 : 0)];
            w1[u] = bp[2 * u + 1 + (u >= R ? 1 : 0)];
        }
        const float4 h0 = *reinterpret_cast<const float4*>(tp), h1 = *reinterpret_cast<const float4*>(tp + 4);
        const float h[R] = {h0.x, h0.y, h0.z, h0.w, h1.x, h1.y, h1.z, h1.w};
#pragma unroll
        for (int k = 0; k < R; ++k)
#pragma unroll
            for (int r = 0; r < R; ++r) {
                acc0[r] = fmaf(h[k], w0[r + k], acc0[r]);
                acc1[r] = fmaf(h[k], w1[r + k], acc1[r]);
            }
#pragma unroll
        for (int u = 0; u < R - 1; ++u) { w0[u] = w0[u + R]; w1[u] = w1[u + R]; }
        bp += P;
        tp += R;
    }

    const int e0 = t * HILH_OUT + a.mid - par;                 // tile element of z[i + T/2] for the thread's first output
    float re[HILH_OUT];
#pragma unroll
    for (int k = 0; k < HILH_OUT; ++k) re[k] = s_tile[(e0 + k) + ((e0 + k) >> 4)];
    // Coalesced stores through shared memory: a thread's 16 results are 128 contiguous bytes, so a direct
    // STG.128 per thread touches 32 different lines per warp instruction (32 L1 wavefronts each — measured:
    // the store pipe, not the FMAs, bounded the first version).  Staged with pitch 17 (conflict-free
    // 64-bit writes), read back as consecutive 16-byte pairs: 512 contiguous bytes per warp instruction.
    __syncthreads();                                           // everyone is done with the input tile
    float2* s_out = reinterpret_cast<float2*>(hil_smem);
#pragma unroll
    for (int r = 0; r < R; ++r) {
        s_out[t * P + 2 * r] = make_float2(re[2 * r], acc0[r]);
        s_out[t * P + 2 * r + 1] = make_float2(re[2 * r + 1], acc1[r]);
    }
    __syncthreads();
    float2* dst = a.out + ob;
    const bool vec = (reinterpret_cast<unsigned long long>(dst) & 15ull) == 0;
#pragma unroll
    for (int k = 0; k < R; ++k) {
        const int o = 2 * (t + k * HILH_NT);                   // even output index inside the tile
        const float2 y0 = s_out[o + (o >> 4)], y1 = s_out[o + 1 + (o >> 4)];
        if (vec && ob + o + 1 < a.n) {
            *reinterpret_cast<float4*>(dst + o) = make_float4(y0.x, y0.y, y1.x, y1.y);
        } else {
            if (ob + o < a.n) dst[o] = y0;
            if (ob + o + 1 < a.n) dst[o + 1] = y1;
        }
    }
}

// history for the next call: z[n .. n + T)  (src/hilbert.rs:123)
__global__ void hilbert_hist_kernel(const HilbertArgs a, float* hist_next) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < a.T) hist_next[k] = hil_z(a, a.n + k);
}

}  // namespace rrc

using namespace rrc;

struct rrc_hilbert {
    int device = 0;
    size_t ntaps = 0;
    int nchunks = 0;
    float* taps_dev = nullptr;
    float* hist[2] = {nullptr, nullptr};
    int cur = 0;
    size_t smem = 0;
    // half-band form (hilbert_half_kernel): every tap h'[j] with j != par (mod 2) is exactly zero
    bool half = false;
    int par = 0, half_chunks = 0;
    float* half_taps_dev = nullptr;
    size_t half_smem = 0;
};

namespace {

// WindowType::make_window (src/window.rs:63-185), all arithmetic in f32 like the reference's `Float`.
void window_f32(int type, float parm, size_t n, float* w) {
    const float pi = (float)M_PI;
    if (n == 1) { w[0] = 1.0f; return; }
    for (size_t i = 0; i < n; ++i) {
        const float x = (float)i;
        switch (type) {
        case RRC_WINDOW_BLACKMAN: {
            const float m = (float)n, A = 0.16f;
            w[i] = (1.0f - A) / 2.0f - 0.5f * cosf(2.0f * pi * x / m) + (A / 2.0f) * cosf(4.0f * pi * x / m);
            break;
        }
        case RRC_WINDOW_BLACKMAN_HARRIS: {
            const float m = (float)n;
            w[i] = 0.35875f - 0.48829f * cosf(2.0f * pi * x / m) + 0.14128f * cosf(4.0f * pi * x / m) -
                   0.01168f * cosf(6.0f * pi * x / m);
            break;
        }
        default: {   // Hamming / HammingParm
            const float a0 = type == RRC_WINDOW_HAMMING_PARM ? parm : 25.0f / 46.0f;
            w[i] = a0 - (1.0f - a0) * cosf(2.0f * pi * x / (float)(n - 1));
        }
        }
    }
}

}  // namespace

extern "C" {

int rrc_make_window(int window_type, float parm, size_t ntaps, float* window_out) {
    if (!window_out && ntaps) return fail(RRC_ERR_INVALID, "window_out is NULL");
    if (window_type < RRC_WINDOW_HAMMING || window_type > RRC_WINDOW_HAMMING_PARM)
        return fail(RRC_ERR_INVALID, "unknown window type %d", window_type);
    if (ntaps) window_f32(window_type, parm, ntaps, window_out);
    return RRC_OK;
}

int rrc_hilbert_taps(const float* window, size_t ntaps, float* taps_out) {   // src/fir.rs:660-680
    if (!window || !taps_out) return fail(RRC_ERR_INVALID, "NULL argument");
    if (ntaps < 2) return fail(RRC_ERR_INVALID, "hilbert() needs a window of at least 2 taps (src/fir.rs:661-662)");
    const size_t mid = (ntaps - 1) / 2;
    float gain = 0.0f;
    std::fill(taps_out, taps_out + ntaps, 0.0f);
    for (size_t i = 1; i <= mid; ++i) {
        if (i & 1) {
            const float x = 1.0f / (float)i;
            taps_out[mid + i] = x * window[mid + i];
            taps_out[mid - i] = -x * window[mid - i];
            gain = taps_out[mid + i] - gain;
        }
    }
    gain = 1.0f / (2.0f * fabsf(gain));
    for (size_t i = 0; i < ntaps; ++i) taps_out[i] = gain * taps_out[i];
    return RRC_OK;
}

int rrc_hilbert_create(int device, const float* taps, size_t ntaps, rrc_hilbert_t** out) {
    if (!out) return fail(RRC_ERR_INVALID, "out is NULL");
    *out = nullptr;
    if (!taps || ntaps <= 1 || (ntaps & 1) == 0)
        return fail(RRC_ERR_INVALID, "hilbert filter len must be odd and greater than 1 (src/hilbert.rs:44-47)");
    if (ntaps > (size_t)HIL_MAX_TAPS) return fail(RRC_ERR_UNSUPPORTED, "Hilbert: at most %d taps", HIL_MAX_TAPS);
    auto* h = new rrc_hilbert();
    h->device = device; h->ntaps = ntaps;
    h->nchunks = (int)((ntaps + HIL_R - 1) / HIL_R);
    std::vector<float> rev((size_t)h->nchunks * HIL_R, 0.0f);
    for (size_t j = 0; j < ntaps; ++j) rev[j] = taps[ntaps - 1 - j];          // Fir::new reverses (src/fir.rs:156-162)
    const size_t tap_floats = (rev.size() + 3) & ~(size_t)3;
    h->smem = (tap_floats + (size_t)(HIL_NT + h->nchunks + 1) * (HIL_R + 1)) * sizeof(float);
    h->par = (int)((ntaps / 2 + 1) & 1);
    h->half = !getenv("RRC_HILBERT_DENSE");
    for (size_t j = 0; j < ntaps && h->half; ++j)
        if ((int)(j & 1) != h->par && rev[j] != 0.0f) h->half = false;
    std::vector<float> g;
    if (h->half) {
        for (size_t j = (size_t)h->par; j < ntaps; j += 2) g.push_back(rev[j]);
        h->half_chunks = (int)((g.size() + HIL_R - 1) / HIL_R);
        g.resize((size_t)h->half_chunks * HIL_R, 0.0f);
        h->half_smem = std::max((((g.size() + 3) & ~(size_t)3) + (size_t)(HILH_NT + h->half_chunks + 2) * (HILH_OUT + 1)) * sizeof(float),
                                (size_t)HILH_NT * (HILH_OUT + 1) * sizeof(float2));      // input tile, later the output staging
        if (h->half_smem > 200 * 1024) h->half = false;
    }
    auto bail = [&](cudaError_t e, const char* what) {
        int s = fail(RRC_ERR_CUDA, "%s failed: %s", what, cudaGetErrorString(e));
        rrc_hilbert_destroy(h);
        return s;
    };
    cudaError_t e;
    if ((e = cudaSetDevice(device)) != cudaSuccess) return bail(e, "cudaSetDevice");
    if ((e = cudaMalloc((void**)&h->taps_dev, rev.size() * sizeof(float))) != cudaSuccess) return bail(e, "cudaMalloc");
    if ((e = upload_sync(h->taps_dev, rev.data(), rev.size() * sizeof(float))) != cudaSuccess) return bail(e, "cudaMemcpy");
    if (h->half) {
        if ((e = cudaMalloc((void**)&h->half_taps_dev, g.size() * sizeof(float))) != cudaSuccess) return bail(e, "cudaMalloc");
        if ((e = upload_sync(h->half_taps_dev, g.data(), g.size() * sizeof(float))) != cudaSuccess) return bail(e, "cudaMemcpy");
        if (h->half_smem > 48 * 1024 &&
            (e = cudaFuncSetAttribute(hilbert_half_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->half_smem)) != cudaSuccess)
            return bail(e, "cudaFuncSetAttribute");
    }
    for (int i = 0; i < 2; ++i) {
        if ((e = cudaMalloc((void**)&h->hist[i], ntaps * sizeof(float))) != cudaSuccess) return bail(e, "cudaMalloc");
        if ((e = zero_sync(h->hist[i], ntaps * sizeof(float))) != cudaSuccess) return bail(e, "cudaMemset");   // history: vec![0.0; ntaps]
    }
    if ((e = cudaStreamSynchronize(0)) != cudaSuccess) return bail(e, "cudaStreamSynchronize");   // callers run on non-blocking streams
    if (h->smem > 48 * 1024 &&
        (e = cudaFuncSetAttribute(hilbert_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->smem)) != cudaSuccess)
        return bail(e, "cudaFuncSetAttribute");
    *out = h;
    return RRC_OK;
}

int rrc_hilbert_destroy(rrc_hilbert_t* h) {
    if (!h) return RRC_OK;
    cudaSetDevice(h->device);
    if (h->taps_dev) cudaFree(h->taps_dev);
    if (h->half_taps_dev) cudaFree(h->half_taps_dev);
    for (int i = 0; i < 2; ++i) if (h->hist[i]) cudaFree(h->hist[i]);
    delete h;
    return RRC_OK;
}

int rrc_hilbert_reset(rrc_hilbert_t* h, void* stream) {
    if (!h) return fail(RRC_ERR_INVALID, "hilbert handle is NULL");
    RRC_CUDA(cudaSetDevice(h->device));
    RRC_CUDA(cudaMemsetAsync(h->hist[h->cur], 0, h->ntaps * sizeof(float), as_stream(stream)));
    return RRC_OK;
}

int rrc_hilbert_run(rrc_hilbert_t* h, const float* in_dev, size_t n, float* out_dev_c32, void* stream) {
    if (!h) return fail(RRC_ERR_INVALID, "hilbert handle is NULL");
    if (n == 0) return RRC_OK;
    if (!in_dev || !out_dev_c32) return fail(RRC_ERR_INVALID, "in/out is NULL");
    RRC_CUDA(cudaSetDevice(h->device));
    cudaStream_t st = as_stream(stream);
    HilbertArgs a{};
    a.in = in_dev; a.hist = h->hist[h->cur]; a.out = reinterpret_cast<float2*>(out_dev_c32);
    a.taps = h->taps_dev; a.n = (long long)n; a.T = (int)h->ntaps; a.nchunks = h->nchunks; a.mid = (int)(h->ntaps / 2);
    const size_t bt = h->half ? (size_t)HILH_NT * HILH_OUT : (size_t)HIL_NT * HIL_R;
    const size_t tiles = (n + bt - 1) / bt;
    if (tiles > 0x7fffffffu) return fail(RRC_ERR_INVALID, "Hilbert: n too large for one launch");
    if (h->half) {
        HilbertArgs b = a;
        b.taps = h->half_taps_dev; b.nchunks = h->half_chunks;
        hilbert_half_kernel<<<(unsigned)tiles, HILH_NT, h->half_smem, st>>>(b, h->par);
    } else {
        hilbert_kernel<<<(unsigned)tiles, HIL_NT, h->smem, st>>>(a);
    }
    RRC_CHECK_LAUNCH();
    hilbert_hist_kernel<<<(unsigned)((h->ntaps + 255) / 256), 256, 0, st>>>(a, h->hist[h->cur ^ 1]);
    RRC_CHECK_LAUNCH();
    count_launch(2);
    h->cur ^= 1;
    return RRC_OK;
}

}  // extern "C"
