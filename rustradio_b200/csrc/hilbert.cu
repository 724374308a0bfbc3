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
// Bytes 4 + 8 per sample; FMA T per sample (every other Hilbert tap is zero — a polyphase-by-2 form
// would halve that; not done).
#include <algorithm>
#include <cmath>
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
    auto bail = [&](cudaError_t e, const char* what) {
        int s = fail(RRC_ERR_CUDA, "%s failed: %s", what, cudaGetErrorString(e));
        rrc_hilbert_destroy(h);
        return s;
    };
    cudaError_t e;
    if ((e = cudaSetDevice(device)) != cudaSuccess) return bail(e, "cudaSetDevice");
    if ((e = cudaMalloc((void**)&h->taps_dev, rev.size() * sizeof(float))) != cudaSuccess) return bail(e, "cudaMalloc");
    if ((e = cudaMemcpy(h->taps_dev, rev.data(), rev.size() * sizeof(float), cudaMemcpyHostToDevice)) != cudaSuccess) return bail(e, "cudaMemcpy");
    for (int i = 0; i < 2; ++i) {
        if ((e = cudaMalloc((void**)&h->hist[i], ntaps * sizeof(float))) != cudaSuccess) return bail(e, "cudaMalloc");
        if ((e = cudaMemset(h->hist[i], 0, ntaps * sizeof(float))) != cudaSuccess) return bail(e, "cudaMemset");   // history: vec![0.0; ntaps]
    }
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
    const size_t bt = (size_t)HIL_NT * HIL_R;
    const size_t tiles = (n + bt - 1) / bt;
    if (tiles > 0x7fffffffu) return fail(RRC_ERR_INVALID, "Hilbert: n too large for one launch");
    hilbert_kernel<<<(unsigned)tiles, HIL_NT, h->smem, st>>>(a);
    RRC_CHECK_LAUNCH();
    hilbert_hist_kernel<<<(unsigned)((h->ntaps + 255) / 256), 256, 0, st>>>(a, h->hist[h->cur ^ 1]);
    RRC_CHECK_LAUNCH();
    count_launch(2);
    h->cur ^= 1;
    return RRC_OK;
}

}  // extern "C"
