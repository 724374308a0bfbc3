// fftstream.cu — batched forward FFT of fixed-size frames (FftStream / Fft) on sm_100a.
//
// Replaces the compute of FftStream::work (rustradio src/fft_stream.rs:71-117: every `size` samples
// are replaced by their unnormalised forward DFT, rustfft `process`) and of Fft::process_one
// (src/fft.rs:31-35); SURVEY 8f rank 2.  rustfft is third-party and not vendored (Cargo.lock:2299),
// so like the FftFilter the bit pattern is unpinned and parity is rel-RMS against the f64 DFT.
//
// One kernel template per log2(size): a Stockham autosort FFT held in shared memory, built from the
// same compile-time in-register DFTs (fft_regs.cuh, radix <= 32) as the FftFilter kernel.
//   size = R_0 * R_1 * ... (1..3 passes, bits split evenly); T = size / Rmax threads per frame;
//   pass p (Ns = R_0...R_{p-1}): task j < size/R_p reads x[j + i*size/R_p], i < R_p, multiplies by
//   W_{Ns R_p}^{i (j mod Ns)} (one table, W_size^m, read through L1/L2), does the radix-R_p DFT in
//   registers and writes X[(j div Ns) Ns R_p + (j mod Ns) + i Ns].
// The first pass reads global memory, the last writes it (both coalesced, natural order in and out);
// the exchange buffer index is skewed by a >> 5 (conflict-free for the stride-R_p stores).
// HBM bound by design: 16 B per sample.
#include <algorithm>
#include <cmath>
#include <vector>

#include "common.cuh"
#include "fft_regs.cuh"
#include "pipeline.cuh"

namespace rrc {

using namespace rrc::fftr;

__host__ __device__ constexpr int fs_npass(int k) { return k <= 5 ? 1 : k <= 10 ? 2 : 3; }
// bits of pass p: as even as possible, larger radices first
__host__ __device__ constexpr int fs_bits(int k, int p) {
    const int np = fs_npass(k), base = k / np, extra = k % np;
    return base + (p < extra ? 1 : 0);
}
__host__ __device__ constexpr int fs_rmax_bits(int k) { return fs_bits(k, 0); }
__host__ __device__ constexpr int fs_skew(int a) { return a + (a >> 5); }

struct FftStreamArgs {
    const float2* in;
    float2* out;
    const float2* tw;        // W_size^m = exp(-2 pi i m / size), m < size
    long long nframes;
};

// First or last pass (the middle pass of a 3-pass plan is written out in the kernel: it exchanges in
// place and needs a barrier between its reads and writes).
template <int K, int P, int RB>
__device__ __forceinline__ void fs_pass(int t, int T, const float2* __restrict__ gin, float2* __restrict__ gout,
                                        float2* sm, const float2* __restrict__ tw, int ns_bits) {
    constexpr int N = 1 << K, R = 1 << RB, NP = fs_npass(K);
    constexpr bool FIRST = P == 0, LAST = P == NP - 1;
    const int Ns = 1 << ns_bits;
    for (int j = t; j < N / R; j += T) {
        float2 v[R];
        const int jm = j & (Ns - 1);
#pragma unroll
        for (int i = 0; i < R; ++i) {
            const int a = j + i * (N / R);
            float2 x = FIRST ? gin[a] : sm[fs_skew(a)];
            if (!FIRST && i > 0) x = cmul(x, tw[(i * jm) << (K - ns_bits - RB)]);
            v[bitrev(i, RB)] = x;
        }
        dit<R, +1>(v);
        const int base = ((j >> ns_bits) << (ns_bits + RB)) + jm;
#pragma unroll
        for (int i = 0; i < R; ++i) {
            const int a = base + (i << ns_bits);
            if (LAST) gout[a] = v[i]; else sm[fs_skew(a)] = v[i];
        }
    }
}

// One CTA processes FPC frames at a time (FPC * T threads), grid-stride over frame groups.
template <int K>
__global__ void __launch_bounds__(K == 14 ? 512 : 256) fftstream_kernel(const FftStreamArgs a, int fpc) {
    constexpr int N = 1 << K, NP = fs_npass(K), RMB = fs_rmax_bits(K), T = N >> RMB;
    extern __shared__ __align__(16) float2 sm_all[];
    const int f = threadIdx.x / T, t = threadIdx.x - f * T;
    float2* sm = sm_all + (size_t)f * fs_skew(N);
    for (long long g = (long long)blockIdx.x * fpc; g < a.nframes; g += (long long)gridDim.x * fpc) {
        const long long frame = g + f;
        const bool live = frame < a.nframes;
        const float2* gin = a.in + frame * N;
        float2* gout = a.out + frame * N;
        if constexpr (NP == 1) {
            if (live) fs_pass<K, 0, K>(t, T, gin, gout, sm, a.tw, 0);
        } else if constexpr (NP == 2) {
            constexpr int B0 = fs_bits(K, 0), B1 = fs_bits(K, 1);
            if (live) fs_pass<K, 0, B0>(t, T, gin, gout, sm, a.tw, 0);
            __syncthreads();
            if (live) fs_pass<K, 1, B1>(t, T, gin, gout, sm, a.tw, B0);
            __syncthreads();                                // buffer reused by the next frame group
        } else {
            constexpr int B0 = fs_bits(K, 0), B1 = fs_bits(K, 1), B2 = fs_bits(K, 2);
            if (live) fs_pass<K, 0, B0>(t, T, gin, gout, sm, a.tw, 0);
            __syncthreads();
            // middle pass: read everything into registers, barrier, then write (in place)
            {
                constexpr int R = 1 << B1, TASKS = (1 << RMB) / R;      // tasks per thread (1 or 2)
                float2 v[TASKS][R];
                if (live) {
#pragma unroll
                    for (int q = 0; q < TASKS; ++q) {
                        const int j = t + q * T, jm = j & ((1 << B0) - 1);
#pragma unroll
                        for (int i = 0; i < R; ++i) {
                            float2 x = sm[fs_skew(j + i * (N / R))];
                            if (i > 0) x = cmul(x, a.tw[(i * jm) << (K - B0 - B1)]);
                            v[q][bitrev(i, B1)] = x;
                        }
                        dit<R, +1>(v[q]);
                    }
                }
                __syncthreads();
                if (live) {
#pragma unroll
                    for (int q = 0; q < TASKS; ++q) {
                        const int j = t + q * T, jm = j & ((1 << B0) - 1);
                        const int base = ((j >> B0) << (B0 + B1)) + jm;
#pragma unroll
                        for (int i = 0; i < R; ++i) sm[fs_skew(base + (i << B0))] = v[q][i];
                    }
                }
            }
            __syncthreads();
            if (live) fs_pass<K, 2, B2>(t, T, gin, gout, sm, a.tw, B0 + B1);
            __syncthreads();
        }
    }
}

}  // namespace rrc

using namespace rrc;

struct rrc_fft {
    int device = 0;
    int k = 0;               // log2(size)
    size_t size = 0;
    float2* tw = nullptr;
    Pipe pipe;
};

namespace {

template <int K>
int launch_k(const rrc_fft* h, const FftStreamArgs& a, cudaStream_t st) {
    constexpr int N = 1 << K, T = N >> fs_rmax_bits(K);
    int fpc = std::max(1, 256 / T);
    fpc = (int)std::min<long long>(fpc, a.nframes);
    const size_t smem = fs_npass(K) == 1 ? 0 : (size_t)fpc * fs_skew(N) * sizeof(float2);
    auto kern = fftstream_kernel<K>;
    if (smem > 48 * 1024) RRC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const long long groups = (a.nframes + fpc - 1) / fpc;
    const int per_sm = smem ? std::max<int>(1, (int)std::min<size_t>(8, (200 * 1024) / smem)) : 8;
    const unsigned grid = (unsigned)std::min<long long>(groups, (long long)sm_count(h->device) * per_sm);
    kern<<<grid, fpc * T, smem, st>>>(a, fpc);
    RRC_CHECK_LAUNCH();
    count_launch();
    return RRC_OK;
}

int launch_fft(const rrc_fft* h, const FftStreamArgs& a, cudaStream_t st) {
    switch (h->k) {
        case 1: return launch_k<1>(h, a, st);   case 2: return launch_k<2>(h, a, st);
        case 3: return launch_k<3>(h, a, st);   case 4: return launch_k<4>(h, a, st);
        case 5: return launch_k<5>(h, a, st);   case 6: return launch_k<6>(h, a, st);
        case 7: return launch_k<7>(h, a, st);   case 8: return launch_k<8>(h, a, st);
        case 9: return launch_k<9>(h, a, st);   case 10: return launch_k<10>(h, a, st);
        case 11: return launch_k<11>(h, a, st); case 12: return launch_k<12>(h, a, st);
        case 13: return launch_k<13>(h, a, st); case 14: return launch_k<14>(h, a, st);
    }
    return fail(RRC_ERR_UNSUPPORTED, "FFT size 2^%d not supported", h->k);
}

}  // namespace

extern "C" {

int rrc_fft_c32_create(int device, size_t size, rrc_fft_t** out) {
    if (!out) return fail(RRC_ERR_INVALID, "out is NULL");
    *out = nullptr;
    if (size == 0) return fail(RRC_ERR_INVALID, "FFT size must be nonzero (src/fft_stream.rs:42, src/fft.rs:27)");
    if (size & (size - 1)) return fail(RRC_ERR_UNSUPPORTED, "FFT size %zu: only powers of two are implemented on the device", size);
    if (size > 16384) return fail(RRC_ERR_UNSUPPORTED, "FFT size %zu > 16384", size);
    RRC_CUDA(cudaSetDevice(device));
    auto* h = new rrc_fft();
    h->device = device; h->size = size;
    while (((size_t)1 << h->k) < size) ++h->k;
    if (size > 1) {
        std::vector<float2> tw(size);
        for (size_t m = 0; m < size; ++m) {
            const double ang = -2.0 * M_PI * (double)m / (double)size;
            tw[m] = make_float2((float)std::cos(ang), (float)std::sin(ang));
        }
        cudaError_t e = cudaMalloc(&h->tw, size * sizeof(float2));
        if (e == cudaSuccess) e = upload_sync(h->tw, tw.data(), size * sizeof(float2));
        if (e != cudaSuccess) { rrc_fft_destroy(h); return fail(RRC_ERR_CUDA, "fft tables: %s", cudaGetErrorString(e)); }
    }
    *out = h;
    return RRC_OK;
}

int rrc_fft_destroy(rrc_fft_t* h) {
    if (!h) return RRC_OK;
    cudaSetDevice(h->device);
    cudaFree(h->tw);
    h->pipe.destroy();
    delete h;
    return RRC_OK;
}

int rrc_fft_size(const rrc_fft_t* h, size_t* size) {
    if (!h || !size) return fail(RRC_ERR_INVALID, "NULL argument");
    *size = h->size;
    return RRC_OK;
}

int rrc_fft_run(rrc_fft_t* h, const float* in_dev, size_t nframes, float* out_dev, void* stream) {
    if (!h) return fail(RRC_ERR_INVALID, "fft handle is NULL");
    if (nframes == 0) return RRC_OK;
    if (!in_dev || !out_dev) return fail(RRC_ERR_INVALID, "in/out is NULL");
    RRC_CUDA(cudaSetDevice(h->device));
    if (h->size == 1) {
        if (in_dev != out_dev) RRC_CUDA(cudaMemcpyAsync(out_dev, in_dev, nframes * sizeof(float2), cudaMemcpyDeviceToDevice, as_stream(stream)));
        return RRC_OK;
    }
    FftStreamArgs a{reinterpret_cast<const float2*>(in_dev), reinterpret_cast<float2*>(out_dev), h->tw, (long long)nframes};
    return launch_fft(h, a, as_stream(stream));
}

int rrc_fftstream_plan(size_t size, size_t in_len, size_t out_free, size_t* len, size_t* wait_need, int* wait_on_output) {
    if (!len || !wait_need || !wait_on_output) return fail(RRC_ERR_INVALID, "NULL argument");
    if (size == 0) return fail(RRC_ERR_INVALID, "FFT size must be nonzero");
    *len = 0; *wait_need = 0; *wait_on_output = 0;
    if (in_len < size) { *wait_need = size; *wait_on_output = 0; return RRC_OK; }      // src/fft_stream.rs:75-77
    if (out_free < size) { *wait_need = size; *wait_on_output = 1; return RRC_OK; }    // :80-82
    const size_t m = std::min(in_len, out_free);
    *len = m - m % size;                                                               // :83-84
    return RRC_OK;
}

int rrc_fft_run_host(rrc_fft_t* h, const float* in_host, size_t n_in, float* out_host, size_t* n_out) {
    if (!h) return fail(RRC_ERR_INVALID, "fft handle is NULL");
    const size_t total = (n_in / h->size) * h->size;
    if (n_out) *n_out = total;
    if (total == 0) return RRC_OK;
    if (!in_host || !out_host) return fail(RRC_ERR_INVALID, "in/out is NULL");
    RRC_TRY(h->pipe.init(h->device));
    const size_t chunk = std::max(h->size, PIPE_CHUNK_SAMPLES / h->size * h->size);
    RRC_TRY(h->pipe.reserve(std::min(chunk, total) * sizeof(float2), std::min(chunk, total) * sizeof(float2)));
    int i = 0;
    for (size_t off = 0; off < total; off += chunk, ++i) {
        const size_t n = std::min(chunk, total - off);
        RRC_TRY(h->pipe.stage_in(i, in_host + 2 * off, n * sizeof(float2)));
        RRC_TRY(rrc_fft_run(h, (const float*)h->pipe.d_in[i & 1], n / h->size, (float*)h->pipe.d_out[i & 1], h->pipe.s_comp));
        RRC_TRY(h->pipe.drain_out(i, out_host + 2 * off, n * sizeof(float2)));
    }
    return h->pipe.finish();
}

}  // extern "C"
