// resample.cu — RationalResampler on sm_100a: a pure index gather.
//
// Replaces RationalResampler::new / work (rustradio src/rational_resampler.rs:
// 125-206).  The reference block has no filter: per input sample
// `counter += interp; while counter > 0 { emit sample; counter -= deci }`.
// Closed form used here, with c0 = counter (<= 0) at the start of the call:
//     out[k] = in[ floor((k*deci - c0) / interp) ]
// Integer arithmetic only; outputs are bit copies.  HBM bound:
// bytes = elem * (N_in + N_out).
#include <algorithm>

#include "common.cuh"
#include "pipeline.cuh"

namespace rrc {

struct U128 { uint64_t a, b; };   // 16-byte element

// (k*D + a) / I without overflowing 64 bits: k = kq*I + kr  =>
// floor((k*D + a)/I) = kq*D + floor((kr*D + a)/I), kr*D + a < 2^63 for I, D < 2^31.
__device__ __forceinline__ void index_of(unsigned long long k, unsigned long long D, unsigned long long I,
                                         unsigned long long a, unsigned long long& idx, unsigned long long& rem) {
    const unsigned long long kq = k / I, kr = k - kq * I;
    const unsigned long long t = kr * D + a;
    const unsigned long long tq = t / I;
    idx = kq * D + tq;
    rem = t - tq * I;
}

constexpr int RS_UNROLL = 4;

// a = -c0 (>= 0).  qG/rG: quotient/remainder of (G*D)/I for the grid stride G.
template <typename E>
__global__ void __launch_bounds__(256) resample_kernel(const E* __restrict__ in, E* __restrict__ out,
                                                       unsigned long long n_out, unsigned long long D,
                                                       unsigned long long I, unsigned long long a,
                                                       unsigned long long qG, unsigned long long rG) {
    const unsigned long long G = (unsigned long long)gridDim.x * blockDim.x;
    unsigned long long k = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n_out) return;
    unsigned long long idx, rem;
    index_of(k, D, I, a, idx, rem);
    // main loop: RS_UNROLL independent gathers in flight per thread
    while (k + (RS_UNROLL - 1) * G < n_out) {
        unsigned long long ix[RS_UNROLL];
#pragma unroll
        for (int u = 0; u < RS_UNROLL; ++u) {
            ix[u] = idx;
            idx += qG; rem += rG;
            if (rem >= I) { rem -= I; ++idx; }
        }
        E v[RS_UNROLL];
#pragma unroll
        for (int u = 0; u < RS_UNROLL; ++u) v[u] = in[ix[u]];
#pragma unroll
        for (int u = 0; u < RS_UNROLL; ++u) out[k + u * G] = v[u];
        k += RS_UNROLL * G;
    }
    while (k < n_out) {
        out[k] = in[idx];
        idx += qG; rem += rG;
        if (rem >= I) { rem -= I; ++idx; }
        k += G;
    }
}

// out[0..n) = *src (flush of the pending sample) ; also used to latch it.
template <typename E>
__global__ void fill_kernel(const E* __restrict__ src, E* __restrict__ out, unsigned long long n) {
    const E v = *src;
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (unsigned long long)gridDim.x * blockDim.x)
        out[i] = v;
}

}  // namespace rrc

using namespace rrc;

struct rrc_resampler {
    int device = 0;
    size_t elem = 0;
    int64_t interp = 1, deci = 1, counter = 0;
    bool has_pending = false;
    void* pending = nullptr;   // 16-byte device slot holding the pending sample
    Pipe pipe;
};

namespace {

uint64_t gcd_u64(uint64_t a, uint64_t b) {   // src/rational_resampler.rs:10-17
    while (b) { uint64_t t = b; b = a % b; a = t; }
    return a;
}

template <typename E>
int launch_gather(rrc_resampler* h, const void* in, void* out, uint64_t n_out, uint64_t a, cudaStream_t st) {
    if (n_out == 0) return RRC_OK;
    const uint64_t per_block = 256;
    uint64_t blocks = (n_out + per_block * RS_UNROLL - 1) / (per_block * RS_UNROLL);
    blocks = std::max<uint64_t>(1, std::min<uint64_t>(blocks, (uint64_t)sm_count(h->device) * 16));
    const uint64_t G = blocks * per_block;
    const uint64_t GD = G * (uint64_t)h->deci;
    resample_kernel<E><<<(unsigned)blocks, 256, 0, st>>>((const E*)in, (E*)out, n_out, (uint64_t)h->deci,
                                                         (uint64_t)h->interp, a, GD / (uint64_t)h->interp,
                                                         GD % (uint64_t)h->interp);
    RRC_CHECK_LAUNCH();
    count_launch();
    return RRC_OK;
}

template <typename E>
int launch_fill(const void* src, void* out, uint64_t n, cudaStream_t st) {
    if (n == 0) return RRC_OK;
    unsigned blocks = (unsigned)std::min<uint64_t>((n + 255) / 256, 1024);
    fill_kernel<E><<<blocks, 256, 0, st>>>((const E*)src, (E*)out, n);
    RRC_CHECK_LAUNCH();
    count_launch();
    return RRC_OK;
}

int gather(rrc_resampler* h, const void* in, void* out, uint64_t n_out, uint64_t a, cudaStream_t st) {
    switch (h->elem) {
    case 1: return launch_gather<uint8_t>(h, in, out, n_out, a, st);
    case 2: return launch_gather<uint16_t>(h, in, out, n_out, a, st);
    case 4: return launch_gather<uint32_t>(h, in, out, n_out, a, st);
    case 8: return launch_gather<uint64_t>(h, in, out, n_out, a, st);
    default: return launch_gather<U128>(h, in, out, n_out, a, st);
    }
}
int fill(rrc_resampler* h, const void* src, void* out, uint64_t n, cudaStream_t st) {
    switch (h->elem) {
    case 1: return launch_fill<uint8_t>(src, out, n, st);
    case 2: return launch_fill<uint16_t>(src, out, n, st);
    case 4: return launch_fill<uint32_t>(src, out, n, st);
    case 8: return launch_fill<uint64_t>(src, out, n, st);
    default: return launch_fill<U128>(src, out, n, st);
    }
}

}  // namespace

extern "C" {

int rrc_resampler_create(int device, size_t elem_size, size_t interp, size_t deci, rrc_resampler_t** out) {
    if (!out) return fail(RRC_ERR_INVALID, "out is NULL");
    *out = nullptr;
    if (deci == 0) return fail(RRC_ERR_INVALID, "RationalResampler created using deci 0");      // :130-132
    if (interp == 0) return fail(RRC_ERR_INVALID, "RationalResampler created using interp 0");  // :133-135
    if (!(elem_size == 1 || elem_size == 2 || elem_size == 4 || elem_size == 8 || elem_size == 16))
        return fail(RRC_ERR_INVALID, "elem_size %zu not in {1,2,4,8,16}", elem_size);
    const uint64_t g = gcd_u64(deci, interp);                                                    // :136-138
    const uint64_t d = deci / g, i = interp / g;
    if (d >= (1ull << 31) || i >= (1ull << 31))
        return fail(RRC_ERR_UNSUPPORTED, "reduced interp/deci must be < 2^31 (got %llu/%llu)",
                    (unsigned long long)i, (unsigned long long)d);
    RRC_CUDA(cudaSetDevice(device));
    auto* h = new rrc_resampler();
    h->device = device; h->elem = elem_size; h->interp = (int64_t)i; h->deci = (int64_t)d;
    cudaError_t e = cudaMalloc(&h->pending, 16);
    if (e != cudaSuccess) { delete h; return fail(RRC_ERR_CUDA, "cudaMalloc failed: %s", cudaGetErrorString(e)); }
    *out = h;
    return RRC_OK;
}

int rrc_resampler_destroy(rrc_resampler_t* h) {
    if (!h) return RRC_OK;
    cudaSetDevice(h->device);
    cudaFree(h->pending);
    h->pipe.destroy();
    delete h;
    return RRC_OK;
}

int rrc_resampler_reset(rrc_resampler_t* h) {
    if (!h) return fail(RRC_ERR_INVALID, "resampler handle is NULL");
    h->counter = 0; h->has_pending = false;
    return RRC_OK;
}

// The reference's carried state (src/rational_resampler.rs:101-105: `counter`, `pending`) made settable, so
// that a time-segment shard can start mid-stream: the shard owning outputs [k_lo, k_hi)
// starts at input sample s = floor(k_lo*deci/interp) with counter = s*interp - k_lo*deci (<= 0; the work()
// loop `counter += interp; while counter > 0 { emit; counter -= deci }` is well defined for any counter).
// pending_host != NULL: that sample is the pending one (:161-173) and counter must be > 0.
int rrc_resampler_set_state(rrc_resampler_t* h, int64_t counter, const void* pending_host) {
    if (!h) return fail(RRC_ERR_INVALID, "resampler handle is NULL");
    if (pending_host) {
        if (counter <= 0) return fail(RRC_ERR_INVALID, "a pending sample needs counter > 0 (got %lld)", (long long)counter);
        RRC_CUDA(cudaSetDevice(h->device));
        RRC_CUDA(cudaMemcpy(h->pending, pending_host, h->elem, cudaMemcpyHostToDevice));
        RRC_CUDA(cudaStreamSynchronize(cudaStreamLegacy));
    } else if (counter > 0) {
        return fail(RRC_ERR_INVALID, "counter %lld > 0 without a pending sample", (long long)counter);
    }
    h->counter = counter;
    h->has_pending = pending_host != nullptr;
    return RRC_OK;
}

int rrc_resampler_state(const rrc_resampler_t* h, int64_t* interp, int64_t* deci, int64_t* counter, int* has_pending) {
    if (!h) return fail(RRC_ERR_INVALID, "resampler handle is NULL");
    if (interp) *interp = h->interp;
    if (deci) *deci = h->deci;
    if (counter) *counter = h->counter;
    if (has_pending) *has_pending = h->has_pending ? 1 : 0;
    return RRC_OK;
}

int rrc_resampler_run(rrc_resampler_t* h, const void* in, size_t n_in, void* out, size_t out_cap,
                      size_t* consumed, size_t* produced, int* wait_on_output, void* stream) {
    if (!h || !consumed || !produced || !wait_on_output) return fail(RRC_ERR_INVALID, "NULL argument");
    *consumed = 0; *produced = 0; *wait_on_output = 1;
    if (out_cap == 0) return RRC_OK;                                         // :157-159
    if (!out) return fail(RRC_ERR_INVALID, "out is NULL");
    RRC_CUDA(cudaSetDevice(h->device));
    cudaStream_t st = as_stream(stream);
    const int64_t I = h->interp, D = h->deci;
    uint64_t opos = 0;
    if (h->has_pending) {                                                    // :161-173
        uint64_t e = h->counter > 0 ? (uint64_t)((h->counter + D - 1) / D) : 0;
        e = std::min<uint64_t>(e, out_cap);
        RRC_TRY(fill(h, h->pending, out, e, st));
        h->counter -= (int64_t)e * D;
        opos = e;
        if (opos == out_cap) { *produced = opos; return RRC_OK; }            // WaitForStream(dst,1)
        h->has_pending = false;
    }
    if (n_in == 0) { *produced = opos; *wait_on_output = 0; return RRC_OK; } // :175-179
    if (!in) return fail(RRC_ERR_INVALID, "in is NULL");
    const uint64_t room = out_cap - opos;
    const int64_t c0 = h->counter;                                           // in (-D, 0]
    // outputs if every input were taken: ceil((c0 + n_in*I)/D), never negative
    const __int128 top = (__int128)c0 + (__int128)n_in * I;
    const uint64_t all = top > 0 ? (uint64_t)((top + D - 1) / D) : 0;
    uint64_t n_main, taken;
    if (all < room) {                                                        // output never fills
        n_main = all; taken = n_in;
        h->counter = (int64_t)(top - (__int128)all * D);
        *wait_on_output = 0;                                                 // WaitForStream(src,1)
    } else {                                                                 // fills on output k = room-1 (:190-195)
        n_main = room;
        const uint64_t k = room - 1;
        const __int128 num = (__int128)k * D - c0;
        const uint64_t s = (uint64_t)(num / I);
        taken = s + 1;
        h->counter = (int64_t)((__int128)c0 + (__int128)taken * I - (__int128)room * D);
        if (h->counter > 0) {
            h->has_pending = true;
            RRC_CUDA(cudaMemcpyAsync(h->pending, (const char*)in + s * h->elem, h->elem, cudaMemcpyDeviceToDevice, st));
        }
        *wait_on_output = 1;
    }
    RRC_TRY(gather(h, in, (char*)out + opos * h->elem, n_main, (uint64_t)(-c0), st));
    *consumed = taken;
    *produced = opos + n_main;
    return RRC_OK;
}

int rrc_resampler_run_host(rrc_resampler_t* h, const void* in_host, size_t n_in, void* out_host, size_t out_cap,
                           size_t* consumed, size_t* produced) {
    if (!h || !consumed || !produced) return fail(RRC_ERR_INVALID, "NULL argument");
    *consumed = 0; *produced = 0;
    RRC_TRY(h->pipe.init(h->device));
    const size_t es = h->elem;
    const size_t chunk_in = pipe_chunk_samples_for(n_in);
    // worst-case outputs of one chunk (+ pending flush)
    const size_t chunk_out = (size_t)(((__int128)chunk_in * h->interp) / h->deci) + (size_t)((h->interp + h->deci - 1) / h->deci) + 2;
    RRC_TRY(h->pipe.reserve(chunk_in * es, chunk_out * es));
    size_t ipos = 0, opos = 0;
    int i = 0;
    for (;; ++i) {
        const size_t ni = n_in > ipos ? pipe_next_chunk((size_t)i, n_in - ipos, chunk_in) : 0;
        const size_t cap = std::min(chunk_out, out_cap - opos);
        if (cap == 0) break;
        RRC_TRY(h->pipe.stage_in(i, (const char*)in_host + ipos * es, ni * es));
        size_t c = 0, p = 0; int w = 0;
        RRC_TRY(rrc_resampler_run(h, h->pipe.d_in[i & 1], ni, h->pipe.d_out[i & 1], cap, &c, &p, &w, h->pipe.s_comp));
        RRC_TRY(h->pipe.drain_out(i, (char*)out_host + opos * es, p * es));
        ipos += c; opos += p;
        if (ipos >= n_in && !h->has_pending) break;
        if (c == 0 && p == 0) break;
    }
    *consumed = ipos; *produced = opos;
    return h->pipe.finish();
}

}  // extern "C"
