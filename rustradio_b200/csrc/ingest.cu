// ingest.cu — RtlSdrDecode / RtlSdrEncode: RTL-SDR's byte format (u8 I, u8 Q) <-> Complex<f32> on sm_100a.
//
// Replaces RtlSdrDecode::work (rustradio src/rtlsdr_decode.rs:18-48):
//   out[k] = Complex((in[2k] - 127.0) * 0.008, (in[2k+1] - 127.0) * 0.008)   (f32, sub then mul)
// bit-exact with the reference (integer -> f32 conversion is exact, both operations are single
// correctly rounded f32 operations and are never contracted).
//
// Stand-alone kernel: HBM bound, 2 B read + 8 B written per sample.  One thread-iteration converts
// 2 samples (32-bit load, 128-bit store; whole contiguous sectors per warp instruction, 8 iterations
// in flight); a head/tail sample and unaligned pointers take the byte kernel.
// The same decode is fused into the first load of the FIR / FftFilter kernels
// (rrc_fir_set_input_u8iq, rrc_fftfilt_set_input_u8iq), which removes the c32 intermediate
// altogether (10 B/sample of HBM traffic and, end to end, 4x fewer bytes over PCIe).
#include <algorithm>

#include "common.cuh"
#include "pipeline.cuh"

namespace rrc {

__device__ __forceinline__ float dec1(unsigned int b) { return __fmul_rn(__fsub_rn((float)b, 127.0f), 0.008f); }

// Fast path: `in` 4-byte aligned, `out` 16-byte aligned.  One thread-iteration = one 32-bit load
// (2 samples) and one 128-bit store, so every warp instruction moves whole, contiguous sectors
// (128 B in, 512 B out); 8 independent iterations per thread are in flight.
__global__ void __launch_bounds__(256) rtlsdr_decode_kernel(const unsigned int* __restrict__ in, float4* __restrict__ out, long long npairs) {
    const long long nthreads = (long long)gridDim.x * blockDim.x;
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    for (; i + 7 * nthreads < npairs; i += 8 * nthreads) {
        unsigned int w[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) w[k] = __ldcs(in + i + k * nthreads);
#pragma unroll
        for (int k = 0; k < 8; ++k)
            __stcs(out + i + k * nthreads, make_float4(dec1(w[k] & 0xffu), dec1((w[k] >> 8) & 0xffu), dec1((w[k] >> 16) & 0xffu), dec1(w[k] >> 24)));
    }
    for (; i < npairs; i += nthreads) {
        const unsigned int w = in[i];
        out[i] = make_float4(dec1(w & 0xffu), dec1((w >> 8) & 0xffu), dec1((w >> 16) & 0xffu), dec1(w >> 24));
    }
}
// Any alignment: one sample per thread-iteration (byte loads, 64-bit stores).
__global__ void __launch_bounds__(256) rtlsdr_decode_bytes_kernel(const unsigned char* __restrict__ in, float2* __restrict__ out, long long n) {
    const long long nthreads = (long long)gridDim.x * blockDim.x;
    for (long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += nthreads)
        out[k] = make_float2(dec1(in[2 * k]), dec1(in[2 * k + 1]));
}

// ---- RtlSdrEncode (src/rtlsdr_encode.rs:22-26): ((s / 0.008) + 127).round().clamp(0, 255) as u8.  Single correctly rounded
// f32 division and addition (never a multiply by the reciprocal, never contracted), round half away from zero, NaN -> 0
// like Rust's saturating float -> int cast (fmaxf(NaN, 0) = 0).
__device__ __forceinline__ unsigned int enc1(float s) {
    const float v = roundf(__fadd_rn(__fdiv_rn(s, 0.008f), 127.0f));
    return (unsigned int)fminf(fmaxf(v, 0.0f), 255.0f);
}
// Fast path: `in` 16-byte aligned, `out` 4-byte aligned: 2 samples per thread-iteration (128-bit load, 32-bit store).
__global__ void __launch_bounds__(256) rtlsdr_encode_kernel(const float4* __restrict__ in, unsigned int* __restrict__ out, long long npairs) {
    const long long nthreads = (long long)gridDim.x * blockDim.x;
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    for (; i + 3 * nthreads < npairs; i += 4 * nthreads) {
        float4 v[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) v[k] = __ldcs(in + i + k * nthreads);
#pragma unroll
        for (int k = 0; k < 4; ++k)
            __stcs(out + i + k * nthreads, enc1(v[k].x) | (enc1(v[k].y) << 8) | (enc1(v[k].z) << 16) | (enc1(v[k].w) << 24));
    }
    for (; i < npairs; i += nthreads) {
        const float4 v = in[i];
        out[i] = enc1(v.x) | (enc1(v.y) << 8) | (enc1(v.z) << 16) | (enc1(v.w) << 24);
    }
}
// Any alignment: one sample per thread-iteration (64-bit load, two byte stores).
__global__ void __launch_bounds__(256) rtlsdr_encode_bytes_kernel(const float2* __restrict__ in, unsigned char* __restrict__ out, long long n) {
    const long long nthreads = (long long)gridDim.x * blockDim.x;
    for (long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += nthreads) {
        const float2 v = in[k];
        out[2 * k] = (unsigned char)enc1(v.x);
        out[2 * k + 1] = (unsigned char)enc1(v.y);
    }
}

}  // namespace rrc

using namespace rrc;

extern "C" {

int rrc_rtlsdr_decode_plan(size_t in_len_bytes, size_t out_free, size_t* consume_bytes, size_t* produce,
                           size_t* wait_need, int* wait_on_output) {
    if (!consume_bytes || !produce || !wait_need || !wait_on_output) return fail(RRC_ERR_INVALID, "NULL argument");
    // The loop of src/rtlsdr_decode.rs:20-46 run to its WaitForStream.
    const size_t usable = in_len_bytes & ~(size_t)1;                      // :23
    const size_t take = std::min(usable, out_free * 2);                   // :32
    *consume_bytes = take;
    *produce = take / 2;
    if (((in_len_bytes - take) & ~(size_t)1) == 0) { *wait_on_output = 0; *wait_need = 2; }   // :24-26
    else { *wait_on_output = 1; *wait_need = 1; }                         // :28-30
    return RRC_OK;
}

int rrc_rtlsdr_decode_run(int device, const unsigned char* in_dev, size_t n_bytes, float* out_dev, void* stream) {
    const size_t n = n_bytes / 2;
    if (n == 0) return RRC_OK;
    if (!in_dev || !out_dev) return fail(RRC_ERR_INVALID, "in/out is NULL");
    RRC_CUDA(cudaSetDevice(device));
    cudaStream_t st = as_stream(stream);
    const uintptr_t a = reinterpret_cast<uintptr_t>(in_dev), o = reinterpret_cast<uintptr_t>(out_dev);
    float2* out = reinterpret_cast<float2*>(out_dev);
    const int max_grid = sm_count(device) * 8;
    // head: 0 or 1 samples so that the body's input is 4-byte and its output 16-byte aligned
    // (possible iff both pointers need the same parity of head samples).
    const bool in_odd = (a & 3) == 2, out_odd = (o & 15) == 8;
    const bool fast = (a & 1) == 0 && (o & 7) == 0 && in_odd == out_odd;
    size_t done = 0;
    if (fast) {
        const size_t head = in_odd ? 1 : 0;
        const size_t npairs = (n - std::min(n, head)) / 2;
        if (npairs) {
            const unsigned grid = (unsigned)std::min<size_t>((npairs + 256 * 8 - 1) / (256 * 8), (size_t)max_grid);
            rtlsdr_decode_kernel<<<grid, 256, 0, st>>>(reinterpret_cast<const unsigned int*>(in_dev + 2 * head),
                                                       reinterpret_cast<float4*>(out + head), (long long)npairs);
            RRC_CHECK_LAUNCH();
            count_launch();
        }
        if (head && n) {                                        // the one head sample
            rtlsdr_decode_bytes_kernel<<<1, 32, 0, st>>>(in_dev, out, 1);
            RRC_CHECK_LAUNCH();
            count_launch();
        }
        done = std::min(n, head + 2 * npairs);
    }
    if (done < n) {                                             // tail sample, or everything when unaligned
        const size_t rest = n - done;
        const unsigned grid = (unsigned)std::min<size_t>((rest + 255) / 256, (size_t)max_grid * 4);
        rtlsdr_decode_bytes_kernel<<<grid, 256, 0, st>>>(in_dev + 2 * done, out + done, (long long)rest);
        RRC_CHECK_LAUNCH();
        count_launch();
    }
    return RRC_OK;
}

int rrc_rtlsdr_decode_run_host(int device, const unsigned char* in_host, size_t n_bytes, float* out_host, size_t* n_out) {
    const size_t total = n_bytes / 2;
    if (n_out) *n_out = total;
    if (total == 0) return RRC_OK;
    if (!in_host || !out_host) return fail(RRC_ERR_INVALID, "in/out is NULL");
    Pipe pipe;
    int s = pipe.init(device);
    if (s == RRC_OK) {
        const size_t chunk = PIPE_CHUNK_SAMPLES;
        s = pipe.reserve(std::min(chunk, total) * 2, std::min(chunk, total) * sizeof(float2));
        int i = 0;
        for (size_t off = 0; s == RRC_OK && off < total; off += chunk, ++i) {
            const size_t n = std::min(chunk, total - off);
            s = pipe.stage_in(i, in_host + 2 * off, n * 2);
            if (s == RRC_OK) s = rrc_rtlsdr_decode_run(device, (const unsigned char*)pipe.d_in[i & 1], n * 2, (float*)pipe.d_out[i & 1], pipe.s_comp);
            if (s == RRC_OK) s = pipe.drain_out(i, out_host + 2 * off, n * sizeof(float2));
        }
        if (s == RRC_OK) s = pipe.finish();
    }
    pipe.destroy();
    return s;
}

int rrc_rtlsdr_encode_plan(size_t in_len, size_t out_free_bytes, size_t* consume, size_t* produce_bytes,
                           size_t* wait_need, int* wait_on_output) {
    if (!consume || !produce_bytes || !wait_need || !wait_on_output) return fail(RRC_ERR_INVALID, "NULL argument");
    // The loop of src/rtlsdr_encode.rs:30-51 run to its WaitForStream.
    const size_t take = std::min(in_len, out_free_bytes / 2);             // :41
    *consume = take;
    *produce_bytes = 2 * take;
    if (in_len - take == 0) { *wait_on_output = 0; *wait_need = 1; }      // :33-35
    else { *wait_on_output = 1; *wait_need = 2; }                         // :37-39
    return RRC_OK;
}

int rrc_rtlsdr_encode_run(int device, const float* in_dev_c32, size_t n, unsigned char* out_dev, void* stream) {
    if (n == 0) return RRC_OK;
    if (!in_dev_c32 || !out_dev) return fail(RRC_ERR_INVALID, "in/out is NULL");
    RRC_CUDA(cudaSetDevice(device));
    cudaStream_t st = as_stream(stream);
    const uintptr_t a = reinterpret_cast<uintptr_t>(in_dev_c32), o = reinterpret_cast<uintptr_t>(out_dev);
    const float2* in = reinterpret_cast<const float2*>(in_dev_c32);
    const int max_grid = sm_count(device) * 8;
    // head: 0 or 1 samples so that the body's input is 16-byte and its output 4-byte aligned
    const bool in_odd = (a & 15) == 8, out_odd = (o & 3) == 2;
    const bool fast = (a & 7) == 0 && (o & 1) == 0 && in_odd == out_odd;
    size_t done = 0;
    if (fast) {
        const size_t head = in_odd ? 1 : 0;
        const size_t npairs = (n - std::min(n, head)) / 2;
        if (npairs) {
            const unsigned grid = (unsigned)std::min<size_t>((npairs + 256 * 4 - 1) / (256 * 4), (size_t)max_grid);
            rtlsdr_encode_kernel<<<grid, 256, 0, st>>>(reinterpret_cast<const float4*>(in + head),
                                                       reinterpret_cast<unsigned int*>(out_dev + 2 * head), (long long)npairs);
            RRC_CHECK_LAUNCH();
            count_launch();
        }
        if (head && n) {
            rtlsdr_encode_bytes_kernel<<<1, 32, 0, st>>>(in, out_dev, 1);
            RRC_CHECK_LAUNCH();
            count_launch();
        }
        done = std::min(n, head + 2 * npairs);
    }
    if (done < n) {                                             // tail sample, or everything when unaligned
        const size_t rest = n - done;
        const unsigned grid = (unsigned)std::min<size_t>((rest + 255) / 256, (size_t)max_grid * 4);
        rtlsdr_encode_bytes_kernel<<<grid, 256, 0, st>>>(in + done, out_dev + 2 * done, (long long)rest);
        RRC_CHECK_LAUNCH();
        count_launch();
    }
    return RRC_OK;
}

int rrc_rtlsdr_encode_run_host(int device, const float* in_host_c32, size_t n, unsigned char* out_host, size_t* n_out_bytes) {
    if (n_out_bytes) *n_out_bytes = 2 * n;
    if (n == 0) return RRC_OK;
    if (!in_host_c32 || !out_host) return fail(RRC_ERR_INVALID, "in/out is NULL");
    Pipe pipe;
    int s = pipe.init(device);
    if (s == RRC_OK) {
        const size_t chunk = pipe_chunk_samples_for(n);
        s = pipe.reserve(std::min(chunk, n) * sizeof(float2), std::min(chunk, n) * 2);
        int i = 0;
        for (size_t off = 0, m = 0; s == RRC_OK && off < n; off += m, ++i) {
            m = pipe_next_chunk((size_t)i, n - off, chunk);
            s = pipe.stage_in(i, reinterpret_cast<const char*>(in_host_c32) + off * sizeof(float2), m * sizeof(float2));
            if (s == RRC_OK) s = rrc_rtlsdr_encode_run(device, (const float*)pipe.d_in[i & 1], m, (unsigned char*)pipe.d_out[i & 1], pipe.s_comp);
            if (s == RRC_OK) s = pipe.drain_out(i, out_host + 2 * off, m * 2);
        }
        if (s == RRC_OK) s = pipe.finish();
    }
    pipe.destroy();
    return s;
}

}  // extern "C"
