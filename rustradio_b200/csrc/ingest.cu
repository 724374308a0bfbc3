// ingest.cu — RtlSdrDecode: RTL-SDR's byte format (u8 I, u8 Q) -> Complex<f32> on sm_100a.
//
// Replaces RtlSdrDecode::work (rustradio src/rtlsdr_decode.rs:18-48):
//   out[k] = Complex((in[2k] - 127.0) * 0.008, (in[2k+1] - 127.0) * 0.008)   (f32, sub then mul)
// bit-exact with the reference (integer -> f32 conversion is exact, both operations are single
// correctly rounded f32 operations and are never contracted).
//
// Stand-alone kernel: HBM bound, 2 B read + 8 B written per sample.  Each thread converts 8 samples:
// one 128-bit load, four 128-bit stores; unaligned heads/tails and odd pointers take the 2-byte path.
// The same decode is fused into the first load of the FIR / FftFilter kernels
// (rrc_fir_set_input_u8iq, rrc_fftfilt_set_input_u8iq), which removes the c32 intermediate
// altogether (10 B/sample of HBM traffic and, end to end, 4x fewer bytes over PCIe).
#include <algorithm>

#include "common.cuh"
#include "pipeline.cuh"

namespace rrc {

__device__ __forceinline__ float dec1(unsigned int b) { return __fmul_rn(__fsub_rn((float)b, 127.0f), 0.008f); }

__global__ void __launch_bounds__(256) rtlsdr_decode_kernel(const unsigned char* __restrict__ in, float2* __restrict__ out,
                                                            long long n /* samples */, long long head /* samples before the aligned body */) {
    const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long nthreads = (long long)gridDim.x * blockDim.x;
    // body: groups of 8 samples, 16-byte aligned in `in`, starting at sample `head`
    const long long ngroups = n > head ? (n - head) / 8 : 0;
    const uint4* in16 = reinterpret_cast<const uint4*>(in + 2 * head);
    for (long long g = tid; g < ngroups; g += nthreads) {
        const uint4 w = __ldcs(in16 + g);
        float2* o = out + head + g * 8;
        const unsigned int ww[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const float2 a = make_float2(dec1(ww[q] & 0xffu), dec1((ww[q] >> 8) & 0xffu));
            const float2 b = make_float2(dec1((ww[q] >> 16) & 0xffu), dec1(ww[q] >> 24));
            if ((reinterpret_cast<uintptr_t>(o) & 15) == 0) {
                __stcs(reinterpret_cast<float4*>(o + 2 * q), make_float4(a.x, a.y, b.x, b.y));
            } else {
                o[2 * q] = a; o[2 * q + 1] = b;
            }
        }
    }
    // head and tail: one sample per thread
    const long long tail0 = head + ngroups * 8;
    for (long long s = tid; s < head + (n - tail0); s += nthreads) {
        const long long k = s < head ? s : tail0 + (s - head);
        out[k] = make_float2(dec1(in[2 * k]), dec1(in[2 * k + 1]));
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
    // samples before `in` reaches 16-byte alignment (only possible when `in` is even)
    const uintptr_t a = reinterpret_cast<uintptr_t>(in_dev);
    long long head = (a & 1) ? (long long)n : (long long)(((16 - (a & 15)) & 15) / 2);
    head = std::min<long long>(head, (long long)n);
    const long long work = std::max<long long>(((long long)n - head) / 8, 1);
    const unsigned grid = (unsigned)std::min<long long>((work + 255) / 256, (long long)sm_count(device) * 16);
    rtlsdr_decode_kernel<<<grid, 256, 0, as_stream(stream)>>>(in_dev, reinterpret_cast<float2*>(out_dev), (long long)n, head);
    RRC_CHECK_LAUNCH();
    count_launch();
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

}  // extern "C"
