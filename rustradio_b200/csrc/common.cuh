// common.cuh — shared plumbing for the rustradio-cuda C ABI implementation.
#pragma once
#include <cuda_runtime.h>

#include <atomic>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstring>

#include "../../include/rustradio_cuda.h"

namespace rrc {

// Thread-local last-error text (rrc_last_error()).
char* err_buf();
int fail(int code, const char* fmt, ...);

extern std::atomic<uint64_t> g_launches;
inline void count_launch(uint64_t n = 1) { g_launches.fetch_add(n, std::memory_order_relaxed); }

#define RRC_CUDA(expr)                                                                   \
    do {                                                                                 \
        cudaError_t _e = (expr);                                                         \
        if (_e != cudaSuccess)                                                           \
            return ::rrc::fail(RRC_ERR_CUDA, "%s failed: %s (%s:%d)", #expr,             \
                               cudaGetErrorString(_e), __FILE__, __LINE__);              \
    } while (0)

#define RRC_CHECK_LAUNCH()                                                               \
    do {                                                                                 \
        cudaError_t _e = cudaGetLastError();                                             \
        if (_e != cudaSuccess)                                                           \
            return ::rrc::fail(RRC_ERR_CUDA, "kernel launch failed: %s (%s:%d)",         \
                               cudaGetErrorString(_e), __FILE__, __LINE__);              \
    } while (0)

#define RRC_TRY(expr)                 \
    do {                              \
        int _s = (expr);              \
        if (_s != RRC_OK) return _s;  \
    } while (0)

inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

// Handle-creation uploads / fills.  cudaMemcpy from pageable host memory and cudaMemset run on the
// legacy default stream and may return before the device has the data; every block and *_run_host
// pipeline runs on cudaStreamNonBlocking streams, which do NOT wait for the legacy stream.  These
// helpers therefore drain the legacy stream before they return, so a handle is complete when its
// constructor returns, whatever stream it is used on next.
inline cudaError_t upload_sync(void* dev, const void* host, size_t bytes) {
    cudaError_t e = cudaMemcpy(dev, host, bytes, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaStreamSynchronize(cudaStreamLegacy);
    return e;
}
inline cudaError_t zero_sync(void* dev, size_t bytes) {
    cudaError_t e = cudaMemset(dev, 0, bytes);
    if (e == cudaSuccess) e = cudaStreamSynchronize(cudaStreamLegacy);
    return e;
}

// Number of SMs of a device (cached).
int sm_count(int device);
int max_smem_optin(int device);

// splitmix64 counter generator shared with oracle/rr_oracle.c::orc_synth_f32.
__host__ __device__ inline uint64_t splitmix64(uint64_t z) {
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
__host__ __device__ inline float synth_value(uint64_t seed, uint64_t index) {
    uint64_t r = splitmix64(seed ^ (index * 0xD1342543DE82EF95ull));
    return (float)(r >> 40) * (1.0f / 8388608.0f) - 1.0f;
}

}  // namespace rrc
