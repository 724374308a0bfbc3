// pipeline.cuh — double-buffered H2D -> kernel -> D2H helper used by the
// *_run_host entry points (the end-to-end path: host buffers in, host buffers
// out, copies overlapped with compute on three CUDA streams).
#pragma once
#include <algorithm>
#include <cstdlib>

#include "common.cuh"

namespace rrc {

struct Pipe {
    int device = -1;
    cudaStream_t s_h2d = nullptr, s_comp = nullptr, s_d2h = nullptr;
    cudaEvent_t in_ready[2] = {nullptr, nullptr}, comp_done[2] = {nullptr, nullptr}, d2h_done[2] = {nullptr, nullptr};
    void* d_in[2] = {nullptr, nullptr};
    void* d_out[2] = {nullptr, nullptr};
    size_t in_bytes = 0, out_bytes = 0;

    int init(int dev) {
        if (device == dev) return RRC_OK;
        device = dev;
        RRC_CUDA(cudaSetDevice(dev));
        RRC_CUDA(cudaStreamCreateWithFlags(&s_h2d, cudaStreamNonBlocking));
        RRC_CUDA(cudaStreamCreateWithFlags(&s_comp, cudaStreamNonBlocking));
        RRC_CUDA(cudaStreamCreateWithFlags(&s_d2h, cudaStreamNonBlocking));
        for (int i = 0; i < 2; ++i) {
            RRC_CUDA(cudaEventCreateWithFlags(&in_ready[i], cudaEventDisableTiming));
            RRC_CUDA(cudaEventCreateWithFlags(&comp_done[i], cudaEventDisableTiming));
            RRC_CUDA(cudaEventCreateWithFlags(&d2h_done[i], cudaEventDisableTiming));
        }
        return RRC_OK;
    }
    int reserve(size_t in_b, size_t out_b) {
        if (in_b > in_bytes) {
            for (int i = 0; i < 2; ++i) {
                if (d_in[i]) RRC_CUDA(cudaFree(d_in[i]));
                d_in[i] = nullptr;
                RRC_CUDA(cudaMalloc(&d_in[i], in_b));
            }
            in_bytes = in_b;
        }
        if (out_b > out_bytes) {
            for (int i = 0; i < 2; ++i) {
                if (d_out[i]) RRC_CUDA(cudaFree(d_out[i]));
                d_out[i] = nullptr;
                RRC_CUDA(cudaMalloc(&d_out[i], out_b));
            }
            out_bytes = out_b;
        }
        return RRC_OK;
    }
    // Stage chunk `i`'s input; returns after enqueueing.
    int stage_in(int i, const void* host_src, size_t bytes) {
        const int b = i & 1;
        RRC_CUDA(cudaStreamWaitEvent(s_h2d, comp_done[b], 0));   // compute of chunk i-2 done with d_in[b]
        if (bytes) RRC_CUDA(cudaMemcpyAsync(d_in[b], host_src, bytes, cudaMemcpyHostToDevice, s_h2d));
        RRC_CUDA(cudaEventRecord(in_ready[b], s_h2d));
        RRC_CUDA(cudaStreamWaitEvent(s_comp, in_ready[b], 0));
        RRC_CUDA(cudaStreamWaitEvent(s_comp, d2h_done[b], 0));   // D2H of chunk i-2 done with d_out[b]
        return RRC_OK;
    }
    // After the kernel for chunk `i` was enqueued on s_comp: copy its output back.
    int drain_out(int i, void* host_dst, size_t bytes) {
        const int b = i & 1;
        RRC_CUDA(cudaEventRecord(comp_done[b], s_comp));
        RRC_CUDA(cudaStreamWaitEvent(s_d2h, comp_done[b], 0));
        if (bytes) RRC_CUDA(cudaMemcpyAsync(host_dst, d_out[b], bytes, cudaMemcpyDeviceToHost, s_d2h));
        RRC_CUDA(cudaEventRecord(d2h_done[b], s_d2h));
        return RRC_OK;
    }
    int finish() {
        RRC_CUDA(cudaStreamSynchronize(s_h2d));
        RRC_CUDA(cudaStreamSynchronize(s_comp));
        RRC_CUDA(cudaStreamSynchronize(s_d2h));
        return RRC_OK;
    }
    void destroy() {
        if (device < 0) return;
        cudaSetDevice(device);
        for (int i = 0; i < 2; ++i) {
            if (d_in[i]) cudaFree(d_in[i]);
            if (d_out[i]) cudaFree(d_out[i]);
            if (in_ready[i]) cudaEventDestroy(in_ready[i]);
            if (comp_done[i]) cudaEventDestroy(comp_done[i]);
            if (d2h_done[i]) cudaEventDestroy(d2h_done[i]);
        }
        if (s_h2d) cudaStreamDestroy(s_h2d);
        if (s_comp) cudaStreamDestroy(s_comp);
        if (s_d2h) cudaStreamDestroy(s_d2h);
        device = -1;
    }
};

// Samples per host-pipeline chunk (bytes = this * element size).  RRC_PIPE_CHUNK_LOG2 overrides
// the default 2^23 (64 MiB of c32) for experiments.
inline size_t pipe_chunk_samples() {
    const char* e = getenv("RRC_PIPE_CHUNK_LOG2");     // read per call: tests shrink it to force many chunks
    int l = e ? atoi(e) : 23;
    if (l < 12) l = 12;
    if (l > 28) l = 28;
    return (size_t)1 << l;
}
#define PIPE_CHUNK_SAMPLES (::rrc::pipe_chunk_samples())
// Chunk for a call that moves `total` samples: the first chunk's H2D and the last chunk's D2H are not overlapped, so a
// call of only a few default chunks loses a large part of the copy time (config 1, 2^24 samples = 2 chunks: 0.72 of
// the bare-copy ceiling; with 2^21-sample chunks 0.89), while chunks below 2^21 samples cost more per copy than they
// hide (config 2: 0.94 at 2^23, 0.88 at 2^21; profiles/r02_e2e_chunk_sweep.txt).  Aim for 8 chunks inside [2^21, 2^23].
inline size_t pipe_chunk_samples_for(size_t total) {
    if (getenv("RRC_PIPE_CHUNK_LOG2")) return pipe_chunk_samples();
    size_t c = (size_t)1 << 23;
    while (c > ((size_t)1 << 21) && c * 8 > total) c >>= 1;
    return c;
}

// Tapered schedule: the first chunk's H2D copy and the last chunk's D2H copy have nothing to overlap with, so the
// chunks ramp up from maxc/8 (x2 per chunk) and ramp down again at the end: the exposed copies shrink 8x for four extra
// chunks at each end.  `k` = chunks already issued, `remaining` = units left; returns the size of the next chunk.
inline size_t pipe_next_chunk(size_t k, size_t remaining, size_t maxc) {
    if (getenv("RRC_PIPE_CHUNK_LOG2")) return std::min(maxc, remaining);       // tests: fixed chunks
    const size_t minc = std::max<size_t>(1, maxc >> 3);
    size_t c = k < 3 ? std::min(maxc, minc << k) : maxc;                       // ramp up
    while (c > minc && c * 2 > remaining) c >>= 1;                             // ramp down: at most half of what is left
    if (remaining <= minc + minc / 2) c = remaining;                           // no crumbs
    return std::min(c, remaining);
}

}  // namespace rrc
