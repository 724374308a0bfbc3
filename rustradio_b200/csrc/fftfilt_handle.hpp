// fftfilt_handle.hpp — the FftFilter handle shared by fftfilt.cu (plain overlap-save kernels) and
// fftfilt_fold.cu (decimate-by-8 fused kernel with the pruned inverse transform).
#pragma once
#include <vector>

#include "common.cuh"
#include "pipeline.cuh"
#include "epilogue.cuh"

// Tables of the decimate-by-8 fold kernel (fftfilt_fold_core.cuh), built lazily on first use.
struct rrc_fold_tables {
    int nc = 0;                 // CTAs per cluster: 1 (N = 16384) or 4 (N = 65536); 0 = not built
    float2* Hc = nullptr;       // [nc][16384] spectrum rows, phase-C order
    float2* gc = nullptr;       // [nc][512]   W_{16384 nc}^{c t}
    float2* twc = nullptr;      // [nc][32]    W_128^{c n1}
    float2* twm = nullptr;      // [2048]      W_{2048 nc}^{m2}
    int max_clusters = 0;       // co-resident clusters (cudaOccupancyMaxActiveClusters)
};

// Tables of the polyphase decimating kernel (fftfilt_poly_core.cuh), built lazily for the decimation of the first call.
struct rrc_poly_tables {
    int D = 0;                  // decimation the tables were built for; 0 = not built
    float2* Hph[16] = {};       // by skip mod D: [D][16384] spectra of the polyphase branches, phase-C order
    float4* scratch = nullptr;  // [clusters][C][4][8192] partial sums on their way to the block's finisher CTA
    size_t scratch_bytes = 0;
    int max_clusters[3] = {0, 0, 0};   // co-resident clusters for C = 1, 2, 4
    cudaStream_t side = nullptr;       // second launch (two-CTA clusters on the SMs the four-CTA clusters leave idle)
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
};

struct rrc_fftfilt {
    int device = 0;
    size_t ntaps = 0;
    std::vector<float> taps_host;     // interleaved c32 taps (kept for lazily built tables)
    int T1 = 0;                       // ntaps - 1 (history length)
    // Long filters are split into tap partitions of <= PART_TAPS taps; partition p filters the
    // input delayed by p*PART_TAPS and accumulates into the output (y = sum_p h_p * x(n - p*L)).
    std::vector<int> part_T1;         // taps of partition p, minus 1
    std::vector<float2*> part_Hp;     // spectrum of partition p (512-thread layout)
    std::vector<float2*> part_Hd;     // spectrum of partition p (1024-thread layout)
    std::vector<float2*> part_Hq;     // spectrum of partition p (packed kernel: re-pairs / im-pairs per thread, fftfilt_pk.cuh)
    float2* tw2p = nullptr;           // packed W_512 table of the packed kernel
    float2* Hp = nullptr;             // == part_Hp[0]
    float2* tw1_16 = nullptr;
    float2* tw2_16 = nullptr;
    float2* tw3_16 = nullptr;
    int real = 0;                     // real stream + real taps (rrc_fftfilt_f32_create): f32 in / out / history
    int in_u8 = 0;                    // 1: run() inputs are u8 I/Q pairs (rrc_fftfilt_set_input_u8iq)
    rrc::Epi epi;                     // fused store epilogue (rrc_fftfilt_set_epilogue)
    // kernel variant (RRC_FFTFILT_VARIANT): fftfilt_tmh_kernel = 36 with per-thread constants in tensor memory: 40 = the whole
    // spectrum, 41 = spectrum + the phase-A / A' twiddle powers, 42 = spectrum + the phase-B / B' twiddles (default),
    // 37 = PACKED FP32 lanes + TMA-staged input (fftfilt_pk.cuh),
    // 36 = 512 threads x 32 points with TMA-staged input, half the spectrum in shared memory, half from L2,
    // 32 = the same with LDG input + L2 prefetch, 16 = 1024 threads x 16 points, 33/34/35 = experiments
    int variant = 42;
    float2* tw1 = nullptr;
    float2* tw2 = nullptr;
    float2* hist[2] = {nullptr, nullptr};
    int cur = 0;
    // One-shot external history (rrc_fftfilt_set_history_ptr): the next run reads its left halo through
    // this pointer — e.g. the tail of the neighbouring GPU's input over NVLink — instead of hist[cur].
    const float2* hist_ext = nullptr;
    // reset()/set_history() enqueue on the caller's stream; the *_run_host pipelines run on their own
    // non-blocking streams and wait for this event first.
    cudaEvent_t state_ev = nullptr;
    bool state_dirty = false;
    rrc::Pipe pipe;
    rrc_fold_tables fold;
    rrc_poly_tables poly;
};

namespace rrc {
// fftfilt_fold.cu: FftFilter + decimate-by-8 with folded spectrum.  Returns RRC_ERR_UNSUPPORTED
// (without setting the error text) when the geometry is not covered, so the caller can fall back
// to the store-predicate path.
int fold_supported(const rrc_fftfilt* h, size_t deci);
int fold_launch(rrc_fftfilt* h, const float* in, size_t n, float* out, size_t n_out,
                size_t skip, cudaStream_t st);
void fold_destroy(rrc_fftfilt* h);
// fftfilt_poly.cu: FftFilter + decimate-by-D as a polyphase filter (D forward transforms, one inverse per block).
// Same convention: RRC_ERR_UNSUPPORTED = not covered / nothing launched.
int poly_supported(const rrc_fftfilt* h, size_t deci);
int poly_launch(rrc_fftfilt* h, const float* in, size_t n, float* out, size_t n_out, size_t deci,
                size_t skip, cudaStream_t st);
void poly_destroy(rrc_fftfilt* h);
}  // namespace rrc
