// fftfilt_fold.cu — FftFilter fused with decimate-by-8, folded spectrum + pruned inverse
// (fftfilt_fold_core.cuh).  NC = 1: one CTA per 16384-point block; NC = 4: one 4-CTA cluster per
// 65536-point block, partial inverse results combined over distributed shared memory.
//
// Replaces FftFilter::work's engine step (rustradio src/fft_filter.rs:172-176,281-287,331-348)
// followed by RationalResampler(1, 8) (src/rational_resampler.rs:155-206) — BASELINE config 5.
#include <cooperative_groups.h>

#include <algorithm>
#include <cmath>
#include <complex>
#include <cstdlib>
#include <vector>

#include "common.cuh"
#include "fftfilt_fold_core.cuh"
#include "fftfilt_handle.hpp"
#include "fftfilt_tables.hpp"

namespace cg = cooperative_groups;

namespace rrc {

using fftf::FoldIO;

constexpr int FOLD_U_OFF = 4096;      // float2 offset of the u array inside the exchange buffer (past P2)
static_assert(fftf::P2_ELEMS <= FOLD_U_OFF && FOLD_U_OFF + fftf::LU <= fftk::SMEM_ELEMS, "fold staging layout");
constexpr int FOLD_NSTAMP = 12, FOLD_TRACE_BLK = 6;
constexpr size_t FOLD_SMEM = (size_t)(fftk::SMEM_ELEMS + 512 + 512 + fftk::HRES_ELEMS + 512 + 32) * sizeof(float2);

__device__ __forceinline__ void task_barrier() { asm volatile("bar.sync 3, 64;" ::: "memory"); }

template <int NC>
__global__ void __launch_bounds__(fftk::NT, 1)
fftfilt_fold_kernel(const FoldIO io, const float2* __restrict__ Hc, const float2* __restrict__ tw1g,
                    const float2* __restrict__ tw2g, const float2* __restrict__ gcg,
                    const float2* __restrict__ twcg, const float2* __restrict__ twm, long long nblocks,
                    long long* __restrict__ trace) {
    extern __shared__ __align__(16) float2 sm[];
    float2* s_tw2 = sm + fftk::SMEM_ELEMS;
    float2* s_tw1 = s_tw2 + 512;
    float2* s_hres = s_tw1 + 512;
    float2* s_gc = s_hres + fftk::HRES_ELEMS;
    float2* s_twc = s_gc + 512;
    const int tid = threadIdx.x;
    int c = 0;
    if constexpr (NC > 1) c = (int)cg::this_cluster().block_rank();
    const float2* Hp = Hc + (size_t)c * fftk::N;
    s_tw2[tid] = tw2g[tid];
    s_tw1[tid] = tw1g[tid];
    if constexpr (NC > 1) {
        s_gc[tid] = gcg[c * 512 + tid];
        if (tid < 32) s_twc[tid] = twcg[c * 32 + tid];
    }
    fftk::load_hres(tid, Hp, s_hres);
    if (io.hist_next && blockIdx.x == gridDim.x - 1) fftf::update_history(io, tid, fftk::NT);
    __syncthreads();
    const float2* uc[NC];
    if constexpr (NC > 1) {
        auto cluster = cg::this_cluster();
#pragma unroll
        for (int i = 0; i < NC; ++i) uc[i] = cluster.map_shared_rank(sm + FOLD_U_OFF, i);
    } else {
        uc[0] = sm + FOLD_U_OFF;
    }
    const long long cl = blockIdx.x / NC, ncl = gridDim.x / NC;
    bool pending = false;         // a relaxed cluster-barrier arrival of the previous block is outstanding
    // RRC_FFTFILT_TRACE (debug): every warp of CTA 1 of cluster 3 stamps clock64 at the phase boundaries of blocks 2..7
    int it = 0;
    auto stamp = [&](int i) {
        if (trace && cl == 3 && c == 1 % NC && it >= 2 && it < 2 + FOLD_TRACE_BLK && (tid & 31) == 0)
            trace[((size_t)(it - 2) * 16 + (tid >> 5)) * FOLD_NSTAMP + i] = clock64();
    };
    for (long long blk = cl; blk < nblocks; blk += ncl, ++it) {
        stamp(0);
        // The other CTAs of the cluster may still be reading this CTA's u array (previous block); the
        // matching wait sits after this block's loads and DFT32, right before the first shared-memory
        // write, so it is normally already satisfied.
        fftf::phase_a<NC>(tid, c, blk, io, s_tw1, s_gc, s_twc, sm, [&]() {
            stamp(1);
            if constexpr (NC > 1) { if (pending) asm volatile("barrier.cluster.wait.aligned;" ::: "memory"); }
            stamp(2);
        });
        stamp(3);
        __syncthreads();
        stamp(4);
        {   // next block's segment -> L2: every CTA of the cluster pulls one quarter (16 warps x 8 KiB)
            const long long nb = blk + ncl;
            const long long seg0 = fftf::seg_start<NC>(nb, io) + (long long)c * fftk::N + (long long)(tid >> 5) * 1024;
            if ((tid & 31) == 0 && nb < nblocks && seg0 >= 0 && seg0 + 1024 <= io.n_in) {
                const unsigned long long esz = io.in_u8 ? 2 : 8;
                const unsigned long long a = (reinterpret_cast<unsigned long long>(io.in) + (unsigned long long)seg0 * esz + 15ull) & ~15ull;
                asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(a), "r"((unsigned)(1024 * esz - 16)) : "memory");
            }
        }
        fftk::phase_mid_b(tid, s_tw2, sm);
        __syncwarp();
        stamp(5);
        fftf::phase_c_fold(tid, Hp, s_hres, sm);
        stamp(6);
        __syncthreads();
        stamp(7);
        if (tid < 64) {                                         // 2048-point inverse: 64 tasks x 32 points, twice
            float2 v[32];
            fftf::inv1_load(tid, sm, v);
            task_barrier();                                     // P2 overlaps folded cells other tasks still read
            fftf::inv1_compute_store(tid, s_tw1, v, sm);
            task_barrier();
            fftf::inv2_load(tid, sm, v);
            fftf::inv2_compute_store(tid, v, sm + FOLD_U_OFF);
        }
        stamp(8);
        if constexpr (NC > 1) cg::this_cluster().sync(); else __syncthreads();
        stamp(9);
        fftf::combine_store<NC>(tid, c, blk, io, uc, twm);
        stamp(10);
        // The next phase A overwrites the u array other CTAs are still reading: write-after-read only,
        // so a RELAXED arrival is enough (no release fence waiting for the global stores to drain).
        if constexpr (NC > 1) { asm volatile("barrier.cluster.arrive.relaxed.aligned;" ::: "memory"); pending = true; }
        else __syncthreads();
    }
    if constexpr (NC > 1) { if (pending) asm volatile("barrier.cluster.wait.aligned;" ::: "memory"); }
}

namespace {

template <int NC>
int launch_fold(rrc_fftfilt* h, const FoldIO& io, long long nblocks, cudaStream_t st) {
    auto kern = fftfilt_fold_kernel<NC>;
    RRC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)FOLD_SMEM));
    cudaLaunchConfig_t cfg = {};
    cfg.blockDim = dim3(fftk::NT);
    cfg.dynamicSmemBytes = FOLD_SMEM;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = NC; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    if (h->fold.max_clusters == 0) {
        int mc = 0;
        cfg.gridDim = dim3(sm_count(h->device) / NC * NC);
        if (NC > 1) {
            RRC_CUDA(cudaOccupancyMaxActiveClusters(&mc, kern, &cfg));
            if (mc < 1) return fail(RRC_ERR_CUDA, "no %d-CTA cluster of the fold kernel fits on device %d", NC, h->device);
        } else {
            mc = sm_count(h->device);
        }
        h->fold.max_clusters = mc;
    }
    const long long ncl = std::min<long long>(nblocks, h->fold.max_clusters);
    cfg.gridDim = dim3((unsigned)(ncl * NC));
    static const bool want_trace = getenv("RRC_FFTFILT_TRACE") != nullptr;
    long long* dtrace = nullptr;
    const size_t trace_n = (size_t)FOLD_TRACE_BLK * 16 * FOLD_NSTAMP;
    if (want_trace) { RRC_CUDA(cudaMalloc((void**)&dtrace, trace_n * 8)); RRC_CUDA(cudaMemsetAsync(dtrace, 0, trace_n * 8, st)); }
    RRC_CUDA(cudaLaunchKernelEx(&cfg, kern, io, (const float2*)h->fold.Hc, (const float2*)h->tw1, (const float2*)h->tw2,
                                (const float2*)h->fold.gc, (const float2*)h->fold.twc, (const float2*)h->fold.twm, nblocks, dtrace));
    count_launch();
    if (want_trace) {                                           // debug only: synchronous dump of the per-phase cycle table
        std::vector<long long> tr(trace_n);
        RRC_CUDA(cudaStreamSynchronize(st));
        RRC_CUDA(cudaMemcpy(tr.data(), dtrace, trace_n * 8, cudaMemcpyDeviceToHost));
        cudaFree(dtrace);
        static const char* names[] = {"A: loads, combine, DFT32, tw", "cluster wait (prev u read)", "A: store", "barrier 1", "B", "C + fold",
                                      "barrier 2", "inverse 2048 (warps 0-1) / idle", "cluster sync", "combine + store"};
        static int dumps = 0;
        if (nblocks > ncl * 8 && dumps++ < 1) {
            for (int b = 0; b < FOLD_TRACE_BLK; ++b) {
                long long t0 = tr[(size_t)(b * 16) * FOLD_NSTAMP], tend = 0;
                for (int w = 0; w < 16; ++w) { t0 = std::min(t0, tr[(size_t)(b * 16 + w) * FOLD_NSTAMP]); tend = std::max(tend, tr[(size_t)(b * 16 + w) * FOLD_NSTAMP + 10]); }
                fprintf(stderr, "fold trace block iter %d: total %lld cycles\n", b + 2, tend - t0);
                for (int p = 0; p < 10; ++p) {
                    std::vector<long long> d;
                    for (int w = 0; w < 16; ++w) d.push_back(tr[(size_t)(b * 16 + w) * FOLD_NSTAMP + p + 1] - tr[(size_t)(b * 16 + w) * FOLD_NSTAMP + p]);
                    std::sort(d.begin(), d.end());
                    fprintf(stderr, "   %-34s min %6lld  med %6lld  max %6lld\n", names[p], d[0], d[8], d[15]);
                }
            }
        }
    }
    return RRC_OK;
}

int build_fold_tables(rrc_fftfilt* h, int nc) {
    std::vector<float2> Hc, gc, twc, twm;
    fftf::build_fold_tables(h->taps_host.data(), h->ntaps, nc, Hc, gc, twc, twm);
    auto up = [&](float2** d, const std::vector<float2>& v) -> cudaError_t {
        cudaError_t e = cudaMalloc((void**)d, v.size() * sizeof(float2));
        if (e != cudaSuccess) return e;
        return upload_sync(*d, v.data(), v.size() * sizeof(float2));
    };
    cudaError_t e;
    if ((e = up(&h->fold.Hc, Hc)) != cudaSuccess || (e = up(&h->fold.gc, gc)) != cudaSuccess ||
        (e = up(&h->fold.twc, twc)) != cudaSuccess || (e = up(&h->fold.twm, twm)) != cudaSuccess)
        return fail(RRC_ERR_CUDA, "FftFilter fold table upload failed: %s", cudaGetErrorString(e));
    h->fold.nc = nc;
    return RRC_OK;
}

int fold_nc(size_t ntaps) {
    if (ntaps <= 12289) return 1;                 // 16384-point block, valid fraction >= 25 %
    if (ntaps <= 49153) return 4;                 // 65536-point block
    return 0;
}

}  // namespace

int fold_supported(const rrc_fftfilt* h, size_t deci) {
    if (deci != (size_t)fftf::FOLD_D || fold_nc(h->ntaps) == 0) return RRC_ERR_UNSUPPORTED;
    if (const char* v = getenv("RRC_FFTFILT_NO_FOLD")) if (atoi(v) != 0) return RRC_ERR_UNSUPPORTED;
    return RRC_OK;
}

int fold_launch(rrc_fftfilt* h, const float* in, size_t n, float* out, size_t n_out, size_t skip, cudaStream_t st) {
    const int nc = fold_nc(h->ntaps);
    if (h->fold.nc == 0) RRC_TRY(build_fold_tables(h, nc));
    FoldIO io;
    io.in = reinterpret_cast<const float2*>(in);
    io.hist = h->hist_ext ? h->hist_ext : h->hist[h->cur];
    io.out = reinterpret_cast<float2*>(out);
    io.n_in = (long long)n;
    io.n_out = (long long)n_out;
    io.T1_total = h->T1;
    io.T1eff = (h->T1 + 7) & ~7;
    io.V = nc * fftk::N - io.T1eff;
    io.r = (int)(skip % fftf::FOLD_D);
    io.jbias = (long long)(skip / fftf::FOLD_D);
    io.in_u8 = h->in_u8;
    io.epi = h->epi;
    io.hist_next = h->T1 > 0 ? h->hist[h->cur ^ 1] : nullptr;
    if ((long long)n <= io.r) return RRC_ERR_UNSUPPORTED;      // nothing to launch: the caller updates the history
    const long long nblocks = ((long long)n - io.r + io.V - 1) / io.V;
    return nc == 1 ? launch_fold<1>(h, io, nblocks, st) : launch_fold<4>(h, io, nblocks, st);
}

void fold_destroy(rrc_fftfilt* h) {
    cudaFree(h->fold.Hc); cudaFree(h->fold.gc); cudaFree(h->fold.twc); cudaFree(h->fold.twm);
    h->fold = rrc_fold_tables();
}

}  // namespace rrc
