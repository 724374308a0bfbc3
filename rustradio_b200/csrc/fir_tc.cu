// fir_tc.cu — launchers of the tensor-core FIR kernels (fir_tc.cuh); its own translation unit so that the ~70 kernel
// instantiations compile next to fir.cu instead of after it.  Geometry comes from plan_tc in fir.cu.
#include <algorithm>

#include "common.cuh"
#include "fir_tc.cuh"

namespace rrc {
namespace {

// Opt-in shared memory + occupancy of one kernel instantiation, queried once per device (the launchers run on every
// work() call of a block).
template <typename K>
int ctas_per_sm(K k, size_t smem, int device, int* cache) {
    if (device < 0 || device >= 16) device = 0;
    if (cache[device] == 0) {
        if (cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return -1;
        int per_sm = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k, FIR_TC_THREADS, smem) != cudaSuccess) return -1;
        cache[device] = per_sm > 0 ? per_sm : -1;
    }
    return cache[device];
}

template <int NTILE, bool DEMOD, int NLD>
int launch_tc_k(const FirTcGeom& g, const FirTcArgs& a, cudaStream_t st) {
    auto k = fir_tc_kernel<NTILE, DEMOD, NLD>;
    RRC_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)g.smem));
    int per_sm = 0;
    RRC_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k, FIR_TC_THREADS, g.smem));
    if (per_sm < 1) return fail(RRC_ERR_CUDA, "fir_tc: kernel does not fit an SM (%zu bytes of shared memory)", g.smem);
    const long long cap = (long long)sm_count(g.device) * per_sm;
    const long long ctas = (a.total_tiles + FIR_TC_THREADS / 32 - 1) / (FIR_TC_THREADS / 32);
    const unsigned grid = (unsigned)std::min<long long>(ctas, cap);
    k<<<grid, FIR_TC_THREADS, g.smem, st>>>(a);
    RRC_CHECK_LAUNCH();
    count_launch();
    return RRC_OK;
}
template <bool DEMOD>
int launch_tc(const FirTcGeom& g, const FirTcArgs& a, cudaStream_t st) {
    switch (g.ntile * 100 + g.nld) {
    case 109: return launch_tc_k<1, DEMOD, 9>(g, a, st);
    case 114: return launch_tc_k<1, DEMOD, 14>(g, a, st);
    case 209: return launch_tc_k<2, DEMOD, 9>(g, a, st);
    case 214: return launch_tc_k<2, DEMOD, 14>(g, a, st);
    case 409: return launch_tc_k<4, DEMOD, 9>(g, a, st);
    case 414: return launch_tc_k<4, DEMOD, 14>(g, a, st);
    default: return fail(RRC_ERR_INVALID, "fir_tc: no kernel for ntile %d nld %d", g.ntile, g.nld);
    }
}


template <int KS, bool DEMOD, bool U8, int D>
int launch_tc1_k2(const FirTcGeom& g, const FirTc1Args& a, cudaStream_t st) {
    auto k = fir_tc1_kernel<KS, DEMOD, U8, D>;
    constexpr size_t smem = fir_tc1_smem(KS, DEMOD);
    static int cache[16] = {};
    const int per_sm = ctas_per_sm(k, smem, g.device, cache);
    if (per_sm < 1) return fail(RRC_ERR_CUDA, "fir_tc1: kernel does not fit an SM");
    const long long cap = (long long)sm_count(g.device) * per_sm;
    const long long ctas = (a.total_tiles + FIR_TC_THREADS / 32 - 1) / (FIR_TC_THREADS / 32);
    k<<<(unsigned)std::min<long long>(ctas, cap), FIR_TC_THREADS, smem, st>>>(a);
    RRC_CHECK_LAUNCH();
    count_launch();
    return RRC_OK;
}
template <int KS, int D>
int launch_tc1_k(const FirTcGeom& g, const FirTc1Args& a, bool demod, cudaStream_t st) {
    if (demod) return a.in_u8 ? launch_tc1_k2<KS, true, true, D>(g, a, st) : launch_tc1_k2<KS, true, false, D>(g, a, st);
    return a.in_u8 ? launch_tc1_k2<KS, false, true, D>(g, a, st) : launch_tc1_k2<KS, false, false, D>(g, a, st);
}
template <int D>
int launch_tc1_even(const FirTcGeom& g, const FirTc1Args& a, bool demod, cudaStream_t st) {
    switch (g.KS) {
    case 2: return launch_tc1_k<2, D>(g, a, demod, st);
    case 4: return launch_tc1_k<4, D>(g, a, demod, st);
    case 6: return launch_tc1_k<6, D>(g, a, demod, st);
    case 8: return launch_tc1_k<8, D>(g, a, demod, st);
    case 10: return launch_tc1_k<10, D>(g, a, demod, st);
    case 12: return launch_tc1_k<12, D>(g, a, demod, st);
    case 14: return launch_tc1_k<14, D>(g, a, demod, st);
    case 16: return launch_tc1_k<16, D>(g, a, demod, st);
    case 18: return launch_tc1_k<18, D>(g, a, demod, st);
    case 20: return launch_tc1_k<20, D>(g, a, demod, st);
    default: return fail(RRC_ERR_INVALID, "fir_tc1: no deci-%d kernel for %d k-steps", D, g.KS);
    }
}

template <int KS, int D>
int launch_tcf_k(const FirTcGeom& g, const FirTcfArgs& a, cudaStream_t st) {
    auto k = fir_tcf_kernel<KS, D>;
    constexpr size_t smem = fir_tcf_smem(KS);
    static int cache[16] = {};
    const int per_sm = ctas_per_sm(k, smem, g.device, cache);
    if (per_sm < 1) return fail(RRC_ERR_CUDA, "fir_tcf: kernel does not fit an SM");
    const long long cap = (long long)sm_count(g.device) * per_sm;
    const long long ctas = (a.total_tiles + FIR_TC_THREADS / 32 - 1) / (FIR_TC_THREADS / 32);
    k<<<(unsigned)std::min<long long>(ctas, cap), FIR_TC_THREADS, smem, st>>>(a);
    RRC_CHECK_LAUNCH();
    count_launch();
    return RRC_OK;
}
template <int D>
int launch_tcf_even(const FirTcGeom& g, const FirTcfArgs& a, cudaStream_t st) {
    switch (g.KS) {
    case 2: return launch_tcf_k<2, D>(g, a, st);
    case 4: return launch_tcf_k<4, D>(g, a, st);
    case 6: return launch_tcf_k<6, D>(g, a, st);
    case 8: return launch_tcf_k<8, D>(g, a, st);
    case 10: return launch_tcf_k<10, D>(g, a, st);
    case 12: return launch_tcf_k<12, D>(g, a, st);
    case 14: return launch_tcf_k<14, D>(g, a, st);
    case 16: return launch_tcf_k<16, D>(g, a, st);
    case 18: return launch_tcf_k<18, D>(g, a, st);
    case 20: return launch_tcf_k<20, D>(g, a, st);
    default: return fail(RRC_ERR_INVALID, "fir_tcf: no deci-%d kernel for %d k-steps", D, g.KS);
    }
}

}  // namespace

int fir_tcf_launch(const FirTcGeom& g, const FirTcfArgs& a, cudaStream_t st) {
    if (g.deci == 2) return launch_tcf_even<2>(g, a, st);
    if (g.deci == 4) return launch_tcf_even<4>(g, a, st);
    if (g.deci == 8) return launch_tcf_even<8>(g, a, st);
    if (g.deci != 1) return fail(RRC_ERR_INVALID, "fir_tcf: deci %d", g.deci);
    switch (g.KS) {
    case 2: return launch_tcf_k<2, 1>(g, a, st);
    case 3: return launch_tcf_k<3, 1>(g, a, st);
    case 4: return launch_tcf_k<4, 1>(g, a, st);
    case 5: return launch_tcf_k<5, 1>(g, a, st);
    case 6: return launch_tcf_k<6, 1>(g, a, st);
    case 7: return launch_tcf_k<7, 1>(g, a, st);
    case 8: return launch_tcf_k<8, 1>(g, a, st);
    case 9: return launch_tcf_k<9, 1>(g, a, st);
    case 10: return launch_tcf_k<10, 1>(g, a, st);
    case 11: return launch_tcf_k<11, 1>(g, a, st);
    case 12: return launch_tcf_k<12, 1>(g, a, st);
    case 13: return launch_tcf_k<13, 1>(g, a, st);
    case 14: return launch_tcf_k<14, 1>(g, a, st);
    case 15: return launch_tcf_k<15, 1>(g, a, st);
    case 16: return launch_tcf_k<16, 1>(g, a, st);
    case 17: return launch_tcf_k<17, 1>(g, a, st);
    case 18: return launch_tcf_k<18, 1>(g, a, st);
    case 19: return launch_tcf_k<19, 1>(g, a, st);
    case 20: return launch_tcf_k<20, 1>(g, a, st);
    default: return fail(RRC_ERR_INVALID, "fir_tcf: no kernel for %d k-steps", g.KS);
    }
}

int fir_tc_launch(const FirTcGeom& g, const FirTcArgs& a, bool demod, cudaStream_t st) {
    return demod ? launch_tc<true>(g, a, st) : launch_tc<false>(g, a, st);
}

int fir_tc1_launch(const FirTcGeom& g, const FirTc1Args& a, bool demod, cudaStream_t st) {
    if (g.deci == 2) return launch_tc1_even<2>(g, a, demod, st);
    if (g.deci == 4) return launch_tc1_even<4>(g, a, demod, st);
    if (g.deci == 8) return launch_tc1_even<8>(g, a, demod, st);
    if (g.deci != 1) return fail(RRC_ERR_INVALID, "fir_tc1: deci %d", g.deci);
    switch (g.KS) {
    case 2: return launch_tc1_k<2, 1>(g, a, demod, st);
    case 3: return launch_tc1_k<3, 1>(g, a, demod, st);
    case 4: return launch_tc1_k<4, 1>(g, a, demod, st);
    case 5: return launch_tc1_k<5, 1>(g, a, demod, st);   // (4 CTAs per SM at 64 registers measured slower: 77.6 vs 62.8 us on config 1)
    case 6: return launch_tc1_k<6, 1>(g, a, demod, st);
    case 7: return launch_tc1_k<7, 1>(g, a, demod, st);
    case 8: return launch_tc1_k<8, 1>(g, a, demod, st);
    case 9: return launch_tc1_k<9, 1>(g, a, demod, st);
    case 10: return launch_tc1_k<10, 1>(g, a, demod, st);
    case 11: return launch_tc1_k<11, 1>(g, a, demod, st);
    case 12: return launch_tc1_k<12, 1>(g, a, demod, st);
    case 13: return launch_tc1_k<13, 1>(g, a, demod, st);
    case 14: return launch_tc1_k<14, 1>(g, a, demod, st);
    case 15: return launch_tc1_k<15, 1>(g, a, demod, st);
    case 16: return launch_tc1_k<16, 1>(g, a, demod, st);
    case 17: return launch_tc1_k<17, 1>(g, a, demod, st);
    case 18: return launch_tc1_k<18, 1>(g, a, demod, st);
    case 19: return launch_tc1_k<19, 1>(g, a, demod, st);
    case 20: return launch_tc1_k<20, 1>(g, a, demod, st);
    default: return fail(RRC_ERR_INVALID, "fir_tc1: no kernel for %d k-steps", g.KS);
    }
}

}  // namespace rrc
