// fir_tcc.cu — launchers of the complex-tap tensor-core FIR kernel (fir_tcc.cuh); own translation unit (compile time).
#include <algorithm>

#include "common.cuh"
#include "fir_tcc.cuh"

namespace rrc {
namespace {

template <int KS, bool DEMOD, int D>
int launch_tcc_k2(const FirTcGeom& g, const FirTccArgs& a, cudaStream_t st) {
    auto k = fir_tcc_kernel<KS, DEMOD, D>;
    constexpr size_t smem = fir_tc1_smem(KS, DEMOD) + (size_t)KS * 512;       // a second set of B fragments
    static int cache[16] = {};
    int dev = (g.device < 0 || g.device >= 16) ? 0 : g.device;
    if (cache[dev] == 0) {
        RRC_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        int per_sm = 0;
        RRC_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k, FIR_TC_THREADS, smem));
        cache[dev] = per_sm > 0 ? per_sm : -1;
    }
    if (cache[dev] < 1) return fail(RRC_ERR_CUDA, "fir_tcc: kernel does not fit an SM");
    const long long cap = (long long)sm_count(g.device) * cache[dev];
    const long long ctas = (a.total_tiles + FIR_TC_THREADS / 32 - 1) / (FIR_TC_THREADS / 32);
    k<<<(unsigned)std::min<long long>(ctas, cap), FIR_TC_THREADS, smem, st>>>(a);
    RRC_CHECK_LAUNCH();
    count_launch();
    return RRC_OK;
}
template <int KS, int D>
int launch_tcc_k(const FirTcGeom& g, const FirTccArgs& a, bool demod, cudaStream_t st) {
    return demod ? launch_tcc_k2<KS, true, D>(g, a, st) : launch_tcc_k2<KS, false, D>(g, a, st);
}
template <int D>
int launch_tcc_even(const FirTcGeom& g, const FirTccArgs& a, bool demod, cudaStream_t st) {
    switch (g.KS) {
    case 2: return launch_tcc_k<2, D>(g, a, demod, st);
    case 4: return launch_tcc_k<4, D>(g, a, demod, st);
    case 6: return launch_tcc_k<6, D>(g, a, demod, st);
    case 8: return launch_tcc_k<8, D>(g, a, demod, st);
    case 10: return launch_tcc_k<10, D>(g, a, demod, st);
    case 12: return launch_tcc_k<12, D>(g, a, demod, st);
    case 14: return launch_tcc_k<14, D>(g, a, demod, st);
    case 16: return launch_tcc_k<16, D>(g, a, demod, st);
    case 18: return launch_tcc_k<18, D>(g, a, demod, st);
    case 20: return launch_tcc_k<20, D>(g, a, demod, st);
    default: return fail(RRC_ERR_INVALID, "fir_tcc: no deci-%d kernel for %d k-steps", D, g.KS);
    }
}

}  // namespace

int fir_tcc_launch(const FirTcGeom& g, const FirTccArgs& a, bool demod, cudaStream_t st) {
    if (g.deci == 2) return launch_tcc_even<2>(g, a, demod, st);
    if (g.deci == 4) return launch_tcc_even<4>(g, a, demod, st);
    if (g.deci == 8) return launch_tcc_even<8>(g, a, demod, st);
    if (g.deci != 1) return fail(RRC_ERR_INVALID, "fir_tcc: deci %d", g.deci);
    switch (g.KS) {
    case 2: return launch_tcc_k<2, 1>(g, a, demod, st);
    case 3: return launch_tcc_k<3, 1>(g, a, demod, st);
    case 4: return launch_tcc_k<4, 1>(g, a, demod, st);
    case 5: return launch_tcc_k<5, 1>(g, a, demod, st);
    case 6: return launch_tcc_k<6, 1>(g, a, demod, st);
    case 7: return launch_tcc_k<7, 1>(g, a, demod, st);
    case 8: return launch_tcc_k<8, 1>(g, a, demod, st);
    case 9: return launch_tcc_k<9, 1>(g, a, demod, st);
    case 10: return launch_tcc_k<10, 1>(g, a, demod, st);
    case 11: return launch_tcc_k<11, 1>(g, a, demod, st);
    case 12: return launch_tcc_k<12, 1>(g, a, demod, st);
    case 13: return launch_tcc_k<13, 1>(g, a, demod, st);
    case 14: return launch_tcc_k<14, 1>(g, a, demod, st);
    case 15: return launch_tcc_k<15, 1>(g, a, demod, st);
    case 16: return launch_tcc_k<16, 1>(g, a, demod, st);
    case 17: return launch_tcc_k<17, 1>(g, a, demod, st);
    case 18: return launch_tcc_k<18, 1>(g, a, demod, st);
    case 19: return launch_tcc_k<19, 1>(g, a, demod, st);
    case 20: return launch_tcc_k<20, 1>(g, a, demod, st);
    default: return fail(RRC_ERR_INVALID, "fir_tcc: no kernel for %d k-steps", g.KS);
    }
}

}  // namespace rrc
