// CPU emulation of the FftFilter CUDA kernel's five phases (fftfilt_core.cuh):
// every "thread" of the 512-thread CTA is run in a loop, phase by phase, with
// a plain array standing in for shared memory.  Lets the index math, swizzle
// and twiddle logic be checked against the oracle without a GPU.
// Built by tests/test_emul.py with `nvcc -x cu` (host code only).
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../../rustradio_b200/csrc/fftfilt_tables.hpp"

using namespace rrc::fftk;

extern "C" int emul_fftfilt(const float* taps, long long ntaps, const float* in, long long n,
                            const float* hist /* (ntaps-1) c32 or NULL = zeros */, float* out,
                            long long deci, long long skip, long long n_out) {
    const int T1_total = (int)ntaps - 1;
    std::vector<float2> h(T1_total > 0 ? T1_total : 1, make_float2(0.f, 0.f));
    if (hist && T1_total > 0) memcpy(h.data(), hist, sizeof(float2) * T1_total);
    const long long part = ntaps <= 12289 ? ntaps : 8193;     // same split as fftfilt.cu
    long long shift = 0;
    const bool decim = !(deci == 1 && skip == 0);
    for (long long off = 0; off < ntaps; off += part) {
        const long long len = std::min(part, ntaps - off);
        std::vector<float2> Hp, tw1, tw2;
        build_tables(taps + 2 * off, (size_t)len, Hp, tw1, tw2);
        BlockIO io;
        io.in = reinterpret_cast<const float2*>(in);
        io.hist = h.data();
        io.out = reinterpret_cast<float2*>(out);
        io.n_in = n; io.n_out = n_out; io.T1 = (int)len - 1; io.V = N - io.T1; io.deci = (int)deci; io.skip = skip;
        io.T1_total = T1_total; io.shift = shift;
        std::vector<float2> sm(SMEM_ELEMS), hres(HRES_ELEMS);
        for (int t = 0; t < NT; ++t) load_hres(t, Hp.data(), hres.data());
        const long long nblocks = (n + io.V - 1) / io.V;
        // RRC_EMUL_FFTFILT_MODE: 0 = table twiddles in B/B' (default), 1 = twiddles from powers in C
        // (TWC), 2 = TWC + staged input (stage_input / phase_a_staged).
        const char* me = getenv("RRC_EMUL_FFTFILT_MODE");
        const int mode = me ? atoi(me) : 0;
        if (mode == 4) {                                       // packed kernel (fftfilt_pk.cuh)
            namespace pk = rrc::fftp;
            std::vector<float2> Hq, tw2p, smp(pk::SMEM_WORDS), hresp(pk::HRES_WORDS);
            build_tables_pk(taps + 2 * off, (size_t)len, Hq, tw2p);
            for (int t = 0; t < NT; ++t) pk::load_hres(t, Hq.data(), hresp.data());
            std::vector<pk::C2> regs((size_t)NT * 16);
            auto save = [&](int t, const pk::C2 (&v)[16]) { for (int i = 0; i < 16; ++i) regs[(size_t)t * 16 + i] = v[i]; };
            auto load = [&](int t, pk::C2 (&v)[16]) { for (int i = 0; i < 16; ++i) v[i] = regs[(size_t)t * 16 + i]; };
            for (long long blk = 0; blk < nblocks; ++blk) {
                if (pk::bulk_ok(blk, io)) {                        // 32 bulk copies of 4 KiB, one per row n1
                    for (int n1 = 0; n1 < 32; ++n1) memcpy(smp.data() + n1 * pk::PP, io.in + pk::seg0_of(blk, io) + 512 * n1, sizeof(float2) * 512);
                } else {
                    for (int t = 0; t < NT; ++t) pk::stage_fallback(t, blk, io, smp.data());
                }
                for (int t = 0; t < NT; ++t) { pk::C2 v[16]; pk::phase_a_compute(t, tw1.data(), smp.data(), v); save(t, v); }
                for (int t = 0; t < NT; ++t) { pk::C2 v[16]; load(t, v); pk::phase_a_store(t, smp.data(), v); }
                for (int t = 0; t < NT; ++t) { pk::C2 v[16]; pk::phase_b_compute(t, tw2p.data(), smp.data(), v); save(t, v); }
                for (int t = 0; t < NT; ++t) { pk::C2 v[16]; load(t, v); pk::phase_b_store(t, smp.data(), v); }
                for (int t = 0; t < NT; ++t) pk::phase_c(t, Hq.data(), hresp.data(), smp.data());
                for (int t = 0; t < NT; ++t) { pk::C2 v[16]; pk::phase_bi_compute(t, tw2p.data(), smp.data(), v); save(t, v); }
                for (int t = 0; t < NT; ++t) { pk::C2 v[16]; load(t, v); pk::phase_bi_store(t, smp.data(), v); }
                for (int t = 0; t < NT; ++t) {
                    if (off == 0) { if (decim) pk::phase_ai<true, false>(t, blk, io, tw1.data(), smp.data()); else pk::phase_ai<false, false>(t, blk, io, tw1.data(), smp.data()); }
                    else          { if (decim) pk::phase_ai<true, true>(t, blk, io, tw1.data(), smp.data()); else pk::phase_ai<false, true>(t, blk, io, tw1.data(), smp.data()); }
                }
            }
            shift += len;
            continue;
        }
        for (long long blk = 0; blk < nblocks; ++blk) {
            if (mode == 2) {
                for (int t = 0; t < NT; ++t) stage_input(t, blk, io, sm.data());
                for (int t = 0; t < NT; ++t) phase_a_staged(t, tw1.data(), sm.data());
            } else if (mode == 3) {                            // linear (TMA) staging, fftfilt_tma_kernel
                if (stage_linear_bulk_ok(blk, io)) memcpy(sm.data(), io.in + stage_linear_seg0(blk, io), sizeof(float2) * N);   // cp.async.bulk
                else for (int t = 0; t < NT; ++t) stage_linear_fallback(t, blk, io, sm.data());
                std::vector<float2> regs((size_t)NT * 32);
                for (int t = 0; t < NT; ++t) { float2 v[32]; phase_a_linear_load(t, sm.data(), v); phase_a_linear_compute(t, tw1.data(), v); memcpy(&regs[(size_t)t * 32], v, sizeof v); }
                for (int t = 0; t < NT; ++t) { float2 v[32]; memcpy(v, &regs[(size_t)t * 32], sizeof v); phase_a_linear_store(t, sm.data(), v); }
            } else {
                for (int t = 0; t < NT; ++t) phase_a(t, blk, io, tw1.data(), sm.data());
            }
            if (mode == 0 || mode == 3) {
                for (int t = 0; t < NT; ++t) phase_mid_b(t, tw2.data(), sm.data());
                for (int t = 0; t < NT; ++t) phase_mid_c(t, Hp.data(), hres.data(), sm.data());
                for (int t = 0; t < NT; ++t) phase_mid_bi(t, tw2.data(), sm.data());
            } else {
                for (int t = 0; t < NT; ++t) phase_mid_b<false>(t, tw2.data(), sm.data());
                for (int t = 0; t < NT; ++t) phase_mid_c<true>(t, Hp.data(), hres.data(), sm.data(), NoTurn(), tw2.data());
                for (int t = 0; t < NT; ++t) phase_mid_bi<false>(t, tw2.data(), sm.data());
            }
            for (int t = 0; t < NT; ++t) {
                if (off == 0) { if (decim) phase_ai<true, false>(t, blk, io, tw1.data(), sm.data()); else phase_ai<false, false>(t, blk, io, tw1.data(), sm.data()); }
                else          { if (decim) phase_ai<true, true>(t, blk, io, tw1.data(), sm.data()); else phase_ai<false, true>(t, blk, io, tw1.data(), sm.data()); }
            }
        }
        shift += len;
    }
    return 0;
}

// Real-stream mode (FftFilterFloat): f32 in / hist / out, two real blocks per complex transform.
extern "C" int emul_fftfilt_real(const float* taps_f32, long long ntaps, const float* in, long long n,
                                 const float* hist /* (ntaps-1) f32 or NULL */, float* out) {
    const int T1_total = (int)ntaps - 1;
    std::vector<float> h(T1_total > 0 ? T1_total : 1, 0.f);
    if (hist && T1_total > 0) memcpy(h.data(), hist, sizeof(float) * T1_total);
    std::vector<float> ct(2 * ntaps, 0.f);
    for (long long i = 0; i < ntaps; ++i) ct[2 * i] = taps_f32[i];
    const long long part = ntaps <= 12289 ? ntaps : 8193;
    long long shift = 0;
    for (long long off = 0; off < ntaps; off += part) {
        const long long len = std::min(part, ntaps - off);
        std::vector<float2> Hp, tw1, tw2;
        build_tables(ct.data() + 2 * off, (size_t)len, Hp, tw1, tw2);
        BlockIO io;
        io.in = reinterpret_cast<const float2*>(in);
        io.hist = reinterpret_cast<const float2*>(h.data());
        io.out = reinterpret_cast<float2*>(out);
        io.n_in = n; io.n_out = n; io.T1 = (int)len - 1; io.V = N - io.T1; io.deci = 1; io.skip = 0;
        io.T1_total = T1_total; io.shift = shift; io.real = 1;
        std::vector<float2> sm(SMEM_ELEMS), hres(HRES_ELEMS);
        for (int t = 0; t < NT; ++t) load_hres(t, Hp.data(), hres.data());
        const long long nreal = (n + io.V - 1) / io.V, nblocks = (nreal + 1) / 2;
        for (long long blk = 0; blk < nblocks; ++blk) {
            for (int t = 0; t < NT; ++t) phase_a(t, blk, io, tw1.data(), sm.data());
            for (int t = 0; t < NT; ++t) phase_mid_b(t, tw2.data(), sm.data());
            for (int t = 0; t < NT; ++t) phase_mid_c(t, Hp.data(), hres.data(), sm.data());
            for (int t = 0; t < NT; ++t) phase_mid_bi(t, tw2.data(), sm.data());
            for (int t = 0; t < NT; ++t) {
                if (off == 0) phase_ai<false, false>(t, blk, io, tw1.data(), sm.data());
                else phase_ai<false, true>(t, blk, io, tw1.data(), sm.data());
            }
        }
        shift += len;
    }
    return 0;
}

// Same for the 1024-thread x 16-point variant (fftfilt16_core.cuh).
extern "C" int emul_fftfilt16(const float* taps, long long ntaps, const float* in, long long n,
                              const float* hist, float* out, long long deci, long long skip, long long n_out) {
    namespace k16 = rrc::fftk16;
    const int T1_total = (int)ntaps - 1;
    std::vector<float2> h(T1_total > 0 ? T1_total : 1, make_float2(0.f, 0.f));
    if (hist && T1_total > 0) memcpy(h.data(), hist, sizeof(float2) * T1_total);
    const long long part = ntaps <= 12289 ? ntaps : 8193;
    long long shift = 0;
    const bool decim = !(deci == 1 && skip == 0);
    for (long long off = 0; off < ntaps; off += part) {
        const long long len = std::min(part, ntaps - off);
        std::vector<float2> Hd, tw1, tw2, tw3;
        build_tables16(taps + 2 * off, (size_t)len, Hd, tw1, tw2, tw3);
        BlockIO io;
        io.in = reinterpret_cast<const float2*>(in);
        io.hist = h.data();
        io.out = reinterpret_cast<float2*>(out);
        io.n_in = n; io.n_out = n_out; io.T1 = (int)len - 1; io.V = N - io.T1; io.deci = (int)deci; io.skip = skip;
        io.T1_total = T1_total; io.shift = shift;
        std::vector<float2> sm(k16::SMEM16_ELEMS), hres(k16::HRES16_ELEMS);
        for (int t = 0; t < k16::NT16; ++t) k16::load_hres(t, Hd.data(), hres.data());
        const long long nblocks = (n + io.V - 1) / io.V;
        for (long long blk = 0; blk < nblocks; ++blk) {
            for (int t = 0; t < k16::NT16; ++t) k16::phase_a(t, blk, io, tw1.data(), sm.data());
            for (int t = 0; t < k16::NT16; ++t) k16::phase_b(t, tw2.data(), sm.data());
            for (int t = 0; t < k16::NT16; ++t) k16::phase_c(t, tw3.data(), sm.data());
            for (int t = 0; t < k16::NT16; ++t) k16::phase_d(t, Hd.data(), hres.data(), sm.data());
            for (int t = 0; t < k16::NT16; ++t) k16::phase_ci(t, tw3.data(), sm.data());
            for (int t = 0; t < k16::NT16; ++t) k16::phase_bi(t, tw2.data(), sm.data());
            for (int t = 0; t < k16::NT16; ++t) {
                if (off == 0) { if (decim) k16::phase_ai<true, false>(t, blk, io, tw1.data(), sm.data()); else k16::phase_ai<false, false>(t, blk, io, tw1.data(), sm.data()); }
                else          { if (decim) k16::phase_ai<true, true>(t, blk, io, tw1.data(), sm.data()); else k16::phase_ai<false, true>(t, blk, io, tw1.data(), sm.data()); }
            }
        }
        shift += len;
    }
    return 0;
}

// Decimate-by-8 fold kernel (fftfilt_fold_core.cuh): nc = 1 (16384-point) or nc = 4 (65536-point,
// the four CTAs of a cluster emulated one after the other with four shared-memory arrays).
template <int NC>
static int emul_fold_t(const float* taps, long long ntaps, const float* in, long long n, const float* hist,
                       float* out, long long skip, long long n_out) {
    namespace ff = rrc::fftf;
    const int T1_total = (int)ntaps - 1;
    std::vector<float2> h(T1_total > 0 ? T1_total : 1, make_float2(0.f, 0.f));
    if (hist && T1_total > 0) memcpy(h.data(), hist, sizeof(float2) * T1_total);
    std::vector<float2> Hp0, tw1, tw2, Hc, gc, twc, twm;
    build_tables(taps, (size_t)std::min<long long>(ntaps, 8193), Hp0, tw1, tw2);      // tw1, tw2 only
    ff::build_fold_tables(taps, (size_t)ntaps, NC, Hc, gc, twc, twm);
    ff::FoldIO io;
    io.in = reinterpret_cast<const float2*>(in); io.hist = h.data(); io.out = reinterpret_cast<float2*>(out);
    io.n_in = n; io.n_out = n_out; io.T1_total = T1_total; io.T1eff = (T1_total + 7) & ~7;
    io.V = NC * N - io.T1eff; io.r = (int)(skip % 8); io.jbias = skip / 8;
    if (n <= io.r) return 0;
    const long long nblocks = (n - io.r + io.V - 1) / io.V;
    const int U_OFF = 4096;
    std::vector<std::vector<float2>> sm(NC, std::vector<float2>(SMEM_ELEMS)), hres(NC, std::vector<float2>(HRES_ELEMS));
    for (int c = 0; c < NC; ++c)
        for (int t = 0; t < NT; ++t) load_hres(t, Hc.data() + (size_t)c * N, hres[c].data());
    for (long long blk = 0; blk < nblocks; ++blk) {
        for (int c = 0; c < NC; ++c) {
            float2* s = sm[c].data();
            const float2* Hp = Hc.data() + (size_t)c * N;
            for (int t = 0; t < NT; ++t) ff::phase_a<NC>(t, c, blk, io, tw1.data(), gc.data() + c * 512, twc.data() + c * 32, s);
            for (int t = 0; t < NT; ++t) phase_mid_b(t, tw2.data(), s);
            for (int t = 0; t < NT; ++t) ff::phase_c_fold(t, Hp, hres[c].data(), s);
            std::vector<float2> regs(64 * 32);
            for (int t = 0; t < 64; ++t) { float2 v[32]; ff::inv1_load(t, s, v); memcpy(&regs[t * 32], v, sizeof v); }
            for (int t = 0; t < 64; ++t) { float2 v[32]; memcpy(v, &regs[t * 32], sizeof v); ff::inv1_compute_store(t, tw1.data(), v, s); }
            for (int t = 0; t < 64; ++t) { float2 v[32]; ff::inv2_load(t, s, v); memcpy(&regs[t * 32], v, sizeof v); }
            for (int t = 0; t < 64; ++t) { float2 v[32]; memcpy(v, &regs[t * 32], sizeof v); ff::inv2_compute_store(t, v, s + U_OFF); }
        }
        const float2* uc[NC];
        for (int c = 0; c < NC; ++c) uc[c] = sm[c].data() + U_OFF;
        for (int c = 0; c < NC; ++c)
            for (int t = 0; t < NT; ++t) ff::combine_store<NC>(t, c, blk, io, uc, twm.data());
    }
    return 0;
}

extern "C" int emul_fftfilt_fold(int nc, const float* taps, long long ntaps, const float* in, long long n,
                                 const float* hist, float* out, long long skip, long long n_out) {
    return nc == 1 ? emul_fold_t<1>(taps, ntaps, in, n, hist, out, skip, n_out)
                   : emul_fold_t<4>(taps, ntaps, in, n, hist, out, skip, n_out);
}

// Polyphase decimating kernel (fftfilt_poly_core.cuh): D forward transforms, the running sum over the
// branches in a per-thread accumulator array (tensor memory on the GPU), one inverse transform.  Pairs of branches are
// fetched with 128-bit loads where the alignment allows, the second one parked in the stash.
extern "C" int emul_fftfilt_poly(const float* taps, long long ntaps, const float* in, long long n, const float* hist,
                                 float* out, long long deci, long long skip, long long n_out) {
    namespace fp = rrc::fftp;
    const int T1_total = (int)ntaps - 1;
    std::vector<float2> h(T1_total > 0 ? T1_total : 1, make_float2(0.f, 0.f));
    if (hist && T1_total > 0) memcpy(h.data(), hist, sizeof(float2) * T1_total);
    std::vector<float2> Hp0, tw1, tw2, Hph;
    build_tables(taps, (size_t)std::min<long long>(ntaps, 8193), Hp0, tw1, tw2);      // tw1, tw2 only
    const int smod = (int)(skip % deci);
    fp::build_poly_tables(taps, (size_t)ntaps, (int)deci, smod, Hph);
    fp::PolyIO io;
    io.b.in = reinterpret_cast<const float2*>(in); io.b.hist = h.data(); io.b.out = reinterpret_cast<float2*>(out);
    io.b.n_in = n; io.b.n_out = n_out; io.b.T1_total = T1_total;
    io.b.T1 = fp::poly_T1(ntaps, (int)deci, smod); io.b.V = N - io.b.T1; io.b.shift = 0; io.b.deci = 1; io.b.skip = 0;
    io.D = (int)deci; io.sbase = skip - smod;
    if (n_out <= 0) return 0;
    const long long nblocks = (n_out + io.b.V - 1) / io.b.V;
    std::vector<float2> sm(SMEM_ELEMS), accv(512 * 32), stashv(512 * 32), regs(512 * 32);
    fp::HostAcc acc{accv.data()}, stash{stashv.data()};
    for (long long blk = 0; blk < nblocks; ++blk) {
        bool stashed = false;
        for (int r = 0; r < io.D; ++r) {
            const bool pair = !stashed && fp::poly_pair_ok(io, blk, r);
            for (int t = 0; t < NT; ++t) {
                float2 v[32];
                if (stashed) fp::poly_load_stash(t, v, stash);
                else if (pair) fp::poly_load_pair(t, blk, r, io, v, stash);
                else fp::poly_load(t, blk, r, io, v);
                phase_a_linear_compute(t, tw1.data(), v);
                memcpy(&regs[t * 32], v, sizeof v);
            }
            stashed = pair;
            for (int t = 0; t < NT; ++t) { float2 v[32]; memcpy(v, &regs[t * 32], sizeof v); phase_a_linear_store(t, sm.data(), v); }
            for (int t = 0; t < NT; ++t) phase_mid_b(t, tw2.data(), sm.data());
            for (int t = 0; t < NT; ++t) fp::phase_c_acc(t, Hph.data() + (size_t)r * N, Hph.data() + (size_t)io.D * N + (size_t)r * HRES_ELEMS, sm.data(), acc, r == 0, r == io.D - 1);
        }
        for (int t = 0; t < NT; ++t) phase_mid_bi(t, tw2.data(), sm.data());
        for (int t = 0; t < NT; ++t) phase_ai<false, false>(t, blk, io.b, tw1.data(), sm.data());
    }
    return 0;
}
