// fftfilt_core.cuh — one 16384-point overlap-save block of the FftFilter
// kernel, written as three barrier-separated phases that are pure functions of
// (thread id, shared memory, parameters).  __host__ __device__ so that the
// same code runs under the CPU emulator in tests/emul (index-math check
// without a GPU) and inside the CUDA kernel.
//
// Replaces Engine::run + sum_vec (rustradio src/fft_filter.rs:172-176,281-287):
// IFFT(FFT(x) * H), with 1/N folded into H like :153-162.
//
// Transform: N = 16384 = N1*N2*N3 = 32*32*16, 512 threads x 32 points.
//   n = n1*512 + n2*16 + n3,  k = k1 + 32*k2 + 1024*k3
//   A  : global -> DFT32 over n1 -> * W_N^{t*k1}            (t = n2*16+n3 = tid)
//   MID: tid = k1*16 + l (a half-warp owns one k1 plane of 32x16 points)
//        B : DFT32 over n2 -> * W_512^{n3*k2}      (column n3 = l)
//        C : DFT16 over n3 -> * H[k] -> IDFT16     (rows k2 = l, l+16)
//        B': * conj W_512^{n3*k2} -> IDFT32 over k2 (column n3 = l)
//        B->C->B' exchange data only inside the half-warp: __syncwarp(), no CTA barrier.
//   A' : * conj W_N^{t*k1} -> IDFT32 over k1 -> global      (tid = t)
// Shared-memory exchange buffer: plane k1 (32 rows x 16 columns of float2) at
// k1*544, row pitch 17 (one pad element per row):
//     phys(k1, r, c) = k1*544 + r*17 + c
// Column access (fixed c, lanes = rows, stride 17) and row access (fixed r,
// lanes = columns) are both bank-conflict free for 64-bit words, and every
// per-register offset is a compile-time immediate.  Every phase writes back
// exactly the words it read (in place), so there are only two CTA barriers
// per block: A -> MID and MID -> A'.
#pragma once
#include "fft_regs.cuh"
#include "epilogue.cuh"

namespace rrc { namespace fftk {

using namespace rrc::fftr;

constexpr int N = 16384;
constexpr int NT = 512;          // threads per CTA
constexpr int ROW_PITCH = 17;
constexpr int PLANE_PITCH = 32 * ROW_PITCH;        // 544
constexpr int SMEM_ELEMS = 32 * PLANE_PITCH;       // 17408 float2

struct BlockIO {
    const float2* in;        // this call's input samples, x[0..n_in)
    const float2* hist;      // previous T1 samples, hist[i] = x[i - T1]
    float2* out;             // y[0..n_out)
    long long n_in;          // valid input samples
    long long n_out;         // outputs to write
    int T1;                  // taps of THIS partition - 1
    int V;                   // valid outputs per block = N - T1
    int T1_total;            // ntaps - 1 of the whole filter = length of hist
    long long shift;         // input delay of this tap partition (p * partition length); 0 for a single partition
    int deci;                // output decimation (fused RationalResampler(1,deci)); 1 = none
    long long skip;          // first kept filter output index (decimation phase)
    int in_u8 = 0;           // 1: `in` is u8 I/Q pairs (RtlSdrDecode fused into the load); hist is always c32
    // 1: REAL stream (FftFilterFloat, src/fft_filter.rs:365-491) with real taps: in / hist / out are f32
    // arrays and kernel block b carries the two consecutive real blocks 2b (real part) and 2b+1
    // (imaginary part) through ONE complex transform — h real => h*(a + ib) = h*a + i h*b — so a real
    // stream costs half the transforms of the reference's widen -> complex filter -> .re.
    int real = 0;
    // Non-NULL: the kernel itself writes the history the NEXT call starts from (the last T1_total samples of
    // hist ++ in) into this buffer — done by the grid's last CTA before its first block, so a run() is ONE
    // launch (matters for small work() windows and for time-segment shards, where a step is ~0.25 ms).
    float2* hist_next = nullptr;
    // Store epilogue (epilogue.cuh): MultiplyConst / AddConst / ComplexToMag2 fused into the output store
    // (complex streams, single tap partition).  MAG2 makes `out` an f32 array.
    rrc::Epi epi;
};

// hist_next[i] = x[n_in - T1_total + i] over the concatenation (hist ++ in), i < T1_total.
RRC_HD void update_history(const BlockIO& io, int tid, int nthreads) {
    if (io.real) {
        const float* hc = reinterpret_cast<const float*>(io.hist);
        const float* in = reinterpret_cast<const float*>(io.in);
        float* hn = reinterpret_cast<float*>(io.hist_next);
        for (int i = tid; i < io.T1_total; i += nthreads) {
            const long long s = io.n_in - io.T1_total + i;
            hn[i] = s >= 0 ? in[s] : hc[s + io.T1_total];
        }
        return;
    }
    for (int i = tid; i < io.T1_total; i += nthreads) {
        const long long s = io.n_in - io.T1_total + i;
        io.hist_next[i] = s >= 0 ? ld_iq(io.in, s, io.in_u8) : io.hist[s + io.T1_total];
    }
}

// Sample g of a real stream with its carried history (g < 0) and zero fill past the end.
RRC_HD float fetch_real(const BlockIO& io, long long g) {
    if (g < 0) return g + io.T1_total >= 0 ? reinterpret_cast<const float*>(io.hist)[g + io.T1_total] : 0.f;
    return g < io.n_in ? reinterpret_cast<const float*>(io.in)[g] : 0.f;
}

RRC_HD int phys(int k1, int r, int c) { return k1 * PLANE_PITCH + r * ROW_PITCH + c; }

// FP-pipe "turn" policy.  Every phase below is  loads -> acquire() -> arithmetic -> release() -> stores.
// NoTurn: no-ops (plain kernel, CPU emulator).  The ping-pong kernel (fftfilt.cu, PingPong) passes a
// token between two groups of 8 warps so that one group's arithmetic burst runs while the other
// group's shared-memory / global burst is in flight, instead of all 16 warps doing the same kind of
// work in lock-step.
// acquire() returns 1.0f: the forward phases feed it to dit_g so that their arithmetic depends on a
// value produced after the barrier (see dit_g in fft_regs.cuh); the inverse phases start with a
// multiplication by a twiddle that is loaded after acquire().
struct NoTurn {
    RRC_HD float acquire() const { return 1.0f; }
    RRC_HD void release() const {}
};

#if defined(__CUDACC__)
// Ping-pong: the 16 warps form two groups of 8 (two warps of each group on every SM sub-partition)
// that pass an "FP turn" token through named barriers 1 and 2, so one group's arithmetic burst
// overlaps the other group's shared-memory / global burst.
struct PingPong {
    int g;               // group 0 or 1
    unsigned one_addr;   // shared-memory address of a float 1.0f
    __device__ __forceinline__ float acquire() const {
        float one;
        asm volatile("bar.sync %1, 512;\n\tld.volatile.shared.f32 %0, [%2];" : "=f"(one) : "r"(1 + g), "r"(one_addr) : "memory");
        return one;
    }
    __device__ __forceinline__ void release() const { asm volatile("bar.arrive %0, 512;" ::"r"(2 - g) : "memory"); }
};
#endif

#if defined(__CUDA_ARCH__)
#define RRC_SYNCWARP() __syncwarp()
#else
#define RRC_SYNCWARP() ((void)0)
#endif

// Powers p[k] = w^k, k = 0..31, depth <= 5 multiplications each.
RRC_HD void powers32(float2 w, float2 (&p)[32]) {
    float2 wp[5];
    wp[0] = w;
#pragma unroll
    for (int i = 1; i < 5; ++i) wp[i] = csqr(wp[i - 1]);
    p[0] = make_float2(1.f, 0.f);
#pragma unroll
    for (int k = 1; k < 32; ++k) {
        const int low = k & (-k);
        const int rest = k & (k - 1);
        const int b = low == 1 ? 0 : low == 2 ? 1 : low == 4 ? 2 : low == 8 ? 3 : 4;
        p[k] = rest == 0 ? wp[b] : cmul(p[rest], wp[b]);
    }
}

// Powers p[k] = w^k, k = 0..15, depth <= 4 multiplications each.
RRC_HD void powers16(float2 w, float2 (&p)[16]) {
    float2 wp[4];
    wp[0] = w;
#pragma unroll
    for (int i = 1; i < 4; ++i) wp[i] = csqr(wp[i - 1]);
    p[0] = make_float2(1.f, 0.f);
#pragma unroll
    for (int k = 1; k < 16; ++k) {
        const int low = k & (-k);
        const int rest = k & (k - 1);
        const int b = low == 1 ? 0 : low == 2 ? 1 : low == 4 ? 2 : 3;
        p[k] = rest == 0 ? wp[b] : cmul(p[rest], wp[b]);
    }
}

// Phase A: load segment of block `blk`, DFT32 over n1, twiddle, write smem.
// tw1[t] = W_N^t = exp(-2 pi i t / N), t < 512.
template <class Turn = NoTurn>
RRC_HD void phase_a(int tid, long long blk, const BlockIO& io, const float2* tw1, float2* sm, Turn turn = Turn()) {
    float2 v[32];
    // input index of segment element 0 (tap partition p filters the input delayed by `shift`)
    const long long seg0 = (io.real ? 2 * blk : blk) * (long long)io.V - io.T1 - io.shift;
    if (io.real) {
        if (seg0 >= 0 && seg0 + io.V + N <= io.n_in) {          // both real blocks interior
            const float* p = reinterpret_cast<const float*>(io.in) + seg0 + tid;
            float a[32], b[32];
#pragma unroll
            for (int n1 = 0; n1 < 32; ++n1) { a[n1] = p[512 * n1]; b[n1] = p[512 * n1 + io.V]; }
#pragma unroll
            for (int n1 = 0; n1 < 32; ++n1) v[bitrev(n1, 5)] = make_float2(a[n1], b[n1]);
        } else {
#pragma unroll
            for (int n1 = 0; n1 < 32; ++n1) {
                const long long g = seg0 + tid + 512 * n1;
                v[bitrev(n1, 5)] = make_float2(fetch_real(io, g), fetch_real(io, g + io.V));
            }
        }
    } else if (seg0 >= 0 && seg0 + N <= io.n_in) {              // interior block: no bounds checks
        if (io.in_u8) {
            const unsigned short* p = reinterpret_cast<const unsigned short*>(io.in) + seg0 + tid;
            unsigned int w[32];
#pragma unroll
            for (int n1 = 0; n1 < 32; ++n1) w[n1] = p[512 * n1];
#pragma unroll
            for (int n1 = 0; n1 < 32; ++n1) v[bitrev(n1, 5)] = decode_u8iq(w[n1] & 0xffu, w[n1] >> 8);
        } else {
            const float2* p = io.in + seg0 + tid;
#pragma unroll
            for (int n1 = 0; n1 < 32; ++n1) v[bitrev(n1, 5)] = p[512 * n1];
        }
    } else {
        const long long g0 = seg0 + tid;
#pragma unroll
        for (int n1 = 0; n1 < 32; ++n1) {
            const long long g = g0 + 512 * n1;
            float2 x = make_float2(0.f, 0.f);
            if (g < 0) { if (g + io.T1_total >= 0) x = io.hist[g + io.T1_total]; }
            else if (g < io.n_in) x = ld_iq(io.in, g, io.in_u8);
            v[bitrev(n1, 5)] = x;
        }
    }
    const float one = turn.acquire();
    const float2 w1 = tw1[tid];
    dit_g<32, +1>(v, one);                                      // bit-reversed in, natural k1 out
    float2 p[32];
    powers32(w1, p);
#pragma unroll
    for (int k1 = 0; k1 < 32; ++k1) v[k1] = cmul(v[k1], p[k1]);
    turn.release();
    float2* s = sm + (tid >> 4) * ROW_PITCH + (tid & 15);       // (k1 = 0, r = n2, c = n3)
#pragma unroll
    for (int k1 = 0; k1 < 32; ++k1) s[k1 * PLANE_PITCH] = v[k1];
}

// ---- staged input (fftfilt_st_kernel) ----------------------------------------------------------
// stage_input: thread t copies ITS 32 input samples of block `blk` (n = t + 512*n1) into the 32
// exchange-buffer words it owns in phase A (phys(n1, t>>4, t&15) — the very words its phase-A
// results go to), asynchronously with cp.async (SASS LDGSTS) for interior blocks.  The copy is
// issued right after phase A' of the previous block has read the buffer, so the HBM/L2 latency of
// the next block's input is hidden behind the arithmetic and stores of phase A' instead of being
// exposed at the top of phase A.  No thread ever reads another thread's staged words, so the only
// synchronisation is cp.async.wait_group by the issuing thread.
RRC_HD void stage_input(int tid, long long blk, const BlockIO& io, float2* sm) {
    float2* s = sm + (tid >> 4) * ROW_PITCH + (tid & 15);
    const long long seg0 = blk * (long long)io.V - io.T1 - io.shift;
    if (seg0 >= 0 && seg0 + N <= io.n_in && !io.in_u8) {
        const float2* p = io.in + seg0 + tid;
#if defined(__CUDA_ARCH__)
        const unsigned dst = (unsigned)__cvta_generic_to_shared(s);
#pragma unroll
        for (int n1 = 0; n1 < 32; ++n1)
            asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst + n1 * PLANE_PITCH * 8), "l"(p + 512 * n1) : "memory");
#else
        for (int n1 = 0; n1 < 32; ++n1) s[n1 * PLANE_PITCH] = p[512 * n1];
#endif
    } else {
        const long long g0 = seg0 + tid;
#pragma unroll 4
        for (int n1 = 0; n1 < 32; ++n1) {
            const long long g = g0 + 512 * n1;
            float2 x = make_float2(0.f, 0.f);
            if (g < 0) { if (g + io.T1_total >= 0) x = io.hist[g + io.T1_total]; }
            else if (g < io.n_in) x = ld_iq(io.in, g, io.in_u8);
            s[n1 * PLANE_PITCH] = x;
        }
    }
#if defined(__CUDA_ARCH__)
    asm volatile("cp.async.commit_group;" ::: "memory");
#endif
}
RRC_HD void stage_wait() {
#if defined(__CUDA_ARCH__)
    asm volatile("cp.async.wait_group 0;" ::: "memory");
#endif
}
// Phase A on staged input: the thread's own words -> DFT32 over n1 -> twiddle -> the same words.
template <class Turn = NoTurn>
RRC_HD void phase_a_staged(int tid, const float2* tw1, float2* sm, Turn turn = Turn()) {
    float2* s = sm + (tid >> 4) * ROW_PITCH + (tid & 15);
    float2 v[32];
    stage_wait();
#pragma unroll
    for (int n1 = 0; n1 < 32; ++n1) v[bitrev(n1, 5)] = s[n1 * PLANE_PITCH];
    const float one = turn.acquire();
    const float2 w1 = tw1[tid];
    dit_g<32, +1>(v, one);
    float2 p[32];
    powers32(w1, p);
#pragma unroll
    for (int k1 = 0; k1 < 32; ++k1) v[k1] = cmul(v[k1], p[k1]);
    turn.release();
#pragma unroll
    for (int k1 = 0; k1 < 32; ++k1) s[k1 * PLANE_PITCH] = v[k1];
}

// ---- linear staging (fftfilt_tma_kernel) --------------------------------------------------------
// The next block's 16384 input samples land in the exchange buffer in NATURAL order, sm[n] = x[seg0+n],
// n < N — the layout one cp.async.bulk (TMA) stream can produce: 128 KiB copied by the copy engine
// with no LSU instructions, issued by one thread right after phase A' of the previous block has read
// the buffer, so the HBM/L2 latency AND the per-SM load bandwidth (22-26 B/clk/SM, 5-6 K cycles for
// 128 KiB) are hidden behind A' instead of being the first 7 K cycles of phase A.  Phase A then reads
// sm[tid + 512*n1] (conflict free), transforms, and — after a CTA barrier, because the padded layout
// the rest of the block uses overlaps other threads' linear words — writes phys(k1, tid>>4, tid&15).
// stage_linear_bulk_ok(): CTA-uniform test that block `blk` is an interior block whose segment is a
// 16-byte aligned run of c32 samples; other blocks (history at the start, zero fill at the end, odd
// segment start) are staged by stage_linear_fallback: every thread stores the 32 elements it will
// read back itself.
RRC_HD long long stage_linear_seg0(long long blk, const BlockIO& io) { return blk * (long long)io.V - io.T1 - io.shift; }
RRC_HD bool stage_linear_bulk_ok(long long blk, const BlockIO& io) {
    const long long seg0 = stage_linear_seg0(blk, io);
    return seg0 >= 0 && seg0 + N <= io.n_in && !io.in_u8 && !io.real &&
           ((reinterpret_cast<unsigned long long>(io.in) + (unsigned long long)seg0 * 8ull) & 15ull) == 0;
}
RRC_HD void stage_linear_fallback(int tid, long long blk, const BlockIO& io, float2* sm) {
    const long long g0 = stage_linear_seg0(blk, io) + tid;
#pragma unroll 4
    for (int n1 = 0; n1 < 32; ++n1) {
        const long long g = g0 + 512 * n1;
        float2 x = make_float2(0.f, 0.f);
        if (g < 0) { if (g + io.T1_total >= 0) x = io.hist[g + io.T1_total]; }
        else if (g < io.n_in) x = ld_iq(io.in, g, io.in_u8);
        sm[tid + 512 * n1] = x;
    }
}
// Phase A on linearly staged input, in two halves around the CTA barrier.
RRC_HD void phase_a_linear_load(int tid, const float2* sm, float2 (&v)[32]) {
#pragma unroll
    for (int n1 = 0; n1 < 32; ++n1) v[bitrev(n1, 5)] = sm[tid + 512 * n1];
}
RRC_HD void phase_a_linear_compute(int tid, const float2* tw1, float2 (&v)[32]) {
    const float2 w1 = tw1[tid];
    dit<32, +1>(v);
    float2 p[32];
    powers32(w1, p);
#pragma unroll
    for (int k1 = 0; k1 < 32; ++k1) v[k1] = cmul(v[k1], p[k1]);
}
RRC_HD void phase_a_linear_store(int tid, float2* sm, const float2 (&v)[32]) {
    float2* s = sm + (tid >> 4) * ROW_PITCH + (tid & 15);
#pragma unroll
    for (int k1 = 0; k1 < 32; ++k1) s[k1 * PLANE_PITCH] = v[k1];
}

// Phase MID: B, C, B' on one k1 plane per half-warp.  tw2[k2*16 + n3] = W_512^{n3*k2}.
// Hp[(k1*32 + k2)*16 + k3] = H[k1 + 32*k2 + 1024*k3] / N.
// TW = false: the W_512^{n3*k2} twiddle is NOT applied here but in phase C (TWC = true there), from
// powers of one per-row value, so the arithmetic bursts of B and B' contain no shared-memory loads:
// a table load issued inside an arithmetic burst queues behind the other warps' exchange traffic
// and couples the FP pipe to the shared-memory pipe (no overlap between them).
template <bool TW = true, class Turn = NoTurn>
RRC_HD void phase_mid_b(int tid, const float2* tw2, float2* sm, Turn turn = Turn()) {
    const int k1 = tid >> 4, l = tid & 15;
    float2* col = sm + k1 * PLANE_PITCH + l;                    // (k1, r = 0, c = l)
    const float2* tw = tw2 + l;
    float2 v[32];
#pragma unroll
    for (int n2 = 0; n2 < 32; ++n2) v[bitrev(n2, 5)] = col[n2 * ROW_PITCH];
    const float one = turn.acquire();
    dit_g<32, +1>(v, one);
    if constexpr (TW) {
#pragma unroll
        for (int k2 = 0; k2 < 32; ++k2) v[k2] = cmul(v[k2], tw[k2 * 16]);
    }
    turn.release();
#pragma unroll
    for (int k2 = 0; k2 < 32; ++k2) col[k2 * ROW_PITCH] = v[k2];
}
// Spectrum residency: the row each thread multiplies first (k2 = l) lives in shared memory for
// the whole kernel (Hres, 512 rows, pitch HRES_PITCH so the per-thread 128-bit reads are
// conflict free); the second row (k2 = l + 16) is fetched from L2 into registers at the top of
// phase C and consumed ~400 instructions later.  (The whole spectrum, 128 KiB, does not fit
// beside the 136 KiB exchange buffer.)
constexpr int HRES_PITCH = 18;                                  // float2 per resident row (144 B)
constexpr int HRES_ELEMS = NT * HRES_PITCH;

RRC_HD void load_hres(int tid, const float2* Hp, float2* Hres) {
    const int k1 = tid >> 4, l = tid & 15;
    const float4* src = reinterpret_cast<const float4*>(Hp + (size_t)(k1 * 32 + l) * 16);
    float4* dst = reinterpret_cast<float4*>(Hres + tid * HRES_PITCH);
#pragma unroll
    for (int i = 0; i < 8; ++i) dst[i] = src[i];
}

template <bool TWC = false, class Turn = NoTurn>
RRC_HD void phase_mid_c(int tid, const float2* Hp, const float2* Hres, float2* sm, Turn turn = Turn(), const float2* tw2 = nullptr) {
    const int k1 = tid >> 4, l = tid & 15;
    const float4* hp1 = reinterpret_cast<const float4*>(Hp + (size_t)(k1 * 32 + l + 16) * 16);
    float4 h1[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) h1[i] = hp1[i];
    const float4* hres = reinterpret_cast<const float4*>(Hres + tid * HRES_PITCH);
#pragma unroll
    for (int half = 0; half < 2; ++half) {
        const int k2 = l + 16 * half;
        float2* row = sm + k1 * PLANE_PITCH + k2 * ROW_PITCH;   // (k1, r = k2, c = 0)
        float2 v[16];
        float2 g = make_float2(1.f, 0.f);
        if constexpr (TWC) g = tw2[k2 * 16 + 1];                // W_512^{k2}
#pragma unroll
        for (int n3 = 0; n3 < 16; ++n3) v[bitrev(n3, 4)] = row[n3];
        const float one = turn.acquire();
        float2 pw[16];
        if constexpr (TWC) {
            powers16(g, pw);                                    // W_512^{n3*k2}
#pragma unroll
            for (int n3 = 1; n3 < 16; ++n3) v[bitrev(n3, 4)] = cmul(v[bitrev(n3, 4)], pw[n3]);
        }
        dit_g<16, +1>(v, one);                                  // v[k3], natural order
        float2 u[16];
#pragma unroll
        for (int k3 = 0; k3 < 16; k3 += 2) {
            const float4 h = half == 0 ? hres[k3 >> 1] : h1[k3 >> 1];
            u[bitrev(k3, 4)] = cmul(v[k3], make_float2(h.x, h.y));
            u[bitrev(k3 + 1, 4)] = cmul(v[k3 + 1], make_float2(h.z, h.w));
        }
        dit<16, -1>(u);                                         // u[n3], natural order
        if constexpr (TWC) {
#pragma unroll
            for (int n3 = 1; n3 < 16; ++n3) u[n3] = cmul_conj(u[n3], pw[n3]);
        }
        turn.release();
#pragma unroll
        for (int n3 = 0; n3 < 16; ++n3) row[n3] = u[n3];
    }
}
template <bool TW = true, class Turn = NoTurn>
RRC_HD void phase_mid_bi(int tid, const float2* tw2, float2* sm, Turn turn = Turn()) {
    const int k1 = tid >> 4, l = tid & 15;
    float2* col = sm + k1 * PLANE_PITCH + l;
    const float2* tw = tw2 + l;
    float2 v[32];
#pragma unroll
    for (int k2 = 0; k2 < 32; ++k2) v[bitrev(k2, 5)] = col[k2 * ROW_PITCH];
    const float one = turn.acquire();
    if constexpr (TW) {
#pragma unroll
        for (int k2 = 0; k2 < 32; ++k2) v[bitrev(k2, 5)] = cmul_conj(v[bitrev(k2, 5)], tw[k2 * 16]);
        dit<32, -1>(v);
    } else {
        dit_g<32, -1>(v, one);
    }
    turn.release();
#pragma unroll
    for (int n2 = 0; n2 < 32; ++n2) col[n2 * ROW_PITCH] = v[n2];
}
template <bool TWC = false, class Turn = NoTurn>
RRC_HD void phase_mid(int tid, const float2* tw2, const float2* Hp, const float2* Hres, float2* sm, Turn turn = Turn()) {
    phase_mid_b<!TWC>(tid, tw2, sm, turn);
    RRC_SYNCWARP();
    phase_mid_c<TWC>(tid, Hp, Hres, sm, turn, tw2);
    RRC_SYNCWARP();
    phase_mid_bi<!TWC>(tid, tw2, sm, turn);
}

// Phase A': tid = t; conj twiddle, IDFT32 over k1 -> n1; store valid outputs.
struct NoHook { RRC_HD void operator()() const {} };

// v[n1] is segment element n = tid + 512*n1 of block `blk`; elements n >= T1 are valid outputs, filter
// output index o = o0 + n (shared by the scalar and the packed kernels).
template <bool DECIM, bool ACCUM>
RRC_HD void store_outputs(int tid, long long blk, const BlockIO& io, const float2 (&v)[32]) {
    const long long o0 = (io.real ? 2 * blk : blk) * (long long)io.V - io.T1;
    const int tq = io.T1 >> 9, tr = io.T1 & 511;
    const int first = tq + (tid < tr ? 1 : 0);                 // element n1 of this thread is valid iff n1 >= first
    if (io.real) {                                              // real stream: .re -> block 2b, .im -> block 2b+1
        float* q = reinterpret_cast<float*>(io.out) + o0 + tid;
#pragma unroll
        for (int n1 = 0; n1 < 32; ++n1) {
            if (n1 >= first) {
                const long long o = o0 + tid + 512 * n1;
                if (o < io.n_out) q[512 * n1] = ACCUM ? q[512 * n1] + v[n1].x : v[n1].x;
                if (o + io.V < io.n_out) q[512 * n1 + io.V] = ACCUM ? q[512 * n1 + io.V] + v[n1].y : v[n1].y;
            }
        }
        return;
    }
    if (io.epi.kind != RRC_EPI_NONE) {                          // fused neighbour (never with ACCUM: single partition only)
        const long long D = DECIM ? io.deci : 1;
        const long long skip = DECIM ? io.skip : 0;
        float* qf = reinterpret_cast<float*>(io.out);
#pragma unroll
        for (int n1 = 0; n1 < 32; ++n1) {
            const long long r = o0 + tid + 512 * n1 - skip;
            if (n1 >= first && r >= 0 && r % D == 0 && r / D < io.n_out) {
                if (io.epi.kind == RRC_EPI_MAG2) qf[r / D] = rrc::epi_mag2(v[n1]);
                else io.out[r / D] = rrc::epi_c32(v[n1], io.epi);
            }
        }
        return;
    }
    if constexpr (!DECIM) {
        float2* q = io.out + o0 + tid;
        if (o0 + N <= io.n_out) {                               // interior: only the n >= T1 test
            // (written with tq / tr: the single-compare form `n1 >= first` measured 3.9 % SLOWER on config 2)
#pragma unroll
            for (int n1 = 0; n1 < 32; ++n1)
                if (n1 > tq || (n1 == tq && tid >= tr)) q[512 * n1] = ACCUM ? cadd(q[512 * n1], v[n1]) : v[n1];
        } else {
            // last valid element of this thread: o0 + tid + 512*n1 < n_out  <=>  n1 < lim
            const long long rem = io.n_out - o0 - tid;
            const int lim = rem <= 0 ? 0 : (int)((rem + 511) >> 9 > 32 ? 32 : (rem + 511) >> 9);
#pragma unroll
            for (int n1 = 0; n1 < 32; ++n1)
                if (n1 >= first && n1 < lim) q[512 * n1] = ACCUM ? cadd(q[512 * n1], v[n1]) : v[n1];
        }
    } else {
        // keep outputs with (o - skip) >= 0 and (o - skip) % deci == 0, at index (o - skip)/deci.
        const long long D = io.deci;
        long long r = o0 + tid - io.skip;                       // for n1 = 0
        long long qd = r >= 0 ? r / D : -((-r + D - 1) / D);    // floor division
        long long m = r - qd * D;                               // in [0, D)
        const long long sq = 512 / D, sm_ = 512 % D;
#pragma unroll
        for (int n1 = 0; n1 < 32; ++n1) {
            if (n1 >= first && m == 0 && qd >= 0 && qd < io.n_out)
                io.out[qd] = ACCUM ? cadd(io.out[qd], v[n1]) : v[n1];
            qd += sq; m += sm_;
            if (m >= D) { m -= D; ++qd; }
        }
    }
}

// after_load() runs once the thread has read its 32 exchange-buffer words (the staged kernel puts a
// CTA barrier and the next block's stage_input there).
template <bool DECIM, bool ACCUM, class Turn = NoTurn, class AfterLoad = NoHook>
RRC_HD void phase_ai(int tid, long long blk, const BlockIO& io, const float2* tw1, const float2* sm, Turn turn = Turn(),
                     AfterLoad after_load = AfterLoad()) {
    const float2* s = sm + (tid >> 4) * ROW_PITCH + (tid & 15);
    float2 v[32];
#pragma unroll
    for (int k1 = 0; k1 < 32; ++k1) v[bitrev(k1, 5)] = s[k1 * PLANE_PITCH];
    after_load();
    turn.acquire();
    const float2 w1 = tw1[tid];
    float2 p[32];
    powers32(w1, p);
#pragma unroll
    for (int k1 = 0; k1 < 32; ++k1) v[bitrev(k1, 5)] = cmul_conj(v[bitrev(k1, 5)], p[k1]);
    dit<32, -1>(v);
    turn.release();
    store_outputs<DECIM, ACCUM>(tid, blk, io, v);
}

}}  // namespace rrc::fftk
