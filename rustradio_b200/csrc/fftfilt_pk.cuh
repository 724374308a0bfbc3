// fftfilt_pk.cuh — one 16384-point overlap-save block of the FftFilter kernel with PACKED FP32
// arithmetic (sm_100a FFMA2 / FADD2 / FMUL2: two FP32 lanes per issue slot).
//
// Same transform, thread roles and 3 x (32, 32, 16) factorisation as fftfilt_core.cuh:
//   n = n1*512 + n2*16 + n3,  t = n2*16 + n3,  k = k1 + 32*k2 + 1024*k3
//   A  (thread t):        a[k1;t]      = DFT32_{n1}(x[n1*512+t]) * W_N^{t k1}
//   B  (thread k1,n3):    b[k1,k2;n3]  = DFT32_{n2}(a[k1;n2*16+n3]) * W_512^{n3 k2}
//   C  (thread k1,l):     rows k2 = l, l+16:  IDFT16_{k3}( DFT16_{n3}(b) * H )
//   B' (thread k1,n3):    z[k1;n2,n3]  = IDFT32_{k2}( . * conj W_512^{n3 k2} )
//   A' (thread t):        y[n1*512+t]  = IDFT32_{k1}( z * conj W_N^{t k1} )
// but inside every thread the 32 (16) points of a transform are TWO LANES of 16 (two rows in C)
// that go through identical butterflies with identical compile-time twiddles, so they ride in the two
// halves of 64-bit register pairs (fft_regs2.cuh): 4 of the 5 radix-2 stages of a 32-point transform,
// all of a 16-point one, every twiddle / spectrum multiplication and the twiddle powers are packed.
// The one stage that crosses the lanes is done with scalar FFMA on the halves (no moves): it is the
// FIRST stage of the decimation-in-frequency transforms (A, B': inputs arrive as (re,im) / (j,j+16)
// words, outputs leave as pairs of ADJACENT indices) and the LAST stage of the decimation-in-time
// ones (B, A': inputs arrive as pairs of adjacent indices, outputs leave as (k,k+16) pairs / (re,im)).
// Issue slots per thread and block: ~2500 against ~3900 for the scalar kernel; FP32-pipe work is the
// same (+4 % for the DIF forms), which is the point: the FP pipe can stay busy while the freed issue
// slots carry the shared-memory traffic.
//
// Shared-memory exchange layouts (64-bit words; plane pitch PP = 544 words, plane k1 at k1*PP):
//   L0  landing / A input : x[n1*512 + t]                     word n1*PP + t          (re, im)
//   L1  A -> B            : (a[k1;t], a[k1;t+16]) for n2 even  re: k1*PP + 32q + n3    q = n2/2
//                                                              im: k1*PP + 32q + 16 + n3
//         thread t = 32q + 16h + n3 and its partner t^16 (same warp) own exactly these two words of
//         every plane in L0 as well, so phase A works IN PLACE pairwise: __syncwarp, no CTA barrier.
//   L2  B -> C -> B'      : (b[k1,j;n3], b[k1,j+16;n3])        k1*PP + c*272 + j*17 + n3   c = re/im
//         16 x 16 words per component, row pitch 17: column (B, B') and row (C) accesses conflict free.
//   L3  B' -> A'          : (z[2m;t], z[2m+1;t])               m*1088 + c*512 + t
//         the pair of planes (2m, 2m+1) belongs to ONE warp, which fills it with 32-bit stores (the two
//         half-warps write the two halves of the same 16 words: 32 distinct banks).
// CTA barriers per block: A | MID | A' loads  -> 3 (the scalar TMA kernel needs 4).
//
// Replaces Engine::run + sum_vec (rustradio src/fft_filter.rs:172-176,281-287): IFFT(FFT(x) * H), 1/N in H.
#pragma once
#include "fft_regs2.cuh"
#include "fftfilt_core.cuh"

namespace rrc { namespace fftp {

using namespace rrc::fftr;
using rrc::fftk::BlockIO;
using rrc::fftk::N;
using rrc::fftk::NT;

constexpr int PP = 544;                        // plane pitch in 64-bit words
constexpr int SMEM_WORDS = 32 * PP;            // 17408 words = 136 KiB
constexpr int L2_COMP = 272;                   // 16 rows x pitch 17
constexpr int L3_PAIR = 2 * PP;                // 1088: region of the plane pair (2m, 2m+1)
constexpr int HRES_PITCH = 18;                 // words per thread of the resident re-pairs (16 + 2 pad: conflict-free LDS.128)
constexpr int HRES_WORDS = NT * HRES_PITCH;

// ---- 64-bit word <-> packed pair ---------------------------------------------------------------
RRC_HD F2 ld_pair(const float2* sm, int w) {
#if defined(__CUDA_ARCH__)
    F2 r; r.v = reinterpret_cast<const unsigned long long*>(sm)[w]; return r;
#else
    return f2(sm[w].x, sm[w].y);
#endif
}
RRC_HD void st_pair(float2* sm, int w, F2 v) {
#if defined(__CUDA_ARCH__)
    reinterpret_cast<unsigned long long*>(sm)[w] = v.v;
#else
    sm[w] = make_float2(f2_lo(v), f2_hi(v));
#endif
}
RRC_HD void st_half(float2* sm, int w, int h, float x) { reinterpret_cast<float*>(sm)[2 * w + h] = x; }

#if defined(__CUDA_ARCH__)
#define RRC_PK_SYNCWARP() __syncwarp()
#else
#define RRC_PK_SYNCWARP() ((void)0)
#endif

// ---- FP-turn policy ------------------------------------------------------------------------------
// Every phase is  loads -> turn.acquire() -> arithmetic -> turn.release() -> stores.  NoTurn: no-ops.
// PingPong (device): the 16 warps form two groups of 8 (two warps of each group on every SM sub-partition)
// that pass an "FP turn" token through named barriers 1 and 2: one group's arithmetic burst runs while the
// other group's shared-memory burst is in flight, instead of all warps convoying through the same kind of
// work.  (With scalar FFMA the arithmetic group needs every issue slot and starves the other group's LDS/STS
// issue — measured slower in round 1; the packed arithmetic needs only half the slots.)
// pin(): an empty volatile asm with the value as a read-write operand.  Volatile asms keep their order, so
// arithmetic on a pinned value cannot be hoisted above the acquire (ptxas otherwise moves register-only FP
// freely across BAR.SYNC), and results pinned before release() are computed before the token is handed on.
// Stagger (device): after every CTA barrier all 16 warps issue their 32-64 shared-memory loads at once; the
// pipe serves them interleaved, so every warp gets its data at the END of the burst (~2 K cycles), all start
// their arithmetic together, all store together: a convoy in which the FP pipes idle during the memory
// bursts and the memory pipe idles during the arithmetic (measured: smem 47 % + FP 47 % = the whole time).
// loads_begin() holds warp w back by w * step cycles before its FIRST load burst after a CTA barrier (step = the
// time one warp's burst occupies the shared-memory pipe), so the bursts are served one warp after the other:
// warp 0 has its data after ~130 cycles and computes while the others are still being served — nobody gets
// its data later than in the interleaved case — and the skew persists through the barrier-free MID phases.
// (A chain of named barriers did the same but ptxas sinks BAR.ARV below the arithmetic: 1 K cycles per link.)
struct NoTurn {
    RRC_HD void acquire() const {}
    RRC_HD void release() const {}
    RRC_HD void loads_begin() const {}
    RRC_HD void loads_end() const {}
};
#if defined(__CUDACC__)
struct PingPong {
    int g;               // group 0 or 1
    __device__ __forceinline__ void acquire() const { asm volatile("bar.sync %0, 512;" ::"r"(1 + g) : "memory"); }
    __device__ __forceinline__ void release() const { asm volatile("bar.arrive %0, 512;" ::"r"(2 - g) : "memory"); }
    __device__ __forceinline__ void loads_begin() const {}
    __device__ __forceinline__ void loads_end() const {}
};
struct Stagger {
    int delay;           // cycles this warp holds back its first load burst after a CTA barrier: warp index * step
    __device__ __forceinline__ void acquire() const {}
    __device__ __forceinline__ void release() const {}
    __device__ __forceinline__ void loads_begin() const {
        if (delay > 0) {
            long long t0, t;
            asm volatile("mov.u64 %0, %%clock64;" : "=l"(t0) :: "memory");
            do { asm volatile("mov.u64 %0, %%clock64;" : "=l"(t) :: "memory"); } while (t - t0 < delay);
        }
    }
    __device__ __forceinline__ void loads_end() const {}
};
#endif
#if defined(__CUDA_ARCH__)
__device__ __forceinline__ void pin(F2& x) { asm volatile("" : "+l"(x.v)); }
__device__ __forceinline__ void pin(float2& x) { asm volatile("" : "+f"(x.x), "+f"(x.y)); }
#else
inline void pin(F2&) {}
inline void pin(float2&) {}
#endif
RRC_HD void pin(C2& x) { pin(x.re); pin(x.im); }
template <int CNT, class T>
RRC_HD void pin_all(T* v) {
#pragma unroll
    for (int i = 0; i < CNT; ++i) pin(v[i]);
}

// ---- packed decimation-in-frequency (natural in, bit-reversed out) -----------------------------
// difb2<NUM, DEN, DIR>(a, b): a <- a + b, b <- (a - b) * w, w = exp(-DIR*2*pi*i*NUM/DEN); all packed,
// the twiddle is an immediate (its sign is folded into the constants: no negated operand needed).
template <int NUM, int DEN, int DIR>
RRC_HD void difb2(C2& a, C2& b) {
    constexpr int n64 = NUM * (64 / DEN);
    const C2 s = c2_add(a, b);
    if constexpr (n64 == 0) {
        b = c2_sub(a, b);
    } else if constexpr (n64 == 16) {              // * (-i) fwd: (d.im, -d.re);  * (+i) inv: (-d.im, d.re)
        C2 r;
        if constexpr (DIR > 0) { r.re = a.im - b.im; r.im = b.re - a.re; }
        else                   { r.re = b.im - a.im; r.im = a.re - b.re; }
        b = r;
    } else {
        constexpr float wr = (float)cos64(n64);
        constexpr float wi = (float)(-DIR * sin64(n64));
        const C2 d = c2_sub(a, b);
        C2 r;
        r.re = fmac(d.im, -wi, d.re * splat(wr));
        r.im = fmac(d.im, wr, d.re * splat(wi));
        b = r;
    }
    a = s;
}
template <int NN, int DIR, int I>
struct DifLevel2 {
    static RRC_HD void run(C2* v) {
        difb2<I, NN, DIR>(v[I], v[I + NN / 2]);
        if constexpr (I + 1 < NN / 2) DifLevel2<NN, DIR, I + 1>::run(v);
    }
};
template <int NN, int DIR>
RRC_HD void dif2(C2* v) {
    if constexpr (NN >= 2) {
        DifLevel2<NN, DIR, 0>::run(v);
        dif2<NN / 2, DIR>(v);
        dif2<NN / 2, DIR>(v + NN / 2);
    }
}
// packed decimation-in-time without the dit_g "one" operand
template <int NN, int DIR>
RRC_HD void dit2p(C2* v) {
    if constexpr (NN >= 2) {
        dit2p<NN / 2, DIR>(v);
        dit2p<NN / 2, DIR>(v + NN / 2);
        DitLevel2<NN, DIR, 0>::run(v);
    }
}

// Scalar lane-crossing stages on the halves of 16 packed values.
// DIF first stage: lane0[i] = x[i] + x[i+16], lane1[i] = (x[i] - x[i+16]) * W_32^{+-i}.
template <int DIR, int I>
struct DifFirst {
    static RRC_HD void run(const float2* x, C2* p) {
        float2 a = x[I], b = x[I + 16];
        const float2 s = cadd(a, b);
        const float2 d = mul_w<I, 32, DIR>(csub(a, b));
        p[I].re = f2(s.x, d.x);
        p[I].im = f2(s.y, d.y);
        if constexpr (I + 1 < 16) DifFirst<DIR, I + 1>::run(x, p);
    }
};
// The same first stage when the 32 inputs already sit in packed (j, j+16) pairs (phase B').
template <int DIR, int I>
struct DifFirstPk {
    static RRC_HD void run(C2* p) {
        const float2 a = make_float2(f2_lo(p[I].re), f2_lo(p[I].im)), b = make_float2(f2_hi(p[I].re), f2_hi(p[I].im));
        const float2 s = cadd(a, b);
        const float2 d = mul_w<I, 32, DIR>(csub(a, b));
        p[I].re = f2(s.x, d.x);
        p[I].im = f2(s.y, d.y);
        if constexpr (I + 1 < 16) DifFirstPk<DIR, I + 1>::run(p);
    }
};
// DIT last stage: out[i] = E[i] + W_32^{+-i} O[i], out[i+16] = E[i] - W O[i]  (E = lane0, O = lane1).
template <int DIR, int I>
struct DitLast {
    static RRC_HD void run(const C2* p, float2* out) {
        float2 a = make_float2(f2_lo(p[I].re), f2_lo(p[I].im)), b = make_float2(f2_hi(p[I].re), f2_hi(p[I].im));
        bfly<I, 32, DIR>(a, b);
        out[I] = a; out[I + 16] = b;
        if constexpr (I + 1 < 16) DitLast<DIR, I + 1>::run(p, out);
    }
};
template <int DIR, int I>
struct DitLastPk {                              // results back into (i, i+16) packed pairs (phase B)
    static RRC_HD void run(C2* p) {
        float2 a = make_float2(f2_lo(p[I].re), f2_lo(p[I].im)), b = make_float2(f2_hi(p[I].re), f2_hi(p[I].im));
        bfly<I, 32, DIR>(a, b);
        p[I].re = f2(a.x, b.x);
        p[I].im = f2(a.y, b.y);
        if constexpr (I + 1 < 16) DitLastPk<DIR, I + 1>::run(p);
    }
};

// Q[m] = (w^{2m}, w^{2m+1}), m < 16, w = W_N^t: a binary tree of packed multiplications by the
// lane-broadcast powers u, u^2, u^4, u^8 of u = w^2 (depth <= 4 multiplications + 4 squarings).
RRC_HD void pair_powers(float2 w, C2 (&q)[16]) {
    float2 u[4];
    u[0] = csqr(w);
#pragma unroll
    for (int i = 1; i < 4; ++i) u[i] = csqr(u[i - 1]);
    q[0].re = f2(1.f, w.x);
    q[0].im = f2(0.f, w.y);
#pragma unroll
    for (int m = 1; m < 16; ++m) {
        const int low = m & (-m);
        const int rest = m & (m - 1);
        const int b = low == 1 ? 0 : low == 2 ? 1 : low == 4 ? 2 : 3;
        q[m] = c2_mul(q[rest], C2{splat(u[b].x), splat(u[b].y)});
    }
}

// ---- staging into L0 ---------------------------------------------------------------------------
RRC_HD long long seg0_of(long long blk, const BlockIO& io) { return blk * (long long)io.V - io.T1 - io.shift; }
// CTA-uniform: block `blk` is an interior block whose 32 rows of 512 samples are 16-byte aligned c32 runs.
RRC_HD bool bulk_ok(long long blk, const BlockIO& io) {
    const long long s0 = seg0_of(blk, io);
    return s0 >= 0 && s0 + N <= io.n_in && !io.in_u8 && !io.real &&
           ((reinterpret_cast<unsigned long long>(io.in) + (unsigned long long)s0 * 8ull) & 15ull) == 0;
}
// Edge blocks (carried history, zero fill, unaligned segment, u8 input): every thread stores the 32 words it
// reads back itself.
RRC_HD void stage_fallback(int tid, long long blk, const BlockIO& io, float2* sm) {
    const long long g0 = seg0_of(blk, io) + tid;
#pragma unroll 4
    for (int n1 = 0; n1 < 32; ++n1) {
        const long long g = g0 + 512 * n1;
        float2 x = make_float2(0.f, 0.f);
        if (g < 0) { if (g + io.T1_total >= 0) x = io.hist[g + io.T1_total]; }
        else if (g < io.n_in) x = ld_iq(io.in, g, io.in_u8);
        sm[n1 * PP + tid] = x;
    }
}

// ---- phase A -----------------------------------------------------------------------------------
// tw1[t] = W_N^t (t < 512).
// Every exchanging phase is written as compute (loads + arithmetic, results in registers) and store, with a
// __syncwarp between them on the device; the CPU emulator runs all computes of a phase before its stores.
template <class Turn = NoTurn>
RRC_HD void phase_a_compute(int tid, const float2* tw1, const float2* sm, C2 (&p)[16], Turn turn = Turn()) {
    float2 x[32];
    turn.loads_begin();
#pragma unroll
    for (int n1 = 0; n1 < 32; ++n1) x[n1] = sm[n1 * PP + tid];
    float2 w = tw1[tid];
    turn.loads_end();
    turn.acquire();
    pin_all<32>(x); pin(w);
    DifFirst<+1, 0>::run(x, p);                 // lane0 = even k1 branch, lane1 = odd k1 branch
    dif2<16, +1>(p);                            // p[r] = (X[2 br4(r)], X[2 br4(r) + 1])
    C2 q[16];
    pair_powers(w, q);
#pragma unroll
    for (int r = 0; r < 16; ++r) p[r] = c2_mul(p[r], q[bitrev(r, 4)]);
    pin_all<16>(p);
    turn.release();
}
RRC_HD void phase_a_store(int tid, float2* sm, const C2 (&p)[16]) {
    const int qq = tid >> 5, h = (tid >> 4) & 1, n3 = tid & 15;
    const int wre = 32 * qq + n3, wim = wre + 16;
#pragma unroll
    for (int r = 0; r < 16; ++r) {
        const int m = bitrev(r, 4);
        st_half(sm, (2 * m) * PP + wre, h, f2_lo(p[r].re));
        st_half(sm, (2 * m) * PP + wim, h, f2_lo(p[r].im));
        st_half(sm, (2 * m + 1) * PP + wre, h, f2_hi(p[r].re));
        st_half(sm, (2 * m + 1) * PP + wim, h, f2_hi(p[r].im));
    }
}
template <class Turn = NoTurn>
RRC_HD void phase_a(int tid, const float2* tw1, float2* sm, Turn turn = Turn()) {
    C2 p[16];
    phase_a_compute(tid, tw1, sm, p, turn);
    RRC_PK_SYNCWARP();                          // the partner lane (tid ^ 16) has read its L0 words
    phase_a_store(tid, sm, p);
}

// ---- phase B -----------------------------------------------------------------------------------
// tw2p[(j*2 + c)*16 + n3] = component c of the pair (W_512^{n3 j}, W_512^{n3 (j+16)}).
template <class Turn = NoTurn>
RRC_HD void phase_b_compute(int tid, const float2* tw2p, const float2* sm, C2 (&v)[16], Turn turn = Turn()) {
    const int k1 = tid >> 4, l = tid & 15;
    const float2* pl = sm + k1 * PP;
    turn.loads_begin();
#pragma unroll
    for (int q = 0; q < 16; ++q) {              // pair q = (n2 = 2q, 2q+1) -> DIT register br4(q)
        v[bitrev(q, 4)].re = ld_pair(pl, 32 * q + l);
        v[bitrev(q, 4)].im = ld_pair(pl, 32 * q + 16 + l);
    }
    C2 w[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) w[j] = C2{ld_pair(tw2p, (2 * j) * 16 + l), ld_pair(tw2p, (2 * j + 1) * 16 + l)};
    turn.loads_end();
    turn.acquire();
    pin_all<16>(v);
    dit2p<16, +1>(v);                           // lane0 = E[kk], lane1 = O[kk]
    DitLastPk<+1, 0>::run(v);                   // v[j] = (b[k2 = j], b[k2 = j + 16])
#pragma unroll
    for (int j = 0; j < 16; ++j) v[j] = c2_mul(v[j], w[j]);
    pin_all<16>(v);
    turn.release();
}
RRC_HD void phase_b_store(int tid, float2* sm, const C2 (&v)[16]) {
    const int k1 = tid >> 4, l = tid & 15;
    float2* pl = sm + k1 * PP;
#pragma unroll
    for (int j = 0; j < 16; ++j) {
        st_pair(pl, j * 17 + l, v[j].re);
        st_pair(pl, L2_COMP + j * 17 + l, v[j].im);
    }
}
template <class Turn = NoTurn>
RRC_HD void phase_b(int tid, const float2* tw2p, float2* sm, Turn turn = Turn()) {
    C2 v[16];
    phase_b_compute(tid, tw2p, sm, v, turn);
    RRC_PK_SYNCWARP();                          // every lane of the plane has read its L1 words
    phase_b_store(tid, sm, v);
}

// ---- phase C -----------------------------------------------------------------------------------
// Hq[tid*32 + k3]      = (Re H[k1 + 32 l + 1024 k3], Re H[k1 + 32 (l+16) + 1024 k3]) / N
// Hq[tid*32 + 16 + k3] = the imaginary parts.  The re-pairs are resident in shared memory (Hres), the
// im-pairs come from L2 at the top of the phase.
RRC_HD void load_hres(int tid, const float2* Hq, float2* Hres) {
    const float4* src = reinterpret_cast<const float4*>(Hq + (size_t)tid * 32);
    float4* dst = reinterpret_cast<float4*>(Hres + tid * HRES_PITCH);
#pragma unroll
    for (int i = 0; i < 8; ++i) dst[i] = src[i];
}
template <class Turn = NoTurn>
RRC_HD void phase_c(int tid, const float2* Hq, const float2* Hres, float2* sm, Turn turn = Turn()) {
    const int k1 = tid >> 4, l = tid & 15;
    const float4* hg = reinterpret_cast<const float4*>(Hq + (size_t)tid * 32 + 16);
    float4 him[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) him[i] = hg[i];
    const float4* hr = reinterpret_cast<const float4*>(Hres + tid * HRES_PITCH);
    float2* row = sm + k1 * PP + l * 17;
    C2 v[16];
#pragma unroll
    for (int n3 = 0; n3 < 16; ++n3) {
        v[bitrev(n3, 4)].re = ld_pair(row, n3);
        v[bitrev(n3, 4)].im = ld_pair(row, L2_COMP + n3);
    }
    float4 hre[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) hre[i] = hr[i];
    turn.acquire();
    pin_all<16>(v);
    dit2p<16, +1>(v);                           // v[k3], both rows
    C2 u[16];
#pragma unroll
    for (int k3 = 0; k3 < 16; k3 += 2) {
        const float4 a = hre[k3 >> 1], b = him[k3 >> 1];
        u[bitrev(k3, 4)] = c2_mul(v[k3], C2{f2(a.x, a.y), f2(b.x, b.y)});
        u[bitrev(k3 + 1, 4)] = c2_mul(v[k3 + 1], C2{f2(a.z, a.w), f2(b.z, b.w)});
    }
    dit2p<16, -1>(u);                           // u[n3]
    pin_all<16>(u);
    turn.release();
#pragma unroll
    for (int n3 = 0; n3 < 16; ++n3) {
        st_pair(row, n3, u[n3].re);
        st_pair(row, L2_COMP + n3, u[n3].im);
    }
}

// ---- phase B' ----------------------------------------------------------------------------------
template <class Turn = NoTurn>
RRC_HD void phase_bi_compute(int tid, const float2* tw2p, const float2* sm, C2 (&v)[16], Turn turn = Turn()) {
    const int k1 = tid >> 4, l = tid & 15;
    const float2* pl = sm + k1 * PP;
    C2 w[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) {
        v[j] = C2{ld_pair(pl, j * 17 + l), ld_pair(pl, L2_COMP + j * 17 + l)};
        w[j] = C2{ld_pair(tw2p, (2 * j) * 16 + l), ld_pair(tw2p, (2 * j + 1) * 16 + l)};
    }
    turn.acquire();
    pin_all<16>(v);
#pragma unroll
    for (int j = 0; j < 16; ++j) v[j] = c2_mul_conj(v[j], w[j]);
    DifFirstPk<-1, 0>::run(v);
    dif2<16, -1>(v);                            // v[r] = (z[n2 = 2 br4(r)], z[2 br4(r) + 1])
    pin_all<16>(v);
    turn.release();
}
RRC_HD void phase_bi_store(int tid, float2* sm, const C2 (&v)[16]) {
    const int k1 = tid >> 4, l = tid & 15;
    float2* dst = sm + (k1 >> 1) * L3_PAIR + l;
    const int h = k1 & 1;
#pragma unroll
    for (int r = 0; r < 16; ++r) {
        const int n2 = 2 * bitrev(r, 4);
        st_half(dst, n2 * 16, h, f2_lo(v[r].re));
        st_half(dst, 512 + n2 * 16, h, f2_lo(v[r].im));
        st_half(dst, (n2 + 1) * 16, h, f2_hi(v[r].re));
        st_half(dst, 512 + (n2 + 1) * 16, h, f2_hi(v[r].im));
    }
}
template <class Turn = NoTurn>
RRC_HD void phase_bi(int tid, const float2* tw2p, float2* sm, Turn turn = Turn()) {
    C2 v[16];
    phase_bi_compute(tid, tw2p, sm, v, turn);
    RRC_PK_SYNCWARP();                          // both planes of this warp have read their L2 words
    phase_bi_store(tid, sm, v);
}

// MID = B, C, B' on the two planes of a warp; half-warp local except the last exchange (warp local).
template <class Turn = NoTurn>
RRC_HD void phase_mid(int tid, const float2* tw2p, const float2* Hq, const float2* Hres, float2* sm, Turn turn = Turn()) {
    phase_b(tid, tw2p, sm, turn);
    RRC_PK_SYNCWARP();
    phase_c(tid, Hq, Hres, sm, turn);
    RRC_PK_SYNCWARP();
    phase_bi(tid, tw2p, sm, turn);
}

// ---- phase A' ----------------------------------------------------------------------------------
// after_load() runs once the thread has read its 32 words (the kernel puts the CTA barrier and the next
// block's staging there).  The store of the valid outputs is the scalar kernel's (fftk::store_outputs).
template <bool DECIM, bool ACCUM, class AfterLoad = rrc::fftk::NoHook, class Turn = NoTurn>
RRC_HD void phase_ai(int tid, long long blk, const BlockIO& io, const float2* tw1, const float2* sm,
                     AfterLoad after_load = AfterLoad(), Turn turn = Turn()) {
    C2 z[16];
    turn.loads_begin();
#pragma unroll
    for (int m = 0; m < 16; ++m) {              // pair m = (k1 = 2m, 2m+1) -> DIT register br4(m)
        z[bitrev(m, 4)].re = ld_pair(sm, m * L3_PAIR + tid);
        z[bitrev(m, 4)].im = ld_pair(sm, m * L3_PAIR + 512 + tid);
    }
    float2 w = tw1[tid];
    turn.loads_end();
    after_load();
    turn.acquire();
    pin_all<16>(z); pin(w);
    C2 q[16];
    pair_powers(w, q);
#pragma unroll
    for (int m = 0; m < 16; ++m) z[bitrev(m, 4)] = c2_mul_conj(z[bitrev(m, 4)], q[m]);
    dit2p<16, -1>(z);
    float2 y[32];
    DitLast<-1, 0>::run(z, y);                  // y[n1], natural order
    pin_all<32>(y);
    turn.release();
    rrc::fftk::store_outputs<DECIM, ACCUM>(tid, blk, io, y);
}

}}  // namespace rrc::fftp
