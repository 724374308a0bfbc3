// fir_tc5.cu — FirFilter (reference src/fir.rs:166-197, Fir::filter_n; block src/fir.rs:492-527) for c32 samples, real
// taps, decimation 1 and ntaps <= 65 on the 5th-generation tensor cores: tcgen05.mma (kind::f16) with the accumulators
// in tensor memory.  Same arithmetic contract as fir_tc.cuh (block-scaled fp16 hi + lo split of samples and taps, the
// three products hi*hi + hi*lo + lo*hi accumulated in FP32), different machinery:
//
//   * A CTA tile is 8192 outputs = 128 block-rows of 64.  The staged fp16 plane of one component (re or im; hi or lo
//     part) is 129 rows of 64 samples, i.e. 129 rows of 128 bytes: exactly the rows of SWIZZLE_128B K-major atoms
//     (8 rows x 128 B, 16-byte chunk c of row r stored at chunk c ^ (r % 8)).  The Toeplitz operand
//     A[b][k] = z[64 b + k], k < 128, is never built: for k < 64 it IS the staged plane (rows 0..127), for k >= 64 it
//     is the same plane advanced by one row (descriptor start address + 128 B, rows 1..128).
//   * B[n][k] = w'[k - n] (reversed, scaled taps; 64 x 128, K-major, two 64-column halves, hi and lo parts) is built
//     once on the host in the swizzled shared-memory image and copied in at kernel start.
//   * D = A * B^T: M = 128, N = 64, K = 16 per instruction, up to 8 k-steps x 3 products x 2 components = 48
//     tcgen05.mma per tile, issued by one thread; completion arrives on an mbarrier (tcgen05.commit).
//   * Epilogue: tcgen05.ld.16x256b (the mma C-fragment distribution: a lane gets two adjacent columns of one row), so
//     the re and im accumulators of an output meet in one lane and four lanes write 64 contiguous bytes.
//
// Two CTAs per SM (2 x ~101 KB of shared memory, 2 x 128 TMEM columns) overlap each other's phases; inside a CTA the
// next tile's global loads are in flight while the current tile's MMAs run and its outputs are stored.
#include <cuda_fp16.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "common.cuh"
#include "fir_tc.hpp"

namespace rrc {
namespace {

constexpr int TC5_THREADS = 256;
constexpr int TC5_ROWS = 129;                       // staged rows of 64 samples
constexpr int TC5_PLANE = 17 * 1024;                // bytes per plane (129 * 128 rounded up to the 1024-byte atom)
constexpr int TC5_BIMG = 32 * 1024;                 // B image: {hi, lo} x {k < 64, k >= 64} x 64 rows x 128 B
constexpr int TC5_NLD = 17;                         // float4 loads per lane: rows warp, warp + 8, ...
constexpr unsigned TC5_IDESC = 0x08100010u;         // kind::f16: D f32, A/B f16 K-major, N = 64 (>>3 at bit 17), M = 128 (>>4 at bit 24)
constexpr size_t TC5_SMEM = 1024 + 4 * TC5_PLANE + TC5_BIMG + 64;

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

// Shared-memory matrix descriptor, SWIZZLE_128B K-major (cute/arch/mma_sm100_desc.hpp): start address >> 4 at [0,14),
// LBO (unused for a swizzled K-major operand whose K extent stays inside the atom) = 1 at [16,30), SBO = 1024 B >> 4
// at [32,46), version 1 at [46,48), base offset at [49,52), layout type 2 at [61,64).  Measured on B200: with base
// offset 0 the XOR pattern follows the ABSOLUTE shared-memory address bits [7,10), so an operand that starts one
// 128-byte row into an atom needs nothing but the advanced start address (base offset 1 shifts the pattern by one
// row and reads the wrong chunks: tools/gpu/tc5_check.py prints the map).
constexpr unsigned TC5_DESC_HI = 0x40004040u;          // SBO 1024 B (>>4) | version 1 | SWIZZLE_128B: bits [32,64) of every descriptor here
__device__ __forceinline__ unsigned tc5_desc_lo(unsigned addr) { return ((addr >> 4) & 0x3fffu) | (1u << 16); }

__device__ __forceinline__ void tc5_mma(unsigned d_tmem, unsigned a_lo, unsigned a_hi, unsigned b_lo, unsigned accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\tsetp.ne.b32 p, %5, 0;\n\t"
        "mov.b64 da, {%1, %2};\n\tmov.b64 db, {%3, %6};\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %4, p;\n\t}"
        :: "r"(d_tmem), "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(TC5_IDESC), "r"(accumulate), "r"(TC5_DESC_HI) : "memory");
}

__device__ __forceinline__ unsigned tc5_elect() {
    unsigned pred;
    asm volatile("{\n\t.reg .pred px;\n\telect.sync _|px, 0xffffffff;\n\tselp.b32 %0, 1, 0, px;\n\t}" : "=r"(pred));
    return pred;
}

__device__ __forceinline__ void tc5_ld16(unsigned taddr, unsigned (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.16x256b.x4.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr) : "memory");
}

__device__ __forceinline__ void tc5_split2(float a, float b, unsigned& hi, unsigned& lo) {
    const __half2 h = __floats2half2_rn(a, b);
    const float2 hf = __half22float2(h);
    const __half2 l = __floats2half2_rn(a - hf.x, b - hf.y);
    hi = *reinterpret_cast<const unsigned*>(&h);
    lo = *reinterpret_cast<const unsigned*>(&l);
}

// Tile loads: row (warp + 8u) of the tile, samples 2*lane, 2*lane + 1 of that row.
__device__ __forceinline__ void tc5_load(const FirTc5Args& a, long long tile, int warp, int lane, float4 (&v)[TC5_NLD]) {
    const long long ch = tile / a.tiles_x, tx = tile - ch * a.tiles_x;
    const long long s0 = tx * FIR_TC5_BT;
    const float2* in = a.in + ch * a.in_stride + s0;
    const long long avail = a.need - s0;
    const bool fast = avail >= 64ll * TC5_ROWS && (reinterpret_cast<uintptr_t>(in) & 15) == 0;
    if (fast) {
        const float4* p = reinterpret_cast<const float4*>(in) + lane;
#pragma unroll
        for (int u = 0; u < TC5_NLD; ++u) {
            const int row = warp + 8 * u;
            v[u] = row < TC5_ROWS ? __ldg(p + row * 32) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
    } else {
#pragma unroll
        for (int u = 0; u < TC5_NLD; ++u) {
            const long long s = 64ll * (warp + 8 * u) + 2 * lane;
            v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (warp + 8 * u < TC5_ROWS) {
                if (s < avail) { const float2 p = __ldg(in + s); v[u].x = p.x; v[u].y = p.y; }
                if (s + 1 < avail) { const float2 q = __ldg(in + s + 1); v[u].z = q.x; v[u].w = q.y; }
            }
        }
    }
}

constexpr int TC5_NSTAMP = 9, TC5_TRACE_IT = 4;
#define TC5_STAMP(k) do { if (trace && blockIdx.x == 2 && it >= 1 && it <= TC5_TRACE_IT && lane == 0) trace[((it - 1) * 8 + warp) * TC5_NSTAMP + (k)] = clock64(); } while (0)

__global__ void __launch_bounds__(TC5_THREADS, 2) fir_tc5_kernel(const FirTc5Args a, long long* __restrict__ trace) {
    extern __shared__ unsigned char smem_raw[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const unsigned raw = smem_u32(smem_raw);
    unsigned char* sm = smem_raw + ((1024u - (raw & 1023u)) & 1023u);       // swizzle atoms need 1024-byte alignment
    unsigned char* s_planes = sm;                                          // plane p = 2 * (im?) + (lo?)
    unsigned char* s_b = sm + 4 * TC5_PLANE;
    unsigned long long* s_bar = reinterpret_cast<unsigned long long*>(s_b + TC5_BIMG);
    unsigned* s_tmem = reinterpret_cast<unsigned*>(s_bar + 1);
    unsigned* s_red = s_tmem + 1;                                          // 8 words
    const unsigned planes_u = smem_u32(s_planes), b_u = smem_u32(s_b), bar_u = smem_u32(s_bar);

    {   // B image -> shared memory, zero the plane padding rows once (never written again, never read by a valid output)
        const uint4* src = a.bimg;
        uint4* dst = reinterpret_cast<uint4*>(s_b);
        for (int i = tid; i < TC5_BIMG / 16; i += TC5_THREADS) dst[i] = __ldg(src + i);
        uint4* pl = reinterpret_cast<uint4*>(s_planes);
        for (int i = tid; i < 4 * TC5_PLANE / 16; i += TC5_THREADS) pl[i] = make_uint4(0u, 0u, 0u, 0u);
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(bar_u) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 128;" :: "r"(smem_u32(s_tmem)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const unsigned tmem = *s_tmem;

    float4 v[TC5_NLD];
    long long tile = blockIdx.x;
    if (tile < a.total_tiles) tc5_load(a, tile, warp, lane, v);
    unsigned phase = 0;
    int it = 0;
    for (; tile < a.total_tiles; tile += gridDim.x, ++it) {
        TC5_STAMP(0);
        // ---- A. largest finite magnitude of the tile -> power-of-two scale (CTA-wide)
        float mx = 0.f;
#pragma unroll
        for (int u = 0; u < TC5_NLD; ++u)
            mx = fmaxf(fmaxf(mx, fmaxf(fabsf(v[u].x), fabsf(v[u].y))), fmaxf(fabsf(v[u].z), fabsf(v[u].w)));
        unsigned wmax = __reduce_max_sync(0xffffffffu, __float_as_uint(mx));          // NaN never wins fmaxf; Inf does
        if (lane == 0) s_red[warp] = wmax;
        __syncthreads();                    // also: every warp is past the previous tile's TMEM loads and plane use
        unsigned ex = 0;
#pragma unroll
        for (int i = 0; i < TC5_THREADS / 32; ++i) ex = max(ex, s_red[i]);
        ex >>= 23;
        if (ex == 255u) {                   // a non-finite sample: scale by the largest finite one (block-uniform branch)
            __syncthreads();
            float m2 = 0.f;
            auto fin = [](float c) { const float q = fabsf(c); return q <= 3.4028234e38f ? q : 0.f; };
#pragma unroll
            for (int u = 0; u < TC5_NLD; ++u)
                m2 = fmaxf(fmaxf(m2, fmaxf(fin(v[u].x), fin(v[u].y))), fmaxf(fin(v[u].z), fin(v[u].w)));
            wmax = __reduce_max_sync(0xffffffffu, __float_as_uint(m2));
            if (lane == 0) s_red[warp] = wmax;
            __syncthreads();
            ex = 0;
#pragma unroll
            for (int i = 0; i < TC5_THREADS / 32; ++i) ex = max(ex, s_red[i]);
            ex >>= 23;
        }
        const bool scaled = ex >= 14u && ex < 255u;
        const float sc = scaled ? __uint_as_float((267u - ex) << 23) : 1.0f;          // 2^(13 - (ex - 127))
        const float isc = scaled ? __uint_as_float((ex - 13u) << 23) : 1.0f;
        const float inv = isc * a.tap_inv_scale;
        TC5_STAMP(1);

        // ---- B. split into the four swizzled fp16 planes
        {
            const unsigned col = ((unsigned)(lane & 3)) << 2;
#pragma unroll
            for (int u = 0; u < TC5_NLD; ++u) {
                const int row = warp + 8 * u;
                if (row < TC5_ROWS) {
                    unsigned rh, rl, ih, il;
                    tc5_split2(v[u].x * sc, v[u].z * sc, rh, rl);
                    tc5_split2(v[u].y * sc, v[u].w * sc, ih, il);
                    unsigned char* p = s_planes + row * 128 + ((((unsigned)lane >> 2) ^ ((unsigned)row & 7u)) << 4) + col;
                    *reinterpret_cast<unsigned*>(p) = rh;
                    *reinterpret_cast<unsigned*>(p + TC5_PLANE) = rl;
                    *reinterpret_cast<unsigned*>(p + 2 * TC5_PLANE) = ih;
                    *reinterpret_cast<unsigned*>(p + 3 * TC5_PLANE) = il;
                }
            }
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // generic-proxy stores -> visible to the tensor core's reads
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        TC5_STAMP(2);
        __syncthreads();
        TC5_STAMP(3);

        // ---- C. one thread issues the tile's MMAs; completion -> mbarrier
        if (warp == 0) {
            if (tc5_elect()) {
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const unsigned a_hi1 = TC5_DESC_HI | (((unsigned)a.base_off & 7u) << 17);     // base offset: bits [49,52)
                const unsigned pa = tc5_desc_lo(planes_u), pb = tc5_desc_lo(b_u);
                constexpr unsigned PL16 = TC5_PLANE / 16;
#pragma unroll
                for (int s = 0; s < 8; ++s) {                   // re and im accumulators alternate: consecutive MMAs are independent
                    if (s < a.KS) {
                        const unsigned half = (unsigned)s >> 2, ko = ((unsigned)s & 3u) * 2u;  // all offsets in 16-byte units
                        const unsigned ah = half ? a_hi1 : TC5_DESC_HI;
                        const unsigned xa = pa + half * 8u + ko;
                        const unsigned b_hi = pb + half * 512u + ko, b_lo = b_hi + 1024u;
                        const unsigned acc = s ? 1u : 0u;
                        tc5_mma(tmem, xa, ah, b_hi, acc);
                        tc5_mma(tmem + 64u, xa + 2 * PL16, ah, b_hi, acc);
                        tc5_mma(tmem, xa, ah, b_lo, 1u);
                        tc5_mma(tmem + 64u, xa + 2 * PL16, ah, b_lo, 1u);
                        tc5_mma(tmem, xa + PL16, ah, b_hi, 1u);
                        tc5_mma(tmem + 64u, xa + 3 * PL16, ah, b_hi, 1u);
                    }
                }
                asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(bar_u) : "memory");
            }
        }
        __syncwarp();
        TC5_STAMP(4);

        // ---- D. next tile's loads go out while the tensor core works
        const long long ch = tile / a.tiles_x, tx = tile - ch * a.tiles_x;
        const long long next = tile + gridDim.x;
        if (next < a.total_tiles) tc5_load(a, next, warp, lane, v);

        TC5_STAMP(5);
        // ---- E. wait for the accumulators, scale, store
        {
            unsigned done = 0;
            while (!done) {
                asm volatile(
                    "{\n\t.reg .pred p;\n\t"
                    "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                    "selp.u32 %0, 1, 0, p;\n\t}"
                    : "=r"(done) : "r"(bar_u), "r"(phase) : "memory");
            }
            phase ^= 1u;
        }
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        TC5_STAMP(6);
        {
            const long long o0 = tx * FIR_TC5_BT;
            float2* out = a.out + ch * a.out_stride + o0;
            const long long cnt = a.out_n - o0;                                    // outputs of this tile that exist (may exceed 8192)
            const bool fast = cnt >= FIR_TC5_BT && (reinterpret_cast<uintptr_t>(out) & 15) == 0;
            const int q = warp & 3, chalf = warp >> 2;
            const int i = lane >> 2, t = lane & 3;
#pragma unroll
            for (int hh = 0; hh < 2; ++hh) {
                unsigned re[16], im[16];
                const unsigned taddr = tmem + ((unsigned)(32 * q + 16 * hh) << 16) + 32u * chalf;
                tc5_ld16(taddr, re);
                tc5_ld16(taddr + 64u, im);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                for (int g = 0; g < 4; ++g)
#pragma unroll
                    for (int v1 = 0; v1 < 2; ++v1) {
                        const int row = 32 * q + 16 * hh + 8 * v1 + i;
                        const int c = 32 * chalf + 8 * g + 2 * t;
                        const int r0 = 4 * g + 2 * v1;
                        const float4 y = make_float4(__uint_as_float(re[r0]) * inv, __uint_as_float(im[r0]) * inv,
                                                     __uint_as_float(re[r0 + 1]) * inv, __uint_as_float(im[r0 + 1]) * inv);
                        const long long o = 64ll * row + c;
                        if (fast) {
                            *reinterpret_cast<float4*>(out + o) = y;
                        } else {
                            if (o < cnt) out[o] = make_float2(y.x, y.y);
                            if (o + 1 < cnt) out[o + 1] = make_float2(y.z, y.w);
                        }
                    }
            }
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        TC5_STAMP(7);
    }
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 128;" :: "r"(tmem) : "memory");
}


// ------------------------------------------------------------------------------------------------------------------
// Pipelined form: one CTA per SM, warp-specialised.  Two producer groups of 8 warps stage alternate tiles into two
// plane sets (global loads of a group's next tile are in flight while it waits and converts), one warp issues the
// MMAs into two TMEM accumulator sets, four warps (one per TMEM lane quarter) drain finished accumulators to global
// memory.  mbarriers: full[g] (256 producer arrivals) -> MMA warp; done[g] (tcgen05.commit) -> epilogue warps and the
// producer group that reuses plane set g; accfree[g] (128 epilogue arrivals) -> MMA warp.
constexpr int TC5P_THREADS = (2 * 8 + 4 + 1) * 32;
constexpr size_t TC5P_SMEM = 1024 + 8 * TC5_PLANE + TC5_BIMG + 512;
constexpr int TC5P_NSTAMP = 8, TC5P_TRACE_IT = 6;

__device__ __forceinline__ void mbar_init(unsigned bar, unsigned count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(bar), "r"(count) : "memory"); }
__device__ __forceinline__ void mbar_arrive(unsigned bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(bar) : "memory"); }
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity) {
    unsigned done = 0;
    while (!done) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    }
}
__device__ __forceinline__ float4 ldg_stream(const float4* p) {
    float4 r;
    asm("ld.global.nc.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
    return r;
}

__device__ __forceinline__ void tc5p_load(const FirTc5Args& a, long long tile, int pw, int lane, float4 (&v)[TC5_NLD]) {
    const long long ch = tile / a.tiles_x, tx = tile - ch * a.tiles_x;
    const long long s0 = tx * FIR_TC5_BT;
    const float2* in = a.in + ch * a.in_stride + s0;
    const long long avail = a.need - s0;
    const bool fast = avail >= 64ll * TC5_ROWS && (reinterpret_cast<uintptr_t>(in) & 15) == 0;
    if (fast) {
        const float4* p = reinterpret_cast<const float4*>(in) + lane;
#pragma unroll
        for (int u = 0; u < TC5_NLD; ++u) {
            const int row = pw + 8 * u;
            v[u] = row < TC5_ROWS ? ldg_stream(p + row * 32) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
    } else {
#pragma unroll
        for (int u = 0; u < TC5_NLD; ++u) {
            const long long s = 64ll * (pw + 8 * u) + 2 * lane;
            v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (pw + 8 * u < TC5_ROWS) {
                if (s < avail) { const float2 p = __ldg(in + s); v[u].x = p.x; v[u].y = p.y; }
                if (s + 1 < avail) { const float2 q = __ldg(in + s + 1); v[u].z = q.x; v[u].w = q.y; }
            }
        }
    }
}

#define TC5P_STAMP(role, k) do { if (trace && blockIdx.x == 2 && j >= 2 && j < 2 + TC5P_TRACE_IT && lane == 0) trace[((j - 2) * 4 + (role)) * TC5P_NSTAMP + (k)] = clock64(); } while (0)

__global__ void __launch_bounds__(TC5P_THREADS, 1) fir_tc5p_kernel(const FirTc5Args a, long long* __restrict__ trace) {
    extern __shared__ unsigned char smem_raw[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const unsigned raw = smem_u32(smem_raw);
    unsigned char* sm = smem_raw + ((1024u - (raw & 1023u)) & 1023u);
    unsigned char* s_planes = sm;                                          // [2 sets][4 planes]
    unsigned char* s_b = sm + 8 * TC5_PLANE;
    unsigned long long* s_bar = reinterpret_cast<unsigned long long*>(s_b + TC5_BIMG);   // full[2], done[2], accfree[2]
    unsigned* s_tmem = reinterpret_cast<unsigned*>(s_bar + 6);
    float* s_inv = reinterpret_cast<float*>(s_tmem + 1);                   // [4]
    unsigned* s_red = reinterpret_cast<unsigned*>(s_inv + 4);              // [2 groups][3][8]
    const unsigned planes_u = smem_u32(s_planes), b_u = smem_u32(s_b), bar_u = smem_u32(s_bar);
    const unsigned full_u = bar_u, done_u = bar_u + 16, accfree_u = bar_u + 32;

    {
        const uint4* src = a.bimg;
        uint4* dst = reinterpret_cast<uint4*>(s_b);
        for (int i = tid; i < TC5_BIMG / 16; i += TC5P_THREADS) dst[i] = __ldg(src + i);
    }
    if (tid == 0) {
        for (unsigned g = 0; g < 2; ++g) { mbar_init(full_u + 8 * g, 256); mbar_init(done_u + 8 * g, 1); mbar_init(accfree_u + 8 * g, 128); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 20) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 256;" :: "r"(smem_u32(s_tmem)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");          // the B image was written through the generic proxy
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const unsigned tmem = *s_tmem;
    const long long first = blockIdx.x;
    const int njobs = first < a.total_tiles ? (int)((a.total_tiles - first + gridDim.x - 1) / gridDim.x) : 0;

    if (warp < 16) {
        // ================= producers: group g stages tiles j = g, g + 2, ... into plane set g
        const int g = warp >> 3, pw = warp & 7;
        unsigned char* planes = s_planes + g * 4 * TC5_PLANE;
        unsigned* red = s_red + g * 24;
        float4 v[TC5_NLD];
        if (g < njobs) tc5p_load(a, first + (long long)g * gridDim.x, pw, lane, v);
        for (int j = g, u = 0; j < njobs; j += 2, ++u) {
            TC5P_STAMP(g, 0);
            float mx = 0.f;
#pragma unroll
            for (int i = 0; i < TC5_NLD; ++i)
                mx = fmaxf(fmaxf(mx, fmaxf(fabsf(v[i].x), fabsf(v[i].y))), fmaxf(fabsf(v[i].z), fabsf(v[i].w)));
            unsigned wmax = __reduce_max_sync(0xffffffffu, __float_as_uint(mx));      // NaN never wins fmaxf; Inf does
            unsigned* rb = red + 8 * (u & 1);
            if (lane == 0) rb[pw] = wmax;
            asm volatile("bar.sync %0, 256;" :: "r"(1 + g) : "memory");
            unsigned ex = 0;
#pragma unroll
            for (int i = 0; i < 8; ++i) ex = max(ex, rb[i]);
            ex >>= 23;
            if (ex == 255u) {               // a non-finite sample: scale by the largest finite one (group-uniform branch)
                float m2 = 0.f;
                auto fin = [](float c) { const float q = fabsf(c); return q <= 3.4028234e38f ? q : 0.f; };
#pragma unroll
                for (int i = 0; i < TC5_NLD; ++i)
                    m2 = fmaxf(fmaxf(m2, fmaxf(fin(v[i].x), fin(v[i].y))), fmaxf(fin(v[i].z), fin(v[i].w)));
                wmax = __reduce_max_sync(0xffffffffu, __float_as_uint(m2));
                if (lane == 0) red[16 + pw] = wmax;
                asm volatile("bar.sync %0, 256;" :: "r"(1 + g) : "memory");
                ex = 0;
#pragma unroll
                for (int i = 0; i < 8; ++i) ex = max(ex, red[16 + i]);
                ex >>= 23;
            }
            const bool scaled = ex >= 14u && ex < 255u;
            const float sc = scaled ? __uint_as_float((267u - ex) << 23) : 1.0f;      // 2^(13 - (ex - 127))
            const float isc = scaled ? __uint_as_float((ex - 13u) << 23) : 1.0f;
            TC5P_STAMP(g, 1);
            if (u >= 1) mbar_wait(done_u + 8 * g, (unsigned)(u - 1) & 1u);            // the MMAs of tile j - 2 have read plane set g
            TC5P_STAMP(g, 2);
            {
                const unsigned col = ((unsigned)(lane & 3)) << 2;
#pragma unroll
                for (int i = 0; i < TC5_NLD; ++i) {
                    const int row = pw + 8 * i;
                    if (row < TC5_ROWS) {
                        unsigned rh, rl, ih, il;
                        tc5_split2(v[i].x * sc, v[i].z * sc, rh, rl);
                        tc5_split2(v[i].y * sc, v[i].w * sc, ih, il);
                        unsigned char* p = planes + row * 128 + ((((unsigned)lane >> 2) ^ ((unsigned)row & 7u)) << 4) + col;
                        *reinterpret_cast<unsigned*>(p) = rh;
                        *reinterpret_cast<unsigned*>(p + TC5_PLANE) = rl;
                        *reinterpret_cast<unsigned*>(p + 2 * TC5_PLANE) = ih;
                        *reinterpret_cast<unsigned*>(p + 3 * TC5_PLANE) = il;
                    }
                }
            }
            if (pw == 0 && lane == 0) s_inv[j & 3] = isc * a.tap_inv_scale;
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            mbar_arrive(full_u + 8 * g);
            TC5P_STAMP(g, 3);
            if (j + 2 < njobs) tc5p_load(a, first + (long long)(j + 2) * gridDim.x, pw, lane, v);
            TC5P_STAMP(g, 4);
        }
    } else if (warp < 20) {
        // ================= epilogue: warp q drains TMEM lanes [32 q, 32 q + 32) of finished accumulator sets
        const int q = warp & 3, i = lane >> 2, t = lane & 3;
        for (int j = 0; j < njobs; ++j) {
            const int g = j & 1, u = j >> 1;
            if (q == 0) TC5P_STAMP(3, 0);
            mbar_wait(done_u + 8 * g, (unsigned)u & 1u);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            if (q == 0) TC5P_STAMP(3, 1);
            const float inv = s_inv[j & 3];
            const long long tile = first + (long long)j * gridDim.x;
            const long long ch = tile / a.tiles_x, tx = tile - ch * a.tiles_x;
            const long long o0 = tx * FIR_TC5_BT;
            float2* out = a.out + ch * a.out_stride + o0;
            const long long cnt = a.out_n - o0;
            const bool fast = cnt >= FIR_TC5_BT && (reinterpret_cast<uintptr_t>(out) & 15) == 0;
#pragma unroll 1
            for (int part = 0; part < 4; ++part) {
                const int hh = part >> 1, chalf = part & 1;
                unsigned re[16], im[16];
                const unsigned taddr = tmem + 128u * g + ((unsigned)(32 * q + 16 * hh) << 16) + 32u * chalf;
                tc5_ld16(taddr, re);
                tc5_ld16(taddr + 64u, im);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                for (int gg = 0; gg < 4; ++gg)
#pragma unroll
                    for (int v1 = 0; v1 < 2; ++v1) {
                        const int row = 32 * q + 16 * hh + 8 * v1 + i;
                        const int c = 32 * chalf + 8 * gg + 2 * t;
                        const int r0 = 4 * gg + 2 * v1;
                        const float4 y = make_float4(__uint_as_float(re[r0]) * inv, __uint_as_float(im[r0]) * inv,
                                                     __uint_as_float(re[r0 + 1]) * inv, __uint_as_float(im[r0 + 1]) * inv);
                        const long long o = 64ll * row + c;
                        if (fast) {
                            *reinterpret_cast<float4*>(out + o) = y;
                        } else {
                            if (o < cnt) out[o] = make_float2(y.x, y.y);
                            if (o + 1 < cnt) out[o + 1] = make_float2(y.z, y.w);
                        }
                    }
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            mbar_arrive(accfree_u + 8 * g);
            if (q == 0) TC5P_STAMP(3, 2);
        }
    } else {
        // ================= MMA warp
        const unsigned a_hi1 = TC5_DESC_HI | (((unsigned)a.base_off & 7u) << 17);
        constexpr unsigned PL16 = TC5_PLANE / 16;
        const unsigned pb = tc5_desc_lo(b_u);
        for (int j = 0; j < njobs; ++j) {
            const int g = j & 1, u = j >> 1;
            TC5P_STAMP(2, 0);
            mbar_wait(full_u + 8 * g, (unsigned)u & 1u);
            TC5P_STAMP(2, 1);
            if (u >= 1) mbar_wait(accfree_u + 8 * g, (unsigned)(u - 1) & 1u);
            TC5P_STAMP(2, 2);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            if (tc5_elect()) {
                const unsigned pa = tc5_desc_lo(planes_u + (unsigned)g * 4u * TC5_PLANE);
                const unsigned d = tmem + 128u * (unsigned)g;
#pragma unroll
                for (int s = 0; s < 8; ++s) {
                    if (s < a.KS) {
                        const unsigned half = (unsigned)s >> 2, ko = ((unsigned)s & 3u) * 2u;  // offsets in 16-byte units
                        const unsigned ah = half ? a_hi1 : TC5_DESC_HI;
                        const unsigned xa = pa + half * 8u + ko;
                        const unsigned b_hi = pb + half * 512u + ko, b_lo = b_hi + 1024u;
                        const unsigned acc = s ? 1u : 0u;
                        tc5_mma(d, xa, ah, b_hi, acc);
                        tc5_mma(d + 64u, xa + 2 * PL16, ah, b_hi, acc);
                        tc5_mma(d, xa, ah, b_lo, 1u);
                        tc5_mma(d + 64u, xa + 2 * PL16, ah, b_lo, 1u);
                        tc5_mma(d, xa + PL16, ah, b_hi, 1u);
                        tc5_mma(d + 64u, xa + 3 * PL16, ah, b_hi, 1u);
                    }
                }
                asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(done_u + 8 * g) : "memory");
            }
            __syncwarp();
            TC5P_STAMP(2, 3);
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 20) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 256;" :: "r"(tmem) : "memory");
}


// ------------------------------------------------------------------------------------------------------------------
// Tap-stationary form (the one that wins): the TAPS are the tcgen05 A operand and live in tensor memory for the whole
// kernel, the samples are the B operand, D[m][n] = sum_k A[m][k] * B[n][k] with
//     A[m][k] = w'[k - m]        m < 128 outputs of a block-row, k < 192          (TMEM, written once, hi and lo parts)
//     B[n][k] = z[128 n + k]     n < 64 block-rows of 128 samples                 (shared memory, SWIZZLE_128B K-major)
// so an MMA reads only 64 rows x 32 B of shared memory (the operand reads of the sample-stationary form above were
// what saturated the shared-memory / L1 data path: 6 KB per MMA, 48 cycles each, next to the staging stores and the
// global loads and stores).  A block-row of 128 samples is two 128-byte atom rows kept in two regions (samples 0-63
// and 64-127 of every row); k-steps 0-3 read region 0, 4-7 region 1, 8-11 region 0 advanced by one row.  The
// accumulator lane is the output index inside the block-row, its column the block-row: tcgen05.ld.32x32b gives every
// lane one output of 16 consecutive block-rows, a warp stores 256 contiguous bytes per block-row.
constexpr int TC5T_REGION = 9 * 1024;               // 65 rows x 128 B, rounded up to the 1024-byte atom
constexpr int TC5T_PLANE = 2 * TC5T_REGION;
constexpr size_t TC5T_SMEM = 1024 + 8 * TC5T_PLANE + 4096;
constexpr int TC5T_NPW = 9;                         // warps per producer group
constexpr int TC5T_NLD = (TC5_ROWS + TC5T_NPW - 1) / TC5T_NPW;   // half-rows (float4 loads) per producer lane
constexpr int TC5T_EPI0 = 2 * TC5T_NPW, TC5T_MMAW = TC5T_EPI0 + 4;
constexpr int TC5T_THREADS = (TC5T_MMAW + 1) * 32;  // two producer groups, 4 epilogue warps (one per TMEM lane quarter), 1 MMA warp
constexpr int TC5T_TAB = 160;                       // fp16x2 words per tap table (even / odd alignment, hi / lo part)

__device__ __forceinline__ void tc5_mma_ts(unsigned d_tmem, unsigned a_tmem, unsigned b_lo, unsigned accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .b64 db;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "mov.b64 db, {%2, %5};\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], db, %3, p;\n\t}"
        :: "r"(d_tmem), "r"(a_tmem), "r"(b_lo), "r"(TC5_IDESC), "r"(accumulate), "r"(TC5_DESC_HI) : "memory");
}
__device__ __forceinline__ void tc5_ld32(unsigned taddr, unsigned (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tc5_st8(unsigned taddr, const unsigned (&r)[8]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
                 :: "r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]) : "memory");
}

__device__ __forceinline__ void tc5t_load(const FirTc5Args& a, long long tile, int pw, int lane, float4 (&v)[TC5T_NLD]) {
    const long long ch = tile / a.tiles_x, tx = tile - ch * a.tiles_x;
    const long long s0 = tx * FIR_TC5_BT;
    const float2* in = a.in + ch * a.in_stride + s0;
    const long long avail = a.need - s0;
    const bool fast = avail >= 64ll * TC5_ROWS && (reinterpret_cast<uintptr_t>(in) & 15) == 0;
    if (fast) {
        const float4* p = reinterpret_cast<const float4*>(in) + lane;
#pragma unroll
        for (int u = 0; u < TC5T_NLD; ++u) {
            const int hr = pw + TC5T_NPW * u;
            v[u] = hr < TC5_ROWS ? ldg_stream(p + hr * 32) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
    } else {
#pragma unroll
        for (int u = 0; u < TC5T_NLD; ++u) {
            const long long s = 64ll * (pw + TC5T_NPW * u) + 2 * lane;
            v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (pw + TC5T_NPW * u < TC5_ROWS) {
                if (s < avail) { const float2 p = __ldg(in + s); v[u].x = p.x; v[u].y = p.y; }
                if (s + 1 < avail) { const float2 q = __ldg(in + s + 1); v[u].z = q.x; v[u].w = q.y; }
            }
        }
    }
}

#define TC5T_STAMP(role, k) do { if (trace && blockIdx.x == 2 && j >= 2 && j < 2 + TC5P_TRACE_IT && lane == 0) trace[((j - 2) * 4 + (role)) * TC5P_NSTAMP + (k)] = clock64(); } while (0)

__global__ void __launch_bounds__(TC5T_THREADS, 1) fir_tc5t_kernel(const FirTc5Args a, long long* __restrict__ trace) {
    extern __shared__ unsigned char smem_raw[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    long long* const ktr = (trace && blockIdx.x == 2) ? trace + TC5P_TRACE_IT * 4 * TC5P_NSTAMP : nullptr;   // whole-kernel stamps
    if (ktr && tid == 0) ktr[0] = clock64();
    const unsigned raw = smem_u32(smem_raw);
    unsigned char* sm = smem_raw + ((1024u - (raw & 1023u)) & 1023u);
    unsigned char* s_planes = sm;                                          // [2 sets][4 planes][2 regions]
    unsigned long long* s_bar = reinterpret_cast<unsigned long long*>(sm + 8 * TC5T_PLANE);   // full[2], done[2], accfree[2], taps
    unsigned* s_tmem = reinterpret_cast<unsigned*>(s_bar + 7);
    float* s_inv = reinterpret_cast<float*>(s_tmem + 1);                   // [4]
    unsigned* s_red = reinterpret_cast<unsigned*>(s_inv + 4);              // [2 groups][3][16]
    unsigned* s_tab = s_red + 96;                                          // [hi, lo][even, odd][TC5T_TAB]
    const unsigned planes_u = smem_u32(s_planes), bar_u = smem_u32(s_bar);
    const unsigned full_u = bar_u, done_u = bar_u + 16, accfree_u = bar_u + 32, taps_u = bar_u + 48;

    const long long first = blockIdx.x;
    const int njobs = first < a.total_tiles ? (int)((a.total_tiles - first + gridDim.x - 1) / gridDim.x) : 0;
    if (tid == 0) {
        for (unsigned g = 0; g < 2; ++g) { mbar_init(full_u + 8 * g, 32 * TC5T_NPW); mbar_init(done_u + 8 * g, 1); mbar_init(accfree_u + 8 * g, 128); }
        mbar_init(taps_u, 128);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();                        // the only CTA-wide barrier before the tail: the mbarriers exist
    if (ktr && tid == 0) ktr[1] = clock64();
    // Set-up off the producers' path: the MMA warp allocates tensor memory while the epilogue warps stage the tap words in
    // shared memory; those 160 threads meet on named barrier 3, the epilogue warps write the tap operand into TMEM and
    // arrive on `taps`, which the MMA warp waits for before its first MMA.  The producers go straight to their first tile.
    unsigned tmem = 0;
    if (warp == TC5T_MMAW) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" :: "r"(smem_u32(s_tmem)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
        if (ktr && lane == 0) ktr[9] = clock64();
    }
    if (warp >= TC5T_EPI0) {
        if (warp < TC5T_MMAW)
            for (int i = tid - 32 * TC5T_EPI0; i < 4 * TC5T_TAB; i += 128) s_tab[i] = __ldg(reinterpret_cast<const unsigned*>(a.bimg) + i);
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        asm volatile("bar.sync 3, 160;" ::: "memory");
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        tmem = *s_tmem;
        if (ktr && warp == TC5T_EPI0 && lane == 0) ktr[10] = clock64();
    }
    if (warp >= TC5T_EPI0 && warp < TC5T_MMAW) {          // the tap operand: lane m, column c of part p holds (w'[2c - m], w'[2c + 1 - m]) as fp16x2
        const int m = 32 * (warp & 3) + lane;
        // host tables of fp16x2 words: E[q] = (w'[2q], w'[2q + 1]), O[q] = (w'[2q + 1], w'[2q + 2]), q + 64 in [0, 160)
        const unsigned* tab = s_tab + (m & 1) * TC5T_TAB - ((m + 1) >> 1) + 64;
        for (int part = 0; part < 2; ++part)
            for (int c8 = 0; c8 < 12; ++c8) {
                unsigned r[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) r[i] = tab[part * 2 * TC5T_TAB + 8 * c8 + i];
                tc5_st8(tmem + ((unsigned)(32 * (warp & 3)) << 16) + 96u * part + 8u * c8, r);
            }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        mbar_arrive(taps_u);
        if (ktr && warp == TC5T_EPI0 && lane == 0) ktr[11] = clock64();
    }

    if (warp < TC5T_EPI0) {
        // ================= producers: group g stages tiles j = g, g + 2, ... into plane set g
        const int g = warp >= TC5T_NPW ? 1 : 0, pw = warp - g * TC5T_NPW;
        float4 v[TC5T_NLD];
        if (g < njobs) tc5t_load(a, first + (long long)g * gridDim.x, pw, lane, v);
        if (ktr && warp == 0 && lane == 0) ktr[8] = clock64();
        unsigned char* planes = s_planes + g * 4 * TC5T_PLANE;
        unsigned* red = s_red + g * 48;
        for (int j = g, u = 0; j < njobs; j += 2, ++u) {
            TC5T_STAMP(g, 0);
            float mx = 0.f;
#pragma unroll
            for (int i = 0; i < TC5T_NLD; ++i)
                mx = fmaxf(fmaxf(mx, fmaxf(fabsf(v[i].x), fabsf(v[i].y))), fmaxf(fabsf(v[i].z), fabsf(v[i].w)));
            unsigned wmax = __reduce_max_sync(0xffffffffu, __float_as_uint(mx));      // NaN never wins fmaxf; Inf does
            unsigned* rb = red + 16 * (u & 1);
            if (lane == 0) rb[pw] = wmax;
            asm volatile("bar.sync %0, %1;" :: "r"(1 + g), "n"(32 * TC5T_NPW) : "memory");
            unsigned ex = 0;
#pragma unroll
            for (int i = 0; i < TC5T_NPW; ++i) ex = max(ex, rb[i]);
            ex >>= 23;
            if (ex == 255u) {               // a non-finite sample: scale by the largest finite one (group-uniform branch)
                float m2 = 0.f;
                auto fin = [](float c) { const float q = fabsf(c); return q <= 3.4028234e38f ? q : 0.f; };
#pragma unroll
                for (int i = 0; i < TC5T_NLD; ++i)
                    m2 = fmaxf(fmaxf(m2, fmaxf(fin(v[i].x), fin(v[i].y))), fmaxf(fin(v[i].z), fin(v[i].w)));
                wmax = __reduce_max_sync(0xffffffffu, __float_as_uint(m2));
                if (lane == 0) red[32 + pw] = wmax;
                asm volatile("bar.sync %0, %1;" :: "r"(1 + g), "n"(32 * TC5T_NPW) : "memory");
                ex = 0;
#pragma unroll
                for (int i = 0; i < TC5T_NPW; ++i) ex = max(ex, red[32 + i]);
                ex >>= 23;
            }
            const bool scaled = ex >= 14u && ex < 255u;
            const float sc = scaled ? __uint_as_float((267u - ex) << 23) : 1.0f;      // 2^(13 - (ex - 127))
            const float isc = scaled ? __uint_as_float((ex - 13u) << 23) : 1.0f;
            TC5T_STAMP(g, 1);
            if (u >= 1) mbar_wait(done_u + 8 * g, (unsigned)(u - 1) & 1u);            // the MMAs of tile j - 2 have read plane set g
            TC5T_STAMP(g, 2);
            {
                const unsigned col = ((unsigned)(lane & 3)) << 2, c8 = (unsigned)lane >> 2;
#pragma unroll
                for (int i = 0; i < TC5T_NLD; ++i) {
                    const unsigned hr = (unsigned)(pw + TC5T_NPW * i);                 // half-row: 64 samples
                    if (hr < (unsigned)TC5_ROWS) {
                        unsigned rh, rl, ih, il;
                        tc5_split2(v[i].x * sc, v[i].z * sc, rh, rl);
                        tc5_split2(v[i].y * sc, v[i].w * sc, ih, il);
                        const unsigned n = hr >> 1;
                        unsigned char* p = planes + (hr & 1u) * TC5T_REGION + n * 128u + ((c8 ^ (n & 7u)) << 4) + col;
                        *reinterpret_cast<unsigned*>(p) = rh;
                        *reinterpret_cast<unsigned*>(p + TC5T_PLANE) = rl;
                        *reinterpret_cast<unsigned*>(p + 2 * TC5T_PLANE) = ih;
                        *reinterpret_cast<unsigned*>(p + 3 * TC5T_PLANE) = il;
                    }
                }
            }
            if (pw == 0 && lane == 0) s_inv[j & 3] = isc * a.tap_inv_scale;
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            mbar_arrive(full_u + 8 * g);
            TC5T_STAMP(g, 3);
            if (j + 2 < njobs) tc5t_load(a, first + (long long)(j + 2) * gridDim.x, pw, lane, v);
            TC5T_STAMP(g, 4);
        }
        if (ktr && lane == 0 && pw == 0) ktr[2 + g] = clock64();
    } else if (warp < TC5T_MMAW) {
        // ================= epilogue: warp q holds outputs 32 q + lane of every block-row
        const int q = warp & 3;
        for (int j = 0; j < njobs; ++j) {
            const int g = j & 1, u = j >> 1;
            if (q == 0) TC5T_STAMP(3, 0);
            mbar_wait(done_u + 8 * g, (unsigned)u & 1u);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            if (q == 0) TC5T_STAMP(3, 1);
            const float inv = s_inv[j & 3];
            const long long tile = first + (long long)j * gridDim.x;
            const long long ch = tile / a.tiles_x, tx = tile - ch * a.tiles_x;
            const long long o0 = tx * FIR_TC5_BT;
            float2* out = a.out + ch * a.out_stride + o0 + 32 * q + lane;
            const long long cnt = a.out_n - o0 - (32 * q + lane);                   // this lane's outputs exist for 128 n < cnt
            const bool fast = a.out_n - o0 >= FIR_TC5_BT;
#pragma unroll 1
            for (int part = 0; part < 4; ++part) {
                unsigned re[16], im[16];
                const unsigned taddr = tmem + 256u + 128u * g + ((unsigned)(32 * q) << 16) + 16u * part;
                tc5_ld32(taddr, re);
                tc5_ld32(taddr + 64u, im);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                for (int r = 0; r < 16; ++r) {
                    const long long o = 128ll * (16 * part + r);
                    if (fast || o < cnt) out[o] = make_float2(__uint_as_float(re[r]) * inv, __uint_as_float(im[r]) * inv);
                }
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            mbar_arrive(accfree_u + 8 * g);
            if (q == 0) TC5T_STAMP(3, 2);
            if (ktr && q == 0 && lane == 0 && j < 2) ktr[6 + j] = clock64();
        }
        if (ktr && q == 0 && lane == 0) ktr[4] = clock64();
    } else {
        // ================= MMA warp
        constexpr unsigned PL16 = TC5T_PLANE / 16, RG16 = TC5T_REGION / 16;
        mbar_wait(taps_u, 0u);
        for (int j = 0; j < njobs; ++j) {
            const int g = j & 1, u = j >> 1;
            TC5T_STAMP(2, 0);
            mbar_wait(full_u + 8 * g, (unsigned)u & 1u);
            TC5T_STAMP(2, 1);
            if (u >= 1) mbar_wait(accfree_u + 8 * g, (unsigned)(u - 1) & 1u);          // tile j - 2 has left accumulator set g
            TC5T_STAMP(2, 2);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            if (tc5_elect()) {
                const unsigned pb = tc5_desc_lo(planes_u + (unsigned)g * 4u * TC5T_PLANE);
                const unsigned d = tmem + 256u + 128u * (unsigned)g;
#pragma unroll
                for (int s = 0; s < 12; ++s) {
                    if (s < a.KS) {
                        // k-block 0: region 0; 1: region 1; 2: region 0 advanced by one 128-byte row (offsets in 16-byte units)
                        const unsigned kb = (unsigned)s >> 2;
                        const unsigned xb = pb + (kb == 1 ? RG16 : 0u) + (kb == 2 ? 8u : 0u) + ((unsigned)s & 3u) * 2u;
                        const unsigned a_hi = tmem + 8u * s, a_lo = a_hi + 96u;
                        const unsigned acc = s ? 1u : 0u;
                        tc5_mma_ts(d, a_hi, xb, acc);                       // re: hi * hi
                        tc5_mma_ts(d + 64u, a_hi, xb + 2 * PL16, acc);      // im: hi * hi
                        tc5_mma_ts(d, a_hi, xb + PL16, 1u);                 // taps hi * samples lo
                        tc5_mma_ts(d + 64u, a_hi, xb + 3 * PL16, 1u);
                        tc5_mma_ts(d, a_lo, xb, 1u);                        // taps lo * samples hi
                        tc5_mma_ts(d + 64u, a_lo, xb + 2 * PL16, 1u);
                    }
                }
                asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(done_u + 8 * g) : "memory");
            }
            __syncwarp();
            TC5T_STAMP(2, 3);
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == TC5T_MMAW) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" :: "r"(tmem) : "memory");
    if (ktr && tid == 0) ktr[5] = clock64();
}

}  // namespace

size_t fir_tc5_bimg_bytes() { return TC5_BIMG; }
size_t fir_tc5_tab_entries() { return TC5T_TAB; }

// Shared-memory image of B for reversed taps w[0..T) already scaled (fp16 hi / lo parts given as bit patterns by the
// callbacks): part (0 hi, 1 lo), half (k < 64, k >= 64), row n, k-local kk -> byte offset.
size_t fir_tc5_bimg_offset(int part, int half, int n, int kk) {
    return (size_t)part * 16384 + (size_t)half * 8192 + (size_t)n * 128 + (size_t)((((unsigned)kk >> 3) ^ ((unsigned)n & 7u)) << 4) + (size_t)(kk & 7) * 2;
}

namespace {
int fir_tc5p_launch(int device, const FirTc5Args& a, cudaStream_t st, bool ts) {
    static bool ready[16] = {};
    const int dv = (device < 0 || device >= 16) ? 0 : device;
    if (!ready[dv]) {
        RRC_CUDA(cudaFuncSetAttribute(fir_tc5p_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TC5P_SMEM));
        RRC_CUDA(cudaFuncSetAttribute(fir_tc5t_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TC5T_SMEM));
        ready[dv] = true;
    }
    static const bool want_trace = getenv("RRC_FIR_TC5_TRACE") != nullptr;
    long long* dtrace = nullptr;
    const size_t trace_n = (size_t)TC5P_TRACE_IT * 4 * TC5P_NSTAMP + 16;
    if (want_trace) { RRC_CUDA(cudaMalloc((void**)&dtrace, trace_n * 8)); RRC_CUDA(cudaMemsetAsync(dtrace, 0, trace_n * 8, st)); }
    const unsigned grid = (unsigned)std::min<long long>(a.total_tiles, sm_count(device));
    if (ts) fir_tc5t_kernel<<<grid, TC5T_THREADS, TC5T_SMEM, st>>>(a, dtrace);
    else fir_tc5p_kernel<<<grid, TC5P_THREADS, TC5P_SMEM, st>>>(a, dtrace);
    RRC_CHECK_LAUNCH();
    count_launch();
    if (want_trace) {                                           // debug only: CTA 2's stamps of tiles 2..7, relative cycles
        std::vector<long long> tr(trace_n);
        RRC_CUDA(cudaStreamSynchronize(st));
        RRC_CUDA(cudaMemcpy(tr.data(), dtrace, trace_n * 8, cudaMemcpyDeviceToHost));
        cudaFree(dtrace);
        static int dumps = 0;
        if (a.total_tiles > 148 * 10 && dumps++ < 2) {
            long long t0 = 0;
            for (long long x : tr) if (x && (!t0 || x < t0)) t0 = x;
            static const char* role[] = {"producers 0: start, scale known, plane set free, staged + arrived, next loads issued",
                                         "producers 1: (same)", "MMA warp: start, planes full, accumulators free, issued + committed",
                                         "epilogue warp 0: start, MMAs done, stored + arrived"};
            for (int j = 0; j < TC5P_TRACE_IT; ++j)
                for (int r = 0; r < 4; ++r) {
                    const long long* p = &tr[(size_t)(j * 4 + r) * TC5P_NSTAMP];
                    if (!p[0]) continue;
                    fprintf(stderr, "tc5p tile %d %-14.14s", j + 2, role[r]);
                    for (int k = 0; k < TC5P_NSTAMP && p[k]; ++k) fprintf(stderr, " %7lld", p[k] - t0);
                    fprintf(stderr, "\n");
                }
            for (int r = 0; r < 4; ++r) fprintf(stderr, "   %s\n", role[r]);
            const long long* kt = &tr[(size_t)TC5P_TRACE_IT * 4 * TC5P_NSTAMP];
            if (kt[0]) fprintf(stderr, "tc5p kernel (CTA 2, %lld tiles in all): start %lld, prologue done %lld, first two tiles stored %lld %lld, producers done %lld %lld, epilogue done %lld, exit %lld\n",
                               (long long)a.total_tiles, kt[0] - t0, kt[1] - t0, kt[6] - t0, kt[7] - t0, kt[2] - t0, kt[3] - t0, kt[4] - t0, kt[5] - t0);
            if (kt[0]) fprintf(stderr, "tc5p prologue: first loads issued (warp 0) %lld, TMEM allocated %lld, barrier 1 passed %lld, taps in TMEM %lld\n", kt[8] - t0, kt[9] - t0, kt[10] - t0, kt[11] - t0);
        }
    }
    return RRC_OK;
}
}  // namespace

int fir_tc5_launch(int device, const FirTc5Args& a, cudaStream_t st) {
    static const int variant = [] { const char* e = getenv("RRC_FIR_TCGEN05"); return e ? atoi(e) : 0; }();
    if (variant != 1) return fir_tc5p_launch(device, a, st, variant != 2);   // 1: simple two-CTAs-per-SM kernel, 2: pipelined, samples as A; else: taps in TMEM
    static int cache[16] = {};
    static const bool want_dbg = getenv("RRC_FIR_TC5_TRACE") != nullptr;
    const int dv = (device < 0 || device >= 16) ? 0 : device;
    if (cache[dv] == 0) {
        RRC_CUDA(cudaFuncSetAttribute(fir_tc5_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TC5_SMEM));
        int per_sm = 0;
        RRC_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fir_tc5_kernel, TC5_THREADS, TC5_SMEM));
        if (per_sm < 1) return fail(RRC_ERR_CUDA, "fir_tc5: kernel does not fit an SM");
        cache[dv] = std::min(per_sm, 4);                 // 128 TMEM columns per CTA: at most 4 CTAs can allocate
        if (want_dbg) fprintf(stderr, "tc5: occupancy API says %d CTAs/SM\n", per_sm);
        if (const char* e = getenv("RRC_FIR_TC5_CTAS")) { const int c = atoi(e); if (c >= 1 && c <= 4) cache[dv] = c; }
    }
    const long long cap = (long long)sm_count(device) * cache[dv];
    static const bool want_trace = getenv("RRC_FIR_TC5_TRACE") != nullptr;
    if (want_trace) fprintf(stderr, "tc5: %d CTAs/SM, grid %lld, smem %zu\n", cache[dv], std::min<long long>(a.total_tiles, cap), TC5_SMEM);
    long long* dtrace = nullptr;
    const size_t trace_n = (size_t)TC5_TRACE_IT * 8 * TC5_NSTAMP;
    if (want_trace) { RRC_CUDA(cudaMalloc((void**)&dtrace, trace_n * 8)); RRC_CUDA(cudaMemsetAsync(dtrace, 0, trace_n * 8, st)); }
    fir_tc5_kernel<<<(unsigned)std::min<long long>(a.total_tiles, cap), TC5_THREADS, TC5_SMEM, st>>>(a, dtrace);
    RRC_CHECK_LAUNCH();
    count_launch();
    if (want_trace) {                                           // debug only: synchronous dump of CTA 2's per-phase cycle table
        std::vector<long long> tr(trace_n);
        RRC_CUDA(cudaStreamSynchronize(st));
        RRC_CUDA(cudaMemcpy(tr.data(), dtrace, trace_n * 8, cudaMemcpyDeviceToHost));
        cudaFree(dtrace);
        static const char* names[] = {"max, barrier, scale", "split + STS + proxy fence", "barrier 2", "MMA issue (thread 0) / syncwarp",
                                      "next tile's LDG issue", "mbarrier wait (MMA done)", "TMEM ld + scale + STG"};
        static int dumps = 0;
        if (a.total_tiles > 148 * 2 * 5 && dumps++ < 2)
            for (int b = 0; b < TC5_TRACE_IT; ++b) {
                long long t0 = tr[(size_t)(b * 8) * TC5_NSTAMP], tend = 0;
                for (int w = 0; w < 8; ++w) { t0 = std::min(t0, tr[(size_t)(b * 8 + w) * TC5_NSTAMP]); tend = std::max(tend, tr[(size_t)(b * 8 + w) * TC5_NSTAMP + 7]); }
                fprintf(stderr, "tc5 trace tile iter %d: total %lld cycles\n", b + 1, tend - t0);
                for (int p = 0; p < 7; ++p) {
                    std::vector<long long> d;
                    for (int w = 0; w < 8; ++w) d.push_back(tr[(size_t)(b * 8 + w) * TC5_NSTAMP + p + 1] - tr[(size_t)(b * 8 + w) * TC5_NSTAMP + p]);
                    std::sort(d.begin(), d.end());
                    fprintf(stderr, "   %-34s min %6lld  med %6lld  max %6lld\n", names[p], d[0], d[4], d[7]);
                }
            }
    }
    return RRC_OK;
}

}  // namespace rrc
