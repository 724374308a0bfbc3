// tmem.cuh — tensor memory (TMEM) as thread-private scratch: tcgen05.ld / tcgen05.st.32x32b moves between a thread's registers and
// consecutive 32-bit columns of ITS OWN TMEM lane (warp w reaches lanes 32 (w & 3) .. + 31).  Used by fftfilt_poly_kernel (running
// sum over the polyphase branches, stashed second branch of a 128-bit gather) and fftfilt_tmh_kernel (the filter spectrum).
#pragma once
#include <cuda_runtime.h>

namespace rrc {

// 32 consecutive 32-bit columns of the thread's own TMEM lane <-> 16 float2
__device__ __forceinline__ void tm_ld32(unsigned taddr, float2 (&x)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n\t"
        "tcgen05.wait::ld.sync.aligned;"
        : "=f"(x[0].x), "=f"(x[0].y), "=f"(x[1].x), "=f"(x[1].y), "=f"(x[2].x), "=f"(x[2].y), "=f"(x[3].x), "=f"(x[3].y),
          "=f"(x[4].x), "=f"(x[4].y), "=f"(x[5].x), "=f"(x[5].y), "=f"(x[6].x), "=f"(x[6].y), "=f"(x[7].x), "=f"(x[7].y),
          "=f"(x[8].x), "=f"(x[8].y), "=f"(x[9].x), "=f"(x[9].y), "=f"(x[10].x), "=f"(x[10].y), "=f"(x[11].x), "=f"(x[11].y),
          "=f"(x[12].x), "=f"(x[12].y), "=f"(x[13].x), "=f"(x[13].y), "=f"(x[14].x), "=f"(x[14].y), "=f"(x[15].x), "=f"(x[15].y)
        : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tm_st32(unsigned taddr, const float2 (&x)[16]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
        :: "r"(taddr),
           "f"(x[0].x), "f"(x[0].y), "f"(x[1].x), "f"(x[1].y), "f"(x[2].x), "f"(x[2].y), "f"(x[3].x), "f"(x[3].y),
           "f"(x[4].x), "f"(x[4].y), "f"(x[5].x), "f"(x[5].y), "f"(x[6].x), "f"(x[6].y), "f"(x[7].x), "f"(x[7].y),
           "f"(x[8].x), "f"(x[8].y), "f"(x[9].x), "f"(x[9].y), "f"(x[10].x), "f"(x[10].y), "f"(x[11].x), "f"(x[11].y),
           "f"(x[12].x), "f"(x[12].y), "f"(x[13].x), "f"(x[13].y), "f"(x[14].x), "f"(x[14].y), "f"(x[15].x), "f"(x[15].y)
        : "memory");
}
__device__ __forceinline__ void tm_st16(unsigned taddr, const float2 (&x)[8]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
        :: "r"(taddr),
           "f"(x[0].x), "f"(x[0].y), "f"(x[1].x), "f"(x[1].y), "f"(x[2].x), "f"(x[2].y), "f"(x[3].x), "f"(x[3].y),
           "f"(x[4].x), "f"(x[4].y), "f"(x[5].x), "f"(x[5].y), "f"(x[6].x), "f"(x[6].y), "f"(x[7].x), "f"(x[7].y)
        : "memory");
}
// Split form: issue the load, do other work, then tm_ld32_wait — which takes the 32 registers as read-write operands so that no
// use of them can be scheduled above the wait.
__device__ __forceinline__ void tm_ld32_issue(unsigned taddr, float2 (&x)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=f"(x[0].x), "=f"(x[0].y), "=f"(x[1].x), "=f"(x[1].y), "=f"(x[2].x), "=f"(x[2].y), "=f"(x[3].x), "=f"(x[3].y),
          "=f"(x[4].x), "=f"(x[4].y), "=f"(x[5].x), "=f"(x[5].y), "=f"(x[6].x), "=f"(x[6].y), "=f"(x[7].x), "=f"(x[7].y),
          "=f"(x[8].x), "=f"(x[8].y), "=f"(x[9].x), "=f"(x[9].y), "=f"(x[10].x), "=f"(x[10].y), "=f"(x[11].x), "=f"(x[11].y),
          "=f"(x[12].x), "=f"(x[12].y), "=f"(x[13].x), "=f"(x[13].y), "=f"(x[14].x), "=f"(x[14].y), "=f"(x[15].x), "=f"(x[15].y)
        : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tm_ld32_wait(float2 (&x)[16]) {
    asm volatile(
        "tcgen05.wait::ld.sync.aligned;"
        : "+f"(x[0].x), "+f"(x[0].y), "+f"(x[1].x), "+f"(x[1].y), "+f"(x[2].x), "+f"(x[2].y), "+f"(x[3].x), "+f"(x[3].y),
          "+f"(x[4].x), "+f"(x[4].y), "+f"(x[5].x), "+f"(x[5].y), "+f"(x[6].x), "+f"(x[6].y), "+f"(x[7].x), "+f"(x[7].y),
          "+f"(x[8].x), "+f"(x[8].y), "+f"(x[9].x), "+f"(x[9].y), "+f"(x[10].x), "+f"(x[10].y), "+f"(x[11].x), "+f"(x[11].y),
          "+f"(x[12].x), "+f"(x[12].y), "+f"(x[13].x), "+f"(x[13].y), "+f"(x[14].x), "+f"(x[14].y), "+f"(x[15].x), "+f"(x[15].y)
        :: "memory");
}
__device__ __forceinline__ void tm_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tm_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tm_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// 512 columns (the whole 256 KiB) for this CTA; call from ONE warp, then publish *slot through a CTA barrier between
// tm_fence_before() / tm_fence_after().
__device__ __forceinline__ void tm_alloc_all(unsigned* slot) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"((unsigned)__cvta_generic_to_shared(slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tm_dealloc_all(unsigned tmem) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
}
// First column of thread (warp, lane)'s private 64-column strip in a 16-warp CTA: lane quarter warp & 3, strip warp >> 2.
__device__ __forceinline__ unsigned tm_strip64(unsigned tmem, int warp) { return tmem + ((unsigned)(32 * (warp & 3)) << 16) + 64u * (unsigned)(warp >> 2); }

}  // namespace rrc
