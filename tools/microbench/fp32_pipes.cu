// Microbenchmark ("measure an FMA-chain microbenchmark on the box"): FP32 issue rates on B200.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp32_pipes fp32_pipes.cu && ./fp32_pipes
// Reports warp-instructions / clk / SM and the equivalent FMA lanes for FFMA, FFMA2 (fma.rn.f32x2),
// FADD, FADD2, FMUL2, a complex-MAC pattern, and shared-memory LDS/STS wavefront rates.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
typedef unsigned long long u64;
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("cuda error %s line %d\n", cudaGetErrorString(e), __LINE__); return 1;}}while(0)

template<int MODE> __global__ void __launch_bounds__(512) fp_kernel(float* out, int iters, float seed) {
  // 8 independent chains per thread
  float a[16]; u64 A[8];
  for (int i = 0; i < 16; i++) a[i] = seed + threadIdx.x * 1e-3f + i;
  for (int i = 0; i < 8; i++) { float2 t = make_float2(a[2*i], a[2*i+1]); A[i] = *reinterpret_cast<u64*>(&t); }
  float b = seed * 0.5f, c = seed * 0.25f;
  float2 bb = make_float2(b, c); u64 B = *reinterpret_cast<u64*>(&bb);
  float2 cc = make_float2(c, b); u64 C = *reinterpret_cast<u64*>(&cc);
  for (int it = 0; it < iters; it++) {
    if (MODE == 0) {        // FFMA, 16 chains, 3 distinct source regs
      #pragma unroll
      for (int i = 0; i < 16; i++) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a[i]) : "f"(b), "f"(c));
    } else if (MODE == 1) { // FFMA2, 8 chains (16 FMAs)
      #pragma unroll
      for (int i = 0; i < 8; i++) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(A[i]) : "l"(B), "l"(C));
    } else if (MODE == 2) { // FADD
      #pragma unroll
      for (int i = 0; i < 16; i++) asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(b));
    } else if (MODE == 3) { // FADD2
      #pragma unroll
      for (int i = 0; i < 8; i++) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(A[i]) : "l"(B));
    } else if (MODE == 4) { // FMUL2
      #pragma unroll
      for (int i = 0; i < 8; i++) asm volatile("mul.rn.f32x2 %0, %0, %1;" : "+l"(A[i]) : "l"(B));
    } else if (MODE == 5) { // FFMA with the accumulate-into-other pattern d = a*b + d (2 distinct + acc)
      #pragma unroll
      for (int i = 0; i < 16; i++) asm volatile("fma.rn.f32 %0, %1, %2, %0;" : "+f"(a[i]) : "f"(b), "f"(c));
    } else if (MODE == 6) { // FFMA2 accumulate pattern
      #pragma unroll
      for (int i = 0; i < 8; i++) asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(A[i]) : "l"(B), "l"(C));
    } else if (MODE == 7) { // alternating FFMA + IADD (alu pipe) : do they dual-issue?
      #pragma unroll
      for (int i = 0; i < 8; i++) {
        asm volatile("fma.rn.f32 %0, %1, %2, %0;" : "+f"(a[i]) : "f"(b), "f"(c));
        asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(A[i]) : "l"(B));
      }
    }
  }
  float s = 0;
  for (int i = 0; i < 16; i++) s += a[i];
  for (int i = 0; i < 8; i++) { float2 t = *reinterpret_cast<float2*>(&A[i]); s += t.x + t.y; }
  if (s == 12345.678f) out[0] = s;
}

// shared memory: each thread does W-bit loads, conflict-free, unrolled
template<int BYTES, bool STORE> __global__ void __launch_bounds__(512) smem_kernel(float* out, int iters) {
  extern __shared__ __align__(16) unsigned char sm[];
  const int tid = threadIdx.x;
  for (int i = tid; i < 512 * 16 / 4 * 4; i += 512) reinterpret_cast<float*>(sm)[i] = i;
  __syncthreads();
  float acc = 0;
  for (int it = 0; it < iters; it++) {
    #pragma unroll
    for (int j = 0; j < 8; j++) {
      const int off = (tid * BYTES + j * 512 * BYTES) % (512 * 16 * 4);
      if (BYTES == 4) { if (STORE) *reinterpret_cast<volatile float*>(sm + off) = acc; else acc += *reinterpret_cast<volatile float*>(sm + off); }
      if (BYTES == 8) { if (STORE) { float2 v = make_float2(acc, acc); asm volatile("st.shared.v2.f32 [%0], {%1,%2};" :: "r"((unsigned)__cvta_generic_to_shared(sm + off)), "f"(v.x), "f"(v.y)); }
                        else { float2 v; asm volatile("ld.shared.v2.f32 {%0,%1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"((unsigned)__cvta_generic_to_shared(sm + off))); acc += v.x + v.y; } }
      if (BYTES == 16) { if (STORE) { asm volatile("st.shared.v4.f32 [%0], {%1,%1,%1,%1};" :: "r"((unsigned)__cvta_generic_to_shared(sm + off)), "f"(acc)); }
                        else { float4 v; asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"((unsigned)__cvta_generic_to_shared(sm + off))); acc += v.x + v.y + v.z + v.w; } }
    }
  }
  if (acc == 12345.678f) out[0] = acc;
}

int main() {
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
  int clk_khz = 0; cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
  printf("device %s, %d SMs, nominal clock %d MHz\n", p.name, p.multiProcessorCount, clk_khz / 1000);
  float* out; CK(cudaMalloc(&out, 1024));
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int sms = p.multiProcessorCount;
  const char* names[] = {"FFMA d=d*b+c (16 chains)", "FFMA2 d=d*b+c (8 chains)", "FADD", "FADD2", "FMUL2", "FFMA d=b*c+d", "FFMA2 d=b*c+d", "FFMA + FADD2 interleaved"};
  const int fmas_per_iter[] = {16, 16, 16, 16, 16, 16, 16, 24};
  const int instr_per_iter[] = {16, 8, 16, 8, 8, 16, 8, 16};
  for (int threads : {256, 512}) for (int mode = 0; mode < 8; mode++) {
    const int iters = 20000, grid = sms * (1024 / threads);
    float ms = 0;
    for (int rep = 0; rep < 2; rep++) {
      cudaEventRecord(e0);
      switch (mode) {
        case 0: fp_kernel<0><<<grid, threads>>>(out, iters, 1.0f); break;
        case 1: fp_kernel<1><<<grid, threads>>>(out, iters, 1.0f); break;
        case 2: fp_kernel<2><<<grid, threads>>>(out, iters, 1.0f); break;
        case 3: fp_kernel<3><<<grid, threads>>>(out, iters, 1.0f); break;
        case 4: fp_kernel<4><<<grid, threads>>>(out, iters, 1.0f); break;
        case 5: fp_kernel<5><<<grid, threads>>>(out, iters, 1.0f); break;
        case 6: fp_kernel<6><<<grid, threads>>>(out, iters, 1.0f); break;
        case 7: fp_kernel<7><<<grid, threads>>>(out, iters, 1.0f); break;
      }
      cudaEventRecord(e1); CK(cudaEventSynchronize(e1)); cudaEventElapsedTime(&ms, e0, e1);
    }
    const double warps_per_sm = 1024 / 32;
    const double winstr = (double)iters * instr_per_iter[mode] * warps_per_sm;     // per SM
    const double clk = ms * 1e-3 * clk_khz * 1e3;
    printf("threads/CTA %4d  %-28s %8.3f ms  %6.3f warp-instr/clk/SM  %7.1f fp32-lane-ops/clk/SM  %6.2f Tops/s chip\n", threads, names[mode], ms,
           winstr / clk, (double)iters * fmas_per_iter[mode] * 1024 / clk, (double)iters * fmas_per_iter[mode] * 1024 * sms / (ms * 1e-3) / 1e12);
  }
  // shared memory
  CK(cudaFuncSetAttribute(smem_kernel<4,false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
  const int iters = 20000;
  auto report = [&](const char* nm, int bytes, float ms) {
    const double clk = ms * 1e-3 * clk_khz * 1e3;
    const double b = (double)iters * 8 * 1024 * bytes;   // 2 CTAs x 512 threads per SM
    printf("smem %-10s %8.3f ms  %7.1f B/clk/SM\n", nm, ms, b / clk);
  };
  float ms;
#define RUN(B, S, NM) { for (int rep = 0; rep < 2; rep++) { cudaEventRecord(e0); smem_kernel<B,S><<<sms * 2, 512, 32768>>>(out, iters); cudaEventRecord(e1); CK(cudaEventSynchronize(e1)); cudaEventElapsedTime(&ms, e0, e1);} report(NM, B, ms); }
  RUN(4, false, "LDS.32") RUN(8, false, "LDS.64") RUN(16, false, "LDS.128") RUN(4, true, "STS.32") RUN(8, true, "STS.64") RUN(16, true, "STS.128")
  return 0;
}
