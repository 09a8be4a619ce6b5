// Issue-rate microbenchmark of the epilogue's instruction mix on one SM sub-partition set (B200): cycles per warp-instruction
// for MUFU.EX2, F2FP.F16.F32.PACK_AB, FHFMA (fma.rn.f32.f16), HADD2.F32 (cvt.f32.f16) and FADD, alone and mixed.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipes pipes.cu && ./pipes
#include <cstdio>
#include <cuda_runtime.h>
#define ITERS 4096
template <int K>
__global__ void bench(float* out, long long* cyc, float seed) {
  float a[8];
  unsigned u[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { a[i] = seed + threadIdx.x * 0.001f + i; u[i] = 0x3c003c00u + i; }
  __syncthreads();
  long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (K == 0) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
      if (K == 1) asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(u[i]) : "f"(a[i]), "f"(a[(i + 1) & 7]));
      if (K == 2) asm volatile("{.reg .b16 l,h; mov.b32 {l,h}, %1; fma.rn.f32.f16 %0, l, h, %0;}" : "+f"(a[i]) : "r"(u[i]));
      if (K == 3) asm volatile("{.reg .b16 l,h; mov.b32 {l,h}, %1; cvt.f32.f16 %0, l;}" : "=f"(a[i]) : "r"(u[i]));
      if (K == 4) asm volatile("add.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(seed));
      if (K == 5) {  // MUFU + F2FP interleaved: shared pipe?
        asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
        asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(u[i]) : "f"(a[i]), "f"(a[(i + 1) & 7]));
      }
      if (K == 6) {  // MUFU + 7 FADD: can the FMA pipe issue between MUFUs?
        asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
#pragma unroll
        for (int j = 0; j < 7; ++j) asm volatile("add.f32 %0, %0, %1;" : "+f"(a[(i + j + 1) & 7]) : "f"(seed));
      }
      if (K == 7) {  // F2FP + FHFMA x2 + F2FP (the split of one pair)
        asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(u[i]) : "f"(a[i]), "f"(a[(i + 1) & 7]));
        asm volatile("{.reg .b16 l,h; mov.b32 {l,h}, %1; fma.rn.f32.f16 %0, l, h, %0;}" : "+f"(a[i]) : "r"(u[i]));
        asm volatile("{.reg .b16 l,h; mov.b32 {l,h}, %1; fma.rn.f32.f16 %0, h, l, %0;}" : "+f"(a[(i + 1) & 7]) : "r"(u[i]));
        asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(u[(i + 3) & 7]) : "f"(a[i]), "f"(a[(i + 1) & 7]));
      }
    }
  }
  long long t1 = clock64();
  float s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += a[i] + __uint_as_float(u[i]);
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
template <int K>
void run(const char* name, int per_iter, int warps_per_smsp) {
  float* out; long long* cyc;
  cudaMalloc(&out, 4096 * 4); cudaMalloc(&cyc, 64);
  bench<K><<<1, 128 * warps_per_smsp>>>(out, cyc, 0.5f);
  bench<K><<<1, 128 * warps_per_smsp>>>(out, cyc, 0.5f);
  long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
  printf("%-34s warps/SMSP %d: %.2f cycles per warp-instruction per SMSP\n", name, warps_per_smsp,
         (double)c / ((double)ITERS * per_iter * warps_per_smsp));
  cudaFree(out); cudaFree(cyc);
}
int main() {
  for (int w = 1; w <= 4; w *= 4) {
    run<0>("MUFU.EX2", 8, w);
    run<1>("F2FP.F16.F32.PACK_AB", 8, w);
    run<2>("FHFMA (fma.rn.f32.f16)", 8, w);
    run<3>("HADD2.F32 (cvt.f32.f16)", 8, w);
    run<4>("FADD", 8, w);
    run<5>("MUFU + F2FP", 16, w);
    run<6>("MUFU + 7 FADD", 64, w);
    run<7>("split pair: F2FP FHFMA FHFMA F2FP", 32, w);
  }
  return 0;
}
