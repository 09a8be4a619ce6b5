// Latency of the synchronisation instructions the layer machines use per k-slab group / per hand-off round (B200):
// cycles per back-to-back call from ONE thread, result consumed where there is one.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o sync_lat sync_lat.cu && ./sync_lat
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
#define REP 64
__global__ void k(long long* out) {
  __shared__ __align__(8) uint64_t bars[4];
  __shared__ uint32_t tslot;
  const uint32_t b0 = smem_u32(&bars[0]), b1 = smem_u32(&bars[1]), b2 = smem_u32(&bars[2]);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(b0), "r"(1));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(b1), "r"((1 << 20) - 1));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(b2), "r"((1 << 20) - 1));
    asm volatile("fence.mbarrier_init.release.cluster;");
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(b0) : "memory");   // phase 0 of b0 complete
  }
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tslot)), "r"(32));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    long long t0, t1;
    uint32_t acc = 0;
    // a) try_wait on a completed phase
    t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < REP; ++i) {
      uint32_t ok;
      asm volatile("{.reg .pred P; mbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2; selp.u32 %0, 1, 0, P;}" : "=r"(ok) : "r"(b0), "r"(0) : "memory");
      acc += ok;
      if (!ok) break;
    }
    t1 = clock64(); out[0] = (t1 - t0) / REP;
    // b) arrive
    t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < REP; ++i) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(b1) : "memory");
    t1 = clock64(); out[1] = (t1 - t0) / REP;
    // c) tcgen05.commit, nothing outstanding
    t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < REP; ++i) asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(b2) : "memory");
    t1 = clock64(); out[2] = (t1 - t0) / REP;
    // d) tcgen05 fences
    t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < REP; ++i) asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    t1 = clock64(); out[3] = (t1 - t0) / REP;
    t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < REP; ++i) asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    t1 = clock64(); out[4] = (t1 - t0) / REP;
    // e) proxy fence
    t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < REP; ++i) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    t1 = clock64(); out[5] = (t1 - t0) / REP;
    // f) try_wait (completed) followed by a dependent arrive: the pair an epilogue round costs at minimum
    t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < REP; ++i) {
      uint32_t ok;
      asm volatile("{.reg .pred P; mbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2; selp.u32 %0, 1, 0, P;}" : "=r"(ok) : "r"(b0), "r"(0) : "memory");
      if (ok) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(b1) : "memory");
    }
    t1 = clock64(); out[6] = (t1 - t0) / REP;
    // g) test_wait (non-blocking) on a completed phase
    t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < REP; ++i) {
      uint32_t ok;
      asm volatile("{.reg .pred P; mbarrier.test_wait.parity.shared::cta.b64 P, [%1], %2; selp.u32 %0, 1, 0, P;}" : "=r"(ok) : "r"(b0), "r"(0) : "memory");
      acc += ok;
      if (!ok) break;
    }
    t1 = clock64(); out[7] = (t1 - t0) / REP;
    out[15] = acc;
  }
  if (threadIdx.x < 32) {
    __syncwarp();
    // h) warp-wide: tcgen05.wait::st / ::ld with nothing outstanding, elect.sync
    long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < REP; ++i) asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    long long t1 = clock64(); if (threadIdx.x == 0) out[8] = (t1 - t0) / REP;
    t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < REP; ++i) asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    t1 = clock64(); if (threadIdx.x == 0) out[9] = (t1 - t0) / REP;
    uint32_t p = 0;
    t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < REP; ++i) { uint32_t q; asm volatile("{.reg .pred P; elect.sync _|P, 0xffffffff; selp.u32 %0, 1, 0, P;}" : "=r"(q)); p += q; }
    t1 = clock64(); if (threadIdx.x == 0) { out[10] = (t1 - t0) / REP; out[14] = p; }
    t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < REP; ++i) __syncwarp();
    t1 = clock64(); if (threadIdx.x == 0) out[11] = (t1 - t0) / REP;
  }
  __syncthreads();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tslot), "r"(32));
}
int main() {
  long long* d; cudaMalloc(&d, 16 * 8); cudaMemset(d, 0, 128);
  k<<<1, 128>>>(d); k<<<1, 128>>>(d);
  long long h[16]; cudaMemcpy(h, d, 128, cudaMemcpyDeviceToHost);
  const char* names[] = {"mbarrier.try_wait (phase complete)", "mbarrier.arrive", "tcgen05.commit (nothing outstanding)", "tcgen05.fence::after_thread_sync",
                         "tcgen05.fence::before_thread_sync", "fence.proxy.async.shared::cta", "try_wait + dependent arrive", "mbarrier.test_wait (phase complete)",
                         "tcgen05.wait::st (nothing outstanding)", "tcgen05.wait::ld (nothing outstanding)", "elect.sync", "__syncwarp"};
  for (int i = 0; i < 12; ++i) printf("%-42s %5lld cycles\n", names[i], h[i]);
  printf("err: %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
