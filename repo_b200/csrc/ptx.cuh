// Thin inline-PTX wrappers for sm_100a: mbarrier, bulk async copy (TMA engine, 1-D),
// tcgen05 (TMEM alloc, MMA, commit, ld) and the proxy fences that tie them together.
// Only what the RSSM kernels need; every wrapper is one instruction.
#pragma once
#include <cstdint>
#include <cuda_fp16.h>

namespace rb {

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile(
      "{\n\t.reg .pred P;\n\t"
      "elect.sync _|P, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, P;\n\t}\n"
      : "=r"(pred));
  return pred != 0;
}

// ----------------------------------------------------------------------------- mbarrier
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// (A suspendTimeHint of 100 us was measured on the layer machines: the waiting warps no longer re-issue the wait loop, but
// every wake-up got ~900 cycles slower — 133k -> 142k cycles per 128-row step.  The default limit stays.)
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred P;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, P;\n\t}\n"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ uint64_t global_timer_ns() {
  uint64_t t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
// Bounded wait: a protocol bug traps (launch error) after ~2 s instead of hanging the GPU.  The timer is only read every 256
// polls: with the check in every iteration the wait loop was 12 instructions long, and the 16 epilogue warps of the rows
// kernel spent a quarter of their schedulers' issue slots in it (ncu: 37 % of all dynamic instructions) — slots the warps
// that still had work on the same sub-partition, and the MMA issuer, were waiting for.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  const uint64_t t0 = global_timer_ns();
  for (;;) {
#pragma unroll 1
    for (int i = 0; i < 64; ++i) {
      if (mbar_try_wait(bar, parity)) return;
      if (mbar_try_wait(bar, parity)) return;
      if (mbar_try_wait(bar, parity)) return;
      if (mbar_try_wait(bar, parity)) return;
    }
    if (global_timer_ns() - t0 > 2000000000ull) __trap();
  }
}

// ----------------------------------------------------------------------------- bulk copy
// global -> shared::cta, completion counted in bytes on an mbarrier (TMA engine, UBLKCP).
__device__ __forceinline__ void bulk_g2s(uint32_t dst_smem, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst_smem),
      "l"(src), "r"(bytes), "r"(bar)
      : "memory");
}

// generic-proxy smem writes -> visible to the async proxy (tensor core / TMA reads)
__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

// ----------------------------------------------------------------------------- tcgen05
__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
// D[tmem] (+)= A[smem] * B[smem]^T, fp16 inputs, fp32 accumulate. One thread issues.
__device__ __forceinline__ void umma_f16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                         uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive on an mbarrier once all previously issued MMAs of this thread have completed
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() {
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
// 32 lanes x 16 consecutive 32-bit columns: thread i <- TMEM[lane_base + i][col .. col+16)
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float* v) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}

// ----------------------------------------------------------------------------- descriptors
// Shared-memory matrix descriptor, K-major, no swizzle ("interleaved" 8x16B core matrices):
//   element (row, k) lives at  (k/8)*LBO + (row/8)*SBO + (row%8)*16 + (k%8)*2   [fp16]
// bits [0,14) addr>>4 | [16,30) LBO>>4 | [32,46) SBO>>4 | [46,48) version=1 | [61,64) layout=0
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((addr >> 4) & 0x3FFF);
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  return d;
}
// Instruction descriptor for kind::f16: 16-bit A/B (K-major both; format 0 = fp16, 1 = bf16, chosen per
// operand), fp32 D, shape M x N x 16.
__host__ __device__ constexpr uint32_t make_idesc_f16(int M, int N, uint32_t a_fmt = 0, uint32_t b_fmt = 0) {
  return (1u << 4)                                  // D format: f32
         | (a_fmt << 7) | (b_fmt << 10)             // A, B format
         | (0u << 15) | (0u << 16)                  // A, B K-major
         | (static_cast<uint32_t>(N >> 3) << 17)    // N / 8
         | (static_cast<uint32_t>(M >> 4) << 24);   // M / 16
}

// ----------------------------------------------------------------------------- fp16 hi/lo split
// x ~= hi + lo with both in fp16 (saturating); hi*hi + lo*hi + hi*lo reproduces an fp32 product
// to ~2^-22 relative.
__device__ __forceinline__ void split_f16(float x, __half& hi, __half& lo) {
  uint16_t h, l;
  float hf, r;
  asm("cvt.rn.satfinite.f16.f32 %0, %1;" : "=h"(h) : "f"(x));
  asm("cvt.f32.f16 %0, %1;" : "=f"(hf) : "h"(h));
  r = x - hf;
  asm("cvt.rn.satfinite.f16.f32 %0, %1;" : "=h"(l) : "f"(r));
  hi = __ushort_as_half(h);
  lo = __ushort_as_half(l);
}

}  // namespace rb
