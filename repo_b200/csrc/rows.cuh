// Layer machine, "rows on M" variant — the bulk-throughput kernel for observe / imagine.
//
// One CTA owns 128 rows (batch rows = independent RSSM sequences) = the 128 TMEM lanes = the M
// dimension of every tcgen05.mma; a layer's output features are the N dimension (16..256 per
// instruction), so narrow layers (action head N=32, scalar heads N=16) cost what they are worth
// and nothing is padded to 128-feature tiles.  Per row everything is thread-local in the epilogue:
// thread <-> TMEM lane <-> row.
//
// Operand homes:
//   X = [belief | state | action]  fp16 hi/lo in SHARED memory (K-major core matrices) — A operand
//       of the layers that read it (SS-mode MMA);
//   H = hidden activations          fp16 hi/lo packed pairs in TENSOR memory — A operand of the layers that read it
//       (TS-mode MMA).  TMEM is two 256-column regions that swap roles every hidden layer: a layer accumulates into the
//       region that does not hold its H operand and its epilogue rewrites the accumulators IN PLACE as the next H
//       (16 fp32 columns of a feature chunk -> 8 columns of hi pairs + 8 of lo pairs, tcgen05.ld / tcgen05.st);
//   W = weights                     B operand, pre-packed (pack.cuh, PackRowsJob) slabs [hi | lo] of
//       Npad x 16 K, streamed L2 -> smem ring by the TMA engine (1-D bulk copies);
//   accumulators                    the other TMEM region (RStage::regs).
// Overlap of tensor pipe and epilogue (round 2): K-CHAINING.  A hidden layer's epilogue rewrites its accumulators as H in
// ROUNDS of four 16-column chunks (one per epilogue warp of a lane quadrant) and signals each round on a ring of
// mbarriers; the MMA issuer runs the NEXT layer's full-width GEMM (N = 208: tensor-bound — tcgen05.mma is bound by
// fetching the 128 x 16 A tile, ~64 cycles, whenever N < 128, which is what made N-split layers slow) k-slab by k-slab
// right behind those rounds.  Stages that only read X and whose inputs were final earlier (RStage::xback) start while the
// previous stage's epilogue is still running, accumulating into the dead H region.
// x*W = hi*hi + lo*hi + hi*lo (three MMAs, fp32 accumulate) as in vm.cuh.
//
// Warps: 0 = weight loader, 1 = MMA issuer + TMEM owner (warpgroup 0 gives its registers away with
// setmaxnreg), 4..19 = epilogue: FOUR warps per TMEM lane quadrant (`part` 0..3 — all four sit on the SM
// sub-partition that owns the quadrant, so each scheduler has four epilogue warps to hide TMEM / MUFU latency
// with); they split the 16-column chunks of wide layers, the 16-unit groups of a GRU chunk and the 8-state
// groups of the Gaussian heads.  Round 1 ran two warps per quadrant at 208 registers: the epilogue warps were
// busy ~113k of the 160k cycles of a step, each stalled ~65 % of the time.
#pragma once
#include <type_traits>
#include "vm.cuh"

namespace rb {

constexpr int kRowsM = 128;
constexpr int kRowsThreads = 640;   // warpgroup 0: loader, MMA issuer, 2 spare; warpgroups 1-4: epilogue
constexpr int kRowsEpiWarp0 = 4;
constexpr int kRowsEpiThreads = 512;
constexpr int kEpiParts = 4;        // epilogue warps per TMEM lane quadrant
constexpr int kRowsEpiWarps = kRowsEpiThreads / 32;
constexpr int kRoundBars = 8;       // ring of round barriers; the epilogue is never more than 5 rounds ahead of the issuer
constexpr int kRSlotBytes = 28672;     // two 208-wide slabs; the weight stream is not what bounds the kernel (profiling flag 8)
constexpr int kRSlots = 3;
constexpr int kRMaxStages = 32;
constexpr int kRMaxGemms = 56;
constexpr int kBiasCap = 5248;         // floats of shared memory for the biases of the WHOLE program, resident for the launch
constexpr uint32_t kAccCol = 256;     // accumulators live at TMEM columns [256, 512)
constexpr uint32_t kXLBO = kRowsM * 16;  // bytes between k-groups of X (128 rows x 16 B)
constexpr int kRowsMiscBytes = 192 + kEpiParts * 128 * 4;   // 18 mbarrier slots + GruCtx (48 bytes), cross-warp partial sums (floats[parts][128])
// Shared-memory map: everything of fixed size first, so that its addresses are the laundered base plus a constant.
constexpr uint32_t kOffBias = 0;                                     // kBiasCap floats: every stage's bias vector (RStage::bias_off)
constexpr uint32_t kOffBars = kOffBias + kBiasCap * 4;                // 24 x 8 bytes: mbarriers, TMEM base slot
constexpr uint32_t kOffScratch = kOffBars + 192;                      // [kEpiParts][128] floats
constexpr uint32_t kOffRing = (kOffScratch + kEpiParts * 128 * 4 + 127) / 128 * 128;
constexpr uint32_t kOffX = kOffRing + kRSlots * kRSlotBytes;          // X hi plane, then the lo plane (kx16 k-slabs each)

enum RowsEpi : uint8_t {
  R_ACT_H = 0,   // H[:, f] = act(acc + bias (+ addend))            -> TMEM H
  R_ACTION = 1,  // tanh-Normal action sample                        -> X action slot
  R_GRU = 2,     // one chunk of GRU units -> beliefs[t]; last chunk refreshes X belief slot
  R_PRIOR = 3,
  R_POST = 4,
  R_SCALAR = 5,
  R_ACT_DOT = 6,  // last hidden layer of a scalar head fused with its 1-output layer: out = w . act(acc + b) + b0
};
enum RowsFlags : uint8_t { RF_LAST_CHUNK = 16 };  // plus SF_* from vm.cuh
// Belief refresh (R_GRU): every chunk's hh MMAs read the OLD belief from X, so the new units cannot go into X before the last
// chunk's MMAs have completed.  Each earlier chunk therefore writes its new units, already split into fp16 hi/lo and in X's
// shared-memory layout, to a per-SM scratch in global memory (coalesced 16-byte stores, L2-resident), and the last chunk's
// epilogue pulls them into X with two 1-D bulk copies (TMA engine) that run under its own gate math.  Round 1 parked them in
// free TMEM columns instead (re-reading beliefs[t] from global, ~9k exposed cycles per step).
struct RGemm {            // acc[:, acc_col ..+n) (+)= A(128 x 16*ksl) * W(n x 16*ksl)^T
  uint32_t w_off16;       // weight blob offset / 16
  uint16_t slab_bytes16;  // slab bytes / 16  (= n * 4)
  uint16_t n;             // padded output width (multiple of 16)
  uint8_t ksl;            // k16 slabs
  uint8_t a_src;          // 0 = X (smem), 1 = H (tmem)
  uint8_t a_k16;          // first k16 slab inside the source
  uint8_t accumulate;
  uint16_t acc_col;       // column offset inside the accumulator region
  uint16_t init_cols;     // != 0: the GEMM accumulates (accumulate = 1) except into its first init_cols columns, which its very
                          // first MMA overwrites — that MMA is issued as two (N = init_cols fresh, N = n - init_cols accumulating)
};
struct RStage {
  uint8_t gemm_begin, gemm_end;
  uint8_t epi, flags;
  uint8_t act;
  uint8_t regs;        // TMEM regions: bit0 = accumulator region, bit1 = region holding this stage's H operand
  uint16_t nfeat;      // valid output features (R_ACT_H) / units in this chunk (R_GRU)
  uint16_t bias_off;   // float offset into the bias blob
  uint16_t unit0;      // R_GRU: first unit of the chunk
  uint16_t width;      // R_GRU: padded units per chunk (gate stride in the accumulator); heads: padded half width
  uint16_t bias_n;     // floats this stage reads from the bias blob (staged in smem by the epilogue warps)
  uint8_t rounds;      // hand-off rounds this stage's epilogue signals: R_ACT_H = ceil(chunks / 4), every other kind 1
  uint8_t xback;       // rounds (of the stages in between) back from the end of the previous stage to the end of the stage that
                       // last wrote the X columns this stage reads: 0 = the previous stage (the default)
  uint16_t round0;     // rounds signalled by the stages before this one within a step (global round = 1 + t * rounds_per_step + round0 + i)
};

struct RowsParams {
  VmParams v;            // dims, scalars, I/O pointers (stage/gemm tables inside are unused here)
  int n_rstages;
  int kh_cols;           // 8 * kh16 (kept for the host-side size checks; H is interleaved hi/lo per 16-feature chunk)
  uint8_t* scr;          // belief refresh scratch: per SM (slot = %smid) an fp16 [hi plane | lo plane] image of X's belief k-groups
  uint32_t scr_plane;    // bytes per plane (= ceil(D / 8) * kXLBO)
  int scr_slots;         // slots allocated (SMs whose %smid is not below this use the global re-read path)
  int rounds_per_step;   // sum of RStage::rounds over the program
  int gru_tsc_col;       // R_GRU: first of 32 free TMEM columns behind the H operand (transposed beliefs[t] stores), or -1
  int n_bias;            // floats in the bias blob (<= kBiasCap: resident in shared memory)
  RStage stages[kRMaxStages];
  RGemm gemms[kRMaxGemms];
};

// ----------------------------------------------------------------------------- PTX extras
__device__ __forceinline__ void umma_f16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(d_tmem),
      "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t* r) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(r[0]),
               "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}
__device__ __forceinline__ void tmem_st16f(uint32_t taddr, const float* v) {
  const uint32_t* r = reinterpret_cast<const uint32_t*>(v);
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]),
      "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
// 16 lanes x 16 columns in the mma-accumulator fragment layout: thread t gets, for rows t/4 and t/4 + 8 of the 16-lane
// window, columns {2(t%4), 2(t%4)+1} (v[0..1] / v[2..3]) and {8 + 2(t%4), 9 + 2(t%4)} (v[4..5] / v[6..7]) — i.e. four
// neighbouring threads hold 8 consecutive columns of one row, which turns per-row stores into 32-byte segments.
__device__ __forceinline__ void tmem_ld_16x256b_x2(uint32_t taddr, float* v) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile("tcgen05.ld.sync.aligned.16x256b.x2.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr)
               : "memory");
}
// 16 lanes x 8 columns, same fragment layout: v[0..1] = row t/4, columns {2(t%4), 2(t%4)+1}; v[2..3] = row t/4 + 8
__device__ __forceinline__ void tmem_ld_16x256b_x1(uint32_t taddr, float* v) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile("tcgen05.ld.sync.aligned.16x256b.x1.b32 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(taddr)
               : "memory");
}
// 32 lanes x 8 consecutive 32-bit columns
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float* v) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr)
               : "memory");
}
__device__ __forceinline__ void tmem_st8f(uint32_t taddr, const float* v) { tmem_st8(taddr, reinterpret_cast<const uint32_t*>(v)); }
// 4-byte asynchronous global -> shared copy (LDGSTS): the bias staging needs no register round trip, so the issuing
// warp does not stall on the L2 latency
__device__ __forceinline__ void cp_async4(uint32_t dst_smem, const void* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst_smem), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// Shared memory is addressed through 32-bit shared-window addresses derived from ONE base register (the kernel launders it
// through an asm so that it is not rematerialised): generic pointers into dynamic shared memory made the compiler rebuild
// the window base (S2UR SR_CgaCtaId + two ULEAs, a scoreboard stall each time) in front of nearly every access.
__device__ __forceinline__ float4 lds_f4(uint32_t a) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a) : "memory");
  return v;
}
__device__ __forceinline__ uint4 lds_u4(uint32_t a) {
  uint4 v;
  asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a) : "memory");
  return v;
}
__device__ __forceinline__ float lds_f(uint32_t a) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a) : "memory");
  return v;
}
__device__ __forceinline__ uint32_t lds_u32(uint32_t a) {
  uint32_t v;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a) : "memory");
  return v;
}
__device__ __forceinline__ void sts_u4(uint32_t a, uint4 v) {
  asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ void sts_u32(uint32_t a, uint32_t v) { asm volatile("st.shared.u32 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ void sts_f(uint32_t a, float v) { asm volatile("st.shared.f32 [%0], %1;" ::"r"(a), "f"(v) : "memory"); }
__device__ __forceinline__ void sts_u16(uint32_t a, uint16_t v) { asm volatile("st.shared.u16 [%0], %1;" ::"r"(a), "h"(v) : "memory"); }

// two floats -> packed fp16 hi pair and lo pair (element 0 in the low half).  Both halves must be fp16:
// tcgen05 kind::f16 rejects mixed fp16 x bf16 operands (tried: illegal instruction), so a cheap
// bf16-by-truncation lo half is not an option.
__device__ __forceinline__ void split2_f16(float x0, float x1, uint32_t& hi, uint32_t& lo) {
  // (fma.rn.f32.f16 — SASS FHFMA, x - hi in one instruction — was measured: it issues at half rate, no gain over
  // HADD2.F32 + FADD; scripts/ubench/pipes.cu)
  asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(hi) : "f"(x1), "f"(x0));
  const __half2 h = *reinterpret_cast<const __half2*>(&hi);
  const float2 hf = __half22float2(h);
  asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(lo) : "f"(x1 - hf.y), "f"(x0 - hf.x));
}

// element (row, k) of X: byte offset inside the hi (or lo) buffer
__device__ __forceinline__ uint32_t x_off(int row, int k) {
  return (uint32_t)(k >> 3) * kXLBO + (uint32_t)row * 16u + (uint32_t)(k & 7) * 2u;
}
__device__ __forceinline__ void x_put(uint32_t hi, uint32_t lo, int row, int k, float v) {
  __half h, l;
  split_f16(v, h, l);
  const uint32_t o = x_off(row, k);
  sts_u16(hi + o, __half_as_ushort(h));
  sts_u16(lo + o, __half_as_ushort(l));
}
// 8 consecutive k (one k-group) of one row: a single 16-byte store per half
__device__ __forceinline__ void x_put8(uint32_t hi, uint32_t lo, int row, int kgroup, const float* v) {
  uint4 h, l;
  split2_f16(v[0], v[1], h.x, l.x);
  split2_f16(v[2], v[3], h.y, l.y);
  split2_f16(v[4], v[5], h.z, l.z);
  split2_f16(v[6], v[7], h.w, l.w);
  const uint32_t o = (uint32_t)kgroup * kXLBO + (uint32_t)row * 16u;
  sts_u4(hi + o, h);
  sts_u4(lo + o, l);
}

__host__ __device__ inline size_t rows_smem_bytes(int kx16) {
  return (size_t)kOffX + 2 * (size_t)kx16 * 2 * kXLBO;
}

// 16 contiguous floats of one row (guarded tail / alignment handled outside the fast path)
__device__ __forceinline__ void ld_row16(float* dst, const float* base, int n_valid, bool row_ok) {
  if (row_ok && n_valid >= 16 && (reinterpret_cast<uintptr_t>(base) & 15) == 0) {
    const float4* p = reinterpret_cast<const float4*>(base);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float4 v = p[i];
      dst[4 * i] = v.x; dst[4 * i + 1] = v.y; dst[4 * i + 2] = v.z; dst[4 * i + 3] = v.w;
    }
  } else {
#pragma unroll
    for (int i = 0; i < 16; ++i) dst[i] = (row_ok && i < n_valid) ? base[i] : 0.f;
  }
}
// same for rows whose stride only guarantees 8-byte alignment (state rows: S = 30 floats)
__device__ __forceinline__ void ld_row16_v2(float* dst, const float* base, int n_valid, bool row_ok) {
  if (row_ok && n_valid >= 16 && (reinterpret_cast<uintptr_t>(base) & 7) == 0) {
    const float2* p = reinterpret_cast<const float2*>(base);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float2 v = __ldg(p + i);
      dst[2 * i] = v.x; dst[2 * i + 1] = v.y;
    }
  } else if (row_ok && (reinterpret_cast<uintptr_t>(base) & 7) == 0) {
    const float2* p = reinterpret_cast<const float2*>(base);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (2 * i + 1 < n_valid) {
        const float2 v = __ldg(p + i);
        dst[2 * i] = v.x; dst[2 * i + 1] = v.y;
      } else {
        dst[2 * i] = (2 * i < n_valid) ? __ldg(base + 2 * i) : 0.f;
        dst[2 * i + 1] = 0.f;
      }
    }
  } else {
#pragma unroll
    for (int i = 0; i < 16; ++i) dst[i] = (row_ok && i < n_valid) ? __ldg(base + i) : 0.f;
  }
}
// 8 floats of one row whose start is 8-byte aligned (state rows, S even); PLAIN loads (not the read-only path): the KL
// term reads prior_m / prior_sd that this kernel wrote a few stages earlier
template <bool NC>
__device__ __forceinline__ void ld_row8_v2(float* dst, const float* base, int n_valid, bool row_ok) {
  // ONE code path (predicated float2 loads): with an aligned and an unaligned variant the eight values met in local memory.
  // The host routes odd state sizes / unaligned tensors to the vm kernel (api.cu: rows_state_rows_aligned).
  const float2* p = reinterpret_cast<const float2*>(base);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float2 v = make_float2(0.f, 0.f);
    if (row_ok && 2 * i + 1 < n_valid) v = NC ? __ldg(p + i) : p[i];
    dst[2 * i] = v.x; dst[2 * i + 1] = v.y;
  }
}
__device__ __forceinline__ void st_row16_v2(float* dst, const float* v, int n_valid) {
  if ((reinterpret_cast<uintptr_t>(dst) & 7) == 0) {
    float2* p = reinterpret_cast<float2*>(dst);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (2 * i + 1 < n_valid) p[i] = make_float2(v[2 * i], v[2 * i + 1]);
      else if (2 * i < n_valid) dst[2 * i] = v[2 * i];
    }
  } else {
#pragma unroll
    for (int i = 0; i < 16; ++i)
      if (i < n_valid) dst[i] = v[i];
  }
}
// 16 floats every lane reads alike, from the smem bias staging buffer (broadcast LDS.128)
__device__ __forceinline__ void ld_uni16(float* dst, uint32_t base) {
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float4 v = lds_f4(base + 16u * i);
    dst[4 * i] = v.x; dst[4 * i + 1] = v.y; dst[4 * i + 2] = v.z; dst[4 * i + 3] = v.w;
  }
}
// state pair (k even) of one row: one 4-byte store per half
__device__ __forceinline__ void x_put2(uint32_t hi, uint32_t lo, int row, int k, float v0, float v1) {
  uint32_t h, l;
  split2_f16(v0, v1, h, l);
  const uint32_t o = x_off(row, k);
  sts_u32(hi + o, h);
  sts_u32(lo + o, l);
}
__device__ __forceinline__ void st_row16(float* dst, const float* v, int n_valid) {
  if (n_valid >= 16 && (reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
    float4* p = reinterpret_cast<float4*>(dst);
#pragma unroll
    for (int i = 0; i < 4; ++i) p[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
  } else {
#pragma unroll
    for (int i = 0; i < 16; ++i)
      if (i < n_valid) dst[i] = v[i];
  }
}
// branch-free activations (all 16 lanes of a chunk stay independent -> ILP across elements).
// (An FMA-pipe polynomial exp for every other element was tried to unload the XU pipe: it made the
// stage 45% slower — the epilogue is issue/latency-bound, extra instructions cost more than MUFU.)
template <int ACT>
__device__ __forceinline__ float act_bf(float x) {
  if (ACT == ACT_ELU) {
    const float m = ex2_f(x * 1.4426950408889634f) - 1.f;  // garbage (inf) for large x is selected away
    return x > 0.f ? x : m;
  }
  return fmaxf(x, 0.f);
}

// H = act(acc + bias (+ addend)) for ONE 16-column chunk of this thread's row.
// Pad columns need no guard: their weight rows and bias are zero, so they come out as act(0) = 0.
// DOT: instead of storing H, reduce it against a weight vector (the scalar head's last layer).
// H is written IN PLACE: accumulator columns [16 ch, 16 ch + 16) of this thread's lane become the packed fp16 hi
// pairs (8 columns) followed by the lo pairs (8 columns) of the same 16 features, so the region that held the
// accumulators is the next layer's A operand and the region that held this layer's operand is free for its accumulators.
// the tiled addend of one chunk (four planes of 128 rows x 4 floats, 512 floats apart) into registers; zeros where the row or
// the feature quad does not exist (n_valid is a multiple of 4: Hd % 4 == 0 is what selects the tiled layout)
__device__ __forceinline__ void rows_addend_prefetch(float4* pf, const float* ad, int n_valid, bool row_ok) {
#pragma unroll
  for (int j = 0; j < 4; ++j)
    pf[j] = (row_ok && 4 * j < n_valid) ? __ldg(reinterpret_cast<const float4*>(ad + 512 * j)) : make_float4(0.f, 0.f, 0.f, 0.f);
}
// `tcol`, `bias`, `wdot`, `adrow` address THIS chunk (the callers step them from chunk to chunk in registers: recomputing
// them from the chunk index cost ~50 instructions per chunk, a quarter of the issue-bound epilogue).
// ADDEND == 2: the addend comes in the TILED layout — per (time step, 128-row tile, 16-feature chunk) four planes of
// 128 rows x 4 floats — so each of the four loads is 32 lanes x 16 contiguous bytes.  Row-major (ADDEND == 1) puts the 32
// lanes of a load on 32 different cache lines (rows are Hd floats apart), and at 13 chunks x 16 warps that made the posterior
// layer's epilogue 15.3k cycles against 3.5k for a layer without an addend.
template <int ACT, bool DOT, int ADDEND>
__device__ __forceinline__ float rows_act_chunk(uint32_t bias, uint32_t wdot, const float* adrow, uint32_t tcol,
                                                int n_valid, bool row_ok, const float4* pf = nullptr) {
  float v[16], bz[16];
  float dot = 0.f;
  tmem_ld16(tcol, v);
  ld_uni16(bz, bias);
  if (ADDEND == 1) {
    float ad[16];
    ld_row16(ad, adrow, n_valid, row_ok);
#pragma unroll
    for (int i = 0; i < 16; ++i) bz[i] += ad[i];
  } else if (ADDEND == 2) {   // requested one round ahead by the caller (rows_addend_prefetch): already in registers
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      bz[4 * j] += pf[j].x; bz[4 * j + 1] += pf[j].y; bz[4 * j + 2] += pf[j].z; bz[4 * j + 3] += pf[j].w;
    }
  }
  tmem_ld_wait();
  if (DOT) {
    float w[16];
    ld_uni16(w, wdot);
#pragma unroll
    for (int i = 0; i < 16; ++i) dot = fmaf(w[i], act_bf<ACT>(v[i] + bz[i]), dot);
  } else {
    uint32_t hi[8], lo[8];
#pragma unroll
    for (int i = 0; i < 16; i += 2)
      split2_f16(act_bf<ACT>(v[i] + bz[i]), act_bf<ACT>(v[i + 1] + bz[i + 1]), hi[i >> 1], lo[i >> 1]);
    tmem_st8(tcol, hi);
    tmem_st8(tcol + 8, lo);
  }
  return dot;
}

// (A leading quarter-chunk round — chunk 0 split over the four warps of a quadrant, so that the next layer's first k-slab is
// ready ~300 instead of ~850 cycles into the epilogue — was measured as well: the layer period went from 6.45k to 6.95k
// cycles; the MMAs of the last, now four-chunk, round trail the epilogue instead.)
// (Splitting the 13th chunk of a 208-wide layer four ways, so that no warp of a quadrant takes 4 chunks against 3, was
// measured: slower, 6.5k against 5.5k cycles per layer — the four warps share one scheduler, which is what limits them, and
// the quarter pieces add instructions plus a quadrant barrier for the in-place rewrite.)

// In-kernel clock64 stamps (scripts/stage_clock.py) exist only in the profiling build (-DRB_STAGE_CLOCK): in the product
// build they would cost registers in the epilogue warps, which run at the spill threshold.
#ifdef RB_STAGE_CLOCK
#define RB_STAMP(k) do { if (V.dbg_clock && blockIdx.x == 0 && et == 0 && t == 5) V.dbg_clock[700 + s * 8 + (k)] = clock64(); } while (0)
#define RB_STAMP_X(k) do { if (V.dbg_clock && blockIdx.x == 0 && et == 0 && t == 5) V.dbg_clock[600 + s * 2 + (k)] = clock64(); } while (0)
// (flag 64: also the first lane of EVERY epilogue warp at step 5 -> [1200 + warp * 64 + stage * 2 + {0, 1}]: which warp is last?)
#define RB_STAGE_BEGIN() do { if (V.dbg_clock && blockIdx.x == 0 && et == 0) V.dbg_clock[(t * P.n_rstages + s) * 2] = clock64(); \
    if ((V.dbg_flags & 64) && V.dbg_clock && blockIdx.x == 0 && lane == 0 && t == 5) V.dbg_clock[1200 + (et >> 5) * 64 + s * 2] = clock64(); } while (0)
#define RB_STAGE_END() do { if (V.dbg_clock && blockIdx.x == 0 && et == 0) V.dbg_clock[(t * P.n_rstages + s) * 2 + 1] = clock64(); \
    if ((V.dbg_flags & 64) && V.dbg_clock && blockIdx.x == 0 && lane == 0 && t == 5) V.dbg_clock[1200 + (et >> 5) * 64 + s * 2 + 1] = clock64(); } while (0)
#else
#define RB_STAMP(k) do { } while (0)
#define RB_STAMP_X(k) do { } while (0)
#define RB_STAGE_BEGIN() do { } while (0)
#define RB_STAGE_END() do { } while (0)
#endif

// One chunk of GRU units (R_GRU).  accumulator: i_n at [0,W), r at [W,2W), z at [2W,3W), h_n at [3W,4W): W_hh . b goes into
// [W,4W) as ONE N = 3W GEMM over X (it only reads the old belief, so the issuer runs it ahead of the layer that produces the
// other operand), W_ih . h accumulates into [0,3W).  Bias vectors stay in the order r, z, i_n, h_n.
// The stage is handed back as soon as the accumulators have left tensor memory, so the next chunk's MMAs run under the
// gate math.  Holding all four gate accumulators (64 registers per thread) across that point spilled into every stage
// of the kernel (hidden-layer epilogues 4.9k -> 10k cycles), so r * (W_hn b + b_hn) is folded before the hand-off:
// 48 registers (t, z_pre, i_n) cross it.
__device__ __forceinline__ void rows_gru_stage(uint32_t x_hi, uint32_t x_lo, uint8_t* scr, uint32_t scr_plane, uint32_t bar_xb,
                                               int D, uint32_t tacc, uint32_t bias, float* bel_row,
                                               uint32_t bar_handoff, uint32_t xb_parity, int r, int part, bool row_ok, int flags,
                                               int W, int u0, int nu, uint32_t tsc, int rows_left, long long* dbg = nullptr, int dbg_skip = 0) {
#ifdef RB_STAGE_CLOCK
#define RB_GRU_STAMP(k) do { if (dbg) dbg[k] = clock64(); } while (0)
#else
#define RB_GRU_STAMP(k) do { } while (0)
#endif
  RB_GRU_STAMP(0);
  const bool last_chunk = (flags & RF_LAST_CHUNK) != 0;
  const int c = part * 16;         // this warp's 16-unit group of the chunk
  const bool mine = c < nu;        // warp-uniform
  const bool refresh = last_chunk && u0 > 0 && scr != nullptr;
  if (refresh && threadIdx.x == 32 * kRowsEpiWarp0) {
    // Every hh MMA has completed (this stage's accumulator barrier) and every earlier chunk's new units are in the scratch
    // (their writers fenced towards the async proxy and have passed the epilogue barrier at this stage's start): pull belief
    // k-groups [0, u0/8) into X while this chunk's gate math runs.
    const uint32_t bytes = (uint32_t)(u0 >> 3) * kXLBO;
    mbar_arrive_expect_tx(bar_xb, 2u * bytes);
    bulk_g2s(x_hi, scr, bytes, bar_xb);
    bulk_g2s(x_lo, scr + scr_plane, bytes, bar_xb);
  }
  // Two 16-register arrays cross the hand-off: the candidate's pre-activation i_n + b_in + r * (W_hn b + b_hn), folded as soon
  // as its terms are out of tensor memory, and the z pre-activation (three arrays spilled: 16 L2 round trips per chunk).
  float vt[16], vz[16];
  if (mine) {
    float vh[16];
    tmem_ld16(tacc + W + c, vt);        // r
    tmem_ld16(tacc + 3 * W + c, vh);    // h_n
    tmem_ld_wait();
    float vi[16];
    tmem_ld16(tacc + 2 * W + c, vz);    // in flight under the r-gate math
    tmem_ld16(tacc + c, vi);
    {
      float bb[16];
      ld_uni16(bb, bias + 4u * c);
#pragma unroll
      for (int i = 0; i < 16; ++i) vt[i] = sigmoid_f(vt[i] + bb[i]);
      ld_uni16(bb, bias + 4u * (3 * W + c));
#pragma unroll
      for (int i = 0; i < 16; ++i) vt[i] = vt[i] * (vh[i] + bb[i]);
      ld_uni16(bb, bias + 4u * (2 * W + c));
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 16; ++i) vt[i] += vi[i] + bb[i];
    }
  }
  // The accumulators have left tensor memory and X is untouched by this chunk: hand the stage back now, so that the
  // next chunk's MMAs run under the rest of the gate math and the stores.
  if (!last_chunk) {
    tc_fence_before();
    __syncwarp();
    if ((threadIdx.x & 31) == 0) mbar_arrive(bar_handoff);
  }
  RB_GRU_STAMP(1);
#ifdef RB_STAGE_CLOCK
  if ((dbg_skip & 1) && !last_chunk) return;   // profiling only (flag 32): no gate math under the next chunk's MMAs (results are garbage)
#endif
  // Rest of the gate math in two halves of 8 units: new units go to beliefs[t] (fp32) and, split into fp16 hi/lo, either
  // to the scratch image of X (earlier chunks) or straight into X (last chunk: every chunk's MMAs are done, the belief
  // slot may be overwritten).
  const bool to_x = last_chunk && (refresh || u0 == 0);
  if (mine) {
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      float bn[8];
      const int kg = (u0 + c) / 8 + j;
      {
        float g[8];
        const uint32_t pb = bias + 4u * (c + 8 * j);
#pragma unroll
        for (int i = 0; i < 8; ++i) bn[i] = tanh_f(vt[8 * j + i]);   // candidate n
        const float4 b0 = lds_f4(pb + 4u * W), b1 = lds_f4(pb + 4u * W + 16u);
        const float bzz[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
        for (int i = 0; i < 8; ++i) g[i] = sigmoid_f(vz[8 * j + i] + bzz[i]);                 // z
        // b_prev = hi + lo from the belief slot of X (exact to 2^-22 relative).  The refresh overwrites k-groups below
        // u0/8 only in the last chunk, whose own units lie above them.
        if (kg * 8 < D) {
          const uint4 h4 = lds_u4(x_hi + (uint32_t)kg * kXLBO + (uint32_t)r * 16u);
          const uint4 l4 = lds_u4(x_lo + (uint32_t)kg * kXLBO + (uint32_t)r * 16u);
          const uint32_t hw[4] = {h4.x, h4.y, h4.z, h4.w}, lw[4] = {l4.x, l4.y, l4.z, l4.w};
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float2 hf = __half22float2(*reinterpret_cast<const __half2*>(&hw[e]));
            const float2 lf = __half22float2(*reinterpret_cast<const __half2*>(&lw[e]));
            bn[2 * e] = bn[2 * e] + g[2 * e] * ((hf.x + lf.x) - bn[2 * e]);               // (1-z)*n + z*b_prev
            bn[2 * e + 1] = bn[2 * e + 1] + g[2 * e + 1] * ((hf.y + lf.y) - bn[2 * e + 1]);
          }
        } else {
#pragma unroll
          for (int e = 0; e < 8; ++e) bn[e] = bn[e] - g[e] * bn[e];
        }
      }
      RB_GRU_STAMP(2 + 3 * j);
      const int u = u0 + c + 8 * j;       // first unit of this half
      const int nval = min(8, u0 + nu - u);
#ifdef RB_STAGE_CLOCK
      if ((dbg_skip & 2) && !last_chunk) continue;   // profiling only (flag 512): gate math but no global stores
#endif
      if (tsc != 0u && nval == 8) {
        // beliefs[t] rows are D floats apart: a store with lane = row touches 32 cache lines per instruction, and the L1
        // pipeline those stores occupy is the one the tensor core fetches its shared-memory operands through (profiling
        // flag 512: without the stores a chunk's MMAs take 10.9k instead of 13k cycles, its gate math 4.5k instead of 7.7k).
        // So the 32 x 8 block goes through 8 free TMEM columns (lane = row in, 16x256b fragments out): four neighbouring
        // threads then hold 8 consecutive floats of one row — 8 rows x 32 bytes per store instruction.
        tmem_st8f(tsc, bn);
        tmem_st_wait();
        const int ln = threadIdx.x & 31, cq = 2 * (ln & 3);
        float tq[2][4];
        tmem_ld_16x256b_x1(tsc, tq[0]);
        tmem_ld_16x256b_x1(tsc + (16u << 16), tq[1]);
        tmem_ld_wait();
#pragma unroll
        for (int hb = 0; hb < 2; ++hb) {
#pragma unroll
          for (int rb = 0; rb < 2; ++rb) {
            const int rr = 16 * hb + 8 * rb + (ln >> 2);   // row within this warp's 32
            if (rr < rows_left)
              *reinterpret_cast<float2*>(bel_row + (ptrdiff_t)(rr - ln) * D + u + cq) = make_float2(tq[hb][2 * rb], tq[hb][2 * rb + 1]);
          }
        }
      } else if (row_ok && nval > 0) {
        float* dst = bel_row + u;
        if (nval == 8 && (reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
          reinterpret_cast<float4*>(dst)[0] = make_float4(bn[0], bn[1], bn[2], bn[3]);
          reinterpret_cast<float4*>(dst)[1] = make_float4(bn[4], bn[5], bn[6], bn[7]);
        } else {
#pragma unroll
          for (int i = 0; i < 8; ++i)
            if (i < nval) dst[i] = bn[i];
        }
      }
      RB_GRU_STAMP(3 + 3 * j);
      if (to_x) {
        if (u + 8 <= D) x_put8(x_hi, x_lo, r, u >> 3, bn);
        else
          for (int i = 0; u + i < D && i < 8; ++i) x_put(x_hi, x_lo, r, u + i, bn[i]);
      } else if (!last_chunk && scr != nullptr) {
        // scratch image of X: a warp writes 512 contiguous bytes per store
        uint4 h, l;
        split2_f16(bn[0], bn[1], h.x, l.x);
        split2_f16(bn[2], bn[3], h.y, l.y);
        split2_f16(bn[4], bn[5], h.z, l.z);
        split2_f16(bn[6], bn[7], h.w, l.w);
        const uint32_t o = (uint32_t)kg * kXLBO + (uint32_t)r * 16u;
        *reinterpret_cast<uint4*>(scr + o) = h;
        *reinterpret_cast<uint4*>(scr + scr_plane + o) = l;
      }
      RB_GRU_STAMP(4 + 3 * j);
    }
    if (!last_chunk && scr != nullptr)
      asm volatile("fence.proxy.async.global;" ::: "memory");   // generic-proxy global writes -> visible to the bulk copy
  }
  RB_GRU_STAMP(8);
  if (!last_chunk) return;
  // ---- last chunk ----
  if (to_x) {
    if (refresh) mbar_wait(bar_xb, xb_parity);   // the earlier chunks' units have landed (async proxy, like the MMAs' reads)
  } else {
    // no scratch slot for this SM: a row's units were written by all four warps of the quadrant, so sync the epilogue
    // warps, then re-read beliefs[t]
    __threadfence_block();
    asm volatile("bar.sync 1, 512;" ::: "memory");
    const int nkg = D >> 3;  // full k-groups
    for (int kg = part; kg < nkg; kg += kEpiParts) {
      if (row_ok) {
        float v[8];
        const float4 lo4 = *reinterpret_cast<const float4*>(bel_row + kg * 8);
        const float4 hi4 = *reinterpret_cast<const float4*>(bel_row + kg * 8 + 4);
        v[0] = lo4.x; v[1] = lo4.y; v[2] = lo4.z; v[3] = lo4.w;
        v[4] = hi4.x; v[5] = hi4.y; v[6] = hi4.z; v[7] = hi4.w;
        x_put8(x_hi, x_lo, r, kg, v);
      }
    }
    if (part == 0 && row_ok)
      for (int k = nkg * 8; k < D; ++k) x_put(x_hi, x_lo, r, k, bel_row[k]);
  }
}

// PROG specialises the epilogue for the program family, so that the register allocation of one family does not carry the
// other's stage kinds (ptxas spills a value everywhere once ANY path is short of registers; the posterior layer's per-row
// addend alone is 16 registers): 1 = imagine (no addend, no posterior head), 2 = observe (no actor / scalar-head stages),
// 0 = everything (one-step cell programs, conditional variants).
template <int PROG>
__global__ void __launch_bounds__(kRowsThreads, 1) rssm_rows_kernel(const __grid_constant__ RowsParams P) {
  constexpr bool kHasAddend = PROG != 1, kHasPost = PROG != 1, kHasActor = PROG != 2;
  extern __shared__ __align__(1024) uint8_t smem[];
  const VmParams& V = P.v;
  uint32_t sb;   // shared-window address of the dynamic shared memory; volatile asm = never rematerialised from the symbol
  asm volatile("mov.u32 %0, %1;" : "=r"(sb) : "r"(smem_u32(smem)));
  const uint32_t ring_a = sb + kOffRing;
  const uint32_t x_hi = sb + kOffX;
  const uint32_t x_bytes = (uint32_t)V.kx16 * 2u * kXLBO;
  const uint32_t x_lo = x_hi + x_bytes;
  const uint32_t bars = sb + kOffBars;                           // 8-byte slots
  const uint32_t tmem_slot = bars + 8 * 15;
  const uint32_t scratch = sb + kOffScratch;                     // [kEpiParts][128] cross-warp partial sums
  const uint32_t bias_s = sb + kOffBias;                         // 2 x kBiasStage floats

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row0 = blockIdx.x * kRowsM;
  const uint32_t bar_full = bars, bar_empty = bars + 8 * kRSlots;
  const uint32_t bar_acc = bars + 8 * (2 * kRSlots);         // a stage's accumulators are complete
  const uint32_t bar_round = bars + 8 * (2 * kRSlots + 1);   // kRoundBars hand-off barriers: global round g -> slot g % 8
  const uint32_t bar_xb = bars + 8 * 16;                     // belief refresh bulk copies have landed

  if (threadIdx.x == 0) {
    for (int i = 0; i < kRSlots; ++i) {
      mbar_init(bar_full + 8 * i, 1);
      mbar_init(bar_empty + 8 * i, 1);
    }
    mbar_init(bar_acc, 1);
    for (int i = 0; i < kRoundBars; ++i) mbar_init(bar_round + 8 * i, kRowsEpiWarps);   // one arrival per epilogue warp
    mbar_init(bar_xb, 1);
    mbar_fence_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = lds_u32(tmem_slot);

  // Each role changes its register budget INSIDE its own branch (ptxas sizes a region by the setmaxnreg that dominates
  // it).  640 threads launch with 96 registers each; warpgroup 0 keeps 32, the four epilogue warpgroups get 112.
  // (576 threads x 112 registers without a hand-over does not launch: registers are granted per 4-warp group.  Stage
  // bodies as non-inlined functions are sized for the 96-register launch bound, not for 112 — so everything is inlined
  // and each stage is written to stay inside 112: see the R_GRU case.)
  if (warp == 0) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 32;");
    // ================================ weight loader ================================
    uint32_t slot = 0, phase = 0;
    for (int t = 0; t < V.n_steps; ++t) {
      for (int s = 0; s < P.n_rstages; ++s) {
        const int g0 = P.stages[s].gemm_begin, g1 = P.stages[s].gemm_end;
        for (int g = g0; g < g1; ++g) {
          const RGemm gm = P.gemms[g];
          const uint32_t slab_bytes = (uint32_t)gm.slab_bytes16 * 16u;
          const int per_slot = kRSlotBytes / slab_bytes;
          const uint8_t* src = V.wblob + (size_t)gm.w_off16 * 16u;
          for (int c0 = 0; c0 < gm.ksl; c0 += per_slot) {
            const int nsl = min(per_slot, (int)gm.ksl - c0);
            mbar_wait(bar_empty + 8 * slot, phase ^ 1);
            if (elect_one()) {
              const uint32_t bytes = (uint32_t)nsl * slab_bytes;
#ifdef RB_STAGE_CLOCK
              if (V.dbg_flags & 8) {   // profiling only: no weight traffic (results are garbage)
                mbar_arrive(bar_full + 8 * slot);
              } else
#endif
              {
                mbar_arrive_expect_tx(bar_full + 8 * slot, bytes);
                bulk_g2s(ring_a + slot * kRSlotBytes, src + (size_t)c0 * slab_bytes, bytes, bar_full + 8 * slot);
              }
            }
            __syncwarp();
            if (++slot == kRSlots) { slot = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 32;");
    // ================================ MMA issuer ================================
    // Hand-off rounds are numbered globally (round 0 = the initial staging of X; then every stage's rounds in program
    // order) and complete in order; `consumed` = how many this warp has observed.  Waits are lazy but never skip a round.
    // (An mbarrier wait answers ~160 cycles after it is issued even when the phase completed long ago — scripts/ubench/
    // sync_lat.cu — and a k-slab group needs up to two: ~1.7k cycles per 13-slab layer.  Firing non-blocking test_waits for
    // the NEXT round / ring slot ahead of a group's MMAs was measured: slower, 4.37 against 4.24 ms — the barrier unit
    // serialises them, so they delay the MMAs instead of hiding behind them.  A scout warp doing all the waiting and
    // publishing "k-slab groups ready" as a shared-memory word the issuer polls with ld.acquire (~30 cycles) was built and
    // measured too: correct, 138.6k against 133.8k cycles per step — its and the issuer's spinning cost the epilogue warps
    // on their schedulers more than the shorter waits gave back.)
    uint32_t slot = 0, phase = 0, consumed = 0, next_round = 1;
    auto wait_round = [&](uint32_t idx) {
      while (consumed <= idx) {
        mbar_wait(bar_round + 8 * (consumed & (kRoundBars - 1)), (consumed / kRoundBars) & 1);
        ++consumed;
      }
      tc_fence_after();
    };
    const uint64_t x_desc = make_smem_desc(x_hi, kXLBO, 128);
    const uint64_t x_lo_delta = x_bytes >> 4;
    constexpr uint64_t kX_slab = (2u * kXLBO) >> 4;
    for (int t = 0; t < V.n_steps; ++t) {
      for (int s = 0; s < P.n_rstages; ++s) {
        const int g0 = P.stages[s].gemm_begin, g1 = P.stages[s].gemm_end;
        const uint32_t tacc = tmem_base + ((P.stages[s].regs & 1) ? kAccCol : 0u);
        const uint32_t th = tmem_base + ((P.stages[s].regs & 2) ? kAccCol : 0u);
        const uint32_t my_first = next_round;
        next_round += P.stages[s].rounds;
        const uint32_t prev_rounds = (t == 0 && s == 0) ? 1u : (uint32_t)P.stages[s ? s - 1 : P.n_rstages - 1].rounds;
        const uint32_t prev_first = my_first - prev_rounds;
        // X-reading GEMMs: the stage that last wrote their columns must have finished (default: the previous stage)
        const uint32_t xb = P.stages[s].xback;
        const uint32_t need_x = (my_first >= 1u + xb) ? my_first - 1u - xb : 0u;
        for (int g = g0; g < g1; ++g) {
          const RGemm gm = P.gemms[g];
          const uint32_t slab_bytes = (uint32_t)gm.slab_bytes16 * 16u;
          const int per_slot = kRSlotBytes / slab_bytes;
          const uint32_t idesc = make_idesc_f16(128, gm.n);
          const uint32_t d = tacc + gm.acc_col;
          const uint32_t w_lbo = (uint32_t)gm.n * 16u;  // bytes between the two k-groups of a weight slab
          const uint32_t w_lo = w_lbo * 2u;             // lo half follows the hi half
          uint32_t acc = gm.accumulate;
          uint32_t kk = gm.a_k16;
          if (gm.a_src == 0) wait_round(need_x);
          for (int c0 = 0; c0 < gm.ksl; c0 += per_slot) {
            const int nsl = min(per_slot, (int)gm.ksl - c0);
            // H-reading GEMMs trail the producing epilogue: k-slab k = features [16k, 16k+16) = chunk k = round k / 4
            if (gm.a_src == 1) wait_round(prev_first + min((kk + (uint32_t)nsl - 1u) >> 2, prev_rounds - 1u));
#ifdef RB_STAGE_CLOCK
            if ((V.dbg_flags & 64) && V.dbg_clock && blockIdx.x == 0 && t == 5 && lane == 0) V.dbg_clock[900 + s * 4 + 1] = clock64();   // inputs of this group ready (last one stays)
#endif
            mbar_wait(bar_full + 8 * slot, phase);
            tc_fence_after();
#ifdef RB_STAGE_CLOCK
            if ((V.dbg_flags & 64) && V.dbg_clock && blockIdx.x == 0 && t == 5 && lane == 0) V.dbg_clock[900 + s * 4 + 2] = clock64();   // its weights have landed
#endif
            if (elect_one()) {
              uint32_t wa = ring_a + slot * kRSlotBytes;
              int nsl_issue = nsl;
#ifdef RB_STAGE_CLOCK
              if (V.dbg_flags & 16) nsl_issue = 0;   // profiling only: no MMAs
#endif
              // Descriptors are built once per slot and stepped by adding to their address fields (14 bits of address / 16;
              // shared memory ends below 2^18, so nothing carries into the stride fields): the MMA issue rate is set by this
              // one thread's instruction chain — it shares its scheduler with four epilogue warps — so the per-MMA work is
              // kept to the instruction itself plus two additions per k-slab.
              uint64_t b_hi = make_smem_desc(wa, w_lbo, 128);
              const uint64_t b_lo_delta = w_lo >> 4, b_step = slab_bytes >> 4;
              if (gm.a_src == 0) {
                uint64_t a_hi = x_desc + (uint64_t)kk * kX_slab;
                for (int j = 0; j < nsl_issue; ++j) {
                  umma_f16(d, a_hi, b_hi, idesc, (j == 0) ? acc : 1u);
                  umma_f16(d, a_hi + x_lo_delta, b_hi, idesc, 1u);
                  umma_f16(d, a_hi, b_hi + b_lo_delta, idesc, 1u);
                  a_hi += kX_slab;
                  b_hi += b_step;
                }
              } else {
                uint32_t a_hi = th + kk * 16u;   // per k-slab: 8 hi columns, 8 lo columns
                for (int j = 0; j < nsl_issue; ++j) {
                  if (j == 0 && c0 == 0 && gm.init_cols) {
                    const uint32_t ic = gm.init_cols;
                    umma_f16_ts(d, a_hi, b_hi, make_idesc_f16(128, ic), 0u);
                    umma_f16_ts(d + ic, a_hi, b_hi + ic, make_idesc_f16(128, gm.n - ic), 1u);   // B rows are 16 bytes apart: + ic in the descriptor's address field
                  } else
                    umma_f16_ts(d, a_hi, b_hi, idesc, (j == 0) ? acc : 1u);
                  umma_f16_ts(d, a_hi + 8u, b_hi, idesc, 1u);
                  umma_f16_ts(d, a_hi, b_hi + b_lo_delta, idesc, 1u);
#ifdef RB_STAGE_CLOCK
                  if ((V.dbg_flags & 64) && V.dbg_clock && blockIdx.x == 0 && t == 5 && s == 1) V.dbg_clock[1100 + kk + j] = clock64();   // A2: slab issued
#endif
                  a_hi += 16u;
                  b_hi += b_step;
                }
              }
              umma_commit(bar_empty + 8 * slot);
            }
            __syncwarp();
            acc = 1u;
            kk += nsl;
            if (++slot == kRSlots) { slot = 0; phase ^= 1; }
          }
        }
        if (elect_one()) umma_commit(bar_acc);
        __syncwarp();
#ifdef RB_STAGE_CLOCK
        if ((V.dbg_flags & 64) && V.dbg_clock && blockIdx.x == 0 && t == 5 && lane == 0) V.dbg_clock[900 + s * 4 + 3] = clock64();       // stage committed
#endif
      }
    }
  } else if (warp < kRowsEpiWarp0) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 32;");  // spare warps of warpgroup 0: the dec is warpgroup-wide
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 112;");
    // ================================ epilogue warps ================================
    const int et = threadIdx.x - 32 * kRowsEpiWarp0;      // 0..511
    const int q = warp & 3;                // TMEM lane quadrant (= the SM sub-partition this warp runs on)
    const int part = (warp - kRowsEpiWarp0) >> 2;      // which of the four warps of this quadrant
    const int r = q * 32 + lane;           // row within the tile = TMEM lane
    const int row = row0 + r;
    const int N = V.N, D = V.D, S = V.S, A = V.A;
    const bool row_ok = row < N;
    const uint32_t tl = tmem_base + ((uint32_t)(q * 32) << 16);
    auto epi_sync = [] { asm volatile("bar.sync 1, 512;" ::: "memory"); };
    uint32_t smid;
    asm("mov.u32 %0, %%smid;" : "=r"(smid));
    uint8_t* const scr = (P.scr && (int)smid < P.scr_slots) ? P.scr + (size_t)smid * 2u * P.scr_plane : nullptr;

    // ---- init: zero X, then stage [belief | state*nonterm[0] | action[0]] ----
    {
      const uint32_t words = (2 * x_bytes) / 16;
      for (uint32_t i = et; i < words; i += kRowsEpiThreads) sts_u4(x_hi + 16u * i, make_uint4(0, 0, 0, 0));
      epi_sync();
      if (row_ok) {
        if (V.init_belief) {
          const float* b = V.init_belief + (size_t)row * D;
          for (int kg = part; kg * 8 < D; kg += kEpiParts) {
            float v[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] = (kg * 8 + i < D) ? b[kg * 8 + i] : 0.f;
            if (kg * 8 + 8 <= D) x_put8(x_hi, x_lo, r, kg, v);
            else
              for (int i = 0; kg * 8 + i < D; ++i) x_put(x_hi, x_lo, r, kg * 8 + i, v[i]);
          }
        }
        if (part == 0 && V.init_state) {
          const float m = V.nonterm ? V.nonterm[row] : 1.f;
          for (int j = 0; j < S; ++j) x_put(x_hi, x_lo, r, D + j, V.init_state[(size_t)row * S + j] * m);
        }
        if (part == 1 && V.actions_in)
          for (int j = 0; j < A; ++j) x_put(x_hi, x_lo, r, D + S + j, V.actions_in[(size_t)row * A + j]);
      }
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_round);   // global round 0
    }

    // The biases of every stage stay in shared memory for the whole launch (RStage::bias_off indexes them).  Round 1 staged
    // each stage's vector through a double buffer while the previous stage ran: an LDGSTS per thread, a wait for the L2 round
    // trip and a 512-thread barrier in front of EVERY stage — ~700 cycles of the ~1.6k that separate two dependent layers.
    for (int i = et; i < P.n_bias; i += kRowsEpiThreads) sts_f(bias_s + 4u * i, __ldg(V.bias + i));
    epi_sync();
    uint32_t par = 0;   // parity of the stage counter = phase of the accumulator barrier
    for (int t = 0; t < V.n_steps; ++t) {
      const size_t trow = (size_t)t * N;
      const bool has_next = (t + 1) < V.n_steps;
      for (int s = 0; s < P.n_rstages; ++s) {
        const RStage& st = P.stages[s];
        const uint32_t bias = bias_s + 4u * st.bias_off;
        const uint32_t tacc = tl + ((st.regs & 1) ? kAccCol : 0u);
        RB_STAMP(0);
        // Per-row noise of this stage (this warp's 8 columns), requested before the wait for the accumulators so that the L2
        // round trip runs under the stage's MMAs.  Declared per stage: a value carried around the stage loop would hold 9
        // registers in every other stage's epilogue too.
        float pre[8];
        float pre_nt = 1.f;
        if (st.epi == R_PRIOR || st.epi == R_POST) {
          const int c = part * 8;   // the four warps of a quadrant take one 8-state group each
          const float* eps = ((kHasPost && st.epi == R_POST) ? V.eps_post : V.eps_prior) + (trow + row) * S + c;
          ld_row8_v2<true>(pre, eps, S - c, row_ok && c < S);
          if ((st.flags & SF_WRITES_STATE) && V.nonterm && has_next && row_ok) pre_nt = __ldg(V.nonterm + trow + N + row);
        } else if (kHasActor && st.epi == R_ACTION) {
          const int c = part * 8;
          const float* eps = V.eps_action + (trow + row) * A + c;
#pragma unroll
          for (int i = 0; i < 8; ++i) pre[i] = (row_ok && c + i < A) ? __ldg(eps + i) : 0.f;
        }
        // The posterior layer's per-row addend (tiled layout) comes from HBM — ~1.5k cycles away: the first chunk's is
        // requested here, under the layer's MMAs, each later chunk's one round ahead.
        // (Holding a chunk's addend in registers one round ahead — 16 registers across the layer — was measured: the layer's
        // epilogue 11.8k -> 8.6k cycles, but the spills it caused in this specialisation cost every other stage more.)  So the
        // tile's addend block (contiguous: chunks x 8 KB) is only pulled into L2 here, one 128-byte line per thread and pass.
        if constexpr (kHasAddend) {
          if (st.epi == R_ACT_H && (st.flags & SF_ADDEND) && V.addend_tiled) {
            const char* blk = reinterpret_cast<const char*>(V.addend + ((size_t)t * gridDim.x + blockIdx.x) * (size_t)((st.nfeat + 15) >> 4) * 2048u);
            const int lines = ((st.nfeat + 15) >> 4) * 64;   // 8 KB per chunk
            for (int i = et; i < lines; i += kRowsEpiThreads) asm volatile("prefetch.global.L2 [%0];" ::"l"(blk + (size_t)i * 128));
          }
        }
        mbar_wait(bar_acc, par);
        tc_fence_after();
        RB_STAMP(1);
        RB_STAMP(2);
        RB_STAGE_BEGIN();
        // One hand-off round: this warp's TMEM / shared-memory writes of the round are done.  The proxy fence also waits
        // for the thread's outstanding GLOBAL stores, and the output rows are `feature`-strided (32 sectors per store
        // instruction) — stages with outputs therefore write their smem/TMEM operands first, hand the stage back, and only
        // then store the outputs, which drain while the next stage's MMAs run.
        // Rounds are numbered globally; the number is recomputed from (t, stage) rather than carried: a running counter
        // lived in local memory, i.e. an L2 round trip in front of every signal.
        const uint32_t ground0 = 1u + (uint32_t)t * (uint32_t)P.rounds_per_step + st.round0;
        bool handed = false;
        auto signal_round = [&](bool wrote_smem, uint32_t i) {
          if (wrote_smem) fence_proxy_async_smem();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(bar_round + 8 * ((ground0 + i) & (kRoundBars - 1)));
        };
        auto handoff = [&](bool wrote_smem) {
          RB_STAMP(3);
          RB_STAGE_END();
          signal_round(wrote_smem, 0);
          handed = true;
        };

        switch (st.epi) {
          case R_ACT_H: {
            // rounds of four chunks (one per warp of the quadrant); the next layer's MMAs start behind each round
            const bool elu = st.act == ACT_ELU, addend = kHasAddend && (st.flags & SF_ADDEND) != 0;
            const int nfeat = st.nfeat, nch = (nfeat + 15) >> 4;
            // row-major: this row's Hd floats; tiled: (t, tile = this CTA, chunk 0, quarter 0, row r) — 2048 floats per chunk
            const float* adrow = !addend ? nullptr
                               : V.addend_tiled ? V.addend + ((size_t)t * gridDim.x + blockIdx.x) * (size_t)nch * 2048u + 4 * r
                                                : V.addend + (trow + row) * V.Hd;
            // the activation / addend variant is chosen once per layer, not per chunk
            auto layer = [&](auto act_tag, auto addend_tag) {
              constexpr int ACT = decltype(act_tag)::value;
              constexpr int AD = decltype(addend_tag)::value;   // 0 none, 1 row-major, 2 tiled
              // per-chunk state stepped in registers; the empty asm keeps the compiler from re-deriving it from the chunk index
              uint32_t tcol = tacc + 16u * part, baddr = bias + 64u * part, bar = ground0;
              int left = nch - part;                       // this warp has a chunk in the round while left > 0
              const float* ad = AD == 2 ? adrow + 2048 * part : (AD ? adrow + 16 * part : nullptr);
              int nval = nfeat - 16 * part;
              for (int rd = st.rounds; rd > 0; --rd) {
                asm volatile("" : "+r"(tcol), "+r"(baddr), "+r"(bar), "+r"(left));
                if (left > 0) {
                  if constexpr (AD == 2) {
                    float4 cur[4];
                    rows_addend_prefetch(cur, ad, nval, row_ok);
                    rows_act_chunk<ACT, false, AD>(baddr, 0u, ad, tcol, nval, row_ok, cur);
                  } else {
                    rows_act_chunk<ACT, false, AD>(baddr, 0u, ad, tcol, nval, row_ok);
                  }
                  tmem_st_wait();
                }
                if (rd == 1) {
                  RB_STAMP(3);
                  RB_STAGE_END();
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(bar_round + 8 * (bar & (kRoundBars - 1)));
                tcol += 16u * kEpiParts; baddr += 64u * kEpiParts; ++bar; left -= kEpiParts;
                if (AD) { ad += (AD == 2 ? 2048 : 16) * kEpiParts; nval -= 16 * kEpiParts; }
              }
            };
            using IC_ELU = std::integral_constant<int, ACT_ELU>;
            using IC_RELU = std::integral_constant<int, ACT_RELU>;
            if constexpr (kHasAddend) {
              if (addend && V.addend_tiled) {
                if (elu) layer(IC_ELU{}, std::integral_constant<int, 2>{});
                else layer(IC_RELU{}, std::integral_constant<int, 2>{});
              } else if (addend) {
                if (elu) layer(IC_ELU{}, std::integral_constant<int, 1>{});
                else layer(IC_RELU{}, std::integral_constant<int, 1>{});
              }
            }
            if (!addend) {
              if (elu) layer(IC_ELU{}, std::integral_constant<int, 0>{});
              else layer(IC_RELU{}, std::integral_constant<int, 0>{});
            }
            handed = true;
          } break;

          case R_ACT_DOT: if constexpr (kHasActor) {
            const int nfeat = st.nfeat, nch = (nfeat + 15) >> 4;
            const bool elu = st.act == ACT_ELU;
            const uint32_t wdot = bias + 4u * (nch * 16);   // the 1-output layer's weight row follows the bias
            float dot = 0.f;
            if (elu) {
              for (int ch = part; ch < nch; ch += kEpiParts)
                dot += rows_act_chunk<ACT_ELU, true, 0>(bias + 64u * ch, wdot + 64u * ch, nullptr, tacc + 16u * ch, 16, row_ok);
            } else {
              for (int ch = part; ch < nch; ch += kEpiParts)
                dot += rows_act_chunk<ACT_RELU, true, 0>(bias + 64u * ch, wdot + 64u * ch, nullptr, tacc + 16u * ch, 16, row_ok);
            }
            handoff(false);   // the accumulators are consumed; nothing in X / H changes
            if (part) sts_f(scratch + 4u * (part * 128 + r), dot);
            epi_sync();
            if (part == 0 && row_ok) {
              float* dst = (st.flags & SF_SCALAR_VALUE) ? V.values : V.rewards;
              dst[trow + row] = ((dot + lds_f(scratch + 4u * (128 + r))) + (lds_f(scratch + 4u * (256 + r)) + lds_f(scratch + 4u * (384 + r)))) + lds_f(bias + 4u * (2 * nch * 16));
            }
            // The next writer of `scratch` (the other scalar head, three stages on) is ordered behind these reads by the round
            // barriers — its accumulators need every warp's rounds of the stages in between — but only through mbarriers,
            // which compute-sanitizer's racecheck cannot follow: this barrier (~60 cycles, twice per step) keeps it clean.
            epi_sync();
          } break;

          case R_GRU: {
            const bool last_chunk = (st.flags & RF_LAST_CHUNK) != 0;
            // the belief refresh (bulk copy scratch -> X, issued by one thread) needs every warp's scratch writes of the
            // earlier chunks: each warp fences them at the end of its chunk, this barrier orders all warps behind that
            if (last_chunk && st.unit0 > 0) epi_sync();
            if (!last_chunk) {
              RB_STAMP(3);
          RB_STAGE_END();
            }
            rows_gru_stage(x_hi, x_lo, scr, P.scr_plane, bar_xb, D, tacc, bias, V.beliefs + (trow + row) * D,
                           bar_round + 8 * (ground0 & (kRoundBars - 1)), (uint32_t)(t & 1), r, part, row_ok, st.flags, st.width,
                           st.unit0, st.nfeat,
                           (P.gru_tsc_col >= 0 && (D & 1) == 0) ? tl + ((st.regs & 2) ? kAccCol : 0u) + (uint32_t)P.gru_tsc_col + 8u * part : 0u,
                           N - (row0 + q * 32)
#ifdef RB_STAGE_CLOCK
                           , (V.dbg_clock && blockIdx.x == 0 && et == 0 && t == 5 && st.unit0 == 64) ? V.dbg_clock + 860 : nullptr,
                           ((V.dbg_flags & 32) ? 1 : 0) | ((V.dbg_flags & 512) ? 2 : 0)
#endif
                           );
            if (!last_chunk) handed = true;   // handed back right after the accumulators were read
          } break;

          case R_PRIOR:
          case R_POST: {
            // S <= 32 here (wider states run on the vm.cuh kernel): warp `part` of the quadrant owns states
            // [8 part, 8 part + 8) of its row — mean at acc [0,W), raw std at [W,2W)
            const bool post = kHasPost && st.epi == R_POST;
            const int W = st.width;
            const int c = part * 8;
            const int nv = max(0, min(8, S - c));   // warp-uniform
            const bool want_kl = post && V.kl != nullptr;
            float kl = 0.f;
            float vm[8], vs[8], smp[8];
            if (nv > 0) {
              float pm[8], psd[8];
              const size_t o = (trow + row) * S + c;
              if (want_kl) {
                // The prior's mean / std of this very (row, state group) were computed by this thread two stages ago.  If the
                // prior stage parked them in tensor memory for its transposed stores (see below: same condition, same
                // columns — nothing has written them since: the posterior hidden layer works in the other region, this
                // stage's accumulators occupy columns [0, 2W) of this one), they are read back from there; the global
                // re-read (lane = row, 30-float rows: 32 cache lines per load) made this epilogue 9.3k cycles against 1.5k
                // for the prior's.
                bool parked = false;
                if (s >= 2 && P.stages[s - 2].epi == R_PRIOR && 2 * (int)P.stages[s - 2].width <= 128) {
                  const RStage& pn = P.stages[s - 1];   // the stage that followed the prior head
                  bool pn_reads_h = false;
                  for (int g = pn.gemm_begin; g < pn.gemm_end; ++g) pn_reads_h |= P.gemms[g].a_src == 1;
                  if (!pn_reads_h) {
                    const uint32_t tpr = tl + ((pn.regs & 1) ? 0u : kAccCol) + 128u + (uint32_t)(part * 32);
                    tmem_ld8(tpr + 8, pm);
                    tmem_ld8(tpr + 16, psd);
                    parked = true;
                  }
                }
                if (!parked) {
                  ld_row8_v2<false>(pm, V.prior_m + o, nv, row_ok);
                  ld_row8_v2<false>(psd, V.prior_sd + o, nv, row_ok);
                }
              }
              tmem_ld8(tacc + c, vm);
              tmem_ld8(tacc + W + c, vs);
              const float4 bm0 = lds_f4(bias + 4u * c), bm1 = lds_f4(bias + 4u * c + 16u);
              const float4 bs0 = lds_f4(bias + 4u * (W + c)), bs1 = lds_f4(bias + 4u * (W + c) + 16u);
              const float bm[8] = {bm0.x, bm0.y, bm0.z, bm0.w, bm1.x, bm1.y, bm1.z, bm1.w};
              const float bsd[8] = {bs0.x, bs0.y, bs0.z, bs0.w, bs1.x, bs1.y, bs1.z, bs1.w};
              tmem_ld_wait();
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                vm[i] += bm[i];
                vs[i] = softplus_f(vs[i] + bsd[i]) + V.min_std;
                smp[i] = vm[i] + vs[i] * pre[i];
              }
              if (want_kl) {
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                  if (i < nv && row_ok) {
                    const float inv = rcp_f(psd[i]);   // MUFU forms: ~1e-7 relative, far inside the KL tolerance
                    const float ratio = vs[i] * inv, vr = ratio * ratio;
                    const float dm = (vm[i] - pm[i]) * inv;
                    kl += 0.5f * (vr + dm * dm - 1.f - __logf(vr));
                  }
                }
              }
              if (st.flags & SF_WRITES_STATE) {
                if (((D + c) & 1) == 0) {
#pragma unroll
                  for (int i = 0; i < 8; i += 2) {
                    if (i + 1 < nv) x_put2(x_hi, x_lo, r, D + c + i, smp[i] * pre_nt, smp[i + 1] * pre_nt);
                    else if (i < nv) x_put(x_hi, x_lo, r, D + c + i, smp[i] * pre_nt);
                  }
                } else {
#pragma unroll
                  for (int i = 0; i < 8; ++i)
                    if (i < nv) x_put(x_hi, x_lo, r, D + c + i, smp[i] * pre_nt);
                }
              }
            }
            if (part == kEpiParts - 1 && (st.flags & SF_LOADS_ACTION) && has_next && row_ok)
              for (int j = 0; j < A; ++j) x_put(x_hi, x_lo, r, D + S + j, __ldg(V.actions_in + (trow + N + row) * A + j));
            if (want_kl && part) sts_f(scratch + 4u * (part * 128 + r), kl);
            handoff(true);
            // Output rows are S floats apart, so a store with lane = row touches 32 cache lines per instruction and
            // kept the LSU busy for ~9k cycles.  Transpose through tensor memory instead: park the three 32 x 8 blocks
            // (lane = row) in columns nobody uses right now and read them back in the 16x256b fragment layout, where
            // four neighbouring threads hold 8 consecutive floats of a row — 8 rows x 32 bytes per store instruction.
            // Free columns: the region the NEXT stage does not accumulate into (this stage's dead H operand or its own
            // consumed accumulators, hence the offset past the 2W accumulator columns), provided the next stage reads X.
            const RStage& ns = P.stages[(s + 1 == P.n_rstages) ? 0 : s + 1];
            bool ns_reads_h = false;
            for (int g = ns.gemm_begin; g < ns.gemm_end; ++g) ns_reads_h |= P.gemms[g].a_src == 1;
            if (nv > 0 && (S & 1) == 0 && !ns_reads_h && 2 * W <= 128) {
              const uint32_t tsc = tl + ((ns.regs & 1) ? 0u : kAccCol) + 128u + (uint32_t)(part * 32);
              tmem_st8f(tsc, smp);
              tmem_st8f(tsc + 8, vm);
              tmem_st8f(tsc + 16, vs);
              tmem_st_wait();
              float* const outs[3] = {post ? V.post_s : V.prior_s, post ? V.post_m : V.prior_m, post ? V.post_sd : V.prior_sd};
              const int cq = 2 * (lane & 3);
#pragma unroll
              for (int a = 0; a < 3; ++a) {
#pragma unroll
                for (int hb = 0; hb < 2; ++hb) {
                  float tq[4];
                  tmem_ld_16x256b_x1(tsc + ((uint32_t)(16 * hb) << 16) + 8 * a, tq);
                  tmem_ld_wait();
#pragma unroll
                  for (int rb = 0; rb < 2; ++rb) {
                    const int rr = row0 + q * 32 + 16 * hb + 8 * rb + (lane >> 2);
                    if (rr < N && cq + 1 < nv)
                      *reinterpret_cast<float2*>(outs[a] + (trow + rr) * S + c + cq) = make_float2(tq[2 * rb], tq[2 * rb + 1]);
                  }
                }
              }
            } else if (nv > 0 && row_ok) {   // odd state sizes (rows not 8-byte aligned) and unusual programs
              const size_t o = (trow + row) * S + c;
              float* const outs[3] = {post ? V.post_s : V.prior_s, post ? V.post_m : V.prior_m, post ? V.post_sd : V.prior_sd};
              const float* const vals[3] = {smp, vm, vs};
#pragma unroll
              for (int a = 0; a < 3; ++a)
#pragma unroll
                for (int i = 0; i < 8; ++i)
                  if (i < nv) outs[a][o + i] = vals[a][i];
            }
            if (want_kl) {   // uniform across the CTA: V.kl and the stage kind are kernel-wide
              epi_sync();
              if (part == 0 && row_ok) V.kl[trow + row] = (kl + lds_f(scratch + 4u * (128 + r))) + (lds_f(scratch + 4u * (256 + r)) + lds_f(scratch + 4u * (384 + r)));
              epi_sync();   // as for the scalar heads: orders the next step's partial sums behind these reads for racecheck
            }
          } break;

          case R_ACTION: if constexpr (kHasActor) {
            const int c = part * 8;   // A <= 16 here: warps 0 and 1 of the quadrant take 8 action dims each
            if (c < A) {
              const int W = st.width;
              const float inv_ms = 1.f / V.a_mean_scale;
              float vm[8], vs[8];
              const size_t o = (trow + row) * A + c;
              tmem_ld8(tacc + c, vm);
              tmem_ld8(tacc + W + c, vs);
              const float4 bm0 = lds_f4(bias + 4u * c), bm1 = lds_f4(bias + 4u * c + 16u);
              const float4 bs0 = lds_f4(bias + 4u * (W + c)), bs1 = lds_f4(bias + 4u * (W + c) + 16u);
              const float bm[8] = {bm0.x, bm0.y, bm0.z, bm0.w, bm1.x, bm1.y, bm1.z, bm1.w};
              const float bsd[8] = {bs0.x, bs0.y, bs0.z, bs0.w, bs1.x, bs1.y, bs1.z, bs1.w};
              tmem_ld_wait();
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                const float mean = V.a_mean_scale * tanh_f((vm[i] + bm[i]) * inv_ms);
                const float sd = softplus_f(vs[i] + bsd[i] + V.a_init_std) + V.a_min_std;
                vm[i] = tanh_f(mean + sd * pre[i]);
              }
#pragma unroll
              for (int i = 0; i < 8; ++i)
                if (c + i < A) x_put(x_hi, x_lo, r, D + S + c + i, vm[i]);
              handoff(true);
              if (row_ok && V.actions_out) {
#pragma unroll
                for (int i = 0; i < 8; ++i)
                  if (c + i < A) V.actions_out[o + i] = vm[i];
              }
            }
          } break;

          case R_SCALAR: if constexpr (kHasActor) {
            if (part == 0) {
              float v[16];
              tmem_ld16(tacc, v);
              tmem_ld_wait();
              float* dst = (st.flags & SF_SCALAR_VALUE) ? V.values : V.rewards;
              if (row_ok) dst[trow + row] = v[0] + lds_f(bias);
            }
          } break;
          default: break;
        }

        if (!handed) handoff(st.epi != R_SCALAR);
        RB_STAMP_X(0);
        par ^= 1;
        RB_STAMP_X(1);
      }
    }

    // ---- lambda-return: each row's rewards/values were written by this thread (part 0) ----
    if (part == 0 && row_ok && V.returns && V.rewards && V.values && V.n_steps >= 2) {
      const int T = V.n_steps;
      const float g = V.gamma, lam = V.lambda;
      float last = V.values[(size_t)(T - 1) * N + row];
      float next_v = last;
      for (int t = T - 2; t >= 0; --t) {
        const float rw = V.rewards[(size_t)t * N + row];
        const float inp = rw + g * next_v * V.one_minus_lambda;
        last = inp + g * lam * last;
        V.returns[(size_t)t * N + row] = last;
        next_v = V.values[(size_t)t * N + row];
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, kTmemCols);
}

}  // namespace rb
