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
//   accumulators                    the other TMEM region (RStage::regs); wide hidden layers run as two GEMMs over
//       output-feature halves with separate commit barriers (RF_SPLIT), so the first half's epilogue overlaps the
//       second half's MMAs.
// x*W = hi*hi + lo*hi + hi*lo (three MMAs, fp32 accumulate) as in vm.cuh.
//
// Warps: 0 = weight loader, 1 = MMA issuer + TMEM owner (warpgroup 0 gives its registers away with
// setmaxnreg), 4..11 = epilogue (two warps per TMEM lane quadrant; they split the columns of wide
// layers) running with 208 registers (128*80 + 256*208 <= the 168*384 registers the CTA was launched with).
#pragma once
#include "vm.cuh"

namespace rb {

constexpr int kRowsM = 128;
constexpr int kRowsThreads = 384;   // warpgroup 0: loader, MMA issuer, 2 spare; warpgroups 1-2: epilogue
constexpr int kRowsEpiThreads = 256;
constexpr int kRSlotBytes = 32768;
constexpr int kRSlots = 3;
constexpr int kRMaxStages = 32;
constexpr int kRMaxGemms = 56;
constexpr int kBiasStage = 512;        // floats per smem bias staging buffer (double-buffered)
constexpr uint32_t kAccCol = 256;     // accumulators live at TMEM columns [256, 512)
constexpr uint32_t kXLBO = kRowsM * 16;  // bytes between k-groups of X (128 rows x 16 B)

enum RowsEpi : uint8_t {
  R_ACT_H = 0,   // H[:, f] = act(acc + bias (+ addend))            -> TMEM H
  R_ACTION = 1,  // tanh-Normal action sample                        -> X action slot
  R_GRU = 2,     // one chunk of GRU units -> beliefs[t]; last chunk refreshes X belief slot
  R_PRIOR = 3,
  R_POST = 4,
  R_SCALAR = 5,
  R_ACT_DOT = 6,  // last hidden layer of a scalar head fused with its 1-output layer: out = w . act(acc + b) + b0
};
enum RowsFlags : uint8_t { RF_LAST_CHUNK = 16, RF_SPLIT = 32, RF_PARK = 64, RF_UNPARK = 128 };  // plus SF_* from vm.cuh
// RF_PARK (the GRU chunk before the last) / RF_UNPARK (the last): the new belief units of all earlier chunks are parked,
// already split into fp16 hi/lo, in the accumulator columns the narrow last chunk leaves free, and the belief slot of X
// is refreshed from there (thread-local TMEM loads) instead of from a global read-back.
// RF_SPLIT (R_ACT_H / R_ACT_DOT): the layer runs as two GEMMs over output features [0, 16*width) and the rest; the
// first half's accumulators are committed on their own barrier, so its epilogue runs under the second half's MMAs.

struct RGemm {            // acc[:, acc_col ..+n) (+)= A(128 x 16*ksl) * W(n x 16*ksl)^T
  uint32_t w_off16;       // weight blob offset / 16
  uint16_t slab_bytes16;  // slab bytes / 16  (= n * 4)
  uint16_t n;             // padded output width (multiple of 16)
  uint8_t ksl;            // k16 slabs
  uint8_t a_src;          // 0 = X (smem), 1 = H (tmem)
  uint8_t a_k16;          // first k16 slab inside the source
  uint8_t accumulate;
  uint16_t acc_col;       // column offset inside the accumulator region
  uint16_t pad;
};
struct RStage {
  uint8_t gemm_begin, gemm_end;
  uint8_t epi, flags;
  uint8_t act;
  uint8_t regs;        // TMEM regions: bit0 = accumulator region, bit1 = region holding this stage's H operand
  uint16_t nfeat;      // valid output features (R_ACT_H) / units in this chunk (R_GRU)
  uint16_t bias_off;   // float offset into the bias blob
  uint16_t unit0;      // R_GRU: first unit of the chunk
  uint16_t width;      // R_GRU: padded units per chunk (gate stride in the accumulator); heads: padded half width;
                       // RF_SPLIT: 16-column chunks in the first half
  uint16_t bias_n;     // floats this stage reads from the bias blob (staged in smem by the epilogue warps)
};

struct RowsParams {
  VmParams v;            // dims, scalars, I/O pointers (stage/gemm tables inside are unused here)
  int n_rstages;
  int kh_cols;           // 8 * kh16 (kept for the host-side size checks; H is interleaved hi/lo per 16-feature chunk)
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
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// two floats -> packed fp16 hi pair and lo pair (element 0 in the low half).  Both halves must be fp16:
// tcgen05 kind::f16 rejects mixed fp16 x bf16 operands (tried: illegal instruction), so a cheap
// bf16-by-truncation lo half is not an option.
__device__ __forceinline__ void split2_f16(float x0, float x1, uint32_t& hi, uint32_t& lo) {
  asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(hi) : "f"(x1), "f"(x0));
  const __half2 h = *reinterpret_cast<const __half2*>(&hi);
  const float2 hf = __half22float2(h);
  asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(lo) : "f"(x1 - hf.y), "f"(x0 - hf.x));
}

// element (row, k) of X: byte offset inside the hi (or lo) buffer
__device__ __forceinline__ uint32_t x_off(int row, int k) {
  return (uint32_t)(k >> 3) * kXLBO + (uint32_t)row * 16u + (uint32_t)(k & 7) * 2u;
}
__device__ __forceinline__ void x_put(uint8_t* hi, uint8_t* lo, int row, int k, float v) {
  __half h, l;
  split_f16(v, h, l);
  const uint32_t o = x_off(row, k);
  *reinterpret_cast<__half*>(hi + o) = h;
  *reinterpret_cast<__half*>(lo + o) = l;
}
// 8 consecutive k (one k-group) of one row: a single 16-byte store per half
__device__ __forceinline__ void x_put8(uint8_t* hi, uint8_t* lo, int row, int kgroup, const float* v) {
  uint4 h, l;
  split2_f16(v[0], v[1], h.x, l.x);
  split2_f16(v[2], v[3], h.y, l.y);
  split2_f16(v[4], v[5], h.z, l.z);
  split2_f16(v[6], v[7], h.w, l.w);
  const uint32_t o = (uint32_t)kgroup * kXLBO + (uint32_t)row * 16u;
  *reinterpret_cast<uint4*>(hi + o) = h;
  *reinterpret_cast<uint4*>(lo + o) = l;
}

__host__ __device__ inline size_t rows_smem_bytes(int kx16) {
  return (size_t)kRSlots * kRSlotBytes + 2 * (size_t)kx16 * 2 * kXLBO + 1024 + 2 * kBiasStage * sizeof(float);
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
__device__ __forceinline__ void ld_uni16(float* dst, const float* base) {
  const float4* p = reinterpret_cast<const float4*>(base);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float4 v = p[i];
    dst[4 * i] = v.x; dst[4 * i + 1] = v.y; dst[4 * i + 2] = v.z; dst[4 * i + 3] = v.w;
  }
}
// state pair (k even) of one row: one 4-byte store per half
__device__ __forceinline__ void x_put2(uint8_t* hi, uint8_t* lo, int row, int k, float v0, float v1) {
  uint32_t h, l;
  split2_f16(v0, v1, h, l);
  const uint32_t o = x_off(row, k);
  *reinterpret_cast<uint32_t*>(hi + o) = h;
  *reinterpret_cast<uint32_t*>(lo + o) = l;
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

// H = act(acc + bias (+ addend)): this warp handles 16-column chunks ch = half, half+2, ...
// Pad columns need no guard: their weight rows and bias are zero, so they come out as act(0) = 0.
// The TMEM load of the next chunk is issued before the current one is processed (latency hidden).
// DOT: instead of storing H, reduce it against a weight vector (the scalar head's last layer).
template <int ACT, bool DOT, bool ADDEND>
__device__ __forceinline__ float rows_act_h(const RowsParams& P, const RStage& st, const float* bias, uint32_t tacc,
                                            int ch_begin, int ch_end, int half, int row, bool row_ok, size_t trow) {
  // H is written IN PLACE: accumulator columns [16 ch, 16 ch + 16) of this thread's lane become the packed fp16 hi
  // pairs (8 columns) followed by the lo pairs (8 columns) of the same 16 features, so the region that held the
  // accumulators is the next layer's A operand and the region that held this layer's operand is free for its accumulators.
  const int nfeat = st.nfeat, nch = (nfeat + 15) >> 4;
  const float* wdot = bias + nch * 16;  // DOT: the 1-output layer's weight row follows the bias
  const float* adrow = ADDEND ? P.v.addend + (trow + row) * P.v.Hd : nullptr;
  float dot = 0.f;
  for (int ch = ch_begin + ((half ^ ch_begin) & 1); ch < ch_end; ch += 2) {   // this warp's chunks: ch = half (mod 2)
    float v[16], bz[16];
    const int f0 = ch * 16;
    tmem_ld16(tacc + f0, v);
    ld_uni16(bz, bias + f0);
    if (ADDEND) {
      float ad[16];
      ld_row16(ad, adrow + f0, nfeat - f0, row_ok);
#pragma unroll
      for (int i = 0; i < 16; ++i) bz[i] += ad[i];
    }
    tmem_ld_wait();
    if (DOT) {
      float w[16];
      ld_uni16(w, wdot + f0);
#pragma unroll
      for (int i = 0; i < 16; ++i) dot = fmaf(w[i], act_bf<ACT>(v[i] + bz[i]), dot);
    } else {
      uint32_t hi[8], lo[8];
#pragma unroll
      for (int i = 0; i < 16; i += 2)
        split2_f16(act_bf<ACT>(v[i] + bz[i]), act_bf<ACT>(v[i + 1] + bz[i + 1]), hi[i >> 1], lo[i >> 1]);
      tmem_st8(tacc + f0, hi);
      tmem_st8(tacc + f0 + 8, lo);
    }
  }
  return dot;
}

__global__ void __launch_bounds__(kRowsThreads, 1) rssm_rows_kernel(const __grid_constant__ RowsParams P) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const VmParams& V = P.v;
  uint8_t* ring = smem;
  uint8_t* x_hi = ring + kRSlots * kRSlotBytes;
  const uint32_t x_bytes = (uint32_t)V.kx16 * 2u * kXLBO;
  uint8_t* x_lo = x_hi + x_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(x_lo + x_bytes);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kRSlots + 2);
  float* scratch = reinterpret_cast<float*>(bars + 2 * kRSlots + 4);  // 128 floats: cross-warp partial sums
  float* bias_s = scratch + 128 + 4;                                  // 2 x kBiasStage floats, 16-byte aligned

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row0 = blockIdx.x * kRowsM;
  const uint32_t bar_full = smem_u32(bars), bar_empty = smem_u32(bars + kRSlots);
  const uint32_t bar_acc = smem_u32(bars + 2 * kRSlots), bar_act = smem_u32(bars + 2 * kRSlots + 1);
  const uint32_t bar_acc_a = smem_u32(bars + 2 * kRSlots + 3);   // first part of an RF_SPLIT stage
  const uint32_t bar_acc_b = smem_u32(scratch + 128);            // second part of a three-way split

  if (threadIdx.x == 0) {
    for (int i = 0; i < kRSlots; ++i) {
      mbar_init(bar_full + 8 * i, 1);
      mbar_init(bar_empty + 8 * i, 1);
    }
    mbar_init(bar_acc, 1);
    mbar_init(bar_acc_a, 1);
    mbar_init(bar_acc_b, 1);
    mbar_init(bar_act, kRowsEpiThreads);
    mbar_fence_init();
  }
  if (warp == 1) {
    tmem_alloc(smem_u32(tmem_slot), kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // Each role changes its register budget INSIDE its own branch (ptxas sizes a region by the
  // setmaxnreg that dominates it; a shared if/else before the role split would cap everything at 80).
  if (warp == 0) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 80;");
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
              mbar_arrive_expect_tx(bar_full + 8 * slot, bytes);
              bulk_g2s(smem_u32(ring + slot * kRSlotBytes), src + (size_t)c0 * slab_bytes, bytes, bar_full + 8 * slot);
            }
            __syncwarp();
            if (++slot == kRSlots) { slot = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 80;");
    // ================================ MMA issuer ================================
    uint32_t slot = 0, phase = 0, act_phase = 0;
    const uint64_t x_desc = make_smem_desc(smem_u32(x_hi), kXLBO, 128);
    const uint64_t x_lo_delta = x_bytes >> 4;
    constexpr uint64_t kX_slab = (2u * kXLBO) >> 4;
    const uint32_t ring_a = smem_u32(ring);
    for (int t = 0; t < V.n_steps; ++t) {
      for (int s = 0; s < P.n_rstages; ++s) {
        const int g0 = P.stages[s].gemm_begin, g1 = P.stages[s].gemm_end;
        const uint32_t tacc = tmem_base + ((P.stages[s].regs & 1) ? kAccCol : 0u);
        const uint32_t th = tmem_base + ((P.stages[s].regs & 2) ? kAccCol : 0u);
        const bool split = (P.stages[s].flags & RF_SPLIT) != 0;
        mbar_wait(bar_act, act_phase);
        act_phase ^= 1;
        tc_fence_after();
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
          for (int c0 = 0; c0 < gm.ksl; c0 += per_slot) {
            const int nsl = min(per_slot, (int)gm.ksl - c0);
            mbar_wait(bar_full + 8 * slot, phase);
            tc_fence_after();
            if (elect_one()) {
              uint32_t wa = ring_a + slot * kRSlotBytes;
              for (int j = 0; j < nsl; ++j) {
                const uint64_t b_hi = make_smem_desc(wa, w_lbo, 128);
                const uint64_t b_lo = make_smem_desc(wa + w_lo, w_lbo, 128);
                if (gm.a_src == 0) {
                  const uint64_t a_hi = x_desc + (uint64_t)(kk + j) * kX_slab;
                  umma_f16(d, a_hi, b_hi, idesc, (j == 0) ? acc : 1u);
                  umma_f16(d, a_hi + x_lo_delta, b_hi, idesc, 1u);
                  umma_f16(d, a_hi, b_lo, idesc, 1u);
                } else {
                  const uint32_t a_hi = th + (kk + j) * 16u, a_lo = a_hi + 8u;   // per k-slab: 8 hi columns, 8 lo columns
                  umma_f16_ts(d, a_hi, b_hi, idesc, (j == 0) ? acc : 1u);
                  umma_f16_ts(d, a_lo, b_hi, idesc, 1u);
                  umma_f16_ts(d, a_hi, b_lo, idesc, 1u);
                }
                wa += slab_bytes;
              }
              umma_commit(bar_empty + 8 * slot);
            }
            __syncwarp();
            acc = 1u;
            kk += nsl;
            if (++slot == kRSlots) { slot = 0; phase ^= 1; }
          }
          if (split && g + 1 < g1) {   // this part's accumulators are complete: its epilogue may start
            if (elect_one()) umma_commit(g == g0 ? bar_acc_a : bar_acc_b);
            __syncwarp();
          }
        }
        if (elect_one()) umma_commit(bar_acc);
        __syncwarp();
      }
    }
  } else if (warp < 4) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 80;");  // spare warps of warpgroup 0: the dec is warpgroup-wide
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 208;");
    // ================================ epilogue warps ================================
    const int et = threadIdx.x - 128;      // 0..255
    const int q = warp & 3;                // TMEM lane quadrant
    const int half = (warp - 4) >> 2;      // which of the two warps of this quadrant
    const int r = q * 32 + lane;           // row within the tile = TMEM lane
    const int row = row0 + r;
    const int N = V.N, D = V.D, S = V.S, A = V.A;
    const bool row_ok = row < N;
    const uint32_t tl = tmem_base + ((uint32_t)(q * 32) << 16);
    auto epi_sync = [] { asm volatile("bar.sync 1, 256;" ::: "memory"); };

    // ---- init: zero X, then stage [belief | state*nonterm[0] | action[0]] ----
    {
      const uint32_t words = (2 * x_bytes) / 16;
      uint4* z = reinterpret_cast<uint4*>(x_hi);
      for (uint32_t i = et; i < words; i += kRowsEpiThreads) z[i] = make_uint4(0, 0, 0, 0);
      epi_sync();
      if (row_ok) {
        if (V.init_belief) {
          const float* b = V.init_belief + (size_t)row * D;
          for (int kg = half; kg * 8 < D; kg += 2) {
            float v[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] = (kg * 8 + i < D) ? b[kg * 8 + i] : 0.f;
            if (kg * 8 + 8 <= D) x_put8(x_hi, x_lo, r, kg, v);
            else
              for (int i = 0; kg * 8 + i < D; ++i) x_put(x_hi, x_lo, r, kg * 8 + i, v[i]);
          }
        }
        if (half == 0 && V.init_state) {
          const float m = V.nonterm ? V.nonterm[row] : 1.f;
          for (int j = 0; j < S; ++j) x_put(x_hi, x_lo, r, D + j, V.init_state[(size_t)row * S + j] * m);
        }
        if (half == 1 && V.actions_in)
          for (int j = 0; j < A; ++j) x_put(x_hi, x_lo, r, D + S + j, V.actions_in[(size_t)row * A + j]);
      }
      fence_proxy_async_smem();
      mbar_arrive(bar_act);
    }

    // With ~220 KB of shared memory in use there is next to no L1: every global load is an L2
    // round trip.  So while the MMAs of a stage run, the epilogue warps already fetch what that
    // stage's epilogue will need: its biases into a smem staging buffer, its per-row inputs
    // (noise, previous belief) into registers.
    float pre[16];  // per-row inputs (noise) of the upcoming stage; the GRU reads b_prev back from X
    float pre_nt = 1.f;
    auto prefetch = [&](int t, int s, int buf) {
      const RStage& st = P.stages[s];
      const size_t trow = (size_t)t * N;
      float* dstb = bias_s + buf * kBiasStage;
      for (int i = et; i < st.bias_n; i += kRowsEpiThreads) dstb[i] = __ldg(V.bias + st.bias_off + i);
      switch (st.epi) {
        case R_PRIOR:
        case R_POST: {
          {  // the two warps of a quadrant take one 16-state chunk each
            const int c = half * 16;
            const float* eps = (st.epi == R_POST ? V.eps_post : V.eps_prior) + (trow + row) * S + c;
            ld_row16_v2(pre, eps, S - c, row_ok && c < S);
            pre_nt = 1.f;
            if ((st.flags & SF_WRITES_STATE) && V.nonterm && (t + 1) < V.n_steps && row_ok) pre_nt = __ldg(V.nonterm + trow + N + row);
          }
        } break;
        case R_ACTION: {
          if (half == 0) {
            const float* eps = V.eps_action + (trow + row) * A;
#pragma unroll
            for (int i = 0; i < 16; ++i) pre[i] = (row_ok && i < A) ? __ldg(eps + i) : 0.f;
          }
        } break;
        default: break;
      }
    };

    uint32_t acc_phase = 0, acc_a_phase = 0, acc_b_phase = 0;
    int buf = 0;
    {  // stage 0 of step 0 never needs per-row inputs ahead of time in either program; stage its biases
      const RStage& st0 = P.stages[0];
      for (int i = et; i < st0.bias_n; i += kRowsEpiThreads) bias_s[i] = __ldg(V.bias + st0.bias_off + i);
      if (st0.epi == R_GRU || st0.epi == R_PRIOR || st0.epi == R_POST || st0.epi == R_ACTION) __trap();
    }
    for (int t = 0; t < V.n_steps; ++t) {
      const size_t trow = (size_t)t * N;
      const bool has_next = (t + 1) < V.n_steps;
      for (int s = 0; s < P.n_rstages; ++s) {
        const RStage& st = P.stages[s];
        const float* bias = bias_s + buf * kBiasStage;
        const uint32_t tacc = tl + ((st.regs & 1) ? kAccCol : 0u);
        const bool split = (st.flags & RF_SPLIT) != 0;
        if (split) {
          mbar_wait(bar_acc_a, acc_a_phase);
          acc_a_phase ^= 1;
        } else {
          mbar_wait(bar_acc, acc_phase);
          acc_phase ^= 1;
        }
        tc_fence_after();
        epi_sync();  // staged biases visible to every epilogue warp
        if (V.dbg_clock && blockIdx.x == 0 && et == 0) V.dbg_clock[(t * P.n_rstages + s) * 2] = clock64();
        // The proxy fence before the hand-off also waits for this thread's outstanding GLOBAL stores, and the
        // output rows are `feature`-strided (32 sectors per store instruction).  Stages with outputs therefore
        // write their smem/TMEM operands first, hand the stage back, and only then store the outputs, which
        // drain while the next stage's MMAs run.
        bool handed = false, prefetched = false;
        auto handoff = [&] {
          fence_proxy_async_smem();
          tc_fence_before();
          if (V.dbg_clock && blockIdx.x == 0 && et == 0) V.dbg_clock[(t * P.n_rstages + s) * 2 + 1] = clock64();
          mbar_arrive(bar_act);
          handed = true;
        };

        switch (st.epi) {
          case R_ACT_H: {
            const bool elu = st.act == ACT_ELU, addend = (st.flags & SF_ADDEND) != 0;
            const int nch = (st.nfeat + 15) >> 4;
            auto part = [&](int c0, int c1) {
              if (addend) {
                if (elu) rows_act_h<ACT_ELU, false, true>(P, st, bias, tacc, c0, c1, half, row, row_ok, trow);
                else rows_act_h<ACT_RELU, false, true>(P, st, bias, tacc, c0, c1, half, row, row_ok, trow);
              } else {
                if (elu) rows_act_h<ACT_ELU, false, false>(P, st, bias, tacc, c0, c1, half, row, row_ok, trow);
                else rows_act_h<ACT_RELU, false, false>(P, st, bias, tacc, c0, c1, half, row, row_ok, trow);
              }
            };
            if (split) {
              part(0, st.width);               // under the next part's MMAs
              int c = st.width;
              if (st.unit0) {
                mbar_wait(bar_acc_b, acc_b_phase);
                acc_b_phase ^= 1;
                tc_fence_after();
                part(c, st.unit0);
                c = st.unit0;
              }
              mbar_wait(bar_acc, acc_phase);
              acc_phase ^= 1;
              tc_fence_after();
              part(c, nch);
            } else {
              part(0, nch);
            }
            tmem_st_wait();
          } break;

          case R_ACT_DOT: {
            const int nch = (st.nfeat + 15) >> 4;
            const bool elu = st.act == ACT_ELU;
            float dot = elu ? rows_act_h<ACT_ELU, true, false>(P, st, bias, tacc, 0, split ? (int)st.width : nch, half, row, row_ok, trow)
                            : rows_act_h<ACT_RELU, true, false>(P, st, bias, tacc, 0, split ? (int)st.width : nch, half, row, row_ok, trow);
            if (split) {
              int c = st.width;
              if (st.unit0) {
                mbar_wait(bar_acc_b, acc_b_phase);
                acc_b_phase ^= 1;
                tc_fence_after();
                dot += elu ? rows_act_h<ACT_ELU, true, false>(P, st, bias, tacc, c, st.unit0, half, row, row_ok, trow)
                           : rows_act_h<ACT_RELU, true, false>(P, st, bias, tacc, c, st.unit0, half, row, row_ok, trow);
                c = st.unit0;
              }
              mbar_wait(bar_acc, acc_phase);
              acc_phase ^= 1;
              tc_fence_after();
              dot += elu ? rows_act_h<ACT_ELU, true, false>(P, st, bias, tacc, c, nch, half, row, row_ok, trow)
                         : rows_act_h<ACT_RELU, true, false>(P, st, bias, tacc, c, nch, half, row, row_ok, trow);
            }
            if (half == 1) scratch[r] = dot;
            epi_sync();
            if (half == 0 && row_ok) {
              const int nch = (st.nfeat + 15) >> 4;
              float* dst = (st.flags & SF_SCALAR_VALUE) ? V.values : V.rewards;
              dst[trow + row] = dot + scratch[r] + bias[2 * nch * 16];
            }
          } break;

          case R_GRU: {
            // accumulator: r at [0,W), z at [W,2W), i_n at [2W,3W), h_n at [3W,4W), W = st.width
            const int W = st.width, u0 = st.unit0, nu = st.nfeat;  // nu valid units in this chunk
            const bool last_chunk = (st.flags & RF_LAST_CHUNK) != 0;
            float bnew[2][16];
#pragma unroll
            for (int k = 0; k < 2; ++k) {
              const int c = (half + 2 * k) * 16;
              if (c < nu) {
                float vr[16], vz[16], vi[16], vh[16], bb[16];
                tmem_ld16(tacc + c, vr);
                tmem_ld16(tacc + W + c, vz);
                tmem_ld16(tacc + 2 * W + c, vi);
                tmem_ld16(tacc + 3 * W + c, vh);
                tmem_ld_wait();
                ld_uni16(bb, bias + c);
#pragma unroll
                for (int i = 0; i < 16; ++i) vr[i] = sigmoid_f(vr[i] + bb[i]);
                ld_uni16(bb, bias + W + c);
#pragma unroll
                for (int i = 0; i < 16; ++i) vz[i] = sigmoid_f(vz[i] + bb[i]);
                ld_uni16(bb, bias + 3 * W + c);
#pragma unroll
                for (int i = 0; i < 16; ++i) vh[i] = vr[i] * (vh[i] + bb[i]);
                ld_uni16(bb, bias + 2 * W + c);
                // b_prev = hi + lo from the belief slot of X (exact to 2^-22 relative; this thread's units, which the
                // refresh below overwrites only after every warp has passed the epi_sync that follows the last chunk)
                float bp[16];
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                  const int kg = (u0 + c) / 8 + j;
                  if (kg * 8 < D) {
                    const uint4 h4 = *reinterpret_cast<const uint4*>(x_hi + (uint32_t)kg * kXLBO + (uint32_t)r * 16u);
                    const uint4 l4 = *reinterpret_cast<const uint4*>(x_lo + (uint32_t)kg * kXLBO + (uint32_t)r * 16u);
                    const uint32_t hw[4] = {h4.x, h4.y, h4.z, h4.w}, lw[4] = {l4.x, l4.y, l4.z, l4.w};
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                      const float2 hf = __half22float2(*reinterpret_cast<const __half2*>(&hw[e]));
                      const float2 lf = __half22float2(*reinterpret_cast<const __half2*>(&lw[e]));
                      bp[8 * j + 2 * e] = hf.x + lf.x;
                      bp[8 * j + 2 * e + 1] = hf.y + lf.y;
                    }
                  } else {
#pragma unroll
                    for (int e = 0; e < 8; ++e) bp[8 * j + e] = 0.f;
                  }
                }
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                  const float nn = tanh_f(vi[i] + bb[i] + vh[i]);
                  bnew[k][i] = nn + vz[i] * (bp[i] - nn);   // (1-z)*n + z*b_prev
                }
              }
            }
            if (!last_chunk) handoff();  // X is untouched by this chunk: the accumulators are all the issuer waits for
#pragma unroll
            for (int k = 0; k < 2; ++k) {
              const int c = (half + 2 * k) * 16;
              if (c < nu && row_ok) st_row16(V.beliefs + (trow + row) * D + u0 + c, bnew[k], min(16, nu - c));
            }
            if (st.flags & RF_PARK) {
              // The last chunk is narrow: its MMAs only touch accumulator columns [0, 4 W_last).  Park this row's new
              // belief units of this and the earlier chunks (re-read: this thread's own stores) behind them as packed
              // fp16 hi/lo — 16 columns per 16-unit group, groups g = half (mod 2) are this thread's — while the last
              // chunk's MMAs run.  The sibling warp shares these TMEM lanes: wait until it has drained its accumulators.
              // (Re-reading both earlier chunks in one batch, or ahead of this chunk's stores, was tried: the 64 extra
              // live registers spill into the hidden-layer epilogues, 160.7k -> 184k cycles per step.)
              epi_sync();
              const uint32_t tpark = tacc + 4u * P.stages[s + 1].width;
              const int nprev = u0 >> 6;
              const float* brow = V.beliefs + (trow + row) * D;
              for (int pc = 0; pc <= nprev; ++pc) {
                float pv[2][16];
                if (pc < nprev) {
#pragma unroll
                  for (int k = 0; k < 2; ++k) ld_row16(pv[k], brow + 64 * pc + (half + 2 * k) * 16, 16, row_ok);
                } else {
#pragma unroll
                  for (int k = 0; k < 2; ++k)
#pragma unroll
                    for (int i = 0; i < 16; ++i) pv[k][i] = bnew[k][i];
                }
#pragma unroll
                for (int k = 0; k < 2; ++k) {
                  uint32_t hl[16];
#pragma unroll
                  for (int i = 0; i < 8; ++i) split2_f16(pv[k][2 * i], pv[k][2 * i + 1], hl[i], hl[8 + i]);
                  tmem_st16f(tpark + 16u * (uint32_t)(4 * pc + half + 2 * k), reinterpret_cast<const float*>(hl));
                }
              }
              tmem_st_wait();
            }
            if ((st.flags & RF_LAST_CHUNK) && (st.flags & RF_UNPARK)) {
              // every chunk's MMAs are done: the belief slot of X may be overwritten.  Earlier chunks: from the parked
              // columns (this thread's own groups — the same ones whose old values it read for the blend, so no
              // cross-thread hazard and no barrier); this chunk: from registers.
              const uint32_t tpark = tacc + 4u * W;
              const int ngroups = u0 >> 4;
              for (int g = half; g < ngroups; g += 2) {
                uint32_t hl[16];
                tmem_ld16(tpark + 16u * g, reinterpret_cast<float*>(hl));
                tmem_ld_wait();
                const uint32_t o = (uint32_t)(2 * g) * kXLBO + (uint32_t)r * 16u;
                *reinterpret_cast<uint4*>(x_hi + o) = make_uint4(hl[0], hl[1], hl[2], hl[3]);
                *reinterpret_cast<uint4*>(x_hi + o + kXLBO) = make_uint4(hl[4], hl[5], hl[6], hl[7]);
                *reinterpret_cast<uint4*>(x_lo + o) = make_uint4(hl[8], hl[9], hl[10], hl[11]);
                *reinterpret_cast<uint4*>(x_lo + o + kXLBO) = make_uint4(hl[12], hl[13], hl[14], hl[15]);
              }
#pragma unroll
              for (int k = 0; k < 2; ++k) {
                const int c = (half + 2 * k) * 16;
                if (c < nu) {
#pragma unroll
                  for (int j = 0; j < 2; ++j) {
                    const int u = u0 + c + 8 * j;
                    if (u + 8 <= D) x_put8(x_hi, x_lo, r, u >> 3, bnew[k] + 8 * j);
                    else
                      for (int i = 0; u + i < D && i < 8; ++i) x_put(x_hi, x_lo, r, u + i, bnew[k][8 * j + i]);
                  }
                }
              }
            } else if (st.flags & RF_LAST_CHUNK) {
              // every chunk's MMAs are done: now the belief slot of X may be overwritten.  Rows were
              // written by both warps of the quadrant, so sync the epilogue warps first; the loads
              // are issued in batches so the L2 latency is paid once per batch, not once per k-group.
              __threadfence_block();
              epi_sync();
              const float* b = V.beliefs + (trow + row) * D;
              const int nkg = D >> 3;  // full k-groups
              for (int kg0 = half; kg0 < nkg; kg0 += 8) {
                float v[4][8];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                  const int kg = kg0 + 2 * j;
                  if (kg < nkg && row_ok) {
                    const float4 lo4 = *reinterpret_cast<const float4*>(b + kg * 8);
                    const float4 hi4 = *reinterpret_cast<const float4*>(b + kg * 8 + 4);
                    v[j][0] = lo4.x; v[j][1] = lo4.y; v[j][2] = lo4.z; v[j][3] = lo4.w;
                    v[j][4] = hi4.x; v[j][5] = hi4.y; v[j][6] = hi4.z; v[j][7] = hi4.w;
                  }
                }
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                  const int kg = kg0 + 2 * j;
                  if (kg < nkg && row_ok) x_put8(x_hi, x_lo, r, kg, v[j]);
                }
              }
              if (half == 0 && row_ok)
                for (int k = nkg * 8; k < D; ++k) x_put(x_hi, x_lo, r, k, b[k]);
            }
          } break;

          case R_PRIOR:
          case R_POST: {
            // S <= 32 here (wider states run on the vm.cuh kernel): warp `half` of the quadrant owns states
            // [16 half, 16 half + 16) of its row — mean at acc [0,W), raw std at [W,2W)
            const bool post = st.epi == R_POST;
            const int W = st.width;
            const int c = half * 16;
            const int nv = min(16, S - c);
            const bool want_kl = post && V.kl != nullptr;
            float kl = 0.f;
            float vm[16], vs[16], smp[16];
            if (nv > 0) {
              float pm[16], psd[16];
              const size_t o = (trow + row) * S + c;
              if (want_kl) {
                ld_row16_v2(pm, V.prior_m + o, nv, row_ok);
                ld_row16_v2(psd, V.prior_sd + o, nv, row_ok);
              }
              tmem_ld16(tacc + c, vm);
              tmem_ld16(tacc + W + c, vs);
              tmem_ld_wait();
#pragma unroll
              for (int i = 0; i < 16; ++i) {
                vm[i] += bias[c + i];
                vs[i] = softplus_f(vs[i] + bias[W + c + i]) + V.min_std;
                smp[i] = vm[i] + vs[i] * pre[i];
              }
              if (want_kl) {
#pragma unroll
                for (int i = 0; i < 16; ++i) {
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
                  for (int i = 0; i < 16; i += 2) {
                    if (i + 1 < nv) x_put2(x_hi, x_lo, r, D + c + i, smp[i] * pre_nt, smp[i + 1] * pre_nt);
                    else if (i < nv) x_put(x_hi, x_lo, r, D + c + i, smp[i] * pre_nt);
                  }
                } else {
#pragma unroll
                  for (int i = 0; i < 16; ++i)
                    if (i < nv) x_put(x_hi, x_lo, r, D + c + i, smp[i] * pre_nt);
                }
              }
            }
            if (half == 1) {
              if ((st.flags & SF_LOADS_ACTION) && has_next && row_ok)
                for (int j = 0; j < A; ++j) x_put(x_hi, x_lo, r, D + S + j, __ldg(V.actions_in + (trow + N + row) * A + j));
              if (want_kl) scratch[r] = kl;
            }
            handoff();
            // Output rows are S floats apart, so a store with lane = row touches 32 cache lines per instruction and the
            // 24 of them kept the LSU busy for ~9k cycles.  Transpose through tensor memory instead: park the three
            // 32 x 16 blocks (lane = row) in columns nobody uses right now and read them back in the 16x256b fragment
            // layout, where four neighbouring threads hold 8 consecutive floats of a row — 8 rows x 32 bytes per store.
            // Free columns: the region the NEXT stage does not accumulate into (this stage's dead H operand or its own
            // consumed accumulators, hence the offset past the 2W accumulator columns), provided the next stage reads X.
            const RStage& ns = P.stages[(s + 1 == P.n_rstages) ? 0 : s + 1];
            bool ns_reads_h = false;
            for (int g = ns.gemm_begin; g < ns.gemm_end; ++g) ns_reads_h |= P.gemms[g].a_src == 1;
            if (nv > 0 && (S & 1) == 0 && !ns_reads_h && 2 * W <= 128) {
              const uint32_t tsc = tl + ((ns.regs & 1) ? 0u : kAccCol) + 128u + (uint32_t)(half * 64);
              tmem_st16f(tsc, smp);
              tmem_st16f(tsc + 16, vm);
              tmem_st16f(tsc + 32, vs);
              tmem_st_wait();
              float* const outs[3] = {post ? V.post_s : V.prior_s, post ? V.post_m : V.prior_m, post ? V.post_sd : V.prior_sd};
              const int cq = 2 * (lane & 3);
#pragma unroll
              for (int a = 0; a < 3; ++a) {
#pragma unroll
                for (int hb = 0; hb < 2; ++hb) {
                  float tq[8];
                  tmem_ld_16x256b_x2(tsc + ((uint32_t)(16 * hb) << 16) + 16 * a, tq);
                  tmem_ld_wait();
#pragma unroll
                  for (int rb = 0; rb < 2; ++rb) {
                    const int rr = row0 + q * 32 + 16 * hb + 8 * rb + (lane >> 2);
                    if (rr < N) {
                      float* dst = outs[a] + (trow + rr) * S + c + cq;
#pragma unroll
                      for (int g = 0; g < 2; ++g)
                        if (cq + 8 * g + 1 < nv)
                          *reinterpret_cast<float2*>(dst + 8 * g) = make_float2(tq[4 * g + 2 * rb], tq[4 * g + 2 * rb + 1]);
                    }
                  }
                }
              }
            } else if (nv > 0 && row_ok) {   // odd state sizes (rows not 8-byte aligned) and unusual programs
              const size_t o = (trow + row) * S + c;
              st_row16_v2((post ? V.post_s : V.prior_s) + o, smp, nv);
              st_row16_v2((post ? V.post_m : V.prior_m) + o, vm, nv);
              st_row16_v2((post ? V.post_sd : V.prior_sd) + o, vs, nv);
            }
            if (want_kl) {   // uniform across the CTA: V.kl and the stage kind are kernel-wide
              epi_sync();
              if (half == 0 && row_ok) V.kl[trow + row] = kl + scratch[r];
            }
          } break;

          case R_ACTION: {
            if (half == 0) {   // A <= 16 here
              const int W = st.width;
              const float inv_ms = 1.f / V.a_mean_scale;
              float vm[16], vs[16];
              const size_t o = (trow + row) * A;
              tmem_ld16(tacc, vm);
              tmem_ld16(tacc + W, vs);
              tmem_ld_wait();
#pragma unroll
              for (int i = 0; i < 16; ++i) {
                const float mean = V.a_mean_scale * tanh_f((vm[i] + bias[i]) * inv_ms);
                const float sd = softplus_f(vs[i] + bias[W + i] + V.a_init_std) + V.a_min_std;
                vm[i] = tanh_f(mean + sd * pre[i]);
              }
#pragma unroll
              for (int i = 0; i < 16; ++i)
                if (i < A) x_put(x_hi, x_lo, r, D + S + i, vm[i]);
              handoff();
              if (row_ok && V.actions_out) st_row16(V.actions_out + o, vm, A);
            }
          } break;

          case R_SCALAR: {
            if (half == 0) {
              float v[16];
              tmem_ld16(tacc, v);
              tmem_ld_wait();
              float* dst = (st.flags & SF_SCALAR_VALUE) ? V.values : V.rewards;
              if (row_ok) dst[trow + row] = v[0] + bias[0];
            }
          } break;
          default: break;
        }

        if (!handed) handoff();
        if (V.dbg_clock && blockIdx.x == 0 && et == 0 && t == 5) V.dbg_clock[600 + s * 2] = clock64();
        // fetch for the next stage while its MMAs run
        buf ^= 1;
        {
          const bool wrap = (s + 1 == P.n_rstages);
          if (!prefetched && (!wrap || has_next)) prefetch(wrap ? t + 1 : t, wrap ? 0 : s + 1, buf);
        }
        if (V.dbg_clock && blockIdx.x == 0 && et == 0 && t == 5) V.dbg_clock[600 + s * 2 + 1] = clock64();
      }
    }

    // ---- lambda-return: each row's rewards/values were written by this thread (half 0) ----
    if (half == 0 && row_ok && V.returns && V.rewards && V.values && V.n_steps >= 2) {
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
