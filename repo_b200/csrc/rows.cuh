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
//   H = hidden activations          fp16 hi/lo packed pairs in TENSOR memory, columns [0,256) — A operand
//       of the layers that read it (TS-mode MMA), written by the epilogue with tcgen05.st;
//   W = weights                     B operand, pre-packed (pack.cuh, PackRowsJob) slabs [hi | lo] of
//       Npad x 16 K, streamed L2 -> smem ring by the TMA engine (1-D bulk copies);
//   accumulators                    TMEM columns [256,512).
// x*W = hi*hi + lo*hi + hi*lo (three MMAs, fp32 accumulate) as in vm.cuh.
//
// Warps: 0 = weight loader, 1 = MMA issuer + TMEM owner, 2..9 = epilogue (two warps per TMEM lane
// quadrant; they split the columns of wide layers).
#pragma once
#include "vm.cuh"

namespace rb {

constexpr int kRowsM = 128;
constexpr int kRowsThreads = 320;
constexpr int kRowsEpiThreads = 256;
constexpr int kRSlotBytes = 32768;
constexpr int kRSlots = 3;
constexpr int kRMaxStages = 32;
constexpr int kRMaxGemms = 56;
constexpr uint32_t kAccCol = 256;     // accumulators live at TMEM columns [256, 512)
constexpr uint32_t kXLBO = kRowsM * 16;  // bytes between k-groups of X (128 rows x 16 B)

enum RowsEpi : uint8_t {
  R_ACT_H = 0,   // H[:, f] = act(acc + bias (+ addend))            -> TMEM H
  R_ACTION = 1,  // tanh-Normal action sample                        -> X action slot
  R_GRU = 2,     // one chunk of GRU units -> beliefs[t]; last chunk refreshes X belief slot
  R_PRIOR = 3,
  R_POST = 4,
  R_SCALAR = 5,
};
enum RowsFlags : uint8_t { RF_LAST_CHUNK = 16 };  // plus SF_* from vm.cuh

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
  uint8_t act, pad0;
  uint16_t nfeat;      // valid output features (R_ACT_H) / units in this chunk (R_GRU)
  uint16_t bias_off;   // float offset into the bias blob
  uint16_t unit0;      // R_GRU: first unit of the chunk
  uint16_t width;      // R_GRU: padded units per chunk (gate stride in the accumulator); heads: padded half width
  uint16_t pad1;
};

struct RowsParams {
  VmParams v;            // dims, scalars, I/O pointers (stage/gemm tables inside are unused here)
  int n_rstages;
  int kh_cols;           // TMEM columns per H half = 8 * kh16
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
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// two floats -> packed fp16 hi pair and lo pair (element 0 in the low half)
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
  return (size_t)kRSlots * kRSlotBytes + 2 * (size_t)kx16 * 2 * kXLBO + 256;
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
// 16 floats every lane reads alike (bias blob; offsets are multiples of 16 floats -> 64-byte aligned)
__device__ __forceinline__ void ld_uni16(float* dst, const float* base) {
  const float4* p = reinterpret_cast<const float4*>(base);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float4 v = __ldg(p + i);
    dst[4 * i] = v.x; dst[4 * i + 1] = v.y; dst[4 * i + 2] = v.z; dst[4 * i + 3] = v.w;
  }
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
// branch-free activations (all 16 lanes of a chunk stay independent -> ILP across elements)
template <int ACT>
__device__ __forceinline__ float act_bf(float x) {
  if (ACT == ACT_ELU) return fmaxf(x, 0.f) + (ex2_f(fminf(x, 0.f) * 1.4426950408889634f) - 1.f);
  return fmaxf(x, 0.f);
}

// H = act(acc + bias (+ addend)): this warp handles 16-column chunks ch = half, half+2, ...
// Pad columns need no guard: their weight rows and bias are zero, so they come out as act(0) = 0.
template <int ACT>
__device__ __forceinline__ void rows_act_h(const RowsParams& P, const RStage& st, uint32_t tacc, uint32_t th_hi,
                                           uint32_t th_lo, int half, int row, bool row_ok, size_t trow) {
  const int nfeat = st.nfeat, nch = (nfeat + 15) >> 4;
  const bool addend = (st.flags & SF_ADDEND) != 0;
  const float* bias = P.v.bias + st.bias_off;
  for (int ch = half; ch < nch; ch += 2) {
    float v[16], bz[16];
    const int f0 = ch * 16;
    tmem_ld16(tacc + f0, v);
    ld_uni16(bz, bias + f0);
    if (addend) {
      float ad[16];
      ld_row16(ad, P.v.addend + (trow + row) * P.v.Hd + f0, nfeat - f0, row_ok);
#pragma unroll
      for (int i = 0; i < 16; ++i) bz[i] += ad[i];
    }
    tmem_ld_wait();
    uint32_t hi[8], lo[8];
#pragma unroll
    for (int i = 0; i < 16; i += 2)
      split2_f16(act_bf<ACT>(v[i] + bz[i]), act_bf<ACT>(v[i + 1] + bz[i + 1]), hi[i >> 1], lo[i >> 1]);
    tmem_st8(th_hi + ch * 8, hi);
    tmem_st8(th_lo + ch * 8, lo);
  }
  tmem_st_wait();
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

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row0 = blockIdx.x * kRowsM;
  const uint32_t bar_full = smem_u32(bars), bar_empty = smem_u32(bars + kRSlots);
  const uint32_t bar_acc = smem_u32(bars + 2 * kRSlots), bar_act = smem_u32(bars + 2 * kRSlots + 1);

  if (threadIdx.x == 0) {
    for (int i = 0; i < kRSlots; ++i) {
      mbar_init(bar_full + 8 * i, 1);
      mbar_init(bar_empty + 8 * i, 1);
    }
    mbar_init(bar_acc, 1);
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

  if (warp == 0) {
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
    // ================================ MMA issuer ================================
    uint32_t slot = 0, phase = 0, act_phase = 0;
    const uint64_t x_desc = make_smem_desc(smem_u32(x_hi), kXLBO, 128);
    const uint64_t x_lo_delta = x_bytes >> 4;
    constexpr uint64_t kX_slab = (2u * kXLBO) >> 4;
    const uint32_t th_hi = tmem_base, th_lo = tmem_base + (uint32_t)P.kh_cols;
    const uint32_t tacc = tmem_base + kAccCol;
    const uint32_t ring_a = smem_u32(ring);
    for (int t = 0; t < V.n_steps; ++t) {
      for (int s = 0; s < P.n_rstages; ++s) {
        const int g0 = P.stages[s].gemm_begin, g1 = P.stages[s].gemm_end;
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
                  const uint32_t a_hi = th_hi + (kk + j) * 8u, a_lo = th_lo + (kk + j) * 8u;
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
        }
        if (elect_one()) umma_commit(bar_acc);
        __syncwarp();
      }
    }
  } else {
    // ================================ epilogue warps ================================
    const int et = threadIdx.x - 64;       // 0..255
    const int q = warp & 3;                // TMEM lane quadrant
    const int half = (warp - 2) >> 2;      // which of the two warps of this quadrant
    const int r = q * 32 + lane;           // row within the tile = TMEM lane
    const int row = row0 + r;
    const int N = V.N, D = V.D, S = V.S, A = V.A;
    const bool row_ok = row < N;
    const uint32_t tl = tmem_base + ((uint32_t)(q * 32) << 16);
    const uint32_t th_hi = tl, th_lo = tl + (uint32_t)P.kh_cols, tacc = tl + kAccCol;
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

    uint32_t acc_phase = 0;
    for (int t = 0; t < V.n_steps; ++t) {
      const size_t trow = (size_t)t * N;
      const bool has_next = (t + 1) < V.n_steps;
      for (int s = 0; s < P.n_rstages; ++s) {
        const RStage& st = P.stages[s];
        mbar_wait(bar_acc, acc_phase);
        acc_phase ^= 1;
        tc_fence_after();

        switch (st.epi) {
          case R_ACT_H: {
            if (st.act == ACT_ELU) rows_act_h<ACT_ELU>(P, st, tacc, th_hi, th_lo, half, row, row_ok, trow);
            else rows_act_h<ACT_RELU>(P, st, tacc, th_hi, th_lo, half, row, row_ok, trow);
          } break;

          case R_GRU: {
            // accumulator: r at [0,W), z at [W,2W), i_n at [2W,3W), h_n at [3W,4W), W = st.width
            const int W = st.width, u0 = st.unit0, nu = st.nfeat;  // nu valid units in this chunk
            const float* bias = V.bias + st.bias_off;              // [r | z | in | hn] each W floats
            const float* bprev = (t == 0) ? V.init_belief : (V.beliefs + (trow - N) * D);
            for (int sub = half; sub * 16 < nu; sub += 2) {
              const int c = sub * 16, nv = min(16, nu - c);
              float vr[16], vz[16], vi[16], vh[16], bo[16], bb[16];
              tmem_ld16(tacc + c, vr);
              tmem_ld16(tacc + W + c, vz);
              tmem_ld16(tacc + 2 * W + c, vi);
              tmem_ld16(tacc + 3 * W + c, vh);
              if (bprev) ld_row16(bo, bprev + (size_t)row * D + u0 + c, nv, row_ok);
              else {
#pragma unroll
                for (int i = 0; i < 16; ++i) bo[i] = 0.f;
              }
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
#pragma unroll
              for (int i = 0; i < 16; ++i) {
                const float nn = tanh_f(vi[i] + bb[i] + vh[i]);
                bo[i] = nn + vz[i] * (bo[i] - nn);   // (1-z)*n + z*b
              }
              if (row_ok) st_row16(V.beliefs + (trow + row) * D + u0 + c, bo, nv);
            }
            if (st.flags & RF_LAST_CHUNK) {
              // every chunk's MMAs are done: now the belief slot of X may be overwritten.  Rows were
              // written by both warps of the quadrant, so sync the epilogue warps first.
              __threadfence_block();
              epi_sync();
              if (row_ok) {
                const float* b = V.beliefs + (trow + row) * D;
                for (int kg = half; kg * 8 < D; kg += 2) {
                  float v[8];
#pragma unroll
                  for (int i = 0; i < 8; ++i) v[i] = (kg * 8 + i < D) ? b[kg * 8 + i] : 0.f;
                  if (kg * 8 + 8 <= D) x_put8(x_hi, x_lo, r, kg, v);
                  else
                    for (int i = 0; kg * 8 + i < D; ++i) x_put(x_hi, x_lo, r, kg * 8 + i, v[i]);
                }
              }
            }
          } break;

          case R_PRIOR:
          case R_POST: {
            if (half == 0) {
              const bool post = st.epi == R_POST;
              const int W = st.width;  // mean at [0,W), raw std at [W,2W)
              const float* bias = V.bias + st.bias_off;
              const float* eps = post ? V.eps_post : V.eps_prior;
              float* o_s = post ? V.post_s : V.prior_s;
              float* o_m = post ? V.post_m : V.prior_m;
              float* o_sd = post ? V.post_sd : V.prior_sd;
              const bool want_kl = post && V.kl != nullptr;
              float nt = 1.f;
              if ((st.flags & SF_WRITES_STATE) && V.nonterm && has_next && row_ok) nt = V.nonterm[trow + N + row];
              float kl = 0.f;
              for (int c = 0; c < S; c += 16) {
                const int nv = min(16, S - c);
                float vm[16], vs[16], e[16], pm[16], psd[16];
                const size_t o = (trow + row) * S + c;
                ld_row16(e, eps + o, nv, row_ok);
                if (want_kl) {
                  ld_row16(pm, V.prior_m + o, nv, row_ok);
                  ld_row16(psd, V.prior_sd + o, nv, row_ok);
                }
                tmem_ld16(tacc + c, vm);
                tmem_ld16(tacc + W + c, vs);
                tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                  if (i < nv) {
                    const float m = vm[i] + __ldg(bias + c + i);
                    const float sd = softplus_f(vs[i] + __ldg(bias + W + c + i)) + V.min_std;
                    const float smp = m + sd * e[i];
                    if (row_ok) {
                      o_s[o + i] = smp;
                      o_m[o + i] = m;
                      o_sd[o + i] = sd;
                      if (want_kl) {
                        const float ratio = sd / psd[i], vr = ratio * ratio;
                        const float dm = (m - pm[i]) / psd[i];
                        kl += 0.5f * (vr + dm * dm - 1.f - logf(vr));
                      }
                    }
                    if (st.flags & SF_WRITES_STATE) x_put(x_hi, x_lo, r, D + c + i, smp * nt);
                  }
                }
              }
              if (want_kl && row_ok) V.kl[trow + row] = kl;
            } else if ((st.flags & SF_LOADS_ACTION) && has_next && row_ok) {
              for (int j = 0; j < A; ++j) x_put(x_hi, x_lo, r, D + S + j, __ldg(V.actions_in + (trow + N + row) * A + j));
            }
          } break;

          case R_ACTION: {
            if (half == 0) {
              const int W = st.width;
              const float* bias = V.bias + st.bias_off;
              const float inv_ms = 1.f / V.a_mean_scale;
              for (int c = 0; c < A; c += 16) {
                const int nv = min(16, A - c);
                float vm[16], vs[16], e[16];
                const size_t o = (trow + row) * A + c;
                ld_row16(e, V.eps_action + o, nv, row_ok);
                tmem_ld16(tacc + c, vm);
                tmem_ld16(tacc + W + c, vs);
                tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                  if (i < nv) {
                    const float mean = V.a_mean_scale * tanh_f((vm[i] + __ldg(bias + c + i)) * inv_ms);
                    const float sd = softplus_f(vs[i] + __ldg(bias + W + c + i) + V.a_init_std) + V.a_min_std;
                    const float a = tanh_f(mean + sd * e[i]);
                    if (row_ok && V.actions_out) V.actions_out[o + i] = a;
                    x_put(x_hi, x_lo, r, D + S + c + i, a);
                  }
                }
              }
            }
          } break;

          case R_SCALAR: {
            if (half == 0) {
              float v[16];
              tmem_ld16(tacc, v);
              tmem_ld_wait();
              float* dst = (st.flags & SF_SCALAR_VALUE) ? V.values : V.rewards;
              if (row_ok) dst[trow + row] = v[0] + __ldg(V.bias + st.bias_off);
            }
          } break;
          default: break;
        }

        fence_proxy_async_smem();
        tc_fence_before();
        mbar_arrive(bar_act);
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
