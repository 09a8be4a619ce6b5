// The persistent "layer machine" behind observe / imagine / linear.
//
// One CTA owns NT rows (batch rows = independent RSSM sequences) for the whole time loop.
// Activations stay on-chip: X = [belief | state | action] and H = hidden, both kept as fp16
// hi/lo pairs in shared memory in the tcgen05 K-major core-matrix layout, so they are the
// B operand (N = NT rows) of every MMA.  Weights are the A operand (M = 128 output features
// per tile); they are pre-packed (pack.cuh) into 8 KB k16-slabs [hi 4 KB | lo 4 KB] already in
// the smem image, and streamed L2 -> smem by the TMA engine (1-D bulk copies) through a ring.
// Accumulators live in TMEM: tile j at columns [j*NT, (j+1)*NT), lane = output feature.
//
// Every x*W is three tcgen05.mma (hi*hi + lo*hi + hi*lo, fp32 accumulate) — fp32-grade
// accuracy on the fp16 tensor pipe (see DESIGN.md "Arithmetic").
//
// Warp roles: warp 0 = weight loader (one thread), warp 1 = MMA issuer (one thread) + TMEM
// owner, warps 2..5 = epilogue (thread <-> TMEM lane <-> output feature).  The three roles walk
// the same stage table (VmParams, passed by value), so they cannot disagree about the order.
#pragma once
#include "ptx.cuh"

namespace rb {

constexpr int kSlabBytes = 8192;  // one k16 slab of a 128-feature tile: hi 4096 B then lo 4096 B
constexpr int kSlotSlabs = 4;     // slabs per ring slot
constexpr int kSlots = 3;
constexpr int kSlotBytes = kSlotSlabs * kSlabBytes;
constexpr int kRingBytes = kSlots * kSlotBytes;
constexpr int kMaxStages = 20;
constexpr int kMaxGemms = 48;
constexpr int kThreads = 192;
constexpr int kEpiThreads = 128;
constexpr uint32_t kTmemCols = 512;

enum EpiKind : uint8_t {
  EPI_ACT_H = 0,   // H[:, f] = act(acc + bias (+ addend[t]))
  EPI_ACTION = 1,  // a = tanh(ms*tanh(m/ms) + (softplus(s+init)+min)*eps) -> X action slot
  EPI_GRU = 2,     // GRU gates -> belief' -> beliefs[t], X belief slot
  EPI_PRIOR = 3,   // prior mean/std/sample -> outputs (+ X state slot when it feeds the recurrence)
  EPI_POST = 4,    // posterior mean/std/sample + KL -> outputs, X state slot
  EPI_SCALAR = 5,  // lane 0: scalar head output (reward / value)
  EPI_STORE = 6,   // out[row, f] = acc + bias   (plain linear layer)
};
enum StageFlags : uint8_t {
  SF_WRITES_STATE = 1,  // epilogue writes the sampled state into X (masked by nonterm[t+1])
  SF_LOADS_ACTION = 2,  // epilogue stages actions_in[t+1] into X
  SF_ADDEND = 4,        // EPI_ACT_H adds addend[t, row, f]
  SF_SCALAR_VALUE = 8,  // EPI_SCALAR writes `values` instead of `rewards`
};
enum ActKind : int { ACT_RELU = 0, ACT_ELU = 1 };

struct VmGemm {        // acc[acc_tile] (+)= Wtile(128 x 16*ksl) * Src[:, 16*src_k16 ...)^T
  uint32_t w_slab;     // first slab of this tile in the packed weight blob
  uint8_t ksl;         // k16 slabs
  uint8_t src;         // 0 = X, 1 = H
  uint8_t src_k16;     // first k16 slab inside the source buffer
  uint8_t acc_tile;    // accumulator tile
  uint8_t accumulate;  // keep what the tile already holds
  uint8_t pad[3];
};
struct VmStage {
  uint8_t gemm_begin, gemm_end;
  uint8_t epi, flags;
  uint8_t ntiles;     // feature tiles the epilogue walks (per gate for GRU)
  uint8_t acc_tile0;  // first accumulator tile
  uint16_t nfeat;     // valid output features
  uint16_t bias_tile; // first 128-float bias tile
  uint8_t act;        // ActKind of EPI_ACT_H (the actor is always ELU, actor_critic.py:58)
  uint8_t pad;
  uint16_t stash_off; // EPI_ACT_H / EPI_GRU: float offset inside a (t,row) stash record, 0xFFFF = not stashed
  uint16_t pad2;
};

// Implicit-GEMM view of a (transposed) convolution (conv.cuh): GEMM row n is one output position, its columns are
// the receptive field gathered straight from the input tensor (no materialised im2col), and the store epilogue
// scatters features back onto the output grid.  Stride-2 Conv2d: sy=sx=2, dy=dx=+1.  Stride-2 ConvTranspose2d:
// stride-1 gather with dy=dx=-1 whose features are the four sub-pixel classes (shuffle = 1).
struct ConvMap {
  int enabled;
  int RA, RB;                      // per-frame row grid: n -> (frame, a, b), n = (frame*RA + a)*RB + b
  int in_nchw, C, H, W;            // input tensor layout (0 = NHWC, 1 = NCHW) and dims
  int TH, TW, tap0, ntaps;         // tap grid; this launch covers taps [tap0, tap0+ntaps), column = (tap-tap0)*C + ci
  int sy, sx, dy, dx, y0, x0;      // input pixel: iy = a*sy + ty*dy + y0, ix = b*sx + tx*dx + x0 (out of range -> 0)
  int out_nchw, Ho, Wo, osy, osx, oy0, ox0;  // output pixel (a*osy + oy0, b*osx + ox0) on an Ho x Wo grid
  int relu, accumulate;            // store: out = (accumulate ? out : 0) + acc + bias, then ReLU
  int shuffle;                     // conv.cuh only: features are (py, px, cout) sub-pixel classes of a 2x finer output grid
  int pix;                         // NHWC input: elements between neighbouring pixels (>= C; lets a GEMM read a column
                                   // window of a wider row-major matrix).  0 = C.
};

struct VmParams {
  int n_steps, n_stages;
  int N;                 // rows
  int D, S, A, Hd;       // belief, state, action, hidden sizes (A = width of X's action slot)
  int A_act;             // sampled action width; A - A_act trailing columns of the slot hold a constant per-row condition
  const float* cond;     // (N, A - A_act) or null: ConditionalTransitionModel.imagine (rssm.py:225-236)
  int kx16, kh16;        // k16 slabs held by X / H
  float min_std;         // RSSM min std
  float a_mean_scale, a_init_std, a_min_std;
  float gamma, lambda, one_minus_lambda;  // (1 - lambda) rounded from double like the reference's Python scalar
  const uint8_t* wblob;
  const float* bias;
  // initial contents of X
  const float* init_belief;  // (N, D) or null (zeros)
  const float* init_state;   // (N, S) or null
  const float* init_x;       // linear mode: (N, init_x_cols) row-major, ld = init_x_ld
  int init_x_cols, init_x_ld;
  // per-step inputs, time-major
  const float* actions_in;   // (T, N, A) or null
  const float* nonterm;      // (T, N) or null
  const float* addend;       // (T, N, Hd) or null; rows kernel only, when addend_tiled: (T, tiles, chunks of 16, 4, 128 rows, 4)
  int addend_tiled;          // see rssm_rows_kernel: the hoisted projection is stored per 128-row tile, quarter-chunk major
  const float* eps_action;   // (T, N, A)
  const float* eps_prior;    // (T, N, S)
  const float* eps_post;     // (T, N, S)
  // outputs, time-major
  float* beliefs;            // (T, N, D)
  float* prior_s; float* prior_m; float* prior_sd;  // (T, N, S)
  float* post_s; float* post_m; float* post_sd;     // (T, N, S)
  float* kl;                 // (T, N) or null
  float* actions_out;        // (T, N, A) or null
  float* rewards; float* values;  // (T, N) or null
  float* returns;            // (T-1, N) or null
  float* out; int out_ld;    // EPI_STORE
  int dbg_flags;             // bring-up only: bit0 swaps LBO/SBO in the smem descriptors
  float* stash; int stash_ld; // backward stash: (T, N, stash_ld) activations the reverse pass needs, or null
  long long* dbg_clock;      // profiling only: CTA 0 writes [step][stage][2] clock64 stamps (epilogue begin/end)
  VmStage stages[kMaxStages];
  VmGemm gemms[kMaxGemms];
};

// ------------------------------------------------------------------------------------------
// Epilogue math: MUFU-based (ex2/lg2/rcp) forms with ~1e-7 absolute error — two orders of
// magnitude inside the 1e-3 parity budget — instead of the multi-instruction libm slow paths.
__device__ __forceinline__ float ex2_f(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float rcp_f(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float exp_f(float x) { return ex2_f(x * 1.4426950408889634f); }
template <int ACT>
__device__ __forceinline__ float act_t(float x) {
  if (ACT == ACT_ELU) return x > 0.f ? x : exp_f(x) - 1.f;
  return fmaxf(x, 0.f);
}
__device__ __forceinline__ float softplus_f(float x) { return x > 20.f ? x : __logf(1.f + exp_f(x)); }
__device__ __forceinline__ float sigmoid_f(float x) { return rcp_f(1.f + exp_f(-x)); }
// 1 - 2 / (1 + e^2x) cancels for small |x| (absolute error ~2e-7 from the approximate exp / reciprocal, i.e. a relative
// error of 1e-3 at |x| = 2e-4): below 0.05 the odd Taylor polynomial takes over (next term 62/2835 x^9: < 1e-12 relative).
__device__ __forceinline__ float tanh_f(float x) {
  const float x2 = x * x;
  const float poly = x * fmaf(x2, fmaf(x2, fmaf(x2, -17.f / 315.f, 2.f / 15.f), -1.f / 3.f), 1.f);
  const float big = 1.f - 2.f * rcp_f(1.f + exp_f(2.f * x));
  return x2 < 0.0025f ? poly : big;
}
// 16 strided read-only loads issued back to back (one memory round trip, not sixteen)
__device__ __forceinline__ void ldg16(float* dst, const float* __restrict__ base, size_t stride, int row_first,
                                      int n_rows, bool lane_ok) {
#pragma unroll
  for (int i = 0; i < 16; ++i)
    dst[i] = (lane_ok && (row_first + i) < n_rows) ? __ldg(base + (size_t)i * stride) : 0.f;
}

template <int NT>
struct Tile {
  static constexpr uint32_t LBO = NT * 16 + 16;  // bytes between k-groups (8 columns); +16 spreads banks
  static constexpr uint32_t SBO = 128;           // bytes between 8-row groups
  // byte offset of element (row n, column k) inside an activation buffer
  __device__ static __forceinline__ uint32_t off(int n, int k) {
    return (uint32_t)(k >> 3) * LBO + (uint32_t)(n >> 3) * SBO + (uint32_t)(n & 7) * 16 + (uint32_t)(k & 7) * 2;
  }
  __device__ static __forceinline__ void put(uint8_t* hi, uint8_t* lo, int n, int k, float v) {
    __half h, l;
    split_f16(v, h, l);
    const uint32_t o = off(n, k);
    *reinterpret_cast<__half*>(hi + o) = h;
    *reinterpret_cast<__half*>(lo + o) = l;
  }
  __host__ __device__ static constexpr uint32_t buf_bytes(int k16) { return (uint32_t)k16 * 2u * LBO; }
};

__host__ __device__ inline size_t vm_smem_bytes(int NT, int kx16, int kh16) {
  const size_t lbo = (size_t)NT * 16 + 16;
  return (size_t)kRingBytes + 2 * (size_t)kx16 * 2 * lbo + 2 * (size_t)kh16 * 2 * lbo + 256;
}


// H[:, f] = act(acc + bias (+ addend)) for every feature tile of the stage; thread <-> feature.
template <int NT, int ACT>
__device__ __forceinline__ void epi_act_h(const VmParams& P, const VmStage& st, uint8_t* h_hi, uint8_t* h_lo,
                                          uint32_t tacc, int lf, int row0, size_t trow) {
  using TL = Tile<NT>;
  const bool addend = (st.flags & SF_ADDEND) != 0;
  const int ntiles = st.ntiles, nfeat = st.nfeat, N = P.N;
  for (int tile = 0; tile < ntiles; ++tile) {
    const int f = tile * 128 + lf;
    const bool vf = f < nfeat;
    const float bias = P.bias[(st.bias_tile + tile) * 128 + lf];
    float* stash = (P.stash && st.stash_off != 0xFFFF) ? P.stash + trow * P.stash_ld + st.stash_off : nullptr;
    uint8_t* bh = h_hi + TL::off(0, f);
    uint8_t* bl = h_lo + TL::off(0, f);
#pragma unroll 1
    for (int c = 0; c < NT / 16; ++c) {
      float v[16], ad[16];
      const int r0 = row0 + c * 16;
      if (addend) ldg16(ad, P.addend + (trow + r0) * P.Hd + f, P.Hd, r0, N, vf);
      tmem_ld16(tacc + tile * NT + c * 16, v);
      tmem_ld_wait();
      if (vf) {
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          float x = v[i] + bias;
          if (addend) x += ad[i];
          __half h, l;
          const float y = act_t<ACT>(x);
          split_f16(y, h, l);
          const uint32_t o = (uint32_t)c * 256u + (uint32_t)(i >> 3) * 128u + (uint32_t)(i & 7) * 16u;
          *reinterpret_cast<__half*>(bh + o) = h;
          *reinterpret_cast<__half*>(bl + o) = l;
          if (stash && r0 + i < N) stash[(size_t)(r0 + i) * P.stash_ld + f] = y;
        }
      }
    }
  }
}

template <int NT>
__global__ void __launch_bounds__(kThreads, 1) rssm_vm_kernel(const __grid_constant__ VmParams P) {
  using TL = Tile<NT>;
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* ring = smem;
  uint8_t* x_hi = ring + kRingBytes;
  uint8_t* x_lo = x_hi + TL::buf_bytes(P.kx16);
  uint8_t* h_hi = x_lo + TL::buf_bytes(P.kx16);
  uint8_t* h_lo = h_hi + TL::buf_bytes(P.kh16);
  uint8_t* tail = h_lo + TL::buf_bytes(P.kh16);
  uint64_t* bars = reinterpret_cast<uint64_t*>((reinterpret_cast<uintptr_t>(tail) + 15) & ~uintptr_t(15));
  // bars[0..kSlots) full, [kSlots..2kSlots) empty, [2kSlots] acc_full, [2kSlots+1] act_ready
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kSlots + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row0 = blockIdx.x * NT;
  const uint32_t bar_full = smem_u32(bars), bar_empty = smem_u32(bars + kSlots);
  const uint32_t bar_acc = smem_u32(bars + 2 * kSlots), bar_act = smem_u32(bars + 2 * kSlots + 1);

  if (threadIdx.x == 0) {
    for (int i = 0; i < kSlots; ++i) {
      mbar_init(bar_full + 8 * i, 1);
      mbar_init(bar_empty + 8 * i, 1);
    }
    mbar_init(bar_acc, 1);
    mbar_init(bar_act, kEpiThreads);
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
    // The whole warp walks the table (warp-uniform control flow); one elected lane issues.
    uint32_t slot = 0, phase = 0;
    for (int t = 0; t < P.n_steps; ++t) {
      for (int s = 0; s < P.n_stages; ++s) {
        const int g0 = P.stages[s].gemm_begin, g1 = P.stages[s].gemm_end;
        for (int g = g0; g < g1; ++g) {
          const uint32_t w_slab = P.gemms[g].w_slab;
          const int ksl = P.gemms[g].ksl;
          for (int c0 = 0; c0 < ksl; c0 += kSlotSlabs) {
            const int nsl = min(kSlotSlabs, ksl - c0);
            mbar_wait(bar_empty + 8 * slot, phase ^ 1);
            if (elect_one()) {
              const uint32_t bytes = (uint32_t)nsl * kSlabBytes;
              mbar_arrive_expect_tx(bar_full + 8 * slot, bytes);
              bulk_g2s(smem_u32(ring + slot * kSlotBytes), P.wblob + (size_t)(w_slab + c0) * kSlabBytes, bytes,
                       bar_full + 8 * slot);
            }
            __syncwarp();
            if (++slot == kSlots) { slot = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ================================ MMA issuer ================================
    // Warp-uniform loops; `elect.sync` picks the single issuing lane so the descriptors stay in
    // uniform registers (no per-instruction convergence loops around UTCHMMA).
    constexpr uint32_t idesc = make_idesc_f16(128, NT);
    uint32_t slot = 0, phase = 0, act_phase = 0;
    const bool swap = (P.dbg_flags & 1) != 0;
    // descriptors are built once; the loops below only add to their (14-bit, >>4) address field
    const uint64_t a_ring = swap ? make_smem_desc(smem_u32(ring), 128, 2048) : make_smem_desc(smem_u32(ring), 2048, 128);
    const uint64_t b_x = swap ? make_smem_desc(smem_u32(x_hi), TL::SBO, TL::LBO) : make_smem_desc(smem_u32(x_hi), TL::LBO, TL::SBO);
    const uint64_t b_h = swap ? make_smem_desc(smem_u32(h_hi), TL::SBO, TL::LBO) : make_smem_desc(smem_u32(h_hi), TL::LBO, TL::SBO);
    const uint64_t x_lo_delta = TL::buf_bytes(P.kx16) >> 4, h_lo_delta = TL::buf_bytes(P.kh16) >> 4;
    constexpr uint64_t kA_lo = 4096 >> 4, kA_slab = kSlabBytes >> 4, kA_slot = kSlotBytes >> 4;
    constexpr uint64_t kB_slab = (2u * TL::LBO) >> 4;
    for (int t = 0; t < P.n_steps; ++t) {
      for (int s = 0; s < P.n_stages; ++s) {
        const int g0 = P.stages[s].gemm_begin, g1 = P.stages[s].gemm_end;
        mbar_wait(bar_act, act_phase);  // previous epilogue: activations written, TMEM drained
        act_phase ^= 1;
        tc_fence_after();
        for (int g = g0; g < g1; ++g) {
          const VmGemm gm = P.gemms[g];
          const uint64_t lo_delta = gm.src ? h_lo_delta : x_lo_delta;
          uint64_t bd = (gm.src ? b_h : b_x) + (uint64_t)gm.src_k16 * kB_slab;
          const uint32_t d = tmem_base + (uint32_t)gm.acc_tile * NT;
          uint32_t acc = gm.accumulate;
          for (int c0 = 0; c0 < gm.ksl; c0 += kSlotSlabs) {
            const int nsl = min(kSlotSlabs, (int)gm.ksl - c0);
            mbar_wait(bar_full + 8 * slot, phase);
            tc_fence_after();
            if (elect_one()) {
              uint64_t ad = a_ring + (uint64_t)slot * kA_slot;
              uint64_t bj = bd;
#pragma unroll
              for (int j = 0; j < kSlotSlabs; ++j) {
                if (j < nsl) {
                  umma_f16(d, ad, bj, idesc, (j == 0) ? acc : 1u);
                  umma_f16(d, ad, bj + lo_delta, idesc, 1u);
                  umma_f16(d, ad + kA_lo, bj, idesc, 1u);
                  ad += kA_slab;
                  bj += kB_slab;
                }
              }
              umma_commit(bar_empty + 8 * slot);  // slot reusable once these MMAs retire
            }
            __syncwarp();
            acc = 1u;
            bd += (uint64_t)nsl * kB_slab;
            if (++slot == kSlots) { slot = 0; phase ^= 1; }
          }
        }
        if (elect_one()) umma_commit(bar_acc);  // accumulators of this stage complete
        __syncwarp();
      }
    }
  } else {
    // ================================ epilogue warps ================================
    const int et = threadIdx.x - 64;                 // 0..127
    const int q = warp & 3;                          // TMEM lane quadrant this warp may read
    const int lf = q * 32 + lane;                    // lane / feature-within-tile
    const uint32_t tlane = tmem_base + ((uint32_t)(q * 32) << 16);
    const int N = P.N, D = P.D, S = P.S, A = P.A;
    auto epi_sync = [] { asm volatile("bar.sync 1, 128;" ::: "memory"); };

    // ---- init: zero both buffers (pad columns must stay finite), then stage X ----
    {
      const uint32_t words = (2 * TL::buf_bytes(P.kx16) + 2 * TL::buf_bytes(P.kh16)) / 16;
      uint4* z = reinterpret_cast<uint4*>(x_hi);
      for (uint32_t i = et; i < words; i += kEpiThreads) z[i] = make_uint4(0, 0, 0, 0);
      epi_sync();
      if (P.init_x) {
        for (int idx = et; idx < NT * P.init_x_cols; idx += kEpiThreads) {
          const int n = idx / P.init_x_cols, k = idx - n * P.init_x_cols, row = row0 + n;
          if (row < N) TL::put(x_hi, x_lo, n, k, P.init_x[(size_t)row * P.init_x_ld + k]);
        }
      } else {
        if (P.init_belief)
          for (int idx = et; idx < NT * D; idx += kEpiThreads) {
            const int n = idx / D, k = idx - n * D, row = row0 + n;
            if (row < N) TL::put(x_hi, x_lo, n, k, P.init_belief[(size_t)row * D + k]);
          }
        if (P.init_state)
          for (int idx = et; idx < NT * S; idx += kEpiThreads) {
            const int n = idx / S, k = idx - n * S, row = row0 + n;
            if (row < N) {
              float v = P.init_state[(size_t)row * S + k];
              if (P.nonterm) v *= P.nonterm[row];
              TL::put(x_hi, x_lo, n, D + k, v);
            }
          }
        if (P.actions_in)
          for (int idx = et; idx < NT * A; idx += kEpiThreads) {
            const int n = idx / A, k = idx - n * A, row = row0 + n;
            if (row < N) TL::put(x_hi, x_lo, n, D + S + k, P.actions_in[(size_t)row * A + k]);
          }
        if (P.cond) {   // constant tail of the action slot: written once, the action epilogue only touches the first A_act columns
          const int Cn = A - P.A_act;
          for (int idx = et; idx < NT * Cn; idx += kEpiThreads) {
            const int n = idx / Cn, k = idx - n * Cn, row = row0 + n;
            if (row < N) TL::put(x_hi, x_lo, n, D + S + P.A_act + k, P.cond[(size_t)row * Cn + k]);
          }
        }
      }
      fence_proxy_async_smem();
      mbar_arrive(bar_act);
    }

    uint32_t acc_phase = 0;
    for (int t = 0; t < P.n_steps; ++t) {
      const size_t trow = (size_t)t * N;  // row offset of time step t in time-major tensors
      const bool has_next = (t + 1) < P.n_steps;
      for (int s = 0; s < P.n_stages; ++s) {
        const VmStage& st = P.stages[s];
        mbar_wait(bar_acc, acc_phase);
        acc_phase ^= 1;
        tc_fence_after();
        const uint32_t tacc = tlane + (uint32_t)st.acc_tile0 * NT;

        switch (st.epi) {
          case EPI_ACT_H: {
            if (st.act == ACT_ELU) epi_act_h<NT, ACT_ELU>(P, st, h_hi, h_lo, tacc, lf, row0, trow);
            else epi_act_h<NT, ACT_RELU>(P, st, h_hi, h_lo, tacc, lf, row0, trow);
          } break;

          case EPI_STORE: {
            for (int tile = 0; tile < st.ntiles; ++tile) {
              const int f = tile * 128 + lf;
              const bool vf = f < st.nfeat;
              const float bias = P.bias[(st.bias_tile + tile) * 128 + lf];
#pragma unroll 1
              for (int c = 0; c < NT / 16; ++c) {
                float v[16];
                tmem_ld16(tacc + tile * NT + c * 16, v);
                tmem_ld_wait();
                if (vf) {
#pragma unroll
                  for (int i = 0; i < 16; ++i) {
                    const int row = row0 + c * 16 + i;
                    if (row < N) P.out[(size_t)row * P.out_ld + f] = v[i] + bias;
                  }
                }
              }
            }
          } break;

          case EPI_GRU: {
            const int mt = st.ntiles;
            for (int tile = 0; tile < mt; ++tile) {
              const int u = tile * 128 + lf;
              const bool vu = u < D;
              const float br = P.bias[(st.bias_tile + tile) * 128 + lf];
              const float bz = P.bias[(st.bias_tile + mt + tile) * 128 + lf];
              const float bin = P.bias[(st.bias_tile + 2 * mt + tile) * 128 + lf];
              const float bhn = P.bias[(st.bias_tile + 3 * mt + tile) * 128 + lf];
              // previous belief in fp32: the start belief at t=0, else what this very thread stored at t-1
              const float* bprev = (t == 0) ? P.init_belief : (P.beliefs + (trow - N) * D);
#pragma unroll 1
              for (int c = 0; c < NT / 16; ++c) {
                float vr[16], vz[16], vi[16], vh[16], bo[16];
                const int r0 = row0 + c * 16;
                if (bprev) {
#pragma unroll
                  for (int i = 0; i < 16; ++i) bo[i] = (vu && r0 + i < N) ? bprev[(size_t)(r0 + i) * D + u] : 0.f;
                } else {
#pragma unroll
                  for (int i = 0; i < 16; ++i) bo[i] = 0.f;
                }
                tmem_ld16(tacc + (tile)*NT + c * 16, vr);
                tmem_ld16(tacc + (mt + tile) * NT + c * 16, vz);
                tmem_ld16(tacc + (2 * mt + tile) * NT + c * 16, vi);
                tmem_ld16(tacc + (3 * mt + tile) * NT + c * 16, vh);
                tmem_ld_wait();
                if (vu) {
                  float bn[16];
#pragma unroll
                  for (int i = 0; i < 16; ++i) {
                    const float r = sigmoid_f(vr[i] + br);
                    const float z = sigmoid_f(vz[i] + bz);
                    const float hn = vh[i] + bhn;
                    const float nn = tanh_f(vi[i] + bin + r * hn);
                    bn[i] = (1.f - z) * nn + z * bo[i];
                    TL::put(x_hi, x_lo, c * 16 + i, u, bn[i]);
                    if (P.stash && st.stash_off != 0xFFFF && r0 + i < N) {
                      float* sp = P.stash + (trow + r0 + i) * P.stash_ld + st.stash_off + u;
                      sp[0] = r; sp[D] = z; sp[2 * D] = nn; sp[3 * D] = hn;
                    }
                  }
#pragma unroll
                  for (int i = 0; i < 16; ++i)
                    if (r0 + i < N) P.beliefs[(trow + r0 + i) * D + u] = bn[i];
                }
              }
            }
          } break;

          case EPI_PRIOR:
          case EPI_POST: {
            const bool post = st.epi == EPI_POST;
            const int j = lf;
            const bool vj = j < S;
            const float bm = P.bias[(st.bias_tile) * 128 + lf];
            const float bs = P.bias[(st.bias_tile + 1) * 128 + lf];
            const float* eps = post ? P.eps_post : P.eps_prior;
            float* o_s = post ? P.post_s : P.prior_s;
            float* o_m = post ? P.post_m : P.prior_m;
            float* o_sd = post ? P.post_sd : P.prior_sd;
            const bool any_valid_in_warp = (q * 32) < S;
            const bool want_kl = post && P.kl != nullptr;
            const bool mask_next = (st.flags & SF_WRITES_STATE) && P.nonterm && has_next;
#pragma unroll 1
            for (int c = 0; c < NT / 16; ++c) {
              float vm[16], vs[16], e[16], pm[16], psd[16], nt[16];
              const int r0 = row0 + c * 16;
              const size_t o0 = (trow + r0) * S + j;
              if (any_valid_in_warp) {
                ldg16(e, eps + o0, S, r0, N, vj);
                if (want_kl) {  // written by this same thread in EPI_PRIOR of this step
#pragma unroll
                  for (int i = 0; i < 16; ++i) {
                    const bool ok = vj && r0 + i < N;
                    pm[i] = ok ? P.prior_m[o0 + (size_t)i * S] : 0.f;
                    psd[i] = ok ? P.prior_sd[o0 + (size_t)i * S] : 1.f;
                  }
                }
                if (mask_next) ldg16(nt, P.nonterm + trow + N + r0, 1, r0, N, true);
              }
              tmem_ld16(tacc + c * 16, vm);
              tmem_ld16(tacc + NT + c * 16, vs);
              tmem_ld_wait();
              if (any_valid_in_warp) {
                float smp[16], m[16], sd[16];
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                  m[i] = vm[i] + bm;
                  sd[i] = softplus_f(vs[i] + bs) + P.min_std;
                  smp[i] = m[i] + sd[i] * e[i];
                }
                if (want_kl) {
#pragma unroll
                  for (int i = 0; i < 16; ++i) {
                    float klj = 0.f;
                    if (vj && r0 + i < N) {
                      const float ratio = sd[i] / psd[i], vr = ratio * ratio;
                      const float dm = (m[i] - pm[i]) / psd[i];
                      klj = 0.5f * (vr + dm * dm - 1.f - logf(vr));
                    }
#pragma unroll
                    for (int sh = 16; sh > 0; sh >>= 1) klj += __shfl_xor_sync(0xffffffffu, klj, sh);
                    if (lane == 0 && r0 + i < N) {
                      if (S <= 32) P.kl[trow + r0 + i] = klj;
                      else atomicAdd(&P.kl[trow + r0 + i], klj);
                    }
                  }
                }
                if (vj) {
#pragma unroll
                  for (int i = 0; i < 16; ++i) {
                    if (r0 + i < N) {
                      const size_t o = o0 + (size_t)i * S;
                      o_s[o] = smp[i];
                      o_m[o] = m[i];
                      o_sd[o] = sd[i];
                    }
                  }
                  if (st.flags & SF_WRITES_STATE) {
#pragma unroll
                    for (int i = 0; i < 16; ++i)
                      TL::put(x_hi, x_lo, c * 16 + i, D + j, mask_next ? smp[i] * nt[i] : smp[i]);
                  }
                }
              }
            }
            if ((st.flags & SF_LOADS_ACTION) && has_next) {
              for (int idx = et; idx < NT * A; idx += kEpiThreads) {
                const int n = idx / A, k = idx - n * A, row = row0 + n;
                if (row < N) TL::put(x_hi, x_lo, n, D + S + k, __ldg(P.actions_in + (trow + N + row) * A + k));
              }
            }
          } break;

          case EPI_ACTION: {
            const int j = lf;
            const int Aa = P.A_act;   // sampled action width (== A unless a condition rides in the slot)
            const bool vj = j < Aa;
            const float bm = P.bias[(st.bias_tile) * 128 + lf];
            const float bs = P.bias[(st.bias_tile + 1) * 128 + lf];
            const bool any_valid_in_warp = (q * 32) < Aa;
            const float inv_ms = 1.f / P.a_mean_scale;
#pragma unroll 1
            for (int c = 0; c < NT / 16; ++c) {
              float vm[16], vs[16], e[16];
              const int r0 = row0 + c * 16;
              const size_t o0 = (trow + r0) * Aa + j;
              if (any_valid_in_warp) ldg16(e, P.eps_action + o0, Aa, r0, N, vj);
              tmem_ld16(tacc + c * 16, vm);
              tmem_ld16(tacc + NT + c * 16, vs);
              tmem_ld_wait();
              if (any_valid_in_warp && vj) {
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                  const float mean = P.a_mean_scale * tanh_f((vm[i] + bm) * inv_ms);
                  const float sd = softplus_f(vs[i] + bs + P.a_init_std) + P.a_min_std;
                  const float a = tanh_f(mean + sd * e[i]);
                  if (r0 + i < N && P.actions_out) P.actions_out[o0 + (size_t)i * Aa] = a;
                  if (P.stash && st.stash_off != 0xFFFF && r0 + i < N) {
                    float* sp = P.stash + (trow + r0 + i) * P.stash_ld + st.stash_off + j;
                    sp[0] = mean; sp[A] = sd;
                  }
                  TL::put(x_hi, x_lo, c * 16 + i, D + S + j, a);
                }
              }
            }
          } break;

          case EPI_SCALAR: {
            const float b0 = P.bias[(st.bias_tile) * 128 + lf];
            float* dst = (st.flags & SF_SCALAR_VALUE) ? P.values : P.rewards;
#pragma unroll 1
            for (int c = 0; c < NT / 16; ++c) {
              float v[16];
              tmem_ld16(tacc + c * 16, v);
              tmem_ld_wait();
              if (lf == 0) {
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                  const int row = row0 + c * 16 + i;
                  if (row < N) dst[trow + row] = v[i] + b0;
                }
              }
            }
          } break;
          default: break;
        }

        // hand the buffers / TMEM back to the MMA issuer
        fence_proxy_async_smem();
        tc_fence_before();
        mbar_arrive(bar_act);
      }
    }

    // ---- lambda-return: reverse scan over the horizon, one thread per row ----
    if (P.returns && P.rewards && P.values && P.n_steps >= 2) {
      __threadfence_block();
      epi_sync();
      const int row = row0 + et;
      if (et < NT && row < N) {
        const int T = P.n_steps;
        const float g = P.gamma, lam = P.lambda;
        float last = P.values[(size_t)(T - 1) * N + row];  // bootstrap = values[-1]
        float next_v = last;
        for (int t = T - 2; t >= 0; --t) {
          const float r = P.rewards[(size_t)t * N + row];
          const float inp = r + g * next_v * P.one_minus_lambda;
          last = inp + g * lam * last;
          P.returns[(size_t)t * N + row] = last;
          next_v = P.values[(size_t)t * N + row];
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, kTmemCols);
}

}  // namespace rb
