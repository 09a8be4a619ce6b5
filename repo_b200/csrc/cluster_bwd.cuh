// Reverse-time pass of TransitionModel.observe for small batches, on the same 16-CTA cluster cut as cluster.cuh: CTA c owns
// features / units [16c, 16c+16) and (c < nS8) state dimensions [8c, 8c+8), its slices of the TRANSPOSED weights stay in shared
// memory for the whole sequence, and every product dx = dy W runs on the tensor cores (dy is the 16-row A operand, M = 128
// with aliased row groups; the weight slice is the B operand).  Same outputs as observe_bwd_kernel (bwd.cuh): the gradient of
// every pre-activation per (t, b); the weight gradients stay the host's batched GEMMs.
//
// A reverse step is four exchanges (bulk copies into the peers' shared memory, landing on their mbarriers):
//   1    (owners of state dimensions, elementwise) d[mean | raw std] of posterior and prior        -> DQP of every CTA
//   2/3  dh_q = (W_q2^T dq) act'(h_q),  dh_p = (W_p2^T dp) act'(h_p)                                -> DH of every CTA
//   4    dbelief = g + carried + W_q1[:, :D]^T dh_q + W_p1^T dh_p;  GRU gate gradients             -> DG of every CTA
//   6    de = (W_ih^T dgi) act'(e)  -> DE of the owners;   carried dbelief = dbelief z + W_hh^T dgh   (stays in registers)
//   7    (owners) dstate = (W_e^T de)[:S] nonterm[t]                                                 (stays in registers)
//
// Gradients are far below fp16's normal range, and the recurrence is linear in them: every row (sequence) runs in units of
// its own power of two (the largest incoming gradient of the row lands at 2^5, `observe_bwd_scale_kernel`), incoming
// gradients are multiplied on load, everything written to global memory is divided again.  The fp16 hi/lo pair then holds
// 2^-25 of the row's largest entry or 2^-22 relative, whichever is larger, over a 2^10 drift of the magnitudes either way.
#pragma once
#include "cluster.cuh"

namespace rb {

struct ClBwdGeom {
  int nK, nS8, K;
  uint32_t off_dqp, off_dh, off_dg, off_w;
  uint32_t w2, w3, w4, w6rz, w6n, w6h, w7;
  uint32_t cta_bytes, off_bar, smem_bytes;
};

__host__ __device__ inline bool clb_geometry(int D, int S, int A, int Hd, ClBwdGeom& g) {
  const int nD = (D + 15) / 16, nH = (Hd + 15) / 16;
  if (nD != nH || nD > kClSize) return false;
  if ((D & 3) || (Hd & 3) || (S & 1)) return false;   // a thread's 4 features / 2 state dimensions: one vector access
  g.nK = nD;
  g.K = nD * 16;
  g.nS8 = (S + 7) / 8;
  if (g.nS8 > g.nK) return false;
  uint32_t o = 0;
  g.off_dqp = o; o += 2u * (uint32_t)g.nS8 * kClSlab;
  g.off_dh = o;  o += 2u * (uint32_t)g.nK * kClSlab;   // DE (nK slabs) aliases its first half
  g.off_dg = o;  o += 4u * (uint32_t)g.nK * kClSlab;
  g.off_w = o;
  uint32_t w = 0;
  g.w2 = w;   w += 4u * 16u * 16u * (uint32_t)g.nS8;
  g.w3 = w;   w += 4u * 16u * 16u * (uint32_t)g.nS8;
  g.w4 = w;   w += 4u * 16u * 32u * (uint32_t)g.nK;
  g.w6rz = w; w += 4u * 32u * 32u * (uint32_t)g.nK;
  g.w6n = w;  w += 4u * 16u * 16u * (uint32_t)g.nK;
  g.w6h = w;  w += 4u * 16u * 16u * (uint32_t)g.nK;
  g.w7 = w;   w += 4u * 16u * 16u * (uint32_t)g.nK;
  g.cta_bytes = w;
  o += w;
  if (w < 2048) return false;
  g.off_bar = o; o += 16 * 8 + 16;
  g.smem_bytes = o;
  return o <= 227u * 1024u;
}

struct ClBwdParams {
  int T, N, D, S, A, Hd;
  int act, with_obs;
  float min_std;
  const uint8_t* wblob;
  const float* scales;   // (N, 2): power-of-two unit of the row, and its reciprocal
  const float *init_belief, *beliefs, *prior_sd, *post_sd, *eps_prior, *eps_post, *nonterm, *stash;
  int stash_ld;
  const float *g_beliefs, *g_prior_s, *g_prior_m, *g_prior_sd, *g_post_s, *g_post_m, *g_post_sd;
  float *d_q, *d_hq, *d_p, *d_hp, *d_gi, *d_gh, *d_e, *d_init_belief, *d_init_state;
  long long* dbg_clock;   // profiling build only: CTA 0 writes [step][32] clock64 stamps
};

struct ClBwdPackArgs {
  int D, S, A, Hd, E, with_obs;
  const float *w_e, *w_ih, *w_hh, *w_pp, *w_prior, *w_pq, *w_post;
  uint8_t* wblob;
};

// One block per (kind, cluster rank): the TRANSPOSED slices as fp16 B-operand blocks of 2N rows (hi rows, then lo rows, see
// cluster.cuh; rows = outputs of the product, k = the buffer column of the gradient operand in the DQP / DH / DG / DE order).
__global__ void __launch_bounds__(256) pack_cluster_bwd_weights_kernel(const __grid_constant__ ClBwdPackArgs a) {
  ClBwdGeom g;
  if (!clb_geometry(a.D, a.S, a.A, a.Hd, g)) return;
  const int kind = blockIdx.x, c = blockIdx.y;
  const int D = a.D, S = a.S, A = a.A, Hd = a.Hd;
  int N, K;
  uint32_t off;
  switch (kind) {
    case 0: N = 16; K = 16 * g.nS8; off = g.w2; break;
    case 1: N = 16; K = 16 * g.nS8; off = g.w3; break;
    case 2: N = 16; K = 32 * g.nK; off = g.w4; break;
    case 3: N = 32; K = 32 * g.nK; off = g.w6rz; break;
    case 4: N = 16; K = 16 * g.nK; off = g.w6n; break;
    case 5: N = 16; K = 16 * g.nK; off = g.w6h; break;
    default: N = 16; K = 16 * g.nK; off = g.w7; break;
  }
  uint8_t* dst = a.wblob + (size_t)c * g.cta_bytes + off;
  for (int idx = threadIdx.x; idx < N * K; idx += blockDim.x) {
    // k fastest would read the sources with stride (they are walked down a column): n fastest keeps the reads coalesced
    const int k = idx / N, n = idx - k * N;
    float v = 0.f;
    switch (kind) {
      case 0:
      case 1: {  // fc_state_posterior / fc_state_prior (2S, H): k = (owner, mean | raw std, dimension), n = hidden feature
        const int f = 16 * c + n, o = k >> 4, s = (k >> 3) & 1, j = 8 * o + (k & 7);
        const float* w = kind == 0 ? a.w_post : a.w_prior;
        if (w && (kind == 1 || a.with_obs) && f < Hd && j < S) v = w[(size_t)(s * S + j) * Hd + f];
      } break;
      case 2: {  // belief columns of fc_embed_belief_posterior (H, D+E) | fc_embed_belief_prior (H, D); n = belief unit
        const int u = 16 * c + n, o = k >> 5, which = (k >> 4) & 1, f = 16 * o + (k & 15);
        if (u < D && f < Hd) {
          if (which == 0) { if (a.with_obs) v = a.w_pq[(size_t)f * (D + a.E) + u]; }
          else v = a.w_pp[(size_t)f * D + u];
        }
      } break;
      case 3: {  // gates r, z of rnn.weight_ih (n < 16: embedding feature) and rnn.weight_hh (n >= 16: belief unit)
        const int col = 16 * c + (n & 15), o = k >> 5, gate = (k >> 4) & 1, i = 16 * o + (k & 15);
        const float* w = n < 16 ? a.w_ih : a.w_hh;
        if (col < D && i < D) v = w[(size_t)(gate * D + i) * D + col];
      } break;
      case 4:    // gate n of rnn.weight_ih; n = embedding feature
      case 5: {  // gate n of rnn.weight_hh; n = belief unit
        const int col = 16 * c + n, i = k;
        const float* w = kind == 4 ? a.w_ih : a.w_hh;
        if (col < D && i < D) v = w[(size_t)(2 * D + i) * D + col];
      } break;
      default: {  // state columns of fc_embed_state_action (D, S+A); n < 8 = state dimension
        const int j = 8 * c + n, f = k;
        if (n < 8 && j < S && f < D) v = a.w_e[(size_t)f * (S + A) + j];
      } break;
    }
    cl_pack_store(dst, N, n, k, v);
  }
}

// scales[b] = {2^k, 2^-k} with the largest |incoming gradient| of sequence b (all steps, all seven tensors) in [2^4, 2^5]
__global__ void __launch_bounds__(256) observe_bwd_scale_kernel(ClBwdParams P, float* scales) {
  const int b = blockIdx.x;
  float m = 0.f;
  const float* gs[7] = {P.g_beliefs, P.g_prior_s, P.g_prior_m, P.g_prior_sd, P.g_post_s, P.g_post_m, P.g_post_sd};
  for (int q = 0; q < 7; ++q) {
    const float* g = gs[q];
    if (!g) continue;
    const int F = q == 0 ? P.D : P.S;
    for (int idx = threadIdx.x; idx < P.T * F; idx += blockDim.x) {
      const int t = idx / F, f = idx - t * F;
      m = fmaxf(m, fabsf(g[((size_t)t * P.N + b) * F + f]));
    }
  }
#pragma unroll
  for (int s = 16; s > 0; s >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, s));
  __shared__ float red[8];
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int i = 1; i < 8; ++i) m = fmaxf(m, red[i]);
    float s = 1.f;
    if (m > 0.f && m < 3.0e38f) {
      int ex;
      frexpf(m, &ex);                                // m = f * 2^ex, f in [0.5, 1)
      s = ldexpf(1.f, max(-100, min(100, 5 - ex)));  // m * s in [2^4, 2^5)
    }
    scales[2 * b] = s;
    scales[2 * b + 1] = 1.f / s;
  }
}

#ifdef RB_STAGE_CLOCK
#define CLB_STAMP(slot) do { if (P.dbg_clock && blockIdx.x == 0 && lane == 0) P.dbg_clock[(size_t)k * 32 + (slot)] = clock64(); } while (0)
#else
#define CLB_STAMP(slot) do { } while (0)
#endif

enum ClBwdBar { BB_W = 0, BB_IN_23, BB_IN_4, BB_IN_6, BB_IN_7, BB_ACC_23, BB_ACC_4, BB_ACC_6, BB_ACC_7, BB_COUNT };

__global__ void __launch_bounds__(kClThreads, 1) rssm_cluster_observe_bwd_kernel(const __grid_constant__ ClBwdParams P) {
  extern __shared__ __align__(1024) uint8_t smem[];
  ClBwdGeom g;
  clb_geometry(P.D, P.S, P.A, P.Hd, g);
  uint8_t* dqp = smem + g.off_dqp;
  uint8_t* dh = smem + g.off_dh;
  uint8_t* de_buf = dh;   // lands after every reader of DH(t) is done, is read before DH(t-1) arrives (see the header)
  uint8_t* dg = smem + g.off_dg;
  uint8_t* wsm = smem + g.off_w;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + g.off_bar);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + BB_COUNT);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int c = (int)cl_ctarank();
  const int row0 = (int)(blockIdx.x / kClSize) * kClRows;
  const int T = P.T, N = P.N, D = P.D, S = P.S, Hd = P.Hd;
  const int nK = g.nK, nS8 = g.nS8;
  const bool active = c < nK, owner = c < nS8, with_obs = P.with_obs != 0;
  auto bar = [&](int i) { return smem_u32(bars + i); };

  if (tid == 0) {
    mbar_init(bar(BB_W), 1);
    mbar_init(bar(BB_IN_23), owner ? 2 : 1);   // the issuer's expect_tx (+ this CTA's own slice where it has one)
    mbar_init(bar(BB_IN_4), 2);
    mbar_init(bar(BB_IN_6), 2);
    mbar_init(bar(BB_IN_7), 2);
    for (int i = BB_ACC_23; i <= BB_ACC_7; ++i) mbar_init(bar(i), 1);
    mbar_fence_init();
  }
  if (warp == 1) {
    tmem_alloc(smem_u32(tmem_slot), 256);
    tmem_relinquish();
  }
  {
    uint4* z = reinterpret_cast<uint4*>(smem);
    for (uint32_t i = tid; i < g.off_w / 16; i += kClThreads) z[i] = make_uint4(0, 0, 0, 0);
  }
  __syncthreads();
  if (active && tid == 0) {
    mbar_arrive_expect_tx(bar(BB_W), g.cta_bytes);
    const uint8_t* src = P.wblob + (size_t)c * g.cta_bytes;
    for (uint32_t o = 0; o < g.cta_bytes; o += 32768u)
      bulk_g2s(smem_u32(wsm + o), src + o, min(32768u, g.cta_bytes - o), bar(BB_W));
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  cl_sync_all();
  tc_fence_after();
  const uint32_t tb = *tmem_slot;

  if (active && warp == 1) {
    // ================================ MMA issuer ================================
    const uint64_t a_dqp = make_smem_desc(smem_u32(dqp), 512, 128), a_dh = make_smem_desc(smem_u32(dh), 512, 128);
    const uint64_t a_dg = make_smem_desc(smem_u32(dg), 512, 128);
    auto wdesc = [&](uint32_t off, int n) { return make_smem_desc(smem_u32(wsm + off), (uint32_t)(2 * n * 16), 128); };
    const uint64_t w2 = wdesc(g.w2, 16), w3 = wdesc(g.w3, 16), w4 = wdesc(g.w4, 16), w6rz = wdesc(g.w6rz, 32);
    const uint64_t w6n = wdesc(g.w6n, 16), w6h = wdesc(g.w6h, 16), w7 = wdesc(g.w7, 16);
    constexpr uint32_t id32 = make_idesc_f16(128, 32), id64 = make_idesc_f16(128, 64);
    // D[128 x 2n] (+)= A * [W_hi; W_lo]^T over `steps` k16 slabs, one MMA each (accumulator lanes 16..31 = lo rows of A);
    // the i-th step reads A slab first + (i / inner) * outer + i % inner
    auto chain = [&](uint32_t d, uint64_t a_base, int first, int inner, int outer, uint64_t w, int n, int steps,
                     uint32_t idesc) {
      const uint64_t w_step = (uint64_t)(2 * n * 32) >> 4;
      for (int i = 0; i < steps; ++i) {
        const int slab = first + (inner == 1 ? i * outer : (i >> 1) * outer + (i & 1));
        umma_f16(d, a_base + (uint64_t)slab * (kClSlab >> 4), w, idesc, i == 0 ? 0u : 1u);
        w += w_step;
      }
    };
    auto arm_wait = [&](int b, uint32_t tx, uint32_t parity) {
      if (elect_one()) mbar_arrive_expect_tx(bar(b), tx);
      __syncwarp();
      mbar_wait(bar(b), parity);
      tc_fence_after();
    };
    const uint32_t tx_23 = (uint32_t)(nS8 - (owner ? 1 : 0)) * 2u * kClSlab;
    mbar_wait(bar(BB_W), 0);
    for (int k = 0; k < T; ++k) {
      const uint32_t ph = (uint32_t)k & 1u;
      arm_wait(BB_IN_23, tx_23, ph);
      CLB_STAMP(16);
      if (elect_one()) {
        if (with_obs) chain(tb + 0, a_dqp, 0, 1, 2, w2, 16, nS8, id32);
        chain(tb + 32, a_dqp, 1, 1, 2, w3, 16, nS8, id32);
        umma_commit(bar(BB_ACC_23));
      }
      __syncwarp();
      CLB_STAMP(17);
      arm_wait(BB_IN_4, (uint32_t)(nK - 1) * 2u * kClSlab, ph);
      CLB_STAMP(18);
      if (elect_one()) {
        chain(tb + 64, a_dh, 0, 1, 1, w4, 16, 2 * nK, id32);
        umma_commit(bar(BB_ACC_4));
      }
      __syncwarp();
      CLB_STAMP(19);
      arm_wait(BB_IN_6, (uint32_t)(nK - 1) * 4u * kClSlab, ph);
      CLB_STAMP(20);
      if (elect_one()) {
        chain(tb + 96, a_dg, 0, 2, 4, w6rz, 32, 2 * nK, id64);   // de | dbelief  <-  d r, d z
        chain(tb + 160, a_dg, 2, 1, 4, w6n, 16, nK, id32);       // de's share of W_ih[n]^T dn        (summed in the epilogue)
        chain(tb + 192, a_dg, 3, 1, 4, w6h, 16, nK, id32);       // dbelief's share of W_hh[n]^T (dn r)
        umma_commit(bar(BB_ACC_6));
      }
      __syncwarp();
      CLB_STAMP(21);
      if (owner) {
        arm_wait(BB_IN_7, (uint32_t)(nK - 1) * kClSlab, ph);
        CLB_STAMP(22);
        if (elect_one()) {
          chain(tb + 224, a_dh, 0, 1, 1, w7, 16, nK, id32);
          umma_commit(bar(BB_ACC_7));
        }
        __syncwarp();
        CLB_STAMP(23);
      }
    }
  } else if (active && (warp & 3) == 0) {
    // ================================ epilogue warps ================================
    const int e = warp >> 2, r = lane & 15, row = row0 + r;
    const bool row_ok = lane < 16 && row < N;
    const int f0 = 16 * c + 4 * e, j0 = 8 * c + 2 * e;
    const int act = P.act;
    const float sc = row_ok ? P.scales[2 * row] : 1.f, inv = row_ok ? P.scales[2 * row + 1] : 1.f;
    auto epi_sync = [] { asm volatile("bar.sync 1, 128;" ::: "memory"); };
    // lane l < nK - 1 talks to peer (c + 1 + l) % nK (every sender starts with a different receiver); the four warps share
    // the lanes: a bulk copy is issued from the warp's uniform datapath, one at a time, and occupies the sender's port for
    // bytes / ~20 cycles
    const uint32_t smem0 = smem_u32(smem);
    const int peer = (c + 1 + (lane & 15)) % nK;
    const uint32_t peer0 = cl_mapa(smem0, (uint32_t)peer);
    auto send = [&](uint32_t local_addr, uint32_t bytes, int npeers, int b) {
      if (lane < nK - 1 && peer < npeers && (lane & 3) == e)
        bulk_s2c(peer0 + (local_addr - smem0), local_addr, bytes, peer0 + (bar(b) - smem0));
    };
    auto ldq = [&](const float* p, size_t o, float mul) { return p ? __ldg(p + o) * mul : 0.f; };
    float db[4] = {0.f, 0.f, 0.f, 0.f}, ds[2] = {0.f, 0.f};
    // what stage 1 reads from global memory is the first thing a reverse step needs: it is requested one step ahead
    struct HeadIn { float gqs[2], gqm[2], gqsd[2], eq[2], sdq[2], gps[2], gpm[2], gpsd[2], ep[2], sdp[2], nt; };
    auto load_heads = [&](int t) {
      HeadIn h;
      h.nt = 1.f;
#pragma unroll
      for (int i = 0; i < 2; ++i) h.gqs[i] = h.gqm[i] = h.gqsd[i] = h.eq[i] = h.gps[i] = h.gpm[i] = h.gpsd[i] = h.ep[i] = 0.f, h.sdq[i] = h.sdp[i] = 1.f;
      if (owner && row_ok && t >= 0 && j0 < S) {   // S is even: both dimensions exist or neither
        const size_t tr = (size_t)t * N + row, o = tr * S + j0;
        if (P.nonterm) h.nt = __ldg(P.nonterm + tr);
        if (with_obs) {
          if (P.g_post_s) ldg2(h.gqs, P.g_post_s + o);
          if (P.g_post_m) ldg2(h.gqm, P.g_post_m + o);
          if (P.g_post_sd) ldg2(h.gqsd, P.g_post_sd + o);
          ldg2(h.eq, P.eps_post + o);
          ldg2(h.sdq, P.post_sd + o);
        }
        if (P.g_prior_s) ldg2(h.gps, P.g_prior_s + o);
        if (P.g_prior_m) ldg2(h.gpm, P.g_prior_m + o);
        if (P.g_prior_sd) ldg2(h.gpsd, P.g_prior_sd + o);
        ldg2(h.ep, P.eps_prior + o);
        ldg2(h.sdp, P.prior_sd + o);
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          h.gqs[i] *= sc; h.gqm[i] *= sc; h.gqsd[i] *= sc; h.gps[i] *= sc; h.gpm[i] *= sc; h.gpsd[i] *= sc;
        }
      } else if (owner && row_ok && t >= 0 && P.nonterm) {
        h.nt = __ldg(P.nonterm + (size_t)t * N + row);
      }
      return h;
    };
    HeadIn hn = load_heads(T - 1);
    for (int k = 0; k < T; ++k) {
      const int t = T - 1 - k;
      const uint32_t ph = (uint32_t)k & 1u;
      const size_t tr = (size_t)t * N + row;
      if (e == 0) CLB_STAMP(0);
      // ---- everything this step reads from global memory is independent of the recurrence: requested up front ----
      // the stashed activations a stage needs are requested one stage ahead (short live ranges: 128 registers per thread)
      const float* st = P.stash + tr * P.stash_ld;
      float she[4], sr[4], sz[4], sn[4], shn[4], shp[4], shq[4], bprev[4], gg[4];
      const HeadIn hd = hn;
      const float nt = hd.nt;
      // ---- 1: Gaussian heads (elementwise, owners of state dimensions) -> DQP slice [dq mean | dq raw | dp mean | dp raw]
      if (owner) {
        float dqm[2] = {0.f, 0.f}, dqr[2] = {0.f, 0.f}, dpm[2] = {0.f, 0.f}, dpr[2] = {0.f, 0.f};
        if (row_ok && j0 < S) {
#pragma unroll
          for (int i = 0; i < 2; ++i) {
            if (with_obs) {
              const float gs = hd.gqs[i] + ds[i];   // the posterior sample feeds step t+1
              dqm[i] = hd.gqm[i] + gs;
              const float dsd = hd.gqsd[i] + gs * hd.eq[i];
              dqr[i] = dsd * (1.f - __expf(-(hd.sdq[i] - P.min_std)));   // softplus' = 1 - exp(-softplus)
            }
            const float gsp = hd.gps[i] + (with_obs ? 0.f : ds[i]);
            dpm[i] = hd.gpm[i] + gsp;
            const float dsdp = hd.gpsd[i] + gsp * hd.ep[i];
            dpr[i] = dsdp * (1.f - __expf(-(hd.sdp[i] - P.min_std)));
          }
          if (with_obs) {
            stg2(P.d_q + tr * 2 * S + j0, dqm[0] * inv, dqm[1] * inv);
            stg2(P.d_q + tr * 2 * S + S + j0, dqr[0] * inv, dqr[1] * inv);
          }
          stg2(P.d_p + tr * 2 * S + j0, dpm[0] * inv, dpm[1] * inv);
          stg2(P.d_p + tr * 2 * S + S + j0, dpr[0] * inv, dpr[1] * inv);
        }
        if (lane < 16) {
          cl_put2(dqp, r, 32 * c + 2 * e, dqm[0], dqm[1]);
          cl_put2(dqp, r, 32 * c + 8 + 2 * e, dqr[0], dqr[1]);
          cl_put2(dqp, r, 32 * c + 16 + 2 * e, dpm[0], dpm[1]);
          cl_put2(dqp, r, 32 * c + 24 + 2 * e, dpr[0], dpr[1]);
        }
        fence_proxy_async_smem();
        epi_sync();
        send(smem_u32(dqp) + (uint32_t)c * 2u * kClSlab, 2u * kClSlab, nK, BB_IN_23);
        if (e == 0) {
          if (lane == 0) mbar_arrive(bar(BB_IN_23));
          CLB_STAMP(1);
        }
      }
      // (requested only now: fence.proxy.async above would wait for every load still in flight)
#pragma unroll
      for (int i = 0; i < 4; ++i) shp[i] = shq[i] = 0.f;
      if (row_ok && f0 < Hd) {
        ldg4(shp, st + 5 * D + f0);
        if (with_obs) ldg4(shq, st + 5 * D + Hd + f0);
      }
      hn = load_heads(t - 1);
      // ---- 2/3: hidden layers of the two heads -> DH slice [dh_q | dh_p]
      {
        mbar_wait(bar(BB_ACC_23), ph);
        tc_fence_after();
        if (e == 0) CLB_STAMP(2);
        float vq[4] = {0.f, 0.f, 0.f, 0.f}, lq[4] = {0.f, 0.f, 0.f, 0.f}, vp[4], lp[4];
        if (with_obs) {
          tmem_ld4(tb + 0 + 4 * e, vq);
          tmem_ld4(tb + 16 + 4 * e, lq);
        }
        tmem_ld4(tb + 32 + 4 * e, vp);
        tmem_ld4(tb + 48 + 4 * e, lp);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const bool okh = row_ok && f0 + i < Hd;
          vq[i] = cl_sum3(vq[i], lq[i]);
          vp[i] = cl_sum3(vp[i], lp[i]);
          vq[i] = (okh && with_obs) ? vq[i] * act_grad_from_output(shq[i], act) : 0.f;
          vp[i] = okh ? vp[i] * act_grad_from_output(shp[i], act) : 0.f;
          lq[i] = vq[i] * inv;
          lp[i] = vp[i] * inv;
        }
        if (row_ok && f0 < Hd) {
          if (with_obs) stg4(P.d_hq + tr * Hd + f0, lq);
          stg4(P.d_hp + tr * Hd + f0, lp);
        }
        if (lane < 16) {
          cl_put4(dh, r, 32 * c + 4 * e, vq);
          cl_put4(dh, r, 32 * c + 16 + 4 * e, vp);
        }
        fence_proxy_async_smem();
        tc_fence_before();
        epi_sync();
        send(smem_u32(dh) + (uint32_t)c * 2u * kClSlab, 2u * kClSlab, nK, BB_IN_4);
        if (e == 0) {
          if (lane == 0) mbar_arrive(bar(BB_IN_4));
          CLB_STAMP(3);
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) sr[i] = sz[i] = sn[i] = shn[i] = bprev[i] = gg[i] = 0.f;
        if (row_ok && f0 < D) {
          ldg4(sr, st + D + f0);
          ldg4(sz, st + 2 * D + f0);
          ldg4(sn, st + 3 * D + f0);
          ldg4(shn, st + 4 * D + f0);
          if (t > 0) ldg4(bprev, P.beliefs + (tr - N) * D + f0);
          else if (P.init_belief) ldg4(bprev, P.init_belief + (size_t)row * D + f0);
          if (P.g_beliefs) ldg4(gg, P.g_beliefs + tr * D + f0);
#pragma unroll
          for (int i = 0; i < 4; ++i) gg[i] *= sc;
        }
      }
      // ---- 4: dbelief_t complete -> GRU gate gradients -> DG slice [dr | dz | dn | dn r]
      float dbd[4];
      {
        mbar_wait(bar(BB_ACC_4), ph);
        tc_fence_after();
        if (e == 0) CLB_STAMP(4);
        float v[4], vl[4], drp[4], dzp[4], dnp[4], dnr[4];
        tmem_ld4(tb + 64 + 4 * e, v);
        tmem_ld4(tb + 80 + 4 * e, vl);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const bool ok = row_ok && f0 + i < D;
          const float s4 = cl_sum3(v[i], vl[i]);
          const float g_ = ok ? gg[i] + db[i] + s4 : 0.f;
          const float z = sz[i], n = sn[i], rr = sr[i];
          dnp[i] = g_ * (1.f - z) * (1.f - n * n);
          dzp[i] = g_ * (bprev[i] - n) * z * (1.f - z);
          drp[i] = dnp[i] * shn[i] * rr * (1.f - rr);
          dnr[i] = dnp[i] * rr;
          dbd[i] = g_ * z;   // direct path to belief_{t-1}; W_hh^T dgh joins in stage 6
        }
        if (row_ok && f0 < D) {
          float o1[4], o2[4], o3[4], o4[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) { o1[i] = drp[i] * inv; o2[i] = dzp[i] * inv; o3[i] = dnp[i] * inv; o4[i] = dnr[i] * inv; }
          float* o_gi = P.d_gi + tr * 3 * D + f0;
          float* o_gh = P.d_gh + tr * 3 * D + f0;
          stg4(o_gi, o1); stg4(o_gi + D, o2); stg4(o_gi + 2 * D, o3);
          stg4(o_gh, o1); stg4(o_gh + D, o2); stg4(o_gh + 2 * D, o4);
        }
        if (lane < 16) {
          cl_put4(dg, r, 64 * c + 4 * e, drp);
          cl_put4(dg, r, 64 * c + 16 + 4 * e, dzp);
          cl_put4(dg, r, 64 * c + 32 + 4 * e, dnp);
          cl_put4(dg, r, 64 * c + 48 + 4 * e, dnr);
        }
        fence_proxy_async_smem();
        tc_fence_before();
        epi_sync();
        send(smem_u32(dg) + (uint32_t)c * 4u * kClSlab, 4u * kClSlab, nK, BB_IN_6);
        if (e == 0) {
          if (lane == 0) mbar_arrive(bar(BB_IN_6));
          CLB_STAMP(5);
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) she[i] = 0.f;
        if (row_ok && f0 < D) ldg4(she, st + f0);
      }
      // ---- 6: de -> DE slab c of the owners; carried dbelief
      {
        mbar_wait(bar(BB_ACC_6), ph);
        tc_fence_after();
        if (e == 0) CLB_STAMP(6);
        float vde[4], vdb[4], lde[4], ldb[4], ne[4], nl[4], hb[4], hl[4];
        tmem_ld4(tb + 96 + 4 * e, vde);
        tmem_ld4(tb + 112 + 4 * e, vdb);
        tmem_ld4(tb + 128 + 4 * e, lde);
        tmem_ld4(tb + 144 + 4 * e, ldb);
        tmem_ld4(tb + 160 + 4 * e, ne);
        tmem_ld4(tb + 176 + 4 * e, nl);
        tmem_ld4(tb + 192 + 4 * e, hb);
        tmem_ld4(tb + 208 + 4 * e, hl);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const bool ok = row_ok && f0 + i < D;
          vde[i] = cl_sum3(vde[i], lde[i]) + cl_sum3(ne[i], nl[i]);
          vdb[i] = cl_sum3(vdb[i], ldb[i]) + cl_sum3(hb[i], hl[i]);
          vde[i] = ok ? vde[i] * act_grad_from_output(she[i], act) : 0.f;
          db[i] = ok ? dbd[i] + vdb[i] : 0.f;
          lde[i] = vde[i] * inv;
        }
        if (row_ok && f0 < D) stg4(P.d_e + tr * D + f0, lde);
        if (lane < 16) cl_put4(de_buf, r, 16 * c + 4 * e, vde);
        fence_proxy_async_smem();
        tc_fence_before();
        epi_sync();
        send(smem_u32(de_buf) + (uint32_t)c * kClSlab, kClSlab, nS8, BB_IN_7);
        if (e == 0) {
          if (owner && lane == 0) mbar_arrive(bar(BB_IN_7));
          CLB_STAMP(7);
        }
      }
      // ---- 7: the state that entered this step: s_{t-1} nonterm[t]
      if (owner) {
        mbar_wait(bar(BB_ACC_7), ph);
        tc_fence_after();
        if (e == 0) CLB_STAMP(8);
        float v[2], vl[2];
        tmem_ld2(tb + 224 + 2 * e, v);
        tmem_ld2(tb + 240 + 2 * e, vl);
        tmem_ld_wait();
        tc_fence_before();
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          const float s7 = cl_sum3(v[i], vl[i]);
          ds[i] = (row_ok && j0 + i < S) ? s7 * nt : 0.f;
        }
      }
    }
    if (row_ok) {
#pragma unroll
      for (int i = 0; i < 4; ++i)
        if (P.d_init_belief && f0 + i < D) P.d_init_belief[(size_t)row * D + f0 + i] = db[i] * inv;
      if (owner && P.d_init_state) {
#pragma unroll
        for (int i = 0; i < 2; ++i)
          if (j0 + i < S) P.d_init_state[(size_t)row * S + j0 + i] = ds[i] * inv;
      }
    }
  }

  tc_fence_before();
  cl_sync_all();   // nobody exits while a peer may still write into its shared memory
  if (warp == 1) tmem_dealloc(tb, 256);
}

}  // namespace rb
