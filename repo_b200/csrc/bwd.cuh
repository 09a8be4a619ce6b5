// Reverse-time pass of TransitionModel.observe (BPTT; the reference gets it from autograd over
// rssm.py:116-133).  Small-batch design: one CTA per sequence (batch column), the whole time loop
// on-chip, fp32.  Every product here is dx = W^T dy, i.e. thread k walks column k of the caller's
// row-major W — consecutive threads read consecutive addresses, so the weights (L2-resident, 2.2 MB)
// need no transposed copy.  The kernel emits the gradient of every pre-activation per (t,b); the
// weight gradients are then plain batched GEMMs over those tensors (host side, repo_b200/autograd.py).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace rb {

struct ObsBwdParams {
  int T, B, D, S, A, Hd, E;
  int act, with_obs;
  float min_std;
  // weights: caller's fp32 tensors, row-major [out, in]
  const float *w_e, *w_ih, *w_hh, *w_p1, *w_p2, *w_q1, *w_q2;
  // forward tensors
  const float *init_belief;          // (B, D) or null (zeros)
  const float *beliefs;              // (T, B, D)
  const float *prior_sd, *post_sd;   // (T, B, S)
  const float *eps_prior, *eps_post; // (T, B, S)
  const float *nonterm;              // (T, B) or null
  const float *stash; int stash_ld;  // (T, B, ld): [e D][r D][z D][n D][h_n D][hp Hd][hq Hd]
  // incoming gradients, each (T, B, feature) or null
  const float *g_beliefs, *g_prior_s, *g_prior_m, *g_prior_sd, *g_post_s, *g_post_m, *g_post_sd;
  // outputs: gradients of the pre-activations, time-major
  float *d_q;   // (T, B, 2S)  posterior [mean | raw std]
  float *d_hq;  // (T, B, Hd)  posterior hidden pre-activation
  float *d_p;   // (T, B, 2S)
  float *d_hp;  // (T, B, Hd)
  float *d_gi;  // (T, B, 3D)  W_ih x + b_ih
  float *d_gh;  // (T, B, 3D)  W_hh h + b_hh
  float *d_e;   // (T, B, D)   fc_embed_state_action pre-activation
  float *d_init_belief, *d_init_state;  // (B, D), (B, S) or null
};

__device__ __forceinline__ float act_grad_from_output(float y, int act) {
  // derivative of the activation expressed through its OUTPUT: relu' = [y > 0]; elu' = y > 0 ? 1 : y + 1
  if (act == 1) return y > 0.f ? 1.f : y + 1.f;
  return y > 0.f ? 1.f : 0.f;
}

// out[k] = sum_j W[j*ld + k] * dy[j], j in [j0, j1)  (dy in shared memory, W column walk is coalesced over k).
// The kernel is bound by the LATENCY of these L2 reads, not by their bandwidth (one CTA per sequence: 1.4 MB of weights per
// time step at 18 B/clk/SM with 8 loads in flight per thread), so 16 independent loads are issued before the first FMA.
__device__ __forceinline__ float col_dot(const float* __restrict__ W, int ld, int k, const float* dy, int j0, int j1) {
  float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
  int j = j0;
  for (; j + 16 <= j1; j += 16) {
    float w[16];
#pragma unroll
    for (int u = 0; u < 16; ++u) w[u] = __ldg(W + (size_t)(j + u) * ld + k);
#pragma unroll
    for (int u = 0; u < 16; u += 4) {
      a0 = fmaf(w[u], dy[j + u], a0);
      a1 = fmaf(w[u + 1], dy[j + u + 1], a1);
      a2 = fmaf(w[u + 2], dy[j + u + 2], a2);
      a3 = fmaf(w[u + 3], dy[j + u + 3], a3);
    }
  }
  for (; j + 4 <= j1; j += 4) {
    const float w0 = __ldg(W + (size_t)j * ld + k), w1 = __ldg(W + (size_t)(j + 1) * ld + k);
    const float w2 = __ldg(W + (size_t)(j + 2) * ld + k), w3 = __ldg(W + (size_t)(j + 3) * ld + k);
    a0 = fmaf(w0, dy[j], a0);
    a1 = fmaf(w1, dy[j + 1], a1);
    a2 = fmaf(w2, dy[j + 2], a2);
    a3 = fmaf(w3, dy[j + 3], a3);
  }
  for (; j < j1; ++j) a0 = fmaf(__ldg(W + (size_t)j * ld + k), dy[j], a0);
  return (a0 + a1) + (a2 + a3);
}

constexpr int kObsGroups = 4;  // thread groups of 256 splitting the reduction range (loads in flight x4)

// Every thread calls this; group g walks its quarter of j for column k, partials meet in shared memory.
// Returns the full sum to group 0 (other groups get garbage).  Contains two __syncthreads.
__device__ __forceinline__ float col_dot_split(const float* __restrict__ W, int ld, int k, int kmax, const float* dy,
                                               int n, float* part, int g) {
  const int per = (n + kObsGroups - 1) / kObsGroups;
  const int j0 = min(n, g * per), j1 = min(n, j0 + per);
  part[g * 256 + k] = (k < kmax) ? col_dot(W, ld, k, dy, j0, j1) : 0.f;
  __syncthreads();
  float v = 0.f;
  if (g == 0) {
#pragma unroll
    for (int i = 0; i < kObsGroups; ++i) v += part[i * 256 + k];
  }
  __syncthreads();
  return v;
}

__global__ void __launch_bounds__(256 * kObsGroups) observe_bwd_kernel(const __grid_constant__ ObsBwdParams P) {
  extern __shared__ float sm[];
  const int D = P.D, S = P.S, A = P.A, Hd = P.Hd, T = P.T, B = P.B;
  float* db = sm;            // D   recurrent dL/d belief_{t}
  float* ds = db + D;        // S   recurrent dL/d state_{t} (the state fed into step t+1)
  float* dq = ds + S;        // 2S
  float* dp = dq + 2 * S;    // 2S
  float* dh = dp + 2 * S;    // Hd  hidden pre-activation gradient (posterior, then prior)
  float* dba = dh + Hd;      // D   accumulated dL/d belief_t
  float* dgi = dba + D;      // 3D
  float* dgh = dgi + 3 * D;  // 3D
  float* de = dgh + 3 * D;   // D
  float* part = de + D;      // kObsGroups * 256 partial sums
  const int b = blockIdx.x, k = threadIdx.x & 255, g = threadIdx.x >> 8;
  const bool lead = g == 0;
  if (lead && k < D) db[k] = 0.f;
  if (lead && k < S) ds[k] = 0.f;
  __syncthreads();

  for (int t = T - 1; t >= 0; --t) {
    const size_t tb = (size_t)t * B + b;
    const float* st = P.stash + tb * P.stash_ld;
    // ---- Gaussian heads: gradient of [mean | raw std] ----
    if (lead && k < S) {
      const int j = k;
      const size_t o = tb * S + j;
      if (P.with_obs) {
        const float gs = (P.g_post_s ? P.g_post_s[o] : 0.f) + ds[j];  // the posterior sample feeds step t+1
        const float dmu = (P.g_post_m ? P.g_post_m[o] : 0.f) + gs;
        const float dsd = (P.g_post_sd ? P.g_post_sd[o] : 0.f) + gs * P.eps_post[o];
        const float draw = dsd * (1.f - __expf(-(P.post_sd[o] - P.min_std)));  // softplus' = 1 - exp(-softplus)
        dq[j] = dmu; dq[S + j] = draw;
        P.d_q[tb * 2 * S + j] = dmu; P.d_q[tb * 2 * S + S + j] = draw;
      }
      const float gsp = (P.g_prior_s ? P.g_prior_s[o] : 0.f) + (P.with_obs ? 0.f : ds[j]);
      const float dmup = (P.g_prior_m ? P.g_prior_m[o] : 0.f) + gsp;
      const float dsdp = (P.g_prior_sd ? P.g_prior_sd[o] : 0.f) + gsp * P.eps_prior[o];
      const float drawp = dsdp * (1.f - __expf(-(P.prior_sd[o] - P.min_std)));
      dp[j] = dmup; dp[S + j] = drawp;
      P.d_p[tb * 2 * S + j] = dmup; P.d_p[tb * 2 * S + S + j] = drawp;
    }
    if (lead && k < D) dba[k] = (P.g_beliefs ? P.g_beliefs[tb * D + k] : 0.f) + db[k];
    __syncthreads();
    // ---- posterior hidden layer ----
    if (P.with_obs) {
      float v = col_dot_split(P.w_q2, Hd, k, Hd, dq, 2 * S, part, g);
      if (lead && k < Hd) {
        v *= act_grad_from_output(st[5 * D + Hd + k], P.act);
        dh[k] = v;
        P.d_hq[tb * Hd + k] = v;
      }
      __syncthreads();
      v = col_dot_split(P.w_q1, D + P.E, k, D, dh, Hd, part, g);
      if (lead && k < D) dba[k] += v;
      __syncthreads();
    }
    // ---- prior hidden layer ----
    {
      float v = col_dot_split(P.w_p2, Hd, k, Hd, dp, 2 * S, part, g);
      if (lead && k < Hd) {
        v *= act_grad_from_output(st[5 * D + k], P.act);
        dh[k] = v;
        P.d_hp[tb * Hd + k] = v;
      }
      __syncthreads();
      v = col_dot_split(P.w_p1, D, k, D, dh, Hd, part, g);
      if (lead && k < D) dba[k] += v;
      __syncthreads();
    }
    // ---- GRU cell ----
    if (lead && k < D) {
      const int i = k;
      const float r = st[D + i], z = st[2 * D + i], n = st[3 * D + i], hn = st[4 * D + i];
      const float bprev = t > 0 ? P.beliefs[(tb - B) * D + i] : (P.init_belief ? P.init_belief[(size_t)b * D + i] : 0.f);
      const float gg = dba[i];
      const float dnp = gg * (1.f - z) * (1.f - n * n);
      const float dzp = gg * (bprev - n) * z * (1.f - z);
      const float drp = dnp * hn * r * (1.f - r);
      dgi[i] = drp; dgi[D + i] = dzp; dgi[2 * D + i] = dnp;
      dgh[i] = drp; dgh[D + i] = dzp; dgh[2 * D + i] = dnp * r;
      db[i] = gg * z;  // direct path to belief_{t-1}; W_hh^T dgh is added below
      float* o_gi = P.d_gi + tb * 3 * D;
      float* o_gh = P.d_gh + tb * 3 * D;
      o_gi[i] = drp; o_gi[D + i] = dzp; o_gi[2 * D + i] = dnp;
      o_gh[i] = drp; o_gh[D + i] = dzp; o_gh[2 * D + i] = dnp * r;
    }
    __syncthreads();
    {
      float v = col_dot_split(P.w_ih, D, k, D, dgi, 3 * D, part, g);
      if (lead && k < D) {
        v *= act_grad_from_output(st[k], P.act);
        de[k] = v;
        P.d_e[tb * D + k] = v;
      }
      v = col_dot_split(P.w_hh, D, k, D, dgh, 3 * D, part, g);
      if (lead && k < D) db[k] += v;
      __syncthreads();
    }
    // ---- state that entered this step: s_{t-1} * nonterm[t] ----
    {
      const float nt = P.nonterm ? P.nonterm[tb] : 1.f;
      const float v = col_dot_split(P.w_e, S + A, k, S, de, D, part, g);
      if (lead && k < S) ds[k] = v * nt;
      __syncthreads();
    }
  }
  if (lead && P.d_init_belief && k < D) P.d_init_belief[(size_t)b * D + k] = db[k];
  if (lead && P.d_init_state && k < S) P.d_init_state[(size_t)b * S + k] = ds[k];
}

}  // namespace rb

namespace rb {

// =====================================================================================================
// Reverse-time pass of TransitionModel.imagine (rssm.py:167-176) incl. the tanh-Normal actor
// (actor_critic.py:76-102).  RB rows per CTA share every weight read (the rollout has thousands of
// rows, so L2 traffic — not latency — is what matters here).  The actor's INPUTS are detached in the
// reference (rssm.py:170): gradients flow  outputs -> dynamics -> action -> actor parameters  and
// outputs -> dynamics -> earlier (belief, state), never through the actor's inputs.
// =====================================================================================================
struct ImgBwdParams {
  int A_act;   // sampled action width (A = slot width incl. a trailing condition)
  int T, N, D, S, A, Hd;
  int act;
  float min_std, a_mean_scale, a_min_std;
  const float *w_e, *w_ih, *w_hh, *w_p1, *w_p2;   // transition weights
  const float *w_a2, *w_a3, *w_a4, *w_a5;         // actor fc2..fc5 (fc1 only receives d1 as a weight gradient)
  const float *start_belief;                      // (N, D)
  const float *beliefs, *actions;                 // (T, N, D), (T, N, A)
  const float *prior_sd, *eps_prior, *eps_action; // (T, N, S), (T, N, S), (T, N, A)
  const float *stash; int stash_ld;               // [e D][r D][z D][n D][h_n D][hp H][h1..h4 4H][mean A][std A]
  const float *g_beliefs, *g_prior_s, *g_prior_m, *g_prior_sd;  // incoming gradients (nullable)
  float *d_p, *d_hp, *d_gi, *d_gh, *d_e;          // transition pre-activation gradients
  float *d_a5, *d_a4, *d_a3, *d_a2, *d_a1;        // actor pre-activation gradients: (T,N,2A), 4 x (T,N,H)
  float *d_start_belief, *d_start_state;          // (N, D), (N, S) or null
};

template <int RB>
__device__ __forceinline__ void col_dot_rows(const float* __restrict__ W, int ld, int k, const float* dy, int n,
                                             float (&acc)[RB]) {
#pragma unroll
  for (int r = 0; r < RB; ++r) acc[r] = 0.f;
  int j = 0;
  for (; j + 8 <= n; j += 8) {   // 8 independent weight loads in flight per thread
    float w[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) w[u] = __ldg(W + (size_t)(j + u) * ld + k);
#pragma unroll
    for (int u = 0; u < 8; ++u) {
#pragma unroll
      for (int r = 0; r < RB; ++r) acc[r] = fmaf(w[u], dy[(j + u) * RB + r], acc[r]);
    }
  }
  for (; j < n; ++j) {
    const float w = __ldg(W + (size_t)j * ld + k);
#pragma unroll
    for (int r = 0; r < RB; ++r) acc[r] = fmaf(w, dy[j * RB + r], acc[r]);
  }
}

// Up to 12 rows per CTA two CTAs share an SM (<= 128 registers, 2 x 110 KB of shared memory): the kernel's time is set by
// the instruction stream of a CTA (0.55 ms + 0.077 ms per row it carries, measured at 4 / 8 / 12 / 20 rows), and two
// warps per scheduler leave issue slots empty that a second CTA fills.
template <int RB>
__global__ void __launch_bounds__(256, RB <= 12 ? 2 : 1) imagine_bwd_kernel(const __grid_constant__ ImgBwdParams P) {
  extern __shared__ float sm[];
  // A = width of the [action | condition] slot that enters the embedding layer, Aa = sampled action width (Aa == A
  // unless the model is conditional: rssm.py:225-236; the condition columns carry no gradient)
  const int D = P.D, S = P.S, A = P.A, Aa = P.A_act, Hd = P.Hd, T = P.T, N = P.N;
  // all vectors are [feature][RB] so one weight element meets RB rows with a vector LDS
  float* db = sm;                  // D
  float* ds = db + D * RB;         // S
  float* dp = ds + S * RB;         // 2S
  float* dh = dp + 2 * S * RB;     // Hd (prior hidden, then actor hidden ping)
  float* dh2 = dh + Hd * RB;       // Hd (actor hidden pong)
  float* dba = dh2 + Hd * RB;      // D
  float* dgi = dba + D * RB;       // 3D
  float* dgh = dgi + 3 * D * RB;   // 3D
  float* de = dgh + 3 * D * RB;    // D
  float* d5 = de + D * RB;         // 2A
  const int k = threadIdx.x, nth = blockDim.x;
  const int row0 = blockIdx.x * RB;
  for (int i = k; i < D * RB; i += nth) db[i] = 0.f;
  for (int i = k; i < S * RB; i += nth) ds[i] = 0.f;
  __syncthreads();

  for (int t = T - 1; t >= 0; --t) {
    // ---- prior head ----
    for (int idx = k; idx < S * RB; idx += nth) {
      const int j = idx / RB, r = idx - j * RB, row = row0 + r;
      float dmu = 0.f, draw = 0.f;
      if (row < N) {
        const size_t o = ((size_t)t * N + row) * S + j;
        const float gs = (P.g_prior_s ? P.g_prior_s[o] : 0.f) + ds[idx];
        dmu = (P.g_prior_m ? P.g_prior_m[o] : 0.f) + gs;
        const float dsd = (P.g_prior_sd ? P.g_prior_sd[o] : 0.f) + gs * P.eps_prior[o];
        draw = dsd * (1.f - __expf(-(P.prior_sd[o] - P.min_std)));
        P.d_p[((size_t)t * N + row) * 2 * S + j] = dmu;
        P.d_p[((size_t)t * N + row) * 2 * S + S + j] = draw;
      }
      dp[j * RB + r] = dmu;
      dp[(S + j) * RB + r] = draw;
    }
    for (int idx = k; idx < D * RB; idx += nth) {
      const int i = idx / RB, r = idx - i * RB, row = row0 + r;
      dba[idx] = db[idx] + ((row < N && P.g_beliefs) ? P.g_beliefs[((size_t)t * N + row) * D + i] : 0.f);
    }
    __syncthreads();
    for (int i = k; i < Hd; i += nth) {
      float acc[RB];
      col_dot_rows<RB>(P.w_p2, Hd, i, dp, 2 * S, acc);
#pragma unroll
      for (int r = 0; r < RB; ++r) {
        const int row = row0 + r;
        float g = 0.f;
        if (row < N) {
          const size_t tr = (size_t)t * N + row;
          g = acc[r] * act_grad_from_output(P.stash[tr * P.stash_ld + 5 * D + i], P.act);
          P.d_hp[tr * Hd + i] = g;
        }
        dh[i * RB + r] = g;
      }
    }
    __syncthreads();
    for (int i = k; i < D; i += nth) {
      float acc[RB];
      col_dot_rows<RB>(P.w_p1, D, i, dh, Hd, acc);
      // ---- GRU cell (same thread owns unit i of every row) ----
#pragma unroll
      for (int r = 0; r < RB; ++r) {
        const int row = row0 + r;
        float drp = 0.f, dzp = 0.f, dnp = 0.f, dnr = 0.f, dbz = 0.f;
        if (row < N) {
          const size_t tr = (size_t)t * N + row;
          const float* st = P.stash + tr * P.stash_ld;
          const float rr = st[D + i], z = st[2 * D + i], n = st[3 * D + i], hn = st[4 * D + i];
          const float bprev = t > 0 ? P.beliefs[(tr - N) * D + i] : P.start_belief[(size_t)row * D + i];
          const float g = dba[i * RB + r] + acc[r];
          dnp = g * (1.f - z) * (1.f - n * n);
          dzp = g * (bprev - n) * z * (1.f - z);
          drp = dnp * hn * rr * (1.f - rr);
          dnr = dnp * rr;
          dbz = g * z;
          float* o_gi = P.d_gi + tr * 3 * D;
          float* o_gh = P.d_gh + tr * 3 * D;
          o_gi[i] = drp; o_gi[D + i] = dzp; o_gi[2 * D + i] = dnp;
          o_gh[i] = drp; o_gh[D + i] = dzp; o_gh[2 * D + i] = dnr;
        }
        dgi[i * RB + r] = drp; dgi[(D + i) * RB + r] = dzp; dgi[(2 * D + i) * RB + r] = dnp;
        dgh[i * RB + r] = drp; dgh[(D + i) * RB + r] = dzp; dgh[(2 * D + i) * RB + r] = dnr;
        db[i * RB + r] = dbz;
      }
    }
    __syncthreads();
    for (int i = k; i < D; i += nth) {
      float acc[RB], acc2[RB];
      col_dot_rows<RB>(P.w_ih, D, i, dgi, 3 * D, acc);
      col_dot_rows<RB>(P.w_hh, D, i, dgh, 3 * D, acc2);
#pragma unroll
      for (int r = 0; r < RB; ++r) {
        const int row = row0 + r;
        float g = 0.f;
        if (row < N) {
          const size_t tr = (size_t)t * N + row;
          g = acc[r] * act_grad_from_output(P.stash[tr * P.stash_ld + i], P.act);
          P.d_e[tr * D + i] = g;
        }
        de[i * RB + r] = g;
        db[i * RB + r] += acc2[r];
      }
    }
    __syncthreads();
    // ---- [state | action] that entered this step ----
    for (int j = k; j < S + A; j += nth) {
      float acc[RB];
      col_dot_rows<RB>(P.w_e, S + A, j, de, D, acc);
#pragma unroll
      for (int r = 0; r < RB; ++r) {
        const int row = row0 + r;
        if (j < S) ds[j * RB + r] = acc[r];
        else if (j - S < Aa) {
          const int a = j - S;
          float dm = 0.f, dsr = 0.f;
          if (row < N) {
            const size_t tr = (size_t)t * N + row;
            const float av = P.actions[tr * Aa + a];
            const float du = acc[r] * (1.f - av * av);           // a = tanh(u)
            const float* st = P.stash + tr * P.stash_ld + 5 * D + 5 * Hd;
            const float mean = st[a], sd = st[A + a];
            const float mm = mean / P.a_mean_scale;
            dm = du * (1.f - mm * mm);                            // mean = ms * tanh(m / ms)
            dsr = du * P.eps_action[tr * Aa + a] * (1.f - __expf(-(sd - P.a_min_std)));
            P.d_a5[tr * 2 * Aa + a] = dm;
            P.d_a5[tr * 2 * Aa + Aa + a] = dsr;
          }
          d5[a * RB + r] = dm;
          d5[(Aa + a) * RB + r] = dsr;
        }
      }
    }
    __syncthreads();
    // ---- actor chain fc5 -> fc2 (ELU everywhere); its inputs are detached, so it stops at d1.  Nothing in the recurrence
    // reads it (rssm.py:170): with d_a4 == NULL the caller runs it AFTER the time loop as dense GEMMs over all (t, row). ----
    if (P.d_a4 == nullptr) continue;
    const float* Wk[4] = {P.w_a5, P.w_a4, P.w_a3, P.w_a2};
    float* outk[4] = {P.d_a4, P.d_a3, P.d_a2, P.d_a1};
    const float* src = d5;
    int nsrc = 2 * Aa;
    float* dst = dh;
    for (int l = 0; l < 4; ++l) {
      const int hoff = 5 * D + Hd + (3 - l) * Hd;  // h4, h3, h2, h1
      for (int i = k; i < Hd; i += nth) {
        float acc[RB];
        col_dot_rows<RB>(Wk[l], Hd, i, src, nsrc, acc);
#pragma unroll
        for (int r = 0; r < RB; ++r) {
          const int row = row0 + r;
          float g = 0.f;
          if (row < N) {
            const size_t tr = (size_t)t * N + row;
            g = acc[r] * act_grad_from_output(P.stash[tr * P.stash_ld + hoff + i], 1);
            outk[l][tr * Hd + i] = g;
          }
          dst[i * RB + r] = g;
        }
      }
      __syncthreads();
      src = dst;
      nsrc = Hd;
      dst = (dst == dh) ? dh2 : dh;
    }
  }
  for (int idx = k; idx < D * RB; idx += nth) {
    const int i = idx / RB, r = idx - i * RB, row = row0 + r;
    if (P.d_start_belief && row < N) P.d_start_belief[(size_t)row * D + i] = db[idx];
  }
  for (int idx = k; idx < S * RB; idx += nth) {
    const int j = idx / RB, r = idx - j * RB, row = row0 + r;
    if (P.d_start_state && row < N) P.d_start_state[(size_t)row * S + j] = ds[idx];
  }
}

}  // namespace rb

namespace rb {

// =====================================================================================================
// Backward of an L-layer MLP on [belief | state] (RewardModel / ValueModel: L=4, out 1; ActorModel.forward:
// L=5, out 2A).  RB rows per CTA.  Emits the pre-activation gradient of every layer (for the weight-gradient
// GEMMs) and, if wanted, the gradient of the input rows.
// =====================================================================================================
struct MlpBwdParams {
  int N, in_f, Hd, out_f, L;          // L layers: in_f -> Hd x (L-1) -> out_f
  int act;
  const float* w[5];                  // fc1..fcL, row-major [out, in]
  const float* stash; int stash_ld;   // (N, (L-1)*Hd): post-activation h1..h_{L-1}
  const float* g_out;                 // (N, out_f)
  float* d_h[4];                      // pre-activation gradients of fc1..fc_{L-1}: each (N, Hd)
  float* d_x;                         // (N, in_f) or null
};

template <int RB>
__global__ void __launch_bounds__(256) mlp_bwd_kernel(const __grid_constant__ MlpBwdParams P) {
  extern __shared__ float sm[];
  const int Hd = P.Hd, N = P.N, L = P.L;
  float* ga = sm;                   // max(out_f, Hd) x RB
  float* gb = ga + max(P.out_f, Hd) * RB;
  const int k = threadIdx.x, nth = blockDim.x;
  const int row0 = blockIdx.x * RB;
  for (int idx = k; idx < P.out_f * RB; idx += nth) {
    const int j = idx / RB, r = idx - j * RB, row = row0 + r;
    ga[idx] = row < N ? P.g_out[(size_t)row * P.out_f + j] : 0.f;
  }
  __syncthreads();
  const float* src = ga;
  float* dst = gb;
  int nsrc = P.out_f;
  for (int l = L - 1; l >= 1; --l) {   // d_{l} = W_{l+1}^T d_{l+1} * act'(h_l)
    for (int i = k; i < Hd; i += nth) {
      float acc[RB];
      col_dot_rows<RB>(P.w[l], Hd, i, src, nsrc, acc);
#pragma unroll
      for (int r = 0; r < RB; ++r) {
        const int row = row0 + r;
        float g = 0.f;
        if (row < N) {
          g = acc[r] * act_grad_from_output(P.stash[(size_t)row * P.stash_ld + (l - 1) * Hd + i], P.act);
          P.d_h[l - 1][(size_t)row * Hd + i] = g;
        }
        dst[i * RB + r] = g;
      }
    }
    __syncthreads();
    const float* tmp = src;
    src = dst;
    dst = const_cast<float*>(tmp);
    nsrc = Hd;
  }
  if (P.d_x) {
    for (int i = k; i < P.in_f; i += nth) {
      float acc[RB];
      col_dot_rows<RB>(P.w[0], P.in_f, i, src, Hd, acc);
#pragma unroll
      for (int r = 0; r < RB; ++r)
        if (row0 + r < N) P.d_x[(size_t)(row0 + r) * P.in_f + i] = acc[r];
    }
  }
}

// =====================================================================================================
// Backward of the MC tanh-Normal entropy (elementwise.cuh): d entropy[m] / d mean[m,a], d std[m,a].
//   x = atanh(clamp(tanh(u))), u = mean + std*eps;  dx/du = (1 - y^2) / (1 - yc^2) inside the clamp, else 0
//   log p = -(x-mu)^2/(2 s^2) - log s - c + 2x + 2 softplus(-2x) - 2 log 2     (d/dx of the tail = 2 tanh(x))
// =====================================================================================================
__global__ void __launch_bounds__(256) tanh_normal_entropy_bwd_kernel(const float* __restrict__ mean,
                                                                      const float* __restrict__ std_,
                                                                      const float* __restrict__ eps,
                                                                      const float* __restrict__ g_ent,
                                                                      float* __restrict__ d_mean, float* __restrict__ d_std,
                                                                      int M, int A, int K) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;  // one thread per (m, a)
  if (idx >= M * A) return;
  const int m = idx / A, a = idx - m * A;
  const float mu = mean[idx], sd = std_[idx];
  const float inv_var = 1.f / (sd * sd);
  float gm = 0.f, gs = 0.f;
  for (int k = 0; k < K; ++k) {
    const float e = __ldg(eps + ((size_t)k * M + m) * A + a);
    const float y = tanhf(mu + sd * e);
    const bool inside = fabsf(y) <= 0.99999997f;
    const float yc = fabsf(y) <= 1.f ? fminf(fmaxf(y, -0.99999997f), 0.99999997f) : y;
    const float x = atanhf(yc);
    const float dxdu = inside ? (1.f - y * y) / (1.f - yc * yc) : 0.f;
    const float d = x - mu;
    const float dlp_dx = -d * inv_var + 2.f * tanhf(x);
    gm += dlp_dx * dxdu + d * inv_var;                          // d log p / d mu
    gs += dlp_dx * dxdu * e + d * d * inv_var / sd - 1.f / sd;  // d log p / d std
  }
  const float w = -g_ent[m] / (float)K;
  d_mean[idx] = w * gm;
  d_std[idx] = w * gs;
}

}  // namespace rb
