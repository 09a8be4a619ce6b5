// Host side of librepo_b200.so: builds the stage table ("program") for observe / imagine /
// linear from the model sizes, packs the caller's fp32 weights into the workspace, and launches
// the layer machine (vm.cuh).  C-ABI declared in include/repo_b200.h.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cmath>
#include <cstring>
#include <vector>

#include "../../include/repo_b200.h"
#include "pack.cuh"
#include "vm.cuh"
#include "rows.cuh"
#include "cluster.cuh"
#include "conv.cuh"
#include "elementwise.cuh"
#include "bwd.cuh"
#include "cluster_bwd.cuh"
#include "optim.cuh"

using namespace rb;

namespace {

thread_local char g_err[512] = "";

int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

#define CUDA_OK(expr)                                                                        \
  do {                                                                                       \
    cudaError_t e_ = (expr);                                                                 \
    if (e_ != cudaSuccess) return fail(-2, "%s failed: %s", #expr, cudaGetErrorString(e_));  \
  } while (0)

inline int cdiv(int a, int b) { return (a + b - 1) / b; }

// Builds the stage/gemm tables and the matching pack jobs in lock-step.
struct Builder {
  VmParams P;
  PackArgs pack;
  BiasArgs bias;
  int n_gemms = 0, n_stages = 0, n_bias_tiles = 0;
  uint32_t n_slabs = 0;
  int blk = 0;
  int max_acc_tiles = 0;
  bool overflow = false;

  Builder() {
    std::memset(&P, 0, sizeof(P));
    std::memset(&pack, 0, sizeof(pack));
    std::memset(&bias, 0, sizeof(bias));
  }

  // one 128-row tile of a weight matrix -> one gemm + one pack job
  void gemm_tile(const float* w, int ld, int row0, int nrows, int col0, int ncols, int kofs, int ksl, int src,
                 int src_k16, int acc_tile, int accumulate) {
    if (n_gemms >= kMaxGemms || pack.n_jobs >= kMaxPackJobs) { overflow = true; return; }
    VmGemm& g = P.gemms[n_gemms++];
    g.w_slab = n_slabs;
    g.ksl = (uint8_t)ksl;
    g.src = (uint8_t)src;
    g.src_k16 = (uint8_t)src_k16;
    g.acc_tile = (uint8_t)acc_tile;
    g.accumulate = (uint8_t)accumulate;
    PackJob& j = pack.jobs[pack.n_jobs++];
    j.w = w; j.ld = ld; j.row0 = row0; j.nrows = nrows; j.col0 = col0; j.ncols = ncols;
    j.kofs = kofs; j.ksl = ksl; j.w_slab = n_slabs; j.blk0 = blk;
    n_slabs += ksl;
    blk += ksl;
    max_acc_tiles = std::max(max_acc_tiles, acc_tile + 1);
  }
  // all tiles of rows [row0, row0+nrows) of a matrix; accumulator tiles acc0, acc0+1, ...
  void gemm_rows(const float* w, int ld, int row0, int nrows, int col0, int ncols, int kofs, int ksl, int src,
                 int src_k16, int acc0, int accumulate) {
    for (int t = 0; t * 128 < nrows; ++t)
      gemm_tile(w, ld, row0 + t * 128, std::min(128, nrows - t * 128), col0, ncols, kofs, ksl, src, src_k16,
                acc0 + t, accumulate);
  }
  void bias_tile(const float* a, int a_off, const float* b, int b_off, int n) {
    if (bias.n_jobs >= kMaxPackJobs) { overflow = true; return; }
    BiasJob& j = bias.jobs[bias.n_jobs++];
    j.a = a; j.a_off = a_off; j.b = b; j.b_off = b_off; j.n = std::max(0, std::min(128, n));
    j.dst_tile = n_bias_tiles++;
  }
  void bias_rows(const float* a, int a_off, const float* b, int b_off, int n) {
    for (int t = 0; t * 128 < n; ++t) bias_tile(a, a_off + t * 128, b, b ? b_off + t * 128 : 0, n - t * 128);
  }
  VmStage& begin_stage() {
    VmStage& s = P.stages[std::min(n_stages, kMaxStages - 1)];
    if (n_stages >= kMaxStages) overflow = true;
    s.gemm_begin = (uint8_t)n_gemms;
    s.bias_tile = (uint16_t)n_bias_tiles;
    s.stash_off = 0xFFFF;
    return s;
  }
  void end_stage(VmStage& s, int epi, int flags, int ntiles, int acc_tile0, int nfeat, int act) {
    s.gemm_end = (uint8_t)n_gemms;
    s.epi = (uint8_t)epi;
    s.flags = (uint8_t)flags;
    s.ntiles = (uint8_t)ntiles;
    s.acc_tile0 = (uint8_t)acc_tile0;
    s.nfeat = (uint16_t)nfeat;
    s.act = (uint8_t)act;
    ++n_stages;
  }

  // dense layer with activation: H = act(W src + b)
  void dense_to_h(const float* w, const float* b, int ld, int out_f, int col0, int ncols, int kofs, int ksl, int src,
                  int src_k16, int act, int flags = 0, int stash_off = 0xFFFF) {
    VmStage& s = begin_stage();
    gemm_rows(w, ld, 0, out_f, col0, ncols, kofs, ksl, src, src_k16, 0, 0);
    bias_rows(b, 0, nullptr, 0, out_f);
    s.stash_off = (uint16_t)stash_off;
    end_stage(s, EPI_ACT_H, flags, cdiv(out_f, 128), 0, out_f, act);
  }
  // Gaussian head: rows [0,n) = mean -> tile 0, rows [n,2n) = raw std -> tile 1
  void gaussian_head(const float* w, const float* b, int ld, int n, int ksl, int epi, int flags, int stash_off = 0xFFFF) {
    VmStage& s = begin_stage();
    s.stash_off = (uint16_t)stash_off;
    gemm_tile(w, ld, 0, n, 0, ld, 0, ksl, 1, 0, 0, 0);
    gemm_tile(w, ld, n, n, 0, ld, 0, ksl, 1, 0, 1, 0);
    bias_tile(b, 0, nullptr, 0, n);
    bias_tile(b, n, nullptr, 0, n);
    end_stage(s, epi, flags, 1, 0, n, 0);
  }
  // fc_embed_state_action + GRUCell  (rssm.py:34-40)
  // stash record per (t,row): [e D][r D][z D][n D][h_n D][prior hidden Hd][posterior hidden Hd]
  void belief_update(const repo_b200_rssm_weights* W, int D, int S, int A, int act, bool stash = false) {
    const int kD16 = cdiv(D, 16), kx16 = cdiv(D + S + A, 16), kSA0 = D / 16, mtD = cdiv(D, 128);
    dense_to_h(W->fc_embed_state_action_w, W->fc_embed_state_action_b, S + A, D, 0, S + A, D - 16 * kSA0,
               kx16 - kSA0, 0, kSA0, act, 0, stash ? 0 : 0xFFFF);
    VmStage& s = begin_stage();
    if (stash) s.stash_off = (uint16_t)D;
    for (int g = 0; g < 3; ++g)  // W_ih * hidden: r, z, i_n
      gemm_rows(W->rnn_w_ih, D, g * D, D, 0, D, 0, kD16, 1, 0, g * mtD, 0);
    for (int g = 0; g < 3; ++g)  // W_hh * belief: r, z accumulate; h_n separate
      gemm_rows(W->rnn_w_hh, D, g * D, D, 0, D, 0, kD16, 0, 0, (g < 2 ? g : 3) * mtD, g < 2 ? 1 : 0);
    bias_rows(W->rnn_b_ih, 0, W->rnn_b_hh, 0, D);
    bias_rows(W->rnn_b_ih, D, W->rnn_b_hh, D, D);
    bias_rows(W->rnn_b_ih, 2 * D, nullptr, 0, D);
    bias_rows(W->rnn_b_hh, 2 * D, nullptr, 0, D);
    end_stage(s, EPI_GRU, 0, mtD, 0, D, 0);
  }
  // 4-layer scalar head on [belief|state]  (decoder.py:189-195 / actor_critic.py:20-26)
  void scalar_head(const repo_b200_mlp_weights* M, int D, int S, int Hd, int act, int flags) {
    const int kBS16 = cdiv(D + S, 16), kH16 = cdiv(Hd, 16);
    dense_to_h(M->w[0], M->b[0], D + S, Hd, 0, D + S, 0, kBS16, 0, 0, act);
    dense_to_h(M->w[1], M->b[1], Hd, Hd, 0, Hd, 0, kH16, 1, 0, act);
    dense_to_h(M->w[2], M->b[2], Hd, Hd, 0, Hd, 0, kH16, 1, 0, act);
    VmStage& s = begin_stage();
    gemm_tile(M->w[3], Hd, 0, 1, 0, Hd, 0, kH16, 1, 0, 0, 0);
    bias_tile(M->b[3], 0, nullptr, 0, 1);
    end_stage(s, EPI_SCALAR, flags, 1, 0, 1, 0);
  }

  size_t wblob_bytes() const { return (size_t)n_slabs * kSlabBytes; }
  size_t bias_bytes() const { return (size_t)n_bias_tiles * 128 * sizeof(float); }
  size_t packed_bytes() const { return wblob_bytes() + bias_bytes(); }

  // point the program at a workspace and (optionally) enqueue the packing kernels
  int bind_and_pack(void* ws, size_t ws_bytes, bool do_pack, cudaStream_t st) {
    if (overflow) return fail(-3, "program too large (stages %d gemms %d)", n_stages, n_gemms);
    if (ws_bytes < packed_bytes()) return fail(-4, "workspace too small: need %zu bytes, got %zu", packed_bytes(), ws_bytes);
    if ((reinterpret_cast<uintptr_t>(ws) & 15) != 0) return fail(-4, "workspace must be 16-byte aligned");
    uint8_t* wb = static_cast<uint8_t*>(ws);
    float* bb = reinterpret_cast<float*>(wb + wblob_bytes());
    P.wblob = wb;
    P.bias = bb;
    P.n_stages = n_stages;
    if (do_pack) {
      pack.wblob = wb;
      bias.bias = bb;
      pack_weights_kernel<<<blk, 128, 0, st>>>(pack);
      pack_bias_kernel<<<bias.n_jobs, 128, 0, st>>>(bias);
      CUDA_OK(cudaGetLastError());
    }
    return 0;
  }
};

int g_dbg_flags = 0;
long long* g_dbg_clock = nullptr;
int g_sm_count = 0;
int sm_count() {
  if (g_sm_count == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 0;
    cudaDeviceGetAttribute(&g_sm_count, cudaDevAttrMultiProcessorCount, dev);
  }
  return g_sm_count;
}

template <int NT>
int launch_nt(const VmParams& P, cudaStream_t st) {
  const size_t smem = vm_smem_bytes(NT, P.kx16, P.kh16);
  if (smem > 227 * 1024) return fail(-5, "shared memory budget exceeded: %zu bytes at row tile %d", smem, NT);
  static size_t configured = 0;
  if (smem > configured) {
    CUDA_OK(cudaFuncSetAttribute(rssm_vm_kernel<NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = smem;
  }
  const int grid = cdiv(P.N, NT);
  VmParams Q = P;
  Q.dbg_flags = g_dbg_flags;
  rssm_vm_kernel<NT><<<grid, kThreads, smem, st>>>(Q);
  CUDA_OK(cudaGetLastError());
  return 0;
}

// rows per CTA for the vm kernel.  Measured per-step cost of one CTA is ~(60 + NT) us (the MMA phase is bound by
// re-reading the 128-feature weight tiles, the epilogue scales with the rows), and CTAs run in waves of one per
// SM — so pick the tile that minimises waves x per-CTA cost (2450 rows: 32-row tiles = 77 CTAs in one wave beat
// 16-row tiles = 154 CTAs in two).
int pick_row_tile(int rows, int max_acc_tiles, int kx16, int kh16, int requested) {
  int nt = 64;
  if (requested == 16 || requested == 32 || requested == 64) nt = requested;
  else {
    const int sms = std::max(1, sm_count());
    double best = 1e30;
    for (int cand : {64, 32, 16}) {
      const int waves = cdiv(cdiv(rows, cand), sms);
      const double cost = waves * (60.0 + cand);
      if (cost < best - 1e-9) { best = cost; nt = cand; }
    }
  }
  while (nt > 16 && (max_acc_tiles * nt > (int)kTmemCols || vm_smem_bytes(nt, kx16, kh16) > 227 * 1024)) nt >>= 1;
  return nt;
}

int launch(const VmParams& P, int max_acc_tiles, int requested, cudaStream_t st) {
  if (P.N <= 0 || P.n_steps <= 0) return 0;
  if (requested == 128) requested = 64;  // shapes the 128-row kernel does not take
  const int nt = pick_row_tile(P.N, max_acc_tiles, P.kx16, P.kh16, requested);
  if (max_acc_tiles * nt > (int)kTmemCols) return fail(-5, "TMEM budget exceeded (%d accumulator tiles x %d rows)", max_acc_tiles, nt);
  switch (nt) {
    case 16: return launch_nt<16>(P, st);
    case 32: return launch_nt<32>(P, st);
    default: return launch_nt<64>(P, st);
  }
}

int check_dims(const repo_b200_dims* d) {
  if (!d) return fail(-1, "dims is NULL");
  if (d->belief < 1 || d->belief > 256) return fail(-1, "belief_size %d unsupported (1..256)", d->belief);
  if (d->hidden < 1 || d->hidden > 256) return fail(-1, "hidden_size %d unsupported (1..256)", d->hidden);
  if (d->state < 1 || d->state > 128) return fail(-1, "state_size %d unsupported (1..128)", d->state);
  if (d->action < 1 || d->action > 128) return fail(-1, "action_size %d unsupported (1..128)", d->action);
  if (d->belief + d->state + d->action > 255 * 16) return fail(-1, "input width too large");
  return 0;
}
int check_act(int act) {
  if (act != REPO_B200_ACT_RELU && act != REPO_B200_ACT_ELU)
    return fail(-1, "activation kind %d unsupported (relu=0, elu=1)", act);
  return 0;
}

void set_dims(VmParams& P, const repo_b200_dims* d) {
  P.D = d->belief; P.S = d->state; P.A = d->action; P.Hd = d->hidden;
  P.A_act = d->action;
  P.kx16 = cdiv(d->belief + d->state + d->action, 16);
  P.kh16 = cdiv(std::max(d->belief, d->hidden), 16);
}

// imagine stash record per (t,row): [e D][r D][z D][n D][h_n D][prior hidden H][actor h1..h4 4H][action mean A][action std A]
void build_imagine(Builder& b, const repo_b200_dims* d, const repo_b200_rssm_weights* W,
                   const repo_b200_mlp_weights* actor, const repo_b200_mlp_weights* reward,
                   const repo_b200_mlp_weights* value, int act, bool stash = false, int cond = 0) {
  // cond > 0: ConditionalTransitionModel / ConditionalActorModel (rssm.py:187-248, actor_critic.py:105-148).  X's action slot is
  // [sampled action (A - cond) | condition (cond)]: the embedding layer reads all of it ([state | action | condition], the
  // reference's concat order), the actor reads [belief | state] and, as a second accumulating window, the condition columns.
  const int D = d->belief, S = d->state, A = d->action, Hd = d->hidden, Aa = A - cond;
  const int kD16 = cdiv(D, 16), kBS16 = cdiv(D + S, 16), kH16 = cdiv(Hd, 16);
  const int so = 5 * D + Hd;  // actor block of the stash
  set_dims(b.P, d);
  b.P.A_act = Aa;
  // actor (always ELU: actor_critic.py:58 + the positional-arg quirk at dreamer.py:99-105)
  if (cond > 0) {
    const int kc0 = (D + S + Aa) / 16, kx16 = cdiv(D + S + A, 16);
    VmStage& s = b.begin_stage();
    b.gemm_rows(actor->w[0], D + S + cond, 0, Hd, 0, D + S, 0, kBS16, 0, 0, 0, 0);
    b.gemm_rows(actor->w[0], D + S + cond, 0, Hd, D + S, cond, (D + S + Aa) - 16 * kc0, kx16 - kc0, 0, kc0, 0, 1);
    b.bias_rows(actor->b[0], 0, nullptr, 0, Hd);
    s.stash_off = (uint16_t)(stash ? so : 0xFFFF);
    b.end_stage(s, EPI_ACT_H, 0, cdiv(Hd, 128), 0, Hd, ACT_ELU);
  } else {
    b.dense_to_h(actor->w[0], actor->b[0], D + S, Hd, 0, D + S, 0, kBS16, 0, 0, ACT_ELU, 0, stash ? so : 0xFFFF);
  }
  for (int i = 1; i < 4; ++i)
    b.dense_to_h(actor->w[i], actor->b[i], Hd, Hd, 0, Hd, 0, kH16, 1, 0, ACT_ELU, 0, stash ? so + i * Hd : 0xFFFF);
  b.gaussian_head(actor->w[4], actor->b[4], Hd, Aa, kH16, EPI_ACTION, 0, stash ? so + 4 * Hd : 0xFFFF);
  b.belief_update(W, D, S, A, act, stash);
  b.dense_to_h(W->fc_embed_belief_prior_w, W->fc_embed_belief_prior_b, D, Hd, 0, D, 0, kD16, 0, 0, act, 0,
               stash ? 5 * D : 0xFFFF);
  b.gaussian_head(W->fc_state_prior_w, W->fc_state_prior_b, Hd, S, kH16, EPI_PRIOR, SF_WRITES_STATE);
  if (reward) b.scalar_head(reward, D, S, Hd, act, 0);
  if (value) b.scalar_head(value, D, S, Hd, act, SF_SCALAR_VALUE);
}

void build_observe(Builder& b, const repo_b200_dims* d, const repo_b200_rssm_weights* W, bool with_obs, int act,
                   bool stash = false) {
  const int D = d->belief, S = d->state, A = d->action, Hd = d->hidden, E = d->embed;
  const int kD16 = cdiv(D, 16), kH16 = cdiv(Hd, 16);
  set_dims(b.P, d);
  b.belief_update(W, D, S, A, act, stash);
  b.dense_to_h(W->fc_embed_belief_prior_w, W->fc_embed_belief_prior_b, D, Hd, 0, D, 0, kD16, 0, 0, act, 0,
               stash ? 5 * D : 0xFFFF);
  b.gaussian_head(W->fc_state_prior_w, W->fc_state_prior_b, Hd, S, kH16, EPI_PRIOR,
                  with_obs ? 0 : (SF_WRITES_STATE | SF_LOADS_ACTION));
  if (with_obs) {
    // belief half of fc_embed_belief_posterior; the embedding half was hoisted into `addend`
    b.dense_to_h(W->fc_embed_belief_posterior_w, W->fc_embed_belief_posterior_b, D + E, Hd, 0, D, 0, kD16, 0, 0, act,
                 SF_ADDEND, stash ? 5 * D + Hd : 0xFFFF);
    b.gaussian_head(W->fc_state_posterior_w, W->fc_state_posterior_b, Hd, S, kH16, EPI_POST,
                    SF_WRITES_STATE | SF_LOADS_ACTION);
  }
}

void build_linear(Builder& b, const float* w, const float* bias, int ld, int col0, int in_f, int out_f) {
  b.P.kx16 = cdiv(in_f, 16);
  b.P.kh16 = 0;
  VmStage& s = b.begin_stage();
  b.gemm_rows(w, ld, 0, out_f, col0, in_f, 0, cdiv(in_f, 16), 0, 0, 0, 0);
  b.bias_rows(bias, 0, nullptr, 0, out_f);
  b.end_stage(s, EPI_STORE, 0, cdiv(out_f, 128), 0, out_f, 0);
}

size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// ============================================================================================
// "rows on M" machine (rows.cuh): program builder
// ============================================================================================
inline int r16(int v) { return (v + 15) / 16 * 16; }

struct RBuilder {
  RowsParams P;
  PackRowsArgs pack;
  BiasRowsArgs bias;
  int n_gemms = 0, n_stages = 0, n_bias = 0, blk = 0;
  int cur_h = 0;        // TMEM region (0: columns [0,256), 1: [256,512)) holding the latest hidden activations
  size_t w_bytes = 0;
  bool overflow = false;

  RBuilder() {
    std::memset(&P, 0, sizeof(P));
    std::memset(&pack, 0, sizeof(pack));
    std::memset(&bias, 0, sizeof(bias));
  }
  struct Seg { int src, n, dst; };

  void gemm(const float* w, int ld, std::initializer_list<Seg> segs, int n_pad, int col0, int ncols, int kofs, int ksl,
            int a_src, int a_k16, int acc_col, int accumulate, int init_cols = 0) {
    if (n_gemms >= kRMaxGemms || pack.n_jobs >= kMaxRowsJobs || n_pad > 256 || n_pad * 64 > kRSlotBytes) { overflow = true; return; }
    RGemm& g = P.gemms[n_gemms++];
    g.w_off16 = (uint32_t)(w_bytes / 16);
    g.slab_bytes16 = (uint16_t)(n_pad * 4);
    g.n = (uint16_t)n_pad;
    g.ksl = (uint8_t)ksl;
    g.a_src = (uint8_t)a_src;
    g.a_k16 = (uint8_t)a_k16;
    g.accumulate = (uint8_t)accumulate;
    g.acc_col = (uint16_t)acc_col;
    g.init_cols = (uint16_t)init_cols;
    PackRowsJob& j = pack.jobs[pack.n_jobs++];
    j.w = w; j.ld = ld; j.col0 = col0; j.ncols = ncols; j.kofs = kofs; j.ksl = ksl; j.n_pad = n_pad;
    j.nseg = 0;
    for (const Seg& s : segs) { j.seg_src[j.nseg] = s.src; j.seg_n[j.nseg] = s.n; j.seg_dst[j.nseg] = s.dst; ++j.nseg; }
    j.dst_off16 = g.w_off16;
    j.blk0 = blk;
    blk += ksl;
    w_bytes += (size_t)ksl * n_pad * 64;
  }
  void bias_job(const float* a, int a_off, const float* b, int b_off, int n, int n_pad) {
    if (bias.n_jobs >= 64) { overflow = true; return; }
    BiasRowsJob& j = bias.jobs[bias.n_jobs++];
    j.a = a; j.b = b; j.a_off = a_off; j.b_off = b_off; j.n = n; j.n_pad = n_pad; j.dst_off = n_bias;
    n_bias += n_pad;
  }
  RStage& begin_stage() {
    RStage& s = P.stages[std::min(n_stages, kRMaxStages - 1)];
    if (n_stages >= kRMaxStages) overflow = true;
    s.gemm_begin = (uint8_t)n_gemms;
    s.bias_off = (uint16_t)n_bias;
    return s;
  }
  int xdeps[kRMaxStages] = {};
  void end_stage(RStage& s, int epi, int flags, int nfeat, int act, int unit0 = 0, int width = 0, int xdep = 1) {
    s.gemm_end = (uint8_t)n_gemms;
    s.bias_n = (uint16_t)(n_bias - s.bias_off);
    if (n_bias > kBiasCap) overflow = true;
    s.epi = (uint8_t)epi; s.flags = (uint8_t)flags; s.act = (uint8_t)act;
    s.nfeat = (uint16_t)nfeat; s.unit0 = (uint16_t)unit0; s.width = (uint16_t)width;
    s.rounds = (uint8_t)(epi == R_ACT_H ? cdiv(cdiv(nfeat, 16), kEpiParts) : 1);
    xdeps[std::min(n_stages, kRMaxStages - 1)] = xdep;
    // accumulators go to the region that does NOT hold the current H; an R_ACT_H epilogue converts them in place,
    // so that region then holds H and the other one is free for the next layer's accumulators.  An early stage
    // (xdep > 1) runs under the previous stage's epilogue, which is still reading the OTHER region: it takes the region
    // of the dead H operand, and its own H then lives there (no flip).
    if (xdep > 1 && epi == R_ACT_H) {
      s.regs = (uint8_t)(cur_h | (cur_h << 1));
    } else {
      s.regs = (uint8_t)((1 - cur_h) | (cur_h << 1));
      if (epi == R_ACT_H) cur_h = 1 - cur_h;
    }
    ++n_stages;
  }
  // One hidden layer = ONE full-width GEMM (N = r16(out_f)): tcgen05.mma is bound by fetching the 128 x 16 A tile (~64
  // cycles) whenever N < 128, so splitting a 208-wide layer over output features (round 1: 112 + 96 with separate commit
  // barriers; 64 + 48 + 48 + 48 measured this round: 9.4k cycles per layer against 7.8k two-way) wastes tensor time.
  // Overlap comes from K-chaining instead (rows.cuh): the consumer's k-slabs trail the producer's epilogue rounds.
  // `xdep` > 1: the X columns this layer reads were final `xdep` stages ago (e.g. the value head's first layer reads
  // [belief | state], last written by the prior head 4 stages earlier), so its MMAs may run under the previous stage's
  // epilogue — they accumulate into the region of the dead H operand instead of the previous stage's accumulator region.
  void dense_to_h(const float* w, const float* b, int ld, int out_f, int col0, int ncols, int kofs, int ksl, int a_src,
                  int a_k16, int act, int flags = 0, int xdep = 1) {
    RStage& s = begin_stage();
    gemm(w, ld, {{0, out_f, 0}}, r16(out_f), col0, ncols, kofs, ksl, a_src, a_k16, 0, 0);
    bias_job(b, 0, nullptr, 0, out_f, r16(out_f));
    end_stage(s, R_ACT_H, flags, out_f, act, 0, 0, (a_src == 0 && xdep > 1) ? xdep : 1);
  }
  void gaussian_head(const float* w, const float* b, int ld, int n, int ksl, int epi, int flags) {
    const int np = r16(n);
    RStage& s = begin_stage();
    gemm(w, ld, {{0, n, 0}, {n, n, np}}, 2 * np, 0, ld, 0, ksl, 1, 0, 0, 0);
    bias_job(b, 0, nullptr, 0, n, np);
    bias_job(b, n, nullptr, 0, n, np);
    end_stage(s, epi, flags, n, 0, 0, np);
  }
  void belief_update(const repo_b200_rssm_weights* W, int D, int S, int A, int act) {
    const int kD16 = cdiv(D, 16), kx16 = cdiv(D + S + A, 16), kSA0 = D / 16;
    dense_to_h(W->fc_embed_state_action_w, W->fc_embed_state_action_b, S + A, D, 0, S + A, D - 16 * kSA0, kx16 - kSA0,
               0, kSA0, act);
    // chunks of 64 units; the last one is only as wide as the units that are left (200 units: 64 + 64 + 64 + 16
    // padded columns instead of 4 x 64 — the tail chunk's MMAs and weight stream shrink to a quarter)
    for (int u0 = 0, Wd = 0; u0 < D; u0 += Wd) {
      Wd = std::min(64, r16(D - u0));
      const int nu = std::min(Wd, D - u0);
      RStage& s = begin_stage();
      // accumulator columns [i_n | r | z | h_n]: W_hh . b -> [r z h_n] first (X only: it can run before the embedding layer
      // has finished), then W_ih . h -> [i_n r z] on top (its first MMA initialises the i_n columns)
      gemm(W->rnn_w_hh, D, {{u0, nu, 0}, {D + u0, nu, Wd}, {2 * D + u0, nu, 2 * Wd}}, 3 * Wd, 0, D, 0, kD16, 0, 0, Wd, 0);
      gemm(W->rnn_w_ih, D, {{2 * D + u0, nu, 0}, {u0, nu, Wd}, {D + u0, nu, 2 * Wd}}, 3 * Wd, 0, D, 0, kD16, 1, 0, 0, 1, Wd);
      bias_job(W->rnn_b_ih, u0, W->rnn_b_hh, u0, nu, Wd);
      bias_job(W->rnn_b_ih, D + u0, W->rnn_b_hh, D + u0, nu, Wd);
      bias_job(W->rnn_b_ih, 2 * D + u0, nullptr, 0, nu, Wd);
      bias_job(W->rnn_b_hh, 2 * D + u0, nullptr, 0, nu, Wd);
      end_stage(s, R_GRU, (u0 + Wd >= D) ? RF_LAST_CHUNK : 0, nu, 0, u0, Wd);
    }
  }
  void scalar_head(const repo_b200_mlp_weights* M, int D, int S, int Hd, int act, int flags, int xdep = 1) {
    const int kBS16 = cdiv(D + S, 16), kH16 = cdiv(Hd, 16);
    dense_to_h(M->w[0], M->b[0], D + S, Hd, 0, D + S, 0, kBS16, 0, 0, act, 0, xdep);
    dense_to_h(M->w[1], M->b[1], Hd, Hd, 0, Hd, 0, kH16, 1, 0, act);
    // fc3 + fc4 in one stage: the epilogue reduces act(fc3) against fc4's single weight row in fp32
    RStage& s = begin_stage();
    gemm(M->w[2], Hd, {{0, Hd, 0}}, r16(Hd), 0, Hd, 0, kH16, 1, 0, 0, 0);
    bias_job(M->b[2], 0, nullptr, 0, Hd, r16(Hd));
    bias_job(M->w[3], 0, nullptr, 0, Hd, r16(Hd));
    bias_job(M->b[3], 0, nullptr, 0, 1, 16);
    end_stage(s, R_ACT_DOT, flags, Hd, act, 0, 0);
  }
  // xback of every stage from the xdeps recorded while building (rounds of the stages in between, wrapping into the
  // previous time step); the first stage of the program may depend on a stage of the previous step.
  void finish_deps() {
    // stage 0 follows the LAST stage of the previous step: starting early is only safe when the two accumulate into
    // different TMEM regions (the region bookkeeping is periodic only if the program flips it an even number of times)
    if (n_stages && xdeps[0] > 1 && (P.stages[0].regs & 1) == (P.stages[n_stages - 1].regs & 1)) xdeps[0] = 1;
    for (int i = 0; i < n_stages; ++i) {
      int back = 0;
      for (int j = 1; j < xdeps[i]; ++j) back += P.stages[((i - j) % n_stages + n_stages) % n_stages].rounds;
      P.stages[i].xback = (uint8_t)std::min(back, 255);
    }
    // The first GRU chunk's W_hh GEMM reads the belief slot of X, final since the previous step: it may start as soon as the
    // previous stage's FIRST round has been signalled (by then every warp has left the stage before that one, including
    // its post-hand-off use of free TMEM columns as a transpose buffer — the region those MMAs accumulate into).
    for (int i = 0; i < n_stages; ++i) {
      const RStage& s = P.stages[i];
      const RStage& prev = P.stages[(i + n_stages - 1) % n_stages];
      if (s.epi == R_GRU && s.unit0 == 0 && prev.epi == R_ACT_H && prev.rounds > 1 && n_stages > 2)
        P.stages[i].xback = (uint8_t)(prev.rounds - 1);
    }
    int r0 = 0;
    for (int i = 0; i < n_stages; ++i) {
      P.stages[i].round0 = (uint16_t)r0;
      r0 += P.stages[i].rounds;
    }
    P.rounds_per_step = r0;
  }
  // belief refresh scratch (rows.cuh, R_GRU): one [hi | lo] plane pair per SM
  size_t scr_plane() const { return (size_t)cdiv(std::max(P.v.D, 1), 8) * kXLBO; }
  size_t scratch_bytes() const { return (size_t)std::max(1, sm_count()) * 2 * scr_plane(); }
  size_t packed_bytes() const { return align_up_(w_bytes, 256) + align_up_((size_t)n_bias * sizeof(float), 256) + scratch_bytes(); }
  static size_t align_up_(size_t v, size_t a) { return (v + a - 1) / a * a; }

  int bind_and_pack(void* ws, size_t ws_bytes, bool do_pack, cudaStream_t st) {
    if (overflow) return fail(-3, "rows program too large (stages %d gemms %d)", n_stages, n_gemms);
    if (ws_bytes < packed_bytes()) return fail(-4, "workspace too small: need %zu bytes, got %zu", packed_bytes(), ws_bytes);
    uint8_t* wb = static_cast<uint8_t*>(ws);
    float* bb = reinterpret_cast<float*>(wb + align_up_(w_bytes, 256));
    P.v.wblob = wb;
    P.v.bias = bb;
    P.n_rstages = n_stages;
    P.n_bias = n_bias;
    // transposed beliefs[t] stores: 32 free TMEM columns behind the GRU's H operand (the embedding layer's output)
    P.gru_tsc_col = (r16(P.v.D) + 8 * kEpiParts <= 256) ? r16(P.v.D) : -1;
    P.scr = wb + align_up_(w_bytes, 256) + align_up_((size_t)n_bias * sizeof(float), 256);
    P.scr_plane = (uint32_t)scr_plane();
    P.scr_slots = std::max(1, sm_count());
    finish_deps();
    if (do_pack) {
      pack.wblob = wb;
      bias.bias = bb;
      pack_rows_weights_kernel<<<blk, 256, 0, st>>>(pack);
      pack_rows_bias_kernel<<<bias.n_jobs, 256, 0, st>>>(bias);
      CUDA_OK(cudaGetLastError());
    }
    return 0;
  }
};

void rows_set_dims(RowsParams& P, const repo_b200_dims* d) {
  set_dims(P.v, d);
  P.kh_cols = 8 * P.v.kh16;
}

void build_imagine_rows(RBuilder& b, const repo_b200_dims* d, const repo_b200_rssm_weights* W,
                        const repo_b200_mlp_weights* actor, const repo_b200_mlp_weights* reward,
                        const repo_b200_mlp_weights* value, int act) {
  const int D = d->belief, S = d->state, A = d->action, Hd = d->hidden;
  const int kD16 = cdiv(D, 16), kBS16 = cdiv(D + S, 16), kH16 = cdiv(Hd, 16);
  rows_set_dims(b.P, d);
  // [belief | state] is final once the prior head of the previous step has run: with the scalar heads in the program, the
  // actor's first layer (next step) and the value head's first layer start under the epilogue of the stage before them
  const int n_head_stages = (reward ? 3 : 0) + (value ? 3 : 0);
  const bool dbg_no_early = (g_dbg_flags & 32) != 0;
  b.dense_to_h(actor->w[0], actor->b[0], D + S, Hd, 0, D + S, 0, kBS16, 0, 0, ACT_ELU, 0, dbg_no_early ? 1 : 1 + n_head_stages);
  for (int i = 1; i < 4; ++i) b.dense_to_h(actor->w[i], actor->b[i], Hd, Hd, 0, Hd, 0, kH16, 1, 0, ACT_ELU);
  b.gaussian_head(actor->w[4], actor->b[4], Hd, A, kH16, R_ACTION, 0);
  b.belief_update(W, D, S, A, act);
  b.dense_to_h(W->fc_embed_belief_prior_w, W->fc_embed_belief_prior_b, D, Hd, 0, D, 0, kD16, 0, 0, act);
  b.gaussian_head(W->fc_state_prior_w, W->fc_state_prior_b, Hd, S, kH16, R_PRIOR, SF_WRITES_STATE);
  if (reward) b.scalar_head(reward, D, S, Hd, act, 0);
  if (value) b.scalar_head(value, D, S, Hd, act, SF_SCALAR_VALUE, (reward && !dbg_no_early) ? 4 : 1);
}

void build_observe_rows(RBuilder& b, const repo_b200_dims* d, const repo_b200_rssm_weights* W, bool with_obs, int act) {
  const int D = d->belief, S = d->state, A = d->action, Hd = d->hidden, E = d->embed;
  const int kD16 = cdiv(D, 16), kH16 = cdiv(Hd, 16);
  rows_set_dims(b.P, d);
  b.belief_update(W, D, S, A, act);
  b.dense_to_h(W->fc_embed_belief_prior_w, W->fc_embed_belief_prior_b, D, Hd, 0, D, 0, kD16, 0, 0, act);
  b.gaussian_head(W->fc_state_prior_w, W->fc_state_prior_b, Hd, S, kH16, R_PRIOR,
                  with_obs ? 0 : (SF_WRITES_STATE | SF_LOADS_ACTION));
  if (with_obs) {
    b.dense_to_h(W->fc_embed_belief_posterior_w, W->fc_embed_belief_posterior_b, D + E, Hd, 0, D, 0, kD16, 0, 0, act,
                 SF_ADDEND);
    b.gaussian_head(W->fc_state_posterior_w, W->fc_state_posterior_b, Hd, S, kH16, R_POST,
                    SF_WRITES_STATE | SF_LOADS_ACTION);
  }
}

int launch_rows(const RowsParams& P, cudaStream_t st) {
  if (P.v.N <= 0 || P.v.n_steps <= 0) return 0;
  const size_t smem = rows_smem_bytes(P.v.kx16);
  if (smem > 227 * 1024) return fail(-5, "rows kernel: shared memory budget exceeded (%zu bytes)", smem);
  // program family -> kernel specialisation (rows.cuh, PROG): imagine-like programs have no addend / posterior stages,
  // observe-like ones no actor / scalar-head stages
  bool has_addend_post = false, has_actor = false;
  for (int i = 0; i < P.n_rstages; ++i) {
    const RStage& s = P.stages[i];
    has_addend_post |= s.epi == R_POST || (s.epi == R_ACT_H && (s.flags & SF_ADDEND));
    has_actor |= s.epi == R_ACT_DOT || s.epi == R_ACTION || s.epi == R_SCALAR;
  }
  const int prog = !has_addend_post ? 1 : (!has_actor ? 2 : 0);
  auto kern = prog == 1 ? rssm_rows_kernel<1> : (prog == 2 ? rssm_rows_kernel<2> : rssm_rows_kernel<0>);
  static size_t configured[3] = {0, 0, 0};
  if (smem > configured[prog]) {
    CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured[prog] = smem;
  }
  RowsParams Q = P;
  Q.v.dbg_flags = g_dbg_flags;
  Q.v.dbg_clock = g_dbg_clock;
  kern<<<cdiv(P.v.N, kRowsM), kRowsThreads, smem, st>>>(Q);
  CUDA_OK(cudaGetLastError());
  return 0;
}

// (T, rows, state) tensors the 128-row kernel touches with float2 accesses: 8-byte aligned bases (state sizes are even there)
bool rows_state_rows_aligned(std::initializer_list<const void*> ptrs) {
  for (const void* p : ptrs)
    if (reinterpret_cast<uintptr_t>(p) & 7) return false;
  return true;
}
// the 128-row kernel runs every no-stash program it supports; vm.cuh keeps the stash-writing training forwards, wide states /
// actions, the conditional model and single rows
bool use_rows_kernel(const repo_b200_dims* d, int n_rows, int row_tile) {
  if (d->state > 32 || d->action > 16) return false;  // the 128-row kernel keeps one row's Gaussian heads in registers
  if (d->state & 1) return false;                      // ... and reads / writes a row's states as float2 pairs
  // its other budgets: the bias staging buffer holds a scalar head's fc3 bias + fc4 row (2 * r16(hidden) + 16 floats), H and
  // the GRU's embedding operand fill one 256-column TMEM region, X for 128 rows must fit shared memory next to the ring
  // (upper bound of the bias floats of the largest program, imagine with both scalar heads: ten hidden layers, the GRU's four
  // gate vectors padded per chunk, two fused fc3+fc4 stages, the Gaussian / action heads)
  const int bias_floats = 10 * r16(d->hidden) + 4 * (r16(d->belief) + 48) + 2 * (2 * r16(d->hidden) + 16) + 4 * r16(d->state) + 64;
  if (bias_floats > kBiasCap || r16(d->belief) > 256 ||
      rows_smem_bytes(cdiv(d->belief + d->state + d->action, 16)) > 227 * 1024)
    return false;
  if (row_tile == 128) return true;
  if (row_tile == 16 || row_tile == 32 || row_tile == 64) return false;
  if (g_dbg_flags & 2) return false;
  if (g_dbg_flags & 4) return true;
  // measured (scripts/tile_crossover.py / observe_crossover.py, round 2): a launch of the 128-row kernel takes 0.97-0.99 ms for
  // any imagine (horizon 15) up to one wave of CTAs — 16 rows or 18,944 — against 1.19 ms for 16-row tiles of the vm kernel,
  // and 2.24-2.39 ms against 2.38-2.39 ms for a 49-step observe of 16..128 sequences: the per-step latency of ONE 128-row CTA
  // is no longer above that of a 16-row vm CTA, so the rows kernel runs everything but single rows (the acting step).
  return n_rows >= 8;
}


// ------------------------------------------------------------------------------------------ cluster observe (cluster.cuh)
// Largest batch the cluster kernel takes on its own initiative: 16 sequences per 16-CTA cluster, about six clusters resident
// at a time (a cluster needs 16 free SMs inside one GPC), 0.43 ms per wave of a 49-step observe against 2.2-2.5 ms for ANY
// batch up to 18,944 sequences on the 128-row kernel.  Measured (scripts/cluster_crossover.py): 16/50 sequences 0.43 ms,
// 128: 0.78, 256: 1.14, 512: 1.89, 768: 2.60 (128-row kernel: 2.47), 1024: 3.67 (2.50).  The backward has no such limit:
// the per-sequence fp32 kernel costs 12 us per sequence (1.77 ms per wave of 148), the cluster kernel 6 us.
constexpr int kClusterAutoBatch = 640;

int cluster_max_active() {   // co-resident 16-CTA clusters of the kernel on this device (0: cannot launch)
  static int cached = -1;
  if (cached >= 0) return cached;
  ClGeom g;
  if (!cl_geometry(200, 30, 6, 200, g)) return cached = 0;
  if (cudaFuncSetAttribute(rssm_cluster_observe_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) != cudaSuccess ||
      cudaFuncSetAttribute(rssm_cluster_observe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess) {
    cudaGetLastError();
    return cached = 0;
  }
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(kClSize * 8);
  cfg.blockDim = dim3(kClThreads);
  cfg.dynamicSmemBytes = g.smem_bytes;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = kClSize; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  int n = 0;
  if (cudaOccupancyMaxActiveClusters(&n, rssm_cluster_observe_kernel, &cfg) != cudaSuccess) { cudaGetLastError(); n = 0; }
  return cached = n;
}

bool ptrs_aligned(std::initializer_list<const void*> ptrs, uintptr_t a) {
  for (const void* p : ptrs)
    if (reinterpret_cast<uintptr_t>(p) & (a - 1)) return false;
  return true;
}
bool cluster_observe_fits(const repo_b200_dims* d) {
  ClGeom g;
  return cl_geometry(d->belief, d->state, d->action, d->hidden, g);
}
// row_tile 1 asks for the cluster kernel; auto (0) takes it for small batches
bool use_cluster_observe(const repo_b200_dims* d, int batch, int row_tile) {
  if (row_tile != 0 && row_tile != 1) return false;
  if (!cluster_observe_fits(d)) return false;
  if (row_tile == 0 && ((g_dbg_flags & 8) || batch > kClusterAutoBatch)) return false;
  return cluster_max_active() > 0;
}
size_t cluster_blob_bytes(const repo_b200_dims* d) {
  ClGeom g;
  if (!cl_geometry(d->belief, d->state, d->action, d->hidden, g)) return 0;
  return (size_t)kClSize * g.cta_bytes;
}

int launch_cluster_observe(const repo_b200_dims* d, const repo_b200_rssm_weights* W, ClParams& P, void* ws, bool do_pack,
                           cudaStream_t st) {
  ClGeom g;
  if (!cl_geometry(d->belief, d->state, d->action, d->hidden, g)) return fail(-5, "cluster observe: sizes unsupported");
  if (do_pack) {
    ClPackArgs a{};
    a.D = d->belief; a.S = d->state; a.A = d->action; a.Hd = d->hidden; a.E = d->embed; a.with_obs = P.with_obs;
    a.w_e = W->fc_embed_state_action_w; a.w_ih = W->rnn_w_ih; a.w_hh = W->rnn_w_hh;
    a.w_pp = W->fc_embed_belief_prior_w; a.w_prior = W->fc_state_prior_w;
    a.w_pq = W->fc_embed_belief_posterior_w; a.w_post = W->fc_state_posterior_w;
    a.wblob = static_cast<uint8_t*>(ws);
    pack_cluster_weights_kernel<<<dim3(8, kClSize), 256, 0, st>>>(a);
    CUDA_OK(cudaGetLastError());
  }
  P.wblob = static_cast<const uint8_t*>(ws);
  P.b_e = W->fc_embed_state_action_b; P.b_ih = W->rnn_b_ih; P.b_hh = W->rnn_b_hh;
  P.b_pp = W->fc_embed_belief_prior_b; P.b_prior = W->fc_state_prior_b;
  P.b_pq = W->fc_embed_belief_posterior_b; P.b_post = W->fc_state_posterior_b;
  P.dbg_clock = g_dbg_clock;
  if (cluster_max_active() <= 0) return fail(-5, "cluster observe: a 16-CTA cluster cannot be scheduled on this device");
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(kClSize * cdiv(P.N, kClRows));
  cfg.blockDim = dim3(kClThreads);
  cfg.dynamicSmemBytes = g.smem_bytes;
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = kClSize; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  CUDA_OK(cudaLaunchKernelEx(&cfg, rssm_cluster_observe_kernel, P));
  return 0;
}

}  // namespace
int conv_gemm_impl(const void* input_, const float* w_mat, int w_ld, int w_col0, const float* bias, const void* relu_mask_,
                   const float* scales, void* out_, int frames, int n_total, ConvMap cm, int hl_flags, const int* dense_opts,
                   void* ws, size_t ws_bytes, cudaStream_t st, int out_tile_rows = 0);
extern "C" size_t repo_b200_conv_workspace_bytes(int K, int n_total);
namespace {

// y = x @ W[:, w_col0 : w_col0 + in_f]^T + b for row-major operands.  From 256 rows (and 16-byte aligned, 4-float
// strided operands) this is one launch of the tcgen05 conv kernel run as a plain GEMM (K streams through its ring, so
// the 1024-wide embedding projection keeps 128-row tiles); below that the vm machine's one-step program.
// out_tile_rows > 0: the output goes out in the tiled layout of ConvParams::out_tile_rows (conv path only: the caller checks
// linear_tiled_ok first).
bool linear_tiled_ok(const float* x, int x_ld, int rows, int in_f, int out_f, size_t ws_bytes) {
  return rows >= 256 && !(reinterpret_cast<uintptr_t>(x) & 15) && !(x_ld & 3) && !(in_f & 3) && !(out_f & 3) &&
         ws_bytes >= repo_b200_conv_workspace_bytes(in_f, out_f) && !(g_dbg_flags & 256);
}
int run_linear(const float* x, int x_ld, int rows, int in_f, const float* w, int w_ld, int w_col0, const float* b,
               int out_f, float* y, int y_ld, void* ws, size_t ws_bytes, int row_tile, cudaStream_t st, int out_tile_rows = 0) {
  const bool aligned = !(reinterpret_cast<uintptr_t>(x) & 15) && !(reinterpret_cast<uintptr_t>(y) & 15) && !(x_ld & 3) &&
                       !(y_ld & 3) && !(in_f & 3);
  if (rows >= 256 && row_tile == 0 && aligned && ws_bytes >= repo_b200_conv_workspace_bytes(in_f, out_f) && !(g_dbg_flags & 256)) {
    ConvMap cm{};
    cm.RA = cm.RB = 1; cm.C = in_f; cm.pix = x_ld; cm.H = cm.W = 1; cm.TH = cm.TW = 1; cm.ntaps = 1;
    cm.sy = cm.sx = cm.dy = cm.dx = 1; cm.Ho = cm.Wo = 1; cm.osy = cm.osx = 1;
    const int opts[4] = {0, 0, y_ld != out_f ? y_ld : 0, 0};
    return conv_gemm_impl(x, w, w_ld, w_col0, b, nullptr, nullptr, y, rows, out_f, cm, 0, opts, ws, ws_bytes, st, out_tile_rows);
  }
  if (out_tile_rows) return fail(-1, "linear: tiled output needs the GEMM path");
  if (in_f < 1 || in_f > 1900) return fail(-1, "linear: in_features %d unsupported (1..1900)", in_f);
  if (out_f < 1 || out_f > 128 * 8) return fail(-1, "linear: out_features %d unsupported (1..1024)", out_f);
  Builder bl;
  build_linear(bl, w, b, w_ld, w_col0, in_f, out_f);
  int rc = bl.bind_and_pack(ws, ws_bytes, true, st);
  if (rc) return rc;
  VmParams& P = bl.P;
  P.n_steps = 1;
  P.N = rows;
  P.init_x = x; P.init_x_cols = in_f; P.init_x_ld = x_ld;
  P.out = y; P.out_ld = y_ld;
  // big K: small row tiles keep X inside shared memory
  int nt = row_tile;
  if (nt == 0) nt = pick_row_tile(rows, bl.max_acc_tiles, P.kx16, 0, 0);
  return launch(P, bl.max_acc_tiles, nt, st);
}

}  // namespace

extern "C" {

int repo_b200_version(void) { return 1; }
void repo_b200_debug_flags(int flags) { g_dbg_flags = flags; }
void repo_b200_debug_clock(void* device_buffer) { g_dbg_clock = static_cast<long long*>(device_buffer); }
const char* repo_b200_last_error(void) { return g_err; }

int repo_b200_device_info(int* sms, int* major, int* minor) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return fail(-2, "no CUDA device");
  cudaDeviceProp p;
  if (cudaGetDeviceProperties(&p, dev) != cudaSuccess) return fail(-2, "cudaGetDeviceProperties failed");
  if (sms) *sms = p.multiProcessorCount;
  if (major) *major = p.major;
  if (minor) *minor = p.minor;
  return 0;
}

size_t repo_b200_linear_workspace_bytes(int in_f, int out_f) {
  const size_t vm = (size_t)cdiv(out_f, 128) * cdiv(in_f, 16) * kSlabBytes + (size_t)cdiv(out_f, 128) * 512 + 64;
  return std::max(vm, repo_b200_conv_workspace_bytes(in_f, out_f));
}

int repo_b200_linear_fwd(const float* x, int x_ld, int rows, int in_f, const float* w, const float* b, int out_f,
                         float* y, int y_ld, void* ws, size_t ws_bytes, int row_tile, void* stream) {
  if (!x || !w || !y || !ws) return fail(-1, "linear: NULL pointer");
  return run_linear(x, x_ld, rows, in_f, w, in_f, 0, b, out_f, y, y_ld, ws, ws_bytes, row_tile,
                    static_cast<cudaStream_t>(stream));
}

size_t repo_b200_imagine_workspace_bytes(const repo_b200_dims* d) {
  if (check_dims(d)) return 0;
  // size the program with dummy (null) weights: the layout depends on the sizes only
  repo_b200_rssm_weights W{};
  repo_b200_mlp_weights m{};
  Builder b;
  build_imagine(b, d, &W, &m, &m, &m, ACT_ELU);
  RBuilder rb_;
  build_imagine_rows(rb_, d, &W, &m, &m, &m, ACT_ELU);
  return align_up(std::max(std::max(b.packed_bytes(), rb_.packed_bytes()), cluster_blob_bytes(d)), 256);
}

int repo_b200_imagine_fwd(const repo_b200_dims* d, const repo_b200_rssm_weights* W, const repo_b200_mlp_weights* actor,
                          const repo_b200_mlp_weights* reward, const repo_b200_mlp_weights* value,
                          const float* start_belief, const float* start_state, const float* eps_action,
                          const float* eps_prior, float* beliefs, float* prior_states, float* prior_means,
                          float* prior_std_devs, float* actions, float* rewards, float* values, float* returns,
                          int horizon, int n_rows, int act_kind, float min_std, float a_mean_scale, float a_init_std,
                          float a_min_std, float gamma, float lambda_, float* stash, void* ws, size_t ws_bytes,
                          int flags, int row_tile, void* stream) {
  return repo_b200_imagine_cond_fwd(d, W, actor, reward, value, start_belief, start_state, nullptr, 0, eps_action, eps_prior,
                                    beliefs, prior_states, prior_means, prior_std_devs, actions, rewards, values, returns,
                                    horizon, n_rows, act_kind, min_std, a_mean_scale, a_init_std, a_min_std, gamma, lambda_,
                                    stash, ws, ws_bytes, flags, row_tile, stream);
}

int repo_b200_imagine_cond_fwd(const repo_b200_dims* d, const repo_b200_rssm_weights* W, const repo_b200_mlp_weights* actor,
                               const repo_b200_mlp_weights* reward, const repo_b200_mlp_weights* value,
                               const float* start_belief, const float* start_state, const float* condition, int cond_size,
                               const float* eps_action, const float* eps_prior, float* beliefs, float* prior_states,
                               float* prior_means, float* prior_std_devs, float* actions, float* rewards, float* values,
                               float* returns, int horizon, int n_rows, int act_kind, float min_std, float a_mean_scale,
                               float a_init_std, float a_min_std, float gamma, float lambda_, float* stash, void* ws,
                               size_t ws_bytes, int flags, int row_tile, void* stream) {
  int rc = check_dims(d);
  if (cond_size < 0 || cond_size >= (d ? d->action : 1) || (cond_size > 0 && !condition))
    return fail(-1, "imagine: condition size %d must be in [0, action slot) and come with a tensor", cond_size);
  if (rc) return rc;
  if ((rc = check_act(act_kind))) return rc;
  if (!W || !actor) return fail(-1, "imagine: rssm / actor weights are required");
  if (actor->n_layers != 5) return fail(-1, "imagine: policy must be an ActorModel with fc1..fc5 (got %d layers)", actor->n_layers);
  if ((reward && reward->n_layers != 4) || (value && value->n_layers != 4)) return fail(-1, "imagine: reward/value heads must have fc1..fc4");
  if (horizon < 1 || n_rows < 0) return fail(-1, "imagine: bad horizon %d / rows %d", horizon, n_rows);
  if (horizon == 1 || n_rows == 0) return 0;  // nothing to roll out (the reference returns empty stacks)
  if (!start_belief || !start_state || !eps_action || !eps_prior || !beliefs || !prior_states || !prior_means || !prior_std_devs)
    return fail(-1, "imagine: NULL input/output pointer");
  if ((reward && !rewards) || (value && !values)) return fail(-1, "imagine: rewards/values output missing");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  // (the activation stash and the conditional variant are served by the vm kernel)
  const bool rows = !stash && cond_size == 0 && use_rows_kernel(d, n_rows, row_tile) &&
                    rows_state_rows_aligned({eps_prior, prior_states, prior_means, prior_std_devs});
  Builder b;
  RBuilder rbld;
  if (rows) {
    build_imagine_rows(rbld, d, W, actor, reward, value, act_kind);
    if ((rc = rbld.bind_and_pack(ws, ws_bytes, !(flags & REPO_B200_WEIGHTS_PACKED), st))) return rc;
  } else {
    build_imagine(b, d, W, actor, reward, value, act_kind, stash != nullptr, cond_size);
    if ((rc = b.bind_and_pack(ws, ws_bytes, !(flags & REPO_B200_WEIGHTS_PACKED), st))) return rc;
    b.P.stash = stash;
    b.P.stash_ld = 5 * d->belief + 5 * d->hidden + 2 * d->action;
  }
  VmParams& P = rows ? rbld.P.v : b.P;
  P.n_steps = horizon - 1;
  P.N = n_rows;
  P.min_std = min_std;
  P.a_mean_scale = a_mean_scale; P.a_init_std = a_init_std; P.a_min_std = a_min_std;
  P.gamma = gamma; P.lambda = lambda_;
  P.one_minus_lambda = (float)(1.0 - (double)lambda_);
  P.init_belief = start_belief; P.init_state = start_state;
  P.cond = cond_size > 0 ? condition : nullptr;
  P.eps_action = eps_action; P.eps_prior = eps_prior;
  P.beliefs = beliefs; P.prior_s = prior_states; P.prior_m = prior_means; P.prior_sd = prior_std_devs;
  P.actions_out = actions;
  P.rewards = reward ? rewards : nullptr;
  P.values = value ? values : nullptr;
  P.returns = (reward && value) ? returns : nullptr;
  if (rows) return launch_rows(rbld.P, st);
  return launch(P, b.max_acc_tiles, row_tile, st);
}

int repo_b200_tanh_normal_entropy_fwd(const float* mean, const float* std_dev, const float* eps, float* entropy, int m,
                                      int action, int samples, void* stream) {
  if (m < 0 || action < 1 || samples < 1) return fail(-1, "entropy: bad sizes m=%d action=%d samples=%d", m, action, samples);
  if (m == 0) return 0;
  if (!mean || !std_dev || !eps || !entropy) return fail(-1, "entropy: NULL pointer");
  const long long threads = (long long)m * kEntSlices;
  tanh_normal_entropy_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      mean, std_dev, eps, entropy, m, action, samples);
  CUDA_OK(cudaGetLastError());
  return 0;
}

int repo_b200_replay_gather(const uint8_t* obs, const float* actions, const float* rewards, const float* dones,
                            const long long* start_inds, int batch, int seq_len, long long pos, int full,
                            long long length, int frame_bytes, int action_dim, float* obs_out, float* actions_out,
                            float* rewards_out, float* nonterminals_out, long long* index_out, void* stream) {
  if (batch < 0 || seq_len < 0 || frame_bytes < 1 || action_dim < 1 || length < 1) return fail(-1, "replay_gather: bad sizes");
  if (batch == 0 || seq_len == 0) return 0;
  if (!obs || !actions || !rewards || !dones || !start_inds || !obs_out || !actions_out || !rewards_out || !nonterminals_out)
    return fail(-1, "replay_gather: NULL pointer");
  replay_gather_kernel<<<batch * seq_len, 128, 0, static_cast<cudaStream_t>(stream)>>>(
      obs, actions, rewards, dones, start_inds, batch, seq_len, pos, full, length, frame_bytes, action_dim, obs_out,
      actions_out, rewards_out, nonterminals_out, index_out);
  CUDA_OK(cudaGetLastError());
  return 0;
}

int repo_b200_observe_stash_floats(const repo_b200_dims* d) { return d ? 5 * d->belief + 2 * d->hidden : 0; }

int repo_b200_observe_bwd(const repo_b200_dims* d, const repo_b200_rssm_weights* W, const float* prev_belief,
                          const float* beliefs, const float* prior_std_devs, const float* post_std_devs,
                          const float* eps_prior, const float* eps_post, const float* nonterminals, const float* stash,
                          const float* g_beliefs, const float* g_prior_states, const float* g_prior_means,
                          const float* g_prior_std_devs, const float* g_post_states, const float* g_post_means,
                          const float* g_post_std_devs, float* d_q, float* d_hq, float* d_p, float* d_hp, float* d_gi,
                          float* d_gh, float* d_e, float* d_prev_belief, float* d_prev_state, int t1, int batch,
                          int with_obs, int act_kind, float min_std, void* stream) {
  int rc = check_dims(d);
  if (rc) return rc;
  if ((rc = check_act(act_kind))) return rc;
  if (t1 < 0 || batch < 0) return fail(-1, "observe_bwd: bad sizes");
  if (t1 == 0 || batch == 0) return 0;
  if (!W || !beliefs || !prior_std_devs || !eps_prior || !stash || !d_p || !d_hp || !d_gi || !d_gh || !d_e)
    return fail(-1, "observe_bwd: NULL pointer");
  if (with_obs && (!post_std_devs || !eps_post || !d_q || !d_hq)) return fail(-1, "observe_bwd: posterior buffers missing");
  ObsBwdParams P{};
  P.T = t1; P.B = batch; P.D = d->belief; P.S = d->state; P.A = d->action; P.Hd = d->hidden; P.E = d->embed;
  P.act = act_kind; P.with_obs = with_obs; P.min_std = min_std;
  P.w_e = W->fc_embed_state_action_w; P.w_ih = W->rnn_w_ih; P.w_hh = W->rnn_w_hh;
  P.w_p1 = W->fc_embed_belief_prior_w; P.w_p2 = W->fc_state_prior_w;
  P.w_q1 = W->fc_embed_belief_posterior_w; P.w_q2 = W->fc_state_posterior_w;
  P.init_belief = prev_belief; P.beliefs = beliefs; P.prior_sd = prior_std_devs; P.post_sd = post_std_devs;
  P.eps_prior = eps_prior; P.eps_post = eps_post; P.nonterm = nonterminals;
  P.stash = stash; P.stash_ld = 5 * d->belief + 2 * d->hidden;
  P.g_beliefs = g_beliefs; P.g_prior_s = g_prior_states; P.g_prior_m = g_prior_means; P.g_prior_sd = g_prior_std_devs;
  P.g_post_s = g_post_states; P.g_post_m = g_post_means; P.g_post_sd = g_post_std_devs;
  P.d_q = d_q; P.d_hq = d_hq; P.d_p = d_p; P.d_hp = d_hp; P.d_gi = d_gi; P.d_gh = d_gh; P.d_e = d_e;
  P.d_init_belief = d_prev_belief; P.d_init_state = d_prev_state;
  const size_t smem = (size_t)(10 * P.D + 5 * P.S + P.Hd + kObsGroups * 256) * sizeof(float);
  observe_bwd_kernel<<<batch, 256 * kObsGroups, smem, static_cast<cudaStream_t>(stream)>>>(P);
  CUDA_OK(cudaGetLastError());
  return 0;
}

size_t repo_b200_observe_bwd_workspace_bytes(const repo_b200_dims* d, int batch) {
  if (check_dims(d) || batch < 0) return 0;
  ClBwdGeom g;
  if (!clb_geometry(d->belief, d->state, d->action, d->hidden, g)) return 256;
  return align_up((size_t)kClSize * g.cta_bytes, 256) + align_up((size_t)std::max(batch, 1) * 2 * sizeof(float), 256);
}

static int cluster_bwd_max_active() {
  static int cached = -1;
  if (cached >= 0) return cached;
  ClBwdGeom g;
  if (!clb_geometry(200, 30, 6, 200, g)) return cached = 0;
  if (cudaFuncSetAttribute(rssm_cluster_observe_bwd_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) != cudaSuccess ||
      cudaFuncSetAttribute(rssm_cluster_observe_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess) {
    cudaGetLastError();
    return cached = 0;
  }
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(kClSize * 8);
  cfg.blockDim = dim3(kClThreads);
  cfg.dynamicSmemBytes = g.smem_bytes;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = kClSize; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  int n = 0;
  if (cudaOccupancyMaxActiveClusters(&n, rssm_cluster_observe_bwd_kernel, &cfg) != cudaSuccess) { cudaGetLastError(); n = 0; }
  return cached = n;
}

int repo_b200_observe_bwd_ws(const repo_b200_dims* d, const repo_b200_rssm_weights* W, const float* prev_belief,
                             const float* beliefs, const float* prior_std_devs, const float* post_std_devs,
                             const float* eps_prior, const float* eps_post, const float* nonterminals, const float* stash,
                             const float* g_beliefs, const float* g_prior_states, const float* g_prior_means,
                             const float* g_prior_std_devs, const float* g_post_states, const float* g_post_means,
                             const float* g_post_std_devs, float* d_q, float* d_hq, float* d_p, float* d_hp, float* d_gi,
                             float* d_gh, float* d_e, float* d_prev_belief, float* d_prev_state, int t1, int batch,
                             int with_obs, int act_kind, float min_std, void* ws, size_t ws_bytes, int mode, void* stream) {
  int rc = check_dims(d);
  if (rc) return rc;
  if (mode < 0 || mode > 2) return fail(-1, "observe_bwd: mode %d unknown (0 auto, 1 cluster, 2 per-sequence kernel)", mode);
  ClBwdGeom g;
  const bool fits = clb_geometry(d->belief, d->state, d->action, d->hidden, g);
  bool cluster = mode != 2 && fits && ws && ws_bytes >= repo_b200_observe_bwd_workspace_bytes(d, batch) &&
                 (mode == 1 || !(g_dbg_flags & 8)) &&
                 ptrs_aligned({ws, prev_belief, beliefs, stash, g_beliefs, d_hq, d_hp, d_gi, d_gh, d_e}, 16) &&
                 ptrs_aligned({prior_std_devs, post_std_devs, eps_prior, eps_post, g_prior_states, g_prior_means, g_prior_std_devs,
                               g_post_states, g_post_means, g_post_std_devs, d_q, d_p}, 8);
  if (cluster && t1 > 0 && batch > 0 && cluster_bwd_max_active() <= 0) cluster = false;
  if (mode == 1 && !cluster)
    return fail(-5, "observe_bwd: the cluster kernel (mode 1) does not take these sizes / this workspace / this device");
  if (!cluster)
    return repo_b200_observe_bwd(d, W, prev_belief, beliefs, prior_std_devs, post_std_devs, eps_prior, eps_post, nonterminals,
                                 stash, g_beliefs, g_prior_states, g_prior_means, g_prior_std_devs, g_post_states,
                                 g_post_means, g_post_std_devs, d_q, d_hq, d_p, d_hp, d_gi, d_gh, d_e, d_prev_belief,
                                 d_prev_state, t1, batch, with_obs, act_kind, min_std, stream);
  if ((rc = check_act(act_kind))) return rc;
  if (t1 < 0 || batch < 0) return fail(-1, "observe_bwd: bad sizes");
  if (t1 == 0 || batch == 0) return 0;
  if (!W || !beliefs || !prior_std_devs || !eps_prior || !stash || !d_p || !d_hp || !d_gi || !d_gh || !d_e)
    return fail(-1, "observe_bwd: NULL pointer");
  if (with_obs && (!post_std_devs || !eps_post || !d_q || !d_hq)) return fail(-1, "observe_bwd: posterior buffers missing");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  uint8_t* base = static_cast<uint8_t*>(ws);
  float* scales = reinterpret_cast<float*>(base + align_up((size_t)kClSize * g.cta_bytes, 256));
  ClBwdPackArgs a{};
  a.D = d->belief; a.S = d->state; a.A = d->action; a.Hd = d->hidden; a.E = d->embed; a.with_obs = with_obs;
  a.w_e = W->fc_embed_state_action_w; a.w_ih = W->rnn_w_ih; a.w_hh = W->rnn_w_hh;
  a.w_pp = W->fc_embed_belief_prior_w; a.w_prior = W->fc_state_prior_w;
  a.w_pq = W->fc_embed_belief_posterior_w; a.w_post = W->fc_state_posterior_w;
  a.wblob = base;
  pack_cluster_bwd_weights_kernel<<<dim3(7, kClSize), 256, 0, st>>>(a);
  CUDA_OK(cudaGetLastError());
  ClBwdParams P{};
  P.T = t1; P.N = batch; P.D = d->belief; P.S = d->state; P.A = d->action; P.Hd = d->hidden;
  P.act = act_kind; P.with_obs = with_obs; P.min_std = min_std;
  P.wblob = base; P.scales = scales;
  P.init_belief = prev_belief; P.beliefs = beliefs; P.prior_sd = prior_std_devs; P.post_sd = post_std_devs;
  P.eps_prior = eps_prior; P.eps_post = eps_post; P.nonterm = nonterminals;
  P.stash = stash; P.stash_ld = 5 * d->belief + 2 * d->hidden;
  P.g_beliefs = g_beliefs; P.g_prior_s = g_prior_states; P.g_prior_m = g_prior_means; P.g_prior_sd = g_prior_std_devs;
  P.g_post_s = g_post_states; P.g_post_m = g_post_means; P.g_post_sd = g_post_std_devs;
  P.d_q = d_q; P.d_hq = d_hq; P.d_p = d_p; P.d_hp = d_hp; P.d_gi = d_gi; P.d_gh = d_gh; P.d_e = d_e;
  P.d_init_belief = d_prev_belief; P.d_init_state = d_prev_state;
  P.dbg_clock = g_dbg_clock;
  observe_bwd_scale_kernel<<<batch, 256, 0, st>>>(P, scales);
  CUDA_OK(cudaGetLastError());
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(kClSize * cdiv(batch, kClRows));
  cfg.blockDim = dim3(kClThreads);
  cfg.dynamicSmemBytes = g.smem_bytes;
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = kClSize; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  CUDA_OK(cudaLaunchKernelEx(&cfg, rssm_cluster_observe_bwd_kernel, P));
  return 0;
}

int repo_b200_imagine_stash_floats(const repo_b200_dims* d) { return d ? 5 * d->belief + 5 * d->hidden + 2 * d->action : 0; }

int repo_b200_imagine_bwd(const repo_b200_dims* d, const repo_b200_rssm_weights* W, const repo_b200_mlp_weights* actor,
                          const float* start_belief, const float* beliefs, const float* actions,
                          const float* prior_std_devs, const float* eps_prior, const float* eps_action, const float* stash,
                          const float* g_beliefs, const float* g_prior_states, const float* g_prior_means,
                          const float* g_prior_std_devs, float* d_p, float* d_hp, float* d_gi, float* d_gh, float* d_e,
                          float* d_a5, float* d_a4, float* d_a3, float* d_a2, float* d_a1, float* d_start_belief,
                          float* d_start_state, int horizon, int n_rows, int act_kind, float min_std,
                          float a_mean_scale, float a_min_std, void* stream) {
  return repo_b200_imagine_cond_bwd(d, W, actor, 0, start_belief, beliefs, actions, prior_std_devs, eps_prior, eps_action, stash,
                                    g_beliefs, g_prior_states, g_prior_means, g_prior_std_devs, d_p, d_hp, d_gi, d_gh, d_e, d_a5,
                                    d_a4, d_a3, d_a2, d_a1, d_start_belief, d_start_state, horizon, n_rows, act_kind, min_std,
                                    a_mean_scale, a_min_std, stream);
}

int repo_b200_imagine_cond_bwd(const repo_b200_dims* d, const repo_b200_rssm_weights* W, const repo_b200_mlp_weights* actor,
                               int cond_size, const float* start_belief, const float* beliefs, const float* actions,
                               const float* prior_std_devs, const float* eps_prior, const float* eps_action,
                               const float* stash, const float* g_beliefs, const float* g_prior_states,
                               const float* g_prior_means, const float* g_prior_std_devs, float* d_p, float* d_hp, float* d_gi,
                               float* d_gh, float* d_e, float* d_a5, float* d_a4, float* d_a3, float* d_a2, float* d_a1,
                               float* d_start_belief, float* d_start_state, int horizon, int n_rows, int act_kind,
                               float min_std, float a_mean_scale, float a_min_std, void* stream) {
  int rc = check_dims(d);
  if (cond_size < 0 || cond_size >= (d ? d->action : 1)) return fail(-1, "imagine_bwd: bad condition size %d", cond_size);
  if (rc) return rc;
  if ((rc = check_act(act_kind))) return rc;
  if (horizon < 1 || n_rows < 0) return fail(-1, "imagine_bwd: bad sizes");
  if (horizon == 1 || n_rows == 0) return 0;
  if (!W || !actor || actor->n_layers != 5) return fail(-1, "imagine_bwd: weights missing");
  if (!start_belief || !beliefs || !actions || !prior_std_devs || !eps_prior || !eps_action || !stash || !d_p || !d_hp ||
      !d_gi || !d_gh || !d_e || !d_a5)
    return fail(-1, "imagine_bwd: NULL pointer");
  // d_a4 .. d_a1 all NULL: the caller hoists the actor's hidden layers out of the time loop (they feed nothing back)
  if ((!d_a4 || !d_a3 || !d_a2 || !d_a1) && (d_a4 || d_a3 || d_a2 || d_a1))
    return fail(-1, "imagine_bwd: d_a4 .. d_a1 must be given together or all be NULL");
  ImgBwdParams P{};
  P.T = horizon - 1; P.N = n_rows; P.D = d->belief; P.S = d->state; P.A = d->action; P.Hd = d->hidden;
  P.A_act = d->action - cond_size;
  P.act = act_kind; P.min_std = min_std; P.a_mean_scale = a_mean_scale; P.a_min_std = a_min_std;
  P.w_e = W->fc_embed_state_action_w; P.w_ih = W->rnn_w_ih; P.w_hh = W->rnn_w_hh;
  P.w_p1 = W->fc_embed_belief_prior_w; P.w_p2 = W->fc_state_prior_w;
  P.w_a2 = actor->w[1]; P.w_a3 = actor->w[2]; P.w_a4 = actor->w[3]; P.w_a5 = actor->w[4];
  P.start_belief = start_belief; P.beliefs = beliefs; P.actions = actions; P.prior_sd = prior_std_devs;
  P.eps_prior = eps_prior; P.eps_action = eps_action;
  P.stash = stash; P.stash_ld = repo_b200_imagine_stash_floats(d);
  P.g_beliefs = g_beliefs; P.g_prior_s = g_prior_states; P.g_prior_m = g_prior_means; P.g_prior_sd = g_prior_std_devs;
  P.d_p = d_p; P.d_hp = d_hp; P.d_gi = d_gi; P.d_gh = d_gh; P.d_e = d_e;
  P.d_a5 = d_a5; P.d_a4 = d_a4; P.d_a3 = d_a3; P.d_a2 = d_a2; P.d_a1 = d_a1;
  P.d_start_belief = d_start_belief; P.d_start_state = d_start_state;
  // rows per CTA: every CTA re-reads all weights each step, so prefer the smallest block that still fits the rollout
  // into ONE wave of CTAs (2450 rows on 148 SMs -> 20 rows, 123 CTAs: 3.2 ms instead of 3.8 ms with 8 rows).
  // Tried and dropped: streaming the weights through a shared-memory ring with bulk copies from a producer warp —
  // same time for 8, 12 and 20 rows per CTA, i.e. the kernel is bound by its FMA / shared-load issue rate and the
  // per-stage stash traffic, not by the weight fetch latency.  Round 2 confirmed it from the other side: splitting every
  // reduction over two thread groups (512 threads, 16 warps instead of 8 to hide the L2 latency with) — 3.31 vs 3.34 ms.
  const int sms = std::max(1, sm_count());
  // (few rows, e.g. a data-parallel shard of 343: 4 rows per CTA spread the shard over 86 SMs instead of 43 — the per-step
  // time of a CTA is what such a launch costs, and it grows with the rows the CTA carries)
  // Up to 12 rows per CTA two CTAs are resident per SM (bwd.cuh): the block size is the smallest that keeps the launch
  // inside one wave of 2 x SMs CTAs (2,450 rows -> 12 rows, 205 CTAs).
  const int rb = n_rows <= 4 * sms ? 4 : (n_rows <= 16 * sms ? 8 : (n_rows <= 24 * sms ? 12 : 20));
  const size_t smem = (size_t)(9 * P.D + 3 * P.S + 2 * P.Hd + 2 * P.A) * rb * sizeof(float);
  if (smem > 220 * 1024) return fail(-1, "imagine_bwd: model too wide for the %d-row backward block", rb);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  auto go = [&](auto kernel, size_t& configured) -> int {
    if (smem > configured) {
      CUDA_OK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      configured = smem;
    }
    kernel<<<cdiv(n_rows, rb), 256, smem, st>>>(P);
    CUDA_OK(cudaGetLastError());
    return 0;
  };
  static size_t c4 = 0, c8 = 0, c12 = 0, c20 = 0;
  if (rb == 4) return go(imagine_bwd_kernel<4>, c4);
  if (rb == 8) return go(imagine_bwd_kernel<8>, c8);
  if (rb == 12) return go(imagine_bwd_kernel<12>, c12);
  return go(imagine_bwd_kernel<20>, c20);
}

int repo_b200_colsum(const float* x, long long rows, int cols, long long ld, float* out, void* stream) {
  if (rows < 0 || cols < 0 || ld < cols) return fail(-1, "colsum: bad sizes");
  if (cols == 0) return 0;
  if (!out || (rows > 0 && !x)) return fail(-1, "colsum: NULL pointer");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  CUDA_OK(cudaMemsetAsync(out, 0, (size_t)cols * sizeof(float), st));
  if (rows == 0) return 0;
  // enough row slices to fill the machine (the column tiles alone are 1..19 blocks), at least 64 rows each
  const int ct = cdiv(cols, 32);
  const int slices = (int)std::max<long long>(1, std::min<long long>((rows + 63) / 64, std::max(1, 4 * sm_count() / ct)));
  colsum_kernel<<<dim3(ct, slices), 256, 0, st>>>(x, rows, cols, ld, out);
  CUDA_OK(cudaGetLastError());
  return 0;
}

int repo_b200_sqnorm_accumulate(const float* grad, long long n, float* sqnorm, void* stream) {
  if (n < 0) return fail(-1, "sqnorm: bad size");
  if (n == 0) return 0;
  if (!grad || !sqnorm) return fail(-1, "sqnorm: NULL pointer");
  const int blocks = (int)std::min<long long>((n + 255) / 256, 148 * 8);
  sqnorm_kernel<<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(grad, n, sqnorm);
  CUDA_OK(cudaGetLastError());
  return 0;
}

int repo_b200_adam_clip_step(float* param, float* grad, float* exp_avg, float* exp_avg_sq, long long n, const float* sqnorm,
                             float max_norm, float lr, float beta1, float beta2, float eps, int step, void* stream) {
  if (n < 0 || step < 1) return fail(-1, "adam: bad size / step");
  if (n == 0) return 0;
  if (!param || !grad || !exp_avg || !exp_avg_sq) return fail(-1, "adam: NULL pointer");
  const double bc1 = 1.0 - std::pow((double)beta1, step), bc2 = 1.0 - std::pow((double)beta2, step);
  const int blocks = (int)std::min<long long>((n + 255) / 256, 148 * 8);
  adam_clip_kernel<<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(param, grad, exp_avg, exp_avg_sq, n, sqnorm, max_norm,
                                                                         lr, beta1, beta2, eps, (float)bc1, (float)std::sqrt(bc2));
  CUDA_OK(cudaGetLastError());
  return 0;
}

int repo_b200_adam_clip_step_dev(float* param, float* grad, float* exp_avg, float* exp_avg_sq, long long n, const float* sqnorm,
                                 float max_norm, float lr, float beta1, float beta2, float eps, int* step_dev, void* stream) {
  if (n < 0 || !step_dev) return fail(-1, "adam: bad size / NULL step counter");
  if (n == 0) return 0;
  if (!param || !grad || !exp_avg || !exp_avg_sq) return fail(-1, "adam: NULL pointer");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  adam_step_inc_kernel<<<1, 1, 0, st>>>(step_dev);
  const int blocks = (int)std::min<long long>((n + 255) / 256, 148 * 8);
  adam_clip_dev_kernel<<<blocks, 256, 0, st>>>(param, grad, exp_avg, exp_avg_sq, n, sqnorm, max_norm, lr, beta1, beta2, eps, step_dev);
  CUDA_OK(cudaGetLastError());
  return 0;
}

// ---- (transposed) convolution as implicit GEMM on the vm machine
// ---- implicit-GEMM convolution (conv.cuh).  Workspace = packed fp16 hi/lo weights + padded bias.
static void conv_geometry(int K, int n_total, int& k16, int& NP, int& n_tiles) {
  k16 = cdiv(K, 16);
  NP = n_total <= 256 ? cdiv(n_total, 16) * 16 : 256;
  n_tiles = cdiv(n_total, NP);
}

size_t repo_b200_conv_workspace_bytes(int K, int n_total) {
  if (K < 1 || n_total < 1) return 0;
  int k16, NP, n_tiles;
  conv_geometry(K, n_total, k16, NP, n_tiles);
  return (size_t)n_tiles * k16 * NP * 64 + (size_t)n_tiles * NP * sizeof(float) + 256;
}

}  // extern "C"

// w_mat: (n_total, K) window of a row-major matrix with row stride w_ld (0 = K) starting at column w_col0
int conv_gemm_impl(const void* input_, const float* w_mat, int w_ld, int w_col0, const float* bias, const void* relu_mask_,
                   const float* scales, void* out_, int frames, int n_total, ConvMap cm, int hl_flags, const int* dense_opts,
                   void* ws, size_t ws_bytes, cudaStream_t st, int out_tile_rows) {
  const float* input = static_cast<const float*>(input_);
  const float* relu_mask = static_cast<const float*>(relu_mask_);
  float* out = static_cast<float*>(out_);
  if (!input || !w_mat || !out || !ws) return fail(-1, "conv: NULL pointer");
  cm.enabled = 1;
  if (cm.pix == 0) cm.pix = cm.C;
  if (cm.pix < cm.C) return fail(-1, "conv: pixel stride %d < channels %d", cm.pix, cm.C);
  if (cm.tap0 != 0 || cm.ntaps != cm.TH * cm.TW || cm.accumulate) return fail(-1, "conv: partial tap windows are not supported");
  if (cm.ntaps < 1 || cm.C < 1 || n_total < 1 || n_total > 256 * kMaxRowsJobs) return fail(-1, "conv: bad sizes");
  if (cm.shuffle && (n_total % 4)) return fail(-1, "conv: sub-pixel store needs 4*cout features");
  if ((reinterpret_cast<uintptr_t>(input) & 15) || (reinterpret_cast<uintptr_t>(out) & 15) ||
      (relu_mask && (reinterpret_cast<uintptr_t>(relu_mask) & 15)))
    return fail(-1, "conv: tensors must be 16-byte aligned");
  const int K = cm.ntaps * cm.C;
  const long long rows = (long long)frames * cm.RA * cm.RB;
  if (rows <= 0) return 0;
  if (rows > 0x7fffffffLL - 128) return fail(-1, "conv: too many rows");
  ConvParams P{};
  conv_geometry(K, n_total, P.k16, P.NP, P.n_tiles);
  if (ws_bytes < repo_b200_conv_workspace_bytes(K, n_total)) return fail(-2, "conv: workspace too small");
  uint8_t* wblob = static_cast<uint8_t*>(ws);
  float* bias_p = reinterpret_cast<float*>(wblob + (size_t)P.n_tiles * P.k16 * P.NP * 64);
  {
    PackRowsArgs pa{};
    BiasRowsArgs ba{};
    pa.wblob = wblob;
    pa.scale = scales ? scales + 1 : nullptr;
    ba.bias = bias_p;
    for (int nt = 0; nt < P.n_tiles; ++nt) {
      PackRowsJob& j = pa.jobs[nt];
      j.w = w_mat; j.ld = w_ld ? w_ld : K; j.col0 = w_col0; j.ncols = K; j.kofs = 0; j.ksl = P.k16; j.n_pad = P.NP; j.nseg = 1;
      j.seg_src[0] = nt * P.NP; j.seg_n[0] = std::min(P.NP, n_total - nt * P.NP); j.seg_dst[0] = 0;
      j.dst_off16 = (uint32_t)((size_t)nt * P.k16 * P.NP * 4);
      j.blk0 = nt * P.k16;
      BiasRowsJob& b = ba.jobs[nt];
      b.a = bias; b.b = nullptr; b.a_off = nt * P.NP; b.b_off = 0; b.n = j.seg_n[0]; b.n_pad = P.NP; b.dst_off = nt * P.NP;
    }
    pa.n_jobs = ba.n_jobs = P.n_tiles;
    pack_rows_both_kernel<<<P.n_tiles * P.k16 + P.n_tiles, 256, 0, st>>>(pa, ba, P.n_tiles * P.k16);
    CUDA_OK(cudaGetLastError());
  }
  P.x = input; P.wblob = wblob; P.bias = bias_p; P.relu_mask = relu_mask; P.scales = scales; P.out = out;
  P.cm = cm;
  P.n_rows = (int)rows; P.K = K; P.n_total = n_total;
  P.cout = cm.shuffle ? n_total / 4 : n_total;
  // split-activation (fp16 hi/lo plane) operands: bit0 input, bit1 output, bit2 relu mask
  P.in_hl = hl_flags & 1; P.out_hl = (hl_flags >> 1) & 1; P.mask_hl = (hl_flags >> 2) & 1;
  if (P.in_hl && (cm.in_nchw || (cm.C & 7) || cm.pix != cm.C || scales)) return fail(-1, "conv: HL input needs NHWC, C %% 8 == 0 and no operand scaling");
  if ((P.out_hl || P.mask_hl) && (cm.out_nchw || (P.cout & 15) || (P.NP & 15)))
    return fail(-1, "conv: HL output / mask needs an NHWC output with cout %% 16 == 0");
  if (P.mask_hl && !relu_mask) return fail(-1, "conv: mask_hl without a mask");
  P.in_lo_bytes = (long long)frames * cm.H * cm.W * cm.C * 2;
  P.out_lo_elems = P.mask_lo_elems = (long long)frames * cm.Ho * cm.Wo * P.cout;
  P.a_lbo = P.in_hl ? kCvALboHL : kCvALbo;
  P.dbg = g_dbg_clock;
  if (dense_opts) {
    P.act_elu = dense_opts[0]; P.mask_elu = dense_opts[1]; P.out_ld = dense_opts[2]; P.mask_ld = dense_opts[3];
    if ((P.out_ld || P.mask_ld) && (cm.RA != 1 || cm.RB != 1 || cm.Ho != 1 || cm.Wo != 1 || cm.shuffle || cm.out_nchw || hl_flags))
      return fail(-1, "conv: row strides are for plain fp32 GEMM maps only");
    if ((P.out_ld && (P.out_ld < n_total || (P.out_ld & 3))) || (P.mask_ld && (P.mask_ld < n_total || (P.mask_ld & 3))))
      return fail(-1, "conv: row strides must be multiples of 4 and >= n_total");
  }
  P.out_tile_rows = out_tile_rows;
  if (out_tile_rows && (cm.RA != 1 || cm.RB != 1 || cm.Ho != 1 || cm.Wo != 1 || cm.shuffle || cm.out_nchw || hl_flags || P.out_ld || (n_total & 3)))
    return fail(-1, "conv: tiled output is for plain fp32 GEMM maps with n_total % 4 == 0 only");
  P.kc16 = P.NP <= 128 ? 4 : 2;
  P.stage_bytes = conv_stage_bytes(P.kc16, P.NP);
  const int n_ent = conv_table_entries(cm, P.k16, P.in_hl);
  if (n_ent > 2048) return fail(-1, "conv: K = %d needs %d gather-table entries (max 2048)", K, n_ent);
  P.bias_smem = (P.n_tiles * P.NP <= 4096) ? 1 : 0;
  const size_t fixed = 128 + 4 * 128 * sizeof(ConvRowInfo) + kCvColEntries * sizeof(ConvCol) + (size_t)n_ent * sizeof(ConvTap) +
                       (P.bias_smem ? (size_t)P.n_tiles * P.NP * sizeof(float) + 32 : 0);
  P.n_stages = std::min<int>(kCvMaxStages, (int)((227 * 1024 - fixed) / P.stage_bytes));
  if (P.n_stages < 2) return fail(-1, "conv: ring does not fit shared memory");
  const size_t smem = (size_t)P.n_stages * P.stage_bytes + fixed;
  static size_t configured = 0;
  if (smem > configured) {
    CUDA_OK(cudaFuncSetAttribute(conv_rows_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CUDA_OK(cudaFuncSetAttribute(conv_rows_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = smem;
  }
  const int items = cdiv((int)rows, 128) * P.n_tiles;
  const int grid = std::min(items, std::max(1, sm_count()));
  if (scales) conv_rows_kernel<true><<<grid, kCvThreads, smem, st>>>(P);
  else conv_rows_kernel<false><<<grid, kCvThreads, smem, st>>>(P);
  CUDA_OK(cudaGetLastError());
  return 0;
}

extern "C" {

int repo_b200_conv_gemm(const void* input_, const float* w_mat, const float* bias, const void* relu_mask_,
                        const float* scales, void* out_, int frames, int n_total, const int* map /* ConvMap as 28 ints */,
                        int hl_flags, const int* dense_opts /* nullable: act_elu, mask_elu, out_ld, mask_ld */, void* ws,
                        size_t ws_bytes, void* stream) {
  if (!map) return fail(-1, "conv: NULL pointer");
  ConvMap cm;
  static_assert(sizeof(ConvMap) == 28 * sizeof(int), "ConvMap layout");
  std::memcpy(&cm, map, sizeof(cm));
  return conv_gemm_impl(input_, w_mat, 0, 0, bias, relu_mask_, scales, out_, frames, n_total, cm, hl_flags, dense_opts, ws,
                        ws_bytes, static_cast<cudaStream_t>(stream));
}

int repo_b200_conv_wgrad(const void* input_, const float* grad_rows, const float* scales, float* dw, int frames,
                         int n_total, int g_ld, const int* map /* ConvMap as 27 ints */, int input_hl, void* stream) {
  const float* input = static_cast<const float*>(input_);
  if (!input || !grad_rows || !dw || !map) return fail(-1, "conv_wgrad: NULL pointer");
  ConvMap cm;
  std::memcpy(&cm, map, sizeof(cm));
  cm.enabled = 1;
  if (cm.pix == 0) cm.pix = cm.C;
  if (cm.pix < cm.C) return fail(-1, "conv: pixel stride %d < channels %d", cm.pix, cm.C);
  if (cm.tap0 != 0 || cm.ntaps != cm.TH * cm.TW) return fail(-1, "conv_wgrad: partial tap windows are not supported");
  if (cm.ntaps < 1 || cm.C < 1 || n_total < 1 || (n_total & 3) || g_ld < n_total || (g_ld & 3))
    return fail(-1, "conv_wgrad: n_total must be a multiple of 4 (got %d, row stride %d)", n_total, g_ld);
  // more than 256 output features: slices of 256 (the accumulator tile), all in one launch (blockIdx.y = slice)
  const int n_all = n_total;
  const int n_slices = cdiv(n_all, 256);
  if (n_slices > 1) n_total = 256;
  const int n_last = n_all - (n_slices - 1) * 256;
  if ((reinterpret_cast<uintptr_t>(input) & 15) || (reinterpret_cast<uintptr_t>(grad_rows) & 15))
    return fail(-1, "conv_wgrad: tensors must be 16-byte aligned");
  const int K = cm.ntaps * cm.C;
  const long long rows = (long long)frames * cm.RA * cm.RB;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  CUDA_OK(cudaMemsetAsync(dw, 0, (size_t)n_all * K * sizeof(float), st));
  if (rows <= 0) return 0;
  if (rows > 0x7fffffffLL - 128) return fail(-1, "conv_wgrad: too many rows");
  WgradParams P{};
  P.x = input; P.g = grad_rows; P.dw = dw; P.scales = scales; P.cm = cm;
  P.n_rows = (int)rows; P.K = K; P.k16 = cdiv(K, 16); P.n_total = n_total; P.g_ld = g_ld;
  P.NP = cdiv(n_total, 16) * 16;
  P.n_last = n_last;
  P.x_hl = input_hl ? 1 : 0;
  P.dbg = g_dbg_clock;
  if (P.x_hl && (cm.in_nchw || (cm.C & 7) || cm.pix != cm.C)) return fail(-1, "conv_wgrad: HL input needs NHWC and C %% 8 == 0");
  P.x_lo_bytes = (long long)frames * cm.H * cm.W * cm.C * 2;
  const int n_ent = conv_table_entries(cm, P.k16, P.x_hl);
  if (n_ent > 2048) return fail(-1, "conv_wgrad: K = %d needs %d gather-table entries (max 2048)", K, n_ent);
  // super tiles: mt k-tiles of 128 share one CTA's TMEM (mt * NP <= 512 columns) and ring stage
  const int m_tiles = cdiv(K, 128);
  const int mt_max = P.NP <= 128 ? 4 : 2;
  P.n_super = cdiv(m_tiles, mt_max);
  if (P.n_super > kWgMaxSuper) return fail(-1, "conv_wgrad: K = %d too large", K);
  const int base = m_tiles / P.n_super, extra = m_tiles % P.n_super;
  const int stages_total = cdiv((int)rows, kWgRows);
  const int sms = std::max(1, sm_count());
  int m = 0, cta = 0, mt_top = 0;
  for (int s = 0; s < P.n_super; ++s) {
    P.m0[s] = m;
    P.mt[s] = base + (s < extra ? 1 : 0);
    m += P.mt[s];
    mt_top = std::max(mt_top, P.mt[s]);
    P.cta0[s] = cta;
    // row splits: fill the SMs once over all (slice, super tile) pairs — every split adds one atomic pass over its dW tile
    const int splits = std::max(1, std::min(stages_total, (sms * P.mt[s] + m_tiles * n_slices / 2) / (m_tiles * n_slices)));
    cta += splits;
  }
  P.cta0[P.n_super] = cta;
  P.kt_max = 128 * mt_top;
  P.stage_bytes = wgrad_stage_bytes(P.kt_max, P.NP);
  P.n_stages = std::min(kCvMaxStages, (200 * 1024) / P.stage_bytes);
  if (P.n_stages < 1) return fail(-1, "conv_wgrad: stage does not fit shared memory");
  const size_t smem = (size_t)P.n_stages * P.stage_bytes + 128 + (size_t)n_ent * sizeof(ConvTap);
  static size_t configured = 0;
  if (smem > configured) {
    CUDA_OK(cudaFuncSetAttribute(conv_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = smem;
  }
  conv_wgrad_kernel<<<dim3(cta, n_slices), kCvThreads, smem, st>>>(P);
  CUDA_OK(cudaGetLastError());
  return 0;
}

int repo_b200_pow2_scale(const float* x, long long n, float target, int which, float* scales, void* scratch, void* stream) {
  if (!x || !scales || !scratch || (which != 0 && which != 1)) return fail(-1, "pow2_scale: bad argument");
  if (reinterpret_cast<uintptr_t>(x) & 15) return fail(-1, "pow2_scale: tensor must be 16-byte aligned");
  if (n <= 0) return 0;
  const int blocks = (int)std::min<long long>((n / 4 + 255) / 256 + 1, (long long)std::max(1, sm_count()) * 8);
  absmax_scale_kernel<<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(x, n, static_cast<unsigned*>(scratch), target, scales, which);
  CUDA_OK(cudaGetLastError());
  return 0;
}

int repo_b200_tia_mix_fwd(const float* t_out, const float* d_out, const float* w, const float* b, float* recon, float* mask,
                          long long frames, int hw, void* stream) {
  if (!t_out || !d_out || !w || !b || !recon || !mask) return fail(-1, "tia_mix_fwd: NULL pointer");
  const long long total = frames * hw;
  if (total <= 0) return 0;
  const int blocks = (int)std::min<long long>((total + 255) / 256, (long long)std::max(1, sm_count()) * 16);
  tia_mix_fwd_kernel<<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(t_out, d_out, w, b, recon, mask, frames, hw);
  CUDA_OK(cudaGetLastError());
  return 0;
}

int repo_b200_tia_mix_bwd(const float* t_out, const float* d_out, const float* w, const float* mask, const float* g_recon,
                          float* g_t_out, float* g_d_out, float* g_wb, long long frames, int hw, void* stream) {
  if (!t_out || !d_out || !w || !mask || !g_recon || !g_t_out || !g_d_out || !g_wb) return fail(-1, "tia_mix_bwd: NULL pointer");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  CUDA_OK(cudaMemsetAsync(g_wb, 0, 7 * sizeof(float), st));
  const long long total = frames * hw;
  if (total <= 0) return 0;
  const int blocks = (int)std::min<long long>((total + 255) / 256, (long long)std::max(1, sm_count()) * 8);
  tia_mix_bwd_kernel<<<blocks, 256, 0, st>>>(t_out, d_out, w, mask, g_recon, g_t_out, g_d_out, g_wb, frames, hw);
  CUDA_OK(cudaGetLastError());
  return 0;
}

int repo_b200_grad_unshuffle(const float* g, int g_nchw, float* G, float* db, int frames, int RA, int RB, int Ho, int Wo,
                             int channels, int cpad, void* stream) {
  if (!g || !G || !db) return fail(-1, "grad_unshuffle: NULL pointer");
  if (cpad < 4 * channels || cpad > 256 || (256 % cpad)) return fail(-1, "grad_unshuffle: cpad must divide 256 and hold 4*channels");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  CUDA_OK(cudaMemsetAsync(db, 0, channels * sizeof(float), st));
  const long long rows = (long long)frames * RA * RB;
  if (rows <= 0) return 0;
  const long long total = rows * cpad;
  const int blocks = (int)std::min<long long>((total + 255) / 256, (long long)std::max(1, sm_count()) * 16);
  if (channels <= 64 && (cpad & 15) == 0 && rows < (1ll << 30) && (reinterpret_cast<uintptr_t>(g) & 15) == 0 &&
      (reinterpret_cast<uintptr_t>(G) & 15) == 0) {
    const int rows_per_pass = 256 / (cpad / 4);
    const int blocks4 = (int)std::min<long long>((rows + rows_per_pass - 1) / rows_per_pass, (long long)std::max(1, sm_count()) * 16);
    grad_unshuffle4_kernel<<<blocks4, 256, 0, st>>>(g, g_nchw, G, db, (int)rows, RA, RB, Ho, Wo, channels, cpad);
    CUDA_OK(cudaGetLastError());
    return 0;
  }
  grad_unshuffle_kernel<<<blocks, 256, 0, st>>>(g, g_nchw, G, db, rows, RA, RB, Ho, Wo, channels, cpad);
  CUDA_OK(cudaGetLastError());
  return 0;
}

int repo_b200_im2col(const float* input, float* col, int frames, const int* map, void* stream) {
  if (!input || !col || !map) return fail(-1, "im2col: NULL pointer");
  ConvMap cm;
  std::memcpy(&cm, map, sizeof(cm));
  if (cm.pix == 0) cm.pix = cm.C;
  const long long rows = (long long)frames * cm.RA * cm.RB;
  if (rows <= 0) return 0;
  const long long total = rows * cm.ntaps * cm.C;
  const int blocks = (int)std::min<long long>((total + 255) / 256, 148 * 16);
  im2col_kernel<<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(input, col, rows, cm);
  CUDA_OK(cudaGetLastError());
  return 0;
}

// ---- generic MLP on [belief | state]: forward through the layer machine, backward on the SIMT kernel
static void build_mlp(Builder& b, const repo_b200_dims* d, const repo_b200_mlp_weights* M, int out_f, int act, bool stash) {
  const int D = d->belief, S = d->state, Hd = d->hidden;
  const int kBS16 = cdiv(D + S, 16), kH16 = cdiv(Hd, 16);
  const int L = M->n_layers;
  set_dims(b.P, d);
  b.dense_to_h(M->w[0], M->b[0], D + S, Hd, 0, D + S, 0, kBS16, 0, 0, act, 0, stash ? 0 : 0xFFFF);
  for (int i = 1; i < L - 1; ++i) b.dense_to_h(M->w[i], M->b[i], Hd, Hd, 0, Hd, 0, kH16, 1, 0, act, 0, stash ? i * Hd : 0xFFFF);
  VmStage& s = b.begin_stage();
  b.gemm_rows(M->w[L - 1], Hd, 0, out_f, 0, Hd, 0, kH16, 1, 0, 0, 0);
  b.bias_rows(M->b[L - 1], 0, nullptr, 0, out_f);
  b.end_stage(s, EPI_STORE, 0, cdiv(out_f, 128), 0, out_f, 0);
}

size_t repo_b200_mlp_workspace_bytes(const repo_b200_dims* d, int n_layers, int out_features) {
  if (check_dims(d) || n_layers < 2 || n_layers > 5) return 0;
  repo_b200_mlp_weights m{};
  m.n_layers = n_layers;
  Builder b;
  build_mlp(b, d, &m, out_features, ACT_ELU, false);
  return align_up(b.packed_bytes(), 256);
}

int repo_b200_mlp_fwd(const repo_b200_dims* d, const repo_b200_mlp_weights* mlp, const float* belief, const float* state,
                      float* out, int out_features, float* stash, int n_rows, int act_kind, void* ws, size_t ws_bytes,
                      void* stream) {
  int rc = check_dims(d);
  if (rc) return rc;
  if ((rc = check_act(act_kind))) return rc;
  if (n_rows < 0 || out_features < 1 || out_features > 256) return fail(-1, "mlp: bad sizes");
  if (n_rows == 0) return 0;
  if (!mlp || mlp->n_layers < 2 || mlp->n_layers > 5) return fail(-1, "mlp: expected 2..5 layers");
  if (!belief || !state || !out) return fail(-1, "mlp: NULL pointer");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  Builder b;
  build_mlp(b, d, mlp, out_features, act_kind, stash != nullptr);
  if ((rc = b.bind_and_pack(ws, ws_bytes, true, st))) return rc;
  VmParams& P = b.P;
  P.n_steps = 1;
  P.N = n_rows;
  P.init_belief = belief; P.init_state = state;
  P.out = out; P.out_ld = out_features;
  P.stash = stash; P.stash_ld = (mlp->n_layers - 1) * d->hidden;
  return launch(P, b.max_acc_tiles, 0, st);
}

int repo_b200_mlp_bwd(const repo_b200_dims* d, const repo_b200_mlp_weights* mlp, const float* stash, const float* g_out,
                      int out_features, float* d_h1, float* d_h2, float* d_h3, float* d_h4, float* d_x, int n_rows,
                      int act_kind, void* stream) {
  int rc = check_dims(d);
  if (rc) return rc;
  if ((rc = check_act(act_kind))) return rc;
  if (n_rows < 0) return fail(-1, "mlp_bwd: bad sizes");
  if (n_rows == 0) return 0;
  if (!mlp || mlp->n_layers < 2 || mlp->n_layers > 5 || !stash || !g_out) return fail(-1, "mlp_bwd: bad arguments");
  MlpBwdParams P{};
  P.N = n_rows; P.in_f = d->belief + d->state; P.Hd = d->hidden; P.out_f = out_features; P.L = mlp->n_layers; P.act = act_kind;
  for (int i = 0; i < P.L; ++i) P.w[i] = mlp->w[i];
  P.stash = stash; P.stash_ld = (P.L - 1) * d->hidden;
  P.g_out = g_out;
  float* dh[4] = {d_h1, d_h2, d_h3, d_h4};
  for (int i = 0; i < P.L - 1; ++i) {
    if (!dh[i]) return fail(-1, "mlp_bwd: d_h%d missing", i + 1);
    P.d_h[i] = dh[i];
  }
  P.d_x = d_x;
  constexpr int RB = 8;
  const size_t smem = (size_t)2 * std::max(out_features, d->hidden) * RB * sizeof(float);
  mlp_bwd_kernel<RB><<<cdiv(n_rows, RB), 256, smem, static_cast<cudaStream_t>(stream)>>>(P);
  CUDA_OK(cudaGetLastError());
  return 0;
}

int repo_b200_tanh_normal_entropy_bwd(const float* mean, const float* std_dev, const float* eps, const float* g_entropy,
                                      float* d_mean, float* d_std, int m, int action, int samples, void* stream) {
  if (m < 0 || action < 1 || samples < 1) return fail(-1, "entropy_bwd: bad sizes");
  if (m == 0) return 0;
  if (!mean || !std_dev || !eps || !g_entropy || !d_mean || !d_std) return fail(-1, "entropy_bwd: NULL pointer");
  const long long n = (long long)m * action;
  tanh_normal_entropy_bwd_kernel<<<(unsigned)((n + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      mean, std_dev, eps, g_entropy, d_mean, d_std, m, action, samples);
  CUDA_OK(cudaGetLastError());
  return 0;
}

// ---- standalone Gaussian cells: compute_prior_state (rssm.py:42-50) / compute_posterior_state (rssm.py:52-64)
static void build_cell(Builder& b, const repo_b200_dims* d, const repo_b200_rssm_weights* W, bool posterior, int act) {
  const int D = d->belief, S = d->state, Hd = d->hidden, E = d->embed;
  const int kD16 = cdiv(D, 16), kH16 = cdiv(Hd, 16);
  set_dims(b.P, d);
  if (!posterior) {
    b.dense_to_h(W->fc_embed_belief_prior_w, W->fc_embed_belief_prior_b, D, Hd, 0, D, 0, kD16, 0, 0, act);
    b.gaussian_head(W->fc_state_prior_w, W->fc_state_prior_b, Hd, S, kH16, EPI_PRIOR, 0);
  } else {
    b.dense_to_h(W->fc_embed_belief_posterior_w, W->fc_embed_belief_posterior_b, D + E, Hd, 0, D, 0, kD16, 0, 0, act, SF_ADDEND);
    b.gaussian_head(W->fc_state_posterior_w, W->fc_state_posterior_b, Hd, S, kH16, EPI_POST, 0);
  }
}

size_t repo_b200_cell_workspace_bytes(const repo_b200_dims* d, int n_rows) {
  if (check_dims(d)) return 0;
  repo_b200_rssm_weights W{};
  Builder b;
  build_cell(b, d, &W, true, ACT_ELU);
  return align_up(b.packed_bytes(), 256) + align_up(repo_b200_linear_workspace_bytes(d->embed, d->hidden), 256) +
         align_up((size_t)std::max(n_rows, 0) * d->hidden * sizeof(float), 256);
}

int repo_b200_cell_fwd(const repo_b200_dims* d, const repo_b200_rssm_weights* W, const float* belief, const float* embed,
                       const float* eps, float* state, float* mean, float* std_dev, int n_rows, int act_kind,
                       float min_std, void* ws, size_t ws_bytes, void* stream) {
  int rc = check_dims(d);
  if (rc) return rc;
  if ((rc = check_act(act_kind))) return rc;
  if (n_rows < 0) return fail(-1, "cell: bad row count");
  if (n_rows == 0) return 0;
  if (!W || !belief || !eps || !state || !mean || !std_dev) return fail(-1, "cell: NULL pointer");
  if (ws_bytes < repo_b200_cell_workspace_bytes(d, n_rows)) return fail(-4, "cell: workspace too small");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const bool posterior = embed != nullptr;
  Builder b;
  build_cell(b, d, W, posterior, act_kind);
  const size_t main = align_up(b.packed_bytes(), 256);
  const size_t lin = align_up(repo_b200_linear_workspace_bytes(d->embed, d->hidden), 256);
  uint8_t* base = static_cast<uint8_t*>(ws);
  if ((rc = b.bind_and_pack(base, main, true, st))) return rc;
  float* addend = reinterpret_cast<float*>(base + main + lin);
  VmParams& P = b.P;
  if (posterior) {
    rc = run_linear(embed, d->embed, n_rows, d->embed, W->fc_embed_belief_posterior_w, d->belief + d->embed, d->belief,
                    nullptr, d->hidden, addend, d->hidden, base + main, lin, 0, st);
    if (rc) return rc;
    P.addend = addend;
    P.eps_post = eps; P.post_s = state; P.post_m = mean; P.post_sd = std_dev;
  } else {
    P.eps_prior = eps; P.prior_s = state; P.prior_m = mean; P.prior_sd = std_dev;
  }
  P.n_steps = 1;
  P.N = n_rows;
  P.min_std = min_std;
  P.init_belief = belief;
  return launch(P, b.max_acc_tiles, 0, st);
}

size_t repo_b200_head_workspace_bytes(const repo_b200_dims* d) {
  if (check_dims(d)) return 0;
  repo_b200_mlp_weights m{};
  Builder b;
  set_dims(b.P, d);
  b.scalar_head(&m, d->belief, d->state, d->hidden, ACT_ELU, 0);
  return align_up(b.packed_bytes(), 256);
}

int repo_b200_head_fwd(const repo_b200_dims* d, const repo_b200_mlp_weights* head, const float* belief,
                       const float* state, float* out, int n_rows, int act_kind, void* ws, size_t ws_bytes,
                       int flags, int row_tile, void* stream) {
  int rc = check_dims(d);
  if (rc) return rc;
  if ((rc = check_act(act_kind))) return rc;
  if (n_rows < 0) return fail(-1, "head: bad row count %d", n_rows);
  if (n_rows == 0) return 0;
  if (!head || head->n_layers != 4) return fail(-1, "head: expected fc1..fc4");
  if (!belief || !state || !out) return fail(-1, "head: NULL input/output pointer");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  Builder b;
  set_dims(b.P, d);
  b.scalar_head(head, d->belief, d->state, d->hidden, act_kind, 0);
  if ((rc = b.bind_and_pack(ws, ws_bytes, !(flags & REPO_B200_WEIGHTS_PACKED), st))) return rc;
  VmParams& P = b.P;
  P.n_steps = 1;
  P.N = n_rows;
  P.init_belief = belief; P.init_state = state;
  P.rewards = out;
  return launch(P, b.max_acc_tiles, row_tile, st);
}

static size_t observe_main_bytes(const repo_b200_dims* d) {
  repo_b200_rssm_weights W{};
  Builder b;
  build_observe(b, d, &W, true, ACT_ELU);
  RBuilder rb_;
  build_observe_rows(rb_, d, &W, true, ACT_ELU);
  return align_up(std::max(b.packed_bytes(), rb_.packed_bytes()), 256);
}

size_t repo_b200_observe_workspace_bytes(const repo_b200_dims* d, int t1, int batch) {
  if (check_dims(d)) return 0;
  const size_t main = observe_main_bytes(d);
  const size_t lin = align_up(repo_b200_linear_workspace_bytes(d->embed, d->hidden), 256);
  // row-major (t1 * batch, hidden), or the rows kernel's tiled layout: batch padded to 128-row tiles, hidden to 16
  const size_t addend = align_up((size_t)std::max(t1, 0) * cdiv(std::max(batch, 0), 128) * 128 * r16(d->hidden) * sizeof(float), 256);
  return main + lin + addend;
}

int repo_b200_observe_fwd(const repo_b200_dims* d, const repo_b200_rssm_weights* W, const float* prev_belief,
                          const float* prev_state, const float* actions, const float* embeds, const float* nonterm,
                          const float* eps_prior, const float* eps_post, float* beliefs, float* prior_states,
                          float* prior_means, float* prior_std_devs, float* post_states, float* post_means,
                          float* post_std_devs, float* kl, float* stash, int t1, int batch, int act_kind, float min_std,
                          void* ws, size_t ws_bytes, int flags, int row_tile, void* stream) {
  int rc = check_dims(d);
  if (rc) return rc;
  if ((rc = check_act(act_kind))) return rc;
  if (t1 < 0 || batch < 0) return fail(-1, "observe: bad sizes");
  if (t1 == 0 || batch == 0) return 0;
  if (!W || !prev_belief || !prev_state || !actions || !eps_prior || !beliefs || !prior_states || !prior_means || !prior_std_devs)
    return fail(-1, "observe: NULL input/output pointer");
  const bool with_obs = embeds != nullptr;
  if (with_obs && (!eps_post || !post_states || !post_means || !post_std_devs)) return fail(-1, "observe: posterior buffers missing");
  if (ws_bytes < repo_b200_observe_workspace_bytes(d, t1, batch)) return fail(-4, "observe: workspace too small");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  // the cluster kernel moves a thread's 4 features / 2 state dimensions with one vector access
  const bool cl_aligned = ptrs_aligned({prev_belief, beliefs, stash, ws}, 16) &&
                          ptrs_aligned({eps_prior, eps_post, prior_states, prior_means, prior_std_devs, post_states, post_means, post_std_devs}, 8);
  if (row_tile == 1 && !(cl_aligned && use_cluster_observe(d, batch, row_tile)))
    return fail(-5, "observe: the cluster kernel (row_tile 1) does not take these sizes / this alignment / this device");
  if (cl_aligned && use_cluster_observe(d, batch, row_tile)) {
    // small batches: weights resident in the shared memory of a 16-CTA cluster (cluster.cuh), stash or not
    const size_t main_c = observe_main_bytes(d);
    const size_t lin_c = align_up(repo_b200_linear_workspace_bytes(d->embed, d->hidden), 256);
    uint8_t* base_c = static_cast<uint8_t*>(ws);
    float* addend_c = reinterpret_cast<float*>(base_c + main_c + lin_c);
    if (with_obs) {
      rc = run_linear(embeds, d->embed, t1 * batch, d->embed, W->fc_embed_belief_posterior_w, d->belief + d->embed,
                      d->belief, nullptr, d->hidden, addend_c, d->hidden, base_c + main_c, lin_c, 0, st, 0);
      if (rc) return rc;
    }
    ClParams C{};
    C.T = t1; C.N = batch; C.D = d->belief; C.S = d->state; C.A = d->action; C.Hd = d->hidden;
    C.act = act_kind == REPO_B200_ACT_ELU ? ACT_ELU : ACT_RELU;
    C.with_obs = with_obs ? 1 : 0;
    C.min_std = min_std;
    C.init_belief = prev_belief; C.init_state = prev_state; C.actions = actions; C.nonterm = nonterm;
    C.addend = with_obs ? addend_c : nullptr;
    C.eps_prior = eps_prior; C.eps_post = eps_post;
    C.beliefs = beliefs; C.prior_s = prior_states; C.prior_m = prior_means; C.prior_sd = prior_std_devs;
    C.post_s = post_states; C.post_m = post_means; C.post_sd = post_std_devs;
    C.kl = with_obs ? kl : nullptr;
    C.stash = stash; C.stash_ld = 5 * d->belief + 2 * d->hidden;
    return launch_cluster_observe(d, W, C, base_c, !(flags & REPO_B200_WEIGHTS_PACKED), st);
  }
  const bool rows = !stash && use_rows_kernel(d, batch, row_tile) &&   // the activation stash is written by the vm kernel
                    rows_state_rows_aligned({eps_prior, eps_post, prior_states, prior_means, prior_std_devs, post_states, post_means, post_std_devs});
  Builder b;
  RBuilder rbld;
  const size_t main = observe_main_bytes(d);
  const size_t lin = align_up(repo_b200_linear_workspace_bytes(d->embed, d->hidden), 256);
  uint8_t* base = static_cast<uint8_t*>(ws);
  if (rows) {
    build_observe_rows(rbld, d, W, with_obs, act_kind);
    if ((rc = rbld.bind_and_pack(base, main, !(flags & REPO_B200_WEIGHTS_PACKED), st))) return rc;
  } else {
    build_observe(b, d, W, with_obs, act_kind, stash != nullptr);
    if ((rc = b.bind_and_pack(base, main, !(flags & REPO_B200_WEIGHTS_PACKED), st))) return rc;
    b.P.stash = stash;
    b.P.stash_ld = 5 * d->belief + 2 * d->hidden;
  }
  float* addend = reinterpret_cast<float*>(base + main + lin);
  bool addend_tiled = false;
  if (with_obs) {
    // hoisted, non-recurrent half of the posterior layer: all (t, b) rows in one pass
    addend_tiled = rows && linear_tiled_ok(embeds, d->embed, t1 * batch, d->embed, d->hidden, lin);
    rc = run_linear(embeds, d->embed, t1 * batch, d->embed, W->fc_embed_belief_posterior_w, d->belief + d->embed,
                    d->belief, nullptr, d->hidden, addend, d->hidden, base + main, lin, 0, st, addend_tiled ? batch : 0);
    if (rc) return rc;
  }
  VmParams& P = rows ? rbld.P.v : b.P;
  P.addend_tiled = addend_tiled ? 1 : 0;
  P.n_steps = t1;
  P.N = batch;
  P.min_std = min_std;
  P.init_belief = prev_belief; P.init_state = prev_state;
  P.actions_in = actions; P.nonterm = nonterm;
  P.addend = with_obs ? addend : nullptr;
  P.eps_prior = eps_prior; P.eps_post = eps_post;
  P.beliefs = beliefs; P.prior_s = prior_states; P.prior_m = prior_means; P.prior_sd = prior_std_devs;
  P.post_s = post_states; P.post_m = post_means; P.post_sd = post_std_devs;
  P.kl = with_obs ? kl : nullptr;
  if (rows) return launch_rows(rbld.P, st);
  if (P.kl && d->state > 32) CUDA_OK(cudaMemsetAsync(kl, 0, (size_t)t1 * batch * sizeof(float), st));
  return launch(P, b.max_acc_tiles, row_tile, st);
}

}  // extern "C"
