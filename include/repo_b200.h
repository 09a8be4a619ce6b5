/* repo_b200 — C-ABI of the B200-native RSSM hot path (librepo_b200.so).
 *
 * The reference (zchuning/repo) has no FFI for this path: the boundary it exposes is the Python
 * class algorithms/repo/models/rssm.py::TransitionModel (rssm.py:8-184) plus the helpers it
 * calls (ActorModel actor_critic.py:50-102, RewardModel decoder.py:178-195, ValueModel
 * actor_critic.py:9-26, lambda_return common/utils.py:61-71, the KL of repo.py:63-83).  Each
 * entry point below names the reference symbol whose arithmetic it replaces.  The Python side
 * (repo_b200/rssm.py) binds these with ctypes and keeps the reference's module interface.
 *
 * Conventions: every pointer is a DEVICE pointer to contiguous fp32, row-major, time-major
 * (T, rows, feature) memory owned by the caller; `stream` is a cudaStream_t; no call
 * synchronises or allocates; the caller provides the workspace (size from *_workspace_bytes).
 * Return 0 on success, negative on error; repo_b200_last_error() gives the message
 * (thread-local).  There is no CPU fallback: without a CUDA device every call fails.
 */
#ifndef REPO_B200_H
#define REPO_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define REPO_B200_ACT_RELU 0
#define REPO_B200_ACT_ELU 1

/* flags */
#define REPO_B200_WEIGHTS_PACKED 1 /* workspace already holds the packed weights of these exact tensors */

typedef struct repo_b200_dims {
  int belief, state, action, hidden, embed; /* TransitionModel ctor sizes, rssm.py:9-18 */
} repo_b200_dims;

/* TransitionModel.state_dict() (rssm.py:21-32): nn.Linear weights are [out, in]; nn.GRUCell gate
 * order r,z,n. */
typedef struct repo_b200_rssm_weights {
  const float *fc_embed_state_action_w, *fc_embed_state_action_b;         /* (D, S+A), (D)   */
  const float *rnn_w_ih, *rnn_w_hh, *rnn_b_ih, *rnn_b_hh;                 /* (3D, D) x2, (3D) x2 */
  const float *fc_embed_belief_prior_w, *fc_embed_belief_prior_b;         /* (H, D), (H)     */
  const float *fc_state_prior_w, *fc_state_prior_b;                       /* (2S, H), (2S)   */
  const float *fc_embed_belief_posterior_w, *fc_embed_belief_posterior_b; /* (H, D+E), (H)   */
  const float *fc_state_posterior_w, *fc_state_posterior_b;               /* (2S, H), (2S)   */
} repo_b200_rssm_weights;

/* fc1..fcN of ActorModel (N=5, out 2A), RewardModel / ValueModel (N=4, out 1); input = [belief|state]. */
typedef struct repo_b200_mlp_weights {
  const float* w[5];
  const float* b[5];
  int n_layers;
} repo_b200_mlp_weights;

int repo_b200_version(void);
const char* repo_b200_last_error(void);
/* number of SMs / compute capability of the current device; <0 if no device */
int repo_b200_device_info(int* sm_count, int* cc_major, int* cc_minor);
/* bring-up only: bit0 swaps LBO/SBO in the tcgen05 shared-memory descriptors */
void repo_b200_debug_flags(int flags);
/* profiling only: CTA 0 of the rows kernel writes [step][stage][2] clock64 stamps into this device buffer (NULL = off) */
void repo_b200_debug_clock(void* device_buffer);

/* ---- imagine: TransitionModel.imagine (rssm.py:148-184) with policy = ActorModel.get_action
 * (actor_critic.py:97-102), fused with RewardModel / ValueModel on every imagined state
 * (dreamer.py:315-317) and lambda_return (common/utils.py:61-71 as called at dreamer.py:342-349).
 *   start_belief (N,D), start_state (N,S); eps_action (H-1,N,A), eps_prior (H-1,N,S) standard normal;
 *   outputs beliefs (H-1,N,D), prior_* (H-1,N,S), actions (H-1,N,A, nullable), rewards/values (H-1,N),
 *   returns (H-2,N).  reward/value may be NULL (then rewards/values/returns are not written).
 *   row_tile: 0 = auto, else 16/32/64 rows per CTA. */
size_t repo_b200_imagine_workspace_bytes(const repo_b200_dims* dims);
int repo_b200_imagine_fwd(const repo_b200_dims* dims, const repo_b200_rssm_weights* rssm,
                          const repo_b200_mlp_weights* actor, const repo_b200_mlp_weights* reward,
                          const repo_b200_mlp_weights* value, const float* start_belief, const float* start_state,
                          const float* eps_action, const float* eps_prior, float* beliefs, float* prior_states,
                          float* prior_means, float* prior_std_devs, float* actions, float* rewards, float* values,
                          float* returns, int horizon, int n_rows, int act_kind, float min_std_dev,
                          float actor_mean_scale, float actor_init_std, float actor_min_std, float gamma,
                          float lambda_, float* stash, void* workspace, size_t workspace_bytes, int flags,
                          int row_tile, void* stream);

/* ---- imagine backward: what autograd derives from rssm.py:167-176 + actor_critic.py:76-102 in the reference,
 * with the actor's inputs detached (rssm.py:170).  `stash` (H-1,N,repo_b200_imagine_stash_floats) is written by
 * repo_b200_imagine_fwd when non-NULL: per (t,row) [embed hidden D][r D][z D][n D][h_n D][prior hidden H]
 * [actor h1..h4 4H][action mean A][action std A].  g_*: incoming gradients of the four outputs (nullable).
 * Outputs: pre-activation gradients of the transition layers (d_p (.,2S), d_hp (.,H), d_gi/d_gh (.,3D), d_e (.,D))
 * and of the actor's fc5..fc1 (d_a5 (.,2A), d_a4..d_a1 (.,H)), plus gradients of the start rows (nullable).
 * d_a4..d_a1 may all be NULL: the actor's inputs are detached (rssm.py:170), so nothing in the recurrence reads its hidden
 * layers' gradients and the caller may compute them after the time loop from d_a5 (dense GEMMs over all (t,row)). */
/* ---- conditional (multitask) variants: ConditionalTransitionModel.imagine (rssm.py:225-248) with a
 * ConditionalActorModel (actor_critic.py:105-148).  dims.action is the pseudo-action width (action + condition, what the
 * reference passes to the base class, rssm.py:198-206); `condition` is (n_rows, cond_size), constant over the horizon;
 * actor fc1 has belief + state + cond_size input columns ([belief | state | condition], actor_critic.py:131-133);
 * eps_action / actions / d_a5 use the sampled width dims.action - cond_size.  cond_size = 0 is the plain call. */
int repo_b200_imagine_cond_fwd(const repo_b200_dims* d, const repo_b200_rssm_weights* W, const repo_b200_mlp_weights* actor,
                               const repo_b200_mlp_weights* reward, const repo_b200_mlp_weights* value,
                               const float* start_belief, const float* start_state, const float* condition, int cond_size,
                               const float* eps_action, const float* eps_prior, float* beliefs, float* prior_states,
                               float* prior_means, float* prior_std_devs, float* actions, float* rewards, float* values,
                               float* returns, int horizon, int n_rows, int act_kind, float min_std, float a_mean_scale,
                               float a_init_std, float a_min_std, float gamma, float lambda_, float* stash, void* workspace,
                               size_t workspace_bytes, int flags, int row_tile, void* stream);
int repo_b200_imagine_cond_bwd(const repo_b200_dims* d, const repo_b200_rssm_weights* W, const repo_b200_mlp_weights* actor,
                               int cond_size, const float* start_belief, const float* beliefs, const float* actions,
                               const float* prior_std_devs, const float* eps_prior, const float* eps_action,
                               const float* stash, const float* g_beliefs, const float* g_prior_states,
                               const float* g_prior_means, const float* g_prior_std_devs, float* d_p, float* d_hp, float* d_gi,
                               float* d_gh, float* d_e, float* d_a5, float* d_a4, float* d_a3, float* d_a2, float* d_a1,
                               float* d_start_belief, float* d_start_state, int horizon, int n_rows, int act_kind,
                               float min_std, float a_mean_scale, float a_min_std, void* stream);

int repo_b200_imagine_stash_floats(const repo_b200_dims* dims);
int repo_b200_imagine_bwd(const repo_b200_dims* dims, const repo_b200_rssm_weights* rssm,
                          const repo_b200_mlp_weights* actor, const float* start_belief, const float* beliefs,
                          const float* actions, const float* prior_std_devs, const float* eps_prior,
                          const float* eps_action, const float* stash, const float* g_beliefs,
                          const float* g_prior_states, const float* g_prior_means, const float* g_prior_std_devs,
                          float* d_p, float* d_hp, float* d_gi, float* d_gh, float* d_e, float* d_a5, float* d_a4,
                          float* d_a3, float* d_a2, float* d_a1, float* d_start_belief, float* d_start_state,
                          int horizon, int n_rows, int act_kind, float min_std_dev, float actor_mean_scale,
                          float actor_min_std, void* stream);

/* ---- observe: TransitionModel.observe (rssm.py:76-146) fused with the per-(t,b) Gaussian KL
 * KL(posterior || prior).sum(state) used by dreamer.py:278-282 / repo.py:63-83.
 *   prev_belief (B,D), prev_state (B,S), actions (T1,B,A), embeds (T1,B,E) or NULL (prior-only
 *   rollout), nonterminals (T1,B) or NULL, eps_prior/eps_post (T1,B,S).
 *   outputs: beliefs (T1,B,D), prior_states/means/std_devs, posterior_states/means/std_devs
 *   (T1,B,S) (posterior_* ignored when embeds == NULL), kl (T1,B) nullable. */
size_t repo_b200_observe_workspace_bytes(const repo_b200_dims* dims, int t1, int batch);
int repo_b200_observe_fwd(const repo_b200_dims* dims, const repo_b200_rssm_weights* rssm, const float* prev_belief,
                          const float* prev_state, const float* actions, const float* embeds,
                          const float* nonterminals, const float* eps_prior, const float* eps_post, float* beliefs,
                          float* prior_states, float* prior_means, float* prior_std_devs, float* posterior_states,
                          float* posterior_means, float* posterior_std_devs, float* kl, float* stash, int t1,
                          int batch, int act_kind, float min_std_dev, void* workspace, size_t workspace_bytes,
                          int flags, int row_tile, void* stream);

/* ---- observe backward (BPTT): what autograd derives from rssm.py:116-133 in the reference.
 * `stash` (T1,B,repo_b200_observe_stash_floats) is written by repo_b200_observe_fwd when non-NULL:
 * per (t,b) [embed hidden D][r D][z D][n D][W_hn h + b_hn D][prior hidden H][posterior hidden H].
 * Inputs: the forward tensors and the incoming gradients g_* of the seven outputs (each nullable).
 * Outputs: the gradient of every pre-activation per (t,b) — d_q/d_p (T1,B,2S) for the posterior/prior
 * [mean | raw std], d_hq/d_hp (T1,B,H), d_gi/d_gh (T1,B,3D) for the GRU's W_ih x + b_ih / W_hh h + b_hh,
 * d_e (T1,B,D) — plus d_prev_belief (B,D) / d_prev_state (B,S) (nullable).  Weight gradients are plain
 * GEMMs of these against the stashed layer inputs (repo_b200/autograd.py). */
int repo_b200_observe_stash_floats(const repo_b200_dims* dims);
int repo_b200_observe_bwd(const repo_b200_dims* dims, const repo_b200_rssm_weights* rssm, const float* prev_belief,
                          const float* beliefs, const float* prior_std_devs, const float* posterior_std_devs,
                          const float* eps_prior, const float* eps_post, const float* nonterminals, const float* stash,
                          const float* g_beliefs, const float* g_prior_states, const float* g_prior_means,
                          const float* g_prior_std_devs, const float* g_posterior_states,
                          const float* g_posterior_means, const float* g_posterior_std_devs, float* d_q, float* d_hq,
                          float* d_p, float* d_hp, float* d_gi, float* d_gh, float* d_e, float* d_prev_belief,
                          float* d_prev_state, int t1, int batch, int with_obs, int act_kind, float min_std_dev,
                          void* stream);
/* Same pass with a caller-provided workspace (>= repo_b200_observe_bwd_workspace_bytes, 16-byte aligned), which lets small
 * batches run on the cluster kernel (csrc/cluster_bwd.cuh: transposed weight slices resident in the shared memory of a
 * 16-CTA cluster, every dx = dy W on the tensor cores, per-sequence power-of-two units).  mode: 0 = auto (cluster kernel when
 * it takes the sizes and alignments, else the kernel above), 1 = cluster kernel or an error, 2 = the kernel above. */
size_t repo_b200_observe_bwd_workspace_bytes(const repo_b200_dims* dims, int batch);
int repo_b200_observe_bwd_ws(const repo_b200_dims* dims, const repo_b200_rssm_weights* rssm, const float* prev_belief,
                             const float* beliefs, const float* prior_std_devs, const float* posterior_std_devs,
                             const float* eps_prior, const float* eps_post, const float* nonterminals, const float* stash,
                             const float* g_beliefs, const float* g_prior_states, const float* g_prior_means,
                             const float* g_prior_std_devs, const float* g_posterior_states,
                             const float* g_posterior_means, const float* g_posterior_std_devs, float* d_q, float* d_hq,
                             float* d_p, float* d_hp, float* d_gi, float* d_gh, float* d_e, float* d_prev_belief,
                             float* d_prev_state, int t1, int batch, int with_obs, int act_kind, float min_std_dev,
                             void* workspace, size_t workspace_bytes, int mode, void* stream);

/* ---- cell: TransitionModel.compute_prior_state (rssm.py:42-50; embed == NULL) or
 * compute_posterior_state (rssm.py:52-64; embed (N,E)).  belief (N,D), eps (N,S) -> state, mean, std_dev (N,S). */
size_t repo_b200_cell_workspace_bytes(const repo_b200_dims* dims, int n_rows);
int repo_b200_cell_fwd(const repo_b200_dims* dims, const repo_b200_rssm_weights* rssm, const float* belief,
                       const float* embed, const float* eps, float* state, float* mean, float* std_dev, int n_rows,
                       int act_kind, float min_std_dev, void* workspace, size_t workspace_bytes, void* stream);

/* ---- head: RewardModel.forward (decoder.py:189-195) / ValueModel.forward (actor_critic.py:20-26):
 * [belief|state] -> 3 x (Linear + act) -> Linear(1), squeezed.  belief (N,D), state (N,S), out (N). */
size_t repo_b200_head_workspace_bytes(const repo_b200_dims* dims);
int repo_b200_head_fwd(const repo_b200_dims* dims, const repo_b200_mlp_weights* head, const float* belief,
                       const float* state, float* out, int n_rows, int act_kind, void* workspace,
                       size_t workspace_bytes, int flags, int row_tile, void* stream);

/* ---- mlp: fc1..fcL on [belief|state] with L in 2..5 — RewardModel / ValueModel (decoder.py:178-195,
 * actor_critic.py:9-26; L=4, out 1) and ActorModel.forward's trunk (actor_critic.py:76-82; L=5, out 2A).
 * fwd: out (N,out_features); stash (N,(L-1)*H) nullable receives the hidden activations.
 * bwd: g_out (N,out_features) -> pre-activation gradients d_h1..d_h{L-1} (N,H) and d_x (N,D+S, nullable);
 * weight gradients are GEMMs of those against the stashed inputs (repo_b200/autograd.py). */
size_t repo_b200_mlp_workspace_bytes(const repo_b200_dims* dims, int n_layers, int out_features);
int repo_b200_mlp_fwd(const repo_b200_dims* dims, const repo_b200_mlp_weights* mlp, const float* belief,
                      const float* state, float* out, int out_features, float* stash, int n_rows, int act_kind,
                      void* workspace, size_t workspace_bytes, void* stream);
int repo_b200_mlp_bwd(const repo_b200_dims* dims, const repo_b200_mlp_weights* mlp, const float* stash,
                      const float* g_out, int out_features, float* d_h1, float* d_h2, float* d_h3, float* d_h4,
                      float* d_x, int n_rows, int act_kind, void* stream);

/* ---- MC entropy of the tanh-Normal policy: SampleDist.entropy (models/utils.py:160-163) over
 * Independent(TransformedDistribution(Normal(mean,std), TanhBijector), 1) (actor_critic.py:89-95;
 * TanhBijector models/utils.py:112-134), called at dreamer.py:320-324.
 *   mean, std_dev (M,A); eps (K,M,A) standard normal draws; entropy (M). */
int repo_b200_tanh_normal_entropy_fwd(const float* mean, const float* std_dev, const float* eps, float* entropy, int m,
                                      int action, int samples, void* stream);

int repo_b200_tanh_normal_entropy_bwd(const float* mean, const float* std_dev, const float* eps,
                                      const float* g_entropy, float* d_mean, float* d_std, int m, int action,
                                      int samples, void* stream);

/* ---- conv_gemm: one Conv2d / ConvTranspose2d layer as an implicit GEMM on tcgen05 (VisualEncoder encoder.py:21-41,
 * VisualObservationModel decoder.py:28-48), also used for the data gradients of those layers.  `map` is the 28-int
 * ConvMap of repo_b200/csrc/vm.cuh (row grid, input layout/dims, tap window, input/output pixel maps, relu, shuffle);
 * w_mat is (n_total, ntaps*C) with columns ordered (tap, cin); with map.shuffle the n_total = 4*cout features are the
 * (py, px, cout) sub-pixel classes of a stride-2 transposed convolution.  relu_mask (nullable, laid out like out)
 * zeroes outputs where mask <= 0.  scales (nullable) = device floats [s_x, s_w, 1/(s_x*s_w)]: input and weights are
 * multiplied by s_x / s_w before the fp16 hi/lo split and the accumulator by the third entry (powers of two that keep
 * small-magnitude gradients inside fp16's normal range).  workspace >= repo_b200_conv_workspace_bytes(ntaps*C, n_total).
 * Built by repo_b200/conv.py. */
size_t repo_b200_conv_workspace_bytes(int k, int n_total);
/* hl_flags: operands in split-activation format — a tensor stored as two fp16 planes, hi then lo (x = hi + lo, the
 * exact operand pair the tensor cores consume), each laid out like the fp32 tensor, lo plane directly after the hi
 * plane.  bit0: input, bit1: output, bit2: relu_mask.  Used for the activations between layers: the producing
 * epilogue splits once and every consumer (next layer, weight gradient) gathers with plain 16-byte copies. */
/* dense_opts (nullable, 4 ints) turns the same kernel into the dense layers of the MLP heads (actor_critic.py:20-26,76-83,
 * decoder.py:189-195) and their data gradients: [0] apply ELU to the output (instead of map.relu), [1] treat relu_mask
 * as the OUTPUT h of an ELU layer and multiply by its derivative (h > 0 ? 1 : h + 1), [2] row stride of `out`, [3] row
 * stride of `relu_mask` (plain GEMM maps only; lets both be column windows of a wider row-major activation stash). */
int repo_b200_conv_gemm(const void* input, const float* w_mat, const float* bias, const void* relu_mask,
                        const float* scales, void* out, int frames, int n_total, const int* map, int hl_flags,
                        const int* dense_opts, void* workspace, size_t workspace_bytes, void* stream);

/* weight gradient of the same implicit GEMM: dw (n_total, ntaps*C) = grad_rows^T @ gather(input), grad_rows being the
 * (frames*RA*RB, g_ld) output-gradient rows (n_total <= 256 columns used).  Row slices are summed with fp32 atomics
 * (dw is zeroed by the call).  scales (nullable) = [s_input, s_grad, 1/(s_input*s_grad)] as in conv_gemm; input_hl:
 * the input is in split-activation format (then s_input must be 1). */
int repo_b200_conv_wgrad(const void* input, const float* grad_rows, const float* scales, float* dw, int frames,
                         int n_total, int g_ld, const int* map, int input_hl, void* stream);

/* scales[which] = 2^floor(log2(target / max|x|)) and scales[2] = 1 / (scales[0] * scales[1]) in one read-only pass
 * (the `scales` triple of conv_gemm / conv_wgrad; initialise scales[0..2] to {1, 1, 1}); scales[3..5] receive the same
 * triple with the two operand slots swapped, so `scales` must hold 6 floats.  scratch = 8 zeroed device bytes,
 * left zeroed again.  No host synchronisation. */
int repo_b200_pow2_scale(const float* x, long long n, float target, int which, float* scales, void* scratch, void* stream);

/* ---- TIA mask mixing (tia.py:72,124-127): t_out / d_out are the (frames,6,H,W) outputs of the task / distractor
 * TIAObservationModel (decoder.py:154-175; channels 0-2 reconstruction, 3-5 mask features), w (6) and b (1) the 1x1
 * mask-head convolution.  mask = sigmoid(w . [t_mask | d_mask] + b) (frames,H,W); recon = t*mask + d*(1-mask)
 * (frames,3,H,W).  hw = H*W.  The backward writes both (frames,6,H,W) input gradients and g_wb = [d_w(6), d_b]. */
int repo_b200_tia_mix_fwd(const float* t_out, const float* d_out, const float* w, const float* b, float* recon, float* mask,
                          long long frames, int hw, void* stream);
int repo_b200_tia_mix_bwd(const float* t_out, const float* d_out, const float* w, const float* mask, const float* g_recon,
                          float* g_t_out, float* g_d_out, float* g_wb, long long frames, int hw, void* stream);

/* backward glue of a stride-2 ConvTranspose2d: G (frames,RA,RB,cpad) with G[(f,a,b),(py,px,c)] = g[f,2a+py,2b+px,c]
 * (zero outside the Ho x Wo grid and in the padding columns; g is NHWC, or NCHW when g_nchw) and the bias gradient
 * db[c] = sum of g over frames and pixels, in one pass.  cpad >= 4*channels must divide 256. */
int repo_b200_grad_unshuffle(const float* g, int g_nchw, float* G, float* db, int frames, int RA, int RB, int Ho, int Wo,
                             int channels, int cpad, void* stream);

/* backward helper with the same map: im2col materialises the gathered rows (rows = frames*RA*RB, ntaps*C columns)
 * for the weight-gradient GEMM. */
int repo_b200_im2col(const float* input, float* col, int frames, const int* map, void* stream);

/* ---- optimiser tail over one flat fp32 bucket: nn.utils.clip_grad_norm_ + Adam.step as the trainers call them
 * (dreamer.py:286-289, 356-359, 370-373; repo.py:87-96), torch defaults (no weight decay, no amsgrad).
 * repo_b200_sqnorm_accumulate adds sum(grad^2) to *sqnorm (device scalar, zero it first);
 * repo_b200_adam_clip_step scales grad by min(1, max_norm / (sqrt(*sqnorm) + 1e-6)) (sqnorm NULL or max_norm <= 0:
 * no clipping) and applies the bias-corrected Adam update of 1-based `step`.  No host synchronisation. */
int repo_b200_sqnorm_accumulate(const float* grad, long long n, float* sqnorm, void* stream);
/* bias gradient of a linear / conv layer, i.e. autograd's grad_output.sum(0) behind nn.Linear / nn.Conv2d in the reference
 * (dreamer.py:286, 356, 370: loss.backward()): out[c] = sum_r x[r * ld + c] for a (rows, cols) row-major window with row
 * stride ld.  Partial sums of row slices meet through fp32 atomics (summation order varies from run to run). */
int repo_b200_colsum(const float* x, long long rows, int cols, long long ld, float* out, void* stream);
int repo_b200_adam_clip_step(float* param, float* grad, float* exp_avg, float* exp_avg_sq, long long n,
                             const float* sqnorm, float max_norm, float lr, float beta1, float beta2, float eps,
                             int step, void* stream);
/* same update with the step count kept on the device (*step_dev is incremented, then used for the bias corrections):
 * safe to capture in a CUDA graph and replay. */
int repo_b200_adam_clip_step_dev(float* param, float* grad, float* exp_avg, float* exp_avg_sq, long long n,
                                 const float* sqnorm, float max_norm, float lr, float beta1, float beta2, float eps,
                                 int* step_dev, void* stream);

/* ---- replay gather: the index bookkeeping of SequenceReplayBuffer.sample (common/buffers.py:156-166) after
 * `np.random.choice` (start_inds, drawn on the host so the RNG stream is the reference's), fused with
 * preprocess (common/utils.py:74-80: x/255*2-1 in numpy's operation order) and nonterms = 1 - dones
 * (dreamer.py:391).  Ring arrays on the device: obs (cap, frame_bytes) uint8, actions (cap, action_dim),
 * rewards (cap), dones (cap); outputs time-major (L, B, ...); index_out (L*B) nullable. */
int repo_b200_replay_gather(const uint8_t* obs, const float* actions, const float* rewards, const float* dones,
                            const long long* start_inds, int batch, int seq_len, long long pos, int full,
                            long long length, int frame_bytes, int action_dim, float* obs_out, float* actions_out,
                            float* rewards_out, float* nonterminals_out, long long* index_out, void* stream);

/* ---- linear: y = x W^T + b on the same machine (nn.Linear as used by rssm.py:23-32); building block
 * and bring-up test.  x (rows, in_f) ld = x_ld; w (out_f, in_f); b (out_f) nullable; y ld = y_ld. */
size_t repo_b200_linear_workspace_bytes(int in_features, int out_features);
int repo_b200_linear_fwd(const float* x, int x_ld, int rows, int in_features, const float* w, const float* b,
                         int out_features, float* y, int y_ld, void* workspace, size_t workspace_bytes,
                         int row_tile, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* REPO_B200_H */
