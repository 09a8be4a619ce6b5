"""The two update functions every algorithm in algorithms/repo runs, composed from this package's modules
(reference: `Dreamer.build_models` dreamer.py:50-139, `Dreamer.train_dynamics` :241-303, `RePo.train_dynamics`
repo.py:25-112, `Dreamer.train_actor_critic` dreamer.py:304-381).

`Agent` owns the same sub-modules under the same attribute names (`encoder`, `transition_model`, `obs_model`,
`reward_model`, `actor_model`, `value_model`, `log_beta`), groups the parameters the way the reference's three Adam
instances do, and returns the reference's `train/*` log entries as 0-d device tensors in `self.logs` (no `.item()`
syncs inside the update).  Noise can be injected (`eps_*`) for parity; otherwise it is drawn on the device.

Data parallel: every rank builds the same Agent and passes its own batch columns (`parallel.shard_rows`,
uneven shards allowed).  `config.batch_size` stays the GLOBAL batch: each rank scales its losses by
B_local / batch_size, the flat gradient buckets are SUM-all-reduced inside `FlatAdam.step` (one collective per
parameter group; `log_beta`'s 1-element gradient separately), and every rank applies the identical clip + Adam.
Logged scalars are stored pre-weighted, so a SUM-reduce (`reduced_logs`) gives the global values.

Scope: pixel observations (VisualEncoder / VisualObservationModel) — the configuration every headline run uses
(`pixel_obs=True`, train_repo.py:18).  `algo` selects the KL / reconstruction wiring:
  "dreamer": recon reads the latents (gradients reach the RSSM), KL = mean(max(kl, free_nats))
  "repo"   : recon reads DETACHED latents, KL is the dual-variable constraint with split stop-gradients
  "tia"    : task + distractor RSSMs observed on the same embeddings, two TIA decoders mixed by the mask head, a
             distractor-only decoder, adversarial distractor reward head (`TIA.build_models` tia.py:18-91,
             `TIA.train_dynamics` tia.py:93-201)
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, Optional

import math

import numpy as np
import torch
import torch.nn as nn

from . import losses
from .conv import TIAObservationModel, VisualEncoder, VisualObservationModel, tia_mix
from .models import ActorModel, EnsembleDynamicsModel, InverseDynamicsModel, RewardModel, ValueModel, bottle
from .optim import FlatAdam
from .rssm import TransitionModel


@dataclass
class Config:
    """experiments/train_repo.py:8-76 defaults (only the fields the two updates read)."""
    belief_size: int = 200
    state_size: int = 30
    hidden_size: int = 200
    embedding_size: int = 1024
    dense_activation_function: str = "elu"
    cnn_activation_function: str = "relu"
    batch_size: int = 50
    chunk_size: int = 50
    horizon: int = 15
    action_noise: float = 0.0
    gamma: float = 0.99
    gae_lambda: float = 0.95
    action_ent_coef: float = 3e-4
    latent_ent_coef: float = 0.0
    free_nats: float = 3.0
    model_lr: float = 3e-4
    actor_lr: float = 8e-5
    value_lr: float = 8e-5
    grad_clip_norm: float = 100.0
    target_kl: float = 3.0
    beta_lr: float = 1e-4
    init_beta: float = 1e-5
    prior_train_steps: int = 5
    tia_obs_coef: float = 1.0
    tia_adv_coef: float = 1.0
    tia_reward_train_steps: int = 1
    disag_model: bool = False          # dreamer.py:116-128 (optional disagreement ensemble)
    ensemble_size: int = 6
    disag_lr: float = 3e-4
    disag_coef: float = 0.0
    inv_dynamics: bool = False         # dreamer.py:130-141 (optional inverse-dynamics head)
    inv_dynamics_hidden_size: int = 512
    inv_dynamics_lr: float = 3e-4


class _Frozen:
    """FreezeParameters (common/utils.py:47-58): requires_grad False inside the block, restored after."""

    def __init__(self, modules):
        self.params = [p for m in modules for p in m.parameters()]

    def __enter__(self):
        self.old = [p.requires_grad for p in self.params]
        for p in self.params:
            p.requires_grad = False

    def __exit__(self, *exc):
        for p, o in zip(self.params, self.old):
            p.requires_grad = o


class Agent:
    def __init__(self, config: Config, action_size: int, algo: str = "repo", device="cuda"):
        if algo not in ("repo", "dreamer", "tia"):
            raise ValueError(f"algo {algo!r}: expected 'repo', 'dreamer' or 'tia'")
        c = self.c = config
        self.algo = algo
        self.device = torch.device(device)
        dev = self.device
        self.encoder = VisualEncoder(c.embedding_size, c.cnn_activation_function).to(dev)
        self.transition_model = TransitionModel(c.belief_size, c.state_size, action_size, c.hidden_size, c.embedding_size,
                                                c.dense_activation_function).to(dev)
        self.obs_model = VisualObservationModel(c.belief_size, c.state_size, c.embedding_size, c.cnn_activation_function).to(dev)
        self.reward_model = RewardModel(c.belief_size, c.state_size, c.hidden_size, c.dense_activation_function).to(dev)
        self.actor_model = ActorModel(c.belief_size, c.state_size, c.hidden_size, action_size, c.dense_activation_function).to(dev)
        self.value_model = ValueModel(c.belief_size, c.state_size, c.hidden_size, c.dense_activation_function).to(dev)
        self.log_beta = torch.tensor(np.log(c.init_beta), dtype=torch.float32, device=dev, requires_grad=True)  # repo.py:17-23
        if algo == "tia":  # tia.py:26-76
            self.obs_model = TIAObservationModel(c.belief_size, c.state_size, c.embedding_size, c.cnn_activation_function).to(dev)
            self.distractor_transition_model = TransitionModel(c.belief_size, c.state_size, action_size, c.hidden_size,
                                                               c.embedding_size, c.dense_activation_function).to(dev)
            self.distractor_obs_model = TIAObservationModel(c.belief_size, c.state_size, c.embedding_size,
                                                            c.cnn_activation_function).to(dev)
            self.distractor_only_obs_model = VisualObservationModel(c.belief_size, c.state_size, c.embedding_size,
                                                                    c.cnn_activation_function).to(dev)
            self.distractor_reward_model = RewardModel(c.belief_size, c.state_size, c.hidden_size, c.dense_activation_function).to(dev)
            self.mask_head = nn.Sequential(nn.Conv2d(6, 1, 1), nn.Sigmoid()).to(dev)  # parameter holder; runs in tia_mix
        if c.disag_model:
            self.disag_model = EnsembleDynamicsModel(c.belief_size, c.state_size, action_size, c.hidden_size, c.ensemble_size,
                                                     c.dense_activation_function).to(dev)
        if c.inv_dynamics:
            self.inv_dynamics = InverseDynamicsModel(c.belief_size, c.state_size, action_size, c.inv_dynamics_hidden_size,
                                                     c.dense_activation_function).to(dev)
        self.logs: Dict[str, torch.Tensor] = {}
        self._opt = None

    # ------------------------------------------------------------------ data parallel helpers
    @staticmethod
    def _world():
        import torch.distributed as dist
        return dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1

    def _weight(self, local_batch: int) -> float:
        """rows_local / rows_global for the means over (t, b): 1 on a single process."""
        return local_batch / self.c.batch_size if self._world() > 1 else 1.0

    def _reduce_scalar_grad(self, t: torch.Tensor):
        if self._world() > 1 and t.grad is not None:
            import torch.distributed as dist
            dist.all_reduce(t.grad, op=dist.ReduceOp.SUM)

    def reduced_logs(self) -> Dict[str, torch.Tensor]:
        """Global values of the logged scalars (one all-reduce of the stacked, pre-weighted entries)."""
        keys = sorted(self.logs)
        if self._world() == 1 or not keys:
            return dict(self.logs)
        import torch.distributed as dist
        flat = torch.stack([self.logs[k].float() for k in keys])
        dist.all_reduce(flat, op=dist.ReduceOp.SUM)
        return {k: flat[i] for i, k in enumerate(keys)}

    # parameter groups of the reference's optimisers (dreamer.py:89-96, 106, 114; repo.py:23)
    @property
    def model_params(self):
        if self.algo == "tia":  # tia.py:74-85
            mods = [self.encoder, self.transition_model, self.reward_model, self.obs_model, self.distractor_transition_model,
                    self.distractor_reward_model, self.distractor_obs_model, self.distractor_only_obs_model, self.mask_head]
        else:
            mods = [self.encoder, self.transition_model, self.obs_model, self.reward_model]
        return [p for m in mods for p in m.parameters()]

    def optimizers(self):
        """Flat-bucket clip + Adam per parameter group, created on first use (re-points the parameters)."""
        if self._opt is None:
            c = self.c
            self._opt = {
                "model": FlatAdam(self.model_params, c.model_lr, max_grad_norm=c.grad_clip_norm),
                "actor": FlatAdam(self.actor_model.parameters(), c.actor_lr, max_grad_norm=c.grad_clip_norm),
                "value": FlatAdam(self.value_model.parameters(), c.value_lr, max_grad_norm=c.grad_clip_norm),
                # RePo's dual variable (repo.py:23): the same fused Adam (no clipping) — two launches instead of the ~25
                # foreach kernels of a capturable torch.optim.Adam on one scalar (0.3 ms of a 7-sequence shard's update)
                "beta": FlatAdam([self.log_beta], c.beta_lr, max_grad_norm=None),
            }
            if c.disag_model:
                self._opt["disag"] = FlatAdam(self.disag_model.parameters(), c.disag_lr, max_grad_norm=c.grad_clip_norm)
            if c.inv_dynamics:
                self._opt["inv_dynamics"] = FlatAdam(self.inv_dynamics.parameters(), c.inv_dynamics_lr,
                                                     max_grad_norm=c.grad_clip_norm)
        return self._opt

    # ------------------------------------------------------------------ CUDA graphs
    def graphed(self, obs, actions, rewards, nonterms):
        """Capture `train_dynamics` and `train_actor_critic` (optimiser steps included) for batches shaped like the
        example; returns (wm_step, ac_step): `beliefs, states = wm_step(obs, actions, rewards, nonterms)` and
        `ac_step(beliefs.flatten(0, 1), states.flatten(0, 1))`, each ONE graph launch.  Outputs and `self.logs`
        entries are static tensors overwritten by every replay.

        Side effect: capture needs warm-up runs on a side stream, so building the graphs APPLIES four real updates on the
        example batch (three warm-up calls + the captured one: parameters, Adam moments, step counters and `log_beta`
        all move; twice that for TIA's two optimiser phases).  Pass a real batch, or snapshot `state_dict()`s around the
        call if the example must not train.  The optional Dreamer heads select rows with a boolean mask (a host
        synchronisation) and cannot be captured: `disag_model` / `inv_dynamics` raise here."""
        if self.c.disag_model or self.c.inv_dynamics:
            raise RuntimeError("Agent.graphed: the disag_model / inv_dynamics heads index by a boolean mask and cannot be "
                               "captured in a CUDA graph; run those configurations eagerly")
        from .graphs import GraphedStep
        self.optimizers()
        wm = GraphedStep(lambda o, a, r, n: self.train_dynamics(o, a, r, n), [obs, actions, rewards, nonterms])
        b, s = wm.static_outputs
        ac = GraphedStep(lambda bb, ss: self.train_actor_critic(bb, ss), [b.flatten(0, 1), s.flatten(0, 1)])
        return wm, ac

    def graphed_acting(self, obs):
        """Capture `update_latent_and_select_action` for one frame (B=1): `belief, state, action = act(belief, state,
        action, obs)` is one graph launch."""
        from .graphs import GraphedStep
        lat = list(self.init_latent_and_action())
        return GraphedStep(lambda b, s, a, o: self.update_latent_and_select_action(b, s, a, o), [*lat, obs])

    # ------------------------------------------------------------------ acting path
    def init_latent_and_action(self):
        """dreamer.py:169-173."""
        z = lambda n: torch.zeros(1, n, device=self.device)
        return z(self.c.belief_size), z(self.c.state_size), z(self.transition_model.action_size)

    @torch.no_grad()
    def update_latent_and_select_action(self, belief, posterior_state, action, obs, explore=False, *, eps_prior=None,
                                        eps_post=None):
        """dreamer.py:175-196: one posterior step on the encoded frame (T=1, B=1, no nonterminal mask), then the actor:
        the most likely of 100 samples when evaluating, one sample (+ clipped Gaussian action noise) when exploring."""
        embed = self.encoder(obs).unsqueeze(0)
        outs = self.transition_model.observe(belief, posterior_state, action.unsqueeze(0), embed, eps_prior=eps_prior, eps_post=eps_post)
        belief, posterior_state = outs[0].squeeze(0), outs[4].squeeze(0)
        action = self.actor_model.get_action(belief, posterior_state, det=not explore)
        if explore:
            action = torch.clamp(action + torch.randn_like(action) * self.c.action_noise, -1, 1)
        return belief, posterior_state, action

    # ------------------------------------------------------------------ world model
    def train_dynamics(self, obs, actions, rewards, nonterms, *, eps_prior=None, eps_post=None, step=True, **tia_eps):
        """obs (T,B,3,64,64) preprocessed floats, actions (T,B,A), rewards (T,B,1), nonterms (T,B,1).
        Returns (beliefs.detach(), posterior_states.detach()) like the reference; logs in `self.logs`."""
        if self.algo == "tia":
            return self._train_dynamics_tia(obs, actions, rewards, nonterms, eps_prior, eps_post, step, **tia_eps)
        c = self.c
        B = obs.shape[1]
        init_belief = torch.zeros(B, c.belief_size, device=self.device)
        init_state = torch.zeros(B, c.state_size, device=self.device)
        embeds = bottle(self.encoder, (obs,))
        (beliefs, prior_states, prior_means, prior_std_devs, posterior_states, posterior_means,
         posterior_std_devs) = self.transition_model.observe(init_belief, init_state, actions[:-1], embeds[1:], nonterms[:-1],
                                                             eps_prior=eps_prior, eps_post=eps_post)
        if self.algo == "repo":  # repo.py:46-48: reconstruction only probes the latents
            recon = bottle(self.obs_model, (beliefs.detach(), posterior_states.detach()))
        else:
            recon = bottle(self.obs_model, (beliefs, posterior_states))
        obs_loss = losses.normal_unit_nll(recon, obs[1:]).sum((2, 3, 4)).mean((0, 1))
        reward_loss = losses.reward_loss(bottle(self.reward_model, (beliefs, posterior_states)), rewards, nonterms)
        logs = {"train/obs_loss": obs_loss.detach(), "train/reward_loss": reward_loss.detach()}
        if self.algo == "repo":
            kl_prior = losses.kl_normal(posterior_means.detach(), posterior_std_devs.detach(), prior_means, prior_std_devs).sum(2).mean((0, 1))
            kl_post = losses.kl_normal(posterior_means, posterior_std_devs, prior_means.detach(), prior_std_devs.detach()).sum(2).mean((0, 1))
            alpha = c.prior_train_steps / (1 + c.prior_train_steps)
            kl_div = alpha * kl_prior + (1 - alpha) * kl_post
            kl_viol = kl_div - c.target_kl
            kl_loss = self.log_beta.exp().detach() * kl_viol
            beta_loss = -self.log_beta * kl_viol.detach()
            logs.update({"train/kl_div": kl_div.detach(), "train/beta": self.log_beta.exp().detach(),
                         "train/beta_loss": beta_loss.detach()})
        else:
            kl_div = losses.kl_normal(posterior_means, posterior_std_devs, prior_means, prior_std_devs).sum(2)
            kl_loss = torch.clamp(kl_div, min=c.free_nats).mean((0, 1))
            beta_loss = None
        model_loss = obs_loss + reward_loss + kl_loss
        logs.update({"train/kl_loss": kl_loss.detach(), "train/model_loss": model_loss.detach()})
        w = self._weight(B)
        opt = self.optimizers() if step else None
        if step:
            opt["model"].zero_grad()
        (model_loss * w).backward()
        if step:
            opt["model"].step()
        if beta_loss is not None:
            if step:
                opt["beta"].zero_grad()
            (beta_loss * w).backward()
            if step:
                opt["beta"].step()          # (FlatAdam SUM-all-reduces its bucket under data parallelism)
            else:
                self._reduce_scalar_grad(self.log_beta)
        if w != 1.0:  # kl_loss / beta_loss / beta carry constants: weight them too so that the SUM over ranks is global
            logs = {k: v * w for k, v in logs.items()}
        self.logs.update(logs)
        if self.algo in ("dreamer", "repo"):  # dreamer.py:297-301 and repo.py:106-110 (TIA's train_dynamics has neither head)
            if c.disag_model:
                self.train_disag(beliefs, posterior_states, actions, nonterms, step=step)
            if c.inv_dynamics:
                self.train_inv_dynamics(beliefs, posterior_states, actions, nonterms, step=step)
        return beliefs.detach(), posterior_states.detach()

    # ------------------------------------------------------------------ optional heads (dreamer.py:198-239)
    @staticmethod
    def _transition_rows(beliefs, states, actions, nonterms):
        """Rows (t, b) with nonterms[1 + t, b] == 1 of (actions[1:-1], beliefs[:-1], states[:-1], beliefs[1:]),
        detached and flattened time-major (dreamer.py:199-209 / :221-231)."""
        keep = nonterms[1:-1].flatten() == 1
        return [x.detach().flatten(0, 1)[keep] for x in (actions[1:-1], beliefs[:-1], states[:-1], beliefs[1:])]

    def _head_step(self, key, loss, local_rows, step):
        """backward + clip + Adam of one optional head; the loss is a mean over this rank's kept rows, so under data
        parallelism it is weighted kept_local / kept_global before FlatAdam's SUM all-reduce.  Returns the weight."""
        w = 1.0
        if self._world() > 1:
            import torch.distributed as dist
            total = torch.tensor([float(local_rows)], device=self.device)
            dist.all_reduce(total, op=dist.ReduceOp.SUM)
            w = local_rows / total.clamp(min=1.0)[0]
        opt = self.optimizers()[key] if step else None
        if step:
            opt.zero_grad()
        (loss * w).backward()
        if step:
            opt.step()
        return w

    def train_disag(self, beliefs, states, actions, nonterms, *, step=True):
        """dreamer.py:198-217: every ensemble member regresses the next belief, unit-variance Normal NLL summed over
        members and features, mean over the kept rows."""
        actions_in, beliefs_in, states_in, beliefs_out = self._transition_rows(beliefs, states, actions, nonterms)
        ens_preds = self.disag_model(beliefs_in, states_in, actions_in)
        disag_loss = losses.normal_unit_nll(ens_preds, beliefs_out.unsqueeze(0)).sum(2).sum(0).mean()
        w = self._head_step("disag", disag_loss, beliefs_in.shape[0], step)
        self.logs["train/disag_loss"] = disag_loss.detach() * w
        return disag_loss.detach()

    def train_inv_dynamics(self, beliefs, states, actions, nonterms, *, step=True):
        """dreamer.py:219-239: Normal(mean, std) NLL of the taken action given (belief, state, next belief)."""
        actions_in, beliefs_in, states_in, beliefs_out = self._transition_rows(beliefs, states, actions, nonterms)
        act_mean, act_std = self.inv_dynamics(beliefs_in, states_in, beliefs_out)
        nll = 0.5 * ((actions_in - act_mean) / act_std) ** 2 + act_std.log() + 0.5 * math.log(2 * math.pi)
        inv_dyn_loss = nll.sum(1).mean()
        w = self._head_step("inv_dynamics", inv_dyn_loss, beliefs_in.shape[0], step)
        self.logs["train/inv_dyn_loss"] = inv_dyn_loss.detach() * w
        return inv_dyn_loss.detach()

    def _train_dynamics_tia(self, obs, actions, rewards, nonterms, eps_prior, eps_post, step, eps_prior_d=None, eps_post_d=None):
        """tia.py:93-201."""
        c = self.c
        B = obs.shape[1]
        init_belief = torch.zeros(B, c.belief_size, device=self.device)
        init_state = torch.zeros(B, c.state_size, device=self.device)
        embeds = bottle(self.encoder, (obs,))
        (t_beliefs, _, t_prior_means, t_prior_std_devs, t_post_states, t_post_means, t_post_std_devs) = \
            self.transition_model.observe(init_belief, init_state, actions[:-1], embeds[1:], nonterms[:-1],
                                          eps_prior=eps_prior, eps_post=eps_post)
        (d_beliefs, _, d_prior_means, d_prior_std_devs, d_post_states, d_post_means, d_post_std_devs) = \
            self.distractor_transition_model.observe(init_belief, init_state, actions[:-1], embeds[1:], nonterms[:-1],
                                                     eps_prior=eps_prior_d, eps_post=eps_post_d)
        # reconstruction through the mask head (one fused kernel for sigmoid(conv1x1) and the mix)
        t_full = bottle(self.obs_model.forward_full, (t_beliefs, t_post_states))
        d_full = bottle(self.distractor_obs_model.forward_full, (d_beliefs, d_post_states))
        recon, _ = tia_mix(t_full.flatten(0, 1), d_full.flatten(0, 1), self.mask_head)
        recon = recon.reshape(obs.shape[0] - 1, B, *recon.shape[1:])
        obs_loss = losses.normal_unit_nll(recon, obs[1:]).sum((2, 3, 4)).mean((0, 1))
        d_only_recon = bottle(self.distractor_only_obs_model, (d_beliefs, d_post_states))
        d_obs_loss = losses.normal_unit_nll(d_only_recon, obs[1:]).sum((2, 3, 4)).mean((0, 1))
        # rewards: the distractor head is frozen here so that only the latents are pushed away from reward information
        rewards_tgt = rewards[:-1].squeeze(-1)
        mask = nonterms[:-1].squeeze(-1)
        t_reward = bottle(self.reward_model, (t_beliefs, t_post_states))
        with _Frozen([self.distractor_reward_model]):
            d_reward = bottle(self.distractor_reward_model, (d_beliefs, d_post_states))
        t_reward_loss = (losses.normal_unit_nll(t_reward, rewards_tgt) * mask).mean((0, 1))
        d_reward_loss = (-losses.normal_unit_nll(d_reward, rewards_tgt) * mask).mean((0, 1))
        reward_loss = t_reward_loss + c.tia_adv_coef * d_reward_loss
        t_kl_div = losses.kl_normal(t_post_means, t_post_std_devs, t_prior_means, t_prior_std_devs).sum(2)
        d_kl_div = losses.kl_normal(d_post_means, d_post_std_devs, d_prior_means, d_prior_std_devs).sum(2)
        kl_loss = torch.clamp(t_kl_div, min=c.free_nats).mean((0, 1)) + torch.clamp(d_kl_div, min=c.free_nats).mean((0, 1))
        model_loss = obs_loss + c.tia_obs_coef * d_obs_loss + reward_loss + kl_loss
        w = self._weight(B)
        opt = self.optimizers() if step else None
        if step:
            opt["model"].zero_grad()
        (model_loss * w).backward()
        if step:
            opt["model"].step()
        logs = {"train/obs_loss": obs_loss.detach(), "train/d_obs_loss": d_obs_loss.detach(), "train/reward_loss": reward_loss.detach(),
                "train/t_reward_loss": t_reward_loss.detach(), "train/kl_loss": kl_loss.detach(),
                "train/t_kl_div": t_kl_div.mean().detach(), "train/d_kl_div": d_kl_div.mean().detach(),
                "train/model_loss": model_loss.detach()}
        # second phase (tia.py:180-191): fit the distractor reward head on detached latents.  The reference runs the whole
        # model optimiser again; with the pinned torch 1.12.1 (`zero_grad` leaves zero tensors) that Adam step also moves every
        # other model parameter by its momentum, which is what the flat bucket reproduces.
        for _ in range(c.tia_reward_train_steps):
            d_reward = bottle(self.distractor_reward_model, (d_beliefs.detach(), d_post_states.detach()))
            d_reward_loss = (losses.normal_unit_nll(d_reward, rewards_tgt) * mask).mean((0, 1))
            if step:
                opt["model"].zero_grad()
                (d_reward_loss * w).backward()
                opt["model"].step()
        logs["train/d_reward_loss"] = d_reward_loss.detach()
        if w != 1.0:
            logs = {k: v * w for k, v in logs.items()}
        self.logs.update(logs)
        return t_beliefs.detach(), t_post_states.detach()

    # ------------------------------------------------------------------ actor-critic
    def train_actor_critic(self, beliefs, states, *, eps_action=None, eps_prior=None, eps_entropy=None, eps_disag=None,
                           step=True):
        """dreamer.py:304-381 on flattened start rows (N, D), (N, S)."""
        c = self.c
        opt = self.optimizers() if step else None
        with _Frozen([self.transition_model, self.reward_model]):
            imag_b, imag_s, imag_m, imag_sd = self.transition_model.imagine(beliefs, states, self.actor_model, c.horizon,
                                                                             eps_action=eps_action, eps_prior=eps_prior)
            with _Frozen([self.value_model]):
                reward_preds = bottle(self.reward_model, (imag_b, imag_s))
                value_preds = bottle(self.value_model, (imag_b, imag_s))
        action_entropy = self.actor_model.get_action_dist(imag_b.flatten(0, 1), imag_s.flatten(0, 1)).entropy(eps_entropy).mean()
        # Independent(Normal(mean, std), 1).entropy().mean() (dreamer.py:325-327), closed form: torch.distributions'
        # argument validation synchronises with the host, which a CUDA-graph capture forbids
        latent_entropy = (0.5 + 0.5 * math.log(2 * math.pi) + imag_sd.log()).sum(-1).mean()
        disag = None
        if c.disag_model and c.disag_coef > 0:  # dreamer.py:330-339: ensemble spread as an intrinsic reward
            dist_ = self.actor_model.get_action_dist(imag_b.flatten(0, 1), imag_s.flatten(0, 1))
            with _Frozen([self.disag_model]):
                ens_preds = self.disag_model(imag_b.flatten(0, 1), imag_s.flatten(0, 1), dist_.rsample(eps_disag))
            disag = ens_preds.std(0).mean(-1).reshape(reward_preds.shape)
            reward_preds = reward_preds + c.disag_coef * disag
        discounts = c.gamma * torch.ones_like(reward_preds)
        returns = losses.lambda_return(reward_preds[:-1], value_preds[:-1], discounts[:-1], value_preds[-1], c.gae_lambda)
        actor_loss = losses.actor_loss(returns, action_entropy, latent_entropy, c.action_ent_coef, c.latent_ent_coef)
        w = self._weight(beliefs.shape[0] / max(1, c.chunk_size - 1))  # start rows = (T-1) * B_local
        if step:
            opt["actor"].zero_grad()
        (actor_loss * w).backward()
        if step:
            opt["actor"].step()
        value_pred = bottle(self.value_model, (imag_b[:-1].detach(), imag_s[:-1].detach()))
        value_loss = losses.value_loss(value_pred, returns.detach())
        if step:
            opt["value"].zero_grad()
        (value_loss * w).backward()
        if step:
            opt["value"].step()
        self.logs.update({"train/actor_loss": actor_loss.detach() * w, "train/value_loss": value_loss.detach() * w,
                          "train/action_entropy": action_entropy.detach() * w, "train/latent_entropy": latent_entropy.detach() * w})
        if disag is not None:
            self.logs["train/disagreement"] = disag.detach().mean() * w
        return returns.detach()
