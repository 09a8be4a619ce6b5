"""Drop-in `TransitionModel` (reference: algorithms/repo/models/rssm.py:8-184).

Same constructor, same submodule / parameter names (so `state_dict()` round-trips with the
reference's checkpoints, dreamer.py:501-547) and the same methods; the time loops run inside one
persistent sm_100a kernel launch (repo_b200/csrc/vm.cuh) instead of ~50 Python iterations of
~45 ATen dispatches.  `obs_step` / `img_step` are the north_star's names for the cell methods.

Noise: the reference draws `torch.randn_like` inline (rssm.py:49,62) and `Normal.rsample` inside
the policy (actor_critic.py:97-102).  Here the same standard-normal tensors are drawn up front with
`torch.randn` on the device (optionally in the reference's exact per-step order, `rng_compat`),
or injected by the caller (`eps_*=` keyword arguments) — that is how parity tests feed identical
noise to the reference, the oracle and this module.
"""
from __future__ import annotations

from typing import Dict, List, Optional

import torch
import torch.nn as nn

from . import ops


def _named(module: nn.Module) -> Dict[str, torch.Tensor]:
    return {k: v for k, v in module.named_parameters()}


class TransitionModel(nn.Module):
    def __init__(self, belief_size, state_size, action_size, hidden_size, embedding_size,
                 activation_function="relu", min_std_dev=0.1):
        super().__init__()
        ops.act_kind(activation_function)  # raise early on unsupported activations
        self.activation_function = activation_function
        self.min_std_dev = min_std_dev
        self.belief_size, self.state_size, self.action_size = belief_size, state_size, action_size
        self.hidden_size, self.embedding_size = hidden_size, embedding_size
        # parameter holders: identical names, shapes, init order and init distribution as rssm.py:21-32
        self.fc_embed_state_action = nn.Linear(state_size + action_size, belief_size)
        self.rnn = nn.GRUCell(belief_size, belief_size)
        self.fc_embed_belief_prior = nn.Linear(belief_size, hidden_size)
        self.fc_state_prior = nn.Linear(hidden_size, 2 * state_size)
        self.fc_embed_belief_posterior = nn.Linear(belief_size + embedding_size, hidden_size)
        self.fc_state_posterior = nn.Linear(hidden_size, 2 * state_size)
        # draw noise step by step in the reference's order (bit-identical RNG consumption on the
        # same device/seed) instead of one batched draw per call
        self.rng_compat = False
        self._ws: Dict[str, torch.Tensor] = {}
        self.last_kl: Optional[torch.Tensor] = None

    # ------------------------------------------------------------------ noise
    def _randn(self, T: int, B: int, F: int, like: torch.Tensor) -> torch.Tensor:
        return torch.randn(T, B, F, device=like.device, dtype=torch.float32)

    def _observe_noise(self, T1, B, like, with_obs):
        S = self.state_size
        if not self.rng_compat:
            return self._randn(T1, B, S, like), (self._randn(T1, B, S, like) if with_obs else None)
        pri, post = [], []
        for _ in range(T1):  # rssm.py:121-132: prior draw then posterior draw, every step
            pri.append(torch.randn(B, S, device=like.device))
            if with_obs:
                post.append(torch.randn(B, S, device=like.device))
        return torch.stack(pri), (torch.stack(post) if with_obs else None)

    # `prefetch_noise = True` (off by default): when imagine() draws its own noise, the draw for the NEXT call of the same shape
    # is enqueued on a side stream right after this call's kernel launch, so it runs under that kernel instead of in front of
    # the next one (38 M normals per 75,776 x 14 launch: ~0.25 ms of a 4.3 ms call).  The random stream is consumed in the same
    # order; the only difference is one unused draw left behind by the last call.
    prefetch_noise = False

    def _imagine_noise(self, T, N, like, action_width=None):
        S, A = self.state_size, (self.action_size if action_width is None else action_width)
        if self.prefetch_noise and not self.rng_compat and like.is_cuda:
            key = (T, N, A, S, like.device)
            cur = torch.cuda.current_stream(like.device)
            nxt = getattr(self, "_noise_next", None)
            if nxt is not None and nxt[0] == key:
                _, ea, ep, ev = nxt
                cur.wait_event(ev)
                ea.record_stream(cur)
                ep.record_stream(cur)
            else:
                ea, ep = self._randn(T, N, A, like), self._randn(T, N, S, like)
            self._noise_key = key
            return ea, ep
        if not self.rng_compat:
            return self._randn(T, N, A, like), self._randn(T, N, S, like)
        ea, ep = [], []
        for _ in range(T):  # rssm.py:170-176: action rsample then prior randn_like
            ea.append(torch.randn(N, A, device=like.device))
            ep.append(torch.randn(N, S, device=like.device))
        return torch.stack(ea), torch.stack(ep)

    # ------------------------------------------------------------------ reference API
    def observe(self, prev_belief: torch.Tensor, prev_state: torch.Tensor, actions: torch.Tensor,
                observations: Optional[torch.Tensor] = None, nonterminals: Optional[torch.Tensor] = None,
                *, eps_prior: Optional[torch.Tensor] = None, eps_post: Optional[torch.Tensor] = None) -> List[torch.Tensor]:
        """rssm.py:76-146.  Returns [beliefs, prior_states, prior_means, prior_std_devs,
        posterior_states, posterior_means, posterior_std_devs] (first four when observations is None),
        each with T-1 leading entries.  The per-(t,b) KL(posterior||prior) computed in the same launch is
        left in `self.last_kl` (T-1, B)."""
        T1, B = actions.shape[0], actions.shape[1]
        with_obs = observations is not None
        if eps_prior is None:
            eps_prior, eps_post_d = self._observe_noise(T1, B, actions, with_obs)
            if eps_post is None:
                eps_post = eps_post_d
        elif with_obs and eps_post is None:
            raise ValueError("eps_post must be given together with eps_prior when observations are passed")
        if self._wants_grad([prev_belief, prev_state, observations]):
            from . import autograd as _ag  # forward kernel + stash, hand-written BPTT kernel in backward
            return _ag.observe(self, prev_belief, prev_state, actions, observations, nonterminals, eps_prior, eps_post)
        outs, kl, ws = ops.observe_fwd(_named(self), prev_belief, prev_state, actions, observations, nonterminals,
                                       eps_prior, eps_post, act=self.activation_function, min_std=self.min_std_dev,
                                       workspace=self._ws.get("observe"))
        self._ws["observe"] = ws
        self.last_kl = kl
        return outs

    def imagine(self, prev_belief, prev_state, policy, horizon, *, eps_action=None, eps_prior=None,
                reward_model=None, value_model=None, gamma=0.99, lambda_=0.95, return_extras=False, _cond=None):
        """rssm.py:148-184.  `policy` must be an ActorModel-like module (fc1..fc5, `_mean_scale`, `_init_std`,
        `_min_std`: actor_critic.py:50-74); anything else raises (no fallback).  With `reward_model` /
        `value_model` (fc1..fc4) the same launch also produces rewards, values and lambda-returns
        (`return_extras=True` returns them as a dict after the reference's 4-list)."""
        for need in ("fc1", "fc2", "fc3", "fc4", "fc5", "_mean_scale", "_init_std", "_min_std"):
            if not hasattr(policy, need):
                raise TypeError(f"imagine: policy {type(policy).__name__} lacks {need!r}; only ActorModel-style "
                                "tanh-Normal policies are supported by the fused kernel")
        if getattr(policy, "_dist", "tanh_normal") not in ("tanh_normal", "elu", "relu"):
            # the trainers pass the activation name in the `dist` slot (dreamer.py:99-105); any value ends
            # up as a tanh-Normal policy in the reference, so only reject things we cannot interpret
            raise TypeError(f"imagine: unsupported policy dist {policy._dist!r}")
        N, T = prev_belief.shape[0], horizon - 1
        csz = 0 if _cond is None else _cond.shape[1]
        if eps_action is None or eps_prior is None:
            ea, ep = self._imagine_noise(T, N, prev_belief, self.action_size - csz)
            eps_action = ea if eps_action is None else eps_action
            eps_prior = ep if eps_prior is None else eps_prior
        if self._wants_grad([prev_belief, prev_state]) or (torch.is_grad_enabled() and any(p.requires_grad for p in policy.parameters())):
            if reward_model is not None or value_model is not None:
                raise NotImplementedError("imagine under autograd returns the reference's four lists; evaluate the heads on them")
            from . import autograd as _ag
            traj, actions = _ag.imagine(self, prev_belief, prev_state, policy, horizon, eps_action, eps_prior, _cond)
            return (traj, {"actions": actions}) if return_extras else traj
        out = ops.imagine_fwd(_named(self), _named(policy),
                              _named(reward_model) if reward_model is not None else None,
                              _named(value_model) if value_model is not None else None,
                              prev_belief, prev_state, eps_action, eps_prior, horizon,
                              act=self.activation_function, min_std=self.min_std_dev,
                              mean_scale=float(policy._mean_scale), init_std=float(policy._init_std),
                              actor_min_std=float(policy._min_std), gamma=gamma, lambda_=lambda_,
                              workspace=self._ws.get("imagine"), cond=_cond)
        self._ws["imagine"] = out.pop("workspace")
        if self.prefetch_noise and getattr(self, "_noise_key", None) is not None:   # next call's noise, under this call's kernel
            key, self._noise_key = self._noise_key, None
            side = getattr(self, "_noise_stream", None)
            if side is None:
                side = self._noise_stream = torch.cuda.Stream(device=key[4])
            with torch.cuda.stream(side):
                ea2 = torch.randn(key[0], key[1], key[2], device=key[4])
                ep2 = torch.randn(key[0], key[1], key[3], device=key[4])
                ev2 = torch.cuda.Event()
                ev2.record(side)
            self._noise_next = (key, ea2, ep2, ev2)
        traj = [out["beliefs"], out["prior_states"], out["prior_means"], out["prior_std_devs"]]
        if return_extras:
            return traj, out
        return traj

    # north_star vocabulary
    def img_step(self, prev_belief, state, action, *, eps=None):
        belief = self.compute_belief(prev_belief, state, action)
        return (belief,) + tuple(self.compute_prior_state(belief, eps=eps))

    def obs_step(self, prev_belief, state, action, observation, *, eps_prior=None, eps_post=None):
        belief = self.compute_belief(prev_belief, state, action)
        prior = self.compute_prior_state(belief, eps=eps_prior)
        post = self.compute_posterior_state(belief, observation, eps=eps_post)
        return (belief,) + tuple(prior) + tuple(post)

    # cell-level methods (rssm.py:34-64).  Without gradients they run as one-step programs of the fused machine; when a
    # gradient is required they are composed from the differentiable GEMM op (autograd.LinearFn: tcgen05 forward, data and
    # weight gradients) plus elementwise torch ops, so standalone calls stay autograd-connected like the reference's.
    def _act(self, x):
        return getattr(torch.nn.functional, self.activation_function)(x)

    def compute_belief(self, prev_belief, state, action):
        if self._wants_grad([prev_belief, state, action]):
            from .autograd import LinearFn
            fc, rnn = self.fc_embed_state_action, self.rnn
            h = self._act(LinearFn.apply(torch.cat([state, action], dim=1), fc.weight, fc.bias))
            gi = LinearFn.apply(h, rnn.weight_ih, rnn.bias_ih)          # torch GRUCell: gates ordered r, z, n
            gh = LinearFn.apply(prev_belief, rnn.weight_hh, rnn.bias_hh)
            i_r, i_z, i_n = gi.chunk(3, dim=1)
            h_r, h_z, h_n = gh.chunk(3, dim=1)
            r, z = torch.sigmoid(i_r + h_r), torch.sigmoid(i_z + h_z)
            n = torch.tanh(i_n + r * h_n)
            return n + z * (prev_belief - n)
        with torch.no_grad():
            outs = self.observe(prev_belief, state, action.unsqueeze(0),
                                eps_prior=torch.zeros(1, state.shape[0], self.state_size, device=state.device))
        return outs[0][0]

    def _gaussian_head(self, x, fc1, fc2, eps):
        from .autograd import LinearFn
        hidden = self._act(LinearFn.apply(x, fc1.weight, fc1.bias))
        mean, raw = LinearFn.apply(hidden, fc2.weight, fc2.bias).chunk(2, dim=1)      # mean first (rssm.py:45-48)
        std = torch.nn.functional.softplus(raw) + self.min_std_dev
        return mean + std * eps, mean, std

    def compute_prior_state(self, belief, *, eps=None):
        """rssm.py:42-50 -> (prior_state, prior_mean, prior_std_dev)."""
        if eps is None:
            eps = torch.randn(belief.shape[0], self.state_size, device=belief.device)
        if self._wants_grad([belief]):
            return self._gaussian_head(belief, self.fc_embed_belief_prior, self.fc_state_prior, eps)
        return ops.cell_fwd(_named(self), belief, None, eps, act=self.activation_function, min_std=self.min_std_dev)

    def compute_posterior_state(self, belief, observation, *, eps=None):
        """rssm.py:52-64 -> (posterior_state, posterior_mean, posterior_std_dev)."""
        if eps is None:
            eps = torch.randn(belief.shape[0], self.state_size, device=belief.device)
        if self._wants_grad([belief, observation]):
            return self._gaussian_head(torch.cat([belief, observation], dim=1), self.fc_embed_belief_posterior,
                                       self.fc_state_posterior, eps)
        return ops.cell_fwd(_named(self), belief, observation, eps, act=self.activation_function, min_std=self.min_std_dev)

    # ------------------------------------------------------------------ helpers
    def _wants_grad(self, tensors) -> bool:
        if not torch.is_grad_enabled():
            return False
        return any(p.requires_grad for p in self.parameters()) or any(t is not None and t.requires_grad for t in tensors)


class ConditionalTransitionModel(TransitionModel):
    """rssm.py:187-248 (multitask variants): the task condition rides next to the action.  `observe` concatenates
    (actions, conditions) into pseudo-actions; `imagine` keeps the (constant) condition in the tail of the kernel's action
    slot, where the embedding layer reads [state | action | condition] and the ConditionalActorModel's first layer reads
    [belief | state | condition].  Same fused launches, same hand-written backward."""

    def __init__(self, belief_size, state_size, action_size, hidden_size, embedding_size, condition_size,
                 activation_function="relu", min_std_dev=0.1):
        super().__init__(belief_size, state_size, action_size + condition_size, hidden_size, embedding_size,
                         activation_function, min_std_dev)
        self.condition_size = condition_size

    def observe(self, prev_belief, prev_state, actions, conditions, observations=None, nonterminals=None, **eps):
        return super().observe(prev_belief, prev_state, torch.cat((actions, conditions), dim=2), observations, nonterminals, **eps)

    def imagine(self, prev_belief, prev_state, condition, policy, horizon, **kw):
        if condition.shape != (prev_belief.shape[0], self.condition_size):
            raise ValueError(f"condition must be ({prev_belief.shape[0]}, {self.condition_size}), got {tuple(condition.shape)}")
        return super().imagine(prev_belief, prev_state, policy, horizon, _cond=condition.float().contiguous(), **kw)
