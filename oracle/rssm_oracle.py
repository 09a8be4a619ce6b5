"""CPU oracle for the RSSM hot path (TEST INFRASTRUCTURE — never the product path).

This file restates, in plain functional torch-on-CPU code with *explicit* noise
arguments, the algorithm of the reference's RSSM recurrence and the reductions
hanging off it.  Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s
`cpu_baseline` / `--impl reference` legs may import it.  `repo_b200/` must not.

Parity status: PINNED against the live reference.  The reference ships no tests
or golden vectors for this path, so `oracle/make_golden.py`
imports the unmodified reference modules from /root/reference in the build
container, injects fixed noise (patching `torch.randn_like` and
`torch.distributions.normal._standard_normal`), and commits the input seeds +
outputs under `tests/golden/`.  `tests/test_oracle_golden.py` checks every
function here against those fixtures (bit-exact or <=1e-6 in fp32).

The arithmetic itself lives in a third-party dependency of the reference:
torch==1.12.1 (requirements.txt:18; installed here: torch 2.11).  Formulas
restated from its published semantics: `nn.Linear` (y = x W^T + b),
`nn.GRUCell` (gate order r,z,n; n = tanh(W_in x + b_in + r*(W_hn h + b_hn));
h' = (1-z)*n + z*h), `F.elu`, `F.softplus` (threshold 20),
`kl_divergence(Normal, Normal)`, `Normal.log_prob`.

Reference citations (relative to /root/reference):
  compute_belief            algorithms/repo/models/rssm.py:34-40
  compute_prior_state       algorithms/repo/models/rssm.py:42-50
  compute_posterior_state   algorithms/repo/models/rssm.py:52-64
  observe                   algorithms/repo/models/rssm.py:76-146
  imagine                   algorithms/repo/models/rssm.py:148-184
  ActorModel                algorithms/repo/models/actor_critic.py:50-102
  ValueModel / RewardModel  actor_critic.py:9-26 / decoder.py:178-195
  TanhBijector / SampleDist algorithms/repo/models/utils.py:112-163
  lambda_return             common/utils.py:61-71
  KL terms                  algorithms/repo/repo.py:63-83, dreamer.py:278-282
  replay index math         common/buffers.py:156-166
  ensemble / inverse dyn.   algorithms/repo/models/utils.py:19-109, dreamer.py:198-239, 330-339
  conv encoder / decoder    algorithms/repo/models/encoder.py:21-41, decoder.py:28-48
"""
from __future__ import annotations

import math
from typing import Dict, Optional, Tuple

import numpy as np
import torch
import torch.nn.functional as F

Params = Dict[str, torch.Tensor]

# seeded synthetic weights / inputs live in repo_b200/synth.py (pure numpy generators, no arithmetic)
from repo_b200.synth import (DEFAULT_DIMS, make_transition_params, make_mlp_params, make_observe_inputs,  # noqa: E402,F401
                             make_imagine_inputs, make_conv_params, make_frames, make_train_batch, make_mask_head_params)


def cast_params(p: Params, dtype) -> Params:
    return {k: v.to(dtype) for k, v in p.items()}


# ----------------------------------------------------------------------------
# optional emulation of the device arithmetic (fp16 hi/lo split, 3 products, fp32
# accumulate) — used by tests to predict how far the CUDA path may sit from fp32
# ----------------------------------------------------------------------------

class Precision:
    """`exact`: plain matmul in the tensor dtype.  `f16x3`: each operand is split
    x = hi + lo (both fp16, lo subnormal-safe); the product keeps hi*hi + lo*hi +
    hi*lo, accumulated in fp32 — what the tcgen05 kernels do."""

    def __init__(self, mode: str = "exact"):
        assert mode in ("exact", "f16x3", "bf16x3", "f16", "bf16")
        self.mode = mode

    def _split(self, x, dt):
        hi = x.to(dt).to(torch.float32)
        lo = (x - hi).to(dt).to(torch.float32)
        return hi, lo

    def linear(self, x, w, b):
        if self.mode == "exact":
            return F.linear(x, w, b)
        dt = torch.float16 if self.mode.startswith("f16") else torch.bfloat16
        x32, w32 = x.float(), w.float()
        xh, xl = self._split(x32, dt)
        wh, wl = self._split(w32, dt)
        if self.mode in ("f16", "bf16"):
            y = xh.double() @ wh.double().t()
        else:
            y = xh.double() @ wh.double().t() + xl.double() @ wh.double().t() + xh.double() @ wl.double().t()
        y = y.float()
        return (y + b.float()).to(x.dtype) if b is not None else y.to(x.dtype)


EXACT = Precision("exact")

# ----------------------------------------------------------------------------
# cell-level functions
# ----------------------------------------------------------------------------

def _act(name: str):
    return getattr(F, name)


def compute_belief(p: Params, prev_belief, state, action, act="elu", prec: Precision = EXACT):
    """rssm.py:34-40 — h = act(W_sa [s|a] + b); belief' = GRUCell(h, belief)."""
    x = torch.cat([state, action], dim=1)
    hid = _act(act)(prec.linear(x, p["fc_embed_state_action.weight"], p["fc_embed_state_action.bias"]))
    gi = prec.linear(hid, p["rnn.weight_ih"], p["rnn.bias_ih"])
    gh = prec.linear(prev_belief, p["rnn.weight_hh"], p["rnn.bias_hh"])
    i_r, i_z, i_n = gi.chunk(3, dim=1)
    h_r, h_z, h_n = gh.chunk(3, dim=1)
    r = torch.sigmoid(i_r + h_r)
    z = torch.sigmoid(i_z + h_z)
    n = torch.tanh(i_n + r * h_n)
    return (1.0 - z) * n + z * prev_belief


def _gaussian_head(p: Params, pre: str, out: str, x, eps, act, min_std, prec):
    hid = _act(act)(prec.linear(x, p[pre + ".weight"], p[pre + ".bias"]))
    o = prec.linear(hid, p[out + ".weight"], p[out + ".bias"])
    mean, raw = o.chunk(2, dim=1)  # mean first, raw std second (rssm.py:45-47)
    std = F.softplus(raw) + min_std
    return mean + std * eps, mean, std


def compute_prior_state(p: Params, belief, eps, act="elu", min_std=0.1, prec: Precision = EXACT):
    """rssm.py:42-50."""
    return _gaussian_head(p, "fc_embed_belief_prior", "fc_state_prior", belief, eps, act, min_std, prec)


def compute_posterior_state(p: Params, belief, embed, eps, act="elu", min_std=0.1, prec: Precision = EXACT):
    """rssm.py:52-64."""
    x = torch.cat([belief, embed], dim=1)
    return _gaussian_head(p, "fc_embed_belief_posterior", "fc_state_posterior", x, eps, act, min_std, prec)


# ----------------------------------------------------------------------------
# observe / imagine
# ----------------------------------------------------------------------------

def observe(p: Params, prev_belief, prev_state, actions, embeds: Optional[torch.Tensor],
            nonterms: Optional[torch.Tensor], eps_prior, eps_post: Optional[torch.Tensor],
            act="elu", min_std=0.1, prec: Precision = EXACT):
    """rssm.py:76-146.  actions (T1,B,A); embeds (T1,B,E) already shifted by the caller;
    nonterms (T1,B,1); eps_* (T1,B,S) in the reference's RNG draw order (prior first,
    posterior second, every step).  Returns the reference's list of 7 (or 4) tensors."""
    T1 = actions.shape[0]
    belief, state = prev_belief, prev_state
    outs = [[] for _ in range(7)]
    for t in range(T1):
        s_in = state if nonterms is None else state * nonterms[t]  # only the state is masked (rssm.py:118-119)
        belief = compute_belief(p, belief, s_in, actions[t], act, prec)
        ps, pm, pd = compute_prior_state(p, belief, eps_prior[t], act, min_std, prec)
        outs[0].append(belief); outs[1].append(ps); outs[2].append(pm); outs[3].append(pd)
        if embeds is not None:
            qs, qm, qd = compute_posterior_state(p, belief, embeds[t], eps_post[t], act, min_std, prec)
            outs[4].append(qs); outs[5].append(qm); outs[6].append(qd)
            state = qs
        else:
            state = ps
    n = 7 if embeds is not None else 4
    return [torch.stack(o, 0) for o in outs[:n]]


def mlp(p: Params, x, n_layers: int, act="elu", prec: Precision = EXACT):
    for i in range(1, n_layers):
        x = _act(act)(prec.linear(x, p[f"fc{i}.weight"], p[f"fc{i}.bias"]))
    return prec.linear(x, p[f"fc{n_layers}.weight"], p[f"fc{n_layers}.bias"])


def actor_forward(ap: Params, belief, state, mean_scale=5.0, init_std=0.0, min_std=0.1, prec: Precision = EXACT):
    """actor_critic.py:76-87 — hidden activation is always ELU."""
    o = mlp(ap, torch.cat([belief, state], 1), 5, "elu", prec)
    m, s = o.chunk(2, dim=1)
    mean = mean_scale * torch.tanh(m / mean_scale)
    std = F.softplus(s + init_std) + min_std
    return mean, std


def head_forward(hp: Params, belief, state, act="elu", prec: Precision = EXACT):
    """RewardModel/ValueModel forward (decoder.py:189-195, actor_critic.py:20-26)."""
    return mlp(hp, torch.cat([belief, state], 1), 4, act, prec).squeeze(1)


def imagine(p: Params, ap: Params, prev_belief, prev_state, eps_action, eps_prior, horizon: int,
            act="elu", min_std=0.1, prec: Precision = EXACT):
    """rssm.py:148-184 with ActorModel.get_action = tanh(mean + std*eps) (actor_critic.py:97-102).
    eps_action (H-1,N,A), eps_prior (H-1,N,S).  Returns [beliefs, prior_states, prior_means,
    prior_std_devs] each with H-1 leading entries (start row excluded) plus the actions taken."""
    belief, state = prev_belief, prev_state
    outs = [[] for _ in range(5)]
    for t in range(horizon - 1):
        mean, std = actor_forward(ap, belief, state, prec=prec)
        action = torch.tanh(mean + std * eps_action[t])
        belief = compute_belief(p, belief, state, action, act, prec)
        state, pm, pd = compute_prior_state(p, belief, eps_prior[t], act, min_std, prec)
        for o, v in zip(outs, (belief, state, pm, pd, action)):
            o.append(v)
    return [torch.stack(o, 0) for o in outs]


# ----------------------------------------------------------------------------
# optional Dreamer heads
# ----------------------------------------------------------------------------
def ensemble_forward(ep: Params, belief, state, action, act="elu"):
    """EnsembleDynamicsModel.forward (models/utils.py:52-80) on EnsembleLinearLayer weights (E, in, out), biases
    (E, 1, out) (models/utils.py:19-49): (E, rows, belief) next-belief predictions."""
    h = torch.cat([belief, state, action], 1)
    for i in (1, 2, 3):
        h = _act(act)(torch.matmul(h, ep[f"fc{i}.weight"]) + ep[f"fc{i}.bias"])
    return torch.matmul(h, ep["fc4.weight"]) + ep["fc4.bias"]


def _transition_rows(beliefs, states, actions, nonterms):
    """dreamer.py:199-209 / :221-231: rows with nonterms[1:-1] == 1 of (actions[1:-1], beliefs[:-1], states[:-1],
    beliefs[1:]), flattened time-major."""
    keep = nonterms[1:-1].flatten() == 1
    return [x.flatten(0, 1)[keep] for x in (actions[1:-1], beliefs[:-1], states[:-1], beliefs[1:])]


def disag_loss(ep: Params, beliefs, states, actions, nonterms, act="elu"):
    """Dreamer.train_disag (dreamer.py:198-217): -Independent(Normal(pred, 1), 1).log_prob(target).sum(0).mean()."""
    a, b, s, b_next = _transition_rows(beliefs, states, actions, nonterms)
    pred = ensemble_forward(ep, b, s, a, act)
    nll = 0.5 * (pred - b_next.unsqueeze(0)) ** 2 + 0.5 * math.log(2 * math.pi)
    return nll.sum(2).sum(0).mean()


def inverse_dynamics_forward(ip: Params, belief, state, next_belief, act="elu", min_std=0.1):
    """InverseDynamicsModel.forward (models/utils.py:83-109): chunk -> mean first, std = softplus(raw) + min_std."""
    o = mlp(ip, torch.cat([belief, state, next_belief], 1), 4, act)
    mean, raw = o.chunk(2, dim=1)
    return mean, F.softplus(raw) + min_std


def inv_dyn_loss(ip: Params, beliefs, states, actions, nonterms, act="elu"):
    """Dreamer.train_inv_dynamics (dreamer.py:219-239): -Independent(Normal(mean, std), 1).log_prob(action).mean()."""
    a, b, s, b_next = _transition_rows(beliefs, states, actions, nonterms)
    mean, std = inverse_dynamics_forward(ip, b, s, b_next, act)
    nll = 0.5 * ((a - mean) / std) ** 2 + std.log() + 0.5 * math.log(2 * math.pi)
    return nll.sum(1).mean()


def disagreement(ep: Params, beliefs, states, actions, act="elu"):
    """dreamer.py:333-338: ensemble spread ens_preds.std(0).mean(-1) (torch's unbiased std over the members)."""
    return ensemble_forward(ep, beliefs, states, actions, act).std(0).mean(-1)


# ----------------------------------------------------------------------------
# reductions hanging off the recurrence
# ----------------------------------------------------------------------------

def lambda_return(rewards, values, discounts, bootstrap, lambda_=0.95):
    """common/utils.py:61-71."""
    next_values = torch.cat([values[1:], bootstrap[None]], 0)
    inputs = rewards + discounts * next_values * (1 - lambda_)
    last = bootstrap
    out = [None] * inputs.shape[0]
    for t in range(inputs.shape[0] - 1, -1, -1):
        last = inputs[t] + discounts[t] * lambda_ * last
        out[t] = last
    return torch.stack(out, 0)


def imagine_conditional(p: Params, ap: Params, prev_belief, prev_state, condition, eps_action, eps_prior, horizon: int,
                        act="elu", min_std=0.1, prec: Precision = EXACT):
    """ConditionalTransitionModel.imagine (rssm.py:225-248) with a ConditionalActorModel (actor_critic.py:131-148): the
    actor sees [belief | state | condition] (detached), the dynamics see the pseudo-action [action | condition]."""
    belief, state = prev_belief, prev_state
    outs = [[] for _ in range(5)]
    for t in range(horizon - 1):
        mean, std = actor_forward(ap, belief.detach(), torch.cat([state.detach(), condition], 1), prec=prec)
        action = torch.tanh(mean + std * eps_action[t])
        belief = compute_belief(p, belief, state, torch.cat([action, condition], 1), act, prec)
        state, pm, pd = compute_prior_state(p, belief, eps_prior[t], act, min_std, prec)
        for lst, v in zip(outs, (belief, state, pm, pd, action)):
            lst.append(v)
    return [torch.stack(o, 0) for o in outs]


def imagine_returns(reward_preds, value_preds, gamma=0.99, lambda_=0.95):
    """dreamer.py:341-349 — H-2 rows, bootstrap = value_preds[-1]."""
    disc = gamma * torch.ones_like(reward_preds)
    return lambda_return(reward_preds[:-1], value_preds[:-1], disc[:-1], value_preds[-1], lambda_)


def kl_normal(mp, sp, mq, sq):
    """KL(N(mp,sp) || N(mq,sq)) elementwise (torch/distributions/kl.py _kl_normal_normal)."""
    var_ratio = (sp / sq) ** 2
    t1 = ((mp - mq) / sq) ** 2
    return 0.5 * (var_ratio + t1 - 1 - var_ratio.log())


def kl_sum(post_mean, post_std, prior_mean, prior_std):
    """KL(post||prior).sum over the state dim -> (T1,B)."""
    return kl_normal(post_mean, post_std, prior_mean, prior_std).sum(2)


def dreamer_kl_loss(kl_tb, free_nats=3.0):
    """dreamer.py:278-282 — max(kl, free_nats) per (t,b), then mean."""
    return torch.clamp(kl_tb, min=free_nats).mean()


def repo_kl_terms(kl_tb, log_beta, prior_train_steps=5, target_kl=3.0):
    """repo.py:63-96 forward values: kl_div, kl_viol, kl_loss, beta_loss."""
    kl_mean = kl_tb.mean()
    alpha = prior_train_steps / (1 + prior_train_steps)
    kl_div = alpha * kl_mean + (1 - alpha) * kl_mean
    kl_viol = kl_div - target_kl
    kl_loss = math.exp(log_beta) * kl_viol
    beta_loss = -log_beta * kl_viol
    return kl_div, kl_viol, kl_loss, beta_loss


def tanh_normal_entropy(mean, std, eps):
    """SampleDist.entropy over Independent(Transformed(Normal, TanhBijector),1)
    (models/utils.py:112-163). eps (K,M,A) standard-normal draws. Returns (M,)."""
    x = mean[None] + std[None] * eps
    y = torch.tanh(x)
    yc = torch.where(y.abs() <= 1.0, torch.clamp(y, -0.99999997, 0.99999997), y)
    xi = torch.atanh(yc)
    var = std[None] ** 2
    base_lp = -((xi - mean[None]) ** 2) / (2 * var) - std[None].log() - math.log(math.sqrt(2 * math.pi))
    ladj = 2.0 * (math.log(2.0) - xi - F.softplus(-2.0 * xi))
    lp = (base_lp - ladj).sum(-1)
    return -lp.mean(0)


# ----------------------------------------------------------------------------
# sequence bookkeeping (bit-exact integer work)
# ----------------------------------------------------------------------------

def replay_indices(start_inds: np.ndarray, seq_len: int, pos: int, full: bool, length: int) -> np.ndarray:
    """common/buffers.py:156-166 after `np.random.choice` — time-major flat indices."""
    inds = np.stack([np.arange(s, s + seq_len) for s in start_inds], 0)
    inds = inds.transpose().reshape(-1)
    if full:
        inds = (inds + pos) % length
    return inds


def shift_for_observe(actions, embeds, nonterms):
    """dreamer.py:253-258 — the caller-side time shift paired with observe()."""
    return actions[:-1], embeds[1:], nonterms[:-1]




# ----------------------------------------------------------------------------
# conv stacks (second tier): restated with torch's conv ops, which is where the
# reference's arithmetic lives (encoder.py:26-41, decoder.py:35-48)
# ----------------------------------------------------------------------------

def visual_encoder(p: Params, obs):
    """encoder.py:34-41; fc = Identity at the default embedding_size == 1024, else Linear(1024, embedding_size)."""
    h = obs
    for i in range(1, 5):
        h = F.relu(F.conv2d(h, p[f"conv{i}.weight"], p[f"conv{i}.bias"], stride=2))
    h = h.reshape(-1, 1024)
    if "fc.weight" in p:
        h = F.linear(h, p["fc.weight"], p["fc.bias"])
    return h


def visual_decoder(p: Params, belief, state):
    """decoder.py:41-48: fc1 without activation, then 4 transposed convolutions (ReLU on the first three)."""
    h = F.linear(torch.cat([belief, state], 1), p["fc1.weight"], p["fc1.bias"]).reshape(-1, p["fc1.weight"].shape[0], 1, 1)
    for i in range(1, 4):
        h = F.relu(F.conv_transpose2d(h, p[f"conv{i}.weight"], p[f"conv{i}.bias"], stride=2))
    return F.conv_transpose2d(h, p["conv4.weight"], p["conv4.bias"], stride=2)
