"""Generate tests/golden/*.npz by running the UNMODIFIED reference (/root/reference).

Run in the build container only (the GPU box has no /root/reference):

    python oracle/make_golden.py

The reference modules are loaded by file path (the package import pulls in
matplotlib, which is not installed — SURVEY.md §8c).  Noise is injected by
patching `torch.randn_like` (rssm.py:49,62) and
`torch.distributions.normal._standard_normal` (Normal.rsample) so that reference,
oracle and CUDA path all consume the same epsilon tensors.  Weights and inputs come
from numpy RandomState seeds (oracle/rssm_oracle.py make_*), so the fixtures only
need to store seeds + outputs.
"""
from __future__ import annotations

import importlib.util
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import rssm_oracle as O  # noqa: E402

REF = "/root/reference/"
OUT = os.path.join(ROOT, "tests", "golden")


def _load(name, path):
    spec = importlib.util.spec_from_file_location(name, path)
    m = importlib.util.module_from_spec(spec)
    sys.modules[name] = m
    spec.loader.exec_module(m)
    return m


def load_reference():
    pkg = types.ModuleType("refmodels")
    pkg.__path__ = [REF + "algorithms/repo/models"]
    sys.modules["refmodels"] = pkg
    mods = types.SimpleNamespace()
    mods.utils = _load("refmodels.utils", REF + "algorithms/repo/models/utils.py")
    mods.actor_critic = _load("refmodels.actor_critic", REF + "algorithms/repo/models/actor_critic.py")
    mods.rssm = _load("refmodels.rssm", REF + "algorithms/repo/models/rssm.py")
    mods.decoder = _load("refmodels.decoder", REF + "algorithms/repo/models/decoder.py")
    mods.common_utils = _load("ref_common_utils", REF + "common/utils.py")
    mods.buffers = _load("ref_common_buffers", REF + "common/buffers.py")
    return mods


class NoiseInjector:
    """Pops pre-drawn tensors in the order the reference asks for them."""

    def __init__(self, queue):
        self.queue = list(queue)
        self._orig_randn_like = torch.randn_like
        self._orig_std_normal = torch.distributions.normal._standard_normal

    def _pop(self, shape):
        t = self.queue.pop(0)
        assert tuple(t.shape) == tuple(shape), (t.shape, shape)
        return t

    def __enter__(self):
        torch.randn_like = lambda x, **kw: self._pop(x.shape)
        torch.distributions.normal._standard_normal = lambda shape, dtype, device: self._pop(shape)
        return self

    def __exit__(self, *a):
        torch.randn_like = self._orig_randn_like
        torch.distributions.normal._standard_normal = self._orig_std_normal
        assert not self.queue, f"{len(self.queue)} noise tensors unused"


def np_(d):
    return {k: (v.detach().numpy() if isinstance(v, torch.Tensor) else np.asarray(v)) for k, v in d.items()}


def build_models(R, dims, seeds, scale):
    D, S, A, H, E = dims["belief"], dims["state"], dims["action"], dims["hidden"], dims["embed"]
    tm = R.rssm.TransitionModel(D, S, A, H, E, "elu")
    tm.load_state_dict(O.make_transition_params(seeds[0], dims, scale))
    # positional-arg quirk reproduced on purpose: 5th positional lands in `dist` (dreamer.py:99-105)
    actor = R.actor_critic.ActorModel(D, S, H, A, "elu")
    actor.load_state_dict(O.make_mlp_params(seeds[1], D + S, H, 2 * A, 4, scale))
    reward = R.decoder.RewardModel(D, S, H, "elu")
    reward.load_state_dict(O.make_mlp_params(seeds[2], D + S, H, 1, 3, scale))
    value = R.actor_critic.ValueModel(D, S, H, "elu")
    value.load_state_dict(O.make_mlp_params(seeds[3], D + S, H, 1, 3, scale))
    return tm, actor, reward, value


def golden_observe(R, name, dims, T, B, seed, scale, embed_scale, p_done, keep="all", use_obs=True, use_nt=True):
    tm, _, _, _ = build_models(R, dims, (seed, seed + 1, seed + 2, seed + 3), scale)
    x = O.make_observe_inputs(seed + 10, T, B, dims, p_done=p_done, embed_scale=embed_scale)
    T1 = T - 1
    queue = []
    for t in range(T1):
        queue.append(x["eps_prior"][t])
        if use_obs:
            queue.append(x["eps_post"][t])
    with torch.no_grad(), NoiseInjector(queue):
        outs = tm.observe(x["prev_belief"], x["prev_state"], x["actions"],
                          x["embeds"] if use_obs else None, x["nonterms"] if use_nt else None)
    names = ["beliefs", "prior_states", "prior_means", "prior_std_devs",
             "posterior_states", "posterior_means", "posterior_std_devs"][: len(outs)]
    res = dict(zip(names, outs))
    save = {}
    if use_obs:
        kl = torch.distributions.kl.kl_divergence(
            torch.distributions.Normal(res["posterior_means"], res["posterior_std_devs"]),
            torch.distributions.Normal(res["prior_means"], res["prior_std_devs"])).sum(2)
        save["kl_tb"] = kl
        save["kl_dreamer"] = torch.max(kl, torch.full((1,), 3.0)).mean((0, 1))
        save["kl_mean"] = kl.mean((0, 1))
    for k, v in res.items():
        save[k] = v if keep == "all" else v[-2:]
    meta = dict(T=T, B=B, seed=seed, scale=scale, embed_scale=embed_scale, p_done=p_done,
                keep=0 if keep == "all" else 2, use_obs=int(use_obs), use_nt=int(use_nt),
                **{"dim_" + k: v for k, v in dims.items()})
    np.savez(os.path.join(OUT, name + ".npz"), **np_(save), **{("meta_" + k): np.asarray(v) for k, v in meta.items()})
    print(name, {k: tuple(v.shape) for k, v in save.items()})


def golden_imagine(R, name, dims, N, H, seed, scale):
    tm, actor, reward, value = build_models(R, dims, (seed, seed + 1, seed + 2, seed + 3), scale)
    x = O.make_imagine_inputs(seed + 20, N, H, dims)
    queue = []
    for t in range(H - 1):
        queue.append(x["eps_action"][t])
        queue.append(x["eps_prior"][t])
    bottle = R.utils.bottle
    with torch.no_grad(), NoiseInjector(queue):
        outs = tm.imagine(x["belief"], x["state"], actor, H)
        rew = bottle(reward, (outs[0], outs[1]))
        val = bottle(value, (outs[0], outs[1]))
        disc = 0.99 * torch.ones_like(rew)
        ret = R.common_utils.lambda_return(rew[:-1], val[:-1], disc[:-1], val[-1], 0.95)
    save = dict(beliefs=outs[0], prior_states=outs[1], prior_means=outs[2], prior_std_devs=outs[3],
                rewards=rew, values=val, returns=ret)
    meta = dict(N=N, H=H, seed=seed, scale=scale, **{"dim_" + k: v for k, v in dims.items()})
    np.savez(os.path.join(OUT, name + ".npz"), **np_(save), **{("meta_" + k): np.asarray(v) for k, v in meta.items()})
    print(name, {k: tuple(v.shape) for k, v in save.items()})


def golden_conditional(R, name, N, H, T, B, C, seed):
    """ConditionalTransitionModel (rssm.py:187-248) + ConditionalActorModel (actor_critic.py:105-148): observe on
    (actions, conditions) and imagine with a per-row condition, default sizes, injected noise."""
    dims = dict(O.DEFAULT_DIMS)
    D, S, A, Hd, E = dims["belief"], dims["state"], dims["action"], dims["hidden"], dims["embed"]
    pdims = dict(dims, action=A + C)
    tm = R.rssm.ConditionalTransitionModel(D, S, A, Hd, E, C, "elu")
    tm.load_state_dict(O.make_transition_params(seed, pdims))
    actor = R.actor_critic.ConditionalActorModel(D, S, Hd, A, C, "elu")
    actor.load_state_dict(O.make_mlp_params(seed + 1, D + S + C, Hd, 2 * A, 4))
    rs = np.random.RandomState(seed + 2)
    cond = torch.from_numpy(rs.standard_normal((N, C)).astype(np.float32))
    x = O.make_imagine_inputs(seed + 20, N, H, dims)
    queue = []
    for t in range(H - 1):
        queue += [x["eps_action"][t], x["eps_prior"][t]]
    with torch.no_grad(), NoiseInjector(queue):
        im = tm.imagine(x["belief"], x["state"], cond, actor, H)
    xo = O.make_observe_inputs(seed + 10, T, B, dims)
    conds = torch.from_numpy(rs.standard_normal((T - 1, B, C)).astype(np.float32))
    queue = []
    for t in range(T - 1):
        queue += [xo["eps_prior"][t], xo["eps_post"][t]]
    with torch.no_grad(), NoiseInjector(queue):
        ob = tm.observe(xo["prev_belief"], xo["prev_state"], xo["actions"], conds, xo["embeds"], xo["nonterms"])
    save = dict(im_beliefs=im[0], im_prior_states=im[1], im_prior_means=im[2], im_prior_std_devs=im[3],
                ob_beliefs=ob[0], ob_posterior_states=ob[4], ob_posterior_means=ob[5], ob_prior_std_devs=ob[3])
    meta = dict(N=N, H=H, T=T, B=B, C=C, seed=seed)
    np.savez(os.path.join(OUT, name + ".npz"), **np_(save), **{("meta_" + k): np.asarray(v) for k, v in meta.items()})
    print(name, {k: tuple(v.shape) for k, v in save.items()})


def golden_entropy(R, name, seed, M, A, K):
    rs = np.random.RandomState(seed)
    mean = torch.from_numpy((rs.standard_normal((M, A)) * 2.0).astype(np.float32))
    std = torch.from_numpy((rs.uniform(0.1, 1.5, (M, A))).astype(np.float32))
    eps = torch.from_numpy(rs.standard_normal((K, M, A)).astype(np.float32))
    from torch.distributions import Normal, Independent, TransformedDistribution
    dist = R.utils.SampleDist(Independent(TransformedDistribution(Normal(mean, std), R.utils.TanhBijector()), 1), samples=K)
    with torch.no_grad(), NoiseInjector([eps]):
        ent = dist.entropy()
    np.savez(os.path.join(OUT, name + ".npz"), mean=mean.numpy(), std=std.numpy(), eps=eps.numpy(), entropy=ent.numpy())
    print(name, tuple(ent.shape))


def golden_lambda_return(R, name):
    # 3-step hand-checkable toy + a random case
    r = torch.tensor([[1.0], [2.0], [3.0]])
    v = torch.tensor([[0.5], [0.25], [0.125]])
    d = 0.9 * torch.ones_like(r)
    boot = torch.tensor([4.0])
    toy = R.common_utils.lambda_return(r, v, d, boot, 0.8)
    rs = np.random.RandomState(7)
    r2 = torch.from_numpy(rs.standard_normal((13, 37)).astype(np.float32))
    v2 = torch.from_numpy(rs.standard_normal((13, 37)).astype(np.float32))
    b2 = torch.from_numpy(rs.standard_normal((37,)).astype(np.float32))
    d2 = 0.99 * torch.ones_like(r2)
    out2 = R.common_utils.lambda_return(r2, v2, d2, b2, 0.95)
    np.savez(os.path.join(OUT, name + ".npz"), toy_r=r.numpy(), toy_v=v.numpy(), toy_boot=boot.numpy(), toy_out=toy.numpy(),
             r=r2.numpy(), v=v2.numpy(), boot=b2.numpy(), out=out2.numpy())
    print(name, toy.flatten().tolist())


def golden_replay(R, name):
    """SequenceReplayBuffer.sample index bookkeeping (buffers.py:156-166) incl. the wrapped/full case."""
    cases = {}
    for tag, cap, n_push, B, L, seed in [("partial", 64, 40, 5, 7, 3), ("full", 64, 64 + 23, 6, 9, 4), ("fullwrap", 50, 50 * 3 + 49, 8, 10, 5)]:
        buf = R.buffers.SequenceReplayBuffer(cap, (2,), (1,))
        for i in range(n_push):
            buf.push(np.array([i, -i], np.float32), np.array([i * 0.5], np.float32), float(i), float(i % 11 == 0))
        np.random.seed(seed)
        obs, act, rew, done = buf.sample(B, L)
        np.random.seed(seed)
        starts = np.random.choice(len(buf) - L, size=B)
        cases.update({f"{tag}_obs": obs, f"{tag}_act": act, f"{tag}_rew": rew, f"{tag}_done": done, f"{tag}_starts": starts,
                      f"{tag}_meta": np.array([cap, n_push, B, L, buf.pos, int(buf.full), len(buf)])})
    np.savez(os.path.join(OUT, name + ".npz"), **cases)
    print(name, list(cases))


def main():
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(1)  # deterministic summation order for the fixtures
    R = load_reference()
    dims = O.DEFAULT_DIMS
    small = dict(belief=32, state=8, action=3, hidden=24, embed=40)
    # observe: full tensors for a short sequence, tail only for the default shape
    golden_observe(R, "observe_T8_B10", dims, T=8, B=10, seed=100, scale=1.0, embed_scale=1.0, p_done=0.15)
    golden_observe(R, "observe_T8_B10_hot", dims, T=8, B=10, seed=110, scale=2.0, embed_scale=0.5, p_done=0.1)
    golden_observe(R, "observe_default_tail", dims, T=50, B=50, seed=120, scale=1.0, embed_scale=1.0, p_done=1 / 500.0, keep="tail")
    golden_observe(R, "observe_prior_only", dims, T=6, B=7, seed=130, scale=1.0, embed_scale=1.0, p_done=0.2, use_obs=False)
    golden_observe(R, "observe_no_nonterm", dims, T=5, B=4, seed=140, scale=1.0, embed_scale=1.0, p_done=0.0, use_nt=False)
    golden_observe(R, "observe_tiny_dims", small, T=9, B=3, seed=150, scale=1.5, embed_scale=1.0, p_done=0.2)
    golden_observe(R, "observe_T2_B1", dims, T=2, B=1, seed=160, scale=1.0, embed_scale=1.0, p_done=0.0)
    # imagine
    golden_imagine(R, "imagine_N24_H6", dims, N=24, H=6, seed=200, scale=1.0)
    golden_imagine(R, "imagine_N8_H15", dims, N=8, H=15, seed=210, scale=1.0)
    golden_imagine(R, "imagine_N16_H15_hot", dims, N=16, H=15, seed=220, scale=2.0)
    golden_imagine(R, "imagine_tiny_dims", small, N=5, H=4, seed=230, scale=1.5)
    golden_conditional(R, "conditional_N12_H7", N=12, H=7, T=5, B=6, C=4, seed=260)
    golden_entropy(R, "entropy_M50_A6_K100", seed=300, M=50, A=6, K=100)
    golden_lambda_return(R, "lambda_return")
    golden_replay(R, "replay_indices")


if __name__ == "__main__":
    main()
