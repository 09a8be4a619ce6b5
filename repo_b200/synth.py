"""Seeded synthetic weights and inputs of the DMC shape: numpy RandomState, so the
same bytes come out in the build container and on the GPU box.  Used by bench.py, smoke(), the tests
and (re-exported) the oracle; contains no model arithmetic."""
from __future__ import annotations

import math
from typing import Dict

import numpy as np
import torch

Params = Dict[str, torch.Tensor]

DEFAULT_DIMS = dict(belief=200, state=30, action=6, hidden=200, embed=1024)


def _uniform(rs: np.random.RandomState, shape, bound) -> torch.Tensor:
    return torch.from_numpy(rs.uniform(-bound, bound, size=shape).astype(np.float32))


def make_transition_params(seed: int, dims=DEFAULT_DIMS, scale: float = 1.0) -> Params:
    """Same key names/shapes as TransitionModel.state_dict() (rssm.py:9-32).

    Values follow torch's default init *distribution* (U(+-1/sqrt(fan_in))) but are
    drawn from numpy so fixtures are reproducible anywhere. `scale` > 1 makes the
    recurrence livelier (stresses numerics more than a fresh init does)."""
    rs = np.random.RandomState(seed)
    D, S, A, H, E = dims["belief"], dims["state"], dims["action"], dims["hidden"], dims["embed"]
    p: Params = {}

    def lin(name, out_f, in_f):
        b = scale / math.sqrt(in_f)
        p[name + ".weight"] = _uniform(rs, (out_f, in_f), b)
        p[name + ".bias"] = _uniform(rs, (out_f,), b)

    lin("fc_embed_state_action", D, S + A)
    b = scale / math.sqrt(D)
    p["rnn.weight_ih"] = _uniform(rs, (3 * D, D), b)
    p["rnn.weight_hh"] = _uniform(rs, (3 * D, D), b)
    p["rnn.bias_ih"] = _uniform(rs, (3 * D,), b)
    p["rnn.bias_hh"] = _uniform(rs, (3 * D,), b)
    lin("fc_embed_belief_prior", H, D)
    lin("fc_state_prior", 2 * S, H)
    lin("fc_embed_belief_posterior", H, D + E)
    lin("fc_state_posterior", 2 * S, H)
    return p


def make_mlp_params(seed: int, in_f: int, hidden: int, out_f: int, n_hidden: int, scale: float = 1.0) -> Params:
    """fc1..fc{n_hidden+1}: in_f -> hidden x n_hidden -> out_f  (ActorModel n_hidden=4,
    out=2A; RewardModel/ValueModel n_hidden=3, out=1)."""
    rs = np.random.RandomState(seed)
    p: Params = {}
    sizes = [in_f] + [hidden] * n_hidden + [out_f]
    for i in range(len(sizes) - 1):
        b = scale / math.sqrt(sizes[i])
        p[f"fc{i + 1}.weight"] = _uniform(rs, (sizes[i + 1], sizes[i]), b)
        p[f"fc{i + 1}.bias"] = _uniform(rs, (sizes[i + 1],), b)
    return p

def make_ensemble_params(seed: int, in_f: int, hidden: int, out_f: int, ensemble: int) -> Params:
    """EnsembleDynamicsModel fc1..fc4 (models/utils.py:52-80): weight (E, in, out), bias (E, 1, out).  Drawn in
    +-1/sqrt(fan_in) instead of the reference's U(0,1) default so that four layers keep O(1) activations."""
    rs = np.random.RandomState(seed)
    p: Params = {}
    sizes = [in_f, hidden, hidden, hidden, out_f]
    for i in range(4):
        b = 1.0 / math.sqrt(sizes[i])
        p[f"fc{i + 1}.weight"] = _uniform(rs, (ensemble, sizes[i], sizes[i + 1]), b)
        p[f"fc{i + 1}.bias"] = _uniform(rs, (ensemble, 1, sizes[i + 1]), b)
    return p


def make_head_rollout(seed: int, T: int, B: int, dims=DEFAULT_DIMS, p_done=0.1):
    """Inputs of Dreamer.train_disag / train_inv_dynamics (dreamer.py:198-239): beliefs/states (T-1,B,.) as observe
    returns them, actions / nonterms (T,B,.) as the replay buffer does."""
    rs = np.random.RandomState(seed)
    f = lambda a: torch.from_numpy(a.astype(np.float32))
    D, S, A = dims["belief"], dims["state"], dims["action"]
    return dict(beliefs=f(np.clip(rs.standard_normal((T - 1, B, D)) * 0.3, -1, 1)), states=f(rs.standard_normal((T - 1, B, S))),
                actions=f(rs.uniform(-1, 1, (T, B, A))), nonterms=f(rs.uniform(0, 1, (T, B, 1)) >= p_done))


def make_observe_inputs(seed: int, T: int, B: int, dims=DEFAULT_DIMS, p_done=1 / 500.0, embed_scale=1.0):
    rs = np.random.RandomState(seed)
    D, S, A, E = dims["belief"], dims["state"], dims["action"], dims["embed"]
    T1 = T - 1
    f = lambda a: torch.from_numpy(a.astype(np.float32))
    return dict(
        prev_belief=torch.zeros(B, D), prev_state=torch.zeros(B, S),
        actions=f(rs.uniform(-1, 1, (T1, B, A))),
        embeds=f(rs.standard_normal((T1, B, E)) * embed_scale),
        nonterms=f((rs.uniform(0, 1, (T1, B, 1)) >= p_done)),
        eps_prior=f(rs.standard_normal((T1, B, S))),
        eps_post=f(rs.standard_normal((T1, B, S))),
    )


def make_imagine_inputs(seed: int, N: int, horizon: int, dims=DEFAULT_DIMS):
    rs = np.random.RandomState(seed)
    D, S, A = dims["belief"], dims["state"], dims["action"]
    f = lambda a: torch.from_numpy(a.astype(np.float32))
    return dict(
        belief=f(np.clip(rs.standard_normal((N, D)) * 0.3, -1, 1)),
        state=f(rs.standard_normal((N, S))),
        eps_action=f(rs.standard_normal((horizon - 1, N, A))),
        eps_prior=f(rs.standard_normal((horizon - 1, N, S))),
    )


def make_mask_head_params(seed: int) -> Params:
    """nn.Sequential(nn.Conv2d(6, 1, 1), nn.Sigmoid()) (tia.py:72)."""
    rs = np.random.RandomState(seed)
    b = 1.0 / math.sqrt(6)
    return {"0.weight": _uniform(rs, (1, 6, 1, 1), b), "0.bias": _uniform(rs, (1,), b)}


def make_conv_params(which: str, seed: int, out_channels: int = 3, embedding_size: int = 1024) -> Params:
    """VisualEncoder (encoder.py:26-29) / VisualObservationModel (decoder.py:35-39) parameters with torch's default
    init distribution (U(+-1/sqrt(fan_in))), from numpy."""
    rs = np.random.RandomState(seed)
    p: Params = {}
    if which == "encoder":
        for i, (ci, co) in enumerate([(3, 32), (32, 64), (64, 128), (128, 256)], 1):
            b = 1.0 / math.sqrt(ci * 16)
            p[f"conv{i}.weight"] = _uniform(rs, (co, ci, 4, 4), b)
            p[f"conv{i}.bias"] = _uniform(rs, (co,), b)
        if embedding_size != 1024:   # encoder.py:30
            b = 1.0 / math.sqrt(1024)
            p["fc.weight"], p["fc.bias"] = _uniform(rs, (embedding_size, 1024), b), _uniform(rs, (embedding_size,), b)
    else:
        b = 1.0 / math.sqrt(230)
        p["fc1.weight"], p["fc1.bias"] = _uniform(rs, (embedding_size, 230), b), _uniform(rs, (embedding_size,), b)
        for i, (ci, co, k) in enumerate([(embedding_size, 128, 5), (128, 64, 5), (64, 32, 6), (32, out_channels, 6)], 1):
            b = 1.0 / math.sqrt(co * k * k)  # ConvTranspose2d fan_in is computed on dim 1 of its (cin, cout, k, k) weight
            p[f"conv{i}.weight"] = _uniform(rs, (ci, co, k, k), b)
            p[f"conv{i}.bias"] = _uniform(rs, (co,), b)
    return p


def make_frames(seed: int, n: int, hw=(64, 64)) -> torch.Tensor:
    """n preprocessed frames: uint8 U{0..255} -> x/255*2-1 (common/utils.py:74-80)."""
    rs = np.random.RandomState(seed)
    u8 = rs.randint(0, 256, (n, 3) + tuple(hw)).astype(np.uint8)
    return torch.from_numpy(((u8.astype(np.float32) / 255) * 2) - 1.0)


def make_train_batch(seed: int, T: int, B: int, action: int = 6, p_done=1 / 500.0):
    """One replay batch of SURVEY §8(d) Config 1: preprocessed frames (T,B,3,64,64), actions U(-1,1), rewards U(0,2),
    nonterms = 1 - Bernoulli(1/500)."""
    rs = np.random.RandomState(seed)
    f = lambda a: torch.from_numpy(a.astype(np.float32))
    u8 = rs.randint(0, 256, (T, B, 3, 64, 64)).astype(np.uint8)
    return dict(obs=f((u8.astype(np.float32) / 255) * 2 - 1.0), actions=f(rs.uniform(-1, 1, (T, B, action))),
                rewards=f(rs.uniform(0, 2, (T, B, 1))), nonterms=f(rs.uniform(0, 1, (T, B, 1)) >= p_done))
