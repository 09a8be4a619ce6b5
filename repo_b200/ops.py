"""Tensor-level entry points over the C-ABI (forward kernels, no autograd here).

Every function takes contiguous fp32 CUDA tensors, allocates the outputs and a workspace with
torch (PyTorch is the device-memory / stream plumbing), and launches on the current stream.
Anything else raises — there is no CPU path (north_star: "no CPU fallback")."""
from __future__ import annotations

import ctypes as C
from typing import Dict, List, Optional, Sequence

import torch

from . import _lib
from ._lib import Dims, MlpWeights, RssmWeights

_RSSM_KEYS = [
    ("fc_embed_state_action_w", "fc_embed_state_action.weight"), ("fc_embed_state_action_b", "fc_embed_state_action.bias"),
    ("rnn_w_ih", "rnn.weight_ih"), ("rnn_w_hh", "rnn.weight_hh"), ("rnn_b_ih", "rnn.bias_ih"), ("rnn_b_hh", "rnn.bias_hh"),
    ("fc_embed_belief_prior_w", "fc_embed_belief_prior.weight"), ("fc_embed_belief_prior_b", "fc_embed_belief_prior.bias"),
    ("fc_state_prior_w", "fc_state_prior.weight"), ("fc_state_prior_b", "fc_state_prior.bias"),
    ("fc_embed_belief_posterior_w", "fc_embed_belief_posterior.weight"), ("fc_embed_belief_posterior_b", "fc_embed_belief_posterior.bias"),
    ("fc_state_posterior_w", "fc_state_posterior.weight"), ("fc_state_posterior_b", "fc_state_posterior.bias"),
]


def _chk(t: torch.Tensor, name: str, shape: Optional[Sequence[int]] = None) -> torch.Tensor:
    if not isinstance(t, torch.Tensor):
        raise TypeError(f"{name}: expected a torch.Tensor, got {type(t)}")
    if not t.is_cuda:
        raise RuntimeError(f"{name}: repo_b200 runs on CUDA tensors only (got {t.device}); there is no CPU fallback")
    if t.dtype != torch.float32:
        raise RuntimeError(f"{name}: expected float32, got {t.dtype}")
    if shape is not None and tuple(t.shape) != tuple(shape):
        raise RuntimeError(f"{name}: expected shape {tuple(shape)}, got {tuple(t.shape)}")
    return t.contiguous()


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else C.c_void_p(t.data_ptr())


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def dims_of(params: Dict[str, torch.Tensor]) -> Dims:
    """Recover (belief, state, action, hidden, embed) from TransitionModel state_dict shapes."""
    D = params["rnn.weight_hh"].shape[1]
    Hd = params["fc_embed_belief_prior.weight"].shape[0]
    S = params["fc_state_prior.weight"].shape[0] // 2
    A = params["fc_embed_state_action.weight"].shape[1] - S
    E = params["fc_embed_belief_posterior.weight"].shape[1] - D
    return Dims(D, S, A, Hd, E)


class _Keep:
    """Keeps the contiguous views alive while the C structs point at them."""

    def __init__(self):
        self.t: List[torch.Tensor] = []

    def __call__(self, t: torch.Tensor, name: str, shape=None):
        t = _chk(t.detach(), name, shape)
        self.t.append(t)
        return t.data_ptr()


def rssm_struct(params: Dict[str, torch.Tensor], keep: _Keep) -> RssmWeights:
    d = dims_of(params)
    D, S, A, Hd, E = d.belief, d.state, d.action, d.hidden, d.embed
    shapes = {
        "fc_embed_state_action.weight": (D, S + A), "fc_embed_state_action.bias": (D,),
        "rnn.weight_ih": (3 * D, D), "rnn.weight_hh": (3 * D, D), "rnn.bias_ih": (3 * D,), "rnn.bias_hh": (3 * D,),
        "fc_embed_belief_prior.weight": (Hd, D), "fc_embed_belief_prior.bias": (Hd,),
        "fc_state_prior.weight": (2 * S, Hd), "fc_state_prior.bias": (2 * S,),
        "fc_embed_belief_posterior.weight": (Hd, D + E), "fc_embed_belief_posterior.bias": (Hd,),
        "fc_state_posterior.weight": (2 * S, Hd), "fc_state_posterior.bias": (2 * S,),
    }
    w = RssmWeights()
    for field, key in _RSSM_KEYS:
        setattr(w, field, keep(params[key], key, shapes[key]))
    return w


def mlp_struct(params: Dict[str, torch.Tensor], n_layers: int, keep: _Keep, name: str) -> MlpWeights:
    m = MlpWeights()
    m.n_layers = n_layers
    for i in range(n_layers):
        m.w[i] = keep(params[f"fc{i + 1}.weight"], f"{name}.fc{i + 1}.weight")
        m.b[i] = keep(params[f"fc{i + 1}.bias"], f"{name}.fc{i + 1}.bias")
    return m


def act_kind(name: str) -> int:
    if name not in _lib.ACT_KINDS:
        raise RuntimeError(f"activation_function {name!r} is not supported by the CUDA path (relu, elu)")
    return _lib.ACT_KINDS[name]


# ----------------------------------------------------------------------------------------------
def linear(x: torch.Tensor, w: torch.Tensor, b: Optional[torch.Tensor], row_tile: int = 0) -> torch.Tensor:
    """y = x W^T + b: the tcgen05 conv kernel run as a plain GEMM from 256 rows, the layer machine's one-step program
    below that (building block of the hoisted embedding projection and the decoder's fc1)."""
    L = _lib.lib()
    x = _chk(x, "x")
    w = _chk(w, "w")
    rows, in_f = x.shape
    out_f = w.shape[0]
    assert w.shape[1] == in_f
    b = _chk(b, "b", (out_f,)) if b is not None else None
    y = torch.empty(rows, out_f, device=x.device, dtype=torch.float32)
    if rows == 0:
        return y
    ws = torch.empty(L.repo_b200_linear_workspace_bytes(in_f, out_f), dtype=torch.uint8, device=x.device)
    rc = L.repo_b200_linear_fwd(_ptr(x), in_f, rows, in_f, _ptr(w), _ptr(b), out_f, _ptr(y), out_f, _ptr(ws), ws.numel(),
                                row_tile, _stream())
    _lib.check(rc, "repo_b200_linear_fwd")
    return y


def observe_fwd(params: Dict[str, torch.Tensor], prev_belief, prev_state, actions, embeds, nonterms,
                eps_prior, eps_post, act: str = "elu", min_std: float = 0.1, want_kl: bool = True,
                row_tile: int = 0, workspace: Optional[torch.Tensor] = None, packed: bool = False,
                stash: Optional[torch.Tensor] = None):
    """TransitionModel.observe forward (rssm.py:76-146). Returns (list of 7 or 4 tensors, kl (T1,B) or None,
    workspace).  `stash` (T1,B,5D+2H), when given, receives the activations the backward pass needs."""
    L = _lib.lib()
    keep = _Keep()
    d = dims_of(params)
    W = rssm_struct(params, keep)
    T1, B = actions.shape[0], actions.shape[1]
    dev = actions.device
    prev_belief = _chk(prev_belief, "prev_belief", (B, d.belief))
    prev_state = _chk(prev_state, "prev_state", (B, d.state))
    actions = _chk(actions, "actions", (T1, B, d.action))
    with_obs = embeds is not None
    if with_obs:
        embeds = _chk(embeds, "observations", (T1, B, d.embed))
        eps_post = _chk(eps_post, "eps_post", (T1, B, d.state))
    if nonterms is not None:
        nonterms = _chk(nonterms.reshape(T1, B), "nonterminals", (T1, B))
    eps_prior = _chk(eps_prior, "eps_prior", (T1, B, d.state))
    mk = lambda f: torch.empty(T1, B, f, device=dev, dtype=torch.float32)
    outs = [mk(d.belief)] + [mk(d.state) for _ in range(6 if with_obs else 3)]
    kl = torch.empty(T1, B, device=dev, dtype=torch.float32) if (with_obs and want_kl) else None
    if T1 == 0 or B == 0:
        return outs, kl, workspace
    need = L.repo_b200_observe_workspace_bytes(C.byref(d), T1, B)
    if workspace is None or workspace.numel() < need:
        workspace = torch.empty(need, dtype=torch.uint8, device=dev)
        packed = False
    o = outs + [None] * (7 - len(outs))
    rc = L.repo_b200_observe_fwd(
        C.byref(d), C.byref(W), _ptr(prev_belief), _ptr(prev_state), _ptr(actions), _ptr(embeds), _ptr(nonterms),
        _ptr(eps_prior), _ptr(eps_post), *[_ptr(t) for t in o], _ptr(kl), _ptr(stash), T1, B, act_kind(act), float(min_std),
        _ptr(workspace), workspace.numel(), _lib.WEIGHTS_PACKED if packed else 0, row_tile, _stream())
    _lib.check(rc, "repo_b200_observe_fwd")
    return outs, kl, workspace


def imagine_fwd(params: Dict[str, torch.Tensor], actor: Dict[str, torch.Tensor],
                reward: Optional[Dict[str, torch.Tensor]], value: Optional[Dict[str, torch.Tensor]],
                belief, state, eps_action, eps_prior, horizon: int, act: str = "elu", min_std: float = 0.1,
                mean_scale: float = 5.0, init_std: float = 0.0, actor_min_std: float = 0.1,
                gamma: float = 0.99, lambda_: float = 0.95, row_tile: int = 0,
                workspace: Optional[torch.Tensor] = None, packed: bool = False, want_actions: bool = True,
                stash: Optional[torch.Tensor] = None, cond: Optional[torch.Tensor] = None):
    """TransitionModel.imagine (rssm.py:148-184) + reward/value heads + lambda-return in one launch.
    Returns dict(beliefs, prior_states, prior_means, prior_std_devs, actions, rewards, values, returns).
    `cond` (N, C): ConditionalTransitionModel.imagine (rssm.py:225-248) — `params` then belong to a model built with
    action_size + C pseudo-actions and `actor` to a ConditionalActorModel (fc1 over [belief | state | condition])."""
    L = _lib.lib()
    keep = _Keep()
    d = dims_of(params)
    W = rssm_struct(params, keep)
    Am = mlp_struct(actor, 5, keep, "actor")
    Rm = mlp_struct(reward, 4, keep, "reward") if reward is not None else None
    Vm = mlp_struct(value, 4, keep, "value") if value is not None else None
    N = belief.shape[0]
    T = horizon - 1
    dev = belief.device
    belief = _chk(belief, "prev_belief", (N, d.belief))
    state = _chk(state, "prev_state", (N, d.state))
    csz = 0 if cond is None else cond.shape[1]
    a_act = d.action - csz
    if cond is not None:
        cond = _chk(cond, "condition", (N, csz))
        if a_act < 1 or actor["fc1.weight"].shape[1] != d.belief + d.state + csz or actor["fc5.weight"].shape[0] != 2 * a_act:
            raise RuntimeError("conditional imagine: model / actor / condition sizes do not fit together")
    eps_action = _chk(eps_action, "eps_action", (T, N, a_act))
    eps_prior = _chk(eps_prior, "eps_prior", (T, N, d.state))
    mk = lambda *s: torch.empty(*s, device=dev, dtype=torch.float32)
    out = dict(beliefs=mk(T, N, d.belief), prior_states=mk(T, N, d.state), prior_means=mk(T, N, d.state),
               prior_std_devs=mk(T, N, d.state))
    out["actions"] = mk(T, N, a_act) if want_actions else None
    out["rewards"] = mk(T, N) if Rm is not None else None
    out["values"] = mk(T, N) if Vm is not None else None
    out["returns"] = mk(max(T - 1, 0), N) if (Rm is not None and Vm is not None) else None
    if N == 0 or T == 0:
        out["workspace"] = workspace
        return out
    need = L.repo_b200_imagine_workspace_bytes(C.byref(d))
    if workspace is None or workspace.numel() < need:
        workspace = torch.empty(need, dtype=torch.uint8, device=dev)
        packed = False
    rc = L.repo_b200_imagine_cond_fwd(
        C.byref(d), C.byref(W), C.byref(Am), C.byref(Rm) if Rm is not None else None,
        C.byref(Vm) if Vm is not None else None, _ptr(belief), _ptr(state), _ptr(cond), csz, _ptr(eps_action), _ptr(eps_prior),
        _ptr(out["beliefs"]), _ptr(out["prior_states"]), _ptr(out["prior_means"]), _ptr(out["prior_std_devs"]),
        _ptr(out["actions"]), _ptr(out["rewards"]), _ptr(out["values"]), _ptr(out["returns"]),
        horizon, N, act_kind(act), float(min_std), float(mean_scale), float(init_std), float(actor_min_std),
        float(gamma), float(lambda_), _ptr(stash), _ptr(workspace), workspace.numel(), _lib.WEIGHTS_PACKED if packed else 0,
        row_tile, _stream())
    _lib.check(rc, "repo_b200_imagine_cond_fwd")
    out["workspace"] = workspace
    return out


def colsum(x: torch.Tensor) -> torch.Tensor:
    """x.sum(0) of a 2-D fp32 CUDA tensor whose rows are contiguous (a column window of a wider row-major matrix is fine):
    the bias gradients of the hand-written backward passes (autograd's grad_output.sum(0) in the reference)."""
    if x.dim() != 2 or x.dtype != torch.float32 or not x.is_cuda or (x.shape[1] > 1 and x.stride(1) != 1):
        raise RuntimeError("colsum: expected a 2-D fp32 CUDA tensor with contiguous rows")
    rows, cols = x.shape
    ld = x.stride(0) if rows > 1 else max(cols, 1)
    if ld < cols:
        raise RuntimeError("colsum: overlapping rows")
    out = torch.empty(cols, device=x.device, dtype=torch.float32)
    rc = _lib.lib().repo_b200_colsum(_ptr(x), rows, cols, ld, _ptr(out), _stream())
    _lib.check(rc, "repo_b200_colsum")
    return out


def head_fwd(head: Dict[str, torch.Tensor], belief: torch.Tensor, state: torch.Tensor, act: str = "elu",
             row_tile: int = 0) -> torch.Tensor:
    """RewardModel / ValueModel forward on (N, D), (N, S) -> (N,)  (decoder.py:189-195, actor_critic.py:20-26)."""
    L = _lib.lib()
    keep = _Keep()
    N = belief.shape[0]
    D, S, Hd = belief.shape[1], state.shape[1], head["fc1.weight"].shape[0]
    if head["fc1.weight"].shape[1] != D + S:
        raise RuntimeError(f"head: fc1 expects {head['fc1.weight'].shape[1]} inputs, got belief {D} + state {S}")
    d = Dims(D, S, 1, Hd, 1)
    M = mlp_struct(head, 4, keep, "head")
    belief = _chk(belief, "belief", (N, D))
    state = _chk(state, "state", (N, S))
    out = torch.empty(N, device=belief.device, dtype=torch.float32)
    if N == 0:
        return out
    ws = torch.empty(L.repo_b200_head_workspace_bytes(C.byref(d)), dtype=torch.uint8, device=belief.device)
    rc = L.repo_b200_head_fwd(C.byref(d), C.byref(M), _ptr(belief), _ptr(state), _ptr(out), N, act_kind(act),
                              _ptr(ws), ws.numel(), 0, row_tile, _stream())
    _lib.check(rc, "repo_b200_head_fwd")
    return out


def tanh_normal_entropy(mean: torch.Tensor, std: torch.Tensor, eps: torch.Tensor) -> torch.Tensor:
    """SampleDist.entropy of the tanh-Normal policy (models/utils.py:137-163): mean, std (M,A), eps (K,M,A) -> (M,)."""
    L = _lib.lib()
    M, A = mean.shape
    K = eps.shape[0]
    mean, std = _chk(mean, "mean", (M, A)), _chk(std, "std", (M, A))
    eps = _chk(eps, "eps", (K, M, A))
    out = torch.empty(M, device=mean.device, dtype=torch.float32)
    if M == 0:
        return out
    rc = L.repo_b200_tanh_normal_entropy_fwd(_ptr(mean), _ptr(std), _ptr(eps), _ptr(out), M, A, K, _stream())
    _lib.check(rc, "repo_b200_tanh_normal_entropy_fwd")
    return out


def cell_fwd(params: Dict[str, torch.Tensor], belief: torch.Tensor, embed: Optional[torch.Tensor], eps: torch.Tensor,
             act: str = "elu", min_std: float = 0.1):
    """compute_prior_state (embed None, rssm.py:42-50) or compute_posterior_state (rssm.py:52-64):
    returns (state, mean, std_dev), each (N, S)."""
    L = _lib.lib()
    keep = _Keep()
    d = dims_of(params)
    W = rssm_struct(params, keep)
    N = belief.shape[0]
    belief = _chk(belief, "belief", (N, d.belief))
    eps = _chk(eps, "eps", (N, d.state))
    if embed is not None:
        embed = _chk(embed, "observation", (N, d.embed))
    outs = [torch.empty(N, d.state, device=belief.device, dtype=torch.float32) for _ in range(3)]
    if N == 0:
        return tuple(outs)
    ws = torch.empty(L.repo_b200_cell_workspace_bytes(C.byref(d), N), dtype=torch.uint8, device=belief.device)
    rc = L.repo_b200_cell_fwd(C.byref(d), C.byref(W), _ptr(belief), _ptr(embed), _ptr(eps), *[_ptr(o) for o in outs], N,
                              act_kind(act), float(min_std), _ptr(ws), ws.numel(), _stream())
    _lib.check(rc, "repo_b200_cell_fwd")
    return tuple(outs)
