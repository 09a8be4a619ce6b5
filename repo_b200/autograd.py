"""Autograd glue for `TransitionModel.observe`.

forward  : the persistent observe kernel, which additionally stashes the per-step activations;
backward : `repo_b200_observe_bwd` — one persistent reverse-time kernel (bwd.cuh) that carries the
           recurrent gradients (belief, state) through the GRU / Gaussian heads and emits the gradient
           of every pre-activation per (t,b);
weights  : dW = dpre^T @ layer_input are plain batched GEMMs over those tensors (cuBLAS via
           torch.matmul — library GEMMs, not part of the fused recurrence), bias grads are sums.

Reproduces what autograd gives the reference for rssm.py:116-133: gradients reach every
TransitionModel parameter via BPTT over all T-1 steps and reach `observations` (the encoder);
`prior_states` is treated like any other output (its incoming gradient, normally None, is honoured).
Parameters frozen at call time (`requires_grad=False`, e.g. under FreezeParameters) get no gradient.
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional

import torch

from . import _lib, ops

PARAM_KEYS = [k for _, k in ops._RSSM_KEYS]


class ObserveFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, act, min_std, prev_belief, prev_state, actions, observations, nonterminals, eps_prior, eps_post,
                *params):
        named = dict(zip(PARAM_KEYS, params))
        d = ops.dims_of(named)
        T1, B = actions.shape[0], actions.shape[1]
        L = _lib.lib()
        stash = torch.empty(T1, B, L.repo_b200_observe_stash_floats(C.byref(d)), device=actions.device, dtype=torch.float32)
        outs, kl, _ = ops.observe_fwd({k: v.detach() for k, v in named.items()}, prev_belief.detach(), prev_state.detach(),
                                      actions.detach(), None if observations is None else observations.detach(),
                                      nonterminals, eps_prior, eps_post, act=act, min_std=min_std, stash=stash)
        ctx.act, ctx.min_std, ctx.with_obs, ctx.dims = act, min_std, observations is not None, d
        ctx.save_for_backward(prev_belief, prev_state, actions, observations, nonterminals, eps_prior, eps_post, stash,
                              *outs, *params)
        ctx.n_out = len(outs)
        ctx.mark_non_differentiable(kl) if kl is not None else None
        return (*outs, kl) if kl is not None else tuple(outs)

    @staticmethod
    def backward(ctx, *grads):
        sv = ctx.saved_tensors
        prev_belief, prev_state, actions, observations, nonterminals, eps_prior, eps_post, stash = sv[:8]
        outs = sv[8:8 + ctx.n_out]
        params = sv[8 + ctx.n_out:]
        named = dict(zip(PARAM_KEYS, params))
        d = ctx.dims
        D, S, A, Hd, E = d.belief, d.state, d.action, d.hidden, d.embed
        T1, B = actions.shape[0], actions.shape[1]
        dev = actions.device
        g = [None if (gi is None) else gi.contiguous().float() for gi in grads[:ctx.n_out]]
        g += [None] * (7 - len(g))
        beliefs, prior_sd = outs[0], outs[3]
        post_s, post_sd = (outs[4], outs[6]) if ctx.with_obs else (None, None)
        mk = lambda f: torch.empty(T1, B, f, device=dev, dtype=torch.float32)
        d_q, d_hq = (mk(2 * S), mk(Hd)) if ctx.with_obs else (None, None)
        d_p, d_hp, d_gi, d_gh, d_e = mk(2 * S), mk(Hd), mk(3 * D), mk(3 * D), mk(D)
        need_init = ctx.needs_input_grad[2] or ctx.needs_input_grad[3]
        d_b0 = torch.empty(B, D, device=dev) if need_init else None
        d_s0 = torch.empty(B, S, device=dev) if need_init else None
        keep = ops._Keep()
        W = ops.rssm_struct({k: v for k, v in named.items()}, keep)
        nt = None if nonterminals is None else nonterminals.reshape(T1, B).contiguous()
        p = ops._ptr
        L = _lib.lib()
        rc = L.repo_b200_observe_bwd(
            C.byref(d), C.byref(W), p(prev_belief.contiguous()), p(beliefs), p(prior_sd), p(post_sd), p(eps_prior), p(eps_post),
            p(nt), p(stash), p(g[0]), p(g[1]), p(g[2]), p(g[3]), p(g[4]), p(g[5]), p(g[6]), p(d_q), p(d_hq), p(d_p), p(d_hp),
            p(d_gi), p(d_gh), p(d_e), p(d_b0), p(d_s0), T1, B, int(ctx.with_obs), ops.act_kind(ctx.act), float(ctx.min_std),
            ops._stream())
        _lib.check(rc, "repo_b200_observe_bwd")

        # ---- weight gradients: dW = dpre^T @ input over all (t,b) rows (plain GEMMs) ----
        flat = lambda x: x.reshape(T1 * B, -1)
        e, hp, hq = stash[..., :D], stash[..., 5 * D:5 * D + Hd], stash[..., 5 * D + Hd:]
        state_seq = post_s if ctx.with_obs else outs[1]
        state_in = torch.cat([prev_state.unsqueeze(0), state_seq[:-1]], 0)
        if nt is not None:
            state_in = state_in * nt.unsqueeze(-1)
        x_sa = torch.cat([state_in, actions], -1)
        b_prev = torch.cat([prev_belief.unsqueeze(0), beliefs[:-1]], 0)
        need = dict(zip(PARAM_KEYS, ctx.needs_input_grad[9:]))
        gp = {k: None for k in PARAM_KEYS}

        def lin(wkey, bkey, dpre, inp):
            if need[wkey]:
                gp[wkey] = flat(dpre).t() @ flat(inp)
            if need[bkey]:
                gp[bkey] = flat(dpre).sum(0)

        lin("fc_embed_state_action.weight", "fc_embed_state_action.bias", d_e, x_sa)
        lin("rnn.weight_ih", "rnn.bias_ih", d_gi, e)
        lin("rnn.weight_hh", "rnn.bias_hh", d_gh, b_prev)
        lin("fc_embed_belief_prior.weight", "fc_embed_belief_prior.bias", d_hp, beliefs)
        lin("fc_state_prior.weight", "fc_state_prior.bias", d_p, hp)
        g_obs = None
        if ctx.with_obs:
            lin("fc_embed_belief_posterior.weight", "fc_embed_belief_posterior.bias", d_hq, torch.cat([beliefs, observations], -1))
            lin("fc_state_posterior.weight", "fc_state_posterior.bias", d_q, hq)
            if ctx.needs_input_grad[5]:
                g_obs = (flat(d_hq) @ named["fc_embed_belief_posterior.weight"][:, D:]).reshape(T1, B, E)
        return (None, None, d_b0 if ctx.needs_input_grad[2] else None, d_s0 if ctx.needs_input_grad[3] else None,
                None, g_obs, None, None, None, *[gp[k] for k in PARAM_KEYS])


def observe(model, prev_belief, prev_state, actions, observations, nonterminals, eps_prior, eps_post) -> List[torch.Tensor]:
    params = [dict(model.named_parameters())[k] for k in PARAM_KEYS]
    res = ObserveFn.apply(model.activation_function, model.min_std_dev, prev_belief, prev_state, actions, observations,
                          nonterminals, eps_prior, eps_post, *params)
    if observations is not None:
        *outs, kl = res
        model.last_kl = kl
        return list(outs)
    model.last_kl = None
    return list(res)
