"""Autograd glue for `TransitionModel.observe`.

forward  : the persistent observe kernel, which additionally stashes the per-step activations;
backward : `repo_b200_observe_bwd` — one persistent reverse-time kernel (bwd.cuh) that carries the
           recurrent gradients (belief, state) through the GRU / Gaussian heads and emits the gradient
           of every pre-activation per (t,b);
weights  : dW = dpre^T @ layer_input over all (t,row) samples runs on the tcgen05 weight-gradient kernel
           (`conv.wgrad_gemm`: same fp16 hi/lo arithmetic, contraction over rows); bias grads are sums.

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


def _wgrad(dpre, x):
    """dW = dpre^T @ x.  Feature counts that are multiples of 4 run on the tcgen05 weight-gradient kernel; the 1-wide
    scalar heads (a 1 x K result) stay a library GEMV."""
    if dpre.shape[1] % 4 == 0:
        from .conv import wgrad_gemm
        return wgrad_gemm(dpre, x)
    return dpre.t() @ x

PARAM_KEYS = [k for _, k in ops._RSSM_KEYS]


# 0 = auto (cluster kernel for small batches), 1 = cluster kernel or an error, 2 = the per-sequence fp32 kernel (tests)
OBSERVE_BWD_MODE = 0

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
        # small batches: the cluster kernel (transposed weight slices resident in shared memory) needs a workspace
        ws = torch.empty(L.repo_b200_observe_bwd_workspace_bytes(C.byref(d), B), dtype=torch.uint8, device=dev)
        rc = L.repo_b200_observe_bwd_ws(
            C.byref(d), C.byref(W), p(prev_belief.contiguous()), p(beliefs), p(prior_sd), p(post_sd), p(eps_prior), p(eps_post),
            p(nt), p(stash), p(g[0]), p(g[1]), p(g[2]), p(g[3]), p(g[4]), p(g[5]), p(g[6]), p(d_q), p(d_hq), p(d_p), p(d_hp),
            p(d_gi), p(d_gh), p(d_e), p(d_b0), p(d_s0), T1, B, int(ctx.with_obs), ops.act_kind(ctx.act), float(ctx.min_std),
            p(ws), ws.numel(), OBSERVE_BWD_MODE, ops._stream())
        _lib.check(rc, "repo_b200_observe_bwd_ws")

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
                gp[wkey] = _wgrad(flat(dpre), flat(inp))
            if need[bkey]:
                gp[bkey] = ops.colsum(flat(dpre))

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
                w_obs = named["fc_embed_belief_posterior.weight"][:, D:]
                if T1 * B >= 256 and E % 16 == 0 and Hd % 4 == 0:
                    from .conv import _as_input_side, dense_layer, grad_scales
                    g_obs = torch.empty(T1 * B, E, device=dev, dtype=torch.float32)
                    dq = flat(d_hq)
                    dense_layer(dq, w_obs.detach().t().contiguous(), None, g_obs, scales=_as_input_side(grad_scales(dq)))
                    g_obs = g_obs.reshape(T1, B, E)
                else:
                    g_obs = (flat(d_hq) @ w_obs).reshape(T1, B, E)
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


ACTOR_KEYS = [f"fc{i}.{w}" for i in range(1, 6) for w in ("weight", "bias")]


class ImagineFn(torch.autograd.Function):
    """TransitionModel.imagine under autograd (rssm.py:148-184).  Gradients reach the actor's parameters
    (outputs -> frozen-or-not dynamics -> action sample -> actor) and the start rows; the actor's inputs
    are detached exactly like `policy.get_action(belief.detach(), state.detach())` (rssm.py:170)."""

    @staticmethod
    def forward(ctx, act, min_std, horizon, a_scalars, prev_belief, prev_state, eps_action, eps_prior, cond, *params):
        tp, ap = params[:len(PARAM_KEYS)], params[len(PARAM_KEYS):]
        named, anamed = dict(zip(PARAM_KEYS, tp)), dict(zip(ACTOR_KEYS, ap))
        d = ops.dims_of(named)
        N, T = prev_belief.shape[0], horizon - 1
        L = _lib.lib()
        stash = torch.empty(T, N, L.repo_b200_imagine_stash_floats(C.byref(d)), device=prev_belief.device, dtype=torch.float32)
        mean_scale, init_std, a_min_std = a_scalars
        out = ops.imagine_fwd({k: v.detach() for k, v in named.items()}, {k: v.detach() for k, v in anamed.items()}, None, None,
                              prev_belief.detach(), prev_state.detach(), eps_action, eps_prior, horizon, act=act,
                              min_std=min_std, mean_scale=mean_scale, init_std=init_std, actor_min_std=a_min_std, stash=stash,
                              cond=None if cond is None else cond.detach())
        ctx.meta = (act, min_std, horizon, a_scalars, d)
        ctx.cond = None if cond is None else cond.detach()
        outs = (out["beliefs"], out["prior_states"], out["prior_means"], out["prior_std_devs"])
        ctx.save_for_backward(prev_belief, prev_state, eps_action, eps_prior, stash, out["actions"], *outs, *params)
        ctx.mark_non_differentiable(out["actions"])
        return (*outs, out["actions"])

    @staticmethod
    def backward(ctx, g_b, g_s, g_m, g_sd, _g_actions):
        act, min_std, horizon, (mean_scale, init_std, a_min_std), d = ctx.meta
        sv = ctx.saved_tensors
        prev_belief, prev_state, eps_action, eps_prior, stash, actions = sv[:6]
        beliefs, prior_s, prior_m, prior_sd = sv[6:10]
        params = sv[10:]
        tp, ap = params[:len(PARAM_KEYS)], params[len(PARAM_KEYS):]
        named, anamed = dict(zip(PARAM_KEYS, tp)), dict(zip(ACTOR_KEYS, ap))
        D, S, A, Hd = d.belief, d.state, d.action, d.hidden
        cond = ctx.cond
        csz = 0 if cond is None else cond.shape[1]
        A = A - csz                                      # sampled action width
        N, T = prev_belief.shape[0], horizon - 1
        dev = prev_belief.device
        c = lambda g: None if g is None else g.contiguous().float()
        mk = lambda f: torch.empty(T, N, f, device=dev, dtype=torch.float32)
        d_p, d_hp, d_gi, d_gh, d_e = mk(2 * S), mk(Hd), mk(3 * D), mk(3 * D), mk(D)
        # The actor's hidden layers feed nothing back into the recurrence (its inputs are detached, rssm.py:170): from
        # `_DENSE_MIN_ROWS` (t,row) samples the kernel stops at d_a5 and fc5 -> fc2 run afterwards as four dense tcgen05
        # GEMMs over all samples (activation derivative from the stashed h4..h1 fused in the epilogue) — 27 % of the fp32
        # SIMT kernel's per-step work, moved to the tensor cores.
        hoist = T * N >= _DENSE_MIN_ROWS and Hd % 4 == 0 and (2 * A) % 4 == 0
        d_a5 = mk(2 * A)
        d_a4, d_a3, d_a2, d_a1 = (None,) * 4 if hoist else (mk(Hd), mk(Hd), mk(Hd), mk(Hd))
        need_b0, need_s0 = ctx.needs_input_grad[4], ctx.needs_input_grad[5]
        d_b0 = torch.empty(N, D, device=dev) if (need_b0 or need_s0) else None
        d_s0 = torch.empty(N, S, device=dev) if (need_b0 or need_s0) else None
        keep = ops._Keep()
        W = ops.rssm_struct(named, keep)
        Am = ops.mlp_struct(anamed, 5, keep, "actor")
        p = ops._ptr
        rc = _lib.lib().repo_b200_imagine_cond_bwd(
            C.byref(d), C.byref(W), C.byref(Am), csz, p(prev_belief.contiguous()), p(beliefs), p(actions), p(prior_sd), p(eps_prior),
            p(eps_action), p(stash), p(c(g_b)), p(c(g_s)), p(c(g_m)), p(c(g_sd)), p(d_p), p(d_hp), p(d_gi), p(d_gh), p(d_e),
            p(d_a5), p(d_a4), p(d_a3), p(d_a2), p(d_a1), p(d_b0), p(d_s0), horizon, N, ops.act_kind(act), float(min_std),
            float(mean_scale), float(a_min_std), ops._stream())
        _lib.check(rc, "repo_b200_imagine_cond_bwd")
        if hoist:
            from .conv import _as_input_side, dense_layer, grad_scales
            off_a = 5 * D + Hd
            st2 = stash.reshape(T * N, -1)
            g_prev = d_a5.reshape(T * N, 2 * A)
            outs_a = []
            for li, wkey in zip((4, 3, 2, 1), ("fc5.weight", "fc4.weight", "fc3.weight", "fc2.weight")):
                dst = torch.empty(T * N, Hd, device=dev, dtype=torch.float32)
                dense_layer(g_prev, anamed[wkey].detach().t().contiguous(), None, dst,
                            mask=st2[:, off_a + (li - 1) * Hd: off_a + li * Hd], mask_act="elu", scales=_as_input_side(grad_scales(g_prev)))
                outs_a.append(dst.reshape(T, N, Hd))
                g_prev = dst
            d_a4, d_a3, d_a2, d_a1 = outs_a

        flat = lambda x: x.reshape(T * N, -1)
        need = dict(zip(PARAM_KEYS + ["actor." + k for k in ACTOR_KEYS], ctx.needs_input_grad[9:]))
        gp = {k: None for k in need}

        def lin(wkey, bkey, dpre, inp):
            if need[wkey]:
                gp[wkey] = _wgrad(flat(dpre), flat(inp))
            if need[bkey]:
                gp[bkey] = ops.colsum(flat(dpre))

        b_in = torch.cat([prev_belief.unsqueeze(0), beliefs[:-1]], 0)
        s_in = torch.cat([prev_state.unsqueeze(0), prior_s[:-1]], 0)
        off = 5 * D + Hd
        e, hp = stash[..., :D], stash[..., 5 * D:5 * D + Hd]
        h = [stash[..., off + i * Hd: off + (i + 1) * Hd] for i in range(4)]
        crep = [] if cond is None else [cond.unsqueeze(0).expand(T, N, csz)]   # the condition is constant over the horizon
        lin("fc_embed_state_action.weight", "fc_embed_state_action.bias", d_e, torch.cat([s_in, actions] + crep, -1))
        lin("rnn.weight_ih", "rnn.bias_ih", d_gi, e)
        lin("rnn.weight_hh", "rnn.bias_hh", d_gh, b_in)
        lin("fc_embed_belief_prior.weight", "fc_embed_belief_prior.bias", d_hp, beliefs)
        lin("fc_state_prior.weight", "fc_state_prior.bias", d_p, hp)
        lin("actor.fc1.weight", "actor.fc1.bias", d_a1, torch.cat([b_in, s_in] + crep, -1))
        lin("actor.fc2.weight", "actor.fc2.bias", d_a2, h[0])
        lin("actor.fc3.weight", "actor.fc3.bias", d_a3, h[1])
        lin("actor.fc4.weight", "actor.fc4.bias", d_a4, h[2])
        lin("actor.fc5.weight", "actor.fc5.bias", d_a5, h[3])
        return (None, None, None, None, d_b0 if need_b0 else None, d_s0 if need_s0 else None, None, None, None,
                *[gp[k] for k in PARAM_KEYS], *[gp["actor." + k] for k in ACTOR_KEYS])


def imagine(model, prev_belief, prev_state, policy, horizon, eps_action, eps_prior, cond=None):
    tparams = [dict(model.named_parameters())[k] for k in PARAM_KEYS]
    aparams = [dict(policy.named_parameters())[k] for k in ACTOR_KEYS]
    scal = (float(policy._mean_scale), float(policy._init_std), float(policy._min_std))
    *outs, actions = ImagineFn.apply(model.activation_function, model.min_std_dev, horizon, scal, prev_belief, prev_state,
                                     eps_action, eps_prior, cond, *tparams, *aparams)
    return list(outs), actions


_DENSE_MIN_ROWS = 1024   # from here on one tcgen05 GEMM launch per layer beats the fused small-batch kernels


class MlpFn(torch.autograd.Function):
    """fc1..fcL on [belief|state] (RewardModel / ValueModel / ActorModel trunk).  Small batches: forward on the layer
    machine with the hidden activations stashed, backward on the SIMT chain kernel.  From `_DENSE_MIN_ROWS` rows (the
    (H-1)*N imagined rows of an actor-critic update) every layer — forward and data gradient — is one dense tcgen05 GEMM
    (`conv.dense_layer`: bias + ELU/ReLU, or the activation derivative, fused in the epilogue; activations written
    straight into the stash).  Weight gradients: `conv.wgrad_gemm`."""

    @staticmethod
    def forward(ctx, act, out_f, belief, state, *params):
        L_layers = len(params) // 2
        named = {f"fc{i + 1}.{w}": params[2 * i + j] for i in range(L_layers) for j, w in enumerate(("weight", "bias"))}
        N, D, S, Hd = belief.shape[0], belief.shape[1], state.shape[1], params[0].shape[0]
        d = _lib.Dims(D, S, 1, Hd, 1)
        lib = _lib.lib()
        b, s = ops._chk(belief.detach(), "belief", (N, D)), ops._chk(state.detach(), "state", (N, S))
        out = torch.empty(N, out_f, device=b.device, dtype=torch.float32)
        stash = torch.empty(N, (L_layers - 1) * Hd, device=b.device, dtype=torch.float32)
        dense = N >= _DENSE_MIN_ROWS and Hd % 4 == 0
        x0 = None
        if dense:
            from .conv import dense_layer
            pad = (-(D + S)) % 4
            x0 = torch.cat([b, s] + ([torch.zeros(N, pad, device=b.device)] if pad else []), 1)
            h = x0
            for i in range(L_layers):
                w = params[2 * i].detach()
                if i == 0 and pad:
                    w = torch.nn.functional.pad(w, (0, pad))
                last = i == L_layers - 1
                dst = out if last else stash[:, i * Hd:(i + 1) * Hd]
                dense_layer(h, w, params[2 * i + 1].detach().contiguous(), dst, act="none" if last else act)
                h = dst
        elif N:
            keep = ops._Keep()
            M = ops.mlp_struct({k: v.detach() for k, v in named.items()}, L_layers, keep, "mlp")
            ws = torch.empty(lib.repo_b200_mlp_workspace_bytes(C.byref(d), L_layers, out_f), dtype=torch.uint8, device=b.device)
            rc = lib.repo_b200_mlp_fwd(C.byref(d), C.byref(M), ops._ptr(b), ops._ptr(s), ops._ptr(out), out_f, ops._ptr(stash), N,
                                       ops.act_kind(act), ops._ptr(ws), ws.numel(), ops._stream())
            _lib.check(rc, "repo_b200_mlp_fwd")
        ctx.meta = (act, out_f, L_layers, d, dense)
        ctx.save_for_backward(b, s, stash, *params)
        return out

    @staticmethod
    def backward(ctx, g_out):
        act, out_f, L_layers, d, dense = ctx.meta
        b, s, stash, *params = ctx.saved_tensors
        N, Hd = b.shape[0], d.hidden
        g_out = g_out.contiguous().float()
        need_x = ctx.needs_input_grad[2] or ctx.needs_input_grad[3]
        dh = [None] * 4
        dx = None
        if dense:
            from .conv import _as_input_side, dense_layer, grad_scales
            g = g_out
            for i in range(L_layers - 1, 0, -1):       # d h_i = (d pre_{i+1} @ W_{i+1}) * act'(h_i)
                dst = torch.empty(N, Hd, device=b.device, dtype=torch.float32)
                dense_layer(g, params[2 * i].detach().t().contiguous(), None, dst, mask=stash[:, (i - 1) * Hd:i * Hd], mask_act=act,
                            scales=_as_input_side(grad_scales(g)))
                dh[i - 1] = dst
                g = dst
            if need_x:
                nin = d.belief + d.state
                npad = (nin + 15) // 16 * 16
                w1t = torch.nn.functional.pad(params[0].detach().t(), (0, 0, 0, npad - nin)).contiguous()   # (npad, Hd)
                dx = torch.empty(N, npad, device=b.device, dtype=torch.float32)
                dense_layer(g, w1t, None, dx, scales=_as_input_side(grad_scales(g)))
        else:
            named = {f"fc{i + 1}.{w}": params[2 * i + j] for i in range(L_layers) for j, w in enumerate(("weight", "bias"))}
            keep = ops._Keep()
            M = ops.mlp_struct(named, L_layers, keep, "mlp")
            dh = [torch.empty(N, Hd, device=b.device) for _ in range(L_layers - 1)] + [None] * (5 - L_layers)
            dx = torch.empty(N, d.belief + d.state, device=b.device) if need_x else None
            if N:
                rc = _lib.lib().repo_b200_mlp_bwd(C.byref(d), C.byref(M), ops._ptr(stash), ops._ptr(g_out), out_f, ops._ptr(dh[0]),
                                                  ops._ptr(dh[1]), ops._ptr(dh[2]), ops._ptr(dh[3]), ops._ptr(dx), N,
                                                  ops.act_kind(act), ops._stream())
                _lib.check(rc, "repo_b200_mlp_bwd")
        grads = []
        inputs = [torch.cat([b, s], 1)] + [stash[:, i * Hd:(i + 1) * Hd] for i in range(L_layers - 1)]
        dpre = list(dh[:L_layers - 1]) + [g_out]
        for i in range(L_layers):
            need_w, need_b = ctx.needs_input_grad[4 + 2 * i], ctx.needs_input_grad[5 + 2 * i]
            grads.append(_wgrad(dpre[i], inputs[i]) if need_w else None)
            grads.append(ops.colsum(dpre[i]) if need_b else None)
        gb = dx[:, :d.belief].contiguous() if (need_x and ctx.needs_input_grad[2]) else None
        gs = dx[:, d.belief:d.belief + d.state].contiguous() if (need_x and ctx.needs_input_grad[3]) else None
        return (None, None, gb, gs, *grads)


def mlp(module, n_layers, out_f, belief, state, act):
    params = []
    for i in range(1, n_layers + 1):
        fc = getattr(module, f"fc{i}")
        params += [fc.weight, fc.bias]
    return MlpFn.apply(act, out_f, belief, state, *params)


class LinearFn(torch.autograd.Function):
    """y = x W^T + b as a differentiable op on the package's GEMM kernels (encoder `fc` when embedding_size != 1024)."""

    @staticmethod
    def forward(ctx, x, weight, bias):
        ctx.save_for_backward(x, weight)
        return ops.linear(x.detach().contiguous(), weight.detach().contiguous(), bias.detach().contiguous())

    @staticmethod
    def backward(ctx, g):
        x, weight = ctx.saved_tensors
        g = g.contiguous().float()
        gx = gw = gb = None
        if ctx.needs_input_grad[0]:
            from .conv import _as_input_side, dense_layer, grad_scales
            gx = torch.empty(x.shape, device=x.device, dtype=torch.float32)
            # gradients sit below fp16's normal range: rescale the gathered operand by a power of two (conv.grad_scales)
            dense_layer(g, weight.detach().t().contiguous(), None, gx, scales=_as_input_side(grad_scales(g)))
        if ctx.needs_input_grad[1]:
            gw = _wgrad(g, x.detach()) if x.shape[0] else torch.zeros_like(weight)
        if ctx.needs_input_grad[2]:
            gb = ops.colsum(g)
        return gx, gw, gb


class EntropyFn(torch.autograd.Function):
    """SampleDist.entropy (models/utils.py:160-163) of the tanh-Normal policy with explicit noise."""

    @staticmethod
    def forward(ctx, mean, std, eps):
        ctx.save_for_backward(mean, std, eps)
        return ops.tanh_normal_entropy(mean.detach(), std.detach(), eps)

    @staticmethod
    def backward(ctx, g):
        mean, std, eps = ctx.saved_tensors
        M, A = mean.shape
        dm, dsd = torch.empty_like(mean), torch.empty_like(std)
        if M:
            rc = _lib.lib().repo_b200_tanh_normal_entropy_bwd(ops._ptr(mean.contiguous()), ops._ptr(std.contiguous()), ops._ptr(eps),
                                                              ops._ptr(g.contiguous().float()), ops._ptr(dm), ops._ptr(dsd), M, A,
                                                              eps.shape[0], ops._stream())
            _lib.check(rc, "repo_b200_tanh_normal_entropy_bwd")
        return dm, dsd, None
