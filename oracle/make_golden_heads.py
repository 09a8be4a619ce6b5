"""Golden fixture for the optional heads: the reference's Dreamer.train_disag / train_inv_dynamics (dreamer.py:198-239)
and the disagreement bonus of train_actor_critic (dreamer.py:330-339) run UNMODIFIED on CPU; records the logged losses
and the gradients just before clip_grad_norm_ (optimizer steps disabled).  TEST INFRASTRUCTURE ONLY.

    python oracle/make_golden_heads.py      (build container only: needs /root/reference)
"""
import os
import types

import numpy as np
import torch
import torch.nn as nn

from make_golden_trainer import OUT, Dreamer, FakeLogger, Space, config, set_gpu_mode  # noqa: E402  (stubs matplotlib)
from oracle import rssm_oracle as O  # noqa: E402
from oracle.make_golden import NoiseInjector  # noqa: E402
from repo_b200.synth import make_ensemble_params, make_head_rollout  # noqa: E402


def main():
    torch.set_num_threads(1)
    set_gpu_mode(False)
    cfg = config()
    cfg.update(disag_model=True, inv_dynamics=True, disag_coef=0.5)
    env = types.SimpleNamespace(observation_space=Space((24,)), action_space=Space((6,)))
    log = FakeLogger()
    algo = Dreamer(cfg, env, env, log)
    D, S, A, Hd, E = 200, 30, 6, 200, cfg.ensemble_size
    seed = 700
    algo.disag_model.load_state_dict(make_ensemble_params(seed, D + S + A, Hd, D, E))
    algo.inv_dynamics.load_state_dict(O.make_mlp_params(seed + 1, 2 * D + S, cfg.inv_dynamics_hidden_size, 2 * A, 3))
    for opt in (algo.disag_optimizer, algo.inv_dynamics_optimizer, algo.actor_optimizer, algo.value_optimizer):
        opt.step = lambda *a, **k: None
    T, B = 9, 37
    x = make_head_rollout(seed + 2, T, B)
    save = {}
    orig_clip = nn.utils.clip_grad_norm_
    for name, fn, mod in (("disag", algo.train_disag, algo.disag_model), ("inv", algo.train_inv_dynamics, algo.inv_dynamics)):
        grads = {}

        def capture(params, max_norm, *a, _g=grads, _m=mod, **k):
            for kname, p in _m.named_parameters():
                _g[kname] = p.grad.detach().clone().numpy()
            return orig_clip(params, max_norm, *a, **k)

        nn.utils.clip_grad_norm_ = capture
        try:
            fn(x["beliefs"], x["states"], x["actions"], x["nonterms"])
        finally:
            nn.utils.clip_grad_norm_ = orig_clip
        for k, v in grads.items():      # big tensors: every 97th element + the L2 norm (keeps the fixture small)
            flat = v.reshape(-1)
            save[f"{name}_gradnorm_{k}"] = np.float64(np.sqrt((flat.astype(np.float64) ** 2).sum()))
            save[f"{name}_grad_{k}"] = flat[::97].copy() if flat.size > 8192 else v
    # disagreement bonus inside train_actor_critic
    algo.transition_model.load_state_dict(O.make_transition_params(seed + 3))
    algo.actor_model.load_state_dict(O.make_mlp_params(seed + 4, D + S, Hd, 2 * A, 4))
    algo.reward_model.load_state_dict(O.make_mlp_params(seed + 5, D + S, Hd, 1, 3))
    algo.value_model.load_state_dict(O.make_mlp_params(seed + 6, D + S, Hd, 1, 3))
    N, H = 24, cfg.horizon
    y = O.make_imagine_inputs(seed + 7, N, H)
    rs = np.random.RandomState(seed + 8)
    eps_ent = torch.from_numpy(rs.standard_normal((100, (H - 1) * N, A)).astype(np.float32))
    eps_disag = torch.from_numpy(rs.standard_normal(((H - 1) * N, A)).astype(np.float32))
    queue = []
    for t in range(H - 1):
        queue += [y["eps_action"][t], y["eps_prior"][t]]
    queue += [eps_ent, eps_disag]
    with NoiseInjector(queue):
        algo.train_actor_critic(y["belief"], y["state"])
    for k, p in algo.actor_model.named_parameters():
        flat = p.grad.numpy().reshape(-1)
        save["actor_gradnorm_" + k] = np.float64(np.sqrt((flat.astype(np.float64) ** 2).sum()))
        save["actor_grad_" + k] = flat[::97].copy() if flat.size > 8192 else p.grad.numpy()
    save.update({"log_" + k.split("/")[1]: np.float64(v) for k, v in log.rec.items()})
    save.update(meta_seed=seed, meta_T=T, meta_B=B, meta_N=N, meta_H=H, meta_disag_coef=cfg.disag_coef)
    np.savez_compressed(os.path.join(OUT, "train_heads.npz"), **save)
    print({k: float(v) for k, v in log.rec.items()})


if __name__ == "__main__":
    main()
