"""Golden fixture from the reference's own trainer code: Dreamer.train_actor_critic (dreamer.py:304-381)
run UNMODIFIED on CPU with injected noise; records the logged scalars and the gradients it computes for
the actor and the critic (optimizer steps are disabled so weights stay at their seeded values).

    python oracle/make_golden_trainer.py      (build container only: needs /root/reference)
"""
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")
from oracle import rssm_oracle as O  # noqa: E402
from oracle.make_golden import NoiseInjector  # noqa: E402

# matplotlib is not installed; common/logger.py:11 imports pyplot
mpl = types.ModuleType("matplotlib")
mpl.pyplot = types.ModuleType("matplotlib.pyplot")
mpl.pyplot.figure = lambda *a, **k: None
sys.modules["matplotlib"] = mpl
sys.modules["matplotlib.pyplot"] = mpl.pyplot

from algorithms.repo.dreamer import Dreamer  # noqa: E402
from common.utils import set_gpu_mode  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


class AttrDict(dict):
    __getattr__ = dict.__getitem__
    __setattr__ = dict.__setitem__


class FakeLogger:
    def __init__(self):
        self.rec, self.dir = {}, "/tmp"

    def record(self, k, v, *a, **kw):
        self.rec[k] = v


class Space:
    def __init__(self, shape):
        self.shape = shape


def config():
    c = AttrDict()  # experiments/train_repo.py:8-76 defaults (symbolic observations keep the conv stacks out)
    c.update(algo="dreamer", env_id="dmc-walker-walk", expr_name="golden", seed=0, use_gpu=False, gpu_id=0,
             pixel_obs=False, num_steps=1, replay_size=64, prefill=1, train_every=500, train_steps=100, eval_every=5000,
             checkpoint_every=25000, log_every=500, embedding_size=1024, hidden_size=200, belief_size=200, state_size=30,
             dense_activation_function="elu", cnn_activation_function="relu", batch_size=50, chunk_size=50, horizon=15,
             gamma=0.99, gae_lambda=0.95, action_noise=0.0, action_ent_coef=3e-4, latent_ent_coef=0.0, free_nats=3,
             model_lr=3e-4, actor_lr=8e-5, value_lr=8e-5, grad_clip_norm=100.0, load_checkpoint=False, load_offline=False,
             offline_dir="data", offline_truncate_size=1000000, save_buffer=False, target_kl=3.0, beta_lr=1e-4,
             init_beta=1e-5, prior_train_steps=5, disag_model=False, ensemble_size=6, disag_lr=3e-4, disag_coef=0.0,
             inv_dynamics=False, inv_dynamics_lr=3e-4, inv_dynamics_hidden_size=512, share_repr=False, tia_obs_coef=1.0,
             tia_adv_coef=1.0, tia_reward_train_steps=1)
    return c


def train_dynamics_fixture(algo_name):
    """`Dreamer.train_dynamics` (dreamer.py:241-303) / `RePo.train_dynamics` (repo.py:25-112) run UNMODIFIED on pixel
    observations at a small batch (B=3, T=5); the gradients are captured just before `clip_grad_norm_` and the
    optimiser steps are disabled."""
    from algorithms.repo.repo import RePo
    from algorithms.repo.tia import TIA
    cfg = config()
    # free_nats / init_beta are moved off their defaults so the KL term carries visible gradient at this tiny batch
    cfg.update(algo=algo_name, pixel_obs=True, batch_size=3, chunk_size=5, free_nats=0.1, init_beta=0.3)
    env = types.SimpleNamespace(observation_space=Space((3, 64, 64)), action_space=Space((6,)))
    log = FakeLogger()
    algo = {"repo": RePo, "dreamer": Dreamer, "tia": TIA}[algo_name](cfg, env, env, log)
    D, S, A, Hd = 200, 30, 6, 200
    seed = 500
    algo.transition_model.load_state_dict(O.make_transition_params(seed))
    algo.reward_model.load_state_dict(O.make_mlp_params(seed + 2, D + S, Hd, 1, 3))
    algo.encoder.load_state_dict(O.make_conv_params("encoder", seed + 4))
    algo.obs_model.load_state_dict(O.make_conv_params("decoder", seed + 5, out_channels=6 if algo_name == "tia" else 3))
    groups = [("encoder", algo.encoder), ("transition_model", algo.transition_model), ("obs_model", algo.obs_model),
              ("reward_model", algo.reward_model)]
    if algo_name == "tia":
        algo.distractor_transition_model.load_state_dict(O.make_transition_params(seed + 6))
        algo.distractor_obs_model.load_state_dict(O.make_conv_params("decoder", seed + 7, out_channels=6))
        algo.distractor_only_obs_model.load_state_dict(O.make_conv_params("decoder", seed + 8))
        algo.distractor_reward_model.load_state_dict(O.make_mlp_params(seed + 9, D + S, Hd, 1, 3))
        algo.mask_head.load_state_dict(O.make_mask_head_params(seed + 12))
        groups += [("distractor_transition_model", algo.distractor_transition_model), ("distractor_obs_model", algo.distractor_obs_model),
                   ("distractor_only_obs_model", algo.distractor_only_obs_model), ("distractor_reward_model", algo.distractor_reward_model),
                   ("mask_head", algo.mask_head)]
    algo.model_optimizer.step = lambda *a, **k: None
    if algo_name == "repo":
        algo.beta_optimizer.step = lambda *a, **k: None
    T, B = cfg.chunk_size, cfg.batch_size
    batch = O.make_train_batch(seed + 10, T, B, A)
    eps = O.make_observe_inputs(seed + 11, T, B)
    queue = []
    for t in range(T - 1):
        queue += [eps["eps_prior"][t], eps["eps_post"][t]]
    if algo_name == "tia":  # the distractor model's observe draws after the task model's (tia.py:100-121)
        eps_d = O.make_observe_inputs(seed + 13, T, B)
        for t in range(T - 1):
            queue += [eps_d["eps_prior"][t], eps_d["eps_post"][t]]
    grads = {}
    import torch.nn as nn
    orig_clip = nn.utils.clip_grad_norm_

    def capture(params, max_norm, *a, **k):
        params = list(params)
        if not grads:  # first call = the model loss; TIA's later calls belong to the distractor-reward phase
            for prefix, mod in groups:
                for kname, p in mod.named_parameters():
                    if p.grad is not None:
                        grads[prefix + "." + kname] = p.grad.detach().clone().numpy()
        return orig_clip(params, max_norm, *a, **k)

    nn.utils.clip_grad_norm_ = capture
    try:
        with NoiseInjector(queue):
            beliefs, states = algo.train_dynamics(batch["obs"], batch["actions"], batch["rewards"], batch["nonterms"])
    finally:
        nn.utils.clip_grad_norm_ = orig_clip
    save = {"log_" + k.split("/")[1]: np.float64(v) for k, v in log.rec.items()}
    for k, v in grads.items():  # big tensors: every 97th element + the L2 norm (keeps the fixture small)
        flat = v.reshape(-1)
        save["gradnorm_" + k] = np.float64(np.sqrt((flat.astype(np.float64) ** 2).sum()))
        save["grad_" + k] = flat[::97].copy() if flat.size > 65536 else v
    save["beliefs"] = beliefs.numpy()
    save["posterior_states"] = states.numpy()
    if algo_name == "repo":
        save["grad_log_beta"] = algo.log_beta.grad.numpy()
    save.update(meta_seed=seed, meta_T=T, meta_B=B, meta_free_nats=cfg.free_nats, meta_init_beta=cfg.init_beta)
    np.savez_compressed(os.path.join(OUT, f"train_dynamics_{algo_name}.npz"), **save)
    print(algo_name, {k: float(v) for k, v in log.rec.items()})
    print("  model grad norm", float(np.sqrt(sum((g.astype(np.float64) ** 2).sum() for g in grads.values()))))


def main():
    torch.set_num_threads(1)
    set_gpu_mode(False)
    train_dynamics_fixture("dreamer")
    train_dynamics_fixture("repo")
    train_dynamics_fixture("tia")
    cfg = config()
    env = types.SimpleNamespace(observation_space=Space((24,)), action_space=Space((6,)))
    log = FakeLogger()
    algo = Dreamer(cfg, env, env, log)
    D, S, A, Hd = 200, 30, 6, 200
    seed = 400
    algo.transition_model.load_state_dict(O.make_transition_params(seed))
    algo.actor_model.load_state_dict(O.make_mlp_params(seed + 1, D + S, Hd, 2 * A, 4))
    algo.reward_model.load_state_dict(O.make_mlp_params(seed + 2, D + S, Hd, 1, 3))
    algo.value_model.load_state_dict(O.make_mlp_params(seed + 3, D + S, Hd, 1, 3))
    for opt in (algo.actor_optimizer, algo.value_optimizer):  # keep the gradients, do not move the weights
        opt.step = lambda *a, **k: None
    N, H = 40, cfg.horizon
    x = O.make_imagine_inputs(seed + 20, N, H)
    rs = np.random.RandomState(seed + 30)
    eps_ent = torch.from_numpy(rs.standard_normal((100, (H - 1) * N, A)).astype(np.float32))
    queue = []
    for t in range(H - 1):
        queue += [x["eps_action"][t], x["eps_prior"][t]]
    queue.append(eps_ent)
    with NoiseInjector(queue):
        algo.train_actor_critic(x["belief"], x["state"])
    save = {"log_" + k.split("/")[1]: np.float64(v) for k, v in log.rec.items()}
    for k, p in algo.actor_model.named_parameters():
        save["actor_grad_" + k] = p.grad.numpy()
    for k, p in algo.value_model.named_parameters():
        save["value_grad_" + k] = p.grad.numpy()
    for k, p in algo.transition_model.named_parameters():
        assert p.grad is None, k  # frozen at call time (dreamer.py:306)
    save.update(meta_seed=seed, meta_N=N, meta_H=H)
    np.savez_compressed(os.path.join(OUT, "train_actor_critic.npz"), **save)  # entropy noise is regenerated from meta_seed + 30
    print({k: float(v) for k, v in log.rec.items()})
    print("actor grad norm", float(torch.sqrt(sum((p.grad ** 2).sum() for p in algo.actor_model.parameters()))))


if __name__ == "__main__":
    main()
