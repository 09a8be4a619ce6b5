"""CPU: host-side wiring of the optional Dreamer heads (Agent.train_disag / train_inv_dynamics: row selection by
nonterms[1:-1], loss, gradients) against the fixture made by the reference's own methods (oracle/make_golden_heads.py).
The CUDA GEMM op is replaced by torch.nn.functional.linear for this test only — it checks the host logic, not the
kernels (tests/test_trainer_gpu.py::test_optional_heads_match_reference_trainer runs the real path on a GPU)."""
import numpy as np
import torch

from oracle import rssm_oracle as O
from tests import _cases as C


def test_train_disag_and_inv_dynamics_host_logic(monkeypatch):
    import repo_b200.autograd as ag
    from repo_b200 import synth
    from repo_b200.trainer import Agent, Config

    class TorchLinear:
        @staticmethod
        def apply(x, w, b):
            return torch.nn.functional.linear(x, w, b)

    monkeypatch.setattr(ag, "LinearFn", TorchLinear)
    g, meta = C.load("train_heads")
    seed, T, B = int(meta["seed"]), int(meta["T"]), int(meta["B"])
    agent = Agent(Config(disag_model=True, inv_dynamics=True, disag_coef=float(meta["disag_coef"])), 6, algo="dreamer", device="cpu")
    agent.disag_model.load_state_dict(synth.make_ensemble_params(seed, 236, 200, 200, 6))
    agent.inv_dynamics.load_state_dict(O.make_mlp_params(seed + 1, 430, 512, 12, 3))
    x = synth.make_head_rollout(seed + 2, T, B)
    kept = int((x["nonterms"][1:-1].flatten() == 1).sum())
    assert 0 < kept < (T - 2) * B          # the fixture exercises the row selection
    agent.train_disag(x["beliefs"], x["states"], x["actions"], x["nonterms"], step=False)
    agent.train_inv_dynamics(x["beliefs"], x["states"], x["actions"], x["nonterms"], step=False)
    np.testing.assert_allclose(agent.logs["train/disag_loss"].item(), g["log_disag_loss"], rtol=1e-5)
    np.testing.assert_allclose(agent.logs["train/inv_dyn_loss"].item(), g["log_inv_dyn_loss"], rtol=1e-5)
    for prefix, mod in (("disag", agent.disag_model), ("inv", agent.inv_dynamics)):
        for k, p in mod.named_parameters():
            want = g[f"{prefix}_grad_{k}"]
            got = p.grad.numpy()
            if want.shape != got.shape:
                got = got.reshape(-1)[::97]
            np.testing.assert_allclose(got, want, rtol=1e-3, atol=1e-6 * np.abs(want).max(), err_msg=f"{prefix} {k}")
