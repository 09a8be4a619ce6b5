"""CPU: the host-side loss reductions (repo_b200/losses.py, pure tensor ops) against torch.distributions written the way
the reference writes them, and against the reference-made lambda-return fixture."""
import math

import numpy as np
import torch
from torch.distributions import Independent, Normal
from torch.distributions.kl import kl_divergence

from repo_b200 import losses
from tests import _cases as C


def _rand(*shape, seed=0):
    return torch.from_numpy(np.random.RandomState(seed).standard_normal(shape).astype(np.float32))


def test_kl_normal_matches_torch_kl_divergence_and_gradients():
    mp, mq = _rand(7, 5, 30, seed=1).requires_grad_(True), _rand(7, 5, 30, seed=2).requires_grad_(True)
    sp = (_rand(7, 5, 30, seed=3).abs() + 0.1).requires_grad_(True)
    sq = (_rand(7, 5, 30, seed=4).abs() + 0.1).requires_grad_(True)
    got = losses.kl_normal(mp, sp, mq, sq).sum(2)
    want = kl_divergence(Independent(Normal(mp, sp), 1), Independent(Normal(mq, sq), 1))     # dreamer.py:278-281
    np.testing.assert_allclose(got.detach().numpy(), want.detach().numpy(), rtol=1e-5, atol=1e-6)
    g1 = torch.autograd.grad(got.mean(), (mp, sp, mq, sq))
    g2 = torch.autograd.grad(want.mean(), (mp, sp, mq, sq))
    for a, b in zip(g1, g2):
        np.testing.assert_allclose(a.numpy(), b.numpy(), rtol=1e-4, atol=1e-7)


def test_free_nats_and_repo_dual_terms():
    kl = _rand(49, 50, seed=5).abs() * 4
    np.testing.assert_allclose(losses.dreamer_kl_loss(kl, 3.0).item(), torch.max(kl, torch.full_like(kl, 3.0)).mean().item(), rtol=1e-6)
    log_beta = torch.tensor(math.log(1e-5), requires_grad=True)
    t = losses.repo_kl_terms(kl, log_beta, prior_train_steps=5, target_kl=3.0)       # repo.py:63-96
    np.testing.assert_allclose(t["kl_div"].item(), kl.mean().item(), rtol=1e-6)
    np.testing.assert_allclose(t["kl_loss"].item(), 1e-5 * (kl.mean().item() - 3.0), rtol=1e-5)
    (g,) = torch.autograd.grad(t["beta_loss"], log_beta)
    np.testing.assert_allclose(g.item(), -(kl.mean().item() - 3.0), rtol=1e-6)       # d/d log_beta of -log_beta * viol


def test_unit_variance_nlls_keep_the_constant():
    pred, tgt = _rand(13, 37, seed=6), _rand(13, 37, seed=7)
    want = -Normal(pred, 1).log_prob(tgt)                                            # repo.py:60-61, dreamer.py:365-368
    np.testing.assert_allclose(losses.normal_unit_nll(pred, tgt).numpy(), want.numpy(), rtol=1e-6, atol=1e-6)
    np.testing.assert_allclose(losses.value_loss(pred, tgt).item(), want.mean().item(), rtol=1e-6)
    rewards, nonterms = _rand(14, 37, 1, seed=8), (_rand(14, 37, 1, seed=9) > -1).float()
    rl = losses.reward_loss(pred, rewards, nonterms)
    want_r = (-Normal(pred, 1).log_prob(rewards[:-1].squeeze(-1)) * nonterms[:-1].squeeze(-1)).mean((0, 1))
    np.testing.assert_allclose(rl.item(), want_r.item(), rtol=1e-6)
    assert 0 < nonterms.mean() < 1                                                   # the mask is exercised


def test_lambda_return_matches_reference_fixture_bit_exact():
    g, _ = C.load("lambda_return")
    for pre, gamma, lam in (("toy_", 0.9, 0.8), ("", 0.99, 0.95)):      # parameters of oracle/make_golden.py
        r, v, boot = (torch.from_numpy(g[pre + k]) for k in ("r", "v", "boot"))
        out = losses.lambda_return(r, v, gamma * torch.ones_like(r), boot, lam)
        assert np.array_equal(out.numpy(), g[pre + "out"])


def test_actor_loss_signs():
    ret, ent, lat = _rand(13, 37, seed=10), torch.tensor(4.0), torch.tensor(35.0)
    np.testing.assert_allclose(losses.actor_loss(ret, ent, lat, 3e-4, 0.0).item(), -ret.mean().item() - 3e-4 * 4.0, rtol=1e-6)
