"""Scalar loss reductions that hang off the observe kernel's per-(t,b) KL.

The heavy part — KL(posterior||prior) summed over the state dimension for every (t,b) — is produced
by the observe launch (`TransitionModel.last_kl`).  What is left are O(T*B) reductions to scalars,
kept on the device as 0-d tensors (no `.item()` syncs)."""
from __future__ import annotations

import math
from typing import Dict

import torch


def dreamer_kl_loss(kl_tb: torch.Tensor, free_nats: float = 3.0) -> torch.Tensor:
    """dreamer.py:278-282 / tia.py:160-170: max(kl, free_nats) per (t,b), then mean."""
    return torch.clamp(kl_tb, min=free_nats).mean()


def repo_kl_terms(kl_tb: torch.Tensor, log_beta: torch.Tensor, prior_train_steps: int = 5,
                  target_kl: float = 3.0) -> Dict[str, torch.Tensor]:
    """repo.py:63-96 forward values.  kl_prior and kl_post have the same forward value (they differ only in
    their stop-gradients), so kl_div = alpha*kl + (1-alpha)*kl; kl_loss uses beta = exp(log_beta) detached;
    beta_loss drives the dual variable."""
    kl = kl_tb.mean()
    alpha = prior_train_steps / (1 + prior_train_steps)
    kl_div = alpha * kl + (1 - alpha) * kl
    kl_viol = kl_div - target_kl
    beta = log_beta.detach().exp()
    return {"kl_div": kl_div, "kl_viol": kl_viol, "kl_loss": beta * kl_viol, "beta": beta,
            "beta_loss": -log_beta * kl_viol.detach()}


def kl_normal(mean_p, std_p, mean_q, std_q) -> torch.Tensor:
    """KL(N(mean_p, std_p) || N(mean_q, std_q)) elementwise, torch's `_kl_normal_normal` formula
    (torch/distributions/kl.py; called at repo.py:63-79, dreamer.py:278-281).  Differentiable tensor ops."""
    var_ratio = (std_p / std_q) ** 2
    t1 = ((mean_p - mean_q) / std_q) ** 2
    return 0.5 * (var_ratio + t1 - 1 - var_ratio.log())


def normal_unit_nll(pred: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
    """-Normal(pred, 1).log_prob(target) elementwise, constant kept (repo.py:60-61, dreamer.py:365-368)."""
    return 0.5 * (pred - target) ** 2 + 0.5 * math.log(2 * math.pi)


def reward_loss(reward_pred: torch.Tensor, rewards: torch.Tensor, nonterms: torch.Tensor) -> torch.Tensor:
    """repo.py:58-61: rewards[:-1], masked by nonterms[:-1], mean over (t,b)."""
    tgt = rewards[:-1].squeeze(-1)
    mask = nonterms[:-1].squeeze(-1)
    return (normal_unit_nll(reward_pred, tgt) * mask).mean()


def actor_loss(returns: torch.Tensor, action_entropy: torch.Tensor, latent_entropy: torch.Tensor,
               action_ent_coef: float = 3e-4, latent_ent_coef: float = 0.0) -> torch.Tensor:
    """dreamer.py:350-354."""
    return -returns.mean() - action_ent_coef * action_entropy - latent_ent_coef * latent_entropy


def value_loss(value_pred: torch.Tensor, returns: torch.Tensor) -> torch.Tensor:
    """dreamer.py:362-368 on imag[:-1]."""
    return normal_unit_nll(value_pred, returns).mean()


def lambda_return(rewards: torch.Tensor, values: torch.Tensor, discounts: torch.Tensor, bootstrap: torch.Tensor,
                  lambda_: float = 0.95) -> torch.Tensor:
    """common/utils.py:61-71 — differentiable (plain tensor ops over the H-2 horizon rows).  The forward-only
    fused version lives at the end of the imagine kernel (`TransitionModel.imagine(..., reward_model=, value_model=)`)."""
    next_values = torch.cat([values[1:], bootstrap[None]], 0)
    inputs = rewards + discounts * next_values * (1 - lambda_)
    last = bootstrap
    outputs = []
    for t in reversed(range(inputs.shape[0])):
        last = inputs[t] + discounts[t] * lambda_ * last
        outputs.append(last)
    return torch.stack(list(reversed(outputs)), 0)
