"""Parameter holders with the reference's names for the modules the RSSM kernels read:
ActorModel (actor_critic.py:50-102), ValueModel (actor_critic.py:9-26), RewardModel
(decoder.py:178-195).  Constructors, submodule names and `state_dict()` match the reference so its
checkpoints load; `TransitionModel.imagine` accepts either these or the reference's own instances
(it only reads fc1..fc5 and the `_mean_scale/_init_std/_min_std` attributes).
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import ops


def bottle(f, xs):
    """models/utils.py:9-16 — apply f on (T*B, ...) views of (T, B, ...) tensors."""
    horizon, batch_size = xs[0].shape[:2]
    ys = f(*(x.reshape(horizon * batch_size, *x.shape[2:]) for x in xs))
    if isinstance(ys, tuple):
        return tuple(y.reshape(horizon, batch_size, *y.shape[1:]) for y in ys)
    return ys.reshape(horizon, batch_size, *ys.shape[1:])


class _ScalarHead(nn.Module):
    def __init__(self, belief_size, state_size, hidden_size, activation_function="relu"):
        super().__init__()
        ops.act_kind(activation_function)
        self.activation_function = activation_function
        self.fc1 = nn.Linear(belief_size + state_size, hidden_size)
        self.fc2 = nn.Linear(hidden_size, hidden_size)
        self.fc3 = nn.Linear(hidden_size, hidden_size)
        self.fc4 = nn.Linear(hidden_size, 1)

    def forward(self, belief, state):
        if torch.is_grad_enabled() and (belief.requires_grad or state.requires_grad or
                                        any(p.requires_grad for p in self.parameters())):
            raise NotImplementedError(f"{type(self).__name__}.forward: backward kernels are not built yet; "
                                      "call under torch.no_grad()")
        return ops.head_fwd({k: v for k, v in self.named_parameters()}, belief, state, act=self.activation_function)


class RewardModel(_ScalarHead):
    """decoder.py:178-195."""


class ValueModel(_ScalarHead):
    """actor_critic.py:9-26."""


class ActorModel(nn.Module):
    """actor_critic.py:50-102 — tanh-Normal policy; hidden activation fixed by `activation_function`
    (the trainers' positional-arg slip leaves it at "elu": dreamer.py:99-105)."""

    def __init__(self, belief_size, state_size, hidden_size, action_size, dist="tanh_normal",
                 activation_function="elu", min_std=0.1, init_std=0.0, mean_scale=5):
        super().__init__()
        if activation_function != "elu":
            raise RuntimeError("the fused imagine kernel implements the ELU actor the reference trainers build")
        self.fc1 = nn.Linear(belief_size + state_size, hidden_size)
        self.fc2 = nn.Linear(hidden_size, hidden_size)
        self.fc3 = nn.Linear(hidden_size, hidden_size)
        self.fc4 = nn.Linear(hidden_size, hidden_size)
        self.fc5 = nn.Linear(hidden_size, 2 * action_size)
        self._dist = dist
        self._min_std = min_std
        self._init_std = init_std
        self._mean_scale = mean_scale
