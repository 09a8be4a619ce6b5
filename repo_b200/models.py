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
            from . import autograd as _ag
            return _ag.mlp(self, 4, 1, belief, state, self.activation_function).squeeze(1)
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

    def forward(self, belief, state):
        """actor_critic.py:76-87 -> (action_mean, action_std).  The 5-layer trunk runs on the layer machine
        (differentiable: hand-written backward); the two 2A-wide squashing ops are elementwise torch."""
        from . import autograd as _ag
        A2 = self.fc5.out_features
        raw = _ag.mlp(self, 5, A2, belief, state, "elu")
        m_raw, s_raw = torch.chunk(raw, 2, dim=1)
        mean = self._mean_scale * torch.tanh(m_raw / self._mean_scale)
        std = torch.nn.functional.softplus(s_raw + self._init_std) + self._min_std
        return mean, std

    def get_action_dist(self, belief, state):
        return TanhNormalDist(*self.forward(belief, state))

    def get_action(self, belief, state, det=False):
        dist = self.get_action_dist(belief, state)
        return dist.mode() if det else dist.rsample()


class ConditionalActorModel(ActorModel):
    """actor_critic.py:105-148: ActorModel over [belief | state | condition]."""

    def __init__(self, belief_size, state_size, hidden_size, action_size, condition_size, dist="tanh_normal",
                 activation_function="elu", min_std=0.1, init_std=0.0, mean_scale=5):
        super().__init__(belief_size, state_size + condition_size, hidden_size, action_size, dist, activation_function, min_std,
                         init_std, mean_scale)
        self.condition_size = condition_size

    def forward(self, belief, state, condition):
        return super().forward(belief, torch.cat((state, condition), dim=1))

    def get_action_dist(self, belief, state, condition):
        return TanhNormalDist(*self.forward(belief, state, condition))

    def get_action(self, belief, state, condition, det=False):
        dist = self.get_action_dist(belief, state, condition)
        return dist.mode() if det else dist.rsample()


class TanhNormalDist:
    """SampleDist(Independent(TransformedDistribution(Normal(mean,std), TanhBijector), 1)) with the reference's
    100-sample Monte-Carlo statistics (models/utils.py:112-163, actor_critic.py:89-95)."""

    def __init__(self, mean, std, samples=100):
        self.mean_, self.std_, self._samples = mean, std, samples

    def rsample(self, eps=None):
        eps = torch.randn_like(self.mean_) if eps is None else eps
        return torch.tanh(self.mean_ + self.std_ * eps)

    sample = rsample

    def entropy(self, eps=None):
        """-(1/K) sum_k log p(y_k): fused CUDA kernel forward and backward (models/utils.py:160-163)."""
        from . import autograd as _ag
        if eps is None:
            eps = torch.randn((self._samples,) + tuple(self.mean_.shape), device=self.mean_.device)
        return _ag.EntropyFn.apply(self.mean_, self.std_, eps)

    def log_prob(self, y):
        import math
        yc = torch.where(y.abs() <= 1.0, torch.clamp(y, -0.99999997, 0.99999997), y)
        x = torch.atanh(yc)
        base = -((x - self.mean_) ** 2) / (2 * self.std_ ** 2) - self.std_.log() - math.log(math.sqrt(2 * math.pi))
        ladj = 2.0 * (math.log(2.0) - x - torch.nn.functional.softplus(-2.0 * x))
        return (base - ladj).sum(-1)

    def mean(self):
        eps = torch.randn((self._samples,) + tuple(self.mean_.shape), device=self.mean_.device)
        return torch.tanh(self.mean_ + self.std_ * eps).mean(0)

    def mode(self):
        """models/utils.py:149-158: the most likely of 100 samples (acting path, B=1)."""
        eps = torch.randn((self._samples,) + tuple(self.mean_.shape), device=self.mean_.device)
        samples = torch.tanh(self.mean_ + self.std_ * eps)
        idx = torch.argmax(self.log_prob(samples), 0)
        return samples[idx, torch.arange(samples.shape[1], device=samples.device)]


# ------------------------------------------------------------------------------------------------ symbolic observations
class _Mlp3(nn.Module):
    """fc1 -> act -> fc2 -> act -> fc3 on the package's GEMM kernels (differentiable: autograd.LinearFn)."""

    def __init__(self, in_f, hidden, out_f, activation_function):
        super().__init__()
        ops.act_kind(activation_function)
        self.act_fn = getattr(torch.nn.functional, activation_function)
        self.fc1 = nn.Linear(in_f, hidden)
        self.fc2 = nn.Linear(hidden, hidden)
        self.fc3 = nn.Linear(hidden, out_f)

    def _run(self, x):
        from .autograd import LinearFn
        h = self.act_fn(LinearFn.apply(x, self.fc1.weight, self.fc1.bias))
        h = self.act_fn(LinearFn.apply(h, self.fc2.weight, self.fc2.bias))
        return LinearFn.apply(h, self.fc3.weight, self.fc3.bias)


class SymbolicEncoder(_Mlp3):
    """encoder.py:6-18 (pixel_obs=False)."""

    def __init__(self, observation_size, embedding_size, activation_function="relu"):
        super().__init__(observation_size, embedding_size, embedding_size, activation_function)

    def forward(self, observation):
        return self._run(observation)


class SymbolicObservationModel(_Mlp3):
    """decoder.py:6-25 (pixel_obs=False)."""

    def __init__(self, observation_size, belief_size, state_size, embedding_size, activation_function="relu"):
        super().__init__(belief_size + state_size, embedding_size, observation_size, activation_function)

    def forward(self, belief, state):
        return self._run(torch.cat([belief, state], dim=1))


def Encoder(symbolic, observation_size, embedding_size, activation_function="relu"):
    """encoder.py:44-48."""
    if symbolic:
        return SymbolicEncoder(observation_size, embedding_size, activation_function)
    from .conv import VisualEncoder
    return VisualEncoder(embedding_size, activation_function)


def ObservationModel(symbolic, observation_size, belief_size, state_size, embedding_size, activation_function="relu"):
    """decoder.py:51-66."""
    if symbolic:
        return SymbolicObservationModel(observation_size, belief_size, state_size, embedding_size, activation_function)
    from .conv import VisualObservationModel
    return VisualObservationModel(belief_size, state_size, embedding_size, activation_function)


# ------------------------------------------------------------------------------------------------ optional heads
class EnsembleLinearLayer(nn.Module):
    """models/utils.py:19-49: `ensemble_size` independent linear layers, weight (E, in, out), bias (E, 1, out), applied to a
    shared (rows, in) input or to per-member (E, rows, in) inputs.  Each member is one GEMM launch (autograd.LinearFn)."""

    def __init__(self, in_dim, out_dim, ensemble_size, bias=True):
        super().__init__()
        self.in_dim, self.out_dim, self.ensemble_size = in_dim, out_dim, ensemble_size
        self.weight = nn.Parameter(torch.rand(ensemble_size, in_dim, out_dim))
        if bias:
            self.bias = nn.Parameter(torch.rand(ensemble_size, 1, out_dim))
        else:
            self.register_parameter("bias", None)

    def forward(self, x):
        from .autograd import LinearFn
        outs = []
        for e in range(self.ensemble_size):
            xe = x if x.dim() == 2 else x[e]
            b = self.bias[e, 0] if self.bias is not None else torch.zeros(self.out_dim, device=x.device)
            outs.append(LinearFn.apply(xe, self.weight[e].t(), b))
        return torch.stack(outs, 0)


class EnsembleDynamicsModel(nn.Module):
    """models/utils.py:52-80 (disagreement ensemble, `disag_model` flag): (E, rows, belief) next-belief predictions."""

    def __init__(self, belief_size, state_size, action_size, hidden_size, ensemble_size, activation_function="relu", min_std_dev=0.1):
        super().__init__()
        ops.act_kind(activation_function)
        self.act_fn = getattr(torch.nn.functional, activation_function)
        self.min_std_dev = min_std_dev
        self.fc1 = EnsembleLinearLayer(belief_size + state_size + action_size, hidden_size, ensemble_size)
        self.fc2 = EnsembleLinearLayer(hidden_size, hidden_size, ensemble_size)
        self.fc3 = EnsembleLinearLayer(hidden_size, hidden_size, ensemble_size)
        self.fc4 = EnsembleLinearLayer(hidden_size, belief_size, ensemble_size)

    def forward(self, belief, state, action):
        x = torch.cat((belief, state, action), dim=1)
        h = self.act_fn(self.fc1(x))
        h = self.act_fn(self.fc2(h))
        h = self.act_fn(self.fc3(h))
        return self.fc4(h)


class InverseDynamicsModel(nn.Module):
    """models/utils.py:83-109 (`inv_dynamics` flag): action distribution from (belief, state, next belief)."""

    def __init__(self, belief_size, state_size, action_size, hidden_size, activation_function="relu", min_std_dev=0.1):
        super().__init__()
        ops.act_kind(activation_function)
        self.act_fn = getattr(torch.nn.functional, activation_function)
        self.min_std_dev = min_std_dev
        self.fc1 = nn.Linear(belief_size + state_size + belief_size, hidden_size)
        self.fc2 = nn.Linear(hidden_size, hidden_size)
        self.fc3 = nn.Linear(hidden_size, hidden_size)
        self.fc4 = nn.Linear(hidden_size, 2 * action_size)

    def forward(self, belief, state, next_belief):
        from .autograd import LinearFn
        h = torch.cat((belief, state, next_belief), dim=1)
        for fc in (self.fc1, self.fc2, self.fc3):
            h = self.act_fn(LinearFn.apply(h, fc.weight, fc.bias))
        mean, std_dev = torch.chunk(LinearFn.apply(h, self.fc4.weight, self.fc4.bias), chunks=2, dim=1)
        return mean, torch.nn.functional.softplus(std_dev) + self.min_std_dev
