"""repo_b200 — B200-native (sm_100a) RSSM hot path behind the reference's TransitionModel interface.

    from repo_b200.rssm import TransitionModel          # drop-in for algorithms/repo/models/rssm.py
    from repo_b200.models import ActorModel, ValueModel, RewardModel, bottle

The CUDA library (repo_b200/librepo_b200.so, C-ABI in include/repo_b200.h) is required: importing the
op layer without it raises — there is no CPU / PyTorch fallback."""
from . import _lib  # noqa: F401

__all__ = ["rssm", "models", "ops"]
