"""Fused optimiser tail for the RSSM path: bucket all-reduce -> global-norm clip -> Adam, one flat fp32
bucket per parameter group, no host synchronisation (reference: `clip_grad_norm_` + `Adam.step` at
dreamer.py:286-289, 356-359, 370-373 and repo.py:87-96; torch defaults: betas (0.9, 0.999), eps 1e-8).

`FlatAdam` re-points every parameter (and its `.grad`) at a slice of one contiguous buffer, so
* the data-parallel exchange is ONE `all_reduce` on the gradient bucket, no packing pass,
* clip + Adam are two kernels over contiguous memory instead of per-tensor launches and a `.item()`.
`state_dict()` / `load_state_dict()` keep torch.optim.Adam's layout (per-parameter `exp_avg`, `exp_avg_sq`,
`step`) so the reference's checkpoints (dreamer.py:501-542) round-trip.

One deviation: the step counter (bias correction) is ONE per bucket, advanced by every `step()`, whereas
torch.optim.Adam keeps it per parameter and skips parameters whose `.grad` is None.  They differ only for a parameter
that sat out some steps — in the reference that is TIA's distractor reward head, which has no gradient in phase 1 of
the very first iteration (tia.py:150-201), so its count lags the others by one for the whole run (bias correction
1 - 0.9^(n-1) instead of 1 - 0.9^n: 26 % on its first update, < 1e-4 after 90 steps).  `load_state_dict` takes the
largest per-parameter count of a checkpoint and warns when they are not all equal."""
from __future__ import annotations

import ctypes as C
from typing import Iterable, List, Optional

import torch
import torch.distributed as dist

from . import _lib


class FlatAdam:
    def __init__(self, params: Iterable[torch.nn.Parameter], lr: float, betas=(0.9, 0.999), eps: float = 1e-8,
                 max_grad_norm: Optional[float] = None):
        self.params: List[torch.nn.Parameter] = [p for p in params]
        if not self.params:
            raise ValueError("FlatAdam got an empty parameter list")
        dev = self.params[0].device
        if dev.type != "cuda":
            raise RuntimeError("FlatAdam runs on CUDA parameters only (no CPU fallback)")
        self.lr, self.betas, self.eps, self.max_grad_norm = lr, betas, eps, max_grad_norm
        self.numel = sum(p.numel() for p in self.params)
        self.flat = torch.empty(self.numel, device=dev, dtype=torch.float32)
        self.flat_grad = torch.zeros(self.numel, device=dev, dtype=torch.float32)
        self.exp_avg = torch.zeros_like(self.flat)
        self.exp_avg_sq = torch.zeros_like(self.flat)
        self.sqnorm = torch.zeros(1, device=dev, dtype=torch.float32)
        self.step_dev = torch.zeros(1, device=dev, dtype=torch.int32)  # step count on the device: CUDA-graph safe
        off = 0
        self.offsets = []
        for p in self.params:
            n = p.numel()
            self.flat[off:off + n].copy_(p.data.reshape(-1))
            p.data = self.flat[off:off + n].view_as(p)
            p.grad = self.flat_grad[off:off + n].view_as(p)
            self.offsets.append(off)
            off += n

    def zero_grad(self, set_to_none: bool = False):
        # gradients live in the flat bucket; autograd accumulates into the views in place
        self.flat_grad.zero_()
        for p, off in zip(self.params, self.offsets):
            if p.grad is None or p.grad.data_ptr() != self.flat_grad.data_ptr() + 4 * off:
                p.grad = self.flat_grad[off:off + p.numel()].view_as(p)

    def _regather(self):
        """autograd replaces `.grad` when it was None: fold such gradients back into the bucket."""
        for p, off in zip(self.params, self.offsets):
            if p.grad is None:
                p.grad = self.flat_grad[off:off + p.numel()].view_as(p)
            elif p.grad.data_ptr() != self.flat_grad.data_ptr() + 4 * off:
                self.flat_grad[off:off + p.numel()].copy_(p.grad.reshape(-1))
                p.grad = self.flat_grad[off:off + p.numel()].view_as(p)

    @torch.no_grad()
    def step(self):
        self._regather()
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            dist.all_reduce(self.flat_grad, op=dist.ReduceOp.SUM)  # losses are pre-weighted rows_local/rows_global
        L = _lib.lib()
        s = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        p = lambda t: C.c_void_p(t.data_ptr())
        sq = None
        if self.max_grad_norm is not None:
            self.sqnorm.zero_()
            _lib.check(L.repo_b200_sqnorm_accumulate(p(self.flat_grad), self.numel, p(self.sqnorm), s), "repo_b200_sqnorm_accumulate")
            sq = p(self.sqnorm)
        rc = L.repo_b200_adam_clip_step_dev(p(self.flat), p(self.flat_grad), p(self.exp_avg), p(self.exp_avg_sq), self.numel, sq,
                                            float(self.max_grad_norm or 0.0), float(self.lr), float(self.betas[0]),
                                            float(self.betas[1]), float(self.eps), p(self.step_dev), s)
        _lib.check(rc, "repo_b200_adam_clip_step_dev")

    @property
    def step_count(self) -> int:
        """Number of steps taken (reads the device counter: synchronises; checkpoints / tests only)."""
        return int(self.step_dev.item())

    @step_count.setter
    def step_count(self, v: int):
        self.step_dev.fill_(int(v))

    def grad_norm(self) -> torch.Tensor:
        """Global gradient norm seen by the last clipped step (device scalar; no sync)."""
        return self.sqnorm.sqrt()

    # ---- torch.optim.Adam-compatible checkpoints
    def state_dict(self):
        state = {}
        for i, (p, off) in enumerate(zip(self.params, self.offsets)):
            n = p.numel()
            state[i] = {"step": torch.tensor(float(self.step_count)),
                        "exp_avg": self.exp_avg[off:off + n].view_as(p).clone(),
                        "exp_avg_sq": self.exp_avg_sq[off:off + n].view_as(p).clone()}
        group = {"lr": self.lr, "betas": self.betas, "eps": self.eps, "weight_decay": 0, "amsgrad": False,
                 "params": list(range(len(self.params)))}
        return {"state": state, "param_groups": [group]}

    def load_state_dict(self, sd):
        steps = []
        for i, (p, off) in enumerate(zip(self.params, self.offsets)):
            st = sd["state"].get(i)
            if st is None:
                continue
            n = p.numel()
            self.exp_avg[off:off + n].copy_(st["exp_avg"].reshape(-1))
            self.exp_avg_sq[off:off + n].copy_(st["exp_avg_sq"].reshape(-1))
            steps.append(int(float(st["step"])))
        if steps:
            if min(steps) != max(steps):
                import warnings
                warnings.warn(f"FlatAdam keeps one step count per bucket; the checkpoint has per-parameter counts "
                              f"{min(steps)}..{max(steps)} — using {max(steps)} for all of them")
            self.step_count = max(steps)
        g = sd["param_groups"][0]
        self.lr, self.betas, self.eps = g["lr"], tuple(g["betas"]), g["eps"]
