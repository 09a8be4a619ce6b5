"""CUDA-graph capture of the update functions (no tracing compiler: the eager code path is recorded once and replayed).

One `train_dynamics` is ~400 kernel launches and one `train_actor_critic` ~350, most of them a few microseconds long; at
the default shapes the host cannot issue them as fast as the GPU retires them.  `GraphedStep` warms a callable up on a
side stream, captures one invocation into a `torch.cuda.CUDAGraph` with static input buffers, and afterwards every call
is: copy the new inputs into those buffers, one graph launch.  Requirements the package meets for this to be valid:
no host synchronisation inside the updates (losses and logs stay device tensors), noise drawn by the CUDA generator
(graph-safe Philox offsets), Adam's step count kept on the device (`FlatAdam.step_dev`), the C-ABI launches on the
current stream and allocates nothing.
"""
from __future__ import annotations

from typing import Any, Callable, Sequence

import torch


class GraphedStep:
    def __init__(self, fn: Callable[..., Any], example_inputs: Sequence[torch.Tensor], warmup: int = 3):
        self.fn = fn
        self.static_inputs = [x.clone() for x in example_inputs]
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(max(1, warmup)):   # also runs every one-time cudaFuncSetAttribute / lazy allocation
                fn(*self.static_inputs)
        torch.cuda.current_stream().wait_stream(side)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.static_outputs = fn(*self.static_inputs)

    def __call__(self, *inputs: torch.Tensor):
        if len(inputs) != len(self.static_inputs):
            raise ValueError(f"expected {len(self.static_inputs)} inputs, got {len(inputs)}")
        for dst, src in zip(self.static_inputs, inputs):
            if dst.shape != src.shape:
                raise ValueError(f"graphed step was captured for shape {tuple(dst.shape)}, got {tuple(src.shape)}")
            if dst.data_ptr() != src.data_ptr():
                dst.copy_(src, non_blocking=True)
        self.graph.replay()
        return self.static_outputs
