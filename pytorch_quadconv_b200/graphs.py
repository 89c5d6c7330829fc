"""
CUDA-graph replay of a whole step (forward, backward, optimiser) built from the drop-in modules.

Small meshes are launch-latency bound: a QuadConv layer step is ~25 kernel launches of a few microseconds each
(SURVEY.md hard part 6), so the host cannot keep the GPU busy -- C1 0.98 ms eager vs 0.19 ms replayed, the 10k-point
autoencoder step 7 ms vs 2.6 ms.  Every op of this package is stream-ordered and allocation-free in steady state
(pair tables and pool indices are cached on the first call), so a step can be captured once and replayed:

    step = graphed(train_step, x_example)        # train_step(x) -> loss ; warm-up + capture
    for x in loader:
        loss = step(x)                            # copy x into the static input, replay, return the static output

Rules of CUDA graphs apply: shapes are fixed, the callable must not synchronise or touch the host, tensors it
returns are static buffers that the next replay overwrites, optimisers need `capturable=True`.  The first forward of
every layer must have happened before capture (the neighbour search reads a count back to the host, like
torch.nonzero in the reference) -- the warm-up calls take care of that.
"""
from __future__ import annotations

from typing import Callable

import torch


class GraphedStep:
    def __init__(self, fn: Callable, example_inputs, warmup: int = 3):
        self._fn = fn
        self._static_in = [t.detach().clone().requires_grad_(t.requires_grad) if isinstance(t, torch.Tensor) else t
                           for t in example_inputs]
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(max(warmup, 1)):
                fn(*self._static_in)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self._static_out = fn(*self._static_in)

    @property
    def static_inputs(self):
        return self._static_in

    def __call__(self, *inputs):
        if len(inputs) != len(self._static_in):
            raise ValueError(f"expected {len(self._static_in)} inputs, got {len(inputs)}")
        for s, t in zip(self._static_in, inputs):
            if isinstance(s, torch.Tensor) and t is not s:
                if t.shape != s.shape or t.dtype != s.dtype:
                    raise ValueError(f"graphed step captured {tuple(s.shape)} {s.dtype}, got {tuple(t.shape)} {t.dtype}")
                s.data.copy_(t.detach(), non_blocking=True)
        self.graph.replay()
        return self._static_out


def graphed(fn: Callable, *example_inputs, warmup: int = 3) -> GraphedStep:
    """Capture `fn(*example_inputs)` (any composition of QuadConv / Mesh_MaxPool layers, loss, backward, optimiser
    step) in a CUDA graph after `warmup` eager calls; the returned object replays it on new inputs of the same shape."""
    if not torch.cuda.is_available():
        raise RuntimeError("pytorch_quadconv_b200.graphed needs a CUDA device")
    return GraphedStep(fn, example_inputs, warmup)
