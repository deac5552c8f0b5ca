"""CUDA-graph capture of a fixed-shape optimisation step.

A 1000-ray mapping step is ~2 ms of GPU work but several hundred host-side launches (autograd
nodes, loss glue, pose maths); replaying a captured graph removes the launch-bound host time.
The step closure must be capture-safe: static shapes, no host sync (no boolean-mask indexing,
``.item()``, ``nonzero``), tensors it reads updated in place between replays.
"""
from __future__ import annotations

import torch


class GraphedStep:
    def __init__(self, fn, warmup: int = 3, device=None, before_capture=None):
        """`before_capture` runs after the warm-up calls and before the capture -- e.g. setting ``.grad = None`` on the
        leaves, so the captured backward hands its gradient buffers over instead of adding into the warm-up's."""
        self.fn = fn
        dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        cur = torch.cuda.current_stream(dev)
        side = torch.cuda.Stream(dev)
        side.wait_stream(cur)
        with torch.cuda.stream(side):
            for _ in range(warmup):
                fn()
        cur.wait_stream(side)
        torch.cuda.synchronize(dev)
        if before_capture is not None:
            before_capture()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.out = fn()

    def __call__(self):
        self.graph.replay()
        return self.out
