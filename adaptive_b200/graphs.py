"""CUDA-graph capture of a whole training step (forward -> cross-entropy -> backward).

The step is ~150 kernel launches of a few microseconds each; replaying it as one graph removes
the per-launch host cost (and the per-GEMM TMA descriptor encoding) from the critical path.
Shapes, caption lengths and the parameter tensors are frozen at capture time: use one
``GraphedTrainStep`` per (batch, T, lengths) bucket.
"""
from __future__ import annotations

from typing import Dict, Sequence

import torch

from . import functional as F_aa


class GraphedTrainStep:
    """``step(batch) -> loss`` with gradients left in ``p.grad`` of ``model.decoder`` parameters.

    ``batch`` is a dict of device tensors ``V, v_g, h0, c0, captions, tgt`` (``tgt`` = packed
    targets); its values are copied into static buffers before each replay."""

    KEYS = ("V", "v_g", "h0", "c0", "captions", "tgt")

    def __init__(self, model, example: Dict[str, torch.Tensor], lengths: Sequence[int], warmup: int = 3):
        self.model = model
        self.lengths = [int(x) for x in lengths]
        self.static = {k: example[k].clone() for k in self.KEYS}
        self.params = [p for p in model.decoder.parameters()]
        cur = torch.cuda.current_stream()
        side = torch.cuda.Stream()
        side.wait_stream(cur)
        with torch.cuda.stream(side):
            for _ in range(max(1, warmup)):      # also runs every one-time initialisation outside the capture
                self._eager()
        cur.wait_stream(side)
        torch.cuda.synchronize()
        for p in self.params:
            p.grad = None
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.loss = self._eager()
        self.grads = [p.grad for p in self.params]     # static gradient buffers rewritten by every replay

    def _eager(self):
        b = self.static
        for p in self.params:
            p.grad = None
        packed = self.model((b["V"], b["v_g"], (b["h0"], b["c0"])), b["captions"], self.lengths)
        loss = F_aa.cross_entropy(packed.data, b["tgt"])
        loss.backward()
        return loss

    def __call__(self, batch: Dict[str, torch.Tensor]) -> torch.Tensor:
        F_aa.copy_multi([self.static[k] for k in self.KEYS], [batch[k] for k in self.KEYS])   # one launch for the six inputs
        for p, g in zip(self.params, self.grads):
            p.grad = g
        self.graph.replay()
        return self.loss
