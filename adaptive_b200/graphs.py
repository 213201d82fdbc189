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
    targets); its values are copied into static buffers before each replay.

    Construct it while no autograd graph of an earlier EAGER step over the same parameters is still alive (a kept ``loss`` /
    ``packed`` tensor): torch reuses such a graph's AccumulateGrad nodes, which stay bound to the stream they were created on,
    and the capture (on its own stream) is then invalidated by the synchronisation with that stream."""

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
        # (bf16 path: the loss is fused into the vocabulary projection's epilogue, the packed logits are never written)
        loss = self.model.forward_loss((b["V"], b["v_g"], (b["h0"], b["c0"])), b["captions"], self.lengths, b["tgt"])
        loss.backward()
        return loss

    def __call__(self, batch: Dict[str, torch.Tensor]) -> torch.Tensor:
        F_aa.copy_multi([self.static[k] for k in self.KEYS], [batch[k] for k in self.KEYS])   # one launch for the six inputs
        for p, g in zip(self.params, self.grads):
            p.grad = g
        self.graph.replay()
        return self.loss


class GraphedSampler:
    """``Encoder2Decoder.sampler`` (adaptive_attention.py:168-216) for a fixed batch shape, captured once as a CUDA graph.

    ``aa_greedy_decode`` is a host loop that enqueues ~9 launches per step and never synchronises, so the whole
    ``max_len``-step loop (prologue included) captures into one graph: one launch per batch instead of ~180, which is what
    bounds small batches (the reference's evaluation batch is 400, ``cfg_wzn.py:84``).  ``beam >= 1`` captures
    ``beam_sampler`` instead.  ``__call__(encoded) -> (ids, attention, Beta)`` copies the inputs into the static buffers and
    replays; the returned tensors are the graph's static outputs (overwritten by the next call)."""

    KEYS = ("V", "v_g", "h0", "c0")

    def __init__(self, model, example: Dict[str, torch.Tensor], max_len: int = 30, beam: int = 0, warmup: int = 2):
        self.model, self.max_len, self.beam = model, int(max_len), int(beam)
        self.static = {k: example[k].clone() for k in self.KEYS}
        cur = torch.cuda.current_stream()
        side = torch.cuda.Stream()
        side.wait_stream(cur)
        with torch.cuda.stream(side):
            for _ in range(max(1, warmup)):
                self._eager()
        cur.wait_stream(side)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.out = self._eager()

    def _eager(self):
        b = self.static
        enc = (b["V"], b["v_g"], (b["h0"], b["c0"]))
        if self.beam >= 1:
            return self.model.beam_sampler(enc, beam=self.beam, max_len=self.max_len)
        return self.model.sampler(enc, max_len=self.max_len)

    def __call__(self, encoded):
        if isinstance(encoded, dict):
            src = [encoded[k] for k in self.KEYS]
        else:
            V, v_g, (h0, c0) = encoded
            src = [V, v_g, h0, c0]
        src = [s.reshape(d.shape) for s, d in zip(src, (self.static[k] for k in self.KEYS))]
        F_aa.copy_multi([self.static[k] for k in self.KEYS], src)
        self.graph.replay()
        return self.out

    def replay(self):
        """Replay on whatever ``self.static`` holds (a caller that uploads its batches straight into the static buffers saves the
        device-to-device copy of ``__call__`` -- 411 MB of V per batch of 4096 at config 3)."""
        self.graph.replay()
        return self.out
