"""Optimizer step next to the hot path.

The reference (``train.py:199-216``, ``model_factory.py:62-77``, ``cfg_wzn.py:47-51``) clips the LSTM's gradient norm at 5
and steps ``torch.optim.Adam(params, lr=1e-3, betas=(0.8, 0.999), weight_decay=0)`` over
``encoder.affine_a/affine_b`` + all decoder parameters.  ``FusedClipAdam`` does both in two launches of
``libadaptive_sm100.so`` (``aa_clip_adam_step``) over every parameter tensor at once; state (``exp_avg``, ``exp_avg_sq``, step)
follows ``torch.optim.Adam`` so that a ``state_dict`` converts either way."""
from __future__ import annotations

import ctypes
from typing import Iterable, Sequence

import torch

from . import _lib

MAX_TENSORS = 24


class _OptTensors(ctypes.Structure):
    _fields_ = [("param", ctypes.c_void_p * MAX_TENSORS), ("grad", ctypes.c_void_p * MAX_TENSORS), ("m", ctypes.c_void_p * MAX_TENSORS),
                ("v", ctypes.c_void_p * MAX_TENSORS), ("n", ctypes.c_longlong * MAX_TENSORS), ("clip", ctypes.c_int * MAX_TENSORS)]


class FusedClipAdam:
    """``step()`` == ``clip_grad_norm_(clip_params, max_norm); Adam.step()`` of the reference.

    ``params``: every tensor to update (fp32, CUDA, contiguous, with ``.grad`` set at step time);
    ``clip_params``: the subset whose joint gradient norm is clipped (the reference: ``model.decoder.LSTM.parameters()``)."""

    def __init__(self, params: Iterable[torch.Tensor], clip_params: Iterable[torch.Tensor] = (), lr: float = 1e-3,
                 betas: Sequence[float] = (0.8, 0.999), eps: float = 1e-8, weight_decay: float = 0.0, max_norm: float = 5.0,
                 write_clipped_grads: bool = True):
        self.params = [p for p in params]
        clip_ids = {id(p) for p in clip_params}
        if not self.params:
            raise ValueError("no parameters")
        for p in self.params:
            if not p.is_cuda or p.dtype != torch.float32 or not p.is_contiguous():
                raise RuntimeError("FusedClipAdam needs contiguous float32 CUDA parameters (no CPU fallback)")
        self.clip = [1 if id(p) in clip_ids else 0 for p in self.params]
        self.lr, self.betas, self.eps, self.weight_decay, self.max_norm = lr, tuple(betas), eps, weight_decay, max_norm
        self.write_clipped_grads = write_clipped_grads
        self.exp_avg = [torch.zeros_like(p) for p in self.params]
        self.exp_avg_sq = [torch.zeros_like(p) for p in self.params]
        self.step_count = 0
        dev = self.params[0].device
        self._scratch = torch.zeros(1, device=dev, dtype=torch.float32)
        self.grad_norm = torch.zeros(1, device=dev, dtype=torch.float32)    # norm of the clipped group at the last step

    def zero_grad(self, set_to_none: bool = True):
        for p in self.params:
            if set_to_none:
                p.grad = None
            elif p.grad is not None:
                p.grad.zero_()

    @torch.no_grad()
    def step(self):
        lib = _lib.load()
        self.step_count += 1
        dev = self.params[0].device
        st = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        live = [(p, m, v, c) for p, m, v, c in zip(self.params, self.exp_avg, self.exp_avg_sq, self.clip) if p.grad is not None]
        with torch.cuda.device(dev):
            for i0 in range(0, len(live), MAX_TENSORS):
                chunk = live[i0:i0 + MAX_TENSORS]
                t = _OptTensors()
                for i, (p, m, v, c) in enumerate(chunk):
                    g = p.grad
                    if g.dtype != torch.float32 or not g.is_contiguous():
                        raise RuntimeError("FusedClipAdam: gradients must be contiguous float32")
                    t.param[i], t.grad[i], t.m[i], t.v[i] = p.data_ptr(), g.data_ptr(), m.data_ptr(), v.data_ptr()
                    t.n[i], t.clip[i] = p.numel(), c
                # (a clipped group must sit inside one chunk: 24 tensors cover the reference's 17)
                _lib.check(lib.aa_clip_adam_step(ctypes.byref(t), len(chunk), self.step_count, self.lr, self.betas[0], self.betas[1],
                                                 self.eps, self.weight_decay, self.max_norm, int(self.write_clipped_grads),
                                                 ctypes.c_void_p(self._scratch.data_ptr()), ctypes.c_void_p(self.grad_norm.data_ptr()),
                                                 st), "aa_clip_adam_step")

    def state_dict(self):
        return {"step": self.step_count, "exp_avg": [m.clone() for m in self.exp_avg], "exp_avg_sq": [v.clone() for v in self.exp_avg_sq]}

    def load_state_dict(self, sd):
        self.step_count = int(sd["step"])
        for dst, src in zip(self.exp_avg, sd["exp_avg"]):
            dst.copy_(src)
        for dst, src in zip(self.exp_avg_sq, sd["exp_avg_sq"]):
            dst.copy_(src)


def reference_optimizer(model, **kw) -> FusedClipAdam:
    """The reference's decoder optimizer (``model_factory.py:62-66,97-101``): encoder.affine_a/affine_b + every decoder
    parameter, LSTM gradient norm clipped at ``cf.train_lstm_maxnormal`` (5)."""
    params = list(model.encoder.affine_a.parameters()) + list(model.encoder.affine_b.parameters()) + list(model.decoder.parameters())
    return FusedClipAdam(params, clip_params=list(model.decoder.LSTM.parameters()), **kw)
