"""Host-input pipelining for the two entry points a caller drives in a loop (training step, sampler).

The reference's loops (``train.py:199-216``, ``utils.py:160-178``) move every batch host -> device and read a result back
once per iteration, serialised with the compute.  On a B200 the train step takes ~0.7 ms and its inputs ~0.17 ms of
PCIe time; a greedy decode of 4096 images takes ~9 ms and its feature maps (432 MB) about as long to upload.
``HostPipeline`` keeps the same per-iteration traffic (every batch is copied from pinned host memory, every result is
read back) but double-buffers it: the copy of batch i+1 runs on a copy stream while batch i computes, and the result
of batch i is read after batch i+1 has been launched.  PyTorch supplies streams, events and pinned buffers; the
compute is whatever ``step`` launches (the C-ABI operators)."""
from __future__ import annotations

from typing import Callable, Dict, Iterable, List, Sequence

import torch


class HostPipeline:
    """``run(host_batches) -> [result per batch]``.

    ``step(device_batch) -> tensor`` launches the work for one batch (e.g. a ``GraphedTrainStep``, a data-parallel step or
    ``lambda b: model.sampler(...)[0]``) and returns the device tensor to read back (loss scalar, sampled ids).  It may
    return a tensor it will overwrite on the next call (a CUDA graph's static output): the pipeline copies it out
    before launching the next step.  ``example`` gives the shapes/dtypes of one host batch (dict of pinned tensors)."""

    def __init__(self, step: Callable[[Dict[str, torch.Tensor]], torch.Tensor], example: Dict[str, torch.Tensor], device,
                 depth: int = 2):
        self.step = step
        self.device = torch.device(device)
        self.depth = depth
        self.keys: Sequence[str] = tuple(example.keys())
        self.stage: List[Dict[str, torch.Tensor]] = [
            {k: torch.empty(v.shape, dtype=v.dtype, device=self.device) for k, v in example.items()} for _ in range(depth)]
        self.copy_stream = torch.cuda.Stream(device=self.device)
        self.copied = [torch.cuda.Event() for _ in range(depth)]
        self.consumed = [torch.cuda.Event() for _ in range(depth)]
        self.read = [torch.cuda.Event() for _ in range(depth)]
        self.out_host: List[torch.Tensor] = [None] * depth   # pinned, allocated on first use (shape of step's result)
        self.h2d_bytes = sum(v.numel() * v.element_size() for v in example.values())
        self.d2h_bytes = 0

    def _upload(self, slot: int, hb: Dict[str, torch.Tensor], first_use: bool):
        with torch.cuda.stream(self.copy_stream):
            if not first_use:
                self.copy_stream.wait_event(self.consumed[slot])      # the step that read this slot has finished
            for k in self.keys:
                self.stage[slot][k].copy_(hb[k], non_blocking=True)
            self.copied[slot].record(self.copy_stream)

    def run(self, host_batches: Iterable[Dict[str, torch.Tensor]]) -> List[torch.Tensor]:
        main = torch.cuda.current_stream(self.device)
        it = iter(host_batches)
        results: List[torch.Tensor] = []
        pending = []            # (slot, index) of steps launched whose result is not read yet
        nxt = next(it, None)
        i = 0
        if nxt is not None:
            self._upload(0, nxt, True)
        while nxt is not None:
            slot = i % self.depth
            cur, nxt = nxt, next(it, None)
            if nxt is not None:                                        # upload of batch i+1 overlaps step i
                self._upload((i + 1) % self.depth, nxt, i + 1 < self.depth)
            main.wait_event(self.copied[slot])
            out = self.step(self.stage[slot])
            self.consumed[slot].record(main)
            if self.out_host[slot] is None:
                self.out_host[slot] = torch.empty(out.shape, dtype=out.dtype, pin_memory=True)
                self.d2h_bytes = out.numel() * out.element_size()
            self.out_host[slot].copy_(out, non_blocking=True)
            self.read[slot].record(main)
            pending.append(slot)
            if len(pending) >= self.depth:                             # read the oldest result (one step late)
                s = pending.pop(0)
                self.read[s].synchronize()
                results.append(self.out_host[s].clone())
            i += 1
        for s in pending:
            self.read[s].synchronize()
            results.append(self.out_host[s].clone())
        return results
