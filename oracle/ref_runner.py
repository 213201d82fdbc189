"""Workloads of BASELINE.json run on the UNMODIFIED reference modules staged in ``oracle/_ref``
(see ``oracle/build_ref.py``).  TEST / BASELINE INFRASTRUCTURE ONLY — never imported by
``adaptive_b200``.

Used by ``bench.py`` for (a) the CPU arm (``--impl reference`` and ``cpu_baseline``,
``kind = "reference"``) and (b) ``gpu_eager_baseline``: the same modules on ``cuda:0`` through stock
PyTorch eager (cuDNN / cuBLAS), the bar BASELINE.md section 1 names.  Environment shims (the
reference files themselves are untouched):

* ``baseline_attention.Decoder.forward`` allocates its ``hiddens``/``cells`` with ``.cuda()`` whenever
  ``torch.cuda.is_available()`` (``baseline_attention.py:159-164``) and wraps its blocks in
  ``nn.DataParallel`` whenever ``torch.cuda.device_count() > 1`` (``:184-187``).  A CPU run on a GPU box
  therefore reports "no CUDA" to the reference for its duration, and the single-GPU eager run reports
  one device (one process per GPU is this repo's scheme; the reference's DataParallel is not the bar).
* gradients are taken through ``Decoder`` with leaf ``V, v_g, h0, c0`` (``Encoder2Decoder
  .forward`` transposes the encoder states in place, which breaks autograd), batched greedy decoding is
  the loop body of ``sampler`` (``adaptive_attention.py:186-216``) with ``[1,B,H]`` states (Q9).
"""
from __future__ import annotations

import contextlib
import os
import sys
from typing import Sequence

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from adaptive_b200.synth import make_inputs, make_weights  # noqa: E402
from oracle import build_ref  # noqa: E402


@contextlib.contextmanager
def visible_cuda(n_devices: int):
    """What the reference sees of CUDA while it runs: 0 = none (CPU arm), 1 = a single device (no DataParallel)."""
    avail, count = torch.cuda.is_available, torch.cuda.device_count
    torch.cuda.is_available = (lambda: False) if n_devices == 0 else avail
    torch.cuda.device_count = lambda: n_devices
    try:
        yield
    finally:
        torch.cuda.is_available, torch.cuda.device_count = avail, count


def load_decoder(dims, w, device, baseline: bool = False):
    """Reference ``Decoder`` with the synthetic weights of ``adaptive_b200.synth`` loaded (strict)."""
    ada, base = build_ref.import_reference()
    dec = base.Decoder(dims.E, dims.Vc, dims.H) if baseline else ada.Decoder(dims.E, dims.Vc, dims.H, None)
    dec.load_state_dict({k: torch.from_numpy(np.asarray(v, np.float32)) for k, v in w.items()}, strict=True)
    return dec.to(device)


def train_step_fn(dims, B: int, T: int, lengths: Sequence[int], device="cpu", seed_w: int = 123, seed_in: int = 1234,
                  autocast_bf16: bool = False):
    """-> step() running Decoder.forward -> pack_padded_sequence -> mean CE -> backward (``train.py:205-210``) on the
    reference modules; returns the loss (python float on CPU, device scalar on CUDA)."""
    from torch.nn.utils.rnn import pack_padded_sequence

    dev = torch.device(device)
    ncuda = 0 if dev.type == "cpu" else 1
    w = make_weights(dims, seed=seed_w)
    inp = make_inputs(dims, B, T, seed=seed_in)
    dec = load_decoder(dims, w, dev)
    V = torch.from_numpy(inp["V"]).to(dev).requires_grad_(True)
    v_g = torch.from_numpy(inp["v_g"]).to(dev).requires_grad_(True)
    h0 = torch.from_numpy(inp["h0"])[None].to(dev).requires_grad_(True)
    c0 = torch.from_numpy(inp["c0"])[None].to(dev).requires_grad_(True)
    cap = torch.from_numpy(inp["captions"]).to(dev)
    lengths = [int(x) for x in lengths]
    tgt = pack_padded_sequence(cap[:, 1:], lengths, batch_first=True).data
    crit = torch.nn.CrossEntropyLoss()
    leaves = [V, v_g, h0, c0] + list(dec.parameters())

    def step():
        for t in leaves:
            t.grad = None
        with visible_cuda(ncuda):
            if autocast_bf16:
                with torch.autocast("cuda", dtype=torch.bfloat16):
                    scores = dec(V, v_g, cap, (h0, c0))[0]
            else:
                scores = dec(V, v_g, cap, (h0, c0))[0]
            packed = pack_padded_sequence(scores, lengths, batch_first=True)
            loss = crit(packed.data.float(), tgt)
            loss.backward()
        return float(loss.detach()) if ncuda == 0 else loss.detach()

    return step


def greedy_fn(dims, B: int, L: int, device="cpu", seed_w: int = 123, seed_in: int = 4321):
    """-> run() = the sampler loop body of adaptive_attention.py:186-216 for a batch (states [1,B,H]); returns ids [B,L]."""
    dev = torch.device(device)
    ncuda = 0 if dev.type == "cpu" else 1
    w = make_weights(dims, seed=seed_w)
    inp = make_inputs(dims, B, 1, seed=seed_in)
    dec = load_decoder(dims, w, dev).eval()
    V = torch.from_numpy(inp["V"]).to(dev)
    v_g = torch.from_numpy(inp["v_g"]).to(dev)
    h0 = torch.from_numpy(inp["h0"])[None].to(dev)
    c0 = torch.from_numpy(inp["c0"])[None].to(dev)

    @torch.no_grad()
    def run():
        with visible_cuda(ncuda):
            captions = torch.full((B, 1), 1, dtype=torch.int64, device=dev)
            states = (h0.contiguous(), c0.contiguous())
            ids = []
            for _ in range(L):
                scores, _, _, states = dec(V, v_g, captions, states)
                captions = scores.max(2)[1]
                ids.append(captions)
            return torch.cat(ids, dim=1)

    return run


def config1_fn(dims, B: int = 4, T: int = 18, seed_w: int = 123, seed_in: int = 1234):
    """BASELINE config 1 exactly: the reference's ``Encoder2Decoder.forward`` on CPU, ``resnet_conv = Identity`` (synthetic
    ``[B,2048,7,7]`` features enter the heads directly), ``no_grad``; -> run() returning the PackedSequence data."""
    from adaptive_b200.synth import make_encoder_weights, make_features, make_lengths

    ada, _ = build_ref.import_reference()

    class Cf:
        adaptive_word_embed_size, adaptive_lstm_hidden_size, vocab_length = dims.E, dims.H, dims.Vc

    with visible_cuda(0):
        model = build_ref.make_encoder2decoder(ada, Cf())
    sd = {"decoder." + k: torch.from_numpy(v) for k, v in make_weights(dims, seed=seed_w).items()}
    sd.update({"encoder." + k: torch.from_numpy(v) for k, v in make_encoder_weights(dims, 2048, seed=321).items()})
    missing, unexpected = model.load_state_dict(sd, strict=False)
    assert not unexpected and all(m.startswith("encoder.resnet_conv") for m in missing), (missing, unexpected)
    model.eval()
    feats = torch.from_numpy(make_features(B, 2048, (7, 7), seed=4321))
    cap = torch.from_numpy(make_inputs(dims, B, T, seed=seed_in)["captions"])
    lengths = make_lengths(B, T, seed=seed_in)

    @torch.no_grad()
    def run():
        with visible_cuda(0):
            return model(feats, cap, lengths).data

    return run
