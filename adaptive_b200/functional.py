"""Host-side operators over the C ABI: tensors in, tensors out, autograd wired by hand.

PyTorch is used for device memory, streams and autograd bookkeeping only; all arithmetic of
the hot path runs in ``libadaptive_sm100.so``.  Every function requires CUDA tensors and
raises otherwise — there is no CPU path (the CPU restatement lives in ``oracle/`` and is
test infrastructure).
"""
from __future__ import annotations

import ctypes
import os
import weakref
from typing import Dict, Optional, Sequence, Tuple

import torch

from . import _lib
from ._lib import AADims, AAWeightGrads, AAWeights, KEY_TO_FIELD, WEIGHT_FIELDS, check

ATT_DIM = 49
# aa_weights fields the baseline model (baseline_attention.py:66-194) does not have: passed as None / NULL
SENTINEL_FIELDS = ("sen_wx", "sen_wh", "att_ws")


def baseline_weights(w: Sequence[Optional[torch.Tensor]]) -> Tuple[Optional[torch.Tensor], ...]:
    """13-tuple in ``aa_weights`` order with the three sentinel weights dropped (None): the baseline decoder."""
    return tuple(None if name in SENTINEL_FIELDS else t for name, t in zip(WEIGHT_FIELDS, w))



def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def _stream(device) -> ctypes.c_void_p:
    return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def _need_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise RuntimeError("adaptive_b200 operators need CUDA tensors (no CPU fallback); got a %s tensor" % t.device)


def _f32c(t: Optional[torch.Tensor]) -> Optional[torch.Tensor]:
    if t is None:
        return None
    if t.dtype != torch.float32:
        t = t.float()
    return t.contiguous()


def _states2d(st: Optional[torch.Tensor], B: int, H: int) -> Optional[torch.Tensor]:
    """Accept [1,B,H] (nn.LSTM layout), [B,1,H] (what the reference encoder returns, Q9) or [B,H]."""
    if st is None:
        return None
    if st.dim() == 3:
        if st.shape[0] == 1 and st.shape[1] == B:
            st = st[0]
        elif st.shape[1] == 1 and st.shape[0] == B:
            st = st[:, 0]
        else:
            raise ValueError("state of shape %s does not match batch %d" % (tuple(st.shape), B))
    if st.shape != (B, H):
        raise ValueError("state of shape %s, expected (%d, %d)" % (tuple(st.shape), B, H))
    return _f32c(st)


PRECISIONS = {"fp32": _lib.PREC_FP32, "bf16": _lib.PREC_BF16}
# decoding: "fp32" = exact SIMT contractions; "tf32x3" = tcgen05 tensor cores over (hi, lo)-split fp32 operands
DECODE_PRECISIONS = {"fp32": _lib.PREC_FP32, "tf32x3": _lib.PREC_TF32X3}


def make_dims(B, T, k, H, E, Vc, a=ATT_DIM, precision=_lib.PREC_FP32) -> AADims:
    return AADims(B=B, T=T, k=k, a=a, H=H, E=E, Vc=Vc, precision=precision)


def weights_struct(w: Sequence[torch.Tensor]) -> AAWeights:
    s = AAWeights()
    for name, t in zip(WEIGHT_FIELDS, w):
        if t is not None:      # (None = NULL: the sentinel weights of the baseline model)
            setattr(s, name, t.data_ptr())
    return s


def ordered_weights(named: Dict[str, torch.Tensor]) -> Tuple[torch.Tensor, ...]:
    """Order a ``{state_dict key (without 'decoder.') : tensor}`` mapping like ``aa_weights``."""
    inv = {v: k for k, v in KEY_TO_FIELD.items()}
    return tuple(named[inv[f]] for f in WEIGHT_FIELDS)


def _check_weights(w: Sequence[torch.Tensor], H, E, Vc, a):
    shapes = [(Vc, E), (4 * H, 2 * E), (4 * H, H), (4 * H,), (4 * H,), (H, 2 * E), (H, H), (a, H), (a, H), (a, H),
              (1, a), (Vc, H), (Vc,)]
    absent = [name for name, t in zip(WEIGHT_FIELDS, w) if t is None]
    if absent and sorted(absent) != sorted(SENTINEL_FIELDS):
        raise ValueError("weights %s are missing: only sen_wx, sen_wh and att_ws may be None, and only together "
                         "(the sentinel-less baseline decoder)" % absent)
    for name, t, s in zip(WEIGHT_FIELDS, w, shapes):
        if t is None:
            continue
        if tuple(t.shape) != s:
            raise ValueError("weight %s has shape %s, expected %s" % (name, tuple(t.shape), s))
        if t.dtype != torch.float32 or not t.is_contiguous():
            raise ValueError("weight %s must be contiguous float32" % name)
        _need_cuda(t)


class _DecoderFn(torch.autograd.Function):
    """``Decoder.forward`` (baseline_attention.py:148-194) + hand-written backward."""

    @staticmethod
    def forward(ctx, prec, V, v_g, captions, h0, c0, *w):
        lib = _lib.load()
        B, k, H = V.shape
        T = captions.shape[1]
        E = v_g.shape[1]
        Vc = w[0].shape[0]
        a = w[7].shape[0]
        _check_weights(w, H, E, Vc, a)
        dev = V.device
        d = make_dims(B, T, k, H, E, Vc, a, prec)
        scores = torch.empty(B, T, Vc, device=dev, dtype=torch.float32)
        alpha = torch.empty(B, T, k, device=dev, dtype=torch.float32)
        beta = torch.empty(B, T, 1, device=dev, dtype=torch.float32)
        hT = torch.empty(B, H, device=dev, dtype=torch.float32)
        cT = torch.empty(B, H, device=dev, dtype=torch.float32)
        nbytes = lib.aa_decoder_saved_bytes(ctypes.byref(d))
        saved = torch.empty(nbytes, device=dev, dtype=torch.uint8)
        ws = weights_struct(w)
        with torch.cuda.device(dev):
            check(lib.aa_decoder_forward(ctypes.byref(d), ctypes.byref(ws), _ptr(V), _ptr(v_g), _ptr(captions), _ptr(h0),
                                         _ptr(c0), _ptr(scores), _ptr(alpha), _ptr(beta), _ptr(hT), _ptr(cT), _ptr(saved),
                                         nbytes, _stream(dev)), "aa_decoder_forward")
        ctx.dims = (B, T, k, H, E, Vc, a, prec)
        ctx.has_state = (h0 is not None, c0 is not None)
        ctx.save_for_backward(V, v_g, captions, h0 if h0 is not None else V.new_empty(0),
                              c0 if c0 is not None else V.new_empty(0), alpha, beta, saved, *w)
        ctx.set_materialize_grads(False)
        return scores, alpha, beta, hT, cT

    @staticmethod
    def backward(ctx, d_scores, d_alpha, d_beta, d_hT, d_cT):
        lib = _lib.load()
        V, v_g, captions, h0, c0, alpha, beta, saved = ctx.saved_tensors[:8]
        w = ctx.saved_tensors[8:]
        B, T, k, H, E, Vc, a, prec = ctx.dims
        h0 = h0 if ctx.has_state[0] else None
        c0 = c0 if ctx.has_state[1] else None
        dev = V.device
        d = make_dims(B, T, k, H, E, Vc, a, prec)
        if d_scores is None:
            d_scores = torch.zeros(B, T, Vc, device=dev, dtype=torch.float32)
        d_scores, d_alpha, d_beta, d_hT, d_cT = (_f32c(x) for x in (d_scores, d_alpha, d_beta, d_hT, d_cT))
        grads = [torch.empty_like(t) if t is not None else None for t in w]
        gs = AAWeightGrads()
        for name, t in zip(WEIGHT_FIELDS, grads):
            if t is not None:
                setattr(gs, name, t.data_ptr())
        dV = torch.empty_like(V)
        dvg = torch.empty_like(v_g)
        dh0 = torch.empty(B, H, device=dev, dtype=torch.float32) if h0 is not None else None
        dc0 = torch.empty(B, H, device=dev, dtype=torch.float32) if c0 is not None else None
        sbytes = lib.aa_decoder_bwd_scratch_bytes(ctypes.byref(d))
        scratch = torch.empty(sbytes, device=dev, dtype=torch.uint8)
        ws = weights_struct(w)
        with torch.cuda.device(dev):
            check(lib.aa_decoder_backward(ctypes.byref(d), ctypes.byref(ws), _ptr(V), _ptr(v_g), _ptr(captions), _ptr(h0),
                                          _ptr(c0), _ptr(alpha), _ptr(beta), _ptr(saved), saved.numel(), _ptr(d_scores),
                                          _ptr(d_alpha), _ptr(d_beta), _ptr(d_hT), _ptr(d_cT), ctypes.byref(gs), _ptr(dV),
                                          _ptr(dvg), _ptr(dh0), _ptr(dc0), _ptr(scratch), sbytes, _stream(dev)),
                  "aa_decoder_backward")
        return (None, dV, dvg, None, dh0, dc0) + tuple(grads)


def decoder_forward(w: Sequence[torch.Tensor], V, v_g, captions, h0=None, c0=None, precision: str = "fp32"):
    """scores [B,T,Vc], alpha [B,T,k], beta [B,T,1], hT [B,H], cT [B,H] (differentiable).
    ``precision``: "fp32" (exact, the parity path) or "bf16" (tcgen05 tensor cores, fp32 accumulate)."""
    _need_cuda(V, v_g, captions, h0, c0)
    V, v_g = _f32c(V), _f32c(v_g)
    B, _, H = V.shape
    captions = captions.to(torch.int64).contiguous()
    h0, c0 = _states2d(h0, B, H), _states2d(c0, B, H)
    return _DecoderFn.apply(PRECISIONS[precision], V, v_g, captions, h0, c0, *w)


class _DecoderPackedFn(torch.autograd.Function):
    """``Encoder2Decoder.forward`` body (baseline_attention.py:219-230): ``Decoder.forward`` followed by
    ``pack_padded_sequence(scores, lengths, batch_first=True).data``, with the vocabulary projection computed for the packed
    rows only (the reference computes all B*T rows and drops the rest, Q13) -- same values, same order."""

    @staticmethod
    def forward(ctx, prec, V, v_g, captions, h0, c0, row_index, *w):
        lib = _lib.load()
        B, k, H = V.shape
        T = captions.shape[1]
        E = v_g.shape[1]
        Vc = w[0].shape[0]
        a = w[7].shape[0]
        _check_weights(w, H, E, Vc, a)
        dev = V.device
        n = row_index.numel()
        d = make_dims(B, T, k, H, E, Vc, a, prec)
        packed = torch.empty(n, Vc, device=dev, dtype=torch.float32)
        alpha = torch.empty(B, T, k, device=dev, dtype=torch.float32)
        beta = torch.empty(B, T, 1, device=dev, dtype=torch.float32)
        hT = torch.empty(B, H, device=dev, dtype=torch.float32)
        cT = torch.empty(B, H, device=dev, dtype=torch.float32)
        nbytes = lib.aa_decoder_saved_bytes(ctypes.byref(d))
        saved = torch.empty(nbytes, device=dev, dtype=torch.uint8)
        ws = weights_struct(w)
        with torch.cuda.device(dev):
            check(lib.aa_decoder_forward_packed(ctypes.byref(d), ctypes.byref(ws), _ptr(V), _ptr(v_g), _ptr(captions), _ptr(h0),
                                                _ptr(c0), _ptr(row_index), n, _ptr(packed), _ptr(alpha), _ptr(beta), _ptr(hT), _ptr(cT),
                                                _ptr(saved), nbytes, _stream(dev)), "aa_decoder_forward_packed")
        ctx.dims = (B, T, k, H, E, Vc, a, prec)
        ctx.has_state = (h0 is not None, c0 is not None)
        ctx.save_for_backward(V, v_g, captions, h0 if h0 is not None else V.new_empty(0),
                              c0 if c0 is not None else V.new_empty(0), alpha, beta, saved, row_index, *w)
        ctx.set_materialize_grads(False)
        return packed, alpha, beta, hT, cT

    @staticmethod
    def backward(ctx, d_packed, d_alpha, d_beta, d_hT, d_cT):
        lib = _lib.load()
        V, v_g, captions, h0, c0, alpha, beta, saved, row_index = ctx.saved_tensors[:9]
        w = ctx.saved_tensors[9:]
        B, T, k, H, E, Vc, a, prec = ctx.dims
        h0 = h0 if ctx.has_state[0] else None
        c0 = c0 if ctx.has_state[1] else None
        dev = V.device
        n = row_index.numel()
        d = make_dims(B, T, k, H, E, Vc, a, prec)
        if d_packed is None:
            d_packed = torch.zeros(n, Vc, device=dev, dtype=torch.float32)
        # bf16 mirror of the incoming gradient, if the producer left one on the tensor (cross_entropy below does): valid only for
        # this very tensor object, so anything autograd put in between (an accumulation, a copy) simply loses it
        mirror = getattr(d_packed, "_aa_bf16_mirror", None)
        d_packed, d_alpha, d_beta, d_hT, d_cT = (_f32c(x) for x in (d_packed, d_alpha, d_beta, d_hT, d_cT))
        if mirror is not None and not (mirror.shape == d_packed.shape and mirror.dtype == torch.bfloat16 and mirror.is_contiguous()
                                       and getattr(d_packed, "_aa_bf16_mirror", None) is mirror
                                       and getattr(d_packed, "_aa_bf16_mirror_version", -1) == d_packed._version):
            mirror = None
        grads = [torch.empty_like(t) if t is not None else None for t in w]
        gs = AAWeightGrads()
        for name, t in zip(WEIGHT_FIELDS, grads):
            if t is not None:
                setattr(gs, name, t.data_ptr())
        dV = torch.empty_like(V)
        dvg = torch.empty_like(v_g)
        dh0 = torch.empty(B, H, device=dev, dtype=torch.float32) if h0 is not None else None
        dc0 = torch.empty(B, H, device=dev, dtype=torch.float32) if c0 is not None else None
        sbytes = lib.aa_decoder_bwd_scratch_bytes(ctypes.byref(d))
        scratch = torch.empty(sbytes, device=dev, dtype=torch.uint8)
        ws = weights_struct(w)
        with torch.cuda.device(dev):
            check(lib.aa_decoder_backward_packed(ctypes.byref(d), ctypes.byref(ws), _ptr(V), _ptr(v_g), _ptr(captions), _ptr(h0),
                                                 _ptr(c0), _ptr(alpha), _ptr(beta), _ptr(saved), saved.numel(), _ptr(row_index), n,
                                                 _ptr(d_packed), _ptr(d_alpha), _ptr(d_beta), _ptr(d_hT), _ptr(d_cT), ctypes.byref(gs),
                                                 _ptr(dV), _ptr(dvg), _ptr(dh0), _ptr(dc0), _ptr(scratch), sbytes, _stream(dev), None, None,
                                                 None, _ptr(mirror)), "aa_decoder_backward_packed")
        return (None, dV, dvg, None, dh0, dc0, None) + tuple(grads)


def decoder_forward_packed(w: Sequence[torch.Tensor], V, v_g, captions, lengths: Sequence[int], h0=None, c0=None,
                           precision: str = "fp32"):
    """-> (PackedSequence of scores over the valid positions, alpha [B,T,k], beta [B,T,1], hT, cT), differentiable.
    Equal to ``pack_scores(decoder_forward(...)[0], lengths)`` without ever computing the discarded rows."""
    _need_cuda(V, v_g, captions, h0, c0)
    V, v_g = _f32c(V), _f32c(v_g)
    B, _, H = V.shape
    captions = captions.to(torch.int64).contiguous()
    h0, c0 = _states2d(h0, B, H), _states2d(c0, B, H)
    row_index, batch_sizes = cached_row_index(lengths, captions.shape[1], V.device)
    data, alpha, beta, hT, cT = _DecoderPackedFn.apply(PRECISIONS[precision], V, v_g, captions, h0, c0, row_index, *w)
    return torch.nn.utils.rnn.PackedSequence(data, batch_sizes), alpha, beta, hT, cT


class _DecoderLossFn(torch.autograd.Function):
    """``Encoder2Decoder.forward`` + mean cross-entropy over the packed positions (``train.py:205-208``) in ONE operator, the loss
    fused into the vocabulary projection's epilogue (``aa_decoder_forward_loss``): the ``[n_rows, Vc]`` logits are never written."""

    @staticmethod
    def forward(ctx, prec, denom, V, v_g, captions, h0, c0, row_index, targets, *w):
        lib = _lib.load()
        B, k, H = V.shape
        T = captions.shape[1]
        E = v_g.shape[1]
        Vc = w[0].shape[0]
        a = w[7].shape[0]
        _check_weights(w, H, E, Vc, a)
        dev = V.device
        n = row_index.numel()
        d = make_dims(B, T, k, H, E, Vc, a, prec)
        loss = torch.empty((), device=dev, dtype=torch.float32)
        alpha = torch.empty(B, T, k, device=dev, dtype=torch.float32)
        beta = torch.empty(B, T, 1, device=dev, dtype=torch.float32)
        hT = torch.empty(B, H, device=dev, dtype=torch.float32)
        cT = torch.empty(B, H, device=dev, dtype=torch.float32)
        nbytes = lib.aa_decoder_saved_bytes(ctypes.byref(d))
        saved = torch.empty(nbytes, device=dev, dtype=torch.uint8)
        ws = weights_struct(w)
        with torch.cuda.device(dev):
            check(lib.aa_decoder_forward_loss(ctypes.byref(d), ctypes.byref(ws), _ptr(V), _ptr(v_g), _ptr(captions), _ptr(h0), _ptr(c0),
                                              _ptr(row_index), n, _ptr(targets), int(denom), _ptr(loss), _ptr(alpha), _ptr(beta), _ptr(hT),
                                              _ptr(cT), _ptr(saved), nbytes, _stream(dev)), "aa_decoder_forward_loss")
        ctx.dims = (B, T, k, H, E, Vc, a, prec)
        ctx.has_state = (h0 is not None, c0 is not None)
        ctx.save_for_backward(V, v_g, captions, h0 if h0 is not None else V.new_empty(0),
                              c0 if c0 is not None else V.new_empty(0), alpha, beta, saved, row_index, *w)
        ctx.set_materialize_grads(False)
        ctx.consumed = False
        return loss, alpha, beta, hT, cT

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, g_loss, d_alpha, d_beta, d_hT, d_cT):
        lib = _lib.load()
        V, v_g, captions, h0, c0, alpha, beta, saved, row_index = ctx.saved_tensors[:9]
        w = ctx.saved_tensors[9:]
        B, T, k, H, E, Vc, a, prec = ctx.dims
        if ctx.consumed:
            raise RuntimeError("adaptive_b200.decoder_forward_loss: backward twice through the same loss is not supported "
                               "(the stored gradient of the logits is consumed in place)")
        ctx.consumed = True
        h0 = h0 if ctx.has_state[0] else None
        c0 = c0 if ctx.has_state[1] else None
        dev = V.device
        n = row_index.numel()
        d = make_dims(B, T, k, H, E, Vc, a, prec)
        d_alpha, d_beta, d_hT, d_cT = (_f32c(x) for x in (d_alpha, d_beta, d_hT, d_cT))
        grads = [torch.empty_like(t) if t is not None else None for t in w]
        gs = AAWeightGrads()
        for name, t in zip(WEIGHT_FIELDS, grads):
            if t is not None:
                setattr(gs, name, t.data_ptr())
        dV = torch.empty_like(V)
        dvg = torch.empty_like(v_g)
        dh0 = torch.empty(B, H, device=dev, dtype=torch.float32) if h0 is not None else None
        dc0 = torch.empty(B, H, device=dev, dtype=torch.float32) if c0 is not None else None
        sbytes = lib.aa_decoder_bwd_scratch_bytes(ctypes.byref(d))
        scratch = torch.empty(sbytes, device=dev, dtype=torch.uint8)
        ws = weights_struct(w)
        with torch.cuda.device(dev):
            if g_loss is None:
                g_loss = torch.zeros((), device=dev, dtype=torch.float32)
            g_loss = g_loss.to(torch.float32).contiguous()
            check(lib.aa_decoder_loss_grad_scale(ctypes.byref(d), _ptr(saved), saved.numel(), n, _ptr(g_loss), _stream(dev)),
                  "aa_decoder_loss_grad_scale")
            check(lib.aa_decoder_backward_packed(ctypes.byref(d), ctypes.byref(ws), _ptr(V), _ptr(v_g), _ptr(captions), _ptr(h0),
                                                 _ptr(c0), _ptr(alpha), _ptr(beta), _ptr(saved), saved.numel(), _ptr(row_index), n,
                                                 None, _ptr(d_alpha), _ptr(d_beta), _ptr(d_hT), _ptr(d_cT), ctypes.byref(gs),
                                                 _ptr(dV), _ptr(dvg), _ptr(dh0), _ptr(dc0), _ptr(scratch), sbytes, _stream(dev), None, None,
                                                 None, None), "aa_decoder_backward_packed")
        return (None, None, dV, dvg, None, dh0, dc0, None, None) + tuple(grads)


L2_BYTES = 96 << 20      # what of the 126 MB L2 a [rows, Vc] fp32 tensor can count on next to the step's other operands


def fused_loss_pays(n_rows: int, Vc: int) -> bool:
    """Whether ``forward_loss`` takes the fused operator.  ``AA_FUSED_CE`` = 1 / 0 forces it on / off; by default it is taken
    when the packed fp32 logits would not stay in L2: at config 2 (943 x 10 000 x 4 B = 38 MB) the two-step route's logits
    never really leave the chip and its 13 us loss kernel beats the fused epilogue + merge + fix-up passes by ~10 us per step
    (382.6 vs 371 us, profiles/r02_timeline_n1_fusedce_v2.txt); at config 5 (3 000 x 20 000 x 4 B = 240 MB) they do."""
    mode = os.environ.get("AA_FUSED_CE", "auto")
    if mode in ("0", "1"):
        return mode == "1"
    return int(n_rows) * int(Vc) * 4 > L2_BYTES


def decoder_forward_loss(w: Sequence[torch.Tensor], V, v_g, captions, lengths: Sequence[int], targets=None, h0=None, c0=None,
                         denom: int = 0):
    """-> (mean cross-entropy over the packed positions, alpha, beta, hT, cT), differentiable; bf16 tensor-core path only.
    Equal (to bf16 accuracy) to ``cross_entropy(decoder_forward_packed(...)[0].data, targets)`` without the logits ever reaching
    memory.  ``targets`` defaults to the packed next words ``pack(captions[:, 1:], lengths)`` (``train.py:102``)."""
    _need_cuda(V, v_g, captions, h0, c0)
    V, v_g = _f32c(V), _f32c(v_g)
    B, _, H = V.shape
    captions = captions.to(torch.int64).contiguous()
    h0, c0 = _states2d(h0, B, H), _states2d(c0, B, H)
    T = captions.shape[1]
    row_index, _ = cached_row_index(lengths, T, V.device)
    if targets is None:
        if max(int(x) for x in lengths) > T - 1:
            raise ValueError("default targets are the next words captions[:, 1:]: lengths must not exceed T - 1")
        tidx, _ = cached_row_index(lengths, T - 1, V.device)
        targets = captions[:, 1:].reshape(-1)[tidx]
    targets = targets.to(torch.int64).contiguous()
    if targets.numel() != row_index.numel():
        raise ValueError("targets has %d entries, the packed rows %d" % (targets.numel(), row_index.numel()))
    return _DecoderLossFn.apply(PRECISIONS["bf16"], int(denom), V, v_g, captions, h0, c0, row_index, targets, *w)


class _PackRowsFn(torch.autograd.Function):
    """``pack_padded_sequence(scores, lengths, batch_first=True).data`` (baseline_attention.py:228)."""

    @staticmethod
    def forward(ctx, scores, row_index):
        lib = _lib.load()
        B, T, Vc = scores.shape
        n = row_index.numel()
        out = torch.empty(n, Vc, device=scores.device, dtype=torch.float32)
        with torch.cuda.device(scores.device):
            check(lib.aa_pack_rows(_ptr(scores), Vc, _ptr(row_index), n, _ptr(out), _stream(scores.device)), "aa_pack_rows")
        ctx.save_for_backward(row_index)
        ctx.shape = (B, T, Vc)
        return out

    @staticmethod
    def backward(ctx, d_out):
        lib = _lib.load()
        (row_index,) = ctx.saved_tensors
        B, T, Vc = ctx.shape
        d_out = _f32c(d_out)
        d_scores = torch.empty(B, T, Vc, device=d_out.device, dtype=torch.float32)
        with torch.cuda.device(d_out.device):
            check(lib.aa_unpack_rows(_ptr(d_out), Vc, _ptr(row_index), row_index.numel(), B * T, _ptr(d_scores),
                                     _stream(d_out.device)), "aa_unpack_rows")
        return d_scores, None


def packed_row_index(lengths: Sequence[int], T: int):
    """Host logic of ``pack_padded_sequence``: time-major row order b*T+t and batch_sizes."""
    lengths = [int(x) for x in lengths]
    if any(lengths[i] < lengths[i + 1] for i in range(len(lengths) - 1)):
        raise RuntimeError("`lengths` array must be sorted in decreasing order")   # torch's own error
    if lengths and (lengths[-1] <= 0 or lengths[0] > T):
        raise RuntimeError("lengths must be in [1, T]")
    idx, bs = [], []
    for t in range(lengths[0] if lengths else 0):
        n = sum(1 for L in lengths if L > t)
        bs.append(n)
        idx.extend(b * T + t for b in range(n))
    return idx, bs


def packed_targets(captions, lengths: Sequence[int]):
    """``pack_padded_sequence(captions[:, 1:], lengths, batch_first=True).data`` (train.py:102);
    works on numpy arrays and torch tensors."""
    idx, _ = packed_row_index(lengths, captions.shape[1] - 1)
    return captions[:, 1:].reshape(-1)[idx]


_ROW_INDEX_CACHE: Dict[tuple, tuple] = {}


def cached_row_index(lengths: Sequence[int], T: int, device):
    """(device row index b*T+t of every packed row, host batch_sizes).  The device copy is cached per
    (lengths, T, device): repeated shapes cost no host->device copy, which also keeps the callers
    capturable into a CUDA graph."""
    device = torch.device(device)
    key = (tuple(int(x) for x in lengths), int(T), device.type, device.index)
    hit = _ROW_INDEX_CACHE.get(key)
    if hit is None:
        idx, bs = packed_row_index(lengths, T)
        hit = (torch.tensor(idx, dtype=torch.int64, device=device), torch.tensor(bs, dtype=torch.int64))
        if len(_ROW_INDEX_CACHE) > 256:
            _ROW_INDEX_CACHE.clear()
        _ROW_INDEX_CACHE[key] = hit
    return hit


def pack_scores(scores: torch.Tensor, lengths: Sequence[int]):
    row_index, batch_sizes = cached_row_index(lengths, scores.shape[1], scores.device)
    data = _PackRowsFn.apply(scores, row_index)
    return torch.nn.utils.rnn.PackedSequence(data, batch_sizes)


class _CrossEntropyFn(torch.autograd.Function):
    """Mean CE over rows with the gradient produced in the same pass (train.py:63,208)."""

    @staticmethod
    def forward(ctx, logits, targets):
        lib = _lib.load()
        n, Vc = logits.shape
        loss = torch.empty((), device=logits.device, dtype=torch.float32)
        dlog = torch.empty_like(logits)
        dlog16 = torch.empty(n, Vc, device=logits.device, dtype=torch.bfloat16)
        written = ctypes.c_int(0)
        with torch.cuda.device(logits.device):
            check(lib.aa_cross_entropy_mirror(_ptr(logits), n, Vc, _ptr(targets), n, _ptr(loss), _ptr(dlog), _ptr(dlog16),
                                              ctypes.byref(written), _stream(logits.device)), "aa_cross_entropy_mirror")
        ctx.save_for_backward(dlog)
        ctx.dlog16 = dlog16 if written.value else None
        ctx.consumed = False
        return loss

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, g):
        (dlog,) = ctx.saved_tensors
        if ctx.consumed:      # the saved gradient is scaled in place below: a second pass (retain_graph=True) would apply g twice
            raise RuntimeError("adaptive_b200.cross_entropy: backward through the same loss twice is not supported "
                               "(its gradient buffer is consumed in place); recompute the loss instead")
        ctx.consumed = True
        # upstream gradient: a device scalar, 1 when the loss is the root (train.py:210).  dlog is this function's own buffer
        # and is consumed once, so it is scaled in place, and only when g != 1 (decided on the device: no host sync)
        g = g.to(torch.float32).contiguous()
        with torch.cuda.device(dlog.device):
            check(_lib.load().aa_scale_unless_one(_ptr(dlog), _ptr(g), dlog.numel(), _ptr(ctx.dlog16), _stream(dlog.device)),
                  "aa_scale_unless_one")
        if ctx.dlog16 is not None:      # the vocabulary projection's backward contracts with the bf16 copy: hand it along
            dlog._aa_bf16_mirror = ctx.dlog16
            dlog._aa_bf16_mirror_version = dlog._version      # an in-place edit of dlog after this point (a gradient hook) voids the mirror
        return dlog, None


def copy_multi(dsts: Sequence[torch.Tensor], srcs: Sequence[torch.Tensor]) -> None:
    """dst[i].copy_(src[i]) for up to 8 contiguous same-shaped device tensor pairs per launch (one kernel instead of one
    memcpy node each)."""
    pairs = [(d, s) for d, s in zip(dsts, srcs) if d is not s and d.numel() > 0]
    for d, s in pairs:
        if not (d.is_cuda and s.is_cuda and d.is_contiguous() and s.is_contiguous() and d.dtype == s.dtype and d.shape == s.shape):
            raise ValueError("copy_multi: contiguous CUDA tensors of equal shape and dtype expected")
    lib = _lib.load()
    for i in range(0, len(pairs), 8):
        grp = pairs[i:i + 8]
        n = len(grp)
        src = (ctypes.c_void_p * n)(*[s.data_ptr() for _, s in grp])
        dst = (ctypes.c_void_p * n)(*[d.data_ptr() for d, _ in grp])
        nbytes = (ctypes.c_int64 * n)(*[d.numel() * d.element_size() for d, _ in grp])
        with torch.cuda.device(grp[0][0].device):
            check(lib.aa_copy_multi(n, src, dst, nbytes, _stream(grp[0][0].device)), "aa_copy_multi")


def cross_entropy(logits: torch.Tensor, targets: torch.Tensor) -> torch.Tensor:
    _need_cuda(logits, targets)
    return _CrossEntropyFn.apply(_f32c(logits), targets.to(torch.int64).contiguous())


@torch.no_grad()
def greedy_decode(w: Sequence[torch.Tensor], V, v_g, h0=None, c0=None, max_len: int = 30, return_logits: bool = False,
                  precision: str = "tf32x3", engine: str = "pipeline"):
    """``Encoder2Decoder.sampler`` loop (adaptive_attention.py:186-216) on the device.
    Returns ids [B,L] int64, attention [B,L,k], Beta [B,L,1] (+ logits [L,B,Vc]).
    ``precision``: "tf32x3" (per-step contractions on tensor cores, fp32-accurate) or "fp32" (exact SIMT).
    ``engine``: "pipeline" (per-step launches, any batch), "persistent" (one cooperative launch, V resident on chip; raises if
    the batch / shape does not fit) or "auto" (persistent when it fits and nothing else was asked for)."""
    if engine not in ("pipeline", "persistent", "auto"):
        raise ValueError("engine must be 'pipeline', 'persistent' or 'auto'")
    if engine == "persistent" or (engine == "auto" and not return_logits and precision == "tf32x3" and V.shape[0] > 0
                                  and persistent_decode_supported(w, V, v_g, max_len)):
        if return_logits or precision != "tf32x3":
            raise ValueError("the persistent engine returns no logits and runs the tf32x3 precision only")
        return greedy_decode_persistent(w, V, v_g, h0, c0, max_len)
    if engine == "auto" and not return_logits and precision == "tf32x3" and V.is_cuda:
        # up to two images per SM: two persistent launches (0.67 ms each at cfgA) still beat the per-step pipeline (1.8 ms)
        B = V.shape[0]
        sms = torch.cuda.get_device_properties(V.device).multi_processor_count
        half = (B + 1) // 2
        if sms < B <= 2 * sms and persistent_decode_supported(w, V[:half], v_g[:half], max_len):
            cut = lambda t, lo, hi: None if t is None else (t[:, lo:hi] if (t.dim() == 3 and t.shape[0] == 1 and t.shape[1] == B) else t[lo:hi])
            parts = [greedy_decode_persistent(w, V[lo:hi], v_g[lo:hi], cut(h0, lo, hi), cut(c0, lo, hi), max_len)
                     for lo, hi in ((0, half), (half, B))]
            return tuple(torch.cat([a, b], dim=0) for a, b in zip(*parts))
    lib = _lib.load()
    _need_cuda(V, v_g, h0, c0)
    V, v_g = _f32c(V), _f32c(v_g)
    B, k, H = V.shape
    E = v_g.shape[1]
    Vc, a = w[0].shape[0], w[7].shape[0]
    _check_weights(w, H, E, Vc, a)
    h0, c0 = _states2d(h0, B, H), _states2d(c0, B, H)
    dev = V.device
    d = make_dims(B, max_len, k, H, E, Vc, a, DECODE_PRECISIONS[precision])
    ids = torch.empty(B, max_len, device=dev, dtype=torch.int64)
    att = torch.empty(B, max_len, k, device=dev, dtype=torch.float32)
    bet = torch.empty(B, max_len, 1, device=dev, dtype=torch.float32)
    logits = torch.empty(max_len, B, Vc, device=dev, dtype=torch.float32) if return_logits else None
    nbytes = lib.aa_decode_workspace_bytes(ctypes.byref(d), 0)
    wsb = torch.empty(nbytes, device=dev, dtype=torch.uint8)
    ws = weights_struct(w)
    with torch.cuda.device(dev):
        check(lib.aa_greedy_decode(ctypes.byref(d), ctypes.byref(ws), _ptr(V), _ptr(v_g), _ptr(h0), _ptr(c0), max_len,
                                   _ptr(ids), _ptr(att), _ptr(bet), _ptr(logits), _ptr(wsb), nbytes, _stream(dev)),
              "aa_greedy_decode")
    return (ids, att, bet, logits) if return_logits else (ids, att, bet)


_PERSIST_WS: Dict[tuple, tuple] = {}


def persistent_decode_supported(w: Sequence[torch.Tensor], V, v_g, max_len: int = 30) -> bool:
    """True when ``aa_decode_persistent`` can take this batch: at most one image per SM and V, P and the in-kernel
    contraction stages fit in shared memory (the C ABI's ``aa_decode_persistent_supported``)."""
    if not V.is_cuda:
        return False
    B, k, H = V.shape
    d = make_dims(B, max_len, k, H, v_g.shape[1], w[0].shape[0], w[7].shape[0], _lib.PREC_TF32X3)
    with torch.cuda.device(V.device):
        return bool(_lib.load().aa_decode_persistent_supported(ctypes.byref(d)))


@torch.no_grad()
def greedy_decode_persistent(w: Sequence[torch.Tensor], V, v_g, h0=None, c0=None, max_len: int = 30, return_candidates: bool = False):
    """The sampler loop as ONE cooperative launch with V, P and the cell state resident in shared memory (one image per SM):
    ``aa_decode_persistent``.  Same outputs as ``greedy_decode``.  The weight-derived operands (packed gate weights, bf16
    projection, row norms) live at the head of a cached workspace and are rebuilt only when a weight tensor changed
    (``Tensor._version``) -- the per-call prologue is then P = V W_v^T, the static gate terms and the initial state."""
    lib = _lib.load()
    _need_cuda(V, v_g, h0, c0)
    V, v_g = _f32c(V), _f32c(v_g)
    B, k, H = V.shape
    E = v_g.shape[1]
    Vc, a = w[0].shape[0], w[7].shape[0]
    _check_weights(w, H, E, Vc, a)
    h0, c0 = _states2d(h0, B, H), _states2d(c0, B, H)
    dev = V.device
    d = make_dims(B, max_len, k, H, E, Vc, a, _lib.PREC_TF32X3)
    ids = torch.empty(B, max_len, device=dev, dtype=torch.int64)
    att = torch.empty(B, max_len, k, device=dev, dtype=torch.float32)
    bet = torch.empty(B, max_len, 1, device=dev, dtype=torch.float32)
    cand = torch.empty(B, max_len, device=dev, dtype=torch.int32) if return_candidates else None
    with torch.cuda.device(dev):
        nbytes = lib.aa_decode_persistent_workspace_bytes(ctypes.byref(d))
        key = (dev.index, B, max_len, k, H, E, Vc, a)
        # the packed operands are valid for THESE tensor objects at THESE versions: weak references, not addresses (a reloaded model
        # can land on the addresses of the one it replaced, at version 0 again)
        stamp = tuple(t._version if t is not None else None for t in w)
        hit = _PERSIST_WS.get(key)
        flags = 0
        same = hit is not None and hit[1] == stamp and len(hit[2]) == len(w) and all(
            (r is None and t is None) or (r is not None and t is not None and r() is t) for r, t in zip(hit[2], w))
        if same and hit[0].numel() >= nbytes and not torch.cuda.is_current_stream_capturing():
            wsb, flags = hit[0], 1                      # AA_DECODE_REUSE_PACKED_WEIGHTS
        else:
            wsb = torch.empty(nbytes, device=dev, dtype=torch.uint8)
            if not torch.cuda.is_current_stream_capturing():
                if len(_PERSIST_WS) > 16:
                    _PERSIST_WS.clear()
                _PERSIST_WS[key] = (wsb, stamp, tuple(weakref.ref(t) if t is not None else None for t in w))
        ws = weights_struct(w)
        check(lib.aa_decode_persistent(ctypes.byref(d), ctypes.byref(ws), _ptr(V), _ptr(v_g), _ptr(h0), _ptr(c0), max_len, _ptr(ids),
                                       _ptr(att), _ptr(bet), flags, _ptr(cand), _ptr(wsb), nbytes, _stream(dev)), "aa_decode_persistent")
    return (ids, att, bet, cand) if return_candidates else (ids, att, bet)


@torch.no_grad()
def beam_decode(w: Sequence[torch.Tensor], V, v_g, h0=None, c0=None, beam: int = 3, max_len: int = 20,
                precision: str = "tf32x3"):
    """Beam search (not in the reference; definition in oracle.beam_decode / SURVEY §8c).
    Returns ids [B,L], attention [B,L,k], Beta [B,L,1], score [B] of the best hypothesis."""
    lib = _lib.load()
    _need_cuda(V, v_g, h0, c0)
    V, v_g = _f32c(V), _f32c(v_g)
    B, k, H = V.shape
    E = v_g.shape[1]
    Vc, a = w[0].shape[0], w[7].shape[0]
    _check_weights(w, H, E, Vc, a)
    h0, c0 = _states2d(h0, B, H), _states2d(c0, B, H)
    dev = V.device
    d = make_dims(B, max_len, k, H, E, Vc, a, DECODE_PRECISIONS[precision])
    ids = torch.empty(B, max_len, device=dev, dtype=torch.int64)
    att = torch.empty(B, max_len, k, device=dev, dtype=torch.float32)
    bet = torch.empty(B, max_len, 1, device=dev, dtype=torch.float32)
    score = torch.empty(B, device=dev, dtype=torch.float32)
    nbytes = lib.aa_decode_workspace_bytes(ctypes.byref(d), beam)
    wsb = torch.empty(nbytes, device=dev, dtype=torch.uint8)
    ws = weights_struct(w)
    with torch.cuda.device(dev):
        check(lib.aa_beam_decode(ctypes.byref(d), ctypes.byref(ws), _ptr(V), _ptr(v_g), _ptr(h0), _ptr(c0), beam, max_len,
                                 _ptr(ids), _ptr(att), _ptr(bet), _ptr(score), _ptr(wsb), nbytes, _stream(dev)),
              "aa_beam_decode")
    return ids, att, bet, score


# ---- sub-block operators (forward only; training goes through decoder_forward) -----------
@torch.no_grad()
def sentinel_forward(sen_wx, sen_wh, x_t, h_t_1, cell_t):
    """``Sentinel.forward`` (adaptive_attention.py:75-85). x_t [B,T,2E]; h_t_1, cell_t [B,T,H]."""
    lib = _lib.load()
    _need_cuda(x_t, h_t_1, cell_t, sen_wx, sen_wh)
    x_t, h_t_1, cell_t = _f32c(x_t), _f32c(h_t_1), _f32c(cell_t)
    B, T, H = cell_t.shape
    E = x_t.shape[2] // 2
    d = make_dims(B, T, 1, H, E, 1)
    gate = torch.empty_like(cell_t)
    s = torch.empty_like(cell_t)
    with torch.cuda.device(x_t.device):
        check(lib.aa_sentinel_forward(ctypes.byref(d), _ptr(sen_wx), _ptr(sen_wh), _ptr(x_t), _ptr(h_t_1), _ptr(cell_t),
                                      _ptr(gate), _ptr(s), _stream(x_t.device)), "aa_sentinel_forward")
    return s


@torch.no_grad()
def atten_forward(att_wv, att_wg, att_ws, att_wh, V, h_t, s_t):
    """``Atten.forward`` (adaptive_attention.py:26-58) -> c_hat [B,T,H], alpha [B,T,k], beta [B,T,1]."""
    lib = _lib.load()
    _need_cuda(V, h_t, s_t, att_wv)
    V, h_t, s_t = _f32c(V), _f32c(h_t), _f32c(s_t)
    B, k, H = V.shape
    T = h_t.shape[1]
    a = att_wv.shape[0]
    d = make_dims(B, T, k, H, 4, 1, a)
    dev = V.device
    c_hat = torch.empty(B, T, H, device=dev, dtype=torch.float32)
    alpha = torch.empty(B, T, k, device=dev, dtype=torch.float32)
    beta = torch.empty(B, T, 1, device=dev, dtype=torch.float32)
    nbytes = lib.aa_atten_workspace_bytes(ctypes.byref(d))
    wsb = torch.empty(nbytes, device=dev, dtype=torch.uint8)
    with torch.cuda.device(dev):
        check(lib.aa_atten_forward(ctypes.byref(d), _ptr(att_wv), _ptr(att_wg), _ptr(att_ws), _ptr(att_wh), _ptr(V),
                                   _ptr(h_t), _ptr(s_t), _ptr(c_hat), _ptr(alpha), _ptr(beta), _ptr(wsb), nbytes,
                                   _stream(dev)), "aa_atten_forward")
    return c_hat, alpha, beta


@torch.no_grad()
def adaptive_forward(w: Sequence[torch.Tensor], x, hiddens, cells, V):
    """``AdaptiveBlock.forward`` (adaptive_attention.py:110-134) -> scores, alpha, beta [B,T,1]."""
    lib = _lib.load()
    _need_cuda(x, hiddens, cells, V)
    x, hiddens, cells, V = _f32c(x), _f32c(hiddens), _f32c(cells), _f32c(V)
    B, k, H = V.shape
    T = hiddens.shape[1]
    E = x.shape[2] // 2
    Vc, a = w[11].shape[0], w[7].shape[0]
    d = make_dims(B, T, k, H, E, Vc, a)
    dev = V.device
    scores = torch.empty(B, T, Vc, device=dev, dtype=torch.float32)
    alpha = torch.empty(B, T, k, device=dev, dtype=torch.float32)
    beta = torch.empty(B, T, 1, device=dev, dtype=torch.float32)
    nbytes = lib.aa_adaptive_workspace_bytes(ctypes.byref(d))
    wsb = torch.empty(nbytes, device=dev, dtype=torch.uint8)
    ws = weights_struct(w)
    with torch.cuda.device(dev):
        check(lib.aa_adaptive_forward(ctypes.byref(d), ctypes.byref(ws), _ptr(x), _ptr(hiddens), _ptr(cells), _ptr(V),
                                      _ptr(scores), _ptr(alpha), _ptr(beta), _ptr(wsb), nbytes, _stream(dev)),
              "aa_adaptive_forward")
    return scores, alpha, beta


# ---- encoder heads -------------------------------------------------------------------------
class _EncoderFn(torch.autograd.Function):
    """``AttentiveCNN.forward`` after the trunk (baseline_attention.py:46-62) + hand-written backward."""

    @staticmethod
    def forward(ctx, prec, A, *w):
        lib = _lib.load()
        B, C, hw = A.shape
        H, E = w[0].shape[0], w[2].shape[0]
        shapes = [(H, C), (H,), (E, C), (E,), (H, C), (H,), (H, C), (H,)]
        for name, t, s in zip(_lib.ENC_FIELDS, w, shapes):
            if tuple(t.shape) != s or t.dtype != torch.float32 or not t.is_contiguous():
                raise ValueError("encoder weight %s must be contiguous float32 of shape %s (got %s)" % (name, s, tuple(t.shape)))
            _need_cuda(t)
        dev = A.device
        d = _lib.AAEncDims(B=B, C=C, hw=hw, H=H, E=E, precision=prec)
        V = torch.empty(B, hw, H, device=dev, dtype=torch.float32)
        v_g = torch.empty(B, E, device=dev, dtype=torch.float32)
        h0 = torch.empty(B, H, device=dev, dtype=torch.float32)
        c0 = torch.empty(B, H, device=dev, dtype=torch.float32)
        nbytes = lib.aa_encoder_saved_bytes(ctypes.byref(d))
        saved = torch.empty(nbytes, device=dev, dtype=torch.uint8)
        ws = _lib.AAEncWeights()
        for name, t in zip(_lib.ENC_FIELDS, w):
            setattr(ws, name, t.data_ptr())
        with torch.cuda.device(dev):
            check(lib.aa_encoder_forward(ctypes.byref(d), ctypes.byref(ws), _ptr(A), _ptr(V), _ptr(v_g), _ptr(h0), _ptr(c0),
                                         _ptr(saved), nbytes, _stream(dev)), "aa_encoder_forward")
        ctx.dims = (B, C, hw, H, E, prec)
        ctx.save_for_backward(V, v_g, h0, c0, saved, *w)
        return V, v_g, h0, c0

    @staticmethod
    def backward(ctx, dV, dv_g, dh0, dc0):
        lib = _lib.load()
        V, v_g, h0, c0, saved = ctx.saved_tensors[:5]
        w = ctx.saved_tensors[5:]
        B, C, hw, H, E, prec = ctx.dims
        dev = V.device
        d = _lib.AAEncDims(B=B, C=C, hw=hw, H=H, E=E, precision=prec)
        dV, dv_g, dh0, dc0 = (_f32c(g) for g in (dV, dv_g, dh0, dc0))
        want_dA = bool(ctx.needs_input_grad[1])
        grads = [torch.empty_like(t) for t in w]
        gs = _lib.AAEncWeightGrads()
        ws = _lib.AAEncWeights()
        for name, t, g in zip(_lib.ENC_FIELDS, w, grads):
            setattr(ws, name, t.data_ptr())
            setattr(gs, name, g.data_ptr())
        dA = torch.empty(B, C, hw, device=dev, dtype=torch.float32) if want_dA else None
        nbytes = lib.aa_encoder_bwd_scratch_bytes(ctypes.byref(d), 1 if want_dA else 0)
        scratch = torch.empty(nbytes, device=dev, dtype=torch.uint8)
        with torch.cuda.device(dev):
            check(lib.aa_encoder_backward(ctypes.byref(d), ctypes.byref(ws), _ptr(saved), saved.numel(), _ptr(V), _ptr(v_g), _ptr(h0),
                                          _ptr(c0), _ptr(dV), _ptr(dv_g), _ptr(dh0), _ptr(dc0), ctypes.byref(gs), _ptr(dA),
                                          _ptr(scratch), nbytes, _stream(dev)), "aa_encoder_backward")
        return (None, dA, *grads)


def encoder_forward(w: Sequence[torch.Tensor], A: torch.Tensor, precision: str = "fp32"):
    """Encoder heads: ``w`` = (affine_a.weight, .bias, affine_b.weight, .bias, affine_h0.weight, .bias, affine_c0.weight, .bias),
    ``A`` the last-conv feature map ``[B,C,h,w]`` or ``[B,C,hw]`` -> ``V [B,hw,H], v_g [B,E], h0 [B,H], c0 [B,H]``."""
    _need_cuda(A)
    if A.dim() == 4:
        A = A.reshape(A.shape[0], A.shape[1], -1)
    return _EncoderFn.apply(PRECISIONS[precision], _f32c(A), *w)
