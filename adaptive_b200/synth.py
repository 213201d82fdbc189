"""Synthetic weights and inputs for the caption-decoder hot path (numpy only).

The CNN encoder is out of scope (BASELINE.json north_star), so every test and bench feeds
synthetic region features.  Value distributions follow what the reference would produce:

* weights: the init helpers of ``code_src/models/model_utils.py:4-74`` as used at
  ``adaptive_attention.py:23-24,73,108`` and ``baseline_attention.py:137,146``
  (xavier-uniform / kaiming-normal / orthogonal LSTM with forget bias 0.5+0.5);
* inputs: ``V, v_g`` post-ReLU (``baseline_attention.py:51-53``), ``h0, c0`` post-tanh
  (``baseline_attention.py:56-58``), captions with ``<start>``=1 first and words drawn from
  ``[4, vocab)`` (``code_src/data/build_vocab.py:48-51``).

Everything is drawn from ``numpy.random.Generator(PCG64(seed))`` so the same tensors can be
rebuilt bit-for-bit on any box without shipping them.
"""
from __future__ import annotations

import dataclasses
import math
from typing import Dict, List, Tuple

import numpy as np

ATT_DIM = 49  # hard-coded attention inner dim of the reference (adaptive_attention.py:16-19)

# state_dict keys of the reference decoder, in registration order
DECODER_KEYS: Tuple[str, ...] = (
    "embed.weight",
    "LSTM.weight_ih_l0",
    "LSTM.weight_hh_l0",
    "LSTM.bias_ih_l0",
    "LSTM.bias_hh_l0",
    "adaptive.sentinel.affine_x.weight",
    "adaptive.sentinel.affine_h.weight",
    "adaptive.atten.affine_v.weight",
    "adaptive.atten.affine_g.weight",
    "adaptive.atten.affine_s.weight",
    "adaptive.atten.affine_h.weight",
    "adaptive.mlp.weight",
    "adaptive.mlp.bias",
)


@dataclasses.dataclass(frozen=True)
class Dims:
    """Shape of one decoder instance: H hidden, E word-embedding, Vc vocab, k regions."""

    H: int = 512
    E: int = 256
    Vc: int = 10000
    k: int = 49
    a: int = ATT_DIM

    def shapes(self) -> Dict[str, Tuple[int, ...]]:
        H, E, Vc, a = self.H, self.E, self.Vc, self.a
        return {
            "embed.weight": (Vc, E),
            "LSTM.weight_ih_l0": (4 * H, 2 * E),
            "LSTM.weight_hh_l0": (4 * H, H),
            "LSTM.bias_ih_l0": (4 * H,),
            "LSTM.bias_hh_l0": (4 * H,),
            "adaptive.sentinel.affine_x.weight": (H, 2 * E),
            "adaptive.sentinel.affine_h.weight": (H, H),
            "adaptive.atten.affine_v.weight": (a, H),
            "adaptive.atten.affine_g.weight": (a, H),
            "adaptive.atten.affine_s.weight": (a, H),
            "adaptive.atten.affine_h.weight": (1, a),
            "adaptive.mlp.weight": (Vc, H),
            "adaptive.mlp.bias": (Vc,),
        }


CFG_A = Dims(H=512, E=256, Vc=10000, k=49)      # BASELINE configs 1-4
CFG_B = Dims(H=1024, E=512, Vc=20000, k=196)    # BASELINE config 5 (E = H/2 as in cfg_wzn.py:115-116)


def _xavier_uniform(rng, shape, gain):
    fan_out, fan_in = shape
    b = gain * math.sqrt(6.0 / (fan_in + fan_out))
    return rng.uniform(-b, b, size=shape)


def _kaiming_normal(rng, shape):
    fan_in = shape[1]
    return rng.standard_normal(shape) * math.sqrt(2.0 / fan_in)


def _orthogonal(rng, shape):
    rows, cols = shape
    m = rng.standard_normal((max(rows, cols), min(rows, cols)))
    q, r = np.linalg.qr(m)
    q = q * np.sign(np.diag(r))
    return q if rows >= cols else q.T


def make_weights(dims: Dims, seed: int = 123, dtype=np.float32, bias_scale: float = 0.0) -> Dict[str, np.ndarray]:
    """Decoder weights keyed like the reference ``Decoder.state_dict()``.

    ``bias_scale`` > 0 adds noise to the biases (the reference initialises them to 0 /
    0.5; tests use a non-zero scale so that bias handling is actually exercised).
    """
    rng = np.random.Generator(np.random.PCG64(seed))
    H, E = dims.H, dims.E
    sh = dims.shapes()
    w: Dict[str, np.ndarray] = {}
    w["embed.weight"] = rng.standard_normal(sh["embed.weight"])
    w["LSTM.weight_ih_l0"] = _orthogonal(rng, sh["LSTM.weight_ih_l0"])
    w["LSTM.weight_hh_l0"] = _orthogonal(rng, sh["LSTM.weight_hh_l0"])
    for name in ("LSTM.bias_ih_l0", "LSTM.bias_hh_l0"):
        b = np.zeros(4 * H)
        b[H:2 * H] = 0.5
        if bias_scale:
            b = b + bias_scale * rng.standard_normal(4 * H)
        w[name] = b
    w["adaptive.sentinel.affine_x.weight"] = _xavier_uniform(rng, sh["adaptive.sentinel.affine_x.weight"], 1.0)
    w["adaptive.sentinel.affine_h.weight"] = _xavier_uniform(rng, sh["adaptive.sentinel.affine_h.weight"], 1.0)
    for n in ("v", "g", "s"):
        key = "adaptive.atten.affine_%s.weight" % n
        w[key] = _xavier_uniform(rng, sh[key], 5.0 / 3.0)
    w["adaptive.atten.affine_h.weight"] = _kaiming_normal(rng, sh["adaptive.atten.affine_h.weight"])
    w["adaptive.mlp.weight"] = _kaiming_normal(rng, sh["adaptive.mlp.weight"])
    b = np.zeros(dims.Vc)
    if bias_scale:
        b = bias_scale * rng.standard_normal(dims.Vc)
    w["adaptive.mlp.bias"] = b
    return {k: np.ascontiguousarray(v, dtype=dtype) for k, v in w.items()}


def make_lengths(B: int, T: int, seed: int = 1234, full: bool = False) -> List[int]:
    """Caption lengths already reduced by one (what ``train.py:101`` hands to forward),
    sorted descending as ``collate_fn`` does (``data_loader.py:64-98``).  ``full`` gives
    ``[T-1]*B``; otherwise lengths follow the COCO statistic (mean 10.47 words,
    ``statics:12``) clipped to the batch maximum."""
    if full:
        return [T - 1] * B
    rng = np.random.Generator(np.random.PCG64(seed + 7))
    raw = np.clip(np.rint(rng.normal(10.5, 2.5, size=B)) + 2, 7, T).astype(np.int64)
    raw[0] = T  # the longest caption defines T
    ls = np.sort(raw)[::-1] - 1
    return [int(x) for x in ls]


def make_inputs(dims: Dims, B: int, T: int, seed: int = 1234, dtype=np.float32) -> Dict[str, np.ndarray]:
    """``V [B,k,H]``, ``v_g [B,E]``, ``h0/c0 [B,H]``, ``captions [B,T]`` int64."""
    rng = np.random.Generator(np.random.PCG64(seed))
    V = np.maximum(rng.standard_normal((B, dims.k, dims.H)), 0.0)
    v_g = np.maximum(rng.standard_normal((B, dims.E)), 0.0)
    h0 = np.tanh(rng.standard_normal((B, dims.H)))
    c0 = np.tanh(rng.standard_normal((B, dims.H)))
    cap = rng.integers(4, dims.Vc, size=(B, T), dtype=np.int64)
    cap[:, 0] = 1
    return {
        "V": np.ascontiguousarray(V, dtype=dtype),
        "v_g": np.ascontiguousarray(v_g, dtype=dtype),
        "h0": np.ascontiguousarray(h0, dtype=dtype),
        "c0": np.ascontiguousarray(c0, dtype=dtype),
        "captions": cap,
    }


# ---- sentinel-less baseline decoder (baseline_attention.py:66-194) --------------------------
BASELINE_KEYS: Tuple[str, ...] = tuple(k for k in DECODER_KEYS if "sentinel" not in k and "affine_s" not in k)


def baseline_weights(w: Dict[str, np.ndarray]) -> Dict[str, np.ndarray]:
    """The subset of ``make_weights`` the baseline ``Decoder.state_dict()`` holds (no sentinel, no ``affine_s``)."""
    return {k: w[k] for k in BASELINE_KEYS}


# ---- encoder heads (baseline_attention.py:21-34) --------------------------------------------
ENCODER_KEYS: Tuple[str, ...] = ("affine_a.weight", "affine_a.bias", "affine_b.weight", "affine_b.bias",
                                 "affine_h0.weight", "affine_h0.bias", "affine_c0.weight", "affine_c0.bias")


def make_encoder_weights(dims: Dims, C: int = 2048, seed: int = 321, dtype=np.float32, bias_scale: float = 0.0):
    """Head weights keyed like ``AttentiveCNN.state_dict()`` minus the trunk: kaiming-uniform(relu) for ``affine_a/b``
    (``baseline_attention.py:29``), xavier-uniform(tanh) for ``affine_h0/c0`` (``:34``), zero biases (+ noise for tests)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    w: Dict[str, np.ndarray] = {}
    for name, rows in (("affine_a", dims.H), ("affine_b", dims.E)):
        b = math.sqrt(2.0) * math.sqrt(3.0 / C)
        w[name + ".weight"] = rng.uniform(-b, b, size=(rows, C))
    for name in ("affine_h0", "affine_c0"):
        w[name + ".weight"] = _xavier_uniform(rng, (dims.H, C), 5.0 / 3.0)
    for name, rows in (("affine_a", dims.H), ("affine_b", dims.E), ("affine_h0", dims.H), ("affine_c0", dims.H)):
        w[name + ".bias"] = bias_scale * rng.standard_normal(rows) if bias_scale else np.zeros(rows)
    return {k: np.ascontiguousarray(w[k], dtype=dtype) for k in ENCODER_KEYS}


def make_features(B: int, C: int = 2048, hw: Tuple[int, int] = (7, 7), seed: int = 4321, dtype=np.float32) -> np.ndarray:
    """Synthetic last-conv feature maps ``[B,C,h,w]``: post-ReLU like the trunk's output, scaled so that the pooled
    pre-activations stay O(1)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    return np.ascontiguousarray(np.maximum(rng.standard_normal((B, C) + tuple(hw)), 0.0), dtype=dtype)
