"""Shared helpers for the test-suite (golden loading, tolerances)."""
import os

import numpy as np

from adaptive_b200.synth import Dims, make_inputs, make_weights

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
GOLDEN_CASES = ("tiny", "k196", "odd", "cfgA")


def load_golden(name):
    g = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    H, E, Vc, k, B, T, L = [int(x) for x in g["meta"]]
    return g, Dims(H=H, E=E, Vc=Vc, k=k), B, T, L


def golden_setup(name, dtype=np.float32):
    """Weights and inputs exactly as oracle/gen_golden.py built them."""
    g, dims, B, T, L = load_golden(name)
    w = make_weights(dims, seed=123, dtype=np.float32, bias_scale=0.1)
    inp = make_inputs(dims, B, T, seed=1234, dtype=np.float32)
    w = {k: v.astype(dtype) for k, v in w.items()}
    inp = {k: (v.astype(dtype) if v.dtype != np.int64 else v) for k, v in inp.items()}
    return g, dims, B, T, L, w, inp


def upstream(shape_scores, shape_alpha, shape_beta, shape_h, dtype):
    """The fixed random upstream gradients used by oracle/gen_golden.py."""
    rng = np.random.Generator(np.random.PCG64(99))
    dS = rng.standard_normal(shape_scores).astype(dtype) / shape_scores[-1]
    dA = rng.standard_normal(shape_alpha).astype(dtype) * 0.1
    dB = rng.standard_normal(shape_beta).astype(dtype) * 0.1
    dH = rng.standard_normal((1,) + tuple(shape_h)).astype(dtype)[0] * 0.1
    dC = rng.standard_normal((1,) + tuple(shape_h)).astype(dtype)[0] * 0.1
    return dS, dA, dB, dH, dC


def rel_err(a, b):
    """max |a-b| / max(|b|) — the 'relative' of BASELINE.json's 1e-4 / 2e-2 tolerances."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


# ---- encoder heads / baseline decoder fixtures (oracle/gen_golden.py: run_encoder_case, run_baseline_case) ----
ENC_CASES = ("enc_tiny", "enc_odd", "enc_cfgA")
BASE_CASES = ("base_tiny", "base_odd")


def encoder_setup(name, dtype=np.float32):
    from adaptive_b200.synth import make_encoder_weights, make_features

    g = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    H, E, C, B, fh, fw = [int(x) for x in g["meta"]]
    dims = Dims(H=H, E=E, Vc=8, k=fh * fw)
    w = {k: v.astype(dtype) for k, v in make_encoder_weights(dims, C, seed=321, bias_scale=0.1).items()}
    A = make_features(B, C, (fh, fw), seed=4321).astype(dtype)
    rng = np.random.Generator(np.random.PCG64(77))
    ups = [rng.standard_normal(s).astype(dtype) for s in ((B, fh * fw, H), (B, E), (B, 1, H), (B, 1, H))]
    ups[2], ups[3] = ups[2][:, 0], ups[3][:, 0]
    return g, dims, C, B, w, A, ups


def baseline_setup(name, dtype=np.float32):
    from adaptive_b200.synth import baseline_weights

    g, dims, B, T, L, w, inp = golden_setup(name, dtype)
    return g, dims, B, T, L, baseline_weights(w), inp
