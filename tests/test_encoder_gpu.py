"""GPU parity tests of the encoder heads (affine_a / affine_b / affine_h0 / affine_c0 + average pool,
baseline_attention.py:21-34, 46-62): ``aa_encoder_forward`` / ``aa_encoder_backward`` through the C ABI against the golden
vectors of the reference's AttentiveCNN (ResNet trunk replaced by Identity) and against the numpy oracle.

Tolerances: fp32 path 1e-4 relative (max-abs error over max-abs value), bf16 path 2e-2 (BASELINE.json north_star)."""
import numpy as np
import pytest
import torch

from adaptive_b200 import functional as F_aa
from adaptive_b200 import modules
from adaptive_b200._lib import ENC_FIELDS, ENC_KEY_TO_FIELD
from adaptive_b200.synth import Dims, make_encoder_weights, make_features
from oracle import adaptive_oracle as orc
from tests.helpers import ENC_CASES, encoder_setup, rel_err

pytestmark = pytest.mark.gpu
TOL = 1e-4
TOL_BF16 = 2e-2


def _dev_enc_weights(w, requires_grad=True):
    inv = {v: k for k, v in ENC_KEY_TO_FIELD.items()}
    out = []
    for f in ENC_FIELDS:
        t = torch.from_numpy(np.ascontiguousarray(w[inv[f]], dtype=np.float32)).cuda()
        out.append(t.requires_grad_(requires_grad))
    return tuple(out), [inv[f] for f in ENC_FIELDS]


def _run(w, A, ups, precision):
    W, keys = _dev_enc_weights(w)
    At = torch.from_numpy(A.astype(np.float32)).cuda().requires_grad_(True)
    V, v_g, h0, c0 = F_aa.encoder_forward(W, At, precision)
    loss = sum((t * torch.from_numpy(u.astype(np.float32)).cuda()).sum() for t, u in zip((V, v_g, h0, c0), ups))
    loss.backward()
    outs = {k: t.detach().cpu().numpy() for k, t in (("V", V), ("v_g", v_g), ("h0", h0), ("c0", c0))}
    grads = {k: t.grad.cpu().numpy() for k, t in zip(keys, W)}
    grads["A"] = At.grad.cpu().numpy()
    return outs, grads


@pytest.mark.parametrize("case", ENC_CASES)
def test_encoder_vs_golden(case):
    """fp32 path against the reference run in float64 (the arbiter)."""
    g, dims, C, B, w, A, ups = encoder_setup(case, np.float32)
    outs, grads = _run(w, A, ups, "fp32")
    tag = "f64"
    for key in ("V", "v_g", "h0", "c0"):
        assert rel_err(outs[key], g[tag + "_" + key]) < TOL, key
    for key in w:
        got = grads[key]
        if (tag + "_grad_" + key) in g.files:
            assert rel_err(got, g[tag + "_grad_" + key]) < TOL, key
        else:
            assert rel_err(got.reshape(-1)[::251], g[tag + "_grad_sub_" + key]) < TOL, key
            nrm = np.sqrt((got.astype(np.float64) ** 2).sum())
            assert abs(nrm - float(g[tag + "_grad_norm_" + key])) < TOL * nrm, key
    if (tag + "_grad_A") in g.files:
        assert rel_err(grads["A"].reshape(g[tag + "_grad_A"].shape), g[tag + "_grad_A"]) < TOL
    else:
        assert rel_err(grads["A"].reshape(-1)[::251], g[tag + "_grad_A_sub"]) < TOL


@pytest.mark.parametrize("precision,tol", [("fp32", TOL), ("bf16", TOL_BF16)])
@pytest.mark.parametrize("B,C,hw,dims", [(16, 2048, (7, 7), Dims(H=512, E=256, Vc=8, k=49)),       # BASELINE config 2 shapes
                                         (3, 256, (14, 14), Dims(H=128, E=64, Vc=8, k=196)),      # 14x14 maps (config 5)
                                         (5, 72, (3, 5), Dims(H=40, E=24, Vc=8, k=15))])          # ragged tails
def test_encoder_vs_oracle(precision, tol, B, C, hw, dims):
    w = make_encoder_weights(dims, C, seed=11, bias_scale=0.1)
    A = make_features(B, C, hw, seed=12)
    rng = np.random.Generator(np.random.PCG64(13))
    ups = [rng.standard_normal(s).astype(np.float32) for s in ((B, hw[0] * hw[1], dims.H), (B, dims.E), (B, dims.H), (B, dims.H))]
    w64 = {k: v.astype(np.float64) for k, v in w.items()}
    V, v_g, h0, c0, cache = orc.encoder_forward(w64, A.astype(np.float64), want_cache=True)
    outs, grads = _run(w, A, ups, precision)
    for key, ref in (("V", V), ("v_g", v_g), ("h0", h0), ("c0", c0)):
        assert rel_err(outs[key], ref) < tol, key
    if precision == "bf16":
        # ReLU is discontinuous in its gradient: a pre-activation within bf16 rounding of 0 may land on the other side, and each
        # such element moves a weight-gradient row by a whole upstream value.  The backward is therefore checked for the ReLU
        # masks the device's own forward produced (a handful of elements out of B*hw*H differ from the fp64 masks).
        flips = int(((outs["V"] > 0) != (V > 0)).sum())
        assert flips <= 0.02 * V.size, flips
        cache["V"], cache["v_g"] = outs["V"].astype(np.float64), outs["v_g"].astype(np.float64)
    G = orc.encoder_backward(w64, cache, *[u.astype(np.float64) for u in ups])
    for key in list(w) + ["A"]:
        assert rel_err(grads[key].reshape(G[key].shape), G[key]) < tol, key


def test_encoder_without_feature_gradient_and_module_surface():
    """A without requires_grad (the frozen trunk before fine-tuning): no dA is computed; the module returns the reference's
    ([B,hw,H], [B,E], ([B,1,H], [B,1,H])) and its parameters receive the same gradients as the functional call."""
    dims, B, C = Dims(H=64, E=32, Vc=8, k=49), 4, 128
    w = make_encoder_weights(dims, C, seed=3, bias_scale=0.1)
    A = make_features(B, C, (7, 7), seed=4)
    enc = modules.AttentiveCNN(dims.E, dims.H, None, feat_dim=C).cuda()
    enc.load_state_dict({k: torch.from_numpy(v) for k, v in w.items()}, strict=True)
    At = torch.from_numpy(A).cuda()
    V, v_g, (h0, c0) = enc(At)
    assert V.shape == (B, 49, dims.H) and v_g.shape == (B, dims.E) and h0.shape == (B, 1, dims.H) and c0.shape == (B, 1, dims.H)
    (V.sum() + v_g.sum() + h0.sum() + c0.sum()).backward()
    assert At.grad is None
    V_o, vg_o, h0_o, c0_o, cache = orc.encoder_forward({k: v.astype(np.float64) for k, v in w.items()}, A.astype(np.float64), True)
    G = orc.encoder_backward({k: v.astype(np.float64) for k, v in w.items()}, cache, np.ones_like(V_o), np.ones_like(vg_o),
                             np.ones_like(h0_o), np.ones_like(c0_o))
    assert rel_err(V.detach().cpu().numpy(), V_o) < TOL and rel_err(h0.detach().cpu().numpy()[:, 0], h0_o) < TOL
    for key, p in enc.named_parameters():
        assert rel_err(p.grad.cpu().numpy(), G[key]) < TOL, key


def test_encoder_errors_are_loud():
    dims, C = Dims(H=64, E=32, Vc=8, k=49), 128
    W, _ = _dev_enc_weights(make_encoder_weights(dims, C, seed=3), requires_grad=False)
    with pytest.raises(RuntimeError):       # CPU tensor: no fallback
        F_aa.encoder_forward(W, torch.zeros(2, C, 7, 7))
    with pytest.raises(ValueError):         # channel count does not match the weights
        F_aa.encoder_forward(W, torch.zeros(2, C + 4, 7, 7, device="cuda"))
    bad = tuple(t[:, :-1].contiguous() if t.dim() == 2 else t for t in W)   # C = 127: not a multiple of 4
    with pytest.raises(RuntimeError):
        F_aa.encoder_forward(bad, torch.zeros(2, C - 1, 7, 7, device="cuda"))
