"""GPU parity tests of the sentinel-less baseline decoder (``baseline_attention.py:66-283``): the adaptive
operators with ``sen_wx = sen_wh = att_ws = NULL`` against the golden vectors of the reference's ``baseline_attention.Decoder``
(``tests/golden/base_*.npz``) and against the oracle's baseline mode.

Tolerances: fp32 1e-4 relative, bf16 2e-2 (BASELINE.json north_star); greedy ids exact away from near-ties."""
import numpy as np
import pytest
import torch

from adaptive_b200 import baseline
from adaptive_b200 import functional as F_aa
from adaptive_b200.synth import CFG_A, Dims, baseline_weights, make_inputs, make_lengths, make_weights
from oracle import adaptive_oracle as orc
from tests.gpu_utils import dev_inputs, dev_weights, grad_key_order, near_tie_report
from tests.helpers import BASE_CASES, baseline_setup, rel_err, upstream

pytestmark = pytest.mark.gpu
TOL = 1e-4
TOL_BF16 = 2e-2


def _dev_base_weights(w_full, requires_grad=True):
    """13-tuple in aa_weights order, sentinel entries None."""
    return F_aa.baseline_weights(dev_weights(w_full, requires_grad=requires_grad))


def _full(w_base, dims):
    """re-insert dummy sentinel entries so that gpu_utils.dev_weights can order the dict (they are dropped again)."""
    full = make_weights(dims, seed=123, bias_scale=0.1)
    full.update(w_base)
    return full


def _oracle_greedy(w64, i64, L):
    """fp64 oracle greedy ids / attention and the top-1 - top-2 logit gap per position (near-tie arbiter)."""
    ids, att, _, sc = orc.greedy_decode(w64, i64["V"], i64["v_g"], i64["h0"], i64["c0"], L, want_scores=True)
    top2 = np.sort(sc, axis=-1)[..., -2:]
    return ids, att, top2[..., 1] - top2[..., 0]


@pytest.mark.parametrize("case", BASE_CASES)
def test_baseline_forward_backward_greedy_vs_golden(case):
    g, dims, B, T, L, w, inp = baseline_setup(case, np.float32)
    W = _dev_base_weights(_full(w, dims))
    V, v_g, h0, c0, cap = dev_inputs(inp, requires_grad=True)
    scores, alpha, beta, hT, cT = F_aa.decoder_forward(W, V, v_g, cap, h0, c0)
    tag = "f64"
    assert rel_err(scores.detach().cpu().numpy(), g[tag + "_scores"]) < TOL
    assert rel_err(alpha.detach().cpu().numpy(), g[tag + "_alpha"]) < TOL
    assert rel_err(hT.detach().cpu().numpy(), g[tag + "_hT"]) < TOL
    assert rel_err(cT.detach().cpu().numpy(), g[tag + "_cT"]) < TOL
    assert float(beta.abs().max()) == 0.0
    dS, dA, _, _, _ = upstream(tuple(scores.shape), tuple(alpha.shape), tuple(beta.shape), tuple(hT.shape), np.float32)
    ((scores * torch.from_numpy(dS).cuda()).sum() + (alpha * torch.from_numpy(dA).cuda()).sum()).backward()
    for key, t in zip(grad_key_order(), W):
        if t is None:
            continue
        assert rel_err(t.grad.cpu().numpy(), g[tag + "_grad_" + key]) < TOL, key
    for key, t in (("V", V), ("v_g", v_g), ("h0", h0), ("c0", c0)):
        assert rel_err(t.grad.cpu().numpy(), g[tag + "_grad_" + key]) < TOL, key
    # greedy sampler, both decode engines
    Wd = tuple(None if t is None else t.detach() for t in W)
    w64 = {k: v.astype(np.float64) for k, v in w.items()}
    i64 = {k: (v.astype(np.float64) if v.dtype != np.int64 else v) for k, v in inp.items()}
    ids_o, att_o, gap = _oracle_greedy(w64, i64, L)
    assert np.array_equal(ids_o, g[tag + "_greedy_ids"])          # (the oracle itself is pinned by test_oracle_golden.py)
    for prec in ("fp32", "tf32x3"):
        ids, att, bet = F_aa.greedy_decode(Wd, V.detach(), v_g.detach(), h0.detach(), c0.detach(), L, precision=prec)
        hard, near = near_tie_report(ids.cpu().numpy(), g[tag + "_greedy_ids"], gap, 1e-4)
        assert not hard, (prec, hard)
        same = (ids.cpu().numpy() == g[tag + "_greedy_ids"]).all(1)
        assert same.any() and rel_err(att.cpu().numpy()[same], g[tag + "_greedy_alpha"][same]) < TOL, prec
        assert float(bet.abs().max()) == 0.0


@pytest.mark.parametrize("precision,tol", [("fp32", TOL), ("bf16", TOL_BF16)])
@pytest.mark.parametrize("B,T,dims", [(80, 18, CFG_A), (9, 5, Dims(H=128, E=64, Vc=1000, k=196))])
def test_baseline_training_vs_oracle(precision, tol, B, T, dims):
    """BASELINE config 2 shapes (cluster recurrence kernels in bf16) in baseline mode, packed entry point included."""
    w = baseline_weights(make_weights(dims, seed=5, bias_scale=0.1))
    inp = make_inputs(dims, B, T, seed=6)
    w64 = {k: v.astype(np.float64) for k, v in w.items()}
    i64 = {k: (v.astype(np.float64) if v.dtype != np.int64 else v) for k, v in inp.items()}
    s_o, a_o, b_o, (h_o, c_o), cache = orc.decoder_forward(w64, i64["V"], i64["v_g"], i64["captions"], i64["h0"], i64["c0"],
                                                           want_cache=True)
    lengths = make_lengths(B, T, seed=3)
    rows = np.concatenate([np.arange(sum(1 for L in lengths if L > t)) * T + t for t in range(max(lengths))])
    rng = np.random.Generator(np.random.PCG64(1))
    dP = rng.standard_normal((rows.size, dims.Vc)) / dims.Vc
    dS = np.zeros((B * T, dims.Vc))
    dS[rows] = dP
    G = orc.decoder_backward(w64, cache, dS.reshape(B, T, dims.Vc))
    W = _dev_base_weights(_full(w, dims))
    V, v_g, h0, c0, cap = dev_inputs(inp, requires_grad=True)
    packed, alpha, beta, hT, cT = F_aa.decoder_forward_packed(W, V, v_g, cap, lengths, h0, c0, precision)
    assert rel_err(packed.data.detach().cpu().numpy(), s_o.reshape(B * T, -1)[rows]) < tol
    assert rel_err(alpha.detach().cpu().numpy(), a_o) < tol
    assert float(beta.abs().max()) == 0.0
    (packed.data * torch.from_numpy(dP.astype(np.float32)).cuda()).sum().backward()
    for key, t in zip(grad_key_order(), W):
        if t is None:
            continue
        assert rel_err(t.grad.cpu().numpy(), G[key]) < tol, key
    for key, t in (("V", V), ("v_g", v_g), ("h0", h0), ("c0", c0)):
        assert rel_err(t.grad.cpu().numpy(), G[key]) < tol, key


def test_baseline_modules_surface_and_beam():
    """state_dict keys and return tuples of the reference's baseline classes; stage operators; beam search runs."""
    dims, B, T, L = Dims(H=64, E=32, Vc=200, k=49), 6, 5, 6

    class Cf:
        base_word_embed_size, base_lstm_hidden_size, vocab_length = dims.E, dims.H, dims.Vc

    model = baseline.Encoder2Decoder(Cf()).cuda()
    keys = set(k for k in model.decoder.state_dict())
    assert keys == {"embed.weight", "LSTM.weight_ih_l0", "LSTM.weight_hh_l0", "LSTM.bias_ih_l0", "LSTM.bias_hh_l0",
                    "adaptive.atten.affine_v.weight", "adaptive.atten.affine_g.weight", "adaptive.atten.affine_h.weight",
                    "adaptive.mlp.weight", "adaptive.mlp.bias"}
    w = baseline_weights(make_weights(dims, seed=8, bias_scale=0.1))
    model.decoder.load_state_dict({k: torch.from_numpy(v) for k, v in w.items()}, strict=True)
    inp = make_inputs(dims, B, T, seed=9)
    V, v_g, h0, c0, cap = dev_inputs(inp)
    w64 = {k: v.astype(np.float64) for k, v in w.items()}
    i64 = {k: (v.astype(np.float64) if v.dtype != np.int64 else v) for k, v in inp.items()}
    s_o, a_o, _, _, cache = orc.decoder_forward(w64, i64["V"], i64["v_g"], i64["captions"], i64["h0"], i64["c0"], want_cache=True)
    scores, alpha, (hT, cT) = model.decoder(V, v_g, cap, (h0[None], c0[None]))
    assert rel_err(scores.detach().cpu().numpy(), s_o) < TOL and rel_err(alpha.detach().cpu().numpy(), a_o) < TOL
    assert hT.shape == (1, B, dims.H)
    # stage operators: Atten.forward(V, h_t) and AdaptiveBlock.forward(x, hiddens, cells, V)
    hid = torch.from_numpy(cache["hiddens"].astype(np.float32)).cuda()
    cel = torch.from_numpy(cache["cells"].astype(np.float32)).cuda()
    x = torch.from_numpy(cache["x"].astype(np.float32)).cuda()
    c_t, al = model.decoder.adaptive.atten(V, hid)
    assert rel_err(c_t.cpu().numpy(), cache["ctx"]) < TOL and rel_err(al.cpu().numpy(), a_o) < TOL
    sc2, al2 = model.decoder.adaptive(x, hid, cel, V)
    assert rel_err(sc2.cpu().numpy(), s_o) < TOL
    # packed forward and sampler through the shell, features entering through the encoder heads
    lengths = make_lengths(B, T, seed=2)
    packed = model((V, v_g, (h0, c0)), cap, lengths)
    assert rel_err(packed.data.detach().cpu().numpy(), orc.pack_padded(s_o, lengths)[0]) < TOL
    ids, att = model.sampler((V, v_g, (h0, c0)), max_len=L)
    ids_o, att_o, gap = _oracle_greedy(w64, i64, L)
    hard, near = near_tie_report(ids.cpu().numpy(), ids_o, gap, 1e-4)
    assert not hard and att.shape == (B, L, dims.k)
    # beam search over the baseline step (definition = oracle.beam_decode in baseline mode; parity unpinned, Q14)
    bi, ba, bb, bs = F_aa.beam_decode(model.decoder.weights(), V, v_g, h0, c0, beam=3, max_len=L)
    bi_o, ba_o, _, bs_o = orc.beam_decode(w64, i64["V"], i64["v_g"], i64["h0"], i64["c0"], beam=3, max_len=L)
    assert np.array_equal(bi.cpu().numpy(), bi_o) and rel_err(bs.cpu().numpy(), bs_o) < TOL


def test_partial_sentinel_weights_are_rejected():
    dims = Dims(H=64, E=32, Vc=200, k=49)
    W = list(dev_weights(make_weights(dims, seed=1), requires_grad=False))
    W[5] = None      # sen_wx only: neither model
    V, v_g, h0, c0, cap = dev_inputs(make_inputs(dims, 2, 3, seed=2))
    with pytest.raises(ValueError):
        F_aa.decoder_forward(tuple(W), V, v_g, cap, h0, c0)
