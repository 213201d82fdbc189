"""GPU parity tests of the fused forward + cross-entropy operator (``aa_decoder_forward_loss``: the loss of ``train.py:63,208``
inside the vocabulary projection's tcgen05 epilogue, SURVEY section 8f rank 1 -- the ``[n_rows, Vc]`` logits are never written).

Bar: the bf16 training path's tolerance (BASELINE.json north_star: 2e-2 relative for gradients); the loss itself is an fp32
reduction of fp32 accumulators and is held to 1e-3."""
import numpy as np
import pytest
import torch

import adaptive_b200
from adaptive_b200 import functional as F_aa
from adaptive_b200.synth import CFG_A, Dims, make_inputs, make_lengths, make_weights
from oracle import adaptive_oracle as orc
from tests.gpu_utils import dev_inputs, dev_weights, grad_key_order
from tests.helpers import rel_err

pytestmark = pytest.mark.gpu
BF16_TOL = 2e-2


def _oracle_step(w, inp, lengths, B, T, dims):
    w64 = {k: v.astype(np.float64) for k, v in w.items()}
    i64 = {k: (v.astype(np.float64) if v.dtype != np.int64 else v) for k, v in inp.items()}
    s_o, _, _, _, cache = orc.decoder_forward(w64, i64["V"], i64["v_g"], i64["captions"], i64["h0"], i64["c0"], want_cache=True)
    data, _ = orc.pack_padded(s_o, lengths)
    tgt = orc.packed_targets(inp["captions"], lengths)
    loss, dlog = orc.cross_entropy(data, tgt)
    idx, _ = F_aa.packed_row_index(lengths, T)
    dS = np.zeros((B * T, dims.Vc))
    dS[idx] = dlog
    return float(loss), orc.decoder_backward(w64, cache, dS.reshape(B, T, dims.Vc)), tgt


@pytest.mark.parametrize("B,T,dims", [(80, 18, CFG_A), (11, 9, Dims(H=128, E=64, Vc=504, k=49)),      # 504 = 15.75 chunks of 32 columns
                                      (6, 5, Dims(H=64, E=32, Vc=40, k=10)), (33, 6, Dims(H=128, E=64, Vc=1000, k=196))])
def test_fused_loss_vs_oracle_and_two_step_route(B, T, dims):
    w = make_weights(dims, seed=61, bias_scale=0.1)
    inp = make_inputs(dims, B, T, seed=62)
    lengths = make_lengths(B, T, seed=63)
    loss_o, G, tgt_np = _oracle_step(w, inp, lengths, B, T, dims)
    tgt = torch.from_numpy(np.ascontiguousarray(tgt_np)).cuda()

    def run(fused):
        W = dev_weights(w, requires_grad=True)
        V, v_g, h0, c0, cap = dev_inputs(inp, requires_grad=True)
        if fused:
            loss = F_aa.decoder_forward_loss(W, V, v_g, cap, lengths, tgt, h0, c0)[0]
        else:
            packed = F_aa.decoder_forward_packed(W, V, v_g, cap, lengths, h0, c0, precision="bf16")[0]
            loss = F_aa.cross_entropy(packed.data, tgt)
        loss.backward()
        torch.cuda.synchronize()
        out = {"loss": np.array([loss.item()])}
        for key, t in zip(grad_key_order(), W):
            out["d" + key] = t.grad.cpu().numpy()
        for key, t in (("V", V), ("v_g", v_g), ("h0", h0), ("c0", c0)):
            out["d" + key] = t.grad.cpu().numpy()
        return out

    a, b = run(True), run(False)
    assert abs(a["loss"][0] - loss_o) < 1e-3 * abs(loss_o), (a["loss"], loss_o)
    assert abs(a["loss"][0] - b["loss"][0]) < 1e-3 * abs(loss_o)
    errs = {k: rel_err(a[k], G[k[1:]]) for k in a if k != "loss"}
    bad = {k: v for k, v in errs.items() if not v < BF16_TOL}
    assert not bad, (bad, errs)
    # against the two-step bf16 route: same operands, only the rounding of the logits' gradient differs
    bad2 = {k: rel_err(a[k], b[k]) for k in a if k != "loss" and not rel_err(a[k], b[k]) < 1e-2}
    assert not bad2, bad2


def test_fused_route_is_chosen_by_size(monkeypatch):
    monkeypatch.delenv("AA_FUSED_CE", raising=False)
    assert not F_aa.fused_loss_pays(943, 10000)          # config 2: 38 MB of logits stay in L2
    assert F_aa.fused_loss_pays(3000, 20000)             # config 5: 240 MB do not
    monkeypatch.setenv("AA_FUSED_CE", "1")
    assert F_aa.fused_loss_pays(10, 10)
    monkeypatch.setenv("AA_FUSED_CE", "0")
    assert not F_aa.fused_loss_pays(3000, 20000)


def test_default_targets_upstream_scale_and_double_backward():
    dims, B, T = Dims(H=128, E=64, Vc=504, k=49), 9, 7
    w = make_weights(dims, seed=71, bias_scale=0.1)
    inp = make_inputs(dims, B, T, seed=72)
    lengths = make_lengths(B, T, seed=73)
    W1 = dev_weights(w, requires_grad=True)
    V, v_g, h0, c0, cap = dev_inputs(inp)
    l1 = F_aa.decoder_forward_loss(W1, V, v_g, cap, lengths, None, h0, c0)[0]                  # targets default to the packed next words
    tgt = torch.from_numpy(np.ascontiguousarray(F_aa.packed_targets(inp["captions"], lengths))).cuda()
    l2 = F_aa.decoder_forward_loss(W1, V, v_g, cap, lengths, tgt, h0, c0)[0]
    assert abs(l1.item() - l2.item()) < 1e-5 * abs(l2.item())      # (the row terms are summed with atomics: order varies)
    l1.backward()
    with pytest.raises(RuntimeError, match="twice|second time"):
        l1.backward()
    # loss not the root: gradients scale with the upstream factor (decided on the device)
    W2 = dev_weights(w, requires_grad=True)
    (2.5 * F_aa.decoder_forward_loss(W2, V, v_g, cap, lengths, tgt, h0, c0)[0]).backward()
    for a, b in zip(W1, W2):
        assert rel_err(b.grad.cpu().numpy(), 2.5 * a.grad.cpu().numpy()) < 1e-2
    # the exact path has no fused loss: the operator says so, the module takes the two-step route
    with pytest.raises(RuntimeError, match="AA_PREC_BF16|bf16"):
        lib, d = F_aa._lib.load(), F_aa.make_dims(B, T, dims.k, dims.H, dims.E, dims.Vc, dims.a, F_aa._lib.PREC_FP32)
        F_aa.check(lib.aa_decoder_forward_loss(F_aa.ctypes.byref(d), F_aa.ctypes.byref(F_aa.weights_struct(W1)), None, None, None, None, None,
                                               F_aa._ptr(tgt), 1, F_aa._ptr(tgt), 0, F_aa._ptr(V), None, None, None, None, None, 0, None),
                   "aa_decoder_forward_loss")


@pytest.mark.parametrize("prec", ["bf16", "fp32"])
def test_module_forward_loss_and_graphed_step(prec, monkeypatch):
    """``Encoder2Decoder.forward_loss`` == criterion(forward(...).data, targets) of train.py:205-208 on both precisions, and the
    graphed training step (which goes through it) leaves the same gradients as the eager two-step route."""
    from adaptive_b200.graphs import GraphedTrainStep

    monkeypatch.setenv("AA_FUSED_CE", "1")      # (by default small problems take the two-step route: functional.fused_loss_pays)
    dims, B, T = Dims(H=128, E=64, Vc=1000, k=49), 16, 8

    class Cf:
        adaptive_word_embed_size, adaptive_lstm_hidden_size, vocab_length = dims.E, dims.H, dims.Vc
        precision = prec

    model = adaptive_b200.Encoder2Decoder(Cf()).cuda()
    w = make_weights(dims, seed=81, bias_scale=0.1)
    model.load_state_dict({"decoder." + k: torch.from_numpy(v) for k, v in w.items()}, strict=False)
    inp = make_inputs(dims, B, T, seed=82)
    lengths = make_lengths(B, T, seed=83)
    b = {k: torch.from_numpy(v).cuda() for k, v in inp.items()}
    b["tgt"] = torch.from_numpy(np.ascontiguousarray(F_aa.packed_targets(inp["captions"], lengths))).cuda()
    enc = (b["V"], b["v_g"], (b["h0"], b["c0"]))
    params = list(model.decoder.parameters())
    packed = model(enc, b["captions"], lengths)
    ref = F_aa.cross_entropy(packed.data, b["tgt"])
    ref.backward()
    g_ref = [p.grad.clone() for p in params]
    for p in params:
        p.grad = None
    loss = model.forward_loss(enc, b["captions"], lengths, b["tgt"])
    loss.backward()
    tol = 1e-2 if prec == "bf16" else 1e-6
    ref_v = ref.item()
    assert abs(loss.item() - ref_v) < 1e-3 * abs(ref_v)
    for p, g in zip(params, g_ref):
        assert rel_err(p.grad.cpu().numpy(), g.cpu().numpy()) < tol
    # (an autograd graph of an eager step that is still alive keeps its AccumulateGrad nodes bound to the stream they were created on:
    #  capturing a step on another stream then synchronises with the default stream and invalidates the capture)
    del packed, ref, loss
    step = GraphedTrainStep(model, b, lengths)
    l3 = step(b)
    torch.cuda.synchronize()
    assert abs(l3.item() - ref_v) < 1e-3 * abs(ref_v)
    for p, g in zip(params, g_ref):
        assert rel_err(p.grad.cpu().numpy(), g.cpu().numpy()) < tol


def test_fused_loss_sentinel_less_baseline_model(monkeypatch):
    """The baseline decoder (baseline_attention.py:66-194: no sentinel, beta = 0) through the fused operator: NULL sentinel weights like
    every other entry point; loss and gradients against the oracle's baseline mode, and ``baseline.Encoder2Decoder.forward_loss``."""
    from adaptive_b200 import baseline
    from adaptive_b200.synth import baseline_weights

    monkeypatch.setenv("AA_FUSED_CE", "1")
    dims, B, T = Dims(H=128, E=64, Vc=504, k=49), 10, 7
    w_full = make_weights(dims, seed=91, bias_scale=0.1)
    wb = baseline_weights(w_full)
    inp = make_inputs(dims, B, T, seed=92)
    lengths = make_lengths(B, T, seed=93)
    loss_o, G, tgt_np = _oracle_step(wb, inp, lengths, B, T, dims)
    tgt = torch.from_numpy(np.ascontiguousarray(tgt_np)).cuda()
    W = F_aa.baseline_weights(dev_weights(w_full, requires_grad=True))
    V, v_g, h0, c0, cap = dev_inputs(inp)
    loss = F_aa.decoder_forward_loss(W, V, v_g, cap, lengths, tgt, h0, c0)[0]
    loss.backward()
    assert abs(loss.item() - loss_o) < 1e-3 * abs(loss_o)
    for key, t in zip(grad_key_order(), W):
        if t is not None:
            assert rel_err(t.grad.cpu().numpy(), G[key]) < BF16_TOL, key

    class Cf:
        base_word_embed_size, base_lstm_hidden_size, vocab_length = dims.E, dims.H, dims.Vc
        precision = "bf16"

    m = baseline.Encoder2Decoder(Cf()).cuda()
    m.decoder.load_state_dict({k: torch.from_numpy(v) for k, v in wb.items()}, strict=True)
    l2 = m.forward_loss((V, v_g, (h0, c0)), cap, lengths, tgt)
    assert abs(l2.item() - loss_o) < 1e-3 * abs(loss_o)
