"""nn.Module drop-in surface on the GPU: Encoder2Decoder.forward / .sampler against the oracle."""
import numpy as np
import pytest
import torch

import adaptive_b200
from adaptive_b200 import functional as F_aa
from adaptive_b200.synth import Dims, make_inputs, make_lengths, make_weights
from oracle import adaptive_oracle as orc
from tests.helpers import rel_err

pytestmark = pytest.mark.gpu


class Cf:
    adaptive_word_embed_size = 32
    adaptive_lstm_hidden_size = 64
    vocab_length = 120


def _model_with(w):
    m = adaptive_b200.Encoder2Decoder(Cf()).cuda()
    sd = {"decoder." + k: torch.from_numpy(v) for k, v in w.items()}
    missing, unexpected = m.load_state_dict(sd, strict=False)
    assert not unexpected and all(k.startswith("encoder.") for k in missing)
    return m


def test_e2d_forward_backward_and_sampler():
    dims = Dims(H=64, E=32, Vc=120, k=49)
    B, T = 6, 7
    w = make_weights(dims, seed=31, bias_scale=0.1)
    inp = make_inputs(dims, B, T, seed=32)
    m = _model_with(w)
    V = torch.from_numpy(inp["V"]).cuda()
    v_g = torch.from_numpy(inp["v_g"]).cuda()
    states = (torch.from_numpy(inp["h0"]).cuda()[:, None], torch.from_numpy(inp["c0"]).cuda()[:, None])   # [B,1,H] like the encoder
    cap = torch.from_numpy(inp["captions"]).cuda()
    lengths = make_lengths(B, T, seed=5)
    packed = m((V, v_g, states), cap, lengths)
    data_ref, bs_ref = orc.e2d_forward(w, inp["V"], inp["v_g"], inp["captions"], lengths, inp["h0"], inp["c0"])
    assert np.array_equal(packed.batch_sizes.numpy(), bs_ref)
    assert rel_err(packed.data.detach().cpu().numpy(), data_ref) < 1e-4
    tgt = torch.from_numpy(orc.packed_targets(inp["captions"], lengths)).cuda()
    loss = torch.nn.CrossEntropyLoss()(packed.data, tgt)      # the caller's own loss (train.py:63,208)
    loss.backward()
    total_norm = torch.nn.utils.clip_grad_norm_(m.decoder.LSTM.parameters(), 5.0)   # train.py:214
    assert torch.isfinite(total_norm)
    assert all(p.grad is not None for p in m.decoder.parameters())
    ids, att, Beta = m.sampler((V, v_g, states), max_len=5)
    r_ids, r_att, r_bet = orc.greedy_decode(w, inp["V"], inp["v_g"], inp["h0"], inp["c0"], 5)
    assert ids.shape == (B, 5) and att.shape == (B, 5, 49) and Beta.shape == (B, 5, 1)
    assert np.array_equal(ids.cpu().numpy(), r_ids)
    assert rel_err(att.cpu().numpy(), r_att) < 1e-4 and rel_err(Beta.cpu().numpy(), r_bet) < 1e-4


def test_feature_map_entry_and_decoder_states():
    m = adaptive_b200.Encoder2Decoder(Cf()).cuda()
    feats = torch.relu(torch.randn(3, 2048, 7, 7, device="cuda"))
    cap = torch.randint(4, 120, (3, 4), device="cuda")
    cap[:, 0] = 1
    out = m(feats, cap, [3, 3, 2])
    assert out.data.shape == (8, 120)
    ids, att, Beta = m.sampler(feats, max_len=3)
    assert ids.shape == (3, 3)
    V, v_g, states = m.encoder(feats)
    s, a, b, (hn, cn) = m.decoder(V, v_g, cap, states)
    assert s.shape == (3, 4, 120) and b.shape == (3, 4, 1) and hn.shape == (1, 3, 64) and cn.shape == (1, 3, 64)


def test_host_pipeline_matches_sequential_loop():
    """HostPipeline (double-buffered H2D upload, results read one step late) returns, per batch, exactly what the plain
    copy -> step -> read loop returns."""
    import torch
    from adaptive_b200.pipeline import HostPipeline

    torch.manual_seed(0)
    batches = [{"x": torch.randn(64, 33).pin_memory(), "y": torch.randint(0, 9, (64,)).pin_memory()} for _ in range(7)]
    w = torch.randn(33, 5, device="cuda")
    static_out = torch.empty(5, device="cuda")

    def step(b):                 # returns a tensor it overwrites on the next call, like a CUDA graph's static output
        static_out.copy_(((b["x"] @ w) * b["y"].float().unsqueeze(1)).sum(0))
        return static_out

    want = [step({k: v.cuda() for k, v in hb.items()}).cpu().clone() for hb in batches]
    pipe = HostPipeline(step, batches[0], "cuda")
    got = pipe.run(iter(batches))
    assert len(got) == len(want)
    for g, wv in zip(got, want):
        assert torch.equal(g, wv)
    assert pipe.h2d_bytes == 64 * 33 * 4 + 64 * 8 and pipe.d2h_bytes == 20


def test_fused_clip_adam_matches_torch_clip_and_adam():
    """aa_clip_adam_step == clip_grad_norm_(LSTM params, 5) + torch.optim.Adam(lr 1e-3, betas (0.8, 0.999), weight_decay 0) of
    the reference (train.py:213-214, model_factory.py:69-77, cfg_wzn.py:47-51), over several steps, with and without the clip
    being active; with weight decay the update of elements where g + wd*p cancels is ill-conditioned (it is ~lr*sign), so that
    variant is compared on the well-conditioned elements only."""
    import torch
    from adaptive_b200.optim import FusedClipAdam

    shapes = [(300, 40), (2048, 96), (2048,), (17,), (5, 7, 3)]
    clip_idx = [1, 2]
    for wd in (0.0, 0.01):
        torch.manual_seed(1)
        ours = [torch.randn(s, device="cuda") for s in shapes]
        ref = [p.clone().requires_grad_(True) for p in ours]
        opt_ref = torch.optim.Adam(ref, lr=1e-3, betas=(0.8, 0.999), weight_decay=wd)
        opt = FusedClipAdam(ours, clip_params=[ours[i] for i in clip_idx], lr=1e-3, betas=(0.8, 0.999), weight_decay=wd, max_norm=5.0)
        for step in range(4):
            scale = 10.0 if step % 2 == 0 else 0.001          # clip active / inactive
            grads = [torch.randn(s, device="cuda") * scale for s in shapes]
            for p, r, g in zip(ours, ref, grads):
                p.grad = g.clone()
                r.grad = g.clone()
            want_norm = torch.nn.utils.clip_grad_norm_([ref[i] for i in clip_idx], 5.0)
            opt_ref.step()
            opt.step()
            torch.cuda.synchronize()
            assert abs(float(opt.grad_norm) - float(want_norm)) <= 1e-4 * float(want_norm)
            for i, (p, r) in enumerate(zip(ours, ref)):
                assert torch.allclose(p.grad, r.grad, rtol=1e-6, atol=1e-9), (step, i)     # clipped gradients written back like torch
                diff = (p - r.detach()).abs()
                if wd == 0.0:
                    assert float(diff.max()) <= 1e-6, (step, i, float(diff.max()))
                else:
                    assert float(diff.max()) <= 2e-5 and float(diff.median()) <= 2e-7, (step, i, float(diff.max()), float(diff.median()))


@pytest.mark.parametrize("n,Vc", [(37, 10000), (5, 1000), (9, 20000), (3, 48000), (11, 1003), (1, 8)])
def test_cross_entropy_variants_vs_torch(n, Vc):
    """Mean CE + gradient (train.py:63,208): the register-resident kernels (one read of the logits, one exp per element) for
    every row length class and the three-pass fallback (Vc % 4 != 0), against torch's fp64 cross_entropy; the upstream
    gradient is applied on the device only when it is not 1."""
    g = torch.Generator().manual_seed(n * 7 + Vc)
    logits = (torch.randn(n, Vc, generator=g) * 3.0).cuda()
    tgt = torch.randint(0, Vc, (n,), generator=g).cuda()
    ref_in = logits.double().requires_grad_(True)
    ref = torch.nn.functional.cross_entropy(ref_in, tgt)
    ref.backward()
    for scale in (1.0, -2.5):
        x = logits.clone().requires_grad_(True)
        loss = F_aa.cross_entropy(x, tgt)
        (loss * scale).backward()
        assert abs(float(loss) - float(ref)) < 1e-5 * max(1.0, abs(float(ref)))
        err = (x.grad.double() - scale * ref_in.grad).abs().max() / ref_in.grad.abs().max()
        assert float(err) < 1e-5, (scale, float(err))


def test_copy_multi_and_scale_unless_one():
    """One-launch multi-segment copy (inputs of a captured step) with mixed dtypes, odd sizes and unaligned views; in-place
    scaling by a device scalar that is skipped bit-exactly when the scalar is 1."""
    g = torch.Generator().manual_seed(5)
    srcs = [torch.randn(80, 49, 512, generator=g).cuda(), torch.randn(80, 256, generator=g).cuda(),
            torch.randint(0, 10000, (80, 18), generator=g).cuda(), torch.randn(1003, generator=g).cuda()[1:],        # 4-byte aligned only
            torch.randint(0, 255, (77,), generator=g, dtype=torch.uint8).cuda()[3:], torch.randn(0).cuda(),
            torch.randn(7, 3, generator=g).cuda().half(), torch.randn(2, 2, generator=g).cuda(),
            torch.randn(33, generator=g).cuda().double(), torch.randn(5, generator=g).cuda()]                     # 10 pairs: two launches
    dsts = [torch.zeros_like(s) for s in srcs]
    dsts[3] = torch.zeros(1003, device="cuda")[1:]
    dsts[4] = torch.zeros(77, dtype=torch.uint8, device="cuda")[3:]
    F_aa.copy_multi(dsts, srcs)
    torch.cuda.synchronize()
    for d, s in zip(dsts, srcs):
        assert torch.equal(d, s)
    with pytest.raises(ValueError):
        F_aa.copy_multi([torch.zeros(4, device="cuda")], [torch.zeros(5, device="cuda")])
    from adaptive_b200 import _lib
    lib = _lib.load()
    x = torch.randn(100003, generator=g).cuda()
    x0 = x.clone()
    one, half = torch.ones((), device="cuda"), torch.full((), 0.5, device="cuda")
    x16 = x.to(torch.bfloat16)
    _lib.check(lib.aa_scale_unless_one(F_aa._ptr(x), F_aa._ptr(one), x.numel(), F_aa._ptr(x16), F_aa._stream(x.device)), "scale")
    assert torch.equal(x, x0) and torch.equal(x16, x0.to(torch.bfloat16))
    _lib.check(lib.aa_scale_unless_one(F_aa._ptr(x), F_aa._ptr(half), x.numel(), F_aa._ptr(x16), F_aa._stream(x.device)), "scale")
    assert torch.equal(x, x0 * 0.5) and torch.equal(x16, (x0 * 0.5).to(torch.bfloat16))
    _lib.check(lib.aa_scale_unless_one(F_aa._ptr(x), F_aa._ptr(half), x.numel(), None, F_aa._stream(x.device)), "scale")
    assert torch.equal(x, x0 * 0.25)


@pytest.mark.parametrize("scale", [1.0, 0.25])
def test_bf16_step_with_fused_loss_matches_torch_loss(scale):
    """bf16 training step through the module surface: the library's loss (whose backward hands the bf16 mirror of its
    gradient straight to the vocabulary projection's backward) against torch's own CrossEntropyLoss on the same packed scores
    (plain autograd: the backward makes its own bf16 copy).  Same fp32 gradient, same bf16 rounding: every parameter gradient
    must agree to fp32 accuracy, also when the loss is scaled before backward() (upstream gradient != 1)."""
    dims = Dims(H=128, E=64, Vc=504, k=49)
    B, T = 11, 9

    class Cf2:
        adaptive_word_embed_size, adaptive_lstm_hidden_size, vocab_length = dims.E, dims.H, dims.Vc

    w = make_weights(dims, seed=71, bias_scale=0.1)
    inp = make_inputs(dims, B, T, seed=72)
    lengths = make_lengths(B, T, seed=73)
    tgt = torch.from_numpy(orc.packed_targets(inp["captions"], lengths)).cuda()
    grads = {}
    for which in ("fused", "torch"):
        m = adaptive_b200.Encoder2Decoder(Cf2()).cuda()
        m.load_state_dict({"decoder." + k: torch.from_numpy(v) for k, v in w.items()}, strict=False)
        m.decoder.precision = "bf16"
        states = (torch.from_numpy(inp["h0"]).cuda()[:, None], torch.from_numpy(inp["c0"]).cuda()[:, None])
        packed = m((torch.from_numpy(inp["V"]).cuda(), torch.from_numpy(inp["v_g"]).cuda(), states), torch.from_numpy(inp["captions"]).cuda(),
                   lengths)
        loss = F_aa.cross_entropy(packed.data, tgt) if which == "fused" else torch.nn.CrossEntropyLoss()(packed.data, tgt)
        (loss * scale).backward()
        grads[which] = {n: p.grad.detach().cpu().numpy() for n, p in m.decoder.named_parameters()}
        grads[which]["loss"] = np.float64(float(loss))
    assert abs(grads["fused"]["loss"] - grads["torch"]["loss"]) < 1e-5 * abs(grads["torch"]["loss"])
    # (split-K contractions sum their partial tiles in a run-to-run varying order: 1e-4, not bit equality)
    bad = {n: rel_err(grads["fused"][n], grads["torch"][n]) for n in grads["torch"] if n != "loss" and not rel_err(grads["fused"][n], grads["torch"][n]) < 1e-4}
    assert not bad, bad
