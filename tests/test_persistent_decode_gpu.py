"""GPU parity tests of the persistent greedy decoder (``aa_decode_persistent``: the whole sampler loop of
``adaptive_attention.py:186-216`` in one cooperative launch, V / P / cell state resident in shared memory) and of the
CUDA-graph replay of the per-step pipeline (``graphs.GraphedSampler``).

Bar (BASELINE.json north_star): greedy ids exact against the reference's fp64 run (near-ties, top-1/top-2 gap < 1e-4, logged),
alpha / beta within 1e-4 relative, attention arg-max exact wherever the ids agree."""
import numpy as np
import pytest
import torch

import adaptive_b200
from adaptive_b200 import functional as F_aa
from adaptive_b200.synth import CFG_A, Dims, make_inputs, make_weights
from oracle import adaptive_oracle as orc
from tests.gpu_utils import dev_inputs, dev_weights, near_tie_report
from tests.helpers import GOLDEN_CASES, golden_setup, rel_err

pytestmark = pytest.mark.gpu
TOL = 1e-4
NEAR_TIE = 1e-4


@pytest.mark.parametrize("case", GOLDEN_CASES)
def test_persistent_vs_golden(case):
    """Ids, alpha, beta of the UNMODIFIED reference (fp64 run) on the golden cases: tiny / odd (ragged sizes) / k196 / cfgA."""
    g, dims, B, T, L, w, inp = golden_setup(case, np.float32)
    W = dev_weights(w)
    V, v_g, h0, c0, _ = dev_inputs(inp)
    if not F_aa.persistent_decode_supported(W, V, v_g, L):
        assert dims.H % 8 != 0 or dims.k * dims.H * 4 > 150 * 1024, "this case should fit the persistent kernel"
        with pytest.raises(RuntimeError, match="does not fit|multiples of 4"):
            F_aa.greedy_decode(W, V, v_g, h0, c0, L, engine="persistent")
        return
    ids, att, bet, cand = F_aa.greedy_decode_persistent(W, V, v_g, h0, c0, L, return_candidates=True)
    ids, att, bet = ids.cpu().numpy(), att.cpu().numpy(), bet.cpu().numpy()
    ref_ids, gap = g["f64_greedy_ids"], g["f64_greedy_gap"]
    hard, near = near_tie_report(ids, ref_ids, gap, NEAR_TIE)
    assert not hard, hard
    same = (ids == ref_ids).all(1)
    assert same.any()
    assert rel_err(att[same], g["f64_greedy_alpha"][same]) < TOL
    assert rel_err(bet[same], g["f64_greedy_beta"][same]) < TOL
    assert np.array_equal(att[same].argmax(-1), g["f64_greedy_alpha"][same].argmax(-1))
    c = cand.cpu().numpy()
    assert (c >= 1).all() and c.mean() < 0.25 * dims.Vc      # the filter keeps the winner and rejects most columns


@pytest.mark.parametrize("B", [1, 5, 64, 148])
def test_persistent_matches_pipeline_and_oracle(B):
    """Config-3 shapes at the batch sizes the persistent kernel serves (one image per SM): identical ids to the per-step
    pipeline (both fp32-accurate; near-ties excepted), and to the fp64 oracle; alpha / beta within 1e-4."""
    dims, L = CFG_A, 20
    w = make_weights(dims, seed=123)
    inp = make_inputs(dims, B, 1, seed=77 + B)
    W = dev_weights(w)
    V, v_g, h0, c0, _ = dev_inputs(inp)
    if not F_aa.persistent_decode_supported(W, V, v_g, L):
        pytest.skip("B=%d exceeds one image per SM on this device" % B)
    ids_p, att_p, bet_p = F_aa.greedy_decode(W, V, v_g, h0, c0, L, engine="pipeline")
    ids, att, bet = F_aa.greedy_decode(W, V, v_g, h0, c0, L, engine="persistent")
    ids2, att2, bet2 = F_aa.greedy_decode(W, V, v_g, h0, c0, L, engine="auto")          # takes the persistent kernel, reuses the packed weights
    assert torch.equal(ids, ids2) and torch.equal(att, att2) and torch.equal(bet, bet2)   # deterministic, also with the reused workspace
    w64 = {k: v.astype(np.float64) for k, v in w.items()}
    i64 = {k: (v.astype(np.float64) if v.dtype != np.int64 else v) for k, v in inp.items()}
    nb = min(B, 32)                                                                       # (the fp64 oracle is the slow part)
    ref_ids, ref_att, ref_bet, ref_sc = orc.greedy_decode(w64, i64["V"][:nb], i64["v_g"][:nb], i64["h0"][:nb], i64["c0"][:nb], L,
                                                          want_scores=True)
    top2 = np.sort(ref_sc, axis=-1)[..., -2:]
    gap = top2[..., 1] - top2[..., 0]
    hard, near = near_tie_report(ids.cpu().numpy()[:nb], ref_ids, gap, NEAR_TIE)
    assert not hard, hard
    same = (ids.cpu().numpy()[:nb] == ref_ids).all(1)
    assert same.mean() > 0.9
    assert rel_err(att.cpu().numpy()[:nb][same], ref_att[same]) < TOL
    assert rel_err(bet.cpu().numpy()[:nb][same].reshape(-1), ref_bet[same].reshape(-1)) < TOL
    assert np.array_equal(att.cpu().numpy()[:nb][same].argmax(-1), ref_att[same].argmax(-1))
    # against the pipeline: rows whose ids agree everywhere carry the same attention to fp32 accuracy
    agree = (ids == ids_p).all(1)
    assert agree.float().mean() > 0.9
    assert rel_err(att[agree].cpu().numpy(), att_p[agree].cpu().numpy()) < 2e-5
    assert rel_err(bet[agree].cpu().numpy(), bet_p[agree].cpu().numpy()) < 2e-5


def test_persistent_sharding_and_weight_updates():
    """An image's caption does not depend on its batch mates (shard == slice of the whole: the property decode sharding rests
    on), and the cached packed weights are rebuilt when a weight tensor changes in place."""
    dims, B, L = Dims(H=128, E=64, Vc=1000, k=49), 40, 8
    w = make_weights(dims, seed=5, bias_scale=0.1)
    inp = make_inputs(dims, B, 1, seed=6)
    W = dev_weights(w)
    V, v_g, h0, c0, _ = dev_inputs(inp)
    ids, att, bet = F_aa.greedy_decode_persistent(W, V, v_g, h0, c0, L)
    lo, hi = 13, 29
    ids_s, att_s, bet_s = F_aa.greedy_decode_persistent(W, V[lo:hi], v_g[lo:hi], h0[lo:hi], c0[lo:hi], L)
    assert torch.equal(ids[lo:hi], ids_s) and torch.equal(att[lo:hi], att_s) and torch.equal(bet[lo:hi], bet_s)
    # no initial state = zeros
    z = torch.zeros_like(h0)
    a1 = F_aa.greedy_decode_persistent(W, V, v_g, None, None, L)
    a2 = F_aa.greedy_decode_persistent(W, V, v_g, z, z, L)
    assert all(torch.equal(x, y) for x, y in zip(a1, a2))
    # in-place weight update (what an optimizer step does): the reused workspace must not serve stale packed weights
    W[11].mul_(-1.0)                                   # mlp.weight
    ids_n = F_aa.greedy_decode_persistent(W, V, v_g, h0, c0, L)[0]
    ids_ref = F_aa.greedy_decode(W, V, v_g, h0, c0, L, engine="pipeline")[0]
    assert (ids_n == ids_ref).float().mean() > 0.95 and not torch.equal(ids_n, ids)


def test_auto_engine_takes_two_persistent_launches_up_to_two_images_per_sm():
    """B in (SMs, 2 SMs]: the auto engine decodes the batch as two persistent launches; same ids as the pipeline (near-ties excepted),
    and row i of the result is image i (no reordering by the split)."""
    dims, L = CFG_A, 10
    sms = torch.cuda.get_device_properties(0).multi_processor_count
    B = 2 * sms - 3
    w = make_weights(dims, seed=123)
    W = dev_weights(w)
    inp = make_inputs(dims, B, 1, seed=17)
    V, v_g, h0, c0, _ = dev_inputs(inp)
    ids_a, att_a, bet_a = F_aa.greedy_decode(W, V, v_g, h0[None], c0[None], L, engine="auto")      # [1,B,H] states: cut along dim 1
    ids_p, att_p, bet_p = F_aa.greedy_decode(W, V, v_g, h0, c0, L, engine="pipeline")
    assert ids_a.shape == (B, L) and att_a.shape == (B, L, dims.k)
    agree = (ids_a == ids_p).all(1)
    assert agree.float().mean() > 0.9
    assert rel_err(att_a[agree].cpu().numpy(), att_p[agree].cpu().numpy()) < 2e-5
    half = (B + 1) // 2
    lo = F_aa.greedy_decode_persistent(W, V[half:], v_g[half:], h0[half:], c0[half:], L)[0]
    assert torch.equal(ids_a[half:], lo)


def test_persistent_rejects_what_it_cannot_hold():
    dims = CFG_A
    w = make_weights(dims, seed=123)
    W = dev_weights(w)
    B = 2 * torch.cuda.get_device_properties(0).multi_processor_count + 1
    inp = make_inputs(dims, B, 1, seed=3)
    V, v_g, h0, c0, _ = dev_inputs(inp)
    assert not F_aa.persistent_decode_supported(W, V, v_g, 5)
    with pytest.raises(RuntimeError, match="does not fit"):
        F_aa.greedy_decode(W, V, v_g, h0, c0, 5, engine="persistent")
    ids_a = F_aa.greedy_decode(W, V, v_g, h0, c0, 5, engine="auto")[0]        # more than two images per SM: falls back to the pipeline
    ids_p = F_aa.greedy_decode(W, V, v_g, h0, c0, 5, engine="pipeline")[0]
    assert torch.equal(ids_a, ids_p)


@pytest.mark.parametrize("engine", ["pipeline", "persistent"])
def test_graphed_sampler_equals_eager(engine):
    """The whole sampler loop captured into one CUDA graph (one launch per batch) returns exactly what the eager calls return,
    for new inputs copied into its static buffers."""
    from adaptive_b200.graphs import GraphedSampler

    dims, B, L = Dims(H=128, E=64, Vc=1000, k=49), 24, 7

    class Cf:
        adaptive_word_embed_size, adaptive_lstm_hidden_size, vocab_length = dims.E, dims.H, dims.Vc
        decode_engine = engine

    model = adaptive_b200.Encoder2Decoder(Cf()).cuda()
    w = make_weights(dims, seed=21, bias_scale=0.1)
    model.load_state_dict({"decoder." + k: torch.from_numpy(v) for k, v in w.items()}, strict=False)
    batches = []
    for s in (1, 2):
        inp = make_inputs(dims, B, 1, seed=30 + s)
        V, v_g, h0, c0, _ = dev_inputs(inp)
        batches.append({"V": V, "v_g": v_g, "h0": h0, "c0": c0})
    gs = GraphedSampler(model, batches[0], max_len=L)
    for b in batches[::-1]:
        want = model.sampler((b["V"], b["v_g"], (b["h0"], b["c0"])), max_len=L)
        got = gs(b)
        assert all(torch.equal(x, y) for x, y in zip(want, got))
