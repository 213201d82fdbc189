"""CPU-side tests: C-ABI library loads and exports every declared symbol, host logic, synth."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

import adaptive_b200
from adaptive_b200 import _lib
from adaptive_b200 import functional as F_aa
from adaptive_b200.synth import CFG_A, DECODER_KEYS, Dims, make_inputs, make_lengths, make_weights

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "adaptive_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(aa_[a-z_A-Z0-9]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    lib = ctypes.CDLL(_lib.LIB_PATH)
    names = _declared_symbols()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), "missing export " + n
    assert sorted(_lib.SIGNATURES) == names          # the ctypes table covers the header exactly
    assert _lib.load().aa_version() == 100


def test_no_cpu_fallback():
    dims = Dims(H=32, E=16, Vc=40, k=49)
    w = make_weights(dims)
    W = tuple(torch.from_numpy(w[k]) for k in DECODER_KEYS)
    inp = make_inputs(dims, 2, 3)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        F_aa.decoder_forward(W, torch.from_numpy(inp["V"]), torch.from_numpy(inp["v_g"]), torch.from_numpy(inp["captions"]))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        F_aa.greedy_decode(W, torch.from_numpy(inp["V"]), torch.from_numpy(inp["v_g"]))


def test_size_queries_work_without_gpu():
    lib = _lib.load()
    d = F_aa.make_dims(80, 18, 49, 512, 256, 10000)
    assert lib.aa_decoder_saved_bytes(ctypes.byref(d)) > 80 * 18 * 512 * 4 * 10
    assert lib.aa_decoder_bwd_scratch_bytes(ctypes.byref(d)) > 0
    assert lib.aa_decode_workspace_bytes(ctypes.byref(d), 3) > lib.aa_decode_workspace_bytes(ctypes.byref(d), 0)


def test_state_dict_keys_match_reference():
    m = adaptive_b200.Encoder2Decoder()
    keys = [k[len("decoder."):] for k in m.state_dict() if k.startswith("decoder.")]
    assert tuple(keys) == DECODER_KEYS                                     # SURVEY.md §8b, registration order
    shapes = CFG_A.shapes()
    for k in keys:
        assert tuple(m.state_dict()["decoder." + k].shape) == shapes[k]
    enc = {k for k in m.state_dict() if k.startswith("encoder.")}
    assert enc == {"encoder.affine_%s.%s" % (n, p) for n in ("a", "b", "h0", "c0") for p in ("weight", "bias")}
    assert sum(p.numel() for p in m.decoder.parameters()) == 10390849
    # LSTM init: forget-gate slices of both biases are 0.5 (model_utils.py:62-74)
    H = 512
    assert torch.all(m.decoder.LSTM.bias_ih_l0[H:2 * H] == 0.5) and torch.all(m.decoder.LSTM.bias_hh_l0[:H] == 0)


def test_reference_checkpoint_loads_through_helper():
    """A state_dict produced by the reference's own ``Encoder2Decoder`` (ResNet-152 trunk included: 930 extra keys) loads through
    ``load_reference_state_dict`` -- decoder + encoder-head tensors by key, trunk dropped -- and plain strict loading says why it
    cannot.  Needs ``oracle/_ref`` (staged by build() where /root/reference exists)."""
    from oracle import build_ref

    if not build_ref.available():
        pytest.skip("oracle/_ref not staged")
    ada, _ = build_ref.import_reference()

    class Cf:
        adaptive_word_embed_size, adaptive_lstm_hidden_size, vocab_length = 16, 32, 50

    ref_model = build_ref.make_encoder2decoder(ada, Cf(), identity_trunk=False)
    sd = ref_model.state_dict()
    assert sum(k.startswith("encoder.resnet_conv.") for k in sd) > 900
    m = adaptive_b200.Encoder2Decoder(Cf())
    with pytest.raises(RuntimeError, match="Unexpected key"):
        m.load_state_dict(sd)
    m.load_reference_state_dict(sd)
    for k, v in m.state_dict().items():
        assert torch.equal(v, sd[k]), k
    bad = dict(sd)
    del bad["decoder.adaptive.mlp.bias"]
    with pytest.raises(RuntimeError, match="Missing key"):
        m.load_reference_state_dict(bad)


def test_packed_row_index_matches_torch():
    for lengths, T in (([5, 5, 3, 1], 6), ([17] * 4, 18), ([2], 2)):
        idx, bs = F_aa.packed_row_index(lengths, T)
        x = torch.arange(len(lengths) * T).view(len(lengths), T)
        ref = torch.nn.utils.rnn.pack_padded_sequence(x, lengths, batch_first=True)
        assert idx == ref.data.tolist() and bs == ref.batch_sizes.tolist()
    with pytest.raises(RuntimeError):
        F_aa.packed_row_index([1, 2], 3)


def test_synth_is_reproducible_and_shaped():
    a, b = make_weights(Dims(H=32, E=16, Vc=40, k=49)), make_weights(Dims(H=32, E=16, Vc=40, k=49))
    assert all(np.array_equal(a[k], b[k]) for k in a) and tuple(a) == DECODER_KEYS
    q = a["LSTM.weight_hh_l0"]
    assert np.allclose(q.T @ q, np.eye(32), atol=1e-5)        # orthogonal columns
    inp = make_inputs(CFG_A, 3, 18)
    assert inp["captions"][:, 0].tolist() == [1, 1, 1] and inp["captions"][:, 1:].min() >= 4
    assert inp["V"].min() >= 0 and np.abs(inp["h0"]).max() <= 1
    ls = make_lengths(80, 18)
    assert ls == sorted(ls, reverse=True) and ls[0] == 17 and min(ls) >= 6


def test_ids_to_captions_and_vocabulary_format():
    """tools/utils.py:180-195 semantics: cut at the first <end>, keep everything before it (incl. <start>/<unk>/<pad>), one
    record per image; Vocabulary ids of the specials as in build_vocab.py:48-51."""
    from adaptive_b200.postprocess import Vocabulary, caption_results, ids_to_captions

    vocab = Vocabulary(["a", "dog", "runs"])
    assert [vocab(w) for w in ("<pad>", "<start>", "<end>", "<unk>", "a", "dog", "runs", "zebra")] == [0, 1, 2, 3, 4, 5, 6, 3]
    assert len(vocab) == 7 and vocab.idx2word[5] == "dog"
    ids = np.array([[4, 5, 6, 2, 4, 4], [2, 4, 5, 6, 4, 4], [4, 3, 6, 6, 6, 6]])
    assert ids_to_captions(ids, vocab) == ["a dog runs", "", "a <unk> runs runs runs runs"]
    res = caption_results([11, 12, 13], ids, vocab)
    assert res[0] == {"image_id": 11, "caption": "a dog runs"} and [r["image_id"] for r in res] == [11, 12, 13]
    with pytest.raises(ValueError):
        caption_results([1, 2], ids, vocab)


def test_baseline_model_surface_and_weight_tuples():
    """the sentinel-less baseline classes keep the reference's state_dict keys (baseline_attention.py:66-194:
    no sentinel, no affine_s) and travel to the C ABI as the 13-tuple with three NULL entries, all three together."""
    from adaptive_b200 import baseline
    from adaptive_b200.synth import BASELINE_KEYS, baseline_weights

    m = baseline.Encoder2Decoder()
    keys = tuple(k[len("decoder."):] for k in m.state_dict() if k.startswith("decoder."))
    assert keys == BASELINE_KEYS and len(keys) == 10
    assert not any("sentinel" in k or "affine_s" in k for k in keys)
    w13 = m.decoder.weights()
    assert len(w13) == 13 and [i for i, t in enumerate(w13) if t is None] == [5, 6, 9]      # sen_wx, sen_wh, att_ws
    assert [_lib.WEIGHT_FIELDS[i] for i in (5, 6, 9)] == list(F_aa.SENTINEL_FIELDS)
    s = F_aa.weights_struct(w13)
    assert s.sen_wx is None and s.sen_wh is None and s.att_ws is None and s.att_wv is not None
    dims = Dims(H=32, E=16, Vc=40, k=49)
    w = make_weights(dims)
    assert tuple(baseline_weights(w)) == BASELINE_KEYS
    full = tuple(torch.from_numpy(w[k]) for k in DECODER_KEYS)
    assert F_aa.baseline_weights(full)[5] is None and F_aa.baseline_weights(full)[7] is full[7]
    partial = list(full)
    partial[5] = None                                        # only sen_wx missing: neither model
    inp = make_inputs(dims, 2, 3)
    with pytest.raises(ValueError, match="only together"):
        F_aa._check_weights(partial, dims.H, dims.E, dims.Vc, dims.a)
    with pytest.raises(RuntimeError, match="no CPU fallback"):                                   # baseline weights on the CPU: no fallback either
        F_aa.decoder_forward(F_aa.baseline_weights(full), torch.from_numpy(inp["V"]), torch.from_numpy(inp["v_g"]),
                             torch.from_numpy(inp["captions"]))


def test_encoder_heads_surface_and_size_queries():
    """AttentiveCNN's head parameters under the reference's names, the C-ABI structs in that order, workspace
    queries without a GPU, and no CPU path."""
    from adaptive_b200.modules import AttentiveCNN
    from adaptive_b200.synth import ENCODER_KEYS, make_encoder_weights, make_features

    enc = AttentiveCNN(16, 32, None, feat_dim=64)
    assert tuple(k for k in enc.state_dict()) == ENCODER_KEYS
    assert [_lib.ENC_KEY_TO_FIELD[k] for k in ENCODER_KEYS] == list(_lib.ENC_FIELDS)
    assert [tuple(t.shape) for t in enc.weights()] == [(32, 64), (32,), (16, 64), (16,), (32, 64), (32,), (32, 64), (32,)]
    lib = _lib.load()
    d32 = _lib.AAEncDims(B=80, C=2048, hw=49, H=512, E=256, precision=_lib.PREC_FP32)
    d16 = _lib.AAEncDims(B=80, C=2048, hw=49, H=512, E=256, precision=_lib.PREC_BF16)
    assert lib.aa_encoder_saved_bytes(ctypes.byref(d32)) >= 80 * 49 * 2048 * 4             # the fp32 transpose of the map
    assert lib.aa_encoder_saved_bytes(ctypes.byref(d16)) >= 80 * 49 * 2048 * 2 + (512 * 3 + 256) * 2048 * 2
    assert lib.aa_encoder_bwd_scratch_bytes(ctypes.byref(d32), 1) > lib.aa_encoder_bwd_scratch_bytes(ctypes.byref(d32), 0)
    w = make_encoder_weights(Dims(H=32, E=16, Vc=8, k=49), 64)
    A = make_features(2, 64, (7, 7))
    assert A.shape == (2, 64, 7, 7) and A.min() >= 0 and tuple(w) == ENCODER_KEYS
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        enc(torch.from_numpy(A))


def _round_bf16(x):
    """fp32 -> bf16 (round to nearest even) -> fp32, like __floats2bfloat162_rn."""
    b = np.asarray(x, np.float32).view(np.uint32).astype(np.uint64)
    b = (b + 0x7FFF + ((b >> 16) & 1)) & 0xFFFF0000
    return b.astype(np.uint32).view(np.float32)


def _round_tf32(x):
    """fp32 -> tf32 (round to nearest, ties away: cvt.rna.tf32.f32) -> fp32."""
    b = np.asarray(x, np.float32).view(np.uint32).astype(np.uint64)
    b = (b + 0x1000) & 0xFFFFE000
    return b.astype(np.uint32).view(np.float32)


@pytest.mark.parametrize("rounder,c", [(_round_bf16, 2.1 / 256), (_round_tf32, 1.1 / 1024)])
def test_argmax_filter_bound_never_drops_the_exact_argmax(rounder, c):
    """The claim csrc/vocab_refine.cu rests on, checked in numpy: with |approx_j - exact_j| <= c ||u|| ||W_j|| (c = 2.1 * 2^-8 for a bf16
    first pass -- unit roundoff 2^-8 per operand --, 1.1 * 2^-10 for tf32) the 16-column tile holding the exact arg-max always passes the filter
    `tile max + bound >= max over tiles of (tile max - bound)`, also for rows built to have near-ties, and the filter keeps few tiles."""
    rng = np.random.Generator(np.random.PCG64(5))
    H, Vc, R, TN = 256, 2000, 300, 16
    W = (rng.standard_normal((Vc, H)) * np.sqrt(2.0 / H)).astype(np.float32)
    bias = (0.1 * rng.standard_normal(Vc)).astype(np.float32)
    u = (0.4 + rng.standard_normal((R, H))).astype(np.float32)                    # common-mode component like c_hat + h
    # adversarial rows: make two far-apart columns (different tiles) tie to ~1e-6
    exact0 = u.astype(np.float64) @ W.astype(np.float64).T + bias
    for r in range(0, R, 3):
        j1, j2 = np.argsort(exact0[r])[-2:]
        if abs(j1 - j2) >= TN:
            d = W[j2].astype(np.float64) - W[j1].astype(np.float64)
            gap = exact0[r, j2] - exact0[r, j1]
            u[r] = (u[r].astype(np.float64) - (gap - 1e-6) * d / (d @ d)).astype(np.float32)
    exact = u.astype(np.float64) @ W.astype(np.float64).T + bias
    approx = (rounder(u).astype(np.float64) @ rounder(W).astype(np.float64).T).astype(np.float32) + bias     # fp32 accumulate
    assert np.abs(approx - exact).max() <= (c * np.linalg.norm(u, axis=1)[:, None] * np.linalg.norm(W, axis=1)[None, :]).max()
    assert (np.abs(approx - exact) <= c * np.linalg.norm(u, axis=1)[:, None] * np.linalg.norm(W, axis=1)[None, :] + 1e-7).all()
    tiles = Vc // TN
    tmax = approx.reshape(R, tiles, TN).max(-1)
    wn = np.linalg.norm(W, axis=1).reshape(tiles, TN).max(-1)
    b = c * np.linalg.norm(u, axis=1)[:, None] * wn[None, :]
    L = (tmax - b).max(1, keepdims=True)
    keep = tmax + b >= L
    win = exact.argmax(1) // TN
    assert keep[np.arange(R), win].all()                                          # the exact winner's tile is always refined
    # every tile holding a column within 1e-6 of the exact maximum is kept too (ties are resolved on exact values afterwards)
    near = (exact >= exact.max(1, keepdims=True) - 1e-6).reshape(R, tiles, TN).any(-1)
    assert (keep | ~near).all()
    assert 1.0 <= keep.sum(1).mean() < 0.2 * tiles


@pytest.mark.parametrize("rounder,c,ulp", [(_round_bf16, 2.1 / 256, 2.0 ** -8), (_round_tf32, 1.1 / 1024, 2.0 ** -11)])
def test_argmax_filter_bound_holds_for_sparse_worst_case_vectors(rounder, c, ulp):
    """Dense random vectors sit ~sqrt(H) below the Cauchy-Schwarz bound, so they cannot tell a correct constant from one that
    is 2x too small.  One-hot rows whose single entry sits just below a rounding midpoint attain it: u = W_j = (1 + ulp(1 - eps)) e_k
    rounds DOWN in both operands, |approx - exact| ~ 2 ulp ||u|| ||W_j||.  The constant must cover that (the round-1 value
    1.1 * 2^-8 for bf16 did not), and the filter must keep the exact winner's tile when the rows are built from such vectors."""
    H, Vc, TN = 64, 256, 16
    x = np.float32(1.0 + ulp * (1.0 - 2.0 ** -10))        # just below the midpoint between 1 and the next representable value
    assert rounder(np.array([x], np.float32))[0] == np.float32(1.0)
    err = abs(float(rounder(np.array([x]))[0]) ** 2 - float(x) ** 2)
    assert err <= c * float(x) * float(x)                  # the bound itself at its worst case ...
    assert err > 0.9 * 2 * ulp                             # ... which really is ~2 ulp: a constant of 1.1 ulp-pairs would fail here
    # rows / columns built from such vectors: column j = x e_{j mod H} (+ a competitor tile whose entries are exactly representable)
    W = np.zeros((Vc, H), np.float32)
    for j in range(Vc):
        W[j, j % H] = x if j < Vc // 2 else np.float32(1.0 + 1.5 * ulp)   # second half: representable-ish competitors, slightly smaller exact logit
    u = np.zeros((H, H), np.float32)
    u[np.arange(H), np.arange(H)] = x
    exact = u.astype(np.float64) @ W.astype(np.float64).T
    approx = (rounder(u).astype(np.float64) @ rounder(W).astype(np.float64).T).astype(np.float32)
    nu, nw = np.linalg.norm(u.astype(np.float64), axis=1), np.linalg.norm(W.astype(np.float64), axis=1)
    assert (np.abs(approx - exact) <= c * nu[:, None] * nw[None, :]).all()
    tiles = Vc // TN
    tmax = approx.reshape(H, tiles, TN).max(-1)
    b = c * nu[:, None] * nw.reshape(tiles, TN).max(-1)[None, :]
    keep = tmax + b >= (tmax - b).max(1, keepdims=True)
    win = exact.argmax(1) // TN
    assert keep[np.arange(H), win].all()


def test_argmax_filter_exact_decomposition_bound():
    """The bound the bf16 first pass is filtered with (csrc/vocab_refine.cu, round 2):
        u.W_j - u^.W^_j  =  u^.(W_j - W^_j) + (u - u^).W_j          (exact, u^ / W^ = the bf16 mirrors)
        |approx_j - exact_j| <= ||u^|| ||W_j - W^_j|| + ||u - u^|| ||W_j|| + 2^-13 ||u|| ||W_j||   (Cauchy-Schwarz + fp32 accumulation)
    checked in numpy on dense random data, on rows built to have near-ties across tiles and on the one-hot vectors that attain the
    worst-case relative bound; it must hold everywhere, keep the exact winner's tile, and be about twice as tight as the
    relative bound 2.1 * 2^-8 ||u|| ||W_j|| on dense data."""
    rng = np.random.Generator(np.random.PCG64(11))
    H, Vc, R, TN = 256, 2000, 300, 16
    W = (rng.standard_normal((Vc, H)) * np.sqrt(2.0 / H)).astype(np.float32)
    bias = (0.1 * rng.standard_normal(Vc)).astype(np.float32)
    u = (0.4 + rng.standard_normal((R, H))).astype(np.float32)
    exact0 = u.astype(np.float64) @ W.astype(np.float64).T + bias
    for r in range(0, R, 3):                                  # near-ties between far-apart columns
        j1, j2 = np.argsort(exact0[r])[-2:]
        if abs(j1 - j2) >= TN:
            d = W[j2].astype(np.float64) - W[j1].astype(np.float64)
            u[r] = (u[r].astype(np.float64) - (exact0[r, j2] - exact0[r, j1] - 1e-6) * d / (d @ d)).astype(np.float32)
    x = np.float32(1.0 + 2.0 ** -8 * (1.0 - 2.0 ** -10))      # worst-case one-hot rows / columns (both operands round down by ~half an ulp)
    for i in range(8):
        u[i] = 0
        u[i, i] = x
        W[i] = 0
        W[i, i] = x
    uh, Wh = _round_bf16(u), _round_bf16(W)
    exact = u.astype(np.float64) @ W.astype(np.float64).T + bias
    approx = (uh.astype(np.float64) @ Wh.astype(np.float64).T).astype(np.float32) + bias
    n = lambda a: np.linalg.norm(a.astype(np.float64), axis=1)
    bound = n(uh)[:, None] * n(W - Wh)[None, :] + (n(u - uh) + 2.0 ** -13 * n(u))[:, None] * n(W)[None, :]
    assert (np.abs(approx - exact) <= bound).all()
    tiles = Vc // TN
    tmax = approx.reshape(R, tiles, TN).max(-1)
    b = (n(uh)[:, None] * n(W - Wh).reshape(tiles, TN).max(-1)[None, :]
         + (n(u - uh) + 2.0 ** -13 * n(u))[:, None] * n(W).reshape(tiles, TN).max(-1)[None, :])
    keep = tmax + b >= (tmax - b).max(1, keepdims=True)
    assert keep[np.arange(R), exact.argmax(1) // TN].all()
    near = (exact >= exact.max(1, keepdims=True) - 1e-6).reshape(R, tiles, TN).any(-1)
    assert (keep | ~near).all()
    rel = 2.1 / 256 * n(u)[:, None] * n(W).reshape(tiles, TN).max(-1)[None, :]
    dense = slice(8, None)
    assert (b[dense, 1:] < 0.6 * rel[dense, 1:]).all()        # ~2x tighter than the relative bound on dense rows and tiles ...
    keep_rel = tmax + rel >= (tmax - rel).max(1, keepdims=True)
    assert keep[dense].sum() < 0.85 * keep_rel[dense].sum()   # ... which shows in the number of tiles handed to the refinement


def test_refine_butterfly_sums_do_not_depend_on_rows_per_warp():
    """csrc/vocab_refine.cu ``refine_rows``: a warp reduces NR x 16 per-lane partial sums with a halving exchange (lane bits 4, 3, 2, 1, 0
    in this order).  The claim the decode sharding rests on: the value a (row, column) pair ends up with is the same fp32 number whether
    the pair was refined by a CTA unit (NR = 4 rows per warp) or a warp unit (NR = 2) -- the same reduction tree over the 32 lanes'
    partials.  Emulated here in numpy with the kernel's exact order of fp32 additions."""
    rng = np.random.default_rng(5)
    TN = 16

    def butterfly(acc):                      # acc [32 lanes, NR * 16] float32 -> {(row, col): value}, as the kernel leaves them
        acc = acc.copy()
        n = acc.shape[1]
        half, bit = n // 2, 16
        while bit >= 1:
            new = acc.copy()
            for lane in range(32):
                up = (lane & bit) != 0
                for i in range(half):
                    keep = acc[lane, i + half] if up else acc[lane, i]
                    partner_up = ((lane ^ bit) & bit) != 0      # the partner sends acc[i] if IT is `up`, else acc[i + half]
                    recv = acc[lane ^ bit, i] if partner_up else acc[lane ^ bit, i + half]
                    new[lane, i] = np.float32(keep) + np.float32(recv)
            acc = new
            half //= 2
            bit //= 2
        nr = n // TN
        lpr, cpl = 32 // nr, TN // (32 // nr)
        out = {}
        for lane in range(32):
            m, c0 = lane // lpr, (lane % lpr) * cpl
            for e in range(cpl):
                out[(m, c0 + e)] = acc[lane, e]
        return out

    part4 = (rng.standard_normal((32, 4 * TN)) * 3).astype(np.float32)          # rows 0..3 of a CTA unit
    r4 = butterfly(part4)
    for pair_of_rows in ((0, 1), (2, 3), (1, 3)):
        part2 = np.concatenate([part4[:, m * TN:(m + 1) * TN] for m in pair_of_rows], axis=1)    # the same two rows as a warp unit
        r2 = butterfly(part2)
        for k, m in enumerate(pair_of_rows):
            for j in range(TN):
                assert r2[(k, j)].tobytes() == r4[(m, j)].tobytes(), (m, j)
    # and the tree really sums all 32 lanes
    ref = part4.astype(np.float64).sum(0)
    got = np.array([r4[(m, j)] for m in range(4) for j in range(TN)], dtype=np.float64)
    assert np.allclose(got, ref, rtol=1e-5, atol=1e-4)
