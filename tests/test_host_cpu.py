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
