#!/usr/bin/env python
"""Debug aid: forward cluster recurrence against the grid-barrier kernel, per step count (hT, cT only)."""
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from adaptive_b200 import _lib  # noqa: E402
from adaptive_b200 import functional as F_aa  # noqa: E402
from adaptive_b200.synth import CFG_A, make_inputs, make_weights  # noqa: E402
from tests.gpu_utils import dev_inputs, dev_weights  # noqa: E402

lib = _lib.load()
w = make_weights(CFG_A, seed=61, bias_scale=0.1)
for B, T in ((80, 1), (80, 2), (80, 3), (80, 18), (5, 4)):
    inp = make_inputs(CFG_A, B, T, seed=62)
    out = {}
    for mode in (1, 0):
        lib.aa_debug_set_lstm_cluster(mode, 0)
        W = dev_weights(w, requires_grad=False)
        V, v_g, h0, c0, cap = dev_inputs(inp, requires_grad=False)
        with torch.no_grad():
            scores, alpha, beta, hT, cT = F_aa.decoder_forward(W, V, v_g, cap, h0, c0, precision="bf16")
        torch.cuda.synchronize()
        out[mode] = (hT.cpu().numpy(), cT.cpu().numpy(), scores.cpu().numpy())
    lib.aa_debug_set_lstm_cluster(1, 0)
    h1, c1, s1 = out[1]
    h0_, c0_, s0 = out[0]
    bad = ~np.isfinite(h1)
    print("B=%d T=%d  nonfinite hT: %d  cT: %d  scores: %d (old path: %d)" % (B, T, bad.sum(), (~np.isfinite(c1)).sum(), (~np.isfinite(s1)).sum(), (~np.isfinite(s0)).sum()))
    if bad.any():
        rows, cols = np.where(bad)
        print("   rows", np.unique(rows)[:20], "units", np.unique(cols)[:40])
    d = np.abs(np.nan_to_num(h1) - h0_)
    print("   max |hT diff| %.3e (max |hT| %.3e)   per 32-unit slice: %s" % (d.max(), np.abs(h0_).max(), np.array2string(d.reshape(B, 16, 32).max(axis=(0, 2)), precision=3)))
    print("   per row:", np.array2string(d.max(axis=1)[:20], precision=3))
