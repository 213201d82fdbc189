#!/usr/bin/env python
"""Time of one greedy decode call split into prologue-dominated (L = 1) and per-step parts: python tools/prof_prologue.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from adaptive_b200 import functional as F_aa  # noqa: E402
from adaptive_b200.synth import CFG_A, make_inputs, make_weights  # noqa: E402
from tests.gpu_utils import dev_inputs, dev_weights  # noqa: E402

W = dev_weights(make_weights(CFG_A, seed=123))
V, v_g, h0, c0, _ = dev_inputs(make_inputs(CFG_A, 4096, 1, seed=1234))
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for L in (1, 20):
    for _ in range(2):
        F_aa.greedy_decode(W, V, v_g, h0, c0, L)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(5):
        F_aa.greedy_decode(W, V, v_g, h0, c0, L)
    e1.record()
    torch.cuda.synchronize()
    print("L=%d: %.3f ms per call" % (L, e0.elapsed_time(e1) / 5))
