#!/usr/bin/env python
"""Small driver for profilers: BASELINE config 3 greedy decode (B=4096, cfgA) for a few steps.
   ncu --set full -k regex:dec_atten_tma -s 4 -c 2 python tools/prof_decode.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from adaptive_b200 import functional as F_aa  # noqa: E402
from adaptive_b200.synth import CFG_A, make_inputs, make_weights  # noqa: E402
from tests.gpu_utils import dev_inputs, dev_weights  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
L = int(sys.argv[2]) if len(sys.argv) > 2 else 4
beam = int(sys.argv[3]) if len(sys.argv) > 3 else 1
W = dev_weights(make_weights(CFG_A, seed=123))
V, v_g, h0, c0, _ = dev_inputs(make_inputs(CFG_A, B, 1, seed=1234))
for _ in range(2):
    if beam == 1:
        out = F_aa.greedy_decode(W, V, v_g, h0, c0, L)
    else:
        out = F_aa.beam_decode(W, V, v_g, h0, c0, beam, L)
torch.cuda.synchronize()
print("ok", out[0].shape)
