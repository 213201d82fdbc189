#!/usr/bin/env python
"""Small driver for profilers: the fused forward + loss operator (aa_decoder_forward_loss) at BASELINE config 5 shapes, two calls.
   ncu --set full -k regex:"ce_fixup|ce_merge|gemm_tc_kernel<128, 2, 5, 0, 0, 0, 3" python tools/prof_fused_loss.py"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from adaptive_b200 import functional as F_aa  # noqa: E402
from adaptive_b200.synth import CFG_A, CFG_B, make_lengths  # noqa: E402

small = len(sys.argv) > 1 and sys.argv[1] == "cfgA"
dims, B, T = (CFG_A, 80, 18) if small else (CFG_B, 256, 18)
g = torch.Generator(device="cpu").manual_seed(5)
sh = dims.shapes()
order = ["embed.weight", "LSTM.weight_ih_l0", "LSTM.weight_hh_l0", "LSTM.bias_ih_l0", "LSTM.bias_hh_l0", "adaptive.sentinel.affine_x.weight",
         "adaptive.sentinel.affine_h.weight", "adaptive.atten.affine_v.weight", "adaptive.atten.affine_g.weight", "adaptive.atten.affine_s.weight",
         "adaptive.atten.affine_h.weight", "adaptive.mlp.weight", "adaptive.mlp.bias"]
W = tuple((torch.randn(sh[k], generator=g) / (sh[k][-1] ** 0.5)).cuda() for k in order)
V = torch.relu(torch.randn(B, dims.k, dims.H, generator=g)).cuda()
v_g = torch.relu(torch.randn(B, dims.E, generator=g)).cuda()
h0 = torch.tanh(torch.randn(B, dims.H, generator=g)).cuda()
c0 = torch.tanh(torch.randn(B, dims.H, generator=g)).cuda()
cap = torch.randint(4, dims.Vc, (B, T), generator=g).cuda()
cap[:, 0] = 1
lengths = make_lengths(B, T, seed=1234)
for _ in range(2):
    loss = F_aa.decoder_forward_loss(W, V, v_g, cap, lengths, None, h0, c0)[0]
torch.cuda.synchronize()
print("ok", float(loss))
