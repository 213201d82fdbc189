"""How many 64-column tiles can hold a row's arg-max as a function of the error bound (sizing of vocab_refine.cu)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from adaptive_b200 import functional as F_aa
from adaptive_b200.synth import CFG_A, make_inputs, make_weights
from tests.gpu_utils import dev_inputs, dev_weights

dims, B, L = CFG_A, 4096, 3
w = make_weights(dims, seed=123)
inp = make_inputs(dims, B, 1, seed=4321)
W = dev_weights(w)
V, v_g, h0, c0, _ = dev_inputs(inp)
ids, att, bet, logits = F_aa.greedy_decode(W, V, v_g, h0, c0, L, return_logits=True)
wn = float(np.linalg.norm(w["adaptive.mlp.weight"], axis=1).max())
for t in range(L):
    lg = logits[t]
    sd = float(lg.std(dim=1).mean())
    un = sd / 0.0625
    pad = torch.full((B, 157 * 64 - dims.Vc), -1e30, device="cuda")
    tm = torch.cat([lg, pad], 1).view(B, 157, 64).max(-1).values
    rm = tm.max(-1, keepdim=True).values
    print("step", t, "logit std %.3f  est ||u|| %.1f  max||w|| %.2f  bound(tf32) %.4f" % (sd, un, wn, 1.1 / 1024 * un * wn))
    for d in (0.01, 0.03, 0.05, 0.1, 0.2, 0.4):
        print("   tiles within %.2f of the row max: %.2f per row" % (d, float((tm >= rm - d).float().sum(1).mean())))

from adaptive_b200 import _lib
lib = _lib.load()
lib.aa_debug_refine_pairs(1)
F_aa.greedy_decode(W, V, v_g, h0, c0, 20)
print("refined (row, tile) pairs per row and step: %.3f" % (lib.aa_debug_refine_pairs(1) / (B * 20.0)))
