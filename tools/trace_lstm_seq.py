#!/usr/bin/env python
"""Per-step timeline of the persistent LSTM kernels (CTA (0,0)), from the library's trace hook."""
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from adaptive_b200 import _lib  # noqa: E402
from adaptive_b200 import functional as F_aa  # noqa: E402
from adaptive_b200.synth import CFG_A, make_inputs, make_weights  # noqa: E402
from tests.gpu_utils import dev_inputs, dev_weights  # noqa: E402

B, T = 80, 18
w = make_weights(CFG_A)
inp = make_inputs(CFG_A, B, T)
W = dev_weights(w, requires_grad=True)
V, v_g, h0, c0, cap = dev_inputs(inp, requires_grad=True)
lib = _lib.load()
import os  # noqa: E402
if "AA_TRACE_NACC" in os.environ:   # accumulators of the forward cluster kernel's MMA chain (1, 2 or 4)
    _lib.check(lib.aa_debug_set_lstm_cluster(1, int(os.environ["AA_TRACE_NACC"])), "nacc")
names = ["bar_passed", "tma_issued", "stage0_landed", "mma_commit", "acc_seen", "cell_done", "warps_met", "published"]
# cluster kernels (lstm_cluster.cu, the default where H in {128,256,512}): 0 operand complete (MMA warp), 1 MMAs issued,
#  2 accumulator seen, 3 activations staged, 4 cell done (fwd) / dgates staged (bwd), 5 slice delivered / partials sent,
#  6 partials received (bwd), 7 outputs stored.  AA_LSTM_CLUSTER=0 in the environment traces the grid-barrier kernels.
for it in range(3):
    buf = torch.zeros(64 * 8, dtype=torch.int64, device="cuda")
    _lib.check(lib.aa_debug_set_trace_buffer(F_aa._ptr(buf)), "trace")
    scores = F_aa.decoder_forward(W, V, v_g, cap, h0, c0, precision="bf16")[0]
    torch.cuda.synchronize()
    fwd = buf.cpu().numpy().reshape(64, 8).copy()
    buf.zero_()
    scores.sum().backward()
    torch.cuda.synchronize()
    bwd = buf.cpu().numpy().reshape(64, 8).copy()
_lib.check(lib.aa_debug_set_trace_buffer(None), "trace")
for label, tr, rng in (("forward", fwd, range(0, T)), ("backward", bwd, range(0, T + 1))):
    print("==", label, "(ns relative to the step's first stamp; last column = step period)")
    print("step " + " ".join("%13s" % n for n in names))
    prev = None
    for s in rng:
        row = tr[s]
        nz = row[row > 0]
        if nz.size == 0:
            continue
        base = nz.min()
        rel = [(int(x - base) if x > 0 else -1) for x in row]
        per = int(base - prev) if prev is not None else 0
        prev = base
        print("%4d " % s + " ".join("%13d" % x for x in rel) + "   period=%d" % per)
    k = tr[40]
    if k[0] > 0:   # cluster kernels: whole-kernel stamps of CTA (0,0)
        print("kernel: entry 0, prologue done +%d ns, loop done +%d ns, exit +%d ns" % (k[1] - k[0], k[2] - k[0], k[3] - k[0]))
