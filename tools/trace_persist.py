"""Per-step timeline of the persistent decoder (CTA 0) + cost of its per-call prologue.
    python tools/trace_persist.py [B]"""
import ctypes
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from adaptive_b200 import _lib  # noqa: E402
from adaptive_b200 import functional as F_aa  # noqa: E402
from adaptive_b200.synth import CFG_A, make_inputs, make_weights  # noqa: E402
from tests.gpu_utils import dev_inputs, dev_weights  # noqa: E402

NAMES = ["phase A entered", "own units done", "barrier 1 passed", "arg-max done", "owner step done", "barrier 2 passed"]


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 148
    L = 20
    W = dev_weights(make_weights(CFG_A, seed=123))
    V, v_g, h0, c0, _ = dev_inputs(make_inputs(CFG_A, B, 1, seed=5))
    lib = _lib.load()
    for _ in range(3):
        F_aa.greedy_decode_persistent(W, V, v_g, h0, c0, L)
    buf = torch.zeros(L * 8, dtype=torch.int64, device="cuda")
    lib.aa_debug_set_persist_trace(ctypes.c_void_p(buf.data_ptr()))
    F_aa.greedy_decode_persistent(W, V, v_g, h0, c0, L)
    torch.cuda.synchronize()
    lib.aa_debug_set_persist_trace(None)
    t = buf.cpu().numpy().reshape(L, 8).astype(np.float64)
    print("B = %d; microseconds per step, CTA 0 (globaltimer)" % B)
    print("step " + " ".join("%18s" % n for n in ["A->units", "units->bar1", "bar1->argmax", "argmax->owner", "owner->bar2", "total"]))
    rows = []
    for s in range(L - 1):
        d = [(t[s, i + 1] - t[s, i]) / 1e3 for i in range(5)]
        tot = (t[s + 1, 0] - t[s, 0]) / 1e3
        rows.append(d + [tot])
        print("%4d " % s + " ".join("%18.2f" % x for x in d + [tot]))
    print("mean " + " ".join("%18.2f" % x for x in np.mean(np.array(rows)[2:], axis=0)))
    # whole call vs kernel: the per-call prologue (P, static gate terms, initial state)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        F_aa.greedy_decode_persistent(W, V, v_g, h0, c0, L)
    e1.record()
    torch.cuda.synchronize()
    print("whole call: %.1f us; kernel loop (20 steps, trace): %.1f us" % (e0.elapsed_time(e1) / 20 * 1e3, (t[L - 1, 5] - t[0, 0]) / 1e3))


if __name__ == "__main__":
    main()
