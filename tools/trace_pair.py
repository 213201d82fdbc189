#!/usr/bin/env python
"""Where the CTA-pair contraction spends a tile: cycles the MMA warp waits for a free accumulator / for operands, cycles it issues, and
what one epilogue warp waits and works -- from a development build of the library (gemm_tc.cu compiled with -DAA_PAIR_TRACE):

    make -C adaptive_b200/csrc && tools/build_pair_trace.sh && python tools/trace_pair.py [vocab|gate]

vocab: the last traced launch of a greedy decode is the bf16 maxima pass; gate: with the arg-max refinement off it is the gate GEMM."""
import ctypes
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from adaptive_b200 import _lib  # noqa: E402

_lib.LIB_PATH = os.path.join(ROOT, "adaptive_b200", "csrc", "build", "libadaptive_trace.so")
from adaptive_b200 import functional as F_aa  # noqa: E402
from adaptive_b200.synth import CFG_A, make_inputs, make_weights  # noqa: E402
from tests.gpu_utils import dev_inputs, dev_weights  # noqa: E402


def main():
    kind = sys.argv[1] if len(sys.argv) > 1 else "vocab"
    lib = _lib.load()
    if kind == "gate":
        lib.aa_debug_set_decode_argmax_refine(0)
    W = dev_weights(make_weights(CFG_A, seed=123))
    V, v_g, h0, c0, _ = dev_inputs(make_inputs(CFG_A, 4096, 1, seed=1234))
    for _ in range(2):
        F_aa.greedy_decode(W, V, v_g, h0, c0, 3, engine="pipeline")
    torch.cuda.synchronize()
    n = 148 * 16 * 8
    buf = (ctypes.c_longlong * n)()
    fn = lib.aa_debug_pair_trace
    fn.restype = ctypes.c_int
    assert fn(buf, n, 1) == 0
    t = np.frombuffer(buf, dtype=np.int64).reshape(148, 16, 8)
    lead = t[0::2]                       # leader CTAs hold the MMA warp's records
    print(kind, "tiles per pair:", (lead[:, :, 2] > 0).sum(1).min(), "..", (lead[:, :, 2] > 0).sum(1).max())
    print("  lt | MMA warp: wait free accumulator, wait operands, whole tile | epilogue warp 2 (leader CTA): wait, work   [cycles, mean over pairs (max)]")
    for lt in range(16):
        m = lead[:, lt, 2] > 0
        if not m.any():
            break
        f = lambda a: "%7.0f (%6d)" % (a[m].mean(), a[m].max())
        print("  %2d | %s %s %s | %s %s   pairs %d" % (lt, f(lead[:, lt, 0]), f(lead[:, lt, 1]), f(lead[:, lt, 2]), f(lead[:, lt, 4]), f(lead[:, lt, 5]),
                                                      int(m.sum())))
    g = t[:, :4, 7].astype(np.float64)      # globaltimer (ns): entry, after set-up, loops done, after the final cluster barrier + dealloc
    t0 = g[:, 0].min()
    print("  kernel (globaltimer, us from the first CTA's entry): entry max %.1f | set-up done mean %.1f max %.1f | loops done mean %.1f max %.1f | exit max %.1f"
          % ((g[:, 0].max() - t0) / 1e3, (g[:, 1].mean() - t0) / 1e3, (g[:, 1].max() - t0) / 1e3, (g[:, 2].mean() - t0) / 1e3, (g[:, 2].max() - t0) / 1e3,
             (g[:, 3].max() - t0) / 1e3))
    peer = t[1::2]
    m = peer[:, 0, 5] > 0
    print("  peer CTA epilogue work, tile 0: %.0f mean" % peer[m][:, 0, 5].mean())


if __name__ == "__main__":
    main()
