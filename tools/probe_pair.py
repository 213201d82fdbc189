#!/usr/bin/env python
"""Single-CTA against CTA-pair (tcgen05.mma.cta_group::2) 3xTF32 contraction: us per launch of aa_gemm_split3 at the decode step's gate
shape and at longer reductions (steady-state rate per k-block).  Run once with AA_GEMM_PAIR=0 and once with AA_GEMM_PAIR=1."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from adaptive_b200 import _lib  # noqa: E402
from adaptive_b200.functional import _ptr, _stream  # noqa: E402


def main():
    lib = _lib.load()
    for (M, N, K) in ((4096, 2048, 512), (4096, 2048, 2048), (4096, 2048, 8192), (4096, 10240, 512), (256, 256, 8192), (256, 256, 512)):
        A = torch.randn(M, K, device="cuda")
        B = torch.randn(N, K, device="cuda")
        As, Bs = torch.empty(M, 2 * K, device="cuda"), torch.empty(N, 2 * K, device="cuda")
        D = torch.empty(M, N, device="cuda")
        st = _stream(D.device)
        _lib.check(lib.aa_split_tf32(_ptr(A), K, M, K, _ptr(As), K, st), "split")
        _lib.check(lib.aa_split_tf32(_ptr(B), K, N, K, _ptr(Bs), K, st), "split")
        for _ in range(3):
            _lib.check(lib.aa_gemm_split3(M, N, K, _ptr(As), _ptr(Bs), None, _ptr(D), N, st), "gemm")
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 20
        e0.record()
        for _ in range(reps):
            _lib.check(lib.aa_gemm_split3(M, N, K, _ptr(As), _ptr(Bs), None, _ptr(D), N, st), "gemm")
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 1e3 / reps
        print(json.dumps({"pair": os.environ.get("AA_GEMM_PAIR", "1"), "M": M, "N": N, "K": K, "us": round(us, 1),
                          "tf32_tflops_x3": round(3 * 2.0 * M * N * K / us / 1e6, 1), "us_per_kblock32": round(us / (K / 32), 3)}))


if __name__ == "__main__":
    main()
