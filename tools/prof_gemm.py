#!/usr/bin/env python
"""Profiler driver: the training step's largest bf16 contractions through aa_gemm, one launch each after a warm-up launch
(odd launches are the ones to read).
   ncu --set full -k regex:gemm_tc_kernel --import-source on -o gpurun_out/gemm python tools/prof_gemm.py"""
import sys

import torch

sys.path.insert(0, ".")
from adaptive_b200 import _lib  # noqa: E402
from adaptive_b200.functional import _ptr, _stream  # noqa: E402

SHAPES = {  # name: (M, N, K, a_kmajor, b_kmajor)
    "vocab_fwd": (840, 10000, 512, 1, 1), "vocab_dx": (840, 512, 10000, 1, 0), "vocab_dw": (10000, 512, 840, 0, 0),
    "gates_in": (1440, 2048, 512, 1, 1), "lstm_dw": (2048, 512, 1440, 0, 0), "lstm_dx": (1440, 512, 2048, 1, 0),
}
lib = _lib.load()
names = sys.argv[1:] or list(SHAPES)
for name in names:
    M, N, K, ak, bk = SHAPES[name]
    A = torch.randn((M, K) if ak else (K, M), device="cuda").to(torch.bfloat16)
    B = torch.randn((N, K) if bk else (K, N), device="cuda").to(torch.bfloat16)
    D = torch.zeros(M, N, device="cuda")
    for i in range(2):
        _lib.check(lib.aa_gemm(1, M, N, K, _ptr(A), A.stride(0), ak, _ptr(B), B.stride(0), bk, None, N, 1.0, None, _ptr(D), N,
                               _stream(D.device)), "aa_gemm")
    torch.cuda.synchronize()
    print(name, "ok")
