#!/usr/bin/env python
"""Micro-benchmark of the contraction engines through aa_gemm (CUDA events, L2 flushed between
launches).  Prints one JSON line per (engine, shape, layout)."""
import json
import sys

import torch

sys.path.insert(0, ".")
from adaptive_b200 import _lib  # noqa: E402
from adaptive_b200.functional import _ptr, _stream  # noqa: E402

SHAPES = {  # name: (M, N, K, a_kmajor, b_kmajor, with_c)
    "vocab_fwd": (1440, 10000, 512, 1, 1, 0), "vocab_dx": (1440, 512, 10000, 1, 0, 0), "vocab_dw": (10000, 512, 1440, 0, 0, 0),
    "lstm_rec": (80, 2048, 512, 1, 1, 1), "bptt_rec": (80, 512, 2048, 1, 0, 0), "gates_in": (1440, 2048, 512, 1, 1, 0),
    "dec_vocab": (4096, 10000, 512, 1, 1, 0), "dec_gate": (4096, 2560, 768, 1, 1, 1), "cfgB_vocab": (4608, 20000, 1024, 1, 1, 0),
}


def main():
    lib = _lib.load()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    for name, (M, N, K, ak, bk, wc) in SHAPES.items():
        for engine in (1, 2, 0):
            if engine == 2 and not (ak and bk):
                continue
            dt = torch.bfloat16 if engine == 1 else torch.float32
            A = torch.randn((M, K) if ak else (K, M), device="cuda").to(dt)
            B = torch.randn((N, K) if bk else (K, N), device="cuda").to(dt)
            C = torch.randn(M, N, device="cuda") if wc else None
            D = torch.empty(M, N, device="cuda")
            ts = []
            for i in range(8):
                flush.zero_()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                _lib.check(lib.aa_gemm(engine, M, N, K, _ptr(A), A.stride(0), ak, _ptr(B), B.stride(0), bk, _ptr(C), N, 1.0, None, _ptr(D), N,
                                       _stream(D.device)), "aa_gemm")
                e1.record()
                torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1))
            t = sorted(ts[2:])[len(ts[2:]) // 2] * 1e-3
            print(json.dumps({"shape": name, "M": M, "N": N, "K": K, "engine": ["fp32_simt", "tc_bf16", "tc_tf32"][engine], "us": t * 1e6,
                              "tflops": 2.0 * M * N * K / t / 1e12}))
        if ak and bk and not wc:          # fp32-accurate 3xTF32 over pre-split operands (the decode engine)
            Kp = (K + 31) // 32 * 32
            A = torch.randn(M, K, device="cuda")
            B = torch.randn(N, K, device="cuda")
            As, Bs = torch.empty(M, 2 * Kp, device="cuda"), torch.empty(N, 2 * Kp, device="cuda")
            D = torch.empty(M, N, device="cuda")
            st = _stream(D.device)
            _lib.check(lib.aa_split_tf32(_ptr(A), K, M, K, _ptr(As), Kp, st), "split")
            _lib.check(lib.aa_split_tf32(_ptr(B), K, N, K, _ptr(Bs), Kp, st), "split")
            ts = []
            for i in range(8):
                flush.zero_()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                _lib.check(lib.aa_gemm_split3(M, N, Kp, _ptr(As), _ptr(Bs), None, _ptr(D), N, st), "aa_gemm_split3")
                e1.record()
                torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1))
            t = sorted(ts[2:])[len(ts[2:]) // 2] * 1e-3
            print(json.dumps({"shape": name, "M": M, "N": N, "K": K, "engine": "tc_tf32x3", "us": t * 1e6,
                              "tflops": 2.0 * M * N * K / t / 1e12, "tensor_tflops": 6.0 * M * N * Kp / t / 1e12}))


if __name__ == "__main__":
    main()
