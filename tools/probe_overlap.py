"""What slows the peer-memory all-reduce down when it runs next to compute -- and the compute next to it?

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 tools/probe_overlap.py

A 10 M-float (40 MB fp32 / 20 MB bf16-on-the-wire) all-reduce is timed alone and concurrently with three kinds of
co-runners on a second stream, each also timed alone and next to the all-reduce:
  alu     a kernel of 148 x 2 CTAs spinning on register arithmetic (issue slots only; torch elementwise on a tiny tensor, many iterations)
  hbm     a device-to-device copy of 1 GB (HBM + L2 bandwidth)
  gemm    the library's tcgen05 bf16 contraction [1440 x 10000 x 512] in a loop (TMA + tensor pipe + mbarrier spins, persistent CTAs)
One JSON line per row on rank 0 (times in microseconds, max over ranks)."""
import ctypes
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from adaptive_b200 import _lib  # noqa: E402
from adaptive_b200.parallel import SymmetricBuffer  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
    dist.init_process_group("nccl", device_id=dev)
    lib = _lib.load()
    n = 10 * (1 << 20)
    sb = SymmetricBuffer(n, dev)
    sb.max_blocks = int(os.environ.get("AA_AR_BLOCKS", "16"))
    view = sb.payload[:n]
    view.normal_()
    side = torch.cuda.Stream()
    main_s = torch.cuda.Stream(priority=int(os.environ.get('AA_DP_COMM_PRIORITY', '-100')))       # the all-reduce's lane
    torch.cuda.set_stream(main_s)

    # co-runners
    M, N, K = 1440, 10000, 512
    A = torch.randn(M, K, device=dev).bfloat16()
    Bm = torch.randn(N, K, device=dev).bfloat16()
    D = torch.empty(M, N, device=dev)
    src = torch.empty(256 << 20, dtype=torch.float32, device=dev)
    dst = torch.empty_like(src)
    small = torch.randn(148 * 2 * 256, device=dev)

    def gemm(st, reps=12):
        for _ in range(reps):
            _lib.check(lib.aa_gemm(1, M, N, K, ctypes.c_void_p(A.data_ptr()), K, 1, ctypes.c_void_p(Bm.data_ptr()), K, 1, None, 0, 0.0, None,
                                   ctypes.c_void_p(D.data_ptr()), N, ctypes.c_void_p(st.cuda_stream)), "aa_gemm")

    def hbm(st):
        with torch.cuda.stream(st):
            dst.copy_(src)

    def alu(st):
        with torch.cuda.stream(st):
            x = small
            for _ in range(40):
                x = torch.sin(x)

    def ar(st, bf16):
        sb.all_reduce_(view, channel=1, stream=st, bf16=bf16)

    def timed(fn_main, fn_side=None, reps=5):
        """-> (us of fn_main on the main stream, us of fn_side on the side stream), both started together"""
        outs = []
        for _ in range(reps + 1):
            dist.barrier()
            torch.cuda.synchronize()
            e = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
            side.wait_stream(main_s)
            e[0].record(main_s)
            if fn_side is not None:
                e[2].record(side)
                fn_side(side)
                e[3].record(side)
            fn_main(main_s)
            e[1].record(main_s)
            torch.cuda.synchronize()
            outs.append((e[0].elapsed_time(e[1]) * 1e3, e[2].elapsed_time(e[3]) * 1e3 if fn_side is not None else 0.0))
        outs = outs[1:]
        t = torch.tensor([sorted(o[0] for o in outs)[len(outs) // 2], sorted(o[1] for o in outs)[len(outs) // 2]], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return [round(float(x), 1) for x in t]

    rows = {"world": world, "ar_blocks": sb.max_blocks, "threads": os.environ.get("AA_AR_THREADS", "256"),
            "ar_stream_priority": main_s.priority}
    for bf16 in (False, True):
        tag = "bf16" if bf16 else "fp32"
        rows["ar_%s_alone" % tag] = timed(lambda st: ar(st, bf16))[0]
        for name, co in (("alu", alu), ("hbm", hbm), ("gemm", gemm)):
            if not bf16:
                rows[name + "_alone"] = timed(co)[0]
            a, c = timed(lambda st: ar(st, bf16), co)
            rows["ar_%s_with_%s" % (tag, name)] = a
            rows["%s_with_ar_%s" % (name, tag)] = c
    if rank == 0:
        print(json.dumps(rows), flush=True)
    dist.barrier()
    torch.cuda.synchronize()
    os._exit(0)


if __name__ == "__main__":
    main()
