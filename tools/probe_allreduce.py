"""Correctness + latency probe of the peer-memory all-reduce kernels (csrc/allreduce.cu) against NCCL.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 tools/probe_allreduce.py

Per size: max |ours - nccl| (inputs are integers scaled by 1/8: sums are exact in fp32, so the difference must be 0), then the
time of one collective, eager and replayed from a CUDA graph, max over ranks, for both engines (multimem when the fabric offers
a multicast mapping, two-shot peer loads, NCCL).  One JSON line per row on rank 0."""
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from adaptive_b200.parallel import SymmetricBuffer  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
    dist.init_process_group("nccl", device_id=dev)
    sizes = [1024, 65536, 1 << 20, 4651264, 5 << 20]          # floats: 4 KB ... 18.6 MB (the tail bucket) ... 20 MB
    sb = SymmetricBuffer(max(sizes), dev)
    if rank == 0:
        print(json.dumps({"world": world, "multicast": bool(sb.multicast_ptr), "peer_ptrs": len(sb.peer_ptrs)}), flush=True)
    modes = [("multimem" if sb.multicast_ptr else "twoshot", sb.multicast_ptr)]
    if sb.multicast_ptr:
        modes.append(("twoshot", 0))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    def timed(fn, n=20):
        fn()
        dist.barrier()
        torch.cuda.synchronize()
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / n], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()) * 1e3

    for n in sizes:
        g = torch.Generator(device="cpu").manual_seed(1000 * rank + 7)
        src = (torch.randint(-64, 64, (n,), generator=g).float() / 8).to(dev)
        ref = src.clone()
        dist.all_reduce(ref)
        row = {"floats": n, "bytes": 4 * n}
        view = sb.payload[:n]
        for name, mc in modes:
            saved = sb.multicast_ptr
            sb.multicast_ptr = mc
            view.copy_(src)
            torch.cuda.synchronize()
            dist.barrier()
            sb.all_reduce_(view, channel=1)
            torch.cuda.synchronize()
            row[name + "_max_abs_err"] = float((view - ref).abs().max())
            for blocks in (8, 16, 32, 64):
                sb.max_blocks = blocks
                row["%s_us_eager_b%d" % (name, blocks)] = timed(lambda: sb.all_reduce_(view, channel=1))
            sb.max_blocks = 32
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                sb.all_reduce_(view, channel=2)
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            gr = torch.cuda.CUDAGraph()
            with torch.cuda.graph(gr):
                sb.all_reduce_(view, channel=2)
            row[name + "_us_graph"] = timed(gr.replay)
            sb.multicast_ptr = saved
        buf = src.clone()
        row["nccl_us_eager"] = timed(lambda: dist.all_reduce(buf))
        try:
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                dist.all_reduce(buf)
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            gr = torch.cuda.CUDAGraph()
            with torch.cuda.graph(gr):
                dist.all_reduce(buf)
            row["nccl_us_graph"] = timed(gr.replay)
        except Exception as e:
            row["nccl_us_graph"] = "failed: %s" % type(e).__name__
        if rank == 0:
            print(json.dumps(row), flush=True)
    dist.barrier()
    torch.cuda.synchronize()
    os._exit(0)


if __name__ == "__main__":
    main()
