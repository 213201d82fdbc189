#!/usr/bin/env python
"""In-situ kernel timeline of the graphed training step (BASELINE config 2) from torch.profiler / CUPTI:
start offset, duration, stream and name of every kernel and memset of ONE graph replay, in start order.  Unlike the ncu launch
list (cold cache, serialised) these are the times inside the replayed graph, with the lanes running next to each other.
   python tools/timeline_train.py > gpurun_out/timeline.txt"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import adaptive_b200  # noqa: E402
from adaptive_b200 import functional as F_aa  # noqa: E402
from adaptive_b200.graphs import GraphedTrainStep  # noqa: E402
from adaptive_b200.synth import CFG_A, make_inputs, make_lengths, make_weights  # noqa: E402

B, T = 80, 18
dims = CFG_A
world = int(os.environ.get("WORLD_SIZE", "1"))      # under torchrun: the data-parallel step (rank 0 prints)
rank = int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    import torch.distributed as dist
    os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
    dist.init_process_group("nccl", device_id=dev)


class Cf:
    adaptive_word_embed_size, adaptive_lstm_hidden_size, vocab_length = dims.E, dims.H, dims.Vc


model = adaptive_b200.Encoder2Decoder(Cf()).to(dev)
model.decoder.precision = "bf16"
w = make_weights(dims, seed=123)
model.load_state_dict({"decoder." + k: torch.from_numpy(v) for k, v in w.items()}, strict=False)
lengths = make_lengths(B, T, seed=1234)
inp = make_inputs(dims, B, T, seed=1234)
b = {k: torch.from_numpy(v).to(dev) for k, v in inp.items()}
b["tgt"] = torch.from_numpy(np.ascontiguousarray(F_aa.packed_targets(inp["captions"], lengths))).to(dev)
if world > 1 or os.environ.get('AA_TIMELINE_DP_PATH') == '1':
    from adaptive_b200.parallel import DataParallelTrainer, GraphedDPStep
    step = GraphedDPStep(DataParallelTrainer(model, overlap=True), b, lengths)
else:
    step = GraphedTrainStep(model, b, lengths)
for _ in range(10):
    step(b)
torch.cuda.synchronize()
from torch.profiler import ProfilerActivity, profile  # noqa: E402

with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(3):
        step(b)
    torch.cuda.synchronize()
if rank != 0:
    torch.cuda.synchronize()
    os._exit(0)
evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
evs.sort(key=lambda e: e.time_range.start)
# split into replays at gaps > 20 us
groups, cur, last_end = [], [], None
for e in evs:
    if last_end is not None and e.time_range.start - last_end > 20 and cur:
        groups.append(cur)
        cur = []
    cur.append(e)
    last_end = max(last_end or 0, e.time_range.end)
if cur:
    groups.append(cur)
g = max(groups, key=len) if groups else []
g = groups[-1] if groups and len(groups[-1]) >= 0.9 * len(max(groups, key=len)) else g
t0 = g[0].time_range.start if g else 0
streams = {}
print("# one replay: %d device activities, span %.1f us" % (len(g), (max(e.time_range.end for e in g) - t0) if g else 0))
print("# %9s %9s %9s  lane  name" % ("start_us", "dur_us", "end_us"))
for e in g:
    sid = getattr(e, "stream", None)
    if sid is None:
        sid = getattr(e, "device_resource_id", -1)
    lane = streams.setdefault(sid, len(streams))
    print("%11.1f %9.1f %9.1f  %4d  %s" % (e.time_range.start - t0, e.time_range.end - e.time_range.start, e.time_range.end - t0, lane, e.name[:110]))

sys.stdout.flush()
if world > 1:
    os._exit(0)
