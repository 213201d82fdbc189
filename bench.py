#!/usr/bin/env python
"""Benchmark of the adaptive-attention decoder hot path (BASELINE.json metric:
decoder tokens/sec, teacher-forced train fwd+bwd and greedy decode).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Headline workload (N=1): BASELINE config 2 — one training step (forward through
``Encoder2Decoder.forward`` -> packed scores -> mean cross-entropy -> backward) at batch 80,
49 regions, hidden 512, vocab 10k, caption length 18, synthetic features and random-init
weights.  BASELINE config 3 (greedy sampler, batch 4096, max_len 20) is measured in the same
run and reported under ``"decode"``.  With N > 1 (torchrun) every rank keeps the same
per-GPU batch (weak scaling); training adds an NCCL all-reduce of the decoder gradients,
decoding shards images with no communication.

``--impl reference`` times the oracle port of the reference's CPU path (numpy, all host
threads) on the same workload, on rank 0 only.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from adaptive_b200.synth import CFG_A, make_inputs, make_lengths, make_weights  # noqa: E402

TRAIN_B, TRAIN_T = 80, 18          # BASELINE config 2
DECODE_B, DECODE_L = 4096, 20      # BASELINE config 3
METRIC = "decoder tokens/sec (train fwd+bwd)"


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"hbm_gbs": d["hbm_gbs"], "bf16_tflops": d["bf16_tflops"], "bf16_tflops_sustained": d.get("bf16_tflops_sustained"),
                "source": "measured (MEASURED_PEAKS.json)"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback (B200_PROFILING.md)"}


# ---------------------------------------------------------------------------------------------
# CPU arm: the oracle port of the reference path (numpy, multi-threaded BLAS)
# ---------------------------------------------------------------------------------------------
def oracle_train_step_fn(B, T, dims):
    from adaptive_b200.functional import packed_row_index
    from oracle import adaptive_oracle as orc

    w = make_weights(dims, seed=123)
    inp = make_inputs(dims, B, T, seed=1234)
    lengths = make_lengths(B, T, seed=1234)
    idx, _ = packed_row_index(lengths, T)
    tgt = orc.packed_targets(inp["captions"], lengths)

    def step():
        s, _, _, _, cache = orc.decoder_forward(w, inp["V"], inp["v_g"], inp["captions"], inp["h0"], inp["c0"], want_cache=True)
        data, _ = orc.pack_padded(s, lengths)
        loss, dlog = orc.cross_entropy(data, tgt)
        dS = np.zeros((B * T, dims.Vc), dtype=np.float32)
        dS[idx] = dlog
        orc.decoder_backward(w, cache, dS.reshape(B, T, dims.Vc))
        return float(loss)

    return step


def time_cpu(step, steps, warmup):
    """Seconds per step with the BLAS pool at every host core (torchrun exports OMP_NUM_THREADS=1 to its workers, which
    would otherwise pin the numpy port to one thread)."""
    from threadpoolctl import threadpool_limits

    with threadpool_limits(limits=os.cpu_count() or 1):
        for _ in range(warmup):
            step()
        t0 = time.perf_counter()
        for _ in range(steps):
            step()
        return (time.perf_counter() - t0) / steps


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    step = oracle_train_step_fn(TRAIN_B, TRAIN_T, CFG_A)
    steps = max(1, min(args.steps, 40))
    sec = time_cpu(step, steps, max(1, min(args.warmup, 3)))
    v = TRAIN_B * TRAIN_T / sec
    out = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": "tokens/s", "n_gpus": args.gpus, "steps": steps,
        "warmup": max(1, min(args.warmup, 3)), "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args.gpus),
        "cpu_baseline": {"value": v, "unit": "tokens/s", "cores": cores, "kind": "port",
                         "sample": "%d full steps of the same workload (B=%d, T=%d) on the numpy oracle port" % (steps, TRAIN_B, TRAIN_T)},
        "e2e": {"value": v, "unit": "tokens/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    emit(out)


def workload_config(n):
    return {"workload": "BASELINE config 2: train step fwd+CE+bwd, batch %d per GPU, 49 regions, hidden 512, embed 256, vocab 10000, "
                        "caption len 18 (variable lengths, mean 10.5 words)" % TRAIN_B,
            "global_batch": TRAIN_B * n, "seq_len": TRAIN_T, "parallelism": "dp%d" % n,
            "l2": "per-step working set (weights+grads 83 MB, logits+dlogits ~110 MB, saved activations ~56 MB) exceeds the 126 MB L2; "
                  "4 rotating input batches"}


# ---------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clock / throttle sampling during the timed region (B200_PROFILING.md recipe)."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx = float(f[1])
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


def kernel_models(dims, B, T, decode_B):
    """Algorithmic work per launch of each profiled kernel tag (DESIGN.md section 'Kernels')."""
    H, E, k, a, Vc = dims.H, dims.E, dims.k, dims.a, dims.Vc
    N = B * T
    fl = lambda m, n, kk: 2.0 * m * n * kk
    # attention stream kernel of the tensor-core decode pipeline, per (row, step): reads V, P, [h|s], [q|r]; writes u, alpha, beta
    step_bytes = k * H * 4 + k * a * 4 + 2 * H * 4 + 2 * a * 4 + H * 4 + k * 4 + 4          # 116 692 B at cfgA
    cell_bytes = (5 * H + H) * 4 + (H + 2 * H) * 4                                          # gates, c in; c, h, s out
    return {
        "lstm_seq_fwd": ("tensor", fl(B, 4 * H, H) * T), "lstm_seq_bwd": ("tensor", fl(B, H, 4 * H) * T),
        "dec_cell": ("hbm", float(cell_bytes) * decode_B), "dec_qr_gemm": ("tf32x3", fl(decode_B, 2 * a, 2 * H)),
        "gemm_vocab_fwd": ("tensor", fl(N, Vc, H)), "gemm_vocab_dx": ("tensor", fl(N, H, Vc)), "gemm_vocab_dw": ("tensor", fl(Vc, H, N)),
        "lstm_rec_gemm": ("tensor", fl(B, 4 * H, H)), "bptt_rec_gemm": ("tensor", fl(B, H, 4 * H)),
        "dec_vocab_gemm": ("tf32x3", fl(decode_B, Vc, H)), "dec_vocab_gemm1": ("bf16x1", fl(decode_B, Vc, H)), "dec_gate_gemm": ("tf32x3", fl(decode_B, 5 * H, E + H)),
        "dec_step_fused": ("hbm", float(step_bytes) * decode_B),
        # final reduction over the row's refined candidates (~2 entries) + gather of the next word's embedding into the (hi | lo) A operand
        "dec_argmax": ("hbm", float(decode_B) * (E * 4 + 2 * E * 4 + 8 + 4 + 3 * 8)),
        # candidate filter: one pass over the first pass's maxima [B, Vc/16] and over u (hi | lo) for the row norms
        "dec_argmax_filter": ("hbm", float(decode_B) * (((Vc + 15) // 16) * 4 + 2 * H * 4)),
        # training attention, per launch over the whole batch: V + P + per-step rows in, u/ctx/alpha/beta out
        "atten_fwd": ("hbm", float(B) * (k * H * 4 + k * a * 4 + T * (2 * a + 2 * H + 2 * H + k + 1) * 4)),
        "atten_bwd": ("hbm", float(B) * (2 * k * H * 4 + 2 * k * a * 4 + T * (2 * a + 3 * H + H + k + 1 + 2 * a) * 4)),
    }


def load_traffic():
    """dram bytes per launch of the profiled kernels, from the committed `ncu --set full` captures (profiles/traffic.json)."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        return {k: v["dram_bytes_per_launch"] for k, v in json.load(open(p)).items() if isinstance(v, dict)}
    except (OSError, ValueError, KeyError):
        return {}


TRAFFIC = load_traffic()


def rooflines(report, models, peaks):
    out = {}
    for tag, (ms, n) in report.items():
        if tag not in models or n == 0 or ms <= 0:
            out[tag] = {"ms_total": ms, "launches": n}
            continue
        bound, work = models[tag]
        per = ms / n * 1e-3
        extra = {}
        if bound == "hbm":
            ach, peak, unit = work / per / 1e9, peaks["hbm_gbs"], "GB/s"
        elif bound == "tf32x3":
            # fp32-accurate contraction executed as three tf32 tensor-core products: the tensor pipe does 3x the
            # algorithmic flops; there is no measured tf32 peak, so half of the measured bf16 peak (the nominal ratio) is used
            bound = "tensor"
            ach, peak, unit = 3.0 * work / per / 1e12, peaks["bf16_tflops"] / 2.0, "TFLOP/s"
            extra = {"engine": "tcgen05 kind::tf32 x3 (hi/lo split)", "algorithmic_tflops": work / per / 1e12,
                     "peak_note": "tf32 peak taken as measured bf16 peak / 2"}
        elif bound == "bf16x1":
            # single bf16 pass of the filter-and-refine arg-max (the exact logits of the candidate tiles come from dec_argmax_refine)
            bound = "tensor"
            ach, peak, unit = work / per / 1e12, peaks["bf16_tflops"], "TFLOP/s"
            extra = {"engine": "tcgen05 kind::f16 over bf16 mirrors, maxima per 16 columns only (no logits written)"}
        else:
            ach, peak, unit = work / per / 1e12, peaks["bf16_tflops"], "TFLOP/s"
        out[tag] = dict({"bound": bound, "achieved": ach, "peak": peak, "unit": unit, "frac": ach / peak, "traffic": TRAFFIC.get(tag),
                         "ms_total": ms, "launches": n, "us_per_launch": per * 1e6}, **extra)
    return out


def measure_widened(torch, dev, dims, peaks, steps):
    """SURVEY 8f rows built beyond the headline path, measured on one GPU at BASELINE config 2/3 sizes (N = 1 only):
    the encoder heads (aa_encoder_forward / aa_encoder_backward, bf16) and the sentinel-less baseline decoder (train step
    and greedy decode).  Each timed as a replayed CUDA graph (training) / as the plain call (decode), CUDA events."""
    import adaptive_b200
    from adaptive_b200 import _lib, baseline
    from adaptive_b200 import functional as F_aa
    from adaptive_b200.graphs import GraphedTrainStep
    from adaptive_b200.synth import baseline_weights, make_encoder_weights, make_features

    out = {}
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    def timed(fn, n):
        torch.cuda.synchronize()
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n

    # ---- encoder heads: B = 80 feature maps [2048, 7, 7] -> V, v_g, h0, c0 and back (all 8 parameter gradients + dA) ----
    C, hw = 2048, (7, 7)
    ew = make_encoder_weights(dims, C, seed=321)
    W = tuple(torch.from_numpy(ew[k]).to(dev).requires_grad_(True) for k in
              ("affine_a.weight", "affine_a.bias", "affine_b.weight", "affine_b.bias", "affine_h0.weight", "affine_h0.bias",
               "affine_c0.weight", "affine_c0.bias"))
    A = torch.from_numpy(make_features(TRAIN_B, C, hw, seed=4321)).to(dev).requires_grad_(True)
    ups = None

    def enc_step():
        nonlocal ups
        for t in W + (A,):
            t.grad = None
        outs = F_aa.encoder_forward(W, A, "bf16")
        if ups is None:
            ups = [torch.randn_like(o) for o in outs]
        torch.autograd.backward(outs, ups)

    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(3):
            enc_step()
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        enc_step()
    for _ in range(3):
        g.replay()
    ms = timed(g.replay, max(steps, 10))
    _lib.profile_reset()
    _lib.profile_enable(True)
    for _ in range(5):
        enc_step()
    torch.cuda.synchronize()
    _lib.profile_enable(False)
    rep = _lib.profile_report()
    _lib.profile_reset()
    M, k = TRAIN_B * hw[0] * hw[1], hw[0] * hw[1]
    fl = lambda m, n, kk: 2.0 * m * n * kk
    models = {
        # one pass over the NCHW map: fp32 in, bf16 transpose + pooled rows (fp32 + bf16) out
        "enc_transpose_pool": ("hbm", float(TRAIN_B) * C * (k * 4 + k * 2 + 4 + 2)),
        "enc_untranspose": ("hbm", float(TRAIN_B) * C * (k * 4 + k * 4 + 4)),
        "enc_gemm_V": ("tensor", fl(M, dims.H, C)), "enc_gemm_dWa": ("tensor", fl(dims.H, C, M)), "enc_gemm_dA": ("tensor", fl(M, C, dims.H)),
    }
    out["encoder_heads"] = {
        "workload": "AttentiveCNN heads fwd+bwd (affine_a/b/h0/c0 + average pool, incl. dA), batch %d, [2048,7,7] maps, bf16" % TRAIN_B,
        "value": TRAIN_B / (ms * 1e-3), "unit": "images/s", "ms_per_step": ms, "cuda_graph": True,
        "algorithmic_tflops": (3 * fl(M, dims.H, C) + 3 * fl(TRAIN_B, dims.E + 2 * dims.H, C)) / (ms * 1e-3) / 1e12,
        "kernels": rooflines(rep, models, peaks),
    }

    # ---- sentinel-less baseline decoder: same shapes as the headline (config 2) and as config 3 ----
    class Cf:
        base_word_embed_size, base_lstm_hidden_size, vocab_length = dims.E, dims.H, dims.Vc
        precision = "bf16"

    bm = baseline.Encoder2Decoder(Cf()).to(dev)
    bw = baseline_weights(make_weights(dims, seed=123))
    bm.decoder.load_state_dict({k: torch.from_numpy(v) for k, v in bw.items()}, strict=True)
    lengths = make_lengths(TRAIN_B, TRAIN_T, seed=1234)
    inp = make_inputs(dims, TRAIN_B, TRAIN_T, seed=1234)
    b = {k: torch.from_numpy(v).to(dev) for k, v in inp.items()}
    b["tgt"] = torch.from_numpy(np.ascontiguousarray(F_aa.packed_targets(inp["captions"], lengths))).to(dev)
    stepper = GraphedTrainStep(bm, b, lengths)
    for _ in range(3):
        stepper(b)
    ms = timed(lambda: stepper(b), max(steps, 10))
    dinp = make_inputs(dims, DECODE_B, 1, seed=4321)
    db = {k: torch.from_numpy(dinp[k]).to(dev) for k in ("V", "v_g", "h0", "c0")}
    dec = lambda: bm.sampler((db["V"], db["v_g"], (db["h0"], db["c0"])), max_len=DECODE_L)
    dec()
    dms = timed(dec, 3)
    out["baseline_decoder"] = {
        "workload": "baseline_attention.py model (no sentinel), config 2 / config 3 shapes",
        "train": {"value": TRAIN_B * TRAIN_T / (ms * 1e-3), "unit": "tokens/s", "ms_per_step": ms, "cuda_graph": True, "dtype": "bf16"},
        "decode": {"value": DECODE_B * DECODE_L / (dms * 1e-3), "unit": "tokens/s", "ms_per_step": dms, "precision": bm.decoder.decode_precision},
    }
    return out


_T0 = time.time()


def stage(msg):
    """Progress line on stderr (rank-tagged) and a re-armed watchdog: a stage that takes longer than AA_BENCH_STAGE_TIMEOUT
    seconds (default 240) dumps every thread's Python stack and exits non-zero instead of hanging the launcher."""
    import faulthandler

    sys.stderr.write("[bench rank %s +%.1fs] %s\n" % (os.environ.get("RANK", "0"), time.time() - _T0, msg))
    sys.stderr.flush()
    faulthandler.cancel_dump_traceback_later()
    faulthandler.dump_traceback_later(float(os.environ.get("AA_BENCH_STAGE_TIMEOUT", "240")), exit=True)


def leave(world):
    """End of a rank's run: flush what was printed, then (N > 1) leave without tearing NCCL down.
    ``destroy_process_group()`` after CUDA-graph-captured collectives blocked forever on the 2-GPU box (every stage done, the
    JSON line still in the stdout buffer); a finished benchmark process has nothing left to release that process exit does
    not release, so the ranks synchronise their devices and exit 0 directly."""
    import faulthandler

    faulthandler.cancel_dump_traceback_later()
    sys.stdout.flush()
    sys.stderr.flush()
    if world > 1:
        import torch

        torch.cuda.synchronize()
        os._exit(0)


def run_ours(args):
    import torch
    import torch.distributed as dist

    import adaptive_b200
    from adaptive_b200 import _lib
    from adaptive_b200 import functional as F_aa

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py --impl ours needs a CUDA device (there is no CPU fallback)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    stage("init (world %d)" % world)
    if world > 1:
        # NCCL writes its version / debug lines to stdout by default; stdout carries the ONE JSON line
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=dev)
    n_gpus = world
    peaks = load_peaks()
    dims = CFG_A

    class Cf:
        adaptive_word_embed_size, adaptive_lstm_hidden_size, vocab_length = dims.E, dims.H, dims.Vc

    from adaptive_b200.graphs import GraphedTrainStep

    model = adaptive_b200.Encoder2Decoder(Cf()).to(dev)
    model.decoder.precision = "bf16" if args.precision == "bf16" else "fp32"
    w = make_weights(dims, seed=123)
    model.load_state_dict({"decoder." + k: torch.from_numpy(v) for k, v in w.items()}, strict=False)
    params = [p for p in model.decoder.parameters()]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---------------- training workload (headline) ----------------
    NB = 4
    host, devb = [], []
    lengths = make_lengths(TRAIN_B, TRAIN_T, seed=1234)
    for i in range(NB):
        inp = make_inputs(dims, TRAIN_B, TRAIN_T, seed=1234 + 97 * rank + i)
        tgt = np.ascontiguousarray(F_aa.packed_targets(inp["captions"], lengths))
        hb = {k: torch.from_numpy(v).pin_memory() for k, v in inp.items()}
        hb["tgt"] = torch.from_numpy(tgt).pin_memory()
        host.append(hb)
        devb.append({k: v.to(dev) for k, v in hb.items()})
    h2d_train = sum(v.numel() * v.element_size() for v in host[0].values())

    def train_step_eager(b):
        for p in params:
            p.grad = None
        states = (b["h0"], b["c0"])
        packed = model((b["V"], b["v_g"], states), b["captions"], lengths)
        loss = F_aa.cross_entropy(packed.data, b["tgt"])
        loss.backward()
        return loss

    stage("first eager training step")
    # one eager step: counts the kernels of a step (graph replays bypass the library's launch counter)
    l0 = _lib.launch_count()
    train_step_eager(devb[0])
    torch.cuda.synchronize()
    launches_per_step = _lib.launch_count() - l0

    stage("building the stepper (graph=%d)" % args.graph)
    dp_mode = None
    if world == 1:
        stepper = GraphedTrainStep(model, devb[0], lengths) if args.graph else None

        def train_step(b):
            return stepper(b) if stepper is not None else train_step_eager(b)
    else:
        # data parallel: C-ABI forward / loss / hooked backward with the four gradient buckets all-reduced (NCCL, sum) on a
        # communication stream as soon as each is final, overlapping the rest of the backward; loss and gradients are
        # normalised by the GLOBAL packed-token count (adaptive_b200/parallel.py)
        from adaptive_b200.parallel import DataParallelTrainer, GraphedDPStep

        trainer = DataParallelTrainer(model, overlap=bool(args.overlap))

        def dp_eager(b):
            return trainer.step((b["V"], b["v_g"], (b["h0"], b["c0"])), b["captions"], lengths, b["tgt"])

        stepper = None
        dp_mode = "eager"
        if args.graph:
            ok = torch.ones(1, device=dev)
            try:
                stepper = GraphedDPStep(trainer, devb[0], lengths)
            except Exception as e:      # capture of the NCCL collectives refused: fall back to eager launches on every rank
                sys.stderr.write("rank %d: CUDA-graph capture of the data-parallel step failed (%s); running eagerly\n" % (rank, e))
                ok.zero_()
                stepper = None
            dist.all_reduce(ok, op=dist.ReduceOp.MIN)
            if float(ok.item()) < 1:
                stepper = None
            dp_mode = "graph" if stepper is not None else "eager"

        def train_step(b):
            return stepper(b) if stepper is not None else dp_eager(b)

    stage("training warm-up")
    for i in range(args.warmup):
        train_step(devb[i % NB])
    barrier()
    stage("training timed region")
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for i in range(args.steps):
        train_step(devb[i % NB])
    e1.record()
    barrier()
    train_ms = max_over_ranks(e0.elapsed_time(e1)) / args.steps
    launches = launches_per_step * args.steps
    clocks = sampler.stop() if rank == 0 else None
    train_tok = TRAIN_B * TRAIN_T * n_gpus / (train_ms * 1e-3)

    stage("training e2e")
    # e2e: through adaptive_b200.pipeline.HostPipeline (the loop a caller runs): every step's inputs are copied from pinned
    # host memory and its loss is read back, inside the timed region; the copy of batch i+1 overlaps step i
    from adaptive_b200.pipeline import HostPipeline

    pipe = HostPipeline(lambda b: train_step(b), host[0], dev)
    pipe.run(host[i % NB] for i in range(min(3, args.warmup)))
    barrier()
    t0 = time.perf_counter()
    losses = pipe.run(host[i % NB] for i in range(args.steps))
    torch.cuda.synchronize()
    e2e_train_s = max_over_ranks((time.perf_counter() - t0)) / args.steps
    assert len(losses) == args.steps and all(np.isfinite(float(x)) for x in losses)
    barrier()

    stage("per-kernel timing pass")
    # per-kernel timing pass (CUDA events around the kernels, same steps; not used for `value`)
    _lib.profile_reset()
    _lib.profile_enable(True)
    for i in range(args.steps):
        train_step_eager(devb[i % NB])
    torch.cuda.synchronize()
    _lib.profile_enable(False)
    train_report = _lib.profile_report()
    _lib.profile_reset()

    stage("decode")
    # ---------------- decode workload (BASELINE config 3) ----------------
    dsteps = max(2, min(args.steps, 5))
    dinp = make_inputs(dims, DECODE_B, 1, seed=4321 + rank)
    dhost = {k: torch.from_numpy(dinp[k]).pin_memory() for k in ("V", "v_g", "h0", "c0")}
    ddev = {k: v.to(dev) for k, v in dhost.items()}
    h2d_dec = sum(v.numel() * v.element_size() for v in dhost.values())

    def decode_step(b):
        return model.sampler((b["V"], b["v_g"], (b["h0"], b["c0"])), max_len=DECODE_L)

    for _ in range(2):
        decode_step(ddev)
    barrier()
    l0d = _lib.launch_count()
    e0.record()
    for _ in range(dsteps):
        decode_step(ddev)
    e1.record()
    barrier()
    dec_ms = max_over_ranks(e0.elapsed_time(e1)) / dsteps
    dec_launches = _lib.launch_count() - l0d
    dec_tok = DECODE_B * DECODE_L * n_gpus / (dec_ms * 1e-3)
    dpipe = HostPipeline(lambda b: decode_step(b)[0], dhost, dev)
    dpipe.run(dhost for _ in range(2))
    barrier()
    t0 = time.perf_counter()
    all_ids = dpipe.run(dhost for _ in range(dsteps))
    torch.cuda.synchronize()
    e2e_dec_s = max_over_ranks(time.perf_counter() - t0) / dsteps
    ids = all_ids[-1]
    d2h_dec = ids.numel() * ids.element_size()
    barrier()
    _lib.profile_enable(True)
    for _ in range(dsteps):
        decode_step(ddev)
    torch.cuda.synchronize()
    _lib.profile_enable(False)
    dec_report = _lib.profile_report()
    _lib.profile_reset()
    # filter-and-refine arg-max: (row, 16-column tile) pairs recomputed exactly, per row and step
    lib = _lib.load()
    lib.aa_debug_refine_pairs(1)
    decode_step(ddev)
    refine_pairs = lib.aa_debug_refine_pairs(1) / float(DECODE_B * DECODE_L)

    widened = None
    if world == 1 and not args.no_widened:
        stage("widened rows (encoder heads, baseline decoder)")
        widened = measure_widened(torch, dev, dims, peaks, args.steps)

    stage("report")
    if world > 1:
        dist.barrier()
    if rank != 0:
        leave(world)
        return

    models = kernel_models(dims, TRAIN_B, TRAIN_T, DECODE_B)
    k_train = rooflines(train_report, models, peaks)
    k_dec = rooflines(dec_report, models, peaks)
    dom = max((t for t in k_train if "frac" in k_train[t]), key=lambda t: k_train[t]["ms_total"])
    roof = {kk: k_train[dom][kk] for kk in ("bound", "achieved", "peak", "unit", "frac", "traffic")}
    roof["kernel"] = dom
    if dom.startswith("lstm_seq"):
        roof["note"] = ("latency-bound: %d dependent recurrence steps of a [%d x %d] x [%d x %d] contraction inside one launch (8 clusters "
                        "of 16 CTAs, weights resident in tensor memory, state exchanged through distributed shared memory: nothing "
                        "of a step touches HBM); the HBM-bound kernels of this step and of decoding are listed under kernels" %
                        (TRAIN_T, TRAIN_B, 4 * dims.H, 4 * dims.H, dims.H))
    roof["peak_source"] = peaks["source"]
    roof["share_of_step"] = k_train[dom]["ms_total"] / max(sum(v["ms_total"] for v in k_train.values()), 1e-9)

    out = {
        "metric": METRIC, "value": train_tok, "unit": "tokens/s", "n_gpus": n_gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": train_ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": args.precision, "data": "synthetic",
        "config": dict(workload_config(n_gpus), cuda_graph=bool(stepper is not None),
                       **({"dp": "4 gradient buckets, NCCL sum all-reduce %s the backward (%s launches)" %
                           ("overlapped with" if args.overlap else "after", dp_mode)} if world > 1 else {})),
        "clocks": clocks,
        "e2e": {"value": TRAIN_B * TRAIN_T * n_gpus / e2e_train_s, "unit": "tokens/s", "h2d_bytes_per_step": h2d_train,
                "d2h_bytes_per_step": 4, "ms_per_step": e2e_train_s * 1e3},
        "gpu_launches": int(launches),
        "roofline": roof,
        "kernels": {"train": k_train, "decode": k_dec},
        "decode": {"workload": "BASELINE config 3: greedy sampler, batch %d per GPU, max_len %d, fp32; V (411 MB) > L2" % (DECODE_B, DECODE_L),
                   "value": dec_tok, "unit": "tokens/s", "ms_per_step": dec_ms, "steps": dsteps, "gpu_launches": int(dec_launches),
                   "e2e": {"value": DECODE_B * DECODE_L * n_gpus / e2e_dec_s, "unit": "tokens/s", "h2d_bytes_per_step": h2d_dec,
                           "d2h_bytes_per_step": d2h_dec, "ms_per_step": e2e_dec_s * 1e3},
                   "precision": model.decoder.decode_precision,
                   "vocab_argmax": {"method": "one bf16 tensor-core pass (maxima per 16 columns) + exact fp32 recompute of the tiles that can "
                                              "hold the row maximum under a rigorous error bound (vocab_refine.cu); ids equal an exact fp32 projection's",
                                    "tiles_refined_per_row_and_step": refine_pairs, "tiles_per_row": (dims.Vc + 15) // 16},
                   "roofline_fused_step": k_dec.get("dec_step_fused")},
    }
    if widened is not None:
        out["widened"] = widened
    if n_gpus == 1:
        cores = os.cpu_count() or 1
        step = oracle_train_step_fn(TRAIN_B, TRAIN_T, dims)
        ncpu = 20
        sec = time_cpu(step, ncpu, 2)
        out["cpu_baseline"] = {"value": TRAIN_B * TRAIN_T / sec, "unit": "tokens/s", "cores": cores, "kind": "port",
                               "sample": "%d full steps of the same workload (B=%d, T=%d) on the numpy oracle port, %.1f s" %
                                         (ncpu, TRAIN_B, TRAIN_T, sec * ncpu)}
    emit(out)
    leave(world)


_JSON_OUT = None


def protect_stdout():
    """stdout carries exactly ONE JSON line.  Libraries write there too (NCCL prints its version banner to stdout whatever
    NCCL_DEBUG_FILE says), so the real stdout is kept on a private descriptor for the JSON line and fd 1 is pointed at stderr."""
    global _JSON_OUT
    if _JSON_OUT is None:
        sys.stdout.flush()
        _JSON_OUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def emit(obj):
    out = _JSON_OUT if _JSON_OUT is not None else sys.stdout
    out.write(json.dumps(obj) + "\n")
    out.flush()


def main():
    protect_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32"], help="training path: bf16 tensor cores (BASELINE config 2) or exact fp32")
    ap.add_argument("--graph", type=int, default=1, help="replay the training step as one CUDA graph (1) or launch eagerly (0)")
    ap.add_argument("--overlap", type=int, default=1, help="N>1: start each gradient bucket's all-reduce as soon as the backward finishes it")
    ap.add_argument("--no-widened", action="store_true", help="skip the extra measurements of the SURVEY 8f rows (N=1 only)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
