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
import ctypes
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
# CPU arm: the reference's own modules (oracle/_ref, staged by build()) on the host cores; the numpy oracle port only if
# the staging is missing (kind says which ran)
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


def cpu_train_step_fn(B, T, dims):
    """-> (step, kind): config 2's training step on the unmodified reference modules (CPU), or on the numpy port."""
    try:
        from oracle import ref_runner

        step = ref_runner.train_step_fn(dims, B, T, make_lengths(B, T, seed=1234), "cpu")
        step()
        return step, "reference"
    except Exception as e:      # staging missing / import failed: say so and time the port
        sys.stderr.write("bench: reference modules unavailable (%s: %s); timing the numpy oracle port\n" % (type(e).__name__, e))
        return oracle_train_step_fn(B, T, dims), "port"


def time_cpu(step, steps, warmup):
    """Seconds per step with every host core (torchrun exports OMP_NUM_THREADS=1 to its workers, which would otherwise pin
    the CPU arm to one thread): torch's intra-op pool for the reference modules, the BLAS pool for the numpy port."""
    import torch
    from threadpoolctl import threadpool_limits

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    with threadpool_limits(limits=cores):
        for _ in range(warmup):
            step()
        t0 = time.perf_counter()
        for _ in range(steps):
            step()
        return (time.perf_counter() - t0) / steps


def cpu_config1():
    """BASELINE config 1 exactly: the reference's Encoder2Decoder.forward on CPU, batch 4, [4,2048,7,7] synthetic features through
    resnet_conv = Identity, hidden 512, vocab 10k, caption len 18, no_grad.  -> dict or None."""
    try:
        from oracle import ref_runner

        run = ref_runner.config1_fn(CFG_A, 4, TRAIN_T)
        sec = time_cpu(run, 20, 3)
        return {"workload": "BASELINE config 1: reference Encoder2Decoder.forward on CPU, batch 4, [4,2048,7,7] features, hidden 512, "
                            "vocab 10000, caption len 18, no_grad", "value": 4 * TRAIN_T / sec, "unit": "tokens/s", "ms_per_step": sec * 1e3,
                "cores": os.cpu_count() or 1, "kind": "reference", "steps": 20}
    except Exception as e:
        sys.stderr.write("bench: config 1 on the reference failed (%s: %s)\n" % (type(e).__name__, e))
        return None


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    step, kind = cpu_train_step_fn(TRAIN_B, TRAIN_T, CFG_A)
    steps = max(1, min(args.steps, 200))
    sec = time_cpu(step, steps, args.warmup)
    v = TRAIN_B * TRAIN_T / sec
    n_real = int(sum(make_lengths(TRAIN_B, TRAIN_T, seed=1234)))
    what = ("the unmodified reference modules (oracle/_ref: adaptive_attention.Decoder fwd -> pack_padded_sequence -> CrossEntropyLoss -> "
            "backward, torch CPU kernels)") if kind == "reference" else "the numpy oracle port"
    out = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": "tokens/s", "n_gpus": args.gpus, "steps": steps,
        "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args.gpus),
        "cpu_baseline": {"value": v, "unit": "tokens/s", "cores": cores, "kind": kind,
                         "sample": "%d full steps of the same workload (B=%d, T=%d) on %s" % (steps, TRAIN_B, TRAIN_T, what)},
        "e2e": {"value": v, "unit": "tokens/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "tokens_per_step": {"positions_B_x_T": TRAIN_B * TRAIN_T, "packed_rows_sum_lengths": n_real,
                            "value_over_packed_rows": n_real / sec,
                            "note": "the metric counts B*T decoder positions (SURVEY 8d); the reference computes all of them"},
        "config1": cpu_config1(),
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
    # cell: recurrent gates (4H) + static sentinel block (H) + the word's EG row (5H, L2-resident) + c in; c, [h | s], tf32 (hi, lo) of h and s out
    cell_bytes = (4 * H + H + 5 * H + H) * 4 + (H + 2 * H + 4 * H) * 4
    return {
        "lstm_seq_fwd": ("tensor", fl(B, 4 * H, H) * T), "lstm_seq_bwd": ("tensor", fl(B, H, 4 * H) * T),
        "dec_cell": ("hbm", float(cell_bytes) * decode_B), "dec_qr_gemm": ("tf32x3", fl(decode_B, 2 * a, H)),
        "dec_prologue_P": ("tf32x3", fl(decode_B * k, a, H)), "dec_gate_table": ("tf32x3", fl(Vc, 5 * H, E)),
        "gemm_vocab_fwd": ("tensor", fl(N, Vc, H)), "gemm_vocab_dx": ("tensor", fl(N, H, Vc)), "gemm_vocab_dw": ("tensor", fl(Vc, H, N)),
        "lstm_rec_gemm": ("tensor", fl(B, 4 * H, H)), "bptt_rec_gemm": ("tensor", fl(B, H, 4 * H)),
        "dec_vocab_gemm": ("tf32x3", fl(decode_B, Vc, H)), "dec_vocab_gemm1": ("bf16x1", fl(decode_B, Vc, H)), "dec_gate_gemm": ("tf32x3", fl(decode_B, 4 * H, H)),
        "dec_step_fused": ("hbm", float(step_bytes) * decode_B),
        # final reduction over the row's refined candidates (~2 entries) + gather of the next word's embedding into the (hi | lo) A operand
        "dec_argmax": ("hbm", float(decode_B) * (E * 4 + 2 * E * 4 + 8 + 4 + 3 * 8)),
        # candidate filter: one pass over the first pass's maxima [B, Vc/16] and over u (hi | lo) for the row norms
        "dec_argmax_filter": ("hbm", float(decode_B) * (((Vc + 15) // 16) * 4 + 2 * H * 4 + H * 2)),
        # the small contractions of the step, grouped by tag: (flops per step) / (launches per step)
        "gemm_att_dw": ("tensor", (fl(a, H, B * k) + 2 * fl(a, H, N)) / 3), "gemm_att_dx": ("tensor", (fl(B * k, H, a) + 2 * fl(N, H, a)) / 3),
        "gemm_sent_dw": ("tensor", (fl(H, 2 * E, N) + fl(H, H, N)) / 2), "gemm_sent_dx": ("tensor", (fl(N, 2 * E, H) + fl(N, H, H)) / 2),
        "gemm_lstm_dw": ("tensor", (fl(4 * H, 2 * E, N) + fl(4 * H, H, N)) / 2), "gemm_lstm_dx": ("tensor", fl(N, 2 * E, 4 * H)),
        "gemm_gates_in": ("tensor", (fl(N, 4 * H, 2 * E) + fl(N, H, 2 * E)) / 2), "gemm_qr": ("tensor", fl(N, a, H)),
        "gemm_P": ("tensor", fl(B * k, a, H)), "gemm_sentinel_h": ("tensor", fl(N, H, H)),
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



def _events(torch):
    return torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)


def time_fn(torch, fn, n, warm=1):
    """ms per call of `fn` over n calls, CUDA events on the current stream, synchronised on both sides."""
    for _ in range(warm):
        fn()
    e0, e1 = _events(torch)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def gpu_eager_baseline(torch, dev, dims, lengths):
    """The bar BASELINE.md section 1 names: the UNMODIFIED reference modules (oracle/_ref) on this B200 through stock PyTorch
    eager (cuDNN LSTM, cuBLAS), same synthetic weights / inputs / shapes as our arm.  Config 2 train step in true fp32
    (allow_tf32 off), with TF32 allowed, and under bf16 autocast; config 3 greedy loop (sampler loop body, [1,B,H] states)."""
    try:
        from oracle import ref_runner
    except Exception as e:
        return {"unavailable": "%s: %s" % (type(e).__name__, e)}
    out = {"what": "reference modules from oracle/_ref on cuda:%d, PyTorch %s eager (no torch.compile, no CUDA graph)" % (dev.index, torch.__version__)}
    mm, cd = torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32
    try:
        train = {}
        for tag, tf32, ac in (("fp32", False, False), ("tf32", True, False), ("bf16_autocast", True, True)):
            torch.backends.cuda.matmul.allow_tf32 = tf32
            torch.backends.cudnn.allow_tf32 = tf32
            try:
                step = ref_runner.train_step_fn(dims, TRAIN_B, TRAIN_T, lengths, dev, autocast_bf16=ac)
                ms = time_fn(torch, step, 20, warm=5)
                train[tag] = {"value": TRAIN_B * TRAIN_T / (ms * 1e-3), "unit": "tokens/s", "ms_per_step": ms}
            except Exception as e:
                train[tag] = {"failed": "%s: %s" % (type(e).__name__, str(e)[:200])}
        out["train_config2"] = train
        dec = {}
        for tag, tf32 in (("fp32", False), ("tf32", True)):
            torch.backends.cuda.matmul.allow_tf32 = tf32
            torch.backends.cudnn.allow_tf32 = tf32
            try:
                run = ref_runner.greedy_fn(dims, DECODE_B, DECODE_L, dev)
                ms = time_fn(torch, run, 3, warm=1)
                dec[tag] = {"value": DECODE_B * DECODE_L / (ms * 1e-3), "unit": "tokens/s", "ms_per_batch": ms}
                if tag == "fp32":
                    run400 = ref_runner.greedy_fn(dims, 400, DECODE_L, dev)
                    ms4 = time_fn(torch, run400, 5, warm=2)
                    dec["fp32_batch400"] = {"value": 400 * DECODE_L / (ms4 * 1e-3), "unit": "tokens/s", "ms_per_batch": ms4}
            except Exception as e:
                dec[tag] = {"failed": "%s: %s" % (type(e).__name__, str(e)[:200])}
        out["greedy_config3"] = dec
    finally:
        torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = mm, cd
    return out


def measure_decode_small(torch, dev, model, dims, peaks):
    """Greedy decoding below config 3's batch (the reference's evaluation batch is 400, cfg_wzn.py:84): eager per-step
    launches, the same pipeline replayed as ONE CUDA graph, and the persistent kernel (one cooperative launch, V / P / cell
    state resident in shared memory) where the batch fits one image per SM."""
    from adaptive_b200 import _lib
    from adaptive_b200 import functional as F_aa
    from adaptive_b200.graphs import GraphedSampler

    k, a, H, E = dims.k, dims.a, dims.H, dims.E
    # SURVEY 8d, persistent variant: per (image, step) 18 632 B of step traffic + V and P read once per image
    out = {"max_len": DECODE_L, "algorithmic_bytes_per_image_step": {"per_launch_pipeline": 128588, "persistent": 18632 + (k * H * 4 + k * a * 4) / DECODE_L}}
    sms = torch.cuda.get_device_properties(dev).multi_processor_count
    rows = []
    eng0 = model.decoder.decode_engine
    try:
        for B in (1, 64, sms, 296, 400):
            inp = make_inputs(dims, B, 1, seed=900 + B)
            b = {kk: torch.from_numpy(inp[kk]).to(dev) for kk in ("V", "v_g", "h0", "c0")}
            enc = (b["V"], b["v_g"], (b["h0"], b["c0"]))
            row = {"batch": B}
            model.decoder.decode_engine = "pipeline"
            l0 = _lib.launch_count()
            ids_pipe = model.sampler(enc, max_len=DECODE_L)[0]
            torch.cuda.synchronize()
            row["launches_pipeline"] = _lib.launch_count() - l0
            ms = time_fn(torch, lambda: model.sampler(enc, max_len=DECODE_L), 10, warm=2)
            row["pipeline_eager"] = {"tokens_per_s": B * DECODE_L / (ms * 1e-3), "ms": ms}
            gs = GraphedSampler(model, b, max_len=DECODE_L)
            ms = time_fn(torch, lambda: gs(b), 20, warm=3)
            row["pipeline_cuda_graph"] = {"tokens_per_s": B * DECODE_L / (ms * 1e-3), "ms": ms, "launches": 2}
            del gs
            if F_aa.persistent_decode_supported(model.decoder.weights(), b["V"], b["v_g"], DECODE_L):
                model.decoder.decode_engine = "persistent"
                l0 = _lib.launch_count()
                ids_pers = model.sampler(enc, max_len=DECODE_L)[0]
                torch.cuda.synchronize()
                row["launches_persistent"] = _lib.launch_count() - l0
                ms = time_fn(torch, lambda: model.sampler(enc, max_len=DECODE_L), 20, warm=3)
                alg = B * DECODE_L * out["algorithmic_bytes_per_image_step"]["persistent"]
                gp = GraphedSampler(model, b, max_len=DECODE_L)
                msg = time_fn(torch, lambda: gp(b), 20, warm=3)
                del gp
                row["persistent_cuda_graph"] = {"tokens_per_s": B * DECODE_L / (msg * 1e-3), "ms": msg}
                row["persistent"] = {"tokens_per_s": B * DECODE_L / (ms * 1e-3), "ms": ms, "us_per_step": ms * 1e3 / DECODE_L,
                                     "ids_equal_pipeline": float((ids_pers == ids_pipe).all(1).float().mean()),
                                     "roofline": {"bound": "hbm", "achieved": alg / (ms * 1e-3) / 1e9, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                                                  "frac": alg / (ms * 1e-3) / 1e9 / peaks["hbm_gbs"],
                                                  "note": "latency-bound: 4 grid barriers + two in-kernel contractions per step; the HBM traffic it is "
                                                          "charged with is V, P once per image + 18.6 KB per step"}}
            else:
                row["persistent"] = None
                if sms < B <= 2 * sms:
                    # two persistent launches (functional.greedy_decode, engine "auto"): what model.sampler does by default here
                    model.decoder.decode_engine = "auto"
                    ids_auto = model.sampler(enc, max_len=DECODE_L)[0]
                    ms = time_fn(torch, lambda: model.sampler(enc, max_len=DECODE_L), 20, warm=3)
                    row["auto_two_persistent_launches"] = {"tokens_per_s": B * DECODE_L / (ms * 1e-3), "ms": ms,
                                                           "ids_equal_pipeline": float((ids_auto == ids_pipe).all(1).float().mean())}
            rows.append(row)
    finally:
        model.decoder.decode_engine = eng0
    out["rows"] = rows
    return out


def measure_beam(torch, dev, model, dims, barrier, max_over_ranks, n_gpus, rank):
    """BASELINE config 4: beam search (beam 3), images sharded over the ranks (contiguous ranges, no collective), 49 regions,
    max_len 20.  tokens/s = images * max_len over the max-over-ranks device time."""
    from adaptive_b200 import _lib

    B, beam = DECODE_B, 3
    inp = make_inputs(dims, B, 1, seed=4321 + rank)
    b = {kk: torch.from_numpy(inp[kk]).to(dev) for kk in ("V", "v_g", "h0", "c0")}
    enc = (b["V"], b["v_g"], (b["h0"], b["c0"]))
    run = lambda: model.beam_sampler(enc, beam=beam, max_len=DECODE_L)
    run()
    barrier()
    l0 = _lib.launch_count()
    e0, e1 = _events(torch)
    e0.record()
    for _ in range(3):
        out = run()
    e1.record()
    barrier()
    ms = max_over_ranks(e0.elapsed_time(e1)) / 3
    launches = (_lib.launch_count() - l0) // 3
    _lib.profile_reset()
    _lib.profile_enable(True)
    run()
    torch.cuda.synchronize()
    _lib.profile_enable(False)
    rep = _lib.profile_report()
    _lib.profile_reset()
    tot = sum(v[0] for v in rep.values()) or 1.0
    return {"workload": "BASELINE config 4: beam search, beam %d, batch %d per GPU (sharded by image, no collective), 49 regions, max_len %d"
                        % (beam, B, DECODE_L),
            "value": B * DECODE_L * n_gpus / (ms * 1e-3), "unit": "tokens/s", "ms_per_batch": ms, "gpu_launches_per_batch": int(launches),
            "parity": "unpinned: the reference has no beam search (SURVEY Q14); definition = oracle.beam_decode",
            "kernel_shares_profiled": {k: v[0] / tot for k, v in sorted(rep.items(), key=lambda kv: -kv[1][0])},
            "profiled_ms_note": "shares of the tagged kernels only (beam_select / row_lse / backtrack are untagged): profiled %.2f ms of %.2f ms"
                                % (tot, ms)}


def dp_check(torch, dist, dev, model, dims, world, rank):
    """Multi-GPU correctness on the hardware the numbers come from:
    (1) one data-parallel step on per-rank batches == the single-GPU step on the concatenated batch (global-count mean CE,
        train.py:63,208): loss and all 13 reduced gradients, in exact fp32 (<= 1e-5 relative) and in bf16 (<= 2e-2);
    (2) image-sharded greedy ids == the unsharded run, bit-exact."""
    from adaptive_b200 import functional as F_aa
    from adaptive_b200._lib import WEIGHT_FIELDS
    from adaptive_b200.parallel import DataParallelTrainer, shard_range

    B, T = 16, TRAIN_T
    res = {"world": world}
    lengths = make_lengths(B, T, seed=55)                       # same (sorted) lengths on every rank
    inp = make_inputs(dims, B, T, seed=7000 + rank)
    b = {k: torch.from_numpy(v).to(dev) for k, v in inp.items()}
    tgt = torch.from_numpy(np.ascontiguousarray(F_aa.packed_targets(inp["captions"], lengths))).to(dev)
    prec0 = model.decoder.precision
    # the concatenated batch, rows sorted by length (stable) as pack_padded_sequence demands
    cat_len = [L for _ in range(world) for L in lengths]
    order = sorted(range(len(cat_len)), key=lambda i: -cat_len[i])
    gathered = {}
    for k in ("V", "v_g", "h0", "c0", "captions"):
        parts = [torch.empty_like(b[k]) for _ in range(world)]
        dist.all_gather(parts, b[k].contiguous())
        gathered[k] = torch.cat(parts, 0)[order].contiguous()
    glen = [cat_len[i] for i in order]
    params = list(model.decoder.weights())
    try:
        for prec, tol in (("fp32", 1e-5), ("bf16", 2e-2)):
            model.decoder.precision = prec
            trainer = DataParallelTrainer(model, overlap=True)
            loss = trainer.step((b["V"], b["v_g"], (b["h0"], b["c0"])), b["captions"], lengths, tgt)
            torch.cuda.synchronize()
            dp_grads = [p.grad.detach().clone() for p in params]
            dp_loss = float(loss)
            worst, worst_name, ref_loss = 0.0, "", dp_loss
            if rank == 0:
                for p in params:
                    p.grad = None
                gt = torch.from_numpy(np.ascontiguousarray(F_aa.packed_targets(gathered["captions"].cpu().numpy(), glen))).to(dev)
                packed = model((gathered["V"], gathered["v_g"], (gathered["h0"], gathered["c0"])), gathered["captions"], glen)
                l1 = F_aa.cross_entropy(packed.data, gt)
                l1.backward()
                torch.cuda.synchronize()
                ref_loss = float(l1.detach())
                for name, p, g in zip(WEIGHT_FIELDS, params, dp_grads):
                    err = float((p.grad - g).abs().max() / p.grad.abs().max().clamp_min(1e-30))
                    if err > worst:
                        worst, worst_name = err, name
            for p in params:
                p.grad = None
            res[prec] = {"max_rel_grad_err": worst, "worst": worst_name, "loss_dp": dp_loss, "loss_single": ref_loss,
                         "loss_rel_err": abs(dp_loss - ref_loss) / max(abs(ref_loss), 1e-30), "tol": tol,
                         "ok": bool(worst <= tol and abs(dp_loss - ref_loss) <= tol * max(abs(ref_loss), 1.0))}
            del trainer
    finally:
        model.decoder.precision = prec0
    # sharded greedy decode vs the whole batch on rank 0
    n_img, L = 8 * world + 3, 12
    dinp = make_inputs(dims, n_img, 1, seed=4242)               # same on every rank
    full = {k: torch.from_numpy(dinp[k]).to(dev) for k in ("V", "v_g", "h0", "c0")}
    lo, hi = shard_range(n_img, rank, world)
    ids_ok = {}
    for eng in ("pipeline", "auto"):
        model.decoder.decode_engine = eng
        mine = model.sampler((full["V"][lo:hi], full["v_g"][lo:hi], (full["h0"][lo:hi], full["c0"][lo:hi])), max_len=L)[0]
        pad = torch.zeros(8 + 3, L, dtype=torch.int64, device=dev)
        pad[: hi - lo] = mine
        parts = [torch.empty_like(pad) for _ in range(world)]
        dist.all_gather(parts, pad)
        if rank == 0:
            got = torch.cat([parts[r][: shard_range(n_img, r, world)[1] - shard_range(n_img, r, world)[0]] for r in range(world)], 0)
            want = model.sampler((full["V"], full["v_g"], (full["h0"], full["c0"])), max_len=L)[0]
            ids_ok[eng] = bool(torch.equal(got, want))
    model.decoder.decode_engine = os.environ.get("AA_DECODE_ENGINE", "auto")
    res["ids_equal"] = all(ids_ok.values()) if rank == 0 else True
    res["ids_equal_by_engine"] = ids_ok
    res["max_rel"] = max(res["fp32"]["max_rel_grad_err"], 0.0)
    res["ok"] = bool(res["fp32"]["ok"] and res["bf16"]["ok"] and res["ids_equal"])
    flag = torch.tensor([1.0 if res["ok"] else 0.0], device=dev)
    dist.broadcast(flag, 0)
    res["ok_rank0"] = bool(flag.item() > 0)
    return res


def measure_config5(torch, dist, dev, world, rank, barrier, max_over_ranks, peaks, steps):
    """BASELINE config 5: data-parallel training, 14x14 = 196 regions, hidden 1024, vocab 20k, batch 256 per GPU, T = 18, bf16,
    E = 512 (assumed: the reference's E = H/2).  One graphed step per rank (+ NCCL all-reduce at N > 1)."""
    import adaptive_b200
    from adaptive_b200 import functional as F_aa
    from adaptive_b200.synth import CFG_B

    dims, B, T = CFG_B, 256, TRAIN_T

    class Cf:
        adaptive_word_embed_size, adaptive_lstm_hidden_size, vocab_length = dims.E, dims.H, dims.Vc
        precision = "bf16"

    model = adaptive_b200.Encoder2Decoder(Cf()).to(dev)
    g = torch.Generator(device="cpu").manual_seed(5)
    with torch.no_grad():       # (the orthogonal LSTM init of make_weights costs seconds at H=1024: scaled gaussians of the same variance)
        for name, p in model.decoder.named_parameters():
            if p.dim() >= 2:
                p.copy_((torch.randn(p.shape, generator=g) / (p.shape[1] ** 0.5)).to(dev))
    lengths = make_lengths(B, T, seed=1234)
    rng = np.random.Generator(np.random.PCG64(99 + rank))
    b = {"V": torch.relu(torch.randn(B, dims.k, dims.H, generator=g)).to(dev), "v_g": torch.relu(torch.randn(B, dims.E, generator=g)).to(dev),
         "h0": torch.tanh(torch.randn(B, dims.H, generator=g)).to(dev), "c0": torch.tanh(torch.randn(B, dims.H, generator=g)).to(dev)}
    cap = rng.integers(4, dims.Vc, size=(B, T), dtype=np.int64)
    cap[:, 0] = 1
    b["captions"] = torch.from_numpy(cap).to(dev)
    b["tgt"] = torch.from_numpy(np.ascontiguousarray(F_aa.packed_targets(cap, lengths))).to(dev)
    if world == 1:
        from adaptive_b200.graphs import GraphedTrainStep

        stepper = GraphedTrainStep(model, b, lengths)
        mode = "graph"
    else:
        from adaptive_b200.parallel import DataParallelTrainer, GraphedDPStep

        trainer = DataParallelTrainer(model, overlap=True)
        stepper = GraphedDPStep(trainer, b, lengths)
        if trainer.buckets.symm is None:
            NCCL_CAPTURED[0] = True
        mode = "graph+exchange"
    for _ in range(3):
        stepper(b)
    barrier()
    e0, e1 = _events(torch)
    e0.record()
    n = max(5, min(steps, 20))
    for _ in range(n):
        loss = stepper(b)
    e1.record()
    barrier()
    ms = max_over_ranks(e0.elapsed_time(e1)) / n
    assert np.isfinite(float(loss))
    flops = 191e6 * B * T               # 191 MFLOP per token fwd+bwd at cfgB
    n_par = sum(p.numel() for p in model.decoder.parameters())
    out = {"workload": "BASELINE config 5: DP training, 196 regions, hidden 1024, embed 512, vocab 20000, batch %d per GPU, caption len %d, bf16" % (B, T),
           "value": B * T * world / (ms * 1e-3), "unit": "tokens/s", "ms_per_step": ms, "n_gpus": world, "exchange": (trainer.engine if world > 1 else "none"), "mode": mode, "steps": n,
           "gradient_bytes_fp32": 4 * n_par,
           "whole_step_roofline": {"bound": "tensor", "algorithmic_tflops": flops / (ms * 1e-3) / 1e12, "peak": peaks["bf16_tflops_sustained"] or peaks["bf16_tflops"],
                                   "frac": flops / (ms * 1e-3) / 1e12 / (peaks["bf16_tflops_sustained"] or peaks["bf16_tflops"]), "unit": "TFLOP/s"}}
    del stepper
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


NCCL_CAPTURED = [False]     # set when a CUDA graph holding NCCL collectives was built in this process


def leave(world, rc=0):
    """End of a rank's run: flush what was printed, then (N > 1) tear the process group down.
    Root cause of round 1's hang, re-examined in round 2 (``gpurun_out/r02_clean_exit*.err``): ``destroy_process_group()`` blocks
    forever exactly when NCCL collectives were captured into a CUDA graph in this process (the fallback exchange engine); with the
    exchange in our own peer-memory kernels nothing of NCCL is ever captured and the teardown returns.  So: clean teardown unless an
    NCCL-holding graph exists -- then, as before, synchronise and leave through ``os._exit`` (a finished benchmark process has
    nothing left to release that process exit does not release)."""
    import faulthandler

    faulthandler.cancel_dump_traceback_later()
    sys.stdout.flush()
    sys.stderr.flush()
    if world > 1:
        import torch
        import torch.distributed as dist

        torch.cuda.synchronize()
        if not NCCL_CAPTURED[0] and os.environ.get("AA_BENCH_CLEAN_EXIT", "1") == "1":
            try:
                dist.barrier()
                dist.destroy_process_group()
            except Exception as e:      # (a peer already gone after a failed check: nothing left to tear down together)
                sys.stderr.write("[bench rank %s] destroy_process_group(): %s\n" % (os.environ.get("RANK", "0"), e))
                os._exit(rc)
            sys.stderr.flush()
            sys.exit(rc)
        os._exit(rc)
    if rc:
        sys.exit(rc)


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
    if world == 1 and not args.dp_path:
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
            if world > 1:
                dist.all_reduce(ok, op=dist.ReduceOp.MIN)
            if float(ok.item()) < 1:
                stepper = None
            dp_mode = "graph" if stepper is not None else "eager"
            if stepper is not None and trainer.buckets.symm is None:
                NCCL_CAPTURED[0] = True

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
    # Blocks of EXACTLY --steps steps, each bracketed by barrier + synchronize on both sides and timed with CUDA events; the
    # per-block time is the max over ranks; `ms_per_step` is the MEDIAN block (>= 50 blocks and >= ~1 s under the clock
    # sampler: one 20-step block is 7 ms, too short for nvidia-smi's 100 ms sampling and for a stable number).
    barrier()
    e0.record()
    for i in range(args.steps):
        train_step(devb[i % NB])
    e1.record()
    barrier()
    first_block = e0.elapsed_time(e1)
    nblocks = int(min(400, max(50, np.ceil(1000.0 / max(first_block, 1e-3))))) if args.blocks <= 0 else args.blocks
    if world > 1:
        t = torch.tensor([float(nblocks)], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        nblocks = int(t.item())
    block_ms = torch.zeros(nblocks, dtype=torch.float64)
    step_i = 0
    for blk in range(nblocks):
        barrier()
        e0.record()
        for _ in range(args.steps):
            train_step(devb[step_i % NB])
            step_i += 1
        e1.record()
        barrier()
        block_ms[blk] = e0.elapsed_time(e1)
    if world > 1:
        bm = block_ms.to(dev)
        dist.all_reduce(bm, op=dist.ReduceOp.MAX)
        block_ms = bm.cpu()
    per_step = (block_ms / args.steps).numpy()
    train_ms = float(np.median(per_step))
    spread = {"blocks": int(nblocks), "steps_per_block": args.steps, "min": float(per_step.min()), "p10": float(np.percentile(per_step, 10)),
              "median": train_ms, "p90": float(np.percentile(per_step, 90)), "max": float(per_step.max()),
              "timed_region_s": float(block_ms.sum() / 1e3)}
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
    # (a multiple of --steps long enough for ~0.3 s: with 20 steps the un-overlapped upload of the first batch and the drain of the
    #  last result were ~5 % of the region -- the loop a caller runs is long)
    n_e2e = args.steps * max(1, min(100, int(np.ceil(300.0 / max(train_ms * args.steps, 1e-3)))))
    t0 = time.perf_counter()
    losses = pipe.run(host[i % NB] for i in range(n_e2e))
    torch.cuda.synchronize()
    e2e_train_s = max_over_ranks((time.perf_counter() - t0)) / n_e2e
    assert len(losses) == n_e2e and all(np.isfinite(float(x)) for x in losses)
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
        return model.sampler((b["V"], b["v_g"], (b["h0"], b["c0"])), max_len=DECODE_L)     # (B = 4096: the engine choice falls on the pipeline)

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
    # the same loop captured once as ONE CUDA graph (graphs.GraphedSampler), replayed on inputs resident in its static buffers
    from adaptive_b200.graphs import GraphedSampler
    gsamp = GraphedSampler(model, ddev, max_len=DECODE_L)
    for _ in range(2):
        gsamp.replay()
    barrier()
    e0.record()
    for _ in range(dsteps):
        ids_graph = gsamp.replay()[0]
    e1.record()
    barrier()
    dec_graph_ms = max_over_ranks(e0.elapsed_time(e1)) / dsteps
    dec_graph_equal = bool(torch.equal(ids_graph, decode_step(ddev)[0]))
    del gsamp
    dpipe = HostPipeline(lambda b: decode_step(b)[0], dhost, dev)
    dpipe.run(dhost for _ in range(2))
    barrier()
    n_e2e_dec = 4 * dsteps      # (the first batch's 432 MB upload cannot overlap anything: amortised over 20 batches instead of 5)
    t0 = time.perf_counter()
    all_ids = dpipe.run(dhost for _ in range(n_e2e_dec))
    torch.cuda.synchronize()
    e2e_dec_s = max_over_ranks(time.perf_counter() - t0) / n_e2e_dec
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
    ru = (ctypes.c_longlong * 4)()
    lib.aa_debug_refine_units(ctypes.cast(ru, ctypes.c_void_p), 1)
    ids_refine = decode_step(ddev)[0]
    refine_pairs = lib.aa_debug_refine_pairs(1) / float(DECODE_B * DECODE_L)
    lib.aa_debug_refine_units(ctypes.cast(ru, ctypes.c_void_p), 1)
    refine_units = {"cta_units_per_step": ru[0] / float(DECODE_L), "warp_units_per_step": ru[1] / float(DECODE_L),
                    "tiles_by_cta_units_per_step": ru[2] / float(DECODE_L), "tiles_by_warp_units_per_step": ru[3] / float(DECODE_L)}
    # near-ties of config 3, counted on the device: the 3xTF32 projection of EVERY logit (refinement off, logits returned) gives
    # the top-1 / top-2 gap of each (image, step); ids of the default path must equal its arg-max wherever the gap is not tiny
    near = None
    if not args.quick:
        stage("near-tie census (config 3)")
        lib.aa_debug_set_decode_argmax_refine(0)
        ids_full, _, _, logits = F_aa.greedy_decode(model.decoder.weights(), ddev["V"], ddev["v_g"], ddev["h0"], ddev["c0"], DECODE_L,
                                                    return_logits=True)
        lib.aa_debug_set_decode_argmax_refine(2)
        top2 = torch.topk(logits, 2, dim=-1).values                    # [L, B, 2]
        gap = (top2[..., 0] - top2[..., 1]).t()                        # [B, L]
        differ = ids_refine != ids_full
        first = differ.float().cumsum(1).cumsum(1) == 1               # first differing position of a row (later ones follow from it)
        near = {"positions": int(gap.numel()), "gap_lt_1e-4": int((gap < 1e-4).sum()), "gap_lt_1e-5": int((gap < 1e-5).sum()),
                "rows_whose_ids_differ_refine_vs_full_projection": int(differ.any(1).sum()),
                "first_differences_with_gap_ge_1e-4": int((first & (gap >= 1e-4)).sum()),
                "note": "gap = top-1 minus top-2 logit of the fp32-accurate projection of every column; a first difference at a gap >= 1e-4 "
                        "would be a parity failure (tests/test_gpu_parity.py applies the same rule against the fp64 oracle)"}
        del logits, top2, gap

    eager_ref = small = beam = None
    if not args.quick:
        stage("beam search (config 4)")
        beam = measure_beam(torch, dev, model, dims, barrier, max_over_ranks, n_gpus, rank)
        if world == 1:
            stage("small-batch decode: eager / CUDA graph / persistent kernel")
            small = measure_decode_small(torch, dev, model, dims, peaks)
            stage("eager-PyTorch reference on this GPU")
            eager_ref = gpu_eager_baseline(torch, dev, dims, lengths)
        barrier()

    check = None
    if world > 1:
        stage("multi-GPU correctness (dp_check)")
        check = dp_check(torch, dist, dev, model, dims, world, rank)

    cfg5 = None
    if not args.quick and not args.no_config5:
        stage("config 5 (cfgB data-parallel training)")
        try:
            cfg5 = measure_config5(torch, dist, dev, world, rank, barrier, max_over_ranks, peaks, args.steps)
        except Exception as e:      # never lose the headline line to the optional section
            cfg5 = {"failed": "%s: %s" % (type(e).__name__, str(e)[:300])}
            sys.stderr.write("config 5 failed on rank %d: %s\n" % (rank, cfg5["failed"]))

    widened = None
    if world == 1 and not args.no_widened and not args.quick:
        stage("widened rows (encoder heads, baseline decoder)")
        widened = measure_widened(torch, dev, dims, peaks, args.steps)

    stage("report")
    if world > 1:
        dist.barrier()
    if rank != 0:
        leave(world, 0 if (check is None or check["ok_rank0"]) else 3)
        return

    models = kernel_models(dims, TRAIN_B, TRAIN_T, DECODE_B)
    k_train = rooflines(train_report, models, peaks)
    k_dec = rooflines(dec_report, models, peaks)
    # dominant kernel of the headline step: the tag with the largest total time among ALL tags of the per-kernel pass
    dom = max(k_train, key=lambda t: k_train[t]["ms_total"])
    roof = {kk: k_train[dom].get(kk) for kk in ("bound", "achieved", "peak", "unit", "frac", "traffic")}
    roof["kernel"] = dom
    if dom.startswith("lstm_seq"):
        roof["note"] = ("latency-bound: %d dependent recurrence steps of a [%d x %d] x [%d x %d] contraction inside one launch (8 clusters "
                        "of 16 CTAs, weights resident in tensor memory, state exchanged through distributed shared memory: nothing "
                        "of a step touches HBM); the HBM-bound kernels of this step and of decoding are listed under kernels" %
                        (TRAIN_T, TRAIN_B, 4 * dims.H, 4 * dims.H, dims.H))
    roof["peak_source"] = peaks["source"]
    serial_ms = max(sum(v["ms_total"] for v in k_train.values()), 1e-9)
    roof["share_of_step"] = k_train[dom]["ms_total"] / serial_ms
    roof["launches_per_step"] = k_train[dom]["launches"] / max(args.steps, 1)
    # whole-step roofline: algorithmic flops of fwd+bwd (47.4 MFLOP per token at cfgA) over the measured step
    alg_flops = 47.4e6 * TRAIN_B * TRAIN_T
    sus = peaks["bf16_tflops_sustained"] or peaks["bf16_tflops"]
    n_real = int(sum(lengths))
    whole = {"bound": "tensor", "algorithmic_tflops": alg_flops / (train_ms * 1e-3) / 1e12, "peak": sus, "unit": "TFLOP/s",
             "frac": alg_flops / (train_ms * 1e-3) / 1e12 / sus, "peak_kind": "sustained bf16 (MEASURED_PEAKS.json)",
             "serialised_kernel_ms_per_step": serial_ms / max(args.steps, 1),
             "note": "nominal flops of all B*T positions; the packed projection skips the padded ones (see tokens_per_step)"}
    # whole decode step against the per-launch algorithmic bytes of SURVEY 8d (128 588 B per image and step)
    dec_alg = 128588.0 * DECODE_B * DECODE_L
    dec_whole = {"bound": "hbm", "achieved": dec_alg / (dec_ms * 1e-3) / 1e9, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                 "frac": dec_alg / (dec_ms * 1e-3) / 1e9 / peaks["hbm_gbs"]}

    out = {
        "metric": METRIC, "value": train_tok, "unit": "tokens/s", "n_gpus": n_gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": train_ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": args.precision, "data": "synthetic",
        "config": workload_config(n_gpus),
        "cuda_graph": bool(stepper is not None),
        **({"dp": "4 gradient buckets, sum all-reduce by %s, %s the backward (%s launches)" %
            (trainer.engine, "overlapped with" if args.overlap else "after", dp_mode),
            "dp_env": {k: os.environ.get(k) for k in ("AA_AR_BLOCKS", "AA_AR_THREADS", "AA_DP_P2P", "AA_AR_MULTICAST", "AA_DP_TAIL_BUCKETS") if os.environ.get(k)}}
           if dp_mode is not None else {}),
        "timing": spread,
        "tokens_per_step": {"positions_B_x_T": TRAIN_B * TRAIN_T, "packed_rows_sum_lengths": n_real,
                            "value_over_packed_rows": n_real * n_gpus / (train_ms * 1e-3),
                            "note": "the metric counts B*T decoder positions per step (SURVEY 8d) for both arms; the vocabulary projection, CE and "
                                    "their backward run over the packed rows only (pack_padded_sequence keeps those, Q13), the reference computes all B*T"},
        "clocks": clocks,
        "e2e": {"value": TRAIN_B * TRAIN_T * n_gpus / e2e_train_s, "unit": "tokens/s", "h2d_bytes_per_step": h2d_train,
                "d2h_bytes_per_step": 4, "ms_per_step": e2e_train_s * 1e3, "steps_timed": int(n_e2e)},
        "gpu_launches": int(launches),
        "gpu_launches_note": "kernels of libadaptive_sm100 per block of --steps steps (counted on one eager step; the timed blocks replay them as a CUDA graph)",
        "roofline": roof,
        "roofline_whole_step": whole,
        "kernels": {"train": k_train, "decode": k_dec},
        "decode": {"workload": "BASELINE config 3: greedy sampler, batch %d per GPU, max_len %d, fp32; V (411 MB) > L2" % (DECODE_B, DECODE_L),
                   "value": dec_tok, "unit": "tokens/s", "ms_per_step": dec_ms, "steps": dsteps, "gpu_launches": int(dec_launches),
                   "cuda_graph": {"value": DECODE_B * DECODE_L * n_gpus / (dec_graph_ms * 1e-3), "unit": "tokens/s", "ms_per_step": dec_graph_ms,
                                  "launches": 1, "ids_equal_eager": dec_graph_equal,
                                  "note": "graphs.GraphedSampler.replay(): the whole max_len-step loop as one graph launch, inputs resident in its static buffers"},
                   "e2e": {"value": DECODE_B * DECODE_L * n_gpus / e2e_dec_s, "unit": "tokens/s", "h2d_bytes_per_step": h2d_dec,
                           "d2h_bytes_per_step": d2h_dec, "ms_per_step": e2e_dec_s * 1e3, "steps_timed": int(n_e2e_dec)},
                   "precision": model.decoder.decode_precision,
                   "vocab_argmax": {"method": "one bf16 tensor-core pass (maxima per 16 columns) + exact fp32 recompute of the tiles that can "
                                              "hold the row maximum under a rigorous error bound (vocab_refine.cu); ids equal an exact fp32 projection's",
                                    "tiles_refined_per_row_and_step": refine_pairs, "tiles_per_row": (dims.Vc + 15) // 16,
                                    "refine_units": refine_units},
                   "near_ties": near,
                   "roofline_fused_step": k_dec.get("dec_step_fused"),
                   "roofline_whole_step": dec_whole},
    }
    if beam is not None:
        out["beam"] = beam
    if small is not None:
        out["decode_small"] = small
    if eager_ref is not None:
        out["gpu_eager_baseline"] = eager_ref
        try:
            out["gpu_eager_baseline"]["ours_over_eager"] = {
                "train_vs_fp32": train_tok / n_gpus / eager_ref["train_config2"]["fp32"]["value"],
                "train_vs_best": train_tok / n_gpus / max(v["value"] for v in eager_ref["train_config2"].values() if "value" in v),
                "greedy_vs_fp32": dec_tok / n_gpus / eager_ref["greedy_config3"]["fp32"]["value"]}
        except (KeyError, ValueError, TypeError):
            pass
    if check is not None:
        out["dp_check"] = check
    if cfg5 is not None:
        out["config5"] = cfg5
    if widened is not None:
        out["widened"] = widened
    if n_gpus == 1:
        cores = os.cpu_count() or 1
        stage("cpu baseline (reference modules on the host cores)")
        step, kind = cpu_train_step_fn(TRAIN_B, TRAIN_T, dims)
        ncpu = 20
        sec = time_cpu(step, ncpu, 2)
        what = "the unmodified reference modules (oracle/_ref, torch CPU kernels)" if kind == "reference" else "the numpy oracle port"
        out["cpu_baseline"] = {"value": TRAIN_B * TRAIN_T / sec, "unit": "tokens/s", "cores": cores, "kind": kind,
                               "sample": "%d full steps of the same workload (B=%d, T=%d) on %s, %.1f s" % (ncpu, TRAIN_B, TRAIN_T, what, sec * ncpu)}
        if not args.quick:
            out["cpu_baseline"]["config1"] = cpu_config1()
    emit(out)
    rc = 0
    if check is not None and not check["ok"]:
        sys.stderr.write("dp_check FAILED: %s\n" % json.dumps(check))
        rc = 3
    leave(world, rc)


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
    ap.add_argument("--blocks", type=int, default=0, help="timed blocks of --steps steps (0 = enough for >= 50 blocks and >= 1 s)")
    ap.add_argument("--quick", action="store_true", help="headline + decode only: skip the eager-GPU reference arm, small-batch decode, beam, config 5, widened rows")
    ap.add_argument("--dp-path", action="store_true", help="N=1: run the data-parallel trainer's code path (no collectives) instead of the autograd one")
    ap.add_argument("--no-config5", action="store_true", help="skip the BASELINE config 5 section")
    ap.add_argument("--no-widened", action="store_true", help="skip the extra measurements of the SURVEY 8f rows (N=1 only)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
