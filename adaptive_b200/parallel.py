"""Multi-GPU drivers: one process per GPU (``torchrun``), ``torch.distributed`` for the plumbing.

The reference's only multi-GPU scheme is ``nn.DataParallel`` around the encoder and the adaptive
block (``baseline_attention.py:184-187,215-218``; ``adaptive_attention.py:178-181``): single
process, parameters re-broadcast on every forward, logits gathered to GPU 0, the LSTM loop not
parallelised at all.  It is not reproduced.  Here:

* **decoding** shards the images into contiguous ranges, one per rank, with *no* data-path
  collective (images are independent; beams never cross images).  Gathering the ids at the end
  is control plane.
* **training** is data parallel over the batch: weights replicated, one exchange step — a
  sum all-reduce of the 13 decoder gradients.  The gradients live in four flat buckets in the
  order in which the hand-written backward finishes them (``AA_BUCKET_*`` in
  ``include/adaptive_b200.h``: mlp -> attention/sentinel -> LSTM -> embedding);
  ``aa_decoder_backward_hooked`` reports each bucket as soon as its last kernel is enqueued and
  the bucket's all-reduce starts on the communication stream while the rest of the backward
  (notably the whole BPTT) is still running.  The loss and its gradient are normalised by the
  GLOBAL packed-token count, so the reduced gradients equal those of the reference's
  single-process mean cross-entropy (``train.py:63,208``) on the concatenated batch.

Everything that does not touch the device (sharding arithmetic, bucket layout, the reducer) works
with the ``gloo`` backend on CPU tensors and is tested that way (``tests/test_parallel_cpu.py``).
"""
from __future__ import annotations

import ctypes
import os
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist

from ._lib import WEIGHT_FIELDS

# gradient buckets in backward-ready order (must match AA_BUCKET_* in include/adaptive_b200.h)
BUCKETS: Tuple[Tuple[str, ...], ...] = (
    ("mlp_w", "mlp_b"),
    ("att_wv", "att_wg", "att_ws", "att_wh", "sen_wx", "sen_wh"),
    ("w_ih", "w_hh", "b_ih", "b_hh"),
    ("embed",),
)
BUCKET_OF: Dict[str, int] = {f: b for b, fs in enumerate(BUCKETS) for f in fs}
EVENT_BPTT_DONE = len(BUCKETS)        # AA_EVENT_BPTT_DONE: index of the "recurrence enqueued" event / callback
N_READY_EVENTS = len(BUCKETS) + 1     # AA_NUM_READY_EVENTS
assert sorted(BUCKET_OF) == sorted(WEIGHT_FIELDS)


# ---------------------------------------------------------------------------------------------
# sharding arithmetic (decode)
# ---------------------------------------------------------------------------------------------
def shard_range(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous near-equal split of ``n`` images: the first ``n % world`` ranks get one more."""
    if not (0 <= rank < world):
        raise ValueError("rank %d outside world of %d" % (rank, world))
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_encoded(encoded, rank: int, world: int):
    """Slice an encoded batch ``(V, v_g, (h0, c0))`` to this rank's image range.
    States may be ``[1,B,H]`` (nn.LSTM layout), ``[B,1,H]`` (what the reference encoder returns) or ``[B,H]``."""
    V, v_g, states = encoded
    B = V.shape[0]
    lo, hi = shard_range(B, rank, world)

    def cut(s):
        if s is None:
            return None
        if s.dim() == 3 and s.shape[0] == 1 and s.shape[1] == B:
            return s[:, lo:hi]
        return s[lo:hi]

    st = None if states is None else tuple(cut(s) for s in states)
    return V[lo:hi], v_g[lo:hi], st


def gather_rows(local: torch.Tensor, n_total: int, group=None) -> torch.Tensor:
    """Concatenate every rank's rows in rank order (ranges from ``shard_range``): the end-of-decode
    collection of ids.  Uneven shards are padded to the largest one for the all_gather."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        return local
    sizes = [shard_range(n_total, r, world) for r in range(world)]
    most = max(hi - lo for lo, hi in sizes)
    pad = local.new_zeros((most,) + tuple(local.shape[1:]))
    pad[: local.shape[0]] = local
    out = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(out, pad, group=group)
    return torch.cat([o[: hi - lo] for o, (lo, hi) in zip(out, sizes)], dim=0)


def sharded_sampler(model, encoded, max_len: int = 30, beam: int = 0, gather: bool = False, group=None):
    """``Encoder2Decoder.sampler`` over this rank's contiguous image range (no collective on the
    data path).  Returns the local ``(ids, attention, Beta)``; with ``gather=True`` the ids of
    all ranks, in image order, instead of the local ones."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    n_total = encoded[0].shape[0]
    local = shard_encoded(encoded, rank, world)
    out = model.beam_sampler(local, beam=beam, max_len=max_len) if beam >= 1 else model.sampler(local, max_len=max_len)
    if gather:
        return (gather_rows(out[0], n_total, group),) + tuple(out[1:])
    return out


# ---------------------------------------------------------------------------------------------
# gradient buckets + reducer (training)
# ---------------------------------------------------------------------------------------------
class SymmetricBuffer:
    """One fp32 allocation per rank of identical size, mapped into every peer of the group (and, where the fabric supports it,
    behind one NVLS multicast address): the memory the hand-written all-reduce kernels (``csrc/allreduce.cu``) work on.
    The mappings come from ``torch.distributed._symmetric_memory`` (cuMem allocation + handle exchange: plumbing); the tail of
    the allocation holds the kernels' cross-GPU flags, zero-filled once here."""

    def __init__(self, n_floats: int, device, group=None):
        import torch.distributed._symmetric_memory as symm

        from . import _lib

        lib = _lib.load()
        self.group = group if group is not None else dist.group.WORLD
        self.world, self.rank = dist.get_world_size(self.group), dist.get_rank(self.group)
        if self.world > 8:
            raise RuntimeError("peer-memory all-reduce is built for one NVSwitch box (<= 8 ranks)")
        flag_floats = int(lib.aa_allreduce_flag_bytes()) // 4
        self.n = (int(n_floats) + 63) // 64 * 64
        # layout: [fp32 payload (n) | bf16 staging of the payload (n/2 floats) | flags]
        self.stage_offset_bytes = self.n * 4
        self.flag_offset_bytes = self.n * 4 + self.n * 2
        self.buf = symm.empty(self.n + self.n // 2 + flag_floats, dtype=torch.float32, device=device)
        self.buf.zero_()
        torch.cuda.synchronize(device)
        self.handle = symm.rendezvous(self.buf, self.group)
        self.peer_ptrs = [int(p) for p in self.handle.buffer_ptrs]
        mc = getattr(self.handle, "multicast_ptr", 0) or 0
        self.multicast_ptr = int(mc) if os.environ.get("AA_AR_MULTICAST", "1") != "0" else 0
        self._peer_arr = (ctypes.c_void_p * self.world)(*self.peer_ptrs)
        self.lib = lib
        # CTAs per exchange kernel: measured best 16 on 2 GPUs (495 vs 518 us per step), 8 on 8 GPUs (468 vs 476)
        self.max_blocks = int(os.environ.get("AA_AR_BLOCKS", "16" if self.world <= 2 else "8"))
        dist.barrier(self.group)          # every rank's flags are zero before anyone signals

    @property
    def payload(self) -> torch.Tensor:
        return self.buf[: self.n]

    def all_reduce_(self, view: torch.Tensor, channel: int, stream=None, bf16: bool = False):
        """In-place sum over the ranks of ``view`` (a contiguous slice of ``payload`` starting on a 32-byte boundary),
        asynchronous on ``stream`` (default: the current one).  ``bf16``: the bucket crosses NVLink as bf16 (fp32 accumulation),
        half the bytes -- the exchange of the mixed-precision training path."""
        from ._lib import check

        off = (view.data_ptr() - self.buf.data_ptr()) // 4
        n = (view.numel() + 7) // 8 * 8
        if off % 8 or off < 0 or off + n > self.n:
            raise ValueError("all_reduce_: view must be a 32-byte aligned slice of the symmetric payload")
        st = stream if stream is not None else torch.cuda.current_stream(self.buf.device)
        mc = ctypes.c_void_p(self.multicast_ptr) if self.multicast_ptr else None
        with torch.cuda.device(self.buf.device):
            if bf16:
                check(self.lib.aa_allreduce_sum_bf16(self._peer_arr, mc, self.flag_offset_bytes, self.stage_offset_bytes, self.rank,
                                                     self.world, off, n, channel, self.max_blocks, ctypes.c_void_p(st.cuda_stream)),
                      "aa_allreduce_sum_bf16")
            else:
                check(self.lib.aa_allreduce_sum_f32(self._peer_arr, mc, self.flag_offset_bytes, self.rank, self.world, off, n, channel,
                                                    self.max_blocks, ctypes.c_void_p(st.cuda_stream)), "aa_allreduce_sum_f32")

    def link_bytes(self, n_elems: int, elem_bytes: int = 4) -> int:
        """NVLink bytes one GPU sends (= receives) for one all-reduce of ``n_elems``: its (W-1)/W share of the bucket once for
        the reduction and once for the broadcast."""
        return int(2 * elem_bytes * n_elems * (self.world - 1) / self.world)


def try_symmetric_buffer(n_floats: int, device, group=None) -> Optional["SymmetricBuffer"]:
    """``SymmetricBuffer`` if every rank of the group can build one (same answer on all ranks), else None (NCCL is used)."""
    if not dist.is_initialized() or dist.get_world_size(group) < 2 or os.environ.get("AA_DP_P2P", "1") == "0":
        return None
    sb, ok = None, 1.0
    try:
        sb = SymmetricBuffer(n_floats, device, group)
    except Exception as e:      # no symmetric-memory support in this build / container / fabric
        import sys

        sys.stderr.write("adaptive_b200.parallel: symmetric memory unavailable on rank %d (%s: %s); using NCCL\n"
                         % (dist.get_rank(group), type(e).__name__, str(e)[:300]))
        ok = 0.0
    flag = torch.tensor([ok], device=device)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)
    return sb if float(flag.item()) > 0 else None


class GradBuckets:
    """One flat fp32 buffer per bucket with a view per parameter (``aa_weights`` field order).

    The buckets are carved out of ONE allocation, in bucket order, followed by a slot for the loss scalar: the last two
    buckets (LSTM and embedding, final within a few microseconds of each other at the end of the backward) and the loss are
    then one contiguous range, ``tail``, and go through one collective instead of three -- each NCCL all-reduce of the
    captured step costs ~45 us whatever its size (``profiles/r01_v47_timeline_n2.txt``)."""

    def __init__(self, shapes: Dict[str, Sequence[int]], device, dtype=torch.float32, symmetric_group=None, want_symmetric: bool = False):
        self.flat: List[torch.Tensor] = []
        self.symm: Optional[SymmetricBuffer] = None
        self.views: Dict[str, torch.Tensor] = {}
        layout, total = [], 0
        for fields in BUCKETS:
            sizes = [int(torch.Size(shapes[f]).numel()) for f in fields]
            # 64-element (256 B) alignment of every view: the kernels store 128-bit vectors
            offs, start = [], total
            for n in sizes:
                offs.append(total)
                total += (n + 63) // 64 * 64
            layout.append((fields, sizes, offs, start, total))
        if want_symmetric and dtype == torch.float32:
            self.symm = try_symmetric_buffer(total + 64, device, symmetric_group)
        if self.symm is not None:
            self.all = self.symm.payload[: total + 64]        # gradients are written straight into peer-mapped memory
        else:
            self.all = torch.zeros(total + 64, device=device, dtype=dtype)
        for fields, sizes, offs, start, end in layout:
            self.flat.append(self.all[start:end])
            for f, o, n in zip(fields, offs, sizes):
                self.views[f] = self.all[o:o + n].view(*shapes[f])
        self.loss = self.all[total:total + 1].view(())           # summed over ranks together with ``tail``
        self.tail_first = max(0, len(BUCKETS) - int(os.environ.get("AA_DP_TAIL_BUCKETS", "2")))   # first bucket of the merged range
        self.tail = self.all[layout[self.tail_first][3]:total + 64]
        self.tail_grads = self.all[layout[self.tail_first][3]:total]       # the same range without the loss slot
        self.loss_slot = self.all[total:total + 64]

    def ordered(self) -> Tuple[torch.Tensor, ...]:
        return tuple(self.views[f] for f in WEIGHT_FIELDS)

    def nbytes(self) -> int:
        return sum(b.numel() * b.element_size() for b in self.flat)


class BucketReducer:
    """Sum all-reduce of the buckets, each started as soon as it is reported ready.

    On CUDA the collective is issued under ``comm_stream`` after that stream has been made to
    wait for ``ready_event`` (recorded by the backward on whichever of its lanes finished the
    bucket), so it overlaps the remaining backward kernels; ``finish`` makes the caller's
    current stream wait for all of them.  On CPU (gloo) it degenerates to async all-reduces."""

    def __init__(self, buckets: GradBuckets, group=None, average: bool = False, bf16_exchange: bool = False):
        self.buckets, self.group, self.average = buckets, group, average
        self.bf16_exchange = bool(bf16_exchange) and buckets.symm is not None
        self.bytes_reduced = 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.cuda = buckets.flat[0].is_cuda
        self.works: List = []
        self.order: List[int] = []
        if self.cuda:
            # HIGHEST priority: the exchange kernels are a handful of CTAs that every rank must have RUNNING before any of them can
            # finish; on a default-priority lane they queue behind the pending CTAs of whatever large grid the backward launched
            # before them (tools/probe_overlap.py: a 145 us exchange took 480 us next to a 345 us copy kernel -- it simply ran after it)
            prio = int(os.environ.get("AA_DP_COMM_PRIORITY", "-100"))
            self.comm_stream = torch.cuda.Stream(device=buckets.flat[0].device, priority=prio)
            # the small attention / sentinel bucket gets a lane of its own: behind the 20 MB vocabulary bucket on ONE lane it waited
            # ~50 us for a 25 us exchange (profiles/r02_timeline_n8_v2.txt)
            self.comm_stream2 = torch.cuda.Stream(device=buckets.flat[0].device, priority=prio)
            self.events = [torch.cuda.Event() for _ in range(N_READY_EVENTS)]      # four buckets + "recurrence enqueued"
            for e in self.events:      # the raw cudaEvent_t exists only after a first record
                e.record()
        else:
            self.comm_stream, self.events = None, [None] * N_READY_EVENTS
        # "bucket" = start every bucket's exchange as soon as it is final (default); "after_bptt" = the first exchange (vocabulary +
        # attention buckets, one launch) starts when the backward's recurrence has finished instead of next to it (measured slower:
        # 530 vs 495 us per step on 2 GPUs)
        self.schedule = os.environ.get("AA_DP_SCHEDULE", "bucket")

    def event_handles(self):
        """``void* ready_events[AA_NUM_READY_EVENTS]`` for ``aa_decoder_backward_hooked`` (CUDA only)."""
        arr = (ctypes.c_void_p * N_READY_EVENTS)()
        for i, e in enumerate(self.events):
            arr[i] = e.cuda_event
        return arr

    def start(self):
        self.works, self.order = [], []
        self._deferred: List[int] = []
        self._early: List[int] = []
        self.bytes_reduced = 0
        self._lanes_used = set()

    def _all_reduce(self, flat: torch.Tensor, buckets: Sequence[int], lane=None):
        self.bytes_reduced += flat.numel() * (2 if self.bf16_exchange else flat.element_size())
        if self.cuda and self.buckets.symm is not None:
            # hand-written all-reduce over peer-mapped memory (csrc/allreduce.cu), one channel per bucket
            if lane is None:
                lane = self.comm_stream2 if tuple(buckets) == (1,) else self.comm_stream
            self._lanes_used.add(lane)
            with torch.cuda.stream(lane):
                for b in buckets:
                    lane.wait_event(self.events[b])
                self.buckets.symm.all_reduce_(flat, channel=min(buckets[0], 3), stream=lane, bf16=self.bf16_exchange)
        elif self.cuda:
            with torch.cuda.stream(self.comm_stream):
                for b in buckets:
                    self.comm_stream.wait_event(self.events[b])
                self.works.append(dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group, async_op=True))
        else:
            self.works.append(dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group, async_op=True))

    def on_ready(self, bucket: int):
        if bucket == EVENT_BPTT_DONE:
            if self.world > 1 and self.schedule == "after_bptt" and self._early:
                # buckets 0..1 are contiguous in the allocation: one exchange, started when the recurrence is done
                first, last_b = min(self._early), max(self._early)
                lo = self.buckets.flat[first].data_ptr()
                hi = self.buckets.flat[last_b].data_ptr() + self.buckets.flat[last_b].numel() * 4
                span = self.buckets.all[(lo - self.buckets.all.data_ptr()) // 4:(hi - self.buckets.all.data_ptr()) // 4]
                early, self._early = tuple(self._early) + (EVENT_BPTT_DONE,), []
                self._all_reduce(span, early, lane=self.comm_stream)
            return
        self.order.append(bucket)
        if self.world == 1:
            return
        if self.schedule == "after_bptt" and self.cuda and self.buckets.symm is not None and bucket < self.buckets.tail_first:
            self._early.append(bucket)
            return
        last = len(BUCKETS) - 1
        tf = self.buckets.tail_first
        if tf <= bucket < last:                      # reduced together with the last bucket and the loss slot
            self._deferred.append(bucket)
        elif bucket == last and self._deferred == list(range(tf, last)):
            self._deferred = []
            # (bf16 exchange: the loss scalar must not be rounded -- it went through reduce_loss() in fp32 right after the loss kernel)
            self._all_reduce(self.buckets.tail_grads if self.bf16_exchange else self.buckets.tail, tuple(range(tf, last + 1)))
        else:
            self._all_reduce(self.buckets.flat[bucket], (bucket,))

    def reduce_loss(self, ready_event=None):
        """bf16 exchange only: sum the loss scalar over the ranks in fp32 on the second lane, as soon as the loss kernel is done
        (long before the backward ends: fully hidden)."""
        if not (self.bf16_exchange and self.world > 1):
            return
        lane = self.comm_stream2
        self._lanes_used.add(lane)
        with torch.cuda.stream(lane):
            if ready_event is not None:
                lane.wait_event(ready_event)
            self.buckets.symm.all_reduce_(self.buckets.loss_slot, channel=3, stream=lane, bf16=False)

    def finish(self):
        for b in getattr(self, "_early", []) + getattr(self, "_deferred", []):       # (deferred buckets whose trigger never came)
            self._all_reduce(self.buckets.flat[b], (b,))
        self._deferred, self._early = [], []
        if self.cuda and self.world > 1:
            with torch.cuda.stream(self.comm_stream):
                for w in self.works:
                    w.wait()                      # comm_stream waits for NCCL's stream
                if self.average:
                    for flat in self.buckets.flat:
                        flat.div_(self.world)
            torch.cuda.current_stream().wait_stream(self.comm_stream)
            for lane in getattr(self, "_lanes_used", ()):
                if lane is not self.comm_stream:
                    torch.cuda.current_stream().wait_stream(lane)
        else:
            for w in self.works:
                w.wait()
            if self.average and self.world > 1:
                for flat in self.buckets.flat:
                    flat.div_(self.world)
        self.works = []


def global_token_count(lengths: Sequence[int], group=None) -> int:
    """Sum over ranks of this rank's packed-row count ``sum(lengths)``: the denominator of the global mean CE.

    One small host-side collective per call, entered by EVERY rank on EVERY call: the value depends on all ranks' lengths,
    so nothing here may be cached under a rank-local key (a rank that hit such a cache would skip a collective its peers
    enter).  A loop that cannot afford it (a captured CUDA graph) computes the count once for its frozen lengths and hands
    it to ``DataParallelTrainer.step(..., denom=...)``, which is what ``GraphedDPStep`` does."""
    n_local = int(sum(int(x) for x in lengths))
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return n_local
    counts: List[Optional[int]] = [None] * dist.get_world_size(group)
    dist.all_gather_object(counts, n_local, group=group)
    return int(sum(counts))


class DataParallelTrainer:
    """Data-parallel training step of the decoder over the C ABI, no autograd in the loop.

    ``step(encoded, captions, lengths, targets) -> loss`` runs forward -> packed rows -> cross-entropy
    (global-count normalised) -> hooked backward with bucketed, overlapped all-reduce, and leaves
    the reduced gradients in ``p.grad`` of ``model.decoder``'s parameters (views of the flat
    buckets; the caller's optimizer / ``clip_grad_norm_(model.decoder.LSTM.parameters(), 5)`` of
    ``train.py:213-214`` work on them as usual).  ``loss`` is the global mean CE (device scalar).
    With one process it is simply the fused single-GPU training step."""

    def __init__(self, model, group=None, overlap: bool = True):
        from . import _lib

        self.lib = _lib.load()
        self.model, self.group, self.overlap = model, group, overlap
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.weights = model.decoder.weights()
        if any(t is None for t in self.weights):
            raise ValueError("DataParallelTrainer buckets the 13 gradients of the adaptive decoder; the sentinel-less baseline model "
                             "(adaptive_b200.baseline) trains through Encoder2Decoder.forward on one GPU per process and all-reduces "
                             "its p.grad tensors itself")
        dev = self.weights[0].device
        if dev.type != "cuda":
            raise RuntimeError("DataParallelTrainer needs the decoder on a CUDA device (no CPU fallback)")
        self.device = dev
        self.buckets = GradBuckets({f: tuple(t.shape) for f, t in zip(WEIGHT_FIELDS, self.weights)}, dev, symmetric_group=group,
                                   want_symmetric=self.world > 1)
        # the mixed-precision path exchanges its gradients as bf16 (fp32 accumulation); the exact path keeps fp32 on the wire
        want16 = os.environ.get("AA_DP_BF16", "1") != "0" and getattr(model.decoder, "precision", "fp32") == "bf16"
        self.reducer = BucketReducer(self.buckets, group, bf16_exchange=want16)
        self.engine = "peer-memory kernels (%s, %s on the wire)" % ("NVLS multimem" if self.buckets.symm.multicast_ptr else "two-shot peer loads",
                                                                     "bf16" if self.reducer.bf16_exchange else "fp32") \
            if self.buckets.symm is not None else ("nccl" if self.world > 1 else "none")
        for p, g in zip(self.weights, self.buckets.ordered()):
            p.grad = g
        self._cb_type = ctypes.CFUNCTYPE(None, ctypes.c_int, ctypes.c_void_p)
        self._cb = self._cb_type(lambda bucket, _user: self.reducer.on_ready(int(bucket)))
        self._bufs: Dict[tuple, dict] = {}
        self._loss_ready = torch.cuda.Event()

    def _buffers(self, B, T, k, H, E, Vc, a, n_rows, prec):
        from . import functional as F_aa

        key = (B, T, k, n_rows, prec)
        hit = self._bufs.get(key)
        if hit is None:
            dev = self.device
            d = F_aa.make_dims(B, T, k, H, E, Vc, a, prec)
            f32 = lambda *s: torch.empty(*s, device=dev, dtype=torch.float32)
            u8 = lambda n: torch.empty(n, device=dev, dtype=torch.uint8)
            hit = {
                "d": d, "alpha": f32(B, T, k), "beta": f32(B, T, 1), "hT": f32(B, H), "cT": f32(B, H),
                "saved": u8(self.lib.aa_decoder_saved_bytes(ctypes.byref(d))),
                "scratch": u8(self.lib.aa_decoder_bwd_scratch_bytes(ctypes.byref(d))),
                "packed": f32(n_rows, Vc), "dpacked": f32(n_rows, Vc),
                "dpacked16": torch.empty(n_rows, Vc, device=dev, dtype=torch.bfloat16),
                "loss": self.buckets.loss, "dV": f32(B, k, H), "dvg": f32(B, E),
                "dh0": f32(B, H), "dc0": f32(B, H),
            }
            if len(self._bufs) > 8:
                self._bufs.clear()
            self._bufs[key] = hit
        return hit

    def step(self, encoded, captions: torch.Tensor, lengths: Sequence[int], targets: torch.Tensor,
             denom: Optional[int] = None) -> torch.Tensor:
        """``denom``: the GLOBAL packed-row count if the caller already knows it (same value on every rank); by default it is
        gathered from all ranks on every call (``global_token_count``)."""
        from . import _lib as _lib_mod
        from . import functional as F_aa
        from ._lib import AAWeightGrads, check

        lib = self.lib
        V, v_g, states = encoded
        V, v_g = F_aa._f32c(V), F_aa._f32c(v_g)
        B, k, H = V.shape
        E, T = v_g.shape[1], captions.shape[1]
        w = self.weights
        Vc, a = w[0].shape[0], w[7].shape[0]
        h0 = c0 = None
        if states is not None:
            h0, c0 = F_aa._states2d(states[0], B, H), F_aa._states2d(states[1], B, H)
        captions = captions.to(torch.int64).contiguous()
        targets = targets.to(torch.int64).contiguous()
        prec = F_aa.PRECISIONS[self.model.decoder.precision]
        row_index, _ = F_aa.cached_row_index(lengths, T, self.device)
        n_rows = row_index.numel()
        if denom is None:
            denom = global_token_count(lengths, self.group)
        b = self._buffers(B, T, k, H, E, Vc, a, n_rows, prec)
        d, st, P = b["d"], F_aa._stream(self.device), F_aa._ptr
        ws = F_aa.weights_struct(w)
        gs = AAWeightGrads()
        for name, t in zip(WEIGHT_FIELDS, self.buckets.ordered()):
            setattr(gs, name, t.data_ptr())
        with torch.cuda.device(self.device):
            # forward over the packed rows only (pack_padded_sequence fused, Q13), loss + its gradient in one pass, backward
            fused = prec == _lib_mod.PREC_BF16 and F_aa.fused_loss_pays(n_rows, Vc)
            written = ctypes.c_int(0)
            if fused:      # the loss rides in the vocabulary projection's epilogue; the gradient of the logits stays inside `saved` (bf16)
                check(lib.aa_decoder_forward_loss(ctypes.byref(d), ctypes.byref(ws), P(V), P(v_g), P(captions), P(h0), P(c0), P(row_index),
                                                  n_rows, P(targets), denom, P(b["loss"]), P(b["alpha"]), P(b["beta"]), P(b["hT"]), P(b["cT"]),
                                                  P(b["saved"]), b["saved"].numel(), st), "aa_decoder_forward_loss")
            else:
                check(lib.aa_decoder_forward_packed(ctypes.byref(d), ctypes.byref(ws), P(V), P(v_g), P(captions), P(h0), P(c0), P(row_index),
                                                    n_rows, P(b["packed"]), P(b["alpha"]), P(b["beta"]), P(b["hT"]), P(b["cT"]), P(b["saved"]),
                                                    b["saved"].numel(), st), "aa_decoder_forward_packed")
                check(lib.aa_cross_entropy_mirror(P(b["packed"]), n_rows, Vc, P(targets), denom, P(b["loss"]), P(b["dpacked"]), P(b["dpacked16"]),
                                                  ctypes.byref(written), st), "aa_cross_entropy_mirror")
            self.reducer.start()
            if self.reducer.bf16_exchange and self.world > 1:
                self._loss_ready.record(torch.cuda.current_stream(self.device))
                self.reducer.reduce_loss(self._loss_ready)
            hooked = self.overlap and self.world > 1
            check(lib.aa_decoder_backward_packed(
                ctypes.byref(d), ctypes.byref(ws), P(V), P(v_g), P(captions), P(h0), P(c0), P(b["alpha"]), P(b["beta"]), P(b["saved"]),
                b["saved"].numel(), P(row_index), n_rows, None if fused else P(b["dpacked"]), None, None, None, None, ctypes.byref(gs), P(b["dV"]),
                P(b["dvg"]), P(b["dh0"]) if h0 is not None else None, P(b["dc0"]) if c0 is not None else None, P(b["scratch"]),
                b["scratch"].numel(), st, self.reducer.event_handles() if hooked else None,
                ctypes.cast(self._cb, ctypes.c_void_p) if hooked else None, None, P(b["dpacked16"]) if written.value else None),
                "aa_decoder_backward_packed")
            if self.world > 1 and not hooked:          # non-overlapped variants, after the backward
                if self.buckets.symm is not None and os.environ.get("AA_DP_SINGLE", "1") != "0":
                    # ONE exchange of all four buckets (they are contiguous in the symmetric allocation)
                    self.reducer.events[0].record()
                    whole = self.buckets.all[: self.buckets.loss_slot.data_ptr() // 4 - self.buckets.all.data_ptr() // 4] \
                        if self.reducer.bf16_exchange else self.buckets.all
                    self.reducer._all_reduce(whole, (0,))
                    self.reducer.order = list(range(len(BUCKETS)))
                else:                                  # the same buckets as the overlapped schedule, one after the other
                    for i in range(len(BUCKETS)):
                        self.reducer.events[i].record()
                        self.reducer.on_ready(i)
            self.reducer.finish()
            loss = b["loss"]      # (summed over ranks by the tail collective)
        for p, g in zip(self.weights, self.buckets.ordered()):
            p.grad = g
        self.input_grads = {"V": b["dV"], "v_g": b["dvg"], "h0": b["dh0"] if h0 is not None else None,
                            "c0": b["dc0"] if c0 is not None else None}
        return loss


class GraphedDPStep:
    """``DataParallelTrainer.step`` captured once as a CUDA graph (forward, loss, hooked backward and the
    bucketed NCCL all-reduces on their communication stream) and replayed per batch.  Shapes, caption lengths
    and parameter tensors are frozen at capture; inputs are copied into static buffers before each replay.
    Every rank must construct it at the same point (capture issues the same collectives on all ranks)."""

    KEYS = ("V", "v_g", "h0", "c0", "captions", "tgt")

    def __init__(self, trainer: DataParallelTrainer, example: Dict[str, torch.Tensor], lengths: Sequence[int], warmup: int = 3):
        self.trainer = trainer
        self.lengths = [int(x) for x in lengths]
        self.static = {k: example[k].clone() for k in self.KEYS}
        self.denom = global_token_count(self.lengths, trainer.group)     # lengths are frozen: one collective, at construction
        cur = torch.cuda.current_stream()
        side = torch.cuda.Stream()
        side.wait_stream(cur)
        with torch.cuda.stream(side):
            for _ in range(max(1, warmup)):       # one-time initialisation (attributes, side streams, NCCL) outside the capture
                self._eager()
        cur.wait_stream(side)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.loss = self._eager()

    def _eager(self):
        b = self.static
        return self.trainer.step((b["V"], b["v_g"], (b["h0"], b["c0"])), b["captions"], self.lengths, b["tgt"], denom=self.denom)

    def __call__(self, batch: Dict[str, torch.Tensor]) -> torch.Tensor:
        from . import functional as F_aa

        F_aa.copy_multi([self.static[k] for k in self.KEYS], [batch[k] for k in self.KEYS])   # one launch for the six inputs
        self.graph.replay()
        return self.loss
