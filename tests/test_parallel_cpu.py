"""Host-side logic of the multi-GPU drivers on CPU: sharding arithmetic, bucket layout, and the
bucket reducer / id gathering over a real world_size-2 ``gloo`` process group."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from adaptive_b200 import parallel as par
from adaptive_b200._lib import WEIGHT_FIELDS
from adaptive_b200.synth import CFG_A, DECODER_KEYS, Dims


def test_shard_range_partitions_exactly():
    for n in (0, 1, 7, 8, 4096, 4099):
        for world in (1, 2, 3, 8):
            rs = [par.shard_range(n, r, world) for r in range(world)]
            assert rs[0][0] == 0 and rs[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(rs, rs[1:]))
            sizes = [hi - lo for lo, hi in rs]
            assert max(sizes) - min(sizes) <= 1 and sizes == sorted(sizes, reverse=True)
    with pytest.raises(ValueError):
        par.shard_range(4, 2, 2)


def test_shard_encoded_handles_all_state_layouts():
    B, k, H, E = 5, 3, 4, 2
    V, v_g = torch.arange(B * k * H).view(B, k, H).float(), torch.arange(B * E).view(B, E).float()
    h = torch.arange(B * H).view(B, H).float()
    for st in (h, h.unsqueeze(0), h.unsqueeze(1)):
        got = [par.shard_encoded((V, v_g, (st, st)), r, 2) for r in range(2)]
        assert torch.equal(torch.cat([g[0] for g in got]), V) and torch.equal(torch.cat([g[1] for g in got]), v_g)
        cat_dim = 1 if (st.dim() == 3 and st.shape[0] == 1) else 0
        assert torch.equal(torch.cat([g[2][0] for g in got], dim=cat_dim), st)
    assert par.shard_encoded((V, v_g, None), 0, 2)[2] is None


def _shapes(dims: Dims):
    from adaptive_b200._lib import KEY_TO_FIELD
    return {KEY_TO_FIELD[k]: v for k, v in dims.shapes().items()}


def test_bucket_layout_covers_every_parameter_once():
    assert [f for fs in par.BUCKETS for f in fs].count("mlp_w") == 1
    assert sorted(f for fs in par.BUCKETS for f in fs) == sorted(WEIGHT_FIELDS)
    gb = par.GradBuckets(_shapes(CFG_A), "cpu")
    n_params = sum(int(np.prod(s)) for s in CFG_A.shapes().values())
    assert n_params == 10390849
    assert n_params * 4 <= gb.nbytes() < n_params * 4 + 13 * 256
    # views alias the flat buffers, are 256-byte aligned, and do not overlap
    for f, v in gb.views.items():
        base = gb.flat[par.BUCKET_OF[f]].data_ptr()
        assert (v.data_ptr() - base) % 256 == 0 and v.is_contiguous()
        v.fill_(1.0)
    assert sum(float(b.sum()) for b in gb.flat) == n_params
    # backward-ready order: the vocabulary projection (half of all gradient bytes) is the first bucket
    assert gb.flat[0].numel() >= 0.49 * n_params and par.BUCKETS[-1] == ("embed",)
    assert [t.shape for t in gb.ordered()] == [torch.Size(_shapes(CFG_A)[f]) for f in WEIGHT_FIELDS]


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        dims = Dims(H=16, E=8, Vc=40, k=49)
        gb = par.GradBuckets(_shapes(dims), "cpu")
        red = par.BucketReducer(gb, average=False)
        g = torch.Generator().manual_seed(100 + rank)
        local = {f: torch.randn(v.shape, generator=g) for f, v in gb.views.items()}
        red.start()
        gb.loss.fill_(rank + 1.0)                            # the loss scalar rides in the last collective
        for b, fields in enumerate(par.BUCKETS):            # buckets become ready one by one, as in the backward
            for f in fields:
                gb.views[f].copy_(local[f])
            red.on_ready(b)
        n_coll = len(red.works)                              # LSTM + embedding + loss go through ONE all-reduce
        red.finish()
        ok_loss = abs(float(gb.loss) - sum(r + 1.0 for r in range(world))) < 1e-6 and n_coll == len(par.BUCKETS) - 1
        # expected: sum over ranks of each rank's seeded gradients
        exp = {}
        for r in range(world):
            gg = torch.Generator().manual_seed(100 + r)
            for f, v in gb.views.items():
                exp[f] = exp.get(f, 0) + torch.randn(v.shape, generator=gg)
        ok_sum = all(torch.allclose(gb.views[f], exp[f], atol=1e-6) for f in exp)
        # loss normalisation: global packed-token count
        lengths = [5, 3, 2] if rank == 0 else [4, 4, 4, 1]
        n_glob = par.global_token_count(lengths)
        # second step: rank 0 sees the SAME lengths again while rank 1's batch changed (last batch of an epoch, a length-bucketed
        # sampler): every rank must enter the collective again and get the new global count -- nothing may be cached per rank
        lengths2 = lengths if rank == 0 else [2, 1]
        n_glob2 = par.global_token_count(lengths2)
        # decode: every rank "decodes" its contiguous range; gathered ids are in image order
        n_img = 7
        lo, hi = par.shard_range(n_img, rank, world)
        ids_local = (torch.arange(lo, hi).view(-1, 1) * 10 + torch.arange(3).view(1, -1)).long()
        ids_all = par.gather_rows(ids_local, n_img)
        exp_ids = (torch.arange(n_img).view(-1, 1) * 10 + torch.arange(3).view(1, -1)).long()
        out[rank] = (ok_sum and ok_loss, red.order == [0, 1, 2, 3], n_glob, n_glob2, torch.equal(ids_all, exp_ids))
    finally:
        dist.destroy_process_group()


def test_reducer_count_and_gather_over_gloo_world2():
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    assert len(out) == world
    for r in range(world):
        ok_sum, order_ok, n_glob, n_glob2, ids_ok = out[r]
        assert ok_sum and order_ok and ids_ok
        assert n_glob == 10 + 13 and n_glob2 == 10 + 3


def test_dp_trainer_rejects_the_baseline_model_clearly():
    """The gradient buckets are laid out for the adaptive decoder's 13 tensors; the baseline model says so instead of failing obscurely."""
    from adaptive_b200 import baseline
    from adaptive_b200.parallel import DataParallelTrainer

    with pytest.raises(ValueError, match="baseline"):
        DataParallelTrainer(baseline.Encoder2Decoder())
