"""CPU, world_size 2 over gloo: the host-side logic of the batch-sharded (N > 1) path."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from neighbour_feature_pooling_b200 import sharding
from oracle import nfp_oracle as O


def test_shard_range_is_a_partition():
    for n in (0, 1, 7, 256, 257):
        for world in (1, 2, 3, 8):
            spans = [sharding.shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        sharding.shard_range(4, 2, 2)
    with pytest.raises(RuntimeError, match="averages over the batch"):
        sharding.check_shardable("scs")
    sharding.check_shardable("cosine")


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    os.environ.update(RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(world),
                      MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.set_num_threads(1)
    r, _, w = sharding.init_from_env("gloo")
    assert (r, w) == (rank, world)
    gen = torch.Generator().manual_seed(0)          # every rank builds the same global batch
    x = torch.randn(5, 6, 7, 7, generator=gen, dtype=torch.float64)
    xs = sharding.shard_batch(x, rank, world)
    # the data path has no collective: each rank computes its slice alone (oracle stands in for the kernel here)
    ys = O.nfp_forward(xs, R=1, measure="cosine", padding=1).mean(dim=(2, 3))
    full = sharding.gather_rows(ys)                 # ragged: 3 + 2 rows
    t = sharding.max_over_ranks(1.0 + rank)
    n = sharding.sum_over_ranks(float(xs.shape[0]))
    if rank == 0:
        ref = O.nfp_forward(x, R=1, measure="cosine", padding=1).mean(dim=(2, 3))
        out.put((torch.equal(full, ref), t, n))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharded_pass_equals_single_rank():
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    same, t, n = out.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert same, "sharded result differs from the single-rank result"
    assert t == 2.0 and n == 5.0


def test_train_harness_cpu_reference_arm():
    """bench_train's CPU baseline model: reference composition (backbone -> GAP(x) * proj(GAP(NFP(x))) -> fc)."""
    import bench_train
    cfg = bench_train.CONFIGS["eurosat"]
    model = bench_train.make_model(cfg, torch.device("cpu"), impl="reference")
    out = model(torch.randn(2, cfg["in_chans"], cfg["size"], cfg["size"]))
    assert out.shape == (2, cfg["classes"])
    r = bench_train.run_cpu_baseline("eurosat", batch=2, steps=1)
    assert r["images_per_s"] > 0 and r["kind"] in ("reference", "port")
