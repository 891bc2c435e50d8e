"""N>1 path on CPU: world_size-2 (and 3) gloo process groups exercise the site sharding and the one
collective of the path (sum of per-rank scaler increments).  The per-rank compute stand-in is
the oracle (test infrastructure); on the GPU box the same code runs with the CUDA kernel and
NCCL (bench.py)."""
from __future__ import annotations

import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT, load_pkg


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import oracle
        pkg = load_pkg()
        from plf_b200 import sharding
        lo, cnt = sharding.shard_for_rank(n, rank, world)
        ev, left, right, x1, x2, wgt = oracle.host_mem_inputs(n, seed=42)   # same on every rank
        x3, sc, inc = oracle.COracle().newview(x1[lo:lo + cnt], x2[lo:lo + cnt], ev, left, right,
                                               wgt[lo:lo + cnt])
        total = sharding.reduce_scaler_increment(inc)
        slowest = sharding.max_over_ranks(float(rank + 1))
        np.savez(os.path.join(out_dir, f"r{rank}.npz"), lo=lo, cnt=cnt, x3=x3, sc=sc, inc=inc,
                 total=total, slowest=slowest)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,n", [(2, 1001), (3, 100), (2, 1)])
def test_sharded_newview_matches_single_run(tmp_path, coracle, world, n):
    import oracle
    port = _free_port()
    mp.spawn(_worker, args=(world, port, n, str(tmp_path)), nprocs=world, join=True)
    ev, left, right, x1, x2, wgt = oracle.host_mem_inputs(n, seed=42)
    full3, fullsc, fullinc = coracle.newview(x1, x2, ev, left, right, wgt)
    covered = 0
    for r in range(world):
        z = np.load(tmp_path / f"r{r}.npz")
        lo, cnt = int(z["lo"]), int(z["cnt"])
        assert lo == covered            # contiguous, ordered, no overlap
        covered += cnt
        assert np.array_equal(z["x3"].view(np.uint32), full3[lo:lo + cnt].view(np.uint32))
        assert np.array_equal(z["sc"], fullsc[lo:lo + cnt])
        assert int(z["total"]) == fullinc           # all-reduced on every rank
        assert float(z["slowest"]) == float(world)
    assert covered == n


def test_shard_rule_is_the_reference_instance_split(pkg):
    from plf_b200 import sharding
    for n, w in ((64 << 20, 8), (1000, 3), (7, 8), (0, 4)):
        shards = sharding.all_shards(n, w)
        assert sum(c for _, c in shards) == n
        assert shards == pkg.partition_sites(n, w)
    assert sharding.all_shards(64 << 20, 8)[3] == (3 * (8 << 20), 8 << 20)
    with pytest.raises(ValueError):
        sharding.shard_for_rank(10, 2, 2)
    assert sharding.reduce_scaler_increment(5) == 5       # no process group: identity


def test_bind_host_to_device_never_narrows_to_nothing(pkg):
    """The NUMA helper is best effort: without NVML it reports why and leaves the affinity mask alone."""
    import importlib
    sh = importlib.import_module(pkg.__name__ + ".sharding")
    before = os.sched_getaffinity(0)
    info = sh.bind_host_to_device(0)
    after = os.sched_getaffinity(0)
    try:
        assert isinstance(info, dict) and "bound" in info
        assert after and after <= before
        if not info["bound"]:
            assert after == before
    finally:
        os.sched_setaffinity(0, before)
