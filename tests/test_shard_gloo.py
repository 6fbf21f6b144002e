"""N > 1 host logic on CPU (world_size 2, gloo): contiguous sharding of the window stream + the final consensus
gather.  The per-shard compute stand-in is the test-only warp simulation of the device code (tiny windows)."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_windows, ret):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from racon_b200 import shard
    from tests import simlib, util
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    ws = util.make_set(123, n_windows, wlen=60, depth=5, err=0.12, partial_frac=0.3)
    lo, hi = shard.shard_bounds(n_windows, rank, world)
    cons, pol, st, _, _ = simlib.sim_consensus(ws.subset(range(lo, hi)))
    assert (st == 0).all()
    full = shard.gather_consensus(cons)
    ret[rank] = full
    dist.barrier()
    dist.destroy_process_group()


def test_shard_bounds_cover_the_stream():
    from racon_b200 import shard
    for n in (0, 1, 7, 10, 10000):
        for world in (1, 2, 3, 8):
            b = [shard.shard_bounds(n, r, world) for r in range(world)]
            assert b[0][0] == 0 and b[-1][1] == n
            assert all(b[i][1] == b[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in b]
            assert max(sizes) - min(sizes) <= 1


def test_two_rank_gloo_shard_and_gather():
    import torch.multiprocessing as mp
    from oracle import bindings as ob
    from tests import util
    n = 9
    port = _free_port()
    mgr = mp.Manager()
    ret = mgr.dict()
    ctx = mp.get_context("spawn")
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n, ret)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(300)
        assert p.exitcode == 0
    ws = util.make_set(123, n, wlen=60, depth=5, err=0.12, partial_frac=0.3)
    ora, _, _ = ob.oracle_consensus(ws, threads=2)
    assert list(ret[0]) == ora and list(ret[1]) == ora
