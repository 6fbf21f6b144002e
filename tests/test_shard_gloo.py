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


class _FakeCudaTorch:
    """torch with a pretend CUDA device: device 'cuda' means the CPU, pinned memory is ordinary memory, events complete one
    query late (so that the not-yet-complete paths of RowGather run too)."""

    def __init__(self):
        import torch
        self._t = torch
        self.events_recorded = 0
        self.pinned_allocations = 0
        outer = self

        class Event:
            def __init__(self):
                self.polls = 0

            def record(self):
                outer.events_recorded += 1

            def query(self):
                self.polls += 1
                return self.polls > 1

            def synchronize(self):
                self.polls = 2

        class Cuda:
            pass
        self.cuda = Cuda()
        self.cuda.Event = Event

    def __getattr__(self, k):
        return getattr(self._t, k)

    @staticmethod
    def _dev(device):
        return None if device is not None and str(device).startswith("cuda") else device

    def empty(self, *a, pin_memory=False, device=None, **kw):
        if pin_memory:
            self.pinned_allocations += 1
        return self._t.empty(*a, device=self._dev(device), **kw)

    def tensor(self, data, device=None, **kw):
        return self._t.tensor(data, device=self._dev(device), **kw)


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
    # the array form the bench uses: one data collective, result on rank 0 only
    stride = 200 + 8 * rank   # the ranks' row strides differ, as they do when the ranks' windows differ
    out = np.zeros((len(cons), stride), np.uint8)
    lens = np.zeros(len(cons), np.uint32)
    for k, c in enumerate(cons):
        out[k, :len(c)] = np.frombuffer(c, np.uint8)
        lens[k] = len(c)
    rows = shard.gather_rows(out, lens, dst=0)
    if rank == 0:
        flat = [rr[k, :ll[k]].tobytes() for rr, ll in rows for k in range(len(ll))]
        assert flat == full
    else:
        assert rows is None
    # the asynchronous form: two steps in flight one after the other, fixed block shape
    rg = shard.RowGather(len(cons), stride, dst=0)
    for step in range(2):
        rg.start(out, lens)
        got = rg.finish()
        if rank == 0:
            assert [rr[k, :ll[k]].tobytes() for rr, ll in got for k in range(len(ll))] == full
        else:
            assert got is None
    # the branch a GPU rank takes (pinned staging in rotation, non-blocking upload, read-back scheduled with an event, no
    # host wait in finish(block=False)), driven here with a stand-in for torch that plays a CUDA device on the CPU:
    # several steps with changing content, results must arrive in order and complete
    fake = _FakeCudaTorch()
    rg = shard.RowGather(len(cons), stride, device="cuda", dst=0, _torch=fake)
    seen = []   # (first byte, rows without their first byte) of every result, copied on receipt: the views are only valid
                # until the next finish()

    def note(got):
        if got is not None:
            seen.append((int(got[0][0][0, 0]), [rr[k, 1:ll[k]].tobytes() for rr, ll in got for k in range(len(ll))]))

    for step in range(6):
        o2 = out.copy()
        o2[:, 0] = step          # every step's rows differ in their first byte
        note(rg.finish(block=False))
        rg.start(o2, lens)
    note(rg.finish())
    if rank == 0:
        assert seen, "no gather result reached rank 0"
        first_bytes = [fb for fb, _ in seen]
        assert first_bytes == sorted(set(first_bytes)) and first_bytes[-1] == 5, first_bytes
        assert all(rows == [c[1:] for c in full] for _, rows in seen)
        assert fake.events_recorded >= 12 and fake.pinned_allocations <= 6
    else:
        assert not seen
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


def test_cost_balanced_ranges():
    """contiguous ranges balanced by the estimated cost sum_s L_s * (len_b + 0.1 * sum_{k<s} L_k)."""
    from racon_b200 import shard, windows
    from tests import util
    ws = util.make_set(5, 40, wlen=120, depth=8, err=0.1, partial_frac=0.4)
    cost = windows.window_costs(ws)
    # the vectorised cost equals the formula evaluated window by window
    for w in (0, 7, 39):
        win = ws.window(w)
        lb = len(win[0][0])
        acc, want = 0.0, 0.0
        for (b, _, _, _) in win[1:]:
            want += len(b) * (lb + 0.1 * acc)
            acc += len(b)
        assert abs(cost[w] - want) < 1e-6 * max(1.0, want)
    heavy = np.concatenate([cost[:20] * 10, cost[20:]])     # the first half ten times more expensive
    for world in (1, 2, 3, 8):
        b = [shard.shard_bounds_by_cost(heavy, r, world) for r in range(world)]
        assert b[0][0] == 0 and b[-1][1] == len(heavy) and all(b[i][1] == b[i + 1][0] for i in range(world - 1))
        loads = [heavy[lo:hi].sum() for lo, hi in b]
        assert max(loads) <= heavy.sum() / world + heavy.max() + 1e-9
    sl = windows.slice_windows(ws, 10, 25)
    assert sl.n_windows == 15 and sl.window(0) == ws.window(10) and sl.window(14) == ws.window(24)


def test_ngs_generator_shape():
    """BASELINE config 4 windows: all layers partial-span with qualities, kNGS, begin < end < backbone length."""
    from racon_b200 import windows
    ws, _ = windows.synth_ngs_windows(50)
    assert (ws.win_type == 0).all() and ws.quals is not None
    for w in (0, 17, 49):
        win = ws.window(w)
        bl = len(win[0][0])
        assert 190 <= bl <= 210 and win[0][1] is None and len(win) > 40
        for (b, q, s, e) in win[1:]:
            assert q is not None and len(q) == len(b) and 4 <= len(b) <= 150 and s < e < bl
            assert min(q) >= 33 + 30 and max(q) <= 33 + 40
