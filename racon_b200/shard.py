"""Multi-GPU plumbing: windows are independent (src/polisher.cpp:495-503), so each rank owns a contiguous shard
of the window stream and the only exchange is the final gather of the consensus bytes.
Works with any torch.distributed backend: "nccl" on the GPU box, "gloo" in the CPU tests."""
import numpy as np


def shard_bounds(n_windows, rank, world):
    """Contiguous, count-balanced window range [lo, hi) of `rank` (output order = window order)."""
    base, rem = divmod(n_windows, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_bounds_by_cost(costs, rank, world):
    """Contiguous, COST-balanced window range [lo, hi) of `rank`: boundaries at the equal quantiles
    of the cumulative per-window cost estimate (windows.window_costs); output order stays window order."""
    c = np.cumsum(np.asarray(costs, dtype=np.float64))
    n = len(c)
    if n == 0 or c[-1] <= 0:
        return shard_bounds(n, rank, world)
    cuts = [0] + [int(np.searchsorted(c, c[-1] * k / world, side="left")) + 1 for k in range(1, world)] + [n]
    cuts = np.minimum(np.maximum.accumulate(cuts), n)
    return int(cuts[rank]), int(cuts[rank + 1])


def pack_consensus(cons):
    """list of bytes -> (flat uint8 array, uint32 lengths)"""
    lens = np.asarray([len(c) for c in cons], dtype=np.uint32)
    flat = np.frombuffer(b"".join(cons), dtype=np.uint8).copy() if len(cons) else np.zeros(0, np.uint8)
    return flat, lens


def gather_packed(flat, lens, device=None, group=None):
    """Array form of the gather: every rank contributes (flat uint8 bytes, uint32 lengths) of its shard and gets
    back the per-rank lists [(flat_r, lens_r)] in rank (= window) order.  Three collectives: counts, lengths, bytes."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    dev = device if device is not None else "cpu"
    n = int(len(lens))
    counts = torch.tensor([n, int(flat.size)], dtype=torch.int64, device=dev)
    all_counts = [torch.zeros_like(counts) for _ in range(world)]
    dist.all_gather(all_counts, counts, group=group)
    max_n = int(max(int(c[0]) for c in all_counts))
    max_b = int(max(int(c[1]) for c in all_counts))
    lbuf = torch.zeros(max(max_n, 1), dtype=torch.int32, device=dev)
    lbuf[:n] = torch.from_numpy(np.asarray(lens).astype(np.int32)).to(dev)
    bbuf = torch.zeros(max(max_b, 1), dtype=torch.uint8, device=dev)
    bbuf[:flat.size] = torch.from_numpy(np.ascontiguousarray(flat)).to(dev)
    all_l = [torch.zeros_like(lbuf) for _ in range(world)]
    all_b = [torch.zeros_like(bbuf) for _ in range(world)]
    dist.all_gather(all_l, lbuf, group=group)
    dist.all_gather(all_b, bbuf, group=group)
    out = []
    for r in range(world):
        nr, br = int(all_counts[r][0]), int(all_counts[r][1])
        out.append((all_b[r][:br].cpu().numpy(), all_l[r][:nr].cpu().numpy().astype(np.uint32)))
    return out


def gather_consensus(cons, device=None, group=None):
    """All ranks call this with their shard's consensus list; every rank gets the full list in window order."""
    flat, lens = pack_consensus(cons)
    out = []
    for bb, ll in gather_packed(flat, lens, device, group):
        off = 0
        for k in range(len(ll)):
            out.append(bb[off:off + int(ll[k])].tobytes())
            off += int(ll[k])
    return out


def gather_rows(out, lens, device=None, group=None, dst=0):
    """The final consensus gather as ONE data collective: every rank contributes its fetch_all() result (`out`: [n, stride]
    uint8 rows, `lens`: uint32 [n]) as a fixed-size block [n_max, 4 + len_max] (length prefix + padded row), gathered with
    all_gather_into_tensor on `device` (NCCL over NVLink on the GPU box; a 2-int all_reduce fixes n_max / len_max first).
    Only rank `dst` copies the gathered block back to the host and returns [(rows_r, lens_r)] in rank (= window) order;
    the other ranks return None."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    dev = device if device is not None else "cpu"
    n = int(len(lens))
    lmax = int(lens.max()) if n else 0
    dims = torch.tensor([n, lmax], dtype=torch.int64, device=dev)
    dist.all_reduce(dims, op=dist.ReduceOp.MAX, group=group)
    n_max, l_max = int(dims[0]), max(4, (int(dims[1]) + 3) // 4 * 4)
    block = np.zeros((n_max + 1, 4 + l_max), dtype=np.uint8)      # row 0: this rank's window count
    block[0, :4] = np.asarray([n], dtype="<u4").view(np.uint8)
    block[1:n + 1, :4] = np.asarray(lens, dtype="<u4").view(np.uint8).reshape(n, 4)
    w = min(l_max, out.shape[1]) if n else 0
    block[1:n + 1, 4:4 + w] = out[:n, :w]
    mine = torch.from_numpy(block).to(dev)
    gathered = torch.empty((world,) + block.shape, dtype=torch.uint8, device=dev)
    try:
        dist.all_gather_into_tensor(gathered, mine, group=group)
    except Exception:  # a backend without the flat variant
        parts = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(parts, mine, group=group)
        gathered = torch.stack(parts)
    if rank != dst:
        return None
    g = gathered.cpu().numpy()
    res = []
    for r in range(world):
        nr = int(g[r, 0, :4].copy().view("<u4")[0])
        ll = g[r, 1:nr + 1, :4].copy().view("<u4").reshape(-1)
        res.append((g[r, 1:nr + 1, 4:], ll.astype(np.uint32)))
    return res


class RowGather:
    """The per-step consensus gather of a multi-rank run, asynchronous: start() queues ONE all_gather_into_tensor of this
    rank's fixed-size block [n_max + 1, 4 + l_max] (row 0: window count; then length prefix + padded consensus per window)
    and returns at once; finish() waits for it and, on rank `dst`, copies the gathered blocks to the host and returns
    [(rows_r, lens_r)] in rank (= window) order (views of a reused pinned buffer: valid until the next finish()).  The block shape is fixed when the object is made (n_max = the largest
    per-rank window count, l_max = the largest row stride of any rank: both agreed with one all-reduce here), so a step needs no shape
    exchange and no rank waits for another one inside its step: the next step's packing and kernels overlap the gather."""

    def __init__(self, n_local, l_max, device=None, group=None, dst=0, _torch=None):
        import torch.distributed as dist
        if _torch is None:   # tests hand in a stand-in that plays a CUDA device on the CPU (tests/test_shard_gloo.py)
            import torch
        else:
            torch = _torch
        self.torch, self.dist, self.group, self.dst = torch, dist, group, dst
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.dev = device if device is not None else "cpu"
        # both dimensions of the block are agreed here, once: every rank must hand the collective the same shape, and the
        # ranks' row strides differ as soon as their windows do (found by tools/mock_bench.py --mock-world 2)
        t = torch.tensor([int(n_local), int(l_max)], dtype=torch.int64, device=self.dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
        self.n_max = int(t[0])
        self.l_max = max(4, (int(t[1]) + 3) // 4 * 4)
        self.pending = None
        self.inflight = []   # read-backs queued on the device: (event, pinned view, device tensor kept alive, pinned buffer)
        self.bufs = []       # up to three pinned buffers, reused
        self.done = None     # newest completed gather not handed out yet
        self.done_buf = None # its pinned buffer: not reused while the caller may still read the returned views
        self.up = []         # pinned staging blocks of start(): [tensor, event of its last upload]
        self.n_started = 0

    def start(self, out, lens):
        torch, dist = self.torch, self.dist
        if self.pending is not None:
            raise RuntimeError("RowGather.start: the previous gather was not finished")
        n = int(len(lens))
        if n > self.n_max:
            raise ValueError("RowGather: %d windows exceed the agreed maximum %d" % (n, self.n_max))
        shape = (self.n_max + 1, 4 + self.l_max)
        on_gpu = str(self.dev).startswith("cuda")
        if on_gpu:
            # pinned staging, three blocks in rotation (a block is rewritten two steps after its upload was queued; the
            # event makes that safe whatever the timing), uploaded without blocking the host
            if len(self.up) < 3:
                self.up.append([torch.empty(shape, dtype=torch.uint8, pin_memory=True), None])
            slot = self.up[self.n_started % len(self.up)] if len(self.up) == 3 else self.up[-1]
            if slot[1] is not None:
                slot[1].synchronize()
            block = slot[0].numpy()
            block[:] = 0
        else:
            block = np.zeros(shape, dtype=np.uint8)
        self.n_started += 1
        block[0, :4] = np.asarray([n], dtype="<u4").view(np.uint8)
        block[1:n + 1, :4] = np.asarray(lens, dtype="<u4").view(np.uint8).reshape(n, 4)
        w = min(self.l_max, out.shape[1]) if n else 0
        block[1:n + 1, 4:4 + w] = out[:n, :w]
        if on_gpu:
            mine = torch.empty(shape, dtype=torch.uint8, device=self.dev)
            mine.copy_(slot[0], non_blocking=True)
            slot[1] = torch.cuda.Event()
            slot[1].record()
        else:
            mine = torch.from_numpy(block).to(self.dev)
        gathered = torch.empty((self.world,) + shape, dtype=torch.uint8, device=self.dev)
        try:
            work = dist.all_gather_into_tensor(gathered, mine, group=self.group, async_op=True)
            parts = None
        except Exception:  # a backend without the flat variant
            parts = [torch.empty_like(mine) for _ in range(self.world)]
            work = dist.all_gather(parts, mine, group=self.group, async_op=True)
        self.pending = (work, gathered, parts, mine)

    def _parse(self, g):
        res = []
        for r in range(self.world):
            nr = int(g[r, 0, :4].copy().view("<u4")[0])
            ll = g[r, 1:nr + 1, :4].copy().view("<u4").reshape(-1)
            res.append((g[r, 1:nr + 1, 4:], ll.astype(np.uint32)))
        return res

    def finish(self, block=True):
        """Completes the gather queued by start().  block=True (default): waits for it and returns, on rank `dst`,
        [(rows_r, lens_r)] in rank order (None on the other ranks).  block=False: never waits on the host — the read-back of
        the gather just queued is only scheduled, and what is returned is the newest EARLIER gather that has completed
        meanwhile (or None); a final finish(block=True) collects the rest.

        On a GPU nothing here launches a kernel or synchronises the host in the non-blocking form: the read-back is one
        device-to-host copy of the (contiguous) gathered block on the copy engine, ordered after the collective by the
        stream.  That matters next to a persistent kernel that fills every SM: any kernel queued here — even a tiny slicing
        or reduction kernel — would wait for that kernel to drain, and a host thread waiting for it cannot launch the next
        batch."""
        torch = self.torch
        if self.pending is not None:
            work, gathered, parts, _mine = self.pending
            self.pending = None
            work.wait()          # NCCL: orders the current stream after the collective, does not block the host
            if self.rank == self.dst:
                g_dev = torch.stack(parts) if parts is not None else gathered
                if str(self.dev).startswith("cuda"):
                    while len(self.inflight) >= 2:   # keep one of the three buffers for the result not handed out yet
                        ev, h, _keep, buf_done = self.inflight.pop(0)
                        ev.synchronize()
                        self.done, self.done_buf = h, buf_done
                    n = g_dev.numel()
                    if len(self.bufs) < 3:
                        self.bufs.append(torch.empty(n, dtype=torch.uint8, pin_memory=True))
                    busy = {id(x[3]) for x in self.inflight}
                    busy.add(id(self.done_buf))
                    buf = next(b for b in self.bufs if id(b) not in busy and b.numel() >= n)
                    h = buf[:n].view(g_dev.shape)
                    h.copy_(g_dev, non_blocking=True)      # plain D2H memcpy (both sides contiguous)
                    ev = torch.cuda.Event()
                    ev.record()
                    self.inflight.append((ev, h, g_dev, buf))
                else:
                    self.done = g_dev
        if self.rank != self.dst:
            return None
        while self.inflight and (block or self.inflight[0][0].query()):
            ev, h, _keep, buf_done = self.inflight.pop(0)
            ev.synchronize()
            self.done, self.done_buf = h, buf_done
        if self.done is None:
            return None
        g, self.done = self.done.numpy(), None   # done_buf stays reserved until the next result replaces it
        return self._parse(g)
