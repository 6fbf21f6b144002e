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
    per-rank window count, agreed with one all-reduce here; l_max = the caller's row stride), so a step needs no shape
    exchange and no rank waits for another one inside its step: the next step's packing and kernels overlap the gather."""

    def __init__(self, n_local, l_max, device=None, group=None, dst=0):
        import torch
        import torch.distributed as dist
        self.torch, self.dist, self.group, self.dst = torch, dist, group, dst
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.dev = device if device is not None else "cpu"
        t = torch.tensor([int(n_local)], dtype=torch.int64, device=self.dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
        self.n_max = int(t[0])
        self.l_max = max(4, (int(l_max) + 3) // 4 * 4)
        self.pending = None
        self.host = None

    def start(self, out, lens):
        torch, dist = self.torch, self.dist
        if self.pending is not None:
            raise RuntimeError("RowGather.start: the previous gather was not finished")
        n = int(len(lens))
        if n > self.n_max:
            raise ValueError("RowGather: %d windows exceed the agreed maximum %d" % (n, self.n_max))
        block = np.zeros((self.n_max + 1, 4 + self.l_max), dtype=np.uint8)
        block[0, :4] = np.asarray([n], dtype="<u4").view(np.uint8)
        block[1:n + 1, :4] = np.asarray(lens, dtype="<u4").view(np.uint8).reshape(n, 4)
        w = min(self.l_max, out.shape[1]) if n else 0
        block[1:n + 1, 4:4 + w] = out[:n, :w]
        mine = torch.from_numpy(block).to(self.dev, non_blocking=False)
        gathered = torch.empty((self.world,) + block.shape, dtype=torch.uint8, device=self.dev)
        try:
            work = dist.all_gather_into_tensor(gathered, mine, group=self.group, async_op=True)
            parts = None
        except Exception:  # a backend without the flat variant
            parts = [torch.empty_like(mine) for _ in range(self.world)]
            work = dist.all_gather(parts, mine, group=self.group, async_op=True)
        self.pending = (work, gathered, parts, mine)

    def finish(self):
        if self.pending is None:
            return None
        work, gathered, parts, _mine = self.pending
        self.pending = None
        work.wait()
        if self.rank != self.dst:
            return None
        torch = self.torch
        g_dev = torch.stack(parts) if parts is not None else gathered
        # only the columns some row uses travel to the host (the block is as wide as the caller's row stride)
        lens_dev = g_dev[:, 1:, :4].contiguous().view(torch.int32)
        w = min(self.l_max, (int(lens_dev.max()) + 3) // 4 * 4) if lens_dev.numel() else 0
        cut = g_dev[:, :, :4 + w].contiguous()
        if cut.is_cuda:
            if self.host is None:
                self.host = torch.empty(self.world * (self.n_max + 1) * (4 + self.l_max), dtype=torch.uint8, pin_memory=True)
            h = self.host[:cut.numel()].view(cut.shape)
            h.copy_(cut, non_blocking=True)
            torch.cuda.current_stream().synchronize()
            g = h.numpy()
        else:
            g = cut.numpy()
        res = []
        for r in range(self.world):
            nr = int(g[r, 0, :4].copy().view("<u4")[0])
            ll = g[r, 1:nr + 1, :4].copy().view("<u4").reshape(-1)
            res.append((g[r, 1:nr + 1, 4:], ll.astype(np.uint32)))
        return res
