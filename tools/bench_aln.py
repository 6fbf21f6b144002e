"""Timing of the batched pre-alignment row (rp_aln_*): pairs/s and cell updates/s through the C ABI.
    python tools/bench_aln.py [--pairs 2000] [--len 8000] [--err 0.12]
(tests/perf_aln_vs_edlib.py times the unmodified edlib on the same pairs and checks the CIGARs.)"""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--pairs", type=int, default=2000)
    ap.add_argument("--len", type=int, default=8000)
    ap.add_argument("--err", type=float, default=0.12)
    ap.add_argument("--cpu-sample", type=int, default=40)
    ap.add_argument("--reps", type=int, default=3)
    a = ap.parse_args()
    from racon_b200 import api
    from tests import util
    rng = np.random.default_rng(5)
    pairs = []
    for _ in range(a.pairs):
        n = int(a.len * rng.uniform(0.5, 1.5))
        t = bytes(util.BASES[i] for i in rng.integers(4, size=n))
        pairs.append((util.mutate(rng, t, a.err), t))
    cells = sum(len(q) * len(t) for q, t in pairs)
    b = api.AlnBatch()
    for q, t in pairs:
        assert b.add(q, t)
    b.upload()
    b.sync()
    best = None
    for _ in range(a.reps):
        t0 = time.perf_counter()
        b.launch()
        b.sync()
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    t0 = time.perf_counter()
    b.reset()
    for q, t in pairs:
        assert b.add(q, t)
    b.run()
    b.sync()
    cig = [b.fetch(i) for i in range(len(pairs))]
    e2e = time.perf_counter() - t0
    fails = sum(1 for c in cig if c[2] != 0)
    out = {"pairs": a.pairs, "mean_len": a.len, "err": a.err, "kernel_s": best, "pairs_per_s": a.pairs / best,
           "gcups_full_matrix": cells / best / 1e9, "e2e_s": e2e, "e2e_pairs_per_s": a.pairs / e2e,
           "soft_failures": fails, "info": b.info()}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
