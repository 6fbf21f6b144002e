"""Not a pytest module (run by hand on the GPU box): the batched aligner against the UNMODIFIED edlib (oracle/_ref) on
the same synthetic overlaps — identical CIGARs, and pairs/s for both (edlib on one host thread).
    python tests/perf_aln_vs_edlib.py [--pairs 12000] [--len 8000] [--err 0.12] [--cpu-sample 40]"""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--pairs", type=int, default=12000)
    ap.add_argument("--len", type=int, default=8000)
    ap.add_argument("--err", type=float, default=0.12)
    ap.add_argument("--cpu-sample", type=int, default=40)
    a = ap.parse_args()
    from oracle import bindings as ob
    from racon_b200 import api
    from tests import util
    rng = np.random.default_rng(5)
    pairs = []
    for _ in range(a.pairs):
        n = int(a.len * rng.uniform(0.5, 1.5))
        t = bytes(util.BASES[i] for i in rng.integers(4, size=n))
        pairs.append((util.mutate(rng, t, a.err), t))
    b = api.AlnBatch()
    for q, t in pairs:
        assert b.add(q, t)
    b.upload()
    b.sync()
    best = 1e9
    for _ in range(3):
        t0 = time.perf_counter()
        b.launch()
        b.sync()
        best = min(best, time.perf_counter() - t0)
    b.download()
    b.sync()
    k = min(a.cpu_sample, len(pairs))
    t0 = time.perf_counter()
    ref = [ob.ref_edlib_cigar(*pairs[i]) for i in range(k)]
    dt = time.perf_counter() - t0
    same = sum(1 for i in range(k) if ref[i][0].encode() == b.fetch(i)[0])
    print(json.dumps({"pairs": a.pairs, "mean_len": a.len, "err": a.err, "gpu_pairs_per_s": a.pairs / best,
                      "edlib_1thread_pairs_per_s": k / dt, "edlib_sample": k, "sample_identical": same}))


if __name__ == "__main__":
    main()
