"""Finds windows whose banded (-b) consensus differs from the full-matrix consensus on the lambda -f flow (all 3461
windows, unit scores) and saves them (gpurun_out/band_divergence.npz) for analysis in the CPU simulation."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from racon_b200 import api, windows  # noqa: E402
from tests.lambda_overlaps import LambdaOverlaps  # noqa: E402

lam = LambdaOverlaps("lambda_frag_overlaps.npz")
m, x, g = (int(v) for v in lam.z["scores"])
n_seq = len(lam.seq_off) - 1
pol = api.MirrorPolisher(lam.bases, lam.quals, lam.seq_off, lam.seq_has_qual, n_targets=n_seq, overlaps=lam.ov,
                         window_length=lam.window_length, quality_threshold=lam.quality_threshold, trim=True,
                         match=m, mismatch=x, gap=g, window_type_tgs=True, fragment_correction=True)
ex = pol.export()
pol.close()
ws = windows.WindowSet(bases=ex["bases"], quals=ex["quals"], seq_off=ex["seq_off"], seq_has_qual=ex["seq_has_qual"],
                       seq_begin=ex["seq_begin"], seq_end=ex["seq_end"], win_first=ex["win_first"], win_type=ex["win_type"])
full, _, st = api.consensus(ws, m, x, g)
out = {}
for k in (8, 4):
    os.environ["RP_POA_BAND_K"] = str(k)
    stats = {}
    band, _, st2 = api.consensus(ws, m, x, g, banded=True, band_stats=stats)
    diff = [w for w in range(ws.n_windows) if band[w] != full[w]]
    print("band K=%d: %d windows differ: %s; stats %s" % (k, len(diff), diff[:10], stats))
    out[k] = diff
pick = sorted(set(out[8]) | set(out[4]))[:6]
if pick:
    sub = ws.subset(pick)
    np.savez_compressed(os.path.join(ROOT, "gpurun_out", "band_divergence.npz"), pick=np.asarray(pick), bases=sub.bases,
                        quals=sub.quals if sub.quals is not None else np.zeros(0, np.uint8), seq_off=sub.seq_off,
                        seq_has_qual=sub.seq_has_qual if sub.seq_has_qual is not None else np.zeros(0, np.uint8),
                        seq_begin=sub.seq_begin, seq_end=sub.seq_end, win_first=sub.win_first, win_type=sub.win_type,
                        scores=np.asarray([m, x, g]), k8=np.asarray(out[8]), k4=np.asarray(out[4]))
    print("saved", pick)
