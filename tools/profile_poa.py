"""Small driver for ncu captures / ad-hoc timing: one warm-up launch + `--launches` timed launches of
rp_poa_kernel.  --shape ont (default): SURVEY.md §8d synthetic windows; --shape ngs: BASELINE config 4 shape
(w=200, 60 x 150-bp reads cut at window edges, qualities, kNGS); --shape real: 30 % partial-span layers + qualities."""
import argparse
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from racon_b200 import api, windows  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--windows", type=int, default=4736)
ap.add_argument("--launches", type=int, default=1)
ap.add_argument("--err", type=float, default=0.12)
ap.add_argument("--shape", default="ont", choices=["ont", "ngs", "real"])
ap.add_argument("--banded", type=int, default=0)
args = ap.parse_args()
wl = 500
if args.shape == "ont":
    ws, _ = windows.synth_windows(args.windows, err=args.err)
else:
    from tests import util
    base = 256
    if args.shape == "ngs":
        wl = 200
        small = util.make_set(7, base, wlen=200, depth=60, err=0.01, partial_frac=0.9, with_qual=True, min_piece=60)
        small.win_type[:] = 0
    else:
        small = util.make_set(7, base, wlen=500, depth=32, err=args.err, partial_frac=0.3, with_qual=True,
                              backbone_qual=True)
    idx = np.arange(args.windows) % base
    ws = small.subset(idx)
b = api.PoaBatch(window_length=wl, banded=bool(args.banded), mem_bytes=int(60e9))
assert b.add_window_set(ws) == args.windows
b.upload()
b.launch()
b.sync()
t0 = time.time()
for _ in range(args.launches):
    b.launch()
b.sync()
dt = time.time() - t0
b.download()
b.sync()
out, lens, pol, st = b.fetch_all(2 * 1100)
print("shape %s: %.0f windows/s (%d windows, %d launches, %.1f ms/launch); statuses %s; mean consensus %.1f"
      % (args.shape, args.windows * args.launches / dt, args.windows, args.launches, 1e3 * dt / args.launches,
         dict(zip(*np.unique(st, return_counts=True))), lens.mean()))
b.close()
