"""Small driver for ncu captures: one warm-up launch + `--launches` timed launches of rp_poa_kernel over
`--windows` synthetic windows (SURVEY.md §8d generator).  Run under ncu with -k regex:rp_poa_kernel."""
import argparse
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from racon_b200 import api, windows  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--windows", type=int, default=4736)
ap.add_argument("--launches", type=int, default=1)
ap.add_argument("--err", type=float, default=0.12)
args = ap.parse_args()
ws, _ = windows.synth_windows(args.windows, err=args.err)
b = api.PoaBatch()
assert b.add_window_set(ws) == args.windows
b.upload()
b.launch()
b.sync()
t0 = time.time()
for _ in range(args.launches):
    b.launch()
b.sync()
print("windows/s %.0f" % (args.windows * args.launches / (time.time() - t0)))
b.close()
