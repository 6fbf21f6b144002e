"""Generates tests/golden/poa_golden.json from oracle/_ref (the UNMODIFIED reference compiled from
/root/reference by oracle/Makefile).  Run in the CPU container:  python tests/golden/make_golden.py
The fixture is small (a few windows per case) so it can be committed; it travels to the GPU box where
/root/reference does not exist."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import bindings as ob  # noqa: E402
from tests import util  # noqa: E402

SPEC = [
    ("fullspan_default", dict(n=3, wlen=500, depth=32, err=0.12), (3, -5, -4), True, 1),
    ("partial_qual_5_4_8", dict(n=3, wlen=500, depth=24, err=0.12, partial_frac=0.4, with_qual=True,
                                backbone_qual=True), (5, -4, -8), True, 1),
    ("ngs_w200", dict(n=4, wlen=200, depth=40, err=0.01, partial_frac=0.9, with_qual=True), (3, -5, -4), True, 0),
    ("acgtn_notrim", dict(n=3, wlen=300, depth=16, err=0.15, partial_frac=0.3, alphabet=b"ACGTN"), (1, -1, -1), False, 1),
    ("shallow", dict(n=6, wlen=100, depth=3, err=0.2, partial_frac=0.3), (3, -5, -4), True, 1),
]


def main():
    cases = []
    for i, (name, kw, scores, trim, wtype) in enumerate(SPEC):
        kw = dict(kw)
        n = kw.pop("n")
        ws = util.make_set(1000 + i, n, **kw)
        ws.win_type[:] = wtype
        cons, pol, _ = ob.ref_consensus(ws, *scores, trim=trim, threads=4)
        wins = []
        for w in range(ws.n_windows):
            wins.append([[b.decode(), q.decode() if q else None, s, e] for (b, q, s, e) in ws.window(w)])
        cases.append(dict(name=name, scores=list(scores), trim=trim, types=[wtype] * n, windows=wins,
                          consensus=[c.decode() for c in cons], polished=[bool(p) for p in pol]))
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "poa_golden.json")
    with open(out, "w") as f:
        json.dump(dict(source="oracle/_ref (unmodified /root/reference, commit a2cfcac)", cases=cases), f)
    print("wrote", out, os.path.getsize(out), "bytes")


if __name__ == "__main__":
    main()
