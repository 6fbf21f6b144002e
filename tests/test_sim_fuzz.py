"""Random small windows through the warp simulation of the device code vs the oracle (which tests/test_oracle.py pins to the
unmodified reference): lengths chosen around the 32-position blocks in which the traceback stores aln[], the 512-column
chunk edge of a DP row and the one-/two-chunk switch of the row loop; depths, error rates, partial spans, qualities, window
types, score sets and trimming drawn at random; full matrix and banded."""
import numpy as np
import pytest

from oracle import bindings as ob
from tests import simlib, util

LENGTHS = [1, 2, 31, 32, 33, 63, 64, 65, 96, 127, 129, 255, 257, 511, 512, 513, 700]


@pytest.mark.parametrize("seed", [11, 12, 13])
def test_sim_random_windows_equal_oracle(seed):
    rng = np.random.default_rng(seed)
    for it in range(9):
        wlen = int(LENGTHS[(it * 3 + seed) % len(LENGTHS)])
        depth = int(rng.integers(2, 12))
        err = float(rng.choice([0.0, 0.02, 0.12, 0.3]))
        pf = float(rng.choice([0.0, 0.3, 0.9]))
        alpha = b"ACGT" if rng.random() < 0.8 else b"ACGTN"
        ws = util.make_set(int(rng.integers(1 << 30)), 2 if wlen > 300 else 4, wlen=wlen, depth=depth, err=err,
                           partial_frac=pf, with_qual=bool(rng.integers(2)), backbone_qual=bool(rng.integers(2)),
                           alphabet=alpha, min_piece=min(20, max(1, wlen // 2)))
        if rng.random() < 0.3:
            ws.win_type[:] = 0
        scores = [(3, -5, -4), (5, -4, -8), (1, -1, -1)][int(rng.integers(3))]
        trim = bool(rng.integers(2))
        ora, opol, _ = ob.oracle_consensus(ws, *scores, trim=trim, threads=2)
        for banded in (0, 1):
            cons, pol, st, _, _ = simlib.sim_consensus(ws, *scores, trim=trim, banded=banded)
            what = "seed %d it %d wlen %d depth %d err %.2f pf %.1f scores %s trim %s banded %d" % (
                seed, it, wlen, depth, err, pf, scores, trim, banded)
            assert (st == 0).all(), what
            assert cons == ora, what
            assert (pol == opol).all(), what
