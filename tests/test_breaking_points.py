"""CIGAR -> breaking points: the oracle restatement against the unmodified reference's own output on
the lambda sample (CPU), and the product's device path against both (GPU)."""
import numpy as np
import pytest

from oracle import bindings as ob
from tests import util
from tests.lambda_overlaps import LambdaOverlaps


@pytest.fixture(scope="module")
def lam():
    return LambdaOverlaps()


def _cigar(q, t):
    return (ob.ref_edlib_cigar(q, t) if ob.have_ref() else ob.oracle_myers_cigar(q, t))[0]


def test_bp_oracle_pinned_on_lambda(lam):
    """edlib CIGAR (unmodified edlib, or its restatement) + oracle walk == the reference's breaking points."""
    idx = range(lam.n_overlaps()) if ob.have_ref() else range(0, lam.n_overlaps(), 12)   # restatement is O(nm)
    for k in idx:
        q, t, t_begin, t_end, q_start = lam.spans(k)
        got = ob.oracle_breaking_points(_cigar(q, t), t_begin, t_end, q_start, lam.window_length)
        assert np.array_equal(got, lam.expected_bp(k)), k


def test_bp_oracle_pinned_on_lambda_fragment_correction():
    """Same pin on the 7 780 all-vs-all overlaps (both strands) of the -f configuration (every 4th with the unmodified
    edlib, every 200th with its restatement)."""
    lam = LambdaOverlaps("lambda_frag_overlaps.npz")
    for k in range(0, lam.n_overlaps(), 4 if ob.have_ref() else 200):
        q, t, t_begin, t_end, q_start = lam.spans(k)
        got = ob.oracle_breaking_points(_cigar(q, t), t_begin, t_end, q_start, lam.window_length)
        assert np.array_equal(got, lam.expected_bp(k)), k


def test_bp_oracle_small_cases():
    # one window, all matches: first match (t_begin, q_start), one past last match
    assert ob.oracle_breaking_points("10M", 0, 10, 0, 500).tolist() == [[0, 0], [10, 10]]
    # window boundary inside a match run: 495..504 crosses i = 500
    assert ob.oracle_breaking_points("10M", 495, 505, 7, 500).tolist() == [[495, 7], [500, 12], [500, 12], [505, 17]]
    # deletion across the boundary, insertion before it; a window without any match emits nothing
    assert ob.oracle_breaking_points("3M2I4D3M", 496, 506, 0, 500).tolist() == [[496, 0], [499, 3], [503, 5], [506, 8]]
    assert ob.oracle_breaking_points("4D6M", 496, 506, 0, 500).tolist() == [[500, 0], [506, 6]]


def test_sim_bp_device_code_vs_oracle_and_reference(lam):
    """The device function (myers_core.cuh: breaking_points) through the warp simulation: random pairs against the
    oracle, and a few real lambda overlaps against the reference's own breaking points."""
    from tests import simlib
    rng = np.random.default_rng(5)
    for wl in (64, 500):
        cases = []
        for _ in range(12):
            n = int(rng.integers(20, 1500))
            t = bytes(util.BASES[i] for i in rng.integers(4, size=n))
            q = util.mutate(rng, t, float(rng.uniform(0.0, 0.3))) or b"A"
            cases.append((q, t, int(rng.integers(0, 3 * wl)), int(rng.integers(0, 500))))
        cases.append((b"ACGT" * 40, b"ACGT" * 40, wl - 3, 0))
        cases.append((b"ACGT" * 40, b"ACGT" * 40, wl, 9))
        got, st = simlib.sim_align_bp(cases, wl)
        assert (st == 0).all()
        for (q, t, tb, qs), (cig, _, bp) in zip(cases, got):
            assert np.array_equal(bp, ob.oracle_breaking_points(cig, tb, tb + len(t), qs, wl))
    short = sorted(range(lam.n_overlaps()), key=lambda k: int(lam.ov[k][7] - lam.ov[k][6]))[:3]
    cases = [(lambda s: (s[0], s[1], s[2], s[4]))(lam.spans(k)) for k in short]
    got, st = simlib.sim_align_bp(cases, lam.window_length)
    assert (st == 0).all()
    for k, (_, _, bp) in zip(short, got):
        assert np.array_equal(bp, lam.expected_bp(k)), k


def test_sim_bp_window_boundary_sweep():
    """Tiny windows (7 bases) and every alignment of overlap start / end against the window grid, with indel-rich
    pairs: device breaking points == oracle walk of the same CIGAR."""
    from tests import simlib
    rng = np.random.default_rng(8)
    wl = 7
    cases = []
    for t_begin in range(0, 15):
        n = int(rng.integers(1, 40))
        t = bytes(util.BASES[i] for i in rng.integers(4, size=n))
        q = util.mutate(rng, t, 0.35) or b"A"
        cases.append((q, t, t_begin, int(rng.integers(0, 50))))
    for n in (7, 14, 21):     # ends exactly on the grid
        t = bytes(util.BASES[i] for i in rng.integers(4, size=n))
        cases.append((util.mutate(rng, t, 0.2) or b"C", t, 7, 3))
        cases.append((util.mutate(rng, t, 0.2) or b"C", t, 0, 0))
    got, st = simlib.sim_align_bp(cases, wl)
    assert (st == 0).all()
    for (q, t, tb, qs), (cig, _, bp) in zip(cases, got):
        assert np.array_equal(bp, ob.oracle_breaking_points(cig, tb, tb + len(t), qs, wl)), (tb, len(t), cig)


@pytest.mark.gpu
def test_gpu_bp_lambda(lam):
    """Real overlaps end to end on the device: CIGAR identical to edlib's, breaking points identical to the
    reference's (tests/golden/lambda_overlaps.npz)."""
    from racon_b200 import api
    b = api.AlnBatch()
    b.set_window_length(lam.window_length)
    for k in range(lam.n_overlaps()):
        q, t, t_begin, t_end, q_start = lam.spans(k)
        assert b.add(q, t, t_begin=t_begin, q_start=q_start)
    b.run()
    b.sync()
    for k in range(lam.n_overlaps()):
        cig, dist, st = b.fetch(k)
        assert st == 0, (k, st)
        q, t, t_begin, t_end, q_start = lam.spans(k)
        if k % 6 == 0:
            assert cig.decode() == _cigar(q, t), k
        assert np.array_equal(b.fetch_breaking_points(k), lam.expected_bp(k)), k
    b.close()


@pytest.mark.gpu
def test_gpu_bp_lambda_fragment_correction():
    """The 7780 all-vs-all overlaps (both strands) of the -f configuration: device breaking points == the reference's."""
    from racon_b200 import api
    lam = LambdaOverlaps("lambda_frag_overlaps.npz")
    b = api.AlnBatch()
    b.set_window_length(lam.window_length)
    for k in range(lam.n_overlaps()):
        q, t, t_begin, t_end, q_start = lam.spans(k)
        assert b.add(q, t, t_begin=t_begin, q_start=q_start)
    b.run()
    b.sync()
    for k in range(lam.n_overlaps()):
        assert b.fetch(k)[2] == 0, k
        assert np.array_equal(b.fetch_breaking_points(k), lam.expected_bp(k)), k
    b.close()


@pytest.mark.gpu
def test_gpu_bp_random_pairs_vs_oracle():
    from racon_b200 import api
    rng = np.random.default_rng(21)
    b = api.AlnBatch()
    cases = []
    for wl in (100, 500, 1000):
        b.reset()
        b.set_window_length(wl)
        cases = []
        for _ in range(40):
            n = int(rng.integers(50, 4000))
            t = bytes(util.BASES[i] for i in rng.integers(4, size=n))
            q = util.mutate(rng, t, float(rng.uniform(0.0, 0.3))) or b"A"
            t_begin = int(rng.integers(0, 3 * wl))
            q_start = int(rng.integers(0, 1000))
            cases.append((q, t, t_begin, q_start))
            assert b.add(q, t, t_begin=t_begin, q_start=q_start)
        b.run()
        b.sync()
        for i, (q, t, t_begin, q_start) in enumerate(cases):
            cig, _, st = b.fetch(i)
            assert st == 0
            exp = ob.oracle_breaking_points(cig, t_begin, t_begin + len(t), q_start, wl)
            assert np.array_equal(b.fetch_breaking_points(i), exp), (wl, i)
    b.close()
