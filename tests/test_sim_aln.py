"""Device code of the batched aligner (myers_core.cuh) through the test-only warp simulation, against the oracle
(and, where built, the unmodified edlib): identical CIGAR strings and distances in both of edlib's regimes."""
import numpy as np
import pytest

from oracle import bindings as ob
from tests import simlib, util


def _pair(rng, n, err, skew=0.0):
    t = bytes(util.BASES[i] for i in rng.integers(4, size=n))
    q = util.mutate(rng, t, err)
    if skew:
        k = int(len(q) * skew)
        q = q[k // 2:len(q) - k // 2]
    return q or b"A", t


def _check(pairs, **kw):
    got, st = simlib.sim_align(pairs, **kw)
    assert (st == 0).all(), st
    for (q, t), g in zip(pairs, got):
        exp = ob.oracle_myers_cigar(q, t)
        assert g == exp, (len(q), len(t), g[1], exp[1])
        if ob.have_ref():
            assert g == ob.ref_edlib_cigar(q, t)


def test_sim_aln_small_and_degenerate():
    pairs = [(b"A", b"A"), (b"A", b"C"), (b"ACGT", b"A"), (b"A", b"ACGT"), (b"AAAA", b"TTTT"), (b"ACGTACGT", b"ACGT"),
             (b"ACGTTGCA" * 9, b"ACGTTGCA" * 9), (b"ACGT" * 40, b"ACGT" * 5)]
    _check(pairs)


@pytest.mark.parametrize("n,err,count", [(30, 0.2, 12), (64, 0.1, 8), (65, 0.3, 8), (200, 0.15, 8), (700, 0.12, 4),
                                         (1500, 0.2, 2)])
def test_sim_aln_traceback_regime(n, err, count):
    rng = np.random.default_rng(n)
    _check([_pair(rng, n, err, skew=rng.choice([0.0, 0.2])) for _ in range(count)])


@pytest.mark.parametrize("n,err", [(2500, 0.12), (3500, 0.05), (3000, 0.3), (5600, 0.1)])
def test_sim_aln_hirschberg_regime(n, err):
    rng = np.random.default_rng(n + 7)
    lib = simlib.lib()
    lib.rp_sim_leaf_pairs.restype = __import__("ctypes").c_ulong
    before = lib.rp_sim_leaf_pairs()
    _check([_pair(rng, n, err, skew=rng.choice([0.0, 0.15]))])
    assert lib.rp_sim_leaf_pairs() > before   # sibling leaves were filled side by side (leaf_pair)


def test_sim_aln_band_growth_multi_round():
    """Unrelated sequences: distance above the first threshold, so k doubles and the band spans several rounds of
    32 words per column (carry hand-over between rounds)."""
    rng = np.random.default_rng(99)
    a = bytes(util.BASES[i] for i in rng.integers(4, size=3900))
    b = bytes(util.BASES[i] for i in rng.integers(4, size=3600))
    _check([(a, b)])


def test_sim_aln_wide_alphabet():
    """Up to 16 distinct characters per pair (IUPAC codes, mixed case); beyond that a soft status."""
    q = b"ACGTNRYKMSWB" * 12 + b"acgt" * 9
    t = b"ACGTNRYKMSW" * 13 + b"acga" * 8
    _check([(q, t)])
    got, st = simlib.sim_align([(bytes(range(65, 85)) * 4, bytes(range(65, 85)) * 4)])
    assert st[0] == 2 and got[0][0] == ""


def test_sim_aln_short_query_long_target_boundary_splits():
    """A 150-base query against an 8 kb target: Hirschberg regime (the 1 MiB rule depends on the target length) with
    split rows at the matrix boundary (r = -1 / r = qlen-1, edlib.cpp:1304-1320) and empty children."""
    rng = np.random.default_rng(123)
    t = bytes(util.BASES[i] for i in rng.integers(4, size=8000))
    for start in (0, 3900, 7850):
        q = util.mutate(rng, t[start:start + 150], 0.1)
        _check([(q, t)])
    _check([(b"ACGT" * 30, b"T" * 9000), (b"A", t)])


def test_sim_aln_stored_distance_pass_with_k_doubling():
    """A pair below the 1 MiB rule whose distance exceeds the first threshold: the bit-vectors are stored during the
    distance pass, and the pass is repeated (and re-stored) with doubled k until it succeeds."""
    rng = np.random.default_rng(77)
    t = bytes(util.BASES[i] for i in rng.integers(4, size=5000))
    q = util.mutate(rng, t[2000:2064], 0.1)
    _check([(q, t), (b"ACGTACGTAC", t[:4000])])
