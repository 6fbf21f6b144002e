"""Real-data parity (BASELINE config 1): the 96 windows the UNMODIFIED reference Polisher builds from its own
lambda-phage sample (test/data/sample_reads.fastq.gz + sample_overlaps.paf.gz + sample_layout.fasta.gz, w=500,
3/-5/-4), committed as tests/golden/lambda_windows.npz by tests/golden/make_lambda_windows.py together with the
reference's per-window consensus and final polished contig (racon stdout md5 b0e2a2788440a4982e544e2e9b3bf378,
the value SURVEY.md §8c records).  13 % of these layers are partial-span (Subgraph path), all carry qualities,
and 9 % are longer than 512 bases (multi-chunk rows)."""
import hashlib
import os

import numpy as np
import pytest

from oracle import bindings as ob
from racon_b200 import windows

MD5 = "b0e2a2788440a4982e544e2e9b3bf378"


def load():
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "lambda_windows.npz"))
    ws = windows.WindowSet(bases=z["bases"], quals=z["quals"], seq_off=z["seq_off"], seq_has_qual=z["seq_has_qual"],
                           seq_begin=z["seq_begin"], seq_end=z["seq_end"], win_first=z["win_first"],
                           win_type=z["win_type"])
    off = np.concatenate([[0], np.cumsum(z["cons_len"].astype(np.int64))])
    ref = [z["cons_flat"][off[i]:off[i + 1]].tobytes() for i in range(len(z["cons_len"]))]
    return ws, ref, z["polished"].tobytes(), z["polished_name"].tobytes().decode()


def load_frag():
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "lambda_frag_windows.npz"))
    ws = windows.WindowSet(bases=z["bases"], quals=z["quals"], seq_off=z["seq_off"], seq_has_qual=z["seq_has_qual"],
                           seq_begin=z["seq_begin"], seq_end=z["seq_end"], win_first=z["win_first"],
                           win_type=z["win_type"])
    off = np.concatenate([[0], np.cumsum(z["cons_len"].astype(np.int64))])
    return ws, [z["cons_flat"][off[i]:off[i + 1]].tobytes() for i in range(len(z["cons_len"]))]


def fasta_md5(name, cons):
    return hashlib.md5((">" + name + "\n").encode() + b"".join(cons) + b"\n").hexdigest()


def test_fixture_is_the_reference_output():
    ws, ref, polished, name = load()
    assert ws.n_windows == 96 and b"".join(ref) == polished
    assert fasta_md5(name, ref) == MD5
    lens = np.diff(ws.seq_off.astype(np.int64))
    assert (lens > 511).sum() > 50, "fixture should exercise multi-chunk rows"


def test_oracle_on_real_windows():
    ws, ref, _, name = load()
    cons, pol, _ = ob.oracle_consensus(ws, 3, -5, -4, threads=8)
    assert cons == ref and fasta_md5(name, cons) == MD5


def load_second_scores():
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "lambda_windows.npz"))
    off = np.concatenate([[0], np.cumsum(z["cons2_len"].astype(np.int64))])
    return [z["cons2_flat"][off[i]:off[i + 1]].tobytes() for i in range(len(z["cons2_len"]))]


def test_oracle_on_real_windows_racon_test_scores():
    """Scores 5/-4/-8 of test/racon_test.cpp:86-107 (the polished contig the generator checked against the
    reference's golden edit distance 1312)."""
    ws, _, _, _ = load()
    cons, _, _ = ob.oracle_consensus(ws, 5, -4, -8, threads=8)
    assert cons == load_second_scores()


def test_oracle_on_fragment_correction_windows():
    """racon -f (kF, all-vs-all overlaps, 1/-1/-1 as in test/racon_test.cpp:243-259, whose golden 236 reads /
    1658216 bases the generator asserted): first 200 windows of the reference run on the lambda reads."""
    ws, ref = load_frag()
    cons, pol, _ = ob.oracle_consensus(ws, 1, -1, -1, threads=8)
    assert cons == ref
    depth = np.diff(ws.win_first.astype(np.int64))
    assert (depth < 3).any() and not pol[depth < 3].any()  # backbone copies are reported as not polished


def test_sim_on_two_real_windows():
    from tests import simlib
    ws, ref, _, _ = load()
    # a shallow window at the contig end and the first window with a partial-span layer that is > 512 bases long
    depth = np.diff(ws.win_first.astype(np.int64))
    pick = [int(np.argmin(np.where(depth >= 4, depth, 10 ** 6)))]
    lens = np.diff(ws.seq_off.astype(np.int64))
    for w in range(ws.n_windows):
        s0, s1 = int(ws.win_first[w]), int(ws.win_first[w + 1])
        if depth[w] <= 16 and (lens[s0 + 1:s1] > 511).any() and w not in pick:
            pick.append(w)
            break
    sub = ws.subset(pick)
    cons, pol, st, _, _ = simlib.sim_consensus(sub, 3, -5, -4)
    assert (st == 0).all() and cons == [ref[w] for w in pick]


@pytest.mark.gpu
def test_gpu_polishes_the_lambda_contig_identically():
    from racon_b200 import api
    ws, ref, polished, name = load()
    cons, pol, st = api.consensus(ws, 3, -5, -4)
    assert (st == 0).all()
    bad = [w for w in range(ws.n_windows) if cons[w] != ref[w]]
    assert not bad, "windows differ: %s" % bad[:8]
    assert b"".join(cons) == polished and fasta_md5(name, cons) == MD5
    mcons, mpol = api.mirror_consensus(ws, 3, -5, -4)  # through createWindow / add_layer / BatchProcessor
    assert mcons == ref


@pytest.mark.gpu
def test_gpu_real_windows_racon_test_scores():
    from racon_b200 import api
    ws, _, _, _ = load()
    cons, pol, st = api.consensus(ws, 5, -4, -8)
    assert (st == 0).all() and cons == load_second_scores()


@pytest.mark.gpu
def test_gpu_fragment_correction_windows():
    from racon_b200 import api
    ws, ref = load_frag()
    cons, pol, st = api.consensus(ws, 1, -1, -1)
    assert (st == 0).all() and cons == ref


# ---- two more of the reference's goldens (tests/golden/make_lambda_more.py) -----------------------------------
def load_scores3():
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "lambda_windows_scores3.npz"))
    off = np.concatenate([[0], np.cumsum(z["cons_len"].astype(np.int64))])
    return [z["cons_flat"][off[i]:off[i + 1]].tobytes() for i in range(len(z["cons_len"]))]


def load_w1000():
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "lambda_w1000_windows.npz"))
    ws = windows.WindowSet(bases=z["bases"], quals=z["quals"], seq_off=z["seq_off"], seq_has_qual=z["seq_has_qual"],
                           seq_begin=z["seq_begin"], seq_end=z["seq_end"], win_first=z["win_first"],
                           win_type=z["win_type"])
    off = np.concatenate([[0], np.cumsum(z["cons_len"].astype(np.int64))])
    return ws, [z["cons_flat"][off[i]:off[i + 1]].tobytes() for i in range(len(z["cons_len"]))], z["polished"].tobytes()


def test_oracle_on_real_windows_edit_distance_scores():
    """1/-1/-1 on the w=500 windows: the reference run behind test/racon_test.cpp:202-223 (golden 1321)."""
    ws, _, _, _ = load()
    cons, _, _ = ob.oracle_consensus(ws, 1, -1, -1, threads=8)
    assert cons == load_scores3()


def test_oracle_on_real_larger_windows():
    """w=1000, 5/-4/-8: the reference run behind test/racon_test.cpp:179-200 (golden 1289).  Graphs of ~2 700 nodes
    with g=-8: several of these alignments are beyond spoa's worst-case int16 bound (its int32 engine)."""
    ws, ref, polished = load_w1000()
    assert ws.n_windows == 48 and b"".join(ref) == polished
    cons, _, _ = ob.oracle_consensus(ws, 5, -4, -8, window_length=1000, threads=8)
    assert cons == ref


def test_sim_on_a_real_larger_window():
    """Device code (warp simulation) on the shallowest real w=1000 window that is polished."""
    from tests import simlib
    ws, ref, _ = load_w1000()
    depth = np.diff(ws.win_first.astype(np.int64))
    w = int(np.argmin(np.where(depth >= 5, depth, 10 ** 6)))
    cons, pol, st, _, _ = simlib.sim_consensus(ws.subset([w]), 5, -4, -8, nmax=6 * 1000 + 64, lmax=2 * 1000 + 23)
    assert (st == 0).all() and cons == [ref[w]]


@pytest.mark.gpu
def test_gpu_real_windows_edit_distance_scores():
    from racon_b200 import api
    ws, _, _, _ = load()
    cons, pol, st = api.consensus(ws, 1, -1, -1)
    assert (st == 0).all() and cons == load_scores3()


@pytest.mark.gpu
def test_gpu_real_larger_windows():
    from racon_b200 import api
    ws, ref, polished = load_w1000()
    cons, pol, st = api.consensus(ws, 5, -4, -8, window_length=1000)
    assert (st == 0).all(), st
    bad = [w for w in range(ws.n_windows) if cons[w] != ref[w]]
    assert not bad, "windows differ: %s" % bad[:8]
    assert b"".join(cons) == polished


# ---- racon -b on the real windows: same outputs, with the band audit (every accepted band result recomputed with the
# full matrix on the device) reporting zero differing alignments -----------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("lanes", [8, 16])
def test_gpu_banded_real_windows_equal_the_reference(lanes, monkeypatch):
    from racon_b200 import api
    monkeypatch.setenv("RP_POA_GROUP", str(lanes))
    monkeypatch.setenv("RP_BAND_AUDIT", "1")
    ws, ref, polished, name = load()
    stats = {}
    cons, pol, st = api.consensus(ws, 3, -5, -4, banded=True, band_stats=stats)
    assert (st == 0).all()
    assert cons == ref and fasta_md5(name, cons) == MD5
    assert stats["band_alignments"] > 1000 and stats["band_audit_mismatches"] == 0, stats
    print("lambda contig, %d-column band: %s" % (16 * lanes, stats))
    stats = {}
    cons, pol, st = api.consensus(ws, 5, -4, -8, banded=True, band_stats=stats)
    assert (st == 0).all() and cons == load_second_scores() and stats["band_audit_mismatches"] == 0
    ws, ref = load_frag()
    stats = {}
    cons, pol, st = api.consensus(ws, 1, -1, -1, banded=True, band_stats=stats)
    assert (st == 0).all() and cons == ref and stats["band_audit_mismatches"] == 0, stats
    print("lambda -f windows, %d-column band: %s" % (16 * lanes, stats))


@pytest.mark.gpu
def test_gpu_banded_real_larger_windows():
    from racon_b200 import api
    ws, ref, polished = load_w1000()
    stats = {}
    cons, pol, st = api.consensus(ws, 5, -4, -8, window_length=1000, banded=True, band_stats=stats)
    assert (st == 0).all(), st
    assert cons == ref and b"".join(cons) == polished


@pytest.mark.parametrize("kb", [8, 4])
def test_sim_banded_on_real_windows(kb):
    """The banded device code (32 lanes x 8 / 4 columns) in the CPU simulation on real lambda windows — the deepest and two
    with partial-span layers — with the band audit on: same consensus as the unmodified reference, no accepted band
    alignment differs from the full-matrix one."""
    from tests import simlib
    ws, ref, _, _ = load()
    depth = np.diff(ws.win_first.astype(np.int64))
    pick = [int(np.argmax(depth))]
    for w in range(ws.n_windows):
        s0, s1 = int(ws.win_first[w]), int(ws.win_first[w + 1])
        blen = int(ws.seq_off[s0 + 1] - ws.seq_off[s0])
        if any(int(ws.seq_begin[s]) > 10 or int(ws.seq_end[s]) < blen - 10 for s in range(s0 + 1, s1)) and w not in pick:
            pick.append(w)
        if len(pick) == 3:
            break
    sub = ws.subset(pick)
    cons, pol, st, _, stats = simlib.sim_consensus(sub, 3, -5, -4, lanes=32, smem=9216, banded=1, band_cols_per_lane=kb,
                                                   debug_flags=2)
    assert (st == 0).all() and cons == [ref[w] for w in pick]
    assert stats[4] > 0 and stats[6] == 0, stats
