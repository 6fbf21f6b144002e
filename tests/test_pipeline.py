"""Whole device-facing pipeline on real data (BASELINE config 1, lambda phage): the reference's own sequences and
filtered overlaps (tests/golden/lambda_overlaps.npz) go through racon_b200::Polisher — device alignment, device
breaking points, window assembly, device consensus, stitching — and every stage is compared with what the UNMODIFIED
reference Polisher produced on the same input (tests/golden/lambda_windows.npz): the window set field by field, each
window's consensus, and the polished contig (racon's golden md5)."""
import hashlib
import os

import numpy as np
import pytest

from tests.lambda_overlaps import LambdaOverlaps

pytestmark = pytest.mark.gpu


def test_pipeline_lambda_matches_reference_polisher():
    from racon_b200 import api
    lam = LambdaOverlaps()
    ref = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "lambda_windows.npz"))
    pol = api.MirrorPolisher(lam.bases, lam.quals, lam.seq_off, lam.seq_has_qual, n_targets=1, overlaps=lam.ov,
                             window_length=lam.window_length, quality_threshold=lam.quality_threshold, trim=True,
                             match=3, mismatch=-5, gap=-4, window_type_tgs=bool(ref["win_type"][0]))
    got = pol.export()
    for k in ("win_first", "win_type", "win_target", "win_rank", "seq_off", "seq_begin", "seq_end", "seq_has_qual",
              "bases", "quals"):
        assert np.array_equal(got[k], ref[k]), k
    cons, polished = pol.polish()
    pol.close()
    off = np.concatenate([[0], np.cumsum(ref["cons_len"].astype(np.int64))])
    flat = ref["cons_flat"].tobytes()
    for w, c in enumerate(cons):
        assert c == flat[off[w]:off[w + 1]], w
    assert len(polished) == 1
    tid, tags, data = polished[0]
    name = ref["polished_name"].tobytes().decode()
    assert tid == 0 and name.endswith(tags)
    assert data == ref["polished"].tobytes()
    fasta = (">" + name + "\n").encode() + data + b"\n"
    assert hashlib.md5(fasta).hexdigest() == "b0e2a2788440a4982e544e2e9b3bf378"   # racon's stdout on this sample


def test_pipeline_lambda_fragment_correction_matches_reference():
    """racon -f on the same sample (BASELINE config 5's flow; test/racon_test.cpp:243-259 settings 1/-1/-1): 7780
    all-vs-all overlaps on both strands -> 3461 windows -> 236 corrected reads.  Checked against the unmodified
    reference: its breaking points for every overlap, the first 200 windows field by field with their consensus, and
    the reference's golden totals (236 sequences, 1 658 216 bases) plus the md5 of all corrected reads and their tags."""
    from racon_b200 import api
    lam = LambdaOverlaps("lambda_frag_overlaps.npz")
    ref = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "lambda_frag_windows.npz"))
    m, x, g = (int(v) for v in lam.z["scores"])
    n_seq = len(lam.seq_off) - 1
    pol = api.MirrorPolisher(lam.bases, lam.quals, lam.seq_off, lam.seq_has_qual, n_targets=n_seq, overlaps=lam.ov,
                             window_length=lam.window_length, quality_threshold=lam.quality_threshold, trim=True,
                             match=m, mismatch=x, gap=g, window_type_tgs=bool(ref["win_type"][0]),
                             fragment_correction=True)
    got = pol.export()
    assert len(got["win_type"]) == int(ref["total_windows"][0])
    nw = len(ref["win_type"])
    ns = int(ref["win_first"][nw])
    nb = int(ref["seq_off"][ns])
    assert np.array_equal(got["win_first"][:nw + 1], ref["win_first"])
    for k in ("seq_begin", "seq_end", "seq_has_qual"):
        assert np.array_equal(got[k][:ns], ref[k]), k
    assert np.array_equal(got["seq_off"][:ns + 1], ref["seq_off"])
    assert np.array_equal(got["bases"][:nb], ref["bases"]) and np.array_equal(got["quals"][:nb], ref["quals"])
    cons, polished = pol.polish()
    pol.close()
    off = np.concatenate([[0], np.cumsum(ref["cons_len"].astype(np.int64))])
    flat = ref["cons_flat"].tobytes()
    for w in range(nw):
        assert cons[w] == flat[off[w]:off[w + 1]], w
    assert len(polished) == int(lam.z["polished_count"][0]) == 236
    assert sum(len(p[2]) for p in polished) == int(lam.z["polished_bases"][0]) == 1658216
    assert hashlib.md5(b"".join(p[2] for p in polished)).hexdigest() == lam.z["polished_md5"].tobytes().decode()
    tags = lam.z["polished_tags"].tobytes().decode().split("\n")
    assert ["r" + t for t in tags] == [p[1] for p in polished] or tags == [p[1][1:] for p in polished]


def test_pipeline_streaming_stitch_writes_the_golden_fasta(tmp_path):
    """the stitch + FASTA writer as a streaming consumer — two batch objects in flight (a 96 MB budget each
    forces several batches over the 96 windows), every sequence written the moment its last window is collected; the file
    is byte-identical to racon's stdout on the sample (md5 b0e2a278...), with and without -b."""
    from racon_b200 import api
    lam = LambdaOverlaps()
    ref = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "lambda_windows.npz"))
    name = ref["polished_name"].tobytes().decode().split(" ")[0]
    for banded in (False, True):
        pol = api.MirrorPolisher(lam.bases, lam.quals, lam.seq_off, lam.seq_has_qual, n_targets=1, overlaps=lam.ov,
                                 window_length=lam.window_length, quality_threshold=lam.quality_threshold, trim=True,
                                 match=3, mismatch=-5, gap=-4, window_type_tgs=bool(ref["win_type"][0]))
        out = str(tmp_path / ("polished_%d.fasta" % banded))
        n = pol.stream_fasta(out, [name], drop_unpolished=True, mem_bytes=96 << 20, banded=banded)
        assert pol.failed() == (0, 0)
        pol.close()
        assert n == 1
        assert hashlib.md5(open(out, "rb").read()).hexdigest() == "b0e2a2788440a4982e544e2e9b3bf378"


def test_pipeline_streaming_stitch_many_batches_fragment_correction(tmp_path):
    """Same consumer on the -f flow with a 40 MB budget per batch object: the 3461 windows need a dozen batches, which
    alternate between the two objects; the 236 corrected reads come out in target order with the reference's bytes."""
    from racon_b200 import api
    lam = LambdaOverlaps("lambda_frag_overlaps.npz")
    ref = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "lambda_frag_windows.npz"))
    m, x, g = (int(v) for v in lam.z["scores"])
    n_seq = len(lam.seq_off) - 1
    pol = api.MirrorPolisher(lam.bases, lam.quals, lam.seq_off, lam.seq_has_qual, n_targets=n_seq, overlaps=lam.ov,
                             window_length=lam.window_length, quality_threshold=lam.quality_threshold, trim=True,
                             match=m, mismatch=x, gap=g, window_type_tgs=bool(ref["win_type"][0]),
                             fragment_correction=True)
    out = str(tmp_path / "corrected.fasta")
    n = pol.stream_fasta(out, ["read%d" % i for i in range(n_seq)], drop_unpolished=False, mem_bytes=40 << 20)
    assert pol.failed() == (0, 0)
    pol.close()
    assert n == 236
    lines = open(out, "rb").read().split(b"\n")
    data = b"".join(lines[1::2])
    assert len(data) == 1658216
    assert hashlib.md5(data).hexdigest() == lam.z["polished_md5"].tobytes().decode()
