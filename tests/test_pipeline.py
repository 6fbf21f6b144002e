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
