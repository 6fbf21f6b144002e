"""Two more of the reference's own goldens (test/racon_test.cpp) on the lambda sample, produced by the UNMODIFIED
reference Polisher (oracle/_ref/refpol_dump):
  * ConsensusWithQualitiesLargerWindow (:179-200): w = 1000, scores 5/-4/-8 -> edit distance 1289 to the reference
    genome; the windows and the reference's per-window consensus go to tests/golden/lambda_w1000_windows.npz.  With
    g = -8 and ~2 700-node graphs this is where spoa's worst-case bound switches to its int32 engine.
  * ConsensusWithQualitiesEditDistance (:202-223): w = 500, scores 1/-1/-1 -> 1321; same windows as
    lambda_windows.npz, only the per-window consensus is stored (tests/golden/lambda_windows_scores3.npz).
Run in the CPU container:  python tests/golden/make_lambda_more.py"""
import gzip
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from make_lambda_windows import DATA, revcomp, run  # noqa: E402


def main():
    from oracle import bindings as ob
    ref_seq = b"".join(l.strip() for l in gzip.open(DATA + "sample_reference.fasta.gz").read().split(b"\n")[1:])
    here = os.path.dirname(os.path.abspath(__file__))

    r = run(DATA + "sample_reads.fastq.gz", DATA + "sample_overlaps.paf.gz", DATA + "sample_layout.fasta.gz",
            0, 1000, 5, -4, -8)
    pol = r["polished"][0][1]
    ed = ob.ref_edlib_cigar(revcomp(pol), ref_seq)[1]
    print("w=1000 5/-4/-8: windows", len(r["cons"]), "edit distance", ed)
    assert ed == 1289   # test/racon_test.cpp:197
    out = os.path.join(here, "lambda_w1000_windows.npz")
    np.savez_compressed(out, bases=r["bases"], quals=r["quals"], seq_off=r["seq_off"], seq_has_qual=r["seq_has_qual"],
                        seq_begin=r["seq_begin"], seq_end=r["seq_end"], win_first=r["win_first"],
                        win_type=r["win_type"], cons_flat=np.frombuffer(b"".join(r["cons"]), np.uint8),
                        cons_len=np.asarray([len(c) for c in r["cons"]], np.uint32),
                        polished=np.frombuffer(pol, np.uint8), scores=np.asarray([5, -4, -8], np.int8),
                        window_length=np.asarray([1000], np.uint32))
    print("wrote", out, os.path.getsize(out), "bytes")

    r3 = run(DATA + "sample_reads.fastq.gz", DATA + "sample_overlaps.paf.gz", DATA + "sample_layout.fasta.gz",
             0, 500, 1, -1, -1)
    base = np.load(os.path.join(here, "lambda_windows.npz"))
    assert (r3["bases"] == base["bases"]).all() and (r3["seq_begin"] == base["seq_begin"]).all()
    ed3 = ob.ref_edlib_cigar(revcomp(r3["polished"][0][1]), ref_seq)[1]
    print("w=500 1/-1/-1: edit distance", ed3)
    assert ed3 == 1321  # test/racon_test.cpp:220
    out = os.path.join(here, "lambda_windows_scores3.npz")
    np.savez_compressed(out, cons_flat=np.frombuffer(b"".join(r3["cons"]), np.uint8),
                        cons_len=np.asarray([len(c) for c in r3["cons"]], np.uint32),
                        scores=np.asarray([1, -1, -1], np.int8))
    print("wrote", out, os.path.getsize(out), "bytes")


if __name__ == "__main__":
    main()
