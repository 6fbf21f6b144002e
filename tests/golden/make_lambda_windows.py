"""Real-data fixture (BASELINE config 1): runs the UNMODIFIED reference Polisher (oracle/_ref/refpol_dump,
built by `make -C oracle refpol`) on the reference's own lambda-phage sample (test/data/sample_*), and stores
  * every window it built (backbone, layers, qualities, positions) as the flat window set, and
  * the reference's per-window consensus + the final polished contig,
in tests/golden/lambda_windows.npz.  Run in the CPU container:  python tests/golden/make_lambda_windows.py"""
import hashlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
DATA = "/root/reference/test/data/"


def run(reads, overlaps, target, fragment, wl, m, x, g, threads=8, q=10.0, e=0.3, trim=True):
    import struct
    import subprocess
    import tempfile
    exe = os.path.join(ROOT, "oracle", "_ref", "refpol_dump")
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "refpol"], stdout=subprocess.DEVNULL)
    with tempfile.TemporaryDirectory() as tmp:
        out = os.path.join(tmp, "dump.bin")
        subprocess.check_call([exe, reads, overlaps, target, str(fragment), str(wl), str(q), str(e),
                               "1" if trim else "0", str(m), str(x), str(g), str(threads), out],
                              stderr=subprocess.DEVNULL)
        raw = open(out, "rb").read()
    nw, ns, nb, npol = struct.unpack_from("<4Q", raw, 0)
    pos = 32

    def take(dtype, n):
        nonlocal pos
        a = np.frombuffer(raw, dtype=dtype, count=n, offset=pos).copy()
        pos += a.nbytes
        return a

    r = dict(bases=take(np.uint8, nb), quals=take(np.uint8, nb), seq_off=take(np.uint64, ns + 1),
             seq_has_qual=take(np.uint8, ns), seq_begin=take(np.uint32, ns), seq_end=take(np.uint32, ns),
             win_first=take(np.uint32, nw + 1), win_type=take(np.uint8, nw), win_target=take(np.uint64, nw),
             win_rank=take(np.uint32, nw))
    cons = []
    for _ in range(nw):
        (n,) = struct.unpack_from("<I", raw, pos)
        pos += 4
        cons.append(raw[pos:pos + n])
        pos += n
    polished = []
    for _ in range(npol):
        (nl,) = struct.unpack_from("<I", raw, pos)
        pos += 4
        name = raw[pos:pos + nl].decode()
        pos += nl
        (n,) = struct.unpack_from("<Q", raw, pos)
        pos += 8
        polished.append((name, raw[pos:pos + n]))
        pos += n
    r["cons"] = cons
    r["polished"] = polished
    return r


def revcomp(b):
    return b.translate(bytes.maketrans(b"ACGT", b"TGCA"))[::-1]


def main():
    import gzip
    from oracle import bindings as ob
    r = run(DATA + "sample_reads.fastq.gz", DATA + "sample_overlaps.paf.gz", DATA + "sample_layout.fasta.gz",
            0, 500, 3, -5, -4)
    cons_flat = np.frombuffer(b"".join(r["cons"]), np.uint8)
    cons_len = np.asarray([len(c) for c in r["cons"]], np.uint32)
    name, pol = r["polished"][0]
    fasta = (">" + name + "\n").encode() + pol + b"\n"
    md5 = hashlib.md5(fasta).hexdigest()
    print("windows", len(r["cons"]), "sequences", len(r["seq_begin"]), "bases", len(r["bases"]))
    print("polished:", name, len(pol), "racon stdout md5", md5)
    assert md5 == "b0e2a2788440a4982e544e2e9b3bf378"

    # same windows, the scores of test/racon_test.cpp:86-107 (5/-4/-8): golden edit distance 1312 to the reference
    r2 = run(DATA + "sample_reads.fastq.gz", DATA + "sample_overlaps.paf.gz", DATA + "sample_layout.fasta.gz",
             0, 500, 5, -4, -8)
    assert (r2["bases"] == r["bases"]).all() and (r2["seq_begin"] == r["seq_begin"]).all()
    ref_seq = b"".join(l.strip() for l in gzip.open(DATA + "sample_reference.fasta.gz").read().split(b"\n")[1:])
    pol2 = r2["polished"][0][1]
    _, ed = ob.ref_edlib_cigar(revcomp(pol2), ref_seq)
    print("5/-4/-8: edit distance to sample_reference =", ed)
    assert ed == 1312  # test/racon_test.cpp:104-106

    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "lambda_windows.npz")
    np.savez_compressed(out, bases=r["bases"], quals=r["quals"], seq_off=r["seq_off"], seq_has_qual=r["seq_has_qual"],
                        seq_begin=r["seq_begin"], seq_end=r["seq_end"], win_first=r["win_first"],
                        win_type=r["win_type"], win_target=r["win_target"], win_rank=r["win_rank"],
                        cons_flat=cons_flat, cons_len=cons_len, polished=np.frombuffer(pol, np.uint8),
                        polished_name=np.frombuffer(name.encode(), np.uint8),
                        scores=np.asarray([3, -5, -4], np.int8),
                        cons2_flat=np.frombuffer(b"".join(r2["cons"]), np.uint8),
                        cons2_len=np.asarray([len(c) for c in r2["cons"]], np.uint32),
                        scores2=np.asarray([5, -4, -8], np.int8))
    print("wrote", out, os.path.getsize(out), "bytes")

    # fragment correction (racon -f, all-vs-all overlaps) with the settings of test/racon_test.cpp:243-259
    # (kF, 1/-1/-1): golden 236 sequences / 1658216 bases.  Only the first 200 windows are committed (the full set
    # is 3461 windows / 17 MB); many of them have < 3 sequences (backbone copied, not polished).
    r = run(DATA + "sample_reads.fastq.gz", DATA + "sample_ava_overlaps.paf.gz", DATA + "sample_reads.fastq.gz",
            1, 500, 1, -1, -1)
    total = sum(len(p[1]) for p in r["polished"])
    print("fragment correction: total windows", len(r["cons"]), "->", len(r["polished"]), "reads,", total, "bases")
    assert len(r["polished"]) == 236 and total == 1658216
    nw = 200
    ns = int(r["win_first"][nw])
    nb = int(r["seq_off"][ns])
    cons = r["cons"][:nw]
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "lambda_frag_windows.npz")
    np.savez_compressed(out, bases=r["bases"][:nb], quals=r["quals"][:nb], seq_off=r["seq_off"][:ns + 1],
                        seq_has_qual=r["seq_has_qual"][:ns], seq_begin=r["seq_begin"][:ns], seq_end=r["seq_end"][:ns],
                        win_first=r["win_first"][:nw + 1], win_type=r["win_type"][:nw],
                        cons_flat=np.frombuffer(b"".join(cons), np.uint8),
                        cons_len=np.asarray([len(c) for c in cons], np.uint32),
                        total_windows=np.asarray([len(r["cons"])], np.uint32), scores=np.asarray([1, -1, -1], np.int8))
    print("wrote", out, os.path.getsize(out), "bytes")


if __name__ == "__main__":
    main()
