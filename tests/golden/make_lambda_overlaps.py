"""Real-data fixture for the pre-alignment -> breaking points -> windows rows: runs the UNMODIFIED
reference Polisher (oracle/_ref/refpol_dump with $REFPOL_OVERLAP_DUMP, see oracle/ref_polisher_harness.cpp) on the
lambda-phage sample (BASELINE config 1: sample_reads.fastq.gz + sample_overlaps.paf.gz + sample_layout.fasta.gz) and
stores, in tests/golden/lambda_overlaps.npz,
  * the sequences as the reference holds them (target first, then reads; bases + qualities),
  * every overlap that survived the reference's filters (ids, strand, coordinates) and
  * the breaking points the reference derived from its edlib CIGAR for each of them.
The windows the reference then builds from these are tests/golden/lambda_windows.npz (make_lambda_windows.py).
Run in the CPU container:  python tests/golden/make_lambda_overlaps.py"""
import os
import struct
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
DATA = "/root/reference/test/data/"


def polished_of(path):
    """(count, total bases, md5 of the concatenated polished data) from refpol_dump's main output file"""
    import hashlib
    raw = open(path, "rb").read()
    nw, ns, nb, npol = struct.unpack_from("<4Q", raw, 0)
    pos = 32 + 2 * nb + 8 * (ns + 1) + ns + 4 * ns + 4 * ns + 4 * (nw + 1) + nw + 8 * nw + 4 * nw
    for _ in range(nw):
        (n,) = struct.unpack_from("<I", raw, pos)
        pos += 4 + n
    h = hashlib.md5()
    total = 0
    tags = []
    for _ in range(npol):
        (nl,) = struct.unpack_from("<I", raw, pos)
        pos += 4
        name = raw[pos:pos + nl].decode()
        pos += nl
        (n,) = struct.unpack_from("<Q", raw, pos)
        pos += 8
        h.update(raw[pos:pos + n])
        total += n
        pos += n
        tags.append(name[name.index(" LN:i:"):] if " LN:i:" in name else "")
    return npol, total, h.hexdigest(), tags


def run(reads, overlaps, target, fragment, scores, out_name):
    exe = os.path.join(ROOT, "oracle", "_ref", "refpol_dump")
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "refpol"], stdout=subprocess.DEVNULL)
    with tempfile.TemporaryDirectory() as tmp:
        dump = os.path.join(tmp, "ov.bin")
        env = dict(os.environ, REFPOL_OVERLAP_DUMP=dump)
        subprocess.check_call([exe, DATA + reads, DATA + overlaps, DATA + target, str(fragment), "500", "10.0", "0.3",
                               "1", str(scores[0]), str(scores[1]), str(scores[2]), "8", os.path.join(tmp, "w.bin")],
                              env=env, stderr=subprocess.DEVNULL)
        raw = open(dump, "rb").read()
        npol, total, md5, tags = polished_of(os.path.join(tmp, "w.bin"))
    pos = 0

    def u32():
        nonlocal pos
        (v,) = struct.unpack_from("<I", raw, pos)
        pos += 4
        return v

    def u64():
        nonlocal pos
        (v,) = struct.unpack_from("<Q", raw, pos)
        pos += 8
        return v

    def blob(n):
        nonlocal pos
        b = raw[pos:pos + n]
        pos += n
        return b

    nseq = u64()
    bases, quals, seq_off, has_qual = [], [], [0], []
    for _ in range(nseq):
        blob(u32())  # name
        d = blob(u64())
        q = blob(u64())
        bases.append(d)
        quals.append(q if q else b"!" * len(d))
        has_qual.append(1 if q else 0)
        seq_off.append(seq_off[-1] + len(d))
    nov = u64()
    ov = np.zeros((nov, 9), np.uint32)
    bp_off, bps = [0], []
    for i in range(nov):
        ov[i] = [u32() for _ in range(9)]
        n = u32()
        bps.append(np.frombuffer(blob(8 * n), np.uint32).reshape(n, 2).copy())
        bp_off.append(bp_off[-1] + n)
    assert pos == len(raw)
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), out_name)
    np.savez_compressed(out, bases=np.frombuffer(b"".join(bases), np.uint8), quals=np.frombuffer(b"".join(quals), np.uint8),
                        seq_off=np.asarray(seq_off, np.uint64), seq_has_qual=np.asarray(has_qual, np.uint8),
                        overlaps=ov, bp_off=np.asarray(bp_off, np.uint64),
                        bp=np.concatenate(bps) if bps else np.zeros((0, 2), np.uint32),
                        params=np.asarray([500, 10.0, 0.3], np.float64), scores=np.asarray(scores, np.int8),
                        polished_count=np.asarray([npol], np.uint64), polished_bases=np.asarray([total], np.uint64),
                        polished_md5=np.frombuffer(md5.encode(), np.uint8),
                        polished_tags=np.frombuffer("\n".join(tags).encode(), np.uint8))
    print(out_name, "sequences", nseq, "bases", seq_off[-1], "overlaps", nov, "breaking point pairs", bp_off[-1] // 2,
          "polished", npol, total, md5)
    print("wrote", out, os.path.getsize(out), "bytes")
    return npol, total


def main():
    run("sample_reads.fastq.gz", "sample_overlaps.paf.gz", "sample_layout.fasta.gz", 0, (3, -5, -4),
        "lambda_overlaps.npz")
    # fragment correction (racon -f, all-vs-all overlaps, both strands), the settings of test/racon_test.cpp:243-259
    npol, total = run("sample_reads.fastq.gz", "sample_ava_overlaps.paf.gz", "sample_reads.fastq.gz", 1, (1, -1, -1),
                      "lambda_frag_overlaps.npz")
    assert npol == 236 and total == 1658216   # the reference's own golden for this configuration


if __name__ == "__main__":
    main()
