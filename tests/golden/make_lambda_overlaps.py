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


def main():
    exe = os.path.join(ROOT, "oracle", "_ref", "refpol_dump")
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "refpol"], stdout=subprocess.DEVNULL)
    with tempfile.TemporaryDirectory() as tmp:
        dump = os.path.join(tmp, "ov.bin")
        env = dict(os.environ, REFPOL_OVERLAP_DUMP=dump)
        subprocess.check_call([exe, DATA + "sample_reads.fastq.gz", DATA + "sample_overlaps.paf.gz",
                               DATA + "sample_layout.fasta.gz", "0", "500", "10.0", "0.3", "1", "3", "-5", "-4", "8",
                               os.path.join(tmp, "w.bin")], env=env, stderr=subprocess.DEVNULL)
        raw = open(dump, "rb").read()
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
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "lambda_overlaps.npz")
    np.savez_compressed(out, bases=np.frombuffer(b"".join(bases), np.uint8), quals=np.frombuffer(b"".join(quals), np.uint8),
                        seq_off=np.asarray(seq_off, np.uint64), seq_has_qual=np.asarray(has_qual, np.uint8),
                        overlaps=ov, bp_off=np.asarray(bp_off, np.uint64),
                        bp=np.concatenate(bps) if bps else np.zeros((0, 2), np.uint32),
                        params=np.asarray([500, 10.0, 0.3], np.float64))
    print("sequences", nseq, "bases", seq_off[-1], "overlaps", nov, "breaking point pairs", bp_off[-1] // 2)
    print("wrote", out, os.path.getsize(out), "bytes")


if __name__ == "__main__":
    main()
