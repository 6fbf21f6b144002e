"""Loader of tests/golden/lambda_overlaps.npz (the reference's own sequences, overlaps and breaking points on the
lambda-phage sample; made by tests/golden/make_lambda_overlaps.py)."""
import os

import numpy as np

_RC = bytes.maketrans(b"ACGT", b"TGCA")


def revcomp(b):
    return b.translate(_RC)[::-1]


class LambdaOverlaps:
    def __init__(self, name="lambda_overlaps.npz"):
        z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", name))
        self.z = z
        self.bases = z["bases"].tobytes()
        self.quals = z["quals"].tobytes()
        self.seq_off = z["seq_off"].astype(np.int64)
        self.seq_has_qual = z["seq_has_qual"]
        self.ov = z["overlaps"].astype(np.int64)   # q_id t_id strand q_begin q_end q_length t_begin t_end t_length
        self.bp_off = z["bp_off"].astype(np.int64)
        self.bp = z["bp"]
        self.window_length, self.quality_threshold, self.error_threshold = z["params"]
        self.window_length = int(self.window_length)
        self._rc = {}

    def n_overlaps(self):
        return len(self.ov)

    def seq(self, i):
        return self.bases[self.seq_off[i]:self.seq_off[i + 1]]

    def qual(self, i):
        return self.quals[self.seq_off[i]:self.seq_off[i + 1]]

    def rc(self, i):
        if i not in self._rc:
            self._rc[i] = revcomp(self.seq(i))
        return self._rc[i]

    def spans(self, k):
        """(query span, target span, t_begin, t_end, q_start) of overlap k as overlap.cpp:193-197 hands them to edlib."""
        q_id, t_id, strand, q_begin, q_end, q_length, t_begin, t_end, _ = (int(v) for v in self.ov[k])
        if strand:
            q_start = q_length - q_end
            q = self.rc(q_id)[q_start:q_start + (q_end - q_begin)]
        else:
            q_start = q_begin
            q = self.seq(q_id)[q_begin:q_end]
        t = self.seq(t_id)[t_begin:t_end]
        return q, t, t_begin, t_end, q_start

    def expected_bp(self, k):
        return self.bp[self.bp_off[k]:self.bp_off[k + 1]]
