"""Flat "window set" arrays — the one input format every consumer reads.

A window set describes many racon windows (reference: racon::Window, /root/reference/src/window.hpp:31-74):
sequence 0 of each window is the backbone, the rest are layers with inclusive backbone coordinates
(begin, end) exactly as passed to Window::add_layer (/root/reference/src/window.cpp:42-63).
"""
import ctypes as C
import dataclasses
import os

import numpy as np

from . import build


@dataclasses.dataclass
class WindowSet:
    bases: np.ndarray          # uint8, concatenated sequence bytes
    quals: np.ndarray          # uint8 (same offsets) or None
    seq_off: np.ndarray        # uint64, n_seq + 1
    seq_has_qual: np.ndarray   # uint8, n_seq (or None)
    seq_begin: np.ndarray      # uint32, n_seq
    seq_end: np.ndarray        # uint32, n_seq
    win_first: np.ndarray      # uint32, n_windows + 1
    win_type: np.ndarray       # uint8, n_windows (0 = kNGS, 1 = kTGS)

    @property
    def n_windows(self):
        return len(self.win_first) - 1

    @property
    def n_seqs(self):
        return len(self.seq_off) - 1

    def window(self, w):
        """Returns [(bases_bytes, quals_bytes_or_None, begin, end), ...] for window w."""
        out = []
        for s in range(int(self.win_first[w]), int(self.win_first[w + 1])):
            a, b = int(self.seq_off[s]), int(self.seq_off[s + 1])
            q = None
            if self.quals is not None and self.seq_has_qual is not None and self.seq_has_qual[s]:
                q = self.quals[a:b].tobytes()
            out.append((self.bases[a:b].tobytes(), q, int(self.seq_begin[s]), int(self.seq_end[s])))
        return out

    def subset(self, idx):
        """New WindowSet holding windows `idx` (in that order)."""
        wins = [self.window(int(w)) for w in idx]
        types = [int(self.win_type[int(w)]) for w in idx]
        return from_lists(wins, types)


def from_lists(windows, types=None):
    """windows: list of [(bases, quals|None, begin, end), ...]; backbone first."""
    bases, quals, off, hasq, beg, end, first = [], [], [0], [], [], [], [0]
    any_q = False
    for win in windows:
        for (b, q, s, e) in win:
            b = b if isinstance(b, bytes) else b.encode()
            bases.append(b)
            if q is not None:
                q = q if isinstance(q, bytes) else q.encode()
                assert len(q) == len(b)
                quals.append(q)
                any_q = True
            else:
                quals.append(b"!" * len(b))
            hasq.append(0 if q is None else 1)
            off.append(off[-1] + len(b))
            beg.append(s)
            end.append(e)
        first.append(len(off) - 1)
    if types is None:
        types = [1] * len(windows)
    return WindowSet(
        bases=np.frombuffer(b"".join(bases), dtype=np.uint8).copy(),
        quals=np.frombuffer(b"".join(quals), dtype=np.uint8).copy() if any_q else None,
        seq_off=np.asarray(off, dtype=np.uint64),
        seq_has_qual=np.asarray(hasq, dtype=np.uint8) if any_q else None,
        seq_begin=np.asarray(beg, dtype=np.uint32),
        seq_end=np.asarray(end, dtype=np.uint32),
        win_first=np.asarray(first, dtype=np.uint32),
        win_type=np.asarray(types, dtype=np.uint8),
    )


_synth = None


def _synth_lib():
    global _synth
    if _synth is None:
        path = build.build_synth()
        lib = C.CDLL(path)
        lib.rp_synth_windows.restype = C.c_uint64
        lib.rp_synth_windows.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_double,
                                         C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p,
                                         C.c_void_p]
        lib.rp_synth_ngs_windows.restype = C.c_uint64
        lib.rp_synth_ngs_windows.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_double,
                                             C.c_double, C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p,
                                             C.c_void_p, C.c_void_p]
        lib.rp_fnv1a64.restype = C.c_uint64
        lib.rp_fnv1a64.argtypes = [C.c_uint64, C.c_void_p, C.c_uint64]
        _synth = lib
    return _synth


def synth_windows(n_windows, truth_len=500, depth=32, err=0.12, state=42):
    """SURVEY.md §8(d) generator.  Returns (WindowSet, new_state)."""
    lib = _synth_lib()
    n_seq = n_windows * (depth + 1)
    cap = int(n_seq * truth_len * (1.0 + err) + 4096 * (depth + 1))
    bases = np.empty(cap, dtype=np.uint8)
    seq_off = np.empty(n_seq + 1, dtype=np.uint64)
    beg = np.empty(n_seq, dtype=np.uint32)
    end = np.empty(n_seq, dtype=np.uint32)
    first = np.empty(n_windows + 1, dtype=np.uint32)
    st = C.c_uint64(state)
    nb = lib.rp_synth_windows(C.byref(st), n_windows, truth_len, depth, float(err), bases.ctypes.data, cap,
                              seq_off.ctypes.data, beg.ctypes.data, end.ctypes.data, first.ctypes.data)
    if nb == 2 ** 64 - 1:
        raise RuntimeError("synthetic generator: capacity too small")
    ws = WindowSet(bases=bases[:nb].copy(), quals=None, seq_off=seq_off, seq_has_qual=None, seq_begin=beg,
                   seq_end=end, win_first=first, win_type=np.ones(n_windows, dtype=np.uint8))
    return ws, st.value


def synth_ngs_windows(n_windows, w=200, depth=60, read_len=150, sub=0.005, bb_err=0.01, state=4242):
    """BASELINE config 4 shape (Illumina mode): w-base windows, `depth` partial-span pieces of read_len-base reads with
    qualities, kNGS.  Returns (WindowSet, new_state).  Generator: racon_b200/csrc/synth.c rp_synth_ngs_windows."""
    lib = _synth_lib()
    n_seq = n_windows * (depth + 1)
    cap = int(n_windows * (2 * w + 64 + depth * read_len))
    bases = np.empty(cap, dtype=np.uint8)
    quals = np.empty(cap, dtype=np.uint8)
    seq_off = np.empty(n_seq + 1, dtype=np.uint64)
    beg = np.empty(n_seq, dtype=np.uint32)
    end = np.empty(n_seq, dtype=np.uint32)
    first = np.empty(n_windows + 1, dtype=np.uint32)
    st = C.c_uint64(state)
    nb = lib.rp_synth_ngs_windows(C.byref(st), n_windows, w, depth, read_len, float(sub), float(bb_err),
                                  bases.ctypes.data, quals.ctypes.data, cap, seq_off.ctypes.data, beg.ctypes.data,
                                  end.ctypes.data, first.ctypes.data)
    if nb == 2 ** 64 - 1:
        raise RuntimeError("synthetic NGS generator: capacity too small")
    ns = int(first[n_windows])
    hasq = np.ones(ns, dtype=np.uint8)
    hasq[first[:-1]] = 0                  # backbones carry the dummy '!' quality (polisher.cpp:174,396-399)
    ws = WindowSet(bases=bases[:nb].copy(), quals=quals[:nb].copy(), seq_off=seq_off[:ns + 1].copy(), seq_has_qual=hasq,
                   seq_begin=beg[:ns].copy(), seq_end=end[:ns].copy(), win_first=first,
                   win_type=np.zeros(n_windows, dtype=np.uint8))
    return ws, st.value


def window_costs(ws):
    """SURVEY.md §8(e) cost estimate per window: sum over layers s of L_s * (len_backbone + 0.1 * sum_{k<s} L_k)."""
    lens = np.diff(ws.seq_off.astype(np.int64)).astype(np.float64)
    first = ws.win_first.astype(np.int64)
    nw = ws.n_windows
    win_of = np.repeat(np.arange(nw), np.diff(first))
    blen = lens[first[:-1]]
    csum = np.cumsum(lens)
    before = csum - lens - (csum[first[:-1]] - lens[first[:-1]])[win_of]    # bases of the window before this sequence
    prior_layers = before - blen[win_of]                                     # layers only (excludes the backbone)
    per_seq = lens * (blen[win_of] + 0.1 * np.maximum(prior_layers, 0.0))
    per_seq[first[:-1]] = 0.0
    return np.bincount(win_of, weights=per_seq, minlength=nw)


def slice_windows(ws, lo, hi):
    """Contiguous windows [lo, hi) of a WindowSet as a new WindowSet (vectorised; subset() is for small picks)."""
    s0, s1 = int(ws.win_first[lo]), int(ws.win_first[hi])
    b0, b1 = int(ws.seq_off[s0]), int(ws.seq_off[s1])
    return WindowSet(bases=ws.bases[b0:b1].copy(), quals=None if ws.quals is None else ws.quals[b0:b1].copy(),
                     seq_off=(ws.seq_off[s0:s1 + 1] - ws.seq_off[s0]).astype(np.uint64),
                     seq_has_qual=None if ws.seq_has_qual is None else ws.seq_has_qual[s0:s1].copy(),
                     seq_begin=ws.seq_begin[s0:s1].copy(), seq_end=ws.seq_end[s0:s1].copy(),
                     win_first=(ws.win_first[lo:hi + 1] - ws.win_first[lo]).astype(np.uint32),
                     win_type=ws.win_type[lo:hi].copy())


def as_refs(ws):
    """The window set described by reference (api.PoaBatch.add_window_set_refs) into a read store built over its own
    arrays (api.ReadStore.from_flat(ws.bases, ws.seq_off, ws.quals, ws.seq_has_qual)): sequence s of the set is
    sequence s of the store, taken whole and forward."""
    n = ws.n_seqs
    return dict(seq_id=np.arange(n, dtype=np.uint32), offset=np.zeros(n, np.uint32),
                length=np.ascontiguousarray(np.diff(ws.seq_off.astype(np.int64)).astype(np.uint32)),
                reverse=np.zeros(n, np.uint8), begin=np.ascontiguousarray(ws.seq_begin, dtype=np.uint32),
                end=np.ascontiguousarray(ws.seq_end, dtype=np.uint32),
                win_first=np.ascontiguousarray(ws.win_first, dtype=np.uint32),
                win_type=np.ascontiguousarray(ws.win_type, dtype=np.uint8))


def fnv1a64(chunks):
    lib = _synth_lib()
    h = 1469598103934665603
    for c in chunks:
        if len(c):
            buf = np.frombuffer(c, dtype=np.uint8) if isinstance(c, (bytes, bytearray)) else np.ascontiguousarray(c)
            h = lib.rp_fnv1a64(h, buf.ctypes.data, buf.size)
    return h
