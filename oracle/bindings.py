"""ctypes bindings for the checkers under oracle/ — TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this.
  ref_*    -> oracle/_ref/libracon_ref.so   (unmodified reference sources, built by oracle/Makefile)
  oracle_* -> oracle/_build/libpoa_oracle.so (the restated CPU oracle, oracle/poa_oracle.cpp)
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SO = os.path.join(HERE, "_ref", "libracon_ref.so")
ORACLE_SO = os.path.join(HERE, "_build", "libpoa_oracle.so")

_ref = None
_oracle = None

_CONS_ARGS = [C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
              C.c_void_p, C.c_void_p, C.c_int8, C.c_int8, C.c_int8, C.c_uint32, C.c_int, C.c_uint32,
              C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p]


def have_ref():
    return os.path.exists(REF_SO)


def ref_lib():
    global _ref
    if _ref is None:
        lib = C.CDLL(REF_SO)
        lib.ref_poa_consensus.restype = C.c_double
        lib.ref_poa_consensus.argtypes = _CONS_ARGS
        lib.ref_edlib_cigar.restype = C.c_int64
        lib.ref_edlib_cigar.argtypes = [C.c_char_p, C.c_uint32, C.c_char_p, C.c_uint32, C.c_void_p,
                                        C.c_uint64, C.c_void_p]
        _ref = lib
    return _ref


def oracle_lib():
    global _oracle
    if _oracle is None:
        srcs = [os.path.join(HERE, f) for f in ("poa_oracle.cpp", "myers_oracle.cpp", "bp_oracle.cpp")]
        stale = not os.path.exists(ORACLE_SO) or any(
            os.path.exists(f) and os.path.getmtime(f) > os.path.getmtime(ORACLE_SO) for f in srcs)
        if stale:
            import subprocess
            subprocess.check_call(["make", "-C", HERE, "oracle"], stdout=subprocess.DEVNULL)
        lib = C.CDLL(ORACLE_SO)
        lib.oracle_poa_consensus.restype = C.c_double
        lib.oracle_poa_consensus.argtypes = _CONS_ARGS + [C.c_void_p]
        lib.oracle_poa_stats.restype = None
        lib.oracle_poa_stats.argtypes = [C.c_void_p]
        if hasattr(lib, "oracle_myers_cigar"):
            lib.oracle_myers_cigar.restype = C.c_int64
            lib.oracle_myers_cigar.argtypes = [C.c_char_p, C.c_uint32, C.c_char_p, C.c_uint32, C.c_void_p,
                                               C.c_uint64, C.c_void_p]
        if hasattr(lib, "oracle_breaking_points"):
            lib.oracle_breaking_points.restype = C.c_int64
            lib.oracle_breaking_points.argtypes = [C.c_char_p, C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32,
                                                   C.c_uint32, C.c_void_p, C.c_uint64]
        _oracle = lib
    return _oracle


def _ptr(a):
    return None if a is None else a.ctypes.data


def _max_nodes_bound(ws):
    # consensus length <= number of graph nodes <= total bases of the window
    lens = np.diff(ws.seq_off.astype(np.int64))
    per_win = np.add.reduceat(lens, ws.win_first[:-1].astype(np.int64)) if ws.n_windows else np.zeros(0)
    return int(per_win.max()) if len(per_win) else 1


def _run(fn, ws, m, x, g, window_length, trim, threads, extra=()):
    n = ws.n_windows
    stride = max(16, _max_nodes_bound(ws))
    out = np.zeros((n, stride), dtype=np.uint8)
    out_len = np.zeros(n, dtype=np.uint32)
    polished = np.zeros(n, dtype=np.uint8)
    secs = fn(n, _ptr(ws.bases), _ptr(ws.quals), _ptr(ws.seq_off), _ptr(ws.seq_has_qual),
              _ptr(ws.seq_begin), _ptr(ws.seq_end), _ptr(ws.win_first), _ptr(ws.win_type),
              m, x, g, window_length, 1 if trim else 0, threads, out.ctypes.data, stride,
              out_len.ctypes.data, polished.ctypes.data, *extra)
    if secs < 0:
        raise RuntimeError("checker failed (code %r)" % secs)
    cons = [out[w, :out_len[w]].tobytes() for w in range(n)]
    return cons, polished.astype(bool), secs


def ref_consensus(ws, m=3, x=-5, g=-4, window_length=500, trim=True, threads=1):
    """The reference itself: Window::generate_consensus over every window."""
    return _run(ref_lib().ref_poa_consensus, ws, m, x, g, window_length, trim, threads)


def oracle_consensus(ws, m=3, x=-5, g=-4, window_length=500, trim=True, threads=1, want_coverage=False):
    """The restated oracle.  With want_coverage also returns per-window uint32 coverage (untrimmed
    consensus coordinates are not exposed; coverage is for the returned, trimmed consensus)."""
    n = ws.n_windows
    stride = max(16, _max_nodes_bound(ws))
    cov = np.zeros((n, stride), dtype=np.uint32)
    cons, pol, secs = _run(oracle_lib().oracle_poa_consensus, ws, m, x, g, window_length, trim, threads,
                           extra=(cov.ctypes.data,))
    if want_coverage:
        return cons, pol, secs, [cov[w, :len(cons[w])].copy() for w in range(n)]
    return cons, pol, secs


def ref_edlib_cigar(q, t):
    cap = 4 * (len(q) + len(t)) + 64
    buf = C.create_string_buffer(cap)
    ed = C.c_int32(0)
    n = ref_lib().ref_edlib_cigar(q, len(q), t, len(t), buf, cap, C.byref(ed))
    if n < 0:
        raise RuntimeError("edlib failed (%d)" % n)
    return buf.raw[:n].decode(), ed.value


def oracle_myers_cigar(q, t):
    cap = 4 * (len(q) + len(t)) + 64
    buf = C.create_string_buffer(cap)
    ed = C.c_int32(0)
    n = oracle_lib().oracle_myers_cigar(q, len(q), t, len(t), buf, cap, C.byref(ed))
    if n < 0:
        raise RuntimeError("oracle myers failed (%d)" % n)
    return buf.raw[:n].decode(), ed.value


def oracle_breaking_points(cigar, t_begin, t_end, q_start, window_length):
    """(n, 2) uint32 array of (t, q) breaking points, two rows per window with a match (overlap.cpp:226-292)."""
    if isinstance(cigar, str):
        cigar = cigar.encode()
    cap = 2 * ((t_end - t_begin) // max(1, window_length) + 3)
    out = np.zeros((cap, 2), np.uint32)
    n = oracle_lib().oracle_breaking_points(cigar, len(cigar), t_begin, t_end, q_start, window_length,
                                            out.ctypes.data, cap)
    if n < 0:
        raise RuntimeError("oracle breaking points failed (%d)" % n)
    return out[:n].copy()


CUDAPOA_SO = os.path.join(HERE, "_ref", "libref_cudapoa.so")


def have_ref_cudapoa():
    return os.path.exists(CUDAPOA_SO)


def ref_cudapoa_consensus(ws, match=3, mismatch=-5, gap=-4, banded=False, max_depth=200, device=0, mem_fraction=0.5):
    """The reference's own GPU path (GenomeWorks cudapoa, unmodified; oracle/ref_cudapoa_harness.cu) on a window set.
    Returns (list of consensus bytes — b"" where cudapoa rejected the window —, n_ok, wall seconds, GPU-call seconds).
    Timing only: cudapoa's consensus is not spoa's."""
    lib = C.CDLL(CUDAPOA_SO)
    lib.ref_cudapoa_consensus.restype = C.c_int64
    lib.ref_cudapoa_consensus.argtypes = [C.c_uint32] + [C.c_void_p] * 6 + [C.c_int8, C.c_int8, C.c_int8, C.c_int,
                                                                            C.c_uint32, C.c_int, C.c_double, C.c_void_p,
                                                                            C.c_uint32, C.c_void_p, C.c_void_p]
    n = ws.n_windows
    stride = 2048
    out = np.zeros((n, stride), np.uint8)
    out_len = np.zeros(n, np.uint32)
    times = np.zeros(2, np.float64)
    seq_off = np.ascontiguousarray(ws.seq_off, dtype=np.uint64)
    ok = lib.ref_cudapoa_consensus(n, _ptr(ws.bases), _ptr(ws.quals), _ptr(seq_off), _ptr(ws.seq_has_qual),
                                   _ptr(ws.seq_begin), _ptr(ws.win_first), match, mismatch, gap, 1 if banded else 0,
                                   max_depth, device, mem_fraction, _ptr(out), stride, _ptr(out_len), _ptr(times))
    if ok < 0:
        raise RuntimeError("reference cudapoa failed (%d)" % ok)
    return [out[i, :out_len[i]].tobytes() for i in range(n)], int(ok), float(times[0]), float(times[1])



def ref_cudapoa_multi(ws, match=3, mismatch=-5, gap=-4, banded=False, max_depth=200, n_devices=1, batches_per_device=2):
    """The reference's own multi-GPU driver shape (CUDAPolisher::polish, cudapolisher.cpp:226-345) over GenomeWorks
    cudapoa, unmodified: `batches_per_device` batch objects on each of the first n_devices GPUs, one host thread each.
    Returns (windows reported ok, wall seconds).  Timing only."""
    lib = C.CDLL(CUDAPOA_SO)
    lib.ref_cudapoa_multi.restype = C.c_int64
    lib.ref_cudapoa_multi.argtypes = [C.c_uint32] + [C.c_void_p] * 6 + [C.c_int8, C.c_int8, C.c_int8, C.c_int,
                                                                       C.c_uint32, C.c_int, C.c_int, C.c_void_p]
    seq_off = ws.seq_off.astype(np.uint64)
    times = np.zeros(2, np.float64)
    ok = lib.ref_cudapoa_multi(ws.n_windows, _ptr(ws.bases), _ptr(ws.quals), _ptr(seq_off), _ptr(ws.seq_has_qual),
                               _ptr(ws.seq_begin), _ptr(ws.win_first), match, mismatch, gap, 1 if banded else 0,
                               max_depth, n_devices, batches_per_device, times.ctypes.data)
    if ok < 0:
        raise RuntimeError("reference cudapoa (multi) failed (%d)" % ok)
    return int(ok), float(times[0])
