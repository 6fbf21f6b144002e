"""ctypes access to the TEST-ONLY host simulation of the device code (racon_b200/lib/libracon_sim.so)."""
import ctypes as C

import numpy as np

from racon_b200 import build

_lib = None


def lib():
    global _lib
    if _lib is None:
        l = C.CDLL(build.build_sim())
        l.rp_sim_poa.restype = C.c_int
        l.rp_sim_poa.argtypes = [C.c_uint32] + [C.c_void_p] * 8 + [C.c_int8, C.c_int8, C.c_int8, C.c_int,
                                                                    C.c_void_p, C.c_void_p, C.c_uint32,
                                                                    C.c_void_p, C.c_void_p, C.c_void_p,
                                                                    C.c_void_p, C.c_void_p]
        _lib = l
    return _lib


def _ptr(a):
    return None if a is None else a.ctypes.data


def sim_consensus(ws, m=3, x=-5, g=-4, trim=True, nmax=4096, lmax=1535, ki=16, ka=8, smem=14336, tile_rows=0, debug_flags=0):
    n = ws.n_windows
    lens = np.diff(ws.seq_off.astype(np.int64))
    stride = int(max(64, 2 * lens.max() + 64))
    out = np.zeros((n, stride), dtype=np.uint8)
    cov = np.zeros((n, stride), dtype=np.uint16)
    out_len = np.zeros(n, dtype=np.uint32)
    pol = np.zeros(n, dtype=np.uint8)
    st = np.zeros(n, dtype=np.uint32)
    stats = np.zeros(8, dtype=np.uint64)
    limits = np.asarray([nmax, lmax, ki, ka, smem, tile_rows, debug_flags], dtype=np.uint32)
    r = lib().rp_sim_poa(n, _ptr(ws.bases), _ptr(ws.quals), _ptr(ws.seq_off), _ptr(ws.seq_has_qual),
                         _ptr(ws.seq_begin), _ptr(ws.seq_end), _ptr(ws.win_first), _ptr(ws.win_type), m, x, g,
                         1 if trim else 0, limits.ctypes.data, out.ctypes.data, stride, out_len.ctypes.data,
                         pol.ctypes.data, st.ctypes.data, cov.ctypes.data, stats.ctypes.data)
    if r != 0:
        raise RuntimeError("rp_sim_poa failed: %d" % r)
    cons = [out[w, :out_len[w]].tobytes() for w in range(n)]
    covs = [cov[w, :out_len[w]].copy() for w in range(n)]
    return cons, pol.astype(bool), st, covs, stats
