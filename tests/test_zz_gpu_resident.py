"""Windows by reference into a device-resident read store (layer extraction on the device;
include/racon_b200.h rp_reads_* / rp_poa_add_window_refs): every layer is a slice of a longer sequence that was uploaded
once — forward, or a slice of its reverse complement with reversed qualities — and the packed arrays the POA kernel reads
are gathered on the device.  Results must equal the pointer path (rp_poa_add_window) and the oracle, byte for byte."""
import numpy as np
import pytest

from oracle import bindings as ob
from racon_b200 import api
from tests import util

pytestmark = pytest.mark.gpu

COMP = bytes.maketrans(b"ACGT", b"TGCA")


def _as_store(ws, seed, junk=b"ACGT", reverse_frac=0.5):
    """Hides every sequence of a window set inside a longer 'read' (random flanks; about half of the layers stored as the
    reverse complement, the way racon reads a '-' strand overlap) and all backbones inside one 'contig'.
    Returns (sequences, qualities, pieces per window)."""
    rng = np.random.default_rng(seed)
    seqs, quals, pieces = [], [], []
    contig, contig_q, any_bq = b"", b"", False
    starts = []
    for w in range(ws.n_windows):
        b, q, _, _ = ws.window(w)[0]
        starts.append(len(contig))
        contig += b
        contig_q += q if q else b"!" * len(b)
        any_bq |= q is not None
    seqs.append(contig)
    quals.append(contig_q if any_bq else None)
    for w in range(ws.n_windows):
        win = ws.window(w)
        pw = [(0, starts[w], len(win[0][0]), 0, 0, 0)]
        for (b, q, beg, end) in win[1:]:
            pre = bytes(rng.choice(list(junk), size=int(rng.integers(0, 30))).astype(np.uint8))
            suf = bytes(rng.choice(list(junk), size=int(rng.integers(0, 30))).astype(np.uint8))
            read = pre + b + suf
            rq = None
            if q is not None:
                rq = bytes(rng.integers(34, 70, size=len(pre)).astype(np.uint8)) + q + \
                    bytes(rng.integers(34, 70, size=len(suf)).astype(np.uint8))
            rev = int(rng.random() < reverse_frac)
            if rev:   # the store holds the other strand; the window reads it back as a reverse complement
                read = read.translate(COMP)[::-1]
                rq = rq[::-1] if rq is not None else None
            seqs.append(read)
            quals.append(rq)
            pw.append((len(seqs) - 1, len(pre), len(b), rev, beg, end))
        pieces.append(pw)
    return seqs, quals, pieces


def _by_reference(ws, seqs, quals, pieces, scores=(3, -5, -4), trim=True, mem_bytes=0, window_length=500):
    store = api.ReadStore(seqs, quals)
    batch = api.PoaBatch(match=scores[0], mismatch=scores[1], gap=scores[2], window_length=window_length,
                         mem_bytes=mem_bytes)
    stride = int(2 * np.diff(ws.seq_off.astype(np.int64)).max() + 64)
    cons, pol, st, batches, h2d = [], [], [], 0, 0
    try:
        w = 0
        while w < ws.n_windows:
            batch.reset()
            first = w
            while w < ws.n_windows:
                r = batch.add_window_refs(store, pieces[w], window_type=int(ws.win_type[w]), trim=trim)
                if r == api.RP_BATCH_FULL:
                    break
                w += 1
            assert w > first, "a window does not fit an empty batch"
            batch.run()
            batch.sync()
            h2d += batch.info()["h2d_bytes"]
            out, lens, p, s = batch.fetch_all(stride)
            cons += [out[i, :lens[i]].tobytes() for i in range(w - first)]
            pol.append(p)
            st.append(s)
            batches += 1
    finally:
        batch.close()
        store.close()
    return cons, np.concatenate(pol), np.concatenate(st), batches, h2d


CASES = {
    "fullspan": dict(n=10, wlen=160, depth=12, err=0.12),
    "partial_qual": dict(n=10, wlen=150, depth=10, err=0.12, partial_frac=0.4, with_qual=True, backbone_qual=True),
    "ngs": dict(n=12, wlen=100, depth=16, err=0.02, partial_frac=0.9, with_qual=True, types=0),
    "acgtn": dict(n=8, wlen=120, depth=10, err=0.2, partial_frac=0.2, alphabet=b"ACGTN"),
    "shallow": dict(n=12, wlen=60, depth=3, err=0.2, partial_frac=0.3),
}


def _mk(name, seed):
    kw = dict(CASES[name])
    n = kw.pop("n")
    t = kw.pop("types", None)
    ws = util.make_set(seed, n, **kw)
    if t is not None:
        ws.win_type[:] = t
    return ws


@pytest.mark.parametrize("name", sorted(CASES))
def test_windows_by_reference_equal_pointer_path_and_oracle(name):
    ws = _mk(name, seed=21)
    seqs, quals, pieces = _as_store(ws, seed=5)
    cons, pol, st, _, h2d = _by_reference(ws, seqs, quals, pieces)
    ora, opol, _ = ob.oracle_consensus(ws, 3, -5, -4, trim=True, threads=4)
    ptr, ppol, pst = api.consensus(ws, 3, -5, -4, trim=True)
    assert (st == 0).all() and (pst == 0).all()
    assert cons == ora == ptr
    assert (pol == opol).all() and (ppol == opol).all()
    # no sequence byte went over the bus with the windows — descriptors and metadata only (by pointer: 2 bytes per base)
    assert h2d < int(ws.seq_off[-1])


def test_bulk_add_over_a_store_made_of_the_window_set_itself():
    """bench.py's shape: the flat arrays of a window set are the store (every packed sequence one 'read'), the windows
    are added in bulk by reference, several batches under a small budget."""
    from racon_b200 import windows
    ws = _mk("partial_qual", seed=4)
    store = api.ReadStore.from_flat(ws.bases, ws.seq_off, ws.quals, ws.seq_has_qual)
    refs = windows.as_refs(ws)
    batch = api.PoaBatch(window_length=200, mem_bytes=48 << 20)
    stride = int(2 * np.diff(ws.seq_off.astype(np.int64)).max() + 64)
    cons, first, rounds = [], 0, 0
    try:
        assert store.device_bytes() >= 2 * len(ws.bases)
        while first < ws.n_windows:
            batch.reset()
            took = batch.add_window_set_refs(store, refs, first)
            assert took > 0
            batch.run()
            batch.sync()
            out, lens, _, st = batch.fetch_all(stride)
            assert (st == 0).all()
            cons += [out[i, :lens[i]].tobytes() for i in range(took)]
            first += took
            rounds += 1
    finally:
        batch.close()
        store.close()
    ora, _, _ = ob.oracle_consensus(ws, 3, -5, -4, trim=True, threads=4)
    assert cons == ora


def test_by_reference_other_scores_no_trim_and_junk_flanks():
    """Flanks full of characters the windows do not contain: the per-read character sets overflow the 8-code alphabet,
    the pieces are then scanned exactly — same result as by pointer."""
    ws = _mk("partial_qual", seed=9)
    seqs, quals, pieces = _as_store(ws, seed=6, junk=b"ACGTNRYKMSWBDHV")
    for scores, trim in (((1, -1, -1), True), ((5, -4, -8), False)):
        cons, pol, st, _, _ = _by_reference(ws, seqs, quals, pieces, scores=scores, trim=trim)
        ora, opol, _ = ob.oracle_consensus(ws, *scores, trim=trim, threads=4)
        assert (st == 0).all() and cons == ora and (pol == opol).all()


def test_by_reference_budget_back_pressure_and_statuses():
    from racon_b200 import windows
    ws = _mk("fullspan", seed=3)
    seqs, quals, pieces = _as_store(ws, seed=8)
    store = api.ReadStore(seqs, quals)
    batch = api.PoaBatch(window_length=200)
    try:
        # a window with < 3 sequences: backbone copy, not polished (window.cpp:68-71) — needs the host bytes of the store
        assert batch.add_window_refs(store, pieces[0][:2]) == 0
        # a mix of the two ways to add windows is refused, and so is a piece outside its sequence
        with pytest.raises(api.RaconB200Error, match="call order"):
            batch.add_window(ws.window(1))
        with pytest.raises(api.RaconB200Error, match="outside its sequence"):
            batch.add_window_refs(store, [pieces[1][0], (1, len(seqs[1]) - 3, 10, 0, 0, 9)] + pieces[1][1:])
        with pytest.raises(api.RaconB200Error, match="backbone"):
            batch.add_window_refs(store, [(0, 0, 100, 1, 0, 0)] + pieces[1][1:])
        assert batch.add_window_refs(store, pieces[1]) == 0
        batch.run()
        batch.sync()
        c0, _, p0 = batch.fetch(0)[0], None, batch.fetch(0)[2]
        assert c0 == ws.window(0)[0][0] and not p0
        ora, _, _ = ob.oracle_consensus(ws.subset([1]), 3, -5, -4, trim=True, threads=1)
        assert batch.fetch(1)[0] == ora[0]
        # after a reset the object takes windows by pointer again
        batch.reset()
        assert batch.add_window(ws.window(1)) == 0
        batch.run()
        batch.sync()
        assert batch.fetch(0)[0] == ora[0]
    finally:
        batch.close()
        store.close()
    # nine distinct characters inside one window: reported, like the pointer path
    bb = bytes(np.random.default_rng(1).choice(list(b"ACGT"), size=90).astype(np.uint8))
    odd = bytearray(bb)
    odd[10:15] = b"NRYKM"
    w9 = windows.from_lists([[(bb, None, 0, 0), (bytes(odd), None, 0, 89), (bb, None, 0, 89), (bb, None, 0, 89)]])
    s9, q9, p9 = _as_store(w9, seed=2)
    _, pol, st, _, _ = _by_reference(w9, s9, q9, p9)
    assert (st != 0).all() and not pol.any()


def test_aligner_by_reference_equals_by_pointer_and_edlib():
    """rp_aln_add_overlap_ref: both spans of an overlap cut out of the store on the device, the query as a slice of the
    reverse complement for a '-' strand overlap — CIGAR, edit distance and breaking points equal to the by-pointer call on
    the same bytes and to the unmodified edlib (when it was built here)."""
    rng = np.random.default_rng(17)
    reads, pairs = [], []
    for k in range(10):
        t = bytes(rng.choice(list(b"ACGT"), size=int(rng.integers(200, 700))).astype(np.uint8))
        q = util.mutate(rng, t, 0.12) or t[:1]
        pre_t = bytes(rng.choice(list(b"ACGT"), size=int(rng.integers(0, 600))).astype(np.uint8))
        pre_q = bytes(rng.choice(list(b"ACGTN"), size=int(rng.integers(0, 40))).astype(np.uint8))
        suf_q = bytes(rng.choice(list(b"ACGT"), size=int(rng.integers(0, 40))).astype(np.uint8))
        rev = k % 2
        stored_q = pre_q + q + suf_q
        if rev:
            stored_q = stored_q.translate(COMP)[::-1]
        reads += [pre_t + t + b"GATTACA", stored_q]
        pairs.append(dict(q=q, t=t, q_id=2 * k + 1, q_start=len(pre_q), rev=rev, t_id=2 * k, t_begin=len(pre_t)))
    store = api.ReadStore(reads)
    a, b = api.AlnBatch(), api.AlnBatch()
    try:
        for batch in (a, b):
            batch.set_window_length(100)
        for p in pairs:
            assert a.add_ref(store, p["q_id"], p["q_start"], len(p["q"]), p["rev"], p["t_id"], p["t_begin"], len(p["t"]))
            assert b.add(p["q"], p["t"], t_begin=p["t_begin"], q_start=p["q_start"])
        with pytest.raises(api.RaconB200Error, match="call order"):
            a.add(pairs[0]["q"], pairs[0]["t"])
        with pytest.raises(api.RaconB200Error, match="outside its sequence"):
            a.add_ref(store, 1, len(reads[1]) - 5, 10, 0, 0, 0, 10)
        for batch in (a, b):
            batch.run()
            batch.sync()
        assert a.info()["h2d_bytes"] < b.info()["h2d_bytes"] - sum(len(p["q"]) + len(p["t"]) for p in pairs) // 2
        for i, p in enumerate(pairs):
            ca, cb = a.fetch(i), b.fetch(i)
            assert ca == cb and ca[2] == 0
            assert (a.fetch_breaking_points(i) == b.fetch_breaking_points(i)).all()
            if ob.have_ref():
                assert ca[0].decode() == ob.ref_edlib_cigar(p["q"], p["t"])[0]
    finally:
        a.close()
        b.close()
        store.close()
