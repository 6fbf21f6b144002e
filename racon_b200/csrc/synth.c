/*
 * synth.c — deterministic synthetic POA-window generator (bench / test input maker).
 *
 * Implements, bit for bit, the generator specified in SURVEY.md §8(d):
 *   splitmix64 (global state, never reseeded), u01 = (x>>11)*2^-53, base = "ACGT"[x&3],
 *   mutate(): per truth base draw u; u<pd delete; u<pd+pi insert a random base *before* it and
 *   re-draw for the same base; u<pd+pi+ps substitute with a different base; else copy.
 *   Per window: truth T = `truth_len` iid bases; backbone = mutate(T) truncated to `truth_len`;
 *   `depth` layers = independent mutate(T); all layers full-span (begin 0, end |B|-1), no
 *   qualities (weight 1), backbone dummy quality '!' (weight 0), window type kTGS.
 * The reference consumes such windows through racon::createWindow / Window::add_layer
 * (/root/reference/src/window.cpp:15-63).
 *
 * This file has no dependency on CUDA or on the checkers.  It is used by bench.py and tests/ to
 * build the flat "window set" arrays that every consumer (the C ABI and the checkers) reads.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
    uint64_t state;
} rp_rng;

static inline uint64_t rng_next(rp_rng* r) {
    r->state += 0x9E3779B97F4A7C15ull;
    uint64_t z = r->state;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

static inline double rng_u01(rp_rng* r) {
    return (double)(rng_next(r) >> 11) * (1.0 / 9007199254740992.0);
}

static const char kBases[4] = {'A', 'C', 'G', 'T'};

/* returns length written to `out` (capacity must be >= 4*len+16; insertion runs are geometric) */
static uint32_t mutate(rp_rng* r, const char* t, uint32_t len, double ps, double pi, double pd,
                       char* out, uint32_t cap) {
    uint32_t n = 0;
    for (uint32_t i = 0; i < len; ++i) {
        for (;;) {
            double u = rng_u01(r);
            if (u < pd) {
                break; /* deletion */
            } else if (u < pd + pi) {
                char c = kBases[rng_next(r) & 3];
                if (n < cap) out[n++] = c; /* insertion before t[i]; redo same i */
                continue;
            } else if (u < pd + pi + ps) {
                char c;
                do {
                    c = kBases[rng_next(r) & 3];
                } while (c == t[i]);
                if (n < cap) out[n++] = c;
                break;
            } else {
                if (n < cap) out[n++] = t[i];
                break;
            }
        }
    }
    return n;
}

/*
 * Generates `n_windows` windows.  Outputs (caller-allocated):
 *   bases      : concatenated sequence bytes, capacity `bases_cap`
 *   seq_off    : (n_windows*(depth+1) + 1) offsets into bases
 *   seq_begin / seq_end : per sequence layer coordinates (backbone entries are 0,0)
 *   win_first  : n_windows+1 indices into the per-sequence arrays
 * `state_io` carries the splitmix64 state across calls (start with 42).
 * Returns total bases written, or (uint64_t)-1 if `bases_cap` is too small.
 */
uint64_t rp_synth_windows(uint64_t* state_io, uint32_t n_windows, uint32_t truth_len,
                          uint32_t depth, double err, char* bases, uint64_t bases_cap,
                          uint64_t* seq_off, uint32_t* seq_begin, uint32_t* seq_end,
                          uint32_t* win_first) {
    rp_rng r;
    r.state = *state_io;
    const double p = err / 3.0;
    const uint32_t tmp_cap = 4 * truth_len + 64;
    char* truth = (char*)malloc(truth_len);
    char* tmp = (char*)malloc(tmp_cap);
    uint64_t nb = 0;
    uint64_t ns = 0;
    seq_off[0] = 0;
    for (uint32_t w = 0; w < n_windows; ++w) {
        win_first[w] = (uint32_t)ns;
        for (uint32_t k = 0; k < truth_len; ++k) truth[k] = kBases[rng_next(&r) & 3];
        uint32_t bl = mutate(&r, truth, truth_len, p, p, p, tmp, tmp_cap);
        if (bl > truth_len) bl = truth_len;
        if (bl == 0) { /* degenerate; keep one base so the window is valid */
            tmp[0] = truth[0];
            bl = 1;
        }
        if (nb + bl > bases_cap) goto fail;
        memcpy(bases + nb, tmp, bl);
        nb += bl;
        seq_begin[ns] = 0;
        seq_end[ns] = 0;
        seq_off[++ns] = nb;
        for (uint32_t d = 0; d < depth; ++d) {
            uint32_t rl = mutate(&r, truth, truth_len, p, p, p, tmp, tmp_cap);
            if (nb + rl > bases_cap) goto fail;
            memcpy(bases + nb, tmp, rl);
            nb += rl;
            seq_begin[ns] = 0;
            seq_end[ns] = bl - 1;
            seq_off[++ns] = nb;
        }
    }
    win_first[n_windows] = (uint32_t)ns;
    *state_io = r.state;
    free(truth);
    free(tmp);
    return nb;
fail:
    free(truth);
    free(tmp);
    return (uint64_t)-1;
}

/*
 * Illumina-mode windows (BASELINE.json config 4; SURVEY.md §8(d) config map): window length `w` (200), `depth` (60)
 * layers that are pieces of `read_len`-base (150) reads cut at the window edges — so every layer is partial-span and
 * goes through the Subgraph path of Window::generate_consensus (/root/reference/src/window.cpp:89-108) —, substitution
 * errors at rate `sub` (0.005) in the reads, Phred 30-40 qualities on every layer, window type kNGS (no trimming,
 * window.cpp:125).  The backbone is the truth with `bb_err` (0.01) errors, a third each substitutions / insertions /
 * deletions, which is what the reads have to polish away; layer coordinates are mapped through that edit script onto
 * the backbone, inclusive, as Polisher::initialize derives them from the breaking points (polisher.cpp:440-457).
 * Pieces shorter than 0.02 * w are dropped like polisher.cpp:415 drops them.  Same splitmix64 stream as above.
 * Outputs as rp_synth_windows plus `quals` (same offsets; backbone quality bytes are '!').  seq arrays need room for
 * n_windows * (depth + 1) + 1 entries; fewer sequences may be produced (win_first tells).  Returns bases written.
 */
uint64_t rp_synth_ngs_windows(uint64_t* state_io, uint32_t n_windows, uint32_t w, uint32_t depth, uint32_t read_len,
                              double sub, double bb_err, char* bases, char* quals, uint64_t bases_cap,
                              uint64_t* seq_off, uint32_t* seq_begin, uint32_t* seq_end, uint32_t* win_first) {
    rp_rng r;
    r.state = *state_io;
    char* truth = (char*)malloc(w);
    char* bb = (char*)malloc(4 * w + 64);
    uint32_t* map = (uint32_t*)malloc(sizeof(uint32_t) * (w + 1)); /* truth index -> backbone index of the base that
                                                                      represents it (or of the next kept base) */
    const double p = bb_err / 3.0;
    const uint32_t min_piece = (uint32_t)(0.02 * w) + 1;
    uint64_t nb = 0, ns = 0;
    seq_off[0] = 0;
    for (uint32_t wi = 0; wi < n_windows; ++wi) {
        win_first[wi] = (uint32_t)ns;
        for (uint32_t k = 0; k < w; ++k) truth[k] = kBases[rng_next(&r) & 3];
        /* backbone = edited truth, remembering where every truth base went */
        uint32_t bl = 0;
        for (uint32_t i = 0; i < w; ++i) {
            for (;;) {
                double u = rng_u01(&r);
                if (u < p) { /* deleted from the backbone */
                    map[i] = bl;
                    break;
                } else if (u < 2 * p) {
                    bb[bl++] = kBases[rng_next(&r) & 3];
                    continue;
                } else if (u < 3 * p) {
                    char c;
                    do {
                        c = kBases[rng_next(&r) & 3];
                    } while (c == truth[i]);
                    map[i] = bl;
                    bb[bl++] = c;
                    break;
                } else {
                    map[i] = bl;
                    bb[bl++] = truth[i];
                    break;
                }
            }
        }
        if (bl < 2) {
            bb[0] = truth[0];
            bb[1] = truth[1 % w];
            bl = 2;
            for (uint32_t i = 0; i < w; ++i) map[i] = i < 2 ? i : 1;
        }
        if (nb + bl > bases_cap) goto fail;
        memcpy(bases + nb, bb, bl);
        memset(quals + nb, '!', bl);
        nb += bl;
        seq_begin[ns] = 0;
        seq_end[ns] = 0;
        seq_off[++ns] = nb;
        for (uint32_t d = 0; d < depth; ++d) {
            /* a read_len-base read placed uniformly among the positions where it overlaps the window */
            const int64_t start = (int64_t)(rng_next(&r) % (uint64_t)(w + read_len - 1)) - (int64_t)(read_len - 1);
            int64_t a = start < 0 ? 0 : start;
            int64_t b = start + read_len - 1 >= (int64_t)w ? (int64_t)w - 1 : start + read_len - 1;
            if (b - a + 1 < (int64_t)min_piece) continue;
            uint32_t begin = map[a], end = map[b];
            if (end >= bl) end = bl - 1;
            if (begin >= end) continue;
            const uint32_t len = (uint32_t)(b - a + 1);
            if (nb + len > bases_cap) goto fail;
            for (uint32_t k = 0; k < len; ++k) {
                char c = truth[a + k];
                if (rng_u01(&r) < sub) {
                    char m;
                    do {
                        m = kBases[rng_next(&r) & 3];
                    } while (m == c);
                    c = m;
                }
                bases[nb + k] = c;
                quals[nb + k] = (char)(33 + 30 + (rng_next(&r) % 11));
            }
            nb += len;
            seq_begin[ns] = begin;
            seq_end[ns] = end;
            seq_off[++ns] = nb;
        }
    }
    win_first[n_windows] = (uint32_t)ns;
    *state_io = r.state;
    free(truth);
    free(bb);
    free(map);
    return nb;
fail:
    free(truth);
    free(bb);
    free(map);
    return (uint64_t)-1;
}

/* FNV-1a-64 over a byte range, continuing from `h` (start: 1469598103934665603). */
uint64_t rp_fnv1a64(uint64_t h, const unsigned char* p, uint64_t n) {
    for (uint64_t i = 0; i < n; ++i) {
        h ^= p[i];
        h *= 1099511628211ull;
    }
    return h;
}
