/*
 * poa_pack.hpp — host-side packing of racon windows into the flat HBM layout the POA kernel reads.
 *
 * Mirrors, on the host, the parts of the reference that decide WHAT is aligned and in which order:
 *   racon::createWindow / Window::add_layer validation and skipping  (src/window.cpp:15-63)
 *   "< 3 sequences => consensus = backbone, false"                     (src/window.cpp:68-71)
 *   layer order = std::sort of indices by begin position (unstable!)   (src/window.cpp:79-86)
 *   full-span test begin < 0.01*len && end > len - 0.01*len            (src/window.cpp:88-94)
 *   weights = quality - 33, or 1 without quality                       (vendor/spoa/src/graph.cpp:121-146)
 * Pure C++ (no CUDA): shared by the product library (pinned buffers) and the test-only host simulation.
 */
#pragma once
#include <stdint.h>

#if defined(__SSE2__)
#include <emmintrin.h>
#endif
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

namespace rp {

/* growable POD buffer whose storage comes from a pluggable allocator (pinned host memory in the product) */
struct HostAllocator {
    void* (*alloc)(size_t);
    void (*release)(void*);
};

template <typename T>
struct GrowBuf {
    T* data = nullptr;
    size_t size = 0, cap = 0;
    const HostAllocator* al = nullptr;
    explicit GrowBuf(const HostAllocator* a) : al(a) {}
    GrowBuf(const GrowBuf&) = delete;
    GrowBuf& operator=(const GrowBuf&) = delete;
    ~GrowBuf() {
        if (data) al->release(data);
    }
    bool reserve(size_t n) {
        if (n <= cap) return true;
        size_t nc = cap ? cap : 1024;
        while (nc < n) nc *= 2;
        T* nd = static_cast<T*>(al->alloc(nc * sizeof(T)));
        if (!nd) return false;
        if (size) std::memcpy(nd, data, size * sizeof(T));
        if (data) al->release(data);
        data = nd;
        cap = nc;
        return true;
    }
    bool push(const T& v) {
        if (!reserve(size + 1)) return false;
        data[size++] = v;
        return true;
    }
    T* extend(size_t n) {  // returns pointer to n new (uninitialised) elements
        if (!reserve(size + n)) return nullptr;
        T* p = data + size;
        size += n;
        return p;
    }
    void clear() { size = 0; }
    size_t bytes() const { return size * sizeof(T); }
};

enum PackResult { kPackOk = 0, kPackFull = 1, kPackInvalid = -1, kPackNoMem = -3 };

struct PackedBatch {
    /* device-bound arrays */
    GrowBuf<uint8_t> bases, weights, seq_flags, win_flags;
    GrowBuf<uint32_t> seq_off, seq_begin, seq_end, win_first, out_off, out_cap, queue;
    GrowBuf<uint64_t> win_alpha;
    /* by-reference batches (layers named as slices of a device-resident read store, rp_reads): instead of bases and
     * weights, per packed sequence the store position of the byte that becomes its first base and how to read it */
    GrowBuf<uint64_t> src_pos;
    GrowBuf<uint8_t> src_flags;            // kSrcReverse | kSrcHasQuality | kSrcBackbone
    bool by_ref = false;
    uint64_t n_bases = 0;                  // bases of all packed sequences (== bases.size unless by_ref)
    /* host-only bookkeeping, one entry per ADDED window (in add order) */
    std::vector<int32_t> gpu_index;        // index among GPU windows, or -1 (trivial), -2 (alphabet limit)
    std::vector<std::string> trivial;      // consensus of trivial windows (backbone copy), indexed by added order
    std::vector<uint32_t> cost;            // per GPU window: total bases (queue ordering)
    uint64_t out_total = 0;
    /* limits */
    uint64_t max_bases = 0xfff00000ull;    // uint32 offsets; rp_poa_create lowers these to the object's memory budget
    uint64_t max_out = 0xfff00000ull;      // consensus bytes reserved (uint32 out_off)
    uint32_t max_windows = 1u << 22;
    uint32_t max_seq_len = 65000;

    explicit PackedBatch(const HostAllocator* a)
        : bases(a), weights(a), seq_flags(a), win_flags(a), seq_off(a), seq_begin(a), seq_end(a), win_first(a),
          out_off(a), out_cap(a), queue(a), win_alpha(a), src_pos(a), src_flags(a) {
        reset();
    }

    void reset() {
        bases.clear(); weights.clear(); seq_flags.clear(); win_flags.clear();
        seq_off.clear(); seq_begin.clear(); seq_end.clear(); win_first.clear();
        out_off.clear(); out_cap.clear(); queue.clear(); win_alpha.clear();
        src_pos.clear(); src_flags.clear();
        by_ref = false;
        n_bases = 0;
        seq_off.push(0);
        win_first.push(0);
        gpu_index.clear();
        trivial.clear();
        cost.clear();
        out_total = 0;
    }

    uint32_t n_added() const { return static_cast<uint32_t>(gpu_index.size()); }
    uint32_t n_gpu() const { return static_cast<uint32_t>(win_flags.size); }

    /* A window is packed in three steps so that the bulk path can run the byte-heavy ones on many host
     * threads: prepare() (validation, layer order, alphabet — no shared state), commit() (serial: reserves
     * space and writes the per-sequence / per-window metadata), fill() (copies bases and weights). */
    struct Prep {
        int status = kPackOk;          // kPackOk or kPackInvalid
        int kind = 0;                  // 0 GPU window, -1 trivial (< 3 sequences), -2 alphabet limit
        uint32_t blen = 0;
        uint64_t tot = 0;
        uint64_t alpha = 0;
        std::vector<uint32_t> order;   // caller's sequence indices in processing order (backbone first)
        std::vector<uint8_t> full;     // per entry of `order`: full-span flag
        uint64_t base_off = 0;         // set by commit(): where fill() writes
    };

    static void prepare(Prep& pr, uint32_t n_seq, const char* const* seq, const uint32_t* len, const uint32_t* begin,
                        const uint32_t* end, uint32_t max_seq_len) {
        pr = Prep();
        if (!seq) {
            pr.status = kPackInvalid;
            return;
        }
        prepare_layout(pr, n_seq, seq, len, begin, end, max_seq_len);
        if (pr.status != kPackOk || pr.kind != 0) return;
        /* alphabet: the distinct characters of the window (any order: only equality matters on the device) */
        uint8_t seen[256];
        std::memset(seen, 0, sizeof(seen));
        for (uint32_t i = 0; i < pr.order.size(); ++i)
            scan_alphabet(reinterpret_cast<const uint8_t*>(seq[pr.order[i]]), len[pr.order[i]], seen);
        alphabet_from_seen(pr, seen);
    }

    /* validation, skipped layers, layer order, full-span flags — everything of prepare() that does not look at the bases.
     * seq may be NULL (by-reference windows): then no per-sequence pointer is checked. */
    static void prepare_layout(Prep& pr, uint32_t n_seq, const char* const* seq, const uint32_t* len,
                               const uint32_t* begin, const uint32_t* end, uint32_t max_seq_len) {
        if (n_seq == 0 || !len || (seq && !seq[0]) || len[0] == 0 || len[0] > max_seq_len) {  // window.cpp:19-23
            pr.status = kPackInvalid;
            return;
        }
        const uint32_t blen = len[0];
        pr.blen = blen;
        /* Window::add_layer, window.cpp:42-63 */
        std::vector<uint32_t> kept;
        kept.reserve(n_seq);
        kept.push_back(0);
        uint64_t tot = blen;
        for (uint32_t k = 1; k < n_seq; ++k) {
            if (len[k] == 0 || begin[k] == end[k]) continue;
            if ((seq && !seq[k]) || begin[k] >= end[k] || begin[k] > blen || end[k] > blen || len[k] > max_seq_len) {
                pr.status = kPackInvalid;
                return;
            }
            kept.push_back(k);
            tot += len[k];
        }
        pr.tot = tot;
        if (kept.size() < 3) {  // window.cpp:68-71
            pr.kind = -1;
            return;
        }
        /* layer order, window.cpp:79-86 — the same std::sort call on the same element type */
        std::vector<uint32_t> rank(kept.size());
        for (uint32_t i = 0; i < rank.size(); ++i) rank[i] = i;
        std::sort(rank.begin() + 1, rank.end(),
                  [&](uint32_t l, uint32_t r) { return begin[kept[l]] < begin[kept[r]]; });
        const uint32_t offset = static_cast<uint32_t>(0.01 * blen);  // window.cpp:88
        pr.order.resize(rank.size());
        pr.full.resize(rank.size());
        for (uint32_t i = 0; i < rank.size(); ++i) {
            uint32_t k = kept[rank[i]];
            bool full = i > 0 && begin[k] < offset && end[k] > blen - offset;
            if (i > 0 && !full && end[k] >= blen) {  // Subgraph(begin, end) needs backbone node `end`
                pr.status = kPackInvalid;
                return;
            }
            pr.order[i] = k;
            pr.full[i] = full ? 1 : 0;
        }
    }

    static void alphabet_from_seen(Prep& pr, const uint8_t* seen) {
        if (seen[0]) {
            pr.status = kPackInvalid;
            return;
        }
        uint32_t ncodes = 0;
        for (uint32_t c = 1; c < 256; ++c) {
            if (!seen[c]) continue;
            if (ncodes == 8) {
                pr.kind = -2;
                return;
            }
            pr.alpha |= static_cast<uint64_t>(c) << (8 * ncodes++);
        }
    }

    /* seen[c] = 1 for every byte value c of sq[0..n).  Sequences are almost always over the four bases: 16 bytes at a
     * time are compared with A, C, G and T, and only a block holding anything else is walked byte by byte. */
    static void scan_alphabet(const uint8_t* sq, uint32_t n, uint8_t* seen) {
        uint32_t j = 0;
#if defined(__SSE2__)
        const __m128i ca = _mm_set1_epi8('A'), cc = _mm_set1_epi8('C'), cg = _mm_set1_epi8('G'), ct = _mm_set1_epi8('T');
        int any_a = 0, any_c = 0, any_g = 0, any_t = 0;
        for (; j + 16 <= n; j += 16) {
            const __m128i v = _mm_loadu_si128(reinterpret_cast<const __m128i*>(sq + j));
            const int ma = _mm_movemask_epi8(_mm_cmpeq_epi8(v, ca)), mc = _mm_movemask_epi8(_mm_cmpeq_epi8(v, cc));
            const int mg = _mm_movemask_epi8(_mm_cmpeq_epi8(v, cg)), mt = _mm_movemask_epi8(_mm_cmpeq_epi8(v, ct));
            any_a |= ma; any_c |= mc; any_g |= mg; any_t |= mt;
            if ((ma | mc | mg | mt) != 0xffff)
                for (uint32_t t = 0; t < 16; ++t) seen[sq[j + t]] = 1;
        }
        if (any_a) seen[static_cast<uint8_t>('A')] = 1;
        if (any_c) seen[static_cast<uint8_t>('C')] = 1;
        if (any_g) seen[static_cast<uint8_t>('G')] = 1;
        if (any_t) seen[static_cast<uint8_t>('T')] = 1;
#endif
        for (; j < n; ++j) seen[sq[j]] = 1;
    }

    /* serial; returns kPackOk / kPackFull / kPackNoMem (or pr.status when the window is malformed) */
    int commit(Prep& pr, const char* const* seq, const uint32_t* len, const uint32_t* begin, const uint32_t* end,
               int window_type, int trim) {
        if (pr.status != kPackOk) return pr.status;
        if (pr.kind == -1) {
            gpu_index.push_back(-1);
            trivial.emplace_back(seq[0], pr.blen);
            return kPackOk;
        }
        if (pr.kind == -2) {
            gpu_index.push_back(-2);
            trivial.emplace_back();
            return kPackOk;
        }
        const uint32_t cap = 2 * pr.blen + 64;
        /* the batch limits (memory budget, uint32 offsets) make a batch FULL; an empty batch takes any window so
         * that the caller's add-until-full loop always makes progress (cudabatch.cpp:126-132) */
        if (n_gpu() > 0 && (n_bases + pr.tot > max_bases || n_gpu() >= max_windows || out_total + cap > max_out))
            return kPackFull;
        if (n_bases + pr.tot > 0xfff00000ull || out_total + cap > 0xfff00000ull) return kPackFull;
        /* reserve everything first: a failed allocation must leave the batch as it was */
        const size_t ns = pr.order.size();
        if ((!by_ref && (!bases.reserve(bases.size + pr.tot) || !weights.reserve(weights.size + pr.tot))) ||
            (by_ref && (!src_pos.reserve(src_pos.size + ns) || !src_flags.reserve(src_flags.size + ns))) ||
            !seq_off.reserve(seq_off.size + ns) || !seq_begin.reserve(seq_begin.size + ns) ||
            !seq_end.reserve(seq_end.size + ns) || !seq_flags.reserve(seq_flags.size + ns) ||
            !win_first.reserve(win_first.size + 1) || !win_flags.reserve(win_flags.size + 1) ||
            !win_alpha.reserve(win_alpha.size + 1) || !out_off.reserve(out_off.size + 1) ||
            !out_cap.reserve(out_cap.size + 1))
            return kPackNoMem;
        pr.base_off = n_bases;
        n_bases += pr.tot;
        if (!by_ref) {
            bases.extend(pr.tot);
            weights.extend(pr.tot);
        }
        uint32_t off = seq_off.data[seq_off.size - 1];
        for (uint32_t i = 0; i < ns; ++i) {
            uint32_t k = pr.order[i];
            off += len[k];
            seq_off.push(off);
            seq_begin.push(i ? begin[k] : 0);
            seq_end.push(i ? end[k] : 0);
            seq_flags.push(pr.full[i]);
        }
        win_first.push(static_cast<uint32_t>(seq_off.size - 1));
        win_flags.push((window_type == 1 && trim) ? 1 : 0);
        win_alpha.push(pr.alpha);
        out_off.push(static_cast<uint32_t>(out_total));
        out_cap.push(cap);
        out_total += cap;
        gpu_index.push_back(static_cast<int32_t>(n_gpu() - 1));
        trivial.emplace_back();
        cost.push_back(static_cast<uint32_t>(pr.tot));
        return kPackOk;
    }

    /* thread-safe for distinct windows once every commit() of the batch is done (buffers no longer move) */
    void fill(const Prep& pr, const char* const* seq, const uint32_t* len, const char* const* qual) {
        if (pr.status != kPackOk || pr.kind != 0) return;
        uint8_t* b = bases.data + pr.base_off;
        uint8_t* w = weights.data + pr.base_off;
        for (uint32_t i = 0; i < pr.order.size(); ++i) {
            uint32_t k = pr.order[i];
            std::memcpy(b, seq[k], len[k]);
            const char* q = qual ? qual[k] : nullptr;
            if (q) {
                for (uint32_t j = 0; j < len[k]; ++j) w[j] = static_cast<uint8_t>(q[j] - 33);  // graph.cpp:141-143
            } else {
                std::memset(w, i == 0 ? 0 : 1, len[k]);  // dummy '!' backbone => 0; layer without quality => 1
            }
            b += len[k];
            w += len[k];
        }
    }

    /* One window; same argument meaning as rp_poa_add_window (include/racon_b200.h). */
    int add(uint32_t n_seq, const char* const* seq, const uint32_t* len, const char* const* qual,
            const uint32_t* begin, const uint32_t* end, int window_type, int trim) {
        Prep pr;
        prepare(pr, n_seq, seq, len, begin, end, max_seq_len);
        int r = commit(pr, seq, len, begin, end, window_type, trim);
        if (r == kPackOk) fill(pr, seq, len, qual);
        return r;
    }

    /* processing order: most expensive windows first (persistent warps pull from this queue) */
    bool build_queue() {
        uint32_t n = n_gpu();
        queue.clear();
        if (!queue.reserve(n ? n : 1)) return false;
        std::vector<uint32_t> idx(n);
        for (uint32_t i = 0; i < n; ++i) idx[i] = i;
        std::stable_sort(idx.begin(), idx.end(), [&](uint32_t a, uint32_t c) { return cost[a] > cost[c]; });
        for (uint32_t i = 0; i < n; ++i) queue.push(idx[i]);
        return true;
    }

    uint32_t max_layer_len() const {
        uint32_t m = 0;
        for (size_t s = 0; s + 1 < seq_off.size; ++s) m = std::max(m, seq_off.data[s + 1] - seq_off.data[s]);
        return m;
    }
};

}  // namespace rp
