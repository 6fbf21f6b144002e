/*
 * poa_core.cuh — B200-native per-window partial-order-alignment consensus (device code).
 *
 * What the reference does per window (Window::generate_consensus, /root/reference/src/window.cpp:65-149,
 * driving spoa: vendor/spoa/src/graph.cpp and sisd/simd alignment engines) is done here by ONE WARP per
 * window, many windows in flight per SM, in one persistent kernel launch:
 *
 *   for every layer:   [subgraph marking]  ->  DP "program" build  ->  NW-linear graph x read DP
 *                      ->  traceback  ->  graph merge  ->  incremental topological order update
 *   then:              spoa-order DFS sort -> heaviest-bundle consensus -> coverage -> TGS trim
 *
 * Design (not a port):
 *   * The POA graph lives in HBM as flat per-node slot arrays (in-edge slots with weights, aligned-node
 *     slots, per-node sequence counters instead of per-edge label lists).  Every graph phase is a
 *     lane-strided data-parallel loop, so HBM latency is overlapped 32-wide.
 *   * The DP keeps a whole 512-column row chunk in registers: lane l owns 16 consecutive columns as
 *     8 packed int16x2 registers; predecessor rows come from a shared-memory ring of the most recent
 *     rows (conflict-free swizzled 128-bit loads) or, for far predecessors, from the HBM copy; the cell
 *     update is two DPX VIADDMNMX.S16x2 per predecessor per register; the in-row gap recurrence is a
 *     lane-local max-plus chain plus a 5-step warp shuffle scan.
 *   * Rows are processed in an incrementally maintained topological order (aligned clusters kept
 *     contiguous).  Cell values and the traceback do not depend on which valid topological order is
 *     used; spoa's own DFS order (graph.cpp:249-303) is only needed to break ties between equally
 *     scoring sink rows and for the final consensus, and is computed on demand.
 *   * H rows are streamed once to HBM (coalesced 1 KB row stores) for the traceback.
 *
 * Exact semantics reproduced: A.1 DP / sink choice / traceback priority,
 * A.2 graph merge, A.3 topological order (on demand), A.4 heaviest bundle + branch completion,
 * A.5 coverage + trim, A.6 subgraph.
 */
#pragma once
#include "rp_warp.cuh"

namespace rp {

constexpr uint16_t kNone = 0xffffu;
constexpr int kChunkCols = 512;          // columns per register-resident row chunk (32 lanes x 16)
constexpr int32_t kNegDiag = -32640;     // "column -1" sentinel: + any int8 profile value stays >= INT16_MIN and
                                         // below every legal score (legal >= INT16_MIN + 1024, the int16 criterion)
constexpr int32_t kNeg32 = -(1 << 28);
constexpr int32_t kMaxGapInt16 = 64;     // |gap| bound of the packed-int16 path (sentinel + 16*gap must not wrap)

/* window status codes (soft, per window; mirrored in include/racon_b200.h) */
enum : uint32_t {
    kWinOk = 0,
    kWinNodeLimit = 1,
    kWinEdgeLimit = 2,
    kWinAlignedLimit = 3,
    kWinNeedsInt32 = 4,
    kWinSeqTooLong = 5,
    kWinStackLimit = 6,
    kWinAlphabetLimit = 7,
    kWinInternal = 8,
};

struct PoaLimits {
    uint32_t nmax;   // max graph nodes per window
    uint32_t lmax;   // max layer length
    uint32_t lp;     // padded row length in int16 cells: multiple of kChunkCols, >= lmax + 1
    uint32_t ki;     // in-edge slots per node
    uint32_t ka;     // aligned-node slots per node
    uint32_t stack_cap;
};

struct SlotLayout {
    uint64_t code, flags, in_cnt, al_cnt, cov, in_tail, in_w, al, order_a, order_b, rank_of, dp_order, dp_rank,
        rec, pred_ovf, H, aln, cur, wts_unused, member, has_out_sub, score, cpred, sorder, srank, marks, stack,
        newlist, ccarry, bytes;
};

RP_HD uint64_t align_up(uint64_t x, uint64_t a) { return (x + a - 1) / a * a; }

RP_HD SlotLayout make_layout(const PoaLimits& L) {
    SlotLayout s;
    uint64_t o = 0;
    auto take = [&](uint64_t bytes) {
        uint64_t r = o;
        o = align_up(o + bytes, 256);
        return r;
    };
    uint64_t n = L.nmax;
    s.code = take(n);
    s.flags = take(n);
    s.in_cnt = take(n);
    s.al_cnt = take(n);
    s.cov = take(n * 2);
    s.in_tail = take(n * L.ki * 2);
    s.in_w = take(n * L.ki * 4);
    s.al = take(n * L.ka * 2);
    s.order_a = take((n + 2 + 64) * 2);
    s.order_b = take((n + 2 + 64) * 2);
    s.rank_of = take(n * 2);
    s.dp_order = take((n + 2 + 64) * 2);
    s.dp_rank = take(n * 2);
    s.rec = take((n + 1 + 160) * 16);
    s.pred_ovf = take((n + 1) * L.ki * 2);
    s.H = take((n + 1) * static_cast<uint64_t>(L.lp) * 2);
    s.aln = take((static_cast<uint64_t>(L.lmax) + 1) * 2);
    s.cur = take((static_cast<uint64_t>(L.lmax) + 1) * 2);
    s.wts_unused = o;
    s.member = take(n);
    s.has_out_sub = take(n);
    s.score = take(n * 8);
    s.cpred = take(n * 2);
    s.sorder = take(n * 2);
    s.srank = take(n * 2);
    s.marks = take(n);
    s.stack = take(static_cast<uint64_t>(L.stack_cap) * 2);
    s.newlist = take((static_cast<uint64_t>(L.lmax) + 1) * 4);
    s.ccarry = take((n + 2) * 2 * 2);
    s.bytes = align_up(o, 4096);
    return s;
}

/* Launch-wide parameters (plain pointers into HBM). */
struct PoaParams {
    int32_t match, mismatch, gap;
    uint32_t n_windows;
    /* inputs, packed by the host in processing order (backbone first, layers sorted as window.cpp:85-86) */
    const uint8_t* bases;
    const uint8_t* weights;       // per base: quality - 33, or 1 (layers without quality), 0 (dummy backbone)
    const uint32_t* seq_off;      // n_seq + 1
    const uint32_t* seq_begin;
    const uint32_t* seq_end;
    const uint8_t* seq_flags;     // bit0: full-span layer (window.cpp:92-93 test evaluated on the host)
    const uint32_t* win_first;    // n_windows + 1
    const uint8_t* win_flags;     // bit0: trim (type == kTGS && trim)
    const uint64_t* win_alpha;    // up to 8 distinct characters of the window, low byte first
    const uint32_t* queue;        // processing order of windows (longest first)
    uint32_t* queue_head;         // atomic cursor
    /* outputs */
    uint8_t* cons;                // window w at cons + out_off[w], capacity out_cap[w]
    uint16_t* cons_cov;           // same offsets (per-base coverage of the returned consensus)
    const uint32_t* out_off;
    const uint32_t* out_cap;
    uint32_t* cons_len;
    uint32_t* status;
    uint64_t* stats;              // optional device counters (may be null): [0] alignments, [1] dp cells, ...
    /* scratch */
    uint8_t* scratch;
    PoaLimits lim;
    SlotLayout lay;
    uint32_t smem_per_warp;       // bytes of shared memory owned by each warp
    uint32_t tile_rows;           // traceback tile height in ranks (0 = default 96)
    uint32_t debug_flags;         // tests only: bit0 = use the HBM-resident variants of the order DFS / bundle
};

/* Row layout.  A lane owns 16 consecutive columns; inside that 32-byte block register r (0..7) packs
 * column r in its low half and column 8+r in its high half, so the in-row gap recurrence runs as two
 * packed 8-long chains and the diagonal operand of register r is simply register r-1 of the predecessor.
 * perm(): logical column -> element index in a row stored in that register order (HBM copy);
 * swz(): additionally XOR-swizzles 16-byte granules so the two LDS.128 of a lane are bank-conflict free. */
RP_DEV uint32_t perm(uint32_t c) { return (c & ~15u) | ((c & 7u) << 1) | ((c >> 3) & 1u); }
RP_DEV uint32_t swz_e(uint32_t e) {
    uint32_t q = e >> 3;
    return ((q ^ ((q >> 3) & 1u)) << 3) | (e & 7u);
}
RP_DEV uint32_t swz(uint32_t c) { return swz_e(perm(c)); }

struct Row8 {
    uint32_t r[8];
};

/* One DP program record per row: a = code index (8) | in-degree (7) | sink (1) | pred0 | pred1 | pred2,
 * b = pred3..pred6 (DP ranks, 0 = virtual root row).  In-degrees above 7 spill to pred_ovf. */
struct alignas(16) Rec {
    uint64_t a, b;
};
RP_DEV uint32_t rec_pred(const Rec& r, uint32_t k) {  // k < 7
    return k < 3 ? (static_cast<uint32_t>(r.a >> (16 + 16 * k)) & 0xffffu)
                 : (static_cast<uint32_t>(r.b >> (16 * (k - 3))) & 0xffffu);
}

struct alignas(16) U4 {
    uint32_t x, y, z, w;
};

RP_DEV Row8 load_row_smem(const int16_t* row, uint32_t chunk, int lane) {
    uint32_t q0 = chunk * 64u + 2u * static_cast<uint32_t>(lane);
    uint32_t p0 = q0 ^ ((q0 >> 3) & 1u);
    uint32_t p1 = (q0 + 1u) ^ (((q0 + 1u) >> 3) & 1u);
    const U4* b = reinterpret_cast<const U4*>(row);
    U4 a = b[p0], c = b[p1];
    Row8 o;
    o.r[0] = a.x; o.r[1] = a.y; o.r[2] = a.z; o.r[3] = a.w;
    o.r[4] = c.x; o.r[5] = c.y; o.r[6] = c.z; o.r[7] = c.w;
    return o;
}
RP_DEV void store_row_smem(int16_t* row, uint32_t chunk, int lane, const Row8& v) {
    uint32_t q0 = chunk * 64u + 2u * static_cast<uint32_t>(lane);
    uint32_t p0 = q0 ^ ((q0 >> 3) & 1u);
    uint32_t p1 = (q0 + 1u) ^ (((q0 + 1u) >> 3) & 1u);
    U4* b = reinterpret_cast<U4*>(row);
    b[p0] = U4{v.r[0], v.r[1], v.r[2], v.r[3]};
    b[p1] = U4{v.r[4], v.r[5], v.r[6], v.r[7]};
}
RP_DEV Row8 load_row_gmem(const int16_t* row, uint32_t chunk, int lane) {
    const U4* b = reinterpret_cast<const U4*>(row) + chunk * 64u + 2u * static_cast<uint32_t>(lane);
    U4 a = b[0], c = b[1];
    Row8 o;
    o.r[0] = a.x; o.r[1] = a.y; o.r[2] = a.z; o.r[3] = a.w;
    o.r[4] = c.x; o.r[5] = c.y; o.r[6] = c.z; o.r[7] = c.w;
    return o;
}
RP_DEV void store_row_gmem(int16_t* row, uint32_t chunk, int lane, const Row8& v) {
    U4* b = reinterpret_cast<U4*>(row) + chunk * 64u + 2u * static_cast<uint32_t>(lane);
    b[0] = U4{v.r[0], v.r[1], v.r[2], v.r[3]};
    b[1] = U4{v.r[4], v.r[5], v.r[6], v.r[7]};
}

/* One warp's view of its scratch slot + shared memory. All scalar members are warp-uniform. */
struct PoaWarp {
    const PoaParams* P;
    int lane;
    /* graph (HBM) */
    uint8_t *code, *flags, *in_cnt, *al_cnt, *member, *has_out_sub, *marks;
    uint16_t *cov, *in_tail, *al, *order, *order_nxt, *rank_of, *dp_order, *dp_rank, *pred_ovf, *aln, *cur, *cpred,
        *sorder, *srank, *stack;
    int32_t* in_w;
    Rec* rec;
    int16_t* H;
    int64_t* score;
    uint32_t* newlist;
    int16_t* ccarry;
    /* shared memory */
    int16_t* ring;       // ring_rows x lp_a
    int16_t* prof;       // ncodes x lp_a
    uint8_t* smem;
    uint32_t smem_bytes;
    /* state */
    uint32_t N;          // nodes
    uint32_t ki, ka, nmax;
    uint32_t status;
    uint64_t alpha;
    uint32_t ncodes;
    uint32_t n_added;    // sequences merged so far (incl. backbone)

    RP_DEV void bind(const PoaParams* p, uint8_t* slot, uint8_t* sm, uint32_t sm_bytes) {
        P = p;
        lane = lane_id();
        const SlotLayout& y = p->lay;
        code = slot + y.code;
        flags = slot + y.flags;
        in_cnt = slot + y.in_cnt;
        al_cnt = slot + y.al_cnt;
        cov = reinterpret_cast<uint16_t*>(slot + y.cov);
        in_tail = reinterpret_cast<uint16_t*>(slot + y.in_tail);
        in_w = reinterpret_cast<int32_t*>(slot + y.in_w);
        al = reinterpret_cast<uint16_t*>(slot + y.al);
        order = reinterpret_cast<uint16_t*>(slot + y.order_a);
        order_nxt = reinterpret_cast<uint16_t*>(slot + y.order_b);
        rank_of = reinterpret_cast<uint16_t*>(slot + y.rank_of);
        dp_order = reinterpret_cast<uint16_t*>(slot + y.dp_order);
        dp_rank = reinterpret_cast<uint16_t*>(slot + y.dp_rank);
        rec = reinterpret_cast<Rec*>(slot + y.rec);
        pred_ovf = reinterpret_cast<uint16_t*>(slot + y.pred_ovf);
        H = reinterpret_cast<int16_t*>(slot + y.H);
        aln = reinterpret_cast<uint16_t*>(slot + y.aln);
        cur = reinterpret_cast<uint16_t*>(slot + y.cur);
        member = slot + y.member;
        has_out_sub = slot + y.has_out_sub;
        score = reinterpret_cast<int64_t*>(slot + y.score);
        cpred = reinterpret_cast<uint16_t*>(slot + y.cpred);
        sorder = reinterpret_cast<uint16_t*>(slot + y.sorder);
        srank = reinterpret_cast<uint16_t*>(slot + y.srank);
        marks = slot + y.marks;
        stack = reinterpret_cast<uint16_t*>(slot + y.stack);
        newlist = reinterpret_cast<uint32_t*>(slot + y.newlist);
        ccarry = reinterpret_cast<int16_t*>(slot + y.ccarry);
        smem = sm;
        smem_bytes = sm_bytes;
        ki = p->lim.ki;
        ka = p->lim.ka;
        nmax = p->lim.nmax;
    }

    RP_DEV void fail(uint32_t st) {
        if (status == kWinOk) status = st;
    }

    RP_DEV uint32_t code_index(uint8_t c) const {  // position of character c in the window alphabet
        uint32_t k = 0;
        for (; k < ncodes; ++k)
            if (static_cast<uint8_t>(alpha >> (8 * k)) == c) break;
        return k;
    }

    /* ---------------------------------------------------------------- backbone (graph.cpp:186-190, 93-110) */
    RP_DEV void init_backbone(const uint8_t* seq, const uint8_t* w, uint32_t len) {
        for (uint32_t v = lane; v < len; v += 32) {
            code[v] = seq[v];
            al_cnt[v] = 0;
            flags[v] = (v + 1 < len) ? 1 : 0;
            cov[v] = len >= 2 ? 1 : 0;
            if (v > 0) {
                in_cnt[v] = 1;
                in_tail[v * ki] = static_cast<uint16_t>(v - 1);
                in_w[v * ki] = static_cast<int32_t>(w[v - 1]) + static_cast<int32_t>(w[v]);
            } else {
                in_cnt[v] = 0;
            }
            order[v + 1] = static_cast<uint16_t>(v);
            rank_of[v] = static_cast<uint16_t>(v + 1);
        }
        if (lane == 0) order[0] = kNone;
        N = len;
        n_added = 1;
        syncwarp();
    }

    /* ---------------------------------------------------------------- subgraph (graph.cpp:518-539) */
    /* Backward DFS from backbone node `end` over in-edges and aligned nodes, keeping ids >= begin.
     * Serial (lane 0); only partial-span layers take this path. */
    RP_DEV void mark_subgraph(uint32_t begin, uint32_t end) {
        const uint32_t npad = (N + 15) & ~15u;
        const bool staged = ki <= 31 && ka == 8 && 2 * npad + 512 <= smem_bytes && !(P->debug_flags & 1);
        if (!staged) {  /* HBM-resident variant */
            for (uint32_t v = lane; v < N; v += 32) member[v] = 0;
            syncwarp();
            if (lane == 0) {
                uint32_t sp = 0;
                const uint32_t cap = P->lim.stack_cap;
                stack[sp++] = static_cast<uint16_t>(end);
                while (sp > 0) {
                    uint32_t c = stack[--sp];
                    if (member[c] || c < begin) continue;
                    uint32_t ni = in_cnt[c], na = al_cnt[c];
                    if (sp + ni + na > cap) {
                        fail(kWinStackLimit);
                        break;
                    }
                    for (uint32_t k = 0; k < ni; ++k) stack[sp++] = in_tail[c * ki + k];
                    for (uint32_t k = 0; k < na; ++k) stack[sp++] = al[c * ka + k];
                    member[c] = 1;
                }
            }
            status = shfl(status, 0);
            syncwarp();
            return;
        }
        /* membership bytes, per-node degrees, the first two in-edge tails of every node and the stack live in
         * shared memory: a node costs an HBM round trip only if it has more than two in-edges or aligned nodes */
        uint8_t* smb = smem;
        uint8_t* sme = smem + npad;
        const bool have_t2 = 6 * npad + 1024 <= smem_bytes;
        uint32_t* st2 = reinterpret_cast<uint32_t*>(smem + 2 * npad);
        const uint32_t fixed = have_t2 ? 6 * npad : 2 * npad;
        uint16_t* sst = reinterpret_cast<uint16_t*>(smem + fixed);
        const uint32_t cap = (smem_bytes - fixed) / 2;
        for (uint32_t v = lane; v < N; v += 32) {
            smb[v] = 0;
            sme[v] = static_cast<uint8_t>(in_cnt[v] | (al_cnt[v] << 5));
            if (have_t2) st2[v] = *reinterpret_cast<const uint32_t*>(in_tail + v * ki);
        }
        syncwarp();
        if (lane == 0) {
            uint32_t sp = 0;
            sst[sp++] = static_cast<uint16_t>(end);
            while (sp > 0) {
                const uint32_t c = sst[--sp];
                if (smb[c] || c < begin) continue;
                const uint32_t ni = sme[c] & 31u, na = sme[c] >> 5;
                uint64_t t4 = have_t2 ? st2[c] : 0;
                if (!have_t2 || ni > 2) t4 = *reinterpret_cast<const uint64_t*>(in_tail + c * ki);
                uint64_t a_lo = 0, a_hi = 0;
                if (na) {
                    const U4 a8 = *reinterpret_cast<const U4*>(al + c * ka);
                    a_lo = a8.x | (static_cast<uint64_t>(a8.y) << 32);
                    a_hi = a8.z | (static_cast<uint64_t>(a8.w) << 32);
                }
                if (sp + ni + na > cap) {
                    fail(kWinStackLimit);
                    break;
                }
                for (uint32_t k = 0; k < ni; ++k) {
                    const uint32_t t = k < 4 ? (static_cast<uint32_t>(t4 >> (16 * k)) & 0xffffu) : in_tail[c * ki + k];
                    if (!smb[t] && t >= begin) sst[sp++] = static_cast<uint16_t>(t);
                }
                for (uint32_t k = 0; k < na; ++k) {
                    const uint32_t a = static_cast<uint32_t>((k < 4 ? a_lo : a_hi) >> (16 * (k & 3))) & 0xffffu;
                    if (!smb[a] && a >= begin) sst[sp++] = static_cast<uint16_t>(a);
                }
                smb[c] = 1;
            }
        }
        status = shfl(status, 0);
        syncwarp();
        for (uint32_t v = lane; v < N; v += 32) member[v] = smb[v];
        syncwarp();
    }

    /* Compact DP order over the member nodes (rank order preserved). Returns number of DP rows. */
    RP_DEV uint32_t build_dp_order_subgraph() {
        uint32_t base = 0;
        for (uint32_t r0 = 1; r0 <= N; r0 += 32) {
            uint32_t r = r0 + lane;
            uint32_t v = r <= N ? order[r] : 0;
            bool in = r <= N && member[v];
            uint32_t tot;
            uint32_t pos = warp_rank(in, &tot);
            if (in) {
                dp_order[base + pos + 1] = static_cast<uint16_t>(v);
                dp_rank[v] = static_cast<uint16_t>(base + pos + 1);
            }
            base += tot;
        }
        for (uint32_t v = lane; v < N; v += 32) has_out_sub[v] = 0;
        syncwarp();
        /* sinks of the subgraph: members without an out-edge to another member (graph.cpp:575-580) */
        for (uint32_t v = lane; v < N; v += 32) {
            if (!member[v]) continue;
            uint32_t ni = in_cnt[v];
            for (uint32_t k = 0; k < ni; ++k) {
                uint32_t t = in_tail[v * ki + k];
                if (member[t]) has_out_sub[t] = 1;
            }
        }
        syncwarp();
        return base;
    }

    /* ---------------------------------------------------------------- DP program (one 8-byte record per row)
     * byte0 code index, byte1 = npred(7 bits) | sink<<7, bytes2..7 = first three predecessor DP ranks
     * (0 = virtual root row).  Further predecessors spill to pred_ovf. */
    RP_DEV uint32_t build_program(uint32_t nrows, bool sub) {
        const uint16_t* ord = sub ? dp_order : order;
        const uint16_t* rk = sub ? dp_rank : rank_of;
        uint32_t pred_rows = 0;
        /* The graph lives in HBM, so every level of the chain row -> node -> in-edges -> ranks costs a full
         * memory latency.  Rows are handled kU per lane at a time with all loads of one level issued together. */
        constexpr int kU = 4;
        for (uint32_t r0 = 1; r0 <= nrows; r0 += 32 * kU) {
            uint32_t r[kU], v[kU], ni[kU], cd[kU], fl[kU];
            uint64_t t4[kU];
            uint32_t pk[kU][3];
            bool use[kU][3];
#pragma unroll
            for (int u = 0; u < kU; ++u) {
                r[u] = r0 + u * 32 + lane;
                v[u] = r[u] <= nrows ? ord[r[u]] : ord[1];
            }
#pragma unroll
            for (int u = 0; u < kU; ++u) {
                ni[u] = in_cnt[v[u]];
                cd[u] = code[v[u]];
                fl[u] = sub ? has_out_sub[v[u]] : (flags[v[u]] & 1u);
                t4[u] = *reinterpret_cast<const uint64_t*>(in_tail + v[u] * ki);  // first four in-edge tails
            }
#pragma unroll
            for (int u = 0; u < kU; ++u)
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    uint32_t t = static_cast<uint32_t>(t4[u] >> (16 * k)) & 0xffffu;
                    use[u][k] = static_cast<uint32_t>(k) < ni[u] && (!sub || member[t]);
                    pk[u][k] = t;
                }
#pragma unroll
            for (int u = 0; u < kU; ++u)
#pragma unroll
                for (int k = 0; k < 3; ++k) pk[u][k] = use[u][k] ? rk[pk[u][k]] : 0u;
#pragma unroll
            for (int u = 0; u < kU; ++u) {
                if (r[u] > nrows) continue;
                uint32_t np = 0;
                uint64_t pa = 0, pb = 0;
                auto put = [&](uint32_t pr) {
                    if (np < 3)
                        pa |= static_cast<uint64_t>(pr) << (16 + 16 * np);
                    else if (np < 7)
                        pb |= static_cast<uint64_t>(pr) << (16 * (np - 3));
                    else
                        pred_ovf[r[u] * ki + np] = static_cast<uint16_t>(pr);
                    ++np;
                };
#pragma unroll
                for (int k = 0; k < 3; ++k)
                    if (use[u][k]) put(pk[u][k]);
                for (uint32_t k = 3; k < ni[u]; ++k) {  // less common: in-degree > 3
                    uint32_t t = in_tail[v[u] * ki + k];
                    if (sub && !member[t]) continue;
                    put(rk[t]);
                }
                bool sink = !fl[u];
                Rec rc;
                rc.a = static_cast<uint64_t>(code_index(static_cast<uint8_t>(cd[u]))) |
                       (static_cast<uint64_t>(np & 0x7f) << 8) | (static_cast<uint64_t>(sink ? 1 : 0) << 15) | pa;
                rc.b = pb;
                rec[r[u]] = rc;
                pred_rows += np ? np : 1;
            }
        }
        if (lane == 0) rec[0] = Rec{0, 0};
        syncwarp();
        return pred_rows;  // per-lane partial sum (only summed when the device counters are on)
    }

    /* ---------------------------------------------------------------- the DP (sisd :292-360, simd :760-906)
     * Column c of a row holds H[row][c] for c = 0..len (column 0 is spoa's first_column); column -1 is
     * -infinity, which makes column 0 follow the general recurrence.  Returns the best sink row
     * (first strictly greater in processing order) in *best_row, its score, and the number of sink rows
     * that reach that score. */
    /* one 512-column chunk of the match/mismatch profile (sisd :123-131) into shared memory */
    RP_DEV void build_profile(const uint8_t* seq, uint32_t len, uint32_t ch) {
        int32_t m = P->match, x = P->mismatch;
        for (uint32_t k = 0; k < ncodes; ++k) {
            uint8_t c = static_cast<uint8_t>(alpha >> (8 * k));
            int16_t* row = prof + k * kChunkCols;
            for (uint32_t cc = lane; cc < kChunkCols; cc += 32) {
                uint32_t col = ch * kChunkCols + cc;
                int16_t v = static_cast<int16_t>(x);
                if (col >= 1 && col <= len && seq[col - 1] == c) v = static_cast<int16_t>(m);
                row[swz(cc)] = v;
            }
        }
    }

    /* True iff no cell of the finished matrix can have left the range int16 arithmetic is exact in.  Every cell
     * satisfies H[i][j] >= H[i][0] + j*g (the horizontal move is always a candidate), column 0 falls by exactly |g|
     * per level along a shortest path (so it cannot jump over the limit unnoticed), and the largest value is bounded
     * by match * min(rows, columns), checked by the caller.  Limit: spoa's own margin, -32768 + 1024. */
    RP_DEV bool matrix_in_range(uint32_t nrows, uint32_t len, uint32_t lpa) {
        int32_t mn = 0;
        const uint32_t e0 = perm(0);
        for (uint32_t i = 1 + lane; i <= nrows; i += 32) {
            const int32_t v = H[static_cast<uint64_t>(i) * lpa + e0];
            mn = v < mn ? v : mn;
        }
        mn = -warp_incl_max(-mn);
        mn = shfl(mn, 31);
        return mn + static_cast<int32_t>(len) * P->gap >= -32768 + 1024;
    }

    RP_DEV void dp(const uint8_t* seq, uint32_t nrows, uint32_t len, uint32_t lpa, uint32_t ring_rows,
                   uint32_t* best_row, int32_t* best_score, uint32_t* n_best) {
        /* Row-synchronous: the whole warp computes one row (512-column chunk) at a time; lane l owns columns
         * 16l..16l+15.  Rows longer than 512 columns are done chunk by chunk (one pass over all rows per chunk,
         * the carry between chunks goes through a small per-row array), so shared memory only ever holds one
         * chunk: the profile chunk and a ring of the last `ring_rows` rows. */
        const int32_t g = P->gap;
        const uint32_t g2 = pack16(g, g);
        const uint32_t nch = lpa / kChunkCols;
        const int32_t negsafe = -32768 - 16 * g;  // see kMaxGapInt16
        uint32_t gb[8], gc[8];  // bridge / carry offsets per register
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            gb[r] = pack16(0, (r + 1) * g);
            gc[r] = pack16((r + 1) * g, (r + 9) * g);
        }
        for (uint32_t col = lane; col < lpa; col += 32) H[perm(col)] = static_cast<int16_t>(static_cast<int32_t>(col) * g);
        const uint32_t sink_ch = len / kChunkCols, sink_e = swz(len % kChunkCols);
        int32_t best = kNeg32;
        uint32_t bi = 0, nb = 0;
        int16_t* cc_prev = ccarry;        // last column of the previous chunk, per row (multi-chunk rows only)
        int16_t* cc_cur = ccarry + (nmax + 2);
        const uint32_t lanem = static_cast<uint32_t>(lane);
        for (uint32_t ch = 0; ch < nch; ++ch) {
            build_profile(seq, len, ch);
            for (uint32_t c = lane; c < kChunkCols; c += 32)  // root row (rank 0) -> ring slot 0
                ring[swz(c)] = static_cast<int16_t>(static_cast<int32_t>(ch * kChunkCols + c) * g);
            syncwarp();
            const bool multi = ch > 0;
            uint32_t rec_a_lo = 0, rec_a_hi = 0, rec_b_lo = 0, rec_b_hi = 0;
            int16_t* hrow = H + static_cast<uint64_t>(ch) * kChunkCols;  // row i of this chunk = hrow + i*lpa
            uint32_t myslot = 0;  // i % ring_rows, kept incrementally (any ring size, no division)
            for (uint32_t i = 1; i <= nrows; ++i) {
                myslot = myslot + 1 == ring_rows ? 0u : myslot + 1;
                const uint32_t ti = (i - 1) & 31u;
                if (ti == 0) {
                    Rec t = rec[i + lane];  // rec[] is padded
                    rec_a_lo = static_cast<uint32_t>(t.a);
                    rec_a_hi = static_cast<uint32_t>(t.a >> 32);
                    rec_b_lo = static_cast<uint32_t>(t.b);
                    rec_b_hi = static_cast<uint32_t>(t.b >> 32);
                }
                const uint32_t lo = shfl(rec_a_lo, ti);
                const uint32_t cidx = lo & 0xff;
                const uint32_t np = (lo >> 8) & 0x7f;
                const bool sink = (lo >> 15) & 1;
                const Row8 pf = load_row_smem(prof + cidx * kChunkCols, 0, lane);
                /* max over predecessors distributes over both terms of the recurrence:
                 *   max_p(H[p][c-1] + s(c), H[p][c] + g) = max(max_p H[p][c-1] + s(c), max_p H[p][c] + g),
                 * so predecessor rows are first combined with a packed max (8 ops per extra predecessor) and the
                 * diagonal/vertical update (16 DPX ops, one shuffle) runs once per row */
                Row8 pm;
                int32_t lvm = kNegDiag;
                auto load_pred = [&](uint32_t p, Row8& pr) {
                    const uint32_t dist = i - p;
                    if (dist < ring_rows) {  // warp-uniform
                        const uint32_t slot = myslot >= dist ? myslot - dist : myslot + ring_rows - dist;
                        pr = load_row_smem(ring + slot * kChunkCols, 0, lane);
                    } else
                        pr = load_row_gmem(hrow + static_cast<uint64_t>(p) * lpa, 0, lane);
                    if (multi) {
                        int32_t lv = cc_prev[p];
                        lvm = lv > lvm ? lv : lvm;
                    }
                };
                auto more = [&](uint32_t p) {
                    Row8 pr;
                    load_pred(p, pr);
#pragma unroll
                    for (int r = 0; r < 8; ++r) pm.r[r] = vmax_s16x2(pm.r[r], pr.r[r]);
                };
                if (np <= 1) {
                    load_pred(np ? (lo >> 16) : 0u, pm);
                } else {
                    const uint32_t hi = shfl(rec_a_hi, ti);
                    load_pred(lo >> 16, pm);
                    more(hi & 0xffff);
                    if (np > 2) more(hi >> 16);
                    if (np > 3) {
                        const uint32_t blo = shfl(rec_b_lo, ti), bhi = shfl(rec_b_hi, ti);
                        more(blo & 0xffff);
                        if (np > 4) more(blo >> 16);
                        if (np > 5) more(bhi & 0xffff);
                        if (np > 6) more(bhi >> 16);
                        for (uint32_t k = 7; k < np; ++k) more(pred_ovf[i * ki + k]);
                    }
                }
                uint32_t acc[8];
                {
                    uint32_t left = shfl_up(pm.r[7], 1);  // hi half = previous lane's last column
                    left = lanem == 0 ? (static_cast<uint32_t>(lvm) << 16) : left;
                    /* diagonal operand of register 0 = (column -1 of the block, column 7) */
                    const uint32_t d0 = byte_perm(left, pm.r[7], 0x5432);
                    acc[0] = viaddmax_s16x2(pm.r[0], g2, viaddmax_s16x2(d0, pf.r[0], 0x80008000u));
#pragma unroll
                    for (int r = 1; r < 8; ++r)
                        acc[r] = viaddmax_s16x2(pm.r[r], g2, viaddmax_s16x2(pm.r[r - 1], pf.r[r], 0x80008000u));
                }
                /* in-row gap recurrence H[c] = max(H[c], H[c-1] + g), all in packed int16:
                 * two 8-long chains (low halves = columns 0..7, high halves = columns 8..15 of the block) */
#pragma unroll
                for (int r = 1; r < 8; ++r) acc[r] = viaddmax_s16x2(acc[r - 1], g2, acc[r]);
                /* bridge: column 7 feeds columns 8..15 */
                const uint32_t bridge = byte_perm(acc[7], 0x80008000u, 0x1076);  // lo = INT16_MIN, hi = acc[7].lo
#pragma unroll
                for (int r = 0; r < 8; ++r) acc[r] = viaddmax_s16x2(bridge, gb[r], acc[r]);
                /* warp max-plus scan of the lane totals (column 15 of each block), decay 16*g per lane */
                int32_t chunk_carry = kNeg32;
                if (multi) chunk_carry = cc_prev[i];
                int32_t t = hi16(acc[7]);
                t = viaddmax_s32(lanem == 0 ? chunk_carry : kNeg32, 16 * g, t);
#pragma unroll
                for (int dd = 1; dd < 32; dd <<= 1) {
                    int32_t o = shfl_up(t, dd);
                    t = viaddmax_s32(lanem >= static_cast<uint32_t>(dd) ? o : kNeg32, dd * 16 * g, t);
                }
                int32_t carry = shfl_up(t, 1);
                carry = lanem == 0 ? chunk_carry : carry;
                carry = carry < negsafe ? negsafe : carry;
                const uint32_t c2 = pack16(carry, carry);
#pragma unroll
                for (int r = 0; r < 8; ++r) acc[r] = viaddmax_s16x2(c2, gc[r], acc[r]);
                Row8 out;
#pragma unroll
                for (int r = 0; r < 8; ++r) out.r[r] = acc[r];
                int16_t* myrow_s = ring + myslot * kChunkCols;
                store_row_smem(myrow_s, 0, lane, out);
                store_row_gmem(hrow + static_cast<uint64_t>(i) * lpa, 0, lane, out);
                if (nch > 1 && lanem == 31) cc_cur[i] = static_cast<int16_t>(hi16(acc[7]));
                syncwarp();
                if (sink && ch == sink_ch) {  // warp-uniform
                    int32_t sc = myrow_s[sink_e];
                    if (sc > best) {
                        best = sc;
                        bi = i;
                        nb = 1;
                    } else if (sc == best) {
                        ++nb;
                    }
                }
            }
            if (nch > 1) {
                if (lane == 0) cc_cur[0] = static_cast<int16_t>((ch + 1) * kChunkCols * g - g);  // root row, last column
                int16_t* tmp = cc_prev;
                cc_prev = cc_cur;
                cc_cur = tmp;
                syncwarp();
            }
        }
        *best_row = bi;
        *best_score = best;
        *n_best = nb;
        syncwarp();
    }

    /* ---------------------------------------------------------------- spoa's DFS order (graph.cpp:249-303)
     * Serial (lane 0).  With `sub`, runs on the subgraph exactly as spoa would on Graph::Subgraph():
     * nodes in ascending id, in-edges/aligned nodes filtered to members, original list orders kept. */
    RP_DEV void spoa_sort_hbm(bool sub) {
        for (uint32_t v = lane; v < N; v += 32) marks[v] = 0;  // bits0-1 mark, bit2 ignored
        syncwarp();
        if (lane == 0) {
            const uint32_t cap = P->lim.stack_cap;
            uint32_t out = 0, sp = 0;
            for (uint32_t root = 0; root < N && status == kWinOk; ++root) {
                if (sub && !member[root]) continue;
                if ((marks[root] & 3) != 0) continue;
                stack[sp++] = static_cast<uint16_t>(root);
                while (sp > 0) {
                    uint32_t c = stack[sp - 1];
                    bool valid = true;
                    uint8_t mk = marks[c];
                    if ((mk & 3) != 2) {
                        uint32_t ni = in_cnt[c], na = al_cnt[c];
                        if (sp + ni + na > cap) {
                            fail(kWinStackLimit);
                            break;
                        }
                        for (uint32_t k = 0; k < ni; ++k) {
                            uint32_t t = in_tail[c * ki + k];
                            if (sub && !member[t]) continue;
                            if ((marks[t] & 3) != 2) {
                                stack[sp++] = static_cast<uint16_t>(t);
                                valid = false;
                            }
                        }
                        if (!(mk & 4)) {
                            for (uint32_t k = 0; k < na; ++k) {
                                uint32_t a = al[c * ka + k];
                                if (sub && !member[a]) continue;
                                if ((marks[a] & 3) != 2) {
                                    stack[sp++] = static_cast<uint16_t>(a);
                                    marks[a] |= 4;
                                    valid = false;
                                }
                            }
                        }
                        if (valid) {
                            marks[c] = static_cast<uint8_t>((marks[c] & 4) | 2);
                            if (!(marks[c] & 4)) {
                                srank[c] = static_cast<uint16_t>(out);
                                sorder[out++] = static_cast<uint16_t>(c);
                                for (uint32_t k = 0; k < na; ++k) {
                                    uint32_t a = al[c * ka + k];
                                    if (sub && !member[a]) continue;
                                    srank[a] = static_cast<uint16_t>(out);
                                    sorder[out++] = static_cast<uint16_t>(a);
                                }
                            }
                        } else {
                            marks[c] = static_cast<uint8_t>((marks[c] & 4) | 1);
                        }
                    }
                    if (valid) --sp;
                }
            }
        }
        status = shfl(status, 0);
        syncwarp();
    }

    /* Same DFS with the per-node state the walk touches on every step (marks, in-degree, aligned count) and
     * the stack staged in shared memory (idle outside the DP): one HBM round trip per examined node (its
     * in-edge tails and aligned list are fetched together) instead of five dependent ones. */
    RP_DEV void spoa_sort(bool sub) {
        const uint32_t npad = (N + 15) & ~15u;
        if (ki > 31 || ka != 8 || 2 * npad + 512 > smem_bytes || (P->debug_flags & 1)) {
            spoa_sort_hbm(sub);
            return;
        }
        uint8_t* smk = smem;          // bits0-1 mark, bit2 ignored, bit3 member
        uint8_t* sme = smem + npad;   // in_cnt | al_cnt << 5
        /* the first two in-edge tails of every node too, when they fit: a node then costs an HBM round trip
         * only if it has more than two in-edges or aligned nodes */
        const bool have_t2 = 6 * npad + 1536 <= smem_bytes;
        uint32_t* st2 = reinterpret_cast<uint32_t*>(smem + 2 * npad);
        const uint32_t fixed = have_t2 ? 6 * npad : 2 * npad;
        uint16_t* sst = reinterpret_cast<uint16_t*>(smem + fixed);
        const uint32_t cap = (smem_bytes - fixed) / 2;
        for (uint32_t v = lane; v < N; v += 32) {
            smk[v] = (!sub || member[v]) ? 8 : 0;
            sme[v] = static_cast<uint8_t>(in_cnt[v] | (al_cnt[v] << 5));
            if (have_t2) st2[v] = *reinterpret_cast<const uint32_t*>(in_tail + v * ki);
        }
        syncwarp();
        if (lane == 0) {
            uint32_t out = 0, sp = 0;
            for (uint32_t root = 0; root < N && status == kWinOk; ++root) {
                if (smk[root] != 8) continue;  // not a member, or already marked
                sst[sp++] = static_cast<uint16_t>(root);
                while (sp > 0) {
                    const uint32_t c = sst[sp - 1];
                    bool valid = true;
                    const uint8_t mk = smk[c];
                    if ((mk & 3) != 2) {
                        const uint32_t ni = sme[c] & 31u, na = sme[c] >> 5;
                        uint64_t t4 = have_t2 ? st2[c] : 0;
                        if (!have_t2 || ni > 2) t4 = *reinterpret_cast<const uint64_t*>(in_tail + c * ki);
                        uint64_t a_lo = 0, a_hi = 0;
                        if (na) {
                            const U4 a8 = *reinterpret_cast<const U4*>(al + c * ka);
                            a_lo = a8.x | (static_cast<uint64_t>(a8.y) << 32);
                            a_hi = a8.z | (static_cast<uint64_t>(a8.w) << 32);
                        }
                        if (sp + ni + na > cap) {
                            fail(kWinStackLimit);
                            break;
                        }
                        for (uint32_t k = 0; k < ni; ++k) {
                            const uint32_t t = k < 4 ? (static_cast<uint32_t>(t4 >> (16 * k)) & 0xffffu)
                                                     : in_tail[c * ki + k];
                            const uint8_t mt = smk[t];
                            if (!(mt & 8)) continue;
                            if ((mt & 3) != 2) {
                                sst[sp++] = static_cast<uint16_t>(t);
                                valid = false;
                            }
                        }
                        if (!(mk & 4)) {
                            for (uint32_t k = 0; k < na; ++k) {
                                const uint32_t a = static_cast<uint32_t>((k < 4 ? a_lo : a_hi) >> (16 * (k & 3))) & 0xffffu;
                                const uint8_t ma = smk[a];
                                if (!(ma & 8)) continue;
                                if ((ma & 3) != 2) {
                                    sst[sp++] = static_cast<uint16_t>(a);
                                    smk[a] = ma | 4;
                                    valid = false;
                                }
                            }
                        }
                        if (valid) {
                            const uint8_t now = smk[c];
                            smk[c] = static_cast<uint8_t>((now & 12) | 2);
                            if (!(now & 4)) {
                                srank[c] = static_cast<uint16_t>(out);
                                sorder[out++] = static_cast<uint16_t>(c);
                                for (uint32_t k = 0; k < na; ++k) {
                                    const uint32_t a = static_cast<uint32_t>((k < 4 ? a_lo : a_hi) >> (16 * (k & 3))) & 0xffffu;
                                    if (!(smk[a] & 8)) continue;
                                    srank[a] = static_cast<uint16_t>(out);
                                    sorder[out++] = static_cast<uint16_t>(a);
                                }
                            }
                        } else {
                            smk[c] = static_cast<uint8_t>((smk[c] & 12) | 1);
                        }
                    }
                    if (valid) --sp;
                }
            }
        }
        status = shfl(status, 0);
        syncwarp();
    }

    /* Among sink rows whose last-column score equals `best`, the one spoa visits first (sisd :353-355). */
    RP_DEV uint32_t resolve_sink_tie(uint32_t nrows, uint32_t len, uint32_t lpa, int32_t best, bool sub) {
        spoa_sort(sub);
        const uint16_t* ord = sub ? dp_order : order;
        uint32_t bestkey = 0xffffffffu;
        for (uint32_t r = 1 + lane; r <= nrows; r += 32) {
            uint64_t rc = rec[r].a;
            if (!((rc >> 15) & 1)) continue;
            if (H[static_cast<uint64_t>(r) * lpa + perm(len)] != best) continue;
            uint32_t key = (static_cast<uint32_t>(srank[ord[r]]) << 16) | r;
            if (key < bestkey) bestkey = key;
        }
        for (int d = 16; d > 0; d >>= 1) {
            uint32_t o = shfl_down(bestkey, d);
            if (o < bestkey) bestkey = o;
        }
        bestkey = shfl(bestkey, 0);
        return bestkey & 0xffffu;
    }

    /* ---------------------------------------------------------------- traceback (sisd :366-459)
     * Priority: diagonal over predecessors in in-edge order, then vertical in the same order, then
     * horizontal.  Lanes test predecessors in parallel; the lowest lane that matches wins.
     * The walk is a chain of dependent reads of H, so it runs out of a shared-memory TILE: the warp
     * copies kTileRows ranks x 32 columns of H (plus those rows' program records) from HBM with one
     * coalesced burst, walks until the path leaves the tile (about 30-40 steps), and re-anchors.
     * Predecessors below the tile (rare long edges) are read from HBM directly.
     * Output: aln[j] = node aligned to read position j, or kNone (new node). */
    /* > 32 predecessors (escalated windows only): every diagonal candidate of every group outranks any
     * vertical one.  Reads straight from the HBM copy; kept out of line so the common loop stays small. */
    static RP_DEV_NOINLINE int traceback_step_wide(const int16_t* H, const Rec* rec, const uint16_t* pred_ovf, uint32_t ki,
                                                   int lane, int32_t g, uint32_t i, uint32_t j, uint32_t npe,
                                                   int32_t hij, int32_t mc, uint32_t lpa, uint32_t* found) {
        for (int pass = 1; pass <= 2; ++pass) {
            if (pass == 1 && j == 0) continue;
            for (uint32_t k0 = 0; k0 < npe; k0 += 32) {
                const uint32_t k = k0 + lane;
                uint32_t p = 0;
                bool ok = false;
                if (k < npe) {
                    p = k < 7 ? rec_pred(rec[i], k) : pred_ovf[i * ki + k];
                    const int16_t* pr = H + static_cast<uint64_t>(p) * lpa;
                    ok = pass == 1 ? (hij == pr[perm(j - 1)] + mc) : (hij == pr[perm(j)] + g);
                }
                const uint32_t msk = ballot(ok);
                if (msk) {
                    *found = shfl(p, ffs_(msk) - 1);
                    return pass;
                }
            }
        }
        return 0;
    }

    static constexpr uint32_t kTileCols = 32;

    RP_DEV void traceback(uint32_t best_row, uint32_t len, uint32_t lpa, const uint8_t* seq, bool sub) {
        const int32_t g = P->gap, m = P->match, x = P->mismatch;
        const uint16_t* ord = sub ? dp_order : order;
        const uint32_t kTileRows = P->tile_rows ? P->tile_rows : 96;
        int16_t* tile = reinterpret_cast<int16_t*>(smem);                                        // [kTileRows][32]
        Rec* trec = reinterpret_cast<Rec*>(smem + kTileRows * kTileCols * 2);                      // [kTileRows]
        uint16_t* tnode = reinterpret_cast<uint16_t*>(smem + kTileRows * kTileCols * 2 + kTileRows * 16);
        uint8_t* tseq = smem + kTileRows * kTileCols * 2 + kTileRows * 16 + ((kTileRows * 2 + 15) & ~15u);
        for (uint32_t c = lane; c < len; c += 32) tseq[c] = seq[c];  // the walk reads seq[j-1] every step
        uint32_t i = best_row, j = len;
        uint32_t t_top = 0, t_rows = 0, t_col0 = 0;  // tile covers ranks (t_top - t_rows, t_top], cols [t_col0, t_col0+32)
        bool have_tile = false;
        while (i != 0) {
            if (!have_tile || i + t_rows <= t_top || (j > 0 && j - 1 < t_col0)) {
                syncwarp();
                t_top = i;
                t_rows = i + 1 < kTileRows ? i + 1 : kTileRows;  // down to rank 0 at most
                uint32_t cb = j >> 4;
                t_col0 = (cb ? cb - 1 : 0) << 4;
                for (uint32_t q = lane; q < t_rows; q += 32) {
                    uint32_t rk = t_top - q;
                    const U4* src = reinterpret_cast<const U4*>(H + static_cast<uint64_t>(rk) * lpa + t_col0);
                    U4* dst = reinterpret_cast<U4*>(tile + q * kTileCols);
                    U4 a = src[0], b = src[1], c = src[2], d = src[3];
                    dst[0] = a; dst[1] = b; dst[2] = c; dst[3] = d;
                    trec[q] = rec[rk];
                    tnode[q] = ord[rk];
                }
                have_tile = true;
                syncwarp();
            }
            const uint32_t q = t_top - i;
            const Rec rc = trec[q];
            const uint32_t lo = static_cast<uint32_t>(rc.a);
            const uint32_t node = tnode[q];
            const uint32_t cidx = lo & 0xff, np = (lo >> 8) & 0x7f;
            const uint32_t npe = np ? np : 1;
            const uint32_t ej = perm(j) - t_col0;                      // element of column j inside a tile row
            const uint32_t ejm = j > 0 ? perm(j - 1) - t_col0 : 0;     // column j-1
            const int32_t hij = tile[q * kTileCols + ej];
            int32_t mc = 0;
            if (j > 0) mc = (static_cast<uint8_t>(alpha >> (8 * cidx)) == tseq[j - 1]) ? m : x;
            uint32_t found_p = 0;
            int move = 0;  // 1 diag, 2 vert, 3 horiz
            if (npe <= 32) {
                /* one predecessor per lane; all diagonal candidates outrank any vertical one (sisd :392-442) */
                uint32_t p = 0;
                bool okd = false, okv = false;
                if (static_cast<uint32_t>(lane) < npe) {
                    if (np) p = static_cast<uint32_t>(lane) < 7 ? rec_pred(rc, lane) : pred_ovf[i * ki + lane];
                    int32_t a = 0, b;
                    if (p + t_rows > t_top) {  // predecessor row is inside the tile
                        const int16_t* pr = tile + (t_top - p) * kTileCols;
                        a = pr[ejm];
                        b = pr[ej];
                    } else {
                        const int16_t* pr = H + static_cast<uint64_t>(p) * lpa;
                        a = pr[perm(j > 0 ? j - 1 : 0)];
                        b = pr[perm(j)];
                    }
                    okd = j > 0 && hij == a + mc;
                    okv = hij == b + g;
                }
                const uint32_t md = ballot(okd);
                const uint32_t mv = ballot(okv);
                const uint32_t sel = md ? md : mv;
                if (sel) {
                    found_p = shfl(p, ffs_(sel) - 1);
                    move = md ? 1 : 2;
                }
            } else {
                move = traceback_step_wide(H, rec, pred_ovf, ki, lane, g, i, j, npe, hij, mc, lpa, &found_p);
            }
            if (!move) move = 3;
            if (move == 1) {
                if (lane == 0) aln[j - 1] = static_cast<uint16_t>(node);
                i = found_p;
                --j;
            } else if (move == 2) {
                i = found_p;
            } else {
                if (j == 0) {  // unreachable for a consistent matrix (column 0 always has a vertical move)
                    fail(kWinInternal);
                    break;
                }
                if (lane == 0) aln[j - 1] = kNone;
                --j;
            }
        }
        /* row 0: remaining columns are horizontal moves (insertions at the start of the read) */
        for (uint32_t c = lane; c < j; c += 32) aln[c] = kNone;
        syncwarp();
    }

    /* ---------------------------------------------------------------- graph merge (graph.cpp:155-247) +
     * incremental order update.  aln[j] = aligned node or kNone for every read position j. */
    RP_DEV void add_alignment(const uint8_t* seq, const uint8_t* w, uint32_t len) {
        const uint32_t n_old = N;
        uint16_t* delta = reinterpret_cast<uint16_t*>(smem);  // n_old + 2 counters (ring is idle now)
        /* All phases handle kU positions per lane at a time with the loads of one dependency level issued
         * together (the graph is in HBM: every level costs a full memory latency). */
        constexpr int kU = 4;
        /* Phase A: target node per position (existing node, aligned sibling with the same character, or new) */
        uint32_t n_new = 0;
        for (uint32_t j0 = 0; j0 < len; j0 += 32 * kU) {
            uint32_t a[kU], c[kU], ca[kU], na[kU];
#pragma unroll
            for (int u = 0; u < kU; ++u) {
                uint32_t j = j0 + u * 32 + lane;
                a[u] = j < len ? aln[j] : kNone;
                c[u] = j < len ? seq[j] : 0;
            }
#pragma unroll
            for (int u = 0; u < kU; ++u) {
                ca[u] = a[u] != kNone ? code[a[u]] : 0;
                na[u] = a[u] != kNone ? al_cnt[a[u]] : 0;
            }
#pragma unroll
            for (int u = 0; u < kU; ++u) {
                uint32_t j = j0 + u * 32 + lane;
                bool is_new = false;
                uint32_t tgt = kNone, anchor = kNone;
                if (j < len) {
                    if (a[u] == kNone) {
                        is_new = true;
                    } else if (ca[u] == c[u]) {
                        tgt = a[u];
                    } else {
                        for (uint32_t k = 0; k < na[u]; ++k) {
                            uint32_t sb = al[a[u] * ka + k];
                            if (code[sb] == c[u]) {
                                tgt = sb;
                                break;
                            }
                        }
                        if (tgt == kNone) {
                            is_new = true;
                            anchor = a[u];
                        }
                    }
                }
                uint32_t tot;
                uint32_t pos = warp_rank(is_new, &tot);
                if (is_new) {
                    uint32_t id = n_old + n_new + pos;
                    tgt = id;
                    if (id < nmax) {
                        code[id] = static_cast<uint8_t>(c[u]);
                        flags[id] = 0;
                        in_cnt[id] = 0;
                        cov[id] = 0;
                        al_cnt[id] = 0;
                    }
                    newlist[n_new + pos] = (j << 16) | anchor;  // anchor kNone => unaligned insertion
                }
                if (j < len) cur[j] = static_cast<uint16_t>(tgt);
                n_new += tot;
            }
        }
        if (n_old + n_new > nmax) {
            fail(kWinNodeLimit);
            return;
        }
        syncwarp();
        /* Phase B: aligned-cluster membership of new nodes (graph.cpp:221-229) */
        bool lim_a = false;
        for (uint32_t k0 = 0; k0 < n_new; k0 += 32) {
            uint32_t k = k0 + lane;
            if (k < n_new) {
                uint32_t e = newlist[k];
                uint32_t anchor = e & 0xffffu;
                uint32_t id = n_old + k;
                if (anchor != kNone) {
                    uint32_t na = al_cnt[anchor];
                    if (na + 1 > ka) {
                        lim_a = true;
                    } else {
                        for (uint32_t q = 0; q < na; ++q) {
                            uint32_t sb = al[anchor * ka + q];
                            al[id * ka + q] = static_cast<uint16_t>(sb);
                            uint32_t ns = al_cnt[sb];
                            al[sb * ka + ns] = static_cast<uint16_t>(id);
                            al_cnt[sb] = static_cast<uint8_t>(ns + 1);
                        }
                        al[id * ka + na] = static_cast<uint16_t>(anchor);
                        al_cnt[id] = static_cast<uint8_t>(na + 1);
                        al[anchor * ka + na] = static_cast<uint16_t>(id);
                        al_cnt[anchor] = static_cast<uint8_t>(na + 1);
                    }
                }
            }
        }
        if (ballot(lim_a)) {
            fail(kWinAlignedLimit);
            return;
        }
        syncwarp();
        /* order keys K_j (non-decreasing along the read): for an old target, or the anchor of an aligned new
         * node, the last rank of its aligned-cluster block; an unaligned insertion inherits K_{j-1} */
        int32_t run = 0;  // K_{-1} = 0: right after the root
        for (uint32_t j0 = 0; j0 < len; j0 += 32 * kU) {
            uint32_t bn[kU], rk0[kU], nal[kU];
#pragma unroll
            for (int u = 0; u < kU; ++u) {
                uint32_t j = j0 + u * 32 + lane;
                bn[u] = j < len ? cur[j] : kNone;
            }
#pragma unroll
            for (int u = 0; u < kU; ++u)
                if (bn[u] != kNone && bn[u] >= n_old) bn[u] = newlist[bn[u] - n_old] & 0xffffu;  // anchor or kNone
#pragma unroll
            for (int u = 0; u < kU; ++u) {
                rk0[u] = bn[u] != kNone ? rank_of[bn[u]] : 0;
                nal[u] = bn[u] != kNone ? al_cnt[bn[u]] : 0;
            }
#pragma unroll
            for (int u = 0; u < kU; ++u) {
                uint32_t j = j0 + u * 32 + lane;
                uint32_t rmax = rk0[u];
                for (uint32_t q = 0; q < nal[u]; ++q) {
                    uint32_t sb = al[bn[u] * ka + q];
                    if (sb < n_old) {
                        uint32_t rs = rank_of[sb];
                        if (rs > rmax) rmax = rs;
                    }
                }
                int32_t inc = warp_incl_max(static_cast<int32_t>(rmax));
                if (inc < run) inc = run;
                if (j < len) aln[j] = static_cast<uint16_t>(inc);  // aln now holds K_j
                run = shfl(inc, 31);
            }
        }
        syncwarp();
        /* Phase C: edges (graph.cpp:81-91,236-243) + per-node sequence counters (Node::Coverage, :32-47) */
        bool lim_e = false;
        for (uint32_t j0 = 0; j0 < len; j0 += 32 * kU) {
            uint32_t c[kU], pv[kU], ni[kU], cv[kU];
            int32_t wt[kU];
            uint64_t t4[kU];
#pragma unroll
            for (int u = 0; u < kU; ++u) {
                uint32_t j = j0 + u * 32 + lane;
                c[u] = j < len ? cur[j] : kNone;
                pv[u] = (j < len && j > 0) ? cur[j - 1] : kNone;
                wt[u] = (j < len && j > 0) ? static_cast<int32_t>(w[j - 1]) + static_cast<int32_t>(w[j]) : 0;
            }
#pragma unroll
            for (int u = 0; u < kU; ++u) {
                ni[u] = c[u] != kNone ? in_cnt[c[u]] : 0;
                cv[u] = c[u] != kNone ? cov[c[u]] : 0;
                t4[u] = c[u] != kNone ? *reinterpret_cast<const uint64_t*>(in_tail + c[u] * ki) : 0;
            }
#pragma unroll
            for (int u = 0; u < kU; ++u) {
                if (c[u] == kNone) continue;
                if (len >= 2) cov[c[u]] = static_cast<uint16_t>(cv[u] + 1);
                if (pv[u] == kNone) continue;
                uint32_t q = 0;
                for (; q < ni[u]; ++q) {
                    uint32_t t = q < 4 ? (static_cast<uint32_t>(t4[u] >> (16 * q)) & 0xffffu) : in_tail[c[u] * ki + q];
                    if (t == pv[u]) break;
                }
                if (q < ni[u]) {
                    in_w[c[u] * ki + q] += wt[u];
                } else if (ni[u] < ki && ni[u] < 127) {
                    in_tail[c[u] * ki + ni[u]] = static_cast<uint16_t>(pv[u]);
                    in_w[c[u] * ki + ni[u]] = wt[u];
                    in_cnt[c[u]] = static_cast<uint8_t>(ni[u] + 1);
                    flags[pv[u]] |= 1;  // every pv is a distinct node, so lanes never touch the same byte
                } else {
                    lim_e = true;
                }
            }
        }
        if (ballot(lim_e)) {
            fail(kWinEdgeLimit);
            return;
        }
        syncwarp();
        /* Phase D: merge new nodes into the processing order.  New node k (read order) with key K goes to
         * rank K + 1 + k; an old node at rank q moves to q + #{new nodes with key < q}. */
        if (n_new > 0) {
            if ((n_old + 4) * 2 > smem_bytes) {
                fail(kWinNodeLimit);
                return;
            }
            for (uint32_t q = lane; q < n_old + 2; q += 32) delta[q] = 0;
            syncwarp();
            for (uint32_t k = lane; k < n_new; k += 32) {
                uint32_t j = newlist[k] >> 16;
                uint32_t key = aln[j];  // aligned new node: end of its cluster block; insertion: K_{j-1}
                uint32_t id = n_old + k;
                uint32_t nr = key + 1 + k;
                order_nxt[nr] = static_cast<uint16_t>(id);
                rank_of[id] = static_cast<uint16_t>(nr);
#if defined(RP_HOST_SIM)
                delta[key + 1] = static_cast<uint16_t>(delta[key + 1] + 1);
#else
                atomicAdd(reinterpret_cast<unsigned int*>(delta + ((key + 1) & ~1u)), (key + 1) & 1u ? 0x10000u : 1u);
#endif
            }
            syncwarp();
            /* inclusive prefix of the counters in shared memory, then a remap pass whose HBM loads are batched */
            uint32_t carry = 0;
            for (uint32_t q0 = 1; q0 <= n_old; q0 += 32) {
                uint32_t q = q0 + lane;
                uint32_t d = q <= n_old ? delta[q] : 0;
                uint32_t inc = warp_incl_sum(d) + carry;
                if (q <= n_old) delta[q] = static_cast<uint16_t>(inc);
                carry = shfl(inc, 31);
            }
            syncwarp();
            constexpr int kV = 8;
            for (uint32_t q0 = 1; q0 <= n_old; q0 += 32 * kV) {
                uint32_t v[kV];
#pragma unroll
                for (int u = 0; u < kV; ++u) {
                    uint32_t q = q0 + u * 32 + lane;
                    v[u] = q <= n_old ? order[q] : 0;
                }
#pragma unroll
                for (int u = 0; u < kV; ++u) {
                    uint32_t q = q0 + u * 32 + lane;
                    if (q <= n_old) {
                        uint32_t nr = q + delta[q];
                        order_nxt[nr] = static_cast<uint16_t>(v[u]);
                        rank_of[v[u]] = static_cast<uint16_t>(nr);
                    }
                }
            }
            if (lane == 0) order_nxt[0] = kNone;
            uint16_t* t = order;
            order = order_nxt;
            order_nxt = t;
            N = n_old + n_new;
        }
        ++n_added;
        syncwarp();
    }

    /* ---------------------------------------------------------------- consensus (graph.cpp:433-516, 377-398)
     * Heaviest bundle.  The pass over the nodes in spoa's rank order is inherently serial (lane 0), but what it
     * reads is not: the warp stages 32 nodes at a time (id, in-degree, first four in-edge tails and weights)
     * into shared memory with coalesced gathers, and scores / best predecessors live in shared memory too
     * when the graph is small enough (int32 scores are exact while the sum of all edge weights < 2^31). */
    struct CStage {
        uint32_t w[4];
        uint16_t t[4];
        uint16_t it;
        uint16_t ni;
        uint32_t pad;
    };

    template <typename ScoreT>
    RP_DEV void better_t(ScoreT* sc, uint16_t* cp, uint32_t it, uint32_t t, int64_t wgt) {
        int64_t si = sc[it];
        if (si < wgt || (si == wgt && cp[it] != kNone && sc[cp[it]] <= sc[t])) {
            sc[it] = static_cast<ScoreT>(wgt);
            cp[it] = static_cast<uint16_t>(t);
        }
    }

    /* one pass over ranks [r_begin, N): graph.cpp:442-456 (skip_dead = false) or :493-513 (true) */
    template <typename ScoreT>
    RP_DEV uint32_t bundle_pass(ScoreT* sc, uint16_t* cp, CStage* stg, uint32_t r_begin, bool skip_dead) {
        uint32_t mx = kNone;
        for (uint32_t r0 = r_begin; r0 < N; r0 += 32) {
            const uint32_t r = r0 + lane;
            if (r < N) {
                const uint32_t it = sorder[r];
                CStage e;
                e.it = static_cast<uint16_t>(it);
                e.ni = in_cnt[it];
                const uint64_t t4 = *reinterpret_cast<const uint64_t*>(in_tail + it * ki);
                const U4 w4 = *reinterpret_cast<const U4*>(in_w + it * ki);
                e.t[0] = static_cast<uint16_t>(t4);
                e.t[1] = static_cast<uint16_t>(t4 >> 16);
                e.t[2] = static_cast<uint16_t>(t4 >> 32);
                e.t[3] = static_cast<uint16_t>(t4 >> 48);
                e.w[0] = w4.x; e.w[1] = w4.y; e.w[2] = w4.z; e.w[3] = w4.w;
                e.pad = 0;
                stg[lane] = e;
            }
            syncwarp();
            if (lane == 0) {
                const uint32_t cnt = N - r0 < 32 ? N - r0 : 32;
                for (uint32_t q = 0; q < cnt; ++q) {
                    const CStage& e = stg[q];
                    const uint32_t it = e.it;
                    if (skip_dead) {
                        sc[it] = -1;
                        cp[it] = kNone;
                    }
                    for (uint32_t k = 0; k < e.ni; ++k) {
                        const uint32_t t = k < 4 ? e.t[k] : in_tail[it * ki + k];
                        if (skip_dead && sc[t] == -1) continue;
                        const int64_t wgt = k < 4 ? static_cast<int32_t>(e.w[k]) : in_w[it * ki + k];
                        better_t(sc, cp, it, t, wgt);
                    }
                    if (cp[it] != kNone) sc[it] = static_cast<ScoreT>(sc[it] + sc[cp[it]]);
                    if (mx == kNone || sc[mx] < sc[it]) mx = it;
                }
            }
            syncwarp();
        }
        return shfl(mx, 0);
    }

    template <typename ScoreT>
    RP_DEV uint32_t bundle(ScoreT* sc, uint16_t* cp, CStage* stg) {
        for (uint32_t v = lane; v < N; v += 32) {
            sc[v] = -1;
            cp[v] = kNone;
        }
        syncwarp();
        uint32_t mx = bundle_pass(sc, cp, stg, 0, false);
        /* branch completion (graph.cpp:478-516) while the best node still has out-edges */
        while (flags[mx] & 1) {
            /* heads of mx's out-edges = nodes with an in-edge from mx; their other tails are invalidated */
            for (uint32_t v = lane; v < N; v += 32) {
                uint32_t ni = in_cnt[v];
                bool hit = false;
                for (uint32_t k = 0; k < ni; ++k) hit |= (in_tail[v * ki + k] == mx);
                if (hit)
                    for (uint32_t k = 0; k < ni; ++k) {
                        uint32_t t = in_tail[v * ki + k];
                        if (t != mx) sc[t] = -1;
                    }
            }
            syncwarp();
            uint32_t nm = bundle_pass(sc, cp, stg, static_cast<uint32_t>(srank[mx]) + 1, true);
            if (nm == kNone) break;  // cannot happen: a node with out-edges has successors of higher rank
            mx = nm;
        }
        /* hand the predecessor chain over in the HBM array the emitter reads */
        if (cp != cpred)
            for (uint32_t v = lane; v < N; v += 32) cpred[v] = cp[v];
        syncwarp();
        return mx;
    }

    RP_DEV uint32_t consensus(uint8_t* out, uint16_t* out_cov, uint32_t out_cap, bool trim, uint32_t n_seq) {
        spoa_sort(false);
        if (status != kWinOk) return 0;
        /* int32 scores in shared memory are exact iff the total edge weight fits */
        uint64_t wsum = 0;
        for (uint32_t v = lane; v < N; v += 32) {
            uint32_t ni = in_cnt[v];
            for (uint32_t k = 0; k < ni; ++k) wsum += static_cast<uint32_t>(in_w[v * ki + k]);
        }
        for (int d = 16; d > 0; d >>= 1) wsum += shfl_down(wsum, d);
        wsum = shfl(wsum, 0);
        const uint32_t npad = (N + 7) & ~7u;
        uint32_t mx;
        if (wsum < 0x7fffffffull && npad * 6 + sizeof(CStage) * 32 <= smem_bytes && !(P->debug_flags & 1)) {
            int32_t* sc = reinterpret_cast<int32_t*>(smem);
            uint16_t* cp = reinterpret_cast<uint16_t*>(smem + npad * 4);
            CStage* stg = reinterpret_cast<CStage*>(smem + npad * 6);
            mx = bundle(sc, cp, stg);
        } else {
            mx = bundle(score, cpred, reinterpret_cast<CStage*>(smem));
        }
        /* walk predecessors back; the path is stored reversed in `stack`, then emitted forward */
        uint32_t clen = 0;
        if (lane == 0) {
            uint32_t v = mx;
            while (v != kNone && clen < P->lim.stack_cap) {
                stack[clen++] = static_cast<uint16_t>(v);
                v = cpred[v];
            }
        }
        clen = shfl(clen, 0);
        syncwarp();
        /* coverage (graph.cpp:388-394): node + its aligned nodes, counted as sequences through the node */
        uint32_t thr = (n_seq - 1) / 2;
        uint32_t first = 0xffffffffu, last = 0;
        for (uint32_t k = lane; k < clen; k += 32) {
            uint32_t v = stack[clen - 1 - k];
            uint32_t c = cov[v];
            uint32_t na = al_cnt[v];
            for (uint32_t q = 0; q < na; ++q) c += cov[al[v * ka + q]];
            if (c > 0xffffu) c = 0xffffu;
            dp_rank[k] = static_cast<uint16_t>(c);  // dp_rank is idle here: per-base coverage scratch
            if (c >= thr) {
                if (k < first) first = k;
                if (k + 1 > last) last = k + 1;
            }
        }
        return finish_consensus(out, out_cov, out_cap, trim, clen, first, last);
    }

    RP_DEV uint32_t finish_consensus(uint8_t* out, uint16_t* out_cov, uint32_t out_cap, bool trim, uint32_t clen,
                                      uint32_t first, uint32_t last) {
        for (int d = 16; d > 0; d >>= 1) {
            uint32_t of = shfl_down(first, d), ol = shfl_down(last, d);
            if (of < first) first = of;
            if (ol > last) last = ol;
        }
        first = shfl(first, 0);
        last = shfl(last, 0);
        uint32_t b = 0, e = clen;  // [b, e)
        if (trim) {
            /* window.cpp:125-146: begin = first index with coverage >= thr, end = last such index;
             * keep [begin, end] only when begin < end */
            if (first != 0xffffffffu && last >= 1 && first < last - 1) {
                b = first;
                e = last;
            }
        }
        uint32_t n = e - b;
        if (n > out_cap) {
            fail(kWinInternal);
            return 0;
        }
        for (uint32_t k = b + lane; k < e; k += 32) {
            uint32_t v = stack[clen - 1 - k];
            out[k - b] = code[v];
            out_cov[k - b] = cur_cov(k);
        }
        syncwarp();
        return n;
    }

    RP_DEV uint16_t cur_cov(uint32_t k) const { return dp_rank[k]; }
};

/* int16 is safe iff spoa's own criterion holds (alignment_engine.cpp:101-110, simd impl :699-745) */
RP_DEV bool fits_int16(int32_t m, int32_t g, int64_t len, int64_t nodes) {
    if (g < -kMaxGapInt16) return false;
    int64_t i = len + 8, j = nodes;
    int64_t mn = i < j ? i : j;
    int64_t df = i > j ? i - j : j - i;
    int64_t a = -1 * (m * mn + g * df);
    int64_t b = g * i + g * j;
    int64_t worst = a < b ? a : b;
    return worst >= static_cast<int64_t>(-32768 + 1024);
}

/* Processes one window end to end. `slot`: this warp's HBM scratch, `smem`: this warp's shared memory. */
RP_DEV void poa_window(const PoaParams& P, uint32_t w, uint8_t* slot, uint8_t* smem) {
    PoaWarp W;
    W.bind(&P, slot, smem, P.smem_per_warp);
    W.status = kWinOk;
    const uint32_t s0 = P.win_first[w], s1 = P.win_first[w + 1];
    W.alpha = P.win_alpha[w];
    uint32_t nc = 0;
    while (nc < 8 && ((W.alpha >> (8 * nc)) & 0xff) != 0) ++nc;
    W.ncodes = nc;
    const uint32_t lane = W.lane;
    uint8_t* out = P.cons + P.out_off[w];
    uint16_t* out_cov = P.cons_cov + P.out_off[w];

    uint32_t boff = P.seq_off[s0];
    uint32_t blen = P.seq_off[s0 + 1] - boff;
    if (blen > P.lim.nmax) {
        if (lane == 0) {
            P.status[w] = kWinNodeLimit;
            P.cons_len[w] = 0;
        }
        return;
    }
    W.init_backbone(P.bases + boff, P.weights + boff, blen);

    for (uint32_t s = s0 + 1; s < s1 && W.status == kWinOk; ++s) {
        uint32_t off = P.seq_off[s];
        uint32_t len = P.seq_off[s + 1] - off;
        const uint8_t* seq = P.bases + off;
        const uint8_t* wts = P.weights + off;
        if (len > P.lim.lmax) {
            W.fail(kWinSeqTooLong);
            break;
        }
        bool sub = !(P.seq_flags[s] & 1);
        uint32_t nrows = W.N;
        if (sub) {
            W.mark_subgraph(P.seq_begin[s], P.seq_end[s]);
            if (W.status != kWinOk) break;
            nrows = W.build_dp_order_subgraph();
        }
        /* spoa picks its int32 engine from a worst-case bound over ALL nodes of the graph (fits_int16).  That bound
         * is far from what a real matrix holds (column 0 only falls by |g| per level of graph DEPTH), and exact
         * arithmetic gives the same answer in either width, so a window that fails the bound is still computed in
         * int16 and the finished matrix is checked for having stayed in range (matrix_in_range). */
        bool verify_range = false;
        if (!fits_int16(P.match, P.gap, len, nrows)) {
            const int64_t short_side = static_cast<int64_t>(len) + 8 < nrows ? static_cast<int64_t>(len) + 8 : nrows;
            if (P.gap < -kMaxGapInt16 || static_cast<int64_t>(P.match) * short_side > 32767 - 1024) {
                W.fail(kWinNeedsInt32);
                break;
            }
            verify_range = true;
        }
        uint32_t lpa = (len + 1 + kChunkCols - 1) / kChunkCols * kChunkCols;
        /* shared memory split: profile rows first, the rest is the ring of recent DP rows */
        /* shared memory (chunk-local): profile chunk | ring of the last R DP rows */
        const uint32_t prof_bytes = W.ncodes * kChunkCols * 2;
        const uint32_t tile_rows = P.tile_rows ? P.tile_rows : 96;
        const uint32_t tb_bytes = tile_rows * (64 + 16 + 2) + 16 + len;  // traceback: H tile | records | nodes | read
        if (prof_bytes + 2 * kChunkCols * 2 > P.smem_per_warp || tb_bytes > P.smem_per_warp) {
            W.fail(kWinSeqTooLong);
            break;
        }
        uint32_t ring_rows = (P.smem_per_warp - prof_bytes) / (kChunkCols * 2);
        if (ring_rows > 32) ring_rows = 32;
        W.prof = reinterpret_cast<int16_t*>(smem);
        W.ring = reinterpret_cast<int16_t*>(smem + prof_bytes);
        uint32_t pred_rows = W.build_program(nrows, sub);
        uint32_t best_row, n_best;
        int32_t best;
        W.dp(seq, nrows, len, lpa, ring_rows, &best_row, &best, &n_best);
        if (verify_range && !W.matrix_in_range(nrows, len, lpa)) {
            W.fail(kWinNeedsInt32);
            break;
        }
        if (n_best > 1) {
            best_row = W.resolve_sink_tie(nrows, len, lpa, best, sub);
            if (W.status != kWinOk) break;
        }
        W.traceback(best_row, len, lpa, seq, sub);
        W.add_alignment(seq, wts, len);
        if (P.stats) pred_rows = warp_incl_sum(pred_rows);
        if (P.stats) pred_rows = shfl(pred_rows, 31);
        if (P.stats && lane == 0) {
#if !defined(RP_HOST_SIM)
            atomicAdd(reinterpret_cast<unsigned long long*>(P.stats + 3),
                      static_cast<unsigned long long>(pred_rows) * (len + 1));
            atomicAdd(reinterpret_cast<unsigned long long*>(P.stats), 1ull);
            atomicAdd(reinterpret_cast<unsigned long long*>(P.stats + 1),
                      static_cast<unsigned long long>(nrows + 1) * (len + 1));
            if (n_best > 1) atomicAdd(reinterpret_cast<unsigned long long*>(P.stats + 2), 1ull);
#else
            P.stats[0] += 1;
            P.stats[3] += static_cast<uint64_t>(pred_rows) * (len + 1);
            P.stats[1] += static_cast<uint64_t>(nrows + 1) * (len + 1);
            if (n_best > 1) P.stats[2] += 1;
#endif
        }
    }
    uint32_t clen = 0;
    if (W.status == kWinOk) clen = W.consensus(out, out_cov, P.out_cap[w], (P.win_flags[w] & 1) != 0, s1 - s0);
    if (lane == 0) {
        P.status[w] = W.status;
        P.cons_len[w] = W.status == kWinOk ? clen : 0;
    }
    syncwarp();
}

}  // namespace rp
