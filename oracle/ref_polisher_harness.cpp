/*
 * ref_polisher_harness.cpp — TEST INFRASTRUCTURE ONLY.
 *
 * Drives the UNMODIFIED reference Polisher (/root/reference/src/polisher.cpp: createPolisher, initialize,
 * polish) on real input files and exports (a) the windows it built — backbone + layers + qualities +
 * positions exactly as Window::generate_consensus sees them — as the flat "window set" arrays, and (b) the
 * reference's own per-window consensus and final polished sequences.  Used by
 * tests/golden/make_lambda_windows.py to create the real-data fixture (λ-phage sample, BASELINE config 1).
 * The `#define private/protected public` below only opens the reference headers for reading; no reference
 * source is modified or copied.
 */
#include <cstdint>
#include <cstring>
#include <memory>
#include <string>
#include <vector>

#define private public
#define protected public
#include "overlap.hpp"
#include "polisher.hpp"
#include "window.hpp"
#undef private
#undef protected
#include <cstdio>
#include <cstdlib>

#include "bioparser/parser.hpp"
#include "logger.hpp"
#include "sequence.hpp"
#include "spoa/spoa.hpp"
#include "thread_pool/thread_pool.hpp"

namespace {
struct Handle {
    std::unique_ptr<racon::Polisher> polisher;
    std::vector<std::shared_ptr<racon::Window>> windows;  // copies keep the windows alive across polish()
    std::vector<std::unique_ptr<racon::Sequence>> polished;
};

/* Same Polisher, one hook: after the reference has aligned every overlap and found its breaking points
 * (Polisher::find_overlap_breaking_points -> Overlap::find_breaking_points, src/overlap.cpp:172-292), write the
 * sequences and the overlaps (coordinates + breaking points) to the file named by $REFPOL_OVERLAP_DUMP.  This is
 * the fixture the CIGAR->breaking-points->windows row is pinned against. */
struct DumpPolisher : public racon::Polisher {
    using racon::Polisher::Polisher;
    void find_overlap_breaking_points(std::vector<std::unique_ptr<racon::Overlap>>& overlaps) override {
        racon::Polisher::find_overlap_breaking_points(overlaps);
        const char* path = std::getenv("REFPOL_OVERLAP_DUMP");
        if (!path) return;
        FILE* f = std::fopen(path, "wb");
        if (!f) return;
        auto u32 = [&](uint32_t v) { std::fwrite(&v, 4, 1, f); };
        auto u64 = [&](uint64_t v) { std::fwrite(&v, 8, 1, f); };
        auto rc = [](const std::string& r) {
            std::string d(r.rbegin(), r.rend());
            for (auto& c : d) {
                switch (c) {
                    case 'A': c = 'T'; break;
                    case 'C': c = 'G'; break;
                    case 'G': c = 'C'; break;
                    case 'T': c = 'A'; break;
                    default: break;
                }
            }
            return d;
        };
        u64(sequences_.size());
        for (const auto& sq : sequences_) {
            std::string data = !sq->data().empty() ? sq->data() : rc(sq->reverse_complement());
            std::string qual = !sq->quality().empty()
                                   ? sq->quality()
                                   : std::string(sq->reverse_quality().rbegin(), sq->reverse_quality().rend());
            u32(static_cast<uint32_t>(sq->name().size()));
            std::fwrite(sq->name().data(), 1, sq->name().size(), f);
            u64(data.size());
            std::fwrite(data.data(), 1, data.size(), f);
            u64(qual.size());
            std::fwrite(qual.data(), 1, qual.size(), f);
        }
        u64(overlaps.size());
        for (const auto& o : overlaps) {
            u32(o->q_id_); u32(o->t_id_); u32(o->strand_);
            u32(o->q_begin_); u32(o->q_end_); u32(o->q_length_);
            u32(o->t_begin_); u32(o->t_end_); u32(o->t_length_);
            u32(static_cast<uint32_t>(o->breaking_points_.size()));
            for (const auto& bp : o->breaking_points_) {
                u32(bp.first);
                u32(bp.second);
            }
        }
        std::fclose(f);
    }
};
}  // namespace

extern "C" {

void* ref_polisher_open(const char* reads, const char* overlaps, const char* target, int fragment_correction,
                        uint32_t window_length, double quality_threshold, double error_threshold, int trim,
                        int8_t match, int8_t mismatch, int8_t gap, uint32_t threads) {
    Handle* h = new Handle();
    h->polisher = racon::createPolisher(reads, overlaps, target,
                                        fragment_correction ? racon::PolisherType::kF : racon::PolisherType::kC,
                                        window_length, quality_threshold, error_threshold, trim != 0, match, mismatch,
                                        gap, threads);
    if (std::getenv("REFPOL_OVERLAP_DUMP")) {
        /* re-seat the parsers the factory chose into the hooked subclass (same constructor arguments) */
        racon::Polisher* p = h->polisher.get();
        std::unique_ptr<racon::Polisher> hooked(new DumpPolisher(
            std::move(p->sparser_), std::move(p->oparser_), std::move(p->tparser_), p->type_, window_length,
            quality_threshold, error_threshold, trim != 0, match, mismatch, gap, threads));
        h->polisher = std::move(hooked);
    }
    h->polisher->initialize();
    h->windows = h->polisher->windows_;
    return h;
}

/* counts[0] windows, [1] sequences (backbones + layers), [2] bases */
void ref_polisher_counts(void* hv, uint64_t* counts) {
    Handle* h = static_cast<Handle*>(hv);
    counts[0] = h->windows.size();
    counts[1] = counts[2] = 0;
    for (const auto& w : h->windows) {
        counts[1] += w->sequences_.size();
        for (const auto& s : w->sequences_) counts[2] += s.second;
    }
}

void ref_polisher_export(void* hv, char* bases, char* quals, uint64_t* seq_off, uint8_t* seq_has_qual,
                         uint32_t* seq_begin, uint32_t* seq_end, uint32_t* win_first, uint8_t* win_type,
                         uint64_t* win_target, uint32_t* win_rank) {
    Handle* h = static_cast<Handle*>(hv);
    uint64_t nb = 0, ns = 0;
    seq_off[0] = 0;
    for (size_t w = 0; w < h->windows.size(); ++w) {
        const auto& win = h->windows[w];
        win_first[w] = static_cast<uint32_t>(ns);
        win_type[w] = win->type_ == racon::WindowType::kTGS ? 1 : 0;
        win_target[w] = win->id_;
        win_rank[w] = win->rank_;
        for (size_t s = 0; s < win->sequences_.size(); ++s) {
            uint32_t len = win->sequences_[s].second;
            std::memcpy(bases + nb, win->sequences_[s].first, len);
            const char* q = win->qualities_[s].first;
            if (q) {
                std::memcpy(quals + nb, q, len);
                seq_has_qual[ns] = 1;
            } else {
                std::memset(quals + nb, '!', len);
                seq_has_qual[ns] = 0;
            }
            seq_begin[ns] = win->positions_[s].first;
            seq_end[ns] = win->positions_[s].second;
            nb += len;
            seq_off[++ns] = nb;
        }
    }
    win_first[h->windows.size()] = static_cast<uint32_t>(ns);
}

/* runs Polisher::polish; returns the number of polished sequences */
uint32_t ref_polisher_polish(void* hv) {
    Handle* h = static_cast<Handle*>(hv);
    h->polisher->polish(h->polished, false);
    return static_cast<uint32_t>(h->polished.size());
}

uint32_t ref_polisher_window_consensus(void* hv, uint32_t w, char* out, uint32_t cap) {
    Handle* h = static_cast<Handle*>(hv);
    const std::string& c = h->windows[w]->consensus();
    if (c.size() > cap) return 0xffffffffu;
    std::memcpy(out, c.data(), c.size());
    return static_cast<uint32_t>(c.size());
}

uint64_t ref_polisher_polished(void* hv, uint32_t i, char* name, uint32_t name_cap, char* data, uint64_t data_cap) {
    Handle* h = static_cast<Handle*>(hv);
    const auto& s = h->polished[i];
    std::strncpy(name, s->name().c_str(), name_cap - 1);
    name[name_cap - 1] = 0;
    if (s->data().size() > data_cap) return ~0ull;
    std::memcpy(data, s->data().data(), s->data().size());
    return s->data().size();
}

void ref_polisher_close(void* hv) { delete static_cast<Handle*>(hv); }

}  // extern "C"
