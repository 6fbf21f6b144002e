/*
 * host_mirror.cpp — see host_mirror.hpp.  Thin: all work happens behind the C ABI on the GPU.
 */
#include "host_mirror.hpp"

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <iterator>

#include "racon_b200.h"
#include "reads_io.hpp"

namespace racon_b200 {

std::shared_ptr<Window> createWindow(uint64_t id, uint32_t rank, WindowType type, const char* backbone,
                                     uint32_t backbone_length, const char* quality, uint32_t quality_length) {
    if (backbone_length == 0 || backbone_length != quality_length) {  // window.cpp:19-23
        fprintf(stderr, "[racon_b200::createWindow] error: empty backbone sequence/unequal quality length!\n");
        exit(1);
    }
    return std::shared_ptr<Window>(new Window(id, rank, type, backbone, backbone_length, quality, quality_length));
}

Window::Window(uint64_t id, uint32_t rank, WindowType type, const char* backbone, uint32_t backbone_length,
               const char* quality, uint32_t quality_length)
    : id_(id), rank_(rank), type_(type), consensus_(), sequences_(), qualities_(), positions_() {
    sequences_.emplace_back(backbone, backbone_length);
    qualities_.emplace_back(quality, quality_length);
    positions_.emplace_back(0, 0);
}

Window::~Window() {}

void Window::add_layer(const char* sequence, uint32_t sequence_length, const char* quality, uint32_t quality_length,
                       uint32_t begin, uint32_t end) {
    if (sequence_length == 0 || begin == end) return;  // window.cpp:45-47
    if (quality != nullptr && sequence_length != quality_length) {
        fprintf(stderr, "[racon_b200::Window::add_layer] error: unequal quality size!\n");
        exit(1);
    }
    if (begin >= end || begin > sequences_.front().second || end > sequences_.front().second) {
        fprintf(stderr, "[racon_b200::Window::add_layer] error: layer begin and end positions are invalid!\n");
        exit(1);
    }
    sequences_.emplace_back(sequence, sequence_length);
    qualities_.emplace_back(quality, quality_length);
    positions_.emplace_back(begin, end);
}

void Window::add_layer(const char* sequence, uint32_t sequence_length, const char* quality, uint32_t quality_length,
                       uint32_t begin, uint32_t end, const Origin& origin) {
    const size_t before = sequences_.size();
    add_layer(sequence, sequence_length, quality, quality_length, begin, end);
    if (sequences_.size() > before && origins_.size() == before) origins_.push_back(origin);
}

std::atomic<uint32_t> BatchProcessor::batches{0};

std::unique_ptr<BatchProcessor> createBatch(uint32_t max_window_depth, uint32_t device, size_t avail_mem, int8_t gap,
                                            int8_t mismatch, int8_t match, bool cuda_banded_alignment,
                                            uint32_t window_length, bool trim) {
    return std::unique_ptr<BatchProcessor>(new BatchProcessor(max_window_depth, device, avail_mem, gap, mismatch,
                                                              match, cuda_banded_alignment, window_length, trim));
}

BatchProcessor::BatchProcessor(uint32_t max_window_depth, uint32_t device, size_t avail_mem, int8_t gap,
                               int8_t mismatch, int8_t match, bool cuda_banded_alignment, uint32_t window_length,
                               bool trim)
    : trim_(trim) {
    bid_ = BatchProcessor::batches++;
    rp_status s = rp_poa_create(&poa_, static_cast<int>(device), avail_mem, match, mismatch, gap,
                                cuda_banded_alignment ? 1 : 0, window_length, max_window_depth);
    if (s != RP_OK) {  // hard failure, like GW_CU_CHECK_ERR / std::runtime_error in the reference
        fprintf(stderr, "[racon_b200::BatchProcessor] error: %s (%s)\n", rp_strerror(s), rp_last_error());
        exit(1);
    }
}

BatchProcessor::~BatchProcessor() { rp_poa_destroy(poa_); }

bool BatchProcessor::addWindow(std::shared_ptr<Window> window) {
    const size_t n = window->sequences_.size();
    std::vector<const char*> seq(n), qual(n);
    std::vector<uint32_t> len(n), begin(n), end(n);
    for (size_t i = 0; i < n; ++i) {
        seq[i] = window->sequences_[i].first;
        len[i] = window->sequences_[i].second;
        qual[i] = window->qualities_[i].first;
        begin[i] = window->positions_[i].first;
        end[i] = window->positions_[i].second;
    }
    rp_status s;
    if (reads_ != nullptr && window->origins_.size() == n) {
        std::vector<uint32_t> id(n), offset(n);
        std::vector<uint8_t> reverse(n);
        for (size_t i = 0; i < n; ++i) {
            id[i] = window->origins_[i].seq_id;
            offset[i] = window->origins_[i].offset;
            reverse[i] = window->origins_[i].reverse ? 1 : 0;
        }
        s = rp_poa_add_window_refs(poa_, reads_, static_cast<uint32_t>(n), id.data(), offset.data(), len.data(),
                                   reverse.data(), begin.data(), end.data(),
                                   window->type_ == WindowType::kTGS ? RP_WINDOW_TGS : RP_WINDOW_NGS, trim_ ? 1 : 0);
    } else {
        s = rp_poa_add_window(poa_, static_cast<uint32_t>(n), seq.data(), len.data(), qual.data(), begin.data(),
                              end.data(), window->type_ == WindowType::kTGS ? RP_WINDOW_TGS : RP_WINDOW_NGS,
                              trim_ ? 1 : 0);
    }
    if (s == RP_BATCH_FULL) return false;
    if (s != RP_OK) {
        fprintf(stderr, "[racon_b200::BatchProcessor::addWindow] error: %s (%s)\n", rp_strerror(s), rp_last_error());
        exit(1);
    }
    windows_.push_back(window);
    return true;
}

bool BatchProcessor::hasWindows() const { return rp_poa_size(poa_) > 0; }

const std::vector<bool>& BatchProcessor::generateConsensus() {
    launch();
    return collect();
}

void BatchProcessor::launch() {
    const rp_status s = rp_poa_run(poa_);
    if (s != RP_OK) {
        fprintf(stderr, "[racon_b200::BatchProcessor::launch] error: %s (%s)\n", rp_strerror(s), rp_last_error());
        exit(1);
    }
}

const std::vector<bool>& BatchProcessor::collect() {
    const rp_status s = rp_poa_sync(poa_);
    if (s != RP_OK) {
        fprintf(stderr, "[racon_b200::BatchProcessor::collect] error: %s (%s)\n", rp_strerror(s), rp_last_error());
        exit(1);
    }
    window_consensus_status_.clear();
    failed_.clear();
    for (uint32_t i = 0; i < windows_.size(); ++i) {
        const char* c = nullptr;
        uint32_t l = 0;
        int polished = 0;
        uint32_t st = 0;
        rp_poa_fetch(poa_, i, &c, &l, nullptr, &polished);
        rp_poa_window_status(poa_, i, &st);
        if (st != RP_WIN_OK) {
            /* a window that exceeded a device limit keeps its backbone (never silently re-run on the CPU) */
            fprintf(stderr, "[racon_b200::BatchProcessor] warning: window %lu exceeded device limit %u\n",
                    static_cast<unsigned long>(windows_[i]->id_), st);
            windows_[i]->consensus_ = std::string(windows_[i]->sequences_.front().first,
                                                  windows_[i]->sequences_.front().second);
            window_consensus_status_.push_back(false);
            failed_.push_back(i);
            continue;
        }
        windows_[i]->consensus_.assign(c, l);
        window_consensus_status_.push_back(polished != 0);
    }
    return window_consensus_status_;
}

void BatchProcessor::reset() {
    windows_.clear();
    window_consensus_status_.clear();
    rp_poa_reset(poa_);
}

/* ---- BatchAligner ---- */
std::atomic<uint32_t> BatchAligner::batches{0};

std::unique_ptr<BatchAligner> createBatchAligner(uint32_t max_query_size, uint32_t max_target_size,
                                                 uint32_t max_alignments, uint32_t device_id) {
    return std::unique_ptr<BatchAligner>(new BatchAligner(max_query_size, max_target_size, max_alignments, device_id));
}

BatchAligner::BatchAligner(uint32_t max_query_size, uint32_t max_target_size, uint32_t max_alignments,
                           uint32_t device_id)
    : max_alignments_(max_alignments) {
    bid_ = batches++;
    rp_status s = rp_aln_create(&aln_, static_cast<int>(device_id), 0,
                                max_query_size > max_target_size ? max_query_size : max_target_size);
    if (s != RP_OK) {
        fprintf(stderr, "[racon_b200::BatchAligner] error: %s (%s)\n", rp_strerror(s), rp_last_error());
        exit(1);
    }
}

BatchAligner::~BatchAligner() { rp_aln_destroy(aln_); }

bool BatchAligner::addOverlap(const char* q, uint32_t q_len, const char* t, uint32_t t_len, std::string* cigar) {
    if (max_alignments_ && cigars_.size() >= max_alignments_) return false;
    rp_status s = rp_aln_add(aln_, q, q_len, t, t_len);
    if (s == RP_BATCH_FULL) return false;
    if (s != RP_OK) {
        fprintf(stderr, "[racon_b200::BatchAligner::addOverlap] error: %s (%s)\n", rp_strerror(s), rp_last_error());
        exit(1);
    }
    cigars_.push_back(cigar);
    return true;
}

void BatchAligner::alignAll() {
    rp_status s = rp_aln_run(aln_);
    if (s != RP_OK) {
        fprintf(stderr, "[racon_b200::BatchAligner::alignAll] error: %s (%s)\n", rp_strerror(s), rp_last_error());
        exit(1);
    }
}

void BatchAligner::generate_cigar_strings() {
    rp_status s = rp_aln_sync(aln_);
    if (s != RP_OK) {
        fprintf(stderr, "[racon_b200::BatchAligner::generate_cigar_strings] error: %s (%s)\n", rp_strerror(s),
                rp_last_error());
        exit(1);
    }
    for (uint32_t i = 0; i < cigars_.size(); ++i) {
        const char* c = nullptr;
        uint32_t l = 0, st = 0;
        rp_aln_fetch_cigar(aln_, i, &c, &l, nullptr, &st);
        if (cigars_[i]) cigars_[i]->assign(st == RP_ALN_OK ? c : "", st == RP_ALN_OK ? l : 0);
    }
}

void BatchAligner::reset() {
    cigars_.clear();
    rp_aln_reset(aln_);
}

std::string format_fasta(const std::string& target_name, const PolishedSequence& s) {
    std::string out;
    out.reserve(target_name.size() + s.tags.size() + s.data.size() + 3);
    out += '>';
    out += target_name;
    out += s.tags;
    out += '\n';
    out += s.data;
    out += '\n';
    return out;
}

/* ---- Polisher ---- */
Polisher::Polisher(std::vector<SequenceView> sequences, uint64_t targets_size, WindowType window_type,
                   bool fragment_correction, uint32_t window_length, double quality_threshold, bool trim, int8_t match,
                   int8_t mismatch, int8_t gap, uint32_t device)
    : sequences_(std::move(sequences)), targets_size_(targets_size), window_type_(window_type),
      fragment_correction_(fragment_correction), window_length_(window_length), quality_threshold_(quality_threshold),
      trim_(trim), match_(match), mismatch_(mismatch), gap_(gap), device_(device),
      dummy_quality_(window_length, '!'),  // polisher.cpp:174
      reverse_complement_(sequences_.size()), reverse_quality_(sequences_.size()) {
    if (window_length == 0) {  // polisher.cpp:71-75
        fprintf(stderr, "[racon_b200::Polisher] error: invalid window length!\n");
        exit(1);
    }
}

Polisher::~Polisher() { rp_reads_destroy(store_); }

const rp_reads* Polisher::read_store() {
    if (store_ == nullptr) {
        std::vector<const char*> data(sequences_.size()), quality(sequences_.size());
        std::vector<uint32_t> length(sequences_.size());
        for (size_t i = 0; i < sequences_.size(); ++i) {
            data[i] = sequences_[i].data;
            quality[i] = sequences_[i].quality;
            length[i] = sequences_[i].length;
        }
        const rp_status s = rp_reads_create(&store_, static_cast<int>(device_), static_cast<uint32_t>(data.size()),
                                            data.data(), quality.data(), length.data());
        if (s != RP_OK) {
            fprintf(stderr, "[racon_b200::Polisher] error: read store: %s (%s)\n", rp_strerror(s), rp_last_error());
            exit(1);
        }
    }
    return store_;
}

const char* Polisher::reverse_complement(uint32_t id) {  // Sequence::create_reverse_complement, sequence.cpp:58-93
    std::string& r = reverse_complement_[id];
    if (r.empty() && sequences_[id].length) {
        const SequenceView& s = sequences_[id];
        r.resize(s.length);
        for (uint32_t i = 0; i < s.length; ++i) {
            char c = s.data[s.length - 1 - i];
            switch (c) {
                case 'A': c = 'T'; break;
                case 'C': c = 'G'; break;
                case 'G': c = 'C'; break;
                case 'T': c = 'A'; break;
                default: break;
            }
            r[i] = c;
        }
    }
    return r.data();
}

const char* Polisher::reverse_quality(uint32_t id) {
    std::string& r = reverse_quality_[id];
    const SequenceView& s = sequences_[id];
    if (r.empty() && s.quality && s.length) r.assign(std::reverse_iterator<const char*>(s.quality + s.length),
                                                     std::reverse_iterator<const char*>(s.quality));
    return s.quality ? r.data() : nullptr;
}

void Polisher::find_overlap_breaking_points(std::vector<Overlap>& overlaps) {
    /* an overlap that came with an alignment (SAM input) keeps it, as in Overlap::find_breaking_points
     * (overlap.cpp:186-203); only the others are aligned, in batches, on the device */
    std::vector<size_t> todo;
    todo.reserve(overlaps.size());
    uint32_t longest = 1;
    for (size_t k = 0; k < overlaps.size(); ++k) {
        Overlap& o = overlaps[k];
        if (!o.breaking_points_.empty()) continue;
        if (!o.cigar.empty()) {
            breaking_points_from_cigar(o, window_length_);
            std::string().swap(o.cigar);
        } else {
            todo.push_back(k);
            longest = std::max(longest, std::max(o.q_end - o.q_begin, o.t_end - o.t_begin));
        }
    }
    if (todo.empty()) return;
    rp_aln* aln = nullptr;
    rp_status s = rp_aln_create(&aln, static_cast<int>(device_), 0, longest);
    if (s == RP_OK) s = rp_aln_set_window_length(aln, window_length_);
    if (s != RP_OK) {
        fprintf(stderr, "[racon_b200::Polisher::find_overlap_breaking_points] error: %s (%s)\n", rp_strerror(s),
                rp_last_error());
        exit(1);
    }
    size_t i = 0;
    while (i < todo.size()) {
        const size_t first = i;
        for (; i < todo.size(); ++i) {
            const Overlap& o = overlaps[todo[i]];
            /* the spans racon hands to the aligner (overlap.cpp:193-197) */
            const uint32_t q_start = o.strand ? o.q_length - o.q_end : o.q_begin;
            if (resident_reads_) {
                s = rp_aln_add_overlap_ref(aln, read_store(), o.q_id, q_start, o.q_end - o.q_begin, o.strand != 0, o.t_id,
                                           o.t_begin, o.t_end - o.t_begin);
            } else {
                const char* q = (o.strand ? reverse_complement(o.q_id) : sequences_[o.q_id].data) + q_start;
                const char* t = sequences_[o.t_id].data + o.t_begin;
                s = rp_aln_add_overlap(aln, q, o.q_end - o.q_begin, t, o.t_end - o.t_begin, o.t_begin, q_start);
            }
            if (s == RP_BATCH_FULL) break;
            if (s != RP_OK) {
                fprintf(stderr, "[racon_b200::Polisher::find_overlap_breaking_points] error: %s (%s)\n",
                        rp_strerror(s), rp_last_error());
                exit(1);
            }
        }
        if (i == first) {
            fprintf(stderr, "[racon_b200::Polisher::find_overlap_breaking_points] error: overlap does not fit an empty batch\n");
            exit(1);
        }
        s = rp_aln_run(aln);
        if (s == RP_OK) s = rp_aln_sync(aln);
        if (s != RP_OK) {
            fprintf(stderr, "[racon_b200::Polisher::find_overlap_breaking_points] error: %s (%s)\n", rp_strerror(s),
                    rp_last_error());
            exit(1);
        }
        for (size_t j = first; j < i; ++j) {
            const size_t k = todo[j];
            const uint32_t* pts = nullptr;
            uint32_t n = 0, st = 0;
            rp_aln_fetch_cigar(aln, static_cast<uint32_t>(j - first), nullptr, nullptr, nullptr, &st);
            if (st != RP_ALN_OK) {
                /* the reference re-aligns such overlaps with its CPU edlib (cudapolisher.cpp:213); this path has no CPU
                 * aligner, so the overlap contributes no layers and the caller is told */
                fprintf(stderr, "[racon_b200::Polisher] warning: overlap %zu exceeded device limit %u\n", k, st);
                failed_overlaps_.push_back(k);
                continue;
            }
            rp_aln_fetch_breaking_points(aln, static_cast<uint32_t>(j - first), &pts, &n);
            overlaps[k].breaking_points_.clear();
            for (uint32_t b = 0; b < n; ++b) overlaps[k].breaking_points_.emplace_back(pts[2 * b], pts[2 * b + 1]);
        }
        rp_aln_reset(aln);
    }
    rp_aln_destroy(aln);
}

void Polisher::initialize(std::vector<Overlap>& overlaps) {
    find_overlap_breaking_points(overlaps);
    build_windows(overlaps);
}

void Polisher::build_windows(const std::vector<Overlap>& overlaps) {

    /* one window per window_length_ bases of every target (polisher.cpp:383-401) */
    std::vector<uint64_t> id_to_first_window_id(targets_size_ + 1, 0);
    for (uint64_t i = 0; i < targets_size_; ++i) {
        uint32_t k = 0;
        const SequenceView& tgt = sequences_[i];
        for (uint32_t j = 0; j < tgt.length; j += window_length_, ++k) {
            const uint32_t length = std::min(j + window_length_, tgt.length) - j;
            windows_.emplace_back(createWindow(i, k, window_type_, tgt.data + j, length,
                                               tgt.quality ? tgt.quality + j : dummy_quality_.data(), length));
            windows_.back()->set_backbone_origin(static_cast<uint32_t>(i), j);
        }
        id_to_first_window_id[i + 1] = id_to_first_window_id[i] + k;
    }

    targets_coverages_.assign(targets_size_, 0);

    /* one layer per pair of breaking points (polisher.cpp:405-458) */
    for (uint64_t i = 0; i < overlaps.size(); ++i) {
        const Overlap& o = overlaps[i];
        ++targets_coverages_[o.t_id];
        const SequenceView& sequence = sequences_[o.q_id];
        const auto& breaking_points = o.breaking_points();
        for (uint32_t j = 0; j + 1 < breaking_points.size(); j += 2) {
            if (breaking_points[j + 1].second - breaking_points[j].second < 0.02 * window_length_) continue;

            const char* quality = o.strand ? reverse_quality(o.q_id) : sequence.quality;
            if (quality != nullptr) {
                double average_quality = 0;
                for (uint32_t k = breaking_points[j].second; k < breaking_points[j + 1].second; ++k) {
                    average_quality += static_cast<uint32_t>(quality[k]) - 33;
                }
                average_quality /= breaking_points[j + 1].second - breaking_points[j].second;
                if (average_quality < quality_threshold_) continue;
            }

            const uint64_t window_id = id_to_first_window_id[o.t_id] + breaking_points[j].first / window_length_;
            const uint32_t window_start = (breaking_points[j].first / window_length_) * window_length_;
            const char* data = (o.strand ? reverse_complement(o.q_id) : sequence.data) + breaking_points[j].second;
            const uint32_t data_length = breaking_points[j + 1].second - breaking_points[j].second;
            const char* layer_quality = quality == nullptr ? nullptr : quality + breaking_points[j].second;
            const uint32_t quality_length = quality == nullptr ? 0 : data_length;

            windows_[window_id]->add_layer(data, data_length, layer_quality, quality_length,
                                           breaking_points[j].first - window_start,
                                           breaking_points[j + 1].first - window_start - 1,
                                           Window::Origin{o.q_id, breaking_points[j].second, o.strand != 0});
        }
    }
}

void Polisher::polish(std::vector<PolishedSequence>& dst, bool drop_unpolished_sequences) {
    auto batch = createBatch(0, device_, 0, gap_, mismatch_, match_, false, window_length_, trim_);
    std::vector<bool> polished(windows_.size(), false);
    size_t i = 0;
    while (i < windows_.size()) {  // cudapolisher.cpp:254-276: fill, run, collect, reset
        const size_t first = i;
        while (i < windows_.size() && batch->addWindow(windows_[i])) ++i;
        if (i == first) {
            fprintf(stderr, "[racon_b200::Polisher::polish] error: window does not fit an empty batch\n");
            exit(1);
        }
        const std::vector<bool>& flags = batch->generateConsensus();
        for (size_t k = 0; k < flags.size(); ++k) polished[first + k] = flags[k];
        for (uint32_t k : batch->failedWindows()) failed_windows_.push_back(first + k);
        batch->reset();
    }

    /* stitch (polisher.cpp:504-537) */
    std::string polished_data;
    uint32_t num_polished_windows = 0;
    for (uint64_t w = 0; w < windows_.size(); ++w) {
        num_polished_windows += polished[w] ? 1 : 0;
        polished_data += windows_[w]->consensus();
        if (w == windows_.size() - 1 || windows_[w + 1]->rank() == 0) {
            const double polished_ratio = num_polished_windows / static_cast<double>(windows_[w]->rank() + 1);
            if (!drop_unpolished_sequences || polished_ratio > 0) {
                std::string tags = fragment_correction_ ? "r" : "";
                tags += " LN:i:" + std::to_string(polished_data.size());
                tags += " RC:i:" + std::to_string(targets_coverages_[windows_[w]->id()]);
                tags += " XC:f:" + std::to_string(polished_ratio);
                dst.push_back(PolishedSequence{windows_[w]->id(), tags, polished_data});
            }
            num_polished_windows = 0;
            polished_data.clear();
        }
    }
}

void Polisher::polish_streaming(const std::function<void(const PolishedSequence&)>& sink,
                                bool drop_unpolished_sequences, size_t memory, bool banded, bool resident_reads) {
    constexpr int kObjects = 2;
    struct InFlight {
        std::unique_ptr<BatchProcessor> batch;
        size_t first = 0, count = 0;
        bool busy = false;
    };
    /* resident reads: every sequence goes to the device once; the windows (whose pieces know their origin, build_windows)
     * are then added by reference and their layers extracted on the device */
    const rp_reads* store = (resident_reads || resident_reads_) ? read_store() : nullptr;
    InFlight obj[kObjects];
    for (auto& o : obj) {
        o.batch = createBatch(0, device_, memory, gap_, mismatch_, match_, banded, window_length_, trim_);
        o.batch->useReadStore(store);
    }
    std::string polished_data;
    uint32_t num_polished_windows = 0;
    auto stitch = [&](size_t w, bool polished) {  // polisher.cpp:504-530, one window at a time, in window order
        num_polished_windows += polished ? 1 : 0;
        polished_data += windows_[w]->consensus();
        if (w == windows_.size() - 1 || windows_[w + 1]->rank() == 0) {
            const double polished_ratio = num_polished_windows / static_cast<double>(windows_[w]->rank() + 1);
            if (!drop_unpolished_sequences || polished_ratio > 0) {
                std::string tags = fragment_correction_ ? "r" : "";
                tags += " LN:i:" + std::to_string(polished_data.size());
                tags += " RC:i:" + std::to_string(targets_coverages_[windows_[w]->id()]);
                tags += " XC:f:" + std::to_string(polished_ratio);
                sink(PolishedSequence{windows_[w]->id(), tags, polished_data});
            }
            num_polished_windows = 0;
            polished_data.clear();
        }
        windows_[w].reset();  // polisher.cpp:531: the window is not needed any more
    };
    size_t next = 0;
    int fill = 0, drain = 0;  // objects are filled and drained in the same round-robin order => window order
    while (next < windows_.size() || obj[drain].busy) {
        InFlight& f = obj[fill];
        if (!f.busy && next < windows_.size()) {
            f.batch->reset();
            f.first = next;
            while (next < windows_.size() && f.batch->addWindow(windows_[next])) ++next;
            f.count = next - f.first;
            if (f.count == 0) {
                fprintf(stderr, "[racon_b200::Polisher::polish_streaming] error: window does not fit an empty batch\n");
                exit(1);
            }
            f.batch->launch();  // asynchronous: the host goes on filling the other object / stitching
            f.busy = true;
            fill = (fill + 1) % kObjects;
            if (!obj[fill].busy && next < windows_.size()) continue;  // keep both in flight before draining
        }
        InFlight& d = obj[drain];
        if (d.busy) {
            const std::vector<bool>& flags = d.batch->collect();
            for (uint32_t k : d.batch->failedWindows()) failed_windows_.push_back(d.first + k);
            for (size_t k = 0; k < d.count; ++k) stitch(d.first + k, flags[k]);
            d.busy = false;
            drain = (drain + 1) % kObjects;
        }
    }
}

}  // namespace racon_b200

/* Test hooks with the shape of oracle/ref_polisher_harness.cpp (open -> counts -> export -> polish -> polished), so
 * that tests/test_pipeline.py compares this pipeline and the unmodified reference Polisher field by field. */
namespace {
struct PolHandle {
    std::unique_ptr<racon_b200::InputSet> input;   // owns the sequence bytes when the polisher was opened on files
    std::unique_ptr<racon_b200::Polisher> polisher;
    std::vector<racon_b200::PolishedSequence> polished;
};
}  // namespace

extern "C" void* rp_mirror_polisher_open(uint32_t n_seq, const char* bases, const char* quals, const uint64_t* seq_off,
                                         const uint8_t* seq_has_qual, uint32_t n_targets, int window_type_tgs,
                                         int fragment_correction, uint32_t n_overlaps, const uint32_t* overlaps,
                                         uint32_t window_length, double quality_threshold, int trim, int8_t match,
                                         int8_t mismatch, int8_t gap, uint32_t device) {
    using namespace racon_b200;
    std::vector<SequenceView> seqs(n_seq);
    for (uint32_t i = 0; i < n_seq; ++i) {
        seqs[i].data = bases + seq_off[i];
        seqs[i].quality = seq_has_qual[i] ? quals + seq_off[i] : nullptr;
        seqs[i].length = static_cast<uint32_t>(seq_off[i + 1] - seq_off[i]);
    }
    std::vector<Overlap> ovl(n_overlaps);
    for (uint32_t i = 0; i < n_overlaps; ++i) {
        const uint32_t* o = overlaps + 9ull * i;
        ovl[i] = Overlap{o[0], o[1], o[2], o[3], o[4], o[5], o[6], o[7], o[8], {}, {}};
    }
    PolHandle* h = new PolHandle();
    h->polisher.reset(new Polisher(std::move(seqs), n_targets, window_type_tgs ? WindowType::kTGS : WindowType::kNGS,
                                   fragment_correction != 0, window_length, quality_threshold, trim != 0, match,
                                   mismatch, gap, device));
    h->polisher->initialize(ovl);
    return h;
}

/* Host logic only (no device call): the same, with the breaking points given (bp_off in points, bp = (t, q) pairs). */
extern "C" void* rp_mirror_polisher_open_with_bp(uint32_t n_seq, const char* bases, const char* quals,
                                                 const uint64_t* seq_off, const uint8_t* seq_has_qual,
                                                 uint32_t n_targets, int window_type_tgs, int fragment_correction,
                                                 uint32_t n_overlaps, const uint32_t* overlaps, const uint64_t* bp_off,
                                                 const uint32_t* bp, uint32_t window_length, double quality_threshold) {
    using namespace racon_b200;
    std::vector<SequenceView> seqs(n_seq);
    for (uint32_t i = 0; i < n_seq; ++i) {
        seqs[i].data = bases + seq_off[i];
        seqs[i].quality = seq_has_qual[i] ? quals + seq_off[i] : nullptr;
        seqs[i].length = static_cast<uint32_t>(seq_off[i + 1] - seq_off[i]);
    }
    std::vector<Overlap> ovl(n_overlaps);
    for (uint32_t i = 0; i < n_overlaps; ++i) {
        const uint32_t* o = overlaps + 9ull * i;
        ovl[i] = Overlap{o[0], o[1], o[2], o[3], o[4], o[5], o[6], o[7], o[8], {}, {}};
        for (uint64_t b = bp_off[i]; b < bp_off[i + 1]; ++b) ovl[i].breaking_points_.emplace_back(bp[2 * b], bp[2 * b + 1]);
    }
    PolHandle* h = new PolHandle();
    h->polisher.reset(new Polisher(std::move(seqs), n_targets, window_type_tgs ? WindowType::kTGS : WindowType::kNGS,
                                   fragment_correction != 0, window_length, quality_threshold, true, 3, -5, -4, 0));
    h->polisher->build_windows(ovl);
    return h;
}

/* Files in: createPolisher + initialize (polisher.cpp:55-160,200-461) — reads_io's load_input, then alignment and breaking
 * points on the device (overlaps that carry a CIGAR — SAM input — keep it and need no device) and the window building.
 * NULL on an input error (message in err).  rp_mirror_polisher_stream_fasta(names = NULL) then is racon's main loop. */
extern "C" void* rp_mirror_polisher_open_files(const char* reads, const char* overlaps, const char* targets,
                                               int fragment_correction, uint32_t window_length, double quality_threshold,
                                               double error_threshold, int trim, int8_t match, int8_t mismatch, int8_t gap,
                                               uint32_t device, int resident_reads, char* err, uint32_t err_cap) {
    using namespace racon_b200;
    try {
        std::unique_ptr<PolHandle> h(new PolHandle());
        h->input.reset(new InputSet(load_input(reads, overlaps, targets, fragment_correction != 0, error_threshold)));
        h->polisher.reset(new Polisher(h->input->views(), h->input->targets_size, h->input->window_type,
                                       fragment_correction != 0, window_length, quality_threshold, trim != 0, match,
                                       mismatch, gap, device));
        h->polisher->use_resident_reads(resident_reads != 0);
        h->polisher->initialize(h->input->overlaps);
        return h.release();
    } catch (const std::exception& e) {
        if (err && err_cap) {
            std::strncpy(err, e.what(), err_cap - 1);
            err[err_cap - 1] = '\0';
        }
        return nullptr;
    }
}

/* failed[0] = overlaps, failed[1] = windows the device could not finish (Polisher::failed_overlaps/failed_windows) */
extern "C" void rp_mirror_polisher_failed(void* hv, uint64_t* failed) {
    PolHandle* h = static_cast<PolHandle*>(hv);
    failed[0] = h->polisher->failed_overlaps().size();
    failed[1] = h->polisher->failed_windows().size();
}

extern "C" void rp_mirror_polisher_counts(void* hv, uint64_t* counts) {
    PolHandle* h = static_cast<PolHandle*>(hv);
    counts[0] = h->polisher->windows().size();
    counts[1] = counts[2] = 0;
    for (const auto& w : h->polisher->windows()) {
        counts[1] += w->sequences().size();
        for (const auto& s : w->sequences()) counts[2] += s.second;
    }
}

extern "C" void rp_mirror_polisher_export(void* hv, char* bases, char* quals, uint64_t* seq_off, uint8_t* seq_has_qual,
                                          uint32_t* seq_begin, uint32_t* seq_end, uint32_t* win_first, uint8_t* win_type,
                                          uint64_t* win_target, uint32_t* win_rank) {
    PolHandle* h = static_cast<PolHandle*>(hv);
    uint64_t nb = 0, ns = 0;
    seq_off[0] = 0;
    const auto& windows = h->polisher->windows();
    for (size_t w = 0; w < windows.size(); ++w) {
        const auto& win = windows[w];
        win_first[w] = static_cast<uint32_t>(ns);
        win_type[w] = win->type() == racon_b200::WindowType::kTGS ? 1 : 0;
        win_target[w] = win->id();
        win_rank[w] = win->rank();
        for (size_t s = 0; s < win->sequences().size(); ++s) {
            const uint32_t len = win->sequences()[s].second;
            std::memcpy(bases + nb, win->sequences()[s].first, len);
            const char* q = win->qualities()[s].first;
            if (q) {
                std::memcpy(quals + nb, q, len);
                seq_has_qual[ns] = 1;
            } else {
                std::memset(quals + nb, '!', len);
                seq_has_qual[ns] = 0;
            }
            seq_begin[ns] = win->positions()[s].first;
            seq_end[ns] = win->positions()[s].second;
            nb += len;
            seq_off[++ns] = nb;
        }
    }
    win_first[windows.size()] = static_cast<uint32_t>(ns);
}

extern "C" uint32_t rp_mirror_polisher_polish(void* hv, int drop_unpolished) {
    PolHandle* h = static_cast<PolHandle*>(hv);
    h->polisher->polish(h->polished, drop_unpolished != 0);
    return static_cast<uint32_t>(h->polished.size());
}

extern "C" uint32_t rp_mirror_polisher_window_consensus(void* hv, uint32_t w, char* out, uint32_t cap) {
    PolHandle* h = static_cast<PolHandle*>(hv);
    const std::string& c = h->polisher->windows()[w]->consensus();
    if (c.size() > cap) return 0xffffffffu;
    std::memcpy(out, c.data(), c.size());
    return static_cast<uint32_t>(c.size());
}

extern "C" uint64_t rp_mirror_polisher_polished(void* hv, uint32_t i, uint64_t* target_id, char* tags, uint32_t tags_cap,
                                                char* data, uint64_t data_cap) {
    PolHandle* h = static_cast<PolHandle*>(hv);
    const auto& s = h->polished[i];
    *target_id = s.id;
    std::strncpy(tags, s.tags.c_str(), tags_cap - 1);
    tags[tags_cap - 1] = 0;
    if (s.data.size() > data_cap) return ~0ull;
    std::memcpy(data, s.data.data(), s.data.size());
    return s.data.size();
}

/* Streaming polish straight into a FASTA file (names: NUL-separated target names, one per target id; NULL = the names
 * of the files the polisher was opened on); returns the
 * number of records written.  mem_bytes: budget per batch object (small budgets force many batches). */
extern "C" uint32_t rp_mirror_polisher_stream_fasta(void* hv, int drop_unpolished, const char* path, const char* names,
                                                     uint64_t mem_bytes, int banded, int resident_reads) {
    PolHandle* h = static_cast<PolHandle*>(hv);
    std::vector<std::string> name_of;
    for (const char* p = names; p && *p; p += std::strlen(p) + 1) name_of.emplace_back(p);
    if (!names && h->input) {
        for (uint64_t i = 0; i < h->input->targets_size; ++i) name_of.push_back(h->input->sequences[i].name);
    }
    FILE* f = std::fopen(path, "wb");
    if (!f) return 0;
    uint32_t n = 0;
    h->polisher->polish_streaming(
        [&](const racon_b200::PolishedSequence& s) {
            const std::string rec = racon_b200::format_fasta(s.id < name_of.size() ? name_of[s.id] : "?", s);
            std::fwrite(rec.data(), 1, rec.size(), f);
            ++n;
        },
        drop_unpolished != 0, static_cast<size_t>(mem_bytes), banded != 0, resident_reads != 0);
    std::fclose(f);
    return n;
}

extern "C" void rp_mirror_polisher_close(void* hv) { delete static_cast<PolHandle*>(hv); }

/* host-only: one FASTA record; returns its length (out may be NULL to size it) */
extern "C" uint64_t rp_mirror_format_fasta(const char* name, const char* tags, const char* data, uint64_t data_len,
                                           char* out, uint64_t cap) {
    racon_b200::PolishedSequence s{0, tags, std::string(data, data_len)};
    const std::string r = racon_b200::format_fasta(name, s);
    if (out && r.size() <= cap) std::memcpy(out, r.data(), r.size());
    return r.size();
}

/* Test hook: CUDAPolisher::find_overlap_breaking_points' batch loop (cudapolisher.cpp:100-213) over flat pairs:
 * fill a batch until addOverlap refuses, alignAll, generate_cigar_strings, reset, continue.  out: NUL-terminated
 * CIGARs at out + i * stride (empty when the device could not take the overlap). */
extern "C" int rp_mirror_align(uint32_t n_pairs, const char* bases, const uint64_t* q_off, const uint32_t* q_len,
                               const uint64_t* t_off, const uint32_t* t_len, uint32_t max_alignments, uint32_t device,
                               char* out, uint32_t stride) {
    using namespace racon_b200;
    uint32_t mq = 1, mt = 1;
    for (uint32_t i = 0; i < n_pairs; ++i) {
        if (q_len[i] > mq) mq = q_len[i];
        if (t_len[i] > mt) mt = t_len[i];
    }
    auto batch = createBatchAligner(mq, mt, max_alignments, device);
    std::vector<std::string> cigars(n_pairs);
    uint32_t i = 0;
    while (i < n_pairs) {
        uint32_t first = i;
        while (i < n_pairs && batch->addOverlap(bases + q_off[i], q_len[i], bases + t_off[i], t_len[i], &cigars[i])) ++i;
        if (i == first) return -1;
        batch->alignAll();
        batch->generate_cigar_strings();
        batch->reset();
    }
    for (uint32_t k = 0; k < n_pairs; ++k) {
        if (cigars[k].size() + 1 > stride) return -2;
        std::memcpy(out + static_cast<uint64_t>(k) * stride, cigars[k].c_str(), cigars[k].size() + 1);
    }
    return 0;
}

/* Test hook: drives the mirror classes exactly like CUDAPolisher::polish drives CUDABatchProcessor
 * (cudapolisher.cpp:254-276), over a flat window set. Same contract as oracle/ref_harness.cpp. */
extern "C" double rp_mirror_consensus(uint32_t n_windows, const char* bases, const char* quals,
                                      const uint64_t* seq_off, const uint8_t* seq_has_qual, const uint32_t* seq_begin,
                                      const uint32_t* seq_end, const uint32_t* win_first, const uint8_t* win_type,
                                      int8_t match, int8_t mismatch, int8_t gap, uint32_t window_length, int trim,
                                      uint32_t device, char* out, uint32_t out_stride, uint32_t* out_len,
                                      uint8_t* polished) {
    using namespace racon_b200;
    uint32_t max_bl = 1;
    for (uint32_t w = 0; w < n_windows; ++w) {
        uint32_t s0 = win_first[w];
        uint32_t bl = static_cast<uint32_t>(seq_off[s0 + 1] - seq_off[s0]);
        if (bl > max_bl) max_bl = bl;
    }
    std::string dummy(max_bl, '!');
    std::vector<std::shared_ptr<Window>> windows;
    for (uint32_t w = 0; w < n_windows; ++w) {
        uint32_t s0 = win_first[w], s1 = win_first[w + 1];
        uint64_t o = seq_off[s0];
        uint32_t bl = static_cast<uint32_t>(seq_off[s0 + 1] - o);
        bool bq = quals && seq_has_qual && seq_has_qual[s0];
        auto win = createWindow(w, w, win_type[w] ? WindowType::kTGS : WindowType::kNGS, bases + o, bl,
                                bq ? quals + o : dummy.data(), bl);
        for (uint32_t s = s0 + 1; s < s1; ++s) {
            uint64_t so = seq_off[s];
            uint32_t sl = static_cast<uint32_t>(seq_off[s + 1] - so);
            bool q = quals && seq_has_qual && seq_has_qual[s];
            win->add_layer(bases + so, sl, q ? quals + so : nullptr, q ? sl : 0, seq_begin[s], seq_end[s]);
        }
        windows.push_back(win);
    }
    auto batch = createBatch(0, device, 0, gap, mismatch, match, false, window_length, trim != 0);
    std::vector<uint8_t> flags(n_windows, 0);
    uint32_t next = 0;
    while (next < n_windows) {
        uint32_t first = next;
        batch->reset();
        while (next < n_windows && batch->addWindow(windows[next])) ++next;
        if (next == first) return -1.0;  // a single window does not fit
        const std::vector<bool>& st = batch->generateConsensus();
        for (uint32_t k = 0; k < st.size(); ++k) flags[first + k] = st[k] ? 1 : 0;
    }
    for (uint32_t w = 0; w < n_windows; ++w) {
        const std::string& c = windows[w]->consensus();
        if (c.size() > out_stride) return -2.0;
        std::memcpy(out + static_cast<uint64_t>(w) * out_stride, c.data(), c.size());
        out_len[w] = static_cast<uint32_t>(c.size());
        polished[w] = flags[w];
    }
    return 0.0;
}
