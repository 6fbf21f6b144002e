/*
 * host_mirror.cpp — see host_mirror.hpp.  Thin: all work happens behind the C ABI on the GPU.
 */
#include "host_mirror.hpp"

#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "racon_b200.h"

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
    rp_status s = rp_poa_add_window(poa_, static_cast<uint32_t>(n), seq.data(), len.data(), qual.data(), begin.data(),
                                    end.data(), window->type_ == WindowType::kTGS ? RP_WINDOW_TGS : RP_WINDOW_NGS,
                                    trim_ ? 1 : 0);
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
    rp_status s = rp_poa_run(poa_);
    if (s == RP_OK) s = rp_poa_sync(poa_);
    if (s != RP_OK) {
        fprintf(stderr, "[racon_b200::BatchProcessor::generateConsensus] error: %s (%s)\n", rp_strerror(s),
                rp_last_error());
        exit(1);
    }
    window_consensus_status_.clear();
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

}  // namespace racon_b200

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
