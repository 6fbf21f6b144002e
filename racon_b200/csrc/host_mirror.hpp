/*
 * host_mirror.hpp — C++ host-side mirror of the reference's interface for the POA hot path, written
 * over the C ABI (include/racon_b200.h).  Same names, argument meaning and error behaviour as
 *   racon::Window / createWindow / Window::add_layer        (/root/reference/src/window.hpp:25-74, window.cpp:15-63)
 *   racon::CUDABatchProcessor / createCUDABatch             (/root/reference/src/cuda/cudabatch.hpp:27-122)
 * so that racon's CUDAPolisher (src/cuda/cudapolisher.cpp:216-413) can drive it unchanged.
 */
#pragma once
#include <stdint.h>

#include <atomic>
#include <memory>
#include <string>
#include <utility>
#include <vector>

struct rp_poa;
struct rp_aln;

namespace racon_b200 {

enum class WindowType { kNGS, kTGS };  // src/window.hpp:20-23

class Window;
std::shared_ptr<Window> createWindow(uint64_t id, uint32_t rank, WindowType type, const char* backbone,
                                     uint32_t backbone_length, const char* quality, uint32_t quality_length);

class Window {
public:
    ~Window();
    uint64_t id() const { return id_; }
    uint32_t rank() const { return rank_; }
    const std::string& consensus() const { return consensus_; }

    /* same checks, same silent skips and the same fatal errors as the reference (window.cpp:42-63) */
    void add_layer(const char* sequence, uint32_t sequence_length, const char* quality, uint32_t quality_length,
                   uint32_t begin, uint32_t end);

    friend std::shared_ptr<Window> createWindow(uint64_t id, uint32_t rank, WindowType type, const char* backbone,
                                                uint32_t backbone_length, const char* quality,
                                                uint32_t quality_length);
    friend class BatchProcessor;

private:
    Window(uint64_t id, uint32_t rank, WindowType type, const char* backbone, uint32_t backbone_length,
           const char* quality, uint32_t quality_length);
    Window(const Window&) = delete;
    const Window& operator=(const Window&) = delete;

    uint64_t id_;
    uint32_t rank_;
    WindowType type_;
    std::string consensus_;
    std::vector<std::pair<const char*, uint32_t>> sequences_;
    std::vector<std::pair<const char*, uint32_t>> qualities_;
    std::vector<std::pair<uint32_t, uint32_t>> positions_;
};

class BatchProcessor;
/* createCUDABatch (cudabatch.cpp:23-39) + the two things the reference fixes elsewhere: window length
 * (Polisher ctor) and the trim flag (Window::generate_consensus argument). */
std::unique_ptr<BatchProcessor> createBatch(uint32_t max_window_depth, uint32_t device, size_t avail_mem, int8_t gap,
                                            int8_t mismatch, int8_t match, bool cuda_banded_alignment,
                                            uint32_t window_length = 500, bool trim = true);

class BatchProcessor {
public:
    ~BatchProcessor();
    /* false = batch full: run generateConsensus(), reset(), add the window again (cudabatch.cpp:126-132) */
    bool addWindow(std::shared_ptr<Window> window);
    bool hasWindows() const;
    /* writes window->consensus_ for every accepted window; flag = what Window::generate_consensus returns */
    const std::vector<bool>& generateConsensus();
    void reset();
    uint32_t getBatchID() const { return bid_; }

    friend std::unique_ptr<BatchProcessor> createBatch(uint32_t, uint32_t, size_t, int8_t, int8_t, int8_t, bool,
                                                       uint32_t, bool);

private:
    BatchProcessor(uint32_t max_window_depth, uint32_t device, size_t avail_mem, int8_t gap, int8_t mismatch,
                   int8_t match, bool cuda_banded_alignment, uint32_t window_length, bool trim);
    BatchProcessor(const BatchProcessor&) = delete;
    const BatchProcessor& operator=(const BatchProcessor&) = delete;

    static std::atomic<uint32_t> batches;
    uint32_t bid_ = 0;
    rp_poa* poa_ = nullptr;
    bool trim_ = true;
    std::vector<std::shared_ptr<Window>> windows_;
    std::vector<bool> window_consensus_status_;
};

/* Mirror of racon::CUDABatchAligner (src/cuda/cudaaligner.hpp:21-92).  The reference's addOverlap takes an
 * Overlap* plus the sequence table and derives the two spans (cudaaligner.cpp:53-58); the mirror takes the spans and a
 * pointer to the std::string that plays Overlap::cigar_. */
class BatchAligner;
std::unique_ptr<BatchAligner> createBatchAligner(uint32_t max_query_size, uint32_t max_target_size,
                                                 uint32_t max_alignments, uint32_t device_id);

class BatchAligner {
public:
    ~BatchAligner();
    /* false = batch full (exceeded_max_alignments / exceeded_max_length are soft: an over-long overlap is
     * accepted and simply keeps an empty cigar, as in cudaaligner.cpp:61-77) */
    bool addOverlap(const char* q, uint32_t q_len, const char* t, uint32_t t_len, std::string* cigar);
    bool hasOverlaps() const { return !cigars_.empty(); }
    void alignAll();                  // asynchronous launch (cudaaligner.cpp:80-84)
    void generate_cigar_strings();    // sync + fill every *cigar (cudaaligner.cpp:86-104)
    void reset();
    uint32_t getBatchID() const { return bid_; }

    friend std::unique_ptr<BatchAligner> createBatchAligner(uint32_t, uint32_t, uint32_t, uint32_t);

private:
    BatchAligner(uint32_t max_query_size, uint32_t max_target_size, uint32_t max_alignments, uint32_t device_id);
    BatchAligner(const BatchAligner&) = delete;
    const BatchAligner& operator=(const BatchAligner&) = delete;

    static std::atomic<uint32_t> batches;
    uint32_t bid_ = 0;
    uint32_t max_alignments_ = 0;
    rp_aln* aln_ = nullptr;
    std::vector<std::string*> cigars_;
};

}  // namespace racon_b200
