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
#include <functional>
#include <memory>
#include <string>
#include <utility>
#include <vector>

struct rp_poa;
struct rp_aln;
struct rp_reads;

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
    WindowType type() const { return type_; }
    const std::vector<std::pair<const char*, uint32_t>>& sequences() const { return sequences_; }
    const std::vector<std::pair<const char*, uint32_t>>& qualities() const { return qualities_; }
    const std::vector<std::pair<uint32_t, uint32_t>>& positions() const { return positions_; }

    /* same checks, same silent skips and the same fatal errors as the reference (window.cpp:42-63) */
    void add_layer(const char* sequence, uint32_t sequence_length, const char* quality, uint32_t quality_length,
                   uint32_t begin, uint32_t end);

    /* Where a piece comes from: sequence `seq_id` of the caller's sequence table from `offset` on — of its reverse
     * complement when `reverse`.  A window whose backbone and every layer carry their origin can be handed to a batch by
     * reference into a device-resident read store (BatchProcessor::useReadStore) instead of by its bytes. */
    struct Origin {
        uint32_t seq_id, offset;
        bool reverse;
    };
    void set_backbone_origin(uint32_t seq_id, uint32_t offset) { origins_.assign(1, Origin{seq_id, offset, false}); }
    /* add_layer + the layer's origin (the origin is dropped with the layer when add_layer skips it) */
    void add_layer(const char* sequence, uint32_t sequence_length, const char* quality, uint32_t quality_length,
                   uint32_t begin, uint32_t end, const Origin& origin);
    const std::vector<Origin>& origins() const { return origins_; }

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
    std::vector<Origin> origins_;   // empty, or one per entry of sequences_
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
    /* the two halves of generateConsensus (generatePOA / getConsensus of the reference, cudabatch.cpp:193-270):
     * launch() queues H2D + kernel + D2H on the object's stream and returns, collect() waits and reads back */
    void launch();
    const std::vector<bool>& collect();
    void reset();
    uint32_t getBatchID() const { return bid_; }
    /* From now on windows that know the origin of all their pieces (Window::origins) are added by reference into
     * `reads` (rp_reads_create over the same sequence table, same device): no sequence byte is copied on the host, the
     * layers are extracted on the device.  nullptr switches back.  Call between batches (after reset()). */
    void useReadStore(const rp_reads* reads) { reads_ = reads; }
    /* indices (within the batch) of the windows of the last generateConsensus() that hit a device limit (soft RP_WIN_*
     * status): their flag is false and their consensus is the backbone — not what racon would produce; the reference's
     * caller re-runs such windows on the CPU (cudapolisher.cpp:354-370), a caller without a CPU path must treat them as
     * an error */
    const std::vector<uint32_t>& failedWindows() const { return failed_; }

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
    const rp_reads* reads_ = nullptr;
    bool trim_ = true;
    std::vector<std::shared_ptr<Window>> windows_;
    std::vector<bool> window_consensus_status_;
    std::vector<uint32_t> failed_;
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

/* ------------------------------------------------------------------------------------------------------------------
 * Mirror of the device-facing half of racon::CUDAPolisher: overlaps -> (device) alignment + breaking points ->
 * windows -> (device) consensus -> stitched sequences.  Parsing and overlap filtering are reads_io.hpp's (load_input);
 * everything that touches sequence bytes in bulk runs on the GPU.
 *   find_overlap_breaking_points : src/cuda/cudapolisher.cpp:72-213  (batched aligner loop)
 *                                  + src/overlap.cpp:226-292          (breaking points, on the device here)
 *   initialize (window building) : src/polisher.cpp:383-461
 *   polish                       : src/cuda/cudapolisher.cpp:215-395 (batch loop), src/polisher.cpp:504-537 (stitch)
 * ------------------------------------------------------------------------------------------------------------------ */
struct SequenceView {        // racon::Sequence as the polisher needs it; the bytes stay owned by the caller
    const char* data;
    const char* quality;     // nullptr when the sequence has no quality string
    uint32_t length;
};

struct Overlap {             // racon::Overlap after transmute() (src/overlap.hpp:82-98)
    uint32_t q_id, t_id, strand, q_begin, q_end, q_length, t_begin, t_end, t_length;
    std::string cigar;       // only when the overlap file carried an alignment (SAM): then it is used as it is
    std::vector<std::pair<uint32_t, uint32_t>> breaking_points_;
    const std::vector<std::pair<uint32_t, uint32_t>>& breaking_points() const { return breaking_points_; }
};

/* Breaking points of an overlap that arrived WITH an alignment (SAM): the (target, query) coordinates of the first and
 * one-past-the-last aligned base pair inside every window of the target the alignment crosses — what
 * Overlap::find_breaking_points_from_cigar (overlap.cpp:226-292) finds base by base, computed run by run here.
 * (Overlaps without a CIGAR get theirs from the device: rp_aln_fetch_breaking_points.) */
void breaking_points_from_cigar(Overlap& overlap, uint32_t window_length);

struct PolishedSequence {
    uint64_t id;             // target index
    std::string tags;        // " LN:i:.. RC:i:.. XC:f:.." exactly as polisher.cpp:521-524 appends them to the name
    std::string data;
};

/* One FASTA record as racon's main prints a polished sequence (src/main.cpp:159-161): ">name tags\ndata\n". */
std::string format_fasta(const std::string& target_name, const PolishedSequence& s);

class Polisher {
public:
    /* sequences: the targets first (targets_size of them), then the reads — the order racon::Polisher::sequences_ has */
    Polisher(std::vector<SequenceView> sequences, uint64_t targets_size, WindowType window_type, bool fragment_correction,
             uint32_t window_length, double quality_threshold, bool trim, int8_t match, int8_t mismatch, int8_t gap,
             uint32_t device);
    ~Polisher();
    /* Device-resident reads: every sequence is uploaded once (rp_reads_create, at first use) and both device stages name
     * their inputs instead of copying them — the aligner its two spans per overlap (rp_aln_add_overlap_ref), the consensus
     * its window pieces (rp_poa_add_window_refs).  Call before initialize(). */
    void use_resident_reads(bool on) { resident_reads_ = on; }
    void find_overlap_breaking_points(std::vector<Overlap>& overlaps);
    void initialize(std::vector<Overlap>& overlaps);
    /* the window-building half of initialize() alone, for overlaps whose breaking points are already known */
    void build_windows(const std::vector<Overlap>& overlaps);
    void polish(std::vector<PolishedSequence>& dst, bool drop_unpolished_sequences);
    /* Streaming form (polisher.cpp:504-537 + main.cpp:159-161 as a consumer overlapped with the compute): two batch
     * objects stay in flight; while one runs on the GPU the other's finished windows are stitched in window order, and
     * `sink` gets every polished sequence the moment its last window is collected (e.g. a FASTA writer).  `memory` =
     * device budget per batch object (0 = library default), `banded` = racon -b.  `resident_reads`: upload every sequence
     * once (rp_reads) and add the windows by reference — layer extraction on the device, no host copy of layer bytes. */
    void polish_streaming(const std::function<void(const PolishedSequence&)>& sink, bool drop_unpolished_sequences,
                          size_t memory = 0, bool banded = false, bool resident_reads = false);
    const std::vector<std::shared_ptr<Window>>& windows() const { return windows_; }
    /* Items the device could not finish (soft RP_ALN_* / RP_WIN_* status).  The reference hands such items to its CPU
     * code (cudapolisher.cpp:213 edlib, :354-370 spoa); this library has no CPU path, so a failed overlap contributes
     * no layers and a failed window keeps its backbone — the result then DIFFERS from racon's, which is why the indices
     * are reported here instead of a stderr line: a caller with a CPU implementation (the integration shim's host is
     * the reference itself) must process them, any other caller should treat a non-empty list as an error. */
    const std::vector<size_t>& failed_overlaps() const { return failed_overlaps_; }
    const std::vector<size_t>& failed_windows() const { return failed_windows_; }

private:
    Polisher(const Polisher&) = delete;
    const Polisher& operator=(const Polisher&) = delete;
    const char* reverse_complement(uint32_t id);
    const char* reverse_quality(uint32_t id);

    std::vector<SequenceView> sequences_;
    uint64_t targets_size_;
    WindowType window_type_;
    bool fragment_correction_;
    uint32_t window_length_;
    double quality_threshold_;
    bool trim_;
    int8_t match_, mismatch_, gap_;
    uint32_t device_;
    std::string dummy_quality_;
    std::vector<std::string> reverse_complement_, reverse_quality_;
    std::vector<uint32_t> targets_coverages_;
    std::vector<std::shared_ptr<Window>> windows_;
    std::vector<size_t> failed_overlaps_, failed_windows_;
    bool resident_reads_ = false;
    rp_reads* store_ = nullptr;
    const rp_reads* read_store();
};

}  // namespace racon_b200
