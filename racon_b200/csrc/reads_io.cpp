/* reads_io.cpp — see reads_io.hpp.  Host code; not part of libracon_b200.so. */
#include "reads_io.hpp"

#include <zlib.h>

#include <cctype>
#include <cstdlib>
#include <cstring>
#include <stdexcept>
#include <unordered_map>

namespace racon_b200 {

namespace {

bool ends_with(const std::string& s, const char* suffix) {
    const size_t n = std::strlen(suffix);
    return s.size() >= n && s.compare(s.size() - n, n, suffix) == 0;
}

bool has_extension(const std::string& path, std::initializer_list<const char*> exts) {
    for (const char* e : exts) {
        if (ends_with(path, e) || ends_with(path, (std::string(e) + ".gz").c_str())) return true;
    }
    return false;
}

/* Lines of a plain or gzip-compressed file.  A line handed out stays valid until the next call. */
class GzLines {
public:
    explicit GzLines(const std::string& path) : path_(path), buf_(size_t(1) << 22) {
        file_ = gzopen(path.c_str(), "rb");
        if (!file_) throw std::runtime_error("[racon_b200::reads_io] error: unable to open file " + path);
        gzbuffer(file_, 1u << 20);
    }
    ~GzLines() {
        if (file_) gzclose(file_);
    }
    GzLines(const GzLines&) = delete;
    GzLines& operator=(const GzLines&) = delete;

    bool next(const char** line, size_t* length) {
        for (;;) {
            if (const void* nl = std::memchr(buf_.data() + pos_, '\n', end_ - pos_)) {
                *line = buf_.data() + pos_;
                *length = static_cast<const char*>(nl) - *line;
                pos_ += *length + 1;
                return true;
            }
            if (eof_) {
                if (pos_ == end_) return false;
                *line = buf_.data() + pos_;   // last line without a newline
                *length = end_ - pos_;
                pos_ = end_;
                return true;
            }
            if (pos_ > 0) {
                std::memmove(buf_.data(), buf_.data() + pos_, end_ - pos_);
                end_ -= pos_;
                pos_ = 0;
            }
            if (end_ == buf_.size()) buf_.resize(2 * buf_.size());
            const size_t room = std::min<size_t>(buf_.size() - end_, size_t(1) << 30);
            const int got = gzread(file_, buf_.data() + end_, static_cast<unsigned>(room));
            if (got < 0) throw std::runtime_error("[racon_b200::reads_io] error: unable to read file " + path_);
            if (got == 0) eof_ = true;
            end_ += static_cast<size_t>(got);
        }
    }

private:
    std::string path_;
    gzFile file_ = nullptr;
    std::vector<char> buf_;
    size_t pos_ = 0, end_ = 0;
    bool eof_ = false;
};

size_t right_strip(const char* s, size_t n) {
    while (n > 0 && std::isspace(static_cast<unsigned char>(s[n - 1]))) --n;
    return n;
}

size_t first_token(const char* s, size_t n) {   // bioparser "shorten": up to the first white space
    size_t i = 0;
    while (i < n && !std::isspace(static_cast<unsigned char>(s[i]))) ++i;
    return i;
}

[[noreturn]] void bad_format(const char* what, const std::string& path) {
    throw std::runtime_error(std::string("[racon_b200::reads_io] error: invalid ") + what + " record in " + path);
}

/* sequence.cpp:20-45 */
void finish_sequence(OwnedSequence& s) {
    for (char& c : s.data) c = static_cast<char>(std::toupper(static_cast<unsigned char>(c)));
    uint32_t quality_sum = 0;
    for (char c : s.quality) quality_sum += static_cast<uint32_t>(c - '!');
    if (quality_sum == 0) std::string().swap(s.quality);
}

std::vector<OwnedSequence> parse_fasta(const std::string& path) {
    GzLines in(path);
    std::vector<OwnedSequence> dst;
    const char* p;
    size_t n;
    bool open = false;
    OwnedSequence cur;
    auto flush = [&]() {
        if (!open) return;
        if (cur.data.empty()) bad_format("FASTA", path);
        finish_sequence(cur);
        dst.emplace_back(std::move(cur));
        cur = OwnedSequence();
        open = false;
    };
    while (in.next(&p, &n)) {
        if (n > 0 && p[0] == '>') {
            flush();
            cur.name.assign(p + 1, first_token(p + 1, n - 1));
            open = true;
            continue;
        }
        n = right_strip(p, n);
        if (n == 0) continue;
        if (!open) bad_format("FASTA", path);
        cur.data.append(p, n);
    }
    flush();
    return dst;
}

std::vector<OwnedSequence> parse_fastq(const std::string& path) {
    GzLines in(path);
    std::vector<OwnedSequence> dst;
    const char* p;
    size_t n;
    enum { kName, kData, kQuality } state = kName;
    OwnedSequence cur;
    while (in.next(&p, &n)) {
        switch (state) {
            case kName:
                if (right_strip(p, n) == 0) break;   // blank line between records
                if (p[0] != '@') bad_format("FASTQ", path);
                cur = OwnedSequence();
                cur.name.assign(p + 1, first_token(p + 1, n - 1));
                state = kData;
                break;
            case kData:
                if (n > 0 && p[0] == '+') {
                    if (cur.data.empty()) bad_format("FASTQ", path);
                    state = kQuality;
                    break;
                }
                cur.data.append(p, right_strip(p, n));
                break;
            case kQuality:
                cur.quality.append(p, right_strip(p, n));
                if (cur.quality.size() > cur.data.size()) bad_format("FASTQ", path);
                if (cur.quality.size() == cur.data.size()) {
                    finish_sequence(cur);
                    dst.emplace_back(std::move(cur));
                    state = kName;
                }
                break;
        }
    }
    if (state != kName) bad_format("FASTQ", path);
    return dst;
}

/* a line cut into fields in place (the separators become NULs, so atoi() stops where bioparser's does) */
struct Fields {
    std::string line;
    std::vector<std::pair<size_t, size_t>> at;   // offset, length
    void split(const char* p, size_t n, char sep, size_t max_fields) {
        line.assign(p, right_strip(p, n));
        at.clear();
        size_t b = 0;
        for (;;) {
            size_t e = b;
            while (e < line.size() && line[e] != sep) ++e;
            at.emplace_back(b, e - b);
            if (e == line.size() || at.size() == max_fields) {
                if (e < line.size()) line[e] = '\0';
                break;
            }
            line[e] = '\0';
            b = e + 1;
        }
    }
    const char* str(size_t i) const { return line.c_str() + at[i].first; }
    size_t len(size_t i) const { return at[i].second; }
    uint32_t u32(size_t i) const { return static_cast<uint32_t>(std::atoi(str(i))); }
};

void span_length_and_error(RawOverlap& o, uint32_t q_span, uint32_t t_span) {   // overlap.cpp:26-28
    o.length = std::max(q_span, t_span);
    o.error = 1 - std::min(q_span, t_span) / static_cast<double>(o.length);
}

bool is_cigar_op(char c) {
    return c == 'M' || c == '=' || c == 'X' || c == 'I' || c == 'D' || c == 'N' || c == 'S' || c == 'H' || c == 'P';
}

/* Coordinates of a SAM record from its CIGAR (overlap.cpp:46-118): the leading clip is where the query starts, M/=/X
 * advance both sequences, I the query, D/N the target, clips count towards the read length; a reverse-strand record
 * is turned back into forward read coordinates. */
void sam_coordinates(RawOverlap& o) {
    const std::string& cg = o.cigar;
    if (cg.size() < 2 && o.valid) {
        throw std::runtime_error("[racon_b200::reads_io] error: missing alignment from SAM object!");
    }
    uint32_t q_aligned = 0, q_clipped = 0, t_aligned = 0;
    bool first_op = true;
    for (size_t i = 0, run = 0; i < cg.size(); ++i) {
        const char c = cg[i];
        if (!is_cigar_op(c)) continue;
        const uint32_t count = static_cast<uint32_t>(std::atoi(cg.c_str() + run));
        if (first_op && (c == 'S' || c == 'H')) o.q_begin = static_cast<uint32_t>(std::atoi(cg.c_str()));
        first_op = false;
        switch (c) {
            case 'M': case '=': case 'X': q_aligned += count; t_aligned += count; break;
            case 'I': q_aligned += count; break;
            case 'D': case 'N': t_aligned += count; break;
            case 'S': case 'H': q_clipped += count; break;
            default: break;   // 'P'
        }
        run = i + 1;
    }
    o.q_end = o.q_begin + q_aligned;
    o.q_length = q_clipped + q_aligned;
    if (o.strand) {
        const uint32_t b = o.q_begin;
        o.q_begin = o.q_length - o.q_end;
        o.q_end = o.q_length - b;
    }
    o.t_end = o.t_begin + t_aligned;
    span_length_and_error(o, q_aligned, t_aligned);
}

}  // namespace

bool sequence_format_of(const std::string& path, SequenceFormat* fmt) {
    if (has_extension(path, {".fasta", ".fna", ".fa"})) {
        *fmt = SequenceFormat::kFasta;
        return true;
    }
    if (has_extension(path, {".fastq", ".fq"})) {
        *fmt = SequenceFormat::kFastq;
        return true;
    }
    return false;
}

bool overlap_format_of(const std::string& path, OverlapFormat* fmt) {
    if (has_extension(path, {".mhap"})) {
        *fmt = OverlapFormat::kMhap;
        return true;
    }
    if (has_extension(path, {".paf"})) {
        *fmt = OverlapFormat::kPaf;
        return true;
    }
    if (has_extension(path, {".sam"})) {
        *fmt = OverlapFormat::kSam;
        return true;
    }
    return false;
}

std::vector<OwnedSequence> parse_sequences(const std::string& path) {
    SequenceFormat fmt;
    if (!sequence_format_of(path, &fmt)) {
        throw std::runtime_error("[racon_b200::reads_io] error: file " + path + " has unsupported format extension "
                                 "(valid extensions: .fasta, .fasta.gz, .fna, .fna.gz, .fa, .fa.gz, .fastq, .fastq.gz, "
                                 ".fq, .fq.gz)!");
    }
    return fmt == SequenceFormat::kFasta ? parse_fasta(path) : parse_fastq(path);
}

std::vector<RawOverlap> parse_overlaps(const std::string& path) {
    OverlapFormat fmt;
    if (!overlap_format_of(path, &fmt)) {
        throw std::runtime_error("[racon_b200::reads_io] error: file " + path + " has unsupported format extension "
                                 "(valid extensions: .mhap, .mhap.gz, .paf, .paf.gz, .sam, .sam.gz)!");
    }
    GzLines in(path);
    std::vector<RawOverlap> dst;
    Fields f;
    const char* p;
    size_t n;
    while (in.next(&p, &n)) {
        if (right_strip(p, n) == 0) continue;
        RawOverlap o;
        if (fmt == OverlapFormat::kMhap) {
            /* a_id b_id error minmers a_rc a_begin a_end a_length b_rc b_begin b_end b_length; ids are 1-based ordinals */
            f.split(p, n, ' ', 12);
            if (f.at.size() != 12) bad_format("MHAP", path);
            o.q_id = static_cast<uint64_t>(std::atoll(f.str(0))) - 1;
            o.t_id = static_cast<uint64_t>(std::atoll(f.str(1))) - 1;
            o.strand = f.u32(4) ^ f.u32(8);
            o.q_begin = f.u32(5); o.q_end = f.u32(6); o.q_length = f.u32(7);
            o.t_begin = f.u32(9); o.t_end = f.u32(10); o.t_length = f.u32(11);
            span_length_and_error(o, o.q_end - o.q_begin, o.t_end - o.t_begin);
        } else if (fmt == OverlapFormat::kPaf) {
            /* qname qlen qbegin qend strand tname tlen tbegin tend matches alignment-length mapq [tags] */
            f.split(p, n, '\t', 12);
            if (f.at.size() != 12) bad_format("PAF", path);
            o.q_name.assign(f.str(0), first_token(f.str(0), f.len(0)));
            o.t_name.assign(f.str(5), first_token(f.str(5), f.len(5)));
            if (o.q_name.empty() || o.t_name.empty()) bad_format("PAF", path);
            o.q_length = f.u32(1); o.q_begin = f.u32(2); o.q_end = f.u32(3);
            o.strand = f.str(4)[0] == '-';
            o.t_length = f.u32(6); o.t_begin = f.u32(7); o.t_end = f.u32(8);
            span_length_and_error(o, o.q_end - o.q_begin, o.t_end - o.t_begin);
        } else {
            if (p[0] == '@') continue;   // header
            /* qname flag tname pos mapq cigar rnext pnext tlen seq qual [tags] */
            f.split(p, n, '\t', 11);
            if (f.at.size() != 11) bad_format("SAM", path);
            o.q_name.assign(f.str(0), first_token(f.str(0), f.len(0)));
            o.t_name.assign(f.str(2), first_token(f.str(2), f.len(2)));
            const size_t cigar_len = right_strip(f.str(5), f.len(5));
            const size_t next_len = first_token(f.str(6), f.len(6));
            const size_t data_len = right_strip(f.str(9), f.len(9)), quality_len = right_strip(f.str(10), f.len(10));
            if (o.q_name.empty() || o.t_name.empty() || cigar_len == 0 || next_len == 0 || data_len == 0 ||
                quality_len == 0 || (data_len > 1 && quality_len > 1 && data_len != quality_len)) {
                bad_format("SAM", path);
            }
            const uint32_t flag = f.u32(1);
            o.valid = !(flag & 0x4);
            o.strand = (flag & 0x10) ? 1 : 0;
            o.t_begin = f.u32(3) - 1;
            o.cigar.assign(f.str(5), cigar_len);
            sam_coordinates(o);
        }
        dst.emplace_back(std::move(o));
    }
    return dst;
}

std::vector<SequenceView> InputSet::views() const {
    std::vector<SequenceView> v;
    v.reserve(sequences.size());
    for (const OwnedSequence& s : sequences) {
        v.push_back(SequenceView{s.data.data(), s.quality.empty() ? nullptr : s.quality.data(),
                                 static_cast<uint32_t>(s.data.size())});
    }
    return v;
}

InputSet load_input(const std::string& reads_path, const std::string& overlaps_path, const std::string& targets_path,
                    bool fragment_correction, double error_threshold) {
    /* createPolisher checks the three extensions before anything is read (polisher.cpp:84-141) */
    SequenceFormat sf;
    OverlapFormat of;
    if (!sequence_format_of(reads_path, &sf)) parse_sequences(reads_path);        // throws the extension error
    if (!overlap_format_of(overlaps_path, &of)) parse_overlaps(overlaps_path);
    if (!sequence_format_of(targets_path, &sf)) parse_sequences(targets_path);

    InputSet in;
    in.sequences = parse_sequences(targets_path);
    in.targets_size = in.sequences.size();
    if (in.targets_size == 0) throw std::runtime_error("[racon_b200::load_input] error: empty target sequences set!");

    /* a read that is also a target (same name) is kept once, as the target (polisher.cpp:229-268) */
    std::unordered_map<std::string, uint64_t> target_of_name, read_of_name;
    for (uint64_t i = 0; i < in.targets_size; ++i) target_of_name[in.sequences[i].name] = i;
    std::vector<uint64_t> read_of_ordinal;
    uint64_t total_length = 0;
    {
        std::vector<OwnedSequence> reads = parse_sequences(reads_path);
        if (reads.empty()) throw std::runtime_error("[racon_b200::load_input] error: empty sequences set!");
        read_of_ordinal.reserve(reads.size());
        for (OwnedSequence& r : reads) {
            total_length += r.data.size();
            uint64_t id;
            const auto it = target_of_name.find(r.name);
            if (it != target_of_name.end()) {
                const OwnedSequence& t = in.sequences[it->second];
                if (r.data.size() != t.data.size() || r.quality.size() != t.quality.size()) {
                    throw std::runtime_error("[racon_b200::load_input] error: duplicate sequence " + r.name +
                                             " with unequal data");
                }
                id = it->second;
                read_of_name[r.name] = id;
            } else {
                id = in.sequences.size();
                read_of_name[r.name] = id;
                in.sequences.emplace_back(std::move(r));
            }
            read_of_ordinal.push_back(id);
        }
        in.window_type = static_cast<double>(total_length) / read_of_ordinal.size() <= 1000 ? WindowType::kNGS
                                                                                              : WindowType::kTGS;
    }

    /* overlaps: names / ordinals -> indices (Overlap::transmute, overlap.cpp:133-178), then, per run of consecutive
     * overlaps of one read, the filter of polisher.cpp:284-309 */
    struct Candidate {
        Overlap o;
        uint32_t length;
        double error;
        bool alive;
    };
    std::vector<Candidate> run;
    auto flush_run = [&]() {
        for (size_t i = 0; i < run.size(); ++i) {
            if (!run[i].alive) continue;
            if (run[i].error > error_threshold || run[i].o.q_id == run[i].o.t_id) {
                run[i].alive = false;
                continue;
            }
            if (fragment_correction) continue;
            /* contig polishing keeps one overlap per read: a later one replaces this one only when strictly longer */
            for (size_t j = i + 1; j < run.size(); ++j) {
                if (!run[j].alive) continue;
                if (run[i].length >= run[j].length) {
                    run[j].alive = false;
                } else {
                    run[i].alive = false;
                    break;
                }
            }
        }
        for (Candidate& c : run) {
            if (c.alive) in.overlaps.emplace_back(std::move(c.o));
        }
        run.clear();
    };

    for (RawOverlap& r : parse_overlaps(overlaps_path)) {
        if (!r.valid) continue;
        uint64_t q_id, t_id;
        if (!r.q_name.empty()) {
            const auto it = read_of_name.find(r.q_name);
            if (it == read_of_name.end()) continue;
            q_id = it->second;
        } else {
            if (r.q_id >= read_of_ordinal.size()) continue;
            q_id = read_of_ordinal[r.q_id];
        }
        if (r.q_length != in.sequences[q_id].data.size()) {
            throw std::runtime_error("[racon_b200::load_input] error: unequal lengths in sequence and overlap file for "
                                     "sequence " + in.sequences[q_id].name + "!");
        }
        if (!r.t_name.empty()) {
            const auto it = target_of_name.find(r.t_name);
            if (it == target_of_name.end()) continue;
            t_id = it->second;
        } else {
            if (r.t_id >= in.targets_size) continue;
            t_id = r.t_id;
        }
        if (r.t_length != 0 && r.t_length != in.sequences[t_id].data.size()) {
            throw std::runtime_error("[racon_b200::load_input] error: unequal lengths in target and overlap file for "
                                     "target " + in.sequences[t_id].name + "!");
        }
        Candidate c;
        c.o.q_id = static_cast<uint32_t>(q_id);
        c.o.t_id = static_cast<uint32_t>(t_id);
        c.o.strand = r.strand;
        c.o.q_begin = r.q_begin; c.o.q_end = r.q_end; c.o.q_length = r.q_length;
        c.o.t_begin = r.t_begin; c.o.t_end = r.t_end;
        c.o.t_length = static_cast<uint32_t>(in.sequences[t_id].data.size());
        c.o.cigar = std::move(r.cigar);
        c.length = r.length;
        c.error = r.error;
        c.alive = true;
        if (!run.empty() && run.front().o.q_id != c.o.q_id) flush_run();
        run.emplace_back(std::move(c));
    }
    flush_run();
    if (in.overlaps.empty()) throw std::runtime_error("[racon_b200::load_input] error: empty overlap set!");
    return in;
}

void breaking_points_from_cigar(Overlap& o, uint32_t window_length) {
    o.breaking_points_.clear();
    if (o.t_end == 0 || window_length == 0) return;
    /* t, q: coordinates of the NEXT base of either sequence.  A window of the target ends at the last base before a
     * multiple of window_length, and at the overlap's last base. */
    uint64_t t = o.t_begin, q = o.strand ? o.q_length - o.q_end : o.q_begin;
    const uint64_t last = o.t_end - 1;
    auto window_end = [&](uint64_t pos) { return std::min<uint64_t>((pos / window_length + 1) * window_length - 1, last); };
    bool open = false;
    std::pair<uint32_t, uint32_t> first{0, 0}, past{0, 0};
    auto close_window = [&]() {
        if (open) {
            o.breaking_points_.push_back(first);
            o.breaking_points_.push_back(past);
        }
        open = false;
    };
    const std::string& cg = o.cigar;
    for (size_t i = 0, run = 0; i < cg.size(); ++i) {
        const char c = cg[i];
        if (!is_cigar_op(c)) continue;
        uint64_t count = static_cast<uint32_t>(std::atoi(cg.c_str() + run));
        run = i + 1;
        if (c == 'M' || c == '=' || c == 'X') {
            while (count > 0) {
                const uint64_t we = window_end(t);
                const uint64_t step = std::min<uint64_t>(count, t <= we ? we - t + 1 : count);
                if (!open) {
                    open = true;
                    first = {static_cast<uint32_t>(t), static_cast<uint32_t>(q)};
                }
                t += step;
                q += step;
                count -= step;
                past = {static_cast<uint32_t>(t), static_cast<uint32_t>(q)};
                if (t - 1 == we) close_window();
            }
        } else if (c == 'I') {
            q += count;
        } else if (c == 'D' || c == 'N') {
            while (count > 0) {
                const uint64_t we = window_end(t);
                const uint64_t step = std::min<uint64_t>(count, t <= we ? we - t + 1 : count);
                t += step;
                count -= step;
                if (t - 1 == we) close_window();
            }
        }
    }
}

}  // namespace racon_b200

/* ------------------------------------------------------------------------------------------------------------------
 * Flat-array hooks over the input layer and the file-to-file pipeline, for tests/ and tools (ctypes).
 * ------------------------------------------------------------------------------------------------------------------ */
namespace {
void put_error(const std::exception& e, char* err, uint32_t cap) {
    if (err && cap) {
        std::strncpy(err, e.what(), cap - 1);
        err[cap - 1] = '\0';
    }
}
}  // namespace

/* NULL on failure (message in err) */
extern "C" void* rp_mirror_input_open(const char* reads, const char* overlaps, const char* targets,
                                      int fragment_correction, double error_threshold, char* err, uint32_t err_cap) {
    try {
        return new racon_b200::InputSet(
            racon_b200::load_input(reads, overlaps, targets, fragment_correction != 0, error_threshold));
    } catch (const std::exception& e) {
        put_error(e, err, err_cap);
        return nullptr;
    }
}

/* counts: [0] sequences, [1] targets, [2] bases, [3] overlaps, [4] window type (1 = kTGS), [5] bytes of all names + NULs,
 * [6] bytes of all CIGARs + NULs */
extern "C" void rp_mirror_input_counts(void* hv, uint64_t* counts) {
    const racon_b200::InputSet* in = static_cast<const racon_b200::InputSet*>(hv);
    counts[0] = in->sequences.size();
    counts[1] = in->targets_size;
    counts[2] = counts[5] = counts[6] = 0;
    for (const auto& s : in->sequences) {
        counts[2] += s.data.size();
        counts[5] += s.name.size() + 1;
    }
    counts[3] = in->overlaps.size();
    counts[4] = in->window_type == racon_b200::WindowType::kTGS ? 1 : 0;
    for (const auto& o : in->overlaps) counts[6] += o.cigar.size() + 1;
}

/* quals: '!' where a sequence has none; overlaps: 9 x uint32 per overlap in the order of struct Overlap */
extern "C" void rp_mirror_input_export(void* hv, char* bases, char* quals, uint64_t* seq_off, uint8_t* seq_has_qual,
                                       char* names, uint32_t* overlaps, char* cigars) {
    const racon_b200::InputSet* in = static_cast<const racon_b200::InputSet*>(hv);
    uint64_t nb = 0;
    seq_off[0] = 0;
    for (size_t i = 0; i < in->sequences.size(); ++i) {
        const auto& s = in->sequences[i];
        std::memcpy(bases + nb, s.data.data(), s.data.size());
        if (!s.quality.empty()) {
            std::memcpy(quals + nb, s.quality.data(), s.quality.size());
        } else {
            std::memset(quals + nb, '!', s.data.size());
        }
        seq_has_qual[i] = s.quality.empty() ? 0 : 1;
        nb += s.data.size();
        seq_off[i + 1] = nb;
        std::memcpy(names, s.name.c_str(), s.name.size() + 1);
        names += s.name.size() + 1;
    }
    for (size_t i = 0; i < in->overlaps.size(); ++i) {
        const auto& o = in->overlaps[i];
        const uint32_t v[9] = {o.q_id, o.t_id, o.strand, o.q_begin, o.q_end, o.q_length, o.t_begin, o.t_end, o.t_length};
        std::memcpy(overlaps + 9 * i, v, sizeof(v));
        std::memcpy(cigars, o.cigar.c_str(), o.cigar.size() + 1);
        cigars += o.cigar.size() + 1;
    }
}

/* host only: breaking points of overlap `index` from the CIGAR it came with; returns the number of points (pairs of
 * uint32 written to out while they fit `cap` points) */
extern "C" uint32_t rp_mirror_input_cigar_breaking_points(void* hv, uint32_t index, uint32_t window_length, uint32_t* out,
                                                          uint32_t cap) {
    const racon_b200::InputSet* in = static_cast<const racon_b200::InputSet*>(hv);
    racon_b200::Overlap o = in->overlaps[index];
    racon_b200::breaking_points_from_cigar(o, window_length);
    for (uint32_t i = 0; i < o.breaking_points_.size() && i < cap; ++i) {
        out[2 * i] = o.breaking_points_[i].first;
        out[2 * i + 1] = o.breaking_points_[i].second;
    }
    return static_cast<uint32_t>(o.breaking_points_.size());
}

extern "C" void rp_mirror_input_close(void* hv) { delete static_cast<racon_b200::InputSet*>(hv); }
