/*
 * bp_oracle.cpp — CPU ORACLE for CIGAR -> breaking points.  TEST INFRASTRUCTURE ONLY.
 *
 * Restates Overlap::find_breaking_points_from_cigar (/root/reference/src/overlap.cpp:226-292) base by base:
 *   window ends  : every i-1 with i a multiple of window_length, t_begin < i < t_end, then t_end-1   (:229-235)
 *   walk         : q pointer starts at (strand ? q_length - q_end : q_begin) - 1, t pointer at t_begin - 1 (:241-242);
 *                  'M'/'='/'X' advance both and remember the first match of the current window and the position
 *                  after the latest match (:245-267); 'I' advances q (:268-270); 'D'/'N' advance t (:271-286);
 *                  when the t pointer reaches a window end and the window saw a match, the pair
 *                  (first match, one past last match) is emitted (:258-264, :276-282).
 * PINNING: tests/test_breaking_points.py feeds it the unmodified edlib's CIGAR of every overlap of the lambda sample
 * and compares with the breaking points the unmodified reference produced (tests/golden/lambda_overlaps.npz).
 */
#include <cstdint>
#include <cstdlib>
#include <vector>

extern "C" {

/* out: (t, q) uint32 pairs, two pairs per window with a match.  Returns the number of PAIRS written, -2 if cap
 * (in pairs) is too small. */
int64_t oracle_breaking_points(const char* cigar, uint64_t cigar_len, uint32_t t_begin, uint32_t t_end,
                               uint32_t q_start, uint32_t window_length, uint32_t* out, uint64_t cap) {
    std::vector<int64_t> ends;
    for (uint64_t i = 0; i < t_end; i += window_length)
        if (i > t_begin) ends.push_back(static_cast<int64_t>(i) - 1);
    ends.push_back(static_cast<int64_t>(t_end) - 1);
    size_t w = 0;
    bool found = false;
    uint32_t first_t = 0, first_q = 0, last_t = 0, last_q = 0;
    int64_t q = static_cast<int64_t>(q_start) - 1, t = static_cast<int64_t>(t_begin) - 1;
    uint64_t n = 0;
    auto emit = [&]() -> bool {
        if (n + 2 > cap) return false;
        out[2 * n] = first_t; out[2 * n + 1] = first_q; ++n;
        out[2 * n] = last_t; out[2 * n + 1] = last_q; ++n;
        return true;
    };
    uint64_t num = 0;
    for (uint64_t i = 0; i < cigar_len; ++i) {
        const char c = cigar[i];
        if (c >= '0' && c <= '9') {
            num = num * 10 + static_cast<uint64_t>(c - '0');
            continue;
        }
        if (c == 'M' || c == '=' || c == 'X') {
            for (uint64_t k = 0; k < num; ++k) {
                ++q; ++t;
                if (!found) {
                    found = true;
                    first_t = static_cast<uint32_t>(t);
                    first_q = static_cast<uint32_t>(q);
                }
                last_t = static_cast<uint32_t>(t + 1);
                last_q = static_cast<uint32_t>(q + 1);
                if (w < ends.size() && t == ends[w]) {
                    if (found && !emit()) return -2;
                    found = false;
                    ++w;
                }
            }
        } else if (c == 'I') {
            q += static_cast<int64_t>(num);
        } else if (c == 'D' || c == 'N') {
            for (uint64_t k = 0; k < num; ++k) {
                ++t;
                if (w < ends.size() && t == ends[w]) {
                    if (found && !emit()) return -2;
                    found = false;
                    ++w;
                }
            }
        }
        num = 0;
    }
    return static_cast<int64_t>(n);
}

}  // extern "C"
