/*
 * ref_edlib_harness.cpp — TEST INFRASTRUCTURE ONLY.
 *
 * extern "C" driver over the UNMODIFIED vendored edlib, calling it exactly the way the reference
 * does in Overlap::align_overlaps (/root/reference/src/overlap.cpp:205-224):
 *   edlibAlign(q, ql, t, tl, edlibNewAlignConfig(-1, EDLIB_MODE_NW, EDLIB_TASK_PATH, nullptr, 0))
 *   edlibAlignmentToCigar(alignment, alignmentLength, EDLIB_CIGAR_STANDARD)
 * Compiled by oracle/Makefile into oracle/_ref/libracon_ref.so.
 */
#include <cstdint>
#include <cstdlib>
#include <cstring>

#include "edlib.h"

extern "C" {

/* Returns CIGAR length (excluding NUL), -1 on edlib failure, -2 if `cap` is too small.
 * `edit_distance` (may be null) receives edlib's editDistance. */
int64_t ref_edlib_cigar(const char* q, uint32_t ql, const char* t, uint32_t tl, char* out,
                        uint64_t cap, int32_t* edit_distance) {
    EdlibAlignResult result =
        edlibAlign(q, ql, t, tl, edlibNewAlignConfig(-1, EDLIB_MODE_NW, EDLIB_TASK_PATH, nullptr, 0));
    if (result.status != EDLIB_STATUS_OK) {
        edlibFreeAlignResult(result);
        return -1;
    }
    if (edit_distance) *edit_distance = result.editDistance;
    char* cigar = edlibAlignmentToCigar(result.alignment, result.alignmentLength, EDLIB_CIGAR_STANDARD);
    size_t n = std::strlen(cigar);
    int64_t ret = static_cast<int64_t>(n);
    if (n + 1 > cap) {
        ret = -2;
    } else {
        std::memcpy(out, cigar, n + 1);
    }
    std::free(cigar);
    edlibFreeAlignResult(result);
    return ret;
}

}  // extern "C"
