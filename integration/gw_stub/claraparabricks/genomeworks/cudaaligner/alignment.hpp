/* Stand-in for <claraparabricks/genomeworks/cudaaligner/alignment.hpp> (included by src/cuda/cudaaligner.hpp:8): the
 * shim reads CIGARs straight from the C ABI (rp_aln_fetch_cigar), so no Alignment type is needed. */
#pragma once
