/*
 * gather_core.cuh — layer extraction on the device: the packed bases / weights arrays the POA kernel reads
 * are filled from a device-resident read store instead of being copied together on the host and uploaded.
 * One warp per packed sequence; a layer is a slice of a read, forward or as its reverse complement
 * (Sequence::create_reverse_complement, /root/reference/src/sequence.cpp:58-93: A<->T, C<->G, anything else unchanged,
 * qualities reversed), weights are phred - 33 (spoa graph.cpp:141-143), 1 for a layer without qualities, 0 for a backbone
 * without qualities (racon's dummy '!' string, polisher.cpp:174,396-399).
 * HBM-bound byte work: 1 byte read (2 with qualities) and 2 bytes written per base, coalesced across the warp.
 */
#pragma once
#include <stdint.h>

#include "rp_warp.cuh"

namespace rp {

constexpr uint8_t kSrcReverse = 1, kSrcHasQuality = 2, kSrcBackbone = 4;

struct GatherParams {
    const uint8_t* store_bases;   // every sequence of the store, concatenated
    const uint8_t* store_quals;   // same offsets; only read for sequences that have qualities
    uint8_t* bases;               // packed batch arrays (PoaParams::bases / weights)
    uint8_t* weights;             // NULL: bases only (the aligner's inputs)
    const uint32_t* seq_off;      // packed offsets, n_seqs + 1
    const uint64_t* src_pos;      // store position of the byte that becomes the sequence's first base
    const uint8_t* src_flags;
    uint32_t n_seqs;
};

RP_HD uint8_t complement_base(uint8_t c) {
    switch (c) {
        case 'A': return 'T';
        case 'T': return 'A';
        case 'C': return 'G';
        case 'G': return 'C';
        default: return c;
    }
}

RP_DEV void gather_sequence(const GatherParams& P, uint32_t s) {
    const uint32_t first = P.seq_off[s], n = P.seq_off[s + 1] - first;
    const uint64_t pos = P.src_pos[s];
    const uint8_t flags = P.src_flags[s];
    const bool reverse = flags & kSrcReverse;
    const uint8_t flat_weight = (flags & kSrcBackbone) ? 0 : 1;
    for (uint32_t j = static_cast<uint32_t>(lane_id()); j < n; j += 32) {
        const uint64_t at = reverse ? pos - j : pos + j;
        const uint8_t c = P.store_bases[at];
        P.bases[first + j] = reverse ? complement_base(c) : c;
        if (P.weights)
            P.weights[first + j] = (flags & kSrcHasQuality) ? static_cast<uint8_t>(P.store_quals[at] - 33) : flat_weight;
    }
}

}  // namespace rp
