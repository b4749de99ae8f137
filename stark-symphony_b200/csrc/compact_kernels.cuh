// SPDX-License-Identifier: MIT
//
// Compact transport form of packed Stwo proofs (include/ssym.h "compact transport form"): per tree, every distinct sibling digest
// once + one bit per path slot + one index per repeated slot.  The host link carries the compact bytes; stwo_expand_kernel rebuilds the packed records in HBM.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/ssym.h"

namespace ssym {

enum { COMPACT_HDR_WORDS = 8, COMPACT_MAX_TREES = SSYM_MAX_FRI_LAYERS + 2, COMPACT_MAX_BITMAP_WORDS = 256 };

// Section geometry of a compact record for one configuration (host-computed, passed by value).
struct CompactShape {
    uint32_t trees;                              // T = n_fri_layers + 3
    uint32_t slots;                              // sibling slots of a proof, all trees
    uint32_t slot_first[COMPACT_MAX_TREES + 1];  // first slot of tree t; [T] = slots
    uint32_t head_slots;                         // slots of the trace + composition trees (contiguous from off_trace_sib); the rest from off_fri_sib[0]
    uint32_t idx_bytes;                          // 1 or 2
    uint32_t fixed_words;                        // packed [0, off_trace_sib)
    uint32_t wit_words;                          // packed [off_fri_wit, off_fri_sib[0])
    uint32_t bitmap_words;                       // ceil(slots / 32)
    uint32_t off_wit, off_bitmap, off_refs;      // word offsets inside the compact record (the fixed part starts at COMPACT_HDR_WORDS); the
                                                 // table follows the R back references: off_refs + compact_refs_words(R)
    uint32_t off_bitmap2, off_refs3;             // version 3 records: the "derived" bitmap, and where their references start (one bitmap further)
    uint32_t max_words;                          // a record with no repeated sibling (version 3: the larger of the two)
    uint32_t n_queries;                          // Q (partner numbers of derived slots are < Q)
};
int compact_shape(const ssym_stwo_config_t &cfg, const ssym_stwo_layout_t &lo, CompactShape &sh);
// words of the reference section: R back references of idx_bytes each, then (version 3) X partner bytes
__host__ __device__ inline uint32_t compact_refs_words(const CompactShape &sh, uint32_t refs, uint32_t derived = 0) {
    return ((refs * sh.idx_bytes + derived + 31u) / 32u) * 8u;
}

struct CompactParams {
    CompactShape sh;
    ssym_stwo_layout_t lo;
    const uint32_t *blob;    // record i at blob + (offsets[i] - base)
    const uint64_t *offsets; // n + 1, device
    uint64_t base;
    uint32_t *packed;        // n * stride_words
    uint32_t *flags;         // n or nullptr
    uint32_t n;
    uint8_t *derive;         // nullptr, or n * sh.slots bytes: the derive table of StwoParams (0xff, or the partner query of a derived slot); a
                             // version 3 record with derived slots is malformed without it
    uint32_t mode;           // the mode derived slots are expanded under: a record with derived slots must carry the same in header word 5
};
void launch_stwo_expand(const CompactParams &p, cudaStream_t s);
// status[i] |= SSYM_ST_SHAPE, accept bit i cleared, where flags[i] != 0
void launch_compact_apply_flags(const uint32_t *flags, uint32_t *status, uint32_t *accept_bits, uint32_t n, cudaStream_t s);

} // namespace ssym
