// SPDX-License-Identifier: MIT
//
// Batched verify_proof of the Stwo wide-Fibonacci verifier (stwo-verifier/src/verifier.simf:32-58)
// as four kernels over a batch of packed proofs (layout: include/ssym.h):
//
//   K1 stwo_channel_kernel     one thread per proof     Fiat-Shamir channel: ~46 dependent compressions, PoW, queries; the per-proof
//                                                      scalars: OODS point, composition-polynomial check, powers of the DEEP coefficient
//   K2 stwo_query_kernel       one warp per proof       DEEP line coefficients (lane = column), fri_answer + the 1+L folds (lane = query)
//   K3 stwo_merkle_kernel      one thread per hash chain  all 2*Q + (L+1)*Q Merkle decommitments
//      (default: stwo_plan_kernel, stwo_merkle_shared_kernel, stwo_check_kernel, stwo_merkle_shared_kernel (round 2), stwo_resolve_kernel:
//       the same decommitments with the nodes that several paths of a tree run through hashed once — see StwoDedup below)
//   K4 stwo_finalize_kernel    status words -> accept bitmap
//
// No early exit anywhere: every check is evaluated and OR-ed into the proof's status word, so a
// rejected proof costs exactly what an accepted one does.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/ssym.h"
#include "field.cuh"
#include "sha256.cuh"

namespace ssym {

// Per-proof context written by K1 (the channel), read by K2 / K3 (u32 words).
struct StwoCtxLayout {
    enum : uint32_t {
        QUERIES = 0,     // [16]   fri/queries.simf:30-43
        CP_ALPHA = 16,   // [4]    evals/commit.simf:29
        DEEP_ALPHA = 20, // [4]    deep/oods.simf:61
        FRI_ALPHA = 24,  // [9][4] fri/commit.simf:42
        PX = 60,         // [4]    OODS point, channel.simf:143-151
        PY = 64,         // [4]
        P2X = 68,        // [4]    sample point of the 16 CP columns: P (REF_LITERAL) or 2P (PROVER_CONSISTENT, Appendix A item 1)
        P2Y = 72,        // [4]
        ALPHA_POW = 76,  // [C + 17][4] deep_alpha^(k+1), k = 0..C+16 (the last: the batch coefficient of fri/answers.simf:126); C <= 16
        N_USED = 76 + 4 * (SSYM_MAX_COLUMNS + SSYM_NUM_CP_PARTITIONS + 1), // [1] U: queries verified (= n_queries, or the distinct ones under SSYM_MODE_QUERY_DEDUP)
        WORDS = N_USED + 4
    };
};

// Domain tables, built once per (lde_log, n_fri_layers) with the literal reference functions so
// every entry is bit-identical to what the .simf code recomputes per query:
//   point[q]        = circle_position_to_m31_point(circle_domain(G), bit_reverse_position(q, G))   fri/answers.simf:108-110
//   fold_inv[0][j]  = m31_inv(y of the point at bit_reverse_position(2j, G))                       fri/folding.simf:18-20
//   fold_inv[l][j]  = m31_inv(line_position_to_x_coord(line_domain(G-l), bit_reverse_position(2j, G-l)))  fri/folding.simf:31-34
struct StwoTables {
    const uint2 *point;       // [2^G]
    const uint32_t *fold_inv; // concatenated; layer l starts at fold_off[l], has 2^(G-l-1) entries
    uint32_t fold_off[SSYM_MAX_FRI_LAYERS];
};

// Node sharing between the Merkle paths of one tree (SURVEY.md 8a "merkle.simf": the reference hashes every query's path to the root
// on its own; two paths that have met run through the same nodes from there on).  With STWO_DEDUP_MAX_DEPTH >= tree depth the Merkle work of a
// proof is planned per tree: query q follows the lowest-numbered query r < q whose path it meets first, at height h_q, and hashes only the
// h_q - 1 levels below the two children of the meeting node.  There q's node must be r's sibling and q's sibling r's node: then the meeting
// node has bit-identical inputs on both paths, and if q's siblings above it are r's too, q's root IS r's root.  All of this is checked word for
// word (followers compare proof data in their own thread, a check kernel compares the one hashed node); a follower that fails any comparison
// — a corrupted proof — is hashed to the root after all, in a second round of the same hashing kernel, so every per-query result is exactly
// what the per-query schedule gives.
enum { STWO_DEDUP_MAX_DEPTH = 16, STWO_DEDUP_MAX_BINS = 64 };
struct StwoDedup {
    // per chain c = tree * Q + q of proof i (tree 0 = trace, 1 = composition, 2 + l = FRI layer l), index i * chains + c:
    uint32_t *plan;      // bits 0-4 lv (levels hashed in round 1), bit 5 sibling mismatch, bit 6 check failed, bit 7 follower, bits 8-11 leader query,
                         // bit 12 same leaf as the leader, bits 16-31 checkpoint mask (bit 16 + k: this chain's node at height k is wanted)
    uint64_t *ckpt_to;   // nibble k: the follower query that needs this chain's node at height k
    uint32_t *own;       // [8] round 1: the chain's last node = the root (full chains) or the node at height lv (followers)
    uint32_t *ckpt;      // [8] written by the LEADER of this chain: the leader's node at this chain's height lv; after the check, round 2 puts
                         //     the root of a follower that failed it here
    uint32_t *bin_count; // [2][MAX_BINS] tasks per bin and round; a bin = (kind, steps) so that a warp's 32 tasks have one length; longest first
    uint32_t *bin_list;  // chain ids (i * chains + c); bin b of round r at bin_base[r][b] .. + its capacity
    uint32_t bin_base[2][STWO_DEDUP_MAX_BINS];
    uint8_t bin_of[2][4][STWO_DEDUP_MAX_DEPTH + 1]; // round, kind (0 trace, 1 composition, 2 FRI, 3 = from a stored node), steps -> bin; 0xff = none
    uint32_t n_bins[2], chains;
    uint32_t enabled;
};

struct StwoParams {
    ssym_stwo_config_t cfg;
    ssym_stwo_layout_t lo;
    StwoTables tab;
    const uint32_t *packed; // n * stride_words
    uint32_t *ctx;          // n * StwoCtxLayout::WORDS
    uint32_t *fri_evals;    // n * (L+1) * Q * 4 : evaluation entering layer l for query q
    uint32_t *status;       // n
    ssym_stwo_trace_t *trace; // n or nullptr
    uint32_t n;
    StwoDedup dd;
    // Compact transport form, version 3 (compact_kernels.cuh): one byte per sibling slot of a proof (slot order = the compact record's: trace [Q][G],
    // composition [Q][G], FRI layer l [Q][G-1-l]), at derive + i * derive_stride.
    //   derive_mode 1 (verify / expand): byte p < Q = the slot's sibling is NOT in the record: it is the node of query p's path of the same tree at
    //     the same level (the Merkle kernel takes it from p's lane and writes it into the packed record); 0xff = the sibling is in the packed record.
    //   derive_mode 2 (pack): the kernel WRITES the table: for every slot the lowest p whose node equals the slot's sibling, else 0xff.
    uint8_t *derive;
    uint32_t derive_stride, derive_mode;
    uint32_t *packed_rw; // = packed, writable (derive_mode 1)
    // Records packed under another semantics than the call's (launch_stwo_verify_cross): the kernels of one pass run on the transcript context of
    // the other.  ctx_mode = the mode K1 ran under (K2 re-derives the sample point of the CP columns when its own semantics differ);
    // derive_kinds: which trees take derived siblings in this pass (bit 0 trace + composition, bit 1 FRI); fri_only: K3 runs the FRI chains only.
    uint32_t ctx_mode, derive_kinds, fri_only;
};

// Fills the static part of StwoDedup (bins, capacities) for a configuration and a chunk capacity of `cap` proofs; returns the number of
// bin_list entries needed, or 0 when the configuration is outside the planner's range (the per-query kernel is used then).
size_t stwo_dedup_layout(const ssym_stwo_config_t &cfg, size_t cap, StwoDedup &dd);

// Optional per-kernel event timing (ssym_profile_enable): begin/end bracket one kernel launch on stream s.
struct Profiler {
    virtual void begin(int kernel_id, cudaStream_t s) = 0;
    virtual void end(int kernel_id, cudaStream_t s) = 0;
};

cudaError_t stwo_kernels_init_device(); // once per device, before the first launch (opt-in shared memory of the transcript kernel)
void launch_stwo_tables(uint32_t lde_log, uint32_t n_fri_layers, uint2 *point, uint32_t *fold_inv, const uint32_t *fold_off,
                        uint32_t *zero_flag, cudaStream_t s);
void launch_stwo_verify(const StwoParams &p, uint32_t *accept_bits, cudaStream_t s, uint64_t *launch_counter, Profiler *prof,
                        cudaStream_t front = nullptr, cudaEvent_t front_done = nullptr, int front_kernels = 0);
// Version 3 compact records whose derived siblings were packed under `rec_mode`, verified under p.cfg.mode (same flags, other semantics): the
// transcript once; the evaluations and the FRI chains of the records' mode, which complete the FRI siblings in the packed records (that pass has no
// status: its verdicts are not wanted); then the verification proper.  p.derive / derive_stride as for derive_mode 1.
void launch_stwo_verify_cross(const StwoParams &p, uint32_t rec_mode, uint32_t *accept_bits, cudaStream_t s, uint64_t *launch_counter);

} // namespace ssym
