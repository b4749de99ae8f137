// SPDX-License-Identifier: MIT
//
// Batched verify_proof of the Stwo wide-Fibonacci verifier (stwo-verifier/src/verifier.simf:32-58)
// as four kernels over a batch of packed proofs (layout: include/ssym.h):
//
//   K1 stwo_channel_kernel     one thread per proof     Fiat-Shamir channel: ~46 dependent compressions, PoW, queries; the per-proof
//                                                      scalars: OODS point, composition-polynomial check, powers of the DEEP coefficient
//   K2 stwo_query_kernel       one warp per proof       DEEP line coefficients (lane = column), fri_answer + the 1+L folds (lane = query)
//   K3 stwo_merkle_kernel      one thread per hash chain  all 2*Q + (L+1)*Q Merkle decommitments
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
        ALPHA_POW = 76,  // [21][4] deep_alpha^(k+1), k = 0..20 (k = 20: the batch coefficient of fri/answers.simf:126)
        WORDS = 160
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
};

// Optional per-kernel event timing (ssym_profile_enable): begin/end bracket one kernel launch on stream s.
struct Profiler {
    virtual void begin(int kernel_id, cudaStream_t s) = 0;
    virtual void end(int kernel_id, cudaStream_t s) = 0;
};

void launch_stwo_tables(uint32_t lde_log, uint32_t n_fri_layers, uint2 *point, uint32_t *fold_inv, const uint32_t *fold_off,
                        uint32_t *zero_flag, cudaStream_t s);
void launch_stwo_verify(const StwoParams &p, uint32_t *accept_bits, cudaStream_t s, uint64_t *launch_counter, Profiler *prof,
                        cudaStream_t front = nullptr, cudaEvent_t front_done = nullptr, int front_kernels = 0);

} // namespace ssym
