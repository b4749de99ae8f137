// SPDX-License-Identifier: MIT
//
// Batched Stwo PROVER for the wide-Fibonacci AIR that stwo-verifier/src/verifier.simf:32-58 checks
// (the reference ships no prover; its fixtures came from an external fork).
// It produces packed proofs (include/ssym.h layout) that the verifier kernels accept in
// SSYM_MODE_PROVER_CONSISTENT, so BASELINE configs 3 and 5 (2^16 / 2^20 DISTINCT proofs) can be generated
// on the GPU that verifies them.  Bit-exact twin of oracle/stwo_prover_ref.c (the CPU checker).
//
// Pipeline over a chunk of m proofs (everything stays in HBM; ~3 MB of scratch per proof at the prod preset):
//   P1 prv_trace_kernel       CTA per proof         trace rows from the seed, circle IFFT (2^T) and LDE FFT (2^G) in shared memory
//   P2 leaf + tree kernels    thread per node       SHA-256 Merkle trees in heap order (node 1 = root, 2^G + q = leaf q)
//   P3 prv_cp_kernel          CTA per (proof,coord) composition polynomial on the LDE domain, IFFT, split into the 4 (x,y)-parity
//                                                   sub-polynomials, 4 LDE FFTs                  (evals/composition_poly.simf:38-59)
//   P4 prv_oods_kernel        warp per (proof,col)  samples at the OODS point P (trace) and at 2P (CP columns)
//   P5 prv_quotient_kernel    thread per (proof,q)  DEEP quotient = fri_answer of every LDE position (fri/answers.simf:97-129 + Appendix A)
//   P6 per FRI layer          leaf/tree kernels, channel step, prv_fold_kernel (fri/folding.simf:15-41)
//   P7 prv_final_kernel       thread per proof      last layer, proof-of-work grind (pow.simf:22-35), queries (fri/queries.simf:30-43)
//   P8 prv_decommit_kernel    warp per (proof,query) gathers values, witnesses and authentication paths into the packed record
//   channel kernels (thread per proof) replay channel.simf between the phases.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/ssym.h"

namespace ssym {

struct PrvCtx { // per-proof prover context, u32 words
    enum : uint32_t {
        CH = 0,          // [9]  channel digest + n_sent
        CP_ALPHA = 12,   // [4]
        PX = 16, PY = 20, P2X = 24, P2Y = 28, // OODS point P and 2P
        TW = 32,         // [14][4]  P.y, P.x, pi(P.x), pi^2(P.x), ...  (basis factors of the circle-FFT basis at P)
        DEEP_ALPHA = 88, // [4]
        KB = 92,         // [16 + C][4]  line coefficient b of column k with alpha^(k+1): 16 CP columns (at 2P), then the C <= 16 trace columns (at P)
        SUMS = 220,      // [4][4]   sum a / sum c of batch A (CP), sum a / sum c of batch B (trace)
        FRI_ALPHA = 236, // [9][4]
        QUERIES = 272,   // [16]
        N_USED = 288,    // [1] slots the decommitments fill: n_queries, or the distinct queries under SSYM_MODE_QUERY_DEDUP
        WORDS = 292
    };
};

#define SSYM_PRV_MAX_LOG 13 /* LDE log size the shared-memory FFTs are sized for (4 columns x 2^13 words = 128 KB) */

// Twiddle tables of one canonic-coset domain of log size n.  Layer l (0 = the circle layer, y coordinates; l >= 1 = x coordinates of
// line_domain(n - l)) has 2^(n-l-1) entries starting at 2^n - 2^(n-l):  tw[l][j] = coordinate at bit_reverse(2j, n - l).
struct PrvDomain {
    const uint32_t *tw, *itw;
    uint32_t log, scale; // scale = 2^-log mod p
};

struct PrvParams {
    ssym_stwo_config_t cfg;
    ssym_stwo_layout_t lo;
    PrvDomain tr, lde;
    const uint2 *point;         // [2^G] domain point of LDE slot q (shared with the verifier's tables)
    const uint32_t *vanish_inv; // [2^G] 1 / vanishing(trace_log) at that point
    const uint64_t *seeds;      // m
    uint32_t *out;              // m packed records
    uint32_t *pctx;             // m * PrvCtx::WORDS
    uint32_t *tcoef;            // m * C * 2^T   (C = NUM_COLUMNS)
    uint32_t *tlde;             // m * C * 2^G
    uint32_t *cpcoef;           // m * 4 * 2^(T+1)
    uint32_t *cplde;            // m * 16 * 2^G
    uint32_t *tree_t, *tree_c;  // m * 2^(G+1) * 8
    uint32_t *fev;              // m * fev_stride : per proof, layer l at fev_off[l] (QM31 = 4 words each)
    uint32_t *ftree;            // m * ftree_stride : per proof, layer l tree at ftree_off[l]
    uint32_t fev_off[SSYM_MAX_FRI_LAYERS + 1], fev_stride;
    uint32_t ftree_off[SSYM_MAX_FRI_LAYERS], ftree_stride;
    uint32_t *flag;             // bit 0: low-degree self check failed, bit 1: draw exhausted / zero inverse (never for honest traces)
    uint32_t m;
};

// Builds tw / itw of one domain (2^n entries each) and, for the LDE domain, vanish_inv.
void launch_prv_tables(uint32_t n, uint32_t *tw, uint32_t *itw, cudaStream_t s);
void launch_prv_vanish(uint32_t trace_log, uint32_t lde_log, const uint2 *point, uint32_t *vanish_inv, cudaStream_t s);
// Proves p.m proofs; returns the number of kernels launched.  cudaFuncSetAttribute for the large-shared-memory kernels is done inside.
int launch_prv_prove(const PrvParams &p, cudaStream_t s, uint64_t *launch_counter);

} // namespace ssym
