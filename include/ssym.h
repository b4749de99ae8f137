/* SPDX-License-Identifier: MIT
 *
 * libssym — B200-native batched STARK verifier: the C-ABI drop-in boundary.
 *
 * Every entry point below replaces one `.simf` function (or one Simplicity jet
 * family) of starkware-bitcoin/stark-symphony; the reference location each one
 * stands in for is cited as `file:line` relative to the reference root.  The
 * reference has no FFI of its own for this path (it is one `simfony run` process
 * per proof, simfony-cli/src/main.rs:163-209): these are the symbols a
 * `verify-batch` sub-command next to `build/run` would bind.
 *
 * Conventions
 *  - plain C, no torch / C++ types; all sizes in elements unless stated.
 *  - return value: 0 = ok, <0 = usage or CUDA error (ssym_last_error() has text).
 *    A rejected proof is DATA (accept bit 0 / non-zero status), never an error.
 *  - field elements are raw `uint32_t` exactly as the jets see them; nothing is
 *    canonicalised on load (m31.simf:17-44 semantics for arbitrary u32).
 *  - a 256-bit digest is 8 `uint32_t` words, word 0 = most significant 32 bits of
 *    the big-endian u256 (channel.simf:48-58 `split_256` order).
 *  - `memspace`: SSYM_MEM_DEVICE pointers are device pointers on the handle's
 *    GPU (inputs already resident in HBM); SSYM_MEM_HOST pointers are host
 *    pointers (pinned or pageable) and the call performs the H2D / D2H copies.
 *  - no CPU fallback exists: every compute entry point runs CUDA kernels on an
 *    sm_100-class device or fails with SSYM_ERR_CUDA.
 */
#ifndef SSYM_H
#define SSYM_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif
#if defined(__GNUC__)
#pragma GCC visibility push(default) /* the library is built with -fvisibility=hidden; only this API is exported */
#endif

#define SSYM_OK 0
#define SSYM_ERR_USAGE (-1)
#define SSYM_ERR_CUDA (-2)
#define SSYM_ERR_PARSE (-3)
#define SSYM_ERR_NOMEM (-4)
#define SSYM_ERR_INTERNAL (-5)

#define SSYM_MEM_DEVICE 0
#define SSYM_MEM_HOST 1

/* ------------------------------------------------------------------------- */
/* Stwo verifier configuration (stwo-verifier/src/config.simf:10-51)          */
/* ------------------------------------------------------------------------- */

/* Semantics switch (SURVEY.md "Read this first", finding 3 / Appendix A):
 *  REF_LITERAL        exactly what the .simf text computes at reference HEAD.
 *  PROVER_CONSISTENT  the three HEAD inconsistencies (F1-F3) resolved the way
 *                     the prover of tests/data/proof.json behaves. */
#define SSYM_MODE_REF_LITERAL 0
#define SSYM_MODE_PROVER_CONSISTENT 1
/* Flag, OR-ed into either of the two: the queries fri_generate_queries draws are SORTED and DE-DUPLICATED before use — the step
 * fri/queries.simf:41 says the reference leaves out "to simplify the implementation" (upstream stwo's Queries::generate collects them into
 * an ordered set).  U <= n_queries positions remain.  The witness types are fixed-size arrays, so the shapes do not change: slot j of
 * DECOMMITMENTS / of every FRI layer decommitment belongs to the j-th smallest distinct query for j < U, and slots j >= U are not looked at
 * (a prover zero-fills them; the compact transport form stores their repeated digests once).  Every per-query check — decommitments, DEEP
 * quotient, folds, last layer — runs for j < U only.  This is configuration the reference announces but does not define; its results are
 * pinned by the oracle (which implements the same rule) and by the two provers, not by a reference vector. */
#define SSYM_MODE_QUERY_DEDUP 2
#define SSYM_MODE_SEMANTICS(mode) ((mode) & 1u)

#define SSYM_NUM_COLUMNS 4        /* config.simf:14 NUM_COLUMNS at reference HEAD (the default)   */
#define SSYM_MAX_COLUMNS 16       /* largest supported NUM_COLUMNS (a power of two, config.simf:12-14) */
#define SSYM_NUM_CP_PARTITIONS 16 /* evals/composition_poly.simf:13                 */
#define SSYM_MAX_QUERIES 16       /* config.simf:42 NUM_FRI_QUERIES (prod)          */
#define SSYM_MAX_FRI_LAYERS 9     /* first layer + NUM_FRI_LAYERS (config.simf:47)  */

typedef struct ssym_stwo_config {
    uint32_t trace_log;    /* TRACE_LOG_SIZE      config.simf:17,35 */
    uint32_t lde_log;      /* LDE_LOG_SIZE        config.simf:21,39 */
    uint32_t n_queries;    /* NUM_FRI_QUERIES     config.simf:25,43  (1..16)        */
    uint32_t n_fri_layers; /* NUM_FRI_LAYERS      config.simf:29,47  (inner, 0..8)  */
    uint32_t mode;         /* SSYM_MODE_*                                           */
    uint32_t n_columns;    /* NUM_COLUMNS         config.simf:14     (4, 8 or 16; 0 = 4, the value at reference HEAD).
                            * The AIR stays the wide-Fibonacci one (constraints/wide_fibonacci.simf:24-62: column i >= 2 is
                            * constrained by c_i = c_{i-1}^2 + c_{i-2}^2), only its width changes: the `CONFIG:` comments of the
                            * reference (fri/answers.simf:118, evals/verify.simf, hasher.simf:85-90) name the macros to re-point. */
    uint64_t pow_target;   /* POW_TARGET_64       config.simf:32,51 */
} ssym_stwo_config_t;

/* NUM_COLUMNS of a configuration (n_columns == 0 means the reference's 4). */
#define SSYM_STWO_COLUMNS(cfg) ((cfg)->n_columns ? (cfg)->n_columns : (uint32_t)SSYM_NUM_COLUMNS)

/* The two presets of config.simf (TESTING / production). */
int ssym_stwo_config_preset(const char *name /* "prod" | "testing" */, uint32_t mode,
                            ssym_stwo_config_t *out);

/* Packed wire format of one Stwo proof (all little-endian u32 words; digests as
 * 8 words, most significant first).  Sections, in order, each starting on a
 * 32-byte boundary (Q = n_queries, L = n_fri_layers, G = lde_log, C = NUM_COLUMNS):
 *   header   : commit[3][8] | oods_trace[C][4] | oods_cp[16][4] | fri_first_root[8]
 *              | fri_inner_root[L][8] | last_coeff[4] | pow_nonce {hi, lo}
 *   qvals    : per query { trace_vals[C], cp_vals[16] }
 *   trace_sib: [Q][G][8]         (leaf -> root order, merkle.simf:39-44)
 *   cp_sib   : [Q][G][8]
 *   fri_wit  : [L+1][Q][4]
 *   fri_sib  : layer l = 0..L : [Q][G-1-l][8]
 * It mirrors the witness tuple of stwo-verifier/src/main.simf:9-25 (COMMITMENTS,
 * DECOMMITMENTS, OODS_EVALS, FRI_COMMITMENTS, FRI_DECOMMITMENTS, POW_NONCE). */
typedef struct ssym_stwo_layout {
    uint32_t off_commit, off_oods_trace, off_oods_cp, off_fri_first_root;
    uint32_t off_fri_inner_root, off_last_coeff, off_pow_nonce;
    uint32_t off_qvals, off_trace_sib, off_cp_sib, off_fri_wit;
    uint32_t off_fri_sib[SSYM_MAX_FRI_LAYERS];
    uint32_t stride_words;    /* distance between consecutive proofs, in u32 words  */
    uint32_t algorithmic_bytes; /* payload bytes without alignment padding          */
} ssym_stwo_layout_t;        /* all offsets in u32 words from the proof's base      */

int ssym_stwo_layout(const ssym_stwo_config_t *cfg, ssym_stwo_layout_t *out);

/* Per-proof status word: 0 = accept.  Bits are ordered by the reference's
 * program order (verifier.simf:32-58) at stage granularity. */
#define SSYM_ST_DRAW_EXHAUSTED (1u << 0)   /* channel.simf:125,132 unwrap_left after 256 tries */
#define SSYM_ST_OODS_INV_ZERO (1u << 1)    /* m31.simf:118-122 inside channel.simf:146 / wide_fibonacci.simf:61 */
#define SSYM_ST_OODS_CP_MISMATCH (1u << 2) /* deep/oods.simf:58 */
#define SSYM_ST_POW_FAIL (1u << 3)         /* pow.simf:32 */
#define SSYM_ST_TRACE_MERKLE (1u << 4)     /* evals/verify.simf:56 -> merkle.simf:42-43 */
#define SSYM_ST_CP_MERKLE (1u << 5)        /* evals/verify.simf:66 -> merkle.simf:42-43 */
#define SSYM_ST_ANSWER_INV_ZERO (1u << 6)  /* deep/quotients.simf:22 */
#define SSYM_ST_FRI_MERKLE(l) (1u << (7 + (l))) /* fri/layers.simf:47, layer l = 0..8 */
#define SSYM_ST_FOLD_INV_ZERO (1u << 16)   /* fri/folding.simf:20,34 */
#define SSYM_ST_FINAL_LOG (1u << 17)       /* fri/verify.simf:127  (REF_LITERAL only) */
#define SSYM_ST_LAST_QUERY (1u << 18)      /* fri/layers.simf:75   (REF_LITERAL only) */
#define SSYM_ST_LAST_EVAL (1u << 19)       /* fri/layers.simf:76 */
#define SSYM_ST_SHAPE (1u << 31)           /* witness shape cannot satisfy merkle.simf:42 / the types */

/* Optional per-proof trace: every value-bearing intermediate of verify_proof,
 * for bit-exact parity diffs against the oracle (the reference's equivalent is
 * `simfony debug` + Tracker, simfony-cli/src/tracker.rs:48-80). */
typedef struct ssym_stwo_trace {
    uint32_t status;
    uint32_t first_fail; /* 0 = accept, else (stage<<16 | layer<<8 | query), stage = bit index of status */
    uint32_t digest_commit[8]; /* after evals_commit      evals/commit.simf:20-35 */
    uint32_t cp_alpha[4];
    uint32_t oods_x[4], oods_y[4]; /* channel.simf:143-150 */
    uint32_t cp_eval[4];           /* wide_fibonacci.simf:56-62 */
    uint32_t cp_sampled[4];        /* composition_poly.simf:47-59 */
    uint32_t digest_oods[8];       /* after channel_mix_oods_evals deep/oods.simf:23-39 */
    uint32_t deep_alpha[4];
    uint32_t fri_alpha[SSYM_MAX_FRI_LAYERS][4]; /* fri/commit.simf:36-45 */
    uint32_t digest_fri[8];                     /* after fri_commit fri/commit.simf:72-85 */
    uint32_t digest_pow[8];                     /* after check_proof_of_work pow.simf:22-35 */
    uint32_t pow_value[2];                      /* {hi, lo} of the compared u64 */
    uint32_t queries[SSYM_MAX_QUERIES];         /* fri/queries.simf:30-43 */
    uint32_t fri_answer[SSYM_MAX_QUERIES][4];   /* fri/answers.simf:97-129 */
    uint32_t folded[SSYM_MAX_FRI_LAYERS][SSYM_MAX_QUERIES][4]; /* fri/folding.simf:15-41 */
    uint32_t trace_root[SSYM_MAX_QUERIES][8];   /* recomputed roots, merkle.simf:41 */
    uint32_t cp_root[SSYM_MAX_QUERIES][8];
    uint32_t fri_root[SSYM_MAX_FRI_LAYERS][SSYM_MAX_QUERIES][8];
    uint32_t mask_trace, mask_cp, mask_answer_inv; /* bit q = query q failed that check */
    uint32_t mask_fri[SSYM_MAX_FRI_LAYERS];
    uint32_t mask_fold_inv[SSYM_MAX_FRI_LAYERS];
    uint32_t mask_last_query, mask_last_eval;
    uint32_t draw_retries; /* felt draws that had to be repeated (channel.simf:115-141: a word >= 2p, probability 2^-29 each) */
    uint32_t n_queries_used; /* U: = n_queries, or the number of distinct queries under SSYM_MODE_QUERY_DEDUP (queries[0..U) sorted, the rest 0) */
    uint32_t pad_[1];
} ssym_stwo_trace_t;

/* ------------------------------------------------------------------------- */
/* Cost model                                               */
/* ------------------------------------------------------------------------- */

/* What the reference PROGRAM executes for one proof — the dynamic counterpart of the static `Node bounds` that `simfony run`
 * prints (simfony-cli/src/main.rs:142-154,193-203) — counted at the level of the functions that carry the cost:
 *  - the sha_256_ctx_8_* jets (calls per jet, message bytes, compression-function invocations inside them);
 *  - the M31 operations (fields/m31.simf): m31_mul = 1 multiply_32 + 1 modulo_64, m31_add (counting the add inside m31_sub) = 1 add_32 +
 *    1 modulo_32, m31_neg = 1 subtract_32, m31_inv = 1 is_zero_32 + 37 m31_mul (those 37 are INCLUDED in m31_mul);
 *  - eq_256 (Merkle root comparisons), index -> point conversions (groups/m31_point.simf:58-106: 32 doublings + one point addition per
 *    set bit of the index, which is what makes m31_mul / m31_add depend on the drawn queries).
 * The O(queries x layers) control jets (eq_32, and_32, shifts, divide_32 / divides_32 of the index algebra) are not modelled.
 * The numbers are those of a proof on which no inversion meets zero (status bits OODS_INV_ZERO / ANSWER_INV_ZERO / FOLD_INV_ZERO clear);
 * the reference's early exit at the first failed assert is NOT modelled: this is the cost of running verify_proof to its end.
 * Closed form in csrc/cost.cpp, function by function from the .simf sources; tests/test_cost_model.py compares every field with the
 * oracle's per-thread counters on the fixtures, on random configurations and on random queries. */
#define SSYM_COST_FIELDS 14
typedef struct ssym_cost {
    uint64_t sha_compressions;
    uint64_t sha_init, sha_add_4, sha_add_8, sha_add_32, sha_finalize; /* calls of sha_256_ctx_8_{init,add_4,add_8,add_32,finalize} */
    uint64_t sha_bytes;
    uint64_t m31_mul, m31_add, m31_neg, m31_inv;
    uint64_t eq_256;
    uint64_t point_from_index;
    uint64_t draw_retries;
} ssym_cost_t;
/* Cost of verify_proof (stwo-verifier/src/verifier.simf:32-58) for one proof of configuration `cfg` whose transcript drew `queries`
 * (the first `n_queries_used` words are used — ssym_stwo_trace_t.queries / .n_queries_used; = cfg->n_queries without SSYM_MODE_QUERY_DEDUP) and
 * repeated `draw_retries` felt draws.  Host-only, no GPU involved. */
int ssym_stwo_cost(const ssym_stwo_config_t *cfg, const uint32_t *queries, uint32_t n_queries_used, uint32_t draw_retries, ssym_cost_t *out);

/* ------------------------------------------------------------------------- */
/* stark101 (stark101/src/verifier.simf:17-42)                                 */
/* ------------------------------------------------------------------------- */

#define SSYM_S101_MAX_LIST 31 /* List<_, 32> holds 0..31 items (merkle.simf:19, fri.simf:49) */

/* Packed stark101 proof: a variable-length record of u32 words.
 *   [0]      total_words of this record (including this header)
 *   [1]      n_layers (0..31)
 *   [2..4]   sibling counts of the three trace decommitments (f(x), f(gx), f(g^2 x))
 *   [5]      fri_last_layer
 *   [6]      query ordinal k (0..SSYM_S101_MAX_ORDINAL; 0 = the reference's program): the query index is the (k+1)-th
 *            channel_draw_32(state, DOMAIN_EX_SIZE) after the commitments (verifier.simf:32 draws the first; channel.simf:102-105: every
 *            draw re-hashes the state).  See ssym_stark101_verify_multi_batch.
 *   [7]      reserved (0)
 *   [8..15]  p_mt_root
 *   [16..18] f(x), f(gx), f(g^2 x)      [19] reserved
 *   [20 ..]  siblings of eval 0, eval 1, eval 2 (8 words each, leaf -> root)
 *   then per FRI layer: root[8] | beta | cpa | cpb | n_sib_a | n_sib_b | 0 0 0 |
 *                       siblings a | siblings b
 * mirroring witnesses P_MT_ROOT, P_EVALS, FRI_LAYERS, FRI_LAST_LAYER
 * (stark101/src/main.simf:12-20).  A batch is the concatenation of records plus
 * an offsets array (u64 word offsets, n+1 entries). */
#define SSYM_S101_ST_TRACE_MERKLE(i) (1u << (i))     /* air.simf:41, eval i = 0..2 */
#define SSYM_S101_ST_BETA (1u << 3)                  /* fri.simf:43 */
#define SSYM_S101_ST_DIV (1u << 4)                   /* field.simf:46 gcd != 1 / loop exhausted */
#define SSYM_S101_ST_LAYER_CP (1u << 5)              /* fri.simf:77 */
#define SSYM_S101_ST_LAYER_MERKLE_A (1u << 6)        /* fri.simf:79 */
#define SSYM_S101_ST_LAYER_MERKLE_B (1u << 7)        /* fri.simf:80 */
#define SSYM_S101_ST_LAST (1u << 8)                  /* fri.simf:90 */
#define SSYM_S101_ST_GROUP (1u << 9)                 /* multi-query: ordinal != slot, or commitments differ from the proof's first record */
#define SSYM_S101_MAX_ORDINAL 255u
#define SSYM_S101_ST_SHAPE (1u << 31)

typedef struct ssym_s101_trace {
    uint32_t status;
    uint32_t first_fail_layer; /* first FRI layer with any failing check, or 0xffffffff */
    uint32_t alpha[3];         /* air.simf:30-36 */
    uint32_t idx;              /* verifier.simf:32 */
    uint32_t x;                /* air.simf:58-60 */
    uint32_t cp0;              /* air.simf:94-101 */
    uint32_t n_layers;
    uint32_t beta_drawn[SSYM_S101_MAX_LIST]; /* fri.simf:42 */
    uint32_t cp_ev[SSYM_S101_MAX_LIST + 1];  /* cp value entering layer i; [n_layers] = final */
    uint32_t layer_mask[SSYM_S101_MAX_LIST]; /* per layer: bit0 cp!=cpa, bit1 merkle a, bit2 merkle b, bit3 beta */
    uint32_t state_final[8];                 /* channel state after the three evaluations are mixed */
    uint32_t trace_root[3][8];               /* recomputed roots of the three trace decommitments */
    uint32_t query_ordinal;                  /* record word 6 */
    uint32_t commit_state[8];                /* channel state after fri_read_commitments_32 + the last layer (verifier.simf:30), before any query draw */
} ssym_s101_trace_t;

/* ------------------------------------------------------------------------- */
/* Handle                                                                     */
/* ------------------------------------------------------------------------- */

typedef struct ssym_ctx ssym_ctx_t;

/* One handle per GPU; thread-safe per handle (one host thread per GPU). */
int ssym_create(int device, ssym_ctx_t **out);
void ssym_destroy(ssym_ctx_t *ctx);
const char *ssym_last_error(void);
const char *ssym_version(void);
/* Stream ordering (read this before passing device buffers).  SSYM_MEM_DEVICE calls are ASYNCHRONOUS and run on the handle's stream,
 * which by default is a stream of its own created with cudaStreamNonBlocking: work on it is NOT ordered after the caller's other
 * streams, not even after the legacy default stream.  A caller that fills the input buffers with kernels or copies on its own stream
 * (torch, thrust, ...) must therefore either
 *   - hand that stream to the handle: ssym_set_stream(ctx, stream) — every later device-resident call is then enqueued on it (after
 *     the producers already in it, before whatever the caller enqueues next); cudaStreamLegacy / cudaStreamPerThread name the default
 *     streams; or
 *   - synchronise the producers itself (cudaStreamSynchronize / an event the handle's stream cannot see is NOT enough) and call
 *     ssym_synchronize(ctx) before reading the outputs.
 * The Python binding does the first automatically for torch tensors (Verifier._space: torch's current stream).  NULL restores the
 * handle's own stream.  SSYM_MEM_HOST calls are synchronous unless ssym_set_host_async is on. */
int ssym_set_stream(ssym_ctx_t *ctx, void *cuda_stream);
/* Page-locked host memory for SSYM_MEM_HOST buffers (packed proofs, witness text): copies from it run at the link rate and, in the
 * asynchronous host mode, truly overlap.  Plain malloc'ed buffers are accepted everywhere too, only slower.  Returns NULL on failure. */
void *ssym_pinned_alloc(size_t bytes);
void ssym_pinned_free(void *ptr);
int ssym_synchronize(ssym_ctx_t *ctx);
/* Pipeline depth D (1..8, default 1) for device-resident ssym_stwo_verify_batch calls: call k runs on internal
 * stream k % D (forked from the handle's stream at call time), so up to D consecutive batches are in flight and the
 * latency-bound channel kernel of one batch overlaps the Merkle kernel of the previous one (the channel and query
 * kernels of a pipelined call run on a high-priority stream so they are dispatched ahead of pending Merkle CTAs).  With D > 1 results are
 * ordered into the handle's stream only by ssym_join (device-side wait, no host sync) or ssym_synchronize, and the
 * caller must not reuse an input / output buffer within D consecutive calls without a join in between. */
int ssym_set_pipeline_depth(ssym_ctx_t *ctx, int depth);
/* Asynchronous SSYM_MEM_HOST mode for ssym_stwo_verify_batch (default off = the call returns with the results in place).
 * With on = 1 a host-buffer call only ENQUEUES its chunked H2D copies, kernels and D2H copies and returns; consecutive calls
 * then overlap (the H2D and the kernels of call k+1 run under the kernels of call k: the device-side result buffers are a ring over four calls),
 * and every output is valid after ssym_synchronize.
 * The caller's buffers must be pinned (cudaHostAlloc / torch pin_memory) and must not be touched until then. */
int ssym_set_host_async(ssym_ctx_t *ctx, int on);
/* Witness texts the GPU tokeniser does not keep on its fast path (JSON escapes, `_` separators, upper-case hex, redundant parentheses,
 * another list length, anything malformed) are re-read by the host parser, whose verdict is the call's (default, on = 1).  on = 0
 * makes ssym_stwo_*_wit_batch / ssym_stark101_verify_wit_batch GPU-only: such a witness keeps the flag SSYM_WIT_SLOW and is reported as
 * rejected without any host parsing — a strict mode for inputs known to come from the reference's generators, and how the tests see
 * which formattings stay on the GPU. */
int ssym_set_wit_host_fallback(ssym_ctx_t *ctx, int on);
/* Merkle schedule of ssym_stwo_verify_batch.  The reference hashes every query's path to the root on its own (merkle.simf:39-44); paths of
 * one tree that have met run through the same nodes from there on.  policy 0: one hash chain per query, as the reference; 2: every distinct
 * node is hashed once and a query takes over another's nodes only after a bitwise comparison of everything its own computation would have
 * read (per-query roots, fail masks and verdicts are identical by construction; csrc/stwo_kernels.cuh StwoDedup); 1 (default): policy 2
 * under SSYM_MODE_PROVER_CONSISTENT, policy 0 under SSYM_MODE_REF_LITERAL, where no FRI path can share (finding F1) and planning does not
 * pay.  The default can also be set with the environment variable SSYM_MERKLE_DEDUP. */
int ssym_set_merkle_sharing(ssym_ctx_t *ctx, int policy);
int ssym_join(ssym_ctx_t *ctx);
/* Number of kernels this handle has launched so far (bench.py's gpu_launches). */
uint64_t ssym_launch_count(const ssym_ctx_t *ctx);

/* Per-kernel device timing (CUDA events recorded around every verifier kernel on the launching stream).
 * Kernel ids: 0 stwo channel, 1 stwo query, 2 stwo merkle, 3 stwo finalize, 4 stark101 transcript,
 * 5 stark101 merkle, 6 stark101 finalize.  ssym_profile_read synchronises the stream, adds up the elapsed
 * time and launch count per kernel id since the last read (arrays of SSYM_PROFILE_KERNELS) and resets. */
#define SSYM_PROFILE_KERNELS 8
int ssym_profile_enable(ssym_ctx_t *ctx, int on);
int ssym_profile_read(ssym_ctx_t *ctx, double *ms_per_kernel, uint64_t *launches_per_kernel);

/* ------------------------------------------------------------------------- */
/* Whole-proof batch verification                                             */
/* ------------------------------------------------------------------------- */

/* verify_proof for n packed Stwo proofs (stwo-verifier/src/verifier.simf:32-58).
 *  packed      : n * layout.stride_words u32 words
 *  accept_bits : (n+31)/32 u32 words; bit i = 1 iff proof i is accepted
 *  status      : NULL or n status words (SSYM_ST_*)
 *  trace       : NULL or n ssym_stwo_trace_t
 * All four pointers live in `memspace`.  Asynchronous on the handle's stream
 * for SSYM_MEM_DEVICE — see "Stream ordering" at ssym_set_stream: the inputs must have been produced on (or ordered into) that stream,
 * and with ssym_set_pipeline_depth > 1 the outputs are ordered into it only by ssym_join / ssym_synchronize; synchronous for
 * SSYM_MEM_HOST (copies included). */
int ssym_stwo_verify_batch(ssym_ctx_t *ctx, const ssym_stwo_config_t *cfg, const uint32_t *packed,
                           size_t n, uint32_t *accept_bits, uint32_t *status,
                           ssym_stwo_trace_t *trace, int memspace);

/* ---- compact transport form of packed Stwo proofs ---------------------------
 * 92 % of a packed proof is authentication paths, and the reference's witness repeats every node that several of a
 * tree's paths run through (merkle.simf:39-44 takes one full path per query).  The compact record stores, per tree, each
 * distinct 32-byte sibling once plus one bit per path slot and one index per repeated slot; expanding it gives back the packed record bit for bit,
 * for ANY packed record (equal digests are found by comparing data, nothing is assumed about the proof being honest).
 * It is what host buffers should hold when the host link is the bottleneck (DESIGN.md "Host buffers").
 * Record (u32 words, a multiple of 8; S = sibling slots of a proof in packed order: trace [Q][G], composition [Q][G], FRI layer l [Q][G-1-l]):
 *   [0] record length in words   [1] D = digests in the table   [2] SSYM_COMPACT_MAGIC   [3] R = S - D back references   [4 .. 8) 0
 *   packed words [0, off_trace_sib)                  (header + per-query values)
 *   packed fri_wit section, padded to 8 words
 *   S bits, slot s = bit s & 31 of word s >> 5: 1 = the slot's sibling is the next digest of the table (digests are stored in the order
 *     of their first slot), 0 = it repeats an earlier sibling of the same tree; padded to 8 words
 *   R back references in slot order: the digest's position among its tree's table entries; 1 byte each if Q * G <= 256, else 2; padded to 8 words
 *   D digests of 8 words
 * Record offsets are multiples of 8 words (ssym_stwo_compact_pack produces them so); a blob in device memory is 16-byte aligned. */
#define SSYM_COMPACT_MAGIC 0x32435353u  /* "SSC2" */
#define SSYM_COMPACT_MAGIC3 0x33435353u /* "SSC3": version 3, below */
/* Words an n-proof compact blob can need at most (no two siblings equal). */
size_t ssym_stwo_compact_bound(const ssym_stwo_config_t *cfg, size_t n);
/* Host function: n packed records -> compact records, concatenated in `out`; offsets[0..n] (u32-word offsets, offsets[0] = 0). */
int ssym_stwo_compact_pack(const ssym_stwo_config_t *cfg, const uint32_t *packed, size_t n, uint32_t *out,
                           size_t out_cap_words, uint64_t *offsets);
/* Version 3 ("SSC3"): siblings another query computes are left out.  Where two paths of a tree meet, the sibling each of them needs at that
 * level is the node the OTHER path has just computed — upstream stwo's decommitments leave those out (its `hash_witness` is minimal; the fork
 * that wrote the reference's fixtures expanded them, stwo-verifier/scripts/generate_wit.py:38-42,150-160).  A version 3 record marks such a
 * slot "derived" and stores one byte, the query whose path supplies the node:
 *   [2] SSYM_COMPACT_MAGIC3   [4] X = derived slots (S = D + R + X)   [5] cfg->mode the record was packed under   [6 .. 8) 0
 *   a second S-bit bitmap behind the first (1 = derived; never set together with "new"); the X partner bytes behind the R back references.
 * What a derived slot expands to is DEFINED as the node verify_proof computes on that path from this very record — from the queries its
 * transcript draws, the queried values (trace / composition trees) or the evaluations of fri_answer and the folds (FRI trees), and the
 * siblings below — so expansion is a function of the record and of the semantics alone, and ssym_stwo_compact_pack_hinted only marks a slot
 * whose stored digest IS that node (ssym_stwo_compact_hints compares them on the GPU): lossless for any record, honest or not.  FRI leaves
 * are hashes of COMPUTED evaluations, so the derived slots of a record are expanded under the mode it was packed under ([5]), whatever mode it
 * is then verified under: from HOST buffers ssym_stwo_compact_expand / ssym_stwo_verify_compact_batch read that word (all records with derived
 * slots of one call must carry the same one; another is reported malformed) and, where it differs from cfg->mode, verification takes two
 * passes over the kernels — complete the records under their own mode, then verify under the call's.  When only the semantics differ (same
 * flags: the transcript does not depend on the semantics) the first pass is the records' FRI evaluations and FRI chains alone, on the one
 * transcript both passes share: records packed under PROVER_CONSISTENT (the smallest form) verify under REF_LITERAL at the rate of the host
 * link, with the statuses of the packed path.  Records in DEVICE memory are expanded
 * under cfg->mode only (a record with derived slots of another mode is reported malformed).  A proof packed under REF_LITERAL has derivable
 * siblings in its trace / composition trees only (its FRI paths do not verify, finding F1): pack under PROVER_CONSISTENT.  Requires 32 % Q == 0
 * (the Merkle kernel resolves a derived sibling with a warp shuffle between the Q chains of a tree); otherwise X = 0.
 * ssym_stwo_compact_hints: hints[i * S + s] = the lowest query whose node equals sibling slot s of proof i, or 0xff (S = the slots of a proof
 * in record order).  ssym_stwo_compact_pack_hinted: as ssym_stwo_compact_pack, hints == NULL gives version 2 records.
 * ssym_stwo_compact_pack_gpu = the two together for packed records in HOST memory (the GPU does the hashing, the host the assembly). */
int ssym_stwo_compact_hints(ssym_ctx_t *ctx, const ssym_stwo_config_t *cfg, const uint32_t *packed, size_t n, uint8_t *hints, int memspace);
int ssym_stwo_compact_pack_hinted(const ssym_stwo_config_t *cfg, const uint32_t *packed, const uint8_t *hints, size_t n, uint32_t *out,
                                  size_t out_cap_words, uint64_t *offsets);
int ssym_stwo_compact_pack_gpu(ssym_ctx_t *ctx, const ssym_stwo_config_t *cfg, const uint32_t *packed, size_t n, uint32_t *out,
                               size_t out_cap_words, uint64_t *offsets);
/* GPU: compact -> packed (n * layout.stride_words words).  flags: NULL or n words, 1 where a record is malformed (wrong length /
 * magic / counts that do not match the bitmap / a back reference that does not point to an earlier digest of its tree); such a record expands to zeros.  blob holds offsets[n] words. */
int ssym_stwo_compact_expand(ssym_ctx_t *ctx, const ssym_stwo_config_t *cfg, const uint32_t *blob, const uint64_t *offsets,
                             size_t n, uint32_t *packed_out, uint32_t *flags, int memspace);
/* ssym_stwo_verify_batch on compact records: expanded on the GPU, then verified; a malformed record is rejected with
 * SSYM_ST_SHAPE.  SSYM_MEM_HOST: the compact bytes are what crosses the host link (chunked, double-buffered, and
 * enqueue-only under ssym_set_host_async, like ssym_stwo_verify_batch). */
int ssym_stwo_verify_compact_batch(ssym_ctx_t *ctx, const ssym_stwo_config_t *cfg, const uint32_t *blob,
                                   const uint64_t *offsets, size_t n, uint32_t *accept_bits, uint32_t *status, int memspace);

/* verify_proof for n packed stark101 proofs (stark101/src/verifier.simf:24-42).
 *  blob / offsets : concatenated records and n+1 word offsets */
int ssym_stark101_verify_batch(ssym_ctx_t *ctx, const uint32_t *blob, const uint64_t *offsets,
                               size_t n, uint32_t *accept_bits, uint32_t *status,
                               ssym_s101_trace_t *trace, int memspace);

/* Multi-query stark101.  The reference draws ONE query (verifier.simf:32; prover.py:138): ~3 bits of soundness at
 * rate 1/8.  A Q-query proof is Q records of the reference's own witness shape — same P_MT_ROOT, FRI roots, betas and last layer, query phase
 * k (P_EVALS and the FRI decommitments) in record k — and record k is verified by verify_proof with its query index taken from the (k+1)-th
 * draw (record word 6 = k): the query phase of verifier.simf:32-41 repeated on one channel.  Records are proof-major: record i * Q + k is
 * query k of proof i.  A proof is accepted iff every one of its records is, carries ordinal k in slot k, and reaches the query phase in the
 * channel state of the proof's first record (= identical commitments, bound by SHA-256); a record that breaks the last two gets
 * SSYM_S101_ST_GROUP.  accept_bits: one bit per PROOF (ceil(n_proofs / 32) words); status / trace: one per RECORD (n_proofs * n_queries).
 * tests/golden/stark101_multiquery.json: the reference's prover run unmodified, its trees asked for three more positions. */
int ssym_stark101_verify_multi_batch(ssym_ctx_t *ctx, const uint32_t *blob, const uint64_t *offsets, size_t n_proofs, uint32_t n_queries,
                                     uint32_t *accept_bits, uint32_t *status, ssym_s101_trace_t *trace, int memspace);

/* ------------------------------------------------------------------------- */
/* Batched prover for the AIR verify_proof checks */
/* ------------------------------------------------------------------------- */

/* Proves n instances of the wide-Fibonacci AIR of stwo-verifier/src/constraints/wide_fibonacci.simf:24-62
 * (trace of 2^trace_log rows: c0 = 1, c1 = SplitMix64(seed, row) mod p, c2 = c0^2 + c1^2, c3 = c1^2 + c2^2) and
 * writes n packed proofs (ssym_stwo_layout) that ssym_stwo_verify_batch accepts in SSYM_MODE_PROVER_CONSISTENT:
 * trace / composition commitments (evals/commit.simf:20-35), samples at the OODS point and its double
 * (deep/oods.simf:44-64, evals/composition_poly.simf:47-59), DEEP quotient (fri/answers.simf:97-129 with
 * Appendix A item 1), circle + line FRI layers (fri/commit.simf:72-85, fri/folding.simf:15-41), proof of work
 * (pow.simf:22-35; the smallest passing nonce) and the decommitments of the drawn queries.  The reference has no
 * prover (its fixtures come from an external fork): this exists so that BASELINE configs 3 and 5 can run on
 * DISTINCT proofs.  Requires n_fri_layers == trace_log - 1 and trace_log < lde_log <= 13 (both presets).
 *  seeds      : n u64, in `memspace`
 *  packed_out : n * layout.stride_words u32 words, in `memspace`
 * Synchronous (returns after the proofs are written).  cfg->mode is ignored. */
int ssym_stwo_prove_batch(ssym_ctx_t *ctx, const ssym_stwo_config_t *cfg, const uint64_t *seeds, size_t n,
                          uint32_t *packed_out, int memspace);

/* ------------------------------------------------------------------------- */
/* Element-wise jets / .simf functions (parity + config-4 microbenchmarks)    */
/* All arrays have n elements (CM31 = 2 words, QM31 = 4 words per element,     */
/* interleaved as the .simf tuples are).  `fail` (NULL or n bytes) receives 1   */
/* where the .simf function would hit assert!(false) (inverse of bitwise 0).   */
/* ------------------------------------------------------------------------- */
int ssym_m31_add(ssym_ctx_t *, const uint32_t *a, const uint32_t *b, uint32_t *out, size_t n, int memspace); /* fields/m31.simf:22-26 */
int ssym_m31_sub(ssym_ctx_t *, const uint32_t *a, const uint32_t *b, uint32_t *out, size_t n, int memspace); /* fields/m31.simf:35-37 */
int ssym_m31_neg(ssym_ctx_t *, const uint32_t *a, uint32_t *out, size_t n, int memspace);                   /* fields/m31.simf:29-32 */
int ssym_m31_mul(ssym_ctx_t *, const uint32_t *a, const uint32_t *b, uint32_t *out, size_t n, int memspace); /* fields/m31.simf:40-45 */
int ssym_m31_inv(ssym_ctx_t *, const uint32_t *a, uint32_t *out, uint8_t *fail, size_t n, int memspace);    /* fields/m31.simf:117-132 */
int ssym_cm31_mul(ssym_ctx_t *, const uint32_t *a, const uint32_t *b, uint32_t *out, size_t n, int memspace); /* fields/cm31.simf:79-86 */
int ssym_cm31_inv(ssym_ctx_t *, const uint32_t *a, uint32_t *out, uint8_t *fail, size_t n, int memspace);     /* fields/cm31.simf:88-93 */
int ssym_qm31_add(ssym_ctx_t *, const uint32_t *a, const uint32_t *b, uint32_t *out, size_t n, int memspace); /* fields/qm31.simf:36-40 */
int ssym_qm31_sub(ssym_ctx_t *, const uint32_t *a, const uint32_t *b, uint32_t *out, size_t n, int memspace); /* fields/qm31.simf:49-53 */
int ssym_qm31_mul(ssym_ctx_t *, const uint32_t *a, const uint32_t *b, uint32_t *out, size_t n, int memspace); /* fields/qm31.simf:73-80 */
int ssym_qm31_inv(ssym_ctx_t *, const uint32_t *a, uint32_t *out, uint8_t *fail, size_t n, int memspace);     /* fields/qm31.simf:87-98 */
int ssym_qm31_mul_m31(ssym_ctx_t *, const uint32_t *a, const uint32_t *b, uint32_t *out, size_t n, int memspace);  /* fields/qm31.simf:56-59 */
int ssym_qm31_mul_cm31(ssym_ctx_t *, const uint32_t *a, const uint32_t *b, uint32_t *out, size_t n, int memspace); /* fields/qm31.simf:62-65 */

/* circle_point_index_to_m31_point (groups/m31_point.simf:103-106): out = n * {x, y}. */
int ssym_circle_point(ssym_ctx_t *, const uint32_t *index, uint32_t *out_xy, size_t n, int memspace);

/* circle_fold / line_fold (fri/folding.simf:15-41): per element a left-leaf
 * position, f(p), f(-p) and the fold alpha; one log_size for the whole call. */
int ssym_circle_fold(ssym_ctx_t *, const uint32_t *position, const uint32_t *f_p, const uint32_t *f_neg_p,
                     const uint32_t *alpha, uint32_t log_size, uint32_t *out, uint8_t *fail, size_t n, int memspace);
int ssym_line_fold(ssym_ctx_t *, const uint32_t *position, const uint32_t *f_p, const uint32_t *f_neg_p,
                   const uint32_t *alpha, uint32_t log_size, uint32_t *out, uint8_t *fail, size_t n, int memspace);

/* sha256_pair (hasher.simf:27-32): out[i] = SHA-256(left[i] || right[i]). */
int ssym_sha256_pair(ssym_ctx_t *, const uint32_t *left, const uint32_t *right, uint32_t *out, size_t n, int memspace);

/* merkle_verify_32 without the asserts (merkle.simf:39-44): n paths of the same
 * depth; siblings[i][depth][8]; writes the recomputed root and the final `path`
 * (== 1 iff merkle.simf:42 holds).  ok_bits (NULL or (n+31)/32 words): bit i = 1
 * iff path == 1 and root == expected_root[i] (expected_root may be NULL). */
int ssym_merkle_root_from_path(ssym_ctx_t *, const uint32_t *leaf, const uint32_t *auth_path,
                               const uint32_t *siblings, uint32_t depth, const uint32_t *expected_root,
                               uint32_t *out_root, uint32_t *out_path, uint32_t *ok_bits, size_t n, int memspace);

/* Channel transitions (channel.simf:31-172, fri/queries.simf:14-43).  A channel
 * state is 9 words: digest[8] | n_sent.  In-place on `state`. */
int ssym_channel_mix_u256(ssym_ctx_t *, uint32_t *state, const uint32_t *input, size_t n, int memspace);
int ssym_channel_mix_u64(ssym_ctx_t *, uint32_t *state, const uint32_t *input_hi_lo, size_t n, int memspace);
int ssym_channel_draw_qm31(ssym_ctx_t *, uint32_t *state, uint32_t *out, uint8_t *fail, size_t n, int memspace);
int ssym_channel_draw_queries(ssym_ctx_t *, uint32_t *state, uint32_t log_size, uint32_t n_queries,
                              uint32_t *out /* n * n_queries */, size_t n, int memspace);

/* stark101 field jets (stark101/src/field.simf:14-94), p = 3*2^30 + 1. */
int ssym_s101_mul_mod(ssym_ctx_t *, const uint32_t *a, const uint32_t *b, uint32_t *out, size_t n, int memspace);
int ssym_s101_div_mod(ssym_ctx_t *, const uint32_t *a, const uint32_t *b, uint32_t *out, uint8_t *fail, size_t n, int memspace);

/* Register-only IADD3 / LOP3 / SHF probe: returns measured 32-bit integer
 * ops/s on this device (the INT32 roofline denominator). */
int ssym_int32_peak_probe(ssym_ctx_t *, double *out_ops_per_s, double *out_ms);

/* ------------------------------------------------------------------------- */
/* Witness ingestion (host side; simfony-cli/src/main.rs:77-81 parse_witness)  */
/* ------------------------------------------------------------------------- */

/* Parse the text of a `.wit` JSON file produced by
 * stwo-verifier/scripts/generate_wit.py:106-245 and pack it (one proof).
 * `out` must hold layout.stride_words words.  Returns 0, or SSYM_ERR_PARSE when
 * the text is not a witness of the program's types.  *shape_reject is set to 1
 * when the witness is well-typed but a list length makes merkle.simf:42 fail
 * (the record is then zero-filled and must be reported as rejected). */
int ssym_stwo_pack_wit(const ssym_stwo_config_t *cfg, const char *json_text, size_t len,
                       uint32_t *out, int *shape_reject);

/* Same for stark101 `.wit` (stark101/scripts/generate_wit.py:7-30).  On entry
 * *out_words is the capacity of `out`; on return the record length. */
int ssym_s101_pack_wit(const char *json_text, size_t len, uint32_t *out, size_t *out_words);

/* ------------------------------------------------------------------------- */
/* Witness ingestion on the GPU                  */
/* ------------------------------------------------------------------------- */

/* Per-witness ingestion flag. */
#define SSYM_WIT_OK 0    /* packed                                                              */
#define SSYM_WIT_SHAPE 1 /* well-typed but ill-shaped (= *shape_reject of ssym_stwo_pack_wit)     */
#define SSYM_WIT_PARSE 2 /* not a witness of the program's types (= SSYM_ERR_PARSE)              */
#define SSYM_WIT_SLOW 3  /* text outside the GPU tokeniser's fast path: re-parsed on the host before the call returns, so never seen by callers unless
                            ssym_set_wit_host_fallback(ctx, 0) */

/* n `.wit` JSON texts (the files `simfony run --witness` reads, simfony-cli/src/main.rs:77-81; witness i = bytes
 * [offsets[i], offsets[i+1]) of `text`) -> n packed proofs, tokenised and packed ON THE GPU (one CTA per witness,
 * csrc/wit_kernels.cu).  Same result as n calls of ssym_stwo_pack_wit: flags[i] = SSYM_WIT_OK / _SHAPE / _PARSE, and the
 * record of a flagged witness is zero-filled.  Texts using grammar the generator never emits (JSON escapes, `_` separators,
 * upper-case hex, redundant parentheses, trailing commas) and malformed ones are detected on
 * the GPU and re-parsed by the host parser, so the accepted grammar is exactly that of ssym_stwo_pack_wit.
 *  text / offsets / packed_out (n * stride_words) / flags (n u32) all live in `memspace`.  Synchronous. */
int ssym_stwo_pack_wit_batch(ssym_ctx_t *ctx, const ssym_stwo_config_t *cfg, const char *text, const uint64_t *offsets,
                             size_t n, uint32_t *packed_out, uint32_t *flags, int memspace);

/* The token skeleton the GPU tokeniser checks witness value `name` (0 COMMITMENTS, 1 DECOMMITMENTS, 2 OODS_EVALS,
 * 3 FRI_COMMITMENTS, 4 FRI_DECOMMITMENTS, 5 POW_NONCE; stwo-verifier/src/main.simf:9-25) against, one byte per token
 * ( ) [ ] , L = `list!`  N = integer literal, and for the k-th literal the packed word it lands in | width << 28 (0 u32, 1 u64, 2 u256).
 * On entry *skel_len / *slot_cnt are the capacities, on return the lengths (SSYM_ERR_NOMEM if too small; NULL buffers query the sizes). */
int ssym_stwo_wit_skeleton(const ssym_stwo_config_t *cfg, int name, uint8_t *skel, size_t *skel_len, uint32_t *slots, size_t *slot_cnt);

/* `simfony run verifier --witness w_i` for n witness TEXTS: ssym_stwo_pack_wit_batch + ssym_stwo_verify_batch without the
 * packed proofs ever leaving the GPU.  For SSYM_MEM_HOST the text is streamed in chunks (H2D of chunk k+1 under the
 * tokeniser + verifier kernels of chunk k).  A witness whose flag is not SSYM_WIT_OK is rejected with SSYM_ST_SHAPE set.
 *  accept_bits ((n+31)/32 words), status (NULL or n), flags (NULL or n) live in `memspace`.  Synchronous. */
int ssym_stwo_verify_wit_batch(ssym_ctx_t *ctx, const ssym_stwo_config_t *cfg, const char *text, const uint64_t *offsets,
                               size_t n, uint32_t *accept_bits, uint32_t *status, uint32_t *flags, int memspace);

/* The same for stark101 witness texts (stark101/src/main.simf:12-20, stark101/scripts/generate_wit.py:13-29).  The lists of a stark101
 * witness have no fixed length, so the GPU tokeniser works with the shape (number of FRI layers, every sibling count) of the first witness
 * of the batch the host parser accepts; witnesses of another shape or another formatting, and malformed ones, go through the host parser
 * (a well-typed witness of another shape is verified in a batch of its own), so every verdict is the one ssym_s101_pack_wit +
 * ssym_stark101_verify_batch give.  flags: SSYM_WIT_OK / SSYM_WIT_PARSE. */
int ssym_stark101_verify_wit_batch(ssym_ctx_t *ctx, const char *text, const uint64_t *offsets, size_t n, uint32_t *accept_bits,
                                   uint32_t *status, uint32_t *flags, int memspace);

#if defined(__GNUC__)
#pragma GCC visibility pop
#endif
#ifdef __cplusplus
}
#endif
#endif /* SSYM_H */
