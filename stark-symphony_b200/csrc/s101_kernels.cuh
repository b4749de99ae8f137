// SPDX-License-Identifier: MIT
// Batched stark101 verifier kernels (see s101_kernels.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/ssym.h"
#include "stwo_kernels.cuh"

namespace ssym {

// per-record scratch: layer offsets at [2, 33); the query ordinal and the channel state after the commitments (multi-query grouping)
enum : uint32_t { S101_CTX_NLAYERS = 0, S101_CTX_IDX = 1, S101_CTX_LAYER_OFF = 2, S101_CTX_ORD = 33, S101_CTX_COMMIT = 34, S101_CTX_WORDS = 48 };

struct S101Params {
    const uint32_t *blob;    // concatenated records (include/ssym.h)
    const uint64_t *offsets; // n + 1 word offsets
    uint32_t *ctx;           // n * S101_CTX_WORDS
    uint32_t *status;        // n
    ssym_s101_trace_t *trace; // n or nullptr
    uint32_t n;
    uint32_t max_layers; // upper bound of n_layers over the batch (31 if unknown)
};

void launch_s101_verify(const S101Params &p, uint32_t *accept_bits, cudaStream_t s, uint64_t *launch_counter, Profiler *prof);
// ssym_stark101_verify_multi_batch: p.n = n_proofs * n_queries records (proof-major); accept_bits: one bit per proof
void launch_s101_verify_multi(const S101Params &p, uint32_t n_queries, uint32_t *accept_bits, cudaStream_t s, uint64_t *launch_counter, Profiler *prof);

} // namespace ssym
