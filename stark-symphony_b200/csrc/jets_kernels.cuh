// SPDX-License-Identifier: MIT
// Launchers of the element-wise jet kernels (jets_kernels.cu).
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

namespace ssym {

enum {
    JET_M31_ADD = 0, JET_M31_SUB, JET_M31_NEG, JET_M31_MUL, JET_M31_INV,
    JET_CM31_MUL, JET_CM31_INV,
    JET_QM31_ADD, JET_QM31_SUB, JET_QM31_MUL, JET_QM31_INV, JET_QM31_MUL_M31, JET_QM31_MUL_CM31,
};
enum { CHAN_MIX_U256 = 0, CHAN_MIX_U64, CHAN_DRAW_QM31, CHAN_DRAW_QUERIES };

int launch_field_jet(int op, const uint32_t *a, const uint32_t *b, uint32_t *out, uint8_t *fail, size_t n, cudaStream_t s);
void launch_circle_point(const uint32_t *index, uint32_t *out_xy, size_t n, cudaStream_t s);
void launch_fold_table(bool circle, uint32_t log_size, uint32_t *table, cudaStream_t s);
void launch_fold(bool circle, const uint32_t *position, const uint32_t *f_p, const uint32_t *f_neg_p, const uint32_t *alpha,
                 uint32_t log_size, const uint32_t *table, uint32_t *out, uint8_t *fail, size_t n, cudaStream_t s);
void launch_sha256_pair(const uint32_t *left, const uint32_t *right, uint32_t *out, size_t n, cudaStream_t s);
void launch_merkle_path(const uint32_t *leaf, const uint32_t *auth_path, const uint32_t *siblings, uint32_t depth,
                        const uint32_t *expected_root, uint32_t *out_root, uint32_t *out_path, uint32_t *ok_bits, size_t n,
                        cudaStream_t s);
void launch_channel(int op, uint32_t *state, const uint32_t *input, uint32_t *out, uint8_t *fail, uint32_t log_size,
                    uint32_t n_queries, size_t n, cudaStream_t s);
void launch_s101_field(int op, const uint32_t *a, const uint32_t *b, uint32_t *out, uint8_t *fail, size_t n, cudaStream_t s);
double launch_int32_probe(uint32_t *sink, cudaStream_t s, int *blocks_out);

} // namespace ssym
