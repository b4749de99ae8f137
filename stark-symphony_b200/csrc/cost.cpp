// SPDX-License-Identifier: MIT
//
// Cost model of the stwo-verifier program (include/ssym.h "Cost model"): how many times verify_proof
// (stwo-verifier/src/verifier.simf:32-58) calls the jets / field functions that carry its cost, as a closed form in the configuration and
// in the drawn queries.  Every function below returns the cost of the `.simf` function it is named after, composed exactly as that function
// composes its callees (file:line given); nothing is measured, nothing runs on the GPU.
#include <cstdint>
#include <cstring>

#include "../../include/ssym.h"

namespace {

struct Cost {
    uint64_t v[SSYM_COST_FIELDS] = {0};
    Cost &operator+=(const Cost &o) { for (int i = 0; i < SSYM_COST_FIELDS; i++) v[i] += o.v[i]; return *this; }
    Cost operator+(const Cost &o) const { Cost r = *this; r += o; return r; }
    Cost operator*(uint64_t n) const { Cost r; for (int i = 0; i < SSYM_COST_FIELDS; i++) r.v[i] = v[i] * n; return r; }
};
enum { COMP, INIT, ADD4, ADD8, ADD32, FIN, BYTES, MUL, ADD, NEG, INV, EQ256, POINT, RETRY };
Cost unit(int field, uint64_t n = 1) { Cost c; c.v[field] = n; return c; }

// ---- fields/m31.simf ---------------------------------------------------------------------------------------
Cost m31_add() { return unit(ADD); }                         // :22-26
Cost m31_neg() { return unit(NEG); }                         // :29-32
Cost m31_sub() { return m31_add() + m31_neg(); }             // :35-37
Cost m31_mul() { return unit(MUL); }                         // :40-45
Cost m31_inv() { return unit(INV) + m31_mul() * 37; }        // :117-132: the addition chain a^(p-2) = 30 squarings + 7 products
// ---- fields/cm31.simf --------------------------------------------------------------------------------------
Cost cm31_add() { return m31_add() * 2; }                    // :30-34
Cost cm31_neg() { return m31_neg() * 2; }                    // :37-40
Cost cm31_sub() { return m31_sub() * 2; }                    // :43-47
Cost cm31_sub_m31() { return m31_sub(); }                    // :50-53
Cost cm31_mul_m31() { return m31_mul() * 2; }                // :56-59
Cost cm31_div_m31() { return cm31_mul_m31() + m31_inv(); }   // :62-65
Cost cm31_conj() { return m31_neg(); }                       // :73-76
Cost cm31_mul() { return m31_mul() * 4 + m31_sub() + m31_add(); }                                   // :79-86
Cost cm31_inv() { return cm31_conj() + m31_mul() * 2 + m31_add() + cm31_div_m31(); }                // :88-93
Cost cm31_dbl() { return cm31_add(); }                       // :102-104
// ---- fields/qm31.simf --------------------------------------------------------------------------------------
Cost qm31_add() { return cm31_add() * 2; }                   // :36-40
Cost qm31_sub() { return cm31_sub() * 2; }                   // :49-53
Cost qm31_mul_m31() { return cm31_mul_m31() * 2; }           // :56-59
Cost qm31_mul_cm31() { return cm31_mul() * 2; }              // :62-65
Cost qm31_mul() { return cm31_mul() * 5 + cm31_add() * 2; }  // :73-80: ar*br + (ai*bi)*(2,1), ar*bi + ai*br
Cost qm31_inv() {                                            // :87-98
    return cm31_mul() * 2 + cm31_add() /* ai_sq_dbl */ + m31_neg() /* ai_sq_rev */ + cm31_add() + cm31_neg() + cm31_add() /* den */ + cm31_inv() +
           cm31_mul() + cm31_neg() + cm31_mul();
}
Cost qm31_div() { return qm31_mul() + qm31_inv(); }          // :101-104
// ---- groups/m31_point.simf, qm31_point.simf -------------------------------------------------------------------
Cost m31_point_dbl() { return m31_mul() /* xy */ + (m31_mul() + m31_add() + m31_sub()) /* dbl_x :33-37 */ + m31_add(); } // :49-55
Cost m31_point_add() { return m31_mul() * 4 + m31_sub() + m31_add(); }                                               // :40-46
Cost point_from_index(uint32_t index) {                      // :58-106: 32 LSB-first steps, a point addition per set bit
    return unit(POINT) + m31_point_dbl() * 32 + m31_point_add() * (uint64_t)__builtin_popcount(index);
}
Cost qm31_point_dbl_x() { return qm31_mul() + qm31_add() + qm31_sub(); } // qm31_point.simf:27-31

// ---- groups/coset.simf, circle_domain.simf, line_domain.simf: the index algebra itself (plain integers, no modelled jets) ----
uint32_t bit_reverse_position(uint32_t pos, uint32_t log_size) { // coset.simf:20-25
    uint32_t r = 0;
    for (int i = 0; i < 32; i++) r |= ((pos >> i) & 1u) << (31 - i);
    const uint32_t sh = (32u - log_size) & 0xff;
    return sh >= 32 ? 0 : r >> sh;
}
uint32_t subgroup_gen(uint32_t log_size) { const uint32_t sh = (31u - log_size) & 0xff; return sh >= 32 ? 0 : 1u << sh; } // coset.simf:28-31
uint32_t idx_add(uint32_t l, uint32_t r) { return (l + r) & 0x7fffffffu; }                                               // :34-37
uint32_t idx_mul(uint32_t l, uint32_t r) { return (uint32_t)((uint64_t)l * r) & 0x7fffffffu; }                         // :40-45
uint32_t idx_neg(uint32_t i) { return (0x80000000u - i) & 0x7fffffffu; }                                                 // :48-51
uint32_t shl(uint32_t sh, uint32_t x) { sh &= 0xff; return sh >= 32 ? 0 : x << sh; }
uint32_t circle_position_to_point_index(uint32_t log_size, uint32_t position) { // circle_domain.simf:17-37
    const uint32_t half = shl(log_size - 1, 1), offset = subgroup_gen((log_size + 1) & 0xff), step = subgroup_gen((log_size - 1) & 0xff);
    if (position < half) return idx_add(offset, idx_mul(step, position));
    return idx_neg(idx_add(offset, idx_mul(step, position - half)));
}
uint32_t line_position_to_index(uint32_t log_size, uint32_t position) { // line_domain.simf:18-31
    return idx_add(subgroup_gen((log_size + 2) & 0xff), idx_mul(subgroup_gen(log_size), position));
}

// ---- the sha_256_ctx_8_* jets: one hash of `bytes` message bytes fed by n4 / n8 / n32 add calls ------------------------
Cost hash(uint64_t n4, uint64_t n8, uint64_t n32) {
    const uint64_t bytes = 4 * n4 + 8 * n8 + 32 * n32;
    return unit(INIT) + unit(ADD4, n4) + unit(ADD8, n8) + unit(ADD32, n32) + unit(FIN) + unit(BYTES, bytes) + unit(COMP, (bytes + 8) / 64 + 1);
}
Cost sha256_pair() { return hash(0, 0, 2); }                 // hasher.simf:27-32
Cost hash_node_qm31() { return hash(4, 0, 0); }              // hasher.simf:100-104
Cost merkle_verify_32(uint32_t n_sib) { return sha256_pair() * n_sib + unit(EQ256); } // merkle.simf:22-44
// ---- channel.simf -------------------------------------------------------------------------------------------
Cost channel_draw_u256() { return hash(1, 0, 1); }           // :36-44
Cost channel_draw_qm31() { return channel_draw_u256(); }     // :115-141 (first attempt; repeats are added once, from draw_retries)
Cost channel_mix_u256() { return hash(0, 0, 2); }            // :154-162
Cost channel_mix_u64() { return hash(0, 1, 1); }             // :165-173

// deep/quotients.simf
Cost denominator_inverse() { return cm31_sub_m31() * 2 + cm31_mul() * 2 + cm31_sub() + cm31_inv(); } // :15-22
Cost interpolant_coefficients() {                                                                 // :25-35
    return (cm31_dbl() + cm31_neg()) * 2 + qm31_mul() * 2 + qm31_sub() + qm31_mul() * 3;
}
Cost nominator() { return qm31_mul_m31() * 2 + qm31_add() + qm31_sub(); }                            // :38-44
Cost numerator_aggregate_column() { return interpolant_coefficients() + nominator() + qm31_add() + qm31_mul(); } // fri/answers.simf:40-58
Cost fold(uint32_t index) { // fri/folding.simf:15-41 (circle and line folds cost the same; they differ in the index)
    return point_from_index(index) + m31_inv() + qm31_add() + qm31_sub() + qm31_mul_m31() + qm31_mul() + qm31_add();
}

} // namespace

extern "C" int ssym_stwo_cost(const ssym_stwo_config_t *cfg, const uint32_t *queries, uint32_t n_queries_used, uint32_t draw_retries, ssym_cost_t *out) {
    ssym_stwo_layout_t lo;
    if (!cfg || !queries || !out) return SSYM_ERR_USAGE;
    int rc = ssym_stwo_layout(cfg, &lo); // validates the configuration
    if (rc) return rc;
    const uint32_t QD = cfg->n_queries; // queries DRAWN; Q of them are verified (all, or the distinct ones under SSYM_MODE_QUERY_DEDUP)
    if (n_queries_used > QD || (!(cfg->mode & SSYM_MODE_QUERY_DEDUP) && n_queries_used != QD)) return SSYM_ERR_USAGE;
    const uint32_t Q = n_queries_used, L = cfg->n_fri_layers, G = cfg->lde_log, T = cfg->trace_log, C = SSYM_STWO_COLUMNS(cfg), NCOL = C + SSYM_NUM_CP_PARTITIONS;
    Cost c;
    // evals_commit                                    evals/commit.simf:20-35
    c += channel_mix_u256() * 3 + channel_draw_qm31();
    // oods                                            deep/oods.simf:44-64
    c += channel_draw_qm31();                                                                                      // the OODS parameter t
    c += qm31_mul() + qm31_add() + qm31_inv() + qm31_sub() + qm31_mul() + qm31_add() + qm31_mul();                 // channel.simf:143-151
    c += hash(4 * NCOL, 0, 1);                                                                                     // channel_mix_oods_evals :23-39
    c += (qm31_mul() * 2 + qm31_add() + qm31_sub() + qm31_mul() + qm31_add()) * (C > 2 ? C - 2 : 0);               // wide_fibonacci.simf:24-55
    c += qm31_point_dbl_x() * (((T - 1) & 0xff)) + qm31_div();                                                     // vanishing poly composition_poly.simf:66-71, :61
    c += (qm31_mul() * 3 + qm31_add() * 3) * 4 + qm31_mul() * 4 + qm31_add() * 3;                                  // composition_poly.simf:38-59
    c += channel_draw_qm31();                                                                                      // the DEEP coefficient
    // fri_commit                                      fri/commit.simf:36-85
    c += (channel_mix_u256() + channel_draw_qm31()) * (L + 1) + hash(4, 0, 1);
    // check_proof_of_work                             pow.simf:22-35
    c += channel_mix_u64();
    // fri_generate_queries                            fri/queries.simf:30-43
    c += channel_draw_u256() * ((QD + 7) / 8);
    // repeated felt draws (channel.simf:125-137): each is one more channel_draw_u256
    c += channel_draw_u256() * draw_retries + unit(RETRY, draw_retries);
    // evals_verify                                    evals/verify.simf:50-78: trace leaf + path, composition leaf + path, per query
    c += (hash(C, 0, 0) + merkle_verify_32(G) + hash(SSYM_NUM_CP_PARTITIONS, 0, 0) + merkle_verify_32(G)) * Q;
    for (uint32_t q = 0; q < Q; q++) {
        // fri_answer                                  fri/answers.simf:97-129 / SURVEY Appendix A item 1
        c += point_from_index(circle_position_to_point_index(G, bit_reverse_position(queries[q], G)));
        c += numerator_aggregate_column() * NCOL;
        if (SSYM_MODE_SEMANTICS(cfg->mode) == SSYM_MODE_REF_LITERAL) c += denominator_inverse() + qm31_mul_cm31() + qm31_mul();
        else c += qm31_point_dbl_x() + qm31_mul() + qm31_add() /* 2P */ + denominator_inverse() * 2 + qm31_mul_cm31() * 2 + qm31_add();
        // fri_verify                                  fri/verify.simf:114-129, fri/layers.simf:29-69
        uint32_t fq = queries[q];
        for (uint32_t l = 0; l <= L; l++) {
            const uint32_t log = (G - l) & 0xff, position = fq & ~1u; // adjacent_leaves: the left leaf
            c += hash_node_qm31() * 2 + sha256_pair() + merkle_verify_32(G - 1 - l);
            const uint32_t rev = bit_reverse_position(position, log);
            c += fold(l == 0 ? circle_position_to_point_index(log, rev) : line_position_to_index(log, rev));
            fq = position / 2;
        }
    }
    memcpy(out, c.v, sizeof c.v);
    return SSYM_OK;
}
