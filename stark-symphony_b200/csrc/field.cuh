// SPDX-License-Identifier: MIT
//
// M31 / CM31 / QM31 and circle-group device functions with the EXACT semantics of the reference's
// jets-over-u32 definitions (stwo-verifier/src/fields/*.simf, groups/*.simf), including their
// behaviour on non-canonical inputs: witness values are never canonicalised, `neg`/`conj` return the
// unreduced `p - a`, `add` wraps at 2^32 before reducing, equality is bitwise.  Reduction is the
// Mersenne fold (x & p) + (x >> 31) with one conditional subtract (no Barrett / Montgomery).
#pragma once
#include <stdint.h>

namespace ssym {

#define SSYM_P 2147483647u

typedef uint32_t M31;

// m31(v) = v mod p for any u32 v                                   fields/m31.simf:17-19
__device__ __forceinline__ M31 m31_reduce(uint32_t v) {
    uint32_t r = (v & SSYM_P) + (v >> 31); // <= p + 1
    return r >= SSYM_P ? r - SSYM_P : r;
}
// ((a + b) mod 2^32) mod p                                          fields/m31.simf:22-26
__device__ __forceinline__ M31 m31_add(M31 a, M31 b) { return m31_reduce(a + b); }
// (p - a) mod 2^32, NOT reduced (neg(0) = p)                        fields/m31.simf:29-32
__device__ __forceinline__ M31 m31_neg(M31 a) { return SSYM_P - a; }
// add(a, neg(b))                                                    fields/m31.simf:35-37
__device__ __forceinline__ M31 m31_sub(M31 a, M31 b) { return m31_reduce(a + (SSYM_P - b)); }
// (a * b as u64) mod p, exact for any u32 a, b                      fields/m31.simf:40-45
__device__ __forceinline__ M31 m31_mul(M31 a, M31 b) {
    uint64_t x = (uint64_t)a * b;                      // < 2^64
    uint64_t y = (x & SSYM_P) + (x >> 31);             // < 2^31 + 2^33
    uint32_t z = ((uint32_t)y & SSYM_P) + (uint32_t)(y >> 31); // < 2^31 + 8
    return z >= SSYM_P ? z - SSYM_P : z;
}
__device__ __forceinline__ M31 m31_pow2(M31 a) { return m31_mul(a, a); }
template <int N>
__device__ __forceinline__ M31 m31_sqn(M31 a) {
#pragma unroll
    for (int i = 0; i < N; i++) a = m31_mul(a, a);
    return a;
}
// a^(p-2) by the reference's addition chain; *fail |= (a == 0 bitwise)   fields/m31.simf:117-132
// (the chain maps 0 -> 0, which is the documented continuation value)
__device__ __forceinline__ M31 m31_inv(M31 a, bool &fail) {
    fail = fail || (a == 0);
    M31 t0 = m31_mul(m31_sqn<2>(a), a);      // a^5
    M31 t1 = m31_mul(m31_sqn<1>(t0), t0);    // a^15
    M31 t2 = m31_mul(m31_sqn<3>(t1), t0);    // a^125
    M31 t3 = m31_mul(m31_sqn<1>(t2), t0);    // a^255
    M31 t4 = m31_mul(m31_sqn<8>(t3), t3);    // a^65535
    M31 t5 = m31_mul(m31_sqn<8>(t4), t3);    // a^16777215
    return m31_mul(m31_sqn<7>(t5), t2);      // a^2147483645
}

struct CM31 {
    M31 a, b; // a + b i
};
__device__ __forceinline__ CM31 cm31(M31 a, M31 b) { CM31 r; r.a = a; r.b = b; return r; }
__device__ __forceinline__ CM31 cm31_add(CM31 x, CM31 y) { return cm31(m31_add(x.a, y.a), m31_add(x.b, y.b)); }       // cm31.simf:30-34
__device__ __forceinline__ CM31 cm31_neg(CM31 x) { return cm31(m31_neg(x.a), m31_neg(x.b)); }                         // cm31.simf:37-40
__device__ __forceinline__ CM31 cm31_sub(CM31 x, CM31 y) { return cm31(m31_sub(x.a, y.a), m31_sub(x.b, y.b)); }       // cm31.simf:43-47
__device__ __forceinline__ CM31 cm31_sub_m31(CM31 x, M31 y) { return cm31(m31_sub(x.a, y), x.b); }                    // cm31.simf:50-53
__device__ __forceinline__ CM31 cm31_mul_m31(CM31 x, M31 y) { return cm31(m31_mul(x.a, y), m31_mul(x.b, y)); }        // cm31.simf:56-59
__device__ __forceinline__ CM31 cm31_conj(CM31 x) { return cm31(x.a, m31_neg(x.b)); }                                 // cm31.simf:73-76
__device__ __forceinline__ CM31 cm31_mul(CM31 x, CM31 y) {                                                            // cm31.simf:79-86
    return cm31(m31_sub(m31_mul(x.a, y.a), m31_mul(x.b, y.b)), m31_add(m31_mul(x.a, y.b), m31_mul(x.b, y.a)));
}
__device__ __forceinline__ CM31 cm31_inv(CM31 x, bool &fail) {                                                        // cm31.simf:88-93
    CM31 cj = cm31_conj(x);
    M31 norm = m31_add(m31_pow2(x.a), m31_pow2(x.b));
    return cm31_mul_m31(cj, m31_inv(norm, fail));
}
__device__ __forceinline__ CM31 cm31_dbl(CM31 x) { return cm31_add(x, x); }                                          // cm31.simf:102-104

struct QM31 {
    CM31 r, i; // r + i j,  j^2 = 2 + i
};
__device__ __forceinline__ QM31 qm31(M31 a, M31 b, M31 c, M31 d) { QM31 q; q.r = cm31(a, b); q.i = cm31(c, d); return q; }
__device__ __forceinline__ QM31 qm31c(CM31 r, CM31 i) { QM31 q; q.r = r; q.i = i; return q; }
__device__ __forceinline__ QM31 qm31_zero() { return qm31(0, 0, 0, 0); }
__device__ __forceinline__ QM31 qm31_one() { return qm31(1, 0, 0, 0); }
__device__ __forceinline__ QM31 qm31_load(const uint32_t *w) { return qm31(w[0], w[1], w[2], w[3]); }
__device__ __forceinline__ QM31 qm31_load4(const uint32_t *w) { // 16-byte aligned
    uint4 v = *reinterpret_cast<const uint4 *>(w);
    return qm31(v.x, v.y, v.z, v.w);
}
__device__ __forceinline__ void qm31_store(uint32_t *w, QM31 q) { w[0] = q.r.a; w[1] = q.r.b; w[2] = q.i.a; w[3] = q.i.b; }
__device__ __forceinline__ void qm31_store4(uint32_t *w, QM31 q) { *reinterpret_cast<uint4 *>(w) = make_uint4(q.r.a, q.r.b, q.i.a, q.i.b); }
__device__ __forceinline__ QM31 qm31_add(QM31 x, QM31 y) { return qm31c(cm31_add(x.r, y.r), cm31_add(x.i, y.i)); }            // qm31.simf:36-40
__device__ __forceinline__ QM31 qm31_neg(QM31 x) { return qm31c(cm31_neg(x.r), cm31_neg(x.i)); }                              // qm31.simf:43-46
__device__ __forceinline__ QM31 qm31_sub(QM31 x, QM31 y) { return qm31c(cm31_sub(x.r, y.r), cm31_sub(x.i, y.i)); }            // qm31.simf:49-53
__device__ __forceinline__ QM31 qm31_mul_m31(QM31 x, M31 y) { return qm31c(cm31_mul_m31(x.r, y), cm31_mul_m31(x.i, y)); }     // qm31.simf:56-59
__device__ __forceinline__ QM31 qm31_mul_cm31(QM31 x, CM31 y) { return qm31c(cm31_mul(x.r, y), cm31_mul(x.i, y)); }           // qm31.simf:62-65
__device__ __forceinline__ QM31 qm31_mul(QM31 x, QM31 y) {                                                                    // qm31.simf:73-80
    CM31 re = cm31_add(cm31_mul(x.r, y.r), cm31_mul(cm31_mul(x.i, y.i), cm31(2, 1)));
    CM31 im = cm31_add(cm31_mul(x.r, y.i), cm31_mul(x.i, y.r));
    return qm31c(re, im);
}
__device__ __forceinline__ QM31 qm31_inv(QM31 x, bool &fail) {                                                                // qm31.simf:87-98
    CM31 ar_sq = cm31_mul(x.r, x.r);
    CM31 ai_sq = cm31_mul(x.i, x.i);
    CM31 ai_sq_dbl = cm31_add(ai_sq, ai_sq);
    CM31 ai_sq_rev = cm31(m31_neg(ai_sq.b), ai_sq.a);
    CM31 den = cm31_add(ar_sq, cm31_neg(cm31_add(ai_sq_dbl, ai_sq_rev)));
    CM31 den_inv = cm31_inv(den, fail);
    return qm31c(cm31_mul(x.r, den_inv), cm31_mul(cm31_neg(x.i), den_inv));
}
__device__ __forceinline__ bool qm31_eq(QM31 x, QM31 y) { return x.r.a == y.r.a && x.r.b == y.r.b && x.i.a == y.i.a && x.i.b == y.i.b; } // qm31.simf:117-124 (bitwise)

// ---- circle group over M31                                         groups/m31_point.simf ----
struct M31Point {
    M31 x, y;
};
__device__ __forceinline__ M31Point m31_point(M31 x, M31 y) { M31Point p; p.x = x; p.y = y; return p; }
__device__ __forceinline__ M31 m31_point_dbl_x(M31 x) { M31 s = m31_pow2(x); return m31_sub(m31_add(s, s), 1); }  // m31_point.simf:33-37
__device__ __forceinline__ M31Point m31_point_add(M31Point l, M31Point r) {                                        // m31_point.simf:40-46
    return m31_point(m31_sub(m31_mul(l.x, r.x), m31_mul(l.y, r.y)), m31_add(m31_mul(l.x, r.y), m31_mul(l.y, r.x)));
}
__device__ __forceinline__ M31Point m31_point_dbl(M31Point p) {                                                    // m31_point.simf:49-55
    M31 xy = m31_mul(p.x, p.y);
    return m31_point(m31_point_dbl_x(p.x), m31_add(xy, xy));
}
// fixed 32-step LSB-first double-and-add from the generator (2, 1268011823)   m31_point.simf:58-106
__device__ inline M31Point circle_point_index_to_m31_point(uint32_t index) {
    M31Point res = m31_point(1, 0), cur = m31_point(2, 1268011823u);
#pragma unroll 1
    for (int bit = 0; bit < 32; bit++) {
        M31Point sum = m31_point_add(res, cur);
        if ((index >> bit) & 1) res = sum;
        cur = m31_point_dbl(cur);
    }
    return res;
}

// ---- index algebra / domains                 groups/coset.simf, circle_domain.simf, line_domain.simf ----
__device__ __forceinline__ uint32_t shl32(uint32_t s, uint32_t x) { return (s & 0xff) >= 32 ? 0u : x << (s & 0xff); } // left_shift_32 jet, u8 amount
__device__ __forceinline__ uint32_t shr32(uint32_t s, uint32_t x) { return (s & 0xff) >= 32 ? 0u : x >> (s & 0xff); }
__device__ __forceinline__ uint32_t bit_reverse_position(uint32_t pos, uint32_t log_size) { return shr32((32u - log_size) & 0xff, __brev(pos)); } // coset.simf:20-25
__device__ __forceinline__ uint32_t circle_subgroup_gen(uint32_t log_size) { return shl32((31u - log_size) & 0xff, 1u); }                       // coset.simf:28-31
__device__ __forceinline__ uint32_t cpi_add(uint32_t l, uint32_t r) { return (l + r) & 0x7fffffffu; }                                           // coset.simf:34-37
__device__ __forceinline__ uint32_t cpi_mul(uint32_t l, uint32_t r) { return (l * r) & 0x7fffffffu; }                                           // coset.simf:40-45
__device__ __forceinline__ uint32_t cpi_neg(uint32_t i) { return (0x80000000u - i) & 0x7fffffffu; }                                             // coset.simf:48-51
// circle_position_to_point_index(circle_domain(log_size), position)                        circle_domain.simf:17-37
__device__ __forceinline__ uint32_t circle_position_to_point_index(uint32_t log_size, uint32_t position) {
    uint32_t half_size = shl32((log_size - 1u) & 0xff, 1u);
    uint32_t offset = circle_subgroup_gen((log_size + 1u) & 0xff);
    uint32_t step = circle_subgroup_gen((log_size - 1u) & 0xff);
    if (position < half_size) return cpi_add(offset, cpi_mul(step, position));
    return cpi_neg(cpi_add(offset, cpi_mul(step, position - half_size)));
}
// point index of line_position_to_x_coord(line_domain(log_size), position)                  line_domain.simf:18-31
__device__ __forceinline__ uint32_t line_position_to_point_index(uint32_t log_size, uint32_t position) {
    uint32_t offset = circle_subgroup_gen((log_size + 2u) & 0xff);
    uint32_t step = circle_subgroup_gen(log_size & 0xff);
    return cpi_add(offset, cpi_mul(step, position));
}

// ---- circle group over QM31 (only what verify_proof needs)          groups/qm31_point.simf ----
__device__ __forceinline__ QM31 qm31_point_dbl_x(QM31 x) { QM31 s = qm31_mul(x, x); return qm31_sub(qm31_add(s, s), qm31_one()); } // qm31_point.simf:27-31

} // namespace ssym
