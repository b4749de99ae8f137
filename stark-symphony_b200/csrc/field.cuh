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
// m31_mul for CANONICAL operands (a, b < p): the product is < 2^62, so one fold suffices.  Same value as m31_mul.
__device__ __forceinline__ M31 m31_mul_c(M31 a, M31 b) {
    const uint64_t x = (uint64_t)a * b;
    const uint32_t s = ((uint32_t)x & SSYM_P) + (uint32_t)(x >> 31); // < 2^32
    return s >= SSYM_P ? s - SSYM_P : s;
}
// The same product with one operand pre-doubled: (2a) * b = 2ab puts floor(ab / 2^31) in the high word and 2 * (ab mod 2^31) in the low
// word, so the Mersenne fold is one shift-and-add (LEA.HI) instead of and + funnel shift + add.  a, b canonical; same value as m31_mul.
__device__ __forceinline__ M31 m31_mul_d(uint32_t a_doubled, M31 b) {
    const uint64_t x = (uint64_t)a_doubled * b;
    const uint32_t s = (uint32_t)(x >> 32) + ((uint32_t)x >> 1); // < 2^32
    return s >= SSYM_P ? s - SSYM_P : s;
}
template <int N>
__device__ __forceinline__ M31 m31_sqn(M31 a) { // a canonical
#pragma unroll
    for (int i = 0; i < N; i++) a = m31_mul_d(a << 1, a);
    return a;
}
// a^(p-2) by the reference's addition chain; *fail |= (a == 0 bitwise)   fields/m31.simf:117-132
// (the chain maps 0 -> 0, which is the documented continuation value).  Every step of the chain is an m31_mul, whose result
// depends only on the residue of a: a is canonicalised once (p -> 0, as m31_mul(p, p) = 0) and the chain runs on the cheaper
// canonical-operand product.
__device__ __forceinline__ M31 m31_inv(M31 a, bool &fail) {
    fail = fail || (a == 0);
    a = m31_reduce(a);
    M31 t0 = m31_mul_d(a << 1, m31_sqn<2>(a));    // a^5
    const uint32_t t0d = t0 << 1;
    M31 t1 = m31_mul_d(t0d, m31_sqn<1>(t0));       // a^15
    M31 t2 = m31_mul_d(t0d, m31_sqn<3>(t1));       // a^125
    M31 t3 = m31_mul_d(t0d, m31_sqn<1>(t2));       // a^255
    const uint32_t t3d = t3 << 1;
    M31 t4 = m31_mul_d(t3d, m31_sqn<8>(t3));       // a^65535
    M31 t5 = m31_mul_d(t3d, m31_sqn<8>(t4));       // a^16777215
    return m31_mul_d(t2 << 1, m31_sqn<7>(t5));     // a^2147483645
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
__device__ __forceinline__ M31 m31_reduce64(uint64_t x);
// cm31.simf:79-86.  Both components are sums of m31_mul results, i.e. exact residues in canonical form for any u32 inputs: computed
// with canonicalised operands, 64-bit accumulation and one reduction per component.
__device__ __forceinline__ CM31 cm31_mul(CM31 x, CM31 y) {
    const uint32_t a0 = m31_reduce(x.a), a1 = m31_reduce(x.b), b0 = m31_reduce(y.a), b1 = m31_reduce(y.b);
    return cm31(m31_reduce64((uint64_t)a0 * b0 + (uint64_t)a1 * (SSYM_P - b1)), m31_reduce64((uint64_t)a0 * b1 + (uint64_t)a1 * b0));
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
// x mod p for any 64-bit x: 2^31 = 2^62 = 1 (mod p), so x = (x & p) + ((x >> 31) & p) + (x >> 62)
__device__ __forceinline__ M31 m31_reduce64(uint64_t x) {
    const uint32_t lo = (uint32_t)x & SSYM_P, mid = (uint32_t)(x >> 31) & SSYM_P, hi = (uint32_t)(x >> 62);
    uint32_t s = lo + mid;                 // <= 2^32 - 2
    s = (s & SSYM_P) + (s >> 31) + hi;     // <= p + 4
    return s >= SSYM_P ? s - SSYM_P : s;
}
// qm31.simf:73-80:  re = x.r*y.r + (x.i*y.i)*(2+i),  im = x.r*y.i + x.i*y.r,  every product a cm31_mul (cm31.simf:79-86).
// Every term of the reference formula passes through m31_mul, whose result depends only on the residues of its operands and is
// canonical, and sums of canonical values stay canonical: the output is the exact product mod p in canonical form for ANY u32
// inputs.  It is computed here with the operands canonicalised once, the 16 products accumulated in 64 bits (each < 2^62, at most
// four per sum) and 6 reductions instead of 20 products + 14 modular adds.
__device__ __forceinline__ QM31 qm31_mul(QM31 x, QM31 y) {
    const uint32_t a0 = m31_reduce(x.r.a), a1 = m31_reduce(x.r.b), a2 = m31_reduce(x.i.a), a3 = m31_reduce(x.i.b);
    const uint32_t b0 = m31_reduce(y.r.a), b1 = m31_reduce(y.r.b), b2 = m31_reduce(y.i.a), b3 = m31_reduce(y.i.b);
    const uint32_t n1 = SSYM_P - b1, n3 = SSYM_P - b3; // -b1, -b3 (in [1, p]: products stay < 2^62)
    // Y = x.i * y.i
    const uint32_t ya = m31_reduce64((uint64_t)a2 * b2 + (uint64_t)a3 * n3);
    const uint32_t yb = m31_reduce64((uint64_t)a2 * b3 + (uint64_t)a3 * b2);
    // re = x.r * y.r + Y * (2 + i) = (Xa + 2 Ya - Yb, Xb + 2 Yb + Ya)
    const uint32_t ra = m31_reduce64((uint64_t)a0 * b0 + (uint64_t)a1 * n1 + ((uint64_t)ya << 1) + (SSYM_P - yb));
    const uint32_t rb = m31_reduce64((uint64_t)a0 * b1 + (uint64_t)a1 * b0 + ((uint64_t)yb << 1) + ya);
    // im = x.r * y.i + x.i * y.r
    const uint32_t ia = m31_reduce64((uint64_t)a0 * b2 + (uint64_t)a1 * n3 + (uint64_t)a2 * b0 + (uint64_t)a3 * n1);
    const uint32_t ib = m31_reduce64((uint64_t)a0 * b3 + (uint64_t)a1 * b2 + (uint64_t)a2 * b1 + (uint64_t)a3 * b0);
    return qm31(ra, rb, ia, ib);
}
// circle_fold / line_fold without the twiddle lookup (fri/folding.simf:21-26, 35-40):
//     f0 = e0 + e1,  f1 = (e0 - e1) * inv,  result = f0 + alpha * f1          (qm31_add, qm31_sub, qm31_mul_m31, qm31_mul, qm31_add)
// for ANY u32 inputs.  Every step of the reference formula is exact modulo p on the residues of its operands except that the
// 32-bit adds wrap first ((a + b) mod 2^32, a + (p - b) mod 2^32: m31.simf:22-37) — so the wrapped sums are formed literally and
// everything after them is residue arithmetic with a canonical result: alpha * ((e0 - e1) * inv) = (alpha * inv) * (e0 - e1), the 16
// products and the wrapped f0 accumulated in 64 bits, 6 reductions in all (the step-by-step form has 20 products, 4 + 8 + 4 + 6 + 4).
__device__ __forceinline__ QM31 qm31_fold(QM31 e0, QM31 e1, M31 inv, QM31 alpha) {
    const uint32_t s0 = e0.r.a + e1.r.a, s1 = e0.r.b + e1.r.b, s2 = e0.i.a + e1.i.a, s3 = e0.i.b + e1.i.b; // f0 before its reduction
    const uint32_t b0 = m31_reduce(e0.r.a + (SSYM_P - e1.r.a)), b1 = m31_reduce(e0.r.b + (SSYM_P - e1.r.b));
    const uint32_t b2 = m31_reduce(e0.i.a + (SSYM_P - e1.i.a)), b3 = m31_reduce(e0.i.b + (SSYM_P - e1.i.b));
    const uint32_t ic = m31_reduce(inv);
    const uint32_t a0 = m31_mul_c(m31_reduce(alpha.r.a), ic), a1 = m31_mul_c(m31_reduce(alpha.r.b), ic);
    const uint32_t a2 = m31_mul_c(m31_reduce(alpha.i.a), ic), a3 = m31_mul_c(m31_reduce(alpha.i.b), ic);
    const uint32_t n1 = SSYM_P - b1, n3 = SSYM_P - b3;
    const uint32_t ya = m31_reduce64((uint64_t)a2 * b2 + (uint64_t)a3 * n3);
    const uint32_t yb = m31_reduce64((uint64_t)a2 * b3 + (uint64_t)a3 * b2);
    const uint32_t ra = m31_reduce64((uint64_t)a0 * b0 + (uint64_t)a1 * n1 + ((uint64_t)ya << 1) + (SSYM_P - yb) + s0);
    const uint32_t rb = m31_reduce64((uint64_t)a0 * b1 + (uint64_t)a1 * b0 + ((uint64_t)yb << 1) + ya + s1);
    const uint32_t ia = m31_reduce64((uint64_t)a0 * b2 + (uint64_t)a1 * n3 + (uint64_t)a2 * b0 + (uint64_t)a3 * n1 + s2); // < 2^64 - 2^34 + 2^32
    const uint32_t ib = m31_reduce64((uint64_t)a0 * b3 + (uint64_t)a1 * b2 + (uint64_t)a2 * b1 + (uint64_t)a3 * b0 + s3);
    return qm31(ra, rb, ia, ib);
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
