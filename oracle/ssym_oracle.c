/* SPDX-License-Identifier: MIT
 *
 * ssym_oracle.c — CPU restatement of the stark-symphony verifier programs.
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load it, and only as the
 * checker (or the timed CPU baseline) — never on the shipped GPU path.
 *
 * It is a plain-C, function-by-function restatement of the reference's SimplicityHL
 * sources; every function cites the `.simf` file:line it follows (paths relative to
 * the reference root).  The jets' C implementation (libsimplicity inside simplicity-sys
 * 0.4.0 @ m-kus/rust-simplicity 7c43d07c12a47906fb1de6af35f8d0f2c2f93127, a git
 * dependency that is NOT vendored in the reference tree, Cargo.toml:6-7) is restated
 * from the Simplicity core specification: add/subtract = wrap-around + discarded
 * carry, multiply_32 = exact u64 product, divide/modulo = truncated unsigned with
 * x/0 = 0 and x%0 = x, shifts = logical with amounts >= width giving 0,
 * sha_256_ctx_8_* = streaming FIPS 180-4 SHA-256 fed big-endian integers.
 *
 * Parity pin: the 86 in-source known-answer tests of the reference (tests/ transcribes
 * every value-bearing one), the fixtures stwo-verifier/tests/data/proof{,_test}.json
 * and the regenerated stark101 proof (= stark101/src/verifier.simf:44-388).  The
 * reference binary itself (`simfony run`) cannot be built here (no Rust toolchain,
 * un-vendored crates), so REF_LITERAL behaviour of `fri_answer` and of the last-layer
 * asserts is pinned by source text only — see DESIGN.md.
 *
 * Continuation semantics (needed because the batch verifier never early-exits): a
 * failed assert only records a status bit; `m31_inv(0)` continues with 0 (which is
 * also what the addition chain of m31.simf:124-130 evaluates to), an exhausted draw
 * loop continues with the last attempt's words, stark101 `div_mod` continues with
 * the current `t`.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "../include/ssym.h"

#define EXPORT __attribute__((visibility("default")))

/* ------------------------------------------------------------------------- */
/* assertion bookkeeping                                                      */
/* ------------------------------------------------------------------------- */
static __thread int t_fail; /* set by any failed assert since it was last cleared */
#define ORACLE_ASSERT(c) do { if (!(c)) t_fail = 1; } while (0)

/* ------------------------------------------------------------------------- */
/* jets (Simplicity core spec)                                                */
/* ------------------------------------------------------------------------- */
static inline uint32_t jet_add_32(uint32_t a, uint32_t b) { return a + b; }
static inline uint32_t jet_subtract_32(uint32_t a, uint32_t b) { return a - b; }
static inline uint8_t jet_add_8(uint8_t a, uint8_t b) { return (uint8_t)(a + b); }
static inline uint8_t jet_subtract_8(uint8_t a, uint8_t b) { return (uint8_t)(a - b); }
static inline uint64_t jet_multiply_32(uint32_t a, uint32_t b) { return (uint64_t)a * b; }
static inline uint32_t jet_modulo_32(uint32_t a, uint32_t b) { return b ? a % b : a; }
static inline uint64_t jet_modulo_64(uint64_t a, uint64_t b) { return b ? a % b : a; }
static inline uint32_t jet_divide_32(uint32_t a, uint32_t b) { return b ? a / b : 0; }
static inline int jet_divides_32(uint32_t a, uint32_t b) { return a ? (b % a == 0) : (b == 0); } /* a | b */
static inline uint32_t jet_left_shift_32(uint8_t s, uint32_t x) { return s >= 32 ? 0 : x << s; }
static inline uint32_t jet_right_shift_32(uint8_t s, uint32_t x) { return s >= 32 ? 0 : x >> s; }
static inline uint64_t jet_left_shift_64(uint8_t s, uint64_t x) { return s >= 64 ? 0 : x << s; }

/* ------------------------------------------------------------------------- */
/* SHA-256 (FIPS 180-4) as the sha_256_ctx_8_* jets use it                     */
/* ------------------------------------------------------------------------- */
typedef struct { uint32_t w[8]; } u256; /* w[0] = most significant word */

typedef struct {
    uint32_t h[8];
    uint8_t buf[64];
    uint64_t len; /* bytes absorbed */
} Ctx8;

static const uint32_t SHA_K[64] = {
    0x428a2f98, 0x71374491, 0xb5c0fbcf, 0xe9b5dba5, 0x3956c25b, 0x59f111f1, 0x923f82a4, 0xab1c5ed5,
    0xd807aa98, 0x12835b01, 0x243185be, 0x550c7dc3, 0x72be5d74, 0x80deb1fe, 0x9bdc06a7, 0xc19bf174,
    0xe49b69c1, 0xefbe4786, 0x0fc19dc6, 0x240ca1cc, 0x2de92c6f, 0x4a7484aa, 0x5cb0a9dc, 0x76f988da,
    0x983e5152, 0xa831c66d, 0xb00327c8, 0xbf597fc7, 0xc6e00bf3, 0xd5a79147, 0x06ca6351, 0x14292967,
    0x27b70a85, 0x2e1b2138, 0x4d2c6dfc, 0x53380d13, 0x650a7354, 0x766a0abb, 0x81c2c92e, 0x92722c85,
    0xa2bfe8a1, 0xa81a664b, 0xc24b8b70, 0xc76c51a3, 0xd192e819, 0xd6990624, 0xf40e3585, 0x106aa070,
    0x19a4c116, 0x1e376c08, 0x2748774c, 0x34b0bcb5, 0x391c0cb3, 0x4ed8aa4a, 0x5b9cca4f, 0x682e6ff3,
    0x748f82ee, 0x78a5636f, 0x84c87814, 0x8cc70208, 0x90befffa, 0xa4506ceb, 0xbef9a3f7, 0xc67178f2};

/* instrumentation (per thread, so that the multi-threaded CPU baseline does not bounce a shared cache line): compression-function calls and
 * M31 operations of the calling thread — the per-proof work figures of SURVEY.md section 8d come from these */
static _Thread_local uint64_t g_compressions, g_m31_mul, g_m31_add, g_m31_inv;
/* cost-model counters (tests/test_cost_model.py checks the product's closed-form model, csrc/cost.cpp, against these): calls of the
 * sha_256_ctx_8_* jets, message bytes, unreduced negations (= subtract_32), eq_256, index -> point conversions, repeated felt draws */
static _Thread_local uint64_t g_sha_init, g_sha_add_4, g_sha_add_8, g_sha_add_32, g_sha_finalize, g_sha_bytes, g_m31_neg, g_eq_256, g_point_from_index,
    g_draw_retries;

static inline uint32_t rotr(uint32_t x, int n) { return (x >> n) | (x << (32 - n)); }

/* CPU-baseline speed switch (bench.py `cpu_baseline_fast`): 0 = the literal port (byte-at-a-time absorb, scalar compression — the default, and what
 * every test pins); 1 = the same functions with 4-byte absorbs and, where the host CPU has them, the SHA-NI compression instructions.  Same digests
 * (tests/test_oracle_fixtures.py::test_fast_sha_mode_is_bit_identical); only the time differs. */
static int g_fast_sha;
#if defined(__x86_64__)
#include <cpuid.h>
#include <immintrin.h>
static int cpu_has_sha_ni(void) {
    unsigned a, b, c, d;
    if (!__get_cpuid_count(7, 0, &a, &b, &c, &d)) return 0;
    return (b >> 29) & 1;
}
__attribute__((target("sha,sse4.1,ssse3"))) static void sha_compress_shani(uint32_t h[8], const uint8_t blk[64]) {
    const __m128i bswap = _mm_set_epi64x(0x0c0d0e0f08090a0bULL, 0x0405060700010203ULL);
    __m128i t = _mm_loadu_si128((const __m128i *)&h[0]), s1 = _mm_loadu_si128((const __m128i *)&h[4]);
    t = _mm_shuffle_epi32(t, 0xB1);             /* CDAB */
    s1 = _mm_shuffle_epi32(s1, 0x1B);           /* EFGH */
    __m128i s0 = _mm_alignr_epi8(t, s1, 8);     /* ABEF */
    s1 = _mm_blend_epi16(s1, t, 0xF0);          /* CDGH */
    const __m128i save0 = s0, save1 = s1;
    __m128i m[4];
    for (int i = 0; i < 4; i++) m[i] = _mm_shuffle_epi8(_mm_loadu_si128((const __m128i *)(blk + 16 * i)), bswap);
    for (int r = 0; r < 16; r++) {              /* 4 rounds per iteration */
        __m128i w = m[r & 3];
        __m128i k = _mm_add_epi32(w, _mm_loadu_si128((const __m128i *)&SHA_K[4 * r]));
        s1 = _mm_sha256rnds2_epu32(s1, s0, k);
        s0 = _mm_sha256rnds2_epu32(s0, s1, _mm_shuffle_epi32(k, 0x0E));
        if (r < 12) {                           /* schedule W[4r+16 .. 4r+19] into the slot of W[4r .. 4r+3] */
            __m128i x = _mm_sha256msg1_epu32(m[r & 3], m[(r + 1) & 3]);
            x = _mm_add_epi32(x, _mm_alignr_epi8(m[(r + 3) & 3], m[(r + 2) & 3], 4));
            m[r & 3] = _mm_sha256msg2_epu32(x, m[(r + 3) & 3]);
        }
    }
    s0 = _mm_add_epi32(s0, save0);
    s1 = _mm_add_epi32(s1, save1);
    t = _mm_shuffle_epi32(s0, 0x1B);            /* FEBA */
    s1 = _mm_shuffle_epi32(s1, 0xB1);           /* DCHG */
    _mm_storeu_si128((__m128i *)&h[0], _mm_blend_epi16(t, s1, 0xF0)); /* DCBA */
    _mm_storeu_si128((__m128i *)&h[4], _mm_alignr_epi8(s1, t, 8));    /* HGFE */
}
#else
static int cpu_has_sha_ni(void) { return 0; }
static void sha_compress_shani(uint32_t h[8], const uint8_t blk[64]) { (void)h; (void)blk; }
#endif
static int g_sha_ni = -1;

static void sha_compress(uint32_t h[8], const uint8_t blk[64]) {
    if (g_fast_sha && g_sha_ni > 0) {
        sha_compress_shani(h, blk);
        g_compressions++;
        return;
    }
    uint32_t w[64];
    for (int i = 0; i < 16; i++)
        w[i] = ((uint32_t)blk[4 * i] << 24) | ((uint32_t)blk[4 * i + 1] << 16) | ((uint32_t)blk[4 * i + 2] << 8) | blk[4 * i + 3];
    for (int i = 16; i < 64; i++) {
        uint32_t s0 = rotr(w[i - 15], 7) ^ rotr(w[i - 15], 18) ^ (w[i - 15] >> 3);
        uint32_t s1 = rotr(w[i - 2], 17) ^ rotr(w[i - 2], 19) ^ (w[i - 2] >> 10);
        w[i] = w[i - 16] + s0 + w[i - 7] + s1;
    }
    uint32_t a = h[0], b = h[1], c = h[2], d = h[3], e = h[4], f = h[5], g = h[6], hh = h[7];
    for (int i = 0; i < 64; i++) {
        uint32_t S1 = rotr(e, 6) ^ rotr(e, 11) ^ rotr(e, 25);
        uint32_t ch = (e & f) ^ (~e & g);
        uint32_t t1 = hh + S1 + ch + SHA_K[i] + w[i];
        uint32_t S0 = rotr(a, 2) ^ rotr(a, 13) ^ rotr(a, 22);
        uint32_t mj = (a & b) ^ (a & c) ^ (b & c);
        uint32_t t2 = S0 + mj;
        hh = g; g = f; f = e; e = d + t1; d = c; c = b; b = a; a = t1 + t2;
    }
    h[0] += a; h[1] += b; h[2] += c; h[3] += d; h[4] += e; h[5] += f; h[6] += g; h[7] += hh;
    g_compressions++;
}

static Ctx8 sha_256_ctx_8_init(void) {
    Ctx8 c;
    static const uint32_t iv[8] = {0x6a09e667, 0xbb67ae85, 0x3c6ef372, 0xa54ff53a, 0x510e527f, 0x9b05688c, 0x1f83d9ab, 0x5be0cd19};
    memcpy(c.h, iv, sizeof iv);
    c.len = 0;
    g_sha_init++;
    return c;
}
static void ctx_add_byte(Ctx8 *c, uint8_t b) {
    c->buf[c->len % 64] = b;
    c->len++;
    if (c->len % 64 == 0) sha_compress(c->h, c->buf);
}
static Ctx8 sha_256_ctx_8_add_4(Ctx8 c, uint32_t v) {
    g_sha_add_4++;
    if (g_fast_sha && c.len % 4 == 0) { /* word-wise absorb: same bytes, one store */
        const uint32_t be = __builtin_bswap32(v);
        memcpy(c.buf + c.len % 64, &be, 4);
        c.len += 4;
        if (c.len % 64 == 0) sha_compress(c.h, c.buf);
        return c;
    }
    for (int i = 3; i >= 0; i--) ctx_add_byte(&c, (uint8_t)(v >> (8 * i)));
    return c;
}
static Ctx8 sha_256_ctx_8_add_8(Ctx8 c, uint64_t v) {
    g_sha_add_8++;
    for (int i = 7; i >= 0; i--) ctx_add_byte(&c, (uint8_t)(v >> (8 * i)));
    return c;
}
static Ctx8 sha_256_ctx_8_add_32(Ctx8 c, u256 v) {
    g_sha_add_32++;
    g_sha_add_4 -= 8; /* the literal path below feeds the eight words through sha_256_ctx_8_add_4: those are not jet calls of the program */
    if (g_fast_sha && c.len % 4 == 0 && c.len % 64 <= 32) { /* the 32 bytes fit the current block */
        g_sha_add_4 += 8;
        uint32_t be[8];
        for (int i = 0; i < 8; i++) be[i] = __builtin_bswap32(v.w[i]);
        memcpy(c.buf + c.len % 64, be, 32);
        c.len += 32;
        if (c.len % 64 == 0) sha_compress(c.h, c.buf);
        return c;
    }
    for (int i = 0; i < 8; i++) c = sha_256_ctx_8_add_4(c, v.w[i]);
    return c;
}
static u256 sha_256_ctx_8_finalize(Ctx8 c) {
    uint64_t bits = c.len * 8;
    g_sha_finalize++;
    g_sha_bytes += c.len;
    if (g_fast_sha) { /* FIPS 180-4 padding written block-wise */
        size_t at = c.len % 64;
        c.buf[at++] = 0x80;
        if (at > 56) {
            memset(c.buf + at, 0, 64 - at);
            sha_compress(c.h, c.buf);
            at = 0;
        }
        memset(c.buf + at, 0, 56 - at);
        const uint64_t be = __builtin_bswap64(bits);
        memcpy(c.buf + 56, &be, 8);
        sha_compress(c.h, c.buf);
        u256 r;
        memcpy(r.w, c.h, sizeof r.w);
        return r;
    }
    ctx_add_byte(&c, 0x80);
    while (c.len % 64 != 56) ctx_add_byte(&c, 0);
    for (int i = 7; i >= 0; i--) ctx_add_byte(&c, (uint8_t)(bits >> (8 * i)));
    u256 r;
    memcpy(r.w, c.h, sizeof r.w);
    return r;
}
static int eq_256(u256 a, u256 b) { g_eq_256++; return memcmp(a.w, b.w, sizeof a.w) == 0; }

/* ========================================================================= */
/* stwo-verifier                                                              */
/* ========================================================================= */

/* ---- fields/m31.simf ---------------------------------------------------- */
#define M31_MODULUS 2147483647u
typedef uint32_t M31;

/* exact x mod (2^31 - 1) for any 64-bit x by Mersenne folding: the fast-mode (oracle_set_fast_sha) replacement of the literal `%` */
static inline uint32_t m31_fold64(uint64_t x) {
    x = (x & M31_MODULUS) + (x >> 31); /* < 2^34 */
    x = (x & M31_MODULUS) + (x >> 31); /* <= 2^31 + 6 */
    x = (x & M31_MODULUS) + (x >> 31); /* <= 2^31 - 1 */
    return x == M31_MODULUS ? 0 : (uint32_t)x;
}
static M31 m31(uint32_t v) { return g_fast_sha ? m31_fold64(v) : jet_modulo_32(v, M31_MODULUS); }                 /* m31.simf:17-19 */
static M31 m31_add(M31 a, M31 b) { g_m31_add++; return m31(jet_add_32(a, b)); }     /* m31.simf:22-26 */
static M31 m31_neg(M31 a) { g_m31_neg++; return jet_subtract_32(M31_MODULUS, a); }               /* m31.simf:29-32 (unreduced) */
static M31 m31_sub(M31 a, M31 b) { return m31_add(a, m31_neg(b)); }                 /* m31.simf:35-37 */
static M31 m31_mul(M31 a, M31 b) {                                                  /* m31.simf:40-45 */
    g_m31_mul++;
    if (g_fast_sha) return m31_fold64(jet_multiply_32(a, b));
    return (uint32_t)jet_modulo_64(jet_multiply_32(a, b), M31_MODULUS);
}
static M31 m31_exp(M31 a, M31 b) { /* m31.simf:57-80: square-and-multiply, <= 65536 steps */
    uint32_t res = 1, base = a, e = b;
    for (uint32_t counter = 0; counter < 65536; counter++) {
        if (e == 0) return res;
        if (!jet_divides_32(2, e)) res = m31_mul(res, base);
        base = m31_mul(base, base);
        e = jet_divide_32(e, 2);
    }
    t_fail = 1; /* unwrap_left on Right */
    return res;
}
static M31 m31_pow2(M31 a) { return m31_mul(a, a); }        /* m31.simf:83-85 */
static M31 m31_pow4(M31 a) { return m31_pow2(m31_pow2(a)); } /* m31.simf:88-90 */
static M31 m31_pow8(M31 a) { return m31_pow2(m31_pow4(a)); } /* m31.simf:93-95 */
static M31 m31_pow16(M31 a) { return m31_pow4(m31_pow4(a)); } /* m31.simf:98-100 */
static int m31_eq(M31 a, M31 b) { return a == b; }          /* m31.simf:103-105 (bitwise) */
static M31 m31_inv(M31 a) {                                 /* m31.simf:117-132 */
    g_m31_inv++;
    if (a == 0) { /* is_zero_32 -> assert!(false); 0 */
        t_fail = 1;
        return 0;
    }
    M31 t0 = m31_mul(m31_pow4(a), a);
    M31 t1 = m31_mul(m31_pow2(t0), t0);
    M31 t2 = m31_mul(m31_pow8(t1), t0);
    M31 t3 = m31_mul(m31_pow2(t2), t0);
    M31 t4 = m31_mul(m31_pow16(m31_pow16(t3)), t3);
    M31 t5 = m31_mul(m31_pow16(m31_pow16(t4)), t3);
    return m31_mul(m31_pow16(m31_pow8(t5)), t2);
}
static M31 m31_div(M31 a, M31 b) { return m31_mul(a, m31_inv(b)); } /* m31.simf:135-137 */

/* ---- fields/cm31.simf --------------------------------------------------- */
typedef struct { M31 a, b; } CM31; /* a + bi */
static CM31 cm31_mk(M31 a, M31 b) { CM31 r = {a, b}; return r; }
static CM31 cm31_zero(void) { return cm31_mk(0, 0); }
static CM31 cm31_one(void) { return cm31_mk(1, 0); }
static CM31 cm31_add(CM31 x, CM31 y) { return cm31_mk(m31_add(x.a, y.a), m31_add(x.b, y.b)); } /* cm31.simf:30-34 */
static CM31 cm31_neg(CM31 x) { return cm31_mk(m31_neg(x.a), m31_neg(x.b)); }                   /* cm31.simf:37-40 */
static CM31 cm31_sub(CM31 x, CM31 y) { return cm31_mk(m31_sub(x.a, y.a), m31_sub(x.b, y.b)); } /* cm31.simf:43-47 */
static CM31 cm31_sub_m31(CM31 x, M31 y) { return cm31_mk(m31_sub(x.a, y), x.b); }              /* cm31.simf:50-53 */
static CM31 cm31_mul_m31(CM31 x, M31 y) { return cm31_mk(m31_mul(x.a, y), m31_mul(x.b, y)); }  /* cm31.simf:56-59 */
static CM31 cm31_div_m31(CM31 x, M31 y) { return cm31_mul_m31(x, m31_inv(y)); }                /* cm31.simf:62-65 */
static CM31 cm31_conj(CM31 x) { return cm31_mk(x.a, m31_neg(x.b)); }                           /* cm31.simf:73-76 */
static CM31 cm31_mul(CM31 x, CM31 y) {                                                         /* cm31.simf:79-86 */
    M31 re = m31_sub(m31_mul(x.a, y.a), m31_mul(x.b, y.b));
    M31 im = m31_add(m31_mul(x.a, y.b), m31_mul(x.b, y.a));
    return cm31_mk(re, im);
}
static CM31 cm31_inv(CM31 x) { /* cm31.simf:88-93 */
    CM31 cj = cm31_conj(x);
    M31 norm = m31_add(m31_pow2(x.a), m31_pow2(x.b));
    return cm31_div_m31(cj, norm);
}
static CM31 cm31_div(CM31 x, CM31 y) { return cm31_mul(x, cm31_inv(y)); } /* cm31.simf:96-99 */
static CM31 cm31_dbl(CM31 x) { return cm31_add(x, x); }                  /* cm31.simf:102-104 */
static int cm31_eq(CM31 x, CM31 y) { return x.a == y.a && x.b == y.b; }  /* cm31.simf:107-114 */

/* ---- fields/qm31.simf --------------------------------------------------- */
typedef struct { CM31 r, i; } QM31; /* (a + bi) + (c + di) j */
static QM31 qm31(M31 a, M31 b, M31 c, M31 d) { QM31 q = {{a, b}, {c, d}}; return q; }
static QM31 qm31_from_w(const uint32_t *w) { return qm31(w[0], w[1], w[2], w[3]); }
static void qm31_to_w(QM31 q, uint32_t *w) { w[0] = q.r.a; w[1] = q.r.b; w[2] = q.i.a; w[3] = q.i.b; }
static QM31 qm31_zero(void) { QM31 q = {cm31_zero(), cm31_zero()}; return q; }
static QM31 qm31_one(void) { QM31 q = {cm31_one(), cm31_zero()}; return q; }
static QM31 qm31_add(QM31 x, QM31 y) { QM31 q = {cm31_add(x.r, y.r), cm31_add(x.i, y.i)}; return q; } /* qm31.simf:36-40 */
static QM31 qm31_neg(QM31 x) { QM31 q = {cm31_neg(x.r), cm31_neg(x.i)}; return q; }                   /* qm31.simf:43-46 */
static QM31 qm31_sub(QM31 x, QM31 y) { QM31 q = {cm31_sub(x.r, y.r), cm31_sub(x.i, y.i)}; return q; } /* qm31.simf:49-53 */
static QM31 qm31_mul_m31(QM31 x, M31 y) { QM31 q = {cm31_mul_m31(x.r, y), cm31_mul_m31(x.i, y)}; return q; } /* qm31.simf:56-59 */
static QM31 qm31_mul_cm31(QM31 x, CM31 y) { QM31 q = {cm31_mul(x.r, y), cm31_mul(x.i, y)}; return q; }       /* qm31.simf:62-65 */
static QM31 qm31_conj(QM31 x) { QM31 q = {x.r, cm31_neg(x.i)}; return q; }                                   /* qm31.simf:68-71 */
static QM31 qm31_mul(QM31 x, QM31 y) { /* qm31.simf:73-80 */
    CM31 re = cm31_add(cm31_mul(x.r, y.r), cm31_mul(cm31_mul(x.i, y.i), cm31_mk(2, 1)));
    CM31 im = cm31_add(cm31_mul(x.r, y.i), cm31_mul(x.i, y.r));
    QM31 q = {re, im};
    return q;
}
static QM31 qm31_pow2(QM31 x) { return qm31_mul(x, x); } /* qm31.simf:83-85 */
static QM31 qm31_inv(QM31 x) {                           /* qm31.simf:87-98 */
    CM31 ar_sq = cm31_mul(x.r, x.r);
    CM31 ai_sq = cm31_mul(x.i, x.i);
    CM31 ai_sq_dbl = cm31_add(ai_sq, ai_sq);
    CM31 ai_sq_rev = cm31_mk(m31_neg(ai_sq.b), ai_sq.a);
    CM31 den = cm31_add(ar_sq, cm31_neg(cm31_add(ai_sq_dbl, ai_sq_rev)));
    CM31 den_inv = cm31_inv(den);
    QM31 q = {cm31_mul(x.r, den_inv), cm31_mul(cm31_neg(x.i), den_inv)};
    return q;
}
static QM31 qm31_div(QM31 x, QM31 y) { return qm31_mul(x, qm31_inv(y)); } /* qm31.simf:101-104 */
static QM31 qm31_from_m31(M31 a) { return qm31(a, 0, 0, 0); }            /* qm31.simf:112-114 */
static int qm31_eq(QM31 x, QM31 y) { return cm31_eq(x.r, y.r) && cm31_eq(x.i, y.i); } /* qm31.simf:117-124 */

/* ---- groups/m31_point.simf ---------------------------------------------- */
typedef struct { M31 x, y; } M31Point;
static M31Point m31_point_mk(M31 x, M31 y) { M31Point p = {x, y}; return p; }
static M31 m31_point_dbl_x(M31 x) { /* m31_point.simf:33-37 */
    M31 x_sq = m31_pow2(x);
    return m31_sub(m31_add(x_sq, x_sq), 1);
}
static M31Point m31_point_add(M31Point l, M31Point r) { /* m31_point.simf:40-46 */
    M31 r0 = m31_sub(m31_mul(l.x, r.x), m31_mul(l.y, r.y));
    M31 r1 = m31_add(m31_mul(l.x, r.y), m31_mul(l.y, r.x));
    return m31_point_mk(r0, r1);
}
static M31Point m31_point_dbl(M31Point p) { /* m31_point.simf:49-55 */
    M31 xy = m31_mul(p.x, p.y);
    return m31_point_mk(m31_point_dbl_x(p.x), m31_add(xy, xy));
}
static M31Point circle_point_index_to_m31_point(uint32_t index) { /* m31_point.simf:58-106: 32 LSB-first steps */
    M31Point res = m31_point_mk(1, 0), cur = m31_point_mk(2, 1268011823u);
    g_point_from_index++;
    for (int bit = 0; bit < 32; bit++) {
        if ((index >> bit) & 1) res = m31_point_add(res, cur);
        cur = m31_point_dbl(cur);
    }
    return res;
}
static M31Point m31_point_neg(M31Point p) { return m31_point_mk(p.x, m31_neg(p.y)); } /* m31_point.simf:109-112 */

/* ---- groups/qm31_point.simf --------------------------------------------- */
typedef struct { QM31 x, y; } QM31Point;
static QM31 qm31_point_dbl_x(QM31 x) { /* qm31_point.simf:27-31 */
    QM31 x_sq = qm31_mul(x, x);
    return qm31_sub(qm31_add(x_sq, x_sq), qm31_one());
}
static QM31Point qm31_point_add(QM31Point l, QM31Point r) { /* qm31_point.simf:34-40 */
    QM31Point p;
    p.x = qm31_sub(qm31_mul(l.x, r.x), qm31_mul(l.y, r.y));
    p.y = qm31_add(qm31_mul(l.x, r.y), qm31_mul(l.y, r.x));
    return p;
}
static QM31Point qm31_point_neg(QM31Point p) { p.y = qm31_neg(p.y); return p; } /* qm31_point.simf:43-46 */
static QM31Point qm31_point_add_m31_point(QM31Point l, M31Point r) {          /* qm31_point.simf:66-72 */
    QM31Point p;
    p.x = qm31_sub(qm31_mul_m31(l.x, r.x), qm31_mul_m31(l.y, r.y));
    p.y = qm31_add(qm31_mul_m31(l.x, r.y), qm31_mul_m31(l.y, r.x));
    return p;
}

/* ---- groups/coset.simf, circle_domain.simf, line_domain.simf ------------- */
#define M31_CIRCLE_LOG_ORDER 31
#define M31_CIRCLE_ORDER 0x80000000u
#define M31_CIRCLE_ORDER_BIT_MASK 0x7fffffffu
static uint8_t bit_reverse_u8(uint8_t v) { /* coset.simf:14-17 */
    uint8_t r = 0;
    for (int i = 0; i < 8; i++) r |= (uint8_t)(((v >> i) & 1) << (7 - i));
    return r;
}
static uint32_t bit_reverse_position(uint32_t pos, uint8_t log_size) { /* coset.simf:20-25 */
    uint8_t l0 = (uint8_t)(pos >> 24), l1 = (uint8_t)(pos >> 16), l2 = (uint8_t)(pos >> 8), l3 = (uint8_t)pos;
    uint32_t res = ((uint32_t)bit_reverse_u8(l3) << 24) | ((uint32_t)bit_reverse_u8(l2) << 16) |
                   ((uint32_t)bit_reverse_u8(l1) << 8) | bit_reverse_u8(l0);
    return jet_right_shift_32(jet_subtract_8(32, log_size), res);
}
static uint32_t circle_subgroup_gen(uint8_t log_size) { /* coset.simf:28-31 */
    return jet_left_shift_32(jet_subtract_8(M31_CIRCLE_LOG_ORDER, log_size), 1);
}
static uint32_t circle_point_index_add(uint32_t l, uint32_t r) { return jet_add_32(l, r) & M31_CIRCLE_ORDER_BIT_MASK; } /* coset.simf:34-37 */
static uint32_t circle_point_index_mul(uint32_t l, uint32_t r) { return (uint32_t)jet_multiply_32(l, r) & M31_CIRCLE_ORDER_BIT_MASK; } /* coset.simf:40-45 */
static uint32_t circle_point_index_neg(uint32_t i) { return jet_subtract_32(M31_CIRCLE_ORDER, i) & M31_CIRCLE_ORDER_BIT_MASK; } /* coset.simf:48-51 */

typedef struct { uint32_t half_size, offset, step; } CircleDomain;
static CircleDomain circle_domain(uint8_t log_size) { /* circle_domain.simf:17-24 */
    CircleDomain d;
    d.half_size = jet_left_shift_32(jet_subtract_8(log_size, 1), 1);
    d.offset = circle_subgroup_gen(jet_add_8(log_size, 1));
    d.step = circle_subgroup_gen(jet_subtract_8(log_size, 1));
    return d;
}
static uint32_t circle_position_to_point_index(CircleDomain d, uint32_t position) { /* circle_domain.simf:27-37 */
    if (position < d.half_size) return circle_point_index_add(d.offset, circle_point_index_mul(d.step, position));
    position = jet_subtract_32(position, d.half_size);
    return circle_point_index_neg(circle_point_index_add(d.offset, circle_point_index_mul(d.step, position)));
}
static M31Point circle_position_to_m31_point(CircleDomain d, uint32_t position) { /* circle_domain.simf:40-43 */
    return circle_point_index_to_m31_point(circle_position_to_point_index(d, position));
}
typedef struct { uint32_t offset, step; } LineDomain;
static LineDomain line_domain(uint8_t log_size) { /* line_domain.simf:18-23 */
    LineDomain d;
    d.offset = circle_subgroup_gen(jet_add_8(log_size, 2));
    d.step = circle_subgroup_gen(log_size);
    return d;
}
static M31 line_position_to_x_coord(LineDomain d, uint32_t position) { /* line_domain.simf:26-31 */
    uint32_t index = circle_point_index_add(d.offset, circle_point_index_mul(d.step, position));
    return circle_point_index_to_m31_point(index).x;
}

/* ---- hasher.simf --------------------------------------------------------- */
static u256 sha256(u256 in) { return sha_256_ctx_8_finalize(sha_256_ctx_8_add_32(sha_256_ctx_8_init(), in)); }       /* hasher.simf:13-17 */
static u256 sha256_32(uint32_t in) { return sha_256_ctx_8_finalize(sha_256_ctx_8_add_4(sha_256_ctx_8_init(), in)); } /* hasher.simf:20-24 */
static u256 sha256_pair(u256 l, u256 r) { /* hasher.simf:27-32 */
    Ctx8 c = sha_256_ctx_8_init();
    c = sha_256_ctx_8_add_32(c, l);
    c = sha_256_ctx_8_add_32(c, r);
    return sha_256_ctx_8_finalize(c);
}
static Ctx8 hasher_add_qm31(QM31 v, Ctx8 c) { /* hasher.simf:57-64 */
    c = sha_256_ctx_8_add_4(c, v.r.a);
    c = sha_256_ctx_8_add_4(c, v.r.b);
    c = sha_256_ctx_8_add_4(c, v.i.a);
    c = sha_256_ctx_8_add_4(c, v.i.b);
    return c;
}
static u256 hash_node_m31_trace(const M31 *evals /* NUM_COLUMNS x 1 */, uint32_t n_columns) { /* hasher.simf:85-90 */
    Ctx8 c = sha_256_ctx_8_init();
    for (uint32_t i = 0; i < n_columns; i++) c = sha_256_ctx_8_add_4(c, evals[i]);
    return sha_256_ctx_8_finalize(c);
}
static u256 hash_node_m31_cp(const M31 *evals /* 16 */) { /* hasher.simf:93-97 */
    Ctx8 c = sha_256_ctx_8_init();
    for (int i = 0; i < SSYM_NUM_CP_PARTITIONS; i++) c = sha_256_ctx_8_add_4(c, evals[i]);
    return sha_256_ctx_8_finalize(c);
}
static u256 hash_node_qm31(QM31 v) { /* hasher.simf:100-104 */
    return sha_256_ctx_8_finalize(hasher_add_qm31(v, sha_256_ctx_8_init()));
}

/* ---- merkle.simf --------------------------------------------------------- */
/* merkle.simf:22-30 + 39-44; returns the recomputed root and final path, asserts recorded in t_fail */
static u256 merkle_fold(u256 leaf, uint32_t auth_path, const uint32_t *sib, uint32_t n_sib, uint32_t *path_out) {
    u256 cur = leaf;
    uint32_t path = auth_path;
    for (uint32_t i = 0; i < n_sib; i++) {
        u256 s;
        memcpy(s.w, sib + 8 * i, 32);
        cur = jet_divides_32(2, path) ? sha256_pair(cur, s) : sha256_pair(s, cur);
        path = jet_divide_32(path, 2);
    }
    *path_out = path;
    return cur;
}
static u256 merkle_verify_32(u256 leaf, uint32_t auth_path, const uint32_t *sib, uint32_t n_sib, u256 root) {
    uint32_t path;
    u256 computed = merkle_fold(leaf, auth_path, sib, n_sib, &path);
    ORACLE_ASSERT(path == 1);            /* merkle.simf:42 */
    ORACLE_ASSERT(eq_256(computed, root)); /* merkle.simf:43 */
    return computed;
}

/* ---- channel.simf -------------------------------------------------------- */
typedef struct { u256 digest; uint32_t n_sent; } ChannelState;
#define DBL_P 4294967294u
static ChannelState channel_init(void) { ChannelState s; memset(&s, 0, sizeof s); return s; } /* channel.simf:31-33 */
static u256 channel_draw_u256(ChannelState *s) { /* channel.simf:36-44 */
    Ctx8 c = sha_256_ctx_8_init();
    c = sha_256_ctx_8_add_32(c, s->digest);
    c = sha_256_ctx_8_add_4(c, s->n_sent);
    u256 res = sha_256_ctx_8_finalize(c);
    s->n_sent = jet_add_32(s->n_sent, 1);
    return res;
}
/* channel.simf:48-58 split_256: our u256 is already the 8 big-endian limbs. */
static int is_uniform_n(const uint32_t *w, int n) { /* channel.simf:68-100 */
    for (int i = 0; i < n; i++)
        if (!(w[i] < DBL_P)) return 0;
    return 1;
}
static void channel_draw_m31xn(ChannelState *s, int n, M31 *out) { /* channel.simf:103-137 (n = 4 or 8) */
    u256 v;
    for (int counter = 0; counter < 256; counter++) {
        v = channel_draw_u256(s);
        if (counter) g_draw_retries++;
        if (is_uniform_n(v.w, n)) {
            for (int i = 0; i < n; i++) out[i] = m31(v.w[i]);
            return;
        }
    }
    t_fail = 1; /* unwrap_left on Right after 256 attempts */
    for (int i = 0; i < n; i++) out[i] = m31(v.w[i]);
}
static QM31 channel_draw_qm31(ChannelState *s) { /* channel.simf:138-141 */
    M31 v[4];
    channel_draw_m31xn(s, 4, v);
    return qm31(v[0], v[1], v[2], v[3]);
}
static QM31Point channel_draw_qm31_point(ChannelState *s) { /* channel.simf:143-151 */
    QM31 t = channel_draw_qm31(s);
    QM31 t_sq = qm31_pow2(t);
    QM31 inv = qm31_inv(qm31_add(qm31_one(), t_sq));
    QM31Point p;
    p.x = qm31_mul(qm31_sub(qm31_one(), t_sq), inv);
    p.y = qm31_mul(qm31_add(t, t), inv);
    return p;
}
static void channel_mix_u256(ChannelState *s, u256 in) { /* channel.simf:154-162 */
    Ctx8 c = sha_256_ctx_8_init();
    c = sha_256_ctx_8_add_32(c, s->digest);
    c = sha_256_ctx_8_add_32(c, in);
    s->digest = sha_256_ctx_8_finalize(c);
    s->n_sent = 0;
}
static void channel_mix_u64(ChannelState *s, uint64_t in) { /* channel.simf:165-173 */
    Ctx8 c = sha_256_ctx_8_init();
    c = sha_256_ctx_8_add_32(c, s->digest);
    c = sha_256_ctx_8_add_8(c, in);
    s->digest = sha_256_ctx_8_finalize(c);
    s->n_sent = 0;
}

/* ---- pow.simf ------------------------------------------------------------ */
static uint32_t reverse_bytes_32(uint32_t v) { /* pow.simf:12-19 */
    return (v >> 24) | ((v >> 8) & 0xff00u) | ((v << 8) & 0xff0000u) | (v << 24);
}
static uint64_t check_proof_of_work(ChannelState *s, uint64_t nonce, uint64_t target) { /* pow.simf:22-35 */
    channel_mix_u64(s, nonce);
    uint32_t g = s->digest.w[6], h = s->digest.w[7];
    uint64_t value = ((uint64_t)reverse_bytes_32(h) << 32) | reverse_bytes_32(g);
    ORACLE_ASSERT(value < target);
    return value;
}

/* ---- evals/commit.simf ---------------------------------------------------- */
static QM31 evals_commit(ChannelState *s, const u256 commitments[3]) { /* evals/commit.simf:20-35 */
    channel_mix_u256(s, commitments[0]);
    channel_mix_u256(s, commitments[1]);
    QM31 random_coeff = channel_draw_qm31(s);
    channel_mix_u256(s, commitments[2]);
    return random_coeff;
}

/* ---- evals/composition_poly.simf ----------------------------------------- */
static QM31 composition_poly_eval_from_partitions(QM31 cp0, QM31 cp1, QM31 cp2, QM31 cp3) { /* composition_poly.simf:38-44 */
    QM31 res = qm31_add(cp0, qm31_mul(cp1, qm31(0, 1, 0, 0)));
    res = qm31_add(res, qm31_mul(cp2, qm31(0, 0, 1, 0)));
    res = qm31_add(res, qm31_mul(cp3, qm31(0, 0, 0, 1)));
    return res;
}
static QM31 composition_poly_eval_from_decomposed(const QM31 e[16], QM31Point p) { /* composition_poly.simf:47-59 */
    /* (a0,b0,c0,d0, a1,b1,c1,d1, ...) : index = 4*coord + poly */
    QM31 cpa = composition_poly_eval_from_partitions(e[0], e[4], e[8], e[12]);
    QM31 cpb = composition_poly_eval_from_partitions(e[1], e[5], e[9], e[13]);
    QM31 cpc = composition_poly_eval_from_partitions(e[2], e[6], e[10], e[14]);
    QM31 cpd = composition_poly_eval_from_partitions(e[3], e[7], e[11], e[15]);
    QM31 res = qm31_add(cpa, qm31_mul(cpb, p.y));
    res = qm31_add(res, qm31_mul(cpc, p.x));
    return qm31_add(res, qm31_mul(cpd, qm31_mul(p.x, p.y)));
}
static QM31 vanishing_poly_eval(uint8_t log_size, QM31Point p) { /* composition_poly.simf:66-71 + pi_fn :26-35 */
    uint8_t n_iter = jet_subtract_8(log_size, 1);
    QM31 acc = p.x;
    for (int counter = 0; counter < 256; counter++) {
        if ((uint8_t)counter == n_iter) return acc;
        acc = qm31_point_dbl_x(acc);
    }
    t_fail = 1;
    return acc;
}

/* ---- constraints/wide_fibonacci.simf -------------------------------------- */
static QM31 eval_composition_poly(uint8_t log_size, QM31Point p, const QM31 *oods_trace /* NUM_COLUMNS */, uint32_t n_columns, QM31 random_coeff) { /* wide_fibonacci.simf:24-62 */
    QM31 acc = qm31_zero(), a = qm31_zero(), b = qm31_zero();
    uint8_t skip_2 = 0;
    for (uint32_t col = 0; col < n_columns; col++) { /* eval_column, left-to-right fold */
        QM31 c = oods_trace[col];
        if (skip_2 == 2) {
            QM31 constraint = qm31_sub(c, qm31_add(qm31_pow2(b), qm31_pow2(a)));
            acc = qm31_add(qm31_mul(acc, random_coeff), constraint);
        } else {
            skip_2 = jet_add_8(skip_2, 1);
        }
        a = b;
        b = c;
    }
    return qm31_div(acc, vanishing_poly_eval(log_size, p));
}

/* ---- deep/oods.simf -------------------------------------------------------- */
static void channel_mix_oods_evals(ChannelState *s, const QM31 *oods_trace /* NUM_COLUMNS */, uint32_t n_columns, const QM31 oods_cp[16]) { /* deep/oods.simf:23-39 */
    Ctx8 c = sha_256_ctx_8_init();
    c = sha_256_ctx_8_add_32(c, s->digest);
    for (uint32_t i = 0; i < n_columns; i++) c = hasher_add_qm31(oods_trace[i], c);
    for (int i = 0; i < SSYM_NUM_CP_PARTITIONS; i++) c = hasher_add_qm31(oods_cp[i], c);
    s->digest = sha_256_ctx_8_finalize(c);
    s->n_sent = 0;
}

/* ---- deep/quotients.simf ---------------------------------------------------- */
static CM31 deep_quotient_denominator_inverse(QM31Point sp, M31Point qp) { /* quotients.simf:15-22 */
    CM31 prx = sp.x.r, pix = sp.x.i, pry = sp.y.r, piy = sp.y.i;
    CM31 dx = cm31_sub_m31(prx, qp.x);
    CM31 dy = cm31_sub_m31(pry, qp.y);
    CM31 d = cm31_sub(cm31_mul(dx, piy), cm31_mul(dy, pix));
    return cm31_inv(d);
}
typedef struct { QM31 a, b, c; } LineCoeffs;
static LineCoeffs deep_quotient_interpolant_coefficients(QM31Point sp, QM31 sv, QM31 alpha_i) { /* quotients.simf:25-35 */
    QM31 py = sp.y;
    QM31 a = {cm31_zero(), cm31_neg(cm31_dbl(sv.i))};
    QM31 b = {cm31_zero(), cm31_neg(cm31_dbl(py.i))};
    QM31 a_py = qm31_mul(a, py);
    QM31 b_val = qm31_mul(b, sv);
    QM31 c = qm31_sub(b_val, a_py);
    LineCoeffs r = {qm31_mul(alpha_i, a), qm31_mul(alpha_i, b), qm31_mul(alpha_i, c)};
    return r;
}
static QM31 deep_quotient_nominator(LineCoeffs k, M31Point qp, M31 qv) { /* quotients.simf:38-44 */
    QM31 b_val = qm31_mul_m31(k.b, qv);
    QM31 a_py = qm31_mul_m31(k.a, qp.y);
    return qm31_sub(b_val, qm31_add(a_py, k.c));
}

/* ---- fri/folding.simf -------------------------------------------------------- */
static QM31 circle_fold(uint32_t position, QM31 f_p, QM31 f_neg_p, uint8_t log_size_ex, QM31 alpha) { /* folding.simf:15-27 */
    CircleDomain d = circle_domain(log_size_ex);
    M31 y = circle_position_to_m31_point(d, bit_reverse_position(position, log_size_ex)).y;
    M31 y_inv = m31_inv(y);
    QM31 f0 = qm31_add(f_p, f_neg_p);
    QM31 f1 = qm31_mul_m31(qm31_sub(f_p, f_neg_p), y_inv);
    return qm31_add(f0, qm31_mul(alpha, f1));
}
static QM31 line_fold(uint32_t position, QM31 f_p, QM31 f_neg_p, uint8_t log_size_ex, QM31 alpha) { /* folding.simf:30-41 */
    LineDomain d = line_domain(log_size_ex);
    M31 x = line_position_to_x_coord(d, bit_reverse_position(position, log_size_ex));
    M31 x_inv = m31_inv(x);
    QM31 f0 = qm31_add(f_p, f_neg_p);
    QM31 f1 = qm31_mul_m31(qm31_sub(f_p, f_neg_p), x_inv);
    return qm31_add(f0, qm31_mul(alpha, f1));
}

/* ---- fri/answers.simf --------------------------------------------------------- */
/* REF_LITERAL: fri/answers.simf:97-129. */
static QM31 fri_answer_literal(uint32_t query, const M31 *trace_evals /* NUM_COLUMNS */, uint32_t n_columns, const M31 cp_evals[16], QM31 random_coeff,
                               QM31Point oods_point, const QM31 *oods_trace, const QM31 oods_cp[16], uint8_t log_size_ex) {
    CircleDomain domain = circle_domain(log_size_ex);
    uint32_t position = bit_reverse_position(query, log_size_ex);
    M31Point dp = circle_position_to_m31_point(domain, position);
    CM31 den_inv = deep_quotient_denominator_inverse(oods_point, dp);
    QM31 acc = qm31(0, 0, 0, 0), alpha_i = random_coeff;
    for (uint32_t c = 0; c < n_columns; c++) { /* trace_quotient_numerator_aggregate, offset 0 */
        LineCoeffs k = deep_quotient_interpolant_coefficients(oods_point, oods_trace[c], alpha_i);
        acc = qm31_add(acc, deep_quotient_nominator(k, dp, trace_evals[c]));
        alpha_i = qm31_mul(alpha_i, random_coeff);
    }
    for (int c = 0; c < SSYM_NUM_CP_PARTITIONS; c++) {
        LineCoeffs k = deep_quotient_interpolant_coefficients(oods_point, oods_cp[c], alpha_i);
        acc = qm31_add(acc, deep_quotient_nominator(k, dp, cp_evals[c]));
        alpha_i = qm31_mul(alpha_i, random_coeff);
    }
    return qm31_mul(qm31_mul_cm31(acc, den_inv), alpha_i); /* batch_coeff = alpha^(NUM_COLUMNS + 17) */
}
/* PROVER_CONSISTENT: SURVEY.md Appendix A item 1 (the behaviour of the prover that produced tests/data/proof.json). */
static QM31 fri_answer_prover(uint32_t query, const M31 *trace_evals /* NUM_COLUMNS */, uint32_t n_columns, const M31 cp_evals[16], QM31 random_coeff,
                              QM31Point oods_point, const QM31 *oods_trace, const QM31 oods_cp[16], uint8_t log_size_ex) {
    CircleDomain domain = circle_domain(log_size_ex);
    uint32_t position = bit_reverse_position(query, log_size_ex);
    M31Point dp = circle_position_to_m31_point(domain, position);
    QM31Point p2;
    p2.x = qm31_point_dbl_x(oods_point.x);
    QM31 xy = qm31_mul(oods_point.x, oods_point.y);
    p2.y = qm31_add(xy, xy);
    QM31 num_a = qm31(0, 0, 0, 0), num_b = qm31(0, 0, 0, 0), alpha_i = random_coeff;
    for (int c = 0; c < SSYM_NUM_CP_PARTITIONS; c++) {
        LineCoeffs k = deep_quotient_interpolant_coefficients(p2, oods_cp[c], alpha_i);
        num_a = qm31_add(num_a, deep_quotient_nominator(k, dp, cp_evals[c]));
        alpha_i = qm31_mul(alpha_i, random_coeff);
    }
    for (uint32_t c = 0; c < n_columns; c++) {
        LineCoeffs k = deep_quotient_interpolant_coefficients(oods_point, oods_trace[c], alpha_i);
        num_b = qm31_add(num_b, deep_quotient_nominator(k, dp, trace_evals[c]));
        alpha_i = qm31_mul(alpha_i, random_coeff);
    }
    CM31 den_a = deep_quotient_denominator_inverse(p2, dp);
    CM31 den_b = deep_quotient_denominator_inverse(oods_point, dp);
    return qm31_add(qm31_mul_cm31(num_a, den_a), qm31_mul_cm31(num_b, den_b));
}

/* ---- packed layout (include/ssym.h) -------------------------------------------- */
static uint32_t align8(uint32_t w) { return (w + 7u) & ~7u; } /* 32-byte sections */

EXPORT int oracle_stwo_layout(const ssym_stwo_config_t *cfg, ssym_stwo_layout_t *o) {
    uint32_t Q = cfg->n_queries, L = cfg->n_fri_layers, G = cfg->lde_log, C = SSYM_STWO_COLUMNS(cfg);
    if (Q < 1 || Q > SSYM_MAX_QUERIES || L + 1 > SSYM_MAX_FRI_LAYERS || G < L + 1 || G > 30) return -1;
    if (C != 4 && C != 8 && C != 16) return -1;
    memset(o, 0, sizeof *o);
    uint32_t w = 0, alg = 0;
    o->off_commit = w; w += 24;
    o->off_oods_trace = w; w += 4 * C;
    o->off_oods_cp = w; w += 64;
    o->off_fri_first_root = w; w += 8;
    o->off_fri_inner_root = w; w += 8 * L;
    o->off_last_coeff = w; w += 4;
    o->off_pow_nonce = w; w += 2;
    alg += w; w = align8(w);
    o->off_qvals = w; w += Q * (C + 16); alg += Q * (C + 16); w = align8(w);
    o->off_trace_sib = w; w += Q * G * 8; alg += Q * G * 8;
    o->off_cp_sib = w; w += Q * G * 8; alg += Q * G * 8;
    o->off_fri_wit = w; w += (L + 1) * Q * 4; alg += (L + 1) * Q * 4; w = align8(w);
    for (uint32_t l = 0; l <= L; l++) {
        o->off_fri_sib[l] = w;
        w += Q * (G - 1 - l) * 8;
        alg += Q * (G - 1 - l) * 8;
    }
    o->stride_words = align8(w);
    o->algorithmic_bytes = alg * 4;
    return 0;
}

/* ---- verifier.simf:32-58 on one packed proof ------------------------------------- */
static u256 load_u256(const uint32_t *w) { u256 r; memcpy(r.w, w, 32); return r; }
static void store_u256(uint32_t *w, u256 v) { memcpy(w, v.w, 32); }

static uint32_t first_fail_code(const ssym_stwo_config_t *cfg, const ssym_stwo_trace_t *t);

EXPORT void oracle_stwo_verify_one(const ssym_stwo_config_t *cfg, const uint32_t *pk, ssym_stwo_trace_t *tr) {
    ssym_stwo_layout_t lo;
    memset(tr, 0, sizeof *tr);
    if (oracle_stwo_layout(cfg, &lo) != 0) { tr->status = SSYM_ST_SHAPE; return; }
    const uint32_t QL = cfg->n_queries; /* query SLOTS of the record (layout) */
    uint32_t Q = QL;                    /* queries verified: U <= QL under SSYM_MODE_QUERY_DEDUP */
    const uint32_t L = cfg->n_fri_layers, G = cfg->lde_log, C = SSYM_STWO_COLUMNS(cfg), QV = C + SSYM_NUM_CP_PARTITIONS;
    uint32_t status = 0;
    const uint64_t retries0 = g_draw_retries;

    u256 commitments[3];
    for (int i = 0; i < 3; i++) commitments[i] = load_u256(pk + lo.off_commit + 8 * i);
    QM31 oods_trace[SSYM_MAX_COLUMNS], oods_cp[16];
    for (uint32_t i = 0; i < C; i++) oods_trace[i] = qm31_from_w(pk + lo.off_oods_trace + 4 * i);
    for (int i = 0; i < 16; i++) oods_cp[i] = qm31_from_w(pk + lo.off_oods_cp + 4 * i);

    /* verifier.simf:36 */
    ChannelState state = channel_init();
    /* verifier.simf:39 evals_commit */
    t_fail = 0;
    QM31 cp_alpha = evals_commit(&state, commitments);
    if (t_fail) status |= SSYM_ST_DRAW_EXHAUSTED;
    store_u256(tr->digest_commit, state.digest);
    qm31_to_w(cp_alpha, tr->cp_alpha);

    /* verifier.simf:42 oods (deep/oods.simf:44-64) */
    t_fail = 0;
    QM31 t = channel_draw_qm31(&state);
    if (t_fail) status |= SSYM_ST_DRAW_EXHAUSTED;
    t_fail = 0;
    QM31Point oods_point;
    { /* channel.simf:143-151 with the draw split out so the failure class is attributable */
        QM31 t_sq = qm31_pow2(t);
        QM31 inv = qm31_inv(qm31_add(qm31_one(), t_sq));
        oods_point.x = qm31_mul(qm31_sub(qm31_one(), t_sq), inv);
        oods_point.y = qm31_mul(qm31_add(t, t), inv);
    }
    channel_mix_oods_evals(&state, oods_trace, C, oods_cp);
    QM31 cp_eval = eval_composition_poly((uint8_t)cfg->trace_log, oods_point, oods_trace, C, cp_alpha);
    if (t_fail) status |= SSYM_ST_OODS_INV_ZERO;
    QM31 sampled = composition_poly_eval_from_decomposed(oods_cp, oods_point);
    if (!qm31_eq(cp_eval, sampled)) status |= SSYM_ST_OODS_CP_MISMATCH; /* deep/oods.simf:58 */
    t_fail = 0;
    QM31 deep_alpha = channel_draw_qm31(&state);
    if (t_fail) status |= SSYM_ST_DRAW_EXHAUSTED;
    qm31_to_w(oods_point.x, tr->oods_x);
    qm31_to_w(oods_point.y, tr->oods_y);
    qm31_to_w(cp_eval, tr->cp_eval);
    qm31_to_w(sampled, tr->cp_sampled);
    store_u256(tr->digest_oods, state.digest);
    qm31_to_w(deep_alpha, tr->deep_alpha);

    /* verifier.simf:45 fri_commit (fri/commit.simf:72-85) */
    u256 fri_root[SSYM_MAX_FRI_LAYERS];
    QM31 fri_alpha[SSYM_MAX_FRI_LAYERS];
    t_fail = 0;
    for (uint32_t l = 0; l <= L; l++) {
        fri_root[l] = load_u256(l == 0 ? pk + lo.off_fri_first_root : pk + lo.off_fri_inner_root + 8 * (l - 1));
        channel_mix_u256(&state, fri_root[l]); /* fri/commit.simf:38 */
        fri_alpha[l] = channel_draw_qm31(&state);
        qm31_to_w(fri_alpha[l], tr->fri_alpha[l]);
    }
    if (t_fail) status |= SSYM_ST_DRAW_EXHAUSTED;
    QM31 last_coeff = qm31_from_w(pk + lo.off_last_coeff);
    { /* channel_mix_line_poly fri/commit.simf:48-57 */
        Ctx8 c = sha_256_ctx_8_init();
        c = sha_256_ctx_8_add_32(c, state.digest);
        c = hasher_add_qm31(last_coeff, c);
        state.digest = sha_256_ctx_8_finalize(c);
        state.n_sent = 0;
    }
    store_u256(tr->digest_fri, state.digest);

    /* verifier.simf:48 check_proof_of_work */
    uint64_t nonce = ((uint64_t)pk[lo.off_pow_nonce] << 32) | pk[lo.off_pow_nonce + 1];
    t_fail = 0;
    uint64_t pow_value = check_proof_of_work(&state, nonce, cfg->pow_target);
    if (t_fail) status |= SSYM_ST_POW_FAIL;
    store_u256(tr->digest_pow, state.digest);
    tr->pow_value[0] = (uint32_t)(pow_value >> 32);
    tr->pow_value[1] = (uint32_t)pow_value;

    /* verifier.simf:51 evals_verify -> fri_generate_queries (fri/queries.simf:30-43) */
    uint32_t query_mask = jet_subtract_32(jet_left_shift_32((uint8_t)G, 1), 1);
    uint32_t queries[SSYM_MAX_QUERIES];
    for (uint32_t q = 0; q < Q; q += 8) {
        u256 w = channel_draw_u256(&state);
        for (uint32_t j = 0; j < 8 && q + j < Q; j++) queries[q + j] = w.w[j] & query_mask;
    }
    /* SSYM_MODE_QUERY_DEDUP (include/ssym.h): sort, drop duplicates; slot j < U belongs to the j-th smallest distinct query, slots >= U are ignored */
    const uint32_t Q_drawn = Q;
    if (cfg->mode & SSYM_MODE_QUERY_DEDUP) {
        for (uint32_t a = 1; a < Q_drawn; a++) /* insertion sort */
            for (uint32_t b = a; b > 0 && queries[b] < queries[b - 1]; b--) { uint32_t t_ = queries[b]; queries[b] = queries[b - 1]; queries[b - 1] = t_; }
        uint32_t u = 0;
        for (uint32_t a = 0; a < Q_drawn; a++)
            if (a == 0 || queries[a] != queries[u - 1]) queries[u++] = queries[a];
        for (uint32_t a = u; a < Q_drawn; a++) queries[a] = 0;
        Q = u;
    }
    tr->n_queries_used = Q;
    uint32_t domain_size = jet_left_shift_32((uint8_t)G, 1); /* evals/verify.simf:119 */
    for (uint32_t q = 0; q < Q; q++) {
        tr->queries[q] = queries[q];
        const uint32_t *qv = pk + lo.off_qvals + QV * q;
        /* verify_trace_evals evals/verify.simf:50-58 */
        t_fail = 0;
        u256 r = merkle_verify_32(hash_node_m31_trace(qv, C), jet_add_32(queries[q], domain_size),
                                  pk + lo.off_trace_sib + q * G * 8, G, commitments[1]);
        if (t_fail) { status |= SSYM_ST_TRACE_MERKLE; tr->mask_trace |= 1u << q; }
        store_u256(tr->trace_root[q], r);
        /* verify_cp_evals evals/verify.simf:60-68 */
        t_fail = 0;
        r = merkle_verify_32(hash_node_m31_cp(qv + C), jet_add_32(queries[q], domain_size),
                             pk + lo.off_cp_sib + q * G * 8, G, commitments[2]);
        if (t_fail) { status |= SSYM_ST_CP_MERKLE; tr->mask_cp |= 1u << q; }
        store_u256(tr->cp_root[q], r);
    }

    /* verifier.simf:54 fri_answers */
    QM31 evals[SSYM_MAX_QUERIES];
    uint32_t fq[SSYM_MAX_QUERIES];
    for (uint32_t q = 0; q < Q; q++) {
        const uint32_t *qv = pk + lo.off_qvals + QV * q;
        t_fail = 0;
        evals[q] = (SSYM_MODE_SEMANTICS(cfg->mode) == SSYM_MODE_REF_LITERAL)
                       ? fri_answer_literal(queries[q], qv, C, qv + C, deep_alpha, oods_point, oods_trace, oods_cp, (uint8_t)G)
                       : fri_answer_prover(queries[q], qv, C, qv + C, deep_alpha, oods_point, oods_trace, oods_cp, (uint8_t)G);
        if (t_fail) { status |= SSYM_ST_ANSWER_INV_ZERO; tr->mask_answer_inv |= 1u << q; }
        qm31_to_w(evals[q], tr->fri_answer[q]);
        fq[q] = queries[q];
    }

    /* verifier.simf:57 fri_verify (fri/verify.simf:114-129) */
    uint8_t log_size_ex = (uint8_t)G;
    for (uint32_t l = 0; l <= L; l++) {
        uint32_t n_sib = G - 1 - l;
        for (uint32_t q = 0; q < Q; q++) { /* fri_verify_query fri/layers.simf:51-69 */
            QM31 witness = qm31_from_w(pk + lo.off_fri_wit + (l * QL + q) * 4);
            uint32_t position;
            QM31 e0, e1;
            if (jet_divides_32(2, fq[q])) { position = fq[q]; e0 = evals[q]; e1 = witness; } /* adjacent_leaves layers.simf:29-37 */
            else { position = jet_subtract_32(fq[q], 1); e0 = witness; e1 = evals[q]; }
            /* verify_decommitment layers.simf:40-48 */
            uint32_t dsize = jet_left_shift_32(log_size_ex, 1);
            u256 node = sha256_pair(hash_node_qm31(e0), hash_node_qm31(e1));
            uint32_t auth = jet_divide_32(jet_add_32(position, dsize), 2);
            t_fail = 0;
            u256 r = merkle_verify_32(node, auth, pk + lo.off_fri_sib[l] + q * n_sib * 8, n_sib, fri_root[l]);
            if (t_fail) { status |= SSYM_ST_FRI_MERKLE(l); tr->mask_fri[l] |= 1u << q; }
            store_u256(tr->fri_root[l][q], r);
            t_fail = 0;
            QM31 folded = (l == 0) ? circle_fold(position, e0, e1, log_size_ex, fri_alpha[l])
                                   : line_fold(position, e0, e1, log_size_ex, fri_alpha[l]);
            if (t_fail) { status |= SSYM_ST_FOLD_INV_ZERO; tr->mask_fold_inv[l] |= 1u << q; }
            qm31_to_w(folded, tr->folded[l][q]);
            evals[q] = folded;
            fq[q] = jet_divide_32(position, 2);
        }
        log_size_ex = jet_subtract_8(log_size_ex, 1); /* fri/verify.simf:76 */
    }
    if (SSYM_MODE_SEMANTICS(cfg->mode) == SSYM_MODE_REF_LITERAL && log_size_ex != 0) status |= SSYM_ST_FINAL_LOG; /* fri/verify.simf:127 */
    for (uint32_t q = 0; q < Q; q++) { /* fri_verify_last_layer fri/layers.simf:73-78 */
        if (SSYM_MODE_SEMANTICS(cfg->mode) == SSYM_MODE_REF_LITERAL && fq[q] != 0) { status |= SSYM_ST_LAST_QUERY; tr->mask_last_query |= 1u << q; }
        if (!qm31_eq(evals[q], last_coeff)) { status |= SSYM_ST_LAST_EVAL; tr->mask_last_eval |= 1u << q; }
    }
    tr->status = status;
    tr->first_fail = first_fail_code(cfg, tr);
    tr->draw_retries = (uint32_t)(g_draw_retries - retries0);
}

/* First failing assert in the reference's program order (verifier.simf:32-58), derived from the masks. */
static uint32_t ff(uint32_t stage, uint32_t layer, uint32_t q) { return (stage << 16) | (layer << 8) | q; }
static uint32_t first_fail_code(const ssym_stwo_config_t *cfg, const ssym_stwo_trace_t *t) {
    uint32_t s = t->status;
    if (!s) return 0;
    if (s & SSYM_ST_SHAPE) return ff(31, 0, 0);
    for (uint32_t b = 0; b <= 3; b++)
        if (s & (1u << b)) return ff(b, 0, 0);
    for (uint32_t q = 0; q < cfg->n_queries; q++) { /* evals/verify.simf:71-78: trace then cp, per query */
        if (t->mask_trace & (1u << q)) return ff(4, 0, q);
        if (t->mask_cp & (1u << q)) return ff(5, 0, q);
    }
    for (uint32_t q = 0; q < cfg->n_queries; q++)
        if (t->mask_answer_inv & (1u << q)) return ff(6, 0, q);
    for (uint32_t l = 0; l <= cfg->n_fri_layers; l++)
        for (uint32_t q = 0; q < cfg->n_queries; q++) {
            if (t->mask_fri[l] & (1u << q)) return ff(7 + l, l, q);
            if (t->mask_fold_inv[l] & (1u << q)) return ff(16, l, q);
        }
    if (s & SSYM_ST_FINAL_LOG) return ff(17, 0, 0);
    for (uint32_t q = 0; q < cfg->n_queries; q++) {
        if (t->mask_last_query & (1u << q)) return ff(18, 0, q);
        if (t->mask_last_eval & (1u << q)) return ff(19, 0, q);
    }
    return ff(30, 0, 0);
}

/* Batch entry used by tests and by bench.py's CPU baseline (rows [begin, end) so callers can shard over cores). */
EXPORT void oracle_stwo_verify_batch(const ssym_stwo_config_t *cfg, const uint32_t *packed, size_t begin, size_t end,
                                     uint32_t *accept_bits, uint32_t *status, ssym_stwo_trace_t *trace) {
    ssym_stwo_layout_t lo;
    if (oracle_stwo_layout(cfg, &lo) != 0) return;
    ssym_stwo_trace_t tmp;
    for (size_t i = begin; i < end; i++) {
        ssym_stwo_trace_t *tr = trace ? &trace[i] : &tmp;
        oracle_stwo_verify_one(cfg, packed + i * (size_t)lo.stride_words, tr);
        if (status) status[i] = tr->status;
        if (accept_bits) {
            if (tr->status == 0) __atomic_fetch_or(&accept_bits[i / 32], 1u << (i % 32), __ATOMIC_RELAXED);
            else __atomic_fetch_and(&accept_bits[i / 32], ~(1u << (i % 32)), __ATOMIC_RELAXED);
        }
    }
}

/* 0 = literal port (default); 1 = word-wise absorb + SHA-NI where available.  Returns 2 if SHA-NI is in use, 1 if only the word-wise absorb, 0 if off.
 * Process-wide: set it before starting worker threads. */
EXPORT int oracle_set_fast_sha(int on) {
    if (g_sha_ni < 0) g_sha_ni = cpu_has_sha_ni();
    g_fast_sha = on != 0;
    return g_fast_sha ? (g_sha_ni > 0 ? 2 : 1) : 0;
}
EXPORT uint64_t oracle_compression_count(void) { return g_compressions; }
EXPORT void oracle_compression_reset(void) { g_compressions = 0; }
/* M31 operations of the calling thread since the last reset: out = {mul (those inside inversions included), add / sub, inversions} */
EXPORT void oracle_field_op_counts(uint64_t out[3]) { out[0] = g_m31_mul; out[1] = g_m31_add; out[2] = g_m31_inv; }
EXPORT void oracle_field_op_reset(void) { g_m31_mul = g_m31_add = g_m31_inv = 0; }
/* Test-vector search (tests/golden/make_retry_fixture.py): a felt draw is repeated when one of the first four words of the drawn digest is >= 2p
 * (channel.simf:115-141), which happens once in 2^29 draws — no fixture of the reference exercises it.  Given the channel state BEFORE the trace root is
 * mixed (evals/commit.simf:23), finds a trace root of the form {base[0..5], hi, lo} for which the cp_alpha draw (evals/commit.simf:29) has to be repeated
 * at least once.  Scans counters [start, start + count); returns 1 and the root, or 0. */
EXPORT int oracle_grind_draw_retry(const uint32_t digest_before[8], const uint32_t base[8], uint64_t start, uint64_t count, uint32_t root_out[8]) {
    const int fast0 = g_fast_sha;
    if (g_sha_ni < 0) g_sha_ni = cpu_has_sha_ni();
    g_fast_sha = 1;
    int found = 0;
    for (uint64_t k = start; k < start + count && !found; k++) {
        ChannelState st;
        memcpy(st.digest.w, digest_before, 32);
        st.n_sent = 0;
        u256 root;
        memcpy(root.w, base, 32);
        root.w[6] = (uint32_t)(k >> 32);
        root.w[7] = (uint32_t)k;
        channel_mix_u256(&st, root);
        u256 v = channel_draw_u256(&st);
        if (!is_uniform_n(v.w, 4)) {
            memcpy(root_out, root.w, 32);
            found = 1;
        }
    }
    g_fast_sha = fast0;
    return found;
}

/* Everything the cost model predicts, in the field order of ssym_cost_t (include/ssym.h), for the calling thread since the last reset. */
EXPORT void oracle_cost_counts(uint64_t out[SSYM_COST_FIELDS]) {
    const uint64_t v[SSYM_COST_FIELDS] = {g_compressions, g_sha_init, g_sha_add_4, g_sha_add_8, g_sha_add_32, g_sha_finalize, g_sha_bytes,
                                          g_m31_mul, g_m31_add, g_m31_neg, g_m31_inv, g_eq_256, g_point_from_index, g_draw_retries};
    memcpy(out, v, sizeof v);
}
EXPORT void oracle_cost_reset(void) {
    g_compressions = g_sha_init = g_sha_add_4 = g_sha_add_8 = g_sha_add_32 = g_sha_finalize = g_sha_bytes = 0;
    g_m31_mul = g_m31_add = g_m31_neg = g_m31_inv = g_eq_256 = g_point_from_index = g_draw_retries = 0;
}

/* ------------------------------------------------------------------------- */
/* Function-level exports for the known-answer tests (ctypes)                 */
/* ------------------------------------------------------------------------- */
EXPORT uint32_t oracle_m31(uint32_t v) { return m31(v); }
EXPORT uint32_t oracle_m31_add(uint32_t a, uint32_t b) { return m31_add(a, b); }
EXPORT uint32_t oracle_m31_neg(uint32_t a) { return m31_neg(a); }
EXPORT uint32_t oracle_m31_sub(uint32_t a, uint32_t b) { return m31_sub(a, b); }
EXPORT uint32_t oracle_m31_mul(uint32_t a, uint32_t b) { return m31_mul(a, b); }
EXPORT uint32_t oracle_m31_exp(uint32_t a, uint32_t b) { return m31_exp(a, b); }
EXPORT uint32_t oracle_m31_inv(uint32_t a, int *fail) { t_fail = 0; uint32_t r = m31_inv(a); if (fail) *fail = t_fail; return r; }
EXPORT void oracle_cm31_add(const uint32_t *a, const uint32_t *b, uint32_t *o) { CM31 r = cm31_add(cm31_mk(a[0], a[1]), cm31_mk(b[0], b[1])); o[0] = r.a; o[1] = r.b; }
EXPORT void oracle_cm31_sub(const uint32_t *a, const uint32_t *b, uint32_t *o) { CM31 r = cm31_sub(cm31_mk(a[0], a[1]), cm31_mk(b[0], b[1])); o[0] = r.a; o[1] = r.b; }
EXPORT void oracle_cm31_mul(const uint32_t *a, const uint32_t *b, uint32_t *o) { CM31 r = cm31_mul(cm31_mk(a[0], a[1]), cm31_mk(b[0], b[1])); o[0] = r.a; o[1] = r.b; }
EXPORT void oracle_cm31_div(const uint32_t *a, const uint32_t *b, uint32_t *o, int *fail) { t_fail = 0; CM31 r = cm31_div(cm31_mk(a[0], a[1]), cm31_mk(b[0], b[1])); o[0] = r.a; o[1] = r.b; if (fail) *fail = t_fail; }
EXPORT void oracle_cm31_inv(const uint32_t *a, uint32_t *o, int *fail) { t_fail = 0; CM31 r = cm31_inv(cm31_mk(a[0], a[1])); o[0] = r.a; o[1] = r.b; if (fail) *fail = t_fail; }
EXPORT void oracle_qm31_add(const uint32_t *a, const uint32_t *b, uint32_t *o) { qm31_to_w(qm31_add(qm31_from_w(a), qm31_from_w(b)), o); }
EXPORT void oracle_qm31_sub(const uint32_t *a, const uint32_t *b, uint32_t *o) { qm31_to_w(qm31_sub(qm31_from_w(a), qm31_from_w(b)), o); }
EXPORT void oracle_qm31_neg(const uint32_t *a, uint32_t *o) { qm31_to_w(qm31_neg(qm31_from_w(a)), o); }
EXPORT void oracle_qm31_conj(const uint32_t *a, uint32_t *o) { qm31_to_w(qm31_conj(qm31_from_w(a)), o); }
EXPORT void oracle_qm31_mul(const uint32_t *a, const uint32_t *b, uint32_t *o) { qm31_to_w(qm31_mul(qm31_from_w(a), qm31_from_w(b)), o); }
EXPORT void oracle_qm31_mul_m31(const uint32_t *a, uint32_t b, uint32_t *o) { qm31_to_w(qm31_mul_m31(qm31_from_w(a), b), o); }
EXPORT void oracle_qm31_mul_cm31(const uint32_t *a, const uint32_t *b, uint32_t *o) { qm31_to_w(qm31_mul_cm31(qm31_from_w(a), cm31_mk(b[0], b[1])), o); }
EXPORT void oracle_qm31_inv(const uint32_t *a, uint32_t *o, int *fail) { t_fail = 0; qm31_to_w(qm31_inv(qm31_from_w(a)), o); if (fail) *fail = t_fail; }
EXPORT void oracle_qm31_div(const uint32_t *a, const uint32_t *b, uint32_t *o, int *fail) { t_fail = 0; qm31_to_w(qm31_div(qm31_from_w(a), qm31_from_w(b)), o); if (fail) *fail = t_fail; }

EXPORT void oracle_m31_point_add(const uint32_t *a, const uint32_t *b, uint32_t *o) { M31Point r = m31_point_add(m31_point_mk(a[0], a[1]), m31_point_mk(b[0], b[1])); o[0] = r.x; o[1] = r.y; }
EXPORT void oracle_m31_point_dbl(const uint32_t *a, uint32_t *o) { M31Point r = m31_point_dbl(m31_point_mk(a[0], a[1])); o[0] = r.x; o[1] = r.y; }
EXPORT void oracle_m31_point_neg(const uint32_t *a, uint32_t *o) { M31Point r = m31_point_neg(m31_point_mk(a[0], a[1])); o[0] = r.x; o[1] = r.y; }
EXPORT void oracle_circle_point_index_to_m31_point(uint32_t idx, uint32_t *o) { M31Point r = circle_point_index_to_m31_point(idx); o[0] = r.x; o[1] = r.y; }
static QM31Point qp_from_w(const uint32_t *w) { QM31Point p = {qm31_from_w(w), qm31_from_w(w + 4)}; return p; }
static void qp_to_w(QM31Point p, uint32_t *w) { qm31_to_w(p.x, w); qm31_to_w(p.y, w + 4); }
EXPORT void oracle_qm31_point_add(const uint32_t *a, const uint32_t *b, uint32_t *o) { qp_to_w(qm31_point_add(qp_from_w(a), qp_from_w(b)), o); }
EXPORT void oracle_qm31_point_neg(const uint32_t *a, uint32_t *o) { qp_to_w(qm31_point_neg(qp_from_w(a)), o); }
EXPORT void oracle_qm31_point_add_m31_point(const uint32_t *a, const uint32_t *b, uint32_t *o) { qp_to_w(qm31_point_add_m31_point(qp_from_w(a), m31_point_mk(b[0], b[1])), o); }
EXPORT uint32_t oracle_bit_reverse_position(uint32_t pos, uint32_t log_size) { return bit_reverse_position(pos, (uint8_t)log_size); }
EXPORT uint32_t oracle_circle_point_index_add(uint32_t a, uint32_t b) { return circle_point_index_add(a, b); }
EXPORT uint32_t oracle_circle_point_index_mul(uint32_t a, uint32_t b) { return circle_point_index_mul(a, b); }
EXPORT uint32_t oracle_circle_point_index_neg(uint32_t a) { return circle_point_index_neg(a); }
EXPORT void oracle_circle_domain(uint32_t log_size, uint32_t *o) { CircleDomain d = circle_domain((uint8_t)log_size); o[0] = d.half_size; o[1] = d.offset; o[2] = d.step; }
EXPORT uint32_t oracle_circle_position_to_point_index(uint32_t log_size, uint32_t pos) { return circle_position_to_point_index(circle_domain((uint8_t)log_size), pos); }
EXPORT uint32_t oracle_line_position_to_x_coord(uint32_t log_size, uint32_t pos) { return line_position_to_x_coord(line_domain((uint8_t)log_size), pos); }

EXPORT void oracle_sha256(const uint32_t *in, uint32_t *o) { store_u256(o, sha256(load_u256(in))); }
EXPORT void oracle_sha256_32(uint32_t in, uint32_t *o) { store_u256(o, sha256_32(in)); }
EXPORT void oracle_sha256_pair(const uint32_t *l, const uint32_t *r, uint32_t *o) { store_u256(o, sha256_pair(load_u256(l), load_u256(r))); }
EXPORT void oracle_sha256_bytes(const uint8_t *data, size_t len, uint32_t *o) { Ctx8 c = sha_256_ctx_8_init(); for (size_t i = 0; i < len; i++) ctx_add_byte(&c, data[i]); store_u256(o, sha_256_ctx_8_finalize(c)); }
EXPORT void oracle_hash_node_m31_trace(const uint32_t *e, uint32_t *o) { store_u256(o, hash_node_m31_trace(e, SSYM_NUM_COLUMNS)); }
EXPORT void oracle_hash_node_m31_cp(const uint32_t *e, uint32_t *o) { store_u256(o, hash_node_m31_cp(e)); }
EXPORT void oracle_hash_node_qm31(const uint32_t *e, uint32_t *o) { store_u256(o, hash_node_qm31(qm31_from_w(e))); }
/* merkle.simf:39-44; returns 1 iff both asserts hold */
EXPORT int oracle_merkle_verify_32(const uint32_t *leaf, uint32_t auth_path, const uint32_t *sib, uint32_t n_sib, const uint32_t *root,
                                   uint32_t *computed_root, uint32_t *final_path) {
    uint32_t path;
    u256 c = merkle_fold(load_u256(leaf), auth_path, sib, n_sib, &path);
    if (computed_root) store_u256(computed_root, c);
    if (final_path) *final_path = path;
    return path == 1 && eq_256(c, load_u256(root));
}
static ChannelState st_from_w(const uint32_t *w) { ChannelState s; memcpy(s.digest.w, w, 32); s.n_sent = w[8]; return s; }
static void st_to_w(ChannelState s, uint32_t *w) { memcpy(w, s.digest.w, 32); w[8] = s.n_sent; }
EXPORT void oracle_channel_mix_u256(uint32_t *st, const uint32_t *in) { ChannelState s = st_from_w(st); channel_mix_u256(&s, load_u256(in)); st_to_w(s, st); }
EXPORT void oracle_channel_mix_u64(uint32_t *st, uint32_t hi, uint32_t lo) { ChannelState s = st_from_w(st); channel_mix_u64(&s, ((uint64_t)hi << 32) | lo); st_to_w(s, st); }
EXPORT void oracle_channel_draw_qm31(uint32_t *st, uint32_t *o, int *fail) { ChannelState s = st_from_w(st); t_fail = 0; qm31_to_w(channel_draw_qm31(&s), o); if (fail) *fail = t_fail; st_to_w(s, st); }
EXPORT void oracle_channel_draw_m31x8(uint32_t *st, uint32_t *o, int *fail) { ChannelState s = st_from_w(st); t_fail = 0; channel_draw_m31xn(&s, 8, o); if (fail) *fail = t_fail; st_to_w(s, st); }
EXPORT void oracle_channel_draw_qm31_point(uint32_t *st, uint32_t *o, int *fail) { ChannelState s = st_from_w(st); t_fail = 0; qp_to_w(channel_draw_qm31_point(&s), o); if (fail) *fail = t_fail; st_to_w(s, st); }
EXPORT void oracle_channel_draw_queries(uint32_t *st, uint32_t log_size, uint32_t n_queries, uint32_t *o) { /* fri/queries.simf:14-43 */
    ChannelState s = st_from_w(st);
    uint32_t mask = jet_subtract_32(jet_left_shift_32((uint8_t)log_size, 1), 1);
    for (uint32_t q = 0; q < n_queries; q += 8) {
        u256 w = channel_draw_u256(&s);
        for (uint32_t j = 0; j < 8 && q + j < n_queries; j++) o[q + j] = w.w[j] & mask;
    }
    st_to_w(s, st);
}
EXPORT uint32_t oracle_reverse_bytes_32(uint32_t v) { return reverse_bytes_32(v); }
EXPORT int oracle_check_proof_of_work(uint32_t *st, uint32_t hi, uint32_t lo, uint64_t target) { ChannelState s = st_from_w(st); t_fail = 0; check_proof_of_work(&s, ((uint64_t)hi << 32) | lo, target); st_to_w(s, st); return !t_fail; }
EXPORT void oracle_evals_commit(uint32_t *st, const uint32_t *commitments /* 3x8 */, uint32_t *coeff) {
    ChannelState s = st_from_w(st);
    u256 c[3];
    for (int i = 0; i < 3; i++) c[i] = load_u256(commitments + 8 * i);
    qm31_to_w(evals_commit(&s, c), coeff);
    st_to_w(s, st);
}
EXPORT void oracle_composition_poly_eval_from_partitions(const uint32_t *p /* 4x4 */, uint32_t *o) { qm31_to_w(composition_poly_eval_from_partitions(qm31_from_w(p), qm31_from_w(p + 4), qm31_from_w(p + 8), qm31_from_w(p + 12)), o); }
EXPORT void oracle_vanishing_poly_eval(uint32_t log_size, const uint32_t *point, uint32_t *o) { qm31_to_w(vanishing_poly_eval((uint8_t)log_size, qp_from_w(point)), o); }
EXPORT void oracle_eval_composition_poly(uint32_t log_size, const uint32_t *point, const uint32_t *oods_trace /* 4x4 */, const uint32_t *coeff, uint32_t *o, int *fail) {
    QM31 tr[4];
    for (int i = 0; i < 4; i++) tr[i] = qm31_from_w(oods_trace + 4 * i);
    t_fail = 0;
    qm31_to_w(eval_composition_poly((uint8_t)log_size, qp_from_w(point), tr, SSYM_NUM_COLUMNS, qm31_from_w(coeff)), o);
    if (fail) *fail = t_fail;
}
EXPORT void oracle_channel_mix_oods_evals(uint32_t *st, const uint32_t *oods_trace, const uint32_t *oods_cp) {
    ChannelState s = st_from_w(st);
    QM31 tr[4], cp[16];
    for (int i = 0; i < 4; i++) tr[i] = qm31_from_w(oods_trace + 4 * i);
    for (int i = 0; i < 16; i++) cp[i] = qm31_from_w(oods_cp + 4 * i);
    channel_mix_oods_evals(&s, tr, SSYM_NUM_COLUMNS, cp);
    st_to_w(s, st);
}
/* deep/oods.simf:44-64; returns 1 iff the CP assert (:58) holds and no inverse-of-zero occurred */
EXPORT int oracle_oods(uint32_t *st, uint32_t log_size, const uint32_t *oods_trace, const uint32_t *oods_cp, const uint32_t *cp_alpha,
                       uint32_t *deep_alpha, uint32_t *point) {
    ChannelState s = st_from_w(st);
    QM31 tr[4], cp[16];
    for (int i = 0; i < 4; i++) tr[i] = qm31_from_w(oods_trace + 4 * i);
    for (int i = 0; i < 16; i++) cp[i] = qm31_from_w(oods_cp + 4 * i);
    t_fail = 0;
    QM31Point p = channel_draw_qm31_point(&s);
    channel_mix_oods_evals(&s, tr, SSYM_NUM_COLUMNS, cp);
    QM31 cp_eval = eval_composition_poly((uint8_t)log_size, p, tr, SSYM_NUM_COLUMNS, qm31_from_w(cp_alpha));
    QM31 sampled = composition_poly_eval_from_decomposed(cp, p);
    ORACLE_ASSERT(qm31_eq(cp_eval, sampled));
    qm31_to_w(channel_draw_qm31(&s), deep_alpha);
    if (point) qp_to_w(p, point);
    st_to_w(s, st);
    return !t_fail;
}
/* fri/commit.simf:72-85 */
EXPORT void oracle_fri_commit(uint32_t *st, const uint32_t *first_root, const uint32_t *inner_roots, uint32_t n_inner, const uint32_t *last_coeff,
                              uint32_t *alphas /* (1+n_inner) x 4 */) {
    ChannelState s = st_from_w(st);
    for (uint32_t l = 0; l <= n_inner; l++) {
        channel_mix_u256(&s, load_u256(l == 0 ? first_root : inner_roots + 8 * (l - 1)));
        qm31_to_w(channel_draw_qm31(&s), alphas + 4 * l);
    }
    Ctx8 c = sha_256_ctx_8_init();
    c = sha_256_ctx_8_add_32(c, s.digest);
    c = hasher_add_qm31(qm31_from_w(last_coeff), c);
    s.digest = sha_256_ctx_8_finalize(c);
    s.n_sent = 0;
    st_to_w(s, st);
}
EXPORT void oracle_deep_quotient_denominator_inverse(const uint32_t *sample_point, const uint32_t *query_point, uint32_t *o, int *fail) {
    t_fail = 0;
    CM31 r = deep_quotient_denominator_inverse(qp_from_w(sample_point), m31_point_mk(query_point[0], query_point[1]));
    o[0] = r.a; o[1] = r.b;
    if (fail) *fail = t_fail;
}
EXPORT void oracle_deep_quotient_interpolant_coefficients(const uint32_t *sample_point, const uint32_t *sample_value, const uint32_t *alpha_i, uint32_t *o /* 3x4 */) {
    LineCoeffs k = deep_quotient_interpolant_coefficients(qp_from_w(sample_point), qm31_from_w(sample_value), qm31_from_w(alpha_i));
    qm31_to_w(k.a, o); qm31_to_w(k.b, o + 4); qm31_to_w(k.c, o + 8);
}
EXPORT void oracle_deep_quotient_nominator(const uint32_t *coeffs /* 3x4 */, const uint32_t *query_point, uint32_t query_value, uint32_t *o) {
    LineCoeffs k = {qm31_from_w(coeffs), qm31_from_w(coeffs + 4), qm31_from_w(coeffs + 8)};
    qm31_to_w(deep_quotient_nominator(k, m31_point_mk(query_point[0], query_point[1]), query_value), o);
}
EXPORT void oracle_fri_answer(uint32_t mode, uint32_t query, const uint32_t *trace_evals, const uint32_t *cp_evals, const uint32_t *coeff,
                              const uint32_t *point, const uint32_t *oods_trace, const uint32_t *oods_cp, uint32_t log_size_ex, uint32_t *o, int *fail) {
    QM31 tr[4], cp[16];
    for (int i = 0; i < 4; i++) tr[i] = qm31_from_w(oods_trace + 4 * i);
    for (int i = 0; i < 16; i++) cp[i] = qm31_from_w(oods_cp + 4 * i);
    t_fail = 0;
    QM31 r = mode == SSYM_MODE_REF_LITERAL
                 ? fri_answer_literal(query, trace_evals, SSYM_NUM_COLUMNS, cp_evals, qm31_from_w(coeff), qp_from_w(point), tr, cp, (uint8_t)log_size_ex)
                 : fri_answer_prover(query, trace_evals, SSYM_NUM_COLUMNS, cp_evals, qm31_from_w(coeff), qp_from_w(point), tr, cp, (uint8_t)log_size_ex);
    qm31_to_w(r, o);
    if (fail) *fail = t_fail;
}
EXPORT void oracle_circle_fold(uint32_t position, const uint32_t *f_p, const uint32_t *f_neg_p, uint32_t log_size, const uint32_t *alpha, uint32_t *o, int *fail) {
    t_fail = 0;
    qm31_to_w(circle_fold(position, qm31_from_w(f_p), qm31_from_w(f_neg_p), (uint8_t)log_size, qm31_from_w(alpha)), o);
    if (fail) *fail = t_fail;
}
EXPORT void oracle_line_fold(uint32_t position, const uint32_t *f_p, const uint32_t *f_neg_p, uint32_t log_size, const uint32_t *alpha, uint32_t *o, int *fail) {
    t_fail = 0;
    qm31_to_w(line_fold(position, qm31_from_w(f_p), qm31_from_w(f_neg_p), (uint8_t)log_size, qm31_from_w(alpha)), o);
    if (fail) *fail = t_fail;
}
/* verify_decommitment fri/layers.simf:40-48; returns 1 iff both merkle asserts hold */
EXPORT int oracle_verify_decommitment(uint32_t position, const uint32_t *eval0, const uint32_t *eval1, uint32_t log_size_ex, const uint32_t *sib, uint32_t n_sib, const uint32_t *root) {
    uint32_t dsize = jet_left_shift_32((uint8_t)log_size_ex, 1);
    u256 node = sha256_pair(hash_node_qm31(qm31_from_w(eval0)), hash_node_qm31(qm31_from_w(eval1)));
    uint32_t auth = jet_divide_32(jet_add_32(position, dsize), 2);
    t_fail = 0;
    merkle_verify_32(node, auth, sib, n_sib, load_u256(root));
    return !t_fail;
}

/* ========================================================================= */
/* stark101                                                                    */
/* ========================================================================= */
#define FIELD_MODULUS 3221225473u
#define FIELD_GEN 5u
#define IDX_OFFSET 8u
#define DOMAIN_EX_SIZE 8192u
#define CANONIC_COSET_GEN 1734477367u

static uint32_t add_mod(uint32_t a, uint32_t b) { return (uint32_t)jet_modulo_64((uint64_t)a + (uint64_t)b, FIELD_MODULUS); } /* field.simf:14-21 */
static uint32_t sub_mod(uint32_t a, uint32_t b) { return add_mod(a, jet_subtract_32(FIELD_MODULUS, b)); }                   /* field.simf:24-27 */
static uint32_t mul_mod(uint32_t a, uint32_t b) { return (uint32_t)jet_modulo_64(jet_multiply_32(a, b), FIELD_MODULUS); }    /* field.simf:30-35 */
static uint32_t div_mod(uint32_t a, uint32_t b) { /* field.simf:42-66: extended Euclid, <= 65536 steps */
    uint32_t t = 0, r = FIELD_MODULUS, new_t = 1, new_r = b;
    int done = 0;
    for (uint32_t counter = 0; counter < 65536; counter++) {
        if (new_r == 0) {
            ORACLE_ASSERT(r == 1); /* field.simf:46 */
            done = 1;
            break;
        }
        uint32_t q = jet_divide_32(r, new_r);
        uint32_t nt = sub_mod(t, mul_mod(q, new_t));
        uint32_t nr = sub_mod(r, mul_mod(q, new_r));
        t = new_t; new_t = nt;
        r = new_r; new_r = nr;
    }
    if (!done) t_fail = 1; /* unwrap_left on Right */
    return mul_mod(a, t);
}
static uint32_t exp_mod(uint32_t a, uint32_t b) { /* field.simf:76-94 */
    uint32_t res = 1, base = a, e = b;
    for (uint32_t counter = 0; counter < 65536; counter++) {
        if (e == 0) return res;
        if (!jet_divides_32(2, e)) res = mul_mod(res, base);
        base = mul_mod(base, base);
        e = jet_divide_32(e, 2);
    }
    t_fail = 1;
    return res;
}
/* stark101/src/channel.simf */
static u256 s101_channel_mix_32(u256 st, uint32_t in) { /* channel.simf:22-27 */
    Ctx8 c = sha_256_ctx_8_init();
    c = sha_256_ctx_8_add_32(c, st);
    c = sha_256_ctx_8_add_4(c, in);
    return sha_256_ctx_8_finalize(c);
}
static u256 s101_channel_mix_256(u256 st, u256 in) { return sha256_pair(st, in); } /* channel.simf:35-40 */
static uint32_t reduce_limb_32_mod_32(uint32_t limb, uint32_t r, uint32_t modulo) { /* channel.simf:67-75 */
    uint64_t v = jet_left_shift_64(32, (uint64_t)r) + (uint64_t)limb;
    return (uint32_t)jet_modulo_64(v, (uint64_t)modulo);
}
static uint32_t reduce_256_mod_32(u256 v, uint32_t modulo) { /* channel.simf:83-94 */
    uint32_t r = 0;
    for (int i = 0; i < 8; i++) r = reduce_limb_32_mod_32(v.w[i], r, modulo);
    return r;
}
static uint32_t s101_channel_draw_32(u256 *st, uint32_t max) { /* channel.simf:102-105 */
    uint32_t v = reduce_256_mod_32(*st, max);
    *st = sha256(*st);
    return v;
}
/* stark101/src/merkle.simf:39-43 (no path == 1 assert) */
static u256 s101_merkle_verify_32(u256 leaf, uint32_t auth_path, const uint32_t *sib, uint32_t n_sib, u256 root) {
    uint32_t path;
    u256 c = merkle_fold(leaf, auth_path, sib, n_sib, &path);
    ORACLE_ASSERT(eq_256(c, root));
    return c;
}
/* air.simf */
static uint32_t fibsquare_calc_x(uint32_t idx) { return mul_mod(FIELD_GEN, exp_mod(CANONIC_COSET_GEN, idx)); } /* air.simf:58-60 */
static uint32_t fibsquare_eval_p0(uint32_t x, uint32_t f_x) { return div_mod(sub_mod(f_x, 1), sub_mod(x, 1)); } /* air.simf:63-66 */
static uint32_t fibsquare_eval_p1(uint32_t x, uint32_t f_x) { return div_mod(sub_mod(f_x, 2338775057u), sub_mod(x, 2450347685u)); } /* air.simf:69-72 */
static uint32_t fibsquare_eval_p2(uint32_t x, uint32_t f_x, uint32_t f_gx, uint32_t f_ggx) { /* air.simf:75-83 */
    uint32_t num0 = sub_mod(f_ggx, add_mod(mul_mod(f_x, f_x), mul_mod(f_gx, f_gx)));
    uint32_t num1 = mul_mod(mul_mod(sub_mod(x, 2342081930u), sub_mod(x, 2450347685u)), sub_mod(x, 532203874u));
    uint32_t den = sub_mod(exp_mod(x, 1024), 1);
    return div_mod(mul_mod(num0, num1), den);
}
static uint32_t fibsquare_eval_cp(uint32_t x, uint32_t a0, uint32_t a1, uint32_t a2, uint32_t f_x, uint32_t f_gx, uint32_t f_ggx) { /* air.simf:86-91 */
    uint32_t p0 = fibsquare_eval_p0(x, f_x);
    uint32_t p1 = fibsquare_eval_p1(x, f_x);
    uint32_t p2 = fibsquare_eval_p2(x, f_x, f_gx, f_ggx);
    return add_mod(add_mod(mul_mod(p0, a0), mul_mod(p1, a1)), mul_mod(p2, a2));
}
/* fri.simf */
static uint32_t fri_eval_cp_next(uint32_t cpa, uint32_t cpb, uint32_t x, uint32_t beta) { /* fri.simf:55-59 */
    uint32_t op0 = div_mod(add_mod(cpa, cpb), 2);
    uint32_t op1 = div_mod(sub_mod(cpa, cpb), mul_mod(x, 2));
    return add_mod(op0, mul_mod(op1, beta));
}
static void compute_auth_path(uint32_t idx, uint32_t domain_size, uint32_t *a, uint32_t *b) { /* fri.simf:63-68 */
    *a = jet_add_32(jet_modulo_32(idx, domain_size), domain_size);
    uint32_t cpb_idx = jet_add_32(idx, jet_divide_32(domain_size, 2));
    *b = jet_add_32(jet_modulo_32(cpb_idx, domain_size), domain_size);
}

/* verify_proof stark101/src/verifier.simf:24-42 on one packed record (include/ssym.h) */
EXPORT void oracle_s101_verify_one(const uint32_t *rec, ssym_s101_trace_t *tr) {
    memset(tr, 0, sizeof *tr);
    tr->first_fail_layer = 0xffffffffu;
    uint32_t total = rec[0], n_layers = rec[1];
    uint32_t ns[3] = {rec[2], rec[3], rec[4]};
    if (n_layers > SSYM_S101_MAX_LIST || ns[0] > SSYM_S101_MAX_LIST || ns[1] > SSYM_S101_MAX_LIST || ns[2] > SSYM_S101_MAX_LIST || total < 20 ||
        rec[6] > SSYM_S101_MAX_ORDINAL) {
        tr->status = SSYM_S101_ST_SHAPE;
        return;
    }
    uint32_t status = 0;
    uint32_t last_layer = rec[5];
    u256 root = load_u256(rec + 8);
    const uint32_t *ev_sib[3];
    uint32_t w = 20;
    for (int i = 0; i < 3; i++) { ev_sib[i] = rec + w; w += 8 * ns[i]; }
    const uint32_t *layers = rec + w;
    for (uint32_t l = 0; l < n_layers; l++) { /* record must be self-consistent */
        if (w + 16 > total || rec[w + 11] > SSYM_S101_MAX_LIST || rec[w + 12] > SSYM_S101_MAX_LIST) { tr->status = SSYM_S101_ST_SHAPE; return; }
        w += 16 + 8 * (rec[w + 11] + rec[w + 12]);
    }
    if (w != total) { tr->status = SSYM_S101_ST_SHAPE; return; }
    tr->query_ordinal = rec[6];

    u256 state = sha256(root); /* verifier.simf:27 */
    /* fibsquare_read_coefficients air.simf:30-36 */
    uint32_t a0 = s101_channel_draw_32(&state, FIELD_MODULUS);
    uint32_t a1 = s101_channel_draw_32(&state, FIELD_MODULUS);
    uint32_t a2 = s101_channel_draw_32(&state, FIELD_MODULUS);
    tr->alpha[0] = a0; tr->alpha[1] = a1; tr->alpha[2] = a2;
    /* fri_read_commitments_32 fri.simf:49-53 */
    tr->n_layers = n_layers;
    const uint32_t *lp = layers;
    for (uint32_t l = 0; l < n_layers; l++) { /* fri_read_commitment fri.simf:37-45 */
        state = s101_channel_mix_256(state, load_u256(lp));
        uint32_t random = s101_channel_draw_32(&state, FIELD_MODULUS);
        tr->beta_drawn[l] = random;
        if (random != lp[8]) { status |= SSYM_S101_ST_BETA; tr->layer_mask[l] |= 8; }
        lp += 16 + 8 * (lp[11] + lp[12]);
    }
    state = s101_channel_mix_32(state, last_layer);
    store_u256(tr->commit_state, state);
    /* verifier.simf:32; query ordinal k (include/ssym.h): the (k+1)-th draw, the query phase repeated on one channel */
    uint32_t idx = s101_channel_draw_32(&state, DOMAIN_EX_SIZE);
    for (uint32_t k = 0; k < rec[6]; k++) idx = s101_channel_draw_32(&state, DOMAIN_EX_SIZE);
    tr->idx = idx;
    /* fibsquare_read_evaluations_checked air.simf:47-56 */
    uint32_t f[3], cur_idx = idx;
    for (int i = 0; i < 3; i++) {
        f[i] = rec[16 + i];
        t_fail = 0;
        u256 c = s101_merkle_verify_32(sha256_32(f[i]), jet_add_32(cur_idx, DOMAIN_EX_SIZE), ev_sib[i], ns[i], root); /* air.simf:39-44 */
        if (t_fail) status |= SSYM_S101_ST_TRACE_MERKLE(i);
        store_u256(tr->trace_root[i], c);
        state = s101_channel_mix_32(state, f[i]);
        cur_idx = jet_add_32(cur_idx, IDX_OFFSET);
    }
    store_u256(tr->state_final, state);
    uint32_t x = fibsquare_calc_x(idx); /* verifier.simf:36 */
    tr->x = x;
    t_fail = 0;
    uint32_t cp_ev = fibsquare_eval_cp(x, a0, a1, a2, f[0], f[1], f[2]); /* verifier.simf:38 */
    if (t_fail) status |= SSYM_S101_ST_DIV;
    tr->cp0 = cp_ev;
    /* fri_verify_32 fri.simf:87-91 */
    uint32_t domain_size = DOMAIN_EX_SIZE;
    lp = layers;
    for (uint32_t l = 0; l < n_layers; l++) { /* fri_verify_layer fri.simf:71-84 */
        u256 lroot = load_u256(lp);
        uint32_t beta = lp[8], cpa = lp[9], cpb = lp[10], na = lp[11], nb = lp[12];
        const uint32_t *sa = lp + 16, *sb = lp + 16 + 8 * na;
        tr->cp_ev[l] = cp_ev;
        if (cp_ev != cpa) { status |= SSYM_S101_ST_LAYER_CP; tr->layer_mask[l] |= 1; }
        uint32_t pa, pb;
        compute_auth_path(idx, domain_size, &pa, &pb);
        t_fail = 0;
        s101_merkle_verify_32(sha256_32(cpa), pa, sa, na, lroot);
        if (t_fail) { status |= SSYM_S101_ST_LAYER_MERKLE_A; tr->layer_mask[l] |= 2; }
        t_fail = 0;
        s101_merkle_verify_32(sha256_32(cpb), pb, sb, nb, lroot);
        if (t_fail) { status |= SSYM_S101_ST_LAYER_MERKLE_B; tr->layer_mask[l] |= 4; }
        t_fail = 0;
        cp_ev = fri_eval_cp_next(cpa, cpb, x, beta);
        if (t_fail) { status |= SSYM_S101_ST_DIV; tr->layer_mask[l] |= 16; }
        x = mul_mod(x, x);
        domain_size = jet_divide_32(domain_size, 2);
        lp += 16 + 8 * (na + nb);
    }
    tr->cp_ev[n_layers] = cp_ev;
    if (cp_ev != last_layer) status |= SSYM_S101_ST_LAST; /* fri.simf:90 */
    for (uint32_t l = 0; l < n_layers; l++)
        if (tr->layer_mask[l]) { tr->first_fail_layer = l; break; }
    tr->status = status;
}

EXPORT void oracle_s101_verify_batch(const uint32_t *blob, const uint64_t *offsets, size_t begin, size_t end, uint32_t *accept_bits,
                                     uint32_t *status, ssym_s101_trace_t *trace) {
    ssym_s101_trace_t tmp;
    for (size_t i = begin; i < end; i++) {
        ssym_s101_trace_t *tr = trace ? &trace[i] : &tmp;
        oracle_s101_verify_one(blob + offsets[i], tr);
        if (status) status[i] = tr->status;
        if (accept_bits) {
            if (tr->status == 0) __atomic_fetch_or(&accept_bits[i / 32], 1u << (i % 32), __ATOMIC_RELAXED);
            else __atomic_fetch_and(&accept_bits[i / 32], ~(1u << (i % 32)), __ATOMIC_RELAXED);
        }
    }
}

/* ssym_stark101_verify_multi_batch (include/ssym.h): records proof-major, n_queries per proof; per-record status (with SSYM_S101_ST_GROUP),
 * one accept bit per proof.  `trace` as produced by oracle_s101_verify_batch over all n_proofs * n_queries records (its status fields are updated). */
EXPORT void oracle_s101_group(size_t n_proofs, uint32_t n_queries, uint32_t *status, ssym_s101_trace_t *trace, uint32_t *accept_bits) {
    for (size_t i = 0; i < n_proofs; i++) {
        int ok = 1;
        const ssym_s101_trace_t *first = &trace[i * n_queries];
        for (uint32_t k = 0; k < n_queries; k++) {
            ssym_s101_trace_t *tr = &trace[i * n_queries + k];
            if (!(tr->status & SSYM_S101_ST_SHAPE) &&
                (tr->query_ordinal != k || (first->status & SSYM_S101_ST_SHAPE) || memcmp(tr->commit_state, first->commit_state, 32) != 0))
                tr->status |= SSYM_S101_ST_GROUP;
            if (status) status[i * n_queries + k] = tr->status;
            if (tr->status) ok = 0;
        }
        if (ok) accept_bits[i / 32] |= 1u << (i % 32);
        else accept_bits[i / 32] &= ~(1u << (i % 32));
    }
}

EXPORT uint32_t oracle_s101_add_mod(uint32_t a, uint32_t b) { return add_mod(a, b); }
EXPORT uint32_t oracle_s101_sub_mod(uint32_t a, uint32_t b) { return sub_mod(a, b); }
EXPORT uint32_t oracle_s101_mul_mod(uint32_t a, uint32_t b) { return mul_mod(a, b); }
EXPORT uint32_t oracle_s101_div_mod(uint32_t a, uint32_t b, int *fail) { t_fail = 0; uint32_t r = div_mod(a, b); if (fail) *fail = t_fail; return r; }
EXPORT uint32_t oracle_s101_exp_mod(uint32_t a, uint32_t b) { return exp_mod(a, b); }
EXPORT uint32_t oracle_s101_channel_draw_32(uint32_t *st, uint32_t max) { u256 s = load_u256(st); uint32_t v = s101_channel_draw_32(&s, max); store_u256(st, s); return v; }
EXPORT void oracle_s101_channel_mix_32(uint32_t *st, uint32_t in) { store_u256(st, s101_channel_mix_32(load_u256(st), in)); }
EXPORT void oracle_s101_channel_mix_256(uint32_t *st, const uint32_t *in) { store_u256(st, s101_channel_mix_256(load_u256(st), load_u256(in))); }
EXPORT int oracle_s101_merkle_verify_32(const uint32_t *leaf, uint32_t auth, const uint32_t *sib, uint32_t n_sib, const uint32_t *root) {
    t_fail = 0;
    s101_merkle_verify_32(load_u256(leaf), auth, sib, n_sib, load_u256(root));
    return !t_fail;
}
EXPORT uint32_t oracle_s101_calc_x(uint32_t idx) { return fibsquare_calc_x(idx); }
EXPORT uint32_t oracle_s101_eval_p0(uint32_t x, uint32_t f_x) { return fibsquare_eval_p0(x, f_x); }
EXPORT uint32_t oracle_s101_eval_cp(uint32_t x, uint32_t a0, uint32_t a1, uint32_t a2, uint32_t f_x, uint32_t f_gx, uint32_t f_ggx) { return fibsquare_eval_cp(x, a0, a1, a2, f_x, f_gx, f_ggx); }
EXPORT uint32_t oracle_s101_fri_eval_cp_next(uint32_t cpa, uint32_t cpb, uint32_t x, uint32_t beta) { return fri_eval_cp_next(cpa, cpb, x, beta); }
EXPORT void oracle_s101_compute_auth_path(uint32_t idx, uint32_t domain_size, uint32_t *o) { compute_auth_path(idx, domain_size, &o[0], &o[1]); }
/* fri_read_commitment fri.simf:37-45 on a bare (root, beta); returns 1 iff the drawn beta equals the witness beta */
EXPORT int oracle_s101_fri_read_commitment(uint32_t *st, const uint32_t *root, uint32_t beta) {
    u256 s = s101_channel_mix_256(load_u256(st), load_u256(root));
    uint32_t random = s101_channel_draw_32(&s, FIELD_MODULUS);
    store_u256(st, s);
    return random == beta;
}
EXPORT size_t oracle_sizeof_stwo_trace(void) { return sizeof(ssym_stwo_trace_t); }
EXPORT size_t oracle_sizeof_s101_trace(void) { return sizeof(ssym_s101_trace_t); }

/* CPU reference prover for the wide-Fibonacci AIR (test infrastructure; the checker of the GPU prover). */
#include "stwo_prover_ref.c"
