// SPDX-License-Identifier: MIT
//
// Device SHA-256 for the commitment hash of stark-symphony (the sha_256_ctx_8_* jets as used by
// stwo-verifier/src/hasher.simf:13-104, channel.simf:36-172 and stark101/src/sha256.simf:11-29).
//
// Everything lives in registers: 8 state words + a rolling 16-word message schedule, rounds fully
// unrolled so K[t] (and, for constant blocks, K[t]+W[t]) become immediates.  Rotates are funnel
// shifts (SHF.R.W), Ch/Maj/xor3 are single LOP3s, the 5-operand round sum is two IADD3s.
//
// A digest is 8 uint32_t, word 0 = most significant (big-endian u256, channel.simf:48-58).
#pragma once
#include <stdint.h>

namespace ssym {

// constexpr copy so fully unrolled code sees immediates
struct ShaK {
    uint32_t k[64];
};
__host__ __device__ constexpr ShaK sha_k_table() {
    return ShaK{{0x428a2f98, 0x71374491, 0xb5c0fbcf, 0xe9b5dba5, 0x3956c25b, 0x59f111f1, 0x923f82a4, 0xab1c5ed5,
                 0xd807aa98, 0x12835b01, 0x243185be, 0x550c7dc3, 0x72be5d74, 0x80deb1fe, 0x9bdc06a7, 0xc19bf174,
                 0xe49b69c1, 0xefbe4786, 0x0fc19dc6, 0x240ca1cc, 0x2de92c6f, 0x4a7484aa, 0x5cb0a9dc, 0x76f988da,
                 0x983e5152, 0xa831c66d, 0xb00327c8, 0xbf597fc7, 0xc6e00bf3, 0xd5a79147, 0x06ca6351, 0x14292967,
                 0x27b70a85, 0x2e1b2138, 0x4d2c6dfc, 0x53380d13, 0x650a7354, 0x766a0abb, 0x81c2c92e, 0x92722c85,
                 0xa2bfe8a1, 0xa81a664b, 0xc24b8b70, 0xc76c51a3, 0xd192e819, 0xd6990624, 0xf40e3585, 0x106aa070,
                 0x19a4c116, 0x1e376c08, 0x2748774c, 0x34b0bcb5, 0x391c0cb3, 0x4ed8aa4a, 0x5b9cca4f, 0x682e6ff3,
                 0x748f82ee, 0x78a5636f, 0x84c87814, 0x8cc70208, 0x90befffa, 0xa4506ceb, 0xbef9a3f7, 0xc67178f2}};
}

__host__ __device__ constexpr uint32_t c_rotr(uint32_t x, int n) { return (x >> n) | (x << (32 - n)); }
__host__ __device__ constexpr uint32_t c_s0(uint32_t x) { return c_rotr(x, 7) ^ c_rotr(x, 18) ^ (x >> 3); }
__host__ __device__ constexpr uint32_t c_s1(uint32_t x) { return c_rotr(x, 17) ^ c_rotr(x, 19) ^ (x >> 10); }

// K[t] + W[t] for the padding block that follows a 64-byte message (0x80, zeros, bit length 512):
// the second compression of sha256_pair / of a 64-byte leaf has a constant schedule.
struct ShaKW {
    uint32_t kw[64];
};
__host__ __device__ constexpr ShaKW sha_pad64_kw() {
    ShaKW r{};
    uint32_t w[64] = {};
    w[0] = 0x80000000u;
    w[15] = 512u;
    for (int t = 16; t < 64; t++) w[t] = w[t - 16] + c_s0(w[t - 15]) + w[t - 7] + c_s1(w[t - 2]);
    ShaK k = sha_k_table();
    for (int t = 0; t < 64; t++) r.kw[t] = k.k[t] + w[t];
    return r;
}

__device__ __forceinline__ uint32_t rotr32(uint32_t x, int n) { return __funnelshift_r(x, x, n); }
__device__ __forceinline__ uint32_t Sig0(uint32_t x) { return rotr32(x, 2) ^ rotr32(x, 13) ^ rotr32(x, 22); }
__device__ __forceinline__ uint32_t Sig1(uint32_t x) { return rotr32(x, 6) ^ rotr32(x, 11) ^ rotr32(x, 25); }
__device__ __forceinline__ uint32_t sig0(uint32_t x) { return rotr32(x, 7) ^ rotr32(x, 18) ^ (x >> 3); }
__device__ __forceinline__ uint32_t sig1(uint32_t x) { return rotr32(x, 17) ^ rotr32(x, 19) ^ (x >> 10); }
__device__ __forceinline__ uint32_t Ch(uint32_t e, uint32_t f, uint32_t g) { return (e & f) ^ (~e & g); }
__device__ __forceinline__ uint32_t Maj(uint32_t a, uint32_t b, uint32_t c) { return (a & b) ^ (a & c) ^ (b & c); }

#define SSYM_SHA_IV0 0x6a09e667u
#define SSYM_SHA_IV1 0xbb67ae85u
#define SSYM_SHA_IV2 0x3c6ef372u
#define SSYM_SHA_IV3 0xa54ff53au
#define SSYM_SHA_IV4 0x510e527fu
#define SSYM_SHA_IV5 0x9b05688cu
#define SSYM_SHA_IV6 0x1f83d9abu
#define SSYM_SHA_IV7 0x5be0cd19u

__device__ __forceinline__ void sha_iv(uint32_t (&h)[8]) {
    h[0] = SSYM_SHA_IV0; h[1] = SSYM_SHA_IV1; h[2] = SSYM_SHA_IV2; h[3] = SSYM_SHA_IV3;
    h[4] = SSYM_SHA_IV4; h[5] = SSYM_SHA_IV5; h[6] = SSYM_SHA_IV6; h[7] = SSYM_SHA_IV7;
}

// ---- pipe balancing -------------------------------------------------------------------------------------
// On sm_100 the rotates / LOP3s / IADD3s of SHA-256 all issue to the ALU pipe (64 lanes/clk/SM), which is what
// bounds the Merkle kernels (ncu: alu pipe 94 % busy, fma pipe 6 %).  An integer add can instead be issued as
// IMAD x*1+y on the otherwise idle FMA pipe — but only if ptxas cannot see that the multiplier is 1, so the `1`
// arrives as a kernel parameter (`one`).  ADDMODE selects which adds are moved:
//   0  plain C adds (ptxas decides; it picks IADD3/VIADD for ~70 % of them)
//   1  every add of the round function and of the message schedule as IMAD
//   2  round-function adds as IMAD, message-schedule adds left to ptxas
//   3  only the T1 chain (h + K + W + Sigma1 + Ch) as IMAD
//   4  as 1, but a' = T1 + Sigma0 + Maj is ONE 3-input IADD3 on the ALU pipe instead of two IMADs: with everything on the FMA pipe the
//      kernel is bound by total issue slots (ncu: 0.72 IPC), so trading two FMA-pipe instructions for one ALU-pipe instruction pays
//   5  as 4, and the message schedule's w + sigma0 + w[-7] is an IADD3 as well
//   6  as 1, and the two plain shifts of the message schedule (x >> 3, x >> 10) are IMAD.HI by 2^29 / 2^22 on the FMA pipe
//      (a 1:1 trade of an ALU-pipe SHF for an FMA-pipe instruction); the multipliers arrive as kernel parameters as well
//   7  as 6, and one rotation of each big Sigma is built on the FMA pipe: rotr(x, n) = x * 2^(32-n) + hi(x * 2^(32-n))
//      (IMAD.HI + IMAD replace one SHF)
//   8  as 1, with a second opaque 1 for the additions of a constant-bank word (ShaAdd::tk)
//   9  as 8, with the plain shifts of the message schedule as IMAD.HI (mode 6's trade)
//   10 as 8, with one rotation of Sigma1 as IMAD.HI + IMAD (half of mode 7's trade)
struct ShaMul { // opaque (kernel-parameter) constants: 1, 2^29, 2^22, 2^10, 2^7, and a second 1 (ADDMODE 8)
    uint32_t one, m29, m22, m10, m7, onek;
};
__host__ __device__ inline ShaMul sha_mul_consts() { return ShaMul{1u, 1u << 29, 1u << 22, 1u << 10, 1u << 7, 1u}; }
template <int ADDMODE>
struct ShaAdd {
    uint32_t one, m29, m22, m10, m7, onek;
    __device__ __forceinline__ ShaAdd(uint32_t o) : one(o), m29(0), m22(0), m10(0), m7(0), onek(o) {}
    __device__ __forceinline__ ShaAdd(const ShaMul &m) : one(m.one), m29(m.m29), m22(m.m22), m10(m.m10), m7(m.m7), onek(m.onek) {}
    __device__ __forceinline__ uint32_t fma(uint32_t a, uint32_t b) const {
        uint32_t d;
        asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(one), "r"(b));
        return d;
    }
    static __device__ __forceinline__ uint32_t mulhi(uint32_t a, uint32_t m) {
        uint32_t d;
        asm("mul.hi.u32 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(m));
        return d;
    }
    static __device__ __forceinline__ uint32_t rot_fma(uint32_t x, uint32_t m) { // rotr(x, n) with m = 2^(32-n)
        uint32_t d;
        asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(x), "r"(m), "r"(mulhi(x, m)));
        return d;
    }
    __device__ __forceinline__ uint32_t ssig0(uint32_t x) const { return rotr32(x, 7) ^ rotr32(x, 18) ^ ((ADDMODE == 6 || ADDMODE == 7 || ADDMODE == 9) ? mulhi(x, m29) : (x >> 3)); }
    __device__ __forceinline__ uint32_t ssig1(uint32_t x) const { return rotr32(x, 17) ^ rotr32(x, 19) ^ ((ADDMODE == 6 || ADDMODE == 7 || ADDMODE == 9) ? mulhi(x, m22) : (x >> 10)); }
    __device__ __forceinline__ uint32_t bSig0(uint32_t x) const { return rotr32(x, 2) ^ rotr32(x, 13) ^ (ADDMODE == 7 ? rot_fma(x, m10) : rotr32(x, 22)); }
    __device__ __forceinline__ uint32_t bSig1(uint32_t x) const { return rotr32(x, 6) ^ rotr32(x, 11) ^ ((ADDMODE == 7 || ADDMODE == 10) ? rot_fma(x, m7) : rotr32(x, 25)); }
    __device__ __forceinline__ uint32_t t1(uint32_t a, uint32_t b) const { return ADDMODE >= 1 ? fma(a, b) : a + b; }    // T1 chain
    // the additions whose addend is K[t] / K[t] + W[t] from the constant bank: an IMAD takes ONE operand from the uniform datapath, so here the
    // multiplier 1 has to sit in a vector register.  With the same `one` everywhere ptxas then reads it from that register in EVERY addition
    // (three vector-register operands each); a second opaque 1 for these leaves the others at two vector operands + a uniform register.
    __device__ __forceinline__ uint32_t tk(uint32_t a, uint32_t k) const {
        if (ADDMODE < 8) return t1(a, k);
        uint32_t d;
        asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(onek), "r"(k));
        return d;
    }
    __device__ __forceinline__ uint32_t rnd(uint32_t a, uint32_t b) const { return (ADDMODE == 1 || ADDMODE == 2 || ADDMODE >= 4) ? fma(a, b) : a + b; } // rest of the round
    __device__ __forceinline__ uint32_t sch(uint32_t a, uint32_t b) const { return (ADDMODE == 1 || ADDMODE >= 4) ? fma(a, b) : a + b; }  // message schedule
    // a' = t1 + Sigma0 + Maj
    __device__ __forceinline__ uint32_t rnd3(uint32_t t1v, uint32_t s0, uint32_t mj) const { return (ADDMODE == 4 || ADDMODE == 5) ? t1v + s0 + mj : rnd(t1v, rnd(s0, mj)); }
    // w + sigma0 + w9 + sigma1
    __device__ __forceinline__ uint32_t sch4(uint32_t w, uint32_t s0, uint32_t w9, uint32_t s1) const {
        return ADDMODE == 5 ? fma(w + s0 + w9, s1) : sch(sch(w, s0), sch(w9, s1));
    }
};

// T1 is associated as ((h + K + W) + Ch) + Sigma1: Sigma1(e) (two dependent ALU operations) is the last term to arrive, so only ONE addition sits
// between it and the new e — the dependent chain of a round is SHF, LOP3, IMAD, IMAD instead of SHF, LOP3, IMAD, IMAD, IMAD.  Same operation
// count; it matters where a lone warp runs a chain of compressions (the transcript kernel), not in the throughput-bound Merkle kernels.
#define SSYM_SHA_ROUND(A, a, b, c, d, e, f, g, h, kw)                        \
    {                                                                        \
        uint32_t t1_ = A.t1(A.t1(A.t1(h, kw), Ch(e, f, g)), A.bSig1(e));     \
        (d) = A.rnd(d, t1_);                                                 \
        (h) = A.rnd3(t1_, A.bSig0(a), Maj(a, b, c));                         \
    }
// the same round with K[t] + W[t] read from the constant bank (the padding block)
#define SSYM_SHA_ROUNDK(A, a, b, c, d, e, f, g, h, kw)                       \
    {                                                                        \
        uint32_t t1_ = A.t1(A.t1(A.tk(h, kw), Ch(e, f, g)), A.bSig1(e));     \
        (d) = A.rnd(d, t1_);                                                 \
        (h) = A.rnd3(t1_, A.bSig0(a), Maj(a, b, c));                         \
    }

// One compression of `h` with the 16-word block `w` (w is consumed: it becomes the rolling schedule).
template <int ADDMODE>
__device__ __forceinline__ void sha_compress(uint32_t (&h)[8], uint32_t (&w)[16], const ShaAdd<ADDMODE> A) {
    constexpr ShaK K = sha_k_table();
    uint32_t a = h[0], b = h[1], c = h[2], d = h[3], e = h[4], f = h[5], g = h[6], hh = h[7];
#pragma unroll
    for (int t = 0; t < 64; t += 8) {
        if (t >= 16) {
#pragma unroll
            for (int j = 0; j < 8; j++) {
                const int i = (t + j) & 15;
                w[i] = A.sch4(w[i], A.ssig0(w[(i + 1) & 15]), w[(i + 9) & 15], A.ssig1(w[(i + 14) & 15]));
            }
        }
        SSYM_SHA_ROUND(A, a, b, c, d, e, f, g, hh, A.tk(w[(t + 0) & 15], K.k[t + 0]));
        SSYM_SHA_ROUND(A, hh, a, b, c, d, e, f, g, A.tk(w[(t + 1) & 15], K.k[t + 1]));
        SSYM_SHA_ROUND(A, g, hh, a, b, c, d, e, f, A.tk(w[(t + 2) & 15], K.k[t + 2]));
        SSYM_SHA_ROUND(A, f, g, hh, a, b, c, d, e, A.tk(w[(t + 3) & 15], K.k[t + 3]));
        SSYM_SHA_ROUND(A, e, f, g, hh, a, b, c, d, A.tk(w[(t + 4) & 15], K.k[t + 4]));
        SSYM_SHA_ROUND(A, d, e, f, g, hh, a, b, c, A.tk(w[(t + 5) & 15], K.k[t + 5]));
        SSYM_SHA_ROUND(A, c, d, e, f, g, hh, a, b, A.tk(w[(t + 6) & 15], K.k[t + 6]));
        SSYM_SHA_ROUND(A, b, c, d, e, f, g, hh, a, A.tk(w[(t + 7) & 15], K.k[t + 7]));
    }
    h[0] += a; h[1] += b; h[2] += c; h[3] += d; h[4] += e; h[5] += f; h[6] += g; h[7] += hh;
}
__device__ __forceinline__ void sha_compress(uint32_t (&h)[8], uint32_t (&w)[16]) { sha_compress<0>(h, w, ShaAdd<0>(1u)); }

// Compression of the constant padding block that ends every 64-byte message.
template <int ADDMODE>
__device__ __forceinline__ void sha_compress_pad64(uint32_t (&h)[8], const ShaAdd<ADDMODE> A) {
    constexpr ShaKW KW = sha_pad64_kw();
    uint32_t a = h[0], b = h[1], c = h[2], d = h[3], e = h[4], f = h[5], g = h[6], hh = h[7];
#pragma unroll
    for (int t = 0; t < 64; t += 8) {
        SSYM_SHA_ROUNDK(A, a, b, c, d, e, f, g, hh, KW.kw[t + 0]);
        SSYM_SHA_ROUNDK(A, hh, a, b, c, d, e, f, g, KW.kw[t + 1]);
        SSYM_SHA_ROUNDK(A, g, hh, a, b, c, d, e, f, KW.kw[t + 2]);
        SSYM_SHA_ROUNDK(A, f, g, hh, a, b, c, d, e, KW.kw[t + 3]);
        SSYM_SHA_ROUNDK(A, e, f, g, hh, a, b, c, d, KW.kw[t + 4]);
        SSYM_SHA_ROUNDK(A, d, e, f, g, hh, a, b, c, KW.kw[t + 5]);
        SSYM_SHA_ROUNDK(A, c, d, e, f, g, hh, a, b, KW.kw[t + 6]);
        SSYM_SHA_ROUNDK(A, b, c, d, e, f, g, hh, a, KW.kw[t + 7]);
    }
    h[0] += a; h[1] += b; h[2] += c; h[3] += d; h[4] += e; h[5] += f; h[6] += g; h[7] += hh;
}

// ---- rolled variants --------------------------------------------------------------------------------------
// The fully unrolled pair hash is ~2300 instructions = 37 KB of SASS, more than the 32 KB L1.5 instruction cache:
// ncu shows 11-19 % instruction-cache misses and `no_instruction` stalls in the Merkle kernels.  The rolled form
// keeps 16 rounds (one period of the rolling message schedule, two rotations of the 8 working variables) unrolled
// and loops over the four groups, reading K[t] (or K[t]+W[t] of the padding block) from constant memory.
struct ShaK4 {
    uint4 v[16];
};
static __constant__ ShaK4 c_sha_k4 = {{
    {0x428a2f98, 0x71374491, 0xb5c0fbcf, 0xe9b5dba5}, {0x3956c25b, 0x59f111f1, 0x923f82a4, 0xab1c5ed5},
    {0xd807aa98, 0x12835b01, 0x243185be, 0x550c7dc3}, {0x72be5d74, 0x80deb1fe, 0x9bdc06a7, 0xc19bf174},
    {0xe49b69c1, 0xefbe4786, 0x0fc19dc6, 0x240ca1cc}, {0x2de92c6f, 0x4a7484aa, 0x5cb0a9dc, 0x76f988da},
    {0x983e5152, 0xa831c66d, 0xb00327c8, 0xbf597fc7}, {0xc6e00bf3, 0xd5a79147, 0x06ca6351, 0x14292967},
    {0x27b70a85, 0x2e1b2138, 0x4d2c6dfc, 0x53380d13}, {0x650a7354, 0x766a0abb, 0x81c2c92e, 0x92722c85},
    {0xa2bfe8a1, 0xa81a664b, 0xc24b8b70, 0xc76c51a3}, {0xd192e819, 0xd6990624, 0xf40e3585, 0x106aa070},
    {0x19a4c116, 0x1e376c08, 0x2748774c, 0x34b0bcb5}, {0x391c0cb3, 0x4ed8aa4a, 0x5b9cca4f, 0x682e6ff3},
    {0x748f82ee, 0x78a5636f, 0x84c87814, 0x8cc70208}, {0x90befffa, 0xa4506ceb, 0xbef9a3f7, 0xc67178f2}}};
// K[t] + W[t] of the 64-byte padding block (generated from sha_pad64_kw(); checked by a static_assert below)
static __constant__ ShaK4 c_sha_kwpad4 = {{
    {0xc28a2f98, 0x71374491, 0xb5c0fbcf, 0xe9b5dba5}, {0x3956c25b, 0x59f111f1, 0x923f82a4, 0xab1c5ed5},
    {0xd807aa98, 0x12835b01, 0x243185be, 0x550c7dc3}, {0x72be5d74, 0x80deb1fe, 0x9bdc06a7, 0xc19bf374},
    {0x649b69c1, 0xf0fe4786, 0x0fe1edc6, 0x240cf254}, {0x4fe9346f, 0x6cc984be, 0x61b9411e, 0x16f988fa},
    {0xf2c65152, 0xa88e5a6d, 0xb019fc65, 0xb9d99ec7}, {0x9a1231c3, 0xe70eeaa0, 0xfdb1232b, 0xc7353eb0},
    {0x3069bad5, 0xcb976d5f, 0x5a0f118f, 0xdc1eeefd}, {0x0a35b689, 0xde0b7a04, 0x58f4ca9d, 0xe15d5b16},
    {0x007f3e86, 0x37088980, 0xa507ea32, 0x6fab9537}, {0x17406110, 0x0d8cd6f1, 0xcdaa3b6d, 0xc0bbbe37},
    {0x83613bda, 0xdb48a363, 0x0b02e931, 0x6fd15ca7}, {0x521afaca, 0x31338431, 0x6ed41a95, 0x6d437890},
    {0xc39c91f2, 0x9eccabbd, 0xb5c9a0e6, 0x532fb63c}, {0xd2c741c6, 0x07237ea3, 0xa4954b68, 0x4c191d76}}};

// ADDMODE 8: the multiplier 1 of the round additions, read per 16-round group with a uniform index — a load ptxas keeps in a uniform register
// (it moves a kernel parameter it has hoisted out of the loop into a vector register instead)
static __constant__ uint32_t c_sha_ones[4] = {1u, 1u, 1u, 1u};
#define SSYM_SHA_ROUND4(A, a, b, c, d, e, f, g, h, k0, k1, k2, k3) \
    SSYM_SHA_ROUND(A, a, b, c, d, e, f, g, h, k0)                   \
    SSYM_SHA_ROUND(A, h, a, b, c, d, e, f, g, k1)                   \
    SSYM_SHA_ROUND(A, g, h, a, b, c, d, e, f, k2)                   \
    SSYM_SHA_ROUND(A, f, g, h, a, b, c, d, e, k3)
#define SSYM_SHA_ROUND4K(A, a, b, c, d, e, f, g, h, k0, k1, k2, k3) \
    SSYM_SHA_ROUNDK(A, a, b, c, d, e, f, g, h, k0)                   \
    SSYM_SHA_ROUNDK(A, h, a, b, c, d, e, f, g, k1)                   \
    SSYM_SHA_ROUNDK(A, g, h, a, b, c, d, e, f, k2)                   \
    SSYM_SHA_ROUNDK(A, f, g, h, a, b, c, d, e, k3)

template <int ADDMODE>
__device__ __forceinline__ void sha_compress_rolled(uint32_t (&h)[8], uint32_t (&w)[16], const ShaAdd<ADDMODE> A0) {
    uint32_t a = h[0], b = h[1], c = h[2], d = h[3], e = h[4], f = h[5], g = h[6], hh = h[7];
    ShaAdd<ADDMODE> A = A0;
#pragma unroll 1
    for (int grp = 0; grp < 4; grp++) {
        if (ADDMODE >= 8) A.one = c_sha_ones[grp];
        if (grp) {
#pragma unroll
            for (int i = 0; i < 16; i++)
                w[i] = A.sch4(w[i], A.ssig0(w[(i + 1) & 15]), w[(i + 9) & 15], A.ssig1(w[(i + 14) & 15]));
        }
        const uint4 k0 = c_sha_k4.v[grp * 4 + 0], k1 = c_sha_k4.v[grp * 4 + 1], k2 = c_sha_k4.v[grp * 4 + 2], k3 = c_sha_k4.v[grp * 4 + 3];
        SSYM_SHA_ROUND4(A, a, b, c, d, e, f, g, hh, A.tk(w[0], k0.x), A.tk(w[1], k0.y), A.tk(w[2], k0.z), A.tk(w[3], k0.w));
        SSYM_SHA_ROUND4(A, e, f, g, hh, a, b, c, d, A.tk(w[4], k1.x), A.tk(w[5], k1.y), A.tk(w[6], k1.z), A.tk(w[7], k1.w));
        SSYM_SHA_ROUND4(A, a, b, c, d, e, f, g, hh, A.tk(w[8], k2.x), A.tk(w[9], k2.y), A.tk(w[10], k2.z), A.tk(w[11], k2.w));
        SSYM_SHA_ROUND4(A, e, f, g, hh, a, b, c, d, A.tk(w[12], k3.x), A.tk(w[13], k3.y), A.tk(w[14], k3.z), A.tk(w[15], k3.w));
    }
    h[0] += a; h[1] += b; h[2] += c; h[3] += d; h[4] += e; h[5] += f; h[6] += g; h[7] += hh;
}
// NP independent compressions in lockstep (same rolled structure, the rounds of the NP states interleaved in one instruction stream):
// a lone warp per scheduler is bound by the dependency chain of a compression (ncu: 0.34 instructions per cycle, every issue followed by
// ~1.2 cycles of fixed-latency wait); a second independent chain fills those slots.  Used by the transcript kernel.
template <int ADDMODE, int NP>
__device__ __forceinline__ void sha_compress_rolled_n(uint32_t (&h)[NP][8], uint32_t (&w)[NP][16], const ShaAdd<ADDMODE> A) {
    uint32_t v[NP][8];
#pragma unroll
    for (int p = 0; p < NP; p++)
#pragma unroll
        for (int k = 0; k < 8; k++) v[p][k] = h[p][k];
#pragma unroll 1
    for (int grp = 0; grp < 4; grp++) {
        if (grp) {
#pragma unroll
            for (int i = 0; i < 16; i++)
#pragma unroll
                for (int p = 0; p < NP; p++) w[p][i] = A.sch4(w[p][i], A.ssig0(w[p][(i + 1) & 15]), w[p][(i + 9) & 15], A.ssig1(w[p][(i + 14) & 15]));
        }
        const uint4 kq[4] = {c_sha_k4.v[grp * 4 + 0], c_sha_k4.v[grp * 4 + 1], c_sha_k4.v[grp * 4 + 2], c_sha_k4.v[grp * 4 + 3]};
#pragma unroll
        for (int t = 0; t < 16; t++) { // round t of the group: the working variables rotate through v[][(j - t) & 7]
            const uint4 k4 = kq[t >> 2];
            const uint32_t kt = (t & 3) == 0 ? k4.x : (t & 3) == 1 ? k4.y : (t & 3) == 2 ? k4.z : k4.w;
#pragma unroll
            for (int p = 0; p < NP; p++) {
                uint32_t &a = v[p][(0 - t) & 7], &b = v[p][(1 - t) & 7], &c = v[p][(2 - t) & 7], &d = v[p][(3 - t) & 7];
                uint32_t &e = v[p][(4 - t) & 7], &f = v[p][(5 - t) & 7], &g = v[p][(6 - t) & 7], &hh = v[p][(7 - t) & 7];
                SSYM_SHA_ROUND(A, a, b, c, d, e, f, g, hh, A.tk(w[p][t], kt));
            }
        }
    }
#pragma unroll
    for (int p = 0; p < NP; p++)
#pragma unroll
        for (int k = 0; k < 8; k++) h[p][k] += v[p][k];
}
template <int ADDMODE>
__device__ __forceinline__ void sha_compress_pad64_rolled(uint32_t (&h)[8], const ShaAdd<ADDMODE> A0) {
    uint32_t a = h[0], b = h[1], c = h[2], d = h[3], e = h[4], f = h[5], g = h[6], hh = h[7];
    ShaAdd<ADDMODE> A = A0;
#pragma unroll 1
    for (int grp = 0; grp < 4; grp++) {
        if (ADDMODE >= 8) A.one = c_sha_ones[grp];
        const uint4 k0 = c_sha_kwpad4.v[grp * 4 + 0], k1 = c_sha_kwpad4.v[grp * 4 + 1], k2 = c_sha_kwpad4.v[grp * 4 + 2], k3 = c_sha_kwpad4.v[grp * 4 + 3];
        SSYM_SHA_ROUND4K(A, a, b, c, d, e, f, g, hh, k0.x, k0.y, k0.z, k0.w);
        SSYM_SHA_ROUND4K(A, e, f, g, hh, a, b, c, d, k1.x, k1.y, k1.z, k1.w);
        SSYM_SHA_ROUND4K(A, a, b, c, d, e, f, g, hh, k2.x, k2.y, k2.z, k2.w);
        SSYM_SHA_ROUND4K(A, e, f, g, hh, a, b, c, d, k3.x, k3.y, k3.z, k3.w);
    }
    h[0] += a; h[1] += b; h[2] += c; h[3] += d; h[4] += e; h[5] += f; h[6] += g; h[7] += hh;
}
template <int ADDMODE>
__device__ __forceinline__ void sha256_64B_rolled(uint32_t (&w)[16], uint32_t (&out)[8], const ShaAdd<ADDMODE> A) {
    sha_iv(out);
    sha_compress_rolled<ADDMODE>(out, w, A);
    sha_compress_pad64_rolled<ADDMODE>(out, A);
}

// SHA-256 of a 64-byte message given as 16 big-endian words (2 compressions).  Used by
// sha256_pair (hasher.simf:27-32), channel_mix_u256 (channel.simf:154-162) and the CP leaf
// (hash_node_m31_cp, hasher.simf:93-97).  `w` is clobbered.
template <int ADDMODE>
__device__ __forceinline__ void sha256_64B(uint32_t (&w)[16], uint32_t (&out)[8], const ShaAdd<ADDMODE> A) {
    sha_iv(out);
    sha_compress<ADDMODE>(out, w, A);
    sha_compress_pad64<ADDMODE>(out, A);
}
__device__ __forceinline__ void sha256_64B(uint32_t (&w)[16], uint32_t (&out)[8]) { sha256_64B<0>(w, out, ShaAdd<0>(1u)); }

// sha256_pair(left, right)
template <int ADDMODE>
__device__ __forceinline__ void sha256_pair(const uint32_t (&l)[8], const uint32_t (&r)[8], uint32_t (&out)[8], const ShaAdd<ADDMODE> A) {
    uint32_t w[16];
#pragma unroll
    for (int i = 0; i < 8; i++) { w[i] = l[i]; w[8 + i] = r[i]; }
    sha256_64B<ADDMODE>(w, out, A);
}
__device__ __forceinline__ void sha256_pair(const uint32_t (&l)[8], const uint32_t (&r)[8], uint32_t (&out)[8]) { sha256_pair<0>(l, r, out, ShaAdd<0>(1u)); }

// SHA-256 of a short message of `NBYTES` (multiple of 4, <= 52) given as words: one compression.
// Covers sha256(u256) (32 B), sha256_32 (4 B), trace / QM31 leaves (16 B), channel draws (36 B),
// channel_mix_u64 (40 B), channel_mix_line_poly (48 B), stark101 channel_mix_32 (36 B).
template <int NWORDS, int ADDMODE>
__device__ __forceinline__ void sha256_short(const uint32_t (&m)[NWORDS], uint32_t (&out)[8], const ShaAdd<ADDMODE> A) {
    static_assert(NWORDS >= 1 && NWORDS <= 13, "single-block messages only");
    uint32_t w[16];
#pragma unroll
    for (int i = 0; i < 16; i++) w[i] = 0;
#pragma unroll
    for (int i = 0; i < NWORDS; i++) w[i] = m[i];
    w[NWORDS] = 0x80000000u;
    w[15] = NWORDS * 32u;
    sha_iv(out);
    sha_compress<ADDMODE>(out, w, A);
}
template <int NWORDS>
__device__ __forceinline__ void sha256_short(const uint32_t (&m)[NWORDS], uint32_t (&out)[8]) { sha256_short<NWORDS, 0>(m, out, ShaAdd<0>(1u)); }

// SHA-256 of an arbitrary message of `nwords` 32-bit big-endian words, word i supplied by `get(i)`.
// One compression body in a rolled block loop: compact code for the transcript kernels, where
// latency, not ALU throughput, is what matters (channel.simf:36-172, deep/oods.simf:23-39).
template <class F>
__device__ __forceinline__ void sha256_msg(int nwords, F get, uint32_t (&out)[8]) {
    sha_iv(out);
    const int nblocks = (nwords + 3 + 15) >> 4; // + 0x80 word + 64-bit length
#pragma unroll 1
    for (int b = 0; b < nblocks; b++) {
        uint32_t w[16];
#pragma unroll
        for (int j = 0; j < 16; j++) {
            const int i = b * 16 + j;
            uint32_t v = 0;
            if (i < nwords) v = get(i);
            else if (i == nwords) v = 0x80000000u;
            else if (i == nblocks * 16 - 1) v = (uint32_t)nwords * 32u;
            w[j] = v;
        }
        sha_compress(out, w);
    }
}

} // namespace ssym
