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

#define SSYM_SHA_ROUND(a, b, c, d, e, f, g, h, kw)           \
    {                                                        \
        uint32_t t1_ = (h) + Sig1(e) + Ch(e, f, g) + (kw);   \
        uint32_t t2_ = Sig0(a) + Maj(a, b, c);               \
        (d) += t1_;                                          \
        (h) = t1_ + t2_;                                     \
    }

// One compression of `h` with the 16-word block `w` (w is consumed: it becomes the rolling schedule).
__device__ __forceinline__ void sha_compress(uint32_t (&h)[8], uint32_t (&w)[16]) {
    constexpr ShaK K = sha_k_table();
    uint32_t a = h[0], b = h[1], c = h[2], d = h[3], e = h[4], f = h[5], g = h[6], hh = h[7];
#pragma unroll
    for (int t = 0; t < 64; t += 8) {
        if (t >= 16) {
#pragma unroll
            for (int j = 0; j < 8; j++) {
                const int i = (t + j) & 15;
                w[i] = w[i] + sig0(w[(i + 1) & 15]) + w[(i + 9) & 15] + sig1(w[(i + 14) & 15]);
            }
        }
        SSYM_SHA_ROUND(a, b, c, d, e, f, g, hh, K.k[t + 0] + w[(t + 0) & 15]);
        SSYM_SHA_ROUND(hh, a, b, c, d, e, f, g, K.k[t + 1] + w[(t + 1) & 15]);
        SSYM_SHA_ROUND(g, hh, a, b, c, d, e, f, K.k[t + 2] + w[(t + 2) & 15]);
        SSYM_SHA_ROUND(f, g, hh, a, b, c, d, e, K.k[t + 3] + w[(t + 3) & 15]);
        SSYM_SHA_ROUND(e, f, g, hh, a, b, c, d, K.k[t + 4] + w[(t + 4) & 15]);
        SSYM_SHA_ROUND(d, e, f, g, hh, a, b, c, K.k[t + 5] + w[(t + 5) & 15]);
        SSYM_SHA_ROUND(c, d, e, f, g, hh, a, b, K.k[t + 6] + w[(t + 6) & 15]);
        SSYM_SHA_ROUND(b, c, d, e, f, g, hh, a, K.k[t + 7] + w[(t + 7) & 15]);
    }
    h[0] += a; h[1] += b; h[2] += c; h[3] += d; h[4] += e; h[5] += f; h[6] += g; h[7] += hh;
}

// Compression of the constant padding block that ends every 64-byte message.
__device__ __forceinline__ void sha_compress_pad64(uint32_t (&h)[8]) {
    constexpr ShaKW KW = sha_pad64_kw();
    uint32_t a = h[0], b = h[1], c = h[2], d = h[3], e = h[4], f = h[5], g = h[6], hh = h[7];
#pragma unroll
    for (int t = 0; t < 64; t += 8) {
        SSYM_SHA_ROUND(a, b, c, d, e, f, g, hh, KW.kw[t + 0]);
        SSYM_SHA_ROUND(hh, a, b, c, d, e, f, g, KW.kw[t + 1]);
        SSYM_SHA_ROUND(g, hh, a, b, c, d, e, f, KW.kw[t + 2]);
        SSYM_SHA_ROUND(f, g, hh, a, b, c, d, e, KW.kw[t + 3]);
        SSYM_SHA_ROUND(e, f, g, hh, a, b, c, d, KW.kw[t + 4]);
        SSYM_SHA_ROUND(d, e, f, g, hh, a, b, c, KW.kw[t + 5]);
        SSYM_SHA_ROUND(c, d, e, f, g, hh, a, b, KW.kw[t + 6]);
        SSYM_SHA_ROUND(b, c, d, e, f, g, hh, a, KW.kw[t + 7]);
    }
    h[0] += a; h[1] += b; h[2] += c; h[3] += d; h[4] += e; h[5] += f; h[6] += g; h[7] += hh;
}

// SHA-256 of a 64-byte message given as 16 big-endian words (2 compressions).  Used by
// sha256_pair (hasher.simf:27-32), channel_mix_u256 (channel.simf:154-162) and the CP leaf
// (hash_node_m31_cp, hasher.simf:93-97).  `w` is clobbered.
__device__ __forceinline__ void sha256_64B(uint32_t (&w)[16], uint32_t (&out)[8]) {
    sha_iv(out);
    sha_compress(out, w);
    sha_compress_pad64(out);
}

// sha256_pair(left, right)
__device__ __forceinline__ void sha256_pair(const uint32_t (&l)[8], const uint32_t (&r)[8], uint32_t (&out)[8]) {
    uint32_t w[16];
#pragma unroll
    for (int i = 0; i < 8; i++) { w[i] = l[i]; w[8 + i] = r[i]; }
    sha256_64B(w, out);
}

// SHA-256 of a short message of `NBYTES` (multiple of 4, <= 52) given as words: one compression.
// Covers sha256(u256) (32 B), sha256_32 (4 B), trace / QM31 leaves (16 B), channel draws (36 B),
// channel_mix_u64 (40 B), channel_mix_line_poly (48 B), stark101 channel_mix_32 (36 B).
template <int NWORDS>
__device__ __forceinline__ void sha256_short(const uint32_t (&m)[NWORDS], uint32_t (&out)[8]) {
    static_assert(NWORDS >= 1 && NWORDS <= 13, "single-block messages only");
    uint32_t w[16];
#pragma unroll
    for (int i = 0; i < 16; i++) w[i] = 0;
#pragma unroll
    for (int i = 0; i < NWORDS; i++) w[i] = m[i];
    w[NWORDS] = 0x80000000u;
    w[15] = NWORDS * 32u;
    sha_iv(out);
    sha_compress(out, w);
}

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
