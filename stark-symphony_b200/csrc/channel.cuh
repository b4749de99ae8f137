// SPDX-License-Identifier: MIT
//
// Device Fiat-Shamir channel of the stwo-verifier program (stwo-verifier/src/channel.simf:18-172):
// ChannelState = (u256 digest, u32 n_sent); shared by the verifier's transcript kernel (stwo_kernels.cu)
// and by the prover's transcript kernels (prover_kernels.cu), which must replay exactly the same chain.
#pragma once
#include "field.cuh"
#include "sha256.cuh"

namespace ssym {

// ------------------------------------------------------------------------------------------
// Transcript helpers (K1).  The compression function is deliberately NOT inlined here: a
// transcript is ~46 dependent compressions per proof, so one hot copy in the I-cache beats
// 46 cold ones.
// ------------------------------------------------------------------------------------------
static __device__ __noinline__ void sha_compress_call(uint32_t *h, const uint32_t *blk) {
    uint32_t hh[8], w[16];
#pragma unroll
    for (int i = 0; i < 8; i++) hh[i] = h[i];
#pragma unroll
    for (int i = 0; i < 16; i++) w[i] = blk[i];
    sha_compress_rolled<0>(hh, w, ShaAdd<0>(1u)); // 4 x 16 rounds: the body stays resident in the instruction cache of an SM that runs one warp of this
#pragma unroll
    for (int i = 0; i < 8; i++) h[i] = hh[i];
}

// SHA-256 of (a[0..na) || b[0..nb)) big-endian words.
static __device__ __noinline__ void sha256_2part(const uint32_t *a, int na, const uint32_t *b, int nb, uint32_t *out) {
    uint32_t h[8];
    sha_iv(h);
    const int nwords = na + nb;
    const int nblocks = (nwords + 3 + 15) >> 4;
    for (int blk = 0; blk < nblocks; blk++) {
        uint32_t w[16];
#pragma unroll
        for (int j = 0; j < 16; j++) {
            const int i = blk * 16 + j;
            uint32_t v = 0;
            if (i < na) v = a[i];
            else if (i < nwords) v = b[i - na];
            else if (i == nwords) v = 0x80000000u;
            else if (i == nblocks * 16 - 1) v = (uint32_t)nwords * 32u;
            w[j] = v;
        }
        sha_compress_call(h, w);
    }
#pragma unroll
    for (int i = 0; i < 8; i++) out[i] = h[i];
}

struct Channel { // channel.simf:18-22 ChannelState = (u256 digest, u32 n_sent)
    uint32_t d[8];
    uint32_t n_sent;
};
__device__ __forceinline__ void channel_draw_u256(Channel &c, uint32_t *out) { // channel.simf:36-44
    uint32_t ns = c.n_sent;
    sha256_2part(c.d, 8, &ns, 1, out);
    c.n_sent = c.n_sent + 1u;
}
__device__ __forceinline__ void channel_mix(Channel &c, const uint32_t *in, int nwords) { // channel.simf:154-173, fri/commit.simf:48-57, deep/oods.simf:23-39
    uint32_t out[8];
    sha256_2part(c.d, 8, in, nwords, out);
#pragma unroll
    for (int i = 0; i < 8; i++) c.d[i] = out[i];
    c.n_sent = 0;
}
// channel_draw_qm31 = channel_draw_m31x4 (channel.simf:115-141): retry (<= 256 draws) until the first four words are < 2p
static __device__ __noinline__ QM31 channel_draw_qm31(Channel &c, bool &exhausted) {
    uint32_t w[8];
    bool ok = false;
    for (int counter = 0; counter < 256 && !ok; counter++) {
        channel_draw_u256(c, w);
        ok = w[0] < 4294967294u && w[1] < 4294967294u && w[2] < 4294967294u && w[3] < 4294967294u;
    }
    exhausted = exhausted || !ok;
    return qm31(m31_reduce(w[0]), m31_reduce(w[1]), m31_reduce(w[2]), m31_reduce(w[3]));
}

} // namespace ssym
