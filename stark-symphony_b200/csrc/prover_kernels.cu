// SPDX-License-Identifier: MIT
// Kernels of the batched Stwo prover; see prover_kernels.cuh for the decomposition.
#include "prover_kernels.cuh"

#include "channel.cuh"
#include "deep.cuh"
#include "field.cuh"
#include "sha256.cuh"

namespace ssym {

typedef PrvCtx PC;
typedef ShaAdd<8> ShaA; // adds on the FMA pipe, as in the verifier's Merkle kernel (sha256.cuh)

// ------------------------------------------------------------------------------------------
// small helpers
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint64_t prv_splitmix(uint64_t seed, uint64_t row) { // = pr_splitmix, oracle/stwo_prover_ref.c
    uint64_t z = seed * 0x9E3779B97F4A7C15ull + (row + 1) * 0xBF58476D1CE4E5B9ull;
    z ^= z >> 30; z *= 0xBF58476D1CE4E5B9ull;
    z ^= z >> 27; z *= 0x94D049BB133111EBull;
    z ^= z >> 31;
    return z;
}
__device__ __forceinline__ void store_digest(uint32_t *dst, const uint32_t (&d)[8]) { // 32-byte aligned
    reinterpret_cast<uint4 *>(dst)[0] = make_uint4(d[0], d[1], d[2], d[3]);
    reinterpret_cast<uint4 *>(dst)[1] = make_uint4(d[4], d[5], d[6], d[7]);
}
__device__ __forceinline__ void copy_digest(uint32_t *dst, const uint32_t *src) {
    reinterpret_cast<uint4 *>(dst)[0] = reinterpret_cast<const uint4 *>(src)[0];
    reinterpret_cast<uint4 *>(dst)[1] = reinterpret_cast<const uint4 *>(src)[1];
}
// SHA-256 of a 16-byte message (trace leaf hasher.simf:85-90, QM31 leaf hasher.simf:100-104): one compression
__device__ __forceinline__ void hash_16B(uint4 v, uint32_t (&out)[8], const ShaA A) {
    uint32_t w[16];
    w[0] = v.x; w[1] = v.y; w[2] = v.z; w[3] = v.w;
    w[4] = 0x80000000u;
#pragma unroll
    for (int k = 5; k < 15; k++) w[k] = 0;
    w[15] = 128u;
    sha_iv(out);
    sha_compress_rolled<8>(out, w, A);
}

// Block-wide circle FFT over `ncols` columns in shared memory (column c at v + c * col_stride), 2^n points each, in the basis
// b_j = y^j0 x^j1 pi(x)^j2 ... (j0 = least significant bit): evaluations in bit-reversed slots, coefficients in natural order.
// INV: evaluations -> coefficients (without the 2^-n scale); else coefficients -> evaluations.  `tw` = the domain's tw / itw table.
template <bool INV>
__device__ void cfft_block(uint32_t *v, uint32_t ncols, uint32_t col_stride, uint32_t n, const uint32_t *__restrict__ tw) {
    const uint32_t half = 1u << (n - 1);
    for (uint32_t step = 0; step < n; step++) {
        const uint32_t l = INV ? step : n - 1 - step;
        const uint32_t stride = 1u << l;
        const uint32_t *twl = tw + ((1u << n) - (1u << (n - l)));
        for (uint32_t idx = threadIdx.x; idx < ncols * half; idx += blockDim.x) {
            const uint32_t col = idx >> (n - 1), t = idx & (half - 1);
            const uint32_t s = ((t >> l) << (l + 1)) | (t & (stride - 1));
            const uint32_t w = __ldg(twl + (t >> l));
            uint32_t *a = v + col * col_stride + s, *b = a + stride;
            const uint32_t x = *a, y = *b;
            if (INV) {
                *a = m31_add(x, y);
                *b = m31_mul(m31_sub(x, y), w);
            } else {
                const uint32_t ty = m31_mul(y, w);
                *a = m31_add(x, ty);
                *b = m31_sub(x, ty);
            }
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------
// tables
// ------------------------------------------------------------------------------------------
__global__ void prv_tables_kernel(uint32_t n, uint32_t *tw, uint32_t *itw) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x; // entry index in the concatenated table
    if (t >= (1u << n) - 1u) return;
    uint32_t l = 0, j = t;
    while (j >= (1u << (n - l - 1))) { j -= 1u << (n - l - 1); l++; }
    const uint32_t log = n - l;
    M31 v;
    if (l == 0) v = circle_point_index_to_m31_point(circle_position_to_point_index(log, bit_reverse_position(2 * j, log))).y;
    else v = circle_point_index_to_m31_point(line_position_to_point_index(log, bit_reverse_position(2 * j, log))).x;
    bool fail = false;
    tw[t] = v;
    itw[t] = m31_inv(v, fail);
}
__global__ void prv_vanish_kernel(uint32_t trace_log, uint32_t lde_log, const uint2 *point, uint32_t *vanish_inv) {
    const uint32_t q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= (1u << lde_log)) return;
    M31 x = point[q].x;
    for (uint32_t k = 1; k < trace_log; k++) x = m31_point_dbl_x(x); // pi_fn, composition_poly.simf:26-35
    bool fail = false;
    vanish_inv[q] = m31_inv(x, fail);
}
void launch_prv_tables(uint32_t n, uint32_t *tw, uint32_t *itw, cudaStream_t s) {
    prv_tables_kernel<<<((1u << n) + 127) / 128, 128, 0, s>>>(n, tw, itw);
}
void launch_prv_vanish(uint32_t trace_log, uint32_t lde_log, const uint2 *point, uint32_t *vanish_inv, cudaStream_t s) {
    prv_vanish_kernel<<<((1u << lde_log) + 127) / 128, 128, 0, s>>>(trace_log, lde_log, point, vanish_inv);
}

// ------------------------------------------------------------------------------------------
// P1: trace -> coefficients -> LDE
// ------------------------------------------------------------------------------------------
// One CTA per (proof, group of four columns): the shared-memory FFTs hold 4 columns x 2^G words; a group re-runs the row recurrence
// from c0, c1 up to its own columns (at most 14 squarings per row).
__global__ void __launch_bounds__(512) prv_trace_kernel(PrvParams p) {
    extern __shared__ uint32_t sm[];
    const uint32_t T = p.cfg.trace_log, G = p.cfg.lde_log, NT = 1u << T, NG = 1u << G, C = SSYM_STWO_COLUMNS(&p.cfg);
    const uint32_t i = blockIdx.x, c_first = 4 * blockIdx.y;
    const uint64_t seed = p.seeds[i];
    for (uint32_t r = threadIdx.x; r < NT; r += blockDim.x) { // row: c_k = c_{k-1}^2 + c_{k-2}^2 (wide_fibonacci.simf:24-62)
        const uint32_t raw = (uint32_t)(prv_splitmix(seed, r) >> 33);
        uint32_t a = 1u, b = raw == SSYM_P ? 0u : raw; // c0, c1
        for (uint32_t k = 0; k < c_first + 4; k++) {    // a = c_k
            if (k >= c_first) sm[(k - c_first) * NG + r] = a;
            const uint32_t nx = m31_add(m31_mul(b, b), m31_mul(a, a));
            a = b;
            b = nx;
        }
    }
    __syncthreads();
    cfft_block<true>(sm, 4, NG, T, p.tr.itw);
    uint32_t *tc = p.tcoef + ((size_t)i * C + c_first) * NT;
    for (uint32_t idx = threadIdx.x; idx < 4 * NG; idx += blockDim.x) {
        const uint32_t c = idx >> G, j = idx & (NG - 1);
        uint32_t v = 0;
        if (j < NT) {
            v = m31_mul(sm[c * NG + j], p.tr.scale);
            tc[c * NT + j] = v;
        }
        sm[c * NG + j] = v; // zero-extended coefficient vector
    }
    __syncthreads();
    cfft_block<false>(sm, 4, NG, G, p.lde.tw);
    uint32_t *tl = p.tlde + ((size_t)i * C + c_first) * NG;
    for (uint32_t idx = threadIdx.x; idx < 4 * NG; idx += blockDim.x) tl[idx] = sm[idx];
}

// ------------------------------------------------------------------------------------------
// P2: Merkle leaves and levels
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) prv_leaf_trace_kernel(PrvParams p, uint32_t one) {
    const ShaA A{one};
    const uint32_t G = p.cfg.lde_log, NG = 1u << G;
    const uint32_t gid = blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= p.m << G) return;
    const uint32_t i = gid >> G, q = gid & (NG - 1), C = SSYM_STWO_COLUMNS(&p.cfg);
    const uint32_t *tl = p.tlde + (size_t)i * C * NG + q;
    uint32_t d[8];
    if (C == 4) {
        hash_16B(make_uint4(tl[0], tl[NG], tl[2 * NG], tl[3 * NG]), d, A);
    } else { // hash_node_m31_trace (hasher.simf:85-90) over C words: 8 -> one block with the padding behind the data, 16 -> a 64-byte message
        uint32_t w[16];
#pragma unroll
        for (int k = 0; k < 16; k++) w[k] = (uint32_t)k < C ? tl[(size_t)k * NG] : 0u;
        if (C == 16) {
            sha256_64B_rolled<8>(w, d, A);
        } else {
            w[8] = 0x80000000u;
            w[15] = 256u;
            sha_iv(d);
            sha_compress_rolled<8>(d, w, A);
        }
    }
    store_digest(p.tree_t + ((size_t)i * 2 * NG + NG + q) * 8, d);
}
__global__ void __launch_bounds__(128) prv_leaf_cp_kernel(PrvParams p, uint32_t one) { // hash_node_m31_cp hasher.simf:93-97
    const ShaA A{one};
    const uint32_t G = p.cfg.lde_log, NG = 1u << G;
    const uint32_t gid = blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= p.m << G) return;
    const uint32_t i = gid >> G, q = gid & (NG - 1);
    const uint32_t *cl = p.cplde + (size_t)i * 16 * NG + q;
    uint32_t w[16], d[8];
#pragma unroll
    for (int k = 0; k < 16; k++) w[k] = cl[(size_t)k * NG];
    sha256_64B_rolled<8>(w, d, A);
    store_digest(p.tree_c + ((size_t)i * 2 * NG + NG + q) * 8, d);
}
__global__ void __launch_bounds__(128) prv_leaf_fri_kernel(PrvParams p, uint32_t layer, uint32_t one) {
    const ShaA A{one};
    const uint32_t n = p.cfg.lde_log - layer, N = 1u << n;
    const uint32_t gid = blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= p.m << n) return;
    const uint32_t i = gid >> n, q = gid & (N - 1);
    const uint4 v = *reinterpret_cast<const uint4 *>(p.fev + (size_t)i * p.fev_stride + p.fev_off[layer] + 4 * q);
    uint32_t d[8];
    hash_16B(v, d, A);
    store_digest(p.ftree + (size_t)i * p.ftree_stride + p.ftree_off[layer] + (size_t)(N + q) * 8, d);
}
// nodes [2^k, 2^(k+1)) of every proof's tree: node = sha256_pair(left child, right child)  (merkle.simf:22-30, hasher.simf:27-32)
__global__ void __launch_bounds__(128) prv_tree_level_kernel(uint32_t *base, size_t stride_words, uint32_t off_words, uint32_t k, uint32_t m, uint32_t one) {
    const ShaA A{one};
    const uint32_t gid = blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= m << k) return;
    const uint32_t i = gid >> k, node = (1u << k) + (gid & ((1u << k) - 1));
    uint32_t *tree = base + (size_t)i * stride_words + off_words;
    const uint4 *ch = reinterpret_cast<const uint4 *>(tree + (size_t)node * 16);
    uint32_t w[16], d[8];
#pragma unroll
    for (int j = 0; j < 4; j++) {
        const uint4 v = ch[j];
        w[4 * j] = v.x; w[4 * j + 1] = v.y; w[4 * j + 2] = v.z; w[4 * j + 3] = v.w;
    }
    sha256_64B_rolled<8>(w, d, A);
    store_digest(tree + (size_t)node * 8, d);
}

// ------------------------------------------------------------------------------------------
// channel kernels (thread per proof)
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void ch_load(Channel &c, const uint32_t *pc) {
#pragma unroll
    for (int k = 0; k < 8; k++) c.d[k] = pc[PC::CH + k];
    c.n_sent = pc[PC::CH + 8];
}
__device__ __forceinline__ void ch_store(const Channel &c, uint32_t *pc) {
#pragma unroll
    for (int k = 0; k < 8; k++) pc[PC::CH + k] = c.d[k];
    pc[PC::CH + 8] = c.n_sent;
}
__device__ __noinline__ QM31 qm31_mul_pn(QM31 x, QM31 y) { return qm31_mul(x, y); }

// evals/commit.simf:20-35 up to the composition-polynomial coefficient
__global__ void __launch_bounds__(64) prv_ch_commit_kernel(PrvParams p) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= p.m) return;
    const uint32_t NG = 1u << p.cfg.lde_log;
    uint32_t *pc = p.pctx + (size_t)i * PC::WORDS, *out = p.out + (size_t)i * p.lo.stride_words;
    // the constant-column tree is never decommitted by the verifier; its root is SHA-256("") in the reference fixtures
    const uint32_t empty[8] = {0xe3b0c442u, 0x98fc1c14u, 0x9afbf4c8u, 0x996fb924u, 0x27ae41e4u, 0x649b934cu, 0xa495991bu, 0x7852b855u};
    for (int k = 0; k < 8; k++) {
        out[p.lo.off_commit + k] = empty[k];
        out[p.lo.off_commit + 8 + k] = p.tree_t[((size_t)i * 2 * NG + 1) * 8 + k];
    }
    Channel ch;
    for (int k = 0; k < 8; k++) ch.d[k] = 0;
    ch.n_sent = 0;
    bool ex = false;
    channel_mix(ch, out + p.lo.off_commit, 8);
    channel_mix(ch, out + p.lo.off_commit + 8, 8);
    qm31_store4(pc + PC::CP_ALPHA, channel_draw_qm31(ch, ex));
    ch_store(ch, pc);
    if (ex) atomicOr(p.flag, 2u);
}
// mix the composition root; draw the OODS point (channel.simf:143-151); basis factors at the point
__global__ void __launch_bounds__(64) prv_ch_oods_point_kernel(PrvParams p) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= p.m) return;
    const uint32_t NG = 1u << p.cfg.lde_log, T = p.cfg.trace_log;
    uint32_t *pc = p.pctx + (size_t)i * PC::WORDS, *out = p.out + (size_t)i * p.lo.stride_words;
    for (int k = 0; k < 8; k++) out[p.lo.off_commit + 16 + k] = p.tree_c[((size_t)i * 2 * NG + 1) * 8 + k];
    Channel ch;
    ch_load(ch, pc);
    bool ex = false, iz = false;
    channel_mix(ch, out + p.lo.off_commit + 16, 8);
    const QM31 t = channel_draw_qm31(ch, ex);
    const QM31 t_sq = qm31_mul_pn(t, t);
    const QM31 inv = qm31_inv(qm31_add(qm31_one(), t_sq), iz);
    const QM31 px = qm31_mul_pn(qm31_sub(qm31_one(), t_sq), inv), py = qm31_mul_pn(qm31_add(t, t), inv);
    const QM31 pxy = qm31_mul_pn(px, py);
    qm31_store4(pc + PC::PX, px);
    qm31_store4(pc + PC::PY, py);
    qm31_store4(pc + PC::P2Y, qm31_add(pxy, pxy));
    qm31_store4(pc + PC::TW, py);
    qm31_store4(pc + PC::TW + 4, px);
    QM31 v = px;
    for (uint32_t k = 2; k <= T; k++) {
        const QM31 sq = qm31_mul_pn(v, v);
        v = qm31_sub(qm31_add(sq, sq), qm31_one()); // qm31_point_dbl_x
        qm31_store4(pc + PC::TW + 4 * k, v);
        if (k == 2) qm31_store4(pc + PC::P2X, v);
    }
    ch_store(ch, pc);
    if (ex || iz) atomicOr(p.flag, 2u);
}
// mix the samples (deep/oods.simf:23-39), draw the DEEP coefficient, line coefficients of the C + 16 columns (deep/quotients.simf:25-35)
__global__ void __launch_bounds__(64) prv_ch_deep_kernel(PrvParams p) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= p.m) return;
    uint32_t *pc = p.pctx + (size_t)i * PC::WORDS;
    const uint32_t *out = p.out + (size_t)i * p.lo.stride_words;
    Channel ch;
    ch_load(ch, pc);
    bool ex = false;
    const uint32_t C = SSYM_STWO_COLUMNS(&p.cfg);
    channel_mix(ch, out + p.lo.off_oods_trace, 4 * (C + 16));
    const QM31 deep_alpha = channel_draw_qm31(ch, ex);
    qm31_store4(pc + PC::DEEP_ALPHA, deep_alpha);
    const QM31 py = qm31_load4(pc + PC::PY), p2y = qm31_load4(pc + PC::P2Y);
    QM31 alpha_i = deep_alpha;
    QM31 sa_a = qm31_zero(), sa_c = qm31_zero(), sb_a = qm31_zero(), sb_c = qm31_zero();
#pragma unroll 1
    for (uint32_t k = 0; k < C + 16; k++) { // aggregation order of Appendix A item 1: 16 CP columns at 2P, then the C trace columns at P
        const bool in_a = k < 16;
        const QM31 sv = qm31_load4(in_a ? out + p.lo.off_oods_cp + 4 * k : out + p.lo.off_oods_trace + 4 * (k - 16));
        const LineCoeffs lc = interpolant_coefficients(in_a ? p2y : py, sv, alpha_i);
        qm31_store4(pc + PC::KB + 4 * k, lc.b);
        if (in_a) { sa_a = qm31_add(sa_a, lc.a); sa_c = qm31_add(sa_c, lc.c); }
        else { sb_a = qm31_add(sb_a, lc.a); sb_c = qm31_add(sb_c, lc.c); }
        alpha_i = qm31_mul_pn(alpha_i, deep_alpha);
    }
    qm31_store4(pc + PC::SUMS, sa_a);
    qm31_store4(pc + PC::SUMS + 4, sa_c);
    qm31_store4(pc + PC::SUMS + 8, sb_a);
    qm31_store4(pc + PC::SUMS + 12, sb_c);
    ch_store(ch, pc);
    if (ex) atomicOr(p.flag, 2u);
}
// fri_layer_commit fri/commit.simf:36-45: mix the layer root, draw the folding coefficient
__global__ void __launch_bounds__(64) prv_ch_fri_kernel(PrvParams p, uint32_t layer) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= p.m) return;
    uint32_t *pc = p.pctx + (size_t)i * PC::WORDS, *out = p.out + (size_t)i * p.lo.stride_words;
    uint32_t *root = out + (layer == 0 ? p.lo.off_fri_first_root : p.lo.off_fri_inner_root + 8 * (layer - 1));
    const uint32_t *tr = p.ftree + (size_t)i * p.ftree_stride + p.ftree_off[layer] + 8;
    for (int k = 0; k < 8; k++) root[k] = tr[k];
    Channel ch;
    ch_load(ch, pc);
    bool ex = false;
    channel_mix(ch, root, 8);
    qm31_store4(pc + PC::FRI_ALPHA + 4 * layer, channel_draw_qm31(ch, ex));
    ch_store(ch, pc);
    if (ex) atomicOr(p.flag, 2u);
}

// ------------------------------------------------------------------------------------------
// P3: composition polynomial columns
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(512) prv_cp_kernel(PrvParams p) {
    extern __shared__ uint32_t sm[]; // [0, NG): evaluations / coefficients of this coordinate; [NG, 5 NG): the four sub-polynomial columns
    const uint32_t T = p.cfg.trace_log, G = p.cfg.lde_log, NT = 1u << T, NG = 1u << G;
    const uint32_t i = blockIdx.x, coord = blockIdx.y, C = SSYM_STWO_COLUMNS(&p.cfg);
    __shared__ uint32_t s_al[SSYM_MAX_COLUMNS]; // s_al[j] = this coordinate of cp_alpha^j: constraint k carries alpha^(C-1-k) in the Horner fold
    if (threadIdx.x == 0) {
        const QM31 alpha = qm31_load4(p.pctx + (size_t)i * PC::WORDS + PC::CP_ALPHA);
        QM31 pw = qm31_one();
        for (uint32_t j = 0; j + 2 < C; j++) {
            s_al[j] = coord == 0 ? pw.r.a : coord == 1 ? pw.r.b : coord == 2 ? pw.i.a : pw.i.b;
            pw = qm31_mul_pn(pw, alpha);
        }
    }
    __syncthreads();
    const uint32_t *tl = p.tlde + (size_t)i * C * NG;
    for (uint32_t q = threadIdx.x; q < NG; q += blockDim.x) {
        // (sum_k alpha^(C-1-k) (c_k - c_{k-1}^2 - c_{k-2}^2)) / vanishing, this QM31 coordinate (wide_fibonacci.simf:24-62)
        uint32_t a = tl[q], b = tl[NG + q], as = m31_mul(a, a), bs = m31_mul(b, b), acc = 0;
        for (uint32_t k = 2; k < C; k++) {
            const uint32_t ck = tl[(size_t)k * NG + q];
            acc = m31_add(acc, m31_mul(s_al[C - 1 - k], m31_sub(ck, m31_add(bs, as))));
            as = bs;
            bs = m31_mul(ck, ck);
        }
        sm[q] = m31_mul(acc, __ldg(p.vanish_inv + q));
    }
    __syncthreads();
    cfft_block<true>(sm, 1, NG, G, p.lde.itw);
    bool bad = false;
    uint32_t *cc = p.cpcoef + ((size_t)i * 4 + coord) * 2 * NT;
    for (uint32_t j = threadIdx.x; j < NG; j += blockDim.x) {
        const uint32_t v = m31_mul(sm[j], p.lde.scale);
        sm[j] = v;
        if (j < 2 * NT) cc[j] = v;
        if (j > NT && v) bad = true; // CP has total degree <= 2^(T-1): coefficients live in [0, 2^T]
    }
    if (bad) atomicOr(p.flag, 1u);
    for (uint32_t idx = threadIdx.x; idx < 4 * NG; idx += blockDim.x) sm[NG + idx] = 0;
    __syncthreads();
    for (uint32_t j = threadIdx.x; j <= NT; j += blockDim.x) // sub-polynomial (j & 3) in the (X, Y) = 2P basis: no Y, X-bits shifted down
        sm[NG + (j & 3) * NG + 2 * (j >> 2)] = sm[j];
    __syncthreads();
    cfft_block<false>(sm + NG, 4, NG, G, p.lde.tw);
    uint32_t *cl = p.cplde + ((size_t)i * 16 + 4 * coord) * NG;
    for (uint32_t idx = threadIdx.x; idx < 4 * NG; idx += blockDim.x) cl[idx] = sm[NG + idx];
}

// ------------------------------------------------------------------------------------------
// P4: samples at the OODS point.  One warp per (proof, column): sum_m c[m] * prod_k t_k^(bit k of m).
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ QM31 qm31_shfl_down(QM31 v, uint32_t d) {
    return qm31(__shfl_down_sync(0xffffffffu, v.r.a, d), __shfl_down_sync(0xffffffffu, v.r.b, d), __shfl_down_sync(0xffffffffu, v.i.a, d),
                __shfl_down_sync(0xffffffffu, v.i.b, d));
}
#define PRV_OODS_WARPS 4
__global__ void __launch_bounds__(32 * PRV_OODS_WARPS) prv_oods_kernel(PrvParams p) {
    __shared__ __align__(16) uint32_t s_basis[PRV_OODS_WARPS][16 * 4];
    const uint32_t T = p.cfg.trace_log, NT = 1u << T;
    const uint32_t wib = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t wid = blockIdx.x * PRV_OODS_WARPS + wib;
    const uint32_t C = SSYM_STWO_COLUMNS(&p.cfg), NCOL = C + 16;
    if (wid >= p.m * NCOL) return; // warp-uniform
    const uint32_t i = wid / NCOL, col = wid % NCOL;
    const uint32_t *pc = p.pctx + (size_t)i * PC::WORDS;
    const uint32_t *c, *tw;
    uint32_t stride, bits;
    if (col < C) { c = p.tcoef + ((size_t)i * C + col) * NT; stride = 1; bits = T; tw = pc + PC::TW; }
    else { // CP column k = 4 * coord + poly: coefficients c[4m + poly] of the coordinate, factors pi^k(2P.x) = TW[2 + k]
        const uint32_t k = col - C;
        c = p.cpcoef + ((size_t)i * 4 + (k >> 2)) * 2 * NT + (k & 3); stride = 4; bits = T - 1; tw = pc + PC::TW + 8;
    }
    const uint32_t blk_bits = bits > 5 ? bits - 5 : 0, lane_bits = bits - blk_bits;
    { // basis products of the low blk_bits bits (<= 16 entries)
        QM31 b = qm31_one();
        for (uint32_t k = 0; k < blk_bits; k++)
            if ((lane >> k) & 1u) b = qm31_mul_pn(b, qm31_load4(tw + 4 * k));
        if (lane < (1u << blk_bits)) qm31_store4(&s_basis[wib][4 * lane], b);
    }
    __syncwarp();
    QM31 v = qm31_zero();
    if (lane < (1u << lane_bits)) {
        const uint32_t base = lane << blk_bits;
        for (uint32_t m = 0; m < (1u << blk_bits); m++)
            v = qm31_add(v, qm31_mul_m31(qm31_load4(&s_basis[wib][4 * m]), c[(size_t)(base + m) * stride]));
    }
    for (uint32_t k = 0; k < lane_bits; k++) {
        const QM31 o = qm31_shfl_down(v, 1u << k);
        v = qm31_add(v, qm31_mul_pn(qm31_load4(tw + 4 * (blk_bits + k)), o));
    }
    if (lane == 0) {
        uint32_t *out = p.out + (size_t)i * p.lo.stride_words;
        qm31_store4(col < C ? out + p.lo.off_oods_trace + 4 * col : out + p.lo.off_oods_cp + 4 * (col - C), v);
    }
}

// ------------------------------------------------------------------------------------------
// P5: DEEP quotient on the whole LDE domain
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) prv_quotient_kernel(PrvParams p) {
    const uint32_t G = p.cfg.lde_log, NG = 1u << G;
    const uint32_t gid = blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= p.m << G) return;
    const uint32_t i = gid >> G, q = gid & (NG - 1);
    const uint32_t *pc = p.pctx + (size_t)i * PC::WORDS;
    const uint2 rp = __ldg(p.point + q);
    const M31Point R = m31_point(rp.x, rp.y);
    const uint32_t C = SSYM_STWO_COLUMNS(&p.cfg);
    const uint32_t *cl = p.cplde + (size_t)i * 16 * NG + q, *tl = p.tlde + (size_t)i * C * NG + q;
    QM31 na = qm31_zero(), nb = qm31_zero();
#pragma unroll 4
    for (int k = 0; k < 16; k++) na = qm31_add(na, qm31_mul_m31(qm31_load4(pc + PC::KB + 4 * k), cl[(size_t)k * NG]));
#pragma unroll 4
    for (uint32_t k = 0; k < C; k++) nb = qm31_add(nb, qm31_mul_m31(qm31_load4(pc + PC::KB + 4 * (16 + k)), tl[(size_t)k * NG]));
    na = qm31_sub(na, qm31_add(qm31_mul_m31(qm31_load4(pc + PC::SUMS), R.y), qm31_load4(pc + PC::SUMS + 4)));
    nb = qm31_sub(nb, qm31_add(qm31_mul_m31(qm31_load4(pc + PC::SUMS + 8), R.y), qm31_load4(pc + PC::SUMS + 12)));
    bool iz = false;
    const CM31 den_a = denominator_inverse(qm31_load4(pc + PC::P2X), qm31_load4(pc + PC::P2Y), R, iz);
    const CM31 den_b = denominator_inverse(qm31_load4(pc + PC::PX), qm31_load4(pc + PC::PY), R, iz);
    const QM31 h = qm31_add(qm31_mul_cm31(na, den_a), qm31_mul_cm31(nb, den_b));
    qm31_store4(p.fev + (size_t)i * p.fev_stride + p.fev_off[0] + 4 * q, h);
    if (iz) atomicOr(p.flag, 2u);
}

// ------------------------------------------------------------------------------------------
// P6: circle_fold (layer 0) / line_fold (fri/folding.simf:15-41) of a whole layer
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) prv_fold_kernel(PrvParams p, uint32_t layer) {
    const uint32_t G = p.cfg.lde_log, n = G - layer, half = 1u << (n - 1);
    const uint32_t gid = blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= p.m << (n - 1)) return;
    const uint32_t i = gid >> (n - 1), j = gid & (half - 1);
    const uint32_t *src = p.fev + (size_t)i * p.fev_stride + p.fev_off[layer] + 8 * j;
    const QM31 e0 = qm31_load4(src), e1 = qm31_load4(src + 4);
    const uint32_t inv = __ldg(p.lde.itw + ((1u << G) - (1u << n)) + j);
    const QM31 alpha = qm31_load4(p.pctx + (size_t)i * PC::WORDS + PC::FRI_ALPHA + 4 * layer);
    qm31_store4(p.fev + (size_t)i * p.fev_stride + p.fev_off[layer + 1] + 4 * j, qm31_fold(e0, e1, inv, alpha));
}

// ------------------------------------------------------------------------------------------
// P7: last layer, proof of work, queries
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(64) prv_final_kernel(PrvParams p) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= p.m) return;
    const uint32_t G = p.cfg.lde_log, L = p.cfg.n_fri_layers, Q = p.cfg.n_queries;
    uint32_t *pc = p.pctx + (size_t)i * PC::WORDS, *out = p.out + (size_t)i * p.lo.stride_words;
    const uint32_t *last = p.fev + (size_t)i * p.fev_stride + p.fev_off[L + 1];
    const QM31 coeff = qm31_load4(last);
    bool bad = false;
    for (uint32_t j = 1; j < (1u << (G - 1 - L)); j++) bad = bad || !qm31_eq(qm31_load4(last + 4 * j), coeff); // fri/layers.simf:73-78
    if (bad) atomicOr(p.flag, 1u);
    qm31_store4(out + p.lo.off_last_coeff, coeff);
    Channel ch;
    ch_load(ch, pc);
    channel_mix(ch, out + p.lo.off_last_coeff, 4); // channel_mix_line_poly fri/commit.simf:48-57
    uint64_t nonce = 0;
    for (;; nonce++) { // smallest nonce passing check_proof_of_work pow.simf:22-35
        Channel c2 = ch;
        const uint32_t nn[2] = {(uint32_t)(nonce >> 32), (uint32_t)nonce};
        channel_mix(c2, nn, 2);
        const uint64_t value = ((uint64_t)__byte_perm(c2.d[7], 0, 0x0123) << 32) | __byte_perm(c2.d[6], 0, 0x0123);
        if (value < p.cfg.pow_target || nonce == 0xffffffffull) { ch = c2; break; }
    }
    out[p.lo.off_pow_nonce] = (uint32_t)(nonce >> 32);
    out[p.lo.off_pow_nonce + 1] = (uint32_t)nonce;
    const uint32_t mask = (1u << G) - 1u;
    for (uint32_t q0 = 0; q0 < Q; q0 += 8) { // fri_generate_queries fri/queries.simf:30-43
        uint32_t w[8];
        channel_draw_u256(ch, w);
        for (uint32_t j = 0; j < 8 && q0 + j < Q; j++) pc[PC::QUERIES + q0 + j] = w[j] & mask;
    }
    uint32_t used = Q;
    if (p.cfg.mode & SSYM_MODE_QUERY_DEDUP) { // include/ssym.h: sorted distinct queries in slots [0, U), the other slots zero-filled
        uint32_t *qs = pc + PC::QUERIES;
        for (uint32_t a = 1; a < Q; a++) {
            const uint32_t v = qs[a];
            uint32_t b = a;
            for (; b > 0 && qs[b - 1] > v; b--) qs[b] = qs[b - 1];
            qs[b] = v;
        }
        used = 0;
        for (uint32_t a = 0; a < Q; a++) {
            const uint32_t v = qs[a];
            if (a == 0 || v != qs[used - 1]) qs[used++] = v;
        }
    }
    pc[PC::N_USED] = used;
    ch_store(ch, pc);
}

// ------------------------------------------------------------------------------------------
// P8: decommitments (evals/verify.simf:50-68, fri/layers.simf:40-69): one warp per (proof, query)
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) prv_decommit_kernel(PrvParams p) {
    const uint32_t G = p.cfg.lde_log, L = p.cfg.n_fri_layers, Q = p.cfg.n_queries, NG = 1u << G;
    const uint32_t wid = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (wid >= p.m * Q) return;
    const uint32_t i = wid / Q, qi = wid % Q;
    const uint32_t q = p.pctx[(size_t)i * PC::WORDS + PC::QUERIES + qi];
    uint32_t *out = p.out + (size_t)i * p.lo.stride_words;
    const uint32_t C = SSYM_STWO_COLUMNS(&p.cfg);
    if (qi >= p.pctx[(size_t)i * PC::WORDS + PC::N_USED]) { // an unused slot (SSYM_MODE_QUERY_DEDUP): zeros, as the reference prover writes
        if (lane < C + 16) out[p.lo.off_qvals + (C + 16) * qi + lane] = 0;
        if (lane <= L) *reinterpret_cast<uint4 *>(out + p.lo.off_fri_wit + (lane * Q + qi) * 4) = make_uint4(0, 0, 0, 0);
        const uint32_t n_slots = 2 * G + (L + 1) * (G - 1) - L * (L + 1) / 2;
        for (uint32_t job = lane; job < n_slots; job += 32) {
            uint32_t *dst;
            if (job < 2 * G) {
                dst = out + (job < G ? p.lo.off_trace_sib : p.lo.off_cp_sib) + (qi * G + (job < G ? job : job - G)) * 8;
            } else {
                uint32_t r = job - 2 * G, l = 0;
                while (r >= G - 1 - l) { r -= G - 1 - l; l++; }
                dst = out + p.lo.off_fri_sib[l] + (qi * (G - l - 1) + r) * 8;
            }
            reinterpret_cast<uint4 *>(dst)[0] = make_uint4(0, 0, 0, 0);
            reinterpret_cast<uint4 *>(dst)[1] = make_uint4(0, 0, 0, 0);
        }
        return;
    }
    if (lane < C + 16)
        out[p.lo.off_qvals + (C + 16) * qi + lane] = lane < C ? p.tlde[((size_t)i * C + lane) * NG + q] : p.cplde[((size_t)i * 16 + (lane - C)) * NG + q];
    if (lane <= L) { // the sibling evaluation of layer `lane` (adjacent_leaves fri/layers.simf:29-37)
        const uint32_t fq = q >> lane;
        *reinterpret_cast<uint4 *>(out + p.lo.off_fri_wit + (lane * Q + qi) * 4) =
            *reinterpret_cast<const uint4 *>(p.fev + (size_t)i * p.fev_stride + p.fev_off[lane] + 4 * (fq ^ 1u));
    }
    const uint32_t n_fri = (L + 1) * (G - 1) - L * (L + 1) / 2;
    for (uint32_t job = lane; job < 2 * G + n_fri; job += 32) {
        const uint32_t *src;
        uint32_t *dst;
        if (job < 2 * G) {
            const uint32_t lev = job < G ? job : job - G;
            const uint32_t sib = ((NG + q) >> lev) ^ 1u;
            src = (job < G ? p.tree_t : p.tree_c) + ((size_t)i * 2 * NG + sib) * 8;
            dst = out + (job < G ? p.lo.off_trace_sib : p.lo.off_cp_sib) + (qi * G + lev) * 8;
        } else {
            uint32_t r = job - 2 * G, l = 0;
            while (r >= G - 1 - l) { r -= G - 1 - l; l++; }
            const uint32_t n = G - l, fq = q >> l;
            const uint32_t sib = ((((1u << n) + fq) >> 1) >> r) ^ 1u;
            src = p.ftree + (size_t)i * p.ftree_stride + p.ftree_off[l] + (size_t)sib * 8;
            dst = out + p.lo.off_fri_sib[l] + (qi * (n - 1) + r) * 8;
        }
        copy_digest(dst, src);
    }
}

// ------------------------------------------------------------------------------------------
// host-side sequencing
// ------------------------------------------------------------------------------------------
int launch_prv_prove(const PrvParams &p, cudaStream_t s, uint64_t *launch_counter) {
    if (p.m == 0) return 0;
    const uint32_t T = p.cfg.trace_log, G = p.cfg.lde_log, L = p.cfg.n_fri_layers, Q = p.cfg.n_queries, NG = 1u << G, C = SSYM_STWO_COLUMNS(&p.cfg);
    (void)T;
    int launches = 0;
    cudaFuncSetAttribute(prv_trace_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 * (4 << SSYM_PRV_MAX_LOG));
    cudaFuncSetAttribute(prv_cp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 5 * (4 << SSYM_PRV_MAX_LOG));
    const uint32_t one = 1u; // opaque to ptxas: see sha256.cuh ShaAdd
    const uint32_t per_proof_blocks = (p.m + 63) / 64;
    auto blocks = [](uint64_t threads) { return (uint32_t)((threads + 127) / 128); };
    auto tree = [&](uint32_t *base, size_t stride, uint32_t off, uint32_t n) {
        for (uint32_t k = n; k-- > 0;) {
            prv_tree_level_kernel<<<blocks((uint64_t)p.m << k), 128, 0, s>>>(base, stride, off, k, p.m, one);
            launches++;
        }
    };
    prv_trace_kernel<<<dim3(p.m, C / 4), 512, 4 * sizeof(uint32_t) * NG, s>>>(p);
    prv_leaf_trace_kernel<<<blocks((uint64_t)p.m << G), 128, 0, s>>>(p, one);
    launches += 2;
    tree(p.tree_t, (size_t)2 * NG * 8, 0, G);
    prv_ch_commit_kernel<<<per_proof_blocks, 64, 0, s>>>(p);
    prv_cp_kernel<<<dim3(p.m, 4), 512, 5 * sizeof(uint32_t) * NG, s>>>(p);
    prv_leaf_cp_kernel<<<blocks((uint64_t)p.m << G), 128, 0, s>>>(p, one);
    launches += 3;
    tree(p.tree_c, (size_t)2 * NG * 8, 0, G);
    prv_ch_oods_point_kernel<<<per_proof_blocks, 64, 0, s>>>(p);
    prv_oods_kernel<<<(p.m * (C + 16) + PRV_OODS_WARPS - 1) / PRV_OODS_WARPS, 32 * PRV_OODS_WARPS, 0, s>>>(p);
    prv_ch_deep_kernel<<<per_proof_blocks, 64, 0, s>>>(p);
    prv_quotient_kernel<<<blocks((uint64_t)p.m << G), 128, 0, s>>>(p);
    launches += 4;
    for (uint32_t l = 0; l <= L; l++) {
        const uint32_t n = G - l;
        prv_leaf_fri_kernel<<<blocks((uint64_t)p.m << n), 128, 0, s>>>(p, l, one);
        launches++;
        tree(p.ftree, p.ftree_stride, p.ftree_off[l], n);
        prv_ch_fri_kernel<<<per_proof_blocks, 64, 0, s>>>(p, l);
        prv_fold_kernel<<<blocks((uint64_t)p.m << (n - 1)), 128, 0, s>>>(p, l);
        launches += 2;
    }
    prv_final_kernel<<<per_proof_blocks, 64, 0, s>>>(p);
    prv_decommit_kernel<<<blocks((uint64_t)p.m * Q * 32), 128, 0, s>>>(p);
    launches += 2;
    if (launch_counter) *launch_counter += launches;
    return launches;
}

} // namespace ssym
