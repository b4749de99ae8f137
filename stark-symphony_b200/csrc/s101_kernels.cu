// SPDX-License-Identifier: MIT
//
// Batched verify_proof of the STARK-101 Fibonacci-square verifier (stark101/src/verifier.simf:24-42).
//
//   S1 s101_transcript_kernel  one thread per proof   channel (sha256 chain), beta checks, query index,
//                                                     composition polynomial, FRI fold chain
//   S2 s101_merkle_kernel      one thread per Merkle path (3 trace + 2 per FRI layer); a warp = 32
//                                                     proofs x one path slot, so replicated batches do not diverge
//   S3 s101_finalize_kernel    status -> accept bitmap
//
// Field: p = 3*2^30 + 1 (not Mersenne) with the reference's exact jet semantics on arbitrary u32
// (field.simf:14-94): 64-bit add / mul followed by modulo_64, sub = add(a, (p - b) mod 2^32), division by
// extended Euclid in a bounded loop whose gcd assert makes division by zero a rejection.
#include "s101_kernels.cuh"

#include "sha256.cuh"

namespace ssym {

#define S101_P 3221225473u

__device__ __forceinline__ uint32_t add_mod(uint32_t a, uint32_t b) { return (uint32_t)(((uint64_t)a + b) % S101_P); } // field.simf:14-21
__device__ __forceinline__ uint32_t sub_mod(uint32_t a, uint32_t b) { return add_mod(a, S101_P - b); }                  // field.simf:24-27
__device__ __forceinline__ uint32_t mul_mod(uint32_t a, uint32_t b) { return (uint32_t)(((uint64_t)a * b) % S101_P); } // field.simf:30-35
__device__ __noinline__ uint32_t div_mod(uint32_t a, uint32_t b, bool &fail) {                                          // field.simf:42-66
    uint32_t t = 0, r = S101_P, new_t = 1, new_r = b;
    bool done = false;
#pragma unroll 1
    for (uint32_t counter = 0; counter < 65536u; counter++) {
        if (new_r == 0) {
            if (r != 1) fail = true; // field.simf:46
            done = true;
            break;
        }
        const uint32_t q = r / new_r;
        const uint32_t nt = sub_mod(t, mul_mod(q, new_t));
        const uint32_t nr = sub_mod(r, mul_mod(q, new_r));
        t = new_t; new_t = nt;
        r = new_r; new_r = nr;
    }
    if (!done) fail = true; // unwrap_left on Right
    return mul_mod(a, t);
}
__device__ __noinline__ uint32_t exp_mod(uint32_t a, uint32_t b) { // field.simf:76-94 (32 halvings always reach 0)
    uint32_t res = 1, base = a, e = b;
#pragma unroll 1
    while (e != 0) {
        if (e & 1u) res = mul_mod(res, base);
        base = mul_mod(base, base);
        e >>= 1;
    }
    return res;
}

__device__ __noinline__ void s101_compress(uint32_t *h, const uint32_t *blk) {
    uint32_t hh[8], w[16];
#pragma unroll
    for (int i = 0; i < 8; i++) hh[i] = h[i];
#pragma unroll
    for (int i = 0; i < 16; i++) w[i] = blk[i];
    sha_compress(hh, w);
#pragma unroll
    for (int i = 0; i < 8; i++) h[i] = hh[i];
}
// state <- SHA-256(state || extra[0..n))                      channel.simf:22-45, sha256.simf:11-15
__device__ __noinline__ void s101_hash_state(uint32_t *state, const uint32_t *extra, int n) {
    uint32_t h[8];
    sha_iv(h);
    const int nwords = 8 + n;
    const int nblocks = (nwords + 3 + 15) >> 4;
    for (int blk = 0; blk < nblocks; blk++) {
        uint32_t w[16];
#pragma unroll
        for (int j = 0; j < 16; j++) {
            const int i = blk * 16 + j;
            uint32_t v = 0;
            if (i < 8) v = state[i];
            else if (i < nwords) v = extra[i - 8];
            else if (i == nwords) v = 0x80000000u;
            else if (i == nblocks * 16 - 1) v = (uint32_t)nwords * 32u;
            w[j] = v;
        }
        s101_compress(h, w);
    }
#pragma unroll
    for (int i = 0; i < 8; i++) state[i] = h[i];
}
// channel_draw_32: state mod max limb-wise, then state <- sha256(state)     channel.simf:67-105
__device__ __forceinline__ uint32_t s101_draw(uint32_t *state, uint32_t max) {
    uint32_t r = 0;
#pragma unroll
    for (int k = 0; k < 8; k++) {
        const uint64_t v = ((uint64_t)r << 32) + state[k];
        r = max ? (uint32_t)(v % max) : (uint32_t)v; // modulo_64(x, 0) = x, truncated by <u64>::into
    }
    s101_hash_state(state, nullptr, 0);
    return r;
}

__global__ void __launch_bounds__(64) s101_transcript_kernel(S101Params p) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= p.n) return;
    const uint32_t *rec = p.blob + p.offsets[i];
    uint32_t *ctx = p.ctx + (size_t)i * S101_CTX_WORDS;
    ssym_s101_trace_t *tr = p.trace ? p.trace + i : nullptr;
    // The record's own length word is trusted only after it has been compared with the offsets array (a device-resident caller may hand over
    // anything): nothing of the record is read before that, and every later read stays below `total` <= the record's slot.  (A record may be
    // shorter than its slot: the `.wit` path packs into fixed-stride slots.)
    const uint64_t len = p.offsets[i + 1] - p.offsets[i];
    bool shape_ok = p.offsets[i + 1] >= p.offsets[i] && len >= 20 && len <= 0xffffffffull && rec[0] <= (uint32_t)len;
    const uint32_t total = shape_ok ? rec[0] : 0, n_layers = shape_ok ? rec[1] : 0;
    const uint32_t ns0 = shape_ok ? rec[2] : 0, ns1 = shape_ok ? rec[3] : 0, ns2 = shape_ok ? rec[4] : 0;
    shape_ok = shape_ok && n_layers <= SSYM_S101_MAX_LIST && ns0 <= SSYM_S101_MAX_LIST && ns1 <= SSYM_S101_MAX_LIST && ns2 <= SSYM_S101_MAX_LIST;
    const uint32_t ordinal = shape_ok ? rec[6] : 0;
    shape_ok = shape_ok && ordinal <= SSYM_S101_MAX_ORDINAL;
    uint32_t w = 20 + 8 * (ns0 + ns1 + ns2);
    if (shape_ok) {
        for (uint32_t l = 0; l < n_layers; l++) {
            if (w + 16 > total || rec[w + 11] > SSYM_S101_MAX_LIST || rec[w + 12] > SSYM_S101_MAX_LIST) { shape_ok = false; break; }
            ctx[S101_CTX_LAYER_OFF + l] = w;
            w += 16 + 8 * (rec[w + 11] + rec[w + 12]);
        }
        if (w != total) shape_ok = false;
    }
    if (!shape_ok) {
        ctx[S101_CTX_NLAYERS] = 0xffffffffu; // S2 skips this proof
        p.status[i] = SSYM_S101_ST_SHAPE;
        if (tr) tr->first_fail_layer = 0xffffffffu;
        return;
    }
    uint32_t status = 0;
    const uint32_t last_layer = rec[5];
    uint32_t state[8];
#pragma unroll
    for (int k = 0; k < 8; k++) state[k] = 0;
    { // state = sha256(p_mt_root)                              verifier.simf:27
        uint32_t h[8];
#pragma unroll
        for (int k = 0; k < 8; k++) h[k] = rec[8 + k];
        s101_hash_state(h, nullptr, 0);
#pragma unroll
        for (int k = 0; k < 8; k++) state[k] = h[k];
    }
    // fibsquare_read_coefficients                              air.simf:30-36
    const uint32_t a0 = s101_draw(state, S101_P), a1 = s101_draw(state, S101_P), a2 = s101_draw(state, S101_P);
    // fri_read_commitments_32                                  fri.simf:37-53
#pragma unroll 1
    for (uint32_t l = 0; l < n_layers; l++) {
        const uint32_t *lp = rec + ctx[S101_CTX_LAYER_OFF + l];
        s101_hash_state(state, lp, 8); // channel_mix_256
        const uint32_t random = s101_draw(state, S101_P);
        if (tr) tr->beta_drawn[l] = random;
        if (random != lp[8]) { // fri.simf:43
            status |= SSYM_S101_ST_BETA;
            if (tr) atomicOr(&tr->layer_mask[l], 8u);
        }
    }
    s101_hash_state(state, &last_layer, 1); // channel_mix_32
#pragma unroll
    for (int k = 0; k < 8; k++) ctx[S101_CTX_COMMIT + k] = state[k];
    ctx[S101_CTX_ORD] = ordinal;
    if (tr) {
        tr->query_ordinal = ordinal;
        for (int k = 0; k < 8; k++) tr->commit_state[k] = state[k];
    }
    uint32_t idx = s101_draw(state, 8192u); // verifier.simf:32
#pragma unroll 1
    for (uint32_t k = 0; k < ordinal; k++) idx = s101_draw(state, 8192u); // query ordinal k (include/ssym.h): the query phase repeated on one channel
    const uint32_t f0 = rec[16], f1 = rec[17], f2 = rec[18];
    if (tr) { // the channel keeps absorbing the three evaluations (air.simf:43); nothing downstream reads it
        uint32_t st[8];
#pragma unroll
        for (int k = 0; k < 8; k++) st[k] = state[k];
        s101_hash_state(st, &f0, 1);
        s101_hash_state(st, &f1, 1);
        s101_hash_state(st, &f2, 1);
        for (int k = 0; k < 8; k++) tr->state_final[k] = st[k];
        tr->alpha[0] = a0; tr->alpha[1] = a1; tr->alpha[2] = a2;
        tr->idx = idx;
        tr->n_layers = n_layers;
    }
    // fibsquare_calc_x / fibsquare_compose                     air.simf:58-101
    uint32_t x = mul_mod(5u, exp_mod(1734477367u, idx));
    bool div_fail = false;
    uint32_t cp_ev;
    {
        const uint32_t p0 = div_mod(sub_mod(f0, 1u), sub_mod(x, 1u), div_fail);
        const uint32_t p1 = div_mod(sub_mod(f0, 2338775057u), sub_mod(x, 2450347685u), div_fail);
        const uint32_t num0 = sub_mod(f2, add_mod(mul_mod(f0, f0), mul_mod(f1, f1)));
        const uint32_t num1 = mul_mod(mul_mod(sub_mod(x, 2342081930u), sub_mod(x, 2450347685u)), sub_mod(x, 532203874u));
        const uint32_t den = sub_mod(exp_mod(x, 1024u), 1u);
        const uint32_t p2 = div_mod(mul_mod(num0, num1), den, div_fail);
        cp_ev = add_mod(add_mod(mul_mod(p0, a0), mul_mod(p1, a1)), mul_mod(p2, a2));
    }
    if (div_fail) status |= SSYM_S101_ST_DIV;
    if (tr) { tr->x = x; tr->cp0 = cp_ev; }
    // fri_verify_32 without the Merkle halves                  fri.simf:71-91
#pragma unroll 1
    for (uint32_t l = 0; l < n_layers; l++) {
        const uint32_t *lp = rec + ctx[S101_CTX_LAYER_OFF + l];
        const uint32_t beta = lp[8], cpa = lp[9], cpb = lp[10];
        if (tr) tr->cp_ev[l] = cp_ev;
        if (cp_ev != cpa) { // fri.simf:77
            status |= SSYM_S101_ST_LAYER_CP;
            if (tr) atomicOr(&tr->layer_mask[l], 1u);
        }
        bool f = false; // fri_eval_cp_next fri.simf:55-59
        const uint32_t op0 = div_mod(add_mod(cpa, cpb), 2u, f);
        const uint32_t op1 = div_mod(sub_mod(cpa, cpb), mul_mod(x, 2u), f);
        cp_ev = add_mod(op0, mul_mod(op1, beta));
        if (f) {
            status |= SSYM_S101_ST_DIV;
            if (tr) atomicOr(&tr->layer_mask[l], 16u);
        }
        x = mul_mod(x, x);
    }
    if (tr) tr->cp_ev[n_layers] = cp_ev;
    if (cp_ev != last_layer) status |= SSYM_S101_ST_LAST; // fri.simf:90
    ctx[S101_CTX_NLAYERS] = n_layers;
    ctx[S101_CTX_IDX] = idx;
    p.status[i] = status;
}

__device__ __forceinline__ void s101_load_digest(const uint32_t *src, uint32_t (&d)[8]) { // records are only 4-byte aligned
#pragma unroll
    for (int k = 0; k < 8; k++) d[k] = __ldg(src + k);
}

// Path slots: 0..2 = trace decommitments of f(x), f(gx), f(g^2 x) (air.simf:39-56); 3 + 2l, 4 + 2l = cpa / cpb of
// FRI layer l (fri.simf:78-80).  merkle_verify_32 here has no `path == 1` assert (stark101/src/merkle.simf:39-43).
__global__ void __launch_bounds__(128, 8) s101_merkle_kernel(S101Params p, uint32_t groups, ShaMul mul) {
    const ShaAdd<8> A(mul); // adds on the FMA pipe, rounds rolled 4 x 16, ONE hashing loop: the core of stwo_merkle_kernel (sha256.cuh)
    const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    const uint32_t slot = warp / groups, group = warp % groups;
    const uint32_t i = group * 32 + lane;
    if (i >= p.n) return;
    const uint32_t *ctx = p.ctx + (size_t)i * S101_CTX_WORDS;
    const uint32_t n_layers = ctx[S101_CTX_NLAYERS];
    if (n_layers == 0xffffffffu || slot >= 3 + 2 * n_layers) return;
    const uint32_t *rec = p.blob + p.offsets[i];
    const uint32_t idx = ctx[S101_CTX_IDX];
    uint32_t value, path, n_sib, fail_bit, layer = 0, mask_bit = 0;
    const uint32_t *sib, *root;
    if (slot < 3) {
        value = rec[16 + slot];
        n_sib = rec[2 + slot];
        sib = rec + 20;
        for (uint32_t k = 0; k < slot; k++) sib += 8 * rec[2 + k];
        root = rec + 8;
        path = idx + 8u * slot + 8192u; // air.simf:40,50-54
        fail_bit = SSYM_S101_ST_TRACE_MERKLE(slot);
    } else {
        layer = (slot - 3) >> 1;
        const bool is_b = (slot - 3) & 1;
        const uint32_t *lp = rec + ctx[S101_CTX_LAYER_OFF + layer];
        const uint32_t na = lp[11], nb = lp[12];
        value = is_b ? lp[10] : lp[9];
        n_sib = is_b ? nb : na;
        sib = lp + 16 + (is_b ? 8 * na : 0);
        root = lp;
        const uint32_t domain_size = layer >= 32 ? 0u : 8192u >> layer; // divide_32(domain_size, 2) per layer
        // compute_auth_path fri.simf:63-68 (modulo_32(x, 0) = x, divide_32(x, 0) = 0)
        const uint32_t base = is_b ? idx + (domain_size >> 1) : idx;
        path = (domain_size ? base % domain_size : base) + domain_size;
        fail_bit = is_b ? SSYM_S101_ST_LAYER_MERKLE_B : SSYM_S101_ST_LAYER_MERKLE_A;
        mask_bit = is_b ? 4u : 2u;
    }
    uint32_t cur[8], nxt[8]; // nxt: the next sibling, loaded one level ahead
#pragma unroll
    for (int k = 0; k < 8; k++) cur[k] = nxt[k] = 0;
#pragma unroll 1
    for (uint32_t step = 0; step <= n_sib; step++) {
        uint32_t w[16];
        if (step == 0) { // sha256_32(value): the leaf (merkle.simf:39, sha256.simf:17-21)
            w[0] = value;
            w[1] = 0x80000000u;
#pragma unroll
            for (int k = 2; k < 15; k++) w[k] = 0;
            w[15] = 32u;
            if (n_sib) s101_load_digest(sib, nxt);
        } else {
            const bool cur_left = (path & 1u) == 0;
#pragma unroll
            for (int k = 0; k < 8; k++) {
                w[k] = cur_left ? cur[k] : nxt[k];
                w[8 + k] = cur_left ? nxt[k] : cur[k];
            }
            if (step < n_sib) s101_load_digest(sib + 8 * step, nxt);
            path >>= 1;
        }
        sha_iv(cur);
        sha_compress_rolled<8>(cur, w, A);
        if (step) sha_compress_pad64_rolled<8>(cur, A);
    }
    bool ok = true;
#pragma unroll
    for (int k = 0; k < 8; k++) ok = ok && (cur[k] == __ldg(root + k));
    if (!ok) atomicOr(&p.status[i], fail_bit);
    if (p.trace) {
        ssym_s101_trace_t *tr = p.trace + i;
        if (slot < 3)
            for (int k = 0; k < 8; k++) tr->trace_root[slot][k] = cur[k];
        else if (!ok)
            atomicOr(&tr->layer_mask[layer], mask_bit);
    }
}

__global__ void __launch_bounds__(256) s101_finalize_kernel(S101Params p, uint32_t *accept_bits) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t s = i < p.n ? p.status[i] : 1u;
    const uint32_t ballot = __ballot_sync(0xffffffffu, s == 0);
    if ((threadIdx.x & 31) == 0 && i < p.n && accept_bits) accept_bits[i >> 5] = ballot;
    if (i < p.n && p.trace) {
        ssym_s101_trace_t *tr = p.trace + i;
        tr->status = s;
        if (!(s & SSYM_S101_ST_SHAPE)) {
            uint32_t ff = 0xffffffffu;
            for (uint32_t l = 0; l < tr->n_layers && l < SSYM_S101_MAX_LIST; l++)
                if (tr->layer_mask[l]) { ff = l; break; }
            tr->first_fail_layer = ff;
        }
    }
}

// Multi-query grouping (ssym_stark101_verify_multi_batch): one thread per proof over its n_queries records — ordinal k in slot k, the channel state
// after the commitments equal to the first record's (identical commitments), every record accepted.
__global__ void __launch_bounds__(256) s101_group_kernel(S101Params p, uint32_t n_queries, uint32_t n_proofs, uint32_t *accept_bits) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    bool ok = i < n_proofs;
    if (i < n_proofs) {
        const uint32_t r0 = i * n_queries;
        const uint32_t *c0 = p.ctx + (size_t)r0 * S101_CTX_WORDS;
        const bool first_shape = (p.status[r0] & SSYM_S101_ST_SHAPE) != 0;
        for (uint32_t k = 0; k < n_queries; k++) {
            const uint32_t r = r0 + k;
            uint32_t st = p.status[r];
            if (!(st & SSYM_S101_ST_SHAPE)) {
                const uint32_t *c = p.ctx + (size_t)r * S101_CTX_WORDS;
                bool same = !first_shape && c[S101_CTX_ORD] == k;
                for (int w = 0; w < 8; w++) same = same && c[S101_CTX_COMMIT + w] == c0[S101_CTX_COMMIT + w];
                if (!same) {
                    st |= SSYM_S101_ST_GROUP;
                    p.status[r] = st;
                    if (p.trace) p.trace[r].status = st;
                }
            }
            ok = ok && st == 0;
        }
    }
    const uint32_t ballot = __ballot_sync(0xffffffffu, ok);
    if ((threadIdx.x & 31) == 0 && i < n_proofs) accept_bits[i >> 5] = ballot;
}

void launch_s101_verify(const S101Params &p, uint32_t *accept_bits, cudaStream_t s, uint64_t *launch_counter, Profiler *prof) {
    if (p.n == 0) return;
    if (prof) prof->begin(4, s);
    s101_transcript_kernel<<<(p.n + 63) / 64, 64, 0, s>>>(p);
    if (prof) { prof->end(4, s); prof->begin(5, s); }
    const uint32_t groups = (p.n + 31) / 32;
    const uint32_t slots = 3 + 2 * p.max_layers;
    const uint64_t warps = (uint64_t)groups * slots;
    s101_merkle_kernel<<<(uint32_t)((warps + 3) / 4), 128, 0, s>>>(p, groups, sha_mul_consts());
    if (prof) { prof->end(5, s); prof->begin(6, s); }
    s101_finalize_kernel<<<(p.n + 255) / 256, 256, 0, s>>>(p, accept_bits);
    if (prof) prof->end(6, s);
    if (launch_counter) *launch_counter += 3;
}

void launch_s101_verify_multi(const S101Params &p, uint32_t n_queries, uint32_t *accept_bits, cudaStream_t s, uint64_t *launch_counter, Profiler *prof) {
    if (p.n == 0 || n_queries == 0) return;
    launch_s101_verify(p, nullptr, s, launch_counter, prof); // every record on its own (per-record status and trace), then the proofs
    const uint32_t n_proofs = p.n / n_queries;
    s101_group_kernel<<<(n_proofs + 255) / 256, 256, 0, s>>>(p, n_queries, n_proofs, accept_bits);
    if (launch_counter) *launch_counter += 1;
}

} // namespace ssym
