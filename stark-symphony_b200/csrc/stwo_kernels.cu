// SPDX-License-Identifier: MIT
// Kernels of the batched Stwo verifier; see stwo_kernels.cuh for the decomposition.
#include "stwo_kernels.cuh"
#include "channel.cuh"
#include "deep.cuh"

#include <algorithm>
#include <cstdlib>

#ifndef SSYM_DEFAULT_ADDMODE
#define SSYM_DEFAULT_ADDMODE 8
#endif
#ifndef SSYM_MERKLE_THREADS
#define SSYM_MERKLE_THREADS 32 // threads per CTA of the per-query Merkle kernel.  A warp never talks to another, so any multiple of 32 computes the same; with
                              // one warp per CTA a finished chain frees its slot at once for the next (shorter) one — a lone 1024-proof launch 230 -> 222 us
                              // (64: 227, 256: 234), the pipelined rate is the same
#endif
#ifndef SSYM_MERKLE_MINB
#define SSYM_MERKLE_MINB (1024 / SSYM_MERKLE_THREADS) // resident CTAs per SM the Merkle kernels are compiled for (32 warps -> 64 registers)
#endif
#ifndef SSYM_DEFAULT_ROLLED
#define SSYM_DEFAULT_ROLLED 1
#endif

namespace ssym {

typedef StwoCtxLayout CX;

__device__ __noinline__ QM31 qm31_mul_nl(QM31 x, QM31 y) { return qm31_mul(x, y); }
__device__ __noinline__ QM31 qm31_inv_nl(QM31 x, bool &fail) { return qm31_inv(x, fail); }

// composition_poly_eval_from_partitions                       evals/composition_poly.simf:38-44
__device__ QM31 cp_from_partitions(QM31 c0, QM31 c1, QM31 c2, QM31 c3) {
    QM31 res = qm31_add(c0, qm31_mul_nl(c1, qm31(0, 1, 0, 0)));
    res = qm31_add(res, qm31_mul_nl(c2, qm31(0, 0, 1, 0)));
    res = qm31_add(res, qm31_mul_nl(c3, qm31(0, 0, 0, 1)));
    return res;
}


// ------------------------------------------------------------------------------------------
// K1: the Fiat-Shamir channel of verify_proof (verifier.simf:36-51).  The channel state only ever
// absorbs proof data (roots, samples, nonce): none of the field arithmetic feeds back into it, so the
// whole transcript is one dependent chain of ~46 SHA-256 compressions per proof — one thread per
// proof, raw draws written to the per-proof context for K2.
// ------------------------------------------------------------------------------------------
// The transcript as data: one loop, ONE copy of the (rolled) compression, digest and message block in registers.  A step either absorbs
// proof words (digest <- SHA-256(digest || words), channel.simf:154-173) or draws (SHA-256(digest || n_sent), channel.simf:36-44); what
// happens with a draw (felt retry loop channel.simf:115-141, queries fri/queries.simf:14-43) is decided after the hash.
enum TxState : uint32_t {
    TX_MIX_CONST, TX_MIX_TRACE, TX_DRAW_CP_ALPHA, TX_MIX_CP,  // evals_commit           evals/commit.simf:20-35
    TX_DRAW_OODS_T, TX_MIX_OODS, TX_DRAW_DEEP_ALPHA,          // oods                   deep/oods.simf:44-64
    TX_MIX_FRI_ROOT, TX_DRAW_FRI_ALPHA, TX_MIX_LAST,          // fri_commit             fri/commit.simf:72-85
    TX_MIX_NONCE,                                              // check_proof_of_work    pow.simf:22-35
    TX_DRAW_QUERIES, TX_DONE                                   // fri_generate_queries   fri/queries.simf:30-43
};

// The per-proof scalars of verify_proof (a thread serves one proof here; the query kernel would redo them in every lane): the OODS point
// (channel.simf:143-151), the composition-polynomial check (deep/oods.simf:52-58), the sample point of the CP columns, the powers of the
// DEEP coefficient.  Adds its failures to `status` and stores the proof's status word.
__device__ __noinline__ uint32_t stwo_scalars(const StwoParams &p, uint32_t i, QM31 oods_t, QM31 cp_alpha, QM31 deep_alpha, uint32_t status) {
    const ssym_stwo_layout_t &lo = p.lo;
    const uint32_t *pk = p.packed + (size_t)i * lo.stride_words;
    uint32_t *ctx = p.ctx + (size_t)i * CX::WORDS;
    ssym_stwo_trace_t *tr = p.trace ? p.trace + i : nullptr;
    const uint32_t C = SSYM_STWO_COLUMNS(&p.cfg), NCOL = C + SSYM_NUM_CP_PARTITIONS;
    bool inv_zero = false;
    QM31 px, py;
    { // channel_draw_qm31_point channel.simf:143-151
        const QM31 t_sq = qm31_mul_nl(oods_t, oods_t);
        const QM31 inv = qm31_inv_nl(qm31_add(qm31_one(), t_sq), inv_zero);
        px = qm31_mul_nl(qm31_sub(qm31_one(), t_sq), inv);
        py = qm31_mul_nl(qm31_add(oods_t, oods_t), inv);
    }
    const QM31 pxy = qm31_mul_nl(px, py);
    {
        // composition_poly_eval_from_decomposed composition_poly.simf:47-59 (index = 4*coord + poly): F_a + y F_b + x F_c + xy F_d
        const uint32_t *e = pk + lo.off_oods_cp;
        QM31 sampled = qm31_zero();
#pragma unroll 1
        for (uint32_t poly = 0; poly < 4; poly++) {
            const QM31 part = cp_from_partitions(qm31_load4(e + 4 * poly), qm31_load4(e + 4 * (4 + poly)), qm31_load4(e + 4 * (8 + poly)),
                                                 qm31_load4(e + 4 * (12 + poly)));
            const QM31 factor = poly == 1 ? py : poly == 2 ? px : pxy;
            sampled = poly == 0 ? part : qm31_add(sampled, qm31_mul_nl(part, factor));
        }
        // eval_composition_poly wide_fibonacci.simf:24-62
        QM31 acc = qm31_zero(), a = qm31_zero(), b = qm31_zero();
        uint32_t skip_2 = 0;
#pragma unroll 1
        for (uint32_t col = 0; col < C; col++) {
            const QM31 c = qm31_load4(pk + lo.off_oods_trace + 4 * col);
            if (skip_2 == 2) {
                const QM31 constraint = qm31_sub(c, qm31_add(qm31_mul_nl(b, b), qm31_mul_nl(a, a)));
                acc = qm31_add(qm31_mul_nl(acc, cp_alpha), constraint);
            } else {
                skip_2++;
            }
            a = b;
            b = c;
        }
        // vanishing_poly_eval composition_poly.simf:66-71: pi^(log_size-1)(x)
        const uint32_t n_iter = (p.cfg.trace_log - 1u) & 0xff;
        QM31 v = px;
#pragma unroll 1
        for (uint32_t counter = 0; counter < 256 && counter != n_iter; counter++) {
            const QM31 sq = qm31_mul_nl(v, v);
            v = qm31_sub(qm31_add(sq, sq), qm31_one());
        }
        const QM31 cp_eval = qm31_mul_nl(acc, qm31_inv_nl(v, inv_zero));
        if (!qm31_eq(cp_eval, sampled)) status |= SSYM_ST_OODS_CP_MISMATCH; // deep/oods.simf:58
        if (inv_zero) status |= SSYM_ST_OODS_INV_ZERO;
        if (tr) {
            qm31_store(tr->oods_x, px);
            qm31_store(tr->oods_y, py);
            qm31_store(tr->cp_eval, cp_eval);
            qm31_store(tr->cp_sampled, sampled);
        }
    }
    qm31_store4(ctx + CX::PX, px);
    qm31_store4(ctx + CX::PY, py);
    if (SSYM_MODE_SEMANTICS(p.cfg.mode) == SSYM_MODE_REF_LITERAL) { // all C + 16 columns are sampled at P (fri/answers.simf:116-125)
        qm31_store4(ctx + CX::P2X, px);
        qm31_store4(ctx + CX::P2Y, py);
    } else { // the 16 CP partitions are sampled at 2*P
        qm31_store4(ctx + CX::P2X, qm31_point_dbl_x(px));
        qm31_store4(ctx + CX::P2Y, qm31_add(pxy, pxy));
    }
    { // alpha^(k+1): the running product of fri/answers.simf:52,70
        QM31 a = deep_alpha;
#pragma unroll 1
        for (uint32_t k = 0; k <= NCOL; k++) {
            qm31_store4(ctx + CX::ALPHA_POW + 4 * k, a);
            a = qm31_mul_nl(a, deep_alpha);
        }
    }
    return status;
}

// One thread runs the transcripts of NP proofs in lockstep (the steps of the program are the same for every proof of a configuration; only a
// felt draw that has to be repeated — probability 2^-29 — makes one proof wait for the other).
#ifndef K1_ADDMODE
#define K1_ADDMODE 1 // adds as IMAD, as in the Merkle kernels.  Plain adds (0: 3-input IADD3s, shorter chains) make a lone launch 4 % faster (0.123 vs
                     // 0.129 ms), but in the pipelined loop this kernel runs under Merkle kernels that are bound by the ALU pipe, where every ALU slot counts
#endif
#ifndef K1_R_ADDMODE
#define K1_R_ADDMODE 4 // the round warp of the warp-specialised kernel: a lone dependent chain whose latency is what a lone launch costs.  a' = T1 + Sigma0 + Maj
                       // as ONE 3-input IADD3 shortens it: 88.6 -> 82.5 us per launch (mode 1: 88.6, plain adds: 84.7, T1 chain only: 82.8); its ALU
                       // slots are 1.5 % of a pass
#endif
template <int NP>
__global__ void __launch_bounds__(64) stwo_channel_kernel(StwoParams p, ShaMul mul) {
    const ShaAdd<K1_ADDMODE> A(mul);
    const uint32_t i0 = (blockIdx.x * blockDim.x + threadIdx.x) * NP;
    if (p.dd.enabled && blockIdx.x == 0) // the bin counters of stwo_plan_kernel / stwo_check_kernel, later in this stream
        for (uint32_t t = threadIdx.x; t < 2 * STWO_DEDUP_MAX_BINS; t += blockDim.x) p.dd.bin_count[t] = 0;
    if (i0 >= p.n) return;
    const ssym_stwo_layout_t &lo = p.lo;
    const uint32_t Q = p.cfg.n_queries, L = p.cfg.n_fri_layers, G = p.cfg.lde_log;
    const uint32_t NCOL = SSYM_STWO_COLUMNS(&p.cfg) + SSYM_NUM_CP_PARTITIONS; // columns entering the DEEP quotient
    // proof k of this thread; a thread at the end of an odd batch runs its last proof twice and stores it once
    uint32_t idx[NP];
    bool live[NP];
    const uint32_t *pk[NP];
    uint32_t *ctx[NP];
    ssym_stwo_trace_t *tr[NP];
    uint32_t status[NP], n_sent[NP], tries[NP], retries[NP], d[NP][8]; // channel_init channel.simf:31-33: ChannelState = (digest, n_sent) = (0, 0)
    bool exhausted[NP], settled[NP];
    QM31 felt[NP], oods_t[NP], deep_alpha[NP], cp_alpha[NP];
#pragma unroll
    for (int k = 0; k < NP; k++) {
        live[k] = i0 + k < p.n;
        idx[k] = live[k] ? i0 + k : i0;
        pk[k] = p.packed + (size_t)idx[k] * lo.stride_words;
        ctx[k] = p.ctx + (size_t)idx[k] * CX::WORDS;
        tr[k] = p.trace && live[k] ? p.trace + idx[k] : nullptr;
        status[k] = 0; n_sent[k] = 0; tries[k] = 0; retries[k] = 0;
        exhausted[k] = false; settled[k] = false;
        felt[k] = oods_t[k] = deep_alpha[k] = cp_alpha[k] = qm31_zero();
#pragma unroll
        for (int j = 0; j < 8; j++) d[k][j] = 0;
    }
    uint32_t state = TX_MIX_CONST, layer = 0, q0 = 0;
    const uint32_t query_mask = shl32(G & 0xff, 1u) - 1u;
#pragma unroll 1
    while (state != TX_DONE) {
        // ---- what this step hashes after the digest: n words at word offset `off` of the packed proof, or the draw counter ----
        uint32_t off = 0, n = 1;
        bool draw = false;
        switch (state) {
        case TX_MIX_CONST: off = lo.off_commit; n = 8; break;
        case TX_MIX_TRACE: off = lo.off_commit + 8; n = 8; break;
        case TX_MIX_CP: off = lo.off_commit + 16; n = 8; break;
        case TX_MIX_OODS: off = lo.off_oods_trace; n = 4 * NCOL; break; // C trace + 16 CP samples are contiguous in the packed header
        case TX_MIX_FRI_ROOT: off = layer == 0 ? lo.off_fri_first_root : lo.off_fri_inner_root + 8 * (layer - 1); n = 8; break;
        case TX_MIX_LAST: off = lo.off_last_coeff; n = 4; break;        // channel_mix_line_poly
        case TX_MIX_NONCE: off = lo.off_pow_nonce; n = 2; break;        // channel_mix_u64: {hi, lo} big-endian
        default: draw = true; break;
        }
        // ---- SHA-256(digest || words), FIPS 180-4 padding; message word m = 8 + j of the stream ----
        uint32_t h[NP][8];
#pragma unroll
        for (int k = 0; k < NP; k++) sha_iv(h[k]);
        const uint32_t nwords = 8 + n, nblocks = (nwords + 3 + 15) >> 4;
#pragma unroll 1
        for (uint32_t b = 0; b < nblocks; b++) {
            if (nwords == 16 && b == 1) { // the 12 digest-sized mixes: the second block is the constant padding block of a 64-byte message
#pragma unroll
                for (int k = 0; k < NP; k++) sha_compress_pad64_rolled<K1_ADDMODE>(h[k], A);
                break;
            }
            uint32_t w[NP][16];
#pragma unroll
            for (int j = 0; j < 16; j++) {
                const uint32_t m = b * 16 + j;
#pragma unroll
                for (int k = 0; k < NP; k++) {
                    uint32_t v = 0;
                    if (m < 8) v = d[k][j & 7]; // only in block 0, where m == j
                    else if (m < nwords) v = draw ? n_sent[k] : __ldg(pk[k] + off + (m - 8));
                    else if (m == nwords) v = 0x80000000u;
                    else if (m == nblocks * 16 - 1) v = nwords * 32u;
                    w[k][j] = v;
                }
            }
            sha_compress_rolled_n<K1_ADDMODE, NP>(h, w, A);
        }
        // ---- what the step does with the hash ----
        if (!draw) { // channel_mix_*: the digest moves, the counter restarts
#pragma unroll
            for (int k = 0; k < NP; k++) {
#pragma unroll
                for (int j = 0; j < 8; j++) d[k][j] = h[k][j];
                n_sent[k] = 0;
                if (state == TX_MIX_CP && tr[k])
                    for (int j = 0; j < 8; j++) tr[k]->digest_commit[j] = d[k][j];
                if (state == TX_MIX_LAST && tr[k])
                    for (int j = 0; j < 8; j++) tr[k]->digest_fri[j] = d[k][j];
                if (state == TX_MIX_NONCE) { // check_proof_of_work pow.simf:22-35
                    const uint64_t value = ((uint64_t)__byte_perm(d[k][7], 0, 0x0123) << 32) | __byte_perm(d[k][6], 0, 0x0123);
                    if (!(value < p.cfg.pow_target)) status[k] |= SSYM_ST_POW_FAIL;
                    if (tr[k]) {
                        for (int j = 0; j < 8; j++) tr[k]->digest_pow[j] = d[k][j];
                        tr[k]->pow_value[0] = (uint32_t)(value >> 32);
                        tr[k]->pow_value[1] = (uint32_t)value;
                    }
                }
            }
            switch (state) {
            case TX_MIX_CONST: state = TX_MIX_TRACE; break;
            case TX_MIX_TRACE: state = TX_DRAW_CP_ALPHA; break;
            case TX_MIX_CP: state = TX_DRAW_OODS_T; break;
            case TX_MIX_OODS: state = TX_DRAW_DEEP_ALPHA; break;
            case TX_MIX_FRI_ROOT: state = TX_DRAW_FRI_ALPHA; break;
            case TX_MIX_LAST: state = TX_MIX_NONCE; break;
            default: state = TX_DRAW_QUERIES; break; // TX_MIX_NONCE
            }
            continue;
        }
        if (state == TX_DRAW_QUERIES) { // channel_draw_u256 x ceil(Q / 8): 8 queries per draw, masked to the LDE domain; no sort, no dedup (fri/queries.simf:41)
#pragma unroll
            for (int k = 0; k < NP; k++) {
                n_sent[k] = n_sent[k] + 1u;
                for (uint32_t j = 0; live[k] && j < 8 && q0 + j < Q; j++) {
                    ctx[k][CX::QUERIES + q0 + j] = h[k][j] & query_mask;
                    if (tr[k]) tr[k]->queries[q0 + j] = h[k][j] & query_mask;
                }
            }
            q0 += 8;
            if (q0 >= Q) state = TX_DONE;
            continue;
        }
        // channel_draw_qm31 = channel_draw_m31x4 (channel.simf:115-141): retry (<= 256 draws) until the first four words are < 2p
        bool all_settled = true;
#pragma unroll
        for (int k = 0; k < NP; k++) {
            if (!settled[k]) { // this hash was a draw of proof k
                const bool ok = h[k][0] < 4294967294u && h[k][1] < 4294967294u && h[k][2] < 4294967294u && h[k][3] < 4294967294u;
                tries[k]++;
                if (ok || tries[k] == 256) {
                    retries[k] += tries[k] - 1u;
                    settled[k] = true;
                    exhausted[k] = exhausted[k] || !ok;
                    felt[k] = qm31(m31_reduce(h[k][0]), m31_reduce(h[k][1]), m31_reduce(h[k][2]), m31_reduce(h[k][3]));
                }
                n_sent[k] = n_sent[k] + 1u; // channel_draw_u256 channel.simf:36-44
            }
            all_settled = all_settled && settled[k];
        }
        if (!all_settled) continue; // a settled proof hashes along while the other repeats its draw; that hash is not looked at
#pragma unroll
        for (int k = 0; k < NP; k++) {
            settled[k] = false;
            tries[k] = 0;
            switch (state) {
            case TX_DRAW_CP_ALPHA:
                cp_alpha[k] = felt[k];
                if (live[k]) qm31_store4(ctx[k] + CX::CP_ALPHA, felt[k]);
                if (tr[k]) qm31_store(tr[k]->cp_alpha, felt[k]);
                break;
            case TX_DRAW_OODS_T: oods_t[k] = felt[k]; break; // channel.simf:143-144
            case TX_DRAW_DEEP_ALPHA:
                deep_alpha[k] = felt[k];
                if (live[k]) qm31_store4(ctx[k] + CX::DEEP_ALPHA, felt[k]);
                if (tr[k]) {
                    for (int j = 0; j < 8; j++) tr[k]->digest_oods[j] = d[k][j];
                    qm31_store(tr[k]->deep_alpha, felt[k]);
                }
                break;
            default: // TX_DRAW_FRI_ALPHA
                if (live[k]) qm31_store4(ctx[k] + CX::FRI_ALPHA + 4 * layer, felt[k]);
                if (tr[k]) qm31_store(tr[k]->fri_alpha[layer], felt[k]);
                break;
            }
        }
        switch (state) {
        case TX_DRAW_CP_ALPHA: state = TX_MIX_CP; break;
        case TX_DRAW_OODS_T: state = TX_MIX_OODS; break;
        case TX_DRAW_DEEP_ALPHA: state = TX_MIX_FRI_ROOT; break;
        default: layer++; state = layer <= L ? TX_MIX_FRI_ROOT : TX_MIX_LAST; break;
        }
    }
#pragma unroll 1
    for (int k = 0; k < NP; k++) {
        if (!live[k]) continue;
        uint32_t used = Q;
        if (p.cfg.mode & SSYM_MODE_QUERY_DEDUP) { // include/ssym.h: sort the drawn queries, keep the distinct ones in slots [0, U), zero the rest
            uint32_t *qs = ctx[k] + CX::QUERIES;
            for (uint32_t a = 1; a < Q; a++) { // insertion sort of <= 16 words in the proof's own context
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
            for (uint32_t a = used; a < Q; a++) qs[a] = 0;
            if (tr[k])
                for (uint32_t a = 0; a < Q; a++) tr[k]->queries[a] = qs[a];
        }
        ctx[k][CX::N_USED] = used;
        if (tr[k]) tr[k]->n_queries_used = used;
        uint32_t st = status[k];
        if (exhausted[k]) st |= SSYM_ST_DRAW_EXHAUSTED;
        if (SSYM_MODE_SEMANTICS(p.cfg.mode) == SSYM_MODE_REF_LITERAL && ((G - (L + 1u)) & 0xff) != 0) st |= SSYM_ST_FINAL_LOG; // fri/verify.simf:127
        if (tr[k]) tr[k]->draw_retries = retries[k];
        p.status[idx[k]] = stwo_scalars(p, idx[k], oods_t[k], cp_alpha[k], deep_alpha[k], st);
    }
}

// ------------------------------------------------------------------------------------------
// K1, warp-specialised (the default).  A transcript is a dependent chain, and a lone warp issues at most one instruction every ~2 cycles per pipe
// (ncu on the one-thread-per-proof kernel above: 0.34 instructions per cycle, 1.15 cycles of fixed-latency wait per issue): its latency is its
// instruction count on the chain.  Two cuts:
//  * only the MIXES are a chain.  A draw hashes digest || counter and leaves the digest alone (channel.simf:36-44); a mix hashes digest || proof data
//    and restarts the counter (:154-172); and what is mixed is proof data only.  So of the 46 compressions 32 (the mixes) depend on one another and
//    the 14 draws each hang off the digest of the mix before them: they run on a warp of their own, beside the chain;
//  * the instruction stream of the chain itself is cut across two warps on two schedulers.
//   warp R  the 64 rounds of every compression of a mix and nothing else: K[t] + W[t] comes from shared memory       (~1000 instructions / compression)
//   warp S  which mix comes next, the PoW check and the message schedule: it assembles each block and produces K + W for rounds 16 g .. 16 g + 15 one
//           group AHEAD of warp R (double-buffered; one named barrier per group); after a mix it publishes the digest for warp D (one mbarrier per
//           digest: arrive = release, no reuse, so the chain never waits for the draws)
//   warp D  the draws, in order: cp_alpha, oods_t, deep_alpha, one folding coefficient per FRI layer, the queries — retries, sorting / de-duplication
//           of the queries included; each draw is one block, compressed here with its own schedule
//   warp F  the per-proof scalars (stwo_scalars: OODS point, composition-polynomial check, powers of the DEEP coefficient) as soon as the DEEP
//           coefficient is drawn, i.e. while fri_commit / PoW / queries are still running
// Same values, same hash inputs; lane l of every warp serves proof 32 * blockIdx.x + l.
// ------------------------------------------------------------------------------------------
// bar.sync / bar.arrive are warp-aligned instructions: the lanes of a warp that went separate ways (a lane-0 store, a per-lane loop) reconverge first
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
    __syncwarp();
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
__device__ __forceinline__ void named_bar_arrive(int id, int nthreads) {
    __syncwarp();
    asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
#define K1_BAR_RS 1 // warps R + S: group produced / group consumed / digest ready / control ready (k1_bar_rs)
#define K1_BAR_F_GO 2   // D arrives, F waits: the three drawn elements the scalars need are in shared memory
#define K1_BAR_F_DONE 3 // F arrives, S waits: the scalars' status bits are in shared memory
#define K1_BAR_D_DONE 4 // D arrives, S waits: the draws are stored, their status bits are in shared memory
// The R <-> S hand-over barrier.  compute-sanitizer's synccheck reports "divergent thread(s) in block" whenever the warps of a named barrier arrive
// from different program locations (tools/synccheck_probe.cu: the plain two-warp producer / consumer pattern of the PTX manual is flagged, the same
// barrier behind one location is not), so a -DSSYM_SYNCCHECK build (build.py --synccheck, tools/sanitize.sh synccheck) puts it into ONE non-inlined
// function and the tool then checks the protocol itself: clean.  The release build inlines it: the call costs a lone launch 12 us (0.101 vs 0.089 ms).
#ifdef SSYM_SYNCCHECK
__device__ __noinline__ void k1_bar_rs() {
    __syncwarp();
    asm volatile("bar.sync 1, 64;" ::: "memory");
}
#else
__device__ __forceinline__ void k1_bar_rs() { named_bar_sync(K1_BAR_RS, 64); }
#endif
// mbarrier (shared-memory barrier object, phase 0 only): the S -> D hand-over of a digest.  32 arrivals complete the phase; arrive has release, a
// successful try_wait acquire semantics at CTA scope; compute-sanitizer's racecheck / synccheck follow it.
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" ::"r"((uint32_t)__cvta_generic_to_shared(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile("{\n\t.reg .pred p;\n\tWAIT:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra DONE;\n\tbra WAIT;\n\tDONE:\n\t}" ::"r"(
                     (uint32_t)__cvta_generic_to_shared(bar)),
                 "r"(parity)
                 : "memory");
}
#define K1_HDR_WORDS (24 + 4 * SSYM_MAX_COLUMNS + 64 + 8 + 8 * (SSYM_MAX_FRI_LAYERS - 1) + 4 + 2 + 6) // the largest header: ssym_stwo_layout's off_qvals

__global__ void __launch_bounds__(128) stwo_channel_ws_kernel(StwoParams p, ShaMul mul) {
    __shared__ uint32_t s_kw[2][16][32]; // K + W of one 16-round group per buffer, [round][lane]
    __shared__ uint32_t s_dig[8][32];    // digest of the finished message (R -> S)
    __shared__ uint32_t s_ctl[4];        // {blocks of the next message, second block is the constant padding block, done}
    __shared__ uint32_t s_felt[12][32];  // oods_t, cp_alpha, deep_alpha (S -> F)
    __shared__ uint32_t s_fbits[32];     // status bits of the scalars (F -> S)
    __shared__ uint32_t s_dbits[32];     // status bits of the draws (D -> S)
    __shared__ __align__(8) uint64_t s_mbar[SSYM_MAX_FRI_LAYERS + 5]; // one mbarrier per published digest (S arrives, D waits): no reuse, no flow control
    extern __shared__ __align__(16) uint32_t s_dyn[]; // the published digests, (L + 5) x [8][32] words
    // Everything the channel absorbs lies in the proof's header (roots, OODS samples, last coefficient, nonce: words [0, off_qvals)).  The CTA
    // copies the 32 headers into shared memory first (coalesced 128-bit loads, all in flight together), so that no global-memory latency sits
    // between two compressions of the chain: warp S assembles a block from shared memory.  [word][proof], rows padded to 33 against bank conflicts.
    __shared__ uint32_t s_hw[K1_HDR_WORDS][33];
    const ShaAdd<K1_ADDMODE> A(mul);
    const uint32_t role = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t i = blockIdx.x * 32 + lane;
    const bool live = i < p.n;
    const uint32_t idx = live ? i : p.n - 1; // idle lanes of the last CTA replay the last proof and store nothing
    if (p.dd.enabled && blockIdx.x == 0) // the bin counters of stwo_plan_kernel / stwo_check_kernel, later in this stream
        for (uint32_t t = threadIdx.x; t < 2 * STWO_DEDUP_MAX_BINS; t += blockDim.x) p.dd.bin_count[t] = 0;
    {
        const uint32_t hq = p.lo.off_qvals / 4, total = 32 * hq; // uint4s per header (sections are 32-byte aligned)
#pragma unroll 1
        if (threadIdx.x < p.cfg.n_fri_layers + 5) mbar_init(&s_mbar[threadIdx.x], 32);
        for (uint32_t t0 = threadIdx.x; t0 < total; t0 += 8 * 128) { // eight loads in flight per thread before the first store
            uint4 v[8];
#pragma unroll
            for (int u = 0; u < 8; u++) {
                const uint32_t t = t0 + u * 128;
                v[u] = make_uint4(0, 0, 0, 0);
                if (t < total) {
                    const uint32_t ip = min(blockIdx.x * 32 + t / hq, p.n - 1);
                    v[u] = __ldg(reinterpret_cast<const uint4 *>(p.packed + (size_t)ip * p.lo.stride_words) + t % hq);
                }
            }
#pragma unroll
            for (int u = 0; u < 8; u++) {
                const uint32_t t = t0 + u * 128;
                if (t < total) {
                    const uint32_t pr = t / hq, q4 = t % hq;
                    s_hw[4 * q4 + 0][pr] = v[u].x; s_hw[4 * q4 + 1][pr] = v[u].y; s_hw[4 * q4 + 2][pr] = v[u].z; s_hw[4 * q4 + 3][pr] = v[u].w;
                }
            }
        }
        __syncthreads();
    }

    if (role == 0) { // ---- warp R: rounds ------------------------------------------------------------------------------------
        const ShaAdd<K1_R_ADDMODE> AR(mul);
        uint32_t h[8];
        for (;;) {
            k1_bar_rs(); // control ready
            const uint32_t nblocks = s_ctl[0], pad64 = s_ctl[1];
            if (s_ctl[2]) break;
            sha_iv(h);
#pragma unroll 1
            for (uint32_t b = 0; b < nblocks; b++) {
                if (pad64 && b == 1) { sha_compress_pad64_rolled<K1_R_ADDMODE>(h, AR); continue; }
                uint32_t a = h[0], bb = h[1], c = h[2], d = h[3], e = h[4], f = h[5], g = h[6], hh = h[7];
#pragma unroll 1
                for (uint32_t grp = 0; grp < 4; grp++) {
                    k1_bar_rs(); // group `grp` produced (and group grp - 1 consumed: its buffer may be overwritten)
                    const uint32_t(*kw)[32] = s_kw[grp & 1];
                    SSYM_SHA_ROUND4(AR, a, bb, c, d, e, f, g, hh, kw[0][lane], kw[1][lane], kw[2][lane], kw[3][lane]);
                    SSYM_SHA_ROUND4(AR, e, f, g, hh, a, bb, c, d, kw[4][lane], kw[5][lane], kw[6][lane], kw[7][lane]);
                    SSYM_SHA_ROUND4(AR, a, bb, c, d, e, f, g, hh, kw[8][lane], kw[9][lane], kw[10][lane], kw[11][lane]);
                    SSYM_SHA_ROUND4(AR, e, f, g, hh, a, bb, c, d, kw[12][lane], kw[13][lane], kw[14][lane], kw[15][lane]);
                }
                h[0] += a; h[1] += bb; h[2] += c; h[3] += d; h[4] += e; h[5] += f; h[6] += g; h[7] += hh;
            }
#pragma unroll
            for (int k = 0; k < 8; k++) s_dig[k][lane] = h[k];
            k1_bar_rs(); // digest ready
        }
        return;
    }
    if (role == 2) { // ---- warp F: the per-proof scalars ----------------------------------------------------------------------
        named_bar_sync(K1_BAR_F_GO, 64);
        const QM31 oods_t = qm31(s_felt[0][lane], s_felt[1][lane], s_felt[2][lane], s_felt[3][lane]);
        const QM31 cp_alpha = qm31(s_felt[4][lane], s_felt[5][lane], s_felt[6][lane], s_felt[7][lane]);
        const QM31 deep_alpha = qm31(s_felt[8][lane], s_felt[9][lane], s_felt[10][lane], s_felt[11][lane]);
        uint32_t bits = 0;
        if (live) bits = stwo_scalars(p, idx, oods_t, cp_alpha, deep_alpha, 0u);
        s_fbits[lane] = bits;
        __threadfence_block();
        named_bar_arrive(K1_BAR_F_DONE, 64);
        return;
    }
    const ssym_stwo_layout_t &lo = p.lo;
    const uint32_t Q = p.cfg.n_queries, L = p.cfg.n_fri_layers, G = p.cfg.lde_log;
    uint32_t *ctx = p.ctx + (size_t)idx * CX::WORDS;
    ssym_stwo_trace_t *tr = p.trace && live ? p.trace + idx : nullptr;
    uint32_t(*s_dq)[8][32] = reinterpret_cast<uint32_t(*)[8][32]>(s_dyn); // digest after the mix that precedes draw group k (S -> D), k = 0 .. L + 4

    if (role == 3) { // ---- warp D: the draws ----------------------------------------------------------------------------------
        // A draw hashes digest || counter and leaves the digest alone (channel.simf:36-44), and a mix hashes digest || proof data and restarts the
        // counter (:154-172): the MIXES are the transcript's one dependent chain (32 of its 46 compressions), every draw hangs off the digest of the mix
        // before it.  This warp takes those digests from warp S in order and runs the draws beside the chain: cp_alpha, oods_t, deep_alpha, one
        // folding coefficient per FRI layer, the queries.  A draw is one block, digest || n_sent || padding, compressed here with its own schedule.
        uint32_t retries = 0, dbits = 0;
        const uint32_t query_mask = shl32(G & 0xff, 1u) - 1u;
#pragma unroll 1
        for (uint32_t k = 0; k < L + 5; k++) {
            mbar_wait(&s_mbar[k], 0);
            uint32_t d[8];
#pragma unroll
            for (int j = 0; j < 8; j++) d[j] = s_dq[k][j][lane];
            uint32_t n_sent = 0;
            if (k == L + 4) { // fri_generate_queries fri/queries.simf:30-43: 8 queries per draw, masked to the LDE domain
#pragma unroll 1
                for (uint32_t q0 = 0; q0 < Q; q0 += 8) {
                    uint32_t w[16], h[8];
#pragma unroll
                    for (int j = 0; j < 8; j++) w[j] = d[j];
                    w[8] = n_sent; w[9] = 0x80000000u;
#pragma unroll
                    for (int j = 10; j < 15; j++) w[j] = 0;
                    w[15] = 9u * 32u;
                    sha_iv(h);
                    sha_compress_rolled<K1_ADDMODE>(h, w, A);
                    n_sent = n_sent + 1u;
                    for (uint32_t j = 0; live && j < 8 && q0 + j < Q; j++) {
                        ctx[CX::QUERIES + q0 + j] = h[j] & query_mask;
                        if (tr) tr->queries[q0 + j] = h[j] & query_mask;
                    }
                }
                break;
            }
            // channel_draw_qm31 (channel.simf:115-141): retry (<= 256 draws) until the first four words are < 2p.  The step is uniform over the warp:
            // a settled proof hashes along (its hash is not looked at) while another repeats its draw.
            bool settled = false;
            uint32_t tries = 0;
            QM31 felt = qm31_zero();
#pragma unroll 1
            do {
                uint32_t w[16], h[8];
#pragma unroll
                for (int j = 0; j < 8; j++) w[j] = d[j];
                w[8] = n_sent; w[9] = 0x80000000u;
#pragma unroll
                for (int j = 10; j < 15; j++) w[j] = 0;
                w[15] = 9u * 32u;
                sha_iv(h);
                sha_compress_rolled<K1_ADDMODE>(h, w, A);
                if (!settled) {
                    const bool ok = h[0] < 4294967294u && h[1] < 4294967294u && h[2] < 4294967294u && h[3] < 4294967294u;
                    tries++;
                    if (ok || tries == 256) {
                        retries += tries - 1u;
                        settled = true;
                        if (!ok) dbits |= SSYM_ST_DRAW_EXHAUSTED;
                        felt = qm31(m31_reduce(h[0]), m31_reduce(h[1]), m31_reduce(h[2]), m31_reduce(h[3]));
                    }
                    n_sent = n_sent + 1u; // channel_draw_u256 channel.simf:36-44
                }
            } while (!__all_sync(0xffffffffu, settled));
            if (k == 0) {
                if (live) qm31_store4(ctx + CX::CP_ALPHA, felt);
                if (tr) qm31_store(tr->cp_alpha, felt);
                s_felt[4][lane] = felt.r.a; s_felt[5][lane] = felt.r.b; s_felt[6][lane] = felt.i.a; s_felt[7][lane] = felt.i.b;
            } else if (k == 1) { // channel.simf:143-144
                s_felt[0][lane] = felt.r.a; s_felt[1][lane] = felt.r.b; s_felt[2][lane] = felt.i.a; s_felt[3][lane] = felt.i.b;
            } else if (k == 2) {
                if (live) qm31_store4(ctx + CX::DEEP_ALPHA, felt);
                if (tr) qm31_store(tr->deep_alpha, felt);
                s_felt[8][lane] = felt.r.a; s_felt[9][lane] = felt.r.b; s_felt[10][lane] = felt.i.a; s_felt[11][lane] = felt.i.b;
                __threadfence_block();
                named_bar_arrive(K1_BAR_F_GO, 64); // warp F starts on the scalars
            } else {
                if (live) qm31_store4(ctx + CX::FRI_ALPHA + 4 * (k - 3), felt);
                if (tr) qm31_store(tr->fri_alpha[k - 3], felt);
            }
        }
        uint32_t used = Q;
        if (live && (p.cfg.mode & SSYM_MODE_QUERY_DEDUP)) { // include/ssym.h: sort the drawn queries, keep the distinct ones in slots [0, U), zero the rest
            uint32_t *qs = ctx + CX::QUERIES;
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
            for (uint32_t a = used; a < Q; a++) qs[a] = 0;
            if (tr)
                for (uint32_t a = 0; a < Q; a++) tr->queries[a] = qs[a];
        }
        if (live) {
            ctx[CX::N_USED] = used;
            if (tr) { tr->n_queries_used = used; tr->draw_retries = retries; }
        }
        s_dbits[lane] = dbits;
        __threadfence_block();
        named_bar_arrive(K1_BAR_D_DONE, 64);
        return;
    }

    // ---- warp S: the mixes (the dependent chain) + their message schedule ------------------------------------------------------
    const uint32_t NCOL = SSYM_STWO_COLUMNS(&p.cfg) + SSYM_NUM_CP_PARTITIONS;
    uint32_t status = 0, d[8]; // channel_init channel.simf:31-33
#pragma unroll
    for (int j = 0; j < 8; j++) d[j] = 0;
    // mix m: 0 constant root, 1 trace root, 2 composition root (evals/commit.simf:20-35), 3 OODS samples (deep/oods.simf:44-64), 4 .. 4 + L the FRI
    // roots, 5 + L the last-layer coefficient (fri/commit.simf:72-85), 6 + L the nonce (pow.simf:22-35)
#pragma unroll 1
    for (uint32_t m = 0; m < L + 7; m++) {
        uint32_t off, n = 8;
        if (m < 3) off = lo.off_commit + 8 * m;
        else if (m == 3) { off = lo.off_oods_trace; n = 4 * NCOL; }
        else if (m < 5 + L) off = m == 4 ? lo.off_fri_first_root : lo.off_fri_inner_root + 8 * (m - 5);
        else if (m == 5 + L) { off = lo.off_last_coeff; n = 4; }
        else { off = lo.off_pow_nonce; n = 2; }
        const uint32_t nwords = 8 + n, nblocks = (nwords + 3 + 15) >> 4;
        const bool pad64 = nwords == 16; // the 12 digest-sized mixes: the second block is the constant padding block of a 64-byte message
        if (lane == 0) { s_ctl[0] = nblocks; s_ctl[1] = pad64 ? 1u : 0u; s_ctl[2] = 0u; }
        k1_bar_rs(); // control ready
#pragma unroll 1
        for (uint32_t b = 0; b < nblocks; b++) {
            if (pad64 && b == 1) continue;
            uint32_t w[16];
#pragma unroll
            for (int j = 0; j < 16; j++) {
                const uint32_t mw = b * 16 + j;
                uint32_t v = 0;
                if (mw < 8) v = d[j & 7]; // only in block 0, where mw == j
                else if (mw < nwords) v = s_hw[off + (mw - 8)][lane];
                else if (mw == nwords) v = 0x80000000u;
                else if (mw == nblocks * 16 - 1) v = nwords * 32u;
                w[j] = v;
            }
#pragma unroll 1
            for (uint32_t grp = 0; grp < 4; grp++) {
                if (grp) {
#pragma unroll
                    for (int t = 0; t < 16; t++) w[t] = A.sch4(w[t], A.ssig0(w[(t + 1) & 15]), w[(t + 9) & 15], A.ssig1(w[(t + 14) & 15]));
                }
                uint32_t(*kw)[32] = s_kw[grp & 1];
#pragma unroll
                for (int t4 = 0; t4 < 4; t4++) {
                    const uint4 k4 = c_sha_k4.v[grp * 4 + t4];
                    kw[4 * t4 + 0][lane] = A.t1(w[4 * t4 + 0], k4.x);
                    kw[4 * t4 + 1][lane] = A.t1(w[4 * t4 + 1], k4.y);
                    kw[4 * t4 + 2][lane] = A.t1(w[4 * t4 + 2], k4.z);
                    kw[4 * t4 + 3][lane] = A.t1(w[4 * t4 + 3], k4.w);
                }
                k1_bar_rs(); // group produced
            }
        }
        k1_bar_rs(); // digest ready
#pragma unroll
        for (int k = 0; k < 8; k++) d[k] = s_dig[k][lane]; // channel_mix_*: the digest moves, the counter restarts
        // which draw group hangs off this digest: trace root -> cp_alpha, composition root -> oods_t, OODS samples -> deep_alpha, FRI root l -> its
        // folding coefficient, nonce -> the queries
        const uint32_t slot = m == 0 || m == 5 + L ? 0xffffffffu : m <= 4 + L ? m - 1 : L + 4;
        if (slot != 0xffffffffu) {
#pragma unroll
            for (int k = 0; k < 8; k++) s_dq[slot][k][lane] = d[k];
            mbar_arrive(&s_mbar[slot]); // release: the digest is visible to the waiter
        }
        if (tr) {
            uint32_t *dst = m == 2 ? tr->digest_commit : m == 3 ? tr->digest_oods : m == 5 + L ? tr->digest_fri : m == 6 + L ? tr->digest_pow : nullptr;
            if (dst)
                for (int j = 0; j < 8; j++) dst[j] = d[j];
        }
        if (m == 6 + L) { // check_proof_of_work pow.simf:22-35
            const uint64_t value = ((uint64_t)__byte_perm(d[7], 0, 0x0123) << 32) | __byte_perm(d[6], 0, 0x0123);
            if (!(value < p.cfg.pow_target)) status |= SSYM_ST_POW_FAIL;
            if (tr) {
                tr->pow_value[0] = (uint32_t)(value >> 32);
                tr->pow_value[1] = (uint32_t)value;
            }
        }
    }
    if (lane == 0) s_ctl[2] = 1u;
    k1_bar_rs(); // releases warp R
    if (SSYM_MODE_SEMANTICS(p.cfg.mode) == SSYM_MODE_REF_LITERAL && ((G - (L + 1u)) & 0xff) != 0) status |= SSYM_ST_FINAL_LOG; // fri/verify.simf:127
    named_bar_sync(K1_BAR_F_DONE, 64); // the scalars are done
    named_bar_sync(K1_BAR_D_DONE, 64); // the draws are done
    if (live) p.status[idx] = status | s_fbits[lane] | s_dbits[lane];
}

// ------------------------------------------------------------------------------------------
// K2: the per-query field arithmetic of verify_proof, one warp per proof (the per-proof scalars are K1's):
//   phase B  lane k < C + 16: DEEP line coefficients of column k with alpha^(k+1) (deep/quotients.simf:25-35); the
//            coefficients depend only on the proof, not on the query, so they are computed once, in parallel
//   phase C  lane q < Q: fri_answer of query q (fri/answers.simf:97-129) and its 1+L folds (fri/layers.simf:51-78,
//            fri/folding.simf:15-41); the Merkle halves of those functions are K3
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ QM31 qm31_shfl(QM31 v, int src) {
    return qm31(__shfl_sync(0xffffffffu, v.r.a, src), __shfl_sync(0xffffffffu, v.r.b, src), __shfl_sync(0xffffffffu, v.i.a, src),
                __shfl_sync(0xffffffffu, v.i.b, src));
}
// Sum of canonical field elements over the warp (exact field addition is associative, so the butterfly order gives
// the same canonical value as the reference's left-to-right fold, fri/answers.simf:52).
__device__ __forceinline__ QM31 qm31_warp_sum(QM31 v) {
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) {
        QM31 o = qm31(__shfl_xor_sync(0xffffffffu, v.r.a, off), __shfl_xor_sync(0xffffffffu, v.r.b, off),
                      __shfl_xor_sync(0xffffffffu, v.i.a, off), __shfl_xor_sync(0xffffffffu, v.i.b, off));
        v = qm31_add(v, o);
    }
    return v;
}

// sum_k b_k * v_k - (R.y * sum_a + sum_c): the batch's quotient numerator accumulator (= the fold of
// quotient_numerator_aggregate, fri/answers.simf:40-58, regrouped; all operands are exact field values and only the
// b_k * v_k products see raw witness words, exactly as in deep_quotient_nominator, deep/quotients.simf:38-44)
__device__ __forceinline__ QM31 batch_numerator(const uint32_t *bcoef, const uint32_t *vals, int n, QM31 sum_a, QM31 sum_c, M31 ry) {
    QM31 s = qm31_zero();
#pragma unroll 4
    for (int k = 0; k < n; k++) s = qm31_add(s, qm31_mul_m31(qm31_load4(bcoef + 4 * k), vals[k]));
    return qm31_sub(s, qm31_add(qm31_mul_m31(sum_a, ry), sum_c));
}

#define K2_WARPS 4
template <int C> // NUM_COLUMNS (config.simf:14): the DEEP quotient runs over NCOL = C + 16 <= 32 columns, one lane each
__global__ void __launch_bounds__(32 * K2_WARPS, 8) stwo_query_kernel(StwoParams p) { // <= 64 registers: a CTA fits the slot a Merkle CTA frees
    constexpr int NCOL = C + SSYM_NUM_CP_PARTITIONS;
    static_assert(NCOL <= 32 && C % 4 == 0, "one lane per column; the per-query values are read as uint4");
    __shared__ __align__(16) uint32_t s_b[K2_WARPS][NCOL * 4];
    const uint32_t Q = p.cfg.n_queries, L = p.cfg.n_fri_layers;
    const uint32_t wib = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t i = blockIdx.x * K2_WARPS + wib;
    if (i >= p.n) return; // warp-uniform
    const ssym_stwo_layout_t &lo = p.lo;
    const uint32_t *pk = p.packed + (size_t)i * lo.stride_words;
    const uint32_t *ctx = p.ctx + (size_t)i * CX::WORDS;
    ssym_stwo_trace_t *tr = p.trace ? p.trace + i : nullptr;
    const bool literal = SSYM_MODE_SEMANTICS(p.cfg.mode) == SSYM_MODE_REF_LITERAL;
    uint32_t status = 0;

    // the per-proof scalars come from K1: the OODS point P, the sample point of batch A, the powers of the DEEP coefficient
    const QM31 px = qm31_load4(ctx + CX::PX), py = qm31_load4(ctx + CX::PY);
    QM31 p2x = qm31_load4(ctx + CX::P2X), p2y = qm31_load4(ctx + CX::P2Y);
    if (SSYM_MODE_SEMANTICS(p.ctx_mode) != SSYM_MODE_SEMANTICS(p.cfg.mode)) { // the transcript ran under the other semantics: stwo_scalars' other branch
        const QM31 pxy = qm31_mul_nl(px, py);
        p2x = literal ? px : qm31_point_dbl_x(px);
        p2y = literal ? py : qm31_add(pxy, pxy);
    }

    // ---- phase B: lane k computes the line coefficients of column k (aggregation order) with alpha^(k+1) ----
    QM31 sum_a_A, sum_c_A, sum_a_B, sum_c_B;
    const QM31 batch_coeff = qm31_load4(ctx + CX::ALPHA_POW + 4 * NCOL); // alpha^(NCOL + 1)
    {
        const uint32_t k = lane < NCOL ? lane : NCOL - 1;
        const QM31 alpha_pow = qm31_load4(ctx + CX::ALPHA_POW + 4 * k);
        // literal: columns = C trace then 16 CP, all at P (the trace samples and the CP samples are contiguous in the header).
        // prover-consistent: 16 CP at 2P, then C trace at P.
        const bool in_A = literal || k < 16;
        const uint32_t *sv = literal ? pk + lo.off_oods_trace + 4 * k : (k < 16 ? pk + lo.off_oods_cp + 4 * k : pk + lo.off_oods_trace + 4 * (k - 16));
        const LineCoeffs lc = interpolant_coefficients(in_A ? p2y : py, qm31_load4(sv), alpha_pow);
        if (lane < NCOL) qm31_store4(&s_b[wib][4 * lane], lc.b);
        const QM31 zero = qm31_zero();
        sum_a_A = qm31_warp_sum(lane < NCOL && in_A ? lc.a : zero);
        sum_c_A = qm31_warp_sum(lane < NCOL && in_A ? lc.c : zero);
        sum_a_B = qm31_warp_sum(lane < NCOL && !in_A ? lc.a : zero);
        sum_c_B = qm31_warp_sum(lane < NCOL && !in_A ? lc.c : zero);
    }
    __syncwarp();

    // ---- phase C: lane q = query q ----
    if (lane < ctx[CX::N_USED]) { // U = Q unless SSYM_MODE_QUERY_DEDUP
        const uint32_t q = lane;
        const uint32_t query = ctx[CX::QUERIES + q];
        const uint2 rp = p.tab.point[query]; // domain point of the query, fri/answers.simf:108-110
        const M31Point R = m31_point(rp.x, rp.y);
        uint32_t vals[NCOL];
        {
            const uint4 *v4 = reinterpret_cast<const uint4 *>(pk + lo.off_qvals + NCOL * q); // 4 * NCOL-byte records: 16-byte aligned
#pragma unroll
            for (int k = 0; k < NCOL / 4; k++) {
                const uint4 v = __ldg(v4 + k);
                vals[4 * k] = v.x; vals[4 * k + 1] = v.y; vals[4 * k + 2] = v.z; vals[4 * k + 3] = v.w;
            }
        }
        QM31 eval;
        bool inv_fail = false;
        if (literal) {
            const CM31 den_inv = denominator_inverse(px, py, R, inv_fail);
            const QM31 acc = batch_numerator(s_b[wib], vals, NCOL, sum_a_A, sum_c_A, R.y);
            eval = qm31_mul(qm31_mul_cm31(acc, den_inv), batch_coeff); // fri/answers.simf:126
        } else {
            const CM31 den_a = denominator_inverse(p2x, p2y, R, inv_fail);
            const CM31 den_b = denominator_inverse(px, py, R, inv_fail);
            const QM31 num_a = batch_numerator(s_b[wib], vals + C, 16, sum_a_A, sum_c_A, R.y);
            const QM31 num_b = batch_numerator(s_b[wib] + 64, vals, C, sum_a_B, sum_c_B, R.y);
            eval = qm31_add(qm31_mul_cm31(num_a, den_a), qm31_mul_cm31(num_b, den_b));
        }
        if (inv_fail) {
            status |= SSYM_ST_ANSWER_INV_ZERO;
            if (tr) atomicOr(&tr->mask_answer_inv, 1u << q);
        }
        if (tr) qm31_store(tr->fri_answer[q], eval);

        uint32_t *ev_out = p.fri_evals + (size_t)i * (L + 1) * Q * 4;
        uint32_t fq = query;
#pragma unroll 1
        for (uint32_t l = 0; l <= L; l++) { // fri_verify_query fri/layers.simf:51-69 (without verify_decommitment)
            qm31_store4(ev_out + (l * Q + q) * 4, eval);
            const QM31 witness = qm31_load4(pk + lo.off_fri_wit + (l * Q + q) * 4);
            const bool even = (fq & 1u) == 0; // adjacent_leaves fri/layers.simf:29-37
            const QM31 e0 = even ? eval : witness, e1 = even ? witness : eval;
            const M31 inv = p.tab.fold_inv[p.tab.fold_off[l] + (fq >> 1)];
            if (inv == 0) { // m31_inv(0): fri/folding.simf:20,34 assert
                status |= SSYM_ST_FOLD_INV_ZERO;
                if (tr) atomicOr(&tr->mask_fold_inv[l], 1u << q);
            }
            eval = qm31_fold(e0, e1, inv, qm31_load4(ctx + CX::FRI_ALPHA + 4 * l)); // circle_fold / line_fold, fri/folding.simf:15-41
            if (tr) qm31_store(tr->folded[l][q], eval);
            fq >>= 1; // divide_32(position, 2)
        }
        // fri_verify_last_layer fri/layers.simf:73-78
        if (literal && fq != 0) {
            status |= SSYM_ST_LAST_QUERY;
            if (tr) atomicOr(&tr->mask_last_query, 1u << q);
        }
        if (!qm31_eq(eval, qm31_load4(pk + lo.off_last_coeff))) {
            status |= SSYM_ST_LAST_EVAL;
            if (tr) atomicOr(&tr->mask_last_eval, 1u << q);
        }
    }
    status = __reduce_or_sync(0xffffffffu, status);
    if (lane == 0 && status && p.status) atomicOr(&p.status[i], status); // (no status: the records' pass of launch_stwo_verify_cross)
}

// ------------------------------------------------------------------------------------------
// K3: every Merkle decommitment of the proof as an independent hash chain
// (merkle.simf:22-44 under evals/verify.simf:50-68 and fri/layers.simf:40-48).
// A warp's 32 lanes are 32 consecutive (proof, query) pairs of ONE chain type, so the lanes of a
// warp run chains of identical length: no divergence.  Chain types are scheduled longest first.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void load_digest(const uint32_t *src, uint32_t (&d)[8]) { // 32-byte aligned
    const uint4 a = __ldg(reinterpret_cast<const uint4 *>(src));
    const uint4 b = __ldg(reinterpret_cast<const uint4 *>(src) + 1);
    d[0] = a.x; d[1] = a.y; d[2] = a.z; d[3] = a.w;
    d[4] = b.x; d[5] = b.y; d[6] = b.z; d[7] = b.w;
}

// Every chain is run by ONE hashing loop, so the kernel contains a single copy of the compression code (the fully
// unrolled pair hash alone is 37 KB of SASS; several inlined copies thrash the instruction cache).  A chain is a few
// "pre" steps that produce its first node, then one pair hash per sibling:
//   trace : pre 0 = SHA-256(4 words)                       hash_node_m31_trace   hasher.simf:85-90
//   cp    : pre 0 = SHA-256(16 words)                      hash_node_m31_cp      hasher.simf:93-97
//   fri   : pre 0 = SHA-256(e0), pre 1 = SHA-256(e1), pre 2 = sha256_pair        fri/layers.simf:40-48
// KMODE 0: the kernel of every packed batch.  KMODE 1 / 2 (compact transport form, version 3; StwoParams::derive): the siblings a record left out
// are the nodes of other queries' paths — the 16 (Q) chains of one tree of one proof sit in one warp and step through the levels together, so a
// derived sibling is a warp shuffle away; KMODE 2 finds, for the packer, which siblings could be left out.  Needs 32 % Q == 0.
template <int ADDMODE, bool ROLLED, int KMODE = 0>
__global__ void __launch_bounds__(SSYM_MERKLE_THREADS, SSYM_MERKLE_MINB) stwo_merkle_kernel(StwoParams p, uint32_t groups_per_type, ShaMul mul) {
    const ShaAdd<ADDMODE> A(mul);
    const uint32_t Q = p.cfg.n_queries, L = p.cfg.n_fri_layers, G = p.cfg.lde_log;
    const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    // warp -> (chain type rank, group); ranks are ordered longest chain first: 0 = CP, 1 = FRI layer 0, 2 = trace,
    // 3.. = FRI layers 1..L
    uint32_t rank = warp / groups_per_type;
    const uint32_t group = warp % groups_per_type;
    if (KMODE == 1 && p.fri_only) rank = rank == 0 ? 1u : rank + 2u; // FRI layer 0, then layers 1..L
    if (rank >= L + 3) return;
    const uint32_t item = group * 32 + lane;
    const bool in_batch = item < p.n * Q;
    const uint32_t i = in_batch ? item / Q : 0, q = in_batch ? item % Q : 0;
    const uint32_t used = p.ctx[(size_t)i * CX::WORDS + CX::N_USED];
    const bool active = in_batch && q < used; // slots >= U are not looked at (SSYM_MODE_QUERY_DEDUP)
    const ssym_stwo_layout_t &lo = p.lo;
    const uint32_t *pk = p.packed + (size_t)i * lo.stride_words;
    const uint32_t query = p.ctx[(size_t)i * CX::WORDS + CX::QUERIES + q];

    const uint32_t *sib, *root, *msg; // msg: where the chain's leaf data lives
    uint32_t n_sib, n_pre, path, fail_bit, layer = 0;
    bool fri_even = true;
    int kind; // 0 trace, 1 cp, 2 fri
    const uint32_t C = SSYM_STWO_COLUMNS(&p.cfg); // trace leaf = C words (hasher.simf:85-90): 4 / 8 in one block, 16 = a 64-byte message like the CP leaf
    if (rank == 0 || rank == 2) {
        msg = pk + lo.off_qvals + (C + SSYM_NUM_CP_PARTITIONS) * q + (rank == 0 ? C : 0);
        if (rank == 2) {
            kind = 0;
            sib = pk + lo.off_trace_sib + q * G * 8;
            root = pk + lo.off_commit + 8;
            fail_bit = SSYM_ST_TRACE_MERKLE;
        } else {
            kind = 1;
            sib = pk + lo.off_cp_sib + q * G * 8;
            root = pk + lo.off_commit + 16;
            fail_bit = SSYM_ST_CP_MERKLE;
        }
        n_pre = 1;
        n_sib = G;
        path = query + shl32(G & 0xff, 1u); // evals/verify.simf:54,64
    } else {
        kind = 2;
        layer = rank == 1 ? 0 : rank - 2;
        const uint32_t fq = query >> layer;
        msg = pk + lo.off_fri_wit + (layer * Q + q) * 4; // the witness; the evaluation comes from K2's scratch
        fri_even = (fq & 1u) == 0;                        // adjacent_leaves fri/layers.simf:29-37
        n_pre = 3;
        n_sib = G - 1 - layer;
        sib = pk + lo.off_fri_sib[layer] + q * n_sib * 8;
        root = layer == 0 ? pk + lo.off_fri_first_root : pk + lo.off_fri_inner_root + 8 * (layer - 1);
        fail_bit = SSYM_ST_FRI_MERKLE(layer);
        path = ((fq & ~1u) + shl32((G - layer) & 0xff, 1u)) >> 1;
    }
    const uint32_t *evp = p.fri_evals + ((size_t)i * (L + 1) * Q + layer * Q + q) * 4;
    // compact form: this chain's bytes in the proof's derive table (slot = tree's first slot + q * n_sib + level)
    uint8_t *dv = nullptr;
    if (KMODE != 0) {
        const uint32_t tree_first = kind == 0 ? 0u : kind == 1 ? Q * G : Q * (2u * G + layer * (G - 1u) - layer * (layer - 1u) / 2u);
        dv = p.derive + (size_t)i * p.derive_stride + tree_first + q * n_sib;
    }
    const uint32_t grp_base = lane & ~(Q - 1u); // first lane of this proof's tree in the warp (32 % Q == 0)

    uint32_t cur[8], nxt[8]; // nxt: next sibling (prefetched one level ahead); during the FRI pre steps: the first leaf hash
#pragma unroll
    for (int k = 0; k < 8; k++) cur[k] = nxt[k] = 0;
    const uint32_t total = n_pre + n_sib;
#pragma unroll 1
    for (uint32_t step = 0; step < total; step++) {
        uint32_t w[16];
        bool two_blocks = true; // 64-byte message: data block + the constant padding block
        if (step < n_pre) {
            if (kind == 1 || (kind == 0 && C == 16)) { // a 64-byte leaf: 16 composition values, or 16 trace columns
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    const uint4 v = __ldg(reinterpret_cast<const uint4 *>(msg) + k);
                    w[4 * k] = v.x; w[4 * k + 1] = v.y; w[4 * k + 2] = v.z; w[4 * k + 3] = v.w;
                }
            } else if (kind == 2 && step == 2) {
#pragma unroll
                for (int k = 0; k < 8; k++) { w[k] = nxt[k]; w[8 + k] = cur[k]; }
            } else { // 16-byte message in one block: trace leaf, or the left (step 0) / right (step 1) FRI leaf
                const bool take_eval = active && kind == 2 && ((step == 0) == fri_even); // (an idle lane has no evaluation: K2 wrote none for it)
                const uint4 v = take_eval ? *reinterpret_cast<const uint4 *>(evp) : __ldg(reinterpret_cast<const uint4 *>(msg));
                w[0] = v.x; w[1] = v.y; w[2] = v.z; w[3] = v.w;
                uint4 v2 = make_uint4(0x80000000u, 0u, 0u, 0u); // FIPS 180-4 padding right after a 16-byte message ...
                const bool wide = kind == 0 && C == 8;          // ... or after the 32 bytes of an 8-column trace leaf
                if (wide) v2 = __ldg(reinterpret_cast<const uint4 *>(msg) + 1);
                w[4] = v2.x; w[5] = v2.y; w[6] = v2.z; w[7] = v2.w;
                w[8] = wide ? 0x80000000u : 0u;
#pragma unroll
                for (int k = 9; k < 15; k++) w[k] = 0;
                w[15] = wide ? 256u : 128u;
                two_blocks = false;
                if (kind == 2 && step == 1) {
#pragma unroll
                    for (int k = 0; k < 8; k++) nxt[k] = cur[k];
                }
            }
            if (step + 1 == n_pre && n_sib) load_digest(sib, nxt);
        } else { // merkle_compute_step merkle.simf:22-30
            const uint32_t lvl = step - n_pre;
            const bool cur_left = (path & 1u) == 0; // divides_32(2, path): sha256_pair(cur, sib) else (sib, cur)
            if (KMODE == 1) { // a sibling the record left out = the node another query's path has reached at this level
                const uint32_t pb = active && ((p.derive_kinds >> (kind == 2 ? 1 : 0)) & 1u) ? dv[lvl] : 0xffu;
                const uint32_t src = grp_base + (pb < Q ? pb : 0u);
                uint32_t got[8];
#pragma unroll
                for (int k = 0; k < 8; k++) got[k] = __shfl_sync(0xffffffffu, cur[k], src);
                if (pb < Q) {
#pragma unroll
                    for (int k = 0; k < 8; k++) nxt[k] = got[k];
                    uint4 *dst = reinterpret_cast<uint4 *>(p.packed_rw + (sib - p.packed) + 8 * lvl); // the packed record becomes complete
                    dst[0] = make_uint4(nxt[0], nxt[1], nxt[2], nxt[3]);
                    dst[1] = make_uint4(nxt[4], nxt[5], nxt[6], nxt[7]);
                }
            }
            if (KMODE == 2) { // which query's node IS this sibling?  (lowest such query; 0xff = none: the digest has to be shipped)
                uint32_t found = 0xffu;
#pragma unroll 1
                for (uint32_t j = 0; j < Q; j++) {
                    const uint32_t c0 = __shfl_sync(0xffffffffu, cur[0], grp_base + j);
                    if (__any_sync(0xffffffffu, c0 == nxt[0] && j != q && j < used && found == 0xffu)) {
                        bool same = c0 == nxt[0];
#pragma unroll
                        for (int k = 1; k < 8; k++) same = (__shfl_sync(0xffffffffu, cur[k], grp_base + j) == nxt[k]) && same;
                        if (same && j != q && j < used && found == 0xffu) found = j;
                    }
                }
                if (active) dv[lvl] = (uint8_t)found;
            }
#pragma unroll
            for (int k = 0; k < 8; k++) {
                w[k] = cur_left ? cur[k] : nxt[k];
                w[8 + k] = cur_left ? nxt[k] : cur[k];
            }
            if (lvl + 1 < n_sib) load_digest(sib + 8 * (lvl + 1), nxt);
            path >>= 1;
        }
        sha_iv(cur);
        if (ROLLED) {
            sha_compress_rolled<ADDMODE>(cur, w, A);
            if (two_blocks) sha_compress_pad64_rolled<ADDMODE>(cur, A);
        } else {
            sha_compress<ADDMODE>(cur, w, A);
            if (two_blocks) sha_compress_pad64<ADDMODE>(cur, A);
        }
    }
    uint32_t r[8];
    load_digest(root, r);
    bool ok = path == 1u; // merkle.simf:42
#pragma unroll
    for (int k = 0; k < 8; k++) ok = ok && (cur[k] == r[k]); // merkle.simf:43
    if (active && !ok && p.status) atomicOr(&p.status[i], fail_bit);
    if (active && p.trace) {
        ssym_stwo_trace_t *tr = p.trace + i;
        uint32_t *dst = kind == 0 ? tr->trace_root[q] : kind == 1 ? tr->cp_root[q] : tr->fri_root[layer][q];
#pragma unroll
        for (int k = 0; k < 8; k++) dst[k] = cur[k];
        if (!ok) atomicOr(kind == 0 ? &tr->mask_trace : kind == 1 ? &tr->mask_cp : &tr->mask_fri[layer], 1u << q);
    }
}

// ------------------------------------------------------------------------------------------
// K3 with shared nodes (StwoDedup, stwo_kernels.cuh): plan -> hash the distinct nodes (round 1) -> check the followers ->
// hash what did not match (round 2; nothing for an honest proof) -> resolve.
// ------------------------------------------------------------------------------------------
#define DD_NONE 0xffu
#define DD_FOLLOWER 0x80u
#define DD_INVALID 0x40u
#define DD_SIBDIFF 0x20u
#define DD_SAMELEAF 0x1000u // follower whose leaf position IS its leader's (h = 0): nothing to hash in round 1
#define DD_UNUSED 0x2000u   // query slot >= U under SSYM_MODE_QUERY_DEDUP: no chain

// Appends chain `c` to bin `bin` of round `round`: counters privatised in shared memory, one global atomic per (CTA, bin).
// Every thread of the CTA calls this (want = false for threads with nothing to append; up to two appends per thread).
template <int MAXAPP>
__device__ __forceinline__ void dd_append(const StwoDedup &dd, uint32_t round, const uint32_t (&bin)[MAXAPP], const uint32_t (&chain)[MAXAPP], uint32_t n_app,
                                          uint32_t *s_cnt, uint32_t *s_base) {
    for (uint32_t t = threadIdx.x; t < STWO_DEDUP_MAX_BINS; t += blockDim.x) s_cnt[t] = 0;
    __syncthreads();
    uint32_t slot[MAXAPP];
#pragma unroll
    for (int k = 0; k < MAXAPP; k++)
        if ((uint32_t)k < n_app) slot[k] = atomicAdd(&s_cnt[bin[k]], 1u);
    __syncthreads();
    for (uint32_t t = threadIdx.x; t < STWO_DEDUP_MAX_BINS; t += blockDim.x)
        if (s_cnt[t]) s_base[t] = atomicAdd(&dd.bin_count[round * STWO_DEDUP_MAX_BINS + t], s_cnt[t]);
    __syncthreads();
#pragma unroll
    for (int k = 0; k < MAXAPP; k++)
        if ((uint32_t)k < n_app) dd.bin_list[dd.bin_base[round][bin[k]] + s_base[bin[k]] + slot[k]] = chain[k];
}

// One thread per (proof, plan, query slot): plan 0 = the trace and composition trees (leaf position = query, depth G), plan 1 + l = FRI layer
// l (leaf position = the pair index (query >> l) >> 1, depth G - 1 - l).  16 lanes per plan so that a plan's queries talk by shuffles.
__global__ void __launch_bounds__(512) stwo_plan_kernel(StwoParams p) {
    __shared__ uint32_t s_cnt[STWO_DEDUP_MAX_BINS], s_base[STWO_DEDUP_MAX_BINS];
    const uint32_t Q = p.cfg.n_queries, L = p.cfg.n_fri_layers, G = p.cfg.lde_log;
    const uint32_t plans = L + 2;
    const uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t grp = g >> 4, q = g & 15u;
    const uint32_t i = grp / plans, pl = grp % plans;
    const bool in_range = i < p.n; // uniform over the 16-lane group
    const uint32_t lane = threadIdx.x & 31u, gmask = 0xffffu << (lane & 16u);
    const uint32_t d = pl == 0 ? G : G - pl, shift = pl;
    const bool slot = in_range && q < Q;
    const bool act = slot && q < p.ctx[(size_t)i * CX::WORDS + CX::N_USED]; // slots >= U are not looked at (SSYM_MODE_QUERY_DEDUP)
    const uint32_t pos = act ? p.ctx[(size_t)i * CX::WORDS + CX::QUERIES + q] >> shift : 0u;
    uint32_t h = d, lead = DD_NONE;
    for (uint32_t r = 0; r < 15; r++) {
        const uint32_t pr = __shfl_sync(gmask, pos, r, 16);
        if (act && r < q) {
            const uint32_t x = pos ^ pr, mh = x ? 32u - __clz(x) : 0u;
            if (mh < h) { h = mh; lead = r; }
        }
    }
    // REF_LITERAL: fri_answer (fri/answers.simf:116-126) gives evaluations no prover's FRI trees were built from, so every query's FRI
    // path starts from a leaf of its own and nothing can be shared there: plan those trees per query straight away.
    if (pl >= 1 && SSYM_MODE_SEMANTICS(p.cfg.mode) == SSYM_MODE_REF_LITERAL) { h = d; lead = DD_NONE; }
    const bool follower = act && lead != DD_NONE;
    // leaders: which of my nodes the followers need (at most one follower per height >= 1: three paths cannot first meet in one node)
    uint32_t mask = 0;
    uint64_t to = 0;
    for (uint32_t r = 1; r < 16; r++) {
        const uint32_t hr = __shfl_sync(gmask, h, r, 16), lr = __shfl_sync(gmask, follower ? lead : DD_NONE, r, 16);
        if (act && lr == q && hr >= 1) { mask |= 1u << (hr - 1); to |= (uint64_t)r << (4 * (hr - 1)); }
    }
    // A follower stops one level BELOW the meeting node (lv = h - 1 levels): there its node must be the leader's sibling and its sibling the
    // leader's node, which makes the meeting node — and everything above it — the leader's.  Checkpoint bit / nibble k = the leader's node at height k.
    const bool same_leaf = follower && h == 0;
    const uint32_t lv = follower ? (h ? h - 1 : 0u) : h;
    const uint32_t word = lv | (follower ? DD_FOLLOWER : 0u) | (same_leaf ? DD_SAMELEAF : 0u) | ((lead & 15u) << 8) | (mask << 16);
    const uint32_t CH = p.dd.chains;
    uint32_t bin[2] = {0, 0}, chain[2] = {0, 0}, n_app = 0;
    if (act) {
        const uint32_t n_trees = pl == 0 ? 2u : 1u;
        for (uint32_t t = 0; t < n_trees; t++) {
            const uint32_t tree = pl == 0 ? t : pl + 1, kind = tree < 2 ? tree : 2u;
            const uint32_t c = i * CH + tree * Q + q;
            p.dd.plan[c] = word;
            p.dd.ckpt_to[c] = to;
            if (!same_leaf) { bin[n_app] = p.dd.bin_of[0][kind][lv]; chain[n_app] = c; n_app++; }
        }
    }
    if (slot && !act) { // an unused slot: no chain; the check / resolve kernels skip it
        const uint32_t n_trees = pl == 0 ? 2u : 1u;
        for (uint32_t t = 0; t < n_trees; t++) p.dd.plan[i * CH + (pl == 0 ? t : pl + 1) * Q + q] = DD_UNUSED;
    }
    dd_append<2>(p.dd, 0, bin, chain, n_app, s_cnt, s_base);
}

// One thread per task, a warp = 32 tasks of one bin = one (kind, number of steps): no divergence; bins are ordered longest first.
// Round 1: chain c from its leaf up lv levels (plan: a follower stops one level below its meeting node), storing the nodes its followers need.
// Round 2: the followers whose check failed, from their node at height lv (same-leaf followers: from their leaf) to the root with their own siblings.
template <int ADDMODE>
__global__ void __launch_bounds__(128, 8) stwo_merkle_shared_kernel(StwoParams p, uint32_t round, ShaMul mul) {
    const ShaAdd<ADDMODE> A(mul);
    const uint32_t Q = p.cfg.n_queries, L = p.cfg.n_fri_layers, G = p.cfg.lde_log;
    const uint32_t lane = threadIdx.x & 31, grid_warps = (gridDim.x * blockDim.x) >> 5;
    const uint32_t CH = p.dd.chains;
    // round 2 runs on a small grid (it has nothing to do for honest proofs) and strides over the warps of work
#pragma unroll 1
    for (uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;; warp += grid_warps) {
    uint32_t base_w = 0, cnt = 0, bin = DD_NONE;
    for (uint32_t b = 0; b < p.dd.n_bins[round]; b++) {
        cnt = p.dd.bin_count[round * STWO_DEDUP_MAX_BINS + b];
        const uint32_t nw = (cnt + 31u) >> 5;
        if (warp < base_w + nw) { bin = b; break; }
        base_w += nw;
    }
    if (bin == DD_NONE) return;
    const uint32_t off = (warp - base_w) * 32u + lane;
    if (off >= cnt) continue;
    const uint32_t cg = p.dd.bin_list[p.dd.bin_base[round][bin] + off];
    const uint32_t i = cg / CH, c = cg % CH, tree = c / Q, q = c % Q;
    const uint32_t pw = p.dd.plan[cg];
    const uint32_t lv = pw & 31u, ckmask = round == 0 ? pw >> 16 : 0u;
    const uint64_t ckto = round == 0 ? p.dd.ckpt_to[cg] : 0ull;
    const ssym_stwo_layout_t &lo = p.lo;
    const uint32_t *pk = p.packed + (size_t)i * lo.stride_words;
    const uint32_t query = p.ctx[(size_t)i * CX::WORDS + CX::QUERIES + q];

    const uint32_t *sib, *msg;
    uint32_t n_pre, path, layer = 0, depth;
    bool fri_even = true;
    int kind;
    const uint32_t C = SSYM_STWO_COLUMNS(&p.cfg);
    if (tree < 2) {
        kind = (int)tree;
        msg = pk + lo.off_qvals + (C + SSYM_NUM_CP_PARTITIONS) * q + (tree == 1 ? C : 0);
        sib = pk + (tree == 0 ? lo.off_trace_sib : lo.off_cp_sib) + q * G * 8;
        n_pre = 1;
        path = query;
        depth = G;
    } else {
        kind = 2;
        layer = tree - 2;
        const uint32_t fq = query >> layer;
        msg = pk + lo.off_fri_wit + (layer * Q + q) * 4;
        fri_even = (fq & 1u) == 0;
        n_pre = 3;
        depth = G - 1 - layer;
        sib = pk + lo.off_fri_sib[layer] + q * depth * 8;
        path = fq >> 1;
    }
    const uint32_t *evp = p.fri_evals + ((size_t)i * (L + 1) * Q + layer * Q + q) * 4;
    // levels [start, end): round 1 = [0, lv); round 2 = from the stored node at height lv (or, for a same-leaf follower, from the leaf) to the root
    const bool from_node = round != 0 && !(pw & DD_SAMELEAF);
    const uint32_t start = from_node ? lv : 0u, end = round == 0 ? lv : depth;
    if (from_node) n_pre = 0;

    uint32_t cur[8], nxt[8];
#pragma unroll
    for (int k = 0; k < 8; k++) cur[k] = nxt[k] = 0;
    if (from_node) {
        load_digest(p.dd.own + (size_t)cg * 8, cur);
        load_digest(sib + 8 * start, nxt);
        path >>= start;
    }
    const uint32_t total = n_pre + (end - start);
#pragma unroll 1
    for (uint32_t step = 0; step < total; step++) {
        uint32_t w[16];
        bool two_blocks = true;
        if (step < n_pre) {
            if (kind == 1 || (kind == 0 && C == 16)) { // a 64-byte leaf: 16 composition values, or 16 trace columns
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    const uint4 v = __ldg(reinterpret_cast<const uint4 *>(msg) + k);
                    w[4 * k] = v.x; w[4 * k + 1] = v.y; w[4 * k + 2] = v.z; w[4 * k + 3] = v.w;
                }
            } else if (kind == 2 && step == 2) {
#pragma unroll
                for (int k = 0; k < 8; k++) { w[k] = nxt[k]; w[8 + k] = cur[k]; }
            } else {
                const bool take_eval = kind == 2 && ((step == 0) == fri_even);
                const uint4 v = take_eval ? *reinterpret_cast<const uint4 *>(evp) : __ldg(reinterpret_cast<const uint4 *>(msg));
                w[0] = v.x; w[1] = v.y; w[2] = v.z; w[3] = v.w;
                uint4 v2 = make_uint4(0x80000000u, 0u, 0u, 0u); // FIPS 180-4 padding right after a 16-byte message ...
                const bool wide = kind == 0 && C == 8;          // ... or after the 32 bytes of an 8-column trace leaf
                if (wide) v2 = __ldg(reinterpret_cast<const uint4 *>(msg) + 1);
                w[4] = v2.x; w[5] = v2.y; w[6] = v2.z; w[7] = v2.w;
                w[8] = wide ? 0x80000000u : 0u;
#pragma unroll
                for (int k = 9; k < 15; k++) w[k] = 0;
                w[15] = wide ? 256u : 128u;
                two_blocks = false;
                if (kind == 2 && step == 1) {
#pragma unroll
                    for (int k = 0; k < 8; k++) nxt[k] = cur[k];
                }
            }
            if (step + 1 == n_pre && end) load_digest(sib, nxt);
        } else {
            const uint32_t lvl = start + step - n_pre;
            if ((ckmask >> lvl) & 1u) { // a follower needs this node (height lvl)
                uint32_t *dst = p.dd.ckpt + ((size_t)i * CH + tree * Q + (uint32_t)((ckto >> (4 * lvl)) & 15u)) * 8;
#pragma unroll
                for (int k = 0; k < 8; k++) dst[k] = cur[k];
            }
            const bool cur_left = (path & 1u) == 0;
#pragma unroll
            for (int k = 0; k < 8; k++) {
                w[k] = cur_left ? cur[k] : nxt[k];
                w[8 + k] = cur_left ? nxt[k] : cur[k];
            }
            if (lvl + 1 < end) load_digest(sib + 8 * (lvl + 1), nxt);
            path >>= 1;
        }
        sha_iv(cur);
        sha_compress_rolled<ADDMODE>(cur, w, A);
        if (two_blocks) sha_compress_pad64_rolled<ADDMODE>(cur, A);
    }
    uint32_t *dst = (round == 0 ? p.dd.own : p.dd.ckpt) + (size_t)cg * 8; // round 2 reuses the follower's checkpoint slot for its root
#pragma unroll
    for (int k = 0; k < 8; k++) dst[k] = cur[k];
    if (round == 0 && (pw & DD_FOLLOWER)) { // proof data only: my node (height lv) must be the leader's sibling there, and above it my siblings the leader's
        const uint4 *sq = reinterpret_cast<const uint4 *>(sib), *sr = reinterpret_cast<const uint4 *>(sib + ((int)((pw >> 8) & 15u) - (int)q) * (int)(depth * 8));
        const uint4 l0 = __ldg(sr + 2 * lv), l1 = __ldg(sr + 2 * lv + 1);
        uint32_t diff = (cur[0] ^ l0.x) | (cur[1] ^ l0.y) | (cur[2] ^ l0.z) | (cur[3] ^ l0.w) | (cur[4] ^ l1.x) | (cur[5] ^ l1.y) | (cur[6] ^ l1.z) | (cur[7] ^ l1.w);
#pragma unroll 2
        for (uint32_t k = lv + 1; k < depth; k++) {
            const uint4 a0 = __ldg(sq + 2 * k), a1 = __ldg(sq + 2 * k + 1), b0 = __ldg(sr + 2 * k), b1 = __ldg(sr + 2 * k + 1);
            diff |= (a0.x ^ b0.x) | (a0.y ^ b0.y) | (a0.z ^ b0.z) | (a0.w ^ b0.w) | (a1.x ^ b1.x) | (a1.y ^ b1.y) | (a1.z ^ b1.z) | (a1.w ^ b1.w);
        }
        if (diff) p.dd.plan[cg] = pw | DD_SIBDIFF;
    }
    }
}

__device__ __forceinline__ bool eq8(const uint32_t *a, const uint32_t *b) {
    const uint4 a0 = *reinterpret_cast<const uint4 *>(a), a1 = *(reinterpret_cast<const uint4 *>(a) + 1);
    const uint4 b0 = *reinterpret_cast<const uint4 *>(b), b1 = *(reinterpret_cast<const uint4 *>(b) + 1);
    return a0.x == b0.x && a0.y == b0.y && a0.z == b0.z && a0.w == b0.w && a1.x == b1.x && a1.y == b1.y && a1.z == b1.z && a1.w == b1.w;
}

// One thread per chain: a follower keeps its leader's result only if, one level below the meeting node, its node is the leader's sibling and its
// sibling the leader's node (same-leaf followers: its leaf data is the leader's), and every sibling above is bit-identical to the leader's;
// otherwise (corrupted proofs only) it is queued for round 2.
__global__ void __launch_bounds__(256) stwo_check_kernel(StwoParams p) {
    const uint32_t Q = p.cfg.n_queries, L = p.cfg.n_fri_layers, G = p.cfg.lde_log;
    const uint32_t CH = p.dd.chains;
    const uint32_t cg = blockIdx.x * blockDim.x + threadIdx.x;
    if (cg >= p.n * CH) return;
    const uint32_t pw = p.dd.plan[cg];
    if (!(pw & DD_FOLLOWER)) return;
    const uint32_t i = cg / CH, c = cg % CH, tree = c / Q, q = c % Q;
    const uint32_t lv = pw & 31u, lead = (pw >> 8) & 15u;
    const bool same_leaf = (pw & DD_SAMELEAF) != 0;
    const ssym_stwo_layout_t &lo = p.lo;
    const uint32_t *pk = p.packed + (size_t)i * lo.stride_words;
    const uint32_t kind = tree < 2 ? tree : 2u, layer = tree < 2 ? 0u : tree - 2u;
    const uint32_t d = tree < 2 ? G : G - 1 - layer;
    const uint32_t *sib0 = pk + (tree == 0 ? lo.off_trace_sib : tree == 1 ? lo.off_cp_sib : lo.off_fri_sib[layer]);
    bool valid = true;
    if (!same_leaf) { // round 1 compared its node with the leader's sibling and its upper siblings with the leader's; here: its sibling at lv = the leader's node
        valid = !(pw & DD_SIBDIFF) && eq8(sib0 + (q * d + lv) * 8, p.dd.ckpt + (size_t)cg * 8);
    } else if (kind == 2) { // same leaf pair: the two 16-byte leaves, in tree order (adjacent_leaves fri/layers.simf:29-37), must agree
        const uint32_t *qs = p.ctx + (size_t)i * CX::WORDS + CX::QUERIES;
        const uint32_t *ev = p.fri_evals + ((size_t)i * (L + 1) + layer) * Q * 4;
        const uint32_t *wq = pk + lo.off_fri_wit + (layer * Q + q) * 4, *wl = pk + lo.off_fri_wit + (layer * Q + lead) * 4;
        const bool even_q = ((qs[q] >> layer) & 1u) == 0, even_l = ((qs[lead] >> layer) & 1u) == 0;
        const uint32_t *a0 = even_q ? ev + 4 * q : wq, *a1 = even_q ? wq : ev + 4 * q;
        const uint32_t *b0 = even_l ? ev + 4 * lead : wl, *b1 = even_l ? wl : ev + 4 * lead;
        for (int j = 0; j < 4; j++) valid = valid && a0[j] == b0[j] && a1[j] == b1[j];
    } else {
        const uint32_t C = SSYM_STWO_COLUMNS(&p.cfg), QV = C + SSYM_NUM_CP_PARTITIONS;
        const uint32_t nw = kind == 0 ? C : 16u, o = kind == 0 ? 0u : C;
        for (uint32_t j = 0; j < nw; j++) valid = valid && pk[lo.off_qvals + QV * q + o + j] == pk[lo.off_qvals + QV * lead + o + j];
    }
    const uint4 *sq = reinterpret_cast<const uint4 *>(sib0 + q * d * 8), *sr = reinterpret_cast<const uint4 *>(sib0 + lead * d * 8);
    uint32_t diff = 0; // branch-free so that the loads of all levels are in flight together
#pragma unroll 4
    for (uint32_t k = 0; same_leaf && k < d; k++) { // (followers that met at the leaf have no thread in round 1)
        const uint4 a0 = __ldg(sq + 2 * k), a1 = __ldg(sq + 2 * k + 1), b0 = __ldg(sr + 2 * k), b1 = __ldg(sr + 2 * k + 1);
        diff |= (a0.x ^ b0.x) | (a0.y ^ b0.y) | (a0.z ^ b0.z) | (a0.w ^ b0.w) | (a1.x ^ b1.x) | (a1.y ^ b1.y) | (a1.z ^ b1.z) | (a1.w ^ b1.w);
    }
    valid = valid && diff == 0;
    if (!valid) {
        p.dd.plan[cg] = pw | DD_INVALID;
        const uint32_t bin = p.dd.bin_of[1][same_leaf ? kind : 3u][same_leaf ? d : d - lv];
        const uint32_t slot = atomicAdd(&p.dd.bin_count[STWO_DEDUP_MAX_BINS + bin], 1u);
        p.dd.bin_list[p.dd.bin_base[1][bin] + slot] = cg;
    }
}

// One thread per chain: the root the query's path leads to — its own (full chains), its leader's (followers that passed the check; the
// leader may itself follow another) or the one round 2 hashed — against the commitment (merkle.simf:43).  Sets the tree's status bit,
// the per-query fail mask and the trace's recomputed roots.
__global__ void __launch_bounds__(256) stwo_resolve_kernel(StwoParams p) {
    const uint32_t Q = p.cfg.n_queries;
    const uint32_t CH = p.dd.chains;
    const uint32_t cg = blockIdx.x * blockDim.x + threadIdx.x;
    if (cg >= p.n * CH) return;
    const uint32_t i = cg / CH, c = cg % CH, tree = c / Q, q = c % Q;
    const ssym_stwo_layout_t &lo = p.lo;
    const uint32_t *pk = p.packed + (size_t)i * lo.stride_words;
    const uint32_t layer = tree < 2 ? 0u : tree - 2u;
    const uint32_t *root = tree == 0 ? pk + lo.off_commit + 8 : tree == 1 ? pk + lo.off_commit + 16
                           : layer == 0 ? pk + lo.off_fri_first_root : pk + lo.off_fri_inner_root + 8 * (layer - 1);
    uint32_t src = cg, pw = p.dd.plan[src];
    if (pw & DD_UNUSED) return;
    for (uint32_t hop = 0; hop < SSYM_MAX_QUERIES && (pw & DD_FOLLOWER) && !(pw & DD_INVALID); hop++) { // leaders have lower query numbers
        src = i * CH + tree * Q + ((pw >> 8) & 15u);
        pw = p.dd.plan[src];
    }
    const uint32_t *r = ((pw & DD_FOLLOWER) ? p.dd.ckpt : p.dd.own) + (size_t)src * 8;
    const bool ok = eq8(r, root); // merkle.simf:42 (path == 1) holds by construction of the positions
    if (!ok) atomicOr(&p.status[i], tree == 0 ? SSYM_ST_TRACE_MERKLE : tree == 1 ? SSYM_ST_CP_MERKLE : SSYM_ST_FRI_MERKLE(layer));
    if (p.trace) {
        ssym_stwo_trace_t *tr = p.trace + i;
        uint32_t *dst = tree == 0 ? tr->trace_root[q] : tree == 1 ? tr->cp_root[q] : tr->fri_root[layer][q];
        for (int k = 0; k < 8; k++) dst[k] = r[k];
        if (!ok) atomicOr(tree == 0 ? &tr->mask_trace : tree == 1 ? &tr->mask_cp : &tr->mask_fri[layer], 1u << q);
    }
}

size_t stwo_dedup_layout(const ssym_stwo_config_t &cfg, size_t cap, StwoDedup &dd) {
    const uint32_t Q = cfg.n_queries, L = cfg.n_fri_layers, G = cfg.lde_log;
    dd.enabled = 0;
    dd.chains = (L + 3) * Q;
    if (G > STWO_DEDUP_MAX_DEPTH || Q > 16) return 0;
    // Bins (kind, levels), ordered by decreasing number of compressions.  Round 1: kind 0 trace (1 + 2 lv), 1 composition (2 + 2 lv), 2 FRI (4 + 2 lv),
    // lv = 0..depth.  Round 2: kind 3 = from a stored node at height lv, depth - lv levels; kinds 0..2 = whole paths of followers that met at the leaf.
    struct B { uint32_t kind, steps, work, capacity; };
    size_t total = 0;
    for (uint32_t round = 0; round < 2; round++) {
        B bins[STWO_DEDUP_MAX_BINS];
        uint32_t nb = 0;
        for (uint32_t kind = 0; kind < 4; kind++)
            for (uint32_t steps = 0; steps <= G; steps++) {
                // trees that can put a chain here
                uint32_t trees = 0;
                for (uint32_t tree = 0; tree < L + 3; tree++) {
                    const uint32_t tk = tree < 2 ? tree : 2u, d = tree < 2 ? G : G - 1 - (tree - 2);
                    if (round == 0) trees += (kind == tk && steps <= d) ? 1u : 0u;             // lv = steps
                    else if (kind == 3) trees += (steps >= 1 && steps <= d) ? 1u : 0u;      // lv = d - steps
                    else trees += (kind == tk && steps == d) ? 1u : 0u;                      // same leaf: the whole path
                }
                if (!trees || (round == 0 && kind == 3)) continue;
                const uint32_t pre = kind == 0 ? (SSYM_STWO_COLUMNS(&cfg) == 16 ? 2u : 1u) : kind == 1 ? 2u : kind == 2 ? 4u : 0u;
                if (nb == STWO_DEDUP_MAX_BINS) return 0;
                bins[nb++] = B{kind, steps, pre + 2u * steps, (uint32_t)(trees * Q * cap)};
            }
        for (uint32_t a = 1; a < nb; a++)
            for (uint32_t b = a; b > 0 && bins[b].work > bins[b - 1].work; b--) { const B t = bins[b]; bins[b] = bins[b - 1]; bins[b - 1] = t; }
        for (uint32_t k = 0; k < 4; k++)
            for (uint32_t h = 0; h <= STWO_DEDUP_MAX_DEPTH; h++) dd.bin_of[round][k][h] = DD_NONE;
        for (uint32_t b = 0; b < nb; b++) {
            dd.bin_of[round][bins[b].kind][bins[b].steps] = (uint8_t)b;
            dd.bin_base[round][b] = (uint32_t)total;
            total += bins[b].capacity;
        }
        dd.n_bins[round] = nb;
    }
    dd.enabled = 1;
    return total;
}

// ------------------------------------------------------------------------------------------
// K4: status -> accept bitmap (+ first failing assert in reference program order for the trace)
// ------------------------------------------------------------------------------------------
__device__ uint32_t first_fail_code(const ssym_stwo_config_t &cfg, const ssym_stwo_trace_t *t, uint32_t s) {
    if (!s) return 0;
    if (s & SSYM_ST_SHAPE) return 31u << 16;
    for (uint32_t b = 0; b <= 3; b++)
        if (s & (1u << b)) return b << 16;
    for (uint32_t q = 0; q < cfg.n_queries; q++) { // evals/verify.simf:71-78
        if (t->mask_trace & (1u << q)) return (4u << 16) | q;
        if (t->mask_cp & (1u << q)) return (5u << 16) | q;
    }
    for (uint32_t q = 0; q < cfg.n_queries; q++)
        if (t->mask_answer_inv & (1u << q)) return (6u << 16) | q;
    for (uint32_t l = 0; l <= cfg.n_fri_layers; l++)
        for (uint32_t q = 0; q < cfg.n_queries; q++) {
            if (t->mask_fri[l] & (1u << q)) return ((7u + l) << 16) | (l << 8) | q;
            if (t->mask_fold_inv[l] & (1u << q)) return (16u << 16) | (l << 8) | q;
        }
    if (s & SSYM_ST_FINAL_LOG) return 17u << 16;
    for (uint32_t q = 0; q < cfg.n_queries; q++) {
        if (t->mask_last_query & (1u << q)) return (18u << 16) | q;
        if (t->mask_last_eval & (1u << q)) return (19u << 16) | q;
    }
    return 30u << 16;
}

__global__ void __launch_bounds__(256) stwo_finalize_kernel(StwoParams p, uint32_t *accept_bits) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t s = i < p.n ? p.status[i] : 1u;
    const uint32_t ballot = __ballot_sync(0xffffffffu, s == 0);
    if ((threadIdx.x & 31) == 0 && i < p.n) accept_bits[i >> 5] = ballot;
    if (i < p.n && p.trace) {
        p.trace[i].status = s;
        p.trace[i].first_fail = first_fail_code(p.cfg, p.trace + i, s);
    }
}

// ------------------------------------------------------------------------------------------
// Domain tables
// ------------------------------------------------------------------------------------------
__global__ void stwo_tables_kernel(uint32_t G, uint32_t L, uint2 *point, uint32_t *fold_inv, StwoTables offs, uint32_t *zero_flag) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t n_point = 1u << G;
    if (t < n_point) {
        const M31Point pt = circle_point_index_to_m31_point(circle_position_to_point_index(G, bit_reverse_position(t, G)));
        point[t] = make_uint2(pt.x, pt.y);
        return;
    }
    uint32_t j = t - n_point;
    for (uint32_t l = 0; l <= L; l++) {
        const uint32_t cnt = 1u << (G - l - 1);
        if (j < cnt) {
            const uint32_t log = G - l, position = 2 * j;
            M31 v;
            if (l == 0) v = circle_point_index_to_m31_point(circle_position_to_point_index(log, bit_reverse_position(position, log))).y;
            else v = circle_point_index_to_m31_point(line_position_to_point_index(log, bit_reverse_position(position, log))).x;
            bool fail = false;
            fold_inv[offs.fold_off[l] + j] = m31_inv(v, fail);
            if (fail) atomicOr(zero_flag, 1u);
            return;
        }
        j -= cnt;
    }
}

void launch_stwo_tables(uint32_t G, uint32_t L, uint2 *point, uint32_t *fold_inv, const uint32_t *fold_off, uint32_t *zero_flag,
                        cudaStream_t s) {
    StwoTables offs{};
    uint32_t total = 1u << G;
    for (uint32_t l = 0; l <= L; l++) {
        offs.fold_off[l] = fold_off[l];
        total += 1u << (G - l - 1);
    }
    stwo_tables_kernel<<<(total + 127) / 128, 128, 0, s>>>(G, L, point, fold_inv, offs, zero_flag);
}

// dynamic shared memory of the transcript kernel: one [8][32]-word digest per draw group; with its ~38 KB of static arrays the CTA passes 48 KB,
// so every device that runs it opts in once (ssym_create)
static size_t k1_dyn_smem(const StwoParams &p) { return (size_t)(p.cfg.n_fri_layers + 5) * 8 * 32 * sizeof(uint32_t); }
cudaError_t stwo_kernels_init_device() {
    return cudaFuncSetAttribute(stwo_channel_ws_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (SSYM_MAX_FRI_LAYERS + 5) * 8 * 32 * (int)sizeof(uint32_t));
}

void launch_stwo_verify(const StwoParams &p, uint32_t *accept_bits, cudaStream_t s, uint64_t *launch_counter, Profiler *prof,
                        cudaStream_t front, cudaEvent_t front_done, int front_kernels) {
    if (p.n == 0) return;
    const uint32_t Q = p.cfg.n_queries, L = p.cfg.n_fri_layers;
    // Pipelined calls run the small latency-bound kernels (K1, optionally K2) on a high-priority `front` stream so that their
    // CTAs are dispatched ahead of the pending Merkle CTAs of older batches: the next batch's Merkle kernel is then ready
    // to fill the SMs the moment the current one drains.
    const bool use_front = front && front_done && front_kernels > 0 && !prof;
    cudaStream_t s1 = use_front ? front : s, s2 = use_front && front_kernels >= 2 ? front : s;
    if (prof) prof->begin(0, s);
    {
        // NP = 2 (two transcripts per thread in lockstep) raises a warp's issue rate from 0.34 to 0.46 per cycle, but a batch of 1024 proofs has
        // fewer warps than the GPU has schedulers: the launch takes 0.211 ms instead of 0.143 ms.  A knob for very large batches only.
#ifdef SSYM_TUNING // experiment builds only (build.py --tuning): SSYM_CHANNEL_NP = 1 / 2 runs the one-thread-per-proof kernel with 1 / 2 transcripts per thread
        static const int np = [] { const char *e = getenv("SSYM_CHANNEL_NP"); return e ? atoi(e) : 0; }();
        if (np == 2 && p.n > 1) stwo_channel_kernel<2><<<((p.n + 1) / 2 + 63) / 64, 64, 0, s1>>>(p, sha_mul_consts());
        else if (np == 1) stwo_channel_kernel<1><<<(p.n + 63) / 64, 64, 0, s1>>>(p, sha_mul_consts());
        else
#endif
            stwo_channel_ws_kernel<<<(p.n + 31) / 32, 128, k1_dyn_smem(p), s1>>>(p, sha_mul_consts());
    }
    if (use_front && front_kernels < 2) { cudaEventRecord(front_done, front); cudaStreamWaitEvent(s, front_done, 0); }
    if (prof) { prof->end(0, s); prof->begin(1, s); }
    const uint32_t items = p.n * Q;
    switch (SSYM_STWO_COLUMNS(&p.cfg)) { // ssym_stwo_layout admits 4, 8, 16
    case 8: stwo_query_kernel<8><<<(p.n + K2_WARPS - 1) / K2_WARPS, 32 * K2_WARPS, 0, s2>>>(p); break;
    case 16: stwo_query_kernel<16><<<(p.n + K2_WARPS - 1) / K2_WARPS, 32 * K2_WARPS, 0, s2>>>(p); break;
    default: stwo_query_kernel<SSYM_NUM_COLUMNS><<<(p.n + K2_WARPS - 1) / K2_WARPS, 32 * K2_WARPS, 0, s2>>>(p); break;
    }
    if (p.dd.enabled && !p.derive_mode) { // decided by the caller (ssym_set_merkle_sharing); the compact form's derived siblings need the per-query kernel
        // shared-node schedule: plan (on the front stream with K1 / K2 when pipelined), hash the distinct nodes, check the followers, hash
        // what did not match, resolve every query
        stwo_plan_kernel<<<(p.n * (L + 2) * 16 + 511) / 512, 512, 0, s2>>>(p);
        if (use_front && front_kernels >= 2) { cudaEventRecord(front_done, front); cudaStreamWaitEvent(s, front_done, 0); }
        if (prof) { prof->end(1, s); prof->begin(2, s); }
        const uint32_t n_chains = p.n * p.dd.chains;
        const uint64_t max_warps = ((uint64_t)n_chains + 31) / 32 + STWO_DEDUP_MAX_BINS;
        stwo_merkle_shared_kernel<SSYM_DEFAULT_ADDMODE><<<(uint32_t)((max_warps + 3) / 4), 128, 0, s>>>(p, 0u, sha_mul_consts());
        stwo_check_kernel<<<(n_chains + 255) / 256, 256, 0, s>>>(p);
        stwo_merkle_shared_kernel<SSYM_DEFAULT_ADDMODE><<<(uint32_t)std::min<uint64_t>((max_warps + 3) / 4, 148 * 8), 128, 0, s>>>(p, 1u, sha_mul_consts());
        stwo_resolve_kernel<<<(n_chains + 255) / 256, 256, 0, s>>>(p);
        if (prof) { prof->end(2, s); prof->begin(3, s); }
        stwo_finalize_kernel<<<(p.n + 255) / 256, 256, 0, s>>>(p, accept_bits);
        if (prof) prof->end(3, s);
        if (launch_counter) *launch_counter += 8;
        return;
    }
    if (use_front && front_kernels >= 2) { cudaEventRecord(front_done, front); cudaStreamWaitEvent(s, front_done, 0); }
    if (prof) { prof->end(1, s); prof->begin(2, s); }
    const uint32_t groups = (items + 31) / 32;
    const uint64_t warps = (uint64_t)groups * (L + 3);
    const uint32_t wpc = SSYM_MERKLE_THREADS / 32, grid = (uint32_t)((warps + wpc - 1) / wpc);
#define SSYM_LAUNCH_MERKLE(AM, RL) stwo_merkle_kernel<AM, RL><<<grid, SSYM_MERKLE_THREADS, 0, s>>>(p, groups, sha_mul_consts())
    if (p.derive_mode) { // compact transport form, version 3: derived siblings (1) / the packer's scan (2)
        if (p.derive_mode == 1) stwo_merkle_kernel<SSYM_DEFAULT_ADDMODE, true, 1><<<grid, SSYM_MERKLE_THREADS, 0, s>>>(p, groups, sha_mul_consts());
        else stwo_merkle_kernel<SSYM_DEFAULT_ADDMODE, true, 2><<<grid, SSYM_MERKLE_THREADS, 0, s>>>(p, groups, sha_mul_consts());
    } else {
#ifdef SSYM_TUNING
    // experiment builds only (build.py --tuning; sha256.cuh): which adds go to the FMA pipe, and whether the 64 rounds are rolled into 4 x 16.
    // The release library contains the one variant DESIGN.md section 4 arrives at.
    static const int addmode = [] { const char *e = getenv("SSYM_ADDMODE"); return e ? atoi(e) : SSYM_DEFAULT_ADDMODE; }();
    static const int rolled = [] { const char *e = getenv("SSYM_ROLLED"); return e ? atoi(e) : SSYM_DEFAULT_ROLLED; }();
    switch (addmode * 2 + (rolled ? 1 : 0)) { // the multipliers (1, 2^k) must stay opaque to ptxas
    case 1: SSYM_LAUNCH_MERKLE(0, true); break;
    case 2: SSYM_LAUNCH_MERKLE(1, false); break;
    case 3: SSYM_LAUNCH_MERKLE(1, true); break;
    case 4: SSYM_LAUNCH_MERKLE(2, false); break;
    case 5: SSYM_LAUNCH_MERKLE(2, true); break;
    case 6: SSYM_LAUNCH_MERKLE(3, false); break;
    case 7: SSYM_LAUNCH_MERKLE(3, true); break;
    case 9: SSYM_LAUNCH_MERKLE(4, true); break;
    case 11: SSYM_LAUNCH_MERKLE(5, true); break;
    case 13: SSYM_LAUNCH_MERKLE(6, true); break;
    case 15: SSYM_LAUNCH_MERKLE(7, true); break;
    case 17: SSYM_LAUNCH_MERKLE(8, true); break;
    case 19: SSYM_LAUNCH_MERKLE(9, true); break;
    case 21: SSYM_LAUNCH_MERKLE(10, true); break;
    default: SSYM_LAUNCH_MERKLE(0, false); break;
    }
#else
    SSYM_LAUNCH_MERKLE(SSYM_DEFAULT_ADDMODE, SSYM_DEFAULT_ROLLED != 0);
#endif
    }
    if (prof) { prof->end(2, s); prof->begin(3, s); }
    stwo_finalize_kernel<<<(p.n + 255) / 256, 256, 0, s>>>(p, accept_bits);
    if (prof) prof->end(3, s);
    if (launch_counter) *launch_counter += 4;
}

void launch_stwo_verify_cross(const StwoParams &p, uint32_t rec_mode, uint32_t *accept_bits, cudaStream_t s, uint64_t *launch_counter) {
    if (p.n == 0) return;
    const uint32_t Q = p.cfg.n_queries, L = p.cfg.n_fri_layers;
    const uint32_t groups = (p.n * Q + 31) / 32;
    auto k2 = [&](const StwoParams &q) {
        switch (SSYM_STWO_COLUMNS(&q.cfg)) {
        case 8: stwo_query_kernel<8><<<(q.n + K2_WARPS - 1) / K2_WARPS, 32 * K2_WARPS, 0, s>>>(q); break;
        case 16: stwo_query_kernel<16><<<(q.n + K2_WARPS - 1) / K2_WARPS, 32 * K2_WARPS, 0, s>>>(q); break;
        default: stwo_query_kernel<SSYM_NUM_COLUMNS><<<(q.n + K2_WARPS - 1) / K2_WARPS, 32 * K2_WARPS, 0, s>>>(q); break;
        }
    };
    auto k3 = [&](const StwoParams &q, uint32_t ranks) {
        const uint64_t warps = (uint64_t)groups * ranks;
        const uint32_t wpc = SSYM_MERKLE_THREADS / 32;
        stwo_merkle_kernel<SSYM_DEFAULT_ADDMODE, true, 1><<<(uint32_t)((warps + wpc - 1) / wpc), SSYM_MERKLE_THREADS, 0, s>>>(q, groups, sha_mul_consts());
    };
    StwoParams pc = p; // the call's pass
    pc.ctx_mode = p.cfg.mode;
    pc.derive_mode = 1;
    pc.derive_kinds = 1u; // the trace and composition trees derive the same nodes under either semantics; the FRI siblings are complete by then
    pc.fri_only = 0;
    memset(&pc.dd, 0, sizeof pc.dd);
    stwo_channel_ws_kernel<<<(pc.n + 31) / 32, 128, k1_dyn_smem(pc), s>>>(pc, sha_mul_consts());
    StwoParams pr = pc; // the records' pass: evaluations and FRI chains only
    pr.cfg.mode = rec_mode;
    pr.status = nullptr; // its verdicts are not wanted
    pr.trace = nullptr;
    pr.derive_kinds = 2u;
    pr.fri_only = 1;
    k2(pr);
    k3(pr, L + 1);
    k2(pc);
    k3(pc, L + 3);
    stwo_finalize_kernel<<<(p.n + 255) / 256, 256, 0, s>>>(pc, accept_bits);
    if (launch_counter) *launch_counter += 6;
}

} // namespace ssym
