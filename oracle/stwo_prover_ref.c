/* SPDX-License-Identifier: MIT
 *
 * stwo_prover_ref.c — CPU reference PROVER for the wide-Fibonacci AIR of the stwo-verifier program.
 * (#included at the end of ssym_oracle.c: it reuses that file's static field / channel / hashing restatements.)
 *
 * THIS IS TEST INFRASTRUCTURE.  It is the checker for the GPU prover (csrc/prover_kernels.cu): tests compare the
 * packed proofs of both byte for byte.  Only tests/, __graft_entry__.smoke() and bench.py's checks may load it.
 *
 * The reference has NO prover for this AIR: its fixtures stwo-verifier/tests/data/proof{,_test}.json came from an
 * external stwo fork (hard part (b); section 8f rank 1).  What pins this file is therefore the
 * reference VERIFIER: every proof it emits must be accepted by oracle_stwo_verify_one in PROVER_CONSISTENT mode (the
 * semantics under which the shipped fixtures verify), i.e. pass stwo-verifier/src/verifier.simf:32-58
 * with F1-F3 resolved.  The statement proven is the one the verifier checks:
 *   - C = NUM_COLUMNS (config.simf:14; 4 at reference HEAD, also 8 and 16) trace columns over the canonic coset of size 2^trace_log,
 *     each row c_i = c_{i-1}^2 + c_{i-2}^2 for i >= 2 (constraints/wide_fibonacci.simf:24-62, all masks at offset 0),
 *     c0 = 1, c1 = SplitMix64(seed, row) mod p;
 *   - commitments = (SHA-256(""), trace root, composition root) mixed as in evals/commit.simf:20-35;
 *   - composition polynomial CP = (sum_{i>=2} alpha^(C-1-i) K_i) / vanishing(trace_log), K_i = c_i - c_{i-1}^2 - c_{i-2}^2 (the Horner fold of
 *     eval_column, wide_fibonacci.simf:33-52; alpha*K2 + K3 for C = 4), split into 16 M31 columns
 *     (index 4*coord + poly; poly = low two bits of the circle-FFT coefficient index, so that
 *     F(P) = Fa(2P) + y Fb(2P) + x Fc(2P) + xy Fd(2P), evals/composition_poly.simf:47-59) sampled at the doubled point;
 *   - DEEP quotient = fri_answer (fri/answers.simf:97-129 with Appendix A item 1) on the whole LDE domain;
 *   - FRI: circle_fold then n_fri_layers line_folds (fri/folding.simf:15-41) down to a constant (fri/layers.simf:73-78);
 *   - proof of work pow.simf:22-35, queries fri/queries.simf:30-43, decommitments evals/verify.simf, fri/layers.simf.
 * Requires n_fri_layers == trace_log - 1 (so that the last layer is a constant: the relation both presets of
 * config.simf:10-51 satisfy) and lde_log > trace_log.
 */

#include <pthread.h>

/* canonical-representative M31 arithmetic (inputs and outputs in [0, p)); equals m31_* of fields/m31.simf on canonical inputs */
static inline uint32_t fadd(uint32_t a, uint32_t b) { uint32_t s = a + b; return s >= M31_MODULUS ? s - M31_MODULUS : s; }
static inline uint32_t fsub(uint32_t a, uint32_t b) { return a >= b ? a - b : a + M31_MODULUS - b; }
static inline uint32_t fmul(uint32_t a, uint32_t b) {
    uint64_t t = (uint64_t)a * b;
    uint32_t s = (uint32_t)(t & M31_MODULUS) + (uint32_t)(t >> 31);
    return s >= M31_MODULUS ? s - M31_MODULUS : s;
}
static uint32_t finv(uint32_t a) { int keep = t_fail; uint32_t r = m31_inv(a); t_fail = keep; return r; }
static QM31 qcanon(QM31 q) { return qm31(m31(q.r.a), m31(q.r.b), m31(q.i.a), m31(q.i.b)); }

/* word-level SHA-256 of a message of n_words big-endian u32 (n_words <= 16 + 13) */
static void pr_compress(uint32_t h[8], const uint32_t blk[16]) {
    uint8_t b[64];
    for (int i = 0; i < 16; i++) { b[4 * i] = (uint8_t)(blk[i] >> 24); b[4 * i + 1] = (uint8_t)(blk[i] >> 16); b[4 * i + 2] = (uint8_t)(blk[i] >> 8); b[4 * i + 3] = (uint8_t)blk[i]; }
    sha_compress(h, b);
}
static void pr_sha_words(const uint32_t *w, uint32_t n_words, uint32_t out[8]) {
    static const uint32_t iv[8] = {0x6a09e667, 0xbb67ae85, 0x3c6ef372, 0xa54ff53a, 0x510e527f, 0x9b05688c, 0x1f83d9ab, 0x5be0cd19};
    uint32_t h[8], blk[16];
    memcpy(h, iv, sizeof iv);
    uint32_t done = 0;
    while (n_words - done >= 16) { pr_compress(h, w + done); done += 16; }
    uint32_t rem = n_words - done;
    memset(blk, 0, sizeof blk);
    memcpy(blk, w + done, rem * 4);
    blk[rem] = 0x80000000u;
    if (rem >= 14) { pr_compress(h, blk); memset(blk, 0, sizeof blk); }
    blk[15] = n_words * 32;
    pr_compress(h, blk);
    memcpy(out, h, 32);
}

/* ---- per-config tables ------------------------------------------------------------------------------- */
typedef struct {
    uint32_t log;        /* domain log size n */
    uint32_t *tw[32];    /* tw[l][j]: layer 0 = y of circle_domain(n) at bit_reverse(2j, n); layer l >= 1 = x of line_domain(n-l) at bit_reverse(2j, n-l) */
    uint32_t *itw[32];   /* inverses */
    M31Point *pt;        /* pt[q] = domain point of the bit-reversed slot q (= what fri_answer sees for query q) */
} PrDomain;

typedef struct {
    uint32_t trace_log, lde_log;
    PrDomain tr, lde;
    uint32_t *vanish_inv; /* 1 / vanishing_poly(trace_log)(pt[q]) on the LDE domain */
} PrTables;

static void pr_domain_build(PrDomain *d, uint32_t n) {
    d->log = n;
    CircleDomain cd = circle_domain((uint8_t)n);
    d->pt = (M31Point *)malloc(sizeof(M31Point) << n);
    for (uint32_t q = 0; q < (1u << n); q++) d->pt[q] = circle_position_to_m31_point(cd, bit_reverse_position(q, (uint8_t)n));
    for (uint32_t l = 0; l < n; l++) {
        uint32_t cnt = 1u << (n - l - 1);
        d->tw[l] = (uint32_t *)malloc(4u * cnt);
        d->itw[l] = (uint32_t *)malloc(4u * cnt);
        for (uint32_t j = 0; j < cnt; j++) {
            uint32_t t;
            if (l == 0) t = d->pt[2 * j].y;
            else { LineDomain ld = line_domain((uint8_t)(n - l)); t = line_position_to_x_coord(ld, bit_reverse_position(2 * j, (uint8_t)(n - l))); }
            d->tw[l][j] = m31(t);
            d->itw[l][j] = finv(d->tw[l][j]);
        }
    }
}

static PrTables *g_pr_tables[128]; /* one entry per (trace_log, lde_log) ever used: fewer than 128 pairs with lde_log <= 16 */
static pthread_mutex_t g_pr_lock = PTHREAD_MUTEX_INITIALIZER;
static const PrTables *pr_tables(uint32_t trace_log, uint32_t lde_log) {
    pthread_mutex_lock(&g_pr_lock);
    PrTables *t = NULL;
    for (int i = 0; i < 128; i++) {
        if (g_pr_tables[i] && g_pr_tables[i]->trace_log == trace_log && g_pr_tables[i]->lde_log == lde_log) { t = g_pr_tables[i]; break; }
        if (!g_pr_tables[i]) {
            t = (PrTables *)calloc(1, sizeof *t);
            t->trace_log = trace_log; t->lde_log = lde_log;
            pr_domain_build(&t->tr, trace_log);
            pr_domain_build(&t->lde, lde_log);
            t->vanish_inv = (uint32_t *)malloc(4u << lde_log);
            for (uint32_t q = 0; q < (1u << lde_log); q++) {
                uint32_t x = t->lde.pt[q].x;
                for (uint32_t k = 1; k < trace_log; k++) x = fsub(fadd(fmul(x, x), fmul(x, x)), 1); /* pi_fn, composition_poly.simf:26-35 */
                t->vanish_inv[q] = finv(x);
            }
            g_pr_tables[i] = t;
            break;
        }
    }
    pthread_mutex_unlock(&g_pr_lock);
    return t;
}

/* circle FFT in the basis b_j = y^j0 x^j1 pi(x)^j2 pi^2(x)^j3 ... (j0 = least significant bit of j):
 * evaluations live in bit-reversed slots of the canonic coset, coefficients in natural order. */
static void pr_ifft(const PrDomain *d, uint32_t *v) {
    uint32_t n = d->log, N = 1u << n;
    for (uint32_t l = 0; l < n; l++) {
        uint32_t stride = 1u << l;
        for (uint32_t s = 0; s < N; s++) {
            if (s & stride) continue;
            uint32_t a = v[s], b = v[s + stride];
            v[s] = fadd(a, b);
            v[s + stride] = fmul(fsub(a, b), d->itw[l][s >> (l + 1)]);
        }
    }
    uint32_t scale = finv(m31((uint32_t)1 << n)); /* n <= 30 */
    for (uint32_t s = 0; s < N; s++) v[s] = fmul(v[s], scale);
}
static void pr_fft(const PrDomain *d, uint32_t *v) {
    uint32_t n = d->log, N = 1u << n;
    for (uint32_t l = n; l-- > 0;) {
        uint32_t stride = 1u << l;
        for (uint32_t s = 0; s < N; s++) {
            if (s & stride) continue;
            uint32_t e = v[s], t = fmul(v[s + stride], d->tw[l][s >> (l + 1)]);
            v[s] = fadd(e, t);
            v[s + stride] = fsub(e, t);
        }
    }
}
/* sum_m c[m * stride] * prod_k t[k]^(bit k of m), m < 2^bits */
static QM31 pr_eval_at(const uint32_t *c, uint32_t stride, uint32_t bits, const QM31 *t) {
    uint32_t n = 1u << bits;
    QM31 *q = (QM31 *)malloc(sizeof(QM31) * n);
    for (uint32_t m = 0; m < n; m++) q[m] = qm31(c[(size_t)m * stride], 0, 0, 0);
    for (uint32_t k = 0; k < bits; k++) {
        n >>= 1;
        for (uint32_t m = 0; m < n; m++) q[m] = qm31_add(q[2 * m], qm31_mul(t[k], q[2 * m + 1]));
    }
    QM31 r = qcanon(q[0]);
    free(q);
    return r;
}

/* binary SHA-256 tree in heap order: node[1] = root, node[2^n + q] = leaf q (the verifier's auth_path numbering, merkle.simf:39-44) */
static void pr_tree_build(uint32_t *node, uint32_t n) {
    for (uint32_t i = (1u << n) - 1; i >= 1; i--) pr_sha_words(node + 16 * (size_t)i, 16, node + 8 * (size_t)i); /* children 2i, 2i+1 are adjacent */
}

static uint64_t pr_splitmix(uint64_t seed, uint64_t row) {
    uint64_t z = seed * 0x9E3779B97F4A7C15ull + (row + 1) * 0xBF58476D1CE4E5B9ull;
    z ^= z >> 30; z *= 0xBF58476D1CE4E5B9ull;
    z ^= z >> 27; z *= 0x94D049BB133111EBull;
    z ^= z >> 31;
    return z;
}

/* one trace row: c0 = 1, c1 = SplitMix64(seed, row) mod p, c_i = c_{i-1}^2 + c_{i-2}^2 */
static void pr_trace_row(uint64_t seed, uint32_t row, uint32_t n_columns, uint32_t *out) {
    out[0] = 1;
    out[1] = (uint32_t)(pr_splitmix(seed, row) >> 33) % M31_MODULUS;
    for (uint32_t i = 2; i < n_columns; i++) out[i] = fadd(fmul(out[i - 1], out[i - 1]), fmul(out[i - 2], out[i - 2]));
}

EXPORT int oracle_stwo_prove_one(const ssym_stwo_config_t *cfg, uint64_t seed, uint32_t *out) {
    ssym_stwo_layout_t lo;
    if (oracle_stwo_layout(cfg, &lo) != 0) return -1;
    const uint32_t T = cfg->trace_log, G = cfg->lde_log, Q = cfg->n_queries, L = cfg->n_fri_layers, C = SSYM_STWO_COLUMNS(cfg), QV = C + SSYM_NUM_CP_PARTITIONS;
    if (T < 2 || G <= T || L != T - 1) return -1;
    const PrTables *tb = pr_tables(T, G);
    if (!tb) return -1;
    const uint32_t NT = 1u << T, NG = 1u << G;
    memset(out, 0, 4u * lo.stride_words);
    int keep_fail = t_fail;

    /* 1. trace -> coefficients -> LDE */
    uint32_t *tcoef[SSYM_MAX_COLUMNS], *tlde[SSYM_MAX_COLUMNS];
    for (uint32_t c = 0; c < C; c++) { tcoef[c] = (uint32_t *)malloc(4u * NT); tlde[c] = (uint32_t *)calloc(NG, 4); }
    for (uint32_t r = 0; r < NT; r++) {
        uint32_t row[SSYM_MAX_COLUMNS];
        pr_trace_row(seed, r, C, row);
        for (uint32_t c = 0; c < C; c++) tcoef[c][r] = row[c];
    }
    for (uint32_t c = 0; c < C; c++) {
        pr_ifft(&tb->tr, tcoef[c]);
        memcpy(tlde[c], tcoef[c], 4u * NT);
        pr_fft(&tb->lde, tlde[c]);
    }
    /* 2. trace tree (hasher.simf:85-90 leaves) */
    uint32_t *tree_t = (uint32_t *)malloc(32u * 2 * NG), *tree_c = (uint32_t *)malloc(32u * 2 * NG);
    for (uint32_t q = 0; q < NG; q++) { uint32_t w[SSYM_MAX_COLUMNS]; for (uint32_t c = 0; c < C; c++) w[c] = tlde[c][q]; pr_sha_words(w, C, tree_t + 8 * (size_t)(NG + q)); }
    pr_tree_build(tree_t, G);
    /* 3. channel: evals/commit.simf:20-35 */
    u256 commitments[3];
    { Ctx8 c = sha_256_ctx_8_init(); commitments[0] = sha_256_ctx_8_finalize(c); } /* the constant tree is never decommitted; SHA-256("") as in the fixtures */
    commitments[1] = load_u256(tree_t + 8);
    ChannelState st = channel_init();
    channel_mix_u256(&st, commitments[0]);
    channel_mix_u256(&st, commitments[1]);
    QM31 cp_alpha = channel_draw_qm31(&st);
    uint32_t al[SSYM_MAX_COLUMNS][4]; /* al[j] = the four coordinates of cp_alpha^j: constraint i is weighted by alpha^(C-1-i) */
    { QM31 pw = qm31_one(); for (uint32_t j = 0; j + 2 < C; j++) { qm31_to_w(qcanon(pw), al[j]); pw = qm31_mul(pw, cp_alpha); } }
    /* 4. composition polynomial on the LDE domain, per QM31 coordinate; interpolate; split by the low two coefficient-index bits */
    uint32_t *cpc[4], *cplde[16];
    for (int k = 0; k < 4; k++) cpc[k] = (uint32_t *)malloc(4u * NG);
    for (uint32_t q = 0; q < NG; q++) {
        uint32_t acc[4] = {0, 0, 0, 0};
        for (uint32_t i = 2; i < C; i++) {
            uint32_t a = tlde[i - 2][q], b = tlde[i - 1][q];
            uint32_t ki = fsub(tlde[i][q], fadd(fmul(b, b), fmul(a, a)));
            for (int k = 0; k < 4; k++) acc[k] = fadd(acc[k], fmul(al[C - 1 - i][k], ki));
        }
        for (int k = 0; k < 4; k++) cpc[k][q] = fmul(acc[k], tb->vanish_inv[q]);
    }
    int degree_ok = 1;
    for (int k = 0; k < 4; k++) {
        pr_ifft(&tb->lde, cpc[k]);
        for (uint32_t j = NT + 1; j < NG; j++) if (cpc[k][j]) degree_ok = 0; /* CP has total degree <= 2^(T-1): coefficients live in [0, 2^T] */
        for (int p = 0; p < 4; p++) {
            uint32_t *col = (uint32_t *)calloc(NG, 4);
            for (uint32_t m = 0; 4 * m + p <= NT; m++) col[2 * m] = cpc[k][4 * m + p]; /* (X,Y)-basis: no Y, X-bits shifted down by one */
            pr_fft(&tb->lde, col);
            cplde[4 * k + p] = col;
        }
    }
    for (uint32_t q = 0; q < NG; q++) { uint32_t w[16]; for (int k = 0; k < 16; k++) w[k] = cplde[k][q]; pr_sha_words(w, 16, tree_c + 8 * (size_t)(NG + q)); }
    pr_tree_build(tree_c, G);
    commitments[2] = load_u256(tree_c + 8);
    channel_mix_u256(&st, commitments[2]);
    /* 5. OODS (deep/oods.simf:44-64): draw P, sample trace at P and CP columns at 2P */
    QM31Point P = channel_draw_qm31_point(&st);
    P.x = qcanon(P.x); P.y = qcanon(P.y);
    QM31Point P2;
    P2.x = qm31_point_dbl_x(P.x);
    { QM31 xy = qm31_mul(P.x, P.y); P2.y = qm31_add(xy, xy); }
    QM31 tw[32];
    tw[0] = P.y; tw[1] = P.x;
    for (uint32_t k = 2; k < 32; k++) tw[k] = qm31_point_dbl_x(tw[k - 1]);
    QM31 oods_trace[SSYM_MAX_COLUMNS], oods_cp[16];
    for (uint32_t c = 0; c < C; c++) oods_trace[c] = pr_eval_at(tcoef[c], 1, T, tw);
    QM31 tw2[32];
    tw2[0] = P2.x;
    for (uint32_t k = 1; k < 32; k++) tw2[k] = qm31_point_dbl_x(tw2[k - 1]);
    for (int k = 0; k < 4; k++)
        for (int p = 0; p < 4; p++) oods_cp[4 * k + p] = pr_eval_at(cpc[k] + p, 4, T - 1, tw2); /* m < 2^(T-1) covers 4m+p <= 2^T */
    channel_mix_oods_evals(&st, oods_trace, C, oods_cp);
    QM31 deep_alpha = channel_draw_qm31(&st);
    /* 6. DEEP quotient on the whole LDE domain = fri_answer of every position (Appendix A item 1) */
    QM31 *h = (QM31 *)malloc(sizeof(QM31) * NG);
    {
        LineCoeffs ka[16], kb[SSYM_MAX_COLUMNS];
        QM31 alpha_i = deep_alpha;
        QM31 sa_a = qm31_zero(), sa_c = qm31_zero(), sb_a = qm31_zero(), sb_c = qm31_zero();
        for (int c = 0; c < 16; c++) { ka[c] = deep_quotient_interpolant_coefficients(P2, oods_cp[c], alpha_i); sa_a = qm31_add(sa_a, ka[c].a); sa_c = qm31_add(sa_c, ka[c].c); alpha_i = qm31_mul(alpha_i, deep_alpha); }
        for (uint32_t c = 0; c < C; c++) { kb[c] = deep_quotient_interpolant_coefficients(P, oods_trace[c], alpha_i); sb_a = qm31_add(sb_a, kb[c].a); sb_c = qm31_add(sb_c, kb[c].c); alpha_i = qm31_mul(alpha_i, deep_alpha); }
        for (uint32_t q = 0; q < NG; q++) {
            M31Point dp = tb->lde.pt[q];
            QM31 na = qm31_zero(), nb = qm31_zero();
            for (int c = 0; c < 16; c++) na = qm31_add(na, qm31_mul_m31(ka[c].b, cplde[c][q]));
            for (uint32_t c = 0; c < C; c++) nb = qm31_add(nb, qm31_mul_m31(kb[c].b, tlde[c][q]));
            na = qm31_sub(na, qm31_add(qm31_mul_m31(sa_a, dp.y), sa_c));
            nb = qm31_sub(nb, qm31_add(qm31_mul_m31(sb_a, dp.y), sb_c));
            h[q] = qcanon(qm31_add(qm31_mul_cm31(na, deep_quotient_denominator_inverse(P2, dp)), qm31_mul_cm31(nb, deep_quotient_denominator_inverse(P, dp))));
        }
    }
    /* 7. FRI commit (fri/commit.simf:72-85) + folding (fri/folding.simf:15-41) */
    uint32_t *ftree[SSYM_MAX_FRI_LAYERS];
    QM31 *fev[SSYM_MAX_FRI_LAYERS + 1];
    fev[0] = h;
    for (uint32_t l = 0; l <= L; l++) {
        uint32_t n = G - l, N = 1u << n;
        ftree[l] = (uint32_t *)malloc(32u * 2 * N);
        for (uint32_t q = 0; q < N; q++) { uint32_t w[4]; qm31_to_w(fev[l][q], w); pr_sha_words(w, 4, ftree[l] + 8 * (size_t)(N + q)); }
        pr_tree_build(ftree[l], n);
        memcpy(out + (l == 0 ? lo.off_fri_first_root : lo.off_fri_inner_root + 8 * (l - 1)), ftree[l] + 8, 32);
        channel_mix_u256(&st, load_u256(ftree[l] + 8));
        QM31 a = channel_draw_qm31(&st);
        fev[l + 1] = (QM31 *)malloc(sizeof(QM31) * (N / 2));
        for (uint32_t j = 0; j < N / 2; j++) {
            QM31 f0 = qm31_add(fev[l][2 * j], fev[l][2 * j + 1]);
            QM31 f1 = qm31_mul_m31(qm31_sub(fev[l][2 * j], fev[l][2 * j + 1]), tb->lde.itw[l][j]);
            fev[l + 1][j] = qcanon(qm31_add(f0, qm31_mul(a, f1)));
        }
    }
    uint32_t last_n = 1u << (G - 1 - L);
    QM31 last_coeff = fev[L + 1][0];
    for (uint32_t j = 1; j < last_n; j++) if (!qm31_eq(fev[L + 1][j], last_coeff)) degree_ok = 0;
    { /* channel_mix_line_poly fri/commit.simf:48-57 */
        Ctx8 c = sha_256_ctx_8_init();
        c = sha_256_ctx_8_add_32(c, st.digest);
        c = hasher_add_qm31(last_coeff, c);
        st.digest = sha_256_ctx_8_finalize(c);
        st.n_sent = 0;
    }
    /* 8. proof of work (pow.simf:22-35): smallest nonce that passes */
    uint64_t nonce = 0;
    for (;; nonce++) {
        ChannelState s2 = st;
        t_fail = 0;
        check_proof_of_work(&s2, nonce, cfg->pow_target);
        if (!t_fail) { st = s2; break; }
    }
    /* 9. queries (fri/queries.simf:30-43) and decommitments */
    uint32_t queries[SSYM_MAX_QUERIES];
    for (uint32_t q = 0; q < Q; q += 8) {
        u256 w = channel_draw_u256(&st);
        for (uint32_t j = 0; j < 8 && q + j < Q; j++) queries[q + j] = w.w[j] & (NG - 1);
    }
    uint32_t U = Q;
    if (cfg->mode & SSYM_MODE_QUERY_DEDUP) { /* include/ssym.h: sorted distinct queries in the first U slots, the other slots zero-filled */
        for (uint32_t a = 1; a < Q; a++)
            for (uint32_t b = a; b > 0 && queries[b] < queries[b - 1]; b--) { uint32_t t_ = queries[b]; queries[b] = queries[b - 1]; queries[b - 1] = t_; }
        U = 0;
        for (uint32_t a = 0; a < Q; a++)
            if (a == 0 || queries[a] != queries[U - 1]) queries[U++] = queries[a];
    }
    for (int i = 0; i < 3; i++) store_u256(out + lo.off_commit + 8 * i, commitments[i]);
    for (uint32_t c = 0; c < C; c++) qm31_to_w(oods_trace[c], out + lo.off_oods_trace + 4 * c);
    for (int c = 0; c < 16; c++) qm31_to_w(oods_cp[c], out + lo.off_oods_cp + 4 * c);
    qm31_to_w(last_coeff, out + lo.off_last_coeff);
    out[lo.off_pow_nonce] = (uint32_t)(nonce >> 32);
    out[lo.off_pow_nonce + 1] = (uint32_t)nonce;
    for (uint32_t qi = 0; qi < U; qi++) { /* slots >= U stay zero (the record was cleared on entry) */
        uint32_t q = queries[qi];
        uint32_t *qv = out + lo.off_qvals + QV * qi;
        for (uint32_t c = 0; c < C; c++) qv[c] = tlde[c][q];
        for (int c = 0; c < 16; c++) qv[C + c] = cplde[c][q];
        uint32_t node = NG + q;
        for (uint32_t lev = 0; lev < G; lev++, node >>= 1) {
            memcpy(out + lo.off_trace_sib + (qi * G + lev) * 8, tree_t + 8 * (size_t)(node ^ 1), 32);
            memcpy(out + lo.off_cp_sib + (qi * G + lev) * 8, tree_c + 8 * (size_t)(node ^ 1), 32);
        }
        uint32_t fq = q;
        for (uint32_t l = 0; l <= L; l++) {
            uint32_t n = G - l, N = 1u << n;
            qm31_to_w(fev[l][fq ^ 1], out + lo.off_fri_wit + (l * Q + qi) * 4);
            uint32_t nd = (N + fq) >> 1;
            for (uint32_t lev = 0; lev + 1 < n; lev++, nd >>= 1) memcpy(out + lo.off_fri_sib[l] + (qi * (n - 1) + lev) * 8, ftree[l] + 8 * (size_t)(nd ^ 1), 32);
            fq >>= 1;
        }
    }
    for (uint32_t c = 0; c < C; c++) { free(tcoef[c]); free(tlde[c]); }
    for (int k = 0; k < 4; k++) free(cpc[k]);
    for (int c = 0; c < 16; c++) free(cplde[c]);
    for (uint32_t l = 0; l <= L; l++) free(ftree[l]);
    for (uint32_t l = 0; l <= L + 1; l++) free(fev[l]);
    free(tree_t); free(tree_c);
    t_fail = keep_fail;
    return degree_ok ? 0 : 1; /* 1: the low-degree sanity checks failed (a bug in this file, never data) */
}

EXPORT int oracle_stwo_prove_batch(const ssym_stwo_config_t *cfg, const uint64_t *seeds, size_t begin, size_t end, uint32_t *out) {
    ssym_stwo_layout_t lo;
    if (oracle_stwo_layout(cfg, &lo) != 0) return -1;
    int rc = 0;
    for (size_t i = begin; i < end; i++) {
        int r = oracle_stwo_prove_one(cfg, seeds[i], out + i * (size_t)lo.stride_words);
        if (r) rc = r;
    }
    return rc;
}

/* The trace row of a seed (what both provers commit to), for tests. */
EXPORT void oracle_stwo_trace_row(uint64_t seed, uint32_t row, uint32_t out[4]) { pr_trace_row(seed, row, SSYM_NUM_COLUMNS, out); }
EXPORT void oracle_stwo_trace_row_n(uint64_t seed, uint32_t row, uint32_t n_columns, uint32_t *out) { pr_trace_row(seed, row, n_columns, out); }
