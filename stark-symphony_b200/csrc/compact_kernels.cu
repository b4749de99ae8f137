// SPDX-License-Identifier: MIT
#include "compact_kernels.cuh"

namespace ssym {

int compact_shape(const ssym_stwo_config_t &cfg, const ssym_stwo_layout_t &lo, CompactShape &sh) {
    const uint32_t Q = cfg.n_queries, L = cfg.n_fri_layers, G = cfg.lde_log;
    sh.trees = L + 3;
    uint32_t s = 0;
    for (uint32_t t = 0; t < sh.trees; t++) {
        sh.slot_first[t] = s;
        s += Q * (t < 2 ? G : G - 1 - (t - 2));
    }
    for (uint32_t t = sh.trees; t <= COMPACT_MAX_TREES; t++) sh.slot_first[t] = s;
    sh.slots = s;
    sh.head_slots = 2 * Q * G;
    sh.idx_bytes = Q * G <= 256 ? 1 : 2;
    sh.fixed_words = lo.off_trace_sib;
    sh.wit_words = lo.off_fri_sib[0] - lo.off_fri_wit;
    sh.bitmap_words = (sh.slots + 31u) / 32u;
    sh.off_wit = COMPACT_HDR_WORDS + sh.fixed_words;
    sh.off_bitmap = sh.off_wit + sh.wit_words;
    const uint32_t bm_padded = ((sh.bitmap_words + 7u) / 8u) * 8u;
    sh.off_refs = sh.off_bitmap + bm_padded;
    sh.off_bitmap2 = sh.off_refs;
    sh.off_refs3 = sh.off_bitmap2 + bm_padded;
    sh.max_words = sh.off_refs3 + 8u * sh.slots;
    sh.n_queries = Q;
    if (sh.bitmap_words > COMPACT_MAX_BITMAP_WORDS) return -1;
    // the packed sibling sections are contiguous in slot order (trace | composition) and (FRI layer 0 | 1 | ...)
    if (lo.off_cp_sib != lo.off_trace_sib + Q * G * 8 || lo.off_fri_wit != lo.off_cp_sib + Q * G * 8 || (sh.fixed_words & 7u) || (sh.wit_words & 7u)) return -1;
    return 0;
}

namespace {

__global__ void __launch_bounds__(256) stwo_expand_kernel(CompactParams p) {
    __shared__ uint32_t s_hdr[COMPACT_HDR_WORDS];
    __shared__ uint32_t s_bm[COMPACT_MAX_BITMAP_WORDS], s_pre[COMPACT_MAX_BITMAP_WORDS + 1];   // "new digest" bits, and how many before each word
    __shared__ uint32_t s_bm2[COMPACT_MAX_BITMAP_WORDS], s_pre2[COMPACT_MAX_BITMAP_WORDS + 1]; // version 3: "derived" bits
    __shared__ int s_ok;
    const CompactShape &sh = p.sh;
    const uint32_t i = blockIdx.x;
    const uint64_t o0 = p.offsets[i], o1 = p.offsets[i + 1];
    const uint32_t *rec = p.blob + (o0 - p.base);
    uint32_t *out = p.packed + (size_t)i * p.lo.stride_words;
    uint8_t *derive = p.derive ? p.derive + (size_t)i * sh.slots : nullptr;
    if (threadIdx.x == 0) {
        bool ok = o0 >= p.base && o1 >= o0 + sh.off_refs && ((o0 - p.base) & 7u) == 0 && o1 - o0 <= 0xffffffffull;
        if (ok) {
            for (int k = 0; k < COMPACT_HDR_WORDS; k++) s_hdr[k] = rec[k];
            const uint32_t D = s_hdr[1], R = s_hdr[3];
            if (s_hdr[2] == SSYM_COMPACT_MAGIC) { // version 2: new digests and back references only
                ok = s_hdr[0] == (uint32_t)(o1 - o0) && D <= sh.slots && R == sh.slots - D && s_hdr[0] == sh.off_refs + compact_refs_words(sh, R) + 8u * D;
                s_hdr[4] = 0;
            } else { // version 3: + X derived slots, bound to the semantics they were found under
                const uint32_t X = s_hdr[4];
                ok = s_hdr[2] == SSYM_COMPACT_MAGIC3 && s_hdr[0] == (uint32_t)(o1 - o0) && D <= sh.slots && X <= sh.slots - D && R == sh.slots - D - X &&
                     s_hdr[0] == sh.off_refs3 + compact_refs_words(sh, R, X) + 8u * D && (X == 0 || (derive != nullptr && s_hdr[5] == p.mode && (32u % sh.n_queries) == 0));
            }
        }
        s_ok = ok;
    }
    __syncthreads();
    bool bad = !s_ok;
    const bool v3 = !bad && s_hdr[2] == SSYM_COMPACT_MAGIC3;
    if (!bad) {
        for (uint32_t k = threadIdx.x; k < sh.bitmap_words; k += blockDim.x) {
            uint32_t w = rec[sh.off_bitmap + k], w2 = v3 ? rec[sh.off_bitmap2 + k] : 0u;
            if (k == sh.bitmap_words - 1 && (sh.slots & 31u)) { // bits behind the last slot do not count
                w &= (1u << (sh.slots & 31u)) - 1u;
                w2 &= (1u << (sh.slots & 31u)) - 1u;
            }
            s_bm[k] = w;
            s_bm2[k] = w2;
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            uint32_t acc = 0, acc2 = 0, clash = 0;
            for (uint32_t k = 0; k < sh.bitmap_words; k++) {
                s_pre[k] = acc; acc += __popc(s_bm[k]);
                s_pre2[k] = acc2; acc2 += __popc(s_bm2[k]);
                clash |= s_bm[k] & s_bm2[k];
            }
            s_pre[sh.bitmap_words] = acc;
            s_pre2[sh.bitmap_words] = acc2;
            if (acc != s_hdr[1] || acc2 != s_hdr[4] || clash) s_ok = 0; // the bitmaps do not announce D digests and X derived slots
        }
        __syncthreads();
        bad = !s_ok;
    }
    if (!bad) {
        const uint4 *src = reinterpret_cast<const uint4 *>(rec + COMPACT_HDR_WORDS);
        uint4 *dst = reinterpret_cast<uint4 *>(out);
        for (uint32_t k = threadIdx.x; k < sh.fixed_words / 4; k += blockDim.x) dst[k] = __ldg(src + k);
        src = reinterpret_cast<const uint4 *>(rec + sh.off_wit);
        dst = reinterpret_cast<uint4 *>(out + p.lo.off_fri_wit);
        for (uint32_t k = threadIdx.x; k < sh.wit_words / 4; k += blockDim.x) dst[k] = __ldg(src + k);
        const uint32_t R = s_hdr[3], X = s_hdr[4], refs_at = v3 ? sh.off_refs3 : sh.off_refs;
        const uint8_t *ref8 = reinterpret_cast<const uint8_t *>(rec + refs_at);
        const uint16_t *ref16 = reinterpret_cast<const uint16_t *>(rec + refs_at);
        const uint8_t *partner = ref8 + (size_t)R * sh.idx_bytes; // version 3: X partner bytes behind the back references
        const uint4 *tab = reinterpret_cast<const uint4 *>(rec + refs_at + compact_refs_words(sh, R, X));
        for (uint32_t s = threadIdx.x; s < sh.slots; s += blockDim.x) {
            const uint32_t below = (1u << (s & 31u)) - 1u;
            const uint32_t w = s_bm[s >> 5], before = s_pre[s >> 5] + __popc(w & below); // new digests before slot s
            const uint32_t w2 = s_bm2[s >> 5], before2 = s_pre2[s >> 5] + __popc(w2 & below); // derived slots before slot s
            uint32_t t = 0;
            while (t + 1 < sh.trees && s >= sh.slot_first[t + 1]) t++;
            uint32_t e = before; // a new digest: the next table entry
            bool ok = true, derived = false;
            uint8_t dbyte = 0xff;
            if ((w2 >> (s & 31u)) & 1u) { // derived: the node of query `partner`'s path of this tree at this level; filled in by the Merkle kernel
                derived = true;
                dbyte = partner[before2];
                const uint32_t depth = (sh.slot_first[t + 1] - sh.slot_first[t]) / sh.n_queries, own = (s - sh.slot_first[t]) / depth;
                ok = dbyte < sh.n_queries && dbyte != own;
            } else if (!((w >> (s & 31u)) & 1u)) { // a repeat: back reference number (s - before - before2), relative to the first table entry of the slot's tree
                const uint32_t f = sh.slot_first[t], tree_first = s_pre[f >> 5] + __popc(s_bm[f >> 5] & ((1u << (f & 31u)) - 1u));
                const uint32_t r = s - before - before2;
                e = tree_first + (sh.idx_bytes == 1 ? (uint32_t)ref8[r] : (uint32_t)ref16[r]);
                ok = e < before; // an earlier digest of the same tree (e >= tree_first by construction)
            }
            uint4 a = make_uint4(0, 0, 0, 0), b = a;
            if (ok && !derived) { a = __ldg(tab + 2 * e); b = __ldg(tab + 2 * e + 1); }
            if (!ok) bad = true;
            uint4 *d = reinterpret_cast<uint4 *>(out + (s < sh.head_slots ? p.lo.off_trace_sib + 8 * s : p.lo.off_fri_sib[0] + 8 * (s - sh.head_slots)));
            d[0] = a;
            d[1] = b;
            if (derive) derive[s] = ok ? dbyte : (uint8_t)0xff;
        }
    }
    const int any_bad = __syncthreads_or(bad); // malformed: the record expands to zeros and is flagged
    if (any_bad) {
        uint4 *dst = reinterpret_cast<uint4 *>(out);
        for (uint32_t k = threadIdx.x; k < p.lo.stride_words / 4; k += blockDim.x) dst[k] = make_uint4(0, 0, 0, 0);
        if (derive)
            for (uint32_t s = threadIdx.x; s < sh.slots; s += blockDim.x) derive[s] = 0xff;
    }
    if (p.flags && threadIdx.x == 0) p.flags[i] = any_bad ? 1u : 0u;
}

__global__ void compact_apply_flags_kernel(const uint32_t *flags, uint32_t *status, uint32_t *accept_bits, uint32_t n) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n || flags[i] == 0) return;
    if (status) status[i] |= SSYM_ST_SHAPE;
    atomicAnd(&accept_bits[i >> 5], ~(1u << (i & 31)));
}

} // namespace

void launch_stwo_expand(const CompactParams &p, cudaStream_t s) {
    if (p.n) stwo_expand_kernel<<<p.n, 256, 0, s>>>(p);
}
void launch_compact_apply_flags(const uint32_t *flags, uint32_t *status, uint32_t *accept_bits, uint32_t n, cudaStream_t s) {
    if (n) compact_apply_flags_kernel<<<(n + 255) / 256, 256, 0, s>>>(flags, status, accept_bits, n);
}

} // namespace ssym
