// SPDX-License-Identifier: MIT
// GPU `.wit` tokeniser / packer; see wit_kernels.cuh for the scheme.
#include "wit_kernels.cuh"

namespace ssym {
namespace {

constexpr int WIT_THREADS = 256;
constexpr int WIT_MAXQ = 128; // quotes per witness file handled on the GPU (the generator's output has 60)

enum : uint8_t { C_BAD = 0, C_WS, C_NUM, C_STRUCT, C_L, C_LCONT };

__device__ __forceinline__ bool is_ws(uint8_t c) { return c == ' ' || c == '\n' || c == '\t' || c == '\r'; }

// exclusive block scan of a pair of counters; `total` = block sums (valid in every thread)
__device__ __forceinline__ uint2 block_excl_scan(uint2 v, uint2 *s_warp, uint2 &total) {
    const uint32_t lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    uint2 inc = v;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        const uint32_t x = __shfl_up_sync(0xffffffffu, inc.x, off), y = __shfl_up_sync(0xffffffffu, inc.y, off);
        if (lane >= (uint32_t)off) { inc.x += x; inc.y += y; }
    }
    if (lane == 31) s_warp[wid] = inc;
    __syncthreads();
    uint2 woff = make_uint2(0, 0);
    total = make_uint2(0, 0);
#pragma unroll
    for (int w = 0; w < WIT_THREADS / 32; w++) {
        const uint2 t = s_warp[w];
        if ((uint32_t)w < wid) { woff.x += t.x; woff.y += t.y; }
        total.x += t.x;
        total.y += t.y;
    }
    __syncthreads();
    return make_uint2(inc.x - v.x + woff.x, inc.y - v.y + woff.y);
}

__device__ bool str_is(const uint8_t *t, uint32_t s, uint32_t e, const char *lit, uint32_t n) {
    if (e - s != n) return false;
    for (uint32_t k = 0; k < n; k++)
        if (t[s + k] != (uint8_t)lit[k]) return false;
    return true;
}
__device__ int name_id(const uint8_t *t, uint32_t s, uint32_t e) { // stwo-verifier/src/main.simf:9-25
    if (str_is(t, s, e, "COMMITMENTS", 11)) return 0;
    if (str_is(t, s, e, "DECOMMITMENTS", 13)) return 1;
    if (str_is(t, s, e, "OODS_EVALS", 10)) return 2;
    if (str_is(t, s, e, "FRI_COMMITMENTS", 15)) return 3;
    if (str_is(t, s, e, "FRI_DECOMMITMENTS", 17)) return 4;
    if (str_is(t, s, e, "POW_NONCE", 9)) return 5;
    return -1;
}

// The JSON level, by one thread: { NAME: { "value": "<text>", "type": "<text>" }, ... } (simfony-cli/src/main.rs:77-81).  `q` holds the
// sorted positions of the `nq` quote characters (there are no backslashes in the file).  Fills the value span of every name.
__device__ bool json_walk(const uint8_t *t, uint32_t len, const uint32_t *q, uint32_t nq, uint32_t *vstart, uint32_t *vend) {
    uint32_t pos = 0, k = 0, seen = 0;
    auto skip = [&]() { while (pos < len && is_ws(t[pos])) pos++; };
    auto expect = [&](uint8_t ch) { skip(); if (pos < len && t[pos] == ch) { pos++; return true; } return false; };
    auto string = [&](uint32_t &s, uint32_t &e) {
        skip();
        if (k + 1 >= nq || q[k] != pos) return false;
        s = q[k] + 1;
        e = q[k + 1];
        pos = e + 1;
        k += 2;
        return true;
    };
    if (!expect('{')) return false;
    for (;;) {
        uint32_t ns, ne;
        if (!string(ns, ne) || !expect(':') || !expect('{')) return false;
        bool have = false;
        uint32_t vs = 0, ve = 0;
        for (;;) {
            uint32_t ks, ke, s, e;
            if (!string(ks, ke) || !expect(':') || !string(s, e)) return false; // members other than strings: host parser
            if (str_is(t, ks, ke, "value", 5)) {
                if (have) return false;
                have = true;
                vs = s;
                ve = e;
            }
            skip();
            if (pos < len && t[pos] == ',') { pos++; continue; }
            break;
        }
        if (!expect('}') || !have || vs == ve) return false;
        const int id = name_id(t, ns, ne);
        if (id < 0 || (seen >> id & 1)) return false;
        seen |= 1u << id;
        vstart[id] = vs;
        vend[id] = ve;
        skip();
        if (pos < len && t[pos] == ',') { pos++; continue; }
        if (!expect('}')) return false;
        break;
    }
    skip();
    return pos == len && k == nq && seen == (1u << WIT_NAMES) - 1;
}

// One integer literal starting at t[pos] (inside a value that ends at `end`) into its packed slot.
__device__ bool parse_number(const uint8_t *t, uint32_t pos, uint32_t end, const uint8_t *cls, uint32_t slot, uint32_t *rec) {
    const uint32_t off = slot & 0x0fffffffu, kind = slot >> 28;
    const uint32_t kw = kind == WIT_KIND_U32 ? 1u : kind == WIT_KIND_U64 ? 2u : 8u;
    if (t[pos] == '0' && pos + 1 < end && t[pos + 1] == 'x') {
        const uint32_t p = pos + 2;
        uint32_t n = 0;
        while (p + n < end && cls[t[p + n]] == C_NUM) n++;
        if (n == 0 || n > 64) return false;
        uint32_t acc = 0, rem = n;
        for (uint32_t j = 0; j < n; j++) {
            const uint8_t c = t[p + j];
            uint32_t d;
            if (c >= '0' && c <= '9') d = c - '0';
            else if (c >= 'a' && c <= 'f') d = c - 'a' + 10;
            else return false; // a second 'x'
            acc = (acc << 4) | d;
            rem--;
            if ((rem & 7u) == 0) {
                const uint32_t widx = 7u - rem / 8u; // index in the 8-word big-endian value
                if (widx >= 8u - kw) rec[off + widx - (8u - kw)] = acc;
                else if (acc) return false; // does not fit the slot's type
                acc = 0;
            }
        }
        return true;
    }
    uint64_t v = 0;
    for (uint32_t p = pos; p < end && cls[t[p]] == C_NUM; p++) {
        const uint8_t c = t[p];
        if (c < '0' || c > '9') return false;
        const uint32_t d = c - '0';
        if (v > (0xffffffffffffffffull - d) / 10u) return false; // above 64 bits: host parser
        v = v * 10u + d;
    }
    if (kw == 1) {
        if (v >> 32) return false;
        rec[off] = (uint32_t)v;
    } else {
        rec[off + kw - 2] = (uint32_t)(v >> 32);
        rec[off + kw - 1] = (uint32_t)v;
    }
    return true;
}

struct Spans { // sorted by position
    uint32_t start[WIT_NAMES], end[WIT_NAMES], name[WIT_NAMES];
    uint32_t owner[WIT_NAMES], ltok[WIT_NAMES], lnum[WIT_NAMES]; // thread whose chunk holds the span start, its local counts there
    uint32_t tokbase[WIT_NAMES], numbase[WIT_NAMES];              // global token / literal index of the span's first token
};

// One pass of a thread over its chunk [lo, hi) of the file: tokens of the value spans.  COUNT pass: token / literal counts (returned) and
// the local counts at every span start inside the chunk.  EMIT pass: tokens are checked against the skeleton, literals are parsed.
template <bool EMIT>
__device__ __forceinline__ uint2 scan_values(const uint8_t *t, uint32_t lo, uint32_t hi, const uint8_t *cls, Spans &sp, const WitTables &tab, uint2 base,
                                             uint32_t *rec, int *bad) {
    uint32_t tok = 0, num = 0;
    for (int s = 0; s < WIT_NAMES; s++) {
        const uint32_t ss = sp.start[s], se = sp.end[s];
        const uint32_t a = max(lo, ss), b = min(hi, se);
        if (a >= b) continue;
        uint8_t prevc = 0;
        bool prevnum = false;
        if (a == ss) {
            if (!EMIT) { sp.owner[s] = threadIdx.x; sp.ltok[s] = tok; sp.lnum[s] = num; }
        } else {
            prevc = __ldg(t + a - 1);
            prevnum = cls[prevc] == C_NUM;
        }
        const uint32_t name = sp.name[s];
        const uint8_t *skel = tab.skel + tab.skel_off[name];
        const uint32_t *slots = tab.slots + tab.slot_off[name];
        const uint32_t skel_len = tab.skel_len[name], slot_cnt = tab.slot_cnt[name];
        const uint32_t tb = EMIT ? sp.tokbase[s] : 0, nb = EMIT ? sp.numbase[s] : 0;
        for (uint32_t pos = a; pos < b; pos++) {
            const uint8_t c = __ldg(t + pos), k = cls[c];
            if (k == C_NUM) {
                if (!prevnum) {
                    if (EMIT) {
                        const uint32_t rel = base.x + tok - tb, nrel = base.y + num - nb;
                        if (rel >= skel_len || skel[rel] != 'N' || nrel >= slot_cnt || !parse_number(t, pos, se, cls, slots[nrel], rec)) *bad = 1;
                    }
                    tok++;
                    num++;
                }
                prevnum = true;
            } else {
                prevnum = false;
                if (k == C_STRUCT || k == C_L) {
                    if (EMIT) {
                        const uint32_t rel = base.x + tok - tb;
                        if (rel >= skel_len || skel[rel] != (k == C_L ? (uint8_t)'L' : c)) *bad = 1;
                        if (k == C_L && !(pos + 4 < se && t[pos + 1] == 'i' && t[pos + 2] == 's' && t[pos + 3] == 't' && t[pos + 4] == '!')) *bad = 1;
                    }
                    tok++;
                } else if (k == C_LCONT) { // i s t ! : only as the tail of `list!`
                    const uint8_t need = c == 'i' ? 'l' : c == 's' ? 'i' : c == 't' ? 's' : 't';
                    if (prevc != need) *bad = 1;
                } else if (k != C_WS) {
                    *bad = 1;
                }
            }
            prevc = c;
        }
    }
    return make_uint2(tok, num);
}

__global__ void __launch_bounds__(WIT_THREADS) wit_pack_kernel(WitParams p) {
    __shared__ uint8_t s_cls[256];
    __shared__ uint32_t s_q[WIT_MAXQ], s_qsorted[WIT_MAXQ];
    __shared__ uint2 s_warp[WIT_THREADS / 32];
    __shared__ Spans sp;
    __shared__ uint32_t s_nq;
    __shared__ int s_bad;
    const uint32_t i = blockIdx.x, tid = threadIdx.x;
    const uint64_t b0 = p.offsets[i], b1 = p.offsets[i + 1];
    const uint8_t *t = p.text + b0;
    uint32_t *rec = p.packed + (size_t)i * p.stride_words;
    if (b1 < b0 || b1 - b0 >= 0x7fffffffull) {
        if (tid == 0) p.flags[i] = SSYM_WIT_SLOW;
        return;
    }
    const uint32_t len = (uint32_t)(b1 - b0);
    {
        const uint8_t c = (uint8_t)tid;
        uint8_t k = C_BAD;
        if (is_ws(c)) k = C_WS;
        else if ((c >= '0' && c <= '9') || (c >= 'a' && c <= 'f') || c == 'x') k = C_NUM;
        else if (c == '(' || c == ')' || c == '[' || c == ']' || c == ',') k = C_STRUCT;
        else if (c == 'l') k = C_L;
        else if (c == 'i' || c == 's' || c == 't' || c == '!') k = C_LCONT;
        s_cls[tid] = k;
    }
    if (tid == 0) { s_nq = 0; s_bad = 0; }
    __syncthreads();
    const uint32_t chunk = (len + WIT_THREADS - 1) / WIT_THREADS;
    const uint32_t lo = min(len, tid * chunk), hi = min(len, lo + chunk);

    // ---- phase 0: the JSON strings.  Quote positions, no escapes. ----
    for (uint32_t pos = lo; pos < hi; pos++) {
        const uint8_t c = __ldg(t + pos);
        if (c == '"') {
            const uint32_t k = atomicAdd(&s_nq, 1u);
            if (k < WIT_MAXQ) s_q[k] = pos;
        } else if (c == '\\') {
            s_bad = 1;
        }
    }
    __syncthreads();
    const uint32_t nq = s_nq;
    if (nq > WIT_MAXQ || (nq & 1u) || s_bad) {
        if (tid == 0) p.flags[i] = SSYM_WIT_SLOW;
        return;
    }
    if (tid < nq) { // rank sort (positions are distinct)
        const uint32_t v = s_q[tid];
        uint32_t r = 0;
        for (uint32_t k = 0; k < nq; k++) r += s_q[k] < v;
        s_qsorted[r] = v;
    }
    __syncthreads();
    if (tid == 0) {
        uint32_t vs[WIT_NAMES], ve[WIT_NAMES];
        if (!json_walk(t, len, s_qsorted, nq, vs, ve)) {
            s_bad = 1;
        } else { // spans in file order
            uint32_t order[WIT_NAMES];
            for (int a = 0; a < WIT_NAMES; a++) order[a] = a;
            for (int a = 1; a < WIT_NAMES; a++)
                for (int b = a; b > 0 && vs[order[b]] < vs[order[b - 1]]; b--) { const uint32_t x = order[b]; order[b] = order[b - 1]; order[b - 1] = x; }
            for (int a = 0; a < WIT_NAMES; a++) { sp.start[a] = vs[order[a]]; sp.end[a] = ve[order[a]]; sp.name[a] = order[a]; }
        }
    }
    __syncthreads();
    if (s_bad) {
        if (tid == 0) p.flags[i] = SSYM_WIT_SLOW;
        return;
    }

    // ---- phase 1: count the tokens of every value, check the totals against the skeletons ----
    int bad = 0;
    const uint2 mine = scan_values<false>(t, lo, hi, s_cls, sp, p.tab, make_uint2(0, 0), rec, &bad);
    if (bad) s_bad = 1;
    uint2 total;
    const uint2 base = block_excl_scan(mine, s_warp, total); // (has the barriers that publish sp.owner / ltok / lnum and s_bad)
    // owners publish the global index of their span's first token
    for (int s = 0; s < WIT_NAMES; s++)
        if (sp.owner[s] == tid) { sp.tokbase[s] = base.x + sp.ltok[s]; sp.numbase[s] = base.y + sp.lnum[s]; }
    __syncthreads();
    if (tid == 0) {
        for (int s = 0; s < WIT_NAMES; s++) {
            const uint32_t ntok = (s + 1 < WIT_NAMES ? sp.tokbase[s + 1] : total.x) - sp.tokbase[s];
            const uint32_t nnum = (s + 1 < WIT_NAMES ? sp.numbase[s + 1] : total.y) - sp.numbase[s];
            if (ntok != p.tab.skel_len[sp.name[s]] || nnum != p.tab.slot_cnt[sp.name[s]]) s_bad = 1;
        }
    }
    __syncthreads();
    if (s_bad) {
        if (tid == 0) p.flags[i] = SSYM_WIT_SLOW;
        return;
    }

    // ---- phase 2: check every token against the skeleton, parse and scatter the literals ----
    scan_values<true>(t, lo, hi, s_cls, sp, p.tab, base, rec, &bad);
    if (bad) s_bad = 1;
    __syncthreads();
    if (tid == 0) p.flags[i] = s_bad ? SSYM_WIT_SLOW : SSYM_WIT_OK;
}

__global__ void wit_apply_flags_kernel(const uint32_t *flags, uint32_t *status, uint32_t *accept_bits, uint32_t n) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n || flags[i] == SSYM_WIT_OK) return;
    if (status) status[i] |= SSYM_ST_SHAPE;
    atomicAnd(&accept_bits[i >> 5], ~(1u << (i & 31)));
}

} // namespace

void launch_wit_pack(const WitParams &p, cudaStream_t s) {
    if (p.n) wit_pack_kernel<<<p.n, WIT_THREADS, 0, s>>>(p);
}
void launch_wit_apply_flags(const uint32_t *flags, uint32_t *status, uint32_t *accept_bits, uint32_t n, cudaStream_t s) {
    if (n) wit_apply_flags_kernel<<<(n + 255) / 256, 256, 0, s>>>(flags, status, accept_bits, n);
}

} // namespace ssym
