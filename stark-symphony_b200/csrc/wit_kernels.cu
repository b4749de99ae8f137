// SPDX-License-Identifier: MIT
// GPU `.wit` tokeniser / packer; see wit_kernels.cuh for the scheme.
#include "wit_kernels.cuh"

#include <algorithm>

namespace ssym {
namespace {

constexpr int WIT_MAXQ = 128; // quotes per witness file handled on the GPU (the generator's output has 60)

__device__ __forceinline__ bool is_ws(uint8_t c) { return c == ' ' || c == '\n' || c == '\t' || c == '\r'; }

__device__ bool str_is(const uint8_t *t, uint32_t s, uint32_t e, const char *lit, uint32_t n) {
    if (e - s != n) return false;
    for (uint32_t k = 0; k < n; k++)
        if (t[s + k] != (uint8_t)lit[k]) return false;
    return true;
}
__device__ int name_id(const WitTables &tab, const uint8_t *t, uint32_t s, uint32_t e) {
    for (uint32_t k = 0; k < tab.n_names; k++)
        if (str_is(t, s, e, tab.name[k], tab.name_len[k])) return (int)k;
    return -1;
}

// The JSON level, by one thread: { NAME: { "value": "<text>", "type": "<text>" }, ... } (simfony-cli/src/main.rs:77-81).  `q` holds the
// sorted positions of the `nq` quote characters (there are no backslashes in the file).  Fills the value span of every name.
__device__ bool json_walk(const WitTables &tab, const uint8_t *t, uint32_t len, const uint32_t *q, uint32_t nq, uint32_t *vstart, uint32_t *vend) {
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
        bool have = false, have_type = false;
        uint32_t vs = 0, ve = 0;
        for (;;) {
            uint32_t ks, ke, s, e;
            if (!string(ks, ke) || !expect(':') || !string(s, e)) return false; // members other than strings: host parser
            if (str_is(t, ks, ke, "value", 5)) {
                if (have) return false;
                have = true;
                vs = s;
                ve = e;
            } else if (str_is(t, ks, ke, "type", 4)) {
                if (have_type) return false;
                have_type = true;
            } else {
                return false; // a member serde's WitnessValues does not know: the host parser decides
            }
            skip();
            if (pos < len && t[pos] == ',') { pos++; continue; }
            break;
        }
        if (!expect('}') || !have || !have_type || vs == ve) return false; // both members are required (the host parser reports the error)
        const int id = name_id(tab, t, ns, ne);
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
    return pos == len && k == nq && seen == (1u << tab.n_names) - 1;
}

// ---- SWAR helpers (4 text bytes per 32-bit word, little endian: byte 0 = first character) ---------------------------
// 0x80 in every byte of x that is zero (exact, no cross-byte borrow)
__device__ __forceinline__ uint32_t zero_bytes(uint32_t x) { return ~(((x & 0x7f7f7f7fu) + 0x7f7f7f7fu) | x | 0x7f7f7f7fu); }
// bit 0 of each byte (b0, b1, b2, b3) -> the 4-bit value b0 | b1 << 1 | b2 << 2 | b3 << 3
__device__ __forceinline__ uint32_t gather4(uint32_t y) { return ((y & 0x01010101u) * 0x01020408u) >> 24; }
__device__ __forceinline__ uint32_t word_of(const uint4 &v, int k) { return k == 0 ? v.x : k == 1 ? v.y : k == 2 ? v.z : v.w; }
__device__ __forceinline__ bool is_numchar(uint8_t c) { return (c >= '0' && c <= '9') || (c >= 'a' && c <= 'f') || c == 'x'; }

// bits of the class code in the lookup table
enum : uint8_t { K_NUM = 1, K_STR = 2, K_LC = 4, K_BAD = 8 };

// A witness file seen as aligned 16-byte blocks (the text is read with 128-bit loads only; `head` = offset of its first byte in block 0)
struct Blocks {
    const uint4 *base;
    uint32_t head, nblk, len;
    uint32_t safe_blk;  // blocks [0, safe_blk) lie entirely inside the text buffer; the one block that may straddle its end is read byte by byte
    uint32_t safe_bytes; // bytes from `base` to the end of the text buffer (saturated)
    __device__ __noinline__ uint4 load_tail(uint32_t blk) const {
        uint32_t w[4] = {0, 0, 0, 0};
        const uint8_t *b = reinterpret_cast<const uint8_t *>(base + blk);
        for (uint32_t k = 0; k < 16 && blk * 16u + k < safe_bytes; k++) w[k >> 2] |= (uint32_t)b[k] << (8 * (k & 3));
        return make_uint4(w[0], w[1], w[2], w[3]);
    }
    __device__ __forceinline__ uint4 load(uint32_t blk) const {
        if (blk >= nblk) return make_uint4(0, 0, 0, 0);
        return blk < safe_blk ? __ldg(base + blk) : load_tail(blk);
    }
    // 16-bit mask of the bytes of block `blk` whose file position lies in [lo, hi)
    __device__ __forceinline__ uint32_t valid(uint32_t blk, uint32_t lo, uint32_t hi) const {
        const int rel0 = (int)(blk * 16u) - (int)head;
        const int a = max(0, (int)lo - rel0), b = min(16, (int)hi - rel0);
        return a < b ? ((1u << b) - 1u) & ~((1u << a) - 1u) : 0u;
    }
};

constexpr int LEX_WARPS = 4;

// Kernel 1: one warp per witness.  Phase 0 finds the JSON strings (quote positions; a backslash anywhere hands the file to the host parser),
// lane 0 walks the JSON level, phase A streams every value through the warp 512 bytes at a time: bytes are classified through a lookup
// table, token starts are numbered with a warp scan, every token is compared with the skeleton and the position of every integer
// literal is recorded for kernel 2.
__global__ void __launch_bounds__(32 * LEX_WARPS) wit_lex_kernel(WitParams p) {
    __shared__ uint8_t s_lut[256];
    __shared__ uint32_t s_q[LEX_WARPS][WIT_MAXQ];
    __shared__ uint32_t s_span[LEX_WARPS][3 * WIT_MAX_NAMES];
    const uint32_t wib = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (uint32_t c = threadIdx.x; c < 256; c += blockDim.x) {
        uint8_t k = K_BAD;
        if (is_ws((uint8_t)c)) k = 0;
        else if (is_numchar((uint8_t)c)) k = K_NUM;
        else if (c == '(' || c == ')' || c == '[' || c == ']' || c == ',' || c == 'l') k = K_STR;
        else if (c == 'i' || c == 's' || c == 't' || c == '!') k = K_LC;
        s_lut[c] = k;
    }
    __syncthreads();
    const uint32_t i = blockIdx.x * LEX_WARPS + wib;
    if (i >= p.n) return;
    const uint64_t b0 = p.offsets[i], b1 = p.offsets[i + 1];
    if (b1 <= b0 || b1 - b0 >= 0x40000000ull) {
        if (lane == 0) p.flags[i] = SSYM_WIT_SLOW;
        return;
    }
    const uint8_t *t = p.text + b0;
    Blocks B;
    B.len = (uint32_t)(b1 - b0);
    B.head = (uint32_t)(reinterpret_cast<uintptr_t>(t) & 15u);
    B.base = reinterpret_cast<const uint4 *>(t - B.head);
    B.nblk = (B.head + B.len + 15u) / 16u;
    { // the buffer holds p.offsets[p.n] bytes: nothing behind them is read (the last block of the last witness usually straddles the end)
        const uint64_t room = p.offsets[p.n] - b0 + B.head;
        B.safe_bytes = room > 0xffffffffull ? 0xffffffffu : (uint32_t)room;
        B.safe_blk = B.safe_bytes / 16u;
    }
    const uint32_t FULL = 0xffffffffu;
    bool bad = false;

    // ---- phase 0: quote positions (in order), no backslashes ----
    uint32_t nq = 0;
    for (uint32_t blk0 = 0; blk0 < B.nblk; blk0 += 32) {
        const uint32_t blk = blk0 + lane;
        const uint4 v = B.load(blk);
        uint32_t zq[4], zb[4], any = 0;
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const uint32_t w = word_of(v, k);
            zq[k] = zero_bytes(w ^ 0x22222222u); // '"'
            zb[k] = zero_bytes(w ^ 0x5c5c5c5cu); // backslash
            any |= zq[k] | zb[k];
        }
        if (__any_sync(FULL, any != 0)) {
            const uint32_t vm = B.valid(blk, 0, B.len);
            uint32_t qm = 0, bm = 0;
#pragma unroll
            for (int k = 0; k < 4; k++) { qm |= gather4(zq[k] >> 7) << (4 * k); bm |= gather4(zb[k] >> 7) << (4 * k); }
            qm &= vm;
            if (bm & vm) bad = true;
            const uint32_t cnt = __popc(qm);
            uint32_t inc = cnt;
#pragma unroll
            for (int off = 1; off < 32; off <<= 1) {
                const uint32_t x = __shfl_up_sync(FULL, inc, off);
                if (lane >= (uint32_t)off) inc += x;
            }
            uint32_t at = nq + inc - cnt;
            const int rel0 = (int)(blk * 16u) - (int)B.head;
            for (uint32_t m = qm; m; m &= m - 1, at++)
                if (at < WIT_MAXQ) s_q[wib][at] = (uint32_t)(rel0 + __ffs(m) - 1);
            nq += __shfl_sync(FULL, inc, 31);
        }
    }
    bad = __any_sync(FULL, bad) || nq > WIT_MAXQ || (nq & 1u);
    __syncwarp();
    if (!bad && lane == 0) {
        uint32_t vs[WIT_MAX_NAMES], ve[WIT_MAX_NAMES];
        const int NN = (int)p.tab.n_names;
        if (!json_walk(p.tab, t, B.len, s_q[wib], nq, vs, ve)) {
            bad = true;
        } else { // spans in file order
            uint32_t order[WIT_MAX_NAMES];
            for (int a = 0; a < NN; a++) order[a] = a;
            for (int a = 1; a < NN; a++)
                for (int b = a; b > 0 && vs[order[b]] < vs[order[b - 1]]; b--) { const uint32_t x = order[b]; order[b] = order[b - 1]; order[b - 1] = x; }
            for (int a = 0; a < NN; a++) { s_span[wib][3 * a] = vs[order[a]]; s_span[wib][3 * a + 1] = ve[order[a]]; s_span[wib][3 * a + 2] = order[a]; }
        }
    }
    bad = __any_sync(FULL, bad);
    __syncwarp();
    if (bad) {
        if (lane == 0) p.flags[i] = SSYM_WIT_SLOW;
        return;
    }

    // ---- phase A: the six values, in file order ----
    uint32_t *numpos = p.numpos + (size_t)i * p.total_slots;
    uint32_t n_l = 0, n_lc = 0; // `l` tokens seen (each checked to start "list!") / i s t ! characters seen: must be 4 per `l`
    for (int sidx = 0; sidx < (int)p.tab.n_names; sidx++) {
        const uint32_t ss = s_span[wib][3 * sidx], se = s_span[wib][3 * sidx + 1], name = s_span[wib][3 * sidx + 2];
        const uint8_t *skel = p.tab.skel + p.tab.skel_off[name];
        const uint32_t skel_len = p.tab.skel_len[name], slot_cnt = p.tab.slot_cnt[name], slot_off = p.tab.slot_off[name];
        uint32_t tokbase = 0, numbase = 0, carry = 0;
        const uint32_t fb = (B.head + ss) / 16u, lb = (B.head + se - 1u) / 16u;
        for (uint32_t blk0 = fb; blk0 <= lb; blk0 += 32) {
            const uint32_t blk = blk0 + lane;
            const uint4 v = blk <= lb ? B.load(blk) : make_uint4(0, 0, 0, 0);
            const uint32_t vm = blk <= lb ? B.valid(blk, ss, se) : 0u;
            uint32_t numm = 0, strm = 0, lcm = 0, badm = 0;
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const uint32_t w = word_of(v, k);
                const uint32_t kw = (uint32_t)s_lut[w & 0xff] | (uint32_t)s_lut[(w >> 8) & 0xff] << 8 | (uint32_t)s_lut[(w >> 16) & 0xff] << 16 |
                                    (uint32_t)s_lut[w >> 24] << 24;
                numm |= gather4(kw) << (4 * k);
                strm |= gather4(kw >> 1) << (4 * k);
                lcm |= gather4(kw >> 2) << (4 * k);
                badm |= gather4(kw >> 3) << (4 * k);
            }
            numm &= vm; strm &= vm; lcm &= vm;
            if (badm & vm) bad = true;
            n_lc += __popc(lcm);
            uint32_t up = __shfl_up_sync(FULL, numm >> 15, 1);
            if (lane == 0) up = carry;
            const uint32_t nstart = numm & ~((numm << 1) | up) & 0xffffu;
            const uint32_t tstart = nstart | strm;
            const uint32_t cnt = __popc(tstart) | (__popc(nstart) << 16);
            uint32_t inc = cnt;
#pragma unroll
            for (int off = 1; off < 32; off <<= 1) {
                const uint32_t x = __shfl_up_sync(FULL, inc, off);
                if (lane >= (uint32_t)off) inc += x;
            }
            const uint32_t total = __shfl_sync(FULL, inc, 31);
            carry = __shfl_sync(FULL, numm >> 15, 31);
            uint32_t ord = tokbase + ((inc - cnt) & 0xffffu), nidx = numbase + ((inc - cnt) >> 16);
            const int rel0 = (int)(blk * 16u) - (int)B.head;
            for (uint32_t m = tstart; m; m &= m - 1, ord++) {
                const int j = __ffs(m) - 1;
                const uint8_t ch = (uint8_t)(word_of(v, j >> 2) >> (8 * (j & 3)));
                const bool isnum = (nstart >> j) & 1u;
                const uint8_t act = isnum ? (uint8_t)'N' : ch == 'l' ? (uint8_t)'L' : ch;
                if (ord >= skel_len || __ldg(skel + ord) != act) bad = true;
                if (isnum) {
                    if (nidx < slot_cnt) numpos[slot_off + nidx] = (uint32_t)(rel0 + j);
                    nidx++;
                } else if (ch == 'l') {
                    const uint32_t at = (uint32_t)(rel0 + j);
                    n_l++;
                    if (!(at + 4 < se && t[at + 1] == 'i' && t[at + 2] == 's' && t[at + 3] == 't' && t[at + 4] == '!')) bad = true;
                }
            }
            tokbase += total & 0xffffu;
            numbase += total >> 16;
        }
        if (tokbase != skel_len || numbase != slot_cnt) bad = true;
    }
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) {
        n_l += __shfl_xor_sync(FULL, n_l, off);
        n_lc += __shfl_xor_sync(FULL, n_lc, off);
    }
    if (n_lc != 4u * n_l) bad = true;
    bad = __any_sync(FULL, bad);
    if (lane == 0) p.flags[i] = bad ? SSYM_WIT_SLOW : SSYM_WIT_OK;
}

// Kernel 2: one thread per integer literal.  A 64-digit hex literal for a u256 slot (a digest: 87 % of the text) is converted 4 characters
// per operation; everything else goes through the general loop.
__device__ __forceinline__ uint32_t hex8(uint32_t w0, uint32_t w1) { // 8 validated lower-case hex characters -> their 32-bit value
    const uint32_t n0 = (w0 & 0x0f0f0f0fu) + 9u * ((w0 >> 6) & 0x01010101u), n1 = (w1 & 0x0f0f0f0fu) + 9u * ((w1 >> 6) & 0x01010101u);
    const uint32_t x0 = ((n0 << 4) | (n0 >> 8)) & 0x00ff00ffu, x1 = ((n1 << 4) | (n1 >> 8)) & 0x00ff00ffu;
    return __byte_perm(x0, x1, 0x0246);
}
__device__ __forceinline__ bool all_hex(uint32_t w) { // every byte in [0-9a-f]
    const uint32_t H = 0x80808080u;
    if (w & H) return false;
    const uint32_t dig = ((w + 0x50505050u) & ~(w + 0x46464646u)) & H; // >= '0' and not >= ':'
    const uint32_t let = ((w + 0x1f1f1f1fu) & ~(w + 0x19191919u)) & H; // >= 'a' and not >= 'g'
    return (dig | let) == H;
}

__device__ bool parse_number(const uint8_t *s, const uint8_t *text_end, uint32_t slot, uint32_t *rec) {
    const uint32_t off = slot & 0x0fffffffu, kind = slot >> 28;
    const uint32_t kw = kind == WIT_KIND_U32 ? 1u : kind == WIT_KIND_U64 ? 2u : 8u;
    if (s[0] == '0' && s[1] == 'x') {
        const uint8_t *d = s + 2;
        const uint32_t sh = (uint32_t)(reinterpret_cast<uintptr_t>(d) & 3u);
        if (kw == 8 && d - sh + 68 <= text_end) { // 64 digits and a terminator?  (17 aligned words, all inside the text buffer)
            const uint32_t *a = reinterpret_cast<const uint32_t *>(d - sh);
            uint32_t A[17], W[16];
#pragma unroll
            for (int k = 0; k < 17; k++) A[k] = __ldg(a + k);
            bool ok = true;
#pragma unroll
            for (int k = 0; k < 16; k++) {
                W[k] = __funnelshift_r(A[k], A[k + 1], 8 * sh);
                ok = ok && all_hex(W[k]);
            }
            if (ok && !is_numchar((uint8_t)(A[16] >> (8 * sh)))) {
#pragma unroll
                for (int k = 0; k < 8; k++) rec[off + k] = hex8(W[2 * k], W[2 * k + 1]);
                return true;
            }
        }
        uint32_t n = 0;
        while (is_numchar(d[n])) n++;
        if (n == 0 || n > 64) return false;
        uint32_t acc = 0, rem = n;
        for (uint32_t j = 0; j < n; j++) {
            const uint8_t c = d[j];
            uint32_t v;
            if (c >= '0' && c <= '9') v = c - '0';
            else if (c >= 'a' && c <= 'f') v = c - 'a' + 10;
            else return false; // a second 'x'
            acc = (acc << 4) | v;
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
    if (kw == 8) { // decimal u256 (stark101's digests): 9 digits per multiply-add over the 8 words
        uint32_t n = 0;
        while (is_numchar(s[n])) n++;
        if (n > 78) return false;
        uint32_t w[8] = {0, 0, 0, 0, 0, 0, 0, 0}; // little-endian words
        uint32_t j = 0;
        while (j < n) {
            const uint32_t take = (n - j) % 9u ? (n - j) % 9u : 9u;
            uint32_t chunk = 0, mul = 1;
            for (uint32_t k = 0; k < take; k++, j++) {
                const uint8_t c = s[j];
                if (c < '0' || c > '9') return false;
                chunk = chunk * 10u + (c - '0');
                mul *= 10u;
            }
            uint64_t carry = chunk;
#pragma unroll
            for (int k = 0; k < 8; k++) {
                const uint64_t x = (uint64_t)w[k] * mul + carry;
                w[k] = (uint32_t)x;
                carry = x >> 32;
            }
            if (carry) return false; // exceeds 256 bits
        }
#pragma unroll
        for (int k = 0; k < 8; k++) rec[off + k] = w[7 - k];
        return true;
    }
    uint64_t v = 0;
    for (const uint8_t *q = s; is_numchar(*q); q++) {
        const uint8_t c = *q;
        if (c < '0' || c > '9') return false;
        const uint32_t dgt = c - '0';
        if (v > (0xffffffffffffffffull - dgt) / 10u) return false; // above 64 bits: host parser
        v = v * 10u + dgt;
    }
    if (kw == 1) {
        if (v >> 32) return false;
        rec[off] = (uint32_t)v;
    } else {
        rec[off] = (uint32_t)(v >> 32);
        rec[off + 1] = (uint32_t)v;
    }
    return true;
}

__global__ void __launch_bounds__(256) wit_numbers_kernel(WitParams p) {
    const uint64_t g = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t i = (uint32_t)(g / p.total_slots), k = (uint32_t)(g % p.total_slots);
    if (i >= p.n || p.flags[i] != SSYM_WIT_OK) return;
    const uint8_t *t = p.text + p.offsets[i];
    if (!parse_number(t + p.numpos[(size_t)i * p.total_slots + k], p.text + p.offsets[p.n], __ldg(p.tab.slots + k), p.packed + (size_t)i * p.stride_words))
        p.flags[i] = SSYM_WIT_SLOW;
}

// packed[i] = template for every i (the words of a record that are not literals: lengths, counts)
__global__ void wit_fill_template_kernel(uint32_t *packed, const uint32_t *templ, uint32_t stride_words, uint64_t total_words) {
    for (uint64_t g = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; g < total_words; g += (uint64_t)gridDim.x * blockDim.x) packed[g] = templ[g % stride_words];
}
// status[idx[j]] = st[j], accept bit idx[j] = (st[j] == 0)
__global__ void wit_scatter_status_kernel(const uint32_t *idx, const uint32_t *st, uint32_t m, uint32_t *status, uint32_t *accept_bits) {
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= m) return;
    const uint32_t i = idx[j];
    if (status) status[i] = st[j];
    if (st[j] == 0) atomicOr(&accept_bits[i >> 5], 1u << (i & 31));
    else atomicAnd(&accept_bits[i >> 5], ~(1u << (i & 31)));
}

__global__ void wit_apply_flags_kernel(const uint32_t *flags, uint32_t *status, uint32_t *accept_bits, uint32_t n) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n || flags[i] == SSYM_WIT_OK) return;
    if (status) status[i] |= SSYM_ST_SHAPE;
    atomicAnd(&accept_bits[i >> 5], ~(1u << (i & 31)));
}

} // namespace

void wit_set_names(WitTables &t, const char *const *names, uint32_t n) {
    t.n_names = n;
    for (uint32_t k = 0; k < n; k++) {
        uint32_t len = 0;
        for (; names[k][len] && len < WIT_NAME_CHARS; len++) t.name[k][len] = names[k][len];
        t.name_len[k] = (uint8_t)len;
    }
}

void launch_wit_pack(const WitParams &p, cudaStream_t s) {
    if (!p.n) return;
    wit_lex_kernel<<<(p.n + LEX_WARPS - 1) / LEX_WARPS, 32 * LEX_WARPS, 0, s>>>(p);
    const uint64_t threads = (uint64_t)p.n * p.total_slots;
    wit_numbers_kernel<<<(uint32_t)((threads + 255) / 256), 256, 0, s>>>(p);
}
void launch_wit_fill_template(uint32_t *packed, const uint32_t *templ, uint32_t stride_words, size_t n, cudaStream_t s) {
    const uint64_t total = (uint64_t)n * stride_words;
    if (total) wit_fill_template_kernel<<<(uint32_t)std::min<uint64_t>((total + 255) / 256, 148 * 16), 256, 0, s>>>(packed, templ, stride_words, total);
}
void launch_wit_scatter_status(const uint32_t *idx, const uint32_t *st, uint32_t m, uint32_t *status, uint32_t *accept_bits, cudaStream_t s) {
    if (m) wit_scatter_status_kernel<<<(m + 127) / 128, 128, 0, s>>>(idx, st, m, status, accept_bits);
}
void launch_wit_apply_flags(const uint32_t *flags, uint32_t *status, uint32_t *accept_bits, uint32_t n, cudaStream_t s) {
    if (n) wit_apply_flags_kernel<<<(n + 255) / 256, 256, 0, s>>>(flags, status, accept_bits, n);
}

} // namespace ssym
