// SPDX-License-Identifier: MIT
//
// Witness ingestion: the `.wit` JSON that `simfony run --witness` reads (simfony-cli/src/main.rs:77-81:
// NAME -> {"value": <SimplicityHL value text>, "type": <type text>}) parsed with the value grammar the
// reference's generators emit (stwo-verifier/scripts/generate_wit.py:32-35,139-243,
// stark101/scripts/generate_wit.py:7-30): decimal / 0x integers, tuples, arrays, `list![...]`, and packed
// into the binary wire format of include/ssym.h.  Shapes are derived from the program's witness types
// (stwo-verifier/src/main.simf:9-25, stark101/src/main.simf:12-20), never from the "type" strings
// (stark101's FRI_LAYERS type string is malformed upstream).  Field values are NOT canonicalised.
#include <cstdint>
#include <cstring>
#include <map>
#include <memory>
#include <new>
#include <string>
#include <vector>

#include "../../include/ssym.h"

namespace {

struct ParseError {
    std::string msg;
};

struct U256 {
    uint32_t w[8]; // w[0] most significant
};

struct Val {
    enum Kind { INT, TUPLE, ARRAY, LIST } kind = INT;
    U256 num{};
    std::vector<Val> items;
};

// ---- value grammar ----------------------------------------------------------------------------------
// Witness text is untrusted: nesting is bounded (the deepest value of either program nests 7 levels), so a hostile text cannot overflow the stack.
constexpr int MAX_NESTING = 64;
struct DepthGuard {
    int &d;
    explicit DepthGuard(int &d_) : d(d_) { if (++d > MAX_NESTING) throw ParseError{"nesting too deep"}; }
    ~DepthGuard() { --d; }
};

struct ValueParser {
    const char *s;
    size_t n, pos = 0;
    int depth = 0;
    ValueParser(const char *s_, size_t n_) : s(s_), n(n_) {}
    void ws() {
        while (pos < n && (s[pos] == ' ' || s[pos] == '\n' || s[pos] == '\t' || s[pos] == '\r')) pos++;
    }
    bool eat(char c) {
        ws();
        if (pos < n && s[pos] == c) { pos++; return true; }
        return false;
    }
    static void mul_add(U256 &v, uint32_t mul, uint32_t add) {
        uint64_t carry = add;
        for (int i = 7; i >= 0; i--) {
            uint64_t t = (uint64_t)v.w[i] * mul + carry;
            v.w[i] = (uint32_t)t;
            carry = t >> 32;
        }
        if (carry) throw ParseError{"integer literal exceeds 256 bits"};
    }
    Val number() {
        Val v;
        v.kind = Val::INT;
        bool hex = false;
        if (pos + 1 < n && s[pos] == '0' && (s[pos + 1] == 'x' || s[pos + 1] == 'X')) { hex = true; pos += 2; }
        size_t digits = 0;
        for (; pos < n; pos++) {
            char c = s[pos];
            int d;
            if (c == '_') continue;
            if (c >= '0' && c <= '9') d = c - '0';
            else if (hex && c >= 'a' && c <= 'f') d = c - 'a' + 10;
            else if (hex && c >= 'A' && c <= 'F') d = c - 'A' + 10;
            else break;
            mul_add(v.num, hex ? 16 : 10, (uint32_t)d);
            digits++;
        }
        if (!digits) throw ParseError{"empty integer literal"};
        return v;
    }
    std::vector<Val> seq(char close, bool *trailing_comma = nullptr) {
        std::vector<Val> items;
        if (eat(close)) return items;
        for (;;) {
            items.push_back(value());
            if (eat(',')) {
                if (eat(close)) { // trailing comma
                    if (trailing_comma) *trailing_comma = true;
                    return items;
                }
                continue;
            }
            if (eat(close)) return items;
            throw ParseError{"expected ',' or closing bracket"};
        }
    }
    Val value() {
        DepthGuard guard(depth);
        ws();
        if (pos >= n) throw ParseError{"unexpected end of value"};
        char c = s[pos];
        if (c == '(') {
            pos++;
            bool trailing = false;
            std::vector<Val> items = seq(')', &trailing);
            if (items.size() == 1) {
                // `(x)` is a parenthesised expression; `(x,)` is a 1-tuple, which is a value of none of the programs' witness types
                if (trailing) throw ParseError{"1-tuple is not a value of any witness type"};
                return items[0];
            }
            Val v;
            v.kind = Val::TUPLE;
            v.items = std::move(items);
            return v;
        }
        if (c == '[') {
            pos++;
            Val v;
            v.kind = Val::ARRAY;
            v.items = seq(']');
            return v;
        }
        if (n - pos >= 5 && !strncmp(s + pos, "list!", 5)) {
            pos += 5;
            if (!eat('[')) throw ParseError{"list! must be followed by '['"};
            Val v;
            v.kind = Val::LIST;
            v.items = seq(']');
            return v;
        }
        if (c >= '0' && c <= '9') return number();
        throw ParseError{std::string("unexpected character '") + c + "' in value"};
    }
    Val parse() {
        Val v = value();
        ws();
        if (pos != n) throw ParseError{"trailing characters after value"};
        return v;
    }
};

// ---- minimal JSON (object of objects with string members) ----------------------------------------------
struct Json {
    const char *s;
    size_t n, pos = 0;
    int depth = 0;
    Json(const char *s_, size_t n_) : s(s_), n(n_) {}
    void ws() {
        while (pos < n && (s[pos] == ' ' || s[pos] == '\n' || s[pos] == '\t' || s[pos] == '\r')) pos++;
    }
    void expect(char c) {
        ws();
        if (pos >= n || s[pos] != c) throw ParseError{std::string("JSON: expected '") + c + "'"};
        pos++;
    }
    bool peek(char c) {
        ws();
        return pos < n && s[pos] == c;
    }
    std::string str() {
        expect('"');
        std::string out;
        while (pos < n && s[pos] != '"') {
            char c = s[pos++];
            if (c == '\\') {
                if (pos >= n) break;
                char e = s[pos++];
                switch (e) {
                case 'n': out += '\n'; break;
                case 't': out += '\t'; break;
                case 'r': out += '\r'; break;
                case 'b': out += '\b'; break;
                case 'f': out += '\f'; break;
                case 'u': { // only ASCII escapes can occur in a value text
                    if (pos + 4 > n) throw ParseError{"JSON: bad \\u escape"};
                    unsigned cp = 0;
                    for (int i = 0; i < 4; i++) {
                        const char h = s[pos++];
                        const bool dec = h >= '0' && h <= '9', hexl = (h | 32) >= 'a' && (h | 32) <= 'f';
                        if (!dec && !hexl) throw ParseError{"JSON: bad \\u escape"};
                        cp = cp * 16 + (dec ? h - '0' : (h | 32) - 'a' + 10);
                    }
                    if (cp > 0x7f) throw ParseError{"JSON: non-ASCII escape in witness"};
                    out += (char)cp;
                    break;
                }
                default: out += e;
                }
            } else {
                out += c;
            }
        }
        if (pos >= n) throw ParseError{"JSON: unterminated string"};
        pos++;
        return out;
    }
    void skip_value() {
        DepthGuard guard(depth);
        ws();
        if (pos >= n) throw ParseError{"JSON: unexpected end"};
        char c = s[pos];
        if (c == '"') { str(); return; }
        if (c == '{' || c == '[') {
            char close = c == '{' ? '}' : ']';
            pos++;
            if (peek(close)) { pos++; return; }
            for (;;) {
                if (c == '{') { str(); expect(':'); }
                skip_value();
                ws();
                if (pos < n && s[pos] == ',') { pos++; continue; }
                expect(close);
                return;
            }
        }
        while (pos < n && s[pos] != ',' && s[pos] != '}' && s[pos] != ']') pos++; // number / literal
    }
    // top level: { NAME: { "value": "...", ... }, ... }
    std::map<std::string, std::string> witness_values() {
        std::map<std::string, std::string> out;
        expect('{');
        if (peek('}')) { pos++; return out; }
        for (;;) {
            std::string name = str();
            expect(':');
            expect('{');
            // serde's WitnessValues (simfony-cli/src/main.rs:77-81): both members are strings and both are required; a repeated member or
            // witness name is an error there, not "last one wins"
            bool have = false, have_type = false;
            if (out.count(name)) throw ParseError{"duplicate witness " + name};
            if (!peek('}')) {
                for (;;) {
                    std::string key = str();
                    expect(':');
                    if (key == "value") {
                        if (have) throw ParseError{"witness " + name + " has two \"value\" members"};
                        out[name] = str();
                        have = true;
                    } else if (key == "type") {
                        if (have_type) throw ParseError{"witness " + name + " has two \"type\" members"};
                        str(); // shapes come from the program, not from this text (stark101's FRI_LAYERS type string is malformed upstream)
                        have_type = true;
                    } else {
                        skip_value();
                    }
                    ws();
                    if (pos < n && s[pos] == ',') { pos++; continue; }
                    break;
                }
            }
            expect('}');
            if (!have) throw ParseError{"witness " + name + " has no \"value\""};
            if (!have_type) throw ParseError{"witness " + name + " has no \"type\""};
            ws();
            if (pos < n && s[pos] == ',') { pos++; continue; }
            expect('}');
            break;
        }
        ws();
        if (pos != n) throw ParseError{"JSON: trailing characters after the witness object"};
        return out;
    }
};

// ---- typed access ------------------------------------------------------------------------------------
const Val &tuple_of(const Val &v, size_t n, const char *what) {
    if (v.kind != Val::TUPLE || v.items.size() != n) throw ParseError{std::string("expected ") + std::to_string(n) + "-tuple for " + what};
    return v;
}
const Val &array_of(const Val &v, size_t n, const char *what) {
    if (v.kind != Val::ARRAY || v.items.size() != n) throw ParseError{std::string("expected array of ") + std::to_string(n) + " for " + what};
    return v;
}
const Val &list32(const Val &v, const char *what) {
    if (v.kind != Val::LIST || v.items.size() >= 32) throw ParseError{std::string("expected List<_, 32> for ") + what};
    return v;
}
uint32_t u32_of(const Val &v, const char *what) {
    if (v.kind != Val::INT) throw ParseError{std::string("expected u32 for ") + what};
    for (int i = 0; i < 7; i++)
        if (v.num.w[i]) throw ParseError{std::string("value does not fit u32: ") + what};
    return v.num.w[7];
}
void u64_of(const Val &v, uint32_t &hi, uint32_t &lo, const char *what) {
    if (v.kind != Val::INT) throw ParseError{std::string("expected u64 for ") + what};
    for (int i = 0; i < 6; i++)
        if (v.num.w[i]) throw ParseError{std::string("value does not fit u64: ") + what};
    hi = v.num.w[6];
    lo = v.num.w[7];
}
void put_u256(uint32_t *dst, const Val &v, const char *what) {
    if (v.kind != Val::INT) throw ParseError{std::string("expected u256 for ") + what};
    memcpy(dst, v.num.w, 32);
}
void put_qm31(uint32_t *dst, const Val &v, const char *what) {
    const Val &t = tuple_of(v, 2, what);
    const Val &re = tuple_of(t.items[0], 2, what), &im = tuple_of(t.items[1], 2, what);
    dst[0] = u32_of(re.items[0], what);
    dst[1] = u32_of(re.items[1], what);
    dst[2] = u32_of(im.items[0], what);
    dst[3] = u32_of(im.items[1], what);
}

const Val &need(const std::map<std::string, Val> &w, const char *name) {
    auto it = w.find(name);
    if (it == w.end()) throw ParseError{std::string("missing witness ") + name};
    return it->second;
}

std::map<std::string, Val> parse_wit(const char *text, size_t len) {
    Json j(text, len);
    std::map<std::string, Val> out;
    for (auto &kv : j.witness_values()) out[kv.first] = ValueParser(kv.second.data(), kv.second.size()).parse();
    return out;
}

} // namespace

extern "C" int ssym_stwo_pack_wit(const ssym_stwo_config_t *cfg, const char *json_text, size_t len, uint32_t *out, int *shape_reject) {
    if (!cfg || !json_text || !out) return SSYM_ERR_USAGE;
    ssym_stwo_layout_t lo;
    int rc = ssym_stwo_layout(cfg, &lo);
    if (rc) return rc;
    const uint32_t Q = cfg->n_queries, L = cfg->n_fri_layers, G = cfg->lde_log, C = SSYM_STWO_COLUMNS(cfg), QV = C + SSYM_NUM_CP_PARTITIONS;
    memset(out, 0, (size_t)lo.stride_words * 4);
    bool reject = false;
    try {
        std::map<std::string, Val> w = parse_wit(json_text, len);
        const Val &com = tuple_of(need(w, "COMMITMENTS"), 3, "COMMITMENTS"); // evals/commit.simf:16
        for (int i = 0; i < 3; i++) put_u256(out + lo.off_commit + 8 * i, com.items[i], "COMMITMENTS");
        const Val &oods = tuple_of(need(w, "OODS_EVALS"), 2, "OODS_EVALS"); // deep/oods.simf:20
        const Val &ot = array_of(oods.items[0], C, "OODS trace evals");
        for (uint32_t i = 0; i < C; i++) put_qm31(out + lo.off_oods_trace + 4 * i, array_of(ot.items[i], 1, "ColEvalsQM31").items[0], "OODS trace eval");
        const Val &oc = array_of(oods.items[1], SSYM_NUM_CP_PARTITIONS, "OODS CP evals");
        for (int i = 0; i < SSYM_NUM_CP_PARTITIONS; i++) put_qm31(out + lo.off_oods_cp + 4 * i, oc.items[i], "OODS CP eval");
        const Val &fc = tuple_of(need(w, "FRI_COMMITMENTS"), 3, "FRI_COMMITMENTS"); // fri/commit.simf:19-23
        put_u256(out + lo.off_fri_first_root, fc.items[0], "FRI first root");
        const Val &inner = array_of(fc.items[1], L, "FRI inner roots");
        for (uint32_t i = 0; i < L; i++) put_u256(out + lo.off_fri_inner_root + 8 * i, inner.items[i], "FRI inner root");
        put_qm31(out + lo.off_last_coeff, fc.items[2], "FRI last layer");
        u64_of(need(w, "POW_NONCE"), out[lo.off_pow_nonce], out[lo.off_pow_nonce + 1], "POW_NONCE");

        const Val &dec = array_of(need(w, "DECOMMITMENTS"), Q, "DECOMMITMENTS"); // evals/verify.simf:20-36
        for (uint32_t q = 0; q < Q; q++) {
            const Val &d = tuple_of(dec.items[q], 2, "Decommitment");
            const Val &td = tuple_of(d.items[0], 2, "TraceDecommitment"), &cd = tuple_of(d.items[1], 2, "CpDecommitment");
            const Val &tv = array_of(td.items[0], C, "TraceEvalsM31");
            for (uint32_t i = 0; i < C; i++) out[lo.off_qvals + QV * q + i] = u32_of(array_of(tv.items[i], 1, "ColEvalsM31").items[0], "trace eval");
            const Val &cv = array_of(cd.items[0], SSYM_NUM_CP_PARTITIONS, "CPEvalM31");
            for (int i = 0; i < SSYM_NUM_CP_PARTITIONS; i++) out[lo.off_qvals + QV * q + C + i] = u32_of(cv.items[i], "cp eval");
            const Val *proofs[2] = {&list32(td.items[1], "trace MerkleProof32"), &list32(cd.items[1], "cp MerkleProof32")};
            const uint32_t offs[2] = {lo.off_trace_sib, lo.off_cp_sib};
            for (int t = 0; t < 2; t++) {
                if (proofs[t]->items.size() != G) { reject = true; continue; } // merkle.simf:42 cannot hold
                for (uint32_t k = 0; k < G; k++) put_u256(out + offs[t] + (q * G + k) * 8, proofs[t]->items[k], "Merkle sibling");
            }
        }
        const Val &fd = tuple_of(need(w, "FRI_DECOMMITMENTS"), 2, "FRI_DECOMMITMENTS"); // fri/verify.simf:15-21
        const Val &inner_d = array_of(fd.items[1], L, "FRI inner decommitments");
        for (uint32_t l = 0; l <= L; l++) {
            const Val &layer = array_of(l == 0 ? fd.items[0] : inner_d.items[l - 1], Q, "FriLayerDecommitment");
            const uint32_t n_sib = G - 1 - l;
            for (uint32_t q = 0; q < Q; q++) {
                const Val &item = tuple_of(layer.items[q], 2, "FriQueryDecommitment");
                put_qm31(out + lo.off_fri_wit + (l * Q + q) * 4, item.items[0], "FRI witness");
                const Val &proof = list32(item.items[1], "FRI MerkleProof32");
                if (proof.items.size() != n_sib) { reject = true; continue; }
                for (uint32_t k = 0; k < n_sib; k++) put_u256(out + lo.off_fri_sib[l] + (q * n_sib + k) * 8, proof.items[k], "FRI sibling");
            }
        }
    } catch (const ParseError &) {
        memset(out, 0, (size_t)lo.stride_words * 4);
        return SSYM_ERR_PARSE;
    } catch (const std::bad_alloc &) { // nothing may escape the extern "C" boundary: a hostile witness is a reject, not a crash
        memset(out, 0, (size_t)lo.stride_words * 4);
        return SSYM_ERR_NOMEM;
    } catch (...) {
        memset(out, 0, (size_t)lo.stride_words * 4);
        return SSYM_ERR_PARSE;
    }
    if (reject) memset(out, 0, (size_t)lo.stride_words * 4);
    if (shape_reject) *shape_reject = reject ? 1 : 0;
    return SSYM_OK;
}

extern "C" int ssym_s101_pack_wit(const char *json_text, size_t len, uint32_t *out, size_t *out_words) {
    if (!json_text || !out || !out_words) return SSYM_ERR_USAGE;
    try {
        std::map<std::string, Val> w = parse_wit(json_text, len);
        std::vector<uint32_t> rec(20, 0);
        const Val &evals = tuple_of(need(w, "P_EVALS"), 3, "P_EVALS");        // air.simf:24-27
        const Val &layers = list32(need(w, "FRI_LAYERS"), "FRI_LAYERS");      // fri.simf:49
        rec[1] = (uint32_t)layers.items.size();
        rec[5] = u32_of(need(w, "FRI_LAST_LAYER"), "FRI_LAST_LAYER");
        put_u256(rec.data() + 8, need(w, "P_MT_ROOT"), "P_MT_ROOT");
        for (int i = 0; i < 3; i++) {
            const Val &e = tuple_of(evals.items[i], 2, "Eval");
            rec[16 + i] = u32_of(e.items[0], "Eval value");
            rec[2 + i] = (uint32_t)list32(e.items[1], "Eval proof").items.size();
        }
        auto push_sibs = [&](const Val &proof) {
            for (const Val &s : proof.items) {
                size_t at = rec.size();
                rec.resize(at + 8);
                put_u256(rec.data() + at, s, "Merkle sibling");
            }
        };
        for (int i = 0; i < 3; i++) push_sibs(evals.items[i].items[1]);
        for (const Val &lv : layers.items) { // fri.simf:31
            const Val &l = tuple_of(lv, 6, "FriLayer");
            const Val &pa = list32(l.items[3], "cpa proof"), &pb = list32(l.items[5], "cpb proof");
            size_t at = rec.size();
            rec.resize(at + 16, 0);
            put_u256(rec.data() + at, l.items[0], "FRI layer root");
            rec[at + 8] = u32_of(l.items[1], "beta");
            rec[at + 9] = u32_of(l.items[2], "cpa");
            rec[at + 10] = u32_of(l.items[4], "cpb");
            rec[at + 11] = (uint32_t)pa.items.size();
            rec[at + 12] = (uint32_t)pb.items.size();
            push_sibs(pa);
            push_sibs(pb);
        }
        rec[0] = (uint32_t)rec.size();
        if (rec.size() > *out_words) {
            *out_words = rec.size();
            return SSYM_ERR_NOMEM;
        }
        memcpy(out, rec.data(), rec.size() * 4);
        *out_words = rec.size();
    } catch (const ParseError &) {
        return SSYM_ERR_PARSE;
    } catch (const std::bad_alloc &) {
        return SSYM_ERR_NOMEM;
    } catch (...) {
        return SSYM_ERR_PARSE;
    }
    return SSYM_OK;
}
