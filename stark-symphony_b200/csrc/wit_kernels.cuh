// SPDX-License-Identifier: MIT
//
// `.wit` ingestion on the GPU: the JSON text `simfony run --witness` reads
// (simfony-cli/src/main.rs:77-81, emitted by stwo-verifier/scripts/generate_wit.py:106-245) is tokenised and packed
// into the wire format of include/ssym.h by two kernels: a lexer (one warp per witness, 128-bit coalesced loads, table-driven byte
// classes, warp scans for the token numbering) and a literal converter (one thread per integer literal).
//
// The value grammar is fixed by the program's witness types (stwo-verifier/src/main.simf:9-25), so for a given
// configuration the token sequence of every witness value is known in advance up to whitespace: the host builds that
// "skeleton" (one byte per token: ( ) [ ] , L = `list!`  N = integer literal) plus, for the k-th integer literal, the
// packed word it lands in and its width (u32 / u64 / u256).  The kernel checks the text against the skeleton and scatters
// the literals.  Anything the fast path does not cover (JSON escapes, `_` digit separators, upper-case hex, redundant parentheses, trailing commas, a list of another length, any malformed text) sets the witness's
// flag to WIT_SLOW and the host re-parses exactly that witness with the full grammar (csrc/witness.cpp), so the result is
// always the one ssym_stwo_pack_wit gives.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/ssym.h"

namespace ssym {

enum { WIT_NAMES = 6 /* stwo-verifier/src/main.simf:9-25 */, WIT_MAX_NAMES = 8, WIT_NAME_CHARS = 20 };
enum { WIT_KIND_U32 = 0, WIT_KIND_U64 = 1, WIT_KIND_U256 = 2 };

struct WitTables {
    const uint8_t *skel;   // token skeletons of the six witness values, concatenated
    const uint32_t *slots; // per integer literal: packed word offset | kind << 28
    uint32_t skel_off[WIT_MAX_NAMES], skel_len[WIT_MAX_NAMES];
    uint32_t slot_off[WIT_MAX_NAMES], slot_cnt[WIT_MAX_NAMES];
    // the program's witness names (main.simf): every one must be present exactly once, in any order
    uint32_t n_names;
    uint8_t name_len[WIT_MAX_NAMES];
    char name[WIT_MAX_NAMES][WIT_NAME_CHARS];
};
void wit_set_names(WitTables &t, const char *const *names, uint32_t n);

struct WitParams {
    const uint8_t *text;     // concatenated witness texts (device)
    const uint64_t *offsets; // n + 1 byte offsets into text (device)
    uint32_t n;
    uint32_t stride_words;
    uint32_t *packed;        // n * stride_words, pre-filled by the caller with everything that is not a literal (zeros for Stwo)
    uint32_t *flags;         // n: SSYM_WIT_OK or SSYM_WIT_SLOW (internal: host re-parse)
    uint32_t *numpos;        // scratch, n * total_slots: file position of every integer literal (kernel 1 -> kernel 2)
    uint32_t total_slots;    // integer literals per witness (sum of slot_cnt)
    WitTables tab;
};

void launch_wit_pack(const WitParams &p, cudaStream_t s);
void launch_wit_fill_template(uint32_t *packed, const uint32_t *templ, uint32_t stride_words, size_t n, cudaStream_t s);
void launch_wit_scatter_status(const uint32_t *idx, const uint32_t *st, uint32_t m, uint32_t *status, uint32_t *accept_bits, cudaStream_t s);
// status[i] |= SSYM_ST_SHAPE and accept bit i cleared for every witness with a non-zero flag
void launch_wit_apply_flags(const uint32_t *flags, uint32_t *status, uint32_t *accept_bits, uint32_t n, cudaStream_t s);

} // namespace ssym
