"""Witness / proof ingestion on the host: the reference's proof JSON -> `.wit` -> packed wire format.

Mirrors stwo-verifier/scripts/generate_wit.py:106-245 (`build_witness_from_json`) and
stark101/scripts/generate_wit.py:7-30 field for field, so the Python proof scripts of the reference can drive
the batched verifier; `.wit` text is parsed by the C++ parser behind the C-ABI (csrc/witness.cpp)."""
from __future__ import annotations

import ctypes as C
import json
from typing import Any, Dict, List, Optional, Sequence, Tuple

import numpy as np

from ._lib import SsymError, StwoConfig, check, load
from .verifier import stwo_layout


# ---- proof JSON -> witness dict (same names / value syntax as the reference emits) -----------------------
def _u256_hex(byte_list: Sequence[int]) -> str:
    if len(byte_list) != 32:
        raise SsymError("expected a 32-byte digest")
    return "0x" + bytes(byte_list).hex()


def _qm31(node: Any) -> Tuple[int, int, int, int]:
    x = node
    while isinstance(x, list) and len(x) == 1 and isinstance(x[0], list):
        x = x[0]
    (a, b), (c, d) = x
    return int(a), int(b), int(c), int(d)


def _qm31_str(q) -> str:
    return f"(({q[0]}, {q[1]}), ({q[2]}, {q[3]}))"


def _chunks(lst: list, n: int) -> List[list]:
    if n <= 0 or len(lst) % n:
        raise SsymError("list length must be divisible by the number of queries")
    k = len(lst) // n
    return [lst[i * k:(i + 1) * k] for i in range(n)]


def stwo_wit_from_proof_json(data: Dict[str, Any]) -> Dict[str, Dict[str, str]]:
    """stwo proof JSON (serde `StarkProof`) -> the six witnesses of stwo-verifier/src/main.simf:9-25."""
    n_queries = int(data.get("config", {}).get("fri_config", {}).get("n_queries", 1))
    commitments = "(" + ", ".join(_u256_hex(c) for c in data["commitments"][:3]) + ")"
    sampled = data["sampled_values"]
    trace_oods = [_qm31(col) for col in sampled[1]]
    cp_oods = [_qm31(p) for p in sampled[2]]
    oods = "([" + ", ".join("[" + _qm31_str(q) + "]" for q in trace_oods) + "], [" + ", ".join(_qm31_str(q) for q in cp_oods) + "])"
    dec = data["decommitments"]
    t_hash, c_hash = _chunks(dec[1]["hash_witness"], n_queries), _chunks(dec[2]["hash_witness"], n_queries)
    t_vals, c_vals = _chunks([int(x) for x in data["queried_values"][1]], n_queries), _chunks([int(x) for x in data["queried_values"][2]], n_queries)
    items = []
    for i in range(n_queries):
        tv = "[" + ", ".join(f"[{x}]" for x in t_vals[i]) + "]"
        cv = "[" + ", ".join(str(x) for x in c_vals[i]) + "]"
        tp = "list![" + ", ".join(_u256_hex(x) for x in t_hash[i]) + "]"
        cp = "list![" + ", ".join(_u256_hex(x) for x in c_hash[i]) + "]"
        items.append(f"(({tv}, {tp}), ({cv}, {cp}))")
    fri = data["fri_proof"]
    inner = fri.get("inner_layers", [])

    def layer_str(layer) -> str:
        wits = [_qm31(w) for w in layer["fri_witness"]]
        hashes = _chunks(layer["decommitment"]["hash_witness"], n_queries)
        return "[" + ", ".join(f"({_qm31_str(wits[i])}, list![" + ", ".join(_u256_hex(x) for x in hashes[i]) + "])" for i in range(n_queries)) + "]"

    coeffs = fri["last_layer_poly"]["coeffs"]
    if len(coeffs) != 1:
        raise SsymError("expected a degree-0 last layer")
    fri_commitments = f"({_u256_hex(fri['first_layer']['commitment'])}, [" + ", ".join(_u256_hex(l["commitment"]) for l in inner) + f"], {_qm31_str(_qm31(coeffs[0]))})"
    fri_decommitments = f"({layer_str(fri['first_layer'])}, [" + ", ".join(layer_str(l) for l in inner) + "])"
    vals = {
        "COMMITMENTS": commitments, "DECOMMITMENTS": "[" + ", ".join(items) + "]", "OODS_EVALS": oods,
        "FRI_COMMITMENTS": fri_commitments, "FRI_DECOMMITMENTS": fri_decommitments, "POW_NONCE": str(int(data.get("proof_of_work", 0))),
    }
    return {k: {"value": v, "type": ""} for k, v in vals.items()}


def stark101_wit_from_proof_json(proof: Dict[str, Any]) -> Dict[str, Dict[str, str]]:
    """stark101 proof JSON (scripts/fibsquare/prover.py:94-171) -> the four witnesses of stark101/src/main.simf:12-20."""
    p_evals = ", ".join(f"({x[0]}, list!{[int(s) for s in x[1]]})" for x in proof["evals"])
    layers = ", ".join(f"(({l[0]}, {l[1]}, {l[2]}, list!{[int(s) for s in l[3]]}, {l[4]}, list!{[int(s) for s in l[5]]}))" for l in proof["fri_layers"])
    vals = {"P_MT_ROOT": str(proof["p_mt_root"]), "P_EVALS": f"({p_evals})", "FRI_LAYERS": f"list![{layers}]", "FRI_LAST_LAYER": str(proof["fri_last_layer"])}
    return {k: {"value": v, "type": ""} for k, v in vals.items()}


# ---- `.wit` text -> packed (C++ parser behind the C-ABI) ----------------------------------------------------
def pack_stwo_wits(wit_texts: Sequence[str], cfg: StwoConfig) -> Tuple[np.ndarray, np.ndarray]:
    """Returns (packed [n * stride_words], bad [n] bool).  bad = ill-typed or ill-shaped witness: `simfony run` would
    refuse it, so it must be reported as rejected (its record is zero-filled)."""
    lib = load()
    lo = stwo_layout(cfg)
    n = len(wit_texts)
    packed = np.zeros(n * lo.stride_words, dtype=np.uint32)
    bad = np.zeros(n, dtype=bool)
    for i, text in enumerate(wit_texts):
        raw = text.encode() if isinstance(text, str) else text
        shape = C.c_int(0)
        rc = lib.ssym_stwo_pack_wit(C.byref(cfg), raw, len(raw), C.c_void_p(packed[i * lo.stride_words:].ctypes.data), C.byref(shape))
        bad[i] = rc != 0 or shape.value != 0
    return packed, bad


def concat_wit_texts(wit_texts: Sequence) -> Tuple[np.ndarray, np.ndarray]:
    """[text, ...] -> (uint8 array of the concatenated texts, uint64 offsets [n + 1]) for Verifier.stwo_*_wit_batch."""
    raws = [t.encode() if isinstance(t, str) else bytes(t) for t in wit_texts]
    offsets = np.zeros(len(raws) + 1, dtype=np.uint64)
    offsets[1:] = np.cumsum([len(r) for r in raws], dtype=np.uint64)
    return np.frombuffer(b"".join(raws), dtype=np.uint8).copy() if raws else np.zeros(0, dtype=np.uint8), offsets


def stwo_wit_from_packed(packed_one: np.ndarray, cfg: StwoConfig) -> Dict[str, Dict[str, str]]:
    """One packed proof -> the six witnesses in the value syntax of stwo-verifier/scripts/generate_wit.py:139-243 (the inverse of the
    packers): how proofs made by Verifier.stwo_prove_batch become `.wit` files."""
    lo = stwo_layout(cfg)
    w = np.asarray(packed_one, dtype=np.uint32).ravel()
    Q, L, G = cfg.n_queries, cfg.n_fri_layers, cfg.lde_log
    NC = cfg.n_columns or 4  # NUM_COLUMNS, config.simf:14
    QV = NC + 16

    def dig(off: int) -> str:
        return "0x" + "".join(f"{int(x):08x}" for x in w[off:off + 8])

    def qm(off: int) -> str:
        return _qm31_str([int(x) for x in w[off:off + 4]])

    def diglist(off: int, count: int) -> str:
        return "list![" + ", ".join(dig(off + 8 * k) for k in range(count)) + "]"

    commitments = "(" + ", ".join(dig(lo.off_commit + 8 * i) for i in range(3)) + ")"
    items = []
    for q in range(Q):
        tv = "[" + ", ".join(f"[{int(w[lo.off_qvals + QV * q + i])}]" for i in range(NC)) + "]"
        cv = "[" + ", ".join(str(int(w[lo.off_qvals + QV * q + NC + i])) for i in range(16)) + "]"
        items.append(f"(({tv}, {diglist(lo.off_trace_sib + q * G * 8, G)}), ({cv}, {diglist(lo.off_cp_sib + q * G * 8, G)}))")
    oods = "([" + ", ".join("[" + qm(lo.off_oods_trace + 4 * i) + "]" for i in range(NC)) + "], [" + ", ".join(qm(lo.off_oods_cp + 4 * i) for i in range(16)) + "])"

    def layer_str(l: int) -> str:
        ns = G - 1 - l
        return "[" + ", ".join(f"({qm(lo.off_fri_wit + (l * Q + q) * 4)}, {diglist(lo.off_fri_sib[l] + q * ns * 8, ns)})" for q in range(Q)) + "]"

    fri_commitments = f"({dig(lo.off_fri_first_root)}, [" + ", ".join(dig(lo.off_fri_inner_root + 8 * i) for i in range(L)) + f"], {qm(lo.off_last_coeff)})"
    fri_decommitments = f"({layer_str(0)}, [" + ", ".join(layer_str(l) for l in range(1, L + 1)) + "])"
    nonce = (int(w[lo.off_pow_nonce]) << 32) | int(w[lo.off_pow_nonce + 1])
    vals = {
        "COMMITMENTS": commitments, "DECOMMITMENTS": "[" + ", ".join(items) + "]", "OODS_EVALS": oods,
        "FRI_COMMITMENTS": fri_commitments, "FRI_DECOMMITMENTS": fri_decommitments, "POW_NONCE": str(nonce),
    }
    return {k: {"value": v, "type": ""} for k, v in vals.items()}


def pack_stwo_proof_json(data: Dict[str, Any], cfg: StwoConfig) -> np.ndarray:
    packed, bad = pack_stwo_wits([json.dumps(stwo_wit_from_proof_json(data))], cfg)
    if bad[0]:
        raise SsymError("proof JSON does not match the configured preset")
    return packed


def pack_stark101_wits(wit_texts: Sequence[str]) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """Returns (blob, offsets [n+1] uint64, bad [n] bool)."""
    lib = load()
    cap = 20 + 8 * 3 * 31 + 31 * (16 + 8 * 62)
    recs, bad = [], np.zeros(len(wit_texts), dtype=bool)
    for i, text in enumerate(wit_texts):
        raw = text.encode() if isinstance(text, str) else text
        buf = np.zeros(cap, dtype=np.uint32)
        words = C.c_size_t(cap)
        rc = lib.ssym_s101_pack_wit(raw, len(raw), C.c_void_p(buf.ctypes.data), C.byref(words))
        if rc != 0:
            bad[i] = True
            rec = np.zeros(20, dtype=np.uint32)
            rec[0] = 20
            recs.append(rec)
        else:
            recs.append(buf[:words.value].copy())
    offsets = np.zeros(len(recs) + 1, dtype=np.uint64)
    offsets[1:] = np.cumsum([len(r) for r in recs])
    blob = np.concatenate(recs) if recs else np.zeros(0, dtype=np.uint32)
    return blob, offsets, bad


def pack_stark101_proof_json(proof: Dict[str, Any]) -> np.ndarray:
    blob, _, bad = pack_stark101_wits([json.dumps(stark101_wit_from_proof_json(proof))])
    if bad[0]:
        raise SsymError("malformed stark101 proof JSON")
    return blob


def pack_stark101_multiquery(queries: Sequence[Dict[str, Any]]) -> Tuple[np.ndarray, np.ndarray]:
    """One multi-query stark101 proof (include/ssym.h, ssym_stark101_verify_multi_batch): queries[k] in the reference's proof-JSON shape
    (scripts/fibsquare/prover.py:94-171; tests/golden/make_s101_multiquery.py) -> (blob, offsets) of len(queries) records, record k with ordinal k."""
    recs = []
    for k, q in enumerate(queries):
        rec = pack_stark101_proof_json(q).copy()
        rec[6] = k
        recs.append(rec)
    offsets = np.zeros(len(recs) + 1, dtype=np.uint64)
    offsets[1:] = np.cumsum([len(r) for r in recs])
    return np.concatenate(recs), offsets


def compact_stwo(packed: np.ndarray, cfg: StwoConfig, out: Optional[np.ndarray] = None, ver=None) -> Tuple[np.ndarray, np.ndarray]:
    """Packed records -> the compact transport form (ssym_stwo_compact_pack, include/ssym.h): per tree every distinct sibling once plus
    one bit per path slot and one back reference per repeated slot; lossless for any record.  Returns (blob of u32 words, u64 word offsets [n + 1]); `out` may be a
    preallocated (e.g. pinned) uint32 array of at least ssym_stwo_compact_bound words.
    ver: a Verifier -> version 3 records (ssym_stwo_compact_pack_gpu): the GPU finds the siblings that are nodes of other queries' paths and
    they are left out too (bound to cfg's semantics; see include/ssym.h)."""
    lib = load()
    lo = stwo_layout(cfg)
    flat = np.ascontiguousarray(np.asarray(packed, dtype=np.uint32).ravel())
    if flat.size % lo.stride_words:
        raise SsymError("packed length is not a multiple of the proof stride")
    n = flat.size // lo.stride_words
    bound = int(lib.ssym_stwo_compact_bound(C.byref(cfg), n))
    buf = out if out is not None else np.zeros(bound, dtype=np.uint32)
    offsets = np.zeros(n + 1, dtype=np.uint64)
    if ver is not None:
        check(lib.ssym_stwo_compact_pack_gpu(ver.h, C.byref(cfg), C.c_void_p(flat.ctypes.data), n, C.c_void_p(buf.ctypes.data), buf.size, C.c_void_p(offsets.ctypes.data)))
    else:
        check(lib.ssym_stwo_compact_pack(C.byref(cfg), C.c_void_p(flat.ctypes.data), n, C.c_void_p(buf.ctypes.data), buf.size, C.c_void_p(offsets.ctypes.data)))
    return (buf if out is not None else buf[:int(offsets[n])].copy()), offsets


# ---- corrupted-proof generators (fault injection; the reference has no negative tests) -----------------------
def stwo_negative_classes(cfg: StwoConfig) -> Dict[str, Tuple[int, int]]:
    """name -> (word offset inside a packed proof, value added mod 2^32).  One representative per class of
    BASELINE.json config 3: flipped Merkle sibling, bad FRI value, wrong OODS eval, bad PoW nonce, wrong last-layer
    coefficient, non-canonical field encoding."""
    lo = stwo_layout(cfg)
    G, Q = cfg.lde_log, cfg.n_queries
    q = Q // 2
    last = cfg.n_fri_layers
    return {
        "trace_sibling_bit": (lo.off_trace_sib + (q * G + G // 2) * 8 + 3, 1 << 11),
        "cp_sibling_bit": (lo.off_cp_sib + (q * G) * 8 + 7, 1),
        "fri_sibling_bit": (lo.off_fri_sib[last // 2] + 5, 1 << 30),
        "fri_witness_plus_1": (lo.off_fri_wit + (1 * Q + q) * 4 + 2, 1),
        "oods_cp_plus_1": (lo.off_oods_cp + 4 * 5, 1),
        "oods_trace_plus_1": (lo.off_oods_trace + 4 * 2 + 1, 1),
        "pow_nonce_plus_1": (lo.off_pow_nonce + 1, 1),
        "last_coeff_plus_1": (lo.off_last_coeff + 1, 1),
        "queried_value_plus_p": (lo.off_qvals + ((cfg.n_columns or 4) + 16) * q + 2, 2147483647),
    }


def apply_mutation(packed_one: np.ndarray, word: int, delta: int) -> np.ndarray:
    out = packed_one.copy()
    out[word] = np.uint32((int(out[word]) + delta) & 0xFFFFFFFF)
    return out
