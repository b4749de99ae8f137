"""Independent Python reader for the reference's witness files — TEST INFRASTRUCTURE ONLY.

Parses the `.wit` JSON that `simfony run --witness` consumes
(simfony-cli/src/main.rs:77-81: a JSON object NAME -> {"value": <SimplicityHL value text>,
"type": <type text>}) using the value grammar actually emitted by
stwo-verifier/scripts/generate_wit.py:32-35,139-243 and stark101/scripts/generate_wit.py:7-30
(decimal / 0x integers, tuples `( )`, arrays `[ ]`, `list![ ]`), and packs the result into the
binary wire format documented in include/ssym.h.  It exists so the tests can check the product's
C++ parser/packer (stark-symphony_b200/csrc/witness.cpp) against a second implementation.
It must never be imported from the product package.
"""
from __future__ import annotations

import json
import re
from typing import Any, Dict, List, Tuple

import numpy as np


class WitnessTypeError(ValueError):
    """The text is not a value of the program's witness types (simfony would refuse it)."""


class ListValue(list):
    """A `list![...]` value (List<T, 32>: 0..31 items), as opposed to a fixed array `[...]`."""


_TOKEN = re.compile(r"\s*(0x[0-9a-fA-F_]+|[0-9][0-9_]*|list!|[A-Za-z_][A-Za-z_0-9]*|[()\[\],])")


def tokenize(text: str) -> List[str]:
    out, pos = [], 0
    text = text.rstrip()
    while pos < len(text):
        m = _TOKEN.match(text, pos)
        if not m:
            raise WitnessTypeError(f"bad token at {pos}: {text[pos:pos+20]!r}")
        out.append(m.group(1))
        pos = m.end()
    return out


def parse_value(text: str) -> Any:
    toks = tokenize(text)
    val, pos = _parse(toks, 0)
    if pos != len(toks):
        raise WitnessTypeError("trailing tokens")
    return val


class _Trailing(list):
    """Items of a sequence that ended in a trailing comma (`(x,)` is a 1-tuple, `(x)` a parenthesised value)."""


def _parse_seq(toks: List[str], pos: int, close: str, depth: int = 0) -> Tuple[List[Any], int]:
    items: List[Any] = []
    if pos < len(toks) and toks[pos] == close:
        return items, pos + 1
    while True:
        v, pos = _parse(toks, pos, depth)
        items.append(v)
        if pos >= len(toks):
            raise WitnessTypeError("unterminated sequence")
        if toks[pos] == ",":
            pos += 1
            if pos < len(toks) and toks[pos] == close:  # trailing comma
                return _Trailing(items), pos + 1
            continue
        if toks[pos] == close:
            return items, pos + 1
        raise WitnessTypeError(f"expected , or {close}, got {toks[pos]}")


MAX_NESTING = 64  # same bound as csrc/witness.cpp


def _parse(toks: List[str], pos: int, depth: int = 0) -> Tuple[Any, int]:
    if pos >= len(toks):
        raise WitnessTypeError("unexpected end")
    depth += 1
    if depth > MAX_NESTING:
        raise WitnessTypeError("nesting too deep")
    t = toks[pos]
    if t == "(":
        items, pos = _parse_seq(toks, pos + 1, ")", depth)
        if len(items) == 1:
            if isinstance(items, _Trailing):  # `(x,)`: a 1-tuple, a value of none of the witness types
                raise WitnessTypeError("1-tuple")
            return items[0], pos  # parenthesised value
        return tuple(items), pos
    if t == "[":
        items, pos = _parse_seq(toks, pos + 1, "]", depth)
        return list(items), pos
    if t == "list!":
        if pos + 1 >= len(toks) or toks[pos + 1] != "[":
            raise WitnessTypeError("list! must be followed by [")
        items, pos = _parse_seq(toks, pos + 2, "]", depth)
        return ListValue(items), pos
    if t == "qm31":  # constructor used in .simf literals (fields/qm31.simf:20-22)
        if toks[pos + 1] != "(":
            raise WitnessTypeError("qm31 needs (")
        items, pos = _parse_seq(toks, pos + 2, ")", depth)
        if len(items) != 4:
            raise WitnessTypeError("qm31 takes 4 values")
        return ((items[0], items[1]), (items[2], items[3])), pos
    if t[0].isdigit():
        return int(t.replace("_", ""), 0), pos + 1
    raise WitnessTypeError(f"unexpected token {t}")


def _no_duplicates(pairs):
    keys = [k for k, _ in pairs]
    if len(set(keys)) != len(keys):
        raise WitnessTypeError("duplicate member")
    return dict(pairs)


def load_wit(text: str) -> Dict[str, Any]:
    """NAME -> {"value": str, "type": str}, both members required, no repeated names or members (serde's WitnessValues, simfony-cli/src/main.rs:77-81)."""
    obj = json.loads(text, object_pairs_hook=_no_duplicates)
    if not isinstance(obj, dict):
        raise WitnessTypeError("witness file must be a JSON object")
    out = {}
    for name, entry in obj.items():
        if not isinstance(entry, dict) or not isinstance(entry.get("value"), str):
            raise WitnessTypeError(f"witness {name} has no value")
        if not isinstance(entry.get("type"), str):
            raise WitnessTypeError(f"witness {name} has no type")
        out[name] = parse_value(entry["value"])
    return out


# --------------------------------------------------------------------------------------
# typed access helpers
# --------------------------------------------------------------------------------------
def _uint(v: Any, bits: int) -> int:
    if not isinstance(v, int) or isinstance(v, bool) or v < 0 or v >> bits:
        raise WitnessTypeError(f"expected u{bits}, got {v!r}")
    return v


def _tuple(v: Any, n: int) -> tuple:
    if not isinstance(v, tuple) or len(v) != n:
        raise WitnessTypeError(f"expected {n}-tuple")
    return v


def _array(v: Any, n: int) -> list:
    if not isinstance(v, list) or isinstance(v, ListValue) or len(v) != n:
        raise WitnessTypeError(f"expected array of {n}")
    return v


def _list32(v: Any) -> ListValue:
    if not isinstance(v, ListValue) or len(v) >= 32:
        raise WitnessTypeError("expected List<_, 32>")
    return v


def u256_words(v: int) -> List[int]:
    """8 big-endian 32-bit limbs, most significant first (channel.simf:48-58)."""
    _uint(v, 256)
    return [(v >> (32 * (7 - i))) & 0xFFFFFFFF for i in range(8)]


def _qm31(v: Any) -> List[int]:
    (a, b), (c, d) = (_tuple(x, 2) for x in _tuple(v, 2))
    return [_uint(a, 32), _uint(b, 32), _uint(c, 32), _uint(d, 32)]


# --------------------------------------------------------------------------------------
# Stwo layout (mirrors include/ssym.h; deliberately re-derived here, not imported)
# --------------------------------------------------------------------------------------
def _align8(w: int) -> int:
    return (w + 7) & ~7


def stwo_layout(n_queries: int, n_fri_layers: int, lde_log: int, n_columns: int = 4) -> Dict[str, Any]:
    Q, L, G, NC = n_queries, n_fri_layers, lde_log, n_columns
    lo: Dict[str, Any] = {}
    w = 0
    lo["commit"] = w; w += 24
    lo["oods_trace"] = w; w += 4 * NC
    lo["oods_cp"] = w; w += 64
    lo["fri_first_root"] = w; w += 8
    lo["fri_inner_root"] = w; w += 8 * L
    lo["last_coeff"] = w; w += 4
    lo["pow_nonce"] = w; w += 2
    alg = w
    w = _align8(w)
    lo["qvals"] = w; w += Q * (NC + 16); alg += Q * (NC + 16); w = _align8(w)
    lo["trace_sib"] = w; w += Q * G * 8; alg += Q * G * 8
    lo["cp_sib"] = w; w += Q * G * 8; alg += Q * G * 8
    lo["fri_wit"] = w; w += (L + 1) * Q * 4; alg += (L + 1) * Q * 4; w = _align8(w)
    lo["fri_sib"] = []
    for l in range(L + 1):
        lo["fri_sib"].append(w)
        w += Q * (G - 1 - l) * 8
        alg += Q * (G - 1 - l) * 8
    lo["stride_words"] = _align8(w)
    lo["algorithmic_bytes"] = alg * 4
    return lo


def pack_stwo(wit: Dict[str, Any], n_queries: int, n_fri_layers: int, lde_log: int, n_columns: int = 4) -> Tuple[np.ndarray, bool]:
    """Pack the six witnesses of stwo-verifier/src/main.simf:9-25 (n_columns = NUM_COLUMNS, config.simf:14).  Returns (words, shape_reject)."""
    Q, L, G, NC = n_queries, n_fri_layers, lde_log, n_columns
    lo = stwo_layout(Q, L, G, NC)
    out = np.zeros(lo["stride_words"], dtype=np.uint32)
    shape_reject = False

    def put(off: int, words: List[int]) -> None:
        out[off:off + len(words)] = np.array(words, dtype=np.uint64).astype(np.uint32)

    for name in ("COMMITMENTS", "DECOMMITMENTS", "OODS_EVALS", "FRI_COMMITMENTS", "FRI_DECOMMITMENTS", "POW_NONCE"):
        if name not in wit:
            raise WitnessTypeError(f"missing witness {name}")
    for i, r in enumerate(_tuple(wit["COMMITMENTS"], 3)):
        put(lo["commit"] + 8 * i, u256_words(r))
    oods_trace, oods_cp = _tuple(wit["OODS_EVALS"], 2)
    for i, col in enumerate(_array(oods_trace, NC)):
        put(lo["oods_trace"] + 4 * i, _qm31(_array(col, 1)[0]))
    for i, v in enumerate(_array(oods_cp, 16)):
        put(lo["oods_cp"] + 4 * i, _qm31(v))
    first_root, inner_roots, last_coeff = _tuple(wit["FRI_COMMITMENTS"], 3)
    put(lo["fri_first_root"], u256_words(first_root))
    for i, r in enumerate(_array(inner_roots, L)):
        put(lo["fri_inner_root"] + 8 * i, u256_words(r))
    put(lo["last_coeff"], _qm31(last_coeff))
    nonce = _uint(wit["POW_NONCE"], 64)
    put(lo["pow_nonce"], [nonce >> 32, nonce & 0xFFFFFFFF])

    for q, dec in enumerate(_array(wit["DECOMMITMENTS"], Q)):
        (tvals, tproof), (cvals, cproof) = (_tuple(x, 2) for x in _tuple(dec, 2))
        put(lo["qvals"] + (NC + 16) * q, [_uint(_array(c, 1)[0], 32) for c in _array(tvals, NC)])
        put(lo["qvals"] + (NC + 16) * q + NC, [_uint(v, 32) for v in _array(cvals, 16)])
        for name, proof in (("trace_sib", tproof), ("cp_sib", cproof)):
            proof = _list32(proof)
            if len(proof) != G:
                shape_reject = True  # merkle.simf:42: path == 1 cannot hold
                continue
            for k, s in enumerate(proof):
                put(lo[name] + (q * G + k) * 8, u256_words(s))
    first_dec, inner_decs = _tuple(wit["FRI_DECOMMITMENTS"], 2)
    layers = [first_dec] + list(_array(inner_decs, L))
    for l, layer in enumerate(layers):
        n_sib = G - 1 - l
        for q, item in enumerate(_array(layer, Q)):
            w4, proof = _tuple(item, 2)
            put(lo["fri_wit"] + (l * Q + q) * 4, _qm31(w4))
            proof = _list32(proof)
            if len(proof) != n_sib:
                shape_reject = True
                continue
            for k, s in enumerate(proof):
                put(lo["fri_sib"][l] + (q * n_sib + k) * 8, u256_words(s))
    if shape_reject:
        out[:] = 0
    return out, shape_reject


def pack_stark101(wit: Dict[str, Any]) -> np.ndarray:
    """Pack P_MT_ROOT, P_EVALS, FRI_LAYERS, FRI_LAST_LAYER (stark101/src/main.simf:12-20)."""
    for name in ("P_MT_ROOT", "P_EVALS", "FRI_LAYERS", "FRI_LAST_LAYER"):
        if name not in wit:
            raise WitnessTypeError(f"missing witness {name}")
    root = u256_words(wit["P_MT_ROOT"])
    evals = [_tuple(e, 2) for e in _tuple(wit["P_EVALS"], 3)]
    layers = _list32(wit["FRI_LAYERS"])
    last = _uint(wit["FRI_LAST_LAYER"], 32)
    words: List[int] = [0, len(layers)]
    ev_vals, ev_sibs = [], []
    for val, proof in evals:
        ev_vals.append(_uint(val, 32))
        ev_sibs.append(_list32(proof))
    words += [len(p) for p in ev_sibs]
    words += [last, 0, 0]
    words += root
    words += ev_vals + [0]
    for p in ev_sibs:
        for s in p:
            words += u256_words(s)
    for layer in layers:
        lroot, beta, cpa, pa, cpb, pb = _tuple(layer, 6)
        pa, pb = _list32(pa), _list32(pb)
        words += u256_words(lroot)
        words += [_uint(beta, 32), _uint(cpa, 32), _uint(cpb, 32), len(pa), len(pb), 0, 0, 0]
        for s in pa:
            words += u256_words(s)
        for s in pb:
            words += u256_words(s)
    words[0] = len(words)
    return np.array(words, dtype=np.uint64).astype(np.uint32)
