"""Differential fuzz of the host witness parser (csrc/witness.cpp behind ssym_stwo_pack_wit) against the independent Python reader
(oracle/witparse.py) on mutated `.wit` texts of the TESTING preset: token-level rewrites that keep the value (spacing, hex / decimal /
underscores, redundant parentheses, trailing commas) and ones that break it (dropped / duplicated / swapped tokens, out-of-range literals).
Both implementations must agree on accept / reject and, when accepted, on every packed word — the grammar is the one `simfony run --witness`
reads (simfony-cli/src/main.rs:77-81; values as emitted by stwo-verifier/scripts/generate_wit.py:139-243)."""
import json
import os
import re

import numpy as np
import pytest
from hypothesis import HealthCheck, given, settings
from hypothesis import strategies as st

from conftest import GOLDEN
from oracle import witparse as W

TOKEN = re.compile(r"0x[0-9a-fA-F]+|[0-9]+|list!|[()\[\],]")


@pytest.fixture(scope="module")
def S():
    import stark_symphony_b200 as S

    S.load()
    return S


@pytest.fixture(scope="module")
def base():
    return json.loads(open(os.path.join(GOLDEN, "stwo_proof_testing.wit")).read())


def _python_result(text, cfg):
    try:
        rec, rej = W.pack_stwo(W.load_wit(text), cfg.n_queries, cfg.n_fri_layers, cfg.lde_log, cfg.n_columns or 4)
        return ("shape", None) if rej else ("ok", rec)
    except (W.WitnessTypeError, ValueError, KeyError, IndexError, TypeError, RecursionError):
        return ("reject", None)


def _render(tokens, seps):
    return "".join(s + t for s, t in zip(seps, tokens)) + seps[-1]


@st.composite
def mutated_value(draw, value):
    toks = TOKEN.findall(value)
    toks = list(toks)
    n_mut = draw(st.integers(0, 3))
    for _ in range(n_mut):
        kind = draw(st.sampled_from(["respell", "paren", "trail", "drop", "dup", "swap", "big", "junk"]))
        i = draw(st.integers(0, len(toks) - 1))
        t = toks[i]
        if t[0].isdigit() and not re.fullmatch(r"0x[0-9a-fA-F]+|[0-9]+", t):
            continue  # a junk token from an earlier mutation
        if kind == "respell" and t[0].isdigit():
            v = int(t, 0)
            form = draw(st.sampled_from(["dec", "hex", "hex0", "us", "HEX"]))
            toks[i] = {"dec": str(v), "hex": hex(v), "hex0": "0x" + "0" * draw(st.integers(1, 3)) + format(v, "x"),
                       "us": (str(v)[0] + "_" + str(v)[1:]) if len(str(v)) > 1 else str(v), "HEX": "0x" + format(v, "X")}[form]
        elif kind == "paren" and t[0].isdigit():
            toks[i:i + 1] = ["(", t, ")"]
        elif kind == "trail" and t in (")", "]") and i > 0 and toks[i - 1] not in ("(", "[", ","):
            toks.insert(i, ",")
        elif kind == "drop":
            del toks[i]
            if not toks:
                toks = ["0"]
        elif kind == "dup":
            toks.insert(i, t)
        elif kind == "swap" and i + 1 < len(toks):
            toks[i], toks[i + 1] = toks[i + 1], toks[i]
        elif kind == "big" and t[0].isdigit():
            toks[i] = str(int(t, 0) + draw(st.sampled_from([2**32, 2**64, 2**256])))
        elif kind == "junk":
            toks.insert(i, draw(st.sampled_from([";", "list", "!", "x", "0x", "-1", "1.5", "{"])))
    seps = [draw(st.sampled_from(["", " ", "  ", "\\n", "\\t "])) for _ in range(len(toks) + 1)]
    return _render(toks, seps)


@st.composite
def mutated_wit(draw, base):
    names = list(base)
    victim = draw(st.sampled_from(names))
    wit = {k: {"value": v["value"], "type": v.get("type", "")} for k, v in base.items()}
    wit[victim]["value"] = draw(mutated_value(base[victim]["value"]))
    if draw(st.booleans()):
        wit = {k: wit[k] for k in draw(st.permutations(names))}
    if draw(st.integers(0, 9)) == 0:
        del wit[draw(st.sampled_from(names))]
    text = json.dumps(wit, indent=draw(st.sampled_from([None, 0, 2])))
    return text.replace("\\\\n", "\\n").replace("\\\\t", "\\t")  # the separators above are JSON escapes inside the value string


@settings(max_examples=300, deadline=None, suppress_health_check=[HealthCheck.function_scoped_fixture, HealthCheck.too_slow, HealthCheck.data_too_large])
@given(data=st.data())
def test_host_parser_agrees_with_python_reader(S, base, data):
    cfg = S.stwo_config("testing", 0)
    text = data.draw(mutated_wit(base))
    packed, bad = S.witness.pack_stwo_wits([text], cfg)
    kind, rec = _python_result(text, cfg)
    if kind == "ok":
        assert not bad[0], text[:300]
        assert (packed == rec).all()
    else:
        assert bad[0], (kind, text[:300])
        assert not packed.any()


@pytest.fixture(scope="module")
def base101():
    return json.loads(open(os.path.join(GOLDEN, "stark101_proof.wit")).read())


@settings(max_examples=200, deadline=None, suppress_health_check=[HealthCheck.function_scoped_fixture, HealthCheck.too_slow, HealthCheck.data_too_large])
@given(data=st.data())
def test_stark101_host_parser_agrees_with_python_reader(S, base101, data):
    """The same for the stark101 witnesses (stark101/src/main.simf:12-20, lists of 0..31 items: dropping or duplicating a sibling or a whole FRI
    layer stays well-typed and must pack identically in both implementations)."""
    text = data.draw(mutated_wit(base101))
    blob, offsets, bad = S.witness.pack_stark101_wits([text])
    try:
        rec = W.pack_stark101(W.load_wit(text))
    except (W.WitnessTypeError, ValueError, KeyError, IndexError, TypeError, RecursionError):
        rec = None
    if rec is None:
        assert bad[0], text[:300]
    else:
        assert not bad[0], text[:300]
        assert len(blob) == len(rec) and (blob == rec).all()


# ---- GPU tokeniser (csrc/wit_kernels.cu) against the host parser on the same kind of mutations --------------------------------
def _random_mutation(rng, value):
    toks = TOKEN.findall(value)
    for _ in range(int(rng.integers(0, 4))):
        kind = rng.choice(["respell", "paren", "trail", "drop", "dup", "swap", "big", "junk"])
        i = int(rng.integers(0, len(toks)))
        t = toks[i]
        if t[0].isdigit() and not re.fullmatch(r"0x[0-9a-fA-F]+|[0-9]+", t):
            continue
        if kind == "respell" and t[0].isdigit():
            v = int(t, 0)
            toks[i] = [str(v), hex(v), "0x00" + format(v, "x"), (str(v)[0] + "_" + str(v)[1:]) if len(str(v)) > 1 else str(v), "0x" + format(v, "X")][int(rng.integers(0, 5))]
        elif kind == "paren" and t[0].isdigit():
            toks[i:i + 1] = ["(", t, ")"]
        elif kind == "trail" and t in (")", "]") and i > 0 and toks[i - 1] not in ("(", "[", ","):
            toks.insert(i, ",")
        elif kind == "drop" and len(toks) > 1:
            del toks[i]
        elif kind == "dup":
            toks.insert(i, t)
        elif kind == "swap" and i + 1 < len(toks):
            toks[i], toks[i + 1] = toks[i + 1], toks[i]
        elif kind == "big" and t[0].isdigit():
            toks[i] = str(int(t, 0) + [2**32, 2**64, 2**256][int(rng.integers(0, 3))])
        elif kind == "junk":
            toks.insert(i, [";", "list", "!", "x", "0x", "-1", "1.5", "{"][int(rng.integers(0, 8))])
    seps = [["", " ", "  ", "\\n", "\\t "][int(rng.integers(0, 5))] for _ in range(len(toks) + 1)]
    return _render(toks, seps)


def _random_wit(rng, base):
    names = list(base)
    wit = {k: {"value": v["value"], "type": v.get("type", "")} for k, v in base.items()}
    victim = names[int(rng.integers(0, len(names)))]
    wit[victim]["value"] = _random_mutation(rng, base[victim]["value"])
    if rng.integers(0, 2):
        wit = {k: wit[k] for k in rng.permutation(names)}
    if rng.integers(0, 10) == 0:
        del wit[names[int(rng.integers(0, len(names)))]]
    text = json.dumps(wit, indent=[None, 0, 2][int(rng.integers(0, 3))])
    return text.replace("\\\\n", "\\n").replace("\\\\t", "\\t")


@pytest.mark.gpu
@pytest.mark.parametrize("preset,n", [("testing", 1500), ("prod", 96)])
def test_gpu_tokeniser_agrees_with_host_parser_on_mutations(S, preset, n):
    """One batch of randomly mutated witnesses through ssym_stwo_pack_wit_batch: record and flag (ok / shape / parse) of every witness equal
    what ssym_stwo_pack_wit gives — whichever of them the GPU kept on its fast path and whichever it handed to the host parser."""
    rng = np.random.default_rng(2025)
    cfg = S.stwo_config(preset, 0)
    lo = S.stwo_layout(cfg)
    base = json.loads(open(os.path.join(GOLDEN, f"stwo_proof_{preset}.wit")).read())
    texts = [_random_wit(rng, base) for _ in range(n)]
    ver = S.Verifier(0)
    blob, offsets = S.witness.concat_wit_texts(texts)
    g_packed, g_flags = ver.stwo_pack_wit_batch(blob, offsets, cfg)
    ver.close()
    lib = S.load()
    import ctypes as C

    n_ok = 0
    for i, text in enumerate(texts):
        raw = text.encode()
        out = np.zeros(lo.stride_words, dtype=np.uint32)
        shape = C.c_int(0)
        rc = lib.ssym_stwo_pack_wit(C.byref(cfg), raw, len(raw), C.c_void_p(out.ctypes.data), C.byref(shape))
        want = 2 if rc != 0 else (1 if shape.value else 0)
        assert int(g_flags[i]) == want, (i, int(g_flags[i]), want, text[:200])
        assert (g_packed[i] == out).all(), i
        n_ok += want == 0
    assert 0 < n_ok < n  # the batch mixes accepted and refused witnesses


# ---- hostile and JSON-level cases (untrusted text must be a reject, never a crash; accept <=> serde would parse it) ------------
def _json_level_cases(base):
    good = {k: {"value": v["value"], "type": v.get("type", "")} for k, v in base.items()}
    text = json.dumps(good)
    first = next(iter(good))
    no_type = json.dumps({k: ({"value": v["value"]} if k == first else v) for k, v in good.items()})
    type_not_string = json.dumps({k: ({"value": v["value"], "type": 5} if k == first else v) for k, v in good.items()})
    dup_value = text.replace('"value":', '"value": "0", "value":', 1)
    dup_type = text.replace('"type":', '"type": "u32", "type":', 1)
    dup_name = text[:-1] + ", " + json.dumps({first: good[first]})[1:]
    return {
        "good": (text, True), "trailing garbage": (text + " x", False), "trailing object": (text + "{}", False), "trailing whitespace": (text + " \n\t", True),
        "no type": (no_type, False), "type not a string": (type_not_string, False), "two values": (dup_value, False), "two types": (dup_type, False),
        "two witnesses of one name": (dup_name, False),
        "bad unicode escape": (text.replace('"value": "', '"value": "\\u00zz', 1), False),
        "ascii unicode escape": (text.replace('"value": "', '"value": "\\u0020', 1), True),
        "unknown member": (text.replace('"type":', '"note": [1, {"a": [2]}], "type":', 1), True),
    }


def test_json_level_strictness_matches_python_reader(S, base):
    cfg = S.stwo_config("testing", 0)
    for name, (text, ok) in _json_level_cases(base).items():
        packed, bad = S.witness.pack_stwo_wits([text], cfg)
        kind, rec = _python_result(text, cfg)
        assert (not bad[0]) == ok, name
        assert (kind == "ok") == ok, name
        if ok:
            assert (packed == rec).all(), name


def test_one_tuple_is_not_a_parenthesised_value(S, base):
    """`(x)` is a parenthesised expression (accepted, = x); `(x,)` is a 1-tuple and no witness type of either program is one."""
    cfg = S.stwo_config("testing", 0)
    wit = {k: dict(v) for k, v in base.items()}
    nonce = wit["POW_NONCE"]["value"].strip()
    for form, ok in ((f"({nonce})", True), (f"(({nonce}))", True), (f"({nonce},)", False), (f"(({nonce}),)", False)):
        wit["POW_NONCE"]["value"] = form
        text = json.dumps(wit)
        _, bad = S.witness.pack_stwo_wits([text], cfg)
        assert (not bad[0]) == ok, form
        assert (_python_result(text, cfg)[0] == "ok") == ok, form


@pytest.mark.parametrize("depth", [60, 70, 1000, 2_000_000])
def test_hostile_nesting_is_a_parse_error_not_a_crash(S, base, base101, depth):
    """A value of `depth` opening brackets, and the same nesting in a JSON member the parser skips: ParseError beyond 64 levels, at any depth,
    in both programs' entry points (2 000 000 '(' used to overflow the stack of ssym_s101_pack_wit)."""
    import ctypes as C

    lib = S.load()
    cfg = S.stwo_config("testing", 0)
    lo = S.stwo_layout(cfg)
    for opener, closer in (("(", ")"), ("[", "]"), ("list![", "]")):
        wit = {k: dict(v) for k, v in base.items()}
        wit["POW_NONCE"]["value"] = opener * depth + "1" + closer * depth
        raw = json.dumps(wit).encode()
        out = np.zeros(lo.stride_words, dtype=np.uint32)
        shape = C.c_int(0)
        rc = lib.ssym_stwo_pack_wit(C.byref(cfg), raw, len(raw), C.c_void_p(out.ctypes.data), C.byref(shape))
        if opener == "(" and depth < 64:
            assert rc == 0  # redundant parentheses within the bound are a parenthesised value
        else:
            assert rc == S.ERR_PARSE and not out.any()
    # nesting inside a member that is skipped (no closing brackets at all for the deepest case: still no crash)
    wit = json.dumps({k: dict(v) for k, v in base101.items()})
    deep = wit.replace('"type":', '"skipped": ' + "[" * depth + "]" * (depth if depth < 10**6 else 0) + ', "type":', 1).encode()
    rec = np.zeros(4096, dtype=np.uint32)
    words = C.c_size_t(rec.size)
    rc = lib.ssym_s101_pack_wit(deep, len(deep), C.c_void_p(rec.ctypes.data), C.byref(words))
    assert rc == (0 if depth < 64 else S.ERR_PARSE)


def test_huge_list_is_rejected_without_escaping_exceptions(S, base101):
    """A list! of two million items is well-formed text but not a List<_, 32>: rejected (parse error), nothing escapes the extern "C" boundary."""
    import ctypes as C

    lib = S.load()
    wit = {k: dict(v) for k, v in base101.items()}
    wit["FRI_LAYERS"]["value"] = "list![" + "1," * 2_000_000 + "1]"
    raw = json.dumps(wit).encode()
    rec = np.zeros(4096, dtype=np.uint32)
    words = C.c_size_t(rec.size)
    assert lib.ssym_s101_pack_wit(raw, len(raw), C.c_void_p(rec.ctypes.data), C.byref(words)) in (S.ERR_PARSE, S.ERR_NOMEM)
