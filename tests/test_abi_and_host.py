"""CPU-side checks (`-m "not gpu"`): the C-ABI library loads and exports every symbol include/ssym.h declares,
the host-side witness ingestion agrees with the independent Python reader, and nothing computes without a GPU."""
import ctypes as C
import json
import os
import re
import subprocess

import numpy as np
import pytest

from conftest import GOLDEN, ROOT
from oracle import oracle as O
from oracle import witparse as W


@pytest.fixture(scope="module")
def S():
    import stark_symphony_b200 as S

    S.load()
    return S


def header_symbols():
    text = open(os.path.join(ROOT, "include", "ssym.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(ssym_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol(S):
    declared = header_symbols()
    assert len(declared) >= 38
    lib = C.CDLL(S._lib.LIB_PATH)
    missing = [s for s in declared if not hasattr(lib, s)]
    assert missing == []
    assert sorted(S._lib.SYMBOLS) == declared  # the ctypes binding covers exactly the header


def test_no_cpu_fallback(S):
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(S.SsymError):
        S.Verifier(0)


def test_product_does_not_use_oracle():
    """The shipped path must never route through oracle/ (no import, no include, no dlopen of liboracle)."""
    pkg = os.path.join(ROOT, "stark-symphony_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+\.*oracle", text, re.M), f
                assert not re.search(r"#include\s*[\"<][^\">]*oracle", text), f
                assert "liboracle" not in text and "oracle_" not in text, f


def test_layout_matches_oracle_and_survey(S, orc):
    for preset in ("prod", "testing"):
        for mode in (0, 1):
            cfg = S.stwo_config(preset, mode)
            lo = S.stwo_layout(cfg)
            olo = orc.stwo_layout(O.make_config(preset, mode))
            assert bytes(lo) == bytes(olo)
    assert S.stwo_layout(S.stwo_config("prod", 0)).algorithmic_bytes == 54488
    bad = S.stwo_config("prod", 0)
    bad.n_queries = 17
    with pytest.raises(S.SsymError):
        S.stwo_layout(bad)
    assert C.sizeof(S.StwoTrace) == orc.lib.oracle_sizeof_stwo_trace()
    assert C.sizeof(S.S101Trace) == orc.lib.oracle_sizeof_s101_trace()


def test_stwo_wit_packer_matches_python_reader(S):
    for preset in ("prod", "testing"):
        p = O.PRESETS[preset]
        text = open(os.path.join(GOLDEN, f"stwo_proof_{preset}.wit")).read()
        packed, bad = S.witness.pack_stwo_wits([text], S.stwo_config(preset, 0))
        ref, rej = W.pack_stwo(W.load_wit(text), p["n_queries"], p["n_fri_layers"], p["lde_log"])
        assert not bad[0] and not rej and (packed == ref).all()


def test_stwo_wit_ill_shaped_and_ill_typed(S):
    cfg = S.stwo_config("testing", 0)
    text = open(os.path.join(GOLDEN, "stwo_proof_testing.wit")).read()
    wit = json.loads(text)
    # a Merkle proof with one sibling too many is well-typed (List<u256,32>) but can never satisfy merkle.simf:42
    v = wit["DECOMMITMENTS"]["value"]
    wit2 = dict(wit, DECOMMITMENTS={"value": v.replace("list![", "list![0x01, ", 1), "type": ""})
    packed, bad = S.witness.pack_stwo_wits([json.dumps(wit2)], cfg)
    assert bad[0] and not packed.any()
    _, rej = W.pack_stwo(W.load_wit(json.dumps(wit2)), 1, 2, 4)
    assert rej
    # ill-typed: value does not fit u32 / missing witness / not JSON / wrong array length
    for broken in (
        dict(wit, POW_NONCE={"value": str(1 << 64), "type": "u64"}),
        {k: v for k, v in wit.items() if k != "OODS_EVALS"},
        dict(wit, COMMITMENTS={"value": "(1, 2)", "type": ""}),
    ):
        _, bad = S.witness.pack_stwo_wits([json.dumps(broken)], cfg)
        assert bad[0]
        with pytest.raises(W.WitnessTypeError):
            W.pack_stwo(W.load_wit(json.dumps(broken)), 1, 2, 4)
    _, bad = S.witness.pack_stwo_wits(["not json"], cfg)
    assert bad[0]


def test_stark101_wit_packer_matches_python_reader(S):
    text = open(os.path.join(GOLDEN, "stark101_proof.wit")).read()
    blob, offs, bad = S.witness.pack_stark101_wits([text, text])
    ref = W.pack_stark101(W.load_wit(text))
    assert not bad.any() and list(offs) == [0, len(ref), 2 * len(ref)] and (blob[: len(ref)] == ref).all() and (blob[len(ref):] == ref).all()
    proof = json.load(open(os.path.join(GOLDEN, "stark101_proof.json")))
    assert (S.witness.pack_stark101_proof_json(proof) == ref).all()
    # the reference's own generator output parses identically to ours
    assert json.loads(text)["P_EVALS"]["value"] == S.witness.stark101_wit_from_proof_json(proof)["P_EVALS"]["value"]
    _, _, bad = S.witness.pack_stark101_wits(['{"P_MT_ROOT": {"value": "1"}}'])
    assert bad[0]


@pytest.mark.parametrize("preset,src", [("testing", "stwo_proof_test.json"), ("prod", "stwo_proof.json")])
def test_stwo_proof_json_path_equals_reference_generator(S, preset, src):
    """The upstream proof-JSON boundary: witness.stwo_wit_from_proof_json mirrors stwo-verifier/scripts/generate_wit.py:106-245, so
    on the reference's two fixtures every witness value must equal, character for character, what the reference's generator wrote
    (tests/golden/*.wit are its output), and pack_stwo_proof_json must equal the packed golden witness."""
    data = json.load(open(os.path.join(GOLDEN, src)))
    golden_text = open(os.path.join(GOLDEN, f"stwo_proof_{preset}.wit")).read()
    golden = json.loads(golden_text)
    ours = S.witness.stwo_wit_from_proof_json(data)
    assert set(ours) == set(golden) == {"COMMITMENTS", "DECOMMITMENTS", "OODS_EVALS", "FRI_COMMITMENTS", "FRI_DECOMMITMENTS", "POW_NONCE"}
    for name in golden:
        assert ours[name]["value"] == golden[name]["value"], name
    cfg = S.stwo_config(preset, 0)
    want, bad = S.witness.pack_stwo_wits([golden_text], cfg)
    assert not bad[0]
    assert (S.witness.pack_stwo_proof_json(data, cfg) == want).all()
    p = O.PRESETS[preset]
    ref, rej = W.pack_stwo(W.load_wit(json.dumps(ours)), p["n_queries"], p["n_fri_layers"], p["lde_log"])
    assert not rej and (ref == want).all()
    with pytest.raises(S.SsymError):  # the other preset's shape
        S.witness.pack_stwo_proof_json(data, S.stwo_config("prod" if preset == "testing" else "testing", 0))


def test_value_grammar_corner_cases():
    assert W.parse_value("(1, (2, 3), [4, 5], list![], list![6,], 0x10, 1_000)") == (1, (2, 3), [4, 5], [], [6], 16, 1000)
    assert W.parse_value("((7))") == 7
    for bad in ("(1, 2", "list!(1)", "foo", "1 2", ""):
        with pytest.raises(W.WitnessTypeError):
            W.parse_value(bad)


def test_cli_usage_and_parse_errors(S):
    cli = os.path.join(ROOT, "stark-symphony_b200", "bin", "verify-batch")
    assert os.path.exists(cli)
    r = subprocess.run([cli, "--help"], capture_output=True, text=True)
    assert r.returncode == 0 and "--program" in r.stderr
    r = subprocess.run([cli, "--program", "stwo"], capture_output=True, text=True)
    assert r.returncode == 2
    r = subprocess.run([cli, "--program", "stwo", "--witness", "/nonexistent.wit"], capture_output=True, text=True)
    assert r.returncode == 1 and "Failed to read witness file" in r.stderr  # same wording as simfony-cli/src/main.rs:180
