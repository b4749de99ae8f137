#!/usr/bin/env python3
"""Regenerate tests/golden/ from the reference tree (runs only where /root/reference exists).

Everything written here is the OUTPUT of running the reference's own Python scripts on the
reference's own fixtures (no reference source is copied):

  stwo_proof_prod.wit      = stwo-verifier/scripts/generate_wit.py tests/data/proof.json      (Makefile:10-11 `make proof-wit`)
  stwo_proof_testing.wit   = stwo-verifier/scripts/generate_wit.py tests/data/proof_test.json
  stwo_proof.json, stwo_proof_test.json = the two upstream proof-JSON fixtures themselves (stwo-verifier/tests/data/, data files)
  stark101_proof.json      = `python -m fibsquare`  (stark101/Makefile:14-15 `make proof`, deterministic: prover.py:27 seed)
  stark101_proof.wit       = stark101/scripts/generate_wit.py stark101_proof.json
  simf_literals.json       = the witness literals embedded in stwo-verifier/src/verifier.simf:62-108
                             (`test_verify_proof`) and stark101/src/verifier.simf:44-388 (`test_verifier`),
                             parsed with oracle/witparse.py and packed — they must equal the packed .wit files.

Usage: python tests/golden/make_golden.py [--skip-prover]
"""
import json
import os
import re
import shutil
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("SSYM_REFERENCE", "/root/reference")
sys.path.insert(0, ROOT)

from oracle import witparse as W  # noqa: E402


def run(cmd, **kw):
    return subprocess.run(cmd, check=True, capture_output=True, text=True, **kw).stdout


def extract_literal(path: str, fn_name: str, type_name: str) -> str:
    """Return the text of `let proof: <type_name> = <value>;` inside `fn <fn_name>()`."""
    src = open(path).read()
    start = src.index(f"fn {fn_name}()")
    m = re.search(r"let\s+proof\s*:\s*" + type_name + r"\s*=", src[start:])
    pos = start + m.end()
    depth, i = 0, pos
    while True:
        ch = src[i]
        if ch in "([":
            depth += 1
        elif ch in ")]":
            depth -= 1
        elif ch == ";" and depth == 0:
            break
        i += 1
    return src[pos:i]


def main():
    skip_prover = "--skip-prover" in sys.argv
    if not os.path.isdir(REF):
        sys.exit(f"{REF} not present: golden vectors can only be regenerated next to the reference")
    stwo = os.path.join(REF, "stwo-verifier")
    for name, src in (("prod", "proof.json"), ("testing", "proof_test.json")):
        out = run([sys.executable, os.path.join(stwo, "scripts", "generate_wit.py"), os.path.join(stwo, "tests", "data", src)])
        open(os.path.join(HERE, f"stwo_proof_{name}.wit"), "w").write(out)
        # the upstream proof JSON itself (data, not source): the input of the proof-JSON boundary test (tests/test_abi_and_host.py)
        shutil.copy(os.path.join(stwo, "tests", "data", src), os.path.join(HERE, f"stwo_{src}"))

    s101_json = os.path.join(HERE, "stark101_proof.json")
    if not skip_prover or not os.path.exists(s101_json):
        with tempfile.TemporaryDirectory() as tmp:
            shutil.copytree(os.path.join(REF, "stark101"), os.path.join(tmp, "stark101"))
            os.makedirs(os.path.join(tmp, "stark101", "target"), exist_ok=True)
            run([sys.executable, "-m", "fibsquare"], cwd=os.path.join(tmp, "stark101", "scripts"))
            shutil.copy(os.path.join(tmp, "stark101", "target", "proof.json"), s101_json)
    out = run([sys.executable, os.path.join(REF, "stark101", "scripts", "generate_wit.py"), s101_json])
    open(os.path.join(HERE, "stark101_proof.wit"), "w").write(out)

    # Cross-check against the literals embedded in the .simf tests.
    lit = {}
    text = extract_literal(os.path.join(stwo, "src", "verifier.simf"), "test_verify_proof", "StarkProof")
    commitments, decommitments, oods, fri_c, fri_d, nonce = W.parse_value(text)
    wit = {"COMMITMENTS": commitments, "DECOMMITMENTS": decommitments, "OODS_EVALS": oods,
           "FRI_COMMITMENTS": fri_c, "FRI_DECOMMITMENTS": fri_d, "POW_NONCE": nonce}
    packed_lit, rej = W.pack_stwo(wit, 1, 2, 4)
    packed_wit, _ = W.pack_stwo(W.load_wit(open(os.path.join(HERE, "stwo_proof_testing.wit")).read()), 1, 2, 4)
    assert not rej and (packed_lit == packed_wit).all(), "verifier.simf literal != proof_test.json witness"
    lit["stwo_testing_packed_hex"] = packed_lit.tobytes().hex()

    text = extract_literal(os.path.join(REF, "stark101", "src", "verifier.simf"), "test_verifier", "FibSquareProof")
    root, evals, layers, last = W.parse_value(text)
    rec_lit = W.pack_stark101({"P_MT_ROOT": root, "P_EVALS": evals, "FRI_LAYERS": layers, "FRI_LAST_LAYER": last})
    rec_wit = W.pack_stark101(W.load_wit(open(os.path.join(HERE, "stark101_proof.wit")).read()))
    assert len(rec_lit) == len(rec_wit) and (rec_lit == rec_wit).all(), "stark101 verifier.simf literal != regenerated proof"
    lit["stark101_packed_hex"] = rec_lit.tobytes().hex()
    json.dump(lit, open(os.path.join(HERE, "simf_literals.json"), "w"))
    print("golden vectors regenerated; .simf literals match the witness files")


if __name__ == "__main__":
    main()
