#!/usr/bin/env python3
"""Extract the constants of the reference's in-source known-answer tests into tests/golden/kats.json.

The reference's whole test suite for the hot path is the set of `fn test_*` functions inside the
`.simf` sources (63 in stwo-verifier/src under `#ifdef TESTING`, 23 in stark101/src; runner
scripts/unit_tests.sh:27-108).  This script does not copy them: it pulls out, per test function,
  * every `let <name>: <type> = <literal>;` binding whose right-hand side is a pure literal, and
  * every `assert!(<pred>(<name>, <literal>))` expectation,
so that tests/test_oracle_kats_*.py can re-run the same checks against the oracle (and the GPU
tests can reuse the same vectors).  Runs only where /root/reference exists; the output is committed.
"""
import json
import os
import re
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("SSYM_REFERENCE", "/root/reference")
sys.path.insert(0, ROOT)

from oracle import witparse as W  # noqa: E402


def jsonable(v):
    if isinstance(v, (tuple, list)):
        return [jsonable(x) for x in v]
    return str(v) if v >= 1 << 53 else v  # u256 / u64 as decimal strings


def split_statements(body: str):
    out, depth, cur = [], 0, []
    for ch in body:
        if ch in "([{":
            depth += 1
        elif ch in ")]}":
            depth -= 1
        if ch == ";" and depth == 0:
            out.append("".join(cur).strip())
            cur = []
        else:
            cur.append(ch)
    tail = "".join(cur).strip()
    if tail:
        out.append(tail)
    return out


def find_eq(stmt: str, start: int) -> int:
    depth = 0
    for i in range(start, len(stmt)):
        ch = stmt[i]
        if ch in "([{<":
            depth += 1
        elif ch in ")]}>":
            depth -= 1
        elif ch == "=" and depth == 0:
            return i
    return -1


def try_literal(text: str):
    text = re.sub(r"//[^\n]*", "", text).strip()
    try:
        return W.parse_value(text)
    except Exception:
        return None


def extract_file(path: str):
    src = open(path).read()
    tests = {}
    for m in re.finditer(r"fn (test_\w+)\(\)\s*\{", src):
        i, depth = m.end(), 1
        while depth:
            depth += {"{": 1, "}": -1}.get(src[i], 0)
            i += 1
        body = re.sub(r"//[^\n]*", "", src[m.end():i - 1])
        lets, asserts = {}, []
        for stmt in split_statements(body):
            stmt = stmt.strip()
            if stmt.startswith("let "):
                colon = stmt.index(":")
                name = stmt[4:colon].strip()
                eq = find_eq(stmt, colon)
                if eq < 0:
                    continue
                val = try_literal(stmt[eq + 1:])
                if val is not None:
                    lets[name] = jsonable(val)
            for am in re.finditer(r"assert!\(\s*([\w:]+)\(\s*(\w+)\s*,\s*(.*)\)\s*\)\s*$", stmt, re.S):
                val = try_literal(am.group(3))
                asserts.append({"pred": am.group(1), "lhs": am.group(2), "rhs": jsonable(val) if val is not None else am.group(3).strip()})
        tests[m.group(1)] = {"let": lets, "assert": asserts, "line": src[:m.start()].count("\n") + 1}
    return tests


def main():
    if not os.path.isdir(REF):
        sys.exit(f"{REF} not present")
    out = {}
    for prog in ("stwo-verifier", "stark101"):
        base = os.path.join(REF, prog, "src")
        for dirpath, _, files in sorted(os.walk(base)):
            for f in sorted(files):
                if not f.endswith(".simf") or f == "padding.simf":
                    continue
                path = os.path.join(dirpath, f)
                tests = extract_file(path)
                if tests:
                    out[os.path.relpath(path, REF)] = tests
    json.dump(out, open(os.path.join(HERE, "kats.json"), "w"), indent=1)
    n = sum(len(v) for v in out.values())
    print(f"{n} test functions extracted from {len(out)} files")


if __name__ == "__main__":
    main()
