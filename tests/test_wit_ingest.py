"""GPU `.wit` ingestion (csrc/wit_kernels.cu, ssym_stwo_pack_wit_batch / ssym_stwo_verify_wit_batch).

CPU part: the token skeleton + literal slots the GPU checks against are validated with the independent Python reader
(oracle/witparse.py): tokenising the reference's fixtures gives exactly the skeleton, and scattering the literals through
the slots gives exactly the packed record.  GPU part: the tokeniser against the host parser and the Python reader on the
fixtures, on re-formatted texts, on texts outside its fast path (which it must hand to the host parser) and on malformed
ones; then text -> accept bits end to end against the oracle."""
import ctypes as C
import json
import os
import re

import numpy as np
import pytest

from conftest import GOLDEN
from oracle import oracle as O
from oracle import witparse as W

NAMES = ["COMMITMENTS", "DECOMMITMENTS", "OODS_EVALS", "FRI_COMMITMENTS", "FRI_DECOMMITMENTS", "POW_NONCE"]


@pytest.fixture(scope="module")
def S():
    import stark_symphony_b200 as S

    S.load()
    return S


def skeleton(S, cfg, name):
    lib = S.load()
    ns, nn = C.c_size_t(0), C.c_size_t(0)
    assert lib.ssym_stwo_wit_skeleton(C.byref(cfg), name, None, C.byref(ns), None, C.byref(nn)) == 0
    skel = np.zeros(ns.value, dtype=np.uint8)
    slots = np.zeros(max(nn.value, 1), dtype=np.uint32)
    assert lib.ssym_stwo_wit_skeleton(C.byref(cfg), name, C.c_void_p(skel.ctypes.data), C.byref(ns), C.c_void_p(slots.ctypes.data), C.byref(nn)) == 0
    return skel.tobytes().decode(), slots[: nn.value]


@pytest.mark.parametrize("preset", ["prod", "testing"])
def test_skeleton_matches_reference_fixture(S, preset):
    """Tokens of the generator's output == skeleton; literals scattered through the slots == the packed record."""
    cfg = S.stwo_config(preset, S.MODE_REF_LITERAL)
    lo = S.stwo_layout(cfg)
    text = open(os.path.join(GOLDEN, f"stwo_proof_{preset}.wit")).read()
    wit = json.loads(text)
    rec = np.zeros(lo.stride_words, dtype=np.uint32)
    for k, name in enumerate(NAMES):
        skel, slots = skeleton(S, cfg, k)
        toks = W.tokenize(wit[name]["value"])
        classes = "".join("L" if t == "list!" else "N" if t[0].isdigit() else t for t in toks)
        assert classes == skel, name
        lits = [int(t, 0) for t in toks if t[0].isdigit()]
        assert len(lits) == len(slots)
        for v, slot in zip(lits, slots):
            off, kw = int(slot) & 0x0FFFFFFF, {0: 1, 1: 2, 2: 8}[int(slot) >> 28]
            assert v < 1 << (32 * kw)
            rec[off:off + kw] = [(v >> (32 * (kw - 1 - j))) & 0xFFFFFFFF for j in range(kw)]
    expect, bad = S.witness.pack_stwo_wits([text], cfg)
    assert not bad[0] and (rec == expect).all()
    o_rec, o_bad = W.pack_stwo(W.load_wit(text), cfg.n_queries, cfg.n_fri_layers, cfg.lde_log)
    assert not o_bad and (rec == o_rec).all()


def test_skeleton_query_and_errors(S):
    lib = S.load()
    cfg = S.stwo_config("prod", 0)
    ns, nn = C.c_size_t(1), C.c_size_t(1)
    buf = np.zeros(1, dtype=np.uint8)
    assert lib.ssym_stwo_wit_skeleton(C.byref(cfg), 4, C.c_void_p(buf.ctypes.data), C.byref(ns), None, C.byref(nn)) == -4  # SSYM_ERR_NOMEM
    assert ns.value > 1000 and nn.value == 9 * 16 * 4 + 16 * sum(12 - l for l in range(9))
    assert lib.ssym_stwo_wit_skeleton(C.byref(cfg), 6, None, C.byref(ns), None, C.byref(nn)) == -1


# ---- GPU --------------------------------------------------------------------------------------------------
def _dumps(wit, **kw):
    """json.dumps of NAME -> {"value": ...} with the "type" member serde's WitnessValues requires (simfony-cli/src/main.rs:77-81) added where a
    variant did not carry one."""
    return json.dumps({k: ({"type": "", **v} if isinstance(v, dict) else v) for k, v in wit.items()}, **kw)


def _variants(text):
    """(label, text, on the fast path?) — every variant has a defined result under the host parser."""
    wit = json.loads(text)
    out = [("generator output", text, True)]
    out.append(("pretty-printed JSON, reordered names", _dumps({k: {"value": wit[k]["value"]} for k in reversed(NAMES)}, indent=2), True))
    spaced = {k: {"value": "  " + re.sub(r"([(\[\]),])", r"  \1   ", v["value"]) + "  ", "type": "x"} for k, v in wit.items()}
    out.append(("spaces around all tokens", _dumps(spaced), True))
    nl = {k: {"value": re.sub(r"([(\[,])", r"\1 \n\t", v["value"]), "type": "x"} for k, v in wit.items()}
    out.append(("newlines / tabs in the value text (JSON escapes)", _dumps(nl), False))
    out.append(("type member before value", _dumps({k: {"type": "u32", "value": v["value"]} for k, v in wit.items()}), True))
    hexed = dict(wit)
    hexed["POW_NONCE"] = {"value": hex(int(wit["POW_NONCE"]["value"]))}
    hexed["COMMITMENTS"] = {"value": re.sub(r"0x0*", "0x", wit["COMMITMENTS"]["value"])}
    out.append(("hex nonce, digests without leading zeros", _dumps(hexed), True))
    dec = dict(wit)
    dec["COMMITMENTS"] = {"value": "(" + ", ".join(str(int(x, 16)) for x in re.findall(r"0x[0-9a-f]+", wit["COMMITMENTS"]["value"])) + ")"}
    out.append(("decimal u256 literals", _dumps(dec), True))
    up = dict(wit)
    up["COMMITMENTS"] = {"value": wit["COMMITMENTS"]["value"].upper().replace("0X", "0x")}
    out.append(("upper-case hex", _dumps(up), False))
    us = dict(wit)
    us["POW_NONCE"] = {"value": wit["POW_NONCE"]["value"][0] + "_" + wit["POW_NONCE"]["value"][1:]}
    out.append(("digit separators", _dumps(us), False))
    tc = dict(wit)
    tc["COMMITMENTS"] = {"value": wit["COMMITMENTS"]["value"][:-1] + ",)"}
    out.append(("trailing comma", _dumps(tc), False))
    par = dict(wit)
    par["POW_NONCE"] = {"value": "(" + wit["POW_NONCE"]["value"] + ")"}
    out.append(("parenthesised value", _dumps(par), False))
    esc = _dumps(wit).replace('"value": "(', '"value": "\\u0028', 1)
    out.append(("JSON escape", esc, False))
    extra = dict(wit)
    extra["EXTRA"] = {"value": "7"}
    out.append(("unknown extra witness", _dumps(extra), False))
    return out


def _bad_variants(text):
    """(label, text, expected flag) — witnesses `simfony run` would refuse."""
    wit = json.loads(text)
    out = []
    longer = dict(wit)
    longer["DECOMMITMENTS"] = {"value": wit["DECOMMITMENTS"]["value"].replace("list![", "list![0x01, ", 1)}
    out.append(("one sibling too many", _dumps(longer), 1))
    shorter = dict(wit)
    shorter["FRI_DECOMMITMENTS"] = {"value": re.sub(r"list!\[0x[0-9a-f]+, ", "list![", wit["FRI_DECOMMITMENTS"]["value"], count=1)}
    out.append(("one FRI sibling too few", _dumps(shorter), 1))
    big = dict(wit)
    big["OODS_EVALS"] = {"value": wit["OODS_EVALS"]["value"].replace("((1, 0)", "((4294967296, 0)", 1)}
    assert big["OODS_EVALS"] != wit["OODS_EVALS"]
    out.append(("u32 literal out of range", _dumps(big), 2))
    out.append(("missing witness", _dumps({k: v for k, v in wit.items() if k != "POW_NONCE"}), 2))
    out.append(("no type member", json.dumps({k: {"value": v["value"]} for k, v in wit.items()}), 2))
    out.append(("trailing bytes after the object", text + "0", 2))
    out.append(("truncated file", text[: len(text) // 2], 2))
    out.append(("not JSON", "hello", 2))
    out.append(("empty", "", 2))
    junk = dict(wit)
    junk["COMMITMENTS"] = {"value": wit["COMMITMENTS"]["value"].replace(", ", "; ", 1)}
    out.append(("bad separator", _dumps(junk), 2))
    stray = dict(wit)
    stray["DECOMMITMENTS"] = {"value": wit["DECOMMITMENTS"]["value"].replace("list![", "ist![", 1)}
    out.append(("broken list! keyword", _dumps(stray), 2))
    arr = dict(wit)
    arr["COMMITMENTS"] = {"value": wit["COMMITMENTS"]["value"].replace("(", "[").replace(")", "]")}
    out.append(("array where a tuple is expected", _dumps(arr), 2))
    return out


def _host_reference(S, cfg, texts):
    lo = S.stwo_layout(cfg)
    lib = S.load()
    packed = np.zeros((len(texts), lo.stride_words), dtype=np.uint32)
    flags = np.zeros(len(texts), dtype=np.uint32)
    for i, t in enumerate(texts):
        raw = t.encode()
        shape = C.c_int(0)
        rc = lib.ssym_stwo_pack_wit(C.byref(cfg), raw, len(raw), C.c_void_p(packed[i].ctypes.data), C.byref(shape))
        flags[i] = 2 if rc else 1 if shape.value else 0
    return packed, flags


@pytest.mark.gpu
@pytest.mark.parametrize("preset", ["prod", "testing"])
def test_gpu_tokeniser_matches_host_parser(S, preset):
    cfg = S.stwo_config(preset, S.MODE_PROVER_CONSISTENT)
    text = open(os.path.join(GOLDEN, f"stwo_proof_{preset}.wit")).read()
    good, bad = _variants(text), _bad_variants(text)
    texts = [t for _, t, _ in good] + [t for _, t, _ in bad]
    ref_packed, ref_flags = _host_reference(S, cfg, texts)
    assert (ref_flags[: len(good)] == 0).all(), [g[0] for g, f in zip(good, ref_flags) if f]
    assert list(ref_flags[len(good):]) == [f for _, _, f in bad], list(zip([b[0] for b in bad], ref_flags[len(good):]))
    # the Python reader agrees with the host parser on every well-formed variant
    for (label, t, _), rec in zip(good, ref_packed):
        o_rec, o_bad = W.pack_stwo(W.load_wit(t), cfg.n_queries, cfg.n_fri_layers, cfg.lde_log)
        assert not o_bad and (o_rec == rec).all(), label
    ver = S.Verifier(0)
    blob, offsets = S.witness.concat_wit_texts(texts)
    packed, flags = ver.stwo_pack_wit_batch(blob, offsets, cfg)
    assert list(flags) == list(ref_flags)
    assert (packed == ref_packed).all()
    # device-resident text: same result
    import torch

    d_packed, d_flags = ver.stwo_pack_wit_batch(torch.from_numpy(blob).cuda(), torch.from_numpy(offsets.astype(np.int64)).cuda(), cfg)
    assert (d_flags.cpu().numpy().view(np.uint32) == ref_flags).all()
    assert (d_packed.cpu().numpy().view(np.uint32) == ref_packed).all()
    # text -> accept bits, against the oracle on the host-packed records
    accept, status, vflags = ver.stwo_verify_wit_batch(blob, offsets, cfg, want_status=True, want_flags=True)
    orc = O.Oracle()
    ocfg = O.StwoConfig(cfg.trace_log, cfg.lde_log, cfg.n_queries, cfg.n_fri_layers, cfg.mode, cfg.n_columns, cfg.pow_target)
    o_accept, o_status, _ = orc.stwo_verify_batch(ocfg, ref_packed.ravel(), len(texts))
    o_status = o_status.copy()
    o_status[ref_flags != 0] |= 1 << 31
    assert (status == o_status).all()
    assert list(vflags) == list(ref_flags)
    for i in range(len(texts)):
        assert ((int(accept[i // 32]) >> (i % 32)) & 1) == int(o_status[i] == 0), i
    assert (status[: len(good)] == 0).all()  # PROVER_CONSISTENT accepts the fixture in every formatting
    acc_list, st = ver.run_stwo_wit(texts, preset, S.MODE_PROVER_CONSISTENT)
    assert acc_list == [bool(s == 0) for s in o_status]
    ver.close()


@pytest.mark.gpu
def test_gpu_tokeniser_fast_path_is_taken(S):
    """The generator's formatting must not fall back to the host parser: with ssym_set_wit_host_fallback(0) a witness that leaves the fast path
    keeps its flag."""
    cfg = S.stwo_config("prod", S.MODE_PROVER_CONSISTENT)
    text = open(os.path.join(GOLDEN, "stwo_proof_prod.wit")).read()
    ver = S.Verifier(0)
    ver.set_wit_host_fallback(False)
    try:
        good = _variants(text)
        blob, offsets = S.witness.concat_wit_texts([t for _, t, _ in good])
        _, flags = ver.stwo_pack_wit_batch(blob, offsets, cfg)
        assert [int(f) == 0 for f in flags] == [fast for _, _, fast in good], list(zip([g[0] for g in good], flags))
    finally:
        ver.set_wit_host_fallback(True)
    ver.close()


@pytest.mark.gpu
def test_wit_batch_of_distinct_proofs_round_trip(S):
    """Prover -> packed -> `.wit` text (generator syntax) -> GPU tokeniser == the packed proofs; spans several streamed chunks."""
    cfg = S.stwo_config("prod", S.MODE_PROVER_CONSISTENT)
    ver = S.Verifier(0)
    n_distinct, n = 24, 1100
    proofs = ver.stwo_prove_batch(np.arange(100, 100 + n_distinct, dtype=np.uint64), cfg)
    texts = [json.dumps(S.witness.stwo_wit_from_packed(proofs[i], cfg)) for i in range(n_distinct)]
    order = [(7 * i) % n_distinct for i in range(n)]
    bad_at = {5: '{"COMMITMENTS": 3}', 600: texts[3].replace("list![", "list![0x5, ", 1), 1099: ""}
    batch = [bad_at.get(i, texts[order[i]]) for i in range(n)]
    blob, offsets = S.witness.concat_wit_texts(batch)
    packed, flags = ver.stwo_pack_wit_batch(blob, offsets, cfg)
    expect_flags = np.zeros(n, dtype=np.uint32)
    expect_flags[5], expect_flags[600], expect_flags[1099] = 2, 1, 2
    assert (flags == expect_flags).all()
    expect = proofs[order].copy()
    expect[list(bad_at)] = 0
    assert (packed == expect).all()
    accept, status, _ = ver.stwo_verify_wit_batch(blob, offsets, cfg, want_status=True)
    ok = np.array([(int(accept[i // 32]) >> (i % 32)) & 1 for i in range(n)], dtype=bool)
    assert (ok == (expect_flags == 0)).all()
    assert (status[expect_flags == 0] == 0).all() and (status[expect_flags != 0] >> 31 == 1).all()
    ver.close()


@pytest.mark.gpu
def test_cli_gpu_ingestion_matches_host_pack(S, tmp_path):
    """bin/verify-batch (the `simfony run --witness` counterpart): GPU ingestion (default) and --host-pack print the same verdicts."""
    import subprocess

    from conftest import ROOT

    cli = os.path.join(ROOT, "stark-symphony_b200", "bin", "verify-batch")
    good = os.path.join(GOLDEN, "stwo_proof_prod.wit")
    text = open(good).read()
    bad1, bad2 = tmp_path / "long_path.wit", tmp_path / "garbage.wit"
    bad1.write_text(_bad_variants(text)[0][1])
    bad2.write_text("{}")
    outs = []
    for extra in ([], ["--host-pack"]):
        r = subprocess.run([cli, "--program", "stwo", "--mode", "prover-consistent", "--witness", good, "--replicate", "70"] + extra, capture_output=True, text=True)
        assert r.returncode == 0 and r.stdout.count("accept") == 70 and "reject" not in r.stdout, r.stderr
        r = subprocess.run([cli, "--program", "stwo", "--mode", "prover-consistent", "--witness", good, str(bad1), str(bad2), good] + extra, capture_output=True, text=True)
        assert r.returncode == 1 and "Error: Failed to run program" in r.stderr
        lines = r.stdout.strip().splitlines()
        assert [l.split()[0] for l in lines] == ["accept", "reject", "reject", "accept"]
        assert all("status=0x8" in l for l in lines[1:3])  # SSYM_ST_SHAPE
        outs.append(r.stdout)
        r = subprocess.run([cli, "--program", "stwo", "--mode", "ref-literal", "--witness", good] + extra, capture_output=True, text=True)
        assert r.returncode == 1 and r.stdout.startswith("reject")  # the reference at HEAD rejects its own fixture
    assert outs[0] == outs[1]
    s101 = os.path.join(GOLDEN, "stark101_proof.wit")
    bad3 = tmp_path / "s101_bad.wit"
    bad3.write_text(open(s101).read().replace("2133065320", "2133065321"))  # wrong last layer: well-typed, rejected by fri.simf:90
    outs = []
    for extra in ([], ["--host-pack"]):
        r = subprocess.run([cli, "--program", "stark101", "--witness", s101, str(bad3), s101, "--replicate", "11"] + extra, capture_output=True, text=True)
        lines = r.stdout.strip().splitlines()
        assert r.returncode == 1 and [l.split()[0] for l in lines] == ["accept", "reject", "accept"] * 11, r.stderr
        assert all(int(l.split("status=")[1], 16) & 0x100 for l in lines[1::3])  # SSYM_S101_ST_LAST among the failed checks
        outs.append(r.stdout)
    assert outs[0] == outs[1]


@pytest.mark.gpu
def test_cli_cost_output_equals_oracle_counters(S, tmp_path):
    """verify-batch --cost: the per-proof program cost printed by the CLI — the closed-form model of csrc/cost.cpp fed with
    the queries the GPU transcript drew — equals the oracle's counters for that proof, field by field, in both semantics and both presets."""
    import ctypes as C
    import subprocess

    from conftest import ROOT
    from test_oracle_fixtures import load_stwo

    cli = os.path.join(ROOT, "stark-symphony_b200", "bin", "verify-batch")
    orc = O.Oracle()
    for preset in ("testing", "prod"):
        for mode_name, mode in (("ref-literal", O.MODE_REF_LITERAL), ("prover-consistent", O.MODE_PROVER_CONSISTENT)):
            wit = os.path.join(GOLDEN, f"stwo_proof_{preset}.wit")
            r = subprocess.run([cli, "--program", "stwo", "--preset", preset, "--mode", mode_name, "--witness", wit, "--replicate", "3", "--cost"], capture_output=True, text=True)
            assert r.returncode == (0 if mode == O.MODE_PROVER_CONSISTENT else 1), r.stderr
            lines = [l for l in r.stdout.splitlines() if l.startswith("cost ")]
            assert len(lines) == 3 and len(set(lines)) == 1
            got = {kv.split("=")[0]: int(kv.split("=")[1]) for kv in lines[0].split()[2:]}
            want = (C.c_uint64 * 14)()
            orc.lib.oracle_cost_reset()
            orc.stwo_verify_batch(O.make_config(preset, mode), load_stwo(preset), 1)
            orc.lib.oracle_cost_counts(want)
            assert [got[k] for k in S._lib.COST_FIELDS] == [int(x) for x in want], (preset, mode_name)
            assert "as jets: multiply_32=" in r.stderr
    r = subprocess.run([cli, "--program", "stark101", "--witness", os.path.join(GOLDEN, "stark101_proof.wit"), "--cost"], capture_output=True, text=True)
    assert r.returncode == 2 and "models the stwo program only" in r.stderr


@pytest.mark.gpu
def test_wit_batch_edge_cases(S):
    cfg = S.stwo_config("testing", S.MODE_PROVER_CONSISTENT)
    text = open(os.path.join(GOLDEN, "stwo_proof_testing.wit")).read()
    ver = S.Verifier(0)
    lo = S.stwo_layout(cfg)
    # empty batch
    packed, flags = ver.stwo_pack_wit_batch(np.zeros(0, dtype=np.uint8), np.zeros(1, dtype=np.uint64), cfg)
    assert packed.shape == (0, lo.stride_words) and flags.size == 0
    # zero-length witnesses between good ones; 33 witnesses (bitmap word boundary)
    batch = [text if i % 3 else "" for i in range(33)]
    blob, offsets = S.witness.concat_wit_texts(batch)
    accept, status, flags = ver.stwo_verify_wit_batch(blob, offsets, cfg, want_status=True, want_flags=True)
    assert list(flags) == [0 if i % 3 else 2 for i in range(33)]
    bits = np.unpackbits(accept.view(np.uint8), bitorder="little")[:33]
    assert list(bits) == [1 if i % 3 else 0 for i in range(33)]
    assert all((status[i] == 0) == bool(i % 3) for i in range(33))
    # offsets must be non-decreasing
    bad = offsets.copy()
    bad[4] = bad[7]
    with pytest.raises(S.SsymError):
        ver.stwo_pack_wit_batch(blob, bad, cfg)
    # a witness of the OTHER preset is well-formed JSON of the wrong shape: the host parser decides (shape or parse), never a crash
    prod = open(os.path.join(GOLDEN, "stwo_proof_prod.wit")).read()
    blob, offsets = S.witness.concat_wit_texts([text, prod, text])
    _, flags = ver.stwo_pack_wit_batch(blob, offsets, cfg)
    assert flags[0] == 0 and flags[2] == 0 and flags[1] in (1, 2)
    _, ref_flags = _host_reference(S, cfg, [text, prod, text])
    assert list(flags) == list(ref_flags)
    ver.close()


def _s101_variants(text):
    wit = json.loads(text)
    names = ["P_MT_ROOT", "P_EVALS", "FRI_LAYERS", "FRI_LAST_LAYER"]
    out = [("generator output", text, True)]
    out.append(("pretty-printed, reordered", _dumps({k: {"value": wit[k]["value"]} for k in reversed(names)}, indent=1), True))
    out.append(("spaces around tokens", _dumps({k: {"value": " " + re.sub(r"([(\[\]),])", r" \1  ", v["value"]) + " "} for k, v in wit.items()}), True))
    hexed = dict(wit)
    hexed["P_MT_ROOT"] = {"value": hex(int(wit["P_MT_ROOT"]["value"]))}
    hexed["FRI_LAST_LAYER"] = {"value": hex(int(wit["FRI_LAST_LAYER"]["value"]))}
    out.append(("hex root and last layer", _dumps(hexed), True))
    single = dict(wit)
    single["FRI_LAYERS"] = {"value": wit["FRI_LAYERS"]["value"].replace("((", "(").replace("))", ")")}
    out.append(("FRI layers without the doubled parentheses", _dumps(single), False))
    fewer = dict(wit)  # one FRI layer dropped: well-typed, another shape (and a proof the verifier rejects)
    fewer["FRI_LAYERS"] = {"value": "list![" + wit["FRI_LAYERS"]["value"][len("list!["):].split(")), ((", 1)[1].join(["((", ""])}
    out.append(("another shape: first FRI layer dropped", _dumps(fewer), False))
    shorter = dict(wit)
    shorter["P_EVALS"] = {"value": re.sub(r"list!\[\d+, ", "list![", wit["P_EVALS"]["value"], count=1)}
    out.append(("another shape: one trace sibling dropped", _dumps(shorter), False))
    bad = [("garbage", "{]", 2), ("no type member", json.dumps({k: {"value": v["value"]} for k, v in wit.items()}), 2), ("missing witness", _dumps({k: v for k, v in wit.items() if k != "P_EVALS"}), 2),
           ("u32 out of range", _dumps({**wit, "FRI_LAST_LAYER": {"value": str(1 << 32)}}), 2),
           ("u256 out of range", _dumps({**wit, "P_MT_ROOT": {"value": str(1 << 256)}}), 2), ("empty", "", 2)]
    return out, bad


@pytest.mark.gpu
def test_stark101_wit_on_gpu(S):
    """stark101 witness texts tokenised on the GPU (ssym_stark101_verify_wit_batch) against host parser + packed verify + oracle."""
    text = open(os.path.join(GOLDEN, "stark101_proof.wit")).read()
    good, bad = _s101_variants(text)
    ver = S.Verifier(0)
    orc = O.Oracle()
    for lead_with_odd in (False, True):  # the batch's shape comes from its first parseable witness: also lead with the other shape
        items = good + [(l, t, False) for l, t, _ in bad]
        if lead_with_odd:
            items = [items[5]] + items[:5] + items[6:]
        texts = [t for _, t, _ in items] + [text] * 40
        blob, offsets, hostbad = S.witness.pack_stark101_wits(texts)
        o_accept, o_status, _ = orc.s101_verify_batch(blob, offsets)
        o_status = o_status.copy()
        o_status[hostbad] |= 1 << 31
        tblob, toffs = S.witness.concat_wit_texts(texts)
        accept, status, flags = ver.stark101_verify_wit_batch(tblob, toffs, want_status=True, want_flags=True)
        assert (status == o_status).all(), [(items[i][0] if i < len(items) else "fixture", hex(status[i]), hex(o_status[i])) for i in np.nonzero(status != o_status)[0][:6]]
        assert list(flags) == [2 if b else 0 for b in hostbad]
        bits = np.unpackbits(accept.view(np.uint8), bitorder="little")[: len(texts)]
        assert (bits == (o_status == 0)).all()
        assert status[len(items):].tolist() == [0] * 40 and (status[[i for i, it in enumerate(items) if it[0] == "generator output"]] == 0).all()
        import torch

        d_acc, d_st, d_fl = ver.stark101_verify_wit_batch(torch.from_numpy(tblob).cuda(), torch.from_numpy(toffs.astype(np.int64)).cuda(), want_status=True, want_flags=True)
        assert (d_st.cpu().numpy().view(np.uint32) == o_status).all() and (d_acc.cpu().numpy().view(np.uint32) == accept).all()
        acc_list, st = ver.run_stark101_wit(texts)
        assert acc_list == [bool(x == 0) for x in o_status]
    # which variants stay on the GPU
    ver.set_wit_host_fallback(False)
    try:
        tblob, toffs = S.witness.concat_wit_texts([t for _, t, _ in good])
        _, _, flags = ver.stark101_verify_wit_batch(tblob, toffs, want_flags=True)
        assert [int(f) == 0 for f in flags] == [fast for _, _, fast in good], list(zip([g[0] for g in good], flags))
    finally:
        ver.set_wit_host_fallback(True)
    ver.close()
