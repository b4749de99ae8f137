"""NUM_COLUMNS (stwo-verifier/src/config.simf:14) other than the 4 of reference HEAD: the same wide-Fibonacci AIR
(constraints/wide_fibonacci.simf:24-62 folds `c_i - (c_{i-1}^2 + c_{i-2}^2)` over every column i >= 2), 8 or 16 columns wide
(SURVEY.md section 8f rank 3).  The reference has no fixture for these widths, so — as for BASELINE configs 3 and 5 — the inputs
come from the CPU reference prover, which the restated verifier pins: its proofs are accepted under PROVER_CONSISTENT and rejected
under REF_LITERAL exactly where the shipped 4-column fixtures are.  GPU tests compare the whole trace with the oracle, byte for byte."""
import ctypes as C
import json

import numpy as np
import pytest

from oracle import oracle as O
from oracle import witparse as W

P = 2**31 - 1
WIDTHS = [8, 16]


@pytest.fixture(scope="module")
def S():
    import stark_symphony_b200 as S

    S.load()
    return S


def ocfg(cfg):
    return O.StwoConfig(cfg.trace_log, cfg.lde_log, cfg.n_queries, cfg.n_fri_layers, cfg.mode, cfg.n_columns, cfg.pow_target)


# ---- CPU: oracle, layout, host parser -------------------------------------------------------------------------------
@pytest.mark.parametrize("nc", WIDTHS)
@pytest.mark.parametrize("preset,seeds", [("testing", list(range(12))), ("prod", [0, 0xDEADBEEF])])
def test_wide_proofs_accept_and_reject_like_the_fixtures(orc, preset, seeds, nc):
    cfg = O.make_config(preset, O.MODE_PROVER_CONSISTENT, nc)
    pk = orc.stwo_prove_batch(cfg, seeds, threads=4)
    _, status, traces = orc.stwo_verify_batch(cfg, pk.ravel(), len(seeds), want_trace=True)
    assert (status == 0).all(), [hex(s) for s in status]
    lit = O.make_config(preset, O.MODE_REF_LITERAL, nc)
    _, st_lit, tr_lit = orc.stwo_verify_batch(lit, pk.ravel(), len(seeds), want_trace=True)
    for s, t in zip(st_lit, tr_lit):
        assert s & (1 << 7) and s & (1 << 17) and (t.first_fail >> 16) == 7  # FRI layer-0 Merkle root first, F2
    for a, b in zip(traces, tr_lit):  # everything before fri_answers is mode independent
        assert bytes(a.digest_pow) == bytes(b.digest_pow) and list(a.queries) == list(b.queries)
    # a proof of one width is not a proof of another: the 4-column verifier reads another layout
    if preset == "testing":
        other = O.make_config(preset, O.MODE_PROVER_CONSISTENT, 4)
        lo4 = orc.stwo_layout(other)
        _, st4, _ = orc.stwo_verify_batch(other, np.ascontiguousarray(pk[:, :lo4.stride_words]).ravel(), len(seeds))
        assert (st4 != 0).all()


@pytest.mark.parametrize("nc", WIDTHS)
def test_wide_trace_rows_satisfy_the_air(orc, nc):
    row = (C.c_uint32 * nc)()
    for seed in (0, 5, 2**40):
        for r in (0, 1, 511):
            orc.lib.oracle_stwo_trace_row_n(C.c_uint64(seed), C.c_uint32(r), C.c_uint32(nc), row)
            c = [int(x) for x in row]
            assert c[0] == 1 and c[1] < P
            for i in range(2, nc):
                assert c[i] == (c[i - 1] ** 2 + c[i - 2] ** 2) % P


@pytest.mark.parametrize("nc", WIDTHS)
def test_every_section_of_a_wide_proof_is_load_bearing(orc, nc):
    cfg = O.make_config("prod", O.MODE_PROVER_CONSISTENT, nc)
    lo = orc.stwo_layout(cfg)
    pk = orc.stwo_prove_batch(cfg, [3])[0]
    qv = nc + 16
    offs = [lo.off_commit + 8, lo.off_commit + 16, lo.off_oods_trace + 4 * (nc - 1) + 3, lo.off_oods_cp + 5, lo.off_fri_first_root, lo.off_last_coeff,
            lo.off_qvals + nc - 1, lo.off_qvals + nc, lo.off_qvals + qv * 15 + qv - 1, lo.off_trace_sib + 3, lo.off_cp_sib + 9, lo.off_fri_wit,
            lo.off_fri_sib[0] + 2, lo.off_fri_inner_root, lo.off_fri_sib[1]]
    recs = []
    for o in offs:
        r = pk.copy()
        r[o] ^= 1
        recs.append(r)
    _, status, _ = orc.stwo_verify_batch(cfg, np.concatenate(recs), len(recs))
    assert (status != 0).all(), [hex(s) for s in status]


def test_layout_of_wide_configurations(S, orc):
    for nc, alg in ((4, 54488), (8, 54808), (16, 55448)):  # + 16 B of OODS sample and + 4 B per query per extra column
        for preset in ("prod", "testing"):
            cfg = S.stwo_config(preset, 1, n_columns=nc)
            lo = S.stwo_layout(cfg)
            assert bytes(lo) == bytes(orc.stwo_layout(ocfg(cfg)))
            wl = W.stwo_layout(cfg.n_queries, cfg.n_fri_layers, cfg.lde_log, nc)
            assert wl["stride_words"] == lo.stride_words and wl["qvals"] == lo.off_qvals and wl["fri_sib"][0] == lo.off_fri_sib[0]
            if preset == "prod":
                assert lo.algorithmic_bytes == alg
    zero = S.stwo_config("prod", 0, n_columns=0)  # 0 = the reference's 4
    assert bytes(S.stwo_layout(zero)) == bytes(S.stwo_layout(S.stwo_config("prod", 0)))
    for bad in (1, 2, 3, 5, 12, 32):
        with pytest.raises(S.SsymError):
            S.stwo_layout(S.stwo_config("prod", 0, n_columns=bad))


@pytest.mark.parametrize("nc", WIDTHS)
def test_wide_wit_text_round_trip_on_the_host(S, orc, nc):
    """packed proof -> `.wit` text (the value syntax of stwo-verifier/scripts/generate_wit.py:139-243 with C-column arrays) -> host parser
    (csrc/witness.cpp) and the independent Python reader: both give the packed proof back; a 4-column text is ill-typed for a wide program."""
    for preset in ("testing", "prod"):
        cfg = S.stwo_config(preset, S.MODE_PROVER_CONSISTENT, n_columns=nc)
        pk = orc.stwo_prove_batch(ocfg(cfg), [11, 12])
        texts = [json.dumps(S.witness.stwo_wit_from_packed(pk[i], cfg)) for i in range(2)]
        packed, bad = S.witness.pack_stwo_wits(texts, cfg)
        assert not bad.any() and (packed.reshape(2, -1) == pk).all()
        for i in range(2):
            ref, rej = W.pack_stwo(W.load_wit(texts[i]), cfg.n_queries, cfg.n_fri_layers, cfg.lde_log, nc)
            assert not rej and (ref == pk[i]).all()
        _, bad4 = S.witness.pack_stwo_wits(texts[:1], S.stwo_config(preset, S.MODE_PROVER_CONSISTENT))
        assert bad4[0]
        # the skeleton the GPU tokeniser follows has C single-element arrays in both column lists
        n = C.c_size_t(0)
        m = C.c_size_t(0)
        S.load().ssym_stwo_wit_skeleton(C.byref(cfg), 2, None, C.byref(n), None, C.byref(m))
        assert m.value == 4 * nc + 64


# ---- GPU ----------------------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def ver(S):
    return S.Verifier(0)


def _assert_same(S, ver, orc, cfg, batch, n):
    accept, status, traces = ver.stwo_verify_batch(batch, cfg, n, want_status=True, want_trace=True)
    o_accept, o_status, o_traces = orc.stwo_verify_batch(ocfg(cfg), batch, n, want_trace=True)
    assert (status == o_status).all(), [(hex(a), hex(b)) for a, b in zip(status, o_status) if a != b][:4]
    assert (accept == o_accept).all()
    for i in range(n):
        assert bytes(memoryview(traces[i]).cast("B")) == bytes(memoryview(o_traces[i]).cast("B")), i
    return status


@pytest.mark.gpu
@pytest.mark.parametrize("nc", WIDTHS)
@pytest.mark.parametrize("preset", ["testing", "prod"])
def test_gpu_verifier_matches_oracle_on_wide_proofs(S, ver, orc, preset, nc):
    """Honest proofs, one corrupted proof per class and random single-bit corruptions, both semantics, both Merkle schedules."""
    cfg = S.stwo_config(preset, S.MODE_PROVER_CONSISTENT, n_columns=nc)
    lo = S.stwo_layout(cfg)
    n_honest = 6
    pk = orc.stwo_prove_batch(ocfg(cfg), list(range(50, 50 + 40)), threads=8)
    recs = [pk[i] for i in range(n_honest)]
    recs += [S.witness.apply_mutation(pk[0], w, d) for (w, d) in S.witness.stwo_negative_classes(cfg).values()]
    rng = np.random.default_rng(nc)
    for k in range(n_honest, 40):
        r = pk[k].copy()
        r[rng.integers(0, lo.stride_words)] ^= np.uint32(1 << rng.integers(0, 32))
        recs.append(r)
    # leaf data of every trace / composition query value (the trace leaf is C words: one block for 8, two for 16)
    qv = nc + 16
    for j in range(0, qv, 3):
        r = pk[1].copy()
        r[lo.off_qvals + qv * (cfg.n_queries - 1) + j] ^= np.uint32(0x80000000)
        recs.append(r)
    batch = np.concatenate(recs)
    n = len(recs)
    for sharing in (0, 2):
        ver.set_merkle_sharing(sharing)
        for mode in (S.MODE_PROVER_CONSISTENT, S.MODE_REF_LITERAL):
            cfg.mode = mode
            status = _assert_same(S, ver, orc, cfg, batch, n)
            if mode == S.MODE_PROVER_CONSISTENT:
                assert (status[:n_honest] == 0).all() and (status[n_honest:n_honest + 9] != 0).all()
    ver.set_merkle_sharing(1)


@pytest.mark.gpu
@pytest.mark.parametrize("nc", WIDTHS)
def test_gpu_wit_ingestion_of_wide_proofs(S, ver, orc, nc):
    cfg = S.stwo_config("prod", S.MODE_PROVER_CONSISTENT, n_columns=nc)
    pk = orc.stwo_prove_batch(ocfg(cfg), [1, 2, 3, 4], threads=4)
    texts = [json.dumps(S.witness.stwo_wit_from_packed(pk[i], cfg)) for i in range(4)]
    texts.append(json.dumps(S.witness.stwo_wit_from_packed(orc.stwo_prove_batch(ocfg(S.stwo_config("prod", 1)), [1])[0], S.stwo_config("prod", 1))))  # 4 columns
    blob, offsets = S.witness.concat_wit_texts(texts)
    packed, flags = ver.stwo_pack_wit_batch(blob, offsets, cfg)
    assert list(flags) == [0, 0, 0, 0, 2] and (packed[:4] == pk).all()
    accept, status, _ = ver.stwo_verify_wit_batch(blob, offsets, cfg, want_status=True)
    assert (status[:4] == 0).all() and status[4] >> 31 == 1 and int(accept[0]) & 31 == 15


@pytest.mark.gpu
@pytest.mark.parametrize("nc", WIDTHS)
@pytest.mark.parametrize("preset,seeds", [("testing", list(range(24)) + [2**64 - 1]), ("prod", [0, 1, 0xDEADBEEF, 2**63 + 5])])
def test_gpu_prover_matches_reference_prover_on_wide_traces(S, ver, orc, preset, seeds, nc):
    cfg = S.stwo_config(preset, S.MODE_PROVER_CONSISTENT, n_columns=nc)
    seeds = np.array(seeds, dtype=np.uint64)
    gpu = ver.stwo_prove_batch(seeds, cfg)
    ref = orc.stwo_prove_batch(ocfg(cfg), seeds, threads=8)
    assert (gpu == ref).all(), np.argwhere(gpu != ref)[:4]
    accept, status, _ = ver.stwo_verify_batch(gpu.ravel(), cfg, len(seeds), want_status=True)
    assert (status == 0).all()


@pytest.mark.gpu
@pytest.mark.parametrize("nc", WIDTHS)
def test_wide_batch_device_resident_with_negatives(S, ver, orc, nc):
    """2048 distinct GPU-proven wide proofs, every 8th corrupted: bitmap == the corruption pattern, a sample of statuses == the oracle."""
    import torch

    cfg = S.stwo_config("prod", S.MODE_PROVER_CONSISTENT, n_columns=nc)
    lo = S.stwo_layout(cfg)
    n = 2048
    seeds = torch.arange(7000, 7000 + n, dtype=torch.int64, device="cuda")
    proofs = ver.stwo_prove_batch(seeds, cfg)
    torch.cuda.synchronize()
    assert proofs.shape == (n, lo.stride_words)
    classes = list(S.witness.stwo_negative_classes(cfg).values())
    bad_rows = list(range(0, n, 8))
    for j, row in enumerate(bad_rows):
        word, delta = classes[j % len(classes)]
        proofs[row, word] += delta
    torch.cuda.synchronize()  # the corruptions run on torch's stream, the verifier on the handle's own
    accept, status, _ = ver.stwo_verify_batch(proofs.view(-1), cfg, n, want_status=True)
    ver.synchronize()
    status = status.cpu().numpy().view(np.uint32)
    expect_bad = np.zeros(n, dtype=bool)
    expect_bad[bad_rows] = True
    assert ((status != 0) == expect_bad).all(), np.flatnonzero((status != 0) != expect_bad)[:10]
    bits = accept.cpu().numpy().view(np.uint32)
    got = np.array([(bits[i // 32] >> (i % 32)) & 1 for i in range(n)], dtype=bool)
    assert (got == ~expect_bad).all()
    sample = proofs[:24].cpu().numpy().view(np.uint32)
    _, o_status, _ = orc.stwo_verify_batch(ocfg(cfg), sample.ravel(), 24)
    assert (o_status == status[:24]).all()


@pytest.mark.gpu
def test_cli_columns_option(S, orc, tmp_path):
    """bin/verify-batch --columns N: `.wit` files written from 8-column proofs verify with --columns 8 (GPU ingestion and --host-pack alike) and
    are refused as ill-typed without it, as `simfony run` refuses a witness that does not match the program's types."""
    import os
    import subprocess

    from conftest import ROOT

    cli = os.path.join(ROOT, "stark-symphony_b200", "bin", "verify-batch")
    cfg = S.stwo_config("prod", S.MODE_PROVER_CONSISTENT, n_columns=8)
    pk = orc.stwo_prove_batch(ocfg(cfg), [21, 22])
    bad = S.witness.apply_mutation(pk[1], *S.witness.stwo_negative_classes(cfg)["oods_trace_plus_1"])
    paths = []
    for name, rec in (("a", pk[0]), ("b", bad), ("c", pk[1])):
        path = tmp_path / f"{name}.wit"
        path.write_text(json.dumps(S.witness.stwo_wit_from_packed(rec, cfg)))
        paths.append(str(path))
    outs = []
    for extra in ([], ["--host-pack"]):
        r = subprocess.run([cli, "--program", "stwo", "--mode", "prover-consistent", "--columns", "8", "--witness"] + paths + extra, capture_output=True, text=True)
        assert r.returncode == 1 and [l.split()[0] for l in r.stdout.strip().splitlines()] == ["accept", "reject", "accept"], r.stderr
        outs.append(r.stdout)
        r = subprocess.run([cli, "--program", "stwo", "--mode", "prover-consistent", "--witness", paths[0]] + extra, capture_output=True, text=True)
        assert r.returncode == 1 and r.stdout.startswith("reject") and "status=0x8" in r.stdout
    assert outs[0] == outs[1]
    r = subprocess.run([cli, "--program", "stwo", "--columns", "5", "--witness", paths[0]], capture_output=True, text=True)
    assert r.returncode == 2 and "n_columns" in r.stderr
