"""GPU parity: the CUDA path, called through the C-ABI (ctypes -> libssym.so), against the CPU oracle on the same
inputs.  Bar: bit-exact (everything here is integer / byte work).  Needs a B200: `pytest -m gpu`."""
import ctypes as C
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def S():
    import stark_symphony_b200 as S

    return S


@pytest.fixture(scope="module")
def ver(S):
    v = S.Verifier(0)
    yield v
    v.close()


def golden_stwo(S, preset, mode):
    cfg = S.stwo_config(preset, mode)
    packed, bad = S.witness.pack_stwo_wits([open(os.path.join(GOLDEN, f"stwo_proof_{preset}.wit")).read()], cfg)
    assert not bad[0]
    return cfg, packed


def ocfg(cfg):
    from oracle import oracle as O

    return O.StwoConfig(cfg.trace_log, cfg.lde_log, cfg.n_queries, cfg.n_fri_layers, cfg.mode, cfg.n_columns, cfg.pow_target)


def trace_bytes(t):
    return bytes(memoryview(t).cast("B")) if not isinstance(t, (bytes, bytearray)) else bytes(t)


def diff_traces(a, b, struct):
    """Field-by-field diff of two ctypes trace structs (for a readable failure)."""
    out = []
    for name, _ in struct._fields_:
        va, vb = getattr(a, name), getattr(b, name)
        ba = bytes(memoryview(va).cast("B")) if hasattr(va, "_length_") else va
        bb = bytes(memoryview(vb).cast("B")) if hasattr(vb, "_length_") else vb
        if ba != bb:
            out.append(name)
    return out


# ---- whole proofs --------------------------------------------------------------------------------------
@pytest.mark.parametrize("preset", ["testing", "prod"])
@pytest.mark.parametrize("mode", [0, 1])
def test_stwo_fixture_trace_bit_exact(S, ver, orc, preset, mode):
    cfg, packed = golden_stwo(S, preset, mode)
    accept, status, traces = ver.stwo_verify_batch(packed, cfg, 1, want_status=True, want_trace=True)
    o_accept, o_status, o_traces = orc.stwo_verify_batch(ocfg(cfg), packed, 1, want_trace=True)
    assert status[0] == o_status[0] and (accept[0] & 1) == (o_accept[0] & 1)
    assert diff_traces(traces[0], o_traces[0], S.StwoTrace) == []
    assert trace_bytes(traces[0]) == trace_bytes(o_traces[0])
    if mode == 1:
        assert status[0] == 0
    else:
        assert traces[0].first_fail == 7 << 16  # FRI layer-0 Merkle root: the reference rejects its own fixture at HEAD


def negatives_batch(S, cfg, packed, extra_random=0, seed=0):
    lo = S.stwo_layout(cfg)
    classes = S.witness.stwo_negative_classes(cfg)
    recs = [packed] + [S.witness.apply_mutation(packed, w, d) for (w, d) in classes.values()]
    rng = np.random.default_rng(seed)
    for _ in range(extra_random):  # random single-word corruption anywhere in the payload
        w = int(rng.integers(0, lo.stride_words))
        recs.append(S.witness.apply_mutation(packed, w, int(rng.integers(1, 2**32))))
    return np.concatenate(recs), len(recs), list(classes)


@pytest.mark.parametrize("preset", ["testing", "prod"])
@pytest.mark.parametrize("mode", [0, 1])
def test_stwo_negatives_match_oracle(S, ver, orc, preset, mode):
    cfg, packed = golden_stwo(S, preset, mode)
    batch, n, names = negatives_batch(S, cfg, packed, extra_random=150 if preset == "prod" else 400, seed=mode)
    accept, status, traces = ver.stwo_verify_batch(batch, cfg, n, want_status=True, want_trace=True)
    o_accept, o_status, o_traces = orc.stwo_verify_batch(ocfg(cfg), batch, n, want_trace=True)
    assert (status == o_status).all(), [(i, hex(status[i]), hex(o_status[i])) for i in np.nonzero(status != o_status)[0][:5]]
    assert (accept == o_accept).all()
    for i in range(n):
        assert trace_bytes(traces[i]) == trace_bytes(o_traces[i]), (i, diff_traces(traces[i], o_traces[i], S.StwoTrace))
    if mode == 1:
        assert status[0] == 0 and all(status[1:1 + len(names)] != 0)  # every corruption class is rejected


def test_stwo_shared_node_schedule_matches_oracle(S, ver, orc):
    """The Merkle paths of one tree share nodes above the height where they meet (StwoDedup, csrc/stwo_kernels.cuh): hashed once, the other
    queries take the result only if their own node and their own remaining siblings are bit-identical.  Corrupt, for every tree and every
    query in turn, (a) the sibling just below the root, (b) a sibling in the middle, (c) the leaf sibling, (d) the query's leaf data: the
    whole trace (every recomputed root, every per-query mask, the first failing assert) must be what the per-query oracle gives.  Also
    proofs whose queries collide: duplicated queries cannot be forged (the channel draws them), so their effect is covered by a proof
    with many distinct seeds below."""
    cfg, packed = golden_stwo(S, "prod", 1)
    lo = S.stwo_layout(cfg)
    G, Q, L = cfg.lde_log, cfg.n_queries, cfg.n_fri_layers
    recs = [packed]
    for q in range(Q):
        for base, d in [(lo.off_trace_sib, G), (lo.off_cp_sib, G)] + [(lo.off_fri_sib[l], G - 1 - l) for l in range(L + 1)]:
            for lvl in {d - 1, d // 2, 0}:
                recs.append(S.witness.apply_mutation(packed, base + (q * d + lvl) * 8 + (q + lvl) % 8, 1 << ((3 * q + lvl) % 32)))
        recs.append(S.witness.apply_mutation(packed, lo.off_qvals + 20 * q + 1, 1))       # trace leaf
        recs.append(S.witness.apply_mutation(packed, lo.off_qvals + 20 * q + 4 + q, 1))   # composition leaf
        for l in range(L + 1):
            recs.append(S.witness.apply_mutation(packed, lo.off_fri_wit + (l * Q + q) * 4 + l % 4, 1))  # FRI witness = the sibling leaf
    batch = np.concatenate(recs)
    n = len(recs)
    accept, status, traces = ver.stwo_verify_batch(batch, cfg, n, want_status=True, want_trace=True)
    o_accept, o_status, o_traces = orc.stwo_verify_batch(ocfg(cfg), batch, n, want_trace=True)
    assert status[0] == 0 and (status[1:] != 0).all()
    assert (status == o_status).all(), [(i, hex(status[i]), hex(o_status[i])) for i in np.nonzero(status != o_status)[0][:5]]
    assert (accept == o_accept).all()
    for i in range(n):
        assert trace_bytes(traces[i]) == trace_bytes(o_traces[i]), (i, diff_traces(traces[i], o_traces[i], S.StwoTrace))
    # honest proofs of many seeds (different query patterns, incl. colliding pair indices in the small FRI layers), full traces
    seeds = np.arange(5000, 5096, dtype=np.uint64)
    proofs = ver.stwo_prove_batch(seeds, cfg)
    accept, status, traces = ver.stwo_verify_batch(proofs.ravel(), cfg, len(seeds), want_status=True, want_trace=True)
    o_accept, o_status, o_traces = orc.stwo_verify_batch(ocfg(cfg), proofs.ravel(), len(seeds), want_trace=True)
    assert (status == 0).all() and (o_status == 0).all()
    for i in range(len(seeds)):
        assert trace_bytes(traces[i]) == trace_bytes(o_traces[i]), (i, diff_traces(traces[i], o_traces[i], S.StwoTrace))


@pytest.mark.parametrize("policy", [0, 2])
@pytest.mark.parametrize("mode", [0, 1])
def test_stwo_both_merkle_schedules_in_both_modes(S, orc, policy, mode):
    """ssym_set_merkle_sharing: the per-query kernel and the shared-node schedule, each under both semantics (the default policy picks one per
    mode), give the oracle's traces on the fixture, its negatives and random corruptions."""
    v = S.Verifier(0)
    v.set_merkle_sharing(policy)
    cfg, packed = golden_stwo(S, "prod", mode)
    batch, n, names = negatives_batch(S, cfg, packed, extra_random=120, seed=7 + policy)
    accept, status, traces = v.stwo_verify_batch(batch, cfg, n, want_status=True, want_trace=True)
    o_accept, o_status, o_traces = orc.stwo_verify_batch(ocfg(cfg), batch, n, want_trace=True)
    assert (status == o_status).all() and (accept == o_accept).all()
    for i in range(n):
        assert trace_bytes(traces[i]) == trace_bytes(o_traces[i]), (i, diff_traces(traces[i], o_traces[i], S.StwoTrace))
    with pytest.raises(S.SsymError):
        v.set_merkle_sharing(3)
    v.close()


def test_stwo_device_resident_and_large_batch(S, ver, orc):
    """BASELINE config 2 shape: proof.json replicated x1024 (+ negatives sprinkled in), inputs resident in HBM."""
    import torch

    cfg, packed = golden_stwo(S, "prod", 1)
    lo = S.stwo_layout(cfg)
    n = 1024
    batch = np.tile(packed, n).reshape(n, lo.stride_words)
    classes = list(S.witness.stwo_negative_classes(cfg).values())
    bad_rows = list(range(5, n, 37))
    for k, r in enumerate(bad_rows):
        w, d = classes[k % len(classes)]
        batch[r] = S.witness.apply_mutation(batch[r], w, d)
    dev = torch.from_numpy(batch.view(np.int32).reshape(-1)).cuda()
    accept, status, _ = ver.stwo_verify_batch(dev, cfg, n, want_status=True)
    ver.synchronize()
    accept = accept.cpu().numpy().view(np.uint32)
    status = status.cpu().numpy().view(np.uint32)
    bits = np.unpackbits(accept.view(np.uint8), bitorder="little")[:n].astype(bool)
    expect = np.ones(n, dtype=bool)
    expect[bad_rows] = False
    assert (bits == expect).all()
    assert ((status == 0) == expect).all()
    # host-buffer entry point gives the same bitmap (exercises the chunked H2D pipeline)
    accept_h, status_h, _ = ver.stwo_verify_batch(batch.reshape(-1), cfg, n, want_status=True)
    assert (accept_h == accept).all() and (status_h == status).all()
    # spot-check a slice against the oracle
    sl = slice(0, 48)
    _, o_status, _ = orc.stwo_verify_batch(ocfg(cfg), batch[sl].reshape(-1), 48)
    assert (o_status == status[sl]).all()


def test_stwo_ragged_batch_sizes(S, ver):
    """Batch sizes around the warp / bitmap-word / chunk boundaries, and the empty batch."""
    cfg, packed = golden_stwo(S, "testing", 1)
    lo = S.stwo_layout(cfg)
    for n in (0, 1, 2, 31, 32, 33, 63, 65, 257, 1000):
        batch = np.tile(packed, max(n, 1))[: n * lo.stride_words]
        if n >= 2:
            batch = batch.copy()
            batch[(n - 1) * lo.stride_words + lo.off_last_coeff] ^= 1  # last proof corrupted
        accept, status, _ = ver.stwo_verify_batch(batch, cfg, n, want_status=True)
        bits = np.unpackbits(accept.view(np.uint8), bitorder="little")[:n].astype(bool)
        expect = np.ones(n, dtype=bool)
        if n >= 2:
            expect[-1] = False
        assert (bits == expect).all(), n


@pytest.mark.gpu
def test_host_chunk_plan_sizes(S, ver):
    """Blocking host-buffer calls taper their chunks (csrc/api.cu host_chunk_plan): sizes around the plan's break points, a corrupted proof every 37,
    packed and compact records, blocking and enqueue-only calls — the bitmap of the device-resident call every time."""
    import torch

    cfg, packed = golden_stwo(S, "testing", 1)
    lo = S.stwo_layout(cfg)
    for n in (256, 257, 289, 511, 1023, 1025, 2047, 2049, 2304):
        batch = np.tile(packed, n)
        for r in range(5, n, 37):
            batch[r * lo.stride_words + lo.off_last_coeff] ^= 1
        d_accept, _, _ = ver.stwo_verify_batch(torch.from_numpy(batch.view(np.int32)).cuda(), cfg, n)
        ver.synchronize()
        want = d_accept.cpu().numpy().view(np.uint32)
        expect = np.ones(n, dtype=bool)
        expect[5::37] = False
        assert (np.unpackbits(want.view(np.uint8), bitorder="little")[:n].astype(bool) == expect).all(), n
        blob, offsets = S.witness.compact_stwo(batch, cfg)
        for async_mode in (False, True):
            ver.set_host_async(async_mode)
            a1, _, _ = ver.stwo_verify_batch(batch, cfg, n)
            a2, _ = ver.stwo_verify_compact_batch(blob, offsets, cfg)
            ver.synchronize()
            ver.set_host_async(False)
            assert (a1 == want).all() and (a2 == want).all(), (n, async_mode)


def golden_s101(S):
    return S.witness.pack_stark101_wits([open(os.path.join(GOLDEN, "stark101_proof.wit")).read()])


def test_stark101_fixture_and_negatives(S, ver, orc):
    blob, offs, bad = golden_s101(S)
    assert not bad[0]
    n_words = len(blob)
    rng = np.random.default_rng(3)
    recs = [blob]
    first_layer = 20 + 3 * 13 * 8
    for w, d in [(20 + 5 * 8 + 1, 1), (17, 1), (first_layer + 8, 1), (first_layer + 10, 1), (first_layer + 16 + 3, 1 << 9), (5, 1), (8, 1), (16, 3221225473)]:
        recs.append(S.witness.apply_mutation(blob, w, d))
    for _ in range(300):
        w = int(rng.integers(5, n_words))
        if w in (6, 7, 19):
            continue
        recs.append(S.witness.apply_mutation(blob, w, int(rng.integers(1, 2**32))))
    # structurally different but well-formed records: drop the last FRI layer; drop one sibling of a trace path
    short = blob.copy()
    last_layer_words = 16 + 8 * (int(blob[n_words - (16 + 8 * 8) + 11]) + int(blob[n_words - (16 + 8 * 8) + 12]))
    short = short[: n_words - last_layer_words].copy()
    short[0] = len(short)
    short[1] -= 1
    recs.append(short)
    fewer = np.concatenate([blob[:20 + 12 * 8], blob[20 + 13 * 8:]]).copy()
    fewer[0] = len(fewer)
    fewer[2] = 12
    recs.append(fewer)
    all_blob = np.concatenate(recs)
    offsets = np.zeros(len(recs) + 1, dtype=np.uint64)
    offsets[1:] = np.cumsum([len(r) for r in recs])
    accept, status, traces = ver.stark101_verify_batch(all_blob, offsets, want_status=True, want_trace=True)
    o_accept, o_status, o_traces = orc.s101_verify_batch(all_blob, offsets, want_trace=True)
    assert status[0] == 0 and (status == o_status).all() and (accept == o_accept).all()
    for i in range(len(recs)):
        assert trace_bytes(traces[i]) == trace_bytes(o_traces[i]), (i, diff_traces(traces[i], o_traces[i], S.S101Trace))
    assert all(status[1:9] != 0)


@pytest.mark.parametrize("n", [2048, 40001])  # 40001: a large call is cut into 8192-proof chunks on internal streams (ragged last chunk)
def test_stark101_replicated_device(S, ver, orc, n):
    import torch

    blob, offs, _ = golden_s101(S)
    all_blob = np.tile(blob, n)
    bad = sorted({100, 8191, 8192, 16383, 24576, n - 1} & set(range(n)))
    for r in bad:
        all_blob[r * len(blob) + 5] += 1  # wrong last layer
    offsets = np.arange(n + 1, dtype=np.uint64) * len(blob)
    d_blob = torch.from_numpy(all_blob.view(np.int32)).cuda()
    d_off = torch.from_numpy(offsets.view(np.int64)).cuda()
    accept, status, _ = ver.stark101_verify_batch(d_blob, d_off, want_status=True)
    ver.synchronize()
    bits = np.unpackbits(accept.cpu().numpy().view(np.uint8), bitorder="little")[:n].astype(bool)
    expect = np.ones(n, dtype=bool)
    expect[bad] = False
    assert (bits == expect).all()
    st = status.cpu().numpy().view(np.uint32)
    _, o_status, _ = orc.s101_verify_batch(all_blob[: 101 * len(blob)], offsets[:102])
    assert (st[:101] == o_status).all() and (st[bad] == o_status[100]).all() and (st[expect] == 0).all()


def test_stark101_device_records_that_lie_about_their_length(S, ver):
    """Device-resident records are untrusted too: a record whose length word disagrees with the offsets array, or one shorter than the
    fixed header, is SHAPE-rejected without a single read outside its own [offsets[i], offsets[i+1]) range (memcheck-clean under
    tools/sanitize.sh).  Its neighbours are unaffected."""
    import torch

    blob, _, _ = golden_s101(S)
    L = len(blob)
    recs = [blob.copy() for _ in range(6)]
    recs[1][0] = L + 4096            # declares more words than it has
    recs[2][0] = 0xFFFFFFFF          # ... far more
    recs[3] = blob[:12].copy()       # shorter than the header; its length word still says L
    recs[4][0] = L - 8               # declares fewer
    all_blob = np.concatenate(recs)
    offsets = np.concatenate([[0], np.cumsum([len(r) for r in recs])]).astype(np.uint64)
    d_blob = torch.from_numpy(all_blob.view(np.int32)).cuda()  # exact-size allocation: an over-read would leave the buffer
    d_off = torch.from_numpy(offsets.view(np.int64)).cuda()
    accept, status, _ = ver.stark101_verify_batch(d_blob, d_off, want_status=True)
    ver.synchronize()
    st = status.cpu().numpy().view(np.uint32)
    assert st[0] == 0 and st[5] == 0
    assert all(st[i] != 0 for i in (1, 2, 3, 4))
    assert int(accept.cpu().numpy().view(np.uint32)[0]) & 0x3F == 0b100001


# ---- jets ---------------------------------------------------------------------------------------------------
def rand_u32(rng, n, canonical_frac=0.5):
    P = 2147483647
    x = rng.integers(0, 2**32, n, dtype=np.uint64)
    canon = rng.random(n) < canonical_frac
    x[canon] %= P
    edge = np.array([0, 1, P - 1, P, P + 1, 2**31, 2**32 - 1, 2**32 - 2, 2], dtype=np.uint64)
    x[: len(edge)] = edge
    return x.astype(np.uint32)


def test_m31_jets(ver, orc):
    rng = np.random.default_rng(11)
    n = 4099  # not a multiple of 4: exercises the vector tail
    a, b = rand_u32(rng, n), rand_u32(rng, n)
    rng.shuffle(b)
    for name in ("m31_add", "m31_sub", "m31_mul"):
        got = getattr(ver, name)(a, b)
        want = np.array([getattr(orc, name)(int(x), int(y)) for x, y in zip(a, b)], dtype=np.uint32)
        assert (got == want).all(), name
    got = ver.m31_neg(a)
    assert (got == np.array([orc.m31_neg(int(x)) for x in a], dtype=np.uint32)).all()
    got, failv = ver.m31_inv(a)
    want = [orc.m31_inv(int(x)) for x in a]
    assert (got == np.array([w[0] for w in want], dtype=np.uint32)).all()
    assert (failv.astype(bool) == np.array([w[1] for w in want])).all()


def test_ext_field_jets(ver, orc):
    rng = np.random.default_rng(12)
    n = 1500
    a4, b4 = rand_u32(rng, 4 * n), rand_u32(rng, 4 * n)
    rng.shuffle(b4)
    for name in ("qm31_add", "qm31_sub", "qm31_mul"):
        got = getattr(ver, name)(a4, b4).reshape(n, 4)
        for i in range(n):
            assert list(got[i]) == list(getattr(orc, name)(a4[4 * i:4 * i + 4], b4[4 * i:4 * i + 4])), (name, i)
    got, failv = ver.qm31_inv(a4)
    got = got.reshape(n, 4)
    for i in range(n):
        w, f = orc.qm31_inv(a4[4 * i:4 * i + 4])
        assert list(got[i]) == list(w) and bool(failv[i]) == f, i
    # zero / norm-zero inputs must flag (m31.simf:118-122 through cm31_inv)
    z = np.zeros(8, dtype=np.uint32)
    z[4:] = [2147483647, 0, 0, 0]  # p is not bitwise zero, but its norm is 0 mod p
    got, failv = ver.qm31_inv(z)
    assert list(failv) == [1, 1] and list(got) == [0] * 8
    m = rand_u32(rng, n)
    got = ver.qm31_mul_m31(a4, m).reshape(n, 4)
    for i in range(0, n, 7):
        assert list(got[i]) == list(orc.qm31_mul_m31(a4[4 * i:4 * i + 4], int(m[i])))
    c2 = rand_u32(rng, 2 * n)
    got = ver.qm31_mul_cm31(a4, c2).reshape(n, 4)
    for i in range(0, n, 7):
        assert list(got[i]) == list(orc.qm31_mul_cm31(a4[4 * i:4 * i + 4], c2[2 * i:2 * i + 2]))
    a2, b2 = a4[: 2 * n], b4[: 2 * n]
    got = ver.cm31_mul(a2, b2).reshape(n, 2)
    for i in range(0, n, 3):
        assert list(got[i]) == list(orc.cm31_mul(a2[2 * i:2 * i + 2], b2[2 * i:2 * i + 2]))
    got, failv = ver.cm31_inv(a2)
    got = got.reshape(n, 2)
    for i in range(0, n, 3):
        w, f = orc.cm31_inv(a2[2 * i:2 * i + 2])
        assert list(got[i]) == list(w) and bool(failv[i]) == f


def test_points_and_folds(ver, orc, kat):
    rng = np.random.default_rng(13)
    idx = rng.integers(0, 2**32, 700, dtype=np.uint64).astype(np.uint32)
    idx[:4] = [0, 1, 1389, 2**31]
    got = ver.circle_point(idx).reshape(-1, 2)
    for i in range(len(idx)):
        assert list(got[i]) == list(orc.circle_point_index_to_m31_point(int(idx[i])))
    n = 600
    P = 2147483647
    for log, circle in ((13, True), (4, True), (12, False), (5, False), (3, False)):
        pos = (rng.integers(0, 1 << log, n, dtype=np.uint64) & ~np.uint64(1)).astype(np.uint32)
        f_p = (rng.integers(0, P, 4 * n, dtype=np.uint64)).astype(np.uint32)
        f_n = (rng.integers(0, 2**32, 4 * n, dtype=np.uint64)).astype(np.uint32)  # witness side may be non-canonical
        al = (rng.integers(0, P, 4 * n, dtype=np.uint64)).astype(np.uint32)
        got, failv = (ver.circle_fold if circle else ver.line_fold)(pos, f_p, f_n, al, log)
        got = got.reshape(n, 4)
        ofold = orc.circle_fold if circle else orc.line_fold
        for i in range(0, n, 5):
            w, f = ofold(int(pos[i]), f_p[4 * i:4 * i + 4], f_n[4 * i:4 * i + 4], log, al[4 * i:4 * i + 4])
            assert list(got[i]) == list(w) and bool(failv[i]) == f, (log, circle, i)
    # the two in-source fold KATs (fri/folding.simf:45-65)
    lets, asserts = kat("stwo-verifier/src/fri/folding.simf", "test_circle_fold")
    from conftest import flat
    got, _ = ver.circle_fold(np.array([lets["query"]], dtype=np.uint32), np.array(flat(lets["f_p"]), dtype=np.uint32),
                             np.array(flat(lets["f_neg_p"]), dtype=np.uint32), np.array(flat(lets["fold_alpha"]), dtype=np.uint32), lets["log_size_ex"])
    assert list(got) == flat(asserts[0]["rhs"])


def test_sha256_pair_and_merkle(ver, orc, kat):
    from oracle.oracle import u256_words, words_u256

    rng = np.random.default_rng(14)
    n = 777
    l = rng.integers(0, 2**32, 8 * n, dtype=np.uint64).astype(np.uint32)
    r = rng.integers(0, 2**32, 8 * n, dtype=np.uint64).astype(np.uint32)
    got = ver.sha256_pair(l, r).reshape(n, 8)
    import hashlib

    for i in range(n):
        msg = b"".join(int(x).to_bytes(4, "big") for x in np.concatenate([l[8 * i:8 * i + 8], r[8 * i:8 * i + 8]]))
        assert b"".join(int(x).to_bytes(4, "big") for x in got[i]) == hashlib.sha256(msg).digest()
    # Merkle paths: random siblings, depths around the reference's sweep ends plus 0 and 31 (List<u256, 32> bounds)
    for depth in (0, 1, 4, 13, 16, 31):
        m = 200
        leaf = rng.integers(0, 2**32, 8 * m, dtype=np.uint64).astype(np.uint32)
        sib = rng.integers(0, 2**32, 8 * m * depth, dtype=np.uint64).astype(np.uint32)
        auth = (rng.integers(0, 1 << depth, m, dtype=np.uint64) + (1 << depth)).astype(np.uint32)
        auth[0] = 0xFFFFFFFF  # path that cannot end at 1
        root, path, ok = ver.merkle_root_from_path(leaf, auth, sib, depth)
        exp_roots = np.zeros_like(root)
        for i in range(m):
            sibs = [words_u256(sib[(i * depth + k) * 8:(i * depth + k + 1) * 8]) for k in range(depth)]
            _, croot, fpath = orc.merkle_verify_32(words_u256(leaf[8 * i:8 * i + 8]), int(auth[i]), sibs, 0)
            exp_roots[8 * i:8 * i + 8] = u256_words(croot)
            assert fpath == path[i]
        assert (root == exp_roots).all(), depth
        # with the right expected roots every path with final path == 1 verifies; flipping one root bit fails it
        bad = exp_roots.copy()
        bad[8 * 7] ^= 1
        _, _, ok = ver.merkle_root_from_path(leaf, auth, sib, depth, expected_root=bad)
        bits = np.unpackbits(ok.view(np.uint8), bitorder="little")[:m].astype(bool)
        expect = path == 1
        expect[7] = False
        assert (bits == expect).all(), depth
    # the reference's own vector (merkle.simf:58-80)
    lets, _ = kat("stwo-verifier/src/merkle.simf", "test_decommitment")
    leaf = u256_words(orc.sha256_32(2915689030))
    sib = np.concatenate([u256_words(s) for s in lets["proof"]])
    root, path, ok = ver.merkle_root_from_path(leaf, np.array([lets["leaf_id"] + 8192], dtype=np.uint32), sib, 13, expected_root=u256_words(lets["root"]))
    assert words_u256(root) == lets["root"] and path[0] == 1 and ok[0] == 1


def test_channel_jets(ver, orc, kat):
    from oracle.oracle import u256_words

    rng = np.random.default_rng(15)
    n = 300
    st = rng.integers(0, 2**32, 9 * n, dtype=np.uint64).astype(np.uint32)
    st[8::9] = rng.integers(0, 5, n)
    val = rng.integers(0, 2**32, 8 * n, dtype=np.uint64).astype(np.uint32)
    got = ver.channel_mix_u256(st.copy(), val).reshape(n, 9)
    for i in range(0, n, 3):
        want = st[9 * i:9 * i + 9].copy()
        orc.lib.oracle_channel_mix_u256(want.ctypes.data_as(C.POINTER(C.c_uint32)), val[8 * i:8 * i + 8].ctypes.data_as(C.POINTER(C.c_uint32)))
        assert list(got[i]) == list(want)
    hl = rng.integers(0, 2**32, 2 * n, dtype=np.uint64).astype(np.uint32)
    got = ver.channel_mix_u64(st.copy(), hl).reshape(n, 9)
    for i in range(0, n, 3):
        want = orc.channel_mix_u64(st[9 * i:9 * i + 9], (int(hl[2 * i]) << 32) | int(hl[2 * i + 1]))
        assert list(got[i]) == list(want)
    st2 = st.copy()
    out, failv = ver.channel_draw_qm31(st2)
    for i in range(0, n, 3):
        s, v, f = orc.channel_draw_qm31(st[9 * i:9 * i + 9])
        assert list(out[4 * i:4 * i + 4]) == list(v) and list(st2[9 * i:9 * i + 9]) == list(s) and bool(failv[i]) == f
    st3 = st.copy()
    q = ver.channel_draw_queries(st3, 13, 16).reshape(n, 16)
    for i in range(0, n, 3):
        s, v = orc.channel_draw_queries(st[9 * i:9 * i + 9], 13, 16)
        assert list(q[i]) == list(v) and list(st3[9 * i:9 * i + 9]) == list(s)
    # in-source KATs: channel.simf:176-186, fri/queries.simf:47-61
    lets, _ = kat("stwo-verifier/src/channel.simf", "test_channel_draw_qm31")
    from conftest import flat
    s = orc.state(*lets["state"])
    out, _ = ver.channel_draw_qm31(s)
    assert list(out) == flat(lets["first_random_felt"])
    out, _ = ver.channel_draw_qm31(s)
    assert list(out) == flat(lets["second_random_felt"])
    lets, asserts = kat("stwo-verifier/src/fri/queries.simf", "test_channel_draw_queries_8")
    q = ver.channel_draw_queries(orc.state(*lets["state"]), 6, 8)
    assert list(q) == [a["rhs"] for a in asserts]


def test_stark101_field_jets(ver, orc):
    rng = np.random.default_rng(16)
    n = 1200
    a = rng.integers(0, 2**32, n, dtype=np.uint64).astype(np.uint32)
    b = rng.integers(0, 2**32, n, dtype=np.uint64).astype(np.uint32)
    b[:5] = [0, 1, 3221225473, 2, 0xFFFFFFFF]
    got = ver.s101_mul_mod(a, b)
    assert (got == np.array([orc.s101_mul_mod(int(x), int(y)) for x, y in zip(a, b)], dtype=np.uint32)).all()
    got, failv = ver.s101_div_mod(a, b)
    for i in range(n):
        w, f = orc.s101_div_mod(int(a[i]), int(b[i]))
        assert got[i] == w and bool(failv[i]) == f, i


def test_int32_probe_runs(ver):
    ops, ms = ver.int32_peak_probe()
    assert ops > 1e12 and ms > 0


def test_stwo_host_async_mode_matches_sync(S, ver, orc):
    """ssym_set_host_async: back-to-back host-buffer calls, one synchronize; same bitmaps / statuses as the synchronous call."""
    import torch

    cfg, packed = golden_stwo(S, "prod", 1)
    batch, n, _ = negatives_batch(S, cfg, packed, extra_random=690, seed=3)
    pinned = torch.from_numpy(batch.view(np.int32).copy()).pin_memory()
    host = pinned.numpy().view(np.uint32)
    acc_sync, st_sync, _ = ver.stwo_verify_batch(host, cfg, n, want_status=True)
    rows = torch.zeros((5, (n + 31) // 32), dtype=torch.int32).pin_memory().numpy().view(np.uint32)
    sts = torch.zeros((5, n), dtype=torch.int32).pin_memory().numpy().view(np.uint32)
    ver.set_host_async(True)
    try:
        for k in range(5):
            ver.stwo_verify_batch(host, cfg, n, accept_out=rows[k], status_out=sts[k])
        ver.synchronize()
    finally:
        ver.set_host_async(False)
    assert (rows == acc_sync[None, :]).all() and (sts == st_sync[None, :]).all()
    _, o_status, _ = orc.stwo_verify_batch(ocfg(cfg), batch, n)
    assert (st_sync == o_status).all()
