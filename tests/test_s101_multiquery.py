"""Multi-query stark101 (include/ssym.h ssym_stark101_verify_multi_batch): a Q-query proof is Q records of the
reference's witness shape, record k verified under the (k+1)-th query draw.

The fixture (tests/golden/stark101_multiquery.json, made by tests/golden/make_s101_multiquery.py) is the reference's prover run unmodified, its
recorded Merkle trees asked for three more positions; the positions were drawn by the reference's own Channel (channel.py:73-85).  So the golden
query indices pin the ordinal semantics of the oracle, and an independent hashlib restatement of the channel pins its commitment state."""
import hashlib
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN
from oracle import oracle as O
from oracle import witparse as W

ST_GROUP = 1 << 9
ST_SHAPE = 1 << 31


@pytest.fixture(scope="module")
def S():
    import stark_symphony_b200 as S

    return S


@pytest.fixture(scope="module")
def ver(S):
    v = S.Verifier(0)
    yield v
    v.close()


def golden():
    return json.load(open(os.path.join(GOLDEN, "stark101_multiquery.json")))


def wit_text(q):
    """The reference's generate_wit.py formatting (stark101/scripts/generate_wit.py:7-30), via the package's restatement of it."""
    import stark_symphony_b200 as S

    return json.dumps(S.witness.stark101_wit_from_proof_json(q))


def records(g, order=None):
    """Records of one proof, packed by the ORACLE-side reader; record k carries ordinal k (order: which golden query goes into slot k)."""
    recs = []
    for k, src in enumerate(order if order is not None else range(g["n_queries"])):
        rec = W.pack_stark101(W.load_wit(wit_text(g["queries"][src]))).copy()
        rec[6] = k
        recs.append(rec)
    return recs


def batch(recs):
    offsets = np.zeros(len(recs) + 1, dtype=np.uint64)
    offsets[1:] = np.cumsum([len(r) for r in recs])
    return np.concatenate(recs), offsets


def hashlib_channel(q0, n_draws):
    """verifier.simf:27-32 with hashlib: state after the commitments, then n_draws query indices (channel.simf:102-105)."""
    be = lambda v, n: int(v).to_bytes(n, "big")
    state = hashlib.sha256(be(q0["p_mt_root"], 32)).digest()
    draw = lambda st, m: (int.from_bytes(st, "big") % m, hashlib.sha256(st).digest())
    for _ in range(3):
        _, state = draw(state, 3221225473)
    for layer in q0["fri_layers"]:
        state = hashlib.sha256(state + be(layer[0], 32)).digest()
        beta, state = draw(state, 3221225473)
        assert beta == layer[1]
    state = hashlib.sha256(state + be(q0["fri_last_layer"], 4)).digest()
    commit, idx = state, []
    for _ in range(n_draws):
        v, state = draw(state, 8192)
        idx.append(v)
    return commit, idx


def cases(g):
    """(name, records of ONE proof, expected accept, slots expected to carry SSYM_S101_ST_GROUP)"""
    Q = g["n_queries"]
    good = records(g)
    out = [("honest", good, True, [])]
    swapped = records(g, order=[1, 0] + list(range(2, Q)))  # right ordinals, decommitments of other positions: Merkle / FRI checks fail
    out.append(("decommitments swapped", swapped, False, []))
    wrong_ord = [r.copy() for r in good]
    wrong_ord[1][6], wrong_ord[2][6] = 2, 1  # each record verifies under ITS ordinal's position (so: fails), and sits in the wrong slot
    out.append(("ordinals swapped", wrong_ord, False, [1, 2]))
    dup = [good[0].copy() for _ in range(Q)]  # the one-query proof presented Q times
    out.append(("query 0 repeated", dup, False, list(range(1, Q))))
    other = [r.copy() for r in good]
    other[3][5] += 1  # another last layer: another transcript
    out.append(("record 3 of another transcript", other, False, [3]))
    first = [r.copy() for r in good]
    first[0][8] ^= 1  # the FIRST record differs: every other record disagrees with it
    out.append(("record 0 of another transcript", first, False, list(range(1, Q))))
    shape = [r.copy() for r in good]
    shape[2][6] = 256
    out.append(("ordinal out of range", shape, False, []))
    shape0 = [r.copy() for r in good]
    shape0[0][1] = 99
    out.append(("record 0 malformed", shape0, False, list(range(1, Q))))
    return out


def test_multiquery_oracle_pinned_by_reference_channel(orc):
    g = golden()
    Q = g["n_queries"]
    commit, idx = hashlib_channel(g["queries"][0], Q)
    assert idx == g["idx"] == [6160, 5842, 3963, 3462]  # drawn by the reference's Channel in make_s101_multiquery.py
    blob, offsets = batch(records(g))
    accept, status, traces = orc.s101_verify_multi_batch(blob, offsets, Q)
    assert int(accept[0]) & 1 == 1 and not status.any()
    for k in range(Q):
        assert traces[k].idx == g["idx"][k] and traces[k].query_ordinal == k
        assert bytes(np.array(traces[k].commit_state, dtype=">u4").tobytes()) == commit
    # one query = the reference's program: the single-proof call and the multi call with Q = 1 agree, ordinal 0
    a1, s1, t1 = orc.s101_verify_batch(blob[: int(offsets[1])], offsets[:2], want_trace=True)
    am, sm, tm = orc.s101_verify_multi_batch(blob[: int(offsets[1])], offsets[:2], 1)
    assert int(a1[0]) & 1 == 1 and int(am[0]) & 1 == 1 and s1[0] == 0 and sm[0] == 0 and bytes(t1[0]) == bytes(tm[0])


def test_multiquery_oracle_negatives(orc):
    g = golden()
    Q = g["n_queries"]
    for name, recs, ok, group_slots in cases(g):
        blob, offsets = batch(recs)
        accept, status, _ = orc.s101_verify_multi_batch(blob, offsets, Q)
        assert bool(int(accept[0]) & 1) == ok, name
        assert [k for k in range(Q) if status[k] & ST_GROUP] == group_slots, (name, [hex(s) for s in status])
        assert ok == (not status.any()), name


def test_package_packer_equals_oracle_reader(S):
    g = golden()
    blob, offsets = S.witness.pack_stark101_multiquery(g["queries"])
    o_blob, o_offsets = batch(records(g))
    assert (blob == o_blob).all() and (offsets == o_offsets).all()


@pytest.mark.gpu
@pytest.mark.parametrize("device_resident", [False, True])
def test_multiquery_gpu_matches_oracle(S, ver, orc, device_resident):
    g = golden()
    Q = g["n_queries"]
    rng = np.random.default_rng(11)
    all_cases = cases(g)
    recs, expect = [], []
    for name, rs, ok, _ in all_cases:
        recs += rs
        expect.append(ok)
    good = all_cases[0][1]
    for _ in range(120):  # random corruptions of one record of an honest proof, and honest proofs in between
        rs = [r.copy() for r in good]
        if rng.integers(0, 4):
            k = int(rng.integers(0, Q))
            w = int(rng.integers(1, len(rs[k])))
            if w in (7, 19):
                w = 5
            rs[k][w] = (int(rs[k][w]) + int(rng.integers(1, 2**32))) & 0xFFFFFFFF
            if w < 5:  # a length word: keep the record's slot self-consistent is the CALLER's job for host buffers (checked below on device only)
                rs[k][w] = good[k][w]
                rs[k][5] ^= 4
        recs += rs
    blob, offsets = batch(recs)
    n_proofs = len(recs) // Q
    o_accept, o_status, o_traces = orc.s101_verify_multi_batch(blob, offsets, Q)
    if device_resident:
        import torch

        d_blob = torch.from_numpy(blob.view(np.int32)).cuda()
        d_off = torch.from_numpy(offsets.view(np.int64)).cuda()
        accept, status, traces = ver.stark101_verify_multi_batch(d_blob, d_off, Q, want_trace=True)
        ver.synchronize()
        accept, status = accept.cpu().numpy().view(np.uint32), status.cpu().numpy().view(np.uint32)
        raw = traces.cpu().numpy().tobytes()
        sz = len(raw) // len(recs)
        g_traces = [raw[i * sz:(i + 1) * sz] for i in range(len(recs))]
    else:
        accept, status, traces = ver.stark101_verify_multi_batch(blob, offsets, Q, want_trace=True)
        g_traces = [bytes(traces[i]) for i in range(len(recs))]
    assert (status == o_status).all(), [(i, hex(a), hex(b)) for i, (a, b) in enumerate(zip(status, o_status)) if a != b][:5]
    assert (accept == o_accept).all()
    bits = np.unpackbits(accept.view(np.uint8), bitorder="little")[:n_proofs].astype(bool)
    assert list(bits[: len(expect)]) == expect
    for i in range(len(recs)):
        assert g_traces[i] == bytes(o_traces[i]), i


@pytest.mark.gpu
def test_multiquery_replicated_device(S, ver, orc):
    """4096 proofs x 4 queries in HBM, some proofs broken in one record: one accept bit per proof."""
    import torch

    g = golden()
    Q = g["n_queries"]
    good = records(g)
    one, _ = batch(good)
    n = 4096
    blob = np.tile(one, n)
    lens = np.array([len(r) for r in good] * n, dtype=np.uint64)
    offsets = np.zeros(n * Q + 1, dtype=np.uint64)
    offsets[1:] = np.cumsum(lens)
    bad = {0: 3, 31: 0, 32: 1, 1000: 2, n - 1: 3}
    for i, k in bad.items():
        blob[int(offsets[i * Q + k]) + 16] ^= 1  # f(x) of query k
    d_blob = torch.from_numpy(blob.view(np.int32)).cuda()
    d_off = torch.from_numpy(offsets.view(np.int64)).cuda()
    accept, status, _ = ver.stark101_verify_multi_batch(d_blob, d_off, Q)
    ver.synchronize()
    bits = np.unpackbits(accept.cpu().numpy().view(np.uint8), bitorder="little")[:n].astype(bool)
    expect = np.ones(n, dtype=bool)
    expect[list(bad)] = False
    assert (bits == expect).all()
    st = status.cpu().numpy().view(np.uint32).reshape(n, Q)
    for i, k in bad.items():
        assert st[i, k] != 0 and not np.delete(st[i], k).any()


@pytest.mark.gpu
def test_cli_multiquery(S, tmp_path):
    """verify-batch --program stark101 --queries 4: every four consecutive witnesses (the reference's `.wit` shape) are one proof."""
    import subprocess

    g = golden()
    Q = g["n_queries"]
    paths = []
    for k in range(Q):
        p = tmp_path / f"q{k}.wit"
        p.write_text(wit_text(g["queries"][k]))
        paths.append(str(p))
    cli = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "stark-symphony_b200", "bin", "verify-batch")
    ok = subprocess.run([cli, "--program", "stark101", "--queries", str(Q), "--witness", *paths, "--replicate", "3"], capture_output=True, text=True)
    assert ok.returncode == 0 and ok.stdout.count("accept") == 3 * Q, (ok.stdout, ok.stderr)
    swapped = [paths[1], paths[0]] + paths[2:]
    bad = subprocess.run([cli, "--program", "stark101", "--queries", str(Q), "--witness", *swapped], capture_output=True, text=True)
    assert bad.returncode == 1 and bad.stdout.count("reject") == Q and "Error: Failed to run program" in bad.stderr, (bad.stdout, bad.stderr)
    # without --queries each witness is a one-query proof under the FIRST draw: only the reference's own proof (query 0) is accepted
    single = subprocess.run([cli, "--program", "stark101", "--host-pack", "--witness", *paths], capture_output=True, text=True)
    assert single.returncode == 1 and single.stdout.splitlines()[0].startswith("accept") and single.stdout.count("reject") == Q - 1
