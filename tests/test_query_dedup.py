"""SSYM_MODE_QUERY_DEDUP (include/ssym.h): the queries are sorted and de-duplicated before use — the step fri/queries.simf:41 says the
reference leaves out.  The reference defines no vector for it, so the rule is pinned three ways: the oracle verifier and the CPU reference
prover implement it independently of the CUDA code (accept / reject behaviour on honest and corrupted proofs, unused slots ignored), the
GPU verifier's whole trace equals the oracle's, and the GPU prover's proofs equal the CPU prover's byte for byte."""
import numpy as np
import pytest

from oracle import oracle as O

DEDUP = O.MODE_QUERY_DEDUP


def small_cfg(sem, G=5, Q=16, T=3, C=4):
    """16 queries into a domain of 2^G positions: duplicates are certain for G = 5 (and still frequent for G = 6)."""
    return O.StwoConfig(T, G, Q, T - 1, sem | DEDUP, C, 0x07FFFFFFFFFFFFFF)


def test_oracle_dedup_semantics(orc):
    cfg = small_cfg(O.MODE_PROVER_CONSISTENT)
    lo = orc.stwo_layout(cfg)
    proofs = orc.stwo_prove_batch(cfg, np.arange(40, 52, dtype=np.uint64), threads=4)
    accept, status, traces = orc.stwo_verify_batch(cfg, proofs.ravel(), len(proofs), want_trace=True)
    assert (status == 0).all()
    plain = O.StwoConfig(cfg.trace_log, cfg.lde_log, cfg.n_queries, cfg.n_fri_layers, O.MODE_PROVER_CONSISTENT, 4, cfg.pow_target)
    _, _, plain_tr = orc.stwo_verify_batch(plain, proofs.ravel(), len(proofs), want_trace=True)
    saw_dup = False
    for i, t in enumerate(traces):
        U = t.n_queries_used
        q = list(t.queries)
        drawn = list(plain_tr[i].queries)[: cfg.n_queries]  # the same transcript without the flag: the raw draws
        assert q[:U] == sorted(set(drawn)) and q[U:] == [0] * (16 - U)
        saw_dup |= U < cfg.n_queries
        # slots >= U are zero-filled by the prover and ignored by the verifier: noise there changes nothing
        if U < cfg.n_queries:
            rec = proofs[i].copy()
            G, Q = cfg.lde_log, cfg.n_queries
            rec[lo.off_qvals + 20 * U: lo.off_qvals + 20 * Q] = 0xDEADBEEF
            rec[lo.off_trace_sib + U * G * 8: lo.off_trace_sib + Q * G * 8] = 7
            rec[lo.off_fri_wit + U * 4: lo.off_fri_wit + Q * 4] = 9
            assert not proofs[i][lo.off_qvals + 20 * U: lo.off_qvals + 20 * Q].any()
            _, st2, _ = orc.stwo_verify_batch(cfg, rec, 1)
            assert st2[0] == 0
        # a used slot is still checked
        rec = proofs[i].copy()
        rec[lo.off_trace_sib + (U - 1) * cfg.lde_log * 8 + 3] ^= 1
        _, st3, tr3 = orc.stwo_verify_batch(cfg, rec, 1, want_trace=True)
        assert st3[0] & (1 << 4) and tr3[0].mask_trace == 1 << (U - 1)
    assert saw_dup
    # without the flag the same records are (rightly) rejected: their slots follow the sorted order, not the drawn one
    _, st_plain, _ = orc.stwo_verify_batch(plain, proofs.ravel(), len(proofs))
    assert (st_plain != 0).any()


def test_dedup_flag_changes_nothing_when_queries_are_sorted_and_distinct(orc):
    """TESTING preset: one query.  The flag only reorders / drops, so a single query verifies identically."""
    from test_oracle_fixtures import load_stwo

    packed = load_stwo("testing")
    for sem in (O.MODE_REF_LITERAL, O.MODE_PROVER_CONSISTENT):
        a0, s0, t0 = orc.stwo_verify_batch(O.make_config("testing", sem), packed, 1, want_trace=True)
        a1, s1, t1 = orc.stwo_verify_batch(O.make_config("testing", sem | DEDUP), packed, 1, want_trace=True)
        assert s0[0] == s1[0] and bytes(memoryview(t0[0]).cast("B")) == bytes(memoryview(t1[0]).cast("B"))


# ---- GPU ----------------------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def S():
    import stark_symphony_b200 as S

    S.load()
    return S


def scfg(S, c):
    return S.StwoConfig(c.trace_log, c.lde_log, c.n_queries, c.n_fri_layers, c.mode, c.n_columns, c.pow_target)


@pytest.mark.gpu
@pytest.mark.parametrize("G,Q,C_", [(5, 16, 4), (6, 16, 8), (7, 9, 4), (13, 16, 4)])
@pytest.mark.parametrize("sem", [O.MODE_REF_LITERAL, O.MODE_PROVER_CONSISTENT])
@pytest.mark.parametrize("policy", [0, 2])
def test_gpu_dedup_trace_equals_oracle(S, orc, G, Q, C_, sem, policy):
    """Honest proofs (CPU prover), one corrupted proof per class and random single-word corruptions: the whole trace — sorted queries, U,
    every per-query root and mask, the first failing assert — is the oracle's, with both Merkle schedules."""
    T = 3 if G < 13 else 9
    cfg = small_cfg(sem, G=G, Q=Q, T=T, C=C_)
    proofs = orc.stwo_prove_batch(cfg, np.arange(7, 13, dtype=np.uint64), threads=4)
    lo = orc.stwo_layout(cfg)
    rng = np.random.default_rng(G * 100 + Q)
    recs = [p for p in proofs]
    for w, d in S.witness.stwo_negative_classes(scfg(S, cfg)).values():
        recs.append(S.witness.apply_mutation(proofs[0], w, d))
    for _ in range(40):
        r = proofs[int(rng.integers(0, len(proofs)))].copy()
        r[int(rng.integers(0, lo.stride_words))] ^= np.uint32(1 << int(rng.integers(0, 32)))
        recs.append(r)
    batch = np.concatenate(recs)
    ver = S.Verifier(0)
    ver.set_merkle_sharing(policy)
    accept, status, traces = ver.stwo_verify_batch(batch, scfg(S, cfg), len(recs), want_status=True, want_trace=True)
    ver.close()
    o_accept, o_status, o_traces = orc.stwo_verify_batch(cfg, batch, len(recs), want_trace=True)
    assert (status == o_status).all() and (accept == o_accept).all()
    for i in range(len(recs)):
        assert bytes(memoryview(traces[i]).cast("B")) == bytes(memoryview(o_traces[i]).cast("B")), i
    if sem == O.MODE_PROVER_CONSISTENT:
        assert (o_status[: len(proofs)] == 0).all()
    if G <= 6:
        assert any(t.n_queries_used < Q for t in o_traces[: len(proofs)])


@pytest.mark.gpu
@pytest.mark.parametrize("G,C_", [(5, 4), (6, 16), (13, 4)])
def test_gpu_prover_dedup_equals_cpu_prover(S, orc, G, C_):
    T = 3 if G < 13 else 9
    cfg = small_cfg(O.MODE_PROVER_CONSISTENT, G=G, T=T, C=C_)
    seeds = np.arange(100, 108, dtype=np.uint64)
    ver = S.Verifier(0)
    got = ver.stwo_prove_batch(seeds, scfg(S, cfg))
    want = orc.stwo_prove_batch(cfg, seeds, threads=4)
    assert (got == want).all()
    accept, status, _ = ver.stwo_verify_batch(got.ravel(), scfg(S, cfg), len(seeds), want_status=True)
    assert (status == 0).all()
    # device-resident, pipelined, larger batch: every proof accepted, compact form round-trips and is smaller than without the flag
    import torch

    n = 300
    d_seeds = torch.arange(0, n, dtype=torch.int64, device="cuda")
    d_proofs = ver.stwo_prove_batch(d_seeds, scfg(S, cfg))
    ver.set_pipeline_depth(4)
    accs = [ver.stwo_verify_batch(d_proofs.view(-1), scfg(S, cfg), n)[0] for _ in range(6)]
    ver.synchronize()
    ver.set_pipeline_depth(1)
    bits = np.unpackbits(accs[-1].cpu().numpy().view(np.uint8), bitorder="little")[:n]
    assert bits.all() and all((a == accs[0]).all() for a in accs)
    host = d_proofs.cpu().numpy().view(np.uint32)
    blob, offsets = S.witness.compact_stwo(host, scfg(S, cfg))
    expanded, flags = ver.stwo_compact_expand(blob, offsets, scfg(S, cfg), want_flags=True)
    assert (expanded == host).all() and not flags.any()
    ver.close()
