"""The CPU reference prover (oracle/stwo_prover_ref.c) against the reference VERIFIER restatement.

The reference ships no prover for its wide-Fibonacci AIR (SURVEY.md section 7 hard part b), so the prover is pinned by the
verifier: every proof it emits must pass stwo-verifier/src/verifier.simf:32-58 in PROVER_CONSISTENT mode (the semantics
under which the reference's own fixtures verify), and must be rejected exactly where the fixtures are in REF_LITERAL mode."""
import numpy as np
import pytest

from oracle import oracle as O

P = 2**31 - 1
ST_FRI_MERKLE0 = 1 << 7
ST_FINAL_LOG = 1 << 17
ST_LAST_QUERY = 1 << 18


@pytest.mark.parametrize("preset,seeds", [("testing", list(range(24)) + [2**63 + 5]), ("prod", [0, 1, 0xDEADBEEF])])
def test_proofs_accept_in_prover_consistent_mode(orc, preset, seeds):
    cfg = O.make_config(preset, O.MODE_PROVER_CONSISTENT)
    pk = orc.stwo_prove_batch(cfg, seeds, threads=4)
    accept, status, traces = orc.stwo_verify_batch(cfg, pk.ravel(), len(seeds), want_trace=True)
    assert (status == 0).all(), [hex(s) for s in status]
    assert len({bytes(r) for r in pk}) == len(seeds)  # distinct seeds -> distinct proofs
    # the fixtures' behaviour under the literal semantics (SURVEY finding 3): reject at FRI layer 0 Merkle + F2 (+ F3)
    lit = O.make_config(preset, O.MODE_REF_LITERAL)
    _, st_lit, tr_lit = orc.stwo_verify_batch(lit, pk.ravel(), len(seeds), want_trace=True)
    for s, t in zip(st_lit, tr_lit):
        assert s & ST_FRI_MERKLE0 and s & ST_FINAL_LOG
        assert (t.first_fail >> 16) == 7  # first failing assert in program order = FRI layer-0 Merkle root
    # everything before fri_answers is mode independent
    for a, b in zip(traces, tr_lit):
        assert bytes(a.digest_pow) == bytes(b.digest_pow) and list(a.queries) == list(b.queries)


def test_prover_is_deterministic_and_seeded(orc):
    cfg = O.make_config("testing", O.MODE_PROVER_CONSISTENT)
    a = orc.stwo_prove_batch(cfg, [7, 8])
    b = orc.stwo_prove_batch(cfg, [7, 8], threads=2)
    assert (a == b).all() and not (a[0] == a[1]).all()


def test_trace_rows_satisfy_the_air(orc):
    import ctypes as C

    row = (C.c_uint32 * 4)()
    for seed in (0, 5, 2**40):
        for r in (0, 1, 511):
            orc.lib.oracle_stwo_trace_row(C.c_uint64(seed), C.c_uint32(r), row)
            c0, c1, c2, c3 = (int(x) for x in row)
            assert c0 == 1 and c1 < P and c2 == (c0 * c0 + c1 * c1) % P and c3 == (c1 * c1 + c2 * c2) % P


@pytest.mark.parametrize("preset", ["testing", "prod"])
def test_every_section_of_a_proof_is_load_bearing(orc, preset):
    """Flip one word in each section of the packed record: the verifier must reject (soundness of the pin)."""
    cfg = O.make_config(preset, O.MODE_PROVER_CONSISTENT)
    lo = orc.stwo_layout(cfg)
    pk = orc.stwo_prove_batch(cfg, [3])[0]
    offs = [lo.off_commit + 8, lo.off_commit + 16, lo.off_oods_trace + 4, lo.off_oods_cp + 5, lo.off_fri_first_root, lo.off_last_coeff,
            lo.off_qvals + 1, lo.off_qvals + 7, lo.off_trace_sib + 3, lo.off_cp_sib + 9, lo.off_fri_wit, lo.off_fri_sib[0] + 2]
    if cfg.n_fri_layers:
        offs += [lo.off_fri_inner_root, lo.off_fri_wit + 4 * cfg.n_queries, lo.off_fri_sib[1]]
    recs = []
    for o in offs:
        r = pk.copy()
        r[o] ^= 1
        recs.append(r)
    _, status, _ = orc.stwo_verify_batch(cfg, np.concatenate(recs), len(recs))
    assert (status != 0).all(), [hex(s) for s in status]
    # the nonce: nonce - 1 must fail the PoW (the prover returns the smallest passing nonce) unless nonce == 0
    nonce = (int(pk[lo.off_pow_nonce]) << 32) | int(pk[lo.off_pow_nonce + 1])
    if nonce:
        r = pk.copy()
        r[lo.off_pow_nonce + 1] = (nonce - 1) & 0xFFFFFFFF
        _, st, _ = orc.stwo_verify_batch(cfg, r, 1)
        assert st[0] & (1 << 3)
