"""The cost model (include/ssym.h "Cost model", csrc/cost.cpp) against the oracle's per-thread counters: every field of
ssym_cost_t, for the reference's fixtures in both semantics, for other configurations and column counts on proofs of the CPU reference
prover, and for random query vectors (the m31_mul / m31_add counts depend on the popcount of the point indices the queries map to)."""
import ctypes as C

import numpy as np
import pytest

from conftest import GOLDEN
from oracle import oracle as O
from test_oracle_fixtures import load_stwo


@pytest.fixture(scope="module")
def S():
    import stark_symphony_b200 as S

    S.load()
    return S


def oracle_counts(orc, cfg, packed):
    out = (C.c_uint64 * 14)()
    orc.lib.oracle_cost_reset()
    _, status, traces = orc.stwo_verify_batch(cfg, packed, 1, want_trace=True)
    orc.lib.oracle_cost_counts(out)
    return [int(x) for x in out], traces[0], int(status[0])


def model_counts(S, cfg, queries, retries, used=None):
    from stark_symphony_b200 import _lib

    scfg = S.StwoConfig(cfg.trace_log, cfg.lde_log, cfg.n_queries, cfg.n_fri_layers, cfg.mode, cfg.n_columns, cfg.pow_target)
    cost = _lib.Cost()
    q = np.ascontiguousarray(queries, dtype=np.uint32)
    S._lib.check(S.load().ssym_stwo_cost(C.byref(scfg), C.c_void_p(q.ctypes.data), cfg.n_queries if used is None else used, retries, C.byref(cost)))
    return [cost.as_dict()[k] for k in _lib.COST_FIELDS]


@pytest.mark.parametrize("preset", ["testing", "prod"])
@pytest.mark.parametrize("mode", [O.MODE_REF_LITERAL, O.MODE_PROVER_CONSISTENT])
def test_cost_model_equals_oracle_counters_on_fixtures(S, orc, preset, mode):
    cfg = O.make_config(preset, mode)
    want, tr, _ = oracle_counts(orc, cfg, load_stwo(preset))
    got = model_counts(S, cfg, list(tr.queries)[: cfg.n_queries], tr.draw_retries)
    assert got == want
    if preset == "prod" and mode == O.MODE_REF_LITERAL:
        d = dict(zip(S._lib.COST_FIELDS, got))
        assert (d["sha_compressions"], d["m31_mul"], d["m31_add"], d["m31_inv"], d["sha_finalize"], d["sha_bytes"]) == (3806, 65486, 55220, 162, 2061, 117168)


@pytest.mark.parametrize("T,G,Q,C_", [(3, 5, 2, 4), (4, 7, 5, 4), (5, 9, 16, 8), (6, 10, 9, 16), (9, 13, 16, 16)])
def test_cost_model_on_other_configurations(S, orc, T, G, Q, C_):
    """Proofs of the CPU reference prover at other sizes / widths (16 columns: the trace leaf is a two-compression message)."""
    for mode in (O.MODE_REF_LITERAL, O.MODE_PROVER_CONSISTENT):
        cfg = O.StwoConfig(T, G, Q, T - 1, mode, C_, 0x07FFFFFFFFFFFFFF)
        proofs = orc.stwo_prove_batch(cfg, np.array([5, 6], dtype=np.uint64), threads=2)
        for rec in proofs:
            want, tr, _ = oracle_counts(orc, cfg, rec)
            assert model_counts(S, cfg, list(tr.queries)[:Q], tr.draw_retries) == want


def test_cost_depends_on_queries_exactly_as_the_program_does(S, orc):
    """Corrupting the nonce changes the drawn queries and nothing else: the model follows the oracle for every query vector."""
    cfg = O.make_config("prod", O.MODE_REF_LITERAL)
    lo = orc.stwo_layout(cfg)
    base = load_stwo("prod")
    seen = set()
    for k in range(12):
        rec = base.copy()
        rec[lo.off_pow_nonce + 1] += np.uint32(k)
        want, tr, _ = oracle_counts(orc, cfg, rec)
        got = model_counts(S, cfg, list(tr.queries)[:16], tr.draw_retries)
        assert got == want
        seen.add(got[7])
    assert len(seen) > 3  # the m31_mul count really moves with the queries


def test_cost_usage_errors(S):
    cost = S._lib.Cost()
    cfg = S.stwo_config("prod", 0)
    q = np.zeros(16, dtype=np.uint32)
    assert S.load().ssym_stwo_cost(None, C.c_void_p(q.ctypes.data), 16, 0, C.byref(cost)) == S.ERR_USAGE
    assert S.load().ssym_stwo_cost(C.byref(cfg), C.c_void_p(q.ctypes.data), 15, 0, C.byref(cost)) == S.ERR_USAGE  # fewer used queries only under QUERY_DEDUP
    cfg.n_queries = 99
    assert S.load().ssym_stwo_cost(C.byref(cfg), C.c_void_p(q.ctypes.data), 99, 0, C.byref(cost)) == S.ERR_USAGE


def test_cost_model_with_deduplicated_queries(S, orc):
    """SSYM_MODE_QUERY_DEDUP: only the U distinct queries are verified; the model with n_queries_used = U equals the oracle's counters."""
    for G, Q in ((5, 16), (6, 16), (9, 16)):
        for sem in (O.MODE_REF_LITERAL, O.MODE_PROVER_CONSISTENT):
            cfg = O.StwoConfig(3, G, Q, 2, sem | O.MODE_QUERY_DEDUP, 4, 0x07FFFFFFFFFFFFFF)
            for rec in orc.stwo_prove_batch(cfg, np.array([21, 22, 23], dtype=np.uint64)):
                want, tr, _ = oracle_counts(orc, cfg, rec)
                assert tr.n_queries_used <= Q
                assert model_counts(S, cfg, list(tr.queries)[:Q], tr.draw_retries, tr.n_queries_used) == want
