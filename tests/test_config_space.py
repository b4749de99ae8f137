"""Random points of the configuration space (trace / LDE sizes, query count, NUM_COLUMNS; config.simf:10-51) beyond the two presets:
CPU — the reference prover's proofs are accepted by the restated verifier under PROVER_CONSISTENT and rejected under REF_LITERAL at the
FRI layer-0 root, one flipped word is rejected; GPU — prover byte-identical to the CPU one, verifier traces byte-identical to the oracle's."""
import numpy as np
import pytest
from hypothesis import HealthCheck, given, settings
from hypothesis import strategies as st

from oracle import oracle as O

POW = 0x07FFFFFFFFFFFFFF


@st.composite
def configs(draw):
    T = draw(st.integers(2, 6))
    G = T + draw(st.integers(1, 3))
    Q = draw(st.integers(1, 16))
    C = draw(st.sampled_from([4, 8, 16]))
    return T, G, Q, C


@settings(max_examples=25, deadline=None, suppress_health_check=[HealthCheck.function_scoped_fixture, HealthCheck.too_slow])
@given(c=configs(), seed=st.integers(0, 2**64 - 1), word=st.integers(0, 2**31), bit=st.integers(0, 31))
def test_oracle_prover_and_verifier_agree_on_random_configurations(orc, c, seed, word, bit):
    T, G, Q, C = c
    cfg = O.StwoConfig(T, G, Q, T - 1, O.MODE_PROVER_CONSISTENT, C, POW)
    lo = orc.stwo_layout(cfg)
    pk = orc.stwo_prove_batch(cfg, [seed])
    bad = pk[0].copy()
    w = word % lo.stride_words
    bad[w] ^= np.uint32(1 << bit)
    _, status, _ = orc.stwo_verify_batch(cfg, np.concatenate([pk[0], bad]), 2)
    assert status[0] == 0, (c, hex(status[0]))
    lit = O.StwoConfig(T, G, Q, T - 1, O.MODE_REF_LITERAL, C, POW)
    _, st_lit, tr = orc.stwo_verify_batch(lit, pk[0], 1, want_trace=True)
    assert st_lit[0] & (1 << 7) and (tr[0].first_fail >> 16) == 7
    # the flipped word is rejected unless it sits in the alignment padding of the record (zero in every packed proof, read by nothing)
    if pk[0][w] != 0 or _is_payload(lo, cfg, w):
        assert status[1] != 0, (c, w)


def _is_payload(lo, cfg, w):
    Q, L, G, C = cfg.n_queries, cfg.n_fri_layers, cfg.lde_log, cfg.n_columns
    spans = [(0, lo.off_pow_nonce + 2), (lo.off_qvals, lo.off_qvals + Q * (C + 16)), (lo.off_trace_sib, lo.off_fri_wit + (L + 1) * Q * 4)]
    spans += [(lo.off_fri_sib[l], lo.off_fri_sib[l] + Q * (G - 1 - l) * 8) for l in range(L + 1)]
    inside = any(a <= w < b for a, b in spans)
    return inside and not (lo.off_commit <= w < lo.off_commit + 8)  # the constant tree's root is mixed into the channel: flipping it moves every draw... and is still payload


@pytest.mark.gpu
def test_gpu_matches_oracle_on_random_configurations():
    import stark_symphony_b200 as S

    S.load()
    orc = O.Oracle()
    ver = S.Verifier(0)
    rng = np.random.default_rng(5)
    for _ in range(8):
        T = int(rng.integers(2, 8))
        G = min(13, T + int(rng.integers(1, 5)))
        Q = int(rng.integers(1, 17))
        C = int(rng.choice([4, 8, 16]))
        cfg = S.stwo_config("prod", S.MODE_PROVER_CONSISTENT, n_columns=C)
        cfg.trace_log, cfg.lde_log, cfg.n_queries, cfg.n_fri_layers = T, G, Q, T - 1
        ocfg = O.StwoConfig(T, G, Q, T - 1, cfg.mode, C, cfg.pow_target)
        lo = S.stwo_layout(cfg)
        seeds = rng.integers(0, 2**63, size=12, dtype=np.int64).astype(np.uint64)
        gpu = ver.stwo_prove_batch(seeds, cfg)
        ref = orc.stwo_prove_batch(ocfg, seeds, threads=4)
        assert (gpu == ref).all(), (T, G, Q, C)
        batch = gpu.copy()
        for k in range(4, 12):
            batch[k, rng.integers(0, lo.stride_words)] ^= np.uint32(1 << rng.integers(0, 32))
        for sharing in (0, 2):
            ver.set_merkle_sharing(sharing)
            for mode in (S.MODE_PROVER_CONSISTENT, S.MODE_REF_LITERAL):
                cfg.mode = mode
                ocfg.mode = mode
                accept, status, traces = ver.stwo_verify_batch(batch.ravel(), cfg, 12, want_status=True, want_trace=True)
                _, o_status, o_traces = orc.stwo_verify_batch(ocfg, batch.ravel(), 12, want_trace=True)
                assert (status == o_status).all(), (T, G, Q, C, mode, sharing)
                for i in range(12):
                    assert bytes(memoryview(traces[i]).cast("B")) == bytes(memoryview(o_traces[i]).cast("B")), (T, G, Q, C, mode, sharing, i)
                if mode == S.MODE_PROVER_CONSISTENT:
                    assert (status[:4] == 0).all()
        ver.set_merkle_sharing(1)
    ver.close()
