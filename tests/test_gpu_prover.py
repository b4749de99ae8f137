"""GPU prover (ssym_stwo_prove_batch) against the CPU reference prover (oracle/stwo_prover_ref.c) and the verifier.
Bar: the packed proofs are byte-identical, and every one verifies (PROVER_CONSISTENT) on the GPU and on the oracle.
Needs a B200: `pytest -m gpu`."""
import numpy as np
import pytest

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


def ocfg(cfg):
    from oracle import oracle as O

    return O.StwoConfig(cfg.trace_log, cfg.lde_log, cfg.n_queries, cfg.n_fri_layers, cfg.mode, cfg.n_columns, cfg.pow_target)


def first_diff(a, b, lo):
    idx = np.flatnonzero(a != b)
    if not len(idx):
        return None
    names = ["off_commit", "off_oods_trace", "off_oods_cp", "off_fri_first_root", "off_fri_inner_root", "off_last_coeff", "off_pow_nonce", "off_qvals",
             "off_trace_sib", "off_cp_sib", "off_fri_wit"]
    sect = max(((getattr(lo, n), n) for n in names if getattr(lo, n) <= idx[0]), default=(0, "?"))
    return f"{len(idx)} words differ, first at word {idx[0]} (section {sect[1]} + {idx[0] - sect[0]}): gpu {a[idx[0]]:#x} ref {b[idx[0]]:#x}"


@pytest.mark.parametrize("preset,seeds", [("testing", list(range(40)) + [2**63 + 5, 2**64 - 1]), ("prod", [0, 1, 2, 3, 0xDEADBEEF, 2**64 - 1])])
def test_gpu_prover_matches_reference_prover(S, ver, orc, preset, seeds):
    cfg = S.stwo_config(preset, S.MODE_PROVER_CONSISTENT)
    lo = S.stwo_layout(cfg)
    gpu = ver.stwo_prove_batch(np.array(seeds, dtype=np.uint64), cfg)
    ref = orc.stwo_prove_batch(ocfg(cfg), seeds, threads=8)
    for k in range(len(seeds)):
        assert first_diff(gpu[k], ref[k], lo) is None, (preset, seeds[k], first_diff(gpu[k], ref[k], lo))
    accept, status, _ = ver.stwo_verify_batch(gpu.ravel(), cfg, len(seeds), want_status=True)
    assert (status == 0).all()


def test_gpu_prover_device_resident_batch_verifies(S, ver, orc):
    """2048 distinct proofs proven and verified without leaving HBM; negatives made by corrupting one word per class;
    a sample is cross-checked on the oracle."""
    import torch

    n = 2048
    cfg = S.stwo_config("prod", S.MODE_PROVER_CONSISTENT)
    lo = S.stwo_layout(cfg)
    seeds = torch.arange(1000, 1000 + n, dtype=torch.int64, device="cuda")
    proofs = ver.stwo_prove_batch(seeds, cfg)
    torch.cuda.synchronize()
    assert proofs.shape == (n, lo.stride_words)
    classes = list(S.witness.stwo_negative_classes(cfg).values())
    bad_rows = list(range(5, n, 97))
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
    assert len(torch.unique(proofs[:, lo.off_commit + 8:lo.off_commit + 16], dim=0)) == n  # distinct trace roots
    sample = proofs[:64].cpu().numpy().view(np.uint32)
    _, o_status, _ = orc.stwo_verify_batch(ocfg(cfg), sample.ravel(), 64)
    assert (o_status == status[:64]).all()
    # the literal semantics reject every honest proof exactly like the reference fixtures (SURVEY finding 3)
    lit = S.stwo_config("prod", S.MODE_REF_LITERAL)
    _, st_lit, _ = ver.stwo_verify_batch(proofs.view(-1), lit, n, want_status=True)
    ver.synchronize()
    assert (st_lit.cpu().numpy().view(np.uint32) & (1 << 7)).all()


def test_gpu_prover_rejects_unsupported_config(S, ver):
    cfg = S.stwo_config("prod", S.MODE_PROVER_CONSISTENT)
    cfg.n_fri_layers = 5
    with pytest.raises(S.SsymError):
        ver.stwo_prove_batch(np.array([1], dtype=np.uint64), cfg)


def test_config3_full_size_properties(S, ver):
    """BASELINE config 3 at its full size (2^16 distinct proofs, 3.6 GB, proven where they are verified): size-independent properties.
    Every honest proof is accepted; every corrupted one (nine classes, 1/12 of the batch each) is rejected and
    the honest ones stay accepted; verification is idempotent; ragged shards give the bitmap of the whole batch."""
    import torch

    n = 1 << 16
    cfg = S.stwo_config("prod", S.MODE_PROVER_CONSISTENT)
    lo = S.stwo_layout(cfg)
    seeds = torch.arange(7_000_000, 7_000_000 + n, dtype=torch.int64, device="cuda")
    proofs = ver.stwo_prove_batch(seeds, cfg)
    accept0, status0, _ = ver.stwo_verify_batch(proofs.view(-1), cfg, n, want_status=True)
    ver.synchronize()
    assert int((status0 != 0).sum().item()) == 0 and bool((accept0 == -1).all().item())  # all 2^16 bits set
    classes = S.witness.stwo_negative_classes(cfg)
    names = list(classes)
    rows = torch.arange(n, device="cuda")
    cls_of_row = rows % 12  # each of the 9 corruption classes takes 1/12 of the batch, 3/12 stay honest
    for j, name in enumerate(names):
        word, delta = classes[name]
        sel = rows[cls_of_row == j]
        proofs[sel, word] += delta
    accept1, status1, _ = ver.stwo_verify_batch(proofs.view(-1), cfg, n, want_status=True)
    ver.synchronize()
    st = status1.cpu().numpy().view(np.uint32)
    cls = cls_of_row.cpu().numpy()
    assert ((st != 0) == (cls < len(names))).all()
    for j, name in enumerate(names):  # one status pattern per class (the corruption is at the same word of every proof)
        pat = np.unique(st[cls == j] & ~np.uint32(0))
        assert len(pat) >= 1 and (pat != 0).all(), name
    bits = np.unpackbits(accept1.cpu().numpy().view(np.uint8), bitorder="little")[:n].astype(bool)
    assert (bits == (st == 0)).all()
    # idempotence
    accept2, status2, _ = ver.stwo_verify_batch(proofs.view(-1), cfg, n, want_status=True)
    ver.synchronize()
    assert torch.equal(accept1, accept2) and torch.equal(status1, status2)
    # ragged shards (sizes not multiples of 32 except where the bitmap requires it) == the whole batch
    cuts = [0, 32 * 701, 32 * 701 + 32 * 13, n]
    for a, b in zip(cuts[:-1], cuts[1:]):
        acc_s, st_s, _ = ver.stwo_verify_batch(proofs[a:b].reshape(-1), cfg, b - a, want_status=True)
        ver.synchronize()
        assert torch.equal(st_s, status1[a:b]) and torch.equal(acc_s, accept1[a // 32:(b + 31) // 32])
    # a sample of every class on the CPU oracle
    from oracle import oracle as O

    orc = O.Oracle()
    oc = O.StwoConfig(cfg.trace_log, cfg.lde_log, cfg.n_queries, cfg.n_fri_layers, cfg.mode, cfg.n_columns, cfg.pow_target)
    sample = proofs[:32].cpu().numpy().view(np.uint32)
    _, o_status, _ = orc.stwo_verify_batch(oc, sample.ravel(), 32)
    assert (o_status == st[:32]).all()


@pytest.mark.parametrize("T,G,Q", [(5, 7, 8), (4, 9, 16), (6, 8, 3), (2, 4, 5)])
def test_other_configurations_end_to_end(S, ver, orc, T, G, Q):
    """Configurations config.simf has no preset for (other trace / LDE sizes, query counts that are not powers of two): GPU prover == CPU
    reference prover, GPU verifier == oracle on honest and corrupted proofs with full traces (shared-node Merkle schedule included), and
    proof -> `.wit` text -> GPU tokeniser == the packed proof."""
    import json

    cfg = S.stwo_config("prod", S.MODE_PROVER_CONSISTENT)
    cfg.trace_log, cfg.lde_log, cfg.n_queries, cfg.n_fri_layers = T, G, Q, T - 1
    lo = S.stwo_layout(cfg)
    seeds = np.arange(300, 340, dtype=np.uint64)
    gpu = ver.stwo_prove_batch(seeds, cfg)
    ref = orc.stwo_prove_batch(ocfg(cfg), seeds, threads=8)
    assert (gpu == ref).all()
    rng = np.random.default_rng(T * 100 + G)
    batch = gpu.copy()
    for k in range(8, 40):  # one random word corrupted in every proof from the 9th on
        batch[k, rng.integers(0, lo.stride_words)] ^= np.uint32(1 << rng.integers(0, 32))
    for mode in (S.MODE_PROVER_CONSISTENT, S.MODE_REF_LITERAL):
        cfg.mode = mode
        accept, status, traces = ver.stwo_verify_batch(batch.ravel(), cfg, len(seeds), want_status=True, want_trace=True)
        o_accept, o_status, o_traces = orc.stwo_verify_batch(ocfg(cfg), batch.ravel(), len(seeds), want_trace=True)
        assert (status == o_status).all() and (accept == o_accept).all()
        for i in range(len(seeds)):
            assert bytes(memoryview(traces[i]).cast("B")) == bytes(memoryview(o_traces[i]).cast("B")), (mode, i)
        if mode == S.MODE_PROVER_CONSISTENT:
            assert (status[:8] == 0).all()
    cfg.mode = S.MODE_PROVER_CONSISTENT
    texts = [json.dumps(S.witness.stwo_wit_from_packed(gpu[i], cfg)) for i in range(6)]
    blob, offsets = S.witness.concat_wit_texts(texts)
    packed, flags = ver.stwo_pack_wit_batch(blob, offsets, cfg)
    assert (flags == 0).all() and (packed == gpu[:6]).all()
