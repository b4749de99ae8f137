"""Compact transport form of packed Stwo proofs (include/ssym.h): the reference's witness carries one full authentication path per query
(merkle.simf:39-44, evals/verify.simf:20-36, fri/layers.simf:18-24); the compact record keeps every distinct sibling of a tree once.
The host packer is pinned here by an independent numpy expander (CPU); the GPU expander and ssym_stwo_verify_compact_batch are compared
with the packed path and the oracle (GPU)."""
import os

import numpy as np
import pytest

from conftest import GOLDEN
from oracle import oracle as O

MAGIC = 0x32435353


@pytest.fixture(scope="module")
def S():
    import stark_symphony_b200 as S

    S.load()
    return S


def ocfg(cfg):
    return O.StwoConfig(cfg.trace_log, cfg.lde_log, cfg.n_queries, cfg.n_fri_layers, cfg.mode, cfg.n_columns, cfg.pow_target)


def numpy_expand(S, cfg, blob, offsets):
    """Independent restatement of the record format (include/ssym.h): compact -> packed."""
    lo = S.stwo_layout(cfg)
    Q, L, G = cfg.n_queries, cfg.n_fri_layers, cfg.lde_log
    depths = [G, G] + [G - 1 - l for l in range(L + 1)]
    n = len(offsets) - 1
    out = np.zeros((n, lo.stride_words), dtype=np.uint32)
    idx_bytes = 1 if Q * G <= 256 else 2
    fixed, wit = lo.off_trace_sib, lo.off_fri_sib[0] - lo.off_fri_wit
    slots = Q * sum(depths)
    bitmap_words = ((slots + 31) // 32 + 7) // 8 * 8
    for i in range(n):
        rec = blob[int(offsets[i]):int(offsets[i + 1])]
        assert rec[0] == len(rec) and rec[2] == MAGIC and len(rec) % 8 == 0
        D, R = int(rec[1]), int(rec[3])
        assert D + R == slots
        off_wit = 8 + fixed
        off_bitmap = off_wit + wit
        off_refs = off_bitmap + bitmap_words
        refs_words = (R * idx_bytes + 31) // 32 * 8
        off_tab = off_refs + refs_words
        assert len(rec) == off_tab + 8 * D
        out[i, :fixed] = rec[8:8 + fixed]
        out[i, lo.off_fri_wit:lo.off_fri_wit + wit] = rec[off_wit:off_bitmap]
        bits = np.unpackbits(rec[off_bitmap:off_refs].view(np.uint8), bitorder="little")
        assert int(bits[:slots].sum()) == D
        refs = rec[off_refs:off_tab].view(np.uint8 if idx_bytes == 1 else np.uint16)
        tab = rec[off_tab:].reshape(D, 8)
        s = new = ref = 0
        for t, d in enumerate(depths):
            tree_first = new
            for k in range(Q * d):
                if bits[s]:
                    e = new
                    new += 1
                else:
                    e = tree_first + int(refs[ref])
                    ref += 1
                    assert e < new
                dst = lo.off_trace_sib + 8 * s if s < 2 * Q * G else lo.off_fri_sib[0] + 8 * (s - 2 * Q * G)
                out[i, dst:dst + 8] = tab[e]
                s += 1
    return out


def _records(S, orc, cfg, rng):
    """honest proofs, corrupted ones, and pure noise (no two siblings equal: the worst case)."""
    pk = orc.stwo_prove_batch(ocfg(cfg), list(range(90, 96)), threads=4)
    lo = S.stwo_layout(cfg)
    recs = [pk[i] for i in range(6)]
    recs += [S.witness.apply_mutation(pk[0], w, d) for (w, d) in S.witness.stwo_negative_classes(cfg).values()]
    noise = rng.integers(0, 2**32, size=lo.stride_words, dtype=np.uint64).astype(np.uint32)
    recs.append(noise)
    return np.stack(recs)


@pytest.mark.parametrize("preset,nc", [("testing", 4), ("prod", 4), ("prod", 16)])
def test_compact_pack_is_lossless_and_smaller(S, orc, preset, nc):
    cfg = S.stwo_config(preset, S.MODE_PROVER_CONSISTENT, n_columns=nc)
    lo = S.stwo_layout(cfg)
    recs = _records(S, orc, cfg, np.random.default_rng(1))
    blob, offsets = S.witness.compact_stwo(recs, cfg)
    assert offsets[0] == 0 and offsets[-1] == blob.size and (np.diff(offsets.astype(np.int64)) % 8 == 0).all()
    assert (numpy_expand(S, cfg, blob, offsets) == recs).all()
    sizes = np.diff(offsets.astype(np.int64)) * 4
    bound = S.load().ssym_stwo_compact_bound(__import__("ctypes").byref(cfg), 1) * 4
    assert sizes[-1] == bound  # noise: nothing to share
    if preset == "prod":
        assert (sizes[:6] < 0.80 * lo.stride_words * 4).all(), sizes[:6]  # honest proofs: >= 20 % fewer bytes on the link
    # the shipped fixture
    if nc == 4:
        text = open(os.path.join(GOLDEN, f"stwo_proof_{preset}.wit")).read()
        packed, bad = S.witness.pack_stwo_wits([text], cfg)
        b2, o2 = S.witness.compact_stwo(packed, cfg)
        assert not bad[0] and (numpy_expand(S, cfg, b2, o2)[0] == packed).all()


def test_compact_pack_errors(S, orc):
    import ctypes as C

    cfg = S.stwo_config("testing", 1)
    pk = orc.stwo_prove_batch(ocfg(cfg), [1, 2])
    small = np.zeros(40, dtype=np.uint32)
    with pytest.raises(S.SsymError):
        S.witness.compact_stwo(pk, cfg, out=small)
    blob, offsets = S.witness.compact_stwo(pk[:0], cfg)
    assert blob.size == 0 and list(offsets) == [0]
    assert S.load().ssym_stwo_compact_bound(C.byref(S.stwo_config("prod", 0, n_columns=5)), 3) == 0


# ---- GPU ----------------------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def ver(S):
    return S.Verifier(0)


@pytest.mark.gpu
@pytest.mark.parametrize("preset,nc", [("testing", 4), ("prod", 4), ("prod", 8)])
def test_gpu_expand_and_verify_compact(S, ver, orc, preset, nc):
    import torch

    cfg = S.stwo_config(preset, S.MODE_PROVER_CONSISTENT, n_columns=nc)
    recs = _records(S, orc, cfg, np.random.default_rng(2))
    n = len(recs)
    blob, offsets = S.witness.compact_stwo(recs, cfg)
    # host buffers
    packed, flags = ver.stwo_compact_expand(blob, offsets, cfg, want_flags=True)
    assert (packed == recs).all() and not flags.any()
    # device buffers
    d_blob, d_off = torch.from_numpy(blob.view(np.int32)).cuda(), torch.from_numpy(offsets.view(np.int64)).cuda()
    d_packed, d_flags = ver.stwo_compact_expand(d_blob, d_off, cfg, want_flags=True)
    ver.synchronize()
    assert (d_packed.cpu().numpy().view(np.uint32) == recs).all() and not d_flags.cpu().numpy().any()
    for mode in (S.MODE_PROVER_CONSISTENT, S.MODE_REF_LITERAL):
        cfg.mode = mode
        ref_accept, ref_status, _ = ver.stwo_verify_batch(recs.ravel(), cfg, n, want_status=True)
        _, o_status, _ = orc.stwo_verify_batch(ocfg(cfg), recs.ravel(), n)
        assert (ref_status == o_status).all()
        accept, status = ver.stwo_verify_compact_batch(blob, offsets, cfg, want_status=True)
        assert (status == ref_status).all() and (accept == ref_accept).all()
        d_accept, d_status = ver.stwo_verify_compact_batch(d_blob, d_off, cfg, want_status=True)
        ver.synchronize()
        assert (d_status.cpu().numpy().view(np.uint32) == ref_status).all() and (d_accept.cpu().numpy().view(np.uint32) == ref_accept).all()
        if mode == S.MODE_PROVER_CONSISTENT:
            assert (status[:6] == 0).all() and (status[6:] != 0).all()


@pytest.mark.gpu
def test_gpu_malformed_compact_records_are_rejected(S, ver, orc):
    cfg = S.stwo_config("prod", S.MODE_PROVER_CONSISTENT)
    lo = S.stwo_layout(cfg)
    pk = orc.stwo_prove_batch(ocfg(cfg), [5, 6, 7, 8, 9, 10], threads=4)
    blob, offsets = S.witness.compact_stwo(pk, cfg)
    blob = blob.copy()
    o = [int(x) for x in offsets]
    blob[o[1] + 2] ^= 1                  # magic
    blob[o[2] + 1] += 1                  # D does not match the record length
    fixed, wit = lo.off_trace_sib, lo.off_fri_sib[0] - lo.off_fri_wit
    off_bitmap = 8 + fixed + wit
    blob[o[3] + off_bitmap] ^= 2         # one more / one fewer "new digest" bit than D says
    slots = cfg.n_queries * (2 * cfg.lde_log + sum(cfg.lde_log - 1 - l for l in range(cfg.n_fri_layers + 1)))
    off_refs = off_bitmap + ((slots + 31) // 32 + 7) // 8 * 8
    refs = blob[o[4] + off_refs:].view(np.uint8)
    refs[0] = 255                        # a back reference that does not point to an earlier digest of its tree
    packed, flags = ver.stwo_compact_expand(blob, offsets, cfg, want_flags=True)
    assert list(flags) == [0, 1, 1, 1, 1, 0]
    assert (packed[[0, 5]] == pk[[0, 5]]).all() and not packed[1:5].any()
    accept, status = ver.stwo_verify_compact_batch(blob, offsets, cfg, want_status=True)
    assert status[0] == 0 and status[5] == 0 and (status[1:5] >> 31 == 1).all() and int(accept[0]) & 63 == 0b100001
    bad_off = offsets.copy()
    bad_off[2] = bad_off[1] - 8
    with pytest.raises(S.SsymError):
        ver.stwo_verify_compact_batch(blob, bad_off, cfg)


@pytest.mark.gpu
def test_gpu_compact_host_path_chunks_and_async(S, ver, orc):
    """3000 records (several double-buffered chunks), every 7th corrupted; synchronous and enqueue-only calls give the packed path's bitmap."""
    cfg = S.stwo_config("prod", S.MODE_PROVER_CONSISTENT)
    n = 3000
    proofs = ver.stwo_prove_batch(np.arange(n, dtype=np.uint64) + 40000, cfg)
    classes = list(S.witness.stwo_negative_classes(cfg).values())
    for j, row in enumerate(range(3, n, 7)):
        w, d = classes[j % len(classes)]
        proofs[row, w] = np.uint32((int(proofs[row, w]) + d) & 0xFFFFFFFF)
    blob, offsets = S.witness.compact_stwo(proofs, cfg)
    ref_accept, ref_status, _ = ver.stwo_verify_batch(proofs.ravel(), cfg, n, want_status=True)
    accept, status = ver.stwo_verify_compact_batch(blob, offsets, cfg, want_status=True)
    assert (status == ref_status).all() and (accept == ref_accept).all()
    expect_bad = np.zeros(n, dtype=bool)
    expect_bad[3::7] = True
    assert ((status != 0) == expect_bad).all()
    ver.set_host_async(True)
    outs = [ver.stwo_verify_compact_batch(blob, offsets, cfg, want_status=True) for _ in range(3)]
    ver.synchronize()
    ver.set_host_async(False)
    for a, st in outs:
        assert (st == ref_status).all() and (a == ref_accept).all()


def test_compact_pack_fuzz_with_many_duplicates(S):
    """Records whose siblings come from a pool of a few digests (duplicates inside a tree, across trees, runs of equal slots, all-equal
    trees): the packer's tables stay per tree, indices fit their byte, and the independent expander gives every record back."""
    rng = np.random.default_rng(7)
    for preset, nc in (("testing", 4), ("prod", 4), ("prod", 16)):
        cfg = S.stwo_config(preset, 0, n_columns=nc)
        lo = S.stwo_layout(cfg)
        recs = rng.integers(0, 2**32, size=(12, lo.stride_words), dtype=np.uint64).astype(np.uint32)
        first, last = lo.off_trace_sib, lo.stride_words
        for i, pool_size in enumerate((1, 1, 2, 3, 5, 8, 17, 40, 100, 255, 256, 300)):
            pool = rng.integers(0, 2**32, size=(pool_size, 8), dtype=np.uint64).astype(np.uint32)
            if i == 1:
                pool[:] = 0
            for off in list(range(lo.off_trace_sib, lo.off_fri_wit, 8)) + list(range(lo.off_fri_sib[0], last, 8)):
                recs[i, off:off + 8] = pool[rng.integers(0, pool_size)]
        blob, offsets = S.witness.compact_stwo(recs, cfg)
        assert (numpy_expand(S, cfg, blob, offsets) == recs).all()
        sizes = np.diff(offsets.astype(np.int64))
        assert (sizes[:-1] <= sizes[-1]).all() and sizes[0] < sizes[-1]


def test_compact_blob_shards_by_proof_index(S, orc):
    """sharding.shard_compact: every rank's slice expands to exactly its shard_range of the packed batch."""
    from importlib import import_module

    sh = import_module("stark_symphony_b200.sharding")
    cfg = S.stwo_config("testing", 1)
    n = 70
    pk = orc.stwo_prove_batch(ocfg(cfg), list(range(n)), threads=4)
    blob, offsets = S.witness.compact_stwo(pk, cfg)
    for world in (1, 2, 3, 8):
        seen = 0
        for rank in range(world):
            b, e = sh.shard_range(n, rank, world)
            sb, so = sh.shard_compact(blob, offsets, rank, world)
            assert so[0] == 0 and len(so) == e - b + 1 and so[-1] == sb.size
            if e > b:
                assert (numpy_expand(S, cfg, sb, so) == pk[b:e]).all()
            seen += e - b
        assert seen == n


@pytest.mark.gpu
def test_gpu_expand_survives_corrupted_compact_records(S, ver, orc):
    """Random corruptions of the header, the bitmap, the back references and the record lengths of valid compact records: the GPU either flags the
    record (zeros out) or expands it to exactly what the independent expander gives; nothing else (run under compute-sanitizer memcheck too)."""
    cfg = S.stwo_config("prod", S.MODE_PROVER_CONSISTENT)
    lo = S.stwo_layout(cfg)
    pk = orc.stwo_prove_batch(ocfg(cfg), [77], threads=1)
    n = 400
    blob1, off1 = S.witness.compact_stwo(pk, cfg)
    w = int(off1[1])
    blob = np.tile(blob1, n)
    offsets = (np.arange(n + 1, dtype=np.uint64) * np.uint64(w))
    rng = np.random.default_rng(11)
    fixed, wit = lo.off_trace_sib, lo.off_fri_sib[0] - lo.off_fri_wit
    off_bitmap = 8 + fixed + wit
    slots = cfg.n_queries * (2 * cfg.lde_log + sum(cfg.lde_log - 1 - l for l in range(cfg.n_fri_layers + 1)))
    off_refs = off_bitmap + ((slots + 31) // 32 + 7) // 8 * 8
    R = int(blob1[3])
    for i in range(1, n):
        base = i * w
        kind = i % 5
        if kind == 0:
            blob[base + int(rng.integers(0, 4))] = np.uint32(rng.integers(0, 2**32))            # header word
        elif kind == 1:
            blob[base + off_bitmap + int(rng.integers(0, (slots + 31) // 32))] ^= np.uint32(1 << int(rng.integers(0, 32)))  # a bitmap bit
        elif kind == 2:
            blob[base + off_refs:base + off_refs + (R + 3) // 4].view(np.uint8)[int(rng.integers(0, R))] = np.uint8(rng.integers(0, 256))  # a back reference
        elif kind == 3:
            blob[base + 1], blob[base + 3] = blob[base + 3], blob[base + 1]                        # D and R swapped
        else:
            blob[base + off_refs + (R + 31) // 32 * 8 + int(rng.integers(0, 64))] ^= np.uint32(1)  # a digest word: still a valid record
    packed, flags = ver.stwo_compact_expand(blob, offsets, cfg, want_flags=True)
    assert flags[0] == 0 and (packed[0] == pk[0]).all()
    seen = set()
    for i in range(n):
        rec = blob[i * w:(i + 1) * w]
        try:
            want = numpy_expand(S, cfg, rec, np.array([0, w], dtype=np.uint64))[0]
        except (AssertionError, IndexError, ValueError):
            want = None
        if flags[i]:
            assert not packed[i].any()
            assert want is None, i  # the GPU refuses nothing the format allows
        else:
            assert want is not None and (packed[i] == want).all(), i
        seen.add(int(flags[i]))
    assert seen == {0, 1}
