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


MAGIC3 = 0x33435353


def _sha(words):
    import hashlib

    return np.frombuffer(hashlib.sha256(np.asarray(words, dtype=">u4").tobytes()).digest(), dtype=">u4").astype(np.uint32)


def tree_geometry(cfg):
    """[(first sibling word of query 0, depth)] per tree in record order, as a function of the layout."""
    Q, L, G = cfg.n_queries, cfg.n_fri_layers, cfg.lde_log
    return [G, G] + [G - 1 - l for l in range(L + 1)]


def path_nodes(S, orc, cfg, rec, derive=None):
    """Independent restatement (hashlib + the oracle's trace for the queries and the FRI evaluations) of the nodes every query's Merkle path runs
    through: nodes[tree][q][k] = the path's node at level k (k = 0: the leaf / leaf pair), for k = 0 .. depth.  With `derive` (slot -> partner query or
    0xff) the siblings of derived slots are taken from the partner's node — level by level, as the record format defines them — and written into rec."""
    lo = S.stwo_layout(cfg)
    Q, L, G, C_ = cfg.n_queries, cfg.n_fri_layers, cfg.lde_log, cfg.n_columns or 4
    _, _, tr = orc.stwo_verify_batch(ocfg(cfg), rec, 1, want_trace=True)
    t = tr[0]
    U = t.n_queries_used
    queries = list(t.queries)
    depths = tree_geometry(cfg)
    sib_off = [lo.off_trace_sib, lo.off_cp_sib] + [lo.off_fri_sib[l] for l in range(L + 1)]
    nodes, slot0 = [], 0
    for tree, d in enumerate(depths):
        cur, pos = [], []
        for q in range(Q):
            if tree < 2:
                qv = rec[lo.off_qvals + (C_ + 16) * q: lo.off_qvals + (C_ + 16) * (q + 1)]
                cur.append(_sha(qv[:C_] if tree == 0 else qv[C_:]))
                pos.append(queries[q])
            else:
                l = tree - 2
                ev = np.array(t.fri_answer[q] if l == 0 else t.folded[l - 1][q], dtype=np.uint32)
                wit = rec[lo.off_fri_wit + (l * Q + q) * 4: lo.off_fri_wit + (l * Q + q) * 4 + 4]
                fq = queries[q] >> l
                e0, e1 = (ev, wit) if fq % 2 == 0 else (wit, ev)
                cur.append(_sha(np.concatenate([_sha(e0), _sha(e1)])))
                pos.append(fq >> 1)
        levels = [[c.copy() for c in cur]]
        for k in range(d):
            nxt = []
            for q in range(Q):
                at = sib_off[tree] + (q * d + k) * 8
                if derive is not None and derive[slot0 + q * d + k] != 0xFF and q < U:
                    rec[at:at + 8] = levels[k][int(derive[slot0 + q * d + k])]
                sib = rec[at:at + 8]
                pair = (levels[k][q], sib) if (pos[q] >> k) % 2 == 0 else (sib, levels[k][q])
                nxt.append(_sha(np.concatenate(pair)))
            levels.append(nxt)
        nodes.append(levels)
        slot0 += Q * d
    return nodes, U


def python_hints(S, orc, cfg, rec):
    """Independent restatement of ssym_stwo_compact_hints for one packed record."""
    Q = cfg.n_queries
    lo = S.stwo_layout(cfg)
    L = cfg.n_fri_layers
    nodes, U = path_nodes(S, orc, cfg, rec.copy())
    depths = tree_geometry(cfg)
    sib_off = [lo.off_trace_sib, lo.off_cp_sib] + [lo.off_fri_sib[l] for l in range(L + 1)]
    out = []
    for tree, d in enumerate(depths):
        for q in range(Q):
            for k in range(d):
                sib = rec[sib_off[tree] + (q * d + k) * 8: sib_off[tree] + (q * d + k) * 8 + 8]
                hit = 0xFF
                if q < U:
                    for p_ in range(U):
                        if p_ != q and (nodes[tree][k][p_] == sib).all():
                            hit = p_
                            break
                out.append(hit)
    return np.array(out, dtype=np.uint8)


def numpy_expand(S, cfg, blob, offsets, orc=None):
    """Independent restatement of the record format (include/ssym.h): compact -> packed.  Version 3 records (derived slots) need `orc`."""
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
        assert rec[0] == len(rec) and rec[2] in (MAGIC, MAGIC3) and len(rec) % 8 == 0
        v3 = rec[2] == MAGIC3
        D, R, X = int(rec[1]), int(rec[3]), int(rec[4]) if v3 else 0
        assert D + R + X == slots
        off_wit = 8 + fixed
        off_bitmap = off_wit + wit
        off_bitmap2 = off_bitmap + bitmap_words
        off_refs = off_bitmap2 + bitmap_words if v3 else off_bitmap2
        refs_words = (R * idx_bytes + X + 31) // 32 * 8
        off_tab = off_refs + refs_words
        assert len(rec) == off_tab + 8 * D
        out[i, :fixed] = rec[8:8 + fixed]
        out[i, lo.off_fri_wit:lo.off_fri_wit + wit] = rec[off_wit:off_bitmap]
        bits = np.unpackbits(rec[off_bitmap:off_bitmap2].view(np.uint8), bitorder="little")
        bits2 = np.unpackbits(rec[off_bitmap2:off_refs].view(np.uint8), bitorder="little") if v3 else np.zeros(len(bits), dtype=np.uint8)
        assert int(bits[:slots].sum()) == D and int(bits2[:slots].sum()) == X and not (bits[:slots] & bits2[:slots]).any()
        if X:
            assert orc is not None and int(rec[5]) == cfg.mode and 32 % Q == 0
        ref_bytes = rec[off_refs:off_tab].view(np.uint8)
        refs = ref_bytes[:R * idx_bytes].view(np.uint8 if idx_bytes == 1 else np.uint16)
        partners = ref_bytes[R * idx_bytes:R * idx_bytes + X]
        tab = rec[off_tab:].reshape(D, 8)
        derive = np.full(slots, 0xFF, dtype=np.uint8)
        s = new = ref = der = 0
        for t, d in enumerate(depths):
            tree_first = new
            for k in range(Q * d):
                dst = lo.off_trace_sib + 8 * s if s < 2 * Q * G else lo.off_fri_sib[0] + 8 * (s - 2 * Q * G)
                if bits2[s]:
                    derive[s] = partners[der]
                    der += 1
                    assert derive[s] < Q and derive[s] != k // d
                    s += 1
                    continue
                if bits[s]:
                    e = new
                    new += 1
                else:
                    e = tree_first + int(refs[ref])
                    ref += 1
                    assert e < new
                out[i, dst:dst + 8] = tab[e]
                s += 1
        if X:  # the derived siblings: what verify_proof computes on the partner's path from this very record
            path_nodes(S, orc, cfg, out[i], derive)
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
    slots = cfg.n_queries * sum(tree_geometry(cfg))
    assert sizes[-1] == bound - ((slots + 31) // 32 + 7) // 8 * 32  # noise: nothing to share (the bound also covers version 3's second bitmap)
    if preset == "prod":
        assert (sizes[:6] < 0.80 * lo.stride_words * 4).all(), sizes[:6]  # honest proofs: >= 20 % fewer bytes on the link
    # the shipped fixture
    if nc == 4:
        text = open(os.path.join(GOLDEN, f"stwo_proof_{preset}.wit")).read()
        packed, bad = S.witness.pack_stwo_wits([text], cfg)
        b2, o2 = S.witness.compact_stwo(packed, cfg)
        assert not bad[0] and (numpy_expand(S, cfg, b2, o2)[0] == packed).all()


def _pack_hinted(S, cfg, recs, hints):
    import ctypes as C

    lib = S.load()
    lo = S.stwo_layout(cfg)
    flat = np.ascontiguousarray(recs.ravel())
    n = flat.size // lo.stride_words
    buf = np.zeros(int(lib.ssym_stwo_compact_bound(C.byref(cfg), n)), dtype=np.uint32)
    offsets = np.zeros(n + 1, dtype=np.uint64)
    h = np.ascontiguousarray(hints, dtype=np.uint8)
    S._lib.check(lib.ssym_stwo_compact_pack_hinted(C.byref(cfg), C.c_void_p(flat.ctypes.data), C.c_void_p(h.ctypes.data), n, C.c_void_p(buf.ctypes.data), buf.size,
                                                   C.c_void_p(offsets.ctypes.data)))
    return buf[: int(offsets[n])].copy(), offsets


@pytest.mark.parametrize("preset,mode", [("testing", 1), ("prod", 1), ("prod", 0)])
def test_version3_records_leave_out_derivable_siblings(S, orc, preset, mode):
    """The host assembler (ssym_stwo_compact_pack_hinted) fed with hints from the independent Python scan: version 3 records expand — through the
    independent expander, which recomputes the partner's node with hashlib and the oracle's trace — to exactly the packed records, honest or
    corrupted, and an honest prod proof shrinks to about the size of upstream's minimal decommitment."""
    cfg = S.stwo_config(preset, mode)
    lo = S.stwo_layout(cfg)
    recs = _records(S, orc, cfg, np.random.default_rng(3))[[0, 1, 6, 7, 8, 11, 14, 15]] if preset == "prod" else _records(S, orc, cfg, np.random.default_rng(3))
    hints = np.stack([python_hints(S, orc, cfg, r) for r in recs])
    blob, offsets = _pack_hinted(S, cfg, recs, hints)
    assert (numpy_expand(S, cfg, blob, offsets, orc) == recs).all()
    v2_blob, v2_off = S.witness.compact_stwo(recs, cfg)
    sizes, v2_sizes = np.diff(offsets.astype(np.int64)) * 4, np.diff(v2_off.astype(np.int64)) * 4
    if preset == "prod":
        X = [int(blob[int(offsets[i]) + 4]) for i in range(len(recs))]
        if mode == 1:  # honest proofs (records 0, 1): every merge of two paths leaves two siblings out, in all 11 trees
            assert X[0] > 200 and X[1] > 200 and sizes[0] < 0.62 * lo.stride_words * 4 and sizes[0] < v2_sizes[0] - 6000, (X, sizes, v2_sizes)
        else:  # REF_LITERAL: the FRI evaluations are not the prover's (finding F1): only the trace and composition trees have derivable siblings
            assert 100 <= X[0] <= 180 and sizes[0] < v2_sizes[0], (X, sizes)
        assert X[-1] == 0  # noise
    # all-0xff hints give a version 3 record without derived slots; no hints give the version 2 record
    b0, o0 = _pack_hinted(S, cfg, recs[:2], np.full_like(hints[:2], 0xFF))
    assert int(b0[2]) == MAGIC3 and int(b0[4]) == 0 and (numpy_expand(S, cfg, b0, o0, orc) == recs[:2]).all()


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
@pytest.mark.parametrize("preset,nc,mode", [("testing", 4, 1), ("prod", 4, 1), ("prod", 4, 0), ("prod", 16, 1)])
def test_gpu_version3_pack_expand_verify(S, ver, orc, preset, nc, mode):
    """Version 3 records end to end on the GPU: the scan kernel's hints equal the independent Python scan, the GPU-assisted packer's records
    expand (GPU, host and device buffers) to exactly the packed records and (independent expander) too, verdicts from compact buffers equal the
    packed path's and the oracle's, and a record with derived slots presented under the other semantics is refused, not mis-expanded."""
    import torch

    cfg = S.stwo_config(preset, mode, n_columns=nc)
    recs = _records(S, orc, cfg, np.random.default_rng(4))
    n = len(recs)
    hints = ver.stwo_compact_hints(recs.ravel(), cfg, n)
    for i in (0, 1, 6, n - 1):
        assert (hints[i] == python_hints(S, orc, cfg, recs[i])).all(), i
    d_hints = ver.stwo_compact_hints(torch.from_numpy(recs.view(np.int32)).cuda().view(-1), cfg, n)
    ver.synchronize()
    assert (d_hints.cpu().numpy() == hints).all()
    blob, offsets = S.witness.compact_stwo(recs, cfg, ver=ver)
    assert int(blob[2]) == MAGIC3 and (int(blob[4]) > 0) == (cfg.n_queries > 1)
    v2_blob, _ = S.witness.compact_stwo(recs, cfg)
    assert blob.size <= v2_blob.size + n * 64
    assert (numpy_expand(S, cfg, blob[: int(offsets[3])], offsets[:4], orc) == recs[:3]).all()
    packed, flags = ver.stwo_compact_expand(blob, offsets, cfg, want_flags=True)
    assert not flags.any() and (packed == recs).all()
    d_blob, d_off = torch.from_numpy(blob.view(np.int32)).cuda(), torch.from_numpy(offsets.view(np.int64)).cuda()
    d_packed, d_flags = ver.stwo_compact_expand(d_blob, d_off, cfg, want_flags=True)
    ver.synchronize()
    assert (d_packed.cpu().numpy().view(np.uint32) == recs).all() and not d_flags.cpu().numpy().any()
    ref_accept, ref_status, _ = ver.stwo_verify_batch(recs.ravel(), cfg, n, want_status=True)
    _, o_status, _ = orc.stwo_verify_batch(ocfg(cfg), recs.ravel(), n)
    assert (ref_status == o_status).all()
    accept, status = ver.stwo_verify_compact_batch(blob, offsets, cfg, want_status=True)
    assert (status == ref_status).all() and (accept == ref_accept).all()
    d_accept, d_status = ver.stwo_verify_compact_batch(d_blob, d_off, cfg, want_status=True)
    ver.synchronize()
    assert (d_status.cpu().numpy().view(np.uint32) == ref_status).all() and (d_accept.cpu().numpy().view(np.uint32) == ref_accept).all()
    # the other semantics: derived slots are expanded under the mode the record was packed under (header word 5).  From host buffers the call sees
    # the headers and takes two passes (complete the records under their own mode, verify under the call's); the expansion API ignores cfg.mode likewise.
    # Records in DEVICE memory are expanded under the call's mode only: one with derived slots of another mode is refused, not mis-expanded.
    other = S.stwo_config(preset, 1 - mode, n_columns=nc)
    _, want_other, _ = orc.stwo_verify_batch(ocfg(other), recs.ravel(), n)
    acc_other, st_other = ver.stwo_verify_compact_batch(blob, offsets, other, want_status=True)
    assert (st_other == want_other).all()
    packed_o, flags_o = ver.stwo_compact_expand(blob, offsets, other, want_flags=True)
    assert not flags_o.any() and (packed_o == recs).all()
    _, d_st_other = ver.stwo_verify_compact_batch(d_blob, d_off, other, want_status=True)
    ver.synchronize()
    d_st_other = d_st_other.cpu().numpy().view(np.uint32)
    X = np.array([int(blob[int(offsets[i]) + 4]) for i in range(n)])
    assert ((d_st_other >> 31 == 1) == (X > 0)).all() and (d_st_other[X == 0] == want_other[X == 0]).all()


@pytest.mark.gpu
def test_gpu_version3_chunks_async_and_corruptions(S, ver, orc):
    """2100 distinct GPU-proven proofs (several staged chunks), every 7th corrupted, as version 3 records from pinned host buffers: synchronous and
    enqueue-only calls give the packed path's statuses; then random corruptions of the record's bitmaps / partner bytes / header: the GPU flags the
    record or expands it to exactly what the independent expander computes."""
    cfg = S.stwo_config("prod", S.MODE_PROVER_CONSISTENT)
    lo = S.stwo_layout(cfg)
    n = 2100
    proofs = ver.stwo_prove_batch(np.arange(n, dtype=np.uint64) + 7000, cfg)
    classes = list(S.witness.stwo_negative_classes(cfg).values())
    for j, row in enumerate(range(3, n, 7)):
        w, d = classes[j % len(classes)]
        proofs[row, w] = np.uint32((int(proofs[row, w]) + d) & 0xFFFFFFFF)
    blob, offsets = S.witness.compact_stwo(proofs, cfg, ver=ver)
    v2_blob, _ = S.witness.compact_stwo(proofs, cfg)
    assert blob.size < 0.80 * v2_blob.size  # honest proofs lose about a quarter of their version 2 bytes
    ref_accept, ref_status, _ = ver.stwo_verify_batch(proofs.ravel(), cfg, n, want_status=True)
    accept, status = ver.stwo_verify_compact_batch(blob, offsets, cfg, want_status=True)
    assert (status == ref_status).all() and (accept == ref_accept).all()
    expanded, flags = ver.stwo_compact_expand(blob, offsets, cfg, want_flags=True)
    assert (expanded == proofs).all() and not flags.any()
    ver.set_host_async(True)
    outs = [ver.stwo_verify_compact_batch(blob, offsets, cfg, want_status=True) for _ in range(3)]
    ver.synchronize()
    ver.set_host_async(False)
    for a, st in outs:
        assert (st == ref_status).all() and (a == ref_accept).all()
    # the same records (packed under prover-consistent) verified under the reference's literal semantics — what bench.py's e2e leg ships: one
    # transcript, the records' FRI chains to complete them, then the verification proper (launch_stwo_verify_cross); blocking and enqueue-only
    # calls (overlapping: the result buffers are a ring) give the statuses of the packed path under that mode, which are the oracle's
    lit = S.stwo_config("prod", S.MODE_REF_LITERAL)
    lit_accept, lit_status, _ = ver.stwo_verify_batch(proofs.ravel(), lit, n, want_status=True)
    _, o_lit, _ = orc.stwo_verify_batch(ocfg(lit), proofs[:64].ravel(), 64)
    assert (lit_status[:64] == o_lit).all() and lit_status.all()
    accept, status = ver.stwo_verify_compact_batch(blob, offsets, lit, want_status=True)
    assert (status == lit_status).all() and (accept == lit_accept).all()
    ver.set_host_async(True)
    outs = [ver.stwo_verify_compact_batch(blob, offsets, lit if k % 2 == 0 else cfg, want_status=True) for k in range(7)]
    ver.synchronize()
    ver.set_host_async(False)
    for k, (a, st) in enumerate(outs):
        want_st, want_a = (lit_status, lit_accept) if k % 2 == 0 else (ref_status, ref_accept)
        assert (st == want_st).all() and (a == want_a).all(), k
    # corruptions of one record
    rec = blob[: int(offsets[1])].copy()
    w = rec.size
    m = 120
    tiled = np.tile(rec, m)
    toff = np.arange(m + 1, dtype=np.uint64) * np.uint64(w)
    rng = np.random.default_rng(5)
    fixed, wit = lo.off_trace_sib, lo.off_fri_sib[0] - lo.off_fri_wit
    slots = cfg.n_queries * sum(tree_geometry(cfg))
    bm = ((slots + 31) // 32 + 7) // 8 * 8
    off_bitmap = 8 + fixed + wit
    off_refs = off_bitmap + 2 * bm
    R, X = int(rec[3]), int(rec[4])
    for i in range(1, m):
        base = i * w
        kind = i % 5
        if kind == 0:
            tiled[base + int(rng.integers(0, 6))] = np.uint32(rng.integers(0, 2**32))                                      # a header word
        elif kind == 1:
            tiled[base + off_bitmap + int(rng.integers(0, (slots + 31) // 32))] ^= np.uint32(1 << int(rng.integers(0, 32)))  # "new" bitmap
        elif kind == 2:
            tiled[base + off_bitmap + bm + int(rng.integers(0, (slots + 31) // 32))] ^= np.uint32(1 << int(rng.integers(0, 32)))  # "derived" bitmap
        elif kind == 3:
            tiled[base + off_refs:base + w].view(np.uint8)[R + int(rng.integers(0, X))] = np.uint8(rng.integers(0, 20))        # a partner byte
        else:
            tiled[base + off_refs + (R + X + 31) // 32 * 8 + int(rng.integers(0, 64))] ^= np.uint32(1)                         # a digest word
    packed, flags = ver.stwo_compact_expand(tiled, toff, cfg, want_flags=True)
    assert flags[0] == 0 and (packed[0] == proofs[0]).all()
    seen = set()
    for i in range(m):
        r = tiled[i * w:(i + 1) * w]
        try:
            want = numpy_expand(S, cfg, r, np.array([0, w], dtype=np.uint64), orc)[0]
        except (AssertionError, IndexError, ValueError):
            want = None
        if flags[i]:
            assert not packed[i].any() and want is None, i
        else:
            assert want is not None and (packed[i] == want).all(), i
        seen.add(int(flags[i]))
    assert seen == {0, 1}


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
