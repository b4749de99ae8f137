#!/usr/bin/env python3
"""tests/golden/stark101_multiquery.json: the reference's stark101 proof decommitted at MORE query positions.

The reference's prover (stark101/scripts/fibsquare/prover.py:94-171) draws ONE query index after the commitments and decommits there; the
verifier program does the same (stark101/src/verifier.simf:32).  One query of a rate-1/8 code is ~3 bits of soundness; a multi-query proof
repeats the query phase: the k-th query index is the (k+1)-th `receive_random_int(0, 8191)` on the channel after the commitments
(channel.py:73-85 = channel_draw_32, channel.simf:102-105: every draw re-hashes the state; the decommitted values are sent without `mix`,
prover.py:87-91, so they do not move it), and each query is decommitted exactly as prover.py:141-168 does for the first.

This script does not restate the prover: it RUNS the reference's `prove()` unmodified, with `MerkleTree` and `Channel` replaced by recording
subclasses, and then asks the recorded trees (MerkleTree.get_authentication_path, merkle.py:38-54) for the further positions.  Query 0 of the
output must be the reference's own proof (asserted against its return value and against tests/golden/stark101_proof.json).

Usage: python tests/golden/make_s101_multiquery.py [n_queries=4]      (needs /root/reference; ~25 s)
"""
import contextlib
import io
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("SSYM_REFERENCE", "/root/reference")
DOMAIN_EX_MULT = 8


def main():
    n_queries = int(sys.argv[1]) if len(sys.argv) > 1 else 4
    sys.path.insert(0, os.path.join(REF, "stark101", "scripts"))
    import fibsquare.prover as P

    trees, channels = [], []

    class RecTree(P.MerkleTree):
        def __init__(self, data):
            super().__init__(data)
            trees.append(self)

    class RecChannel(P.Channel):
        def __init__(self, *a, **kw):
            super().__init__(*a, **kw)
            channels.append(self)

    P.MerkleTree, P.Channel = RecTree, RecChannel
    sink = io.StringIO()
    with contextlib.redirect_stdout(sink):  # the channel prints every message
        _, res0 = P.prove()
    channel = channels[0]
    p_mt, fri_mts = trees[0], trees[1:]  # trace tree; composition polynomial = FRI layer 0, then one tree per further layer (the last is never sent)
    n_dec = len(res0["fri_layers"])
    assert len(fri_mts) == n_dec + 1
    for i in range(n_dec):
        assert int.from_bytes(fri_mts[i].root, "big") == res0["fri_layers"][i][0]

    def ints(path):
        return [int.from_bytes(x, "big") for x in path[::-1]]  # leaf -> root, as prover.py:146-148

    def decommit(idx):
        if idx + 2 * DOMAIN_EX_MULT >= len(p_mt.data):
            raise SystemExit(f"query {idx}: f(g^2 x) lies outside the evaluation list; the reference's prover cannot decommit it (prover.py:143)")
        out = {"p_mt_root": res0["p_mt_root"], "fri_last_layer": res0["fri_last_layer"]}
        out["evals"] = [[p_mt.data[j].val, ints(p_mt.get_authentication_path(j))] for j in (idx, idx + DOMAIN_EX_MULT, idx + 2 * DOMAIN_EX_MULT)]
        out["fri_layers"] = []
        for i in range(n_dec):  # prover.py:152-166
            layer, mt = fri_mts[i].data, fri_mts[i]
            length = len(layer)
            a, b = idx % length, (idx + length // 2) % length
            out["fri_layers"].append([res0["fri_layers"][i][0], res0["fri_layers"][i][1], layer[a].val, ints(mt.get_authentication_path(a)),
                                      layer[b].val, ints(mt.get_authentication_path(b))])
        return out

    # the channel has drawn query 0 already (prover.py:138); its state now is what the next draw reduces
    idxs, proofs = [], []
    for k in range(n_queries):
        if k == 0:
            # recover idx 0 from the proof itself: the position whose authentication path the reference sent
            idx = next(j for j in range(len(p_mt.data) - 2 * DOMAIN_EX_MULT)
                       if p_mt.data[j].val == res0["evals"][0][0] and ints(p_mt.get_authentication_path(j)) == res0["evals"][0][1])
        else:
            with contextlib.redirect_stdout(sink):
                idx = channel.receive_random_int(0, len(p_mt.data) - 1, f"query #{k}")
        idxs.append(idx)
        proofs.append(decommit(idx))
    # every decommitment of every query against the reference's own checker (merkle.py:73-86 verify_decommitment)
    from fibsquare.field import FieldElement
    from fibsquare.merkle import verify_decommitment

    def as_path(ints_):
        return [int(x).to_bytes(32, "big") for x in ints_[::-1]]  # back to root -> leaf, the order get_authentication_path returns

    for idx, pr in zip(idxs, proofs):
        root = int(pr["p_mt_root"]).to_bytes(32, "big")
        for j, (val, path) in enumerate(pr["evals"]):
            assert verify_decommitment(idx + j * DOMAIN_EX_MULT, FieldElement(val), as_path(path), root)
        for i, (lroot, _beta, cpa, pa, cpb, pb) in enumerate(pr["fri_layers"]):
            length = len(fri_mts[i].data)
            lr = int(lroot).to_bytes(32, "big")
            assert verify_decommitment(idx % length, FieldElement(cpa), as_path(pa), lr)
            assert verify_decommitment((idx + length // 2) % length, FieldElement(cpb), as_path(pb), lr)
    assert proofs[0] == res0, "query 0 must be the reference's own proof"
    golden = json.load(open(os.path.join(HERE, "stark101_proof.json")))
    assert proofs[0] == golden, "query 0 must equal tests/golden/stark101_proof.json"
    out = {"n_queries": n_queries, "idx": idxs, "queries": proofs,
           "note": "queries[k] has the shape of the reference's proof.json (stark101/scripts/fibsquare/__main__.py); queries[0] IS that proof"}
    with open(os.path.join(HERE, "stark101_multiquery.json"), "w") as f:
        json.dump(out, f, separators=(",", ":"))
    print(f"wrote stark101_multiquery.json: query indices {idxs}")


if __name__ == "__main__":
    main()
