"""The reference's in-source known-answer tests for stark101 (stark101/src/*.simf), re-run against the
CPU oracle, plus a cross-check of the oracle against the reference's *Python* prover-side primitives
through the committed golden proof (tests/golden/stark101_proof.json = `python -m fibsquare`)."""
import hashlib
import json
import os

import numpy as np

from conftest import GOLDEN

S = "stark101/src/"
P = 3221225473


def rhs(asserts, lhs):
    return [a["rhs"] for a in asserts if a["lhs"] == lhs]


# ---- field.simf:96-158 ---------------------------------------------------------------------
def test_field_kats(orc, kat):
    f = S + "field.simf"
    for name, op in (("test_add_mod", orc.s101_add_mod), ("test_sub_mod", orc.s101_sub_mod), ("test_mul_mod", orc.s101_mul_mod),
                     ("test_mul_mod_2", orc.s101_mul_mod), ("test_exp_mod", orc.s101_exp_mod), ("test_exp_mod_2", orc.s101_exp_mod)):
        lets, asserts = kat(f, name)
        assert op(lets["a"], lets["b"]) == rhs(asserts, "c")[0], name
    lets, _ = kat(f, "test_div_mod")
    c, fail = orc.s101_div_mod(lets["a"], lets["b"])
    assert not fail and orc.s101_mul_mod(c, lets["b"]) == lets["a"]
    lets, asserts = kat(f, "test_div_mod_2")
    assert orc.s101_div_mod(lets["a"], lets["b"]) == (rhs(asserts, "c")[0], False)


def test_field_literal_semantics(orc):
    """field.simf:14-66 on arbitrary u32: division by 0 / by a non-unit trips the gcd assert (:46)."""
    assert orc.s101_div_mod(5, 0)[1] is True
    assert orc.s101_div_mod(5, P)[1] is True          # p = 0 mod p, not bitwise zero
    assert orc.s101_div_mod(5, 0xFFFFFFFF)[1] is True  # non-canonical divisor: first quotient is 0, r becomes b
    assert orc.s101_sub_mod(0, 0) == 0 and orc.s101_sub_mod(P, 1) == P - 1
    assert orc.s101_add_mod(0xFFFFFFFF, 0xFFFFFFFF) == (2 * 0xFFFFFFFF) % P
    rng = np.random.default_rng(7)
    for a, b in rng.integers(1, P, (200, 2)):
        c, fail = orc.s101_div_mod(int(a), int(b))
        assert not fail and c == int(a) * pow(int(b), -1, P) % P


# ---- channel.simf:107-112, sha256.simf:32-42, merkle.simf:45-79 -------------------------------
def test_channel_draw_32(orc, kat):
    lets, asserts = kat(S + "channel.simf", "test_channel_draw_32")
    st, v = orc.s101_channel_draw_32(lets["state"], 8193)  # channel.simf:109
    assert v == rhs(asserts, "value")[0] and st == rhs(asserts, "state")[0]


def test_sha_merkle_kats(orc, kat):
    lets, asserts = kat(S + "sha256.simf", "test_sha256")
    assert orc.sha256(lets["input"]) == rhs(asserts, "result")[0]
    lets, asserts = kat(S + "sha256.simf", "test_sha256_32")
    assert orc.sha256_32(lets["input"]) == rhs(asserts, "result")[0]
    lets, _ = kat(S + "merkle.simf", "test_merkle")
    assert orc.s101_merkle_verify_32(orc.sha256(0), 4, lets["proof"], lets["root"])
    lets, _ = kat(S + "merkle.simf", "test_decommitment")
    assert orc.s101_merkle_verify_32(orc.sha256_32(2915689030), lets["leaf_id"] + 8192, lets["proof"], lets["root"])
    # stark101's merkle_verify_32 has no `path == 1` assert (merkle.simf:39-43): only the root decides
    assert not orc.s101_merkle_verify_32(orc.sha256_32(2915689030), lets["leaf_id"] + 8192, lets["proof"][:-1], lets["root"])


# ---- air.simf:103-139 ----------------------------------------------------------------------------
def test_air_kats(orc, kat):
    f = S + "air.simf"
    _, asserts = kat(f, "test_fibsquare_calc_x")
    assert orc.s101_calc_x(365) == rhs(asserts, "x")[0]
    lets, asserts = kat(f, "test_fibsquare_eval_p0")
    assert orc.s101_eval_p0(lets["x"], lets["f_x"]) == rhs(asserts, "p0")[0]
    lets, asserts = kat(f, "test_fibsquare_eval_cp")
    assert orc.s101_eval_cp(lets["x"], *lets["(a0, a1, a2)"], *lets["(f_x, f_gx, f_ggx)"]) == rhs(asserts, "cp")[0]
    lets, asserts = kat(f, "test_fibsquare_read_coefficients")
    st, out = lets["state"], []
    for _ in range(3):
        st, v = orc.s101_channel_draw_32(st, P)
        out.append(v)
    assert out == [rhs(asserts, f"alpha{i}")[0] for i in range(3)]


# ---- fri.simf:93-203 -------------------------------------------------------------------------------
def test_fri_kats(orc, kat):
    f = S + "fri.simf"
    _, asserts = kat(f, "test_fri_eval_cp_next")
    assert orc.s101_fri_eval_cp_next(587367660, 786239131, 1944025132, 593409582) == rhs(asserts, "cp_next")[0]  # fri.simf:94
    _, asserts = kat(f, "test_compute_auth_path")
    got = [list(orc.s101_compute_auth_path(365, n)) for n in (8192, 4096, 32, 16)]  # fri.simf:99-110
    assert [g[0] for g in got] == rhs(asserts, "cpa_path") and [g[1] for g in got] == rhs(asserts, "cpb_path")
    # modulo_32(x, 0) = x and divide_32(x, 0) = 0 (jet semantics) once the domain has shrunk to 0
    assert list(orc.s101_compute_auth_path(365, 0)) == [365, 365]

    lets, asserts = kat(f, "test_fri_verify_layer")
    root, beta, cpa, pa, cpb, pb = lets["layer"]
    idx, x, cp_ev, size = lets["acc"]
    assert cp_ev == cpa
    a_path, b_path = orc.s101_compute_auth_path(idx, size)
    assert orc.s101_merkle_verify_32(orc.sha256_32(cpa), int(a_path), pa, root)
    assert orc.s101_merkle_verify_32(orc.sha256_32(cpb), int(b_path), pb, root)
    assert orc.s101_fri_eval_cp_next(cpa, cpb, x, beta) == rhs(asserts, "cp_ev")[0]
    assert orc.s101_mul_mod(x, x) == rhs(asserts, "x")[0] and size // 2 == rhs(asserts, "domain_size")[0]

    lets, asserts = kat(f, "test_fri_read_commitment")
    st, ok = orc.s101_fri_read_commitment(lets["state"], lets["layer"][0], lets["layer"][1])
    assert ok and st == rhs(asserts, "state")[0]
    assert not orc.s101_fri_read_commitment(lets["state"], lets["layer"][0], lets["layer"][1] + 1)[1]


# ---- reference Python cross-check (scripts/fibsquare/{channel,merkle}.py through the golden proof) ----
def test_python_prover_consistency(orc):
    """The golden proof was produced by the reference's Python prover: its Merkle paths (merkle.py:60-87,
    sha256 of the decimal-free 4-byte big-endian leaf) and channel (channel.py:55-84) must agree with the
    oracle's restatement of merkle.simf / channel.simf."""
    proof = json.load(open(os.path.join(GOLDEN, "stark101_proof.json")))
    root = proof["p_mt_root"]
    # channel.py: state = sha256(b'' + root)
    st = int.from_bytes(hashlib.sha256(root.to_bytes(32, "big")).digest(), "big")
    assert orc.sha256(root) == st
    # channel.py:73: num = int(state) % (max - min + 1), then state = sha256(state)
    st2, v = orc.s101_channel_draw_32(st, P)
    assert v == st % P and st2 == int.from_bytes(hashlib.sha256(st.to_bytes(32, "big")).digest(), "big")
    val, sibs = proof["evals"][0]
    leaf = int.from_bytes(hashlib.sha256(val.to_bytes(4, "big")).digest(), "big")
    assert orc.sha256_32(val) == leaf
    assert orc.s101_merkle_verify_32(leaf, 6160 + 8192, sibs, root)
