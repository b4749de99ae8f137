"""The reference's in-source known-answer tests for stwo-verifier, re-run against the CPU oracle.

Each test cites the `fn test_*` it re-states (file:line in the reference); the constants come from
tests/golden/kats.json (extracted by tests/golden/make_kats.py), call arguments that are inline in
the .simf test are written here."""
import numpy as np
import pytest

from conftest import flat

S = "stwo-verifier/src/"


def rhs(asserts, lhs):
    return [a["rhs"] for a in asserts if a["lhs"] == lhs][0]


# ---- fields/m31.simf:142-160 -------------------------------------------------------------
def test_m31_inv(orc, kat):
    lets, _ = kat(S + "fields/m31.simf", "test_m31_inv")
    inv, fail = orc.m31_inv(lets["a"])
    assert not fail and inv == orc.m31_exp(lets["a"], 2147483645)
    assert orc.m31_mul(inv, lets["a"]) == 1


def test_m31_add_sub(orc, kat):
    lets, asserts = kat(S + "fields/m31.simf", "test_m31_add")
    assert orc.m31_add(lets["a"], lets["b"]) == rhs(asserts, "c") == 0
    lets, asserts = kat(S + "fields/m31.simf", "test_m31_sub")
    assert orc.m31_sub(lets["a"], lets["b"]) == rhs(asserts, "c") == 2147483646


def test_m31_literal_semantics_on_non_canonical_inputs(orc):
    """m31.simf:17-45 on arbitrary u32: neg is unreduced, add wraps at 2^32 first."""
    P = 2147483647
    assert orc.m31_neg(0) == P
    assert orc.m31_neg(P) == 0
    assert orc.m31_neg(0xFFFFFFFF) == (P - 0xFFFFFFFF) % 2**32
    assert orc.m31_add(0xFFFFFFFF, 0xFFFFFFFF) == ((2 * 0xFFFFFFFF) % 2**32) % P
    assert orc.m31(P) == 0 and orc.m31(0xFFFFFFFF) == 1
    assert orc.m31_mul(0xFFFFFFFF, 0xFFFFFFFF) == (0xFFFFFFFF * 0xFFFFFFFF) % P
    assert orc.m31_inv(0) == (0, True)
    assert orc.m31_inv(P) == (0, False)  # bitwise non-zero: no assert, a^(p-2) of 0 mod p


# ---- fields/cm31.simf:118-160 ------------------------------------------------------------
@pytest.mark.parametrize("name,op", [("test_cm31_add", "cm31_add"), ("test_cm31_sub", "cm31_sub"), ("test_cm31_mul", "cm31_mul")])
def test_cm31_binary(orc, kat, name, op):
    lets, asserts = kat(S + "fields/cm31.simf", name)
    assert list(getattr(orc, op)(lets["a"], lets["b"])) == rhs(asserts, "c")


def test_cm31_mul_2(orc, kat):
    lets, asserts = kat(S + "fields/cm31.simf", "test_cm31_mul_2")
    assert list(orc.cm31_mul(orc.cm31_mul(lets["a"], lets["b"]), lets["c"])) == rhs(asserts, "d")


def test_cm31_div_inv(orc, kat):
    lets, asserts = kat(S + "fields/cm31.simf", "test_cm31_div")
    out, fail = orc.cm31_div(lets["a"], lets["b"])
    assert not fail and list(out) == rhs(asserts, "c")
    lets, asserts = kat(S + "fields/cm31.simf", "test_cm31_inv")
    inv, fail = orc.cm31_inv(lets["cm"])
    assert not fail and list(orc.cm31_mul(lets["cm"], inv)) == rhs(asserts, "o") == [1, 0]


# ---- fields/qm31.simf:136-180 ------------------------------------------------------------
def test_qm31_inv(orc, kat):
    lets, _ = kat(S + "fields/qm31.simf", "test_qm31_inv")
    a = flat(lets["a"])
    inv, fail = orc.qm31_inv(a)
    assert not fail and list(orc.qm31_mul(a, inv)) == [1, 0, 0, 0]


@pytest.mark.parametrize("name,op", [("test_qm31_add", "qm31_add"), ("test_qm31_sub", "qm31_sub"), ("test_qm31_mul", "qm31_mul")])
def test_qm31_binary(orc, kat, name, op):
    lets, asserts = kat(S + "fields/qm31.simf", name)
    assert list(getattr(orc, op)(flat(lets["a"]), flat(lets["b"]))) == flat(rhs(asserts, "c"))


def test_qm31_mul_m31(orc, kat):
    lets, asserts = kat(S + "fields/qm31.simf", "test_qm31_mul_m31")
    assert list(orc.qm31_mul_m31(flat(lets["a"]), lets["b"])) == flat(rhs(asserts, "c"))


def test_qm31_mul_cm31(orc, kat):
    lets, _ = kat(S + "fields/qm31.simf", "test_qm31_mul_cm31")
    assert list(orc.qm31_mul_cm31(flat(lets["a"]), lets["b"])) == list(orc.qm31_mul(flat(lets["a"]), flat(lets["c"])))


# ---- groups/m31_point.simf:117-158 ---------------------------------------------------------
def test_m31_point_kats(orc, kat):
    f = S + "groups/m31_point.simf"
    lets, _ = kat(f, "test_m31_point_add_1")
    assert list(orc.m31_point_add(lets["g4"], lets["g4"])) == lets["expected"]
    lets, _ = kat(f, "test_m31_point_add_2")
    assert list(orc.m31_point_add(lets["point_1"], lets["point_2"])) == lets["expected"]
    lets, _ = kat(f, "test_m31_point_zero")
    assert lets["expected"] == [1, 0]
    lets, _ = kat(f, "test_m31_point_add_zero")
    assert list(orc.m31_point_add(lets["point_1"], [1, 0])) == lets["point_1"]
    lets, _ = kat(f, "test_m31_point_dbl")
    assert list(orc.m31_point_dbl(lets["point"])) == lets["expected"]
    lets, _ = kat(f, "test_circle_point_index_to_m31_point")
    assert list(orc.circle_point_index_to_m31_point(lets["point_index"])) == lets["expected"]


# ---- groups/qm31_point.simf:77-97 ------------------------------------------------------------
QM31_CIRCLE_GEN = [1, 0, 478637715, 513582971, 992285211, 649143431, 740191619, 1186584352]  # qm31_point.simf:14
M31_CIRCLE_GEN = [2, 1268011823]  # m31_point.simf:13


def test_add_circle_point_m31(orc):
    res = orc.qm31_point_add_m31_point(QM31_CIRCLE_GEN, M31_CIRCLE_GEN)
    as_q = [M31_CIRCLE_GEN[0], 0, 0, 0, M31_CIRCLE_GEN[1], 0, 0, 0]
    assert list(res) == list(orc.qm31_point_add(QM31_CIRCLE_GEN, as_q))


def test_qm31_point_neg(orc):
    """qm31_point.simf:89-97: 3g + (-3g) == zero under qm31_point_eq (bitwise), so the sum is canonical."""
    p = orc.qm31_point_add(orc.qm31_point_add(QM31_CIRCLE_GEN, QM31_CIRCLE_GEN), QM31_CIRCLE_GEN)
    assert list(orc.qm31_point_add(p, orc.qm31_point_neg(p))) == [1, 0, 0, 0, 0, 0, 0, 0]


# ---- groups/coset.simf:55-82, circle_domain.simf:47-68 --------------------------------------
def test_coset_kats(orc, kat):
    f = S + "groups/coset.simf"
    lets, asserts = kat(f, "test_bit_reverse_position")
    assert orc.bit_reverse_position(lets["index"], lets["log_size"]) == rhs(asserts, "reversed")
    lets, asserts = kat(f, "test_circle_point_index_add")
    assert orc.circle_point_index_add(lets["lhs"], lets["rhs"]) == rhs(asserts, "res")
    lets, asserts = kat(f, "test_circle_point_index_mul")
    assert orc.circle_point_index_mul(lets["lhs"], lets["rhs"]) == rhs(asserts, "res")
    lets, asserts = kat(f, "test_circle_point_index_neg")
    assert orc.circle_point_index_neg(lets["index"]) == rhs(asserts, "res")


def test_circle_domain_kats(orc, kat):
    f = S + "groups/circle_domain.simf"
    _, asserts = kat(f, "test_circle_domain")
    assert list(orc.circle_domain(11)) == [rhs(asserts, "half_size"), rhs(asserts, "offset"), rhs(asserts, "step")]
    for name in ("test_circle_position_to_point_index", "test_circle_position_to_point_index_2"):
        lets, asserts = kat(f, name)
        assert orc.circle_position_to_point_index(11, lets["position"]) == rhs(asserts, "point_index")


# ---- hasher.simf:108-118, merkle.simf:48-80 ---------------------------------------------------
def test_sha256_kats(orc, kat):
    lets, asserts = kat(S + "hasher.simf", "test_sha256")
    assert orc.sha256(lets["input"]) == rhs(asserts, "result")
    lets, asserts = kat(S + "hasher.simf", "test_sha256_32")
    assert orc.sha256_32(lets["input"]) == rhs(asserts, "result")


def test_sha256_against_hashlib(orc):
    import hashlib

    rng = np.random.default_rng(1)
    for n in (0, 1, 3, 55, 56, 63, 64, 65, 119, 120, 128, 352):
        data = rng.integers(0, 256, n, dtype=np.uint8).tobytes()
        assert orc.sha256_bytes(data) == hashlib.sha256(data).digest()


def test_merkle_kats(orc, kat):
    lets, _ = kat(S + "merkle.simf", "test_merkle")
    ok, _, path = orc.merkle_verify_32(orc.sha256(0), 4, lets["proof"], lets["root"])
    assert ok and path == 1
    lets, _ = kat(S + "merkle.simf", "test_decommitment")
    leaf = orc.sha256_32(2915689030)  # merkle.simf:60
    ok, _, path = orc.merkle_verify_32(leaf, lets["leaf_id"] + 8192, lets["proof"], lets["root"])
    assert ok and path == 1
    # merkle.simf:42: a proof with one sibling too few / too many cannot end with path == 1
    ok, _, path = orc.merkle_verify_32(leaf, lets["leaf_id"] + 8192, lets["proof"][:-1], lets["root"])
    assert not ok and path in (2, 3)
    bad = list(lets["proof"])
    bad[3] ^= 1
    assert not orc.merkle_verify_32(leaf, lets["leaf_id"] + 8192, bad, lets["root"])[0]


# ---- channel.simf:176-196, pow.simf:39-52, fri/queries.simf:47-61 -----------------------------
def test_channel_draw_qm31(orc, kat):
    lets, _ = kat(S + "channel.simf", "test_channel_draw_qm31")
    st = orc.state(*lets["state"])
    st, v, fail = orc.channel_draw_qm31(st)
    assert not fail and list(v) == flat(lets["first_random_felt"])
    st, v, fail = orc.channel_draw_qm31(st)
    assert not fail and list(v) == flat(lets["second_random_felt"])


def test_channel_draw_qm31_point(orc, kat):
    lets, _ = kat(S + "channel.simf", "test_channel_draw_qm31_point")
    _, p, fail = orc.channel_draw_qm31_point(orc.state(*lets["state"]))
    assert not fail and list(p) == flat(lets["x"]) + flat(lets["y"])


def test_pow_kats(orc, kat):
    lets, asserts = kat(S + "pow.simf", "test_reverse_bytes_32")
    assert orc.reverse_bytes_32(lets["value"]) == rhs(asserts, "reversed")
    lets, asserts = kat(S + "pow.simf", "test_check_proof_of_work")
    st, ok = orc.check_proof_of_work(orc.state(*lets["state"]), lets["nonce"], 0x07FFFFFFFFFFFFFF)
    assert ok and [int(x) for x in st[:8]] == [(rhs(asserts, "digest") >> (32 * (7 - i))) & 0xFFFFFFFF for i in range(8)]
    # an impossible target rejects (pow.simf:32)
    assert not orc.check_proof_of_work(orc.state(*lets["state"]), lets["nonce"], 1)[1]


def test_channel_draw_queries_8(orc, kat):
    lets, asserts = kat(S + "fri/queries.simf", "test_channel_draw_queries_8")
    assert lets["query_mask"] == 63
    _, q = orc.channel_draw_queries(orc.state(*lets["state"]), 6, 8)
    assert list(q) == [rhs(asserts, f"q{i}") for i in range(8)]


# ---- evals/commit.simf:39-51, deep/oods.simf:68-134, fri/commit.simf:89-107 --------------------
def digest_of(st):
    r = 0
    for x in st[:8]:
        r = (r << 32) | int(x)
    return r


def test_evals_commit(orc, kat):
    lets, asserts = kat(S + "evals/commit.simf", "test_evals_commit")
    st, coeff = orc.evals_commit(orc.state(*lets["state"]), lets["commitments"])
    assert digest_of(st) == rhs(asserts, "digest") and list(coeff) == flat(rhs(asserts, "random_coeff"))


def test_oods_kats(orc, kat):
    lets, asserts = kat(S + "deep/oods.simf", "test_oods")
    st, alpha, _, ok = orc.oods(orc.state(*lets["state"]), lets["log_size"], flat(lets["oods_trace_evals"]), flat(lets["oods_cp_eval"]), flat(lets["random_coeff"]))
    assert ok and digest_of(st) == rhs(asserts, "digest") and list(alpha) == flat(rhs(asserts, "deep_alpha"))
    lets, _ = kat(S + "deep/oods.simf", "test_channel_mix_oods_evals")
    st = orc.channel_mix_oods_evals(orc.state(*lets["state"]), flat(lets["oods_trace_evals"]), flat(lets["oods_cp_eval"]))
    assert digest_of(st) == lets["expected"] and int(st[8]) == 0
    # deep/oods.simf:58: a wrong sampled CP value must fail the assert
    lets, _ = kat(S + "deep/oods.simf", "test_oods")
    cp = flat(lets["oods_cp_eval"])
    cp[0] = (cp[0] + 1) % 2147483647
    assert not orc.oods(orc.state(*lets["state"]), lets["log_size"], flat(lets["oods_trace_evals"]), cp, flat(lets["random_coeff"]))[3]


def test_fri_commit(orc, kat):
    lets, asserts = kat(S + "fri/commit.simf", "test_fri_commit")
    first, inner, last = lets["fri_commitments"]
    st, alphas = orc.fri_commit(orc.state(*lets["state"]), first, inner, flat(last))
    assert digest_of(st) == rhs(asserts, "digest") and list(alphas[0]) == flat(rhs(asserts, "first_alpha"))


# ---- constraints / composition poly -------------------------------------------------------------
def test_eval_composition_poly(orc, kat):
    lets, _ = kat(S + "constraints/wide_fibonacci.simf", "test_eval_composition_poly")
    out, fail = orc.eval_composition_poly(lets["log_size"], flat(lets["oods_point"]), flat(lets["oods_trace_evals"]), flat(lets["random_coeff"]))
    assert not fail and list(out) == flat(lets["expected"])


def test_composition_poly_kats(orc, kat):
    lets, _ = kat(S + "evals/composition_poly.simf", "test_composition_poly_eval_from_partitions")
    assert list(orc.composition_poly_eval_from_partitions(flat(lets["partitioned_cp_eval"]))) == flat(lets["expected"])
    lets, _ = kat(S + "evals/composition_poly.simf", "test_vanishing_poly_eval")
    assert list(orc.vanishing_poly_eval(lets["log_size"], flat(lets["point"]))) == flat(lets["expected"])


# ---- deep/quotients.simf:48-79 --------------------------------------------------------------------
def test_quotient_kats(orc, kat):
    f = S + "deep/quotients.simf"
    lets, asserts = kat(f, "test_quotient_denominator_inverse")
    out, fail = orc.deep_quotient_denominator_inverse(flat(lets["sample_point"]), lets["query_point"])
    assert not fail and list(out) == rhs(asserts, "denominator_inv")
    lets, asserts = kat(f, "test_deep_quotient_nominator")
    coeffs = flat(lets["a"]) + flat(lets["b"]) + flat(lets["c"])
    assert list(orc.deep_quotient_nominator(coeffs, lets["query_point"], lets["query_value"])) == flat(rhs(asserts, "nominator"))
    lets, _ = kat(f, "test_deep_quotient_interpolant_coefficients")
    out = orc.deep_quotient_interpolant_coefficients(flat(lets["sample_point"]), flat(lets["sample_value"]), flat(lets["alpha_i"]))
    assert [list(r) for r in out] == [flat(lets["a"]), flat(lets["b"]), flat(lets["c"])]


# ---- evals/verify.simf:127-148 ---------------------------------------------------------------------
def test_verify_query(orc, kat):
    lets, _ = kat(S + "evals/verify.simf", "test_verify_query")
    (tvals, tproof), (cvals, cproof) = lets["decommitment"]
    _, trace_root, cp_root = lets["roots"]
    auth = lets["query"] + lets["domain_size"]
    ok, _, path = orc.merkle_verify_32(orc.hash_node_m31_trace(flat(tvals)), auth, tproof, trace_root)
    assert ok and path == 1
    ok, _, path = orc.merkle_verify_32(orc.hash_node_m31_cp(cvals), auth, cproof, cp_root)
    assert ok and path == 1


# ---- fri/folding.simf:45-65, fri/layers.simf:82-130 --------------------------------------------------
def test_fold_kats(orc, kat):
    lets, asserts = kat(S + "fri/folding.simf", "test_circle_fold")
    out, fail = orc.circle_fold(lets["query"], flat(lets["f_p"]), flat(lets["f_neg_p"]), lets["log_size_ex"], flat(lets["fold_alpha"]))
    assert not fail and list(out) == flat(rhs(asserts, "folded_eval"))
    lets, asserts = kat(S + "fri/folding.simf", "test_line_fold")
    out, fail = orc.line_fold(lets["query"], flat(lets["f_x"]), flat(lets["f_neg_x"]), lets["log_size_ex"], flat(lets["fold_alpha"]))
    assert not fail and list(out) == flat(rhs(asserts, "folded_eval"))


def test_fri_layer_kats(orc, kat):
    f = S + "fri/layers.simf"
    lets, _ = kat(f, "test_verify_decommitment")
    assert orc.verify_decommitment(lets["position"], flat(lets["eval0"]), flat(lets["eval1"]), lets["log_size_ex"], lets["proof"], lets["root"])
    # fri_verify_query (layers.simf:51-69) on the two layer KATs; contexts are inline at layers.simf:106-110,123-127
    ctx = {
        "test_fri_verify_first_layer": (0x26DA5011FE955BE570DA501AB3D42F3903913FA59554A6EC8BFBBC9C66D84B5B, [1516394272, 915498982, 1578049480, 1826337248], 4, True),
        "test_fri_verify_inner_layer": (0x0D11AA22F18AF5F6F5E8F7A844D359A82BFFB5F18C12A29722A09884BC3A7B17, [428468021, 292366470, 1298858467, 227984395], 3, False),
    }
    folded = {}
    for name, (root, alpha, log, first) in ctx.items():
        lets, _ = kat(f, name)
        (query, evaluation), (witness, proof) = lets["data"]
        assert query % 2 == 0  # both KATs query an even (left) leaf
        assert orc.verify_decommitment(query, flat(evaluation), flat(witness), log, proof, root)
        fold = orc.circle_fold if first else orc.line_fold
        folded[name], fail = fold(query, flat(evaluation), flat(witness), log, alpha)
        assert not fail
    # the first layer's folded value is the inner layer KAT's evaluation (layers.simf:115 == folding.simf:52)
    lets, _ = kat(f, "test_fri_verify_inner_layer")
    assert list(folded["test_fri_verify_first_layer"]) == flat(lets["data"][0][1])
