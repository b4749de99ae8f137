"""ctypes wrapper over oracle/liboracle.so — TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module (see oracle/ssym_oracle.c).  The product package never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import Optional, Sequence, Tuple

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle.so")

u32p = C.POINTER(C.c_uint32)
MODE_REF_LITERAL = 0
MODE_PROVER_CONSISTENT = 1
MODE_QUERY_DEDUP = 2  # flag OR-ed into either: sorted, de-duplicated queries (include/ssym.h)


class StwoConfig(C.Structure):
    _fields_ = [
        ("trace_log", C.c_uint32),
        ("lde_log", C.c_uint32),
        ("n_queries", C.c_uint32),
        ("n_fri_layers", C.c_uint32),
        ("mode", C.c_uint32),
        ("n_columns", C.c_uint32),
        ("pow_target", C.c_uint64),
    ]


class StwoLayout(C.Structure):
    _fields_ = [
        ("off_commit", C.c_uint32), ("off_oods_trace", C.c_uint32), ("off_oods_cp", C.c_uint32),
        ("off_fri_first_root", C.c_uint32), ("off_fri_inner_root", C.c_uint32), ("off_last_coeff", C.c_uint32),
        ("off_pow_nonce", C.c_uint32), ("off_qvals", C.c_uint32), ("off_trace_sib", C.c_uint32),
        ("off_cp_sib", C.c_uint32), ("off_fri_wit", C.c_uint32), ("off_fri_sib", C.c_uint32 * 9),
        ("stride_words", C.c_uint32), ("algorithmic_bytes", C.c_uint32),
    ]


class StwoTrace(C.Structure):
    _fields_ = [
        ("status", C.c_uint32), ("first_fail", C.c_uint32),
        ("digest_commit", C.c_uint32 * 8), ("cp_alpha", C.c_uint32 * 4),
        ("oods_x", C.c_uint32 * 4), ("oods_y", C.c_uint32 * 4),
        ("cp_eval", C.c_uint32 * 4), ("cp_sampled", C.c_uint32 * 4),
        ("digest_oods", C.c_uint32 * 8), ("deep_alpha", C.c_uint32 * 4),
        ("fri_alpha", (C.c_uint32 * 4) * 9),
        ("digest_fri", C.c_uint32 * 8), ("digest_pow", C.c_uint32 * 8),
        ("pow_value", C.c_uint32 * 2),
        ("queries", C.c_uint32 * 16),
        ("fri_answer", (C.c_uint32 * 4) * 16),
        ("folded", ((C.c_uint32 * 4) * 16) * 9),
        ("trace_root", (C.c_uint32 * 8) * 16), ("cp_root", (C.c_uint32 * 8) * 16),
        ("fri_root", ((C.c_uint32 * 8) * 16) * 9),
        ("mask_trace", C.c_uint32), ("mask_cp", C.c_uint32), ("mask_answer_inv", C.c_uint32),
        ("mask_fri", C.c_uint32 * 9), ("mask_fold_inv", C.c_uint32 * 9),
        ("mask_last_query", C.c_uint32), ("mask_last_eval", C.c_uint32),
        ("draw_retries", C.c_uint32), ("n_queries_used", C.c_uint32), ("pad_", C.c_uint32 * 1),
    ]


class S101Trace(C.Structure):
    _fields_ = [
        ("status", C.c_uint32), ("first_fail_layer", C.c_uint32),
        ("alpha", C.c_uint32 * 3), ("idx", C.c_uint32), ("x", C.c_uint32), ("cp0", C.c_uint32),
        ("n_layers", C.c_uint32),
        ("beta_drawn", C.c_uint32 * 31), ("cp_ev", C.c_uint32 * 32), ("layer_mask", C.c_uint32 * 31),
        ("state_final", C.c_uint32 * 8), ("trace_root", (C.c_uint32 * 8) * 3),
        ("query_ordinal", C.c_uint32), ("commit_state", C.c_uint32 * 8),
    ]


PRESETS = {
    # stwo-verifier/src/config.simf:16-33 (TESTING) and :34-52 (production)
    "testing": dict(trace_log=3, lde_log=4, n_queries=1, n_fri_layers=2, pow_target=0x07FFFFFFFFFFFFFF),
    "prod": dict(trace_log=9, lde_log=13, n_queries=16, n_fri_layers=8, pow_target=0x07FFFFFFFFFFFFFF),
}


def make_config(preset: str, mode: int, n_columns: int = 4) -> StwoConfig:
    """A preset of config.simf:10-51 with NUM_COLUMNS = n_columns (config.simf:14; 4 at reference HEAD)."""
    p = PRESETS[preset]
    return StwoConfig(p["trace_log"], p["lde_log"], p["n_queries"], p["n_fri_layers"], mode, n_columns, p["pow_target"])


def build(force: bool = False) -> str:
    """Compile liboracle.so from oracle/ssym_oracle.c (gcc).  Building the checker is not using it."""
    src = os.path.join(_HERE, "ssym_oracle.c")
    hdr = os.path.join(_HERE, "..", "include", "ssym.h")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < max(os.path.getmtime(src), os.path.getmtime(hdr), os.path.getmtime(os.path.join(_HERE, "stwo_prover_ref.c"))):
        subprocess.check_call(["make", "-s", "-C", _HERE, "liboracle.so"])
    return _LIB_PATH


def words(x) -> np.ndarray:
    return np.ascontiguousarray(np.asarray(x, dtype=np.uint64).astype(np.uint32).ravel())


def ptr(a: np.ndarray):
    return a.ctypes.data_as(u32p)


def u256_words(v: int) -> np.ndarray:
    return np.array([(v >> (32 * (7 - i))) & 0xFFFFFFFF for i in range(8)], dtype=np.uint32)


def words_u256(w: Sequence[int]) -> int:
    r = 0
    for x in w:
        r = (r << 32) | int(x)
    return r


class Oracle:
    """Thin, typed access to the C restatement.  Arrays are numpy uint32."""

    def __init__(self, path: Optional[str] = None):
        if path is None:
            path = _LIB_PATH
            if not os.path.exists(path):
                build()
        self.lib = C.CDLL(path)
        L = self.lib
        L.oracle_compression_count.restype = C.c_uint64
        L.oracle_sizeof_stwo_trace.restype = C.c_size_t
        L.oracle_sizeof_s101_trace.restype = C.c_size_t
        for name in ("oracle_m31", "oracle_m31_add", "oracle_m31_neg", "oracle_m31_sub", "oracle_m31_mul", "oracle_m31_exp",
                     "oracle_m31_inv", "oracle_bit_reverse_position", "oracle_circle_point_index_add",
                     "oracle_circle_point_index_mul", "oracle_circle_point_index_neg",
                     "oracle_circle_position_to_point_index", "oracle_line_position_to_x_coord", "oracle_reverse_bytes_32",
                     "oracle_s101_add_mod", "oracle_s101_sub_mod", "oracle_s101_mul_mod", "oracle_s101_div_mod",
                     "oracle_s101_exp_mod", "oracle_s101_channel_draw_32", "oracle_s101_calc_x", "oracle_s101_eval_p0",
                     "oracle_s101_eval_cp", "oracle_s101_fri_eval_cp_next"):
            getattr(L, name).restype = C.c_uint32
        L.oracle_check_proof_of_work.argtypes = [u32p, C.c_uint32, C.c_uint32, C.c_uint64]
        L.oracle_stwo_verify_batch.argtypes = [C.POINTER(StwoConfig), u32p, C.c_size_t, C.c_size_t, u32p, u32p, C.c_void_p]
        L.oracle_s101_verify_batch.argtypes = [u32p, C.POINTER(C.c_uint64), C.c_size_t, C.c_size_t, u32p, u32p, C.c_void_p]
        L.oracle_s101_group.argtypes = [C.c_size_t, C.c_uint32, u32p, C.c_void_p, u32p]
        L.oracle_s101_group.restype = None
        L.oracle_sha256_bytes.argtypes = [C.c_char_p, C.c_size_t, u32p]
        assert L.oracle_sizeof_stwo_trace() == C.sizeof(StwoTrace), (L.oracle_sizeof_stwo_trace(), C.sizeof(StwoTrace))
        assert L.oracle_sizeof_s101_trace() == C.sizeof(S101Trace), (L.oracle_sizeof_s101_trace(), C.sizeof(S101Trace))

    # ---- generic helpers -------------------------------------------------------------
    def _call_out(self, fn: str, n_out: int, *args) -> np.ndarray:
        out = np.zeros(n_out, dtype=np.uint32)
        cargs = [ptr(a) if isinstance(a, np.ndarray) else C.c_uint32(int(a)) for a in args]
        getattr(self.lib, fn)(*cargs, ptr(out))
        return out

    def _call_out_fail(self, fn: str, n_out: int, *args) -> Tuple[np.ndarray, bool]:
        out = np.zeros(n_out, dtype=np.uint32)
        fail = C.c_int(0)
        cargs = [ptr(a) if isinstance(a, np.ndarray) else C.c_uint32(int(a)) for a in args]
        getattr(self.lib, fn)(*cargs, ptr(out), C.byref(fail))
        return out, bool(fail.value)

    # ---- fields ------------------------------------------------------------------------
    def m31(self, v): return self.lib.oracle_m31(C.c_uint32(v))
    def m31_add(self, a, b): return self.lib.oracle_m31_add(C.c_uint32(a), C.c_uint32(b))
    def m31_sub(self, a, b): return self.lib.oracle_m31_sub(C.c_uint32(a), C.c_uint32(b))
    def m31_neg(self, a): return self.lib.oracle_m31_neg(C.c_uint32(a))
    def m31_mul(self, a, b): return self.lib.oracle_m31_mul(C.c_uint32(a), C.c_uint32(b))
    def m31_exp(self, a, b): return self.lib.oracle_m31_exp(C.c_uint32(a), C.c_uint32(b))

    def m31_inv(self, a) -> Tuple[int, bool]:
        fail = C.c_int(0)
        r = self.lib.oracle_m31_inv(C.c_uint32(a), C.byref(fail))
        return r, bool(fail.value)

    def cm31_add(self, a, b): return self._call_out("oracle_cm31_add", 2, words(a), words(b))
    def cm31_sub(self, a, b): return self._call_out("oracle_cm31_sub", 2, words(a), words(b))
    def cm31_mul(self, a, b): return self._call_out("oracle_cm31_mul", 2, words(a), words(b))
    def cm31_div(self, a, b): return self._call_out_fail("oracle_cm31_div", 2, words(a), words(b))
    def cm31_inv(self, a): return self._call_out_fail("oracle_cm31_inv", 2, words(a))
    def qm31_add(self, a, b): return self._call_out("oracle_qm31_add", 4, words(a), words(b))
    def qm31_sub(self, a, b): return self._call_out("oracle_qm31_sub", 4, words(a), words(b))
    def qm31_neg(self, a): return self._call_out("oracle_qm31_neg", 4, words(a))
    def qm31_conj(self, a): return self._call_out("oracle_qm31_conj", 4, words(a))
    def qm31_mul(self, a, b): return self._call_out("oracle_qm31_mul", 4, words(a), words(b))
    def qm31_mul_m31(self, a, b): return self._call_out("oracle_qm31_mul_m31", 4, words(a), int(b))
    def qm31_mul_cm31(self, a, b): return self._call_out("oracle_qm31_mul_cm31", 4, words(a), words(b))
    def qm31_inv(self, a): return self._call_out_fail("oracle_qm31_inv", 4, words(a))
    def qm31_div(self, a, b): return self._call_out_fail("oracle_qm31_div", 4, words(a), words(b))

    # ---- groups ------------------------------------------------------------------------
    def m31_point_add(self, a, b): return self._call_out("oracle_m31_point_add", 2, words(a), words(b))
    def m31_point_dbl(self, a): return self._call_out("oracle_m31_point_dbl", 2, words(a))
    def m31_point_neg(self, a): return self._call_out("oracle_m31_point_neg", 2, words(a))
    def circle_point_index_to_m31_point(self, idx): return self._call_out("oracle_circle_point_index_to_m31_point", 2, int(idx))
    def qm31_point_add(self, a, b): return self._call_out("oracle_qm31_point_add", 8, words(a), words(b))
    def qm31_point_neg(self, a): return self._call_out("oracle_qm31_point_neg", 8, words(a))
    def qm31_point_add_m31_point(self, a, b): return self._call_out("oracle_qm31_point_add_m31_point", 8, words(a), words(b))
    def bit_reverse_position(self, pos, log): return self.lib.oracle_bit_reverse_position(C.c_uint32(pos), C.c_uint32(log))
    def circle_point_index_add(self, a, b): return self.lib.oracle_circle_point_index_add(C.c_uint32(a), C.c_uint32(b))
    def circle_point_index_mul(self, a, b): return self.lib.oracle_circle_point_index_mul(C.c_uint32(a), C.c_uint32(b))
    def circle_point_index_neg(self, a): return self.lib.oracle_circle_point_index_neg(C.c_uint32(a))
    def circle_domain(self, log): return self._call_out("oracle_circle_domain", 3, int(log))
    def circle_position_to_point_index(self, log, pos): return self.lib.oracle_circle_position_to_point_index(C.c_uint32(log), C.c_uint32(pos))
    def line_position_to_x_coord(self, log, pos): return self.lib.oracle_line_position_to_x_coord(C.c_uint32(log), C.c_uint32(pos))

    # ---- hashing -----------------------------------------------------------------------
    def sha256(self, v: int) -> int: return words_u256(self._call_out("oracle_sha256", 8, u256_words(v)))
    def sha256_32(self, v: int) -> int: return words_u256(self._call_out("oracle_sha256_32", 8, int(v)))
    def sha256_pair(self, l: int, r: int) -> int: return words_u256(self._call_out("oracle_sha256_pair", 8, u256_words(l), u256_words(r)))

    def sha256_bytes(self, data: bytes) -> bytes:
        out = np.zeros(8, dtype=np.uint32)
        self.lib.oracle_sha256_bytes(data, len(data), ptr(out))
        return b"".join(int(x).to_bytes(4, "big") for x in out)

    def hash_node_m31_trace(self, e): return words_u256(self._call_out("oracle_hash_node_m31_trace", 8, words(e)))
    def hash_node_m31_cp(self, e): return words_u256(self._call_out("oracle_hash_node_m31_cp", 8, words(e)))
    def hash_node_qm31(self, e): return words_u256(self._call_out("oracle_hash_node_qm31", 8, words(e)))

    def merkle_verify_32(self, leaf: int, auth_path: int, proof: Sequence[int], root: int):
        """merkle.simf:39-44 -> (ok, computed_root, final_path)."""
        sib = np.concatenate([u256_words(s) for s in proof]) if len(proof) else np.zeros(0, dtype=np.uint32)
        croot = np.zeros(8, dtype=np.uint32)
        fpath = C.c_uint32(0)
        ok = self.lib.oracle_merkle_verify_32(ptr(u256_words(leaf)), C.c_uint32(auth_path), ptr(sib), C.c_uint32(len(proof)),
                                              ptr(u256_words(root)), ptr(croot), C.byref(fpath))
        return bool(ok), words_u256(croot), fpath.value

    # ---- channel -----------------------------------------------------------------------
    @staticmethod
    def state(digest: int, n_sent: int = 0) -> np.ndarray:
        return np.concatenate([u256_words(digest), np.array([n_sent], dtype=np.uint32)])

    def channel_mix_u256(self, st, v: int):
        st = st.copy(); self.lib.oracle_channel_mix_u256(ptr(st), ptr(u256_words(v))); return st

    def channel_mix_u64(self, st, v: int):
        st = st.copy(); self.lib.oracle_channel_mix_u64(ptr(st), C.c_uint32(v >> 32), C.c_uint32(v & 0xFFFFFFFF)); return st

    def channel_draw_qm31(self, st):
        st = st.copy(); out = np.zeros(4, dtype=np.uint32); fail = C.c_int(0)
        self.lib.oracle_channel_draw_qm31(ptr(st), ptr(out), C.byref(fail)); return st, out, bool(fail.value)

    def channel_draw_m31x8(self, st):
        st = st.copy(); out = np.zeros(8, dtype=np.uint32); fail = C.c_int(0)
        self.lib.oracle_channel_draw_m31x8(ptr(st), ptr(out), C.byref(fail)); return st, out, bool(fail.value)

    def channel_draw_qm31_point(self, st):
        st = st.copy(); out = np.zeros(8, dtype=np.uint32); fail = C.c_int(0)
        self.lib.oracle_channel_draw_qm31_point(ptr(st), ptr(out), C.byref(fail)); return st, out, bool(fail.value)

    def channel_draw_queries(self, st, log_size: int, n_queries: int):
        st = st.copy(); out = np.zeros(n_queries, dtype=np.uint32)
        self.lib.oracle_channel_draw_queries(ptr(st), C.c_uint32(log_size), C.c_uint32(n_queries), ptr(out)); return st, out

    def reverse_bytes_32(self, v): return self.lib.oracle_reverse_bytes_32(C.c_uint32(v))

    def check_proof_of_work(self, st, nonce: int, target: int):
        st = st.copy()
        ok = self.lib.oracle_check_proof_of_work(ptr(st), C.c_uint32(nonce >> 32), C.c_uint32(nonce & 0xFFFFFFFF), C.c_uint64(target))
        return st, bool(ok)

    def evals_commit(self, st, commitments: Sequence[int]):
        st = st.copy(); coeff = np.zeros(4, dtype=np.uint32)
        cw = np.concatenate([u256_words(c) for c in commitments])
        self.lib.oracle_evals_commit(ptr(st), ptr(cw), ptr(coeff)); return st, coeff

    def composition_poly_eval_from_partitions(self, parts): return self._call_out("oracle_composition_poly_eval_from_partitions", 4, words(parts))
    def vanishing_poly_eval(self, log, point): return self._call_out("oracle_vanishing_poly_eval", 4, int(log), words(point))
    def eval_composition_poly(self, log, point, oods_trace, coeff): return self._call_out_fail("oracle_eval_composition_poly", 4, int(log), words(point), words(oods_trace), words(coeff))

    def channel_mix_oods_evals(self, st, oods_trace, oods_cp):
        st = st.copy(); self.lib.oracle_channel_mix_oods_evals(ptr(st), ptr(words(oods_trace)), ptr(words(oods_cp))); return st

    def oods(self, st, log_size, oods_trace, oods_cp, cp_alpha):
        st = st.copy(); alpha = np.zeros(4, dtype=np.uint32); point = np.zeros(8, dtype=np.uint32)
        ok = self.lib.oracle_oods(ptr(st), C.c_uint32(log_size), ptr(words(oods_trace)), ptr(words(oods_cp)), ptr(words(cp_alpha)), ptr(alpha), ptr(point))
        return st, alpha, point, bool(ok)

    def fri_commit(self, st, first_root: int, inner_roots: Sequence[int], last_coeff):
        st = st.copy(); n = len(inner_roots); alphas = np.zeros(4 * (n + 1), dtype=np.uint32)
        inner = np.concatenate([u256_words(r) for r in inner_roots]) if n else np.zeros(0, dtype=np.uint32)
        self.lib.oracle_fri_commit(ptr(st), ptr(u256_words(first_root)), ptr(inner), C.c_uint32(n), ptr(words(last_coeff)), ptr(alphas))
        return st, alphas.reshape(n + 1, 4)

    def deep_quotient_denominator_inverse(self, sp, qp): return self._call_out_fail("oracle_deep_quotient_denominator_inverse", 2, words(sp), words(qp))
    def deep_quotient_interpolant_coefficients(self, sp, sv, alpha): return self._call_out("oracle_deep_quotient_interpolant_coefficients", 12, words(sp), words(sv), words(alpha)).reshape(3, 4)
    def deep_quotient_nominator(self, coeffs, qp, qv): return self._call_out("oracle_deep_quotient_nominator", 4, words(coeffs), words(qp), int(qv))

    def fri_answer(self, mode, query, trace_evals, cp_evals, coeff, point, oods_trace, oods_cp, log_size):
        return self._call_out_fail("oracle_fri_answer", 4, int(mode), int(query), words(trace_evals), words(cp_evals), words(coeff),
                                   words(point), words(oods_trace), words(oods_cp), int(log_size))

    def circle_fold(self, position, f_p, f_neg_p, log_size, alpha): return self._call_out_fail("oracle_circle_fold", 4, int(position), words(f_p), words(f_neg_p), int(log_size), words(alpha))
    def line_fold(self, position, f_p, f_neg_p, log_size, alpha): return self._call_out_fail("oracle_line_fold", 4, int(position), words(f_p), words(f_neg_p), int(log_size), words(alpha))

    def verify_decommitment(self, position, eval0, eval1, log_size, proof: Sequence[int], root: int) -> bool:
        sib = np.concatenate([u256_words(s) for s in proof]) if len(proof) else np.zeros(0, dtype=np.uint32)
        return bool(self.lib.oracle_verify_decommitment(C.c_uint32(position), ptr(words(eval0)), ptr(words(eval1)), C.c_uint32(log_size),
                                                        ptr(sib), C.c_uint32(len(proof)), ptr(u256_words(root))))

    # ---- whole proofs --------------------------------------------------------------------
    def stwo_layout(self, cfg: StwoConfig) -> StwoLayout:
        lo = StwoLayout()
        rc = self.lib.oracle_stwo_layout(C.byref(cfg), C.byref(lo))
        if rc != 0:
            raise ValueError("bad stwo config")
        return lo

    def stwo_verify_batch(self, cfg: StwoConfig, packed: np.ndarray, n: int, want_trace: bool = False, begin: int = 0, end: Optional[int] = None):
        end = n if end is None else end
        packed = np.ascontiguousarray(packed, dtype=np.uint32)
        accept = np.zeros((n + 31) // 32, dtype=np.uint32)
        status = np.zeros(n, dtype=np.uint32)
        traces = (StwoTrace * n)() if want_trace else None
        self.lib.oracle_stwo_verify_batch(C.byref(cfg), ptr(packed), begin, end, ptr(accept), ptr(status),
                                          C.cast(traces, C.c_void_p) if want_trace else None)
        return accept, status, traces

    def stwo_prove_batch(self, cfg: StwoConfig, seeds, threads: int = 1) -> np.ndarray:
        """CPU reference prover (oracle/stwo_prover_ref.c): one packed proof per seed, shape (n, stride_words)."""
        lo = self.stwo_layout(cfg)
        seeds = np.ascontiguousarray(seeds, dtype=np.uint64)
        n = len(seeds)
        out = np.zeros((n, lo.stride_words), dtype=np.uint32)
        fn = self.lib.oracle_stwo_prove_batch
        fn.restype = C.c_int
        sp = seeds.ctypes.data_as(C.POINTER(C.c_uint64))

        def run(b, e):
            return fn(C.byref(cfg), sp, C.c_size_t(b), C.c_size_t(e), ptr(out))

        if threads <= 1 or n <= 1:
            rcs = [run(0, n)]
        else:
            from concurrent.futures import ThreadPoolExecutor
            threads = min(threads, n)
            cuts = [n * i // threads for i in range(threads + 1)]
            with ThreadPoolExecutor(threads) as ex:
                rcs = list(ex.map(lambda i: run(cuts[i], cuts[i + 1]), range(threads)))
        if any(rcs):
            raise RuntimeError(f"reference prover failed its own low-degree checks: {rcs}")
        return out

    def s101_verify_batch(self, blob: np.ndarray, offsets: np.ndarray, want_trace: bool = False):
        n = len(offsets) - 1
        blob = np.ascontiguousarray(blob, dtype=np.uint32)
        offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
        accept = np.zeros((n + 31) // 32, dtype=np.uint32)
        status = np.zeros(n, dtype=np.uint32)
        traces = (S101Trace * n)() if want_trace else None
        self.lib.oracle_s101_verify_batch(ptr(blob), offsets.ctypes.data_as(C.POINTER(C.c_uint64)), 0, n, ptr(accept), ptr(status),
                                          C.cast(traces, C.c_void_p) if want_trace else None)
        return accept, status, traces

    def s101_verify_multi_batch(self, blob: np.ndarray, offsets: np.ndarray, n_queries: int):
        """ssym_stark101_verify_multi_batch (include/ssym.h): records proof-major, n_queries per proof -> (accept bits per proof, status per record, traces)."""
        n = len(offsets) - 1
        assert n % n_queries == 0
        _, status, traces = self.s101_verify_batch(blob, offsets, want_trace=True)
        accept = np.zeros((n // n_queries + 31) // 32, dtype=np.uint32)
        self.lib.oracle_s101_group(C.c_size_t(n // n_queries), C.c_uint32(n_queries), ptr(status), C.cast(traces, C.c_void_p), ptr(accept))
        return accept, status, traces

    # ---- stark101 functions --------------------------------------------------------------
    def s101_add_mod(self, a, b): return self.lib.oracle_s101_add_mod(C.c_uint32(a), C.c_uint32(b))
    def s101_sub_mod(self, a, b): return self.lib.oracle_s101_sub_mod(C.c_uint32(a), C.c_uint32(b))
    def s101_mul_mod(self, a, b): return self.lib.oracle_s101_mul_mod(C.c_uint32(a), C.c_uint32(b))
    def s101_exp_mod(self, a, b): return self.lib.oracle_s101_exp_mod(C.c_uint32(a), C.c_uint32(b))

    def s101_div_mod(self, a, b):
        fail = C.c_int(0)
        r = self.lib.oracle_s101_div_mod(C.c_uint32(a), C.c_uint32(b), C.byref(fail))
        return r, bool(fail.value)

    def s101_channel_draw_32(self, state: int, mx: int):
        st = u256_words(state)
        v = self.lib.oracle_s101_channel_draw_32(ptr(st), C.c_uint32(mx))
        return words_u256(st), v

    def s101_channel_mix_32(self, state: int, v: int):
        st = u256_words(state); self.lib.oracle_s101_channel_mix_32(ptr(st), C.c_uint32(v)); return words_u256(st)

    def s101_merkle_verify_32(self, leaf: int, auth: int, proof: Sequence[int], root: int) -> bool:
        sib = np.concatenate([u256_words(s) for s in proof]) if len(proof) else np.zeros(0, dtype=np.uint32)
        return bool(self.lib.oracle_s101_merkle_verify_32(ptr(u256_words(leaf)), C.c_uint32(auth), ptr(sib), C.c_uint32(len(proof)), ptr(u256_words(root))))

    def s101_calc_x(self, idx): return self.lib.oracle_s101_calc_x(C.c_uint32(idx))
    def s101_eval_p0(self, x, f_x): return self.lib.oracle_s101_eval_p0(C.c_uint32(x), C.c_uint32(f_x))
    def s101_eval_cp(self, x, a0, a1, a2, f0, f1, f2): return self.lib.oracle_s101_eval_cp(*(C.c_uint32(v) for v in (x, a0, a1, a2, f0, f1, f2)))
    def s101_fri_eval_cp_next(self, cpa, cpb, x, beta): return self.lib.oracle_s101_fri_eval_cp_next(*(C.c_uint32(v) for v in (cpa, cpb, x, beta)))
    def s101_compute_auth_path(self, idx, size): return self._call_out("oracle_s101_compute_auth_path", 2, int(idx), int(size))

    def s101_fri_read_commitment(self, state: int, root: int, beta: int):
        st = u256_words(state)
        ok = self.lib.oracle_s101_fri_read_commitment(ptr(st), ptr(u256_words(root)), C.c_uint32(beta))
        return words_u256(st), bool(ok)

    def compression_count(self) -> int: return int(self.lib.oracle_compression_count())
    def compression_reset(self) -> None: self.lib.oracle_compression_reset()

    def set_fast(self, on: bool) -> int:
        """CPU-baseline speed switch (process-wide): word-wise absorbs, block-wise padding, Mersenne folding and, where the host has them, SHA-NI
        compressions instead of the literal byte-at-a-time port.  Same results bit for bit.  Returns 2 (SHA-NI in use), 1 (word-wise only) or 0 (off)."""
        return int(self.lib.oracle_set_fast_sha(1 if on else 0))
