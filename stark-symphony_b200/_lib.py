"""ctypes binding of libssym.so (include/ssym.h).  There is no CPU fallback: if the library cannot be
loaded, or no sm_100 device is present, every compute call raises."""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SSYM_LIB") or os.path.join(HERE, "libssym.so")  # SSYM_LIB: an alternative build of the same ABI (kernel A/B runs)

MEM_DEVICE, MEM_HOST = 0, 1
OK, ERR_USAGE, ERR_CUDA, ERR_PARSE, ERR_NOMEM, ERR_INTERNAL = 0, -1, -2, -3, -4, -5  # include/ssym.h
MODE_REF_LITERAL, MODE_PROVER_CONSISTENT = 0, 1
MODE_QUERY_DEDUP = 2  # flag: sorted, de-duplicated queries (fri/queries.simf:41; include/ssym.h)
MAX_QUERIES, MAX_FRI_LAYERS, S101_MAX_LIST = 16, 9, 31

u32p = C.POINTER(C.c_uint32)
u64p = C.POINTER(C.c_uint64)
u8p = C.POINTER(C.c_uint8)


class SsymError(RuntimeError):
    pass


class StwoConfig(C.Structure):
    _fields_ = [("trace_log", C.c_uint32), ("lde_log", C.c_uint32), ("n_queries", C.c_uint32), ("n_fri_layers", C.c_uint32),
                ("mode", C.c_uint32), ("n_columns", C.c_uint32), ("pow_target", C.c_uint64)]


class StwoLayout(C.Structure):
    _fields_ = [("off_commit", C.c_uint32), ("off_oods_trace", C.c_uint32), ("off_oods_cp", C.c_uint32), ("off_fri_first_root", C.c_uint32),
                ("off_fri_inner_root", C.c_uint32), ("off_last_coeff", C.c_uint32), ("off_pow_nonce", C.c_uint32), ("off_qvals", C.c_uint32),
                ("off_trace_sib", C.c_uint32), ("off_cp_sib", C.c_uint32), ("off_fri_wit", C.c_uint32), ("off_fri_sib", C.c_uint32 * MAX_FRI_LAYERS),
                ("stride_words", C.c_uint32), ("algorithmic_bytes", C.c_uint32)]


class StwoTrace(C.Structure):
    _fields_ = [
        ("status", C.c_uint32), ("first_fail", C.c_uint32),
        ("digest_commit", C.c_uint32 * 8), ("cp_alpha", C.c_uint32 * 4),
        ("oods_x", C.c_uint32 * 4), ("oods_y", C.c_uint32 * 4),
        ("cp_eval", C.c_uint32 * 4), ("cp_sampled", C.c_uint32 * 4),
        ("digest_oods", C.c_uint32 * 8), ("deep_alpha", C.c_uint32 * 4),
        ("fri_alpha", (C.c_uint32 * 4) * MAX_FRI_LAYERS),
        ("digest_fri", C.c_uint32 * 8), ("digest_pow", C.c_uint32 * 8),
        ("pow_value", C.c_uint32 * 2),
        ("queries", C.c_uint32 * MAX_QUERIES),
        ("fri_answer", (C.c_uint32 * 4) * MAX_QUERIES),
        ("folded", ((C.c_uint32 * 4) * MAX_QUERIES) * MAX_FRI_LAYERS),
        ("trace_root", (C.c_uint32 * 8) * MAX_QUERIES), ("cp_root", (C.c_uint32 * 8) * MAX_QUERIES),
        ("fri_root", ((C.c_uint32 * 8) * MAX_QUERIES) * MAX_FRI_LAYERS),
        ("mask_trace", C.c_uint32), ("mask_cp", C.c_uint32), ("mask_answer_inv", C.c_uint32),
        ("mask_fri", C.c_uint32 * MAX_FRI_LAYERS), ("mask_fold_inv", C.c_uint32 * MAX_FRI_LAYERS),
        ("mask_last_query", C.c_uint32), ("mask_last_eval", C.c_uint32),
        ("draw_retries", C.c_uint32), ("n_queries_used", C.c_uint32), ("pad_", C.c_uint32 * 1),
    ]


class S101Trace(C.Structure):
    _fields_ = [
        ("status", C.c_uint32), ("first_fail_layer", C.c_uint32),
        ("alpha", C.c_uint32 * 3), ("idx", C.c_uint32), ("x", C.c_uint32), ("cp0", C.c_uint32), ("n_layers", C.c_uint32),
        ("beta_drawn", C.c_uint32 * S101_MAX_LIST), ("cp_ev", C.c_uint32 * (S101_MAX_LIST + 1)), ("layer_mask", C.c_uint32 * S101_MAX_LIST),
        ("state_final", C.c_uint32 * 8), ("trace_root", (C.c_uint32 * 8) * 3),
        ("query_ordinal", C.c_uint32), ("commit_state", C.c_uint32 * 8),
    ]


# name -> (restype, argtypes): every symbol include/ssym.h declares
_V, _I, _U32, _SZ = C.c_void_p, C.c_int, C.c_uint32, C.c_size_t
SYMBOLS = {
    "ssym_stwo_config_preset": (_I, [C.c_char_p, _U32, C.POINTER(StwoConfig)]),
    "ssym_stwo_layout": (_I, [C.POINTER(StwoConfig), C.POINTER(StwoLayout)]),
    "ssym_create": (_I, [_I, C.POINTER(_V)]),
    "ssym_destroy": (None, [_V]),
    "ssym_last_error": (C.c_char_p, []),
    "ssym_version": (C.c_char_p, []),
    "ssym_set_stream": (_I, [_V, _V]),
    "ssym_pinned_alloc": (_V, [_SZ]),
    "ssym_pinned_free": (None, [_V]),
    "ssym_synchronize": (_I, [_V]),
    "ssym_set_pipeline_depth": (_I, [_V, _I]),
    "ssym_join": (_I, [_V]),
    "ssym_set_host_async": (_I, [_V, _I]),
    "ssym_set_wit_host_fallback": (_I, [_V, _I]),
    "ssym_set_merkle_sharing": (_I, [_V, _I]),
    "ssym_launch_count": (C.c_uint64, [_V]),
    "ssym_profile_enable": (_I, [_V, _I]),
    "ssym_profile_read": (_I, [_V, C.POINTER(C.c_double), C.POINTER(C.c_uint64)]),
    "ssym_stwo_verify_batch": (_I, [_V, C.POINTER(StwoConfig), _V, _SZ, _V, _V, _V, _I]),
    "ssym_stwo_compact_bound": (_SZ, [C.POINTER(StwoConfig), _SZ]),
    "ssym_stwo_compact_pack": (_I, [C.POINTER(StwoConfig), _V, _SZ, _V, _SZ, _V]),
    "ssym_stwo_compact_expand": (_I, [_V, C.POINTER(StwoConfig), _V, _V, _SZ, _V, _V, _I]),
    "ssym_stwo_compact_hints": (_I, [_V, C.POINTER(StwoConfig), _V, _SZ, _V, _I]),
    "ssym_stwo_compact_pack_hinted": (_I, [C.POINTER(StwoConfig), _V, _V, _SZ, _V, _SZ, _V]),
    "ssym_stwo_compact_pack_gpu": (_I, [_V, C.POINTER(StwoConfig), _V, _SZ, _V, _SZ, _V]),
    "ssym_stwo_verify_compact_batch": (_I, [_V, C.POINTER(StwoConfig), _V, _V, _SZ, _V, _V, _I]),
    "ssym_stark101_verify_batch": (_I, [_V, _V, _V, _SZ, _V, _V, _V, _I]),
    "ssym_stark101_verify_multi_batch": (_I, [_V, _V, _V, _SZ, _U32, _V, _V, _V, _I]),
    "ssym_stwo_prove_batch": (_I, [_V, C.POINTER(StwoConfig), _V, _SZ, _V, _I]),
    "ssym_m31_add": (_I, [_V, _V, _V, _V, _SZ, _I]),
    "ssym_m31_sub": (_I, [_V, _V, _V, _V, _SZ, _I]),
    "ssym_m31_neg": (_I, [_V, _V, _V, _SZ, _I]),
    "ssym_m31_mul": (_I, [_V, _V, _V, _V, _SZ, _I]),
    "ssym_m31_inv": (_I, [_V, _V, _V, _V, _SZ, _I]),
    "ssym_cm31_mul": (_I, [_V, _V, _V, _V, _SZ, _I]),
    "ssym_cm31_inv": (_I, [_V, _V, _V, _V, _SZ, _I]),
    "ssym_qm31_add": (_I, [_V, _V, _V, _V, _SZ, _I]),
    "ssym_qm31_sub": (_I, [_V, _V, _V, _V, _SZ, _I]),
    "ssym_qm31_mul": (_I, [_V, _V, _V, _V, _SZ, _I]),
    "ssym_qm31_inv": (_I, [_V, _V, _V, _V, _SZ, _I]),
    "ssym_qm31_mul_m31": (_I, [_V, _V, _V, _V, _SZ, _I]),
    "ssym_qm31_mul_cm31": (_I, [_V, _V, _V, _V, _SZ, _I]),
    "ssym_circle_point": (_I, [_V, _V, _V, _SZ, _I]),
    "ssym_circle_fold": (_I, [_V, _V, _V, _V, _V, _U32, _V, _V, _SZ, _I]),
    "ssym_line_fold": (_I, [_V, _V, _V, _V, _V, _U32, _V, _V, _SZ, _I]),
    "ssym_sha256_pair": (_I, [_V, _V, _V, _V, _SZ, _I]),
    "ssym_merkle_root_from_path": (_I, [_V, _V, _V, _V, _U32, _V, _V, _V, _V, _SZ, _I]),
    "ssym_channel_mix_u256": (_I, [_V, _V, _V, _SZ, _I]),
    "ssym_channel_mix_u64": (_I, [_V, _V, _V, _SZ, _I]),
    "ssym_channel_draw_qm31": (_I, [_V, _V, _V, _V, _SZ, _I]),
    "ssym_channel_draw_queries": (_I, [_V, _V, _U32, _U32, _V, _SZ, _I]),
    "ssym_s101_mul_mod": (_I, [_V, _V, _V, _V, _SZ, _I]),
    "ssym_s101_div_mod": (_I, [_V, _V, _V, _V, _V, _SZ, _I]),
    "ssym_int32_peak_probe": (_I, [_V, C.POINTER(C.c_double), C.POINTER(C.c_double)]),
    "ssym_stwo_pack_wit": (_I, [C.POINTER(StwoConfig), C.c_char_p, _SZ, _V, C.POINTER(_I)]),
    "ssym_s101_pack_wit": (_I, [C.c_char_p, _SZ, _V, C.POINTER(_SZ)]),
    "ssym_stwo_wit_skeleton": (_I, [C.POINTER(StwoConfig), _I, _V, C.POINTER(_SZ), _V, C.POINTER(_SZ)]),
    "ssym_stwo_pack_wit_batch": (_I, [_V, C.POINTER(StwoConfig), _V, _V, _SZ, _V, _V, _I]),
    "ssym_stwo_verify_wit_batch": (_I, [_V, C.POINTER(StwoConfig), _V, _V, _SZ, _V, _V, _V, _I]),
    "ssym_stark101_verify_wit_batch": (_I, [_V, _V, _V, _SZ, _V, _V, _V, _I]),
}

COST_FIELDS = ("sha_compressions", "sha_init", "sha_add_4", "sha_add_8", "sha_add_32", "sha_finalize", "sha_bytes", "m31_mul", "m31_add", "m31_neg",
               "m31_inv", "eq_256", "point_from_index", "draw_retries")  # ssym_cost_t, include/ssym.h


class Cost(C.Structure):
    _fields_ = [(name, C.c_uint64) for name in COST_FIELDS]

    def as_dict(self):
        return {name: int(getattr(self, name)) for name in COST_FIELDS}


SYMBOLS["ssym_stwo_cost"] = (_I, [C.POINTER(StwoConfig), _V, _U32, _U32, C.POINTER(Cost)])

WIT_OK, WIT_SHAPE, WIT_PARSE = 0, 1, 2  # per-witness ingestion flags (include/ssym.h)

_lib = None


def load(build_if_missing: bool = True) -> C.CDLL:
    """Load libssym.so (building it in-tree with nvcc if it is missing)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        if not build_if_missing:
            raise SsymError(f"{LIB_PATH} is missing: run `python stark-symphony_b200/build.py` (no CPU fallback exists)")
        import importlib.util

        spec = importlib.util.spec_from_file_location("_ssym_build", os.path.join(HERE, "build.py"))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        mod.build()
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)  # AttributeError here = the library does not export what the header declares
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc: int) -> None:
    if rc != 0:
        msg = load().ssym_last_error().decode(errors="replace")
        raise SsymError(f"libssym error {rc}: {msg}")
