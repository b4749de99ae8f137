"""Verifier: one handle per GPU over the C-ABI (include/ssym.h).

Inputs may be numpy arrays (host memory: the call performs the H2D / D2H copies) or CUDA tensors
(anything with `.data_ptr()` and `.is_cuda`, e.g. torch tensors: inputs already resident in HBM, the call
is asynchronous on the handle's stream).  Outputs live where the inputs live."""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence, Tuple

import numpy as np

from . import _lib
from ._lib import MEM_DEVICE, MEM_HOST, S101Trace, SsymError, StwoConfig, StwoLayout, StwoTrace, check, load


def stwo_config(preset: str = "prod", mode: int = _lib.MODE_REF_LITERAL, n_columns: int = 4, dedup_queries: bool = False) -> StwoConfig:
    """The two presets of stwo-verifier/src/config.simf:10-51; n_columns = NUM_COLUMNS (config.simf:14: 4 at reference HEAD; 8 and 16 widen
    the same wide-Fibonacci AIR); dedup_queries = the SSYM_MODE_QUERY_DEDUP flag (sorted distinct queries, fri/queries.simf:41)."""
    cfg = StwoConfig()
    if dedup_queries:
        mode |= _lib.MODE_QUERY_DEDUP
    check(load().ssym_stwo_config_preset(preset.encode(), mode, C.byref(cfg)))
    cfg.n_columns = n_columns
    return cfg


_layout_cache: dict = {}


def stwo_layout(cfg: StwoConfig) -> StwoLayout:
    """ssym_stwo_layout, memoised per configuration (a loop of calls on one configuration asks for it every time)."""
    key = bytes(cfg)
    lo = _layout_cache.get(key)
    if lo is None:
        lo = StwoLayout()
        check(load().ssym_stwo_layout(C.byref(cfg), C.byref(lo)))
        if len(_layout_cache) < 64:
            _layout_cache[key] = lo
    return lo


def _is_device(x) -> bool:
    return hasattr(x, "data_ptr") and bool(getattr(x, "is_cuda", False))


def _ptr(x):
    if x is None:
        return None
    if _is_device(x):
        return C.c_void_p(x.data_ptr())
    if isinstance(x, np.ndarray):
        if not x.flags["C_CONTIGUOUS"]:
            raise SsymError("host arrays must be C-contiguous")
        return C.c_void_p(x.ctypes.data)
    if hasattr(x, "data_ptr"):  # CPU torch tensor (e.g. pinned host memory)
        return C.c_void_p(x.data_ptr())
    raise SsymError(f"unsupported buffer type {type(x)}")


class Verifier:
    def __init__(self, device: int = 0):
        self.lib = load()
        h = C.c_void_p()
        check(self.lib.ssym_create(device, C.byref(h)))
        self.h = h
        self.device = device
        self._explicit_stream = False  # set_stream() was called: the caller manages ordering
        self._auto_stream = None       # the stream handle last installed by _space()

    def close(self) -> None:
        if getattr(self, "h", None):
            self.lib.ssym_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- plumbing ------------------------------------------------------------------------------------
    def set_stream(self, cuda_stream: Optional[int]) -> None:
        """Run device-resident calls on a caller-owned stream (e.g. torch.cuda.current_stream().cuda_stream)."""
        if cuda_stream == 0:
            raise SsymError("pass a non-default stream handle (e.g. torch.cuda.Stream().cuda_stream); None restores the default behaviour")
        check(self.lib.ssym_set_stream(self.h, C.c_void_p(cuda_stream) if cuda_stream else None))
        self._explicit_stream = cuda_stream is not None
        self._auto_stream = None

    def synchronize(self) -> None:
        check(self.lib.ssym_synchronize(self.h))

    def set_pipeline_depth(self, depth: int) -> None:
        """Keep up to `depth` device-resident stwo batches in flight (see ssym_set_pipeline_depth in include/ssym.h)."""
        check(self.lib.ssym_set_pipeline_depth(self.h, depth))

    def set_host_async(self, on: bool) -> None:
        """Host-buffer stwo_verify_batch calls only enqueue their copies + kernels (see ssym_set_host_async); outputs are valid
        after synchronize().  Buffers must be pinned."""
        check(self.lib.ssym_set_host_async(self.h, 1 if on else 0))

    def set_wit_host_fallback(self, on: bool) -> None:
        """on=False: `.wit` texts the GPU tokeniser does not keep on its fast path are flagged WIT_SLOW (= rejected) instead of being re-read
        by the host parser (ssym_set_wit_host_fallback, include/ssym.h)."""
        check(self.lib.ssym_set_wit_host_fallback(self.h, 1 if on else 0))

    def set_merkle_sharing(self, policy: int) -> None:
        """0: one hash chain per query (the reference's schedule); 2: paths of a tree share the nodes above their meeting point; 1 (default):
        2 under MODE_PROVER_CONSISTENT, 0 under MODE_REF_LITERAL (ssym_set_merkle_sharing, include/ssym.h)."""
        check(self.lib.ssym_set_merkle_sharing(self.h, int(policy)))

    def join(self) -> None:
        """Order all in-flight batches into the handle's stream (device-side wait only)."""
        check(self.lib.ssym_join(self.h))

    @property
    def launch_count(self) -> int:
        return int(self.lib.ssym_launch_count(self.h))

    KERNEL_NAMES = ("stwo_channel", "stwo_query", "stwo_merkle", "stwo_finalize", "s101_transcript", "s101_merkle", "s101_finalize", "other")

    def profile_enable(self, on: bool = True) -> None:
        """Record CUDA events around every verifier kernel on the launching stream."""
        check(self.lib.ssym_profile_enable(self.h, 1 if on else 0))

    def profile_read(self):
        """-> {kernel name: (total ms, launches)} since the last read (synchronises the stream)."""
        ms = (C.c_double * 8)()
        cnt = (C.c_uint64 * 8)()
        check(self.lib.ssym_profile_read(self.h, ms, cnt))
        return {self.KERNEL_NAMES[i]: (ms[i], int(cnt[i])) for i in range(8) if cnt[i]}

    def _alloc(self, like, shape, dtype=np.uint32):
        if _is_device(like):
            import torch

            tdt = {np.uint32: torch.int32, np.uint8: torch.uint8}[dtype]  # bit patterns; torch has no general uint32 ops
            return torch.empty(shape, dtype=tdt, device=like.device)  # every output element is written by the kernels
        return np.zeros(shape, dtype=dtype)

    CUDA_STREAM_LEGACY = 1  # cudaStreamLegacy: the handle that names the default stream in API calls

    def _space(self, x) -> int:
        """Memory space of a buffer.  For device buffers this is also where the stream-ordering contract of include/ssym.h ("Stream ordering")
        is made safe by default: unless set_stream() was called, the handle is pointed at torch's CURRENT stream for x's device, so the
        call is ordered after whatever produced its inputs and before whatever the caller enqueues next."""
        if not _is_device(x):
            return MEM_HOST
        if not self._explicit_stream:
            try:
                import torch

                handle = int(torch.cuda.current_stream(x.device).cuda_stream) or self.CUDA_STREAM_LEGACY
            except Exception:  # not a torch tensor: the caller owns the ordering (ssym_set_stream)
                handle = None
            if handle is not None and handle != self._auto_stream:
                check(self.lib.ssym_set_stream(self.h, C.c_void_p(handle)))
                self._auto_stream = handle
        return MEM_DEVICE

    # ---- whole proofs --------------------------------------------------------------------------------
    def stwo_verify_batch(self, packed, cfg: StwoConfig, n: Optional[int] = None, want_status: bool = False, want_trace: bool = False,
                          accept_out=None, status_out=None):
        """verify_proof (stwo-verifier/src/verifier.simf:32-58) for a batch of packed proofs.
        Returns (accept_bits, status | None, traces | None); bit i of accept_bits = proof i accepted."""
        lo = stwo_layout(cfg)
        total = packed.numel() if _is_device(packed) or hasattr(packed, "numel") else packed.size
        if n is None:
            if total % lo.stride_words:
                raise SsymError("packed length is not a multiple of the proof stride")
            n = total // lo.stride_words
        if n * lo.stride_words > total:
            raise SsymError("packed buffer too small")
        space = self._space(packed)
        accept = accept_out if accept_out is not None else self._alloc(packed, (n + 31) // 32)
        status = status_out if status_out is not None else (self._alloc(packed, n) if want_status else None)
        traces = None
        tptr = None
        if want_trace:
            if space == MEM_DEVICE:
                traces = self._alloc(packed, n * C.sizeof(StwoTrace), np.uint8)
                tptr = _ptr(traces)
            else:
                traces = (StwoTrace * n)()
                tptr = C.cast(traces, C.c_void_p)
        check(self.lib.ssym_stwo_verify_batch(self.h, C.byref(cfg), _ptr(packed), n, _ptr(accept), _ptr(status), tptr, space))
        return accept, status, traces

    def stwo_compact_expand(self, blob, offsets, cfg: StwoConfig, want_flags: bool = False):
        """Compact records (witness.compact_stwo) -> packed records, on the GPU (ssym_stwo_compact_expand).  Returns (packed (n, stride_words), flags | None)."""
        lo = stwo_layout(cfg)
        n = (offsets.numel() if hasattr(offsets, "numel") else offsets.size) - 1
        packed = self._alloc(blob, (n, lo.stride_words))
        flags = self._alloc(blob, n) if want_flags else None
        check(self.lib.ssym_stwo_compact_expand(self.h, C.byref(cfg), _ptr(blob), _ptr(offsets), n, _ptr(packed), _ptr(flags), self._space(blob)))
        return packed, flags

    def stwo_compact_hints(self, packed, cfg: StwoConfig, n: Optional[int] = None):
        """For every sibling slot of every packed proof: the lowest query whose Merkle path reaches a node equal to that sibling at that level,
        or 0xff (ssym_stwo_compact_hints; slot order = the compact record's).  Returns an (n, slots) uint8 array / tensor."""
        lo = stwo_layout(cfg)
        total = packed.numel() if hasattr(packed, "numel") else packed.size
        n = total // lo.stride_words if n is None else n
        Q, L, G = cfg.n_queries, cfg.n_fri_layers, cfg.lde_log
        slots = Q * (2 * G + sum(G - 1 - l for l in range(L + 1)))
        hints = self._alloc(packed, (n, slots), np.uint8)
        check(self.lib.ssym_stwo_compact_hints(self.h, C.byref(cfg), _ptr(packed), n, _ptr(hints), self._space(packed)))
        return hints

    def stwo_verify_compact_batch(self, blob, offsets, cfg: StwoConfig, want_status: bool = False, accept_out=None, status_out=None):
        """stwo_verify_batch on compact records (include/ssym.h "compact transport form"): with host arrays, the compact bytes are what
        crosses the host link.  Returns (accept_bits, status | None)."""
        n = (offsets.numel() if hasattr(offsets, "numel") else offsets.size) - 1
        accept = accept_out if accept_out is not None else self._alloc(blob, (n + 31) // 32)
        status = status_out if status_out is not None else (self._alloc(blob, n) if want_status else None)
        check(self.lib.ssym_stwo_verify_compact_batch(self.h, C.byref(cfg), _ptr(blob), _ptr(offsets), n, _ptr(accept), _ptr(status), self._space(blob)))
        return accept, status

    def stwo_prove_batch(self, seeds, cfg: StwoConfig, out=None):
        """Batched prover for the wide-Fibonacci AIR that verify_proof checks (ssym_stwo_prove_batch, include/ssym.h): one packed
        proof per u64 seed, accepted by stwo_verify_batch in MODE_PROVER_CONSISTENT.  `seeds`: numpy uint64 array (host) or a CUDA
        int64 tensor (device; the proofs then stay in HBM).  Returns an (n, stride_words) array / tensor."""
        lo = stwo_layout(cfg)
        if _is_device(seeds):
            import torch

            n = seeds.numel()
            if seeds.dtype != torch.int64 or not seeds.is_contiguous():
                raise SsymError("device seeds must be a contiguous int64 tensor (the u64 bit patterns)")
            if out is None:
                out = torch.empty((n, lo.stride_words), dtype=torch.int32, device=seeds.device)
            space = self._space(seeds)
        else:
            seeds = np.ascontiguousarray(seeds, dtype=np.uint64)
            n = seeds.size
            if out is None:
                out = np.zeros((n, lo.stride_words), dtype=np.uint32)
            space = MEM_HOST
        check(self.lib.ssym_stwo_prove_batch(self.h, C.byref(cfg), _ptr(seeds), n, _ptr(out), space))
        return out

    def stark101_verify_batch(self, blob, offsets, want_status: bool = False, want_trace: bool = False):
        """verify_proof (stark101/src/verifier.simf:24-42) for a batch of packed records."""
        n = (offsets.numel() if hasattr(offsets, "numel") else offsets.size) - 1
        space = self._space(blob)
        accept = self._alloc(blob, (n + 31) // 32)
        status = self._alloc(blob, n) if want_status else None
        traces, tptr = None, None
        if want_trace:
            if space == MEM_DEVICE:
                traces = self._alloc(blob, n * C.sizeof(S101Trace), np.uint8)
                tptr = _ptr(traces)
            else:
                traces = (S101Trace * n)()
                tptr = C.cast(traces, C.c_void_p)
        check(self.lib.ssym_stark101_verify_batch(self.h, _ptr(blob), _ptr(offsets), n, _ptr(accept), _ptr(status), tptr, space))
        return accept, status, traces

    def stark101_verify_multi_batch(self, blob, offsets, n_queries: int, want_trace: bool = False):
        """Multi-query stark101 (ssym_stark101_verify_multi_batch, include/ssym.h): records proof-major, `n_queries` records of the reference's
        witness shape per proof, record k verified under the (k+1)-th query draw.  Returns (accept bits per PROOF, status per RECORD, traces)."""
        n = (offsets.numel() if hasattr(offsets, "numel") else offsets.size) - 1
        if n_queries < 1 or n % n_queries:
            raise SsymError("the number of records must be a multiple of n_queries")
        n_proofs = n // n_queries
        space = self._space(blob)
        accept = self._alloc(blob, (n_proofs + 31) // 32)
        status = self._alloc(blob, n)
        traces, tptr = None, None
        if want_trace:
            if space == MEM_DEVICE:
                traces = self._alloc(blob, n * C.sizeof(S101Trace), np.uint8)
                tptr = _ptr(traces)
            else:
                traces = (S101Trace * n)()
                tptr = C.cast(traces, C.c_void_p)
        check(self.lib.ssym_stark101_verify_multi_batch(self.h, _ptr(blob), _ptr(offsets), n_proofs, n_queries, _ptr(accept), _ptr(status), tptr, space))
        return accept, status, traces

    # ---- `.wit` text in (GPU tokeniser, csrc/wit_kernels.cu) ---------------------------------------------
    def stwo_pack_wit_batch(self, text, offsets, cfg: StwoConfig):
        """n `.wit` JSON texts (concatenated in `text`, witness i = bytes [offsets[i], offsets[i+1])) -> (packed (n, stride_words),
        flags (n)) tokenised and packed on the GPU; flags: WIT_OK / WIT_SHAPE / WIT_PARSE (include/ssym.h).  `text`: numpy uint8
        array (host) or CUDA uint8 tensor; `offsets`: uint64 array / int64 tensor in the same memory space."""
        lo = stwo_layout(cfg)
        n = (offsets.numel() if hasattr(offsets, "numel") else offsets.size) - 1
        space = self._space(text)
        if space == MEM_DEVICE:
            import torch

            packed = torch.empty((n, lo.stride_words), dtype=torch.int32, device=text.device)
            flags = torch.empty(n, dtype=torch.int32, device=text.device)
        else:
            packed = np.zeros((n, lo.stride_words), dtype=np.uint32)
            flags = np.zeros(n, dtype=np.uint32)
        check(self.lib.ssym_stwo_pack_wit_batch(self.h, C.byref(cfg), _ptr(text), _ptr(offsets), n, _ptr(packed), _ptr(flags), space))
        return packed, flags

    def stwo_verify_wit_batch(self, text, offsets, cfg: StwoConfig, want_status: bool = False, want_flags: bool = False, accept_out=None):
        """verify_proof for n witness TEXTS: tokenise + pack + verify on the GPU, the packed proofs never leave it.
        Returns (accept_bits, status | None, flags | None)."""
        n = (offsets.numel() if hasattr(offsets, "numel") else offsets.size) - 1
        space = self._space(text)
        accept = accept_out if accept_out is not None else self._alloc(text, (n + 31) // 32)
        status = self._alloc(text, n) if want_status else None
        flags = self._alloc(text, n) if want_flags else None
        check(self.lib.ssym_stwo_verify_wit_batch(self.h, C.byref(cfg), _ptr(text), _ptr(offsets), n, _ptr(accept), _ptr(status), _ptr(flags), space))
        return accept, status, flags

    def stark101_verify_wit_batch(self, text, offsets, want_status: bool = False, want_flags: bool = False):
        """verify_proof (stark101) for n witness TEXTS, tokenised and packed on the GPU (ssym_stark101_verify_wit_batch)."""
        n = (offsets.numel() if hasattr(offsets, "numel") else offsets.size) - 1
        space = self._space(text)
        accept = self._alloc(text, (n + 31) // 32)
        status = self._alloc(text, n) if want_status else None
        flags = self._alloc(text, n) if want_flags else None
        check(self.lib.ssym_stark101_verify_wit_batch(self.h, _ptr(text), _ptr(offsets), n, _ptr(accept), _ptr(status), _ptr(flags), space))
        return accept, status, flags

    # ---- `simfony run`-shaped convenience ---------------------------------------------------------------
    def run_stwo_wit(self, wit_texts: Sequence[str], preset: str = "prod", mode: int = _lib.MODE_REF_LITERAL):
        """`simfony run main.simf --witness x.wit` for many witnesses: returns (accept: list[bool], status: np.ndarray)."""
        from . import witness

        cfg = stwo_config(preset, mode)
        text, offsets = witness.concat_wit_texts(wit_texts)
        _, status, _ = self.stwo_verify_wit_batch(text, offsets, cfg, want_status=True)
        return [bool(s == 0) for s in status], status

    def run_stark101_wit(self, wit_texts: Sequence[str]):
        from . import witness

        text, offsets = witness.concat_wit_texts(wit_texts)
        _, status, _ = self.stark101_verify_wit_batch(text, offsets, want_status=True)
        return [bool(s == 0) for s in status], status

    # ---- element-wise jets --------------------------------------------------------------------------------
    def _jet(self, name: str, words_out: int, a, b=None, with_fail: bool = False, n: Optional[int] = None, words_a: int = 1):
        total = a.numel() if hasattr(a, "numel") else a.size
        n = total // words_a if n is None else n
        out = self._alloc(a, n * words_out)
        failv = self._alloc(a, n, np.uint8) if with_fail else None
        fn = getattr(self.lib, name)
        args = [self.h, _ptr(a)] + ([_ptr(b)] if b is not None else []) + [_ptr(out)] + ([_ptr(failv)] if with_fail else []) + [n, self._space(a)]
        check(fn(*args))
        return (out, failv) if with_fail else out

    def m31_add(self, a, b): return self._jet("ssym_m31_add", 1, a, b)      # fields/m31.simf:22-26
    def m31_sub(self, a, b): return self._jet("ssym_m31_sub", 1, a, b)      # fields/m31.simf:35-37
    def m31_neg(self, a): return self._jet("ssym_m31_neg", 1, a)            # fields/m31.simf:29-32
    def m31_mul(self, a, b): return self._jet("ssym_m31_mul", 1, a, b)      # fields/m31.simf:40-45
    def m31_inv(self, a): return self._jet("ssym_m31_inv", 1, a, with_fail=True)              # fields/m31.simf:117-132
    def cm31_mul(self, a, b): return self._jet("ssym_cm31_mul", 2, a, b, words_a=2)           # fields/cm31.simf:79-86
    def cm31_inv(self, a): return self._jet("ssym_cm31_inv", 2, a, with_fail=True, words_a=2)  # fields/cm31.simf:88-93
    def qm31_add(self, a, b): return self._jet("ssym_qm31_add", 4, a, b, words_a=4)           # fields/qm31.simf:36-40
    def qm31_sub(self, a, b): return self._jet("ssym_qm31_sub", 4, a, b, words_a=4)           # fields/qm31.simf:49-53
    def qm31_mul(self, a, b): return self._jet("ssym_qm31_mul", 4, a, b, words_a=4)           # fields/qm31.simf:73-80
    def qm31_inv(self, a): return self._jet("ssym_qm31_inv", 4, a, with_fail=True, words_a=4)  # fields/qm31.simf:87-98
    def qm31_mul_m31(self, a, b): return self._jet("ssym_qm31_mul_m31", 4, a, b, words_a=4)   # fields/qm31.simf:56-59
    def qm31_mul_cm31(self, a, b): return self._jet("ssym_qm31_mul_cm31", 4, a, b, words_a=4)  # fields/qm31.simf:62-65
    def s101_mul_mod(self, a, b): return self._jet("ssym_s101_mul_mod", 1, a, b)              # stark101/src/field.simf:30-35
    def s101_div_mod(self, a, b): return self._jet("ssym_s101_div_mod", 1, a, b, with_fail=True)  # stark101/src/field.simf:42-66

    def circle_point(self, index):
        """circle_point_index_to_m31_point (groups/m31_point.simf:103-106) -> n x {x, y}."""
        n = index.numel() if hasattr(index, "numel") else index.size
        out = self._alloc(index, 2 * n)
        check(self.lib.ssym_circle_point(self.h, _ptr(index), _ptr(out), n, self._space(index)))
        return out

    def _fold(self, name, position, f_p, f_neg_p, alpha, log_size):
        n = position.numel() if hasattr(position, "numel") else position.size
        out = self._alloc(position, 4 * n)
        failv = self._alloc(position, n, np.uint8)
        check(getattr(self.lib, name)(self.h, _ptr(position), _ptr(f_p), _ptr(f_neg_p), _ptr(alpha), log_size, _ptr(out), _ptr(failv), n,
                                      self._space(position)))
        return out, failv

    def circle_fold(self, position, f_p, f_neg_p, alpha, log_size): return self._fold("ssym_circle_fold", position, f_p, f_neg_p, alpha, log_size)  # fri/folding.simf:15-27
    def line_fold(self, position, f_p, f_neg_p, alpha, log_size): return self._fold("ssym_line_fold", position, f_p, f_neg_p, alpha, log_size)      # fri/folding.simf:30-41

    def sha256_pair(self, left, right):
        """sha256_pair (hasher.simf:27-32) on n pairs of 8-word digests."""
        n = (left.numel() if hasattr(left, "numel") else left.size) // 8
        out = self._alloc(left, 8 * n)
        check(self.lib.ssym_sha256_pair(self.h, _ptr(left), _ptr(right), _ptr(out), n, self._space(left)))
        return out

    def merkle_root_from_path(self, leaf, auth_path, siblings, depth: int, expected_root=None):
        """merkle_verify_32 (merkle.simf:39-44) on n paths of equal depth -> (root, final_path, ok_bits)."""
        n = auth_path.numel() if hasattr(auth_path, "numel") else auth_path.size
        root = self._alloc(leaf, 8 * n)
        path = self._alloc(leaf, n)
        ok = self._alloc(leaf, (n + 31) // 32)
        check(self.lib.ssym_merkle_root_from_path(self.h, _ptr(leaf), _ptr(auth_path), _ptr(siblings), depth, _ptr(expected_root), _ptr(root),
                                                  _ptr(path), _ptr(ok), n, self._space(leaf)))
        return root, path, ok

    def channel_mix_u256(self, state, value):   # channel.simf:154-162 (in place)
        n = (state.numel() if hasattr(state, "numel") else state.size) // 9
        check(self.lib.ssym_channel_mix_u256(self.h, _ptr(state), _ptr(value), n, self._space(state)))
        return state

    def channel_mix_u64(self, state, hi_lo):    # channel.simf:165-173 (in place)
        n = (state.numel() if hasattr(state, "numel") else state.size) // 9
        check(self.lib.ssym_channel_mix_u64(self.h, _ptr(state), _ptr(hi_lo), n, self._space(state)))
        return state

    def channel_draw_qm31(self, state):         # channel.simf:115-141 (state advanced in place)
        n = (state.numel() if hasattr(state, "numel") else state.size) // 9
        out = self._alloc(state, 4 * n)
        failv = self._alloc(state, n, np.uint8)
        check(self.lib.ssym_channel_draw_qm31(self.h, _ptr(state), _ptr(out), _ptr(failv), n, self._space(state)))
        return out, failv

    def channel_draw_queries(self, state, log_size: int, n_queries: int):  # fri/queries.simf:14-43
        n = (state.numel() if hasattr(state, "numel") else state.size) // 9
        out = self._alloc(state, n * n_queries)
        check(self.lib.ssym_channel_draw_queries(self.h, _ptr(state), log_size, n_queries, _ptr(out), n, self._space(state)))
        return out

    def int32_peak_probe(self) -> Tuple[float, float]:
        """Measured 32-bit integer ALU throughput (machine-instruction lanes / s) and the probe's duration in ms."""
        ops, ms = C.c_double(0), C.c_double(0)
        check(self.lib.ssym_int32_peak_probe(self.h, C.byref(ops), C.byref(ms)))
        return ops.value, ms.value
