"""ctypes binding of libsemgate.so (C ABI in include/semgate.h) + a thin torch-facing
wrapper.  torch is used for device memory and streams only; all arithmetic on the
path happens in the library's sm_100a kernels.  There is no CPU fallback: if the
library is missing or the device is not a B200-class GPU, calls raise.
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass
from typing import Optional

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libsemgate.so")

MAX_K = 64            # candidates per query one sweep keeps
MAX_K_TOTAL = 1024    # largest k: above MAX_K the library runs ceil(k / 64) sweeps
FLOOR_NONE = -2**31
DTYPE_F32, DTYPE_F16, DTYPE_BF16 = 0, 1, 2     # SEMGATE_DTYPE_* (semgate.h): element type of descriptors handed to K1
GATE_FLAG, GATE_MASK = 0, 1
EINVAL, EARCH, ENOMEM, EDRIVER, EINDEX = -1, -2, -3, -4, -5

# every symbol include/semgate.h declares (checked by tests/test_abi.py)
SYMBOLS = [
    "semgate_version", "semgate_last_error", "semgate_create", "semgate_destroy", "semgate_device_info",
    "semgate_set_option", "semgate_profile_read", "semgate_launch_count", "semgate_pad_dim", "semgate_normalize_cast", "semgate_normalize_cast_dtype",
    "semgate_find_loop_closures_host_dtype",
    "semgate_topk_workspace_bytes", "semgate_gated_topk", "semgate_merge_topk", "semgate_compact_workspace_bytes",
    "semgate_compact", "semgate_gate_candidates", "semgate_find_loop_closures_host", "semgate_query_host",
    "semgate_gate_candidates_host", "semgate_spatial_workspace_bytes", "semgate_spatial_count", "semgate_spatial_fill",
    "semgate_spatial_candidates_host", "semgate_rerank_scores", "semgate_rerank_select", "semgate_similarity_matrix",
    "semgate_merge_topk_peers", "semgate_compact_valid", "semgate_stats_workspace_bytes", "semgate_candidate_stats",
    "semgate_last_sweep_mode", "semgate_schedule_check", "semgate_last_sweep_overflow", "semgate_merge_topk_peers_rows", "semgate_compact_rows", "semgate_clock_probe_read",
    "semgate_find_loop_closures_device", "semgate_find_loop_closures_device_workspace_bytes",
]


class SemgateError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(f"libsemgate error {code}: {message}")
        self.code = code


class TopkParams(C.Structure):
    _fields_ = [
        ("similarity_threshold", C.c_float),
        ("min_time_gap", C.c_double),
        ("k", C.c_int32),
        ("max_floor_diff", C.c_int32),
        ("gate_mode", C.c_int32),
        ("db_index_offset", C.c_uint32),
        ("cta_group", C.c_int32),
        ("accumulate", C.c_int32),
        ("symmetric", C.c_int32),
        ("part_index", C.c_int32),
        ("part_count", C.c_int32),
    ]


_lib = None


def load_library():
    """Load libsemgate.so; raises ImportError with the build command if absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.isfile(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} not found. Build it with `python multi-level-indoor-slam_b200/build.py` "
            "(nvcc, sm_100a). There is no fallback implementation.")
    lib = C.CDLL(LIB_PATH)
    vp, i32, i64, u32, sz = C.c_void_p, C.c_int32, C.c_int64, C.c_uint32, C.c_size_t
    P = C.POINTER
    lib.semgate_version.restype = C.c_int
    lib.semgate_last_error.restype = C.c_char_p
    lib.semgate_create.argtypes = [P(vp), C.c_int]
    lib.semgate_destroy.argtypes = [vp]
    lib.semgate_device_info.argtypes = [vp, P(C.c_int), P(C.c_int), P(C.c_int)]
    lib.semgate_set_option.argtypes = [vp, C.c_char_p, i64]
    lib.semgate_profile_read.argtypes = [vp, P(C.c_double), P(i64)]
    lib.semgate_launch_count.argtypes = [vp]
    lib.semgate_find_loop_closures_device_workspace_bytes.argtypes = [vp, i64, i32, P(TopkParams)]
    lib.semgate_find_loop_closures_device_workspace_bytes.restype = sz
    lib.semgate_find_loop_closures_device.argtypes = [vp, vp, i64, i32, vp, vp, P(TopkParams), vp, sz, vp, vp, vp, vp, vp, i32, vp]
    lib.semgate_clock_probe_read.argtypes = [vp, P(C.c_double), P(C.c_double), P(C.c_double), P(i32)]
    lib.semgate_launch_count.restype = i64
    lib.semgate_pad_dim.argtypes = [C.c_int]
    lib.semgate_normalize_cast.argtypes = [vp, vp, i64, i32, i64, vp, i32, vp]
    lib.semgate_normalize_cast_dtype.argtypes = [vp, vp, i32, i64, i32, i64, vp, i32, vp]
    lib.semgate_topk_workspace_bytes.argtypes = [vp, i64, i64, i32, P(TopkParams)]
    lib.semgate_topk_workspace_bytes.restype = sz
    lib.semgate_gated_topk.argtypes = [vp, vp, i64, vp, i64, i32, vp, vp, vp, vp, P(TopkParams), vp, sz,
                                       vp, vp, vp, vp, vp, vp]
    lib.semgate_merge_topk.argtypes = [vp, vp, i32, i64, i32, vp, vp, i32, vp, vp, vp, vp, vp, vp]
    lib.semgate_compact_workspace_bytes.argtypes = [i64]
    lib.semgate_compact_workspace_bytes.restype = sz
    lib.semgate_compact.argtypes = [vp, vp, vp, vp, vp, i64, i32, vp, vp, vp, vp, vp, vp, vp]
    lib.semgate_gate_candidates.argtypes = [vp, vp, i64, vp, vp, i64, i32, vp, vp, vp]
    lib.semgate_find_loop_closures_host.argtypes = [vp, vp, i64, i32, vp, vp, P(TopkParams), vp, vp, vp, vp, i64, P(i64)]
    lib.semgate_find_loop_closures_host_dtype.argtypes = [vp, vp, i32, i64, i32, vp, vp, P(TopkParams), vp, vp, vp, vp, i64, P(i64)]
    lib.semgate_query_host.argtypes = [vp, vp, i64, vp, i64, i32, vp, vp, P(TopkParams), vp, vp, vp]
    lib.semgate_gate_candidates_host.argtypes = [vp, vp, i64, vp, vp, i64, i32, vp, vp]
    lib.semgate_spatial_workspace_bytes.argtypes = [i64]
    lib.semgate_spatial_workspace_bytes.restype = sz
    lib.semgate_spatial_count.argtypes = [vp, vp, i64, C.c_double, i64, vp, vp, vp]
    lib.semgate_spatial_fill.argtypes = [vp, vp, i64, C.c_double, i64, vp, vp, vp, vp, i64, vp]
    lib.semgate_spatial_candidates_host.argtypes = [vp, vp, i64, C.c_double, i64, vp, vp, vp, i64, P(i64)]
    lib.semgate_rerank_scores.argtypes = [vp, vp, i64, i32, i32, vp, vp, vp, i64, vp, vp, vp]
    lib.semgate_rerank_select.argtypes = [vp, vp, vp, vp, i64, i32, i32, vp, vp, vp, vp]
    lib.semgate_similarity_matrix.argtypes = [vp, vp, i64, vp, i64, i32, vp, i64, vp]
    lib.semgate_merge_topk_peers.argtypes = [vp, vp, i32, i64, i32, vp, vp, i32, vp, vp, vp, vp, vp, vp]
    lib.semgate_merge_topk_peers_rows.argtypes = [vp, vp, i32, i64, i32, i64, i64, i64, vp, vp, i32, vp, vp, vp, vp, vp, vp, vp]
    lib.semgate_compact_rows.argtypes = [vp, vp, vp, vp, vp, i64, i32, i64, i32, vp, vp, vp, vp, vp, vp, vp]
    lib.semgate_compact_valid.argtypes = [vp, vp, vp, vp, vp, i64, i32, vp, vp, vp, vp, vp, vp, vp]
    lib.semgate_stats_workspace_bytes.argtypes = []
    lib.semgate_stats_workspace_bytes.restype = sz
    lib.semgate_candidate_stats.argtypes = [vp, vp, vp, vp, i64, vp, vp, vp]
    lib.semgate_last_sweep_mode.argtypes = [vp, P(i32), P(i64)]
    lib.semgate_last_sweep_overflow.argtypes = [vp, vp, vp]
    lib.semgate_schedule_check.argtypes = [i64, i64, i32, i32, i32, i32, i32, i32, P(i32), P(i64)]
    for name in SYMBOLS:
        getattr(lib, name)   # AttributeError here = the library is older than the header
    _lib = lib
    return lib


def _check(rc: int):
    if rc != 0:
        raise SemgateError(rc, load_library().semgate_last_error().decode("utf-8", "replace"))


def pad_dim(d: int) -> int:
    return ((int(d) + 63) // 64) * 64


def make_params(k: int, similarity_threshold: float = -np.inf, min_time_gap: float = 10.0, max_floor_diff: int = -1,
                gate_mode: int = GATE_FLAG, db_index_offset: int = 0, cta_group: int = 0,
                accumulate: bool = False, symmetric: int = 0, part_index: int = 0, part_count: int = 0) -> TopkParams:
    if not (1 <= int(k) <= MAX_K_TOTAL):
        raise ValueError(f"k={k} outside 1..{MAX_K_TOTAL}")
    # the reference compares `sim < threshold` in the similarity dtype (fp32): same rounding here
    thr = float(np.float32(similarity_threshold))
    return TopkParams(thr, float(min_time_gap), int(k), int(max_floor_diff), int(gate_mode), int(db_index_offset),
                      int(cta_group), 1 if accumulate else 0, int(symmetric), int(part_index), int(part_count))


def schedule_check(Q: int, N: int, d_pad: int, cta_group: int = 2, sm_count: int = 148, symmetric: bool = False,
                   part_index: int = 0, part_count: int = 1):
    """Host-side walk of the fused kernel's tile schedule (no device needed); raises if an invariant breaks.
    Returns dict(blocks, tiles, rm, s_main, r_last, s_last, window, resident, computed, makespan)."""
    shape, tiles = (C.c_int32 * 8)(), (C.c_int64 * 2)()
    _check(load_library().semgate_schedule_check(int(Q), int(N), int(d_pad), int(cta_group), int(sm_count),
                                                 int(symmetric), int(part_index), int(part_count), shape, tiles))
    names = ("blocks", "tiles", "rm", "s_main", "r_last", "s_last", "window", "resident")
    out = dict(zip(names, list(shape)))
    out["computed"], out["makespan"] = int(tiles[0]), int(tiles[1])
    return out


def _np_ptr(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


@dataclass
class TopkResult:
    scores: "object"   # [Q,k] fp32, descending, -inf padded
    idx: "object"      # [Q,k] int32 global database index, -1 padded
    valid: "object"    # [Q,k] uint8 floor flag
    count: "object"    # [Q] int32
    keys: "object" = None


class Engine:
    """One handle per GPU.  Device-pointer calls take torch CUDA tensors and run on
    torch's current stream; `*_host` calls take numpy arrays."""

    def __init__(self, device: int = 0):
        self.lib = load_library()
        self.device = int(device)
        h = C.c_void_p()
        _check(self.lib.semgate_create(C.byref(h), self.device))
        self._h = h
        sm, ma, mi = C.c_int(), C.c_int(), C.c_int()
        _check(self.lib.semgate_device_info(self._h, C.byref(sm), C.byref(ma), C.byref(mi)))
        self.sm_count, self.cc = sm.value, (ma.value, mi.value)
        self._ws = None   # grow-only torch workspace

    def close(self):
        if getattr(self, "_h", None):
            self.lib.semgate_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ------------------------------------------------------------------ helpers
    def set_option(self, name: str, value: int):
        _check(self.lib.semgate_set_option(self._h, name.encode(), int(value)))

    def profile_read(self):
        """(total K2 milliseconds, launches) since the last read; needs set_option('profile', 1)."""
        ms, n = C.c_double(0.0), C.c_int64(0)
        _check(self.lib.semgate_profile_read(self._h, C.byref(ms), C.byref(n)))
        return ms.value, n.value

    def clock_probe_read(self):
        """(median SM MHz, min SM MHz, span in us, CTAs) of the last sweep's K2 launch, from in-kernel clock64 /
        globaltimer pairs; needs set_option('clock_probe', 1).  Synchronises."""
        med, mn, span, n = C.c_double(0.0), C.c_double(0.0), C.c_double(0.0), C.c_int32(0)
        _check(self.lib.semgate_clock_probe_read(self._h, C.byref(med), C.byref(mn), C.byref(span), C.byref(n)))
        return med.value, mn.value, span.value, n.value

    def last_sweep_mode(self):
        """(mode, tiles) of the last gated_topk: 0 full sweep, 1 symmetric, 2 symmetric overflowed -> full redone."""
        mode, tiles = C.c_int32(0), C.c_int64(0)
        _check(self.lib.semgate_last_sweep_mode(self._h, C.byref(mode), C.byref(tiles)))
        return mode.value, tiles.value

    def last_sweep_overflow(self, out=None):
        """int32 device tensor [1]: non-zero iff the last (part of a) symmetric sweep overflowed its candidate
        buffers.  Stream-ordered, no host synchronisation."""
        torch = self._torch()
        if out is None:
            out = torch.empty((1,), dtype=torch.int32, device=self._dev())
        _check(self.lib.semgate_last_sweep_overflow(self._h, self._ptr(out), self._stream()))
        return out

    @property
    def launch_count(self) -> int:
        return int(self.lib.semgate_launch_count(self._h))

    def _torch(self):
        import torch
        return torch

    def _stream(self):
        torch = self._torch()
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def _dev(self):
        return self._torch().device("cuda", self.device)

    @staticmethod
    def _ptr(t):
        return None if t is None else C.c_void_p(t.data_ptr())

    def _expect(self, t, dtype, name, ndim=None):
        torch = self._torch()
        if t is None:
            return
        if not (isinstance(t, torch.Tensor) and t.is_cuda and t.device.index == self.device):
            raise TypeError(f"{name}: expected a CUDA tensor on device {self.device}")
        if t.dtype != dtype:
            raise TypeError(f"{name}: expected dtype {dtype}, got {t.dtype}")
        if not t.is_contiguous():
            raise ValueError(f"{name}: must be contiguous")
        if ndim is not None and t.dim() != ndim:
            raise ValueError(f"{name}: expected {ndim} dimensions")

    def _workspace(self, nbytes: int):
        torch = self._torch()
        if self._ws is None or self._ws.numel() < nbytes:
            self._ws = torch.empty(max(int(nbytes * 1.25), 1 << 20), dtype=torch.uint8, device=self._dev())
        return self._ws

    # ------------------------------------------------------------------ K1
    def normalize_cast(self, x, out=None):
        """fp32 (or fp16 / bf16) [n,d] CUDA tensor -> row-normalised bf16 [n,pad64(d)] (zero padded).  Half-precision
        rows are widened exactly: the result equals that of their fp32 image bit for bit."""
        torch = self._torch()
        dtypes = {torch.float32: DTYPE_F32, torch.float16: DTYPE_F16, torch.bfloat16: DTYPE_BF16}
        if x.dim() != 2 or x.dtype not in dtypes or not x.is_cuda or x.stride(1) != 1:
            raise TypeError("normalize_cast: expected a 2-D fp32 / fp16 / bf16 CUDA tensor with unit inner stride")
        n, d = x.shape
        dp = pad_dim(d)
        if out is None:
            out = torch.empty((n, dp), dtype=torch.bfloat16, device=x.device)
        self._expect(out, torch.bfloat16, "out", 2)
        if out.shape[0] < n or out.shape[1] != dp:
            raise ValueError("normalize_cast: out has the wrong shape")
        if n:
            _check(self.lib.semgate_normalize_cast_dtype(self._h, self._ptr(x), dtypes[x.dtype], n, d, x.stride(0), self._ptr(out),
                                                         dp, self._stream()))
        return out

    # ------------------------------------------------------------------ K2 + K3
    def gated_topk(self, q_bf16, db_bf16, params: TopkParams, q_ts=None, db_ts=None, q_floor=None, db_floor=None,
                   want_keys: bool = False, want_lists: bool = True, keys=None) -> TopkResult:
        """`keys`: int64 [Q,k] to write the key lists into (and, with params.accumulate, to merge with)."""
        torch = self._torch()
        self._expect(q_bf16, torch.bfloat16, "q_bf16", 2)
        self._expect(db_bf16, torch.bfloat16, "db_bf16", 2)
        self._expect(q_ts, torch.float64, "q_ts", 1)
        self._expect(db_ts, torch.float64, "db_ts", 1)
        self._expect(q_floor, torch.int32, "q_floor", 1)
        self._expect(db_floor, torch.int32, "db_floor", 1)
        Q, dp = q_bf16.shape
        N = db_bf16.shape[0]
        if db_bf16.shape[1] != dp and N > 0:
            raise ValueError("gated_topk: query and database descriptor lengths differ")
        for t, n, name in ((q_ts, Q, "q_ts"), (q_floor, Q, "q_floor"), (db_ts, N, "db_ts"), (db_floor, N, "db_floor")):
            if t is not None and t.shape[0] < n:
                raise ValueError(f"gated_topk: {name} shorter than its matrix")
        k = params.k
        dev = q_bf16.device
        if keys is not None:
            self._expect(keys, torch.int64, "keys", 2)
            if tuple(keys.shape) != (Q, k):
                raise ValueError("gated_topk: keys must be [Q, k]")
        elif want_keys or params.accumulate:
            if params.accumulate:
                raise ValueError("gated_topk: accumulate needs the previous `keys`")
            keys = torch.empty((Q, k), dtype=torch.int64, device=dev)
        scores = idx = valid = count = None
        if want_lists:
            scores = torch.empty((Q, k), dtype=torch.float32, device=dev)
            idx = torch.empty((Q, k), dtype=torch.int32, device=dev)
            valid = torch.empty((Q, k), dtype=torch.uint8, device=dev)
            count = torch.empty((Q,), dtype=torch.int32, device=dev)
        if Q == 0:
            return TopkResult(scores, idx, valid, count, keys)
        wsb = int(self.lib.semgate_topk_workspace_bytes(self._h, Q, N, dp, C.byref(params)))
        ws = self._workspace(wsb)
        _check(self.lib.semgate_gated_topk(
            self._h, self._ptr(q_bf16), Q, self._ptr(db_bf16), N, dp, self._ptr(q_ts), self._ptr(db_ts),
            self._ptr(q_floor), self._ptr(db_floor), C.byref(params), self._ptr(ws), ws.numel(),
            self._ptr(keys), self._ptr(scores), self._ptr(idx), self._ptr(valid), self._ptr(count), self._stream()))
        return TopkResult(scores, idx, valid, count, keys)

    def similarity_matrix(self, q_bf16, db_bf16, out=None):
        """Dense fp32 [Q,N] similarity of normalised bf16 rows (interface parity with
        compute_all_pairwise_similarities / _compute_similarity; the retrieval path never forms it)."""
        torch = self._torch()
        self._expect(q_bf16, torch.bfloat16, "q_bf16", 2)
        self._expect(db_bf16, torch.bfloat16, "db_bf16", 2)
        Q, dp = q_bf16.shape
        N = db_bf16.shape[0]
        if N > 0 and db_bf16.shape[1] != dp:
            raise ValueError("similarity_matrix: query and database descriptor lengths differ")
        if out is None:
            out = torch.empty((Q, N), dtype=torch.float32, device=q_bf16.device)
        self._expect(out, torch.float32, "out", 2)
        if tuple(out.shape) != (Q, N):
            raise ValueError("similarity_matrix: out must be [Q, N]")
        if Q and N:
            _check(self.lib.semgate_similarity_matrix(self._h, self._ptr(q_bf16), Q, self._ptr(db_bf16), N, dp,
                                                      self._ptr(out), out.stride(0), self._stream()))
        return out

    def merge_topk(self, keys_gathered, k: int, q_floor=None, db_floor_all=None, max_floor_diff: int = -1,
                   want_keys: bool = False) -> TopkResult:
        """keys_gathered: int64 [G,Q,k] (all-gathered `keys` of per-shard sweeps)."""
        torch = self._torch()
        self._expect(keys_gathered, torch.int64, "keys_gathered", 3)
        G, Q, kk = keys_gathered.shape
        if kk != k:
            raise ValueError("merge_topk: k mismatch")
        dev = keys_gathered.device
        scores = torch.empty((Q, k), dtype=torch.float32, device=dev)
        idx = torch.empty((Q, k), dtype=torch.int32, device=dev)
        valid = torch.empty((Q, k), dtype=torch.uint8, device=dev)
        count = torch.empty((Q,), dtype=torch.int32, device=dev)
        keys = torch.empty((Q, k), dtype=torch.int64, device=dev) if want_keys else None
        if Q:
            _check(self.lib.semgate_merge_topk(self._h, self._ptr(keys_gathered), G, Q, k, self._ptr(q_floor),
                                               self._ptr(db_floor_all), max_floor_diff, self._ptr(keys), self._ptr(scores),
                                               self._ptr(idx), self._ptr(valid), self._ptr(count), self._stream()))
        return TopkResult(scores, idx, valid, count, keys)

    def merge_topk_peers(self, peer_ptrs_dev: int, G: int, Q: int, k: int, q_floor=None, db_floor_all=None,
                         max_floor_diff: int = -1, want_keys: bool = False) -> TopkResult:
        """`peer_ptrs_dev`: address of a device array of G pointers to the ranks' `[Q,k]` key buffers
        (this rank's own and the peers' mapped over NVLink); merged in place, no gathered copy."""
        torch = self._torch()
        dev = self._dev()
        scores = torch.empty((Q, k), dtype=torch.float32, device=dev)
        idx = torch.empty((Q, k), dtype=torch.int32, device=dev)
        valid = torch.empty((Q, k), dtype=torch.uint8, device=dev)
        count = torch.empty((Q,), dtype=torch.int32, device=dev)
        keys = torch.empty((Q, k), dtype=torch.int64, device=dev) if want_keys else None
        if Q:
            _check(self.lib.semgate_merge_topk_peers(self._h, C.c_void_p(int(peer_ptrs_dev)), G, Q, k, self._ptr(q_floor),
                                                     self._ptr(db_floor_all), max_floor_diff, self._ptr(keys),
                                                     self._ptr(scores), self._ptr(idx), self._ptr(valid), self._ptr(count),
                                                     self._stream()))
        return TopkResult(scores, idx, valid, count, keys)

    def merge_topk_peers_rows(self, peer_ptrs_dev: int, G: int, Q: int, k: int, row_begin: int, row_count: int,
                              q_floor=None, db_floor_all=None, max_floor_diff: int = -1, want_keys: bool = False,
                              flag_offset: int = 0, any_flag=None) -> TopkResult:
        """Rows [row_begin, row_begin + row_count) of the G per-GPU `[Q,k]` key buffers merged in place over NVLink
        (this rank's share of the exchange); `any_flag` (int32 [1] device tensor): receives the OR of the uint32
        words at `flag_offset` keys behind every peer's buffer base (the ranks' overflow flags)."""
        torch = self._torch()
        dev = self._dev()
        n = int(row_count)
        scores = torch.empty((n, k), dtype=torch.float32, device=dev)
        idx = torch.empty((n, k), dtype=torch.int32, device=dev)
        valid = torch.empty((n, k), dtype=torch.uint8, device=dev)
        count = torch.empty((n,), dtype=torch.int32, device=dev)
        keys = torch.empty((n, k), dtype=torch.int64, device=dev) if want_keys else None
        self._expect(any_flag, torch.int32, "any_flag", 1)
        _check(self.lib.semgate_merge_topk_peers_rows(self._h, C.c_void_p(int(peer_ptrs_dev)), G, Q, k, int(row_begin), n,
                                                      int(flag_offset), self._ptr(q_floor), self._ptr(db_floor_all),
                                                      max_floor_diff, self._ptr(keys), self._ptr(scores), self._ptr(idx),
                                                      self._ptr(valid), self._ptr(count), self._ptr(any_flag), self._stream()))
        return TopkResult(scores, idx, valid, count, keys)

    def find_loop_closures_device(self, x_bf16, params: TopkParams, ts=None, floor=None, use_graph: bool = True):
        """find_loop_closures over a device-resident normalised database in ONE library call (K2 + K3 + K4),
        replayed as a CUDA graph from the third call with the same arguments on.  Returns
        (query_idx, match_idx, similarity, is_valid, total) CUDA tensors OWNED BY THE ENGINE: the next call with
        the same shape overwrites them (that is what keeps the pointers, hence the graph, stable)."""
        torch = self._torch()
        self._expect(x_bf16, torch.bfloat16, "x_bf16", 2)
        self._expect(ts, torch.float64, "ts", 1)
        self._expect(floor, torch.int32, "floor", 1)
        n, dp = x_bf16.shape
        k = params.k
        dev = x_bf16.device
        bufs = getattr(self, "_sweep_bufs", None)
        if bufs is None:
            bufs = self._sweep_bufs = {}
        key = (n, dp, k, params.symmetric, params.cta_group)
        if key not in bufs:
            if len(bufs) >= 4:
                bufs.pop(next(iter(bufs)))
            cap = max(n * k, 1)
            wsb = int(self.lib.semgate_find_loop_closures_device_workspace_bytes(self._h, n, dp, C.byref(params)))
            bufs[key] = (torch.empty((max(wsb, 256),), dtype=torch.uint8, device=dev),
                         torch.empty((cap,), dtype=torch.int32, device=dev), torch.empty((cap,), dtype=torch.int32, device=dev),
                         torch.empty((cap,), dtype=torch.float32, device=dev), torch.empty((cap,), dtype=torch.uint8, device=dev),
                         torch.zeros((1,), dtype=torch.int64, device=dev))
        ws, oq, om, os_, ov, total = bufs[key]
        cur = torch.cuda.current_stream(self.device)
        side = None
        if use_graph and cur.cuda_stream == 0:
            # a capture cannot start on the legacy default stream: run on a stream of our own, ordered behind and
            # before the caller's
            side = getattr(self, "_graph_stream", None)
            if side is None:
                side = self._graph_stream = torch.cuda.Stream(device=self.device)
            side.wait_stream(cur)
        st = C.c_void_p(side.cuda_stream) if side is not None else self._stream()
        _check(self.lib.semgate_find_loop_closures_device(
            self._h, self._ptr(x_bf16), n, dp, self._ptr(ts), self._ptr(floor), C.byref(params), self._ptr(ws), ws.numel(),
            self._ptr(oq), self._ptr(om), self._ptr(os_), self._ptr(ov), self._ptr(total), 1 if use_graph else 0, st))
        if side is not None:
            cur.wait_stream(side)
        return oq, om, os_, ov, total

    # ------------------------------------------------------------------ K4
    def compact(self, res: TopkResult, valid_only: bool = False, query_offset: int = 0):
        """Padded lists -> (query_idx, match_idx, similarity, is_valid, total) CUDA tensors;
        the arrays have capacity Q*k, the first `total` entries are meaningful.
        `valid_only`: emit only the floor-consistent candidates (the verifier hand-off list).
        `query_offset`: the lists are rows query_offset.. of a larger sweep (a rank's share); emitted indices are global."""
        torch = self._torch()
        Q, k = res.scores.shape
        dev = res.scores.device
        cap = max(Q * k, 1)
        oq = torch.empty((cap,), dtype=torch.int32, device=dev)
        om = torch.empty((cap,), dtype=torch.int32, device=dev)
        os_ = torch.empty((cap,), dtype=torch.float32, device=dev)
        ov = torch.empty((cap,), dtype=torch.uint8, device=dev)
        total = torch.empty((1,), dtype=torch.int64, device=dev)     # always written by the kernels
        wsb = int(self.lib.semgate_compact_workspace_bytes(Q))
        ws = torch.empty((max(wsb, 256),), dtype=torch.uint8, device=dev)
        if query_offset:
            _check(self.lib.semgate_compact_rows(self._h, self._ptr(res.scores), self._ptr(res.idx), self._ptr(res.valid),
                                                 self._ptr(res.count), Q, k, int(query_offset), 1 if valid_only else 0,
                                                 self._ptr(oq), self._ptr(om), self._ptr(os_), self._ptr(ov), self._ptr(total),
                                                 self._ptr(ws), self._stream()))
            return oq, om, os_, ov, total
        fn = self.lib.semgate_compact_valid if valid_only else self.lib.semgate_compact
        _check(fn(self._h, self._ptr(res.scores), self._ptr(res.idx), self._ptr(res.valid),
                  self._ptr(res.count), Q, k, self._ptr(oq), self._ptr(om), self._ptr(os_),
                  self._ptr(ov), self._ptr(total), self._ptr(ws), self._stream()))
        return oq, om, os_, ov, total

    def candidate_stats(self, similarity, is_valid, total):
        """Device-side get_statistics: float64[4] CUDA tensor = (total, valid, sum(sim), sum(valid sim)).
        `total`: the device int64 tensor `compact` returned, or a Python int."""
        torch = self._torch()
        self._expect(similarity, torch.float32, "similarity", 1)
        self._expect(is_valid, torch.uint8, "is_valid", 1)
        dev = similarity.device
        out = torch.empty((4,), dtype=torch.float64, device=dev)
        ws = torch.empty((int(self.lib.semgate_stats_workspace_bytes()),), dtype=torch.uint8, device=dev)
        if isinstance(total, int):
            tptr, m = None, total
        else:
            self._expect(total, torch.int64, "total", 1)
            tptr, m = self._ptr(total), 0
        _check(self.lib.semgate_candidate_stats(self._h, self._ptr(similarity), self._ptr(is_valid), tptr, m, self._ptr(ws),
                                                self._ptr(out), self._stream()))
        return out

    # ------------------------------------------------------------------ gate
    def gate_candidates(self, floor_labels, query_idx, match_idx, max_floor_diff: int):
        torch = self._torch()
        self._expect(floor_labels, torch.int32, "floor_labels", 1)
        self._expect(query_idx, torch.int32, "query_idx", 1)
        self._expect(match_idx, torch.int32, "match_idx", 1)
        M = query_idx.shape[0]
        if match_idx.shape[0] != M:
            raise ValueError("gate_candidates: index arrays differ in length")
        valid = torch.empty((M,), dtype=torch.uint8, device=query_idx.device)
        counts = torch.zeros((3,), dtype=torch.int64, device=query_idx.device)
        _check(self.lib.semgate_gate_candidates(self._h, self._ptr(floor_labels), floor_labels.shape[0], self._ptr(query_idx),
                                                self._ptr(match_idx), M, max_floor_diff, self._ptr(valid), self._ptr(counts),
                                                self._stream()))
        return valid, counts

    # ------------------------------------------------------------------ K5 re-rank
    def rerank_scores(self, local_feats, query_idx, match_idx, global_sim):
        """local_feats: bf16 [n_feat, P, dl_pad] (rows normalised); index / score vectors [M] on the device.
        Returns (cross[M], combined[M]) fp32; negative indices mean "no cached features"."""
        torch = self._torch()
        self._expect(local_feats, torch.bfloat16, "local_feats", 3)
        self._expect(query_idx, torch.int32, "query_idx", 1)
        self._expect(match_idx, torch.int32, "match_idx", 1)
        self._expect(global_sim, torch.float32, "global_sim", 1)
        M = query_idx.shape[0]
        if match_idx.shape[0] != M or global_sim.shape[0] != M:
            raise ValueError("rerank_scores: vectors differ in length")
        n_feat, P, dlp = local_feats.shape
        cross = torch.empty((M,), dtype=torch.float32, device=query_idx.device)
        comb = torch.empty((M,), dtype=torch.float32, device=query_idx.device)
        if M:
            _check(self.lib.semgate_rerank_scores(self._h, self._ptr(local_feats), n_feat, P, dlp, self._ptr(query_idx),
                                                  self._ptr(match_idx), self._ptr(global_sim), M, self._ptr(cross),
                                                  self._ptr(comb), self._stream()))
        return cross, comb

    def rerank_select(self, cand_idx, combined, count, top_k: int):
        """[Q,kc] padded candidate lists -> ([Q,top_k] idx, [Q,top_k] score, [Q] count), stable sort by score desc."""
        torch = self._torch()
        self._expect(cand_idx, torch.int32, "cand_idx", 2)
        self._expect(combined, torch.float32, "combined", 2)
        self._expect(count, torch.int32, "count", 1)
        Q, kc = cand_idx.shape
        dev = cand_idx.device
        oi = torch.empty((Q, top_k), dtype=torch.int32, device=dev)
        os_ = torch.empty((Q, top_k), dtype=torch.float32, device=dev)
        oc = torch.empty((Q,), dtype=torch.int32, device=dev)
        if Q:
            _check(self.lib.semgate_rerank_select(self._h, self._ptr(cand_idx), self._ptr(combined), self._ptr(count), Q, kc,
                                                  int(top_k), self._ptr(oi), self._ptr(os_), self._ptr(oc), self._stream()))
        return oi, os_, oc

    # ------------------------------------------------------------------ host-buffer calls
    def find_loop_closures_host(self, descriptors: np.ndarray, timestamps: Optional[np.ndarray],
                                floor_labels: Optional[np.ndarray], params: TopkParams, out=None):
        """Whole find_loop_closures on host arrays (H2D, kernels, D2H inside the call).
        Returns (query_idx, match_idx, similarity, is_valid) numpy arrays.  `descriptors`: fp32, or half precision as
        a numpy float16 array or a CPU torch tensor (float16 / bfloat16; numpy has no bf16) -- half the PCIe bytes,
        candidates bit-identical to the call on the widened rows."""
        dtype = DTYPE_F32
        if hasattr(descriptors, "data_ptr"):                   # a CPU torch tensor
            torch = self._torch()
            if descriptors.is_cuda:
                raise TypeError("find_loop_closures_host: descriptors live on the GPU; use normalize_cast + gated_topk")
            if descriptors.dtype in (torch.float16, torch.bfloat16):
                dtype = DTYPE_F16 if descriptors.dtype == torch.float16 else DTYPE_BF16
                keep = descriptors.contiguous()                # keeps the storage alive over the call
                desc = keep.view(torch.int16).numpy()
            else:
                desc = np.ascontiguousarray(descriptors.numpy(), dtype=np.float32)
        elif isinstance(descriptors, np.ndarray) and descriptors.dtype == np.float16:
            dtype = DTYPE_F16
            desc = np.ascontiguousarray(descriptors)
        else:
            desc = np.ascontiguousarray(descriptors, dtype=np.float32)
        n, d = desc.shape if desc.ndim == 2 else (0, 0)
        ts = None if timestamps is None else np.ascontiguousarray(timestamps, dtype=np.float64)
        fl = None if floor_labels is None else np.ascontiguousarray(floor_labels, dtype=np.int32)
        cap = max(n * params.k, 1)
        if out is None:
            out = (np.empty(cap, np.int32), np.empty(cap, np.int32), np.empty(cap, np.float32), np.empty(cap, np.uint8))
        oq, om, os_, ov = out
        total = C.c_int64(0)
        _check(self.lib.semgate_find_loop_closures_host_dtype(self._h, _np_ptr(desc), dtype, n, d, _np_ptr(ts), _np_ptr(fl),
                                                              C.byref(params), _np_ptr(oq), _np_ptr(om), _np_ptr(os_),
                                                              _np_ptr(ov), min(cap, oq.shape[0]), C.byref(total)))
        t = total.value
        return oq[:t], om[:t], os_[:t], ov[:t].astype(bool)

    def query_host(self, queries: np.ndarray, database: np.ndarray, params: TopkParams, q_ts=None, db_ts=None):
        q = np.ascontiguousarray(np.atleast_2d(queries), dtype=np.float32)
        db = np.ascontiguousarray(database, dtype=np.float32)
        nq, d = q.shape
        n = db.shape[0] if db.ndim == 2 else 0
        qt = None if q_ts is None else np.ascontiguousarray(q_ts, dtype=np.float64)
        dt = None if db_ts is None else np.ascontiguousarray(db_ts, dtype=np.float64)
        k = params.k
        scores = np.empty((nq, k), np.float32)
        idx = np.empty((nq, k), np.int32)
        count = np.empty((nq,), np.int32)
        _check(self.lib.semgate_query_host(self._h, _np_ptr(q), nq, _np_ptr(db) if n else None, n, d, _np_ptr(qt), _np_ptr(dt),
                                           C.byref(params), _np_ptr(scores), _np_ptr(idx), _np_ptr(count)))
        return scores, idx, count

    def gate_candidates_host(self, floor_labels: np.ndarray, query_idx: np.ndarray, match_idx: np.ndarray,
                             max_floor_diff: int):
        fl = np.ascontiguousarray(floor_labels, dtype=np.int32)
        qi = np.ascontiguousarray(query_idx, dtype=np.int32)
        mi = np.ascontiguousarray(match_idx, dtype=np.int32)
        M = qi.shape[0]
        valid = np.empty((max(M, 1),), np.uint8)
        counts = np.zeros((3,), np.uint64)
        _check(self.lib.semgate_gate_candidates_host(self._h, _np_ptr(fl), fl.shape[0], _np_ptr(qi), _np_ptr(mi), M,
                                                     max_floor_diff, _np_ptr(valid), _np_ptr(counts)))
        return valid[:M].astype(bool), int(counts[0]), int(counts[1])


    def spatial_candidates_host(self, positions: np.ndarray, radius: float, min_index_gap: int, want_dist: bool = True):
        """Radius join over poses: (i, j, dist) with i < j, j - i >= gap, ||p_i - p_j|| <= radius, sorted by (i, j)."""
        pos = np.ascontiguousarray(positions, dtype=np.float64)
        if pos.ndim != 2 or pos.shape[1] != 3:
            raise ValueError("positions must be [n, 3]")
        n = pos.shape[0]
        total = C.c_int64(0)
        rc = self.lib.semgate_spatial_candidates_host(self._h, _np_ptr(pos), n, float(radius), int(min_index_gap),
                                                      None, None, None, 0, C.byref(total))
        if rc not in (0, ENOMEM):
            _check(rc)
        t = total.value
        oi, oj = np.empty(max(t, 1), np.int32), np.empty(max(t, 1), np.int32)
        od = np.empty(max(t, 1), np.float64) if want_dist else None
        if t:
            _check(self.lib.semgate_spatial_candidates_host(self._h, _np_ptr(pos), n, float(radius), int(min_index_gap),
                                                            _np_ptr(oi), _np_ptr(oj), _np_ptr(od), t, C.byref(total)))
        return oi[:t], oj[:t], (od[:t] if want_dist else None)


_engines = {}


def get_engine(device: int = 0) -> Engine:
    """Process-wide engine per device.  Raises if there is no usable B200."""
    device = int(device)
    if device not in _engines:
        _engines[device] = Engine(device)
    return _engines[device]
