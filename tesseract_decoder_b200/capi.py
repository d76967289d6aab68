"""ctypes binding of the C ABI in include/tesseract_b200.h (libtesseract_b200.so, built in-tree).

This is the thin, torch-free Python face of the boundary: tests and bench.py call the decode path
through it so that what is measured and checked is exactly what a reference maintainer would bind.
There is no CPU fallback: decode calls raise CudaUnavailable without an sm_100 device.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libtesseract_b200.so")

INF_DET_BEAM = 65535
PQLIMIT_UNBOUNDED = (1 << 64) - 1
NO_ERROR_INDEX = (1 << 64) - 1
CREATE_HOST_ONLY = 1

TSR_E_INVALID_ARGUMENT, TSR_E_RUNTIME, TSR_E_CUDA, TSR_E_DEVICE_MEMORY = -1, -2, -3, -4


class CudaUnavailable(RuntimeError):
    pass


class DeviceMemoryError(RuntimeError):
    pass


class TsrConfig(C.Structure):
    _fields_ = [
        ("det_beam", C.c_int32),
        ("beam_climbing", C.c_int32),
        ("no_revisit_dets", C.c_int32),
        ("merge_errors", C.c_int32),
        ("at_most_two_errors_per_detector", C.c_int32),
        ("device", C.c_int32),
        ("pqlimit", C.c_uint64),
        ("det_penalty", C.c_double),
        ("num_det_orders", C.c_uint64),
        ("det_orders", C.POINTER(C.c_uint64)),
        ("det_order_len", C.c_uint64),
        ("search_slots", C.c_uint32),
        ("warps_per_block", C.c_uint32),
        ("node_pool_bytes", C.c_uint64),
        ("max_batch_shots", C.c_uint64),
        ("force_generic_kernel", C.c_int32),
        ("row_scan", C.c_int32),
        ("sparsify_errors", C.c_int32),
        ("sparsify_base_degree", C.c_int32),
        ("sparsify_max_degree", C.c_int32),
        ("sparsify_reactivate_limit", C.c_int32),
        ("create_visualization", C.c_int32),
    ]


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(make -C tesseract_decoder_b200/csrc).  There is no fallback implementation.")
    L = C.CDLL(LIB_PATH)
    u64p, f64p, u8p, i32p = (C.POINTER(t) for t in (C.c_uint64, C.c_double, C.c_uint8, C.c_int32))
    L.tsr_config_init.argtypes = [C.POINTER(TsrConfig)]
    L.tsr_last_error.restype = C.c_char_p
    L.tsr_version.restype = C.c_char_p
    L.tsr_create.argtypes = [C.c_char_p, C.POINTER(TsrConfig), C.c_uint32, C.POINTER(C.c_void_p)]
    L.tsr_destroy.argtypes = [C.c_void_p]
    for name in ("tsr_num_detectors", "tsr_num_observables", "tsr_num_errors", "tsr_num_dem_errors",
                 "tsr_num_det_orders", "tsr_last_batch_nnz"):
        getattr(L, name).restype = C.c_uint64
        getattr(L, name).argtypes = [C.c_void_p]
    L.tsr_error_costs.argtypes = [C.c_void_p, f64p]
    L.tsr_maps.argtypes = [C.c_void_p, u64p, u64p]
    L.tsr_csr.restype = C.c_int64
    L.tsr_csr.argtypes = [C.c_void_p, C.c_int, u64p, i32p]
    L.tsr_compiled_dem_text.argtypes = [C.c_void_p, C.POINTER(C.c_void_p)]
    L.tsr_free.argtypes = [C.c_void_p]
    L.tsr_decode_batch_b8.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p]
    L.tsr_decode_batch_b8_device.argtypes = L.tsr_decode_batch_b8.argtypes
    L.tsr_last_batch_errors.argtypes = [C.c_void_p, u64p, u64p]
    L.tsr_set_collect_errors.argtypes = [C.c_void_p, C.c_int]
    L.tsr_set_instrumented.argtypes = [C.c_void_p, C.c_int]
    L.tsr_cross_check_detcost.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64, u64p]
    L.tsr_cross_check_detcost.restype = C.c_int64
    L.tsr_search_path.argtypes = [C.c_void_p]
    L.tsr_search_path_reason.argtypes = [C.c_void_p]
    L.tsr_search_path_reason.restype = C.c_char_p
    L.tsr_decode.argtypes = [C.c_void_p, u64p, C.c_uint64, C.c_int64, C.c_uint64, u64p, C.c_uint64, u64p, f64p, i32p]
    L.tsr_cost_from_errors.argtypes = [C.c_void_p, u64p, C.c_uint64, f64p]
    L.tsr_flipped_observables.argtypes = [C.c_void_p, u64p, C.c_uint64, u8p]
    L.tsr_counters.argtypes = [C.c_void_p, u64p, C.c_int]
    L.tsr_last_batch_timing.argtypes = [C.c_void_p, f64p]
    L.tsr_last_batch_tail.argtypes = [C.c_void_p, f64p]
    L.tsr_sparsify_reactivate_limit.restype = C.c_int64
    L.tsr_sparsify_reactivate_limit.argtypes = [C.c_void_p]
    L.tsr_suggest_sparsify_reactivate_limit.restype = C.c_int64
    L.tsr_suggest_sparsify_reactivate_limit.argtypes = [C.c_uint64, C.c_int32]
    L.tsr_comm_unique_id.argtypes = [u8p]
    L.tsr_comm_init_rank.argtypes = [C.c_void_p, u8p, C.c_int32, C.c_int32]
    L.tsr_comm_world.argtypes = [C.c_void_p]
    L.tsr_comm_world.restype = C.c_int32
    L.tsr_comm_allreduce_sum_u64.argtypes = [C.c_void_p, u64p, C.c_uint64]
    L.tsr_group_create.argtypes = [C.c_char_p, C.POINTER(TsrConfig), i32p, C.c_uint32, C.POINTER(C.c_void_p)]
    L.tsr_group_destroy.argtypes = [C.c_void_p]
    L.tsr_group_size.argtypes = [C.c_void_p]
    L.tsr_group_size.restype = C.c_uint32
    L.tsr_group_member.argtypes = [C.c_void_p, C.c_uint32]
    L.tsr_group_member.restype = C.c_void_p
    L.tsr_group_decode_b8.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint64, C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p,
                                      C.c_void_p, u64p, u64p]
    L.tsr_group_last_batch_nnz.argtypes = [C.c_void_p]
    L.tsr_group_last_batch_nnz.restype = C.c_uint64
    L.tsr_group_last_batch_errors.argtypes = [C.c_void_p, u64p, u64p]
    L.tsr_visualization_text.argtypes = [C.c_void_p, C.POINTER(C.c_void_p)]
    L.tsr_set_visited_audit.argtypes = [C.c_void_p, C.c_int]
    L.tsr_visited_audit.argtypes = [C.c_void_p, u64p, C.c_int]
    L.tsr_set_profile.argtypes = [C.c_void_p, C.c_int]
    L.tsr_profile.argtypes = [C.c_void_p, u64p, C.c_int]
    L.tsr_stream.restype = C.c_void_p
    L.tsr_stream.argtypes = [C.c_void_p]
    L.tsr_device.argtypes = [C.c_void_p]
    L.tsr_build_det_orders.argtypes = [C.c_char_p, C.c_uint64, C.c_int, C.c_uint64, u64p]
    L.tsr_dem_count_detectors.restype = C.c_uint64
    L.tsr_dem_count_detectors.argtypes = [C.c_char_p]
    L.tsr_dem_count_observables.restype = C.c_uint64
    L.tsr_dem_count_observables.argtypes = [C.c_char_p]
    L.tsr_merge_weights.restype = C.c_double
    L.tsr_merge_weights.argtypes = [C.c_double, C.c_double]
    L.tsr_circuit_to_dem.argtypes = [C.c_char_p, C.POINTER(C.c_void_p)]
    L.tsr_dem_from_counts.argtypes = [C.c_char_p, u64p, C.c_uint64, C.c_uint64, C.POINTER(C.c_void_p)]
    L.tsr_sample_dem_b8.argtypes = [C.c_char_p, C.c_uint64, C.c_uint64, u8p, u8p]
    L.tsr_merge_indistinguishable_errors.argtypes = [C.c_char_p, C.POINTER(C.c_void_p), u64p]
    L.tsr_remove_zero_probability_errors.argtypes = [C.c_char_p, C.POINTER(C.c_void_p), u64p]
    L.tsr_dem_count_errors.restype = C.c_int64
    L.tsr_dem_count_errors.argtypes = [C.c_char_p]
    L.tsr_get_errors_from_dem.restype = C.c_int64
    L.tsr_get_errors_from_dem.argtypes = [C.c_char_p, f64p, u64p, i32p, u64p, i32p, u64p, u64p]
    L.tsr_build_detector_graph.restype = C.c_int64
    L.tsr_build_detector_graph.argtypes = [C.c_char_p, u64p, u64p]
    L.tsr_get_detector_coords.restype = C.c_int64
    L.tsr_get_detector_coords.argtypes = [C.c_char_p, u64p, u64p, f64p]
    _lib = L
    return L


EXPORTED_SYMBOLS = [
    "tsr_config_init", "tsr_last_error", "tsr_version", "tsr_create", "tsr_destroy", "tsr_num_detectors",
    "tsr_num_observables", "tsr_num_errors", "tsr_num_dem_errors", "tsr_num_det_orders", "tsr_error_costs",
    "tsr_maps", "tsr_csr", "tsr_compiled_dem_text", "tsr_free", "tsr_decode_batch_b8",
    "tsr_decode_batch_b8_device", "tsr_set_collect_errors", "tsr_cross_check_detcost", "tsr_set_instrumented", "tsr_search_path",
    "tsr_search_path_reason", "tsr_last_batch_nnz", "tsr_last_batch_errors", "tsr_decode",
    "tsr_cost_from_errors", "tsr_flipped_observables", "tsr_counters", "tsr_last_batch_timing", "tsr_stream",
    "tsr_device", "tsr_build_det_orders", "tsr_dem_count_detectors", "tsr_dem_count_observables", "tsr_merge_weights", "tsr_dem_from_counts", "tsr_circuit_to_dem",
    "tsr_sample_dem_b8", "tsr_set_profile", "tsr_profile", "tsr_sparsify_reactivate_limit",
    "tsr_suggest_sparsify_reactivate_limit", "tsr_comm_unique_id", "tsr_comm_init_rank", "tsr_comm_world",
    "tsr_comm_allreduce_sum_u64", "tsr_nccl_version", "tsr_group_create", "tsr_group_destroy", "tsr_group_size",
    "tsr_group_member", "tsr_group_decode_b8", "tsr_group_last_batch_nnz", "tsr_group_last_batch_errors",
    "tsr_visualization_text", "tsr_set_visited_audit", "tsr_visited_audit", "tsr_last_batch_tail",
    "tsr_merge_indistinguishable_errors", "tsr_remove_zero_probability_errors", "tsr_dem_count_errors",
    "tsr_get_errors_from_dem", "tsr_build_detector_graph", "tsr_get_detector_coords",
]


def _p(arr, ctype):
    return arr.ctypes.data_as(C.POINTER(ctype))


def _raise(rc: int):
    msg = lib().tsr_last_error().decode()
    if rc == TSR_E_INVALID_ARGUMENT:
        raise ValueError(msg)
    if rc == TSR_E_CUDA:
        raise CudaUnavailable(msg)
    if rc == TSR_E_DEVICE_MEMORY:
        raise DeviceMemoryError(msg)
    raise RuntimeError(msg)


def _check(rc: int):
    if rc != 0:
        _raise(rc)


def _take_text(ptr: C.c_void_p) -> str:
    try:
        return C.string_at(ptr).decode()
    finally:
        lib().tsr_free(ptr)


def circuit_to_dem(circuit_text: str) -> str:
    out = C.c_void_p()
    _check(lib().tsr_circuit_to_dem(circuit_text.encode(), C.byref(out)))
    return _take_text(out)


def dem_from_counts(dem_text: str, counts, num_shots: int) -> str:
    """common::dem_from_counts (common.cc:241-263)."""
    counts = np.ascontiguousarray(counts, dtype=np.uint64)
    out = C.c_void_p()
    _check(lib().tsr_dem_from_counts(dem_text.encode(), _p(counts, C.c_uint64), len(counts), num_shots, C.byref(out)))
    return _take_text(out)


def dem_count_errors(dem_text: str) -> int:
    n = lib().tsr_dem_count_errors(dem_text.encode())
    if n < 0:
        _raise(n)
    return n


def _dem_stage(fn, dem_text: str):
    index_map = np.zeros(dem_count_errors(dem_text), dtype=np.uint64)
    out = C.c_void_p()
    _check(fn(dem_text.encode(), C.byref(out), _p(index_map, C.c_uint64)))
    return _take_text(out), index_map


def merge_indistinguishable_errors(dem_text: str):
    """common::merge_indistinguishable_errors (common.cc:139-190) -> (DEM text, error_index_map)."""
    return _dem_stage(lib().tsr_merge_indistinguishable_errors, dem_text)


def remove_zero_probability_errors(dem_text: str):
    """common::remove_zero_probability_errors (common.cc:192-218) -> (DEM text, error_index_map); dropped errors map to
    NO_ERROR_INDEX."""
    return _dem_stage(lib().tsr_remove_zero_probability_errors, dem_text)


def get_errors_from_dem(dem_text: str):
    """get_errors_from_dem (utils.cc:253-261) -> (cost [n], list of detector lists, list of observable lists)."""
    b = dem_text.encode()
    nd, no = C.c_uint64(), C.c_uint64()
    n = lib().tsr_get_errors_from_dem(b, None, None, None, None, None, C.byref(nd), C.byref(no))
    if n < 0:
        _raise(n)
    cost = np.zeros(n, dtype=np.float64)
    doff, ooff = np.zeros(n + 1, dtype=np.uint64), np.zeros(n + 1, dtype=np.uint64)
    dets, obs = np.zeros(max(nd.value, 1), dtype=np.int32), np.zeros(max(no.value, 1), dtype=np.int32)
    n2 = lib().tsr_get_errors_from_dem(b, _p(cost, C.c_double), _p(doff, C.c_uint64), _p(dets, C.c_int32), _p(ooff, C.c_uint64),
                                       _p(obs, C.c_int32), None, None)
    assert n2 == n
    return (cost, [dets[int(doff[i]):int(doff[i + 1])].tolist() for i in range(n)],
            [obs[int(ooff[i]):int(ooff[i + 1])].tolist() for i in range(n)])


def build_detector_graph(dem_text: str):
    """build_detector_graph (utils.cc:60-87) -> one sorted neighbour list per detector."""
    b = dem_text.encode()
    D = dem_count_detectors(dem_text)
    n = lib().tsr_build_detector_graph(b, None, None)
    if n < 0:
        _raise(n)
    off, idx = np.zeros(D + 1, dtype=np.uint64), np.zeros(max(n, 1), dtype=np.uint64)
    lib().tsr_build_detector_graph(b, _p(off, C.c_uint64), _p(idx, C.c_uint64))
    return [idx[int(off[d]):int(off[d + 1])].tolist() for d in range(D)]


def get_detector_coords(dem_text: str):
    """get_detector_coords (utils.cc:32-58) -> the coordinates of every DETECTOR instruction, in DEM order."""
    b = dem_text.encode()
    rows = C.c_uint64()
    n = lib().tsr_get_detector_coords(b, C.byref(rows), None, None)
    if n < 0:
        _raise(n)
    off, xs = np.zeros(rows.value + 1, dtype=np.uint64), np.zeros(max(n, 1), dtype=np.float64)
    lib().tsr_get_detector_coords(b, None, _p(off, C.c_uint64), _p(xs, C.c_double))
    return [xs[int(off[r]):int(off[r + 1])].tolist() for r in range(rows.value)]


def dem_count_detectors(dem_text: str) -> int:
    n = lib().tsr_dem_count_detectors(dem_text.encode())
    if n == (1 << 64) - 1:
        raise ValueError(lib().tsr_last_error().decode())
    return n


def build_det_orders(dem_text: str, num_det_orders: int, method: int = 1, seed: int = 0) -> np.ndarray:
    """utils.cc:192-205; method 0=DetBFS 1=DetIndex 2=DetCoordinate."""
    D = dem_count_detectors(dem_text)
    out = np.zeros((num_det_orders, D), dtype=np.uint64)
    _check(lib().tsr_build_det_orders(dem_text.encode(), num_det_orders, method, seed, _p(out, C.c_uint64)))
    return out


def merge_weights(a: float, b: float) -> float:
    return lib().tsr_merge_weights(a, b)


def dem_count_observables(dem_text: str) -> int:
    n = lib().tsr_dem_count_observables(dem_text.encode())
    if n == (1 << 64) - 1:
        raise ValueError(lib().tsr_last_error().decode())
    return n


def sample_dem_b8(dem_text: str, seed: int, shots: int, num_detectors: int | None = None, num_observables: int | None = None):
    """i.i.d. DEM sampler -> (dets_b8 [shots, ceil(D/8)], obs_b8 [shots, ceil(O/8)]).  The counts are the DEM's own;
    the optional arguments are only cross-checked (the C side writes ceil(D/8) / ceil(O/8) bytes per shot)."""
    D, O = dem_count_detectors(dem_text), dem_count_observables(dem_text)
    if (num_detectors is not None and num_detectors != D) or (num_observables is not None and num_observables != O):
        raise ValueError(f"the DEM has {D} detectors and {O} observables, not {num_detectors} / {num_observables}")
    num_detectors, num_observables = D, O
    dets = np.zeros((shots, (num_detectors + 7) // 8), dtype=np.uint8)
    obs = np.zeros((shots, max((num_observables + 7) // 8, 1)), dtype=np.uint8)
    _check(lib().tsr_sample_dem_b8(dem_text.encode(), seed, shots, _p(dets, C.c_uint8), _p(obs, C.c_uint8)))
    return dets, obs[:, : (num_observables + 7) // 8]


def make_config(*, det_beam: int = 5, beam_climbing: bool = False, no_revisit_dets: bool = True, merge_errors: bool = True,
                pqlimit: int | None = 200000, det_penalty: float = 0.0, det_orders=None,
                at_most_two_errors_per_detector: bool = False, device: int = -1, search_slots: int = 0, warps_per_block: int = 0,
                node_pool_bytes: int = 0, max_batch_shots: int = 0, force_generic_kernel: bool = False, row_scan: int = 0,
                sparsify_errors: bool = False, sparsify_base_degree: int = -1, sparsify_max_degree: int = -1,
                sparsify_reactivate_limit: int = -1, create_visualization: bool = False):
    """(tsr_config, keep-alive object for the det_orders array)."""
    cfg = TsrConfig()
    lib().tsr_config_init(C.byref(cfg))
    cfg.det_beam = det_beam
    cfg.beam_climbing = int(beam_climbing)
    cfg.no_revisit_dets = int(no_revisit_dets)
    cfg.merge_errors = int(merge_errors)
    cfg.at_most_two_errors_per_detector = int(at_most_two_errors_per_detector)
    cfg.device = device
    cfg.pqlimit = PQLIMIT_UNBOUNDED if pqlimit is None or pqlimit >= PQLIMIT_UNBOUNDED else int(pqlimit)
    cfg.det_penalty = det_penalty
    orders = None
    if det_orders is not None and len(det_orders):
        orders = np.ascontiguousarray(det_orders, dtype=np.uint64)
        if orders.ndim != 2:
            raise ValueError("det_orders must be a 2D array [num_det_orders, num_detectors]")
        cfg.num_det_orders = orders.shape[0]
        cfg.det_order_len = orders.shape[1]
        cfg.det_orders = _p(orders, C.c_uint64)
    cfg.search_slots = search_slots
    cfg.warps_per_block = warps_per_block
    cfg.node_pool_bytes = node_pool_bytes
    cfg.max_batch_shots = max_batch_shots
    cfg.force_generic_kernel = int(force_generic_kernel)
    cfg.row_scan = int(row_scan)  # 0 auto, 1 entry per lane, 2 bit-sliced
    cfg.sparsify_errors = int(sparsify_errors)
    cfg.sparsify_base_degree = sparsify_base_degree
    cfg.sparsify_max_degree = sparsify_max_degree
    cfg.sparsify_reactivate_limit = sparsify_reactivate_limit
    cfg.create_visualization = int(create_visualization)
    return cfg, orders


def comm_unique_id() -> bytes:
    """ncclGetUniqueId (rank 0); ship the bytes to the other ranks, then Decoder.comm_init_rank on every rank."""
    out = np.zeros(128, dtype=np.uint8)
    _check(lib().tsr_comm_unique_id(_p(out, C.c_uint8)))
    return out.tobytes()


class Decoder:
    """Owns one tsr_decoder handle."""

    def __init__(self, dem_text: str, *, host_only: bool = False, _borrowed=None, **kw):
        L = lib()
        if _borrowed is not None:  # a member of a Group: the group owns the handle
            self._owns, h = False, C.c_void_p(_borrowed)
        else:
            cfg, self._orders = make_config(**kw)
            self._owns, h = True, C.c_void_p()
            self.h = None
            _check(L.tsr_create(dem_text.encode(), C.byref(cfg), CREATE_HOST_ONLY if host_only else 0, C.byref(h)))
        self.h = h
        self.num_detectors = L.tsr_num_detectors(h)
        self.num_observables = L.tsr_num_observables(h)
        self.num_errors = L.tsr_num_errors(h)
        self.num_dem_errors = L.tsr_num_dem_errors(h)
        self.num_det_orders = L.tsr_num_det_orders(h)
        self.sparsify_reactivate_limit = L.tsr_sparsify_reactivate_limit(h)  # effective value; -1 when sparsify is off

    def close(self):
        if getattr(self, "h", None):
            if self._owns:
                lib().tsr_destroy(self.h)
            self.h = None

    # ---- multi-GPU, one process per GPU (csrc/group.cu) ------------------------------------------------
    def comm_init_rank(self, unique_id: bytes, rank: int, world: int):
        buf = np.frombuffer(unique_id, dtype=np.uint8).copy()
        _check(lib().tsr_comm_init_rank(self.h, _p(buf, C.c_uint8), rank, world))

    @property
    def comm_world(self) -> int:
        return lib().tsr_comm_world(self.h)

    def allreduce_sum(self, values):
        """One ncclAllReduce(sum) of u64 counters over the decoder's communicator; returns the sums."""
        v = np.ascontiguousarray(values, dtype=np.uint64).copy()
        _check(lib().tsr_comm_allreduce_sum_u64(self.h, _p(v, C.c_uint64), len(v)))
        return v

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- tables -------------------------------------------------------------------------------
    def error_costs(self) -> np.ndarray:
        out = np.zeros(self.num_errors, dtype=np.float64)
        _check(lib().tsr_error_costs(self.h, _p(out, C.c_double)))
        return out

    def maps(self):
        a = np.zeros(self.num_dem_errors, dtype=np.uint64)
        b = np.zeros(self.num_errors, dtype=np.uint64)
        _check(lib().tsr_maps(self.h, _p(a, C.c_uint64), _p(b, C.c_uint64)))
        return a, b

    def csr(self, which: int):
        rows = self.num_detectors if which == 0 else self.num_errors
        nnz = lib().tsr_csr(self.h, which, None, None)
        if nnz < 0:
            _raise(int(nnz))
        off = np.zeros(rows + 1, dtype=np.uint64)
        idx = np.zeros(max(nnz, 1), dtype=np.int32)
        lib().tsr_csr(self.h, which, _p(off, C.c_uint64), _p(idx, C.c_int32))
        return off, idx[:nnz]

    def rows(self, which: int):
        off, idx = self.csr(which)
        return [idx[int(off[r]):int(off[r + 1])].tolist() for r in range(len(off) - 1)]

    def visualization_text(self) -> str:
        """The Visualizer's lines (visualization.cc) recorded so far; empty unless create_visualization."""
        out = C.c_void_p()
        _check(lib().tsr_visualization_text(self.h, C.byref(out)))
        return _take_text(out)

    def compiled_dem_text(self) -> str:
        out = C.c_void_p()
        _check(lib().tsr_compiled_dem_text(self.h, C.byref(out)))
        return _take_text(out)

    # ---- decoding -----------------------------------------------------------------------------
    def decode_to_errors(self, detections, order: int = -1, beam: int = 0):
        """(predicted_errors_buffer, cost_from_errors(buffer), low_confidence_flag)."""
        det = np.ascontiguousarray(detections, dtype=np.uint64)
        cap = 1 << 16
        errs = np.zeros(cap, dtype=np.uint64)
        n, cost, low = C.c_uint64(), C.c_double(), C.c_int32()
        _check(lib().tsr_decode(self.h, _p(det, C.c_uint64), len(det), order, beam, _p(errs, C.c_uint64), cap,
                                C.byref(n), C.byref(cost), C.byref(low)))
        return errs[: n.value].tolist(), cost.value, bool(low.value)

    @property
    def search_path(self) -> str:
        return ("generic", "fast", "fast-bits")[lib().tsr_search_path(self.h)]

    @property
    def search_path_reason(self) -> str:
        return lib().tsr_search_path_reason(self.h).decode()

    def cross_check_detcost(self, trials: int, seed: int = 1):
        """(mismatches, evaluations compared) of the host cross-check of the three get_detcost formulations."""
        n = C.c_uint64()
        bad = lib().tsr_cross_check_detcost(self.h, trials, seed, C.byref(n))
        if bad < 0:
            _raise(int(bad))
        return int(bad), n.value

    def set_instrumented(self, on: bool):
        _check(lib().tsr_set_instrumented(self.h, int(on)))

    def set_collect_errors(self, on: bool):
        _check(lib().tsr_set_collect_errors(self.h, int(on)))

    def last_batch_errors(self):
        nnz = lib().tsr_last_batch_nnz(self.h)
        off = np.zeros(self._last_shots + 1, dtype=np.uint64)
        idx = np.zeros(max(nnz, 1), dtype=np.uint64)
        _check(lib().tsr_last_batch_errors(self.h, _p(off, C.c_uint64), _p(idx, C.c_uint64)))
        return off, idx[:nnz]

    def decode_batch_b8(self, dets_b8: np.ndarray, want_errors: bool = True):
        """Host buffers in, host buffers out (H2D and D2H inside the call)."""
        dets_b8 = np.ascontiguousarray(dets_b8, dtype=np.uint8)
        if dets_b8.ndim != 2:
            raise ValueError("dets_b8 must be a 2D uint8 array [shots, ceil(num_detectors/8)]")
        shots, stride = dets_b8.shape
        obs = np.zeros((shots, (self.num_observables + 7) // 8), dtype=np.uint8)
        cost = np.zeros(shots, dtype=np.float64)
        low = np.zeros(shots, dtype=np.uint8)
        _check(lib().tsr_decode_batch_b8(self.h, dets_b8.ctypes.data, stride, shots, obs.ctypes.data, cost.ctypes.data,
                                         low.ctypes.data))
        self._last_shots = shots
        out = {"obs": obs, "cost": cost, "low_confidence": low}
        if want_errors:
            out["err_offsets"], out["err_idx"] = self.last_batch_errors()
        return out

    def decode_batch_b8_raw(self, dets_ptr: int, stride: int, shots: int, obs_ptr: int = 0, cost_ptr: int = 0,
                            low_ptr: int = 0):
        """tsr_decode_batch_b8 on caller-owned HOST buffers given as addresses (e.g. pinned memory)."""
        _check(lib().tsr_decode_batch_b8(self.h, dets_ptr, stride, shots, obs_ptr or None, cost_ptr or None, low_ptr or None))
        self._last_shots = shots

    def decode_batch_b8_device(self, d_dets: int, stride: int, shots: int, d_obs: int = 0, d_cost: int = 0,
                               d_low: int = 0):
        """Device pointers (ints) in and out; inputs already resident in HBM."""
        _check(lib().tsr_decode_batch_b8_device(self.h, d_dets, stride, shots, d_obs or None, d_cost or None,
                                                d_low or None))
        self._last_shots = shots

    def cost_from_errors(self, errs) -> float:
        errs = np.ascontiguousarray(errs, dtype=np.uint64)
        out = C.c_double()
        _check(lib().tsr_cost_from_errors(self.h, _p(errs, C.c_uint64), len(errs), C.byref(out)))
        return out.value

    def flipped_observables(self, errs):
        errs = np.ascontiguousarray(errs, dtype=np.uint64)
        out = np.zeros(max((self.num_observables + 7) // 8, 1), dtype=np.uint8)
        _check(lib().tsr_flipped_observables(self.h, _p(errs, C.c_uint64), len(errs), _p(out, C.c_uint8)))
        bits = np.unpackbits(out, bitorder="little")[: self.num_observables]
        return np.nonzero(bits)[0].tolist()

    def counters(self, reset: bool = False):
        out = np.zeros(8, dtype=np.uint64)
        _check(lib().tsr_counters(self.h, _p(out, C.c_uint64), int(reset)))
        return dict(zip(["Q", "P", "X", "L", "S", "A", "V", "searches"], out.tolist()))

    def set_visited_audit(self, on: bool):
        _check(lib().tsr_set_visited_audit(self.h, int(on)))

    def visited_audit(self, reset: bool = False):
        """(fingerprint hits checked against the real fired sets, hits whose sets differ)."""
        out = np.zeros(2, dtype=np.uint64)
        _check(lib().tsr_visited_audit(self.h, _p(out, C.c_uint64), int(reset)))
        return int(out[0]), int(out[1])

    def set_profile(self, on: bool):
        _check(lib().tsr_set_profile(self.h, int(on)))

    PROFILE_PHASES = ["pop", "rebuild_goal_visited", "node_terms", "min_detector_child_list", "children", "push", "staging_cleanup"]

    def profile(self, reset: bool = False):
        """SM clock ticks per expansion phase summed over search CTAs, plus the expansion count (tsr_profile)."""
        out = np.zeros(8, dtype=np.uint64)
        _check(lib().tsr_profile(self.h, _p(out, C.c_uint64), int(reset)))
        d = dict(zip(self.PROFILE_PHASES, out[:7].tolist()))
        d["expansions"] = int(out[7])
        return d

    def last_batch_timing(self):
        out = np.zeros(4, dtype=np.float64)
        _check(lib().tsr_last_batch_timing(self.h, _p(out, C.c_double)))
        tail = np.zeros(3, dtype=np.float64)
        _check(lib().tsr_last_batch_tail(self.h, _p(tail, C.c_double)))
        return {"search_ms": out[0], "pipeline_ms": out[1], "search_launches": int(out[2]), "rerun_items": int(out[3]),
                "tail_searches": int(tail[0]), "tail_ms": float(tail[1]), "tail_cluster": int(tail[2])}

    @property
    def stream(self) -> int:
        return lib().tsr_stream(self.h) or 0

    @property
    def device(self) -> int:
        return lib().tsr_device(self.h)


class Group:
    """tsr_group: one process driving several GPUs (one decoder, host thread and NCCL communicator per device)."""

    def __init__(self, dem_text: str, devices, **kw):
        cfg, self._orders = make_config(**kw)
        devs = np.ascontiguousarray(list(devices), dtype=np.int32)
        g = C.c_void_p()
        self.g = None
        _check(lib().tsr_group_create(dem_text.encode(), C.byref(cfg), _p(devs, C.c_int32), len(devs), C.byref(g)))
        self.g = g
        self.size = lib().tsr_group_size(g)
        self.members = [Decoder("", _borrowed=lib().tsr_group_member(g, i)) for i in range(self.size)]
        self.num_detectors, self.num_observables = self.members[0].num_detectors, self.members[0].num_observables
        self.num_dem_errors = self.members[0].num_dem_errors

    def close(self):
        if getattr(self, "g", None):
            for m in self.members:
                m.close()
            lib().tsr_group_destroy(self.g)
            self.g = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def decode_batch_b8(self, dets_b8: np.ndarray, obs_true: np.ndarray | None = None, chunk_shots: int = 0, want_error_use: bool = False):
        dets_b8 = np.ascontiguousarray(dets_b8, dtype=np.uint8)
        shots, stride = dets_b8.shape
        obs = np.zeros((shots, (self.num_observables + 7) // 8), dtype=np.uint8)
        cost = np.zeros(shots, dtype=np.float64)
        low = np.zeros(shots, dtype=np.uint8)
        tally = np.zeros(3, dtype=np.uint64)
        use = np.zeros(max(self.num_dem_errors, 1), dtype=np.uint64) if want_error_use else None
        truth = np.ascontiguousarray(obs_true, dtype=np.uint8) if obs_true is not None else None
        _check(lib().tsr_group_decode_b8(self.g, dets_b8.ctypes.data, stride, shots, chunk_shots,
                                         truth.ctypes.data if truth is not None else None, obs.ctypes.data, cost.ctypes.data,
                                         low.ctypes.data, _p(tally, C.c_uint64), _p(use, C.c_uint64) if use is not None else None))
        out = {"obs": obs, "cost": cost, "low_confidence": low,
               "tally": {"shots": int(tally[0]), "logical_errors": int(tally[1]), "low_confidence": int(tally[2])}}
        if use is not None:
            out["error_use"] = use[: self.num_dem_errors]
            nnz = lib().tsr_group_last_batch_nnz(self.g)
            off = np.zeros(shots + 1, dtype=np.uint64)
            idx = np.zeros(max(nnz, 1), dtype=np.uint64)
            _check(lib().tsr_group_last_batch_errors(self.g, _p(off, C.c_uint64), _p(idx, C.c_uint64)))
            out["err_offsets"], out["err_idx"] = off, idx[:nnz]
        return out
