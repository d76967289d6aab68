"""TEST INFRASTRUCTURE — NOT PRODUCT CODE.

ctypes face over oracle/oracle_api.h.  Loads either checker:

  kind="reference"  oracle/_ref/libtesseract_ref.so    (the reference's own sources, prebuilt)
  kind="port"       oracle/_build/libtesseract_port.so (the CPU restatement, oracle/port.cc)

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
PATHS = {
    "reference": os.path.join(HERE, "_ref", "libtesseract_ref.so"),
    "port": os.path.join(HERE, "_build", "libtesseract_port.so"),
}
INF_DET_BEAM = 65535
U64_MAX = (1 << 64) - 1


class OrcConfig(C.Structure):
    _fields_ = [
        ("det_beam", C.c_int32),
        ("beam_climbing", C.c_int32),
        ("no_revisit_dets", C.c_int32),
        ("merge_errors", C.c_int32),
        ("pqlimit", C.c_uint64),
        ("det_penalty", C.c_double),
        ("num_det_orders", C.c_uint64),
        ("det_orders", C.POINTER(C.c_uint64)),
        ("num_threads", C.c_int32),
        ("sparsify_errors", C.c_int32),
        ("sparsify_base_degree", C.c_int32),
        ("sparsify_max_degree", C.c_int32),
        ("sparsify_reactivate_limit", C.c_int32),
        ("at_most_two_errors_per_detector", C.c_int32),
        ("create_visualization", C.c_int32),
    ]


def build(kind: str) -> None:
    """Compile a checker with oracle/Makefile ("ref" needs /root/reference)."""
    target = "ref" if kind == "reference" else "port"
    subprocess.run(["make", "-s", "-C", HERE, target], check=True)


def available(kind: str) -> bool:
    return os.path.exists(PATHS[kind])


_libs: dict[str, C.CDLL] = {}


def _lib(kind: str) -> C.CDLL:
    if kind in _libs:
        return _libs[kind]
    lib = C.CDLL(PATHS[kind])
    u64p, f64p, u8p, i32p = (C.POINTER(t) for t in (C.c_uint64, C.c_double, C.c_uint8, C.c_int32))
    lib.orc_last_error.restype = C.c_char_p
    lib.orc_kind.restype = C.c_char_p
    lib.orc_create.restype = C.c_void_p
    lib.orc_create.argtypes = [C.c_char_p, C.POINTER(OrcConfig)]
    lib.orc_destroy.argtypes = [C.c_void_p]
    for name in ("orc_num_detectors", "orc_num_observables", "orc_num_errors", "orc_num_dem_errors",
                 "orc_last_batch_nnz"):
        getattr(lib, name).restype = C.c_uint64
        getattr(lib, name).argtypes = [C.c_void_p]
    lib.orc_error_costs.argtypes = [C.c_void_p, f64p]
    lib.orc_maps.argtypes = [C.c_void_p, u64p, u64p]
    lib.orc_csr.restype = C.c_int64
    lib.orc_csr.argtypes = [C.c_void_p, C.c_int, u64p, i32p]
    lib.orc_decode.argtypes = [C.c_void_p, u64p, C.c_uint64, C.c_int64, C.c_uint64, u64p, C.c_uint64,
                               u64p, f64p, i32p]
    lib.orc_decode_batch_b8.argtypes = [C.c_void_p, u8p, C.c_uint64, C.c_uint64, u8p, f64p, u8p, f64p, f64p]
    lib.orc_last_batch_errors.argtypes = [C.c_void_p, u64p, u64p]
    lib.orc_cost_from_errors.argtypes = [C.c_void_p, u64p, C.c_uint64, f64p]
    lib.orc_flipped_observables.argtypes = [C.c_void_p, u64p, C.c_uint64, u8p]
    lib.orc_build_det_orders.argtypes = [C.c_char_p, C.c_uint64, C.c_int, C.c_uint64, u64p]
    lib.orc_merge_weights.restype = C.c_double
    lib.orc_merge_weights.argtypes = [C.c_double, C.c_double]
    lib.orc_counters.argtypes = [C.c_void_p, u64p, C.c_int]
    lib.orc_visualization_write.argtypes = [C.c_void_p, C.c_char_p]
    lib.orc_sample_dem_b8.argtypes = [C.c_char_p, C.c_uint64, C.c_uint64, u8p, u8p]
    if kind == "reference":  # the reference's standalone ingestion functions (oracle_api.h: REFERENCE BUILD ONLY)
        lib.orc_dem_dump.restype = C.c_int64
        lib.orc_dem_dump.argtypes = [C.c_char_p, C.c_int, u64p, u8p, u64p, f64p, u64p, u64p, u64p]
        lib.orc_get_errors_from_dem.restype = C.c_int64
        lib.orc_get_errors_from_dem.argtypes = [C.c_char_p, f64p, u64p, i32p, u64p, i32p, u64p, u64p]
        lib.orc_build_detector_graph.restype = C.c_int64
        lib.orc_build_detector_graph.argtypes = [C.c_char_p, u64p, u64p]
        lib.orc_get_detector_coords.restype = C.c_int64
        lib.orc_get_detector_coords.argtypes = [C.c_char_p, u64p, u64p, f64p]
        lib.orc_error_text.argtypes = [C.c_double, i32p, C.c_uint64, i32p, C.c_uint64, C.c_char_p, C.c_uint64, f64p]
        lib.orc_cost_of_probability.argtypes = [C.c_double, f64p]
    _libs[kind] = lib
    return lib


def _p(arr, ctype):
    return arr.ctypes.data_as(C.POINTER(ctype))


class OracleError(RuntimeError):
    pass


def build_det_orders(dem_text: str, num_det_orders: int, method: int = 1, seed: int = 0,
                     kind: str = "reference") -> np.ndarray:
    """utils.cc:192-205; method 0=DetBFS 1=DetIndex 2=DetCoordinate."""
    lib = _lib(kind)
    tmp = OracleDecoder(dem_text, kind=kind)
    D = tmp.num_detectors
    tmp.close()
    out = np.zeros((num_det_orders, D), dtype=np.uint64)
    if lib.orc_build_det_orders(dem_text.encode(), num_det_orders, method, seed, _p(out, C.c_uint64)) != 0:
        raise OracleError(lib.orc_last_error().decode())
    return out


def dem_dump(dem_text: str, which: int = 0):
    """Reference build only.  The DEM's instructions after which = 0 flatten / 1 merge_indistinguishable_errors /
    2 remove_zero_probability_errors (common.cc:135-218) as (instructions, error_index_map): each instruction is
    (kind, args tuple, targets tuple) with kind 0 error / 1 detector / 2 logical_observable, targets as detector id,
    observable id | 1<<63, or U64_MAX for the separator.  Doubles are compared as the bits they are."""
    lib = _lib("reference")
    b = dem_text.encode()
    sizes = np.zeros(4, dtype=np.uint64)
    if lib.orc_dem_dump(b, which, None, None, None, None, None, None, _p(sizes, C.c_uint64)) < 0:
        raise OracleError(lib.orc_last_error().decode())
    n, na, nt, nm = (int(v) for v in sizes)
    kind = np.zeros(max(n, 1), dtype=np.uint8)
    aoff, toff = np.zeros(n + 1, dtype=np.uint64), np.zeros(n + 1, dtype=np.uint64)
    args, tgt = np.zeros(max(na, 1), dtype=np.float64), np.zeros(max(nt, 1), dtype=np.uint64)
    imap = np.zeros(max(nm, 1), dtype=np.uint64)
    lib.orc_dem_dump(b, which, _p(imap, C.c_uint64), _p(kind, C.c_uint8), _p(aoff, C.c_uint64), _p(args, C.c_double),
                     _p(toff, C.c_uint64), _p(tgt, C.c_uint64), _p(sizes, C.c_uint64))
    inst = [(int(kind[i]), tuple(args[int(aoff[i]):int(aoff[i + 1])].tolist()), tuple(tgt[int(toff[i]):int(toff[i + 1])].tolist()))
            for i in range(n)]
    return inst, imap[:nm]


def get_errors_from_dem(dem_text: str):
    """Reference build only.  get_errors_from_dem (utils.cc:253-261) -> (cost [n], detector lists, observable lists)."""
    lib = _lib("reference")
    b = dem_text.encode()
    nd, no = C.c_uint64(), C.c_uint64()
    n = lib.orc_get_errors_from_dem(b, None, None, None, None, None, C.byref(nd), C.byref(no))
    if n < 0:
        raise OracleError(lib.orc_last_error().decode())
    cost = np.zeros(n, dtype=np.float64)
    doff, ooff = np.zeros(n + 1, dtype=np.uint64), np.zeros(n + 1, dtype=np.uint64)
    dets, obs = np.zeros(max(nd.value, 1), dtype=np.int32), np.zeros(max(no.value, 1), dtype=np.int32)
    lib.orc_get_errors_from_dem(b, _p(cost, C.c_double), _p(doff, C.c_uint64), _p(dets, C.c_int32), _p(ooff, C.c_uint64),
                                _p(obs, C.c_int32), None, None)
    return (cost, [dets[int(doff[i]):int(doff[i + 1])].tolist() for i in range(n)],
            [obs[int(ooff[i]):int(ooff[i + 1])].tolist() for i in range(n)])


def build_detector_graph(dem_text: str, num_detectors: int):
    """Reference build only.  build_detector_graph (utils.cc:60-87)."""
    lib = _lib("reference")
    b = dem_text.encode()
    n = lib.orc_build_detector_graph(b, None, None)
    if n < 0:
        raise OracleError(lib.orc_last_error().decode())
    off, idx = np.zeros(num_detectors + 1, dtype=np.uint64), np.zeros(max(n, 1), dtype=np.uint64)
    lib.orc_build_detector_graph(b, _p(off, C.c_uint64), _p(idx, C.c_uint64))
    return [idx[int(off[d]):int(off[d + 1])].tolist() for d in range(num_detectors)]


def get_detector_coords(dem_text: str):
    """Reference build only.  get_detector_coords (utils.cc:32-58)."""
    lib = _lib("reference")
    b = dem_text.encode()
    rows = C.c_uint64()
    n = lib.orc_get_detector_coords(b, C.byref(rows), None, None)
    if n < 0:
        raise OracleError(lib.orc_last_error().decode())
    off, xs = np.zeros(rows.value + 1, dtype=np.uint64), np.zeros(max(n, 1), dtype=np.float64)
    lib.orc_get_detector_coords(b, None, _p(off, C.c_uint64), _p(xs, C.c_double))
    return [xs[int(off[r]):int(off[r + 1])].tolist() for r in range(rows.value)]


def error_text(likelihood_cost: float, detectors, observables):
    """Reference build only.  (common::Error(...).str(), .get_probability()) — common.cc:88-98."""
    lib = _lib("reference")
    d, o = np.asarray(list(detectors) + [0], dtype=np.int32), np.asarray(list(observables) + [0], dtype=np.int32)
    buf = C.create_string_buffer(4096)
    p = C.c_double()
    if lib.orc_error_text(likelihood_cost, _p(d, C.c_int32), len(detectors), _p(o, C.c_int32), len(observables), buf, 4096,
                          C.byref(p)) != 0:
        raise OracleError(lib.orc_last_error().decode())
    return buf.value.decode(), p.value


def cost_of_probability(p: float) -> float:
    """Reference build only.  Error::set_with_probability (common.cc:100-105); raises OracleError outside (0, 1)."""
    lib = _lib("reference")
    c = C.c_double()
    if lib.orc_cost_of_probability(p, C.byref(c)) != 0:
        raise OracleError(lib.orc_last_error().decode())
    return c.value


def sample_dem_b8(dem_text: str, seed: int, shots: int, kind: str = "reference"):
    """Synthetic i.i.d. shots of the DEM (oracle/sampler.inc): (dets [shots, ceil(D/8)], obs [shots, ceil(O/8)])."""
    lib = _lib(kind)
    tmp = OracleDecoder(dem_text, kind=kind)
    D, O = tmp.num_detectors, tmp.num_observables
    tmp.close()
    dets = np.zeros((shots, (D + 7) // 8), dtype=np.uint8)
    obs = np.zeros((shots, (O + 7) // 8), dtype=np.uint8)
    if lib.orc_sample_dem_b8(dem_text.encode(), seed, shots, _p(dets, C.c_uint8), _p(obs, C.c_uint8)) != 0:
        raise OracleError(lib.orc_last_error().decode())
    return dets, obs


def merge_weights(a: float, b: float, kind: str = "reference") -> float:
    return _lib(kind).orc_merge_weights(a, b)


class OracleDecoder:
    """A pool of CPU decoders (one per worker thread) behind oracle_api.h."""

    def __init__(self, dem_text: str, *, det_beam: int = 5, beam_climbing: bool = False,
                 no_revisit_dets: bool = True, merge_errors: bool = True, pqlimit: int = 200000,
                 det_penalty: float = 0.0, det_orders=None, num_threads: int = 1, kind: str = "reference",
                 sparsify_errors: bool = False, sparsify_base_degree: int = -1, sparsify_max_degree: int = -1,
                 sparsify_reactivate_limit: int = -1, at_most_two_errors_per_detector: bool = False,
                 create_visualization: bool = False):
        self.lib = _lib(kind)
        self.kind = kind
        cfg = OrcConfig()
        cfg.det_beam = det_beam
        cfg.beam_climbing = int(beam_climbing)
        cfg.no_revisit_dets = int(no_revisit_dets)
        cfg.merge_errors = int(merge_errors)
        cfg.pqlimit = U64_MAX if pqlimit is None or pqlimit >= U64_MAX else int(pqlimit)
        cfg.det_penalty = det_penalty
        self._orders = None
        if det_orders is not None and len(det_orders):
            self._orders = np.ascontiguousarray(det_orders, dtype=np.uint64)
            cfg.num_det_orders = self._orders.shape[0]
            cfg.det_orders = _p(self._orders, C.c_uint64)
        cfg.num_threads = num_threads
        cfg.sparsify_errors = int(sparsify_errors)
        cfg.sparsify_base_degree = sparsify_base_degree
        cfg.sparsify_max_degree = sparsify_max_degree
        cfg.sparsify_reactivate_limit = sparsify_reactivate_limit
        cfg.at_most_two_errors_per_detector = int(at_most_two_errors_per_detector)
        cfg.create_visualization = int(create_visualization)
        self.h = self.lib.orc_create(dem_text.encode(), C.byref(cfg))
        if not self.h:
            raise OracleError(self.lib.orc_last_error().decode())
        self.num_detectors = self.lib.orc_num_detectors(self.h)
        self.num_observables = self.lib.orc_num_observables(self.h)
        self.num_errors = self.lib.orc_num_errors(self.h)
        self.num_dem_errors = self.lib.orc_num_dem_errors(self.h)

    def close(self):
        if getattr(self, "h", None):
            self.lib.orc_destroy(self.h)
            self.h = None

    def __del__(self):
        self.close()

    def _check(self, rc):
        if rc != 0:
            raise OracleError(self.lib.orc_last_error().decode())

    # ---- tables -------------------------------------------------------------------------------
    def error_costs(self) -> np.ndarray:
        out = np.zeros(self.num_errors, dtype=np.float64)
        self._check(self.lib.orc_error_costs(self.h, _p(out, C.c_double)))
        return out

    def maps(self):
        a = np.zeros(self.num_dem_errors, dtype=np.uint64)
        b = np.zeros(self.num_errors, dtype=np.uint64)
        self._check(self.lib.orc_maps(self.h, _p(a, C.c_uint64), _p(b, C.c_uint64)))
        return a, b

    def csr(self, which: int):
        rows = self.num_detectors if which == 0 else self.num_errors
        nnz = self.lib.orc_csr(self.h, which, None, None)
        if nnz < 0:
            raise OracleError(self.lib.orc_last_error().decode())
        off = np.zeros(rows + 1, dtype=np.uint64)
        idx = np.zeros(max(nnz, 1), dtype=np.int32)
        self.lib.orc_csr(self.h, which, _p(off, C.c_uint64), _p(idx, C.c_int32))
        return off, idx[:nnz]

    def rows(self, which: int):
        off, idx = self.csr(which)
        return [idx[int(off[r]):int(off[r + 1])].tolist() for r in range(len(off) - 1)]

    # ---- decoding -----------------------------------------------------------------------------
    def decode_to_errors(self, hits, order: int = -1, beam: int = 0):
        hits = np.ascontiguousarray(hits, dtype=np.uint64)
        cap = 4096
        errs = np.zeros(cap, dtype=np.uint64)
        n = C.c_uint64()
        cost = C.c_double()
        low = C.c_int32()
        self._check(self.lib.orc_decode(self.h, _p(hits, C.c_uint64), len(hits), order, beam,
                                        _p(errs, C.c_uint64), cap, C.byref(n), C.byref(cost), C.byref(low)))
        return errs[: n.value].tolist(), cost.value, bool(low.value)

    def decode_batch_b8(self, dets_b8: np.ndarray):
        dets_b8 = np.ascontiguousarray(dets_b8, dtype=np.uint8)
        shots, stride = dets_b8.shape
        obs = np.zeros((shots, (self.num_observables + 7) // 8), dtype=np.uint8)
        cost = np.zeros(shots, dtype=np.float64)
        low = np.zeros(shots, dtype=np.uint8)
        wall = C.c_double()
        cpu = C.c_double()
        self._check(self.lib.orc_decode_batch_b8(self.h, _p(dets_b8, C.c_uint8), stride, shots,
                                                 _p(obs, C.c_uint8), _p(cost, C.c_double), _p(low, C.c_uint8),
                                                 C.byref(wall), C.byref(cpu)))
        nnz = self.lib.orc_last_batch_nnz(self.h)
        off = np.zeros(shots + 1, dtype=np.uint64)
        idx = np.zeros(max(nnz, 1), dtype=np.uint64)
        self._check(self.lib.orc_last_batch_errors(self.h, _p(off, C.c_uint64), _p(idx, C.c_uint64)))
        return {"obs": obs, "cost": cost, "low_confidence": low, "err_offsets": off, "err_idx": idx[:nnz],
                "wall_seconds": wall.value, "sum_shot_seconds": cpu.value}

    def cost_from_errors(self, errs) -> float:
        errs = np.ascontiguousarray(errs, dtype=np.uint64)
        out = C.c_double()
        self._check(self.lib.orc_cost_from_errors(self.h, _p(errs, C.c_uint64), len(errs), C.byref(out)))
        return out.value

    def flipped_observables(self, errs):
        errs = np.ascontiguousarray(errs, dtype=np.uint64)
        out = np.zeros((self.num_observables + 7) // 8 or 1, dtype=np.uint8)
        self._check(self.lib.orc_flipped_observables(self.h, _p(errs, C.c_uint64), len(errs), _p(out, C.c_uint8)))
        bits = np.unpackbits(out, bitorder="little")[: self.num_observables]
        return np.nonzero(bits)[0].tolist()

    def visualization_text(self) -> str:
        """Visualizer::write output (visualization.cc) of the pool's first decoder."""
        import tempfile
        with tempfile.NamedTemporaryFile("r", suffix=".log") as f:
            self._check(self.lib.orc_visualization_write(self.h, f.name.encode()))
            return open(f.name).read()

    def counters(self, reset: bool = False):
        out = np.zeros(8, dtype=np.uint64)
        self._check(self.lib.orc_counters(self.h, _p(out, C.c_uint64), int(reset)))
        names = ["Q", "P", "X", "L", "S", "A", "V", "searches"]
        return dict(zip(names, out.tolist()))
