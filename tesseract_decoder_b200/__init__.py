"""tesseract_decoder_b200 — B200-native (sm_100a CUDA) batched A* most-likely-error decode path of Tesseract.

Two faces over the same C ABI (include/tesseract_b200.h, libtesseract_b200.so):
  * `capi`  — thin ctypes binding (what the parity tests and bench.py call);
  * `tesseract`, `utils`, `common`, `tesseract_sinter_compat`, `viz`, `ostream_redirect` — the reference's Python API
    names (src/tesseract.pybind.cc, src/*.pybind.h) from the pybind11 module `_tesseract_b200`.
Both are built in-tree by `__graft_entry__.build()` / `make -C tesseract_decoder_b200/csrc`.  There is no CPU
fallback: decoding without an sm_100 GPU raises.
"""
from . import capi  # noqa: F401

try:
    from . import _tesseract_b200 as _ext
except ImportError as _e:  # the extension has not been built
    _ext = None
    _ext_error = _e

if _ext is not None:
    tesseract = _ext.tesseract
    utils = _ext.utils
    common = _ext.common
    tesseract_sinter_compat = _ext.tesseract_sinter_compat
    viz = _ext.viz
    ostream_redirect = _ext.ostream_redirect
    TesseractSinterDecoder = tesseract_sinter_compat.TesseractSinterDecoder
    make_tesseract_sinter_decoders_dict = tesseract_sinter_compat.make_tesseract_sinter_decoders_dict


def __getattr__(name):
    if name in ("tesseract", "utils", "common", "tesseract_sinter_compat", "viz", "ostream_redirect", "TesseractSinterDecoder",
                "make_tesseract_sinter_decoders_dict"):
        raise ImportError(f"tesseract_decoder_b200.{name} needs the pybind11 extension _tesseract_b200, which is not built "
                          f"(run `make -C tesseract_decoder_b200/csrc pybind`): {_ext_error}")
    raise AttributeError(name)
