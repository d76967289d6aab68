"""One warm-up decode (64 shots) and one decode of a named case — the process `ncu` wraps (scripts/gpu_ncu.sh).
usage: python scripts/one_decode.py <case of scripts/profile_phases.py> [shots]"""
import sys
import os
import time
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import numpy as np
import profile_phases as P
from tesseract_decoder_b200 import capi
import workloads as W

name = sys.argv[1]
cfg, seed, n, sl = P.CASES[name]
dem, kw = W.decoder_kwargs(cfg, capi.build_det_orders)
dec = capi.Decoder(dem, **kw)
dets, _ = capi.sample_dem_b8(dem, seed, n, dec.num_detectors, dec.num_observables)
dets = np.ascontiguousarray(dets[sl])
if len(sys.argv) > 2:
    dets = dets[: int(sys.argv[2])]
dec.set_collect_errors(False)
dec.decode_batch_b8(dets[:64], want_errors=False)
t0 = time.perf_counter()
dec.decode_batch_b8(dets, want_errors=False)
print(f"{name}: {len(dets)} shots in {time.perf_counter() - t0:.3f} s, search kernel {dec.last_batch_timing()['search_ms']:.1f} ms, path {dec.search_path}")
