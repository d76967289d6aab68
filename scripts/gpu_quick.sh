#!/bin/bash
# Shortest possible sanity session (torch-free): smoke() against the CPU oracle, then one headline pass through the C ABI.
timeout 60 python - <<'PY' > gpurun_out/quick.log 2>&1
import sys, time
sys.path.insert(0, "."); sys.path.insert(0, "tests")
t0 = time.time()
import __graft_entry__ as g
g.smoke()
print("smoke wall", round(time.time() - t0, 1), flush=True)
import numpy as np, workloads as W
from tesseract_decoder_b200 import capi
dem, kw = W.decoder_kwargs(W.CONFIGS["C4"], capi.build_det_orders)
dec = capi.Decoder(dem, device=0, **kw)
dec.set_collect_errors(False)
dets, obs = capi.sample_dem_b8(dem, 1000, 65536, dec.num_detectors, dec.num_observables)
for i in range(3):
    t = time.time(); out = dec.decode_batch_b8(dets); dt = time.time() - t
    tm = dec.last_batch_timing()
    print("bb144 65536 shots", round(65536 / dt), "shots/s e2e;", round(65536 / (tm["search_ms"] * 1e-3)), "shots/s search kernel;",
          "wrong", int(((out["obs"] != obs).any(axis=1) & (out["low_confidence"] == 0)).sum()), flush=True)
PY
echo rc=$? >> gpurun_out/quick.log; cat gpurun_out/quick.log
