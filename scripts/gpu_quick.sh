#!/bin/bash
# Shortest possible sanity session (torch-free, ~15 s on the box): smoke() against the CPU oracle, the committed golden
# vectors through the DEM's default kernel variant (first 128 shots of each set), one headline pass through the C ABI.
timeout ${QUICK_TIMEOUT:-50} python - <<'PY' > gpurun_out/quick.log 2>&1
import sys, time
sys.path.insert(0, "."); sys.path.insert(0, "tests")
t0 = time.time()
import __graft_entry__ as g
g.smoke()
import numpy as np, workloads as W
from tesseract_decoder_b200 import capi
bad = 0
for name in W.golden_names() + W.golden_names(sparsify=True):
    t = time.time()
    gold = W.load_golden(name)
    dem, kw = W.decoder_kwargs(gold["config"], capi.build_det_orders)
    dec = capi.Decoder(dem, device=0, **kw)
    n = min(128, len(gold["dets"]))
    out = dec.decode_batch_b8(gold["dets"][:n])
    ok = (np.array_equal(out["obs"], gold["obs"][:n]) and np.array_equal(out["cost"], gold["cost"][:n])
          and np.array_equal(out["low_confidence"], gold["low_confidence"][:n])
          and W.split_lists(out["err_offsets"], out["err_idx"]) == W.split_lists(gold["err_offsets"], gold["err_idx"])[:n])
    bad += not ok
    print(f"golden {name}: {n} shots, path {dec.search_path}, {'identical' if ok else 'MISMATCH'} ({time.time() - t:.1f} s)", flush=True)
    del dec
dem, kw = W.decoder_kwargs(W.CONFIGS["C4"], capi.build_det_orders)
dec = capi.Decoder(dem, device=0, **kw)
dec.set_collect_errors(False)
dets, obs = capi.sample_dem_b8(dem, 1000, 65536, dec.num_detectors, dec.num_observables)
for i in range(2):
    t = time.time(); out = dec.decode_batch_b8(dets); dt = time.time() - t
    tm = dec.last_batch_timing()
    print("bb144 65536 shots", round(65536 / dt), "shots/s e2e;", round(65536 / (tm["search_ms"] * 1e-3)), "shots/s search kernel;",
          "wrong", int(((out["obs"] != obs).any(axis=1) & (out["low_confidence"] == 0)).sum()), flush=True)
print("mismatching golden sets:", bad, "; wall", round(time.time() - t0, 1), "s")
sys.exit(1 if bad else 0)
PY
echo rc=$? >> gpurun_out/quick.log; cat gpurun_out/quick.log
