#!/bin/bash
# Surface code d=11, --beam 23 --beam-climbing (BASELINE configs[2]): throughput against batch size (the tail's share).  Torch-free.
timeout ${QUICK_TIMEOUT:-38} python - <<'PY' > gpurun_out/c3_batch.log 2>&1
import sys, time, json
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import numpy as np, workloads as W
from tesseract_decoder_b200 import capi
dem, kw = W.decoder_kwargs(W.CONFIGS["C3"], capi.build_det_orders)
dec = capi.Decoder(dem, device=0, **kw)
dec.set_collect_errors(False)
dets, obs = capi.sample_dem_b8(dem, 1000, 8192, dec.num_detectors, dec.num_observables)
dec.decode_batch_b8(dets[:256])
for n in (512, 1024, 2048, 4096, 8192):
    t = time.time(); out = dec.decode_batch_b8(dets[:n]); dt = time.time() - t
    tm = dec.last_batch_timing()
    print(json.dumps({"shots": n, "shots_per_s": round(n / dt, 1), "wall_ms": round(1e3 * dt, 1), "search_ms": round(tm["search_ms"], 1),
                      "tail_searches": tm["tail_searches"], "tail_ms": round(tm["tail_ms"], 1), "tail_cluster": tm["tail_cluster"],
                      "low_conf": int(out["low_confidence"].sum()),
                      "wrong": int(((out["obs"][:n] != obs[:n]).any(axis=1) & (out["low_confidence"] == 0)).sum())}), flush=True)
PY
echo rc=$? >> gpurun_out/c3_batch.log; cat gpurun_out/c3_batch.log
