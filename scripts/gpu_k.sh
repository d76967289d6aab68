#!/bin/bash
tag=$1
timeout 600 python -m pytest tests -m gpu -q --durations=8 > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${tag}_pytest.log; tail -4 gpurun_out/${tag}_pytest.log
timeout 400 python scripts/profile_phases.py c3_surface_d11_2048 c2_color_d7_first600 bb144_p1e-3_longbeam_256 c2_color_d7_shot877_1.2M_pushes c2_color_d7_shot631_35.6M_pushes c1_surface_d5_262144 c5_transcx_d5_131072 c4_bb144_p1e-4_65536 > gpurun_out/${tag}_phases.jsonl 2> gpurun_out/${tag}_phases.err
cut -c1-520 gpurun_out/${tag}_phases.jsonl
python - <<'PY'
import sys; sys.path.insert(0,'tests'); sys.path.insert(0,'.')
import numpy as np, workloads as W, time
from tesseract_decoder_b200 import capi
dem, kw = W.decoder_kwargs(W.CONFIGS['C3'], capi.build_det_orders)
dec = capi.Decoder(dem, **kw)
dets,_ = capi.sample_dem_b8(dem, 1000, 8192, dec.num_detectors, dec.num_observables)
for n in (2048, 8192):
    dec.decode_batch_b8(dets[:64], want_errors=False)
    t0=time.perf_counter(); dec.decode_batch_b8(dets[:n], want_errors=False); dt=time.perf_counter()-t0
    print('C3 shots', n, 'rate', n/dt, dec.last_batch_timing())
PY
