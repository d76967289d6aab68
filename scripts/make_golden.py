"""Generates tests/golden/*.npz: seeded shots plus the outputs of the REFERENCE's own decoder on them.

Runs in the build container only (it builds oracle/_ref from /root/reference).  Each file holds the
bit-packed shots, the decoder flags, and what the unmodified reference produced: per-shot observable
bits, cost_from_errors, low_confidence and predicted_errors_buffer.  Parity tests compare both the
CPU port and the CUDA path against these.
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import oracle  # noqa: E402
from tesseract_decoder_b200 import capi  # noqa: E402
import workloads as W  # noqa: E402

oracle.build("reference")
SETS = {
    # name: (config, shots, seed)
    "g_c1_surface_d5": (W.CONFIGS["C1"], 400, 1001),
    "g_c1_surface_d5_norevisit_20orders": (dict(dem="surface_d5_r5_p0.002_X", det_beam=5, pqlimit=200000, no_revisit_dets=True,
                                                num_det_orders=20, det_order_seed=2384753), 200, 11),
    "g_c2_color_d7_beam5": (dict(dem="color_superdense_d7_r7_p0.001_X", det_beam=5, pqlimit=200000, no_revisit_dets=True,
                                 num_det_orders=1), 150, 1002),
    "g_color_d5_climb": (dict(dem="color_superdense_d5_r5_p0.001_Z", det_beam=3, beam_climbing=True, pqlimit=1000,
                              no_revisit_dets=True, num_det_orders=5), 200, 5),
    # BASELINE config 2 at its literal flags (CLI defaults: beam 65535, pqlimit unbounded, revisits allowed;
    # tesseract_main.cc:482, 495-503).  Heavy-tailed: median 837 pushes per shot, but this batch holds searches of
    # 35.6 M (shot 631), 11.0 M (1296), 4.4 M (1395) ... pushes; the reference needs 226 s for shot 631 alone.
    "g_c2_color_d7_literal": (W.CONFIGS["C2"], 2000, 1002),
    "g_c3_surface_d11_climb": (W.CONFIGS["C3"], 256, 1003),
    "g_c4_bb144_p1e-4": (W.CONFIGS["C4"], 256, 1004),
    "g_bb72_pqlimit2000": (dict(dem="bb72_r6_p0.001_X", det_beam=65535, pqlimit=2000,
                                no_revisit_dets=True, num_det_orders=2), 64, 77),
    "g_c5_transcx_d5_penalty": (W.CONFIGS["C5"], 200, 1005),
    "g_transcx_d3_nomerge": (dict(dem="transcx_d3_p0.001_X", det_beam=5, pqlimit=200000, no_revisit_dets=True, merge_errors=False,
                                  num_det_orders=0, det_penalty=0.25), 300, 9),
    # --sparsify-errors (tesseract.cc:256-292, 673-737): reference outputs for the next hot-path row
    "s_color_d5_base3": (dict(dem="color_superdense_d5_r5_p0.001_Z", det_beam=5, pqlimit=200000, no_revisit_dets=True,
                              num_det_orders=3, sparsify_errors=True, sparsify_base_degree=3), 200, 41),
    "s_surface_d5_base2_limit8": (dict(dem="surface_d5_r5_p0.002_X", det_beam=5, pqlimit=200000, no_revisit_dets=True,
                                       num_det_orders=0, sparsify_errors=True, sparsify_base_degree=2, sparsify_max_degree=4,
                                       sparsify_reactivate_limit=8), 300, 42),
    "s_bb72_base3_climb": (dict(dem="bb72_r6_p0.001_X", det_beam=3, beam_climbing=True, pqlimit=20000, no_revisit_dets=True,
                                num_det_orders=4, sparsify_errors=True, sparsify_base_degree=3), 64, 43),
}
only = sys.argv[1:]
for name, (cfg, shots, seed) in SETS.items():
    if only and name not in only:
        continue
    dem, kw = W.decoder_kwargs(cfg, capi.build_det_orders)
    host = capi.Decoder(dem, host_only=True, **{k: v for k, v in kw.items() if not k.startswith("sparsify")})
    dets, obs = capi.sample_dem_b8(dem, seed, shots, host.num_detectors, host.num_observables)
    ref = oracle.OracleDecoder(dem, kind="reference", num_threads=8, **kw)
    out = ref.decode_batch_b8(dets)
    np.savez_compressed(os.path.join(W.GOLDEN, name + ".npz"), config=json.dumps(cfg), seed=seed, dets=dets, sampled_obs=obs,
                        obs=out["obs"], cost=out["cost"], low_confidence=out["low_confidence"],
                        err_offsets=out["err_offsets"], err_idx=out["err_idx"])
    wrong = int(((out["obs"] != obs).any(axis=1) & (out["low_confidence"] == 0)).sum())
    print(f"{name}: shots={shots} low_conf={int(out['low_confidence'].sum())} logical_errors={wrong} "
          f"mean_errs={len(out['err_idx']) / shots:.2f} ref_wall={out['wall_seconds']:.2f}s")
