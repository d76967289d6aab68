"""Per-phase time of the CTA-per-search kernel (tsr_profile) on the heavy-tailed BASELINE configs, and the wall time of
lone searches.  Run on the GPU box:  python scripts/profile_phases.py [case ...] > gpurun_out/phases.json
Each record: shots/s, search-kernel ms, SM-clock ticks per phase summed over CTAs (share in %), ticks per expansion."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402
import workloads as W  # noqa: E402
from tesseract_decoder_b200 import capi  # noqa: E402

LONGBEAM = dict(dem="bb144_r12_p0.001_X", det_beam=20, beam_climbing=True, pqlimit=1000000, no_revisit_dets=True, num_det_orders=21)
CASES = {
    # name: (config, sampler seed, shots sampled, slice)
    "c3_surface_d11_2048": (W.CONFIGS["C3"], 1000, 2048, slice(None)),
    "c3_shot1532_41755_expansions": (W.CONFIGS["C3"], 1000, 2048, slice(1532, 1533)),
    "c3_shot1007_2.5M_pushes": (W.CONFIGS["C3"], 1000, 2048, slice(1007, 1008)),
    "c2_color_d7_first600": (W.CONFIGS["C2"], 1002, 2000, slice(0, 600)),
    "c2_color_d7_shot6_305k_pushes": (W.CONFIGS["C2"], 1002, 2000, slice(6, 7)),
    "c2_color_d7_shot877_1.2M_pushes": (W.CONFIGS["C2"], 1002, 2000, slice(877, 878)),
    "c2_color_d7_shot631_35.6M_pushes": (W.CONFIGS["C2"], 1002, 2000, slice(631, 632)),
    "c2_color_d7_all2000": (W.CONFIGS["C2"], 1002, 2000, slice(None)),
    "bb144_p1e-3_longbeam_256": (LONGBEAM, 1000, 256, slice(None)),
    "bb144_p1e-3_longbeam_one_shot": (LONGBEAM, 1000, 256, slice(0, 1)),
    "c4_bb144_p1e-4_65536": (W.CONFIGS["C4"], 1000, 65536, slice(None)),
    "c1_surface_d5_262144": (W.CONFIGS["C1"], 1000, 262144, slice(None)),
    "c5_transcx_d5_131072": (W.CONFIGS["C5"], 1000, 131072, slice(None)),
}


def run(name, extra=None):
    cfg, seed, n, sl = CASES[name]
    dem, kw = W.decoder_kwargs(cfg, capi.build_det_orders)
    dec = capi.Decoder(dem, **kw, **(extra or {}))
    dets, _ = capi.sample_dem_b8(dem, seed, n, dec.num_detectors, dec.num_observables)
    dets = np.ascontiguousarray(dets[sl])
    dec.set_collect_errors(False)
    dec.decode_batch_b8(dets[: min(len(dets), 64)], want_errors=False)  # warm-up (tables into L2, lazy init)
    rec = {"case": name, "shots": len(dets), "path": dec.search_path, "extra": extra or {}}
    for prof in (False, True):
        dec.set_profile(prof)
        dec.profile(reset=True)
        dec.counters(reset=True)
        t0 = time.perf_counter()
        dec.decode_batch_b8(dets, want_errors=False)
        wall = time.perf_counter() - t0
        t = dec.last_batch_timing()
        if not prof:
            rec.update(wall_s=round(wall, 4), search_ms=round(t["search_ms"], 3), shots_per_s=round(len(dets) / wall, 1),
                       launches=t["search_launches"], rerun=t["rerun_items"], counters=dec.counters(),
                       tail=dict(searches=t["tail_searches"], ms=round(t["tail_ms"], 2), cluster=t["tail_cluster"]))
        else:
            p = dec.profile()
            ticks = sum(v for k, v in p.items() if k != "expansions")
            rec["profiled_search_ms"] = round(t["search_ms"], 3)
            rec["phase_pct"] = {k: round(100.0 * v / max(ticks, 1), 1) for k, v in p.items() if k != "expansions"}
            rec["ticks_per_expansion"] = round(ticks / max(p["expansions"], 1))
            rec["expansions"] = p["expansions"]
    print(json.dumps(rec), flush=True)


if __name__ == "__main__":
    names = sys.argv[1:] or list(CASES)
    for nm in names:
        try:
            run(nm)
        except Exception as e:  # keep going: one failing case should not hide the others
            print(json.dumps({"case": nm, "error": f"{type(e).__name__}: {e}"}), flush=True)
