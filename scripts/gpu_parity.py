"""Ad-hoc parity driver: GPU path (C ABI) vs the reference build on sampled shots.  Dev tool."""
import argparse, gzip, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tesseract_decoder_b200 import capi
from oracle.oracle import OracleDecoder

ap = argparse.ArgumentParser()
ap.add_argument("--dem", default="surface_d5_r5_p0.002_X")
ap.add_argument("--shots", type=int, default=1000)
ap.add_argument("--seed", type=int, default=1001)
ap.add_argument("--beam", type=int, default=5)
ap.add_argument("--pqlimit", type=int, default=200000)
ap.add_argument("--beam-climbing", action="store_true")
ap.add_argument("--revisit", action="store_true")
ap.add_argument("--orders", type=int, default=0)
ap.add_argument("--order-seed", type=int, default=518278944)
ap.add_argument("--det-penalty", type=float, default=0.0)
ap.add_argument("--threads", type=int, default=8)
ap.add_argument("--kind", default="reference")
ap.add_argument("--no-cpu", action="store_true")
a = ap.parse_args()

dem = gzip.open(os.path.join(ROOT, "tests/golden/dems", a.dem + ".dem.gz")).read().decode()
orders = capi.build_det_orders(dem, a.orders, 1, a.order_seed) if a.orders else None
kw = dict(det_beam=a.beam, beam_climbing=a.beam_climbing, no_revisit_dets=not a.revisit, pqlimit=a.pqlimit if a.pqlimit > 0 else None,
          det_penalty=a.det_penalty, det_orders=orders)
t0 = time.time(); g = capi.Decoder(dem, **kw); print("gpu create %.2fs" % (time.time() - t0))
dets, obs = capi.sample_dem_b8(dem, a.seed, a.shots, g.num_detectors, g.num_observables)
print("D", g.num_detectors, "E", g.num_errors, "mean fired", np.unpackbits(dets, axis=1).sum() / a.shots)
t0 = time.time(); out = g.decode_batch_b8(dets); tg = time.time() - t0
print("gpu decode %.3fs  %.1f shots/s" % (tg, a.shots / tg), g.last_batch_timing(), g.counters())
print("low_conf", int(out["low_confidence"].sum()), "logical errors", int(((out["obs"] != obs).any(axis=1) & (out["low_confidence"] == 0)).sum()))
if not a.no_cpu:
    r = OracleDecoder(dem, num_threads=a.threads, kind=a.kind, **kw)
    ref = r.decode_batch_b8(dets)
    print("cpu decode %.3fs  %.1f shots/s (threads=%d) sum_shot_s %.3f" % (ref["wall_seconds"], a.shots / ref["wall_seconds"], a.threads, ref["sum_shot_seconds"]))
    bad = 0
    for k in ("obs", "low_confidence"):
        n = int((out[k] != ref[k]).sum()); bad += n; print(k, "mismatches", n)
    rel = np.abs(out["cost"] - ref["cost"]) / np.maximum(np.abs(ref["cost"]), 1e-300)
    print("cost max rel diff", rel.max() if len(rel) else 0, "exact equal", int((out["cost"] == ref["cost"]).sum()), "/", a.shots)
    same_off = np.array_equal(out["err_offsets"], ref["err_offsets"]); same_idx = same_off and np.array_equal(out["err_idx"], ref["err_idx"])
    print("error lists identical:", same_idx)
    if not same_idx:
        nb = 0
        for s in range(a.shots):
            x = out["err_idx"][int(out["err_offsets"][s]):int(out["err_offsets"][s+1])]; y = ref["err_idx"][int(ref["err_offsets"][s]):int(ref["err_offsets"][s+1])]
            if not np.array_equal(x, y):
                nb += 1
                if nb <= 5: print(" shot", s, "gpu", x.tolist(), "ref", y.tolist(), out["cost"][s], ref["cost"][s])
        print(" shots with differing error lists:", nb); bad += nb
    print("PARITY", "OK" if bad == 0 and same_idx else "FAIL")
    sys.exit(0 if bad == 0 and same_idx else 1)
