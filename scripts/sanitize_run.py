"""Small decode through every search-kernel variant, for compute-sanitizer (memcheck / racecheck) runs:
    compute-sanitizer --tool memcheck python scripts/sanitize_run.py
Checks the answers against the committed golden vectors on the way."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import workloads as W  # noqa: E402
from tesseract_decoder_b200 import capi  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 16
for name in ("g_c1_surface_d5_norevisit_20orders", "g_bb72_pqlimit2000", "g_color_d5_climb", "g_c4_bb144_p1e-4"):
    g = W.load_golden(name)
    dem, kw = W.decoder_kwargs(g["config"], capi.build_det_orders)
    for variant, opts in (("bits", dict(row_scan=2)), ("lanes", dict(row_scan=1)), ("generic", dict(force_generic_kernel=True))):
        dec = capi.Decoder(dem, search_slots=8, node_pool_bytes=256 << 20, **opts, **kw)
        if variant == "lanes":
            dec.set_instrumented(True)
        out = dec.decode_batch_b8(g["dets"][:n])
        ok = np.array_equal(out["obs"], g["obs"][:n]) and np.array_equal(out["cost"], g["cost"][:n])
        print(name, variant, dec.search_path, "ok" if ok else "MISMATCH", flush=True)
        assert ok
        dec.close()
print("sanitize_run done")
