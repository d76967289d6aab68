"""Shared test/bench workload definitions: committed DEM fixtures, decoder flag sets, sampling.

The DEM fixtures under tests/golden/dems/ were produced by scripts/make_dems.py from the reference's
test circuits (the circuits themselves do not travel to the GPU box).
"""
from __future__ import annotations

import gzip
import json
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
INF_BEAM = 65535
CLI_ORDER_SEED = 518278944  # tesseract_main.cc:379


def load_dem(name: str) -> str:
    with gzip.open(os.path.join(GOLDEN, "dems", name + ".dem.gz")) as f:
        return f.read().decode()


# BASELINE.json configs, as decoder keyword sets.  pqlimit None = unbounded (CLI default).
CONFIGS = {
    # C1: surface d5 r5 p=0.002, --beam 5 --pqlimit 200000, CLI defaults otherwise (revisits allowed, 1 order)
    "C1": dict(dem="surface_d5_r5_p0.002_X", det_beam=5, pqlimit=200000, no_revisit_dets=False, num_det_orders=0),
    # C2: colour code d7, --num-det-orders 1 (1 DetIndex order), CLI defaults otherwise
    "C2": dict(dem="color_superdense_d7_r7_p0.001_X", det_beam=INF_BEAM, pqlimit=None, no_revisit_dets=False, num_det_orders=1),
    # C3: surface d11 r11 p=0.001, --beam 23 --beam-climbing
    "C3": dict(dem="surface_d11_r11_p0.001_X", det_beam=23, beam_climbing=True, pqlimit=None, no_revisit_dets=False, num_det_orders=0),
    # C4: BB [[144,12,12]] r12, --num-det-orders 24 --no-revisit-dets; p is not fixed by the config
    "C4": dict(dem="bb144_r12_p0.0001_X", det_beam=INF_BEAM, pqlimit=None, no_revisit_dets=True, num_det_orders=24),
    # C5: transversal-CNOT surface code, --det-penalty
    "C5": dict(dem="transcx_d5_p0.001_X", det_beam=INF_BEAM, pqlimit=None, no_revisit_dets=False, num_det_orders=0, det_penalty=1.0),
}


def decoder_kwargs(cfg: dict, build_det_orders) -> tuple[str, dict]:
    """(dem_text, kwargs for capi.Decoder / OracleDecoder) for a CONFIGS-style dict."""
    cfg = dict(cfg)
    dem = load_dem(cfg.pop("dem"))
    n_orders = cfg.pop("num_det_orders", 0)
    seed = cfg.pop("det_order_seed", CLI_ORDER_SEED)
    method = cfg.pop("det_order_method", 1)  # 0 = DetBFS, 1 = DetIndex, 2 = DetCoordinate (utils.h:43)
    kw = dict(det_beam=cfg.pop("det_beam", 5), beam_climbing=cfg.pop("beam_climbing", False),
              no_revisit_dets=cfg.pop("no_revisit_dets", True), merge_errors=cfg.pop("merge_errors", True),
              pqlimit=cfg.pop("pqlimit", 200000), det_penalty=cfg.pop("det_penalty", 0.0))
    for k in ("sparsify_errors", "sparsify_base_degree", "sparsify_max_degree", "sparsify_reactivate_limit",
              "at_most_two_errors_per_detector"):
        if k in cfg:
            kw[k] = cfg.pop(k)
    assert not cfg, f"unknown config keys {cfg}"
    kw["det_orders"] = build_det_orders(dem, n_orders, method, seed) if n_orders else None
    return dem, kw


def load_golden(name: str) -> dict:
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    out = {k: z[k] for k in z.files}
    out["config"] = json.loads(str(out["config"]))
    return out


def golden_names(sparsify: bool = False) -> list[str]:
    """Committed golden sets; names starting with "s_" decode with --sparsify-errors."""
    names = sorted(f[:-4] for f in os.listdir(GOLDEN) if f.endswith(".npz"))
    return [n for n in names if n.startswith("s_") == sparsify]


def split_lists(offsets, idx):
    return [idx[int(offsets[i]):int(offsets[i + 1])].tolist() for i in range(len(offsets) - 1)]
