"""Generates the committed DEM fixtures under tests/golden/dems/ from the reference's test circuits.

Runs only in the build container (needs /root/reference/testdata, which does not exist on the GPU
box).  The circuit -> DEM conversion is this repo's own analyzer (tesseract_decoder_b200/csrc/
circuit.cc), because stim is not available; the error counts it must reproduce are the reference's
own records in benchmarking/sparsify_errors/aggregated_results.jsonl (num_detectors /
num_compiled_errors), checked below.
"""
import gzip
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tesseract_decoder_b200 import capi  # noqa: E402

T = "/root/reference/testdata/"
# name -> (circuit path, expected (D, E) from the reference's records or None)
CIRCUITS = {
    "surface_d5_r5_p0.002_X": (T + "surfacecodes/r=5,d=5,p=0.002,noise=si1000,c=surface_code_X,q=49,gates=cz.stim", (120, 1677)),
    "color_superdense_d7_r7_p0.001_X": (T + "colorcodes/r=7,d=7,p=0.001,noise=si1000,c=superdense_color_code_X,q=73,gates=cz.stim", (252, 9941)),
    "color_superdense_d5_r5_p0.001_Z": (T + "colorcodes/r=5,d=5,p=0.001,noise=si1000,c=superdense_color_code_Z,q=37,gates=cz.stim", None),
    "surface_d11_r11_p0.001_X": (T + "surfacecodes/r=11,d=11,p=0.001,noise=si1000,c=surface_code_X,q=241,gates=cz.stim", (1320, 24483)),
    "bb144_r12_p0.001_X": (None, (1728, 67824)),
    "bb144_r12_p0.0005_X": (None, (1728, 67824)),
    "bb144_r12_p0.0001_X": (None, (1728, 67824)),
    "bb72_r6_p0.001_X": (None, (432, 16200)),
    "transcx_d3_p0.001_X": (None, None),
    "transcx_d5_p0.001_X": (None, None),
}


def find(sub, *needles):
    hits = [f for f in sorted(os.listdir(T + sub)) if all(n in f for n in needles)]
    assert len(hits) == 1, (sub, needles, hits)
    return T + sub + "/" + hits[0]


CIRCUITS["bb144_r12_p0.001_X"] = (find("bivariatebicyclecodes", "r=12,", "p=0.001,", "bicycle_X", "[[144,12,12]]"), (1728, 67824))
CIRCUITS["bb144_r12_p0.0005_X"] = (find("bivariatebicyclecodes", "r=12,", "p=0.0005,", "bicycle_X", "[[144,12,12]]"), (1728, 67824))
CIRCUITS["bb144_r12_p0.0001_X"] = (find("bivariatebicyclecodes", "r=12,", "p=0.0001,", "bicycle_X", "[[144,12,12]]"), (1728, 67824))
CIRCUITS["bb72_r6_p0.001_X"] = (find("bivariatebicyclecodes", "r=6,", "p=0.001,", "bicycle_X", "[[72,12,6]]"), (432, 16200))
CIRCUITS["transcx_d3_p0.001_X"] = (find("surface_code_trans_cx_circuits", "r=3,d=3,", "p=0.001,", "trans_cx_X"), None)
CIRCUITS["transcx_d5_p0.001_X"] = (find("surface_code_trans_cx_circuits", "r=5,d=5,", "p=0.001,", "trans_cx_X"), None)

out_dir = os.path.join(ROOT, "tests", "golden", "dems")
os.makedirs(out_dir, exist_ok=True)
for name, (path, expect) in CIRCUITS.items():
    dem = capi.circuit_to_dem(open(path).read())
    dec = capi.Decoder(dem, host_only=True)
    got = (dec.num_detectors, dec.num_errors)
    print(f"{name}: D={got[0]} E={got[1]} O={dec.num_observables} expect={expect} <- {os.path.basename(path)}")
    if expect is not None:
        assert got == expect, (name, got, expect)
    with gzip.GzipFile(os.path.join(out_dir, name + ".dem.gz"), "wb", mtime=0) as f:
        f.write(dem.encode())
