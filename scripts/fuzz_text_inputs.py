"""Mutation fuzzing of the text entry points of the host library (DEM parser + every helper that takes DEM text, the
circuit analyzer) — meant to run against the ASan + UBSan build that scripts/host_sanitizers.sh leaves in /tmp/tsr_san:

    bash scripts/host_sanitizers.sh          # builds /tmp/tsr_san/libtesseract_b200.so, runs the host tests under it
    ASAN_OPTIONS=detect_leaks=0:allocator_may_return_null=1 UBSAN_OPTIONS=halt_on_error=1 \\
      LD_PRELOAD="$(gcc -print-file-name=libasan.so) $(gcc -print-file-name=libubsan.so)" \\
      python scripts/fuzz_text_inputs.py dem 0 800       # or: circuit 0 400 (needs /root/reference/testdata for the seed circuit)

Every call must either succeed or raise ValueError / RuntimeError; a sanitizer report or a crash is a bug."""
import glob
import os
import random
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from tesseract_decoder_b200 import capi  # noqa: E402

if os.path.exists("/tmp/tsr_san/libtesseract_b200.so"):
    capi.LIB_PATH = "/tmp/tsr_san/libtesseract_b200.so"
import workloads as W  # noqa: E402

REFUSALS = (ValueError, RuntimeError, UnicodeError, MemoryError)
DEM_TOKENS = ["error(", ")", "D", "L", "^", "repeat", "{", "}", "shift_detectors", "detector(", "logical_observable", "1e400", "-1", "nan",
              "inf", "0.5", "99999999999999999999", "4294967296", "\n", " ", ",", "(", "#", "[tag]", "D-1", "L99999999",
              "repeat 1000000000 {", "1.5", "\x00", "\u00e9"]
CIRCUIT_TOKENS = ["REPEAT 3 {", "}", "DETECTOR", "rec[-1]", "rec[-99999]", "rec[0]", "OBSERVABLE_INCLUDE(0)", "OBSERVABLE_INCLUDE(99999999)",
                  "M", "MR", "MRX", "R", "H", "CZ", "CX", "DEPOLARIZE1(0.5)", "DEPOLARIZE2(2)", "X_ERROR(-1)", "PAULI_CHANNEL_1(0.1,0.2)",
                  "TICK", "QUBIT_COORDS(1,2)", "SHIFT_COORDS(0,0,1)", "99999999999", "-5", "\n", " ", "(", ")", "REPEAT 999999999 {",
                  "MPP X1*Z2", "E(0.1) X1", "ELSE_CORRELATED_ERROR(0.1) Z1", "4294967295", "#", "\x00"]


def mutate(text, tokens, span):
    t = list(text)
    for _ in range(random.randint(1, 6)):
        k, i = random.random(), random.randrange(len(t) + 1)
        if k < 0.3 and t:
            del t[i % len(t): i % len(t) + random.randint(1, span)]
        elif k < 0.6:
            t[i:i] = list(" " + random.choice(tokens) + " ")
        elif k < 0.8 and t:
            t[i % len(t)] = chr(random.randrange(32, 127))
        else:
            t = t[:i]
    return "".join(t)


def main():
    what, seed, n = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
    random.seed(seed)
    ok = bad = 0
    if what == "dem":
        base = "\n".join(W.load_dem("surface_d5_r5_p0.002_X").splitlines()[:60]) + \
            "\nrepeat 2 {\n error(0.1) D1 D2 ^ D3 L0\n shift_detectors(1, 2) 3\n detector(1,2,3) D0\n}\nlogical_observable L1\n"
        for it in range(n):
            text = mutate(base, DEM_TOKENS, 20)
            for fn in (capi.dem_count_detectors, capi.merge_indistinguishable_errors, capi.remove_zero_probability_errors,
                       capi.get_errors_from_dem, capi.build_detector_graph, capi.get_detector_coords,
                       lambda t: capi.build_det_orders(t, 2, it % 3, 1), lambda t: capi.Decoder(t, host_only=True).close()):
                try:
                    fn(text)
                    ok += 1
                except REFUSALS:
                    bad += 1
    else:
        base = open(sorted(glob.glob("/root/reference/testdata/surfacecodes/r=5,d=5,p=0.002*"))[0]).read()
        for _ in range(n):
            try:
                capi.circuit_to_dem(mutate(base, CIRCUIT_TOKENS, 40))
                ok += 1
            except REFUSALS:
                bad += 1
    print(f"{what}: {ok} calls accepted, {bad} refused, library {capi.LIB_PATH}")


if __name__ == "__main__":
    main()
