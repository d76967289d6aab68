"""The `tesseract` CLI (csrc/main.cc): the reference CLI's decode flags, validation and accounting
(tesseract_main.cc:98-180, 350-538, 592-616, 657-699)."""
import json
import os
import subprocess

import numpy as np
import pytest

import workloads as W
from tesseract_decoder_b200 import capi

CLI = os.path.join(W.ROOT, "tesseract_decoder_b200", "tesseract")


def run(*args):
    return subprocess.run([CLI, *map(str, args)], capture_output=True, text=True, timeout=600)


def test_argument_validation():
    assert os.path.exists(CLI), "build the CLI with `make -C tesseract_decoder_b200/csrc cli`"
    r = run("--sample-num-shots", 5)
    assert r.returncode != 0 and "exactly one of --circuit or --dem" in r.stderr
    r = run("--dem", "x.dem")
    assert r.returncode != 0 and "exactly one of --sample-num-shots or --in" in r.stderr
    r = run("--dem", "x.dem", "--sample-num-shots", 5, "--beam-climbing")  # tesseract_main.cc:151-153
    assert r.returncode != 0 and "Beam climbing requires a finite beam" in r.stderr
    r = run("--dem", "x.dem", "--sample-num-shots", 5, "--bogus")
    assert r.returncode != 0 and "unknown argument" in r.stderr
    r = run("--dem", "/nonexistent.dem", "--sample-num-shots", 5)
    assert r.returncode != 0 and "Failed to open" in r.stderr
    dem = os.path.join(W.GOLDEN, "..", "..", "gpurun_out", "_cli_test.dem")
    os.makedirs(os.path.dirname(dem), exist_ok=True)
    with open(dem, "w") as f:
        f.write("error(0.1) D0 L0\nerror(0.1) D0 D1")
    r = run("--dem", dem, "--sample-num-shots", 5, "--sparsify-errors", "--sparsify-base-degree", 0)
    assert r.returncode != 0 and "--sparsify-base-degree must be > 0." in r.stderr  # tesseract_main.cc:170-172
    # Args::validate of the reference (tesseract_main.cc:155-179)
    r = run("--dem", dem, "--sample-num-shots", 5, "--sparsify-base-degree", 2)
    assert r.returncode != 0 and "without --sparsify-errors" in r.stderr
    r = run("--dem", dem, "--sample-num-shots", 5, "--sparsify-errors")
    assert r.returncode != 0 and "Must specify --sparsify-base-degree when --sparsify-errors is enabled." in r.stderr
    r = run("--dem", dem, "--sample-num-shots", 5, "--sparsify-errors", "--sparsify-base-degree", 3, "--sparsify-max-degree", 2)
    assert r.returncode != 0 and "--sparsify-max-degree must be >= --sparsify-base-degree." in r.stderr
    # the rest of Args::validate (tesseract_main.cc:103-147), all before any device is touched
    for args, message in ((("--det-order-bfs", "--det-order-index"), "Only one of --det-order-bfs, --det-order-index, or --det-order-coordinate may be set."),
                          (("--out-format", "foo"), "Invalid format: foo"),
                          (("--obs_in", "x.b8"), "Cannot load observable flips without a corresponding detection event data file."),
                          (("--threads", 0), "--threads must be at least 1."),
                          (("--shot-range-begin", 9, "--shot-range-end", 3), "Provided shot range must have end >= begin."),
                          (("--beam", "abc"), "--beam: 'abc' is not an integer"),
                          (("--pqlimit", -5), "--pqlimit: '-5' is not an unsigned integer"),
                          (("--det-penalty", "1.0x"), "--det-penalty: '1.0x' is not a number")):
        r = run("--dem", dem, "--sample-num-shots", 5, *args)
        assert r.returncode != 0 and message in r.stderr, (args, r.stderr)


@pytest.mark.gpu
def test_cli_matches_the_library(tmp_path):
    dem = W.load_dem("surface_d5_r5_p0.002_X")
    (tmp_path / "m.dem").write_text(dem)
    # sampled shots: counts equal the library's on the same sampler seed
    r = run("--dem", tmp_path / "m.dem", "--sample-num-shots", 3000, "--sample-seed", 7, "--beam", 5, "--pqlimit", 200000,
            "--stats-out", tmp_path / "stats.json")
    assert r.returncode == 0, r.stderr
    dec = capi.Decoder(dem, det_beam=5, pqlimit=200000, no_revisit_dets=False, det_orders=capi.build_det_orders(dem, 1, 1, W.CLI_ORDER_SEED))
    dets, obs = capi.sample_dem_b8(dem, 7, 3000, 120, 1)
    out = dec.decode_batch_b8(dets)
    wrong = int(((out["obs"] != obs).any(axis=1) & (out["low_confidence"] == 0)).sum())
    low = int(out["low_confidence"].sum())
    assert r.stdout.startswith(f"num_shots = 3000 num_low_confidence = {low} num_errors = {wrong} total_time_seconds = ")
    stats = json.loads((tmp_path / "stats.json").read_text())
    assert (stats["num_shots"], stats["num_errors"], stats["num_low_confidence"], stats["det_beam"], stats["pqlimit"]) == \
        (3000, wrong, low, 5, 200000)
    # shots from files, predictions to a file, in both formats
    dets.tofile(tmp_path / "dets.b8")
    obs.tofile(tmp_path / "obs.b8")
    r = run("--dem", tmp_path / "m.dem", "--in", tmp_path / "dets.b8", "--obs_in", tmp_path / "obs.b8", "--beam", 5,
            "--pqlimit", 200000, "--out", tmp_path / "pred.01", "--out-format", "01", "--shot-range-begin", 100, "--shot-range-end", 600)
    assert r.returncode == 0, r.stderr
    lines = (tmp_path / "pred.01").read_text().split()
    assert len(lines) == 500 and [int(x) for x in lines] == out["obs"][100:600, 0].tolist()
    bits01 = "\n".join("".join(str(b) for b in row) for row in np.unpackbits(dets[:50], axis=1, bitorder="little")[:, :120]) + "\n"
    (tmp_path / "dets.01").write_text(bits01)
    r = run("--dem", tmp_path / "m.dem", "--in", tmp_path / "dets.01", "--in-format", "01", "--beam", 5, "--pqlimit", 200000,
            "--out", tmp_path / "pred.b8")
    assert r.returncode == 0, r.stderr
    assert np.fromfile(tmp_path / "pred.b8", dtype=np.uint8).tolist() == out["obs"][:50, 0].tolist()
    assert "num_errors" not in r.stdout  # no observables given (tesseract_main.cc:693-695)
    # --dem-out: usage-frequency DEM over the shots whose prediction matched (tesseract_main.cc:586-590, 625-639)
    r = run("--dem", tmp_path / "m.dem", "--in", tmp_path / "dets.b8", "--obs_in", tmp_path / "obs.b8", "--beam", 5,
            "--pqlimit", 200000, "--dem-out", tmp_path / "use.dem")
    assert r.returncode == 0, r.stderr
    use = np.zeros(dec.num_dem_errors, dtype=np.int64)
    lists = W.split_lists(out["err_offsets"], out["err_idx"])
    good = (out["obs"] == obs).all(axis=1)
    for s_, errs in enumerate(lists):
        if good[s_]:
            use[errs] += 1
    assert (tmp_path / "use.dem").read_text().strip() == capi.dem_from_counts(dem, use, 3000 - wrong).strip()


def test_shot_file_formats_round_trip(tmp_path):
    """stim result formats the CLI reads and writes (--in-format / --out-format): b8, 01, hits, dets, r8, ptb64.
    Host-only (--convert-only): every format re-encodes to the same b8 bytes."""
    dem = W.load_dem("surface_d5_r5_p0.002_X")
    (tmp_path / "m.dem").write_text(dem)
    dets, _ = capi.sample_dem_b8(dem, 3, 200, 120, 1)
    dets[5] = 0       # an empty shot
    dets[6] = 0xFF    # a full one (120 ones: the last byte is entirely detector bits)
    dets.tofile(tmp_path / "a.b8")
    for fmt in ("01", "hits", "dets", "r8"):
        r = run("--convert-only", "--dem", tmp_path / "m.dem", "--in", tmp_path / "a.b8", "--out", tmp_path / f"a.{fmt}", "--out-format", fmt)
        assert r.returncode == 0, r.stderr
        r = run("--convert-only", "--dem", tmp_path / "m.dem", "--in", tmp_path / f"a.{fmt}", "--in-format", fmt, "--out",
                tmp_path / f"back_{fmt}.b8", "--out-format", "b8")
        assert r.returncode == 0, r.stderr
        assert np.array_equal(np.fromfile(tmp_path / f"back_{fmt}.b8", dtype=np.uint8).reshape(200, 15), dets), fmt
    assert (tmp_path / "a.hits").read_text().splitlines()[5] == ""
    assert (tmp_path / "a.dets").read_text().splitlines()[5] == "shot"
    first = np.nonzero(np.unpackbits(dets[0], bitorder="little")[:120])[0]
    assert (tmp_path / "a.dets").read_text().splitlines()[0] == "shot" + "".join(f" D{i}" for i in first)
    raw = (tmp_path / "a.r8").read_bytes()
    assert raw[0] == first[0]  # run of zeros before the first one
    # ptb64 (tesseract_main.cc:259-330 accepts every stim format name): groups of 64 shots, transposed — for each bit k
    # one little-endian u64 whose bit s is bit k of the group's shot s.  Checked against numpy, and round-tripped.
    r = run("--convert-only", "--dem", tmp_path / "m.dem", "--in", tmp_path / "a.b8", "--out", tmp_path / "x", "--out-format", "ptb64")
    assert r.returncode != 0 and "multiple of 64" in r.stderr  # 200 shots
    dets[:192].tofile(tmp_path / "b.b8")
    r = run("--convert-only", "--dem", tmp_path / "m.dem", "--in", tmp_path / "b.b8", "--out", tmp_path / "b.ptb64", "--out-format", "ptb64")
    assert r.returncode == 0, r.stderr
    bits = np.unpackbits(dets[:192], axis=1, bitorder="little")[:, :120]          # [shot, bit]
    expect = np.packbits(bits.reshape(3, 64, 120).transpose(0, 2, 1), axis=2, bitorder="little")  # [group, bit, 8 bytes]
    assert (tmp_path / "b.ptb64").read_bytes() == expect.tobytes()
    r = run("--convert-only", "--dem", tmp_path / "m.dem", "--in", tmp_path / "b.ptb64", "--in-format", "ptb64", "--out",
            tmp_path / "back_ptb64.b8", "--out-format", "b8")
    assert r.returncode == 0, r.stderr
    assert np.array_equal(np.fromfile(tmp_path / "back_ptb64.b8", dtype=np.uint8).reshape(192, 15), dets[:192])
    r = run("--convert-only", "--dem", tmp_path / "m.dem", "--in", tmp_path / "a.b8", "--out", tmp_path / "x", "--out-format", "nope")
    assert r.returncode != 0 and "Invalid format: nope" in r.stderr  # tesseract_main.cc:119-127


def test_malformed_shot_files_are_refused(tmp_path):
    """Every reader fails with a message naming the file (never a crash, never a silently shortened batch); the same
    cases run clean under ASan + UBSan (scripts/host_sanitizers.sh builds the instrumented library)."""
    dem = W.load_dem("surface_d5_r5_p0.002_X")
    (tmp_path / "m.dem").write_text(dem)
    rng = np.random.default_rng(0)
    cases = {
        "b8 truncated": ("b8", bytes(15 * 3 - 7), "multiple of the b8 shot size"),
        "01 short line": ("01", b"0101\n", "does not have 120 bits"),
        "01 long line": ("01", b"1" * 300 + b"\n", "does not have 120 bits"),
        "01 bad character": ("01", b"01x1" * 30 + b"\n", "unexpected character"),
        "hits out of range": ("hits", b"1,5,100000\n", "out of range"),
        "hits negative": ("hits", b"-3,4\n", "not a bit index"),
        "hits trailing junk": ("hits", b"12abc\n", "not a bit index"),
        "hits overflow": ("hits", b"99999999999999999999999\n", "not a bit index"),
        "dets out of range": ("dets", b"shot D5 D99999\n", "out of range"),
        "dets bad token": ("dets", b"shot Dx\n", "not a bit index"),
        "dets no shot": ("dets", b"D1 D2\n", "must start with 'shot'"),
        "r8 truncated": ("r8", bytes([255, 255, 255]), "ends inside a shot"),
        "r8 overrun": ("r8", bytes([200, 200, 200, 200]), "passes the end of a shot"),
        "r8 random": ("r8", rng.integers(0, 256, 1000, dtype=np.uint8).tobytes(), "r8"),
        "ptb64 truncated": ("ptb64", b"\x01" * 100, "multiple of the ptb64 group size"),
    }
    for name, (fmt, data, message) in cases.items():
        (tmp_path / "x.in").write_bytes(data)
        r = run("--convert-only", "--dem", tmp_path / "m.dem", "--in", tmp_path / "x.in", "--in-format", fmt, "--out", tmp_path / "x.out",
                "--out-format", "b8")
        assert r.returncode == 1 and message in r.stderr and "x.in" in r.stderr, (name, r.returncode, r.stderr)
    # lenient where stim is: blanks around hit indices, an empty file is zero shots
    (tmp_path / "x.in").write_bytes(b" 3, 4 \n")
    r = run("--convert-only", "--dem", tmp_path / "m.dem", "--in", tmp_path / "x.in", "--in-format", "hits", "--out", tmp_path / "x.out",
            "--out-format", "hits")
    assert r.returncode == 0 and (tmp_path / "x.out").read_text() == "3,4\n"
    (tmp_path / "x.in").write_bytes(b"")
    r = run("--convert-only", "--dem", tmp_path / "m.dem", "--in", tmp_path / "x.in", "--in-format", "01", "--out", tmp_path / "x.out",
            "--out-format", "b8")
    assert r.returncode == 0 and (tmp_path / "x.out").read_bytes() == b""
