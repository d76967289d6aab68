"""GPU parity proper (-m gpu): the sm_100a decode path, called through the C ABI
(include/tesseract_b200.h via tesseract_decoder_b200/capi.py), against

  * the committed golden outputs of the reference build (tests/golden/*.npz, scripts/make_golden.py),
  * the reference's own known-answer tests restated (tesseract.test.cc, shared_decoding_tests.py),
  * the CPU checkers (oracle/_ref when it travelled, else the port) on fresh seeded shots, and
  * size-independent properties at BASELINE sizes (syndrome reproduction, cost/observable
    consistency, batch-shape / slot-count / node-pool-tier invariance, determinism).

Bar: bit-exact observables, low-confidence flags, error lists (order included) and costs
(np.array_equal on float64 — stricter than the 1e-9 relative the north star allows).
Nothing here reads /root/reference.
"""
import math

import numpy as np
import pytest

import workloads as W
from oracle import oracle
from oracle.oracle import OracleDecoder
from tesseract_decoder_b200 import capi
from test_oracle import EXHAUSTIVE_DEMS, brute_force_min_cost

pytestmark = pytest.mark.gpu
KIND = "reference" if oracle.available("reference") else "port"


def assert_same(out, ref, n=None):
    n = len(out["obs"]) if n is None else n
    assert np.array_equal(out["low_confidence"][:n], ref["low_confidence"][:n])
    assert np.array_equal(out["obs"][:n], ref["obs"][:n])
    assert np.array_equal(out["cost"][:n], ref["cost"][:n])  # bit-exact doubles
    a = W.split_lists(out["err_offsets"], out["err_idx"])[:n]
    b = W.split_lists(ref["err_offsets"], ref["err_idx"])[:n]
    bad = [i for i in range(n) if a[i] != b[i]]
    assert not bad, f"{len(bad)} shots with differing error lists, first {bad[:5]}"


def check_solutions_are_consistent(dec, dets, out):
    """Properties every non-low-confidence answer must have, whatever the search did: the predicted
    errors reproduce the syndrome, cost == cost_from_errors, obs == get_flipped_observables."""
    D, O = dec.num_detectors, dec.num_observables
    d2e, _ = dec.maps()
    off, edets = dec.csr(1)
    ooff, eobs = dec.csr(3)
    costs = dec.error_costs()
    bits = np.unpackbits(dets, axis=1, bitorder="little")[:, :D]
    obits = np.unpackbits(out["obs"], axis=1, bitorder="little")[:, :O] if O else np.zeros((len(dets), 0), np.uint8)
    lists = W.split_lists(out["err_offsets"], out["err_idx"])[:len(dets)]
    for s, errs in enumerate(lists):
        if out["low_confidence"][s]:
            assert errs == [] and out["cost"][s] == 0
            continue
        syn = np.zeros(D, dtype=np.uint8)
        ob = np.zeros(O, dtype=np.uint8)
        total = 0.0
        for e in errs:
            ce = int(d2e[e])
            assert ce != capi.NO_ERROR_INDEX
            syn[edets[int(off[ce]):int(off[ce + 1])]] ^= 1
            ob[eobs[int(ooff[ce]):int(ooff[ce + 1])]] ^= 1
            total += costs[ce]
        assert np.array_equal(syn, bits[s]), s
        assert np.array_equal(ob, obits[s]), s
        assert total == out["cost"][s], s


# ---- golden vectors of the reference build ---------------------------------------------------------

# golden set -> (kernel variant that decodes all of it, shots the other variants decode)
HEAVY_GOLDEN = {"g_c2_color_d7_literal": ("fast", 600)}


@pytest.mark.parametrize("path", ["fast-bits", "fast", "generic"])
@pytest.mark.parametrize("name", W.golden_names() + W.golden_names(sparsify=True))
def test_golden_vectors(name, path):
    """All three CUDA search variants — the CTA-per-search integer-rank kernel every fixture DEM is certified for
    (csrc/fast.h) with bit-sliced and with entry-per-lane row scans, and the generic event-replay kernel —
    reproduce the reference build's outputs.  The s_* sets decode with --sparsify-errors (tesseract.cc:256-292, 673-737):
    per-shot active-entry masks built on the device (csrc/sparsify.cuh)."""
    g = W.load_golden(name)
    dem, kw = W.decoder_kwargs(g["config"], capi.build_det_orders)
    dec = capi.Decoder(dem, force_generic_kernel=(path == "generic"), row_scan={"fast": 1, "fast-bits": 2}.get(path, 0), **kw)
    assert dec.search_path == path, dec.search_path_reason
    n = len(g["dets"])
    if name in HEAVY_GOLDEN and path != HEAVY_GOLDEN[name][0]:
        n = HEAVY_GOLDEN[name][1]  # the multi-million-push searches of this set run on the DEM's default kernel variant only
    out = dec.decode_batch_b8(g["dets"][:n])
    assert_same(out, g, n=n)
    check_solutions_are_consistent(dec, g["dets"][:n], out)
    # one shot at a time through tsr_decode gives the same answers as the batch
    lists = W.split_lists(g["err_offsets"], g["err_idx"])
    bits = np.unpackbits(g["dets"], axis=1, bitorder="little")[:, :dec.num_detectors]
    for s in range(min(4, len(lists))):
        errs, cost, low = dec.decode_to_errors(np.nonzero(bits[s])[0])
        assert (errs, cost, low) == (lists[s], g["cost"][s], bool(g["low_confidence"][s]))


# ---- the reference's known-answer tests, on the GPU ------------------------------------------------

@pytest.mark.parametrize("which", range(4))
def test_exhaustive_small_dems(which):  # tesseract.test.cc:125-202
    dem = EXHAUSTIVE_DEMS[which]
    dec = capi.Decoder(dem)
    D = dec.num_detectors
    best = brute_force_min_cost(dem, D)
    allbits = ((np.arange(1 << D)[:, None] >> np.arange(D)[None, :]) & 1).astype(np.uint8)
    dets = np.packbits(allbits, axis=1, bitorder="little")
    out = dec.decode_batch_b8(dets)
    ref = OracleDecoder(dem, kind=KIND).decode_batch_b8(dets)
    assert_same(out, ref)
    for mask in range(1 << D):
        if out["low_confidence"][mask]:
            assert mask not in best
        else:
            assert abs(out["cost"][mask] - best[mask]) <= 1e-7  # EPSILON, utils.h:34


def test_small_known_answers():
    # tesseract.test.cc:460-476 (unbounded pqlimit and beam)
    dec = capi.Decoder("error(0.1) D0\nerror(0.1) D1\nerror(0.01) D0 D1\ndetector(0,0,0) D0\ndetector(0,0,0) D1",
                       pqlimit=None, det_beam=capi.INF_DET_BEAM)
    errs, cost, low = dec.decode_to_errors([0, 1])
    assert not low and sorted(errs) in ([2], [0, 1]) and abs(cost - min(math.log(99), 2 * math.log(9))) < 1e-9
    # tesseract.test.cc:402-423 (detcost compares ratios): the search must pick the cheaper-per-detector error
    dem = "error(0.005322067133022559) D0 D1 D3\nerror(0.0051237598826648) D0 D1 D2"
    a = capi.Decoder(dem, merge_errors=False).decode_to_errors([0, 1, 3])
    b = OracleDecoder(dem, merge_errors=False, kind=KIND).decode_to_errors([0, 1, 3])
    assert a == b
    # tesseract.test.cc:542-564 (more than 64 observables)
    dec = capi.Decoder("error(0.1) D0 L3 L70\nerror(0.1) D1 L64\ndetector(0,0,0) D0\ndetector(0,0,0) D1")
    errs, _, low = dec.decode_to_errors([0, 1])
    assert not low and dec.flipped_observables(errs) == [3, 64, 70]
    out = dec.decode_batch_b8(np.array([[3]], dtype=np.uint8))
    assert np.nonzero(np.unpackbits(out["obs"][0], bitorder="little"))[0].tolist() == [3, 64, 70]
    # shared_decoding_tests.py:130-156
    for obs_id in (0, 1, 31, 32, 63):
        dec = capi.Decoder(f"error(0.01) D0 L{obs_id}")
        errs, _, _ = dec.decode_to_errors([0])
        assert dec.flipped_observables(errs) == [obs_id]
    # shared_decoding_tests.py:214-237
    dec = capi.Decoder("error(0.5) D0 D1 L0 L2\nerror(0.01) D0\nerror(0.01) D1\nerror(0.05) D2")
    errs, _, _ = dec.decode_to_errors([0, 1])
    assert dec.flipped_observables(errs) == [0, 2]
    # tesseract_test.py:193-201
    dec = capi.Decoder("error(0.1) D0 D1 L0\nerror(0.2) D0\nerror(0.3) D1")
    errs, cost, low = dec.decode_to_errors([0, 1])
    assert not low and cost == pytest.approx(min(math.log(9), math.log(4) + math.log(7 / 3)))
    # shared_decoding_tests.py:332-361 (merging changes the cost, not the answer)
    dem = "error(0.1) D0\nerror(0.01) D0"
    ea, ca, _ = capi.Decoder(dem, merge_errors=False).decode_to_errors([0])
    eb, cb, _ = capi.Decoder(dem, merge_errors=True).decode_to_errors([0])
    p = 0.1 * 0.99 + 0.01 * 0.9
    assert ea == eb and ca == pytest.approx(math.log(9)) and cb == pytest.approx(math.log((1 - p) / p)) and ca != cb


def test_decode_batch_expected_matrices():  # shared_decoding_tests.py:255-329
    dec = capi.Decoder("error(0.1) D0 D1 L0\nerror(0.2) D0 L1\nerror(0.05) D1")
    dets = np.packbits(np.array([[1, 1], [1, 0], [0, 1]], dtype=np.uint8), axis=1, bitorder="little")
    out = dec.decode_batch_b8(dets)
    assert np.unpackbits(out["obs"], axis=1, bitorder="little")[:, :2].tolist() == [[1, 0], [0, 1], [0, 0]]
    dec = capi.Decoder("error(0.1) D0 D1 L0\nerror(0.2) D0 D2 L1\nerror(0.05) D1 D3 L0 L2\nerror(0.15) D2\nerror(0.25) D3")
    dets = np.packbits(np.array([[1, 1, 0, 0], [1, 0, 1, 0], [0, 1, 0, 1], [0, 0, 1, 1]], dtype=np.uint8), axis=1, bitorder="little")
    out = dec.decode_batch_b8(dets)
    assert np.unpackbits(out["obs"], axis=1, bitorder="little")[:, :3].tolist() == [[1, 0, 0], [0, 1, 0], [1, 0, 1], [0, 0, 0]]


def test_error_behaviour():
    dem = ("error(0.1) D0 D1\nerror(0.1) D1 D2\nerror(0.1) D2 D3\ndetector(0, 0, 0) D0\ndetector(1, 0, 0) D1\n"
           "detector(2, 0, 0) D2\ndetector(2, 0, 0) D2")
    dec = capi.Decoder(dem)
    n = dec.num_detectors
    with pytest.raises(RuntimeError, match=rf"Symptom {n} references a detector >= num_detectors \(= {n}\)\."):  # tesseract.test.cc:376-400
        dec.decode_to_errors([n])
    with pytest.raises(ValueError, match="stride"):
        capi.Decoder(W.load_dem("surface_d5_r5_p0.002_X")).decode_batch_b8(np.zeros((2, 3), dtype=np.uint8))
    with pytest.raises(ValueError):
        dec.decode_to_errors([0], order=5, beam=1)
    # a syndrome no error set explains is low confidence, not an exception (tesseract.cc:416-419)
    assert dec.decode_to_errors([0]) == ([], 0.0, True)


def test_empty_and_ragged_inputs():
    dem = W.load_dem("surface_d5_r5_p0.002_X")
    dec = capi.Decoder(dem, det_beam=5, no_revisit_dets=False)
    out = dec.decode_batch_b8(np.zeros((0, 15), dtype=np.uint8))  # empty batch
    assert out["obs"].shape == (0, 1) and len(out["err_offsets"]) == 1
    dets, _ = capi.sample_dem_b8(dem, 3, 257, 120, 1)
    dets[::7] = 0  # shots with no fired detector decode to the empty set at cost 0
    ref = OracleDecoder(dem, det_beam=5, no_revisit_dets=False, kind=KIND).decode_batch_b8(dets)
    out = dec.decode_batch_b8(dets)
    assert_same(out, ref)
    assert all(out["cost"][::7] == 0) and not out["low_confidence"][::7].any()
    # a wider stride with junk in the padding bytes and in the unused high bits is ignored
    wide = np.full((257, 23), 0xFF, dtype=np.uint8)
    wide[:, :15] = dets
    assert_same(dec.decode_batch_b8(wide), ref)
    # D not a multiple of 8 / 32: bits >= D in the last byte are ignored
    dem7 = "\n".join(f"error(0.05) D{i} D{(i + 1) % 7}" for i in range(7)) + "\nerror(0.02) D0 L0"
    d7 = capi.Decoder(dem7)
    syn = np.array([[0b0000011], [0b1000001], [0b0101000]], dtype=np.uint8)
    a = d7.decode_batch_b8(syn)
    b = d7.decode_batch_b8(syn | 0x80)
    r = OracleDecoder(dem7, kind=KIND).decode_batch_b8(syn)
    assert_same(a, r)
    assert_same(b, r)


# ---- fresh seeded shots against the CPU checker, counters included ---------------------------------

FRESH = [
    # (config, shots, seed)
    (W.CONFIGS["C1"], 3000, 21),
    (dict(dem="surface_d5_r5_p0.002_X", det_beam=5, pqlimit=200000, no_revisit_dets=True, num_det_orders=20,
          det_order_seed=2384753), 600, 22),  # Python-side defaults (tesseract.pybind.h:65-67, 144-150)
    (dict(dem="color_superdense_d5_r5_p0.001_Z", det_beam=2, beam_climbing=True, pqlimit=10000, no_revisit_dets=True,
          num_det_orders=3), 800, 23),
    (dict(dem="color_superdense_d7_r7_p0.001_X", det_beam=5, pqlimit=200000, no_revisit_dets=True, num_det_orders=1), 200, 24),
    (dict(dem="transcx_d3_p0.001_X", det_beam=W.INF_BEAM, pqlimit=None, no_revisit_dets=False, num_det_orders=0,
          det_penalty=0.5), 3000, 25),
    (dict(dem="transcx_d5_p0.001_X", det_beam=W.INF_BEAM, pqlimit=None, no_revisit_dets=False, num_det_orders=0,
          det_penalty=1.0), 300, 26),
    (dict(dem="bb72_r6_p0.001_X", det_beam=W.INF_BEAM, pqlimit=500, no_revisit_dets=True, num_det_orders=2), 150, 27),  # pqlimit hits
    (dict(dem="bb144_r12_p0.0001_X", det_beam=W.INF_BEAM, pqlimit=None, no_revisit_dets=True, num_det_orders=24), 40, 28),
    (dict(dem="surface_d11_r11_p0.001_X", det_beam=3, beam_climbing=True, pqlimit=20000, no_revisit_dets=True,
          num_det_orders=0), 60, 29),
    # --sparsify-errors on fresh shots: the sinter presets' settings (tesseract_sinter_compat.pybind.h:454-488) and a tight limit
    (dict(dem="transcx_d5_p0.001_X", det_beam=15, beam_climbing=True, pqlimit=200000, no_revisit_dets=True, num_det_orders=16,
          sparsify_errors=True, sparsify_base_degree=2), 300, 32),
    (dict(dem="bb144_r12_p0.0001_X", det_beam=W.INF_BEAM, pqlimit=None, no_revisit_dets=True, num_det_orders=24,
          sparsify_errors=True, sparsify_base_degree=3), 64, 33),
    (dict(dem="color_superdense_d7_r7_p0.001_X", det_beam=5, pqlimit=200000, no_revisit_dets=True, num_det_orders=2,
          sparsify_errors=True, sparsify_base_degree=3, sparsify_max_degree=5, sparsify_reactivate_limit=25), 300, 34),
    # real ensembles: BFS and coordinate detector orders (utils.cc:89-171) are distinct permutations
    (dict(dem="transcx_d3_p0.001_X", det_beam=4, pqlimit=50000, no_revisit_dets=True, num_det_orders=6, det_order_method=0), 500, 30),
    (dict(dem="color_superdense_d5_r5_p0.001_Z", det_beam=3, beam_climbing=True, pqlimit=20000, no_revisit_dets=True,
          num_det_orders=5, det_order_method=2), 400, 31),
]


@pytest.mark.parametrize("case", range(len(FRESH)))
def test_fresh_shots_against_the_cpu_checker(case):
    cfg, shots, seed = FRESH[case]
    dem, kw = W.decoder_kwargs(cfg, capi.build_det_orders)
    dec = capi.Decoder(dem, **kw)
    dets, _ = capi.sample_dem_b8(dem, seed, shots, dec.num_detectors, dec.num_observables)
    ref = OracleDecoder(dem, kind=KIND, num_threads=8, **kw).decode_batch_b8(dets)
    dec.set_instrumented(True)
    dec.counters(reset=True)
    out = dec.decode_batch_b8(dets)
    assert_same(out, ref)
    check_solutions_are_consistent(dec, dets, out)
    if case % 2 == 0:  # the generic kernel, on the same shots
        assert_same(capi.Decoder(dem, force_generic_kernel=True, **kw).decode_batch_b8(dets), ref)
    else:  # the other row-scan variant of the fast kernel
        other = capi.Decoder(dem, row_scan=1 if dec.search_path == "fast-bits" else 2, **kw)
        assert other.search_path != dec.search_path
        assert_same(other.decode_batch_b8(dets), ref)
    if oracle.available("port") and shots <= 1000:
        # The search itself is the same search: pushes, pops, expansions, chain steps and visited probes
        # agree exactly with the instrumented CPU port.  Scanned heuristic entries (S) only
        # agree approximately: the reference memoises a parent's detector terms lazily, the kernel computes
        # them for every fired detector of an expanded node.  The port runs every literal trial; the GPU
        # runs distinct (order, beam) trials once, so compare only when the two coincide.
        port = OracleDecoder(dem, kind="port", num_threads=8, **kw)
        port.decode_batch_b8(dets)
        pc, gc = port.counters(), dec.counters()
        n_literal = max(len(kw["det_orders"]) if kw["det_orders"] is not None else 1,
                        (kw["det_beam"] + 1) if kw["beam_climbing"] else 0)
        if gc["searches"] == shots * n_literal:
            for k in ("Q", "P", "X", "L", "V"):
                assert gc[k] == pc[k], (k, gc, pc)
            assert abs(gc["S"] - pc["S"]) <= 0.15 * pc["S"], (gc, pc)


def test_uncertified_dems_take_the_generic_kernel_and_still_match():
    """DEMs outside the fast path's certificate (csrc/fast.h) — p > 0.5 (negative costs), more than 128 distinct
    costs, an error on more than 11 detectors — are decoded by the event-replay kernel, with the same answers as
    the CPU checker."""
    rng = np.random.default_rng(5)
    ring = "\n".join(f"error({0.01 + 0.002 * i}) D{i} D{(i + 1) % 150}" for i in range(150)) + "\n" + \
           "\n".join(f"error({0.02 + 0.001 * i}) D{i} L{i % 3}" for i in range(150))
    wide = "error(0.01) " + " ".join(f"D{i}" for i in range(13)) + " L0\n" + "\n".join(f"error(0.03) D{i} D{i + 1}" for i in range(13)) + \
           "\nerror(0.02) D0\nerror(0.02) D13 L1"
    neg = "error(0.7) D0 D1 L0\nerror(0.1) D0\nerror(0.1) D1\nerror(0.6) D1 D2\nerror(0.2) D2 L1"
    for dem, shots in ((ring, 400), (wide, 200), (neg, 8)):
        dec = capi.Decoder(dem, det_beam=W.INF_BEAM, pqlimit=100000)
        assert dec.search_path == "generic" and dec.search_path_reason
        D = dec.num_detectors
        bits = (rng.random((shots, D)) < 0.06).astype(np.uint8)
        if shots == 8:
            bits = ((np.arange(8)[:, None] >> np.arange(D)[None, :]) & 1).astype(np.uint8)
        dets = np.packbits(bits, axis=1, bitorder="little")
        ref = OracleDecoder(dem, det_beam=W.INF_BEAM, pqlimit=100000, kind=KIND, num_threads=4).decode_batch_b8(dets)
        assert_same(dec.decode_batch_b8(dets), ref)


# ---- invariances at larger sizes --------------------------------------------------------------------

def test_results_do_not_depend_on_batching_slots_or_pool_tiers():
    cfg = W.CONFIGS["C1"]
    dem, kw = W.decoder_kwargs(cfg, capi.build_det_orders)
    base = capi.Decoder(dem, **kw)
    dets, obs = capi.sample_dem_b8(dem, 1001, 20000, base.num_detectors, base.num_observables)
    a = base.decode_batch_b8(dets)
    check_solutions_are_consistent(base, dets[:2000], {k: (v[:2000] if k in ("obs", "cost", "low_confidence") else v) for k, v in a.items()})
    assert_same(base.decode_batch_b8(dets), a)  # deterministic
    small = capi.Decoder(dem, max_batch_shots=777, search_slots=96, warps_per_block=2, **kw)
    assert_same(small.decode_batch_b8(dets), a)
    for w in (1, 3, 8, 16):  # warps per search do not change the answer
        assert_same(capi.Decoder(dem, warps_per_block=w, **kw).decode_batch_b8(dets[:4000]), a, n=4000)
    # A starved node pool: searches overflow tier 0 and are re-run in the larger tiers (never on the CPU).
    tiny = capi.Decoder(dem, node_pool_bytes=8 << 20, **kw)
    b = tiny.decode_batch_b8(dets)
    assert tiny.last_batch_timing()["rerun_items"] > 0
    assert_same(b, a)


def test_unbounded_search_that_outgrows_every_tier_fails_loudly():
    dem = W.load_dem("bb72_r6_p0.001_X")
    dec = capi.Decoder(dem, det_beam=W.INF_BEAM, pqlimit=None, no_revisit_dets=True, node_pool_bytes=64 << 10, search_slots=64)
    dets, _ = capi.sample_dem_b8(dem, 5, 64, dec.num_detectors, dec.num_observables)
    with pytest.raises(capi.DeviceMemoryError):
        dec.decode_batch_b8(dets)


def test_device_resident_entry_point_matches_host_entry_point():
    torch = pytest.importorskip("torch")
    cfg = W.CONFIGS["C1"]
    dem, kw = W.decoder_kwargs(cfg, capi.build_det_orders)
    dec = capi.Decoder(dem, device=0, **kw)
    dets, _ = capi.sample_dem_b8(dem, 77, 5000, dec.num_detectors, dec.num_observables)
    a = dec.decode_batch_b8(dets)
    d = torch.from_numpy(dets).cuda()
    obs = torch.zeros((5000, 1), dtype=torch.uint8, device="cuda")
    cost = torch.zeros(5000, dtype=torch.float64, device="cuda")
    low = torch.zeros(5000, dtype=torch.uint8, device="cuda")
    torch.cuda.synchronize()
    dec.decode_batch_b8_device(d.data_ptr(), dets.shape[1], 5000, obs.data_ptr(), cost.data_ptr(), low.data_ptr())
    torch.cuda.synchronize()
    assert np.array_equal(obs.cpu().numpy(), a["obs"]) and np.array_equal(cost.cpu().numpy(), a["cost"])
    assert np.array_equal(low.cpu().numpy(), a["low_confidence"])
    off, idx = dec.last_batch_errors()
    assert np.array_equal(off, a["err_offsets"]) and np.array_equal(idx, a["err_idx"])


def test_at_most_two_errors_per_detector_matches_the_cpu_port():
    """No reference code exists for this flag (PARITY UNPINNED against the reference).  It is pinned to the
    next best thing: the CPU port's independent restatement (oracle/port.cc, itself checked against brute force in
    tests/test_oracle.py) — the three CUDA implementations (generic event replay, entry-per-lane, bit-sliced) must
    reproduce its observables, costs, flags and error lists bit for bit, and no detector may be touched by more than
    two chosen errors."""
    for name, shots, extra in (("transcx_d5_p0.001_X", 400, dict(det_penalty=0.5)), ("bb72_r6_p0.001_X", 120, dict(num_orders=2)),
                               ("surface_d5_r5_p0.002_X", 1500, dict(beam=5))):
        dem = W.load_dem(name)
        kw = dict(det_beam=extra.get("beam", W.INF_BEAM), pqlimit=20000, no_revisit_dets=True, det_penalty=extra.get("det_penalty", 0.0),
                  at_most_two_errors_per_detector=True)
        if extra.get("num_orders"):
            kw["det_orders"] = capi.build_det_orders(dem, extra["num_orders"], 1, W.CLI_ORDER_SEED)
        outs = {}
        for variant, opts in (("generic", dict(force_generic_kernel=True)), ("lanes", dict(row_scan=1)), ("bits", dict(row_scan=2))):
            dec = capi.Decoder(dem, **opts, **kw)
            dets, _ = capi.sample_dem_b8(dem, 31, shots, dec.num_detectors, dec.num_observables)
            outs[variant] = dec.decode_batch_b8(dets)
        ref = OracleDecoder(dem, kind="port", num_threads=8, **kw).decode_batch_b8(dets)
        for variant in outs:
            assert_same(outs[variant], ref)
        out = outs["bits"]
        check_solutions_are_consistent(dec, dets, out)
        d2e, _ = dec.maps()
        off, edets = dec.csr(1)
        for errs in W.split_lists(out["err_offsets"], out["err_idx"]):
            touched = np.zeros(dec.num_detectors, dtype=np.int32)
            for e in errs:
                ce = int(d2e[e])
                touched[edets[int(off[ce]):int(off[ce + 1])]] += 1
            assert touched.max(initial=0) <= 2


@pytest.mark.parametrize("which", range(4))
def test_at_most_two_exhaustive_small_dems(which):
    """Every syndrome of the reference's four literal DEMs (tesseract.test.cc:125-202) with the flag on: the GPU answer
    equals the port's and costs the brute-force optimum over error sets touching each detector at most twice."""
    dem = EXHAUSTIVE_DEMS[which]
    kw = dict(det_beam=W.INF_BEAM, pqlimit=None, at_most_two_errors_per_detector=True)
    dec = capi.Decoder(dem, **kw)
    D = dec.num_detectors
    best = brute_force_min_cost(dem, D, max_errors_per_detector=2)
    allbits = ((np.arange(1 << D)[:, None] >> np.arange(D)[None, :]) & 1).astype(np.uint8)
    dets = np.packbits(allbits, axis=1, bitorder="little")
    out = dec.decode_batch_b8(dets)
    assert_same(out, OracleDecoder(dem, kind="port", **kw).decode_batch_b8(dets))
    for mask in range(1 << D):
        if out["low_confidence"][mask]:
            assert mask not in best
        else:
            assert abs(out["cost"][mask] - best[mask]) <= 1e-7


def test_visualization_log_equals_the_reference():
    """The visualization log (visualization.cc; tesseract.cc:442-445, 478-481): for every node a search expands, and for
    the goal, the chosen errors (leaf to root) and the fired detectors — per trial of the reference's loop, in pop order.
    The GPU records chain indices on the device and rebuilds the lines on the host; the text must equal the CPU decoder's
    after the same sequence of single-shot decodes."""
    cases = [
        ("surface_d5_r5_p0.002_X", dict(det_beam=5, pqlimit=200000, no_revisit_dets=True, num_det_orders=3)),
        ("color_superdense_d5_r5_p0.001_Z", dict(det_beam=2, beam_climbing=True, pqlimit=10000, no_revisit_dets=True, num_det_orders=2)),
        ("transcx_d3_p0.001_X", dict(det_beam=W.INF_BEAM, pqlimit=None, no_revisit_dets=False, num_det_orders=0, det_penalty=0.5)),
    ]
    for name, cfg in cases:
        dem, kw = W.decoder_kwargs(dict(cfg, dem=name), capi.build_det_orders)
        for variant in (dict(), dict(force_generic_kernel=True), dict(row_scan=2)):
            dec = capi.Decoder(dem, create_visualization=True, **variant, **kw)
            ref = OracleDecoder(dem, kind=KIND, create_visualization=True, **kw)
            dets, _ = capi.sample_dem_b8(dem, 17, 12, dec.num_detectors, dec.num_observables)
            bits = np.unpackbits(dets, axis=1, bitorder="little")[:, :dec.num_detectors]
            for s in range(len(bits)):
                hits = np.nonzero(bits[s])[0]
                assert dec.decode_to_errors(hits) == ref.decode_to_errors(hits)
            if kw["det_orders"] is not None:  # the (detections, order, beam) overload logs one search (tesseract.cc:371-378)
                hits = np.nonzero(bits[0])[0]
                assert dec.decode_to_errors(hits, order=1, beam=1) == ref.decode_to_errors(hits, order=1, beam=1)
            a, b = dec.visualization_text(), ref.visualization_text()
            assert a == b, (name, variant, len(a), len(b))
            assert a.count("activated_errors") > 12


def test_visited_set_fingerprints_never_stand_in_for_a_different_set():
    """The reference's visited set holds the fired-detector bitsets themselves (tesseract.cc:394, 475-476, 563-565); the
    device holds 128-bit Zobrist fingerprints of them.  The audit mode checks every fingerprint hit — pops and child probes
    — against the real sets (rebuilt from the error chains): over more than 10^8 probes on the no-revisit workloads there
    is not a single hit whose sets differ, and the answers are those of the unaudited run."""
    total_probes = total_hits = 0
    for cfg, shots in ((W.CONFIGS["C4"], 81920),
                       (dict(dem="surface_d5_r5_p0.002_X", det_beam=5, pqlimit=200000, no_revisit_dets=True, num_det_orders=20,
                             det_order_seed=2384753), 40000),
                       (dict(dem="bb72_r6_p0.001_X", det_beam=W.INF_BEAM, pqlimit=20000, no_revisit_dets=True, num_det_orders=2), 400),
                       (dict(dem="surface_d11_r11_p0.001_X", det_beam=3, beam_climbing=True, pqlimit=20000, no_revisit_dets=True,
                             num_det_orders=0), 300)):
        dem, kw = W.decoder_kwargs(cfg, capi.build_det_orders)
        dec = capi.Decoder(dem, **kw)
        dets, _ = capi.sample_dem_b8(dem, 606, shots, dec.num_detectors, dec.num_observables)
        plain = dec.decode_batch_b8(dets, want_errors=False)
        dec.set_visited_audit(True)
        dec.counters(reset=True)
        audited = dec.decode_batch_b8(dets, want_errors=False)
        for k in ("obs", "cost", "low_confidence"):
            assert np.array_equal(plain[k], audited[k]), k
        hits, differing = dec.visited_audit(reset=True)
        assert differing == 0, (cfg["dem"], hits, differing)
        assert hits > 0
        total_probes += dec.counters()["V"]
        total_hits += hits
    assert total_probes >= 10 ** 8, total_probes
    assert total_hits >= 10 ** 4, total_hits  # revisits are rare (5.7e4 hits in 1.3e8 probes on these workloads); every one is checked
