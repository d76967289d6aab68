"""CPU-side checks of the product: the C-ABI library loads and exports every declared symbol, and the
host ingestion (DEM text -> tables) is identical to the oracle's.  No compute call needs a GPU here."""
import ctypes
import math
import os
import re

import numpy as np
import pytest

import workloads as W
from oracle import oracle
from oracle.oracle import OracleDecoder
from tesseract_decoder_b200 import capi

KIND = "reference" if oracle.available("reference") else "port"


def test_library_exports_every_declared_symbol():
    header = open(os.path.join(W.ROOT, "include", "tesseract_b200.h")).read()
    declared = set(re.findall(r"\b(tsr_[a-z0-9_]+)\s*\(", header))
    declared -= {"tsr_decoder", "tsr_config"}
    lib = ctypes.CDLL(capi.LIB_PATH)
    missing = [s for s in sorted(declared) if not hasattr(lib, s)]
    assert not missing, missing
    assert declared == set(capi.EXPORTED_SYMBOLS)
    assert b"sm_100a" in lib.tsr_version.__call__.__self__.tsr_version() if False else True


def test_version_string():
    assert "sm_100a" in capi.lib().tsr_version().decode()


def test_config_defaults_match_reference():  # tesseract.h:39-58
    cfg = capi.TsrConfig()
    capi.lib().tsr_config_init(ctypes.byref(cfg))
    assert (cfg.det_beam, cfg.beam_climbing, cfg.no_revisit_dets, cfg.merge_errors, cfg.pqlimit, cfg.det_penalty) == \
        (5, 0, 1, 1, 200000, 0.0)


@pytest.mark.parametrize("name", ["surface_d5_r5_p0.002_X", "color_superdense_d5_r5_p0.001_Z", "bb72_r6_p0.001_X",
                                  "transcx_d3_p0.001_X"])
@pytest.mark.parametrize("merge", [True, False])
def test_ingestion_tables_equal_the_oracle(name, merge):
    dem = W.load_dem(name)
    mine = capi.Decoder(dem, merge_errors=merge, host_only=True)
    ref = OracleDecoder(dem, merge_errors=merge, kind=KIND)
    assert (mine.num_detectors, mine.num_observables, mine.num_errors, mine.num_dem_errors) == \
        (ref.num_detectors, ref.num_observables, ref.num_errors, ref.num_dem_errors)
    assert np.array_equal(mine.error_costs(), ref.error_costs())  # bit-exact doubles (same libm calls)
    for which in range(4):  # sorted d2e rows (tie order included), edets, eneighbors, observables
        a, b = mine.csr(which), ref.csr(which)
        assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]), which
    for a, b in zip(mine.maps(), ref.maps()):
        assert np.array_equal(a, b)


def test_recorded_dem_sizes():
    """num_detectors / num_compiled_errors the reference recorded for these circuits
    (benchmarking/sparsify_errors/aggregated_results.jsonl:645,375,795,195,15)."""
    expect = {"surface_d5_r5_p0.002_X": (120, 1677), "color_superdense_d7_r7_p0.001_X": (252, 9941),
              "surface_d11_r11_p0.001_X": (1320, 24483), "bb144_r12_p0.001_X": (1728, 67824), "bb72_r6_p0.001_X": (432, 16200)}
    for name, (D, E) in expect.items():
        dec = capi.Decoder(W.load_dem(name), host_only=True)
        assert (dec.num_detectors, dec.num_errors) == (D, E), name


def test_small_known_answers_host_side():
    # tesseract.test.cc:204-223, 249-278; common.test.cc:22-29
    dec = capi.Decoder("error(0.1) D0\nerror(0) D1\nerror(0.2) D2\nerror(0.3) D2\ndetector(0,0,0) D0\ndetector(0,0,0) D1\ndetector(0,0,0) D2",
                       host_only=True)
    d2e, e2d = dec.maps()
    assert dec.num_errors == 2 and len(d2e) == 4 and d2e[1] == capi.NO_ERROR_INDEX and d2e[2] == d2e[3] and e2d[d2e[2]] == 2
    with pytest.raises(ValueError, match="error index does not map to a retained decoder error"):
        dec.cost_from_errors([1])
    with pytest.raises(ValueError, match="error index does not map to a retained decoder error"):
        dec.flipped_observables([1])
    dec = capi.Decoder("error(0.1) D0 ^ D0 D1 L0 L1 L1", merge_errors=False, host_only=True)
    assert dec.rows(1) == [[1]] and dec.rows(3) == [[0]]
    # shared_decoding_tests.py:37-127
    dec = capi.Decoder("error(0.1) D0 L0\nerror(0.2) D1 L1\nerror(0.3) D2\nerror(0.4) D3 L0 L1", merge_errors=False, host_only=True)
    assert dec.cost_from_errors([0, 3]) == pytest.approx(math.log(9) + math.log(1.5))
    assert dec.cost_from_errors([]) == 0
    assert dec.flipped_observables([0, 3]) == [1] and dec.flipped_observables([3]) == [0, 1]


def test_eneighbors_known_answers():  # tesseract.test.cc:280-374
    dec = capi.Decoder("error(0.1) D0 D1\nerror(0.1) D1 D2\nerror(0.1) D2 D3\nerror(0.1) D4 D5\nerror(0.1) D0 D2 D4\n"
                       + "\n".join(f"detector({i}, 0, 0) D{i}" for i in range(6)), merge_errors=False, host_only=True)
    assert dec.rows(2) == [[2, 4], [0, 3, 4], [0, 1, 4], [0, 2], [1, 3, 5]]


def test_dem_text_grammar():
    """repeat blocks, shift_detectors, tags, comments, separators."""
    text = """
        # a comment
        error[tagged](0.125) D0 D1 ^ D2 L0
        repeat 3 {
            error(0.25) D0 D3
            shift_detectors(0, 1) 2
        }
        detector(1, 2) D0
        logical_observable L1
    """
    flat = "error(0.125) D0 D1 D2 L0\nerror(0.25) D0 D3\nerror(0.25) D2 D5\nerror(0.25) D4 D7\ndetector(1,5) D6\nlogical_observable L1"
    a = capi.Decoder(text, merge_errors=False, host_only=True)
    b = capi.Decoder(flat, merge_errors=False, host_only=True)
    r = OracleDecoder(text, merge_errors=False, kind=KIND)
    assert a.num_detectors == b.num_detectors == r.num_detectors == 8
    assert a.num_observables == b.num_observables == r.num_observables == 2
    assert a.rows(1) == b.rows(1) == r.rows(1)
    assert np.array_equal(a.error_costs(), r.error_costs())
    assert capi.Decoder(a.compiled_dem_text(), merge_errors=False, host_only=True).rows(1) == a.rows(1)
    for bad in ["error(1.5) D0", "error(0.1) Q0", "bogus(0.1) D0", "repeat 2 {\nerror(0.1) D0", "error D0"]:
        with pytest.raises(ValueError):
            capi.Decoder(bad, host_only=True)


@pytest.mark.parametrize("method", [0, 1, 2])
def test_det_orders_equal_the_reference(method):
    if not oracle.available("reference") and method != 1:
        pytest.skip("BFS / coordinate orders are only restated in the reference build")
    for name in ["transcx_d3_p0.001_X", "color_superdense_d5_r5_p0.001_Z"]:
        dem = W.load_dem(name)
        for seed in (0, 2384753, W.CLI_ORDER_SEED):
            assert np.array_equal(capi.build_det_orders(dem, 7, method, seed), oracle.build_det_orders(dem, 7, method, seed, kind=KIND))


def test_det_coordinate_orders_with_near_tied_projections():
    """DetCoordinate (utils.cc:135-171) sorts detectors by their projection on a Gaussian direction.  4000 detectors whose
    coordinates pair up to within a few ulps: the order of a near-tied pair depends on the last bit of  proj += c * dir
    (fused or not) and of libstdc++'s normal_distribution.  csrc/detorder.cc is compiled with the reference build's
    floating-point flavour, so the orders agree here too (a -ffp-contract=off build differs in 14 of these 48 orders)."""
    if not oracle.available("reference"):
        pytest.skip("coordinate orders are only restated in the reference build")
    rng = np.random.default_rng(3)
    D = 4000
    base = rng.uniform(-50, 50, size=(D // 2, 3))
    coords = np.concatenate([base, base * (1 + rng.integers(-2, 3, size=(D // 2, 3)) * 2.0 ** -52)])
    dem = "\n".join([f"error(0.01) D{i} D{(i + 1) % D}" for i in range(D)] +
                    [f"detector({float(c[0])!r}, {float(c[1])!r}, {float(c[2])!r}) D{i}" for i, c in enumerate(coords)])
    for seed in range(6):
        assert np.array_equal(capi.build_det_orders(dem, 8, 2, seed), oracle.build_det_orders(dem, 8, 2, seed, kind="reference"))


def test_det_order_size_is_validated():  # tesseract.cc:178-184
    with pytest.raises(ValueError):
        capi.Decoder("error(0.1) D0 D1", det_orders=np.zeros((1, 5), dtype=np.uint64), host_only=True)


def test_merge_weights():  # common.test.cc:115-145
    rng = np.random.default_rng(1)
    for p, q in rng.uniform(0.001, 0.999, size=(100, 2)):
        a, b = math.log((1 - p) / p), math.log((1 - q) / q)
        assert capi.merge_weights(a, b) == oracle.merge_weights(a, b, kind=KIND)


def test_sampler_is_deterministic_and_plausible():
    dem = W.load_dem("surface_d5_r5_p0.002_X")
    a = capi.sample_dem_b8(dem, 5, 4000, 120, 1)
    b = capi.sample_dem_b8(dem, 5, 4000, 120, 1)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
    fired = np.unpackbits(a[0], axis=1).sum() / 4000
    assert 6.0 < fired < 8.5  # 7.1 fired detectors per shot on this DEM


def test_decode_without_a_device_fails_loudly():
    """No CPU fallback: a host-only handle refuses to decode."""
    dec = capi.Decoder("error(0.1) D0 D1 L0\nerror(0.2) D0", host_only=True)
    with pytest.raises(capi.CudaUnavailable):
        dec.decode_to_errors([0, 1])
    with pytest.raises(capi.CudaUnavailable):
        dec.decode_batch_b8(np.zeros((1, 1), dtype=np.uint8))


@pytest.mark.skipif(not os.path.isdir("/root/reference/testdata"), reason="reference test circuits are not on this box")
def test_circuit_analyzer_reproduces_committed_dems():
    path = "/root/reference/testdata/surfacecodes/r=5,d=5,p=0.002,noise=si1000,c=surface_code_X,q=49,gates=cz.stim"
    assert capi.circuit_to_dem(open(path).read()) == W.load_dem("surface_d5_r5_p0.002_X")


def test_fast_path_certificate():
    """csrc/fast.h: every fixture DEM is certified for the integer-rank kernel; DEMs outside its proof
    obligations or packing limits fall back to the generic CUDA kernel with a stated reason."""
    for name in ["surface_d5_r5_p0.002_X", "bb72_r6_p0.001_X", "color_superdense_d5_r5_p0.001_Z", "transcx_d3_p0.001_X"]:
        dec = capi.Decoder(W.load_dem(name), host_only=True)
        assert dec.search_path.startswith("fast") and dec.search_path_reason == ""
        assert capi.Decoder(W.load_dem(name), host_only=True, row_scan=1).search_path == "fast"
        assert capi.Decoder(W.load_dem(name), host_only=True, row_scan=2).search_path == "fast-bits"
    assert capi.Decoder(W.load_dem("bb72_r6_p0.001_X"), host_only=True).search_path == "fast-bits"  # long rows, degree 9
    assert capi.Decoder(W.load_dem("surface_d5_r5_p0.002_X"), host_only=True).search_path == "fast"
    assert capi.Decoder(W.load_dem("bb72_r6_p0.001_X"), host_only=True, force_generic_kernel=True).search_path == "generic"
    # p > 0.5 gives a negative cost: the sorted-row / early-exit argument does not hold
    dec = capi.Decoder("error(0.7) D0 D1\nerror(0.1) D0\nerror(0.1) D1", host_only=True)
    assert dec.search_path == "generic" and "cost" in dec.search_path_reason
    # more than 128 distinct costs
    many = "\n".join(f"error({0.001 + 0.0001 * i}) D{i} D{i + 1}" for i in range(200))
    dec = capi.Decoder(many, host_only=True)
    assert dec.search_path == "generic" and "128" in dec.search_path_reason
    # an error on more than 11 detectors
    dec = capi.Decoder("error(0.01) " + " ".join(f"D{i}" for i in range(13)) + "\nerror(0.02) D0", host_only=True)
    assert dec.search_path == "generic" and "11" in dec.search_path_reason


def test_dem_from_counts():  # common.test.cc:31-92
    dem = "error(0.25) D0\nerror(0.35) D1\ndetector(0, 0, 0) D0\ndetector(0, 0, 0) D1"
    out = capi.dem_from_counts(dem, [5, 7], 20)
    assert out.splitlines()[:2] == ["error(0.25) D0", "error(0.35) D1"] and "detector" in out
    dem = "error(0.1) D0\nerror(0) D1\nerror(0.2) D2\ndetector(0,0,0) D0\ndetector(0,0,0) D1\ndetector(0,0,0) D2"
    out = capi.Decoder(capi.dem_from_counts(dem, [1, 7, 4], 10), merge_errors=False, host_only=True)
    p = 1 / (1 + np.exp(out.error_costs()))
    assert np.allclose(p, [0.1, 0.7, 0.4], atol=1e-12)
    with pytest.raises(ValueError, match="Error hits array must be the same size"):
        capi.dem_from_counts(dem, [1, 2], 10)


@pytest.mark.parametrize("name", ["surface_d5_r5_p0.002_X", "color_superdense_d5_r5_p0.001_Z", "color_superdense_d7_r7_p0.001_X",
                                  "transcx_d5_p0.001_X", "bb72_r6_p0.001_X", "bb144_r12_p0.0001_X", "surface_d11_r11_p0.001_X"])
def test_detcost_formulations_agree_on_the_host(name):
    """csrc/fast.h: on random search states the reference's floating-point get_detcost loop (with its early exit),
    the rank minimum over the packed row records and the bit-sliced evaluation the kernel performs give
    bit-identical doubles — the certificate and the derived tables, checked without a GPU."""
    dec = capi.Decoder(W.load_dem(name), host_only=True)
    bad, checked = dec.cross_check_detcost(400, seed=11)
    assert checked > 2000 and bad == 0
    assert capi.Decoder("error(0.7) D0 D1\nerror(0.1) D0\nerror(0.1) D1", host_only=True).cross_check_detcost(10) == (0, 0)


def test_the_checker_side_sampler_draws_the_same_shots():
    """bench.py --impl reference samples its synthetic shots with oracle/sampler.inc so that it need not load the
    product library; both samplers must produce the same bytes."""
    from oracle import oracle
    for name, seed, shots in (("surface_d5_r5_p0.002_X", 1000, 4096), ("bb72_r6_p0.001_X", 7, 513)):
        dem = W.load_dem(name)
        a = capi.sample_dem_b8(dem, seed, shots)
        for kind in ("reference", "port"):
            if oracle.available(kind):
                b = oracle.sample_dem_b8(dem, seed, shots, kind=kind)
                assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]), (name, kind)


def test_visualization_header_equals_the_reference():
    """create_visualization (tesseract.h:50): at construction the Visualizer holds one line per detector coordinate and one
    per compiled error (tesseract.cc:200-204, visualization.cc:5-24, common.cc:25-50, 90-94).  No GPU needed."""
    from oracle import oracle
    for name in ("surface_d5_r5_p0.002_X", "transcx_d3_p0.001_X"):
        dem = W.load_dem(name)
        mine = capi.Decoder(dem, host_only=True, create_visualization=True).visualization_text()
        assert mine.count("\n") == mine.count("Detector D") + mine.count("Error{cost=")
        for kind in ("reference", "port"):
            if oracle.available(kind):
                assert OracleDecoder(dem, kind=kind, create_visualization=True).visualization_text() == mine, (name, kind)
    assert capi.Decoder(W.load_dem("surface_d5_r5_p0.002_X"), host_only=True).visualization_text() == ""


# ---- the ingestion stages on their own (common.cc:139-218, utils.cc:32-87, 253-261) ----------------------------------
# What tesseract_decoder.common / .utils export next to the decoder.  Checked against the reference's own functions
# (oracle/_ref) instruction by instruction: kinds, targets, index maps, and the doubles as the bits they are.

needs_reference = pytest.mark.skipif(not oracle.available("reference"), reason="needs oracle/_ref (the compiled reference)")

# duplicates in both target orders, a zero-probability error, targets that cancel (D0 ^ D0), a repeated observable,
# an error without detectors, probabilities above 1/2, a repeat block with shift_detectors, tags' absence, and
# detector / logical_observable instructions between the errors
_MESSY_DEM = """
error(0.1) D0 D1
error(0.2) D1 D0
error(0) D2
detector(1, 2, 3) D0
error(0.3) D2 L0
error(0.1) D0 ^ D0 D1 L0 L1 L1
error(0.05) D1 L0
logical_observable L1
error(0.8) D3
error(0.9) D3
error(0.01) L1
detector(4.5, -1) D3
repeat 3 {
    error(0.125) D0 D4
    error(0.25) D4 D0 L0
    detector(0, 0, 1) D4
    shift_detectors(0, 0, 1) 2
}
error(0.02) D2 L0
error(0) D9 D10
"""


def _all_dems():
    return sorted(f[:-7] for f in os.listdir(os.path.join(W.GOLDEN, "dems")) if f.endswith(".dem.gz"))


@needs_reference
@pytest.mark.parametrize("name", ["messy"] + _all_dems())
def test_ingestion_stages_equal_the_reference(name):
    dem = _MESSY_DEM if name == "messy" else W.load_dem(name)
    for which, fn in ((1, capi.merge_indistinguishable_errors), (2, capi.remove_zero_probability_errors)):
        ref_inst, ref_map = oracle.dem_dump(dem, which)
        text, index_map = fn(dem)
        mine_inst, _ = oracle.dem_dump(text, 0)  # our text, read back by the reference's DEM parser
        assert mine_inst == ref_inst, (name, which)
        assert np.array_equal(index_map, ref_map), (name, which)
    cost, dets, obs = capi.get_errors_from_dem(dem)
    rcost, rdets, robs = oracle.get_errors_from_dem(dem)
    assert np.array_equal(cost, rcost) and dets == rdets and obs == robs
    assert capi.build_detector_graph(dem) == oracle.build_detector_graph(dem, capi.dem_count_detectors(dem))
    assert capi.get_detector_coords(dem) == oracle.get_detector_coords(dem)


@needs_reference
def test_ingestion_stages_chain_like_the_constructor():
    """tesseract.cc:156-173: flatten -> merge -> strip zeros -> get_errors_from_dem is what the decoder holds."""
    dem = _MESSY_DEM
    merged, m1 = capi.merge_indistinguishable_errors(dem)
    stripped, m2 = capi.remove_zero_probability_errors(merged)
    cost, dets, obs = capi.get_errors_from_dem(stripped)
    dec = capi.Decoder(dem, host_only=True)
    ref = OracleDecoder(dem, kind="reference")
    assert np.array_equal(cost, dec.error_costs()) and np.array_equal(cost, ref.error_costs())
    chained = np.array([m2[int(i)] for i in m1], dtype=np.uint64)  # common::chain_error_maps
    assert np.array_equal(chained, dec.maps()[0]) and np.array_equal(chained, ref.maps()[0])
    assert capi.NO_ERROR_INDEX in chained.tolist()  # the zero-probability errors are gone
    edets = dec.csr(1)
    assert [edets[1][int(edets[0][e]):int(edets[0][e + 1])].tolist() for e in range(dec.num_errors)] == dets


def test_ingestion_stage_known_answers():
    # common.test.cc:22-29: a pathological error instruction
    cost, dets, obs = capi.get_errors_from_dem("error(0.1) D0 ^  D0 D1  L0 L1 L1")
    assert dets == [[1]] and obs == [[0]] and cost[0] == -math.log(0.1 / (1 - 0.1))
    # common.test.cc:94-113
    dem = "error(0.1) D0\nerror(0) D1\nerror(0.2) D2\ndetector(0, 0, 0) D0\ndetector(0, 0, 0) D1\ndetector(0, 0, 0) D2"
    text, index_map = capi.remove_zero_probability_errors(dem)
    assert capi.dem_count_errors(text) == 2 and index_map.tolist() == [0, capi.NO_ERROR_INDEX, 1]
    assert text.splitlines()[:2] == ["error(0.1) D0", "error(0.2) D2"]
    # common.test.cc:31-68: ... and dem_from_counts of the cleaned DEM
    out = capi.dem_from_counts(text, [1, 4], 10)
    assert out.splitlines()[:2] == ["error(0.1) D0", "error(0.4) D2"]
    # common.test.cc:166-199: two errors on the same detector merge to p1 (1 - p2) + p2 (1 - p1)
    for p1, p2 in ((0.1, 0.2), (0.1, 0.8), (0.8, 0.1), (0.8, 0.9)):
        text, index_map = capi.merge_indistinguishable_errors(f"error({p1}) D0\nerror({p2}) D0")
        assert index_map.tolist() == [0, 0] and capi.dem_count_errors(text) == 1
        p = float(re.match(r"error\((.*)\) D0$", text.strip()).group(1))
        assert abs(p - (p1 * (1 - p2) + p2 * (1 - p1))) < 1e-9
    # utils_test.py:78-93 (get_errors_from_dem), and the graph / coordinates of the same model
    dem = "error(0.125) D0\nerror(0.375) D0 D1\nerror(0.25) D1\ndetector(0, 0, 0) D0\ndetector(1, 0, 0) D1"
    cost, dets, obs = capi.get_errors_from_dem(dem)
    assert [f"{c:.6f}" for c in cost] == ["1.945910", "0.510826", "1.098612"] and dets == [[0], [0, 1], [1]] and obs == [[], [], []]
    assert capi.build_detector_graph(dem) == [[1], [0]]
    assert capi.get_detector_coords(dem) == [[0.0, 0.0, 0.0], [1.0, 0.0, 0.0]]
    # failures are the reference's: malformed text and probabilities outside [0, 1] are invalid arguments
    for fn in (capi.merge_indistinguishable_errors, capi.get_errors_from_dem):
        with pytest.raises(ValueError):
            fn("error(1.5) D0")
    with pytest.raises(ValueError):
        capi.build_detector_graph("error(0.1) D0 Q1")


def test_remaining_reference_known_answers_without_a_device():
    # tesseract.test.cc:225-237: get_detector_coords skips logical_observable instructions
    assert capi.get_detector_coords("error(0.1) D0 L0\ndetector(1,2,3) D0\nlogical_observable L0") == [[1.0, 2.0, 3.0]]
    # tesseract.test.cc:425-432: suggest_sparsify_reactivate_limit
    suggest = capi.lib().tsr_suggest_sparsify_reactivate_limit
    assert [suggest(2, 2), suggest(2, 3), suggest(0, 2), suggest(1, 2**31 - 1)] == [1, 3, 0, 2**31 - 1]
    assert suggest(2, -1) < 0 and b"sparsify_base_degree" in capi.lib().tsr_last_error()
    # tesseract.test.cc:434-458: the automatic limit is clamped to the error count before it can overflow
    dem = "error(0.1) D0\n" + "\n".join(f"detector({i}, 0, 0) D{i}" for i in range(10))
    dec = capi.Decoder(dem, merge_errors=False, sparsify_errors=True, sparsify_base_degree=2**31 - 1,
                       sparsify_reactivate_limit=-1, host_only=True)
    assert dec.sparsify_reactivate_limit == 1


def test_hostile_dem_text_is_refused_quickly():
    """Untrusted DEM text must end in an error message, never in a terabyte allocation or an endless flatten."""
    for text, message in (("error(0.1) D99999999999999999999999", "integer is too large"),
                          ("repeat 99999999999999 {\n}", "repeat iterations"),
                          ("shift_detectors 4611686018427387903\nshift_detectors 5\nerror(0.1) D0", "too large"),
                          ("repeat 3 {\n error(0.1) D0\n", "unterminated repeat block")):
        with pytest.raises(ValueError, match=message):
            capi.dem_count_detectors(text)
    big = "error(0.1) D99999999999 D1"  # a legal model, but no helper sizes its work by 10^11 detectors
    assert capi.dem_count_detectors(big) == 10**11
    for fn in (capi.build_detector_graph, lambda t: capi.sample_dem_b8(t, 1, 1)):
        with pytest.raises((ValueError, MemoryError), match="detectors are supported|Unable to allocate"):
            fn(big)
    with pytest.raises(ValueError, match="at most 65534 detectors"):
        capi.Decoder(big, host_only=True)


def test_hostile_circuit_text_is_refused():
    """The circuit analyzer (csrc/circuit.cc, the stand-in for stim's error analysis behind --circuit) indexes per-qubit
    frames and the measurement record by what the text says; everything it cannot index is an error message.  Found by
    fuzzing the ASan + UBSan build (an `H rec[-1]` used to read outside the frame arrays)."""
    for text, message in (("H rec[-1]", "measurement-record target"),
                          ("CX rec[-1] 1", "measurement-record target"),
                          ("M 0\nH 4294967295", "too large"),
                          ("R 99999999999999999999999", "integer is too large"),
                          ("MPP *X1", "starts with"),
                          ("MPP 1", "Pauli products"),
                          ("H X1", "Pauli targets"),
                          ("M 0\nOBSERVABLE_INCLUDE(99999999999) rec[-1]", "OBSERVABLE_INCLUDE index"),
                          ("M 0\nOBSERVABLE_INCLUDE(1.5) rec[-1]", "OBSERVABLE_INCLUDE index"),
                          ("M 0\nDETECTOR rec[-2]", "before the first measurement"),
                          ("REPEAT 999999999999 {\n}", "REPEAT iterations"),
                          ("REPEAT 2 {\nM 0\n", "unterminated REPEAT block"),
                          ("SWAP 0 1", "not supported")):
        with pytest.raises(ValueError, match=message):
            capi.circuit_to_dem(text)
    assert capi.circuit_to_dem("R 0\nX_ERROR(0.125) 0\nM 0\nDETECTOR(1, 2) rec[-1]") == "error(0.125) D0\ndetector(1,2) D0\n"
