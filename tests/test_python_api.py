"""The reference's Python surface (src/tesseract.pybind.h, src/tesseract_sinter_compat.pybind.h) as mirrored by the
pybind11 module over the C ABI.  Tests read like the reference's own (src/py/tesseract_test.py,
shared_decoding_tests.py, tesseract_sinter_compat_test.py); a DEM is any object whose str() is DEM text."""
import math
import pickle

import numpy as np
import pytest

import workloads as W
import tesseract_decoder_b200 as td

from tesseract_decoder_b200 import capi

tesseract = td.tesseract
sinter_compat = td.tesseract_sinter_compat
TesseractSinterDecoder = td.tesseract_sinter_compat.TesseractSinterDecoder

_DETECTOR_ERROR_MODEL = "error(0.125) D0\nerror(0.375) D0 D1\nerror(0.25) D1\ndetector(0, 0, 0) D0\ndetector(1, 0, 0) D1"


class Dem:  # stand-in for stim.DetectorErrorModel: only str() crosses the boundary (stim_utils.pybind.h:22-26)
    def __init__(self, text):
        self.text = text

    def __str__(self):
        return self.text


# ---- no GPU needed -------------------------------------------------------------------------------------

def test_config_defaults():  # tesseract_test.py:38-55
    c = tesseract.TesseractConfig(Dem(_DETECTOR_ERROR_MODEL))
    assert (c.det_beam, c.beam_climbing, c.no_revisit_dets, c.verbose, c.merge_errors, c.pqlimit, c.det_penalty) == \
        (5, False, True, False, True, 200000, 0.0)
    assert len(c.det_orders) == 20 and all(len(o) == 2 for o in c.det_orders)  # 20 DetIndex orders, seed 2384753
    assert c.det_orders == td.utils.build_det_orders(Dem(_DETECTOR_ERROR_MODEL), 20, td.utils.DetOrder.DetIndex, 2384753)
    assert tesseract.INF_DET_BEAM == 65535
    assert str(c).startswith("TesseractConfig(dem=DetectorErrorModel_Object, det_beam=5")
    c.det_beam = 9
    c.dem = Dem("error(0.1) D0")
    assert c.det_beam == 9 and c.dem == "error(0.1) D0"
    assert tesseract.TesseractConfig(det_beam=3, pqlimit=7).pqlimit == 7  # the overload without a DEM


def test_det_order_shapes():  # utils_test.py:37-76
    dem = Dem(W.load_dem("transcx_d3_p0.001_X"))
    for method in (td.utils.DetOrder.DetBFS, td.utils.DetOrder.DetIndex, td.utils.DetOrder.DetCoordinate):
        orders = td.utils.build_det_orders(dem, 4, method, 11)
        assert len(orders) == 4 and all(sorted(o) == list(range(96)) for o in orders)


def test_utils_module_functions():  # utils_test.py:33-93
    dem3 = Dem("error(0.125) D0\nerror(0.375) D0 D1\nerror(0.25) D1")
    dem10 = Dem("\n".join(f"error(0.1) D{i}" for i in range(10)))
    assert td.utils.EPSILON <= 1e-7 and not math.isfinite(td.utils.INF)
    assert td.utils.get_detector_coords(dem3) == []
    assert td.utils.get_detector_coords(Dem(_DETECTOR_ERROR_MODEL)) == [[0.0, 0.0, 0.0], [1.0, 0.0, 0.0]]
    assert td.utils.build_detector_graph(dem3) == [[1], [0]]
    asc = list(range(10))
    assert td.utils.build_det_orders(dem10, num_det_orders=1, seed=0) in ([asc], [asc[::-1]])  # the default method is DetIndex
    assert td.utils.build_det_orders(dem10, num_det_orders=1, seed=0) == \
        td.utils.build_det_orders(dem10, num_det_orders=1, method=td.utils.DetOrder.DetIndex, seed=0)
    assert td.utils.build_det_orders(dem3, num_det_orders=1, method=td.utils.DetOrder.DetBFS, seed=0) == [[0, 1]]
    assert td.utils.build_det_orders(dem3, num_det_orders=1, method=td.utils.DetOrder.DetCoordinate, seed=0) == [[0, 1]]
    assert ", ".join(map(str, td.utils.get_errors_from_dem(dem3))) == \
        "Error{cost=1.945910, symptom=Symptom{detectors=[0], observables=[]}}, " \
        "Error{cost=0.510826, symptom=Symptom{detectors=[0 1], observables=[]}}, " \
        "Error{cost=1.098612, symptom=Symptom{detectors=[1], observables=[]}}"


def test_common_module_classes_and_functions():  # common_test.py:42-151
    def set_bits(n):
        return [i for i in range(n.bit_length()) if (n >> i) & 1]

    error = td.common.Error(1.945910, [1, 2], set_bits(4324))
    assert error.likelihood_cost == pytest.approx(1.945910)
    assert error.symptom.detectors == [1, 2] and error.symptom.observables == [2, 5, 6, 7, 12]
    assert str(td.common.Error(5.5, [1, 2], [5, 10])) == "Error{cost=5.500000, symptom=Symptom{detectors=[1 2], observables=[5 10]}}"
    s = td.common.Symptom([1, 2], set_bits(4324))
    assert str(s) == "Symptom{detectors=[1 2], observables=[2 5 6 7 12]}"
    assert s == td.common.Symptom([1, 2], [2, 5, 6, 7, 12]) and s != td.common.Symptom([1], [2, 5, 6, 7, 12])
    assert td.common.Symptom().detectors == [] and td.common.Symptom().observables == []
    assert [str(t) for t in s.as_dem_instruction_targets()] == ["D1", "D2", "L2", "L5", "L6", "L7", "L12"]
    # an Error from a DEM error instruction (here: its text, what str(stim.DemInstruction) gives)
    assert str(td.common.Error("error(0.125) L3")) == "Error{cost=1.945910, symptom=Symptom{detectors=[], observables=[3]}}"
    with pytest.raises(ValueError):
        td.common.Error("detector(1, 2) D0")
    e = td.common.Error()
    for p, cost in ((0.125, 1.9459101490553132), (0.5, 0.0)):
        e.set_with_probability(p)
        assert e.likelihood_cost == pytest.approx(cost) and e.get_probability() == pytest.approx(p)
    for bad in (0.0, 1.0, -0.1, 1.1):
        with pytest.raises(ValueError):
            e.set_with_probability(bad)
    # DEM in, DEM (text) out: stim.DetectorErrorModel(text) where stim is available
    assert td.common.merge_indistinguishable_errors(Dem("")) == ""
    assert td.common.remove_zero_probability_errors(Dem("")) == ""
    assert td.common.dem_from_counts(Dem(""), [], 3) == ""
    merged = td.common.merge_indistinguishable_errors(Dem("error(0.1) D0 D1\nerror(0.2) D1 D0\nerror(0) D2"))
    assert [ln.split(")")[1] for ln in merged.splitlines()] == [" D0 D1", " D2"]
    assert td.common.remove_zero_probability_errors(Dem(merged)).count("error") == 1
    assert td.common.merge_weights(1.0, 2.0) == capi.merge_weights(1.0, 2.0)


def test_module_layout_and_ostream_redirect(tmp_path):  # tesseract.pybind.cc:27-47, visualization.pybind.h:14-18
    for sub in ("common", "utils", "viz", "tesseract", "tesseract_sinter_compat"):
        assert hasattr(td, sub)
    log = tmp_path / "empty.log"
    td.viz.Visualizer().write(str(log))  # a free-standing Visualizer has logged nothing
    assert log.read_text() == ""
    import contextlib
    import io
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        with td.ostream_redirect(stdout=True, stderr=True):  # the C++ side's std::cout lands in Python's sys.stdout
            td.common.merge_indistinguishable_errors(Dem("error(0.1) L0"))
    assert "errors that do not flip any detectors" in buf.getvalue()  # common.cc:154-157


def test_sinter_decoder_object():  # tesseract_sinter_compat_test.py:691-795
    d = TesseractSinterDecoder()
    assert (d.det_beam, d.no_revisit_dets, d.pqlimit, d.num_det_orders, d.seed) == (5, True, 200000, 0, 2384753)
    e = TesseractSinterDecoder(det_beam=20, beam_climbing=True, pqlimit=10**6, num_det_orders=21)
    assert e != d and pickle.loads(pickle.dumps(e)) == e and pickle.loads(pickle.dumps(e)).num_det_orders == 21
    presets = td.make_tesseract_sinter_decoders_dict()
    assert {"tesseract", "tesseract-long-beam", "tesseract-short-beam"} <= set(presets)
    assert presets["tesseract"] == presets["tesseract-long-beam"] and presets["tesseract-short-beam"].det_beam == 15


def test_invalid_inputs_without_decoding():
    with pytest.raises(ValueError):  # tesseract.cc:178-184
        tesseract.TesseractConfig(Dem(_DETECTOR_ERROR_MODEL), det_orders=[[0, 1, 2]]).compile_decoder()
    with pytest.raises(ValueError):
        tesseract.TesseractConfig(Dem("error(1.5) D0")).compile_decoder()


def test_sparsify_options_are_validated():  # tesseract_test.py:295-465, tesseract.cc:256-277
    dem = Dem(_DETECTOR_ERROR_MODEL)
    c = tesseract.TesseractConfig(dem, sparsify_errors=True, sparsify_base_degree=2, sparsify_max_degree=4, sparsify_reactivate_limit=3)
    assert (c.sparsify_errors, c.sparsify_base_degree, c.sparsify_max_degree, c.sparsify_reactivate_limit) == (True, 2, 4, 3)
    assert tesseract.TesseractConfig(dem).sparsify_errors is False and tesseract.TesseractConfig(dem).sparsify_base_degree == -1
    for kw, msg in ((dict(sparsify_base_degree=0), "sparsify_base_degree must be > 0 when sparsify_errors is enabled."),
                    (dict(sparsify_base_degree=2, sparsify_max_degree=-2), "sparsify_max_degree must be >= -1."),
                    (dict(sparsify_base_degree=2, sparsify_reactivate_limit=-2), "sparsify_reactivate_limit must be >= -1."),
                    (dict(sparsify_base_degree=3, sparsify_max_degree=2), "sparsify_max_degree must be >= sparsify_base_degree.")):
        with pytest.raises(ValueError, match=msg.replace(".", r"\.")):
            tesseract.TesseractConfig(dem, sparsify_errors=True, **kw).compile_decoder()
    # tesseract.h:37 / tesseract.cc:45-69, 104-107: round(4.5^(base-2) * D / 3)
    assert tesseract.suggest_sparsify_reactivate_limit(1728, 3) == 2592
    assert tesseract.suggest_sparsify_reactivate_limit(120, 2) == 40
    assert tesseract.suggest_sparsify_reactivate_limit(0, 3) == 0


# ---- on the GPU ----------------------------------------------------------------------------------------

@pytest.mark.gpu
def test_decode_single_shot_api():  # shared_decoding_tests.py:37-237, tesseract_test.py:193-201
    dec = tesseract.TesseractConfig(Dem("error(0.1) D0 D1 L0\nerror(0.2) D0\nerror(0.3) D1")).compile_decoder()
    assert (dec.num_detectors, dec.num_observables, len(dec.errors)) == (2, 1, 3)
    errs = dec.decode_to_errors(np.array([True, True]))
    assert errs == dec.predicted_errors_buffer and not dec.low_confidence_flag
    assert dec.cost_from_errors(errs) == pytest.approx(min(math.log(9), math.log(4) + math.log(7 / 3)))
    assert dec.get_observables_from_errors(errs) == [errs == [0]]
    assert dec.decode(np.array([True, True])).dtype == np.bool_
    assert dec.decode_from_detection_events([0, 1]).tolist() == dec.decode(np.array([True, True])).tolist()
    assert dec.decode_to_errors(np.array([True, False]), 0, 5) == [1]
    assert dec.errors[0].symptom.detectors == [0, 1] and dec.errors[0].likelihood_cost == pytest.approx(math.log(9))
    with pytest.raises(ValueError, match=r"Syndrome array size \(3\) does not match the number of detectors in the decoder \(2\)\."):
        dec.decode(np.array([True, False, True]))  # shared_decoding_tests.py:378-382
    with pytest.raises(RuntimeError, match="Input syndromes must be a 2D NumPy array."):
        dec.decode_batch(np.array([True, False]))  # :251
    with pytest.raises(ValueError, match=r"The number of detectors in the input array \(3\) does not match"):
        dec.decode_batch(np.zeros((2, 3), dtype=bool))  # :399-403
    dec = tesseract.TesseractConfig(Dem("error(0.1) D0\nerror(0) D1\nerror(0.2) D2\nerror(0.3) D2")).compile_decoder()
    with pytest.raises(ValueError, match="error index does not map to a retained decoder error"):
        dec.cost_from_errors([1])


@pytest.mark.gpu
def test_decode_batch_expected_matrices():  # shared_decoding_tests.py:255-329
    dec = tesseract.TesseractDecoder(tesseract.TesseractConfig(Dem("error(0.1) D0 D1 L0\nerror(0.2) D0 L1\nerror(0.05) D1")))
    out = dec.decode_batch(np.array([[True, True], [True, False], [False, True]]))
    assert out.dtype == np.bool_ and out.tolist() == [[True, False], [False, True], [False, False]]
    cfg = tesseract.TesseractConfig()  # no detector orders yet: the decoder falls back to the identity order
    cfg.det_beam = 7
    dec = cfg.compile_decoder_for_dem(Dem("error(0.1) D0 D1 L0\nerror(0.2) D0 D2 L1\nerror(0.05) D1 D3 L0 L2\nerror(0.15) D2\nerror(0.25) D3"))
    out = dec.decode_batch(np.array([[1, 1, 0, 0], [1, 0, 1, 0], [0, 1, 0, 1], [0, 0, 1, 1]], dtype=bool))
    assert out.tolist() == [[True, False, False], [False, True, False], [True, False, True], [False, False, False]]
    assert dec.config.det_beam == 7


@pytest.mark.gpu
def test_sinter_bit_packed_known_answers(tmp_path):  # tesseract_sinter_compat_test.py:75-162, 165-372
    dem = Dem("detector(0, 0, 0) D0\ndetector(0, 0, 1) D1\ndetector(0, 0, 2) D2\nerror(0.1) D0 D1 L0\nerror(0.1) D1 D2 L1")
    compiled = TesseractSinterDecoder().compile_decoder_for_dem(dem=dem)
    assert (compiled.num_detectors, compiled.num_observables) == (3, 2)
    assert compiled.decoder.config.det_beam == 5
    dets = np.array([[0b011], [0b110], [0b101]], dtype=np.uint8)
    out = compiled.decode_shots_bit_packed(bit_packed_detection_event_data=dets)
    assert out.dtype == np.uint8 and out.tolist() == [[0b01], [0b10], [0b11]]
    with pytest.raises(ValueError):
        compiled.decode_shots_bit_packed(bit_packed_detection_event_data=np.zeros((2, 2), dtype=np.uint8))
    (tmp_path / "m.dem").write_text(str(dem))
    dets.tofile(tmp_path / "dets.b8")
    TesseractSinterDecoder().decode_via_files(num_shots=3, num_dets=3, num_obs=2, dem_path=tmp_path / "m.dem",
                                              dets_b8_in_path=tmp_path / "dets.b8",
                                              obs_predictions_b8_out_path=tmp_path / "out.b8", tmp_dir=tmp_path)
    assert np.fromfile(tmp_path / "out.b8", dtype=np.uint8).tolist() == [0b01, 0b10, 0b11]


@pytest.mark.gpu
def test_bit_packed_equals_decode_batch_on_color_code_shots():  # tesseract_sinter_compat_test.py:578-653
    dem_text = W.load_dem("color_superdense_d5_r5_p0.001_Z")
    dets, _ = td.capi.sample_dem_b8(dem_text, 9, 1000, td.capi.dem_count_detectors(dem_text), 1)
    D = td.capi.dem_count_detectors(dem_text)
    bits = np.unpackbits(dets, axis=1, bitorder="little")[:, :D].astype(bool)
    for kw in (dict(), dict(det_beam=2, beam_climbing=True), dict(no_revisit_dets=False), dict(pqlimit=1000, det_penalty=0.1)):
        sinter = TesseractSinterDecoder(num_det_orders=4, **kw).compile_decoder_for_dem(dem=Dem(dem_text))
        packed = sinter.decode_shots_bit_packed(bit_packed_detection_event_data=dets)
        cfg = tesseract.TesseractConfig(Dem(dem_text), det_orders=td.utils.build_det_orders(Dem(dem_text), 4, td.utils.DetOrder.DetIndex, 2384753), **kw)
        batch = cfg.compile_decoder().decode_batch(bits)
        assert np.array_equal(np.unpackbits(packed, axis=1, bitorder="little")[:, :1].astype(bool), batch)


@pytest.mark.gpu
def test_sinter_sparsify_presets_decode():  # tesseract_sinter_compat.pybind.h:442-490
    d = sinter_compat.make_tesseract_sinter_decoders_dict()
    assert set(d) == {"tesseract", "tesseract-long-beam", "tesseract-short-beam", "tesseract-long-beam-sparsify-color-code-like",
                      "tesseract-long-beam-sparsify-surface-code-like", "tesseract-short-beam-sparsify-color-code-like",
                      "tesseract-short-beam-sparsify-surface-code-like"}
    p = d["tesseract-short-beam-sparsify-surface-code-like"]
    assert (p.sparsify_errors, p.sparsify_base_degree, p.sparsify_max_degree, p.sparsify_reactivate_limit, p.det_beam) == (True, 2, -1, -1, 15)
    assert pickle.loads(pickle.dumps(p)) == p
    dem = W.load_dem("surface_d5_r5_p0.002_X")
    compiled = p.compile_decoder_for_dem(dem=Dem(dem))
    assert compiled.decoder.config.sparsify_reactivate_limit == 40  # the suggested limit replaces -1 (tesseract.cc:272-277)
    dets, obs = capi.sample_dem_b8(dem, 12, 400, 120, 1)
    got = compiled.decode_shots_bit_packed(bit_packed_detection_event_data=dets)
    kw = dict(det_beam=15, beam_climbing=True, pqlimit=200000, no_revisit_dets=True, sparsify_errors=True, sparsify_base_degree=2,
              det_orders=capi.build_det_orders(dem, 16, 1, 2384753))
    from oracle import oracle
    kind = "reference" if oracle.available("reference") else "port"
    ref = oracle.OracleDecoder(dem, kind=kind, num_threads=8, **kw).decode_batch_b8(dets)
    assert np.array_equal(got, ref["obs"])
