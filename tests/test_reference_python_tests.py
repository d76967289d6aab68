"""The reference's OWN Python test files, run unmodified against this repo's module.

`tesseract_decoder` is aliased to `tesseract_decoder_b200` and `stim` to a small stand-in (tests/stim_stand_in.py: DEM
text holder, DemTarget, DemInstruction) for the duration of these tests; the test functions are then imported from
/root/reference/src/py and called as they are.  Here (no GPU) only the tests that never decode can run; everything that
decodes is covered by the mirrors in tests/test_python_api.py and tests/test_gpu_parity.py, which do not need the
reference tree (it does not exist on the GPU box)."""
import importlib.util
import os
import sys

import pytest

REF_PY = "/root/reference/src/py"
pytestmark = pytest.mark.skipif(not os.path.isdir(REF_PY), reason="the reference tree is not on this box")

# file -> the tests in it that construct configs / call ingestion helpers but never decode (no device needed)
DEVICE_FREE = {
    "common_test.py": ["test_error_from_direct_constructor", "test_error_str", "test_as_dem_instruction_targets",
                       "test_error_from_dem_instruction", "test_error_get_set_probability",
                       "test_error_set_with_probability_invalid_input", "test_merge_indistinguishable_errors",
                       "test_remove_zero_probability_errors", "test_dem_from_counts"],
    "utils_test.py": ["test_module_has_global_constants", "test_get_detector_coords", "test_build_detector_graph",
                      "test_build_det_orders_default_index", "test_build_det_orders_bfs", "test_build_det_orders_coordinate",
                      "test_build_det_orders_index", "test_get_errors_from_dem"],
    "tesseract_test.py": ["test_create_tesseract_config", "test_create_tesseract_config_with_dem",
                          "test_create_tesseract_config_with_dem_and_custom_args", "test_create_tesseract_config_no_dem",
                          "test_create_tesseract_config_no_dem_with_custom_args", "test_create_tesseract_config_sparsify_defaults",
                          "test_create_tesseract_config_sparsify_custom", "test_suggest_sparsify_reactivate_limit",
                          # rejected by the constructor's validation, which runs before a device is asked for
                          "test_sparsify_negative_sentinels_rejected", "test_decoder_compilation_validation"],
    "tesseract_sinter_compat_test.py": ["test_tesseract_sinter_obj_exists", "test_tesseract_sinter_decoder_sparsify_attributes",
                                        "test_tesseract_sinter_decoder_old_positional_constructor_order",
                                        "test_make_tesseract_sinter_decoders_dict_contains_sparsify"],
}


@pytest.fixture(scope="module")
def reference_modules():
    import stim_stand_in
    import tesseract_decoder_b200
    import types
    names = ("stim", "tesseract_decoder", "shared_decoding_tests", "sinter", "sinter._decoding", "sinter._decoding._decoding")
    saved = {k: sys.modules.get(k) for k in names}
    sys.modules["stim"] = stim_stand_in
    # tesseract_sinter_compat_test.py imports sinter at module level; only its decoding tests (not run here) use it
    for name in names[3:]:
        sys.modules[name] = types.ModuleType(name)
    sys.modules["sinter._decoding._decoding"].sample_decode = None
    sys.modules["sinter"]._decoding = sys.modules["sinter._decoding"]
    sys.modules["sinter._decoding"]._decoding = sys.modules["sinter._decoding._decoding"]
    sys.modules["tesseract_decoder"] = tesseract_decoder_b200
    sys.path.insert(0, REF_PY)  # tesseract_test.py imports shared_decoding_tests from its own directory
    sys.dont_write_bytecode, old_flag = True, sys.dont_write_bytecode  # never write __pycache__ into the reference tree
    mods = {}
    try:
        for fname in DEVICE_FREE:
            spec = importlib.util.spec_from_file_location("reference_" + fname[:-3], os.path.join(REF_PY, fname))
            mod = importlib.util.module_from_spec(spec)
            spec.loader.exec_module(mod)
            mods[fname] = mod
        yield mods
    finally:
        sys.dont_write_bytecode = old_flag
        sys.path.remove(REF_PY)
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v


@pytest.mark.parametrize("fname,test", [(f, t) for f, tests in DEVICE_FREE.items() for t in tests])
def test_reference_test_passes_unmodified(reference_modules, fname, test):
    fn = getattr(reference_modules[fname], test)
    marks = [m for m in getattr(fn, "pytestmark", []) if m.name == "parametrize"]
    if not marks:
        fn()
        return
    assert len(marks) == 1
    names = [n.strip() for n in marks[0].args[0].split(",")]
    for values in marks[0].args[1]:  # every parameter set of @pytest.mark.parametrize, as pytest would call it
        values = values if isinstance(values, (tuple, list)) and len(names) > 1 else (values,)
        fn(**dict(zip(names, values)))


def test_every_device_free_reference_test_is_listed(reference_modules):
    """The lists above are not a cherry-pick: every other test of these files decodes (needs the GPU) or belongs to the
    simplex decoder (out of scope); a test added upstream shows up here."""
    needs_device_or_out_of_scope = {
        "tesseract_test.py": {"test_compile_decoder_for_dem_basic_functionality", "test_compile_decoder_for_dem_sets_dem_on_config",
                              "test_compile_decoder_for_dem_preserves_other_config_params", "test_compile_decoder_for_dem_with_empty_dem",
                              "test_create_tesseract_decoder", "test_tesseract_compile_decoder", "test_tesseract_cost_from_errors",
                              "test_tesseract_get_observables_from_errors", "test_tesseract_decode_from_detection_events",
                              "test_tesseract_decoder_predicts_various_observable_flips", "test_tesseract_decode",
                              "test_tesseract_decode_complex_dem", "test_tesseract_decode_batch_with_invalid_dimensions",
                              "test_tesseract_decode_batch", "test_tesseract_decode_batch_with_complex_model",
                              "test_tesseract_merge_errors_affects_cost", "test_simlpex_decode_with_mismatched_syndrome_size",
                              "test_test_simplex_decode_batch_with_mismatched_syndrome_size",
                              "test_compile_decoder_resolves_auto_sparsify_reactivate_limit",
                              "test_compile_decoder_caps_auto_sparsify_reactivate_limit_at_error_count",
                              "test_compile_decoder_preserves_explicit_sparsify_reactivate_limit",
                              "test_python_sparsify_changes_predicted_error_set"},
        "tesseract_sinter_compat_test.py": {"test_compile_decoder_for_dem", "test_decode_shots_bit_packed", "test_decode_shots_bit_packed_multi_shot",
                                            "test_decode_via_files_sanity_check", "test_decode_via_files", "test_decode_via_files_multi_shot",
                                            "test_sinter_decode_repetition_code", "test_sinter_decode_surface_code", "test_sinter_empty",
                                            "test_sinter_no_observables", "test_sinter_invincible_observables", "test_sinter_detector_counting",
                                            "test_full_scale", "test_full_scale_one_worker", "test_decode_shots_bit_packed_vs_decode_batch",
                                            "test_sinter_collect_different_dems", "test_sinter_compile_sparsify_config_reaches_decoder",
                                            "test_sinter_decode_with_sparsify_decoders"},
    }
    for fname, mod in reference_modules.items():
        found = {n for n in dir(mod) if n.startswith("test_") and callable(getattr(mod, n))}
        assert found == set(DEVICE_FREE[fname]) | needs_device_or_out_of_scope.get(fname, set()), fname
