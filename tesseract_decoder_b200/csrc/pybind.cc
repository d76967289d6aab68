// pybind11 module `_tesseract_b200`: the reference's Python surface for the decode path
// (src/tesseract.pybind.h, src/utils.pybind.h, src/tesseract_sinter_compat.pybind.h) on top of decoder.h.
// Same class / method / argument names, defaults and exception texts; DEMs cross the boundary as
// str(dem), as in the reference (stim_utils.pybind.h:22-26), so python-stim is not required.
#include <pybind11/iostream.h>
#include <pybind11/numpy.h>
#include <pybind11/operators.h>
#include <pybind11/pybind11.h>
#include <pybind11/stl.h>

#include <fstream>

#include "decoder.h"

namespace py = pybind11;
using namespace tesseract_b200;

namespace {

std::string dem_text(const py::object& dem) {
  if (dem.is_none()) return std::string();
  return py::cast<std::string>(py::str(dem));
}

// A DEM leaving the module: stim.DetectorErrorModel(text) where python-stim is importable — what the reference returns
// (stim_utils.pybind.h:62-69) — and the text itself otherwise.
py::object dem_object(const std::string& text) {
  try {
    return py::module_::import("stim").attr("DetectorErrorModel")(text);
  } catch (py::error_already_set& e) {
    if (!e.matches(PyExc_ImportError)) throw;
    return py::str(text);
  }
}

// tesseract.pybind.h:38-73: an empty det_orders argument means 20 DetIndex orders, seed 2384753.
TesseractConfig make_config(const std::string& dem, int det_beam, bool beam_climbing, bool no_revisit_dets, bool verbose,
                            bool merge_errors, size_t pqlimit, std::vector<std::vector<size_t>> det_orders, double det_penalty,
                            bool create_visualization, bool at_most_two, bool sparsify_errors = false, int sparsify_base_degree = -1,
                            int sparsify_max_degree = -1, int sparsify_reactivate_limit = -1) {
  if (det_orders.empty()) det_orders = build_det_orders(dem, 20, DetOrder::DetIndex, 2384753);
  TesseractConfig c;
  c.dem = dem;
  c.det_beam = det_beam;
  c.beam_climbing = beam_climbing;
  c.no_revisit_dets = no_revisit_dets;
  c.verbose = verbose;
  c.merge_errors = merge_errors;
  c.pqlimit = pqlimit;
  c.det_orders = std::move(det_orders);
  c.det_penalty = det_penalty;
  c.create_visualization = create_visualization;
  c.at_most_two_errors_per_detector = at_most_two;
  c.sparsify_errors = sparsify_errors;
  c.sparsify_base_degree = sparsify_base_degree;
  c.sparsify_max_degree = sparsify_max_degree;
  c.sparsify_reactivate_limit = sparsify_reactivate_limit;
  return c;
}

std::vector<uint64_t> fired(const py::array_t<bool>& syndrome, size_t num_detectors) {
  if ((size_t)syndrome.size() != num_detectors)  // tesseract.pybind.h:264-269
    throw std::invalid_argument("Syndrome array size (" + std::to_string(syndrome.size()) +
                                ") does not match the number of detectors in the decoder (" + std::to_string(num_detectors) + ").");
  std::vector<uint64_t> detections;
  auto s = syndrome.unchecked<1>();
  for (py::ssize_t i = 0; i < s.size(); ++i)
    if (s(i)) detections.push_back((uint64_t)i);
  return detections;
}

py::array obs_array(const TesseractDecoder& self, const std::vector<size_t>& errs) {
  std::vector<char> result(self.num_observables, 0);
  for (int o : self.get_flipped_observables(errs)) result[(size_t)o] ^= 1;
  return py::array(py::dtype::of<bool>(), result.size(), result.data());
}

// ---- sinter (tesseract_sinter_compat.pybind.h) ------------------------------------------------------

struct TesseractSinterCompiledDecoder {
  std::shared_ptr<TesseractDecoder> decoder;
  uint64_t num_detectors = 0, num_observables = 0;

  py::array_t<uint8_t> decode_shots_bit_packed(const py::array_t<uint8_t>& data) {
    if (data.ndim() != 2) throw std::invalid_argument("Input `bit_packed_detection_event_data` must be a 2D array.");
    const uint64_t nb = (num_detectors + 7) / 8;
    if (data.shape(1) != (py::ssize_t)nb) throw std::invalid_argument("Input array's second dimension does not match num_detector_bytes.");
    const size_t shots = (size_t)data.shape(0);
    const uint64_t ob = (num_observables + 7) / 8;
    auto result = py::array_t<uint8_t>({(py::ssize_t)shots, (py::ssize_t)ob});
    if (data.strides(1) != 1) throw std::invalid_argument("bit_packed_detection_event_data rows must be contiguous.");
    {
      py::gil_scoped_release release;  // one GPU call for the whole batch
      decoder->decode_batch_b8(data.data(), (size_t)data.strides(0), shots, result.mutable_data(), nullptr, nullptr);
    }
    return result;
  }
};

struct TesseractSinterDecoder {
  int det_beam = DEFAULT_DET_BEAM;
  bool beam_climbing = false, no_revisit_dets = true, verbose = false, merge_errors = true;
  size_t pqlimit = DEFAULT_PQLIMIT;
  double det_penalty = 0;
  bool create_visualization = false;
  size_t num_det_orders = 0;
  DetOrder det_order_method = DetOrder::DetIndex;
  uint64_t seed = 2384753;
  // tesseract_sinter_compat.pybind.h:114-117
  bool sparsify_errors = false;
  int sparsify_base_degree = -1, sparsify_max_degree = -1, sparsify_reactivate_limit = -1;
  bool at_most_two_errors_per_detector = false;  // extension (README.md:59); last, so the reference's positional order holds

  TesseractSinterDecoder() = default;
  TesseractSinterDecoder(int det_beam, bool beam_climbing, bool no_revisit_dets, bool verbose, bool merge_errors, size_t pqlimit,
                         double det_penalty, bool create_visualization, size_t num_det_orders, DetOrder method, uint64_t seed,
                         bool sparsify_errors = false, int sparsify_base_degree = -1, int sparsify_max_degree = -1,
                         int sparsify_reactivate_limit = -1, bool at_most_two = false)
      : det_beam(det_beam), beam_climbing(beam_climbing), no_revisit_dets(no_revisit_dets), verbose(verbose),
        merge_errors(merge_errors), pqlimit(pqlimit), det_penalty(det_penalty), create_visualization(create_visualization),
        num_det_orders(num_det_orders), det_order_method(method), seed(seed), sparsify_errors(sparsify_errors),
        sparsify_base_degree(sparsify_base_degree), sparsify_max_degree(sparsify_max_degree),
        sparsify_reactivate_limit(sparsify_reactivate_limit), at_most_two_errors_per_detector(at_most_two) {}

  bool operator==(const TesseractSinterDecoder& o) const {
    return det_beam == o.det_beam && beam_climbing == o.beam_climbing && no_revisit_dets == o.no_revisit_dets &&
           verbose == o.verbose && merge_errors == o.merge_errors && pqlimit == o.pqlimit && det_penalty == o.det_penalty &&
           create_visualization == o.create_visualization && num_det_orders == o.num_det_orders &&
           det_order_method == o.det_order_method && seed == o.seed && sparsify_errors == o.sparsify_errors &&
           sparsify_base_degree == o.sparsify_base_degree && sparsify_max_degree == o.sparsify_max_degree &&
           sparsify_reactivate_limit == o.sparsify_reactivate_limit &&
           at_most_two_errors_per_detector == o.at_most_two_errors_per_detector;
  }
  bool operator!=(const TesseractSinterDecoder& o) const { return !(*this == o); }

  std::shared_ptr<TesseractDecoder> build(const std::string& dem) const {
    TesseractConfig c;
    c.dem = dem;
    c.det_beam = det_beam;
    c.beam_climbing = beam_climbing;
    c.no_revisit_dets = no_revisit_dets;
    c.verbose = verbose;
    c.merge_errors = merge_errors;
    c.pqlimit = pqlimit;
    c.det_penalty = det_penalty;
    c.create_visualization = create_visualization;
    c.at_most_two_errors_per_detector = at_most_two_errors_per_detector;
    c.sparsify_errors = sparsify_errors;
    c.sparsify_base_degree = sparsify_base_degree;
    c.sparsify_max_degree = sparsify_max_degree;
    c.sparsify_reactivate_limit = sparsify_reactivate_limit;
    c.det_orders = build_det_orders(dem, num_det_orders, det_order_method, seed);  // 0 -> the decoder's iota order
    return std::make_shared<TesseractDecoder>(std::move(c));
  }

  TesseractSinterCompiledDecoder compile_decoder_for_dem(const py::object& dem) const {  // :183-210
    auto d = build(dem_text(dem));
    TesseractSinterCompiledDecoder out;
    out.num_detectors = d->num_detectors;
    out.num_observables = d->num_observables;
    out.decoder = std::move(d);
    return out;
  }

  void decode_via_files(uint64_t num_shots, uint64_t num_dets, uint64_t num_obs, const py::object& dem_path,
                        const py::object& dets_b8_in_path, const py::object& obs_predictions_b8_out_path,
                        const py::object& /*tmp_dir*/) const {  // :213-298
    const std::string dem_path_s = py::cast<std::string>(py::str(dem_path));
    const std::string in_s = py::cast<std::string>(py::str(dets_b8_in_path));
    const std::string out_s = py::cast<std::string>(py::str(obs_predictions_b8_out_path));
    std::ifstream dem_file(dem_path_s);
    if (!dem_file) throw std::runtime_error("Failed to open DEM file: " + dem_path_s);
    std::stringstream ss;
    ss << dem_file.rdbuf();
    auto decoder = build(ss.str());
    const uint64_t nb = (num_dets + 7) / 8, ob = (num_obs + 7) / 8;
    std::ifstream input(in_s, std::ios::binary);
    if (!input) throw std::runtime_error("Failed to open input file: " + in_s);
    std::ofstream output(out_s, std::ios::binary);
    if (!output) throw std::runtime_error("Failed to open output file: " + out_s);
    std::vector<uint8_t> dets((size_t)(num_shots * nb)), obs((size_t)(num_shots * ob) + 1);
    input.read(reinterpret_cast<char*>(dets.data()), (std::streamsize)dets.size());
    if ((uint64_t)input.gcount() != num_shots * nb) throw std::runtime_error("Failed to read a full shot from the input file.");
    const uint64_t own_ob = (decoder->num_observables + 7) / 8;
    std::vector<uint8_t> own((size_t)(num_shots * own_ob) + 1);
    {
      py::gil_scoped_release release;
      decoder->decode_batch_b8(dets.data(), nb, num_shots, own.data(), nullptr, nullptr);
    }
    for (uint64_t s = 0; s < num_shots; ++s)  // keep only observables < num_obs, like :283-287
      for (uint64_t o = 0; o < std::min<uint64_t>(num_obs, decoder->num_observables); ++o)
        if ((own[s * own_ob + o / 8] >> (o % 8)) & 1) obs[s * ob + o / 8] ^= (uint8_t)(1u << (o % 8));
    output.write(reinterpret_cast<char*>(obs.data()), (std::streamsize)(num_shots * ob));
  }
};

}  // namespace

PYBIND11_MODULE(_tesseract_b200, root) {
  root.doc() = "B200-native Tesseract decode path (CUDA through libtesseract_b200.so); mirrors tesseract_decoder.";

  // ---- utils (utils.pybind.h) --------------------------------------------------------------------
  auto u = root.def_submodule("utils", "utility functions");
  py::enum_<DetOrder>(u, "DetOrder")
      .value("DetBFS", DetOrder::DetBFS)
      .value("DetIndex", DetOrder::DetIndex)
      .value("DetCoordinate", DetOrder::DetCoordinate)
      .export_values();
  u.def(
      "build_det_orders",
      [](const py::object& dem, size_t num_det_orders, DetOrder method, uint64_t seed) {
        return build_det_orders(dem_text(dem), num_det_orders, method, seed);
      },
      py::arg("dem"), py::arg("num_det_orders"), py::arg("method") = DetOrder::DetIndex, py::arg("seed") = 0);  // utils.pybind.h:95-96
  u.def("get_detector_coords", [](const py::object& dem) { return get_detector_coords(dem_text(dem)); }, py::arg("dem"));    // utils.pybind.h:43-61
  u.def("build_detector_graph", [](const py::object& dem) { return build_detector_graph(dem_text(dem)); }, py::arg("dem"));  // utils.pybind.h:63-86
  u.def("get_errors_from_dem", [](const py::object& dem) { return get_errors_from_dem(dem_text(dem)); }, py::arg("dem"));    // utils.pybind.h:118-135
  u.def("circuit_to_dem", [](const std::string& circuit) {
    char* out = nullptr;
    check(tsr_circuit_to_dem(circuit.c_str(), &out));
    std::string s(out);
    tsr_free(out);
    return s;
  }, py::arg("circuit_text"), "Detector error model text of a stim circuit (stand-in for stim's ErrorAnalyzer).");
  u.attr("INF") = std::numeric_limits<double>::infinity();
  u.attr("EPSILON") = 1e-7;

  // ---- common (common.pybind.h, the parts the decoder exposes) --------------------------------------
  auto cm = root.def_submodule("common", "classes commonly used by the decoder");
  py::class_<Symptom>(cm, "Symptom")  // common.pybind.h:33-75
      .def(py::init<std::vector<int>, std::vector<int>>(), py::arg("detectors") = std::vector<int>(),
           py::arg("observables") = std::vector<int>())
      .def_readwrite("detectors", &Symptom::detectors)
      .def_readwrite("observables", &Symptom::observables)
      .def("__str__", &Symptom::str)
      .def(py::self == py::self)
      .def(py::self != py::self)
      .def(
          "as_dem_instruction_targets",
          [](const Symptom& s) {  // stim.DemTarget objects where python-stim is importable, their text ("D1", "L2") otherwise
            py::list out;
            py::object make = py::none();
            try {
              make = py::module_::import("stim").attr("DemTarget");
            } catch (py::error_already_set&) {
            }
            auto add = [&](const std::string& t) {
              if (make.is_none())
                out.append(py::str(t));
              else
                out.append(make(t));
            };
            for (int d : s.detectors) add("D" + std::to_string(d));
            for (int o : s.observables) add("L" + std::to_string(o));
            return out;
          });
  py::class_<Error>(cm, "Error")  // common.pybind.h:77-135
      .def(py::init<>())
      .def(py::init<double, std::vector<int>, std::vector<int>>(), py::arg("likelihood_cost"), py::arg("detectors"),
           py::arg("observables"))
      .def(py::init([](const py::object& error) {  // a stim.DemInstruction, or its text: "error(0.125) D0 D1 L0"
             const std::vector<Error> parsed = get_errors_from_dem(py::cast<std::string>(py::str(error)));
             if (parsed.size() != 1)
               throw std::invalid_argument("Error must be loaded from an error dem instruction, but received: " +
                                           py::cast<std::string>(py::str(error)));
             return parsed[0];
           }),
           py::arg("error"))
      .def_readwrite("likelihood_cost", &Error::likelihood_cost)
      .def_readwrite("symptom", &Error::symptom)
      .def("__str__", &Error::str)
      .def("get_probability", &Error::get_probability)
      .def("set_with_probability", &Error::set_with_probability, py::arg("probability"));
  // common.pybind.h:143-191: DEM in (anything whose str() is DEM text), DEM text out
  cm.def(
      "merge_indistinguishable_errors",
      [](const py::object& dem) {
        std::vector<size_t> map;
        return dem_object(merge_indistinguishable_errors(dem_text(dem), map));
      },
      py::arg("dem"));
  cm.def(
      "remove_zero_probability_errors",
      [](const py::object& dem) {
        std::vector<size_t> map;
        return dem_object(remove_zero_probability_errors(dem_text(dem), map));
      },
      py::arg("dem"));
  cm.def("merge_weights", &tsr_merge_weights, py::arg("a"), py::arg("b"));
  cm.def(
      "dem_from_counts",
      [](const py::object& dem, const std::vector<uint64_t>& error_counts, uint64_t num_shots) {
        char* out = nullptr;
        check(tsr_dem_from_counts(dem_text(dem).c_str(), error_counts.data(), error_counts.size(), num_shots, &out));
        std::string text(out);
        tsr_free(out);
        return dem_object(text);
      },
      py::arg("orig_dem"), py::arg("error_counts"), py::arg("num_shots"));

  // ---- tesseract (tesseract.pybind.h) --------------------------------------------------------------
  auto m = root.def_submodule("tesseract", "Module containing the tesseract algorithm");
  m.attr("INF_DET_BEAM") = INF_DET_BEAM;
  m.def(
      "suggest_sparsify_reactivate_limit",  // tesseract.pybind.h:80-82
      [](size_t num_detectors, int sparsify_base_degree) {
        const int64_t v = tsr_suggest_sparsify_reactivate_limit(num_detectors, sparsify_base_degree);
        if (v < 0) throw std::invalid_argument(tsr_last_error());
        return (int)v;
      },
      py::arg("num_detectors"), py::arg("sparsify_base_degree"),
      "Returns the suggested number of optional high-degree errors to reactivate per shot.");

  py::class_<tesseract_b200::Visualizer>(root.def_submodule("viz", "Module containing the visualization tools"), "Visualizer")  // visualization.pybind.h
      .def(py::init<>())  // visualization.pybind.h:16: a free-standing Visualizer writes an empty log
      .def("write", &tesseract_b200::Visualizer::write, py::arg("fpath"))
      .def("text", &tesseract_b200::Visualizer::text);

  py::class_<TesseractConfig>(m, "TesseractConfig")
      .def(py::init<>())
      .def(py::init([](int det_beam, bool beam_climbing, bool no_revisit_dets, bool verbose, bool merge_errors, size_t pqlimit,
                       std::vector<std::vector<size_t>> det_orders, double det_penalty, bool create_visualization, bool sparsify_errors,
                       int sparsify_base_degree, int sparsify_max_degree, int sparsify_reactivate_limit, bool at_most_two) {
             return make_config("", det_beam, beam_climbing, no_revisit_dets, verbose, merge_errors, pqlimit, std::move(det_orders),
                                det_penalty, create_visualization, at_most_two, sparsify_errors, sparsify_base_degree,
                                sparsify_max_degree, sparsify_reactivate_limit);
           }),
           py::arg("det_beam") = 5, py::arg("beam_climbing") = false, py::arg("no_revisit_dets") = true, py::arg("verbose") = false,
           py::arg("merge_errors") = true, py::arg("pqlimit") = 200000, py::arg("det_orders") = std::vector<std::vector<size_t>>(),
           py::arg("det_penalty") = 0.0, py::arg("create_visualization") = false, py::arg("sparsify_errors") = false,
           py::arg("sparsify_base_degree") = -1, py::arg("sparsify_max_degree") = -1, py::arg("sparsify_reactivate_limit") = -1,
           py::arg("at_most_two_errors_per_detector") = false)
      .def(py::init([](const py::object& dem, int det_beam, bool beam_climbing, bool no_revisit_dets, bool verbose, bool merge_errors,
                       size_t pqlimit, std::vector<std::vector<size_t>> det_orders, double det_penalty, bool create_visualization,
                       bool sparsify_errors, int sparsify_base_degree, int sparsify_max_degree, int sparsify_reactivate_limit,
                       bool at_most_two) {
             return make_config(dem_text(dem), det_beam, beam_climbing, no_revisit_dets, verbose, merge_errors, pqlimit,
                                std::move(det_orders), det_penalty, create_visualization, at_most_two, sparsify_errors,
                                sparsify_base_degree, sparsify_max_degree, sparsify_reactivate_limit);
           }),
           py::arg("dem"), py::arg("det_beam") = 5, py::arg("beam_climbing") = false, py::arg("no_revisit_dets") = true,
           py::arg("verbose") = false, py::arg("merge_errors") = true, py::arg("pqlimit") = 200000,
           py::arg("det_orders") = std::vector<std::vector<size_t>>(), py::arg("det_penalty") = 0.0,
           py::arg("create_visualization") = false, py::arg("sparsify_errors") = false, py::arg("sparsify_base_degree") = -1,
           py::arg("sparsify_max_degree") = -1, py::arg("sparsify_reactivate_limit") = -1,
           py::arg("at_most_two_errors_per_detector") = false)
      .def_property(
          "dem", [](const TesseractConfig& c) { return dem_object(c.dem); },
          [](TesseractConfig& c, const py::object& dem) { c.dem = dem_text(dem); },
          "The detector error model: a stim.DetectorErrorModel where python-stim is importable, its text otherwise.")
      .def_readwrite("det_beam", &TesseractConfig::det_beam)
      .def_readwrite("beam_climbing", &TesseractConfig::beam_climbing)
      .def_readwrite("no_revisit_dets", &TesseractConfig::no_revisit_dets)
      .def_readwrite("verbose", &TesseractConfig::verbose)
      .def_readwrite("merge_errors", &TesseractConfig::merge_errors)
      .def_readwrite("pqlimit", &TesseractConfig::pqlimit)
      .def_readwrite("det_orders", &TesseractConfig::det_orders)
      .def_readwrite("det_penalty", &TesseractConfig::det_penalty)
      .def_readwrite("create_visualization", &TesseractConfig::create_visualization)
      .def_readwrite("at_most_two_errors_per_detector", &TesseractConfig::at_most_two_errors_per_detector)
      .def_readwrite("sparsify_errors", &TesseractConfig::sparsify_errors)
      .def_readwrite("sparsify_base_degree", &TesseractConfig::sparsify_base_degree)
      .def_readwrite("sparsify_max_degree", &TesseractConfig::sparsify_max_degree)
      .def_readwrite("sparsify_reactivate_limit", &TesseractConfig::sparsify_reactivate_limit)
      .def_readwrite("device", &TesseractConfig::device, "CUDA device ordinal (-1 = current device)")
      .def("__str__", &TesseractConfig::str)
      .def("compile_decoder", [](const TesseractConfig& self) { return std::make_unique<TesseractDecoder>(self); })
      .def(
          "compile_decoder_for_dem",
          [](TesseractConfig& self, const py::object& dem) {
            self.dem = dem_text(dem);
            return std::make_unique<TesseractDecoder>(self);
          },
          py::arg("dem"));

  py::class_<TesseractDecoder>(m, "TesseractDecoder")
      .def(py::init([](const TesseractConfig& c) { return std::make_unique<TesseractDecoder>(c); }), py::arg("config"))
      .def(
          "decode_to_errors",
          [](TesseractDecoder& self, const py::array_t<bool>& syndrome) {
            self.decode_to_errors(fired(syndrome, self.num_detectors));
            return self.predicted_errors_buffer;
          },
          py::arg("syndrome"))
      .def(
          "decode_to_errors",
          [](TesseractDecoder& self, const py::array_t<bool>& syndrome, size_t det_order, size_t det_beam) {
            self.decode_to_errors(fired(syndrome, self.num_detectors), det_order, det_beam);
            return self.predicted_errors_buffer;
          },
          py::arg("syndrome"), py::arg("det_order"), py::arg("det_beam"))
      .def(
          "get_observables_from_errors",
          [](TesseractDecoder& self, const std::vector<size_t>& predicted_errors) {
            std::vector<bool> result(self.num_observables, false);
            for (int o : self.get_flipped_observables(predicted_errors)) result[(size_t)o] = !result[(size_t)o];
            return result;
          },
          py::arg("predicted_errors"))
      .def("cost_from_errors", &TesseractDecoder::cost_from_errors, py::arg("predicted_errors"))
      .def(
          "decode_from_detection_events",
          [](TesseractDecoder& self, const std::vector<uint64_t>& detections) {
            self.decode(detections);
            return obs_array(self, self.predicted_errors_buffer);
          },
          py::arg("detections"))
      .def(
          "decode",
          [](TesseractDecoder& self, const py::array_t<bool>& syndrome) {
            self.decode(fired(syndrome, self.num_detectors));
            return obs_array(self, self.predicted_errors_buffer);
          },
          py::arg("syndrome"))
      .def(
          "decode_batch",
          [](TesseractDecoder& self, const py::array_t<bool>& syndromes) {
            if (syndromes.ndim() != 2) throw std::runtime_error("Input syndromes must be a 2D NumPy array.");
            auto s = syndromes.unchecked<2>();
            const size_t shots = (size_t)s.shape(0), nd = (size_t)s.shape(1);
            if (nd != self.num_detectors)
              throw std::invalid_argument("The number of detectors in the input array (" + std::to_string(nd) +
                                          ") does not match the number of detectors in the decoder (" +
                                          std::to_string(self.num_detectors) + ").");
            if (self.config.create_visualization) {  // the log is written by single-shot decodes: loop like the reference
              py::array_t<bool> result({shots, self.num_observables});
              auto r = result.mutable_unchecked<2>();
              for (size_t i = 0; i < shots; ++i) {
                std::vector<uint64_t> hits;
                for (size_t j = 0; j < nd; ++j)
                  if (s((py::ssize_t)i, (py::ssize_t)j)) hits.push_back(j);
                std::vector<bool> row(self.num_observables, false);
                for (int o : self.decode(hits)) row[(size_t)o] = true;
                for (size_t o = 0; o < self.num_observables; ++o) r((py::ssize_t)i, (py::ssize_t)o) = row[o];
              }
              return result;
            }
            // pack -> one GPU batch -> unpack (the reference loops decode() per row, tesseract.pybind.h:476-490)
            const size_t stride = std::max<size_t>((nd + 7) / 8, 1), ob = (self.num_observables + 7) / 8;
            std::vector<uint8_t> packed(shots * stride, 0), obs(shots * ob + 1, 0);
            for (size_t i = 0; i < shots; ++i)
              for (size_t j = 0; j < nd; ++j)
                if (s((py::ssize_t)i, (py::ssize_t)j)) packed[i * stride + j / 8] |= (uint8_t)(1u << (j % 8));
            {
              py::gil_scoped_release release;
              self.decode_batch_b8(packed.data(), stride, shots, obs.data(), nullptr, nullptr);
            }
            py::array_t<bool> result({shots, self.num_observables});
            auto r = result.mutable_unchecked<2>();
            for (size_t i = 0; i < shots; ++i)
              for (size_t o = 0; o < self.num_observables; ++o)
                r((py::ssize_t)i, (py::ssize_t)o) = (obs[i * ob + o / 8] >> (o % 8)) & 1;
            return result;
          },
          py::arg("syndromes"))
      .def("get_eneighbors", &TesseractDecoder::get_eneighbors)
      .def_readwrite("config", &TesseractDecoder::config)
      .def_readwrite("low_confidence_flag", &TesseractDecoder::low_confidence_flag)
      .def_readwrite("predicted_errors_buffer", &TesseractDecoder::predicted_errors_buffer)
      .def_readwrite("errors", &TesseractDecoder::errors)
      .def_readwrite("num_observables", &TesseractDecoder::num_observables)
      .def_readwrite("num_detectors", &TesseractDecoder::num_detectors)
      .def_readonly("visualizer", &TesseractDecoder::visualizer)  // tesseract.pybind.h:525-527
      .def_readonly("dem_error_to_error", &TesseractDecoder::dem_error_to_error)
      .def_readonly("error_to_dem_error", &TesseractDecoder::error_to_dem_error);

  // ---- sinter (tesseract_sinter_compat.pybind.h) ---------------------------------------------------------
  auto sc = root.def_submodule("tesseract_sinter_compat", "sinter.Decoder / sinter.CompiledDecoder for the B200 decoder");
  py::class_<TesseractSinterCompiledDecoder>(sc, "TesseractSinterCompiledDecoder")
      .def("decode_shots_bit_packed", &TesseractSinterCompiledDecoder::decode_shots_bit_packed, py::kw_only(),
           py::arg("bit_packed_detection_event_data"))
      .def_readwrite("num_detectors", &TesseractSinterCompiledDecoder::num_detectors)
      .def_readwrite("num_observables", &TesseractSinterCompiledDecoder::num_observables)
      .def_property_readonly(
          "decoder", [](const TesseractSinterCompiledDecoder& self) -> TesseractDecoder& { return *self.decoder; },
          py::return_value_policy::reference_internal);
  py::class_<TesseractSinterDecoder>(sc, "TesseractSinterDecoder")
      .def(py::init<>())
      .def(py::init<int, bool, bool, bool, bool, size_t, double, bool, size_t, DetOrder, uint64_t, bool, int, int, int, bool>(),
           py::arg("det_beam") = DEFAULT_DET_BEAM, py::arg("beam_climbing") = false, py::arg("no_revisit_dets") = true,
           py::arg("verbose") = false, py::arg("merge_errors") = true, py::arg("pqlimit") = DEFAULT_PQLIMIT,
           py::arg("det_penalty") = 0.0, py::arg("create_visualization") = false, py::arg("num_det_orders") = 0,
           py::arg("det_order_method") = DetOrder::DetIndex, py::arg("seed") = 2384753, py::arg("sparsify_errors") = false,
           py::arg("sparsify_base_degree") = -1, py::arg("sparsify_max_degree") = -1, py::arg("sparsify_reactivate_limit") = -1,
           py::arg("at_most_two_errors_per_detector") = false)
      .def("compile_decoder_for_dem", &TesseractSinterDecoder::compile_decoder_for_dem, py::kw_only(), py::arg("dem"))
      .def("decode_via_files", &TesseractSinterDecoder::decode_via_files, py::kw_only(), py::arg("num_shots"), py::arg("num_dets"),
           py::arg("num_obs"), py::arg("dem_path"), py::arg("dets_b8_in_path"), py::arg("obs_predictions_b8_out_path"),
           py::arg("tmp_dir"))
      .def_readwrite("det_beam", &TesseractSinterDecoder::det_beam)
      .def_readwrite("beam_climbing", &TesseractSinterDecoder::beam_climbing)
      .def_readwrite("no_revisit_dets", &TesseractSinterDecoder::no_revisit_dets)
      .def_readwrite("verbose", &TesseractSinterDecoder::verbose)
      .def_readwrite("merge_errors", &TesseractSinterDecoder::merge_errors)
      .def_readwrite("pqlimit", &TesseractSinterDecoder::pqlimit)
      .def_readwrite("det_penalty", &TesseractSinterDecoder::det_penalty)
      .def_readwrite("create_visualization", &TesseractSinterDecoder::create_visualization)
      .def_readwrite("num_det_orders", &TesseractSinterDecoder::num_det_orders)
      .def_readwrite("det_order_method", &TesseractSinterDecoder::det_order_method)
      .def_readwrite("seed", &TesseractSinterDecoder::seed)
      .def_readwrite("sparsify_errors", &TesseractSinterDecoder::sparsify_errors)
      .def_readwrite("sparsify_base_degree", &TesseractSinterDecoder::sparsify_base_degree)
      .def_readwrite("sparsify_max_degree", &TesseractSinterDecoder::sparsify_max_degree)
      .def_readwrite("sparsify_reactivate_limit", &TesseractSinterDecoder::sparsify_reactivate_limit)
      .def_readwrite("at_most_two_errors_per_detector", &TesseractSinterDecoder::at_most_two_errors_per_detector)
      .def(py::self == py::self)
      .def(py::self != py::self)
      .def(py::pickle(
          // the reference's 15-tuple (tesseract_sinter_compat.pybind.h:406-439) plus at_most_two_errors_per_detector
          [](const TesseractSinterDecoder& s) {
            return py::make_tuple(s.det_beam, s.beam_climbing, s.no_revisit_dets, s.verbose, s.merge_errors, s.pqlimit,
                                  s.det_penalty, s.create_visualization, s.num_det_orders, s.det_order_method, s.seed,
                                  s.sparsify_errors, s.sparsify_base_degree, s.sparsify_max_degree, s.sparsify_reactivate_limit,
                                  s.at_most_two_errors_per_detector);
          },
          [](py::tuple t) {
            auto head = [&](size_t orders_at, bool sp, int base, int max, int limit, bool amt) {
              return TesseractSinterDecoder(t[0].cast<int>(), t[1].cast<bool>(), t[2].cast<bool>(), t[3].cast<bool>(),
                                            t[4].cast<bool>(), t[5].cast<size_t>(), t[6].cast<double>(), t[7].cast<bool>(),
                                            t[orders_at].cast<size_t>(), t[orders_at + 1].cast<DetOrder>(),
                                            t[orders_at + 2].cast<uint64_t>(), sp, base, max, limit, amt);
            };
            if (t.size() == 11) return head(8, false, -1, -1, -1, false);                 // legacy 11-tuple
            if (t.size() == 12) return head(8, false, -1, -1, -1, t[11].cast<bool>());    // this module's round-1 state
            if (t.size() != 15 && t.size() != 16) throw std::runtime_error("Invalid state for TesseractSinterDecoder!");
            const bool amt = t.size() == 16 ? t[15].cast<bool>() : false;
            if (py::isinstance<py::bool_>(t[8]))  // legacy order with the sparsify fields at [8..11]
              return head(12, t[8].cast<bool>(), t[9].cast<int>(), t[10].cast<int>(), t[11].cast<int>(), amt);
            return head(8, t[11].cast<bool>(), t[12].cast<int>(), t[13].cast<int>(), t[14].cast<int>(), amt);
          }));
  // Presets of make_tesseract_sinter_decoders_dict (tesseract_sinter_compat.pybind.h:442-490).
  sc.def("make_tesseract_sinter_decoders_dict", []() {
    py::dict d;
    auto preset = [](int beam, size_t pqlimit, size_t orders, bool sparsify, int base) {
      return TesseractSinterDecoder(beam, true, true, false, true, pqlimit, 0.0, false, orders, DetOrder::DetIndex, 2384753, sparsify,
                                    base, -1, -1, false);
    };
    d["tesseract-long-beam"] = preset(20, 1000000, 21, false, -1);
    d["tesseract"] = d["tesseract-long-beam"];
    d["tesseract-long-beam-sparsify-color-code-like"] = preset(20, 1000000, 21, true, 3);
    d["tesseract-long-beam-sparsify-surface-code-like"] = preset(20, 1000000, 21, true, 2);
    d["tesseract-short-beam"] = preset(15, 200000, 16, false, -1);
    d["tesseract-short-beam-sparsify-color-code-like"] = preset(15, 200000, 16, true, 3);
    d["tesseract-short-beam-sparsify-surface-code-like"] = preset(15, 200000, 16, true, 2);
    return d;
  });

  // tesseract.pybind.cc:39-46: `with tesseract_decoder.ostream_redirect(stdout=..., stderr=...)` routes the C++ side's
  // stdout / stderr (the verbose search log, the merge warning) to Python's.
  py::add_ostream_redirect(root, "ostream_redirect");
}
