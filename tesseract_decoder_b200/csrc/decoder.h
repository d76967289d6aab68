// Host-side mirror of the reference's decoder interface (tesseract.h:33-147) over the C ABI
// (include/tesseract_b200.h).  Same names, argument meaning and error behaviour; the search runs on the GPU.
// Header-only C++20; used by the pybind11 module (pybind.cc) and the CLI (main.cc).
#pragma once
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <iomanip>
#include <limits>
#include <memory>
#include <sstream>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "tesseract_b200.h"

namespace tesseract_b200 {

constexpr int INF_DET_BEAM = 65535;       // tesseract.h:33 (std::numeric_limits<uint16_t>::max())
constexpr int DEFAULT_DET_BEAM = 5;       // tesseract.h:34
constexpr size_t DEFAULT_PQLIMIT = 200000;  // tesseract.h:36

[[noreturn]] inline void throw_last(int rc) {
  const std::string msg = tsr_last_error();
  if (rc == TSR_E_INVALID_ARGUMENT) throw std::invalid_argument(msg);
  throw std::runtime_error(msg);
}
inline void check(int rc) {
  if (rc != TSR_OK) throw_last(rc);
}

// tesseract.h:39-58.  The DEM is held as text (what crosses the Python boundary in the reference too,
// stim_utils.pybind.h:22-26).
struct TesseractConfig {
  std::string dem;
  int det_beam = DEFAULT_DET_BEAM;
  bool beam_climbing = false;
  bool no_revisit_dets = true;
  bool verbose = false;
  bool merge_errors = true;
  size_t pqlimit = DEFAULT_PQLIMIT;
  std::vector<std::vector<size_t>> det_orders;
  double det_penalty = 0;
  bool create_visualization = false;
  bool at_most_two_errors_per_detector = false;  // README.md:59,159 (no code in the reference snapshot)
  // tesseract.h:52-55; validated by tsr_create like tesseract.cc:256-277 (same messages).  After construction the
  // decoder's copy holds the effective sparsify_reactivate_limit, as in the reference.
  bool sparsify_errors = false;
  int sparsify_base_degree = -1;
  int sparsify_max_degree = -1;
  int sparsify_reactivate_limit = -1;
  int device = -1;

  std::string str() const {  // tesseract.cc:84-117
    std::stringstream ss;
    ss << "TesseractConfig(dem=DetectorErrorModel_Object, det_beam=" << det_beam << ", no_revisit_dets=" << no_revisit_dets
       << ", verbose=" << verbose << ", merge_errors=" << merge_errors << ", pqlimit=" << pqlimit
       << ", det_penalty=" << det_penalty << ", create_visualization=" << create_visualization << ")";
    return ss.str();
  }
};

// common::Symptom / common::Error (common.h:25-75, common.cc:22-105).
struct Symptom {
  std::vector<int> detectors;
  std::vector<int> observables;
  Symptom() = default;
  Symptom(std::vector<int> detectors_, std::vector<int> observables_)
      : detectors(std::move(detectors_)), observables(std::move(observables_)) {}
  bool operator==(const Symptom& other) const { return detectors == other.detectors && observables == other.observables; }
  bool operator!=(const Symptom& other) const { return !(*this == other); }
  std::string str() const {  // common.cc:22-50
    auto list = [](const std::vector<int>& v) {
      std::string s = "[";
      for (size_t i = 0; i < v.size(); ++i) s += (i ? " " : "") + std::to_string(v[i]);
      return s + "]";
    };
    return "Symptom{detectors=" + list(detectors) + ", observables=" + list(observables) + "}";
  }
};
struct Error {
  double likelihood_cost = 0;
  Symptom symptom;
  Error() = default;
  Error(double likelihood_cost_, std::vector<int> detectors, std::vector<int> observables)
      : likelihood_cost(likelihood_cost_), symptom(std::move(detectors), std::move(observables)) {}
  std::string str() const {  // common.cc:88-94
    std::stringstream ss;
    ss << std::fixed << std::setprecision(6) << likelihood_cost;
    return "Error{cost=" + ss.str() + ", symptom=" + symptom.str() + "}";
  }
  double get_probability() const { return 1.0 / (1.0 + std::exp(likelihood_cost)); }  // common.cc:96-98
  void set_with_probability(double p) {                                               // common.cc:100-105
    if (p <= 0 || p >= 1) throw std::invalid_argument("Probability must be between 0 and 1.");
    likelihood_cost = -std::log(p / (1.0 - p));
  }
};

// TesseractConfig -> tsr_config (`flat` keeps the det_orders storage alive for the tsr_create call).
inline void fill_config(const TesseractConfig& config, tsr_config& c, std::vector<uint64_t>& flat) {
  tsr_config_init(&c);
  c.det_beam = config.det_beam;
  c.beam_climbing = config.beam_climbing;
  c.no_revisit_dets = config.no_revisit_dets;
  c.merge_errors = config.merge_errors;
  c.at_most_two_errors_per_detector = config.at_most_two_errors_per_detector;
  c.device = config.device;
  c.pqlimit = config.pqlimit;
  c.det_penalty = config.det_penalty;
  c.sparsify_errors = config.sparsify_errors;
  c.sparsify_base_degree = config.sparsify_base_degree;
  c.sparsify_max_degree = config.sparsify_max_degree;
  c.sparsify_reactivate_limit = config.sparsify_reactivate_limit;
  c.create_visualization = config.create_visualization;
  for (const auto& o : config.det_orders) {
    if (o.size() != config.det_orders[0].size())
      throw std::invalid_argument("Each detector order list must have a size equal to the number of detectors.");
    flat.insert(flat.end(), o.begin(), o.end());
  }
  c.num_det_orders = config.det_orders.size();
  c.det_orders = flat.data();
  c.det_order_len = config.det_orders.empty() ? 0 : config.det_orders[0].size();
  if (c.num_det_orders && c.det_order_len == 0 && tsr_dem_count_detectors(config.dem.c_str()) != 0)
    throw std::invalid_argument("Each detector order list must have a size equal to the number of detectors.");
}

struct HandleDeleter {
  void operator()(tsr_decoder* h) const { tsr_destroy(h); }
};

// visualization.h:12-23: the log the decoder fills when config.create_visualization is set.
struct Visualizer {
  tsr_decoder* handle = nullptr;  // borrowed from the decoder that owns this object
  std::string text() const {
    if (!handle) return "";
    char* out = nullptr;
    check(tsr_visualization_text(handle, &out));
    std::string t(out);
    tsr_free(out);
    return t;
  }
  void write(const char* fpath) const {
    FILE* f = std::fopen(fpath, "w");
    if (!f) throw std::invalid_argument(std::string("Failed to open ") + fpath);
    const std::string t = text();
    std::fwrite(t.data(), 1, t.size(), f);
    std::fclose(f);
  }
};

struct TesseractDecoder {
  TesseractConfig config;
  Visualizer visualizer;  // tesseract.h:84
  // tesseract.h:112-130
  bool low_confidence_flag = false;
  std::vector<size_t> predicted_errors_buffer;
  std::vector<Error> errors;
  size_t num_detectors = 0, num_observables = 0, num_errors = 0;
  std::vector<size_t> dem_error_to_error, error_to_dem_error;

  explicit TesseractDecoder(TesseractConfig cfg, bool host_only = false) : config(std::move(cfg)) {
    tsr_config c;
    std::vector<uint64_t> flat;
    fill_config(config, c, flat);
    tsr_decoder* raw = nullptr;
    check(tsr_create(config.dem.c_str(), &c, host_only ? TSR_CREATE_HOST_ONLY : 0, &raw));
    h_.reset(raw);
    visualizer.handle = raw;
    num_detectors = tsr_num_detectors(raw);
    num_observables = tsr_num_observables(raw);
    num_errors = tsr_num_errors(raw);
    if (config.sparsify_errors) config.sparsify_reactivate_limit = (int)tsr_sparsify_reactivate_limit(raw);  // tesseract.cc:272-277
    if (config.det_orders.empty()) {  // tesseract.cc:175-177
      config.det_orders.emplace_back(num_detectors);
      for (size_t i = 0; i < num_detectors; ++i) config.det_orders[0][i] = i;
    }
    // errors, maps (tesseract.cc:172-173, 189)
    std::vector<double> costs(num_errors);
    check(tsr_error_costs(raw, costs.data()));
    std::vector<uint64_t> d2e(tsr_num_dem_errors(raw)), e2d(num_errors);
    check(tsr_maps(raw, d2e.data(), e2d.data()));
    dem_error_to_error.assign(d2e.begin(), d2e.end());
    error_to_dem_error.assign(e2d.begin(), e2d.end());
    errors.resize(num_errors);
    auto fill = [&](int which, bool dets) {
      std::vector<uint64_t> off(num_errors + 1);
      const int64_t nnz = tsr_csr(raw, which, nullptr, nullptr);
      std::vector<int32_t> idx((size_t)std::max<int64_t>(nnz, 1));
      tsr_csr(raw, which, off.data(), idx.data());
      for (size_t e = 0; e < num_errors; ++e) {
        auto& v = dets ? errors[e].symptom.detectors : errors[e].symptom.observables;
        v.assign(idx.begin() + (std::ptrdiff_t)off[e], idx.begin() + (std::ptrdiff_t)off[e + 1]);
      }
    };
    fill(1, true);
    fill(3, false);
    for (size_t e = 0; e < num_errors; ++e) errors[e].likelihood_cost = costs[e];
  }

  tsr_decoder* handle() const { return h_.get(); }

  // tesseract.cc:295-347
  void decode_to_errors(const std::vector<uint64_t>& detections) { run(detections, -1, 0); }
  // tesseract.cc:371-378
  void decode_to_errors(const std::vector<uint64_t>& detections, size_t detector_order, size_t detector_beam) {
    run(detections, (int64_t)detector_order, detector_beam);
  }
  // tesseract.cc:660-663
  std::vector<int> decode(const std::vector<uint64_t>& detections) {
    decode_to_errors(detections);
    return get_flipped_observables(predicted_errors_buffer);
  }
  // tesseract.cc:618-628
  double cost_from_errors(const std::vector<size_t>& predicted_errors) const {
    std::vector<uint64_t> e(predicted_errors.begin(), predicted_errors.end());
    double out = 0;
    check(tsr_cost_from_errors(h_.get(), e.data(), e.size(), &out));
    return out;
  }
  // tesseract.cc:630-658
  std::vector<int> get_flipped_observables(const std::vector<size_t>& predicted_errors) const {
    std::vector<uint64_t> e(predicted_errors.begin(), predicted_errors.end());
    std::vector<uint8_t> bits((num_observables + 7) / 8 + 1, 0);
    check(tsr_flipped_observables(h_.get(), e.data(), e.size(), bits.data()));
    std::vector<int> out;
    for (size_t o = 0; o < num_observables; ++o)
      if ((bits[o >> 3] >> (o & 7)) & 1) out.push_back((int)o);
    return out;
  }
  std::vector<std::vector<int>> get_eneighbors() const { return csr_rows(2, num_errors); }
  std::vector<std::vector<int>> get_d2e() const { return csr_rows(0, num_detectors); }

  // The batched hot path: decode_shots (tesseract.cc:665-671) / decode_batch / decode_shots_bit_packed on
  // bit-packed rows (bit k of byte j = detector 8j+k).  Any output may be null.
  void decode_batch_b8(const uint8_t* dets, size_t stride, size_t shots, uint8_t* obs_out, double* cost_out,
                       uint8_t* low_conf_out, bool collect_errors = false) {
    tsr_set_collect_errors(h_.get(), collect_errors ? 1 : 0);
    const int rc = tsr_decode_batch_b8(h_.get(), dets, stride, shots, obs_out, cost_out, low_conf_out);
    tsr_set_collect_errors(h_.get(), 1);
    check(rc);
  }
  // predicted_errors_buffer of every shot of the last decode_batch_b8(collect_errors = true), as CSR.
  void last_batch_errors(size_t shots, std::vector<uint64_t>& offsets, std::vector<uint64_t>& idx) const {
    const uint64_t nnz = tsr_last_batch_nnz(h_.get());
    offsets.assign(shots + 1, 0);
    idx.assign((size_t)nnz + 1, 0);
    check(tsr_last_batch_errors(h_.get(), offsets.data(), idx.data()));
    idx.resize((size_t)nnz);
  }

 private:
  std::unique_ptr<tsr_decoder, HandleDeleter> h_;

  void run(const std::vector<uint64_t>& detections, int64_t order, uint64_t beam) {
    uint64_t n = 0;
    double cost = 0;
    int32_t low = 0;
    std::vector<uint64_t> buf(num_errors + 1);
    check(tsr_decode(h_.get(), detections.data(), detections.size(), order, beam, buf.data(), buf.size(), &n, &cost, &low));
    predicted_errors_buffer.assign(buf.begin(), buf.begin() + (std::ptrdiff_t)n);
    low_confidence_flag = low != 0;
  }
  std::vector<std::vector<int>> csr_rows(int which, size_t rows) const {
    std::vector<uint64_t> off(rows + 1);
    const int64_t nnz = tsr_csr(h_.get(), which, nullptr, nullptr);
    if (nnz < 0) throw_last((int)nnz);
    std::vector<int32_t> idx((size_t)std::max<int64_t>(nnz, 1));
    tsr_csr(h_.get(), which, off.data(), idx.data());
    std::vector<std::vector<int>> out(rows);
    for (size_t r = 0; r < rows; ++r) out[r].assign(idx.begin() + (std::ptrdiff_t)off[r], idx.begin() + (std::ptrdiff_t)off[r + 1]);
    return out;
  }
};

// Several GPUs driven by one process (tsr_group): what the reference does with one decoder per host thread
// (tesseract_main.cc:559-616), with the counters summed by one NCCL all-reduce.
struct DecoderGroup {
  explicit DecoderGroup(const TesseractConfig& config, int num_gpus) {
    tsr_config c;
    std::vector<uint64_t> flat;
    fill_config(config, c, flat);
    tsr_group* raw = nullptr;
    check(tsr_group_create(config.dem.c_str(), &c, nullptr, (uint32_t)num_gpus, &raw));
    g_.reset(raw);
  }
  tsr_decoder* member(uint32_t i) const { return tsr_group_member(g_.get(), i); }
  // Interleaved chunks over the GPUs; tally = {shots, logical errors, low confidence} after the all-reduce.
  void decode_batch_b8(const uint8_t* dets, size_t stride, size_t shots, const uint8_t* obs_true, uint8_t* obs_out, double* cost_out,
                       uint8_t* low_conf_out, uint64_t tally[3], uint64_t* error_use) {
    check(tsr_group_decode_b8(g_.get(), dets, stride, shots, 0, obs_true, obs_out, cost_out, low_conf_out, tally, error_use));
  }
  void last_batch_errors(size_t shots, std::vector<uint64_t>& offsets, std::vector<uint64_t>& idx) const {
    const uint64_t nnz = tsr_group_last_batch_nnz(g_.get());
    offsets.assign(shots + 1, 0);
    idx.assign((size_t)nnz + 1, 0);
    check(tsr_group_last_batch_errors(g_.get(), offsets.data(), idx.data()));
    idx.resize((size_t)nnz);
  }

 private:
  struct Deleter {
    void operator()(tsr_group* g) const { tsr_group_destroy(g); }
  };
  std::unique_ptr<tsr_group, Deleter> g_;
};

// utils.cc:192-205
enum class DetOrder { DetBFS = 0, DetIndex = 1, DetCoordinate = 2 };
inline std::vector<std::vector<size_t>> build_det_orders(const std::string& dem, size_t num_det_orders, DetOrder method,
                                                         uint64_t seed) {
  const uint64_t D = tsr_dem_count_detectors(dem.c_str());
  if (D == UINT64_MAX) throw std::invalid_argument(tsr_last_error());
  std::vector<uint64_t> flat(num_det_orders * D + 1);
  check(tsr_build_det_orders(dem.c_str(), num_det_orders, (int)method, seed, flat.data()));
  std::vector<std::vector<size_t>> out(num_det_orders);
  for (size_t k = 0; k < num_det_orders; ++k) out[k].assign(flat.begin() + (std::ptrdiff_t)(k * D), flat.begin() + (std::ptrdiff_t)((k + 1) * D));
  return out;
}

// The ingestion stages on their own (common.cc:139-218, utils.cc:32-87, 253-261); DEMs as text, like everywhere at this boundary.
inline std::string dem_stage(int (*fn)(const char*, char**, uint64_t*), const std::string& dem, std::vector<size_t>& error_index_map) {
  const int64_t n = tsr_dem_count_errors(dem.c_str());
  if (n < 0) throw_last((int)n);
  std::vector<uint64_t> map((size_t)n + 1);
  char* out = nullptr;
  check(fn(dem.c_str(), &out, map.data()));
  std::string text(out);
  tsr_free(out);
  error_index_map.assign(map.begin(), map.begin() + (std::ptrdiff_t)n);  // UINT64_MAX == std::numeric_limits<size_t>::max()
  return text;
}
inline std::string merge_indistinguishable_errors(const std::string& dem, std::vector<size_t>& error_index_map) {
  return dem_stage(tsr_merge_indistinguishable_errors, dem, error_index_map);
}
inline std::string remove_zero_probability_errors(const std::string& dem, std::vector<size_t>& error_index_map) {
  return dem_stage(tsr_remove_zero_probability_errors, dem, error_index_map);
}
inline std::vector<Error> get_errors_from_dem(const std::string& dem) {
  uint64_t nd = 0, no = 0;
  const int64_t n = tsr_get_errors_from_dem(dem.c_str(), nullptr, nullptr, nullptr, nullptr, nullptr, &nd, &no);
  if (n < 0) throw_last((int)n);
  std::vector<double> cost((size_t)n + 1);
  std::vector<uint64_t> doff((size_t)n + 1), ooff((size_t)n + 1);
  std::vector<int32_t> dets(nd + 1), obs(no + 1);
  tsr_get_errors_from_dem(dem.c_str(), cost.data(), doff.data(), dets.data(), ooff.data(), obs.data(), nullptr, nullptr);
  std::vector<Error> errors((size_t)n);
  for (size_t e = 0; e < (size_t)n; ++e) {
    errors[e].likelihood_cost = cost[e];
    errors[e].symptom.detectors.assign(dets.begin() + (std::ptrdiff_t)doff[e], dets.begin() + (std::ptrdiff_t)doff[e + 1]);
    errors[e].symptom.observables.assign(obs.begin() + (std::ptrdiff_t)ooff[e], obs.begin() + (std::ptrdiff_t)ooff[e + 1]);
  }
  return errors;
}
inline std::vector<std::vector<size_t>> build_detector_graph(const std::string& dem) {
  const uint64_t D = tsr_dem_count_detectors(dem.c_str());
  if (D == UINT64_MAX) throw std::invalid_argument(tsr_last_error());
  const int64_t n = tsr_build_detector_graph(dem.c_str(), nullptr, nullptr);
  if (n < 0) throw_last((int)n);
  std::vector<uint64_t> off(D + 1), idx((size_t)n + 1);
  tsr_build_detector_graph(dem.c_str(), off.data(), idx.data());
  std::vector<std::vector<size_t>> graph(D);
  for (uint64_t d = 0; d < D; ++d) graph[d].assign(idx.begin() + (std::ptrdiff_t)off[d], idx.begin() + (std::ptrdiff_t)off[d + 1]);
  return graph;
}
inline std::vector<std::vector<double>> get_detector_coords(const std::string& dem) {
  uint64_t rows = 0;
  const int64_t n = tsr_get_detector_coords(dem.c_str(), &rows, nullptr, nullptr);
  if (n < 0) throw_last((int)n);
  std::vector<uint64_t> off(rows + 1);
  std::vector<double> xs((size_t)n + 1);
  tsr_get_detector_coords(dem.c_str(), nullptr, off.data(), xs.data());
  std::vector<std::vector<double>> coords(rows);
  for (uint64_t r = 0; r < rows; ++r) coords[r].assign(xs.begin() + (std::ptrdiff_t)off[r], xs.begin() + (std::ptrdiff_t)off[r + 1]);
  return coords;
}

}  // namespace tesseract_b200
