// TEST INFRASTRUCTURE — NOT PRODUCT CODE.
//
// extern "C" face (oracle_api.h) over the reference's OWN classes.  This file is compiled together
// with the unmodified /root/reference/src/{tesseract,common,utils,visualization}.cc (see Makefile)
// into oracle/_ref/libtesseract_ref.so.  It adds no decoding logic: every result comes from
// tesseract_decoder::TesseractDecoder (tesseract.h:82-147).
#include <atomic>
#include <chrono>
#include <memory>
#include <thread>

#include "common.h"
#include "oracle_api.h"
#include "tesseract.h"
#include "utils.h"

using tesseract_decoder::TesseractConfig;
using tesseract_decoder::TesseractDecoder;

struct orc_handle {
  std::vector<std::unique_ptr<TesseractDecoder>> pool;
  std::vector<std::vector<size_t>> last_errors;
};

static thread_local std::string g_err;

template <typename F>
static int guarded(F&& f) {
  try {
    f();
    return 0;
  } catch (const std::exception& e) {
    g_err = e.what();
    return -1;
  }
}

extern "C" {

const char* orc_last_error(void) {
  return g_err.c_str();
}
const char* orc_kind(void) {
  return "reference";
}

orc_handle* orc_create(const char* dem_text, const orc_config* cfg) {
  orc_handle* h = nullptr;
  int rc = guarded([&] {
    if (cfg->at_most_two_errors_per_detector)
      throw std::invalid_argument("the reference has no at_most_two_errors_per_detector (README.md:59 only); use the port checker");
    stim::DetectorErrorModel dem(dem_text);
    TesseractConfig c;
    c.dem = dem;
    c.det_beam = cfg->det_beam;
    c.beam_climbing = cfg->beam_climbing != 0;
    c.no_revisit_dets = cfg->no_revisit_dets != 0;
    c.merge_errors = cfg->merge_errors != 0;
    c.pqlimit = cfg->pqlimit == UINT64_MAX ? std::numeric_limits<size_t>::max() : (size_t)cfg->pqlimit;
    c.det_penalty = cfg->det_penalty;
    c.sparsify_errors = cfg->sparsify_errors != 0;
    c.sparsify_base_degree = cfg->sparsify_base_degree;
    c.sparsify_max_degree = cfg->sparsify_max_degree;
    c.sparsify_reactivate_limit = cfg->sparsify_reactivate_limit;
    c.create_visualization = cfg->create_visualization != 0;
    uint64_t D = tesseract_decoder::common::flatten(dem).count_detectors();
    for (uint64_t k = 0; k < cfg->num_det_orders; ++k)
      c.det_orders.emplace_back(cfg->det_orders + k * D, cfg->det_orders + (k + 1) * D);
    auto out = std::make_unique<orc_handle>();
    int n = cfg->num_threads < 1 ? 1 : cfg->num_threads;
    for (int t = 0; t < n; ++t) out->pool.push_back(std::make_unique<TesseractDecoder>(c));
    h = out.release();
  });
  return rc == 0 ? h : nullptr;
}

void orc_destroy(orc_handle* h) {
  delete h;
}

uint64_t orc_num_detectors(const orc_handle* h) {
  return h->pool[0]->num_detectors;
}
uint64_t orc_num_observables(const orc_handle* h) {
  return h->pool[0]->num_observables;
}
uint64_t orc_num_errors(const orc_handle* h) {
  return h->pool[0]->num_errors;
}
uint64_t orc_num_dem_errors(const orc_handle* h) {
  return h->pool[0]->dem_error_to_error.size();
}

int orc_error_costs(const orc_handle* h, double* out) {
  const auto& d = *h->pool[0];
  for (size_t i = 0; i < d.errors.size(); ++i) out[i] = d.errors[i].likelihood_cost;
  return 0;
}

int orc_maps(const orc_handle* h, uint64_t* dem_error_to_error, uint64_t* error_to_dem_error) {
  const auto& d = *h->pool[0];
  if (dem_error_to_error)
    for (size_t i = 0; i < d.dem_error_to_error.size(); ++i) dem_error_to_error[i] = d.dem_error_to_error[i];
  if (error_to_dem_error)
    for (size_t i = 0; i < d.error_to_dem_error.size(); ++i) error_to_dem_error[i] = d.error_to_dem_error[i];
  return 0;
}

int64_t orc_csr(const orc_handle* h, int which, uint64_t* offsets, int32_t* idx) {
  auto& d = *h->pool[0];
  std::vector<std::vector<int>> obs;
  const std::vector<std::vector<int>>* rows = nullptr;
  switch (which) {
    case 0:
      rows = &d.d2e;
      break;
    case 1:
      rows = &d.edets;
      break;
    case 2:
      rows = &d.get_eneighbors();
      break;
    case 3:
      for (const auto& e : d.errors) obs.push_back(e.symptom.observables);
      rows = &obs;
      break;
    default:
      g_err = "bad table id";
      return -1;
  }
  uint64_t n = 0;
  for (size_t r = 0; r < rows->size(); ++r) {
    if (offsets) offsets[r] = n;
    for (int v : (*rows)[r]) {
      if (idx) idx[n] = v;
      ++n;
    }
  }
  if (offsets) offsets[rows->size()] = n;
  return (int64_t)n;
}

int orc_decode(orc_handle* h, const uint64_t* hits, uint64_t n_hits, int64_t order, uint64_t beam,
               uint64_t* errs_out, uint64_t errs_cap, uint64_t* n_errs, double* cost,
               int32_t* low_confidence) {
  return guarded([&] {
    auto& d = *h->pool[0];
    std::vector<uint64_t> dets(hits, hits + n_hits);
    if (order < 0) {
      d.decode_to_errors(dets);
    } else {
      d.decode_to_errors(dets, (size_t)order, (size_t)beam);
    }
    const auto& buf = d.predicted_errors_buffer;
    if (n_errs) *n_errs = buf.size();
    if (errs_out) {
      if (buf.size() > errs_cap) throw std::runtime_error("errs_out too small");
      for (size_t i = 0; i < buf.size(); ++i) errs_out[i] = buf[i];
    }
    if (cost) *cost = d.cost_from_errors(buf);
    if (low_confidence) *low_confidence = d.low_confidence_flag ? 1 : 0;
  });
}

int orc_decode_batch_b8(orc_handle* h, const uint8_t* dets, uint64_t stride, uint64_t shots,
                        uint8_t* obs_out, double* cost_out, uint8_t* low_conf_out,
                        double* wall_seconds, double* sum_shot_seconds) {
  return guarded([&] {
    const size_t D = h->pool[0]->num_detectors;
    const size_t O = h->pool[0]->num_observables;
    const size_t obs_stride = (O + 7) / 8;
    h->last_errors.assign(shots, {});
    std::atomic<uint64_t> next{0};
    std::vector<double> thread_seconds(h->pool.size(), 0.0);
    std::vector<std::string> thread_err(h->pool.size());
    auto t0 = std::chrono::steady_clock::now();
    std::vector<std::thread> workers;
    for (size_t t = 0; t < h->pool.size(); ++t) {
      workers.emplace_back([&, t] {
        try {
          auto& d = *h->pool[t];
          std::vector<uint64_t> hits;
          for (uint64_t s; (s = next++) < shots;) {
            hits.clear();
            const uint8_t* row = dets + s * stride;
            for (size_t k = 0; k < D; ++k)
              if ((row[k >> 3] >> (k & 7)) & 1) hits.push_back(k);
            auto a = std::chrono::high_resolution_clock::now();
            d.decode_to_errors(hits);
            auto b = std::chrono::high_resolution_clock::now();
            thread_seconds[t] += std::chrono::duration<double>(b - a).count();
            h->last_errors[s] = d.predicted_errors_buffer;
            if (low_conf_out) low_conf_out[s] = d.low_confidence_flag ? 1 : 0;
            if (cost_out) cost_out[s] = d.cost_from_errors(d.predicted_errors_buffer);
            if (obs_out) {
              uint8_t* o = obs_out + s * obs_stride;
              for (size_t k = 0; k < obs_stride; ++k) o[k] = 0;
              for (int ob : d.get_flipped_observables(d.predicted_errors_buffer))
                o[ob >> 3] ^= (uint8_t)(1u << (ob & 7));
            }
          }
        } catch (const std::exception& e) {
          thread_err[t] = e.what();
          next = shots;
        }
      });
    }
    for (auto& w : workers) w.join();
    auto t1 = std::chrono::steady_clock::now();
    for (auto& e : thread_err)
      if (!e.empty()) throw std::runtime_error(e);
    if (wall_seconds) *wall_seconds = std::chrono::duration<double>(t1 - t0).count();
    if (sum_shot_seconds) {
      *sum_shot_seconds = 0;
      for (double v : thread_seconds) *sum_shot_seconds += v;
    }
  });
}

uint64_t orc_last_batch_nnz(const orc_handle* h) {
  uint64_t n = 0;
  for (const auto& v : h->last_errors) n += v.size();
  return n;
}

int orc_last_batch_errors(const orc_handle* h, uint64_t* offsets, uint64_t* idx) {
  uint64_t n = 0;
  for (size_t s = 0; s < h->last_errors.size(); ++s) {
    offsets[s] = n;
    for (size_t v : h->last_errors[s]) idx[n++] = v;
  }
  offsets[h->last_errors.size()] = n;
  return 0;
}

int orc_cost_from_errors(const orc_handle* h, const uint64_t* errs, uint64_t n, double* out) {
  return guarded([&] {
    std::vector<size_t> v(errs, errs + n);
    *out = h->pool[0]->cost_from_errors(v);
  });
}

int orc_flipped_observables(const orc_handle* h, const uint64_t* errs, uint64_t n, uint8_t* obs_bits_out) {
  return guarded([&] {
    std::vector<size_t> v(errs, errs + n);
    size_t O = h->pool[0]->num_observables;
    for (size_t k = 0; k < (O + 7) / 8; ++k) obs_bits_out[k] = 0;
    for (int ob : h->pool[0]->get_flipped_observables(v)) obs_bits_out[ob >> 3] ^= (uint8_t)(1u << (ob & 7));
  });
}

int orc_build_det_orders(const char* dem_text, uint64_t num_det_orders, int method, uint64_t seed,
                         uint64_t* out) {
  return guarded([&] {
    stim::DetectorErrorModel dem(dem_text);
    tesseract_decoder::DetOrder m = method == 0   ? tesseract_decoder::DetOrder::DetBFS
                                    : method == 1 ? tesseract_decoder::DetOrder::DetIndex
                                                  : tesseract_decoder::DetOrder::DetCoordinate;
    auto orders = tesseract_decoder::build_det_orders(dem, num_det_orders, m, seed);
    size_t D = dem.count_detectors();
    for (size_t k = 0; k < orders.size(); ++k)
      for (size_t i = 0; i < D; ++i) out[k * D + i] = orders[k][i];
  });
}

// ---- the ingestion stages on their own (reference build only; oracle_api.h) ------------------------------

int64_t orc_dem_dump(const char* dem_text, int which, uint64_t* index_map, uint8_t* kind, uint64_t* arg_off, double* args,
                     uint64_t* tgt_off, uint64_t* tgt, uint64_t sizes[4]) {
  int64_t n = 0;
  int rc = guarded([&] {
    namespace cm = tesseract_decoder::common;
    stim::DetectorErrorModel dem(dem_text);
    std::vector<size_t> map;
    stim::DetectorErrorModel out = which == 1   ? cm::merge_indistinguishable_errors(dem, map)
                                   : which == 2 ? cm::remove_zero_probability_errors(dem, map)
                                                : cm::flatten(dem);
    uint64_t na = 0, nt = 0, i = 0;
    for (const stim::DemInstruction& inst : out.instructions) {
      if (kind)
        kind[i] = inst.type == stim::DemInstructionType::DEM_ERROR      ? 0
                  : inst.type == stim::DemInstructionType::DEM_DETECTOR ? 1
                  : inst.type == stim::DemInstructionType::DEM_LOGICAL_OBSERVABLE ? 2 : 255;
      if (arg_off) arg_off[i] = na;
      if (tgt_off) tgt_off[i] = nt;
      for (double a : inst.arg_data) {
        if (args) args[na] = a;
        ++na;
      }
      for (const stim::DemTarget& t : inst.target_data) {
        if (tgt) tgt[nt] = t.is_separator() ? UINT64_MAX : (t.is_observable_id() ? (t.raw_id() | (uint64_t{1} << 63)) : t.raw_id());
        ++nt;
      }
      ++i;
    }
    if (arg_off) arg_off[i] = na;
    if (tgt_off) tgt_off[i] = nt;
    if (index_map)
      for (size_t k = 0; k < map.size(); ++k) index_map[k] = map[k] == std::numeric_limits<size_t>::max() ? UINT64_MAX : map[k];
    if (sizes) {
      sizes[0] = i;
      sizes[1] = na;
      sizes[2] = nt;
      sizes[3] = map.size();
    }
    n = (int64_t)i;
  });
  return rc == 0 ? n : -1;
}

int64_t orc_get_errors_from_dem(const char* dem_text, double* cost, uint64_t* det_offsets, int32_t* dets,
                                uint64_t* obs_offsets, int32_t* obs, uint64_t* n_dets, uint64_t* n_obs) {
  int64_t n = 0;
  int rc = guarded([&] {
    stim::DetectorErrorModel dem(dem_text);
    // the function reads dem.instructions as they are (utils.cc:255); every caller hands it a flattened DEM
    auto errors = tesseract_decoder::get_errors_from_dem(tesseract_decoder::common::flatten(dem));
    uint64_t nd = 0, no = 0;
    for (size_t e = 0; e < errors.size(); ++e) {
      if (cost) cost[e] = errors[e].likelihood_cost;
      if (det_offsets) det_offsets[e] = nd;
      if (obs_offsets) obs_offsets[e] = no;
      for (int d : errors[e].symptom.detectors) {
        if (dets) dets[nd] = d;
        ++nd;
      }
      for (int o : errors[e].symptom.observables) {
        if (obs) obs[no] = o;
        ++no;
      }
    }
    if (det_offsets) det_offsets[errors.size()] = nd;
    if (obs_offsets) obs_offsets[errors.size()] = no;
    if (n_dets) *n_dets = nd;
    if (n_obs) *n_obs = no;
    n = (int64_t)errors.size();
  });
  return rc == 0 ? n : -1;
}

int64_t orc_build_detector_graph(const char* dem_text, uint64_t* offsets, uint64_t* neighbors) {
  int64_t n = 0;
  int rc = guarded([&] {
    stim::DetectorErrorModel dem(dem_text);
    auto graph = tesseract_decoder::build_detector_graph(dem);
    uint64_t at = 0;
    for (size_t d = 0; d < graph.size(); ++d) {
      if (offsets) offsets[d] = at;
      for (size_t v : graph[d]) {
        if (neighbors) neighbors[at] = v;
        ++at;
      }
    }
    if (offsets) offsets[graph.size()] = at;
    n = (int64_t)at;
  });
  return rc == 0 ? n : -1;
}

int64_t orc_get_detector_coords(const char* dem_text, uint64_t* n_rows, uint64_t* offsets, double* coords) {
  int64_t n = 0;
  int rc = guarded([&] {
    stim::DetectorErrorModel dem(dem_text);
    auto rows = tesseract_decoder::get_detector_coords(dem);
    uint64_t at = 0;
    for (size_t r = 0; r < rows.size(); ++r) {
      if (offsets) offsets[r] = at;
      for (double v : rows[r]) {
        if (coords) coords[at] = v;
        ++at;
      }
    }
    if (offsets) offsets[rows.size()] = at;
    if (n_rows) *n_rows = rows.size();
    n = (int64_t)at;
  });
  return rc == 0 ? n : -1;
}

int orc_error_text(double likelihood_cost, const int32_t* dets, uint64_t n_dets, const int32_t* obs, uint64_t n_obs, char* out,
                   uint64_t cap, double* probability) {
  return guarded([&] {
    std::vector<int> d(dets, dets + n_dets), o(obs, obs + n_obs);
    tesseract_decoder::common::Error e(likelihood_cost, d, o);
    const std::string s = e.str();
    if (s.size() + 1 > cap) throw std::invalid_argument("buffer too small");
    std::copy(s.begin(), s.end(), out);
    out[s.size()] = 0;
    if (probability) *probability = e.get_probability();
  });
}

int orc_cost_of_probability(double p, double* cost) {
  return guarded([&] {
    tesseract_decoder::common::Error e;
    e.set_with_probability(p);
    *cost = e.likelihood_cost;
  });
}

double orc_merge_weights(double a, double b) {
  return tesseract_decoder::common::merge_weights(a, b);
}

int orc_visualization_write(orc_handle* h, const char* path) {
  return guarded([&] { h->pool[0]->visualizer.write(path); });
}

int orc_counters(orc_handle*, uint64_t out[8], int) {
  for (int i = 0; i < 8; ++i) out[i] = 0;
  return 0;
}

}  // extern "C"

#include "sampler.inc"
