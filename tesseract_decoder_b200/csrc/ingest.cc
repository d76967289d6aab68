#include "ingest.h"

#include <algorithm>
#include <cmath>
#include <iostream>
#include <map>
#include <numeric>
#include <queue>
#include <random>
#include <set>
#include <stdexcept>

namespace tsr {

double CompiledError::probability() const {
  return 1.0 / (1.0 + std::exp(cost));
}

double merge_weights(double a, double b) {
  // common.cc:118-123: the log-likelihood weight of "exactly one of two independent mechanisms".
  double sign = std::copysign(1, a) * std::copysign(1, b);
  double smaller = sign * std::min(std::abs(a), std::abs(b));
  return smaller + std::log(1 + std::exp(-std::abs(a + b))) - std::log(1 + std::exp(-std::abs(a - b)));
}

namespace {

// common.cc:52-88.  Targets are XOR-accumulated (a detector listed twice cancels); separators are
// ignored; the surviving ids come out ascending.
CompiledError compile_error(const DemInstruction& inst) {
  if (inst.kind != DemKind::Error)
    throw std::invalid_argument("Error must be loaded from an error dem instruction, but received: " + inst.str());
  double p = inst.args.at(0);
  if (p < 0 || p > 1)
    throw std::invalid_argument("Probability must be between 0 and 1, but received: " + std::to_string(p));
  std::set<int> dets, obs;
  for (const DemTarget& t : inst.targets) {
    if (t.is_sep()) continue;
    std::set<int>& into = t.is_obs() ? obs : dets;
    int id = (int)t.id();
    if (!into.erase(id)) into.insert(id);
  }
  CompiledError e;
  e.cost = -1 * std::log(p / (1 - p));
  e.detectors.assign(dets.begin(), dets.end());
  e.observables.assign(obs.begin(), obs.end());
  return e;
}

std::vector<DemTarget> symptom_targets(const CompiledError& e) {
  std::vector<DemTarget> t;
  for (int d : e.detectors) t.push_back(DemTarget::det((uint64_t)d));
  for (int o : e.observables) t.push_back(DemTarget::obs((uint64_t)o));
  return t;
}

}  // namespace

DemModel merge_indistinguishable_errors(const DemModel& flat, std::vector<uint64_t>& error_index_map) {
  // First-seen order defines the merged index.  The reference keys an unordered_map by symptom (common.cc:146); only
  // find / insert are used, so any exact map gives the same indices.
  DemModel out;
  error_index_map.clear();
  std::map<std::pair<std::vector<int>, std::vector<int>>, uint64_t> index_of;
  std::vector<CompiledError> merged;
  for (const DemInstruction& inst : flat.instructions) {
    if (inst.kind != DemKind::Error) {
      out.instructions.push_back(inst);
      continue;
    }
    CompiledError e = compile_error(inst);
    if (e.detectors.empty()) std::cout << "Warning: the circuit has errors that do not flip any detectors \n";
    auto key = std::make_pair(e.detectors, e.observables);
    auto it = index_of.find(key);
    if (it == index_of.end()) {
      index_of.emplace(std::move(key), merged.size());
      error_index_map.push_back(merged.size());
      merged.push_back(std::move(e));
    } else {
      merged[it->second].cost = merge_weights(e.cost, merged[it->second].cost);
      error_index_map.push_back(it->second);
    }
  }
  for (const CompiledError& e : merged) out.append_error(e.probability(), symptom_targets(e));
  return out;
}

DemModel remove_zero_probability_errors(const DemModel& flat, std::vector<uint64_t>& error_index_map) {
  DemModel out;
  error_index_map.clear();
  uint64_t kept = 0;
  for (const DemInstruction& inst : flat.instructions) {
    if (inst.kind == DemKind::Error) {
      if (inst.args[0] > 0) {
        error_index_map.push_back(kept++);
        out.instructions.push_back(inst);
      } else {
        error_index_map.push_back(kNoError);
      }
    } else {
      out.instructions.push_back(inst);
    }
  }
  return out;
}

std::vector<CompiledError> get_errors_from_dem(const DemModel& flat) {
  std::vector<CompiledError> errors;
  for (const DemInstruction& inst : flat.instructions)
    if (inst.kind == DemKind::Error && inst.args[0] > 0) errors.push_back(compile_error(inst));
  return errors;
}

DemTables ingest(const DemModel& flat, bool merge_errors) {
  DemTables T;
  T.num_dem_errors = flat.count_errors();
  std::vector<uint64_t> to_compiled(T.num_dem_errors);
  std::iota(to_compiled.begin(), to_compiled.end(), uint64_t{0});

  // Stage 1 (common.cc:139-190, tesseract.cc:159-164): merge errors with identical symptoms.
  DemModel stage1;
  if (merge_errors) {
    std::vector<uint64_t> merge_map;
    stage1 = merge_indistinguishable_errors(flat, merge_map);
    for (uint64_t& v : to_compiled) v = merge_map[v];
  } else {
    stage1 = flat;
  }

  // Stage 2 (common.cc:192-218, tesseract.cc:165-168): drop errors whose probability is not > 0.
  std::vector<uint64_t> keep_map;
  T.compiled_dem = remove_zero_probability_errors(stage1, keep_map);
  for (uint64_t& v : to_compiled) v = keep_map[v];  // common::chain_error_maps; nothing is kNoError before this stage
  T.dem_error_to_error = to_compiled;

  // Stage 3 (utils.cc:253-261): the compiled error list; costs are recomputed from the printed
  // probabilities, i.e. after the cost -> p -> cost round trip.
  T.errors = get_errors_from_dem(T.compiled_dem);
  T.num_errors = T.errors.size();
  T.num_detectors = T.compiled_dem.count_detectors();
  T.num_observables = T.compiled_dem.count_observables();

  // common.cc:228-239: first original index for each compiled error.
  T.error_to_dem_error.assign(T.num_errors, kNoError);
  for (uint64_t i = 0; i < T.dem_error_to_error.size(); ++i) {
    uint64_t c = T.dem_error_to_error[i];
    if (c != kNoError && T.error_to_dem_error[c] == kNoError) T.error_to_dem_error[c] = i;
  }

  const uint64_t D = T.num_detectors, E = T.num_errors;
  if (D >= 0xFFFF) throw std::invalid_argument("tesseract_b200: at most 65534 detectors are supported");

  // Rows (tesseract.cc:211-227).  Filled in ascending error index, then std::sort by cost/degree —
  // an unstable sort whose tie order is part of the contract, so it is the same
  // libstdc++ call on the same input sequence.
  std::vector<std::vector<int>> rows(D);
  std::vector<double> per_detector_cost(E);
  for (uint64_t e = 0; e < E; ++e) {
    for (int d : T.errors[e].detectors) rows[(size_t)d].push_back((int)e);
    per_detector_cost[e] = T.errors[e].cost / T.errors[e].detectors.size();
  }
  for (auto& row : rows)
    std::sort(row.begin(), row.end(), [&](size_t a, size_t b) { return per_detector_cost[a] < per_detector_cost[b]; });

  T.row_off.assign(D + 1, 0);
  for (uint64_t d = 0; d < D; ++d) {
    T.row_off[d + 1] = T.row_off[d] + (uint32_t)rows[d].size();
    T.max_row_len = std::max<uint32_t>(T.max_row_len, (uint32_t)rows[d].size());
  }
  if (T.max_row_len > 0x7FFE) throw std::invalid_argument("tesseract_b200: detector rows longer than 32766 are not supported");
  const size_t nnz = T.row_off[D];
  T.row_err.resize(nnz);
  for (uint64_t d = 0; d < D; ++d) std::copy(rows[d].begin(), rows[d].end(), T.row_err.begin() + T.row_off[d]);

  T.e_off.assign(E + 1, 0);
  T.eo_off.assign(E + 1, 0);
  T.e_cost.resize(E);
  for (uint64_t e = 0; e < E; ++e) {
    T.e_off[e + 1] = T.e_off[e] + (uint32_t)T.errors[e].detectors.size();
    T.eo_off[e + 1] = T.eo_off[e] + (uint32_t)T.errors[e].observables.size();
    T.e_cost[e] = T.errors[e].cost;
    T.max_degree = std::max<uint32_t>(T.max_degree, (uint32_t)T.errors[e].detectors.size());
  }
  T.e_det.resize(nnz);
  T.e_rank.assign(nnz, -1);
  T.eo_obs.resize(T.eo_off[E]);
  for (uint64_t e = 0; e < E; ++e) {
    std::copy(T.errors[e].detectors.begin(), T.errors[e].detectors.end(), T.e_det.begin() + T.e_off[e]);
    std::copy(T.errors[e].observables.begin(), T.errors[e].observables.end(), T.eo_obs.begin() + T.eo_off[e]);
  }
  for (uint64_t d = 0; d < D; ++d) {
    for (uint32_t k = 0; k < rows[d].size(); ++k) {
      uint64_t e = (uint64_t)rows[d][k];
      const auto& dets = T.errors[e].detectors;
      size_t j = (size_t)(std::lower_bound(dets.begin(), dets.end(), (int)d) - dets.begin());
      T.e_rank[T.e_off[e] + j] = (int32_t)k;
    }
  }

  // Row records.
  const RowSlot unused{(uint16_t)D, 0xFFFF};
  T.rec_a.resize(nnz);
  T.rec_b.assign(nnz * 4, unused);
  T.rec_c.assign(nnz * 4, unused);
  for (uint64_t d = 0; d < D; ++d) {
    for (uint32_t k = 0; k < rows[d].size(); ++k) {
      size_t at = T.row_off[d] + k;
      uint64_t e = (uint64_t)rows[d][k];
      const auto& dets = T.errors[e].detectors;
      RowRecordA a;
      a.cost = T.errors[e].cost;
      a.size = (uint16_t)std::min<size_t>(dets.size(), 0xFFFF);
      a.flags = dets.size() > (size_t)kMaxInlineDegree ? 1 : 0;
      a.other0 = unused;
      int slot = 0;
      for (size_t j = 0; j < dets.size() && slot < kMaxInlineDegree - 1; ++j) {
        if ((uint64_t)dets[j] == d) continue;
        RowSlot s{(uint16_t)dets[j], (uint16_t)T.e_rank[T.e_off[e] + j]};
        if (slot == 0) {
          a.other0 = s;
        } else if (slot <= 4) {
          T.rec_b[at * 4 + (size_t)(slot - 1)] = s;
        } else {
          T.rec_c[at * 4 + (size_t)(slot - 5)] = s;
        }
        ++slot;
      }
      T.rec_a[at] = a;
    }
  }

  // Detector adjacency (replaces eneighbors, tesseract.cc:229-254).
  T.adj_words = (uint32_t)((D + 31) / 32);
  T.adj.assign((size_t)D * T.adj_words, 0);
  for (const CompiledError& e : T.errors)
    for (int a : e.detectors)
      for (int b : e.detectors) T.adj[(size_t)a * T.adj_words + ((size_t)b >> 5)] |= 1u << (b & 31);
  return T;
}

std::vector<std::vector<int>> DemTables::eneighbors() const {
  std::vector<std::vector<int>> out(num_errors);
  std::vector<uint32_t> acc(adj_words);
  for (uint64_t e = 0; e < num_errors; ++e) {
    std::fill(acc.begin(), acc.end(), 0u);
    for (int d : errors[e].detectors)
      for (uint32_t w = 0; w < adj_words; ++w) acc[w] |= adj[(size_t)d * adj_words + w];
    for (int d : errors[e].detectors) acc[(size_t)d >> 5] &= ~(1u << (d & 31));
    for (uint32_t w = 0; w < adj_words; ++w)
      for (uint32_t m = acc[w]; m; m &= m - 1) out[e].push_back((int)(w * 32 + (uint32_t)__builtin_ctz(m)));
  }
  return out;
}

}  // namespace tsr
