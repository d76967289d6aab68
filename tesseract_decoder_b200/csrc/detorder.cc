// Detector orders (utils.cc:32-205): DetBFS, DetIndex, DetCoordinate.
//
// This file is compiled with the floating-point flavour of the reference build the parity tests pin against
// (oracle/Makefile: -march=x86-64-v3, GCC's default -ffp-contract=fast — the reference itself builds with
// -march=native, src/BUILD:43-55, i.e. with FMA on any current x86 host), NOT with the -ffp-contract=off of the rest of
// the host code: DetCoordinate sorts detectors by  proj += coord * dir  and by libstdc++'s normal_distribution
// (x*x + y*y), and whether those are fused changes the last bit of the projections and therefore the order of
// near-tied detectors (tests/test_host.py::test_det_coordinate_orders_with_near_tied_projections).
#include <algorithm>
#include <numeric>
#include <queue>
#include <random>
#include <stdexcept>

#include "ingest.h"

namespace tsr {

std::vector<std::vector<double>> get_detector_coords(const DemModel& dem) {
  // One entry per DETECTOR instruction, in order of appearance (utils.cc:42-48).
  std::vector<std::vector<double>> coords;
  for (const DemInstruction& inst : dem.instructions)
    if (inst.kind == DemKind::Detector) coords.push_back(inst.args);
  return coords;
}

std::vector<std::vector<uint64_t>> build_detector_graph(const DemModel& dem) {
  std::vector<std::vector<uint64_t>> nb(dem.count_detectors());
  for (const DemInstruction& inst : dem.instructions) {
    if (inst.kind != DemKind::Error) continue;
    std::vector<uint64_t> dets;
    for (const DemTarget& t : inst.targets)
      if (t.is_det()) dets.push_back(t.id());
    for (size_t i = 0; i < dets.size(); ++i)
      for (size_t j = i + 1; j < dets.size(); ++j) {
        nb[dets[i]].push_back(dets[j]);
        nb[dets[j]].push_back(dets[i]);
      }
  }
  for (auto& v : nb) {
    std::sort(v.begin(), v.end());
    v.erase(std::unique(v.begin(), v.end()), v.end());
  }
  return nb;
}

std::vector<std::vector<uint64_t>> build_det_orders(const DemModel& dem, uint64_t num_det_orders, DetOrder method,
                                                    uint64_t seed) {
  std::mt19937_64 rng(seed);
  const size_t n = dem.count_detectors();
  std::vector<std::vector<uint64_t>> orders(num_det_orders);
  switch (method) {
    case DetOrder::DetIndex: {
      // utils.cc:173-190: one coin per order, heads = descending.
      std::uniform_int_distribution<int> coin(0, 1);
      for (auto& order : orders) {
        order.resize(n);
        bool descending = coin(rng) != 0;
        for (size_t i = 0; i < n; ++i) order[i] = descending ? n - 1 - i : i;
      }
      return orders;
    }
    case DetOrder::DetBFS: {
      // utils.cc:89-133: randomised breadth-first sweeps; the stored list is the INVERSE of the
      // visiting sequence (a quirk the search then reads as position -> detector).
      auto graph = build_detector_graph(dem);
      std::uniform_int_distribution<size_t> pick(0, graph.size() - 1);
      for (auto& order : orders) {
        std::vector<size_t> seq;
        seq.reserve(graph.size());
        std::vector<char> seen(graph.size(), 0);
        std::queue<size_t> frontier;
        size_t start = pick(rng);
        while (seq.size() < graph.size()) {
          if (!seen[start]) {
            seen[start] = 1;
            frontier.push(start);
            seq.push_back(start);
          }
          while (!frontier.empty()) {
            size_t cur = frontier.front();
            frontier.pop();
            std::vector<uint64_t> nb = graph[cur];
            std::shuffle(nb.begin(), nb.end(), rng);
            for (uint64_t v : nb)
              if (!seen[v]) {
                seen[v] = 1;
                frontier.push(v);
                seq.push_back(v);
              }
          }
          if (seq.size() < graph.size()) {
            do {
              start = pick(rng);
            } while (seen[start]);
          }
        }
        order.resize(graph.size());
        for (size_t i = 0; i < seq.size(); ++i) order[seq[i]] = i;
      }
      return orders;
    }
    case DetOrder::DetCoordinate: {
      // utils.cc:135-171: sort detectors by their projection on a random Gaussian direction.
      auto coords = get_detector_coords(dem);
      std::normal_distribution<double> gauss(0, 1);
      if (coords.empty() || coords.at(0).empty()) {
        for (auto& order : orders) {
          order.resize(n);
          std::iota(order.begin(), order.end(), uint64_t{0});
        }
        return orders;
      }
      std::vector<double> proj(n);
      for (auto& order : orders) {
        std::vector<double> dir;
        for (size_t i = 0; i < coords.at(0).size(); ++i) dir.push_back(gauss(rng));
        for (size_t i = 0; i < coords.size(); ++i) {
          proj[i] = 0;
          for (size_t j = 0; j < dir.size(); ++j) proj[i] += coords[i][j] * dir[j];
        }
        std::vector<size_t> seq(n);
        std::iota(seq.begin(), seq.end(), size_t{0});
        std::sort(seq.begin(), seq.end(), [&](const size_t& a, const size_t& b) { return proj[a] > proj[b]; });
        order.resize(n);
        for (size_t i = 0; i < seq.size(); ++i) order[seq[i]] = i;
      }
      return orders;
    }
  }
  throw std::invalid_argument("Unknown det order method");
}

}  // namespace tsr
