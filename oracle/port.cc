// TEST INFRASTRUCTURE — NOT PRODUCT CODE.
//
// CPU restatement ("port") of the reference's batched-shot A* decode path, written from the
// algorithm rather than from the reference's text: flat CSR tables, an explicit binary heap that
// spells out libstdc++'s sift order, and the SURVEY.md §8(d) search counters.  It exists so that
//   (1) the reference build (oracle/_ref, the unmodified sources) can be cross-checked against an
//       independent reading of the same algorithm (tests/test_oracle.py: bit-for-bit agreement), and
//   (2) the counters Q/P/X/L/S/A/V that define "algorithmic bytes per shot" have a CPU source.
// Each function cites the reference lines it follows.  Only tests/, __graft_entry__.smoke() and
// bench.py's cpu_baseline / --impl reference legs may load this library; the product never does.
//
// Parity status: PINNED — checked against the reference's own known-answer tests and, shot for shot,
// against oracle/_ref on every committed workload (tests/test_oracle.py).
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <limits>
#include <iomanip>
#include <map>
#include <sstream>
#include <memory>
#include <numeric>
#include <queue>
#include <random>
#include <set>
#include <thread>
#include <unordered_set>

#include "oracle_api.h"
#include "stim.h"

namespace port {

constexpr uint64_t kGone = std::numeric_limits<uint64_t>::max();
const double kInf = std::numeric_limits<double>::infinity();

struct Mechanism {  // common.h:59-75
  double cost = 0;
  std::vector<int> dets, obs;
};

// common.cc:52-88 — XOR-canonicalised symptom, cost = -log(p/(1-p)).
static Mechanism mechanism_of(const stim::DemInstruction& inst) {
  double p = inst.arg_data[0];
  if (p < 0 || p > 1) throw std::invalid_argument("Probability must be between 0 and 1, but received: " + std::to_string(p));
  std::set<int> d, o;
  for (const auto& t : inst.target_data) {
    if (t.is_observable_id()) {
      if (!o.erase((int)t.val())) o.insert((int)t.val());
    } else if (t.is_relative_detector_id()) {
      if (!d.erase((int)t.val())) d.insert((int)t.val());
    }
  }
  Mechanism m;
  m.cost = -1 * std::log(p / (1 - p));
  m.dets.assign(d.begin(), d.end());
  m.obs.assign(o.begin(), o.end());
  return m;
}

// common.cc:118-123
static double merge_weights(double a, double b) {
  double s = std::copysign(1, a) * std::copysign(1, b);
  double m = s * std::min(std::abs(a), std::abs(b));
  return m + std::log(1 + std::exp(-std::abs(a + b))) - std::log(1 + std::exp(-std::abs(a - b)));
}

struct Config {
  int det_beam = 5;
  bool beam_climbing = false, no_revisit = true, merge = true;
  uint64_t pqlimit = 200000;
  double det_penalty = 0;
  std::vector<std::vector<uint64_t>> orders;
  // tesseract.h:52-55
  bool sparsify = false;
  int sparsify_base_degree = -1, sparsify_max_degree = -1, sparsify_reactivate_limit = -1;
  // --at-most-two-errors-per-detector.  The reference snapshot names the flag (README.md:59,159) but holds no code for
  // it, so there is nothing to follow line by line: PARITY UNPINNED against the reference.  The meaning restated here
  // (and implemented independently by the CUDA kernels): once a chosen error switches a fired detector OFF, no
  // further error may touch that detector — every error of its row is blocked for the node's descendants.  A
  // detector that starts fired is then touched exactly once and one that starts quiet zero or two times (on, off),
  // so the search returns the cheapest explanation among those that use at most two errors per detector
  // (tests/test_oracle.py checks that against brute force on the reference's four literal DEMs).
  bool at_most_two = false;
  bool create_visualization = false;  // tesseract.h:50
};

struct Counters {
  uint64_t Q = 0, P = 0, X = 0, L = 0, S = 0, A = 0, V = 0, searches = 0;
};

struct QueueNode {  // tesseract.h:60-70
  double f;
  uint64_t num_dets, depth;
  int64_t chain;
};
// tesseract.cc:119-121
static bool worse(const QueueNode& a, const QueueNode& b) {
  return a.f > b.f || (a.f == b.f && a.num_dets < b.num_dets);
}

// std::priority_queue<Node, vector<Node>, greater<Node>> spelled out (bits/stl_heap.h:135-148, 224-267).
struct MinQueue {
  std::vector<QueueNode> a;
  bool empty() const { return a.empty(); }
  void sift_up(int64_t hole, int64_t top, QueueNode v) {
    int64_t parent = (hole - 1) / 2;
    while (hole > top && worse(a[(size_t)parent], v)) {
      a[(size_t)hole] = a[(size_t)parent];
      hole = parent;
      parent = (hole - 1) / 2;
    }
    a[(size_t)hole] = v;
  }
  void push(const QueueNode& v) {
    a.push_back(v);
    sift_up((int64_t)a.size() - 1, 0, v);
  }
  QueueNode pop() {
    QueueNode top = a.front();
    if (a.size() > 1) {
      QueueNode v = a.back();
      a.back() = a.front();
      int64_t len = (int64_t)a.size() - 1, hole = 0, child = 0;
      while (child < (len - 1) / 2) {
        child = 2 * (child + 1);
        if (worse(a[(size_t)child], a[(size_t)child - 1])) --child;
        a[(size_t)hole] = a[(size_t)child];
        hole = child;
      }
      if ((len & 1) == 0 && child == (len - 2) / 2) {
        child = 2 * (child + 1);
        a[(size_t)hole] = a[(size_t)child - 1];
        hole = child - 1;
      }
      sift_up(hole, 0, v);
    }
    a.pop_back();
    return top;
  }
};

struct ChainLink {  // common.h:52-56
  uint32_t error, min_det;
  int64_t parent;
  uint32_t switched_off;  // at-most-two: bit j = detector j of `error` was fired when the error was chosen
};

struct Decoder {
  Config cfg;
  uint64_t D = 0, O = 0, E = 0;
  std::vector<Mechanism> errs;
  std::vector<uint64_t> dem2err, err2dem;
  // CSR tables
  std::vector<uint32_t> row_off, e_off, nb_off;
  std::vector<int> row, edet, nb;
  std::vector<double> cost, per_det;  // tesseract.cc:218-221
  // results
  std::vector<uint64_t> predicted;
  bool low_confidence = false;
  Counters ctr;
  // scratch
  std::vector<ChainLink> chain;
  std::vector<uint32_t> blocked, count, child_blocked, child_count;
  std::vector<int> amt_undo;  // at-most-two: errors blocked for the current child only
  std::vector<std::string> viz;  // visualization.cc: the Visualizer's lines
  // Rows the search walks: the full detector rows, or the shot's sparse rows (tesseract.cc:296-298).
  std::vector<uint32_t> aoff;
  std::vector<int> arow;
  std::vector<int> mandatory, optional;  // tesseract.cc:279-289
  std::vector<uint8_t> active;

  Decoder(const stim::DetectorErrorModel& input, Config c) : cfg(std::move(c)) {
    // tesseract.cc:156-205
    stim::DetectorErrorModel flat = input.flattened();
    dem2err.resize(flat.count_errors());
    std::iota(dem2err.begin(), dem2err.end(), uint64_t{0});
    stim::DetectorErrorModel cur = flat;
    if (cfg.merge) {  // common.cc:139-190
      std::map<std::pair<std::vector<int>, std::vector<int>>, size_t> first;
      std::vector<Mechanism> merged;
      std::vector<uint64_t> map;
      stim::DetectorErrorModel out;
      for (const auto& inst : flat.instructions) {
        if (inst.type != stim::DemInstructionType::DEM_ERROR) {
          out.append_dem_instruction(inst);
          continue;
        }
        Mechanism m = mechanism_of(inst);
        auto key = std::make_pair(m.dets, m.obs);
        auto it = first.find(key);
        if (it == first.end()) {
          first[key] = merged.size();
          map.push_back(merged.size());
          merged.push_back(m);
        } else {
          merged[it->second].cost = merge_weights(m.cost, merged[it->second].cost);
          map.push_back(it->second);
        }
      }
      for (const auto& m : merged) {
        std::vector<stim::DemTarget> t;
        for (int d : m.dets) t.push_back(stim::DemTarget::relative_detector_id((uint64_t)d));
        for (int o : m.obs) t.push_back(stim::DemTarget::observable_id((uint64_t)o));
        out.append_error_instruction(1.0 / (1.0 + std::exp(m.cost)), t, "");
      }
      for (auto& v : dem2err) v = map[v];
      cur = out;
    }
    {  // common.cc:192-218
      stim::DetectorErrorModel out;
      std::vector<uint64_t> map;
      uint64_t kept = 0;
      for (const auto& inst : cur.instructions) {
        if (inst.type == stim::DemInstructionType::DEM_ERROR) {
          if (inst.arg_data[0] > 0) {
            out.append_dem_instruction(inst);
            map.push_back(kept++);
          } else {
            map.push_back(kGone);
          }
        } else {
          out.append_dem_instruction(inst);
        }
      }
      for (auto& v : dem2err)
        if (v != kGone) v = map[v];
      cur = out;
    }
    for (const auto& inst : cur.instructions)  // utils.cc:253-261
      if (inst.type == stim::DemInstructionType::DEM_ERROR && inst.arg_data[0] > 0) errs.push_back(mechanism_of(inst));
    D = cur.count_detectors();
    O = cur.count_observables();
    E = errs.size();
    err2dem.assign(E, kGone);  // common.cc:228-239
    for (uint64_t i = 0; i < dem2err.size(); ++i)
      if (dem2err[i] != kGone && err2dem[dem2err[i]] == kGone) err2dem[dem2err[i]] = i;
    if (cfg.orders.empty()) {  // tesseract.cc:175-177
      cfg.orders.emplace_back(D);
      std::iota(cfg.orders[0].begin(), cfg.orders[0].end(), uint64_t{0});
    } else {
      for (const auto& o : cfg.orders)
        if (o.size() != D)
          throw std::invalid_argument("Each detector order list must have a size equal to the number of detectors.");
    }
    build_tables();
    if (cfg.create_visualization) {  // tesseract.cc:200-204, visualization.cc:5-24, common.cc:25-50, 90-94
      std::vector<std::vector<double>> coords(D);
      std::vector<double> shift;
      const stim::DetectorErrorModel flat_cur = cur.flattened();
      size_t next = 0;  // utils.cc:32-58 pushes one entry per detector INSTRUCTION, in instruction order
      for (const auto& inst : flat_cur.instructions)
        if (inst.type == stim::DemInstructionType::DEM_DETECTOR && next < coords.size()) coords[next++] = inst.arg_data;
      for (uint64_t d = 0; d < D; ++d) {
        std::stringstream ss;
        ss << "Detector D" << d << " coordinate (";
        const size_t e = std::min<size_t>(3, coords[d].size());
        for (size_t i = 0; i < e; ++i) ss << coords[d][i] << (i + 1 < e ? ", " : "");
        ss << ")";
        viz.push_back(ss.str());
      }
      auto list = [](const std::vector<int>& v) {
        std::string t = "[";
        for (size_t i = 0; i < v.size(); ++i) t += std::to_string(v[i]) + (i + 1 < v.size() ? " " : "");
        return t + "]";
      };
      for (const auto& m : errs) {
        std::stringstream ss;
        ss << std::fixed << std::setprecision(6) << m.cost;
        viz.push_back("Error{cost=" + ss.str() + ", symptom=Symptom{detectors=" + list(m.dets) + ", observables=" + list(m.obs) + "}}");
      }
    }
  }

  void log_node(int64_t chain_idx, const std::vector<bool>& fired) {  // visualization.cc:26-51
    if (!cfg.create_visualization) return;
    std::string a = "activated_errors = ", b = "activated_detectors = ";
    for (int64_t at = chain_idx; at != -1; at = chain[(size_t)at].parent) a += std::to_string(chain[(size_t)at].error) + ", ";
    for (uint64_t d = 0; d < D; ++d)
      if (fired[d]) b += std::to_string(d) + ", ";
    viz.push_back(a);
    viz.push_back(b);
  }

  void build_tables() {  // tesseract.cc:207-254
    std::vector<std::vector<int>> rows(D);
    cost.resize(E);
    per_det.resize(E);
    e_off.assign(E + 1, 0);
    for (uint64_t e = 0; e < E; ++e) {
      cost[e] = errs[e].cost;
      per_det[e] = errs[e].cost / errs[e].dets.size();
      e_off[e + 1] = e_off[e] + (uint32_t)errs[e].dets.size();
      for (int d : errs[e].dets) {
        edet.push_back(d);
        rows[(size_t)d].push_back((int)e);
      }
    }
    row_off.assign(D + 1, 0);
    for (uint64_t d = 0; d < D; ++d) {
      // The unstable libstdc++ sort on the ascending-index sequence is part of the contract.
      std::sort(rows[d].begin(), rows[d].end(), [this](size_t x, size_t y) { return per_det[x] < per_det[y]; });
      row_off[d + 1] = row_off[d] + (uint32_t)rows[d].size();
      row.insert(row.end(), rows[d].begin(), rows[d].end());
    }
    // eneighbors (tesseract.cc:229-254): detectors of errors sharing a detector, minus own.
    nb_off.assign(E + 1, 0);
    std::vector<char> mark(D, 0);
    std::vector<int> touched;
    for (uint64_t e = 0; e < E; ++e) {
      touched.clear();
      for (int d : errs[e].dets)
        for (uint32_t k = row_off[(size_t)d]; k < row_off[(size_t)d + 1]; ++k)
          for (int od : errs[(size_t)row[k]].dets)
            if (!mark[(size_t)od]) {
              mark[(size_t)od] = 1;
              touched.push_back(od);
            }
      for (int d : errs[e].dets) mark[(size_t)d] = 2;
      std::sort(touched.begin(), touched.end());
      for (int od : touched) {
        if (mark[(size_t)od] == 1) nb.push_back(od);
        mark[(size_t)od] = 0;
      }
      nb_off[e + 1] = (uint32_t)nb.size();
    }
    blocked.resize(E);
    count.resize(E);
    aoff = row_off;  // the search walks the full rows unless sparsification narrows them per shot
    arow = row;
    init_sparsify();
  }

  // tesseract.cc:128-154
  double detcost(uint64_t d, const std::vector<uint32_t>& blk, const std::vector<uint32_t>& cnt) {
    double best = kInf;
    uint32_t best_n = std::numeric_limits<uint32_t>::max();
    for (uint32_t k = aoff[d]; k < aoff[d + 1]; ++k) {
      const int e = arow[k];
      ++ctr.S;
      if (cost[(size_t)e] * best_n >= best * (e_off[(size_t)e + 1] - e_off[(size_t)e])) break;
      if (!blk[(size_t)e] && cost[(size_t)e] * best_n < best * cnt[(size_t)e]) {
        best = cost[(size_t)e];
        best_n = cnt[(size_t)e];
      }
    }
    return best / best_n + cfg.det_penalty;
  }

  // tesseract.cc:380-616
  void search(const std::vector<uint64_t>& hits, size_t order_index, uint64_t beam) {
    ++ctr.searches;
    predicted.clear();
    low_confidence = false;
    chain.clear();
    MinQueue pq;
    std::map<uint64_t, std::unordered_set<std::vector<bool>>> seen;
    std::vector<bool> root(D, false);
    std::fill(count.begin(), count.end(), 0u);
    std::fill(blocked.begin(), blocked.end(), 0u);
    for (uint64_t d : hits) {
      if (d >= D)
        throw std::runtime_error("Symptom " + std::to_string(d) + " references a detector >= num_detectors (= " +
                                 std::to_string(D) + ").");
      root[d] = true;
      for (uint32_t k = aoff[d]; k < aoff[d + 1]; ++k) ++count[(size_t)arow[k]];
    }
    double f0 = 0;
    for (uint64_t d : hits) f0 += detcost(d, blocked, count);
    if (f0 == kInf) {
      low_confidence = true;
      return;
    }
    uint64_t lo = hits.size(), hi = lo + beam;
    pq.push({f0, lo, 0, -1});
    ++ctr.Q;
    uint64_t pushed = 1;
    const std::vector<uint64_t>& order = cfg.orders[order_index];
    std::vector<bool> fired, next_fired;
    std::vector<double> memo(D);

    while (!pq.empty()) {
      const QueueNode node = pq.pop();
      ++ctr.P;
      if (node.num_dets > hi) continue;
      // State of the node from its chain (tesseract.cc:349-369).
      fired = root;
      std::fill(blocked.begin(), blocked.end(), 0u);
      std::fill(count.begin(), count.end(), 0u);
      for (int64_t at = node.chain; at != -1; at = chain[(size_t)at].parent) {
        const ChainLink& link = chain[(size_t)at];
        ++ctr.L;
        for (uint32_t k = aoff[link.min_det]; k < aoff[link.min_det + 1]; ++k) {
          blocked[(size_t)arow[k]] = 1;
          if ((uint32_t)arow[k] == link.error) break;
        }
        for (uint32_t k = e_off[link.error]; k < e_off[link.error + 1]; ++k) fired[(size_t)edet[k]] = !fired[(size_t)edet[k]];
        if (cfg.at_most_two)  // a detector an ancestor switched off takes no further error (see Config::at_most_two)
          for (uint32_t k = e_off[link.error]; k < e_off[link.error + 1]; ++k)
            if ((link.switched_off >> (k - e_off[link.error])) & 1)
              for (uint32_t q = aoff[(size_t)edet[k]]; q < aoff[(size_t)edet[k] + 1]; ++q) blocked[(size_t)arow[q]] = 1;
      }
      if (node.num_dets == 0) {  // tesseract.cc:441-473
        log_node(node.chain, fired);
        predicted.resize(node.depth);
        int64_t at = node.chain;
        for (uint64_t i = 0; i < node.depth; ++i) {
          predicted[node.depth - 1 - i] = err2dem[chain[(size_t)at].error];
          at = chain[(size_t)at].parent;
        }
        return;
      }
      if (cfg.no_revisit) {
        ++ctr.V;
        if (!seen[node.num_dets].insert(fired).second) continue;
      }
      ++ctr.X;
      log_node(node.chain, fired);
      if (node.num_dets < lo) {  // tesseract.cc:503-511
        lo = node.num_dets;
        if (cfg.no_revisit)
          for (uint64_t i = lo + beam + 1; i <= hi; ++i) seen[i].clear();
        hi = std::min(hi, lo + beam);
      }
      for (uint64_t d = 0; d < D; ++d)
        if (fired[d])
          for (uint32_t k = aoff[d]; k < aoff[d + 1]; ++k) ++count[(size_t)arow[k]];
      child_blocked = blocked;
      child_count = count;
      uint64_t min_det = kGone;
      for (uint64_t pos = 0; pos < D; ++pos)
        if (fired[order[pos]]) {
          min_det = order[pos];
          break;
        }
      std::fill(memo.begin(), memo.end(), -1.0);
      int64_t prev = -1;
      amt_undo.clear();
      for (uint32_t k = aoff[min_det]; k < aoff[min_det + 1]; ++k) {
        const int e = arow[k];
        if (blocked[(size_t)e]) continue;
        if (prev >= 0)  // undo the previous sibling's count deltas (tesseract.cc:536-543)
          for (uint32_t j = e_off[(size_t)prev]; j < e_off[(size_t)prev + 1]; ++j) {
            const int d = edet[j], delta = fired[(size_t)d] ? 1 : -1;
            for (uint32_t q = aoff[(size_t)d]; q < aoff[(size_t)d + 1]; ++q) child_count[(size_t)arow[q]] += (uint32_t)delta;
          }
        for (int b : amt_undo) child_blocked[(size_t)b] = 0;  // ... and the rows its switched-off detectors blocked
        amt_undo.clear();
        prev = e;
        next_fired = fired;
        child_blocked[(size_t)e] = 1;  // persists for later siblings
        double f = node.f + cost[(size_t)e];
        uint64_t nd = node.num_dets;
        uint32_t switched_off = 0;
        for (uint32_t j = e_off[(size_t)e]; j < e_off[(size_t)e + 1]; ++j) {
          const int d = edet[j];
          next_fired[(size_t)d] = !next_fired[(size_t)d];
          const int delta = next_fired[(size_t)d] ? 1 : -1;
          nd += (uint64_t)(int64_t)delta;
          for (uint32_t q = aoff[(size_t)d]; q < aoff[(size_t)d + 1]; ++q) child_count[(size_t)arow[q]] += (uint32_t)delta;
          if (fired[(size_t)d]) switched_off |= 1u << (j - e_off[(size_t)e]);
          if (cfg.at_most_two && fired[(size_t)d])
            for (uint32_t q = aoff[(size_t)d]; q < aoff[(size_t)d + 1]; ++q)
              if (!child_blocked[(size_t)arow[q]]) {
                child_blocked[(size_t)arow[q]] = 1;
                amt_undo.push_back(arow[q]);
              }
        }
        if (nd > hi) continue;
        if (cfg.no_revisit) {
          ++ctr.V;
          if (seen[nd].count(next_fired)) continue;
        }
        ctr.A += (e_off[(size_t)e + 1] - e_off[(size_t)e]) + (nb_off[(size_t)e + 1] - nb_off[(size_t)e]);
        for (uint32_t j = e_off[(size_t)e]; j < e_off[(size_t)e + 1]; ++j) {  // tesseract.cc:567-576
          const uint64_t d = (uint64_t)edet[j];
          if (fired[d]) {
            if (memo[d] == -1) memo[d] = detcost(d, blocked, count);
            f -= memo[d];
          } else {
            f += detcost(d, child_blocked, child_count);
          }
        }
        for (uint32_t j = nb_off[(size_t)e]; j < nb_off[(size_t)e + 1]; ++j) {  // tesseract.cc:578-585
          const uint64_t od = (uint64_t)nb[j];
          if (!fired[od] || !next_fired[od]) continue;
          if (memo[od] == -1) memo[od] = detcost(od, blocked, count);
          f -= memo[od];
          f += detcost(od, child_blocked, child_count);
        }
        if (f == kInf) continue;
        chain.push_back({(uint32_t)e, (uint32_t)min_det, node.chain, switched_off});
        pq.push({f, nd, node.depth + 1, (int64_t)chain.size() - 1});
        ++ctr.Q;
        if (++pushed > cfg.pqlimit) {  // tesseract.cc:597-605
          low_confidence = true;
          return;
        }
      }
    }
    low_confidence = true;  // tesseract.cc:609-615
  }

  double cost_of(const std::vector<uint64_t>& dem_errors) const {  // tesseract.cc:618-628
    double total = 0;
    for (uint64_t i : dem_errors) {
      uint64_t e = dem2err.at(i);
      if (e == kGone) throw std::invalid_argument("error index does not map to a retained decoder error");
      total += cost[e];
    }
    return total;
  }
  std::vector<int> observables_of(const std::vector<uint64_t>& dem_errors) const {  // tesseract.cc:630-658
    std::set<int> on;
    for (uint64_t i : dem_errors) {
      uint64_t e = dem2err.at(i);
      if (e == kGone) throw std::invalid_argument("error index does not map to a retained decoder error");
      for (int o : errs[e].obs)
        if (!on.erase(o)) on.insert(o);
    }
    return std::vector<int>(on.begin(), on.end());
  }

  // tesseract.cc:45-69
  static int suggest_limit_capped(size_t num_detectors, int base_degree, int max_limit) {
    if (base_degree < 0) throw std::invalid_argument("sparsify_base_degree must be >= 0.");
    if (num_detectors == 0 || max_limit <= 0) return 0;
    const double exponent = (double)base_degree - 2.0, max_result = (double)max_limit;
    const double log_result = exponent * std::log(4.5) - std::log(3.0) + std::log((double)num_detectors);
    if (!std::isfinite(log_result) || log_result >= std::log(max_result)) return max_limit;
    const double result = std::exp(log_result);
    if (!std::isfinite(result)) return max_limit;
    const double rounded = std::round(result);
    if (rounded >= max_result) return max_limit;
    return (int)rounded;
  }

  void init_sparsify() {  // tesseract.cc:256-292
    if (!cfg.sparsify) return;
    if (cfg.sparsify_base_degree <= 0) throw std::invalid_argument("sparsify_base_degree must be > 0 when sparsify_errors is enabled.");
    if (cfg.sparsify_max_degree < -1) throw std::invalid_argument("sparsify_max_degree must be >= -1.");
    if (cfg.sparsify_reactivate_limit < -1) throw std::invalid_argument("sparsify_reactivate_limit must be >= -1.");
    if (cfg.sparsify_max_degree >= 0 && cfg.sparsify_max_degree < cfg.sparsify_base_degree)
      throw std::invalid_argument("sparsify_max_degree must be >= sparsify_base_degree.");
    if (cfg.sparsify_reactivate_limit == -1)
      cfg.sparsify_reactivate_limit =
          suggest_limit_capped(D, cfg.sparsify_base_degree, (int)std::min<uint64_t>(E, (uint64_t)std::numeric_limits<int>::max()));
    for (uint64_t e = 0; e < E; ++e) {
      const int degree = (int)(e_off[e + 1] - e_off[e]);
      if (degree <= cfg.sparsify_base_degree) mandatory.push_back((int)e);
      else if (cfg.sparsify_max_degree == -1 || degree <= cfg.sparsify_max_degree) optional.push_back((int)e);
    }
    active.assign(E, 0);
  }

  void build_sparse_rows(const std::vector<uint64_t>& hits) {  // tesseract.cc:673-737
    std::vector<uint8_t> shot(D, 0);
    for (uint64_t d : hits)
      if (d < D) shot[d] = 1;
    std::fill(active.begin(), active.end(), 0);
    for (int e : mandatory) active[(size_t)e] = 1;
    struct Cand {
      int e, overlap, degree;
      double cost;
    };
    std::vector<Cand> cands;
    for (int e : optional) {
      int overlap = 0;
      for (uint32_t k = e_off[(size_t)e]; k < e_off[(size_t)e + 1]; ++k) overlap += shot[(size_t)edet[k]];
      if (overlap > 0) cands.push_back({e, overlap, (int)(e_off[(size_t)e + 1] - e_off[(size_t)e]), cost[(size_t)e]});
    }
    std::sort(cands.begin(), cands.end(), [](const Cand& a, const Cand& b) {
      if (a.overlap != b.overlap) return a.overlap > b.overlap;
      if (a.degree != b.degree) return a.degree < b.degree;
      if (a.cost != b.cost) return a.cost < b.cost;
      return a.e < b.e;
    });
    const size_t limit = std::min<size_t>((size_t)cfg.sparsify_reactivate_limit, cands.size());
    for (size_t i = 0; i < limit; ++i) active[(size_t)cands[i].e] = 1;
    aoff.assign(D + 1, 0);
    arow.clear();
    for (uint64_t d = 0; d < D; ++d) {
      for (uint32_t k = row_off[d]; k < row_off[d + 1]; ++k)
        if (active[(size_t)row[k]]) arow.push_back(row[k]);
      aoff[d + 1] = (uint32_t)arow.size();
    }
  }

  void decode(const std::vector<uint64_t>& hits) {  // tesseract.cc:295-347
    if (cfg.sparsify) build_sparse_rows(hits);
    std::vector<uint64_t> best;
    double best_cost = std::numeric_limits<double>::max();
    auto consider = [&] {
      double c = cost_of(predicted);
      if (!low_confidence && c < best_cost) {
        best = predicted;
        best_cost = c;
      }
    };
    if (cfg.beam_climbing) {
      int beam = 0;
      size_t order = 0;
      for (int t = 0; t < std::max(cfg.det_beam + 1, (int)cfg.orders.size()); ++t) {
        search(hits, order, (uint64_t)beam);
        consider();
        beam = (beam + 1) % (cfg.det_beam + 1);
        order = (order + 1) % cfg.orders.size();
      }
    } else {
      for (size_t order = 0; order < cfg.orders.size(); ++order) {
        search(hits, order, (uint64_t)cfg.det_beam);
        consider();
      }
    }
    predicted = best;
    low_confidence = best_cost == std::numeric_limits<double>::max();
  }
};

}  // namespace port

struct orc_handle {
  std::vector<std::unique_ptr<port::Decoder>> pool;
  std::vector<std::vector<uint64_t>> last_errors;
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

const char* orc_last_error(void) { return g_err.c_str(); }
const char* orc_kind(void) { return "port"; }

orc_handle* orc_create(const char* dem_text, const orc_config* cfg) {
  orc_handle* h = nullptr;
  int rc = guarded([&] {
    stim::DetectorErrorModel dem(dem_text);
    port::Config c;
    c.det_beam = cfg->det_beam;
    c.beam_climbing = cfg->beam_climbing != 0;
    c.no_revisit = cfg->no_revisit_dets != 0;
    c.merge = cfg->merge_errors != 0;
    c.pqlimit = cfg->pqlimit;
    c.det_penalty = cfg->det_penalty;
    c.sparsify = cfg->sparsify_errors != 0;
    c.sparsify_base_degree = cfg->sparsify_base_degree;
    c.sparsify_max_degree = cfg->sparsify_max_degree;
    c.sparsify_reactivate_limit = cfg->sparsify_reactivate_limit;
    c.at_most_two = cfg->at_most_two_errors_per_detector != 0;
    c.create_visualization = cfg->create_visualization != 0;
    uint64_t D = dem.flattened().count_detectors();
    for (uint64_t k = 0; k < cfg->num_det_orders; ++k)
      c.orders.emplace_back(cfg->det_orders + k * D, cfg->det_orders + (k + 1) * D);
    auto out = std::make_unique<orc_handle>();
    int n = cfg->num_threads < 1 ? 1 : cfg->num_threads;
    for (int t = 0; t < n; ++t) out->pool.push_back(std::make_unique<port::Decoder>(dem, c));
    h = out.release();
  });
  return rc == 0 ? h : nullptr;
}
void orc_destroy(orc_handle* h) { delete h; }

uint64_t orc_num_detectors(const orc_handle* h) { return h->pool[0]->D; }
uint64_t orc_num_observables(const orc_handle* h) { return h->pool[0]->O; }
uint64_t orc_num_errors(const orc_handle* h) { return h->pool[0]->E; }
uint64_t orc_num_dem_errors(const orc_handle* h) { return h->pool[0]->dem2err.size(); }

int orc_error_costs(const orc_handle* h, double* out) {
  std::copy(h->pool[0]->cost.begin(), h->pool[0]->cost.end(), out);
  return 0;
}
int orc_maps(const orc_handle* h, uint64_t* a, uint64_t* b) {
  const auto& d = *h->pool[0];
  if (a) std::copy(d.dem2err.begin(), d.dem2err.end(), a);
  if (b) std::copy(d.err2dem.begin(), d.err2dem.end(), b);
  return 0;
}
int64_t orc_csr(const orc_handle* h, int which, uint64_t* offsets, int32_t* idx) {
  const auto& d = *h->pool[0];
  std::vector<uint32_t> ooff;
  std::vector<int> oidx;
  const std::vector<uint32_t>* off = nullptr;
  const std::vector<int>* v = nullptr;
  switch (which) {
    case 0: off = &d.row_off; v = &d.row; break;
    case 1: off = &d.e_off; v = &d.edet; break;
    case 2: off = &d.nb_off; v = &d.nb; break;
    case 3:
      ooff.push_back(0);
      for (const auto& e : d.errs) {
        oidx.insert(oidx.end(), e.obs.begin(), e.obs.end());
        ooff.push_back((uint32_t)oidx.size());
      }
      off = &ooff;
      v = &oidx;
      break;
    default:
      g_err = "bad table id";
      return -1;
  }
  if (offsets) std::copy(off->begin(), off->end(), offsets);
  if (idx) std::copy(v->begin(), v->end(), idx);
  return (int64_t)v->size();
}

int orc_decode(orc_handle* h, const uint64_t* hits, uint64_t n_hits, int64_t order, uint64_t beam, uint64_t* errs_out,
               uint64_t errs_cap, uint64_t* n_errs, double* cost, int32_t* low_confidence) {
  return guarded([&] {
    auto& d = *h->pool[0];
    std::vector<uint64_t> dets(hits, hits + n_hits);
    if (order < 0) {
      d.decode(dets);
    } else {
      d.search(dets, (size_t)order, beam);
    }
    if (n_errs) *n_errs = d.predicted.size();
    if (errs_out) {
      if (d.predicted.size() > errs_cap) throw std::runtime_error("errs_out too small");
      std::copy(d.predicted.begin(), d.predicted.end(), errs_out);
    }
    if (cost) *cost = d.cost_of(d.predicted);
    if (low_confidence) *low_confidence = d.low_confidence ? 1 : 0;
  });
}

int orc_decode_batch_b8(orc_handle* h, const uint8_t* dets, uint64_t stride, uint64_t shots, uint8_t* obs_out,
                        double* cost_out, uint8_t* low_conf_out, double* wall_seconds, double* sum_shot_seconds) {
  return guarded([&] {
    const size_t D = h->pool[0]->D, O = h->pool[0]->O, obs_stride = (O + 7) / 8;
    h->last_errors.assign(shots, {});
    std::atomic<uint64_t> next{0};  // utils.h:88-92
    std::vector<double> busy(h->pool.size(), 0.0);
    std::vector<std::string> failure(h->pool.size());
    auto t0 = std::chrono::steady_clock::now();
    std::vector<std::thread> workers;
    for (size_t t = 0; t < h->pool.size(); ++t)
      workers.emplace_back([&, t] {
        try {
          auto& d = *h->pool[t];
          std::vector<uint64_t> hits;
          for (uint64_t s; (s = next++) < shots;) {
            hits.clear();
            const uint8_t* r = dets + s * stride;
            for (size_t k = 0; k < D; ++k)
              if ((r[k >> 3] >> (k & 7)) & 1) hits.push_back(k);
            auto a = std::chrono::high_resolution_clock::now();
            d.decode(hits);
            busy[t] += std::chrono::duration<double>(std::chrono::high_resolution_clock::now() - a).count();
            h->last_errors[s] = d.predicted;
            if (low_conf_out) low_conf_out[s] = d.low_confidence;
            if (cost_out) cost_out[s] = d.cost_of(d.predicted);
            if (obs_out) {
              uint8_t* o = obs_out + s * obs_stride;
              std::fill(o, o + obs_stride, (uint8_t)0);
              for (int ob : d.observables_of(d.predicted)) o[ob >> 3] ^= (uint8_t)(1u << (ob & 7));
            }
          }
        } catch (const std::exception& e) {
          failure[t] = e.what();
          next = shots;
        }
      });
    for (auto& w : workers) w.join();
    if (wall_seconds) *wall_seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    for (auto& f : failure)
      if (!f.empty()) throw std::runtime_error(f);
    if (sum_shot_seconds) *sum_shot_seconds = std::accumulate(busy.begin(), busy.end(), 0.0);
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
    for (uint64_t v : h->last_errors[s]) idx[n++] = v;
  }
  offsets[h->last_errors.size()] = n;
  return 0;
}

int orc_cost_from_errors(const orc_handle* h, const uint64_t* errs, uint64_t n, double* out) {
  return guarded([&] { *out = h->pool[0]->cost_of(std::vector<uint64_t>(errs, errs + n)); });
}
int orc_flipped_observables(const orc_handle* h, const uint64_t* errs, uint64_t n, uint8_t* bits) {
  return guarded([&] {
    size_t O = h->pool[0]->O;
    std::fill(bits, bits + (O + 7) / 8, (uint8_t)0);
    for (int ob : h->pool[0]->observables_of(std::vector<uint64_t>(errs, errs + n))) bits[ob >> 3] ^= (uint8_t)(1u << (ob & 7));
  });
}

// utils.cc:173-205, DetIndex only (the default method); the BFS / coordinate builders are checked
// against the reference build directly.
int orc_build_det_orders(const char* dem_text, uint64_t n_orders, int method, uint64_t seed, uint64_t* out) {
  return guarded([&] {
    if (method != 1) throw std::invalid_argument("port: only DetIndex orders are restated");
    stim::DetectorErrorModel dem(dem_text);
    const uint64_t D = dem.count_detectors();
    std::mt19937_64 rng(seed);
    std::uniform_int_distribution<int> coin(0, 1);
    for (uint64_t k = 0; k < n_orders; ++k) {
      bool descending = coin(rng) != 0;
      for (uint64_t i = 0; i < D; ++i) out[k * D + i] = descending ? D - 1 - i : i;
    }
  });
}
double orc_merge_weights(double a, double b) { return port::merge_weights(a, b); }

int orc_counters(orc_handle* h, uint64_t out[8], int reset) {
  for (int i = 0; i < 8; ++i) out[i] = 0;
  for (auto& d : h->pool) {
    const port::Counters& c = d->ctr;
    const uint64_t v[8] = {c.Q, c.P, c.X, c.L, c.S, c.A, c.V, c.searches};
    for (int i = 0; i < 8; ++i) out[i] += v[i];
    if (reset) d->ctr = port::Counters{};
  }
  return 0;
}

}  // extern "C"

extern "C" int orc_visualization_write(orc_handle* h, const char* path) {
  return guarded([&] {
    FILE* f = fopen(path, "w");
    if (!f) throw std::runtime_error(std::string("cannot open ") + path);
    for (const auto& line : h->pool[0]->viz) fprintf(f, "%s\n", line.c_str());
    fclose(f);
  });
}

#include "sampler.inc"
