// TEST INFRASTRUCTURE — NOT PRODUCT CODE.
//
// Minimal stand-in for the slice of the `stim` C++ API that the reference decoder sources
// (/root/reference/src/{tesseract,common,utils,visualization}.cc) touch, so that those sources can
// be compiled UNMODIFIED into oracle/_ref/ (see oracle/Makefile).  stim itself (pinned by the
// reference at commit bd60b73, MODULE.bazel:40-45) is not vendored and there is no network.
//
// What is implemented for real: detector-error-model text parsing, flattening (repeat blocks and
// shift_detectors), counting, appending and printing — the published DEM file-format semantics.
// What is stubbed to throw at run time: circuits, the DEM sampler and the frame simulator (they
// are referenced only by utils.cc:207-251, which the oracle never calls).
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may use anything built
// from this directory.
#ifndef ORACLE_SHIM_STIM_H
#define ORACLE_SHIM_STIM_H

#include <algorithm>
#include <array>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <iostream>
#include <list>
#include <map>
#include <random>
#include <set>
#include <sstream>
#include <stdexcept>
#include <string>
#include <string_view>
#include <unordered_map>
#include <unordered_set>
#include <utility>
#include <deque>
#include <vector>

namespace stim {

constexpr uint64_t DEM_OBSERVABLE_BIT = uint64_t{1} << 63;
constexpr uint64_t DEM_SEPARATOR_SYGIL = UINT64_MAX;
constexpr size_t MAX_BITWORD_WIDTH = 64;

struct DemTarget {
  uint64_t data = 0;
  static DemTarget observable_id(uint64_t id) {
    return DemTarget{DEM_OBSERVABLE_BIT | id};
  }
  static DemTarget relative_detector_id(uint64_t id) {
    return DemTarget{id};
  }
  static DemTarget separator() {
    return DemTarget{DEM_SEPARATOR_SYGIL};
  }
  bool is_separator() const {
    return data == DEM_SEPARATOR_SYGIL;
  }
  bool is_observable_id() const {
    return data != DEM_SEPARATOR_SYGIL && (data & DEM_OBSERVABLE_BIT);
  }
  bool is_relative_detector_id() const {
    return data != DEM_SEPARATOR_SYGIL && !(data & DEM_OBSERVABLE_BIT);
  }
  uint64_t raw_id() const {
    return data & ~DEM_OBSERVABLE_BIT;
  }
  uint64_t val() const {
    if (is_separator()) throw std::invalid_argument("Separator doesn't have an integer value.");
    return raw_id();
  }
  bool operator==(const DemTarget& o) const {
    return data == o.data;
  }
  std::string str() const {
    if (is_separator()) return "^";
    return (is_observable_id() ? "L" : "D") + std::to_string(raw_id());
  }
};

enum class DemInstructionType : uint8_t {
  DEM_ERROR,
  DEM_SHIFT_DETECTORS,
  DEM_DETECTOR,
  DEM_LOGICAL_OBSERVABLE,
  DEM_REPEAT_BLOCK,
};

inline std::string shim_format_double(double v) {
  // Shortest text that round-trips (what stim's printer aims for).
  char buf[64];
  for (int prec = 1; prec <= 17; ++prec) {
    snprintf(buf, sizeof(buf), "%.*g", prec, v);
    if (strtod(buf, nullptr) == v) break;
  }
  return buf;
}

struct DemInstruction {
  std::vector<double> arg_data;
  std::vector<DemTarget> target_data;
  std::string tag;
  DemInstructionType type = DemInstructionType::DEM_ERROR;

  std::string str() const {
    std::stringstream ss;
    switch (type) {
      case DemInstructionType::DEM_ERROR:
        ss << "error";
        break;
      case DemInstructionType::DEM_SHIFT_DETECTORS:
        ss << "shift_detectors";
        break;
      case DemInstructionType::DEM_DETECTOR:
        ss << "detector";
        break;
      case DemInstructionType::DEM_LOGICAL_OBSERVABLE:
        ss << "logical_observable";
        break;
      case DemInstructionType::DEM_REPEAT_BLOCK:
        ss << "repeat";
        break;
    }
    if (!tag.empty()) ss << "[" << tag << "]";
    if (!arg_data.empty()) {
      ss << "(";
      for (size_t i = 0; i < arg_data.size(); ++i) {
        if (i) ss << ",";
        ss << shim_format_double(arg_data[i]);
      }
      ss << ")";
    }
    if (type == DemInstructionType::DEM_SHIFT_DETECTORS || type == DemInstructionType::DEM_REPEAT_BLOCK) {
      if (!target_data.empty()) ss << " " << target_data[0].data;
    } else {
      for (const auto& t : target_data) ss << " " << t.str();
    }
    return ss.str();
  }
};

struct DetectorErrorModel {
  std::vector<DemInstruction> instructions;
  std::vector<DetectorErrorModel> blocks;

  DetectorErrorModel() = default;
  DetectorErrorModel(const DetectorErrorModel&) = default;
  DetectorErrorModel(DetectorErrorModel&&) = default;
  DetectorErrorModel& operator=(const DetectorErrorModel&) = default;
  DetectorErrorModel& operator=(DetectorErrorModel&&) = default;
  // The reference iterates `for (... : common::flatten(dem).instructions)` (utils.cc:34, used by get_detector_coords):
  // the temporary dies before the loop body runs (no lifetime extension through the member access before C++23), so the
  // loop reads freed storage.  With stim's allocator pattern that happens to work; to keep this checker build
  // well-defined the shim parks a dying model's instruction storage for a while instead of freeing it at once.
  ~DetectorErrorModel() {
    if (instructions.empty()) return;
    thread_local std::deque<std::vector<DemInstruction>> parked;
    parked.push_back(std::move(instructions));
    if (parked.size() > 32) parked.pop_front();
  }
  explicit DetectorErrorModel(std::string_view text) {
    size_t pos = 0;
    parse_block(text, pos, /*nested=*/false);
  }
  explicit DetectorErrorModel(const char* text) : DetectorErrorModel(std::string_view(text)) {}

  void append_dem_instruction(const DemInstruction& instruction) {
    if (instruction.type == DemInstructionType::DEM_REPEAT_BLOCK)
      throw std::invalid_argument("shim: cannot append a repeat block by instruction");
    instructions.push_back(instruction);
  }
  void append_error_instruction(double probability, const std::vector<DemTarget>& targets,
                                std::string_view tag) {
    if (!(probability >= 0 && probability <= 1))
      throw std::invalid_argument("error probability must be in [0,1]");
    DemInstruction inst;
    inst.type = DemInstructionType::DEM_ERROR;
    inst.arg_data = {probability};
    inst.target_data = targets;
    inst.tag = std::string(tag);
    instructions.push_back(std::move(inst));
  }

  // Published semantics of DetectorErrorModel.flattened(): repeat blocks unrolled, detector
  // offsets and coordinate shifts applied, shift_detectors instructions removed.
  DetectorErrorModel flattened() const {
    DetectorErrorModel out;
    uint64_t det_shift = 0;
    std::vector<double> coord_shift;
    flatten_into(out, det_shift, coord_shift);
    return out;
  }

  uint64_t count_errors() const {
    uint64_t n = 0;
    for (const auto& inst : instructions) {
      if (inst.type == DemInstructionType::DEM_ERROR) {
        ++n;
      } else if (inst.type == DemInstructionType::DEM_REPEAT_BLOCK) {
        n += inst.target_data[0].data * blocks[inst.target_data[1].data].count_errors();
      }
    }
    return n;
  }
  uint64_t count_detectors() const {
    uint64_t shift = 0;
    uint64_t n = 0;
    count_detectors_into(shift, n);
    return n;
  }
  uint64_t count_observables() const {
    uint64_t n = 0;
    for (const auto& inst : instructions) {
      if (inst.type == DemInstructionType::DEM_REPEAT_BLOCK) {
        n = std::max(n, blocks[inst.target_data[1].data].count_observables());
      } else if (inst.type == DemInstructionType::DEM_ERROR ||
                 inst.type == DemInstructionType::DEM_LOGICAL_OBSERVABLE) {
        for (const auto& t : inst.target_data)
          if (t.is_observable_id()) n = std::max(n, t.raw_id() + 1);
      }
    }
    return n;
  }

  std::string str() const {
    std::stringstream ss;
    print_into(ss, 0);
    return ss.str();
  }

 private:
  static void skip_ws(std::string_view s, size_t& p) {
    while (p < s.size() && (s[p] == ' ' || s[p] == '\t' || s[p] == '\r')) ++p;
  }
  void parse_block(std::string_view s, size_t& p, bool nested) {
    while (p < s.size()) {
      skip_ws(s, p);
      if (p >= s.size()) break;
      char c = s[p];
      if (c == '\n') {
        ++p;
        continue;
      }
      if (c == '#') {
        while (p < s.size() && s[p] != '\n') ++p;
        continue;
      }
      if (c == '}') {
        if (!nested) throw std::invalid_argument("shim: unmatched '}' in detector error model");
        ++p;
        return;
      }
      size_t b = p;
      while (p < s.size() && (isalnum((unsigned char)s[p]) || s[p] == '_')) ++p;
      std::string name(s.substr(b, p - b));
      for (auto& ch : name) ch = (char)tolower((unsigned char)ch);
      DemInstruction inst;
      if (name == "error") {
        inst.type = DemInstructionType::DEM_ERROR;
      } else if (name == "detector") {
        inst.type = DemInstructionType::DEM_DETECTOR;
      } else if (name == "logical_observable") {
        inst.type = DemInstructionType::DEM_LOGICAL_OBSERVABLE;
      } else if (name == "shift_detectors") {
        inst.type = DemInstructionType::DEM_SHIFT_DETECTORS;
      } else if (name == "repeat") {
        inst.type = DemInstructionType::DEM_REPEAT_BLOCK;
      } else {
        throw std::invalid_argument("shim: unrecognized DEM instruction '" + name + "'");
      }
      if (p < s.size() && s[p] == '[') {
        size_t e = s.find(']', p);
        if (e == std::string_view::npos) throw std::invalid_argument("shim: unterminated tag");
        inst.tag = std::string(s.substr(p + 1, e - p - 1));
        p = e + 1;
      }
      if (p < s.size() && s[p] == '(') {
        ++p;
        while (true) {
          skip_ws(s, p);
          if (p < s.size() && s[p] == ')') {
            ++p;
            break;
          }
          std::string num;
          while (p < s.size() && s[p] != ',' && s[p] != ')' && s[p] != '\n') num.push_back(s[p++]);
          char* end = nullptr;
          double v = strtod(num.c_str(), &end);
          if (end == num.c_str()) throw std::invalid_argument("shim: bad number '" + num + "'");
          inst.arg_data.push_back(v);
          if (p < s.size() && s[p] == ',') ++p;
          if (p >= s.size() || s[p] == '\n') throw std::invalid_argument("shim: unterminated '('");
        }
      }
      if (inst.type == DemInstructionType::DEM_REPEAT_BLOCK) {
        skip_ws(s, p);
        uint64_t reps = parse_uint(s, p);
        skip_ws(s, p);
        if (p >= s.size() || s[p] != '{') throw std::invalid_argument("shim: repeat needs '{'");
        ++p;
        DetectorErrorModel body;
        body.parse_block(s, p, true);
        inst.target_data.push_back(DemTarget{reps});
        inst.target_data.push_back(DemTarget{(uint64_t)blocks.size()});
        blocks.push_back(std::move(body));
        instructions.push_back(std::move(inst));
        continue;
      }
      while (true) {
        skip_ws(s, p);
        if (p >= s.size() || s[p] == '\n' || s[p] == '#' || s[p] == '}') break;
        char t = s[p];
        if (t == '^') {
          ++p;
          inst.target_data.push_back(DemTarget::separator());
        } else if (t == 'D' || t == 'd') {
          ++p;
          inst.target_data.push_back(DemTarget::relative_detector_id(parse_uint(s, p)));
        } else if (t == 'L' || t == 'l') {
          ++p;
          inst.target_data.push_back(DemTarget::observable_id(parse_uint(s, p)));
        } else if (isdigit((unsigned char)t) && inst.type == DemInstructionType::DEM_SHIFT_DETECTORS) {
          inst.target_data.push_back(DemTarget{parse_uint(s, p)});
        } else {
          throw std::invalid_argument(std::string("shim: unexpected target character '") + t + "'");
        }
      }
      if (inst.type == DemInstructionType::DEM_ERROR) {
        if (inst.arg_data.size() != 1)
          throw std::invalid_argument("shim: error instruction needs exactly one probability");
        if (!(inst.arg_data[0] >= 0 && inst.arg_data[0] <= 1))
          throw std::invalid_argument("shim: error probability must be in [0,1]");
      }
      if (inst.type == DemInstructionType::DEM_SHIFT_DETECTORS && inst.target_data.empty())
        inst.target_data.push_back(DemTarget{0});
      instructions.push_back(std::move(inst));
    }
    if (nested) throw std::invalid_argument("shim: unterminated repeat block");
  }
  static uint64_t parse_uint(std::string_view s, size_t& p) {
    if (p >= s.size() || !isdigit((unsigned char)s[p]))
      throw std::invalid_argument("shim: expected an unsigned integer");
    uint64_t v = 0;
    while (p < s.size() && isdigit((unsigned char)s[p])) v = v * 10 + (uint64_t)(s[p++] - '0');
    return v;
  }
  void flatten_into(DetectorErrorModel& out, uint64_t& det_shift, std::vector<double>& coord_shift) const {
    for (const auto& inst : instructions) {
      switch (inst.type) {
        case DemInstructionType::DEM_SHIFT_DETECTORS:
          for (size_t k = 0; k < inst.arg_data.size(); ++k) {
            if (coord_shift.size() <= k) coord_shift.push_back(0);
            coord_shift[k] += inst.arg_data[k];
          }
          det_shift += inst.target_data[0].data;
          break;
        case DemInstructionType::DEM_REPEAT_BLOCK: {
          uint64_t reps = inst.target_data[0].data;
          const auto& body = blocks[inst.target_data[1].data];
          for (uint64_t r = 0; r < reps; ++r) body.flatten_into(out, det_shift, coord_shift);
          break;
        }
        case DemInstructionType::DEM_DETECTOR: {
          DemInstruction c = inst;
          for (size_t k = 0; k < c.arg_data.size(); ++k)
            if (k < coord_shift.size()) c.arg_data[k] += coord_shift[k];
          for (auto& t : c.target_data)
            if (t.is_relative_detector_id()) t.data += det_shift;
          out.instructions.push_back(std::move(c));
          break;
        }
        default: {
          DemInstruction c = inst;
          for (auto& t : c.target_data)
            if (t.is_relative_detector_id()) t.data += det_shift;
          out.instructions.push_back(std::move(c));
        }
      }
    }
  }
  void count_detectors_into(uint64_t& shift, uint64_t& n) const {
    for (const auto& inst : instructions) {
      switch (inst.type) {
        case DemInstructionType::DEM_SHIFT_DETECTORS:
          shift += inst.target_data[0].data;
          break;
        case DemInstructionType::DEM_REPEAT_BLOCK: {
          uint64_t reps = inst.target_data[0].data;
          const auto& body = blocks[inst.target_data[1].data];
          for (uint64_t r = 0; r < reps; ++r) body.count_detectors_into(shift, n);
          break;
        }
        case DemInstructionType::DEM_ERROR:
        case DemInstructionType::DEM_DETECTOR:
          for (const auto& t : inst.target_data)
            if (t.is_relative_detector_id()) n = std::max(n, shift + t.raw_id() + 1);
          break;
        default:
          break;
      }
    }
  }
  void print_into(std::stringstream& ss, int indent) const {
    for (size_t k = 0; k < instructions.size(); ++k) {
      const auto& inst = instructions[k];
      if (k || indent) ss << "\n";
      ss << std::string(indent, ' ');
      if (inst.type == DemInstructionType::DEM_REPEAT_BLOCK) {
        ss << "repeat " << inst.target_data[0].data << " {";
        blocks[inst.target_data[1].data].print_into(ss, indent + 4);
        ss << "\n" << std::string(indent, ' ') << "}";
      } else {
        ss << inst.str();
      }
    }
  }
};

// ---- run-time stubs: present only so the unmodified utils.cc links -----------------------------

template <size_t W>
struct simd_bits {
  std::vector<uint8_t> bits;
  struct bit_ref {
    uint8_t& b;
    operator bool() const {
      return b != 0;
    }
    bit_ref& operator=(bool v) {
      b = v;
      return *this;
    }
    bit_ref& operator^=(bool v) {
      b ^= (uint8_t)v;
      return *this;
    }
  };
  simd_bits() : bits(64) {}
  explicit simd_bits(size_t n) : bits(n) {}
  bit_ref operator[](size_t k) {
    if (k >= bits.size()) bits.resize(k + 1);
    return bit_ref{bits[k]};
  }
  bool operator[](size_t k) const {
    return k < bits.size() && bits[k];
  }
  void clear() {
    std::fill(bits.begin(), bits.end(), 0);
  }
};

template <size_t W>
struct simd_bit_table {
  simd_bit_table transposed() const {
    throw std::runtime_error("shim: simd_bit_table is a stub");
  }
  simd_bits<W> operator[](size_t) const {
    throw std::runtime_error("shim: simd_bit_table is a stub");
  }
};

struct SparseShot {
  std::vector<uint64_t> hits;
  simd_bits<64> obs_mask;
  void clear() {
    hits.clear();
    obs_mask.clear();
  }
};

struct Circuit {
  uint64_t count_detectors() const {
    throw std::runtime_error("shim: stim::Circuit is a stub");
  }
};

template <size_t W>
struct DemSampler {
  size_t num_detectors = 0;
  size_t num_observables = 0;
  std::vector<std::vector<bool>> det_buffer;
  std::vector<std::vector<bool>> obs_buffer;
  DemSampler(const DetectorErrorModel&, std::mt19937_64, size_t) {
    throw std::runtime_error("shim: stim::DemSampler is a stub");
  }
  void resample(bool) {}
};

template <size_t W>
std::pair<simd_bit_table<W>, simd_bit_table<W>> sample_batch_detection_events(const Circuit&, size_t,
                                                                              std::mt19937_64&) {
  throw std::runtime_error("shim: sample_batch_detection_events is a stub");
}

}  // namespace stim

#endif  // ORACLE_SHIM_STIM_H
