#include "dem.h"

#include <algorithm>
#include <cctype>
#include <cstdio>
#include <cstdlib>
#include <stdexcept>

namespace tsr {

std::string format_double(double v) {
  char buf[64];
  for (int prec = 1; prec <= 17; ++prec) {
    std::snprintf(buf, sizeof(buf), "%.*g", prec, v);
    if (std::strtod(buf, nullptr) == v) break;
  }
  return buf;
}

std::string DemInstruction::str() const {
  std::string s = kind == DemKind::Error ? "error" : kind == DemKind::Detector ? "detector" : "logical_observable";
  if (!tag.empty()) s += "[" + tag + "]";
  if (!args.empty()) {
    s += "(";
    for (size_t i = 0; i < args.size(); ++i) {
      if (i) s += ",";
      s += format_double(args[i]);
    }
    s += ")";
  }
  for (const DemTarget& t : targets) {
    s += " ";
    if (t.is_sep()) {
      s += "^";
    } else {
      s += t.is_obs() ? "L" : "D";
      s += std::to_string(t.id());
    }
  }
  return s;
}

namespace {

// A recursive-descent reader that emits flattened instructions as it goes.  `repeat N { ... }` is
// handled by remembering where the body starts and re-reading it N times with the running shifts.
struct Reader {
  std::string_view s;
  size_t p = 0;
  uint64_t det_shift = 0;
  std::vector<double> coord_shift;
  DemModel* out;
  uint64_t work = 0;
  static constexpr uint64_t kMaxId = (uint64_t{1} << 62) - 1;  // ids and counts stay clear of the observable bit and the separator
  static constexpr uint64_t kMaxWork = uint64_t{1} << 28;
  static constexpr uint64_t kMaxInstructions = uint64_t{1} << 24;  // 245 x the largest model of the test set (BB [[144,12,12]] r=12)

  [[noreturn]] void fail(const std::string& why) const {
    size_t line = 1 + (size_t)std::count(s.begin(), s.begin() + (std::ptrdiff_t)std::min(p, s.size()), '\n');
    throw std::invalid_argument("detector error model, line " + std::to_string(line) + ": " + why);
  }
  void blanks() {
    while (p < s.size() && (s[p] == ' ' || s[p] == '\t' || s[p] == '\r')) ++p;
  }
  uint64_t uint() {
    if (p >= s.size() || !std::isdigit((unsigned char)s[p])) fail("expected an unsigned integer");
    uint64_t v = 0;
    while (p < s.size() && std::isdigit((unsigned char)s[p])) {
      const uint64_t digit = (uint64_t)(s[p++] - '0');
      if (v > (kMaxId - digit) / 10) fail("integer is too large");
      v = v * 10 + digit;
    }
    return v;
  }
  // Flattening is bounded: a `repeat 1000000000 { ... }` must be an error message, not an out-of-memory kill.
  void spend() {
    if (++work > kMaxWork) fail("more than " + std::to_string(kMaxWork) + " repeat iterations");
    if (out->instructions.size() >= kMaxInstructions)
      fail("the flattened model would exceed " + std::to_string(kMaxInstructions) + " instructions");
  }
  // Skips a balanced `{ ... }` body starting just after the '{'; returns offset just after '}'.
  size_t skip_body(size_t from) const {
    int depth = 1;
    size_t q = from;
    while (q < s.size()) {
      char c = s[q];
      if (c == '#') {
        while (q < s.size() && s[q] != '\n') ++q;
        continue;
      }
      if (c == '{') ++depth;
      if (c == '}' && --depth == 0) return q + 1;
      ++q;
    }
    fail("unterminated repeat block");
  }

  // Reads instructions until end of text (top level) or the matching '}' (nested).
  void block(bool nested) {
    while (true) {
      blanks();
      if (p >= s.size()) {
        if (nested) fail("unterminated repeat block");
        return;
      }
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
        if (!nested) fail("unmatched '}'");
        ++p;
        return;
      }
      instruction();
    }
  }

  void instruction() {
    size_t b = p;
    while (p < s.size() && (std::isalnum((unsigned char)s[p]) || s[p] == '_')) ++p;
    std::string name(s.substr(b, p - b));
    for (char& ch : name) ch = (char)std::tolower((unsigned char)ch);
    std::string tag;
    if (p < s.size() && s[p] == '[') {
      size_t e = s.find(']', p);
      if (e == std::string_view::npos) fail("unterminated tag");
      tag = std::string(s.substr(p + 1, e - p - 1));
      p = e + 1;
    }
    std::vector<double> args;
    if (p < s.size() && s[p] == '(') {
      ++p;
      while (true) {
        blanks();
        if (p >= s.size() || s[p] == '\n') fail("unterminated '('");
        if (s[p] == ')') {
          ++p;
          break;
        }
        size_t nb = p;
        while (p < s.size() && s[p] != ',' && s[p] != ')' && s[p] != '\n') ++p;
        std::string num(s.substr(nb, p - nb));
        char* end = nullptr;
        double v = std::strtod(num.c_str(), &end);
        if (end == num.c_str()) fail("bad number '" + num + "'");
        args.push_back(v);
        if (p < s.size() && s[p] == ',') ++p;
      }
    }
    if (name == "repeat") {
      blanks();
      uint64_t reps = uint();
      blanks();
      if (p >= s.size() || s[p] != '{') fail("repeat needs '{'");
      size_t body = p + 1;
      size_t after = skip_body(body);
      for (uint64_t r = 0; r < reps; ++r) {
        spend();
        p = body;
        block(true);
      }
      p = after;
      return;
    }
    if (name == "shift_detectors") {
      blanks();
      uint64_t n = 0;
      if (p < s.size() && std::isdigit((unsigned char)s[p])) n = uint();
      for (size_t k = 0; k < args.size(); ++k) {
        if (coord_shift.size() <= k) coord_shift.push_back(0);
        coord_shift[k] += args[k];
      }
      if (n > kMaxId - det_shift) fail("detector shift is too large");
      det_shift += n;
      rest_of_line_must_be_empty();
      return;
    }
    DemInstruction inst;
    if (name == "error") {
      inst.kind = DemKind::Error;
      if (args.size() != 1) fail("error needs exactly one probability");
      if (!(args[0] >= 0 && args[0] <= 1)) fail("error probability must be in [0,1]");
    } else if (name == "detector") {
      inst.kind = DemKind::Detector;
      for (size_t k = 0; k < args.size() && k < coord_shift.size(); ++k) args[k] += coord_shift[k];
    } else if (name == "logical_observable") {
      inst.kind = DemKind::LogicalObservable;
    } else {
      fail("unrecognized instruction '" + name + "'");
    }
    inst.args = std::move(args);
    inst.tag = std::move(tag);
    while (true) {
      blanks();
      if (p >= s.size() || s[p] == '\n' || s[p] == '#' || s[p] == '}') break;
      char t = s[p];
      if (t == '^') {
        ++p;
        inst.targets.push_back(DemTarget::sep());
      } else if (t == 'D' || t == 'd') {
        ++p;
        const uint64_t id = uint();
        if (id > kMaxId - det_shift) fail("detector id is too large");
        inst.targets.push_back(DemTarget::det(id + det_shift));
      } else if (t == 'L' || t == 'l') {
        ++p;
        inst.targets.push_back(DemTarget::obs(uint()));
      } else {
        fail(std::string("unexpected target character '") + t + "'");
      }
    }
    spend();
    out->instructions.push_back(std::move(inst));
  }

  void rest_of_line_must_be_empty() {
    blanks();
    if (p < s.size() && s[p] != '\n' && s[p] != '#' && s[p] != '}') fail("unexpected trailing text");
  }
};

}  // namespace

DemModel DemModel::parse(std::string_view text) {
  DemModel m;
  Reader r{text, 0, 0, {}, &m};
  r.block(false);
  return m;
}

uint64_t DemModel::count_errors() const {
  uint64_t n = 0;
  for (const auto& i : instructions) n += i.kind == DemKind::Error;
  return n;
}

uint64_t DemModel::count_detectors() const {
  uint64_t n = 0;
  for (const auto& i : instructions) {
    if (i.kind == DemKind::LogicalObservable) continue;
    for (const auto& t : i.targets)
      if (t.is_det()) n = std::max(n, t.id() + 1);
  }
  return n;
}

uint64_t DemModel::count_observables() const {
  uint64_t n = 0;
  for (const auto& i : instructions) {
    if (i.kind == DemKind::Detector) continue;
    for (const auto& t : i.targets)
      if (t.is_obs()) n = std::max(n, t.id() + 1);
  }
  return n;
}

std::string DemModel::str() const {
  std::string s;
  for (size_t k = 0; k < instructions.size(); ++k) {
    if (k) s += "\n";
    s += instructions[k].str();
  }
  return s;
}

void DemModel::append_error(double p, const std::vector<DemTarget>& targets, const std::string& tag) {
  DemInstruction inst;
  inst.kind = DemKind::Error;
  inst.args = {p};
  inst.targets = targets;
  inst.tag = tag;
  instructions.push_back(std::move(inst));
}

}  // namespace tsr
