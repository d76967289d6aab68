// Stabilizer-circuit text -> detector error model, and an i.i.d. DEM shot sampler.
//
// The reference gets both from stim (tesseract_main.cc:191-211 circuit_to_detector_error_model with
// decompose_errors=false / approximate_disjoint_errors_threshold=1; utils.cc:207-251 sampling).
// stim is not available, so this file carries the slice of it the reference's test circuits need
// (SURVEY.md App. B.2 / D.2): R RX M MX MPP H CX CZ X_ERROR Z_ERROR DEPOLARIZE1 DEPOLARIZE2 DETECTOR
// OBSERVABLE_INCLUDE REPEAT TICK QUBIT_COORDS SHIFT_COORDS.
//
// Method: one backward sweep over the REPEAT-unrolled circuit keeping, per qubit, the set of
// detectors/observables that an X (resp. Z) fault at the current time would flip.  Every noise
// channel then emits its independent Pauli components with those sets as symptoms; identical
// symptoms are XOR-combined.
#include "circuit.h"

#include <algorithm>
#include <cctype>
#include <cmath>
#include <cstdlib>
#include <map>
#include <stdexcept>

namespace tsr {
namespace {

struct Target {
  uint32_t qubit = 0;
  char pauli = 0;     // 'X','Y','Z' inside an MPP product
  bool combine = false;  // joined to the previous target with '*'
  int64_t rec = 0;    // rec[-k] lookback (negative) for DETECTOR / OBSERVABLE_INCLUDE
  bool is_rec = false;
};
struct Op {
  std::string name;
  std::vector<double> args;
  std::vector<Target> targets;
};

struct Parser {
  std::string_view s;
  size_t p = 0;
  std::vector<Op>* out;
  uint64_t work = 0;
  static constexpr uint64_t kMaxInt = (uint64_t{1} << 62) - 1;
  static constexpr uint64_t kMaxWork = uint64_t{1} << 28, kMaxOps = uint64_t{1} << 24;
  static constexpr uint64_t kMaxQubit = (uint64_t{1} << 24) - 1;  // per-qubit frames are allocated up to the largest index

  [[noreturn]] void fail(const std::string& why) const {
    size_t line = 1 + (size_t)std::count(s.begin(), s.begin() + (std::ptrdiff_t)std::min(p, s.size()), '\n');
    throw std::invalid_argument("circuit, line " + std::to_string(line) + ": " + why);
  }
  void blanks() {
    while (p < s.size() && (s[p] == ' ' || s[p] == '\t' || s[p] == '\r')) ++p;
  }
  uint64_t uint() {
    if (p >= s.size() || !std::isdigit((unsigned char)s[p])) fail("expected an unsigned integer");
    uint64_t v = 0;
    while (p < s.size() && std::isdigit((unsigned char)s[p])) {
      const uint64_t digit = (uint64_t)(s[p++] - '0');
      if (v > (kMaxInt - digit) / 10) fail("integer is too large");
      v = v * 10 + digit;
    }
    return v;
  }
  // Flattening is bounded: REPEAT 1000000000 { ... } must be an error message, not an out-of-memory kill.
  void spend() {
    if (++work > kMaxWork) fail("more than " + std::to_string(kMaxWork) + " REPEAT iterations");
    if (out->size() >= kMaxOps) fail("the flattened circuit would exceed " + std::to_string(kMaxOps) + " instructions");
  }
  size_t skip_body(size_t from) const {
    int depth = 1;
    for (size_t q = from; q < s.size(); ++q) {
      if (s[q] == '#') {
        while (q < s.size() && s[q] != '\n') ++q;
        continue;
      }
      if (s[q] == '{') ++depth;
      if (s[q] == '}' && --depth == 0) return q + 1;
    }
    fail("unterminated REPEAT block");
  }
  void block(bool nested) {
    while (true) {
      blanks();
      if (p >= s.size()) {
        if (nested) fail("unterminated REPEAT block");
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
      op();
    }
  }
  // Which target forms an instruction may carry (anything else would index frames that do not exist).
  void check_targets(const Op& o) const {
    const bool annotation = o.name == "DETECTOR" || o.name == "OBSERVABLE_INCLUDE";
    for (size_t i = 0; i < o.targets.size(); ++i) {
      const Target& t = o.targets[i];
      if (t.is_rec && !annotation) fail(o.name + " with a measurement-record target (classical control) is not supported");
      if (o.name == "MPP") {
        if (!t.pauli) fail("MPP targets must be Pauli products such as X1*Z2");
        if (t.combine && i == 0) fail("MPP product starts with '*'");
      } else if (t.pauli || t.combine) {
        fail(o.name + " with Pauli targets is not supported by this analyzer");
      }
    }
    if (o.name == "OBSERVABLE_INCLUDE" && !o.args.empty() &&
        !(o.args[0] >= 0 && o.args[0] < (double)(1 << 24) && o.args[0] == std::floor(o.args[0])))
      fail("OBSERVABLE_INCLUDE index must be an integer below 16777216");
  }
  void op() {
    size_t b = p;
    while (p < s.size() && (std::isalnum((unsigned char)s[p]) || s[p] == '_')) ++p;
    Op o;
    o.name = std::string(s.substr(b, p - b));
    for (char& ch : o.name) ch = (char)std::toupper((unsigned char)ch);
    if (p < s.size() && s[p] == '[') {  // tag, ignored
      size_t e = s.find(']', p);
      if (e == std::string_view::npos) fail("unterminated tag");
      p = e + 1;
    }
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
        o.args.push_back(std::strtod(num.c_str(), nullptr));
        if (p < s.size() && s[p] == ',') ++p;
      }
    }
    if (o.name == "REPEAT") {
      blanks();
      uint64_t reps = uint();
      blanks();
      if (p >= s.size() || s[p] != '{') fail("REPEAT needs '{'");
      size_t body = p + 1, after = skip_body(body);
      for (uint64_t r = 0; r < reps; ++r) {
        spend();
        p = body;
        block(true);
      }
      p = after;
      return;
    }
    auto qubit = [&]() {
      const uint64_t q = uint();
      if (q > kMaxQubit) fail("qubit index " + std::to_string(q) + " is too large");
      return (uint32_t)q;
    };
    while (true) {
      blanks();
      if (p >= s.size() || s[p] == '\n' || s[p] == '#' || s[p] == '}') break;
      Target t;
      if (s[p] == '*') {
        ++p;
        t.combine = true;
        blanks();
        if (p >= s.size()) fail("'*' needs a target after it");
      }
      char c = s[p];
      if (c == 'r' || c == 'R') {
        if (s.substr(p, 4) != "rec[") fail("bad record target");
        p += 4;
        if (p >= s.size() || s[p] != '-') fail("record lookback must be negative");
        ++p;
        t.rec = -(int64_t)uint();
        t.is_rec = true;
        if (p >= s.size() || s[p] != ']') fail("bad record target");
        ++p;
      } else if (c == 'X' || c == 'Y' || c == 'Z' || c == 'x' || c == 'y' || c == 'z') {
        t.pauli = (char)std::toupper((unsigned char)c);
        ++p;
        t.qubit = qubit();
      } else if (c == '!') {
        ++p;  // inverted result: irrelevant for error propagation
        continue;
      } else {
        t.qubit = qubit();
      }
      o.targets.push_back(t);
    }
    check_targets(o);
    spend();
    out->push_back(std::move(o));
  }
};

using SymSet = std::vector<int32_t>;  // ascending ids: detector d -> d, observable o -> kObsBase + o
constexpr int32_t kObsBase = 1 << 30;

void xor_into(SymSet& a, const SymSet& b) {
  if (b.empty()) return;
  SymSet r;
  r.reserve(a.size() + b.size());
  size_t i = 0, j = 0;
  while (i < a.size() && j < b.size()) {
    if (a[i] < b[j]) {
      r.push_back(a[i++]);
    } else if (b[j] < a[i]) {
      r.push_back(b[j++]);
    } else {
      ++i;
      ++j;
    }
  }
  r.insert(r.end(), a.begin() + (std::ptrdiff_t)i, a.end());
  r.insert(r.end(), b.begin() + (std::ptrdiff_t)j, b.end());
  a.swap(r);
}
SymSet xor_of(const SymSet& a, const SymSet& b) {
  SymSet r = a;
  xor_into(r, b);
  return r;
}

bool is_measure(const std::string& n) {
  return n == "M" || n == "MZ" || n == "MX" || n == "MPP" || n == "MR" || n == "MRZ" || n == "MRX";
}

}  // namespace

std::string circuit_to_dem_text(std::string_view circuit_text) {
  std::vector<Op> ops;
  Parser parser{circuit_text, 0, &ops};
  parser.block(false);

  // Forward pass: number the measurements, collect which detectors/observables read each one.
  size_t num_meas = 0, num_det = 0;
  uint32_t num_qubits = 0;
  for (const Op& o : ops) {
    for (const Target& t : o.targets)
      if (!t.is_rec) num_qubits = std::max(num_qubits, t.qubit + 1);
  }
  std::vector<SymSet> readers;  // per measurement
  std::vector<std::vector<double>> det_coords;
  std::vector<double> coord_shift;
  std::vector<size_t> meas_begin(ops.size(), 0);  // index of the op's first measurement
  for (size_t k = 0; k < ops.size(); ++k) {
    const Op& o = ops[k];
    meas_begin[k] = num_meas;
    if (is_measure(o.name)) {
      size_t n = 0;
      if (o.name == "MPP") {
        for (const Target& t : o.targets) n += !t.combine;
      } else {
        n = o.targets.size();
      }
      num_meas += n;
      readers.resize(num_meas);
    } else if (o.name == "DETECTOR" || o.name == "OBSERVABLE_INCLUDE") {
      int32_t id;
      if (o.name == "DETECTOR") {
        id = (int32_t)num_det++;
        std::vector<double> c = o.args;
        for (size_t i = 0; i < c.size() && i < coord_shift.size(); ++i) c[i] += coord_shift[i];
        det_coords.push_back(std::move(c));
      } else {
        if (o.args.empty()) throw std::invalid_argument("OBSERVABLE_INCLUDE needs an index");
        id = kObsBase + (int32_t)o.args[0];
      }
      for (const Target& t : o.targets) {
        if (!t.is_rec) throw std::invalid_argument(o.name + " targets must be rec[-k]");
        int64_t m = (int64_t)num_meas + t.rec;
        if (m < 0) throw std::invalid_argument("record lookback before the first measurement");
        xor_into(readers[(size_t)m], SymSet{id});
      }
    } else if (o.name == "SHIFT_COORDS") {
      for (size_t i = 0; i < o.args.size(); ++i) {
        if (coord_shift.size() <= i) coord_shift.push_back(0);
        coord_shift[i] += o.args[i];
      }
    }
  }

  // Backward pass.
  std::vector<SymSet> xs(num_qubits), zs(num_qubits);
  std::map<SymSet, double> acc;
  std::vector<SymSet> first_seen;  // emission order (reverse circuit time); reversed at the end
  auto emit = [&](const SymSet& sym, double p) {
    if (sym.empty() || !(p > 0)) return;
    auto it = acc.find(sym);
    if (it == acc.end()) {
      acc.emplace(sym, p);
      first_seen.push_back(sym);
    } else {
      it->second = it->second * (1 - p) + p * (1 - it->second);
    }
  };
  // A reset into the Z (X) basis absorbs a Z (X) error placed right after it.  Detectors or observables that such an
  // error would still flip do not have a deterministic value ("gauge" detectors; stim reports them as error(0.5)
  // mechanisms when asked to allow them, tesseract_main.cc:206-211).  This analyzer does not model them: refuse
  // instead of emitting a DEM that silently lacks them.
  auto require_deterministic = [&](const SymSet& absorbed, const std::string& at) {
    if (!absorbed.empty())
      throw std::invalid_argument("the circuit has a non-deterministic detector or observable (sensitive to the basis of a '" + at +
                                  "'); gauge detectors are not supported by this analyzer");
  };
  for (size_t k = ops.size(); k-- > 0;) {
    const Op& o = ops[k];
    const std::string& n = o.name;
    double arg = o.args.empty() ? 0.0 : o.args[0];
    if (n == "M" || n == "MZ" || n == "MX" || n == "MR" || n == "MRZ" || n == "MRX") {
      // MR / MRX = the measurement followed by a reset of the same qubit: walking backwards the reset comes
      // first (nothing after it depends on the qubit's earlier frame), then the measurement's readers.
      const bool resets = n[1] == 'R', x_basis = n.back() == 'X';
      for (size_t i = o.targets.size(); i-- > 0;) {
        const SymSet& r = readers[meas_begin[k] + i];
        uint32_t q = o.targets[i].qubit;
        if (resets) {
          require_deterministic(x_basis ? xs[q] : zs[q], n);
          xs[q].clear();
          zs[q].clear();
        }
        emit(r, arg);
        xor_into(x_basis ? zs[q] : xs[q], r);
      }
    } else if (n == "MPP") {
      // Walk products from the last to the first; factors of one product share a record.
      size_t prod = meas_begin[k];
      for (const Target& t : o.targets) prod += !t.combine;
      for (size_t i = o.targets.size(); i-- > 0;) {
        const Target& t = o.targets[i];
        const SymSet& r = readers[prod - 1];
        if (t.pauli == 'X' || t.pauli == 'Y') xor_into(zs[t.qubit], r);
        if (t.pauli == 'Z' || t.pauli == 'Y') xor_into(xs[t.qubit], r);
        if (!t.combine) {
          emit(r, arg);
          --prod;
        }
      }
    } else if (n == "R" || n == "RZ" || n == "RX") {
      for (const Target& t : o.targets) {
        require_deterministic(n == "RX" ? xs[t.qubit] : zs[t.qubit], n);
        xs[t.qubit].clear();
        zs[t.qubit].clear();
      }
    } else if (n == "H") {
      for (const Target& t : o.targets) xs[t.qubit].swap(zs[t.qubit]);
    } else if (n == "CX" || n == "CNOT" || n == "ZCX") {
      if (o.targets.size() % 2) throw std::invalid_argument("CX needs qubit pairs");
      for (size_t i = o.targets.size(); i >= 2; i -= 2) {
        uint32_t c = o.targets[i - 2].qubit, t = o.targets[i - 1].qubit;
        xor_into(xs[c], xs[t]);  // X on the control spreads to the target
        xor_into(zs[t], zs[c]);  // Z on the target spreads to the control
      }
    } else if (n == "CZ" || n == "ZCZ") {
      if (o.targets.size() % 2) throw std::invalid_argument("CZ needs qubit pairs");
      for (size_t i = o.targets.size(); i >= 2; i -= 2) {
        uint32_t a = o.targets[i - 2].qubit, b = o.targets[i - 1].qubit;
        SymSet za = zs[a];
        xor_into(xs[a], zs[b]);
        xor_into(xs[b], za);
      }
    } else if (n == "X_ERROR") {
      for (const Target& t : o.targets) emit(xs[t.qubit], arg);
    } else if (n == "Z_ERROR") {
      for (const Target& t : o.targets) emit(zs[t.qubit], arg);
    } else if (n == "Y_ERROR") {
      for (const Target& t : o.targets) emit(xor_of(xs[t.qubit], zs[t.qubit]), arg);
    } else if (n == "DEPOLARIZE1") {
      double q = 0.5 * (1 - std::sqrt(1 - 4 * arg / 3));
      for (size_t i = o.targets.size(); i-- > 0;) {
        uint32_t t = o.targets[i].qubit;
        emit(xs[t], q);
        emit(zs[t], q);
        emit(xor_of(xs[t], zs[t]), q);
      }
    } else if (n == "DEPOLARIZE2") {
      if (o.targets.size() % 2) throw std::invalid_argument("DEPOLARIZE2 needs qubit pairs");
      double q = 0.5 * (1 - std::pow(1 - 16 * arg / 15, 0.125));
      for (size_t i = o.targets.size(); i >= 2; i -= 2) {
        uint32_t a = o.targets[i - 2].qubit, b = o.targets[i - 1].qubit;
        SymSet pa[4] = {SymSet{}, xs[a], xor_of(xs[a], zs[a]), zs[a]};
        SymSet pb[4] = {SymSet{}, xs[b], xor_of(xs[b], zs[b]), zs[b]};
        for (int u = 0; u < 4; ++u)
          for (int v = 0; v < 4; ++v)
            if (u || v) emit(xor_of(pa[u], pb[v]), q);
      }
    } else if (n == "TICK" || n == "QUBIT_COORDS" || n == "SHIFT_COORDS" || n == "DETECTOR" ||
               n == "OBSERVABLE_INCLUDE" || n == "I") {
    } else {
      throw std::invalid_argument("circuit instruction '" + n + "' is not supported by this analyzer");
    }
  }

  for (uint32_t q = 0; q < num_qubits; ++q) require_deterministic(zs[q], "start of the circuit (|0>)");

  std::string text;
  for (size_t k = first_seen.size(); k-- > 0;) {
    const SymSet& sym = first_seen[k];
    text += "error(" + format_double(acc[sym]) + ")";
    for (int32_t id : sym) text += id >= kObsBase ? " L" + std::to_string(id - kObsBase) : " D" + std::to_string(id);
    text += "\n";
  }
  for (size_t d = 0; d < num_det; ++d) {
    text += "detector";
    if (!det_coords[d].empty()) {
      text += "(";
      for (size_t i = 0; i < det_coords[d].size(); ++i) text += (i ? "," : "") + format_double(det_coords[d][i]);
      text += ")";
    }
    text += " D" + std::to_string(d) + "\n";
  }
  return text;
}

// ---- sampler ------------------------------------------------------------------------------------

namespace {
struct Xoshiro {
  uint64_t s[4];
  explicit Xoshiro(uint64_t seed) {
    for (auto& w : s) {  // splitmix64
      seed += 0x9E3779B97F4A7C15ull;
      uint64_t z = seed;
      z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
      z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
      w = z ^ (z >> 31);
    }
  }
  static uint64_t rotl(uint64_t x, int k) { return (x << k) | (x >> (64 - k)); }
  uint64_t next() {
    uint64_t r = rotl(s[1] * 5, 7) * 9, t = s[1] << 17;
    s[2] ^= s[0];
    s[3] ^= s[1];
    s[1] ^= s[2];
    s[0] ^= s[3];
    s[2] ^= t;
    s[3] = rotl(s[3], 45);
    return r;
  }
  double unit_open() { return ((double)(next() >> 11) + 0.5) * (1.0 / 9007199254740992.0); }  // (0,1)
};
}  // namespace

void sample_dem_b8(const DemModel& flat, uint64_t seed, uint64_t shots, uint8_t* dets, uint64_t det_stride,
                   uint8_t* obs, uint64_t obs_stride) {
  Xoshiro rng(seed);
  for (const DemInstruction& inst : flat.instructions) {
    if (inst.kind != DemKind::Error) continue;
    double p = inst.args[0];
    if (!(p > 0)) continue;
    // Shots hit by this mechanism: geometric gaps between consecutive hits.
    double log_q = p < 1 ? std::log1p(-p) : 0;
    uint64_t s = 0;
    while (true) {
      if (p < 1) {
        double gap = std::floor(std::log(rng.unit_open()) / log_q);
        if (gap >= (double)(shots - s)) break;
        s += (uint64_t)gap;
      }
      if (s >= shots) break;
      for (const DemTarget& t : inst.targets) {
        if (t.is_det()) {
          dets[s * det_stride + (t.id() >> 3)] ^= (uint8_t)(1u << (t.id() & 7));
        } else if (t.is_obs() && obs) {
          obs[s * obs_stride + (t.id() >> 3)] ^= (uint8_t)(1u << (t.id() & 7));
        }
      }
      ++s;
    }
  }
}

}  // namespace tsr
