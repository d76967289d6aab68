// Detector-error-model text: parse, flatten, count, print.
//
// Stands where the reference leans on stim::DetectorErrorModel: the reference's
// Python boundary hands the decoder `str(dem)` and re-parses it (stim_utils.pybind.h:22-26), so DEM
// *text* is the hand-off format this layer consumes.  A DemModel is always stored flattened: repeat
// blocks are unrolled and shift_detectors applied while parsing.
#pragma once
#include <cstdint>
#include <string>
#include <string_view>
#include <vector>

namespace tsr {

enum class DemKind : uint8_t { Error, Detector, LogicalObservable };

struct DemTarget {
  // bit 63 set: observable id; all ones: the '^' separator; otherwise an absolute detector id.
  uint64_t raw = 0;
  static constexpr uint64_t kObsBit = uint64_t{1} << 63;
  static constexpr uint64_t kSeparator = ~uint64_t{0};
  static DemTarget det(uint64_t id) { return {id}; }
  static DemTarget obs(uint64_t id) { return {id | kObsBit}; }
  static DemTarget sep() { return {kSeparator}; }
  bool is_sep() const { return raw == kSeparator; }
  bool is_obs() const { return !is_sep() && (raw & kObsBit); }
  bool is_det() const { return !is_sep() && !(raw & kObsBit); }
  uint64_t id() const { return raw & ~kObsBit; }
};

struct DemInstruction {
  DemKind kind = DemKind::Error;
  std::vector<double> args;  // error: {probability}; detector: coordinates (shift applied)
  std::vector<DemTarget> targets;
  std::string tag;
  std::string str() const;
};

struct DemModel {
  std::vector<DemInstruction> instructions;

  // Throws std::invalid_argument on malformed text or probabilities outside [0,1].
  static DemModel parse(std::string_view text);

  uint64_t count_errors() const;
  uint64_t count_detectors() const;    // 1 + largest detector id mentioned anywhere
  uint64_t count_observables() const;  // 1 + largest observable id mentioned anywhere
  std::string str() const;

  void append_error(double p, const std::vector<DemTarget>& targets, const std::string& tag = "");
};

// Shortest decimal text that parses back to exactly `v`.
std::string format_double(double v);

}  // namespace tsr
