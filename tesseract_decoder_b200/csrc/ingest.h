// DEM ingestion: text -> compiled error list -> search tables (host side).
//
// Follows the reference's constructor pipeline (tesseract.cc:156-205, common.cc:52-239,
// utils.cc:253-261) for everything that fixes a number or an order — error canonicalisation,
// merge_weights, the probability round trip, index maps, the std::sort of detector rows — and then
// lays the result out for the GPU search instead of the reference's vector<vector<int>>:
//
//   * CSR detector->error rows in the host-sorted order, plus "row records": for every row entry
//     the error's cost, degree and the (detector, rank-in-that-detector's-row) pairs of the error's
//     OTHER detectors, so that a row scan needs no per-error gather (kernel: search.cu).
//   * CSR error->detectors with the matching ranks (ranks[e][j] = position of e in row edets[e][j]).
//   * The D x D detector adjacency bit matrix that replaces the reference's eneighbors lists
//     (tesseract.cc:229-254): Adj[a][b] = 1 iff some compiled error contains both a and b.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include "dem.h"

namespace tsr {

constexpr uint64_t kNoError = ~uint64_t{0};  // "removed" marker in dem_error_to_error

struct CompiledError {
  double cost = 0;              // -log(p/(1-p)), common.cc:85
  std::vector<int> detectors;   // ascending, XOR-canonicalised (common.cc:63-84)
  std::vector<int> observables; // ascending
  double probability() const;   // 1/(1+exp(cost)), common.cc:96-98
};

double merge_weights(double a, double b);  // common.cc:118-123

// 16-byte row-record plane A; planes B and C are 4 x RowSlot each.
struct RowSlot {
  uint16_t det;   // detector id, or D (the always-quiet dummy detector) when unused
  uint16_t rank;  // position of the error in that detector's row, 0xFFFF when unused
};
struct RowRecordA {
  double cost;
  uint16_t size;   // number of detectors of the error (|edets[e]|)
  uint16_t flags;  // bit0: degree > kMaxInlineDegree, take the CSR slow path
  RowSlot other0;
};
static_assert(sizeof(RowRecordA) == 16, "row record plane A must be 16 bytes");
constexpr int kMaxInlineDegree = 10;  // 1 (own detector) + 1 (plane A) + 4 (plane B) + 4 (plane C)

struct DemTables {
  uint64_t num_detectors = 0, num_observables = 0, num_errors = 0, num_dem_errors = 0;
  DemModel compiled_dem;  // the DEM after merge / zero-strip (what config.dem holds in the reference)
  std::vector<CompiledError> errors;
  std::vector<uint64_t> dem_error_to_error;  // kNoError for removed errors
  std::vector<uint64_t> error_to_dem_error;

  // CSR detector -> errors (rows sorted by cost/degree with std::sort, tesseract.cc:223-227).
  std::vector<uint32_t> row_off;  // [D+1]
  std::vector<int32_t> row_err;   // [nnz]
  std::vector<RowRecordA> rec_a;  // [nnz]
  std::vector<RowSlot> rec_b;     // [nnz*4]
  std::vector<RowSlot> rec_c;     // [nnz*4]
  uint32_t max_row_len = 0;
  uint32_t max_degree = 0;

  // CSR error -> detectors (ascending) with ranks, costs, observables.
  std::vector<uint32_t> e_off;   // [E+1]
  std::vector<int32_t> e_det;    // [nnz]
  std::vector<int32_t> e_rank;   // [nnz]
  std::vector<double> e_cost;    // [E]
  std::vector<uint32_t> eo_off;  // [E+1]
  std::vector<int32_t> eo_obs;   // [sum |observables|]

  // Detector adjacency bit matrix, row-major, adj_words 32-bit words per row.
  uint32_t adj_words = 0;
  std::vector<uint32_t> adj;  // [D * adj_words]

  // Reference-shaped views, for parity checks and the eneighbors getter.
  std::vector<std::vector<int>> eneighbors() const;
};

// Throws std::invalid_argument for the same inputs the reference rejects.
DemTables ingest(const DemModel& flat_dem, bool merge_errors);

// The ingestion stages on their own (the reference exposes them as tesseract_decoder.common / .utils functions).
// common.cc:139-190: errors with identical symptoms merged (merge_weights), first-seen order; DETECTOR and
// LOGICAL_OBSERVABLE instructions keep their order and come first.  error_index_map[i] = merged index of input error i.
DemModel merge_indistinguishable_errors(const DemModel& flat_dem, std::vector<uint64_t>& error_index_map);
// common.cc:192-218: errors whose probability is not > 0 dropped; error_index_map[i] = output index or kNoError.
DemModel remove_zero_probability_errors(const DemModel& flat_dem, std::vector<uint64_t>& error_index_map);
// utils.cc:253-261: the error instructions with probability > 0, canonicalised (common.cc:52-88).
std::vector<CompiledError> get_errors_from_dem(const DemModel& flat_dem);

enum class DetOrder { DetBFS = 0, DetIndex = 1, DetCoordinate = 2 };
// utils.cc:192-205 (std::mt19937_64 + libstdc++ distributions, same draws as the reference).
std::vector<std::vector<uint64_t>> build_det_orders(const DemModel& dem, uint64_t num_det_orders, DetOrder method,
                                                    uint64_t seed);
std::vector<std::vector<uint64_t>> build_detector_graph(const DemModel& dem);  // utils.cc:60-87
std::vector<std::vector<double>> get_detector_coords(const DemModel& dem);     // utils.cc:32-58

}  // namespace tsr
