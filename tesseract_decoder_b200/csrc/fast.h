// "Rank" form of the detector heuristic, and the certificate that makes it exact.
//
// The reference's get_detcost (tesseract.cc:128-154) walks a detector's row keeping (min_cost, min_count)
// and decides everything with one predicate on (cost, small integer) pairs:
//     P(a, b)  :=  fl(cost_a * n_b) < fl(cost_b * n_a)
// (update: a = (cost_i, count_i), b = current minimum; early exit: !P((cost_i, size_i), b)).  A DEM after
// merging has only a few dozen distinct costs, so the pairs ("classes") that can ever meet number a few
// hundred.  build_fast_tables() evaluates P on ALL class pairs with the same double-precision products and
// certifies, for this DEM, that
//   (1) P is a strict weak order, i.e. there is an integer rank with  P(a, b) <=> rank(a) < rank(b)
//       (near-equal costs that differ in the last bit simply share a rank, exactly as P treats them);
//   (2) along every detector row (in the reference's std::sort order) rank(cost_i, size_i) never decreases,
//       so the reference's early exit can never hide an entry that would have replaced the minimum.
// Under (1)+(2) the sequential scan returns the FIRST unblocked row entry of minimal rank(cost_i, count_i),
// whatever the evaluation order, and its value is the host-computed quotient fl(cost/count) of that
// entry's class.  The search kernel (search_fast.cu) then needs no floating point inside a row scan: it
// takes an integer minimum of (rank << 12 | position) and looks the quotient up.  The add/sub chain of the
// child cost stays in the reference's order.  DEMs that fail the certificate (or the packing limits below)
// are decoded by the generic kernel (search.cu), which replays the scan event by event.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include "ingest.h"

namespace tsr {

constexpr uint32_t kFastMaxCosts = 128;      // cost id is 7 bits of the record header
constexpr uint32_t kFastMaxDegree = 11;      // own detector + 5 slots (plane A) + 5 slots (plane B)
constexpr uint32_t kFastMaxDetectors = 16383;  // 14-bit detector field (D itself is the dummy detector)
constexpr uint32_t kFastMaxRowLen = 1024;    // 10-bit rank field; 12-bit position inside a key
constexpr uint32_t kFastSlotDummyRank = 0x3FF;

struct FastTables {
  bool certified = false;
  std::string why_not;  // set when !certified
  // Row records, one 16-byte plane A per row entry (same indexing as DemTables::row_err) and a plane B
  // that is only read for entries whose header has bit 7 set.
  //   plane A: byte 0 = cost_id | (has_plane_b << 7); bytes 1..15 = slots 0..4
  //   plane B: bytes 0..14 = slots 5..9
  //   slot (24 bits, little endian) = detector (14 bits) | rank of the error in that detector's row << 14;
  //   unused slots name the dummy detector D with rank 0x3FF.
  std::vector<uint8_t> rec_a, rec_b;  // [nnz * 16]
  uint32_t a_slots = 0;               // slots of plane A any entry uses: min(5, max_degree - 1)
  bool any_plane_b = false;
  // Classes: cls = cost_id * nstride + n, n = 1..size.
  uint32_t n_costs = 0, nstride = 0;
  std::vector<uint16_t> rank;      // [n_costs * nstride] rank of class; 0xFFFF for unused classes
  std::vector<double> quotient;    // [n_costs * nstride] fl(cost / n)
  std::vector<uint8_t> cost_id;    // [E]
};

FastTables build_fast_tables(const DemTables& T);

}  // namespace tsr
