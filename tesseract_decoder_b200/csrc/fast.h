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

// Bit-sliced row tables (search_fast.cu, "bits" scan): the same rank minimum, evaluated 32 row entries per
// lane instead of one.  For a detector x with row length L (ceil(L/32) words):
//   * for each detector y adjacent to x (the set bits of adj[x], in ascending order; index j = number of set
//     bits of adj[x] below y): the L-bit mask of row[x] entries whose error also contains y.  The fired-count
//     of every row entry is 1 + the bit-sliced sum of the masks of the FIRED neighbours;
//   * for each pair (x, y) and each 32-entry word of row[x]: the entries of that word containing y as
//     (rank in row[y], bit) sorted by rank — the entries a block threshold on y removes are a prefix of that list;
//   * for each 32-entry word of row[x]: the distinct costs among its entries (ascending) as `ncls` slots of
//     (mask of the word's entries with that cost, cost id); unused slots are (0, 0).  A word of a BB [[144,12,12]]
//     row holds 3.4 cost classes on average (at most 9) while the row has 9-12: a lane scans its own word's classes.
struct BitRowMeta {   // 32 bytes, one per detector
  uint32_t mask_off;  // word offset of the neighbour masks [j][word] in `masks`
  uint32_t cmask_off; // word offset (a multiple of 4) of the class slots [word][ncls] x {mask, cost id} in `masks`
  uint32_t pair_base; // pair index of (x, first neighbour); pair lists are pl[pl_off[p] .. pl_off[p+1])
  uint32_t reserved;
  uint32_t len, nw, ncls, pad;  // ncls: class slots per word (even)
};
struct BitTables {
  bool ok = false;
  std::string why_not;
  uint32_t group = 0;     // lanes per target: smallest power of two >= max words per row
  uint32_t max_nbrs = 0;  // largest number of set bits in an adjacency row
  std::vector<BitRowMeta> rowmeta;  // [D]
  std::vector<uint16_t> adjpre;     // [D * adj_words] set bits of adj[x] before word w
  std::vector<uint32_t> masks;
  std::vector<uint32_t> pl_off;     // [masks.size() + 1], indexed like the neighbour masks: per (pair, row word)
  std::vector<uint32_t> pl;         // rank in the neighbour's row << 16 | bit within the row word
};
BitTables build_bit_tables(const DemTables& T, const FastTables& F);

// Host cross-check of the three formulations of get_detcost on `trials` random search states (random fired
// sets and block thresholds): (1) the reference's floating-point loop with its early exit (tesseract.cc:128-154)
// on the plain tables, (2) the rank minimum over packed row records (entry-per-lane kernel), (3) the bit-sliced
// evaluation over BitTables (masks, pair lists, cost classes) exactly as search_fast.cu performs it.  Returns
// the number of (state, detector) evaluations whose three values are not bit-identical doubles; `checked`
// receives how many were compared.  B may be !ok (then (3) is skipped).
uint64_t cross_check_detcost(const DemTables& T, const FastTables& F, const BitTables& B, uint64_t trials, uint64_t seed,
                             uint64_t* checked);

}  // namespace tsr
