// --sparsify-errors on the device (reference: tesseract.cc:256-292 set-up, :673-737 build_sparse_d2e).
//
// The reference rebuilds, per shot, filtered copies of all detector rows holding only the "active" errors: the
// mandatory ones (degree <= sparsify_base_degree) plus the `sparsify_reactivate_limit` best optional errors
// (base < degree <= max) that touch a fired detector, ranked by (overlap with the syndrome descending, degree
// ascending, cost ascending, error index ascending).  In this decoder's state representation an inactive error is
// simply one that no row scan may see: it is skipped by get_detcost and never expanded, and block thresholds keep their
// meaning because they refer to positions in the FULL row (filtering preserves row order, and whether an inactive
// entry counts as blocked is unobservable).  So the per-shot graph is one bit per row entry:
//
//     act[slot][d][w]   bit i of word w = entry 32 w + i of row[d] is active for the shot this slot is searching
//
// kept per search slot in global memory (L2-resident), equal to the static mask of mandatory errors except while a
// search runs.  At the start of a search the CTA (fast path) or warp (generic path) selects the reactivated errors:
//
//   * candidates are enumerated from the rows of the FIRED detectors (an optional error with overlap > 0 appears in
//     the row of each of its fired detectors; it is taken from its lowest one), never by a scan over all errors;
//   * the key (63 - overlap : 6 | degree : 6 | cost rank : 24 | error index : 28) orders candidates exactly like the
//     reference's comparator (cost ranks are dense ranks of the distinct double costs, so a.cost != b.cost <=> ranks
//     differ) and is unique, so "the first `limit` after std::sort" is "the `limit` smallest keys" — an MSB-first
//     radix select over re-enumerations with a 256-bin shared-memory histogram, no sort and no candidate buffer;
//   * the chosen errors' bits are OR-ed into act[slot]; when the search ends the touched words are restored.
#pragma once
#include <cstdint>

#include "search.h"

namespace tsr {
namespace sparsify {

constexpr uint32_t kHistBins = 256;
// Shared scratch: kHistBins histogram words + 8 control words.
constexpr uint32_t kScratchWords = kHistBins + 8;

__device__ __forceinline__ bool bit_of(const uint32_t* bits, uint32_t i) {
  return (bits[i >> 5] >> (i & 31)) & 1u;
}

// Calls f(key, error, first CSR index of the error, degree) for every reactivation candidate of the syndrome `fired`
// ([nwords] bitset in shared memory): warps take the bitset words round-robin, lanes stride a fired detector's row.
template <typename F>
__device__ __forceinline__ void for_each_candidate(const DeviceTables& t, const uint32_t* fired, uint32_t warp, uint32_t nwarps,
                                                   uint32_t lane, F&& f) {
  for (uint32_t w = warp; w < t.nwords; w += nwarps) {
    for (uint32_t m = fired[w]; m; m &= m - 1) {
      const uint32_t d = w * 32 + (uint32_t)__ffs((int)m) - 1;
      const uint32_t roff = __ldg(t.row_off + d), rlen = __ldg(t.row_off + d + 1) - roff;
      const uint32_t* opt = t.sp_opt + (size_t)d * t.sp_rw;
      for (uint32_t idx = lane; idx < rlen; idx += 32) {
        if (!((__ldg(opt + (idx >> 5)) >> (idx & 31)) & 1u)) continue;
        const uint32_t e = (uint32_t)__ldg(t.row_err + roff + idx);
        const uint32_t eb = __ldg(t.e_off + e), deg = __ldg(t.e_off + e + 1) - eb;
        uint32_t overlap = 0, first = 0xFFFFFFFFu;
        for (uint32_t j = 0; j < deg; ++j) {
          const uint32_t dj = (uint32_t)__ldg(t.e_det + eb + j);
          if (bit_of(fired, dj)) {
            if (first == 0xFFFFFFFFu) first = dj;  // detectors ascend
            ++overlap;
          }
        }
        if (first != d) continue;  // this error is taken from the row of its lowest fired detector
        const unsigned long long key = ((unsigned long long)(63u - overlap) << 58) | ((unsigned long long)deg << 52) |
                                       ((unsigned long long)__ldg(t.sp_crank + e) << 28) | (unsigned long long)e;
        f(key, e, eb, deg);
      }
    }
  }
}

// Builds act (this slot's [D][sp_rw] masks, equal to sp_mand on entry) for the syndrome `fired`.  All `nwarps` warps of
// the search call it together; `sync` is the barrier among them (__syncthreads or __syncwarp); scratch is shared memory
// [kScratchWords] visible to all of them.
template <typename Sync>
__device__ __forceinline__ void select(const DeviceTables& t, const uint32_t* fired, uint32_t* act, uint32_t* scratch, uint32_t warp,
                                       uint32_t nwarps, uint32_t lane, Sync&& sync) {
  uint32_t* hist = scratch;
  uint32_t* ctl = scratch + kHistBins;  // [0] candidates, [1..2] prefix, [3] need, [4] done
  const uint32_t limit = t.sp_limit;
  if (limit == 0) return;
  if (warp == 0 && lane == 0) ctl[0] = 0;
  sync();
  for_each_candidate(t, fired, warp, nwarps, lane, [&](unsigned long long, uint32_t, uint32_t, uint32_t) { atomicAdd(&ctl[0], 1u); });
  sync();
  const uint32_t n_cand = ctl[0];
  if (n_cand == 0) return;
  unsigned long long t_max = ~0ull;  // every candidate with key <= t_max is reactivated
  if (n_cand > limit) {
    unsigned long long prefix = 0;
    uint32_t need = limit;
    for (uint32_t pass = 0; pass < 8; ++pass) {
      const uint32_t shift = 56 - 8 * pass;
      for (uint32_t i = warp * 32 + lane; i < kHistBins; i += nwarps * 32) hist[i] = 0;
      sync();
      for_each_candidate(t, fired, warp, nwarps, lane, [&](unsigned long long key, uint32_t, uint32_t, uint32_t) {
        if (pass == 0 || (key >> (shift + 8)) == prefix) atomicAdd(&hist[(uint32_t)(key >> shift) & 255u], 1u);
      });
      sync();
      if (warp == 0) {
        // bucket b with cum(b-1) < need <= cum(b): lane l owns bins 8l .. 8l+7
        uint32_t mine = 0;
        for (uint32_t k = 0; k < 8; ++k) mine += hist[lane * 8 + k];
        uint32_t incl = mine;
        for (int o = 1; o < 32; o <<= 1) {
          const uint32_t v = __shfl_up_sync(0xFFFFFFFFu, incl, o);
          if ((int)lane >= o) incl += v;
        }
        const unsigned hit = __ballot_sync(0xFFFFFFFFu, incl >= need);
        const uint32_t l = (uint32_t)__ffs((int)hit) - 1;  // need <= candidates with this prefix, so some lane hits
        if (lane == l) {
          uint32_t cum = incl - mine, b = lane * 8;
          while (cum + hist[b] < need) cum += hist[b++];
          const unsigned long long np = (prefix << 8) | b;
          ctl[1] = (uint32_t)np;
          ctl[2] = (uint32_t)(np >> 32);
          ctl[3] = need - cum;
          ctl[4] = (need - cum == hist[b]) ? 1u : 0u;
        }
      }
      sync();
      prefix = (unsigned long long)ctl[1] | ((unsigned long long)ctl[2] << 32);
      need = ctl[3];
      const bool done = ctl[4] != 0;
      sync();
      if (done) {
        t_max = ((prefix + 1) << shift) - 1;  // wraps to all ones when the prefix is all ones
        break;
      }
    }
  }
  for_each_candidate(t, fired, warp, nwarps, lane, [&](unsigned long long key, uint32_t, uint32_t eb, uint32_t deg) {
    if (key > t_max) return;
    for (uint32_t j = 0; j < deg; ++j) {
      const uint32_t det = (uint32_t)__ldg(t.e_det + eb + j), r = (uint32_t)__ldg(t.e_rank + eb + j);
      atomicOr(act + (size_t)det * t.sp_rw + (r >> 5), 1u << (r & 31));
    }
  });
  sync();
}

// Restores act to the mandatory masks after a search (same enumeration; every candidate's words are reset).
template <typename Sync>
__device__ __forceinline__ void restore(const DeviceTables& t, const uint32_t* fired, uint32_t* act, uint32_t warp, uint32_t nwarps,
                                        uint32_t lane, Sync&& sync) {
  if (t.sp_limit == 0) return;
  for_each_candidate(t, fired, warp, nwarps, lane, [&](unsigned long long, uint32_t, uint32_t eb, uint32_t deg) {
    for (uint32_t j = 0; j < deg; ++j) {
      const uint32_t det = (uint32_t)__ldg(t.e_det + eb + j), r = (uint32_t)__ldg(t.e_rank + eb + j);
      const size_t at = (size_t)det * t.sp_rw + (r >> 5);
      __stcg(act + at, __ldg(t.sp_mand + at));
    }
  });
  sync();
}

}  // namespace sparsify
}  // namespace tsr
