// See fast.h.
#include "fast.h"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <map>

namespace tsr {
namespace {

struct Cls {
  uint32_t cost_id, n;
  double cost;
};

// The reference's predicate, with its arithmetic: double * (double)integer on both sides, each product
// rounded to double, then compared (tesseract.cc:139-140, 146).
inline bool P(const Cls& a, const Cls& b) {
  volatile double lhs = a.cost * (double)b.n;
  volatile double rhs = b.cost * (double)a.n;
  return lhs < rhs;
}

// Exact order of cost/n: a 53-bit significand times an integer <= 11 is exact in x87 extended precision.
inline bool exact_less(const Cls& a, const Cls& b) {
  static_assert(sizeof(long double) >= 10, "needs a 64-bit significand");
  return (long double)a.cost * (long double)b.n < (long double)b.cost * (long double)a.n;
}

void put_slot(uint8_t* plane, int byte_off, uint32_t det, uint32_t rank) {
  const uint32_t v = (det & 0x3FFF) | (rank << 14);
  plane[byte_off] = (uint8_t)v;
  plane[byte_off + 1] = (uint8_t)(v >> 8);
  plane[byte_off + 2] = (uint8_t)(v >> 16);
}

}  // namespace

FastTables build_fast_tables(const DemTables& T) {
  FastTables F;
  const uint64_t D = T.num_detectors, E = T.num_errors;
  auto fail = [&](std::string why) {
    F.certified = false;
    F.why_not = std::move(why);
    return F;
  };
  if (D > kFastMaxDetectors) return fail("more than 16383 detectors");
  if (T.max_degree > kFastMaxDegree) return fail("an error touches more than 11 detectors");
  if (T.max_row_len > kFastMaxRowLen) return fail("a detector row is longer than 1024 errors");

  std::map<double, uint32_t> ids;
  for (uint64_t e = 0; e < E; ++e) {
    const double c = T.e_cost[e];
    if (!(c > 0) || !std::isfinite(c)) return fail("an error has a non-positive or non-finite cost (p >= 0.5)");
    ids.emplace(c, 0);
  }
  if (ids.size() > kFastMaxCosts) return fail("more than 128 distinct error costs");
  uint32_t next = 0;
  for (auto& kv : ids) kv.second = next++;
  F.n_costs = (uint32_t)ids.size();
  F.nstride = std::max<uint32_t>(T.max_degree, 1) + 1;
  F.cost_id.resize(E);
  std::vector<uint32_t> max_size(F.n_costs, 0);
  std::vector<double> cost_of(F.n_costs, 0);
  for (uint64_t e = 0; e < E; ++e) {
    const uint32_t k = ids[T.e_cost[e]];
    F.cost_id[e] = (uint8_t)k;
    cost_of[k] = T.e_cost[e];
    max_size[k] = std::max<uint32_t>(max_size[k], (uint32_t)T.errors[e].detectors.size());
  }

  // (1) P is a strict weak order on the classes.
  std::vector<Cls> cls;
  for (uint32_t k = 0; k < F.n_costs; ++k)
    for (uint32_t n = 1; n <= max_size[k]; ++n) cls.push_back({k, n, cost_of[k]});
  std::stable_sort(cls.begin(), cls.end(), exact_less);
  std::vector<uint32_t> r(cls.size(), 0);
  for (size_t i = 1; i < cls.size(); ++i) r[i] = r[i - 1] + (P(cls[i - 1], cls[i]) ? 1u : 0u);
  for (size_t i = 0; i < cls.size(); ++i)
    for (size_t j = i; j < cls.size(); ++j)
      if (P(cls[i], cls[j]) != (r[i] < r[j]) || P(cls[j], cls[i]))
        return fail("the heuristic's rounded ratio comparison is not a strict weak order on this DEM's costs");
  if (!cls.empty() && r.back() >= 0xFFFF) return fail("too many distinct cost/count ratios");
  F.rank.assign((size_t)F.n_costs * F.nstride, 0xFFFF);
  F.quotient.assign((size_t)F.n_costs * F.nstride, INFINITY);
  for (size_t i = 0; i < cls.size(); ++i) {
    const size_t at = (size_t)cls[i].cost_id * F.nstride + cls[i].n;
    F.rank[at] = (uint16_t)r[i];
    volatile double q = cls[i].cost / (double)cls[i].n;  // tesseract.cc:153
    F.quotient[at] = q;
  }

  // (2) rank(cost, size) never decreases along a row.
  for (uint64_t d = 0; d < D; ++d) {
    uint32_t prev = 0;
    for (uint32_t k = T.row_off[d]; k < T.row_off[d + 1]; ++k) {
      const uint64_t e = (uint64_t)T.row_err[k];
      const uint32_t rk = F.rank[(size_t)F.cost_id[e] * F.nstride + T.errors[e].detectors.size()];
      if (rk < prev) return fail("a detector row is not sorted by the rank of cost/size");
      prev = rk;
    }
  }

  // Records.
  const size_t nnz = T.row_err.size();
  F.rec_a.assign(nnz * 16 + 16, 0);
  F.rec_b.assign(nnz * 16 + 16, 0);
  F.a_slots = std::min<uint32_t>(5, T.max_degree > 0 ? T.max_degree - 1 : 0);
  for (uint64_t d = 0; d < D; ++d) {
    for (uint32_t at = T.row_off[d]; at < T.row_off[d + 1]; ++at) {
      const uint64_t e = (uint64_t)T.row_err[at];
      const auto& dets = T.errors[e].detectors;
      uint8_t* a = &F.rec_a[(size_t)at * 16];
      uint8_t* b = &F.rec_b[(size_t)at * 16];
      for (int s = 0; s < 5; ++s) {
        put_slot(a, 1 + 3 * s, (uint32_t)D, kFastSlotDummyRank);
        put_slot(b, 3 * s, (uint32_t)D, kFastSlotDummyRank);
      }
      int slot = 0;
      for (size_t j = 0; j < dets.size(); ++j) {
        if ((uint64_t)dets[j] == d) continue;
        const uint32_t rank = (uint32_t)T.e_rank[T.e_off[e] + j];
        if (slot < 5)
          put_slot(a, 1 + 3 * slot, (uint32_t)dets[j], rank);
        else
          put_slot(b, 3 * (slot - 5), (uint32_t)dets[j], rank);
        ++slot;
      }
      const bool has_b = slot > 5;
      F.any_plane_b |= has_b;
      a[0] = (uint8_t)(F.cost_id[e] | (has_b ? 0x80 : 0));
    }
  }
  F.certified = true;
  return F;
}

BitTables build_bit_tables(const DemTables& T, const FastTables& F) {
  BitTables B;
  if (!F.certified) {
    B.why_not = F.why_not;
    return B;
  }
  const uint64_t D = T.num_detectors;
  const uint32_t aw = T.adj_words;
  B.rowmeta.resize(D);
  B.adjpre.assign((size_t)D * aw + 1, 0);
  uint32_t max_nw = 1;
  uint64_t mask_words = 0, pairs = 0;
  std::vector<std::vector<uint32_t>> nbrs(D);
  for (uint64_t x = 0; x < D; ++x) {
    uint32_t run = 0;
    for (uint32_t w = 0; w < aw; ++w) {
      B.adjpre[(size_t)x * aw + w] = (uint16_t)run;
      for (uint32_t m = T.adj[(size_t)x * aw + w]; m; m &= m - 1) nbrs[x].push_back(w * 32 + (uint32_t)__builtin_ctz(m));
      run += (uint32_t)__builtin_popcount(T.adj[(size_t)x * aw + w]);
    }
    if (run > 0xFFFF) {
      B.why_not = "a detector has more than 65535 neighbours";
      return B;
    }
    B.max_nbrs = std::max(B.max_nbrs, run);
    const uint32_t len = T.row_off[x + 1] - T.row_off[x];
    const uint32_t nw = std::max<uint32_t>((len + 31) / 32, 1);
    max_nw = std::max(max_nw, nw);
    uint32_t slots = 0;  // most distinct costs in one 32-entry word of the row
    for (uint32_t w0 = 0; w0 < len; w0 += 32) {
      std::vector<uint8_t> cls;
      for (uint32_t i = w0; i < std::min(len, w0 + 32); ++i) cls.push_back(F.cost_id[(size_t)T.row_err[T.row_off[x] + i]]);
      std::sort(cls.begin(), cls.end());
      slots = std::max<uint32_t>(slots, (uint32_t)(std::unique(cls.begin(), cls.end()) - cls.begin()));
    }
    BitRowMeta& m = B.rowmeta[x];
    m.len = len;
    m.nw = nw;
    m.ncls = std::max<uint32_t>((slots + 1) & ~1u, 2);  // even: the kernel reads two slots per 16-byte load
    m.pad = 0;
    m.reserved = 0;
    m.mask_off = (uint32_t)mask_words;
    mask_words += (uint64_t)nbrs[x].size() * nw;
    mask_words = (mask_words + 3) & ~uint64_t{3};
    m.cmask_off = (uint32_t)mask_words;
    mask_words += (uint64_t)2 * m.ncls * nw;
    m.pair_base = (uint32_t)pairs;
    pairs += nbrs[x].size();
    if (mask_words > 0x7FFFFFFFull) {
      B.why_not = "the bit-sliced row masks exceed 8 GB";
      return B;
    }
  }
  B.group = 1;
  while (B.group < max_nw) B.group *= 2;
  if (B.group > 32) {
    B.why_not = "a detector row is longer than 1024 errors";
    return B;
  }
  B.masks.assign((size_t)mask_words + 4, 0);
  std::vector<std::vector<uint32_t>> lists((size_t)pairs);
  for (uint64_t x = 0; x < D; ++x) {
    const BitRowMeta& m = B.rowmeta[x];
    for (uint32_t w0 = 0; w0 < m.len; w0 += 32) {  // this word's classes, ascending cost id == ascending cost
      std::vector<uint8_t> cls;
      for (uint32_t i = w0; i < std::min(m.len, w0 + 32); ++i) cls.push_back(F.cost_id[(size_t)T.row_err[T.row_off[x] + i]]);
      std::sort(cls.begin(), cls.end());
      cls.erase(std::unique(cls.begin(), cls.end()), cls.end());
      uint32_t* slot = B.masks.data() + m.cmask_off + (size_t)2 * m.ncls * (w0 >> 5);
      for (size_t k = 0; k < cls.size(); ++k) slot[2 * k + 1] = cls[k];
      for (uint32_t i = w0; i < std::min(m.len, w0 + 32); ++i) {
        const uint8_t cid = F.cost_id[(size_t)T.row_err[T.row_off[x] + i]];
        const size_t k = (size_t)(std::lower_bound(cls.begin(), cls.end(), cid) - cls.begin());
        slot[2 * k] |= 1u << (i & 31);
      }
    }
    for (uint32_t i = 0; i < m.len; ++i) {
      const uint64_t e = (uint64_t)T.row_err[T.row_off[x] + i];
      const auto& dets = T.errors[e].detectors;
      for (size_t q = 0; q < dets.size(); ++q) {
        const uint32_t y = (uint32_t)dets[q];
        if (y == x) continue;
        const uint32_t j = (uint32_t)(std::lower_bound(nbrs[x].begin(), nbrs[x].end(), y) - nbrs[x].begin());
        B.masks[m.mask_off + (size_t)j * m.nw + (i >> 5)] |= 1u << (i & 31);
        lists[m.pair_base + j].push_back(((uint32_t)T.e_rank[T.e_off[e] + q] << 16) | i);
      }
    }
  }
  // Pair lists, split by row word: pl_off is indexed like the neighbour masks (mask_off + j * nw + word) and bounds
  // the entries of pair (x, neighbour j) that fall into that word of row[x], sorted by rank in the neighbour's row.
  B.pl_off.assign(B.masks.size() + 1, 0);
  size_t total = 0;
  for (const auto& l : lists) total += l.size();
  B.pl.assign(total + 4, 0);
  uint32_t run = 0;
  size_t next_index = 0;  // pl_off entries below this are final
  for (uint64_t x = 0; x < D; ++x) {
    const BitRowMeta& m = B.rowmeta[x];
    for (size_t j = 0; j < nbrs[x].size(); ++j) {
      std::vector<uint32_t>& l = lists[m.pair_base + j];
      std::sort(l.begin(), l.end(), [](uint32_t a, uint32_t b) {  // by word of row[x], then by rank (ranks are distinct)
        const uint32_t wa = (a & 0xFFFF) >> 5, wb = (b & 0xFFFF) >> 5;
        return wa != wb ? wa < wb : a < b;
      });
      size_t at = 0;
      for (uint32_t w = 0; w < m.nw; ++w) {
        const size_t index = (size_t)m.mask_off + j * m.nw + w;
        for (; next_index <= index; ++next_index) B.pl_off[next_index] = run;
        while (at < l.size() && ((l[at] & 0xFFFF) >> 5) == w) {
          B.pl[run++] = (l[at] & 0xFFFF0000u) | (l[at] & 31);  // rank << 16 | bit within the word
          ++at;
        }
      }
    }
  }
  for (; next_index < B.pl_off.size(); ++next_index) B.pl_off[next_index] = run;
  B.ok = true;
  return B;
}

namespace {

uint64_t mix64(uint64_t& s) {
  s += 0x9E3779B97F4A7C15ull;
  uint64_t z = s;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}

}  // namespace

uint64_t cross_check_detcost(const DemTables& T, const FastTables& F, const BitTables& B, uint64_t trials, uint64_t seed,
                             uint64_t* checked) {
  const uint64_t D = T.num_detectors;
  uint64_t bad = 0, n_checked = 0;
  if (!F.certified || D == 0) {
    if (checked) *checked = 0;
    return 0;
  }
  std::vector<uint8_t> fired(D);
  std::vector<uint32_t> thr(D);
  const uint32_t aw = T.adj_words;
  for (uint64_t trial = 0; trial < trials; ++trial) {
    // a random state: a cluster of fired detectors around a random error neighbourhood, a few thresholds
    std::fill(fired.begin(), fired.end(), 0);
    std::fill(thr.begin(), thr.end(), 0u);
    const uint64_t n_seed_errors = 1 + mix64(seed) % 6;
    for (uint64_t k = 0; k < n_seed_errors && T.num_errors; ++k)
      for (int d : T.errors[mix64(seed) % T.num_errors].detectors) fired[(size_t)d] ^= 1;
    for (uint64_t k = 0; k < 3; ++k) fired[mix64(seed) % D] = 1;
    for (uint64_t k = 0; k < 1 + mix64(seed) % 4; ++k) {
      const uint64_t d = mix64(seed) % D;
      const uint32_t len = T.row_off[d + 1] - T.row_off[d];
      thr[d] = len ? (uint32_t)(mix64(seed) % (len + 1)) : 0;
      if (mix64(seed) % 8 == 0) thr[d] = 0x7FF;  // "block the whole row" (at-most-two)
    }
    for (uint64_t x = 0; x < D; ++x) {
      if (!fired[x]) continue;
      const uint32_t roff = T.row_off[x], len = T.row_off[x + 1] - roff;
      auto count_of = [&](uint64_t e) {
        uint32_t c = 0;
        for (int d : T.errors[e].detectors) c += fired[(size_t)d];
        return c;
      };
      auto blocked_of = [&](uint64_t e) {
        const auto& dets = T.errors[e].detectors;
        for (size_t j = 0; j < dets.size(); ++j)
          if ((uint32_t)T.e_rank[T.e_off[e] + j] < thr[(size_t)dets[j]]) return true;
        return false;
      };
      // (1) the reference loop
      double min_cost = INFINITY;
      uint32_t min_count = 0xFFFFFFFFu;
      for (uint32_t i = 0; i < len; ++i) {
        const uint64_t e = (uint64_t)T.row_err[roff + i];
        volatile double lhs = T.e_cost[e] * (double)min_count, rhs = min_cost * (double)T.errors[e].detectors.size();
        if (lhs >= rhs) break;
        if (!blocked_of(e)) {
          const uint32_t cnt = count_of(e);
          volatile double l2 = T.e_cost[e] * (double)min_count, r2 = min_cost * (double)cnt;
          if (l2 < r2) {
            min_cost = T.e_cost[e];
            min_count = cnt;
          }
        }
      }
      volatile double ref = min_cost / (double)min_count;
      // (2) rank minimum over the packed records
      uint32_t best = 0xFFFFFFFFu, best_cls = 0;
      for (uint32_t i = 0; i < len; ++i) {
        const uint8_t* a = &F.rec_a[(size_t)(roff + i) * 16];
        const uint8_t* b = &F.rec_b[(size_t)(roff + i) * 16];
        uint32_t cnt = 1;
        bool blk = i < (thr[x] & 0x7FF);
        auto slot = [&](const uint8_t* q) {
          const uint32_t v = (uint32_t)q[0] | ((uint32_t)q[1] << 8) | ((uint32_t)q[2] << 16);
          const uint32_t det = v & 0x3FFF, rk = v >> 14;
          if (det < D) {
            cnt += fired[det];
            blk |= rk < thr[det];
          }
        };
        for (int k = 0; k < 5; ++k) slot(a + 1 + 3 * k);
        if (a[0] & 0x80)
          for (int k = 0; k < 5; ++k) slot(b + 3 * k);
        if (blk) continue;
        const uint32_t cls = (uint32_t)(a[0] & 0x7F) * F.nstride + cnt;
        const uint32_t key = ((uint32_t)F.rank[cls] << 12) | i;
        if (key < best) {
          best = key;
          best_cls = cls;
        }
      }
      const double lanes = best == 0xFFFFFFFFu ? INFINITY : F.quotient[best_cls];
      // (3) bit-sliced
      double bits = lanes;
      if (B.ok) {
        const BitRowMeta& m = B.rowmeta[x];
        uint32_t bbest = 0xFFFFFFFFu, bcls = 0;
        for (uint32_t wl = 0; wl < m.nw; ++wl) {
          uint32_t c0 = 0, c1 = 0, c2 = 0, c3 = 0;
          const uint32_t lo = wl * 32;
          const uint32_t valid = m.len - lo >= 32 ? 0xFFFFFFFFu : ((1u << (m.len - lo)) - 1u);
          const uint32_t t0 = thr[x];
          uint32_t Bm = t0 <= lo ? 0u : (t0 - lo >= 32 ? 0xFFFFFFFFu : ((1u << (t0 - lo)) - 1u));
          for (uint32_t w = 0; w < aw; ++w) {
            const uint32_t arow = T.adj[(size_t)x * aw + w];
            for (uint32_t mm = arow; mm; mm &= mm - 1) {
              const uint32_t bit = (uint32_t)__builtin_ctz(mm), y = w * 32 + bit;
              if (y == x) continue;
              const uint32_t j = B.adjpre[(size_t)x * aw + w] + (uint32_t)__builtin_popcount(arow & ((1u << bit) - 1u));
              if (fired[y]) {
                const uint32_t msk = B.masks[m.mask_off + (size_t)j * m.nw + wl];
                const uint32_t a0 = c0 & msk;
                c0 ^= msk;
                const uint32_t a1 = c1 & a0;
                c1 ^= a0;
                const uint32_t a2 = c2 & a1;
                c2 ^= a1;
                c3 ^= a2;
              }
              if (thr[y]) {
                const size_t idx = m.mask_off + (size_t)j * m.nw + wl;
                for (uint32_t at = B.pl_off[idx]; at < B.pl_off[idx + 1]; ++at) {
                  if ((B.pl[at] >> 16) >= thr[y]) break;
                  Bm |= 1u << (B.pl[at] & 31);
                }
              }
            }
          }
          const uint32_t U = valid & ~Bm;
          if (!U) continue;
          for (uint32_t kc = 0; kc < m.ncls; ++kc) {
            const uint32_t* slot = B.masks.data() + m.cmask_off + ((size_t)wl * m.ncls + kc) * 2;
            uint32_t cm = slot[0] & U;
            if (!cm) continue;
            uint32_t q;
            q = cm & c3;
            cm = q ? q : cm;
            q = cm & c2;
            cm = q ? q : cm;
            q = cm & c1;
            cm = q ? q : cm;
            q = cm & c0;
            cm = q ? q : cm;
            const uint32_t v = ((cm & c3) ? 8u : 0u) | ((cm & c2) ? 4u : 0u) | ((cm & c1) ? 2u : 0u) | ((cm & c0) ? 1u : 0u);
            const uint32_t cls = slot[1] * F.nstride + v + 1;
            const uint32_t key = ((uint32_t)F.rank[cls] << 12) | (lo + (uint32_t)__builtin_ctz(cm));
            if (key < bbest) {
              bbest = key;
              bcls = cls;
            }
          }
        }
        bits = bbest == 0xFFFFFFFFu ? INFINITY : F.quotient[bcls];
      }
      ++n_checked;
      const double r = ref;
      if (std::memcmp(&r, &lanes, 8) != 0 || std::memcmp(&r, &bits, 8) != 0) ++bad;
    }
  }
  if (checked) *checked = n_checked;
  return bad;
}

}  // namespace tsr
