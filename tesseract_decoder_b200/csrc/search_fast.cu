// The A* most-likely-error search, fast path: one CTA per (shot, trial) search, W warps per CTA.
//
// Same search as search.cu (reference: tesseract.cc:380-616, 128-154, 349-369, 119-121, bits/stl_heap.h),
// reorganised around what the v0 profile showed (profiles/r01_v0_*): the warp-per-search kernel was
// latency bound (long-scoreboard stalls on L2 row records, 2 % L1 hit rate, 16 warps/SM) and lost half
// the machine to the tail of a batch.  Here
//
//   * a search owns a CTA.  The children of an expanded node are independent (each is the parent state
//     with the error's detectors toggled and a longer blocked prefix of the min-detector's row), so the
//     warps claim them one at a time and warp 0 then pushes the results in row order — which keeps the
//     reference's queue order.  A batch needs 8-16x fewer concurrent searches to fill the machine
//     (smaller tail, larger node pool per search).
//   * get_detcost is integer-only.  Under the host-checked certificate (fast.h) it is "the first unblocked
//     row entry of minimal rank(cost, count)", and the value is the host-computed quotient fl(cost/count).
//     Two scans compute that minimum:
//       - entry per lane (scan_row): lanes stride over 16-byte row records (5 neighbours, +16 bytes with 5
//         more for errors of degree > 6; 24-bit (detector, rank) slots), look the neighbours up in the
//         warp's private copy of the 16-bit-per-detector state, keep min(rank << 12 | position), one
//         REDUX.MIN per row.  kInstr additionally counts the row entries the reference's early-exit loop
//         visits (counter S).
//       - bit-sliced (bs_core / bs_targets, kBits): a group of G lanes owns a target row, a lane 32 row
//         entries as bit masks; fired counts are a carry-save sum of precomputed neighbour masks, blocked
//         entries come from rank-sorted pair lists, and each cost class yields one candidate.  No private
//         state copy: a child is (fired bitset, min detector's threshold) on top of the node's state.
//   * the queue is the libstdc++ binary heap in global memory; the pushes of an expansion are staged in
//     shared memory (one tree level per sub-batch) or committed in windows of 32.
//   * a popped node that is a child of the state already materialised (a descent) costs one chain step
//     instead of a walk from the root.
//
// Exactness: the child cost is still node.cost + cost(e) then -= / += detector terms in the reference's
// order with round-to-nearest adds; pushes happen in row order; the heap performs the same moves.
#include <cooperative_groups.h>
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdlib>

#include "search.h"
#include "sparsify.cuh"

namespace tsr {
namespace {

constexpr unsigned kFull = 0xFFFFFFFFu;
constexpr uint32_t kFired = 0x8000u;
constexpr uint32_t kThr = 0x7FFu;
constexpr uint32_t kKeyInf = 0xFFFFFFFFu;

enum Action : int { kActExpand = 0, kActSkip = 1, kActDone = 2 };
enum ClusterState : uint32_t { kClGo = 1, kClExit = 2 };
enum ChildStatus : uint32_t { kChildOk = 0, kChildDropped = 1 };

struct alignas(16) Ctl {
  // the popped node
  double cost;
  uint32_t chain;
  uint32_t num_dets, depth;
  // search state (owned by warp 0)
  uint32_t min_dets, max_dets;
  uint32_t root_dets;
  int root_pending;
  int terms_cached;
  unsigned long long heap_size, chain_size, pushed, n_visited;
  uint4 key, root_key;
  // expansion
  uint32_t mind, m_off, m_len, nchild;
  uint32_t next_child;  // children are claimed one at a time by the warps
  uint32_t n_fz;        // non-zero words of the node's fired bitset (sm.fz)
  // which node's state pst / fbits / pbb / key currently hold, and how phase 2 gets to the popped node's
  uint32_t state_chain, state_depth;
  int state_valid;
  int rebuild;     // 0 = from the root along the chain, 1 = one step from state_chain, 2 = already there
  uint4 last_raw;  // the popped node's chain entry (rebuild == 1)
  // control
  int action;
  uint32_t status;
  uint32_t work;
  uint32_t item;
  uint32_t n_exp;  // nodes this search has expanded in the current launch (expansion budget)
  // cluster mode: what the helper CTAs read after the "children are ready" barrier
  uint32_t cl_state;  // kClGo / kClExit
  uint32_t cl_slot;   // pool slot of the running search
  // phase profile (SearchParams::prof): clock64 ticks of thread 0 per phase, see tsr_profile
  long long prof_last;
  unsigned long long prof[8];
};

struct Smem {
  Ctl* ctl;
  uint16_t* rank;   // [n_costs * nstride]
  double* fq;       // latency variant: [n_costs * nstride] copy of DeviceTables::fquot (else null)
  uint32_t* rbits;  // [nwords]
  uint32_t* fbits;  // [nwords]
  uint16_t* pst;    // [dpad]
  double* hc;       // [D] (shared or global)
  double* res_cost;  // [max_row_len]
  uint4* res_meta;   // [max_row_len] {error, fired_mask, next_dets | pos << 16, status}
  uint16_t* clist;   // [max_row_len]
  uint16_t* st;      // this warp's private state [dpad]
  // bit-sliced scan (kBits): shared blocking-detector bitset of the node, and this warp's scratch
  uint32_t* pbb;     // [nwords] detectors with a block threshold in the node's state
  uint16_t* fz;      // [nwords] indices of the non-zero words of fbits (ctl->n_fz of them)
  uint32_t* cf;      // [nwords] fired bitset of this warp's current state (node or child)
  uint16_t* tl;      // [kTL] targets
  double* hv;        // [kTL] their heuristic terms (without det_penalty)
  uint16_t* fl;      // [groups][cap] fired-neighbour indices of each group's target
  uint32_t* bl;      // [groups][cap] blocking neighbours: index | threshold << 16
  uint16_t* nzw;     // [nwords] indices of the bitset words that hold a fired or blocking detector in this state
  uint32_t* gc;      // [groups][4] list lengths and the target's own threshold
  // push staging (phase 7): the other warps' regions, idle while warp 0 pushes
  uint4* stage;
  uint32_t stage_cap;  // in 16-byte nodes
  // this search's pool slot and chain arena (visited-set audit)
  uint32_t slot;
  const ChainNode* chain;
  uint32_t cluster;  // CTAs cooperating on this search (1 = this CTA alone)
  uint32_t crank;    // this CTA's rank among them
  // sparsify (sparsify.cuh): this slot's active-entry masks [D][rw] in global memory (null when off), selection scratch
  const uint32_t* act;
  uint32_t rw;
  uint32_t* sp_scratch;
};

constexpr uint32_t kTL = 64;  // targets evaluated per flush

__host__ __device__ inline uint32_t round16(uint32_t bytes) {
  return (bytes + 15u) & ~15u;
}
__host__ __device__ inline uint32_t dpad_of(uint32_t D) {
  return (D + 1 + 7) & ~7u;  // u16 entries, multiple of 16 bytes
}

struct Layout {
  uint32_t off_rank, off_rbits, off_fbits, off_pst, off_hc, off_cost, off_meta, off_clist, off_st, total;
  uint32_t off_pbb, off_fz, warp_stride, w_cf, w_tl, w_hv, w_fl, w_bl, w_gc, w_nzw;  // w_*: offsets inside a warp's region
  uint32_t off_sp, off_fq, stage_extra;
};
__host__ __device__ inline Layout make_layout(uint32_t D, uint32_t nwords, uint32_t rank_entries, uint32_t max_row_len,
                                               uint32_t warps, bool hc_shared, uint32_t bs_group = 0, uint32_t bs_cap = 0,
                                               bool sparsify = false, bool latency = false) {
  Layout L;
  uint32_t at = round16(sizeof(Ctl));
  L.off_rank = at;
  at += round16(rank_entries * 2);
  L.off_rbits = at;
  at += round16(nwords * 4);
  L.off_fbits = at;
  at += round16(nwords * 4);
  L.off_pst = at;
  at += dpad_of(D) * 2;
  L.off_hc = at;
  if (hc_shared) at += round16(D * 8);
  L.off_cost = at;
  at += round16(max_row_len * 8);
  L.off_meta = at;
  at += max_row_len * 16;
  L.off_clist = at;
  at += round16(max_row_len * 2);
  L.off_pbb = at;
  if (bs_group) at += round16(nwords * 4);
  L.off_fz = at;
  at += round16(nwords * 2);
  L.off_sp = at;
  if (sparsify) at += round16(sparsify::kScratchWords * 4);
  L.off_fq = at;
  if (latency) at += round16(rank_entries * 8);
  L.off_st = at;
  // this warp's region: the private state copy (entry-per-lane scans) or the bit-sliced scratch
  uint32_t w = bs_group ? 0 : dpad_of(D) * 2;
  L.w_cf = L.w_tl = L.w_hv = L.w_fl = L.w_bl = L.w_gc = L.w_nzw = 0;
  if (bs_group) {
    const uint32_t groups = 32 / bs_group;
    L.w_cf = w;
    w += round16(nwords * 4);
    L.w_hv = w;
    w += kTL * 8;
    L.w_tl = w;
    w += round16(kTL * 2);
    L.w_bl = w;
    w += round16(groups * bs_cap * 4);
    L.w_fl = w;
    w += round16(groups * bs_cap * 2);
    L.w_gc = w;
    w += round16(groups * 16);
    L.w_nzw = w;
    w += round16(nwords * 2);
  }
  L.warp_stride = w;
  at += warps * w;
  // Push staging (phase 7) = the regions of warps 1.. plus this tail, contiguous: room for a whole row of children and
  // their ancestor ranges (about 2 K + depth nodes of 16 bytes), capped so that small DEMs keep their occupancy.
  const uint32_t want = min(2u * max_row_len + 64u, 768u) * 16u, have = (warps - 1) * w;
  L.stage_extra = have >= want ? 0u : round16(want - have);
  at += L.stage_extra;
  L.total = at;
  return L;
}

__device__ __forceinline__ uint4 ldg16(const uint4* p) {
  return __ldg(p);
}

// ---- row records -----------------------------------------------------------------------------------

__device__ __forceinline__ void slot_apply(uint32_t slot, const uint16_t* st, uint32_t& count, bool& blocked) {
  const uint32_t v = st[slot & 0x3FFF];
  count += v >> 15;
  blocked |= ((slot >> 14) & 0x3FF) < (v & kThr);
}

struct Eval {
  uint32_t count;
  bool blocked;
  uint32_t cost_id;
};

// Entry `idx` of the row at `roff` whose own detector has block threshold Td, in state st.
__device__ __forceinline__ Eval eval_entry(const DeviceTables& t, const uint16_t* st, uint32_t roff, uint32_t idx,
                                           uint32_t Td) {
  const uint4 a = ldg16(t.frec_a + roff + idx);
  Eval e;
  e.count = 1;
  e.blocked = idx < Td;
  e.cost_id = a.x & 0x7F;
  const uint32_t ns = t.fa_slots;
  if (ns > 0) slot_apply(a.x >> 8, st, e.count, e.blocked);
  if (ns > 1) slot_apply(a.y & 0xFFFFFF, st, e.count, e.blocked);
  if (ns > 2) slot_apply((a.y >> 24) | ((a.z & 0xFFFF) << 8), st, e.count, e.blocked);
  if (ns > 3) slot_apply((a.z >> 16) | ((a.w & 0xFF) << 16), st, e.count, e.blocked);
  if (ns > 4) slot_apply(a.w >> 8, st, e.count, e.blocked);
  if (a.x & 0x80) {
    const uint4 b = ldg16(t.frec_b + roff + idx);
    slot_apply(b.x & 0xFFFFFF, st, e.count, e.blocked);
    slot_apply((b.x >> 24) | ((b.y & 0xFFFF) << 8), st, e.count, e.blocked);
    slot_apply((b.y >> 16) | ((b.z & 0xFF) << 16), st, e.count, e.blocked);
    slot_apply(b.z >> 8, st, e.count, e.blocked);
    slot_apply(b.w & 0xFFFFFF, st, e.count, e.blocked);
  }
  return e;
}

// Number of detectors of the error behind a record (instrumented builds only: the early-exit position).
__device__ __forceinline__ uint32_t entry_size(const DeviceTables& t, uint32_t roff, uint32_t idx) {
  const uint32_t e = (uint32_t)__ldg(t.row_err + roff + idx);
  return __ldg(t.e_off + e + 1) - __ldg(t.e_off + e);
}

// get_detcost (tesseract.cc:128-154) without det_penalty for fired detector x in state st; +inf when no
// unblocked error touches x.  kInstr additionally counts the row entries the reference loop would have
// visited before its early exit (counter S of SURVEY.md §8d).
// Row extents of up to 32 detectors at once (lane i asks for detector x: one memory round trip for all of them), with the
// rows' records on their way into L1.  Latency variant of the kernel only (a lone search per SM): there the scans
// that follow are one chain of dependent loads after the other; with many searches per SM the extra instructions
// cost more than the latency they hide (measured: surface d=5 1.41 M -> 1.17 M shots/s).
__device__ __forceinline__ void row_extent(const DeviceTables& t, uint32_t x, bool valid, bool prefetch, uint32_t& roff,
                                           uint32_t& rlen) {
  roff = rlen = 0;
  if (valid) {
    roff = __ldg(t.row_off + x);
    rlen = __ldg(t.row_off + x + 1) - roff;
    if (prefetch) {
      const char* q = reinterpret_cast<const char*>((uintptr_t)(t.frec_a + roff) & ~(uintptr_t)127);
      const char* end = reinterpret_cast<const char*>(t.frec_a + roff + rlen);
      for (; q < end; q += 128) asm volatile("prefetch.global.L1 [%0];" ::"l"(q));
    }
  }
}

template <bool kInstr, bool kLat = false>
__device__ __forceinline__ double scan_row(const DeviceTables& t, const uint16_t* st, const uint16_t* rank, const uint32_t* act,
                                           uint32_t x, uint32_t lane, unsigned long long& scanned, uint32_t roff_in = 0,
                                           uint32_t rlen_in = 0, const double* fq = nullptr) {
  const uint32_t roff = kLat ? roff_in : __ldg(t.row_off + x);
  const uint32_t rlen = kLat ? rlen_in : __ldg(t.row_off + x + 1) - roff;
  const uint32_t Td = st[x] & kThr;
  uint32_t best = kKeyInf, best_cls = 0;
  uint32_t carry = kKeyInf;  // kInstr: best rank before this chunk
  bool counting = true;
#pragma unroll 2
  for (uint32_t base = 0; base < rlen; base += 32) {
    const uint32_t idx = base + lane;
    uint32_t urank = kKeyInf;
    // sparsify: entries that are not active for this shot are not in the row at all (tesseract.cc:729-736)
    const bool present = idx < rlen && (!act || ((__ldcg(act + (size_t)x * t.sp_rw + (base >> 5)) >> lane) & 1u));
    if (present) {
      const Eval e = eval_entry(t, st, roff, idx, Td);
      const uint32_t cls = e.cost_id * t.fnstride + e.count;
      if (!e.blocked) {
        urank = rank[cls];
        const uint32_t key = (urank << 12) | idx;
        if (key < best) {
          best = key;
          best_cls = cls;
        }
      }
    }
    if (kInstr && counting) {
      uint32_t rs = 0;
      if (present) rs = rank[(ldg16(t.frec_a + roff + idx).x & 0x7F) * t.fnstride + entry_size(t, roff, idx)];
      uint32_t incl = urank;
      for (int off = 1; off < 32; off <<= 1) {
        const uint32_t o = __shfl_up_sync(kFull, incl, off);
        if ((int)lane >= off) incl = min(incl, o);
      }
      uint32_t excl = __shfl_up_sync(kFull, incl, 1);
      if (lane == 0) excl = kKeyInf;
      const uint32_t before = min(carry, excl);
      const unsigned mp = __ballot_sync(kFull, present);
      const unsigned mb = __ballot_sync(kFull, present && before != kKeyInf && rs >= before);
      if (mb) {
        scanned += (unsigned long long)__popc(mp & ((2u << ((uint32_t)__ffs((int)mb) - 1)) - 1u));  // up to and including the exit
        counting = false;
      } else {
        scanned += (unsigned long long)__popc(mp);
        carry = min(carry, __shfl_sync(kFull, incl, 31));
      }
    }
  }
  const uint32_t m = __reduce_min_sync(kFull, best);
  if (m == kKeyInf) return __longlong_as_double(0x7FF0000000000000LL);
  const int winner = __ffs((int)__ballot_sync(kFull, best == m)) - 1;
  const uint32_t cls = __shfl_sync(kFull, best_cls, winner);
  return kLat ? fq[cls] : __ldg(t.fquot + cls);
}


// ---- bit-sliced row scan (fast.h: BitTables) ----------------------------------------------------------
//
// A group of G lanes owns one target detector x; lane wl of the group owns entries [32 wl, 32 wl + 32) of
// row[x] as bit masks.  32/G targets are evaluated per warp pass.

struct BsRow {
  uint32_t U;               // unblocked entries of this lane's word
  uint32_t c0, c1, c2, c3;  // bit planes of (fired-detector count - 1) per entry
  uint32_t nw, cmask_off, ncls;
};

// The warp's current state is the node's (sm.pst thresholds, sm.fbits fired, sm.pbb detectors with a threshold)
// or one of its children: fired bitset sm.cf, the min detector's threshold raised to c_thr, and with
// at-most-two the detectors the child's error switched off (fired in the node, not in the child) fully blocked.
struct BsState {
  uint32_t c_mind, c_thr;  // c_mind = 0xFFFFFFFF: the node's own state
  bool amt;
  uint32_t n_nz;           // entries of sm.nzw (bs_list_words)
};

// Lists the bitset words that can contribute a fired or a blocking detector in this state, so that a row scan
// visits those (a dozen) instead of every word of the adjacency row.
__device__ __forceinline__ void bs_list_words(const Smem& sm, BsState& bs, uint32_t nwords, uint32_t lane) {
  uint32_t n = 0;
  for (uint32_t wb = 0; wb < nwords; wb += 32) {
    const uint32_t w = wb + lane;
    bool nz = false;
    if (w < nwords) {
      const uint32_t cfw = sm.cf[w];
      nz = (cfw | sm.pbb[w] | (bs.amt ? (sm.fbits[w] & ~cfw) : 0u)) != 0 || (bs.c_mind >> 5) == w;
    }
    const unsigned m = __ballot_sync(kFull, nz);
    if (nz) sm.nzw[n + __popc(m & ((1u << lane) - 1u))] = (uint16_t)w;
    n += (uint32_t)__popc(m);
  }
  bs.n_nz = n;
  __syncwarp();
}

// Phases A-C for the group's target x.  Must be called by all 32 lanes.
template <bool kCounts>
__device__ __forceinline__ BsRow bs_core(const DeviceTables& t, const Smem& sm, const BsState& bs, bool active, uint32_t x,
                                         uint32_t g, uint32_t wl) {
  const uint32_t G = t.bs_group, cap = t.bs_cap, nwords = t.nwords;
  uint32_t* gc = sm.gc + g * 4;
  if (wl == 0) {
    gc[0] = 0;
    gc[1] = 0;
    gc[2] = 0;
  }
  __syncwarp();
  uint4 ma = make_uint4(0, 0, 0, 0), mb = make_uint4(0, 0, 0, 0);
  if (active) {
    ma = ldg16(t.bs_meta + 2 * (size_t)x);
    mb = ldg16(t.bs_meta + 2 * (size_t)x + 1);
    // A: the fired neighbours of x (their mask index) and the neighbours that carry a block threshold
    const uint32_t* arow = t.adj + (size_t)x * nwords;
    auto scan_word = [&](uint32_t w, uint32_t a) {
      const uint32_t cfw = sm.cf[w];
      uint32_t f = kCounts ? (a & cfw) : 0u;
      uint32_t bsrc = sm.pbb[w];
      if (bs.amt) bsrc |= sm.fbits[w] & ~cfw;
      if ((bs.c_mind >> 5) == w) bsrc |= 1u << (bs.c_mind & 31);
      uint32_t b = a & bsrc;
      if ((x >> 5) == w) f &= ~(1u << (x & 31));
      if (f | b) {
        const uint32_t pre = __ldg(t.bs_adjpre + (size_t)x * nwords + w);
        while (f) {
          const uint32_t bit = (uint32_t)__ffs((int)f) - 1;
          f &= f - 1;
          const uint32_t j = pre + (uint32_t)__popc(a & ((1u << bit) - 1u));
          sm.fl[g * cap + atomicAdd(&gc[0], 1u)] = (uint16_t)j;
        }
        while (b) {
          const uint32_t bit = (uint32_t)__ffs((int)b) - 1;
          b &= b - 1;
          const uint32_t det = w * 32 + bit;
          uint32_t thr = sm.pst[det] & kThr;
          if (det == bs.c_mind) thr = max(thr, bs.c_thr);
          if (bs.amt && (((sm.fbits[w] & ~cfw) >> bit) & 1)) thr = kThr;
          if (det == x) {
            gc[2] = thr;
          } else if (thr) {
            const uint32_t j = pre + (uint32_t)__popc(a & ((1u << bit) - 1u));
            sm.bl[g * cap + atomicAdd(&gc[1], 1u)] = j | (thr << 16);
          }
        }
      }
    };
    for (uint32_t i0 = wl; i0 < bs.n_nz; i0 += 4 * G) {  // four independent loads in flight per lane
      const uint32_t i1 = i0 + G, i2 = i0 + 2 * G, i3 = i0 + 3 * G;
      const uint32_t w0 = sm.nzw[i0];
      const uint32_t w1 = i1 < bs.n_nz ? sm.nzw[i1] : 0u, w2 = i2 < bs.n_nz ? sm.nzw[i2] : 0u, w3 = i3 < bs.n_nz ? sm.nzw[i3] : 0u;
      const uint32_t a0 = __ldg(arow + w0);
      const uint32_t a1 = i1 < bs.n_nz ? __ldg(arow + w1) : 0u;
      const uint32_t a2 = i2 < bs.n_nz ? __ldg(arow + w2) : 0u;
      const uint32_t a3 = i3 < bs.n_nz ? __ldg(arow + w3) : 0u;
      if (a0) scan_word(w0, a0);
      if (a1) scan_word(w1, a1);
      if (a2) scan_word(w2, a2);
      if (a3) scan_word(w3, a3);
    }
  }
  __syncwarp();
  BsRow r;
  r.U = r.c0 = r.c1 = r.c2 = r.c3 = 0;
  r.nw = mb.y;
  r.cmask_off = ma.y;
  r.ncls = mb.z;
  if (active && wl < mb.y) {
    const uint32_t nw = mb.y, len = mb.x, lo = wl * 32;
    const uint32_t nf = gc[0], nb = gc[1], thr0 = gc[2];
    if (kCounts) {
      // B: bit-sliced sum of the fired neighbours' masks
      const uint32_t* mrow = t.bs_masks + ma.x + wl;
      auto add = [&](uint32_t m) {
        const uint32_t t0 = r.c0 & m;
        r.c0 ^= m;
        const uint32_t t1 = r.c1 & t0;
        r.c1 ^= t0;
        const uint32_t t2 = r.c2 & t1;
        r.c2 ^= t1;
        r.c3 ^= t2;
      };
      const uint16_t* fl = sm.fl + g * cap;
      // four loads in flight also for the last (usually the only) group: adding an empty mask changes nothing
      for (uint32_t q = 0; q < nf; q += 4) {
        const uint32_t m0 = __ldg(mrow + (size_t)fl[q] * nw);
        const uint32_t m1 = q + 1 < nf ? __ldg(mrow + (size_t)fl[q + 1] * nw) : 0u;
        const uint32_t m2 = q + 2 < nf ? __ldg(mrow + (size_t)fl[q + 2] * nw) : 0u;
        const uint32_t m3 = q + 3 < nf ? __ldg(mrow + (size_t)fl[q + 3] * nw) : 0u;
        add(m0);
        add(m1);
        add(m2);
        add(m3);
      }
    }
    // C: blocked entries — the row's own prefix, and per blocking neighbour a prefix of the pair list
    uint32_t valid = len - lo >= 32 ? 0xFFFFFFFFu : ((1u << (len - lo)) - 1u);
    if (sm.act) valid &= __ldcg(sm.act + (size_t)x * sm.rw + wl);  // sparsify: this shot's active entries
    uint32_t B = thr0 <= lo ? 0u : (thr0 - lo >= 32 ? 0xFFFFFFFFu : ((1u << (thr0 - lo)) - 1u));
    for (uint32_t q = 0; q < nb; ++q) {
      const uint32_t jb = sm.bl[g * cap + q];
      const uint32_t thr = jb >> 16;
      const uint32_t* po = t.bs_pl_off + ma.x + (size_t)(jb & 0xFFFF) * nw + wl;  // this pair, this lane's word
      const uint32_t pe = __ldg(po + 1);
      for (uint32_t at = __ldg(po); at < pe; ++at) {
        const uint32_t v = __ldg(t.bs_pl + at);
        if ((v >> 16) >= thr) break;
        B |= 1u << (v & 31);
      }
    }
    r.U = valid & ~B;
  }
  return r;
}

// get_detcost (without det_penalty) of targets sm.tl[0..n) in the warp's current state -> sm.hv[0..n).
__device__ __forceinline__ void bs_targets(const DeviceTables& t, const Smem& sm, const BsState& bs, uint32_t n, uint32_t lane) {
  const uint32_t G = t.bs_group, ng = 32 / G;
  const uint32_t g = lane / G, wl = lane % G;
  __syncwarp();
  for (uint32_t base = 0; base < n; base += ng) {
    const uint32_t k = base + g;
    const bool active = k < n;
    const uint32_t x = active ? sm.tl[k] : 0u;
    const BsRow r = bs_core<true>(t, sm, bs, active, x, g, wl);
    uint32_t best = kKeyInf, bcls = 0;
    // D: within a cost class the entry with the most fired detectors has the lowest rank (rank(c, n) falls
    // strictly with n), ties going to the lowest position; the classes then compete on (rank, position).
    // Straight-line per class, so the lanes of a warp stay converged.
    if (r.U) {
      // this lane's word: r.ncls slots of {mask of the word's entries with one cost, cost id}, two slots per 16-byte load
      const uint4* cslots = reinterpret_cast<const uint4*>(t.bs_masks + r.cmask_off) + (size_t)wl * (r.ncls >> 1);
      auto consider = [&](uint32_t cm, uint32_t cid) {
        if (cm) {
          // keep, plane by plane from the top, the entries whose count has that bit set whenever one does: what is left
          // shares the largest count, whose bits are the outcomes of those tests
          uint32_t m = cm, q, v;
          q = m & r.c3;
          v = q ? 8u : 0u;
          m = q ? q : m;
          q = m & r.c2;
          v |= q ? 4u : 0u;
          m = q ? q : m;
          q = m & r.c1;
          v |= q ? 2u : 0u;
          m = q ? q : m;
          q = m & r.c0;
          v |= q ? 1u : 0u;
          m = q ? q : m;
          const uint32_t cls = cid * t.fnstride + v + 1;
          const uint32_t key = ((uint32_t)sm.rank[cls] << 12) | ((uint32_t)__ffs((int)m) - 1);  // the word's offset is added below
          if (key < best) {
            best = key;
            bcls = cls;
          }
        }
      };
      const uint32_t nload = r.ncls >> 1;
      for (uint32_t k = 0; k < nload; k += 2) {  // two loads (four slots) in flight
        const uint4 sa = ldg16(cslots + k);
        const uint4 sb = k + 1 < nload ? ldg16(cslots + k + 1) : make_uint4(0, 0, 0, 0);
        consider(sa.x & r.U, sa.y);
        consider(sa.z & r.U, sa.w);
        consider(sb.x & r.U, sb.y);
        consider(sb.z & r.U, sb.w);
      }
    }
    if (best != kKeyInf) best += wl * 32;
    for (uint32_t off = G >> 1; off >= 1; off >>= 1) {
      const uint32_t ob = __shfl_xor_sync(kFull, best, (int)off);
      const uint32_t oc = __shfl_xor_sync(kFull, bcls, (int)off);
      if (ob < best) {
        best = ob;
        bcls = oc;
      }
    }
    if (active && wl == 0) sm.hv[k] = best == kKeyInf ? __longlong_as_double(0x7FF0000000000000LL) : __ldg(t.fquot + bcls);
  }
  __syncwarp();
}

// Appends the set bits of `word` (detectors base_det + bit) to the warp's target list; the caller flushes
// the list when fewer than 32 free slots remain.
__device__ __forceinline__ void tl_append(const Smem& sm, uint32_t& nt, uint32_t word, uint32_t base_det, uint32_t lane) {
  const uint32_t cnt = (uint32_t)__popc(word);
  // lane b owns bit b: its place in the list is the number of set bits below it
  if ((word >> lane) & 1u) sm.tl[nt + (uint32_t)__popc(word & ((1u << lane) - 1u))] = (uint16_t)(base_det + lane);
  nt += cnt;
}

// ---- queue: exact libstdc++ std::priority_queue<Node, vector<Node>, std::greater<Node>> -------------

__device__ __forceinline__ bool node_greater(const HeapNode& a, const HeapNode& b) {
  return a.cost > b.cost || (a.cost == b.cost && a.num_dets < b.num_dets);  // Node::operator>, tesseract.cc:119-121
}
__device__ __forceinline__ HeapNode heap_ld(const HeapNode* p) {
  const uint4 v = __ldcg(reinterpret_cast<const uint4*>(p));
  HeapNode n;
  n.cost = __hiloint2double((int)v.y, (int)v.x);
  n.chain = v.z;
  n.num_dets = (uint16_t)(v.w & 0xFFFF);
  n.depth = (uint16_t)(v.w >> 16);
  return n;
}
__device__ __forceinline__ void heap_st(HeapNode* p, const HeapNode& n) {
  uint4 v;
  v.x = (uint32_t)__double2loint(n.cost);
  v.y = (uint32_t)__double2hiint(n.cost);
  v.z = n.chain;
  v.w = (uint32_t)n.num_dets | ((uint32_t)n.depth << 16);
  __stcg(reinterpret_cast<uint4*>(p), v);
}

// std::push_heap after push_back (bits/stl_heap.h:135-148), by a whole warp: lane k fetches the ancestor
// k+1 levels above the new leaf, the sift stops at the first ancestor that is not greater than the value,
// and the ancestors below it move down one level — the same moves the sequential loop makes, in one
// memory round trip.
__device__ __forceinline__ void heap_push_warp(HeapNode* heap, unsigned long long size_before, const HeapNode& value,
                                               uint32_t lane) {
  const unsigned long long leaf1 = size_before + 1;  // 1-based index of the new leaf
  const unsigned long long anc1 = leaf1 >> (lane + 1);
  const bool valid = anc1 >= 1;
  HeapNode pn;
  pn.cost = 0;
  pn.chain = 0;
  pn.num_dets = 0;
  pn.depth = 0;
  if (valid) pn = heap_ld(heap + (anc1 - 1));
  const bool up = valid && node_greater(pn, value);
  const unsigned not_up = ~__ballot_sync(kFull, up);
  const uint32_t stop = not_up ? (uint32_t)__ffs((int)not_up) - 1 : 32;  // ancestors 0..stop-1 move down
  if (lane < stop) heap_st(heap + ((leaf1 >> lane) - 1), pn);
  if (lane == 0) heap_st(heap + ((leaf1 >> stop) - 1), value);
  __syncwarp();
}

__device__ __forceinline__ uint4 node_raw(const HeapNode& n) {
  uint4 v;
  v.x = (uint32_t)__double2loint(n.cost);
  v.y = (uint32_t)__double2hiint(n.cost);
  v.z = n.chain;
  v.w = (uint32_t)n.num_dets | ((uint32_t)n.depth << 16);
  return v;
}
__device__ __forceinline__ HeapNode node_of(const uint4& v) {
  HeapNode n;
  n.cost = __hiloint2double((int)v.y, (int)v.x);
  n.chain = v.z;
  n.num_dets = (uint16_t)(v.w & 0xFFFF);
  n.depth = (uint16_t)(v.w >> 16);
  return n;
}
__device__ __forceinline__ uint4 shfl4(const uint4& v, int src) {
  return make_uint4(__shfl_sync(kFull, v.x, src), __shfl_sync(kFull, v.y, src), __shfl_sync(kFull, v.z, src),
                    __shfl_sync(kFull, v.w, src));
}

// top(); pop(): std::pop_heap (bits/stl_heap.h:224-267) then pop_back, by a whole warp; every lane gets the top.
// __adjust_heap walks the hole from the root to a leaf, always taking the child that is not greater than its sibling:
// a chain of dependent loads, one heap level per memory round trip when done by one thread.  Here 30 lanes fetch the
// whole subtree four levels below the hole in ONE round trip, the same walk is replayed on the fetched nodes with
// shuffles (every lane computes the same path), and the moves are stored as the walk goes: depth / 4 round trips.
// The final sift-up of the displaced last element compares it first with the node just moved into the hole's parent,
// which is still in registers (it almost always stops there).
__device__ __noinline__ HeapNode heap_pop(HeapNode* heap, unsigned long long& size, uint32_t lane) {
  uint4 mine = make_uint4(0, 0, 0, 0);
  const long long len = (long long)size - 1;  // elements that remain
  if (lane == 0) mine = __ldcg(reinterpret_cast<const uint4*>(heap));
  if (lane == 1 && len >= 1) mine = __ldcg(reinterpret_cast<const uint4*>(heap + len));      // the displaced last element
  if (lane == 2 && len >= 2) mine = __ldcg(reinterpret_cast<const uint4*>(heap + len - 1));  // a possible single child
  const HeapNode top = node_of(shfl4(mine, 0));
  --size;
  if (len < 1) return top;
  const HeapNode value = node_of(shfl4(mine, 1));
  const uint4 tail_raw = shfl4(mine, 2);
  long long hole = 0;
  const long long last_parent = (len - 1) / 2;  // the walk continues while hole < last_parent: both children exist
  uint4 moved = make_uint4(0, 0, 0, 0);         // the node now stored at the hole's parent
  bool have_moved = false;
  // this lane's place in a 4-level fetch: relative level r (1..4), position j within it
  const uint32_t r = lane < 2 ? 1u : lane < 6 ? 2u : lane < 14 ? 3u : lane < 30 ? 4u : 0u;
  const uint32_t j = lane - ((1u << r) - 2u);
  while (hole < last_parent) {
    uint4 node = make_uint4(0, 0, 0, 0);
    if (r) {
      const long long idx = ((hole + 1) << r) - 1 + (long long)j;
      if (idx < len) node = __ldcg(reinterpret_cast<const uint4*>(heap + idx));
    }
    uint32_t rel = 0;
    for (uint32_t step = 1; step <= 4 && hole < last_parent; ++step) {
      const int lbase = (int)((1u << step) - 2u + 2u * rel);
      const uint4 lraw = shfl4(node, lbase), rraw = shfl4(node, lbase + 1);
      const bool take_left = node_greater(node_of(rraw), node_of(lraw));
      moved = take_left ? lraw : rraw;
      have_moved = true;
      if (lane == 0) __stcg(reinterpret_cast<uint4*>(heap + hole), moved);
      hole = 2 * hole + (take_left ? 1 : 2);
      rel = 2 * rel + (take_left ? 0u : 1u);
    }
  }
  if ((len & 1) == 0 && hole == (len - 2) / 2) {  // the hole's only child is the last element
    moved = tail_raw;
    have_moved = true;
    if (lane == 0) __stcg(reinterpret_cast<uint4*>(heap + hole), moved);
    hole = 2 * hole + 1;
  }
  if (lane == 0) {  // __push_heap of the displaced element from the hole (bits/stl_heap.h:135-148)
    bool first = have_moved;
    while (hole > 0) {
      const long long parent = (hole - 1) / 2;
      const uint4 praw = first ? moved : __ldcg(reinterpret_cast<const uint4*>(heap + parent));
      first = false;
      if (!node_greater(node_of(praw), value)) break;
      __stcg(reinterpret_cast<uint4*>(heap + hole), praw);
      hole = parent;
    }
    __stcg(reinterpret_cast<uint4*>(heap + hole), node_raw(value));
  }
  __syncwarp();
  return top;
}

// ---- visited set: open addressing over 128-bit Zobrist fingerprints of the fired set ----------------

__device__ __forceinline__ bool key_eq(const uint4& a, const uint4& b) {
  return a.x == b.x && a.y == b.y && a.z == b.z && a.w == b.w;
}
__device__ __forceinline__ uint4 key_fix(uint4 k) {
  if ((k.x | k.y | k.z | k.w) == 0) k.w = 1;  // all-zero marks an empty slot
  return k;
}
__device__ bool visited_contains(const uint4* tab, uint64_t cap, uint4 key, uint64_t& where) {
  key = key_fix(key);
  uint64_t i = ((uint64_t)key.x * cap) >> 32;
  while (true) {
    const uint4 v = __ldcg(tab + i);
    if ((v.x | v.y | v.z | v.w) == 0) return false;
    if (key_eq(v, key)) {
      where = i;
      return true;
    }
    if (++i == cap) i = 0;
  }
}
__device__ int visited_insert(uint4* tab, uint64_t cap, uint4 key, uint64_t& where) {
  key = key_fix(key);
  uint64_t i = ((uint64_t)key.x * cap) >> 32;
  while (true) {
    const uint4 v = __ldcg(tab + i);
    if ((v.x | v.y | v.z | v.w) == 0) {
      __stcg(tab + i, key);
      where = i;
      return 1;
    }
    if (key_eq(v, key)) {
      where = i;
      return 0;
    }
    if (++i == cap) i = 0;
  }
}

// Visited-set audit: a fingerprint hit on table entry `where` claims that the candidate state — the node's fired set
// `fbits`, with the `deg` detectors `dj` (one per lane) toggled for a child — equals the state of the node that inserted
// the entry.  Rebuild that node's fired set from its chain and compare: scratch = candidate ^ root ^ toggles(chain) must
// be all zero.
struct AuditArgs {  // passed by value: the cold audit code must not pin the kernel's parameter block in local memory
  const uint32_t* vchain;
  unsigned long long* audit;
  uint32_t* scratch;  // this warp's scratch bitset
  const uint32_t* e_off;
  const int32_t* e_det;
  uint64_t visit_cap;
  uint32_t nwords;
};
__device__ __noinline__ void audit_hit(AuditArgs a, const uint32_t* fbits, const uint32_t* rbits, const ChainNode* chain, uint32_t slot,
                                       uint64_t where, uint32_t deg, uint32_t dj, uint32_t lane) {
  uint32_t* scratch = a.scratch;
  for (uint32_t w = lane; w < a.nwords; w += 32) scratch[w] = fbits[w] ^ rbits[w];
  __syncwarp();
  if (lane < deg) atomicXor(&scratch[dj >> 5], 1u << (dj & 31));
  __syncwarp();
  for (uint32_t at = __ldcg(a.vchain + (size_t)slot * a.visit_cap + where); at != kNoChain;) {
    const uint4 raw = __ldcg(reinterpret_cast<const uint4*>(chain + at));
    const uint32_t eb = __ldg(a.e_off + raw.x), n = __ldg(a.e_off + raw.x + 1) - eb;
    if (lane < n) {
      const uint32_t d = (uint32_t)__ldg(a.e_det + eb + lane);
      atomicXor(&scratch[d >> 5], 1u << (d & 31));
    }
    __syncwarp();
    at = raw.y;
  }
  uint32_t diff = 0;
  for (uint32_t w = lane; w < a.nwords; w += 32) diff |= scratch[w];
  const bool differs = __any_sync(kFull, diff != 0);
  if (lane == 0) {
    atomicAdd(a.audit, 1ull);
    if (differs) atomicAdd(a.audit + 1, 1ull);
  }
}
__device__ __forceinline__ AuditArgs audit_args(const SearchParams& p, uint32_t slot, uint32_t crank) {
  AuditArgs a;
  a.vchain = p.vchain;
  a.audit = p.audit;
  // one scratch bitset per warp of every CTA that can work on the slot's search (up to 8 CTAs of 32 warps in cluster mode)
  a.scratch = p.audit_scratch + (((size_t)slot * 8 + crank) * 32 + (threadIdx.x >> 5)) * p.t.nwords;
  a.e_off = p.t.e_off;
  a.e_det = p.t.e_det;
  a.visit_cap = p.visit_cap;
  a.nwords = p.t.nwords;
  return a;
}
__device__ __forceinline__ uint4 warp_xor(uint4 k) {
  k.x = __reduce_xor_sync(kFull, k.x);
  k.y = __reduce_xor_sync(kFull, k.y);
  k.z = __reduce_xor_sync(kFull, k.z);
  k.w = __reduce_xor_sync(kFull, k.w);
  return k;
}
__device__ __forceinline__ void key_xor(uint4& a, const uint4& b) {
  a.x ^= b.x;
  a.y ^= b.y;
  a.z ^= b.z;
  a.w ^= b.w;
}

struct Counters {
  unsigned long long Q = 0, P = 0, X = 0, L = 0, S = 0, A = 0, V = 0;
};

// Phase profile: thread 0 charges the ticks since the previous mark to `phase`.  Called right after a CTA barrier, so a
// phase's ticks are its critical path including the wait for the slowest warp.
template <bool kDiag>
__device__ __forceinline__ void prof_mark(const SearchParams& p, Ctl& c, uint32_t phase) {
  if (kDiag && p.prof && threadIdx.x == 0) {
    const long long now = clock64();
    c.prof[phase] += (unsigned long long)(now - c.prof_last);
    c.prof_last = now;
  }
}

// ---- one child (one warp): tesseract.cc:533-606 up to, not including, the push ----------------------

template <bool kInstr, bool kBits, bool kLat>
__device__ __forceinline__ void eval_child(const SearchParams& p, const Smem& sm, uint32_t j, uint32_t lane,
                                           const uint4* vtab, Counters& ctr) {
  const DeviceTables& t = p.t;
  const Ctl& c = *sm.ctl;
  uint16_t* st = kBits ? sm.pst : sm.st;  // kBits: the node's state, read only
  const uint32_t D = t.D, nwords = t.nwords;
  const uint32_t pos = sm.clist[j];
  const uint32_t mind = c.mind;
  const uint32_t ei = (uint32_t)__ldg(t.row_err + c.m_off + pos);
  const uint32_t eb = __ldg(t.e_off + ei), deg = __ldg(t.e_off + ei + 1) - eb;
  uint32_t dj = D;
  uint32_t sj = 0;
  if (lane < deg) {
    dj = (uint32_t)__ldg(t.e_det + eb + lane);
    sj = st[dj];
  }
  const unsigned fired_mask = __ballot_sync(kFull, (sj & kFired) != 0);
  const uint32_t next_dets = c.num_dets + deg - 2 * (uint32_t)__popc(fired_mask);
  uint32_t status = kChildOk;
  double cost = 0;
  if (next_dets > c.max_dets) status = kChildDropped;  // tesseract.cc:561
  if (status == kChildOk && p.no_revisit) {            // tesseract.cc:563-565
    uint4 z = make_uint4(0, 0, 0, 0);
    if (lane < deg) z = ldg16(t.zobrist + dj);
    z = warp_xor(z);
    key_xor(z, c.key);
    int seen = 0;
    uint64_t where = 0;
    if (lane == 0) seen = visited_contains(vtab, p.visit_cap, z, where) ? 1 : 0;
    seen = __shfl_sync(kFull, seen, 0);
    ++ctr.V;
    if (seen) status = kChildDropped;
    if (kInstr && seen && p.audit)
      audit_hit(audit_args(p, sm.slot, sm.crank), sm.fbits, sm.rbits, sm.chain, sm.slot, __shfl_sync(kFull, where, 0), deg, dj, lane);
  }
  if (status == kChildOk) {
    // The child's state: toggle the error's detectors, block the min-detector row up to this entry.
    if (kBits) {
      if (lane < deg) atomicXor(&sm.cf[dj >> 5], 1u << (dj & 31));
    } else {
      if (lane < deg) {
        uint32_t nv = sj ^ kFired;
        if (p.at_most_two && (sj & kFired)) nv = (nv & kFired) | kThr;
        st[dj] = (uint16_t)nv;
      }
      __syncwarp();
      if (lane == 0) {
        const uint32_t s = st[mind];
        if ((s & kThr) < pos + 1) st[mind] = (uint16_t)((s & kFired) | (pos + 1));
      }
    }
    __syncwarp();
    BsState bs;
    bs.c_mind = mind;
    bs.c_thr = pos + 1;
    bs.amt = p.at_most_two != 0;
    bs.n_nz = 0;
    if (kBits) bs_list_words(sm, bs, nwords, lane);

    // The reference's add/sub chain, in its order (tesseract.cc:549-585).
    cost = __dadd_rn(c.cost, __ldg(t.e_cost + ei));
    uint32_t nt = 0;  // kBits: targets collected so far (the error's own detectors first)
    if (kBits) {
      // targets: first the error's detectors that become fired, then (below) the fired neighbours; they are
      // evaluated together, 32/G per warp pass, and the chain is applied afterwards in the reference's order
      const unsigned newly = ~fired_mask & ((deg >= 32) ? kFull : ((1u << deg) - 1u));
      if (lane < deg && !((fired_mask >> lane) & 1)) sm.tl[__popc(newly & ((1u << lane) - 1u))] = (uint16_t)dj;
      nt = (uint32_t)__popc(newly);
    } else {
      uint32_t my_off = 0, my_len = 0;  // latency variant: row extents of the error's detectors, fetched together
      if (kLat) row_extent(t, dj, lane < deg, lane < deg && !((fired_mask >> lane) & 1), my_off, my_len);
      for (uint32_t k = 0; k < deg; ++k) {
        const uint32_t d = __shfl_sync(kFull, dj, (int)k);
        const uint32_t roff = kLat ? __shfl_sync(kFull, my_off, (int)k) : 0u, rlen = kLat ? __shfl_sync(kFull, my_len, (int)k) : 0u;
        if ((fired_mask >> k) & 1) {
          cost = __dsub_rn(cost, sm.hc[d]);
        } else {
          const double h = __dadd_rn(scan_row<kInstr, kLat>(t, st, sm.rank, sm.act, d, lane, ctr.S, roff, rlen, sm.fq), p.det_penalty);
          cost = __dadd_rn(cost, h);
        }
      }
    }
    bool own_applied = false;
    auto flush = [&]() {
      bs_targets(t, sm, bs, nt, lane);
      uint32_t at = 0;
      if (!own_applied) {
        for (uint32_t k = 0; k < deg; ++k) {
          const uint32_t d = __shfl_sync(kFull, dj, (int)k);
          if ((fired_mask >> k) & 1) {
            cost = __dsub_rn(cost, sm.hc[d]);
          } else {
            cost = __dadd_rn(cost, __dadd_rn(sm.hv[at++], p.det_penalty));
          }
        }
        own_applied = true;
      }
      for (; at < nt; ++at) {
        cost = __dsub_rn(cost, sm.hc[sm.tl[at]]);
        cost = __dadd_rn(cost, __dadd_rn(sm.hv[at], p.det_penalty));
      }
      nt = 0;
      __syncwarp();
    };
    // eneighbors[ei] ∩ fired: detectors adjacent to one of the error's detectors, fired in the parent
    // (hence in the child), not in the error itself; ascending.
    const uint32_t n_fz = c.n_fz;
    for (uint32_t wb = 0; wb < n_fz; wb += 32) {  // only bitset words that hold a fired detector
      const uint32_t wi = wb + lane;
      const uint32_t w = wi < n_fz ? sm.fz[wi] : 0xFFFFFFFFu;
      uint32_t mine = 0;
      for (uint32_t k = 0; k < deg; k += 4) {  // four adjacency words in flight
        const uint32_t d0 = __shfl_sync(kFull, dj, (int)k), d1 = __shfl_sync(kFull, dj, (int)((k + 1) & 31));
        const uint32_t d2 = __shfl_sync(kFull, dj, (int)((k + 2) & 31)), d3 = __shfl_sync(kFull, dj, (int)((k + 3) & 31));
        if (wi < n_fz) {
          const uint32_t a0 = __ldg(t.adj + (size_t)d0 * nwords + w);
          const uint32_t a1 = k + 1 < deg ? __ldg(t.adj + (size_t)d1 * nwords + w) : 0u;
          const uint32_t a2 = k + 2 < deg ? __ldg(t.adj + (size_t)d2 * nwords + w) : 0u;
          const uint32_t a3 = k + 3 < deg ? __ldg(t.adj + (size_t)d3 * nwords + w) : 0u;
          mine |= a0 | a1 | a2 | a3;
        }
      }
      for (uint32_t k = 0; k < deg; ++k) {
        const uint32_t d = __shfl_sync(kFull, dj, (int)k);
        if ((d >> 5) == w) mine &= ~(1u << (d & 31));
      }
      if (wi < n_fz) mine &= sm.fbits[w];
      unsigned nz = __ballot_sync(kFull, mine != 0);
      if (kBits) {
        while (nz) {
          if (nt + 32 > kTL) flush();
          const uint32_t l = (uint32_t)__ffs((int)nz) - 1;
          nz &= nz - 1;
          tl_append(sm, nt, __shfl_sync(kFull, mine, (int)l), __shfl_sync(kFull, w, (int)l) * 32, lane);
        }
      } else {
        while (nz) {
          const uint32_t l = (uint32_t)__ffs((int)nz) - 1;
          nz &= nz - 1;
          uint32_t word = __shfl_sync(kFull, mine, (int)l);
          if (kLat) {
            // lane i takes the word's i-th set bit: extents of all its rows in one round trip, records prefetched
            const uint32_t wbase = __shfl_sync(kFull, w, (int)l) * 32, cnt = (uint32_t)__popc(word);
            uint32_t my_od = 0;
            if (lane < cnt) {
              uint32_t rest = word;
              for (uint32_t i = 0; i < lane; ++i) rest &= rest - 1;
              my_od = wbase + (uint32_t)__ffs((int)rest) - 1;
            }
            uint32_t my_off, my_len;
            row_extent(t, my_od, lane < cnt, lane < cnt, my_off, my_len);
            for (uint32_t i = 0; i < cnt; ++i) {
              const uint32_t od = __shfl_sync(kFull, my_od, (int)i);
              cost = __dsub_rn(cost, sm.hc[od]);
              const double h = __dadd_rn(scan_row<kInstr, true>(t, st, sm.rank, sm.act, od, lane, ctr.S, __shfl_sync(kFull, my_off, (int)i),
                                                                __shfl_sync(kFull, my_len, (int)i), sm.fq), p.det_penalty);
              cost = __dadd_rn(cost, h);
            }
            word = 0;
          }
          while (word) {
            const uint32_t od = __shfl_sync(kFull, w, (int)l) * 32 + (uint32_t)__ffs((int)word) - 1;
            word &= word - 1;
            cost = __dsub_rn(cost, sm.hc[od]);
            const double h = __dadd_rn(scan_row<kInstr>(t, st, sm.rank, sm.act, od, lane, ctr.S), p.det_penalty);
            cost = __dadd_rn(cost, h);
          }
        }
      }
    }
    if (kBits) flush();
    ctr.A += deg + __ldg(t.e_nnbr + ei);
    if (kBits) {
      if (lane < deg) atomicXor(&sm.cf[dj >> 5], 1u << (dj & 31));  // back to the node's fired set
    } else {
      if (lane < deg) st[dj] = (uint16_t)sj;  // back to the parent state (the min detector is one of them)
    }
    __syncwarp();
    if (cost == __longlong_as_double(0x7FF0000000000000LL)) status = kChildDropped;  // tesseract.cc:587
  }
  if (lane == 0) {
    sm.res_cost[j] = cost;
    sm.res_meta[j] = make_uint4(ei, fired_mask, next_dets | (pos << 16), status);
  }
}

// ---- one search (whole CTA).  Returns the item status; on kItemOk the solution has been written. ------

template <bool kInstr, bool kBits, bool kLat>
__device__ uint32_t run_search(const SearchParams& p, const Smem& sm, uint32_t slot, uint32_t item, uint32_t sol_row,
                               const ResumeRec* resume, Counters& ctr) {
  const DeviceTables& t = p.t;
  const uint32_t tid = threadIdx.x, nthr = blockDim.x;
  const uint32_t lane = tid & 31, warp = tid >> 5, nwarps = nthr >> 5;
  const uint32_t D = t.D, nwords = t.nwords, dpad = dpad_of(D);
  Ctl& c = *sm.ctl;
  const uint32_t shot = item / p.n_trials, trial = item - shot * p.n_trials;
  const uint32_t perm = __ldg(p.trial_perm + trial);
  const uint32_t beam = __ldg(p.trial_beam + trial);
  const uint32_t* order_det = t.order_det + (size_t)perm * D;
  const uint32_t* order_pos = t.order_pos + (size_t)perm * D;
  HeapNode* heap = p.heap + (size_t)slot * p.node_cap;
  ChainNode* chain = p.chain + (size_t)slot * p.node_cap;
  uint4* vtab = p.no_revisit ? p.visited + (size_t)slot * p.visit_cap : nullptr;
  uint32_t* vlist = p.no_revisit ? p.vlist + (size_t)slot * p.vlist_cap : nullptr;
  const double kInf = __longlong_as_double(0x7FF0000000000000LL);

  // Stage the shot: bit-packed syndrome -> root fired bitset (warp 0).
  if (warp == 0) {
    const uint8_t* row = p.dets + (size_t)shot * p.stride;
    const uint32_t nbytes = (D + 7) / 8;
    uint32_t my_fired = 0;
    uint4 root_key = make_uint4(0, 0, 0, 0);
    for (uint32_t w = lane; w < nwords; w += 32) {
      uint32_t word = 0;
      for (uint32_t b = 0; b < 4; ++b) {
        const uint32_t at = w * 4 + b;
        if (at < nbytes) word |= (uint32_t)row[at] << (8 * b);
      }
      if (w * 32 + 32 > D) word &= (1u << (D - w * 32)) - 1u;
      sm.rbits[w] = word;
      my_fired += __popc(word);
      for (uint32_t m = word; m; m &= m - 1) key_xor(root_key, ldg16(t.zobrist + (w * 32 + (uint32_t)__ffs((int)m) - 1)));
    }
    const uint32_t root_dets = __reduce_add_sync(kFull, my_fired);
    root_key = warp_xor(root_key);
    if (lane == 0) {
      c.root_key = root_key;
      c.root_dets = root_dets;
      c.min_dets = root_dets;
      c.max_dets = root_dets + beam;
      c.root_pending = 1;
      c.state_valid = 0;
      c.heap_size = c.chain_size = c.pushed = c.n_visited = 0;
      c.n_exp = 0;
      if (kInstr && p.trace && !resume) {
        p.trace_len[item] = 0;
        p.trace_arena[item] = (unsigned long long)(uintptr_t)chain;
      }
      if (resume) {  // continue a paused search: its queue, arena and visited set are where it left them
        c.min_dets = resume->min_dets;
        c.max_dets = resume->max_dets;
        c.heap_size = resume->heap_size;
        c.chain_size = resume->chain_size;
        c.pushed = resume->pushed;
        c.n_visited = resume->n_visited;
        c.root_pending = 0;
      }
    }
  }
  __syncthreads();
  // sparsify: this shot's active-entry masks (build_sparse_d2e, tesseract.cc:673-737), by the whole CTA
  if (sm.act && !resume) sparsify::select(t, sm.rbits, const_cast<uint32_t*>(sm.act), sm.sp_scratch, warp, nwarps, lane, [] { __syncthreads(); });
  prof_mark<kInstr>(p, c, 6);

  while (true) {
    // ---- phase 1 (warp 0): next node ----------------------------------------------------------------
    if (warp == 0) {
      int action = kActExpand;
      uint32_t status = kItemOk;
      const int root_pending = c.root_pending;
      const unsigned long long heap_size0 = c.heap_size;
      const uint32_t max_dets0 = c.max_dets;
      __syncwarp();
      if (root_pending) {
        if (lane == 0) {
          c.cost = 0;
          c.chain = kNoChain;
          c.num_dets = c.root_dets;
          c.depth = 0;
          c.terms_cached = 0;
          c.rebuild = 0;
        }
      } else if (heap_size0 == 0) {
        action = kActDone;
        status = kItemLowConfidence;  // tesseract.cc:609-615
      } else if (p.expansion_budget && c.n_exp >= p.expansion_budget &&
                 ((kLat && p.resume) || __shfl_sync(kFull, lane == 0 ? __ldcg(p.work_counter) : 0u, 0) >= p.n_work)) {
        // Over budget AND no work left to claim: this search is part of the batch's tail.  Park it for a wider CTA
        // (SearchParams::resume); while the queue still holds work a long search simply runs on — a CTA that retired
        // then would take its slot out of service.  (If the list is full the search also runs on.)
        unsigned int at = 0;
        if (lane == 0) at = atomicAdd(p.paused_count, 1u);
        at = __shfl_sync(kFull, at, 0);
        if (at < p.paused_cap) {
          if (lane == 0) {
            ResumeRec r;
            r.item = item;
            r.slot = slot;
            r.min_dets = c.min_dets;
            r.max_dets = c.max_dets;
            r.heap_size = c.heap_size;
            r.chain_size = c.chain_size;
            r.pushed = c.pushed;
            r.n_visited = c.n_visited;
            p.paused[at] = r;
          }
          action = kActDone;
          status = kItemPaused;
        }
      }
      if (action == kActExpand && !root_pending && heap_size0 != 0) {
        unsigned long long hs = heap_size0;
        const HeapNode node = heap_pop(heap, hs, lane);
        const uint32_t nd = node.num_dets;
        ++ctr.P;
        if (lane == 0) {
          c.heap_size = hs;
          c.cost = node.cost;
          c.chain = node.chain;
          c.num_dets = node.num_dets;
          c.depth = node.depth;
          c.terms_cached = node.chain == kNoChain;  // the root's detector terms are still in hc
          // Most pops continue from the node expanded last (a descent): one chain step instead of a walk.
          int mode = 0;
          if (c.state_valid && node.chain == c.state_chain) {
            mode = 2;
          } else if (c.state_valid && node.chain != kNoChain && nd <= max_dets0 && node.depth == c.state_depth + 1) {
            const uint4 raw = __ldcg(reinterpret_cast<const uint4*>(chain + node.chain));
            if (raw.y == c.state_chain) {
              mode = 1;
              c.last_raw = raw;
            }
          }
          c.rebuild = mode;
        }
        if (nd > max_dets0) action = kActSkip;  // tesseract.cc:434
      }
      if (lane == 0) {
        c.action = action;
        c.status = status;
      }
    }
    __syncthreads();
    prof_mark<kInstr>(p, c, 0);
    if (c.action == kActDone) return c.status;
    if (c.action == kActSkip) {
      __syncthreads();
      continue;
    }

    // ---- phase 2 (all): the node's state from its chain (tesseract.cc:349-369) ------------------------
    const int rebuild = c.rebuild;
    if (rebuild == 0) {
      for (uint32_t i = tid; i < dpad / 2; i += nthr) reinterpret_cast<uint32_t*>(sm.pst)[i] = 0;
      for (uint32_t w = tid; w < nwords; w += nthr) {
        sm.fbits[w] = sm.rbits[w];
        if (kBits) sm.pbb[w] = 0;
      }
      __syncthreads();
      for (uint32_t w = tid; w < nwords; w += nthr)
        for (uint32_t m = sm.rbits[w]; m; m &= m - 1) sm.pst[w * 32 + (uint32_t)__ffs((int)m) - 1] = (uint16_t)kFired;
      __syncthreads();
    }
    if (warp == 0) {
      uint4 key = make_uint4(0, 0, 0, 0);
      auto apply = [&](const uint4& raw) {  // one chain entry: flip the error's detectors, block its row prefix
        const uint32_t err = raw.x, mind_pos = raw.z, off_mask = raw.w;
        const uint32_t eb = __ldg(t.e_off + err), deg = __ldg(t.e_off + err + 1) - eb;
        uint32_t dj = 0;
        if (lane < deg) {
          dj = (uint32_t)__ldg(t.e_det + eb + lane);
          sm.pst[dj] ^= (uint16_t)kFired;
          atomicXor(&sm.fbits[dj >> 5], 1u << (dj & 31));
          key_xor(key, ldg16(t.zobrist + dj));
        }
        __syncwarp();
        if (p.at_most_two && lane < deg && ((off_mask >> lane) & 1)) {
          sm.pst[dj] = (uint16_t)((sm.pst[dj] & kFired) | kThr);
          if (kBits) atomicOr(&sm.pbb[dj >> 5], 1u << (dj & 31));
        }
        __syncwarp();
        if (lane == 0) {
          const uint32_t md = mind_pos >> 16, thr = (mind_pos & 0xFFFF) + 1;
          const uint32_t s = sm.pst[md];
          if ((s & kThr) < thr) sm.pst[md] = (uint16_t)((s & kFired) | thr);
          if (kBits) sm.pbb[md >> 5] |= 1u << (md & 31);
        }
        __syncwarp();
      };
      if (rebuild == 0) {
        for (uint32_t at = c.chain; at != kNoChain;) {
          const uint4 raw = __ldcg(reinterpret_cast<const uint4*>(chain + at));
          apply(raw);
          at = raw.y;
        }
        key = warp_xor(key);
        key_xor(key, c.root_key);
      } else {
        if (rebuild == 1) apply(c.last_raw);
        key = warp_xor(key);
        key_xor(key, c.key);  // the key of the state this step started from
      }
      ctr.L += c.depth;  // the reference walks the whole chain at every pop
      if (lane == 0) {
        c.state_chain = c.chain;
        c.state_depth = c.depth;
        c.state_valid = 1;
      }

      // ---- phase 3 (warp 0): goal, visited, beam window ------------------------------------------------
      int action = kActExpand;
      uint32_t status = kItemOk;
      if (!c.root_pending) {
        if (c.num_dets == 0) {
          // Goal (tesseract.cc:441-473): chain errors in root -> leaf order.
          action = kActDone;
          if (kInstr && p.trace && lane == 0) {
            const uint32_t n = p.trace_len[item]++;
            if (n < p.trace_cap) p.trace[(size_t)item * p.trace_cap + n] = c.chain;
          }
          if (c.depth > p.sol_stride) {
            status = kItemSolutionOverflow;
          } else if (lane == 0) {
            uint32_t* out = p.sol + (size_t)sol_row * p.sol_stride;
            uint32_t at = c.chain;
            for (uint32_t i = 0; i < c.depth; ++i) {
              const uint4 raw = __ldcg(reinterpret_cast<const uint4*>(chain + at));
              out[c.depth - 1 - i] = raw.x;
              at = raw.y;
            }
            p.sol_len[item] = c.depth;
            p.fcost[item] = c.cost;
          }
        } else {
          if (p.no_revisit) {  // tesseract.cc:475-476
            int inserted = 0;
            uint64_t where = 0;
            if (lane == 0) {
              if ((c.n_visited + 1) * 2 > p.visit_cap) {
                inserted = -1;
              } else {
                inserted = visited_insert(vtab, p.visit_cap, key, where);
                if (inserted == 1 && c.n_visited < p.vlist_cap) vlist[c.n_visited] = (uint32_t)where;
                if (kInstr && inserted == 1 && p.vchain) p.vchain[(size_t)slot * p.visit_cap + where] = c.chain;
              }
            }
            inserted = __shfl_sync(kFull, inserted, 0);
            ++ctr.V;
            if (kInstr && inserted == 0 && p.audit)
              audit_hit(audit_args(p, slot, 0), sm.fbits, sm.rbits, chain, slot, __shfl_sync(kFull, where, 0), 0, 0, lane);
            if (inserted < 0) {
              action = kActDone;
              status = kItemNodeOverflow;
            } else if (inserted == 0) {
              action = kActSkip;
            } else if (lane == 0) {
              ++c.n_visited;
            }
          }
          if (kInstr && action == kActExpand && p.trace && lane == 0) {  // tesseract.cc:478-481
            const uint32_t n = p.trace_len[item]++;
            if (n < p.trace_cap) p.trace[(size_t)item * p.trace_cap + n] = c.chain;
          }
          if (action == kActExpand && lane == 0 && c.num_dets < c.min_dets) {  // tesseract.cc:503-511
            c.min_dets = c.num_dets;
            c.max_dets = min(c.max_dets, c.min_dets + beam);
          }
        }
      }
      if (lane == 0) {
        c.key = key;
        c.action = action;
        c.status = status;
      }
    }
    __syncthreads();
    prof_mark<kInstr>(p, c, 1);
    if (c.action == kActDone) return c.status;
    if (c.action == kActSkip) {
      __syncthreads();
      continue;
    }

    // ---- phase 4 (all): private state copies; detector terms of the node ------------------------------
    if (kBits) {
      for (uint32_t w = lane; w < nwords; w += 32) sm.cf[w] = sm.fbits[w];
      __syncwarp();
    } else {
      const uint4* src = reinterpret_cast<const uint4*>(sm.pst);
      uint4* dst = reinterpret_cast<uint4*>(sm.st);
      for (uint32_t i = lane; i < dpad / 8; i += 32) dst[i] = src[i];
      __syncwarp();
    }
    if (warp == 0) {  // the non-zero words of the fired bitset, for the children's neighbour enumeration
      uint32_t n = 0;
      for (uint32_t wb = 0; wb < nwords; wb += 32) {
        const uint32_t w = wb + lane;
        const bool nz = w < nwords && sm.fbits[w] != 0;
        const unsigned m = __ballot_sync(kFull, nz);
        if (nz) sm.fz[n + __popc(m & ((1u << lane) - 1u))] = (uint16_t)w;
        n += (uint32_t)__popc(m);
      }
      if (lane == 0) c.n_fz = n;
    }
    BsState node_state;
    node_state.c_mind = 0xFFFFFFFFu;
    node_state.c_thr = 0;
    node_state.amt = false;
    node_state.n_nz = 0;
    if (kBits) bs_list_words(sm, node_state, nwords, lane);
    if (!c.terms_cached) {
      // (the reference memoises these lazily in detector_cost_cache, tesseract.cc:531,569-582; same values)
      if (kBits) {
        uint32_t nt = 0;
        uint32_t w = warp;
        while (true) {
          if (w < nwords) {
            tl_append(sm, nt, sm.fbits[w], w * 32, lane);
            w += nwarps;
          }
          if (nt == 0 && w >= nwords) break;
          if (w < nwords && nt + 32 <= kTL) continue;
          bs_targets(t, sm, node_state, nt, lane);
          if (lane < nt) sm.hc[sm.tl[lane]] = __dadd_rn(sm.hv[lane], p.det_penalty);
          for (uint32_t i = lane + 32; i < nt; i += 32) sm.hc[sm.tl[i]] = __dadd_rn(sm.hv[i], p.det_penalty);
          __syncwarp();
          nt = 0;
          if (w >= nwords) break;
        }
      } else {
        for (uint32_t w = warp; w < nwords; w += nwarps) {
          uint32_t word = sm.fbits[w];
          while (word) {
            const uint32_t d = w * 32 + (uint32_t)__ffs((int)word) - 1;
            word &= word - 1;
            const double h = __dadd_rn(scan_row<kInstr>(t, sm.st, sm.rank, sm.act, d, lane, ctr.S), p.det_penalty);
            if (lane == 0) sm.hc[d] = h;
          }
        }
      }
    }
    __syncthreads();
    prof_mark<kInstr>(p, c, 2);

    // ---- phase 5 (warp 0): root push, or min detector and the list of children --------------------------
    if (warp == 0) {
      int action = kActExpand;
      uint32_t status = kItemOk;
      if (c.root_pending) {
        // initial cost = in-order sum of the fired detectors' terms (tesseract.cc:411-414)
        double root_cost = 0;
        for (uint32_t w = 0; w < nwords; ++w)
          for (uint32_t m = sm.fbits[w]; m; m &= m - 1) root_cost = __dadd_rn(root_cost, sm.hc[w * 32 + (uint32_t)__ffs((int)m) - 1]);
        action = kActSkip;
        if (root_cost == kInf) {  // tesseract.cc:416-419
          action = kActDone;
          status = kItemLowConfidence;
        } else if (p.node_cap < 1) {
          action = kActDone;
          status = kItemNodeOverflow;
        } else {
          HeapNode hn;
          hn.cost = root_cost;
          hn.chain = kNoChain;
          hn.num_dets = (uint16_t)c.root_dets;
          hn.depth = 0;
          heap_push_warp(heap, 0, hn, lane);
          ++ctr.Q;
          if (lane == 0) {
            c.heap_size = 1;
            c.pushed = 1;
            c.root_pending = 0;  // the loop now pops it back (tesseract.cc:427-432)
          }
        }
      } else {
        ++ctr.X;
        if (lane == 0) ++c.n_exp;
        // min_detector: first fired detector in this trial's order (tesseract.cc:522-528)
        uint32_t best_pos = 0xFFFFFFFFu;
        for (uint32_t w = lane; w < nwords; w += 32)
          for (uint32_t m = sm.fbits[w]; m; m &= m - 1)
            best_pos = min(best_pos, __ldg(order_pos + (w * 32 + (uint32_t)__ffs((int)m) - 1)));
        best_pos = __reduce_min_sync(kFull, best_pos);
        const uint32_t mind = __ldg(order_det + best_pos);
        const uint32_t m_off = __ldg(t.row_off + mind), m_len = __ldg(t.row_off + mind + 1) - m_off;
        const uint32_t m_thr = (kBits ? sm.pst : sm.st)[mind] & kThr;
        // children: unblocked entries of row[min_detector], in row order (tesseract.cc:533-534)
        uint32_t n = 0;
        if (kBits) {
          const uint32_t G = t.bs_group;
          const BsRow r = bs_core<false>(t, sm, node_state, lane < G, mind, lane / G, lane % G);
          uint32_t mine = lane < G ? r.U : 0u, before = 0;
          for (uint32_t l = 0; l < G; ++l) {
            const uint32_t cnt = (uint32_t)__popc(__shfl_sync(kFull, mine, (int)l));
            if (l < lane) before += cnt;
            n += cnt;
          }
          for (uint32_t m = mine; m; m &= m - 1) sm.clist[before++] = (uint16_t)(lane * 32 + (uint32_t)__ffs((int)m) - 1);
        } else {
          for (uint32_t base = 0; base < m_len; base += 32) {
            const uint32_t idx = base + lane;
            bool open = false;
            if (idx < m_len) open = !eval_entry(t, sm.st, m_off, idx, m_thr).blocked;
            if (sm.act) open = open && ((__ldcg(sm.act + (size_t)mind * sm.rw + (base >> 5)) >> lane) & 1u);
            const unsigned mo = __ballot_sync(kFull, open);
            if (open) sm.clist[n + __popc(mo & ((1u << lane) - 1u))] = (uint16_t)idx;
            n += __popc(mo);
          }
        }
        if (lane == 0) {
          c.mind = mind;
          c.m_off = m_off;
          c.m_len = m_len;
          c.nchild = n;
          c.next_child = 0;
        }
      }
      if (lane == 0) {
        c.action = action;
        c.status = status;
      }
    }
    __syncthreads();
    prof_mark<kInstr>(p, c, 3);
    if (c.action == kActDone) return c.status;
    if (c.action == kActSkip) {
      __syncthreads();
      continue;
    }

    // ---- phase 6 (all warps): the children, round-robin ----------------------------------------------------
    // Cluster mode: the helper CTAs wait at the first barrier, then claim children from the same counter (they read this
    // CTA's state and write res_cost / res_meta through distributed shared memory); the second barrier ends the phase.
    const bool clustered = kLat && sm.cluster > 1;
    if (clustered) {
      if (tid == 0) {
        c.cl_state = kClGo;
        c.cl_slot = slot;
      }
      cooperative_groups::this_cluster().sync();
    }
    while (true) {
      uint32_t j = 0;
      if (lane == 0) j = atomicAdd(&c.next_child, 1u);
      j = __shfl_sync(kFull, j, 0);
      if (j >= c.nchild) break;
      eval_child<kInstr, kBits, kLat>(p, sm, j, lane, vtab, ctr);
    }
    if (clustered) cooperative_groups::this_cluster().sync();
    else __syncthreads();
    prof_mark<kInstr>(p, c, 4);

    // ---- phase 7 (warp 0): pushes, in row order (tesseract.cc:590-605) ----------------------------------------
    if (warp == 0) {
      int action = kActSkip;
      uint32_t status = kItemOk;
      unsigned long long heap_size = c.heap_size, chain_size = c.chain_size, pushed = c.pushed;
      const uint32_t n = c.nchild;
      // Staged pushes: the heap nodes K consecutive pushes can touch are the ancestors of leaves [size, size+K),
      // one contiguous index range per level.  They are copied to shared memory (the other warps' scratch, idle
      // in this phase, plus a dedicated tail) in one round trip, the K std::push_heap sifts run there as a pipeline,
      // and the ranges are written back.  A sub-batch that does not fit the staging area takes the windowed path
      // below, which pushes straight to global memory.
      uint32_t K = 0;
      for (uint32_t b0 = 0; b0 < n; b0 += 32) {
        const uint32_t idx = b0 + lane;
        const bool ok = idx < n && sm.res_meta[idx].w == kChildOk;
        const unsigned m = __ballot_sync(kFull, ok);
        if (ok) sm.clist[K + __popc(m & ((1u << lane) - 1u))] = (uint16_t)idx;
        K += (uint32_t)__popc(m);
      }
      __syncwarp();
      uint32_t i0 = 0;  // children (of the compact list) pushed so far
      // (staging costs a few memory round trips of its own: it pays for a long row of children, or a deep heap; measured:
      // staging every expansion slows surface d=5 from 1.8 M to 1.4 M shots/s)
      if ((K >= 96 || (K >= 8 && heap_size >= 4096)) && c.depth != 0xFFFF && pushed + K <= p.pqlimit && heap_size + K <= p.node_cap && chain_size + K <= p.node_cap &&
          chain_size + K < 0xFFFFFFFEull) {
        for (uint32_t i = lane; i < K; i += 32) {  // chain nodes
          const uint4 meta = sm.res_meta[sm.clist[i]];
          __stcg(reinterpret_cast<uint4*>(chain + chain_size + i), make_uint4(meta.x, c.chain, (c.mind << 16) | (meta.z >> 16), meta.y));
        }
        if (K >= 96 && heap_size + K <= sm.stage_cap) {
          // A long row of children onto a small heap (the BB-code regime): the WHOLE heap is staged (node n at slot
          // n - 1) and the K sifts run as a systolic pipeline instead of one after the other.  Push i performs its step s
          // (hole at the s-th ancestor of its leaf: compare with the node one level up, move that node down or settle) at
          // time 2 i + s.  With a lag of two steps between consecutive pushes every node a push reads or overwrites has
          // already seen all the steps the earlier pushes make on it, so the heap ends up exactly as after K sequential
          // std::push_heap calls (bits/stl_heap.h:135-148) - for any mix of leaf levels, including leaves whose parent is
          // an earlier leaf of the same batch - in 2 K + depth steps instead of K x depth.  Lane l runs pushes l, l + 32,
          // ... (a push lasts at most depth + 1 <= 33 steps, the lane is free again after 64).
          const uint32_t hs = (uint32_t)heap_size;
          for (uint32_t q = lane; q < hs; q += 32) sm.stage[q] = __ldcg(reinterpret_cast<const uint4*>(heap + q));
          __syncwarp();
          uint32_t my = lane, hole1 = 0;
          bool active = false;
          uint4 vv = make_uint4(0, 0, 0, 0);
          HeapNode value;
          value.cost = 0;
          value.chain = 0;
          value.num_dets = 0;
          value.depth = 0;
          const uint32_t t_end = 2 * (K - 1) + (32u - (uint32_t)__clz((int)(hs + K))) + 1;
          // Every new node starts at its leaf (push_back, bits/stl_heap.h:135): placed there up front by all lanes, a
          // push then begins with one load of its own slot (nobody else touches that slot before: only later pushes
          // have it on their paths), fetched as soon as the lane's previous push has settled.
          for (uint32_t i = lane; i < K; i += 32) {
            const uint32_t ci = sm.clist[i];
            const uint4 meta = sm.res_meta[ci];
            HeapNode nv;
            nv.cost = sm.res_cost[ci];
            nv.chain = (uint32_t)(chain_size + i);
            nv.num_dets = (uint16_t)(meta.z & 0xFFFF);
            nv.depth = (uint16_t)(c.depth + 1);
            sm.stage[hs + i] = node_raw(nv);
          }
          __syncwarp();
          uint4 nextv = my < K ? sm.stage[hs + my] : make_uint4(0, 0, 0, 0);
          for (uint32_t t = 0; t < t_end; ++t) {
            if (!active && t == 2 * my && my < K) {
              vv = nextv;
              value = node_of(vv);
              active = true;
              hole1 = hs + my + 1;  // 1-based
            }
            if (active) {
              bool up = false;
              uint4 pv = make_uint4(0, 0, 0, 0);
              if (hole1 > 1) {
                pv = sm.stage[(hole1 >> 1) - 1];
                up = node_greater(node_of(pv), value);
              }
              sm.stage[hole1 - 1] = up ? pv : vv;
              if (up) {
                hole1 >>= 1;
              } else {
                active = false;
                my += 32;
                if (my < K) nextv = sm.stage[hs + my];
              }
            }
            __syncwarp();
          }
          for (uint32_t q = lane; q < hs + K; q += 32) __stcg(reinterpret_cast<uint4*>(heap + q), sm.stage[q]);
          __syncwarp();
          heap_size += K;
          chain_size += K;
          pushed += K;
          ctr.Q += K;
          i0 = K;
        }
        while (i0 < K) {
          // one sub-batch = new leaves on ONE tree level, so that "s levels up" names a distinct level for every s
          const unsigned long long lo1 = heap_size + 1;  // 1-based index of the first new leaf
          const uint32_t lvl = 63u - (uint32_t)__clzll((long long)lo1);
          const unsigned long long level_end1 = (2ull << lvl) - 1;
          const uint32_t Kb = (uint32_t)min((unsigned long long)(K - i0), level_end1 - lo1 + 1);
          const unsigned long long hi1 = lo1 + Kb - 1;
          uint32_t lo_s = 0, width_s = 0;  // this lane's level: s = lane + 1 levels above the leaves
          if (lane + 1 <= lvl) {
            lo_s = (uint32_t)(lo1 >> (lane + 1));
            width_s = (uint32_t)(hi1 >> (lane + 1)) - lo_s + 1;
          }
          uint32_t incl = width_s;
          for (int o = 1; o < 32; o <<= 1) {
            const uint32_t v = __shfl_up_sync(kFull, incl, o);
            if ((int)lane >= o) incl += v;
          }
          const uint32_t off_s = Kb + incl - width_s;
          if (Kb + __shfl_sync(kFull, incl, 31) > sm.stage_cap) break;  // does not fit: windowed path for the rest
          // Load the ancestor ranges.  Far above the leaves a range is one or two nodes: every lane fetches its own
          // level's (lane s-1 owns level s), all of them in one memory round trip.  The wide ranges near the leaves go
          // two levels per round trip (both loads are issued before either is stored to shared memory).
          uint32_t n_wide = 0;  // levels 1..n_wide have more than two nodes (widths shrink towards the root)
          {
            const bool narrow = lane + 1 <= lvl && width_s <= 2;
            uint4 g0 = make_uint4(0, 0, 0, 0), g1 = make_uint4(0, 0, 0, 0);
            if (narrow) {
              g0 = __ldcg(reinterpret_cast<const uint4*>(heap + (lo_s - 1)));
              if (width_s == 2) g1 = __ldcg(reinterpret_cast<const uint4*>(heap + lo_s));
            }
            n_wide = (uint32_t)__popc(__ballot_sync(kFull, lane + 1 <= lvl && width_s > 2));
            if (narrow) {
              sm.stage[off_s] = g0;
              if (width_s == 2) sm.stage[off_s + 1] = g1;
            }
          }
          for (uint32_t s0 = 0; s0 < n_wide; s0 += 2) {
            const int sa = (int)s0, sb = (int)min(s0 + 1, 31u);
            const uint32_t wa = __shfl_sync(kFull, width_s, sa), la = __shfl_sync(kFull, lo_s, sa), oa = __shfl_sync(kFull, off_s, sa);
            const uint32_t wb = s0 + 1 < n_wide ? __shfl_sync(kFull, width_s, sb) : 0u;
            const uint32_t lb = __shfl_sync(kFull, lo_s, sb), ob = __shfl_sync(kFull, off_s, sb);
            uint4 ga = make_uint4(0, 0, 0, 0), gb = make_uint4(0, 0, 0, 0);
            if (lane < wa) ga = __ldcg(reinterpret_cast<const uint4*>(heap + (la - 1 + lane)));
            if (lane < wb) gb = __ldcg(reinterpret_cast<const uint4*>(heap + (lb - 1 + lane)));
            if (lane < wa) sm.stage[oa + lane] = ga;
            if (lane < wb) sm.stage[ob + lane] = gb;
            for (uint32_t q = lane + 32; q < wa; q += 32) sm.stage[oa + q] = __ldcg(reinterpret_cast<const uint4*>(heap + (la - 1 + q)));
            for (uint32_t q = lane + 32; q < wb; q += 32) sm.stage[ob + q] = __ldcg(reinterpret_cast<const uint4*>(heap + (lb - 1 + q)));
          }
          __syncwarp();
          // The Kb pushes, 32 at a time.  A push whose value is not smaller than its parent leaves the rest of the heap
          // alone, so every lane checks its own push against its (staged) parent and the leading run of non-sifting
          // pushes commits at once; the first push that does sift is replayed warp-wide (lane s-1 holds the ancestor s
          // levels up, the ballot of `ancestor > value` gives the stop level, the ancestors below it move down one
          // level: the moves of std::push_heap, bits/stl_heap.h:135-148) and the lanes behind it check again.
          {
            const uint32_t lo_1 = __shfl_sync(kFull, lo_s, 0), off_1 = __shfl_sync(kFull, off_s, 0);
            for (uint32_t w0 = 0; w0 < Kb; w0 += 32) {
              const uint32_t j = w0 + lane;
              const bool have = j < Kb;
              HeapNode value;
              value.cost = 0;
              value.chain = 0;
              value.num_dets = 0;
              value.depth = 0;
              if (have) {
                const uint32_t ci = sm.clist[i0 + j];
                const uint4 meta = sm.res_meta[ci];
                value.cost = sm.res_cost[ci];
                value.chain = (uint32_t)(chain_size + j);
                value.num_dets = (uint16_t)(meta.z & 0xFFFF);
                value.depth = (uint16_t)(c.depth + 1);
              }
              const uint4 vv = node_raw(value);
              unsigned remaining = __ballot_sync(kFull, have);
              while (remaining) {
                bool sift = false;
                if (((remaining >> lane) & 1) && lvl >= 1)
                  sift = node_greater(node_of(sm.stage[off_1 + (uint32_t)((lo1 + j) >> 1) - lo_1]), value);
                const unsigned sifting = __ballot_sync(kFull, sift);
                const uint32_t first = sifting ? (uint32_t)__ffs((int)sifting) - 1 : 32u;
                const unsigned run = remaining & (first >= 32 ? kFull : ((1u << first) - 1u));
                if ((run >> lane) & 1) sm.stage[j] = vv;
                remaining &= ~run;
                if (first < 32) {
                  const uint32_t jf = w0 + first;
                  const uint4 braw = shfl4(vv, (int)first);
                  const HeapNode bvalue = node_of(braw);
                  const bool valid = lane + 1 <= lvl;
                  const uint32_t slot = valid ? off_s + (uint32_t)((lo1 + jf) >> (lane + 1)) - lo_s : 0u;
                  uint4 pv = make_uint4(0, 0, 0, 0);
                  if (valid) pv = sm.stage[slot];
                  const bool up = valid && node_greater(node_of(pv), bvalue);
                  const unsigned not_up = ~__ballot_sync(kFull, up);
                  const uint32_t stop = not_up ? (uint32_t)__ffs((int)not_up) - 1 : 32u;  // ancestors 0..stop-1 move down
                  uint32_t below = __shfl_up_sync(kFull, slot, 1);  // slot one level closer to the leaf
                  if (lane == 0) below = jf;
                  if (lane < stop) sm.stage[below] = pv;
                  const uint32_t vslot = stop == 0 ? jf : __shfl_sync(kFull, slot, (int)stop - 1);
                  if (lane == 0) sm.stage[vslot] = braw;
                  remaining &= ~(1u << first);
                }
                __syncwarp();
              }
            }
          }
          for (uint32_t i = lane; i < Kb; i += 32) __stcg(reinterpret_cast<uint4*>(heap + heap_size + i), sm.stage[i]);
          for (uint32_t sft = 0; sft < lvl && sft < 32; ++sft) {  // write the ancestor ranges back
            const uint32_t w_ = __shfl_sync(kFull, width_s, (int)sft);
            const uint32_t lo_ = __shfl_sync(kFull, lo_s, (int)sft), off_ = __shfl_sync(kFull, off_s, (int)sft);
            for (uint32_t q = lane; q < w_; q += 32) __stcg(reinterpret_cast<uint4*>(heap + (lo_ - 1 + q)), sm.stage[off_ + q]);
          }
          __syncwarp();
          heap_size += Kb;
          chain_size += Kb;
          pushed += Kb;
          ctr.Q += Kb;
          i0 += Kb;
        }
      }
      // Otherwise: windows of 32 children.  A push whose value is not smaller than its heap parent leaves the rest
      // of the heap alone, so the leading run of such pushes in a window commits in one round trip (each lane
      // checks its own parent); the first push that has to sift is replayed with the warp-wide sift and the
      // window restarts behind it.
      uint32_t j0 = i0 >= K ? n : sm.clist[i0];
      while (j0 < n) {
        const uint32_t idx = j0 + lane;
        uint4 meta = make_uint4(0, 0, 0, kChildDropped);
        if (idx < n) meta = sm.res_meta[idx];
        const bool ok = meta.w == kChildOk;
        const unsigned okmask = __ballot_sync(kFull, ok);
        const uint32_t cnt = (uint32_t)__popc(okmask);
        if (cnt == 0) {
          j0 += 32;
          continue;
        }
        if (c.depth == 0xFFFF) {
          action = kActDone;
          status = kItemSolutionOverflow;
          break;
        }
        if (pushed + cnt > p.pqlimit) {  // the limit falls inside this window (tesseract.cc:599-605): nothing
          action = kActDone;             // pushed before it is observable once the search is abandoned
          status = kItemLowConfidence;
          ctr.Q += p.pqlimit + 1 - pushed;
          break;
        }
        if (heap_size + cnt > p.node_cap || chain_size + cnt > p.node_cap || chain_size + cnt >= 0xFFFFFFFEull) {
          action = kActDone;
          status = kItemNodeOverflow;
          break;
        }
        const uint32_t rank = (uint32_t)__popc(okmask & ((1u << lane) - 1u));
        HeapNode hn;
        hn.cost = idx < n ? sm.res_cost[idx] : 0.0;
        hn.chain = (uint32_t)(chain_size + rank);
        hn.num_dets = (uint16_t)(meta.z & 0xFFFF);
        hn.depth = (uint16_t)(c.depth + 1);
        bool sift = false;
        if (ok) {
          __stcg(reinterpret_cast<uint4*>(chain + chain_size + rank), make_uint4(meta.x, c.chain, (c.mind << 16) | (meta.z >> 16), meta.y));
          const unsigned long long leaf1 = heap_size + rank + 1;  // 1-based
          if (leaf1 > 1) {
            const unsigned long long parent1 = leaf1 >> 1;
            sift = parent1 > heap_size || node_greater(heap_ld(heap + (parent1 - 1)), hn);
          }
        }
        const unsigned siftmask = __ballot_sync(kFull, sift);
        const uint32_t first = siftmask ? (uint32_t)__ffs((int)siftmask) - 1 : 32u;
        if (ok && lane < first) heap_st(heap + heap_size + rank, hn);
        const uint32_t committed = (uint32_t)__popc(okmask & (first >= 32 ? kFull : ((1u << first) - 1u)));
        heap_size += committed;
        chain_size += committed;
        pushed += committed;
        ctr.Q += committed;
        __syncwarp();
        if (pushed > p.pqlimit) {  // tesseract.cc:599-605
          action = kActDone;
          status = kItemLowConfidence;
          break;
        }
        if (first < 32) {
          HeapNode hb;
          hb.cost = __shfl_sync(kFull, hn.cost, (int)first);
          hb.chain = __shfl_sync(kFull, hn.chain, (int)first);
          const uint32_t packed = __shfl_sync(kFull, (uint32_t)hn.num_dets | ((uint32_t)hn.depth << 16), (int)first);
          hb.num_dets = (uint16_t)(packed & 0xFFFF);
          hb.depth = (uint16_t)(packed >> 16);
          heap_push_warp(heap, heap_size, hb, lane);
          ++heap_size;
          ++chain_size;
          ++pushed;
          ++ctr.Q;
          if (pushed > p.pqlimit) {
            action = kActDone;
            status = kItemLowConfidence;
            break;
          }
          j0 += first + 1;
        } else {
          j0 += 32;
        }
      }
      __syncwarp();
      if (lane == 0) {
        c.heap_size = heap_size;
        c.chain_size = chain_size;
        c.pushed = pushed;
        c.action = action;
        c.status = status;
      }
    }
    __syncthreads();
    prof_mark<kInstr>(p, c, 5);
    if (c.action == kActDone) return c.status;
    __syncthreads();
  }
}

template <bool kInstr, bool kBits, bool kLat>
__global__ void __launch_bounds__(kLat ? 1024 : 512, kLat ? 1 : 2) search_fast_kernel(const SearchParams p) {
  extern __shared__ uint4 smem_raw[];
  uint8_t* base = reinterpret_cast<uint8_t*>(smem_raw);
  const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
  // Cluster mode (wide launches): `C` CTAs per search; CTA 0 of the cluster is the search's owner, the others help with
  // the children.  A "unit" is what owns a pool slot and claims work: a CTA, or a cluster.
  const uint32_t C = kLat ? max(p.cluster, 1u) : 1u;
  const uint32_t crank = C > 1 ? cooperative_groups::this_cluster().block_rank() : 0u;
  const uint32_t unit = blockIdx.x / C, n_units = gridDim.x / C;
  const uint32_t slot = unit;
  const uint32_t D = p.t.D;
  const uint32_t rank_entries = p.t.fn_costs * p.t.fnstride;
  const Layout L = make_layout(D, p.t.nwords, rank_entries, p.t.max_row_len, nwarps, p.hc_shared != 0,
                               kBits ? p.t.bs_group : 0, kBits ? p.t.bs_cap : 0, p.t.sp_on != 0, kLat);
  Smem sm;
  sm.ctl = reinterpret_cast<Ctl*>(base);
  sm.rank = reinterpret_cast<uint16_t*>(base + L.off_rank);
  sm.rbits = reinterpret_cast<uint32_t*>(base + L.off_rbits);
  sm.fbits = reinterpret_cast<uint32_t*>(base + L.off_fbits);
  sm.pst = reinterpret_cast<uint16_t*>(base + L.off_pst);
  sm.hc = p.hc_shared ? reinterpret_cast<double*>(base + L.off_hc) : p.hcache + (size_t)slot * D;
  sm.res_cost = reinterpret_cast<double*>(base + L.off_cost);
  sm.res_meta = reinterpret_cast<uint4*>(base + L.off_meta);
  sm.clist = reinterpret_cast<uint16_t*>(base + L.off_clist);
  uint8_t* wbase = base + L.off_st + (size_t)warp * L.warp_stride;
  sm.st = reinterpret_cast<uint16_t*>(wbase);
  sm.pbb = reinterpret_cast<uint32_t*>(base + L.off_pbb);
  sm.fz = reinterpret_cast<uint16_t*>(base + L.off_fz);
  sm.cf = reinterpret_cast<uint32_t*>(wbase + L.w_cf);
  sm.tl = reinterpret_cast<uint16_t*>(wbase + L.w_tl);
  sm.hv = reinterpret_cast<double*>(wbase + L.w_hv);
  sm.fl = reinterpret_cast<uint16_t*>(wbase + L.w_fl);
  sm.bl = reinterpret_cast<uint32_t*>(wbase + L.w_bl);
  sm.gc = reinterpret_cast<uint32_t*>(wbase + L.w_gc);
  sm.nzw = reinterpret_cast<uint16_t*>(wbase + L.w_nzw);
  sm.act = p.t.sp_on ? p.sp_act + (size_t)slot * D * p.t.sp_rw : nullptr;
  sm.rw = p.t.sp_rw;
  sm.sp_scratch = reinterpret_cast<uint32_t*>(base + L.off_sp);
  sm.stage = reinterpret_cast<uint4*>(base + L.off_st + L.warp_stride);
  sm.stage_cap = ((nwarps - 1) * L.warp_stride + L.stage_extra) / 16;
  sm.fq = kLat ? reinterpret_cast<double*>(base + L.off_fq) : nullptr;
  sm.cluster = C;
  sm.crank = crank;
  for (uint32_t i = tid; i < rank_entries; i += blockDim.x) {
    sm.rank[i] = __ldg(p.t.frank + i);
    if (kLat) sm.fq[i] = __ldg(p.t.fquot + i);
  }
  if (tid == 0) {
    for (auto& v : sm.ctl->prof) v = 0;
    sm.ctl->prof_last = clock64();
  }
  __syncthreads();

  Counters ctr;
  unsigned long long searches = 0;
  if (kLat && crank != 0) {
    // Helper CTA of a cluster: per expansion of the owner's search, evaluate children.  After the "children are ready"
    // barrier the helper copies the node's state — thresholds, fired / blocking bitsets, detector terms, child list, the
    // control block — from the OWNER's shared memory into its own (distributed shared memory, one 16-byte load per
    // thread), so that the scans run on local memory exactly as in the owner; only the claims (an atomicAdd on the
    // owner's counter) and the results (res_cost / res_meta) cross SMs.
    auto cluster = cooperative_groups::this_cluster();
    Smem hs = sm;
    hs.res_cost = cluster.map_shared_rank(sm.res_cost, 0);
    hs.res_meta = cluster.map_shared_rank(sm.res_meta, 0);
    Ctl* owner = cluster.map_shared_rank(sm.ctl, 0);
    const uint4* o_base = reinterpret_cast<const uint4*>(cluster.map_shared_rank(base, 0));
    uint4* l_base = reinterpret_cast<uint4*>(base);
    const uint32_t dpad = dpad_of(D);
    // [off_rbits, off_st): rbits, fbits, pst, hc (when shared), res arrays (skipped below), clist, pbb, fz, ... — contiguous
    const uint32_t copy_lo = L.off_rbits / 16, copy_hi = L.off_st / 16;
    const uint32_t skip_lo = L.off_cost / 16, skip_hi = L.off_clist / 16;  // res_cost / res_meta stay the owner's
    while (true) {
      cluster.sync();  // the owner has a list of children ready, or is done
      if (owner->cl_state == kClExit) {
        cluster.sync();  // the owner's shared memory must outlive this read: nobody leaves before everybody has looked
        break;
      }
      for (uint32_t i = tid; i < round16(sizeof(Ctl)) / 16; i += blockDim.x) l_base[i] = reinterpret_cast<const uint4*>(owner)[i];
      for (uint32_t i = copy_lo + tid; i < copy_hi; i += blockDim.x)
        if (i < skip_lo || i >= skip_hi) l_base[i] = o_base[i];
      __syncthreads();
      const uint32_t cs = hs.ctl->cl_slot;
      hs.slot = cs;
      hs.chain = p.chain + (size_t)cs * p.node_cap;
      if (!p.hc_shared) hs.hc = p.hcache + (size_t)cs * D;
      hs.act = p.t.sp_on ? p.sp_act + (size_t)cs * D * p.t.sp_rw : nullptr;
      const uint4* vtab = p.no_revisit ? p.visited + (size_t)cs * p.visit_cap : nullptr;
      const uint32_t nchild = hs.ctl->nchild;
      bool have_state = false;  // this warp's private copy of the node's state, made when it gets its first child
      while (true) {
        uint32_t j = 0;
        if (lane == 0) j = atomicAdd(&owner->next_child, 1u);
        j = __shfl_sync(kFull, j, 0);
        if (j >= nchild) break;
        if (!have_state) {
          if (kBits) {
            for (uint32_t w = lane; w < p.t.nwords; w += 32) hs.cf[w] = hs.fbits[w];
          } else {
            const uint4* src = reinterpret_cast<const uint4*>(hs.pst);
            uint4* dst = reinterpret_cast<uint4*>(hs.st);
            for (uint32_t i = lane; i < dpad / 8; i += 32) dst[i] = src[i];
          }
          __syncwarp();
          have_state = true;
        }
        eval_child<kInstr, kBits, kLat>(p, hs, j, lane, vtab, ctr);
      }
      cluster.sync();  // all children of this expansion are evaluated
    }
  } else {
  const bool resuming = kLat && p.resume;  // paused searches are continued by the wide variants only
  const uint32_t n_paused = p.resume_n;
  uint32_t resume_at = unit;
  while (true) {
    uint32_t item, cur_slot = slot, sol_row = 0;
    const ResumeRec* rec = nullptr;
    if (resuming) {  // continue paused searches (one per CTA and round), each in the pool slot it was started in
      if (resume_at >= n_paused) break;
      rec = p.resume_list + resume_at;
      resume_at += n_units;
      item = rec->item;
      sol_row = item;
      cur_slot = rec->slot;
      sm.hc = p.hc_shared ? sm.hc : p.hcache + (size_t)cur_slot * D;
      sm.act = p.t.sp_on ? p.sp_act + (size_t)cur_slot * D * p.t.sp_rw : nullptr;
    } else {
      if (tid == 0) sm.ctl->work = atomicAdd(p.work_counter, 1u);
      __syncthreads();
      const uint32_t work = sm.ctl->work;
      if (work >= p.n_work) break;
      item = p.item_ids ? __ldg(p.item_ids + work) : work;
      if (p.shot_order) item = __ldg(p.shot_order + work / p.n_trials) * p.n_trials + work % p.n_trials;
      sol_row = p.sol_by_work ? work : item;
    }
    sm.slot = cur_slot;
    sm.chain = p.chain + (size_t)cur_slot * p.node_cap;
    const uint32_t status = run_search<kInstr, kBits, kLat>(p, sm, cur_slot, item, sol_row, rec, ctr);
    __syncthreads();
    if (!rec) ++searches;
    if (tid == 0) {
      p.status[item] = (uint8_t)status;
      if (status != kItemOk) p.sol_len[item] = 0;
    }
    // A paused search keeps its queue, arena, visited set and masks in its slot.  In the narrow launch that is this
    // CTA's own slot: the CTA retires.  A resume launch works on the searches' slots and simply takes its next one.
    if (status == kItemPaused) {
      if (resuming) continue;
      break;
    }
    if (sm.act)  // back to the mandatory masks for the slot's next search
      sparsify::restore(p.t, sm.rbits, const_cast<uint32_t*>(sm.act), warp, nwarps, lane, [] { __syncthreads(); });
    // Empty the visited table for the next search of this slot.
    const unsigned long long n_visited = sm.ctl->n_visited;
    if (p.no_revisit && n_visited) {
      uint4* vtab = p.visited + (size_t)cur_slot * p.visit_cap;
      const uint4 zero = make_uint4(0, 0, 0, 0);
      if (n_visited <= p.vlist_cap) {
        const uint32_t* vlist = p.vlist + (size_t)cur_slot * p.vlist_cap;
        for (unsigned long long i = tid; i < n_visited; i += blockDim.x) __stcg(vtab + __ldcg(vlist + i), zero);
      } else {
        for (unsigned long long i = tid; i < p.visit_cap; i += blockDim.x) __stcg(vtab + i, zero);
      }
    }
    __syncthreads();
  }
  if (kLat && C > 1) {  // release the helpers; the second barrier keeps this CTA's shared memory alive while they read the flag
    if (tid == 0) sm.ctl->cl_state = kClExit;
    cooperative_groups::this_cluster().sync();
    cooperative_groups::this_cluster().sync();
  }
  }
  if (kInstr && tid == 0 && p.prof && crank == 0) {  // (a helper's control block is a copy of its owner's)
    sm.ctl->prof[7] = ctr.X;
    for (int i = 0; i < 8; ++i) atomicAdd(p.prof + i, sm.ctl->prof[i]);
  }
  if (lane == 0 && p.counters) {
    if (warp == 0) {
      atomicAdd(p.counters + 0, ctr.Q);
      atomicAdd(p.counters + 1, ctr.P);
      atomicAdd(p.counters + 2, ctr.X);
      atomicAdd(p.counters + 3, ctr.L);
      atomicAdd(p.counters + 7, searches);
    }
    atomicAdd(p.counters + 4, ctr.S);
    atomicAdd(p.counters + 5, ctr.A);
    atomicAdd(p.counters + 6, ctr.V);
  }
}

template <bool kInstr, bool kBits, bool kLat>
int configure(size_t smem) {
  if (const char* e = std::getenv("TSR_CARVEOUT"))  // experiments: shared-memory carveout in percent (the rest is L1)
    if (cudaError_t err = cudaFuncSetAttribute(search_fast_kernel<kInstr, kBits, kLat>, cudaFuncAttributePreferredSharedMemoryCarveout, std::atoi(e)))
      return (int)err;
  if (smem > 48 * 1024)
    return (int)cudaFuncSetAttribute(search_fast_kernel<kInstr, kBits, kLat>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  return 0;
}

}  // namespace

// Variants.  `diag` (template kInstr): the diagnostics build — counter S of the entry-per-lane scan, the visited-set audit,
// the visualization trace and the phase profile are compiled in only there, so that production launches carry none of
// that code (measured: the audit's hit index alone cost the headline 3.6 %).  `latency`: the variant for a lone search
// per SM (wide CTAs of up to 1024 threads: the tail and the retry tiers); entry-per-lane scans then use batched row
// extents, L1 prefetch and the quotient table in shared memory.
uint32_t search_fast_smem_bytes(const DeviceTables& t, uint32_t warps, bool hc_shared, bool bits, bool latency) {
  return make_layout(t.D, t.nwords, t.fn_costs * t.fnstride, t.max_row_len, warps, hc_shared, bits ? t.bs_group : 0,
                     bits ? t.bs_cap : 0, t.sp_on != 0, latency).total;
}

namespace {
template <bool kInstr, bool kBits, bool kLat>
int occupancy_of(uint32_t warps, size_t smem) {
  int n = 0;
  if (configure<kInstr, kBits, kLat>(smem) != 0) return 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, search_fast_kernel<kInstr, kBits, kLat>, (int)(warps * 32), smem) != cudaSuccess) return 0;
  return n;
}
template <bool kInstr, bool kBits, bool kLat>
int launch_variant(const SearchParams& p, uint32_t blocks, uint32_t warps, size_t smem, cudaStream_t st) {
  if (int e = configure<kInstr, kBits, kLat>(smem)) return e;
  if (kLat && p.cluster > 1) {  // thread-block clusters: blocks = units x cluster size
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(blocks, 1, 1);
    cfg.blockDim = dim3(warps * 32, 1, 1);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = p.cluster;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return (int)cudaLaunchKernelEx(&cfg, search_fast_kernel<kInstr, kBits, kLat>, p);
  }
  search_fast_kernel<kInstr, kBits, kLat><<<blocks, warps * 32, smem, st>>>(p);
  return (int)cudaGetLastError();
}
}  // namespace

int search_fast_max_blocks_per_sm(const DeviceTables& t, uint32_t warps, bool hc_shared, bool bits, bool latency) {
  const size_t smem = search_fast_smem_bytes(t, warps, hc_shared, bits, latency);
  if (smem > 227 * 1024) return 0;
  // a production variant and its diagnostics twin share the block shape
  if (latency) {
    if (bits) return std::min(occupancy_of<false, true, true>(warps, smem), occupancy_of<true, true, true>(warps, smem));
    return std::min(occupancy_of<false, false, true>(warps, smem), occupancy_of<true, false, true>(warps, smem));
  }
  if (bits) return std::min(occupancy_of<false, true, false>(warps, smem), occupancy_of<true, true, false>(warps, smem));
  return std::min(occupancy_of<false, false, false>(warps, smem), occupancy_of<true, false, false>(warps, smem));
}

int launch_search_fast(const SearchParams& p, uint32_t blocks, uint32_t warps, bool diag, bool bits, bool latency, void* stream) {
  if ((p.n_work == 0 && !p.resume) || blocks == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  const size_t smem = search_fast_smem_bytes(p.t, warps, p.hc_shared != 0, bits, latency);
  switch ((diag ? 4 : 0) | (bits ? 2 : 0) | (latency ? 1 : 0)) {
    case 0: return launch_variant<false, false, false>(p, blocks, warps, smem, st);
    case 1: return launch_variant<false, false, true>(p, blocks, warps, smem, st);
    case 2: return launch_variant<false, true, false>(p, blocks, warps, smem, st);
    case 3: return launch_variant<false, true, true>(p, blocks, warps, smem, st);
    case 4: return launch_variant<true, false, false>(p, blocks, warps, smem, st);
    case 5: return launch_variant<true, false, true>(p, blocks, warps, smem, st);
    case 6: return launch_variant<true, true, false>(p, blocks, warps, smem, st);
    default: return launch_variant<true, true, true>(p, blocks, warps, smem, st);
  }
}

}  // namespace tsr
