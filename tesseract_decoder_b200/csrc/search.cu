// The A* most-likely-error search as a hand-written sm_100a kernel: one warp per (shot, trial).
//
// What it reproduces (reference, file:line):
//   decode_to_errors_with_graph  tesseract.cc:380-616   the search loop, pruning and termination
//   get_detcost                  tesseract.cc:128-154   the admissible per-detector heuristic
//   flip_detectors_and_block_errors tesseract.cc:349-369  node state from its ancestor chain
//   Node::operator>              tesseract.cc:119-121   queue order, with libstdc++'s binary-heap
//                                                       sift order (bits/stl_heap.h) for full ties
//   decode_to_errors             tesseract.cc:295-347   ensemble / beam-climbing resolve (resolve_kernel)
//   cost_from_errors, get_flipped_observables  tesseract.cc:618-658
//
// How it differs in structure (B200-first): the reference re-creates two
// 8*num_errors-byte arrays per popped node.  Here the whole mutable state of a search is one 16-bit
// word per detector in shared memory:
//     bit 15      the detector is fired
//     bits 0..14  block threshold T: the first T entries of this detector's row are blocked
// An error is blocked iff, for one of its detectors, its rank in that detector's row is < T; the
// number of its fired detectors is read off bit 15.  Rows are streamed from L2 as 16-byte records
// that already carry (detector, rank) of the error's other detectors, so a row scan is coalesced
// vector loads plus shared-memory lookups — no per-error gather, no O(num_errors) work anywhere.
// A child state is the parent state with <= degree+1 words patched (and restored afterwards).
//
// Exactness: every floating-point operation of the reference (products in the ratio comparisons,
// the quotient, the add/sub chain of the child cost) is issued in the same order with explicit
// round-to-nearest intrinsics (no FMA contraction); the sequential scan of a row is emulated
// event-by-event (first break / first improvement) so ties resolve as in the reference loop.
#include <cuda_runtime.h>

#include <cstdint>

#include "search.h"
#include "sparsify.cuh"

namespace tsr {
namespace {

constexpr unsigned kFull = 0xFFFFFFFFu;
constexpr uint16_t kFiredBit = 0x8000;
constexpr uint16_t kThreshMask = 0x7FFF;

__device__ __forceinline__ uint4 ldg16(const uint4* p) {
  return __ldg(p);
}

struct Entry {
  double cost;
  uint32_t size;
  uint32_t count;
  bool blocked;
};

__device__ __forceinline__ void slot_apply(uint32_t slot, const uint16_t* st, uint32_t& count, bool& blocked) {
  uint32_t s = st[slot & 0xFFFF];
  count += s >> 15;
  blocked |= (slot >> 16) < (s & kThreshMask);
}

// Decodes the record of entry `idx` of the row starting at `roff` (row detector threshold Td): the
// error's cost and degree, how many of its detectors are fired in `st` (the row's own detector is
// fired by construction) and whether it is blocked.
__device__ __forceinline__ Entry decode_entry(const DeviceTables& t, const uint16_t* st, uint32_t roff, uint32_t idx,
                                              uint32_t Td) {
  Entry e;
  const uint32_t at = roff + idx;
  uint4 a = ldg16(t.rec_a + at);
  e.cost = __hiloint2double((int)a.y, (int)a.x);
  e.size = a.z & 0xFFFF;
  const uint32_t flags = a.z >> 16;
  e.count = 1;
  e.blocked = idx < Td;
  if (flags & 1) {
    // Degree beyond the inline slots: walk the error's CSR lists instead.
    const uint32_t err = (uint32_t)__ldg(t.row_err + at);
    const uint32_t b = __ldg(t.e_off + err), n = __ldg(t.e_off + err + 1) - b;
    uint32_t fired = 0;
    for (uint32_t j = 0; j < n; ++j) {
      uint32_t s = st[__ldg(t.e_det + b + j)];
      fired += s >> 15;
      e.blocked |= (uint32_t)__ldg(t.e_rank + b + j) < (s & kThreshMask);
    }
    e.count = fired;
    return e;
  }
  slot_apply(a.w, st, e.count, e.blocked);
  if (e.size > 2) {
    uint4 b = ldg16(t.rec_b + at);
    slot_apply(b.x, st, e.count, e.blocked);
    slot_apply(b.y, st, e.count, e.blocked);
    slot_apply(b.z, st, e.count, e.blocked);
    slot_apply(b.w, st, e.count, e.blocked);
    if (e.size > 6) {
      uint4 c = ldg16(t.rec_c + at);
      slot_apply(c.x, st, e.count, e.blocked);
      slot_apply(c.y, st, e.count, e.blocked);
      slot_apply(c.z, st, e.count, e.blocked);
      slot_apply(c.w, st, e.count, e.blocked);
    }
  }
  return e;
}

// get_detcost (tesseract.cc:128-154) for fired detector d in state `st`, computed by the whole warp.
// The reference walks the row in order keeping (min_cost, min_count); it stops at the first entry with
// cost*min_count >= min_cost*size and replaces the minimum at unblocked entries with
// cost*min_count < min_cost*count.  Each 32-entry chunk is decoded in parallel and then those two
// event kinds are replayed in row order.
__device__ __forceinline__ double detcost(const DeviceTables& t, const uint16_t* st, const uint32_t* act, uint32_t d,
                                          double det_penalty, uint32_t lane, unsigned long long& scanned) {
  const uint32_t roff = __ldg(t.row_off + d);
  const uint32_t rlen = __ldg(t.row_off + d + 1) - roff;
  const uint32_t Td = st[d] & kThreshMask;
  double best_cost = __longlong_as_double(0x7FF0000000000000LL);
  double best_cnt = 4294967295.0;  // std::numeric_limits<uint32_t>::max()
  bool done = false;
  for (uint32_t base = 0; base < rlen && !done; base += 32) {
    const uint32_t idx = base + lane;
    // sparsify: entries that are not active for this shot are not in the row at all (tesseract.cc:729-736)
    const bool valid = idx < rlen && (!act || ((__ldcg(act + (size_t)d * t.sp_rw + (base >> 5)) >> lane) & 1u));
    Entry e;
    e.cost = 0;
    e.size = 1;
    e.count = 1;
    e.blocked = true;
    if (valid) e = decode_entry(t, st, roff, idx, Td);
    const double size_d = (double)e.size, count_d = (double)e.count;
    unsigned remaining = __ballot_sync(kFull, valid);
    uint32_t examined = __popc(remaining);
    while (true) {
      const bool brk = valid && (__dmul_rn(e.cost, best_cnt) >= __dmul_rn(best_cost, size_d));
      const bool upd = valid && !e.blocked && (__dmul_rn(e.cost, best_cnt) < __dmul_rn(best_cost, count_d));
      const unsigned mb = __ballot_sync(kFull, brk) & remaining;
      const unsigned mu = __ballot_sync(kFull, upd) & remaining;
      const uint32_t fb = mb ? (uint32_t)__ffs((int)mb) - 1 : 32;
      const uint32_t fu = mu ? (uint32_t)__ffs((int)mu) - 1 : 32;
      if (fb <= fu) {
        if (fb < 32) {
          done = true;
          examined = fb + 1;
        }
        break;
      }
      best_cost = __shfl_sync(kFull, e.cost, (int)fu);
      best_cnt = __shfl_sync(kFull, count_d, (int)fu);
      remaining &= fu >= 31 ? 0u : ~((2u << fu) - 1u);
    }
    scanned += examined;
  }
  return __dadd_rn(__ddiv_rn(best_cost, best_cnt), det_penalty);
}

// ---- queue: exact libstdc++ std::priority_queue<Node, vector<Node>, std::greater<Node>> ---------

__device__ __forceinline__ bool node_greater(const HeapNode& a, const HeapNode& b) {
  // Node::operator> (tesseract.cc:119-121)
  return a.cost > b.cost || (a.cost == b.cost && a.num_dets < b.num_dets);
}

// std::push_heap after push_back (bits/stl_heap.h:135-148): sift the new value up while parent > value.
__device__ void heap_push(HeapNode* heap, uint64_t& size, const HeapNode& value) {
  int64_t hole = (int64_t)size++;
  while (hole > 0) {
    int64_t parent = (hole - 1) / 2;
    HeapNode pn = heap[parent];
    if (!node_greater(pn, value)) break;
    heap[hole] = pn;
    hole = parent;
  }
  heap[hole] = value;
}

// top(); pop(): std::pop_heap (bits/stl_heap.h:224-267) then pop_back.
__device__ HeapNode heap_pop(HeapNode* heap, uint64_t& size) {
  HeapNode top = heap[0];
  if (size > 1) {
    const int64_t len = (int64_t)size - 1;
    HeapNode value = heap[len];
    int64_t hole = 0, child = 0;
    while (child < (len - 1) / 2) {
      child = 2 * (child + 1);
      HeapNode r = heap[child], l = heap[child - 1];
      if (node_greater(r, l)) {
        --child;
        r = l;
      }
      heap[hole] = r;
      hole = child;
    }
    if ((len & 1) == 0 && child == (len - 2) / 2) {
      child = 2 * (child + 1);
      heap[hole] = heap[child - 1];
      hole = child - 1;
    }
    while (hole > 0) {
      int64_t parent = (hole - 1) / 2;
      HeapNode pn = heap[parent];
      if (!node_greater(pn, value)) break;
      heap[hole] = pn;
      hole = parent;
    }
    heap[hole] = value;
  }
  --size;
  return top;
}

// ---- visited set: open addressing over 128-bit Zobrist fingerprints of the fired set ------------

__device__ __forceinline__ bool key_eq(const uint4& a, const uint4& b) {
  return a.x == b.x && a.y == b.y && a.z == b.z && a.w == b.w;
}
__device__ __forceinline__ uint4 key_fix(uint4 k) {
  if ((k.x | k.y | k.z | k.w) == 0) k.w = 1;  // all-zero marks an empty slot
  return k;
}
__device__ bool visited_contains(const uint4* tab, uint64_t cap, uint4 key) {
  key = key_fix(key);
  uint64_t i = ((uint64_t)key.x * cap) >> 32;
  while (true) {
    uint4 v = tab[i];
    if ((v.x | v.y | v.z | v.w) == 0) return false;
    if (key_eq(v, key)) return true;
    if (++i == cap) i = 0;
  }
}
// Returns 1 if inserted, 0 if already present.
__device__ int visited_insert(uint4* tab, uint64_t cap, uint4 key, uint64_t& where) {
  key = key_fix(key);
  uint64_t i = ((uint64_t)key.x * cap) >> 32;
  while (true) {
    uint4 v = tab[i];
    if ((v.x | v.y | v.z | v.w) == 0) {
      tab[i] = key;
      where = i;
      return 1;
    }
    if (key_eq(v, key)) return 0;
    if (++i == cap) i = 0;
  }
}

__device__ __forceinline__ uint4 warp_xor(uint4 k) {
  k.x = __reduce_xor_sync(kFull, k.x);
  k.y = __reduce_xor_sync(kFull, k.y);
  k.z = __reduce_xor_sync(kFull, k.z);
  k.w = __reduce_xor_sync(kFull, k.w);
  return k;
}

struct Counters {
  unsigned long long Q = 0, P = 0, X = 0, L = 0, S = 0, A = 0, V = 0;
};

// One search.  Returns the item status; on kItemOk the solution has been written.
__device__ uint8_t run_search(const SearchParams& p, uint32_t slot, uint32_t item, uint32_t sol_row, uint32_t lane, uint32_t* fbits,
                              uint32_t* rbits, uint16_t* st, uint32_t* sp_scratch, Counters& ctr, uint64_t& n_visited) {
  const DeviceTables& t = p.t;
  const uint32_t D = t.D, nwords = t.nwords;
  const uint32_t shot = item / p.n_trials, trial = item - shot * p.n_trials;
  const uint32_t perm = __ldg(p.trial_perm + trial);
  const uint32_t beam = __ldg(p.trial_beam + trial);
  const uint32_t* order_det = t.order_det + (size_t)perm * D;
  const uint32_t* order_pos = t.order_pos + (size_t)perm * D;
  HeapNode* heap = p.heap + (size_t)slot * p.node_cap;
  ChainNode* chain = p.chain + (size_t)slot * p.node_cap;
  uint4* vtab = p.no_revisit ? p.visited + (size_t)slot * p.visit_cap : nullptr;
  uint32_t* vlist = p.no_revisit ? p.vlist + (size_t)slot * p.vlist_cap : nullptr;
  double* hc = p.hcache + (size_t)slot * D;
  const double kInf = __longlong_as_double(0x7FF0000000000000LL);

  // Stage the shot: bit-packed syndrome -> root fired bitset.
  const uint8_t* row = p.dets + (size_t)shot * p.stride;
  const uint32_t nbytes = (D + 7) / 8;
  uint32_t my_fired = 0;
  uint4 root_key = make_uint4(0, 0, 0, 0);
  for (uint32_t w = lane; w < nwords; w += 32) {
    uint32_t word = 0;
    for (uint32_t b = 0; b < 4; ++b) {
      uint32_t at = w * 4 + b;
      if (at < nbytes) word |= (uint32_t)row[at] << (8 * b);
    }
    if (w * 32 + 32 > D) word &= (1u << (D - w * 32)) - 1u;
    rbits[w] = word;
    my_fired += __popc(word);
    for (uint32_t m = word; m; m &= m - 1) {
      uint4 z = ldg16(t.zobrist + (w * 32 + (uint32_t)__ffs((int)m) - 1));
      root_key.x ^= z.x;
      root_key.y ^= z.y;
      root_key.z ^= z.z;
      root_key.w ^= z.w;
    }
  }
  const uint32_t root_dets = __reduce_add_sync(kFull, my_fired);
  root_key = warp_xor(root_key);
  __syncwarp();
  // sparsify: this shot's active-entry masks (build_sparse_d2e, tesseract.cc:673-737)
  uint32_t* act = p.t.sp_on ? p.sp_act + (size_t)slot * D * t.sp_rw : nullptr;
  if (act) sparsify::select(t, rbits, act, sp_scratch, 0, 1, lane, [] { __syncwarp(); });

  if (p.trace && lane == 0) {
    p.trace_len[item] = 0;
    p.trace_arena[item] = (unsigned long long)(uintptr_t)chain;
  }
  auto log_node = [&](uint32_t chain_idx) {  // visualization log (tesseract.cc:442-445, 478-481)
    if (p.trace && lane == 0) {
      const uint32_t n = p.trace_len[item]++;
      if (n < p.trace_cap) p.trace[(size_t)item * p.trace_cap + n] = chain_idx;
    }
  };
  uint64_t heap_size = 0, chain_size = 0, pushed = 0;
  uint32_t min_dets = root_dets;
  uint32_t max_dets = root_dets + beam;
  bool root_pending = true;  // the root is "pushed" once its cost is known
  HeapNode node;

  while (true) {
    if (root_pending) {
      node.cost = 0;
      node.chain = kNoChain;
      node.num_dets = (uint16_t)root_dets;
      node.depth = 0;
    } else {
      if (heap_size == 0) return kItemLowConfidence;  // tesseract.cc:609-615
      if (lane == 0) node = heap_pop(heap, heap_size);
      heap_size = __shfl_sync(kFull, heap_size, 0) ;
      node.cost = __shfl_sync(kFull, node.cost, 0);
      node.chain = __shfl_sync(kFull, node.chain, 0);
      uint32_t packed = __shfl_sync(kFull, (uint32_t)node.num_dets | ((uint32_t)node.depth << 16), 0);
      node.num_dets = (uint16_t)(packed & 0xFFFF);
      node.depth = (uint16_t)(packed >> 16);
      ++ctr.P;
      if (node.num_dets > max_dets) continue;  // tesseract.cc:434
    }

    // Rebuild the node's state from its chain (tesseract.cc:349-369).
    for (uint32_t i = lane; i <= D; i += 32) st[i] = 0;
    for (uint32_t w = lane; w < nwords; w += 32) fbits[w] = rbits[w];
    __syncwarp();
    for (uint32_t w = lane; w < nwords; w += 32)
      for (uint32_t m = rbits[w]; m; m &= m - 1) st[w * 32 + (uint32_t)__ffs((int)m) - 1] = kFiredBit;
    __syncwarp();
    uint4 key = make_uint4(0, 0, 0, 0);
    for (uint32_t at = node.chain; at != kNoChain;) {
      const ChainNode cn = chain[at];
      const uint32_t eb = __ldg(t.e_off + cn.error), deg = __ldg(t.e_off + cn.error + 1) - eb;
      uint32_t dj = 0;
      if (lane < deg) {
        dj = (uint32_t)__ldg(t.e_det + eb + lane);
        st[dj] ^= kFiredBit;
        atomicXor(&fbits[dj >> 5], 1u << (dj & 31));
        uint4 z = ldg16(t.zobrist + dj);
        key.x ^= z.x;
        key.y ^= z.y;
        key.z ^= z.z;
        key.w ^= z.w;
      }
      __syncwarp();
      if (p.at_most_two && lane < deg && ((cn.off_mask >> lane) & 1)) st[dj] = (st[dj] & kFiredBit) | kThreshMask;
      __syncwarp();
      if (lane == 0) {
        const uint32_t mind = cn.mind_pos >> 16, thr = (cn.mind_pos & 0xFFFF) + 1;
        const uint16_t s = st[mind];
        if ((uint32_t)(s & kThreshMask) < thr) st[mind] = (uint16_t)((s & kFiredBit) | thr);
      }
      __syncwarp();
      at = cn.parent;
      ++ctr.L;
    }
    key = warp_xor(key);
    key.x ^= root_key.x;
    key.y ^= root_key.y;
    key.z ^= root_key.z;
    key.w ^= root_key.w;

    if (!root_pending) {
      if (node.num_dets == 0) {
        // Goal (tesseract.cc:441-473): chain errors in root -> leaf order.
        log_node(node.chain);
        if (node.depth > p.sol_stride) return kItemSolutionOverflow;
        if (lane == 0) {
          uint32_t* out = p.sol + (size_t)sol_row * p.sol_stride;
          uint32_t at = node.chain;
          for (uint32_t i = 0; i < node.depth; ++i) {
            const ChainNode cn = chain[at];
            out[node.depth - 1 - i] = cn.error;
            at = cn.parent;
          }
          p.sol_len[item] = node.depth;
          p.fcost[item] = node.cost;
        }
        return kItemOk;
      }
      if (p.no_revisit) {  // tesseract.cc:475-476
        int inserted = 0;
        uint64_t where = 0;
        if (lane == 0) {
          if ((n_visited + 1) * 2 > p.visit_cap) {
            inserted = -1;
          } else {
            inserted = visited_insert(vtab, p.visit_cap, key, where);
            if (inserted == 1 && n_visited < p.vlist_cap) vlist[n_visited] = (uint32_t)where;
          }
        }
        inserted = __shfl_sync(kFull, inserted, 0);
        ++ctr.V;
        if (inserted < 0) return kItemNodeOverflow;
        if (inserted == 0) continue;
        ++n_visited;
      }
      log_node(node.chain);
      if (node.num_dets < min_dets) {  // tesseract.cc:503-511
        min_dets = node.num_dets;
        max_dets = min(max_dets, min_dets + beam);
      }
    }

    // Heuristic terms of every fired detector in this state (the reference memoises them lazily in
    // detector_cost_cache, tesseract.cc:531,569-582; same values).  For the root their in-order sum is
    // the initial cost (tesseract.cc:411-414).
    double root_cost = 0;
    // (The root is popped right after its cost was computed: its terms are still in hc.)
    const bool terms_cached = !root_pending && node.chain == kNoChain;
    for (uint32_t wb = 0; wb < nwords && !terms_cached; wb += 32) {
      const uint32_t w = wb + lane;
      const uint32_t mine = w < nwords ? fbits[w] : 0;
      unsigned nz = __ballot_sync(kFull, mine != 0);
      while (nz) {
        const uint32_t l = (uint32_t)__ffs((int)nz) - 1;
        nz &= nz - 1;
        uint32_t word = __shfl_sync(kFull, mine, (int)l);
        while (word) {
          const uint32_t d = (wb + l) * 32 + (uint32_t)__ffs((int)word) - 1;
          word &= word - 1;
          const double h = detcost(t, st, act, d, p.det_penalty, lane, ctr.S);
          if (lane == 0) hc[d] = h;
          root_cost = __dadd_rn(root_cost, h);
        }
      }
    }
    __syncwarp();

    if (root_pending) {
      root_pending = false;
      if (root_cost == kInf) return kItemLowConfidence;  // tesseract.cc:416-419
      if (p.node_cap < 1) return kItemNodeOverflow;
      node.cost = root_cost;
      if (lane == 0) heap_push(heap, heap_size, node);
      heap_size = __shfl_sync(kFull, heap_size, 0);
      pushed = 1;
      ++ctr.Q;
      continue;  // the loop now pops it back (tesseract.cc:427-432)
    }
    ++ctr.X;

    // min_detector: first fired detector in this trial's order (tesseract.cc:522-528).
    uint32_t best_pos = 0xFFFFFFFFu;
    for (uint32_t w = lane; w < nwords; w += 32)
      for (uint32_t m = fbits[w]; m; m &= m - 1)
        best_pos = min(best_pos, __ldg(order_pos + (w * 32 + (uint32_t)__ffs((int)m) - 1)));
    best_pos = __reduce_min_sync(kFull, best_pos);
    const uint32_t mind = __ldg(order_det + best_pos);
    const uint32_t m_off = __ldg(t.row_off + mind), m_len = __ldg(t.row_off + mind + 1) - m_off;
    const uint32_t m_thr = st[mind] & kThreshMask;

    // Children: unblocked entries of row[min_detector], in row order (tesseract.cc:533-606).
    for (uint32_t base = 0; base < m_len; base += 32) {
      const uint32_t idx = base + lane;
      bool open = false;
      uint32_t my_err = 0;
      if (idx < m_len) {
        open = !decode_entry(t, st, m_off, idx, m_thr).blocked;
        if (act) open = open && ((__ldcg(act + (size_t)mind * t.sp_rw + (base >> 5)) >> lane) & 1u);
        my_err = (uint32_t)__ldg(t.row_err + m_off + idx);
      }
      unsigned todo = __ballot_sync(kFull, open);
      while (todo) {
        const uint32_t c = (uint32_t)__ffs((int)todo) - 1;
        todo &= todo - 1;
        const uint32_t pos = base + c;
        const uint32_t ei = __shfl_sync(kFull, my_err, (int)c);
        const uint32_t eb = __ldg(t.e_off + ei), deg = __ldg(t.e_off + ei + 1) - eb;
        uint32_t dj = D;
        uint16_t sj = 0;
        if (lane < deg) {
          dj = (uint32_t)__ldg(t.e_det + eb + lane);
          sj = st[dj];
        }
        const unsigned fired_mask = __ballot_sync(kFull, (sj & kFiredBit) != 0);
        const uint32_t next_dets = node.num_dets + deg - 2 * (uint32_t)__popc(fired_mask);
        if (next_dets > max_dets) continue;  // tesseract.cc:561
        if (p.no_revisit) {                  // tesseract.cc:563-565
          uint4 z = make_uint4(0, 0, 0, 0);
          if (lane < deg) z = ldg16(t.zobrist + dj);
          z = warp_xor(z);
          z.x ^= key.x;
          z.y ^= key.y;
          z.z ^= key.z;
          z.w ^= key.w;
          int seen = 0;
          if (lane == 0) seen = visited_contains(vtab, p.visit_cap, z) ? 1 : 0;
          seen = __shfl_sync(kFull, seen, 0);
          ++ctr.V;
          if (seen) continue;
        }
        // Patch the state into the child's: toggle the error's detectors, block the row prefix.
        if (lane < deg) {
          uint16_t nv = sj ^ kFiredBit;
          if (p.at_most_two && (sj & kFiredBit)) nv = (uint16_t)((nv & kFiredBit) | kThreshMask);
          st[dj] = nv;
        }
        __syncwarp();
        if (lane == 0) {
          const uint16_t s = st[mind];
          if ((uint32_t)(s & kThreshMask) < pos + 1) st[mind] = (uint16_t)((s & kFiredBit) | (pos + 1));
        }
        __syncwarp();

        // Child cost: the reference's add/sub chain, in its order (tesseract.cc:549-585).
        double cost = __dadd_rn(node.cost, __ldg(t.e_cost + ei));
        for (uint32_t j = 0; j < deg; ++j) {
          const uint32_t d = __shfl_sync(kFull, dj, (int)j);
          if ((fired_mask >> j) & 1) {
            cost = __dsub_rn(cost, hc[d]);
          } else {
            cost = __dadd_rn(cost, detcost(t, st, act, d, p.det_penalty, lane, ctr.S));
          }
        }
        // eneighbors[ei] ∩ fired: detectors adjacent to one of the error's detectors, fired in the
        // parent (hence in the child), not in the error itself; ascending.
        for (uint32_t wb = 0; wb < nwords; wb += 32) {
          const uint32_t w = wb + lane;
          uint32_t mine = 0;
          for (uint32_t j = 0; j < deg; ++j) {
            const uint32_t d = __shfl_sync(kFull, dj, (int)j);
            if (w < nwords) mine |= __ldg(t.adj + (size_t)d * nwords + w);
          }
          for (uint32_t j = 0; j < deg; ++j) {
            const uint32_t d = __shfl_sync(kFull, dj, (int)j);
            if ((d >> 5) == w) mine &= ~(1u << (d & 31));
          }
          if (w < nwords) mine &= fbits[w];
          unsigned nz = __ballot_sync(kFull, mine != 0);
          while (nz) {
            const uint32_t l = (uint32_t)__ffs((int)nz) - 1;
            nz &= nz - 1;
            uint32_t word = __shfl_sync(kFull, mine, (int)l);
            while (word) {
              const uint32_t od = (wb + l) * 32 + (uint32_t)__ffs((int)word) - 1;
              word &= word - 1;
              cost = __dsub_rn(cost, hc[od]);
              cost = __dadd_rn(cost, detcost(t, st, act, od, p.det_penalty, lane, ctr.S));
            }
          }
        }
        ctr.A += deg + __ldg(t.e_nnbr + ei);
        // Restore the parent state.
        if (lane < deg) st[dj] = sj;
        __syncwarp();

        if (cost == kInf) continue;  // tesseract.cc:587
        if (node.depth == 0xFFFF) return kItemSolutionOverflow;
        if (heap_size >= p.node_cap || chain_size >= p.node_cap || chain_size >= 0xFFFFFFFEull) return kItemNodeOverflow;
        if (lane == 0) {
          ChainNode cn;
          cn.error = ei;
          cn.parent = node.chain;
          cn.mind_pos = (mind << 16) | pos;
          cn.off_mask = fired_mask;
          chain[chain_size] = cn;
          HeapNode hn;
          hn.cost = cost;
          hn.chain = (uint32_t)chain_size;
          hn.num_dets = (uint16_t)next_dets;
          hn.depth = (uint16_t)(node.depth + 1);
          heap_push(heap, heap_size, hn);
        }
        heap_size = __shfl_sync(kFull, heap_size, 0);
        ++chain_size;
        ++pushed;
        ++ctr.Q;
        if (pushed > p.pqlimit) return kItemLowConfidence;  // tesseract.cc:599-605
      }
    }
  }
}

__global__ void __launch_bounds__(256) search_kernel(const SearchParams p) {
  extern __shared__ uint32_t smem[];
  const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t slot = blockIdx.x * (blockDim.x >> 5) + warp;
  if (slot >= p.n_slots) return;
  const uint32_t D = p.t.D, nwords = p.t.nwords;
  const uint32_t per_warp_words = 2 * nwords + (D + 2) / 2 + (p.t.sp_on ? sparsify::kScratchWords : 0);
  uint32_t* fbits = smem + (size_t)warp * per_warp_words;
  uint32_t* rbits = fbits + nwords;
  uint16_t* st = reinterpret_cast<uint16_t*>(rbits + nwords);
  uint32_t* sp_scratch = rbits + nwords + (D + 2) / 2;
  Counters ctr;
  unsigned long long searches = 0;
  while (true) {
    uint32_t work = 0;
    if (lane == 0) work = atomicAdd(p.work_counter, 1u);
    work = __shfl_sync(kFull, work, 0);
    if (work >= p.n_work) break;
    uint32_t item = p.item_ids ? __ldg(p.item_ids + work) : work;
    if (p.shot_order) item = __ldg(p.shot_order + work / p.n_trials) * p.n_trials + work % p.n_trials;
    uint64_t n_visited = 0;
    const uint8_t status = run_search(p, slot, item, p.sol_by_work ? work : item, lane, fbits, rbits, st, sp_scratch, ctr, n_visited);
    __syncwarp();
    if (p.t.sp_on)  // back to the mandatory masks for the slot's next search
      sparsify::restore(p.t, rbits, p.sp_act + (size_t)slot * D * p.t.sp_rw, 0, 1, lane, [] { __syncwarp(); });
    ++searches;
    if (lane == 0) {
      p.status[item] = status;
      if (status != kItemOk) p.sol_len[item] = 0;
    }
    // Empty the visited table for the next search of this slot.
    if (p.no_revisit && n_visited) {
      uint4* vtab = p.visited + (size_t)slot * p.visit_cap;
      const uint4 zero = make_uint4(0, 0, 0, 0);
      if (n_visited <= p.vlist_cap) {
        const uint32_t* vlist = p.vlist + (size_t)slot * p.vlist_cap;
        for (uint64_t i = lane; i < n_visited; i += 32) vtab[vlist[i]] = zero;
      } else {
        for (uint64_t i = lane; i < p.visit_cap; i += 32) vtab[i] = zero;
      }
    }
    __syncwarp();
  }
  if (lane == 0 && p.counters) {
    atomicAdd(p.counters + 0, ctr.Q);
    atomicAdd(p.counters + 1, ctr.P);
    atomicAdd(p.counters + 2, ctr.X);
    atomicAdd(p.counters + 3, ctr.L);
    atomicAdd(p.counters + 4, ctr.S);
    atomicAdd(p.counters + 5, ctr.A);
    atomicAdd(p.counters + 6, ctr.V);
    atomicAdd(p.counters + 7, searches);
  }
}

// ---- staging: per-batch syndrome statistics (sizes the solution buffers) ------------------------

__global__ void stage_kernel(const StageParams p) {
  const uint32_t shot = blockIdx.x * blockDim.x + threadIdx.x;
  uint32_t fired = 0;
  if (shot < p.shots) {
    const uint8_t* row = p.dets + (size_t)shot * p.stride;
    const uint32_t nbytes = (p.D + 7) / 8;
    for (uint32_t b = 0; b < nbytes; ++b) {
      uint32_t v = row[b];
      if (b * 8 + 8 > p.D) v &= (1u << (p.D - b * 8)) - 1u;
      fired += __popc(v);
    }
  }
  if (p.fired && shot < p.shots) p.fired[shot] = fired;
  const uint32_t wmax = __reduce_max_sync(kFull, fired);
  const uint32_t wsum = __reduce_add_sync(kFull, fired);
  if ((threadIdx.x & 31) == 0) {
    atomicMax(p.max_fired, wmax);
    atomicAdd(p.total_fired, (unsigned long long)wsum);
  }
}

// ---- order: shots by decreasing syndrome weight (counting sort) ------------------------------------

__global__ void order_hist_kernel(const OrderParams p) {
  const uint32_t shot = blockIdx.x * blockDim.x + threadIdx.x;
  if (shot < p.shots) atomicAdd(p.bin_cursor + p.fired[shot], 1u);
}
__global__ void order_scan_kernel(const OrderParams p) {  // one thread: bins is at most num_detectors + 1
  uint32_t run = 0;
  for (uint32_t b = p.bins; b-- > 0;) {
    const uint32_t n = p.bin_cursor[b];
    p.bin_cursor[b] = run;
    run += n;
  }
}
__global__ void order_scatter_kernel(const OrderParams p) {
  const uint32_t shot = blockIdx.x * blockDim.x + threadIdx.x;
  if (shot < p.shots) p.order[atomicAdd(p.bin_cursor + p.fired[shot], 1u)] = shot;
}

// ---- resolve: ensemble minimum + per-shot outputs (tesseract.cc:295-347, 618-658) --------------

__global__ void resolve_kernel(const ResolveParams p) {
  const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= p.n_shots) return;
  const uint32_t shot = p.shot_ids ? p.shot_ids[k] : k;
  const DeviceTables& t = p.t;
  double best = 1.7976931348623157e308;  // std::numeric_limits<double>::max()
  int64_t best_item = -1, best_row = -1;
  bool overflow = false;
  for (uint32_t u = 0; u < p.n_trials; ++u) overflow |= p.status[(size_t)shot * p.n_trials + u] >= kItemNodeOverflow;
  if (overflow) {
    if (p.retry_shots) p.retry_shots[atomicAdd(p.retry_count, 1u)] = shot;
    return;
  }
  const size_t row0 = p.sol_by_rank ? (size_t)k * p.n_trials : (size_t)shot * p.n_trials;
  for (uint32_t lt = 0; lt < p.n_literal; ++lt) {
    const size_t item = (size_t)shot * p.n_trials + p.literal_to_trial[lt];
    if (p.status[item] != kItemOk) continue;
    // cost_from_errors: sum in buffer order from 0.0 (tesseract.cc:618-628).
    const uint32_t* s = p.sol + (row0 + p.literal_to_trial[lt]) * p.sol_stride;
    const uint32_t n = p.sol_len[item];
    double c = 0;
    for (uint32_t i = 0; i < n; ++i) c = __dadd_rn(c, t.e_cost[s[i]]);
    if (c < best) {  // strict: the earliest trial wins ties (tesseract.cc:314-317)
      best = c;
      best_item = (int64_t)item;
      best_row = (int64_t)(row0 + p.literal_to_trial[lt]);
    }
  }
  const bool low = best_item < 0;
  if (p.low_conf_out) p.low_conf_out[shot] = low ? 1 : 0;
  if (p.cost_out) p.cost_out[shot] = low ? 0.0 : best;
  const uint32_t n = low ? 0 : p.sol_len[best_item];
  const uint32_t* s = low ? nullptr : p.sol + (size_t)best_row * p.sol_stride;
  if (p.obs_out) {
    uint8_t* o = p.obs_out + (size_t)shot * p.obs_stride;
    for (uint32_t b = 0; b < p.obs_stride; ++b) o[b] = 0;
    for (uint32_t i = 0; i < n; ++i) {
      const uint32_t e = s[i];
      for (uint32_t q = t.eo_off[e]; q < t.eo_off[e + 1]; ++q) {
        const uint32_t ob = (uint32_t)t.eo_obs[q];
        o[ob >> 3] ^= (uint8_t)(1u << (ob & 7));
      }
    }
  }
  if (p.err_off) {
    unsigned long long at = n ? atomicAdd(p.err_cursor, (unsigned long long)n) : 0;
    p.err_off[shot] = at;
    p.err_len[shot] = n;
    if (at + n <= p.err_cap)
      for (uint32_t i = 0; i < n; ++i) p.err_out[at + i] = t.e2dem[s[i]];
  }
}

}  // namespace

uint32_t search_smem_bytes_per_warp(uint32_t D, bool sparsify) {
  const uint32_t nwords = (D + 31) / 32;
  return 4 * (2 * nwords + (D + 2) / 2 + (sparsify ? sparsify::kScratchWords : 0));
}

int search_max_blocks_per_sm(uint32_t warps_per_block, uint32_t D, bool sparsify) {
  int n = 0;
  const size_t smem = (size_t)search_smem_bytes_per_warp(D, sparsify) * warps_per_block;
  if (smem > 48 * 1024)
    if (cudaFuncSetAttribute(search_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, search_kernel, (int)(warps_per_block * 32), smem) != cudaSuccess)
    return 0;
  return n;
}

int launch_stage(const StageParams& p, void* stream) {
  if (p.shots == 0) return 0;
  stage_kernel<<<(p.shots + 127) / 128, 128, 0, (cudaStream_t)stream>>>(p);
  return (int)cudaGetLastError();
}

int launch_order(const OrderParams& p, void* stream) {
  if (p.shots == 0) return 0;
  cudaStream_t s = (cudaStream_t)stream;
  order_hist_kernel<<<(p.shots + 255) / 256, 256, 0, s>>>(p);
  order_scan_kernel<<<1, 1, 0, s>>>(p);
  order_scatter_kernel<<<(p.shots + 255) / 256, 256, 0, s>>>(p);
  return (int)cudaGetLastError();
}

int launch_search(const SearchParams& p, uint32_t blocks, uint32_t warps_per_block, void* stream) {
  if (p.n_work == 0 || blocks == 0) return 0;
  const size_t smem = (size_t)search_smem_bytes_per_warp(p.t.D, p.t.sp_on != 0) * warps_per_block;
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(search_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
  }
  search_kernel<<<blocks, warps_per_block * 32, smem, (cudaStream_t)stream>>>(p);
  return (int)cudaGetLastError();
}

int launch_resolve(const ResolveParams& p, void* stream) {
  if (p.n_shots == 0) return 0;
  resolve_kernel<<<(p.n_shots + 127) / 128, 128, 0, (cudaStream_t)stream>>>(p);
  return (int)cudaGetLastError();
}

}  // namespace tsr
