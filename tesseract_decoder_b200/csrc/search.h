// Device-side data layout and launch interface of the A* search kernels (search.cu, search_fast.cu).
#pragma once
#include <cstdint>

namespace tsr {

// ---- device-resident DEM tables (read-only during decoding, L2-resident: tens of MB) --------------
struct DeviceTables {
  uint32_t D = 0, E = 0, O = 0;
  uint32_t nwords = 0;  // ceil(D/32): words per detector bitset and per adjacency row
  // Detector rows (host-sorted), and per-entry records; see ingest.h.
  const uint32_t* row_off = nullptr;  // [D+1]
  const int32_t* row_err = nullptr;   // [nnz]
  const uint4* rec_a = nullptr;       // [nnz]   {cost f64, size u16, flags u16, slot0}
  const uint4* rec_b = nullptr;       // [nnz]   slots 1..4
  const uint4* rec_c = nullptr;       // [nnz]   slots 5..8
  // Errors.
  const uint32_t* e_off = nullptr;    // [E+1]
  const int32_t* e_det = nullptr;     // [nnz]  ascending detectors
  const int32_t* e_rank = nullptr;    // [nnz]  rank of the error in that detector's row
  const double* e_cost = nullptr;     // [E]
  const uint16_t* e_nnbr = nullptr;   // [E]    |eneighbors[e]| of the reference (counter A only)
  const uint32_t* eo_off = nullptr;   // [E+1]
  const int32_t* eo_obs = nullptr;    // observables per error
  const uint64_t* e2dem = nullptr;    // [E]    error_to_dem_error
  // Detector adjacency bit matrix [D][nwords] and 128-bit Zobrist keys [D].
  const uint32_t* adj = nullptr;
  const uint4* zobrist = nullptr;
  // Distinct detector orders: position -> detector and its inverse, [n_perm][D].
  const uint32_t* order_det = nullptr;
  const uint32_t* order_pos = nullptr;
  // Fast path (fast.h; null / 0 when the DEM is not certified for it).
  const uint4* frec_a = nullptr;     // [nnz] packed records, plane A
  const uint4* frec_b = nullptr;     // [nnz] plane B (entries with header bit 7)
  const uint16_t* frank = nullptr;   // [fn_costs * fnstride] rank of class (cost id, count)
  const double* fquot = nullptr;     // [fn_costs * fnstride] fl(cost / count)
  uint32_t fa_slots = 0, fnstride = 0, fn_costs = 0;
  uint32_t max_row_len = 0;
  // Bit-sliced row tables (fast.h: BitTables; null / 0 when absent).
  const uint4* bs_meta = nullptr;       // [D][2] BitRowMeta
  const uint16_t* bs_adjpre = nullptr;  // [D][nwords]
  const uint32_t* bs_masks = nullptr;
  const uint32_t* bs_pl_off = nullptr;
  const uint32_t* bs_pl = nullptr;
  uint32_t bs_group = 0, bs_cap = 0;
  // --sparsify-errors (sparsify.cuh; null / 0 when off): per row word [D][sp_rw], bit i of word w = entry 32 w + i.
  const uint32_t* sp_mand = nullptr;   // errors of degree <= sparsify_base_degree: always active
  const uint32_t* sp_opt = nullptr;    // errors of base < degree <= max: reactivated per shot
  const uint32_t* sp_crank = nullptr;  // [E] dense rank of the error's cost among the distinct costs
  uint32_t sp_rw = 0;                  // words per row: ceil(max_row_len / 32)
  uint32_t sp_limit = 0;               // sparsify_reactivate_limit
  int sp_on = 0;
};

struct alignas(16) HeapNode {
  double cost;       // f = g + h
  uint32_t chain;    // index into the chain arena, kNoChain for the root
  uint16_t num_dets;
  uint16_t depth;
};
struct alignas(16) ChainNode {
  uint32_t error;
  uint32_t parent;
  uint32_t mind_pos;  // (min_detector << 16) | rank of `error` in that detector's row
  uint32_t off_mask;  // bit j: detector j of `error` was fired before it (at-most-two heuristic)
};
constexpr uint32_t kNoChain = 0xFFFFFFFFu;

enum ItemStatus : uint8_t {
  kItemOk = 0,
  kItemLowConfidence = 1,   // the reference's low_confidence_flag
  kItemNodeOverflow = 2,    // outgrew this tier's per-slot heap / chain / visited capacity
  kItemSolutionOverflow = 3, // solution longer than sol_stride
  kItemPaused = 4           // fast path: spent its expansion budget; to be resumed by a wider CTA (ResumeRec)
};

// A paused search (fast path).  Everything else a search owns — heap, chain arena, visited table, sparsify masks — stays
// in its slot of the node pool; the CTA-local state is rebuilt from the popped node's chain.
struct alignas(16) ResumeRec {
  uint32_t item, slot;
  uint32_t min_dets, max_dets;
  unsigned long long heap_size, chain_size, pushed, n_visited;
};

struct SearchParams {
  DeviceTables t;
  // Work: item i = (shot, distinct trial) with i = shot * n_trials + trial; `item_ids` (nullable)
  // restricts the launch to a list of items (retry tiers).
  const uint8_t* dets = nullptr;  // [shots][stride] bit-packed
  uint64_t stride = 0;
  uint32_t n_work = 0;
  const uint32_t* item_ids = nullptr;
  // Tier 0 only (nullable): shots in decreasing order of fired detectors — the k-th claimed work unit is
  // (shot_order[k / n_trials], k % n_trials).  Search cost grows with the syndrome weight and is heavy-tailed, so the
  // longest searches start first instead of setting the tail of the batch.
  const uint32_t* shot_order = nullptr;
  uint32_t n_trials = 1;
  const uint32_t* trial_perm = nullptr;  // [n_trials] distinct order id
  const uint32_t* trial_beam = nullptr;  // [n_trials]
  // Config.
  double det_penalty = 0;
  uint64_t pqlimit = 0;
  int no_revisit = 1;
  int at_most_two = 0;
  // Per-slot scratch (slot = one warp).
  uint32_t n_slots = 0;
  uint64_t node_cap = 0;   // heap and chain entries per slot
  uint64_t visit_cap = 0;  // visited-table entries per slot
  uint32_t vlist_cap = 0;
  HeapNode* heap = nullptr;
  ChainNode* chain = nullptr;
  uint4* visited = nullptr;
  uint32_t* vlist = nullptr;
  uint32_t* sp_act = nullptr;  // [n_slots][D][sp_rw] active-entry masks (sparsify.cuh), == sp_mand between searches
  double* hcache = nullptr;  // [n_slots][D]
  int hc_shared = 0;         // fast path: the per-detector terms of the expanded node live in shared memory
  // Per-item outputs.
  uint32_t sol_stride = 0;
  int sol_by_work = 0;          // retry tiers: the solution row is the work index (position in item_ids), not the item id
  uint32_t* sol = nullptr;      // [items][sol_stride] compiled error ids, root -> leaf
  uint32_t* sol_len = nullptr;  // [items]
  uint8_t* status = nullptr;    // [items] ItemStatus
  double* fcost = nullptr;      // [items] f of the goal node (diagnostic)
  unsigned int* work_counter = nullptr;
  unsigned long long* counters = nullptr;  // [8], see tsr_counters
  unsigned long long* prof = nullptr;      // [8] nullable, fast path only: per-phase ticks, see tsr_profile
  // Visited-set audit (tsr_set_visited_audit; fast path).  The visited set stores 128-bit Zobrist fingerprints of fired
  // sets, not the sets themselves (the reference keeps the bitsets: tesseract.cc:394, 475-476, 563-565).  With the audit
  // on, every table entry also remembers the chain index of the node that inserted it, and every fingerprint HIT is
  // checked against the truth: the hit node's fired set is rebuilt from its chain and compared bit for bit with the
  // candidate's.  audit[0] counts the hits checked, audit[1] the hits whose sets differ (a collision; never observed).
  uint32_t* vchain = nullptr;            // [n_slots][visit_cap] chain index per table entry
  unsigned long long* audit = nullptr;   // [2]
  uint32_t* audit_scratch = nullptr;     // [n_slots][32 warps][nwords]
  // Visualization log (debug decodes of one shot; visualization.cc, tesseract.cc:442-445, 478-481): per item, the chain
  // index of every node the reference would log — each expanded node and the goal — in pop order, and the address of
  // the item's chain arena.  Null = off.
  uint32_t* trace = nullptr;          // [items][trace_cap]
  uint32_t trace_cap = 0;
  uint32_t* trace_len = nullptr;      // [items] entries the search wanted to log (may exceed trace_cap)
  unsigned long long* trace_arena = nullptr;  // [items]
  // Tail widening (fast path).  A launch with expansion_budget > 0 pauses every search that has expanded more nodes
  // than that once the work queue is empty: its record goes to paused[atomicAdd(paused_count)] and the CTA retires (its
  // slot stays occupied).  A launch with resume = 1 runs no new work: unit b continues resume_list[b], [b + units], ...
  // with four times the warps — to the end, or (budget > 0) for another `expansion_budget` expansions each, parking the
  // survivors in its own `paused` list for the next, wider stage.  Heavy-tailed batches then end with a few clusters
  // instead of one narrow CTA.
  uint32_t expansion_budget = 0;
  // Wide launches only: thread-block clusters of this many CTAs (1 = off).  CTA 0 of a cluster owns the search; during
  // the children phase of an expansion the other CTAs claim children from its counter, read the node's state from its
  // shared memory (distributed shared memory) and write their results there.  A lone search is bound by the
  // instruction issue rate of ONE SM (profiles/r02_f_*: 2.2 M warp instructions per expansion on the colour code);
  // a cluster spreads the children over up to 8 SMs.
  uint32_t cluster = 1;
  int resume = 0;                           // this launch continues resume_list[0 .. resume_n) instead of claiming work
  const ResumeRec* resume_list = nullptr;
  uint32_t resume_n = 0;
  ResumeRec* paused = nullptr;              // where this launch parks the searches it pauses (may be null when budget == 0)
  unsigned int* paused_count = nullptr;
  uint32_t paused_cap = 0;
};

struct ResolveParams {
  DeviceTables t;
  uint32_t n_shots = 0;          // shots of the batch (or length of shot_ids)
  const uint32_t* shot_ids = nullptr;  // nullable: restrict to these shots (retry tiers)
  uint32_t n_trials = 1;         // distinct trials per shot
  uint32_t n_literal = 1;        // literal trials of the reference loop
  const uint32_t* literal_to_trial = nullptr;  // [n_literal]
  uint32_t sol_stride = 0;
  int sol_by_rank = 0;           // retry tiers: the row of (k-th listed shot, trial u) is k * n_trials + u
  const uint32_t* sol = nullptr;
  const uint32_t* sol_len = nullptr;
  const uint8_t* status = nullptr;
  // Outputs (any may be null).
  uint8_t* obs_out = nullptr;  // [shots][obs_stride]
  uint32_t obs_stride = 0;
  double* cost_out = nullptr;
  uint8_t* low_conf_out = nullptr;
  uint64_t* err_out = nullptr;           // pooled predicted errors (original DEM indices)
  uint64_t err_cap = 0;
  unsigned long long* err_cursor = nullptr;
  uint64_t* err_off = nullptr;           // [shots] start in err_out
  uint32_t* err_len = nullptr;           // [shots]
  uint32_t* retry_shots = nullptr;       // shots with an overflowed trial
  unsigned int* retry_count = nullptr;
};

struct StageParams {
  const uint8_t* dets = nullptr;
  uint64_t stride = 0;
  uint32_t shots = 0;
  uint32_t D = 0;
  unsigned int* max_fired = nullptr;  // out: largest number of fired detectors in the batch
  unsigned long long* total_fired = nullptr;
  uint32_t* fired = nullptr;          // out (nullable): fired detectors per shot
};

struct OrderParams {
  const uint32_t* fired = nullptr;  // [shots]
  uint32_t shots = 0;
  uint32_t bins = 0;                // max fired + 1
  uint32_t* bin_cursor = nullptr;   // [bins], zeroed by the caller; scratch
  uint32_t* order = nullptr;        // out [shots]: shots by decreasing fired count
};

// Launchers (search.cu).  All enqueue on `stream` and return the cudaError_t of the launch.
int launch_stage(const StageParams& p, void* stream);
int launch_order(const OrderParams& p, void* stream);  // counting sort: three small kernels
int launch_search(const SearchParams& p, uint32_t blocks, uint32_t warps_per_block, void* stream);
int launch_resolve(const ResolveParams& p, void* stream);
// Fast path (search_fast.cu): one CTA of `warps` warps per search.
// `bits` selects the bit-sliced row scan (needs the bs_* tables); otherwise rows are scanned entry per lane, and
// `instrumented` additionally counts the reference's visited row entries.
// `latency` selects the variant tuned for a lone search per SM (wide CTAs: tail and retry tiers; entry-per-lane scan only).
int launch_search_fast(const SearchParams& p, uint32_t blocks, uint32_t warps, bool instrumented, bool bits, bool latency, void* stream);
uint32_t search_fast_smem_bytes(const DeviceTables& t, uint32_t warps, bool hc_shared, bool bits, bool latency = false);  // t.sp_on adds the selection scratch
int search_fast_max_blocks_per_sm(const DeviceTables& t, uint32_t warps, bool hc_shared, bool bits, bool latency = false);
// Shared memory bytes one search warp needs for a DEM with D detectors (sparsify adds the selection scratch).
uint32_t search_smem_bytes_per_warp(uint32_t D, bool sparsify);
// Resident blocks per SM of the search kernel for this block shape (0 on failure).
int search_max_blocks_per_sm(uint32_t warps_per_block, uint32_t D, bool sparsify);

}  // namespace tsr
