// C ABI (include/tesseract_b200.h): host-side decoder object, device memory and the launch pipeline.
//
// Per batch chunk:  stage_kernel (fired detectors per shot)  ->  counting sort of the shots by decreasing
// syndrome weight  ->  search kernel (search_fast.cu: one CTA per (shot, trial) for DEMs certified by fast.h;
// search.cu: one warp per (shot, trial) otherwise; persistent CTAs / warps claiming work from an atomic
// counter, like utils.h:88-92)  ->  resolve_kernel (ensemble minimum, observables, costs, pooled error
// lists).  Searches that outgrow their slot's heap/chain/visited capacity are re-run in the next "tier":
// fewer slots over the same node pool, each proportionally larger — never on the CPU.
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <iomanip>
#include <limits>
#include <sstream>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/tesseract_b200.h"
#include "circuit.h"
#include "dem.h"
#include "fast.h"
#include "ingest.h"
#include "search.h"

#include "handle.h"

using namespace tsr_impl;

namespace tsr_impl {
std::string& last_error() {
  thread_local std::string e;
  return e;
}
}  // namespace tsr_impl

tsr_decoder::~tsr_decoder() {
  tsr_impl::comm_release(*this);
  for (auto& e : ev)
    if (e) cudaEventDestroy(e);
  if (stream) cudaStreamDestroy(stream);
}

namespace {

using tsr::DemModel;

uint64_t splitmix(uint64_t& s) {
  s += 0x9E3779B97F4A7C15ull;
  uint64_t z = s;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}

// suggest_sparsify_reactivate_limit_capped (tesseract.cc:45-69): round(4.5^(base-2) * D / 3), capped.
int suggest_limit_capped(uint64_t num_detectors, int base_degree, int max_limit) {
  if (base_degree < 0) throw std::invalid_argument("sparsify_base_degree must be >= 0.");
  if (num_detectors == 0 || max_limit <= 0) return 0;
  const double exponent = (double)base_degree - 2.0, max_result = (double)max_limit;
  const double log_result = exponent * std::log(4.5) - std::log(3.0) + std::log((double)num_detectors);
  if (!std::isfinite(log_result) || log_result >= std::log(max_result)) return max_limit;
  const double result = std::exp(log_result);
  if (!std::isfinite(result)) return max_limit;
  const double rounded = std::round(result);
  if (rounded >= max_result) return max_limit;
  return (int)rounded;
}

void make_ensemble(const tsr_decoder& h, TrialSet& ts) {
  // tesseract.cc:307-344.
  std::map<std::pair<uint32_t, uint32_t>, uint32_t> seen;
  auto add = [&](size_t order, uint32_t beam) {
    auto key = std::make_pair(h.order_to_perm[order], beam);
    auto it = seen.find(key);
    if (it == seen.end()) {
      it = seen.emplace(key, (uint32_t)ts.perm.size()).first;
      ts.perm.push_back(key.first);
      ts.beam.push_back(beam);
    }
    ts.literal_to_trial.push_back(it->second);
  };
  const size_t n_orders = h.det_orders.size();
  if (h.cfg.beam_climbing) {
    const int trials = std::max(h.cfg.det_beam + 1, (int)n_orders);
    int beam = 0;
    size_t order = 0;
    for (int t = 0; t < trials; ++t) {
      add(order, (uint32_t)beam);
      beam = (beam + 1) % (h.cfg.det_beam + 1);
      order = (order + 1) % n_orders;
    }
  } else {
    for (size_t order = 0; order < n_orders; ++order) add(order, (uint32_t)h.cfg.det_beam);
  }
}

void init_device(tsr_decoder& h) {
  if (h.cfg.device >= 0) {
    CUDA_OK(cudaSetDevice(h.cfg.device));
    h.device = h.cfg.device;
  } else {
    CUDA_OK(cudaGetDevice(&h.device));
  }
  cudaDeviceProp prop{};
  CUDA_OK(cudaGetDeviceProperties(&prop, h.device));
  if (prop.major < 10)
    throw CudaError(std::string("tesseract_b200 needs an sm_100 (Blackwell) device; found ") + prop.name + " (sm_" +
                    std::to_string(prop.major) + std::to_string(prop.minor) + ")");
  h.sm_count = prop.multiProcessorCount;
  CUDA_OK(cudaStreamCreateWithFlags(&h.stream, cudaStreamNonBlocking));
  for (auto& e : h.ev) CUDA_OK(cudaEventCreate(&e));

  const tsr::DemTables& T = h.T;
  const uint32_t D = (uint32_t)T.num_detectors, E = (uint32_t)T.num_errors;
  h.row_off.upload(T.row_off);
  h.row_err.upload(T.row_err);
  h.rec_a.upload(T.rec_a);
  h.rec_b.upload(T.rec_b);
  h.rec_c.upload(T.rec_c);
  h.e_off.upload(T.e_off);
  h.e_det.upload(T.e_det);
  h.e_rank.upload(T.e_rank);
  h.e_cost.upload(T.e_cost);
  h.eo_off.upload(T.eo_off);
  h.eo_obs.upload(T.eo_obs);
  h.e2dem.upload(T.error_to_dem_error);
  h.adj.upload(T.adj);
  {
    // |eneighbors[e]| for counter A, from the adjacency matrix.
    std::vector<uint16_t> nn(E);
    std::vector<uint32_t> acc(T.adj_words);
    for (uint32_t e = 0; e < E; ++e) {
      std::fill(acc.begin(), acc.end(), 0u);
      for (int d : T.errors[e].detectors)
        for (uint32_t w = 0; w < T.adj_words; ++w) acc[w] |= T.adj[(size_t)d * T.adj_words + w];
      uint32_t n = 0;
      for (uint32_t w : acc) n += (uint32_t)__builtin_popcount(w);
      nn[e] = (uint16_t)std::min<uint32_t>(n - (uint32_t)T.errors[e].detectors.size(), 0xFFFF);
    }
    h.e_nnbr.upload(nn);
  }
  {
    std::vector<uint64_t> z((size_t)D * 2 + 2);
    uint64_t s = 0x7E55E7AC7DEC0DE5ull;
    for (auto& v : z) v = splitmix(s);
    h.zobrist.upload(z);
  }
  {
    std::vector<uint32_t> od((size_t)h.n_perms * D + 1), op((size_t)h.n_perms * D + 1, 0xFFFFFFFFu);
    std::vector<char> done(h.n_perms, 0);
    for (size_t k = 0; k < h.det_orders.size(); ++k) {
      uint32_t q = h.order_to_perm[k];
      if (done[q]) continue;
      done[q] = 1;
      for (uint32_t pos = 0; pos < D; ++pos) {
        uint64_t det = h.det_orders[k][pos];
        if (det >= D) throw std::invalid_argument("detector order entries must be < num_detectors");
        od[(size_t)q * D + pos] = (uint32_t)det;
        uint32_t& inv = op[(size_t)q * D + det];
        inv = std::min(inv, pos);
      }
    }
    h.order_det.upload(od);
    h.order_pos.upload(op);
  }
  tsr::DeviceTables& dt = h.dt;
  dt.D = D;
  dt.E = E;
  dt.O = (uint32_t)T.num_observables;
  dt.nwords = T.adj_words;
  dt.row_off = h.row_off.as<uint32_t>();
  dt.row_err = h.row_err.as<int32_t>();
  dt.rec_a = h.rec_a.as<uint4>();
  dt.rec_b = h.rec_b.as<uint4>();
  dt.rec_c = h.rec_c.as<uint4>();
  dt.e_off = h.e_off.as<uint32_t>();
  dt.e_det = h.e_det.as<int32_t>();
  dt.e_rank = h.e_rank.as<int32_t>();
  dt.e_cost = h.e_cost.as<double>();
  dt.e_nnbr = h.e_nnbr.as<uint16_t>();
  dt.eo_off = h.eo_off.as<uint32_t>();
  dt.eo_obs = h.eo_obs.as<int32_t>();
  dt.e2dem = h.e2dem.as<uint64_t>();
  dt.adj = h.adj.as<uint32_t>();
  dt.zobrist = h.zobrist.as<uint4>();
  dt.order_det = h.order_det.as<uint32_t>();
  dt.order_pos = h.order_pos.as<uint32_t>();
  dt.max_row_len = std::max<uint32_t>(T.max_row_len, 1);
  if (h.fast) {
    h.frec_a.upload(h.F.rec_a);
    h.frec_b.upload(h.F.rec_b);
    h.frank.upload(h.F.rank);
    h.fquot.upload(h.F.quotient);
    dt.frec_a = h.frec_a.as<uint4>();
    dt.frec_b = h.frec_b.as<uint4>();
    dt.frank = h.frank.as<uint16_t>();
    dt.fquot = h.fquot.as<double>();
    dt.fa_slots = h.F.a_slots;
    dt.fnstride = h.F.nstride;
    dt.fn_costs = h.F.n_costs;
    if (h.bits) {
      h.bs_meta.upload(h.B.rowmeta);
      h.bs_adjpre.upload(h.B.adjpre);
      h.bs_masks.upload(h.B.masks);
      h.bs_pl_off.upload(h.B.pl_off);
      h.bs_pl.upload(h.B.pl);
      dt.bs_meta = h.bs_meta.as<uint4>();
      dt.bs_adjpre = h.bs_adjpre.as<uint16_t>();
      dt.bs_masks = h.bs_masks.as<uint32_t>();
      dt.bs_pl_off = h.bs_pl_off.as<uint32_t>();
      dt.bs_pl = h.bs_pl.as<uint32_t>();
      dt.bs_group = h.B.group;
      dt.bs_cap = (h.B.max_nbrs + 7) & ~7u;
    }
  }

  if (h.cfg.sparsify_errors) {
    // sparsify.cuh: per row word, which entries are mandatory / optional (tesseract.cc:279-289), and dense cost ranks.
    const uint32_t rw = (dt.max_row_len + 31) / 32;
    std::vector<uint32_t> mand((size_t)D * rw + 1, 0u), opt((size_t)D * rw + 1, 0u);
    for (uint32_t d = 0; d < D; ++d)
      for (uint32_t k = T.row_off[d]; k < T.row_off[d + 1]; ++k) {
        const int degree = (int)T.errors[(size_t)T.row_err[k]].detectors.size();
        const uint32_t idx = k - T.row_off[d];
        if (degree <= h.cfg.sparsify_base_degree) mand[(size_t)d * rw + (idx >> 5)] |= 1u << (idx & 31);
        else if (h.cfg.sparsify_max_degree == -1 || degree <= h.cfg.sparsify_max_degree) opt[(size_t)d * rw + (idx >> 5)] |= 1u << (idx & 31);
      }
    std::vector<double> distinct(T.e_cost);
    std::sort(distinct.begin(), distinct.end());
    distinct.erase(std::unique(distinct.begin(), distinct.end()), distinct.end());
    if (distinct.size() >= (size_t{1} << 24)) throw std::invalid_argument("tesseract_b200: sparsify_errors supports fewer than 2^24 distinct costs");
    std::vector<uint32_t> crank(E + 1, 0u);
    for (uint32_t e = 0; e < E; ++e)
      crank[e] = (uint32_t)(std::lower_bound(distinct.begin(), distinct.end(), T.e_cost[e]) - distinct.begin());
    h.sp_mand.upload(mand);
    h.sp_opt.upload(opt);
    h.sp_crank.upload(crank);
    dt.sp_mand = h.sp_mand.as<uint32_t>();
    dt.sp_opt = h.sp_opt.as<uint32_t>();
    dt.sp_crank = h.sp_crank.as<uint32_t>();
    dt.sp_rw = rw;
    dt.sp_limit = (uint32_t)std::min<int64_t>(h.sparsify_limit, (int64_t)E);
    dt.sp_on = 1;
  }

  make_ensemble(h, h.ensemble);
  h.ensemble.upload();

  // Grid shape: persistent CTAs (fast path: one search each) or warps (generic path) filling every SM.
  uint32_t slots = 0;
  if (h.fast) {
    // Warps per search: enough to split a row's children, few enough that short rows do not idle them.
    // (measured: 8 warps beat 4 by 30 % on surface d=11 where searches are long, lose 6 % on surface d=5)
    uint32_t w = h.cfg.warps_per_block ? h.cfg.warps_per_block : ((T.max_row_len > 96 || D > 512) ? 8 : 4);
    w = std::max<uint32_t>(1, std::min<uint32_t>(w, 16));
    h.hc_shared = (size_t)D * 8 <= 32 * 1024;
    auto fit = [&](uint32_t& w_, bool& hc_, bool bits_) {
      int o = tsr::search_fast_max_blocks_per_sm(dt, w_, hc_, bits_);
      while (o == 0 && (w_ > 1 || hc_)) {
        if (hc_ && w_ <= 2) hc_ = false; else w_ = std::max<uint32_t>(w_ / 2, 1);
        o = tsr::search_fast_max_blocks_per_sm(dt, w_, hc_, bits_);
      }
      return o;
    };
    int occ = 0;
    if (h.bits) {
      uint32_t wb = w;
      bool hcb = h.hc_shared;
      occ = fit(wb, hcb, true);
      if (occ == 0) {
        h.bits = false;
      } else {
        w = wb;
        h.hc_shared = hcb;
        // the entry-per-lane variants (instrumented counters) share the block shape; they need less shared memory
      }
    }
    if (!h.bits) occ = fit(w, h.hc_shared, false);
    if (occ == 0) {
      h.fast = false;  // state does not fit in shared memory: generic kernel
      h.F.why_not = "the per-search state does not fit in shared memory";
    } else {
      h.wpb = w;
      if (const char* e = std::getenv("TSR_CTAS_PER_SM")) occ = std::max(1, std::min(occ, std::atoi(e)));  // experiments
      h.blocks = (uint32_t)(h.sm_count * occ);
      slots = h.blocks;
      if (h.cfg.search_slots && h.cfg.search_slots < slots) slots = h.blocks = std::max<uint32_t>(h.cfg.search_slots, 1);
      // Wide CTAs for the tail: the most warps (<= 32) whose state still fits an SM.
      uint32_t wide = 32;
      if (const char* e = std::getenv("TSR_WIDE_WARPS")) wide = (uint32_t)std::max(0, std::atoi(e));
      int occ_wide = 0;
      while (wide > w && (occ_wide = tsr::search_fast_max_blocks_per_sm(dt, wide, h.hc_shared, h.bits, true)) == 0) wide /= 2;
      if (wide > w && occ_wide > 0) {
        h.wpb_wide = wide;
        h.blocks_wide = (uint32_t)(h.sm_count * occ_wide);
        h.expansion_budget = 256;
        if (const char* e = std::getenv("TSR_EXPANSION_BUDGET")) h.expansion_budget = (uint32_t)std::max(0, std::atoi(e));
        // Clusters pay where an expansion has several CTAs' worth of children AND the children dominate it: the
        // entry-per-lane path on long rows (colour codes: rows of ~177, lone 1.2 M-push search 3.05 s -> 1.50 s with 8
        // CTAs).  Measured slower with them (profiles/r02_l_cluster_stage_sweep.txt): surface d=11 (rows of ~60; its long
        // searches are bound by pops and chain rebuilds, 1.74 s -> 2.11 s) and the bit-sliced path (BB144 long-beam
        // 2.93 s -> 3.56 s) — those keep single wide CTAs.  With clusters the tail runs in stages of 1024 expansions.
        const double mean_row = (double)T.row_err.size() / std::max<uint32_t>(D, 1);
        h.cluster_max = (!h.bits && mean_row >= 128.0) ? 8 : 1;
        if (const char* e = std::getenv("TSR_CLUSTER")) h.cluster_max = (uint32_t)std::min(std::max(1, std::atoi(e)), 8);
        h.stage_budget = h.cluster_max > 1 ? 1024 : 0;
        if (const char* e = std::getenv("TSR_STAGE_BUDGET")) h.stage_budget = (uint32_t)std::max(0, std::atoi(e));
      }
    }
  }
  if (!h.fast) {
    h.wpb = h.cfg.warps_per_block ? std::min<uint32_t>(h.cfg.warps_per_block, 8) : 4;
    int occ = tsr::search_max_blocks_per_sm(h.wpb, D, dt.sp_on != 0);
    while (occ == 0 && h.wpb > 1) {
      h.wpb /= 2;
      occ = tsr::search_max_blocks_per_sm(h.wpb, D, dt.sp_on != 0);
    }
    if (occ == 0) throw CudaError("the search kernel does not fit on this device for this many detectors");
    h.blocks = (uint32_t)(h.sm_count * occ);
    slots = h.blocks * h.wpb;
    if (h.cfg.search_slots && h.cfg.search_slots < slots) {
      slots = std::max<uint32_t>(h.cfg.search_slots, 1);
      h.blocks = (slots + h.wpb - 1) / h.wpb;
      slots = h.blocks * h.wpb;
    }
  }

  // Node pool: 32 B per node (heap + chain) + 16 B per visited entry (half as many as nodes).
  const bool revisit_tab = h.cfg.no_revisit_dets != 0;
  const double per_node = 32.0 + (revisit_tab ? 8.0 : 0.0);
  const uint64_t want_nodes = h.cfg.pqlimit == TSR_PQLIMIT_UNBOUNDED ? ~uint64_t{0} : h.cfg.pqlimit + 1;
  size_t free_b = 0, total_b = 0;
  CUDA_OK(cudaMemGetInfo(&free_b, &total_b));
  uint64_t budget = h.cfg.node_pool_bytes ? h.cfg.node_pool_bytes : (uint64_t)(free_b * 0.30);
  budget = std::min<uint64_t>(budget, (uint64_t)(free_b * 0.85));
  uint64_t cap0 = (uint64_t)((double)budget / slots / per_node);
  cap0 = std::min<uint64_t>(cap0, want_nodes);
  cap0 = std::max<uint64_t>(cap0, 64);
  auto visit_of = [&](uint64_t cap) { return revisit_tab ? std::max<uint64_t>(cap / 2, 64) : 0; };
  h.tiers.push_back({slots, cap0, visit_of(cap0)});
  const uint64_t pool_nodes = (uint64_t)slots * cap0;
  while (h.tiers.back().node_cap < want_nodes && h.tiers.back().slots > 1) {
    uint32_t s = std::max<uint32_t>(h.tiers.back().slots / 16, 1);
    uint64_t cap = std::min<uint64_t>(pool_nodes / s, want_nodes);
    h.tiers.push_back({s, cap, visit_of(cap)});
  }
  // Layout inside the pool for tier k: [heap: slots*cap*16][chain: slots*cap*16][visited: slots*vcap*16].
  size_t pool_bytes = 0;
  for (const Tier& t : h.tiers)
    pool_bytes = std::max<size_t>(pool_bytes, (size_t)t.slots * (t.node_cap * 32 + t.visit_cap * 16));
  h.pool.ensure(pool_bytes + 256);
  h.hcache.ensure((size_t)slots * std::max<uint32_t>(D, 1) * sizeof(double));
  h.vlist.ensure((size_t)slots * h.vlist_cap * sizeof(uint32_t));
  if (dt.sp_on) {  // every slot's active-entry masks start as (and are restored to) the mandatory masks
    const size_t per_slot = (size_t)D * dt.sp_rw * 4;
    h.sp_act.ensure(std::max<size_t>((size_t)slots * per_slot, 16));
    DevBuf::uploads_done();  // sp_mand was uploaded through the legacy stream
    for (uint32_t sl = 0; sl < slots && per_slot; ++sl)
      CUDA_OK(cudaMemcpyAsync(h.sp_act.as<char>() + (size_t)sl * per_slot, h.sp_mand.p, per_slot, cudaMemcpyDeviceToDevice, h.stream));
  }
  CUDA_OK(cudaMemsetAsync(h.pool.p, 0, h.pool.bytes, h.stream));  // visited tables start empty
  h.d_scalars.ensure(64);
  h.d_paused.ensure(std::max<size_t>((size_t)slots * sizeof(tsr::ResumeRec), 64));
  h.d_paused2.ensure(std::max<size_t>((size_t)slots * sizeof(tsr::ResumeRec), 64));
  h.d_stage_counts.ensure(64);
  h.d_counters.ensure(8 * sizeof(unsigned long long));
  CUDA_OK(cudaMemsetAsync(h.d_counters.p, 0, 64, h.stream));
  h.d_prof.ensure(8 * sizeof(unsigned long long));
  CUDA_OK(cudaMemsetAsync(h.d_prof.p, 0, 64, h.stream));
  CUDA_OK(cudaStreamSynchronize(h.stream));
  DevBuf::uploads_done();  // the table uploads above went through the legacy stream
  h.max_batch = h.cfg.max_batch_shots ? h.cfg.max_batch_shots : (uint64_t{1} << 18);
}

}  // namespace
void tsr_impl::require_device(const tsr_decoder& h) {
  if (h.host_only) throw CudaError("CUDA: this decoder was created host-only; decoding needs an sm_100 GPU (no CPU fallback)");
}

namespace {
// d_scalars layout (4-byte units): [0] work counter, [1] max fired, [2..3] total fired (u64),
// [4..5] error cursor (u64), [6] retry count.
struct Scalars {
  unsigned int work, max_fired;
  unsigned long long total_fired, err_cursor;
  unsigned int retry_count, paused_count;
};

// Decodes `shots` shots whose bit-packed syndromes are at device pointer d_dets.  Outputs go to the
// given device pointers (nullable).  Error lists are appended to h.last_off / h.last_idx.
void run_chunk(tsr_decoder& h, const uint8_t* d_dets, uint64_t stride, uint32_t shots, const TrialSet& ts,
               uint8_t* d_obs, double* d_cost, uint8_t* d_low) {
  const tsr::DemTables& T = h.T;
  const uint32_t U = (uint32_t)ts.perm.size();
  const uint64_t items64 = (uint64_t)shots * U;
  if (items64 > 0xFFFFFFF0ull) throw std::invalid_argument("batch too large: shots x trials must fit in 32 bits");
  const uint32_t items = (uint32_t)items64;
  cudaStream_t s = h.stream;
  Scalars* sc = h.d_scalars.as<Scalars>();
  CUDA_OK(cudaMemsetAsync(sc, 0, sizeof(Scalars), s));
  CUDA_OK(cudaEventRecord(h.ev[0], s));

  tsr::StageParams sp;
  sp.dets = d_dets;
  sp.stride = stride;
  sp.shots = shots;
  sp.D = (uint32_t)T.num_detectors;
  sp.max_fired = &sc->max_fired;
  sp.total_fired = &sc->total_fired;
  h.d_fired.ensure((size_t)shots * 4);
  sp.fired = h.d_fired.as<uint32_t>();
  CUDA_OK((cudaError_t)tsr::launch_stage(sp, s));
  Scalars hs{};
  CUDA_OK(cudaMemcpyAsync(&hs, sc, sizeof(Scalars), cudaMemcpyDeviceToHost, s));
  CUDA_OK(cudaStreamSynchronize(s));
  const uint32_t stride_need =
      (uint32_t)std::min<uint64_t>(std::max<uint64_t>(T.num_errors, 1), 4ull * hs.max_fired + 32);
  // Bound the per-chunk solution buffer (items x stride x 4 B) to 2 GiB: split the chunk when it would not fit.
  if ((uint64_t)items * stride_need * 4 > (uint64_t{2} << 30) && shots > 1024) {
    const uint32_t half = shots / 2;
    const uint64_t ob = (T.num_observables + 7) / 8;
    run_chunk(h, d_dets, stride, half, ts, d_obs, d_cost, d_low);
    run_chunk(h, d_dets + (size_t)half * stride, stride, shots - half, ts, d_obs ? d_obs + (size_t)half * ob : nullptr,
              d_cost ? d_cost + half : nullptr, d_low ? d_low + half : nullptr);
    return;
  }
  h.sol_stride = stride_need;

  // Longest searches first: shots in decreasing order of fired detectors (counting sort on the device).
  const uint32_t* d_shot_order = nullptr;
  if (shots > 1) {
    tsr::OrderParams op;
    op.fired = h.d_fired.as<uint32_t>();
    op.shots = shots;
    op.bins = hs.max_fired + 1;
    h.d_bins.ensure((size_t)op.bins * 4);
    h.d_order.ensure((size_t)shots * 4);
    CUDA_OK(cudaMemsetAsync(h.d_bins.p, 0, (size_t)op.bins * 4, s));
    op.bin_cursor = h.d_bins.as<uint32_t>();
    op.order = h.d_order.as<uint32_t>();
    CUDA_OK((cudaError_t)tsr::launch_order(op, s));
    d_shot_order = op.order;
  }

  h.d_status.ensure(items);
  h.d_sol_len.ensure((size_t)items * 4);
  h.d_fcost.ensure((size_t)items * 8);
  h.d_sol.ensure((size_t)items * h.sol_stride * 4);
  h.d_err_off.ensure((size_t)shots * 8);
  h.d_err_len.ensure((size_t)shots * 4);
  const uint64_t err_cap = (uint64_t)shots * h.sol_stride;
  if (h.collect_errors) h.d_err.ensure(err_cap * 8);
  h.d_retry.ensure((size_t)shots * 4 * 2);

  auto search_params = [&](const Tier& tier, const uint32_t* item_ids, uint32_t n_work) {
    tsr::SearchParams p;
    p.t = h.dt;
    p.dets = d_dets;
    p.stride = stride;
    p.n_work = n_work;
    p.item_ids = item_ids;
    p.shot_order = item_ids ? nullptr : d_shot_order;
    p.n_trials = U;
    p.trial_perm = ts.d_perm.as<uint32_t>();
    p.trial_beam = ts.d_beam.as<uint32_t>();
    p.det_penalty = h.cfg.det_penalty;
    p.pqlimit = h.cfg.pqlimit;
    p.no_revisit = h.cfg.no_revisit_dets != 0;
    p.at_most_two = h.cfg.at_most_two_errors_per_detector != 0;
    p.n_slots = tier.slots;
    p.node_cap = tier.node_cap;
    p.visit_cap = tier.visit_cap;
    p.vlist_cap = h.vlist_cap;
    char* base = h.pool.as<char>();
    p.heap = reinterpret_cast<tsr::HeapNode*>(base);
    p.chain = reinterpret_cast<tsr::ChainNode*>(base + (size_t)tier.slots * tier.node_cap * 16);
    p.visited = reinterpret_cast<uint4*>(base + (size_t)tier.slots * tier.node_cap * 32);
    p.vlist = h.vlist.as<uint32_t>();
    p.sp_act = h.sp_act.as<uint32_t>();
    p.hcache = h.hcache.as<double>();
    p.hc_shared = h.hc_shared ? 1 : 0;
    p.sol_stride = h.sol_stride;
    p.sol = h.d_sol.as<uint32_t>();
    p.sol_len = h.d_sol_len.as<uint32_t>();
    p.status = h.d_status.as<uint8_t>();
    p.fcost = h.d_fcost.as<double>();
    p.work_counter = &sc->work;
    p.counters = h.d_counters.as<unsigned long long>();
    p.prof = h.profile ? h.d_prof.as<unsigned long long>() : nullptr;
    if (h.visited_audit && h.fast && tier.visit_cap) {
      p.vchain = h.d_vchain.as<uint32_t>();
      p.audit = h.d_audit.as<unsigned long long>();
      p.audit_scratch = h.d_audit_scratch.as<uint32_t>();
    }
    if (h.trace_on) {
      p.trace = h.d_trace.as<uint32_t>();
      p.trace_cap = h.trace_cap;
      p.trace_len = h.d_trace_len.as<uint32_t>();
      p.trace_arena = h.d_trace_arena.as<unsigned long long>();
    }
    return p;
  };
  auto resolve_params = [&](const uint32_t* shot_ids, uint32_t n, uint32_t* retry_out) {
    tsr::ResolveParams r;
    r.t = h.dt;
    r.n_shots = n;
    r.shot_ids = shot_ids;
    r.n_trials = U;
    r.n_literal = (uint32_t)ts.literal_to_trial.size();
    r.literal_to_trial = ts.d_literal.as<uint32_t>();
    r.sol_stride = h.sol_stride;
    r.sol = h.d_sol.as<uint32_t>();
    r.sol_len = h.d_sol_len.as<uint32_t>();
    r.status = h.d_status.as<uint8_t>();
    r.obs_out = d_obs;
    r.obs_stride = (uint32_t)((T.num_observables + 7) / 8);
    r.cost_out = d_cost;
    r.low_conf_out = d_low;
    r.err_out = h.d_err.as<uint64_t>();
    r.err_cap = err_cap;
    r.err_cursor = &sc->err_cursor;
    r.err_off = h.collect_errors ? h.d_err_off.as<uint64_t>() : nullptr;
    r.err_len = h.d_err_len.as<uint32_t>();
    r.retry_shots = retry_out;
    r.retry_count = &sc->retry_count;
    return r;
  };

  // Tier 0 over every item.
  const Tier& t0 = h.tiers[0];
  // Fast path: tier 0 runs with an expansion budget and is followed by a launch of wide CTAs that finish whatever was
  // paused (no host round trip: CTAs beyond the paused count exit at once); the retry tiers run wide from the start.
  const bool widen = h.fast && h.wpb_wide && h.expansion_budget && !h.instrumented;
  // counter S, the phase profile, the visited audit and the visualization trace exist only in the diagnostics variants
  const bool diag = h.instrumented || h.profile || h.visited_audit || h.trace_on;
  bool resumed = false;
  auto launch = [&](tsr::SearchParams sp_, uint32_t tier_slots, uint32_t n_work, bool tier0) -> int {
    if (!h.fast)
      return tsr::launch_search(sp_, std::min<uint32_t>((tier_slots + h.wpb - 1) / h.wpb, (n_work + h.wpb - 1) / h.wpb), h.wpb, s);
    const bool bits = h.bits && !h.instrumented;
    if (!widen) return tsr::launch_search_fast(sp_, std::min<uint32_t>(tier_slots, n_work), h.wpb, diag, bits, false, s);
    // Few long searches left: give each a thread-block cluster (up to 8 SMs share the children of one search).
    auto cluster_for = [&](uint32_t searches) {
      uint32_t c = h.cluster_max;
      while (c > 1 && (uint64_t)searches * c > h.blocks_wide) c /= 2;
      return std::max<uint32_t>(c, 1);
    };
    if (!tier0) {
      const uint32_t units = std::min<uint32_t>(tier_slots, n_work);
      sp_.cluster = cluster_for(units);
      return tsr::launch_search_fast(sp_, units * sp_.cluster, h.wpb_wide, diag, bits, true, s);
    }
    sp_.expansion_budget = h.expansion_budget;
    sp_.paused = h.d_paused.as<tsr::ResumeRec>();
    sp_.paused_count = &sc->paused_count;
    sp_.paused_cap = tier_slots;
    if (int e = tsr::launch_search_fast(sp_, std::min<uint32_t>(tier_slots, n_work), h.wpb, diag, bits, false, s)) return e;
    // The tail, in stages (a host round trip each, but only at the very end of the chunk's search work): wide CTAs
    // continue the paused searches for another slice of expansions and park the survivors again; once few enough are
    // left for a full-size cluster each, the last stage runs them to the end.
    unsigned int n_paused = 0;
    if (cudaError_t e = cudaMemcpyAsync(&n_paused, &sc->paused_count, 4, cudaMemcpyDeviceToHost, s)) return (int)e;
    if (cudaError_t e = cudaStreamSynchronize(s)) return (int)e;
    if (n_paused == 0) return 0;
    n_paused = std::min<uint32_t>(n_paused, tier_slots);
    h.tail[0] += n_paused;
    resumed = true;
    if (cudaError_t e = cudaEventRecord(h.ev[4], s)) return (int)e;
    const tsr::ResumeRec* in = h.d_paused.as<tsr::ResumeRec>();
    unsigned int* counts = h.d_stage_counts.as<unsigned int>();
    for (int stage = 0; n_paused > 0; ++stage) {
      const uint32_t units = std::min<uint32_t>(n_paused, h.blocks_wide);
      tsr::SearchParams rp = sp_;
      rp.resume = 1;
      rp.resume_list = in;
      rp.resume_n = n_paused;
      rp.cluster = cluster_for(units);
      const bool last = rp.cluster >= h.cluster_max || stage >= 8 || h.stage_budget == 0;
      tsr::ResumeRec* out = in == h.d_paused.as<tsr::ResumeRec>() ? h.d_paused2.as<tsr::ResumeRec>() : h.d_paused.as<tsr::ResumeRec>();
      rp.expansion_budget = last ? 0 : h.stage_budget;
      rp.paused = out;
      rp.paused_count = counts + (stage & 1);
      rp.paused_cap = tier_slots;
      if (cudaError_t e = cudaMemsetAsync(rp.paused_count, 0, 4, s)) return (int)e;
      h.tail[2] = std::max<double>(h.tail[2], rp.cluster);
      if (int e = tsr::launch_search_fast(rp, units * rp.cluster, h.wpb_wide, diag, bits, true, s)) return e;
      if (cudaError_t e = cudaEventRecord(h.ev[5], s)) return (int)e;
      if (last) break;
      if (cudaError_t e = cudaMemcpyAsync(&n_paused, rp.paused_count, 4, cudaMemcpyDeviceToHost, s)) return (int)e;
      if (cudaError_t e = cudaStreamSynchronize(s)) return (int)e;
      n_paused = std::min<uint32_t>(n_paused, tier_slots);
      in = out;
    }
    return 0;
  };
  // An item no CTA got to must never pass for decoded: every status starts as "overflow" (-> retried in the next tier).
  CUDA_OK(cudaMemsetAsync(h.d_status.p, tsr::kItemNodeOverflow, items, s));
  CUDA_OK(cudaEventRecord(h.ev[1], s));
  CUDA_OK((cudaError_t)launch(search_params(t0, nullptr, items), t0.slots, items, true));
  CUDA_OK(cudaEventRecord(h.ev[2], s));
  uint32_t* retry_a = h.d_retry.as<uint32_t>();
  uint32_t* retry_b = retry_a + shots;
  CUDA_OK((cudaError_t)tsr::launch_resolve(resolve_params(nullptr, shots, retry_a), s));
  CUDA_OK(cudaMemcpyAsync(&hs, sc, sizeof(Scalars), cudaMemcpyDeviceToHost, s));
  CUDA_OK(cudaStreamSynchronize(s));
  double launches = 1, rerun = 0;
  float ms_search = 0;
  CUDA_OK(cudaEventElapsedTime(&ms_search, h.ev[1], h.ev[2]));
  if (resumed) {
    float ms_tail = 0;
    CUDA_OK(cudaEventElapsedTime(&ms_tail, h.ev[4], h.ev[5]));
    h.tail[1] += ms_tail;
    launches += 1;
  }

  // Retry tiers for searches that outgrew their slot.
  // Retried shots write their solutions to a separate buffer with longer rows (a search can also be sent here because
  // its solution did not fit the tier-0 row, which is sized from the batch's largest syndrome).
  const uint32_t retry_stride = (uint32_t)std::min<uint64_t>(std::min<uint64_t>(std::max<uint64_t>(T.num_errors, 1), 65535),
                                                              std::max<uint64_t>(8ull * h.sol_stride, 2048));
  for (size_t tier = 1; hs.retry_count > 0; ++tier) {
    if (tier >= h.tiers.size())
      throw DeviceMemoryError("a search outgrew the device node pool, or its solution the " + std::to_string(retry_stride) +
                              "-error solution buffer (" + std::to_string(hs.retry_count) +
                              " shot(s)); raise node_pool_bytes, lower pqlimit or set a finite beam");
    const uint32_t n_retry = hs.retry_count;
    std::vector<uint32_t> shot_ids(n_retry), item_ids((size_t)n_retry * U);
    CUDA_OK(cudaMemcpy(shot_ids.data(), retry_a, (size_t)n_retry * 4, cudaMemcpyDeviceToHost));
    std::sort(shot_ids.begin(), shot_ids.end());
    for (uint32_t i = 0; i < n_retry; ++i)
      for (uint32_t u = 0; u < U; ++u) item_ids[(size_t)i * U + u] = shot_ids[i] * U + u;
    h.d_items.ensure(item_ids.size() * 4 + (size_t)n_retry * 4);
    uint32_t* d_item_ids = h.d_items.as<uint32_t>();
    uint32_t* d_shot_ids = d_item_ids + item_ids.size();
    CUDA_OK(cudaMemcpyAsync(d_item_ids, item_ids.data(), item_ids.size() * 4, cudaMemcpyHostToDevice, s));
    CUDA_OK(cudaMemcpyAsync(d_shot_ids, shot_ids.data(), (size_t)n_retry * 4, cudaMemcpyHostToDevice, s));
    CUDA_OK(cudaMemsetAsync(&sc->work, 0, 4, s));
    CUDA_OK(cudaMemsetAsync(&sc->retry_count, 0, 4, s));
    // A differently-shaped tier reuses the pool: visited tables must start empty.
    const Tier& tk = h.tiers[tier];
    if (tk.visit_cap)
      CUDA_OK(cudaMemsetAsync(h.pool.as<char>() + (size_t)tk.slots * tk.node_cap * 32, 0, (size_t)tk.slots * tk.visit_cap * 16, s));
    CUDA_OK(cudaEventRecord(h.ev[1], s));
    h.d_sol_retry.ensure(item_ids.size() * (size_t)retry_stride * 4);
    tsr::SearchParams rp = search_params(tk, d_item_ids, (uint32_t)item_ids.size());
    rp.sol = h.d_sol_retry.as<uint32_t>();
    rp.sol_stride = retry_stride;
    rp.sol_by_work = 1;
    CUDA_OK((cudaError_t)launch(rp, tk.slots, (uint32_t)item_ids.size(), false));
    CUDA_OK(cudaEventRecord(h.ev[2], s));
    tsr::ResolveParams rr = resolve_params(d_shot_ids, n_retry, retry_b);
    rr.sol = h.d_sol_retry.as<uint32_t>();
    rr.sol_stride = retry_stride;
    rr.sol_by_rank = 1;
    CUDA_OK((cudaError_t)tsr::launch_resolve(rr, s));
    // Restore the tier-0 invariant (empty visited tables in the tier-0 layout).
    if (t0.visit_cap)
      CUDA_OK(cudaMemsetAsync(h.pool.as<char>() + (size_t)t0.slots * t0.node_cap * 32, 0, (size_t)t0.slots * t0.visit_cap * 16, s));
    CUDA_OK(cudaMemcpyAsync(&hs, sc, sizeof(Scalars), cudaMemcpyDeviceToHost, s));
    CUDA_OK(cudaStreamSynchronize(s));
    float ms = 0;
    CUDA_OK(cudaEventElapsedTime(&ms, h.ev[1], h.ev[2]));
    ms_search += ms;
    launches += 1;
    rerun += (double)item_ids.size();
    std::swap(retry_a, retry_b);
  }
  if (hs.err_cursor > err_cap) throw DeviceMemoryError("pooled error lists overflowed their buffer");

  // Error lists -> host CSR.
  if (h.collect_errors) {
    std::vector<uint64_t> off(shots), pooled(hs.err_cursor);
    std::vector<uint32_t> len(shots);
    CUDA_OK(cudaMemcpyAsync(off.data(), h.d_err_off.p, (size_t)shots * 8, cudaMemcpyDeviceToHost, s));
    CUDA_OK(cudaMemcpyAsync(len.data(), h.d_err_len.p, (size_t)shots * 4, cudaMemcpyDeviceToHost, s));
    if (!pooled.empty())
      CUDA_OK(cudaMemcpyAsync(pooled.data(), h.d_err.p, pooled.size() * 8, cudaMemcpyDeviceToHost, s));
    CUDA_OK(cudaEventRecord(h.ev[3], s));
    CUDA_OK(cudaStreamSynchronize(s));
    for (uint32_t k = 0; k < shots; ++k) {
      h.last_idx.insert(h.last_idx.end(), pooled.begin() + (std::ptrdiff_t)off[k], pooled.begin() + (std::ptrdiff_t)(off[k] + len[k]));
      h.last_off.push_back(h.last_idx.size());
    }
  } else {
    CUDA_OK(cudaEventRecord(h.ev[3], s));
    CUDA_OK(cudaStreamSynchronize(s));
    h.last_off.resize(h.last_off.size() + shots, 0);
  }
  float ms_all = 0;
  CUDA_OK(cudaEventElapsedTime(&ms_all, h.ev[0], h.ev[3]));
  h.timing[0] += ms_search;
  h.timing[1] += ms_all;
  h.timing[2] += launches;
  h.timing[3] += rerun;
}

void begin_call(tsr_decoder& h) {
  h.last_off.assign(1, 0);
  h.last_idx.clear();
  for (double& v : h.timing) v = 0;
  for (double& v : h.tail) v = 0;
}

}  // namespace
void tsr_impl::decode_batch(tsr_decoder& h, const uint8_t* dets, uint64_t stride, uint64_t shots, bool on_device,
                            const TrialSet& ts, uint8_t* obs, double* cost, uint8_t* low) {
  require_device(h);
  CUDA_OK(cudaSetDevice(h.device));
  const uint64_t min_stride = (h.T.num_detectors + 7) / 8;
  if (shots && stride < min_stride)
    throw std::invalid_argument("stride (" + std::to_string(stride) + ") is smaller than ceil(num_detectors/8) (" +
                                std::to_string(min_stride) + ")");
  begin_call(h);
  const uint64_t obs_stride = (h.T.num_observables + 7) / 8;
  for (uint64_t at = 0; at < shots; at += h.max_batch) {
    const uint32_t n = (uint32_t)std::min<uint64_t>(h.max_batch, shots - at);
    if (on_device) {
      run_chunk(h, dets + at * stride, stride, n, ts, obs ? obs + at * obs_stride : nullptr, cost ? cost + at : nullptr,
                low ? low + at : nullptr);
    } else {
      h.d_dets.ensure((size_t)n * stride);
      h.d_obs.ensure(std::max<size_t>((size_t)n * obs_stride, 16));
      h.d_cost.ensure((size_t)n * 8);
      h.d_low.ensure(n);
      CUDA_OK(cudaMemcpyAsync(h.d_dets.p, dets + at * stride, (size_t)n * stride, cudaMemcpyHostToDevice, h.stream));
      run_chunk(h, h.d_dets.as<uint8_t>(), stride, n, ts, h.d_obs.as<uint8_t>(), h.d_cost.as<double>(), h.d_low.as<uint8_t>());
      if (obs && obs_stride)
        CUDA_OK(cudaMemcpyAsync(obs + at * obs_stride, h.d_obs.p, (size_t)n * obs_stride, cudaMemcpyDeviceToHost, h.stream));
      if (cost) CUDA_OK(cudaMemcpyAsync(cost + at, h.d_cost.p, (size_t)n * 8, cudaMemcpyDeviceToHost, h.stream));
      if (low) CUDA_OK(cudaMemcpyAsync(low + at, h.d_low.p, n, cudaMemcpyDeviceToHost, h.stream));
      CUDA_OK(cudaStreamSynchronize(h.stream));
    }
  }
}

namespace {
uint64_t mapped_error(const tsr_decoder& h, uint64_t dem_error) {
  // dem_error_to_error.at(i) and the "removed" check (tesseract.cc:620-624).
  if (dem_error >= h.T.dem_error_to_error.size()) throw std::out_of_range("vector::_M_range_check: error index out of range");
  uint64_t e = h.T.dem_error_to_error[dem_error];
  if (e == tsr::kNoError) throw std::invalid_argument("error index does not map to a retained decoder error");
  return e;
}

// Turns the device trace of a single-shot decode into Visualizer lines (visualization.cc:26-51), trial by trial in
// the reference's literal order (identical (order, beam) trials are searched once and logged once per literal trial).
void append_visualization(tsr_decoder& h, const TrialSet& ts, const std::vector<uint8_t>& packed) {
  const size_t U = ts.perm.size();
  std::vector<uint32_t> len(U);
  std::vector<unsigned long long> arena(U);
  CUDA_OK(cudaStreamSynchronize(h.stream));
  CUDA_OK(cudaMemcpy(len.data(), h.d_trace_len.p, U * 4, cudaMemcpyDeviceToHost));
  CUDA_OK(cudaMemcpy(arena.data(), h.d_trace_arena.p, U * 8, cudaMemcpyDeviceToHost));
  std::vector<std::vector<std::string>> per_trial(U);
  const uint64_t D = h.T.num_detectors;
  for (size_t u = 0; u < U; ++u) {
    if (len[u] > h.trace_cap) throw std::runtime_error("the visualization log of this decode exceeds " + std::to_string(h.trace_cap) + " nodes");
    std::vector<uint32_t> trace(len[u]);
    if (len[u]) CUDA_OK(cudaMemcpy(trace.data(), h.d_trace.as<uint32_t>() + u * h.trace_cap, (size_t)len[u] * 4, cudaMemcpyDeviceToHost));
    uint32_t top = 0;
    bool any = false;
    for (uint32_t v : trace)
      if (v != tsr::kNoChain) {
        top = std::max(top, v);
        any = true;
      }
    std::vector<tsr::ChainNode> nodes(any ? (size_t)top + 1 : 0);  // a node's ancestors have smaller indices
    if (any) CUDA_OK(cudaMemcpy(nodes.data(), (const void*)(uintptr_t)arena[u], nodes.size() * sizeof(tsr::ChainNode), cudaMemcpyDeviceToHost));
    for (uint32_t v : trace) {
      std::string errs = "activated_errors = ";
      std::vector<uint8_t> fired(packed);
      for (uint32_t at = v; at != tsr::kNoChain; at = nodes[at].parent) {
        errs += std::to_string(nodes[at].error) + ", ";
        for (int d : h.T.errors[nodes[at].error].detectors) fired[(size_t)d >> 3] ^= (uint8_t)(1u << (d & 7));
      }
      std::string dets = "activated_detectors = ";
      for (uint64_t d = 0; d < D; ++d)
        if ((fired[d >> 3] >> (d & 7)) & 1) dets += std::to_string(d) + ", ";
      per_trial[u].push_back(std::move(errs));
      per_trial[u].push_back(std::move(dets));
    }
  }
  for (uint32_t u : ts.literal_to_trial) h.viz_lines.insert(h.viz_lines.end(), per_trial[u].begin(), per_trial[u].end());
}

char* dup_text(const std::string& s) {
  char* out = (char*)std::malloc(s.size() + 1);
  if (!out) throw std::bad_alloc();
  std::memcpy(out, s.c_str(), s.size() + 1);
  return out;
}

}  // namespace

extern "C" {

void tsr_config_init(tsr_config* cfg) {
  std::memset(cfg, 0, sizeof(*cfg));
  cfg->det_beam = 5;
  cfg->no_revisit_dets = 1;
  cfg->merge_errors = 1;
  cfg->pqlimit = 200000;
  cfg->device = -1;
}

const char* tsr_last_error(void) {
  return last_error().c_str();
}
const char* tsr_version(void) {
  return "tesseract_b200 0.1.0 (sm_100a)";
}

int tsr_create(const char* dem_text, const tsr_config* cfg, uint32_t flags, tsr_decoder** out) {
  *out = nullptr;
  return guarded([&] {
    auto h = std::make_unique<tsr_decoder>();
    h->cfg = *cfg;
    h->cfg.det_orders = nullptr;
    DemModel dem = DemModel::parse(dem_text);
    h->T = tsr::ingest(dem, cfg->merge_errors != 0);
    const uint64_t D = h->T.num_detectors;
    // tesseract.cc:175-188.
    if (cfg->num_det_orders == 0) {
      h->det_orders.emplace_back(D);
      for (uint64_t i = 0; i < D; ++i) h->det_orders[0][i] = i;
    } else {
      if (!cfg->det_orders) throw std::invalid_argument("det_orders is null but num_det_orders > 0");
      if (cfg->det_order_len && cfg->det_order_len != D)  // tesseract.cc:178-184
        throw std::invalid_argument("Each detector order list must have a size equal to the number of detectors.");
      for (uint64_t k = 0; k < cfg->num_det_orders; ++k)
        h->det_orders.emplace_back(cfg->det_orders + k * D, cfg->det_orders + (k + 1) * D);
    }
    if (cfg->det_beam < 0) throw std::invalid_argument("det_beam must be >= 0");
    if (cfg->sparsify_errors) {  // tesseract.cc:256-277
      if (cfg->sparsify_base_degree <= 0)
        throw std::invalid_argument("sparsify_base_degree must be > 0 when sparsify_errors is enabled.");
      if (cfg->sparsify_max_degree < -1) throw std::invalid_argument("sparsify_max_degree must be >= -1.");
      if (cfg->sparsify_reactivate_limit < -1) throw std::invalid_argument("sparsify_reactivate_limit must be >= -1.");
      if (cfg->sparsify_max_degree >= 0 && cfg->sparsify_max_degree < cfg->sparsify_base_degree)
        throw std::invalid_argument("sparsify_max_degree must be >= sparsify_base_degree.");
      h->sparsify_limit = cfg->sparsify_reactivate_limit;
      if (h->sparsify_limit == -1)
        h->sparsify_limit = suggest_limit_capped(D, cfg->sparsify_base_degree,
                                                 (int)std::min<uint64_t>(h->T.num_errors, (uint64_t)std::numeric_limits<int>::max()));
      if (h->T.num_errors >= (uint64_t{1} << 28))
        throw std::invalid_argument("tesseract_b200: sparsify_errors supports fewer than 2^28 errors");
    }
    if (h->T.max_degree > 32)
      throw std::invalid_argument("tesseract_b200: errors touching more than 32 detectors are not supported");
    std::map<std::vector<uint64_t>, uint32_t> perm_ids;
    for (const auto& o : h->det_orders) {
      auto it = perm_ids.emplace(o, (uint32_t)perm_ids.size()).first;
      h->order_to_perm.push_back(it->second);
    }
    h->n_perms = (uint32_t)perm_ids.size();
    h->F = tsr::build_fast_tables(h->T);
    h->fast = h->F.certified && !cfg->force_generic_kernel;
    if (h->fast && cfg->row_scan != 1) {
      // Bit-sliced scans pay a fixed cost per scanned row and win when rows are long and errors touch many
      // detectors (measured: BB [[144,12,12]] 1.7x, colour / surface codes 0.65-0.8x); sum(deg^2)/D is the
      // entry-per-lane work per row scan.
      double work = 0;
      for (const auto& e : h->T.errors) work += (double)e.detectors.size() * (double)e.detectors.size();
      work /= (double)std::max<uint64_t>(D, 1);
      if (cfg->row_scan == 2 || work > 1200.0) {
        h->B = tsr::build_bit_tables(h->T, h->F);
        h->bits = h->B.ok;
      }
    }
    if (h->F.certified && cfg->force_generic_kernel) h->F.why_not = "force_generic_kernel is set";
    if (cfg->create_visualization) {  // tesseract.cc:200-204, visualization.cc:5-24
      DemModel compiled = h->T.compiled_dem;
      for (const auto& coords : tsr::get_detector_coords(compiled)) {
        std::ostringstream ss;
        ss << "Detector D" << h->viz_lines.size() << " coordinate (";
        const size_t e = std::min<size_t>(3, coords.size());
        for (size_t i = 0; i < e; ++i) ss << coords[i] << (i + 1 < e ? ", " : "");
        ss << ")";
        h->viz_lines.push_back(ss.str());
      }
      for (const auto& err : h->T.errors) {  // common.cc:25-50, 90-94
        std::ostringstream ss;
        ss << std::fixed << std::setprecision(6) << err.cost;
        auto list = [](const std::vector<int>& v) {
          std::string t = "[";
          for (size_t i = 0; i < v.size(); ++i) t += std::to_string(v[i]) + (i + 1 < v.size() ? " " : "");
          return t + "]";
        };
        h->viz_lines.push_back("Error{cost=" + ss.str() + ", symptom=Symptom{detectors=" + list(err.detectors) + ", observables=" +
                               list(err.observables) + "}}");
      }
    }
    h->host_only = (flags & TSR_CREATE_HOST_ONLY) != 0;
    if (!h->host_only) init_device(*h);
    *out = h.release();
  });
}

void tsr_destroy(tsr_decoder* h) {
  if (!h) return;
  if (!h->host_only) cudaSetDevice(h->device);
  delete h;
}

uint64_t tsr_num_detectors(const tsr_decoder* h) { return h->T.num_detectors; }
uint64_t tsr_num_observables(const tsr_decoder* h) { return h->T.num_observables; }
uint64_t tsr_num_errors(const tsr_decoder* h) { return h->T.num_errors; }
uint64_t tsr_num_dem_errors(const tsr_decoder* h) { return h->T.num_dem_errors; }
uint64_t tsr_num_det_orders(const tsr_decoder* h) { return h->det_orders.size(); }
int64_t tsr_sparsify_reactivate_limit(const tsr_decoder* h) { return h->sparsify_limit; }
int64_t tsr_suggest_sparsify_reactivate_limit(uint64_t num_detectors, int32_t sparsify_base_degree) {
  int64_t out = TSR_E_INVALID_ARGUMENT;
  guarded([&] { out = suggest_limit_capped(num_detectors, sparsify_base_degree, std::numeric_limits<int>::max()); });
  return out;
}

int tsr_error_costs(const tsr_decoder* h, double* out) {
  std::copy(h->T.e_cost.begin(), h->T.e_cost.end(), out);
  return TSR_OK;
}

int tsr_maps(const tsr_decoder* h, uint64_t* dem_error_to_error, uint64_t* error_to_dem_error) {
  if (dem_error_to_error) std::copy(h->T.dem_error_to_error.begin(), h->T.dem_error_to_error.end(), dem_error_to_error);
  if (error_to_dem_error) std::copy(h->T.error_to_dem_error.begin(), h->T.error_to_dem_error.end(), error_to_dem_error);
  return TSR_OK;
}

int64_t tsr_csr(const tsr_decoder* h, int which, uint64_t* offsets, int32_t* idx) {
  const tsr::DemTables& T = h->T;
  auto emit = [&](const std::vector<uint32_t>& off, const std::vector<int32_t>& v) -> int64_t {
    if (offsets) std::copy(off.begin(), off.end(), offsets);
    if (idx) std::copy(v.begin(), v.end(), idx);
    return (int64_t)v.size();
  };
  switch (which) {
    case 0:
      return emit(T.row_off, T.row_err);
    case 1:
      return emit(T.e_off, T.e_det);
    case 2: {
      auto nb = T.eneighbors();
      uint64_t n = 0;
      for (size_t e = 0; e < nb.size(); ++e) {
        if (offsets) offsets[e] = n;
        for (int v : nb[e]) {
          if (idx) idx[n] = v;
          ++n;
        }
      }
      if (offsets) offsets[nb.size()] = n;
      return (int64_t)n;
    }
    case 3:
      return emit(T.eo_off, T.eo_obs);
    default:
      last_error() = "unknown table id";
      return TSR_E_INVALID_ARGUMENT;
  }
}

int tsr_compiled_dem_text(const tsr_decoder* h, char** out) {
  return guarded([&] { *out = dup_text(h->T.compiled_dem.str()); });
}

void tsr_free(void* p) {
  std::free(p);
}

int tsr_decode_batch_b8(tsr_decoder* h, const uint8_t* dets, uint64_t stride, uint64_t shots, uint8_t* obs_out,
                        double* cost_out, uint8_t* low_conf_out) {
  return guarded([&] { decode_batch(*h, dets, stride, shots, false, h->ensemble, obs_out, cost_out, low_conf_out); });
}

int tsr_decode_batch_b8_device(tsr_decoder* h, const uint8_t* d_dets, uint64_t stride, uint64_t shots,
                               uint8_t* d_obs_out, double* d_cost_out, uint8_t* d_low_conf_out) {
  return guarded([&] { decode_batch(*h, d_dets, stride, shots, true, h->ensemble, d_obs_out, d_cost_out, d_low_conf_out); });
}

int tsr_set_collect_errors(tsr_decoder* h, int on) {
  h->collect_errors = on != 0;
  return TSR_OK;
}

int64_t tsr_cross_check_detcost(const tsr_decoder* h, uint64_t trials, uint64_t seed, uint64_t* checked) {
  int64_t bad = 0;
  const int rc = guarded([&] {
    tsr::BitTables local;
    const tsr::BitTables* B = &h->B;
    if (h->F.certified && !h->B.ok) {  // the handle chose the entry-per-lane scan: build the tables for the check
      local = tsr::build_bit_tables(h->T, h->F);
      B = &local;
    }
    bad = (int64_t)tsr::cross_check_detcost(h->T, h->F, *B, trials, seed, checked);
  });
  return rc == TSR_OK ? bad : rc;
}

int tsr_set_instrumented(tsr_decoder* h, int on) {
  h->instrumented = on != 0;
  return TSR_OK;
}

int tsr_search_path(const tsr_decoder* h) {
  return h->fast ? (h->bits ? 2 : 1) : 0;
}

const char* tsr_search_path_reason(const tsr_decoder* h) {
  return h->fast ? "" : h->F.why_not.c_str();
}

uint64_t tsr_last_batch_nnz(const tsr_decoder* h) {
  return h->last_idx.size();
}

int tsr_last_batch_errors(const tsr_decoder* h, uint64_t* offsets, uint64_t* idx) {
  if (offsets) std::copy(h->last_off.begin(), h->last_off.end(), offsets);
  if (idx) std::copy(h->last_idx.begin(), h->last_idx.end(), idx);
  return TSR_OK;
}

int tsr_decode(tsr_decoder* h, const uint64_t* detections, uint64_t n, int64_t order, uint64_t beam, uint64_t* errs_out,
               uint64_t errs_cap, uint64_t* n_errs, double* cost, int32_t* low_confidence) {
  return guarded([&] {
    require_device(*h);
    const uint64_t D = h->T.num_detectors;
    const uint64_t stride = std::max<uint64_t>((D + 7) / 8, 1);
    std::vector<uint8_t> packed(stride, 0);
    for (uint64_t i = 0; i < n; ++i) {
      const uint64_t d = detections[i];
      if (d >= D)  // tesseract.cc:400-404
        throw std::runtime_error("Symptom " + std::to_string(d) + " references a detector >= num_detectors (= " +
                                 std::to_string(D) + ").");
      packed[d >> 3] |= (uint8_t)(1u << (d & 7));
    }
    double c = 0;
    uint8_t low = 0;
    const TrialSet* used = &h->ensemble;
    TrialSet one;
    if (h->cfg.create_visualization) {  // per distinct trial: the chain index of every node the reference logs
      CUDA_OK(cudaSetDevice(h->device));
      const size_t U = std::max<size_t>(h->ensemble.perm.size(), 1);
      h->trace_cap = (uint32_t)std::min<uint64_t>(h->cfg.pqlimit == TSR_PQLIMIT_UNBOUNDED ? (1u << 20) : h->cfg.pqlimit + 2, 1u << 20);
      h->d_trace.ensure(U * h->trace_cap * 4);
      h->d_trace_len.ensure(U * 4);
      h->d_trace_arena.ensure(U * 8);
      CUDA_OK(cudaMemsetAsync(h->d_trace_len.p, 0, U * 4, h->stream));
      h->trace_on = true;
    }
    struct TraceOff {
      tsr_decoder* h;
      ~TraceOff() { h->trace_on = false; }
    } trace_off{h};
    if (order < 0) {
      decode_batch(*h, packed.data(), stride, 1, false, h->ensemble, nullptr, &c, &low);
    } else {
      if ((uint64_t)order >= h->det_orders.size()) throw std::invalid_argument("detector_order is out of range");
      if (beam > TSR_INF_DET_BEAM) beam = TSR_INF_DET_BEAM;
      one.perm = {h->order_to_perm[(size_t)order]};
      one.beam = {(uint32_t)beam};
      one.literal_to_trial = {0};
      CUDA_OK(cudaSetDevice(h->device));
      one.upload();
      decode_batch(*h, packed.data(), stride, 1, false, one, nullptr, &c, &low);
      used = &one;
    }
    if (h->trace_on) append_visualization(*h, *used, packed);
    if (cost) *cost = c;
    if (low_confidence) *low_confidence = low;
    if (n_errs) *n_errs = h->last_idx.size();
    if (errs_out) {
      if (h->last_idx.size() > errs_cap) throw std::invalid_argument("errs_out is too small for the predicted errors");
      std::copy(h->last_idx.begin(), h->last_idx.end(), errs_out);
    }
  });
}

int tsr_visualization_text(const tsr_decoder* h, char** out) {
  return guarded([&] {
    std::string text;
    for (const std::string& line : h->viz_lines) text += line + "\n";
    *out = dup_text(text);
  });
}

int tsr_cost_from_errors(const tsr_decoder* h, const uint64_t* errs, uint64_t n, double* out) {
  return guarded([&] {
    double total = 0;
    for (uint64_t i = 0; i < n; ++i) total += h->T.e_cost[mapped_error(*h, errs[i])];
    *out = total;
  });
}

int tsr_flipped_observables(const tsr_decoder* h, const uint64_t* errs, uint64_t n, uint8_t* obs_bits_out) {
  return guarded([&] {
    const uint64_t nb = (h->T.num_observables + 7) / 8;
    std::vector<uint8_t> bits(nb, 0);
    for (uint64_t i = 0; i < n; ++i)
      for (int ob : h->T.errors[mapped_error(*h, errs[i])].observables) bits[(size_t)ob >> 3] ^= (uint8_t)(1u << (ob & 7));
    std::copy(bits.begin(), bits.end(), obs_bits_out);
  });
}

int tsr_counters(tsr_decoder* h, uint64_t out[8], int reset) {
  return guarded([&] {
    require_device(*h);
    CUDA_OK(cudaSetDevice(h->device));
    CUDA_OK(cudaStreamSynchronize(h->stream));
    CUDA_OK(cudaMemcpyAsync(out, h->d_counters.p, 64, cudaMemcpyDeviceToHost, h->stream));
    if (reset) CUDA_OK(cudaMemsetAsync(h->d_counters.p, 0, 64, h->stream));
    CUDA_OK(cudaStreamSynchronize(h->stream));
  });
}

int tsr_set_visited_audit(tsr_decoder* h, int on) {
  return guarded([&] {
    require_device(*h);
    CUDA_OK(cudaSetDevice(h->device));
    h->visited_audit = on != 0;
    if (!on) return;
    size_t entries = 0;  // the largest slots x visit_cap of any tier
    for (const Tier& t : h->tiers) entries = std::max<size_t>(entries, (size_t)t.slots * t.visit_cap);
    h->d_vchain.ensure(std::max<size_t>(entries * 4, 16));
    h->d_audit.ensure(16);
    h->d_audit_scratch.ensure(std::max<size_t>((size_t)h->tiers[0].slots * 8 * 32 * h->dt.nwords * 4, 16));
    CUDA_OK(cudaMemsetAsync(h->d_audit.p, 0, 16, h->stream));
    CUDA_OK(cudaStreamSynchronize(h->stream));
  });
}

int tsr_visited_audit(tsr_decoder* h, uint64_t out[2], int reset) {
  return guarded([&] {
    require_device(*h);
    out[0] = out[1] = 0;
    if (!h->d_audit.p) return;
    CUDA_OK(cudaSetDevice(h->device));
    CUDA_OK(cudaMemcpyAsync(out, h->d_audit.p, 16, cudaMemcpyDeviceToHost, h->stream));
    if (reset) CUDA_OK(cudaMemsetAsync(h->d_audit.p, 0, 16, h->stream));
    CUDA_OK(cudaStreamSynchronize(h->stream));
  });
}

int tsr_set_profile(tsr_decoder* h, int on) {
  h->profile = on != 0;
  return TSR_OK;
}

int tsr_profile(tsr_decoder* h, uint64_t out[8], int reset) {
  return guarded([&] {
    require_device(*h);
    CUDA_OK(cudaSetDevice(h->device));
    CUDA_OK(cudaMemcpyAsync(out, h->d_prof.p, 64, cudaMemcpyDeviceToHost, h->stream));
    if (reset) CUDA_OK(cudaMemsetAsync(h->d_prof.p, 0, 64, h->stream));
    CUDA_OK(cudaStreamSynchronize(h->stream));
  });
}

int tsr_last_batch_tail(const tsr_decoder* h, double out[3]) {
  for (int i = 0; i < 3; ++i) out[i] = h->tail[i];
  return TSR_OK;
}

int tsr_last_batch_timing(const tsr_decoder* h, double out[4]) {
  for (int i = 0; i < 4; ++i) out[i] = h->timing[i];
  return TSR_OK;
}

void* tsr_stream(tsr_decoder* h) {
  return h->stream;
}
int tsr_device(const tsr_decoder* h) {
  return h->host_only ? -1 : h->device;
}

// The helpers below size their work by the DEM's detector count (1 + the largest id mentioned): a stray D99999999999 must
// be an error message, not a terabyte allocation.  The decoder itself takes at most 65 534 detectors (ingest.cc).
static void check_detector_count(const DemModel& dem) {
  constexpr uint64_t kMax = uint64_t{1} << 24;
  const uint64_t D = dem.count_detectors();
  if (D > kMax)
    throw std::invalid_argument("the detector error model mentions detector " + std::to_string(D - 1) + "; at most " +
                                std::to_string(kMax) + " detectors are supported here");
}

int tsr_build_det_orders(const char* dem_text, uint64_t num_det_orders, int method, uint64_t seed, uint64_t* out) {
  return guarded([&] {
    if (method < 0 || method > 2) throw std::invalid_argument("Unknown det order method");
    DemModel dem = DemModel::parse(dem_text);
    check_detector_count(dem);
    auto orders = tsr::build_det_orders(dem, num_det_orders, (tsr::DetOrder)method, seed);
    const uint64_t D = dem.count_detectors();
    for (size_t k = 0; k < orders.size(); ++k) std::copy(orders[k].begin(), orders[k].end(), out + k * D);
  });
}

uint64_t tsr_dem_count_detectors(const char* dem_text) {
  uint64_t n = UINT64_MAX;
  guarded([&] { n = DemModel::parse(dem_text).count_detectors(); });
  return n;
}

uint64_t tsr_dem_count_observables(const char* dem_text) {
  uint64_t n = UINT64_MAX;
  guarded([&] { n = DemModel::parse(dem_text).count_observables(); });
  return n;
}

double tsr_merge_weights(double a, double b) {
  return tsr::merge_weights(a, b);
}

int tsr_dem_from_counts(const char* dem_text, const uint64_t* counts, uint64_t n_counts, uint64_t num_shots,
                        char** dem_text_out) {
  return guarded([&] {
    DemModel dem = DemModel::parse(dem_text);  // parsed models are already flat
    if (dem.count_errors() != n_counts)
      throw std::invalid_argument("Error hits array must be the same size as the number of errors in the original DEM.");
    DemModel out;
    uint64_t k = 0;
    for (const tsr::DemInstruction& inst : dem.instructions) {
      if (inst.kind == tsr::DemKind::Error) {
        out.append_error((double)counts[k] / (double)num_shots, inst.targets, inst.tag);
        ++k;
      } else {
        out.instructions.push_back(inst);
      }
    }
    *dem_text_out = dup_text(out.str());
  });
}

int tsr_merge_indistinguishable_errors(const char* dem_text, char** dem_text_out, uint64_t* error_index_map) {
  return guarded([&] {
    std::vector<uint64_t> map;
    DemModel out = tsr::merge_indistinguishable_errors(DemModel::parse(dem_text), map);
    if (error_index_map) std::copy(map.begin(), map.end(), error_index_map);
    if (dem_text_out) *dem_text_out = dup_text(out.str());
  });
}

int tsr_remove_zero_probability_errors(const char* dem_text, char** dem_text_out, uint64_t* error_index_map) {
  return guarded([&] {
    std::vector<uint64_t> map;
    DemModel out = tsr::remove_zero_probability_errors(DemModel::parse(dem_text), map);
    if (error_index_map) std::copy(map.begin(), map.end(), error_index_map);
    if (dem_text_out) *dem_text_out = dup_text(out.str());
  });
}

int64_t tsr_dem_count_errors(const char* dem_text) {
  int64_t n = 0;
  const int rc = guarded([&] { n = (int64_t)DemModel::parse(dem_text).count_errors(); });
  return rc == TSR_OK ? n : rc;
}

int64_t tsr_get_errors_from_dem(const char* dem_text, double* cost, uint64_t* det_offsets, int32_t* dets, uint64_t* obs_offsets,
                                int32_t* obs, uint64_t* n_dets, uint64_t* n_obs) {
  int64_t n = 0;
  const int rc = guarded([&] {
    const std::vector<tsr::CompiledError> errors = tsr::get_errors_from_dem(DemModel::parse(dem_text));
    uint64_t nd = 0, no = 0;
    for (size_t e = 0; e < errors.size(); ++e) {
      if (cost) cost[e] = errors[e].cost;
      if (det_offsets) det_offsets[e] = nd;
      if (obs_offsets) obs_offsets[e] = no;
      for (int d : errors[e].detectors) {
        if (dets) dets[nd] = d;
        ++nd;
      }
      for (int o : errors[e].observables) {
        if (obs) obs[no] = o;
        ++no;
      }
    }
    if (det_offsets) det_offsets[errors.size()] = nd;
    if (obs_offsets) obs_offsets[errors.size()] = no;
    if (n_dets) *n_dets = nd;
    if (n_obs) *n_obs = no;
    n = (int64_t)errors.size();
  });
  return rc == TSR_OK ? n : rc;
}

int64_t tsr_build_detector_graph(const char* dem_text, uint64_t* offsets, uint64_t* neighbors) {
  int64_t n = 0;
  const int rc = guarded([&] {
    const DemModel dem = DemModel::parse(dem_text);
    check_detector_count(dem);
    const auto graph = tsr::build_detector_graph(dem);
    uint64_t at = 0;
    for (size_t d = 0; d < graph.size(); ++d) {
      if (offsets) offsets[d] = at;
      for (uint64_t v : graph[d]) {
        if (neighbors) neighbors[at] = v;
        ++at;
      }
    }
    if (offsets) offsets[graph.size()] = at;
    n = (int64_t)at;
  });
  return rc == TSR_OK ? n : rc;
}

int64_t tsr_get_detector_coords(const char* dem_text, uint64_t* n_rows, uint64_t* offsets, double* coords) {
  int64_t n = 0;
  const int rc = guarded([&] {
    const auto rows = tsr::get_detector_coords(DemModel::parse(dem_text));
    uint64_t at = 0;
    for (size_t r = 0; r < rows.size(); ++r) {
      if (offsets) offsets[r] = at;
      for (double v : rows[r]) {
        if (coords) coords[at] = v;
        ++at;
      }
    }
    if (offsets) offsets[rows.size()] = at;
    if (n_rows) *n_rows = rows.size();
    n = (int64_t)at;
  });
  return rc == TSR_OK ? n : rc;
}

int tsr_circuit_to_dem(const char* circuit_text, char** dem_text_out) {
  return guarded([&] { *dem_text_out = dup_text(tsr::circuit_to_dem_text(circuit_text)); });
}

int tsr_sample_dem_b8(const char* dem_text, uint64_t seed, uint64_t shots, uint8_t* dets_out, uint8_t* obs_out) {
  return guarded([&] {
    DemModel dem = DemModel::parse(dem_text);
    check_detector_count(dem);
    const uint64_t ds = (dem.count_detectors() + 7) / 8, os = (dem.count_observables() + 7) / 8;
    std::memset(dets_out, 0, (size_t)(shots * ds));
    if (obs_out) std::memset(obs_out, 0, (size_t)(shots * os));
    tsr::sample_dem_b8(dem, seed, shots, dets_out, ds, obs_out, os);
  });
}

}  // extern "C"
