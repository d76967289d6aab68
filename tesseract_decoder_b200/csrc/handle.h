// The decoder handle behind the C ABI (include/tesseract_b200.h) and the helpers shared by its translation units:
// capi.cu (construction, the launch pipeline, single-device entry points) and group.cu (NCCL: multi-GPU groups and
// per-rank communicators).
#pragma once
#include <cuda_runtime.h>

#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/tesseract_b200.h"
#include "fast.h"
#include "ingest.h"
#include "search.h"

namespace tsr_impl {

// Message of the last failure on the calling thread (tsr_last_error).
std::string& last_error();

struct CudaError : std::runtime_error {
  using std::runtime_error::runtime_error;
};
struct DeviceMemoryError : std::runtime_error {
  using std::runtime_error::runtime_error;
};

inline void cuda_check(cudaError_t e, const char* what) {
  if (e != cudaSuccess) throw CudaError(std::string("CUDA: ") + what + ": " + cudaGetErrorString(e));
}
#define CUDA_OK(expr) ::tsr_impl::cuda_check((expr), #expr)

template <typename F>
int guarded(F&& f) {
  try {
    f();
    return TSR_OK;
  } catch (const std::invalid_argument& e) {
    last_error() = e.what();
    return TSR_E_INVALID_ARGUMENT;
  } catch (const CudaError& e) {
    last_error() = e.what();
    return TSR_E_CUDA;
  } catch (const DeviceMemoryError& e) {
    last_error() = e.what();
    return TSR_E_DEVICE_MEMORY;
  } catch (const std::exception& e) {
    last_error() = e.what();
    return TSR_E_RUNTIME;
  }
}

// A device allocation that frees itself.
struct DevBuf {
  void* p = nullptr;
  size_t bytes = 0;
  DevBuf() = default;
  DevBuf(const DevBuf&) = delete;
  DevBuf& operator=(const DevBuf&) = delete;
  ~DevBuf() { release(); }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    bytes = 0;
  }
  void ensure(size_t n) {
    if (n <= bytes) return;
    release();
    CUDA_OK(cudaMalloc(&p, n));
    bytes = n;
  }
  template <typename T>
  T* as() const {
    return static_cast<T*>(p);
  }
  // Pageable H2D copy on the legacy stream.  The decoder's own stream is non-blocking, so it is NOT ordered after this
  // copy (and cudaMemcpy may return before the DMA of a small pageable buffer has landed): callers finish a group
  // of uploads with uploads_done() before anything is launched on another stream.
  template <typename T>
  void upload(const std::vector<T>& v) {
    ensure(std::max<size_t>(v.size() * sizeof(T), 16));
    if (!v.empty()) CUDA_OK(cudaMemcpy(p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice));
  }
  static void uploads_done() { CUDA_OK(cudaStreamSynchronize(cudaStreamLegacy)); }
};

struct Tier {
  uint32_t slots;
  uint64_t node_cap, visit_cap;
};

// The (order, beam) trials of one decode call: the reference's literal loop (tesseract.cc:307-344)
// and the distinct searches behind it (identical detector orders are searched once).
struct TrialSet {
  std::vector<uint32_t> perm, beam;        // distinct trials
  std::vector<uint32_t> literal_to_trial;  // literal trial -> distinct trial
  DevBuf d_perm, d_beam, d_literal;
  void upload() {
    d_perm.upload(perm);
    d_beam.upload(beam);
    d_literal.upload(literal_to_trial);
    DevBuf::uploads_done();
  }
};

}  // namespace tsr_impl

struct tsr_comm;  // NCCL communicator of a decoder (group.cu)

struct tsr_decoder {
  tsr_config cfg{};
  tsr::DemTables T;
  tsr::FastTables F;
  tsr::BitTables B;
  bool bits = false;          // fast path: bit-sliced row scans (32 row entries per lane)
  bool fast = false;          // decode with search_fast.cu (CTA per search) instead of search.cu
  bool instrumented = false;  // fast path: also count the reference's scanned row entries (counter S)
  bool profile = false;       // fast path: per-phase clock ticks (tsr_profile)
  bool visited_audit = false; // fast path: check every visited-set fingerprint hit against the real fired sets
  bool hc_shared = false;
  std::vector<std::vector<uint64_t>> det_orders;
  std::vector<uint32_t> order_to_perm;  // det order index -> distinct permutation id
  uint32_t n_perms = 0;
  bool host_only = true;
  int device = 0;
  int sm_count = 0;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};

  // Device tables.
  tsr_impl::DevBuf row_off, row_err, rec_a, rec_b, rec_c, e_off, e_det, e_rank, e_cost, e_nnbr, eo_off, eo_obs, e2dem, adj,
      zobrist, order_det, order_pos, frec_a, frec_b, frank, fquot, bs_meta, bs_adjpre, bs_masks, bs_pl_off, bs_pl,
      sp_mand, sp_opt, sp_crank, sp_act;
  int64_t sparsify_limit = -1;  // effective sparsify_reactivate_limit (tesseract.cc:272-277); -1 = sparsify off
  tsr::DeviceTables dt;
  tsr_impl::TrialSet ensemble;

  // Node pool and tiers.
  uint32_t wpb = 4, blocks = 0;
  // Tail widening (fast path; SearchParams::expansion_budget): searches that outlive the budget are resumed by CTAs of
  // wpb_wide warps, and the retry tiers (few, long searches) run wide from the start.  0 = off.
  uint32_t wpb_wide = 0, blocks_wide = 0, expansion_budget = 0, stage_budget = 0, cluster_max = 1;
  std::vector<tsr_impl::Tier> tiers;
  tsr_impl::DevBuf pool, hcache, vlist;
  uint32_t vlist_cap = 256;

  // Per-batch buffers.
  uint64_t max_batch = 0;
  tsr_impl::DevBuf d_dets, d_status, d_sol_len, d_fcost, d_sol, d_obs, d_cost, d_low, d_err_off, d_err_len, d_err, d_retry,
      d_items, d_scalars, d_counters, d_prof, d_fired, d_order, d_bins, d_paused, d_paused2, d_stage_counts, d_vchain, d_audit, d_audit_scratch, d_sol_retry;
  uint32_t sol_stride = 0;
  bool collect_errors = true;

  // Visualization log (cfg.create_visualization): header + one pair of lines per logged node, like Visualizer::lines.
  std::vector<std::string> viz_lines;
  bool trace_on = false;             // set around the single-shot decode of tsr_decode
  uint32_t trace_cap = 0;
  tsr_impl::DevBuf d_trace, d_trace_len, d_trace_arena;

  // Results of the last decode call.
  std::vector<uint64_t> last_off{0};
  std::vector<uint64_t> last_idx;
  double timing[4] = {0, 0, 0, 0};
  double tail[3] = {0, 0, 0};  // tsr_last_batch_tail: paused searches, ms of the resume launch, largest cluster size used

  tsr_comm* comm = nullptr;  // set by tsr_comm_init_rank / tsr_group_create (group.cu)

  ~tsr_decoder();
};

namespace tsr_impl {
// The batched decode behind tsr_decode_batch_b8[_device] (capi.cu): `dets` is a host pointer, or a device pointer on the
// handle's device when on_device; outputs likewise (nullable).  Error lists are left in h.last_off / h.last_idx.
void decode_batch(tsr_decoder& h, const uint8_t* dets, uint64_t stride, uint64_t shots, bool on_device, const TrialSet& ts,
                  uint8_t* obs, double* cost, uint8_t* low);
void require_device(const tsr_decoder& h);
void comm_release(tsr_decoder& h);  // group.cu: destroys h.comm if any
}  // namespace tsr_impl


