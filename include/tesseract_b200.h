/* tesseract_b200 — C ABI of the B200-native Tesseract decode path.
 *
 * This is the drop-in boundary.  The reference has no FFI of its own for this path
 * (its Python module binds the C++ class directly, src/tesseract.pybind.h); each entry point below
 * names the reference C++ interface it stands in for, so that a maintainer can rebind
 * TesseractDecoder's methods onto it (INTEGRATION.md shows the stubs).
 *
 * Conventions: plain C, opaque handles, caller-owned buffers, no torch types.  Functions returning
 * int return TSR_OK (0) or a negative TSR_E_* code; tsr_last_error() then gives the message of the
 * last failure on the calling thread (texts follow the reference's exception texts where it has
 * one).  A decoder handle is not re-entrant (like the reference object, tesseract.h:109-131): use
 * one handle per host thread or serialise calls.  There is NO CPU fallback: every decode entry
 * point fails with TSR_E_CUDA when no sm_100 device is usable.
 */
#ifndef TESSERACT_B200_H
#define TESSERACT_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TSR_OK 0
#define TSR_E_INVALID_ARGUMENT (-1) /* std::invalid_argument in the reference */
#define TSR_E_RUNTIME (-2)          /* std::runtime_error in the reference */
#define TSR_E_CUDA (-3)             /* no device / CUDA failure */
#define TSR_E_DEVICE_MEMORY (-4)    /* a search outgrew every retry tier of the node pool */

#define TSR_INF_DET_BEAM 65535u          /* tesseract.h:33 */
#define TSR_PQLIMIT_UNBOUNDED UINT64_MAX /* std::numeric_limits<size_t>::max(), tesseract_main.cc:503 */
#define TSR_NO_ERROR_INDEX UINT64_MAX    /* "removed" in dem_error_to_error, common.cc:228-239 */

typedef struct tsr_decoder tsr_decoder;

/* Mirrors TesseractConfig (tesseract.h:39-58) minus the DEM (passed as text) and the
 * verbose/visualization switches, which are host-only debugging aids in the reference. */
typedef struct tsr_config {
  int32_t det_beam;        /* default 5; TSR_INF_DET_BEAM = unbounded */
  int32_t beam_climbing;   /* tesseract.cc:307-328 */
  int32_t no_revisit_dets; /* tesseract.cc:475-476, 563-565 */
  int32_t merge_errors;    /* tesseract.cc:162-166 */
  int32_t at_most_two_errors_per_detector; /* README.md:59 only; no reference code — see DESIGN.md */
  int32_t device;          /* CUDA device ordinal; -1 = current device */
  uint64_t pqlimit;        /* default 200000; TSR_PQLIMIT_UNBOUNDED */
  double det_penalty;      /* tesseract.cc:153 */
  uint64_t num_det_orders; /* 0 = one identity order (tesseract.cc:175-177) */
  const uint64_t* det_orders; /* [num_det_orders][num_detectors] row-major, position -> detector */
  uint64_t det_order_len;  /* entries per order; must equal num_detectors (tesseract.cc:178-184); 0 = unchecked */
  /* Device resources; 0 = choose automatically. */
  uint32_t search_slots;       /* concurrent searches (fast path: one CTA each; generic kernel: one warp each) */
  uint32_t warps_per_block;
  uint64_t node_pool_bytes;    /* HBM set aside for heaps / chains / visited sets */
  uint64_t max_batch_shots;    /* shots staged per kernel launch */
  int32_t force_generic_kernel; /* 1 = always use the event-replay kernel (search.cu), even for certified DEMs */
  int32_t row_scan; /* certified DEMs: 0 = choose by DEM shape, 1 = one row entry per lane, 2 = bit-sliced (32 entries per lane) */
  /* --sparsify-errors (tesseract.h:52-55; validated like tesseract.cc:256-277, same messages): per shot only the errors of
   * degree <= base plus the `limit` best-overlapping errors of degree <= max take part (tesseract.cc:673-737). */
  int32_t sparsify_errors;
  int32_t sparsify_base_degree;      /* must be > 0 when sparsify_errors */
  int32_t sparsify_max_degree;       /* -1 = no upper bound */
  int32_t sparsify_reactivate_limit; /* -1 = suggest_sparsify_reactivate_limit (tesseract.cc:45-69) capped at num_errors */
  /* TesseractConfig::create_visualization (tesseract.h:50): tsr_decode calls record the visualization log of
   * visualization.cc (read it with tsr_visualization_text).  A debugging aid for single shots: batch calls do not log. */
  int32_t create_visualization;
} tsr_config;

/* Fills the reference's C++ defaults (tesseract.h:39-58). */
void tsr_config_init(tsr_config* cfg);

const char* tsr_last_error(void);
const char* tsr_version(void);

/* ---- construction: TesseractDecoder::TesseractDecoder(TesseractConfig) (tesseract.cc:156-205) ----
 * dem_text is what str(stim.DetectorErrorModel) prints (stim_utils.pybind.h:22-26).  Ingestion runs on
 * the host; the resulting tables are uploaded to `device`.  With TSR_CREATE_HOST_ONLY in `flags` no
 * CUDA call is made: the handle serves the table getters, cost_from_errors and
 * get_flipped_observables only, and every decode call fails with TSR_E_CUDA. */
#define TSR_CREATE_HOST_ONLY 1u
int tsr_create(const char* dem_text, const tsr_config* cfg, uint32_t flags, tsr_decoder** out);
void tsr_destroy(tsr_decoder* h);

/* Public fields of TesseractDecoder (tesseract.h:112-130). */
uint64_t tsr_num_detectors(const tsr_decoder* h);
uint64_t tsr_num_observables(const tsr_decoder* h);
uint64_t tsr_num_errors(const tsr_decoder* h);     /* errors.size() */
uint64_t tsr_num_dem_errors(const tsr_decoder* h); /* dem_error_to_error.size() */
uint64_t tsr_num_det_orders(const tsr_decoder* h); /* config.det_orders.size() */
/* config.sparsify_reactivate_limit after construction (the suggested value when -1 was passed, tesseract.cc:272-277);
 * -1 when sparsify_errors is off.  tsr_suggest_sparsify_reactivate_limit is the free function of tesseract.h:37. */
int64_t tsr_sparsify_reactivate_limit(const tsr_decoder* h);
int64_t tsr_suggest_sparsify_reactivate_limit(uint64_t num_detectors, int32_t sparsify_base_degree);
int tsr_error_costs(const tsr_decoder* h, double* out /* [num_errors] */);
int tsr_maps(const tsr_decoder* h, uint64_t* dem_error_to_error /* [num_dem_errors] or NULL */,
             uint64_t* error_to_dem_error /* [num_errors] or NULL */);
/* which: 0 = d2e rows (sorted, tesseract.cc:223-227), 1 = edets, 2 = eneighbors (get_eneighbors()),
 * 3 = per-error observables.  Either pointer may be NULL.  Returns the index count or a negative code. */
int64_t tsr_csr(const tsr_decoder* h, int which, uint64_t* offsets, int32_t* idx);
/* The compiled DEM (config.dem after merge / zero-strip) as text; caller frees with tsr_free. */
int tsr_compiled_dem_text(const tsr_decoder* h, char** out);
void tsr_free(void* p);

/* ---- decode --------------------------------------------------------------------------------------
 * tsr_decode_batch_b8: the batched hot path behind decode_shots (tesseract.cc:665-671),
 * decode_batch (tesseract.pybind.h:450-511) and decode_shots_bit_packed
 * (tesseract_sinter_compat.pybind.h:45-97).  dets: [shots][stride] bytes, bit k of byte j = detector
 * 8j+k.  All outputs are optional (NULL to skip): obs_out [shots][ceil(O/8)] same bit order;
 * cost_out [shots] = cost_from_errors(predicted_errors_buffer); low_conf_out [shots].  The per-shot
 * error lists (predicted_errors_buffer: original flattened-DEM indices, root-to-leaf order) stay in
 * the handle until the next decode and are read with tsr_last_batch_errors.
 * The *_device variant takes and fills device pointers on the handle's device (inputs already in
 * HBM) and enqueues on tsr_stream(h) without a final host sync of the outputs' consumers. */
int tsr_decode_batch_b8(tsr_decoder* h, const uint8_t* dets, uint64_t stride, uint64_t shots,
                        uint8_t* obs_out, double* cost_out, uint8_t* low_conf_out);
int tsr_decode_batch_b8_device(tsr_decoder* h, const uint8_t* d_dets, uint64_t stride, uint64_t shots,
                               uint8_t* d_obs_out, double* d_cost_out, uint8_t* d_low_conf_out);
/* Whether batch decodes keep the per-shot error lists for tsr_last_batch_errors (default 1).  Callers that
 * only consume observables (decode_batch, decode_shots_bit_packed) switch it off and save the D2H copy. */
int tsr_set_collect_errors(tsr_decoder* h, int on);
uint64_t tsr_last_batch_nnz(const tsr_decoder* h);
int tsr_last_batch_errors(const tsr_decoder* h, uint64_t* offsets /* [shots+1] */, uint64_t* idx /* [nnz] */);

/* tsr_decode: one shot from a sparse list of fired detectors.
 * order < 0: decode_to_errors(detections) — the det-order ensemble / beam climbing (tesseract.cc:295-347).
 * order >= 0: decode_to_errors(detections, order, beam) (tesseract.cc:371-378).
 * errs_out receives predicted_errors_buffer (fails with TSR_E_INVALID_ARGUMENT if errs_cap is too
 * small; *n_errs is set either way).  A detector id >= num_detectors fails with TSR_E_RUNTIME and the
 * reference's message (tesseract.cc:400-404). */
int tsr_decode(tsr_decoder* h, const uint64_t* detections, uint64_t n, int64_t order, uint64_t beam,
               uint64_t* errs_out, uint64_t errs_cap, uint64_t* n_errs, double* cost, int32_t* low_confidence);

/* The Visualizer's lines (visualization.cc; what Visualizer::write puts in the file, '\n'-terminated): "Detector D<d>
 * coordinate (...)" and "Error{cost=..., symptom=Symptom{...}}" headers written at construction (tesseract.cc:200-204),
 * then per tsr_decode call and per trial of the reference's loop, for every node the search expanded and for the goal,
 * "activated_errors = e, e, ..., " (leaf to root) and "activated_detectors = d, d, ..., " (tesseract.cc:442-445, 478-481).
 * Empty unless the decoder was created with create_visualization.  Caller frees *out with tsr_free. */
int tsr_visualization_text(const tsr_decoder* h, char** out);

/* cost_from_errors (tesseract.cc:618-628) and get_flipped_observables (tesseract.cc:630-658) over
 * original flattened-DEM error indices; both fail with TSR_E_INVALID_ARGUMENT ("error index does not
 * map to a retained decoder error") on removed errors.  obs_bits_out: ceil(O/8) bytes. */
int tsr_cost_from_errors(const tsr_decoder* h, const uint64_t* errs, uint64_t n, double* out);
int tsr_flipped_observables(const tsr_decoder* h, const uint64_t* errs, uint64_t n, uint8_t* obs_bits_out);

/* Which search kernel this handle decodes with: 2 = the CTA-per-search integer-rank kernel with bit-sliced row
 * scans, 1 = the same kernel scanning one row entry per lane (search_fast.cu; both need the DEM to pass the
 * host-side certificate of csrc/fast.h); 0 = the generic kernel (search.cu) that replays get_detcost's
 * floating-point scan event by event.  All are CUDA; results are identical.
 * tsr_search_path_reason says why a DEM was not certified ("" when it was). */
int tsr_search_path(const tsr_decoder* h);
const char* tsr_search_path_reason(const tsr_decoder* h);
/* Fast path only: also count the row entries the reference's early-exit loop would visit (counter S below);
 * off by default because it costs a warp scan per 32 entries.  The generic kernel always counts. */
int tsr_set_instrumented(tsr_decoder* h, int on);

/* Host-only self check (no GPU): evaluates get_detcost (tesseract.cc:128-154) on `trials` random search states
 * three ways — the reference's floating-point loop, the rank minimum over the packed row records, and the
 * bit-sliced evaluation the sm_100a kernel performs — and returns the number of evaluations that are not
 * bit-identical (0 is the only acceptable answer), or a negative code.  *checked receives how many were compared
 * (0 when the DEM is not certified for the fast path). */
int64_t tsr_cross_check_detcost(const tsr_decoder* h, uint64_t trials, uint64_t seed, uint64_t* checked);

/* Search counters since the last reset, summed over every search this handle ran:
 * [0]=Q pushes [1]=P pops [2]=X expanded nodes [3]=L chain-walk steps [4]=S scanned row entries
 * [5]=A sum(|edets|+|fired neighbours|) over evaluated children [6]=V visited inserts+probes
 * [7]=searches run (distinct (order, beam) trials only). */
int tsr_counters(tsr_decoder* h, uint64_t out[8], int reset);
/* Visited-set audit (certified DEMs, no_revisit_dets).  The device's visited set holds 128-bit Zobrist fingerprints of
 * the fired-detector sets; the reference holds the sets themselves (tesseract.cc:394, 475-476, 563-565).  With the audit
 * on, every fingerprint hit — at a pop and at every child probe — is checked against the truth: the fired set of the node
 * that made the entry is rebuilt from its error chain and compared bit for bit with the candidate's.  out[0] = hits
 * checked, out[1] = hits whose sets differ (a fingerprint collision; two distinct sets collide with probability 2^-128
 * per pair).  Slower; results are unchanged. */
int tsr_set_visited_audit(tsr_decoder* h, int on);
int tsr_visited_audit(tsr_decoder* h, uint64_t out[2], int reset);
/* Phase profile of the CTA-per-search kernel (diagnostic; off by default, a few clock reads per expansion when on).
 * Ticks of the SM clock summed over every search CTA since the last reset, by phase of an expansion
 * (csrc/search_fast.cu): [0]=pop [1]=state rebuild + goal / visited test [2]=detector terms of the node
 * [3]=min detector + child list [4]=children (all warps) [5]=pushes [6]=shot staging, work claiming, table
 * cleanup [7]=expansions (a count, not ticks). */
int tsr_set_profile(tsr_decoder* h, int on);
int tsr_profile(tsr_decoder* h, uint64_t out[8], int reset);
/* Timing of the last batch, measured with CUDA events on tsr_stream(h): [0]=search kernel ms,
 * [1]=whole device pipeline ms (stage + search + resolve), [2]=search kernel launches,
 * [3]=items rerun in a larger node-pool tier. */
int tsr_last_batch_timing(const tsr_decoder* h, double out[4]);
/* The tail of the last batch (certified DEMs): [0]=searches that were still running when the work queue emptied and were
 * handed to wide CTAs, [1]=ms of that second launch (part of the search ms above), [2]=CTAs per search in it (thread-block
 * cluster size; 1 = no clusters). */
int tsr_last_batch_tail(const tsr_decoder* h, double out[3]);
void* tsr_stream(tsr_decoder* h); /* cudaStream_t the decoder launches on */
int tsr_device(const tsr_decoder* h);

/* ---- multi-GPU (csrc/group.cu) -------------------------------------------------------------------------
 * Shots are independent (tesseract.cc:665-671); the reference shards them over host threads with one decoder each
 * (parallel_for_shots_in_order, utils.h:75-115; tesseract_main.cc:559-616) and sums three counters.  Here every GPU
 * holds its own copy of the DEM tables and decodes its own shots; the only exchange is ONE ncclAllReduce(sum) of
 * {shots, logical errors, low confidence} — plus the error-use histogram of --dem-out when asked for
 * (tesseract_main.cc:618-623) — over NVLink.  NCCL (libnccl.so.2) is loaded on first use.
 *
 * One process per GPU: rank 0 calls tsr_comm_unique_id and ships the 128 bytes to the other ranks (launcher's
 * business); every rank then calls tsr_comm_init_rank on its decoder (collective) and, per batch, sums its counters
 * with tsr_comm_allreduce_sum_u64 (in place, host values, one collective on tsr_stream(h)). */
#define TSR_COMM_UNIQUE_ID_BYTES 128
int tsr_comm_unique_id(uint8_t out[TSR_COMM_UNIQUE_ID_BYTES]);
int tsr_comm_init_rank(tsr_decoder* h, const uint8_t id[TSR_COMM_UNIQUE_ID_BYTES], int32_t rank, int32_t world);
int32_t tsr_comm_world(const tsr_decoder* h); /* 0 = no communicator */
int tsr_comm_allreduce_sum_u64(tsr_decoder* h, uint64_t* values, uint64_t n);
int tsr_nccl_version(void); /* ncclGetVersion code of the library in use, 0 if NCCL cannot be loaded */

/* One process, several GPUs (the CLI's --gpus N): one decoder, host thread and communicator per listed device
 * (devices == NULL: 0 .. n_devices-1).  tsr_group_decode_b8 deals the batch out in interleaved chunks of chunk_shots
 * shots (0 = 16384; chunk k goes to member k mod n — per-shot cost is heavy-tailed, so interleaving balances better
 * than n contiguous ranges), decodes them concurrently and writes every shot's outputs at its own position in the
 * caller's host arrays (all nullable).  obs_true (nullable, [shots][ceil(O/8)]) enables the logical-error count:
 * tally_out = {shots, shots whose prediction differs from obs_true and is not low-confidence, low-confidence shots}
 * (tesseract_main.cc:580-601), summed over the members by the all-reduce.  error_use_out (nullable,
 * [num_dem_errors]) receives how often each error was predicted in correctly decoded shots (tesseract_main.cc:586-590). */
typedef struct tsr_group tsr_group;
int tsr_group_create(const char* dem_text, const tsr_config* cfg, const int32_t* devices, uint32_t n_devices, tsr_group** out);
void tsr_group_destroy(tsr_group* g);
uint32_t tsr_group_size(const tsr_group* g);
tsr_decoder* tsr_group_member(tsr_group* g, uint32_t i); /* borrowed; for the table getters, counters and timing */
int tsr_group_decode_b8(tsr_group* g, const uint8_t* dets, uint64_t stride, uint64_t shots, uint64_t chunk_shots,
                        const uint8_t* obs_true, uint8_t* obs_out, double* cost_out, uint8_t* low_conf_out,
                        uint64_t tally_out[3], uint64_t* error_use_out);
/* predicted_errors_buffer of every shot of the last tsr_group_decode_b8 call that asked for error_use_out, in the
 * caller's shot order (offsets [shots+1], idx [nnz]). */
uint64_t tsr_group_last_batch_nnz(const tsr_group* g);
int tsr_group_last_batch_errors(const tsr_group* g, uint64_t* offsets, uint64_t* idx);

/* ---- helpers either side of the path -------------------------------------------------------------*/
/* build_det_orders (utils.cc:192-205); method 0=DetBFS 1=DetIndex 2=DetCoordinate; out [n][D]. */
int tsr_build_det_orders(const char* dem_text, uint64_t num_det_orders, int method, uint64_t seed, uint64_t* out);
uint64_t tsr_dem_count_detectors(const char* dem_text);   /* UINT64_MAX on parse failure */
uint64_t tsr_dem_count_observables(const char* dem_text); /* UINT64_MAX on parse failure */
double tsr_merge_weights(double a, double b);           /* common.cc:118-123 */
/* common::dem_from_counts (common.cc:241-263): the flattened DEM with error k's probability replaced by
 * counts[k] / num_shots; n_counts must equal the number of error instructions (TSR_E_INVALID_ARGUMENT otherwise).
 * Used by the CLI's --dem-out (tesseract_main.cc:625-639).  Caller frees *dem_text_out with tsr_free. */
int tsr_dem_from_counts(const char* dem_text, const uint64_t* counts, uint64_t n_counts, uint64_t num_shots,
                        char** dem_text_out);
/* The ingestion stages on their own — what tesseract_decoder.common / .utils export next to the decoder
 * (common.pybind.h:143-213, utils.pybind.h:43-135); DEM text in, flattened DEM text out (caller frees with tsr_free;
 * dem_text_out may be NULL).  Host only, no device needed.
 *   tsr_merge_indistinguishable_errors  common::merge_indistinguishable_errors (common.cc:139-190): errors with the same
 *       symptom merged with merge_weights, first-seen order, DETECTOR / LOGICAL_OBSERVABLE instructions first;
 *       error_index_map [tsr_dem_count_errors(dem_text)] (may be NULL) = merged index of every input error.
 *   tsr_remove_zero_probability_errors  common::remove_zero_probability_errors (common.cc:192-218);
 *       error_index_map[i] = output index, or UINT64_MAX for a dropped error. */
int tsr_merge_indistinguishable_errors(const char* dem_text, char** dem_text_out, uint64_t* error_index_map);
int tsr_remove_zero_probability_errors(const char* dem_text, char** dem_text_out, uint64_t* error_index_map);
int64_t tsr_dem_count_errors(const char* dem_text); /* error instructions of the flattened DEM; < 0 = status */
/* get_errors_from_dem (utils.cc:253-261): the errors with probability > 0 as common::Error values (common.cc:52-88:
 * likelihood_cost = -log(p/(1-p)); detectors and observables XOR-reduced, ascending).  Returns their number (< 0 =
 * status).  Call once with NULL arrays for the sizes (*n_dets, *n_obs), then with cost [n], det_offsets [n+1],
 * dets [*n_dets], obs_offsets [n+1], obs [*n_obs].  Every pointer may be NULL. */
int64_t tsr_get_errors_from_dem(const char* dem_text, double* cost, uint64_t* det_offsets, int32_t* dets,
                                uint64_t* obs_offsets, int32_t* obs, uint64_t* n_dets, uint64_t* n_obs);
/* build_detector_graph (utils.cc:60-87): detectors are neighbours when one error flips both; sorted, unique.
 * offsets [D+1], neighbors [returned count]; either may be NULL (sizing call).  Returns the count, < 0 = status. */
int64_t tsr_build_detector_graph(const char* dem_text, uint64_t* offsets, uint64_t* neighbors);
/* get_detector_coords (utils.cc:32-58): the coordinate list of every DETECTOR instruction, in DEM order.
 * *n_rows = number of DETECTOR instructions; offsets [*n_rows+1], coords [returned count]; pointers may be NULL. */
int64_t tsr_get_detector_coords(const char* dem_text, uint64_t* n_rows, uint64_t* offsets, double* coords);
/* Stand-ins for the stim services the reference CLI uses (tesseract_main.cc:191-211, utils.cc:207-251):
 * circuit text -> DEM text (caller frees with tsr_free) and an i.i.d. DEM shot sampler into bit-packed
 * rows ([shots][ceil(D/8)], [shots][ceil(O/8)]; obs_out may be NULL). */
int tsr_circuit_to_dem(const char* circuit_text, char** dem_text_out);
int tsr_sample_dem_b8(const char* dem_text, uint64_t seed, uint64_t shots, uint8_t* dets_out, uint8_t* obs_out);

#ifdef __cplusplus
}
#endif
#endif /* TESSERACT_B200_H */
