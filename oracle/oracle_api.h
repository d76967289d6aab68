/* TEST INFRASTRUCTURE — NOT PRODUCT CODE.
 *
 * One C interface, two CPU implementations of the reference decode path, used only as the
 * checker by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs:
 *
 *   oracle/_ref/libtesseract_ref.so   the reference's own, unmodified src/{tesseract,common,utils,
 *                                     visualization}.cc compiled against oracle/shim (ref_capi.cc)
 *   oracle/_build/libtesseract_port.so an independent CPU restatement of the same algorithm with
 *                                     search counters (port.cc)
 *
 * Both export these symbols (the block marked REFERENCE BUILD ONLY excepted).  All functions returning int return 0 on success and -1 on
 * failure, in which case orc_last_error() gives the exception text (thread-local).
 */
#ifndef ORACLE_API_H
#define ORACLE_API_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct orc_handle orc_handle;

typedef struct orc_config {
  int32_t det_beam;        /* 65535 = INF_DET_BEAM (tesseract.h:33) */
  int32_t beam_climbing;   /* tesseract.h:42 */
  int32_t no_revisit_dets; /* tesseract.h:43 */
  int32_t merge_errors;    /* tesseract.h:46 */
  uint64_t pqlimit;        /* UINT64_MAX = unbounded (tesseract.h:47) */
  double det_penalty;      /* tesseract.h:49 */
  uint64_t num_det_orders; /* 0 -> the decoder's default single identity order (tesseract.cc:175-177) */
  const uint64_t* det_orders; /* [num_det_orders][num_detectors], row-major */
  int32_t num_threads;     /* decoders in the pool; one per worker like tesseract_main.cc:559-571 */
  int32_t sparsify_errors; /* tesseract.h:52-55, tesseract.cc:256-292, 673-737 */
  int32_t sparsify_base_degree;
  int32_t sparsify_max_degree;
  int32_t sparsify_reactivate_limit;
  int32_t at_most_two_errors_per_detector; /* README.md:59 only — no reference code: the port implements it (port.cc: Config),
                                              the reference build refuses it */
  int32_t create_visualization; /* tesseract.h:50: record the Visualizer log (visualization.cc) */
} orc_config;

const char* orc_last_error(void);
const char* orc_kind(void); /* "reference" or "port" */

orc_handle* orc_create(const char* dem_text, const orc_config* cfg);
void orc_destroy(orc_handle* h);

uint64_t orc_num_detectors(const orc_handle* h);
uint64_t orc_num_observables(const orc_handle* h);
uint64_t orc_num_errors(const orc_handle* h);     /* compiled (merged, stripped) errors */
uint64_t orc_num_dem_errors(const orc_handle* h); /* error instructions of the flattened input DEM */

/* Ingestion results, for table parity. */
int orc_error_costs(const orc_handle* h, double* out);
int orc_maps(const orc_handle* h, uint64_t* dem_error_to_error, uint64_t* error_to_dem_error);
/* which: 0=d2e (host-sorted rows) 1=edets 2=eneighbors 3=error observables.  Either pointer may be
 * NULL.  Returns the number of stored indices, or -1. */
int64_t orc_csr(const orc_handle* h, int which, uint64_t* offsets, int32_t* idx);

/* One shot, sparse hits in caller order.  order < 0: decode_to_errors(hits) (the ensemble,
 * tesseract.cc:295-347); order >= 0: decode_to_errors(hits, order, beam) (tesseract.cc:371-378).
 * cost = cost_from_errors(predicted_errors_buffer). */
int orc_decode(orc_handle* h, const uint64_t* hits, uint64_t n_hits, int64_t order, uint64_t beam,
               uint64_t* errs_out, uint64_t errs_cap, uint64_t* n_errs, double* cost,
               int32_t* low_confidence);

/* Batch of bit-packed shots (bit k of byte j = detector 8j+k, tesseract_sinter_compat.pybind.h:77),
 * decoded by num_threads workers claiming shots from an atomic counter (utils.h:88-92).
 * obs_out: [shots][ceil(O/8)] same bit order.  The per-shot error lists are kept in the handle and
 * fetched with orc_last_batch_errors.  wall_seconds covers the decode loop only;
 * sum_shot_seconds is the reference's total_time_seconds convention (tesseract_main.cc:574-602). */
int orc_decode_batch_b8(orc_handle* h, const uint8_t* dets, uint64_t stride, uint64_t shots,
                        uint8_t* obs_out, double* cost_out, uint8_t* low_conf_out,
                        double* wall_seconds, double* sum_shot_seconds);
uint64_t orc_last_batch_nnz(const orc_handle* h);
int orc_last_batch_errors(const orc_handle* h, uint64_t* offsets, uint64_t* idx);

int orc_cost_from_errors(const orc_handle* h, const uint64_t* errs, uint64_t n, double* out);
int orc_flipped_observables(const orc_handle* h, const uint64_t* errs, uint64_t n, uint8_t* obs_bits_out);

/* utils.cc:192-205.  method: 0=DetBFS 1=DetIndex 2=DetCoordinate (utils.h:43). */
int orc_build_det_orders(const char* dem_text, uint64_t num_det_orders, int method, uint64_t seed,
                         uint64_t* out);
double orc_merge_weights(double a, double b); /* common.cc:118-123 */

/* ---- REFERENCE BUILD ONLY (libtesseract_ref.so): the reference's standalone ingestion functions, for the parity tests of
 * the product's tsr_merge_indistinguishable_errors / tsr_remove_zero_probability_errors / tsr_get_errors_from_dem /
 * tsr_build_detector_graph / tsr_get_detector_coords.  The port restates these inside its own constructor only. ---- */
/* The instructions of a DEM after `which`: 0 = common::flatten, 1 = common::merge_indistinguishable_errors
 * (common.cc:139-190), 2 = common::remove_zero_probability_errors (common.cc:192-218).  kind [n]: 0 error 1 detector
 * 2 logical_observable; args / targets as CSR (target = detector id, observable id | 1<<63, UINT64_MAX for '^');
 * index_map = the function's error_index_map (UINT64_MAX = dropped); sizes = {n, n_args, n_targets, map length}.
 * Every pointer may be NULL (sizing call).  Returns n or -1. */
int64_t orc_dem_dump(const char* dem_text, int which, uint64_t* index_map, uint8_t* kind, uint64_t* arg_off, double* args,
                     uint64_t* tgt_off, uint64_t* tgt, uint64_t sizes[4]);
/* get_errors_from_dem (utils.cc:253-261) of the flattened DEM; same array convention as tsr_get_errors_from_dem. */
int64_t orc_get_errors_from_dem(const char* dem_text, double* cost, uint64_t* det_offsets, int32_t* dets,
                                uint64_t* obs_offsets, int32_t* obs, uint64_t* n_dets, uint64_t* n_obs);
int64_t orc_build_detector_graph(const char* dem_text, uint64_t* offsets, uint64_t* neighbors); /* utils.cc:60-87 */
int64_t orc_get_detector_coords(const char* dem_text, uint64_t* n_rows, uint64_t* offsets, double* coords); /* utils.cc:32-58 */
/* common::Error(cost, detectors, observables).str() and .get_probability() (common.cc:88-98); set_with_probability
 * (common.cc:100-105). */
int orc_error_text(double likelihood_cost, const int32_t* dets, uint64_t n_dets, const int32_t* obs, uint64_t n_obs, char* out,
                   uint64_t cap, double* probability);
int orc_cost_of_probability(double p, double* cost);

/* Search counters summed over every search run through this handle since the last reset
 * (SURVEY.md §8d): out[0]=Q pushes, [1]=P pops, [2]=X expanded, [3]=L chain-walk steps,
 * [4]=S get_detcost loop iterations, [5]=A sum(|edets|+|eneighbors|) over evaluated children,
 * [6]=V visited inserts+probes, [7]=searches.  The reference build reports zeros. */
int orc_counters(orc_handle* h, uint64_t out[8], int reset);

/* Writes the visualization log of decoder 0 of the pool (Visualizer::write, visualization.cc:53-62): header lines plus
 * whatever orc_decode calls have logged so far. */
int orc_visualization_write(orc_handle* h, const char* path);

/* Synthetic i.i.d. shots of a DEM (oracle/sampler.inc): the bench / test workload generator restated on the
 * checker side; bit-identical to the product's tsr_sample_dem_b8.  dets_out [shots][ceil(D/8)],
 * obs_out [shots][ceil(O/8)] or NULL. */
int orc_sample_dem_b8(const char* dem_text, uint64_t seed, uint64_t shots, uint8_t* dets_out, uint8_t* obs_out);

#ifdef __cplusplus
}
#endif
#endif
