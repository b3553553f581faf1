/* peaksegdisk_b200.h -- C ABI of the B200-native PeakSegFPOP solver (libpeaksegdisk_b200.so).
 *
 * Drop-in boundary: the reference's only native entry for this path is
 *     int PeakSegFPOP_disk(char *bedGraph_file_name, char *penalty_str, char *db_file_name);
 * declared in tdhock/PeakSegDisk src/PeakSegFPOPLog.h:15 and called from src/interface.cpp:15
 * (the .C routine "PeakSegFPOP_interface" that R/PeakSegFPOP_file.R:66-71 invokes).
 * This library exports that exact C++-mangled symbol (_Z16PeakSegFPOP_diskPcS_S_) so an unchanged
 * interface.cpp links against it, plus the plain-C entry points below.  No torch types, only
 * pointers and sizes.  All entry points are thread-safe for distinct plans / distinct file names.
 *
 * There is no CPU fallback: every solve runs the sm_100a kernels and fails with PSD_ERR_CUDA
 * when no usable GPU / driver is present.
 */
#ifndef PEAKSEGDISK_B200_H
#define PEAKSEGDISK_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- status codes ---------------------------------------------------------------------------
 * 0..11 are the reference's (src/PeakSegFPOPLog.h:3-13), same meaning, same order of checks. */
#define PSD_OK 0
#define PSD_ERR_PENALTY_NOT_FINITE 1
#define PSD_ERR_PENALTY_NEGATIVE 2
#define PSD_ERR_UNABLE_TO_OPEN_BEDGRAPH 3
#define PSD_ERR_NOT_ENOUGH_COLUMNS 4
#define PSD_ERR_NON_INTEGER_DATA 5
#define PSD_ERR_INCONSISTENT_CHROMSTART_CHROMEND 6
#define PSD_ERR_WRITING_COST_FUNCTIONS 7
#define PSD_ERR_WRITING_LOSS_OUTPUT 8
#define PSD_ERR_NO_DATA 9
#define PSD_ERR_PENALTY_NOT_NUMERIC 10
#define PSD_ERR_WRITING_SEGMENTS_OUTPUT 11
/* ours (never produced by the reference) */
#define PSD_ERR_PIECE_OVERFLOW 101   /* a cost function outgrew the largest piece-list tier */
#define PSD_ERR_STORE_EXHAUSTED 102  /* the HBM cost-function store cannot hold this problem */
#define PSD_ERR_BACKTRACK 103
#define PSD_ERR_INTERNAL 104
#define PSD_ERR_CUDA 110             /* no device, driver error or kernel fault: see psd_last_error() */
#define PSD_ERR_ARG 111

/* Replaces PeakSegFPOP_disk (src/PeakSegFPOPLog.cpp:143-463): reads the 4-column bedGraph, solves
 * on the GPU, writes <bedGraph>_penalty=<penalty_str>_segments.bed and _loss.tsv byte-identical to
 * the reference's.  db_file is created/truncated on the non-trivial branch exactly as the
 * reference does (error 7 if that fails) and receives a small header; the cost functions
 * themselves live in HBM.  Returns a status code. */
int psd_fpop_disk(const char *bedGraph_file_name, const char *penalty_str, const char *db_file_name);

/* The same for n (bedGraph, penalty, db) triples in ONE batched launch: one warp per problem.
 * status_out[i] receives problem i's code.  Returns 0, or PSD_ERR_CUDA if the device failed. */
int psd_fpop_disk_batch(int n, const char *const *bedGraph_file_names, const char *const *penalty_strs,
                        const char *const *db_file_names, int *status_out);

/* Message for a status code; the 1..11 texts are interface.cpp:16-52's format strings. */
const char *psd_status_message(int status);
const char *psd_last_error(void);

/* ---- in-memory batch ("plan"): rows in, segments out, explicit H2D / solve / D2H steps ------ */
typedef struct psd_plan psd_plan;

typedef struct psd_result {
  int32_t status;
  int32_t trivial;           /* 1: one-segment model (penalty Inf or constant coverage), solved on the host */
  int32_t n_rows;            /* bedGraph.lines */
  int32_t n_segments;
  int32_t n_peaks;
  int32_t n_equality;        /* equality.constraints */
  double penalty;
  double bases;
  double mean_pen_cost;
  double total_loss;
  double mean_intervals;
  double max_intervals;
} psd_result;

typedef struct psd_stats {
  double dp_ms;              /* device time of the DP kernel(s), CUDA events on the plan's stream */
  double backtrack_ms;       /* device time of the backtrack kernel(s) */
  double h2d_ms, d2h_ms;
  int64_t rows_solved;       /* rows of the non-trivial problems */
  int64_t store_bytes_algorithmic;  /* sum over problems of N*24 + 20*total_intervals */
  int64_t store_bytes_written;      /* bytes of records + index actually written to HBM */
  int64_t store_bytes_spilled_host; /* record bytes that overflowed into pinned host memory */
  int64_t backtrack_bytes_read;
  int64_t h2d_bytes, d2h_bytes;
  int32_t n_launches;        /* kernels launched by the last solve */
  int32_t n_waves;           /* store waves */
  int32_t n_overflow_tier;   /* problems that needed the global-memory piece-list tier */
  int32_t piece_cap;         /* shared-memory tier capacity (pieces per function) */
  int32_t warps_per_sm;
  int32_t n_sm;
  /* count-vector problems (psd_plan_add_counts): the device run-length encoding done at upload */
  int32_t n_rle_launches;
  double rle_ms;                    /* device time of the four RLE kernels */
  int64_t rle_positions;            /* count positions encoded */
  int64_t rle_bytes_algorithmic;    /* 4 B per position read + 12 B per row written */
  int32_t n_latency_waves;          /* waves of the last solve that ran the latency kernel (one problem per block) */
  int32_t pad_;
  int64_t store_bytes_drained_dma;  /* of the spilled bytes: copied HBM ring -> pinned host by cudaMemcpyAsync on the side stream */
} psd_stats;

/* device < 0 selects the current CUDA device. */
psd_plan *psd_plan_create(int device);
void psd_plan_destroy(psd_plan *plan);
/* Adds one problem; rows are copied.  penalty_is_inf != 0 selects the no-peaks model.
 * Rows must be contiguous (chromStart[i] == chromEnd[i-1], else -6 like the file path,
 * src/PeakSegFPOPLog.cpp:180-186), of positive width and non-negative coverage (else -PSD_ERR_ARG).
 * Returns the problem id (>= 0) or a negative PSD_ERR_*. */
int psd_plan_add(psd_plan *plan, int64_t n_rows, const int32_t *chromStart, const int32_t *chromEnd,
                 const int32_t *coverage, double penalty, int penalty_is_inf);
/* In-memory front end for a count vector (replaces R/PeakSegFPOP_vec.R:18-25 + PeakSegFPOP_df's
 * bedGraph round trip): counts[i] >= 0 is the coverage of base [i, i+1).  The run-length encoding
 * into bedGraph rows happens on the device at upload.  Results are the same as psd_plan_add() on the
 * rows (chromStart 0-based, chromEnd = cumulative run lengths).  Returns the problem id or -PSD_ERR_*. */
int psd_plan_add_counts(psd_plan *plan, int64_t n_positions, const int32_t *counts, double penalty,
                        int penalty_is_inf);
int psd_plan_size(const psd_plan *plan);
/* stream: a cudaStream_t (0 = default stream).  upload/solve/download enqueue work on it;
 * download synchronizes the stream before returning. */
int psd_plan_upload(psd_plan *plan, void *stream);
int psd_plan_solve(psd_plan *plan, void *stream);
int psd_plan_download(psd_plan *plan, void *stream);
/* upload + solve + download */
int psd_plan_run(psd_plan *plan, void *stream);
int psd_plan_result(const psd_plan *plan, int id, psd_result *out);
/* Segments of problem id, last segment first (the order of _segments.bed).  Arrays need
 * n_segments entries; is_peak is 0 (background) / 1 (peak). */
int psd_plan_segments(const psd_plan *plan, int id, int32_t *chromStart, int32_t *chromEnd,
                      int32_t *is_peak, double *mean);
int psd_plan_get_stats(const psd_plan *plan, psd_stats *out);
/* Replaces problem id's penalty (used by the sequential search to re-solve the same rows). */
int psd_plan_set_penalty(psd_plan *plan, int id, double penalty, int penalty_is_inf);

/* Where the wall time of the process's last psd_fpop_disk / psd_fpop_disk_batch call went. */
typedef struct psd_batch_stats {
  double parse_ms;           /* bedGraph text -> rows (distinct files on all host cores) + penalty parsing */
  double build_ms;           /* pass-1 totals, output files created, scratch db files */
  double run_ms;             /* upload + DP + backtrack + download (wall) */
  double write_ms;           /* _segments.bed / _loss.tsv rendered and written */
  double release_ms;
  double dp_ms, backtrack_ms;      /* device time inside run_ms (max over the GPUs used) */
  int64_t rows_parsed, rows_solved;
  int64_t h2d_bytes, d2h_bytes, store_bytes_algorithmic;
  int32_t n_problems, n_files_parsed, n_devices, n_launches, n_waves, n_latency_waves;
} psd_batch_stats;
int psd_last_batch_stats(psd_batch_stats *out);

/* Writes rows as the 4-column bedGraph text R/writeBedGraph.R:35-37 produces (tab separated, no
 * header, one "chrom\tchromStart\tchromEnd\tcount" line per row).  Returns 0 or PSD_ERR_ARG when the
 * file cannot be written. */
int psd_write_bedgraph(const char *path, const char *chrom, int64_t n_rows, const int32_t *chromStart,
                       const int32_t *chromEnd, const int32_t *count);

/* Inspection of the cost-function store (replaces DiskVector::read, src/PeakSegFPOPLog.cpp:103-117,
 * for tests and tools): the stored function `which` (0 = up, 1 = down) of bedGraph row `row` of
 * problem id, in the reference's record fields (src/PeakSegFPOPLog.cpp:18-34: max_log_mean, data_i,
 * prev_log_mean per piece).  *n_pieces receives the piece count; arrays need `cap` >= that many
 * entries (PSD_ERR_ARG otherwise).  Valid for problems solved in the plan's last store wave (all of them
 * unless the store had to be recycled). */
int psd_plan_store_function(psd_plan *plan, int id, int row, int which, int cap, int *n_pieces, double *max_log_mean,
                            int *data_i, double *prev_log_mean);

/* Tunables (call before psd_plan_create): "piece_cap" (shared-memory tier, default 48),
 * "overflow_cap" (global tier, default 8192), "store_gb" (HBM pool, default 0 = auto),
 * "chunk_kb" (store chunk, default 64), "spill_cap" (per-warp global workspace, default 512),
 * "host_spill_gb" (pinned-host overflow of the store: -1 = automatic, 0 = off),
 * "spill_mode" (how spilled records reach the host: 0 = written into an HBM ring that a host thread drains
 * with cudaMemcpyAsync on a side stream (default), 1 = zero-copy stores through the mapping), "ring_gb",
 * "occupancy_mode" (0 = choose per batch, 1 = one block of 14 warps per SM, 2 = two blocks),
 * "latency_mode" (0 = per wave, whichever of the two kernels a makespan model predicts to finish first:
 * the latency kernel -- one problem per block, one chain per warp -- for few or long problems, never beyond
 * "latency_max_blocks" (16) problems per SM; 1 = always the latency kernel; 2 = never),
 * "devices" (GPUs one psd_fpop_disk_batch call may use: 1 = the current device (default), k = the
 * first k, <= 0 = all; problems are dealt longest-first, one plan and host thread per GPU). */
int psd_set_option(const char *name, double value);

int psd_device_count(void);

/* psd_fpop_disk / psd_fpop_disk_batch keep the plan of their last call parked (device buffers, store
 * pool and pinned staging, up to 100 GB of HBM) so that the next call does not allocate again; this
 * frees it.  Safe to call at any time from any thread. */
void psd_release_cache(void);

#ifdef __cplusplus
}
#endif
#endif
