/*
 * rst_align.h — C ABI of the B200-native RGB-D frame alignment engine.
 *
 * This is the drop-in boundary for the reference's `rs_tracker/align` hot path.
 * Citations are relative to the reference tree (yycho0108/RealsenseTracker):
 *
 *   - `bool AlignIcp3d(const Cloud3f& src, const Cloud3f& dst, int max_iter,
 *      Eigen::Isometry3f* transform)`
 *        rs_tracker/align/include/rs_tracker/align/align_icp.hpp:19-24
 *        rs_tracker/align/src/align_icp.cpp:73-167
 *     -> the pose is READ as the initial guess and OVERWRITTEN with the result
 *        (align_icp.cpp:82,156); it maps src -> dst: p_dst ~ T * p_src
 *        (align_icp.cpp:107).  Every rst_align_* call keeps exactly that
 *        contract (`poses_inout`).
 *   - frame-level inputs come from the driver boundary
 *        rs_tracker/driver/include/rs_tracker/driver/rs_driver.hpp:17-22
 *     depth `CV_16UC1` row-major (rs_driver.cpp:211-212), intrinsics as the
 *     3x3 K = [[fx,0,cx],[0,fy,cy],[0,0,1]] (rs_driver.cpp:264-280).
 *   - failure convention: `false` for < 3 usable points (align_icp.cpp:77-79),
 *     non-finite -> failure (align_gicp.cpp:146-151).  Here: per-pair
 *     `rst_stats.status != RST_STATUS_OK`; the C++ wrapper maps it to `bool`.
 *
 * No C++ types, no exceptions, no ownership transfer: the caller allocates all
 * outputs.  One context per GPU; a context is not thread-safe, different
 * contexts are independent (the reference is stateless / re-entrant).
 *
 * There is NO CPU fallback behind this ABI.  Every compute entry point fails
 * with RST_ERR_CUDA / RST_ERR_NO_DEVICE when no sm_100 device is usable.
 */
#ifndef RST_ALIGN_H_
#define RST_ALIGN_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RST_ABI_VERSION 2
#define RST_MAX_LEVELS 4

/* ---- return codes of the API calls (call-level, not per-pair) ---- */
enum {
  RST_OK = 0,
  RST_ERR_INVALID_ARG = 1,  /* null pointer, bad size, bad params          */
  RST_ERR_NO_DEVICE = 2,    /* no CUDA device / device id out of range     */
  RST_ERR_CUDA = 3,         /* a CUDA runtime call failed (see last_error) */
  RST_ERR_CAPACITY = 4,     /* more frames/pairs/pixels than the ctx holds */
  RST_ERR_ALIGNMENT = 5,    /* device input violates the pitch/alignment rule */
  RST_ERR_ARCH = 6          /* device is not sm_100 (B200)                 */
};

/* ---- per-pair status bits (rst_stats.status: last evaluated iteration; rst_stats.any_status: sticky OR) ---- */
enum {
  RST_STATUS_OK = 0,
  RST_STATUS_TOO_FEW = 1,    /* fewer than params.min_count associations
                                (reference: "< 3 points -> false")          */
  RST_STATUS_DEGENERATE = 2, /* 6x6 Cholesky hit a non-positive pivot       */
  RST_STATUS_NON_FINITE = 4  /* NaN/Inf in the normal equations or the pose */
};

/* ---- robust weight kinds (rst_params.robust_kind) ---- */
enum {
  RST_ROBUST_NONE = 0,
  RST_ROBUST_HUBER = 1,          /* w = min(1, delta/|r|)   (ceres::HuberLoss,
                                    align_gicp.cpp:67)                       */
  RST_ROBUST_GEMAN_MCCLURE = 2   /* w = (mu/(r^2+mu))^2    (align_icp.cpp:116-118) */
};

/* ---- block tiling of the fused ICP kernel (rst_params.tiling) ----
 * Block extents depend on the image size and this switch only — never on the batch size — so
 * results are bit-identical across batch sizes, ranks and entry points for a given tiling; the two
 * tilings differ from each other in the last bits (different fp32 partial-sum extents). */
enum {
  RST_TILING_THROUGHPUT = 0,  /* up to 8192 pixels per block: fewest partials, best for batches   */
  RST_TILING_LATENCY = 1      /* about 148 blocks per pair and level: a single pair fills the GPU in one wave */
};

/* ---- launch schedule of the iteration loop (rst_set_schedule) ---- */
enum {
  RST_SCHEDULE_AUTO = 0,           /* default: the schedule measured fastest on B200 (currently PER_ITERATION at every
                                      batch size; never chosen by the batch size, so results stay bit-identical across
                                      batch sizes)                                                                    */
  RST_SCHEDULE_FUSED = 1,          /* one thread-block cluster per pair runs every iteration of every level in ONE
                                      launch (partial sums combined through distributed shared memory)               */
  RST_SCHEDULE_PER_ITERATION = 2,  /* one launch per iteration per level (block partials in global memory)           */
  RST_SCHEDULE_HYBRID = 3          /* coarse levels fused (one launch), finest level one launch per iteration         */
};

/* Pin-hole intrinsics of the finest level; K = [[fx,0,cx],[0,fy,cy],[0,0,1]]
 * (rs_driver.cpp:264-280).  No distortion. */
typedef struct rst_intrinsics {
  float fx, fy, cx, cy;
} rst_intrinsics;

/* One depth frame (host or device memory depending on the call).
 * depth: uint16 row-major, 0 = invalid (CV_16UC1, rs_driver.cpp:211).
 * rgb:   optional uint8 RGB, 3 bytes per pixel (CV_8UC3, rs_driver.cpp:212);
 *        NULL when unused. Strides are in BYTES. */
typedef struct rst_frame {
  const uint16_t* depth;
  const uint8_t* rgb;
  int32_t width, height;
  int32_t depth_stride_bytes;
  int32_t rgb_stride_bytes;
} rst_frame;

/* Algorithm parameters (the reference hard-codes its constants,
 * align_icp.cpp:91,96-98,165; here they are one POD). Fill with
 * rst_params_default() and override. */
typedef struct rst_params {
  int32_t num_levels;             /* pyramid levels, 1..RST_MAX_LEVELS          */
  int32_t iters[RST_MAX_LEVELS];  /* iterations per level, [0] = finest; run
                                     coarse -> fine; fixed count, no early exit */
  float depth_scale;              /* metres per depth LSB (D435: 0.001)         */
  float z_min, z_max;             /* valid depth range in metres                */
  float dist_max;                 /* association gate |p'-q| <= dist_max (m)    */
  float normal_cos_min;           /* src/dst normal gate; <= -1 disables        */
  float normal_depth_tol;         /* tau_n: |z_nbr - z| <= tau_n * z            */
  int32_t pyr_depth_tol;          /* tau_pyr in LSB for the 2x2 depth pooling   */
  int32_t robust_kind;            /* RST_ROBUST_*                               */
  float robust_scale;             /* Huber delta (m) or Geman-McClure mu (m^2)  */
  int32_t min_count;              /* minimum associations per iteration         */
  float damping;                  /* added to diag(A) before the solve          */
  float photo_weight;             /* lambda of the photometric term (needs rst_frame.rgb), 0 = off:
                                     cost = sum r_geo^2 + lambda * sum (I_dst(pi(p')) - I_src)^2  */
  int32_t tiling;                 /* RST_TILING_*: how pixels are cut into blocks */
  float converge_eps;             /* > 0: a pair skips the remaining iterations of a level once an update
                                     has |omega| < eps (rad) and |v| < eps (m); 0 (default) = fixed count,
                                     like the reference's AlignIcp3d (no convergence test, align_icp.cpp:92) */
  int32_t reserved[2];
} rst_params;

/* Per-pair result statistics. A/b/sum_wr2/count belong to the LAST evaluated
 * iterate, i.e. the correspondences BEFORE the final update — the same
 * "pre-update" semantics as the reference's mean_cost (align_icp.cpp:104-113,157). */
typedef struct rst_stats {
  int32_t status;      /* RST_STATUS_* of the LAST evaluated iteration (plus NON_FINITE if the pose is not
                          finite): the success criterion — a failed solve on a coarse level of sparse depth
                          does not fail a pair whose fine levels converged                               */
  int32_t iterations;  /* iterations executed (failed solves included)                                  */
  int32_t count;       /* associations in the last iteration                                            */
  float rmse;          /* sqrt(sum_wr2 / count) of the last iteration (m)                               */
  int32_t any_status;  /* RST_STATUS_* bit-or over ALL iterations (sticky)                              */
  int32_t failed_iterations; /* iterations whose solve failed (pose left unchanged by them)             */
  double sum_wr2;      /* sum w r^2                                             */
  double A[21];        /* upper triangle of J^T W J, row-major (0,0)..(5,5)     */
  double b[6];         /* J^T W r                                               */
} rst_stats;

typedef struct rst_ctx rst_ctx;

/* -------------------------------------------------------------------------
 * Context
 * ---------------------------------------------------------------------- */

/* Creates a context on CUDA device `device` able to hold `max_frames` frames of
 * up to max_width x max_height and `max_pairs` pairs. `stream` is a
 * cudaStream_t (or NULL to let the context create its own). */
int32_t rst_ctx_create(int32_t device, int32_t max_width, int32_t max_height,
                       int32_t max_frames, int32_t max_pairs, void* stream,
                       rst_ctx** out_ctx);
void rst_ctx_destroy(rst_ctx* ctx);

/* Last error message of this context ("" if none). Never NULL. */
const char* rst_last_error(const rst_ctx* ctx);
/* Message of the last rst_ctx_create failure in this thread. */
const char* rst_last_create_error(void);

int32_t rst_abi_version(void);
void rst_params_default(rst_params* params);

/* -------------------------------------------------------------------------
 * Alignment — host frames in, poses out (H2D and D2H inside the call).
 * Replaces: AlignIcp3d(src, dst, max_iter, &T)  align_icp.cpp:163-167
 *   pair i aligns src[i] onto dst[i]; poses_inout[16*i..] is a column-major
 *   4x4 fp32 matrix (bit-compatible with Eigen::Isometry3f::matrix()).
 * stats_out may be NULL.  Returns RST_OK when the batch ran; per-pair failure
 * is in stats_out[i].status.
 * ---------------------------------------------------------------------- */
int32_t rst_align_pairs(rst_ctx* ctx, const rst_frame* src, const rst_frame* dst,
                        int32_t n_pairs, const rst_intrinsics* intr,
                        const rst_params* params, float* poses_inout,
                        rst_stats* stats_out);

/* Frame-to-frame odometry over a sequence (the loop of rs_replay_app.cpp:211-287):
 * pair i aligns frames[i+1] (src, "curr") onto frames[i] (dst, "prev"), so
 * poses_inout[16*i..] = T_{i <- i+1}; n_frames-1 pairs. Each frame is uploaded
 * and pre-processed once. */
int32_t rst_align_sequence(rst_ctx* ctx, const rst_frame* frames, int32_t n_frames,
                           const rst_intrinsics* intr, const rst_params* params,
                           float* poses_inout, rst_stats* stats_out);

/* Asynchronous forms: enqueue H2D, kernels and the D2H of the results into the context's pinned
 * staging, and return without waiting; rst_wait() blocks until that work has finished and copies
 * the poses (n x 16) / statistics (either may be NULL) out. `poses_in` may be NULL (identity
 * priors). The host frames must stay valid until rst_wait returns. Two contexts driven
 * alternately (submit k+1, wait k) overlap the PCIe copies of one batch with the kernels of the
 * other — the streaming form of the odometry loop. One outstanding call per context. */
int32_t rst_align_pairs_async(rst_ctx* ctx, const rst_frame* src, const rst_frame* dst,
                              int32_t n_pairs, const rst_intrinsics* intr,
                              const rst_params* params, const float* poses_in);
int32_t rst_align_sequence_async(rst_ctx* ctx, const rst_frame* frames, int32_t n_frames,
                                 const rst_intrinsics* intr, const rst_params* params,
                                 const float* poses_in);
int32_t rst_wait(rst_ctx* ctx, float* poses_out, rst_stats* stats_out);

/* The host-frame calls can upload and process their frames in chunks so that the H2D copy of chunk k+1 (on an
 * internal copy stream) overlaps the kernels of chunk k. `frames_per_chunk` > 0: that many frames (pairs) per chunk;
 * 0 (default): automatic — blocking calls on at least 64 frames / 32 pairs run as two halves, asynchronous calls are
 * not chunked (their caller overlaps whole calls on two contexts); < 0: never.
 * The chunk size never changes results: every pair is reduced in image-size-determined blocks. */
int32_t rst_set_pipeline_chunk(rst_ctx* ctx, int32_t frames_per_chunk);

/* Batches of at least `min_pairs` pairs (default 32) run their iteration schedule as two halves on two
 * internal streams, so the launches of one half fill the partial last wave and the latency-bound coarse
 * levels of the other. min_pairs <= 0 disables the split (single stream; used when timing one kernel in
 * isolation). Never changes results. */
int32_t rst_set_stream_split(rst_ctx* ctx, int32_t min_pairs);

/* Blocking host-frame calls (rst_align_pairs / rst_align_sequence) with at most `max_pairs` pairs (default 8) run
 * their kernels as ONE replayed CUDA graph (captured per frame size / parameter set / pair count): the single-pair
 * latency path. 0 disables it. Never changes results. */
int32_t rst_set_graph_max_pairs(rst_ctx* ctx, int32_t max_pairs);

/* Selects how the iteration loop is launched (RST_SCHEDULE_*). Poses agree between the two schedules to fp32
 * round-off of the partial-sum extents; each schedule is bit-reproducible. */
int32_t rst_set_schedule(rst_ctx* ctx, int32_t schedule);

/* CTAs per pair of the fused kernel for a tiling (RST_TILING_*; defaults 4 and 8, at most 16). The cluster size
 * fixes the partial-sum extents, so it is a property of the context, never of the batch. */
int32_t rst_set_cluster_size(rst_ctx* ctx, int32_t tiling, int32_t ctas_per_pair);
/* How many clusters of that size the device holds at once (pairs iterating concurrently); 0 on error. */
int32_t rst_max_active_clusters(rst_ctx* ctx, int32_t ctas_per_pair);

/* -------------------------------------------------------------------------
 * Staged / device-resident interface (what the two calls above are built from;
 * this is what an HBM-resident benchmark or a multi-stage host uses).
 * ---------------------------------------------------------------------- */

/* Declares the geometry of the frames that follow and resets the frame slots. */
int32_t rst_begin(rst_ctx* ctx, int32_t width, int32_t height,
                  const rst_intrinsics* intr, const rst_params* params);

/* Copies host depth frames into slots [first_slot, first_slot+n). Asynchronous
 * on the context stream when the host memory is pinned. */
int32_t rst_upload_frames(rst_ctx* ctx, const rst_frame* frames, int32_t n,
                          int32_t first_slot);

/* Uses `n` depth frames already in device memory, laid out back to back:
 * frame k starts at d_depth + k*frame_stride_px, rows are row_stride_px apart.
 * Requires d_depth 16-byte aligned and row_stride_px, frame_stride_px multiples
 * of 8 (RST_ERR_ALIGNMENT otherwise). The memory is read in place (level 0 is
 * not copied) and must stay valid until the next rst_begin. */
int32_t rst_set_frames_device(rst_ctx* ctx, const uint16_t* d_depth, int32_t n,
                              int32_t row_stride_px, int64_t frame_stride_px,
                              int32_t first_slot);

/* Builds depth pyramid + geometry maps (K1+K2+K6) for slots [first_slot, first_slot+n). */
int32_t rst_preprocess(rst_ctx* ctx, int32_t first_slot, int32_t n);

/* Runs the coarse-to-fine ICP for n_pairs pairs; pair i = (src_slots[i] ->
 * dst_slots[i]) (host int arrays). poses_inout / stats_out are HOST pointers,
 * or NULL to leave the results on the device (see rst_device_results). The call
 * is asynchronous w.r.t. the host unless results are requested. */
int32_t rst_align_slots(rst_ctx* ctx, const int32_t* src_slots,
                        const int32_t* dst_slots, int32_t n_pairs,
                        float* poses_inout, rst_stats* stats_out);

/* Device pointers of the result arrays of the last rst_align_slots:
 * poses: n_pairs x 16 fp32 (column-major 4x4), stats: n_pairs x rst_stats. */
int32_t rst_device_results(rst_ctx* ctx, const float** d_poses,
                           const rst_stats** d_stats);

/* Blocks until all work queued on the context stream has finished. */
int32_t rst_sync(rst_ctx* ctx);

/* -------------------------------------------------------------------------
 * Single-stage entry points (parity tests and staged hosts).  All operate on
 * the slots of the current rst_begin(); outputs are HOST pointers.
 * ---------------------------------------------------------------------- */

/* Level geometry: width/height/pitch (pixels) and intrinsics of pyramid level. */
int32_t rst_level_info(const rst_ctx* ctx, int32_t level, int32_t* width,
                       int32_t* height, int32_t* pitch_px, rst_intrinsics* intr);

/* Reads back the depth pyramid level of a slot, densely packed width*height. */
int32_t rst_read_depth(rst_ctx* ctx, int32_t slot, int32_t level, uint16_t* out);

/* Reads back the geometry map of a slot/level: width*height float4
 * {nx, ny, nz, z}; z = 0 where the vertex or the normal is invalid. */
int32_t rst_read_geometry(rst_ctx* ctx, int32_t slot, int32_t level, float* out);

/* Reads back the intensity map of a slot/level (photometric term on): width*height floats in [0,1]. */
int32_t rst_read_intensity(rst_ctx* ctx, int32_t slot, int32_t level, float* out);

/* One association + normal-equation evaluation at `level` under `pose`
 * (column-major 4x4 fp32, src->dst), without updating any pose:
 * idx_out[width*height] = dst pixel index v'*width+u' or -1 (may be NULL);
 * stats_out gets A, b, sum_wr2, count of that evaluation. */
int32_t rst_evaluate(rst_ctx* ctx, int32_t src_slot, int32_t dst_slot,
                     int32_t level, const float* pose, int32_t* idx_out,
                     rst_stats* stats_out);

/* -------------------------------------------------------------------------
 * Cloud-based alignment with the reference's own algorithm, on the GPU.
 * Literal counterpart of
 *   bool AlignIcp3d(const Cloud3f& src, const Cloud3f& dst, const int max_iter,
 *                   Eigen::Isometry3f* const transform)   align_icp.cpp:73-167
 * exact 1-NN of every transformed source point in dst (kdtree.hpp:51-57, here a uniform
 * grid with ring expansion — exact, ties to the lowest index), Geman-McClure weights with mu
 * annealed /1.4 every 8 iterations (:91,96-98,116-118), UNWEIGHTED dst/src centroids
 * (:85-86,120-122), weighted fp32-product / fp64-accumulated cross-covariance (:125-136),
 * 3x3 SVD -> R = U V^T with the literal reflection patch (:139-145), t = mu_d - R mu_s (:148),
 * Translation * Quaternion round trip (:151), fixed max_iter iterations, returns
 * mean_cost = sqrt(cost/N) of the last iteration's pre-update correspondences (:157-160).
 * Clouds are xyz-interleaved fp32 (the memory of Cloud3f = Eigen 3xN column-major), HOST pointers.
 *
 * What "exact" covers: for the same input pose the neighbour indices and the weights are bit-identical to the
 * reference algorithm (transform and squared distance in the reference's operation order, no contraction). Equal
 * distances resolve to the LOWEST index; nanoflann resolves them in tree-traversal order, so on clouds with exactly
 * tied neighbours the index may differ (the distance never does). Centroids are accumulated in fp64 (the reference:
 * sequential fp32) and the pose composition / determinant / eigen-solvers are ordinary device fp32 (FMA contraction
 * allowed): covariances agree to 1e-4 relative and poses to 1e-4 m / 1e-4 rad, not bit for bit.
 * ---------------------------------------------------------------------- */
typedef struct rst_cloud {
  const float* xyz;  /* n x 3 */
  int32_t n;
} rst_cloud;

typedef struct rst_icp3d_result {
  int32_t ok;        /* the reference's return value: n >= 3 && m >= 3 && mean_cost < 10000 */
  int32_t iterations;
  float mean_cost;   /* sqrt(sum d^2 / N), last iteration, pre-update                       */
  float mu;          /* final Geman-McClure mu                                              */
  double cov[9];     /* last iteration's cross-covariance, row-major                        */
} rst_icp3d_result;

/* pair i aligns src[i] onto dst[i]; poses_inout: n_pairs x 16 column-major fp32, initial guess in,
 * result out. results / nbrs_out / weights_out may be NULL; nbrs_out and weights_out receive the
 * last iteration's correspondences of all pairs back to back (sum of src[i].n entries).
 * grid_cell <= 0 picks the search-grid cell size automatically. */
int32_t rst_icp3d_pairs(rst_ctx* ctx, const rst_cloud* src, const rst_cloud* dst, int32_t n_pairs,
                        int32_t max_iter, float grid_cell, float* poses_inout,
                        rst_icp3d_result* results, int32_t* nbrs_out, float* weights_out);

/* CTAs per pair of the cloud ICP kernel (rst_icp3d_pairs / rst_icp3d_depth): 0 (default) = automatic — one CTA per
 * pair for batches that fill the GPU, a thread-block cluster of up to 16 CTAs per pair for small batches (a single
 * pair, the reference caller's case, then runs on 16 SMs instead of one); 1, 2, 4, 8, 16 force a size. Results of
 * different sizes agree to fp64 round-off of the partial-sum order (neighbour indices: bit-identical for the same
 * pose). */
int32_t rst_set_icp3d_cluster(rst_ctx* ctx, int32_t ctas_per_pair);

/* Neighbour cache of the cloud ICP kernel. The reference queries its k-d tree for every source point in every one of the
 * 128 iterations (align_icp.cpp:105-113). Here a query that has to search scans `margin` further than the candidate
 * neighbour requires and records how far every OTHER target point is proven to be; in later iterations the triangle
 * inequality |p' - nbr| + |p' - p| < L shows, for most points, that the neighbour cannot have changed, and the search is
 * skipped. The neighbour index and its fp32 squared distance are the ones a search would return, bit for bit — only
 * the time changes. margin = clamp(gain * motion of the point since its last search, lo_cells * cell, hi_cells * cell).
 * hi_cells = 0 switches the cache off (every point searches in every iteration). Defaults: 1, 0.05, 0.2. */
int32_t rst_set_icp3d_cache(rst_ctx* ctx, float gain, float lo_cells, float hi_cells);

/* Of the last rst_icp3d_pairs / rst_icp3d_depth call of this context: how many neighbour queries were answered
 * (source points x iterations, over all pairs) and how many of them had to search (the rest were proven by the cache). */
int32_t rst_icp3d_cache_stats(rst_ctx* ctx, uint64_t* searched_out, uint64_t* queried_out);

/* Fixed-point skip of the cloud ICP kernel (default on). The reference runs max_iter iterations whatever happens
 * (align_icp.cpp:92-153). Once an iteration returns the pose it was given, bit for bit, the iterations after it repeat it
 * exactly until mu changes (every 8th iteration, :96-98): same pose in, same neighbours, same sums, same pose out. The
 * kernel detects that state (pose unchanged AND its SVD warm-start basis unchanged, i.e. the complete state an
 * iteration depends on) and continues at the next change of mu; the last iteration is always run. Results are those of
 * running every iteration, bit for bit (test); rst_icp3d_result.iterations still reports max_iter. on = 0 runs them all.
 * rst_icp3d_iteration_stats: iterations actually run / asked for, summed over the pairs of the last call. */
int32_t rst_set_icp3d_fixed_point_skip(rst_ctx* ctx, int32_t on);
int32_t rst_icp3d_iteration_stats(rst_ctx* ctx, uint64_t* run_out, uint64_t* asked_out);

/* bool SolveKabsch(src, dst, indices, weights, &xfm)  (align_icp.hpp:14-18, align_icp.cpp:18-71) on the device:
 * closed-form pose from GIVEN (src index, dst index) pairs — the initialiser rs_align_app.cpp:295 feeds to
 * AlignIcp3d. `pairs`: n_pairs x 2 int32; `weights`: n_pairs floats or NULL (the reference's empty vector);
 * unweighted centroids, weighted covariance, exactly as written there. pose_out: 16 floats column-major;
 * *ok_out = 0 (and RST_OK) when either cloud has fewer than 3 points (:23-25). All pointers are HOST memory. */
int32_t rst_solve_kabsch(rst_ctx* ctx, const rst_cloud* src, const rst_cloud* dst, const int32_t* pairs,
                         int32_t n_pairs, const float* weights, float* pose_out, int32_t* ok_out);

/* ComputeNormals(cloud, tree, k, &normals) + OrientNormals(cloud, viewpoint, &normals)
 * (point_cloud_utils.cpp:176-216) on the device: exact k nearest neighbours (k in [2,32], counting the
 * point itself; rs_align_app.cpp:25 uses 16), fp32 centroid + covariance, eigenvector of the smallest
 * eigenvalue, flipped so that n . (p - viewpoint) <= 0. normals_out: n x 3 floats. HOST pointers. */
int32_t rst_cloud_normals(rst_ctx* ctx, const rst_cloud* cloud, int32_t k, const float* viewpoint,
                          float grid_cell, float* normals_out);

/* The cloud utilities of rs_tracker/common/point_cloud_utils.hpp:11-31 for callers that hold clouds (every caller of
 * the reference does: rs_replay_app.cpp:229,246-247; rs_tracker.cpp:60-87). All pointers are HOST memory; every call
 * builds the search grid of its cloud on the device (the KDTree3f argument of the reference has no counterpart) and
 * then runs as many blocks as the cloud needs. grid_cell <= 0 picks the cell size automatically.
 *
 * FindCorrespondences(tree, source, &indices, &squared_distances)  point_cloud_utils.cpp:70-90: exact 1-NN of every
 *   source point in `target`; ties go to the lowest index; a non-finite source point gets index -1, distance +inf. */
int32_t rst_find_correspondences(rst_ctx* ctx, const rst_cloud* target, const rst_cloud* source, float grid_cell,
                                 int32_t* indices_out, float* sq_dist_out);
/* KDTree3f (kdtree.hpp:11-99; types.hpp) as a device-resident handle: the search structure of one cloud, built once
 * (rst_tree_create: upload + one grid-build launch) and queried any number of times without re-uploading or re-building
 * it — what the reference does with `dst_tree` (align_icp.cpp:163-167) and `tree` (rs_replay_app.cpp:382).
 * rst_tree_query = KDTree3f::query(point, num_closest, out_indices, out_distances_sq) (kdtree.hpp:51-57) for a batch:
 * row i of the n_queries x k outputs holds the k nearest points of query i in ascending distance (ties: lower index
 * first; k = 1 is exactly rst_find_correspondences), squared L2 distances accumulated left to right as nanoflann's
 * L2_Adaptor. Entries past the cloud's size and all entries of a non-finite query are -1 / +inf. k in [1, 33].
 * The tree belongs to the device of the context that made it; any context on that device may query it; the cloud is
 * copied (the caller's buffer is not referenced after rst_tree_create). HOST pointers in and out. */
typedef struct rst_tree rst_tree;
int32_t rst_tree_create(rst_ctx* ctx, const rst_cloud* cloud, float grid_cell, rst_tree** tree_out);
int32_t rst_tree_query(rst_ctx* ctx, const rst_tree* tree, const float* queries_xyz, int32_t n_queries, int32_t k,
                       int32_t* indices_out, float* sq_dist_out);
int32_t rst_tree_size(const rst_tree* tree);
void rst_tree_destroy(rst_tree* tree);

/* ComputeCovariances(tree, cloud, &covs, use_gicp)  point_cloud_utils.cpp:100-161: 32 nearest OTHER points, fp32
 *   centroid + scatter; use_gicp = 0: / 31; use_gicp != 0: singular values replaced by (1, 1, 1e-2) (:139-154).
 *   covs_out: n x 9 floats (row-major symmetric 3x3). */
int32_t rst_cloud_covariances(rst_ctx* ctx, const rst_cloud* cloud, int32_t use_gicp, float grid_cell, float* covs_out);
/* DownsampleVoxel(cloud_in, voxel_size, &cloud_out)  point_cloud_utils.cpp:34-68: key floor(p / voxel_size), the first
 *   point of a voxel wins; output in first-occurrence order (the reference's order is unordered_map iteration order).
 *   xyz_out must hold cloud_in->n points; *n_out = points kept. */
int32_t rst_downsample_voxel(rst_ctx* ctx, const rst_cloud* cloud_in, float voxel_size, float* xyz_out, int32_t* n_out);
/* RemoveNans(cloud_in, &cloud_out)  point_cloud_utils.cpp:163-174: keeps the points whose three coordinates are finite,
 *   in order. xyz_out must hold cloud_in->n points. */
int32_t rst_remove_nans(rst_ctx* ctx, const rst_cloud* cloud_in, float* xyz_out, int32_t* n_out);
/* ComputeCentroid(cloud, &centroid)  point_cloud_utils.cpp:92-98: mean of the points, summed in fp64 in a fixed order
 *   (the reference sums sequentially in fp32: equal to fp32 round-off of that sum). centroid_out: 3 floats. An empty
 *   cloud is RST_ERR_INVALID_ARG (the reference divides by zero). */
int32_t rst_cloud_centroid(rst_ctx* ctx, const rst_cloud* cloud, float* centroid_out);
/* ComputeExtents(cloud, &box)  point_cloud_utils.cpp:26-32: axis-aligned bounding box, lo_out / hi_out: 3 floats each
 *   (bit-exact: min and max are exact). An empty cloud gives Eigen's empty box (lo = FLT_MAX, hi = -FLT_MAX); NaN
 *   coordinates are ignored, infinite ones are not. */
int32_t rst_cloud_extents(rst_ctx* ctx, const rst_cloud* cloud, float* lo_out, float* hi_out);
/* OrientNormals(cloud, viewpoint, &normals)  point_cloud_utils.cpp:205-216: normals_inout[i] is negated where
 *   (cloud[i] - viewpoint) . normals_inout[i] > 0. rst_cloud_normals() already orients its output; this is for normals
 *   that come from elsewhere or for a second viewpoint. viewpoint: 3 floats; normals_inout: n x 3 floats. */
int32_t rst_orient_normals(rst_ctx* ctx, const rst_cloud* cloud, const float* viewpoint, float* normals_inout);

/* GICP plane-to-plane (rs_tracker/align: gicp_cost.hpp:40-73, align_gicp.cpp:41-163) on the device.
 * Residual of correspondence i -> j = dst_indices[i]:  e = C^{-1/2} (R s_i + t - d_j),  C = C_d[j] + R C_s[i] R^T,
 * robustified with ceres::HuberLoss(huber_delta) (the reference uses 0.5, align_gicp.cpp:67; <= 0 = no loss).
 * cost = 1/2 sum rho(|e|^2) (ceres' final_cost, :113); A, b = Gauss-Newton normal equations sum w J^T J, sum w J^T e
 * with J = C^{-1/2} [ -[p']x | I ], omega first (same convention as rst_stats), C held at the current rotation. */
typedef struct rst_gicp_stats {
  double cost;
  double A[21];
  double b[6];
  int32_t count;     /* correspondences evaluated (dst index in range, finite residual) */
  int32_t reserved;
} rst_gicp_stats;

/* The cost of the 7-argument ComputeAlignment (align_gicp.cpp:41-117) at `pose` (16 floats, column-major, src->dst) for
 * GIVEN covariances (n x 9 / m x 9 floats, row-major) and correspondences (n entries, < 0 = none).
 * residuals_out: n x 3 (nullable; rows without a correspondence are left untouched). HOST pointers. */
int32_t rst_gicp_evaluate(rst_ctx* ctx, const rst_cloud* src, const rst_cloud* dst, const float* src_covs, const float* dst_covs,
                          const int32_t* dst_indices, const float* pose, float huber_delta, float* residuals_out,
                          rst_gicp_stats* stats_out);

/* The 7-argument ComputeAlignment (align_gicp.cpp:41-117) itself: the pose that minimises the robustified cost above
 * for GIVEN covariances and correspondences (layouts as rst_gicp_evaluate), starting from the seed in pose_inout. The
 * reference hands the problem to Ceres (Levenberg-Marquardt, at most 1024 iterations; absent, out of scope): here
 * `max_iters` Levenberg-Marquardt steps on the normal equations above, a step that raises the cost being rejected
 * (0 = evaluate the seed only). stats_out (nullable): cost (= ceres' final_cost, the function's return value), A, b and
 * count at the returned pose. */
int32_t rst_gicp_minimize(rst_ctx* ctx, const rst_cloud* src, const rst_cloud* dst, const float* src_covs, const float* dst_covs,
                          const int32_t* dst_indices, int32_t max_iters, float huber_delta, float* pose_inout,
                          rst_gicp_stats* stats_out);

/* The 3-argument ComputeAlignment (align_gicp.cpp:119-163): covariances of both clouds (ComputeCovariances;
 * use_gicp_covariances = 0 is what the reference passes), then `max_outer` (reference: 16) rounds of
 * { FindCorrespondences(dst, pose * src); minimise the robustified cost over the fixed correspondences }. The reference
 * minimises with Ceres (absent, out of scope): here `inner_iters` Levenberg-Marquardt steps on the normal equations
 * above. pose_inout: initial guess in (the reference always starts from the identity, :147), result out. */
int32_t rst_gicp_align(rst_ctx* ctx, const rst_cloud* src, const rst_cloud* dst, int32_t max_outer, int32_t inner_iters,
                       float huber_delta, int32_t use_gicp_covariances, float grid_cell, float* pose_inout,
                       rst_gicp_stats* stats_out);

/* The reference caller's whole per-pair sequence on the device, from depth frames
 * (rs_replay_app.cpp:229,246-251): back-projection with invalid pixels at the origin
 * (rs_driver.cpp:83-88,201-202) -> RemoveNans -> DownsampleVoxel(voxel) (first point per voxel, in
 * first-occurrence order; voxel <= 0 skips the decimation) -> AlignIcp3d(max_iter).
 * `frames`: the n_frames unique HOST frames; pair i aligns frames[src_idx[i]] onto
 * frames[dst_idx[i]] (a sequence is src_idx = 1..n-1, dst_idx = 0..n-2, every frame converted once).
 * counts_out (nullable, n_frames): points in every frame's cloud. */
int32_t rst_icp3d_depth(rst_ctx* ctx, const rst_frame* frames, int32_t n_frames, const int32_t* src_idx,
                        const int32_t* dst_idx, int32_t n_pairs, const rst_intrinsics* intr,
                        float depth_scale, float voxel, int32_t max_iter, float grid_cell,
                        float* poses_inout, rst_icp3d_result* results, int32_t* counts_out);

/* Reads back (host xyz_out, n_points x 3) the cloud of one frame of the last rst_icp3d_depth call. */
int32_t rst_icp3d_read_cloud(rst_ctx* ctx, int32_t frame_index, float* xyz_out, int32_t n_points);

/* Number of kernel launches this context has issued so far (bench evidence). */
int64_t rst_launch_count(const rst_ctx* ctx);

/* Copies the results of the last rst_align_slots into caller DEVICE buffers
 * (n_pairs x 16 fp32, n_pairs x rst_stats; either may be NULL), asynchronously on
 * the context stream — e.g. into the send buffer of an NCCL all-gather. */
int32_t rst_copy_results_device(rst_ctx* ctx, float* d_poses_out, rst_stats* d_stats_out);

/* Per-stage device timing with CUDA events on the context stream. Events bracket each
 * GROUP of launches (one pre-processing launch per level; all iterations of one pyramid
 * level), so enabling it does not perturb the launch sequence. */
typedef struct rst_profile {
  float ms_preprocess[RST_MAX_LEVELS];       /* K1+K2+K6, per level, summed since enable */
  float ms_icp[RST_MAX_LEVELS];              /* K3+K4+K5, per level, summed since enable */
  int32_t launches_preprocess[RST_MAX_LEVELS];
  int32_t launches_icp[RST_MAX_LEVELS];
  int64_t frames_preprocessed[RST_MAX_LEVELS]; /* frames covered by those launches      */
  int64_t pairs_iterated[RST_MAX_LEVELS];      /* sum over launches of pairs per launch */
  float ms_icp_fused;                          /* fused schedule: whole iteration loop, summed since enable */
  int32_t launches_icp_fused;
  int64_t pair_iterations_fused;               /* sum over launches of pairs x iterations of all levels */
} rst_profile;
int32_t rst_profile_enable(rst_ctx* ctx, int32_t on); /* also resets the accumulators */
int32_t rst_profile_read(rst_ctx* ctx, rst_profile* out); /* synchronises the stream  */

#ifdef __cplusplus
}
#endif

#endif /* RST_ALIGN_H_ */
