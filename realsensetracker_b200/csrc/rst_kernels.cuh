/*
 * rst_kernels.cuh — argument blocks and launchers of the sm_100a alignment kernels.
 *
 *   K1+K2+K6  k_preprocess : uint16 depth tile (+1 px halo) -> shared memory by ONE TMA tensor
 *                            copy (cp.async.bulk.tensor.3d, out-of-image pixels zero-filled by
 *                            the hardware); back-projection, normals -> geometry map float4
 *                            {nx,ny,nz,z}; 2x2 integer pooling -> next pyramid level.
 *   K3+K4+K5  k_icp_iter   : one launch per iteration (default): depth tile by bulk async copy,
 *                            projective association + point-to-plane residual/Jacobian, 29
 *                            sums per thread -> warp shuffle tree -> block tree -> block
 *                            partials in global memory -> last block of a pair reduces in
 *                            fp64, 6x6 Cholesky + SE(3) update; also rst_evaluate.
 *             k_icp_fused  : the same pixel pipeline, one thread-block cluster per pair, every
 *                            iteration of the chosen levels in ONE launch: cluster totals
 *                            through distributed shared memory in fixed rank order (fp64),
 *                            solve on the leader CTA, next pose back through DSMEM
 *                            (selectable schedule).
 *   k_init_pairs           : pose upload -> fp64 master pose, state reset.
 *
 * Nearest reference counterparts: align_icp.cpp:101-151 (correspondence loop,
 * covariance accumulation, closed-form solve), rs_driver.cpp:201-202 (back-projection),
 * point_cloud_utils.cpp:176-216 (normals + orientation).
 */
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "rst_align.h"

namespace rst {

constexpr int kAcc = 29;          // 21 (A upper) + 6 (b) + sum w r^2 + count
constexpr int kAccPad = 32;       // partial row stride (floats)
#ifndef RST_ICP_THREADS
#define RST_ICP_THREADS 128
#endif
#ifndef RST_ICP_CPW
#define RST_ICP_CPW 2
#endif
#ifndef RST_ICP_MINB
#define RST_ICP_MINB 4            // resident blocks per SM k_icp_iter is compiled for: 4 -> 128 registers, no spills (5 -> 96: 1.5 % slower)
#endif
#ifndef RST_FUSED_MINB
#define RST_FUSED_MINB 4          // resident CTAs per SM the fused kernel is compiled for (4 -> 128 registers per thread)
#endif
#ifndef RST_ICP_GROUP_PX
#define RST_ICP_GROUP_PX 8192     // pixels per block on large levels (groups = this / block pixels per group)
#endif
constexpr int kIcpThreads = RST_ICP_THREADS;
constexpr int kChunkPx = 64;      // pixels one warp covers per step (2 per lane)
constexpr int kChunksPerWarp = RST_ICP_CPW; // chunks per warp per group -> 2*CPW pixels in flight per thread
constexpr int kMaxGroups = RST_ICP_GROUP_PX / ((RST_ICP_THREADS / 32) * RST_ICP_CPW * 64);  // groups per block on large levels
constexpr int kPxPerStage = 2 * kChunksPerWarp;  // gathers per thread per pipeline stage
constexpr float kMinNormalLen2 = 1e-30f; // |a x b|^2 below this is a degenerate normal
constexpr float kMinProjZ = 1e-6f; // transformed points closer than this to the camera plane are rejected
constexpr int kChunksPerBlock = (kIcpThreads / 32) * kChunksPerWarp;  // 32 -> 2048 px
constexpr int kTileW = 64, kTileH = 32;  // preprocess tile

struct LevelGeom {
  int32_t w, h;
  float fx, fy, cx, cy, ifx, ify;
};

/* one pyramid level of the frame store */
struct LevelStore {
  const uint16_t* depth;     // slot 0, row 0
  int32_t depth_pitch;       // pixels between rows
  int64_t depth_frame;       // pixels between frames
  float4* geom;              // dense w*h per frame; texel [-1] of every frame is an all-zero guard
  int64_t geom_frame;        // float4 between frames
  const float* intensity;    // dense w*h per frame (photometric term only), else nullptr
  int64_t int_frame;
};

/* a CUtensorMap (128 bytes, 64-byte aligned) without pulling cuda.h into every translation unit */
struct alignas(64) TensorMap {
  unsigned long long opaque[16];
};
constexpr int kPreBoxW = 80, kPreBoxH = kTileH + 2;   // k_preprocess depth box: 7 pad | 1 halo | 64 | 1 halo | 7 pad columns (the box must
                                                      // start on a 16-byte boundary of the row: x0 - 8), 1 + 32 + 1 rows

struct PreArgs {
  LevelGeom g;
  LevelStore cur;
  int32_t tmap_slot0;        // slot of the tensor map's frame 0 (0 for the context's own store)
  uint16_t* next_depth;      // next level (nullable)
  int32_t next_pitch;
  int64_t next_frame;
  int32_t next_w, next_h;
  int32_t first_slot;
  float depth_scale, normal_depth_tol;
  uint32_t d_lo, d_span;     // valid raw depth: (d - d_lo) <= d_span  <=>  d != 0 && z_min <= d*scale <= z_max
  int32_t pyr_tol;
};

struct IcpArgs {
  LevelGeom g;
  LevelStore lv;
  const int2* pairs;          // (src slot, dst slot)
  int32_t pair_offset;        // index of the first pair this launch handles
  const float* pose_f32;      // 12 per pair: row-major R, t
  float* partials;            // [pair][max_blocks][kAccPad]
  uint32_t* tickets;          // [pair]
  int32_t max_blocks;
  int32_t blocks_per_pair, chunks_per_row, n_chunks;
  int32_t groups;             // groups of kChunksPerBlock chunks per block
  int32_t group_dv, group_du; // kChunksPerBlock chunks = group_dv rows + group_du columns
  uint32_t cpr_magic;         // ceil(2^32 / chunks_per_row): c / chunks_per_row == umulhi(c, magic)
  uint32_t d_lo, d_span;      // valid raw depth: (d - d_lo) <= d_span  <=>  d != 0 && z_min <= d*scale <= z_max
  uint32_t guard_texel;       // w * h: index of the all-zero texel behind every geometry frame (gather target of rejected pixels)
  float umax, vmax;           // w - 0.5, h - 0.5
  float depth_scale, dmax2, ncos_min, robust_scale;
  float sqrt_lambda;          // sqrt(photo_weight), photometric variant only
  /* finalize */
  double* pose_master;        // 12 per pair
  float* pose_f32_out;        // == pose_f32 (written by the last block)
  float* poses_cm;            // 16 per pair, column-major 4x4 result
  rst_stats* stats;
  int32_t min_count;
  float damping;
  float converge_eps;         // > 0: set done[pair] when an update is smaller than this
  uint8_t* done;              // [pair]: skip the rest of the level (nullptr when converge_eps == 0)
  int32_t update_pose;        // 0: evaluate only
  int32_t pdl;                // launched with programmatic stream serialization: wait for the previous launch in-kernel
  int32_t* idx_out;           // WRITE_IDX: [pair-local][h*w]
};

/* one pyramid level as the fused kernel sees it */
struct FusedLevel {
  LevelGeom g;
  LevelStore lv;
  int32_t chunks_per_row;
  int32_t n_groups;           // groups of kChunksPerBlock chunks covering the level
  int32_t groups_per_cta;     // ceil(n_groups / cluster size): contiguous share of one CTA
  int32_t group_dv, group_du; // kChunksPerBlock chunks = group_dv rows + group_du columns
  int32_t iters;              // iterations on this level
  uint32_t cpr_magic;         // ceil(2^32 / chunks_per_row)
  uint32_t guard_texel;       // w * h
};

/* k_icp_fused: every iteration of levels level_hi .. level_lo of a batch of pairs, one cluster per pair */
struct FusedArgs {
  FusedLevel lvl[RST_MAX_LEVELS];
  int32_t level_hi, level_lo;
  const int2* pairs;          // (src slot, dst slot)
  int32_t pair_offset;
  float* pose_f32;            // 12 per pair, in: pose of the first iteration, out: final
  double* pose_master;        // 12 per pair, in/out
  float* poses_cm;            // 16 per pair, column-major 4x4 result
  rst_stats* stats;           // in: reset by k_init_pairs, out: last evaluated iterate
  uint32_t d_lo, d_span;
  float depth_scale, dmax2, ncos_min, robust_scale, sqrt_lambda;
  int32_t min_count;
  float damping, converge_eps;
};

struct InitArgs {
  const float* poses_cm_in;   // 16 per pair
  double* pose_master;
  float* pose_f32;
  float* poses_cm;
  rst_stats* stats;
  uint32_t* tickets;
  int32_t n_pairs;
};

cudaError_t launch_preprocess(const PreArgs& a, const TensorMap& depth_map, int n_frames, cudaStream_t s);
cudaError_t launch_icp_iter(const IcpArgs& a, int n_pairs, int robust_kind, bool normal_gate,
                            bool write_idx, bool photo, cudaStream_t s);
cudaError_t launch_icp_fused(const FusedArgs& a, int n_pairs, int cluster, int robust_kind, bool normal_gate,
                             bool photo, cudaStream_t s);
int fused_max_active_clusters(int cluster);

/* f2: grey intensity (level 0 from CV_8UC3 RGB, then 2x2 means) */
struct IntensityArgs {
  const uint8_t* rgb;        // [frame][h][w][3] dense (level 0 only)
  const float* in;           // previous level (levels >= 1)
  float* out;
  int32_t w, h;              // OUTPUT size
  int32_t in_w;              // input width (levels >= 1)
  int64_t in_frame, out_frame, rgb_frame;
  int32_t first_slot;
};
cudaError_t launch_intensity(const IntensityArgs& a, int n_frames, cudaStream_t s);
cudaError_t launch_init_pairs(const InitArgs& a, cudaStream_t s);

}  // namespace rst
