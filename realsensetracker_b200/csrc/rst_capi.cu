/*
 * rst_capi.cu — the C ABI of include/rst_align.h over the sm_100a kernels.
 * Host side of the drop-in boundary: context, HBM frame store, launch schedule.
 * No CPU compute path exists here: every entry point either runs the CUDA kernels
 * or returns an error code.
 */
#include <cuda_runtime.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>
#include <vector>

#include "rst_align.h"
#include "rst_internal.h"
#include <cuda.h>   // CUtensorMap + the cuTensorMapEncodeTiled prototype (resolved at run time, libcuda is not linked)

#include "rst_kernels.cuh"

using namespace rst;

struct rst_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  bool own_stream = false;
  int max_w = 0, max_h = 0, max_frames = 0, max_pairs = 0;

  // geometry of the current rst_begin()
  bool begun = false;
  int w = 0, h = 0, num_levels = 0;
  rst_intrinsics K{};
  rst_params P{};
  LevelGeom geom[RST_MAX_LEVELS]{};
  int pitch[RST_MAX_LEVELS]{};
  int64_t dframe[RST_MAX_LEVELS]{};
  int64_t gframe[RST_MAX_LEVELS]{};
  int blocks_per_pair[RST_MAX_LEVELS]{}, chunks_per_row[RST_MAX_LEVELS]{}, n_chunks[RST_MAX_LEVELS]{};
  int groups[RST_MAX_LEVELS]{};
  uint32_t d_lo = 1, d_span = 0;
  float range_scale = -1.f, range_zmin = 0.f, range_zmax = 0.f;  // parameters d_lo/d_span were derived from

  // HBM frame store (allocated once for max_w x max_h x max_frames)
  uint16_t* d_depth[RST_MAX_LEVELS]{};
  size_t depth_bytes[RST_MAX_LEVELS]{};
  float4* d_geom[RST_MAX_LEVELS]{};
  size_t geom_bytes[RST_MAX_LEVELS]{};
  // photometric term (allocated on first use): RGB staging + intensity pyramid
  bool photo = false;
  uint8_t* d_rgb = nullptr;
  float* d_int[RST_MAX_LEVELS]{};
  bool ext0 = false;  // level 0 read in place from caller memory
  const uint16_t* ext_depth0 = nullptr;
  int ext_pitch0 = 0;
  int64_t ext_frame0 = 0;
  int ext_first = 0, ext_count = 0;     // slots [ext_first, ext_first + ext_count) are backed by the caller's memory
  TensorMap tmap[RST_MAX_LEVELS]{};     // (w, h, frames) uint16 view of every depth level for k_preprocess' TMA box
  TensorMap tmap_ext0{};                // the same for level 0 bound in place (frame 0 = slot ext_first)
  bool store_dirty = true;

  // pair state (max_pairs + 1: the last entry is the rst_evaluate scratch pair)
  int2* d_pairs = nullptr;
  float* d_poses_in = nullptr;
  double* d_master = nullptr;
  float* d_pose_f32 = nullptr;
  float* d_poses_cm = nullptr;
  rst_stats* d_stats = nullptr;
  uint32_t* d_tickets = nullptr;
  uint8_t* d_done = nullptr;            // per pair: converged on the current level (converge_eps > 0)
  float* d_partials = nullptr;
  int max_blocks = 0;
  int32_t* d_idx = nullptr;
  size_t idx_bytes = 0;

  // pinned staging
  int2* h_pairs = nullptr;
  float* h_poses = nullptr;
  rst_stats* h_stats = nullptr;

  cudaStream_t copy_stream = nullptr;   // H2D of the next chunk overlaps compute of the current one
  cudaStream_t work_stream[4] = {nullptr, nullptr, nullptr, nullptr};  // chunks alternate between two compute streams
  int schedule = RST_SCHEDULE_AUTO;     // see rst_set_schedule
  int cluster_size[2] = {4, 8};         // CTAs per pair of the fused kernel, by rst_params.tiling (throughput, latency)
  int split_ways = 2;                   // measured on B200: 2, 3 and 4 ways are equal (2.18 ms per 128-pair step)
  int split_min_pairs = 32;             // batches of at least this many pairs iterate as two halves on two streams
  int pipeline_chunk = 0;               // frames (pairs) per upload/compute chunk of the host entry points; 0 = automatic, < 0 = never
  void* ext = nullptr;                  // state of the cloud-based engine (rst_icp3d.cu), created on first use
  void (*ext_free)(void*) = nullptr;
  // CUDA graphs of the kernel part of small blocking calls (the latency path): a replay costs one enqueue instead
  // of ~27 and removes the launch gaps between the 23 dependent kernels of a pair
  struct GraphKey {
    int32_t mode, w, h, n_frames, n_pairs, schedule;   // mode 0: sequence, 1: pairs
    rst_intrinsics K;
    rst_params P;
  };
  struct CachedGraph { GraphKey key; cudaGraphExec_t exec; int launches; uint64_t stamp; };
  std::vector<CachedGraph> graphs;
  int pdl_max_pairs = 8;                // batches up to this size chain their iteration launches with programmatic dependent launch
  int graph_max_pairs = 8;              // blocking calls with at most this many pairs replay a graph (0 = never)
  uint64_t graph_clock = 0;
  int n_pairs_last = 0;
  int fetch_pending = 0;                // pairs whose results sit in the pinned staging of an *_async call
  int64_t launches = 0;
  std::string err;

  // per-stage event timing (rst_profile_*)
  struct ProfRec { int kind, level, launches; int64_t units; cudaEvent_t e0, e1; };
  bool profiling = false;
  std::vector<ProfRec> prof_open;
  std::vector<cudaEvent_t> ev_pool;
  rst_profile prof{};
};

static cudaEvent_t prof_event(rst_ctx* c) {
  cudaEvent_t e = nullptr;
  if (!c->ev_pool.empty()) { e = c->ev_pool.back(); c->ev_pool.pop_back(); }
  else cudaEventCreate(&e);
  return e;
}
/* kind 0 = preprocess, 1 = icp (per-iteration schedule, by level), 2 = fused icp */
static int prof_begin(rst_ctx* c, int kind, int level) {
  if (!c->profiling) return -1;
  rst_ctx::ProfRec r{kind, level, 0, 0, prof_event(c), prof_event(c)};
  cudaEventRecord(r.e0, c->stream);
  c->prof_open.push_back(r);
  return (int)c->prof_open.size() - 1;
}
static void prof_end(rst_ctx* c, int h, int launches, int64_t units) {
  if (h < 0) return;
  c->prof_open[h].launches = launches;
  c->prof_open[h].units = units;
  cudaEventRecord(c->prof_open[h].e1, c->stream);
}

static thread_local std::string g_create_err;

namespace rst {
cudaStream_t ctx_stream(rst_ctx* c) { return c->stream; }
int ctx_device(rst_ctx* c) { return c->device; }
void ctx_set_error(rst_ctx* c, const std::string& msg) { c->err = msg; }
void ctx_count_launches(rst_ctx* c, int n) { c->launches += n; }
void** ctx_ext_slot(rst_ctx* c, void (***free_fn)(void*)) { *free_fn = &c->ext_free; return &c->ext; }
}  // namespace rst

#define RST_CUDA(ctx, expr)                                                              \
  do {                                                                                   \
    cudaError_t e_ = (expr);                                                             \
    if (e_ != cudaSuccess) {                                                             \
      (ctx)->err = std::string(#expr) + ": " + cudaGetErrorString(e_);                   \
      return RST_ERR_CUDA;                                                               \
    }                                                                                    \
  } while (0)

static int fail(rst_ctx* c, int code, const char* msg) {
  if (c) c->err = msg;
  return code;
}

static inline int round_up(int x, int m) { return (x + m - 1) / m * m; }
constexpr int kGeomGuard = 8;  // all-zero float4 texels in front of every geometry frame (index -1 = rejected pixel)

/* pyramid level geometry — identical expressions to the CPU specification */
static void level_geom(const rst_intrinsics& K, int w, int h, int level, LevelGeom* g) {
  g->w = w; g->h = h; g->fx = K.fx; g->fy = K.fy; g->cx = K.cx; g->cy = K.cy;
  for (int l = 0; l < level; ++l) {
    g->w /= 2; g->h /= 2;
    g->fx = g->fx * 0.5f; g->fy = g->fy * 0.5f;
    g->cx = (g->cx + 0.5f) * 0.5f - 0.5f;
    g->cy = (g->cy + 0.5f) * 0.5f - 0.5f;
  }
  g->ifx = 1.0f / g->fx;
  g->ify = 1.0f / g->fy;
}

extern "C" {

int32_t rst_abi_version(void) { return RST_ABI_VERSION; }

void rst_params_default(rst_params* p) {
  if (!p) return;
  std::memset(p, 0, sizeof(*p));
  p->num_levels = 3;
  p->iters[0] = 10; p->iters[1] = 5; p->iters[2] = 4; p->iters[3] = 0;
  p->depth_scale = 0.001f;
  p->z_min = 0.1f; p->z_max = 10.0f;
  p->dist_max = 0.2f;
  p->normal_cos_min = -2.0f;
  p->normal_depth_tol = 0.05f;
  p->pyr_depth_tol = 100;
  p->robust_kind = RST_ROBUST_NONE;
  p->robust_scale = 0.02f;
  p->min_count = 16;
  p->damping = 0.0f;
  p->photo_weight = 0.0f;
}

const char* rst_last_error(const rst_ctx* ctx) { return ctx ? ctx->err.c_str() : "null context"; }
const char* rst_last_create_error(void) { return g_create_err.c_str(); }
int64_t rst_launch_count(const rst_ctx* ctx) { return ctx ? ctx->launches : 0; }

void rst_ctx_destroy(rst_ctx* c) {
  if (!c) return;
  cudaSetDevice(c->device);
  if (c->stream) cudaStreamSynchronize(c->stream);
  for (auto& g : c->graphs) cudaGraphExecDestroy(g.exec);
  if (c->ext && c->ext_free) c->ext_free(c->ext);
  for (int l = 0; l < RST_MAX_LEVELS; ++l) { cudaFree(c->d_depth[l]); cudaFree(c->d_geom[l]); }
  cudaFree(c->d_pairs); cudaFree(c->d_poses_in); cudaFree(c->d_master); cudaFree(c->d_pose_f32);
  cudaFree(c->d_poses_cm); cudaFree(c->d_stats); cudaFree(c->d_tickets); cudaFree(c->d_partials); cudaFree(c->d_done);
  cudaFree(c->d_idx);
  cudaFree(c->d_rgb);
  for (int l = 0; l < RST_MAX_LEVELS; ++l) cudaFree(c->d_int[l]);
  cudaFreeHost(c->h_pairs); cudaFreeHost(c->h_poses); cudaFreeHost(c->h_stats);
  for (auto& r : c->prof_open) { cudaEventDestroy(r.e0); cudaEventDestroy(r.e1); }
  for (auto e : c->ev_pool) cudaEventDestroy(e);
  if (c->copy_stream) cudaStreamDestroy(c->copy_stream);
  for (int k = 0; k < 4; ++k) if (c->work_stream[k]) cudaStreamDestroy(c->work_stream[k]);
  if (c->own_stream && c->stream) cudaStreamDestroy(c->stream);
  delete c;
}

int32_t rst_ctx_create(int32_t device, int32_t max_w, int32_t max_h, int32_t max_frames,
                       int32_t max_pairs, void* stream, rst_ctx** out) {
  g_create_err.clear();
  if (!out) { g_create_err = "out_ctx is null"; return RST_ERR_INVALID_ARG; }
  *out = nullptr;
  if (max_w < 16 || max_h < 16 || max_frames < 2 || max_pairs < 1 || max_w > 16384 || max_h > 16384 ||
      (int64_t)max_w * max_h > (1ll << 27)) {  // texel byte offsets inside a frame are 32-bit
    g_create_err = "bad capacity (need 16 <= max_w,max_h <= 16384, max_w*max_h <= 2^27, max_frames >= 2, max_pairs >= 1)";
    return RST_ERR_INVALID_ARG;
  }
  int n_dev = 0;
  cudaError_t e = cudaGetDeviceCount(&n_dev);
  if (e != cudaSuccess || n_dev <= 0) {
    g_create_err = std::string("no CUDA device: ") + cudaGetErrorString(e);
    return RST_ERR_NO_DEVICE;
  }
  if (device < 0 || device >= n_dev) { g_create_err = "device id out of range"; return RST_ERR_NO_DEVICE; }
  cudaDeviceProp prop;
  if ((e = cudaGetDeviceProperties(&prop, device)) != cudaSuccess) {
    g_create_err = cudaGetErrorString(e);
    return RST_ERR_CUDA;
  }
  if (prop.major != 10) {
    g_create_err = "device is sm_" + std::to_string(prop.major) + std::to_string(prop.minor) +
                   "; this library carries sm_100a code only";
    return RST_ERR_ARCH;
  }
  rst_ctx* c = new (std::nothrow) rst_ctx();
  if (!c) { g_create_err = "out of host memory"; return RST_ERR_INVALID_ARG; }
  c->device = device; c->max_w = max_w; c->max_h = max_h; c->max_frames = max_frames; c->max_pairs = max_pairs;
#define CREATE_TRY(expr)                                                              \
  do {                                                                                \
    cudaError_t e2_ = (expr);                                                         \
    if (e2_ != cudaSuccess) {                                                         \
      g_create_err = std::string(#expr) + ": " + cudaGetErrorString(e2_);             \
      rst_ctx_destroy(c);                                                             \
      return RST_ERR_CUDA;                                                            \
    }                                                                                 \
  } while (0)
  CREATE_TRY(cudaSetDevice(device));
  if (stream) { c->stream = (cudaStream_t)stream; }
  else { CREATE_TRY(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking)); c->own_stream = true; }
  CREATE_TRY(cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
  for (int k = 0; k < 4; ++k) CREATE_TRY(cudaStreamCreateWithFlags(&c->work_stream[k], cudaStreamNonBlocking));
  int w = max_w, h = max_h;
  for (int l = 0; l < RST_MAX_LEVELS; ++l) {
    const size_t px = (size_t)round_up(w, 8) * h;
    c->depth_bytes[l] = px * sizeof(uint16_t) * max_frames;
    CREATE_TRY(cudaMalloc(&c->d_depth[l], c->depth_bytes[l]));
    c->geom_bytes[l] = ((size_t)w * h + kGeomGuard) * sizeof(float4) * max_frames + kGeomGuard * sizeof(float4);
    CREATE_TRY(cudaMalloc(&c->d_geom[l], c->geom_bytes[l]));
    w /= 2; h /= 2;
    if (w < 1) w = 1;
    if (h < 1) h = 1;
  }
  const int np = max_pairs + 1;
  const int cpr = (max_w + kChunkPx - 1) / kChunkPx;
  c->max_blocks = (cpr * max_h + kChunksPerBlock - 1) / kChunksPerBlock;
  CREATE_TRY(cudaMalloc(&c->d_pairs, sizeof(int2) * np));
  CREATE_TRY(cudaMalloc(&c->d_poses_in, sizeof(float) * 16 * np));
  CREATE_TRY(cudaMalloc(&c->d_master, sizeof(double) * 12 * np));
  CREATE_TRY(cudaMalloc(&c->d_pose_f32, sizeof(float) * 12 * np));
  CREATE_TRY(cudaMalloc(&c->d_poses_cm, sizeof(float) * 16 * np));
  CREATE_TRY(cudaMalloc(&c->d_stats, sizeof(rst_stats) * np));
  CREATE_TRY(cudaMalloc(&c->d_tickets, sizeof(uint32_t) * np));
  CREATE_TRY(cudaMemset(c->d_tickets, 0, sizeof(uint32_t) * np));
  CREATE_TRY(cudaMalloc(&c->d_done, np));
  CREATE_TRY(cudaMemset(c->d_done, 0, np));
  CREATE_TRY(cudaMalloc(&c->d_partials, sizeof(float) * kAccPad * (size_t)c->max_blocks * np));
  CREATE_TRY(cudaMallocHost(&c->h_pairs, sizeof(int2) * np));
  CREATE_TRY(cudaMallocHost(&c->h_poses, sizeof(float) * 16 * np));
  CREATE_TRY(cudaMallocHost(&c->h_stats, sizeof(rst_stats) * np));
#undef CREATE_TRY
  if (const char* e = std::getenv("RST_FUSED_CLUSTER")) {   // experiments: CTAs per pair of the fused kernel
    const int v = std::atoi(e);
    if (v >= 1 && v <= 16) c->cluster_size[0] = c->cluster_size[1] = v;
  }
  if (const char* e = std::getenv("RST_PDL_MAX_PAIRS")) c->pdl_max_pairs = std::atoi(e);
  if (const char* e = std::getenv("RST_SCHEDULE")) {
    const int v = std::atoi(e);
    if (v >= RST_SCHEDULE_AUTO && v <= RST_SCHEDULE_HYBRID) c->schedule = v;
  }
  *out = c;
  return RST_OK;
}

/* (w, h, frames) uint16 tensor over a depth level for the TMA box of k_preprocess (kPreBoxW x kPreBoxH x 1). The
 * width is the IMAGE width, not the pitch: row padding is out of bounds for the copy engine and reads as zero. */
static int32_t encode_depth_map(rst_ctx* c, const uint16_t* base, int w, int h, int pitch_px, int64_t frame_px, int frames,
                                TensorMap* out) {
  typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                               const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                               CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  static EncodeFn encode = nullptr;
  if (!encode) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess || !fn) {
      cudaGetLastError();
      return fail(c, RST_ERR_CUDA, "cuTensorMapEncodeTiled is not available in this driver");
    }
    encode = reinterpret_cast<EncodeFn>(fn);
  }
  static_assert(sizeof(TensorMap) == sizeof(CUtensorMap) && alignof(TensorMap) >= alignof(CUtensorMap), "TensorMap must mirror CUtensorMap");
  const cuuint64_t dims[3] = {(cuuint64_t)w, (cuuint64_t)h, (cuuint64_t)frames};
  const cuuint64_t strides[2] = {(cuuint64_t)pitch_px * 2, (cuuint64_t)frame_px * 2};   // bytes, dimensions 1 and 2
  const cuuint32_t box[3] = {(cuuint32_t)kPreBoxW, (cuuint32_t)kPreBoxH, 1u};
  const cuuint32_t estr[3] = {1u, 1u, 1u};
  const CUresult r = encode(reinterpret_cast<CUtensorMap*>(out), CU_TENSOR_MAP_DATA_TYPE_UINT16, 3, const_cast<uint16_t*>(base), dims, strides,
                            box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(c, RST_ERR_CUDA, ("cuTensorMapEncodeTiled failed (" + std::to_string((int)r) + ")").c_str());
  return RST_OK;
}

int32_t rst_begin(rst_ctx* c, int32_t width, int32_t height, const rst_intrinsics* intr,
                  const rst_params* params) {
  if (!c) return RST_ERR_INVALID_ARG;
  if (!intr || !params) return fail(c, RST_ERR_INVALID_ARG, "intr/params is null");
  if (width < 16 || height < 16) return fail(c, RST_ERR_INVALID_ARG, "frame smaller than 16x16");
  if (width > c->max_w || height > c->max_h) return fail(c, RST_ERR_CAPACITY, "frame larger than the context capacity");
  const rst_params& P = *params;
  if (P.num_levels < 1 || P.num_levels > RST_MAX_LEVELS) return fail(c, RST_ERR_INVALID_ARG, "num_levels out of range");
  for (int l = 0; l < P.num_levels; ++l)
    if (P.iters[l] < 0 || P.iters[l] > 10000) return fail(c, RST_ERR_INVALID_ARG, "iters out of range");
  if (!(P.photo_weight >= 0.f)) return fail(c, RST_ERR_INVALID_ARG, "photo_weight must be >= 0");
  if (!(P.depth_scale > 0.f) || !(P.dist_max > 0.f) || !(P.z_max > P.z_min) || !(intr->fx > 0.f) || !(intr->fy > 0.f))
    return fail(c, RST_ERR_INVALID_ARG, "depth_scale, dist_max, z range and focal lengths must be positive");
  if (P.robust_kind < 0 || P.robust_kind > 2) return fail(c, RST_ERR_INVALID_ARG, "unknown robust_kind");
  if (P.tiling < 0 || P.tiling > 1) return fail(c, RST_ERR_INVALID_ARG, "unknown tiling");
  if (!(P.converge_eps >= 0.f)) return fail(c, RST_ERR_INVALID_ARG, "converge_eps must be >= 0");
  if (P.robust_kind != RST_ROBUST_NONE && !(P.robust_scale > 0.f)) return fail(c, RST_ERR_INVALID_ARG, "robust_scale must be positive");
  if ((width >> (P.num_levels - 1)) < 8 || (height >> (P.num_levels - 1)) < 8)
    return fail(c, RST_ERR_INVALID_ARG, "coarsest pyramid level smaller than 8x8");
  RST_CUDA(c, cudaSetDevice(c->device));
  const bool same = c->begun && c->w == width && c->h == height;
  c->w = width; c->h = height; c->K = *intr; c->P = P; c->num_levels = P.num_levels;
  for (int l = 0; l < P.num_levels; ++l) {
    level_geom(*intr, width, height, l, &c->geom[l]);
    c->pitch[l] = round_up(c->geom[l].w, 8);
    c->dframe[l] = (int64_t)c->pitch[l] * c->geom[l].h;
    c->gframe[l] = (int64_t)c->geom[l].w * c->geom[l].h + kGeomGuard;  // guard texels in front of every frame
    c->chunks_per_row[l] = (c->geom[l].w + kChunkPx - 1) / kChunkPx;
    c->n_chunks[l] = c->chunks_per_row[l] * c->geom[l].h;
    // block extent depends on the image size only, so results never depend on the batch size
    {
      // pixels per block ~ level size / 8, a power-of-two number of groups, at most kMaxGroups
      const int group_px = kChunksPerBlock * kChunkPx;
      int g = 1;
      if (P.tiling != RST_TILING_LATENCY) {
        while (g * 2 <= kMaxGroups && (int64_t)g * 2 * group_px * 8 <= (int64_t)c->n_chunks[l] * kChunkPx) g *= 2;
      } else {
        // latency tiling: the smallest blocks that still give one pair about one block per SM (148): a single pair
        // fills the GPU in one wave and the last block sums ~150 partials instead of 600
        while (g * 2 <= kMaxGroups && (int64_t)g * 2 * group_px * 148 <= (int64_t)c->n_chunks[l] * kChunkPx) g *= 2;
      }
      c->groups[l] = g;
    }
    const int cpb = kChunksPerBlock * c->groups[l];
    c->blocks_per_pair[l] = (c->n_chunks[l] + cpb - 1) / cpb;
  }
  if (!(c->range_scale == P.depth_scale && c->range_zmin == P.z_min && c->range_zmax == P.z_max)) {
    // integer form of the depth validity test: d != 0 && z_min <= float(d)*scale <= z_max (monotone in d);
    // the scan over all 65535 raw values is redone only when the three parameters change
    uint32_t lo = 65536u, hi = 0u;
    for (uint32_t d = 1; d <= 65535u; ++d) {
      const float z = (float)d * P.depth_scale;
      if (z >= P.z_min && z <= P.z_max) { if (d < lo) lo = d; hi = d; }
    }
    if (hi < lo) return fail(c, RST_ERR_INVALID_ARG, "no uint16 depth value falls inside [z_min, z_max]");
    c->d_lo = lo; c->d_span = hi - lo;
    c->range_scale = P.depth_scale; c->range_zmin = P.z_min; c->range_zmax = P.z_max;
  }
  if (!same || c->store_dirty) {
    // row padding (columns >= w) must read as invalid depth
    for (int l = 0; l < RST_MAX_LEVELS; ++l) {
      RST_CUDA(c, cudaMemsetAsync(c->d_depth[l], 0, c->depth_bytes[l], c->stream));
      RST_CUDA(c, cudaMemsetAsync(c->d_geom[l], 0, c->geom_bytes[l], c->stream));  // guard texels must read as invalid
    }
    c->store_dirty = false;
  }
  c->ext0 = false; c->ext_depth0 = nullptr; c->ext_first = 0; c->ext_count = 0;
  for (int l = 0; l < P.num_levels; ++l) {
    const int32_t rc = encode_depth_map(c, c->d_depth[l], c->geom[l].w, c->geom[l].h, c->pitch[l], c->dframe[l], c->max_frames, &c->tmap[l]);
    if (rc != RST_OK) return rc;
  }
  c->photo = P.photo_weight > 0.0f;
  if (c->photo && !c->d_rgb) {
    RST_CUDA(c, cudaMalloc(&c->d_rgb, (size_t)c->max_w * c->max_h * 3 * c->max_frames));
    int mw = c->max_w, mh = c->max_h;
    for (int l = 0; l < RST_MAX_LEVELS; ++l) {
      RST_CUDA(c, cudaMalloc(&c->d_int[l], (size_t)mw * mh * sizeof(float) * c->max_frames));
      mw = mw / 2 > 0 ? mw / 2 : 1; mh = mh / 2 > 0 ? mh / 2 : 1;
    }
  }
  c->begun = true;
  c->err.clear();
  return RST_OK;
}

int32_t rst_upload_frames(rst_ctx* c, const rst_frame* frames, int32_t n, int32_t first_slot) {
  if (!c) return RST_ERR_INVALID_ARG;
  if (!c->begun) return fail(c, RST_ERR_INVALID_ARG, "rst_begin has not been called");
  if (!frames || n < 0 || first_slot < 0) return fail(c, RST_ERR_INVALID_ARG, "bad frames/n/first_slot");
  if (first_slot + n > c->max_frames) return fail(c, RST_ERR_CAPACITY, "more frames than the context holds");
  if (c->ext0) return fail(c, RST_ERR_INVALID_ARG, "level 0 is bound to device memory; call rst_begin first");
  RST_CUDA(c, cudaSetDevice(c->device));
  for (int i = 0; i < n; ++i) {
    const rst_frame& f = frames[i];
    if (!f.depth) return fail(c, RST_ERR_INVALID_ARG, "frame.depth is null");
    if (f.width != c->w || f.height != c->h) return fail(c, RST_ERR_INVALID_ARG, "frame size differs from rst_begin");
    if (f.depth_stride_bytes < c->w * 2 || (f.depth_stride_bytes & 1)) return fail(c, RST_ERR_INVALID_ARG, "bad depth stride");
    if (c->photo && (!f.rgb || f.rgb_stride_bytes < c->w * 3)) return fail(c, RST_ERR_INVALID_ARG, "photo_weight > 0 needs an rgb image in every frame");
  }
  if (c->photo) {
    const size_t fb = (size_t)c->w * c->h * 3;
    for (int k = 0; k < n; ++k)
      RST_CUDA(c, cudaMemcpy2DAsync(c->d_rgb + fb * (size_t)(first_slot + k), (size_t)c->w * 3, frames[k].rgb,
                                    (size_t)frames[k].rgb_stride_bytes, (size_t)c->w * 3, (size_t)c->h, cudaMemcpyHostToDevice, c->stream));
  }
  const size_t dpitch = (size_t)c->pitch[0] * 2;
  int i = 0;
  while (i < n) {
    // merge frames that lie back to back in host memory into one copy
    int j = i;
    const size_t stride = (size_t)frames[i].depth_stride_bytes;
    while (j + 1 < n && (size_t)frames[j + 1].depth_stride_bytes == stride &&
           (const uint8_t*)frames[j + 1].depth == (const uint8_t*)frames[j].depth + stride * c->h)
      ++j;
    const int cnt = j - i + 1;
    uint16_t* dst = c->d_depth[0] + (int64_t)(first_slot + i) * c->dframe[0];
    if (stride == dpitch && stride == (size_t)c->w * 2) {
      // dense on both sides: one linear copy (a 2-D copy with a row pitch that is not a multiple of
      // 64 bytes, e.g. 848 px, runs at a fraction of the PCIe rate)
      RST_CUDA(c, cudaMemcpyAsync(dst, frames[i].depth, stride * c->h * cnt, cudaMemcpyHostToDevice, c->stream));
    } else {
      RST_CUDA(c, cudaMemcpy2DAsync(dst, dpitch, frames[i].depth, stride, (size_t)c->w * 2, (size_t)c->h * cnt,
                                    cudaMemcpyHostToDevice, c->stream));
    }
    i = j + 1;
  }
  return RST_OK;
}

int32_t rst_set_frames_device(rst_ctx* c, const uint16_t* d_depth, int32_t n, int32_t row_stride_px,
                              int64_t frame_stride_px, int32_t first_slot) {
  if (!c) return RST_ERR_INVALID_ARG;
  if (!c->begun) return fail(c, RST_ERR_INVALID_ARG, "rst_begin has not been called");
  if (!d_depth || n < 1 || first_slot < 0) return fail(c, RST_ERR_INVALID_ARG, "bad d_depth/n/first_slot");
  if (first_slot + n > c->max_frames) return fail(c, RST_ERR_CAPACITY, "more frames than the context holds");
  if (row_stride_px < c->w || frame_stride_px < (int64_t)row_stride_px * c->h)
    return fail(c, RST_ERR_INVALID_ARG, "strides smaller than the frame");
  if (c->photo) return fail(c, RST_ERR_INVALID_ARG, "the photometric term needs host frames with rgb (rst_upload_frames)");
  if (((uintptr_t)d_depth & 15) || (row_stride_px & 7) || (frame_stride_px & 7))
    return fail(c, RST_ERR_ALIGNMENT, "device depth must be 16-byte aligned with row/frame strides multiples of 8 pixels");
  {
    const int32_t rc = encode_depth_map(c, d_depth, c->w, c->h, row_stride_px, frame_stride_px, n, &c->tmap_ext0);
    if (rc != RST_OK) return rc;
  }
  c->ext0 = true;
  c->ext_depth0 = d_depth - (int64_t)first_slot * frame_stride_px;
  c->ext_pitch0 = row_stride_px;
  c->ext_frame0 = frame_stride_px;
  c->ext_first = first_slot;
  c->ext_count = n;
  return RST_OK;
}

/* while level 0 is read in place from caller memory only the bound slots exist */
static bool slot_bound(const rst_ctx* c, int slot) {
  if (slot < 0 || slot >= c->max_frames) return false;
  return !c->ext0 || (slot >= c->ext_first && slot < c->ext_first + c->ext_count);
}

/* one outstanding *_async call per context: its results sit in the pinned staging until rst_wait */
static int32_t check_idle(rst_ctx* c) {
  if (c->fetch_pending != 0) {
    c->err = "an asynchronous call is outstanding on this context: call rst_wait first";
    return RST_ERR_INVALID_ARG;
  }
  return RST_OK;
}

static LevelStore level_store(const rst_ctx* c, int l) {
  LevelStore s;
  if (l == 0 && c->ext0) { s.depth = c->ext_depth0; s.depth_pitch = c->ext_pitch0; s.depth_frame = c->ext_frame0; }
  else { s.depth = c->d_depth[l]; s.depth_pitch = c->pitch[l]; s.depth_frame = c->dframe[l]; }
  s.geom = c->d_geom[l] + kGeomGuard;
  s.geom_frame = c->gframe[l];
  s.intensity = c->photo ? c->d_int[l] : nullptr;
  s.int_frame = (int64_t)c->geom[l].w * c->geom[l].h;
  return s;
}

static int32_t preprocess_impl(rst_ctx* c, int first_slot, int n, bool write_geom) {
  if (c->photo) {
    for (int l = 0; l < c->num_levels; ++l) {
      IntensityArgs ia{};
      ia.w = c->geom[l].w; ia.h = c->geom[l].h; ia.first_slot = first_slot;
      ia.out = c->d_int[l]; ia.out_frame = (int64_t)ia.w * ia.h;
      if (l == 0) { ia.rgb = c->d_rgb; ia.rgb_frame = (int64_t)c->w * c->h * 3; }
      else { ia.in = c->d_int[l - 1]; ia.in_w = c->geom[l - 1].w; ia.in_frame = (int64_t)c->geom[l - 1].w * c->geom[l - 1].h; }
      RST_CUDA(c, launch_intensity(ia, n, c->stream));
      c->launches += 1;
    }
  }
  for (int l = 0; l < c->num_levels; ++l) {
    const bool has_next = l + 1 < c->num_levels;
    if (!write_geom && !has_next) break;
    PreArgs a{};
    a.g = c->geom[l];
    a.cur = level_store(c, l);
    if (!write_geom) a.cur.geom = nullptr;
    if (has_next) {
      a.next_depth = c->d_depth[l + 1]; a.next_pitch = c->pitch[l + 1]; a.next_frame = c->dframe[l + 1];
      a.next_w = c->geom[l + 1].w; a.next_h = c->geom[l + 1].h;
    }
    a.first_slot = first_slot;
    a.depth_scale = c->P.depth_scale; a.d_lo = c->d_lo; a.d_span = c->d_span;
    a.normal_depth_tol = c->P.normal_depth_tol; a.pyr_tol = c->P.pyr_depth_tol;
    const bool ext = l == 0 && c->ext0;
    a.tmap_slot0 = ext ? c->ext_first : 0;
    const int ph = prof_begin(c, 0, l);
    RST_CUDA(c, launch_preprocess(a, ext ? c->tmap_ext0 : c->tmap[l], n, c->stream));
    prof_end(c, ph, 1, n);
    c->launches += 1;
  }
  return RST_OK;
}

int32_t rst_preprocess(rst_ctx* c, int32_t first_slot, int32_t n) {
  if (!c) return RST_ERR_INVALID_ARG;
  if (!c->begun) return fail(c, RST_ERR_INVALID_ARG, "rst_begin has not been called");
  if (first_slot < 0 || n < 0 || first_slot + n > c->max_frames) return fail(c, RST_ERR_CAPACITY, "slot range out of capacity");
  if (n > 0 && (!slot_bound(c, first_slot) || !slot_bound(c, first_slot + n - 1)))
    return fail(c, RST_ERR_INVALID_ARG, "slot range outside the frames bound by rst_set_frames_device");
  RST_CUDA(c, cudaSetDevice(c->device));
  return preprocess_impl(c, first_slot, n, true);
}

static void fill_icp_args(const rst_ctx* c, int l, IcpArgs* a) {
  a->g = c->geom[l];
  a->lv = level_store(c, l);
  a->pairs = c->d_pairs;
  a->pair_offset = 0;
  a->pose_f32 = c->d_pose_f32;
  a->partials = c->d_partials;
  a->tickets = c->d_tickets;
  a->max_blocks = c->max_blocks;
  a->blocks_per_pair = c->blocks_per_pair[l];
  a->chunks_per_row = c->chunks_per_row[l];
  a->n_chunks = c->n_chunks[l];
  a->groups = c->groups[l];
  a->group_dv = kChunksPerBlock / c->chunks_per_row[l];
  a->group_du = (kChunksPerBlock % c->chunks_per_row[l]) * kChunkPx;
  a->cpr_magic = (uint32_t)(((1ull << 32) + (uint64_t)c->chunks_per_row[l] - 1) / (uint64_t)c->chunks_per_row[l]);
  a->d_lo = c->d_lo; a->d_span = c->d_span;
  a->guard_texel = (uint32_t)(c->geom[l].w * c->geom[l].h);
  a->umax = (float)c->geom[l].w - 0.5f; a->vmax = (float)c->geom[l].h - 0.5f;
  a->depth_scale = c->P.depth_scale;
  a->dmax2 = c->P.dist_max * c->P.dist_max;
  a->ncos_min = c->P.normal_cos_min;
  a->robust_scale = c->P.robust_scale;
  a->sqrt_lambda = sqrtf(c->P.photo_weight);
  a->pose_master = c->d_master;
  a->pose_f32_out = c->d_pose_f32;
  a->poses_cm = c->d_poses_cm;
  a->stats = c->d_stats;
  a->min_count = c->P.min_count;
  a->damping = c->P.damping;
  a->update_pose = 1;
  a->converge_eps = c->P.converge_eps;
  a->done = c->P.converge_eps > 0.f ? c->d_done : nullptr;
  a->idx_out = nullptr;
}

struct StreamScope {  // kernels/copies issued through the ctx go to `s` while this object lives
  rst_ctx* c; cudaStream_t saved;
  StreamScope(rst_ctx* c_, cudaStream_t s) : c(c_), saved(c_->stream) { c->stream = s; }
  ~StreamScope() { c->stream = saved; }
};

static int32_t link_streams(rst_ctx* c, cudaStream_t from, cudaStream_t to) {  // `to` waits for `from`
  cudaEvent_t e = prof_event(c);
  RST_CUDA(c, cudaEventRecord(e, from));
  RST_CUDA(c, cudaStreamWaitEvent(to, e, 0));
  c->ev_pool.push_back(e);  // safe to recycle: the wait is already enqueued
  return RST_OK;
}

/* device half of stage 1: the staged pair table + initial poses -> device, state reset */
static int32_t pairs_begin_enqueue(rst_ctx* c, int32_t n_pairs) {
  RST_CUDA(c, cudaMemcpyAsync(c->d_pairs, c->h_pairs, sizeof(int2) * n_pairs, cudaMemcpyHostToDevice, c->stream));
  RST_CUDA(c, cudaMemcpyAsync(c->d_poses_in, c->h_poses, sizeof(float) * 16 * n_pairs, cudaMemcpyHostToDevice, c->stream));
  InitArgs ia{c->d_poses_in, c->d_master, c->d_pose_f32, c->d_poses_cm, c->d_stats, c->d_tickets, n_pairs};
  RST_CUDA(c, launch_init_pairs(ia, c->stream));
  c->launches += 1;
  return RST_OK;
}

/* stage 1 of an alignment: pair table + initial poses into the pinned staging (and, unless the caller replays a
 * graph that contains it, on to the device) */
static int32_t pairs_begin(rst_ctx* c, const int32_t* src_slots, const int32_t* dst_slots, int32_t n_pairs,
                           const float* poses_in, bool enqueue = true) {
  if (check_idle(c) != RST_OK) return RST_ERR_INVALID_ARG;   // the pinned staging below still belongs to that call
  for (int i = 0; i < n_pairs; ++i) {
    if (!slot_bound(c, src_slots[i]) || !slot_bound(c, dst_slots[i]))
      return fail(c, RST_ERR_INVALID_ARG, "slot index out of range (or outside the frames bound by rst_set_frames_device)");
    c->h_pairs[i] = make_int2(src_slots[i], dst_slots[i]);
  }
  if (poses_in) {
    std::memcpy(c->h_poses, poses_in, sizeof(float) * 16 * n_pairs);
  } else {
    for (int i = 0; i < n_pairs; ++i)
      for (int k = 0; k < 16; ++k) c->h_poses[16 * i + k] = (k % 5 == 0) ? 1.f : 0.f;
  }
  c->n_pairs_last = n_pairs;
  return enqueue ? pairs_begin_enqueue(c, n_pairs) : RST_OK;
}

/* stage 2, fused schedule: ONE launch runs every iteration of every level for pairs [first, first + n);
 * a cluster of cluster_size[tiling] CTAs owns each pair (csrc/rst_icp_kernels.cu, k_icp_fused). The cluster size
 * depends on the tiling switch only — never on the batch — so results are bit-identical across batch sizes. */
static int32_t pairs_iterate_fused(rst_ctx* c, int first, int n, int level_hi, int level_lo) {
  const int C = c->cluster_size[c->P.tiling == RST_TILING_LATENCY ? 1 : 0];
  FusedArgs a{};
  for (int l = 0; l < c->num_levels; ++l) {
    FusedLevel& L = a.lvl[l];
    L.g = c->geom[l];
    L.lv = level_store(c, l);
    L.chunks_per_row = c->chunks_per_row[l];
    L.n_groups = (c->n_chunks[l] + kChunksPerBlock - 1) / kChunksPerBlock;
    L.groups_per_cta = (L.n_groups + C - 1) / C;
    L.group_dv = kChunksPerBlock / c->chunks_per_row[l];
    L.group_du = (kChunksPerBlock % c->chunks_per_row[l]) * kChunkPx;
    L.iters = c->P.iters[l];
    L.cpr_magic = (uint32_t)(((1ull << 32) + (uint64_t)c->chunks_per_row[l] - 1) / (uint64_t)c->chunks_per_row[l]);
    L.guard_texel = (uint32_t)(c->geom[l].w * c->geom[l].h);
  }
  a.level_hi = level_hi; a.level_lo = level_lo;
  a.pairs = c->d_pairs;
  a.pose_f32 = c->d_pose_f32; a.pose_master = c->d_master; a.poses_cm = c->d_poses_cm; a.stats = c->d_stats;
  a.d_lo = c->d_lo; a.d_span = c->d_span;
  a.depth_scale = c->P.depth_scale; a.dmax2 = c->P.dist_max * c->P.dist_max; a.ncos_min = c->P.normal_cos_min;
  a.robust_scale = c->P.robust_scale; a.sqrt_lambda = sqrtf(c->P.photo_weight);
  a.min_count = c->P.min_count; a.damping = c->P.damping; a.converge_eps = c->P.converge_eps;
  const bool ngate = c->P.normal_cos_min > -1.0f;
  int iters_total = 0;
  for (int l = level_lo; l <= level_hi; ++l) iters_total += c->P.iters[l];
  if (iters_total == 0) return RST_OK;
  const int ph = prof_begin(c, 2, 0);
  int nl = 0;
  for (int off = 0; off < n; off += 65535) {
    a.pair_offset = first + off;
    const int cnt = n - off < 65535 ? n - off : 65535;
    RST_CUDA(c, launch_icp_fused(a, cnt, C, c->P.robust_kind, ngate, c->photo, c->stream));
    c->launches += 1;
    ++nl;
  }
  prof_end(c, ph, nl, (int64_t)n * iters_total);
  return RST_OK;
}

/* stage 2: the whole coarse-to-fine schedule for pairs [first, first + n) */
static int32_t pairs_iterate(rst_ctx* c, int first, int n) {
  if (n <= 0) return RST_OK;
  // which levels run inside the fused cluster kernel (the rest: one launch per iteration)
  int fused_lo = c->num_levels;   // none
  if (c->schedule == RST_SCHEDULE_FUSED) fused_lo = 0;
  else if (c->schedule == RST_SCHEDULE_HYBRID) fused_lo = 1;
  // RST_SCHEDULE_AUTO: one launch per iteration — measured fastest on B200 at every batch size tried
  if (fused_lo < c->num_levels) {
    const int32_t rc = pairs_iterate_fused(c, first, n, c->num_levels - 1, fused_lo);
    if (rc != RST_OK) return rc;
  }
  const bool ngate = c->P.normal_cos_min > -1.0f;
  for (int l = fused_lo - 1; l >= 0; --l) {
    IcpArgs a{};
    fill_icp_args(c, l, &a);
    if (a.done) RST_CUDA(c, cudaMemsetAsync(c->d_done + first, 0, (size_t)n, c->stream));  // every level starts active
    const int ph = prof_begin(c, 1, l);
    int nl = 0;
    // consecutive iteration launches are chained with programmatic dependent launch: the next launch's blocks run their
    // pose-independent head (parameter setup, depth staging) while this launch reduces and solves. Small batches (the
    // latency path) let the next launch come up at once (pdl 1); large batches only when a block has left its pixel loop
    // (pdl 2: earlier, the waiting blocks would take SM slots from blocks that still have pixels to process; +2.2 %
    // at 128 pairs). RST_PDL_LARGE=0 restores plain stream order for large batches.
    a.pdl = (c->pdl_max_pairs > 0 && n <= c->pdl_max_pairs && !c->profiling) ? 1 : 0;
    if (!a.pdl && !c->profiling) { static const int big = [] { const char* e = std::getenv("RST_PDL_LARGE"); return e ? std::atoi(e) : 2; }(); a.pdl = big; }
    for (int it = 0; it < c->P.iters[l]; ++it) {
      for (int off = 0; off < n; off += 65535) {
        a.pair_offset = first + off;
        const int cnt = n - off < 65535 ? n - off : 65535;
        RST_CUDA(c, launch_icp_iter(a, cnt, c->P.robust_kind, ngate, false, c->photo, c->stream));
        c->launches += 1;
        ++nl;
      }
    }
    prof_end(c, ph, nl, (int64_t)n * c->P.iters[l]);
  }
  return RST_OK;
}

/* The iteration schedule of a large batch is issued as two halves on the two work streams: the
 * launches of one half fill the partial last wave (and the latency-bound coarse levels) of the other.
 * Pairs are independent and reduced in image-size-determined blocks, so the split never changes a bit. */
static int32_t pairs_iterate_split(rst_ctx* c, int n_pairs) {
  // the fused schedule is one launch whose clusters are all resident at once: nothing to interleave
  if (c->schedule == RST_SCHEDULE_FUSED || n_pairs < c->split_min_pairs) return pairs_iterate(c, 0, n_pairs);
  cudaStream_t main_s = c->stream;
  const int ways = c->split_ways;
  const int part = (n_pairs + ways - 1) / ways;
  int32_t rc;
  for (int k = 0; k < ways; ++k) {
    const int first = k * part, n = n_pairs - first < part ? n_pairs - first : part;
    if (n <= 0) break;
    if ((rc = link_streams(c, main_s, c->work_stream[k])) != RST_OK) return rc;
    StreamScope work(c, c->work_stream[k]);
    if ((rc = pairs_iterate(c, first, n)) != RST_OK) return rc;
  }
  for (int k = 0; k < ways; ++k) if ((rc = link_streams(c, c->work_stream[k], main_s)) != RST_OK) return rc;
  return RST_OK;
}

/* stage 3a: results -> pinned staging, asynchronously on the context stream */
static int32_t fetch_enqueue(rst_ctx* c, int32_t n_pairs) {
  RST_CUDA(c, cudaMemcpyAsync(c->h_poses, c->d_poses_cm, sizeof(float) * 16 * n_pairs, cudaMemcpyDeviceToHost, c->stream));
  RST_CUDA(c, cudaMemcpyAsync(c->h_stats, c->d_stats, sizeof(rst_stats) * n_pairs, cudaMemcpyDeviceToHost, c->stream));
  c->fetch_pending = n_pairs;
  return RST_OK;
}

/* stage 3b: wait for the stream, hand the staged results to the caller */
static int32_t fetch_finish(rst_ctx* c, float* poses_out, rst_stats* stats_out) {
  RST_CUDA(c, cudaStreamSynchronize(c->stream));
  const int n = c->fetch_pending;
  c->fetch_pending = 0;
  if (poses_out) std::memcpy(poses_out, c->h_poses, sizeof(float) * 16 * n);
  if (stats_out) std::memcpy(stats_out, c->h_stats, sizeof(rst_stats) * n);
  return RST_OK;
}

static int32_t pairs_fetch(rst_ctx* c, int32_t n_pairs, float* poses_out, rst_stats* stats_out) {
  if (!poses_out && !stats_out) return RST_OK;
  int32_t rc = fetch_enqueue(c, n_pairs);
  if (rc != RST_OK) return rc;
  return fetch_finish(c, poses_out, stats_out);
}

/* The kernel part of a small blocking host-frame call as one CUDA graph: pair table / pose upload, state reset, the
 * pre-processing launches, every iteration launch and the result copies into the pinned staging. Captured the first
 * time a (mode, size, parameters, pair count) combination is seen, replayed afterwards; the frame upload stays outside
 * (its source is caller memory). `pre` ranges: frames [0, n_frames) with geometry; for mode 1 the source frames
 * [n_pairs, 2 n_pairs) get geometry only with the normal gate. Returns RST_OK with *ran = false when graphs do not
 * apply (the caller then enqueues directly). */
static int32_t run_graphed(rst_ctx* c, int mode, int n_frames, int n_pairs, bool* ran) {
  *ran = false;
  if (c->graph_max_pairs <= 0 || n_pairs > c->graph_max_pairs || c->profiling || c->ext0) return RST_OK;
  rst_ctx::GraphKey key;
  std::memset(&key, 0, sizeof(key));
  key.mode = mode; key.w = c->w; key.h = c->h; key.n_frames = n_frames; key.n_pairs = n_pairs; key.schedule = c->schedule;
  key.K = c->K; key.P = c->P;
  rst_ctx::CachedGraph* hit = nullptr;
  for (auto& g : c->graphs)
    if (std::memcmp(&g.key, &key, sizeof(key)) == 0) { hit = &g; break; }
  if (!hit) {
    const int64_t launches0 = c->launches;
    if (cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeThreadLocal) != cudaSuccess) { cudaGetLastError(); return RST_OK; }
    int32_t rc = pairs_begin_enqueue(c, n_pairs);
    const bool ngate = c->P.normal_cos_min > -1.0f;
    if (rc == RST_OK) {
      if (mode == 0) rc = preprocess_impl(c, 0, n_frames, true);
      else {
        rc = preprocess_impl(c, 0, n_pairs, true);
        if (rc == RST_OK) rc = preprocess_impl(c, n_pairs, n_pairs, ngate);
      }
    }
    if (rc == RST_OK) rc = pairs_iterate_split(c, n_pairs);
    if (rc == RST_OK) {
      // result copies (fetch_enqueue without the pending flag: set by the caller after the launch)
      if (cudaMemcpyAsync(c->h_poses, c->d_poses_cm, sizeof(float) * 16 * n_pairs, cudaMemcpyDeviceToHost, c->stream) != cudaSuccess ||
          cudaMemcpyAsync(c->h_stats, c->d_stats, sizeof(rst_stats) * n_pairs, cudaMemcpyDeviceToHost, c->stream) != cudaSuccess)
        rc = RST_ERR_CUDA;
    }
    cudaGraph_t graph = nullptr;
    const cudaError_t e_end = cudaStreamEndCapture(c->stream, &graph);
    const int launches = (int)(c->launches - launches0);
    c->launches = launches0;
    if (rc != RST_OK || e_end != cudaSuccess || !graph) {
      if (graph) cudaGraphDestroy(graph);
      cudaGetLastError();
      return RST_OK;   // not captured: the direct path reports any real error
    }
    cudaGraphExec_t exec = nullptr;
    const cudaError_t e_inst = cudaGraphInstantiate(&exec, graph, 0);
    cudaGraphDestroy(graph);
    if (e_inst != cudaSuccess) { cudaGetLastError(); return RST_OK; }
    if (c->graphs.size() >= 8) {   // drop the least recently used
      size_t lru = 0;
      for (size_t i = 1; i < c->graphs.size(); ++i) if (c->graphs[i].stamp < c->graphs[lru].stamp) lru = i;
      cudaGraphExecDestroy(c->graphs[lru].exec);
      c->graphs.erase(c->graphs.begin() + lru);
    }
    c->graphs.push_back({key, exec, launches, 0});
    hit = &c->graphs.back();
  }
  hit->stamp = ++c->graph_clock;
  RST_CUDA(c, cudaGraphLaunch(hit->exec, c->stream));
  c->launches += hit->launches;
  c->fetch_pending = n_pairs;
  *ran = true;
  return RST_OK;
}

int32_t rst_set_graph_max_pairs(rst_ctx* c, int32_t max_pairs) {
  if (!c) return RST_ERR_INVALID_ARG;
  c->graph_max_pairs = max_pairs > 0 ? max_pairs : 0;
  return RST_OK;
}

int32_t rst_align_slots(rst_ctx* c, const int32_t* src_slots, const int32_t* dst_slots, int32_t n_pairs,
                        float* poses_inout, rst_stats* stats_out) {
  if (!c) return RST_ERR_INVALID_ARG;
  if (!c->begun) return fail(c, RST_ERR_INVALID_ARG, "rst_begin has not been called");
  if (!src_slots || !dst_slots || n_pairs < 0) return fail(c, RST_ERR_INVALID_ARG, "bad slot arrays / n_pairs");
  if (n_pairs > c->max_pairs) return fail(c, RST_ERR_CAPACITY, "more pairs than the context holds");
  c->n_pairs_last = n_pairs;
  if (n_pairs == 0) return RST_OK;
  RST_CUDA(c, cudaSetDevice(c->device));
  int32_t rc;
  if ((rc = pairs_begin(c, src_slots, dst_slots, n_pairs, poses_inout)) != RST_OK) return rc;
  if ((rc = pairs_iterate_split(c, n_pairs)) != RST_OK) return rc;
  return pairs_fetch(c, n_pairs, poses_inout, stats_out);
}

int32_t rst_set_stream_split(rst_ctx* c, int32_t min_pairs) {
  if (!c) return RST_ERR_INVALID_ARG;
  c->split_min_pairs = min_pairs > 0 ? min_pairs : 0x7fffffff;
  return RST_OK;
}

int32_t rst_set_schedule(rst_ctx* c, int32_t schedule) {
  if (!c) return RST_ERR_INVALID_ARG;
  if (schedule < RST_SCHEDULE_AUTO || schedule > RST_SCHEDULE_HYBRID) return fail(c, RST_ERR_INVALID_ARG, "unknown schedule");
  c->schedule = schedule;
  return RST_OK;
}

int32_t rst_set_cluster_size(rst_ctx* c, int32_t tiling, int32_t ctas_per_pair) {
  if (!c) return RST_ERR_INVALID_ARG;
  if (tiling < 0 || tiling > 1 || ctas_per_pair < 1 || ctas_per_pair > 16) return fail(c, RST_ERR_INVALID_ARG, "tiling must be 0/1, ctas_per_pair 1..16");
  RST_CUDA(c, cudaSetDevice(c->device));
  if (fused_max_active_clusters(ctas_per_pair) < 1) return fail(c, RST_ERR_INVALID_ARG, "this device cannot co-schedule a cluster of that size");
  c->cluster_size[tiling] = ctas_per_pair;
  return RST_OK;
}

int32_t rst_max_active_clusters(rst_ctx* c, int32_t ctas_per_pair) {
  if (!c || ctas_per_pair < 1 || ctas_per_pair > 16) return 0;
  if (cudaSetDevice(c->device) != cudaSuccess) return 0;
  return fused_max_active_clusters(ctas_per_pair);
}

int32_t rst_set_pipeline_chunk(rst_ctx* c, int32_t frames_per_chunk) {
  if (!c) return RST_ERR_INVALID_ARG;
  c->pipeline_chunk = frames_per_chunk;
  return RST_OK;
}

int32_t rst_device_results(rst_ctx* c, const float** d_poses, const rst_stats** d_stats) {
  if (!c) return RST_ERR_INVALID_ARG;
  if (d_poses) *d_poses = c->d_poses_cm;
  if (d_stats) *d_stats = c->d_stats;
  return RST_OK;
}

int32_t rst_copy_results_device(rst_ctx* c, float* d_poses_out, rst_stats* d_stats_out) {
  if (!c) return RST_ERR_INVALID_ARG;
  RST_CUDA(c, cudaSetDevice(c->device));
  const int n = c->n_pairs_last;
  if (n <= 0) return RST_OK;
  if (d_poses_out)
    RST_CUDA(c, cudaMemcpyAsync(d_poses_out, c->d_poses_cm, sizeof(float) * 16 * n, cudaMemcpyDeviceToDevice, c->stream));
  if (d_stats_out)
    RST_CUDA(c, cudaMemcpyAsync(d_stats_out, c->d_stats, sizeof(rst_stats) * n, cudaMemcpyDeviceToDevice, c->stream));
  return RST_OK;
}

int32_t rst_profile_enable(rst_ctx* c, int32_t on) {
  if (!c) return RST_ERR_INVALID_ARG;
  RST_CUDA(c, cudaSetDevice(c->device));
  RST_CUDA(c, cudaStreamSynchronize(c->stream));
  for (auto& r : c->prof_open) { c->ev_pool.push_back(r.e0); c->ev_pool.push_back(r.e1); }
  c->prof_open.clear();
  std::memset(&c->prof, 0, sizeof(c->prof));
  c->profiling = on != 0;
  return RST_OK;
}

int32_t rst_profile_read(rst_ctx* c, rst_profile* out) {
  if (!c || !out) return RST_ERR_INVALID_ARG;
  RST_CUDA(c, cudaSetDevice(c->device));
  RST_CUDA(c, cudaStreamSynchronize(c->stream));
  for (auto& r : c->prof_open) {
    float ms = 0.f;
    RST_CUDA(c, cudaEventElapsedTime(&ms, r.e0, r.e1));
    if (r.kind == 0) { c->prof.ms_preprocess[r.level] += ms; c->prof.launches_preprocess[r.level] += r.launches; c->prof.frames_preprocessed[r.level] += r.units; }
    else if (r.kind == 2) { c->prof.ms_icp_fused += ms; c->prof.launches_icp_fused += r.launches; c->prof.pair_iterations_fused += r.units; }
    else { c->prof.ms_icp[r.level] += ms; c->prof.launches_icp[r.level] += r.launches; c->prof.pairs_iterated[r.level] += r.units; }
    c->ev_pool.push_back(r.e0); c->ev_pool.push_back(r.e1);
  }
  c->prof_open.clear();
  *out = c->prof;
  return RST_OK;
}

int32_t rst_sync(rst_ctx* c) {
  if (!c) return RST_ERR_INVALID_ARG;
  RST_CUDA(c, cudaSetDevice(c->device));
  RST_CUDA(c, cudaStreamSynchronize(c->stream));
  return RST_OK;
}

static int32_t check_frames(rst_ctx* c, const rst_frame* f, int n) {
  for (int i = 0; i < n; ++i)
    if (f[i].width != f[0].width || f[i].height != f[0].height) return fail(c, RST_ERR_INVALID_ARG, "frames differ in size");
  return RST_OK;
}

/* ---- chunked host pipeline --------------------------------------------------------------------
 * The frames of one call are cut into chunks. Chunk k is copied on the copy stream; its kernels run
 * on work stream k % 2, so the H2D of chunk k+1 AND the latency-bound coarse-level launches of one
 * chunk overlap the bandwidth-bound fine-level launches of the other (a single small chunk cannot
 * fill 148 SMs). Results never depend on the chunking: every pair is reduced in blocks fixed by the
 * image size. */
static int32_t align_pairs_impl(rst_ctx* c, const rst_frame* src, const rst_frame* dst, int32_t n_pairs,
                                const rst_intrinsics* intr, const rst_params* params, float* poses_inout,
                                rst_stats* stats_out, bool wait) {
  if (!c) return RST_ERR_INVALID_ARG;
  if (!src || !dst || n_pairs < 0 || (wait && !poses_inout)) return fail(c, RST_ERR_INVALID_ARG, "null src/dst/poses or negative n_pairs");
  if (n_pairs == 0) return RST_OK;
  if (check_idle(c) != RST_OK) return RST_ERR_INVALID_ARG;
  if (2 * (int64_t)n_pairs > c->max_frames || n_pairs > c->max_pairs) return fail(c, RST_ERR_CAPACITY, "batch exceeds the context capacity");
  int32_t rc;
  if ((rc = check_frames(c, src, n_pairs)) != RST_OK) return rc;
  if ((rc = rst_begin(c, src[0].width, src[0].height, intr, params)) != RST_OK) return rc;
  // dst frames in slots [0, n), src frames in [n, 2n)
  std::vector<int32_t> s(n_pairs), d(n_pairs);
  for (int i = 0; i < n_pairs; ++i) { d[i] = i; s[i] = n_pairs + i; }
  // automatic: a BLOCKING call on a large batch runs as two halves, the upload of the second under the kernels of the
  // first (36.5 k -> 40.6 k pairs/s at 128 pairs of 640x480); an asynchronous call is not chunked — its caller overlaps
  // whole calls on two contexts, and half batches only cost partial waves there
  int chunk_cfg = c->pipeline_chunk;
  if (chunk_cfg == 0) chunk_cfg = (wait && n_pairs >= 32 && !c->profiling) ? (n_pairs + 1) / 2 : -1;
  const int chunk = chunk_cfg > 0 && chunk_cfg < n_pairs ? chunk_cfg : n_pairs;
  const bool piped = chunk < n_pairs;
  const bool try_graph = !piped && n_pairs <= c->graph_max_pairs && !c->profiling;
  if ((rc = pairs_begin(c, s.data(), d.data(), n_pairs, poses_inout, !try_graph)) != RST_OK) return rc;
  if (try_graph) {
    if ((rc = rst_upload_frames(c, dst, n_pairs, 0)) != RST_OK) return rc;
    if ((rc = rst_upload_frames(c, src, n_pairs, n_pairs)) != RST_OK) return rc;
    bool ran = false;
    if ((rc = run_graphed(c, 1, 2 * n_pairs, n_pairs, &ran)) != RST_OK) return rc;
    if (ran) return wait ? fetch_finish(c, poses_inout, stats_out) : RST_OK;
    if ((rc = pairs_begin_enqueue(c, n_pairs)) != RST_OK) return rc;   // graphs unavailable: the direct path, frames already uploaded
    if ((rc = preprocess_impl(c, 0, n_pairs, true)) != RST_OK) return rc;
    if ((rc = preprocess_impl(c, n_pairs, n_pairs, c->P.normal_cos_min > -1.0f)) != RST_OK) return rc;
    if ((rc = pairs_iterate_split(c, n_pairs)) != RST_OK) return rc;
    if (!wait) return fetch_enqueue(c, n_pairs);
    return pairs_fetch(c, n_pairs, poses_inout, stats_out);
  }
  const bool ngate = c->P.normal_cos_min > -1.0f;
  cudaStream_t main_s = c->stream;
  if (piped) {
    if ((rc = link_streams(c, main_s, c->copy_stream)) != RST_OK) return rc;
    for (int k = 0; k < 2; ++k) if ((rc = link_streams(c, main_s, c->work_stream[k])) != RST_OK) return rc;
  }
  int ci = 0;
  for (int p0 = 0; p0 < n_pairs; p0 += chunk, ++ci) {
    const int n = n_pairs - p0 < chunk ? n_pairs - p0 : chunk;
    cudaStream_t ws = piped ? c->work_stream[ci & 1] : main_s;
    {
      StreamScope up(c, piped ? c->copy_stream : main_s);
      if ((rc = rst_upload_frames(c, dst + p0, n, p0)) != RST_OK) return rc;
      if ((rc = rst_upload_frames(c, src + p0, n, n_pairs + p0)) != RST_OK) return rc;
    }
    if (piped && (rc = link_streams(c, c->copy_stream, ws)) != RST_OK) return rc;
    StreamScope work(c, ws);
    if ((rc = preprocess_impl(c, p0, n, true)) != RST_OK) return rc;
    if ((rc = preprocess_impl(c, n_pairs + p0, n, ngate)) != RST_OK) return rc;
    if ((rc = piped ? pairs_iterate(c, p0, n) : pairs_iterate_split(c, n_pairs)) != RST_OK) return rc;
  }
  if (piped)
    for (int k = 0; k < 2; ++k) if ((rc = link_streams(c, c->work_stream[k], main_s)) != RST_OK) return rc;
  if (!wait) return fetch_enqueue(c, n_pairs);
  return pairs_fetch(c, n_pairs, poses_inout, stats_out);
}

int32_t rst_align_pairs(rst_ctx* c, const rst_frame* src, const rst_frame* dst, int32_t n_pairs,
                        const rst_intrinsics* intr, const rst_params* params, float* poses_inout,
                        rst_stats* stats_out) {
  return align_pairs_impl(c, src, dst, n_pairs, intr, params, poses_inout, stats_out, true);
}

int32_t rst_align_pairs_async(rst_ctx* c, const rst_frame* src, const rst_frame* dst, int32_t n_pairs,
                              const rst_intrinsics* intr, const rst_params* params, const float* poses_in) {
  return align_pairs_impl(c, src, dst, n_pairs, intr, params, const_cast<float*>(poses_in), nullptr, false);
}

static int32_t align_sequence_impl(rst_ctx* c, const rst_frame* frames, int32_t n_frames, const rst_intrinsics* intr,
                                   const rst_params* params, float* poses_inout, rst_stats* stats_out, bool wait) {
  if (!c) return RST_ERR_INVALID_ARG;
  if (!frames || n_frames < 0 || (wait && !poses_inout)) return fail(c, RST_ERR_INVALID_ARG, "null frames/poses or negative n_frames");
  if (n_frames < 2) return RST_OK;
  if (check_idle(c) != RST_OK) return RST_ERR_INVALID_ARG;
  if (n_frames > c->max_frames || n_frames - 1 > c->max_pairs) return fail(c, RST_ERR_CAPACITY, "sequence exceeds the context capacity");
  int32_t rc;
  if ((rc = check_frames(c, frames, n_frames)) != RST_OK) return rc;
  if ((rc = rst_begin(c, frames[0].width, frames[0].height, intr, params)) != RST_OK) return rc;
  const int n_pairs = n_frames - 1;
  std::vector<int32_t> s(n_pairs), d(n_pairs);
  for (int i = 0; i < n_pairs; ++i) { s[i] = i + 1; d[i] = i; }  // AlignIcp3d(curr, prev): rs_replay_app.cpp:251
  int chunk_cfg = c->pipeline_chunk;   // see align_pairs_impl
  if (chunk_cfg == 0) chunk_cfg = (wait && n_frames >= 64 && !c->profiling) ? (n_frames + 1) / 2 : -1;
  const int chunk = chunk_cfg > 0 && chunk_cfg < n_frames ? chunk_cfg : n_frames;
  const bool piped = chunk < n_frames;
  const bool try_graph = !piped && n_pairs <= c->graph_max_pairs && !c->profiling;
  if ((rc = pairs_begin(c, s.data(), d.data(), n_pairs, poses_inout, !try_graph)) != RST_OK) return rc;
  if (try_graph) {
    if ((rc = rst_upload_frames(c, frames, n_frames, 0)) != RST_OK) return rc;
    bool ran = false;
    if ((rc = run_graphed(c, 0, n_frames, n_pairs, &ran)) != RST_OK) return rc;
    if (ran) return wait ? fetch_finish(c, poses_inout, stats_out) : RST_OK;
    if ((rc = pairs_begin_enqueue(c, n_pairs)) != RST_OK) return rc;   // graphs unavailable: the direct path, frames already uploaded
    if ((rc = preprocess_impl(c, 0, n_frames, true)) != RST_OK) return rc;
    if ((rc = pairs_iterate_split(c, n_pairs)) != RST_OK) return rc;
    if (!wait) return fetch_enqueue(c, n_pairs);
    return pairs_fetch(c, n_pairs, poses_inout, stats_out);
  }
  cudaStream_t main_s = c->stream;
  if (piped) {
    if ((rc = link_streams(c, main_s, c->copy_stream)) != RST_OK) return rc;
    for (int k = 0; k < 2; ++k) if ((rc = link_streams(c, main_s, c->work_stream[k])) != RST_OK) return rc;
  }
  int ci = 0;
  for (int f0 = 0; f0 < n_frames; f0 += chunk, ++ci) {
    const int n = n_frames - f0 < chunk ? n_frames - f0 : chunk;
    cudaStream_t ws = piped ? c->work_stream[ci & 1] : main_s;
    {
      StreamScope up(c, piped ? c->copy_stream : main_s);
      if ((rc = rst_upload_frames(c, frames + f0, n, f0)) != RST_OK) return rc;
    }
    if (piped && (rc = link_streams(c, c->copy_stream, ws)) != RST_OK) return rc;
    StreamScope work(c, ws);
    if ((rc = preprocess_impl(c, f0, n, true)) != RST_OK) return rc;
    // pair i uses frames i and i+1: the first pair of this chunk also needs the last frame of the previous
    // chunk, pre-processed on the other work stream
    const int p0 = f0 > 0 ? f0 - 1 : 0, p1 = f0 + n - 1;
    if (piped && ci > 0 && (rc = link_streams(c, c->work_stream[(ci - 1) & 1], ws)) != RST_OK) return rc;
    if ((rc = piped ? pairs_iterate(c, p0, p1 - p0) : pairs_iterate_split(c, n_pairs)) != RST_OK) return rc;
  }
  if (piped)
    for (int k = 0; k < 2; ++k) if ((rc = link_streams(c, c->work_stream[k], main_s)) != RST_OK) return rc;
  if (!wait) return fetch_enqueue(c, n_pairs);
  return pairs_fetch(c, n_pairs, poses_inout, stats_out);
}

int32_t rst_align_sequence(rst_ctx* c, const rst_frame* frames, int32_t n_frames, const rst_intrinsics* intr,
                           const rst_params* params, float* poses_inout, rst_stats* stats_out) {
  return align_sequence_impl(c, frames, n_frames, intr, params, poses_inout, stats_out, true);
}

int32_t rst_align_sequence_async(rst_ctx* c, const rst_frame* frames, int32_t n_frames, const rst_intrinsics* intr,
                                 const rst_params* params, const float* poses_in) {
  return align_sequence_impl(c, frames, n_frames, intr, params, const_cast<float*>(poses_in), nullptr, false);
}

int32_t rst_wait(rst_ctx* c, float* poses_out, rst_stats* stats_out) {
  if (!c) return RST_ERR_INVALID_ARG;
  RST_CUDA(c, cudaSetDevice(c->device));
  return fetch_finish(c, poses_out, stats_out);
}

int32_t rst_level_info(const rst_ctx* c, int32_t level, int32_t* width, int32_t* height, int32_t* pitch_px,
                       rst_intrinsics* intr) {
  if (!c || !c->begun || level < 0 || level >= c->num_levels) return RST_ERR_INVALID_ARG;
  if (width) *width = c->geom[level].w;
  if (height) *height = c->geom[level].h;
  if (pitch_px) *pitch_px = (level == 0 && c->ext0) ? c->ext_pitch0 : c->pitch[level];
  if (intr) { intr->fx = c->geom[level].fx; intr->fy = c->geom[level].fy; intr->cx = c->geom[level].cx; intr->cy = c->geom[level].cy; }
  return RST_OK;
}

int32_t rst_read_depth(rst_ctx* c, int32_t slot, int32_t level, uint16_t* out) {
  if (!c) return RST_ERR_INVALID_ARG;
  if (!c->begun || !out || level < 0 || level >= c->num_levels || slot < 0 || slot >= c->max_frames)
    return fail(c, RST_ERR_INVALID_ARG, "bad slot/level/out");
  RST_CUDA(c, cudaSetDevice(c->device));
  const LevelStore s = level_store(c, level);
  const int w = c->geom[level].w, h = c->geom[level].h;
  RST_CUDA(c, cudaMemcpy2DAsync(out, (size_t)w * 2, s.depth + (int64_t)slot * s.depth_frame, (size_t)s.depth_pitch * 2,
                                (size_t)w * 2, h, cudaMemcpyDeviceToHost, c->stream));
  RST_CUDA(c, cudaStreamSynchronize(c->stream));
  return RST_OK;
}

int32_t rst_read_geometry(rst_ctx* c, int32_t slot, int32_t level, float* out) {
  if (!c) return RST_ERR_INVALID_ARG;
  if (!c->begun || !out || level < 0 || level >= c->num_levels || slot < 0 || slot >= c->max_frames)
    return fail(c, RST_ERR_INVALID_ARG, "bad slot/level/out");
  RST_CUDA(c, cudaSetDevice(c->device));
  const size_t n = (size_t)c->geom[level].w * c->geom[level].h;
  RST_CUDA(c, cudaMemcpyAsync(out, c->d_geom[level] + kGeomGuard + (int64_t)slot * c->gframe[level], n * sizeof(float4),
                              cudaMemcpyDeviceToHost, c->stream));
  RST_CUDA(c, cudaStreamSynchronize(c->stream));
  return RST_OK;
}

int32_t rst_read_intensity(rst_ctx* c, int32_t slot, int32_t level, float* out) {
  if (!c) return RST_ERR_INVALID_ARG;
  if (!c->begun || !c->photo || !out || level < 0 || level >= c->num_levels || slot < 0 || slot >= c->max_frames)
    return fail(c, RST_ERR_INVALID_ARG, "bad slot/level/out, or the photometric term is off");
  RST_CUDA(c, cudaSetDevice(c->device));
  const size_t n = (size_t)c->geom[level].w * c->geom[level].h;
  RST_CUDA(c, cudaMemcpyAsync(out, c->d_int[level] + n * (size_t)slot, n * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
  RST_CUDA(c, cudaStreamSynchronize(c->stream));
  return RST_OK;
}

int32_t rst_evaluate(rst_ctx* c, int32_t src_slot, int32_t dst_slot, int32_t level, const float* pose,
                     int32_t* idx_out, rst_stats* stats_out) {
  if (!c) return RST_ERR_INVALID_ARG;
  if (!c->begun || !pose || !stats_out || level < 0 || level >= c->num_levels) return fail(c, RST_ERR_INVALID_ARG, "bad level/pose/stats");
  if (!slot_bound(c, src_slot) || !slot_bound(c, dst_slot))
    return fail(c, RST_ERR_INVALID_ARG, "slot index out of range (or outside the frames bound by rst_set_frames_device)");
  if (check_idle(c) != RST_OK) return RST_ERR_INVALID_ARG;
  RST_CUDA(c, cudaSetDevice(c->device));
  const int sp = c->max_pairs;  // scratch pair
  c->h_pairs[sp] = make_int2(src_slot, dst_slot);
  std::memcpy(c->h_poses + 16 * sp, pose, sizeof(float) * 16);
  RST_CUDA(c, cudaMemcpyAsync(c->d_pairs + sp, c->h_pairs + sp, sizeof(int2), cudaMemcpyHostToDevice, c->stream));
  RST_CUDA(c, cudaMemcpyAsync(c->d_poses_in + 16 * sp, c->h_poses + 16 * sp, sizeof(float) * 16, cudaMemcpyHostToDevice, c->stream));
  InitArgs ia{c->d_poses_in + 16 * sp, c->d_master + 12 * sp, c->d_pose_f32 + 12 * sp, c->d_poses_cm + 16 * sp,
              c->d_stats + sp, c->d_tickets + sp, 1};
  RST_CUDA(c, launch_init_pairs(ia, c->stream));
  c->launches += 1;
  const size_t npx = (size_t)c->geom[level].w * c->geom[level].h;
  if (idx_out && c->idx_bytes < npx * 4) {
    cudaFree(c->d_idx); c->d_idx = nullptr; c->idx_bytes = 0;
    RST_CUDA(c, cudaMalloc(&c->d_idx, (size_t)c->max_w * c->max_h * 4));
    c->idx_bytes = (size_t)c->max_w * c->max_h * 4;
  }
  IcpArgs a{};
  fill_icp_args(c, level, &a);
  a.pair_offset = sp;
  a.update_pose = 0;
  a.done = nullptr;
  a.idx_out = idx_out ? c->d_idx : nullptr;
  RST_CUDA(c, launch_icp_iter(a, 1, c->P.robust_kind, c->P.normal_cos_min > -1.0f, idx_out != nullptr, c->photo, c->stream));
  c->launches += 1;
  RST_CUDA(c, cudaMemcpyAsync(c->h_stats + sp, c->d_stats + sp, sizeof(rst_stats), cudaMemcpyDeviceToHost, c->stream));
  if (idx_out) RST_CUDA(c, cudaMemcpyAsync(idx_out, c->d_idx, npx * 4, cudaMemcpyDeviceToHost, c->stream));
  RST_CUDA(c, cudaStreamSynchronize(c->stream));
  *stats_out = c->h_stats[sp];
  return RST_OK;
}

}  // extern "C"
