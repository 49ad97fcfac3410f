/*
 * rst_icp3d.cu — the reference's own cloud-based ICP on the GPU.
 *
 * Literal counterpart of AlignIcp3d (rs_tracker/align/src/align_icp.cpp:73-167): exact nearest
 * neighbour + Geman-McClure/GNC weights + weighted cross-covariance + 3x3 SVD (Kabsch) + quaternion
 * round trip, max_iter fixed iterations. One thread block of 1024 threads owns one pair and runs ALL
 * iterations inside one launch (the pose lives in shared memory, phases are separated by
 * __syncthreads), so a batch of pairs is one kernel launch and there is no host round trip per
 * iteration. The KD-tree (nanoflann, kdtree.hpp:27-57) is replaced by a uniform grid over the dst
 * cloud built in the same kernel, searched in growing rings until the ring bound proves the current
 * best is the exact nearest neighbour (ties: lowest index). From the second iteration on a query searches only
 * the ball through its previous neighbour — and not at all while the triangle inequality proves that neighbour
 * cannot have changed (the neighbour cache, see nn_ball): the result of every query is the one a search would
 * return, index and fp32 distance, bit for bit. A small batch gives every pair a thread-block cluster instead of
 * one CTA (sums through distributed shared memory).
 *
 * fp32 arithmetic that decides a neighbour (transform, squared distance) uses explicit
 * round-to-nearest intrinsics in the reference's operation order (no FMA contraction), so NN indices
 * and weights are bit-identical to the CPU restatement for the same input pose.
 */
#include <cooperative_groups.h>
#include <cuda_runtime.h>

#include <cfloat>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "rst_align.h"
#include "rst_device.cuh"
#include "rst_internal.h"

// -DRST_ICP3D_PROFILE: block 0 prints the cycles thread 0 spent in every phase of the iteration loop (experiments only)
#ifdef RST_ICP3D_PROFILE
#define PHASE(k) do { const long long now_ = clock64(); ph_[k] += now_ - t_; t_ = now_; } while (0)
#else
#define PHASE(k) do { } while (0)
#endif

// -DRST_ICP3D_CHECK: index checks in k_icp3d that trap (compute-sanitizer is not available on the GPU pool; the tests are
// run once against a library built this way)
#ifdef RST_ICP3D_CHECK
#define ICP3D_CHECK(cond) do { if (!(cond)) { printf("k_icp3d check failed: %s (line %d)\n", #cond, __LINE__); __trap(); } } while (0)
#else
#define ICP3D_CHECK(cond) do { } while (0)
#endif

namespace {

constexpr int kThreads = 1024;
constexpr int kWarps = kThreads / 32;
constexpr int kCellCap = 1 << 18;  // grid cells per pair (1 MiB of cell_start)
constexpr int kRingBytes = 8 * kThreads * 16;   // dynamic shared memory of k_icp3d (OwnStream)
constexpr int kUploadChunks = 4;   // rst_icp3d_depth: frame upload / depth -> cloud pipeline depth
constexpr int kGroupScanMax = 2048;   // queued points per CTA up to which a search is spread over kScanLanes lanes
constexpr float kCacheGain = 1.0f, kCacheLo = 0.05f, kCacheHi = 0.2f;   // neighbour-cache scan margin (see nn_ball)

struct PairDesc {
  const float* src;   // n x 3
  const float* dst;   // m x 3
  int n, m;           // point counts, or ...
  const int* n_ptr;   // ... read from device memory when the clouds were produced on the device
  const int* m_ptr;
  int* cell_start;    // [kCellCap + 1]
  int* cell_fill;     // [kCellCap]
  float4* sorted;     // m: x, y, z, original index (bit pattern)
  int* nbr;           // n
  float* w;           // n
  float4* sl;         // n   (source point, neighbour cache: proven radius L | iteration tag)
  float4* qd;         // n   (coordinates of the current neighbour, squared distance to it)
  int* queue;         // n + 16 * kThreads: source points whose neighbour has to be searched this iteration
  unsigned long long* stat;   // [4]: neighbour searches done, neighbour queries answered, iterations run, iterations asked (zeroed by the host)
  float* pose;        // 16, column-major, in/out
  rst_icp3d_result* res;
};

struct Grid {
  float lox, loy, loz, h, inv_h;
  int nx, ny, nz;
};

__device__ __forceinline__ float mulrn(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float addrn(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float subrn(float a, float b) { return __fsub_rn(a, b); }

// hi + lo += x without rounding error (Knuth's TwoSum; adds only, so nothing for the compiler to contract). A thread's
// partial sum kept this way carries ~48 significant bits — the fp64 accumulation of fp32 terms the reference
// specifies (align_icp.cpp:124-136) to within 1e-14 — without a float -> double conversion and a DADD per term
// (those two made the covariance pass the longest phase of an iteration: 13 issue cycles per term on B200).
__device__ __forceinline__ void acc2(float& hi, float& lo, float x) {
  const float s = __fadd_rn(hi, x);
  const float bb = __fsub_rn(s, hi);
  lo = __fadd_rn(lo, __fadd_rn(__fsub_rn(hi, __fsub_rn(s, bb)), __fsub_rn(x, bb)));
  hi = s;
}

// block-wide sum of K doubles, fixed order. Inside a warp a transposed butterfly: at every step a lane passes one half
// of the components it still holds to its partner and adds the partner's other half to its own, so that K (padded to a
// power of two) components over 32 lanes cost about K shuffles instead of 5 K — shuffles were what a 1024-thread CTA
// spent most of a reduction on. Then warp k adds the 32 warp sums of component k (xor tree).
template <int K, bool PUSH = false>
__device__ void block_sum(double (&v)[K], double (*s_part)[16], double* s_out) {
  constexpr int KP = K <= 1 ? 1 : K <= 2 ? 2 : K <= 4 ? 4 : K <= 8 ? 8 : 16;
  static_assert(K <= 16 && kWarps == 32, "block_sum: at most 16 components, 32 warps");
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double x[KP];
#pragma unroll
  for (int k = 0; k < KP; ++k) x[k] = k < K ? v[k] : 0.0;
  int comp = 0;   // the component this lane ends up with
  int o = 16;
#pragma unroll
  for (int cnt = KP; cnt > 1; cnt >>= 1, o >>= 1) {
    const bool upper = (lane & o) != 0;
    comp = 2 * comp + (upper ? 1 : 0);
#pragma unroll
    for (int k = 0; k < cnt / 2; ++k) {
      const double send = upper ? x[k] : x[k + cnt / 2], keep = upper ? x[k + cnt / 2] : x[k];
      x[k] = keep + __shfl_xor_sync(0xffffffffu, send, o);
    }
  }
  for (; o > 0; o >>= 1) x[0] += __shfl_xor_sync(0xffffffffu, x[0], o);
  // lanes that differ only in the bits below the last exchange hold the same sum: the lowest of them stores it
  constexpr int kSteps = KP == 1 ? 0 : KP == 2 ? 1 : KP == 4 ? 2 : KP == 8 ? 3 : 4;
  if ((lane & ((32 >> kSteps) - 1)) == 0 && comp < K) s_part[warp][comp] = x[0];
  __syncthreads();
  if (warp < K) {   // warp k adds the 32 warp sums of component k
    double y = s_part[lane][warp];
#pragma unroll
    for (int q = 16; q > 0; q >>= 1) y += __shfl_xor_sync(0xffffffffu, y, q);
    if (!PUSH) { if (lane == 0) s_out[warp] = y; }
    else {
      // s_out = this CTA's row of the cluster's table [rank][16]: lane r stores the sum into the same row of rank r's table
      // (distributed shared memory); the caller's cluster barrier publishes it
      namespace cg = cooperative_groups;
      cg::cluster_group cluster = cg::this_cluster();
      if (lane < (int)cluster.num_blocks()) cluster.map_shared_rank(s_out, lane)[warp] = y;
    }
  }
  if (!PUSH) __syncthreads();
}

__device__ __forceinline__ int cell_coord(float p, float lo, float inv_h, int n) {
  int c = (int)floorf((p - lo) * inv_h);
  return c < 0 ? 0 : (c >= n ? n - 1 : c);
}

// Builds the uniform grid of a cloud inside one 1024-thread block: bounding box, cell size (grid_cell, or
// extent/64; grown until the cell count fits kCellCap), counting sort of the points into cells
// (cell_start[cells + 1], sorted[] = {x, y, z, original index}). All threads of the block must call it.
__device__ Grid build_grid(const float* __restrict__ pts, int n, float grid_cell, int* cell_start, int* cell_fill, float4* sorted) {
  __shared__ float s_lohi[kWarps][6];
  __shared__ int s_scan[kWarps];
  __shared__ Grid s_grid;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  float lo[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, hi[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
  for (int j = tid; j < n; j += kThreads)
    for (int a = 0; a < 3; ++a) {   // the box of the FINITE coordinates: an infinite one would make the cell counts overflow
      const float v = pts[3 * j + a];
      if (fabsf(v) <= FLT_MAX) { lo[a] = fminf(lo[a], v); hi[a] = fmaxf(hi[a], v); }
    }
  for (int a = 0; a < 3; ++a)
    for (int o = 16; o > 0; o >>= 1) {
      lo[a] = fminf(lo[a], __shfl_xor_sync(0xffffffffu, lo[a], o));
      hi[a] = fmaxf(hi[a], __shfl_xor_sync(0xffffffffu, hi[a], o));
    }
  if (lane == 0) for (int a = 0; a < 3; ++a) { s_lohi[warp][a] = lo[a]; s_lohi[warp][3 + a] = hi[a]; }
  __syncthreads();
  if (tid == 0) {
    float bl[3], bh[3];
    for (int a = 0; a < 3; ++a) {
      bl[a] = FLT_MAX; bh[a] = -FLT_MAX;
      for (int w = 0; w < kWarps; ++w) { bl[a] = fminf(bl[a], s_lohi[w][a]); bh[a] = fmaxf(bh[a], s_lohi[w][3 + a]); }
      if (!(bh[a] >= bl[a])) { bl[a] = 0.f; bh[a] = 0.f; }   // no finite coordinate on this axis: one cell
    }
    Grid g;
    g.lox = bl[0]; g.loy = bl[1]; g.loz = bl[2];
    const float ex = bh[0] - bl[0], ey = bh[1] - bl[1], ez = bh[2] - bl[2];
    float h = grid_cell > 0.f ? grid_cell : fmaxf(fmaxf(ex, fmaxf(ey, ez)) / 64.f, 1e-6f);
    for (;;) {
      // counted in floating point first: a tiny grid_cell must not overflow the int conversion
      const float fx = floorf(ex / h) + 1.f, fy = floorf(ey / h) + 1.f, fz = floorf(ez / h) + 1.f;
      if ((double)fx * (double)fy * (double)fz <= (double)kCellCap) { g.nx = (int)fx; g.ny = (int)fy; g.nz = (int)fz; break; }
      h *= 1.26f;
    }
    g.h = h; g.inv_h = 1.0f / h;
    s_grid = g;
  }
  __syncthreads();
  const Grid g = s_grid;
  const int n_cells = g.nx * g.ny * g.nz;
  auto cell_of = [&](int j) {
    return (cell_coord(pts[3 * j + 2], g.loz, g.inv_h, g.nz) * g.ny + cell_coord(pts[3 * j + 1], g.loy, g.inv_h, g.ny)) * g.nx +
           cell_coord(pts[3 * j], g.lox, g.inv_h, g.nx);
  };
  for (int c = tid; c < n_cells; c += kThreads) cell_fill[c] = 0;
  __syncthreads();
  for (int j = tid; j < n; j += kThreads) atomicAdd(cell_fill + cell_of(j), 1);
  __syncthreads();
  {  // exclusive scan of the cell counts in coalesced tiles of 4 cells per thread; cell_fill is reset on the way
    __shared__ int s_carry;
    if (tid == 0) s_carry = 0;
    __syncthreads();
    for (int base = 0; base < n_cells; base += 4 * kThreads) {
      const int c = base + 4 * tid;
      int k4[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) k4[q] = c + q < n_cells ? cell_fill[c + q] : 0;
      const int local = k4[0] + k4[1] + k4[2] + k4[3];
      int incl = local;
      for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += v; }
      if (lane == 31) s_scan[warp] = incl;
      __syncthreads();
      if (warp == 0) {
        int v = s_scan[lane];
        for (int o = 1; o < 32; o <<= 1) { const int u = __shfl_up_sync(0xffffffffu, v, o); if (lane >= o) v += u; }
        s_scan[lane] = v;
      }
      __syncthreads();
      int run = s_carry + incl - local + (warp > 0 ? s_scan[warp - 1] : 0);
#pragma unroll
      for (int q = 0; q < 4; ++q)
        if (c + q < n_cells) { cell_start[c + q] = run; cell_fill[c + q] = 0; run += k4[q]; }
      __syncthreads();
      if (tid == kThreads - 1) s_carry = run;
      __syncthreads();
    }
    if (tid == 0) cell_start[n_cells] = n;
  }
  __syncthreads();
  for (int j = tid; j < n; j += kThreads) {
    const int c = cell_of(j);
    const int pos = cell_start[c] + atomicAdd(cell_fill + c, 1);
    sorted[pos] = make_float4(pts[3 * j], pts[3 * j + 1], pts[3 * j + 2], __int_as_float(j));
  }
  __syncthreads();
  return g;
}

// exact nearest neighbour of p in the gridded dst cloud; ties go to the lowest original index
__device__ void nn_search(const Grid& g, const int* __restrict__ cell_start, const float4* __restrict__ sorted,
                          float px, float py, float pz, int* best_j, float* best_d2) {
  const int cx = cell_coord(px, g.lox, g.inv_h, g.nx), cy = cell_coord(py, g.loy, g.inv_h, g.ny),
            cz = cell_coord(pz, g.loz, g.inv_h, g.nz);
  float bd = FLT_MAX;
  int bj = 0;  // always a valid index: a query that beats nothing (non-finite coordinates) must not return INT_MAX
  for (int r = 0;; ++r) {
    const int x0 = max(cx - r, 0), x1 = min(cx + r, g.nx - 1);
    const int y0 = max(cy - r, 0), y1 = min(cy + r, g.ny - 1);
    const int z0 = max(cz - r, 0), z1 = min(cz + r, g.nz - 1);
    for (int z = z0; z <= z1; ++z)
      for (int y = y0; y <= y1; ++y) {
        const bool inner_zy = (abs(z - cz) < r) && (abs(y - cy) < r);
        for (int x = x0; x <= x1; ++x) {
          if (inner_zy && abs(x - cx) < r) { x = cx + r - 1; continue; }  // interior was visited by smaller rings
          const int c = (z * g.ny + y) * g.nx + x;
          const int e = cell_start[c + 1];
          for (int k = cell_start[c]; k < e; ++k) {
            const float4 q = sorted[k];
            const float dx = subrn(px, q.x), dy = subrn(py, q.y), dz = subrn(pz, q.z);
            const float d2 = addrn(addrn(mulrn(dx, dx), mulrn(dy, dy)), mulrn(dz, dz));  // left to right, as nanoflann L2
            const int j = __float_as_int(q.w);
            if (d2 < bd || (d2 == bd && j < bj)) { bd = d2; bj = j; }
          }
        }
      }
    if (x0 == 0 && x1 == g.nx - 1 && y0 == 0 && y1 == g.ny - 1 && z0 == 0 && z1 == g.nz - 1) break;  // whole grid seen
    // every unvisited point lies beyond one of the (unclipped) faces of the visited box
    float bound = FLT_MAX;
    if (cx - r > 0) bound = fminf(bound, px - (g.lox + (float)(cx - r) * g.h));
    if (cx + r < g.nx - 1) bound = fminf(bound, (g.lox + (float)(cx + r + 1) * g.h) - px);
    if (cy - r > 0) bound = fminf(bound, py - (g.loy + (float)(cy - r) * g.h));
    if (cy + r < g.ny - 1) bound = fminf(bound, (g.loy + (float)(cy + r + 1) * g.h) - py);
    if (cz - r > 0) bound = fminf(bound, pz - (g.loz + (float)(cz - r) * g.h));
    if (cz + r < g.nz - 1) bound = fminf(bound, (g.loz + (float)(cz + r + 1) * g.h) - pz);
    bound -= 1e-3f * g.h;  // slack for the rounding of the point -> cell assignment
    if (bound > 0.0f && bd <= bound * bound * 0.9999f) break;
  }
  *best_j = bj;
  *best_d2 = bd;
}

// Exact nearest neighbour when a candidate is already known (the previous iteration's neighbour):
// the true nearest neighbour — and every point tying with it — lies inside the ball of radius
// |p - candidate| around p, so only the cells that ball overlaps are scanned (typically 1-8 instead
// of the 27+ of a ring search). Same result as nn_search: smallest fp32 d2, ties to the lowest index.
//
// The ball is scanned `margin` wider than needed, and the scan also tracks the second-smallest distance it
// meets: afterwards every dst point other than the winner is PROVEN to lie at least L = min(second distance,
// scanned radius) away from p. The caller keeps (p, L) with the neighbour (the neighbour cache): while a later query
// p' of the same source point satisfies |p' - nbr| + |p' - p| < L, the triangle inequality makes nbr the strict
// nearest neighbour of p' as well, and the search is skipped — same index, same fp32 d2, bit for bit.
__device__ void nn_ball(const Grid& g, const int* __restrict__ cell_start, const float4* __restrict__ sorted,
                        const float* __restrict__ dst, float px, float py, float pz, int cand, float margin, int* best_j,
                        float* best_d2, float* proven) {
  float bd, sd = FLT_MAX;   // best and second-best squared distance (second: any index other than the best's)
  int bj = cand;
  {
    const float dx = subrn(px, dst[3 * cand]), dy = subrn(py, dst[3 * cand + 1]), dz = subrn(pz, dst[3 * cand + 2]);
    bd = addrn(addrn(mulrn(dx, dx), mulrn(dy, dy)), mulrn(dz, dz));
  }
  const float rad_sure = sqrtf(bd) * 1.0001f + margin;   // every point within rad_sure of p is examined
  const float rad = rad_sure + 1e-3f * g.h;               // inflated: rounding of d2 and of the cell assignment
  const int x0 = cell_coord(px - rad, g.lox, g.inv_h, g.nx), x1 = cell_coord(px + rad, g.lox, g.inv_h, g.nx);
  const int y0 = cell_coord(py - rad, g.loy, g.inv_h, g.ny), y1 = cell_coord(py + rad, g.loy, g.inv_h, g.ny);
  const int z0 = cell_coord(pz - rad, g.loz, g.inv_h, g.nz), z1 = cell_coord(pz + rad, g.loz, g.inv_h, g.nz);
  for (int z = z0; z <= z1; ++z)
    for (int y = y0; y <= y1; ++y) {
      const int row = (z * g.ny + y) * g.nx;
      // the cells x0..x1 of one row are contiguous in the sorted array: one range instead of x1-x0+1
      ICP3D_CHECK(row + x0 >= 0 && x0 <= x1 && row + x1 + 1 <= g.nx * g.ny * g.nz);
      const int e = cell_start[row + x1 + 1];
      for (int k = cell_start[row + x0]; k < e; ++k) {
        const float4 q = sorted[k];
        const float dx = subrn(px, q.x), dy = subrn(py, q.y), dz = subrn(pz, q.z);
        const float d2 = addrn(addrn(mulrn(dx, dx), mulrn(dy, dy)), mulrn(dz, dz));
        const int j = __float_as_int(q.w);
        if (d2 < bd || (d2 == bd && j < bj)) { sd = bd; bd = d2; bj = j; }   // the old best is the new runner-up
        else if (j != bj) sd = fminf(sd, d2);
      }
    }
  *best_j = bj;
  *best_d2 = bd;
  *proven = fminf(sqrtf(sd), rad_sure) * 0.9998f;
}

// nn_ball by kScanLanes consecutive lanes of a warp for ONE query (all of them pass the same arguments; `active` = the group
// has a query): the (z, y) rows of the ball's box are dealt out over the lanes, every lane keeps the best and second
// best of its rows, and a butterfly merges them — the latency of a search is one or two rows instead of all of them.
// Used when few points of a CTA have to search (most iterations): the whole CTA waits for them at a barrier. Every
// lane of the warp must call it (full-mask shuffles). Same results as nn_ball, bit for bit.
constexpr int kScanLanes = 8;
__device__ void nn_ball_group(const Grid& g, const int* __restrict__ cell_start, const float4* __restrict__ sorted,
                              const float* __restrict__ dst, bool active, float px, float py, float pz, int cand, float margin,
                              int* best_j, float* best_d2, float* proven) {
  float bd = FLT_MAX, sd = FLT_MAX, rad_sure = 0.f, dc2 = FLT_MAX;
  int bj = 0x7fffffff;
  if (active) {
    const float cx = subrn(px, dst[3 * cand]), cy = subrn(py, dst[3 * cand + 1]), cz = subrn(pz, dst[3 * cand + 2]);
    dc2 = addrn(addrn(mulrn(cx, cx), mulrn(cy, cy)), mulrn(cz, cz));
    rad_sure = sqrtf(dc2) * 1.0001f + margin;
    const float rad = rad_sure + 1e-3f * g.h;
    const int x0 = cell_coord(px - rad, g.lox, g.inv_h, g.nx), x1 = cell_coord(px + rad, g.lox, g.inv_h, g.nx);
    const int y0 = cell_coord(py - rad, g.loy, g.inv_h, g.ny), y1 = cell_coord(py + rad, g.loy, g.inv_h, g.ny);
    const int z0 = cell_coord(pz - rad, g.loz, g.inv_h, g.nz), z1 = cell_coord(pz + rad, g.loz, g.inv_h, g.nz);
    const int ny = y1 - y0 + 1, rows = ny * (z1 - z0 + 1);
    for (int r = threadIdx.x & (kScanLanes - 1); r < rows; r += kScanLanes) {
      const int row = ((z0 + r / ny) * g.ny + y0 + r % ny) * g.nx;
      ICP3D_CHECK(row + x0 >= 0 && x0 <= x1 && row + x1 + 1 <= g.nx * g.ny * g.nz);
      const int e = cell_start[row + x1 + 1];
      for (int k = cell_start[row + x0]; k < e; ++k) {
        const float4 q = sorted[k];
        const float dx = subrn(px, q.x), dy = subrn(py, q.y), dz = subrn(pz, q.z);
        const float d2 = addrn(addrn(mulrn(dx, dx), mulrn(dy, dy)), mulrn(dz, dz));
        const int j = __float_as_int(q.w);
        if (d2 < bd || (d2 == bd && j < bj)) { sd = bd; bd = d2; bj = j; }
        else sd = fminf(sd, d2);   // a lane meets every point once: j != bj here
      }
    }
  }
#pragma unroll
  for (int o = kScanLanes / 2; o > 0; o >>= 1) {   // the lanes' candidates are disjoint: the loser's best is a runner-up
    const float obd = __shfl_xor_sync(0xffffffffu, bd, o), osd = __shfl_xor_sync(0xffffffffu, sd, o);
    const int obj = __shfl_xor_sync(0xffffffffu, bj, o);
    const bool take = obd < bd || (obd == bd && obj < bj);
    sd = fminf(fminf(sd, osd), take ? bd : obd);
    if (take) { bd = obd; bj = obj; }
  }
  if (bj == 0x7fffffff) { bj = cand; bd = dc2; }   // nothing met (a non-finite candidate): as nn_ball, which starts from it
  *best_j = bj;
  *best_d2 = bd;
  *proven = fminf(sqrtf(sd), rad_sure) * 0.9998f;
}

// R = U V^T of a 3x3 fp64 matrix by one-sided Jacobi (row-major in/out). V (in/out, orthogonal) is the right basis the
// sweeps start from and the one they end with: the identity for a cold start, the previous ICP iteration's result for
// a warm one — consecutive cross-covariances differ little, so B = M V then has nearly orthogonal columns already and
// one or two sweeps finish the job instead of five. The rotation (c, s) that zeroes the column product gamma comes
// without a division or a square root: with tau = beta - alpha, kappa = 2 gamma, rho = |(tau, kappa)|,
//   c = (|tau| + rho) / sqrt(2 rho (rho + |tau|)),  s = sign(tau) kappa / sqrt(2 rho (rho + |tau|))
// (the same angle as t = sign(zeta) / (|zeta| + sqrt(1 + zeta^2)), zeta = tau / kappa), two reciprocal square roots.
__device__ bool svd_uvt(const double* M, double* UVt, double* V) {   // returns whether any rotation was applied (V changed)
  double B[9];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) B[3 * i + j] = M[3 * i] * V[j] + M[3 * i + 1] * V[3 + j] + M[3 * i + 2] * V[6 + j];
  bool any_rotation = false;
  for (int sweep = 0; sweep < 60; ++sweep) {
    bool rotated = false;
    for (int p = 0; p < 2; ++p)
      for (int q = p + 1; q < 3; ++q) {
        double a = 0, b = 0, c = 0;
        for (int i = 0; i < 3; ++i) { a += B[3 * i + p] * B[3 * i + p]; b += B[3 * i + q] * B[3 * i + q]; c += B[3 * i + p] * B[3 * i + q]; }
        if (!(c * c > 1e-30 * (a * b))) continue;   // |c| <= 1e-15 sqrt(a b): the two columns are orthogonal
        rotated = true;
        const double tau = b - a, kappa = 2.0 * c, s2 = tau * tau + kappa * kappa;
        const double rho = s2 * rsqrt(s2), at = fabs(tau);
        const double r = rsqrt(2.0 * rho * (rho + at));
        const double cs = (at + rho) * r, sn = (tau >= 0 ? kappa : -kappa) * r;
        for (int i = 0; i < 3; ++i) {
          const double bp = B[3 * i + p], bq = B[3 * i + q];
          B[3 * i + p] = cs * bp - sn * bq; B[3 * i + q] = sn * bp + cs * bq;
          const double vp = V[3 * i + p], vq = V[3 * i + q];
          V[3 * i + p] = cs * vp - sn * vq; V[3 * i + q] = sn * vp + cs * vq;
        }
      }
    if (!rotated) break;
    any_rotation = true;
  }
  double U[9], s[3], inv[3];
  for (int j = 0; j < 3; ++j) {
    const double n2 = B[j] * B[j] + B[3 + j] * B[3 + j] + B[6 + j] * B[6 + j];
    inv[j] = rsqrt(n2);
    s[j] = n2 > 0 ? n2 * inv[j] : 0.0;
  }
  const double smax = fmax(s[0], fmax(s[1], s[2]));
  int bad = -1;
  for (int j = 0; j < 3; ++j) {
    if (s[j] > 1e-14 * smax && s[j] > 0) { for (int i = 0; i < 3; ++i) U[3 * i + j] = B[3 * i + j] * inv[j]; }
    else bad = j;
  }
  if (bad >= 0) {  // rank-deficient covariance: complete the basis with the cross product
    const int a = (bad + 1) % 3, b = (bad + 2) % 3;
    U[bad] = U[3 + a] * U[6 + b] - U[6 + a] * U[3 + b];
    U[3 + bad] = U[6 + a] * U[b] - U[a] * U[6 + b];
    U[6 + bad] = U[a] * U[3 + b] - U[3 + a] * U[b];
  }
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) {
      double acc = 0;
      for (int k = 0; k < 3; ++k) acc += U[3 * i + k] * V[3 * j + k];
      UVt[3 * i + j] = acc;
    }
  return any_rotation;
}

// xfm = Translation3f{t} * Quaternionf{R} (align_icp.cpp:151), column-major 4x4 out
__device__ void compose_pose(const float* R, const float* t, float* T) {
  float q[4];  // x y z w
  float tr = R[0] + R[4] + R[8];
  if (tr > 0.f) {
    tr = sqrtf(tr + 1.0f);
    q[3] = 0.5f * tr;
    tr = 0.5f / tr;
    q[0] = (R[7] - R[5]) * tr; q[1] = (R[2] - R[6]) * tr; q[2] = (R[3] - R[1]) * tr;
  } else {
    int i = 0;
    if (R[4] > R[0]) i = 1;
    if (R[8] > R[4 * i]) i = 2;
    const int j = (i + 1) % 3, k = (j + 1) % 3;
    tr = sqrtf(R[4 * i] - R[4 * j] - R[4 * k] + 1.0f);
    q[i] = 0.5f * tr;
    tr = 0.5f / tr;
    q[3] = (R[3 * k + j] - R[3 * j + k]) * tr;
    q[j] = (R[3 * j + i] + R[3 * i + j]) * tr;
    q[k] = (R[3 * k + i] + R[3 * i + k]) * tr;
  }
  const float tx = 2.f * q[0], ty = 2.f * q[1], tz = 2.f * q[2];
  const float twx = tx * q[3], twy = ty * q[3], twz = tz * q[3];
  const float txx = tx * q[0], txy = ty * q[0], txz = tz * q[0];
  const float tyy = ty * q[1], tyz = tz * q[1], tzz = tz * q[2];
  const float Rq[9] = {1.f - (tyy + tzz), txy - twz, txz + twy, txy + twz, 1.f - (txx + tzz), tyz - twx,
                       txz - twy, tyz + twx, 1.f - (txx + tyy)};
  for (int r = 0; r < 3; ++r) {
    for (int c = 0; c < 3; ++c) T[r + 4 * c] = Rq[3 * r + c];
    T[12 + r] = t[r];
    T[4 * r + 3] = 0.f;
  }
  T[15] = 1.f;
}

// A thread's own elements (first, first + stride, ...) of NA float4 arrays, streamed through a per-thread ring in shared
// memory with cp.async: 7 (one array) or 3 (two arrays) trips are in flight per thread at no register cost. The
// passes of an iteration are plain sweeps over 0.2-0.5 MB per CTA that lives in L2; with the one or two loads per thread
// a register-bound 1024-thread CTA can keep in flight they ran at a fifth of the L2 bandwidth.
// ring: 8 x kThreads float4 of dynamic shared memory; slot s of array r of thread t at ring[(s * NA + r) * kThreads + t].
template <int NA>
struct OwnStream {
  static constexpr int S = 8 / NA;
  float4* ring;
  const float4* arr[NA];
  int first, stride, n;
  __device__ __forceinline__ void fetch(int trip) {
    const int i = first + trip * stride;
    if (i < n) {
#pragma unroll
      for (int r = 0; r < NA; ++r)
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(ring + ((trip % S) * NA + r) * kThreads + threadIdx.x)),
                     "l"(arr[r] + i) : "memory");
    }
    rst::cp_async_commit();
  }
  __device__ __forceinline__ void start() {
#pragma unroll
    for (int t = 0; t < S - 1; ++t) fetch(t);
  }
  // element `trip` of every array; then the slot read one trip ago (its values are long consumed) takes trip + S - 1
  __device__ __forceinline__ void get(int trip, float4 (&x)[NA]) {
    rst::cp_async_wait<S - 2>();
#pragma unroll
    for (int r = 0; r < NA; ++r) x[r] = ring[((trip % S) * NA + r) * kThreads + threadIdx.x];
    fetch(trip + S - 1);
  }
  __device__ __forceinline__ void finish() { rst::cp_async_wait<0>(); }
};

// Sum of K doubles over every thread of the CTAs that share a pair. CL = false: one CTA, block_sum. CL = true: the
// CTAs of a thread-block cluster; every CTA PUSHES its block sums into row `rank` of a table in every CTA's shared
// memory (distributed shared memory stores), and after one cluster barrier every CTA adds the C rows up in rank order
// from its own copy — the same total, bit for bit, in every CTA, so nothing has to be broadcast back and nobody waits
// for a remote load. `s_tab` alternates between two tables from one call to the next: a CTA writes table b again only
// after the barrier of the call in between, which every CTA reaches after it has read table b.
template <int K, bool CL>
__device__ void pair_sum(double (&v)[K], double (*s_part)[16], double (*s_tab)[16], double* s_out) {
  if (!CL) { block_sum<K>(v, s_part, s_out); return; }
  namespace cg = cooperative_groups;
  cg::cluster_group cluster = cg::this_cluster();
  block_sum<K, true>(v, s_part, s_tab[cluster.block_rank()]);
  cluster.sync();
  if (threadIdx.x < K) {
    double x = 0.0;
    const unsigned C = cluster.num_blocks();
    for (unsigned r = 0; r < C; ++r) x += s_tab[r][threadIdx.x];
    s_out[threadIdx.x] = x;
  }
  __syncthreads();
}

// grid (n_pairs) [CL = false] or (C, n_pairs) in clusters of (C, 1, 1) [CL = true: the C CTAs split the source points
// of one pair — small batches, where one CTA per pair would leave most of the 148 SMs idle].
template <bool CL>
__global__ void __launch_bounds__(kThreads, 1) k_icp3d(const PairDesc* __restrict__ descs, int max_iter, float grid_cell, float3 cache, int skip_fixed) {
  __shared__ double s_part[kWarps][16];
  __shared__ int s_qn;
  __shared__ int s_next;   // the iteration that follows the current one (thread 0's decision, see the solve)
  // s_cum[k & 255]: upper bound of the path length ANY source point has travelled from the initial pose to the pose of
  // iteration k (sum of per-iteration bounds): a point moved at most s_cum[now] - s_cum[k] since iteration k
  __shared__ float s_cum[256];
  __shared__ float s_red[kWarps];
  __shared__ float s_rmax;
  __shared__ double s_sum[16];
  __shared__ double s_tab[2][16][16];   // pair_sum: [call parity][rank][component]
  extern __shared__ float4 s_ring[];   // kRingBytes: the passes' cp.async rings (OwnStream)
  __shared__ float s_T[16];
  __shared__ double s_V[9];   // right singular vectors of the last solve (thread 0)
  __shared__ Grid s_g0;
  namespace cg = cooperative_groups;

  PairDesc P = descs[CL ? blockIdx.y : blockIdx.x];
  if (P.n_ptr) P.n = *P.n_ptr;
  if (P.m_ptr) P.m = *P.m_ptr;
  const int tid = threadIdx.x;
  const int rank = CL ? (int)cg::this_cluster().block_rank() : 0;
  const int stride = CL ? (int)cg::this_cluster().num_blocks() * kThreads : kThreads;
  const int first = rank * kThreads + tid;   // this thread owns source points first, first + stride, ...
  if (P.n < 3 || P.m < 3) {  // align_icp.cpp:77-79: false, pose untouched (uniform over the cluster)
    if (rank == 0 && tid == 0 && P.res) { rst_icp3d_result r{}; *P.res = r; }
    return;
  }

  // ---- uniform grid over dst (built by the first CTA of the pair)
  Grid g;
  if (!CL) {
    g = build_grid(P.dst, P.m, grid_cell, P.cell_start, P.cell_fill, P.sorted);
  } else {
    cg::cluster_group cluster = cg::this_cluster();
    if (rank == 0) {
      g = build_grid(P.dst, P.m, grid_cell, P.cell_start, P.cell_fill, P.sorted);
      if (tid == 0) s_g0 = g;
    }
    cluster.sync();   // release / acquire at cluster scope: the grid arrays in global memory and s_g0 are visible
    g = *cluster.map_shared_rank(&s_g0, 0);
  }

  // ---- src centroid (ComputeCentroid, point_cloud_utils.cpp:92-98), pose -> shared
  float smean[3];
  {
    double acc[3] = {0, 0, 0};
    for (int i = tid; i < P.n; i += kThreads)
      for (int a = 0; a < 3; ++a) acc[a] += (double)P.src[3 * i + a];
    block_sum<3>(acc, s_part, s_sum);
    const float inv = (float)(1.0 / (double)P.n);
    for (int a = 0; a < 3; ++a) smean[a] = (float)s_sum[a] * inv;
  }
  // radius of the source cloud about its centroid (for the neighbour cache's motion bound)
  {
    float r2 = 0.f;
    for (int i = tid; i < P.n; i += kThreads) {
      const float x = P.src[3 * i] - smean[0], y = P.src[3 * i + 1] - smean[1], z = P.src[3 * i + 2] - smean[2];
      const float d = x * x + y * y + z * z;
      r2 = d > r2 ? d : r2;   // a non-finite point never raises it (and never consults the cache)
    }
    for (int o = 16; o > 0; o >>= 1) r2 = fmaxf(r2, __shfl_xor_sync(0xffffffffu, r2, o));
    if ((tid & 31) == 0) s_red[tid >> 5] = r2;
    __syncthreads();
    if (tid == 0) {
      for (int w = 1; w < kWarps; ++w) r2 = fmaxf(r2, s_red[w]);
      s_rmax = sqrtf(r2) * 1.0001f;
      s_cum[0] = 0.f;
    }
    __syncthreads();
  }
  const float rmax = s_rmax;
  if (tid < 16) s_T[tid] = P.pose[tid];
  if (tid < 9) s_V[tid] = tid % 4 == 0 ? 1.0 : 0.0;
  if (tid == 0) s_qn = 0;
  // per source point, two 16-byte records so that no pass of an iteration gathers or chases an index:
  //   sl[i] = (source point, L | k): the proven radius of the neighbour cache with, in its 8 lowest mantissa bits,
  //           the iteration (mod 256) whose pose the point had when L was proven — how far it has moved since is
  //           bounded by the path length s_cum[now] - s_cum[k] instead of storing that position;
  //   qd[i] = (coordinates of the current neighbour, squared distance to it).
  for (int i = first; i < P.n; i += stride) {
    P.sl[i] = make_float4(P.src[3 * i], P.src[3 * i + 1], P.src[3 * i + 2], 0.f);   // L = 0: nothing proven yet
    P.qd[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  __syncthreads();
  // this CTA's part of the search queue: it owns at most n / C + kThreads source points
  int* const queue = P.queue + (CL ? rank * (P.n / (int)cg::this_cluster().num_blocks() + kThreads) : 0);
  const bool cache_on = cache.z > 0.f;
  const float kInf = __int_as_float(0x7f800000);

  float mu = 1.0f;  // align_icp.cpp:91
  double cost = 0.0;
  unsigned long long n_searched = 0;
#ifdef RST_ICP3D_PROFILE
  long long ph_[8] = {0, 0, 0, 0, 0, 0, 0, 0}, t_ = clock64();
#endif
  unsigned iters_run = 0;
  for (int iter = 0; iter < max_iter; iter = s_next) {
    ++iters_run;
    if (iter > 0 && iter % 8 == 0) mu = __fdiv_rn(mu, 1.4f);  // :96-98
    float T[12];
    T[0] = s_T[0]; T[1] = s_T[1]; T[2] = s_T[2]; T[3] = s_T[4]; T[4] = s_T[5]; T[5] = s_T[6];
    T[6] = s_T[8]; T[7] = s_T[9]; T[8] = s_T[10]; T[9] = s_T[12]; T[10] = s_T[13]; T[11] = s_T[14];
    const float moved_now = s_cum[iter & 255];   // written by the previous iteration's solve, before its closing barrier
    PHASE(0);
    // Isometry3f * Vector3f, left to right, no contraction
    auto xform = [](const float* M, float sx, float sy, float sz, float& px, float& py, float& pz) {
      px = addrn(addrn(addrn(mulrn(M[0], sx), mulrn(M[3], sy)), mulrn(M[6], sz)), M[9]);
      py = addrn(addrn(addrn(mulrn(M[1], sx), mulrn(M[4], sy)), mulrn(M[7], sz)), M[10]);
      pz = addrn(addrn(addrn(mulrn(M[2], sx), mulrn(M[5], sy)), mulrn(M[8], sz)), M[11]);
    };
    // ---- correspondences (:105-113) through the neighbour cache, in three passes over this CTA's source points.
    // A: transform, distance to the cached neighbour, cache test (see nn_ball); points that fail it are queued.
    auto pass_a = [&](int i, const float4& a, const float4& b) {
      float px, py, pz;
      xform(T, a.x, a.y, a.z, px, py, pz);
      // a non-finite source point or pose (the reference's callers run RemoveNans first) has no neighbour:
      // index 0, d2 = +inf, hence weight 0, cost = inf and ok = 0 — without walking the whole grid for it
      if (!(isfinite(px) && isfinite(py) && isfinite(pz))) {
        P.nbr[i] = 0;
        P.qd[i] = make_float4(P.dst[0], P.dst[1], P.dst[2], kInf);
        return;
      }
      const int age = (iter - __float_as_int(a.w)) & 255;
      if (cache_on && age < 250) {
        const float dx = subrn(px, b.x), dy = subrn(py, b.y), dz = subrn(pz, b.z);
        const float d2c = addrn(addrn(mulrn(dx, dx), mulrn(dy, dy)), mulrn(dz, dz));   // as the scan computes it
        const float m = __int_as_float(__float_as_int(a.w) & ~255) - (moved_now - s_cum[(iter - age) & 255]);   // L - motion since L was proven
        if (m > 0.f && d2c * 1.0005f < m * m) { P.qd[i].w = d2c; return; }   // |p - nbr| + motion < L: proven, nbr stays
      }
      const int slot = atomicAdd(&s_qn, 1);
      ICP3D_CHECK(slot < (CL ? P.n / (int)cg::this_cluster().num_blocks() + kThreads : P.n));
      queue[slot] = i;
    };
    {
      OwnStream<2> in{s_ring, {P.sl, P.qd}, first, stride, P.n};
      in.start();
      for (int trip = 0, i = first; i < P.n; ++trip, i += stride) {
        float4 x[2];
        in.get(trip, x);
        pass_a(i, x[0], x[1]);
      }
      in.finish();
    }
    PHASE(1);
    __syncthreads();
    PHASE(2);
    // B: the queued points, densely over the CTA's threads: ring search without a candidate (first iteration), then
    // the ball scan that also renews the cache entry. The margin follows the point's motion since its last scan.
    const int qn = s_qn;
    n_searched += qn;
    if (iter > 0 && qn <= kGroupScanMax) {   // few points search: kScanLanes lanes each, so that the barrier below is reached sooner
      for (int q0 = 0; q0 < qn; q0 += kThreads / kScanLanes) {
        const int q = q0 + tid / kScanLanes;
        const bool active = q < qn;
        const int i = active ? queue[q] : 0;
        const float4 a = P.sl[i];
        float px, py, pz;
        xform(T, a.x, a.y, a.z, px, py, pz);
        const int cand = P.nbr[i];
        const float margin = fminf(fmaxf(cache.x * (moved_now - s_cum[__float_as_int(a.w) & 255]), cache.y * g.h), cache.z * g.h);
        int j;
        float d2, L;
        ICP3D_CHECK(i >= 0 && i < P.n && cand >= 0 && cand < P.m);
        nn_ball_group(g, P.cell_start, P.sorted, P.dst, active, px, py, pz, cand, cache_on ? margin : 0.f, &j, &d2, &L);
        ICP3D_CHECK(!active || (j >= 0 && j < P.m));
        if (active && (tid & (kScanLanes - 1)) == 0) {
          P.nbr[i] = j;
          P.sl[i].w = __int_as_float((__float_as_int(L) & ~255) | (iter & 255));
          P.qd[i] = make_float4(P.dst[3 * j], P.dst[3 * j + 1], P.dst[3 * j + 2], d2);
        }
      }
    } else
    for (int q = tid; q < qn; q += kThreads) {
      const int i = queue[q];
      const float4 a = P.sl[i];
      float px, py, pz;
      xform(T, a.x, a.y, a.z, px, py, pz);
      int cand, j;
      float d2, L, margin = cache.y * g.h;
      if (iter == 0) {
        // any target point is a valid candidate for the ball scan; one from the query's own cell makes the ball small
        // (a few cells instead of the 27 of a ring search). An empty cell falls back to the ring search.
        const int c = (cell_coord(pz, g.loz, g.inv_h, g.nz) * g.ny + cell_coord(py, g.loy, g.inv_h, g.ny)) * g.nx + cell_coord(px, g.lox, g.inv_h, g.nx);
        const int k0 = P.cell_start[c];
        if (k0 < P.cell_start[c + 1]) cand = __float_as_int(P.sorted[k0].w);
        else nn_search(g, P.cell_start, P.sorted, px, py, pz, &cand, &d2);
      } else {
        cand = P.nbr[i];
        margin = fminf(fmaxf(cache.x * (moved_now - s_cum[__float_as_int(a.w) & 255]), margin), cache.z * g.h);
      }
      ICP3D_CHECK(i >= 0 && i < P.n && cand >= 0 && cand < P.m);
      nn_ball(g, P.cell_start, P.sorted, P.dst, px, py, pz, cand, cache_on ? margin : 0.f, &j, &d2, &L);
      ICP3D_CHECK(j >= 0 && j < P.m);
      P.nbr[i] = j;
      P.sl[i].w = __int_as_float((__float_as_int(L) & ~255) | (iter & 255));   // L rounded down, tagged with this iteration
      P.qd[i] = make_float4(P.dst[3 * j], P.dst[3 * j + 1], P.dst[3 * j + 2], d2);
    }
    __syncthreads();
    PHASE(3);
    if (tid == 0) s_qn = 0;   // every thread has read qn; visible to the next iteration through pair_sum's barriers
    // C: cost and the neighbour sum in the fixed per-thread order (four points per trip)
    // (four sums: the conversion pipe keeps up with a float -> double conversion per term here; the nine covariance
    // sums below would saturate it and are kept as compensated fp32 pairs instead)
    double a4[4] = {0, 0, 0, 0};  // cost, sum dst_j
    {
      OwnStream<1> in{s_ring, {P.qd}, first, stride, P.n};
      in.start();
      for (int trip = 0, i = first; i < P.n; ++trip, i += stride) {
        float4 b[1];
        in.get(trip, b);
        a4[0] += (double)b[0].w; a4[1] += (double)b[0].x; a4[2] += (double)b[0].y; a4[3] += (double)b[0].z;
      }
      in.finish();
    }
    PHASE(4);
    pair_sum<4, CL>(a4, s_part, s_tab[0], s_sum);
    PHASE(5);
    cost = s_sum[0];
    float dmean[3];
    for (int a = 0; a < 3; ++a) dmean[a] = (float)s_sum[1 + a] / (float)P.n;  // :122, unweighted
    // ---- weights (:114-121) and the weighted cross-covariance: fp32 products, fp64 accumulation (:125-136)
    float hv[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0}, lv[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    const bool last = iter == max_iter - 1;
    {
      OwnStream<2> in{s_ring, {P.sl, P.qd}, first, stride, P.n};
      in.start();
      for (int trip = 0, i = first; i < P.n; ++trip, i += stride) {
        float4 x[2];   // (source point, .), (neighbour, d2)
        in.get(trip, x);
        const float rt = __fdiv_rn(mu, addrn(x[1].w, mu));
        const float w = mulrn(rt, rt);
        if (last) P.w[i] = w;
        const float wd[3] = {mulrn(w, subrn(x[1].x, dmean[0])), mulrn(w, subrn(x[1].y, dmean[1])), mulrn(w, subrn(x[1].z, dmean[2]))};
        const float ds[3] = {subrn(x[0].x, smean[0]), subrn(x[0].y, smean[1]), subrn(x[0].z, smean[2])};
#pragma unroll
        for (int a = 0; a < 3; ++a)
#pragma unroll
          for (int c = 0; c < 3; ++c) acc2(hv[3 * a + c], lv[3 * a + c], mulrn(wd[a], ds[c]));
      }
      in.finish();
    }
    double cv[9];
    for (int k = 0; k < 9; ++k) cv[k] = isfinite(hv[k]) ? (double)hv[k] + (double)lv[k] : (double)hv[k];
    PHASE(6);
    pair_sum<9, CL>(cv, s_part, s_tab[1], s_sum);
    PHASE(5);
    // ---- closed-form pose (:139-151); in a cluster every CTA solves from the same totals (no broadcast)
    if (tid == 0) {
      double cov[9], uvt[9];
      for (int k = 0; k < 9; ++k) cov[k] = s_sum[k];
      const bool v_moved = svd_uvt(cov, uvt, s_V);   // warm start from the previous iteration's right singular vectors
      float R[9], t[3];
      for (int k = 0; k < 9; ++k) R[k] = (float)uvt[k];
      const float det = R[0] * (R[4] * R[8] - R[5] * R[7]) - R[1] * (R[3] * R[8] - R[5] * R[6]) + R[2] * (R[3] * R[7] - R[4] * R[6]);
      if (det < 0) { R[2] *= -1; R[5] *= -1; R[8] *= -1; }  // R.col(2) *= -1, as written (:143-145)
      for (int a = 0; a < 3; ++a) t[a] = dmean[a] - (R[3 * a] * smean[0] + R[3 * a + 1] * smean[1] + R[3 * a + 2] * smean[2]);
      float Tn[16];
      compose_pose(R, t, Tn);
      {  // |T' s - T s| = |(R' - R)(s - c) + T' c - T c| <= |R' - R|_F rmax + |T' c - T c|, c = the source centroid
        float f2 = 0.f, cm[3];
        for (int col = 0; col < 3; ++col)
          for (int r = 0; r < 3; ++r) { const float d = Tn[4 * col + r] - s_T[4 * col + r]; f2 += d * d; }
        for (int r = 0; r < 3; ++r)
          cm[r] = (Tn[r] - s_T[r]) * smean[0] + (Tn[4 + r] - s_T[4 + r]) * smean[1] + (Tn[8 + r] - s_T[8 + r]) * smean[2] + (Tn[12 + r] - s_T[12 + r]);
        const float step = (sqrtf(f2) * rmax + sqrtf(cm[0] * cm[0] + cm[1] * cm[1] + cm[2] * cm[2])) * 1.001f + 1e-7f * rmax;
        // Fixed point: this iteration left the pose bit for bit as it was and the solve left its warm-start basis
        // untouched. The next iteration would then start from the very state this one started from — pose, basis, mu
        // (the cache never changes a result) — and reproduce it, and so would every iteration up to the next change of
        // mu (iterations that are multiples of 8, :96-98). They are not run: the loop continues at that iteration, or at
        // the last one, which is always run so that it leaves its weights and covariance behind. The reference's 128
        // iterations spend a third of their time in such repeats once the alignment has converged in fp32.
        int next = iter + 1;
        bool same = skip_fixed != 0 && !v_moved;
        for (int k = 0; k < 16; ++k) same = same && Tn[k] == s_T[k];
        if (same) {
          const int j = min((iter / 8 + 1) * 8, max_iter - 1);
          if (j > next) next = j;
        }
        s_next = next;
        s_cum[next & 255] = moved_now + step + 2e-7f * moved_now;   // rounded up generously: the sum must not fall short
      }
      for (int k = 0; k < 16; ++k) s_T[k] = Tn[k];
      if (rank == 0 && P.res && iter == max_iter - 1) for (int k = 0; k < 9; ++k) P.res->cov[k] = cov[k];
    }
    __syncthreads();
    PHASE(7);
  }
#ifdef RST_ICP3D_PROFILE
  if (tid == 0 && blockIdx.x == 0 && blockIdx.y == 0)
    printf("k_icp3d phases (cycles of thread 0, %d iterations): head %lld | pass A %lld | wait A %lld | pass B + wait %lld | pass C %lld | sums %lld | cov %lld | solve %lld\n",
           max_iter, ph_[0], ph_[1], ph_[2], ph_[3], ph_[4], ph_[5], ph_[6], ph_[7]);
#endif
  if (tid == 0) atomicAdd(P.stat, n_searched);
  if (CL) cg::this_cluster().sync();   // nobody leaves while its block sums may still be read
  if (rank != 0) return;
  if (tid == 0) {
    P.stat[1] = (unsigned long long)P.n * (unsigned long long)max_iter;
    P.stat[2] = iters_run;
    P.stat[3] = (unsigned long long)max_iter;
  }
  if (tid < 16) P.pose[tid] = s_T[tid];  // :156
  if (tid == 0 && P.res) {
    const float mean_cost = sqrtf((float)cost / (float)P.n);  // :157
    P.res->mean_cost = mean_cost;
    P.res->ok = mean_cost < 10000.f ? 1 : 0;                  // :160
    P.res->iterations = max_iter;
    P.res->mu = mu;
    if (max_iter == 0) for (int k = 0; k < 9; ++k) P.res->cov[k] = 0.0;
  }
}

// One CTA per pair for batches that fill the GPU; for small batches a cluster of C CTAs per pair (C a power of two,
// C * n_pairs <= SM count, at most 16 — beyond 8 is the opt-in cluster size). RST_ICP3D_CLUSTER overrides.
// Measured on B200, 14 k-point clouds, 128 iterations: one pair 13.4 ms (C = 1) -> 4.6 ms (C = 16) including the
// depth -> cloud stage; 8 pairs 14.1 -> 5.9 ms (C = 8; C = 16: 7.8 ms).
cudaError_t launch_icp3d(const PairDesc* descs, int n_pairs, int max_iter, float grid_cell, int forced, float3 cache, int skip_fixed, cudaStream_t stream) {
  static int sm_counts[64] = {0}, c_maxs[64] = {0};   // per device ordinal (function attributes are per device)
  int dev = 0;
  cudaGetDevice(&dev);
  const int di = dev < 64 ? dev : 63;
  if (sm_counts[di] == 0 || dev >= 64) {
    cudaDeviceGetAttribute(&sm_counts[di], cudaDevAttrMultiProcessorCount, dev);
    c_maxs[di] = 8;
    cudaFuncSetAttribute(k_icp3d<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kRingBytes);
    cudaFuncSetAttribute(k_icp3d<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kRingBytes);
    if (cudaFuncSetAttribute(k_icp3d<true>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) == cudaSuccess) {
      cudaLaunchConfig_t cfg{};
      cfg.gridDim = dim3(16, 1, 1);
      cfg.blockDim = dim3(kThreads, 1, 1);
      cfg.dynamicSmemBytes = kRingBytes;
      cudaLaunchAttribute at[1];
      at[0].id = cudaLaunchAttributeClusterDimension;
      at[0].val.clusterDim.x = 16; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
      cfg.attrs = at; cfg.numAttrs = 1;
      int n = 0;
      if (cudaOccupancyMaxActiveClusters(&n, k_icp3d<true>, &cfg) == cudaSuccess && n >= 1) c_maxs[di] = 16;
    }
    cudaGetLastError();
  }
  const int sm_count = sm_counts[di], c_max = c_maxs[di];
  if (forced <= 0)
    if (const char* e = std::getenv("RST_ICP3D_CLUSTER")) forced = std::atoi(e);
  int C = 1;
  if (forced > 0) while (C * 2 <= c_max && C * 2 <= forced) C *= 2;
  else while (C * 2 <= c_max && C * 2 * n_pairs <= (C * 2 > 8 ? sm_count / 2 : sm_count)) C *= 2;   // 16-CTA clusters pack badly: only while they leave half the GPU free
  if (C == 1) {
    k_icp3d<false><<<n_pairs, kThreads, kRingBytes, stream>>>(descs, max_iter, grid_cell, cache, skip_fixed);
    return cudaGetLastError();
  }
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(C, n_pairs, 1);
  cfg.blockDim = dim3(kThreads, 1, 1);
  cfg.dynamicSmemBytes = kRingBytes;
  cfg.stream = stream;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = C; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, k_icp3d<true>, descs, max_iter, grid_cell, cache, skip_fixed);
}

// ----------------------------------------------------------------------------------------------
// depth frame -> the cloud the reference's caller aligns (rs_replay_app.cpp:229,246-247):
// back-projection with invalid pixels mapped to the origin (rs_driver.cpp:83-88,201-202), RemoveNans
// (point_cloud_utils.cpp:163-174; nothing to drop: the back-projection emits no NaN) and
// DownsampleVoxel (point_cloud_utils.cpp:34-68): key = floor(p / voxel), the FIRST point of a voxel
// wins. The reference's output order is unordered_map iteration order (implementation-defined);
// here it is first-occurrence order, which is deterministic. Pass 1 inserts atomicMin(pixel index) per
// voxel into an open-addressing hash table, pass 2 keeps the winners in pixel order (segment counts,
// scan, order-preserving write).
// ----------------------------------------------------------------------------------------------
struct CloudifyDesc {
  const uint16_t* depth;           // dense w*h
  unsigned long long* keys;        // [cap], 0 = empty
  int* vals;                       // [cap]
  float* cloud;                    // out, up to w*h points
  int* count;                      // out
  int* seg;                        // [ceil(w*h / kSegPx)] winners per pixel segment, then their exclusive scan
};

constexpr unsigned long long kEmptyKey = 0ull;

__device__ __forceinline__ void backproject_px(const uint16_t* depth, int i, int w, float fx, float fy, float cx, float cy,
                                               float scale, float* p) {
  const uint32_t d = depth[i];
  if (d == 0u) { p[0] = p[1] = p[2] = 0.f; return; }  // invalid -> origin, kept (rs_driver.cpp:83-88)
  const int v = i / w, u = i - v * w;
  const float z = mulrn((float)d, scale);
  p[0] = __fdiv_rn(mulrn(subrn((float)u, cx), z), fx);
  p[1] = __fdiv_rn(mulrn(subrn((float)v, cy), z), fy);
  p[2] = z;
}

__device__ __forceinline__ unsigned long long voxel_key(const float* p, float voxel) {
  // floor(p / voxel) per axis (point_cloud_utils.cpp:41-42), 21 bits each, +1 so that 0 stays "empty"
  const long long kx = (long long)floorf(__fdiv_rn(p[0], voxel)) + (1 << 20);
  const long long ky = (long long)floorf(__fdiv_rn(p[1], voxel)) + (1 << 20);
  const long long kz = (long long)floorf(__fdiv_rn(p[2], voxel)) + (1 << 20);
  return (((unsigned long long)(kx & 0x1FFFFF) << 42) | ((unsigned long long)(ky & 0x1FFFFF) << 21) | (unsigned long long)(kz & 0x1FFFFF)) + 1ull;
}

__device__ __forceinline__ uint32_t hash_key(unsigned long long k) {
  k ^= k >> 33; k *= 0xff51afd7ed558ccdull; k ^= k >> 33; k *= 0xc4ceb9fe1a85ec53ull; k ^= k >> 33;
  return (uint32_t)k;
}

// Four launches, every one over as many blocks as the frames need (a single frame uses the whole GPU):
//   k_cloudify_insert : one thread per pixel, atomicMin(pixel index) per voxel in the frame's hash table
//   k_cloudify_count  : winners per segment of kSegPx consecutive pixels
//   k_cloudify_scan   : exclusive scan of the segment counts of a frame (one block per frame), point count out
//   k_cloudify_write  : winners of a segment, in pixel order, at the segment's offset
constexpr int kSegThreads = 256, kSegPerThread = 8, kSegPx = kSegThreads * kSegPerThread;

__global__ void __launch_bounds__(256) k_cloudify_insert(const CloudifyDesc* __restrict__ descs, int w, int h, float fx, float fy,
                                                         float cx, float cy, float scale, float voxel, uint32_t cap_mask) {
  const CloudifyDesc D = descs[blockIdx.y];
  const int i = blockIdx.x * 256 + threadIdx.x;
  if (i >= w * h) return;
  float p[3];
  backproject_px(D.depth, i, w, fx, fy, cx, cy, scale, p);
  const unsigned long long key = voxel_key(p, voxel);
  uint32_t slot = hash_key(key) & cap_mask;
  for (;;) {
    const unsigned long long prev = atomicCAS(D.keys + slot, kEmptyKey, key);
    if (prev == kEmptyKey || prev == key) { atomicMin(D.vals + slot, i); break; }
    slot = (slot + 1) & cap_mask;
  }
}

__device__ __forceinline__ bool cloudify_winner(const CloudifyDesc& D, int i, int w, float fx, float fy, float cx, float cy, float scale,
                                                float voxel, uint32_t cap_mask, float* p) {
  backproject_px(D.depth, i, w, fx, fy, cx, cy, scale, p);
  if (!(voxel > 0.f)) return true;
  const unsigned long long key = voxel_key(p, voxel);
  uint32_t slot = hash_key(key) & cap_mask;
  while (D.keys[slot] != key) slot = (slot + 1) & cap_mask;
  return D.vals[slot] == i;
}

// WRITE = false: seg[blockIdx.x] = winners of this segment. WRITE = true: seg[] holds the exclusive scan; the winners
// are written in pixel order (order-preserving compaction: contiguous run per thread, block scan of the run counts).
template <bool WRITE>
__global__ void __launch_bounds__(kSegThreads) k_cloudify_segment(const CloudifyDesc* __restrict__ descs, int w, int h, float fx, float fy,
                                                                   float cx, float cy, float scale, float voxel, uint32_t cap_mask) {
  __shared__ int s_scan[kSegThreads / 32];
  const CloudifyDesc D = descs[blockIdx.y];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int n = w * h;
  const int i0 = min(blockIdx.x * kSegPx + tid * kSegPerThread, n), i1 = min(i0 + kSegPerThread, n);
  float p[3];
  uint32_t win = 0;
  for (int i = i0; i < i1; ++i) win |= (cloudify_winner(D, i, w, fx, fy, cx, cy, scale, voxel, cap_mask, p) ? 1u : 0u) << (i - i0);
  const int local = __popc(win);
  int incl = local;
  for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += v; }
  if (lane == 31) s_scan[warp] = incl;
  __syncthreads();
  int base = 0;
  for (int k = 0; k < warp; ++k) base += s_scan[k];
  if (!WRITE) {
    if (tid == kSegThreads - 1) D.seg[blockIdx.x] = base + incl;
    return;
  }
  int pos = D.seg[blockIdx.x] + base + incl - local;
  for (int i = i0; i < i1; ++i)
    if ((win >> (i - i0)) & 1u) {
      backproject_px(D.depth, i, w, fx, fy, cx, cy, scale, p);
      D.cloud[3 * pos] = p[0]; D.cloud[3 * pos + 1] = p[1]; D.cloud[3 * pos + 2] = p[2];
      ++pos;
    }
}

// exclusive scan of a frame's segment counts in place; the total is the frame's point count
__global__ void __launch_bounds__(kThreads) k_cloudify_scan(const CloudifyDesc* __restrict__ descs, int n_seg) {
  __shared__ int s_scan[kWarps];
  const CloudifyDesc D = descs[blockIdx.x];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int per = (n_seg + kThreads - 1) / kThreads;
  const int k0 = min(tid * per, n_seg), k1 = min(k0 + per, n_seg);
  int local = 0;
  for (int k = k0; k < k1; ++k) local += D.seg[k];
  int incl = local;
  for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += v; }
  if (lane == 31) s_scan[warp] = incl;
  __syncthreads();
  if (warp == 0) {
    int v = s_scan[lane];
    for (int o = 1; o < 32; o <<= 1) { const int u = __shfl_up_sync(0xffffffffu, v, o); if (lane >= o) v += u; }
    s_scan[lane] = v;
  }
  __syncthreads();
  int run = incl - local + (warp > 0 ? s_scan[warp - 1] : 0);
  for (int k = k0; k < k1; ++k) { const int c = D.seg[k]; D.seg[k] = run; run += c; }
  if (tid == kThreads - 1) *D.count = run;
}

// ----------------------------------------------------------------------------------------------
// SolveKabsch (align_icp.cpp:18-71): closed-form pose from GIVEN index pairs, optional weights.
// Centroids over the pairs are unweighted (:28-35); the covariance is the fp32 outer product cast to
// double, multiplied by the weight in double (:46-56); then the same SVD / reflection patch /
// quaternion round trip as AlignIcp3d. One block.
// ----------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads, 1) k_kabsch(const float* __restrict__ src, const float* __restrict__ dst,
                                                        const int2* __restrict__ pairs, int n_pairs,
                                                        const float* __restrict__ weights, float* __restrict__ pose_out) {
  __shared__ double s_part[kWarps][16];
  __shared__ double s_sum[16];
  const int tid = threadIdx.x;
  double m6[6] = {0, 0, 0, 0, 0, 0};
  for (int c = tid; c < n_pairs; c += kThreads) {
    const int2 p = pairs[c];
    for (int a = 0; a < 3; ++a) { m6[a] += (double)src[3 * p.x + a]; m6[3 + a] += (double)dst[3 * p.y + a]; }
  }
  block_sum<6>(m6, s_part, s_sum);
  float sm[3], dm[3];
  for (int a = 0; a < 3; ++a) { sm[a] = (float)s_sum[a] / (float)n_pairs; dm[a] = (float)s_sum[3 + a] / (float)n_pairs; }
  __syncthreads();
  double cv[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
  for (int c = tid; c < n_pairs; c += kThreads) {
    const int2 p = pairs[c];
    const double w = weights ? (double)weights[c] : 1.0;
    float d[3], s3[3];
    for (int a = 0; a < 3; ++a) { d[a] = subrn(dst[3 * p.y + a], dm[a]); s3[a] = subrn(src[3 * p.x + a], sm[a]); }
    for (int a = 0; a < 3; ++a)
      for (int b = 0; b < 3; ++b) cv[3 * a + b] += w * (double)mulrn(d[a], s3[b]);
  }
  block_sum<9>(cv, s_part, s_sum);
  if (tid == 0) {
    double cov[9], uvt[9];
    for (int k = 0; k < 9; ++k) cov[k] = s_sum[k];
    double V0[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
    svd_uvt(cov, uvt, V0);
    float R[9], t[3];
    for (int k = 0; k < 9; ++k) R[k] = (float)uvt[k];
    const float det = R[0] * (R[4] * R[8] - R[5] * R[7]) - R[1] * (R[3] * R[8] - R[5] * R[6]) + R[2] * (R[3] * R[7] - R[4] * R[6]);
    if (det < 0) { R[2] *= -1; R[5] *= -1; R[8] *= -1; }  // :61-63
    for (int a = 0; a < 3; ++a) t[a] = dm[a] - (R[3 * a] * sm[0] + R[3 * a + 1] * sm[1] + R[3 * a + 2] * sm[2]);  // :66
    float T[16];
    compose_pose(R, t, T);  // :69
    for (int k = 0; k < 16; ++k) pose_out[k] = T[k];
  }
}

// ----------------------------------------------------------------------------------------------
// ComputeNormals + OrientNormals (point_cloud_utils.cpp:176-216) on the device: for every point the
// exact k nearest neighbours (k includes the point itself, :184), fp32 centroid and covariance summed
// in ascending-distance order (:187-198), eigenvector of the smallest eigenvalue of the 3x3 covariance
// (SelfAdjointEigenSolver, :201-202; here a cyclic Jacobi in fp32), flipped so that
// n . (p - viewpoint) <= 0 (:210-214). The grid of the cloud is built exactly as for the ICP (one block,
// k_grid_build), then one thread per point over as many blocks as the cloud needs.
// ----------------------------------------------------------------------------------------------
__global__ void k_grid_build(const float* __restrict__ pts, int n, float grid_cell, int* cell_start, int* cell_fill, float4* sorted,
                             struct Grid* out);
constexpr int kMaxK = 33;   // ComputeCovariances asks for 32 neighbours + the point itself

// k nearest neighbours of p by ring expansion; (d2, index) kept sorted ascending, ties to the lower index
__device__ void knn_search(const Grid& g, const int* __restrict__ cell_start, const float4* __restrict__ sorted, float px, float py,
                           float pz, int k, float* bd, int* bj) {
  const int cx = cell_coord(px, g.lox, g.inv_h, g.nx), cy = cell_coord(py, g.loy, g.inv_h, g.ny), cz = cell_coord(pz, g.loz, g.inv_h, g.nz);
  int cnt = 0;
  for (int r = 0;; ++r) {
    const int x0 = max(cx - r, 0), x1 = min(cx + r, g.nx - 1);
    const int y0 = max(cy - r, 0), y1 = min(cy + r, g.ny - 1);
    const int z0 = max(cz - r, 0), z1 = min(cz + r, g.nz - 1);
    for (int z = z0; z <= z1; ++z)
      for (int y = y0; y <= y1; ++y) {
        const bool inner_zy = (abs(z - cz) < r) && (abs(y - cy) < r);
        for (int x = x0; x <= x1; ++x) {
          if (inner_zy && abs(x - cx) < r) { x = cx + r - 1; continue; }
          const int c = (z * g.ny + y) * g.nx + x;
          const int e = cell_start[c + 1];
          for (int q = cell_start[c]; q < e; ++q) {
            const float4 v = sorted[q];
            const float dx = subrn(px, v.x), dy = subrn(py, v.y), dz = subrn(pz, v.z);
            const float d2 = addrn(addrn(mulrn(dx, dx), mulrn(dy, dy)), mulrn(dz, dz));
            const int j = __float_as_int(v.w);
            if (cnt == k && !(d2 < bd[k - 1] || (d2 == bd[k - 1] && j < bj[k - 1]))) continue;
            int i = cnt < k ? cnt : k - 1;   // insertion, as nanoflann's KNNResultSet
            for (; i > 0 && (bd[i - 1] > d2 || (bd[i - 1] == d2 && bj[i - 1] > j)); --i) { bd[i] = bd[i - 1]; bj[i] = bj[i - 1]; }
            bd[i] = d2; bj[i] = j;
            if (cnt < k) ++cnt;
          }
        }
      }
    if (x0 == 0 && x1 == g.nx - 1 && y0 == 0 && y1 == g.ny - 1 && z0 == 0 && z1 == g.nz - 1) break;
    if (cnt < k) continue;
    float bound = FLT_MAX;
    if (cx - r > 0) bound = fminf(bound, px - (g.lox + (float)(cx - r) * g.h));
    if (cx + r < g.nx - 1) bound = fminf(bound, (g.lox + (float)(cx + r + 1) * g.h) - px);
    if (cy - r > 0) bound = fminf(bound, py - (g.loy + (float)(cy - r) * g.h));
    if (cy + r < g.ny - 1) bound = fminf(bound, (g.loy + (float)(cy + r + 1) * g.h) - py);
    if (cz - r > 0) bound = fminf(bound, pz - (g.loz + (float)(cz - r) * g.h));
    if (cz + r < g.nz - 1) bound = fminf(bound, (g.loz + (float)(cz + r + 1) * g.h) - pz);
    bound -= 1e-3f * g.h;
    if (bound > 0.0f && bd[k - 1] <= bound * bound * 0.9999f) break;
  }
  for (int i = cnt; i < k; ++i) { bd[i] = 0.f; bj[i] = bj[0]; }  // fewer than k points in the cloud: pad with the nearest
}

// eigenvector of the smallest eigenvalue of a symmetric 3x3 (cyclic Jacobi, fp32)
__device__ void smallest_eigenvector(const float* C, float* out) {
  float a[9], v[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
  for (int i = 0; i < 9; ++i) a[i] = C[i];
  for (int sweep = 0; sweep < 30; ++sweep) {
    float off = 0.f;
    for (int p = 0; p < 2; ++p)
      for (int q = p + 1; q < 3; ++q) {
        off = fmaxf(off, fabsf(a[3 * p + q]));
        if (fabsf(a[3 * p + q]) <= 1e-37f) continue;
        const float theta = (a[3 * q + q] - a[3 * p + p]) / (2.f * a[3 * p + q]);
        const float t = (theta >= 0 ? 1.f : -1.f) / (fabsf(theta) + sqrtf(1.f + theta * theta));
        const float c = 1.f / sqrtf(1.f + t * t), s = c * t;
        for (int k = 0; k < 3; ++k) { const float akp = a[3 * k + p], akq = a[3 * k + q]; a[3 * k + p] = c * akp - s * akq; a[3 * k + q] = s * akp + c * akq; }
        for (int k = 0; k < 3; ++k) { const float apk = a[3 * p + k], aqk = a[3 * q + k]; a[3 * p + k] = c * apk - s * aqk; a[3 * q + k] = s * apk + c * aqk; }
        for (int k = 0; k < 3; ++k) { const float vkp = v[3 * k + p], vkq = v[3 * k + q]; v[3 * k + p] = c * vkp - s * vkq; v[3 * k + q] = s * vkp + c * vkq; }
      }
    if (off < 1e-12f * (fabsf(a[0]) + fabsf(a[4]) + fabsf(a[8]))) break;
  }
  int m = 0;
  if (a[4] < a[0]) m = 1;
  if (a[8] < a[4 * m]) m = 2;
  out[0] = v[m]; out[1] = v[3 + m]; out[2] = v[6 + m];
}

__global__ void __launch_bounds__(128) k_normals(const Grid* __restrict__ gp, const int* __restrict__ cell_start,
                                                 const float4* __restrict__ sorted, const float* __restrict__ pts, int n, int k,
                                                 float vx, float vy, float vz, float* __restrict__ normals) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const Grid g = *gp;
  const float px = pts[3 * i], py = pts[3 * i + 1], pz = pts[3 * i + 2];
  float bd[kMaxK]; int bj[kMaxK];
  knn_search(g, cell_start, sorted, px, py, pz, k, bd, bj);
  float cen[3] = {0.f, 0.f, 0.f};
  for (int q = 0; q < k; ++q) for (int a = 0; a < 3; ++a) cen[a] += pts[3 * bj[q] + a];   // :188-191
  for (int a = 0; a < 3; ++a) cen[a] /= (float)k;
  float C[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
  for (int q = 0; q < k; ++q) {                                                            // :194-198
    float d[3];
    for (int a = 0; a < 3; ++a) d[a] = pts[3 * bj[q] + a] - cen[a];
    for (int a = 0; a < 3; ++a) for (int b = 0; b < 3; ++b) C[3 * a + b] += d[a] * d[b];
  }
  float nv[3];
  smallest_eigenvector(C, nv);                                                             // :201-202
  const float ray = (px - vx) * nv[0] + (py - vy) * nv[1] + (pz - vz) * nv[2];            // :209-213
  const float sgn = ray > 0.f ? -1.f : 1.f;
  for (int a = 0; a < 3; ++a) normals[3 * i + a] = sgn * nv[a];
}

// ----------------------------------------------------------------------------------------------
// Multi-block cloud utilities: the grid of a cloud is built once (one block, k_grid_build) and then
// queried by as many blocks as the cloud needs, so a single cloud uses the whole GPU.
// ----------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads, 1) k_grid_build(const float* __restrict__ pts, int n, float grid_cell, int* cell_start,
                                                            int* cell_fill, float4* sorted, Grid* out) {
  const Grid g = build_grid(pts, n, grid_cell, cell_start, cell_fill, sorted);
  if (threadIdx.x == 0) *out = g;
}

// FindCorrespondences (point_cloud_utils.cpp:70-90): exact 1-NN of every query point in the gridded target cloud.
// A non-finite query has no neighbour: index -1, squared distance +inf.
__global__ void __launch_bounds__(128) k_nn_query(const Grid* __restrict__ gp, const int* __restrict__ cell_start,
                                                  const float4* __restrict__ sorted, const float* __restrict__ q, int nq,
                                                  int* __restrict__ idx, float* __restrict__ d2) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nq) return;
  const Grid g = *gp;
  const float px = q[3 * i], py = q[3 * i + 1], pz = q[3 * i + 2];
  int j = -1; float d = __int_as_float(0x7f800000);
  if (isfinite(px) && isfinite(py) && isfinite(pz)) nn_search(g, cell_start, sorted, px, py, pz, &j, &d);
  idx[i] = j; d2[i] = d;
}

// KDTree3f::query(point, num_closest, out_indices, out_distances_sq) (kdtree.hpp:51-57) for a batch of query points:
// the k nearest points in ascending distance (ties: lower index first), k x nq outputs, row i = query i. Entries past the
// cloud's size, and every entry of a non-finite query, are index -1 / distance +inf.
__global__ void __launch_bounds__(128) k_knn_query(const Grid* __restrict__ gp, const int* __restrict__ cell_start,
                                                   const float4* __restrict__ sorted, int m, const float* __restrict__ q, int nq, int k,
                                                   int* __restrict__ idx, float* __restrict__ d2) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nq) return;
  const Grid g = *gp;
  const float px = q[3 * i], py = q[3 * i + 1], pz = q[3 * i + 2];
  float bd[kMaxK]; int bj[kMaxK];
  int have = 0;
  if (isfinite(px) && isfinite(py) && isfinite(pz)) {
    knn_search(g, cell_start, sorted, px, py, pz, k, bd, bj);
    have = min(k, m);
  }
  for (int e = 0; e < k; ++e) {
    idx[(size_t)i * k + e] = e < have ? bj[e] : -1;
    d2[(size_t)i * k + e] = e < have ? bd[e] : __int_as_float(0x7f800000);
  }
}

// eigen-decomposition of a symmetric 3x3 (cyclic Jacobi, fp32): values descending, vectors in the columns of V
__device__ void sym_eig3_desc(const float* C, float* val, float* V) {
  float a[9], v[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
  for (int i = 0; i < 9; ++i) a[i] = C[i];
  for (int sweep = 0; sweep < 30; ++sweep) {
    float off = 0.f;
    for (int p = 0; p < 2; ++p)
      for (int q = p + 1; q < 3; ++q) {
        off = fmaxf(off, fabsf(a[3 * p + q]));
        if (fabsf(a[3 * p + q]) <= 1e-37f) continue;
        const float theta = (a[3 * q + q] - a[3 * p + p]) / (2.f * a[3 * p + q]);
        const float t = (theta >= 0 ? 1.f : -1.f) / (fabsf(theta) + sqrtf(1.f + theta * theta));
        const float c = 1.f / sqrtf(1.f + t * t), s = c * t;
        for (int k = 0; k < 3; ++k) { const float akp = a[3 * k + p], akq = a[3 * k + q]; a[3 * k + p] = c * akp - s * akq; a[3 * k + q] = s * akp + c * akq; }
        for (int k = 0; k < 3; ++k) { const float apk = a[3 * p + k], aqk = a[3 * q + k]; a[3 * p + k] = c * apk - s * aqk; a[3 * q + k] = s * apk + c * aqk; }
        for (int k = 0; k < 3; ++k) { const float vkp = v[3 * k + p], vkq = v[3 * k + q]; v[3 * k + p] = c * vkp - s * vkq; v[3 * k + q] = s * vkp + c * vkq; }
      }
    if (off < 1e-12f * (fabsf(a[0]) + fabsf(a[4]) + fabsf(a[8]))) break;
  }
  int o[3] = {0, 1, 2};
  const float d[3] = {a[0], a[4], a[8]};
  if (d[o[0]] < d[o[1]]) { const int t = o[0]; o[0] = o[1]; o[1] = t; }
  if (d[o[1]] < d[o[2]]) { const int t = o[1]; o[1] = o[2]; o[2] = t; }
  if (d[o[0]] < d[o[1]]) { const int t = o[0]; o[0] = o[1]; o[1] = t; }
  for (int k = 0; k < 3; ++k) {
    val[k] = d[o[k]];
    for (int r = 0; r < 3; ++r) V[3 * r + k] = v[3 * r + o[k]];
  }
}

// ComputeCovariances (point_cloud_utils.cpp:100-161): for every point the 32 nearest OTHER points (knnSearch of 33,
// entry 0 = the point itself is skipped, :121-134), fp32 centroid and scatter matrix in ascending-distance order;
// use_gicp = 0: divided by 31 (:157); use_gicp = 1: the scatter's singular vectors U recombined with singular values
// (1, 1, 1e-2) (:139-154). covs_out: n x 9 floats, row-major 3x3 (the matrices are symmetric).
__global__ void __launch_bounds__(128) k_covariances(const Grid* __restrict__ gp, const int* __restrict__ cell_start,
                                                     const float4* __restrict__ sorted, const float* __restrict__ pts, int n,
                                                     int use_gicp, float* __restrict__ covs) {
  constexpr int kNb = 32;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const Grid g = *gp;
  float bd[kMaxK]; int bj[kMaxK];
  knn_search(g, cell_start, sorted, pts[3 * i], pts[3 * i + 1], pts[3 * i + 2], kNb + 1, bd, bj);
  float cen[3] = {0.f, 0.f, 0.f};
  for (int q = 1; q <= kNb; ++q) for (int a = 0; a < 3; ++a) cen[a] += pts[3 * bj[q] + a];
  for (int a = 0; a < 3; ++a) cen[a] /= (float)kNb;
  float C[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
  for (int q = 1; q <= kNb; ++q) {
    float d[3];
    for (int a = 0; a < 3; ++a) d[a] = pts[3 * bj[q] + a] - cen[a];
    for (int a = 0; a < 3; ++a) for (int b = 0; b < 3; ++b) C[3 * a + b] += d[a] * d[b];
  }
  if (use_gicp) {
    float val[3], U[9];
    sym_eig3_desc(C, val, U);
    for (int e = 0; e < 9; ++e) C[e] = 0.f;
    for (int k = 0; k < 3; ++k) {
      const float v = k == 2 ? 1e-2f : 1.f;   // gicp_epsilon for the smallest singular value (:147-151)
      for (int a = 0; a < 3; ++a) for (int b = 0; b < 3; ++b) C[3 * a + b] += v * U[3 * a + k] * U[3 * b + k];
    }
  } else {
    for (int e = 0; e < 9; ++e) C[e] /= (float)(kNb - 1);
  }
  for (int e = 0; e < 9; ++e) covs[9 * (size_t)i + e] = C[e];
}

// ComputeCentroid (point_cloud_utils.cpp:92-98) for a caller that holds a cloud: one block, fp64 partial sums in a
// fixed order (thread-strided, warp butterfly, warps in order), so the result is deterministic; the reference sums
// sequentially in fp32, this sum is the more accurate one (they agree to fp32 round-off of the n-term sum).
__global__ void __launch_bounds__(kThreads, 1) k_centroid(const float* __restrict__ pts, int n, float* __restrict__ out) {
  __shared__ double s_part[kWarps][3];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  double acc[3] = {0.0, 0.0, 0.0};
  for (int j = tid; j < n; j += kThreads)
    for (int a = 0; a < 3; ++a) acc[a] += (double)pts[3 * j + a];
  for (int a = 0; a < 3; ++a)
    for (int o = 16; o > 0; o >>= 1) acc[a] += __shfl_xor_sync(0xffffffffu, acc[a], o);
  if (lane == 0) for (int a = 0; a < 3; ++a) s_part[warp][a] = acc[a];
  __syncthreads();
  if (tid < 3) {
    double t = 0.0;
    for (int w = 0; w < kWarps; ++w) t += s_part[w][tid];
    out[tid] = (float)(t * (1.0 / (double)n));
  }
}

// ComputeExtents (point_cloud_utils.cpp:26-32): axis-aligned bounding box of the cloud, out = {lo xyz, hi xyz}; the empty
// box is {FLT_MAX.., -FLT_MAX..} (Eigen::AlignedBox::setEmpty). min / max are exact, so any order gives the same bits.
__global__ void __launch_bounds__(kThreads, 1) k_extents(const float* __restrict__ pts, int n, float* __restrict__ out) {
  __shared__ float s_lohi[kWarps][6];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  float lo[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, hi[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
  for (int j = tid; j < n; j += kThreads)
    for (int a = 0; a < 3; ++a) { const float v = pts[3 * j + a]; lo[a] = fminf(lo[a], v); hi[a] = fmaxf(hi[a], v); }
  for (int a = 0; a < 3; ++a)
    for (int o = 16; o > 0; o >>= 1) {
      lo[a] = fminf(lo[a], __shfl_xor_sync(0xffffffffu, lo[a], o));
      hi[a] = fmaxf(hi[a], __shfl_xor_sync(0xffffffffu, hi[a], o));
    }
  if (lane == 0) for (int a = 0; a < 3; ++a) { s_lohi[warp][a] = lo[a]; s_lohi[warp][3 + a] = hi[a]; }
  __syncthreads();
  if (tid < 6) {
    float v = s_lohi[0][tid];
    for (int w = 1; w < kWarps; ++w) v = tid < 3 ? fminf(v, s_lohi[w][tid]) : fmaxf(v, s_lohi[w][tid]);
    out[tid] = v;
  }
}

// OrientNormals (point_cloud_utils.cpp:205-216): a normal that points along the viewing ray p - viewpoint is negated.
__global__ void __launch_bounds__(256) k_orient_normals(const float* __restrict__ pts, int n, float vx, float vy, float vz,
                                                        float* __restrict__ normals) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float rx = subrn(pts[3 * i], vx), ry = subrn(pts[3 * i + 1], vy), rz = subrn(pts[3 * i + 2], vz);
  const float nx = normals[3 * i], ny = normals[3 * i + 1], nz = normals[3 * i + 2];
  const float dot = addrn(mulrn(rx, nx), addrn(mulrn(ry, ny), mulrn(rz, nz)));
  if (dot > 0.f) { normals[3 * i] = -nx; normals[3 * i + 1] = -ny; normals[3 * i + 2] = -nz; }
}

// DownsampleVoxel (point_cloud_utils.cpp:34-68) for a cloud: key = floor(p / voxel), the FIRST point of a voxel wins.
// Pass 1 (any number of blocks): atomicMin(point index) per voxel in an open-addressing table.
__global__ void __launch_bounds__(256) k_voxel_insert(const float* __restrict__ pts, int n, float voxel, unsigned long long* keys,
                                                      int* vals, uint32_t cap_mask) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float p[3] = {pts[3 * i], pts[3 * i + 1], pts[3 * i + 2]};
  const unsigned long long key = voxel_key(p, voxel);
  uint32_t slot = hash_key(key) & cap_mask;
  for (;;) {
    const unsigned long long prev = atomicCAS(keys + slot, kEmptyKey, key);
    if (prev == kEmptyKey || prev == key) { atomicMin(vals + slot, i); break; }
    slot = (slot + 1) & cap_mask;
  }
}

// Pass 2 (one block): order-preserving compaction of the points a predicate keeps. MODE 0: voxel winners
// (first-occurrence order — the reference's order is unordered_map iteration order, implementation-defined);
// MODE 1: RemoveNans (point_cloud_utils.cpp:163-174), all three coordinates finite.
template <int MODE>
__global__ void __launch_bounds__(kThreads, 1) k_compact_points(const float* __restrict__ pts, int n, float voxel,
                                                                const unsigned long long* __restrict__ keys, const int* __restrict__ vals,
                                                                uint32_t cap_mask, float* __restrict__ out, int* __restrict__ count) {
  __shared__ int s_scan[kWarps];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int per = (n + kThreads - 1) / kThreads;
  const int i0 = min(tid * per, n), i1 = min(i0 + per, n);
  auto keep = [&](int i) -> bool {
    const float p[3] = {pts[3 * i], pts[3 * i + 1], pts[3 * i + 2]};
    if (MODE == 1) return isfinite(p[0]) && isfinite(p[1]) && isfinite(p[2]);
    const unsigned long long key = voxel_key(p, voxel);
    uint32_t slot = hash_key(key) & cap_mask;
    while (keys[slot] != key) slot = (slot + 1) & cap_mask;
    return vals[slot] == i;
  };
  int local = 0;
  for (int i = i0; i < i1; ++i) local += keep(i) ? 1 : 0;
  int incl = local;
  for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += v; }
  if (lane == 31) s_scan[warp] = incl;
  __syncthreads();
  if (warp == 0) {
    int v = s_scan[lane];
    for (int o = 1; o < 32; o <<= 1) { const int u = __shfl_up_sync(0xffffffffu, v, o); if (lane >= o) v += u; }
    s_scan[lane] = v;
  }
  __syncthreads();
  int pos = incl - local + (warp > 0 ? s_scan[warp - 1] : 0);
  for (int i = i0; i < i1; ++i)
    if (keep(i)) { out[3 * pos] = pts[3 * i]; out[3 * pos + 1] = pts[3 * i + 1]; out[3 * pos + 2] = pts[3 * i + 2]; ++pos; }
  if (tid == kThreads - 1) *count = pos;
}

// ----------------------------------------------------------------------------------------------
// GICP plane-to-plane residual (gicp_cost.hpp:40-73) + ceres::HuberLoss(0.5) (align_gicp.cpp:67), one thread per
// correspondence:  e = C^{-1/2} (R s + t - d),  C = C_d + R C_s R^T,  C^{-1/2} = V diag(lambda^{-1/2}) V^T from the
// eigen-decomposition of the symmetric C (the reference calls the general EigenSolver only because the self-adjoint
// one does not compile with ceres::Jet, :58-62).  Gauss-Newton linearisation with C held at the current rotation:
// J = C^{-1/2} [ -[p']x | I ] (left perturbation, p' = R s + t), robust weight w = rho'(|e|^2).  Accumulates
// cost = 1/2 sum rho(|e|^2) (ceres' final_cost convention, :113), A = sum w J^T J (21), b = sum w J^T e (6), count:
// 29 sums, fp64, block partials in block order (deterministic), summed by k_gicp_finish.
// ----------------------------------------------------------------------------------------------
constexpr int kGicpThreads = 128;
constexpr int kGicpSums = 29;   // A(21) b(6) cost count

__global__ void __launch_bounds__(kGicpThreads) k_gicp_residuals(const float* __restrict__ src, const float* __restrict__ dst,
                                                                 const float* __restrict__ src_cov, const float* __restrict__ dst_cov,
                                                                 const int* __restrict__ dst_idx, int n, int m, const float* __restrict__ pose_cm,
                                                                 float huber, float* __restrict__ resid, double* __restrict__ partials) {
  __shared__ double s_part[kGicpThreads / 32][kGicpSums];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int i = blockIdx.x * kGicpThreads + tid;
  double acc[kGicpSums];
#pragma unroll
  for (int k = 0; k < kGicpSums; ++k) acc[k] = 0.0;
  const int j = i < n ? dst_idx[i] : -1;
  if (i < n && j >= 0 && j < m) {
    float R[9], t[3];
    for (int r = 0; r < 3; ++r) { for (int c = 0; c < 3; ++c) R[3 * r + c] = pose_cm[r + 4 * c]; t[r] = pose_cm[12 + r]; }
    float p[3], delta[3];
    for (int r = 0; r < 3; ++r) {
      p[r] = R[3 * r] * src[3 * i] + R[3 * r + 1] * src[3 * i + 1] + R[3 * r + 2] * src[3 * i + 2] + t[r];
      delta[r] = p[r] - dst[3 * j + r];
    }
    float RC[9], C[9];
    const float* Cs = src_cov + 9 * (size_t)i;
    const float* Cd = dst_cov + 9 * (size_t)j;
    for (int r = 0; r < 3; ++r)
      for (int c = 0; c < 3; ++c) RC[3 * r + c] = R[3 * r] * Cs[c] + R[3 * r + 1] * Cs[3 + c] + R[3 * r + 2] * Cs[6 + c];
    for (int r = 0; r < 3; ++r)
      for (int c = 0; c < 3; ++c) C[3 * r + c] = Cd[3 * r + c] + (RC[3 * r] * R[3 * c] + RC[3 * r + 1] * R[3 * c + 1] + RC[3 * r + 2] * R[3 * c + 2]);
    for (int r = 0; r < 3; ++r)   // symmetrise: the Jacobi sweep reads both triangles
      for (int c = r + 1; c < 3; ++c) { const float v = 0.5f * (C[3 * r + c] + C[3 * c + r]); C[3 * r + c] = C[3 * c + r] = v; }
    float val[3], V[9], W[9];
    sym_eig3_desc(C, val, V);
    for (int r = 0; r < 3; ++r)
      for (int c = 0; c < 3; ++c) {
        float a = 0.f;
        for (int k = 0; k < 3; ++k) a += V[3 * r + k] * rsqrtf(val[k]) * V[3 * c + k];
        W[3 * r + c] = a;   // C^{-1/2}
      }
    float e[3];
    for (int r = 0; r < 3; ++r) e[r] = W[3 * r] * delta[0] + W[3 * r + 1] * delta[1] + W[3 * r + 2] * delta[2];
    if (resid) { resid[3 * i] = e[0]; resid[3 * i + 1] = e[1]; resid[3 * i + 2] = e[2]; }
    const float s = e[0] * e[0] + e[1] * e[1] + e[2] * e[2];
    float w = 1.f, rho = s;
    if (huber > 0.f && s > huber * huber) { const float rt = sqrtf(s); w = huber / rt; rho = 2.f * huber * rt - huber * huber; }
    if (isfinite(s)) {
      // J = W [ -[p']x | I ]: columns 0..2 = W * (-[p']x), columns 3..5 = W
      float J[3][6];
      for (int r = 0; r < 3; ++r) {
        J[r][0] = W[3 * r + 2] * p[1] - W[3 * r + 1] * p[2];    // -[p']x = [[0, p2, -p1], [-p2, 0, p0], [p1, -p0, 0]]
        J[r][1] = W[3 * r] * p[2] - W[3 * r + 2] * p[0];
        J[r][2] = W[3 * r + 1] * p[0] - W[3 * r] * p[1];
        J[r][3] = W[3 * r]; J[r][4] = W[3 * r + 1]; J[r][5] = W[3 * r + 2];
      }
      int k = 0;
      for (int a = 0; a < 6; ++a)
        for (int b = a; b < 6; ++b) {
          acc[k++] = (double)(w * (J[0][a] * J[0][b] + J[1][a] * J[1][b] + J[2][a] * J[2][b]));
        }
      for (int a = 0; a < 6; ++a) acc[21 + a] = (double)(w * (J[0][a] * e[0] + J[1][a] * e[1] + J[2][a] * e[2]));
      acc[27] = 0.5 * (double)rho;
      acc[28] = 1.0;
    }
  }
#pragma unroll
  for (int k = 0; k < kGicpSums; ++k) {
    double x = acc[k];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
    if (lane == 0) s_part[warp][k] = x;
  }
  __syncthreads();
  if (tid < kGicpSums) {
    double x = 0.0;
    for (int w2 = 0; w2 < kGicpThreads / 32; ++w2) x += s_part[w2][tid];
    partials[(size_t)blockIdx.x * kGicpSums + tid] = x;
  }
}

__global__ void k_transform_points(const float* __restrict__ in, int n, const float* __restrict__ pose_cm, float* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float x = in[3 * i], y = in[3 * i + 1], z = in[3 * i + 2];
  for (int r = 0; r < 3; ++r) out[3 * i + r] = pose_cm[r] * x + pose_cm[4 + r] * y + pose_cm[8 + r] * z + pose_cm[12 + r];
}

struct GicpState {
  double sums[kGicpSums];       // of the last evaluation (at Rt)
  double good_sums[kGicpSums];  // of the last ACCEPTED pose (at good_Rt)
  double Rt[12];                // fp64 pose the last evaluation ran at: row-major R then t
  double good_Rt[12];
  float pose_cm[16];            // fp32 copy of Rt, what the kernels read
  double lambda;                // Levenberg-Marquardt damping
  int have_good, accepted, rejected;
};

__device__ void gicp_set_pose(GicpState* st, const double* Rt) {
  for (int k = 0; k < 12; ++k) st->Rt[k] = Rt[k];
  for (int r = 0; r < 3; ++r) {
    for (int c = 0; c < 3; ++c) st->pose_cm[r + 4 * c] = (float)Rt[3 * r + c];
    st->pose_cm[12 + r] = (float)Rt[9 + r];
    st->pose_cm[4 * r + 3] = 0.f;
  }
  st->pose_cm[15] = 1.f;
}

// One warp: sums the block partials in block order, then (lane 0)
//   mode 0: evaluation only;
//   mode 1: Levenberg-Marquardt bookkeeping + step: the pose just evaluated is accepted when it did not raise the
//           cost of the last accepted pose (lambda / 3), else rejected (back to the accepted pose, lambda * 10);
//           the next trial pose is Exp(xi) * accepted pose with (A + lambda diag A) xi = -b;
//   mode 2: as 1 after new correspondences (costs are not comparable: the evaluated pose is accepted as is);
//   mode 3: final check: a trial pose that raised the cost is dropped, sums = those of the accepted pose.
__global__ void __launch_bounds__(32) k_gicp_finish(const double* __restrict__ partials, int n_blocks, GicpState* st, int mode) {
  const int lane = threadIdx.x;
  double s = 0.0;
  if (lane < kGicpSums) for (int b = 0; b < n_blocks; ++b) s += partials[(size_t)b * kGicpSums + lane];
  if (lane < kGicpSums) st->sums[lane] = s;
  __syncwarp();
  if (lane != 0 || mode == 0) return;
  const bool worse = mode != 2 && st->have_good && !(st->sums[27] <= st->good_sums[27]);
  if (worse) {
    st->lambda = fmin(st->lambda * 10.0, 1e6);
    st->rejected += 1;
  } else {
    for (int k = 0; k < kGicpSums; ++k) st->good_sums[k] = st->sums[k];
    for (int k = 0; k < 12; ++k) st->good_Rt[k] = st->Rt[k];
    if (st->have_good && mode != 2) st->lambda = fmax(st->lambda / 3.0, 1e-9);
    st->have_good = 1;
    st->accepted += 1;
  }
  if (mode == 3) {
    if (worse) { gicp_set_pose(st, st->good_Rt); for (int k = 0; k < kGicpSums; ++k) st->sums[k] = st->good_sums[k]; }
    return;
  }
  double A[21], b[6], xi[6];
  for (int k = 0; k < 21; ++k) A[k] = st->good_sums[k];
  for (int k = 0; k < 6; ++k) b[k] = st->good_sums[21 + k];
  const int dg[6] = {0, 6, 11, 15, 18, 20};
  for (int k = 0; k < 6; ++k) A[dg[k]] *= (1.0 + st->lambda);
  double Rn[12];
  for (int k = 0; k < 12; ++k) Rn[k] = st->good_Rt[k];
  if (rst::solve6(A, b, (int)st->good_sums[28], 6, 0.0, xi) == RST_STATUS_OK) {
    rst::se3_update(xi, Rn);
    bool fin = true;
    for (int k = 0; k < 12; ++k) fin &= isfinite(Rn[k]);
    if (!fin) for (int k = 0; k < 12; ++k) Rn[k] = st->good_Rt[k];
  }
  gicp_set_pose(st, Rn);
}

/* grow-only device/pinned arenas of the cloud engine, owned by the context */
struct Icp3dState {
  void* d_arena = nullptr; size_t d_bytes = 0;
  void* h_arena = nullptr; size_t h_bytes = 0;
  // layout of the last rst_icp3d_depth call (for rst_icp3d_read_cloud)
  int last_frames = 0;
  size_t last_npx = 0, last_cloud_off = 0;
  int icp3d_cluster = 0;   // CTAs per pair of k_icp3d, 0 = automatic
  // neighbour cache of k_icp3d: scan margin = clamp(x * motion since the last scan, y * cell, z * cell); z <= 0 = off
  float3 cache = make_float3(kCacheGain, kCacheLo, kCacheHi);
  unsigned long long searched = 0, queried = 0, iters_run = 0, iters_asked = 0;   // of the last rst_icp3d_pairs / rst_icp3d_depth call
  int skip_fixed = 1;   // k_icp3d jumps over iterations that provably repeat the previous one
  // rst_icp3d_depth uploads its frames in chunks on a stream of its own, the depth -> cloud kernels of a chunk waiting
  // for that chunk only
  cudaStream_t copy = nullptr;
  cudaEvent_t ev[kUploadChunks + 1] = {};
};

inline void sum_stats(Icp3dState* st, const unsigned long long* s, int n_pairs) {
  st->searched = st->queried = st->iters_run = st->iters_asked = 0;
  for (int i = 0; i < n_pairs; ++i) {
    st->searched += s[4 * i]; st->queried += s[4 * i + 1]; st->iters_run += s[4 * i + 2]; st->iters_asked += s[4 * i + 3];
  }
}

void icp3d_free(void* p) {
  Icp3dState* s = static_cast<Icp3dState*>(p);
  if (s->copy) cudaStreamDestroy(s->copy);
  for (cudaEvent_t e : s->ev) if (e) cudaEventDestroy(e);
  cudaFree(s->d_arena);
  cudaFreeHost(s->h_arena);
  delete s;
}

inline size_t align_up(size_t x) { return (x + 255) & ~size_t(255); }

}  // namespace

extern "C" int32_t rst_set_icp3d_cluster(rst_ctx* c, int32_t ctas_per_pair) {
  if (!c) return RST_ERR_INVALID_ARG;
  if (ctas_per_pair != 0 && ctas_per_pair != 1 && ctas_per_pair != 2 && ctas_per_pair != 4 && ctas_per_pair != 8 && ctas_per_pair != 16) {
    rst::ctx_set_error(c, "ctas_per_pair must be 0, 1, 2, 4, 8 or 16");
    return RST_ERR_INVALID_ARG;
  }
  void (**free_fn)(void*) = nullptr;
  void** slot = rst::ctx_ext_slot(c, &free_fn);
  if (!*slot) { *slot = new Icp3dState(); *free_fn = icp3d_free; }
  static_cast<Icp3dState*>(*slot)->icp3d_cluster = ctas_per_pair;
  return RST_OK;
}

extern "C" int32_t rst_set_icp3d_cache(rst_ctx* c, float gain, float lo_cells, float hi_cells) {
  if (!c) return RST_ERR_INVALID_ARG;
  if (!(gain >= 0.f) || !(lo_cells >= 0.f) || !(hi_cells >= 0.f) || !(hi_cells <= 16.f) || lo_cells > hi_cells) {
    rst::ctx_set_error(c, "rst_set_icp3d_cache: gain >= 0 and 0 <= lo_cells <= hi_cells <= 16 expected");
    return RST_ERR_INVALID_ARG;
  }
  void (**free_fn)(void*) = nullptr;
  void** slot = rst::ctx_ext_slot(c, &free_fn);
  if (!*slot) { *slot = new Icp3dState(); *free_fn = icp3d_free; }
  static_cast<Icp3dState*>(*slot)->cache = make_float3(gain, lo_cells, hi_cells);
  return RST_OK;
}

extern "C" int32_t rst_set_icp3d_fixed_point_skip(rst_ctx* c, int32_t on) {
  if (!c) return RST_ERR_INVALID_ARG;
  void (**free_fn)(void*) = nullptr;
  void** slot = rst::ctx_ext_slot(c, &free_fn);
  if (!*slot) { *slot = new Icp3dState(); *free_fn = icp3d_free; }
  static_cast<Icp3dState*>(*slot)->skip_fixed = on ? 1 : 0;
  return RST_OK;
}

extern "C" int32_t rst_icp3d_iteration_stats(rst_ctx* c, uint64_t* run_out, uint64_t* asked_out) {
  if (!c) return RST_ERR_INVALID_ARG;
  void (**free_fn)(void*) = nullptr;
  void** slot = rst::ctx_ext_slot(c, &free_fn);
  const Icp3dState* st = static_cast<const Icp3dState*>(*slot);
  if (run_out) *run_out = st ? st->iters_run : 0;
  if (asked_out) *asked_out = st ? st->iters_asked : 0;
  return RST_OK;
}

extern "C" int32_t rst_icp3d_cache_stats(rst_ctx* c, uint64_t* searched_out, uint64_t* queried_out) {
  if (!c) return RST_ERR_INVALID_ARG;
  void (**free_fn)(void*) = nullptr;
  void** slot = rst::ctx_ext_slot(c, &free_fn);
  const Icp3dState* st = static_cast<const Icp3dState*>(*slot);
  if (searched_out) *searched_out = st ? st->searched : 0;
  if (queried_out) *queried_out = st ? st->queried : 0;
  return RST_OK;
}

extern "C" int32_t rst_icp3d_pairs(rst_ctx* c, const rst_cloud* src, const rst_cloud* dst, int32_t n_pairs, int32_t max_iter,
                                   float grid_cell, float* poses_inout, rst_icp3d_result* results, int32_t* nbrs_out,
                                   float* weights_out) {
  if (!c) return RST_ERR_INVALID_ARG;
  auto fail = [&](int code, const std::string& m) { rst::ctx_set_error(c, m); return code; };
  if (!src || !dst || !poses_inout || n_pairs < 0 || max_iter < 0) return fail(RST_ERR_INVALID_ARG, "null clouds/poses or negative count");
  if (n_pairs == 0) return RST_OK;
#define ICP_CUDA(expr)                                                                   \
  do {                                                                                   \
    cudaError_t e_ = (expr);                                                             \
    if (e_ != cudaSuccess) return fail(RST_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e_)); \
  } while (0)
  ICP_CUDA(cudaSetDevice(rst::ctx_device(c)));
  cudaStream_t stream = rst::ctx_stream(c);
  void (**free_fn)(void*) = nullptr;
  void** slot = rst::ctx_ext_slot(c, &free_fn);
  if (!*slot) { *slot = new Icp3dState(); *free_fn = icp3d_free; }
  Icp3dState* st = static_cast<Icp3dState*>(*slot);
  st->last_frames = 0;  // the arena is about to be re-laid out

  // arena layout (same offsets on host staging and device for the uploaded part)
  size_t off = 0;
  const size_t o_desc = off; off = align_up(off + sizeof(PairDesc) * n_pairs);
  const size_t o_pose = off; off = align_up(off + sizeof(float) * 16 * n_pairs);
  std::vector<size_t> o_src(n_pairs), o_dst(n_pairs);
  size_t n_src_total = 0;
  for (int i = 0; i < n_pairs; ++i) {
    if (src[i].n < 0 || dst[i].n < 0 || (src[i].n > 0 && !src[i].xyz) || (dst[i].n > 0 && !dst[i].xyz))
      return fail(RST_ERR_INVALID_ARG, "bad cloud");
    o_src[i] = off; off = align_up(off + sizeof(float) * 3 * (size_t)src[i].n);
    o_dst[i] = off; off = align_up(off + sizeof(float) * 3 * (size_t)dst[i].n);
    n_src_total += (size_t)src[i].n;
  }
  const size_t upload_bytes = off;
  const size_t o_res = off; off = align_up(off + sizeof(rst_icp3d_result) * n_pairs);
  const size_t o_stat = off; off = align_up(off + sizeof(unsigned long long) * 4 * n_pairs);
  std::vector<size_t> o_nbr(n_pairs), o_w(n_pairs);
  const size_t o_nbr0 = off;
  for (int i = 0; i < n_pairs; ++i) { o_nbr[i] = off; off += sizeof(int) * (size_t)src[i].n; }
  off = align_up(off);
  const size_t o_w0 = off;
  for (int i = 0; i < n_pairs; ++i) { o_w[i] = off; off += sizeof(float) * (size_t)src[i].n; }
  off = align_up(off);
  const size_t download_end = off;
  std::vector<size_t> o_cs(n_pairs), o_cf(n_pairs), o_sorted(n_pairs), o_sl(n_pairs), o_qd(n_pairs), o_queue(n_pairs);
  for (int i = 0; i < n_pairs; ++i) {
    o_cs[i] = off; off = align_up(off + sizeof(int) * (kCellCap + 1));
    o_cf[i] = off; off = align_up(off + sizeof(int) * kCellCap);
    o_sorted[i] = off; off = align_up(off + sizeof(float4) * (size_t)dst[i].n);
    o_sl[i] = off; off = align_up(off + sizeof(float4) * (size_t)src[i].n);
    o_qd[i] = off; off = align_up(off + sizeof(float4) * (size_t)src[i].n);
    o_queue[i] = off; off = align_up(off + sizeof(int) * ((size_t)src[i].n + 16 * kThreads));
  }
  const size_t total = off;
  if (st->d_bytes < total) {
    ICP_CUDA(cudaStreamSynchronize(stream));
    cudaFree(st->d_arena); st->d_arena = nullptr; st->d_bytes = 0;
    ICP_CUDA(cudaMalloc(&st->d_arena, total));
    st->d_bytes = total;
  }
  if (st->h_bytes < download_end) {
    ICP_CUDA(cudaStreamSynchronize(stream));
    cudaFreeHost(st->h_arena); st->h_arena = nullptr; st->h_bytes = 0;
    ICP_CUDA(cudaMallocHost(&st->h_arena, download_end));
    st->h_bytes = download_end;
  }
  char* H = static_cast<char*>(st->h_arena);
  char* D = static_cast<char*>(st->d_arena);
  PairDesc* hd = reinterpret_cast<PairDesc*>(H + o_desc);
  for (int i = 0; i < n_pairs; ++i) {
    std::memcpy(H + o_src[i], src[i].xyz, sizeof(float) * 3 * (size_t)src[i].n);
    std::memcpy(H + o_dst[i], dst[i].xyz, sizeof(float) * 3 * (size_t)dst[i].n);
    PairDesc d;
    d.src = reinterpret_cast<const float*>(D + o_src[i]); d.dst = reinterpret_cast<const float*>(D + o_dst[i]);
    d.n = src[i].n; d.m = dst[i].n; d.n_ptr = nullptr; d.m_ptr = nullptr;
    d.cell_start = reinterpret_cast<int*>(D + o_cs[i]); d.cell_fill = reinterpret_cast<int*>(D + o_cf[i]);
    d.sorted = reinterpret_cast<float4*>(D + o_sorted[i]);
    d.nbr = reinterpret_cast<int*>(D + o_nbr[i]); d.w = reinterpret_cast<float*>(D + o_w[i]);
    d.sl = reinterpret_cast<float4*>(D + o_sl[i]); d.qd = reinterpret_cast<float4*>(D + o_qd[i]);
    d.queue = reinterpret_cast<int*>(D + o_queue[i]);
    d.pose = reinterpret_cast<float*>(D + o_pose) + 16 * i;
    d.res = reinterpret_cast<rst_icp3d_result*>(D + o_res) + i;
    d.stat = reinterpret_cast<unsigned long long*>(D + o_stat) + 4 * i;
    hd[i] = d;
  }
  std::memcpy(H + o_pose, poses_inout, sizeof(float) * 16 * n_pairs);
  ICP_CUDA(cudaMemcpyAsync(D, H, upload_bytes, cudaMemcpyHostToDevice, stream));
  ICP_CUDA(cudaMemsetAsync(D + o_res, 0, o_nbr0 - o_res, stream));   // results and cache statistics
  ICP_CUDA(launch_icp3d(reinterpret_cast<const PairDesc*>(D + o_desc), n_pairs, max_iter, grid_cell, st->icp3d_cluster, st->cache, st->skip_fixed, stream));
  rst::ctx_count_launches(c, 1);
  ICP_CUDA(cudaMemcpyAsync(H + o_pose, D + o_pose, sizeof(float) * 16 * n_pairs, cudaMemcpyDeviceToHost, stream));
  const bool want_corr = nbrs_out || weights_out;
  ICP_CUDA(cudaMemcpyAsync(H + o_res, D + o_res, (want_corr ? download_end : o_nbr0) - o_res, cudaMemcpyDeviceToHost, stream));
  ICP_CUDA(cudaStreamSynchronize(stream));
  std::memcpy(poses_inout, H + o_pose, sizeof(float) * 16 * n_pairs);
  if (results) std::memcpy(results, H + o_res, sizeof(rst_icp3d_result) * n_pairs);
  sum_stats(st, reinterpret_cast<const unsigned long long*>(H + o_stat), n_pairs);
  if (nbrs_out) std::memcpy(nbrs_out, H + o_nbr0, sizeof(int) * n_src_total);
  if (weights_out) std::memcpy(weights_out, H + o_w0, sizeof(float) * n_src_total);
#undef ICP_CUDA
  return RST_OK;
}

/* Depth frames in: the reference caller's whole per-pair sequence on the device
 * (rs_replay_app.cpp:229,246-251): back-project -> RemoveNans -> DownsampleVoxel(voxel) -> AlignIcp3d.
 * `frames` are the unique frames (host, dense or strided); pair i aligns frames[src_idx[i]] onto
 * frames[dst_idx[i]]. counts_out (nullable): number of points of every frame's cloud. */
extern "C" int32_t rst_icp3d_depth(rst_ctx* c, const rst_frame* frames, int32_t n_frames, const int32_t* src_idx,
                                   const int32_t* dst_idx, int32_t n_pairs, const rst_intrinsics* intr, float depth_scale,
                                   float voxel, int32_t max_iter, float grid_cell, float* poses_inout,
                                   rst_icp3d_result* results, int32_t* counts_out) {
  if (!c) return RST_ERR_INVALID_ARG;
  auto fail = [&](int code, const std::string& m) { rst::ctx_set_error(c, m); return code; };
  if (!frames || !src_idx || !dst_idx || !intr || !poses_inout || n_frames < 1 || n_pairs < 0 || max_iter < 0)
    return fail(RST_ERR_INVALID_ARG, "null argument or negative count");
  if (!(depth_scale > 0.f) || !(intr->fx > 0.f) || !(intr->fy > 0.f)) return fail(RST_ERR_INVALID_ARG, "depth_scale and focal lengths must be positive");
  const int w = frames[0].width, h = frames[0].height;
  if (w < 1 || h < 1 || (int64_t)w * h > (1 << 24)) return fail(RST_ERR_INVALID_ARG, "bad frame size");
  for (int i = 0; i < n_frames; ++i)
    if (!frames[i].depth || frames[i].width != w || frames[i].height != h || frames[i].depth_stride_bytes < 2 * w)
      return fail(RST_ERR_INVALID_ARG, "frames differ in size / null depth / bad stride");
  for (int i = 0; i < n_pairs; ++i)
    if (src_idx[i] < 0 || src_idx[i] >= n_frames || dst_idx[i] < 0 || dst_idx[i] >= n_frames) return fail(RST_ERR_INVALID_ARG, "frame index out of range");
#define ICP_CUDA(expr)                                                                   \
  do {                                                                                   \
    cudaError_t e_ = (expr);                                                             \
    if (e_ != cudaSuccess) return fail(RST_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e_)); \
  } while (0)
  ICP_CUDA(cudaSetDevice(rst::ctx_device(c)));
  cudaStream_t stream = rst::ctx_stream(c);
  void (**free_fn)(void*) = nullptr;
  void** slot = rst::ctx_ext_slot(c, &free_fn);
  if (!*slot) { *slot = new Icp3dState(); *free_fn = icp3d_free; }
  Icp3dState* st = static_cast<Icp3dState*>(*slot);

  const size_t npx = (size_t)w * h;
  uint32_t cap = 1;
  while (cap < 2 * npx) cap <<= 1;  // load factor <= 0.5 even if every pixel is its own voxel
  const bool decimate = voxel > 0.f;
  size_t off = 0;
  const size_t o_cdesc = off; off = align_up(off + sizeof(CloudifyDesc) * n_frames);
  const size_t o_pdesc = off; off = align_up(off + sizeof(PairDesc) * (size_t)(n_pairs > 0 ? n_pairs : 1));
  const size_t o_pose = off; off = align_up(off + sizeof(float) * 16 * (size_t)(n_pairs > 0 ? n_pairs : 1));
  const size_t upload_bytes = off;
  const size_t o_res = off; off = align_up(off + sizeof(rst_icp3d_result) * (size_t)(n_pairs > 0 ? n_pairs : 1));
  const size_t o_cnt = off; off = align_up(off + sizeof(int) * n_frames);
  const size_t o_stat = off; off = align_up(off + sizeof(unsigned long long) * 4 * (size_t)(n_pairs > 0 ? n_pairs : 1));
  const size_t download_end = off;
  const size_t o_depth = off; off = align_up(off + npx * 2 * n_frames);
  const size_t o_cloud = off; off = align_up(off + npx * 12 * n_frames);
  const size_t o_keys = off; off = align_up(off + (decimate ? (size_t)cap * 8 * n_frames : 0));
  const size_t o_vals = off; off = align_up(off + (decimate ? (size_t)cap * 4 * n_frames : 0));
  const int n_seg = (int)((npx + kSegPx - 1) / kSegPx);
  const size_t o_seg = off; off = align_up(off + sizeof(int) * (size_t)n_seg * n_frames);
  std::vector<size_t> o_cs(n_pairs), o_cf(n_pairs), o_sorted(n_pairs), o_nbr(n_pairs), o_w(n_pairs), o_sl(n_pairs), o_qd(n_pairs),
      o_queue(n_pairs);
  for (int i = 0; i < n_pairs; ++i) {
    o_cs[i] = off; off = align_up(off + sizeof(int) * (kCellCap + 1));
    o_cf[i] = off; off = align_up(off + sizeof(int) * kCellCap);
    o_sorted[i] = off; off = align_up(off + sizeof(float4) * npx);
    o_nbr[i] = off; off = align_up(off + sizeof(int) * npx);
    o_w[i] = off; off = align_up(off + sizeof(float) * npx);
    o_sl[i] = off; off = align_up(off + sizeof(float4) * npx);
    o_qd[i] = off; off = align_up(off + sizeof(float4) * npx);
    o_queue[i] = off; off = align_up(off + sizeof(int) * (npx + 16 * kThreads));
  }
  const size_t total = off;
  if (st->d_bytes < total) {
    ICP_CUDA(cudaStreamSynchronize(stream));
    cudaFree(st->d_arena); st->d_arena = nullptr; st->d_bytes = 0;
    ICP_CUDA(cudaMalloc(&st->d_arena, total));
    st->d_bytes = total;
  }
  if (st->h_bytes < download_end) {
    ICP_CUDA(cudaStreamSynchronize(stream));
    cudaFreeHost(st->h_arena); st->h_arena = nullptr; st->h_bytes = 0;
    ICP_CUDA(cudaMallocHost(&st->h_arena, download_end));
    st->h_bytes = download_end;
  }
  st->last_frames = n_frames; st->last_npx = npx; st->last_cloud_off = o_cloud;
  char* H = static_cast<char*>(st->h_arena);
  char* D = static_cast<char*>(st->d_arena);
  CloudifyDesc* cd = reinterpret_cast<CloudifyDesc*>(H + o_cdesc);
  for (int f = 0; f < n_frames; ++f) {
    cd[f].depth = reinterpret_cast<const uint16_t*>(D + o_depth + npx * 2 * f);
    cd[f].keys = reinterpret_cast<unsigned long long*>(D + o_keys + (size_t)cap * 8 * f);
    cd[f].vals = reinterpret_cast<int*>(D + o_vals + (size_t)cap * 4 * f);
    cd[f].cloud = reinterpret_cast<float*>(D + o_cloud + npx * 12 * f);
    cd[f].count = reinterpret_cast<int*>(D + o_cnt) + f;
    cd[f].seg = reinterpret_cast<int*>(D + o_seg) + (size_t)n_seg * f;
  }
  PairDesc* pd = reinterpret_cast<PairDesc*>(H + o_pdesc);
  for (int i = 0; i < n_pairs; ++i) {
    PairDesc d;
    d.src = cd[src_idx[i]].cloud; d.dst = cd[dst_idx[i]].cloud;
    d.n = 0; d.m = 0; d.n_ptr = cd[src_idx[i]].count; d.m_ptr = cd[dst_idx[i]].count;
    d.cell_start = reinterpret_cast<int*>(D + o_cs[i]); d.cell_fill = reinterpret_cast<int*>(D + o_cf[i]);
    d.sorted = reinterpret_cast<float4*>(D + o_sorted[i]);
    d.nbr = reinterpret_cast<int*>(D + o_nbr[i]); d.w = reinterpret_cast<float*>(D + o_w[i]);
    d.sl = reinterpret_cast<float4*>(D + o_sl[i]); d.qd = reinterpret_cast<float4*>(D + o_qd[i]);
    d.queue = reinterpret_cast<int*>(D + o_queue[i]);
    d.pose = reinterpret_cast<float*>(D + o_pose) + 16 * i;
    d.res = reinterpret_cast<rst_icp3d_result*>(D + o_res) + i;
    d.stat = reinterpret_cast<unsigned long long*>(D + o_stat) + 4 * i;
    pd[i] = d;
  }
  if (n_pairs > 0) std::memcpy(H + o_pose, poses_inout, sizeof(float) * 16 * n_pairs);
  ICP_CUDA(cudaMemcpyAsync(D, H, upload_bytes, cudaMemcpyHostToDevice, stream));
  // frames go up in chunks on the copy stream; the depth -> cloud kernels of chunk k run under the copy of chunk k + 1
  const int n_chunks = n_frames >= 4 * kUploadChunks ? kUploadChunks : 1;
  if (n_chunks > 1 && !st->copy) {
    ICP_CUDA(cudaStreamCreateWithFlags(&st->copy, cudaStreamNonBlocking));
    for (cudaEvent_t& e : st->ev) ICP_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
  }
  if (n_chunks > 1) {   // the arena may still be read by this stream's earlier work
    ICP_CUDA(cudaEventRecord(st->ev[kUploadChunks], stream));
    ICP_CUDA(cudaStreamWaitEvent(st->copy, st->ev[kUploadChunks], 0));
  }
  if (decimate) {
    ICP_CUDA(cudaMemsetAsync(D + o_keys, 0, (size_t)cap * 8 * n_frames, stream));
    ICP_CUDA(cudaMemsetAsync(D + o_vals, 0x7f, (size_t)cap * 4 * n_frames, stream));  // 0x7f7f7f7f > any pixel index
  }
  ICP_CUDA(cudaMemsetAsync(D + o_res, 0, sizeof(rst_icp3d_result) * (size_t)(n_pairs > 0 ? n_pairs : 1), stream));
  ICP_CUDA(cudaMemsetAsync(D + o_stat, 0, sizeof(unsigned long long) * 4 * (size_t)(n_pairs > 0 ? n_pairs : 1), stream));
  for (int ch = 0; ch < n_chunks; ++ch) {
    const int f0 = (int)((long long)n_frames * ch / n_chunks), f1 = (int)((long long)n_frames * (ch + 1) / n_chunks), nf = f1 - f0;
    cudaStream_t up = n_chunks > 1 ? st->copy : stream;
    for (int f = f0; f < f1; ++f)
      ICP_CUDA(cudaMemcpy2DAsync(D + o_depth + npx * 2 * f, (size_t)w * 2, frames[f].depth, (size_t)frames[f].depth_stride_bytes,
                                 (size_t)w * 2, (size_t)h, cudaMemcpyHostToDevice, up));
    if (n_chunks > 1) {
      ICP_CUDA(cudaEventRecord(st->ev[ch], st->copy));
      ICP_CUDA(cudaStreamWaitEvent(stream, st->ev[ch], 0));
    }
    const CloudifyDesc* dd = reinterpret_cast<const CloudifyDesc*>(D + o_cdesc) + f0;
    const dim3 gpx((unsigned)((npx + 255) / 256), nf), gseg(n_seg, nf);
    if (decimate) k_cloudify_insert<<<gpx, 256, 0, stream>>>(dd, w, h, intr->fx, intr->fy, intr->cx, intr->cy, depth_scale, voxel, cap - 1);
    k_cloudify_segment<false><<<gseg, kSegThreads, 0, stream>>>(dd, w, h, intr->fx, intr->fy, intr->cx, intr->cy, depth_scale, voxel, cap - 1);
    k_cloudify_scan<<<nf, kThreads, 0, stream>>>(dd, n_seg);
    k_cloudify_segment<true><<<gseg, kSegThreads, 0, stream>>>(dd, w, h, intr->fx, intr->fy, intr->cx, intr->cy, depth_scale, voxel, cap - 1);
    ICP_CUDA(cudaGetLastError());
    rst::ctx_count_launches(c, decimate ? 4 : 3);
  }
  if (n_pairs > 0) {
    ICP_CUDA(launch_icp3d(reinterpret_cast<const PairDesc*>(D + o_pdesc), n_pairs, max_iter, grid_cell, st->icp3d_cluster, st->cache, st->skip_fixed, stream));
    rst::ctx_count_launches(c, 1);
    ICP_CUDA(cudaMemcpyAsync(H + o_pose, D + o_pose, sizeof(float) * 16 * n_pairs, cudaMemcpyDeviceToHost, stream));
  }
  ICP_CUDA(cudaMemcpyAsync(H + o_res, D + o_res, download_end - o_res, cudaMemcpyDeviceToHost, stream));
  ICP_CUDA(cudaStreamSynchronize(stream));
  if (n_pairs > 0) std::memcpy(poses_inout, H + o_pose, sizeof(float) * 16 * n_pairs);
  if (results && n_pairs > 0) std::memcpy(results, H + o_res, sizeof(rst_icp3d_result) * n_pairs);
  if (counts_out) std::memcpy(counts_out, H + o_cnt, sizeof(int) * n_frames);
  sum_stats(st, reinterpret_cast<const unsigned long long*>(H + o_stat), n_pairs);
#undef ICP_CUDA
  return RST_OK;
}

/* Reads back the first n_points points of the cloud of frame `frame_index` produced by the last
 * rst_icp3d_depth call of this context (parity tests). */
extern "C" int32_t rst_icp3d_read_cloud(rst_ctx* c, int32_t frame_index, float* xyz_out, int32_t n_points) {
  if (!c) return RST_ERR_INVALID_ARG;
  void (**free_fn)(void*) = nullptr;
  void** slot = rst::ctx_ext_slot(c, &free_fn);
  Icp3dState* st = static_cast<Icp3dState*>(*slot);
  if (!st || !xyz_out || frame_index < 0 || frame_index >= st->last_frames || n_points < 0 || (size_t)n_points > st->last_npx) {
    rst::ctx_set_error(c, "rst_icp3d_read_cloud: no rst_icp3d_depth result for that frame / bad count");
    return RST_ERR_INVALID_ARG;
  }
  if (cudaSetDevice(rst::ctx_device(c)) != cudaSuccess) return RST_ERR_CUDA;
  const char* src = static_cast<char*>(st->d_arena) + st->last_cloud_off + st->last_npx * 12 * (size_t)frame_index;
  if (cudaMemcpy(xyz_out, src, sizeof(float) * 3 * (size_t)n_points, cudaMemcpyDeviceToHost) != cudaSuccess) return RST_ERR_CUDA;
  return RST_OK;
}

/* SolveKabsch(src, dst, indices, weights, &xfm)  align_icp.cpp:18-71 on the device. pairs: n_pairs x
 * (src index, dst index), weights nullable (= empty vector), all HOST pointers; pose_out: 16 floats,
 * column-major. Returns RST_OK and *ok_out = 0 for < 3 points in either cloud (:23-25). */
extern "C" int32_t rst_solve_kabsch(rst_ctx* c, const rst_cloud* src, const rst_cloud* dst, const int32_t* pairs, int32_t n_pairs,
                                    const float* weights, float* pose_out, int32_t* ok_out) {
  if (!c) return RST_ERR_INVALID_ARG;
  auto fail = [&](int code, const std::string& m) { rst::ctx_set_error(c, m); return code; };
  if (!src || !dst || !pose_out || !ok_out || n_pairs < 0 || (n_pairs > 0 && !pairs)) return fail(RST_ERR_INVALID_ARG, "null argument");
  *ok_out = 0;
  if (src->n < 3 || dst->n < 3) return RST_OK;
  if (n_pairs == 0) return fail(RST_ERR_INVALID_ARG, "no index pairs");
  for (int i = 0; i < n_pairs; ++i)
    if (pairs[2 * i] < 0 || pairs[2 * i] >= src->n || pairs[2 * i + 1] < 0 || pairs[2 * i + 1] >= dst->n)
      return fail(RST_ERR_INVALID_ARG, "pair index out of range");
#define ICP_CUDA(expr)                                                                   \
  do {                                                                                   \
    cudaError_t e_ = (expr);                                                             \
    if (e_ != cudaSuccess) return fail(RST_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e_)); \
  } while (0)
  ICP_CUDA(cudaSetDevice(rst::ctx_device(c)));
  cudaStream_t stream = rst::ctx_stream(c);
  void (**free_fn)(void*) = nullptr;
  void** slot = rst::ctx_ext_slot(c, &free_fn);
  if (!*slot) { *slot = new Icp3dState(); *free_fn = icp3d_free; }
  Icp3dState* st = static_cast<Icp3dState*>(*slot);
  st->last_frames = 0;
  size_t off = 0;
  const size_t o_pose = off; off = align_up(off + 64);
  const size_t o_src = off; off = align_up(off + sizeof(float) * 3 * (size_t)src->n);
  const size_t o_dst = off; off = align_up(off + sizeof(float) * 3 * (size_t)dst->n);
  const size_t o_pairs = off; off = align_up(off + sizeof(int2) * (size_t)n_pairs);
  const size_t o_w = off; off = align_up(off + sizeof(float) * (size_t)n_pairs);
  const size_t total = off;
  if (st->d_bytes < total) {
    ICP_CUDA(cudaStreamSynchronize(stream));
    cudaFree(st->d_arena); st->d_arena = nullptr; st->d_bytes = 0;
    ICP_CUDA(cudaMalloc(&st->d_arena, total));
    st->d_bytes = total;
  }
  if (st->h_bytes < total) {
    ICP_CUDA(cudaStreamSynchronize(stream));
    cudaFreeHost(st->h_arena); st->h_arena = nullptr; st->h_bytes = 0;
    ICP_CUDA(cudaMallocHost(&st->h_arena, total));
    st->h_bytes = total;
  }
  char* H = static_cast<char*>(st->h_arena);
  char* D = static_cast<char*>(st->d_arena);
  std::memcpy(H + o_src, src->xyz, sizeof(float) * 3 * (size_t)src->n);
  std::memcpy(H + o_dst, dst->xyz, sizeof(float) * 3 * (size_t)dst->n);
  std::memcpy(H + o_pairs, pairs, sizeof(int2) * (size_t)n_pairs);
  if (weights) std::memcpy(H + o_w, weights, sizeof(float) * (size_t)n_pairs);
  ICP_CUDA(cudaMemcpyAsync(D + o_src, H + o_src, total - o_src, cudaMemcpyHostToDevice, stream));
  k_kabsch<<<1, kThreads, 0, stream>>>(reinterpret_cast<const float*>(D + o_src), reinterpret_cast<const float*>(D + o_dst),
                                       reinterpret_cast<const int2*>(D + o_pairs), n_pairs,
                                       weights ? reinterpret_cast<const float*>(D + o_w) : nullptr, reinterpret_cast<float*>(D + o_pose));
  ICP_CUDA(cudaGetLastError());
  rst::ctx_count_launches(c, 1);
  ICP_CUDA(cudaMemcpyAsync(H + o_pose, D + o_pose, 64, cudaMemcpyDeviceToHost, stream));
  ICP_CUDA(cudaStreamSynchronize(stream));
  std::memcpy(pose_out, H + o_pose, 64);
  *ok_out = 1;
#undef ICP_CUDA
  return RST_OK;
}

/* ComputeNormals(cloud, tree, k, &normals) + OrientNormals(cloud, viewpoint, &normals)
 * (point_cloud_utils.cpp:176-216) on the device. k (2..32) counts the point itself, as in the reference
 * (rs_align_app.cpp:25 uses 16). normals_out: n x 3 floats, HOST memory. */
extern "C" int32_t rst_cloud_normals(rst_ctx* c, const rst_cloud* cloud, int32_t k, const float* viewpoint, float grid_cell, float* normals_out) {
  if (!c) return RST_ERR_INVALID_ARG;
  auto fail = [&](int code, const std::string& m) { rst::ctx_set_error(c, m); return code; };
  if (!cloud || !viewpoint || !normals_out || cloud->n < 0 || (cloud->n > 0 && !cloud->xyz)) return fail(RST_ERR_INVALID_ARG, "null argument");
  if (k < 2 || k > 32) return fail(RST_ERR_INVALID_ARG, "k must be in [2, 32]");
  if (cloud->n == 0) return RST_OK;
#define ICP_CUDA(expr)                                                                   \
  do {                                                                                   \
    cudaError_t e_ = (expr);                                                             \
    if (e_ != cudaSuccess) return fail(RST_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e_)); \
  } while (0)
  ICP_CUDA(cudaSetDevice(rst::ctx_device(c)));
  cudaStream_t stream = rst::ctx_stream(c);
  void (**free_fn)(void*) = nullptr;
  void** slot = rst::ctx_ext_slot(c, &free_fn);
  if (!*slot) { *slot = new Icp3dState(); *free_fn = icp3d_free; }
  Icp3dState* st = static_cast<Icp3dState*>(*slot);
  st->last_frames = 0;
  const size_t n = (size_t)cloud->n;
  size_t off = 0;
  const size_t o_pts = off; off = align_up(off + sizeof(float) * 3 * n);
  const size_t upload = off;
  const size_t o_nrm = off; off = align_up(off + sizeof(float) * 3 * n);
  const size_t host_end = off;
  const size_t o_grid = off; off = align_up(off + sizeof(Grid));
  const size_t o_cs = off; off = align_up(off + sizeof(int) * (kCellCap + 1));
  const size_t o_cf = off; off = align_up(off + sizeof(int) * kCellCap);
  const size_t o_sorted = off; off = align_up(off + sizeof(float4) * n);
  const size_t total = off;
  if (st->d_bytes < total) {
    ICP_CUDA(cudaStreamSynchronize(stream));
    cudaFree(st->d_arena); st->d_arena = nullptr; st->d_bytes = 0;
    ICP_CUDA(cudaMalloc(&st->d_arena, total));
    st->d_bytes = total;
  }
  if (st->h_bytes < host_end) {
    ICP_CUDA(cudaStreamSynchronize(stream));
    cudaFreeHost(st->h_arena); st->h_arena = nullptr; st->h_bytes = 0;
    ICP_CUDA(cudaMallocHost(&st->h_arena, host_end));
    st->h_bytes = host_end;
  }
  char* H = static_cast<char*>(st->h_arena);
  char* D = static_cast<char*>(st->d_arena);
  std::memcpy(H + o_pts, cloud->xyz, sizeof(float) * 3 * n);
  ICP_CUDA(cudaMemcpyAsync(D, H, upload, cudaMemcpyHostToDevice, stream));
  k_grid_build<<<1, kThreads, 0, stream>>>(reinterpret_cast<const float*>(D + o_pts), (int)n, grid_cell, reinterpret_cast<int*>(D + o_cs),
                                           reinterpret_cast<int*>(D + o_cf), reinterpret_cast<float4*>(D + o_sorted),
                                           reinterpret_cast<Grid*>(D + o_grid));
  k_normals<<<(unsigned)((n + 127) / 128), 128, 0, stream>>>(reinterpret_cast<const Grid*>(D + o_grid), reinterpret_cast<const int*>(D + o_cs),
                                                             reinterpret_cast<const float4*>(D + o_sorted),
                                                             reinterpret_cast<const float*>(D + o_pts), (int)n, k, viewpoint[0],
                                                             viewpoint[1], viewpoint[2], reinterpret_cast<float*>(D + o_nrm));
  ICP_CUDA(cudaGetLastError());
  rst::ctx_count_launches(c, 2);
  ICP_CUDA(cudaMemcpyAsync(H + o_nrm, D + o_nrm, sizeof(float) * 3 * n, cudaMemcpyDeviceToHost, stream));
  ICP_CUDA(cudaStreamSynchronize(stream));
  std::memcpy(normals_out, H + o_nrm, sizeof(float) * 3 * n);
#undef ICP_CUDA
  return RST_OK;
}

// ------------------------------------------------------------------------------------------------
// Cloud utilities of rs_tracker/common (point_cloud_utils.hpp) for callers that hold clouds — which is every
// caller of the reference (rs_replay_app.cpp:229,246-247; rs_tracker.cpp:60-87). HOST pointers in and out.
// ------------------------------------------------------------------------------------------------
namespace {

// one call's view of the context's grow-only arenas: offsets are taken first, then `commit` sizes the arenas
struct CloudCall {
  rst_ctx* c;
  Icp3dState* st = nullptr;
  cudaStream_t stream = nullptr;
  size_t off = 0;
  char* H = nullptr;
  char* D = nullptr;
  explicit CloudCall(rst_ctx* ctx) : c(ctx) {}
  size_t take(size_t bytes) { const size_t o = off; off = align_up(off + bytes); return o; }
  int fail(int code, const std::string& m) { rst::ctx_set_error(c, m); return code; }
  int cuda(cudaError_t e, const char* what) {
    if (e == cudaSuccess) return RST_OK;
    return fail(RST_ERR_CUDA, std::string(what) + ": " + cudaGetErrorString(e));
  }
  int begin() {
    int rc = cuda(cudaSetDevice(rst::ctx_device(c)), "cudaSetDevice");
    if (rc != RST_OK) return rc;
    stream = rst::ctx_stream(c);
    void (**free_fn)(void*) = nullptr;
    void** slot = rst::ctx_ext_slot(c, &free_fn);
    if (!*slot) { *slot = new Icp3dState(); *free_fn = icp3d_free; }
    st = static_cast<Icp3dState*>(*slot);
    st->last_frames = 0;
    return RST_OK;
  }
  int commit(size_t host_bytes) {   // device arena >= off, pinned arena >= host_bytes
    int rc;
    if (st->d_bytes < off) {
      if ((rc = cuda(cudaStreamSynchronize(stream), "sync")) != RST_OK) return rc;
      cudaFree(st->d_arena); st->d_arena = nullptr; st->d_bytes = 0;
      if ((rc = cuda(cudaMalloc(&st->d_arena, off), "cudaMalloc")) != RST_OK) return rc;
      st->d_bytes = off;
    }
    if (st->h_bytes < host_bytes) {
      if ((rc = cuda(cudaStreamSynchronize(stream), "sync")) != RST_OK) return rc;
      cudaFreeHost(st->h_arena); st->h_arena = nullptr; st->h_bytes = 0;
      if ((rc = cuda(cudaMallocHost(&st->h_arena, host_bytes), "cudaMallocHost")) != RST_OK) return rc;
      st->h_bytes = host_bytes;
    }
    H = static_cast<char*>(st->h_arena);
    D = static_cast<char*>(st->d_arena);
    return RST_OK;
  }
};

#define CLOUD_TRY(expr)                              \
  do {                                               \
    const int rc_ = (expr);                          \
    if (rc_ != RST_OK) return rc_;                   \
  } while (0)

bool cloud_ok(const rst_cloud* cl) { return cl && cl->n >= 0 && (cl->n == 0 || cl->xyz); }

}  // namespace

/* void FindCorrespondences(tree, source, &indices, &squared_distances)  point_cloud_utils.cpp:70-90 */
extern "C" int32_t rst_find_correspondences(rst_ctx* c, const rst_cloud* target, const rst_cloud* source, float grid_cell,
                                            int32_t* indices_out, float* sq_dist_out) {
  if (!c) return RST_ERR_INVALID_ARG;
  CloudCall k(c);
  if (!cloud_ok(target) || !cloud_ok(source) || !indices_out || !sq_dist_out) return k.fail(RST_ERR_INVALID_ARG, "null argument / bad cloud");
  if (source->n == 0) return RST_OK;
  if (target->n == 0) return k.fail(RST_ERR_INVALID_ARG, "the target cloud is empty");
  CLOUD_TRY(k.begin());
  const size_t m = (size_t)target->n, n = (size_t)source->n;
  const size_t o_tgt = k.take(12 * m), o_src = k.take(12 * n);
  const size_t upload = k.off;
  const size_t o_idx = k.take(4 * n), o_d2 = k.take(4 * n);
  const size_t host_end = k.off;
  const size_t o_grid = k.take(sizeof(Grid)), o_cs = k.take(sizeof(int) * (kCellCap + 1)), o_cf = k.take(sizeof(int) * kCellCap);
  const size_t o_sorted = k.take(sizeof(float4) * m);
  CLOUD_TRY(k.commit(host_end));
  std::memcpy(k.H + o_tgt, target->xyz, 12 * m);
  std::memcpy(k.H + o_src, source->xyz, 12 * n);
  CLOUD_TRY(k.cuda(cudaMemcpyAsync(k.D, k.H, upload, cudaMemcpyHostToDevice, k.stream), "H2D"));
  k_grid_build<<<1, kThreads, 0, k.stream>>>(reinterpret_cast<const float*>(k.D + o_tgt), (int)m, grid_cell, reinterpret_cast<int*>(k.D + o_cs),
                                             reinterpret_cast<int*>(k.D + o_cf), reinterpret_cast<float4*>(k.D + o_sorted),
                                             reinterpret_cast<Grid*>(k.D + o_grid));
  k_nn_query<<<(unsigned)((n + 127) / 128), 128, 0, k.stream>>>(reinterpret_cast<const Grid*>(k.D + o_grid), reinterpret_cast<const int*>(k.D + o_cs),
                                                                  reinterpret_cast<const float4*>(k.D + o_sorted),
                                                                  reinterpret_cast<const float*>(k.D + o_src), (int)n,
                                                                  reinterpret_cast<int*>(k.D + o_idx), reinterpret_cast<float*>(k.D + o_d2));
  CLOUD_TRY(k.cuda(cudaGetLastError(), "launch"));
  rst::ctx_count_launches(c, 2);
  CLOUD_TRY(k.cuda(cudaMemcpyAsync(k.H + o_idx, k.D + o_idx, host_end - o_idx, cudaMemcpyDeviceToHost, k.stream), "D2H"));
  CLOUD_TRY(k.cuda(cudaStreamSynchronize(k.stream), "sync"));
  std::memcpy(indices_out, k.H + o_idx, 4 * n);
  std::memcpy(sq_dist_out, k.H + o_d2, 4 * n);
  return RST_OK;
}

/* KDTree3f (kdtree.hpp:11-99, types.hpp) as a device-resident handle: the search grid of one cloud, built once and
 * queried any number of times (the reference builds dst_tree once per AlignIcp3d call and queries it 128 x n times,
 * align_icp.cpp:163-167,112). */
struct rst_tree {
  int device = 0;
  int m = 0;
  char* d_mem = nullptr;
  size_t o_grid = 0, o_cs = 0, o_sorted = 0;
};

extern "C" int32_t rst_tree_create(rst_ctx* c, const rst_cloud* cloud, float grid_cell, rst_tree** tree_out) {
  if (!c) return RST_ERR_INVALID_ARG;
  CloudCall k(c);
  if (!cloud_ok(cloud) || !tree_out) return k.fail(RST_ERR_INVALID_ARG, "null argument / bad cloud");
  *tree_out = nullptr;
  if (cloud->n == 0) return k.fail(RST_ERR_INVALID_ARG, "the cloud is empty");
  CLOUD_TRY(k.begin());
  const size_t m = (size_t)cloud->n;
  const size_t o_stage = k.take(12 * m);
  CLOUD_TRY(k.commit(k.off));
  // the tree's own memory: points | grid | cell_start | cell_fill | sorted
  size_t off = 0;
  auto take = [&](size_t bytes) { const size_t o = off; off = align_up(off + bytes); return o; };
  const size_t t_pts = take(12 * m), t_grid = take(sizeof(Grid)), t_cs = take(sizeof(int) * (kCellCap + 1)), t_cf = take(sizeof(int) * kCellCap);
  const size_t t_sorted = take(sizeof(float4) * m);
  char* mem = nullptr;
  CLOUD_TRY(k.cuda(cudaMalloc(&mem, off), "cudaMalloc"));
  std::memcpy(k.H + o_stage, cloud->xyz, 12 * m);
  int rc = k.cuda(cudaMemcpyAsync(mem + t_pts, k.H + o_stage, 12 * m, cudaMemcpyHostToDevice, k.stream), "H2D");
  if (rc == RST_OK) {
    k_grid_build<<<1, kThreads, 0, k.stream>>>(reinterpret_cast<const float*>(mem + t_pts), (int)m, grid_cell, reinterpret_cast<int*>(mem + t_cs),
                                               reinterpret_cast<int*>(mem + t_cf), reinterpret_cast<float4*>(mem + t_sorted),
                                               reinterpret_cast<Grid*>(mem + t_grid));
    rc = k.cuda(cudaGetLastError(), "launch");
  }
  if (rc == RST_OK) { rst::ctx_count_launches(c, 1); rc = k.cuda(cudaStreamSynchronize(k.stream), "sync"); }
  if (rc != RST_OK) { cudaFree(mem); return rc; }
  rst_tree* t = new rst_tree();
  t->device = rst::ctx_device(c); t->m = (int)m; t->d_mem = mem;
  t->o_grid = t_grid; t->o_cs = t_cs; t->o_sorted = t_sorted;
  *tree_out = t;
  return RST_OK;
}

extern "C" void rst_tree_destroy(rst_tree* tree) {
  if (!tree) return;
  int prev = 0;
  const bool have_prev = cudaGetDevice(&prev) == cudaSuccess;
  if (cudaSetDevice(tree->device) == cudaSuccess) cudaFree(tree->d_mem);
  if (have_prev) cudaSetDevice(prev);
  delete tree;
}

extern "C" int32_t rst_tree_size(const rst_tree* tree) { return tree ? tree->m : 0; }

extern "C" int32_t rst_tree_query(rst_ctx* c, const rst_tree* tree, const float* queries_xyz, int32_t n_queries, int32_t k_nearest,
                                  int32_t* indices_out, float* sq_dist_out) {
  if (!c) return RST_ERR_INVALID_ARG;
  CloudCall k(c);
  if (!tree || n_queries < 0 || (n_queries > 0 && (!queries_xyz || !indices_out || !sq_dist_out)))
    return k.fail(RST_ERR_INVALID_ARG, "null argument");
  if (k_nearest < 1 || k_nearest > kMaxK) return k.fail(RST_ERR_INVALID_ARG, "k must be in [1, 33]");
  if (tree->device != rst::ctx_device(c)) return k.fail(RST_ERR_INVALID_ARG, "the tree lives on another device than the context");
  if (n_queries == 0) return RST_OK;
  CLOUD_TRY(k.begin());
  const size_t n = (size_t)n_queries, kk = (size_t)k_nearest;
  const size_t o_q = k.take(12 * n);
  const size_t upload = k.off;
  const size_t o_idx = k.take(4 * n * kk), o_d2 = k.take(4 * n * kk);
  const size_t host_end = k.off;
  CLOUD_TRY(k.commit(host_end));
  std::memcpy(k.H + o_q, queries_xyz, 12 * n);
  CLOUD_TRY(k.cuda(cudaMemcpyAsync(k.D, k.H, upload, cudaMemcpyHostToDevice, k.stream), "H2D"));
  const Grid* d_grid = reinterpret_cast<const Grid*>(tree->d_mem + tree->o_grid);
  const int* d_cs = reinterpret_cast<const int*>(tree->d_mem + tree->o_cs);
  const float4* d_sorted = reinterpret_cast<const float4*>(tree->d_mem + tree->o_sorted);
  if (k_nearest == 1)   // the search of FindCorrespondences / AlignIcp3d
    k_nn_query<<<(unsigned)((n + 127) / 128), 128, 0, k.stream>>>(d_grid, d_cs, d_sorted, reinterpret_cast<const float*>(k.D + o_q), (int)n,
                                                                    reinterpret_cast<int*>(k.D + o_idx), reinterpret_cast<float*>(k.D + o_d2));
  else
    k_knn_query<<<(unsigned)((n + 127) / 128), 128, 0, k.stream>>>(d_grid, d_cs, d_sorted, tree->m, reinterpret_cast<const float*>(k.D + o_q), (int)n,
                                                                     k_nearest, reinterpret_cast<int*>(k.D + o_idx), reinterpret_cast<float*>(k.D + o_d2));
  CLOUD_TRY(k.cuda(cudaGetLastError(), "launch"));
  rst::ctx_count_launches(c, 1);
  CLOUD_TRY(k.cuda(cudaMemcpyAsync(k.H + o_idx, k.D + o_idx, host_end - o_idx, cudaMemcpyDeviceToHost, k.stream), "D2H"));
  CLOUD_TRY(k.cuda(cudaStreamSynchronize(k.stream), "sync"));
  std::memcpy(indices_out, k.H + o_idx, 4 * n * kk);
  std::memcpy(sq_dist_out, k.H + o_d2, 4 * n * kk);
  return RST_OK;
}

/* void ComputeCovariances(tree, cloud, &covs, use_gicp)  point_cloud_utils.cpp:100-161 */
extern "C" int32_t rst_cloud_covariances(rst_ctx* c, const rst_cloud* cloud, int32_t use_gicp, float grid_cell, float* covs_out) {
  if (!c) return RST_ERR_INVALID_ARG;
  CloudCall k(c);
  if (!cloud_ok(cloud) || !covs_out) return k.fail(RST_ERR_INVALID_ARG, "null argument / bad cloud");
  if (cloud->n == 0) return RST_OK;
  CLOUD_TRY(k.begin());
  const size_t n = (size_t)cloud->n;
  const size_t o_pts = k.take(12 * n);
  const size_t upload = k.off;
  const size_t o_cov = k.take(36 * n);
  const size_t host_end = k.off;
  const size_t o_grid = k.take(sizeof(Grid)), o_cs = k.take(sizeof(int) * (kCellCap + 1)), o_cf = k.take(sizeof(int) * kCellCap);
  const size_t o_sorted = k.take(sizeof(float4) * n);
  CLOUD_TRY(k.commit(host_end));
  std::memcpy(k.H + o_pts, cloud->xyz, 12 * n);
  CLOUD_TRY(k.cuda(cudaMemcpyAsync(k.D, k.H, upload, cudaMemcpyHostToDevice, k.stream), "H2D"));
  const float* pts = reinterpret_cast<const float*>(k.D + o_pts);
  k_grid_build<<<1, kThreads, 0, k.stream>>>(pts, (int)n, grid_cell, reinterpret_cast<int*>(k.D + o_cs), reinterpret_cast<int*>(k.D + o_cf),
                                             reinterpret_cast<float4*>(k.D + o_sorted), reinterpret_cast<Grid*>(k.D + o_grid));
  k_covariances<<<(unsigned)((n + 127) / 128), 128, 0, k.stream>>>(reinterpret_cast<const Grid*>(k.D + o_grid), reinterpret_cast<const int*>(k.D + o_cs),
                                                                     reinterpret_cast<const float4*>(k.D + o_sorted), pts, (int)n, use_gicp ? 1 : 0,
                                                                     reinterpret_cast<float*>(k.D + o_cov));
  CLOUD_TRY(k.cuda(cudaGetLastError(), "launch"));
  rst::ctx_count_launches(c, 2);
  CLOUD_TRY(k.cuda(cudaMemcpyAsync(k.H + o_cov, k.D + o_cov, 36 * n, cudaMemcpyDeviceToHost, k.stream), "D2H"));
  CLOUD_TRY(k.cuda(cudaStreamSynchronize(k.stream), "sync"));
  std::memcpy(covs_out, k.H + o_cov, 36 * n);
  return RST_OK;
}

/* void ComputeCentroid(cloud, &centroid)  point_cloud_utils.cpp:92-98 */
extern "C" int32_t rst_cloud_centroid(rst_ctx* c, const rst_cloud* cloud, float* centroid_out) {
  if (!c) return RST_ERR_INVALID_ARG;
  CloudCall k(c);
  if (!cloud_ok(cloud) || !centroid_out) return k.fail(RST_ERR_INVALID_ARG, "null argument / bad cloud");
  if (cloud->n == 0) return k.fail(RST_ERR_INVALID_ARG, "the cloud is empty");   // the reference divides by zero here
  CLOUD_TRY(k.begin());
  const size_t n = (size_t)cloud->n;
  const size_t o_pts = k.take(12 * n);
  const size_t upload = k.off;
  const size_t o_cen = k.take(12);
  const size_t host_end = k.off;
  CLOUD_TRY(k.commit(host_end));
  std::memcpy(k.H + o_pts, cloud->xyz, 12 * n);
  CLOUD_TRY(k.cuda(cudaMemcpyAsync(k.D, k.H, upload, cudaMemcpyHostToDevice, k.stream), "H2D"));
  k_centroid<<<1, kThreads, 0, k.stream>>>(reinterpret_cast<const float*>(k.D + o_pts), (int)n, reinterpret_cast<float*>(k.D + o_cen));
  CLOUD_TRY(k.cuda(cudaGetLastError(), "launch"));
  rst::ctx_count_launches(c, 1);
  CLOUD_TRY(k.cuda(cudaMemcpyAsync(k.H + o_cen, k.D + o_cen, 12, cudaMemcpyDeviceToHost, k.stream), "D2H"));
  CLOUD_TRY(k.cuda(cudaStreamSynchronize(k.stream), "sync"));
  std::memcpy(centroid_out, k.H + o_cen, 12);
  return RST_OK;
}

/* void ComputeExtents(cloud, &box)  point_cloud_utils.cpp:26-32 */
extern "C" int32_t rst_cloud_extents(rst_ctx* c, const rst_cloud* cloud, float* lo_out, float* hi_out) {
  if (!c) return RST_ERR_INVALID_ARG;
  CloudCall k(c);
  if (!cloud_ok(cloud) || !lo_out || !hi_out) return k.fail(RST_ERR_INVALID_ARG, "null argument / bad cloud");
  for (int a = 0; a < 3; ++a) { lo_out[a] = FLT_MAX; hi_out[a] = -FLT_MAX; }   // the empty box
  if (cloud->n == 0) return RST_OK;
  CLOUD_TRY(k.begin());
  const size_t n = (size_t)cloud->n;
  const size_t o_pts = k.take(12 * n);
  const size_t upload = k.off;
  const size_t o_box = k.take(24);
  const size_t host_end = k.off;
  CLOUD_TRY(k.commit(host_end));
  std::memcpy(k.H + o_pts, cloud->xyz, 12 * n);
  CLOUD_TRY(k.cuda(cudaMemcpyAsync(k.D, k.H, upload, cudaMemcpyHostToDevice, k.stream), "H2D"));
  k_extents<<<1, kThreads, 0, k.stream>>>(reinterpret_cast<const float*>(k.D + o_pts), (int)n, reinterpret_cast<float*>(k.D + o_box));
  CLOUD_TRY(k.cuda(cudaGetLastError(), "launch"));
  rst::ctx_count_launches(c, 1);
  CLOUD_TRY(k.cuda(cudaMemcpyAsync(k.H + o_box, k.D + o_box, 24, cudaMemcpyDeviceToHost, k.stream), "D2H"));
  CLOUD_TRY(k.cuda(cudaStreamSynchronize(k.stream), "sync"));
  std::memcpy(lo_out, k.H + o_box, 12);
  std::memcpy(hi_out, k.H + o_box + 12, 12);
  return RST_OK;
}

/* void OrientNormals(cloud, viewpoint, &normals)  point_cloud_utils.cpp:205-216 */
extern "C" int32_t rst_orient_normals(rst_ctx* c, const rst_cloud* cloud, const float* viewpoint, float* normals_inout) {
  if (!c) return RST_ERR_INVALID_ARG;
  CloudCall k(c);
  if (!cloud_ok(cloud) || !viewpoint || !normals_inout) return k.fail(RST_ERR_INVALID_ARG, "null argument / bad cloud");
  if (cloud->n == 0) return RST_OK;
  CLOUD_TRY(k.begin());
  const size_t n = (size_t)cloud->n;
  const size_t o_pts = k.take(12 * n), o_nrm = k.take(12 * n);
  const size_t host_end = k.off;
  CLOUD_TRY(k.commit(host_end));
  std::memcpy(k.H + o_pts, cloud->xyz, 12 * n);
  std::memcpy(k.H + o_nrm, normals_inout, 12 * n);
  CLOUD_TRY(k.cuda(cudaMemcpyAsync(k.D, k.H, host_end, cudaMemcpyHostToDevice, k.stream), "H2D"));
  k_orient_normals<<<(unsigned)((n + 255) / 256), 256, 0, k.stream>>>(reinterpret_cast<const float*>(k.D + o_pts), (int)n, viewpoint[0],
                                                                      viewpoint[1], viewpoint[2], reinterpret_cast<float*>(k.D + o_nrm));
  CLOUD_TRY(k.cuda(cudaGetLastError(), "launch"));
  rst::ctx_count_launches(c, 1);
  CLOUD_TRY(k.cuda(cudaMemcpyAsync(k.H + o_nrm, k.D + o_nrm, 12 * n, cudaMemcpyDeviceToHost, k.stream), "D2H"));
  CLOUD_TRY(k.cuda(cudaStreamSynchronize(k.stream), "sync"));
  std::memcpy(normals_inout, k.H + o_nrm, 12 * n);
  return RST_OK;
}

static int32_t compact_cloud(rst_ctx* c, const rst_cloud* in, float voxel, bool by_voxel, float* xyz_out, int32_t* n_out) {
  CloudCall k(c);
  if (!cloud_ok(in) || !xyz_out || !n_out) return k.fail(RST_ERR_INVALID_ARG, "null argument / bad cloud");
  if (by_voxel && !(voxel > 0.f)) return k.fail(RST_ERR_INVALID_ARG, "voxel_size must be positive");
  *n_out = 0;
  if (in->n == 0) return RST_OK;
  CLOUD_TRY(k.begin());
  const size_t n = (size_t)in->n;
  uint32_t cap = 1;
  while (cap < 2 * n) cap <<= 1;   // load factor <= 0.5 even if every point is its own voxel
  const size_t o_pts = k.take(12 * n);
  const size_t upload = k.off;
  const size_t o_out = k.take(12 * n), o_cnt = k.take(4);
  const size_t host_end = k.off;
  const size_t o_keys = k.take(by_voxel ? (size_t)cap * 8 : 0), o_vals = k.take(by_voxel ? (size_t)cap * 4 : 0);
  CLOUD_TRY(k.commit(host_end));
  std::memcpy(k.H + o_pts, in->xyz, 12 * n);
  CLOUD_TRY(k.cuda(cudaMemcpyAsync(k.D, k.H, upload, cudaMemcpyHostToDevice, k.stream), "H2D"));
  const float* pts = reinterpret_cast<const float*>(k.D + o_pts);
  unsigned long long* keys = reinterpret_cast<unsigned long long*>(k.D + o_keys);
  int* vals = reinterpret_cast<int*>(k.D + o_vals);
  if (by_voxel) {
    CLOUD_TRY(k.cuda(cudaMemsetAsync(keys, 0, (size_t)cap * 8, k.stream), "memset"));
    CLOUD_TRY(k.cuda(cudaMemsetAsync(vals, 0x7f, (size_t)cap * 4, k.stream), "memset"));   // 0x7f7f7f7f > any point index
    k_voxel_insert<<<(unsigned)((n + 255) / 256), 256, 0, k.stream>>>(pts, (int)n, voxel, keys, vals, cap - 1);
    k_compact_points<0><<<1, kThreads, 0, k.stream>>>(pts, (int)n, voxel, keys, vals, cap - 1, reinterpret_cast<float*>(k.D + o_out),
                                                       reinterpret_cast<int*>(k.D + o_cnt));
    rst::ctx_count_launches(c, 2);
  } else {
    k_compact_points<1><<<1, kThreads, 0, k.stream>>>(pts, (int)n, 0.f, nullptr, nullptr, 0u, reinterpret_cast<float*>(k.D + o_out),
                                                       reinterpret_cast<int*>(k.D + o_cnt));
    rst::ctx_count_launches(c, 1);
  }
  CLOUD_TRY(k.cuda(cudaGetLastError(), "launch"));
  CLOUD_TRY(k.cuda(cudaMemcpyAsync(k.H + o_out, k.D + o_out, host_end - o_out, cudaMemcpyDeviceToHost, k.stream), "D2H"));
  CLOUD_TRY(k.cuda(cudaStreamSynchronize(k.stream), "sync"));
  const int cnt = *reinterpret_cast<const int*>(k.H + o_cnt);
  std::memcpy(xyz_out, k.H + o_out, 12 * (size_t)cnt);
  *n_out = cnt;
  return RST_OK;
}

/* void DownsampleVoxel(cloud_in, voxel_size, &cloud_out)  point_cloud_utils.cpp:34-68 */
extern "C" int32_t rst_downsample_voxel(rst_ctx* c, const rst_cloud* cloud_in, float voxel_size, float* xyz_out, int32_t* n_out) {
  if (!c) return RST_ERR_INVALID_ARG;
  return compact_cloud(c, cloud_in, voxel_size, true, xyz_out, n_out);
}

/* void RemoveNans(cloud_in, &cloud_out)  point_cloud_utils.cpp:163-174 */
extern "C" int32_t rst_remove_nans(rst_ctx* c, const rst_cloud* cloud_in, float* xyz_out, int32_t* n_out) {
  if (!c) return RST_ERR_INVALID_ARG;
  return compact_cloud(c, cloud_in, 0.f, false, xyz_out, n_out);
}

// ------------------------------------------------------------------------------------------------
// GICP (align_gicp.cpp:41-163): the residual / normal-equation evaluation for GIVEN correspondences and covariances
// (the 7-argument ComputeAlignment's cost, :59-77), and the 3-argument driver (:119-163): sample covariances, then
// kMaxIter rounds of { FindCorrespondences, minimise over the fixed correspondences }. The reference minimises with
// Ceres (Levenberg-Marquardt, DENSE_QR, autodiff through C^{-1/2}); Ceres is absent and out of scope, so the inner
// minimisation is Levenberg-Marquardt on the Gauss-Newton normal equations of the same robustified cost.
// ------------------------------------------------------------------------------------------------
extern "C" int32_t rst_gicp_evaluate(rst_ctx* c, const rst_cloud* src, const rst_cloud* dst, const float* src_covs, const float* dst_covs,
                                     const int32_t* dst_indices, const float* pose, float huber_delta, float* residuals_out,
                                     rst_gicp_stats* stats_out) {
  if (!c) return RST_ERR_INVALID_ARG;
  CloudCall k(c);
  if (!cloud_ok(src) || !cloud_ok(dst) || !src_covs || !dst_covs || !dst_indices || !pose || !stats_out)
    return k.fail(RST_ERR_INVALID_ARG, "null argument / bad cloud");
  std::memset(stats_out, 0, sizeof(*stats_out));
  if (src->n == 0 || dst->n == 0) return RST_OK;
  CLOUD_TRY(k.begin());
  const size_t n = (size_t)src->n, m = (size_t)dst->n;
  const int n_blocks = (int)((n + kGicpThreads - 1) / kGicpThreads);
  const size_t o_src = k.take(12 * n), o_dst = k.take(12 * m), o_sc = k.take(36 * n), o_dc = k.take(36 * m), o_idx = k.take(4 * n);
  const size_t o_state = k.take(sizeof(GicpState));
  const size_t upload = k.off;
  const size_t o_res = k.take(12 * n);
  const size_t host_end = k.off;
  const size_t o_part = k.take(sizeof(double) * kGicpSums * (size_t)n_blocks);
  CLOUD_TRY(k.commit(host_end));
  std::memcpy(k.H + o_src, src->xyz, 12 * n); std::memcpy(k.H + o_dst, dst->xyz, 12 * m);
  std::memcpy(k.H + o_sc, src_covs, 36 * n); std::memcpy(k.H + o_dc, dst_covs, 36 * m);
  std::memcpy(k.H + o_idx, dst_indices, 4 * n);
  GicpState* hs = reinterpret_cast<GicpState*>(k.H + o_state);
  std::memset(hs, 0, sizeof(*hs));
  std::memcpy(hs->pose_cm, pose, 64);
  CLOUD_TRY(k.cuda(cudaMemcpyAsync(k.D, k.H, upload, cudaMemcpyHostToDevice, k.stream), "H2D"));
  GicpState* ds = reinterpret_cast<GicpState*>(k.D + o_state);
  k_gicp_residuals<<<n_blocks, kGicpThreads, 0, k.stream>>>(
      reinterpret_cast<const float*>(k.D + o_src), reinterpret_cast<const float*>(k.D + o_dst), reinterpret_cast<const float*>(k.D + o_sc),
      reinterpret_cast<const float*>(k.D + o_dc), reinterpret_cast<const int*>(k.D + o_idx), (int)n, (int)m, ds->pose_cm, huber_delta,
      residuals_out ? reinterpret_cast<float*>(k.D + o_res) : nullptr, reinterpret_cast<double*>(k.D + o_part));
  k_gicp_finish<<<1, 32, 0, k.stream>>>(reinterpret_cast<const double*>(k.D + o_part), n_blocks, ds, 0);
  CLOUD_TRY(k.cuda(cudaGetLastError(), "launch"));
  rst::ctx_count_launches(c, 2);
  CLOUD_TRY(k.cuda(cudaMemcpyAsync(k.H + o_state, k.D + o_state, sizeof(GicpState), cudaMemcpyDeviceToHost, k.stream), "D2H"));
  if (residuals_out) CLOUD_TRY(k.cuda(cudaMemcpyAsync(k.H + o_res, k.D + o_res, 12 * n, cudaMemcpyDeviceToHost, k.stream), "D2H"));
  CLOUD_TRY(k.cuda(cudaStreamSynchronize(k.stream), "sync"));
  for (int i = 0; i < 21; ++i) stats_out->A[i] = hs->sums[i];
  for (int i = 0; i < 6; ++i) stats_out->b[i] = hs->sums[21 + i];
  stats_out->cost = hs->sums[27];
  stats_out->count = (int32_t)hs->sums[28];
  if (residuals_out) std::memcpy(residuals_out, k.H + o_res, 12 * n);
  return RST_OK;
}

/* The 7-argument ComputeAlignment (align_gicp.cpp:41-117): minimise the robustified cost over the pose for GIVEN
 * covariances and correspondences, starting from the seed. */
extern "C" int32_t rst_gicp_minimize(rst_ctx* c, const rst_cloud* src, const rst_cloud* dst, const float* src_covs, const float* dst_covs,
                                     const int32_t* dst_indices, int32_t max_iters, float huber_delta, float* pose_inout,
                                     rst_gicp_stats* stats_out) {
  if (!c) return RST_ERR_INVALID_ARG;
  CloudCall k(c);
  if (!cloud_ok(src) || !cloud_ok(dst) || !src_covs || !dst_covs || !dst_indices || !pose_inout || max_iters < 0)
    return k.fail(RST_ERR_INVALID_ARG, "null argument / bad cloud / negative iteration count");
  if (stats_out) std::memset(stats_out, 0, sizeof(*stats_out));
  if (src->n == 0 || dst->n == 0) return RST_OK;   // no residual block: the seed is the minimiser
  CLOUD_TRY(k.begin());
  const size_t n = (size_t)src->n, m = (size_t)dst->n;
  const int n_blocks = (int)((n + kGicpThreads - 1) / kGicpThreads);
  const size_t o_src = k.take(12 * n), o_dst = k.take(12 * m), o_sc = k.take(36 * n), o_dc = k.take(36 * m), o_idx = k.take(4 * n);
  const size_t o_state = k.take(sizeof(GicpState));
  const size_t upload = k.off;
  const size_t host_end = k.off;
  const size_t o_part = k.take(sizeof(double) * kGicpSums * (size_t)n_blocks);
  CLOUD_TRY(k.commit(host_end));
  std::memcpy(k.H + o_src, src->xyz, 12 * n); std::memcpy(k.H + o_dst, dst->xyz, 12 * m);
  std::memcpy(k.H + o_sc, src_covs, 36 * n); std::memcpy(k.H + o_dc, dst_covs, 36 * m);
  std::memcpy(k.H + o_idx, dst_indices, 4 * n);
  GicpState* hs = reinterpret_cast<GicpState*>(k.H + o_state);
  std::memset(hs, 0, sizeof(*hs));
  std::memcpy(hs->pose_cm, pose_inout, 64);
  for (int r = 0; r < 3; ++r) { for (int cc = 0; cc < 3; ++cc) hs->Rt[3 * r + cc] = pose_inout[r + 4 * cc]; hs->Rt[9 + r] = pose_inout[12 + r]; }
  hs->lambda = 1e-4;
  CLOUD_TRY(k.cuda(cudaMemcpyAsync(k.D, k.H, upload, cudaMemcpyHostToDevice, k.stream), "H2D"));
  GicpState* ds = reinterpret_cast<GicpState*>(k.D + o_state);
  const float* d_src = reinterpret_cast<const float*>(k.D + o_src);
  const float* d_dst = reinterpret_cast<const float*>(k.D + o_dst);
  const float* d_sc = reinterpret_cast<const float*>(k.D + o_sc);
  const float* d_dc = reinterpret_cast<const float*>(k.D + o_dc);
  const int* d_idx = reinterpret_cast<const int*>(k.D + o_idx);
  double* d_part = reinterpret_cast<double*>(k.D + o_part);
  // max_iters Levenberg-Marquardt steps (the first evaluation is accepted unconditionally), then the statistics of the
  // best pose: mode 3 falls back to the last accepted pose when the final step made the cost worse
  for (int it = 0; it <= max_iters; ++it) {
    k_gicp_residuals<<<n_blocks, kGicpThreads, 0, k.stream>>>(d_src, d_dst, d_sc, d_dc, d_idx, (int)n, (int)m, ds->pose_cm, huber_delta, nullptr, d_part);
    k_gicp_finish<<<1, 32, 0, k.stream>>>(d_part, n_blocks, ds, it == max_iters ? 3 : (it == 0 ? 2 : 1));
  }
  CLOUD_TRY(k.cuda(cudaGetLastError(), "launch"));
  rst::ctx_count_launches(c, 2 * (max_iters + 1));
  CLOUD_TRY(k.cuda(cudaMemcpyAsync(k.H + o_state, k.D + o_state, sizeof(GicpState), cudaMemcpyDeviceToHost, k.stream), "D2H"));
  CLOUD_TRY(k.cuda(cudaStreamSynchronize(k.stream), "sync"));
  std::memcpy(pose_inout, hs->pose_cm, 64);
  pose_inout[3] = pose_inout[7] = pose_inout[11] = 0.f; pose_inout[15] = 1.f;
  if (stats_out) {
    for (int i = 0; i < 21; ++i) stats_out->A[i] = hs->sums[i];
    for (int i = 0; i < 6; ++i) stats_out->b[i] = hs->sums[21 + i];
    stats_out->cost = hs->sums[27];
    stats_out->count = (int32_t)hs->sums[28];
  }
  return RST_OK;
}

extern "C" int32_t rst_gicp_align(rst_ctx* c, const rst_cloud* src, const rst_cloud* dst, int32_t max_outer, int32_t inner_iters,
                                  float huber_delta, int32_t use_gicp_covariances, float grid_cell, float* pose_inout, rst_gicp_stats* stats_out) {
  if (!c) return RST_ERR_INVALID_ARG;
  CloudCall k(c);
  if (!cloud_ok(src) || !cloud_ok(dst) || !pose_inout || max_outer < 0 || inner_iters < 1)
    return k.fail(RST_ERR_INVALID_ARG, "null argument / bad cloud / bad iteration counts");
  if (stats_out) std::memset(stats_out, 0, sizeof(*stats_out));
  if (src->n < 3 || dst->n < 3) return k.fail(RST_ERR_INVALID_ARG, "clouds need at least 3 points");
  CLOUD_TRY(k.begin());
  const size_t n = (size_t)src->n, m = (size_t)dst->n;
  const int n_blocks = (int)((n + kGicpThreads - 1) / kGicpThreads);
  const size_t o_src = k.take(12 * n), o_dst = k.take(12 * m), o_state = k.take(sizeof(GicpState));
  const size_t upload = k.off;
  const size_t host_end = k.off;
  const size_t o_sc = k.take(36 * n), o_dc = k.take(36 * m), o_idx = k.take(4 * n), o_d2 = k.take(4 * n), o_tmp = k.take(12 * n);
  const size_t o_part = k.take(sizeof(double) * kGicpSums * (size_t)n_blocks);
  const size_t o_grid = k.take(sizeof(Grid)), o_cs = k.take(sizeof(int) * (kCellCap + 1)), o_cf = k.take(sizeof(int) * kCellCap);
  const size_t o_sorted = k.take(sizeof(float4) * (n > m ? n : m));
  CLOUD_TRY(k.commit(host_end));
  std::memcpy(k.H + o_src, src->xyz, 12 * n); std::memcpy(k.H + o_dst, dst->xyz, 12 * m);
  GicpState* hs = reinterpret_cast<GicpState*>(k.H + o_state);
  std::memset(hs, 0, sizeof(*hs));
  std::memcpy(hs->pose_cm, pose_inout, 64);
  for (int r = 0; r < 3; ++r) { for (int cc = 0; cc < 3; ++cc) hs->Rt[3 * r + cc] = pose_inout[r + 4 * cc]; hs->Rt[9 + r] = pose_inout[12 + r]; }
  hs->lambda = 1e-4;
  CLOUD_TRY(k.cuda(cudaMemcpyAsync(k.D, k.H, upload, cudaMemcpyHostToDevice, k.stream), "H2D"));
  const float* d_src = reinterpret_cast<const float*>(k.D + o_src);
  const float* d_dst = reinterpret_cast<const float*>(k.D + o_dst);
  Grid* d_grid = reinterpret_cast<Grid*>(k.D + o_grid);
  int* d_cs = reinterpret_cast<int*>(k.D + o_cs); int* d_cf = reinterpret_cast<int*>(k.D + o_cf);
  float4* d_sorted = reinterpret_cast<float4*>(k.D + o_sorted);
  GicpState* ds = reinterpret_cast<GicpState*>(k.D + o_state);
  int launches = 0;
  // covariances of both clouds (align_gicp.cpp:136-140; the reference passes use_gicp = false)
  k_grid_build<<<1, kThreads, 0, k.stream>>>(d_src, (int)n, grid_cell, d_cs, d_cf, d_sorted, d_grid);
  k_covariances<<<(unsigned)((n + 127) / 128), 128, 0, k.stream>>>(d_grid, d_cs, d_sorted, d_src, (int)n, use_gicp_covariances ? 1 : 0,
                                                                     reinterpret_cast<float*>(k.D + o_sc));
  k_grid_build<<<1, kThreads, 0, k.stream>>>(d_dst, (int)m, grid_cell, d_cs, d_cf, d_sorted, d_grid);   // stays: the NN target
  k_covariances<<<(unsigned)((m + 127) / 128), 128, 0, k.stream>>>(d_grid, d_cs, d_sorted, d_dst, (int)m, use_gicp_covariances ? 1 : 0,
                                                                     reinterpret_cast<float*>(k.D + o_dc));
  launches += 4;
  for (int outer = 0; outer < max_outer; ++outer) {
    // tmp = estimate * src; FindCorrespondences(dst_tree, tmp) (:149-150,160)
    k_transform_points<<<(unsigned)((n + 255) / 256), 256, 0, k.stream>>>(d_src, (int)n, ds->pose_cm, reinterpret_cast<float*>(k.D + o_tmp));
    k_nn_query<<<(unsigned)((n + 127) / 128), 128, 0, k.stream>>>(d_grid, d_cs, d_sorted, reinterpret_cast<const float*>(k.D + o_tmp), (int)n,
                                                                    reinterpret_cast<int*>(k.D + o_idx), reinterpret_cast<float*>(k.D + o_d2));
    launches += 2;
    for (int it = 0; it < inner_iters; ++it) {
      k_gicp_residuals<<<n_blocks, kGicpThreads, 0, k.stream>>>(d_src, d_dst, reinterpret_cast<const float*>(k.D + o_sc),
                                                                reinterpret_cast<const float*>(k.D + o_dc), reinterpret_cast<const int*>(k.D + o_idx),
                                                                (int)n, (int)m, ds->pose_cm, huber_delta, nullptr, reinterpret_cast<double*>(k.D + o_part));
      k_gicp_finish<<<1, 32, 0, k.stream>>>(reinterpret_cast<const double*>(k.D + o_part), n_blocks, ds, it == 0 ? 2 : 1);
      launches += 2;
    }
  }
  // statistics of the final pose over the last correspondences
  if (max_outer > 0) {
    k_gicp_residuals<<<n_blocks, kGicpThreads, 0, k.stream>>>(d_src, d_dst, reinterpret_cast<const float*>(k.D + o_sc),
                                                              reinterpret_cast<const float*>(k.D + o_dc), reinterpret_cast<const int*>(k.D + o_idx),
                                                              (int)n, (int)m, ds->pose_cm, huber_delta, nullptr, reinterpret_cast<double*>(k.D + o_part));
    k_gicp_finish<<<1, 32, 0, k.stream>>>(reinterpret_cast<const double*>(k.D + o_part), n_blocks, ds, 3);
    launches += 2;
  }
  CLOUD_TRY(k.cuda(cudaGetLastError(), "launch"));
  rst::ctx_count_launches(c, launches);
  CLOUD_TRY(k.cuda(cudaMemcpyAsync(k.H + o_state, k.D + o_state, sizeof(GicpState), cudaMemcpyDeviceToHost, k.stream), "D2H"));
  CLOUD_TRY(k.cuda(cudaStreamSynchronize(k.stream), "sync"));
  std::memcpy(pose_inout, hs->pose_cm, 64);
  if (stats_out) {
    for (int i = 0; i < 21; ++i) stats_out->A[i] = hs->sums[i];
    for (int i = 0; i < 6; ++i) stats_out->b[i] = hs->sums[21 + i];
    stats_out->cost = hs->sums[27];
    stats_out->count = (int32_t)hs->sums[28];
  }
  return RST_OK;
}
