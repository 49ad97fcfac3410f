/*
 * oracle_r.cpp — Oracle-R: dependency-free CPU restatement of the reference's
 * own alignment path (KD-tree nearest-neighbour, Geman-McClure weighted,
 * closed-form Kabsch ICP) and of the steps its caller runs before it.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under realsensetracker_b200/ may include,
 * link or call this file; it is the checker for the cloud-based GPU path and the
 * CPU baseline that bench.py times (`cpu_baseline`, `--impl reference`).
 *
 * PINNING: the reference has no tests / golden vectors / fixtures, and its real
 * dependencies (Eigen3, nanoflann, ChoUtil, fmt — versions unpinned, `find_package(...
 * REQUIRED)` CMakeLists.txt:11-21) are absent, so it cannot be built as shipped. What IS
 * done: its own align_icp.cpp and point_cloud_utils.cpp are compiled UNMODIFIED from
 * /root/reference against stand-in headers (oracle/shim, oracle/Makefile target `ref` ->
 * oracle/_ref/libref.so) and this restatement matches that build BIT FOR BIT on AlignIcp3d
 * (1..128 iterations), SolveKabsch, ComputeCentroid, RemoveNans, FindCorrespondences and, as
 * a set, DownsampleVoxel (tests/test_reference_compiled.py). That pins the reference's own
 * control flow and quirks; the THIRD-PARTY arithmetic stays unpinned and is restated from
 * the published algorithms on both sides:
 *   - nanoflann KDTreeSingleIndexAdaptor<L2> exact 1-NN (kdtree.hpp:51-57):
 *     result = exact nearest neighbour under fp32 squared L2, accumulated
 *     dx^2+dy^2+dz^2 left to right; any exact search agrees off exact ties.
 *   - Eigen::JacobiSVD<Matrix3d> (align_icp.cpp:139-140): U*V^T is unique for a
 *     non-degenerate covariance; computed here by one-sided Jacobi in fp64.
 *   - Eigen::Quaternionf(Matrix3f) and Quaternionf::toRotationMatrix()
 *     (align_icp.cpp:151): Shoemake's branches, restated below.
 * It is cross-checked in tests/ against scipy.spatial.cKDTree, numpy.linalg.svd
 * and known-motion synthetic scenes.
 *
 * Followed line by line (paths relative to the reference tree):
 *   rs_tracker/align/src/align_icp.cpp:18-71    SolveKabsch
 *   rs_tracker/align/src/align_icp.cpp:73-161   AlignIcp3d (5-arg)
 *   rs_tracker/align/src/align_icp.cpp:163-167  AlignIcp3d (4-arg, leaf 16)
 *   rs_tracker/common/src/point_cloud_utils.cpp:34-68    DownsampleVoxel
 *   rs_tracker/common/src/point_cloud_utils.cpp:92-98    ComputeCentroid
 *   rs_tracker/common/src/point_cloud_utils.cpp:163-174  RemoveNans
 *   rs_tracker/app/src/rs_replay_app.cpp:229,246-251     caller sequence
 * Deviation (documented): DownsampleVoxel's output order in the reference is
 * std::unordered_map iteration order (implementation-defined); here it is
 * first-occurrence order, which is deterministic.
 */
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <numeric>
#include <unordered_map>
#include <vector>

namespace {

struct V3 { float x, y, z; };

/* ---------------- exact kd-tree (stands in for nanoflann) ---------------- */
struct KdNode {
  int32_t left, right;  /* children, or -1 */
  int32_t begin, end;   /* leaf point range in `order` */
  int32_t dim;
  float split;
};

struct KdTree {
  const float* pts;  /* 3*N interleaved, borrowed (kdtree.hpp:43) */
  int32_t n;
  int32_t leaf;
  std::vector<int32_t> order;
  std::vector<KdNode> nodes;

  KdTree(const float* p, int32_t n_, int32_t leaf_) : pts(p), n(n_), leaf(leaf_), order(n_) {
    std::iota(order.begin(), order.end(), 0);
    nodes.reserve(2 * (n / std::max(1, leaf)) + 8);
    if (n > 0) build(0, n);
  }

  int32_t build(int32_t b, int32_t e) {
    const int32_t id = (int32_t)nodes.size();
    nodes.push_back(KdNode{-1, -1, b, e, 0, 0.f});
    if (e - b <= leaf) return id;
    float lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
    for (int32_t i = b; i < e; ++i)
      for (int k = 0; k < 3; ++k) {
        const float v = pts[3 * order[i] + k];
        lo[k] = std::min(lo[k], v); hi[k] = std::max(hi[k], v);
      }
    int dim = 0;
    for (int k = 1; k < 3; ++k) if (hi[k] - lo[k] > hi[dim] - lo[dim]) dim = k;
    if (!(hi[dim] > lo[dim])) return id; /* all points identical: keep as leaf */
    const int32_t m = b + (e - b) / 2;
    std::nth_element(order.begin() + b, order.begin() + m, order.begin() + e,
                     [&](int32_t a, int32_t c) { return pts[3 * a + dim] < pts[3 * c + dim]; });
    nodes[id].dim = dim;
    nodes[id].split = pts[3 * order[m] + dim];
    const int32_t l = build(b, m);
    const int32_t r = build(m, e);
    nodes[id].left = l; nodes[id].right = r;
    return id;
  }

  void search(int32_t id, const float* q, int32_t& best, float& best_d) const {
    const KdNode& nd = nodes[id];
    if (nd.left < 0) {
      for (int32_t i = nd.begin; i < nd.end; ++i) {
        const int32_t j = order[i];
        const float dx = q[0] - pts[3 * j], dy = q[1] - pts[3 * j + 1], dz = q[2] - pts[3 * j + 2];
        const float d = dx * dx + dy * dy + dz * dz; /* left-to-right fp32, as nanoflann L2 */
        if (d < best_d || (d == best_d && j < best)) { best_d = d; best = j; }
      }
      return;
    }
    const float diff = q[nd.dim] - nd.split;
    const int32_t near = diff < 0 ? nd.left : nd.right, far = diff < 0 ? nd.right : nd.left;
    search(near, q, best, best_d);
    if (diff * diff <= best_d) search(far, q, best, best_d);
  }

  /* KDTreeChoCloudAdaptor::query(p, 1, &j, &d2)  kdtree.hpp:51-57 */
  void query1(const float* q, int32_t* j, float* d2) const {
    int32_t best = -1; float bd = INFINITY;
    search(0, q, best, bd);
    *j = best; *d2 = bd;
  }
};

/* ---------------- 3x3 fp64 SVD -> U*V^T (stands in for Eigen::JacobiSVD) ---------------- */
void svd_uvt(const double* M /* row-major */, double* UVt) {
  /* one-sided Jacobi on B = M*V: rotate column pairs until orthogonal */
  double B[9], V[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
  std::memcpy(B, M, sizeof(B));
  for (int sweep = 0; sweep < 60; ++sweep) {
    double off = 0;
    for (int p = 0; p < 2; ++p)
      for (int q = p + 1; q < 3; ++q) {
        double a = 0, b = 0, c = 0;
        for (int i = 0; i < 3; ++i) { a += B[3 * i + p] * B[3 * i + p]; b += B[3 * i + q] * B[3 * i + q]; c += B[3 * i + p] * B[3 * i + q]; }
        off = std::max(off, std::fabs(c) / std::sqrt(a * b + 1e-300));
        if (std::fabs(c) <= 1e-300) continue;
        const double zeta = (b - a) / (2.0 * c);
        const double t = (zeta >= 0 ? 1.0 : -1.0) / (std::fabs(zeta) + std::sqrt(1.0 + zeta * zeta));
        const double cs = 1.0 / std::sqrt(1.0 + t * t), sn = cs * t;
        for (int i = 0; i < 3; ++i) {
          const double bp = B[3 * i + p], bq = B[3 * i + q];
          B[3 * i + p] = cs * bp - sn * bq; B[3 * i + q] = sn * bp + cs * bq;
          const double vp = V[3 * i + p], vq = V[3 * i + q];
          V[3 * i + p] = cs * vp - sn * vq; V[3 * i + q] = sn * vp + cs * vq;
        }
      }
    if (off < 1e-15) break;
  }
  /* U columns = normalised B columns; a (near-)null column is completed by the cross product */
  double U[9], s[3];
  for (int j = 0; j < 3; ++j) {
    s[j] = std::sqrt(B[j] * B[j] + B[3 + j] * B[3 + j] + B[6 + j] * B[6 + j]);
  }
  const double smax = std::max(s[0], std::max(s[1], s[2]));
  int bad = -1;
  for (int j = 0; j < 3; ++j) {
    if (s[j] > 1e-14 * smax && s[j] > 0) { for (int i = 0; i < 3; ++i) U[3 * i + j] = B[3 * i + j] / s[j]; }
    else bad = j;
  }
  if (bad >= 0) {
    const int a = (bad + 1) % 3, b = (bad + 2) % 3;
    const double ux = U[3 * 1 + a] * U[3 * 2 + b] - U[3 * 2 + a] * U[3 * 1 + b];
    const double uy = U[3 * 2 + a] * U[3 * 0 + b] - U[3 * 0 + a] * U[3 * 2 + b];
    const double uz = U[3 * 0 + a] * U[3 * 1 + b] - U[3 * 1 + a] * U[3 * 0 + b];
    U[bad] = ux; U[3 + bad] = uy; U[6 + bad] = uz;
  }
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) {
      double acc = 0;
      for (int k = 0; k < 3; ++k) acc += U[3 * i + k] * V[3 * j + k];
      UVt[3 * i + j] = acc;
    }
}

/* R (fp32, row-major) with the literal reflection patch of align_icp.cpp:141-145 */
void rotation_from_cov(const double* cov, float* R) {
  double uvt[9];
  svd_uvt(cov, uvt);
  for (int i = 0; i < 9; ++i) R[i] = (float)uvt[i];
  const float det = R[0] * (R[4] * R[8] - R[5] * R[7]) - R[1] * (R[3] * R[8] - R[5] * R[6]) +
                    R[2] * (R[3] * R[7] - R[4] * R[6]);
  if (det < 0) { R[2] *= -1; R[5] *= -1; R[8] *= -1; } /* R.col(2) *= -1, as written */
}

/* xfm = Translation3f{t} * Quaternionf{R}  (align_icp.cpp:69,151): the rotation is
 * re-expressed through a quaternion. Output: column-major 4x4. */
void compose(const float* R /* row-major */, const float* t, float* T) {
  float q[4]; /* x y z w */
  float tr = R[0] + R[4] + R[8];
  if (tr > 0.f) {
    tr = std::sqrt(tr + 1.0f);
    q[3] = 0.5f * tr;
    tr = 0.5f / tr;
    q[0] = (R[3 * 2 + 1] - R[3 * 1 + 2]) * tr;
    q[1] = (R[3 * 0 + 2] - R[3 * 2 + 0]) * tr;
    q[2] = (R[3 * 1 + 0] - R[3 * 0 + 1]) * tr;
  } else {
    int i = 0;
    if (R[4] > R[0]) i = 1;
    if (R[8] > R[4 * i]) i = 2;
    const int j = (i + 1) % 3, k = (j + 1) % 3;
    tr = std::sqrt(R[4 * i] - R[4 * j] - R[4 * k] + 1.0f);
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
  const float Rq[9] = {1.f - (tyy + tzz), txy - twz, txz + twy,
                       txy + twz, 1.f - (txx + tzz), tyz - twx,
                       txz - twy, tyz + twx, 1.f - (txx + tyy)};
  for (int r = 0; r < 3; ++r) {
    for (int c = 0; c < 3; ++c) T[r + 4 * c] = Rq[3 * r + c];
    T[12 + r] = t[r];
    T[4 * r + 3] = 0.f;
  }
  T[15] = 1.f;
}

inline V3 apply(const float* T, const float* p) { /* Isometry3f * Vector3f: R*p + t */
  V3 o;
  o.x = T[0] * p[0] + T[4] * p[1] + T[8] * p[2] + T[12];
  o.y = T[1] * p[0] + T[5] * p[1] + T[9] * p[2] + T[13];
  o.z = T[2] * p[0] + T[6] * p[1] + T[10] * p[2] + T[14];
  return o;
}

}  // namespace

extern "C" {

/* ComputeCentroid  point_cloud_utils.cpp:92-98: sequential fp32 sum, then *= float(1.0/N). */
void or_centroid(const float* pts, int32_t n, float* c) {
  float s[3] = {0, 0, 0};
  for (int32_t i = 0; i < n; ++i) { s[0] += pts[3 * i]; s[1] += pts[3 * i + 1]; s[2] += pts[3 * i + 2]; }
  const float inv = (float)(1.0 / n);
  c[0] = s[0] * inv; c[1] = s[1] * inv; c[2] = s[2] * inv;
}

/* RemoveNans  point_cloud_utils.cpp:163-174. Returns the output count. */
int32_t or_remove_nans(const float* in, int32_t n, float* out) {
  int32_t m = 0;
  for (int32_t i = 0; i < n; ++i) {
    if (!std::isfinite(in[3 * i]) || !std::isfinite(in[3 * i + 1]) || !std::isfinite(in[3 * i + 2])) continue;
    out[3 * m] = in[3 * i]; out[3 * m + 1] = in[3 * i + 1]; out[3 * m + 2] = in[3 * i + 2];
    ++m;
  }
  return m;
}

/* DownsampleVoxel  point_cloud_utils.cpp:34-68: key = floor(p / voxel), first point
 * wins. Output in first-occurrence order (see header note). Returns the count. */
int32_t or_downsample_voxel(const float* in, int32_t n, float voxel, float* out) {
  struct Key { int32_t x, y, z; bool operator==(const Key& o) const { return x == o.x && y == o.y && z == o.z; } };
  struct Hash { size_t operator()(const Key& k) const {
    uint64_t h = (uint64_t)(uint32_t)k.x * 0x9E3779B97F4A7C15ull;
    h ^= ((uint64_t)(uint32_t)k.y + 0x7F4A7C15ull + (h << 6) + (h >> 2));
    h ^= ((uint64_t)(uint32_t)k.z + 0x94D049BBull + (h << 6) + (h >> 2));
    return (size_t)h; } };
  std::unordered_map<Key, int32_t, Hash> vox;
  vox.reserve((size_t)n / 4 + 16);
  int32_t m = 0;
  for (int32_t i = 0; i < n; ++i) {
    const Key k{(int32_t)std::floor(in[3 * i] / voxel), (int32_t)std::floor(in[3 * i + 1] / voxel),
                (int32_t)std::floor(in[3 * i + 2] / voxel)};
    if (vox.emplace(k, i).second) {
      out[3 * m] = in[3 * i]; out[3 * m + 1] = in[3 * i + 1]; out[3 * m + 2] = in[3 * i + 2];
      ++m;
    }
  }
  return m;
}

/* exact 1-NN of every query in `dst`  (kdtree.hpp:51-57; point_cloud_utils.cpp:70-90) */
void or_nn(const float* dst, int32_t m, const float* queries, int32_t n, int32_t leaf,
           int32_t* idx, float* d2) {
  const KdTree tree(dst, m, leaf);
  for (int32_t i = 0; i < n; ++i) tree.query1(queries + 3 * i, idx + i, d2 + i);
}

/* SolveKabsch  align_icp.cpp:18-71. pairs = (src index, dst index) x n_pairs;
 * weights may be NULL (= empty). T out: column-major 4x4. Returns 1 on success. */
int32_t or_solve_kabsch(const float* src, int32_t ns, const float* dst, int32_t nd,
                        const int32_t* pairs, int32_t n_pairs, const float* weights, float* T) {
  if (ns < 3 || nd < 3) return 0;
  float sm[3] = {0, 0, 0}, dm[3] = {0, 0, 0};
  for (int32_t c = 0; c < n_pairs; ++c)
    for (int k = 0; k < 3; ++k) { sm[k] += src[3 * pairs[2 * c] + k]; dm[k] += dst[3 * pairs[2 * c + 1] + k]; }
  for (int k = 0; k < 3; ++k) { sm[k] /= (float)n_pairs; dm[k] /= (float)n_pairs; }
  double cov[9] = {0};
  for (int32_t c = 0; c < n_pairs; ++c) {
    const float* ps = src + 3 * pairs[2 * c];
    const float* pd = dst + 3 * pairs[2 * c + 1];
    for (int a = 0; a < 3; ++a)
      for (int b = 0; b < 3; ++b) {
        const float prod = (pd[a] - dm[a]) * (ps[b] - sm[b]); /* fp32 outer product */
        /* weighted branch multiplies AFTER the cast (align_icp.cpp:51-53) */
        cov[3 * a + b] += weights ? (double)weights[c] * (double)prod : (double)prod;
      }
  }
  float R[9], t[3];
  rotation_from_cov(cov, R);
  for (int a = 0; a < 3; ++a) t[a] = dm[a] - (R[3 * a] * sm[0] + R[3 * a + 1] * sm[1] + R[3 * a + 2] * sm[2]);
  compose(R, t, T);
  return 1;
}

/* AlignIcp3d  align_icp.cpp:73-167. T: column-major 4x4, initial guess in, result
 * out. mean_cost_out (nullable) receives sqrt(cost/N) of the last iteration's
 * pre-update correspondences (:157). nbrs_out/weights_out (nullable, size n) are the
 * last iteration's correspondences. Returns 1 on success (mean_cost < 10000, :160). */
int32_t or_align_icp3d(const float* src, int32_t n, const float* dst, int32_t m, int32_t max_iter,
                       float* T, float* mean_cost_out, int32_t* nbrs_out, float* weights_out,
                       double* cov_out) {
  if (n < 3 || m < 3) return 0;                  /* :77-79 */
  const KdTree tree(dst, m, 16);                 /* :165 leaf 16 */
  float xfm[16];
  std::memcpy(xfm, T, sizeof(xfm));              /* :82 */
  float src_mean[3];
  or_centroid(src, n, src_mean);                 /* :85-86 */
  float cost = 0.f, mu = 1.0f;                   /* :90-91 */
  std::vector<int32_t> nbrs(n);
  std::vector<float> weights(n);
  double cov[9] = {0};
  for (int iter = 0; iter < max_iter; ++iter) {
    if (iter > 0 && iter % 8 == 0) mu /= 1.4f;   /* :96-98 */
    float dst_mean[3] = {0, 0, 0};
    cost = 0.f;
    for (int32_t i = 0; i < n; ++i) {            /* :105-121 */
      const V3 p = apply(xfm, src + 3 * i);
      int32_t j; float d2;
      tree.query1(&p.x, &j, &d2);
      cost += d2;
      nbrs[i] = j;
      const float rt = mu / (d2 + mu);
      weights[i] = rt * rt;
      dst_mean[0] += dst[3 * j]; dst_mean[1] += dst[3 * j + 1]; dst_mean[2] += dst[3 * j + 2];
    }
    for (int k = 0; k < 3; ++k) dst_mean[k] /= (float)n; /* :122 (unweighted) */
    std::memset(cov, 0, sizeof(cov));
    for (int32_t i = 0; i < n; ++i) {            /* :125-136 */
      const float* pd = dst + 3 * nbrs[i];
      const float* ps = src + 3 * i;
      for (int a = 0; a < 3; ++a) {
        const float wd = weights[i] * (pd[a] - dst_mean[a]);
        for (int b = 0; b < 3; ++b) cov[3 * a + b] += (double)(wd * (ps[b] - src_mean[b]));
      }
    }
    float R[9], t[3];
    rotation_from_cov(cov, R);                   /* :139-145 */
    for (int a = 0; a < 3; ++a)                  /* :148 */
      t[a] = dst_mean[a] - (R[3 * a] * src_mean[0] + R[3 * a + 1] * src_mean[1] + R[3 * a + 2] * src_mean[2]);
    compose(R, t, xfm);                          /* :151 */
  }
  std::memcpy(T, xfm, sizeof(xfm));              /* :156 */
  const float mean_cost = std::sqrt(cost / (float)n); /* :157 */
  if (mean_cost_out) *mean_cost_out = mean_cost;
  if (nbrs_out) std::memcpy(nbrs_out, nbrs.data(), sizeof(int32_t) * n);
  if (weights_out) std::memcpy(weights_out, weights.data(), sizeof(float) * n);
  if (cov_out) std::memcpy(cov_out, cov, sizeof(cov));
  return mean_cost < 10000;                      /* :160 */
}

/* Depth -> full-resolution cloud the way the driver delivers it: librealsense
 * back-projection (rs_driver.cpp:201-202; third-party, restated as the pin-hole
 * model) with invalid pixels mapped to the ORIGIN, not dropped (rs_driver.cpp:83-88). */
void or_backproject(const uint16_t* depth, int32_t w, int32_t h, float fx, float fy, float cx,
                    float cy, float depth_scale, float* cloud) {
  for (int32_t v = 0; v < h; ++v)
    for (int32_t u = 0; u < w; ++u) {
      const uint16_t d = depth[(size_t)v * w + u];
      float* p = cloud + 3 * ((size_t)v * w + u);
      if (d == 0) { p[0] = p[1] = p[2] = 0.f; continue; }
      const float z = (float)d * depth_scale;
      p[0] = ((float)u - cx) * z / fx; p[1] = ((float)v - cy) * z / fy; p[2] = z;
    }
}

/* The caller's per-pair sequence (rs_replay_app.cpp:229,246-251) on depth frames:
 * back-project both, RemoveNans, DownsampleVoxel(voxel) both, AlignIcp3d(curr, prev).
 * voxel <= 0 skips the decimation. n_src_out/n_dst_out: decimated cloud sizes. */
int32_t or_align_depth_pair(const uint16_t* src_depth, const uint16_t* dst_depth, int32_t w,
                            int32_t h, float fx, float fy, float cx, float cy, float depth_scale,
                            float voxel, int32_t max_iter, float* T, float* mean_cost_out,
                            int32_t* n_src_out, int32_t* n_dst_out) {
  const size_t n = (size_t)w * h;
  std::vector<float> a(3 * n), b(3 * n), c(3 * n);
  auto prep = [&](const uint16_t* d, std::vector<float>& out) -> int32_t {
    or_backproject(d, w, h, fx, fy, cx, cy, depth_scale, a.data());
    int32_t m = or_remove_nans(a.data(), (int32_t)n, b.data());
    if (voxel > 0) { m = or_downsample_voxel(b.data(), m, voxel, out.data()); }
    else { std::memcpy(out.data(), b.data(), sizeof(float) * 3 * m); }
    return m;
  };
  std::vector<float> s(3 * n);
  const int32_t ns = prep(src_depth, s);
  const int32_t nd = prep(dst_depth, c);
  if (n_src_out) *n_src_out = ns;
  if (n_dst_out) *n_dst_out = nd;
  return or_align_icp3d(s.data(), ns, c.data(), nd, max_iter, T, mean_cost_out, nullptr, nullptr, nullptr);
}

/* One pair per thread over a batch (the reference itself is single-threaded; this is
 * the "all host cores" leg of the CPU baseline). Frames are dense and back to back. */
void or_align_depth_pairs(const uint16_t* src_depth, const uint16_t* dst_depth, int32_t n_pairs,
                          int32_t w, int32_t h, float fx, float fy, float cx, float cy,
                          float depth_scale, float voxel, int32_t max_iter, int32_t n_threads,
                          float* T /* n_pairs x 16 */, int32_t* ok) {
  const size_t fs = (size_t)w * h;
#pragma omp parallel for schedule(dynamic, 1) num_threads(n_threads)
  for (int32_t i = 0; i < n_pairs; ++i)
    ok[i] = or_align_depth_pair(src_depth + i * fs, dst_depth + i * fs, w, h, fx, fy, cx, cy,
                                depth_scale, voxel, max_iter, T + 16 * i, nullptr, nullptr, nullptr);
}

}  // extern "C"
