// ref_glue.cpp — C entry points over the reference's OWN translation units
// (rs_tracker/align/src/align_icp.cpp, rs_tracker/common/src/point_cloud_utils.cpp), compiled unmodified
// from /root/reference against the stand-in headers in oracle/shim/. TEST INFRASTRUCTURE: used only to
// check Oracle-R (oracle_r.cpp) against the literal reference control flow. See oracle/Makefile (`ref`).
#include <cstring>
#include <vector>

#include "rs_tracker/align/align_icp.hpp"
#include "rs_tracker/common/point_cloud_utils.hpp"

using rs_tracker::Cloud3f;

static void to_cloud(const float* xyz, int n, Cloud3f* c) { c->SetNumPoints(n); if (n) std::memcpy(c->GetPtr(), xyz, sizeof(float) * 3 * n); }
static void to_pose(const float* T, Eigen::Isometry3f* x) { std::memcpy(x->matrix().data(), T, sizeof(float) * 16); }
static void from_pose(const Eigen::Isometry3f& x, float* T) { std::memcpy(T, x.matrix().data(), sizeof(float) * 16); }

namespace rs_tracker { void ComputeExtents(const Cloud3f& cloud, Eigen::AlignedBox3f* const box); }

extern "C" {
int ref_align_icp3d(const float* src, int n, const float* dst, int m, int max_iter, float* T) {
  Cloud3f s, d; to_cloud(src, n, &s); to_cloud(dst, m, &d);
  Eigen::Isometry3f x; to_pose(T, &x);
  const bool ok = rs_tracker::AlignIcp3d(s, d, max_iter, &x);
  from_pose(x, T);
  return ok ? 1 : 0;
}
int ref_solve_kabsch(const float* src, int n, const float* dst, int m, const int* pairs, int n_pairs, const float* weights, float* T) {
  Cloud3f s, d; to_cloud(src, n, &s); to_cloud(dst, m, &d);
  std::vector<std::pair<int, int>> idx(n_pairs);
  for (int i = 0; i < n_pairs; ++i) idx[i] = {pairs[2 * i], pairs[2 * i + 1]};
  std::vector<float> w;
  if (weights) w.assign(weights, weights + n_pairs);
  Eigen::Isometry3f x;
  const bool ok = rs_tracker::SolveKabsch(s, d, idx, w, &x);
  from_pose(x, T);
  return ok ? 1 : 0;
}
void ref_centroid(const float* pts, int n, float* c) {
  Cloud3f s; to_cloud(pts, n, &s);
  Eigen::Vector3f m; rs_tracker::ComputeCentroid(s, &m);
  c[0] = m[0]; c[1] = m[1]; c[2] = m[2];
}
int ref_remove_nans(const float* in, int n, float* out) {
  Cloud3f s, o; to_cloud(in, n, &s);
  rs_tracker::RemoveNans(s, &o);
  std::memcpy(out, o.GetPtr(), sizeof(float) * 3 * o.GetNumPoints());
  return o.GetNumPoints();
}
int ref_downsample_voxel(const float* in, int n, float voxel, float* out) {
  Cloud3f s, o; to_cloud(in, n, &s);
  rs_tracker::DownsampleVoxel(s, voxel, &o);
  std::memcpy(out, o.GetPtr(), sizeof(float) * 3 * o.GetNumPoints());
  return o.GetNumPoints();
}
void ref_find_correspondences(const float* dst, int m, const float* src, int n, int* idx, float* d2) {
  Cloud3f s, d; to_cloud(src, n, &s); to_cloud(dst, m, &d);
  const rs_tracker::KDTree3f tree{std::cref(d), 16};
  std::vector<int> i; std::vector<float> q;
  rs_tracker::FindCorrespondences(tree, s, &i, &q);
  std::memcpy(idx, i.data(), sizeof(int) * n); std::memcpy(d2, q.data(), sizeof(float) * n);
}
/* The caller's per-pair sequence (rs_replay_app.cpp:229,246-251) through the reference's own functions:
 * RemoveNans -> DownsampleVoxel(voxel) x2 -> AlignIcp3d(curr, prev, max_iter). Only the back-projection is
 * ours (in the reference it happens inside librealsense, rs_driver.cpp:201-202; invalid -> origin, :83-88).
 * One pair per OpenMP thread. */
void ref_align_depth_pairs(const unsigned short* src_depth, const unsigned short* dst_depth, int n_pairs, int w, int h,
                           float fx, float fy, float cx, float cy, float depth_scale, float voxel, int max_iter,
                           int n_threads, float* T, int* ok) {
  const size_t npx = (size_t)w * h;
  auto cloud_of = [&](const unsigned short* d, Cloud3f* out) {
    Cloud3f raw, clean;
    raw.SetNumPoints((int)npx);
    float* p = raw.GetPtr();
    for (int v = 0; v < h; ++v)
      for (int u = 0; u < w; ++u, p += 3) {
        const unsigned short dd = d[(size_t)v * w + u];
        if (dd == 0) { p[0] = p[1] = p[2] = 0.f; continue; }
        const float z = (float)dd * depth_scale;
        p[0] = ((float)u - cx) * z / fx; p[1] = ((float)v - cy) * z / fy; p[2] = z;
      }
    rs_tracker::RemoveNans(raw, &clean);
    if (voxel > 0) rs_tracker::DownsampleVoxel(clean, voxel, out); else *out = clean;
  };
#pragma omp parallel for schedule(dynamic, 1) num_threads(n_threads)
  for (int i = 0; i < n_pairs; ++i) {
    Cloud3f s, d;
    cloud_of(src_depth + npx * i, &s);
    cloud_of(dst_depth + npx * i, &d);
    Eigen::Isometry3f x; to_pose(T + 16 * i, &x);
    ok[i] = rs_tracker::AlignIcp3d(s, d, max_iter, &x) ? 1 : 0;
    from_pose(x, T + 16 * i);
  }
}

void ref_covariances(const float* pts, int n, int use_gicp, float* out) {
  Cloud3f s; to_cloud(pts, n, &s);
  const rs_tracker::KDTree3f tree{std::cref(s), 10};
  std::vector<Eigen::Matrix3f> covs;
  rs_tracker::ComputeCovariances(tree, s, &covs, use_gicp != 0);
  for (int i = 0; i < n; ++i)
    for (int r = 0; r < 3; ++r)
      for (int c = 0; c < 3; ++c) out[9 * i + 3 * r + c] = covs[i](r, c);
}

void ref_normals(const float* pts, int n, int k, const float* viewpoint, float* out) {
  Cloud3f s, nrm; to_cloud(pts, n, &s);
  const rs_tracker::KDTree3f tree{std::cref(s), 10};
  rs_tracker::ComputeNormals(s, tree, (float)k, &nrm);
  rs_tracker::OrientNormals(s, Eigen::Vector3f(viewpoint[0], viewpoint[1], viewpoint[2]), &nrm);
  std::memcpy(out, nrm.GetPtr(), sizeof(float) * 3 * n);
}

// ComputeExtents (point_cloud_utils.cpp:26-32; defined there, declared in no header)
void ref_extents(const float* pts, int n, float* lo, float* hi) {
  Cloud3f s; to_cloud(pts, n, &s);
  Eigen::AlignedBox3f box;
  rs_tracker::ComputeExtents(s, &box);
  const Eigen::Vector3f a = box.min(), b = box.max();
  for (int i = 0; i < 3; ++i) { lo[i] = a[i]; hi[i] = b[i]; }
}

// OrientNormals alone (point_cloud_utils.cpp:205-216) on caller-supplied normals, in place
void ref_orient_normals(const float* pts, int n, const float* viewpoint, float* normals_inout) {
  Cloud3f s, nrm; to_cloud(pts, n, &s); to_cloud(normals_inout, n, &nrm);
  rs_tracker::OrientNormals(s, Eigen::Vector3f(viewpoint[0], viewpoint[1], viewpoint[2]), &nrm);
  std::memcpy(normals_inout, nrm.GetPtr(), sizeof(float) * 3 * n);
}
}
