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
void ref_normals(const float* pts, int n, int k, const float* viewpoint, float* out) {
  Cloud3f s, nrm; to_cloud(pts, n, &s);
  const rs_tracker::KDTree3f tree{std::cref(s), 10};
  rs_tracker::ComputeNormals(s, tree, (float)k, &nrm);
  rs_tracker::OrientNormals(s, Eigen::Vector3f(viewpoint[0], viewpoint[1], viewpoint[2]), &nrm);
  std::memcpy(out, nrm.GetPtr(), sizeof(float) * 3 * n);
}
}
