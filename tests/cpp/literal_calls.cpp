// The reference's own call expressions, verbatim, against include/rs_tracker/align/align_rgbd.hpp. Test code: built
// in the build container by `make -C oracle ref` (it includes the reference's own rs_tracker/common/types.hpp for
// Cloud3f / KDTree3f, and the stand-in Eigen / ChoUtil / nanoflann headers of oracle/shim) into
// oracle/_ref/literal_calls, and run on the GPU box by tests/test_cpp_api.py.
//   rs_replay_app.cpp:245-251   DownsampleVoxel(cloud, 0.05f, &curr_cloud_down); AlignIcp3d(curr_cloud_down, prev_cloud_down, 128, &xfm)
//   rs_align_app.cpp:295        SolveKabsch(src_cloud, dst_cloud, indices, weights, &xfm)
//   rs_align_app.cpp:303        AlignIcp3d(src_cloud, dst_cloud, 128, &xfm)
//   align_icp.cpp:165-166       AlignIcp3d(src, dst, dst_tree, max_iter, transform)
//   point_cloud_utils.hpp:14    FindCorrespondences(tree, source, &indices, &squared_distances)
//   rs_replay_app.cpp:382-387   KDTree3f tree{std::cref(cloud_transformed), 16}; ComputeNormals(cloud_transformed, tree, 16, &normals);
//                               OrientNormals(cloud_transformed, viewpoint, &normals)
//   align_gicp.cpp:120-121      std::vector<Eigen::Matrix3f> src_covs(src.GetNumPoints()); ComputeCovariances(*src_tree, src, &src_covs, false)
//   point_cloud_utils.hpp:16    ComputeCentroid(cloud, &centroid)   (align_icp.cpp:86)
//   point_cloud_utils.cpp:26    ComputeExtents(cloud, &box)
//   align_gicp.cpp:141-143      cost = ComputeAlignment(src, dst, src_covs, dst_covs, nn_indices, estimate, &delta_xfm)
//   rs_tracker.cpp:87           rs_tracker::ComputeAlignment(prev_cloud, curr_cloud, &transform)
#include <cmath>
#include <cstdio>
#include <memory>
#include <random>

#include "rs_tracker/common/types.hpp"       // the reference's own header: Cloud3f, KDTree3f
#include "rs_tracker/align/align_rgbd.hpp"   // instead of rs_tracker/align/align_icp.hpp + common/point_cloud_utils.hpp

static float translation_error(const Eigen::Isometry3f& x, const float* t) {
  const float* m = x.matrix().data();   // column-major 4x4
  const float dx = m[12] - t[0], dy = m[13] - t[1], dz = m[14] - t[2];
  return std::sqrt(dx * dx + dy * dy + dz * dz);
}

int main() {
  std::mt19937 rng(11);
  std::uniform_real_distribution<float> U(-1.f, 1.f);
  rs_tracker::Cloud3f src_cloud, dst_cloud;
  const int n = 2000;
  src_cloud.SetNumPoints(n);
  dst_cloud.SetNumPoints(n);
  // dst = R_x(0.02) R_y(-0.04) R_z(0.05) src + t   (the reference's disabled self-test, rs_align_app.cpp:257-263, scaled)
  const float rx = 0.02f, ry = -0.04f, rz = 0.05f;
  const float cxr = std::cos(rx), sxr = std::sin(rx), cyr = std::cos(ry), syr = std::sin(ry), czr = std::cos(rz), szr = std::sin(rz);
  const float R[9] = {cyr * czr, -cyr * szr, syr,
                      sxr * syr * czr + cxr * szr, -sxr * syr * szr + cxr * czr, -sxr * cyr,
                      -cxr * syr * czr + sxr * szr, cxr * syr * szr + sxr * czr, cxr * cyr};
  const float t[3] = {0.05f, -0.03f, 0.02f};
  float* s = src_cloud.GetPtr();
  float* d = dst_cloud.GetPtr();
  for (int i = 0; i < n; ++i) {
    const float p[3] = {U(rng), U(rng), U(rng)};
    for (int a = 0; a < 3; ++a) {
      s[3 * i + a] = p[a];
      d[3 * i + a] = R[3 * a] * p[0] + R[3 * a + 1] * p[1] + R[3 * a + 2] * p[2] + t[a];
    }
  }
  int failures = 0;

  // rs_align_app.cpp:290-303
  std::vector<std::pair<int, int>> indices;
  std::vector<float> weights;
  for (int i = 0; i < n; i += 3) indices.emplace_back(i, i);
  Eigen::Isometry3f xfm = Eigen::Isometry3f::Identity();
  {
    const bool suc =
        rs_tracker::SolveKabsch(src_cloud, dst_cloud, indices, weights, &xfm);
    if (!suc) { std::printf("kabsch failed\n"); ++failures; }
  }
  const float ek = translation_error(xfm, t);
  std::printf("kabsch translation error %.3e\n", ek);
  if (!(ek < 1e-4f)) ++failures;
  {
    const bool suc = rs_tracker::AlignIcp3d(src_cloud, dst_cloud, 128, &xfm);
    if (!suc) { std::printf("icp failed\n"); ++failures; }
  }
  const float ei = translation_error(xfm, t);
  std::printf("icp translation error %.3e\n", ei);
  if (!(ei < 1e-3f)) ++failures;

  // rs_replay_app.cpp:245-251
  {
    const rs_tracker::Cloud3f& cloud = src_cloud;
    const rs_tracker::Cloud3f& prev_cloud = dst_cloud;
    Eigen::Isometry3f xfm = Eigen::Isometry3f::Identity();
    cho::core::PointCloud<float, 3> curr_cloud_down, prev_cloud_down;
    rs_tracker::DownsampleVoxel(cloud, 0.05f, &curr_cloud_down);
    rs_tracker::DownsampleVoxel(prev_cloud, 0.05f, &prev_cloud_down);
    bool suc{false};
    suc =
        rs_tracker::AlignIcp3d(curr_cloud_down, prev_cloud_down, 128, &xfm);
    std::printf("replay-style: %d -> %d / %d points, suc %d\n", n, curr_cloud_down.GetNumPoints(), prev_cloud_down.GetNumPoints(), (int)suc);
    if (!suc || curr_cloud_down.GetNumPoints() < 3 || curr_cloud_down.GetNumPoints() > n) ++failures;
  }

  // align_icp.cpp:163-167 (the 5-argument form) and point_cloud_utils.hpp:14
  {
    const rs_tracker::KDTree3f dst_tree{std::cref(dst_cloud), 16};
    Eigen::Isometry3f x2 = Eigen::Isometry3f::Identity();
    const bool suc = rs_tracker::AlignIcp3d(src_cloud, dst_cloud, dst_tree, 128, &x2);
    if (!suc || translation_error(x2, t) > 1e-3f) { std::printf("5-argument form failed\n"); ++failures; }
    std::vector<int> idx;
    std::vector<float> d2;
    rs_tracker::FindCorrespondences(dst_tree, src_cloud, &idx, &d2);
    int cpu_idx[1];
    float cpu_d2[1];
    int bad = 0;
    for (int i = 0; i < n; i += 97) {
      dst_tree.query(s + 3 * i, 1, cpu_idx, cpu_d2);
      if (cpu_idx[0] != idx[i] || cpu_d2[0] != d2[i]) ++bad;
    }
    std::printf("FindCorrespondences: %zu results, %d mismatches against the k-d tree\n", idx.size(), bad);
    if ((int)idx.size() != n || bad) ++failures;
    rs_tracker::Cloud3f with_nan, clean;
    with_nan.SetNumPoints(n);
    for (int i = 0; i < 3 * n; ++i) with_nan.GetPtr()[i] = s[i];
    with_nan.GetPtr()[3 * 5 + 1] = std::nanf("");
    rs_tracker::RemoveNans(with_nan, &clean);
    if (clean.GetNumPoints() != n - 1) { std::printf("RemoveNans kept %d\n", clean.GetNumPoints()); ++failures; }
  }
  // rs_replay_app.cpp:382-387, align_gicp.cpp:114-121, align_icp.cpp:86
  {
    const rs_tracker::Cloud3f& cloud_transformed = dst_cloud;
    const Eigen::Vector3f viewpoint(0.f, 0.f, 5.f);
    rs_tracker::KDTree3f tree{std::cref(cloud_transformed), 16};
    rs_tracker::Cloud3f normals;
    rs_tracker::ComputeNormals(cloud_transformed, tree, 16, &normals);
    rs_tracker::OrientNormals(cloud_transformed, viewpoint, &normals);
    int bad = normals.GetNumPoints() != n;
    for (int i = 0; i < n && !bad; ++i) {
      const float* nv = normals.GetPtr() + 3 * i;
      const float len = std::sqrt(nv[0] * nv[0] + nv[1] * nv[1] + nv[2] * nv[2]);
      const float ray = (d[3 * i] - 0.f) * nv[0] + (d[3 * i + 1] - 0.f) * nv[1] + (d[3 * i + 2] - 5.f) * nv[2];
      if (!(std::fabs(len - 1.f) < 1e-3f) || ray > 1e-6f) ++bad;
    }
    std::printf("ComputeNormals + OrientNormals: %d normals, %d not unit / not facing the viewpoint\n", normals.GetNumPoints(), bad);
    if (bad) ++failures;

    const rs_tracker::Cloud3f& src = src_cloud;
    std::shared_ptr<rs_tracker::KDTree3f> src_tree = std::make_shared<rs_tracker::KDTree3f>(std::cref(src), 10);
    std::vector<Eigen::Matrix3f> src_covs(src.GetNumPoints());
    rs_tracker::ComputeCovariances(*src_tree, src, &src_covs, false);
    int bad_cov = (int)src_covs.size() != n;
    for (int i = 0; i < n && !bad_cov; i += 53) {
      const Eigen::Matrix3f& C = src_covs[i];
      const float tr = C(0, 0) + C(1, 1) + C(2, 2);
      if (!(tr > 0.f) || !(tr < 1.f) || C(0, 1) != C(1, 0) || C(0, 2) != C(2, 0) || C(1, 2) != C(2, 1)) ++bad_cov;
    }
    std::printf("ComputeCovariances: %zu matrices, %d not symmetric / not positive\n", src_covs.size(), bad_cov);
    if (bad_cov) ++failures;

    // align_gicp.cpp:122-143 (one round) and rs_tracker.cpp:87
    {
      const rs_tracker::Cloud3f& dst = dst_cloud;
      std::shared_ptr<rs_tracker::KDTree3f> dst_tree = std::make_shared<rs_tracker::KDTree3f>(std::cref(dst), 10);
      std::vector<Eigen::Matrix3f> dst_covs(dst.GetNumPoints());
      rs_tracker::ComputeCovariances(*dst_tree, dst, &dst_covs, false);
      // rs_tracker.cpp:87
      const rs_tracker::Cloud3f& prev_cloud = src_cloud;
      const rs_tracker::Cloud3f& curr_cloud = dst_cloud;
      Eigen::Isometry3f transform = Eigen::Isometry3f::Identity();
      const float c3 =
        rs_tracker::ComputeAlignment(prev_cloud, curr_cloud, &transform);
      const float e3 = translation_error(transform, t);
      std::printf("ComputeAlignment (3 arguments): cost %.4e, translation error %.3e\n", c3, e3);
      if (!std::isfinite(c3) || !(e3 < 0.02f)) ++failures;

      // one more round of the loop of align_gicp.cpp:128-159 from that estimate
      Eigen::Isometry3f estimate = transform;
      rs_tracker::Cloud3f tmp;
      tmp.SetNumPoints(src.GetNumPoints());
      const float* E = estimate.matrix().data();
      for (int i = 0; i < n; ++i)
        for (int a = 0; a < 3; ++a)
          tmp.GetPtr()[3 * i + a] = E[a] * s[3 * i] + E[4 + a] * s[3 * i + 1] + E[8 + a] * s[3 * i + 2] + E[12 + a];
      std::vector<int> nn_indices;
      std::vector<float> nn_sq_dists;
      rs_tracker::FindCorrespondences(*dst_tree, tmp, &nn_indices, &nn_sq_dists);
      float cost{0};
      Eigen::Isometry3f delta_xfm = Eigen::Isometry3f::Identity();
      cost = rs_tracker::ComputeAlignment(src, dst, src_covs, dst_covs, nn_indices, estimate,
                              &delta_xfm);
      const float e7 = translation_error(delta_xfm, t);
      std::printf("ComputeAlignment (7 arguments): cost %.4e, translation error %.3e\n", cost, e7);
      if (!(cost >= 0.f) || !std::isfinite(cost) || !(e7 < e3 + 1e-3f)) ++failures;
    }

    // kdtree.hpp:51-57 on the device-resident tree: dst_tree.query(p.data(), 1, &j, &dist_sqr) (align_icp.cpp:112) and k = 4
    {
      const rs_tracker::KDTree3f dst_tree{std::cref(dst_cloud), 16};
      rs_tracker::GpuKDTree3f gpu_tree(rs_tracker::DefaultAlignContext(), dst_cloud);
      int bad_q = gpu_tree.size() != n;
      for (int i = 0; i < n; i += 211) {
        int j = -1, gj[4], cj[4];
        float dist_sqr = 0.f, gd[4], cd[4];
        gpu_tree.query(s + 3 * i, 1, &j, &dist_sqr);
        dst_tree.query(s + 3 * i, 1, cj, cd);
        if (j != cj[0] || dist_sqr != cd[0]) ++bad_q;
        gpu_tree.query(s + 3 * i, 4, gj, gd);
        dst_tree.query(s + 3 * i, 4, cj, cd);
        for (int e = 0; e < 4; ++e) if (gj[e] != cj[e] || gd[e] != cd[e]) ++bad_q;
      }
      std::vector<int> all_idx;
      std::vector<float> all_d2;
      if (!gpu_tree.query(src_cloud, 2, &all_idx, &all_d2) || (int)all_idx.size() != 2 * n) ++bad_q;
      std::printf("GpuKDTree3f::query: %d mismatches against the k-d tree\n", bad_q);
      if (bad_q) ++failures;
    }

    // point_cloud_utils.cpp:26-32
    {
      Eigen::AlignedBox3f box;
      rs_tracker::ComputeExtents(src, &box);
      float lo[3] = {1e30f, 1e30f, 1e30f}, hi[3] = {-1e30f, -1e30f, -1e30f};
      for (int i = 0; i < n; ++i) for (int a = 0; a < 3; ++a) { lo[a] = std::fmin(lo[a], s[3 * i + a]); hi[a] = std::fmax(hi[a], s[3 * i + a]); }
      const Eigen::Vector3f bl = box.min(), bh = box.max();
      int bad_box = 0;
      for (int a = 0; a < 3; ++a) bad_box += (bl[a] != lo[a]) + (bh[a] != hi[a]);
      std::printf("ComputeExtents: %d of 6 bounds differ\n", bad_box);
      if (bad_box) ++failures;
    }

    Eigen::Vector3f src_mean;
    rs_tracker::ComputeCentroid(src, &src_mean);
    double ref[3] = {0, 0, 0};
    for (int i = 0; i < n; ++i) for (int a = 0; a < 3; ++a) ref[a] += s[3 * i + a];
    float ec = 0.f;
    for (int a = 0; a < 3; ++a) ec = std::fmax(ec, std::fabs(src_mean(a) - (float)(ref[a] / n)));
    std::printf("ComputeCentroid error %.3e\n", ec);
    if (!(ec < 1e-6f)) ++failures;
  }
  std::printf("failures %d\n", failures);
  return failures;
}
