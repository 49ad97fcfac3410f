// align_rgbd.hpp — C++ host API of the B200 alignment engine, in the reference's idiom.
//
// Frame-based sibling of the reference's
//   bool AlignIcp3d(const Cloud3f& src, const Cloud3f& dst, const int max_iter,
//                   Eigen::Isometry3f* const transform);
//   (rs_tracker/align/include/rs_tracker/align/align_icp.hpp:22-24)
// Same conventions: free functions in namespace rs_tracker, inputs by const reference, outputs
// last as `T* const`, `bool` = success, and `transform` is READ as the initial guess and
// OVERWRITTEN with the result (align_icp.cpp:82,156); it maps src -> dst (align_icp.cpp:107).
// Frame-level inputs are what the driver delivers (rs_driver.hpp:17-22): CV_16UC1 depth,
// CV_8UC3 colour, and the 3x3 K of rs_driver.cpp:264-280.
//
// Header-only over the C ABI (include/rst_align.h); link with librst_align.so. No Eigen/OpenCV
// dependency: Pose is 16 floats, column-major, bit-compatible with Eigen::Isometry3f::matrix();
// when <Eigen/Geometry> is available an Isometry3f overload is provided.
#pragma once

#include <array>
#include <cstdint>
#include <cstring>
#include <limits>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "rst_align.h"

#if defined(__has_include)
#if __has_include(<Eigen/Geometry>)
#include <Eigen/Geometry>
#define RS_TRACKER_HAVE_EIGEN 1
#endif
#endif

namespace rs_tracker {

/// One RGB-D frame as the driver hands it out (rs_driver.hpp:17-19): depth CV_16UC1, colour CV_8UC3.
struct DepthFrame {
  const std::uint16_t* depth{nullptr};
  const std::uint8_t* rgb{nullptr};
  int width{0}, height{0};
  int depth_stride_bytes{0};  ///< 0 = tightly packed
  int rgb_stride_bytes{0};
  rst_frame c() const {
    return rst_frame{depth, rgb, width, height, depth_stride_bytes ? depth_stride_bytes : width * 2,
                     rgb_stride_bytes ? rgb_stride_bytes : width * 3};
  }
};

/// Pin-hole intrinsics; FromMatrix takes the column-major 3x3 of RsDriver::GetIntrinsicMatrix()
/// (rs_driver.cpp:264-280): K = [[fx,0,ppx],[0,fy,ppy],[0,0,1]].
struct Intrinsics {
  float fx{0}, fy{0}, cx{0}, cy{0};
  static Intrinsics FromMatrix(const float* K_colmajor3x3) {
    return Intrinsics{K_colmajor3x3[0], K_colmajor3x3[4], K_colmajor3x3[6], K_colmajor3x3[7]};
  }
  rst_intrinsics c() const { return rst_intrinsics{fx, fy, cx, cy}; }
};

/// 4x4 rigid transform, column-major fp32 — the memory layout of Eigen::Isometry3f::matrix().
struct Pose {
  std::array<float, 16> m{{1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1}};
  static Pose Identity() { return Pose{}; }
  float& operator()(int r, int c) { return m[r + 4 * c]; }
  float operator()(int r, int c) const { return m[r + 4 * c]; }
  /// this * other (the replay loop chains total_xfm = total_xfm * xfm, rs_replay_app.cpp:267)
  Pose operator*(const Pose& o) const {
    Pose r;
    for (int i = 0; i < 4; ++i)
      for (int j = 0; j < 4; ++j) {
        float s = 0.f;
        for (int k = 0; k < 4; ++k) s += (*this)(i, k) * o(k, j);
        r(i, j) = s;
      }
    return r;
  }
};

/// Algorithm constants (the reference hard-codes its own, align_icp.cpp:91,96-98,165).
struct AlignParams : rst_params {
  AlignParams() { rst_params_default(this); }
};

using AlignStats = rst_stats;

/// One alignment context per GPU (rst_ctx). Not thread-safe; contexts are independent.
class AlignContext {
 public:
  AlignContext(int device, int max_width, int max_height, int max_frames, int max_pairs, void* stream = nullptr) {
    const int rc = rst_ctx_create(device, max_width, max_height, max_frames, max_pairs, stream, &ctx_);
    if (rc != RST_OK) throw std::runtime_error(std::string("rst_ctx_create: ") + rst_last_create_error());
  }
  ~AlignContext() { rst_ctx_destroy(ctx_); }
  AlignContext(const AlignContext&) = delete;
  AlignContext& operator=(const AlignContext&) = delete;
  rst_ctx* get() const { return ctx_; }
  const char* LastError() const { return rst_last_error(ctx_); }

 private:
  rst_ctx* ctx_{nullptr};
};

/// Aligns `src` onto `dst`: p_dst ~ transform * p_src. Returns false when the call fails or the
/// pair's status word is not RST_STATUS_OK (too few associations / degenerate / non-finite —
/// the reference's `false` for < 3 points, align_icp.cpp:77-79).
inline bool AlignRgbd(AlignContext& ctx, const DepthFrame& src, const DepthFrame& dst, const Intrinsics& K,
                      const AlignParams& params, Pose* const transform, AlignStats* const stats = nullptr) {
  const rst_frame s = src.c(), d = dst.c();
  const rst_intrinsics k = K.c();
  AlignStats st{};
  const int rc = rst_align_pairs(ctx.get(), &s, &d, 1, &k, &params, transform->m.data(), &st);
  if (stats) *stats = st;
  return rc == RST_OK && st.status == RST_STATUS_OK;
}

/// Batched overload: pair i aligns src[i] onto dst[i]; transforms in/out, stats optional.
/// Returns true when the call ran and every pair succeeded; per-pair status is in `stats`.
inline bool AlignRgbd(AlignContext& ctx, const std::vector<DepthFrame>& src, const std::vector<DepthFrame>& dst,
                      const Intrinsics& K, const AlignParams& params, std::vector<Pose>* const transforms,
                      std::vector<AlignStats>* const stats = nullptr) {
  const int n = static_cast<int>(src.size());
  if (dst.size() != src.size() || static_cast<int>(transforms->size()) != n) return false;
  std::vector<rst_frame> s(n), d(n);
  for (int i = 0; i < n; ++i) { s[i] = src[i].c(); d[i] = dst[i].c(); }
  std::vector<AlignStats> st(n);
  const rst_intrinsics k = K.c();
  static_assert(sizeof(Pose) == 16 * sizeof(float), "Pose must be 16 packed floats");
  const int rc = rst_align_pairs(ctx.get(), s.data(), d.data(), n, &k, &params,
                                 n ? (*transforms)[0].m.data() : nullptr, st.data());
  bool ok = rc == RST_OK;
  for (const auto& x : st) ok = ok && x.status == RST_STATUS_OK;
  if (stats) *stats = std::move(st);
  return ok;
}

/// Frame-to-frame odometry over a sequence — the loop body of rs_replay_app.cpp:246-251 for all
/// consecutive frames at once: (*transforms)[i] = T_{i <- i+1} (AlignIcp3d(curr, prev)).
inline bool AlignSequence(AlignContext& ctx, const std::vector<DepthFrame>& frames, const Intrinsics& K,
                          const AlignParams& params, std::vector<Pose>* const transforms,
                          std::vector<AlignStats>* const stats = nullptr) {
  const int n = static_cast<int>(frames.size());
  if (n < 2 || static_cast<int>(transforms->size()) != n - 1) return false;
  std::vector<rst_frame> f(n);
  for (int i = 0; i < n; ++i) f[i] = frames[i].c();
  std::vector<AlignStats> st(n - 1);
  const rst_intrinsics k = K.c();
  const int rc = rst_align_sequence(ctx.get(), f.data(), n, &k, &params, (*transforms)[0].m.data(), st.data());
  bool ok = rc == RST_OK;
  for (const auto& x : st) ok = ok && x.status == RST_STATUS_OK;
  if (stats) *stats = std::move(st);
  return ok;
}

// ------------------------------------------------------------------------------------------------
// Cloud-based calls with the reference's own algorithm (align_icp.hpp:14-24), on the GPU.
// `Cloud` is any type with GetPtr() -> const float* (xyz interleaved) and GetNumPoints() — the
// reference's Cloud3f = cho::core::PointCloud<float,3> qualifies as is (rs_viewer.cpp:91-92).
// ------------------------------------------------------------------------------------------------

/// bool AlignIcp3d(src, dst, max_iter, &transform)  align_icp.hpp:22-24 / align_icp.cpp:73-167.
template <class Cloud>
inline bool AlignIcp3d(AlignContext& ctx, const Cloud& src, const Cloud& dst, const int max_iter, Pose* const transform,
                       float* const mean_cost = nullptr) {
  const rst_cloud s{src.GetPtr(), static_cast<std::int32_t>(src.GetNumPoints())};
  const rst_cloud d{dst.GetPtr(), static_cast<std::int32_t>(dst.GetNumPoints())};
  rst_icp3d_result res{};
  const int rc = rst_icp3d_pairs(ctx.get(), &s, &d, 1, max_iter, /*grid_cell=*/0.f, transform->m.data(), &res, nullptr, nullptr);
  if (mean_cost) *mean_cost = res.mean_cost;
  return rc == RST_OK && res.ok != 0;
}

/// bool SolveKabsch(src, dst, indices, weights, &xfm)  align_icp.hpp:14-18 / align_icp.cpp:18-71.
template <class Cloud>
inline bool SolveKabsch(AlignContext& ctx, const Cloud& src, const Cloud& dst, const std::vector<std::pair<int, int>>& indices,
                        const std::vector<float>& weights, Pose* const xfm) {
  static_assert(sizeof(std::pair<int, int>) == 2 * sizeof(std::int32_t), "pair<int,int> must be two packed int32");
  const rst_cloud s{src.GetPtr(), static_cast<std::int32_t>(src.GetNumPoints())};
  const rst_cloud d{dst.GetPtr(), static_cast<std::int32_t>(dst.GetNumPoints())};
  std::int32_t ok = 0;
  const int rc = rst_solve_kabsch(ctx.get(), &s, &d, indices.empty() ? nullptr : &indices[0].first, static_cast<std::int32_t>(indices.size()),
                                  weights.empty() ? nullptr : weights.data(), xfm->m.data(), &ok);
  return rc == RST_OK && ok != 0;
}

// ------------------------------------------------------------------------------------------------
// Cloud utilities of rs_tracker/common/point_cloud_utils.hpp:11-31 on the GPU, same signatures. `Cloud` as above,
// plus SetNumPoints(n) and a mutable GetPtr() for outputs (cho::core::PointCloud has both).
// ------------------------------------------------------------------------------------------------

/// void DownsampleVoxel(cloud_in, voxel_size, &cloud_out)  point_cloud_utils.cpp:34-68 (cloud_out may alias cloud_in).
template <class Cloud>
inline void DownsampleVoxel(AlignContext& ctx, const Cloud& cloud_in, const float voxel_size, Cloud* const cloud_out) {
  const rst_cloud in{cloud_in.GetPtr(), static_cast<std::int32_t>(cloud_in.GetNumPoints())};
  std::vector<float> out(3 * static_cast<std::size_t>(in.n));
  std::int32_t n = 0;
  if (rst_downsample_voxel(ctx.get(), &in, voxel_size, out.data(), &n) != RST_OK) n = 0;
  cloud_out->SetNumPoints(n);
  if (n) std::memcpy(cloud_out->GetPtr(), out.data(), sizeof(float) * 3 * static_cast<std::size_t>(n));
}

/// void RemoveNans(cloud_in, &cloud_out)  point_cloud_utils.cpp:163-174.
template <class Cloud>
inline void RemoveNans(AlignContext& ctx, const Cloud& cloud_in, Cloud* const cloud_out) {
  const rst_cloud in{cloud_in.GetPtr(), static_cast<std::int32_t>(cloud_in.GetNumPoints())};
  std::vector<float> out(3 * static_cast<std::size_t>(in.n));
  std::int32_t n = 0;
  if (rst_remove_nans(ctx.get(), &in, out.data(), &n) != RST_OK) n = 0;
  cloud_out->SetNumPoints(n);
  if (n) std::memcpy(cloud_out->GetPtr(), out.data(), sizeof(float) * 3 * static_cast<std::size_t>(n));
}

/// void FindCorrespondences(tree, source, &indices, &squared_distances)  point_cloud_utils.cpp:70-90; `target` is the
/// cloud the reference's KDTree3f was built over.
template <class Cloud>
inline void FindCorrespondences(AlignContext& ctx, const Cloud& target, const Cloud& source, std::vector<int>* const indices,
                                std::vector<float>* const squared_distances) {
  const rst_cloud t{target.GetPtr(), static_cast<std::int32_t>(target.GetNumPoints())};
  const rst_cloud s{source.GetPtr(), static_cast<std::int32_t>(source.GetNumPoints())};
  static_assert(sizeof(int) == sizeof(std::int32_t), "int must be 32-bit");
  indices->assign(static_cast<std::size_t>(s.n), -1);
  squared_distances->assign(static_cast<std::size_t>(s.n), 0.f);
  if (s.n) rst_find_correspondences(ctx.get(), &t, &s, /*grid_cell=*/0.f, indices->data(), squared_distances->data());
}

/// void ComputeCentroid(cloud, &centroid)  point_cloud_utils.cpp:92-98; centroid_xyz: 3 floats.
template <class Cloud>
inline void ComputeCentroid(AlignContext& ctx, const Cloud& cloud, float* const centroid_xyz) {
  const rst_cloud c{cloud.GetPtr(), static_cast<std::int32_t>(cloud.GetNumPoints())};
  if (rst_cloud_centroid(ctx.get(), &c, centroid_xyz) != RST_OK)
    centroid_xyz[0] = centroid_xyz[1] = centroid_xyz[2] = std::numeric_limits<float>::quiet_NaN();   // 0 / 0 in the reference
}

/// void ComputeExtents(cloud, &box)  point_cloud_utils.cpp:26-32; lo_xyz / hi_xyz: 3 floats each (empty cloud: FLT_MAX / -FLT_MAX).
template <class Cloud>
inline void ComputeExtents(AlignContext& ctx, const Cloud& cloud, float* const lo_xyz, float* const hi_xyz) {
  const rst_cloud c{cloud.GetPtr(), static_cast<std::int32_t>(cloud.GetNumPoints())};
  if (rst_cloud_extents(ctx.get(), &c, lo_xyz, hi_xyz) != RST_OK)
    for (int a = 0; a < 3; ++a) { lo_xyz[a] = std::numeric_limits<float>::max(); hi_xyz[a] = std::numeric_limits<float>::lowest(); }
}

/// void ComputeCovariances(tree, cloud, &covs, use_gicp)  point_cloud_utils.cpp:100-161; covs_9n: n x 9 floats
/// (symmetric 3x3 each, so row- and column-major agree).
template <class Cloud>
inline void ComputeCovariances(AlignContext& ctx, const Cloud& cloud, float* const covs_9n, const bool use_gicp) {
  const rst_cloud c{cloud.GetPtr(), static_cast<std::int32_t>(cloud.GetNumPoints())};
  rst_cloud_covariances(ctx.get(), &c, use_gicp ? 1 : 0, /*grid_cell=*/0.f, covs_9n);
}

/// void ComputeNormals(cloud, tree, num_neighbors, &normals)  point_cloud_utils.cpp:176-203. The reference leaves the
/// sign of each normal to its eigen-solver; here every normal comes back pointing towards the origin (one of those
/// signs), and OrientNormals re-orients it for any other viewpoint.
template <class Cloud>
inline void ComputeNormals(AlignContext& ctx, const Cloud& cloud, const float num_neighbors, Cloud* const normals) {
  const rst_cloud c{cloud.GetPtr(), static_cast<std::int32_t>(cloud.GetNumPoints())};
  const float origin[3] = {0.f, 0.f, 0.f};
  normals->SetNumPoints(c.n);
  if (c.n && rst_cloud_normals(ctx.get(), &c, static_cast<std::int32_t>(num_neighbors), origin, /*grid_cell=*/0.f, normals->GetPtr()) != RST_OK)
    normals->SetNumPoints(0);   // e.g. num_neighbors outside [2, 32]: no normals rather than unwritten ones (ctx.LastError() says why)
}

/// void OrientNormals(cloud, viewpoint, &normals)  point_cloud_utils.cpp:205-216; viewpoint_xyz: 3 floats.
template <class Cloud>
inline void OrientNormals(AlignContext& ctx, const Cloud& cloud, const float* const viewpoint_xyz, Cloud* const normals) {
  const rst_cloud c{cloud.GetPtr(), static_cast<std::int32_t>(cloud.GetNumPoints())};
  if (c.n && normals->GetNumPoints() == cloud.GetNumPoints()) rst_orient_normals(ctx.get(), &c, viewpoint_xyz, normals->GetPtr());
}

/// The reference's KDTree3f (kdtree.hpp:11-99) on the GPU: built once over a cloud, queried any number of times.
/// `query` has the signature of KDTreeChoCloudAdaptor::query (kdtree.hpp:51-57); the batched form answers all points of
/// a cloud in one launch (row i of the n x k outputs = query i).
class GpuKDTree3f {
 public:
  template <class Cloud>
  GpuKDTree3f(AlignContext& ctx, const Cloud& cloud) : ctx_(ctx) {
    const rst_cloud c{cloud.GetPtr(), static_cast<std::int32_t>(cloud.GetNumPoints())};
    if (rst_tree_create(ctx.get(), &c, /*grid_cell=*/0.f, &tree_) != RST_OK) throw std::runtime_error(std::string("rst_tree_create: ") + ctx.LastError());
  }
  ~GpuKDTree3f() { rst_tree_destroy(tree_); }
  GpuKDTree3f(const GpuKDTree3f&) = delete;
  GpuKDTree3f& operator=(const GpuKDTree3f&) = delete;
  inline void query(const float* query_point, const std::size_t num_closest, int* out_indices, float* out_distances_sq,
                    const int /* nChecks_IGNORED */ = 10) const {
    rst_tree_query(ctx_.get(), tree_, query_point, 1, static_cast<std::int32_t>(num_closest), out_indices, out_distances_sq);
  }
  template <class Cloud>
  inline bool query(const Cloud& queries, const std::size_t num_closest, std::vector<int>* const out_indices,
                    std::vector<float>* const out_distances_sq) const {
    const std::size_t n = static_cast<std::size_t>(queries.GetNumPoints());
    out_indices->assign(n * num_closest, -1);
    out_distances_sq->assign(n * num_closest, std::numeric_limits<float>::infinity());
    return rst_tree_query(ctx_.get(), tree_, queries.GetPtr(), static_cast<std::int32_t>(n), static_cast<std::int32_t>(num_closest),
                          out_indices->data(), out_distances_sq->data()) == RST_OK;
  }
  int size() const { return rst_tree_size(tree_); }
  const rst_tree* get() const { return tree_; }

 private:
  AlignContext& ctx_;
  rst_tree* tree_{nullptr};
};

/// The process-wide context the literal (context-free) signatures below run on: created on first use on CUDA device
/// RS_TRACKER_ALIGN_DEVICE (default 0). The cloud engine sizes its own device memory per call, so the frame capacity
/// of this context is minimal; use an explicit AlignContext for the frame-based calls.
#ifndef RS_TRACKER_ALIGN_DEVICE
#define RS_TRACKER_ALIGN_DEVICE 0
#endif
inline AlignContext& DefaultAlignContext() {
  static AlignContext ctx(RS_TRACKER_ALIGN_DEVICE, 16, 16, 2, 1);
  return ctx;
}

#ifdef RS_TRACKER_HAVE_EIGEN
// ------------------------------------------------------------------------------------------------
// The reference's LITERAL signatures (align_icp.hpp:14-24, point_cloud_utils.hpp:11-31): the call sites
// rs_replay_app.cpp:246-251 and rs_align_app.cpp:295,303 compile against this header unchanged.
// ------------------------------------------------------------------------------------------------
namespace detail {
inline Pose ToPose(const Eigen::Isometry3f& x) { Pose p; std::memcpy(p.m.data(), x.matrix().data(), sizeof(float) * 16); return p; }
inline void FromPose(const Pose& p, Eigen::Isometry3f* const x) { std::memcpy(x->matrix().data(), p.m.data(), sizeof(float) * 16); }
}  // namespace detail

/// bool AlignIcp3d(const Cloud3f& src, const Cloud3f& dst, const int max_iter, Eigen::Isometry3f* const transform)
template <class Cloud>
inline bool AlignIcp3d(const Cloud& src, const Cloud& dst, const int max_iter, Eigen::Isometry3f* const transform) {
  Pose p = detail::ToPose(*transform);   // read as the initial guess (align_icp.cpp:82)
  const bool ok = AlignIcp3d(DefaultAlignContext(), src, dst, max_iter, &p);
  detail::FromPose(p, transform);        // overwritten with the result (:156)
  return ok;
}

/// bool AlignIcp3d(src, dst, const KDTree3f& dst_tree, max_iter, transform): the tree is the reference's search
/// structure over `dst`; the GPU builds its own grid over `dst`, so it is accepted and not used.
template <class Cloud, class Tree>
inline bool AlignIcp3d(const Cloud& src, const Cloud& dst, const Tree& /*dst_tree*/, const int max_iter,
                       Eigen::Isometry3f* const transform) {
  return AlignIcp3d(src, dst, max_iter, transform);
}

/// bool SolveKabsch(src, dst, indices, weights, Eigen::Isometry3f* const xfm)
template <class Cloud>
inline bool SolveKabsch(const Cloud& src, const Cloud& dst, const std::vector<std::pair<int, int>>& indices,
                        const std::vector<float>& weights, Eigen::Isometry3f* const xfm) {
  Pose p;
  const bool ok = SolveKabsch(DefaultAlignContext(), src, dst, indices, weights, &p);
  if (ok) detail::FromPose(p, xfm);
  return ok;
}

template <class Cloud>
inline void DownsampleVoxel(const Cloud& cloud_in, const float voxel_size, Cloud* const cloud_out) {
  DownsampleVoxel(DefaultAlignContext(), cloud_in, voxel_size, cloud_out);
}
template <class Cloud>
inline void RemoveNans(const Cloud& cloud_in, Cloud* const cloud_out) {
  RemoveNans(DefaultAlignContext(), cloud_in, cloud_out);
}
/// void FindCorrespondences(const KDTree3f& tree, const Cloud3f& source, ...): the tree names its cloud (kdtree.hpp:43).
template <class Tree, class Cloud>
inline auto FindCorrespondences(const Tree& tree, const Cloud& source, std::vector<int>* const indices,
                                std::vector<float>* const squared_distances) -> decltype(tree.m_cloud.get(), void()) {
  FindCorrespondences(DefaultAlignContext(), tree.m_cloud.get(), source, indices, squared_distances);
}

/// void ComputeCentroid(const Cloud3f& cloud, Eigen::Vector3f* const centroid)  (align_icp.cpp:86 calls it)
template <class Cloud>
inline void ComputeCentroid(const Cloud& cloud, Eigen::Vector3f* const centroid) {
  ComputeCentroid(DefaultAlignContext(), cloud, centroid->data());
}
/// void ComputeExtents(const Cloud3f& cloud, Eigen::AlignedBox3f* const box)  point_cloud_utils.cpp:26-32
template <class Cloud, class Box>
inline auto ComputeExtents(const Cloud& cloud, Box* const box) -> decltype(box->setEmpty(), void()) {
  float lo[3], hi[3];
  ComputeExtents(DefaultAlignContext(), cloud, lo, hi);
  box->setEmpty();
  if (lo[0] <= hi[0] && lo[1] <= hi[1] && lo[2] <= hi[2]) {
    box->extend(Eigen::Vector3f(lo[0], lo[1], lo[2]));
    box->extend(Eigen::Vector3f(hi[0], hi[1], hi[2]));
  }
}
/// void ComputeCovariances(const KDTree3f& tree, const Cloud3f& cloud, std::vector<Eigen::Matrix3f>* const covs,
/// const bool use_gicp)  (align_gicp.cpp:121,123); the tree is accepted and not used (the GPU grids the cloud itself).
template <class Tree, class Cloud>
inline void ComputeCovariances(const Tree& /*tree*/, const Cloud& cloud, std::vector<Eigen::Matrix3f>* const covs, const bool use_gicp) {
  const std::size_t n = static_cast<std::size_t>(cloud.GetNumPoints());
  std::vector<float> flat(9 * n);
  if (n) ComputeCovariances(DefaultAlignContext(), cloud, flat.data(), use_gicp);
  covs->resize(n);
  for (std::size_t i = 0; i < n; ++i) std::memcpy((*covs)[i].data(), flat.data() + 9 * i, sizeof(float) * 9);
}
/// void ComputeNormals(const Cloud3f& cloud, const KDTree3f& tree, const float num_neighbors, Cloud3f* const normals)
/// (rs_replay_app.cpp:386)
template <class Cloud, class Tree>
inline void ComputeNormals(const Cloud& cloud, const Tree& /*tree*/, const float num_neighbors, Cloud* const normals) {
  ComputeNormals(DefaultAlignContext(), cloud, num_neighbors, normals);
}
/// void OrientNormals(const Cloud3f& cloud, const Vector3f& viewpoint, Cloud3f* const normals)  (rs_replay_app.cpp:387)
template <class Cloud>
inline void OrientNormals(const Cloud& cloud, const Eigen::Vector3f& viewpoint, Cloud* const normals) {
  OrientNormals(DefaultAlignContext(), cloud, viewpoint.data(), normals);
}

/// float ComputeAlignment(src, dst, src_covs, dst_covs, dst_indices, seed, transform)  align_gicp.hpp:15-21 /
/// align_gicp.cpp:41-117: the minimiser of the Huber(0.5) GICP cost over the given correspondences, from `seed`;
/// returns the final cost. kGicpInnerIterations Levenberg-Marquardt steps stand where the reference runs Ceres.
#ifndef RS_TRACKER_GICP_INNER_ITERATIONS
#define RS_TRACKER_GICP_INNER_ITERATIONS 32
#endif
template <class Cloud>
inline float ComputeAlignment(const Cloud& src, const Cloud& dst, const std::vector<Eigen::Matrix3f>& src_covs,
                              const std::vector<Eigen::Matrix3f>& dst_covs, const std::vector<int>& dst_indices,
                              const Eigen::Isometry3f& seed, Eigen::Isometry3f* const transform) {
  const rst_cloud s{src.GetPtr(), static_cast<std::int32_t>(src.GetNumPoints())};
  const rst_cloud d{dst.GetPtr(), static_cast<std::int32_t>(dst.GetNumPoints())};
  if (src_covs.size() < static_cast<std::size_t>(s.n) || dst_covs.size() < static_cast<std::size_t>(d.n) ||
      dst_indices.size() < static_cast<std::size_t>(s.n))
    return std::numeric_limits<float>::infinity();
  std::vector<float> sc(9 * static_cast<std::size_t>(s.n)), dc(9 * static_cast<std::size_t>(d.n));
  for (std::size_t i = 0; i < static_cast<std::size_t>(s.n); ++i) std::memcpy(sc.data() + 9 * i, src_covs[i].data(), sizeof(float) * 9);
  for (std::size_t i = 0; i < static_cast<std::size_t>(d.n); ++i) std::memcpy(dc.data() + 9 * i, dst_covs[i].data(), sizeof(float) * 9);
  Pose p = detail::ToPose(seed);
  rst_gicp_stats st{};
  const int rc = rst_gicp_minimize(DefaultAlignContext().get(), &s, &d, sc.data(), dc.data(), dst_indices.data(),
                                   RS_TRACKER_GICP_INNER_ITERATIONS, /*huber_delta=*/0.5f, p.m.data(), &st);
  if (rc != RST_OK) return std::numeric_limits<float>::infinity();
  detail::FromPose(p, transform);
  return static_cast<float>(st.cost);
}

/// float ComputeAlignment(src, dst, transform)  align_gicp.hpp:23-25 / align_gicp.cpp:119-163 (rs_tracker.cpp:87,
/// rs_replay_app.cpp:253): sample covariances, 16 rounds of { FindCorrespondences; minimise }, from the identity.
template <class Cloud>
inline float ComputeAlignment(const Cloud& src, const Cloud& dst, Eigen::Isometry3f* const transform) {
  const rst_cloud s{src.GetPtr(), static_cast<std::int32_t>(src.GetNumPoints())};
  const rst_cloud d{dst.GetPtr(), static_cast<std::int32_t>(dst.GetNumPoints())};
  Pose p;   // the reference always starts from the identity (:147)
  rst_gicp_stats st{};
  const int rc = rst_gicp_align(DefaultAlignContext().get(), &s, &d, /*max_outer=*/16, /*inner_iters=*/8, /*huber_delta=*/0.5f,
                                /*use_gicp_covariances=*/0, /*grid_cell=*/0.f, p.m.data(), &st);
  if (rc != RST_OK) return std::numeric_limits<float>::infinity();   // the reference's answer to a non-finite estimate (:145-150)
  detail::FromPose(p, transform);
  return static_cast<float>(st.cost);
}

/// Drop-in signature for the reference's call sites (rs_replay_app.cpp:251, rs_align_app.cpp:303).
inline bool AlignRgbd(AlignContext& ctx, const DepthFrame& src, const DepthFrame& dst, const Eigen::Matrix3f& K,
                      const AlignParams& params, Eigen::Isometry3f* const transform, AlignStats* const stats = nullptr) {
  Pose p;
  std::memcpy(p.m.data(), transform->matrix().data(), sizeof(float) * 16);
  const bool ok = AlignRgbd(ctx, src, dst, Intrinsics::FromMatrix(K.data()), params, &p, stats);
  std::memcpy(transform->matrix().data(), p.m.data(), sizeof(float) * 16);
  return ok;
}
#endif

}  // namespace rs_tracker
