// replay_synth.cpp — the reference's odometry loop (rs_tracker/app/src/rs_replay_app.cpp:211-287)
// on a synthetic depth sequence, through the C++ host API over the C ABI.
//
//   for each frame: AlignRgbd(curr, prev, K, params, &xfm = I); total = total * xfm
//
// The recorded `.pb` clouds the reference replays are not in its tree, so frames come from the
// seeded synthetic renderer (csrc/rst_synth.h). Prints the per-frame and accumulated pose error
// against the known trajectory; exit code 0 iff every alignment succeeded and the final drift is small.
//
// build:  g++ -std=c++17 -O2 -I include -I realsensetracker_b200/csrc examples/replay_synth.cpp \
//             -L realsensetracker_b200/_lib -lrst_align -lrst_synth -Wl,-rpath,'$ORIGIN/../realsensetracker_b200/_lib' -o examples/replay_synth
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <unordered_map>
#include <array>
#include <vector>

#include "rs_tracker/align/align_rgbd.hpp"
#include "rst_synth.h"

namespace {
struct Mat4d { double m[16]; };  // column-major

Mat4d Mul(const Mat4d& a, const Mat4d& b) {
  Mat4d r{};
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j < 4; ++j) {
      double s = 0;
      for (int k = 0; k < 4; ++k) s += a.m[i + 4 * k] * b.m[k + 4 * j];
      r.m[i + 4 * j] = s;
    }
  return r;
}
Mat4d InvRigid(const Mat4d& a) {
  Mat4d r{};
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) r.m[i + 4 * j] = a.m[j + 4 * i];
  for (int i = 0; i < 3; ++i) r.m[i + 12] = -(r.m[i] * a.m[12] + r.m[i + 4] * a.m[13] + r.m[i + 8] * a.m[14]);
  r.m[15] = 1;
  return r;
}
Mat4d CameraPose(int k) {  // smooth trajectory: ~1.5 cm, ~0.7 deg per frame
  const double s = 0.05 * k;
  const double rx = 0.06 * std::sin(0.9 * s), ry = 0.12 * std::sin(1.1 * s + 0.4), rz = 0.07 * std::sin(0.7 * s + 1.0);
  const double cx = std::cos(rx), sx = std::sin(rx), cy = std::cos(ry), sy = std::sin(ry), cz = std::cos(rz), sz = std::sin(rz);
  Mat4d T{};
  // R = Rx * Ry * Rz
  T.m[0] = cy * cz;                 T.m[4] = -cy * sz;                T.m[8] = sy;
  T.m[1] = sx * sy * cz + cx * sz;  T.m[5] = -sx * sy * sz + cx * cz; T.m[9] = -sx * cy;
  T.m[2] = -cx * sy * cz + sx * sz; T.m[6] = cx * sy * sz + sx * cz;  T.m[10] = cx * cy;
  T.m[12] = 0.3 * std::sin(s); T.m[13] = 0.12 * std::sin(1.3 * s + 0.5); T.m[14] = 0.2 * std::sin(0.8 * s + 0.2);
  T.m[15] = 1;
  return T;
}
void PoseError(const Mat4d& est, const Mat4d& gt, double* t_err, double* r_err) {
  const Mat4d d = Mul(InvRigid(gt), est);
  *t_err = std::sqrt(d.m[12] * d.m[12] + d.m[13] * d.m[13] + d.m[14] * d.m[14]);
  double c = (d.m[0] + d.m[5] + d.m[10] - 1.0) / 2.0;
  c = c > 1 ? 1 : (c < -1 ? -1 : c);
  *r_err = std::acos(c);
}
}  // namespace

// The replay app's voxel map (CloudAccumulator, rs_replay_app.cpp:76-129): the first point that lands in a
// voxel stays; the voxel index TRUNCATES toward zero (`(p * 1/voxel).cast<int>()`, :109-111 — unlike
// DownsampleVoxel, which floors).
class CloudAccumulator {
 public:
  explicit CloudAccumulator(float voxel_size = 0.05f) : inv_(static_cast<float>(1.0 / voxel_size)) {}
  void AddDepth(const Mat4d& xfm, const std::uint16_t* depth, int w, int h, const rs_tracker::Intrinsics& K, int step) {
    for (int v = 0; v < h; v += step)
      for (int u = 0; u < w; u += step) {
        const std::uint16_t d = depth[v * w + u];
        if (d == 0) continue;
        const float z = d * 0.001f, x = (u - K.cx) * z / K.fx, y = (v - K.cy) * z / K.fy;
        const float p[3] = {static_cast<float>(xfm.m[0] * x + xfm.m[4] * y + xfm.m[8] * z + xfm.m[12]),
                            static_cast<float>(xfm.m[1] * x + xfm.m[5] * y + xfm.m[9] * z + xfm.m[13]),
                            static_cast<float>(xfm.m[2] * x + xfm.m[6] * y + xfm.m[10] * z + xfm.m[14])};
        const std::uint64_t key = Pack(static_cast<int>(p[0] * inv_), static_cast<int>(p[1] * inv_), static_cast<int>(p[2] * inv_));
        map_.emplace(key, std::array<float, 3>{p[0], p[1], p[2]});   // emplace keeps the first point (:101-104)
      }
  }
  std::size_t Size() const { return map_.size(); }

 private:
  static std::uint64_t Pack(int x, int y, int z) {
    return (static_cast<std::uint64_t>(x & 0x1FFFFF) << 42) | (static_cast<std::uint64_t>(y & 0x1FFFFF) << 21) | static_cast<std::uint64_t>(z & 0x1FFFFF);
  }
  float inv_;
  std::unordered_map<std::uint64_t, std::array<float, 3>> map_;
};

// minimal stand-in for the reference's Cloud3f accessors (cho::core::PointCloud<float,3>)
struct VecCloud {
  std::vector<float> xyz;
  const float* GetPtr() const { return xyz.data(); }
  int GetNumPoints() const { return static_cast<int>(xyz.size() / 3); }
};

// the pairwise path of rs_align_app.cpp:295-303 on clouds: SolveKabsch from known matches, then AlignIcp3d
static bool CloudSelfTest(rs_tracker::AlignContext& ctx) {
  VecCloud src, dst;
  unsigned s = 12345u;
  auto rnd = [&]() { s = s * 1664525u + 1013904223u; return (float)((s >> 8) & 0xFFFF) / 32768.0f - 1.0f; };
  const float c = std::cos(0.05f), sn = std::sin(0.05f);
  std::vector<std::pair<int, int>> idx;
  for (int i = 0; i < 500; ++i) {
    const float x = rnd(), y = rnd(), z = rnd();
    src.xyz.insert(src.xyz.end(), {x, y, z});
    dst.xyz.insert(dst.xyz.end(), {c * x - sn * y + 0.05f, sn * x + c * y - 0.03f, z + 0.02f});  // R_z(0.05), t
    idx.emplace_back(i, i);
  }
  rs_tracker::Pose xfm;
  if (!rs_tracker::SolveKabsch(ctx, src, dst, idx, {}, &xfm)) return false;
  float mean_cost = 0.f;
  if (!rs_tracker::AlignIcp3d(ctx, src, dst, 32, &xfm, &mean_cost)) return false;
  std::printf("cloud self-test: R01 %.5f (want %.5f), t = %.4f %.4f %.4f, mean cost %.2e\n", xfm(0, 1), -sn, xfm(0, 3), xfm(1, 3), xfm(2, 3), mean_cost);
  return std::fabs(xfm(0, 1) + sn) < 1e-4f && std::fabs(xfm(0, 3) - 0.05f) < 1e-4f && std::fabs(xfm(2, 3) - 0.02f) < 1e-4f;
}

int main(int argc, char** argv) {
  const int n_frames = argc > 1 ? std::atoi(argv[1]) : 20;
  const int w = 640, h = 480;
  const rs_tracker::Intrinsics K{385.f, 385.f, 320.f, 240.f};
  rst_synth_scene scene;
  rst_synth_scene_default(0, &scene);

  rs_tracker::AlignContext ctx(/*device=*/0, w, h, /*max_frames=*/2, /*max_pairs=*/1);
  rs_tracker::AlignParams params;
  if (!CloudSelfTest(ctx)) { std::printf("cloud self-test FAILED\n"); return 2; }

  std::vector<std::uint16_t> prev(w * h), curr(w * h);
  Mat4d total{};  // accumulated estimate: camera k -> camera 0
  total.m[0] = total.m[5] = total.m[10] = total.m[15] = 1;
  const Mat4d T0 = CameraPose(0);
  rst_synth_render(&scene, T0.m, K.fx, K.fy, K.cx, K.cy, w, h, 0.001, nullptr, 0, prev.data(), w, nullptr);
  CloudAccumulator acc(0.05f);
  acc.AddDepth(total, prev.data(), w, h, K, 4);                      // acc.AddCloud(total_xfm, prev_cloud)  :240
  const std::size_t map_first = acc.Size();
  int failures = 0;
  double t_err = 0, r_err = 0;
  for (int k = 1; k < n_frames; ++k) {
    const Mat4d Tk = CameraPose(k);
    rst_synth_render(&scene, Tk.m, K.fx, K.fy, K.cx, K.cy, w, h, 0.001, nullptr, k, curr.data(), w, nullptr);
    rs_tracker::Pose xfm = rs_tracker::Pose::Identity();            // rs_replay_app.cpp:235
    rs_tracker::AlignStats st;
    const rs_tracker::DepthFrame fc{curr.data(), nullptr, w, h, 0, 0}, fp{prev.data(), nullptr, w, h, 0, 0};
    const bool suc = rs_tracker::AlignRgbd(ctx, fc, fp, K, params, &xfm, &st);  // AlignIcp3d(curr, prev, ...) :251
    if (suc) {
      Mat4d x{};
      for (int i = 0; i < 16; ++i) x.m[i] = xfm.m[i];
      total = Mul(total, x);                                         // total_xfm = total_xfm * xfm  :267
      acc.AddDepth(total, curr.data(), w, h, K, 4);                  // acc.AddCloud(total_xfm, cloud) :268
      prev.swap(curr);                                               // prev_cloud = cloud           :270
    } else {
      std::printf("ALIGNMENT FAILED!! frame %d status %d (%s)\n", k, st.status, ctx.LastError());  // :272
      ++failures;
    }
    PoseError(total, Mul(InvRigid(T0), Tk), &t_err, &r_err);
    std::printf("frame %3d  count %6d  rmse %.5f m  drift %.4f m %.4f rad\n", k, st.count, st.rmse, t_err, r_err);
  }
  std::printf("launches %lld, final drift %.4f m / %.4f rad over %d frames, failures %d\n",
              static_cast<long long>(rst_launch_count(ctx.get())), t_err, r_err, n_frames, failures);
  // with correct poses consecutive frames re-observe the same surfaces: the map grows far slower than linearly
  std::printf("voxel map: %zu voxels after the first frame, %zu after %d frames\n", map_first, acc.Size(), n_frames);
  const bool map_ok = acc.Size() < map_first * 2 || n_frames > 40;
  return (failures == 0 && t_err < 0.02 && r_err < 0.02 && map_ok) ? 0 : 1;
}
