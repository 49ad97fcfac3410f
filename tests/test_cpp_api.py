"""The C++ host API (include/rs_tracker/align/align_rgbd.hpp) over the C ABI: compiles with g++
(CPU check) and, on the GPU box, the rs_replay_app-style odometry loop runs through it."""
import subprocess

import pytest

from conftest import ROOT

EXE = ROOT / "examples" / "replay_synth"


def build_example():
    cmd = ["/usr/bin/g++", "-std=c++17", "-O2", "-Wall", "-I", str(ROOT / "include"), "-I", str(ROOT / "realsensetracker_b200" / "csrc"),
           str(ROOT / "examples" / "replay_synth.cpp"), "-L", str(ROOT / "realsensetracker_b200" / "_lib"),
           "-lrst_align", "-lrst_synth", "-Wl,-rpath," + str(ROOT / "realsensetracker_b200" / "_lib"), "-o", str(EXE)]
    res = subprocess.run(cmd, capture_output=True, text=True)
    assert res.returncode == 0, res.stderr
    return EXE


def test_cpp_header_compiles_and_links():
    build_example()


@pytest.mark.gpu
def test_replay_loop_through_the_cpp_api():
    exe = build_example()
    res = subprocess.run([str(exe), "12"], capture_output=True, text=True, timeout=300)
    assert res.returncode == 0, res.stdout + res.stderr
    assert "failures 0" in res.stdout


def test_eigen_overload_compiles_against_the_stand_in_eigen(tmp_path):
    """The Isometry3f / Matrix3f overload of AlignRgbd is only compiled when <Eigen/Geometry> exists; real
    Eigen is absent here, so the syntax is checked against the stand-in headers of oracle/shim (compile only)."""
    src = tmp_path / "eig.cpp"
    src.write_text('#include "rs_tracker/align/align_rgbd.hpp"\n'
                   '#ifndef RS_TRACKER_HAVE_EIGEN\n#error "Eigen overload not enabled"\n#endif\n'
                   'bool f(rs_tracker::AlignContext& c, const rs_tracker::DepthFrame& a, const rs_tracker::DepthFrame& b) {\n'
                   '  Eigen::Matrix3f K = Eigen::Matrix3f::Identity(); Eigen::Isometry3f T = Eigen::Isometry3f::Identity();\n'
                   '  rs_tracker::AlignParams p; return rs_tracker::AlignRgbd(c, a, b, K, p, &T); }\n')
    res = subprocess.run(["/usr/bin/g++", "-std=c++17", "-fsyntax-only", "-I", str(ROOT / "include"), "-I", str(ROOT / "oracle" / "shim"), str(src)],
                         capture_output=True, text=True)
    assert res.returncode == 0, res.stderr


LITERAL = ROOT / "oracle" / "_ref" / "literal_calls"


def test_reference_call_sites_compile_verbatim():
    """rs_replay_app.cpp:246-251, rs_align_app.cpp:295,303, align_icp.cpp:165 and point_cloud_utils.hpp:14 as written
    there (tests/cpp/literal_calls.cpp) compile against align_rgbd.hpp with the reference's own types.hpp supplying
    Cloud3f / KDTree3f. Needs /root/reference (build container only); the binary travels in oracle/_ref."""
    from pathlib import Path
    if not Path("/root/reference/rs_tracker/common/include/rs_tracker/common/types.hpp").exists():
        pytest.skip("the reference tree is not mounted here")
    res = subprocess.run(["make", "-C", str(ROOT / "oracle"), "_ref/literal_calls"], capture_output=True, text=True)
    assert res.returncode == 0, res.stdout + res.stderr
    assert LITERAL.exists()


@pytest.mark.gpu
def test_reference_call_sites_run_on_the_gpu():
    if not LITERAL.exists():
        pytest.skip("oracle/_ref/literal_calls was not built (needs the reference tree at build time)")
    res = subprocess.run([str(LITERAL)], capture_output=True, text=True, timeout=300)
    assert res.returncode == 0 and "failures 0" in res.stdout, res.stdout + res.stderr


def test_every_literal_signature_instantiates_against_the_stand_in_types(tmp_path):
    """Every context-free signature of align_rgbd.hpp (align_icp.hpp:14-24, point_cloud_utils.hpp:9-31 + ComputeExtents,
    align_gicp.hpp:15-25) and GpuKDTree3f, instantiated with the stand-in Cloud3f of oracle/shim and a tree type of the
    reference's shape (kdtree.hpp: public `m_cloud` reference wrapper). Compile + link only: runs where neither the
    reference tree nor a GPU exists; tests/cpp/literal_calls.cpp is the version against the reference's own types.hpp."""
    src = tmp_path / "lit.cpp"
    src.write_text(r'''
#include <functional>
#include <memory>
#include <cho_util/core/geometry/point_cloud.hpp>
#include "rs_tracker/align/align_rgbd.hpp"
#ifndef RS_TRACKER_HAVE_EIGEN
#error "Eigen signatures not enabled"
#endif
using Cloud3f = cho::core::PointCloud<float, 3>;
struct Tree3f { explicit Tree3f(const std::reference_wrapper<const Cloud3f>& c, int = 10) : m_cloud(c) {} const std::reference_wrapper<const Cloud3f> m_cloud; };
float all(const Cloud3f& src, const Cloud3f& dst) {
  Eigen::Isometry3f xfm = Eigen::Isometry3f::Identity(), seed = xfm;
  const Tree3f tree{std::cref(dst), 16};
  std::vector<std::pair<int, int>> pairs; std::vector<float> weights, d2; std::vector<int> idx;
  std::vector<Eigen::Matrix3f> sc, dc;
  Cloud3f out, normals; Eigen::Vector3f centroid; Eigen::AlignedBox3f box;
  bool ok = rs_tracker::SolveKabsch(src, dst, pairs, weights, &xfm);
  ok = rs_tracker::AlignIcp3d(src, dst, 128, &xfm) && ok;
  ok = rs_tracker::AlignIcp3d(src, dst, tree, 128, &xfm) && ok;
  rs_tracker::DownsampleVoxel(src, 0.05f, &out);
  rs_tracker::RemoveNans(src, &out);
  rs_tracker::FindCorrespondences(tree, src, &idx, &d2);
  rs_tracker::ComputeCentroid(src, &centroid);
  rs_tracker::ComputeExtents(src, &box);
  rs_tracker::ComputeCovariances(tree, dst, &dc, false);
  rs_tracker::ComputeCovariances(tree, src, &sc, true);
  rs_tracker::ComputeNormals(dst, tree, 16, &normals);
  rs_tracker::OrientNormals(dst, Eigen::Vector3f(0.f, 0.f, 1.f), &normals);
  float cost = rs_tracker::ComputeAlignment(src, dst, sc, dc, idx, seed, &xfm);
  cost += rs_tracker::ComputeAlignment(src, dst, &xfm);
  rs_tracker::GpuKDTree3f gpu_tree(rs_tracker::DefaultAlignContext(), dst);
  int j; float d; gpu_tree.query(src.GetPtr(), 1, &j, &d);
  gpu_tree.query(src, 4, &idx, &d2);
  return ok ? cost : -1.f;
}
int main() { return 0; }
''')
    exe = tmp_path / "lit"
    res = subprocess.run(["/usr/bin/g++", "-std=c++17", "-O1", "-Wall", "-I", str(ROOT / "include"), "-I", str(ROOT / "oracle" / "shim"), str(src),
                          "-L", str(ROOT / "realsensetracker_b200" / "_lib"), "-lrst_align",
                          "-Wl,-rpath," + str(ROOT / "realsensetracker_b200" / "_lib"), "-o", str(exe)], capture_output=True, text=True)
    assert res.returncode == 0, res.stderr
