"""GPU: the cloud utilities of rs_tracker/common (point_cloud_utils.hpp) for callers that hold clouds, against the
reference's own functions compiled unmodified (oracle/_ref) and, where _ref is absent, the bit-identical restatement:
FindCorrespondences indices + squared distances bit-exact, RemoveNans exact, DownsampleVoxel the same point set
(and exactly Oracle-R's first-occurrence order), ComputeCovariances within fp32 round-off."""
import numpy as np
import pytest

from conftest import ROOT
from oracle import oracle as O
from realsensetracker_b200 import Aligner

pytestmark = pytest.mark.gpu
GOLD = np.load(ROOT / "tests" / "golden" / "oracle_r.npz")
GN = np.load(ROOT / "tests" / "golden" / "oracle_n_160x120.npz")
HAVE_REF = O.ref_lib() is not None


@pytest.fixture(scope="module")
def al():
    a = Aligner(16, 16, 2, 1)
    yield a
    a.close()


def depth_cloud(i, voxel=0.05):
    c = O.remove_nans(O.backproject(GN["frames"][i], tuple(GN["intr"])))
    return O.downsample_voxel(c, voxel) if voxel > 0 else c


def test_find_correspondences_bit_exact(al):
    for tgt, src in ((GOLD["dst"], GOLD["src"]), (depth_cloud(0), depth_cloud(1)), (depth_cloud(0, 0.0), depth_cloud(1))):
        idx_g, d2_g = al.find_correspondences(tgt, src)
        idx_o, d2_o = (O.ref_find_correspondences if HAVE_REF else O.nn)(tgt, src)
        assert np.array_equal(d2_g, d2_o)                       # the squared distance is unique even where the index ties
        diff = idx_g != idx_o
        assert not diff.any() or np.array_equal(                # exact distance ties: any of the tying points is a nearest neighbour
            ((src[diff] - tgt[idx_g[diff]]) ** 2).sum(1).astype(np.float32), ((src[diff] - tgt[idx_o[diff]]) ** 2).sum(1).astype(np.float32))
        assert diff.mean() < 1e-3
    q = GOLD["src"].copy(); q[7] = np.nan
    idx, d2 = al.find_correspondences(GOLD["dst"], q)
    assert idx[7] == -1 and np.isinf(d2[7]) and (idx[:7] >= 0).all()


def test_remove_nans_and_downsample_voxel(al):
    src = GOLD["src"]
    bad = src.copy(); bad[3, 0] = np.nan; bad[10, 2] = np.inf; bad[599] = -np.inf
    want = O.ref_remove_nans(bad) if HAVE_REF else O.remove_nans(bad)
    assert np.array_equal(al.remove_nans(bad), want)
    assert al.remove_nans(np.full((4, 3), np.nan, np.float32)).shape == (0, 3)
    for cloud, voxel in ((src, 0.25), (depth_cloud(0, 0.0), 0.05), (depth_cloud(2, 0.0), 0.013)):
        got = al.downsample_voxel(cloud, voxel)
        assert np.array_equal(got, O.downsample_voxel(cloud, voxel))          # first point per voxel, first-occurrence order
        if HAVE_REF:
            ref = O.ref_downsample_voxel(cloud, voxel)
            key = lambda x: x[np.lexsort(x.T[::-1])]
            assert got.shape == ref.shape and np.array_equal(key(got), key(ref))   # the reference's order is unordered_map's


@pytest.mark.skipif(not HAVE_REF, reason="needs oracle/_ref/libref.so")
def test_compute_covariances(al):
    cloud = depth_cloud(1)
    for gicp in (False, True):
        got = al.cloud_covariances(cloud, use_gicp=gicp)
        ref = O.ref_covariances(cloud, gicp)
        assert got.shape == ref.shape == (len(cloud), 3, 3)
        assert np.allclose(got, np.swapaxes(got, 1, 2), atol=1e-6)
        scale = np.abs(ref).max(axis=(1, 2))
        err = np.abs(got - ref).max(axis=(1, 2)) / scale
        if not gicp:
            assert err.max() < 2e-4, err.max()                   # same neighbours, same fp32 sums up to FMA contraction
        else:
            # singular vectors: where the two smallest singular values nearly tie (edges, corners) the plane normal is
            # ill-conditioned and two correct SVDs differ; everywhere else the regularised covariance must agree
            assert np.median(err) < 1e-4 and (err < 1e-2).mean() > 0.97, (np.median(err), (err < 1e-2).mean())
            ev = np.linalg.eigvalsh(got.astype(np.float64))
            assert np.allclose(ev[:, 0], 1e-2, atol=1e-4) and np.allclose(ev[:, 1:], 1.0, atol=1e-4)
