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


def test_compute_centroid(al):
    """ComputeCentroid (point_cloud_utils.cpp:92-98): the reference sums sequentially in fp32, the device in fp64 in a
    fixed order; both within fp32 round-off of the exact mean, and the device result repeats bit for bit."""
    for cloud in (GOLD["src"], depth_cloud(0), depth_cloud(2, 0.0), GOLD["dst"][:1]):
        got = al.cloud_centroid(cloud)
        exact = cloud.astype(np.float64).mean(axis=0)
        scale = np.abs(cloud).max()
        assert np.abs(got - exact).max() <= 1e-7 * scale + 1e-12
        assert np.array_equal(got, al.cloud_centroid(cloud))
        if HAVE_REF:
            assert np.abs(got - O.ref_centroid(cloud)).max() <= 1e-4 * scale      # the reference's own fp32 running sum
    with pytest.raises(Exception):
        al.cloud_centroid(np.zeros((0, 3), np.float32))


def test_orient_normals(al):
    """OrientNormals (point_cloud_utils.cpp:205-216) on caller-supplied normals: the flips of the compiled reference,
    except where the ray is perpendicular to the normal to within rounding (the sign of a sum of three products)."""
    rng = np.random.default_rng(5)
    cloud = depth_cloud(1)
    nrm = rng.normal(size=cloud.shape).astype(np.float32)
    nrm /= np.linalg.norm(nrm, axis=1, keepdims=True)
    nrm[3] = 0.0                                              # a zero normal is left alone (dot = 0 is not > 0)
    for vp in ((0.0, 0.0, 0.0), (0.3, -0.2, 1.5)):
        got = al.orient_normals(cloud, vp, nrm)
        dot = ((cloud.astype(np.float64) - np.asarray(vp)) * nrm).sum(1)
        want = np.where((dot > 0)[:, None], -nrm, nrm)
        clear = np.abs(dot) > 1e-5
        assert clear.mean() > 0.99
        assert np.array_equal(got[clear], want[clear])
        assert np.array_equal(np.abs(got), np.abs(nrm))      # only signs change
        assert (((cloud - np.asarray(vp, np.float32)) * got).sum(1)[clear] <= 0).all()
        if HAVE_REF:
            assert np.array_equal(got[clear], O.ref_orient_normals(cloud, vp, nrm)[clear])
    # ComputeNormals + OrientNormals in two calls == the fused rst_cloud_normals with that viewpoint
    vp = (0.3, -0.2, 1.5)
    n0 = al.cloud_normals(cloud, 16)
    n1 = al.cloud_normals(cloud, 16, viewpoint=vp)
    n2 = al.orient_normals(cloud, vp, n0)
    d = ((cloud.astype(np.float64) - np.asarray(vp)) * n1).sum(1)
    clear = np.abs(d) > 1e-5
    assert np.array_equal(n2[clear], n1[clear])


def test_compute_extents(al):
    """ComputeExtents (point_cloud_utils.cpp:26-32): bit-exact (min / max are exact), the empty box for an empty cloud."""
    for cloud in (GOLD["src"], depth_cloud(0), depth_cloud(2, 0.0), GOLD["dst"][:1]):
        lo, hi = al.cloud_extents(cloud)
        assert np.array_equal(lo, cloud.min(axis=0)) and np.array_equal(hi, cloud.max(axis=0))
        if HAVE_REF:
            rlo, rhi = O.ref_extents(cloud)
            assert np.array_equal(lo, rlo) and np.array_equal(hi, rhi)
    lo, hi = al.cloud_extents(np.zeros((0, 3), np.float32))
    fmax = np.finfo(np.float32).max
    assert (lo == fmax).all() and (hi == -fmax).all()
    bad = GOLD["src"].copy(); bad[11] = np.nan
    lo, hi = al.cloud_extents(bad)
    keep = np.delete(GOLD["src"], 11, axis=0)
    assert np.array_equal(lo, keep.min(axis=0)) and np.array_equal(hi, keep.max(axis=0))     # NaN coordinates are ignored


def brute_knn(target, queries, k):
    """(d2, index)-ordered k nearest neighbours with nanoflann's fp32 accumulation ((dx^2 + dy^2) + dz^2)."""
    t, q = target.astype(np.float32), queries.astype(np.float32)
    d = q[:, None, :] - t[None, :, :]
    d2 = (d[..., 0] * d[..., 0] + d[..., 1] * d[..., 1]) + d[..., 2] * d[..., 2]
    order = np.lexsort((np.broadcast_to(np.arange(len(t)), d2.shape), d2), axis=1)[:, :k]
    return order.astype(np.int32), np.take_along_axis(d2, order, axis=1)


def test_tree_handle_knn_queries(al):
    """KDTree3f built once, queried many times (kdtree.hpp:26-57): k = 1 is FindCorrespondences bit for bit; k > 1 is the
    brute-force (distance, index) order with bit-identical squared distances; short clouds and non-finite queries pad
    with -1 / +inf; the handle survives other calls on the context."""
    tgt, src = GOLD["dst"], GOLD["src"]
    tree = al.tree_create(tgt)
    try:
        idx1, d21 = al.tree_query(tree, src, 1)
        idx_f, d2_f = al.find_correspondences(tgt, src)
        assert np.array_equal(idx1[:, 0], idx_f) and np.array_equal(d21[:, 0], d2_f)
        al.cloud_normals(depth_cloud(0), 16)                     # re-lays the context's arenas: the tree has its own memory
        far = src[:64] * 3.0 + 5.0                               # queries well outside the cloud's box
        for q, k in ((src, 5), (src[:200], 16), (src[:50], 33), (far, 8), (tgt[:100], 2)):
            idx, d2 = al.tree_query(tree, q, k)
            bi, bd = brute_knn(tgt, q, k)
            assert np.array_equal(d2, bd)
            same = idx == bi
            assert same.mean() > 0.999 and np.array_equal(d2[~same], bd[~same])      # exact distance ties only
        idx, d2 = al.tree_query(tree, tgt[:100], 2)
        assert np.array_equal(idx[:, 0], np.arange(100)) and (d2[:, 0] == 0).all()   # a cloud point's nearest is itself
        q = src[:8].copy(); q[2, 1] = np.nan; q[5] = np.inf
        idx, d2 = al.tree_query(tree, q, 4)
        assert (idx[[2, 5]] == -1).all() and np.isinf(d2[[2, 5]]).all() and (idx[[0, 1, 3, 4, 6, 7]] >= 0).all()
    finally:
        al.tree_destroy(tree)
    small = al.tree_create(tgt[:3])
    try:
        idx, d2 = al.tree_query(small, src[:10], 5)
        bi, bd = brute_knn(tgt[:3], src[:10], 3)
        assert np.array_equal(idx[:, :3], bi) and np.array_equal(d2[:, :3], bd)
        assert (idx[:, 3:] == -1).all() and np.isinf(d2[:, 3:]).all()
        with pytest.raises(Exception):
            al.tree_query(small, src[:10], 34)
    finally:
        al.tree_destroy(small)
    with pytest.raises(Exception):
        al.tree_create(np.zeros((0, 3), np.float32))
