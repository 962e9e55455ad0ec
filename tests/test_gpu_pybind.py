"""The drop-in surface: the pybind11 `misc3d` module (python/) with the reference's function names,
argument names and defaults, driven the way examples/python/*.py drive the reference."""
import os
import sys

import numpy as np
import pytest

from misc3d_b200 import synth

pytestmark = pytest.mark.gpu
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "python"))


class FakeO3DCloud:
    """duck-types open3d.geometry.PointCloud (.points / .normals convertible by np.asarray)"""

    def __init__(self, points, normals=None):
        self.points = points
        if normals is not None:
            self.normals = normals


@pytest.fixture(scope="module")
def m3d():
    import misc3d
    misc3d.set_verbosity_level(misc3d.VerbosityLevel.Error)
    return misc3d


def test_fit_functions_match_oracle(m3d, orc):
    xyz, nrm = synth.make_c2(n=30000, seed=3)
    pc = FakeO3DCloud(xyz, nrm)
    for name, kind in (("fit_plane", orc.PLANE), ("fit_sphere", orc.SPHERE), ("fit_cylinder", orc.CYLINDER)):
        w, idx = getattr(m3d.common, name)(pc, 0.01, 300, 0.9999, seed=5)
        rc, model, inl, st = orc.ransac_fit(kind, xyz, nrm if kind == orc.CYLINDER else None, thr=0.01, max_it=300,
                                            prob=0.9999, seed=5)
        assert isinstance(idx, list) and isinstance(w, np.ndarray) and w.dtype == np.float64
        np.testing.assert_array_equal(np.asarray(idx, dtype=np.uint64), inl)
        np.testing.assert_allclose(w, model, rtol=1e-9, atol=1e-12)
    # defaults (threshold=0.01, max_iteration=1000, probability=0.9999) and a plain ndarray input
    w, idx = m3d.common.fit_plane(xyz)
    assert w.shape == (4,) and len(idx) > 0.3 * len(xyz)


def test_errors_are_runtime_errors(m3d):
    xyz = synth.make_c1(n=100, seed=1)
    with pytest.raises(RuntimeError):
        m3d.common.fit_cylinder(xyz)              # no normals (py_common.cpp:50-52)
    with pytest.raises(RuntimeError):
        m3d.common.fit_plane(xyz[:2])             # lack of points (ransac.h:510-513)
    with pytest.raises(RuntimeError):
        m3d.common.fit_plane(xyz, probability=0)  # ransac.h:483-485
    w, idx = m3d.common.fit_cylinder(FakeO3DCloud(np.tile([[1.0, 2, 3]], (50, 1)), np.tile([[0, 0, 1.0]], (50, 1))))
    assert w.shape == (4,) and not w.any() and idx == []  # FitModel false -> setZero(4) (py_common.cpp:62)


def test_segment_plane_iterative(m3d, orc):
    xyz = synth.make_c3(n=30000, seed=4)
    res = m3d.segmentation.segment_plane_iterative(FakeO3DCloud(xyz), 0.01, 100, 0.05, seed=7)
    rc, planes, labels = orc.segment_plane_iterative(xyz, 0.01, 100, 0.05, seed=7)
    assert len(res) == len(planes)
    for k, (w, cluster) in enumerate(res):
        np.testing.assert_allclose(w, planes[k], rtol=0, atol=1e-12)
        np.testing.assert_array_equal(np.asarray(cluster), xyz[labels == k])


def test_registration_chain(m3d, orc):
    d = synth.make_c4(n=3000, seed=5)
    assert int(m3d.registration.MatchMethod.FLANN) == 0 and int(m3d.registration.MatchMethod.ANNOY) == 1
    corres = m3d.registration.match_correspondence(d["src_feat"], d["dst_feat"])
    o0, o1 = orc.match_correspondence(d["src_feat"], d["dst_feat"])
    assert isinstance(corres, tuple) and isinstance(corres[0], list)
    np.testing.assert_array_equal(corres[0], o0)
    np.testing.assert_array_equal(corres[1], o1)

    class Feature:  # duck-types open3d.pipelines.registration.Feature
        def __init__(self, a):
            self.data = a
    c2 = m3d.registration.match_correspondence(Feature(d["src_feat"]), Feature(d["dst_feat"]),
                                               m3d.registration.MatchMethod.FLANN)
    assert c2 == corres
    T = m3d.registration.compute_transformation_ransac(FakeO3DCloud(d["src"]), FakeO3DCloud(d["dst"]), corres, 0.02,
                                                       2000, 0.9, seed=1)
    rc, oT, st = orc.ransac_registration(d["src"], d["dst"], o0, o1, thr=0.02, max_iter=2000, edge_thr=0.9, seed=1)
    assert T.shape == (4, 4) and np.linalg.norm(T - oT) <= 1e-5
    Tl = m3d.registration.compute_transformation_least_square(d["src"][o0.astype(int)], d["dst"][o1.astype(int)])
    assert np.linalg.norm(Tl - orc.umeyama(d["src"][o0.astype(int)], d["dst"][o1.astype(int)])) <= 1e-9
    # extension: least-squares refit on the inlier correspondences of the RANSAC result
    Tr = m3d.registration.refine_transformation_on_inliers(FakeO3DCloud(d["src"]), FakeO3DCloud(d["dst"]), corres, T, 0.02)
    assert Tr.shape == (4, 4) and np.linalg.norm(Tr - d["T_true"]) <= np.linalg.norm(T - d["T_true"]) + 1e-12


def test_feature_extensions_host_and_device(m3d, orc):
    """extensions of the shim: FPFH on the GPU as an ndarray and as a DeviceFeature; match_correspondence takes either;
    registration_icp"""
    d = synth.make_surface_pair(n=5000, seed=6)
    src, dst = FakeO3DCloud(d["src"], d["src_nrm"]), FakeO3DCloud(d["dst"], d["dst_nrm"])
    fa = m3d.registration.compute_fpfh_feature(src, 0.1, 60)
    fb = m3d.registration.compute_fpfh_feature(dst, 0.1, 60)
    assert fa.shape == (33, 5000) and fa.flags.f_contiguous
    ofa = orc.fpfh(d["src"], d["src_nrm"], 0.1, 60)
    assert np.mean(np.abs(fa - ofa) > 1e-9) < 1e-3            # isolated histogram-edge flips only
    da = m3d.registration.compute_fpfh_feature_device(src, 0.1, 60)
    db = m3d.registration.compute_fpfh_feature_device(dst, 0.1, 60)
    assert (da.dimension(), da.num()) == (33, 5000)
    np.testing.assert_array_equal(da.data, fa)
    host = m3d.registration.match_correspondence(fa, fb)
    dev = m3d.registration.match_correspondence(da, db)
    assert host == dev and len(host[0]) > 200
    T = m3d.registration.compute_transformation_ransac(src, dst, dev, 0.02, 4000, 0.9, seed=1)
    T_icp, fitness, rmse, iterations = m3d.registration.registration_icp(src, dst, 0.02, T, 30)
    assert fitness > 0.9 and np.linalg.norm(T_icp - d["T_true"]) < 5e-3
