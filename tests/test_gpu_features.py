"""GPU tests of the Open3D steps around the registration path: FPFH descriptors and point-to-
point ICP on the device against the CPU checker (oracle/m3d_oracle_features.cpp; both restate Open3D v0.15.1 -- parity
unpinned until tools/pin_open3d.py runs against a real Open3D), and the whole chain FPFH -> match_correspondence ->
compute_transformation_ransac -> ICP on a synthetic pair with a known transform."""
import numpy as np
import pytest

from misc3d_b200 import synth

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n,radius,max_nn", [(4000, 0.12, 100), (1500, 0.3, 30), (300, 0.05, 16)])
def test_fpfh_matches_the_cpu_checker(ctx, capi, orc, n, radius, max_nn):
    d = synth.make_surface_pair(n=n, seed=7)
    f, ms = ctx.compute_fpfh(d["src"], d["src_nrm"], radius, max_nn)
    g = orc.fpfh(d["src"], d["src_nrm"], radius, max_nn)
    assert f.shape == g.shape == (33, n)
    close = np.isclose(f, g, rtol=1e-9, atol=1e-9)
    assert close.mean() > 0.999, close.mean()      # CUDA's atan2 / acos vs libm's: isolated bin-edge flips only
    has_nb = g.sum(0) > 0
    sums = f.T.reshape(n, 3, 11).sum(2)
    assert np.allclose(sums[has_nb], 200.0, atol=1e-8) and np.all(sums[~has_nb] == 0)


def test_fpfh_errors_and_degenerate_inputs(ctx, capi):
    d = synth.make_surface_pair(n=500, seed=1)
    with pytest.raises(capi.M3DError) as e:
        ctx.compute_fpfh(d["src"], None, 0.1, 50)
    assert e.value.code == capi.ERR_NO_NORMALS
    with pytest.raises(capi.M3DError):
        ctx.compute_fpfh(d["src"], d["src_nrm"], 0.1, 1000)
    f, _ = ctx.compute_fpfh(d["src"][:1], d["src_nrm"][:1], 0.1, 50)      # a single point has no neighbour: zeros
    assert f.shape == (33, 1) and not f.any()
    dup = np.repeat(d["src"][:50], 3, axis=0)                            # exact duplicates (zero distances are skipped)
    f, _ = ctx.compute_fpfh(dup, np.repeat(d["src_nrm"][:50], 3, axis=0), 0.2, 20)
    assert np.all(np.isfinite(f))


def test_icp_matches_the_cpu_checker(ctx, capi, orc):
    d = synth.make_surface_pair(n=3000, seed=5, sigma=0.001)
    T0 = np.eye(4)
    T0[:3, 3] = [0.03, -0.02, 0.01]
    Ti = d["T_true"] @ T0
    for max_dist, iters in ((0.08, 30), (0.05, 3)):
        T, fit, rmse, it = ctx.icp_point_to_point(d["src"], d["dst"], max_dist, Ti, iters)
        oT, ofit, ormse, oit = orc.icp(d["src"], d["dst"], max_dist, Ti, iters)
        assert it == oit and abs(fit - ofit) < 1e-12 and abs(rmse - ormse) < 1e-9
        np.testing.assert_allclose(T, oT, rtol=0, atol=1e-9)
    assert np.linalg.norm(T - d["T_true"]) < 0.05
    # no overlap at all: the initial transform comes back, fitness 0
    T, fit, rmse, it = ctx.icp_point_to_point(d["src"], d["dst"] + 100.0, 0.05, None, 10)
    assert fit == 0 and np.array_equal(T, np.eye(4))


def test_feature_to_icp_chain_recovers_the_transform(ctx, capi):
    """what examples/cpp/transform_estimation.cpp does with Open3D + Misc3D, entirely through this library"""
    d = synth.make_surface_pair(n=20000, seed=11, sigma=0.0005)
    fs, _ = ctx.compute_fpfh(d["src"], d["src_nrm"], 0.15, 100)
    fd, _ = ctx.compute_fpfh(d["dst"], d["dst_nrm"], 0.15, 100)
    i0, i1, _ = ctx.match_correspondence(fs, fd)
    assert len(i0) > 2000
    truth = np.empty(len(d["perm"]), dtype=np.int64)
    truth[d["perm"]] = np.arange(len(d["perm"]))          # src index -> dst index
    assert np.mean(truth[i0.astype(np.int64)] == i1.astype(np.int64)) > 0.15   # enough true pairs for RANSAC
    rc, T, st = ctx.ransac_registration(d["src"], d["dst"], i0, i1, 0.02, 20000, 0.9, 0.999, 1)
    assert rc == 1 and np.linalg.norm(T - d["T_true"]) < 0.05
    T2, fit, rmse, it = ctx.icp_point_to_point(d["src"], d["dst"], 0.02, T, 30)
    assert fit > 0.99 and rmse < 0.002 and np.linalg.norm(T2 - d["T_true"]) < 2e-3


def test_device_resident_feature_chain(ctx, capi, orc):
    """f3: FPFH left on the device and matched there == the host round-trip chain, bit for bit"""
    d = synth.make_surface_pair(n=6000, seed=3)
    fa_host, _ = ctx.compute_fpfh(d["src"], d["src_nrm"], 0.1, 60)
    fb_host, _ = ctx.compute_fpfh(d["dst"], d["dst_nrm"], 0.1, 60)
    fa, ms_a = ctx.fpfh_features(d["src"], d["src_nrm"], 0.1, 60)
    fb, ms_b = ctx.fpfh_features(d["dst"], d["dst_nrm"], 0.1, 60)
    assert (fa.dim, fa.n) == (33, 6000) and ms_a > 0
    np.testing.assert_array_equal(fa.download(), fa_host)
    np.testing.assert_array_equal(fb.download(), fb_host)
    i0, i1, _ = ctx.match_correspondence(fa_host, fb_host)
    j0, j1, _ = ctx.match_features(fa, fb)
    np.testing.assert_array_equal(i0, j0)
    np.testing.assert_array_equal(i1, j1)
    # descriptors computed elsewhere: upload once, match on the device
    ua, ub = ctx.upload_features(fa_host), ctx.upload_features(fb_host)
    k0, k1, _ = ctx.match_features(ua, ub)
    np.testing.assert_array_equal(i0, k0)
    np.testing.assert_array_equal(i1, k1)
    with pytest.raises(capi.M3DError):
        ctx.match_features(ua, ctx.upload_features(np.zeros((5, 10))))   # dimensions differ
    for f in (fa, fb, ua, ub):
        f.free()
