"""GPU parity tests of compute_transformation_ransac (RANSACSolver::Solve -> Open3D
RegistrationRANSACBasedOnCorrespondence, restated in the oracle) and of the least-squares
(Umeyama) solver.  north_star tolerance: 4x4 transforms within 1e-5 Frobenius; the loop
statistics (best index / count / stop index) are integer work and must be identical."""
import os

import numpy as np
import pytest

from misc3d_b200 import synth

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")
FROB_TOL = 1e-5


def _check(ctx, orc, src, dst, c0, c1, thr, max_iter, edge, conf, seed):
    rc, T, st = ctx.ransac_registration(src, dst, c0, c1, thr, max_iter, edge, conf, seed)
    orc_rc, oT, ost = orc.ransac_registration(src, dst, c0, c1, thr=thr, max_iter=max_iter, edge_thr=edge,
                                              confidence=conf, seed=seed)
    assert rc == orc_rc
    for k in ("best_index", "best_count", "evaluated", "stop_index"):
        assert st[k] == ost[k], (k, st, ost)
    assert np.linalg.norm(T - oT) <= FROB_TOL
    np.testing.assert_array_equal(T, oT)  # in fact the 3-point Umeyama is replayed bit for bit
    assert abs(st["best_rmse"] - ost["best_rmse"]) <= 1e-12 * max(1.0, ost["best_rmse"])
    return T, st


def test_golden_reg_small(ctx, capi, orc):
    g = np.load(os.path.join(GOLD, "reg_small.npz"))
    d = synth.make_c4(n=3000, seed=5)
    rc, T, st = ctx.ransac_registration(d["src"], d["dst"], g["i0"], g["i1"], 0.02, 2000, 0.9, 0.999, 1)
    assert np.linalg.norm(T - g["T"]) <= FROB_TOL
    assert st["best_index"] == int(g["best_index"]) and st["best_count"] == int(g["best_count"])
    assert st["stop_index"] == int(g["stop_index"]) and st["evaluated"] == int(g["evaluated"])
    assert np.linalg.norm(T - d["T_true"]) < 0.05


@pytest.mark.parametrize("conf", [0.999, 1.0])
@pytest.mark.parametrize("n,seed", [(3000, 2), (20000, 3)])
def test_registration_parity(ctx, capi, orc, n, seed, conf):
    d = synth.make_c4(n=n, seed=seed)
    i0, i1 = orc.match_correspondence(d["src_feat"], d["dst_feat"])
    _check(ctx, orc, d["src"], d["dst"], i0, i1, 0.02, 3000, 0.9, conf, seed)


def test_registration_with_many_outlier_correspondences(ctx, capi, orc):
    """70 % random correspondences: most triples fail the checkers, est_k stays large"""
    rng = np.random.default_rng(1)
    d = synth.make_c4(n=4000, seed=7)
    m = 3000
    c0 = rng.integers(0, 4000, m)
    inv = np.empty(4000, dtype=np.int64)
    inv[d["perm"]] = np.arange(4000)
    c1 = inv[c0].copy()          # true partner of src[c0] in dst
    bad = rng.uniform(size=m) < 0.7
    c1[bad] = rng.integers(0, 4000, int(bad.sum()))
    T, st = _check(ctx, orc, d["src"], d["dst"], c0, c1, 0.02, 20000, 0.9, 0.999, 5)
    assert np.linalg.norm(T - d["T_true"]) < 0.05
    _check(ctx, orc, d["src"] * 1000 + 5e4, d["dst"] * 1000 - 2e4, c0, c1, 20.0, 5000, 0.9, 1.0, 6)


def test_registration_edge_cases(ctx, capi, orc):
    d = synth.make_c4(n=200, seed=1)
    c = np.arange(200)
    with pytest.raises(capi.M3DError) as e:  # transform_estimation.cpp:130-133 throws
        ctx.ransac_registration(d["src"][:2], d["dst"], c[:2], c[:2])
    assert e.value.code == capi.ERR_TOO_FEW_POINTS
    rc, T, st = ctx.ransac_registration(d["src"], d["dst"], c[:2], c[:2])  # < 3 correspondences: default result
    assert rc == 0 and np.array_equal(T, np.eye(4))
    # no hypothesis survives the checkers -> identity, like Open3D's default result
    rng = np.random.default_rng(0)
    rc, T, st = ctx.ransac_registration(d["src"], rng.uniform(-1, 1, (200, 3)), c, c, 1e-4, 500, 0.9, 0.999, 3)
    orc_rc, oT, ost = orc.ransac_registration(d["src"], rng.uniform(-1, 1, (200, 3)), c, c, thr=1e-4, max_iter=500)
    assert np.array_equal(T, np.eye(4)) and st["best_count"] == 0


@pytest.mark.parametrize("scaling", [False, True])
def test_least_squares_transform(ctx, capi, orc, scaling):
    rng = np.random.default_rng(4)
    src = rng.normal(size=(50000, 3))
    R = synth.rotation_about((0.2, 0.9, -0.4), 63.0)
    dst = (1.7 if scaling else 1.0) * src @ R.T + np.array([0.3, -1.2, 4.0]) + rng.normal(0, 1e-3, size=src.shape)
    T = ctx.least_squares_transform(src, dst, scaling)
    oT = orc.umeyama(src, dst, scaling)
    assert np.linalg.norm(T - oT) <= 1e-9


def test_c4_sized_registration_properties(ctx, capi, orc):
    """BASELINE config C4 size: 200k points, mutual matches, 50k hypotheses, no early exit; and the oracle's loop on
    the same ~94k correspondences with 2000 hypotheses (statistics exact, transform bit for bit)"""
    d = synth.make_c4()
    i0, i1, ms = ctx.match_correspondence(d["src_feat"], d["dst_feat"])
    rc, T, st = ctx.ransac_registration(d["src"], d["dst"], i0, i1, 0.02, 50000, 0.9, 1.0, 1)
    assert rc == 1 and st["stop_index"] == 50000
    assert np.linalg.norm(T - d["T_true"]) < 0.02
    p = d["src"][i0.astype(np.int64)] @ T[:3, :3].T + T[:3, 3]
    good = int((((p - d["dst"][i1.astype(np.int64)]) ** 2).sum(1) < 0.02 ** 2).sum())
    assert abs(good - st["best_count"]) <= 2  # numpy evaluates T*p in a different operation order
    for conf in (1.0, 0.999):
        _check(ctx, orc, d["src"], d["dst"], i0, i1, 0.02, 2000, 0.9, conf, 3)


def test_open3d_pin(ctx, capi):
    """the GPU path against a real Open3D (tests/golden/reg_open3d.npz from tools/pin_open3d.py; skipped while absent)"""
    path = os.path.join(GOLD, "reg_open3d.npz")
    if not os.path.exists(path):
        pytest.skip("tests/golden/reg_open3d.npz absent: run tools/pin_open3d.py where open3d is installed")
    g = np.load(path)
    d = synth.make_c4(n=3000, seed=5)
    i0, i1 = g["i0"], g["i1"]
    np.testing.assert_allclose(ctx.least_squares_transform(d["src"][i0], d["dst"][i1], False), g["T_ls"], rtol=0, atol=1e-9)
    np.testing.assert_allclose(ctx.least_squares_transform(d["src"][i0], d["dst"][i1], True), g["T_ls_scale"], rtol=0, atol=1e-9)
    rc, T, st = ctx.ransac_registration(d["src"], d["dst"], i0, i1, float(g["thr"]), int(g["max_iter"]), float(g["edge"]),
                                        0.999, 1)
    assert abs(st["best_count"] / len(i0) - float(g["fitness"])) < 0.02 and np.linalg.norm(T - g["T"]) < 0.05


def test_gpu_agrees_with_the_independent_numpy_restatement(ctx, capi, orc):
    """a19 without a compiled reference: the GPU loop against tests/reg_numpy_ref.py (numpy-only, np.linalg.svd) on the
    recorded sample table -- statistics exactly, transform within 1e-9 (and far inside the 1e-5 Frobenius tolerance)"""
    import reg_numpy_ref as ref
    d = synth.make_c4(n=2500, seed=4)
    i0, i1 = orc.match_correspondence(d["src_feat"], d["dst_feat"])
    picks = orc.reg_sample_table(6, len(i0), 1500)
    T, st = ref.ransac_registration_np(d["src"], d["dst"], i0.astype(np.int64), i1.astype(np.int64), picks, 0.02, 1500, 0.9, 0.999)
    rc, gT, gst = ctx.ransac_registration(d["src"], d["dst"], i0, i1, 0.02, 1500, 0.9, 0.999, 6)
    assert rc == 1
    for k in ("best_index", "best_count", "evaluated", "stop_index"):
        assert gst[k] == st[k], (k, gst, st)
    np.testing.assert_allclose(gT, T, rtol=0, atol=1e-9)


@pytest.mark.parametrize("scaling", [False, True])
def test_refit_on_inlier_correspondences(ctx, capi, orc, scaling):
    """extension f2: Umeyama over the correspondences that are inliers of T (correspondence order), against numpy"""
    d = synth.make_c4(n=20000, seed=5)
    i0, i1, ms = ctx.match_correspondence(d["src_feat"], d["dst_feat"])
    rc, T, st = ctx.ransac_registration(d["src"], d["dst"], i0, i1, 0.02, 2000, 0.9, 1.0, 1)
    assert rc == 1
    T2, n_inl = ctx.registration_refit(d["src"], d["dst"], i0, i1, T, 0.02, scaling)
    s = d["src"][i0.astype(np.int64)]
    q = d["dst"][i1.astype(np.int64)]
    p = s @ T[:3, :3].T + T[:3, 3]
    d2 = ((p - q) ** 2).sum(1)
    inl = d2 < 0.02 ** 2
    near = np.abs(d2 - 0.02 ** 2) < 1e-12          # pairs a different operation order could flip
    assert abs(int(inl.sum()) - n_inl) <= int(near.sum())
    if not near.any():
        oT = orc.umeyama(s[inl], q[inl], scaling)
        assert np.linalg.norm(T2 - oT) <= 1e-9
    # closer to the truth than the 3-point RANSAC estimate, and all of it a rigid (or similarity) transform
    assert np.linalg.norm(T2 - d["T_true"]) <= np.linalg.norm(T - d["T_true"]) + 1e-12
    # fewer than three inliers: the input comes back
    T3, n3 = ctx.registration_refit(d["src"], d["dst"], i0, i1, np.eye(4), 1e-9, scaling)
    assert n3 < 3 and np.array_equal(T3, np.eye(4))
    T4, n4 = ctx.registration_refit(d["src"], d["dst"], i0[:0], i1[:0], T, 0.02, scaling)
    assert n4 == 0 and np.array_equal(T4, T)
