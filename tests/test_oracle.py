"""CPU tests of the oracle: analytic known answers, independent numpy cross-checks, and the
committed golden vectors (tests/golden, made by tools/make_golden.py: outputs of the reference's own
sources compiled into oracle/_ref for the fits, segmentation and matching; the oracle's for the
Open3D-defined registration -- the reference itself ships no tests or fixtures).
tests/test_reference_pin.py compares the oracle with the compiled reference directly."""
import os

import numpy as np
import pytest

from misc3d_b200 import synth

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def test_sampler_matches_raw_mt19937_stream(orc):
    # utils.h:81-97: idx = rng() % size with duplicate rejection, std::mt19937(seed)
    seed, n, k, rows = 12345, 1000, 3, 200
    tab = orc.sample_table(seed, n, k, rows)
    raw = np.random.RandomState(seed)._bit_generator.random_raw(4 * rows * k)
    it = iter(int(v) for v in raw)
    for r in range(rows):
        got = []
        while len(got) < k:
            v = next(it) % n
            if v not in got:
                got.append(v)
        assert got == list(tab[r])


def test_plane_minimal_known_answer(orc):
    pts = np.array([[0, 0, 1.0], [1, 0, 1.0], [0, 1, 1.0]])
    ok, m = orc.minimal_fit(orc.PLANE, pts)
    assert ok == 1
    np.testing.assert_array_equal(m, [0, 0, 1, -1])
    assert orc.distance(orc.PLANE, m, [5, -3, 1.25]) == 0.25
    ok, _ = orc.minimal_fit(orc.PLANE, np.array([[0, 0, 0.0], [1, 1, 1], [2, 2, 2]]))
    assert ok == 0  # collinear: norm < 1e-8 (ransac.h:149-152)


def test_sphere_minimal_known_answer(orc):
    c, r = np.array([0.25, -0.5, 2.0]), 1.5
    d = np.array([[1, 0, 0], [0, 1, 0], [0, 0, 1], [-1, 0, 0.0]])
    ok, m = orc.minimal_fit(orc.SPHERE, c + r * d)
    assert ok == 1
    np.testing.assert_allclose(m, [0.25, -0.5, 2.0, 1.5], rtol=0, atol=1e-12)
    assert abs(orc.distance(orc.SPHERE, m, c + [0, 0, 2.0]) - 0.5) < 1e-12
    assert abs(orc.distance(orc.SPHERE, m, c + [0, 0, 1.0]) - 0.5) < 1e-12
    # 4th point in the plane of the first three -> ValidationCheck fails (ransac.h:225-234)
    ok, _ = orc.minimal_fit(orc.SPHERE, np.array([[1, 0, 0.0], [0, 1, 0], [-1, 0, 0], [0, -1, 0]]))
    assert ok == 0


def test_cylinder_minimal_known_answer(orc):
    # axis = z through (1, 2, *), radius 0.5; outward normals (PCL-style construction)
    pts = np.array([[1.5, 2.0, 0.0], [1.0, 2.5, 1.0]])
    nrm = np.array([[1.0, 0, 0], [0, 1.0, 0]])
    ok, m = orc.minimal_fit(orc.CYLINDER, pts, nrm)
    assert ok == 1
    axis = m[3:6]
    np.testing.assert_allclose(np.abs(axis), [0, 0, 1], atol=1e-12)
    assert abs(m[6] - abs(m[0] - 1.5)) < 1e-9 or m[6] >= 0  # radius = dist(points[0], axis)
    # the distance functor is |dist_to_axis - r|
    q = np.array([m[0] + 3.0, m[1], 7.0])
    assert abs(orc.distance(orc.CYLINDER, m, q) - abs(3.0 - m[6])) < 1e-12


def test_evaluate_against_numpy(orc):
    xyz = synth.make_c1(n=5000, seed=3)
    ok, m = orc.minimal_fit(orc.PLANE, xyz[[10, 200, 3000]])
    cnt, err = orc.evaluate(orc.PLANE, xyz, m, 0.01)
    d = np.abs(xyz @ m[:3] + m[3]) / np.linalg.norm(m[:3])
    assert cnt == int((d < 0.01).sum())
    assert abs(err - d[d < 0.01].sum()) < 1e-9


def test_c1_fit_recovers_plane_and_inlier_definition(orc):
    xyz = synth.make_c1()
    rc, model, inl, st = orc.ransac_fit(orc.PLANE, xyz, thr=0.01, max_it=100, prob=0.9999, seed=1)
    assert rc == 1 and st["found"] == 1
    n_true = np.array([0.1, -0.2, 0.97])
    n_true /= np.linalg.norm(n_true)
    s = np.sign(model[:3] @ n_true)
    np.testing.assert_allclose(s * model[:3], n_true, atol=2e-3)
    assert abs(s * model[3] - 0.3) < 2e-3
    assert np.all(np.diff(inl.astype(np.int64)) > 0)
    assert 0.69 * len(xyz) < len(inl) < 0.72 * len(xyz)
    assert st["best_count"] == len(inl)  # inliers are those of the minimal model (ransac.h:536-543)


def test_sphere_general_fit_matches_lstsq(orc):
    rng = np.random.default_rng(5)
    d = rng.normal(size=(4000, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    pts = np.array([0.3, -0.1, 0.7]) + d * (0.8 + rng.normal(0, 0.002, size=(4000, 1)))
    ok, m = orc.general_fit(orc.SPHERE, pts)
    A = np.c_[2 * pts, np.ones(len(pts))]
    w = np.linalg.lstsq(A, (pts ** 2).sum(1), rcond=None)[0]
    ref = np.r_[w[:3], np.sqrt(w[:3] @ w[:3] + w[3])]
    assert ok == 1
    np.testing.assert_allclose(m, ref, rtol=1e-9, atol=1e-11)


def test_umeyama_matches_numpy_kabsch(orc):
    rng = np.random.default_rng(9)
    src = rng.normal(size=(50, 3))
    R = synth.rotation_about((0.3, -1, 0.5), 41.0)
    t = np.array([0.5, -0.25, 2.0])
    dst = src @ R.T + t
    T = orc.umeyama(src, dst)
    np.testing.assert_allclose(T[:3, :3], R, atol=1e-12)
    np.testing.assert_allclose(T[:3, 3], t, atol=1e-12)
    # 3 points (rank-2 covariance): still a proper rotation mapping the triple
    T3 = orc.umeyama(src[:3], dst[:3])
    assert abs(np.linalg.det(T3[:3, :3]) - 1) < 1e-12
    np.testing.assert_allclose(src[:3] @ T3[:3, :3].T + T3[:3, 3], dst[:3], atol=1e-12)


def test_nearest_matches_bruteforce(orc):
    rng = np.random.default_rng(2)
    a = rng.uniform(0, 100, size=(33, 300))
    b = rng.uniform(0, 100, size=(33, 400))
    nn = orc.nearest(a, b)
    d2 = ((a.T[:, None, :] - b.T[None, :, :]) ** 2).sum(-1)
    np.testing.assert_array_equal(nn, d2.argmin(1))
    i0, i1 = orc.match_correspondence(a, b)
    back = d2.argmin(0)
    keep = [i for i in range(300) if back[d2.argmin(1)[i]] == i]
    np.testing.assert_array_equal(i0, keep)
    np.testing.assert_array_equal(i1, d2.argmin(1)[keep])


def test_segmentation_small_scene(orc):
    xyz = synth.make_c3(n=30000, seed=4)
    rc, planes, labels = orc.segment_plane_iterative(xyz, 0.01, 100, 0.05, seed=7)
    assert rc == 0 and 6 <= len(planes) <= 9
    big = [p for k, p in enumerate(planes) if (labels == k).sum() > 0.05 * len(xyz)]
    for p in big:  # every large cluster is one of the cube faces |x|,|y|,|z| = 1
        ax = np.argmax(np.abs(p[:3]))
        assert abs(abs(p[ax]) - 1) < 1e-3 and abs(abs(p[3]) - 1) < 5e-3


@pytest.mark.parametrize("name", ["c1_plane", "small_sphere", "small_cylinder", "seg_small", "reg_small"])
def test_golden_vectors(orc, name):
    g = np.load(os.path.join(GOLD, name + ".npz"))
    if name == "c1_plane":
        xyz = synth.make_c1()
        rc, model, inl, st = orc.ransac_fit(orc.PLANE, xyz, thr=0.01, max_it=100, prob=0.9999, seed=1)
    elif name == "small_sphere":
        xyz, nrm = synth.make_c2(n=20000, seed=11)
        rc, model, inl, st = orc.ransac_fit(orc.SPHERE, xyz, thr=0.01, max_it=300, prob=0.9999, seed=2)
    elif name == "small_cylinder":
        xyz, nrm = synth.make_c2(n=20000, seed=11)
        rc, model, inl, st = orc.ransac_fit(orc.CYLINDER, xyz, nrm, thr=0.01, max_it=300, prob=0.9999, seed=3)
    elif name == "seg_small":
        xyz = synth.make_c3(n=30000, seed=4)
        rc, planes, labels = orc.segment_plane_iterative(xyz, 0.01, 100, 0.05, seed=7)
        np.testing.assert_array_equal(planes, g["planes"])
        np.testing.assert_array_equal(labels, g["labels"])
        return
    else:
        d = synth.make_c4(n=3000, seed=5)
        i0, i1 = orc.match_correspondence(d["src_feat"], d["dst_feat"])
        np.testing.assert_array_equal(i0, g["i0"])
        np.testing.assert_array_equal(i1, g["i1"])
        rc, T, st = orc.ransac_registration(d["src"], d["dst"], i0, i1, thr=0.02, max_iter=2000, edge_thr=0.9,
                                            confidence=0.999, seed=1)
        np.testing.assert_array_equal(T, g["T"])
        assert st["best_index"] == int(g["best_index"]) and st["best_count"] == int(g["best_count"])
        return
    if name == "small_sphere":   # golden = the reference's SVD least squares; oracle = normal equations
        np.testing.assert_allclose(model, g["model"], rtol=1e-9, atol=1e-12)
    else:
        np.testing.assert_array_equal(model, g["model"])
    np.testing.assert_array_equal(inl, g["inl"])
    for k in ("best_index", "best_count", "iterations_run", "stop_index"):
        assert st[k] == int(g[k]), k


def test_golden_real_scan(orc):
    """tests/golden/real_scan.npz: every 4th vertex of the reference's demo scan + the compiled reference's
    outputs for the demo's parameters (tools/make_golden.py) -- real sensor data"""
    g = np.load(os.path.join(GOLD, "real_scan.npz"))
    xyz = g["xyz"]
    rc, model, inl, st = orc.ransac_fit(orc.PLANE, xyz, thr=0.01, max_it=100, prob=0.9999, seed=1)
    assert rc == int(g["rc"]) and st["iterations_run"] == int(g["iterations_run"])
    np.testing.assert_array_equal(inl, g["inl"])
    np.testing.assert_array_equal(model, g["model"])
    rc, planes, labels = orc.segment_plane_iterative(xyz, 0.01, 100, 0.1, seed=3)
    assert rc == 0
    np.testing.assert_array_equal(planes, g["planes"])
    np.testing.assert_array_equal(labels.astype(np.int64), g["labels"])


@pytest.mark.parametrize("conf", [0.999, 1.0])
@pytest.mark.parametrize("seed,outliers", [(2, 0.0), (5, 0.6)])
def test_registration_two_independent_restatements_agree(orc, seed, outliers, conf):
    """rows a19/a20 have no compiled reference (Open3D is a third-party dependency that is not in the tree): the
    oracle's C++ restatement of Appendix B must agree with an independently written numpy one (tests/reg_numpy_ref.py:
    Kabsch via np.linalg.svd, vectorised scoring) on the same recorded sample table -- loop statistics exactly,
    the 4 x 4 transform to 1e-9."""
    import reg_numpy_ref as ref
    d = synth.make_c4(n=2500, seed=seed)
    i0, i1 = orc.match_correspondence(d["src_feat"], d["dst_feat"])
    i0, i1 = i0.astype(np.int64), i1.astype(np.int64)
    if outliers:
        rng = np.random.default_rng(seed)
        bad = rng.uniform(size=len(i1)) < outliers
        i1 = i1.copy()
        i1[bad] = rng.integers(0, 2500, int(bad.sum()))
    max_iter = 1500
    picks = orc.reg_sample_table(seed, len(i0), max_iter)
    rc, oT, ost = orc.ransac_registration(d["src"], d["dst"], i0, i1, thr=0.02, max_iter=max_iter, edge_thr=0.9,
                                          confidence=conf, seed=seed)
    T, st = ref.ransac_registration_np(d["src"], d["dst"], i0, i1, picks, 0.02, max_iter, 0.9, conf)
    assert rc == 1
    for k in ("best_index", "best_count", "evaluated", "stop_index"):
        assert st[k] == ost[k], (k, st, ost)
    np.testing.assert_allclose(T, oT, rtol=0, atol=1e-9)
    assert abs(st["best_rmse"] - ost["best_rmse"]) <= 1e-12
    # least squares over all inlier pairs, with and without scaling (LeastSquareSolver = Eigen::umeyama)
    for scaling in (False, True):
        np.testing.assert_allclose(orc.umeyama(d["src"][i0], 1.7 * d["dst"][i1] if scaling else d["dst"][i1], scaling),
                                   ref.umeyama_np(d["src"][i0], 1.7 * d["dst"][i1] if scaling else d["dst"][i1], scaling),
                                   rtol=0, atol=1e-9)


def test_open3d_pin(orc):
    """rows a19 / a20 against a real Open3D: consumes tests/golden/reg_open3d.npz, which tools/pin_open3d.py writes on a
    box where `import open3d` works (none in this image: skipped until the file exists)."""
    path = os.path.join(GOLD, "reg_open3d.npz")
    if not os.path.exists(path):
        pytest.skip("tests/golden/reg_open3d.npz absent: run tools/pin_open3d.py where open3d is installed")
    g = np.load(path)
    d = synth.make_c4(n=3000, seed=5)
    i0, i1 = g["i0"], g["i1"]
    for t, T3 in zip(g["triples"], g["T3"]):             # the 3-point solve, transform by transform
        np.testing.assert_allclose(orc.umeyama(d["src"][i0[t]], d["dst"][i1[t]], False), T3, rtol=0, atol=1e-9)
    np.testing.assert_allclose(orc.umeyama(d["src"][i0], d["dst"][i1], False), g["T_ls"], rtol=0, atol=1e-9)
    np.testing.assert_allclose(orc.umeyama(d["src"][i0], d["dst"][i1], True), g["T_ls_scale"], rtol=0, atol=1e-9)
    rc, T, st = orc.ransac_registration(d["src"], d["dst"], i0, i1, thr=float(g["thr"]), max_iter=int(g["max_iter"]),
                                        edge_thr=float(g["edge"]), confidence=0.999, seed=1)
    # different sample streams, one dominant alignment: same inlier population, transforms within the 3-point noise
    assert abs(st["best_count"] / len(i0) - float(g["fitness"])) < 0.02
    assert np.linalg.norm(T - g["T"]) < 0.05


def test_fpfh_and_icp_restatements_agree(orc):
    """f3 / f4 (Open3D's ComputeFPFHFeature and point-to-point RegistrationICP, restated -- parity unpinned): the C++
    checker against an independently written numpy one on a small cloud, plus the histogram-mass property"""
    import fpfh_numpy_ref as ref
    d = synth.make_surface_pair(n=600, seed=3)
    f = orc.fpfh(d["src"], d["src_nrm"], 0.25, 40)
    g = ref.fpfh(d["src"], d["src_nrm"], 0.25, 40)
    assert f.shape == (33, 600)
    # libm's atan2 / acos may put a pair on the other side of a bin edge in one of the two: allow isolated flips
    assert np.mean(np.isclose(f, g, rtol=1e-9, atol=1e-9)) > 0.999
    sums = f.T.reshape(600, 3, 11).sum(2)
    assert np.allclose(sums, 200.0, atol=1e-8)      # SPFH (100) + normalised neighbour sum (100) per sub-histogram
    T0 = np.eye(4)
    T0[:3, 3] = [0.02, -0.01, 0.015]
    src = d["src"][:400]
    dst = (d["src"] @ d["T_true"][:3, :3].T + d["T_true"][:3, 3])
    Ti = d["T_true"] @ np.linalg.inv(T0)             # a start close to the truth
    T, fit, rmse, it = orc.icp(src, dst, 0.1, Ti, 30)
    T2, fit2, rmse2, it2 = ref.icp(src, dst, 0.1, Ti, 30)
    assert it == it2 and abs(fit - fit2) < 1e-12 and abs(rmse - rmse2) < 1e-9
    np.testing.assert_allclose(T, T2, rtol=0, atol=1e-9)
    assert np.linalg.norm(T - d["T_true"]) < 1e-6 and fit == 1.0
