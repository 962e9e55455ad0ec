"""GPU parity tests of the RANSAC path: libm3d_b200.so (through the C-ABI) against the CPU oracle
on the same seeded inputs -- bit-exact for models of the minimal fits, inlier counts, inlier
index sets and loop statistics; tolerance (stated per test) for the least-squares refits."""
import os

import numpy as np
import pytest

from misc3d_b200 import synth

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")
KINDS = [0, 1, 2]


def _cloud(kind, n, seed):
    xyz, nrm = synth.make_c2(n=n, seed=seed)
    return xyz, (nrm if kind == 2 else None)


def _oracle_rows(orc, kind, xyz, nrm, table, thr):
    valid = np.zeros(len(table), np.uint8)
    counts = np.zeros(len(table), np.uint64)
    models = np.zeros((len(table), 8))
    for i, row in enumerate(table):
        idx = np.sort(row)
        ok, m = orc.minimal_fit(kind, xyz[idx], None if nrm is None else nrm[idx])
        if ok:
            valid[i] = 1
            models[i, :len(m)] = m
            counts[i], _ = orc.evaluate(kind, xyz, m, thr)
    return valid, counts, models


@pytest.mark.parametrize("kind", KINDS)
@pytest.mark.parametrize("n,rows", [(20000, 300), (1025, 64), (5000, 2500)])
def test_score_samples_bit_exact(ctx, capi, orc, kind, n, rows):
    xyz, nrm = _cloud(kind, n, seed=100 + kind)
    table = capi.sample_table(7 + kind, n, capi.KSAMPLE[kind], rows)
    cloud = ctx.upload(xyz, nrm)
    counts, models, valid = ctx.score_samples(kind, cloud, table, 0.01)
    ovalid, ocounts, omodels = _oracle_rows(orc, kind, xyz, nrm, table, 0.01)
    np.testing.assert_array_equal(valid, ovalid)
    np.testing.assert_array_equal(models.view(np.uint64), omodels.view(np.uint64))  # bit-exact, NaN included
    np.testing.assert_array_equal(counts, ocounts)
    # the fp64 reference-order kernel gives the same counts
    counts2, _, _ = ctx.score_samples(kind, cloud, table, 0.01, flags=capi.FLAG_EXACT_ONLY, want_models=False)
    np.testing.assert_array_equal(counts2, ocounts)
    cloud.free()


def _check_fit(ctx, capi, orc, kind, xyz, nrm, thr, max_it, prob, seed, rtol_refit=1e-9):
    rc, model, inl, st = ctx.ransac_fit(kind, xyz, nrm, thr, max_it, prob, seed)
    orc_rc, omodel, oinl, ost = orc.ransac_fit(kind, xyz, nrm, thr=thr, max_it=max_it, prob=prob, seed=seed)
    assert rc == orc_rc
    for key in ("best_index", "best_count", "iterations_run", "stop_index", "found", "refit_ok"):
        assert st[key] == ost[key], (key, st, ost)
    np.testing.assert_array_equal(inl, oinl)  # bit-exact inlier index set, ascending
    if ost["found"]:
        assert abs(st["best_rmse"] - ost["best_rmse"]) <= 1e-12 * max(1.0, abs(ost["best_rmse"]))
    if kind == 2:  # GeneralFit is a no-op for the cylinder: the minimal model, bit-exact
        np.testing.assert_array_equal(model, omodel)
    else:          # least-squares refit: summation order differs (the reference's own OMP reduction
        #            is order-nondeterministic) -> 1e-9 relative / 1e-12 absolute
        np.testing.assert_allclose(model, omodel, rtol=rtol_refit, atol=1e-12)
    return st, ost


def test_c1_plane_golden(ctx, capi, orc):
    """BASELINE config C1: fit_plane, 50k points, 100 iterations, seed 1 -- against the committed golden."""
    g = np.load(os.path.join(GOLD, "c1_plane.npz"))
    xyz = synth.make_c1()
    rc, model, inl, st = ctx.ransac_fit(capi.PLANE, xyz, None, 0.01, 100, 0.9999, 1)
    assert rc == int(g["rc"])
    np.testing.assert_array_equal(inl, g["inl"])
    np.testing.assert_allclose(model, g["model"], rtol=0, atol=1e-12)
    for key in ("best_index", "best_count", "iterations_run", "stop_index"):
        assert st[key] == int(g[key]), key
    _check_fit(ctx, capi, orc, capi.PLANE, xyz, None, 0.01, 100, 0.9999, 1)


@pytest.mark.parametrize("kind,name,seed", [(1, "small_sphere", 2), (2, "small_cylinder", 3)])
def test_small_goldens(ctx, capi, orc, kind, name, seed):
    g = np.load(os.path.join(GOLD, name + ".npz"))
    xyz, nrm = synth.make_c2(n=20000, seed=11)
    rc, model, inl, st = ctx.ransac_fit(kind, xyz, nrm if kind == 2 else None, 0.01, 300, 0.9999, seed)
    np.testing.assert_array_equal(inl, g["inl"])
    np.testing.assert_allclose(model, g["model"], rtol=1e-9, atol=1e-12)
    for key in ("best_index", "best_count", "iterations_run", "stop_index"):
        assert st[key] == int(g[key]), key


@pytest.mark.parametrize("kind", KINDS)
@pytest.mark.parametrize("prob", [0.9999, 1.0])
def test_fit_parity_with_and_without_early_exit(ctx, capi, orc, kind, prob):
    xyz, nrm = _cloud(kind, 60000, seed=31)
    st, ost = _check_fit(ctx, capi, orc, kind, xyz, nrm, 0.01, 700, prob, seed=17 + kind)
    if prob < 1 and kind == 0:
        assert ost["stop_index"] < 700


@pytest.mark.parametrize("kind", KINDS)
def test_fit_parity_offset_and_scaled_clouds(ctx, capi, orc, kind):
    """coordinates far from the origin / millimetre scale: the fp32 copy is centred, the guard band scales"""
    xyz, nrm = _cloud(kind, 30000, seed=5)
    _check_fit(ctx, capi, orc, kind, xyz + np.array([1500.0, -2200.0, 870.0]), nrm, 0.01, 200, 1.0, seed=3)
    _check_fit(ctx, capi, orc, kind, xyz * 1000.0, nrm, 10.0, 200, 1.0, seed=4, rtol_refit=1e-8)


def test_fit_parity_nonfinite_points(ctx, capi, orc):
    xyz = synth.make_c1(n=8000, seed=2)
    xyz[17] = np.nan
    xyz[4000, 1] = np.inf
    _check_fit(ctx, capi, orc, capi.PLANE, xyz, None, 0.01, 150, 1.0, seed=9)


def test_degenerate_inputs(ctx, capi, orc):
    # all points identical: every MinimalFit fails -> no model, zero parameters, FitModel false
    xyz = np.tile([[0.5, 0.25, -1.0]], (100, 1))
    rc, model, inl, st = ctx.ransac_fit(capi.PLANE, xyz, None, 0.01, 50, 0.9999, 1)
    orc_rc, omodel, oinl, ost = orc.ransac_fit(orc.PLANE, xyz, thr=0.01, max_it=50, prob=0.9999, seed=1)
    assert (rc, st["found"], st["iterations_run"]) == (orc_rc, ost["found"], ost["iterations_run"]) == (0, 0, 0)
    assert len(inl) == 0 and not model.any()
    # exactly k points; a perfect plane (fitness 1 -> current_iteration = 0, ransac.h:607-609)
    xyz = np.array([[0, 0, 0.0], [1, 0, 0], [0, 1, 0]])
    _check_fit(ctx, capi, orc, capi.PLANE, xyz, None, 0.01, 10, 0.9999, seed=1)
    plane = np.c_[np.random.default_rng(1).uniform(-1, 1, (500, 2)), np.zeros(500)]
    _check_fit(ctx, capi, orc, capi.PLANE, plane, None, 0.01, 100, 0.9999, seed=2)


def test_error_codes_mirror_the_reference_throws(ctx, capi):
    xyz = synth.make_c1(n=100, seed=1)
    with pytest.raises(capi.M3DError) as e:
        ctx.ransac_fit(capi.PLANE, xyz[:2], None, 0.01, 10, 0.9999, 1)
    assert e.value.code == capi.ERR_TOO_FEW_POINTS
    for bad in (0.0, -1.0, 1.5):
        with pytest.raises(capi.M3DError) as e:
            ctx.ransac_fit(capi.PLANE, xyz, None, 0.01, 10, bad, 1)
        assert e.value.code == capi.ERR_PROBABILITY
    with pytest.raises(capi.M3DError) as e:
        ctx.ransac_fit(capi.CYLINDER, xyz, None, 0.01, 10, 0.9999, 1)
    assert e.value.code == capi.ERR_NO_NORMALS


@pytest.mark.parametrize("kind", KINDS)
def test_evaluate_model_sequential_error_is_bit_exact(ctx, capi, orc, kind):
    xyz, nrm = _cloud(kind, 30000, seed=77)
    table = capi.sample_table(5, len(xyz), capi.KSAMPLE[kind], 400)
    cloud = ctx.upload(xyz, nrm)
    counts, models, valid = ctx.score_samples(kind, cloud, table, 0.01)
    best = int(np.argmax(counts * valid))
    m = models[best, :capi.NPARAM[kind]]
    ocnt, oerr = orc.evaluate(kind, xyz, m, 0.01)
    cnt, err = ctx.evaluate_model(kind, cloud, m, 0.01, sequential=True)
    assert cnt == ocnt and err == oerr  # the reference's index-order sum, bit for bit
    cnt, err = ctx.evaluate_model(kind, cloud, m, 0.01, sequential=False)
    assert cnt == ocnt and abs(err - oerr) <= 1e-12 * oerr
    cloud.free()


def test_count_ties_resolved_like_the_sequential_loop(ctx, capi, orc):
    """few points, coarse threshold: many hypotheses share an inlier count, so the strict
    rmse tie-break (ransac.h:595-596) decides -- the GPU path must pick the same hypothesis"""
    rng = np.random.default_rng(4)
    for trial in range(6):
        xyz = np.round(rng.uniform(-1, 1, (40, 3)), 1)  # lattice points -> exact ties and duplicates
        _check_fit(ctx, capi, orc, capi.PLANE, xyz, None, 0.15, 400, 1.0, seed=trial)


def test_segmentation_parity(ctx, capi, orc):
    g = np.load(os.path.join(GOLD, "seg_small.npz"))
    xyz = synth.make_c3(n=30000, seed=4)
    rc, planes, labels, ms = ctx.segment_plane_iterative(xyz, 0.01, 100, 0.05, seed=7)
    assert rc == int(g["rc"]) == 0
    np.testing.assert_array_equal(labels, g["labels"])  # bit-exact cluster membership
    np.testing.assert_allclose(planes, g["planes"], rtol=0, atol=1e-12)
    # a second scene with different parameters against the live oracle
    xyz = synth.make_c3(n=50000, seed=12)
    rc, planes, labels, ms = ctx.segment_plane_iterative(xyz, 0.008, 60, 0.1, seed=3)
    orc_rc, oplanes, olabels = orc.segment_plane_iterative(xyz, 0.008, 60, 0.1, seed=3)
    assert rc == orc_rc
    np.testing.assert_array_equal(labels, olabels)
    np.testing.assert_allclose(planes, oplanes, rtol=0, atol=1e-12)
    # fewer than 3 points: warning + empty result in the reference (:14-17)
    rc, planes, labels, ms = ctx.segment_plane_iterative(xyz[:2], 0.01)
    assert rc == 0 and len(planes) == 0


def _full_size_rows_vs_cpu(orc, kind, xyz, nrm, table, counts, models, valid, thr, inl_of_best=None, best=None):
    """32 rows of a full-size launch (8 with the largest counts, the first 8, 16 random) against the CPU side on the
    SAME full-size cloud: MinimalFit models bit for bit and EvaluateModel counts from the oracle, and -- when the
    reference's own sources were compiled (oracle/_ref) -- the per-point distances of ransac.h itself for the winner,
    i.e. the identical inlier list."""
    rng = np.random.default_rng(99)
    sub = np.unique(np.r_[np.argsort(counts)[-8:], np.arange(8), rng.choice(len(table), 16, replace=False)])
    ovalid, ocounts, omodels = _oracle_rows(orc, kind, xyz, nrm, table[sub], thr)
    np.testing.assert_array_equal(valid[sub], ovalid)
    np.testing.assert_array_equal(models[sub].view(np.uint64), omodels.view(np.uint64))
    np.testing.assert_array_equal(counts[sub], ocounts)
    if inl_of_best is not None:
        import refc
        m = models[best, :{0: 4, 1: 4, 2: 7}[kind]]
        if refc.available():
            d = refc.distances(kind, m, xyz)        # the reference's CalcPointToModelDistance, point by point
        else:
            d = np.array([orc.distance(kind, m, q) for q in xyz[inl_of_best]])  # oracle: at least no false inlier
            assert np.all(d < thr)
            return
        np.testing.assert_array_equal(np.nonzero(d < thr)[0], inl_of_best)


@pytest.mark.parametrize("kind", KINDS)
def test_full_size_properties_c2(ctx, capi, orc, kind):
    """BASELINE config C2 size (1M points): fast counts == fp64 reference-order counts, inliers ascending and exactly
    {d < thr}; and, on the full 1M-point cloud, 32 hypotheses against the oracle + the winner's inlier list against
    the compiled reference's distances."""
    xyz, nrm = synth.make_c2()
    cloud = ctx.upload(xyz, nrm if kind == 2 else None)
    table = capi.sample_table(3, len(xyz), capi.KSAMPLE[kind], 4096)
    counts, models, valid = ctx.score_samples(kind, cloud, table, 0.01)
    sub = np.r_[np.argsort(counts)[-8:], np.arange(8)]
    exact, _, _ = ctx.score_samples(kind, cloud, table[sub], 0.01, flags=capi.FLAG_EXACT_ONLY, want_models=False)
    np.testing.assert_array_equal(counts[sub], exact)
    rc, model, inl, st = ctx.ransac_fit_cloud(kind, cloud, 0.01, 4096, 1.0, seed=3)
    assert st["best_count"] == counts.max() == len(inl)
    assert st["best_index"] == int(np.argmax(np.where(valid > 0, counts, 0)))  # first max, no ties expected
    assert np.all(np.diff(inl.astype(np.int64)) > 0)
    cnt, err = ctx.evaluate_model(kind, cloud, models[st["best_index"]], 0.01)
    assert cnt == len(inl)
    if kind == 0:  # independent numpy check of the inlier definition on the minimal model
        m = models[st["best_index"]]
        d = np.abs((m[0] * xyz[:, 0] + m[2] * xyz[:, 2]) + (m[1] * xyz[:, 1] + m[3])) / np.sqrt(m[0] ** 2 + m[1] ** 2 + m[2] ** 2)
        np.testing.assert_array_equal(np.nonzero(d < 0.01)[0], inl)
    _full_size_rows_vs_cpu(orc, kind, xyz, nrm if kind == 2 else None, table, counts, models, valid, 0.01,
                           inl_of_best=inl, best=st["best_index"])
    cloud.free()


def test_early_exit_in_a_later_wave(ctx, capi, orc):
    """15 % inliers: the adaptive limit is ~2.7k iterations, i.e. the skip test fires inside the
    third GPU wave (256, 1024, 4096, ...) -- stop index and iteration count must still be the
    sequential loop's"""
    xyz = synth.make_c1(n=40000, seed=6, inlier_frac=0.15)
    st, ost = _check_fit(ctx, capi, orc, capi.PLANE, xyz, None, 0.01, 20000, 0.9999, seed=5)
    assert 256 + 1024 < ost["stop_index"] < 20000 and st["evaluated"] >= ost["stop_index"]


def test_cloud_from_device_and_reuse(ctx, capi, orc):
    """a cloud already resident in HBM (torch tensor) is fitted repeatedly without new uploads"""
    import torch
    xyz, nrm = synth.make_c2(n=50000, seed=8)
    dx = torch.from_numpy(xyz).cuda()
    dn = torch.from_numpy(nrm).cuda()
    torch.cuda.synchronize()
    cloud = ctx.cloud_from_device(dx.data_ptr(), dn.data_ptr(), len(xyz))
    del dx, dn  # the library copied them
    for kind in KINDS:
        for seed in (1, 2):
            rc, model, inl, st = ctx.ransac_fit_cloud(kind, cloud, 0.01, 400, 0.9999, seed=seed)
            orc_rc, omodel, oinl, ost = orc.ransac_fit(kind, xyz, nrm if kind == 2 else None, thr=0.01, max_it=400,
                                                       prob=0.9999, seed=seed)
            assert rc == orc_rc and np.array_equal(inl, oinl) and st["best_index"] == ost["best_index"]
    cloud.free()


def test_launch_count_and_no_silent_fallback(ctx, capi):
    xyz = synth.make_c1(n=5000, seed=1)
    before = ctx.launches
    ctx.ransac_fit(capi.PLANE, xyz, None, 0.01, 100, 0.9999, 1)
    assert ctx.launches - before >= 8  # cloud preparation, scoring, resolve, refine passes all ran on the GPU


# ---------------------------------------------------------------------------------------------------
# culling score kernel (score_cull.cuh): Morton-ordered copy + bounding spheres.  Counts must equal
# the dense fp32 kernel's, the fp64 reference-order kernel's and the oracle's on every kind of cloud.
def _special_clouds():
    rng = np.random.default_rng(77)
    n = 9000
    out = {}
    xyz, nrm = synth.make_c2(n=50001, seed=5)             # not a multiple of 32 / 1024
    out["c2_50001"] = (xyz, nrm)
    out["offset_1e4"] = (xyz[:n] + np.array([1.0e4, -2.0e4, 5.0e3]), nrm[:n])   # far from the origin
    out["tiny_scale"] = (xyz[:n] * 1e-3, nrm[:n])
    planar = np.c_[rng.uniform(-1, 1, (n, 2)), np.zeros(n)]          # exactly planar: flat cells
    out["planar"] = (planar, np.tile([0.0, 0, 1.0], (n, 1)))
    dup = np.repeat(rng.uniform(-1, 1, (n // 50, 3)), 50, axis=0)     # heavy duplicates: R = 0 cells
    out["duplicates"] = (dup, rng.normal(size=dup.shape))
    line = np.outer(np.linspace(-1, 1, n), [1.0, 2.0, -0.5]) + rng.normal(0, 1e-4, (n, 3))
    out["near_line"] = (line, rng.normal(size=line.shape))
    clustered = np.r_[rng.normal(0, 1e-3, (n // 2, 3)), rng.normal(0, 1.0, (n // 2, 3)) + 50.0]
    out["two_scales"] = (clustered, rng.normal(size=clustered.shape))
    return out


@pytest.mark.parametrize("kind", KINDS)
@pytest.mark.parametrize("name", ["c2_50001", "offset_1e4", "tiny_scale", "planar", "duplicates", "near_line",
                                  "two_scales"])
def test_cull_kernel_counts_equal_dense_and_exact(ctx, capi, orc, kind, name):
    xyz, nrm = _special_clouds()[name]
    n = len(xyz)
    thr = 0.01 * (1e-3 if name == "tiny_scale" else 1.0)
    rows = 2304 if name == "c2_50001" else 600
    table = capi.sample_table(3 + kind, n, capi.KSAMPLE[kind], rows)
    cloud = ctx.upload(xyz, nrm)
    c_cull, m_cull, v_cull = ctx.score_samples(kind, cloud, table, thr)
    c_dense, m_dense, v_dense = ctx.score_samples(kind, cloud, table, thr, flags=capi.FLAG_DENSE)
    c_exact, _, _ = ctx.score_samples(kind, cloud, table, thr, flags=capi.FLAG_EXACT_ONLY, want_models=False)
    np.testing.assert_array_equal(c_cull, c_exact)
    np.testing.assert_array_equal(c_dense, c_exact)
    np.testing.assert_array_equal(m_cull.view(np.uint64), m_dense.view(np.uint64))
    np.testing.assert_array_equal(v_cull, v_dense)
    # and the oracle, on a subset of the rows (it is a scalar CPU loop)
    sub = table[:40]
    ovalid, ocounts, _ = _oracle_rows(orc, kind, xyz, nrm if kind == 2 else None, sub, thr)
    np.testing.assert_array_equal(c_cull[:40], ocounts)
    cloud.free()


def test_cull_kernel_large_hypothesis_batch(ctx, capi):
    """C2-like shape at reduced size: 200k points x 6000 hypotheses, cull == dense for all three kinds"""
    xyz, nrm = synth.make_c2(n=200000, seed=9)
    cloud = ctx.upload(xyz, nrm)
    for kind in KINDS:
        table = capi.sample_table(11 + kind, len(xyz), capi.KSAMPLE[kind], 6000)
        c_cull, _, _ = ctx.score_samples(kind, cloud, table, 0.01, want_models=False)
        c_dense, _, _ = ctx.score_samples(kind, cloud, table, 0.01, flags=capi.FLAG_DENSE, want_models=False)
        np.testing.assert_array_equal(c_cull, c_dense)
    cloud.free()


@pytest.mark.parametrize("seed", range(12))
def test_cull_kernel_randomized(ctx, capi, seed):
    """random mixtures, scales, offsets, thresholds and sizes: the culling kernel's counts equal the fp64
    reference-order kernel's for all three primitives"""
    rng = np.random.default_rng(1000 + seed)
    n = int(rng.integers(2048, 70000))
    scale = 10.0 ** rng.uniform(-2, 3)
    offset = rng.uniform(-1, 1, 3) * scale * 10.0 ** rng.uniform(-1, 2)
    parts = []
    m = n
    for _ in range(int(rng.integers(1, 5))):          # a few primitives ...
        k = int(m * rng.uniform(0.1, 0.5))
        m -= k
        kind = rng.integers(0, 3)
        if kind == 0:
            u = rng.uniform(-1, 1, (k, 2))
            pts = np.c_[u, 0.01 * rng.normal(size=k) * rng.integers(0, 2)]
            q, _ = np.linalg.qr(rng.normal(size=(3, 3)))
            pts = pts @ q.T
        elif kind == 1:
            d = rng.normal(size=(k, 3))
            pts = d / np.linalg.norm(d, axis=1, keepdims=True) * rng.uniform(0.05, 1.0) + rng.uniform(-0.5, 0.5, 3)
        else:
            a = rng.uniform(0, 2 * np.pi, k)
            pts = np.c_[0.2 * np.cos(a), 0.2 * np.sin(a), rng.uniform(-1, 1, k)]
        parts.append(pts + 0.003 * rng.normal(size=(k, 3)))
    parts.append(rng.uniform(-1, 1, (m, 3)))          # ... and uniform outliers
    xyz = np.concatenate(parts) * scale + offset
    xyz = xyz[rng.permutation(len(xyz))]
    nrm = rng.normal(size=xyz.shape)
    nrm /= np.linalg.norm(nrm, axis=1, keepdims=True)
    thr = scale * 10.0 ** rng.uniform(-3, -1)
    cloud = ctx.upload(xyz, nrm)
    for kind in KINDS:
        table = capi.sample_table(seed * 7 + kind, len(xyz), capi.KSAMPLE[kind], 640)
        c_cull, m_cull, v_cull = ctx.score_samples(kind, cloud, table, thr)
        c_exact, _, _ = ctx.score_samples(kind, cloud, table, thr, flags=capi.FLAG_EXACT_ONLY, want_models=False)
        np.testing.assert_array_equal(c_cull, c_exact, err_msg=f"seed {seed} kind {kind} n {len(xyz)} scale {scale} thr {thr}")
    cloud.free()


@pytest.mark.parametrize("kind", KINDS)
def test_classified_split_counts_equal_exact(ctx, capi, orc, kind):
    """M3D_FLAG_CLASSIFY: hypotheses whose shell passes through much of the cloud go to the dense kernel, the
    rest to the culling kernel (what large waves do by default); the counts are those of the fp64 kernel.
    The C1-style cloud (70 % of the points on one plane) makes the dense part non-trivial."""
    for name, (xyz, nrm) in {"c1": (synth.make_c1(n=60000, seed=3), None), "c2": synth.make_c2(n=60000, seed=4)}.items():
        if nrm is None:
            rng = np.random.default_rng(1)
            nrm = rng.normal(size=xyz.shape)
            nrm /= np.linalg.norm(nrm, axis=1, keepdims=True)
        table = capi.sample_table(21 + kind, len(xyz), capi.KSAMPLE[kind], 3000)
        cloud = ctx.upload(xyz, nrm)
        c_split, m_split, v_split = ctx.score_samples(kind, cloud, table, 0.01, flags=capi.FLAG_CLASSIFY)
        c_exact, m_exact, v_exact = ctx.score_samples(kind, cloud, table, 0.01, flags=capi.FLAG_EXACT_ONLY)
        np.testing.assert_array_equal(c_split, c_exact, err_msg=name)
        np.testing.assert_array_equal(v_split, v_exact)
        np.testing.assert_array_equal(m_split.view(np.uint64), m_exact.view(np.uint64))
        cloud.free()
    # a whole fit through the split path equals the oracle's
    xyz = synth.make_c1(n=60000, seed=3)
    if kind == 0:
        rc, model, inl, st = ctx.ransac_fit(kind, xyz, None, 0.01, 3000, 1.0, seed=5, flags=capi.FLAG_CLASSIFY)
        orc_rc, omodel, oinl, ost = orc.ransac_fit(kind, xyz, None, thr=0.01, max_it=3000, prob=1.0, seed=5)
        assert rc == orc_rc and st["best_index"] == ost["best_index"] and st["best_count"] == ost["best_count"]
        np.testing.assert_array_equal(inl, oinl)


def test_large_wave_default_split_equals_dense(ctx, capi):
    """a wave of >= 12288 rows is pre-sorted into culled / dense hypotheses by default (config C5's shape at
    reduced size: 70 % of the points on one plane); counts equal the dense kernel's"""
    xyz = synth.make_c1(n=300000, seed=12)
    cloud = ctx.upload(xyz, None)
    table = capi.sample_table(5, len(xyz), 3, 14000)
    c_default, _, _ = ctx.score_samples(0, cloud, table, 0.01, want_models=False)
    c_dense, _, _ = ctx.score_samples(0, cloud, table, 0.01, flags=capi.FLAG_DENSE, want_models=False)
    np.testing.assert_array_equal(c_default, c_dense)
    assert (c_default > 0.6 * len(xyz)).sum() > 1000   # many hypotheses ARE the dominant plane
    cloud.free()


def test_full_size_properties_c3(ctx, capi, orc):
    """BASELINE config C3 size (2M points, six-plane scene): size-independent properties of the segmentation --
    clusters disjoint, the stopping rule of iterative_plane_segmentation.cpp:28-29, the large clusters are faces
    of the box, labelled points lie near their plane, the same result when run again -- and, since the
    sequential oracle still finishes in seconds at this size, label-for-label equality with it."""
    xyz = synth.make_c3()
    n = len(xyz)
    rc, planes, labels, ms = ctx.segment_plane_iterative(xyz, 0.01, 100, 0.05, seed=11)
    assert rc == 0 and len(planes) >= 6
    none = np.uint64(0xFFFFFFFFFFFFFFFF)
    assigned = labels != none
    sizes = np.array([int((labels == k).sum()) for k in range(len(planes))])
    assert np.all(sizes > 0) and sizes.sum() == int(assigned.sum())
    target = int((1 - 0.05) * n)                          # while (count < (size_t)((1 - min_ratio) * N))
    assert sizes.sum() >= target and sizes[:-1].sum() < target
    faces = set()
    for k in range(len(planes)):
        nrm, d = planes[k, :3], planes[k, 3]
        assert abs(np.linalg.norm(nrm) - 1) < 1e-9
        dist = np.abs(xyz[labels == k] @ nrm + d)
        assert np.mean(dist < 0.012) > 0.9                # inliers of the minimal model vs the refitted plane
        if sizes[k] > 0.05 * n:                           # a big cluster is (part of) a face of [-1, 1]^3
            axis = int(np.argmax(np.abs(nrm)))
            assert abs(abs(nrm[axis]) - 1) < 1e-3 and abs(abs(d) - 1) < 1e-2
            faces.add((axis, int(np.sign(nrm[axis] * d))))
    assert len(faces) == 6
    rc2, planes2, labels2, _ = ctx.segment_plane_iterative(xyz, 0.01, 100, 0.05, seed=11)
    np.testing.assert_array_equal(labels2, labels)
    np.testing.assert_array_equal(planes2, planes)
    orc_rc, oplanes, olabels = orc.segment_plane_iterative(xyz, 0.01, 100, 0.05, seed=11)
    assert orc_rc == 0 and len(oplanes) == len(planes)
    np.testing.assert_array_equal(labels, olabels)
    np.testing.assert_allclose(planes, oplanes, rtol=1e-9, atol=1e-12)


def test_full_size_properties_c5(ctx, capi, orc):
    """BASELINE config C5 size (4M points x 100k plane hypotheses, one GPU's view of it): the default path
    (waves pre-sorted into culled / dense hypotheses) and the dense kernel agree on the winner, the winner's
    count is the fp64 count of its minimal model, and the inlier list is exactly {d < thr} of that model."""
    xyz = synth.make_c5()
    cloud = ctx.upload(xyz, None)
    rc, model, inl, st = ctx.ransac_fit_cloud(0, cloud, 0.01, 100000, 1.0, seed=1)
    rc_d, model_d, inl_d, st_d = ctx.ransac_fit_cloud(0, cloud, 0.01, 100000, 1.0, seed=1, flags=capi.FLAG_DENSE)
    assert rc == rc_d == 1
    for key in ("best_index", "best_count", "iterations_run", "stop_index"):
        assert st[key] == st_d[key], key
    np.testing.assert_array_equal(inl, inl_d)
    np.testing.assert_array_equal(model, model_d)
    assert st["iterations_run"] <= 100000 and st["stop_index"] == 100000
    table = capi.sample_table(1, len(xyz), 3, st["best_index"] + 1)
    counts, models, valid = ctx.score_samples(0, cloud, table[-1:], 0.01, flags=capi.FLAG_EXACT_ONLY)
    assert valid[0] == 1 and counts[0] == st["best_count"] == len(inl)
    m = models[0]
    d = np.abs((m[0] * xyz[:, 0] + m[2] * xyz[:, 2]) + (m[1] * xyz[:, 1] + m[3])) / np.sqrt(m[0] ** 2 + m[1] ** 2 + m[2] ** 2)
    np.testing.assert_array_equal(np.nonzero(d < 0.01)[0], inl)
    assert np.all(np.diff(inl.astype(np.int64)) > 0) and len(inl) > 0.69 * len(xyz)
    # the full 4M-point cloud against the CPU side: 32 rows of the launch + the winner's inlier list
    table4k = capi.sample_table(1, len(xyz), 3, 4096)
    c4k, m4k, v4k = ctx.score_samples(0, cloud, table4k, 0.01)
    _full_size_rows_vs_cpu(orc, 0, xyz, None, table4k, c4k, m4k, v4k, 0.01)
    import refc
    if refc.available():
        d = refc.distances(0, m[:4], xyz)
        np.testing.assert_array_equal(np.nonzero(d < 0.01)[0], inl)
    cloud.free()


# ---------------------------------------------------------------------------------------------------
# the device-side loop (probability == 1): sample table drawn on the GPU, arg-best record instead of the
# host replay (csrc/loop_kernels.cuh)
@pytest.mark.parametrize("seed,n,k,rows", [(1, 1_000_000, 3, 10_000), (2, 1_000_000, 4, 10_000), (3, 1_000_000, 2, 80_000),
                                           (4, 5000, 4, 20_000), (5, 700, 3, 3000), (6, 40, 3, 400), (7, 2, 2, 100),
                                           (8, 4_000_000, 3, 100_000), (9, 65536, 3, 1), (10, 123457, 2, 624 * 5),
                                           # long tables: segments started from jumped-ahead generator states
                                           (11, 1_000_000, 3, 80_000), (12, 1_000_000, 3, 300_000),
                                           (13, 1_000_000, 4, 30_000), (14, 1_000_000, 3, 18_600),
                                           (15, 1_000_000, 3, 18_500), (4294967295, 2_000_000, 3, 160_000)])
def test_sample_table_device_equals_host(ctx, capi, seed, n, k, rows):
    """utils.h:81-97 (mt19937, % size, duplicate rejection) drawn on the GPU == the host stream, bit for bit;
    small clouds have many rejected draws (row boundaries shift), tiny ones make the device give up (None)"""
    dev = ctx.sample_table_device(seed, n, k, rows)
    host = capi.sample_table(seed, n, k, rows)
    if dev is None:
        assert n < 200  # gave up: duplicate-heavy stream, the host draw is used
    else:
        np.testing.assert_array_equal(dev, host)


def test_device_loop_many_ties_and_perfect_fit(ctx, capi, orc):
    """probability 1: (a) more count ties than the 8 a record lists -> host replay of the wave; (b) a hypothesis
    with fitness 1 stops the reference's loop (ransac.h:607-609) even without the adaptive exit"""
    rng = np.random.default_rng(11)
    xyz = np.round(rng.uniform(-1, 1, (30, 3)), 0)  # 27 lattice sites: massive ties
    _check_fit(ctx, capi, orc, capi.PLANE, xyz, None, 0.3, 600, 1.0, seed=2)
    plane = np.c_[rng.uniform(-1, 1, (3000, 2)), np.zeros(3000)]
    st, ost = _check_fit(ctx, capi, orc, capi.PLANE, plane, None, 0.01, 300, 1.0, seed=3)
    assert ost["stop_index"] < 300
    # resident cloud with device normals: the cylinder's table is drawn on the device too
    xyz, nrm = synth.make_c2(n=30000, seed=12)
    cloud = ctx.upload(xyz, nrm)
    for kind in KINDS:
        rc, model, inl, st = ctx.ransac_fit_cloud(kind, cloud, 0.01, 900, 1.0, seed=21 + kind)
        orc_rc, omodel, oinl, ost = orc.ransac_fit(kind, xyz, nrm if kind == 2 else None, thr=0.01, max_it=900, prob=1.0,
                                                   seed=21 + kind)
        assert rc == orc_rc and np.array_equal(inl, oinl)
        for key in ("best_index", "best_count", "iterations_run", "stop_index", "found"):
            assert st[key] == ost[key], (key, st, ost)
    cloud.free()


def test_device_loop_several_waves(ctx, capi, orc):
    """more hypotheses than one device-side wave holds (2^18 per rank): the best is carried across waves"""
    xyz = synth.make_c1(n=2500, seed=13)
    _check_fit(ctx, capi, orc, capi.PLANE, xyz, None, 0.01, 300_000, 1.0, seed=5)


def test_sphere_refit_of_degenerate_inliers_is_finite(ctx, capi):
    """coplanar inliers make the sphere's least-squares system rank deficient: the refit must stay finite (the
    reference's bdcSvd().solve() returns a minimum-norm solution; here the minimum-norm solution of the centred system)"""
    rng = np.random.default_rng(8)
    ang = rng.uniform(0, 2 * np.pi, 4000)
    # a circle with 1e-7 of out-of-plane noise: minimal fits pass the coplanarity check (ransac.h:225-234, 1e-8), the
    # inlier set is coplanar to 1e-7 -> the 3 x 3 normal matrix has a relative determinant of ~1e-14
    ring = np.c_[0.5 * np.cos(ang), 0.5 * np.sin(ang), 1e-7 * rng.standard_normal(len(ang))]
    xyz = np.r_[ring, rng.uniform(-1, 1, (50, 3)) * np.array([1, 1, 1e-7])]
    rc, model, inl, st = ctx.ransac_fit(capi.SPHERE, xyz, None, 0.01, 300, 1.0, seed=4)
    assert st["found"] == 1 and np.all(np.isfinite(model)), (st, model)
    d = np.abs(np.linalg.norm(ring - model[:3], axis=1) - model[3])
    assert np.median(d) < 0.05


# ---------------------------------------------------------------------------------------------------
# the chunked host-buffer fit (pinned caller buffers): the cloud is uploaded in chunks on a copy stream, each chunk is
# scored as it arrives, the minimal models come from a zero-copy gather of the sample points
def _pinned(a):
    import torch
    t = torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
    return t, t.numpy()


@pytest.mark.parametrize("kind", KINDS)
def test_chunked_pinned_fit_parity(ctx, capi, orc, kind):
    """>= 262144 points in pinned memory, probability 1: the overlapped-upload path == the oracle's sequential loop and
    == the same call on pageable memory (plain upload path)"""
    xyz, nrm = synth.make_c2(n=300_000, seed=41)
    nrm = nrm if kind == 2 else None
    keep, pxyz = _pinned(xyz)
    keepn, pnrm = _pinned(nrm) if nrm is not None else (None, None)
    for seed, H in ((3, 700), (4, 1500)):
        before = ctx.launches
        rc, model, inl, st = ctx.ransac_fit(kind, pxyz, pnrm, 0.01, H, 1.0, seed=seed, flags=capi.FLAG_CHUNKED_UPLOAD)
        n_launch = ctx.launches - before
        rc2, model2, inl2, st2 = ctx.ransac_fit(kind, xyz, nrm, 0.01, H, 1.0, seed=seed, flags=capi.FLAG_PLAIN_UPLOAD)
        assert n_launch > ctx.launches - before - n_launch                                     # chunk pipelines ran
        rc3, model3, inl3, st3 = ctx.ransac_fit(kind, pxyz, pnrm, 0.01, H, 1.0, seed=seed)     # default: one chunk
        assert rc == rc3 and np.array_equal(inl, inl3) and np.array_equal(model, model3) and st3["best_index"] == st["best_index"]
        assert rc == rc2 and np.array_equal(inl, inl2) and np.array_equal(model, model2)
        orc_rc, omodel, oinl, ost = orc.ransac_fit(kind, xyz, nrm, thr=0.01, max_it=H, prob=1.0, seed=seed)
        assert rc == orc_rc and np.array_equal(inl, oinl)
        for key in ("best_index", "best_count", "iterations_run", "stop_index", "found"):
            assert st[key] == ost[key] == st2[key], (key, st, ost)


def test_chunked_pinned_fit_falls_back_on_nonfinite_points(ctx, capi, orc):
    xyz = synth.make_c1(n=300_000, seed=42)
    xyz[123456, 1] = np.nan
    xyz[7] = np.inf
    keep, pxyz = _pinned(xyz)
    orc_rc, omodel, oinl, ost = orc.ransac_fit(orc.PLANE, xyz, thr=0.01, max_it=300, prob=1.0, seed=2)
    for flags in (0, capi.FLAG_CHUNKED_UPLOAD):
        rc, model, inl, st = ctx.ransac_fit(capi.PLANE, pxyz, None, 0.01, 300, 1.0, seed=2, flags=flags)
        assert rc == orc_rc and np.array_equal(inl, oinl) and st["best_index"] == ost["best_index"]


def test_registered_host_buffer_fit_parity(ctx, capi, orc):
    """M3D_FLAG_REGISTER_HOST: the caller's pageable cloud is page-locked in place on first use; same results, the
    buffer counts as pinned afterwards, and the registrations are dropped on request"""
    xyz, nrm = synth.make_c2(n=150000, seed=31)
    xyz = xyz.copy()
    for rep in range(2):
        for kind in KINDS:
            rc, model, inl, st = ctx.ransac_fit(kind, xyz, nrm if kind == 2 else None, 0.01, 3000, 1.0, seed=40 + kind,
                                                flags=capi.FLAG_REGISTER_HOST)
            orc_rc, omodel, oinl, ost = orc.ransac_fit(kind, xyz, nrm if kind == 2 else None, thr=0.01, max_it=3000,
                                                       prob=1.0, seed=40 + kind)
            assert rc == orc_rc and np.array_equal(inl, oinl)
            assert st["best_index"] == ost["best_index"] and st["best_count"] == ost["best_count"]
    ctx.host_unregister_all()
    rc, model, inl2, st = ctx.ransac_fit(capi.PLANE, xyz, None, 0.01, 3000, 1.0, seed=40)   # staged path again
    orc_rc, omodel, oinl, ost = orc.ransac_fit(capi.PLANE, xyz, None, thr=0.01, max_it=3000, prob=1.0, seed=40)
    assert np.array_equal(inl2, oinl)
