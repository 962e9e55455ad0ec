"""GPU parity against the REFERENCE'S OWN code: libm3d_b200.so (through the C-ABI) versus the reference
sources compiled unmodified into oracle/_ref/ (oracle/refc.py; Eigen/Open3D stood in by oracle/shim/,
seed injected, sequential build).  Same seeded inputs; bit-exact inlier index sets, iteration counts,
minimal models, labels and match lists; the least-squares refits within the stated tolerance.
The prebuilt oracle/_ref/*.so travel with the tree to the GPU box (no /root/reference there)."""
import numpy as np
import pytest

from misc3d_b200 import synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def refc():
    import refc as _refc
    if not _refc.build():
        pytest.skip("oracle/_ref is not built and /root/reference is absent")
    return _refc


def _fit_vs_reference(ctx, refc, kind, xyz, nrm, thr, max_it, prob, seed):
    rc, model, inl, st = ctx.ransac_fit(kind, xyz, nrm, thr, max_it, prob, seed)
    r_rc, r_model, r_inl, r_st = refc.ransac_fit(kind, xyz, nrm, thr, max_it, prob, seed)
    assert rc == r_rc
    np.testing.assert_array_equal(inl, r_inl)                      # bit-exact inlier index set
    assert st["iterations_run"] == r_st["iterations_run"]          # `count` of ransac.h:616-619
    assert abs(st["best_count"] / len(xyz) - r_st["fitness"]) < 1e-12
    if kind == 2:   # GeneralFit is a no-op for the cylinder: the minimal model itself, bit-exact
        np.testing.assert_array_equal(model.view(np.uint64), r_model.view(np.uint64))
    else:           # least-squares refit: parallel summation order on the GPU -> 1e-9 rel / 1e-12 abs
        np.testing.assert_allclose(model, r_model, rtol=1e-9, atol=1e-12)
    return st


def test_c1_fit_plane_vs_reference(ctx, refc):
    """BASELINE config C1: fit_plane on the 50k-point plane+noise cloud, 100 iterations, seeded."""
    xyz = synth.make_c1()
    for seed in (1, 2, 3):
        st = _fit_vs_reference(ctx, refc, 0, xyz, None, 0.01, 100, 0.9999, seed)
        assert st["iterations_run"] < 100   # the adaptive exit fired, at the reference's iteration


@pytest.mark.parametrize("kind", [0, 1, 2])
@pytest.mark.parametrize("prob", [0.9999, 1.0])
def test_c2_small_fits_vs_reference(ctx, refc, kind, prob):
    xyz, nrm = synth.make_c2(40000, 31)
    for seed in (5, 6):
        _fit_vs_reference(ctx, refc, kind, xyz, nrm if kind == 2 else None, 0.01, 400, prob, seed)


def test_segmentation_vs_reference(ctx, refc):
    xyz = synth.make_c3(40000, 9)
    rc, planes, labels, _ = ctx.segment_plane_iterative(xyz, 0.01, 100, 0.05, seed=3)
    npl, r_planes, r_labels = refc.segment_plane_iterative(xyz, 0.01, 100, 0.05, 3)
    assert rc == 0 and len(planes) == npl
    np.testing.assert_array_equal(labels, r_labels)
    np.testing.assert_allclose(planes, r_planes, rtol=1e-9, atol=1e-12)


@pytest.mark.parametrize("method", [0, 1])
def test_matching_vs_reference_flann(ctx, refc, method):
    """both MatchMethod values of the product run the exact search = the reference's FLANN branch"""
    d = synth.make_c4(n=4000, seed=12)
    i0, i1, _ = ctx.match_correspondence(d["src_feat"], d["dst_feat"], method=method)
    r0, r1 = refc.match_correspondence(d["src_feat"], d["dst_feat"], refc.FLANN)
    np.testing.assert_array_equal(i0, r0)
    np.testing.assert_array_equal(i1, r1)


def test_real_scan_golden(ctx):
    """real sensor data: every 4th vertex of the reference's demo scan (tests/golden/real_scan.npz holds the
    input and the compiled reference's outputs for the demo's parameters, tools/make_golden.py)"""
    import os
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "real_scan.npz"))
    xyz = g["xyz"]
    rc, model, inl, st = ctx.ransac_fit(0, xyz, None, 0.01, 100, 0.9999, 1)
    assert rc == int(g["rc"]) and st["iterations_run"] == int(g["iterations_run"])
    assert st["best_index"] == int(g["best_index"]) and st["best_count"] == int(g["best_count"])
    np.testing.assert_array_equal(inl, g["inl"])
    np.testing.assert_allclose(model, g["model"], rtol=1e-9, atol=1e-12)
    rc, planes, labels, _ = ctx.segment_plane_iterative(xyz, 0.01, 100, 0.1, seed=3)
    assert rc == 0 and len(planes) == len(g["planes"])
    np.testing.assert_array_equal(labels.astype(np.int64), g["labels"])
    np.testing.assert_allclose(planes, g["planes"], rtol=1e-9, atol=1e-12)
