"""CPU test of the pybind11 drop-in surface (python/py_misc3d.cpp -> module `misc3d`): the names, argument
names and defaults of the reference's bindings (python/py_common.cpp:70-78, py_segmentation.cpp:87-96,
py_registration.cpp:55-106, py_misc3d.cpp:52-62).  No compute call: without a GPU the module must raise, not
fall back."""
import os
import re
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "python"))


@pytest.fixture(scope="module")
def m3d():
    return pytest.importorskip("misc3d")   # built by __graft_entry__.build() (python/Makefile)


def _sig(fn):
    """'name(arg: type = default, ...)' of the first overload -> [(arg, default or None)]"""
    doc = [ln for ln in fn.__doc__.splitlines() if re.match(r"^\s*(\d+\.\s*)?\w+\(", ln)][0]
    args = doc[doc.index("(") + 1: doc.rindex(") ->")]
    out, depth, cur = [], 0, ""
    for ch in args:
        depth += ch in "[(<"
        depth -= ch in "])>"
        if ch == "," and depth == 0:
            out.append(cur)
            cur = ""
        else:
            cur += ch
    out.append(cur)
    res = []
    for a in out:
        a = a.strip()
        if a == "*":
            continue
        name = a.split(":")[0].strip()
        res.append((name, a.rsplit("=", 1)[1].strip() if "=" in a else None))
    return res


def test_module_layout(m3d):
    for sub in ("common", "segmentation", "registration"):
        assert hasattr(m3d, sub)
    for fn in ("fit_plane", "fit_sphere", "fit_cylinder"):
        assert callable(getattr(m3d.common, fn))
    assert callable(m3d.segmentation.segment_plane_iterative)
    assert callable(m3d.registration.match_correspondence)
    assert callable(m3d.registration.compute_transformation_ransac)
    assert int(m3d.registration.MatchMethod.FLANN) == 0 and int(m3d.registration.MatchMethod.ANNOY) == 1
    lv = m3d.VerbosityLevel
    assert [int(lv.Error), int(lv.Warning), int(lv.Info), int(lv.Debug)] == [0, 1, 2, 3]
    old = m3d.get_verbosity_level()
    m3d.set_verbosity_level(lv.Error)
    assert m3d.get_verbosity_level() == lv.Error
    m3d.set_verbosity_level(old)


def test_argument_names_and_defaults(m3d):
    for fn in (m3d.common.fit_plane, m3d.common.fit_sphere, m3d.common.fit_cylinder):
        assert _sig(fn) == [("pc", None), ("threshold", "0.01"), ("max_iteration", "1000"), ("probability", "0.9999"),
                            ("seed", "None")]                      # py_common.cpp:70-78 (+ keyword-only seed)
    assert _sig(m3d.segmentation.segment_plane_iterative) == [
        ("pcd", None), ("threshold", None), ("max_iteration", "100"), ("min_ratio", "0.05"), ("seed", "None")]
    s = _sig(m3d.registration.match_correspondence)
    assert [a for a, _ in s] == ["src", "dst", "method", "n_trees"] and s[3][1] == "4" and "ANNOY" in s[2][1]
    assert _sig(m3d.registration.compute_transformation_ransac) == [
        ("src", None), ("dst", None), ("corres", None), ("threshold", "0.01"), ("max_iter", "100000"),
        ("edge_length_threshold", "0.9"), ("seed", "None")]
    # py_registration.cpp:12-31: (src, dst, scaling=False), point clouds or (n, 3) arrays
    assert _sig(m3d.registration.compute_transformation_least_square) == [("src", None), ("dst", None), ("scaling", "False")]
    assert _sig(m3d.registration.compute_fpfh_feature_device) == [("pcd", None), ("radius", None), ("max_nn", "100")]
    assert hasattr(m3d.registration, "DeviceFeature")
    assert _sig(m3d.registration.refine_transformation_on_inliers) == [
        ("src", None), ("dst", None), ("corres", None), ("T", None), ("threshold", "0.01"), ("scaling", "False")]


def test_no_cpu_fallback(m3d):
    from misc3d_b200 import capi
    if capi.device_count() > 0:
        pytest.skip("GPU present")
    pts = np.random.default_rng(0).uniform(-1, 1, (100, 3))
    with pytest.raises(RuntimeError):
        m3d.common.fit_plane(pts, 0.01, 10, 0.99)
