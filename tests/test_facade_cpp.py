"""The C++ drop-in boundary: a caller written against the reference's class API (tests/cpp/facade_example.cpp,
in the shape of the reference's examples/cpp/*.cpp) compiles and links against include/misc3d/** +
libm3d_b200.so (CPU test), and on the GPU returns what the compiled reference returns (gpu test)."""
import os
import shutil
import subprocess

import numpy as np
import pytest

from misc3d_b200 import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "cpp", "facade_example.cpp")
LIBDIR = os.path.join(ROOT, "misc3d_b200")


def _build(tmp):
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else shutil.which("g++")
    exe = os.path.join(tmp, "facade_example")
    cuda_inc = "/usr/local/cuda/include"
    cmd = [cxx, "-std=c++17", "-O2", "-I", os.path.join(ROOT, "include"), "-I", cuda_inc, SRC, "-o", exe,
           "-L", LIBDIR, "-lm3d_b200", f"-Wl,-rpath,{LIBDIR}", "-L/usr/local/cuda/lib64", "-Wl,-rpath,/usr/local/cuda/lib64"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-2000:]
    return exe


def test_reference_style_caller_compiles_and_links(tmp_path, capi):
    assert os.path.exists(_build(str(tmp_path)))


@pytest.mark.gpu
def test_reference_style_caller_matches_compiled_reference(tmp_path, capi):
    import refc
    if not refc.build():
        pytest.skip("oracle/_ref is not built and /root/reference is absent")
    exe = _build(str(tmp_path))
    xyz = synth.make_c3(30000, 5)
    path = os.path.join(str(tmp_path), "cloud.bin")
    with open(path, "wb") as f:
        f.write(np.uint64(len(xyz)).tobytes())
        f.write(np.ascontiguousarray(xyz, dtype=np.float64).tobytes())
    seed = 9
    r = subprocess.run([exe, path, str(seed)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln.split() for ln in r.stdout.strip().splitlines()]
    fit = next(ln for ln in lines if ln[0] == "fit")
    rc, model, inl, st = refc.ransac_fit(0, xyz, None, 0.01, 100, 0.9999, seed)
    assert int(fit[1]) == rc and int(fit[2]) == len(inl)
    np.testing.assert_allclose([float(v) for v in fit[3:7]], model, rtol=1e-9, atol=1e-12)
    h = 1469598103934665603
    for i in inl.tolist():
        h = ((h ^ i) * 1099511628211) % (1 << 64)
    assert int(next(ln for ln in lines if ln[0] == "hash")[1]) == h          # identical inlier index list
    npl, planes, labels = refc.segment_plane_iterative(xyz, 0.01, 100, 0.1, seed)
    assert int(next(ln for ln in lines if ln[0] == "planes")[1]) == npl
    got = [ln for ln in lines if ln[0] == "plane"]
    for k, ln in enumerate(got):
        assert int(ln[1]) == int((labels == k).sum())
        np.testing.assert_allclose([float(v) for v in ln[2:6]], planes[k], rtol=1e-9, atol=1e-12)
    assert next(ln for ln in lines if ln[0] == "throw")[1] == "1"
    knn = next(ln for ln in lines if ln[0] == "knn")
    d = np.sqrt(((xyz - xyz[0]) ** 2).sum(1))
    order = np.lexsort((np.arange(len(xyz)), d))[:5]
    assert int(knn[1]) == 5 and [int(v) for v in knn[2::2]] == order.tolist()
    np.testing.assert_allclose([float(v) for v in knn[3::2]], d[order], rtol=1e-12, atol=1e-15)
    assert next(ln for ln in lines if ln[0] == "hybrid")[1] == "4"
    dm = next(ln for ln in lines if ln[0] == "devmatch")
    assert dm[1] == "1" and int(dm[2]) > 1000 and dm[3] == "1"   # device-resident descriptors: same matches, same bits
