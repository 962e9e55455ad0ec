"""Generate tests/golden/*.npz.

The reference ships no tests, golden vectors or fixtures with expected outputs for this path
(SURVEY.md §4).  The RANSAC-fit, segmentation and matching vectors are therefore OUTPUTS OF THE
REFERENCE ITSELF RUN HERE: its ransac.h / iterative_plane_segmentation.cpp / correspondence_matching.cpp
compiled unmodified into oracle/_ref (oracle/refc.py; Eigen/Open3D stood in by oracle/shim/, seed
injected, sequential build); the restated oracle must reproduce them (asserted below, and in
tests/test_oracle.py on every run).  Loop statistics the reference does not expose (best index, stop
index, rmse) come from the oracle.  The registration vector is the oracle's alone: that arithmetic is
Open3D v0.15.1's RegistrationRANSACBasedOnCorrespondence, which is not under /root/reference.
Run (where /root/reference exists):  python tools/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import orc  # noqa: E402
import refc  # noqa: E402
from misc3d_b200 import synth  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
os.makedirs(OUT, exist_ok=True)
orc.build()
if not refc.build():
    raise SystemExit("needs /root/reference (or a prebuilt oracle/_ref)")


def ref_fit(kind, xyz, nrm, **kw):
    """the reference's FitModel; the oracle must agree bit for bit on what both expose"""
    r_rc, r_model, r_inl, r_st = refc.ransac_fit(kind, xyz, nrm, kw["thr"], kw["max_it"], kw["prob"], kw["seed"])
    rc, model, inl, st = orc.ransac_fit(kind, xyz, nrm, **kw)
    assert rc == r_rc and np.array_equal(inl, r_inl) and st["iterations_run"] == r_st["iterations_run"]
    assert np.allclose(model, r_model, rtol=1e-9, atol=1e-12)
    return r_rc, r_model, r_inl, st


def save_fit(name, rc, model, inl, st):
    np.savez_compressed(os.path.join(OUT, name + ".npz"), rc=rc, model=model, inl=inl.astype(np.uint32),
                        best_index=st["best_index"], best_count=st["best_count"], best_rmse=st["best_rmse"],
                        iterations_run=st["iterations_run"], stop_index=st["stop_index"])


xyz = synth.make_c1()
save_fit("c1_plane", *ref_fit(orc.PLANE, xyz, None, thr=0.01, max_it=100, prob=0.9999, seed=1))
xyz, nrm = synth.make_c2(n=20000, seed=11)
save_fit("small_sphere", *ref_fit(orc.SPHERE, xyz, None, thr=0.01, max_it=300, prob=0.9999, seed=2))
save_fit("small_cylinder", *ref_fit(orc.CYLINDER, xyz, nrm, thr=0.01, max_it=300, prob=0.9999, seed=3))
xyz = synth.make_c3(n=30000, seed=4)
npl, planes, labels = refc.segment_plane_iterative(xyz, 0.01, 100, 0.05, 7)
rc, oplanes, olabels = orc.segment_plane_iterative(xyz, 0.01, 100, 0.05, seed=7)
assert rc == 0 and np.array_equal(labels, olabels) and np.array_equal(planes, oplanes)
np.savez_compressed(os.path.join(OUT, "seg_small.npz"), rc=rc, planes=planes, labels=labels)
d = synth.make_c4(n=3000, seed=5)
i0, i1 = refc.match_correspondence(d["src_feat"], d["dst_feat"], refc.FLANN)
o0, o1 = orc.match_correspondence(d["src_feat"], d["dst_feat"])
assert np.array_equal(i0, o0) and np.array_equal(i1, o1)
rc, T, st = orc.ransac_registration(d["src"], d["dst"], i0, i1, thr=0.02, max_iter=2000, edge_thr=0.9,
                                    confidence=0.999, seed=1)
np.savez_compressed(os.path.join(OUT, "reg_small.npz"), i0=i0, i1=i1, T=T, best_index=st["best_index"],
                    best_count=st["best_count"], best_rmse=st["best_rmse"], evaluated=st["evaluated"],
                    stop_index=st["stop_index"])
# real sensor data: every 4th vertex of the reference's demo scan (examples/data/segmentation/test.ply, binary PLY,
# double xyz) with the demo's parameters (examples/cpp/segment_plane_iterative.cpp:18); input + the compiled
# reference's outputs travel as one small fixture
ply = os.path.join(refc.REFERENCE_ROOT, "examples", "data", "segmentation", "test.ply")
if os.path.exists(ply):
    raw = open(ply, "rb").read()
    end = raw.index(b"end_header\n") + len(b"end_header\n")
    nv = int([ln for ln in raw[:end].decode().splitlines() if ln.startswith("element vertex")][0].split()[-1])
    scan = np.frombuffer(raw, dtype="<f8", count=3 * nv, offset=end).reshape(nv, 3)[::4].copy()
    r_rc, r_model, r_inl, st = ref_fit(orc.PLANE, scan, None, thr=0.01, max_it=100, prob=0.9999, seed=1)
    npl, planes, labels = refc.segment_plane_iterative(scan, 0.01, 100, 0.1, 3)
    rc, oplanes, olabels = orc.segment_plane_iterative(scan, 0.01, 100, 0.1, seed=3)
    assert rc == 0 and np.array_equal(labels, olabels) and np.array_equal(planes, oplanes)
    np.savez_compressed(os.path.join(OUT, "real_scan.npz"), xyz=scan.astype(np.float64), rc=r_rc, model=r_model,
                        inl=r_inl.astype(np.uint32), iterations_run=st["iterations_run"],
                        best_index=st["best_index"], best_count=st["best_count"], planes=planes,
                        labels=labels.astype(np.int64))
print("golden vectors written to", OUT)
