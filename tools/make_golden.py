"""Generate tests/golden/*.npz from the CPU oracle.

The reference ships no tests, golden vectors or fixtures with expected outputs for this path
 and cannot be built here (Open3D/Eigen absent), so these vectors pin the ORACLE
(parity unpinned against the reference, see oracle/m3d_oracle.h); they guard the oracle against
regressions and give the GPU tests fixed expected outputs that travel to the GPU box.
Run:  python tools/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import orc  # noqa: E402
from misc3d_b200 import synth  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
os.makedirs(OUT, exist_ok=True)
orc.build()


def save_fit(name, rc, model, inl, st):
    np.savez_compressed(os.path.join(OUT, name + ".npz"), rc=rc, model=model, inl=inl.astype(np.uint32),
                        best_index=st["best_index"], best_count=st["best_count"], best_rmse=st["best_rmse"],
                        iterations_run=st["iterations_run"], stop_index=st["stop_index"])


xyz = synth.make_c1()
save_fit("c1_plane", *orc.ransac_fit(orc.PLANE, xyz, thr=0.01, max_it=100, prob=0.9999, seed=1))
xyz, nrm = synth.make_c2(n=20000, seed=11)
save_fit("small_sphere", *orc.ransac_fit(orc.SPHERE, xyz, thr=0.01, max_it=300, prob=0.9999, seed=2))
save_fit("small_cylinder", *orc.ransac_fit(orc.CYLINDER, xyz, nrm, thr=0.01, max_it=300, prob=0.9999, seed=3))
xyz = synth.make_c3(n=30000, seed=4)
rc, planes, labels = orc.segment_plane_iterative(xyz, 0.01, 100, 0.05, seed=7)
np.savez_compressed(os.path.join(OUT, "seg_small.npz"), rc=rc, planes=planes, labels=labels)
d = synth.make_c4(n=3000, seed=5)
i0, i1 = orc.match_correspondence(d["src_feat"], d["dst_feat"])
rc, T, st = orc.ransac_registration(d["src"], d["dst"], i0, i1, thr=0.02, max_iter=2000, edge_thr=0.9,
                                    confidence=0.999, seed=1)
np.savez_compressed(os.path.join(OUT, "reg_small.npz"), i0=i0, i1=i1, T=T, best_index=st["best_index"],
                    best_count=st["best_count"], best_rmse=st["best_rmse"], evaluated=st["evaluated"],
                    stop_index=st["stop_index"])
print("golden vectors written to", OUT)
