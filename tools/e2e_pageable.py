"""pageable-buffer e2e of the three C2 fits (what bench.py reports as e2e_pageable), for upload tuning"""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from misc3d_b200 import capi, synth
xyz, nrm = synth.make_c2()
ctx = capi.Context(0)
buf = np.empty(len(xyz), dtype=np.uint64)
def step(seed):
    for kind in (0, 1, 2):
        ctx.ransac_fit(kind, xyz, nrm if kind == 2 else None, 0.01, 10000, 1.0, seed=seed + kind, inl_buf=buf)
for w in range(3):
    step(w)
t0 = time.perf_counter()
for s in range(10):
    step(100 + 3 * s)
dt = (time.perf_counter() - t0) / 10
print(f"pageable e2e: {dt*1e3:.3f} ms/step  {3e4/dt/1e6:.3f} M hyp/s  threads={os.environ.get('M3D_UPLOAD_THREADS','3')} staged={os.environ.get('M3D_STAGED_UPLOAD','1')}")
