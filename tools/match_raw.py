"""timing-only run of m3d_nearest at C4 size (results are not checked; used for the tensor-core timing experiments recorded in DESIGN.md section 8)"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from misc3d_b200 import capi, synth
d = synth.make_c4()
ctx = capi.Context(0)
for rep in range(3):
    t0 = time.perf_counter()
    nn, ms = ctx.nearest(d["src_feat"], d["dst_feat"])
    print("nearest device_ms", round(ms, 2), "wall", round(1e3 * (time.perf_counter() - t0), 2), flush=True)
