"""Component timing of the registration front end at 200k points: FPFH of both clouds + match_correspondence, with the
descriptors left on the device (m3d_fpfh_create / m3d_match_features) and through host buffers.  Wall ms (device ms)."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from misc3d_b200 import capi, synth
ctx = capi.Context(0)
dp = synth.make_surface_pair(n=200000, seed=2, sigma=0.0005)
for rep in range(4):
    t0 = time.perf_counter(); fa, ma = ctx.fpfh_features(dp["src"], dp["src_nrm"], 0.03, 100)
    t1 = time.perf_counter(); fb, mb = ctx.fpfh_features(dp["dst"], dp["dst_nrm"], 0.03, 100)
    t2 = time.perf_counter(); i0, i1, ms = ctx.match_features(fa, fb)
    t3 = time.perf_counter(); fa.free(); fb.free()
    t4 = time.perf_counter()
    print(f"device: fpfh {1e3*(t1-t0):.1f} ({ma:.1f}) {1e3*(t2-t1):.1f} ({mb:.1f}) match {1e3*(t3-t2):.1f} ({ms:.1f}) free {1e3*(t4-t3):.1f}")
for rep in range(3):
    t0 = time.perf_counter(); fa, ma = ctx.compute_fpfh(dp["src"], dp["src_nrm"], 0.03, 100)
    t1 = time.perf_counter(); fb, mb = ctx.compute_fpfh(dp["dst"], dp["dst_nrm"], 0.03, 100)
    t2 = time.perf_counter(); i0, i1, ms = ctx.match_correspondence(fa, fb)
    t3 = time.perf_counter()
    print(f"host: fpfh {1e3*(t1-t0):.1f} ({ma:.1f}) {1e3*(t2-t1):.1f} ({mb:.1f}) match {1e3*(t3-t2):.1f} ({ms:.1f})")
