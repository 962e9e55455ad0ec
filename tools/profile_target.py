"""ncu target: one warm + one measured launch of the scoring kernel per primitive on the C2 cloud
(10k hypotheses, 1M points), plus one full RANSAC fit each so that the refine passes appear."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from misc3d_b200 import capi, synth  # noqa: E402

xyz, nrm = synth.make_c2()
ctx = capi.Context(0)
cloud = ctx.upload(xyz, nrm)
for kind in (0, 1, 2):
    for rep in range(2):
        rc, model, inl, st = ctx.ransac_fit_cloud(kind, cloud, 0.01, 10000, 1.0, seed=rep)
        print(kind, rep, st["score_ms"], st["device_ms"], st["exact_resolves"])
