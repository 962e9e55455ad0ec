"""ncu target.  `ransac` (default): one warm + one measured RANSAC fit per primitive on the C2 cloud
(10k hypotheses, 1M points).  `c4`: match_correspondence + compute_transformation_ransac at C4 size.  `feat`: FPFH + ICP at 200k points."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from misc3d_b200 import capi, synth  # noqa: E402

what = sys.argv[1] if len(sys.argv) > 1 else "ransac"
ctx = capi.Context(0)
if what == "ransac":
    xyz, nrm = synth.make_c2()
    cloud = ctx.upload(xyz, nrm)
    for kind in (0, 1, 2):
        for rep in range(2):
            rc, model, inl, st = ctx.ransac_fit_cloud(kind, cloud, 0.01, 10000, 1.0, seed=rep)
            print(kind, rep, st["score_ms"], st["device_ms"], st["exact_resolves"])
elif what == "stats":  # work counters of the scoring kernel (statistics build, M3D_FLAG_STATS)
    import json
    xyz, nrm = synth.make_c2()
    cloud = ctx.upload(xyz, nrm)
    for kind in (0, 1, 2):
        rc, model, inl, st = ctx.ransac_fit_cloud(kind, cloud, 0.01, 10000, 1.0, seed=1, flags=capi.FLAG_STATS)
        s = ctx.score_stats()
        s["pairs_evaluated_frac"] = s["cell_pairs"] * 32 / (1e6 * 1e4)
        s["lane_slots"] = 32 * s["passes_1"] + 64 * s["passes_2"]
        s["lane_efficiency"] = s["cell_pairs"] / max(s["lane_slots"], 1)
        print(kind, st["score_ms"], json.dumps(s))
elif what == "feat":  # FPFH (f3) + ICP (f4) at 200k points
    import numpy as np
    dp = synth.make_surface_pair(n=200000, seed=2, sigma=0.0005)
    for rep in range(2):
        f, ms = ctx.compute_fpfh(dp["src"], dp["src_nrm"], 0.03, 100)
        print("fpfh", rep, ms)
    if len(sys.argv) > 2 and sys.argv[2] == "chain":   # FPFH x2 -> match on the device (real descriptors)
        fa, _ = ctx.fpfh_features(dp["src"], dp["src_nrm"], 0.03, 100)
        fb, _ = ctx.fpfh_features(dp["dst"], dp["dst_nrm"], 0.03, 100)
        i0, i1, ms = ctx.match_features(fa, fb)
        print("chain match", ms, len(i0), float(np.mean(dp["perm"][i1.astype(np.int64)] == i0.astype(np.int64))) if len(i0) else 0)
        sys.exit(0)
    T0 = dp["T_true"].copy()
    T0[:3, 3] += 0.01
    for rep in range(2):
        T, fit, rmse, it = ctx.icp_point_to_point(dp["src"], dp["dst"], 0.02, T0, 30)
        print("icp", rep, it, fit, rmse)
else:
    d = synth.make_c4()
    for rep in range(3):
        i0, i1, ms = ctx.match_correspondence(d["src_feat"], d["dst_feat"])
        print("match", rep, ms, len(i0))
    rc, T, st = ctx.ransac_registration(d["src"], d["dst"], i0, i1, 0.02, 50000, 0.9, 1.0, 1)
    print("reg", st)
