"""draw_ms (device-side sample draw inside a fit, CUDA events, warm clocks) for 10k and 80k hypotheses"""
import sys
sys.path.insert(0, ".")
import numpy as np
from misc3d_b200 import capi, synth
xyz, nrm = synth.make_c2()
c = capi.Context(0)
cloud = c.upload(xyz, nrm)
for H in (10000, 80000):
    for kind in (0, 1, 2):
        d = []
        for s in range(6):
            rc, m, inl, st = c.ransac_fit_cloud(kind, cloud, 0.01, H, 1.0, seed=s, want_inliers=False)
            d.append(st["draw_ms"])
        print(f"H={H} kind={kind} draw_ms {np.mean(d[1:]):.4f}  fit_ms {st['device_ms']:.3f} score_ms {st['score_ms']:.3f}")
