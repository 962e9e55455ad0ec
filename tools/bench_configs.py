"""Times every BASELINE.json config (C1..C4; C5 at the single-GPU share) on the GPU through the
C-ABI and, on a bounded sample, with the CPU oracle -- evidence for the SURVEY.md §8 rows that are
not the bench.py headline.  Prints one JSON object per config."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import orc  # noqa: E402
from misc3d_b200 import capi, synth  # noqa: E402

orc.build()
ctx = capi.Context(0)
cores = orc.omp_threads()


def timed(fn, reps=3):
    best = 1e30
    out = None
    for _ in range(reps):
        t0 = time.perf_counter()
        out = fn()
        best = min(best, time.perf_counter() - t0)
    return best, out


def emit(**kw):
    print(json.dumps(kw), flush=True)


# C1 ---------------------------------------------------------------------------------------------
xyz = synth.make_c1()
t, (rc, model, inl, st) = timed(lambda: ctx.ransac_fit(capi.PLANE, xyz, None, 0.01, 100, 0.9999, 1))
tc, (orc_rc, omodel, oinl, ost) = timed(lambda: orc.ransac_fit(orc.PLANE, xyz, thr=0.01, max_it=100, prob=0.9999, seed=1,
                                                              omp=True, faithful=True), 1)
emit(config="C1 fit_plane 50k pts / 100 it (e2e host buffers)", gpu_ms=1e3 * t, device_ms=st["device_ms"],
     cpu_omp_ms=1e3 * tc, cores=cores, inliers_equal=bool(np.array_equal(
         inl, orc.ransac_fit(orc.PLANE, xyz, thr=0.01, max_it=100, prob=0.9999, seed=1)[2])))

# C3 ---------------------------------------------------------------------------------------------
xyz = synth.make_c3()
t, (rc, planes, labels, ms) = timed(lambda: ctx.segment_plane_iterative(xyz, 0.01, 100, 0.05, seed=1))
tc, (orc_rc, oplanes, olabels) = timed(lambda: orc.segment_plane_iterative(xyz, 0.01, 100, 0.05, seed=1, omp=True), 1)
t32, _ = timed(lambda: ctx.segment_plane_iterative(xyz, 0.01, 100, 0.05, seed=1, labels32=True))
emit(config="C3 segment_plane_iterative 2M pts, 6-plane scene, 100 it/round (e2e host buffers)", gpu_ms=1e3 * t,
     gpu_ms_labels32=1e3 * t32,
     device_fit_ms=ms, planes=int(len(planes)), cpu_omp_ms=1e3 * tc, cores=cores, rc=rc,
     note="CPU = oracle OpenMP rounds (not seed-comparable: shared sampler order)")

# C4 ---------------------------------------------------------------------------------------------
d = synth.make_c4()
t, (i0, i1, ms) = timed(lambda: ctx.match_correspondence(d["src_feat"], d["dst_feat"]), 2)
sub = 2000  # CPU brute force on a bounded sample of query rows, both directions
tc, _ = timed(lambda: (orc.nearest(d["src_feat"][:, :sub], d["dst_feat"]), orc.nearest(d["dst_feat"][:, :sub], d["src_feat"])), 1)
emit(config="C4a match_correspondence 200k x 200k x 33-D (e2e host buffers)", gpu_ms=1e3 * t, device_ms=ms,
     matches=int(len(i0)), descriptor_pairs_per_s=2 * 4e10 / t,
     cpu_omp_ms_extrapolated=1e3 * tc * 200000 / sub, cpu_sample=f"{sub} query rows per direction, exact brute force, {cores} threads")
t, (rc, T, st) = timed(lambda: ctx.ransac_registration(d["src"], d["dst"], i0, i1, 0.02, 50000, 0.9, 1.0, 1), 2)
hc = 500
tc, _ = timed(lambda: orc.ransac_registration(d["src"], d["dst"], i0, i1, thr=0.02, max_iter=hc, edge_thr=0.9,
                                              confidence=1.0, seed=1, omp=True), 1)
emit(config="C4b compute_transformation_ransac 50k hypotheses, confidence 1.0 (e2e host buffers)", gpu_ms=1e3 * t,
     device_ms=st["device_ms"], score_ms=st["score_ms"], evaluated=st["evaluated"], correspondences=int(len(i0)),
     hyp_per_s=50000 / t, cpu_omp_hyp_per_s=hc / tc, cpu_sample=f"{hc} hypotheses, {cores} threads",
     err_vs_truth=float(np.linalg.norm(T - d["T_true"])))
t, (rc, T, st) = timed(lambda: ctx.ransac_registration(d["src"], d["dst"], i0, i1, 0.02, 50000, 0.9, 0.999, 1), 2)
emit(config="C4c compute_transformation_ransac default confidence 0.999 (early exit)", gpu_ms=1e3 * t,
     stop_index=st["stop_index"], evaluated=st["evaluated"])

# the Open3D steps around C4 (f3 / f4): FPFH of both clouds and the ICP refinement, 200k points each ----------------
dp = synth.make_surface_pair(n=200000, seed=2, sigma=0.0005)
t, (f, ms) = timed(lambda: ctx.compute_fpfh(dp["src"], dp["src_nrm"], 0.03, 100), 2)
emit(config="f3 compute_fpfh 200k pts, radius 0.03, max_nn 100 (e2e host buffers: 4.8 MB x2 up, 52.8 MB down)", gpu_ms=1e3 * t,
     device_ms=ms)
# the chain FPFH(src) + FPFH(dst) + match_correspondence: through host buffers (descriptors come back and go up again)
# and with the descriptors left on the device (m3d_fpfh_create / m3d_match_features)
def chain_host():
    fa, _ = ctx.compute_fpfh(dp["src"], dp["src_nrm"], 0.03, 100)
    fb, _ = ctx.compute_fpfh(dp["dst"], dp["dst_nrm"], 0.03, 100)
    return ctx.match_correspondence(fa, fb)


def chain_device():
    fa, _ = ctx.fpfh_features(dp["src"], dp["src_nrm"], 0.03, 100)
    fb, _ = ctx.fpfh_features(dp["dst"], dp["dst_nrm"], 0.03, 100)
    out = ctx.match_features(fa, fb)
    fa.free()
    fb.free()
    return out


th, (h0, h1, _) = timed(chain_host, 3)
td, (d0, d1, _) = timed(chain_device, 3)
emit(config="f3 chain: FPFH x2 + match_correspondence, 200k points each", gpu_ms_host_round_trip=1e3 * th,
     gpu_ms_device_resident=1e3 * td, gpu_ms=1e3 * td, matches=int(len(d0)), identical=bool(np.array_equal(h0, d0) and np.array_equal(h1, d1)))
T0 = dp["T_true"].copy()
T0[:3, 3] += 0.01
t, (T, fit, rmse, it) = timed(lambda: ctx.icp_point_to_point(dp["src"], dp["dst"], 0.02, T0, 30), 2)
emit(config="f4 icp_point_to_point 200k <-> 200k pts, max_distance 0.02 (e2e host buffers)", gpu_ms=1e3 * t, iterations=it,
     fitness=fit, inlier_rmse=rmse, err_vs_truth=float(np.linalg.norm(T - dp["T_true"])))

# C5 (single GPU share shown for 1 GPU: all 100k hypotheses) -------------------------------------
xyz = synth.make_c5()
cloud = ctx.upload(xyz)
t, (rc, model, inl, st) = timed(lambda: ctx.ransac_fit_cloud(capi.PLANE, cloud, 0.01, 100000, 1.0, seed=1,
                                                             want_inliers=False), 2)
emit(config="C5 fit_plane 4M pts x 100k hypotheses on ONE GPU (resident cloud)", gpu_ms=1e3 * t,
     score_ms=st["score_ms"], hyp_per_s=100000 / t, point_hyp_per_s=4e6 * 100000 / t)
