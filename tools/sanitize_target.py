"""small end-to-end run of every new kernel family for compute-sanitizer (memcheck / racecheck / synccheck)"""
import sys
sys.path.insert(0, ".")
import numpy as np
from misc3d_b200 import capi, synth
ctx = capi.Context(0)
xyz, nrm = synth.make_c2(n=40000, seed=3)
cloud = ctx.upload(xyz, nrm)
for kind in (0, 1, 2):
    for H, p in ((1500, 1.0), (5000, 1.0), (600, 0.9999)):
        rc, m, inl, st = ctx.ransac_fit_cloud(kind, cloud, 0.01, H, p, seed=kind + 1)
        print("fit", kind, H, p, rc, len(inl), st["best_index"])
rc, m, inl, st = ctx.ransac_fit(0, xyz, None, 0.01, 1200, 1.0, seed=5)
print("host fit", rc, len(inl))
t = ctx.sample_table_device(3, 5000, 4, 3000)
print("table", None if t is None else t.shape)
t = ctx.sample_table_device(5, 1_000_000, 3, 40_000)   # long table: jump-ahead segments (mt_jump / mt_segments kernels)
print("long table", None if t is None else t.shape, int(t.sum()) if t is not None else 0)
rc, planes, labels, ms = ctx.segment_plane_iterative(synth.make_c3(30000, 5), 0.01, 100, 0.1, seed=2, labels32=True)
print("seg", rc, len(planes))
d = synth.make_surface_pair(n=3000, seed=2)
f, _ = ctx.compute_fpfh(d["src"], d["src_nrm"], 0.15, 40)
g, _ = ctx.compute_fpfh(d["dst"], d["dst_nrm"], 0.15, 40)
i0, i1, _ = ctx.match_correspondence(f, g)
rc, T, st = ctx.ransac_registration(d["src"], d["dst"], i0, i1, 0.02, 2000, 0.9, 0.999, 1)
T2, fit, rmse, it = ctx.icp_point_to_point(d["src"], d["dst"], 0.05, T, 10)
print("chain", len(i0), rc, fit, it)
# device-resident descriptors, the refit extension, and the full fp64 search (16 rows per group) on every row
fa, _ = ctx.fpfh_features(d["src"], d["src_nrm"], 0.15, 40)
fb, _ = ctx.fpfh_features(d["dst"], d["dst_nrm"], 0.15, 40)
j0, j1, _ = ctx.match_features(fa, fb)
T3, n_in = ctx.registration_refit(d["src"], d["dst"], i0, i1, T, 0.02)
import numpy as np   # noqa: E402
wide = np.asfortranarray(np.random.default_rng(1).uniform(0, 1, (200, 300)))   # dim 200: the fp64-only path
k0, k1, _ = ctx.match_correspondence(wide, wide[:, ::-1].copy(order="F"))
print("device chain", len(j0), bool((j0 == i0).all()), n_in, len(k0))
fa.free()
fb.free()
idx, dist, cnt = ctx.knn_search(f[:, :500], f[:, :20], 5)
print("knn", cnt[:5])
ctx.close()
print("SANITIZE_TARGET_DONE")
