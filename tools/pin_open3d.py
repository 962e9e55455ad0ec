"""Pins rows a19 / a20 against a REAL Open3D the moment one is importable (there is none in this image and no
network: `import open3d` fails here).  Run on any box that has `pip install open3d`:

    python tools/pin_open3d.py            # writes tests/golden/reg_open3d.npz

It runs exactly what RANSACSolver::Solve runs (src/transform_estimation.cpp:124-164): registration_ransac_based_on_
correspondence with TransformationEstimationPointToPoint(False), ransac_n = 3, the EdgeLength(0.9) and Distance(thr)
checkers and RANSACConvergenceCriteria(max_iter, 0.999), plus Eigen::umeyama through
TransformationEstimationPointToPoint.compute_transformation (with_scaling False / True) -- on the synthetic C4 recipe at
3000 points -- and records inputs and outputs.  tests/test_oracle.py::test_open3d_pin and
tests/test_gpu_registration.py::test_open3d_pin consume the file (skipped while it is absent).

Open3D's RANSAC draws from its own global random engine, so the sample stream cannot be matched; what IS comparable:
the transform estimated from a GIVEN correspondence triple / set (umeyama, bit-level up to the SVD), the checkers'
accept / reject decisions for given triples, the fitness / inlier_rmse of a given transform, and the final result on data
with one dominant alignment (same inlier set => same fitness, transforms within the 3-point noise)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))


def main():
    try:
        import open3d as o3d
    except Exception as e:  # pragma: no cover - depends on the box
        print(f"open3d is not importable here ({e}); nothing written")
        return 1
    import orc
    from misc3d_b200 import synth
    reg = o3d.pipelines.registration
    d = synth.make_c4(n=3000, seed=5)
    i0, i1 = orc.match_correspondence(d["src_feat"], d["dst_feat"])
    src, dst = o3d.geometry.PointCloud(), o3d.geometry.PointCloud()
    src.points = o3d.utility.Vector3dVector(d["src"])
    dst.points = o3d.utility.Vector3dVector(d["dst"])
    corres = o3d.utility.Vector2iVector(np.c_[i0, i1].astype(np.int32))
    thr, max_iter, edge = 0.02, 100000, 0.9
    if hasattr(o3d.utility, "random"):
        o3d.utility.random.seed(1)
    res = reg.registration_ransac_based_on_correspondence(
        src, dst, corres, thr, reg.TransformationEstimationPointToPoint(False), 3,
        [reg.CorrespondenceCheckerBasedOnEdgeLength(edge), reg.CorrespondenceCheckerBasedOnDistance(thr)],
        reg.RANSACConvergenceCriteria(max_iter, 0.999))
    # given triples -> the 3-point transform (Eigen::umeyama inside Open3D), for bit-level comparison of the solve
    rng = np.random.default_rng(3)
    triples = rng.integers(0, len(i0), (256, 3))
    T3 = np.stack([reg.TransformationEstimationPointToPoint(False).compute_transformation(
        src, dst, o3d.utility.Vector2iVector(np.c_[i0[t], i1[t]].astype(np.int32))) for t in triples])
    # fitness / rmse of given transforms (EvaluateRANSACBasedOnCorrespondence is not exposed; evaluate_registration
    # uses nearest neighbours instead, so the per-correspondence figures are recomputed by the tests from T3)
    all_pairs = o3d.utility.Vector2iVector(np.c_[i0, i1].astype(np.int32))
    T_ls = reg.TransformationEstimationPointToPoint(False).compute_transformation(src, dst, all_pairs)
    T_ls_scale = reg.TransformationEstimationPointToPoint(True).compute_transformation(src, dst, all_pairs)
    out = os.path.join(ROOT, "tests", "golden", "reg_open3d.npz")
    np.savez_compressed(out, open3d_version=o3d.__version__, i0=i0, i1=i1, thr=thr, max_iter=max_iter, edge=edge,
                        T=np.asarray(res.transformation), fitness=res.fitness, inlier_rmse=res.inlier_rmse,
                        n_corres_inliers=len(np.asarray(res.correspondence_set)), triples=triples, T3=T3, T_ls=T_ls,
                        T_ls_scale=T_ls_scale)
    print("wrote", out, "open3d", o3d.__version__, "fitness", res.fitness, "rmse", res.inlier_rmse)
    return 0


if __name__ == "__main__":
    sys.exit(main())
