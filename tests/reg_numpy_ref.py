"""A second, independent restatement of the correspondence-RANSAC registration loop (SURVEY.md Appendix B:
Open3D v0.15.1 RegistrationRANSACBasedOnCorrespondence as RANSACSolver::Solve drives it,
src/transform_estimation.cpp:124-164) -- written against numpy only (Kabsch / Umeyama through np.linalg.svd,
vectorised scoring), sharing nothing with oracle/m3d_oracle.cpp except the recorded sample table.  The oracle,
the GPU path and this file have to agree: the registration rows (a19, a20) have no compiled reference to pin them,
so two independently written CPU versions + the GPU is what stands in for it."""
import numpy as np


def umeyama_np(src, dst, with_scaling=False):
    """Umeyama 1991 / Eigen::umeyama: dst ~ c R src + t.  src, dst (n, 3)."""
    src, dst = np.asarray(src, float), np.asarray(dst, float)
    n = len(src)
    mu_s, mu_d = src.mean(0), dst.mean(0)
    sc, dc = src - mu_s, dst - mu_d
    cov = dc.T @ sc / n
    U, D, Vt = np.linalg.svd(cov)
    S = np.ones(3)
    if np.linalg.det(U) * np.linalg.det(Vt) < 0:
        S[2] = -1.0
    R = U @ np.diag(S) @ Vt
    c = 1.0
    if with_scaling:
        c = float((D * S).sum() / (sc ** 2).sum(1).mean())
    T = np.eye(4)
    T[:3, :3] = c * R
    T[:3, 3] = mu_d - c * R @ mu_s
    return T


def ransac_registration_np(src, dst, c0, c1, picks, thr, max_iter, edge_thr, confidence):
    """picks: (rows, 3) correspondence indices in draw order (row r is the r-th iteration that draws).
    Returns (T, stats) with the loop statistics of the sequential loop."""
    src, dst = np.asarray(src, float), np.asarray(dst, float)
    c0, c1 = np.asarray(c0, np.int64), np.asarray(c1, np.int64)
    m = len(c0)
    P, Q = src[c0], dst[c1]                     # corresponding points, (m, 3)
    best_fit, best_rmse = 0.0, 0.0
    T_best = np.eye(4)
    st = {"best_index": 0, "best_count": 0, "best_rmse": 0.0, "evaluated": 0, "stop_index": max_iter}
    est_k = max_iter
    for itr in range(max_iter):
        if not itr < est_k:
            st["stop_index"] = itr
            break
        pk = picks[itr]
        p, q = P[pk], Q[pk]
        T = umeyama_np(p, q)
        ok = True
        for i in range(3):                       # CorrespondenceCheckerBasedOnEdgeLength
            for j in range(i + 1, 3):
                ds, dt = np.linalg.norm(p[i] - p[j]), np.linalg.norm(q[i] - q[j])
                if ds < dt * edge_thr or dt < ds * edge_thr:
                    ok = False
        if ok:                                   # CorrespondenceCheckerBasedOnDistance
            tp = p @ T[:3, :3].T + T[:3, 3]
            if np.any(np.linalg.norm(q - tp, axis=1) > thr):
                ok = False
        if not ok:
            continue
        st["evaluated"] += 1
        d2 = ((P @ T[:3, :3].T + T[:3, 3] - Q) ** 2).sum(1)
        inl = d2 < thr * thr
        good = int(inl.sum())
        fitness = good / m if good else 0.0
        rmse = float(np.sqrt(d2[inl].sum() / good)) if good else 0.0
        if fitness > best_fit or (fitness == best_fit and rmse < best_rmse):
            best_fit, best_rmse, T_best = fitness, rmse, T
            st.update(best_index=itr, best_count=good, best_rmse=rmse)
            with np.errstate(divide="ignore", invalid="ignore"):
                est = np.log(1.0 - confidence) / np.log(1.0 - (good / m) ** 3)
            if est < est_k:
                c = np.ceil(est)
                est_k = int(c) if np.isfinite(c) and -2147483648.0 <= c <= 2147483647.0 else -2147483648
    return T_best, st
