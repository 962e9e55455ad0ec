"""Independent numpy restatement of Open3D's ComputeFPFHFeature (hybrid search by brute force, pair features, SPFH,
FPFH) and of point-to-point ICP, for small clouds -- the second opinion on oracle/m3d_oracle_features.cpp (both are
restatements of third-party code that is not in the reference tree: parity unpinned)."""
import numpy as np


def hybrid(xyz, i, radius, max_nn):
    d2 = ((xyz - xyz[i]) ** 2).sum(1)
    idx = np.lexsort((np.arange(len(xyz)), d2))
    idx = idx[d2[idx] < radius * radius][:max_nn]
    return idx, d2[idx]


def pair_features(p1, n1, p2, n2):
    dp = p2 - p1
    dist = np.linalg.norm(dp)
    if dist == 0:
        return np.zeros(4)
    a, b = n1.copy(), n2.copy()
    angle1, angle2 = a @ dp / dist, b @ dp / dist
    if np.arccos(abs(angle1)) > np.arccos(abs(angle2)):
        a, b = n2.copy(), n1.copy()
        dp = -dp
        f2 = -angle2
    else:
        f2 = angle1
    v = np.cross(dp, a)
    vn = np.linalg.norm(v)
    if vn == 0:
        return np.zeros(4)
    v /= vn
    w = np.cross(a, v)
    return np.array([np.arctan2(w @ b, a @ b), v @ b, f2, dist])


def fpfh(xyz, nrm, radius, max_nn):
    n = len(xyz)
    nb = [hybrid(xyz, i, radius, max_nn) for i in range(n)]
    spfh = np.zeros((n, 33))
    for i, (idx, d2) in enumerate(nb):
        if len(idx) > 1:
            incr = 100.0 / (len(idx) - 1)
            for j in idx[1:]:
                f = pair_features(xyz[i], nrm[i], xyz[j], nrm[j])
                for h, off in ((int(np.floor(11 * (f[0] + np.pi) / (2 * np.pi))), 0), (int(np.floor(11 * (f[1] + 1) * 0.5)), 11),
                               (int(np.floor(11 * (f[2] + 1) * 0.5)), 22)):
                    spfh[i, off + min(max(h, 0), 10)] += incr
    out = np.zeros((n, 33))
    for i, (idx, d2) in enumerate(nb):
        if len(idx) > 1:
            s = np.zeros(3)
            for j, d in zip(idx[1:], d2[1:]):
                if d == 0:
                    continue
                val = spfh[j] / d
                s += val.reshape(3, 11).sum(1)
                out[i] += val
            s = np.where(s != 0, 100.0 / np.where(s != 0, s, 1), 0)
            out[i] = out[i] * np.repeat(s, 11) + spfh[i]
    return np.asfortranarray(out.T)


def kabsch(src, dst):
    mu_s, mu_d = src.mean(0), dst.mean(0)
    U, D, Vt = np.linalg.svd((dst - mu_d).T @ (src - mu_s) / len(src))
    S = np.ones(3)
    if np.linalg.det(U) * np.linalg.det(Vt) < 0:
        S[2] = -1
    R = U @ np.diag(S) @ Vt
    T = np.eye(4)
    T[:3, :3], T[:3, 3] = R, mu_d - R @ mu_s
    return T


def icp(src, dst, max_dist, T_init=None, max_iter=30, rel_fitness=1e-6, rel_rmse=1e-6):
    T = np.eye(4) if T_init is None else np.array(T_init, float)
    pcd = src @ T[:3, :3].T + T[:3, 3]

    def evaluate():
        d2 = ((pcd[:, None, :] - dst[None, :, :]) ** 2).sum(2)
        j = d2.argmin(1)
        m = d2[np.arange(len(pcd)), j]
        ok = m < max_dist * max_dist
        n = int(ok.sum())
        return ok, j, (n / len(pcd) if n else 0.0), (float(np.sqrt(m[ok].sum() / n)) if n else 0.0)
    ok, j, fit, rmse = evaluate()
    it = 0
    while it < max_iter:
        if not ok.any():
            break
        U = kabsch(pcd[ok], dst[j[ok]])
        T = U @ T
        pcd = pcd @ U[:3, :3].T + U[:3, 3]
        bf, br = fit, rmse
        ok, j, fit, rmse = evaluate()
        it += 1
        if abs(bf - fit) < rel_fitness and abs(br - rmse) < rel_rmse:
            break
    return T, fit, rmse, it
