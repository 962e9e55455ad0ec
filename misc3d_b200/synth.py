"""Synthetic workloads C1..C5; numpy default_rng, fp64 AoS."""
import numpy as np

SEED = 20240917


def _unit(v):
    v = np.asarray(v, dtype=np.float64)
    return v / np.linalg.norm(v)


def _plane_points(rng, n, normal, d, sigma, extent=1.0):
    """n points on the plane normal.x + d = 0, in-plane coords U(-extent,extent)^2, noise along normal"""
    nrm = _unit(normal)
    a = np.array([1.0, 0, 0]) if abs(nrm[0]) < 0.9 else np.array([0, 1.0, 0])
    u = _unit(np.cross(nrm, a))
    v = np.cross(nrm, u)
    uv = rng.uniform(-extent, extent, size=(n, 2))
    noise = rng.normal(0.0, sigma, size=n) if sigma > 0 else np.zeros(n)
    return uv[:, :1] * u + uv[:, 1:] * v + (noise - d)[:, None] * nrm


def make_c1(n=50_000, seed=SEED, sigma=0.003, inlier_frac=0.7):
    """70% on plane n=(0.1,-0.2,0.97)/|.|, d=0.3 ; 30% uniform outliers in [-1,1]^3; shuffled"""
    rng = np.random.default_rng(seed)
    n_in = int(round(inlier_frac * n))
    pts = np.concatenate([
        _plane_points(rng, n_in, (0.1, -0.2, 0.97), 0.3, sigma),
        rng.uniform(-1, 1, size=(n - n_in, 3)),
    ])
    return np.ascontiguousarray(pts[rng.permutation(n)])


def make_c2(n=1_000_000, seed=SEED, sigma=0.003):
    """40% plane + 20% sphere + 20% cylinder + 20% outliers, with normals. Returns (xyz, normals)."""
    rng = np.random.default_rng(seed)
    n_pl = int(0.4 * n)
    n_sp = int(0.2 * n)
    n_cy = int(0.2 * n)
    n_out = n - n_pl - n_sp - n_cy
    pl = _plane_points(rng, n_pl, (0.1, -0.2, 0.97), 0.3, sigma)
    pl_n = np.tile(_unit((0.1, -0.2, 0.97)), (n_pl, 1))
    # sphere c=(0.2,0.1,-0.3) r=0.4
    dirs = rng.normal(size=(n_sp, 3))
    dirs /= np.linalg.norm(dirs, axis=1, keepdims=True)
    sp = np.array([0.2, 0.1, -0.3]) + dirs * (0.4 + rng.normal(0, sigma, size=(n_sp, 1)))
    sp_n = dirs
    # cylinder axis through (-0.3,0,0) dir (0,0,1), r=0.15, |z|<0.5
    th = rng.uniform(0, 2 * np.pi, size=n_cy)
    rad = 0.15 + rng.normal(0, sigma, size=n_cy)
    cy = np.stack([-0.3 + rad * np.cos(th), rad * np.sin(th), rng.uniform(-0.5, 0.5, size=n_cy)], 1)
    cy_n = np.stack([np.cos(th), np.sin(th), np.zeros(n_cy)], 1)
    # 2 degrees of angular noise on the cylinder normals
    cy_n = cy_n + np.tan(np.deg2rad(2.0)) * rng.normal(size=(n_cy, 3)) / np.sqrt(3.0)
    cy_n /= np.linalg.norm(cy_n, axis=1, keepdims=True)
    out = rng.uniform(-1, 1, size=(n_out, 3))
    out_n = rng.normal(size=(n_out, 3))
    out_n /= np.linalg.norm(out_n, axis=1, keepdims=True)
    xyz = np.concatenate([pl, sp, cy, out])
    nrm = np.concatenate([pl_n, sp_n, cy_n, out_n])
    perm = rng.permutation(n)
    return np.ascontiguousarray(xyz[perm]), np.ascontiguousarray(nrm[perm])


def make_c3(n=2_000_000, seed=SEED, sigma=0.002):
    """six faces of [-1,1]^3 with shares 30/20/15/12/10/8 % + 5 % outliers"""
    rng = np.random.default_rng(seed)
    shares = [0.30, 0.20, 0.15, 0.12, 0.10, 0.08]
    faces = [((1, 0, 0), -1.0), ((-1, 0, 0), -1.0), ((0, 1, 0), -1.0), ((0, -1, 0), -1.0),
             ((0, 0, 1), -1.0), ((0, 0, -1), -1.0)]
    parts = []
    used = 0
    for s, (nrm, d) in zip(shares, faces):
        k = int(s * n)
        parts.append(_plane_points(rng, k, nrm, d, sigma))
        used += k
    parts.append(rng.uniform(-1, 1, size=(n - used, 3)))
    xyz = np.concatenate(parts)
    return np.ascontiguousarray(xyz[rng.permutation(n)])


def rotation_about(axis, deg):
    a = _unit(axis)
    t = np.deg2rad(deg)
    K = np.array([[0, -a[2], a[1]], [a[2], 0, -a[0]], [-a[1], a[0], 0]])
    return np.eye(3) + np.sin(t) * K + (1 - np.cos(t)) * (K @ K)


def make_c4(n=200_000, seed=SEED, dim=33, sigma=0.002, true_frac=0.3):
    """registration pair + 33-D descriptors.

    Returns dict(src, dst, src_feat (dim,n) F-order, dst_feat (dim,n) F-order, T_true, perm)
    dst[j] = R src[perm[j]] + t + noise ; descriptors of true matches = src descriptor + N(0,1)."""
    rng = np.random.default_rng(seed)
    src = rng.uniform(-1, 1, size=(n, 3))
    R = rotation_about((1, 1, 1), 30.0)
    t = np.array([0.1, -0.2, 0.05])
    perm = rng.permutation(n)
    dst = src[perm] @ R.T + t + rng.normal(0, sigma, size=(n, 3))
    base = rng.uniform(0, 100, size=(n, dim))
    dfeat = rng.uniform(0, 100, size=(n, dim))
    true_mask = rng.uniform(size=n) < true_frac
    dfeat[true_mask] = base[perm][true_mask] + rng.normal(0, 1, size=(int(true_mask.sum()), dim))
    T = np.eye(4)
    T[:3, :3] = R
    T[:3, 3] = t
    return dict(src=np.ascontiguousarray(src), dst=np.ascontiguousarray(dst),
                src_feat=np.asfortranarray(base.T), dst_feat=np.asfortranarray(dfeat.T),
                T_true=T, perm=perm, true_mask=true_mask)


def make_c5(n=4_000_000, seed=SEED):
    return make_c1(n=n, seed=seed)


def make_surface_pair(n=20_000, seed=SEED, sigma=0.0):
    """a registration pair with ANALYTIC normals for the FPFH -> match -> RANSAC -> ICP chain: points on the bumpy
    height field z = 0.25 sin(3x) cos(2y) + 0.1 sin(7x + 5y) over [-1, 1]^2 (locally distinctive, so descriptors
    discriminate), dst = R src[perm] + t (+ noise).  Returns dict(src, src_nrm, dst, dst_nrm, T_true, perm)."""
    rng = np.random.default_rng(seed)
    x, y = rng.uniform(-1, 1, n), rng.uniform(-1, 1, n)
    z = 0.25 * np.sin(3 * x) * np.cos(2 * y) + 0.1 * np.sin(7 * x + 5 * y)
    zx = 0.75 * np.cos(3 * x) * np.cos(2 * y) + 0.7 * np.cos(7 * x + 5 * y)
    zy = -0.5 * np.sin(3 * x) * np.sin(2 * y) + 0.5 * np.cos(7 * x + 5 * y)
    src = np.c_[x, y, z]
    nrm = np.c_[-zx, -zy, np.ones(n)]
    nrm /= np.linalg.norm(nrm, axis=1, keepdims=True)
    R = rotation_about((1, 2, -1), 25.0)
    t = np.array([0.3, -0.1, 0.2])
    perm = rng.permutation(n)
    dst = src[perm] @ R.T + t + rng.normal(0, sigma, size=(n, 3)) if sigma else src[perm] @ R.T + t
    T = np.eye(4)
    T[:3, :3], T[:3, 3] = R, t
    return dict(src=np.ascontiguousarray(src), src_nrm=np.ascontiguousarray(nrm), dst=np.ascontiguousarray(dst),
                dst_nrm=np.ascontiguousarray(nrm[perm] @ R.T), T_true=T, perm=perm)
