"""GPU parity tests of match_correspondence (brute-force L2 + mutual check) against the oracle:
index lists must be identical (integer work: bit-exact)."""
import os

import numpy as np
import pytest

from misc3d_b200 import synth

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def test_golden_reg_small_matches(ctx, capi):
    g = np.load(os.path.join(GOLD, "reg_small.npz"))
    d = synth.make_c4(n=3000, seed=5)
    i0, i1, ms = ctx.match_correspondence(d["src_feat"], d["dst_feat"])
    np.testing.assert_array_equal(i0, g["i0"])
    np.testing.assert_array_equal(i1, g["i1"])


@pytest.mark.parametrize("path", ["tc", "fp32", "fp64"])
@pytest.mark.parametrize("ns,nd,dim", [(1000, 1300, 33), (257, 129, 33), (500, 500, 8), (300, 200, 352), (1, 5, 33),
                                       (130, 1, 3), (700, 900, 47), (400, 300, 48)])
def test_nearest_and_match_equal_oracle(ctx, capi, orc, ns, nd, dim, path, monkeypatch):
    """all three search kernels (tcgen05 bf16x3 GEMM, fp32 CUDA-core tiles, fp64 only) + the fp64
    re-search of close calls must give the oracle's indices"""
    if path != "tc":
        monkeypatch.setenv("M3D_MATCH_PATH", path)
    rng = np.random.default_rng(ns + nd + dim)
    a = np.asfortranarray(rng.uniform(0, 100, size=(dim, ns)))
    b = np.asfortranarray(rng.uniform(0, 100, size=(dim, nd)))
    nn, ms = ctx.nearest(a, b)
    np.testing.assert_array_equal(nn, orc.nearest(a, b))
    for method in (capi.MATCH_FLANN, capi.MATCH_ANNOY):
        i0, i1, ms = ctx.match_correspondence(a, b, method=method)
        o0, o1 = orc.match_correspondence(a, b)
        np.testing.assert_array_equal(i0, o0)
        np.testing.assert_array_equal(i1, o1)
        assert np.all(np.diff(i0.astype(np.int64)) > 0)


@pytest.mark.parametrize("path", ["tc", "fp32"])
def test_exact_ties_pick_the_lowest_index(ctx, capi, orc, path, monkeypatch):
    if path != "tc":
        monkeypatch.setenv("M3D_MATCH_PATH", path)
    """FPFH histograms repeat on real data: duplicated descriptors must resolve to the lowest index"""
    rng = np.random.default_rng(3)
    base = np.round(rng.uniform(0, 10, size=(33, 40)))  # small integer lattice -> many equal distances
    a = np.asfortranarray(base[:, rng.integers(0, 40, 700)])
    b = np.asfortranarray(base[:, rng.integers(0, 40, 900)])
    nn, ms = ctx.nearest(a, b)
    np.testing.assert_array_equal(nn, orc.nearest(a, b))
    i0, i1, ms = ctx.match_correspondence(a, b)
    o0, o1 = orc.match_correspondence(a, b)
    np.testing.assert_array_equal(i0, o0)
    np.testing.assert_array_equal(i1, o1)


def test_offset_descriptors_and_empty_sets(ctx, capi, orc):
    rng = np.random.default_rng(8)
    a = np.asfortranarray(rng.normal(0, 1, size=(33, 600)) + 1.0e4)  # large common offset: centring matters
    b = np.asfortranarray(a[:, rng.permutation(600)] + rng.normal(0, 0.05, size=(33, 600)))
    i0, i1, ms = ctx.match_correspondence(a, b)
    o0, o1 = orc.match_correspondence(a, b)
    np.testing.assert_array_equal(i0, o0)
    np.testing.assert_array_equal(i1, o1)
    e = np.zeros((33, 0), order="F")
    i0, i1, ms = ctx.match_correspondence(e, b)
    assert len(i0) == 0 and len(i1) == 0
    i0, i1, ms = ctx.match_correspondence(a, e)
    assert len(i0) == 0 and len(i1) == 0


def test_c4_sized_properties(ctx, capi, orc):
    """BASELINE config C4 size (200k x 200k x 33): mutual pairs are consistent and recover the
    planted correspondences (30 % of the descriptors are noisy copies); 2048 random query columns of each direction
    against the oracle's exact search over the FULL 200k database (nanoflann accumulation order, ties to the lowest
    index)."""
    d = synth.make_c4()
    i0, i1, ms = ctx.match_correspondence(d["src_feat"], d["dst_feat"])
    assert np.all(np.diff(i0.astype(np.int64)) > 0) and len(np.unique(i1)) == len(i1)
    perm, mask = d["perm"], d["true_mask"]          # dst[j] <- src[perm[j]] where mask[j]
    truth = {int(perm[j]): int(j) for j in np.nonzero(mask)[0]}
    got = dict(zip(i0.tolist(), i1.tolist()))
    hit = sum(1 for s, t in truth.items() if got.get(s) == t)
    assert hit >= 0.999 * len(truth)
    # spot-check 64 rows against a numpy brute force
    rng = np.random.default_rng(0)
    rows = rng.choice(len(i0), 64, replace=False)
    A, B = d["src_feat"], d["dst_feat"]
    for r in rows:
        dist = ((B - A[:, [int(i0[r])]]) ** 2).sum(0)
        assert int(np.argmin(dist)) == int(i1[r])
    # one direction of the search on its own, against the oracle on the full database
    nn, _ = ctx.nearest(A, B)
    q = rng.choice(A.shape[1], 2048, replace=False)
    np.testing.assert_array_equal(nn[q], orc.nearest(A[:, q], B))
    nn_back, _ = ctx.nearest(B, A)
    np.testing.assert_array_equal(nn_back[q], orc.nearest(B[:, q], A))
    mutual = np.nonzero(nn_back[nn.astype(np.int64)] == np.arange(A.shape[1], dtype=np.uint64))[0]
    np.testing.assert_array_equal(mutual, i0.astype(np.int64))      # the mutual filter, recomputed from the two searches
    np.testing.assert_array_equal(nn[mutual], i1)


def test_knn_index_is_exact(ctx, capi):
    """KNearestSearch's backend (m3d_knn_*): exact neighbours in ascending distance, ties to the lower index, Euclidean
    (unsquared) distances, radius cut -- against numpy brute force, for descriptor (33-D) and point (3-D) data"""
    rng = np.random.default_rng(3)
    for dim, n, nq, k in ((33, 5000, 40, 7), (3, 20000, 25, 16), (3, 10, 3, 16)):
        data = rng.uniform(0, 100, (dim, n))
        data[:, 1] = data[:, 0]                      # an exact tie
        q = np.c_[data[:, :2], rng.uniform(0, 100, (dim, nq - 2))]
        idx, dist, cnt = ctx.knn_search(data, q, k)
        for j in range(nq):
            d = np.sqrt(((data - q[:, j:j + 1]) ** 2).sum(0))
            order = np.lexsort((np.arange(n), d))[:k]
            assert cnt[j] == min(k, n)
            np.testing.assert_array_equal(idx[j, :cnt[j]], order)
            np.testing.assert_allclose(dist[j, :cnt[j]], d[order], rtol=1e-12, atol=1e-12)
        r = float(np.median(dist[:, min(k, n) // 2]))
        idx2, dist2, cnt2 = ctx.knn_search(data, q, k, radius=r)
        for j in range(nq):
            assert cnt2[j] == int((dist[j, :cnt[j]] <= r).sum())
            np.testing.assert_array_equal(idx2[j, :cnt2[j]], idx[j, :cnt2[j]])


def test_two_cta_kernel_variant_in_a_subprocess(orc):
    """the experimental cta_group::2 form of the tcgen05 search (M3D_MATCH_TC=2, read once per process) must return the
    same mutual matches as the oracle"""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = ("import sys, numpy as np; sys.path.insert(0, %r); from misc3d_b200 import capi, synth; "
            "d = synth.make_c4(n=20000, seed=3); c = capi.Context(0); "
            "i0, i1, ms = c.match_correspondence(d['src_feat'], d['dst_feat']); "
            "np.save(sys.argv[1], np.stack([i0, i1]))") % root
    out = os.path.join(root, "gpurun_out", "_tc2_match.npy")
    os.makedirs(os.path.dirname(out), exist_ok=True)
    env = dict(os.environ, M3D_MATCH_TC="2")
    r = subprocess.run([sys.executable, "-c", code, out], env=env, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    got = np.load(out)
    d = synth.make_c4(n=20000, seed=3)
    o0, o1 = orc.match_correspondence(d["src_feat"], d["dst_feat"])
    np.testing.assert_array_equal(got[0], o0)
    np.testing.assert_array_equal(got[1], o1)
