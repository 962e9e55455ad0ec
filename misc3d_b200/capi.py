"""ctypes binding of libm3d_b200.so (include/m3d_capi.h).

Thin plumbing only: numpy buffers in, numpy buffers out.  There is no CPU fallback: if the shared
library is missing, or no CUDA device is usable, the calls raise.
"""
import atexit
import ctypes as C
import weakref
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("M3D_LIB") or os.path.join(_HERE, "libm3d_b200.so")  # M3D_LIB: tuning builds only

PLANE, SPHERE, CYLINDER = 0, 1, 2
KSAMPLE = {PLANE: 3, SPHERE: 4, CYLINDER: 2}
NPARAM = {PLANE: 4, SPHERE: 4, CYLINDER: 7}
MATCH_FLANN, MATCH_ANNOY = 0, 1
FLAG_EXACT_ONLY, FLAG_NO_REFIT, FLAG_DENSE, FLAG_CLASSIFY, FLAG_STATS = 1, 2, 4, 8, 16
FLAG_CHUNKED_UPLOAD, FLAG_PLAIN_UPLOAD, FLAG_REGISTER_HOST = 32, 64, 128

OK = 0
ERR_INVALID_ARG, ERR_TOO_FEW_POINTS, ERR_PROBABILITY, ERR_NO_NORMALS = -1, -2, -3, -4
ERR_CUDA, ERR_NCCL, ERR_NO_INLIERS, ERR_CAPACITY, ERR_INTERNAL = -5, -6, -7, -8, -9

EXPORTS = [
    "m3d_abi_version", "m3d_device_count", "m3d_ctx_create", "m3d_ctx_create_on_stream",
    "m3d_ctx_destroy", "m3d_last_error", "m3d_ctx_stream", "m3d_ctx_launch_count", "m3d_probe_fp32_ffma", "m3d_probe_fp64_dfma",
    "m3d_nccl_unique_id", "m3d_ctx_init_nccl", "m3d_ctx_set_exchange", "m3d_ransac_fit",
    "m3d_cloud_upload", "m3d_cloud_from_device", "m3d_cloud_free", "m3d_cloud_size",
    "m3d_ransac_fit_cloud", "m3d_score_samples", "m3d_evaluate_model", "m3d_sample_table",
    "m3d_ordered_scan", "m3d_segment_plane_iterative", "m3d_match_correspondence", "m3d_nearest",
    "m3d_ransac_registration", "m3d_least_squares_transform", "m3d_registration_refit", "m3d_host_unregister_all", "m3d_fpfh_create", "m3d_features_upload",
    "m3d_features_download", "m3d_features_count", "m3d_features_dim", "m3d_features_free", "m3d_match_features", "m3d_shard_rows", "m3d_sample_table_device", "m3d_score_stats",
    "m3d_knn_create", "m3d_knn_free", "m3d_knn_search", "m3d_segment_plane_iterative_u32",
    "m3d_compute_fpfh", "m3d_icp_point_to_point",
]


class RansacParams(C.Structure):
    _fields_ = [("threshold", C.c_double), ("max_iteration", C.c_uint64), ("probability", C.c_double),
                ("seed", C.c_uint32), ("flags", C.c_uint32)]


class _Dictable(C.Structure):
    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


class RansacStats(_Dictable):
    _fields_ = [("best_index", C.c_uint64), ("best_count", C.c_uint64), ("best_rmse", C.c_double),
                ("iterations_run", C.c_uint64), ("stop_index", C.c_uint64), ("evaluated", C.c_uint64),
                ("exact_resolves", C.c_uint64), ("found", C.c_int32), ("refit_ok", C.c_int32),
                ("device_ms", C.c_float), ("score_ms", C.c_float), ("refine_ms", C.c_float), ("draw_ms", C.c_float)]


class RegStats(_Dictable):
    _fields_ = [("best_index", C.c_uint64), ("best_count", C.c_uint64), ("best_rmse", C.c_double),
                ("evaluated", C.c_uint64), ("stop_index", C.c_uint64), ("device_ms", C.c_float),
                ("score_ms", C.c_float)]


ALLGATHER_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_int)


class M3DError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"m3d error {code}: {msg}")
        self.code = code


_lib = None


def lib():
    """Load libm3d_b200.so; raises (never falls back) when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                              "(make -C misc3d_b200/csrc); there is no CPU fallback")
        L = C.CDLL(LIB_PATH)
        L.m3d_last_error.restype = C.c_char_p
        L.m3d_ctx_stream.restype = C.c_void_p
        L.m3d_ctx_launch_count.restype = C.c_uint64
        L.m3d_cloud_size.restype = C.c_size_t
        L.m3d_ctx_destroy.restype = None
        L.m3d_cloud_free.restype = None
        L.m3d_sample_table.restype = None
        L.m3d_ctx_destroy.argtypes = [C.c_void_p]
        L.m3d_cloud_free.argtypes = [C.c_void_p]
        L.m3d_knn_free.restype = None
        L.m3d_knn_free.argtypes = [C.c_void_p]
        L.m3d_features_free.restype = None
        L.m3d_features_free.argtypes = [C.c_void_p]
        L.m3d_features_count.restype = C.c_size_t
        L.m3d_features_count.argtypes = [C.c_void_p]
        L.m3d_features_dim.argtypes = [C.c_void_p]
        L.m3d_features_download.argtypes = [C.c_void_p, C.POINTER(C.c_double)]
        L.m3d_host_unregister_all.restype = None
        L.m3d_host_unregister_all.argtypes = [C.c_void_p]
        L.m3d_last_error.argtypes = [C.c_void_p]
        L.m3d_ctx_stream.argtypes = [C.c_void_p]
        L.m3d_ctx_launch_count.argtypes = [C.c_void_p]
        L.m3d_cloud_size.argtypes = [C.c_void_p]
        _lib = L
    return _lib


def _p(a, t=C.c_double):
    return a.ctypes.data_as(C.POINTER(t)) if a is not None else None


def _f64(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.float64)


def nccl_unique_id():
    buf = C.create_string_buffer(128)
    rc = lib().m3d_nccl_unique_id(buf)
    if rc != 0:
        raise M3DError(rc, "libnccl.so.2 could not be loaded")
    return bytes(buf.raw)


def device_count():
    return int(lib().m3d_device_count())


def shard_rows(rows, rank, world):
    """(wave rows of rank `rank` in its local order, padded per-rank stride) -- m3d_shard_rows"""
    n_local, padded = C.c_size_t(0), C.c_size_t(0)
    L = lib()
    L.m3d_shard_rows.restype = None
    L.m3d_shard_rows(C.c_size_t(rows), C.c_int(rank), C.c_int(world), None, C.byref(n_local), C.byref(padded))
    out = np.empty(max(n_local.value, 1), dtype=np.uint32)
    L.m3d_shard_rows(C.c_size_t(rows), C.c_int(rank), C.c_int(world), out.ctypes.data_as(C.POINTER(C.c_uint32)),
                     C.byref(n_local), C.byref(padded))
    return out[:n_local.value], int(padded.value)


def sample_table(seed, n, k, rows):
    out = np.empty((rows, k), dtype=np.uint32)
    lib().m3d_sample_table(C.c_uint32(seed & 0xFFFFFFFF), C.c_size_t(n), C.c_int(k), C.c_size_t(rows),
                           _p(out, C.c_uint32))
    return out


def ordered_scan(counts, valid, err, n_points, k, probability, max_iteration):
    counts = np.ascontiguousarray(counts, dtype=np.uint64)
    valid = np.ascontiguousarray(valid, dtype=np.uint8)
    e = None if err is None else np.ascontiguousarray(err, dtype=np.float64)
    st = RansacStats()
    rc = lib().m3d_ordered_scan(_p(counts, C.c_uint64), _p(valid, C.c_uint8), _p(e), C.c_size_t(len(counts)),
                                C.c_size_t(n_points), C.c_int(k), C.c_double(probability),
                                C.c_uint64(max_iteration), C.byref(st))
    if rc != 0:
        raise M3DError(rc, "m3d_ordered_scan")
    return st.as_dict()


_live_contexts = weakref.WeakSet()


@atexit.register
def _close_all():
    """destroy every context (and its clouds) while the CUDA runtime and this module are still alive"""
    for c in list(_live_contexts):
        try:
            c.close()
        except Exception:
            pass


class Cloud:
    def __init__(self, ctx, handle, n, has_normals):
        self.ctx, self.handle, self.n, self.has_normals = ctx, handle, n, has_normals
        ctx._clouds.add(self)

    def free(self):
        """releases the device buffers; a cloud is always freed before its context is destroyed
        (Context.close frees the clouds it still owns), whatever order the interpreter tears objects down in"""
        if self.handle:
            if getattr(self.ctx, "h", None):
                lib().m3d_cloud_free(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class Features:
    """device-resident descriptors (m3d_features): dim x n float64 in HBM, owned by this object"""

    def __init__(self, ctx, handle):
        self.ctx, self.handle = ctx, handle
        self.dim = int(lib().m3d_features_dim(handle))
        self.n = int(lib().m3d_features_count(handle))
        ctx._clouds.add(self)   # freed with the context at the latest, like clouds

    def download(self):
        """(dim, n) float64, Fortran order (what match_correspondence takes)"""
        out = np.zeros((max(self.n, 1), self.dim))
        rc = lib().m3d_features_download(self.handle, _p(out))
        if rc != 0:
            self.ctx._check(rc)
        return np.asfortranarray(out[:self.n].T)

    def free(self):
        if self.handle:
            if getattr(self.ctx, "h", None):
                lib().m3d_features_free(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class Context:
    """One m3d_ctx: a CUDA stream + scratch (+ optional exchange)."""

    def __init__(self, device=0, stream=None):
        L = lib()
        h = C.c_void_p()
        if stream is None:
            rc = L.m3d_ctx_create(C.c_int(device), C.byref(h))
        else:
            rc = L.m3d_ctx_create_on_stream(C.c_int(device), C.c_void_p(stream), C.byref(h))
        if rc != 0:
            raise M3DError(rc, "cannot create a CUDA context (no usable GPU? there is no CPU fallback)")
        self.h = h
        self.device = device
        self._cb = None
        self._clouds = weakref.WeakSet()
        _live_contexts.add(self)

    def close(self):
        if getattr(self, "h", None):
            for c in list(getattr(self, "_clouds", ())):
                c.free()
            lib().m3d_ctx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc < 0:
            raise M3DError(rc, (lib().m3d_last_error(self.h) or b"").decode())
        return rc

    @property
    def launches(self):
        return int(lib().m3d_ctx_launch_count(self.h))

    @property
    def stream(self):
        return lib().m3d_ctx_stream(self.h)

    def probe_fp32_ffma(self):
        v = C.c_double(0)
        self._check(lib().m3d_probe_fp32_ffma(self.h, C.byref(v)))
        return float(v.value)

    def probe_fp64_dfma(self):
        v = C.c_double(0)
        self._check(lib().m3d_probe_fp64_dfma(self.h, C.byref(v)))
        return float(v.value)

    def host_unregister_all(self):
        """drop the cudaHostRegister registrations made by fits with FLAG_REGISTER_HOST"""
        lib().m3d_host_unregister_all(self.h)

    def score_stats(self):
        """work counters of the launches run with FLAG_STATS since the last call (m3d_score_stats)"""
        out = np.zeros(8, dtype=np.uint64)
        self._check(lib().m3d_score_stats(self.h, _p(out, C.c_uint64)))
        return {"tile_tests": int(out[0]), "tile_survivors": int(out[1]), "cell_pairs": int(out[2]),
                "passes_1": int(out[3]), "passes_2": int(out[4]), "rescans": int(out[5])}

    def sample_table_device(self, seed, n, k, rows):
        """the sample table drawn on the GPU (m3d_sample_table_device): (rows, k) uint32, or None when the
        device draw is not eligible / gave up for these sizes"""
        out = np.empty((max(rows, 1), k), dtype=np.uint32)
        rc = self._check(lib().m3d_sample_table_device(self.h, C.c_uint32(seed & 0xFFFFFFFF), C.c_size_t(n), C.c_int(k),
                                                       C.c_size_t(rows), _p(out, C.c_uint32)))
        return out[:rows] if rc == 1 else None

    # ------------------------------------------------------------------ multi-GPU
    def init_nccl(self, unique_id, rank, world):
        buf = C.create_string_buffer(bytes(unique_id), 128)
        self._check(lib().m3d_ctx_init_nccl(self.h, buf, C.c_int(rank), C.c_int(world)))

    def set_exchange(self, fn, rank, world, on_device=False):
        """fn(send_ptr, recv_ptr, bytes_per_rank, on_device) -> int"""
        self._cb = ALLGATHER_FN(lambda user, s, r, nb, dev: int(fn(s, r, nb, dev))) if fn else ALLGATHER_FN(0)
        self._check(lib().m3d_ctx_set_exchange(self.h, self._cb, None, C.c_int(1 if on_device else 0),
                                               C.c_int(rank), C.c_int(world)))

    # ------------------------------------------------------------------ clouds
    def upload(self, xyz, normals=None):
        xyz = _f64(xyz).reshape(-1, 3)
        nrm = None if normals is None else _f64(normals).reshape(-1, 3)
        h = C.c_void_p()
        self._check(lib().m3d_cloud_upload(self.h, _p(xyz), _p(nrm), C.c_size_t(len(xyz)), C.byref(h)))
        return Cloud(self, h, len(xyz), nrm is not None)

    def cloud_from_device(self, d_xyz_ptr, d_nrm_ptr, n):
        h = C.c_void_p()
        self._check(lib().m3d_cloud_from_device(self.h, C.c_void_p(d_xyz_ptr),
                                                C.c_void_p(d_nrm_ptr) if d_nrm_ptr else None, C.c_size_t(n),
                                                C.byref(h)))
        return Cloud(self, h, n, bool(d_nrm_ptr))

    # ------------------------------------------------------------------ RANSAC
    @staticmethod
    def _params(threshold, max_iteration, probability, seed, flags):
        return RansacParams(float(threshold), int(max_iteration), float(probability), int(seed) & 0xFFFFFFFF,
                            int(flags))

    def ransac_fit(self, kind, xyz, normals=None, threshold=0.01, max_iteration=1000, probability=0.9999,
                   seed=0, flags=0, want_inliers=True, inl_buf=None):
        """Host-buffer entry point (m3d_ransac_fit). Returns (ret, model[np], inliers, stats).
        inl_buf: optional caller-owned uint64 buffer of >= n entries (e.g. pinned) for the inlier indices;
        the returned index array is then a view into it."""
        xyz = _f64(xyz).reshape(-1, 3)
        nrm = None if normals is None else _f64(normals).reshape(-1, 3)
        n = len(xyz)
        model = np.zeros(8)
        inl = inl_buf if inl_buf is not None else (np.empty(max(n, 1), dtype=np.uint64) if want_inliers else None)
        n_inl = C.c_size_t(0)
        st = RansacStats()
        p = self._params(threshold, max_iteration, probability, seed, flags)
        rc = self._check(lib().m3d_ransac_fit(self.h, C.c_int(kind), _p(xyz), _p(nrm), C.c_size_t(n), C.byref(p),
                                              _p(model), _p(inl, C.c_size_t), C.byref(n_inl), C.byref(st)))
        if inl is None:
            return rc, model[:NPARAM[kind]].copy(), None, st.as_dict()
        out = inl[:n_inl.value] if inl_buf is not None else inl[:n_inl.value].copy()
        return rc, model[:NPARAM[kind]].copy(), out, st.as_dict()

    def ransac_fit_cloud(self, kind, cloud, threshold=0.01, max_iteration=1000, probability=0.9999, seed=0,
                         flags=0, want_inliers=True, inl_buf=None):
        model = np.zeros(8)
        inl = inl_buf if inl_buf is not None else (np.empty(max(cloud.n, 1), dtype=np.uint64) if want_inliers else None)
        n_inl = C.c_size_t(0)
        st = RansacStats()
        p = self._params(threshold, max_iteration, probability, seed, flags)
        rc = self._check(lib().m3d_ransac_fit_cloud(self.h, C.c_int(kind), cloud.handle, C.byref(p), _p(model),
                                                    _p(inl, C.c_size_t), C.byref(n_inl), C.byref(st)))
        out_inl = inl[:n_inl.value] if inl is not None else None
        return rc, model[:NPARAM[kind]].copy(), out_inl, st.as_dict()

    def score_samples(self, kind, cloud, samples, threshold, flags=0, want_models=True):
        samples = np.ascontiguousarray(samples, dtype=np.uint32).reshape(-1, KSAMPLE[kind])
        rows = len(samples)
        models = np.zeros((rows, 8)) if want_models else None
        valid = np.zeros(rows, dtype=np.uint8) if want_models else None
        counts = np.zeros(rows, dtype=np.uint64)
        self._check(lib().m3d_score_samples(self.h, C.c_int(kind), cloud.handle, _p(samples, C.c_uint32),
                                            C.c_size_t(rows), C.c_double(threshold), C.c_uint32(flags),
                                            _p(models), _p(valid, C.c_uint8), _p(counts, C.c_uint64)))
        return counts, models, valid

    def evaluate_model(self, kind, cloud, model, threshold, sequential=False):
        m = np.zeros(8)
        m[:NPARAM[kind]] = np.asarray(model, dtype=np.float64)[:NPARAM[kind]]
        cnt = C.c_uint64(0)
        err = C.c_double(0)
        self._check(lib().m3d_evaluate_model(self.h, C.c_int(kind), cloud.handle, _p(m), C.c_double(threshold),
                                             C.c_int(1 if sequential else 0), C.byref(cnt), C.byref(err)))
        return int(cnt.value), float(err.value)

    # ------------------------------------------------------------------ segmentation
    def segment_plane_iterative(self, xyz, threshold, max_iteration=100, min_ratio=0.05, seed=0, cap_planes=256,
                                labels32=False):
        """Returns (status, planes (P,4), labels (n,) uint64 with UINT64_MAX = unassigned -- uint32 / 0xFFFFFFFF with
        labels32 -- , device_ms)."""
        xyz = _f64(xyz).reshape(-1, 3)
        n = len(xyz)
        planes = np.zeros((cap_planes, 4))
        labels = np.empty(max(n, 1), dtype=np.uint32 if labels32 else np.uint64)
        npl = C.c_size_t(0)
        ms = C.c_float(0)
        fn = lib().m3d_segment_plane_iterative_u32 if labels32 else lib().m3d_segment_plane_iterative
        rc = fn(self.h, _p(xyz), C.c_size_t(n), C.c_double(threshold), C.c_int(max_iteration), C.c_double(min_ratio),
                C.c_uint32(seed & 0xFFFFFFFF), _p(planes), C.c_size_t(cap_planes),
                _p(labels, C.c_uint32 if labels32 else C.c_uint64), C.byref(npl), C.byref(ms))
        if rc in (ERR_INVALID_ARG, ERR_CUDA, ERR_INTERNAL, ERR_NCCL):
            self._check(rc)
        return rc, planes[:npl.value].copy(), labels[:n], float(ms.value)

    # ------------------------------------------------------------------ matching / registration
    def nearest(self, src, dst):
        src = np.asfortranarray(src, dtype=np.float64)
        dst = np.asfortranarray(dst, dtype=np.float64)
        dim, ns = src.shape
        nd = dst.shape[1]
        nn = np.empty(max(ns, 1), dtype=np.uint64)
        ms = C.c_float(0)
        self._check(lib().m3d_nearest(self.h, _p(src), C.c_size_t(ns), _p(dst), C.c_size_t(nd), C.c_int(dim),
                                      _p(nn, C.c_size_t), C.byref(ms)))
        return nn[:ns], float(ms.value)

    def match_correspondence(self, src, dst, method=MATCH_ANNOY, n_trees=4):
        """src, dst: (dim, n) float64 (column = descriptor). Returns (idx0, idx1, device_ms)."""
        src = np.asfortranarray(src, dtype=np.float64)
        dst = np.asfortranarray(dst, dtype=np.float64)
        dim, ns = src.shape
        if dst.shape[0] != dim:
            raise ValueError("descriptor dimensions differ")
        nd = dst.shape[1]
        i0 = np.empty(max(ns, 1), dtype=np.uint64)
        i1 = np.empty(max(ns, 1), dtype=np.uint64)
        n_out = C.c_size_t(0)
        ms = C.c_float(0)
        self._check(lib().m3d_match_correspondence(self.h, _p(src), C.c_size_t(ns), _p(dst), C.c_size_t(nd),
                                                   C.c_int(dim), C.c_int(method), C.c_int(n_trees),
                                                   _p(i0, C.c_size_t), _p(i1, C.c_size_t), C.byref(n_out),
                                                   C.byref(ms)))
        return i0[:n_out.value].copy(), i1[:n_out.value].copy(), float(ms.value)

    def ransac_registration(self, src, dst, c0, c1, threshold=0.01, max_iter=100000, edge_thr=0.9,
                            confidence=0.999, seed=0):
        src = _f64(src).reshape(-1, 3)
        dst = _f64(dst).reshape(-1, 3)
        c0 = np.ascontiguousarray(c0, dtype=np.uint64)
        c1 = np.ascontiguousarray(c1, dtype=np.uint64)
        if len(c0) != len(c1):
            raise ValueError("correspondence index lists differ in length")
        T = np.zeros(16)
        st = RegStats()
        rc = self._check(lib().m3d_ransac_registration(self.h, _p(src), C.c_size_t(len(src)), _p(dst),
                                                       C.c_size_t(len(dst)), _p(c0, C.c_size_t), _p(c1, C.c_size_t),
                                                       C.c_size_t(len(c0)), C.c_double(threshold), C.c_int(max_iter),
                                                       C.c_double(edge_thr), C.c_double(confidence),
                                                       C.c_uint32(seed & 0xFFFFFFFF), _p(T), C.byref(st)))
        return rc, T.reshape(4, 4).copy(), st.as_dict()

    def knn_search(self, data, queries, k, radius=0.0):
        """exact k-NN (m3d_knn_*): data (dim, n), queries (dim, nq) float64 -> (idx (nq, k) uint64, dist (nq, k), counts (nq,))"""
        data = np.asfortranarray(data, dtype=np.float64)
        queries = np.asfortranarray(queries, dtype=np.float64)
        dim, n = data.shape
        nq = queries.shape[1]
        h = C.c_void_p()
        self._check(lib().m3d_knn_create(self.h, _p(data), C.c_int(dim), C.c_size_t(n), C.byref(h)))
        try:
            idx = np.zeros((max(nq, 1), max(k, 1)), dtype=np.uint64)
            dist = np.zeros((max(nq, 1), max(k, 1)))
            cnt = np.zeros(max(nq, 1), dtype=np.int32)
            self._check(lib().m3d_knn_search(self.h, h, _p(queries), C.c_size_t(nq), C.c_int(k), C.c_double(radius),
                                             _p(idx, C.c_size_t), _p(dist), _p(cnt, C.c_int)))
        finally:
            lib().m3d_knn_free(h)
        return idx[:nq, :k], dist[:nq, :k], cnt[:nq]

    def compute_fpfh(self, xyz, normals, radius, max_nn=100):
        """Open3D ComputeFPFHFeature (m3d_compute_fpfh): (33, n) float64 F-order descriptors, device_ms"""
        xyz = _f64(xyz).reshape(-1, 3)
        nrm = None if normals is None else _f64(normals).reshape(-1, 3)
        n = len(xyz)
        out = np.zeros((max(n, 1), 33))
        ms = C.c_float(0)
        self._check(lib().m3d_compute_fpfh(self.h, _p(xyz), _p(nrm), C.c_size_t(n), C.c_double(radius), C.c_int(max_nn),
                                           _p(out), C.byref(ms)))
        return np.asfortranarray(out[:n].T), float(ms.value)

    def fpfh_features(self, xyz, normals, radius, max_nn=100):
        """m3d_fpfh_create: FPFH descriptors left on the device -> (Features, device_ms)"""
        xyz = _f64(xyz).reshape(-1, 3)
        nrm = None if normals is None else _f64(normals).reshape(-1, 3)
        h = C.c_void_p()
        ms = C.c_float(0)
        self._check(lib().m3d_fpfh_create(self.h, _p(xyz), _p(nrm), C.c_size_t(len(xyz)), C.c_double(radius),
                                          C.c_int(max_nn), C.byref(h), C.byref(ms)))
        return Features(self, h), float(ms.value)

    def upload_features(self, feat):
        """(dim, n) float64 descriptors -> Features on the device"""
        feat = np.asfortranarray(feat, dtype=np.float64)
        dim, n = feat.shape
        h = C.c_void_p()
        self._check(lib().m3d_features_upload(self.h, _p(feat), C.c_int(dim), C.c_size_t(n), C.byref(h)))
        return Features(self, h)

    def match_features(self, fa, fb):
        """m3d_match_features: match_correspondence on two device-resident descriptor sets -> (idx0, idx1, device_ms)"""
        i0 = np.empty(max(fa.n, 1), dtype=np.uint64)
        i1 = np.empty(max(fa.n, 1), dtype=np.uint64)
        n_out = C.c_size_t(0)
        ms = C.c_float(0)
        self._check(lib().m3d_match_features(self.h, fa.handle, fb.handle, _p(i0, C.c_size_t), _p(i1, C.c_size_t),
                                             C.byref(n_out), C.byref(ms)))
        return i0[:n_out.value].copy(), i1[:n_out.value].copy(), float(ms.value)

    def icp_point_to_point(self, src, dst, max_distance, T_init=None, max_iteration=30, relative_fitness=1e-6,
                           relative_rmse=1e-6):
        """Open3D RegistrationICP, point to point (m3d_icp_point_to_point): (T, fitness, inlier_rmse, iterations)"""
        src = _f64(src).reshape(-1, 3)
        dst = _f64(dst).reshape(-1, 3)
        T0 = None if T_init is None else np.ascontiguousarray(T_init, dtype=np.float64).reshape(16)
        T = np.zeros(16)
        fit, rmse, it = C.c_double(0), C.c_double(0), C.c_int(0)
        self._check(lib().m3d_icp_point_to_point(self.h, _p(src), C.c_size_t(len(src)), _p(dst), C.c_size_t(len(dst)),
                                                 C.c_double(max_distance), _p(T0), C.c_int(max_iteration),
                                                 C.c_double(relative_fitness), C.c_double(relative_rmse), _p(T),
                                                 C.byref(fit), C.byref(rmse), C.byref(it)))
        return T.reshape(4, 4).copy(), float(fit.value), float(rmse.value), int(it.value)

    def least_squares_transform(self, src, dst, with_scaling=False):
        src = _f64(src).reshape(-1, 3)
        dst = _f64(dst).reshape(-1, 3)
        if len(src) != len(dst):  # transform_estimation.cpp:52-55 "The number of points pair is not equal"
            raise ValueError("The number of points pair is not equal")
        T = np.zeros(16)
        self._check(lib().m3d_least_squares_transform(self.h, _p(src), _p(dst), C.c_size_t(len(src)),
                                                      C.c_int(1 if with_scaling else 0), _p(T)))
        return T.reshape(4, 4).copy()

    def registration_refit(self, src, dst, idx0, idx1, T, threshold, with_scaling=False):
        """Extension: least-squares (Umeyama) refit of T on its inlier correspondences.
        Returns (T_refit[4,4], n_inliers)."""
        src = _f64(src).reshape(-1, 3)
        dst = _f64(dst).reshape(-1, 3)
        c0 = np.ascontiguousarray(idx0, dtype=np.uint64)
        c1 = np.ascontiguousarray(idx1, dtype=np.uint64)
        if len(c0) != len(c1):
            raise ValueError("correspondence index arrays differ in length")
        Tin = _f64(T).reshape(16)
        Tout = np.zeros(16)
        n_inl = C.c_size_t(0)
        self._check(lib().m3d_registration_refit(self.h, _p(src), C.c_size_t(len(src)), _p(dst), C.c_size_t(len(dst)),
                                                 _p(c0, C.c_size_t), _p(c1, C.c_size_t), C.c_size_t(len(c0)), _p(Tin),
                                                 C.c_double(threshold), C.c_int(1 if with_scaling else 0), _p(Tout),
                                                 C.byref(n_inl)))
        return Tout.reshape(4, 4).copy(), int(n_inl.value)
