"""ctypes wrapper of the CPU oracle (oracle/libm3d_oracle.so).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs.  Never imported by the product
package (misc3d_b200).  Parity status (pinned against the compiled reference for the fit /
segmentation / matching functions, unpinned for the Open3D-defined registration): oracle/m3d_oracle.h.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "libm3d_oracle.so")

PLANE, SPHERE, CYLINDER = 0, 1, 2
KSAMPLE = {PLANE: 3, SPHERE: 4, CYLINDER: 2}
NPARAM = {PLANE: 4, SPHERE: 4, CYLINDER: 7}


class Stats(C.Structure):
    _fields_ = [("best_index", C.c_uint64), ("best_count", C.c_uint64), ("best_rmse", C.c_double),
                ("iterations_run", C.c_uint64), ("stop_index", C.c_uint64), ("found", C.c_int32),
                ("refit_ok", C.c_int32)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


class RegStats(C.Structure):
    _fields_ = [("best_index", C.c_uint64), ("best_count", C.c_uint64), ("best_rmse", C.c_double),
                ("evaluated", C.c_uint64), ("stop_index", C.c_uint64)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


def build(force=False):
    src = [os.path.join(_HERE, f) for f in ("m3d_oracle.cpp", "m3d_oracle_features.cpp", "m3d_oracle.h", "Makefile")]
    if (not force and os.path.exists(_LIB)
            and all(os.path.getmtime(_LIB) >= os.path.getmtime(s) for s in src)):
        return _LIB
    subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _LIB


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB):
            build()
        _lib = C.CDLL(_LIB)
        _lib.orc_distance.restype = C.c_double
        _lib.orc_evaluate.restype = C.c_uint64
    return _lib


def _p(a, t=C.c_double):
    return a.ctypes.data_as(C.POINTER(t)) if a is not None else None


def _f64(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.float64)


def sample_table(seed, n, k, rows):
    out = np.empty((rows, k), dtype=np.uint32)
    lib().orc_sample_table(C.c_uint32(seed & 0xFFFFFFFF), C.c_size_t(n), k, C.c_size_t(rows),
                           _p(out, C.c_uint32))
    return out


def reg_sample_table(seed, m, rows):
    out = np.empty((rows, 3), dtype=np.uint32)
    lib().orc_reg_sample_table(C.c_uint32(seed & 0xFFFFFFFF), C.c_size_t(m), C.c_size_t(rows),
                               _p(out, C.c_uint32))
    return out


def minimal_fit(kind, pts, nrm=None):
    pts, nrm = _f64(pts), _f64(nrm)
    m = np.zeros(7)
    ok = lib().orc_minimal_fit(kind, _p(pts), _p(nrm), _p(m))
    return ok, m[:NPARAM[kind]]


def distance(kind, model, q):
    m = np.zeros(7)
    m[:len(model)] = model
    q = _f64(q)
    return lib().orc_distance(kind, _p(m), _p(q))


def evaluate(kind, xyz, model, thr):
    xyz = _f64(xyz)
    m = np.zeros(7)
    m[:len(model)] = model
    err = C.c_double(0)
    cnt = lib().orc_evaluate(kind, _p(xyz), C.c_size_t(len(xyz)), _p(m), C.c_double(thr),
                             C.byref(err))
    return int(cnt), err.value


def general_fit(kind, xyz, model0=None):
    xyz = _f64(xyz)
    m = np.zeros(7)
    if model0 is not None:
        m[:len(model0)] = model0
    ok = lib().orc_general_fit(kind, _p(xyz), C.c_size_t(len(xyz)), _p(m))
    return ok, m[:NPARAM[kind]]


def ransac_fit(kind, xyz, nrm=None, thr=0.01, max_it=1000, prob=0.9999, seed=1, omp=False,
               faithful=False):
    xyz, nrm = _f64(xyz), _f64(nrm)
    n = len(xyz)
    model = np.zeros(7)
    inl = np.empty(max(n, 1), dtype=np.uint64)
    n_inl = C.c_size_t(0)
    st = Stats()
    if omp:
        rc = lib().orc_ransac_fit_omp(kind, _p(xyz), _p(nrm), C.c_size_t(n), C.c_double(thr),
                                      C.c_size_t(max_it), C.c_double(prob),
                                      C.c_uint32(seed & 0xFFFFFFFF), int(faithful), _p(model),
                                      _p(inl, C.c_size_t), C.byref(n_inl), C.byref(st))
    else:
        rc = lib().orc_ransac_fit(kind, _p(xyz), _p(nrm), C.c_size_t(n), C.c_double(thr),
                                  C.c_size_t(max_it), C.c_double(prob),
                                  C.c_uint32(seed & 0xFFFFFFFF), _p(model), _p(inl, C.c_size_t),
                                  C.byref(n_inl), C.byref(st))
    return rc, model[:NPARAM[kind]].copy(), inl[:n_inl.value].copy(), st.as_dict()


def segment_plane_iterative(xyz, thr, max_it=100, min_ratio=0.05, seed=1, omp=False, cap=256):
    xyz = _f64(xyz)
    n = len(xyz)
    planes = np.zeros((cap, 4))
    labels = np.empty(max(n, 1), dtype=np.uint64)
    npl = C.c_size_t(0)
    rc = lib().orc_segment_plane_iterative(_p(xyz), C.c_size_t(n), C.c_double(thr), int(max_it),
                                           C.c_double(min_ratio), C.c_uint32(seed & 0xFFFFFFFF),
                                           int(omp), _p(planes), C.c_size_t(cap),
                                           _p(labels, C.c_uint64), C.byref(npl))
    return rc, planes[:npl.value].copy(), labels[:n].copy()


def _feat(a):
    """(dim, n) array -> contiguous column-major buffer (n rows of dim)"""
    a = np.asarray(a, dtype=np.float64)
    return np.ascontiguousarray(a.T), a.shape[0], a.shape[1]


def nearest(src, dst):
    s, dim, ns = _feat(src)
    d, dim2, nd = _feat(dst)
    assert dim == dim2
    nn = np.empty(ns, dtype=np.uint64)
    lib().orc_nearest(_p(s), C.c_size_t(ns), _p(d), C.c_size_t(nd), dim, _p(nn, C.c_size_t))
    return nn


def match_correspondence(src, dst):
    s, dim, ns = _feat(src)
    d, dim2, nd = _feat(dst)
    assert dim == dim2
    i0 = np.empty(max(ns, 1), dtype=np.uint64)
    i1 = np.empty(max(ns, 1), dtype=np.uint64)
    n = C.c_size_t(0)
    lib().orc_match_correspondence(_p(s), C.c_size_t(ns), _p(d), C.c_size_t(nd), dim,
                                   _p(i0, C.c_size_t), _p(i1, C.c_size_t), C.byref(n))
    return i0[:n.value].copy(), i1[:n.value].copy()


def umeyama(src, dst, with_scaling=False):
    """src, dst: (n, 3) arrays"""
    src, dst = _f64(src), _f64(dst)
    T = np.zeros(16)
    lib().orc_umeyama(_p(src), _p(dst), C.c_size_t(len(src)), int(with_scaling), _p(T))
    return T.reshape(4, 4)


def ransac_registration(src, dst, c0, c1, thr=0.01, max_iter=100000, edge_thr=0.9,
                        confidence=0.999, seed=1, omp=False):
    src, dst = _f64(src), _f64(dst)
    c0 = np.ascontiguousarray(c0, dtype=np.uint64)
    c1 = np.ascontiguousarray(c1, dtype=np.uint64)
    T = np.zeros(16)
    st = RegStats()
    rc = lib().orc_ransac_registration(_p(src), C.c_size_t(len(src)), _p(dst),
                                       C.c_size_t(len(dst)), _p(c0, C.c_size_t),
                                       _p(c1, C.c_size_t), C.c_size_t(len(c0)), C.c_double(thr),
                                       int(max_iter), C.c_double(edge_thr), C.c_double(confidence),
                                       C.c_uint32(seed & 0xFFFFFFFF), int(omp), _p(T),
                                       C.byref(st))
    return rc, T.reshape(4, 4), st.as_dict()


def fpfh(xyz, nrm, radius, max_nn):
    """Open3D ComputeFPFHFeature(cloud, KDTreeSearchParamHybrid(radius, max_nn)) restated: (33, n) F-order array"""
    xyz, nrm = _f64(xyz), _f64(nrm)
    n = len(xyz)
    out = np.zeros((n, 33))
    rc = lib().orc_fpfh(_p(xyz), _p(nrm), C.c_size_t(n), C.c_double(radius), int(max_nn), _p(out))
    if rc != 0:
        raise RuntimeError("orc_fpfh: the cloud has no normals")
    return np.asfortranarray(out.T)


def hybrid_search_all(xyz, radius, max_nn):
    xyz = _f64(xyz)
    n = len(xyz)
    idx = np.zeros((n, max_nn), dtype=np.uint32)
    d2 = np.zeros((n, max_nn))
    cnt = np.zeros(n, dtype=np.uint32)
    lib().orc_hybrid_search_all(_p(xyz), C.c_size_t(n), C.c_double(radius), int(max_nn), _p(idx, C.c_uint32), _p(d2),
                                _p(cnt, C.c_uint32))
    return idx, d2, cnt


def icp(src, dst, max_dist, T_init=None, max_iter=30, rel_fitness=1e-6, rel_rmse=1e-6):
    """Open3D RegistrationICP (point to point) restated: (T, fitness, inlier_rmse, iterations)"""
    src, dst = _f64(src), _f64(dst)
    T0 = np.eye(4) if T_init is None else np.ascontiguousarray(T_init, dtype=np.float64)
    T = np.zeros(16)
    fit, rmse, it = C.c_double(0), C.c_double(0), C.c_int(0)
    lib().orc_icp(_p(src), C.c_size_t(len(src)), _p(dst), C.c_size_t(len(dst)), C.c_double(max_dist), _p(T0), int(max_iter),
                  C.c_double(rel_fitness), C.c_double(rel_rmse), _p(T), C.byref(fit), C.byref(rmse), C.byref(it))
    return T.reshape(4, 4), fit.value, rmse.value, it.value


def omp_threads():
    return lib().orc_omp_threads()


def use_all_cores():
    """OpenMP threads := the cores this process may run on (launchers like torchrun export OMP_NUM_THREADS=1)"""
    n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    lib().orc_set_threads(int(n))
    return omp_threads()
