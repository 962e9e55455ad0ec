/*
 * m3d_oracle.cpp -- CPU oracle for the Misc3D RANSAC / registration hot path.
 *
 * TEST INFRASTRUCTURE ONLY (see m3d_oracle.h).  Parity: the RANSAC fit,
 * segmentation and matching functions are PINNED bit for bit against the
 * reference's own sources compiled into oracle/_ref (tests/test_reference_pin.py);
 * the Open3D-defined registration part is UNPINNED (see m3d_oracle.h).  Every
 * function below cites the reference lines (relative to /root/reference) it
 * restates.  No code is copied: Eigen/Open3D expressions are re-expressed as
 * explicit scalar IEEE-754 operations in the evaluation order documented in
 * SURVEY.md Appendix A/B/D.  Build: g++ -O3 -fopenmp -ffp-contract=off
 * (no -march=native, no -ffast-math: the reference has no FMA contraction).
 */
#include "m3d_oracle.h"

#include <algorithm>
#include <cfloat>
#include <climits>
#include <cmath>
#include <cstring>
#include <limits>
#include <mutex>
#include <random>
#include <vector>

#ifdef _OPENMP
#include <omp.h>
#endif

namespace {

constexpr double kEps = 1.0e-8; /* ransac.h:14 */

struct V3 {
    double x, y, z;
};
inline V3 ld3(const double *p) { return {p[0], p[1], p[2]}; }
inline V3 sub(V3 a, V3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
/* Eigen cross(): each component mul, mul, sub (Appendix D) */
inline V3 cross(V3 a, V3 b) {
    return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}
/* fixed-size-3 dot / squaredNorm: linear order (Appendix D) */
inline double dot3(V3 a, V3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
inline double norm3(V3 a) { return std::sqrt(dot3(a, a)); }
/* fixed-size-4 dot with SSE2 packets of 2: (v0w0+v2w2)+(v1w1+v3w3) (Appendix D) */
inline double dot4(const double *v, const double *w) {
    return (v[0] * w[0] + v[2] * w[2]) + (v[1] * w[1] + v[3] * w[3]);
}

/* ------------------------------------------------------------------ plane */
/* ransac.h:138-162 */
bool plane_minimal(const double *pts, double *m) {
    const V3 p0 = ld3(pts), p1 = ld3(pts + 3), p2 = ld3(pts + 6);
    const V3 e0 = sub(p1, p0);
    const V3 e1 = sub(p2, p0);
    V3 abc = cross(e0, e1);
    const double norm = norm3(abc);
    if (norm < kEps) return false;
    const double n2 = norm3(abc); /* ransac.h:154 recomputes the norm */
    abc = {abc.x / n2, abc.y / n2, abc.z / n2};
    const double d = -dot3(abc, p0);
    m[0] = abc.x;
    m[1] = abc.y;
    m[2] = abc.z;
    m[3] = d;
    return true;
}
/* ransac.h:215-220 : |w.[q,1]| / ||w[0:3]|| */
inline double plane_distance(const double *w, V3 q) {
    const double num = (w[0] * q.x + w[2] * q.z) + (w[1] * q.y + w[3] * 1.0);
    const double nrm = std::sqrt((w[0] * w[0] + w[1] * w[1]) + w[2] * w[2]);
    return std::fabs(num) / nrm;
}
/* ransac.h:164-213 (covariance summed sequentially; the reference uses an
 * order-nondeterministic OpenMP reduction, Appendix A.7) */
bool plane_general(const double *xyz, size_t n, double *m) {
    if (n < 3) return false;
    double mx = 0, my = 0, mz = 0;
    for (size_t i = 0; i < n; ++i) {
        mx += xyz[3 * i];
        my += xyz[3 * i + 1];
        mz += xyz[3 * i + 2];
    }
    mx /= double(n);
    my /= double(n);
    mz /= double(n);
    double xx = 0, xy = 0, xz = 0, yy = 0, yz = 0, zz = 0;
    for (size_t i = 0; i < n; ++i) {
        const double rx = xyz[3 * i] - mx, ry = xyz[3 * i + 1] - my, rz = xyz[3 * i + 2] - mz;
        xx += rx * rx;
        xy += rx * ry;
        xz += rx * rz;
        yy += ry * ry;
        yz += ry * rz;
        zz += rz * rz;
    }
    const double det_x = yy * zz - yz * yz;
    const double det_y = xx * zz - xz * xz;
    const double det_z = xx * yy - xy * xy;
    V3 abc;
    if (det_x > det_y && det_x > det_z) {
        abc = {det_x, xz * yz - xy * zz, xy * yz - xz * yy};
    } else if (det_y > det_z) {
        abc = {xz * yz - xy * zz, det_y, xy * xz - yz * xx};
    } else {
        abc = {xy * yz - xz * yy, xy * xz - yz * xx, det_z};
    }
    const double norm = norm3(abc);
    if (norm < kEps) return false;
    abc = {abc.x / norm, abc.y / norm, abc.z / norm};
    const V3 mean = {mx, my, mz};
    m[0] = abc.x;
    m[1] = abc.y;
    m[2] = abc.z;
    m[3] = -dot3(abc, mean);
    return true;
}

/* ----------------------------------------------------------------- sphere */
/* Eigen 3.4 Matrix4d::determinant() (Appendix D), row-major m[4][4] */
double det4(const double m[4][4]) {
    auto d2 = [&](int i, int j) { return m[i][0] * m[j][1] - m[j][0] * m[i][1]; };
    auto d3 = [&](int i0, double a, int i1, double b, int i2, double c) {
        return m[i0][2] * a + (-m[i1][2] * b + m[i2][2] * c);
    };
    const double d01 = d2(0, 1), d02 = d2(0, 2), d03 = d2(0, 3), d12 = d2(1, 2), d13 = d2(1, 3),
                 d23 = d2(2, 3);
    const double d3_0 = d3(1, d23, 2, d13, 3, d12);
    const double d3_1 = d3(0, d23, 2, d03, 3, d02);
    const double d3_2 = d3(0, d13, 1, d03, 3, d01);
    const double d3_3 = d3(0, d12, 1, d02, 2, d01);
    return (-m[0][3] * d3_0 + m[1][3] * d3_1) + (-m[2][3] * d3_2 + m[3][3] * d3_3);
}
/* ransac.h:225-234 + 239-294 */
bool sphere_minimal(const double *pts, double *out) {
    double pl[4];
    if (!plane_minimal(pts, pl)) return false;
    if (plane_distance(pl, ld3(pts + 9)) < kEps) return false;
    double m[4][4];
    double sq[4];
    for (int i = 0; i < 4; ++i) sq[i] = dot3(ld3(pts + 3 * i), ld3(pts + 3 * i));
    auto fill = [&](int c0, int c1, int c2, int c3) {
        /* column source codes: 0..2 = x,y,z ; 3 = |p|^2 ; 4 = 1 */
        const int cs[4] = {c0, c1, c2, c3};
        for (int i = 0; i < 4; ++i)
            for (int c = 0; c < 4; ++c)
                m[i][c] = cs[c] == 4 ? 1.0 : (cs[c] == 3 ? sq[i] : pts[3 * i + cs[c]]);
    };
    fill(0, 1, 2, 4);
    const double M11 = det4(m);
    fill(3, 1, 2, 4);
    const double M12 = det4(m);
    fill(3, 0, 2, 4);
    const double M13 = det4(m);
    fill(3, 0, 1, 4);
    const double M14 = det4(m);
    fill(3, 0, 1, 2);
    const double M15 = det4(m);
    const V3 c = {0.5 * (M12 / M11), -0.5 * (M13 / M11), 0.5 * (M14 / M11)};
    const double r = std::sqrt(dot3(c, c) - (M15 / M11));
    out[0] = c.x;
    out[1] = c.y;
    out[2] = c.z;
    out[3] = r;
    return true;
}
/* ransac.h:332-343 */
inline double sphere_distance(const double *w, V3 q) {
    const V3 c = {w[0], w[1], w[2]};
    const double r = w[3];
    const double d = norm3(sub(q, c));
    if (d <= r) return r - d;
    return d - r;
}
/* ransac.h:296-330.  The reference solves min ||A w - b|| with a full-U
 * BDCSVD (O(n^2) memory, Appendix A.6); restated as thin Householder QR of
 * [A | b] -- same least-squares solution for full-rank A. */
bool sphere_general(const double *xyz, size_t n, double *out) {
    if (n < 4) return false;
    std::vector<double> a(n * 5);
    for (size_t i = 0; i < n; ++i) {
        const double x = xyz[3 * i], y = xyz[3 * i + 1], z = xyz[3 * i + 2];
        a[5 * i + 0] = x * 2;
        a[5 * i + 1] = y * 2;
        a[5 * i + 2] = z * 2;
        a[5 * i + 3] = 1.0;
        a[5 * i + 4] = (std::pow(x, 2) + std::pow(y, 2)) + std::pow(z, 2);
    }
    for (int j = 0; j < 4; ++j) {
        long double s = 0;
        for (size_t i = j; i < n; ++i) s += (long double)a[5 * i + j] * a[5 * i + j];
        const double nr = (double)sqrtl(s);
        if (nr == 0.0) continue;
        const double ajj = a[5 * (size_t)j + j];
        const double alpha = ajj > 0 ? -nr : nr;
        /* v = a_j - alpha e_j ; H = I - 2 v v^T / (v^T v) */
        a[5 * (size_t)j + j] = ajj - alpha;
        long double vtv = 0;
        for (size_t i = j; i < n; ++i) vtv += (long double)a[5 * i + j] * a[5 * i + j];
        for (int c = j + 1; c < 5; ++c) {
            long double vta = 0;
            for (size_t i = j; i < n; ++i) vta += (long double)a[5 * i + j] * a[5 * i + c];
            const double f = (double)(2.0L * vta / vtv);
            for (size_t i = j; i < n; ++i) a[5 * i + c] -= f * a[5 * i + j];
        }
        /* store R_jj in place of the (no longer needed) head of v: keep v tail
         * untouched, it is never read again */
        a[5 * (size_t)j + j] = alpha;
    }
    double w[4];
    for (int j = 3; j >= 0; --j) {
        double s = a[5 * (size_t)j + 4];
        for (int c = j + 1; c < 4; ++c) s -= a[5 * (size_t)j + c] * w[c];
        w[j] = s / a[5 * (size_t)j + j];
    }
    const double r = std::sqrt(w[0] * w[0] + w[1] * w[1] + w[2] * w[2] + w[3]);
    out[0] = w[0];
    out[1] = w[1];
    out[2] = w[2];
    out[3] = r;
    return true;
}

/* --------------------------------------------------------------- cylinder */
/* utils.h:314-322 */
inline double point2line(V3 q, V3 p1, V3 p2) {
    const V3 a = sub(q, p1);
    const V3 b = sub(q, p2);
    const V3 c = sub(p2, p1);
    return norm3(cross(a, b)) / norm3(c);
}
/* ransac.h:354-417 (normals presence is checked by the caller) */
bool cylinder_minimal(const double *pts, const double *nrm, double *out) {
    const double *P0 = pts, *P1 = pts + 3;
    /* ransac.h:367-374 as the compiler parses it (Appendix A.4) */
    const bool degenerate = (P0[0] - P1[0] <= std::numeric_limits<double>::epsilon()) &&
                            (std::fabs(P0[1] - P1[1]) <= std::numeric_limits<float>::epsilon()) &&
                            (std::fabs(P0[2] - P1[2]) <= std::numeric_limits<float>::epsilon());
    if (std::fabs((double)degenerate) != 0.0) return false;
    const double p1[4] = {P0[0], P0[1], P0[2], 0};
    const double p2[4] = {P1[0], P1[1], P1[2], 0};
    const double n1[4] = {nrm[0], nrm[1], nrm[2], 0};
    const double n2[4] = {nrm[3], nrm[4], nrm[5], 0};
    double w[4];
    for (int i = 0; i < 4; ++i) w[i] = (n1[i] + p1[i]) - p2[i];
    const double a = dot4(n1, n1);
    const double b = dot4(n1, n2);
    const double c = dot4(n2, n2);
    const double d = dot4(n1, w);
    const double e = dot4(n2, w);
    const double denominator = a * c - b * b;
    double sc, tc;
    if (denominator < 1e-8) {
        sc = 0;
        tc = (b > c ? d / b : e / c);
    } else {
        sc = (b * e - c * d) / denominator;
        tc = (a * e - b * d) / denominator;
    }
    double line_pt[4], line_dir[4];
    for (int i = 0; i < 4; ++i) line_pt[i] = (p1[i] + n1[i]) + sc * n1[i];
    for (int i = 0; i < 4; ++i) line_dir[i] = (p2[i] + tc * n2[i]) - line_pt[i];
    /* Vector4d::normalize(): z = squaredNorm (pairwise), if z > 0 divide by sqrt(z) */
    const double z = dot4(line_dir, line_dir);
    if (z > 0) {
        const double s = std::sqrt(z);
        for (int i = 0; i < 4; ++i) line_dir[i] /= s;
    }
    out[0] = line_pt[0];
    out[1] = line_pt[1];
    out[2] = line_pt[2];
    out[3] = line_dir[0];
    out[4] = line_dir[1];
    out[5] = line_dir[2];
    /* ransac.h:413-414: the DIRECTION is passed as the line's second POINT */
    out[6] = point2line({P0[0], P0[1], P0[2]}, {line_pt[0], line_pt[1], line_pt[2]},
                        {line_dir[0], line_dir[1], line_dir[2]});
    return true;
}
/* ransac.h:435-445 */
inline double cylinder_distance(const double *w, V3 q) {
    const V3 center = {w[0], w[1], w[2]};
    const V3 ref = {w[0] + w[3], w[1] + w[4], w[2] + w[5]};
    const double d = point2line(q, center, ref);
    return std::fabs(d - w[6]);
}

/* ------------------------------------------------------------ dispatchers */
inline int ksample(int kind) { return kind == ORC_PLANE ? 3 : (kind == ORC_SPHERE ? 4 : 2); }
inline int nparam(int kind) { return kind == ORC_CYLINDER ? 7 : 4; }

inline double distance(int kind, const double *m, V3 q) {
    switch (kind) {
        case ORC_PLANE:
            return plane_distance(m, q);
        case ORC_SPHERE:
            return sphere_distance(m, q);
        default:
            return cylinder_distance(m, q);
    }
}
bool minimal_fit(int kind, const double *pts, const double *nrm, double *m) {
    switch (kind) {
        case ORC_PLANE:
            return plane_minimal(pts, m);
        case ORC_SPHERE:
            return sphere_minimal(pts, m);
        default:
            return cylinder_minimal(pts, nrm, m);
    }
}
bool general_fit(int kind, const double *xyz, size_t n, double *m) {
    switch (kind) {
        case ORC_PLANE:
            return plane_general(xyz, n, m);
        case ORC_SPHERE:
            return sphere_general(xyz, n, m);
        default:
            return true; /* ransac.h:427-433 no-op */
    }
}

/* ransac.h:626-654 */
template <int KIND>
uint64_t evaluate_t(const double *xyz, size_t n, const double *m, double thr, double *err) {
    uint64_t cnt = 0;
    double e = 0;
    for (size_t i = 0; i < n; ++i) {
        const double d = distance(KIND, m, ld3(xyz + 3 * i));
        if (d < thr) {
            e += d;
            cnt++;
        }
    }
    *err = e;
    return cnt;
}
uint64_t evaluate(int kind, const double *xyz, size_t n, const double *m, double thr, double *err) {
    switch (kind) {
        case ORC_PLANE:
            return evaluate_t<ORC_PLANE>(xyz, n, m, thr, err);
        case ORC_SPHERE:
            return evaluate_t<ORC_SPHERE>(xyz, n, m, thr, err);
        default:
            return evaluate_t<ORC_CYLINDER>(xyz, n, m, thr, err);
    }
}

/* utils.h:72-97 with an injected seed */
struct Sampler {
    size_t size_;
    std::mt19937 rng_;
    Sampler(size_t n, uint32_t seed) : size_(n), rng_(seed) {}
    void draw(int k, size_t *out) {
        int valid = 0;
        while (valid < k) {
            const size_t idx = rng_() % size_;
            if (std::find(out, out + valid, idx) == out + valid) out[valid++] = idx;
        }
    }
};

/* ransac.h:601-610 : size_t current_iteration = min(log(1-p)/log(1-fit^k), max_it).
 * The implicit double->size_t conversion is UB for -inf / NaN; emulate what
 * gcc/x86-64 does (Appendix A.3, probed Appendix C): values outside [0,2^64)
 * behave as "no limit" (2^63), NaN as 0. */
size_t adaptive_limit(double fitness, int k, double prob, size_t max_it) {
    if (!(fitness < 1.0)) return 0;
    const double v =
        std::min(std::log(1 - prob) / std::log(1 - std::pow(fitness, k)), (double)max_it);
    if (v != v) return 0;
    if (v < 0) return (size_t)1 << 63;
    if (v >= 18446744073709551616.0) return (size_t)1 << 63;
    return (size_t)v;
}

/* SelectByIndex(indices) (Appendix B): ascending order, duplicates collapse. */
void gather_sorted(const double *xyz, const double *nrm, const size_t *idx, int k, double *pts,
                   double *nr) {
    size_t s[4];
    for (int i = 0; i < k; ++i) { /* insertion sort, k <= 4 */
        size_t v = idx[i];
        int j = i;
        while (j > 0 && s[j - 1] > v) {
            s[j] = s[j - 1];
            --j;
        }
        s[j] = v;
    }
    for (int i = 0; i < k; ++i) {
        std::memcpy(pts + 3 * i, xyz + 3 * s[i], 3 * sizeof(double));
        if (nrm) std::memcpy(nr + 3 * i, nrm + 3 * s[i], 3 * sizeof(double));
    }
}
/* the same through an O(N) mask pass, as Open3D really does it (timing only) */
void gather_masked(const double *xyz, const double *nrm, size_t n, const size_t *idx, int k,
                   double *pts, double *nr) {
    std::vector<bool> mask(n, false);
    for (int i = 0; i < k; ++i) mask[idx[i]] = true;
    int o = 0;
    for (size_t i = 0; i < n; ++i) {
        if (mask[i]) {
            std::memcpy(pts + 3 * o, xyz + 3 * i, 3 * sizeof(double));
            if (nrm) std::memcpy(nr + 3 * o, nrm + 3 * i, 3 * sizeof(double));
            ++o;
        }
    }
}

/* ransac.h:534-549 */
int refine(int kind, const double *xyz, size_t n, double thr, double *model, size_t *inl,
           size_t *n_inl) {
    size_t c = 0;
    for (size_t i = 0; i < n; ++i) {
        const double d = distance(kind, model, ld3(xyz + 3 * i));
        if (d < thr) inl[c++] = i;
    }
    *n_inl = c;
    std::vector<double> sel(3 * c);
    for (size_t j = 0; j < c; ++j) std::memcpy(&sel[3 * j], xyz + 3 * inl[j], 3 * sizeof(double));
    return general_fit(kind, sel.data(), c, model) ? 1 : 0;
}

int ransac_fit(int kind, const double *xyz, const double *nrm, size_t n, double thr, size_t max_it,
               double prob, uint32_t seed, int omp_mode, int faithful, double *model, size_t *inl,
               size_t *n_inl, orc_stats *st) {
    const int k = ksample(kind), np = nparam(kind);
    orc_stats s{};
    s.stop_index = max_it;
    *n_inl = 0;
    for (int i = 0; i < 7; ++i) model[i] = 0;
    if (!(prob > 0 && prob <= 1)) return -3;       /* ransac.h:482-487 throws */
    if (kind == ORC_CYLINDER && !nrm) return -1;   /* py_common.cpp:50-52 / ransac.h:356-359 */
    if (n < (size_t)k) return -2;                  /* ransac.h:510-513 throws */

    double best_fit = 0, best_rmse = 0; /* ransac.h:459-460, 519-522 */
    double best_model[7] = {0, 0, 0, 0, 0, 0, 0};
    size_t count = 0;
    size_t cur = std::numeric_limits<size_t>::max();
    Sampler sampler(n, seed);

    if (!omp_mode) {
        for (size_t i = 0; i < max_it; ++i) {
            if (count > cur) { /* ransac.h:573-575 */
                if (s.stop_index == max_it) s.stop_index = i;
                continue;
            }
            size_t idx[4];
            sampler.draw(k, idx);
            double pts[12], nr[12], trial[7];
            gather_sorted(xyz, nrm, idx, k, pts, nr);
            if (!minimal_fit(kind, pts, nr, trial)) continue;
            double err;
            const uint64_t cnt = evaluate(kind, xyz, n, trial, thr, &err);
            double fitness, rmse;
            if (cnt == 0) {
                fitness = 0;
                rmse = 1e+10;
            } else {
                fitness = (double)cnt / (double)n;
                rmse = err / std::sqrt((double)cnt);
            }
            if (fitness > best_fit || (fitness == best_fit && rmse < best_rmse)) {
                best_fit = fitness;
                best_rmse = rmse;
                std::memcpy(best_model, trial, sizeof(double) * np);
                s.best_index = i;
                s.best_count = cnt;
                s.best_rmse = rmse;
                s.found = 1;
                cur = adaptive_limit(best_fit, k, prob, max_it);
            }
            count++;
        }
    } else {
        std::mutex mu;
        volatile size_t *pcount = &count;
        volatile size_t *pcur = &cur;
#pragma omp parallel for schedule(static)
        for (long long i = 0; i < (long long)max_it; ++i) {
            if (*pcount > *pcur) continue;
            size_t idx[4];
            {
                std::lock_guard<std::mutex> g(mu); /* utils.h:83 */
                sampler.draw(k, idx);
            }
            double pts[12], nr[12], trial[7];
            if (faithful)
                gather_masked(xyz, nrm, n, idx, k, pts, nr);
            else
                gather_sorted(xyz, nrm, idx, k, pts, nr);
            if (!minimal_fit(kind, pts, nr, trial)) continue;
            double err;
            const uint64_t cnt = evaluate(kind, xyz, n, trial, thr, &err);
            double fitness, rmse;
            if (cnt == 0) {
                fitness = 0;
                rmse = 1e+10;
            } else {
                fitness = (double)cnt / (double)n;
                rmse = err / std::sqrt((double)cnt);
            }
#pragma omp critical
            {
                if (fitness > best_fit || (fitness == best_fit && rmse < best_rmse)) {
                    best_fit = fitness;
                    best_rmse = rmse;
                    std::memcpy(best_model, trial, sizeof(double) * np);
                    s.best_index = (uint64_t)i;
                    s.best_count = cnt;
                    s.best_rmse = rmse;
                    s.found = 1;
                    cur = adaptive_limit(best_fit, k, prob, max_it);
                }
                count++;
            }
        }
    }
    s.iterations_run = count;
    int ret = 0;
    if (s.found) {
        /* ransac.h:621-623 */
        ret = refine(kind, xyz, n, thr, best_model, inl, n_inl);
        std::memcpy(model, best_model, sizeof(double) * np);
        s.refit_ok = ret;
    }
    /* !found: the reference returns an uninitialised VectorXd (Appendix A.9);
     * the oracle defines zeros + no inliers + false. */
    if (st) *st = s;
    return ret;
}

/* ---------------------------------------------------------- registration */
struct Rot {
    double c, s;
};
/* Eigen JacobiRotation::makeJacobi(x, y, z) for the real symmetric 2x2 [x y; y z] */
bool make_jacobi(double x, double y, double z, Rot &r) {
    const double deno = 2.0 * std::fabs(y);
    if (deno < DBL_MIN) {
        r.c = 1;
        r.s = 0;
        return false;
    }
    const double tau = (x - z) / deno;
    const double w = std::sqrt(tau * tau + 1.0);
    const double t = tau > 0 ? 1.0 / (tau + w) : 1.0 / (tau - w);
    const double sign_t = t > 0 ? 1.0 : -1.0;
    const double n = 1.0 / std::sqrt(t * t + 1.0);
    r.s = -sign_t * (y / std::fabs(y)) * std::fabs(t) * n;
    r.c = n;
    return true;
}
/* rows p,q of a 3x3: x' = c x + s y ; y' = -s x + c y  (B = J B) */
void rot_left(double w[3][3], int p, int q, Rot j) {
    for (int i = 0; i < 3; ++i) {
        const double x = w[p][i], y = w[q][i];
        w[p][i] = j.c * x + j.s * y;
        w[q][i] = -j.s * x + j.c * y;
    }
}
/* columns p,q: B = B J, J = [c s; -s c] */
void rot_right(double w[3][3], int p, int q, Rot j) {
    for (int i = 0; i < 3; ++i) {
        const double x = w[i][p], y = w[i][q];
        w[i][p] = j.c * x - j.s * y;
        w[i][q] = j.s * x + j.c * y;
    }
}
/* Eigen JacobiSVD<Matrix3d>(FullU|FullV): two-sided Jacobi, A = U diag(sv) V^T */
void jacobi_svd3(const double a[3][3], double U[3][3], double sv[3], double V[3][3]) {
    const double precision = 2.0 * DBL_EPSILON;
    const double consider_as_zero = DBL_MIN;
    double scale = 0;
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) scale = std::max(scale, std::fabs(a[i][j]));
    if (scale == 0.0) scale = 1.0;
    double w[3][3];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
            w[i][j] = a[i][j] / scale;
            U[i][j] = V[i][j] = (i == j) ? 1.0 : 0.0;
        }
    double max_diag = std::max(std::fabs(w[0][0]), std::max(std::fabs(w[1][1]), std::fabs(w[2][2])));
    bool finished = false;
    int guard = 0;
    while (!finished && guard++ < 100) {
        finished = true;
        for (int p = 1; p < 3; ++p) {
            for (int q = 0; q < p; ++q) {
                const double threshold = std::max(consider_as_zero, precision * max_diag);
                if (std::fabs(w[p][q]) > threshold || std::fabs(w[q][p]) > threshold) {
                    finished = false;
                    /* real_2x2_jacobi_svd */
                    double m00 = w[p][p], m01 = w[p][q], m10 = w[q][p], m11 = w[q][q];
                    Rot rot1;
                    const double t = m00 + m11;
                    const double d = m10 - m01;
                    if (std::fabs(d) < DBL_MIN) {
                        rot1.s = 0;
                        rot1.c = 1;
                    } else {
                        const double u = t / d;
                        const double tmp = std::sqrt(1.0 + u * u);
                        rot1.s = 1.0 / tmp;
                        rot1.c = u / tmp;
                    }
                    /* m.applyOnTheLeft(0,1,rot1) */
                    const double n00 = rot1.c * m00 + rot1.s * m10;
                    const double n01 = rot1.c * m01 + rot1.s * m11;
                    const double n11 = -rot1.s * m01 + rot1.c * m11;
                    Rot jr;
                    make_jacobi(n00, n01, n11, jr);
                    /* j_left = rot1 * j_right^T */
                    Rot jl;
                    jl.c = rot1.c * jr.c + rot1.s * jr.s;
                    jl.s = -rot1.c * jr.s + rot1.s * jr.c;
                    rot_left(w, p, q, jl);
                    rot_right(U, p, q, Rot{jl.c, -jl.s});
                    rot_right(w, p, q, jr);
                    rot_right(V, p, q, jr);
                    max_diag = std::max(max_diag,
                                        std::max(std::fabs(w[p][p]), std::fabs(w[q][q])));
                }
            }
        }
    }
    for (int i = 0; i < 3; ++i) {
        const double aa = std::fabs(w[i][i]);
        sv[i] = aa;
        if (aa != 0.0 && w[i][i] < 0)
            for (int r = 0; r < 3; ++r) U[r][i] = -U[r][i];
    }
    for (int i = 0; i < 3; ++i) sv[i] *= scale;
    for (int i = 0; i < 3; ++i) {
        int pos = i;
        for (int j = i + 1; j < 3; ++j)
            if (sv[j] > sv[pos]) pos = j;
        if (sv[pos] == 0.0) break;
        if (pos != i) {
            std::swap(sv[i], sv[pos]);
            for (int r = 0; r < 3; ++r) {
                std::swap(U[r][i], U[r][pos]);
                std::swap(V[r][i], V[r][pos]);
            }
        }
    }
}
double det3(const double m[3][3]) {
    return m[0][0] * (m[1][1] * m[2][2] - m[1][2] * m[2][1]) -
           m[0][1] * (m[1][0] * m[2][2] - m[1][2] * m[2][0]) +
           m[0][2] * (m[1][0] * m[2][1] - m[1][1] * m[2][0]);
}
/* Eigen::umeyama (Appendix D).  sp/dp: arrays of n pointers to xyz triples. */
void umeyama_pts(const double *const *sp, const double *const *dp, size_t n, bool with_scaling,
                 double T[16]) {
    const double one_over_n = 1.0 / (double)n;
    double sm[3] = {0, 0, 0}, dm[3] = {0, 0, 0};
    for (size_t i = 0; i < n; ++i)
        for (int r = 0; r < 3; ++r) {
            sm[r] += sp[i][r];
            dm[r] += dp[i][r];
        }
    for (int r = 0; r < 3; ++r) {
        sm[r] *= one_over_n;
        dm[r] *= one_over_n;
    }
    double sigma[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
    double src_var = 0;
    for (size_t i = 0; i < n; ++i) {
        double sd[3], dd[3];
        for (int r = 0; r < 3; ++r) {
            sd[r] = sp[i][r] - sm[r];
            dd[r] = dp[i][r] - dm[r];
        }
        src_var += (sd[0] * sd[0] + sd[1] * sd[1]) + sd[2] * sd[2];
        for (int r = 0; r < 3; ++r)
            for (int c = 0; c < 3; ++c) sigma[r][c] += (one_over_n * dd[r]) * sd[c];
    }
    src_var *= one_over_n;
    double U[3][3], V[3][3], sv[3];
    jacobi_svd3(sigma, U, sv, V);
    double S[3] = {1, 1, 1};
    if (det3(U) * det3(V) < 0) S[2] = -1;
    double R[3][3];
    for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c)
            R[r][c] = ((U[r][0] * S[0]) * V[c][0] + (U[r][1] * S[1]) * V[c][1]) +
                      (U[r][2] * S[2]) * V[c][2];
    double cs = 1.0;
    if (with_scaling) cs = 1.0 / src_var * ((sv[0] * S[0] + sv[1] * S[1]) + sv[2] * S[2]);
    for (int r = 0; r < 3; ++r) {
        const double rs = (R[r][0] * sm[0] + R[r][1] * sm[1]) + R[r][2] * sm[2];
        T[4 * r + 3] = with_scaling ? dm[r] - cs * rs : dm[r] - rs;
        for (int c = 0; c < 3; ++c) T[4 * r + c] = with_scaling ? R[r][c] * cs : R[r][c];
    }
    T[12] = T[13] = T[14] = 0;
    T[15] = 1;
}
/* Open3D PointCloud::Transform on one point (Appendix B): T*[p;1], sequential inner order */
inline V3 xform(const double *T, V3 p) {
    V3 o;
    o.x = ((T[0] * p.x + T[1] * p.y) + T[2] * p.z) + T[3] * 1.0;
    o.y = ((T[4] * p.x + T[5] * p.y) + T[6] * p.z) + T[7] * 1.0;
    o.z = ((T[8] * p.x + T[9] * p.y) + T[10] * p.z) + T[11] * 1.0;
    return o;
}

struct RegHyp {
    bool pass;
    double T[16];
};
/* one RANSAC hypothesis: estimate + the two checkers (Appendix B) */
RegHyp reg_hypothesis(const double *sx, const double *dx, const size_t *c0, const size_t *c1,
                      const uint32_t *pick, double thr, double edge_thr) {
    RegHyp h;
    const double *sp[3], *dp[3];
    for (int j = 0; j < 3; ++j) {
        sp[j] = sx + 3 * c0[pick[j]];
        dp[j] = dx + 3 * c1[pick[j]];
    }
    umeyama_pts(sp, dp, 3, false, h.T);
    h.pass = true;
    /* CorrespondenceCheckerBasedOnEdgeLength */
    for (int i = 0; i < 3 && h.pass; ++i)
        for (int j = i + 1; j < 3; ++j) {
            const double ds = norm3(sub(ld3(sp[i]), ld3(sp[j])));
            const double dt = norm3(sub(ld3(dp[i]), ld3(dp[j])));
            if (ds < dt * edge_thr || dt < ds * edge_thr) {
                h.pass = false;
                break;
            }
        }
    if (!h.pass) return h;
    /* CorrespondenceCheckerBasedOnDistance */
    for (int j = 0; j < 3; ++j) {
        const V3 pt = xform(h.T, ld3(sp[j]));
        if (norm3(sub(ld3(dp[j]), pt)) > thr) {
            h.pass = false;
            break;
        }
    }
    return h;
}
/* EvaluateRANSACBasedOnCorrespondence */
uint64_t reg_evaluate(const double *sx, const double *dx, const size_t *c0, const size_t *c1,
                      size_t m, const double *T, double thr, double *err2) {
    const double max_dis2 = thr * thr;
    uint64_t good = 0;
    double e2 = 0;
    for (size_t i = 0; i < m; ++i) {
        const V3 p = xform(T, ld3(sx + 3 * c0[i]));
        const V3 df = sub(p, ld3(dx + 3 * c1[i]));
        const double dis2 = dot3(df, df);
        if (dis2 < max_dis2) {
            good++;
            e2 += dis2;
        }
    }
    *err2 = e2;
    return good;
}
/* est_k update on improvement (Appendix B); (int)ceil of a non-finite value is
 * UB in the reference -- emulated as INT_MIN, which is what x86-64 yields. */
int reg_update_limit(uint64_t good, size_t m, double confidence, int est_k) {
    const double ratio = (double)good / (double)m;
    const double est = std::log(1.0 - confidence) / std::log(1.0 - std::pow(ratio, 3));
    if (est < (double)est_k) {
        const double c = std::ceil(est);
        if (c >= -2147483648.0 && c <= 2147483647.0) return (int)c;
        return INT_MIN;
    }
    return est_k;
}

}  // namespace

/* ================================================================= C API */
extern "C" {

void orc_sample_table(uint32_t seed, size_t n, int k, size_t rows, uint32_t *out) {
    Sampler s(n, seed);
    size_t idx[8];
    for (size_t r = 0; r < rows; ++r) {
        s.draw(k, idx);
        for (int j = 0; j < k; ++j) out[r * k + j] = (uint32_t)idx[j];
    }
}

int orc_minimal_fit(int kind, const double *pts, const double *nrm, double *model) {
    if (kind == ORC_CYLINDER && !nrm) return -1;
    return minimal_fit(kind, pts, nrm, model) ? 1 : 0;
}
double orc_distance(int kind, const double *model, const double *q) {
    return distance(kind, model, ld3(q));
}
uint64_t orc_evaluate(int kind, const double *xyz, size_t n, const double *model, double thr,
                      double *err) {
    double e;
    const uint64_t c = evaluate(kind, xyz, n, model, thr, &e);
    if (err) *err = e;
    return c;
}
int orc_general_fit(int kind, const double *xyz, size_t n, double *model) {
    return general_fit(kind, xyz, n, model) ? 1 : 0;
}

int orc_ransac_fit(int kind, const double *xyz, const double *nrm, size_t n, double thr,
                   size_t max_it, double prob, uint32_t seed, double *model, size_t *inl,
                   size_t *n_inl, orc_stats *st) {
    return ransac_fit(kind, xyz, nrm, n, thr, max_it, prob, seed, 0, 0, model, inl, n_inl, st);
}
int orc_ransac_fit_omp(int kind, const double *xyz, const double *nrm, size_t n, double thr,
                       size_t max_it, double prob, uint32_t seed, int faithful, double *model,
                       size_t *inl, size_t *n_inl, orc_stats *st) {
    return ransac_fit(kind, xyz, nrm, n, thr, max_it, prob, seed, 1, faithful, model, inl, n_inl,
                      st);
}

/* iterative_plane_segmentation.cpp:7-39 */
int orc_segment_plane_iterative(const double *xyz, size_t n, double thr, int max_it,
                                double min_ratio, uint32_t seed, int use_omp, double *planes,
                                size_t cap_planes, uint64_t *labels, size_t *n_planes) {
    *n_planes = 0;
    for (size_t i = 0; i < n; ++i) labels[i] = UINT64_MAX;
    if (n < 3) return 0; /* :14-17 warning + empty result */
    std::vector<double> cur(xyz, xyz + 3 * n);
    std::vector<uint64_t> orig(n);
    for (size_t i = 0; i < n; ++i) orig[i] = i;
    std::vector<size_t> inl(n);
    size_t count = 0;
    const size_t target = (size_t)((1 - min_ratio) * (double)n); /* :28 */
    uint32_t round = 0;
    while (count < target) {
        const size_t m = orig.size();
        if (m < 3) return -2; /* ransac.h:510-513 throws out of the loop */
        if (*n_planes >= cap_planes) return -4;
        double model[7];
        size_t n_inl = 0;
        orc_stats st;
        /* a fresh sampler per FitModel (ransac.h:570): seed_round = seed + round */
        ransac_fit(ORC_PLANE, cur.data(), nullptr, m, thr, (size_t)max_it, 0.9999, seed + round,
                   use_omp, 0, model, inl.data(), &n_inl, &st);
        if (n_inl == 0) return -3; /* reference would loop forever (Appendix A.12) */
        const size_t pid = *n_planes;
        std::memcpy(planes + 4 * pid, model, 4 * sizeof(double));
        /* SelectByIndex(inliers) / SelectByIndex(inliers, true): stable */
        std::vector<double> rest;
        std::vector<uint64_t> rest_orig;
        rest.reserve(3 * (m - n_inl));
        rest_orig.reserve(m - n_inl);
        size_t j = 0;
        for (size_t i = 0; i < m; ++i) {
            if (j < n_inl && inl[j] == i) {
                labels[orig[i]] = pid;
                ++j;
            } else {
                rest.insert(rest.end(), &cur[3 * i], &cur[3 * i] + 3);
                rest_orig.push_back(orig[i]);
            }
        }
        cur.swap(rest);
        orig.swap(rest_orig);
        (*n_planes)++;
        count += n_inl;
        round++;
    }
    return 0;
}

/* nanoflann L2_Adaptor::evalMetric accumulation order (Open3D KDTreeFlann's
 * metric; restated from the published nanoflann source): groups of four,
 * then the remainder one by one. */
static inline double l2_groups4(const double *a, const double *b, int dim) {
    double result = 0;
    int d = 0;
    for (; d + 3 < dim; d += 4) {
        const double d0 = a[d] - b[d], d1 = a[d + 1] - b[d + 1], d2 = a[d + 2] - b[d + 2],
                     d3 = a[d + 3] - b[d + 3];
        result += d0 * d0 + d1 * d1 + d2 * d2 + d3 * d3;
    }
    for (; d < dim; ++d) {
        const double d0 = a[d] - b[d];
        result += d0 * d0;
    }
    return result;
}

/* correspondence_matching.cpp:13-44 (FLANN branch = exact 1-NN); ties -> lowest index */
void orc_nearest(const double *src, size_t ns, const double *dst, size_t nd, int dim, size_t *nn) {
#pragma omp parallel for schedule(static)
    for (long long i = 0; i < (long long)ns; ++i) {
        const double *a = src + (size_t)i * dim;
        double best = std::numeric_limits<double>::infinity();
        size_t bj = 0;
        for (size_t j = 0; j < nd; ++j) {
            const double d = l2_groups4(a, dst + j * dim, dim);
            if (d < best) {
                best = d;
                bj = j;
            }
        }
        nn[i] = bj;
    }
}

/* correspondence_matching.cpp:52-84 */
int orc_match_correspondence(const double *src, size_t ns, const double *dst, size_t nd, int dim,
                             size_t *idx0, size_t *idx1, size_t *n_out) {
    *n_out = 0;
    if (ns == 0 || nd == 0) return 0;
    std::vector<size_t> nn01(ns), nn10(nd);
    orc_nearest(src, ns, dst, nd, dim, nn01.data());
    orc_nearest(dst, nd, src, ns, dim, nn10.data());
    size_t c = 0;
    for (size_t i = 0; i < ns; ++i) { /* :73-78 */
        if (nn10[nn01[i]] == i) {
            idx0[c] = i;
            idx1[c] = nn01[i];
            ++c;
        }
    }
    *n_out = c;
    return 0;
}

void orc_umeyama(const double *src, const double *dst, size_t n, int with_scaling, double *T) {
    std::vector<const double *> sp(n), dp(n);
    for (size_t i = 0; i < n; ++i) {
        sp[i] = src + 3 * i;
        dp[i] = dst + 3 * i;
    }
    umeyama_pts(sp.data(), dp.data(), n, with_scaling != 0, T);
}

void orc_reg_sample_table(uint32_t seed, size_t m, size_t rows, uint32_t *out) {
    std::mt19937 rng(seed);
    std::uniform_int_distribution<int> dist(0, (int)m - 1);
    for (size_t r = 0; r < rows * 3; ++r) out[r] = (uint32_t)dist(rng);
}

int orc_ransac_registration(const double *sx, size_t ns, const double *dx, size_t nd,
                            const size_t *c0, const size_t *c1, size_t m, double thr, int max_iter,
                            double edge_thr, double confidence, uint32_t seed, int use_omp,
                            double *T_out, orc_reg_stats *st) {
    static const double I4[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};
    std::memcpy(T_out, I4, sizeof(I4));
    orc_reg_stats s{};
    s.stop_index = max_iter > 0 ? (uint64_t)max_iter : 0;
    if (st) *st = s;
    if (ns < 3 || nd < 3) return -2; /* transform_estimation.cpp:130-133 throws */
    if (m < 3 || !(thr > 0.0)) return 0; /* Open3D returns the default result */
    double best_fit = 0, best_rmse = 0;
    int est_k = max_iter;

    if (!use_omp) {
        std::mt19937 rng(seed);
        std::uniform_int_distribution<int> dist(0, (int)m - 1);
        for (int itr = 0; itr < max_iter; ++itr) {
            if (!(itr < est_k)) {
                if (s.stop_index == (uint64_t)max_iter) s.stop_index = (uint64_t)itr;
                continue;
            }
            uint32_t pick[3];
            for (int j = 0; j < 3; ++j) pick[j] = (uint32_t)dist(rng);
            const RegHyp h = reg_hypothesis(sx, dx, c0, c1, pick, thr, edge_thr);
            if (!h.pass) continue;
            s.evaluated++;
            double e2;
            const uint64_t good = reg_evaluate(sx, dx, c0, c1, m, h.T, thr, &e2);
            double fitness = 0, rmse = 0;
            if (good) {
                fitness = (double)good / (double)m;
                rmse = std::sqrt(e2 / (double)good);
            }
            if (fitness > best_fit || (fitness == best_fit && rmse < best_rmse)) {
                best_fit = fitness;
                best_rmse = rmse;
                std::memcpy(T_out, h.T, sizeof(h.T));
                s.best_index = (uint64_t)itr;
                s.best_count = good;
                s.best_rmse = rmse;
                est_k = reg_update_limit(good, m, confidence, est_k);
            }
        }
    } else {
        /* timing variant: per-thread generator and best, merged at the end (Appendix B) */
        volatile int est_k_global = max_iter;
        uint64_t evaluated = 0;
#pragma omp parallel reduction(+ : evaluated)
        {
            int tid = 0;
#ifdef _OPENMP
            tid = omp_get_thread_num();
#endif
            std::mt19937 rng(seed + 7919u * (uint32_t)tid);
            std::uniform_int_distribution<int> dist(0, (int)m - 1);
            double lfit = 0, lrmse = 0, lT[16];
            uint64_t lidx = 0, lcnt = 0;
            std::memcpy(lT, I4, sizeof(I4));
            int est_k_local = max_iter;
#pragma omp for nowait
            for (int itr = 0; itr < max_iter; ++itr) {
                if (!(itr < est_k_global)) continue;
                uint32_t pick[3];
                for (int j = 0; j < 3; ++j) pick[j] = (uint32_t)dist(rng);
                const RegHyp h = reg_hypothesis(sx, dx, c0, c1, pick, thr, edge_thr);
                if (!h.pass) continue;
                evaluated++;
                double e2;
                const uint64_t good = reg_evaluate(sx, dx, c0, c1, m, h.T, thr, &e2);
                double fitness = 0, rmse = 0;
                if (good) {
                    fitness = (double)good / (double)m;
                    rmse = std::sqrt(e2 / (double)good);
                }
                if (fitness > lfit || (fitness == lfit && rmse < lrmse)) {
                    lfit = fitness;
                    lrmse = rmse;
                    lidx = (uint64_t)itr;
                    lcnt = good;
                    std::memcpy(lT, h.T, sizeof(lT));
                    est_k_local = reg_update_limit(good, m, confidence, est_k_local);
#pragma omp critical(estk)
                    if (est_k_local < est_k_global) est_k_global = est_k_local;
                }
            }
#pragma omp critical(merge)
            if (lfit > best_fit || (lfit == best_fit && lrmse < best_rmse)) {
                best_fit = lfit;
                best_rmse = lrmse;
                std::memcpy(T_out, lT, sizeof(lT));
                s.best_index = lidx;
                s.best_count = lcnt;
                s.best_rmse = lrmse;
            }
        }
        s.evaluated = evaluated;
    }
    if (st) *st = s;
    return 1;
}

/* launchers such as torchrun export OMP_NUM_THREADS=1; the timing legs ask for the host's cores explicitly */
void orc_set_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}
int orc_omp_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

} /* extern "C" */
