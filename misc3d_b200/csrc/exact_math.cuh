/*
 * exact_math.cuh -- contraction-free fp64 device arithmetic in the reference's evaluation order.
 *
 * Every function here is the device restatement of one reference routine (paths relative to the
 * reference tree).  The reference is compiled by g++ -O3 for x86-64 without -mfma
 * (CMakeLists.txt:16), i.e. every product and every sum is rounded separately; nvcc contracts
 * a*b+c into DFMA by default, so all arithmetic below goes through __dmul_rn / __dadd_rn /
 * __dsub_rn / __ddiv_rn / __dsqrt_rn, which are IEEE-754 correctly rounded and never contracted.
 * Eigen evaluation orders follow SURVEY.md Appendix D.
 */
#pragma once
#include <cfloat>
#include <cstdint>

namespace m3d {

constexpr int kPlane = 0, kSphere = 1, kCylinder = 2;
constexpr double kEps = 1.0e-8; /* EPS, ransac.h:14 */

__host__ __device__ constexpr int sample_size(int kind) { /* ransac.h:136, 237, 352 */
    return kind == kPlane ? 3 : (kind == kSphere ? 4 : 2);
}
__host__ __device__ constexpr int param_count(int kind) { return kind == kCylinder ? 7 : 4; }

namespace ex {

__device__ __forceinline__ double mul(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double add(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double sub(double a, double b) { return __dsub_rn(a, b); }
__device__ __forceinline__ double div(double a, double b) { return __ddiv_rn(a, b); }
__device__ __forceinline__ double sqrt_(double a) { return __dsqrt_rn(a); }

struct V3 {
    double x, y, z;
};
__device__ __forceinline__ V3 ld3(const double *p) { return {p[0], p[1], p[2]}; }
__device__ __forceinline__ V3 sub3(V3 a, V3 b) { return {sub(a.x, b.x), sub(a.y, b.y), sub(a.z, b.z)}; }
/* Eigen cross(): each component mul, mul, sub */
__device__ __forceinline__ V3 cross(V3 a, V3 b) {
    return {sub(mul(a.y, b.z), mul(a.z, b.y)), sub(mul(a.z, b.x), mul(a.x, b.z)),
            sub(mul(a.x, b.y), mul(a.y, b.x))};
}
/* fixed-size-3 dot / squaredNorm: ((a0*b0 + a1*b1) + a2*b2) */
__device__ __forceinline__ double dot3(V3 a, V3 b) {
    return add(add(mul(a.x, b.x), mul(a.y, b.y)), mul(a.z, b.z));
}
__device__ __forceinline__ double norm3(V3 a) { return sqrt_(dot3(a, a)); }
/* fixed-size-4 dot, SSE2 packets of two: (v0*w0 + v2*w2) + (v1*w1 + v3*w3) */
__device__ __forceinline__ double dot4(const double *v, const double *w) {
    return add(add(mul(v[0], w[0]), mul(v[2], w[2])), add(mul(v[1], w[1]), mul(v[3], w[3])));
}

/* ---- PlaneEstimator::MinimalFit, ransac.h:138-162 (pts: 3 points, ascending index order) */
__device__ inline bool plane_minimal(const double *pts, double *m) {
    const V3 p0 = ld3(pts), p1 = ld3(pts + 3), p2 = ld3(pts + 6);
    V3 abc = cross(sub3(p1, p0), sub3(p2, p0));
    const double nrm = norm3(abc);
    if (nrm < kEps) return false;
    abc = {div(abc.x, nrm), div(abc.y, nrm), div(abc.z, nrm)}; /* :154 recomputes the same norm */
    m[0] = abc.x;
    m[1] = abc.y;
    m[2] = abc.z;
    m[3] = -dot3(abc, p0);
    return true;
}
/* Plane::CalcPointToModelDistance, ransac.h:215-220: |w.[q,1]| / ||w[0:3]|| */
__device__ __forceinline__ double plane_norm(const double *w) {
    return sqrt_(add(add(mul(w[0], w[0]), mul(w[1], w[1])), mul(w[2], w[2])));
}
__device__ __forceinline__ double plane_distance(const double *w, double nrm, V3 q) {
    const double num = add(add(mul(w[0], q.x), mul(w[2], q.z)), add(mul(w[1], q.y), mul(w[3], 1.0)));
    return div(fabs(num), nrm);
}

/* ---- Eigen 3.4 Matrix4d::determinant() on the matrix whose columns are c0..c3 */
__device__ inline double det4(const double *c0, const double *c1, const double *c2, const double *c3) {
#define M3D_D2(i, j) sub(mul(c0[i], c1[j]), mul(c0[j], c1[i]))
#define M3D_D3(i0, a, i1, b, i2, c) add(mul(c2[i0], a), add(mul(-c2[i1], b), mul(c2[i2], c)))
    const double d01 = M3D_D2(0, 1), d02 = M3D_D2(0, 2), d03 = M3D_D2(0, 3), d12 = M3D_D2(1, 2),
                 d13 = M3D_D2(1, 3), d23 = M3D_D2(2, 3);
    const double d3_0 = M3D_D3(1, d23, 2, d13, 3, d12);
    const double d3_1 = M3D_D3(0, d23, 2, d03, 3, d02);
    const double d3_2 = M3D_D3(0, d13, 1, d03, 3, d01);
    const double d3_3 = M3D_D3(0, d12, 1, d02, 2, d01);
#undef M3D_D2
#undef M3D_D3
    return add(add(mul(-c3[0], d3_0), mul(c3[1], d3_1)), add(mul(-c3[2], d3_2), mul(c3[3], d3_3)));
}
/* ---- SphereEstimator::ValidationCheck + MinimalFit, ransac.h:225-234, 239-294 (4 points) */
__device__ inline bool sphere_minimal(const double *pts, double *out) {
    double pl[4];
    if (!plane_minimal(pts, pl)) return false;
    if (plane_distance(pl, plane_norm(pl), ld3(pts + 9)) < kEps) return false;
    double cx[4], cy[4], cz[4], sq[4], one[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        cx[i] = pts[3 * i];
        cy[i] = pts[3 * i + 1];
        cz[i] = pts[3 * i + 2];
        sq[i] = dot3(ld3(pts + 3 * i), ld3(pts + 3 * i));
        one[i] = 1.0;
    }
    const double M11 = det4(cx, cy, cz, one);
    const double M12 = det4(sq, cy, cz, one);
    const double M13 = det4(sq, cx, cz, one);
    const double M14 = det4(sq, cx, cy, one);
    const double M15 = det4(sq, cx, cy, cz);
    const V3 c = {mul(0.5, div(M12, M11)), mul(-0.5, div(M13, M11)), mul(0.5, div(M14, M11))};
    out[0] = c.x;
    out[1] = c.y;
    out[2] = c.z;
    out[3] = sqrt_(sub(dot3(c, c), div(M15, M11)));
    return true;
}
/* Sphere::CalcPointToModelDistance, ransac.h:332-343 */
__device__ __forceinline__ double sphere_distance(const double *w, V3 q) {
    const V3 c = {w[0], w[1], w[2]};
    const double d = norm3(sub3(q, c));
    return (d <= w[3]) ? sub(w[3], d) : sub(d, w[3]);
}

/* ---- CalcPoint2LineDistance, utils.h:314-322 */
__device__ __forceinline__ double point2line(V3 q, V3 p1, V3 p2) {
    return div(norm3(cross(sub3(q, p1), sub3(q, p2))), norm3(sub3(p2, p1)));
}
/* ---- CylinderEstimator::MinimalFit, ransac.h:354-417 (2 points + 2 normals) */
__device__ inline bool cylinder_minimal(const double *pts, const double *nrm, double *out) {
    const double *P0 = pts, *P1 = pts + 3;
    /* ransac.h:367-374 as the compiler parses it: signed x test with DBL_EPSILON, float
     * epsilons for y and z */
    const bool degenerate = (sub(P0[0], P1[0]) <= DBL_EPSILON) &&
                            (fabs(sub(P0[1], P1[1])) <= (double)FLT_EPSILON) &&
                            (fabs(sub(P0[2], P1[2])) <= (double)FLT_EPSILON);
    if (degenerate) return false;
    const double p1[4] = {P0[0], P0[1], P0[2], 0.0};
    const double p2[4] = {P1[0], P1[1], P1[2], 0.0};
    const double n1[4] = {nrm[0], nrm[1], nrm[2], 0.0};
    const double n2[4] = {nrm[3], nrm[4], nrm[5], 0.0};
    double w[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) w[i] = sub(add(n1[i], p1[i]), p2[i]);
    const double a = dot4(n1, n1), b = dot4(n1, n2), c = dot4(n2, n2), d = dot4(n1, w),
                 e = dot4(n2, w);
    const double den = sub(mul(a, c), mul(b, b));
    double sc, tc;
    if (den < 1e-8) { /* :393-400 */
        sc = 0.0;
        tc = (b > c) ? div(d, b) : div(e, c);
    } else {
        sc = div(sub(mul(b, e), mul(c, d)), den);
        tc = div(sub(mul(a, e), mul(b, d)), den);
    }
    double lp[4], ld[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) lp[i] = add(add(p1[i], n1[i]), mul(sc, n1[i])); /* :402 (sic) */
#pragma unroll
    for (int i = 0; i < 4; ++i) ld[i] = sub(add(p2[i], mul(tc, n2[i])), lp[i]);
    const double z = dot4(ld, ld); /* Vector4d::normalize() */
    if (z > 0) {
        const double s = sqrt_(z);
#pragma unroll
        for (int i = 0; i < 4; ++i) ld[i] = div(ld[i], s);
    }
    out[0] = lp[0];
    out[1] = lp[1];
    out[2] = lp[2];
    out[3] = ld[0];
    out[4] = ld[1];
    out[5] = ld[2];
    /* :413-414 passes the DIRECTION as the second point of the line */
    out[6] = point2line({P0[0], P0[1], P0[2]}, {lp[0], lp[1], lp[2]}, {ld[0], ld[1], ld[2]});
    return true;
}
/* Cylinder::CalcPointToModelDistance, ransac.h:435-445 */
__device__ __forceinline__ double cylinder_distance(const double *w, V3 q) {
    const V3 center = {w[0], w[1], w[2]};
    const V3 ref = {add(w[0], w[3]), add(w[1], w[4]), add(w[2], w[5])};
    return fabs(sub(point2line(q, center, ref), w[6]));
}

template <int KIND>
__device__ __forceinline__ bool minimal_fit(const double *pts, const double *nrm, double *m) {
    if (KIND == kPlane) return plane_minimal(pts, m);
    if (KIND == kSphere) return sphere_minimal(pts, m);
    return cylinder_minimal(pts, nrm, m);
}

/* Distance functor with the per-model invariants hoisted (the plane norm is recomputed per point
 * by the reference, ransac.h:219, but is the same bits every time). */
template <int KIND>
struct Dist {
    double w[7];
    double nrm;
    __device__ __forceinline__ void set(const double *m) {
#pragma unroll
        for (int i = 0; i < param_count(KIND); ++i) w[i] = m[i];
        nrm = (KIND == kPlane) ? plane_norm(m) : 1.0;
    }
    __device__ __forceinline__ double operator()(V3 q) const {
        if (KIND == kPlane) return plane_distance(w, nrm, q);
        if (KIND == kSphere) return sphere_distance(w, q);
        return cylinder_distance(w, q);
    }
};

}  // namespace ex
}  // namespace m3d
