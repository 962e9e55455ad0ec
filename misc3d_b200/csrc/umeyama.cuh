/*
 * umeyama.cuh -- Eigen::umeyama (3-D, two-sided Jacobi SVD in the operation order of the test suite's CPU checker) and Open3D's point
 * transform as device functions; shared by the registration RANSAC (registration.cu) and the ICP refinement
 * (features.cu).  Reference: src/transform_estimation.cpp:49-66, 124-164 (-> Open3D / Eigen).
 */
#pragma once
#include "exact_math.cuh"

namespace m3d {

namespace rg {
using ex::add;
using ex::div;
using ex::mul;
using ex::sqrt_;
using ex::sub;

struct Rot {
    double c, s;
};
/* Eigen JacobiRotation::makeJacobi(x, y, z) for the symmetric 2x2 [x y; y z] */
__device__ __forceinline__ void make_jacobi(double x, double y, double z, Rot &r) {
    const double deno = mul(2.0, fabs(y));
    if (deno < DBL_MIN) {
        r.c = 1;
        r.s = 0;
        return;
    }
    const double tau = div(sub(x, z), deno);
    const double w = sqrt_(add(mul(tau, tau), 1.0));
    const double t = tau > 0 ? div(1.0, add(tau, w)) : div(1.0, sub(tau, w));
    const double sign_t = t > 0 ? 1.0 : -1.0;
    const double n = div(1.0, sqrt_(add(mul(t, t), 1.0)));
    r.s = mul(mul(mul(-sign_t, div(y, fabs(y))), fabs(t)), n);
    r.c = n;
}
template <int P, int Q>
__device__ __forceinline__ void rot_left(double (&w)[3][3], Rot j) { /* rows P,Q */
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        const double x = w[P][i], y = w[Q][i];
        w[P][i] = add(mul(j.c, x), mul(j.s, y));
        w[Q][i] = add(mul(-j.s, x), mul(j.c, y));
    }
}
template <int P, int Q>
__device__ __forceinline__ void rot_right(double (&w)[3][3], Rot j) { /* columns P,Q */
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        const double x = w[i][P], y = w[i][Q];
        w[i][P] = sub(mul(j.c, x), mul(j.s, y));
        w[i][Q] = add(mul(j.s, x), mul(j.c, y));
    }
}
template <int P, int Q>
__device__ __forceinline__ bool sweep_pq(double (&w)[3][3], double (&U)[3][3], double (&V)[3][3],
                                         double &max_diag) {
    const double precision = 2.0 * DBL_EPSILON, consider_as_zero = DBL_MIN;
    const double threshold = fmax(consider_as_zero, mul(precision, max_diag));
    if (!(fabs(w[P][Q]) > threshold || fabs(w[Q][P]) > threshold)) return false;
    /* real_2x2_jacobi_svd */
    const double m00 = w[P][P], m01 = w[P][Q], m10 = w[Q][P], m11 = w[Q][Q];
    Rot rot1;
    const double t = add(m00, m11), d = sub(m10, m01);
    if (fabs(d) < DBL_MIN) {
        rot1.s = 0;
        rot1.c = 1;
    } else {
        const double u = div(t, d);
        const double tmp = sqrt_(add(1.0, mul(u, u)));
        rot1.s = div(1.0, tmp);
        rot1.c = div(u, tmp);
    }
    const double n00 = add(mul(rot1.c, m00), mul(rot1.s, m10));
    const double n01 = add(mul(rot1.c, m01), mul(rot1.s, m11));
    const double n11 = add(mul(-rot1.s, m01), mul(rot1.c, m11));
    Rot jr;
    make_jacobi(n00, n01, n11, jr);
    Rot jl; /* j_left = rot1 * j_right^T */
    jl.c = add(mul(rot1.c, jr.c), mul(rot1.s, jr.s));
    jl.s = add(mul(-rot1.c, jr.s), mul(rot1.s, jr.c));
    rot_left<P, Q>(w, jl);
    rot_right<P, Q>(U, Rot{jl.c, -jl.s});
    rot_right<P, Q>(w, jr);
    rot_right<P, Q>(V, jr);
    max_diag = fmax(max_diag, fmax(fabs(w[P][P]), fabs(w[Q][Q])));
    return true;
}
/* Eigen JacobiSVD<Matrix3d>(FullU | FullV): A = U diag(sv) V^T */
__device__ inline void jacobi_svd3(const double (&a)[3][3], double (&U)[3][3], double (&sv)[3], double (&V)[3][3]) {
    double scale = 0;
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) scale = fmax(scale, fabs(a[i][j]));
    if (scale == 0.0) scale = 1.0;
    double w[3][3];
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            w[i][j] = div(a[i][j], scale);
            U[i][j] = V[i][j] = (i == j) ? 1.0 : 0.0;
        }
    double max_diag = fmax(fabs(w[0][0]), fmax(fabs(w[1][1]), fabs(w[2][2])));
    bool finished = false;
    int guard = 0;
    while (!finished && guard++ < 100) {
        finished = true;
        if (sweep_pq<1, 0>(w, U, V, max_diag)) finished = false;
        if (sweep_pq<2, 0>(w, U, V, max_diag)) finished = false;
        if (sweep_pq<2, 1>(w, U, V, max_diag)) finished = false;
    }
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        const double aa = fabs(w[i][i]);
        sv[i] = aa;
        if (aa != 0.0 && w[i][i] < 0) {
#pragma unroll
            for (int r = 0; r < 3; ++r) U[r][i] = -U[r][i];
        }
    }
#pragma unroll
    for (int i = 0; i < 3; ++i) sv[i] = mul(sv[i], scale);
    /* sort singular values in decreasing order (selection, with column swaps) */
    bool stop = false;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        if (stop) continue;
        int pos = i;
#pragma unroll
        for (int j = i + 1; j < 3; ++j)
            if (sv[j] > sv[pos]) pos = j;
        if (sv[pos] == 0.0) {
            stop = true;
            continue;
        }
        if (pos != i) {
            const double ts = sv[i];
            sv[i] = sv[pos];
            sv[pos] = ts;
#pragma unroll
            for (int r = 0; r < 3; ++r) {
                const double tu = U[r][i], tv = V[r][i];
                U[r][i] = U[r][pos];
                U[r][pos] = tu;
                V[r][i] = V[r][pos];
                V[r][pos] = tv;
            }
        }
    }
}
__device__ __forceinline__ double det3(const double (&m)[3][3]) {
    return add(sub(mul(m[0][0], sub(mul(m[1][1], m[2][2]), mul(m[1][2], m[2][1]))),
                   mul(m[0][1], sub(mul(m[1][0], m[2][2]), mul(m[1][2], m[2][0])))),
               mul(m[0][2], sub(mul(m[1][0], m[2][1]), mul(m[1][1], m[2][0]))));
}
/* rotation / translation from the means and the covariance (Eigen::umeyama tail) */
__device__ inline void umeyama_finish(const double (&sigma)[3][3], const double (&sm)[3], const double (&dm)[3],
                                      bool with_scaling, double src_var, double *T /*16, row-major*/) {
    double U[3][3], V[3][3], sv[3];
    jacobi_svd3(sigma, U, sv, V);
    double S[3] = {1, 1, 1};
    if (mul(det3(U), det3(V)) < 0) S[2] = -1;
    double R[3][3];
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int c = 0; c < 3; ++c)
            R[r][c] = add(add(mul(mul(U[r][0], S[0]), V[c][0]), mul(mul(U[r][1], S[1]), V[c][1])),
                          mul(mul(U[r][2], S[2]), V[c][2]));
    double cs = 1.0;
    if (with_scaling) cs = mul(div(1.0, src_var), add(add(mul(sv[0], S[0]), mul(sv[1], S[1])), mul(sv[2], S[2])));
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        const double rs = add(add(mul(R[r][0], sm[0]), mul(R[r][1], sm[1])), mul(R[r][2], sm[2]));
        T[4 * r + 3] = with_scaling ? sub(dm[r], mul(cs, rs)) : sub(dm[r], rs);
#pragma unroll
        for (int c = 0; c < 3; ++c) T[4 * r + c] = with_scaling ? mul(R[r][c], cs) : R[r][c];
    }
    T[12] = T[13] = T[14] = 0;
    T[15] = 1;
}
/* Open3D PointCloud::Transform on one point: T*[p;1] */
__device__ __forceinline__ ex::V3 xform(const double *T, ex::V3 p) {
    ex::V3 o;
    o.x = add(add(add(mul(T[0], p.x), mul(T[1], p.y)), mul(T[2], p.z)), mul(T[3], 1.0));
    o.y = add(add(add(mul(T[4], p.x), mul(T[5], p.y)), mul(T[6], p.z)), mul(T[7], 1.0));
    o.z = add(add(add(mul(T[8], p.x), mul(T[9], p.y)), mul(T[10], p.z)), mul(T[11], 1.0));
    return o;
}
__device__ __forceinline__ double dis2(const double *T, ex::V3 p, ex::V3 q) {
    const ex::V3 df = ex::sub3(xform(T, p), q);
    return ex::dot3(df, df);
}
}  // namespace rg

}  // namespace m3d
