/*
 * m3d_oracle_features.cpp -- CPU restatement (TEST INFRASTRUCTURE ONLY, like m3d_oracle.cpp) of the two Open3D steps
 * around the registration path: ComputeFPFHFeature and point-to-point RegistrationICP.
 *
 * PARITY UNPINNED: this is Open3D v0.15.1 code (pipelines/registration/Feature.cpp, Registration.cpp,
 * geometry/KDTreeFlann.cpp), a third-party dependency that is not under /root/reference and cannot be installed here;
 * the reference only CALLS it (examples/cpp/transform_estimation.cpp:20-33, 82-86).  The arithmetic below is restated
 * from the published algorithm; tools/pin_open3d.py records real Open3D outputs the moment one is importable.
 *
 *   KDTreeFlann::SearchHybrid(q, radius, max_nn)   nanoflann knnSearch(max_nn), then cut at d2 < radius^2: here brute
 *                                                  force, ascending (d2, index), d2 = ((dx^2 + dy^2) + dz^2)
 *   ComputeSPFHFeature / ComputeFPFHFeature        33 x n column-major (Feature::data_)
 *   RegistrationICP(src, dst, max_dist, init, PointToPoint(false), criteria)
 */
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <utility>
#include <vector>

extern "C" void orc_umeyama(const double *src, const double *dst, size_t n, int with_scaling, double *T);

namespace {

inline double dist2(const double *a, const double *b) {
    const double dx = a[0] - b[0], dy = a[1] - b[1], dz = a[2] - b[2];
    return (dx * dx + dy * dy) + dz * dz;
}

/* the max_nn nearest items with d2 < r2, ascending (d2, index) */
void hybrid(const double *xyz, size_t n, const double *q, double r2, int max_nn, std::vector<std::pair<double, uint32_t>> &out) {
    out.clear();
    for (size_t j = 0; j < n; ++j) {
        const double d2 = dist2(q, xyz + 3 * j);
        if (d2 < r2) out.emplace_back(d2, (uint32_t)j);
    }
    const size_t k = std::min<size_t>(out.size(), (size_t)max_nn);
    std::partial_sort(out.begin(), out.begin() + k, out.end());
    out.resize(k);
}

void pair_features(const double *p1, const double *n1, const double *p2, const double *n2, double f[4]) {
    double dp[3] = {p2[0] - p1[0], p2[1] - p1[1], p2[2] - p1[2]};
    f[3] = std::sqrt(dp[0] * dp[0] + dp[1] * dp[1] + dp[2] * dp[2]);
    f[0] = f[1] = f[2] = 0;
    if (f[3] == 0.0) {
        f[3] = 0;
        return;
    }
    double a[3] = {n1[0], n1[1], n1[2]}, b[3] = {n2[0], n2[1], n2[2]};
    const double angle1 = (a[0] * dp[0] + a[1] * dp[1] + a[2] * dp[2]) / f[3];
    const double angle2 = (b[0] * dp[0] + b[1] * dp[1] + b[2] * dp[2]) / f[3];
    if (std::acos(std::fabs(angle1)) > std::acos(std::fabs(angle2))) {
        for (int c = 0; c < 3; ++c) {
            a[c] = n2[c];
            b[c] = n1[c];
            dp[c] *= -1.0;
        }
        f[2] = -angle2;
    } else {
        f[2] = angle1;
    }
    double v[3] = {dp[1] * a[2] - dp[2] * a[1], dp[2] * a[0] - dp[0] * a[2], dp[0] * a[1] - dp[1] * a[0]};
    const double vn = std::sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
    if (vn == 0.0) {
        f[0] = f[1] = f[2] = f[3] = 0;
        return;
    }
    for (int c = 0; c < 3; ++c) v[c] /= vn;
    const double w[3] = {a[1] * v[2] - a[2] * v[1], a[2] * v[0] - a[0] * v[2], a[0] * v[1] - a[1] * v[0]};
    f[1] = v[0] * b[0] + v[1] * b[1] + v[2] * b[2];
    f[0] = std::atan2(w[0] * b[0] + w[1] * b[1] + w[2] * b[2], a[0] * b[0] + a[1] * b[1] + a[2] * b[2]);
}

}  // namespace

extern "C" {

/* neighbour lists of every point against the set: idx / d2 [n][max_nn], cnt [n] */
void orc_hybrid_search_all(const double *xyz, size_t n, double radius, int max_nn, uint32_t *idx, double *d2, uint32_t *cnt) {
#pragma omp parallel
    {
        std::vector<std::pair<double, uint32_t>> nb;
#pragma omp for schedule(dynamic, 64)
        for (long i = 0; i < (long)n; ++i) {
            hybrid(xyz, n, xyz + 3 * i, radius * radius, max_nn, nb);
            cnt[i] = (uint32_t)nb.size();
            for (size_t k = 0; k < nb.size(); ++k) {
                idx[(size_t)i * max_nn + k] = nb[k].second;
                d2[(size_t)i * max_nn + k] = nb[k].first;
            }
        }
    }
}

/* ComputeFPFHFeature(cloud, KDTreeSearchParamHybrid(radius, max_nn)); out: 33 x n column-major.  returns 0, or -4 when
 * the cloud has no normals (Open3D: LogError) */
int orc_fpfh(const double *xyz, const double *nrm, size_t n, double radius, int max_nn, double *out) {
    if (!nrm) return -4;
    std::vector<uint32_t> idx(n * (size_t)max_nn), cnt(n);
    std::vector<double> d2(n * (size_t)max_nn), spfh(33 * n, 0.0);
    orc_hybrid_search_all(xyz, n, radius, max_nn, idx.data(), d2.data(), cnt.data());
#pragma omp parallel for schedule(static)
    for (long i = 0; i < (long)n; ++i) {
        if (cnt[i] > 1) {
            const double incr = 100.0 / (double)(cnt[i] - 1);
            for (uint32_t k = 1; k < cnt[i]; ++k) {
                const uint32_t j = idx[(size_t)i * max_nn + k];
                double f[4];
                pair_features(xyz + 3 * i, nrm + 3 * i, xyz + 3 * (size_t)j, nrm + 3 * (size_t)j, f);
                int h = (int)std::floor(11 * (f[0] + M_PI) / (2.0 * M_PI));
                h = std::min(std::max(h, 0), 10);
                spfh[(size_t)i * 33 + h] += incr;
                h = (int)std::floor(11 * (f[1] + 1.0) * 0.5);
                h = std::min(std::max(h, 0), 10);
                spfh[(size_t)i * 33 + 11 + h] += incr;
                h = (int)std::floor(11 * (f[2] + 1.0) * 0.5);
                h = std::min(std::max(h, 0), 10);
                spfh[(size_t)i * 33 + 22 + h] += incr;
            }
        }
    }
#pragma omp parallel for schedule(static)
    for (long i = 0; i < (long)n; ++i) {
        double *f = out + (size_t)i * 33;
        for (int j = 0; j < 33; ++j) f[j] = 0;
        if (cnt[i] > 1) {
            double sum[3] = {0, 0, 0};
            for (uint32_t k = 1; k < cnt[i]; ++k) {
                const double dist = d2[(size_t)i * max_nn + k];
                if (dist == 0.0) continue;
                const double *s = spfh.data() + (size_t)idx[(size_t)i * max_nn + k] * 33;
                for (int j = 0; j < 33; ++j) {
                    const double val = s[j] / dist;
                    sum[j / 11] += val;
                    f[j] += val;
                }
            }
            for (int j = 0; j < 3; ++j)
                if (sum[j] != 0.0) sum[j] = 100.0 / sum[j];
            for (int j = 0; j < 33; ++j) {
                f[j] *= sum[j / 11];
                f[j] += spfh[(size_t)i * 33 + j];
            }
        }
    }
    return 0;
}

/* RegistrationICP, point to point.  T row-major.  returns 0 */
int orc_icp(const double *src, size_t ns, const double *dst, size_t nd, double max_dist, const double *T_init, int max_iter,
            double rel_fitness, double rel_rmse, double *T_out, double *fitness, double *rmse, int *iterations) {
    static const double I4[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};
    double T[16];
    std::memcpy(T, T_init ? T_init : I4, sizeof T);
    std::vector<double> pcd(src, src + 3 * ns);
    auto transform = [&](const double *M) {
        for (size_t i = 0; i < ns; ++i) {
            const double x = pcd[3 * i], y = pcd[3 * i + 1], z = pcd[3 * i + 2];
            const double w = ((M[12] * x + M[13] * y) + M[14] * z) + M[15];
            pcd[3 * i] = (((M[0] * x + M[1] * y) + M[2] * z) + M[3]) / w;
            pcd[3 * i + 1] = (((M[4] * x + M[5] * y) + M[6] * z) + M[7]) / w;
            pcd[3 * i + 2] = (((M[8] * x + M[9] * y) + M[10] * z) + M[11]) / w;
        }
    };
    if (std::memcmp(T, I4, sizeof T) != 0) transform(T);
    std::vector<long> corr(ns);
    double fit = 0, err = 0;
    auto evaluate = [&]() {
        const double r2 = max_dist * max_dist;
        double e2 = 0;
        size_t m = 0;
#pragma omp parallel for schedule(static) reduction(+ : e2, m)
        for (long i = 0; i < (long)ns; ++i) {
            double bd = INFINITY;
            long bj = -1;
            for (size_t j = 0; j < nd; ++j) {
                const double d2 = dist2(&pcd[3 * i], dst + 3 * j);
                if (d2 < r2 && d2 < bd) {
                    bd = d2;
                    bj = (long)j;
                }
            }
            corr[i] = bj;
            if (bj >= 0) {
                e2 += bd;
                ++m;
            }
        }
        fit = m ? (double)m / (double)ns : 0.0;
        err = m ? std::sqrt(e2 / (double)m) : 0.0;
        return m;
    };
    size_t m = evaluate();
    int it = 0;
    for (; it < max_iter; ++it) {
        if (m == 0) break;
        std::vector<double> a, b;
        a.reserve(3 * m);
        b.reserve(3 * m);
        for (size_t i = 0; i < ns; ++i)
            if (corr[i] >= 0) {
                a.insert(a.end(), &pcd[3 * i], &pcd[3 * i] + 3);
                b.insert(b.end(), dst + 3 * corr[i], dst + 3 * corr[i] + 3);
            }
        double U[16], Tn[16];
        orc_umeyama(a.data(), b.data(), m, 0, U);
        for (int r = 0; r < 4; ++r)
            for (int c = 0; c < 4; ++c) {
                double s = 0;
                for (int k = 0; k < 4; ++k) s += U[4 * r + k] * T[4 * k + c];
                Tn[4 * r + c] = s;
            }
        std::memcpy(T, Tn, sizeof T);
        transform(U);
        const double bf = fit, br = err;
        m = evaluate();
        if (std::fabs(bf - fit) < rel_fitness && std::fabs(br - err) < rel_rmse) {
            ++it;
            break;
        }
    }
    std::memcpy(T_out, T, sizeof T);
    if (fitness) *fitness = fit;
    if (rmse) *rmse = err;
    if (iterations) *iterations = it;
    return 0;
}

} /* extern "C" */
