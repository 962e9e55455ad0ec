/*
 * ref_driver.cpp -- TEST INFRASTRUCTURE ONLY.
 *
 * C entry points around the REFERENCE'S OWN sources, compiled from where they lie under
 * /root/reference (oracle/Makefile target `_ref` -> oracle/_ref/libm3d_ref.so):
 *     include/misc3d/common/ransac.h        RANSAC<>, Plane/Sphere/Cylinder estimators (unmodified)
 *     include/misc3d/utils.h                RandomSampler (unmodified)
 *     src/iterative_plane_segmentation.cpp  SegmentPlaneIterative (unmodified)
 *     src/correspondence_matching.cpp       ANNMatcher::Match / NearestSearch (unmodified)
 *     src/knn.cpp + vendored annoylib.h     KNearestSearch: the real Annoy forest of the ANNOY branch
 *     src/logging.cpp                       Logger (unmodified)
 * Eigen and Open3D do not exist in this image; oracle/shim/ provides stand-ins for the few types
 * those sources use (see shim/m3d_eigen_shim.h for what that does and does not prove).
 * Two things are injected from outside, without touching the sources:
 *   - the seed: utils.h:74-77 seeds mt19937 from std::random_device; shim/m3d_seed_hook.h re-points
 *     the token `random_device` at a device that returns seed, seed+1, ... (one per sampler
 *     construction, i.e. per FitModel = per segmentation round);
 *   - the iteration count: only observable through the LogInfo line of ransac.h:616-619, captured
 *     with Logger::SetPrintFunction.
 * Built twice: without -fopenmp (pragmas ignored: the sequential, seed-deterministic semantics the
 * parity tests need) and with -fopenmp (the reference's real OpenMP loop, used as the CPU
 * baseline `kind: "reference"`).
 */
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <string>

/* shim/m3d_seed_hook.h is force-included in front of every TU (oracle/Makefile) */
#include <misc3d/common/ransac.h>
#include <misc3d/registration/correspondence_matching.h>
#include <misc3d/segmentation/iterative_plane_segmentation.h>

extern "C" {
unsigned int m3dref_seed = 0;
}

namespace {
std::string g_last_info;
void capture(const std::string &s) { g_last_info = s; }

void fill_cloud(open3d::geometry::PointCloud &pc, const double *xyz, const double *nrm, size_t n) {
    pc.points_.resize(n);
    for (size_t i = 0; i < n; ++i) pc.points_[i] = Eigen::Vector3d(xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]);
    if (nrm) {
        pc.normals_.resize(n);
        for (size_t i = 0; i < n; ++i) pc.normals_[i] = Eigen::Vector3d(nrm[3 * i], nrm[3 * i + 1], nrm[3 * i + 2]);
    }
}

/* "Find best model with {}% inliers and run {} iterations" */
void parse_info(double *fitness_pct, unsigned long long *iterations) {
    *fitness_pct = 0;
    *iterations = 0;
    const char *s = g_last_info.c_str();
    const char *p = strstr(s, "model with ");
    if (p) *fitness_pct = atof(p + 11);
    p = strstr(s, "and run ");
    if (p) *iterations = strtoull(p + 8, nullptr, 10);
}

template <class R, class M>
int run_fit(const double *xyz, const double *nrm, size_t n, double thr, size_t max_it, double prob,
            double *model_out, size_t *inl_out, size_t *n_inl, double *fitness_pct, unsigned long long *iterations) {
    open3d::geometry::PointCloud pc;
    fill_cloud(pc, xyz, nrm, n);
    R fit;
    fit.SetMaxIteration(max_it);
    fit.SetProbability(prob);
    fit.SetPointCloud(pc);
    M model;
    std::vector<size_t> inliers;
    const bool ret = fit.FitModel(thr, model, inliers);
    for (size_t i = 0; i < (size_t)model.parameters_.size() && i < 8; ++i) model_out[i] = model.parameters_(i);
    *n_inl = inliers.size();
    if (inl_out) memcpy(inl_out, inliers.data(), sizeof(size_t) * inliers.size());
    parse_info(fitness_pct, iterations);
    return ret ? 1 : 0;
}
}  // namespace

extern "C" {

/* 1/0 = FitModel's return value, -1 = the reference threw (message in m3dref_last_error) */
static std::string g_err;
const char *m3dref_last_error() { return g_err.c_str(); }
void m3dref_set_threads(int n) { /* torchrun exports OMP_NUM_THREADS=1: the timing legs ask for the cores explicitly */
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#endif
}
int m3dref_openmp() {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 0;
#endif
}

int m3dref_ransac_fit(int kind, const double *xyz, const double *nrm, size_t n, double thr, size_t max_it,
                      double prob, uint32_t seed, double *model_out, size_t *inl_out, size_t *n_inl,
                      double *fitness_pct, unsigned long long *iterations) {
    using namespace misc3d::common;
    m3dref_seed = seed;
    misc3d::Logger::GetInstance().SetPrintFunction(capture);
    g_last_info.clear();
    for (int i = 0; i < 8; ++i) model_out[i] = 0;
    *n_inl = 0;
    try {
        switch (kind) {
            case 0:
                return run_fit<RANSACPlane, Plane>(xyz, nullptr, n, thr, max_it, prob, model_out, inl_out, n_inl,
                                                   fitness_pct, iterations);
            case 1:
                return run_fit<RANSACShpere, Sphere>(xyz, nullptr, n, thr, max_it, prob, model_out, inl_out, n_inl,
                                                     fitness_pct, iterations);
            default:
                return run_fit<RANSACCylinder, Cylinder>(xyz, nrm, n, thr, max_it, prob, model_out, inl_out, n_inl,
                                                         fitness_pct, iterations);
        }
    } catch (const std::exception &e) {
        g_err = e.what();
        return -1;
    }
}

/* the estimators alone (MinimalFit on k points given in the order SelectByIndex would emit them,
 * and CalcPointToModelDistance) */
int m3dref_minimal_fit(int kind, const double *pts, const double *nrm, int k, double *model_out) {
    using namespace misc3d::common;
    open3d::geometry::PointCloud pc;
    fill_cloud(pc, pts, nrm, (size_t)k);
    bool ret = false;
    try {
        if (kind == 0) {
            Plane m;
            ret = PlaneEstimator().MinimalFit(pc, m);
            for (int i = 0; i < 4; ++i) model_out[i] = m.parameters_(i);
        } else if (kind == 1) {
            Sphere m;
            ret = SphereEstimator().MinimalFit(pc, m);
            for (int i = 0; i < 4; ++i) model_out[i] = m.parameters_(i);
        } else {
            Cylinder m;
            ret = CylinderEstimator().MinimalFit(pc, m);
            for (int i = 0; i < 7; ++i) model_out[i] = m.parameters_(i);
        }
    } catch (const std::exception &e) {
        g_err = e.what();
        return -1;
    }
    return ret ? 1 : 0;
}
void m3dref_distances(int kind, const double *model, const double *xyz, size_t n, double *out) {
    using namespace misc3d::common;
    const int np = kind == 2 ? 7 : 4;
    Model m{Eigen::VectorXd(np)};
    for (int i = 0; i < np; ++i) m.parameters_(i) = model[i];
    PlaneEstimator pe;
    SphereEstimator se;
    CylinderEstimator ce;
    for (size_t i = 0; i < n; ++i) {
        const Eigen::Vector3d q(xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]);
        out[i] = kind == 0 ? pe.CalcPointToModelDistance(q, m)
                           : (kind == 1 ? se.CalcPointToModelDistance(q, m) : ce.CalcPointToModelDistance(q, m));
    }
}
int m3dref_general_fit(int kind, const double *xyz, size_t n, double *model_io) {
    using namespace misc3d::common;
    open3d::geometry::PointCloud pc;
    fill_cloud(pc, xyz, nullptr, n);
    const int np = kind == 2 ? 7 : 4;
    Model m{Eigen::VectorXd(np)};
    for (int i = 0; i < np; ++i) m.parameters_(i) = model_io[i];
    bool ret;
    if (kind == 0)
        ret = PlaneEstimator().GeneralFit(pc, m);
    else if (kind == 1)
        ret = SphereEstimator().GeneralFit(pc, m);
    else
        ret = CylinderEstimator().GeneralFit(pc, m);
    for (int i = 0; i < np; ++i) model_io[i] = m.parameters_(i);
    return ret ? 1 : 0;
}

/* the sampler alone: `rows` consecutive calls of RandomSampler<size_t>(n)(k) */
void m3dref_sample_table(uint32_t seed, size_t n, int k, size_t rows, size_t *out) {
    m3dref_seed = seed;
    misc3d::RandomSampler<size_t> sampler(n);
    for (size_t r = 0; r < rows; ++r) {
        const std::vector<size_t> s = sampler((size_t)k);
        for (int j = 0; j < k; ++j) out[r * k + j] = s[j];
    }
}

/* SegmentPlaneIterative; clusters are returned as point clouds by the reference, so labels are
 * rebuilt here by walking the surviving points in order (SelectByIndex is stable).
 * returns the number of planes, -1 on a throw, -2 if cap_planes is too small */
long m3dref_segment_plane_iterative(const double *xyz, size_t n, double thr, int max_it, double min_ratio,
                                    uint32_t seed, double *planes, size_t cap_planes, size_t *labels) {
    m3dref_seed = seed;
    misc3d::Logger::GetInstance().SetPrintFunction(capture);
    open3d::geometry::PointCloud pc;
    fill_cloud(pc, xyz, nullptr, n);
    std::vector<std::pair<Eigen::Vector4d, open3d::geometry::PointCloud>> res;
    try {
        res = misc3d::segmentation::SegmentPlaneIterative(pc, thr, max_it, min_ratio);
    } catch (const std::exception &e) {
        g_err = e.what();
        return -1;
    }
    if (res.size() > cap_planes) return -2;
    std::vector<size_t> remaining(n);
    for (size_t i = 0; i < n; ++i) remaining[i] = i, labels[i] = (size_t)-1;
    for (size_t c = 0; c < res.size(); ++c) {
        for (int j = 0; j < 4; ++j) planes[4 * c + j] = res[c].first(j);
        const auto &cl = res[c].second.points_;
        std::vector<size_t> next;
        next.reserve(remaining.size());
        size_t j = 0;
        for (size_t i : remaining) {
            if (j < cl.size() && cl[j](0) == xyz[3 * i] && cl[j](1) == xyz[3 * i + 1] && cl[j](2) == xyz[3 * i + 2]) {
                labels[i] = c;
                ++j;
            } else {
                next.push_back(i);
            }
        }
        if (j != cl.size()) return -3; /* cluster is not an in-order subsequence: cannot happen */
        remaining.swap(next);
    }
    return (long)res.size();
}

/* ANNMatcher(method, n_trees).Match(src, dst); descriptors dim x n column-major (= Eigen::MatrixXd memory).
 * method 0 = FLANN (exact; the Open3D kd-tree is stood in by an exact brute-force search),
 * method 1 = ANNOY (the reference's real, vendored Annoy forest: approximate, racy 4-thread build) */
long m3dref_match_correspondence(const double *src, size_t ns, const double *dst, size_t nd, int dim, int method,
                                 int n_trees, size_t *idx0, size_t *idx1) {
    Eigen::MatrixXd a, b;
    a.resize(dim, (long)ns);
    b.resize(dim, (long)nd);
    memcpy(a.data(), src, sizeof(double) * ns * dim);
    memcpy(b.data(), dst, sizeof(double) * nd * dim);
    try {
        misc3d::registration::ANNMatcher matcher(method == 0 ? misc3d::registration::MatchMethod::FLANN
                                                             : misc3d::registration::MatchMethod::ANNOY,
                                                 n_trees);
        const auto res = matcher.Match(a, b);
        memcpy(idx0, res.first.data(), sizeof(size_t) * res.first.size());
        memcpy(idx1, res.second.data(), sizeof(size_t) * res.second.size());
        return (long)res.first.size();
    } catch (const std::exception &e) {
        g_err = e.what();
        return -1;
    }
}

} /* extern "C" */
