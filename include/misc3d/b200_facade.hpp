/*
 * b200_facade.hpp -- C++ facade over libm3d_b200.so that keeps the reference's class names and call
 * shapes for the RANSAC / segmentation / matching / registration path, so existing callers
 * (python/py_*.cpp, examples/cpp/{ransac_and_boundary,segment_plane_iterative,
 * transform_estimation}.cpp, src/pipeline.cpp:800-812) keep compiling against
 *   <misc3d/common/ransac.h>                                 (reference: include/misc3d/common/ransac.h)
 *   <misc3d/segmentation/iterative_plane_segmentation.h>     (.../iterative_plane_segmentation.h)
 *   <misc3d/registration/correspondence_matching.h>          (.../correspondence_matching.h)
 *   <misc3d/registration/transform_estimation.h>             (.../transform_estimation.h)
 *   <misc3d/logging.h>                                       (.../logging.h)
 *
 * Open3D and Eigen are not available in this build environment, so the geometry types are plain
 * standard-library stand-ins with the same memory layout (std::vector<std::array<double,3>> ==
 * std::vector<Eigen::Vector3d>).  Define M3D_WITH_OPEN3D before including to alias the real
 * open3d::geometry::PointCloud instead.
 *
 * Every compute call goes through the C-ABI of include/m3d_capi.h; C-ABI error codes are mapped
 * back to what the reference does (LogError -> throws std::runtime_error; FitModel -> bool).
 */
#pragma once
#include <algorithm>
#include <array>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <functional>
#include <memory>
#include <random>
#include <stdexcept>
#include <string>
#include <tuple>
#include <utility>
#include <vector>

#include "../m3d_capi.h"

namespace misc3d {

/* ------------------------------------------------------------------ logging (logging.h) */
enum class VerbosityLevel { Error = 0, Warning = 1, Info = 2, Debug = 3 };
namespace detail {
inline VerbosityLevel &verbosity() {
    static VerbosityLevel v = VerbosityLevel::Info; /* reference default, logging.cpp:56 */
    return v;
}
}  // namespace detail
inline void SetVerbosityLevel(VerbosityLevel level) { detail::verbosity() = level; }
inline VerbosityLevel GetVerbosityLevel() { return detail::verbosity(); }
[[noreturn]] inline void LogError(const std::string &msg) { /* logging.cpp:64-74: throws */
    throw std::runtime_error("[Misc3D Error] " + msg);
}
inline void LogWarning(const std::string &msg) {
    if (detail::verbosity() >= VerbosityLevel::Warning) std::fprintf(stderr, "[Misc3D Warning] %s\n", msg.c_str());
}
inline void LogInfo(const std::string &msg) {
    if (detail::verbosity() >= VerbosityLevel::Info) std::fprintf(stdout, "[Misc3D Info] %s\n", msg.c_str());
}

/* ------------------------------------------------------------------ geometry stand-ins */
using Vector3d = std::array<double, 3>;
using Vector4d = std::array<double, 4>;
using Matrix4d = std::array<double, 16>; /* row-major */

struct PointCloud { /* the part of open3d::geometry::PointCloud this path touches */
    std::vector<Vector3d> points_;
    std::vector<Vector3d> normals_;
    PointCloud() = default;
    explicit PointCloud(std::vector<Vector3d> pts) : points_(std::move(pts)) {}
    bool HasPoints() const { return !points_.empty(); }
    bool HasNormals() const { return !points_.empty() && normals_.size() == points_.size(); }
    void Clear() {
        points_.clear();
        normals_.clear();
    }
    /* ascending order, duplicates collapse (Open3D mask pass) */
    std::shared_ptr<PointCloud> SelectByIndex(const std::vector<size_t> &indices, bool invert = false) const {
        std::vector<bool> mask(points_.size(), invert);
        for (size_t i : indices) mask[i] = !invert;
        auto out = std::make_shared<PointCloud>();
        const bool nrm = HasNormals();
        for (size_t i = 0; i < points_.size(); ++i)
            if (mask[i]) {
                out->points_.push_back(points_[i]);
                if (nrm) out->normals_.push_back(normals_[i]);
            }
        return out;
    }
};
using PointCloudPtr = std::shared_ptr<PointCloud>;

/* dim x count column-major descriptors (open3d Feature::data_ / Eigen::MatrixXd) */
struct FeatureMatrix {
    int dim = 0;
    size_t count = 0;
    const double *data = nullptr; /* borrowed */
};

/* ------------------------------------------------------------------ the CUDA context */
namespace b200 {
inline m3d_ctx *DefaultContext() {
    struct Holder {
        m3d_ctx *ctx = nullptr;
        Holder() {
            const char *dev = std::getenv("M3D_DEVICE");
            if (m3d_ctx_create(dev ? std::atoi(dev) : 0, &ctx) != M3D_OK)
                throw std::runtime_error("misc3d (B200 build): no usable CUDA device; there is no CPU fallback");
        }
        ~Holder() { m3d_ctx_destroy(ctx); }
    };
    thread_local Holder h; /* one m3d_ctx per host thread */
    return h.ctx;
}
inline uint32_t RandomSeed() { /* the reference seeds every sampler from std::random_device */
    return std::random_device{}();
}
[[noreturn]] inline void Raise(m3d_ctx *ctx) { LogError(m3d_last_error(ctx)); }
}  // namespace b200

/* ------------------------------------------------------------------ common/ransac.h */
namespace common {

class Model { /* ransac.h:24-47 */
public:
    std::vector<double> parameters_; /* plane [a,b,c,d], sphere [x,y,z,r], cylinder [x,y,z,nx,ny,nz,r] */
protected:
    explicit Model(size_t n) : parameters_(n, 0.0) {}
};
class Plane : public Model {
public:
    Plane() : Model(4) {}
};
class Sphere : public Model {
public:
    Sphere() : Model(4) {}
};
class Cylinder : public Model {
public:
    Cylinder() : Model(7) {}
};

/* RANSAC<ModelEstimator, Model, Sampler> (ransac.h:455-664).  The estimator / sampler template
 * arguments of the reference collapse into the primitive id: sampling, minimal solves, scoring and
 * the refit all run inside m3d_ransac_fit. */
template <int KIND, class ModelT>
class RANSAC {
public:
    RANSAC() = default;
    void SetPointCloud(const PointCloud &pc) { pc_ = pc; }       /* deep copy, ransac.h:469-475 */
    void SetPointCloud(PointCloud &&pc) { pc_ = std::move(pc); } /* not in the reference: callers that own a temporary */
    void SetProbability(double probability) {                     /* ransac.h:482-487 */
        if (probability <= 0 || probability > 1) LogError("Probability must be > 0 or <= 1.0");
        probability_ = probability;
    }
    void SetMaxIteration(size_t num) { max_iteration_ = num; }    /* ransac.h:495 */
    /* not in the reference: fixes the sample stream (default: std::random_device, utils.h:74-77) */
    void SetSeed(uint32_t seed) {
        seed_ = seed;
        has_seed_ = true;
    }
    const m3d_ransac_stats &LastStats() const { return stats_; }

    /* ransac.h:506-516 */
    bool FitModel(double threshold, ModelT &model, std::vector<size_t> &inlier_indices) {
        m3d_ctx *ctx = b200::DefaultContext();
        const size_t n = pc_.points_.size();
        m3d_ransac_params p{};
        p.threshold = threshold;
        p.max_iteration = max_iteration_;
        p.probability = probability_;
        p.seed = has_seed_ ? seed_ : b200::RandomSeed();
        double out[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        inlier_indices.assign(n, 0);
        size_t n_inl = 0;
        const double *nrm = pc_.HasNormals() ? pc_.normals_[0].data() : nullptr;
        const int rc = m3d_ransac_fit(ctx, KIND, n ? pc_.points_[0].data() : nullptr, nrm, n, &p, out,
                                      inlier_indices.data(), &n_inl, &stats_);
        inlier_indices.resize(n_inl);
        if (rc < 0) b200::Raise(ctx); /* lack of points / no normals: the reference throws */
        char line[160];
        std::snprintf(line, sizeof line, "Find best model with %g%% inliers and run %llu iterations",
                      n ? 100.0 * (double)stats_.best_count / (double)n : 0.0,
                      (unsigned long long)stats_.iterations_run); /* ransac.h:616-619 */
        LogInfo(line);
        if (stats_.found) model.parameters_.assign(out, out + model.parameters_.size());
        return rc == 1;
    }

private:
    PointCloud pc_;
    double probability_ = 0.9999;  /* ransac.h:462 */
    size_t max_iteration_ = 1000;  /* ransac.h:461 */
    uint32_t seed_ = 0;
    bool has_seed_ = false;
    m3d_ransac_stats stats_{};
};
using RANSACPlane = RANSAC<M3D_PLANE, Plane>;
using RANSACShpere = RANSAC<M3D_SPHERE, Sphere>; /* (sic) ransac.h:667 */
using RANSACCylinder = RANSAC<M3D_CYLINDER, Cylinder>;

/* KNearestSearch (include/misc3d/common/knn.h:24-73, src/knn.cpp:36-139).  The reference wraps an approximate
 * Annoy forest (n_trees, racy 4-thread build); here the same class answers with the EXACT neighbours from a data
 * set resident in HBM (m3d_knn_*).  n_trees is kept for source compatibility and ignored.  Distances are what
 * Annoy reports: Euclidean, not squared (the reference's parameter name `distance2` notwithstanding). */
class KNearestSearch {
public:
    KNearestSearch() : n_trees_(4) {}
    explicit KNearestSearch(int n_trees) : n_trees_(n_trees) {}
    KNearestSearch(const FeatureMatrix &data, int n_trees = 4) : n_trees_(n_trees) { SetMatrixData(data); }
    KNearestSearch(const PointCloud &geometry, int n_trees = 4) : n_trees_(n_trees) { SetGeometry(geometry); }
    ~KNearestSearch() { m3d_knn_free(index_); }
    KNearestSearch(const KNearestSearch &) = delete;
    KNearestSearch &operator=(const KNearestSearch &) = delete;

    bool SetMatrixData(const FeatureMatrix &data) { return SetRawData(data.data, (size_t)data.dim, data.count); }
    bool SetGeometry(const PointCloud &geometry) { /* knn.cpp:57-77: the points as a 3 x n matrix */
        return SetRawData(geometry.points_.empty() ? nullptr : geometry.points_[0].data(), 3, geometry.points_.size());
    }
    bool SetFeature(const FeatureMatrix &feature) { return SetMatrixData(feature); }

    /* knn.cpp:103-113 */
    int SearchKNN(const std::vector<double> &query, int knn, std::vector<size_t> &indices,
                  std::vector<double> &distance2) const {
        if (dataset_size_ == 0 || query.size() != dimension_ || knn < 0) return -1;
        return Query(query, knn, 0.0, indices, distance2);
    }
    /* knn.cpp:115-139, including its off-by-one: of the neighbours within `radius` the LAST one is dropped
     * (`num = i - 1`).  With no neighbour inside the radius the reference resizes to SIZE_MAX (i.e. crashes);
     * here that case returns 0. */
    int SearchHybrid(const std::vector<double> &query, double radius, int knn, std::vector<size_t> &indices,
                     std::vector<double> &distance2) const {
        if (dataset_size_ == 0 || query.size() != dimension_ || knn < 0) return -1;
        Query(query, knn, 0.0, indices, distance2);
        size_t i = 0;
        for (; i < indices.size(); ++i)
            if (distance2[i] > radius) break;
        const size_t num = i == 0 ? 0 : i - 1;
        indices.resize(num);
        distance2.resize(num);
        return (int)num;
    }

private:
    bool SetRawData(const double *data, size_t dim, size_t count) { /* knn.cpp:36-50 */
        dimension_ = dim;
        dataset_size_ = count;
        m3d_knn_free(index_);
        index_ = nullptr;
        if (dimension_ == 0 || dataset_size_ == 0) return false;
        m3d_ctx *ctx = b200::DefaultContext();
        if (m3d_knn_create(ctx, data, (int)dim, count, &index_) != M3D_OK) b200::Raise(ctx);
        return true;
    }
    int Query(const std::vector<double> &query, int knn, double radius, std::vector<size_t> &indices,
              std::vector<double> &distance) const {
        m3d_ctx *ctx = b200::DefaultContext();
        indices.assign((size_t)knn, 0);
        distance.assign((size_t)knn, 0.0);
        int count = 0;
        if (m3d_knn_search(ctx, index_, query.data(), 1, knn, radius, indices.data(), distance.data(), &count) != M3D_OK)
            b200::Raise(ctx);
        indices.resize((size_t)count);
        distance.resize((size_t)count);
        return count;
    }
    int n_trees_;
    m3d_knn *index_ = nullptr;
    size_t dimension_ = 0, dataset_size_ = 0;
};

}  // namespace common

/* ------------------------------------------------------------------ segmentation */
namespace segmentation {
/* iterative_plane_segmentation.h:25-28 / .cpp:7-39 */
inline std::vector<std::pair<Vector4d, PointCloud>> SegmentPlaneIterative(const PointCloud &pcd,
                                                                          const double threshold,
                                                                          const int max_iteration = 100,
                                                                          const double min_ratio = 0.05,
                                                                          const uint32_t *seed = nullptr) {
    std::vector<std::pair<Vector4d, PointCloud>> result;
    const size_t n = pcd.points_.size();
    if (n < 3) {
        LogWarning("No enough points to segment plane."); /* :14-17 */
        return result;
    }
    m3d_ctx *ctx = b200::DefaultContext();
    /* the reference has no limit on the number of planes: the plane buffer grows (the call is deterministic in
     * `seed`, so a retry reproduces the rounds already done) */
    const uint32_t the_seed = seed ? *seed : b200::RandomSeed();
    size_t cap = 1024;
    std::vector<double> planes;
    std::vector<uint32_t> labels(n);
    size_t n_planes = 0;
    for (;;) {
        planes.assign(4 * cap, 0.0);
        const int rc = m3d_segment_plane_iterative_u32(ctx, pcd.points_[0].data(), n, threshold, max_iteration, min_ratio,
                                                       the_seed, planes.data(), cap, labels.data(), &n_planes, nullptr);
        if (rc == M3D_ERR_CAPACITY && cap < n / 3 + 1) {
            cap = std::min(n / 3 + 1, cap * 16);
            continue;
        }
        if (rc != M3D_OK) b200::Raise(ctx);
        break;
    }
    result.resize(n_planes);
    for (size_t k = 0; k < n_planes; ++k)
        for (int i = 0; i < 4; ++i) result[k].first[i] = planes[4 * k + i];
    for (size_t i = 0; i < n; ++i) /* clusters keep ascending original order (stable compaction) */
        if (labels[i] != UINT32_MAX) result[labels[i]].second.points_.push_back(pcd.points_[i]);
    return result;
}
}  // namespace segmentation

/* ------------------------------------------------------------------ registration */
namespace registration {

enum class MatchMethod { FLANN = 0, ANNOY = 1 }; /* correspondence_matching.h:14-17 */

/* Extension: descriptors that live on the device (m3d_features) -- FPFH computed there and matched there, without the
 * 2 x 52.8 MB host round trip of the reference's callers (examples/cpp/transform_estimation.cpp:20-33). */
class DeviceFeature {
public:
    DeviceFeature() = default;
    DeviceFeature(const DeviceFeature &) = delete;
    DeviceFeature &operator=(const DeviceFeature &) = delete;
    DeviceFeature(DeviceFeature &&o) noexcept : h_(o.h_) { o.h_ = nullptr; }
    DeviceFeature &operator=(DeviceFeature &&o) noexcept {
        if (this != &o) {
            m3d_features_free(h_);
            h_ = o.h_;
            o.h_ = nullptr;
        }
        return *this;
    }
    ~DeviceFeature() { m3d_features_free(h_); }
    /* open3d::pipelines::registration::ComputeFPFHFeature(cloud, KDTreeSearchParamHybrid(radius, max_nn)) */
    static DeviceFeature FPFH(const PointCloud &cloud, double radius, int max_nn = 100) {
        if (!cloud.HasNormals() && cloud.HasPoints()) LogError("Failed because input point cloud has no normal.");
        m3d_ctx *ctx = b200::DefaultContext();
        DeviceFeature f;
        const size_t n = cloud.points_.size();
        if (m3d_fpfh_create(ctx, n ? cloud.points_[0].data() : nullptr, n ? cloud.normals_[0].data() : nullptr, n, radius,
                            max_nn, &f.h_, nullptr) != M3D_OK)
            b200::Raise(ctx);
        return f;
    }
    static DeviceFeature Upload(const FeatureMatrix &m) {
        m3d_ctx *ctx = b200::DefaultContext();
        DeviceFeature f;
        if (m3d_features_upload(ctx, m.data, m.dim, m.count, &f.h_) != M3D_OK) b200::Raise(ctx);
        return f;
    }
    int Dimension() const { return m3d_features_dim(h_); }
    size_t Num() const { return m3d_features_count(h_); }
    std::vector<double> Download() const { /* dim x n, column-major (usable as FeatureMatrix{dim, n, data}) */
        std::vector<double> out((size_t)Dimension() * Num());
        if (h_ && !out.empty() && m3d_features_download(h_, out.data()) != M3D_OK) b200::Raise(b200::DefaultContext());
        return out;
    }
    const m3d_features *handle() const { return h_; }

private:
    m3d_features *h_ = nullptr;
};

class ANNMatcher { /* correspondence_matching.h:67-91 */
public:
    ANNMatcher() : method_(MatchMethod::FLANN), n_trees_(4) {}
    explicit ANNMatcher(const MatchMethod &method) : method_(method), n_trees_(4) {}
    ANNMatcher(const MatchMethod &method, int n_trees) : method_(method), n_trees_(n_trees) {}

    std::pair<std::vector<size_t>, std::vector<size_t>> Match(const FeatureMatrix &src,
                                                             const FeatureMatrix &dst) const {
        if (src.dim != dst.dim) LogError("Descriptor dimensions differ");
        m3d_ctx *ctx = b200::DefaultContext();
        std::pair<std::vector<size_t>, std::vector<size_t>> out;
        out.first.resize(src.count);
        out.second.resize(src.count);
        size_t n_out = 0;
        const int rc = m3d_match_correspondence(ctx, src.data, src.count, dst.data, dst.count, src.dim,
                                                (int)method_, n_trees_, out.first.data(), out.second.data(), &n_out,
                                                nullptr);
        if (rc != M3D_OK) b200::Raise(ctx);
        out.first.resize(n_out);
        out.second.resize(n_out);
        return out;
    }

    /* extension: both descriptor sets already on the device */
    std::pair<std::vector<size_t>, std::vector<size_t>> Match(const DeviceFeature &src, const DeviceFeature &dst) const {
        m3d_ctx *ctx = b200::DefaultContext();
        std::pair<std::vector<size_t>, std::vector<size_t>> out;
        out.first.resize(src.Num());
        out.second.resize(src.Num());
        size_t n_out = 0;
        if (!src.handle() || !dst.handle()) LogError("Empty device feature");
        if (m3d_match_features(ctx, src.handle(), dst.handle(), out.first.data(), out.second.data(), &n_out, nullptr) != M3D_OK)
            b200::Raise(ctx);
        out.first.resize(n_out);
        out.second.resize(n_out);
        return out;
    }

private:
    MatchMethod method_;
    int n_trees_;
};

class RANSACSolver { /* transform_estimation.h:114-146 */
public:
    explicit RANSACSolver(double threshold, int max_iter = 100000, double edge_length_threshold = 0.9)
        : threshold_(threshold), max_iter_(max_iter), edge_length_threshold_(edge_length_threshold) {}
    /* the reference's member is self-initialised (transform_estimation.h:126): the argument is
     * honoured here -- documented deviation */
    void SetSeed(uint32_t seed) {
        seed_ = seed;
        has_seed_ = true;
    }
    Matrix4d Solve(const PointCloud &src, const PointCloud &dst,
                   const std::pair<std::vector<size_t>, std::vector<size_t>> &corres) const {
        if (corres.first.size() != corres.second.size()) LogError("Correspondence lists differ in length");
        m3d_ctx *ctx = b200::DefaultContext();
        Matrix4d T{};
        const int rc = m3d_ransac_registration(
            ctx, src.points_.empty() ? nullptr : src.points_[0].data(), src.points_.size(),
            dst.points_.empty() ? nullptr : dst.points_[0].data(), dst.points_.size(), corres.first.data(),
            corres.second.data(), corres.first.size(), threshold_, max_iter_, edge_length_threshold_, 0.999,
            has_seed_ ? seed_ : b200::RandomSeed(), T.data(), nullptr);
        if (rc < 0) b200::Raise(ctx); /* < 3 points: transform_estimation.cpp:130-133 throws */
        return T;
    }

private:
    double threshold_;
    int max_iter_;
    double edge_length_threshold_;
    uint32_t seed_ = 0;
    bool has_seed_ = false;
};

/* Extension (not in the reference): LeastSquareSolver applied behind RANSACSolver -- Umeyama over the correspondences
 * that are inliers of T (|T s - d| < threshold), on the device.  Returns T unchanged when fewer than 3 are. */
inline Matrix4d RefineOnInlierCorrespondences(const PointCloud &src, const PointCloud &dst,
                                              const std::pair<std::vector<size_t>, std::vector<size_t>> &corres,
                                              const Matrix4d &T, double threshold, bool scaling = false,
                                              size_t *n_inliers = nullptr) {
    if (corres.first.size() != corres.second.size()) LogError("Correspondence lists differ in length");
    m3d_ctx *ctx = b200::DefaultContext();
    Matrix4d out{};
    const int rc = m3d_registration_refit(
        ctx, src.points_.empty() ? nullptr : src.points_[0].data(), src.points_.size(),
        dst.points_.empty() ? nullptr : dst.points_[0].data(), dst.points_.size(), corres.first.data(),
        corres.second.data(), corres.first.size(), T.data(), threshold, scaling ? 1 : 0, out.data(), n_inliers);
    if (rc != M3D_OK) b200::Raise(ctx);
    return out;
}

/* The Open3D calls that surround the path in the reference's callers (examples/cpp/transform_estimation.cpp:20-33,
 * 82-86): not Misc3D API, offered so that descriptors and the refinement need not leave the GPU build. */
/* open3d::pipelines::registration::ComputeFPFHFeature(cloud, KDTreeSearchParamHybrid(radius, max_nn)):
 * 33 x n column-major descriptors (usable as FeatureMatrix{33, n, data}) */
inline std::vector<double> ComputeFPFHFeature(const PointCloud &cloud, double radius, int max_nn = 100) {
    if (!cloud.HasNormals() && cloud.HasPoints()) LogError("Failed because input point cloud has no normal.");
    m3d_ctx *ctx = b200::DefaultContext();
    std::vector<double> out(33 * cloud.points_.size());
    const size_t n = cloud.points_.size();
    if (n == 0) return out;
    if (m3d_compute_fpfh(ctx, cloud.points_[0].data(), cloud.normals_[0].data(), n, radius, max_nn, out.data(), nullptr) != M3D_OK)
        b200::Raise(ctx);
    return out;
}
struct ICPResult {
    Matrix4d transformation_{};
    double fitness_ = 0, inlier_rmse_ = 0;
    int iterations_ = 0;
};
/* open3d::pipelines::registration::RegistrationICP, TransformationEstimationPointToPoint(false) */
inline ICPResult RegistrationICP(const PointCloud &source, const PointCloud &target, double max_correspondence_distance,
                                 const Matrix4d &init = Matrix4d{1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1},
                                 int max_iteration = 30, double relative_fitness = 1e-6, double relative_rmse = 1e-6) {
    m3d_ctx *ctx = b200::DefaultContext();
    ICPResult r;
    const int rc = m3d_icp_point_to_point(ctx, source.points_.empty() ? nullptr : source.points_[0].data(), source.points_.size(),
                                          target.points_.empty() ? nullptr : target.points_[0].data(), target.points_.size(),
                                          max_correspondence_distance, init.data(), max_iteration, relative_fitness,
                                          relative_rmse, r.transformation_.data(), &r.fitness_, &r.inlier_rmse_, &r.iterations_);
    if (rc != M3D_OK) b200::Raise(ctx);
    return r;
}

class LeastSquareSolver { /* transform_estimation.cpp:49-66 */
public:
    explicit LeastSquareSolver(bool scaling = false) : scaling_(scaling) {}
    /* src, dst: n corresponding points each */
    Matrix4d Solve(const std::vector<Vector3d> &src, const std::vector<Vector3d> &dst) const {
        /* CheckValid, transform_estimation.cpp:27-37: both conditions throw */
        if (src.size() < 3 || dst.size() < 3) LogError("The number of points pair is less than 3.");
        if (src.size() != dst.size()) LogError("The number of points pair is not equal.");
        m3d_ctx *ctx = b200::DefaultContext();
        Matrix4d T{};
        const int rc = m3d_least_squares_transform(ctx, src.empty() ? nullptr : src[0].data(),
                                                   dst.empty() ? nullptr : dst[0].data(), src.size(), scaling_ ? 1 : 0,
                                                   T.data());
        if (rc != M3D_OK) b200::Raise(ctx);
        return T;
    }

private:
    bool scaling_;
};

}  // namespace registration
}  // namespace misc3d
