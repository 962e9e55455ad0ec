#pragma once
/* stand-in for open3d/pipelines/registration/Feature.h (TEST INFRASTRUCTURE ONLY) */
#include <Eigen/Core>
namespace open3d {
namespace pipelines {
namespace registration {
class Feature {
public:
    void Resize(int dim, int n) { data_.resize(dim, n); }
    size_t Dimension() const { return data_.rows(); }
    size_t Num() const { return data_.cols(); }
    Eigen::MatrixXd data_;
};
}  // namespace registration
}  // namespace pipelines
}  // namespace open3d
