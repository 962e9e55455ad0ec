#pragma once
/*
 * stand-in for open3d/geometry/PointCloud.h (TEST INFRASTRUCTURE ONLY): the members the reference's
 * RANSAC / segmentation sources touch.  SelectByIndex follows Open3D v0.15.1's published behaviour
 * (SURVEY.md Appendix B): a bool mask over all N points, then ONE ascending pass -- so the output
 * is in ascending index order, duplicates collapse, and the cost is O(N) per call.
 */
#include <Eigen/Core>
#include <memory>
#include <vector>

#include "Geometry.h"

namespace open3d {
namespace geometry {
class PointCloud : public Geometry {
public:
    PointCloud() : Geometry(GeometryType::PointCloud) {}
    PointCloud(const std::vector<Eigen::Vector3d> &points) : Geometry(GeometryType::PointCloud), points_(points) {}
    bool HasPoints() const { return points_.size() > 0; }
    bool HasNormals() const { return points_.size() > 0 && normals_.size() == points_.size(); }
    bool HasColors() const { return points_.size() > 0 && colors_.size() == points_.size(); }
    PointCloud &Clear() {
        points_.clear();
        normals_.clear();
        colors_.clear();
        return *this;
    }
    std::shared_ptr<PointCloud> SelectByIndex(const std::vector<size_t> &indices, bool invert = false) const {
        auto output = std::make_shared<PointCloud>();
        const bool has_normals = HasNormals(), has_colors = HasColors();
        std::vector<bool> mask(points_.size(), invert);
        for (size_t i : indices) mask[i] = !invert;
        for (size_t i = 0; i < points_.size(); ++i) {
            if (mask[i]) {
                output->points_.push_back(points_[i]);
                if (has_normals) output->normals_.push_back(normals_[i]);
                if (has_colors) output->colors_.push_back(colors_[i]);
            }
        }
        return output;
    }
    std::vector<Eigen::Vector3d> points_;
    std::vector<Eigen::Vector3d> normals_;
    std::vector<Eigen::Vector3d> colors_;
};
}  // namespace geometry
}  // namespace open3d
