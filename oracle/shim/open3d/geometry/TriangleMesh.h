#pragma once
/* stand-in for open3d/geometry/TriangleMesh.h (TEST INFRASTRUCTURE ONLY): src/knn.cpp names vertices_ */
#include <Eigen/Core>
#include <vector>

#include "Geometry.h"
namespace open3d {
namespace geometry {
class TriangleMesh : public Geometry {
public:
    TriangleMesh() : Geometry(GeometryType::TriangleMesh) {}
    std::vector<Eigen::Vector3d> vertices_;
};
}  // namespace geometry
}  // namespace open3d
