#pragma once
/* stand-in for open3d/geometry/Geometry.h (TEST INFRASTRUCTURE ONLY) */
namespace open3d {
namespace geometry {
class Geometry {
public:
    enum class GeometryType { Unspecified = 0, PointCloud = 1, TriangleMesh = 6 };
    explicit Geometry(GeometryType t = GeometryType::Unspecified) : type_(t) {}
    virtual ~Geometry() {}
    GeometryType GetGeometryType() const { return type_; }

private:
    GeometryType type_;
};
class TriangleMesh; /* defined in TriangleMesh.h; misc3d/utils.h names it in a typedef */
}  // namespace geometry
}  // namespace open3d
