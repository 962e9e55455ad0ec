#pragma once
/* stand-in for open3d/geometry/KDTreeSearchParam.h (TEST INFRASTRUCTURE ONLY) */
namespace open3d {
namespace geometry {
class KDTreeSearchParam {
public:
    enum class SearchType { Knn = 0, Radius = 1, Hybrid = 2 };
    virtual ~KDTreeSearchParam() {}
    SearchType GetSearchType() const { return search_type_; }

protected:
    explicit KDTreeSearchParam(SearchType t) : search_type_(t) {}

private:
    SearchType search_type_;
};
class KDTreeSearchParamKNN : public KDTreeSearchParam {
public:
    explicit KDTreeSearchParamKNN(int knn = 30) : KDTreeSearchParam(SearchType::Knn), knn_(knn) {}
    int knn_;
};
class KDTreeSearchParamHybrid : public KDTreeSearchParam {
public:
    KDTreeSearchParamHybrid(double radius, int max_nn)
        : KDTreeSearchParam(SearchType::Hybrid), radius_(radius), max_nn_(max_nn) {}
    double radius_;
    int max_nn_;
};
}  // namespace geometry
}  // namespace open3d
