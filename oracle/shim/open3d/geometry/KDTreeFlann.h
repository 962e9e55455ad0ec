#pragma once
/*
 * stand-in for open3d/geometry/KDTreeFlann.h (TEST INFRASTRUCTURE ONLY).  Open3D's class is an
 * EXACT L2 nearest-neighbour search (nanoflann kd-tree); this stand-in returns the same exact
 * neighbours by brute force, with nanoflann's distance accumulation order (groups of four, then the
 * remainder) and the lowest index on exact ties (nanoflann's own tie order is unspecified).
 */
#include <Eigen/Core>
#include <algorithm>
#include <limits>
#include <vector>
namespace open3d {
namespace geometry {
class KDTreeFlann {
public:
    KDTreeFlann() {}
    explicit KDTreeFlann(const Eigen::MatrixXd &data) : data_(data) {}
    int SearchKNN(const Eigen::VectorXd &query, int knn, std::vector<int> &indices,
                  std::vector<double> &distance2) const {
        const long dim = data_.rows(), n = data_.cols();
        if (n == 0 || (long)query.size() != dim || knn < 0) return -1;
        std::vector<std::pair<double, int>> all((size_t)n);
        for (long j = 0; j < n; ++j) all[(size_t)j] = {dist(query.data(), data_.data() + j * dim, (int)dim), (int)j};
        const size_t k = std::min<size_t>((size_t)knn, (size_t)n);
        std::partial_sort(all.begin(), all.begin() + k, all.end());
        indices.resize(k);
        distance2.resize(k);
        for (size_t i = 0; i < k; ++i) indices[i] = all[i].second, distance2[i] = all[i].first;
        return (int)k;
    }

private:
    static double dist(const double *a, const double *b, int dim) {
        double result = 0;
        int d = 0;
        for (; d + 3 < dim; d += 4) {
            const double d0 = a[d] - b[d], d1 = a[d + 1] - b[d + 1], d2 = a[d + 2] - b[d + 2], d3 = a[d + 3] - b[d + 3];
            result += d0 * d0 + d1 * d1 + d2 * d2 + d3 * d3;
        }
        for (; d < dim; ++d) {
            const double d0 = a[d] - b[d];
            result += d0 * d0;
        }
        return result;
    }
    Eigen::MatrixXd data_;
};
}  // namespace geometry
}  // namespace open3d
