// A caller written against the reference's C++ API (class names, include paths and call shapes of
// examples/cpp/segment_plane_iterative.cpp and examples/cpp/ransac_and_boundary.cpp), compiled
// unchanged against the B200 facade headers.  Reads a cloud (n, then n x 3 doubles) from a binary
// file, fits a plane with a fixed seed, segments planes, prints the results as text.
#include <misc3d/common/ransac.h>
#include <misc3d/logging.h>
#include <misc3d/registration/correspondence_matching.h>
#include <misc3d/segmentation/iterative_plane_segmentation.h>

#include <cstdio>
#include <cstdlib>
#include <vector>

int main(int argc, char **argv) {
    if (argc < 3) return 2;
    std::FILE *f = std::fopen(argv[1], "rb");
    if (!f) return 3;
    unsigned long long n = 0;
    if (std::fread(&n, sizeof n, 1, f) != 1) return 4;
    misc3d::PointCloud pc;
    pc.points_.resize(n);
    if (std::fread(pc.points_.data(), sizeof(double) * 3, n, f) != n) return 5;
    std::fclose(f);
    const unsigned seed = (unsigned)std::atoi(argv[2]);
    misc3d::SetVerbosityLevel(misc3d::VerbosityLevel::Error);

    misc3d::common::RANSACPlane fit;
    fit.SetMaxIteration(100);
    fit.SetProbability(0.9999);
    fit.SetPointCloud(pc);
    fit.SetSeed(seed); /* the only call the reference does not have */
    misc3d::common::Plane plane;
    std::vector<size_t> inliers;
    const bool ret = fit.FitModel(0.01, plane, inliers);
    std::printf("fit %d %zu %.17g %.17g %.17g %.17g\n", ret ? 1 : 0, inliers.size(), plane.parameters_[0],
                plane.parameters_[1], plane.parameters_[2], plane.parameters_[3]);
    unsigned long long h = 1469598103934665603ull; /* FNV-1a of the index list */
    for (size_t i : inliers) h = (h ^ (unsigned long long)i) * 1099511628211ull;
    std::printf("hash %llu\n", h);

    const auto clusters = misc3d::segmentation::SegmentPlaneIterative(pc, 0.01, 100, 0.1, &seed);
    std::printf("planes %zu\n", clusters.size());
    for (const auto &c : clusters)
        std::printf("plane %zu %.17g %.17g %.17g %.17g\n", c.second.points_.size(), c.first[0], c.first[1], c.first[2],
                    c.first[3]);

    { /* KNearestSearch (knn.h:24-73): 5 nearest points of point 0 in the cloud; the first is the point itself */
        misc3d::common::KNearestSearch knn(pc);
        std::vector<size_t> idx;
        std::vector<double> dist;
        const std::vector<double> q = {pc.points_[0][0], pc.points_[0][1], pc.points_[0][2]};
        const int k = knn.SearchKNN(q, 5, idx, dist);
        std::printf("knn %d", k);
        for (int i = 0; i < k; ++i) std::printf(" %zu %.17g", idx[i], dist[i]);
        std::printf("\n");
        const int kh = knn.SearchHybrid(q, dist.empty() ? 1.0 : dist.back(), 5, idx, dist); /* drops the last one (knn.cpp:129) */
        std::printf("hybrid %d\n", kh);
    }

    { /* ANNMatcher on host descriptors and (extension) on descriptors uploaded to the device: same answer.  The
       * "descriptors" are the first 2000 points (dim 3) against themselves shifted by one */
        const size_t m = pc.points_.size() < 2000 ? pc.points_.size() : 2000;
        std::vector<double> a(3 * m), b(3 * m);
        for (size_t i = 0; i < m; ++i)
            for (int c = 0; c < 3; ++c) {
                a[3 * i + c] = pc.points_[i][c];
                b[3 * i + c] = pc.points_[(i + 1) % m][c];
            }
        misc3d::registration::ANNMatcher matcher(misc3d::registration::MatchMethod::FLANN);
        const misc3d::FeatureMatrix fa{3, m, a.data()}, fb{3, m, b.data()};
        const auto host = matcher.Match(fa, fb);
        const auto da = misc3d::registration::DeviceFeature::Upload(fa), db = misc3d::registration::DeviceFeature::Upload(fb);
        const auto dev = matcher.Match(da, db);
        std::printf("devmatch %d %zu %d\n", host == dev ? 1 : 0, dev.first.size(), da.Download() == a ? 1 : 0);
    }

    try { /* the reference throws std::runtime_error from LogError (ransac.h:483-485) */
        fit.SetProbability(1.5);
        std::printf("throw 0\n");
    } catch (const std::runtime_error &) {
        std::printf("throw 1\n");
    }
    return 0;
}
