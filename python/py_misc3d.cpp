/*
 * py_misc3d.cpp -- pybind11 shim of the B200 build: module `misc3d` with the submodules, function
 * names, argument names, defaults and return types of the reference's python binding for the
 * RANSAC / segmentation / registration path:
 *   common.fit_plane / fit_sphere / fit_cylinder            (reference python/py_common.cpp:11-78)
 *   segmentation.segment_plane_iterative                     (python/py_segmentation.cpp:87-96)
 *   registration.match_correspondence (2 overloads), compute_transformation_ransac,
 *   compute_transformation_least_square, MatchMethod        (python/py_registration.cpp:12-106)
 *   VerbosityLevel, set_verbosity_level, get_verbosity_level (python/py_misc3d.cpp:52-62)
 *
 * Point clouds: the reference takes open3d.geometry.PointCloud.  Here any object with a `.points`
 * (and optional `.normals`) attribute convertible by numpy.asarray is accepted -- a real Open3D
 * cloud where Open3D is installed -- as well as a plain (N,3) float64 ndarray or a (points,
 * normals) tuple.  Features: any object with a `.data` attribute of shape (dim, n), or such an
 * ndarray.  Extra keyword-only argument `seed` (default None = std::random_device, as the
 * reference) fixes the sample stream.
 */
#include <pybind11/numpy.h>
#include <pybind11/pybind11.h>
#include <pybind11/stl.h>

#include <cstring>

#include <misc3d/common/ransac.h>
#include <misc3d/logging.h>
#include <misc3d/registration/correspondence_matching.h>
#include <misc3d/registration/transform_estimation.h>
#include <misc3d/segmentation/iterative_plane_segmentation.h>

namespace py = pybind11;
using namespace misc3d;
using DArray = py::array_t<double, py::array::c_style | py::array::forcecast>;
using FArray = py::array_t<double, py::array::f_style | py::array::forcecast>;

namespace {

std::vector<Vector3d> rows3(const py::handle &h, const char *what) {
    DArray a = DArray::ensure(h);
    if (!a || !((a.ndim() == 2 && a.shape(1) == 3) || a.size() == 0))
        throw py::type_error(std::string(what) + " must be convertible to an (N, 3) float64 array");
    std::vector<Vector3d> out((size_t)(a.size() / 3));
    if (!out.empty()) std::memcpy(out[0].data(), a.data(), sizeof(double) * 3 * out.size());
    return out;
}

/* returns the cloud and whether the caller passed an Open3D-like object (has .points) */
PointCloud cloud_from_py(const py::object &o, bool *is_o3d = nullptr) {
    PointCloud pc;
    if (is_o3d) *is_o3d = false;
    if (py::hasattr(o, "points")) {
        if (is_o3d) *is_o3d = true;
        pc.points_ = rows3(py::module_::import("numpy").attr("asarray")(o.attr("points")), "pc.points");
        if (py::hasattr(o, "normals")) {
            auto nr = rows3(py::module_::import("numpy").attr("asarray")(o.attr("normals")), "pc.normals");
            if (nr.size() == pc.points_.size()) pc.normals_ = std::move(nr);
        }
    } else if (py::isinstance<py::tuple>(o) && py::len(o) == 2) {
        py::tuple t = o.cast<py::tuple>();
        pc.points_ = rows3(t[0], "points");
        if (!t[1].is_none()) pc.normals_ = rows3(t[1], "normals");
    } else {
        pc.points_ = rows3(o, "pc");
    }
    return pc;
}

py::array_t<double> to_numpy(const std::vector<double> &v) {
    py::array_t<double> a((py::ssize_t)v.size());
    std::memcpy(a.mutable_data(), v.data(), sizeof(double) * v.size());
    return a;
}

/* (points, normals) of a cloud argument as float64 C-contiguous (N, 3) arrays WITHOUT a copy when the caller's arrays
 * already have that layout (numpy arrays; Open3D's Vector3dVector converts through numpy.asarray) */
std::pair<DArray, DArray> cloud_arrays(const py::object &o, bool want_normals) {
    py::object np = py::module_::import("numpy");
    py::object pts, nrm = py::none();
    if (py::hasattr(o, "points")) {
        pts = np.attr("asarray")(o.attr("points"));
        if (want_normals && py::hasattr(o, "normals")) nrm = np.attr("asarray")(o.attr("normals"));
    } else if (py::isinstance<py::tuple>(o) && py::len(o) == 2) {
        py::tuple t = o.cast<py::tuple>();
        pts = t[0];
        if (want_normals && !t[1].is_none()) nrm = t[1];
    } else {
        pts = o;
    }
    DArray a = DArray::ensure(pts);
    if (!a || !((a.ndim() == 2 && a.shape(1) == 3) || a.size() == 0))
        throw py::type_error("pc must be convertible to an (N, 3) float64 array");
    DArray b;
    if (!nrm.is_none()) {
        b = DArray::ensure(nrm);
        if (!b || b.size() != a.size()) b = DArray();
    }
    return {a, b};
}

/* FitPlane / FitSphere / FitCylinder (python/py_common.cpp:11-67): SetMaxIteration, SetProbability, SetPointCloud,
 * FitModel -- the same sequence as the facade's RANSAC<> class, but on the caller's buffers (the reference's
 * SetPointCloud deep copy has no observable effect here: the call does not return before the fit is done) */
template <int KIND>
std::tuple<py::array_t<double>, std::vector<size_t>> fit_primitive(const py::object &pc_obj, double threshold,
                                                                   size_t max_iteration, double probability,
                                                                   const py::object &seed, bool need_normals) {
    auto arrays = cloud_arrays(pc_obj, need_normals);
    const DArray &pts = arrays.first, &nrm = arrays.second;
    const size_t n = (size_t)(pts.size() / 3);
    const bool has_normals = nrm && (size_t)(nrm.size() / 3) == n && n > 0;
    if (need_normals && !has_normals) LogError("Fit cylinder requires normals."); /* py_common.cpp:50-52 */
    if (probability <= 0 || probability > 1) LogError("Probability must be > 0 or <= 1.0"); /* ransac.h:482-487 */
    m3d_ransac_params p{};
    p.threshold = threshold;
    p.max_iteration = max_iteration;
    p.probability = probability;
    p.seed = seed.is_none() ? b200::RandomSeed() : seed.cast<uint32_t>();
    constexpr size_t NP = KIND == M3D_CYLINDER ? 7 : 4;
    double out[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    std::vector<size_t> inliers(n);
    size_t n_inl = 0;
    m3d_ransac_stats st{};
    int rc;
    m3d_ctx *ctx = b200::DefaultContext();
    {
        py::gil_scoped_release nogil;
        rc = m3d_ransac_fit(ctx, KIND, n ? pts.data() : nullptr, has_normals ? nrm.data() : nullptr, n, &p, out,
                            inliers.data(), &n_inl, &st);
    }
    inliers.resize(n_inl);
    if (rc < 0) b200::Raise(ctx); /* lack of points: ransac.h:510-513 throws */
    {
        char line[160];
        std::snprintf(line, sizeof line, "Find best model with %g%% inliers and run %llu iterations",
                      n ? 100.0 * (double)st.best_count / (double)n : 0.0, (unsigned long long)st.iterations_run);
        LogInfo(line); /* ransac.h:616-619 */
    }
    std::vector<double> params(rc == 1 ? NP : 4, 0.0); /* failure: setZero(4), also for the cylinder (py_common.cpp:62) */
    if (rc == 1) params.assign(out, out + NP);
    return std::make_tuple(to_numpy(params), std::move(inliers));
}

FArray feature_from_py(const py::object &o) {
    py::object src = py::hasattr(o, "data") && !py::isinstance<py::array>(o) ? py::object(o.attr("data")) : o;
    FArray a = FArray::ensure(src);
    if (!a || a.ndim() != 2) throw py::type_error("descriptors must be a (dim, n) float64 array");
    return a;
}

std::pair<std::vector<size_t>, std::vector<size_t>> match(const py::object &src, const py::object &dst,
                                                         const registration::MatchMethod &method, int n_trees) {
    if (py::isinstance<registration::DeviceFeature>(src) && py::isinstance<registration::DeviceFeature>(dst)) {
        /* extension: both descriptor sets already on the device */
        const auto &da = src.cast<const registration::DeviceFeature &>();
        const auto &db = dst.cast<const registration::DeviceFeature &>();
        registration::ANNMatcher matcher(method, n_trees);
        py::gil_scoped_release nogil;
        return matcher.Match(da, db);
    }
    FArray a = feature_from_py(src), b = feature_from_py(dst);
    registration::ANNMatcher matcher(method, n_trees);
    FeatureMatrix fa{(int)a.shape(0), (size_t)a.shape(1), a.data()};
    FeatureMatrix fb{(int)b.shape(0), (size_t)b.shape(1), b.data()};
    py::gil_scoped_release nogil;
    return matcher.Match(fa, fb);
}

py::array_t<double> mat4(const Matrix4d &T) {
    py::array_t<double> a({4, 4});
    std::memcpy(a.mutable_data(), T.data(), sizeof(double) * 16);
    return a;
}

}  // namespace

PYBIND11_MODULE(py_misc3d, m) {
    m.doc() = "Misc3D RANSAC / segmentation / registration path, B200 (sm_100a) build";

    py::module_ common = m.def_submodule("common");
    common.def(
        "fit_plane",
        [](const py::object &pc, double threshold, size_t max_iteration, double probability, const py::object &seed) {
            return fit_primitive<M3D_PLANE>(pc, threshold, max_iteration, probability, seed, false);
        },
        "Fit a plane from point clouds", py::arg("pc"), py::arg("threshold") = 0.01, py::arg("max_iteration") = 1000,
        py::arg("probability") = 0.9999, py::kw_only(), py::arg("seed") = py::none());
    common.def(
        "fit_sphere",
        [](const py::object &pc, double threshold, size_t max_iteration, double probability, const py::object &seed) {
            return fit_primitive<M3D_SPHERE>(pc, threshold, max_iteration, probability, seed, false);
        },
        "Fit a sphere from point clouds", py::arg("pc"), py::arg("threshold") = 0.01, py::arg("max_iteration") = 1000,
        py::arg("probability") = 0.9999, py::kw_only(), py::arg("seed") = py::none());
    common.def(
        "fit_cylinder",
        [](const py::object &pc, double threshold, size_t max_iteration, double probability, const py::object &seed) {
            return fit_primitive<M3D_CYLINDER>(pc, threshold, max_iteration, probability, seed, true);
        },
        "Fit a cylinder from point clouds", py::arg("pc"), py::arg("threshold") = 0.01,
        py::arg("max_iteration") = 1000, py::arg("probability") = 0.9999, py::kw_only(),
        py::arg("seed") = py::none());

    py::module_ seg = m.def_submodule("segmentation");
    seg.def(
        "segment_plane_iterative",
        [](const py::object &pcd, const double threshold, const int max_iteration, const double min_ratio,
           const py::object &seed) {
            bool is_o3d = false;
            PointCloud pc = cloud_from_py(pcd, &is_o3d);
            uint32_t s = 0;
            if (!seed.is_none()) s = seed.cast<uint32_t>();
            std::vector<std::pair<Vector4d, PointCloud>> res;
            {
                py::gil_scoped_release nogil;
                res = segmentation::SegmentPlaneIterative(pc, threshold, max_iteration, min_ratio,
                                                          seed.is_none() ? nullptr : &s);
            }
            py::list out;
            py::object o3d_cloud = py::none(), o3d_vec = py::none();
            if (is_o3d) { /* hand clusters back as open3d.geometry.PointCloud when Open3D is there */
                try {
                    py::module_ o3d = py::module_::import("open3d");
                    o3d_cloud = o3d.attr("geometry").attr("PointCloud");
                    o3d_vec = o3d.attr("utility").attr("Vector3dVector");
                } catch (py::error_already_set &) {
                    o3d_cloud = py::none();
                }
            }
            for (auto &pr : res) {
                py::array_t<double> plane(4);
                std::memcpy(plane.mutable_data(), pr.first.data(), sizeof(double) * 4);
                const size_t k = pr.second.points_.size();
                py::array_t<double> pts({(py::ssize_t)k, (py::ssize_t)3});
                if (k) std::memcpy(pts.mutable_data(), pr.second.points_[0].data(), sizeof(double) * 3 * k);
                py::object cluster = pts;
                if (!o3d_cloud.is_none()) cluster = o3d_cloud(o3d_vec(pts));
                out.append(py::make_tuple(plane, cluster));
            }
            return out;
        },
        "Segment plane iteratively using RANSAC plane fitting", py::arg("pcd"), py::arg("threshold"),
        py::arg("max_iteration") = 100, py::arg("min_ratio") = 0.05, py::kw_only(), py::arg("seed") = py::none());

    py::module_ reg = m.def_submodule("registration");
    py::enum_<registration::MatchMethod>(reg, "MatchMethod")
        .value("FLANN", registration::MatchMethod::FLANN)
        .value("ANNOY", registration::MatchMethod::ANNOY)
        .export_values();
    reg.def("match_correspondence", &match, "Match corresponding point clouds (exact brute-force search on the GPU)",
            py::arg("src"), py::arg("dst"), py::arg("method") = registration::MatchMethod::ANNOY,
            py::arg("n_trees") = 4);
    reg.def(
        "compute_transformation_ransac",
        [](const py::object &src, const py::object &dst,
           const std::pair<std::vector<size_t>, std::vector<size_t>> &corres, double threshold, int max_iter,
           double edge_length_threshold, const py::object &seed) {
            PointCloud s = cloud_from_py(src), d = cloud_from_py(dst);
            registration::RANSACSolver solver(threshold, max_iter, edge_length_threshold);
            if (!seed.is_none()) solver.SetSeed(seed.cast<uint32_t>());
            Matrix4d T;
            {
                py::gil_scoped_release nogil;
                T = solver.Solve(s, d, corres);
            }
            return mat4(T);
        },
        "Compute 3D rigid transformation from corresponding point clouds using RANSAC", py::arg("src"),
        py::arg("dst"), py::arg("corres"), py::arg("threshold") = 0.01, py::arg("max_iter") = 100000,
        py::arg("edge_length_threshold") = 0.9, py::kw_only(), py::arg("seed") = py::none());
    reg.def(
        "compute_transformation_least_square",
        /* both reference overloads (python/py_registration.cpp:12-31): PointCloud pair or (n, 3) ndarrays, scaling=False */
        [](const py::object &src, const py::object &dst, bool scaling) {
            registration::LeastSquareSolver solver(scaling);
            const auto s = rows3(py::hasattr(src, "points") ? py::object(src.attr("points")) : src, "src");
            const auto d = rows3(py::hasattr(dst, "points") ? py::object(dst.attr("points")) : dst, "dst");
            Matrix4d T;
            {
                py::gil_scoped_release nogil;
                T = solver.Solve(s, d);
            }
            return mat4(T);
        },
        "Compute 3D transformation from corresponding point clouds using Least-Square method (point clouds or "
        "numpy arrays with shape (n, 3))",
        py::arg("src"), py::arg("dst"), py::arg("scaling") = false);

    /* extensions (not in the reference's module) */
    reg.def(
        "refine_transformation_on_inliers",
        [](const py::object &src, const py::object &dst,
           const std::pair<std::vector<size_t>, std::vector<size_t>> &corres, const py::array_t<double> &T,
           double threshold, bool scaling) {
            PointCloud s = cloud_from_py(src), d = cloud_from_py(dst);
            if (T.ndim() != 2 || T.shape(0) != 4 || T.shape(1) != 4) throw py::value_error("T must have shape (4, 4)");
            Matrix4d Tin{};
            for (int r = 0; r < 4; ++r)
                for (int c = 0; c < 4; ++c) Tin.data()[4 * r + c] = T.at(r, c);
            Matrix4d out;
            {
                py::gil_scoped_release nogil;
                out = registration::RefineOnInlierCorrespondences(s, d, corres, Tin, threshold, scaling);
            }
            return mat4(out);
        },
        "Least-squares (Umeyama) refit of a transformation on the correspondences that are its inliers "
        "(|T src - dst| < threshold), on the GPU",
        py::arg("src"), py::arg("dst"), py::arg("corres"), py::arg("T"), py::arg("threshold") = 0.01,
        py::arg("scaling") = false);
    /* ... and the Open3D steps the reference's callers run around this path, on the GPU */
    py::class_<registration::DeviceFeature>(reg, "DeviceFeature",
                                            "descriptors that live on the GPU: accepted by match_correspondence (both arguments)")
        .def("dimension", &registration::DeviceFeature::Dimension)
        .def("num", &registration::DeviceFeature::Num)
        .def_property_readonly(
            "data",
            [](const registration::DeviceFeature &f) {
                std::vector<double> v = f.Download();
                py::array_t<double, py::array::f_style> a({(py::ssize_t)f.Dimension(), (py::ssize_t)f.Num()});
                if (!v.empty()) std::memcpy(a.mutable_data(), v.data(), sizeof(double) * v.size());
                return a;
            },
            "(dim, n) float64 copy on the host (like open3d's Feature.data)");
    reg.def(
        "compute_fpfh_feature_device",
        [](const py::object &pcd, double radius, int max_nn) {
            PointCloud pc = cloud_from_py(pcd);
            py::gil_scoped_release nogil;
            return registration::DeviceFeature::FPFH(pc, radius, max_nn);
        },
        "compute_fpfh_feature with the (33, n) result left on the GPU as a DeviceFeature", py::arg("pcd"), py::arg("radius"),
        py::arg("max_nn") = 100);
    reg.def(
        "compute_fpfh_feature",
        [](const py::object &pcd, double radius, int max_nn) {
            PointCloud pc = cloud_from_py(pcd);
            std::vector<double> f;
            {
                py::gil_scoped_release nogil;
                f = registration::ComputeFPFHFeature(pc, radius, max_nn);
            }
            py::array_t<double, py::array::f_style> a({(py::ssize_t)33, (py::ssize_t)pc.points_.size()});
            if (!f.empty()) std::memcpy(a.mutable_data(), f.data(), sizeof(double) * f.size());
            return a;
        },
        "open3d.pipelines.registration.compute_fpfh_feature(pcd, KDTreeSearchParamHybrid(radius, max_nn)) on the GPU: "
        "(33, n) float64 array, usable as match_correspondence input",
        py::arg("pcd"), py::arg("radius"), py::arg("max_nn") = 100);
    reg.def(
        "registration_icp",
        [](const py::object &src, const py::object &dst, double max_correspondence_distance, const py::object &init,
           int max_iteration, double relative_fitness, double relative_rmse) {
            PointCloud s = cloud_from_py(src), d = cloud_from_py(dst);
            Matrix4d T0{1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};
            if (!init.is_none()) {
                DArray a = DArray::ensure(init);
                if (!a || a.size() != 16) throw py::type_error("init must be a (4, 4) float64 array");
                std::memcpy(T0.data(), a.data(), sizeof(double) * 16);
            }
            registration::ICPResult r;
            {
                py::gil_scoped_release nogil;
                r = registration::RegistrationICP(s, d, max_correspondence_distance, T0, max_iteration, relative_fitness,
                                                  relative_rmse);
            }
            return py::make_tuple(mat4(r.transformation_), r.fitness_, r.inlier_rmse_, r.iterations_);
        },
        "open3d.pipelines.registration.registration_icp (point to point) on the GPU: (T, fitness, inlier_rmse, iterations)",
        py::arg("src"), py::arg("dst"), py::arg("max_correspondence_distance"), py::arg("init") = py::none(),
        py::arg("max_iteration") = 30, py::arg("relative_fitness") = 1e-6, py::arg("relative_rmse") = 1e-6);

    py::enum_<VerbosityLevel>(m, "VerbosityLevel", py::arithmetic(), "VerbosityLevel")
        .value("Error", VerbosityLevel::Error)
        .value("Warning", VerbosityLevel::Warning)
        .value("Info", VerbosityLevel::Info)
        .value("Debug", VerbosityLevel::Debug)
        .export_values();
    m.def("set_verbosity_level", &SetVerbosityLevel, "Set global verbosity level of Misc3D",
          py::arg("verbosity_level"));
    m.def("get_verbosity_level", &GetVerbosityLevel, "Get global verbosity level of Misc3D");
}
