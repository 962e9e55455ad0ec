/*
 * m3d_capi.h -- C-ABI of libm3d_b200.so: the B200 (sm_100a) implementation of
 * Misc3D's RANSAC primitive fitting / iterative plane segmentation /
 * correspondence matching / correspondence-RANSAC registration hot path.
 *
 * The reference (yuecideng/Misc3D) has no FFI layer for this path: its boundary
 * is a C++ class API plus a pybind11 module.  Every entry point below states
 * the reference interface it replaces (paths relative to the reference tree).
 * A C++ facade with the reference's class names sits on top of this ABI in
 * include/misc3d/ and the pybind11 shim in python/ keeps the `misc3d` module
 * signatures.
 *
 * Conventions
 *  - plain pointers and sizes only; every buffer is caller-owned; the library
 *    never returns owned memory except opaque handles with a matching *_free;
 *  - blocking calls; one m3d_ctx per host thread (a ctx owns a CUDA stream,
 *    its scratch arena and, optionally, a NCCL communicator);
 *  - return value: M3D_OK (0) or a negative m3d_status; no exceptions cross
 *    the ABI; m3d_last_error() gives a human-readable message;
 *  - every function that computes needs a CUDA device: there is NO CPU
 *    fallback (M3D_ERR_CUDA is returned when no device is usable);
 *  - point clouds are N x 3 float64 AoS, exactly the memory of the reference's
 *    std::vector<Eigen::Vector3d> (open3d::geometry::PointCloud::points_);
 *    descriptors are dim x count float64 column-major (Eigen::MatrixXd /
 *    open3d Feature::data_); indices are size_t (std::vector<size_t>).
 */
#ifndef M3D_CAPI_H_
#define M3D_CAPI_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define M3D_ABI_VERSION 2

typedef enum m3d_status {
    M3D_OK = 0,
    M3D_ERR_INVALID_ARG = -1,    /* null pointer, bad enum, bad capacity                       */
    M3D_ERR_TOO_FEW_POINTS = -2, /* reference: LogError "Can not fit model due to lack of points"
                                    (ransac.h:510-513), registration <3 points
                                    (transform_estimation.cpp:130-133) -> throws               */
    M3D_ERR_PROBABILITY = -3,    /* reference: SetProbability throws (ransac.h:482-487)        */
    M3D_ERR_NO_NORMALS = -4,     /* reference: fit_cylinder without normals throws
                                    (py_common.cpp:50-52, ransac.h:356-359)                    */
    M3D_ERR_CUDA = -5,           /* CUDA runtime / no device / launch failure                  */
    M3D_ERR_NCCL = -6,           /* NCCL not loadable or a collective failed                   */
    M3D_ERR_NO_INLIERS = -7,     /* segmentation round with zero inliers (reference loops
                                    forever)                              */
    M3D_ERR_CAPACITY = -8,       /* an output capacity given by the caller is too small        */
    M3D_ERR_INTERNAL = -9        /* self-check failed (fast count != exact count)              */
} m3d_status;

typedef enum m3d_primitive { M3D_PLANE = 0, M3D_SPHERE = 1, M3D_CYLINDER = 2 } m3d_primitive;

/* reference: misc3d::registration::MatchMethod (correspondence_matching.h:14-17).  Both values
 * run the exact brute-force search on the GPU (equal to FLANN up to ties, a superset in quality
 * of ANNOY); ties resolve to the lowest index. */
typedef enum m3d_match_method { M3D_MATCH_FLANN = 0, M3D_MATCH_ANNOY = 1 } m3d_match_method;

typedef struct m3d_ctx m3d_ctx;     /* stream + scratch + (optional) communicator */
typedef struct m3d_cloud m3d_cloud; /* a point cloud resident in HBM              */
typedef struct m3d_knn m3d_knn;     /* an exact k-NN index resident in HBM        */

/* ------------------------------------------------------------------ context */
int m3d_abi_version(void);
/* number of visible CUDA devices (0 when there is no driver / GPU); never fails */
int m3d_device_count(void);
int m3d_ctx_create(int device, m3d_ctx **out);
/* same, but all work is enqueued on a caller-owned cudaStream_t (e.g. torch's current stream) so
 * that the caller can time it with events recorded on that stream */
int m3d_ctx_create_on_stream(int device, void *cuda_stream, m3d_ctx **out);
void m3d_ctx_destroy(m3d_ctx *ctx);
const char *m3d_last_error(const m3d_ctx *ctx);
/* cudaStream_t the context launches on */
void *m3d_ctx_stream(const m3d_ctx *ctx);
/* kernels launched by this context since creation (bench.py's gpu_launches) */
uint64_t m3d_ctx_launch_count(const m3d_ctx *ctx);
/* measured fp32 FFMA issue rate of the device (FFMA lane-operations per second): the ALU roofline
 * denominator bench.py reports beside the HBM one */
int m3d_probe_fp32_ffma(m3d_ctx *ctx, double *ffma_per_s);
/* the same for fp64 (DFMA lane-operations per second): the resolve / refine kernels' ALU roofline */
int m3d_probe_fp64_dfma(m3d_ctx *ctx, double *dfma_per_s);

/* work counters of the scoring launches run with M3D_FLAG_STATS since the last call (read and cleared):
 * out[0] (hypothesis, tile) bounding-sphere tests, [1] survivors, [2] surviving (hypothesis, cell) pairs
 * (x32 = point-hypothesis pairs evaluated point by point), [3] / [4] evaluation passes with one / two
 * hypotheses per lane, [5] guard-band re-scans.  bench.py derives the kernel's binding-resource figure
 * from these. */
int m3d_score_stats(m3d_ctx *ctx, uint64_t out[8]);

/* drops every host-buffer registration made under M3D_FLAG_REGISTER_HOST (also done by m3d_ctx_destroy) */
void m3d_host_unregister_all(m3d_ctx *ctx);

/* -------- multi-GPU: hypothesis sharding.  Every rank calls the fit with the SAME cloud
 * and parameters; a wave of hypotheses of the one global sample table is dealt to the ranks in cyclic
 * blocks of 256 rows (m3d_shard_rows).  probability == 1: one all-gather of a 64-byte best record per
 * rank; probability < 1: one all-gather of the per-hypothesis inlier counts per wave; then every rank
 * finishes identically, so results do not depend on R.  With a communicator from m3d_ctx_init_nccl,
 * m3d_ransac_fit also uploads the (replicated) host cloud cooperatively: rank r copies the r-th 1/R
 * slice over its own PCIe link and the slices are all-gathered over NVLink. */
#define M3D_NCCL_ID_BYTES 128
int m3d_nccl_unique_id(char id[M3D_NCCL_ID_BYTES]);
int m3d_ctx_init_nccl(m3d_ctx *ctx, const char id[M3D_NCCL_ID_BYTES], int rank, int world);
/* alternative exchange for callers that own the communicator (torch.distributed): the callback
 * must all-gather `bytes_per_rank` bytes from every rank's `send` into `recv` (rank-major);
 * both are DEVICE pointers valid on the ctx stream when on_device != 0, host pointers otherwise. */
typedef int (*m3d_allgather_fn)(void *user, const void *send, void *recv, size_t bytes_per_rank,
                                int on_device);
int m3d_ctx_set_exchange(m3d_ctx *ctx, m3d_allgather_fn fn, void *user, int on_device, int rank,
                         int world);

/* -------------------------------------------------------------- RANSAC fit */
typedef struct m3d_ransac_params {
    double threshold;       /* FitModel(threshold, ...) ransac.h:506                         */
    uint64_t max_iteration; /* SetMaxIteration (ransac.h:495), default 1000 (ransac.h:461)   */
    double probability;     /* SetProbability (ransac.h:482), default 0.9999 (ransac.h:462); */
                            /* 1.0 disables the adaptive early exit (ransac.h:601-610)       */
    uint32_t seed;          /* seed of the mt19937 sample stream; the reference seeds from   */
                            /* std::random_device (utils.h:74-77)                            */
    uint32_t flags;         /* M3D_FLAG_*                                                    */
} m3d_ransac_params;

#define M3D_FLAG_EXACT_ONLY 1u /* score with the fp64 reference-order kernel only (slow; debug) */
#define M3D_FLAG_NO_REFIT 2u   /* skip RefineModel's GeneralFit (model_out = minimal model)     */
#define M3D_FLAG_DENSE 4u      /* score every point-hypothesis pair (no bounding-sphere culling) */
#define M3D_FLAG_CLASSIFY 8u   /* round-1 culling kernel (M3D_SCORE_PATH=cull) only: always pre-sort the hypotheses into
                                  culled / dense ones (default there: launches of >= 12288 rows) */
#define M3D_FLAG_CHUNKED_UPLOAD 32u /* m3d_ransac_fit, pinned buffers, probability 1: score the cloud chunk by chunk behind
                                       its upload (experimental: measured slower than the default) */
#define M3D_FLAG_PLAIN_UPLOAD 64u   /* m3d_ransac_fit: upload, prepare, then fit, all on one stream (the round-1 path) */
#define M3D_FLAG_REGISTER_HOST 128u /* m3d_ransac_fit: page-lock the caller's (pageable) xyz buffer in place with
                                      * cudaHostRegister the first time it is seen, so that this and later fits of the
                                      * same buffer upload by DMA at the pinned rate.  The registration lives until
                                      * m3d_host_unregister_all / m3d_ctx_destroy; THE CALLER KEEPS THE BUFFER ALLOCATED
                                      * UNTIL THEN.  Without the flag pageable buffers are staged through pinned memory */
#define M3D_FLAG_STATS 16u     /* run the counting build of the scoring kernel (slower); read with m3d_score_stats */

typedef struct m3d_ransac_stats {
    uint64_t best_index;     /* loop index i of the winning minimal model                      */
    uint64_t best_count;     /* its inlier count in EvaluateModel (ransac.h:626-654)           */
    double best_rmse;        /* error / sqrt(count) (ransac.h:650)                             */
    uint64_t iterations_run; /* `count` printed by ransac.h:616-619                            */
    uint64_t stop_index;     /* first i skipped by ransac.h:573 (== max_iteration if none)     */
    uint64_t evaluated;      /* hypotheses the GPU actually scored (>= stop_index)             */
    uint64_t exact_resolves; /* point-hypothesis pairs decided by the fp64 reference-order path */
    int32_t found;           /* 1 if any hypothesis ever became best                           */
    int32_t refit_ok;        /* GeneralFit's return value (ransac.h:548)                       */
    float device_ms;         /* CUDA-event time of the whole fit on the ctx stream             */
    float score_ms;          /* CUDA-event time of the scoring kernel(s) alone                 */
    float refine_ms;         /* CUDA-event time of the RefineModel passes (ransac.h:534-549):
                                2 x 24 N bytes read + 8 n_inl written -- the HBM-bound part        */
    float draw_ms;           /* CUDA-event time of the device-side sample draw (0: host draw)  */
} m3d_ransac_stats;

/* Replaces RANSAC<Estimator,Model,Sampler>::{SetPointCloud, SetProbability, SetMaxIteration,
 * FitModel} (include/misc3d/common/ransac.h:469-516) and therefore FitPlane / FitSphere /
 * FitCylinder of python/py_common.cpp:11-67.
 *   xyz      n x 3 float64 host;  nrm n x 3 float64 host or NULL (required for M3D_CYLINDER)
 *   model_out  8 doubles (4 used for plane/sphere, 7 for cylinder; rest zero)
 *   inl_out    capacity n (may be NULL to skip the copy), ascending; *n_inl always written
 * returns 1 = FitModel true, 0 = FitModel false (no model / GeneralFit failed), <0 = m3d_status */
int m3d_ransac_fit(m3d_ctx *ctx, int kind, const double *xyz, const double *nrm, size_t n,
                   const m3d_ransac_params *p, double *model_out, size_t *inl_out, size_t *n_inl,
                   m3d_ransac_stats *stats);

/* The same against a cloud already resident in HBM (what SetPointCloud's deep copy,
 * ransac.h:469-475, becomes). */
int m3d_cloud_upload(m3d_ctx *ctx, const double *xyz, const double *nrm, size_t n, m3d_cloud **out);
/* d_xyz / d_nrm are DEVICE pointers (n x 3 float64); they are copied, not adopted */
int m3d_cloud_from_device(m3d_ctx *ctx, const double *d_xyz, const double *d_nrm, size_t n,
                          m3d_cloud **out);
void m3d_cloud_free(m3d_cloud *cloud);
size_t m3d_cloud_size(const m3d_cloud *cloud);
int m3d_ransac_fit_cloud(m3d_ctx *ctx, int kind, const m3d_cloud *cloud,
                         const m3d_ransac_params *p, double *model_out, size_t *inl_out,
                         size_t *n_inl, m3d_ransac_stats *stats);

/* Building blocks (also what the parity tests probe one by one).
 * models: rows x 8 doubles, valid: rows bytes, counts: rows uint64 (all host).
 * m3d_score_samples: MinimalFit (ransac.h:138-162 / 239-294 / 354-417) of every row of a caller-
 * supplied sample table (rows x k uint32, draw order; the kernel re-orders each row ascending as
 * SelectByIndex does, ransac.h:578) + EvaluateModel's inlier count (ransac.h:626-654). */
int m3d_score_samples(m3d_ctx *ctx, int kind, const m3d_cloud *cloud, const uint32_t *samples,
                      size_t rows, double threshold, uint32_t flags, double *models,
                      uint8_t *valid, uint64_t *counts);
/* EvaluateModel for one explicit model: count and sum of distances (parallel fp64 sum, or the
 * reference's sequential index-order sum when sequential != 0). */
int m3d_evaluate_model(m3d_ctx *ctx, int kind, const m3d_cloud *cloud, const double *model,
                       double threshold, int sequential, uint64_t *count, double *err);

/* host-only helpers (no GPU needed) --------------------------------------------------------- */
/* utils.h:81-97 RandomSampler<size_t>::operator() on an mt19937(seed) stream: rows x k, draw order */
void m3d_sample_table(uint32_t seed, size_t n, int k, size_t rows, uint32_t *out);

/* the same table drawn ON THE DEVICE (what m3d_ransac_fit* use when probability == 1): bit-identical rows.
 * out is a HOST buffer of rows x k.  returns 1 = drawn on the device, 0 = the device draw is not eligible
 * for these sizes or gave up (too many duplicate draws, i.e. tiny clouds; the host draw is used then),
 * <0 = m3d_status.  Needs a GPU. */
int m3d_sample_table_device(m3d_ctx *ctx, uint32_t seed, size_t n, int k, size_t rows, uint32_t *out);

/* Hypothesis sharding (host only): the rows of a wave of `rows` hypotheses that rank `rank` of
 * `world` scores, in its local order (cyclic blocks of 256 rows).  out (may be NULL) receives the wave rows,
 * *n_local their number, *padded the per-rank stride of the all-gathered count buffer (same on all ranks). */
void m3d_shard_rows(size_t rows, int rank, int world, uint32_t *out, size_t *n_local, size_t *padded);
/* ransac.h:572-613 replayed over per-hypothesis results in loop order (the "ordered scan").
 * counts[i] is the inlier count of hypothesis i, valid[i] MinimalFit's return value.  err may be
 * NULL; when given, err[i] (sum of inlier distances) breaks count ties as inlier_rmse does.
 * Fills best_index/best_count/iterations_run/stop_index/found of *st. Returns 0. */
int m3d_ordered_scan(const uint64_t *counts, const uint8_t *valid, const double *err, size_t rows,
                     size_t n_points, int k, double probability, uint64_t max_iteration,
                     m3d_ransac_stats *st);

/* ---------------------------------------------------- iterative segmentation */
/* Replaces misc3d::segmentation::SegmentPlaneIterative
 * (src/iterative_plane_segmentation.cpp:7-39; python/py_segmentation.cpp:87-96).
 *   planes   cap_planes x 4 doubles; labels n entries = plane id or UINT64_MAX
 *   round r uses sample seed `seed + r` (the reference constructs a new random_device-seeded
 *   sampler per FitModel, ransac.h:570)
 * returns M3D_OK, M3D_ERR_TOO_FEW_POINTS (reference throws from ransac.h:510-513 when < 3 points
 * remain), M3D_ERR_NO_INLIERS, M3D_ERR_CAPACITY.  n < 3 -> M3D_OK with *n_planes = 0 (reference
 * warns and returns {} at :14-17). */
int m3d_segment_plane_iterative(m3d_ctx *ctx, const double *xyz, size_t n, double threshold,
                                int max_iteration, double min_ratio, uint32_t seed,
                                double *planes, size_t cap_planes, uint64_t *labels,
                                size_t *n_planes, float *device_ms);
/* the same with 32-bit labels (0xFFFFFFFF = unassigned): half the device-to-host bytes of the call */
int m3d_segment_plane_iterative_u32(m3d_ctx *ctx, const double *xyz, size_t n, double threshold,
                                    int max_iteration, double min_ratio, uint32_t seed,
                                    double *planes, size_t cap_planes, uint32_t *labels,
                                    size_t *n_planes, float *device_ms);

/* --------------------------------------------------- correspondence matching */
/* Replaces ANNMatcher::Match (src/correspondence_matching.cpp:52-84) / NearestSearch (:13-44);
 * python match_correspondence (python/py_registration.cpp:73-106).
 *   src: dim x ns, dst: dim x nd float64 column-major (host); idx0/idx1 capacity ns.
 * Mutual nearest neighbours in ascending source index. */
int m3d_match_correspondence(m3d_ctx *ctx, const double *src, size_t ns, const double *dst,
                             size_t nd, int dim, int method, int n_trees, size_t *idx0,
                             size_t *idx1, size_t *n_out, float *device_ms);
/* one direction only: nn[i] = argmin_j |src_i - dst_j|^2 (ties -> lowest j) */
int m3d_nearest(m3d_ctx *ctx, const double *src, size_t ns, const double *dst, size_t nd, int dim,
                size_t *nn, float *device_ms);

/* Replaces misc3d::common::KNearestSearch (include/misc3d/common/knn.h:24-73, src/knn.cpp:36-139; an
 * approximate Annoy forest there).  data: dim x n float64 column-major (Eigen::MatrixXd; a point cloud is
 * dim = 3).  m3d_knn_search answers nq queries (dim x nq column-major) with the EXACT k nearest items in
 * ascending distance (ties: lower index): idx_out / dist_out are nq x k (row q holds count_out[q] valid
 * entries), distances are Euclidean -- not squared -- as Annoy reports them (knn.cpp:103-113).  radius > 0
 * keeps only items with distance <= radius (SearchHybrid, knn.cpp:115-139). */
int m3d_knn_create(m3d_ctx *ctx, const double *data, int dim, size_t n, m3d_knn **out);
void m3d_knn_free(m3d_knn *index);
int m3d_knn_search(m3d_ctx *ctx, m3d_knn *index, const double *queries, size_t nq, int k, double radius,
                   size_t *idx_out, double *dist_out, int *count_out);

/* ------------------------------------------------------- RANSAC registration */
typedef struct m3d_reg_stats {
    uint64_t best_index;
    uint64_t best_count;
    double best_rmse;
    uint64_t evaluated;  /* hypotheses that passed both checkers (before stop_index) */
    uint64_t stop_index; /* first itr with itr >= est_k (== max_iter if none)        */
    float device_ms;
    float score_ms;
} m3d_reg_stats;

/* Replaces RANSACSolver::Solve (src/transform_estimation.cpp:124-164), i.e. Open3D's
 * RegistrationRANSACBasedOnCorrespondence with TransformationEstimationPointToPoint(false),
 * ransac_n = 3, checkers EdgeLength(edge_thr) + Distance(threshold), criteria(max_iter,
 * confidence); python compute_transformation_ransac (python/py_registration.cpp:55-67).
 * T_out: 16 doubles row-major.  Honours edge_thr (the reference's member is self-initialised,
 * transform_estimation.h:126 -- documented deviation).
 * returns 1 ok, 0 when Open3D would return its default result (m < 3 or threshold <= 0; T = I),
 * <0 = m3d_status (M3D_ERR_TOO_FEW_POINTS when a cloud has < 3 points). */
int m3d_ransac_registration(m3d_ctx *ctx, const double *src_xyz, size_t ns, const double *dst_xyz,
                            size_t nd, const size_t *c0, const size_t *c1, size_t m,
                            double threshold, int max_iter, double edge_thr, double confidence,
                            uint32_t seed, double *T_out, m3d_reg_stats *stats);

/* ---------------------------------------------- the Open3D steps around the path (f3, f4) */
/* open3d::pipelines::registration::ComputeFPFHFeature(cloud, KDTreeSearchParamHybrid(radius, max_nn)) -- what the
 * reference's callers run on the host to get the descriptors match_correspondence consumes
 * (examples/cpp/transform_estimation.cpp:20-33, examples/python/transform_estimation.py:12-27).
 *   xyz, nrm: n x 3 float64 host (normals are required: M3D_ERR_NO_NORMALS); feat_out: 33 x n float64 column-major
 *   (Feature::data_), directly usable as m3d_match_correspondence input.  1 <= max_nn <= 128.
 * Neighbour sets are exact (uniform grid in HBM, ascending (squared distance, index)); arithmetic fp64. */
int m3d_compute_fpfh(m3d_ctx *ctx, const double *xyz, const double *nrm, size_t n, double radius, int max_nn,
                     double *feat_out, float *device_ms);
/* open3d::pipelines::registration::RegistrationICP(src, dst, max_distance, init,
 * TransformationEstimationPointToPoint(false), ICPConvergenceCriteria(relative_fitness, relative_rmse, max_iteration))
 * -- the refinement the reference's callers run after compute_transformation_ransac
 * (examples/cpp/transform_estimation.cpp:82-86, src/pipeline.cpp:800-812).  T_init: 16 doubles row-major or NULL
 * (identity); T_out row-major; fitness = correspondences / ns, inlier_rmse = sqrt(sum d^2 / correspondences). */
int m3d_icp_point_to_point(m3d_ctx *ctx, const double *src_xyz, size_t ns, const double *dst_xyz, size_t nd,
                           double max_distance, const double *T_init, int max_iteration, double relative_fitness,
                           double relative_rmse, double *T_out, double *fitness, double *inlier_rmse, int *iterations);

/* Device-resident descriptors (the FPFH -> match_correspondence chain without the 2 x 52.8 MB round trip
 * of examples/cpp/transform_estimation.cpp:20-33).  A handle owns dim x n float64 (column-major) in HBM.
 *   m3d_fpfh_create      = m3d_compute_fpfh with the result left on the device
 *   m3d_features_upload  = descriptors computed elsewhere (host, dim x n column-major)
 *   m3d_match_features   = m3d_match_correspondence (ANNMatcher::Match) on two handles of the same context */
typedef struct m3d_features m3d_features;
int m3d_fpfh_create(m3d_ctx *ctx, const double *xyz, const double *nrm, size_t n, double radius, int max_nn,
                    m3d_features **out, float *device_ms);
int m3d_features_upload(m3d_ctx *ctx, const double *host, int dim, size_t n, m3d_features **out);
int m3d_features_download(const m3d_features *f, double *out);
size_t m3d_features_count(const m3d_features *f);
int m3d_features_dim(const m3d_features *f);
void m3d_features_free(m3d_features *f);
int m3d_match_features(m3d_ctx *ctx, const m3d_features *a, const m3d_features *b, size_t *idx0, size_t *idx1,
                       size_t *n_out, float *device_ms);

/* Replaces LeastSquareSolver::Solve = Eigen::umeyama over all pairs
 * (src/transform_estimation.cpp:49-66).  src/dst: n x 3 float64 host.  n < 3 -> M3D_ERR_TOO_FEW_POINTS
 * (the reference throws "The number of points pair is less than 3.", transform_estimation.cpp:29-31). */
int m3d_least_squares_transform(m3d_ctx *ctx, const double *src_xyz, const double *dst_xyz,
                                size_t n, int with_scaling, double *T_out);

/* Extension (the reference has no such step): LeastSquareSolver applied behind RANSACSolver --
 * Eigen::umeyama (src/transform_estimation.cpp:49-66) over the correspondences (c0[i], c1[i]) that are inliers
 * of T_in, |T_in s - d|^2 < threshold^2, taken in correspondence order.  T_in / T_out 4x4 row-major;
 * *n_inliers (may be NULL) = pairs used.  Fewer than 3 inliers: T_out = T_in. */
int m3d_registration_refit(m3d_ctx *ctx, const double *src_xyz, size_t ns, const double *dst_xyz, size_t nd,
                           const size_t *c0, const size_t *c1, size_t m, const double *T_in, double threshold,
                           int with_scaling, double *T_out, size_t *n_inliers);

#ifdef __cplusplus
}
#endif
#endif /* M3D_CAPI_H_ */
