/*
 * m3d_oracle.h -- C API of the CPU oracle.
 *
 * TEST INFRASTRUCTURE ONLY.  This is a dependency-free CPU restatement of the
 * Misc3D RANSAC / segmentation / matching / registration hot path.  Only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference legs may load it.  The product (libm3d_b200.so) never links,
 * includes or calls anything in this directory.
 *
 * PARITY STATUS.  The reference ships no tests / golden vectors for this path
 *  and its normal build needs Open3D + Eigen (absent, no network).
 *   PINNED against outputs of the reference itself run here: ransac.h
 *   (sampler, estimators, FitModel loop, RefineModel), iterative_plane_
 *   segmentation.cpp and correspondence_matching.cpp are compiled UNMODIFIED
 *   from /root/reference into oracle/_ref (Makefile target `_ref`, Eigen/Open3D
 *   stood in by oracle/shim/, seed injected) and tests/test_reference_pin.py
 *   requires this oracle to reproduce them bit for bit.  The stand-ins use the
 *   Eigen evaluation orders of SURVEY.md Appendix D, so what is pinned is the
 *   reference's control flow and formulas, not Eigen's instruction selection.
 *   UNPINNED: orc_ransac_registration / orc_umeyama -- that arithmetic is
 *   Open3D v0.15.1's RegistrationRANSACBasedOnCorrespondence + Eigen::umeyama,
 *   not under /root/reference; restated from the published algorithms
 *    and cross-checked with numpy only.
 */
#ifndef M3D_ORACLE_H_
#define M3D_ORACLE_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum { ORC_PLANE = 0, ORC_SPHERE = 1, ORC_CYLINDER = 2 };

typedef struct orc_stats {
    uint64_t best_index;     /* hypothesis index of the winning minimal model */
    uint64_t best_count;     /* its inlier count during EvaluateModel */
    double best_rmse;        /* its error / sqrt(count) (ransac.h:650) */
    uint64_t iterations_run; /* `count` printed by ransac.h:616-619 */
    uint64_t stop_index;     /* first loop index skipped by ransac.h:573 (== max_it if none) */
    int32_t found;           /* 1 if any hypothesis ever became best */
    int32_t refit_ok;        /* return value of GeneralFit (ransac.h:548) */
} orc_stats;

/* utils.h:81-97 : k distinct indices per row, draw order, rng()%n, mt19937(seed). */
void orc_sample_table(uint32_t seed, size_t n, int k, size_t rows, uint32_t *out);

/* ransac.h:138-162 / 239-294 / 354-417.  pts: k x 3 (already in ascending
 * index order, as SelectByIndex emits them), nrm: k x 3 or NULL. */
int orc_minimal_fit(int kind, const double *pts, const double *nrm, double *model);
/* ransac.h:215-220 / 332-343 / 435-445 */
double orc_distance(int kind, const double *model, const double *q);
/* ransac.h:626-654 : returns inlier count, *err = sum of distances in index order */
uint64_t orc_evaluate(int kind, const double *xyz, size_t n, const double *model, double thr,
                      double *err);
/* ransac.h:164-213 / 296-330 / 427-433 */
int orc_general_fit(int kind, const double *xyz, size_t n, double *model);

/* ransac.h:506-516 + 561-624 + 534-549, sequential, seeded (parity definition). */
int orc_ransac_fit(int kind, const double *xyz, const double *nrm, size_t n, double thr,
                   size_t max_it, double prob, uint32_t seed, double *model, size_t *inl,
                   size_t *n_inl, orc_stats *st);

/* Same loop with `#pragma omp parallel for schedule(static)` and a shared
 * mutex-guarded sampler exactly as the reference (timing baseline only; not
 * deterministic).  faithful!=0 also pays Open3D SelectByIndex's O(N) mask
 * pass per hypothesis (ransac.h:578). */
int orc_ransac_fit_omp(int kind, const double *xyz, const double *nrm, size_t n, double thr,
                       size_t max_it, double prob, uint32_t seed, int faithful, double *model,
                       size_t *inl, size_t *n_inl, orc_stats *st);

/* iterative_plane_segmentation.cpp:7-39.  labels[n] = plane id or UINT64_MAX.
 * returns 0 ok, -2 if a round is entered with < 3 remaining points (the
 * reference throws there), -3 if a round finds zero inliers (reference loops
 * forever). use_omp selects the timing variant. */
int orc_segment_plane_iterative(const double *xyz, size_t n, double thr, int max_it,
                                double min_ratio, uint32_t seed, int use_omp, double *planes,
                                size_t cap_planes, uint64_t *labels, size_t *n_planes);

/* correspondence_matching.cpp:13-84 with exact 1-NN (FLANN branch), ties ->
 * lowest index. src/dst: dim x count column-major. idx0/idx1 capacity ns. */
int orc_match_correspondence(const double *src, size_t ns, const double *dst, size_t nd, int dim,
                             size_t *idx0, size_t *idx1, size_t *n_out);
/* one direction only: nn[i] = argmin_j |src_i - dst_j|^2 */
void orc_nearest(const double *src, size_t ns, const double *dst, size_t nd, int dim, size_t *nn);

/* Eigen::umeyama(src, dst, with_scaling) on 3 x n column-major inputs; T row-major 4x4. */
void orc_umeyama(const double *src, const double *dst, size_t n, int with_scaling, double *T);

typedef struct orc_reg_stats {
    uint64_t best_index;
    uint64_t best_count;
    double best_rmse;
    uint64_t evaluated;  /* hypotheses that passed both checkers */
    uint64_t stop_index; /* first itr with itr >= est_k_global (== max_iter if none) */
} orc_reg_stats;

/* transform_estimation.cpp:124-164 -> Open3D RegistrationRANSACBasedOnCorrespondence
 * (sequential, seeded). T_out row-major 4x4. use_omp: timing variant. */
int orc_ransac_registration(const double *src_xyz, size_t ns, const double *dst_xyz, size_t nd,
                            const size_t *c0, const size_t *c1, size_t m, double thr, int max_iter,
                            double edge_thr, double confidence, uint32_t seed, int use_omp,
                            double *T_out, orc_reg_stats *st);

/* libstdc++ uniform_int_distribution<int>(0,m-1) over mt19937(seed): rows x 3 */
void orc_reg_sample_table(uint32_t seed, size_t m, size_t rows, uint32_t *out);

int orc_omp_threads(void);
void orc_set_threads(int n);

#ifdef __cplusplus
}
#endif
#endif
