/*
 * features.cu -- the two Open3D steps that sit on either side of the registration hot path in every real use of it
 * (reference callers: examples/cpp/transform_estimation.cpp:20-33 and 82-86,
 * examples/python/transform_estimation.py:12-27, src/pipeline.cpp:800-812):
 *
 *   m3d_compute_fpfh        open3d::pipelines::registration::ComputeFPFHFeature(cloud, KDTreeSearchParamHybrid(radius,
 *                           max_nn))  -- the 33-D descriptors match_correspondence consumes
 *   m3d_icp_point_to_point  open3d::pipelines::registration::RegistrationICP(src, dst, max_distance, init,
 *                           TransformationEstimationPointToPoint(false), ICPConvergenceCriteria(...)) -- the
 *                           refinement after compute_transformation_ransac
 *
 * Both are Open3D code, not Misc3D code: the arithmetic restated here (and, independently, in the CPU checker of the
 * test suite) is Open3D v0.15.1's as recalled in SURVEY Appendix B style -- UNPINNED, like the registration RANSAC,
 * until tools/pin_open3d.py can run against a real Open3D.  Neighbour sets are exact (grid3d.cuh);
 * all feature / transform arithmetic is fp64.
 */
#include <algorithm>
#include <cmath>

#include "context.h"
#include "grid3d.cuh"
#include "umeyama.cuh"

namespace m3d {

/* ComputePairFeatures (Open3D Feature.cpp): the Darboux-frame angles of (p1, n1) and (p2, n2) */
__device__ inline void pair_features(const double *p1, const double *n1, const double *p2, const double *n2, double f[4]) {
    double dp[3] = {p2[0] - p1[0], p2[1] - p1[1], p2[2] - p1[2]};
    f[3] = sqrt(dp[0] * dp[0] + dp[1] * dp[1] + dp[2] * dp[2]);
    f[0] = f[1] = f[2] = 0;
    if (f[3] == 0.0) {
        f[3] = 0;
        return;
    }
    double a[3] = {n1[0], n1[1], n1[2]}, b[3] = {n2[0], n2[1], n2[2]};
    const double angle1 = (a[0] * dp[0] + a[1] * dp[1] + a[2] * dp[2]) / f[3];
    const double angle2 = (b[0] * dp[0] + b[1] * dp[1] + b[2] * dp[2]) / f[3];
    if (acos(fabs(angle1)) > acos(fabs(angle2))) {
        for (int c = 0; c < 3; ++c) {
            a[c] = n2[c];
            b[c] = n1[c];
            dp[c] *= -1.0;
        }
        f[2] = -angle2;
    } else {
        f[2] = angle1;
    }
    double v[3] = {dp[1] * a[2] - dp[2] * a[1], dp[2] * a[0] - dp[0] * a[2], dp[0] * a[1] - dp[1] * a[0]};
    const double vn = sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
    if (vn == 0.0) {
        f[0] = f[1] = f[2] = f[3] = 0;
        return;
    }
    for (int c = 0; c < 3; ++c) v[c] /= vn;
    const double w[3] = {a[1] * v[2] - a[2] * v[1], a[2] * v[0] - a[0] * v[2], a[0] * v[1] - a[1] * v[0]};
    f[1] = v[0] * b[0] + v[1] * b[1] + v[2] * b[2];
    f[0] = atan2(w[0] * b[0] + w[1] * b[1] + w[2] * b[2], a[0] * b[0] + a[1] * b[1] + a[2] * b[2]);
}

/* ComputeSPFHFeature: thread = point; spfh is 33 x n column-major (Feature::data_) */
__global__ void __launch_bounds__(128) spfh_kernel(const double *__restrict__ xyz, const double *__restrict__ nrm, uint32_t n,
                                                   int K, const uint32_t *__restrict__ nbr_idx,
                                                   const uint32_t *__restrict__ nbr_cnt, double *__restrict__ spfh) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double h[33];
#pragma unroll
    for (int j = 0; j < 33; ++j) h[j] = 0;
    const uint32_t cnt = nbr_cnt[i];
    if (cnt > 1) {
        const double incr = 100.0 / (double)(cnt - 1);
        const double p[3] = {xyz[3 * (size_t)i], xyz[3 * (size_t)i + 1], xyz[3 * (size_t)i + 2]};
        const double nn[3] = {nrm[3 * (size_t)i], nrm[3 * (size_t)i + 1], nrm[3 * (size_t)i + 2]};
        unsigned char bins[3];
        for (uint32_t k = 1; k < cnt; ++k) { /* the first entry is the point itself */
            const uint32_t j = nbr_idx[(size_t)i * K + k];
            const double q[3] = {xyz[3 * (size_t)j], xyz[3 * (size_t)j + 1], xyz[3 * (size_t)j + 2]};
            const double qn[3] = {nrm[3 * (size_t)j], nrm[3 * (size_t)j + 1], nrm[3 * (size_t)j + 2]};
            double f[4];
            pair_features(p, nn, q, qn, f);
            int b0 = (int)floor(11 * (f[0] + M_PI) / (2.0 * M_PI));
            int b1 = (int)floor(11 * (f[1] + 1.0) * 0.5);
            int b2 = (int)floor(11 * (f[2] + 1.0) * 0.5);
            bins[0] = (unsigned char)min(max(b0, 0), 10);
            bins[1] = (unsigned char)min(max(b1, 0), 10);
            bins[2] = (unsigned char)min(max(b2, 0), 10);
            /* h[bin] += incr with a dynamic index would put h in local memory: select by comparison instead */
#pragma unroll
            for (int j2 = 0; j2 < 11; ++j2) {
                h[j2] += (bins[0] == j2) ? incr : 0.0;
                h[11 + j2] += (bins[1] == j2) ? incr : 0.0;
                h[22 + j2] += (bins[2] == j2) ? incr : 0.0;
            }
        }
    }
#pragma unroll
    for (int j = 0; j < 33; ++j) spfh[(size_t)i * 33 + j] = h[j];
}

/* ComputeFPFHFeature's second loop: weighted sum of the neighbours' SPFH, per-histogram normalisation, + own SPFH */
__global__ void __launch_bounds__(128) fpfh_kernel(uint32_t n, int K, const uint32_t *__restrict__ nbr_idx,
                                                   const double *__restrict__ nbr_d2, const uint32_t *__restrict__ nbr_cnt,
                                                   const double *__restrict__ spfh, double *__restrict__ out) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double f[33];
#pragma unroll
    for (int j = 0; j < 33; ++j) f[j] = 0;
    const uint32_t cnt = nbr_cnt[i];
    if (cnt > 1) {
        double sum[3] = {0, 0, 0};
        for (uint32_t k = 1; k < cnt; ++k) {
            const double dist = nbr_d2[(size_t)i * K + k];
            if (dist == 0.0) continue;
            const double *s = spfh + (size_t)nbr_idx[(size_t)i * K + k] * 33;
#pragma unroll
            for (int j = 0; j < 33; ++j) {
                const double val = s[j] / dist;
                sum[j / 11] += val;
                f[j] += val;
            }
        }
#pragma unroll
        for (int j = 0; j < 3; ++j)
            if (sum[j] != 0.0) sum[j] = 100.0 / sum[j];
#pragma unroll
        for (int j = 0; j < 33; ++j) {
            f[j] *= sum[j / 11];
            f[j] += spfh[(size_t)i * 33 + j];
        }
    }
#pragma unroll
    for (int j = 0; j < 33; ++j) out[(size_t)i * 33 + j] = f[j];
}

/* ---------------------------------------------------------------------------------------------- ICP */
struct IcpAcc { /* device-side state of one ICP run */
    double T[16];      /* accumulated transformation (row-major)                  */
    double update[16]; /* the last estimated update                               */
    double mean[6];
    unsigned long long n_corr;
    double err2;
    double fitness, rmse;
};

/* GetRegistrationResultAndCorrespondences: nearest target point within max_distance of every (already transformed)
 * source point.  corr[i] = target index or 0xffffffff; per-block partial sums for the Umeyama means. */
__global__ void __launch_bounds__(256) icp_match_kernel(const double *__restrict__ src, uint32_t ns,
                                                        const double *__restrict__ dst, Grid3 G, double r2,
                                                        uint32_t *__restrict__ corr, double *__restrict__ part /*[blocks][8]*/) {
    double acc[8] = {0, 0, 0, 0, 0, 0, 0, 0}; /* count, err2, sum p (3), sum q (3) */
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < ns; i += gridDim.x * blockDim.x) {
        const double p[3] = {src[3 * (size_t)i], src[3 * (size_t)i + 1], src[3 * (size_t)i + 2]};
        int c[3];
        G.cell_of(p, c);
        double bd = INFINITY;
        uint32_t bj = 0xffffffffu;
        for (int z = max(c[2] - 1, 0); z <= min(c[2] + 1, G.dim[2] - 1); ++z)
            for (int y = max(c[1] - 1, 0); y <= min(c[1] + 1, G.dim[1] - 1); ++y)
                for (int x = max(c[0] - 1, 0); x <= min(c[0] + 1, G.dim[0] - 1); ++x) {
                    const uint32_t cell = G.cell_index(x, y, z);
                    const uint32_t e = G.cell_start[cell + 1];
                    for (uint32_t k = G.cell_start[cell]; k < e; ++k) {
                        const uint32_t j = G.order[k];
                        const double q[3] = {dst[3 * (size_t)j], dst[3 * (size_t)j + 1], dst[3 * (size_t)j + 2]};
                        const double d2 = dist2_ref(p, q);
                        if (d2 < r2 && knn_less(d2, j, bd, bj)) {
                            bd = d2;
                            bj = j;
                        }
                    }
                }
        corr[i] = bj;
        if (bj != 0xffffffffu) {
            acc[0] += 1.0;
            acc[1] += bd;
            for (int a = 0; a < 3; ++a) {
                acc[2 + a] += p[a];
                acc[5 + a] += dst[3 * (size_t)bj + a];
            }
        }
    }
    __shared__ double sh[8][8];
    for (int k = 0; k < 8; ++k)
        for (int o = 16; o; o >>= 1) acc[k] += __shfl_xor_sync(0xffffffffu, acc[k], o);
    if ((threadIdx.x & 31) == 0)
        for (int k = 0; k < 8; ++k) sh[threadIdx.x >> 5][k] = acc[k];
    __syncthreads();
    if (threadIdx.x == 0)
        for (int k = 0; k < 8; ++k) {
            double r = 0;
            for (int w = 0; w < 8; ++w) r += sh[w][k];
            part[8 * blockIdx.x + k] = r;
        }
}
/* sum of the per-block partials of quantity q by warp q: lane l adds parts l, l + 32, ... in order, then a fixed
 * shuffle tree (deterministic); thread 0 finishes */
template <int Q>
__device__ __forceinline__ void reduce_parts(const double *__restrict__ part, int nparts, double *__restrict__ sh) {
    const int q = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (q < Q) {
        double a = 0;
        for (int k = lane; k < nparts; k += 32) a += part[Q * k + q];
        for (int o = 16; o; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
        if (lane == 0) sh[q] = a;
    }
    __syncthreads();
}
__global__ void __launch_bounds__(256) icp_mean_kernel(const double *__restrict__ part, int nparts, uint32_t ns, IcpAcc *st) {
    __shared__ double r[8];
    reduce_parts<8>(part, nparts, r);
    if (threadIdx.x != 0) return;
    const unsigned long long m = (unsigned long long)(r[0] + 0.5);
    st->n_corr = m;
    st->err2 = r[1];
    if (m == 0) { /* RegistrationResult default: fitness 0, rmse 0 */
        st->fitness = 0;
        st->rmse = 0;
        for (int q = 0; q < 6; ++q) st->mean[q] = 0;
    } else {
        st->fitness = (double)m / (double)ns;
        st->rmse = sqrt(r[1] / (double)m);
        for (int q = 0; q < 6; ++q) st->mean[q] = r[2 + q] / (double)m;
    }
}
/* covariance of the corresponding pairs about their means (Eigen::umeyama: sigma = 1/n * dst_c * src_c^T) */
__global__ void __launch_bounds__(256) icp_cov_kernel(const double *__restrict__ src, uint32_t ns, const double *__restrict__ dst,
                                                      const uint32_t *__restrict__ corr, const IcpAcc *__restrict__ st,
                                                      double *__restrict__ part /*[blocks][9]*/) {
    double acc[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    const double inv = st->n_corr ? 1.0 / (double)st->n_corr : 0.0;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < ns; i += gridDim.x * blockDim.x) {
        const uint32_t j = corr[i];
        if (j == 0xffffffffu) continue;
        double sd[3], dd[3];
        for (int a = 0; a < 3; ++a) {
            sd[a] = src[3 * (size_t)i + a] - st->mean[a];
            dd[a] = dst[3 * (size_t)j + a] - st->mean[3 + a];
        }
        for (int r = 0; r < 3; ++r)
            for (int c = 0; c < 3; ++c) acc[3 * r + c] += (inv * dd[r]) * sd[c];
    }
    __shared__ double sh[8][9];
    for (int k = 0; k < 9; ++k)
        for (int o = 16; o; o >>= 1) acc[k] += __shfl_xor_sync(0xffffffffu, acc[k], o);
    if ((threadIdx.x & 31) == 0)
        for (int k = 0; k < 9; ++k) sh[threadIdx.x >> 5][k] = acc[k];
    __syncthreads();
    if (threadIdx.x == 0)
        for (int k = 0; k < 9; ++k) {
            double r = 0;
            for (int w = 0; w < 8; ++w) r += sh[w][k];
            part[9 * blockIdx.x + k] = r;
        }
}
/* update = umeyama(correspondences); T <- update * T */
__global__ void __launch_bounds__(288) icp_update_kernel(const double *__restrict__ part, int nparts, IcpAcc *st) {
    __shared__ double sg[9];
    reduce_parts<9>(part, nparts, sg);
    if (threadIdx.x != 0) return;
    double sigma[3][3], sm[3], dm[3];
    for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) sigma[r][c] = sg[3 * r + c];
    for (int a = 0; a < 3; ++a) {
        sm[a] = st->mean[a];
        dm[a] = st->mean[3 + a];
    }
    double U[16];
    rg::umeyama_finish(sigma, sm, dm, false, 1.0, U);
    double Tn[16];
    for (int r = 0; r < 4; ++r)
        for (int c = 0; c < 4; ++c) {
            double a = 0;
            for (int k = 0; k < 4; ++k) a += U[4 * r + k] * st->T[4 * k + c];
            Tn[4 * r + c] = a;
        }
    for (int q = 0; q < 16; ++q) {
        st->update[q] = U[q];
        st->T[q] = Tn[q];
    }
}
/* pcd.Transform(T): p <- T * [p; 1] (Open3D divides by the homogeneous coordinate) */
__global__ void __launch_bounds__(256) icp_transform_kernel(double *__restrict__ pts, uint32_t n, const double *__restrict__ T) {
    double t[16];
    for (int q = 0; q < 16; ++q) t[q] = T[q];
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const double x = pts[3 * (size_t)i], y = pts[3 * (size_t)i + 1], z = pts[3 * (size_t)i + 2];
        const double w = ((t[12] * x + t[13] * y) + t[14] * z) + t[15];
        pts[3 * (size_t)i] = (((t[0] * x + t[1] * y) + t[2] * z) + t[3]) / w;
        pts[3 * (size_t)i + 1] = (((t[4] * x + t[5] * y) + t[6] * z) + t[7]) / w;
        pts[3 * (size_t)i + 2] = (((t[8] * x + t[9] * y) + t[10] * z) + t[11]) / w;
    }
}

struct FeatBufs {
    GridBufs grid;
    DevBuf xyz, nrm, nbr_idx, nbr_d2, nbr_cnt, spfh, out, corr, part, state;
};

}  // namespace m3d

using namespace m3d;

struct m3d_feat_scratch {
    FeatBufs b;
};

static FeatBufs *feat_bufs(m3d_ctx *ctx) {
    if (!ctx->feat) ctx->feat = new m3d_feat_scratch();
    return &ctx->feat->b;
}
extern "C" void m3d_feat_scratch_free(m3d_feat_scratch *f) {
    if (!f) return;
    FeatBufs &b = f->b;
    DevBuf *all[] = {&b.grid.cell_id, &b.grid.start, &b.grid.cursor, &b.grid.order, &b.grid.part, &b.xyz, &b.nrm,
                     &b.nbr_idx, &b.nbr_d2, &b.nbr_cnt, &b.spfh, &b.out, &b.corr, &b.part, &b.state};
    for (auto *d : all) d->release();
    delete f;
}

extern "C" {

/* FPFH of a host cloud into d_out (33 x n f64 on the device); events ev[0] (start) .. the caller records the end */
static int fpfh_to_device(m3d_ctx *ctx, const double *xyz, const double *nrm, size_t n, double radius, int max_nn, double *d_out);

int m3d_compute_fpfh(m3d_ctx *ctx, const double *xyz, const double *nrm, size_t n, double radius, int max_nn,
                     double *feat_out, float *device_ms) {
    if (!ctx || (n && (!xyz || !feat_out))) return M3D_ERR_INVALID_ARG;
    if (device_ms) *device_ms = 0;
    if (int rc = fpfh_to_device(ctx, xyz, nrm, n, radius, max_nn, nullptr)) return rc;
    if (n == 0) return M3D_OK;
    FeatBufs &B = *feat_bufs(ctx);
    M3D_CUDA(ctx, cudaMemcpyAsync(feat_out, B.out.p, sizeof(double) * 33 * n, cudaMemcpyDeviceToHost, ctx->stream));
    M3D_CUDA(ctx, cudaEventRecord(ctx->ev[1], ctx->stream));
    M3D_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (device_ms) cudaEventElapsedTime(device_ms, ctx->ev[0], ctx->ev[1]);
    return M3D_OK;
}

/* ---- device-resident descriptors ("kept on device so descriptors never round-trip") */
int m3d_fpfh_create(m3d_ctx *ctx, const double *xyz, const double *nrm, size_t n, double radius, int max_nn,
                    m3d_features **out, float *device_ms) {
    if (!ctx || !out || (n && !xyz)) return M3D_ERR_INVALID_ARG;
    *out = nullptr;
    if (device_ms) *device_ms = 0;
    M3D_CUDA(ctx, cudaSetDevice(ctx->device));
    m3d_features *f = new m3d_features();
    f->ctx = ctx;
    f->dim = 33;
    f->n = n;
    if (f->data.reserve(sizeof(double) * 33 * std::max<size_t>(n, 1)) != cudaSuccess) {
        delete f;
        cudaGetLastError();
        return ctx->fail(M3D_ERR_CUDA, "cudaMalloc of %zu descriptors failed", n);
    }
    if (int rc = fpfh_to_device(ctx, xyz, nrm, n, radius, max_nn, f->data.as<double>())) {
        f->data.release();
        delete f;
        return rc;
    }
    if (n) {
        M3D_CUDA(ctx, cudaEventRecord(ctx->ev[1], ctx->stream));
        M3D_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        if (device_ms) cudaEventElapsedTime(device_ms, ctx->ev[0], ctx->ev[1]);
    }
    *out = f;
    return M3D_OK;
}
int m3d_features_upload(m3d_ctx *ctx, const double *host, int dim, size_t n, m3d_features **out) {
    if (!ctx || !out || dim <= 0 || (n && !host)) return M3D_ERR_INVALID_ARG;
    *out = nullptr;
    M3D_CUDA(ctx, cudaSetDevice(ctx->device));
    m3d_features *f = new m3d_features();
    f->ctx = ctx;
    f->dim = dim;
    f->n = n;
    const size_t bytes = sizeof(double) * (size_t)dim * n;
    if (f->data.reserve(std::max<size_t>(bytes, 8)) != cudaSuccess) {
        delete f;
        cudaGetLastError();
        return ctx->fail(M3D_ERR_CUDA, "cudaMalloc of %zu descriptors failed", n);
    }
    if (n) {
        int rc = host_to_device(ctx, f->data.p, host, bytes, ctx->stream);
        if (rc == M3D_OK && cudaStreamSynchronize(ctx->stream) != cudaSuccess) rc = ctx->fail(M3D_ERR_CUDA, "descriptor upload failed");
        if (rc) {
            f->data.release();
            delete f;
            return rc;
        }
    }
    *out = f;
    return M3D_OK;
}
int m3d_features_download(const m3d_features *f, double *out) {
    if (!f || !f->ctx || (f->n && !out)) return M3D_ERR_INVALID_ARG;
    m3d_ctx *ctx = f->ctx;
    if (f->n == 0) return M3D_OK;
    M3D_CUDA(ctx, cudaSetDevice(ctx->device));
    M3D_CUDA(ctx, cudaMemcpyAsync(out, f->data.p, sizeof(double) * (size_t)f->dim * f->n, cudaMemcpyDeviceToHost, ctx->stream));
    M3D_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return M3D_OK;
}
size_t m3d_features_count(const m3d_features *f) { return f ? f->n : 0; }
int m3d_features_dim(const m3d_features *f) { return f ? f->dim : 0; }
void m3d_features_free(m3d_features *f) {
    if (!f) return;
    if (f->ctx) cudaStreamSynchronize(f->ctx->stream);
    f->data.release();
    delete f;
}

static int fpfh_to_device(m3d_ctx *ctx, const double *xyz, const double *nrm, size_t n, double radius, int max_nn, double *d_out) {
    if (n && !nrm) return ctx->fail(M3D_ERR_NO_NORMALS, "Failed because input point cloud has no normal."); /* Feature.cpp */
    if (!(radius > 0) || max_nn < 1 || max_nn > kKnnCap)
        return ctx->fail(M3D_ERR_INVALID_ARG, "FPFH needs radius > 0 and 1 <= max_nn <= %d", kKnnCap);
    if (n >= (1ull << 31)) return ctx->fail(M3D_ERR_INVALID_ARG, "clouds of >= 2^31 points are not supported");
    if (n == 0) return M3D_OK;
    M3D_CUDA(ctx, cudaSetDevice(ctx->device));
    FeatBufs &B = *feat_bufs(ctx);
    const uint32_t N = (uint32_t)n;
    const int K = max_nn;
    M3D_CUDA(ctx, B.xyz.reserve(sizeof(double) * 3 * n));
    M3D_CUDA(ctx, B.nrm.reserve(sizeof(double) * 3 * n));
    M3D_CUDA(ctx, B.nbr_idx.reserve(sizeof(uint32_t) * n * K));
    M3D_CUDA(ctx, B.nbr_d2.reserve(sizeof(double) * n * K));
    M3D_CUDA(ctx, B.nbr_cnt.reserve(sizeof(uint32_t) * n));
    M3D_CUDA(ctx, B.spfh.reserve(sizeof(double) * 33 * n));
    if (!d_out) M3D_CUDA(ctx, B.out.reserve(sizeof(double) * 33 * n));
    M3D_CUDA(ctx, cudaEventRecord(ctx->ev[0], ctx->stream));
    if (int rc = host_to_device(ctx, B.xyz.p, xyz, sizeof(double) * 3 * n, ctx->stream)) return rc;
    M3D_CUDA(ctx, cudaStreamSynchronize(ctx->stream)); /* one staging buffer for pageable sources */
    if (int rc = host_to_device(ctx, B.nrm.p, nrm, sizeof(double) * 3 * n, ctx->stream)) return rc;
    Grid3 G;
    if (int rc = grid_build(ctx, B.xyz.as<double>(), N, radius, B.grid, &G)) return rc;
    const int nb = (int)((n + 127) / 128);
    hybrid_knn_kernel<<<nb, 128, 0, ctx->stream>>>(B.xyz.as<double>(), N, G, radius * radius, K, B.nbr_idx.as<uint32_t>(),
                                                   B.nbr_d2.as<double>(), B.nbr_cnt.as<uint32_t>());
    M3D_LAUNCHED(ctx);
    spfh_kernel<<<nb, 128, 0, ctx->stream>>>(B.xyz.as<double>(), B.nrm.as<double>(), N, K, B.nbr_idx.as<uint32_t>(),
                                             B.nbr_cnt.as<uint32_t>(), B.spfh.as<double>());
    M3D_LAUNCHED(ctx);
    fpfh_kernel<<<nb, 128, 0, ctx->stream>>>(N, K, B.nbr_idx.as<uint32_t>(), B.nbr_d2.as<double>(), B.nbr_cnt.as<uint32_t>(),
                                             B.spfh.as<double>(), d_out ? d_out : B.out.as<double>());
    M3D_LAUNCHED(ctx);
    return M3D_OK;
}

int m3d_icp_point_to_point(m3d_ctx *ctx, const double *src_xyz, size_t ns, const double *dst_xyz, size_t nd,
                           double max_distance, const double *T_init, int max_iteration, double relative_fitness,
                           double relative_rmse, double *T_out, double *fitness, double *inlier_rmse, int *iterations) {
    if (!ctx || !T_out || (ns && !src_xyz) || (nd && !dst_xyz)) return M3D_ERR_INVALID_ARG;
    static const double I4[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};
    for (int q = 0; q < 16; ++q) T_out[q] = T_init ? T_init[q] : I4[q];
    if (fitness) *fitness = 0;
    if (inlier_rmse) *inlier_rmse = 0;
    if (iterations) *iterations = 0;
    if (!(max_distance > 0)) return ctx->fail(M3D_ERR_INVALID_ARG, "Invalid max_correspondence_distance."); /* Registration.cpp */
    if (ns >= (1ull << 31) || nd >= (1ull << 31)) return ctx->fail(M3D_ERR_INVALID_ARG, "clouds of >= 2^31 points are not supported");
    if (ns == 0 || nd == 0) return M3D_OK;
    M3D_CUDA(ctx, cudaSetDevice(ctx->device));
    FeatBufs &B = *feat_bufs(ctx);
    const uint32_t NS = (uint32_t)ns, ND = (uint32_t)nd;
    const int nb = std::max(1, std::min<int>(ctx->sm_count * 4, (int)((ns + 255) / 256)));
    M3D_CUDA(ctx, B.xyz.reserve(sizeof(double) * 3 * nd)); /* target */
    M3D_CUDA(ctx, B.nrm.reserve(sizeof(double) * 3 * ns)); /* source, transformed in place */
    M3D_CUDA(ctx, B.corr.reserve(sizeof(uint32_t) * ns));
    M3D_CUDA(ctx, B.part.reserve(sizeof(double) * 9 * (size_t)nb));
    M3D_CUDA(ctx, B.state.reserve(sizeof(IcpAcc)));
    M3D_CUDA(ctx, ctx->h_small.reserve(sizeof(IcpAcc) + 4096));
    IcpAcc *st = B.state.as<IcpAcc>();
    IcpAcc *hs = ctx->h_small.as<IcpAcc>();
    double *d_src = B.nrm.as<double>();
    const double *d_dst = B.xyz.as<double>();
    if (int rc = host_to_device(ctx, B.xyz.p, dst_xyz, sizeof(double) * 3 * nd, ctx->stream)) return rc;
    M3D_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (int rc = host_to_device(ctx, B.nrm.p, src_xyz, sizeof(double) * 3 * ns, ctx->stream)) return rc;
    IcpAcc init{};
    for (int q = 0; q < 16; ++q) init.T[q] = T_out[q], init.update[q] = I4[q];
    M3D_CUDA(ctx, cudaMemcpyAsync(st, &init, sizeof init, cudaMemcpyHostToDevice, ctx->stream));
    Grid3 G;
    if (int rc = grid_build(ctx, d_dst, ND, max_distance, B.grid, &G)) return rc;
    /* pcd = source; if (!init.isIdentity()) pcd.Transform(init) */
    bool ident = true;
    for (int q = 0; q < 16; ++q) ident = ident && (T_out[q] == I4[q]);
    if (!ident) {
        icp_transform_kernel<<<nb, 256, 0, ctx->stream>>>(d_src, NS, st->T);
        M3D_LAUNCHED(ctx);
    }
    auto evaluate = [&]() -> int { /* GetRegistrationResultAndCorrespondences */
        icp_match_kernel<<<nb, 256, 0, ctx->stream>>>(d_src, NS, d_dst, G, max_distance * max_distance, B.corr.as<uint32_t>(),
                                                      B.part.as<double>());
        M3D_LAUNCHED(ctx);
        icp_mean_kernel<<<1, 256, 0, ctx->stream>>>(B.part.as<double>(), nb, NS, st);
        M3D_LAUNCHED(ctx);
        return M3D_OK;
    };
    if (int rc = evaluate()) return rc;
    M3D_CUDA(ctx, cudaMemcpyAsync(hs, st, sizeof(IcpAcc), cudaMemcpyDeviceToHost, ctx->stream));
    M3D_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    double fit = hs->fitness, rmse = hs->rmse;
    int it = 0;
    for (; it < max_iteration; ++it) {
        if (hs->n_corr == 0) break; /* nothing to estimate from (Open3D would return an identity update) */
        icp_cov_kernel<<<nb, 256, 0, ctx->stream>>>(d_src, NS, d_dst, B.corr.as<uint32_t>(), st, B.part.as<double>());
        M3D_LAUNCHED(ctx);
        icp_update_kernel<<<1, 288, 0, ctx->stream>>>(B.part.as<double>(), nb, st);
        M3D_LAUNCHED(ctx);
        icp_transform_kernel<<<nb, 256, 0, ctx->stream>>>(d_src, NS, st->update);
        M3D_LAUNCHED(ctx);
        if (int rc = evaluate()) return rc;
        M3D_CUDA(ctx, cudaMemcpyAsync(hs, st, sizeof(IcpAcc), cudaMemcpyDeviceToHost, ctx->stream));
        M3D_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        const double bf = fit, br = rmse;
        fit = hs->fitness, rmse = hs->rmse;
        if (std::fabs(bf - fit) < relative_fitness && std::fabs(br - rmse) < relative_rmse) {
            ++it;
            break;
        }
    }
    for (int q = 0; q < 16; ++q) T_out[q] = hs->T[q];
    if (fitness) *fitness = fit;
    if (inlier_rmse) *inlier_rmse = rmse;
    if (iterations) *iterations = it;
    return M3D_OK;
}

} /* extern "C" */
