/*
 * knn.cu -- exact k-nearest-neighbour index behind misc3d::common::KNearestSearch
 * (reference: include/misc3d/common/knn.h:24-73, src/knn.cpp:36-139 -- an Annoy forest there).
 *
 * The reference's class wraps an approximate Annoy index (random projection trees, racy multi-threaded
 * build).  The B200 build keeps the class and its call signatures but answers every query with
 * the EXACT neighbours: the data set lives in HBM (dim x n float64, column-major as Eigen::MatrixXd), a
 * query batch is compared with all columns in fp64 and the k smallest distances are selected per query
 * (ascending, ties to the lower index).  Distances are Euclidean (not squared), as Annoy reports them
 * (knn.cpp:110 passes Annoy's distances through).  The hot matching path does not go through here
 * (matching.cu's tensor-core search); this is the API class.
 */
#include <algorithm>
#include <cmath>

#include "context.h"

struct m3d_knn {
    m3d_ctx *ctx = nullptr;
    int dim = 0;
    size_t n = 0;
    m3d::DevBuf data;                  /* dim x n f64 */
    m3d::DevBuf q, dist, idx, outd, cnt; /* per-call scratch, grow-only */
};

namespace m3d {

/* dist[q][i] = |data_i - query_q|^2, terms added in dimension order */
__global__ void __launch_bounds__(256) knn_dist_kernel(const double *__restrict__ data, int dim, uint32_t n,
                                                       const double *__restrict__ queries, uint32_t nq,
                                                       double *__restrict__ dist) {
    extern __shared__ double sq[]; /* the query (+ 2 words of slack: the compiler reads it in 16-byte pairs) */
    const uint32_t q = blockIdx.y;
    for (int d = threadIdx.x; d < dim; d += blockDim.x) sq[d] = queries[(size_t)q * dim + d];
    __syncthreads();
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const double *p = data + (size_t)i * dim;
        double s = 0;
        for (int d = 0; d < dim; ++d) {
            const double t = __dsub_rn(p[d], sq[d]);
            s = __dadd_rn(s, __dmul_rn(t, t));
        }
        dist[(size_t)q * n + i] = s;
    }
}

/* one CTA per query: k rounds of (min distance, lowest index) over the row; a taken entry is set to +inf */
__global__ void __launch_bounds__(256) knn_select_kernel(double *__restrict__ dist, uint32_t n, int k, double radius,
                                                         unsigned long long *__restrict__ idx_out,
                                                         double *__restrict__ dist_out, int *__restrict__ count_out) {
    __shared__ double sv[8];
    __shared__ uint32_t si[8];
    __shared__ uint32_t s_best;
    const uint32_t q = blockIdx.x;
    double *row = dist + (size_t)q * n;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    int found = 0;
    const int kk = (int)min((uint32_t)k, n);
    for (int r = 0; r < kk; ++r) {
        double bv = INFINITY;
        uint32_t bi = 0xffffffffu;
        for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) {
            const double v = row[i];
            if (v < bv) { /* strided ascending i per thread: the first minimum wins inside a thread */
                bv = v;
                bi = i;
            }
        }
        for (int o = 16; o; o >>= 1) {
            const double ov = __shfl_xor_sync(0xffffffffu, bv, o);
            const uint32_t oi = __shfl_xor_sync(0xffffffffu, bi, o);
            if (ov < bv || (ov == bv && oi < bi)) {
                bv = ov;
                bi = oi;
            }
        }
        if (lane == 0) {
            sv[w] = bv;
            si[w] = bi;
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            for (int j = 1; j < 8; ++j)
                if (sv[j] < bv || (sv[j] == bv && si[j] < bi)) {
                    bv = sv[j];
                    bi = si[j];
                }
            s_best = bi;
            if (bi != 0xffffffffu) {
                const double d = sqrt(bv);
                if (!(radius > 0) || d <= radius) {
                    idx_out[(size_t)q * k + found] = bi;
                    dist_out[(size_t)q * k + found] = d;
                    ++found;
                } else {
                    s_best = 0xffffffffu; /* everything else is farther: stop */
                }
                if (s_best != 0xffffffffu) row[bi] = INFINITY;
            }
        }
        __syncthreads();
        if (s_best == 0xffffffffu) break;
    }
    if (threadIdx.x == 0) count_out[q] = found;
}

}  // namespace m3d

using namespace m3d;

extern "C" {

int m3d_knn_create(m3d_ctx *ctx, const double *data, int dim, size_t n, m3d_knn **out) {
    if (!ctx || !out || dim <= 0 || (n && !data)) return M3D_ERR_INVALID_ARG;
    if (n >= (1ull << 31)) return ctx->fail(M3D_ERR_INVALID_ARG, "more than 2^31 items");
    M3D_CUDA(ctx, cudaSetDevice(ctx->device));
    m3d_knn *k = new m3d_knn();
    k->ctx = ctx;
    k->dim = dim;
    k->n = n;
    const size_t bytes = sizeof(double) * (size_t)dim * std::max<size_t>(n, 1);
    if (k->data.reserve(bytes) != cudaSuccess) {
        delete k;
        return ctx->fail(M3D_ERR_CUDA, "cudaMalloc of the knn data set failed");
    }
    if (n) {
        cudaError_t e = cudaMemcpyAsync(k->data.p, data, sizeof(double) * (size_t)dim * n, cudaMemcpyHostToDevice, ctx->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
        if (e != cudaSuccess) {
            k->data.release();
            delete k;
            return ctx->fail(M3D_ERR_CUDA, "upload of the knn data set failed: %s", cudaGetErrorString(e));
        }
    }
    *out = k;
    return M3D_OK;
}

void m3d_knn_free(m3d_knn *k) {
    if (!k) return;
    if (k->ctx) cudaStreamSynchronize(k->ctx->stream);
    k->data.release();
    k->q.release();
    k->dist.release();
    k->idx.release();
    k->outd.release();
    k->cnt.release();
    delete k;
}

int m3d_knn_search(m3d_ctx *ctx, m3d_knn *index, const double *queries, size_t nq, int k, double radius,
                   size_t *idx_out, double *dist_out, int *count_out) {
    if (!ctx || !index || k < 0 || (nq && (!queries || !count_out)) || (nq && k && (!idx_out || !dist_out)))
        return M3D_ERR_INVALID_ARG;
    M3D_CUDA(ctx, cudaSetDevice(ctx->device));
    if (nq == 0) return M3D_OK;
    if (index->n == 0 || k == 0) {
        for (size_t i = 0; i < nq; ++i) count_out[i] = 0;
        return M3D_OK;
    }
    const uint32_t n = (uint32_t)index->n;
    const int dim = index->dim;
    /* queries are processed in chunks whose distance rows fit 256 MB */
    const size_t chunk = std::max<size_t>(1, std::min<size_t>(nq, ((size_t)256 << 20) / (sizeof(double) * n)));
    M3D_CUDA(ctx, index->q.reserve(sizeof(double) * chunk * dim));
    M3D_CUDA(ctx, index->dist.reserve(sizeof(double) * chunk * n));
    M3D_CUDA(ctx, index->idx.reserve(sizeof(unsigned long long) * chunk * k));
    M3D_CUDA(ctx, index->outd.reserve(sizeof(double) * chunk * k));
    M3D_CUDA(ctx, index->cnt.reserve(sizeof(int) * chunk));
    static_assert(sizeof(size_t) == sizeof(unsigned long long), "size_t must be 64-bit");
    for (size_t q0 = 0; q0 < nq; q0 += chunk) {
        const uint32_t cq = (uint32_t)std::min(chunk, nq - q0);
        M3D_CUDA(ctx, cudaMemcpyAsync(index->q.p, queries + q0 * dim, sizeof(double) * (size_t)cq * dim,
                                      cudaMemcpyHostToDevice, ctx->stream));
        const int bx = std::max(1, std::min<int>((n + 255) / 256, std::max(1, ctx->sm_count * 8 / (int)std::min<uint32_t>(cq, 64))));
        knn_dist_kernel<<<dim3(bx, cq), 256, sizeof(double) * ((size_t)dim + 2), ctx->stream>>>(index->data.as<double>(), dim, n,
                                                                                  index->q.as<double>(), cq,
                                                                                  index->dist.as<double>());
        M3D_LAUNCHED(ctx);
        knn_select_kernel<<<cq, 256, 0, ctx->stream>>>(index->dist.as<double>(), n, k, radius,
                                                       index->idx.as<unsigned long long>(), index->outd.as<double>(),
                                                       index->cnt.as<int>());
        M3D_LAUNCHED(ctx);
        M3D_CUDA(ctx, cudaMemcpyAsync(idx_out + q0 * k, index->idx.p, sizeof(size_t) * (size_t)cq * k, cudaMemcpyDeviceToHost, ctx->stream));
        M3D_CUDA(ctx, cudaMemcpyAsync(dist_out + q0 * k, index->outd.p, sizeof(double) * (size_t)cq * k, cudaMemcpyDeviceToHost, ctx->stream));
        M3D_CUDA(ctx, cudaMemcpyAsync(count_out + q0, index->cnt.p, sizeof(int) * cq, cudaMemcpyDeviceToHost, ctx->stream));
        M3D_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    }
    return M3D_OK;
}

} /* extern "C" */
