/*
 * grid3d.cuh -- uniform-grid neighbour search over a 3-D point set resident in HBM: the spatial index behind the
 * FPFH features (Open3D KDTreeFlann::SearchHybrid as ComputeFPFHFeature uses it) and the ICP refinement
 * (SearchHybrid(point, max_distance, 1)).  The reference path has no such code: it calls Open3D's nanoflann kd-tree
 * on the host (examples/cpp/transform_estimation.cpp:20-33, 82-86).  Results are defined by the metric, not by the
 * index: the k nearest items in ascending (squared distance, index) with squared distance < radius^2, distances
 * accumulated in nanoflann's order ((dx^2 + dy^2) + dz^2, no FMA).
 *
 *   build    grid_bbox / grid_count / scan_* / grid_scatter : counting sort of the points by cell
 *            (cell edge >= search radius, so a query looks at 27 cells)
 *   query    hybrid_knn_kernel: thread = query, bounded max-heap on (d2, index) in local memory, heap-sorted output
 */
#pragma once
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdint>

#include "context.h"

namespace m3d {

namespace g3 { /* small helpers private to this index (ransac_kernels.cuh / score_cull.cuh have their own copies; those
                * headers define kernels and belong to ransac.cu alone) */
__device__ __forceinline__ double warp_min(double v) {
    for (int o = 16; o; o >>= 1) v = fmin(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ double warp_max(double v) {
    for (int o = 16; o; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
constexpr int kScanBlock = 1024, kScanItems = 2;
__device__ __forceinline__ uint32_t block_exclusive_scan_1024(uint32_t v, uint32_t *total) {
    __shared__ uint32_t wsum[32];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    uint32_t inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    if (lane == 31) wsum[w] = inc;
    __syncthreads();
    if (w == 0) {
        uint32_t s = wsum[lane];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t t = __shfl_up_sync(0xffffffffu, s, o);
            if (lane >= o) s += t;
        }
        wsum[lane] = s;
    }
    __syncthreads();
    const uint32_t before = w ? wsum[w - 1] : 0u;
    if (total) *total = wsum[31];
    return before + inc - v;
}
/* exclusive scan of `blocks` x 2048 counters in three kernels (sums, top, apply) */
__global__ void __launch_bounds__(kScanBlock) scan_sums_kernel(const uint32_t *__restrict__ hist, uint32_t *__restrict__ bsum) {
    const uint2 v = reinterpret_cast<const uint2 *>(hist)[blockIdx.x * kScanBlock + threadIdx.x];
    uint32_t tot;
    block_exclusive_scan_1024(v.x + v.y, &tot);
    if (threadIdx.x == 0) bsum[blockIdx.x] = tot;
}
__global__ void __launch_bounds__(kScanBlock) scan_top_kernel(uint32_t *bsum, int nblocks) {
    const uint32_t v = (int)threadIdx.x < nblocks ? bsum[threadIdx.x] : 0u;
    const uint32_t ex = block_exclusive_scan_1024(v, nullptr);
    if ((int)threadIdx.x < nblocks) bsum[threadIdx.x] = ex;
}
__global__ void __launch_bounds__(kScanBlock) scan_apply_kernel(uint32_t *hist, const uint32_t *__restrict__ bsum) {
    uint2 *h2 = reinterpret_cast<uint2 *>(hist);
    const uint2 v = h2[blockIdx.x * kScanBlock + threadIdx.x];
    const uint32_t ex = block_exclusive_scan_1024(v.x + v.y, nullptr) + bsum[blockIdx.x];
    h2[blockIdx.x * kScanBlock + threadIdx.x] = make_uint2(ex, ex + v.x);
}
}  // namespace g3
using g3::kScanBlock;
using g3::kScanItems;

struct Grid3 {
    double org[3]; /* lower corner of the bounding box */
    double inv_h;  /* 1 / cell edge                    */
    int dim[3];
    uint32_t ncells;
    const uint32_t *cell_start; /* [ncells + 1] */
    const uint32_t *order;      /* point indices sorted by cell */
    __device__ __forceinline__ void cell_of(const double *p, int c[3]) const {
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            const double f = floor((p[a] - org[a]) * inv_h);
            int v = (f != f) ? 0 : (f < 0 ? 0 : (f > 2.0e9 ? 2000000000 : (int)f));
            c[a] = min(max(v, 0), dim[a] - 1);
        }
    }
    __device__ __forceinline__ uint32_t cell_index(int x, int y, int z) const {
        return ((uint32_t)z * (uint32_t)dim[1] + (uint32_t)y) * (uint32_t)dim[0] + (uint32_t)x;
    }
};

constexpr int kGridScanUnit = kScanBlock * kScanItems; /* cells are padded to whole scan blocks */
constexpr uint32_t kGridMaxCells = 1u << 21;
constexpr int kKnnCap = 128; /* largest max_nn the per-thread heap holds */

/* min / max of the finite coordinates: part[block] = {mn[3], mx[3]} */
__global__ void __launch_bounds__(256) grid_bbox_kernel(const double *__restrict__ xyz, uint32_t n, double *__restrict__ part) {
    double mn[3] = {INFINITY, INFINITY, INFINITY}, mx[3] = {-INFINITY, -INFINITY, -INFINITY};
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const double v = xyz[3 * (size_t)i + c];
            if (isfinite(v)) {
                mn[c] = fmin(mn[c], v);
                mx[c] = fmax(mx[c], v);
            }
        }
    __shared__ double sh[8][6];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        mn[c] = g3::warp_min(mn[c]);
        mx[c] = g3::warp_max(mx[c]);
    }
    if ((threadIdx.x & 31) == 0)
        for (int c = 0; c < 3; ++c) {
            sh[threadIdx.x >> 5][c] = mn[c];
            sh[threadIdx.x >> 5][3 + c] = mx[c];
        }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int c = 0; c < 3; ++c)
            for (int k = 1; k < 8; ++k) {
                sh[0][c] = fmin(sh[0][c], sh[k][c]);
                sh[0][3 + c] = fmax(sh[0][3 + c], sh[k][3 + c]);
            }
        for (int c = 0; c < 6; ++c) part[6 * blockIdx.x + c] = sh[0][c];
    }
}
/* 192 threads: warp c reduces component c (min / max are exact whatever the order) */
__global__ void __launch_bounds__(192) grid_bbox_final_kernel(const double *__restrict__ part, int nparts, double *__restrict__ out) {
    const int c = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double r = c < 3 ? INFINITY : -INFINITY;
    for (int k = lane; k < nparts; k += 32) r = c < 3 ? fmin(r, part[6 * k + c]) : fmax(r, part[6 * k + c]);
    for (int o = 16; o; o >>= 1) {
        const double v = __shfl_xor_sync(0xffffffffu, r, o);
        r = c < 3 ? fmin(r, v) : fmax(r, v);
    }
    if (lane == 0) out[c] = r;
}
__global__ void __launch_bounds__(256) grid_count_kernel(const double *__restrict__ xyz, uint32_t n, Grid3 G,
                                                         uint32_t *__restrict__ cell_id, uint32_t *__restrict__ counts) {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const double p[3] = {xyz[3 * (size_t)i], xyz[3 * (size_t)i + 1], xyz[3 * (size_t)i + 2]};
        int c[3];
        G.cell_of(p, c);
        const uint32_t id = G.cell_index(c[0], c[1], c[2]);
        cell_id[i] = id;
        atomicAdd(&counts[id], 1u);
    }
}
__global__ void __launch_bounds__(256) grid_scatter_kernel(uint32_t n, const uint32_t *__restrict__ cell_id,
                                                           uint32_t *__restrict__ cursor, uint32_t *__restrict__ order) {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
        order[atomicAdd(&cursor[cell_id[i]], 1u)] = i;
}

/* device buffers of one grid (grow-only, owned by the caller) */
struct GridBufs {
    DevBuf cell_id, start, cursor, order, part;
};

/* builds the grid of `n` device points for searches of radius `radius`; *G is ready for kernels on ctx->stream */
inline int grid_build(m3d_ctx *ctx, const double *d_xyz, uint32_t n, double radius, GridBufs &B, Grid3 *G) {
    const int nb = std::max(1, std::min<int>(ctx->sm_count * 8, (int)((n + 255) / 256)));
    M3D_CUDA(ctx, B.part.reserve(sizeof(double) * (6 * (size_t)nb + 8)));
    double *part = B.part.as<double>(), *d_box = part + 6 * (size_t)nb;
    grid_bbox_kernel<<<nb, 256, 0, ctx->stream>>>(d_xyz, n, part);
    M3D_LAUNCHED(ctx);
    grid_bbox_final_kernel<<<1, 192, 0, ctx->stream>>>(part, nb, d_box);
    M3D_LAUNCHED(ctx);
    double box[6];
    M3D_CUDA(ctx, cudaMemcpyAsync(box, d_box, sizeof box, cudaMemcpyDeviceToHost, ctx->stream));
    M3D_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    /* cell edge a hair above the radius: two points closer than the radius then never sit two cells apart, whatever
     * the rounding of (p - org) / h at a cell boundary */
    double h = radius > 0 && std::isfinite(radius) ? radius * (1.0 + 1e-9) : 1.0;
    for (int c = 0; c < 3; ++c) {
        if (!(box[c] <= box[3 + c])) box[c] = box[3 + c] = 0; /* no finite coordinate */
        G->org[c] = box[c];
    }
    for (;;) { /* cell edge >= radius, at most kGridMaxCells cells */
        double cells = 1;
        for (int c = 0; c < 3; ++c) {
            const double e = std::floor((box[3 + c] - box[c]) / h) + 1.0;
            G->dim[c] = (int)std::min(e, 2.0e6);
            cells *= (double)G->dim[c];
        }
        if (cells <= (double)kGridMaxCells) break;
        h *= std::cbrt(cells / (double)kGridMaxCells) * 1.02;
    }
    G->inv_h = 1.0 / h;
    G->ncells = (uint32_t)G->dim[0] * (uint32_t)G->dim[1] * (uint32_t)G->dim[2];
    const uint32_t padded = (G->ncells + 1 + kGridScanUnit - 1) / kGridScanUnit * kGridScanUnit;
    const int scan_blocks = (int)(padded / kGridScanUnit);
    M3D_CUDA(ctx, B.cell_id.reserve(sizeof(uint32_t) * (size_t)std::max<uint32_t>(n, 1)));
    M3D_CUDA(ctx, B.order.reserve(sizeof(uint32_t) * (size_t)std::max<uint32_t>(n, 1)));
    M3D_CUDA(ctx, B.start.reserve(sizeof(uint32_t) * (size_t)padded));
    M3D_CUDA(ctx, B.cursor.reserve(sizeof(uint32_t) * ((size_t)padded + scan_blocks)));
    uint32_t *cursor = B.cursor.as<uint32_t>(), *bsum = cursor + padded;
    M3D_CUDA(ctx, cudaMemsetAsync(cursor, 0, sizeof(uint32_t) * (size_t)padded, ctx->stream));
    Grid3 g = *G;
    g.cell_start = nullptr;
    g.order = nullptr;
    grid_count_kernel<<<nb, 256, 0, ctx->stream>>>(d_xyz, n, g, B.cell_id.as<uint32_t>(), cursor);
    M3D_LAUNCHED(ctx);
    g3::scan_sums_kernel<<<scan_blocks, kScanBlock, 0, ctx->stream>>>(cursor, bsum);
    M3D_LAUNCHED(ctx);
    g3::scan_top_kernel<<<1, kScanBlock, 0, ctx->stream>>>(bsum, scan_blocks);
    M3D_LAUNCHED(ctx);
    g3::scan_apply_kernel<<<scan_blocks, kScanBlock, 0, ctx->stream>>>(cursor, bsum);
    M3D_LAUNCHED(ctx);
    M3D_CUDA(ctx, cudaMemcpyAsync(B.start.p, cursor, sizeof(uint32_t) * (size_t)padded, cudaMemcpyDeviceToDevice, ctx->stream));
    grid_scatter_kernel<<<nb, 256, 0, ctx->stream>>>(n, B.cell_id.as<uint32_t>(), cursor, B.order.as<uint32_t>());
    M3D_LAUNCHED(ctx);
    G->cell_start = B.start.as<uint32_t>();
    G->order = B.order.as<uint32_t>();
    return M3D_OK;
}

/* squared distance in nanoflann's accumulation order */
__device__ __forceinline__ double dist2_ref(const double *a, const double *b) {
    const double dx = __dsub_rn(a[0], b[0]), dy = __dsub_rn(a[1], b[1]), dz = __dsub_rn(a[2], b[2]);
    return __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
}
__device__ __forceinline__ bool knn_less(double da, uint32_t ia, double db, uint32_t ib) {
    return da < db || (da == db && ia < ib);
}

/* KDTreeFlann::SearchHybrid(query, radius, max_nn) for every point of the set against the set itself:
 * the K nearest items with d2 < r2 in ascending (d2, index).  thread = query (in cell order, so that a warp walks
 * the same cells).  out: nbr_idx / nbr_d2 [n][K], nbr_cnt [n]. */
__global__ void __launch_bounds__(128) hybrid_knn_kernel(const double *__restrict__ xyz, uint32_t n, Grid3 G, double r2,
                                                         int K, uint32_t *__restrict__ nbr_idx,
                                                         double *__restrict__ nbr_d2, uint32_t *__restrict__ nbr_cnt) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    const uint32_t i = G.order[t];
    const double q[3] = {xyz[3 * (size_t)i], xyz[3 * (size_t)i + 1], xyz[3 * (size_t)i + 2]};
    double hd[kKnnCap];
    uint32_t hi[kKnnCap];
    int size = 0;
    int c[3];
    G.cell_of(q, c);
    for (int z = max(c[2] - 1, 0); z <= min(c[2] + 1, G.dim[2] - 1); ++z)
        for (int y = max(c[1] - 1, 0); y <= min(c[1] + 1, G.dim[1] - 1); ++y)
            for (int x = max(c[0] - 1, 0); x <= min(c[0] + 1, G.dim[0] - 1); ++x) {
                const uint32_t cell = G.cell_index(x, y, z);
                const uint32_t e = G.cell_start[cell + 1];
                for (uint32_t p = G.cell_start[cell]; p < e; ++p) {
                    const uint32_t j = G.order[p];
                    const double pj[3] = {xyz[3 * (size_t)j], xyz[3 * (size_t)j + 1], xyz[3 * (size_t)j + 2]};
                    const double d2 = dist2_ref(q, pj);
                    if (!(d2 < r2)) continue;
                    if (size < K) { /* push, sift up (max-heap on (d2, index)) */
                        int k = size++;
                        while (k > 0) {
                            const int par = (k - 1) >> 1;
                            if (!knn_less(hd[par], hi[par], d2, j)) break;
                            hd[k] = hd[par];
                            hi[k] = hi[par];
                            k = par;
                        }
                        hd[k] = d2;
                        hi[k] = j;
                    } else if (knn_less(d2, j, hd[0], hi[0])) { /* replace the largest, sift down */
                        int k = 0;
                        for (;;) {
                            int ch = 2 * k + 1;
                            if (ch >= size) break;
                            if (ch + 1 < size && knn_less(hd[ch], hi[ch], hd[ch + 1], hi[ch + 1])) ++ch;
                            if (!knn_less(d2, j, hd[ch], hi[ch])) break;
                            hd[k] = hd[ch];
                            hi[k] = hi[ch];
                            k = ch;
                        }
                        hd[k] = d2;
                        hi[k] = j;
                    }
                }
            }
    /* heap sort: repeatedly move the largest to the end -> ascending */
    const int cnt = size;
    for (int end = size - 1; end > 0; --end) {
        const double d2 = hd[end];
        const uint32_t j = hi[end];
        hd[end] = hd[0];
        hi[end] = hi[0];
        int k = 0;
        for (;;) {
            int ch = 2 * k + 1;
            if (ch >= end) break;
            if (ch + 1 < end && knn_less(hd[ch], hi[ch], hd[ch + 1], hi[ch + 1])) ++ch;
            if (!knn_less(d2, j, hd[ch], hi[ch])) break;
            hd[k] = hd[ch];
            hi[k] = hi[ch];
            k = ch;
        }
        hd[k] = d2;
        hi[k] = j;
    }
    nbr_cnt[i] = (uint32_t)cnt;
    for (int k = 0; k < cnt; ++k) {
        nbr_idx[(size_t)i * K + k] = hi[k];
        nbr_d2[(size_t)i * K + k] = hd[k];
    }
}

}  // namespace m3d
