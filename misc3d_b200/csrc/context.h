/*
 * context.h -- internal: m3d_ctx / m3d_cloud definitions, device buffers, launch accounting.
 * Not part of the public ABI (include/m3d_capi.h).
 */
#pragma once
#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <utility>
#include <vector>

#include "../../include/m3d_capi.h"

namespace m3d {

/* grow-only device / pinned-host buffer */
template <bool HOST>
struct Buf {
    void *p = nullptr;
    size_t cap = 0;
    cudaError_t reserve(size_t bytes) {
        if (bytes <= cap) return cudaSuccess;
        release();
        size_t want = bytes + bytes / 4 + 256;
        cudaError_t e = HOST ? cudaMallocHost(&p, want) : cudaMalloc(&p, want);
        if (e != cudaSuccess) {
            p = nullptr;
            cap = 0;
            return e;
        }
        cap = want;
        return cudaSuccess;
    }
    void release() {
        if (p) {
            if (HOST)
                cudaFreeHost(p);
            else
                cudaFree(p);
        }
        p = nullptr;
        cap = 0;
    }
    template <class T>
    T *as() const {
        return reinterpret_cast<T *>(p);
    }
};
using DevBuf = Buf<false>;
using PinBuf = Buf<true>;

/* per-cloud constants the kernels read (device resident) */
struct CloudMeta {
    double center[3]; /* bounding-box centre; the fp32 copy of the cloud is stored relative to it */
    double mc;        /* max |centred coordinate| over the cloud                                   */
    double mraw;      /* max |raw coordinate|                                                      */
    int nonfinite;    /* 1 if any coordinate is NaN/inf -> fp64 reference-order kernels only       */
    int pad;
};

struct NcclApi; /* nccl_dl.cpp */

}  // namespace m3d

struct m3d_features { /* device-resident descriptors: dim x n float64, column-major (one column per point) */
    m3d_ctx *ctx = nullptr;
    int dim = 0;
    size_t n = 0;
    m3d::DevBuf data;
};

struct m3d_cloud {
    m3d_ctx *ctx = nullptr;
    size_t n = 0;
    bool has_normals = false;
    m3d::DevBuf xyz;   /* n x 3 f64 (reference layout)                         */
    m3d::DevBuf nrm;   /* n x 3 f64 or empty                                   */
    m3d::DevBuf pts32; /* n float4: centred fp32 x,y,z and |q|^2 of the centred point */
    m3d::DevBuf meta;  /* CloudMeta                                            */
    m3d::CloudMeta h_meta;
    /* host-buffer entry points only: the caller's normal array stays on the host (borrowed for the
     * call); the normals of the sampled points are uploaded per wave instead of all n of them */
    const double *h_nrm = nullptr;
    /* Morton-ordered copy for the culling score kernel (score_cull.cuh), built on first use */
    mutable m3d::DevBuf blob;  /* tiles x (1024 points + 32 cell spheres + 1 tile sphere) float4 */
    mutable m3d::DevBuf perm;  /* sorted position -> original point index (u32)                   */
    mutable m3d::DevBuf keys, hist; /* build scratch                                              */
    mutable bool sorted = false;
};

struct m3d_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    int sm_count = 148;
    uint64_t launches = 0;
    std::string err;
    cudaEvent_t ev[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};

    /* scratch (grow-only) */
    m3d::DevBuf d_samples, d_counts, d_counts_all, d_blk, d_part, d_small, d_inl, d_models, d_valid;
    std::vector<std::pair<const void *, size_t>> registered; /* caller buffers page-locked under M3D_FLAG_REGISTER_HOST */
    m3d::DevBuf d_tmp0, d_tmp1, d_tmp2, d_tmp3, d_tmp4, d_tmp5, d_queue, d_tiles, d_rownrm, d_rowmap, d_recs, d_draw, d_mtjump;
    m3d::PinBuf h_samples, h_counts, h_small, h_stage, h_rownrm;
    m3d_cloud *scratch_cloud = nullptr; /* staging cloud of the host-buffer entry points */
    /* chunked upload of the host-buffer fit (ransac.cu fit_host_chunked): a copy stream + one event per chunk */
    cudaStream_t copy_stream = nullptr;
    cudaEvent_t ev_chunk[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    m3d::DevBuf d_metas, d_models_all, d_valid_all;
    m3d::PinBuf h_metas, h_upload; /* h_upload: pinned staging of pageable uploads */
    struct CopyPool *pool = nullptr;
    struct m3d_feat_scratch *feat = nullptr; /* device buffers of the FPFH / ICP entry points (features.cu) */

    /* sharding / exchange */
    int rank = 0, world = 1;
    void *nccl_comm = nullptr;
    m3d_allgather_fn xfn = nullptr;
    void *xuser = nullptr;
    int x_on_device = 0;

    int fail(int code, const char *fmt, ...) {
        char buf[512];
        va_list ap;
        va_start(ap, fmt);
        vsnprintf(buf, sizeof buf, fmt, ap);
        va_end(ap);
        err = buf;
        return code;
    }
};

#define M3D_CUDA(ctx, call)                                                                  \
    do {                                                                                     \
        cudaError_t e__ = (call);                                                            \
        if (e__ != cudaSuccess)                                                              \
            return (ctx)->fail(M3D_ERR_CUDA, "%s:%d %s: %s", __FILE__, __LINE__, #call,      \
                               cudaGetErrorString(e__));                                     \
    } while (0)

#define M3D_LAUNCHED(ctx)                                                                    \
    do {                                                                                     \
        (ctx)->launches++;                                                                   \
        cudaError_t e__ = cudaGetLastError();                                                \
        if (e__ != cudaSuccess)                                                              \
            return (ctx)->fail(M3D_ERR_CUDA, "%s:%d kernel launch: %s", __FILE__, __LINE__,  \
                               cudaGetErrorString(e__));                                     \
    } while (0)

namespace m3d {
/* host -> device copy of a caller buffer on `stream`.  Pinned sources go straight to cudaMemcpyAsync; PAGEABLE ones
 * (what numpy / Open3D callers hold) are first copied into a pinned staging buffer by a few worker threads, piece by
 * piece, each piece handed to the copy engine as soon as it is staged -- the driver's own pageable path stages with one
 * thread (~18 GB/s measured), this one is bound by the DMA instead. */
int host_to_device(m3d_ctx *ctx, void *dst, const void *src, size_t bytes, cudaStream_t stream);
/* all-gather `bytes_per_rank` bytes per rank of device memory (NCCL or the caller's callback) */
int exchange_allgather(m3d_ctx *ctx, const void *d_send, void *d_recv, size_t bytes_per_rank);
/* ncclAllGather on `stream` (any stream of this context); only with a direct NCCL communicator (m3d_ctx_init_nccl) */
bool exchange_has_nccl(const m3d_ctx *ctx);
int exchange_allgather_nccl(m3d_ctx *ctx, const void *d_send, void *d_recv, size_t bytes_per_rank, cudaStream_t stream);
}  // namespace m3d
