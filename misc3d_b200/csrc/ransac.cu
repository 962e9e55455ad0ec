/*
 * ransac.cu -- host orchestration + C-ABI of the RANSAC primitive-fitting path
 * (reference: include/misc3d/common/ransac.h RANSAC<>::FitModel / FitModelParallel / RefineModel,
 * src/iterative_plane_segmentation.cpp SegmentPlaneIterative).
 *
 * The reference loop is sequential in its semantics (adaptive early exit, strict "better"
 * comparison in loop order).  Here hypotheses are scored in waves on the GPU; after each wave
 * the host replays the loop's bookkeeping over the wave's inlier counts in loop order
 * (scan.h), which reproduces best model / iteration count / stop index exactly.
 */
#include <algorithm>
#include <cmath>
#include <ctime>
#include <vector>

#include "context.h"
#include "loop_kernels.cuh"
#include "ransac_kernels.cuh"
#include "score_cull.cuh"
#include "score_cell.cuh"
#include "scan.h"

using namespace m3d;

namespace {

constexpr uint32_t kMaxWave = 1u << 16;

struct SmallDev { /* layout of ctx->d_small */
    double model[8];    /* minimal model of the hypothesis under refinement          */
    uint8_t valid[8];   /* its MinimalFit flag                                        */
    RefineMid mid;      /* pass-2 output                                              */
    RefineOut out;      /* pass-4 output                                              */
    double seq_err;     /* seq_err_kernel output                                      */
    unsigned long long resolves;
    uint32_t sample[8]; /* one sample row                                             */
    double row_nrm[12]; /* its normals in draw order (host-normals mode)              */
    BestRec best;       /* merged arg-best record of the last wave (device-side loop) */
    RowBreaks draw;     /* outcome of the device-side sample draw                     */
};

/* bounding box + fp32 copy of n points (3 kernels) */
int convert_points(m3d_ctx *ctx, const double *xyz, uint32_t n, CloudMeta *meta, float4 *pts32) {
    const int nb = std::max(1, std::min<int>(ctx->sm_count * 8, (int)((n + 255) / 256)));
    M3D_CUDA(ctx, ctx->d_part.reserve(sizeof(BBoxPart) * (size_t)nb));
    BBoxPart *bp = ctx->d_part.as<BBoxPart>();
    bbox_kernel<<<nb, 256, 0, ctx->stream>>>(xyz, n, bp);
    M3D_LAUNCHED(ctx);
    bbox_final_kernel<<<1, 256, 0, ctx->stream>>>(bp, nb, meta);
    M3D_LAUNCHED(ctx);
    convert_kernel<<<nb, 256, 0, ctx->stream>>>(xyz, n, meta, pts32);
    M3D_LAUNCHED(ctx);
    return M3D_OK;
}
/* Morton-ordered tile blob of n points: counting sort over a 2^bits cubed grid (histogram, 3-kernel scan, scatter)
 * + bounding spheres.  hist needs (1 << 3 bits) + bins / 2048 words, keys n words. */
size_t morton_hist_words(int bits) { return ((size_t)1 << (3 * bits)) + (((size_t)1 << (3 * bits)) / (kScanBlock * kScanItems)); }
int morton_sort(m3d_ctx *ctx, const float4 *pts32, uint32_t n, const CloudMeta *meta, uint32_t *keys, uint32_t *hist,
                float4 *blob, uint32_t *perm) {
    const int bits = morton_bits(n);
    const uint32_t bins = 1u << (3 * bits);
    const int scan_blocks = (int)(bins / (kScanBlock * kScanItems));
    uint32_t *bsum = hist + bins;
    const uint32_t ntiles = (n + kTile - 1) / kTile;
    M3D_CUDA(ctx, cudaMemsetAsync(hist, 0, sizeof(uint32_t) * (size_t)bins, ctx->stream));
    const int nb = std::max(1, std::min<int>(ctx->sm_count * 8, (int)((n + 255) / 256)));
    morton_hist_kernel<<<nb, 256, 0, ctx->stream>>>(pts32, n, meta, bits, keys, hist);
    M3D_LAUNCHED(ctx);
    scan_sums_kernel<<<scan_blocks, kScanBlock, 0, ctx->stream>>>(hist, bsum);
    M3D_LAUNCHED(ctx);
    scan_top_kernel<<<1, kScanBlock, 0, ctx->stream>>>(bsum, scan_blocks);
    M3D_LAUNCHED(ctx);
    scan_apply_kernel<<<scan_blocks, kScanBlock, 0, ctx->stream>>>(hist, bsum);
    M3D_LAUNCHED(ctx);
    morton_scatter_kernel<<<nb, 256, 0, ctx->stream>>>(pts32, n, keys, hist, blob, perm);
    M3D_LAUNCHED(ctx);
    tile_bounds_kernel<<<ntiles, kTile, 0, ctx->stream>>>(blob, n, meta);
    M3D_LAUNCHED(ctx);
    return M3D_OK;
}

int prepare_cloud(m3d_ctx *ctx, m3d_cloud *c) {
    const uint32_t n = (uint32_t)c->n;
    M3D_CUDA(ctx, c->pts32.reserve(sizeof(float4) * (size_t)std::max<uint32_t>(n, 1)));
    M3D_CUDA(ctx, c->meta.reserve(sizeof(CloudMeta)));
    if (int rc = convert_points(ctx, c->xyz.as<double>(), n, c->meta.as<CloudMeta>(), c->pts32.as<float4>())) return rc;
    M3D_CUDA(ctx, cudaMemcpyAsync(&c->h_meta, c->meta.p, sizeof(CloudMeta), cudaMemcpyDeviceToHost, ctx->stream));
    M3D_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return M3D_OK;
}

/* (re)fills a cloud object: buffers are grow-only, so a cached object costs no cudaMalloc */
int cloud_fill(m3d_ctx *ctx, m3d_cloud *c, const double *xyz, const double *nrm, size_t n, cudaMemcpyKind kind) {
    c->ctx = ctx;
    c->n = n;
    c->sorted = false;
    c->h_nrm = nullptr;
    c->has_normals = nrm != nullptr;
    const size_t bytes = sizeof(double) * 3 * std::max<size_t>(n, 1);
    M3D_CUDA(ctx, c->xyz.reserve(bytes));
    if (nrm) M3D_CUDA(ctx, c->nrm.reserve(bytes));
    if (n && kind == cudaMemcpyHostToDevice) {
        if (int rc = host_to_device(ctx, c->xyz.p, xyz, sizeof(double) * 3 * n, ctx->stream)) return rc;
        if (nrm) {
            M3D_CUDA(ctx, cudaStreamSynchronize(ctx->stream)); /* one staging buffer: the first upload has to drain */
            if (int rc = host_to_device(ctx, c->nrm.p, nrm, sizeof(double) * 3 * n, ctx->stream)) return rc;
        }
    } else if (n) {
        M3D_CUDA(ctx, cudaMemcpyAsync(c->xyz.p, xyz, sizeof(double) * 3 * n, kind, ctx->stream));
        if (nrm) M3D_CUDA(ctx, cudaMemcpyAsync(c->nrm.p, nrm, sizeof(double) * 3 * n, kind, ctx->stream));
    }
    return prepare_cloud(ctx, c);
}

int cloud_create(m3d_ctx *ctx, const double *xyz, const double *nrm, size_t n, cudaMemcpyKind kind,
                 m3d_cloud **out) {
    if (!ctx || !out || (n && !xyz)) return M3D_ERR_INVALID_ARG;
    if (n >= (size_t)kInvalidBit) return ctx->fail(M3D_ERR_INVALID_ARG, "clouds of >= 2^31 points are not supported");
    M3D_CUDA(ctx, cudaSetDevice(ctx->device));
    m3d_cloud *c = new m3d_cloud();
    const int rc = cloud_fill(ctx, c, xyz, nrm, n, kind);
    if (rc != M3D_OK) {
        m3d_cloud_free(c);
        return rc;
    }
    *out = c;
    return M3D_OK;
}

/* scoring path: M3D_SCORE_PATH=dense keeps every point-hypothesis pair (score_kernel); default
 * (cull) uses the Morton-ordered copy for clouds of >= kCullMinPoints finite points */
constexpr size_t kCullMinPoints = 2048;
int score_path() { /* 0 dense (score_kernel), 1 score_cull_kernel (round 1), 2 score_cell_kernel (default) */
    static int v = -1;
    if (v < 0) {
        const char *e = getenv("M3D_SCORE_PATH");
        v = (e && strcmp(e, "dense") == 0) ? 0 : ((e && strcmp(e, "cull") == 0) ? 1 : 2);
    }
    return v;
}
bool cull_enabled() { return score_path() != 0; }

/* builds (once per upload) the Morton-ordered tile blob of score_cull.cuh: counting sort over a
 * 128^3 grid (histogram, 3-kernel scan, scatter) + bounding spheres */
int ensure_sorted(m3d_ctx *ctx, const m3d_cloud *c) {
    if (c->sorted) return M3D_OK;
    if (!cull_enabled() || c->n < kCullMinPoints || c->h_meta.nonfinite) return M3D_OK;
    const uint32_t n = (uint32_t)c->n;
    const uint32_t ntiles = (n + kTile - 1) / kTile;
    M3D_CUDA(ctx, c->blob.reserve(sizeof(float4) * (size_t)ntiles * kBlobF4));
    M3D_CUDA(ctx, c->perm.reserve(sizeof(uint32_t) * (size_t)ntiles * kTile));
    M3D_CUDA(ctx, c->keys.reserve(sizeof(uint32_t) * (size_t)n));
    M3D_CUDA(ctx, c->hist.reserve(sizeof(uint32_t) * morton_hist_words(morton_bits(n))));
    if (int rc = morton_sort(ctx, c->pts32.as<float4>(), n, c->meta.as<CloudMeta>(), c->keys.as<uint32_t>(),
                             c->hist.as<uint32_t>(), c->blob.as<float4>(), c->perm.as<uint32_t>()))
        return rc;
    c->sorted = true;
    return M3D_OK;
}

template <int KIND, int THREADS, int HPT>
int launch_cull_t(m3d_ctx *ctx, const ScoreArgs &a, uint32_t ntiles) {
    const size_t smem = cull_smem_bytes<KIND, THREADS, HPT>();
    M3D_CUDA(ctx, cudaFuncSetAttribute(score_cull_kernel<KIND, THREADS, HPT>,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const uint32_t hb = (a.rows + THREADS * HPT - 1) / (THREADS * HPT);
    int per_sm = 0;
    M3D_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, score_cull_kernel<KIND, THREADS, HPT>,
                                                                 THREADS + 32, smem));
    const uint32_t slots = (uint32_t)ctx->sm_count * (uint32_t)std::max(per_sm, 1);
    uint32_t per_hb = std::max<uint32_t>(1, (slots + hb - 1) / hb);
    per_hb = std::min(per_hb, ntiles);
    M3D_CUDA(ctx, ctx->d_tiles.reserve(sizeof(uint32_t) * (size_t)hb));
    M3D_CUDA(ctx, cudaMemsetAsync(ctx->d_tiles.p, 0, sizeof(uint32_t) * (size_t)hb, ctx->stream));
    ScoreArgs b = a;
    b.tile_counter = ctx->d_tiles.as<uint32_t>();
    dim3 grid(hb, per_hb);
    score_cull_kernel<KIND, THREADS, HPT><<<grid, THREADS + 32, smem, ctx->stream>>>(b);
    M3D_LAUNCHED(ctx);
    return M3D_OK;
}

template <int KIND, int THREADS, int NH, bool STATS, bool PRE>
int launch_cell_tt(m3d_ctx *ctx, const ScoreArgs &a, uint32_t ntiles) {
    auto kern = score_cell_kernel<KIND, THREADS, NH, STATS, PRE>;
    const size_t smem = CellSmem<KIND, THREADS, NH>::bytes();
    M3D_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const uint32_t hb = (a.rows + NH - 1) / NH;
    int per_sm = 0;
    M3D_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, THREADS + 32, smem));
    const uint32_t slots = (uint32_t)ctx->sm_count * (uint32_t)std::max(per_sm, 1);
    uint32_t per_hb = std::max<uint32_t>(1, (slots + hb - 1) / hb);
    per_hb = std::min(per_hb, ntiles);
    M3D_CUDA(ctx, ctx->d_tiles.reserve(sizeof(uint32_t) * (size_t)hb));
    M3D_CUDA(ctx, cudaMemsetAsync(ctx->d_tiles.p, 0, sizeof(uint32_t) * (size_t)hb, ctx->stream));
    ScoreArgs b = a;
    b.tile_counter = ctx->d_tiles.as<uint32_t>();
    dim3 grid(hb, per_hb);
    kern<<<grid, THREADS + 32, smem, ctx->stream>>>(b);
    M3D_LAUNCHED(ctx);
    return M3D_OK;
}
template <int KIND, int THREADS, int NH, bool STATS>
int launch_cell_t(m3d_ctx *ctx, const ScoreArgs &a, uint32_t ntiles) {
    if (a.models_in) return launch_cell_tt<KIND, THREADS, NH, STATS, true>(ctx, a, ntiles);
    return launch_cell_tt<KIND, THREADS, NH, STATS, false>(ctx, a, ntiles);
}

/* ---- launch of the hot kernel for one (wave, kind) */
template <int KIND, int THREADS, int HPT>
int launch_score_t(m3d_ctx *ctx, const ScoreArgs &a, uint32_t ntiles) {
    const size_t smem = (size_t)kStages * kTile * sizeof(float4) + 2 * kStages * sizeof(uint64_t) + kStages * sizeof(uint32_t) + 16;
    M3D_CUDA(ctx, cudaFuncSetAttribute(score_kernel<KIND, THREADS, HPT>,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const uint32_t hb = (a.rows + THREADS * HPT - 1) / (THREADS * HPT);
    /* one resident wave: CTAs per hypothesis block = resident CTA slots / hypothesis blocks; every
     * group pulls tiles from its own cursor */
    int per_sm = 0;
    M3D_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, score_kernel<KIND, THREADS, HPT>,
                                                                 THREADS + 32, smem));
    const uint32_t slots = (uint32_t)ctx->sm_count * (uint32_t)std::max(per_sm, 1);
    uint32_t per_hb = std::max<uint32_t>(1, (slots + hb - 1) / hb); /* late CTAs find the cursor advanced: work-conserving */
    per_hb = std::min(per_hb, ntiles);
    M3D_CUDA(ctx, ctx->d_tiles.reserve(sizeof(uint32_t) * (size_t)hb));
    M3D_CUDA(ctx, cudaMemsetAsync(ctx->d_tiles.p, 0, sizeof(uint32_t) * (size_t)hb, ctx->stream));
    ScoreArgs b = a;
    b.tile_counter = ctx->d_tiles.as<uint32_t>();
    dim3 grid(hb, per_hb);
    score_kernel<KIND, THREADS, HPT><<<grid, THREADS + 32, smem, ctx->stream>>>(b); /* + producer warp */
    M3D_LAUNCHED(ctx);
    return M3D_OK;
}

/* tuning knob for experiments: M3D_SCORE_VARIANT=<threads>x<hypotheses per thread> */
int score_variant_override() {
    static int v = -1;
    if (v < 0) {
        v = 0;
        if (const char *e = getenv("M3D_SCORE_VARIANT")) {
            int t = 0, h = 0;
            if (sscanf(e, "%dx%d", &t, &h) == 2) v = t * 16 + h;
        }
    }
    return v;
}

/* one scoring launch + its guard-band resolve; `cull` selects score_cull_kernel (needs a.blob) */
template <int KIND>
int launch_one(m3d_ctx *ctx, ScoreArgs a, uint32_t ntiles, bool cull) {
    if (a.rows == 0) return M3D_OK;
    if (!cull) a.blob = nullptr; /* resolve_queue_kernel: the dense kernel queues original point indices */
    M3D_CUDA(ctx, cudaMemsetAsync(a.queue_count, 0, sizeof(uint32_t), ctx->stream));
    int rc;
    const int var = score_variant_override();
    if (cull && score_path() == 2) {
        /* 1024 hypotheses per CTA (one CTA of 16 consumer warps per SM) once the launch fills the machine that
         * way: longer per-cell lists (fewer half-empty passes), half the tile tests and barriers per pair */
        static const int cell_var = getenv("M3D_CELL_VARIANT") ? atoi(getenv("M3D_CELL_VARIANT")) : 0;
        if (a.flags & M3D_FLAG_STATS)
            rc = (cell_var == 256) ? launch_cell_t<KIND, 256, 512, true>(ctx, a, ntiles) : launch_cell_t<KIND, 512, 1024, true>(ctx, a, ntiles);
        else if (cell_var == 512 || (cell_var == 0 && a.rows >= 4096))
            rc = launch_cell_t<KIND, 512, 1024, false>(ctx, a, ntiles);
        else if (cell_var == 256 || (cell_var == 0 && a.rows >= 1024))
            rc = launch_cell_t<KIND, 256, 512, false>(ctx, a, ntiles);
        else
            rc = launch_cell_t<KIND, 128, 128, false>(ctx, a, ntiles);
    } else if (cull) {
        switch (var) {
            case 128 * 16 + 1: rc = launch_cull_t<KIND, 128, 1>(ctx, a, ntiles); break;
            case 128 * 16 + 2: rc = launch_cull_t<KIND, 128, 2>(ctx, a, ntiles); break;
            case 256 * 16 + 1: rc = launch_cull_t<KIND, 256, 1>(ctx, a, ntiles); break;
            case 256 * 16 + 4: rc = launch_cull_t<KIND, 256, 4>(ctx, a, ntiles); break;
            default:
                rc = (a.rows >= 2048) ? launch_cull_t<KIND, 256, 2>(ctx, a, ntiles)
                                      : launch_cull_t<KIND, 128, 1>(ctx, a, ntiles);
        }
    } else {
        switch (var) {
            case 128 * 16 + 1: rc = launch_score_t<KIND, 128, 1>(ctx, a, ntiles); break;
            case 128 * 16 + 2: rc = launch_score_t<KIND, 128, 2>(ctx, a, ntiles); break;
            case 128 * 16 + 4: rc = launch_score_t<KIND, 128, 4>(ctx, a, ntiles); break;
            case 256 * 16 + 1: rc = launch_score_t<KIND, 256, 1>(ctx, a, ntiles); break;
            case 256 * 16 + 2: rc = launch_score_t<KIND, 256, 2>(ctx, a, ntiles); break;
            case 256 * 16 + 4: rc = launch_score_t<KIND, 256, 4>(ctx, a, ntiles); break;
            default:
                rc = (a.rows >= 2048) ? launch_score_t<KIND, 256, 2>(ctx, a, ntiles)
                                      : launch_score_t<KIND, 128, 1>(ctx, a, ntiles);
        }
    }
    if (rc) return rc;
    resolve_queue_kernel<KIND><<<ctx->sm_count * 2, 256, 0, ctx->stream>>>(a);
    M3D_LAUNCHED(ctx);
    return M3D_OK;
}

/* launches of at least this many rows (default 12288) are pre-sorted into culled / dense hypotheses (cull_classify_kernel):
 * costs one small kernel + a 8-byte read-back, pays when part of the hypotheses pass through most of the
 * cloud.  A hypothesis goes to the dense kernel when more than kDenseAbove of its sampled cells survive
 * (measured break-even of the two kernels: ~25 % surviving pairs). */
constexpr float kDenseAbove = 0.25f;
uint32_t classify_min_rows() { /* M3D_CLASSIFY_MIN_ROWS overrides (tuning) */
    static uint32_t v = 0;
    if (!v) {
        const char *e = getenv("M3D_CLASSIFY_MIN_ROWS");
        v = e ? (uint32_t)std::max(1l, atol(e)) : 12288u;
    }
    return v;
}

template <int KIND>
int launch_score(m3d_ctx *ctx, const m3d_cloud *c, ScoreArgs a, bool exact_only) {
    const uint32_t ntiles = std::max<uint32_t>(1, (a.n + kTile - 1) / kTile);
    /* per-launch scratch: the minimal models and the queue of guard-band (hypothesis, point) pairs */
    constexpr uint32_t kQueueCap = 1u << 22;
    M3D_CUDA(ctx, ctx->d_models.reserve(sizeof(double) * 8 * (size_t)a.rows));
    M3D_CUDA(ctx, ctx->d_queue.reserve(sizeof(uint2) * (size_t)kQueueCap + 16));
    a.models = ctx->d_models.as<double>();
    a.queue_count = ctx->d_queue.as<uint32_t>();
    a.queue = reinterpret_cast<uint2 *>(ctx->d_queue.as<char>() + 16);
    a.queue_cap = kQueueCap;
    if (exact_only || c->h_meta.nonfinite) {
        const uint32_t hb = (a.rows + 127) / 128;
        uint32_t chunks = std::min<uint32_t>(ntiles, std::max<uint32_t>(1, (4u * ctx->sm_count * 4u) / hb));
        a.chunk_tiles = (ntiles + chunks - 1) / chunks;
        chunks = (ntiles + a.chunk_tiles - 1) / a.chunk_tiles;
        score_exact_kernel<KIND><<<dim3(hb, chunks), 128, 0, ctx->stream>>>(a);
        M3D_LAUNCHED(ctx);
        return M3D_OK;
    }
    const bool cull = a.blob && cull_enabled();
    /* score_cell_kernel evaluates a hypothesis that crosses most of the cloud at dense-kernel cost: only the
     * round-1 culling kernel needs the pre-sort into culled / dense hypotheses */
    if (cull && (score_path() == 1 || (a.flags & M3D_FLAG_CLASSIFY)) &&
        (a.rows >= classify_min_rows() || (a.flags & M3D_FLAG_CLASSIFY))) {
        /* row_map: [rows for the culling kernel ...   ... rows for the dense kernel], part = the two sizes */
        M3D_CUDA(ctx, ctx->d_rowmap.reserve(sizeof(uint32_t) * ((size_t)a.rows + 2)));
        uint32_t *map = ctx->d_rowmap.as<uint32_t>() + 2, *part = ctx->d_rowmap.as<uint32_t>();
        M3D_CUDA(ctx, cudaMemsetAsync(part, 0, 2 * sizeof(uint32_t), ctx->stream));
        const uint32_t stride = std::max<uint32_t>(1, ntiles / 96); /* ~96 sampled tiles per hypothesis */
        cull_classify_kernel<KIND><<<(a.rows + 127) / 128, 128, 0, ctx->stream>>>(a, ntiles, stride, kDenseAbove, map, part);
        M3D_LAUNCHED(ctx);
        uint32_t h_part[2] = {0, 0};
        M3D_CUDA(ctx, cudaMemcpyAsync(h_part, part, sizeof h_part, cudaMemcpyDeviceToHost, ctx->stream));
        M3D_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        if (h_part[0] + h_part[1] != a.rows) return ctx->fail(M3D_ERR_INTERNAL, "hypothesis classification lost rows");
        ScoreArgs lo = a, hi = a;
        lo.row_map = map;
        lo.rows = h_part[0];
        hi.row_map = map + (a.rows - h_part[1]);
        hi.rows = h_part[1];
        if (int rc = launch_one<KIND>(ctx, lo, ntiles, true)) return rc;
        return launch_one<KIND>(ctx, hi, ntiles, false);
    }
    return launch_one<KIND>(ctx, a, ntiles, cull);
}

int launch_score_kind(m3d_ctx *ctx, int kind, const m3d_cloud *c, const ScoreArgs &a, bool exact_only) {
    switch (kind) {
        case kPlane:
            return launch_score<kPlane>(ctx, c, a, exact_only);
        case kSphere:
            return launch_score<kSphere>(ctx, c, a, exact_only);
        default:
            return launch_score<kCylinder>(ctx, c, a, exact_only);
    }
}

/* ---- RefineModel passes */
struct RefineBufs {
    uint32_t nblk;
    uint32_t *blk_cnt, *blk_off;
    double *blk_part, *blk_mom;
};
int refine_bufs(m3d_ctx *ctx, uint32_t n, RefineBufs *rb) {
    rb->nblk = std::max<uint32_t>(1, (n + kRBlockPts - 1) / kRBlockPts);
    const size_t per = 2 * sizeof(uint32_t) + 14 * sizeof(double);
    M3D_CUDA(ctx, ctx->d_blk.reserve(per * rb->nblk + 64));
    char *p = ctx->d_blk.as<char>();
    rb->blk_part = reinterpret_cast<double *>(p);
    rb->blk_mom = rb->blk_part + 4 * (size_t)rb->nblk;
    rb->blk_cnt = reinterpret_cast<uint32_t *>(rb->blk_mom + 10 * (size_t)rb->nblk);
    rb->blk_off = rb->blk_cnt + rb->nblk;
    return M3D_OK;
}

template <int KIND>
int count_pass(m3d_ctx *ctx, const double *xyz, uint32_t n, const double *d_model, double thr,
               const RefineBufs &rb, RefineMid *d_mid) {
    refine_count_kernel<KIND><<<rb.nblk, kRB, 0, ctx->stream>>>(xyz, n, d_model, thr, rb.blk_cnt, rb.blk_part);
    M3D_LAUNCHED(ctx);
    refine_scan_kernel<<<1, 1024, 0, ctx->stream>>>(rb.blk_cnt, rb.blk_part, rb.nblk, rb.blk_off, d_mid);
    M3D_LAUNCHED(ctx);
    return M3D_OK;
}
template <int KIND, bool SEG>
int write_pass(m3d_ctx *ctx, const double *xyz, uint32_t n, const double *d_model, double thr,
               const RefineBufs &rb, const RefineMid *d_mid, unsigned long long *d_inl, RefineOut *d_out,
               const SegArgs &seg) {
    refine_write_kernel<KIND, SEG><<<rb.nblk, kRB, 0, ctx->stream>>>(xyz, n, d_model, thr, rb.blk_off, d_mid,
                                                                    d_inl, rb.blk_mom, seg);
    M3D_LAUNCHED(ctx);
    refine_final_kernel<KIND><<<1, 256, 0, ctx->stream>>>(rb.blk_mom, rb.nblk, d_mid, d_model, d_out);
    M3D_LAUNCHED(ctx);
    return M3D_OK;
}
int count_pass_kind(m3d_ctx *ctx, int kind, const double *xyz, uint32_t n, const double *d_model, double thr,
                    const RefineBufs &rb, RefineMid *d_mid) {
    switch (kind) {
        case kPlane:
            return count_pass<kPlane>(ctx, xyz, n, d_model, thr, rb, d_mid);
        case kSphere:
            return count_pass<kSphere>(ctx, xyz, n, d_model, thr, rb, d_mid);
        default:
            return count_pass<kCylinder>(ctx, xyz, n, d_model, thr, rb, d_mid);
    }
}
int write_pass_kind(m3d_ctx *ctx, int kind, const double *xyz, uint32_t n, const double *d_model, double thr,
                    const RefineBufs &rb, const RefineMid *d_mid, unsigned long long *d_inl, RefineOut *d_out) {
    SegArgs none{};
    switch (kind) {
        case kPlane:
            return write_pass<kPlane, false>(ctx, xyz, n, d_model, thr, rb, d_mid, d_inl, d_out, none);
        case kSphere:
            return write_pass<kSphere, false>(ctx, xyz, n, d_model, thr, rb, d_mid, d_inl, d_out, none);
        default:
            return write_pass<kCylinder, false>(ctx, xyz, n, d_model, thr, rb, d_mid, d_inl, d_out, none);
    }
}
int fit_rows_kind(m3d_ctx *ctx, int kind, const double *xyz, const double *nrm, const uint32_t *d_samples,
                  uint32_t rows, double *d_models, uint8_t *d_valid, const double *d_row_nrm = nullptr) {
    const int nb = (rows + 127) / 128;
    switch (kind) {
        case kPlane:
            minimal_fit_rows_kernel<kPlane><<<nb, 128, 0, ctx->stream>>>(xyz, nrm, d_samples, rows, d_models, d_valid, d_row_nrm);
            break;
        case kSphere:
            minimal_fit_rows_kernel<kSphere><<<nb, 128, 0, ctx->stream>>>(xyz, nrm, d_samples, rows, d_models, d_valid, d_row_nrm);
            break;
        default:
            minimal_fit_rows_kernel<kCylinder><<<nb, 128, 0, ctx->stream>>>(xyz, nrm, d_samples, rows, d_models, d_valid, d_row_nrm);
    }
    M3D_LAUNCHED(ctx);
    return M3D_OK;
}
int seq_err_kind(m3d_ctx *ctx, int kind, const double *xyz, const unsigned long long *d_inl,
                 unsigned long long n_inl, const double *d_model, double *d_out) {
    switch (kind) {
        case kPlane:
            seq_err_kernel<kPlane><<<1, 32, 0, ctx->stream>>>(xyz, d_inl, n_inl, d_model, d_out);
            break;
        case kSphere:
            seq_err_kernel<kSphere><<<1, 32, 0, ctx->stream>>>(xyz, d_inl, n_inl, d_model, d_out);
            break;
        default:
            seq_err_kernel<kCylinder><<<1, 32, 0, ctx->stream>>>(xyz, d_inl, n_inl, d_model, d_out);
    }
    M3D_LAUNCHED(ctx);
    return M3D_OK;
}

/* The device view of a cloud the fit runs on (segmentation swaps these between rounds). */
struct CloudView {
    const double *xyz;
    const double *nrm;
    const float4 *pts32;
    const CloudMeta *meta;
    uint32_t n;
    bool nonfinite;
    const float4 *blob = nullptr; /* Morton-ordered copy (null: dense scoring only) */
    const uint32_t *perm = nullptr;
    const double *h_nrm = nullptr; /* normals left on the host (host-buffer entry point): upload per sample */
    struct ChunkPlan *chunks = nullptr; /* host-buffer fit whose upload is still in flight (fit_host_chunked) */
};

/* The chunked host-buffer fit.  The caller's (pinned) cloud is uploaded in `count` chunks on a copy stream; every
 * chunk is an independent cloud for the purposes of counting -- own bounding box / fp32 copy / Morton-ordered tiles
 * -- and EvaluateModel's count is a sum over points, so the scoring kernel runs once per chunk, as soon as the chunk
 * has arrived, while the next one is still on the PCIe bus.  The minimal models need sample points from anywhere in
 * the cloud: they are solved beforehand by a kernel that gathers the k x H sample points straight from the pinned host
 * buffer (zero-copy).  xyz / pts32 / blob / perm of the view are the buffers of the WHOLE cloud (chunks are tile
 * aligned); RefineModel runs on the complete upload as before. */
struct ChunkPlan {
    int count = 0;
    uint32_t begin[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0}; /* point offsets, multiples of kTile; begin[count] = n */
    const double *dev_of_host_xyz = nullptr;          /* device-mapped address of the caller's pinned points  */
    const double *dev_of_host_nrm = nullptr;          /* ... normals (cylinder) or null                        */
    CloudMeta *metas = nullptr;                       /* device, one per chunk                                 */
    uint32_t *keys = nullptr, *hist = nullptr;        /* Morton sort scratch (re-used chunk after chunk)       */
    bool prepped[8] = {false, false, false, false, false, false, false, false};
    bool pre_models = false;  /* count >= 2: minimal models solved beforehand from the pinned buffer (zero-copy)   */
    const double *h_xyz = nullptr; /* the caller's buffer and its destination: the copies are issued by the fit, */
    double *d_xyz = nullptr;       /* right behind the device-side sample draw (so that the two overlap)          */
    bool copies_issued = false;
    bool shard_upload = false; /* multi-rank fit over NCCL, one chunk: upload 1/R of the cloud, all-gather the rest */
    int issue_copies(m3d_ctx *ctx) {
        if (copies_issued) return M3D_OK;
        copies_issued = true;
        if (shard_upload) {
            /* every rank holds the same cloud in host memory: rank r uploads the r-th slice over its own PCIe link and
             * the slices are all-gathered over NVLink (in place), instead of R identical 24 n-byte uploads that
             * share the host's PCIe uplinks */
            const size_t tot = 3 * (size_t)begin[count], R = (size_t)ctx->world;
            const size_t S = (tot + R - 1) / R, off = S * (size_t)ctx->rank;
            const size_t len = off < tot ? std::min(S, tot - off) : 0;
            if (len)
                if (int rc = host_to_device(ctx, d_xyz + off, h_xyz + off, sizeof(double) * len, ctx->copy_stream)) return rc;
            if (int rc = exchange_allgather_nccl(ctx, d_xyz + off, d_xyz, sizeof(double) * S, ctx->copy_stream)) return rc;
            M3D_CUDA(ctx, cudaEventRecord(ctx->ev_chunk[0], ctx->copy_stream));
            return M3D_OK;
        }
        for (int i = 0; i < count; ++i) {
            const size_t b = begin[i], cnt = begin[i + 1] - begin[i];
            if (cnt)
                if (int rc = host_to_device(ctx, d_xyz + 3 * b, h_xyz + 3 * b, sizeof(double) * 3 * cnt, ctx->copy_stream)) return rc;
            M3D_CUDA(ctx, cudaEventRecord(ctx->ev_chunk[i], ctx->copy_stream));
        }
        return M3D_OK;
    }
};
constexpr int kRetryUnchunked = 1001; /* internal: a chunk holds NaN / inf coordinates, use the plain upload path */

/* one full FitModel on a device-resident cloud.  On return the minimal best model sits in
 * d_small->model, pass 1+2 results in d_small->mid, and (when `seg` is null) the ascending
 * inlier indices in ctx->d_inl and the refined model in d_small->out. */
struct FitResult {
    m3d_ransac_stats st;
    double minimal[8];
    double refined[8];
    unsigned long long n_inl;
    int ret;
};

/* tuning / test knobs: M3D_LOOP=host replays every wave on the host (the round-1 path);
 * M3D_SAMPLER=host draws every sample table on the host */
bool env_is(const char *name, const char *value) {
    const char *e = getenv(name);
    return e && strcmp(e, value) == 0;
}
bool loop_on_host() {
    static const bool v = env_is("M3D_LOOP", "host");
    return v;
}
/* M3D_TRACE=1: host wall-clock per phase of m3d_ransac_fit_cloud, summed and printed at exit (a tuning aid) */
struct HostTrace {
    bool on = getenv("M3D_TRACE") != nullptr;
    double acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    double t = 0;
    unsigned long long calls = 0;
    bool armed = false; /* only fits of a resident cloud (m3d_ransac_fit_cloud) are traced */
    long skip = getenv("M3D_TRACE") ? atol(getenv("M3D_TRACE")) : 0; /* value = warm-up fits to leave out */
    void mark(int i) {
        if (!on || !armed) return;
        if (skip > 0) {
            t = now();
            return;
        }
        const double n = now();
        if (t != 0) acc[i] += n - t;
        t = n;
    }
    static double now() {
        timespec ts;
        clock_gettime(CLOCK_MONOTONIC, &ts);
        return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6;
    }
    ~HostTrace() {
        if (!on || !calls) return;
        const char *r = getenv("RANK");
        fprintf(stderr,
                "[m3d trace rank %s] %llu fits, ms per fit: between-calls %.4f  setup+draw-enqueue %.4f  score-enqueue %.4f  "
                "exchange-enqueue %.4f  refine-enqueue %.4f  sync-wait %.4f  finish %.4f  deliver %.4f\n",
                r ? r : "0", calls, acc[0] / calls, acc[1] / calls, acc[2] / calls, acc[3] / calls, acc[4] / calls,
                acc[5] / calls, acc[6] / calls, acc[7] / calls);
    }
};
HostTrace g_trace;
bool sampler_on_host() {
    static const bool v = env_is("M3D_SAMPLER", "host");
    return v;
}

/* ---- the sample table of rows [0, rows) of the stream (seed, n), k indices per row, drawn ON THE DEVICE
 * into `d_table` (loop_kernels.cuh).  *d_status receives the outcome asynchronously (status != 0: the draw
 * gave up -- too many duplicates -- and the caller has to fall back to the host draw). */
bool device_draw_eligible(uint32_t n, int k, uint64_t rows) {
    if (sampler_on_host() || n < 2 || k < 2 || k > 4) return false;
    const double len = (double)rows * k + 4096.0;
    if (len >= 2.0e9) return false;
    return len * (0.5 * k * (k - 1)) / (double)n <= 0.25 * kDupCap; /* expected rows with a duplicate */
}
int draw_table_device(m3d_ctx *ctx, uint32_t seed, uint32_t n, int k, uint32_t rows, uint32_t *d_table,
                      RowBreaks *d_status) {
    const uint32_t nblocks = (uint32_t)(((uint64_t)rows * k + 4096 + 623) / 624);
    const uint32_t len = nblocks * 624;
    M3D_CUDA(ctx, ctx->d_draw.reserve(sizeof(uint32_t) * ((size_t)len + kDupCap) + sizeof(uint2) * (kDupCap + 2)));
    uint32_t *stream = ctx->d_draw.as<uint32_t>();
    uint32_t *list = stream + len;
    uint2 *brk = reinterpret_cast<uint2 *>(list + kDupCap + ((len + kDupCap) & 1));
    MtInit init;
    init.mt[0] = seed;
    for (uint32_t i = 1; i < 624; ++i) init.mt[i] = 1812433253u * (init.mt[i - 1] ^ (init.mt[i - 1] >> 30)) + i;
    const uint64_t magic = UINT64_MAX / n + 1;
    M3D_CUDA(ctx, cudaMemsetAsync(d_status, 0, sizeof(RowBreaks), ctx->stream));
    /* long tables: the stream is produced in segments that start from jumped-ahead states (loop_kernels.cuh) */
    static const bool no_jump = env_is("M3D_MT_JUMP", "0");
    if (nblocks >= 2u * kMtSegBlocks && !no_jump) {
        const uint32_t nseg = std::min<uint32_t>(kMtMaxSegments, (nblocks + kMtSegBlocks - 1) / kMtSegBlocks);
        if (!ctx->d_mtjump.p) {
            M3D_CUDA(ctx, ctx->d_mtjump.reserve(sizeof kMtJump));
            M3D_CUDA(ctx, cudaMemcpyAsync(ctx->d_mtjump.p, kMtJump, sizeof kMtJump, cudaMemcpyHostToDevice, ctx->stream));
        }
        mt_stream_kernel<<<1, 256, 0, ctx->stream>>>(init, kMtPrefixBlocks, stream, nseg, kMtSegBlocks);
        M3D_LAUNCHED(ctx);
        mt_jump_kernel<<<dim3(kMtJumpChunks, nseg - 1), 640, 0, ctx->stream>>>(stream, ctx->d_mtjump.as<uint32_t>(), kMtSegBlocks);
        M3D_LAUNCHED(ctx);
        mt_segments_kernel<<<nseg, 256, 0, ctx->stream>>>(stream, kMtSegBlocks, nseg, nblocks);
        M3D_LAUNCHED(ctx);
    } else {
        mt_stream_kernel<<<1, 256, 0, ctx->stream>>>(init, nblocks, stream, 0, 0);
        M3D_LAUNCHED(ctx);
    }
    const int nb = std::max(1, std::min<int>(ctx->sm_count * 4, (int)((len + 255) / 256)));
    stream_finish_kernel<<<nb, 256, 0, ctx->stream>>>(stream, len, n, magic);
    M3D_LAUNCHED(ctx);
    dup_positions_kernel<<<nb, 256, 0, ctx->stream>>>(stream, len, k, d_status, list);
    M3D_LAUNCHED(ctx);
    row_breaks_kernel<<<1, 1024, 0, ctx->stream>>>(stream, len, k, rows, d_status, list, brk);
    M3D_LAUNCHED(ctx);
    build_rows_kernel<<<std::max(1, std::min<int>(ctx->sm_count * 4, (int)((rows + 255) / 256))), 256, 0, ctx->stream>>>(
        stream, len, k, rows, d_status, brk, d_table);
    M3D_LAUNCHED(ctx);
    return M3D_OK;
}

constexpr int kRetryOnHost = 1000; /* internal: the device-side loop gave up, run the host loop */

/* bounding box, fp32 copy and Morton-ordered tiles of chunk c (the kernels of prepare_cloud / ensure_sorted on a
 * slice); enqueued behind the chunk's arrival event */
int prep_chunk(m3d_ctx *ctx, const CloudView &v, ChunkPlan &pl, int c) {
    if (pl.prepped[c]) return M3D_OK;
    M3D_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ctx->ev_chunk[c], 0));
    const uint32_t b = pl.begin[c], n = pl.begin[c + 1] - pl.begin[c];
    float4 *pts32 = const_cast<float4 *>(v.pts32) + b;
    if (int rc = convert_points(ctx, v.xyz + 3 * (size_t)b, n, pl.metas + c, pts32)) return rc;
    if (int rc = morton_sort(ctx, pts32, n, pl.metas + c, pl.keys, pl.hist,
                             const_cast<float4 *>(v.blob) + (size_t)(b / kTile) * kBlobF4, const_cast<uint32_t *>(v.perm) + b))
        return rc;
    pl.prepped[c] = true;
    return M3D_OK;
}

struct Fit {
    m3d_ctx *ctx;
    int kind, k;
    const m3d_cloud *cloud_for_flags;
    const CloudView &v;
    const m3d_ransac_params &p;
    SegArgs *seg;
    FitResult *res;
    uint32_t n;
    uint64_t H;
    bool exact_only, host_nrm;
    SmallDev *ds, *hs;
    RefineBufs rb;
    const double *one_row_nrm;
    cudaEvent_t ev_a, ev_b, ev_s0, ev_s1, ev_r0, ev_r1, ev_d0, ev_d1;
    float score_ms = 0;
    bool drew_on_device = false;
    uint64_t evaluated = 0;
    int R, rank;

    /* where sample rows live: the current wave [wave_base, wave_base + wave_rows) in ctx->d_samples (and the
     * normals of its sample points in ctx->d_rownrm in host-normals mode) + a host copy of the best row */
    uint64_t wave_base = 0, wave_rows = 0;
    bool have_saved = false;
    uint64_t saved_idx = 0;
    uint32_t saved_sample[4] = {0, 0, 0, 0};
    double saved_nrm[12];

    Fit(m3d_ctx *c, int kd, const m3d_cloud *cf, const CloudView &vv, const m3d_ransac_params &pp, SegArgs *sg, FitResult *r)
        : ctx(c), kind(kd), k(sample_size(kd)), cloud_for_flags(cf), v(vv), p(pp), seg(sg), res(r) {}

    int setup() {
        H = p.max_iteration;
        n = v.n;
        exact_only = (p.flags & M3D_FLAG_EXACT_ONLY) != 0 || v.nonfinite;
        memset(res, 0, sizeof *res);
        res->st.stop_index = H;
        M3D_CUDA(ctx, ctx->d_small.reserve(sizeof(SmallDev)));
        M3D_CUDA(ctx, ctx->h_small.reserve(sizeof(SmallDev)));
        ds = ctx->d_small.as<SmallDev>();
        hs = ctx->h_small.as<SmallDev>();
        if (int rc = refine_bufs(ctx, n, &rb)) return rc;
        M3D_CUDA(ctx, ctx->d_inl.reserve(sizeof(unsigned long long) * (size_t)std::max<uint32_t>(n, 1)));
        M3D_CUDA(ctx, cudaMemsetAsync(&ds->resolves, 0, sizeof(unsigned long long), ctx->stream));
        ev_a = ctx->ev[0], ev_b = ctx->ev[1], ev_s0 = ctx->ev[2], ev_s1 = ctx->ev[3];
        ev_r0 = ctx->ev[4], ev_r1 = ctx->ev[5], ev_d0 = ctx->ev[6], ev_d1 = ctx->ev[7];
        M3D_CUDA(ctx, cudaEventRecord(ev_a, ctx->stream));
        /* host-normals mode: only the cylinder reads normals, and only those of its sample points */
        host_nrm = (kind == kCylinder) && v.nrm == nullptr && v.h_nrm != nullptr;
        one_row_nrm = host_nrm ? ds->row_nrm : nullptr;
        R = ctx->world, rank = ctx->rank;
        return M3D_OK;
    }

    void save_best(uint64_t idx, const uint32_t *row) {
        have_saved = true;
        saved_idx = idx;
        for (int j = 0; j < k; ++j) saved_sample[j] = row[j];
        if (host_nrm)
            for (int j = 0; j < k; ++j)
                for (int c = 0; c < 3; ++c) saved_nrm[3 * j + c] = v.h_nrm[3 * (size_t)row[j] + c];
    }
    /* sample row `row` (+ its normals) -> ds->sample / ds->row_nrm.  Pageable copies of a few bytes are staged
     * by the driver before cudaMemcpyAsync returns */
    int stage_row(uint64_t row) {
        if (row >= wave_base && row < wave_base + wave_rows) {
            const size_t off = (size_t)(row - wave_base) * k;
            M3D_CUDA(ctx, cudaMemcpyAsync(ds->sample, ctx->d_samples.as<uint32_t>() + off, sizeof(uint32_t) * k,
                                          cudaMemcpyDeviceToDevice, ctx->stream));
            if (host_nrm)
                M3D_CUDA(ctx, cudaMemcpyAsync(ds->row_nrm, ctx->d_rownrm.as<double>() + 3 * off, sizeof(double) * 3 * k,
                                              cudaMemcpyDeviceToDevice, ctx->stream));
            return 0;
        }
        if (!have_saved || saved_idx != row) return ctx->fail(M3D_ERR_INTERNAL, "sample row %llu is not available", (unsigned long long)row);
        M3D_CUDA(ctx, cudaMemcpyAsync(ds->sample, saved_sample, sizeof(uint32_t) * k, cudaMemcpyHostToDevice, ctx->stream));
        if (host_nrm)
            M3D_CUDA(ctx, cudaMemcpyAsync(ds->row_nrm, saved_nrm, sizeof(double) * 3 * k, cudaMemcpyHostToDevice, ctx->stream));
        return 0;
    }
    /* evaluates hypothesis `row` alone (tie-breaks): model -> ds->model, then pass 1+2 -> ds->mid;
     * optionally the index-order error */
    int eval_row(uint64_t row, bool exact, uint64_t expect_cnt, double *rmse) {
        if (int rc = stage_row(row)) return rc;
        if (int rc = fit_rows_kind(ctx, kind, v.xyz, v.nrm, ds->sample, 1, ds->model, ds->valid, one_row_nrm)) return rc;
        if (int rc = count_pass_kind(ctx, kind, v.xyz, n, ds->model, p.threshold, rb, &ds->mid)) return rc;
        if (exact) {
            if (int rc = write_pass_kind(ctx, kind, v.xyz, n, ds->model, p.threshold, rb, &ds->mid,
                                         ctx->d_inl.as<unsigned long long>(), &ds->out))
                return rc;
            if (int rc = seq_err_kind(ctx, kind, v.xyz, ctx->d_inl.as<unsigned long long>(), expect_cnt, ds->model, &ds->seq_err))
                return rc;
        }
        M3D_CUDA(ctx, cudaMemcpyAsync(hs, ds, sizeof(SmallDev), cudaMemcpyDeviceToHost, ctx->stream));
        M3D_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        if (hs->mid.n_inl != expect_cnt)
            return ctx->fail(M3D_ERR_INTERNAL, "hypothesis %llu: scoring kernel counted %llu inliers, fp64 pass %llu",
                             (unsigned long long)row, (unsigned long long)expect_cnt, (unsigned long long)hs->mid.n_inl);
        const double e = exact ? hs->seq_err : hs->mid.err;
        *rmse = expect_cnt ? e / std::sqrt((double)expect_cnt) : 1e10;
        return 0;
    }

    ScoreArgs score_args(uint32_t l0, uint32_t l1) const {
        ScoreArgs a{};
        a.pts32 = v.pts32;
        a.blob = (p.flags & M3D_FLAG_DENSE) ? nullptr : v.blob;
        a.perm = v.perm;
        a.flags = p.flags;
        a.xyz = v.xyz;
        a.nrm = v.nrm;
        a.row_nrm = host_nrm ? ctx->d_rownrm.as<double>() : nullptr;
        a.meta = v.meta;
        a.samples = ctx->d_samples.as<uint32_t>();
        a.counts = ctx->d_counts.as<uint32_t>() + l0;
        a.resolves = &ds->resolves;
        a.thr = p.threshold;
        a.n = n;
        a.row_begin = l0; /* shard-local */
        a.rows = l1 - l0;
        a.shard_world = (uint32_t)R;
        a.shard_rank = (uint32_t)rank;
        return a;
    }

    /* RefineModel (ransac.h:534-549) on the minimal model of the sample row staged in ds->sample */
    int enqueue_refine() {
        M3D_CUDA(ctx, cudaEventRecord(ev_r0, ctx->stream));
        const int rc = enqueue_refine_passes();
        if (rc) return rc;
        M3D_CUDA(ctx, cudaEventRecord(ev_r1, ctx->stream));
        return M3D_OK;
    }
    int enqueue_refine_passes() {
        if (int rc = fit_rows_kind(ctx, kind, v.xyz, v.nrm, ds->sample, 1, ds->model, ds->valid, one_row_nrm)) return rc;
        if (int rc = count_pass_kind(ctx, kind, v.xyz, n, ds->model, p.threshold, rb, &ds->mid)) return rc;
        if (seg) return write_pass<kPlane, true>(ctx, v.xyz, n, ds->model, p.threshold, rb, &ds->mid,
                                                 ctx->d_inl.as<unsigned long long>(), &ds->out, *seg);
        return write_pass_kind(ctx, kind, v.xyz, n, ds->model, p.threshold, rb, &ds->mid,
                               ctx->d_inl.as<unsigned long long>(), &ds->out);
    }

    /* host replay of ransac.h:572-613 over the wave's counts (rank-major gathered buffer `hc`) */
    int replay_wave(OrderedScan &scan, const ShardMap &sm, const uint32_t *hc, uint64_t done, uint32_t rows,
                    const uint32_t *h_table) {
        int rc_scan = 0;
        for (uint32_t r = 0; r < rows && !scan.stopped; ++r) {
            const uint32_t raw = hc[sm.gathered_index(r)]; /* rank-major buffer -> wave row */
            const bool valid = (raw & kInvalidBit) == 0;
            const uint64_t cnt = raw & ~kInvalidBit;
            const uint64_t before = scan.found ? scan.best_index : UINT64_MAX;
            scan.step(done + r, valid, cnt, [&](uint64_t j, bool exact, double *rmse) {
                const uint32_t rawj = (j >= done) ? hc[sm.gathered_index((uint32_t)(j - done))] : 0;
                const uint64_t cj = (j == scan.best_index && scan.found) ? scan.best_count : (uint64_t)(rawj & ~kInvalidBit);
                const int rc = eval_row(j, exact, cj, rmse);
                if (rc) rc_scan = rc;
                return rc;
            });
            if (rc_scan) return rc_scan;
            if (h_table && scan.found && scan.best_index != before) save_best(scan.best_index, h_table + (size_t)r * k);
        }
        return 0;
    }

    /* the wave's counts of all ranks -> ctx->h_counts (rank-major), synchronised */
    int fetch_counts(uint32_t S) {
        const uint32_t *d_all = ctx->d_counts.as<uint32_t>();
        if (R > 1) {
            M3D_CUDA(ctx, ctx->d_counts_all.reserve(sizeof(uint32_t) * (size_t)std::max<uint32_t>(S, 1) * R));
            if (int rc = exchange_allgather(ctx, ctx->d_counts.p, ctx->d_counts_all.p, sizeof(uint32_t) * (size_t)S)) return rc;
            d_all = ctx->d_counts_all.as<uint32_t>();
        }
        M3D_CUDA(ctx, ctx->h_counts.reserve(sizeof(uint32_t) * (size_t)std::max<uint32_t>(S, 1) * R));
        M3D_CUDA(ctx, cudaMemcpyAsync(ctx->h_counts.p, d_all, sizeof(uint32_t) * (size_t)S * R, cudaMemcpyDeviceToHost, ctx->stream));
        M3D_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        return 0;
    }

    int finish(OrderedScan &scan, bool refined_already) {
        scan.fill(&res->st);
        res->st.evaluated = evaluated;
        res->st.score_ms = score_ms;
        int ret = 0;
        if (scan.found) {
            const uint64_t bi = scan.best_index;
            if (!refined_already) {
                if (int rc = stage_row(bi)) return rc;
                if (int rc = enqueue_refine()) return rc;
                M3D_CUDA(ctx, cudaMemcpyAsync(hs, ds, sizeof(SmallDev), cudaMemcpyDeviceToHost, ctx->stream));
                M3D_CUDA(ctx, cudaEventRecord(ev_b, ctx->stream));
                M3D_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
            }
            if (hs->mid.n_inl != scan.best_count)
                return ctx->fail(M3D_ERR_INTERNAL, "best hypothesis %llu: scoring kernel counted %llu inliers, fp64 pass %llu",
                                 (unsigned long long)bi, (unsigned long long)scan.best_count, (unsigned long long)hs->mid.n_inl);
            res->n_inl = hs->mid.n_inl;
            memcpy(res->minimal, hs->model, sizeof res->minimal);
            const bool refit = (p.flags & M3D_FLAG_NO_REFIT) == 0;
            memcpy(res->refined, refit ? hs->out.model : hs->model, sizeof res->refined);
            res->st.refit_ok = refit ? hs->out.ok : 1;
            if (!(scan.best_rmse_known && scan.best_rmse_exact && scan.best_index == bi && scan.best_rmse != 0))
                res->st.best_rmse = hs->mid.err / std::sqrt((double)hs->mid.n_inl);
            res->st.exact_resolves = hs->resolves;
            ret = res->st.refit_ok;
        } else {
            if (!refined_already) {
                M3D_CUDA(ctx, cudaMemcpyAsync(&hs->resolves, &ds->resolves, sizeof(unsigned long long), cudaMemcpyDeviceToHost, ctx->stream));
                M3D_CUDA(ctx, cudaEventRecord(ev_b, ctx->stream));
                M3D_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
            }
            res->st.exact_resolves = hs->resolves;
        }
        float ms = 0;
        cudaEventElapsedTime(&ms, ev_a, ev_b);
        res->st.device_ms = ms;
        if (scan.found) {
            if (cudaEventElapsedTime(&ms, ev_r0, ev_r1) == cudaSuccess) res->st.refine_ms = ms;
            else cudaGetLastError();
        }
        if (drew_on_device) {
            if (cudaEventElapsedTime(&ms, ev_d0, ev_d1) == cudaSuccess) res->st.draw_ms = ms;
            else cudaGetLastError();
        }
        res->ret = ret;
        return M3D_OK;
    }

    /* ---- probability == 1 (no adaptive exit, ransac.h:601-606): the whole loop stays on the device.
     * Sample table drawn on the device (or on the host when the cylinder's normals live there), ONE scoring
     * launch per wave, wave_best_kernel + (R > 1) an all-gather of one 64-byte record per rank +
     * best_merge_kernel, and -- for a single wave -- RefineModel of the provisional winner enqueued behind
     * them, so that a fit costs one stream synchronisation.  Ties in the inlier count (the reference breaks
     * them with inlier_rmse) and a hypothesis with fitness 1 (which stops the reference's loop) are detected
     * in the record and replayed on the host exactly as before. */
    int run_on_device() {
        constexpr uint64_t kFastMaxRows = 1u << 22;
        if (H == 0 || H > kFastMaxRows || (host_nrm && H > (1u << 20))) return kRetryOnHost;
        const uint32_t rows_all = (uint32_t)H;
        OrderedScan scan(n, k, p.probability, H);
        M3D_CUDA(ctx, ctx->d_samples.reserve(sizeof(uint32_t) * (size_t)rows_all * k));
        const bool dev_draw = !host_nrm && device_draw_eligible(n, k, rows_all);
        std::vector<uint32_t> h_table;
        if (v.chunks && !dev_draw)
            if (int rc = v.chunks->issue_copies(ctx)) return rc; /* the host draw below overlaps the DMA (pinned buffers) */
        if (dev_draw) {
            M3D_CUDA(ctx, cudaEventRecord(ev_d0, ctx->stream));
            if (int rc = draw_table_device(ctx, p.seed, n, k, rows_all, ctx->d_samples.as<uint32_t>(), &ds->draw)) return rc;
            M3D_CUDA(ctx, cudaEventRecord(ev_d1, ctx->stream));
            drew_on_device = true;
        } else {
            M3D_CUDA(ctx, cudaMemsetAsync(&ds->draw, 0, sizeof(RowBreaks), ctx->stream));
            M3D_CUDA(ctx, ctx->h_samples.reserve(sizeof(uint32_t) * (size_t)rows_all * k));
            uint32_t *tab = ctx->h_samples.as<uint32_t>();
            SampleStream stream(p.seed, n);
            stream.draw_rows(k, rows_all, tab);
            M3D_CUDA(ctx, cudaMemcpyAsync(ctx->d_samples.p, tab, sizeof(uint32_t) * (size_t)rows_all * k,
                                          cudaMemcpyHostToDevice, ctx->stream));
            if (host_nrm) { /* normals of the sample points only: rows x k x 3 doubles */
                const size_t cnt = (size_t)rows_all * k;
                M3D_CUDA(ctx, ctx->h_rownrm.reserve(sizeof(double) * cnt * 3));
                M3D_CUDA(ctx, ctx->d_rownrm.reserve(sizeof(double) * cnt * 3));
                double *dst = ctx->h_rownrm.as<double>();
                for (size_t e = 0; e < cnt; ++e) {
                    const double *src = v.h_nrm + 3 * (size_t)tab[e];
                    dst[3 * e] = src[0], dst[3 * e + 1] = src[1], dst[3 * e + 2] = src[2];
                }
                M3D_CUDA(ctx, cudaMemcpyAsync(ctx->d_rownrm.p, dst, sizeof(double) * 3 * cnt, cudaMemcpyHostToDevice, ctx->stream));
            }
        }
        ChunkPlan *pl = v.chunks;
        if (pl)
            if (int rc = pl->issue_copies(ctx)) return rc; /* behind the draw kernels: the two overlap */
        if (pl && pl->pre_models) { /* all minimal models now, from the caller's pinned buffer (the upload is still in flight) */
            if (!dev_draw) return ctx->fail(M3D_ERR_INTERNAL, "chunked fit without a device-side sample draw");
            M3D_CUDA(ctx, ctx->d_models_all.reserve(sizeof(double) * 8 * (size_t)rows_all));
            M3D_CUDA(ctx, ctx->d_valid_all.reserve((size_t)rows_all));
            if (int rc = fit_rows_kind(ctx, kind, pl->dev_of_host_xyz, pl->dev_of_host_nrm, ctx->d_samples.as<uint32_t>(),
                                       rows_all, ctx->d_models_all.as<double>(), ctx->d_valid_all.as<uint8_t>()))
                return rc;
        }
        wave_base = 0, wave_rows = rows_all; /* the whole table is resident: stage_row never needs the host */
        M3D_CUDA(ctx, ctx->d_recs.reserve(sizeof(BestRec) * (size_t)(R + 1)));
        BestRec *d_local = ctx->d_recs.as<BestRec>(), *d_all = d_local + 1;

        const uint32_t cap = (uint32_t)std::min<uint64_t>((uint64_t)(1u << 18) * (uint64_t)std::max(R, 1), kFastMaxRows);
        uint64_t done = 0;
        bool refined = false;
        while (done < H && !scan.stopped) {
            const uint32_t rows = (uint32_t)std::min<uint64_t>(cap, H - done);
            const ShardMap sm{rows, (uint32_t)R, (uint32_t)rank};
            const uint32_t S = sm.padded(), mine = sm.local_rows();
            M3D_CUDA(ctx, ctx->d_counts.reserve(sizeof(uint32_t) * (size_t)std::max<uint32_t>(S, 1)));
            M3D_CUDA(ctx, cudaMemsetAsync(ctx->d_counts.p, 0, sizeof(uint32_t) * (size_t)std::max<uint32_t>(S, 1), ctx->stream));
            M3D_CUDA(ctx, cudaEventRecord(ev_s0, ctx->stream));
            g_trace.mark(1);
            if (mine && !pl) {
                ScoreArgs a = score_args(0, mine);
                a.samples = ctx->d_samples.as<uint32_t>() + (size_t)done * k;
                if (host_nrm) a.row_nrm = ctx->d_rownrm.as<double>() + (size_t)done * k * 3;
                if (int rc = launch_score_kind(ctx, kind, cloud_for_flags, a, exact_only)) return rc;
            }
            for (int c = 0; pl && c < pl->count; ++c) { /* one launch per chunk, behind the chunk's arrival */
                if (int rc = prep_chunk(ctx, v, *pl, c)) return rc;
                if (!mine) continue;
                ScoreArgs a = score_args(0, mine);
                a.samples = ctx->d_samples.as<uint32_t>() + (size_t)done * k;
                if (host_nrm) a.row_nrm = ctx->d_rownrm.as<double>() + (size_t)done * k * 3;
                if (pl->pre_models) {
                    a.models_in = ctx->d_models_all.as<double>() + (size_t)done * 8;
                    a.valid_in = ctx->d_valid_all.as<uint8_t>() + (size_t)done;
                    a.nrm = nullptr;
                }
                const uint32_t b = pl->begin[c];
                a.xyz = v.xyz + 3 * (size_t)b;
                a.pts32 = v.pts32 + b;
                a.blob = v.blob + (size_t)(b / kTile) * kBlobF4;
                a.perm = v.perm + b;
                a.meta = pl->metas + c;
                a.n = pl->begin[c + 1] - b;
                if (int rc = launch_score_kind(ctx, kind, cloud_for_flags, a, false)) return rc;
            }
            M3D_CUDA(ctx, cudaEventRecord(ev_s1, ctx->stream));
            g_trace.mark(2);
            wave_best_kernel<<<1, 1024, 0, ctx->stream>>>(ctx->d_counts.as<uint32_t>(), mine, (uint32_t)R, (uint32_t)rank, n,
                                                          &ds->draw, d_local);
            M3D_LAUNCHED(ctx);
            if (int rc = exchange_allgather(ctx, d_local, d_all, sizeof(BestRec))) return rc;
            best_merge_kernel<<<1, 32, 0, ctx->stream>>>(d_all, (uint32_t)R, ctx->d_samples.as<uint32_t>() + (size_t)done * k, k,
                                                         host_nrm ? ctx->d_rownrm.as<double>() + (size_t)done * k * 3 : nullptr,
                                                         &ds->best, ds->sample, ds->row_nrm);
            M3D_LAUNCHED(ctx);
            g_trace.mark(3);
            const bool speculate = (rows == H) && !seg;
            if (speculate)
                if (int rc = enqueue_refine()) return rc;
            M3D_CUDA(ctx, cudaMemcpyAsync(hs, ds, sizeof(SmallDev), cudaMemcpyDeviceToHost, ctx->stream));
            if (pl && done == 0)
                M3D_CUDA(ctx, cudaMemcpyAsync(ctx->h_metas.p, pl->metas, sizeof(CloudMeta) * pl->count, cudaMemcpyDeviceToHost, ctx->stream));
            if (speculate) M3D_CUDA(ctx, cudaEventRecord(ev_b, ctx->stream));
            g_trace.mark(4);
            M3D_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
            g_trace.mark(5);
            if (pl && done == 0) /* NaN / inf coordinates need the fp64 reference-order kernel: plain upload path */
                for (int c = 0; c < pl->count; ++c)
                    if (ctx->h_metas.as<CloudMeta>()[c].nonfinite) return kRetryUnchunked;
            float ms = 0;
            cudaEventElapsedTime(&ms, ev_s0, ev_s1);
            score_ms += ms;
            const BestRec rec = hs->best;
            if (rec.status) return kRetryOnHost; /* the device draw gave up: nothing of this fit is kept */
            evaluated += rows;
            bool clobbered = false;
            if (rec.first_full != kNoRow || rec.n_tied > (uint32_t)kTiedCap) {
                /* fitness 1 (the loop stops behind it) or more ties than the record lists: sequential replay */
                if (int rc = fetch_counts(S)) return rc;
                if (int rc = replay_wave(scan, sm, ctx->h_counts.as<uint32_t>(), done, rows, nullptr)) return rc;
                clobbered = true;
            } else if (rec.max_count == 0) {
                scan.count += rec.n_valid;
            } else {
                /* only rows that reach the wave's maximum can end up as the best (ransac.h:592-613); they are
                 * stepped in loop order, which breaks count ties with inlier_rmse exactly like the full replay */
                uint32_t fed = 0;
                int rc_scan = 0;
                for (uint32_t t = 0; t < rec.n_tied; ++t) {
                    scan.step(done + rec.tied[t], true, rec.max_count, [&](uint64_t j, bool exact, double *rmse) {
                        const uint64_t cj = (j == scan.best_index && scan.found) ? scan.best_count : (uint64_t)rec.max_count;
                        const int rc = eval_row(j, exact, cj, rmse);
                        clobbered = true;
                        if (rc) rc_scan = rc;
                        return rc;
                    });
                    if (rc_scan) return rc_scan;
                    ++fed;
                }
                scan.count += rec.n_valid - fed;
            }
            refined = speculate && !clobbered && scan.found && scan.best_index == done + rec.first_row;
            if (speculate && !scan.found && !clobbered) refined = true; /* nothing to refine; hs->resolves is current */
            done += rows;
        }
        return finish(scan, refined);
    }

    /* ---- probability < 1: hypotheses are scored in growing waves and the host replays the loop's
     * bookkeeping after each (adaptive early exit, ransac.h:601-610) */
    int run_on_host() {
        SampleStream stream(p.seed, n);
        std::vector<uint32_t> table; /* the current wave's rows (draw order) */
        OrderedScan scan(n, k, p.probability, H);
        uint64_t done = 0;
        /* a wave is at most kMaxWave rows PER RANK: one exchange + one host replay per wave, whatever the number of ranks */
        const uint32_t wave_cap = (uint32_t)std::min<uint64_t>((uint64_t)kMaxWave * (uint64_t)std::max(R, 1), 1u << 22);
        uint32_t wave = (p.probability >= 1.0) ? wave_cap : 256;
        while (done < H && !scan.stopped) {
            const uint32_t rows = (uint32_t)std::min<uint64_t>(wave, H - done);
            /* The sample rows come from one sequential mt19937 stream drawn on the host, and every rank needs the
             * whole table (its own rows for the GPU, any row for the replay's tie-breaks and the final refit).
             * Rows are dealt to the ranks in cyclic blocks (ShardMap), so a rank can launch the first part of its
             * shard after drawing a fraction of the table and draws the rest while its GPU works. */
            const ShardMap sm{rows, (uint32_t)R, (uint32_t)rank};
            const uint32_t S = sm.padded(), mine = sm.local_rows();
            table.resize((size_t)rows * k);
            wave_base = done, wave_rows = rows;
            M3D_CUDA(ctx, ctx->h_samples.reserve(sizeof(uint32_t) * (size_t)rows * k));
            M3D_CUDA(ctx, ctx->d_samples.reserve(sizeof(uint32_t) * (size_t)rows * k));
            if (host_nrm) { /* normals of this wave's sample points only: rows x k x 3 doubles */
                M3D_CUDA(ctx, ctx->h_rownrm.reserve(sizeof(double) * (size_t)rows * k * 3));
                M3D_CUDA(ctx, ctx->d_rownrm.reserve(sizeof(double) * (size_t)rows * k * 3));
            }
            uint32_t drawn = 0; /* rows of this wave drawn and uploaded so far */
            auto draw_to = [&](uint32_t upto) -> int {
                if (upto <= drawn) return 0;
                uint32_t *tab = &table[(size_t)drawn * k];
                stream.draw_rows(k, upto - drawn, tab);
                const size_t off = (size_t)drawn * k, cnt = (size_t)(upto - drawn) * k;
                memcpy(ctx->h_samples.as<uint32_t>() + off, tab, sizeof(uint32_t) * cnt);
                M3D_CUDA(ctx, cudaMemcpyAsync(ctx->d_samples.as<uint32_t>() + off, ctx->h_samples.as<uint32_t>() + off,
                                              sizeof(uint32_t) * cnt, cudaMemcpyHostToDevice, ctx->stream));
                if (host_nrm) {
                    double *dst = ctx->h_rownrm.as<double>() + 3 * off;
                    for (size_t e = 0; e < cnt; ++e) {
                        const double *src = v.h_nrm + 3 * (size_t)tab[e];
                        dst[3 * e] = src[0], dst[3 * e + 1] = src[1], dst[3 * e + 2] = src[2];
                    }
                    M3D_CUDA(ctx, cudaMemcpyAsync(ctx->d_rownrm.as<double>() + 3 * off, dst, sizeof(double) * 3 * cnt,
                                                  cudaMemcpyHostToDevice, ctx->stream));
                }
                drawn = upto;
                return 0;
            };
            M3D_CUDA(ctx, ctx->d_counts.reserve(sizeof(uint32_t) * (size_t)std::max<uint32_t>(S, 1)));
            M3D_CUDA(ctx, cudaMemsetAsync(ctx->d_counts.p, 0, sizeof(uint32_t) * (size_t)std::max<uint32_t>(S, 1), ctx->stream));
            bool s0_recorded = false;
            /* one launch on a single GPU (the draw of a 10k-row table is ~50 us there); with R ranks the table is R
             * times longer, so the shard is issued in up to four parts -- of at least classify_min_rows() rows each
             * when the shard is that large (so that every part is still pre-sorted into culled / dense hypotheses),
             * else in two halves */
            uint32_t parts = 1;
            if (R > 1) parts = (mine >= 2 * classify_min_rows()) ? std::min<uint32_t>(4, mine / classify_min_rows())
                                                                  : (mine >= 8192 ? 2 : 1);
            for (uint32_t part = 0; part < parts && mine; ++part) {
                const uint32_t l0 = (uint32_t)((uint64_t)mine * part / parts) / kShardBlock * kShardBlock;
                const uint32_t l1 = (part + 1 == parts) ? mine : (uint32_t)((uint64_t)mine * (part + 1) / parts) / kShardBlock * kShardBlock;
                if (l1 <= l0) continue;
                if (int rc = draw_to(std::min<uint32_t>(rows, sm.wave_row(l1 - 1) + 1))) return rc;
                if (!s0_recorded) { /* score_ms = the launches of the wave (the first part's table draw is not GPU time) */
                    M3D_CUDA(ctx, cudaEventRecord(ev_s0, ctx->stream));
                    s0_recorded = true;
                }
                if (int rc = launch_score_kind(ctx, kind, cloud_for_flags, score_args(l0, l1), exact_only)) return rc;
            }
            if (!s0_recorded) M3D_CUDA(ctx, cudaEventRecord(ev_s0, ctx->stream)); /* a rank without rows in this wave */
            M3D_CUDA(ctx, cudaEventRecord(ev_s1, ctx->stream));
            if (int rc = draw_to(rows)) return rc; /* the rest of the table (host side), while the GPU works */
            if (int rc = fetch_counts(S)) return rc;
            float ms = 0;
            cudaEventElapsedTime(&ms, ev_s0, ev_s1);
            score_ms += ms;
            evaluated += rows;
            if (int rc = replay_wave(scan, sm, ctx->h_counts.as<uint32_t>(), done, rows, table.data())) return rc;
            done += rows;
            if (wave < wave_cap) wave = std::min<uint32_t>(wave_cap, wave * 4);
        }
        /* (when the loop ran to max_iteration: the reference checks `count > current_iteration` only at the
         * top of an iteration, so nothing more to do) */
        return finish(scan, false);
    }
};

int fit_view(m3d_ctx *ctx, int kind, const m3d_cloud *cloud_for_flags, const CloudView &v,
             const m3d_ransac_params &p, SegArgs *seg, FitResult *res) {
    Fit f(ctx, kind, cloud_for_flags, v, p, seg, res);
    if (int rc = f.setup()) return rc;
    if (p.probability >= 1.0 && !loop_on_host()) {
        const int rc = f.run_on_device();
        if (rc != kRetryOnHost) return rc;
        if (v.chunks) return kRetryUnchunked; /* the host loop needs the whole cloud prepared */
        Fit g(ctx, kind, cloud_for_flags, v, p, seg, res); /* fresh state */
        if (int rc2 = g.setup()) return rc2;
        return g.run_on_host();
    }
    return f.run_on_host();
}

int check_params(m3d_ctx *ctx, int kind, size_t n, bool has_normals, const m3d_ransac_params *p) {
    if (!p) return ctx->fail(M3D_ERR_INVALID_ARG, "null params");
    if (kind < 0 || kind > 2) return ctx->fail(M3D_ERR_INVALID_ARG, "unknown primitive %d", kind);
    if (!(p->probability > 0 && p->probability <= 1))
        return ctx->fail(M3D_ERR_PROBABILITY, "Probability must be > 0 or <= 1.0"); /* ransac.h:483-485 */
    if (kind == kCylinder && !has_normals)
        return ctx->fail(M3D_ERR_NO_NORMALS, "Fit cylinder requires normals."); /* py_common.cpp:50-52 */
    if (n < (size_t)sample_size(kind))
        return ctx->fail(M3D_ERR_TOO_FEW_POINTS, "Can not fit model due to lack of points"); /* ransac.h:510-513 */
    return M3D_OK;
}

}  // namespace

extern "C" {

int m3d_cloud_upload(m3d_ctx *ctx, const double *xyz, const double *nrm, size_t n, m3d_cloud **out) {
    return cloud_create(ctx, xyz, nrm, n, cudaMemcpyHostToDevice, out);
}
int m3d_cloud_from_device(m3d_ctx *ctx, const double *d_xyz, const double *d_nrm, size_t n, m3d_cloud **out) {
    return cloud_create(ctx, d_xyz, d_nrm, n, cudaMemcpyDeviceToDevice, out);
}
void m3d_cloud_free(m3d_cloud *c) {
    if (!c) return;
    if (c->ctx) cudaStreamSynchronize(c->ctx->stream);
    c->xyz.release();
    c->nrm.release();
    c->pts32.release();
    c->meta.release();
    c->blob.release();
    c->perm.release();
    c->keys.release();
    c->hist.release();
    delete c;
}
size_t m3d_cloud_size(const m3d_cloud *c) { return c ? c->n : 0; }

} /* extern "C" */

namespace {
/* results of a finished fit -> the caller's buffers (model, ascending inlier indices) */
int deliver(m3d_ctx *ctx, int kind, const FitResult &res, double *model_out, size_t *inl_out, size_t *n_inl,
            m3d_ransac_stats *stats) {
    if (stats) *stats = res.st;
    if (res.st.found) {
        /* model = refined parameters, inliers = those of the minimal model (ransac.h:621-622) */
        for (int i = 0; i < param_count(kind); ++i) model_out[i] = res.refined[i];
        *n_inl = (size_t)res.n_inl;
        if (inl_out && res.n_inl) {
            static_assert(sizeof(size_t) == sizeof(unsigned long long), "size_t must be 64-bit");
            M3D_CUDA(ctx, cudaMemcpyAsync(inl_out, ctx->d_inl.p, sizeof(size_t) * (size_t)res.n_inl,
                                          cudaMemcpyDeviceToHost, ctx->stream));
            M3D_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        }
    }
    return res.ret;
}

/* device-mapped address of a pinned (page-locked, mapped) host buffer, or null for pageable memory */
const double *mapped_address(const double *host) {
    if (!host) return nullptr;
    cudaPointerAttributes at{};
    if (cudaPointerGetAttributes(&at, host) != cudaSuccess) {
        cudaGetLastError();
        return nullptr;
    }
    if (at.type != cudaMemoryTypeHost || !at.devicePointer) return nullptr;
    return static_cast<const double *>(at.devicePointer);
}

int e2e_chunks() { /* M3D_E2E_CHUNKS: 0 = plain upload (round-1 path); 1 (default) = one chunk: the upload runs on the
                    * copy stream while the sample table is drawn and the minimal models are solved from the pinned
                    * buffer; 2..8 = scoring chunk by chunk behind the upload */
    static const int v = getenv("M3D_E2E_CHUNKS") ? std::max(0, std::min(8, atoi(getenv("M3D_E2E_CHUNKS")))) : 1;
    return v;
}

/* The host-buffer fit with the upload overlapped (see ChunkPlan).  Returns kRetryUnchunked when the call does not
 * qualify or has to be redone through the plain upload path. */
int fit_host_chunked(m3d_ctx *ctx, int kind, const double *xyz, const double *nrm, size_t n, const m3d_ransac_params *p,
                     double *model_out, size_t *inl_out, size_t *n_inl, m3d_ransac_stats *stats) {
    const int k = sample_size(kind);
    const int want = (p->flags & M3D_FLAG_CHUNKED_UPLOAD) ? std::max(4, e2e_chunks()) : e2e_chunks();
    if (want < 1 || (p->flags & M3D_FLAG_PLAIN_UPLOAD) || p->probability < 1.0 || loop_on_host() || score_path() != 2 ||
        n < (size_t)want * 64 * kTile ||
        p->max_iteration == 0 || p->max_iteration > (1u << 18) * (uint64_t)std::max(ctx->world, 1) ||
        (p->flags & (M3D_FLAG_EXACT_ONLY | M3D_FLAG_DENSE | M3D_FLAG_CLASSIFY)) || !(p->threshold > 0) ||
        (want >= 2 && !device_draw_eligible((uint32_t)n, k, p->max_iteration)))
        return kRetryUnchunked;
    const bool multi = want >= 2; /* scoring chunk by chunk needs the zero-copy model solve, i.e. pinned buffers */
    const double *dxyz = multi ? mapped_address(xyz) : nullptr;
    const double *dnrm = (multi && kind == kCylinder) ? mapped_address(nrm) : nullptr;
    if (multi && (!dxyz || (kind == kCylinder && !dnrm))) return kRetryUnchunked; /* pageable memory: plain path */
    if (!multi && kind != kCylinder && !device_draw_eligible((uint32_t)n, k, p->max_iteration)) return kRetryUnchunked;
    if (!ctx->copy_stream) {
        M3D_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
        for (auto &e : ctx->ev_chunk) M3D_CUDA(ctx, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    }
    if (!ctx->scratch_cloud) ctx->scratch_cloud = new m3d_cloud();
    m3d_cloud *c = ctx->scratch_cloud;
    c->ctx = ctx;
    c->n = n;
    c->sorted = false;
    c->has_normals = false;
    c->h_nrm = nullptr;
    c->h_meta = CloudMeta{};
    const uint32_t ntiles = (uint32_t)((n + kTile - 1) / kTile);
    ChunkPlan pl;
    pl.count = want;
    const uint32_t per = (ntiles + want - 1) / want;
    for (int i = 0; i <= want; ++i) pl.begin[i] = (uint32_t)std::min<size_t>(n, (size_t)per * i * kTile);
    M3D_CUDA(ctx, c->xyz.reserve(sizeof(double) * (3 * n + (size_t)std::max(ctx->world, 1)))); /* + all-gather padding */
    M3D_CUDA(ctx, c->pts32.reserve(sizeof(float4) * n));
    M3D_CUDA(ctx, c->blob.reserve(sizeof(float4) * ((size_t)ntiles + want) * kBlobF4));
    M3D_CUDA(ctx, c->perm.reserve(sizeof(uint32_t) * ((size_t)ntiles + want) * kTile));
    M3D_CUDA(ctx, c->keys.reserve(sizeof(uint32_t) * ((size_t)per * kTile)));
    M3D_CUDA(ctx, c->hist.reserve(sizeof(uint32_t) * morton_hist_words(morton_bits(per * kTile))));
    M3D_CUDA(ctx, ctx->d_metas.reserve(sizeof(CloudMeta) * 8));
    M3D_CUDA(ctx, ctx->h_metas.reserve(sizeof(CloudMeta) * 8));
    pl.metas = ctx->d_metas.as<CloudMeta>();
    pl.keys = c->keys.as<uint32_t>();
    pl.hist = c->hist.as<uint32_t>();
    pl.dev_of_host_xyz = dxyz;
    pl.dev_of_host_nrm = dnrm;
    pl.pre_models = multi;
    pl.h_xyz = xyz;
    pl.d_xyz = c->xyz.as<double>();
    static const bool no_shard = env_is("M3D_SHARD_UPLOAD", "0");
    pl.shard_upload = !multi && !no_shard && exchange_has_nccl(ctx) && n >= ((size_t)1 << 16);
    CloudView v{c->xyz.as<double>(), dnrm, c->pts32.as<float4>(), pl.metas, (uint32_t)n, false};
    if (!multi) v.h_nrm = nrm; /* single chunk: the cylinder's sample normals are gathered on the host, behind the DMA */
    v.blob = c->blob.as<float4>();
    v.perm = c->perm.as<uint32_t>();
    v.chunks = &pl;
    FitResult res;
    const int rc = fit_view(ctx, kind, c, v, *p, nullptr, &res);
    if (rc != M3D_OK) {
        if (!pl.copies_issued && rc == kRetryUnchunked) return rc;
        cudaStreamSynchronize(ctx->copy_stream); /* nothing of this call may still be writing the staging buffers */
        cudaStreamSynchronize(ctx->stream);
        return rc;
    }
    return deliver(ctx, kind, res, model_out, inl_out, n_inl, stats);
}
}  // namespace

extern "C" {

int m3d_ransac_fit_cloud(m3d_ctx *ctx, int kind, const m3d_cloud *cloud, const m3d_ransac_params *p,
                         double *model_out, size_t *inl_out, size_t *n_inl, m3d_ransac_stats *stats) {
    if (!ctx || !cloud || !model_out || !n_inl) return M3D_ERR_INVALID_ARG;
    for (int i = 0; i < 8; ++i) model_out[i] = 0;
    *n_inl = 0;
    if (stats) memset(stats, 0, sizeof *stats);
    if (int rc = check_params(ctx, kind, cloud->n, cloud->has_normals, p)) return rc;
    M3D_CUDA(ctx, cudaSetDevice(ctx->device));
    g_trace.armed = true;
    g_trace.mark(0);
    CloudView v{cloud->xyz.as<double>(), (cloud->has_normals && !cloud->h_nrm) ? cloud->nrm.as<double>() : nullptr,
                cloud->pts32.as<float4>(), cloud->meta.as<CloudMeta>(), (uint32_t)cloud->n,
                cloud->h_meta.nonfinite != 0};
    v.h_nrm = cloud->h_nrm;
    if (int rc = ensure_sorted(ctx, cloud)) return rc;
    if (cloud->sorted) {
        v.blob = cloud->blob.as<float4>();
        v.perm = cloud->perm.as<uint32_t>();
    }
    FitResult res;
    if (int rc = fit_view(ctx, kind, cloud, v, *p, nullptr, &res)) return rc;
    g_trace.mark(6);
    const int ret = deliver(ctx, kind, res, model_out, inl_out, n_inl, stats);
    g_trace.mark(7);
    g_trace.armed = false;
    if (g_trace.skip > 0) --g_trace.skip;
    else ++g_trace.calls;
    return ret;
}

int m3d_ransac_fit(m3d_ctx *ctx, int kind, const double *xyz, const double *nrm, size_t n,
                   const m3d_ransac_params *p, double *model_out, size_t *inl_out, size_t *n_inl,
                   m3d_ransac_stats *stats) {
    if (!ctx || !model_out || !n_inl) return M3D_ERR_INVALID_ARG;
    for (int i = 0; i < 8; ++i) model_out[i] = 0;
    *n_inl = 0;
    if (stats) memset(stats, 0, sizeof *stats);
    if (int rc = check_params(ctx, kind, n, nrm != nullptr, p)) return rc;
    if (n >= (size_t)kInvalidBit) return ctx->fail(M3D_ERR_INVALID_ARG, "clouds of >= 2^31 points are not supported");
    M3D_CUDA(ctx, cudaSetDevice(ctx->device));
    if ((p->flags & M3D_FLAG_REGISTER_HOST) && n) { /* page-lock the caller's cloud in place (once per buffer) */
        cudaPointerAttributes at{};
        if (cudaPointerGetAttributes(&at, xyz) != cudaSuccess) cudaGetLastError();
        else if (at.type == cudaMemoryTypeUnregistered) {
            if (cudaHostRegister(const_cast<double *>(xyz), sizeof(double) * 3 * n, cudaHostRegisterDefault) == cudaSuccess)
                ctx->registered.emplace_back(xyz, sizeof(double) * 3 * n);
            else
                cudaGetLastError(); /* e.g. a read-only mapping: the staged upload still works */
        }
    }
    {
        const int rc = fit_host_chunked(ctx, kind, xyz, nrm, n, p, model_out, inl_out, n_inl, stats);
        if (rc != kRetryUnchunked) return rc;
        for (int i = 0; i < 8; ++i) model_out[i] = 0;
        *n_inl = 0;
        if (stats) memset(stats, 0, sizeof *stats);
    }
    /* the staging cloud lives in the context: repeated calls re-use its device buffers */
    if (!ctx->scratch_cloud) ctx->scratch_cloud = new m3d_cloud();
    /* the normals stay on the host: the fit reads them at the sample points only (cylinder MinimalFit,
     * ransac.h:376-383), so rows x 2 normals per wave are uploaded instead of all n */
    if (int rc = cloud_fill(ctx, ctx->scratch_cloud, xyz, nullptr, n, cudaMemcpyHostToDevice)) return rc;
    ctx->scratch_cloud->has_normals = nrm != nullptr;
    ctx->scratch_cloud->h_nrm = nrm;
    const int rc = m3d_ransac_fit_cloud(ctx, kind, ctx->scratch_cloud, p, model_out, inl_out, n_inl, stats);
    ctx->scratch_cloud->h_nrm = nullptr; /* borrowed for this call only */
    ctx->scratch_cloud->has_normals = false;
    return rc;
}

/* utils.h:81-97 drawn on the device (loop_kernels.cuh): rows x k indices, bit-identical to m3d_sample_table.
 * returns 1 = drawn on the device, 0 = not eligible / gave up (too many duplicate draws: tiny clouds) */
int m3d_sample_table_device(m3d_ctx *ctx, uint32_t seed, size_t n, int k, size_t rows, uint32_t *out) {
    if (!ctx || !out || k < 1) return M3D_ERR_INVALID_ARG;
    if (rows == 0) return 1;
    if (n >= (size_t)kInvalidBit || rows > (1u << 22) || !device_draw_eligible((uint32_t)n, k, rows)) return 0;
    M3D_CUDA(ctx, cudaSetDevice(ctx->device));
    M3D_CUDA(ctx, ctx->d_small.reserve(sizeof(SmallDev)));
    M3D_CUDA(ctx, ctx->h_small.reserve(sizeof(SmallDev)));
    SmallDev *ds = ctx->d_small.as<SmallDev>();
    SmallDev *hs = ctx->h_small.as<SmallDev>();
    M3D_CUDA(ctx, ctx->d_samples.reserve(sizeof(uint32_t) * rows * k));
    if (int rc = draw_table_device(ctx, seed, (uint32_t)n, k, (uint32_t)rows, ctx->d_samples.as<uint32_t>(), &ds->draw)) return rc;
    M3D_CUDA(ctx, cudaMemcpyAsync(&hs->draw, &ds->draw, sizeof(RowBreaks), cudaMemcpyDeviceToHost, ctx->stream));
    M3D_CUDA(ctx, cudaMemcpyAsync(out, ctx->d_samples.p, sizeof(uint32_t) * rows * k, cudaMemcpyDeviceToHost, ctx->stream));
    M3D_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return hs->draw.status ? 0 : 1;
}

/* counters of the statistics build of score_cell_kernel (M3D_FLAG_STATS), read and cleared:
 * [0] (hypothesis, tile) tests, [1] survivors, [2] (hypothesis, cell) list entries (x32 = point-hypothesis
 * pairs evaluated), [3] one-per-lane passes, [4] two-per-lane passes, [5] guard-band re-scans */
int m3d_score_stats(m3d_ctx *ctx, uint64_t out[8]) {
    if (!ctx || !out) return M3D_ERR_INVALID_ARG;
    unsigned long long z[8] = {0, 0, 0, 0, 0, 0, 0, 0}, v[8];
    M3D_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    M3D_CUDA(ctx, cudaMemcpyFromSymbol(v, g_cell_stats, sizeof v));
    M3D_CUDA(ctx, cudaMemcpyToSymbol(g_cell_stats, z, sizeof z));
    for (int i = 0; i < 8; ++i) out[i] = v[i];
    return M3D_OK;
}

#ifdef M3D_CULL_STATS
int m3d_debug_cull_stats(unsigned long long *out) { /* reads and clears the counters of score_cull_kernel */
    unsigned long long z[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    cudaDeviceSynchronize();
    cudaMemcpyFromSymbol(out, g_cull_stats, sizeof z);
    cudaMemcpyToSymbol(g_cull_stats, z, sizeof z);
    return 0;
}
#endif

int m3d_score_samples(m3d_ctx *ctx, int kind, const m3d_cloud *cloud, const uint32_t *samples, size_t rows,
                      double threshold, uint32_t flags, double *models, uint8_t *valid, uint64_t *counts) {
    if (!ctx || !cloud || !samples || !counts) return M3D_ERR_INVALID_ARG;
    if (kind < 0 || kind > 2) return ctx->fail(M3D_ERR_INVALID_ARG, "unknown primitive %d", kind);
    if (kind == kCylinder && !cloud->has_normals) return ctx->fail(M3D_ERR_NO_NORMALS, "Fit cylinder requires normals.");
    if (cloud->n < (size_t)sample_size(kind)) return ctx->fail(M3D_ERR_TOO_FEW_POINTS, "Can not fit model due to lack of points");
    if (rows == 0) return M3D_OK;
    if (rows >= (1u << 30)) return ctx->fail(M3D_ERR_INVALID_ARG, "too many rows");
    M3D_CUDA(ctx, cudaSetDevice(ctx->device));
    const int k = sample_size(kind);
    M3D_CUDA(ctx, ctx->d_samples.reserve(sizeof(uint32_t) * rows * k));
    M3D_CUDA(ctx, ctx->d_counts.reserve(sizeof(uint32_t) * rows));
    M3D_CUDA(ctx, ctx->h_counts.reserve(sizeof(uint32_t) * rows));
    M3D_CUDA(ctx, ctx->d_models.reserve(sizeof(double) * 8 * rows));
    M3D_CUDA(ctx, ctx->d_valid.reserve(rows));
    M3D_CUDA(ctx, ctx->d_small.reserve(sizeof(SmallDev)));
    SmallDev *ds = ctx->d_small.as<SmallDev>();
    M3D_CUDA(ctx, cudaMemcpyAsync(ctx->d_samples.p, samples, sizeof(uint32_t) * rows * k, cudaMemcpyHostToDevice, ctx->stream));
    M3D_CUDA(ctx, cudaMemsetAsync(ctx->d_counts.p, 0, sizeof(uint32_t) * rows, ctx->stream));
    M3D_CUDA(ctx, cudaMemsetAsync(&ds->resolves, 0, sizeof(unsigned long long), ctx->stream));
    if (int rc = ensure_sorted(ctx, cloud)) return rc;
    ScoreArgs a{};
    if (cloud->sorted && !(flags & M3D_FLAG_DENSE)) {
        a.blob = cloud->blob.as<float4>();
        a.perm = cloud->perm.as<uint32_t>();
    }
    a.flags = flags;
    a.pts32 = cloud->pts32.as<float4>();
    a.xyz = cloud->xyz.as<double>();
    a.nrm = cloud->has_normals ? cloud->nrm.as<double>() : nullptr;
    a.meta = cloud->meta.as<CloudMeta>();
    a.samples = ctx->d_samples.as<uint32_t>();
    a.counts = ctx->d_counts.as<uint32_t>();
    a.resolves = &ds->resolves;
    a.thr = threshold;
    a.n = (uint32_t)cloud->n;
    a.row_begin = 0;
    a.rows = (uint32_t)rows;
    if (int rc = launch_score_kind(ctx, kind, cloud, a, (flags & M3D_FLAG_EXACT_ONLY) != 0)) return rc;
    if (models || valid) {
        if (int rc = fit_rows_kind(ctx, kind, a.xyz, a.nrm, a.samples, a.rows, ctx->d_models.as<double>(),
                                   ctx->d_valid.as<uint8_t>()))
            return rc;
        if (models)
            M3D_CUDA(ctx, cudaMemcpyAsync(models, ctx->d_models.p, sizeof(double) * 8 * rows, cudaMemcpyDeviceToHost, ctx->stream));
        if (valid) M3D_CUDA(ctx, cudaMemcpyAsync(valid, ctx->d_valid.p, rows, cudaMemcpyDeviceToHost, ctx->stream));
    }
    M3D_CUDA(ctx, cudaMemcpyAsync(ctx->h_counts.p, ctx->d_counts.p, sizeof(uint32_t) * rows, cudaMemcpyDeviceToHost, ctx->stream));
    M3D_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    const uint32_t *hc = ctx->h_counts.as<uint32_t>();
    for (size_t r = 0; r < rows; ++r) counts[r] = hc[r] & ~kInvalidBit;
    return M3D_OK;
}

int m3d_evaluate_model(m3d_ctx *ctx, int kind, const m3d_cloud *cloud, const double *model, double threshold,
                       int sequential, uint64_t *count, double *err) {
    if (!ctx || !cloud || !model || !count) return M3D_ERR_INVALID_ARG;
    if (kind < 0 || kind > 2) return ctx->fail(M3D_ERR_INVALID_ARG, "unknown primitive %d", kind);
    M3D_CUDA(ctx, cudaSetDevice(ctx->device));
    const uint32_t n = (uint32_t)cloud->n;
    M3D_CUDA(ctx, ctx->d_small.reserve(sizeof(SmallDev)));
    M3D_CUDA(ctx, ctx->h_small.reserve(sizeof(SmallDev)));
    SmallDev *ds = ctx->d_small.as<SmallDev>();
    SmallDev *hs = ctx->h_small.as<SmallDev>();
    RefineBufs rb;
    if (int rc = refine_bufs(ctx, n, &rb)) return rc;
    M3D_CUDA(ctx, ctx->d_inl.reserve(sizeof(unsigned long long) * (size_t)std::max<uint32_t>(n, 1)));
    double m8[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int i = 0; i < param_count(kind); ++i) m8[i] = model[i];
    M3D_CUDA(ctx, cudaMemcpyAsync(ds->model, m8, sizeof m8, cudaMemcpyHostToDevice, ctx->stream));
    if (int rc = count_pass_kind(ctx, kind, cloud->xyz.as<double>(), n, ds->model, threshold, rb, &ds->mid)) return rc;
    if (sequential) {
        if (int rc = write_pass_kind(ctx, kind, cloud->xyz.as<double>(), n, ds->model, threshold, rb, &ds->mid,
                                     ctx->d_inl.as<unsigned long long>(), &ds->out))
            return rc;
        M3D_CUDA(ctx, cudaMemcpyAsync(hs, ds, sizeof(SmallDev), cudaMemcpyDeviceToHost, ctx->stream));
        M3D_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        if (int rc = seq_err_kind(ctx, kind, cloud->xyz.as<double>(), ctx->d_inl.as<unsigned long long>(),
                                  hs->mid.n_inl, ds->model, &ds->seq_err))
            return rc;
    }
    M3D_CUDA(ctx, cudaMemcpyAsync(hs, ds, sizeof(SmallDev), cudaMemcpyDeviceToHost, ctx->stream));
    M3D_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    *count = hs->mid.n_inl;
    if (err) *err = sequential ? hs->seq_err : hs->mid.err;
    return M3D_OK;
}

} /* extern "C" */

namespace {
/* SegmentPlaneIterative, src/iterative_plane_segmentation.cpp:7-39.  Labels are 32-bit on the device; `wide`
 * selects the size_t labels of the original entry point (widened on the device before the copy). */
int segment_impl(m3d_ctx *ctx, const double *xyz, size_t n, double threshold, int max_iteration, double min_ratio,
                 uint32_t seed, double *planes, size_t cap_planes, void *labels, bool wide, size_t *n_planes,
                 float *device_ms) {
    if (!ctx || !n_planes || (n && (!xyz || !labels)) || (cap_planes && !planes)) return M3D_ERR_INVALID_ARG;
    *n_planes = 0;
    if (device_ms) *device_ms = 0;
    if (n < 3) { /* :14-17 warning + empty result */
        for (size_t i = 0; i < n; ++i) {
            if (wide) static_cast<uint64_t *>(labels)[i] = UINT64_MAX;
            else static_cast<uint32_t *>(labels)[i] = UINT32_MAX;
        }
        return M3D_OK;
    }
    if (n >= (size_t)kInvalidBit) return ctx->fail(M3D_ERR_INVALID_ARG, "clouds of >= 2^31 points are not supported");
    if (max_iteration < 0) max_iteration = 0;
    M3D_CUDA(ctx, cudaSetDevice(ctx->device));
    if (!ctx->scratch_cloud) ctx->scratch_cloud = new m3d_cloud();
    m3d_cloud *c = ctx->scratch_cloud; /* staging cloud, buffers re-used across calls */
    if (int rc = cloud_fill(ctx, c, xyz, nullptr, n, cudaMemcpyHostToDevice)) return rc;

    /* ping-pong buffers for the shrinking cloud + original indices + labels */
    const size_t n3 = sizeof(double) * 3 * n;
    M3D_CUDA(ctx, ctx->d_tmp0.reserve(n3));                                  /* xyz B      */
    M3D_CUDA(ctx, ctx->d_tmp1.reserve(sizeof(float4) * n));                  /* pts32 B    */
    M3D_CUDA(ctx, ctx->d_tmp2.reserve(sizeof(uint32_t) * n));                /* orig A     */
    M3D_CUDA(ctx, ctx->d_tmp3.reserve(sizeof(uint32_t) * n));                /* orig B     */
    M3D_CUDA(ctx, ctx->d_tmp4.reserve(sizeof(uint32_t) * n));                /* labels     */
    M3D_CUDA(ctx, cudaMemsetAsync(ctx->d_tmp4.p, 0xff, sizeof(uint32_t) * n, ctx->stream));
    iota_kernel<<<ctx->sm_count * 4, 256, 0, ctx->stream>>>(ctx->d_tmp2.as<uint32_t>(), (uint32_t)n);
    M3D_LAUNCHED(ctx);
    double *xyz_cur = c->xyz.as<double>(), *xyz_nxt = ctx->d_tmp0.as<double>();
    float4 *p32_cur = c->pts32.as<float4>(), *p32_nxt = ctx->d_tmp1.as<float4>();
    uint32_t *org_cur = ctx->d_tmp2.as<uint32_t>(), *org_nxt = ctx->d_tmp3.as<uint32_t>();

    size_t count = 0, remaining = n;
    const size_t target = (size_t)((1 - min_ratio) * (double)n); /* :28 */
    uint32_t round = 0;
    float total_ms = 0;
    int status = M3D_OK;
    while (count < target) {
        if (remaining < 3) { /* ransac.h:510-513 throws out of the loop */
            status = ctx->fail(M3D_ERR_TOO_FEW_POINTS, "Can not fit model due to lack of points");
            break;
        }
        if (*n_planes >= cap_planes) {
            status = ctx->fail(M3D_ERR_CAPACITY, "more than %zu planes", cap_planes);
            break;
        }
        m3d_ransac_params p{};
        p.threshold = threshold;
        p.max_iteration = (uint64_t)max_iteration;
        p.probability = 0.9999; /* estimator default, ransac.h:462 (SetProbability is not called, :20-21) */
        p.seed = seed + round;
        CloudView v{xyz_cur, nullptr, p32_cur, c->meta.as<CloudMeta>(), (uint32_t)remaining, c->h_meta.nonfinite != 0};
        SegArgs seg{p32_cur, org_cur, xyz_nxt, p32_nxt, org_nxt, ctx->d_tmp4.as<uint32_t>(), (uint32_t)*n_planes};
        FitResult res;
        if (int rc = fit_view(ctx, kPlane, c, v, p, &seg, &res)) return rc;
        total_ms += res.st.device_ms;
        if (!res.st.found || res.n_inl == 0) { /* the reference would loop forever (Appendix A.12) */
            status = ctx->fail(M3D_ERR_NO_INLIERS, "segmentation round %u found no inliers", round);
            break;
        }
        for (int i = 0; i < 4; ++i) planes[4 * (*n_planes) + i] = res.refined[i];
        (*n_planes)++;
        count += (size_t)res.n_inl;
        remaining -= (size_t)res.n_inl;
        std::swap(xyz_cur, xyz_nxt);
        std::swap(p32_cur, p32_nxt);
        std::swap(org_cur, org_nxt);
        round++;
    }
    if (wide) {
        M3D_CUDA(ctx, ctx->d_inl.reserve(sizeof(unsigned long long) * n));
        widen_labels_kernel<<<ctx->sm_count * 4, 256, 0, ctx->stream>>>(ctx->d_tmp4.as<uint32_t>(),
                                                                        ctx->d_inl.as<unsigned long long>(), (uint32_t)n);
        M3D_LAUNCHED(ctx);
        M3D_CUDA(ctx, cudaMemcpyAsync(labels, ctx->d_inl.p, sizeof(uint64_t) * n, cudaMemcpyDeviceToHost, ctx->stream));
    } else {
        M3D_CUDA(ctx, cudaMemcpyAsync(labels, ctx->d_tmp4.p, sizeof(uint32_t) * n, cudaMemcpyDeviceToHost, ctx->stream));
    }
    M3D_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (device_ms) *device_ms = total_ms;
    return status;
}
}  // namespace

extern "C" {

int m3d_segment_plane_iterative(m3d_ctx *ctx, const double *xyz, size_t n, double threshold, int max_iteration,
                                double min_ratio, uint32_t seed, double *planes, size_t cap_planes,
                                uint64_t *labels, size_t *n_planes, float *device_ms) {
    return segment_impl(ctx, xyz, n, threshold, max_iteration, min_ratio, seed, planes, cap_planes, labels, true, n_planes,
                        device_ms);
}
int m3d_segment_plane_iterative_u32(m3d_ctx *ctx, const double *xyz, size_t n, double threshold, int max_iteration,
                                    double min_ratio, uint32_t seed, double *planes, size_t cap_planes,
                                    uint32_t *labels, size_t *n_planes, float *device_ms) {
    return segment_impl(ctx, xyz, n, threshold, max_iteration, min_ratio, seed, planes, cap_planes, labels, false, n_planes,
                        device_ms);
}


} /* extern "C" */
