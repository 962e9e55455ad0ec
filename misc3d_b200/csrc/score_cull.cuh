/*
 * score_cull.cuh -- the hierarchical (culling) form of the all-point inlier scoring
 * (reference: RANSAC<>::EvaluateModel, include/misc3d/common/ransac.h:626-654, once per hypothesis).
 *
 * EvaluateModel visits every point for every hypothesis.  Its result -- the inlier COUNT -- does
 * not depend on the order of the points, and a primitive's inliers live in a thin shell
 * (|distance| < threshold) that misses almost all of space.  So the cloud is kept a second time in
 * Morton order, cut into cells of 32 points and tiles of 32 cells, each with a bounding sphere
 * {centre, R}.  All three distance functions are 1-Lipschitz, hence
 *        |dist(centre)| - R - margin >= threshold   =>   no point of the cell/tile is an inlier
 * and the whole cell (32 point-hypothesis pairs) or tile (1024) is skipped with ONE evaluation.
 * Cells that survive are evaluated point by point exactly like score_kernel does (same fp32 guard
 * band, same fp64 resolve queue), so the counts are bit-identical to the reference's: culling only
 * ever removes pairs that are provably outliers (margins below cover every rounding involved).
 *
 *   preparation  morton_hist / scan_* / morton_scatter / tile_bounds   (counting sort by a 128^3
 *                Morton grid; once per uploaded cloud, ~5 short kernels)
 *   hot kernel   score_cull_kernel<KIND,THREADS,HPT>: persistent CTAs, producer warp + TMA ring as
 *                in score_kernel (one 16.9 KB bulk copy per tile brings 1024 points + 32 cell
 *                spheres + the tile sphere);  lane = hypothesis for the tile test, lane = cell for
 *                the cell test, lane = point inside surviving cells; counts reduced with redux.sync.
 */
#pragma once
#include "ransac_kernels.cuh"

namespace m3d {

constexpr int kCellPts = 32;                 /* points per cell = one warp                          */
constexpr int kTileCells = kTile / kCellPts; /* 32 cells per tile                                   */
constexpr int kBlobF4 = kTile + kTileCells + 1; /* float4 per tile in HBM: points, cell spheres, tile sphere */
constexpr int kStageF4 = kBlobF4 + kCellPts;    /* + one all-NaN dummy cell per smem stage (never TMA-written) */
#ifndef M3D_CULL_STAGES
#define M3D_CULL_STAGES 4
#endif
constexpr int kCullStages = M3D_CULL_STAGES; /* ring depth: slack between the fastest and the slowest warp of a CTA */
constexpr int kGridBits = 7;                 /* finest Morton grid: 128^3 (clouds of >= ~1M points)   */
constexpr uint32_t kBins = 1u << (3 * kGridBits);
/* grid resolution for n points: about one bin per point, 32^3 .. 128^3 (the cost of the counting sort's scans is
 * proportional to the number of bins, which matters when a cloud is sorted chunk by chunk) */
inline int morton_bits(uint32_t n) {
    int b = 5;
    while (b < kGridBits && (1ull << (3 * b)) < (unsigned long long)n) ++b;
    return b;
}
constexpr int kScanBlock = 1024, kScanItems = 2; /* scan: 2048 bins per block                        */

__device__ __forceinline__ uint32_t part1by2(uint32_t x) { /* 7 bits -> every third bit */
    x &= 0x3ffu;
    x = (x | (x << 16)) & 0x30000ffu;
    x = (x | (x << 8)) & 0x300f00fu;
    x = (x | (x << 4)) & 0x30c30c3u;
    x = (x | (x << 2)) & 0x9249249u;
    return x;
}

/* key = Morton code of the point's grid cell inside the cube [-mc, mc]^3 around the bbox centre */
__global__ void __launch_bounds__(256) morton_hist_kernel(const float4 *__restrict__ pts32, uint32_t n,
                                                          const CloudMeta *__restrict__ meta, int bits,
                                                          uint32_t *__restrict__ keys, uint32_t *__restrict__ hist) {
    const float mc = (float)meta->mc;
    const float scale = mc > 0.f ? (float)(1 << (bits - 1)) / mc : 0.f;
    const int gmax = (1 << bits) - 1;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const float4 p = pts32[i];
        const int gx = min(gmax, max(0, (int)((p.x + mc) * scale)));
        const int gy = min(gmax, max(0, (int)((p.y + mc) * scale)));
        const int gz = min(gmax, max(0, (int)((p.z + mc) * scale)));
        const uint32_t key = part1by2(gx) | (part1by2(gy) << 1) | (part1by2(gz) << 2);
        keys[i] = key;
        atomicAdd(&hist[key], 1u);
    }
}

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
        wsum[lane] = s; /* inclusive over warps */
    }
    __syncthreads();
    const uint32_t before = w ? wsum[w - 1] : 0u;
    if (total) *total = wsum[31];
    return before + inc - v;
}

/* pass 1: sum of each 2048-bin block */
__global__ void __launch_bounds__(kScanBlock) scan_sums_kernel(const uint32_t *__restrict__ hist, uint32_t *__restrict__ bsum) {
    const uint2 v = reinterpret_cast<const uint2 *>(hist)[blockIdx.x * kScanBlock + threadIdx.x];
    uint32_t tot;
    block_exclusive_scan_1024(v.x + v.y, &tot);
    if (threadIdx.x == 0) bsum[blockIdx.x] = tot;
}
/* pass 2: exclusive scan of the (<= 1024) block sums */
__global__ void __launch_bounds__(kScanBlock) scan_top_kernel(uint32_t *bsum, int nblocks) {
    const uint32_t v = (int)threadIdx.x < nblocks ? bsum[threadIdx.x] : 0u;
    const uint32_t ex = block_exclusive_scan_1024(v, nullptr);
    if ((int)threadIdx.x < nblocks) bsum[threadIdx.x] = ex;
}
/* pass 3: hist[b] <- number of points with a smaller key (the scatter cursor of bin b) */
__global__ void __launch_bounds__(kScanBlock) scan_apply_kernel(uint32_t *hist, const uint32_t *__restrict__ bsum) {
    uint2 *h2 = reinterpret_cast<uint2 *>(hist);
    const uint2 v = h2[blockIdx.x * kScanBlock + threadIdx.x];
    const uint32_t ex = block_exclusive_scan_1024(v.x + v.y, nullptr) + bsum[blockIdx.x];
    h2[blockIdx.x * kScanBlock + threadIdx.x] = make_uint2(ex, ex + v.x);
}

/* sorted position -> tile blob; the order inside one grid cell is whatever the atomics give
 * (inlier counts do not depend on it) */
__global__ void __launch_bounds__(256) morton_scatter_kernel(const float4 *__restrict__ pts32, uint32_t n,
                                                             const uint32_t *__restrict__ keys, uint32_t *cursor,
                                                             float4 *__restrict__ blob, uint32_t *__restrict__ perm) {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const uint32_t pos = atomicAdd(&cursor[keys[i]], 1u);
        blob[(size_t)(pos / kTile) * kBlobF4 + (pos % kTile)] = pts32[i];
        perm[pos] = i;
    }
}

/* one CTA (1024 threads) per tile: NaN-pads the tail, writes the 32 cell spheres and the tile
 * sphere.  R is inflated so that it also bounds the distance of the REAL (fp64) points, which
 * differ from their fp32 roundings by at most sqrt(3) * 2^-24 * mc. */
__global__ void __launch_bounds__(kTile) tile_bounds_kernel(float4 *__restrict__ blob, uint32_t n,
                                                            const CloudMeta *__restrict__ meta) {
    __shared__ float smn[32][3], smx[32][3], srr[32];
    float4 *tb = blob + (size_t)blockIdx.x * kBlobF4;
    const uint32_t gi = blockIdx.x * kTile + threadIdx.x;
    const bool valid = gi < n;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const float qnan = __int_as_float(0x7fffffff);
    float4 p = valid ? tb[threadIdx.x] : make_float4(qnan, qnan, qnan, qnan);
    if (!valid) tb[threadIdx.x] = p;
    const float slack = (float)meta->mc * 2.4e-7f; /* 2^-22 * mc */
    float mn[3] = {valid ? p.x : INFINITY, valid ? p.y : INFINITY, valid ? p.z : INFINITY};
    float mx[3] = {valid ? p.x : -INFINITY, valid ? p.y : -INFINITY, valid ? p.z : -INFINITY};
#pragma unroll
    for (int c = 0; c < 3; ++c) {
#pragma unroll
        for (int o = 16; o; o >>= 1) {
            mn[c] = fminf(mn[c], __shfl_xor_sync(0xffffffffu, mn[c], o));
            mx[c] = fmaxf(mx[c], __shfl_xor_sync(0xffffffffu, mx[c], o));
        }
    }
    const bool any = mn[0] <= mx[0];
    float cx = 0.5f * (mn[0] + mx[0]), cy = 0.5f * (mn[1] + mx[1]), cz = 0.5f * (mn[2] + mx[2]);
    float r2 = valid ? ((p.x - cx) * (p.x - cx) + (p.y - cy) * (p.y - cy) + (p.z - cz) * (p.z - cz)) : 0.f;
#pragma unroll
    for (int o = 16; o; o >>= 1) r2 = fmaxf(r2, __shfl_xor_sync(0xffffffffu, r2, o));
    if (lane == 0) {
        /* an empty cell gets R = -inf: every test culls it (its NaN points would count nothing anyway) */
        tb[kTile + w] = any ? make_float4(cx, cy, cz, sqrtf(r2) * 1.000002f + slack)
                            : make_float4(0.f, 0.f, 0.f, -INFINITY);
        for (int c = 0; c < 3; ++c) {
            smn[w][c] = mn[c];
            smx[w][c] = mx[c];
        }
    }
    __syncthreads();
    float tmn[3], tmx[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        tmn[c] = smn[lane][c];
        tmx[c] = smx[lane][c];
#pragma unroll
        for (int o = 16; o; o >>= 1) {
            tmn[c] = fminf(tmn[c], __shfl_xor_sync(0xffffffffu, tmn[c], o));
            tmx[c] = fmaxf(tmx[c], __shfl_xor_sync(0xffffffffu, tmx[c], o));
        }
    }
    cx = 0.5f * (tmn[0] + tmx[0]), cy = 0.5f * (tmn[1] + tmx[1]), cz = 0.5f * (tmn[2] + tmx[2]);
    r2 = valid ? ((p.x - cx) * (p.x - cx) + (p.y - cy) * (p.y - cy) + (p.z - cz) * (p.z - cz)) : 0.f;
#pragma unroll
    for (int o = 16; o; o >>= 1) r2 = fmaxf(r2, __shfl_xor_sync(0xffffffffu, r2, o));
    if (lane == 0) srr[w] = r2;
    __syncthreads();
    if (w == 0) {
        r2 = srr[lane];
#pragma unroll
        for (int o = 16; o; o >>= 1) r2 = fmaxf(r2, __shfl_xor_sync(0xffffffffu, r2, o));
        if (lane == 0) tb[kTile + kTileCells] = make_float4(cx, cy, cz, sqrtf(r2) * 1.000002f + slack);
    }
}

/* ---------------------------------------------------------------- the conservative sphere test */
/* plane:            a = T + 2*band, b = L >= ||w||            cull iff |t(centre)| >= a + b*R
 * sphere/cylinder:  a = mid, b = error bound of q(centre), c = r_out, d = r_in, with q = squared
 *                   distance to the centre / axis = t + mid;  cull iff the whole sphere lies outside
 *                   radius r_out (sqrt(q - b) - R >= r_out) or inside r_in (sqrt(q + b) + R <= r_in),
 *                   both tested in squared form (no square root) */
struct CullP {
    float a, b, c, d;
};

template <int KIND>
__device__ inline void make_cull(const Fast<KIND> &f, const double *m, const CloudMeta &M, double thr, CullP &k) {
    const float never = INFINITY;
    if (f.T < 0.f) { /* MinimalFit failed / unusable model: no point is ever an inlier */
        if (KIND == kPlane) {
            k.a = -1.f, k.b = 0.f, k.c = 0.f, k.d = 0.f;
        } else {
            k.a = 0.f, k.b = 0.f, k.c = -INFINITY, k.d = -1.f;
        }
        return;
    }
    if (!isfinite(f.band)) { /* fp32 cannot hold this model: every point goes through the fp64 path */
        if (KIND == kPlane) {
            k.a = never, k.b = 0.f, k.c = 0.f, k.d = 0.f;
        } else {
            k.a = 0.f, k.b = 0.f, k.c = never, k.d = -1.f;
        }
        return;
    }
    if (KIND == kPlane) {
        const double nrm = ex::plane_norm(m);
        k.a = __double2float_ru((double)f.T + 2.0 * (double)f.band);
        k.b = __double2float_ru(nrm * (1.0 + 1e-6));
        k.c = 0.f, k.d = 0.f;
    } else {
        const double r = (KIND == kSphere) ? m[3] : m[6];
        const double Hi = (r + thr) * (r + thr);
        const double Lo = (r >= thr) ? (r - thr) * (r - thr) : -Hi;
        const double mid = 0.5 * (Lo + Hi);
        /* |q_computed(centre) - q_real(centre)| <= band (the evaluation, same bound as for a point)
         * + rounding of mid and of the on-the-fly |centre|^2 */
        const double b = 2.0 * (double)f.band + 2.4e-7 * fabs(mid) + 1e-6 * M.mc * M.mc;
        k.a = (float)mid;
        k.b = __double2float_ru(b);
        k.c = __double2float_ru(sqrt(Hi + b) * (1.0 + 2e-6));
        k.d = (Lo > b) ? __double2float_rd(sqrt(Lo - b) * (1.0 - 2e-6)) : -1.f;
        if (!isfinite(k.a) || !isfinite(k.b) || !isfinite(k.c)) k.c = never, k.d = -1.f;
    }
}

/* true = certainly no inlier (and no guard-band point) inside the sphere {bd.xyz, bd.w} */
template <int KIND>
__device__ __forceinline__ bool cull_test(const float *c, const CullP &k, const float4 bd) {
    if (KIND == kPlane) {
        const float t = fmaf(c[0], bd.x, fmaf(c[1], bd.y, fmaf(c[2], bd.z, c[3])));
        return fabsf(t) >= fmaf(k.b, bd.w, k.a);
    } else {
        const float w = fmaf(bd.x, bd.x, fmaf(bd.y, bd.y, bd.z * bd.z));
        float t = fmaf(c[0], bd.x, fmaf(c[1], bd.y, fmaf(c[2], bd.z, w + c[3])));
        if (KIND == kCylinder) {
            const float b = fmaf(c[4], bd.x, fmaf(c[5], bd.y, fmaf(c[6], bd.z, c[7])));
            t = fmaf(-b, b, t);
        }
        const float q = t + k.a;
        /* sqrt(q - b) - R >= r_out  <=>  q - b >= (r_out + R)^2 ;  sqrt(q + b) + R <= r_in  <=>  q + b <= (r_in - R)^2
         * (r_out + R < 0 only for the always-cull encoding r_out = -inf or an empty cell R = -inf;
         * a NaN on either side compares false = keep) */
        const float so = k.c + bd.w, si = k.d - bd.w;
        const bool outside = (so < 0.f) || (q - k.b >= so * so * 1.000004f);
        const bool inside = (si > 0.f) && (q + k.b <= si * si * 0.999996f);
        return outside || inside;
    }
}

#ifdef M3D_CULL_STATS /* tuning builds only: how much survives each level */
__device__ unsigned long long g_cull_stats[8];
#define M3D_STAT(i, v) do { if (lane == 0) atomicAdd(&g_cull_stats[i], (unsigned long long)(v)); } while (0)
#else
#define M3D_STAT(i, v) do { } while (0)
#endif

/* ------------------------------------------------------------------------------ the hot kernel */
/* per-hypothesis parameters in shared memory, kF4 float4 each, in the order the phases read them:
 * {c[0..3]} [cylinder: {c[4..7]}] {T, band, cull.a, cull.b} [sphere, cylinder: {cull.c, cull.d, -, -}] */
template <int KIND>
struct HypLayout {
    static constexpr int kF4 = KIND == kPlane ? 2 : (KIND == kSphere ? 3 : 4);
};
template <int KIND>
__device__ __forceinline__ void load_hyp(const float4 *hp, Fast<KIND> &g, CullP &gk) {
    constexpr int NC = KIND == kCylinder ? 8 : 4;
    const float4 q0 = hp[0];
    g.c[0] = q0.x, g.c[1] = q0.y, g.c[2] = q0.z, g.c[3] = q0.w;
    if (KIND == kCylinder) {
        const float4 q1 = hp[1];
        g.c[NC - 4] = q1.x, g.c[NC - 3] = q1.y, g.c[NC - 2] = q1.z, g.c[NC - 1] = q1.w;
    }
    const float4 q2 = hp[NC / 4];
    g.T = q2.x, g.band = q2.y, gk.a = q2.z, gk.b = q2.w;
    gk.c = 0.f, gk.d = 0.f;
    if (KIND != kPlane) {
        const float4 q3 = hp[NC / 4 + 1];
        gk.c = q3.x, gk.d = q3.y;
    }
}
constexpr int kQBuf = 96; /* guard-band pairs a warp stages in shared memory before one global atomicAdd */
constexpr int kListLen = kTileCells + 4; /* per-warp list of surviving cells (+ the padding slot), 16 B aligned */

template <int KIND, int THREADS, int HPT>
constexpr size_t cull_smem_bytes() {
    return (size_t)kCullStages * kStageF4 * sizeof(float4)                      /* tile ring            */
           + (size_t)THREADS * HPT * HypLayout<KIND>::kF4 * sizeof(float4)  /* hypothesis parameters */
           + (size_t)(THREADS / 32) * kListLen * sizeof(uint32_t)           /* surviving-cell lists  */
           + (size_t)(THREADS / 32) * kQBuf * sizeof(uint2)                 /* guard-band staging    */
           + (size_t)THREADS * HPT * sizeof(uint32_t)                       /* per-hypothesis counts */
           + (size_t)kCullStages * sizeof(uint32_t)                         /* group cursor per stage */
           + 2 * kCullStages * sizeof(uint64_t) + kCullStages * sizeof(uint32_t) + 16;
}

/* the kernel is latency-bound (dependent test -> ballot -> load -> evaluate chains), so resident
 * warps matter more than registers: at least 2 CTAs of 256+32 threads / 4 of 128+32 per SM */
template <int KIND, int THREADS, int HPT>
__global__ void __launch_bounds__(THREADS + 32, (THREADS >= 256 ? 2 : 4)) score_cull_kernel(const ScoreArgs a) {
    constexpr int NC = KIND == kCylinder ? 8 : 4;
    constexpr int HF4 = HypLayout<KIND>::kF4;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float4 *tiles = reinterpret_cast<float4 *>(smem_raw);
    float4 *hyp = tiles + (size_t)kCullStages * kStageF4;
    uint32_t *lists = reinterpret_cast<uint32_t *>(hyp + (size_t)THREADS * HPT * HF4);
    uint2 *qbufs = reinterpret_cast<uint2 *>(lists + (THREADS / 32) * kListLen);
    uint64_t *full = reinterpret_cast<uint64_t *>(qbufs + (THREADS / 32) * kQBuf);
    uint64_t *empty = full + kCullStages;
    volatile uint32_t *tile_id = reinterpret_cast<volatile uint32_t *>(empty + kCullStages);
    uint32_t *scnt = const_cast<uint32_t *>(tile_id) + kCullStages; /* inlier counts of the CTA's hypotheses */
    uint32_t *grp_next = scnt + THREADS * HPT;                      /* next unclaimed hypothesis group of a stage's tile */

    const int tid = threadIdx.x;
    const uint32_t ntiles = (a.n + kTile - 1) / kTile;

    if (tid == THREADS) {
#pragma unroll
        for (int s = 0; s < kCullStages; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], THREADS / 32);
        }
        mbar_fence_init();
    }
    if (tid < kCullStages * kCellPts) { /* the dummy cell of every stage: NaN points count nothing */
        const float qnan = __int_as_float(0x7fffffff);
        tiles[(size_t)(tid / kCellPts) * kStageF4 + kBlobF4 + (tid % kCellPts)] = make_float4(qnan, qnan, qnan, qnan);
    }
    __syncthreads();

    if (tid >= THREADS) { /* ---------------- producer: claims tiles, one bulk copy per tile */
        if (tid == THREADS) {
            for (uint32_t k = 0;; ++k) {
                const int st = k % kCullStages;
                if (k >= kCullStages) mbar_wait_relaxed(&empty[st], ((k / kCullStages) - 1) & 1);
                const uint32_t t = atomicAdd(&a.tile_counter[blockIdx.x], 1u);
                if (t >= ntiles) {
                    tile_id[st] = kNoTile;
                    mbar_arrive(&full[st]);
                    break;
                }
                tile_id[st] = t;
                grp_next[st] = 0; /* published with tile_id by the barrier's release/acquire */
                tma_load_1d(tiles + (size_t)st * kStageF4, a.blob + (size_t)t * kBlobF4,
                            (uint32_t)(kBlobF4 * sizeof(float4)), &full[st]);
            }
        }
        return;
    }

    /* ---------------- consumers.  Prologue as in score_kernel: gather, MinimalFit (fp64), fp32 form */
    const CloudMeta M = *a.meta;
    const unsigned fullmask = 0xffffffffu;
    const int lane = tid & 31, warp = tid >> 5;
    uint32_t *wlist = lists + warp * kListLen;
    uint2 *qbuf = qbufs + warp * kQBuf;
    uint32_t qn = 0; /* entries staged in qbuf (warp-uniform) */
    /* staged guard-band pairs -> the global queue: one atomicAdd per flush instead of one per cell */
    auto flush_queue = [&]() {
        __syncwarp();
        if (qn) {
            uint32_t pos0 = 0;
            if (lane == 0) pos0 = atomicAdd(a.queue_count, qn);
            pos0 = __shfl_sync(fullmask, pos0, 0);
            for (uint32_t i = lane; i < qn; i += 32) {
                const uint2 e = qbuf[i];
                const uint32_t pos = pos0 + i;
                if (pos < a.queue_cap) {
                    a.queue[pos] = e;
                } else { /* queue full: decide here with the reference arithmetic */
                    const uint32_t prov = e.y >> 31, pt = a.perm[e.y & 0x7fffffffu];
                    double m[8];
                    const bool ok = fit_row<KIND>(a.xyz, a.nrm, a.samples, a.src_row(e.x), m, a.row_nrm);
                    uint32_t in = 0;
                    if (ok) {
                        ex::Dist<KIND> dist;
                        dist.set(m);
                        in = dist(ex::ld3(a.xyz + 3 * (size_t)pt)) < a.thr ? 1u : 0u;
                    }
                    if (in != prov) atomicAdd(&a.counts[a.cnt_row(e.x)], in - prov);
                }
            }
            qn = 0;
        }
        __syncwarp();
    };
    uint32_t row[HPT];
    bool invalid[HPT];
#pragma unroll
    for (int h = 0; h < HPT; ++h) { /* parameters go to shared memory; nothing of them stays in registers */
        row[h] = (blockIdx.x * HPT + h) * THREADS + tid;
        double m[8];
        bool ok = false;
        if (row[h] < a.rows) ok = fit_row<KIND>(a.xyz, a.nrm, a.samples, a.src_row(row[h]), m, a.row_nrm);
        invalid[h] = (row[h] < a.rows) && !ok;
        if (blockIdx.y == 0 && row[h] < a.rows) {
#pragma unroll
            for (int i = 0; i < 8; ++i)
                a.models[(size_t)row[h] * 8 + i] = (ok && i < param_count(KIND)) ? m[i] : 0.0;
        }
        Fast<KIND> f;
        CullP ck;
        make_fast<KIND>(m, ok, M, a.thr, f);
        make_cull<KIND>(f, m, M, a.thr, ck);
        float4 *hp = hyp + (size_t)(h * THREADS + tid) * HF4;
        hp[0] = make_float4(f.c[0], f.c[1], f.c[2], f.c[3]);
        if (KIND == kCylinder) hp[1] = make_float4(f.c[4], f.c[5], f.c[6], f.c[7]);
        hp[NC / 4] = make_float4(f.T, f.band, ck.a, ck.b);
        if (KIND != kPlane) hp[NC / 4 + 1] = make_float4(ck.c, ck.d, 0.f, 0.f);
        scnt[h * THREADS + tid] = 0;
    }
    /* from here on any consumer warp may work on any of the CTA's hypotheses: consumer-only barrier */
    asm volatile("bar.sync 1, %0;" ::"n"(THREADS) : "memory");

    /* The CTA's THREADS*HPT hypotheses form groups of 32 (CTA-local index = h*THREADS + thread).  For every
     * tile the consumer warps claim groups from the stage's cursor, so a tile's work is balanced over the
     * warps whatever the per-hypothesis cost; counts are accumulated in shared memory. */
    constexpr uint32_t kGroups = THREADS * HPT / 32;
    const uint32_t cta_row0 = blockIdx.x * HPT * THREADS; /* local index + cta_row0 = row of the launch */
    uint32_t nres = 0;
    for (uint32_t k = 0;; ++k) {
        const int st = k % kCullStages;
        mbar_wait(&full[st], (k / kCullStages) & 1);
        const uint32_t t = tile_id[st];
        if (t == kNoTile) break;
        const float4 *sp = tiles + (size_t)st * kStageF4;
        const unsigned char *spb = reinterpret_cast<const unsigned char *>(sp) + lane * sizeof(float4);
        const uint32_t base = t * kTile;
        const float4 tb = sp[kTile + kTileCells]; /* tile sphere (broadcast) */
        const float4 cb = sp[kTile + lane];       /* this lane's cell sphere */
        for (;;) {
            uint32_t grp = 0;
            if (lane == 0) grp = atomicAdd(&grp_next[st], 1u);
            grp = __shfl_sync(fullmask, grp, 0);
            if (grp >= kGroups) break;
            const uint32_t hbase = grp * 32; /* CTA-local index of the group's first hypothesis */
            /* lane = hypothesis: which of the group's hypotheses can have inliers in this tile? */
            Fast<KIND> g;
            CullP gk;
            load_hyp<KIND>(hyp + (size_t)(hbase + lane) * HF4, g, gk);
            unsigned live = __ballot_sync(fullmask, !cull_test<KIND>(g.c, gk, tb));
            M3D_STAT(0, 32);
            M3D_STAT(1, __popc(live));
            while (live) {
                const int src = __ffs(live) - 1;
                live &= live - 1;
                /* the hypothesis' parameters, warp-uniform (broadcast loads) */
                load_hyp<KIND>(hyp + (size_t)(hbase + src) * HF4, g, gk);
                /* lane = cell: surviving cells, compacted into the warp's list (byte offsets of the
                 * cells inside the stage) */
                const bool keep = !cull_test<KIND>(g.c, gk, cb);
                const unsigned cells = __ballot_sync(fullmask, keep);
                M3D_STAT(2, __popc(cells));
                if (cells == 0) continue;
                const int ncell = __popc(cells);
                __syncwarp(); /* the previous hypothesis' list has been consumed */
                if (keep) wlist[__popc(cells & ((1u << lane) - 1))] = (uint32_t)(lane * kCellPts * sizeof(float4));
                if (lane == 0) wlist[ncell] = (uint32_t)(kBlobF4 * sizeof(float4)); /* odd count: pad with the NaN cell */
                __syncwarp();
                /* lane = point: four surviving cells per trip, then the remainder two at a time */
                uint32_t cnt = 0;
                float mn = INFINITY;
                int j = 0;
                for (; j + 4 <= ncell; j += 4) {
                    const uint4 off = *reinterpret_cast<const uint4 *>(wlist + j);
                    const float4 p0 = *reinterpret_cast<const float4 *>(spb + off.x);
                    const float4 p1 = *reinterpret_cast<const float4 *>(spb + off.y);
                    const float4 p2 = *reinterpret_cast<const float4 *>(spb + off.z);
                    const float4 p3 = *reinterpret_cast<const float4 *>(spb + off.w);
                    const float v0 = fast_v<KIND>(g, p0), v1 = fast_v<KIND>(g, p1);
                    const float v2 = fast_v<KIND>(g, p2), v3 = fast_v<KIND>(g, p3);
                    cnt += (__float_as_uint(v0) >> 31) + (__float_as_uint(v1) >> 31);
                    cnt += (__float_as_uint(v2) >> 31) + (__float_as_uint(v3) >> 31);
                    mn = fminf(mn, fminf(fabsf(v0), fabsf(v1)));
                    mn = fminf(mn, fminf(fabsf(v2), fabsf(v3)));
                }
                for (; j < ncell; j += 2) {
                    const uint2 off = *reinterpret_cast<const uint2 *>(wlist + j);
                    const float4 p0 = *reinterpret_cast<const float4 *>(spb + off.x);
                    const float4 p1 = *reinterpret_cast<const float4 *>(spb + off.y);
                    const float v0 = fast_v<KIND>(g, p0), v1 = fast_v<KIND>(g, p1);
                    cnt += (__float_as_uint(v0) >> 31) + (__float_as_uint(v1) >> 31);
                    mn = fminf(mn, fminf(fabsf(v0), fabsf(v1)));
                }
                if (__any_sync(fullmask, mn < g.band)) { /* rare: stage the guard-band points of these cells */
                    const uint32_t r = cta_row0 + hbase + src;
                    unsigned cm = cells;
                    while (cm) {
                        const int c = __ffs(cm) - 1;
                        cm &= cm - 1;
                        const float v = fast_v<KIND>(g, sp[c * kCellPts + lane]);
                        const bool amb = fabsf(v) < g.band; /* false for the NaN padding */
                        const unsigned am = __ballot_sync(fullmask, amb);
                        if (am == 0) continue;
                        const uint32_t na = __popc(am);
                        if (qn + na > (uint32_t)kQBuf) flush_queue();
                        if (amb) {
                            const uint32_t prov = __float_as_uint(v) >> 31;
                            const uint32_t pt = base + c * kCellPts + lane; /* SORTED position: resolve_queue_kernel maps it through perm */
                            qbuf[qn + __popc(am & ((1u << lane) - 1))] = make_uint2(r, pt | (prov << 31));
                        }
                        qn += na;
                        nres += (lane == 0) ? na : 0u;
                    }
                }
                const uint32_t tot = __reduce_add_sync(fullmask, cnt);
                if (lane == 0 && tot) atomicAdd(&scnt[hbase + src], tot);
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[st]);
    }

    flush_queue();
    asm volatile("bar.sync 1, %0;" ::"n"(THREADS) : "memory"); /* all shared-memory counts are final */
#pragma unroll
    for (int h = 0; h < HPT; ++h) {
        if (row[h] < a.rows) {
            const uint32_t c = scnt[h * THREADS + tid];
            const uint32_t ci = a.cnt_row(row[h]);
            if (c) atomicAdd(&a.counts[ci], c);
            if (invalid[h] && blockIdx.y == 0) atomicOr(&a.counts[ci], kInvalidBit);
        }
    }
    if (nres) atomicAdd(a.resolves, (unsigned long long)nres);
}

/* ------------------------------------------------------------------------- hypothesis classification
 * Culling cannot remove pairs that are inliers: a hypothesis whose shell passes through most of the cloud
 * (e.g. THE dominant plane of the scene) is cheaper in the dense kernel, where one point load serves 32-64
 * hypotheses.  For large waves one thread per hypothesis estimates the fraction of cells that would survive
 * (every `stride`-th tile: tile sphere, then its 32 cell spheres) and files the row in the front (cull) or
 * back (dense) part of `row_map`.  The order inside the two parts is whatever the atomics give; counts do
 * not depend on it.  part[0] = rows for the culling kernel, part[1] = rows for the dense kernel. */
template <int KIND>
__global__ void __launch_bounds__(128) cull_classify_kernel(const ScoreArgs a, uint32_t ntiles, uint32_t stride,
                                                            float dense_above, uint32_t *__restrict__ row_map,
                                                            uint32_t *__restrict__ part) {
    const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    const bool active = r < a.rows;
    const CloudMeta M = *a.meta;
    double m[8];
    bool ok = false;
    if (active) ok = fit_row<KIND>(a.xyz, a.nrm, a.samples, a.src_row(r), m, a.row_nrm);
    Fast<KIND> f;
    CullP ck;
    make_fast<KIND>(m, ok, M, a.thr, f);
    make_cull<KIND>(f, m, M, a.thr, ck);
    uint32_t seen = 0, kept = 0;
    for (uint32_t t = (r % stride); t < ntiles; t += stride) { /* staggered so that a warp covers all tiles */
        const float4 *tb = a.blob + (size_t)t * kBlobF4 + kTile;
        seen += kTileCells;
        if (cull_test<KIND>(f.c, ck, tb[kTileCells])) continue;
#pragma unroll 4
        for (int c = 0; c < kTileCells; ++c) kept += cull_test<KIND>(f.c, ck, tb[c]) ? 0u : 1u;
    }
    const bool dense = active && seen && ((float)kept > dense_above * (float)seen);
    const unsigned full = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const unsigned md = __ballot_sync(full, dense), mc = __ballot_sync(full, active && !dense);
    uint32_t bd = 0, bc = 0;
    if (lane == 0) {
        if (md) bd = atomicAdd(&part[1], (uint32_t)__popc(md));
        if (mc) bc = atomicAdd(&part[0], (uint32_t)__popc(mc));
    }
    bd = __shfl_sync(full, bd, 0);
    bc = __shfl_sync(full, bc, 0);
    const unsigned lt = (1u << lane) - 1;
    if (dense) row_map[a.rows - 1 - (bd + __popc(md & lt))] = a.row_begin + r;
    if (active && !dense) row_map[bc + __popc(mc & lt)] = a.row_begin + r;
}

}  // namespace m3d
