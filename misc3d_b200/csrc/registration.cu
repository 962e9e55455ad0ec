/*
 * registration.cu -- correspondence-based RANSAC rigid registration on sm_100a.
 *
 * Replaces misc3d::registration::RANSACSolver::Solve (src/transform_estimation.cpp:124-164), which
 * delegates to Open3D v0.15.1 RegistrationRANSACBasedOnCorrespondence:
 * per iteration draw 3 correspondences (with replacement), estimate T by Eigen::umeyama (no
 * scaling), run the edge-length and distance checkers, score T over all correspondences
 * (d^2 < max_dist^2), keep the best by (fitness, rmse) and shrink the iteration budget est_k.
 * And LeastSquareSolver::Solve = Eigen::umeyama over all pairs (:49-66).
 *
 *   reg_solve_kernel    thread = hypothesis: gather, 3-point Umeyama (fp64, two-sided Jacobi SVD in
 *                       the reference's operation order, no FMA contraction), both checkers
 *   reg_score_kernel    thread = surviving hypothesis, correspondences ({p, q} as 2 x float4, centred)
 *                       staged through shared memory by 1-D TMA bulk copies; fp32 guard-banded count
 *                       (sign of |Rp+t-q|^2 - thr^2 and min |.|), ambiguous pairs queued
 *   reg_resolve_kernel  fp64 reference-order decision of the queued pairs
 *   reg_eval_kernel / reg_seq_eval_kernel   (count, sum d^2) of one transform: parallel sum, or the
 *                       reference's index-order sum (tie-breaks only)
 * The sequential bookkeeping (best, est_k, stop index) is replayed on the host in loop order.
 */
#include <algorithm>
#include <cfloat>
#include <climits>
#include <cmath>
#include <random>
#include <vector>

#include "context.h"
#include "exact_math.cuh"
#include "umeyama.cuh"

namespace m3d {


constexpr int kRegTile = 512; /* correspondences per TMA stage: 2 x float4 each = 16 KB */
constexpr int kRegStages = 3;
constexpr int kRegSub = 32;
constexpr int kRegThreads = 128;
constexpr double kRU32 = 5.9604644775390625e-08, kRU64 = 1.1102230246251565e-16;

struct RegMeta {
    double cs[3], cd[3]; /* centres of the gathered source / target points */
    double mp, mq;       /* max |centred coordinate| of p / q              */
    double mraw;         /* max |raw coordinate| over both                 */
};

struct RegArgs {
    const double *src, *dst;   /* raw clouds, n x 3 f64 */
    const uint32_t *c0, *c1;   /* correspondences       */
    uint32_t m;
    const float4 *pq;          /* [m][2]: centred fp32 {p, q} */
    const RegMeta *meta;
    const uint32_t *picks;     /* rows x 3 */
    uint32_t rows;
    double thr, thr2, edge_thr;
    double *T;                 /* [rows][16] */
    uint8_t *pass;             /* [rows]     */
    uint32_t *list;            /* surviving rows (any order) */
    uint32_t *list_count;
    uint32_t *counts;          /* [rows] */
    double *err2;              /* [rows] approximate sum of inlier d^2 (fp32 per sub-tile, fp64 across) */
    uint2 *queue;
    uint32_t *queue_count;
    uint32_t queue_cap;
    uint32_t chunk_tiles;
};

/* gather p_i = src[c0[i]], q_i = dst[c1[i]]: bounding boxes (for centring) */
__global__ void __launch_bounds__(256) reg_bbox_kernel(RegArgs a, unsigned long long *mn, unsigned long long *mx);
__device__ __forceinline__ unsigned long long r_enc(double x) {
    const unsigned long long b = (unsigned long long)__double_as_longlong(x);
    return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}
__device__ __forceinline__ double r_dec(unsigned long long e) {
    const unsigned long long b = (e >> 63) ? (e & 0x7fffffffffffffffull) : ~e;
    return __longlong_as_double((long long)b);
}
/* mnmx: 6 mins then 6 maxs (p xyz, q xyz) */
__global__ void __launch_bounds__(256) reg_bbox_kernel(RegArgs a, unsigned long long *mn, unsigned long long *mx) {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < a.m; i += gridDim.x * blockDim.x) {
        const double *p = a.src + 3 * (size_t)a.c0[i], *q = a.dst + 3 * (size_t)a.c1[i];
#pragma unroll
        for (int c = 0; c < 6; ++c) {
            const double v = c < 3 ? p[c] : q[c - 3];
            if (v != v) continue;
            const unsigned long long e = r_enc(v);
            if (e < mn[c]) atomicMin(&mn[c], e);
            if (e > mx[c]) atomicMax(&mx[c], e);
        }
    }
}
__global__ void reg_meta_kernel(const unsigned long long *mn, const unsigned long long *mx, RegMeta *meta) {
    if (threadIdx.x != 0) return;
    double mp = 0, mq = 0, mraw = 0;
    for (int c = 0; c < 6; ++c) {
        const double lo = r_dec(mn[c]), hi = r_dec(mx[c]);
        double ctr = 0.5 * (lo + hi);
        if (!isfinite(ctr)) ctr = 0;
        const double ext = fmax(fabs(hi - ctr), fabs(lo - ctr));
        if (c < 3) {
            meta->cs[c] = ctr;
            mp = fmax(mp, ext);
        } else {
            meta->cd[c - 3] = ctr;
            mq = fmax(mq, ext);
        }
        mraw = fmax(mraw, fmax(fabs(lo), fabs(hi)));
    }
    meta->mp = mp * (1 + 1e-12);
    meta->mq = mq * (1 + 1e-12);
    meta->mraw = mraw;
}
__global__ void __launch_bounds__(256) reg_gather_kernel(RegArgs a, float4 *pq) {
    const RegMeta M = *a.meta;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < a.m; i += gridDim.x * blockDim.x) {
        const double *p = a.src + 3 * (size_t)a.c0[i], *q = a.dst + 3 * (size_t)a.c1[i];
        pq[2 * (size_t)i] = make_float4((float)(p[0] - M.cs[0]), (float)(p[1] - M.cs[1]), (float)(p[2] - M.cs[2]), 0.f);
        pq[2 * (size_t)i + 1] = make_float4((float)(q[0] - M.cd[0]), (float)(q[1] - M.cd[1]), (float)(q[2] - M.cd[2]), 0.f);
    }
}

/* thread = hypothesis: TransformationEstimationPointToPoint(false)::ComputeTransformation on the 3
 * picked correspondences + CorrespondenceCheckerBasedOnEdgeLength + ...BasedOnDistance */
__global__ void __launch_bounds__(128) reg_solve_kernel(RegArgs a) {
    const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= a.rows) return;
    ex::V3 sp[3], dp[3];
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        const uint32_t pk = a.picks[3 * (size_t)r + j];
        sp[j] = ex::ld3(a.src + 3 * (size_t)a.c0[pk]);
        dp[j] = ex::ld3(a.dst + 3 * (size_t)a.c1[pk]);
    }
    /* Eigen::umeyama on 3 points */
    const double one_over_n = rg::div(1.0, 3.0);
    double sm[3] = {0, 0, 0}, dm[3] = {0, 0, 0};
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        sm[0] = rg::add(sm[0], sp[i].x);
        sm[1] = rg::add(sm[1], sp[i].y);
        sm[2] = rg::add(sm[2], sp[i].z);
        dm[0] = rg::add(dm[0], dp[i].x);
        dm[1] = rg::add(dm[1], dp[i].y);
        dm[2] = rg::add(dm[2], dp[i].z);
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        sm[c] = rg::mul(sm[c], one_over_n);
        dm[c] = rg::mul(dm[c], one_over_n);
    }
    double sigma[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        const double sd[3] = {rg::sub(sp[i].x, sm[0]), rg::sub(sp[i].y, sm[1]), rg::sub(sp[i].z, sm[2])};
        const double dd[3] = {rg::sub(dp[i].x, dm[0]), rg::sub(dp[i].y, dm[1]), rg::sub(dp[i].z, dm[2])};
#pragma unroll
        for (int rr = 0; rr < 3; ++rr)
#pragma unroll
            for (int c = 0; c < 3; ++c)
                sigma[rr][c] = rg::add(sigma[rr][c], rg::mul(rg::mul(one_over_n, dd[rr]), sd[c]));
    }
    double T[16];
    rg::umeyama_finish(sigma, sm, dm, false, 1.0, T);
    bool pass = true;
    /* edge lengths: all pairs i < j, fail if ds < dt*s || dt < ds*s */
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = i + 1; j < 3; ++j) {
            const double ds = ex::norm3(ex::sub3(sp[i], sp[j]));
            const double dt = ex::norm3(ex::sub3(dp[i], dp[j]));
            if (ds < rg::mul(dt, a.edge_thr) || dt < rg::mul(ds, a.edge_thr)) pass = false;
        }
    if (pass) {
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            const ex::V3 pt = rg::xform(T, sp[j]);
            if (ex::norm3(ex::sub3(dp[j], pt)) > a.thr) pass = false;
        }
    }
#pragma unroll
    for (int i = 0; i < 16; ++i) a.T[(size_t)r * 16 + i] = T[i];
    a.pass[r] = pass ? 1 : 0;
    a.counts[r] = 0;
    a.err2[r] = 0.0;
    if (pass) a.list[atomicAdd(a.list_count, 1u)] = r;
}

struct RegFast {
    float R[9], t[3];
    float band;
};
__device__ __forceinline__ float reg_v(const RegFast &f, const float4 p, const float4 q, float thr2) {
    const float rx = fmaf(f.R[0], p.x, fmaf(f.R[1], p.y, fmaf(f.R[2], p.z, __fsub_rn(f.t[0], q.x))));
    const float ry = fmaf(f.R[3], p.x, fmaf(f.R[4], p.y, fmaf(f.R[5], p.z, __fsub_rn(f.t[1], q.y))));
    const float rz = fmaf(f.R[6], p.x, fmaf(f.R[7], p.y, fmaf(f.R[8], p.z, __fsub_rn(f.t[2], q.z))));
    return __fsub_rn(fmaf(rx, rx, fmaf(ry, ry, __fmul_rn(rz, rz))), thr2);
}
__device__ inline void reg_make_fast(const double *T, const RegMeta &M, double thr, double thr2, RegFast &f) {
    double E2 = 0, smax = 0;
    bool fin = true;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        const double tc = T[4 * i + 3] + (T[4 * i] * M.cs[0] + T[4 * i + 1] * M.cs[1] + T[4 * i + 2] * M.cs[2]) - M.cd[i];
        const double l1 = fabs(T[4 * i]) + fabs(T[4 * i + 1]) + fabs(T[4 * i + 2]);
        const double S = l1 * M.mp + fabs(tc) + M.mq;
        const double e = 10 * kRU32 * S + 16 * kRU64 * (l1 * M.mraw + fabs(T[4 * i + 3]) + M.mraw);
        E2 += e * e;
        smax = fmax(smax, S);
        f.t[i] = (float)tc;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            f.R[3 * i + c] = (float)T[4 * i + c];
            fin = fin && isfinite(f.R[3 * i + c]);
        }
        fin = fin && isfinite(f.t[i]);
    }
    const double E = sqrt(E2);
    /* | |r^|^2 - |r|^2 | <= 2|r|E + E^2 (+ roundings of the squares); near the threshold |r| ~ thr */
    const double band = 1.5 * (2 * (thr + E) * E + E * E) + 16 * kRU32 * thr2;
    f.band = __double2float_ru(band);
    if (!fin || !isfinite(f.band) || !(smax < 1e15)) {
#pragma unroll
        for (int i = 0; i < 9; ++i) f.R[i] = 0.f;
        f.t[0] = f.t[1] = f.t[2] = 0.f;
        f.band = INFINITY; /* every pair goes to the fp64 path */
    }
}

__device__ __forceinline__ uint32_t r_smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void r_mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "LAB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra DONE;\n"
        "bra LAB_WAIT;\n"
        "DONE:\n"
        "}\n" ::"r"(r_smem_u32(bar)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void r_tma_load_1d(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
    const uint32_t b = r_smem_u32(bar);
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(bytes) : "memory");
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
            r_smem_u32(dst)),
        "l"(src), "r"(bytes), "r"(b)
        : "memory");
}

__device__ __noinline__ void reg_rescan(const RegArgs &a, uint32_t row, const RegFast f, float thr2f,
                                        const float4 *sp, uint32_t gbase, int cnt) {
    for (int j = 0; j < cnt; ++j) {
        const float v = reg_v(f, sp[2 * j], sp[2 * j + 1], thr2f);
        if (fabsf(v) < f.band) {
            const uint32_t prov = __float_as_uint(v) >> 31;
            const uint32_t pos = atomicAdd(a.queue_count, 1u);
            if (pos < a.queue_cap) {
                a.queue[pos] = make_uint2(row, (gbase + j) | (prov << 31));
            } else { /* queue full: decide here */
                const uint32_t i = gbase + j;
                const double d2 = rg::dis2(a.T + (size_t)row * 16, ex::ld3(a.src + 3 * (size_t)a.c0[i]),
                                           ex::ld3(a.dst + 3 * (size_t)a.c1[i]));
                const uint32_t in = d2 < a.thr2 ? 1u : 0u;
                if (in != prov) atomicAdd(&a.counts[row], in - prov);
            }
        }
    }
}

/* EvaluateRANSACBasedOnCorrespondence for every surviving hypothesis: grid = (hypothesis blocks,
 * correspondence chunks) */
__global__ void __launch_bounds__(kRegThreads) reg_score_kernel(const RegArgs a) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float4 *tiles = reinterpret_cast<float4 *>(smem_raw);
    uint64_t *full = reinterpret_cast<uint64_t *>(smem_raw + (size_t)kRegStages * kRegTile * 2 * sizeof(float4));
    const int tid = threadIdx.x;
    const uint32_t nsurv = *a.list_count;
    if (blockIdx.x * kRegThreads >= nsurv) return; /* whole CTA idle */
    const uint32_t ntiles = (a.m + kRegTile - 1) / kRegTile;
    const uint32_t t0 = blockIdx.y * a.chunk_tiles, t1 = min(t0 + a.chunk_tiles, ntiles);
    auto issue = [&](uint32_t t) {
        const uint32_t base = t * kRegTile, npt = min((uint32_t)kRegTile, a.m - base);
        const int st = (t - t0) % kRegStages;
        r_tma_load_1d(tiles + (size_t)st * kRegTile * 2, a.pq + 2 * (size_t)base, npt * 32u, &full[st]);
    };
    if (tid == 0) {
        for (int s = 0; s < kRegStages; ++s)
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(r_smem_u32(&full[s])), "r"(1) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        for (uint32_t t = t0; t < t1 && t < t0 + kRegStages; ++t) issue(t);
    }
    const uint32_t li = blockIdx.x * kRegThreads + tid;
    const bool active = li < nsurv;
    const uint32_t row = active ? a.list[li] : 0;
    RegFast f;
    const float thr2f = (float)a.thr2;
    if (active) {
        reg_make_fast(a.T + (size_t)row * 16, *a.meta, a.thr, a.thr2, f);
    } else {
        for (int i = 0; i < 9; ++i) f.R[i] = 0.f;
        f.t[0] = f.t[1] = f.t[2] = 0.f;
        f.band = -1.f; /* never ambiguous */
    }
    uint32_t clo = 0;
    float mn = INFINITY;
    double esum = 0; /* sum over provisional inliers of d^2 (only used to rank count ties) */
    __syncthreads();
    for (uint32_t t = t0; t < t1; ++t) {
        const int st = (t - t0) % kRegStages;
        r_mbar_wait(&full[st], ((t - t0) / kRegStages) & 1);
        const float4 *sp = tiles + (size_t)st * kRegTile * 2;
        const uint32_t base = t * kRegTile;
        const int npt = (int)min((uint32_t)kRegTile, a.m - base);
        for (int s0 = 0; s0 < npt; s0 += kRegSub) {
            const int cnt = min(kRegSub, npt - s0);
            const uint32_t c0 = clo;
            float es = 0.f; /* sum of min(v, 0) = sum over inliers of (d^2 - thr^2) */
#pragma unroll 4
            for (int j = 0; j < cnt; ++j) {
                const float v = reg_v(f, sp[2 * (s0 + j)], sp[2 * (s0 + j) + 1], thr2f);
                clo += __float_as_uint(v) >> 31;
                mn = fminf(mn, fabsf(v));
                es += fminf(v, 0.f);
            }
            esum += (double)es + (double)(clo - c0) * (double)thr2f;
            if (mn < f.band) {
                reg_rescan(a, row, f, thr2f, sp + 2 * s0, base + s0, cnt);
                mn = INFINITY;
            }
            __syncwarp();
        }
        __syncthreads();
        if (tid == 0 && t + kRegStages < t1) issue(t + kRegStages);
    }
    if (active && clo) {
        atomicAdd(&a.counts[row], clo);
        atomicAdd(&a.err2[row], esum);
    }
}

__global__ void __launch_bounds__(256) reg_resolve_kernel(const RegArgs a) {
    const uint32_t total = min(*a.queue_count, a.queue_cap);
    for (uint32_t k = blockIdx.x * blockDim.x + threadIdx.x; k < total; k += gridDim.x * blockDim.x) {
        const uint2 e = a.queue[k];
        const uint32_t prov = e.y >> 31, i = e.y & 0x7fffffffu;
        const double d2 = rg::dis2(a.T + (size_t)e.x * 16, ex::ld3(a.src + 3 * (size_t)a.c0[i]),
                                   ex::ld3(a.dst + 3 * (size_t)a.c1[i]));
        const uint32_t in = d2 < a.thr2 ? 1u : 0u;
        if (in != prov) atomicAdd(&a.counts[e.x], in - prov);
    }
}

/* fp64 reference-order scoring (debug / cross-check): thread = surviving hypothesis */
__global__ void __launch_bounds__(128) reg_score_exact_kernel(const RegArgs a) {
    const uint32_t li = blockIdx.x * blockDim.x + threadIdx.x;
    if (li >= *a.list_count) return;
    const uint32_t row = a.list[li];
    double T[16];
    for (int i = 0; i < 16; ++i) T[i] = a.T[(size_t)row * 16 + i];
    uint32_t c = 0;
    for (uint32_t i = 0; i < a.m; ++i)
        c += rg::dis2(T, ex::ld3(a.src + 3 * (size_t)a.c0[i]), ex::ld3(a.dst + 3 * (size_t)a.c1[i])) < a.thr2 ? 1u : 0u;
    a.counts[row] = c;
}

/* (good, sum d^2) of one transform; out[0] = count (as double), out[1] = sum.  Fixed-order
 * block partials, reduced by the last launch below. */
__global__ void __launch_bounds__(256) reg_eval_kernel(const RegArgs a, const double *T, double *part) {
    double Tl[16];
    for (int i = 0; i < 16; ++i) Tl[i] = T[i];
    double s = 0, c = 0;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < a.m; i += gridDim.x * blockDim.x) {
        const double d2 = rg::dis2(Tl, ex::ld3(a.src + 3 * (size_t)a.c0[i]), ex::ld3(a.dst + 3 * (size_t)a.c1[i]));
        if (d2 < a.thr2) {
            s += d2;
            c += 1;
        }
    }
    __shared__ double sh[2][8];
    for (int o = 16; o; o >>= 1) {
        s += __shfl_xor_sync(0xffffffffu, s, o);
        c += __shfl_xor_sync(0xffffffffu, c, o);
    }
    if ((threadIdx.x & 31) == 0) {
        sh[0][threadIdx.x >> 5] = c;
        sh[1][threadIdx.x >> 5] = s;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        double cc = 0, ss = 0;
        for (int k = 0; k < 8; ++k) {
            cc += sh[0][k];
            ss += sh[1][k];
        }
        part[2 * blockIdx.x] = cc;
        part[2 * blockIdx.x + 1] = ss;
    }
}
__global__ void reg_eval_final_kernel(const double *part, int nparts, double *out) {
    if (threadIdx.x != 0) return;
    double c = 0, s = 0;
    for (int k = 0; k < nparts; ++k) {
        c += part[2 * k];
        s += part[2 * k + 1];
    }
    out[0] = c;
    out[1] = s;
}
/* the reference's own accumulation: one accumulator, correspondence order */
__global__ void reg_seq_eval_kernel(const RegArgs a, const double *T, double *out) {
    double Tl[16];
    for (int i = 0; i < 16; ++i) Tl[i] = T[i];
    const int lane = threadIdx.x;
    double acc = 0, cnt = 0;
    for (uint32_t b = 0; b < a.m; b += 32) {
        const uint32_t i = b + lane;
        double d2 = INFINITY;
        if (i < a.m) d2 = rg::dis2(Tl, ex::ld3(a.src + 3 * (size_t)a.c0[i]), ex::ld3(a.dst + 3 * (size_t)a.c1[i]));
        const int n = (int)min(32u, a.m - b);
        for (int k = 0; k < n; ++k) {
            const double v = __shfl_sync(0xffffffffu, d2, k);
            if (v < a.thr2) {
                acc = ex::add(acc, v);
                cnt += 1;
            }
        }
    }
    if (lane == 0) {
        out[0] = cnt;
        out[1] = acc;
    }
}

/* ---- LeastSquareSolver: Eigen::umeyama over all n pairs (two fixed-order reduction passes) */
__global__ void __launch_bounds__(256) lsq_mean_kernel(const double *s, const double *d, uint32_t n, double *part) {
    double acc[6] = {0, 0, 0, 0, 0, 0};
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
        for (int c = 0; c < 3; ++c) {
            acc[c] += s[3 * (size_t)i + c];
            acc[3 + c] += d[3 * (size_t)i + c];
        }
    __shared__ double sh[8][6];
    for (int k = 0; k < 6; ++k)
        for (int o = 16; o; o >>= 1) acc[k] += __shfl_xor_sync(0xffffffffu, acc[k], o);
    if ((threadIdx.x & 31) == 0)
        for (int k = 0; k < 6; ++k) sh[threadIdx.x >> 5][k] = acc[k];
    __syncthreads();
    if (threadIdx.x == 0)
        for (int k = 0; k < 6; ++k) {
            double r = 0;
            for (int w = 0; w < 8; ++w) r += sh[w][k];
            part[6 * blockIdx.x + k] = r;
        }
}
__global__ void lsq_mean_final_kernel(const double *part, int nparts, uint32_t n, double *mean) {
    if (threadIdx.x >= 6) return;
    double r = 0;
    for (int k = 0; k < nparts; ++k) r += part[6 * k + threadIdx.x];
    mean[threadIdx.x] = r * (1.0 / (double)n);
}
__global__ void __launch_bounds__(256) lsq_cov_kernel(const double *s, const double *d, uint32_t n,
                                                      const double *mean, double *part) {
    double acc[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    const double inv = 1.0 / (double)n;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        double sd[3], dd[3];
        for (int c = 0; c < 3; ++c) {
            sd[c] = s[3 * (size_t)i + c] - mean[c];
            dd[c] = d[3 * (size_t)i + c] - mean[3 + c];
        }
        for (int r = 0; r < 3; ++r)
            for (int c = 0; c < 3; ++c) acc[3 * r + c] += (inv * dd[r]) * sd[c];
        acc[9] += sd[0] * sd[0] + sd[1] * sd[1] + sd[2] * sd[2];
    }
    __shared__ double sh[8][10];
    for (int k = 0; k < 10; ++k)
        for (int o = 16; o; o >>= 1) acc[k] += __shfl_xor_sync(0xffffffffu, acc[k], o);
    if ((threadIdx.x & 31) == 0)
        for (int k = 0; k < 10; ++k) sh[threadIdx.x >> 5][k] = acc[k];
    __syncthreads();
    if (threadIdx.x == 0)
        for (int k = 0; k < 10; ++k) {
            double r = 0;
            for (int w = 0; w < 8; ++w) r += sh[w][k];
            part[10 * blockIdx.x + k] = r;
        }
}
__global__ void lsq_final_kernel(const double *part, int nparts, uint32_t n, const double *mean, int with_scaling,
                                 double *T) {
    if (threadIdx.x != 0) return;
    double acc[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    for (int k = 0; k < nparts; ++k)
        for (int q = 0; q < 10; ++q) acc[q] += part[10 * k + q];
    double sigma[3][3], sm[3], dm[3];
    for (int r = 0; r < 3; ++r) {
        sm[r] = mean[r];
        dm[r] = mean[3 + r];
        for (int c = 0; c < 3; ++c) sigma[r][c] = acc[3 * r + c];
    }
    double Tl[16];
    rg::umeyama_finish(sigma, sm, dm, with_scaling != 0, acc[9] / (double)n, Tl);
    for (int i = 0; i < 16; ++i) T[i] = Tl[i];
}

/* ---- refit on the inlier correspondences (LeastSquareSolver as the optional step behind RANSACSolver).
 * One CTA: thread t owns a contiguous run of correspondences, so the selected pairs keep correspondence order
 * (fixed-order sums downstream).  A pair is an inlier when |T s - d|^2 < thr^2 (Open3D's max_correspondence_distance
 * test), T applied like pcd.Transform (divide by the homogeneous coordinate). */
__global__ void __launch_bounds__(1024) refit_select_kernel(const double *__restrict__ src, const double *__restrict__ dst,
                                                            const uint32_t *__restrict__ c0, const uint32_t *__restrict__ c1,
                                                            uint32_t m, const double *__restrict__ T, double thr2,
                                                            double *__restrict__ out_s, double *__restrict__ out_d,
                                                            uint32_t *__restrict__ count) {
    __shared__ uint32_t warp_tot[32];
    double t[16];
    for (int q = 0; q < 16; ++q) t[q] = T[q];
    const uint32_t per = (m + blockDim.x - 1) / blockDim.x;
    const uint32_t b = min(m, threadIdx.x * per), e = min(m, b + per);
    auto inlier = [&](uint32_t i) {
        const double *sp = src + 3 * (size_t)c0[i], *dp = dst + 3 * (size_t)c1[i];
        const double x = sp[0], y = sp[1], z = sp[2];
        const double w = ((t[12] * x + t[13] * y) + t[14] * z) + t[15];
        const double dx = (((t[0] * x + t[1] * y) + t[2] * z) + t[3]) / w - dp[0];
        const double dy = (((t[4] * x + t[5] * y) + t[6] * z) + t[7]) / w - dp[1];
        const double dz = (((t[8] * x + t[9] * y) + t[10] * z) + t[11]) / w - dp[2];
        return (dx * dx + dy * dy) + dz * dz < thr2;
    };
    uint32_t mine = 0;
    for (uint32_t i = b; i < e; ++i) mine += inlier(i) ? 1u : 0u;
    /* exclusive scan of `mine` over the block */
    const int lane = threadIdx.x & 31, wp = threadIdx.x >> 5;
    uint32_t inc = mine;
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t v = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += v;
    }
    if (lane == 31) warp_tot[wp] = inc;
    __syncthreads();
    if (wp == 0) {
        uint32_t v = warp_tot[lane], r = v;
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t u = __shfl_up_sync(0xffffffffu, r, o);
            if (lane >= o) r += u;
        }
        warp_tot[lane] = r - v; /* exclusive */
        if (lane == 31) *count = r;
    }
    __syncthreads();
    uint32_t pos = warp_tot[wp] + inc - mine;
    for (uint32_t i = b; i < e; ++i)
        if (inlier(i)) {
            const double *sp = src + 3 * (size_t)c0[i], *dp = dst + 3 * (size_t)c1[i];
            for (int a = 0; a < 3; ++a) {
                out_s[3 * (size_t)pos + a] = sp[a];
                out_d[3 * (size_t)pos + a] = dp[a];
            }
            ++pos;
        }
}

/* ------------------------------------------------------------------------------ host side */
/* Eigen::umeyama over n dense pairs already on the device; part = scratch of 10 nb + 32 doubles, the 4x4 result is
 * left at part + 10 nb + 8 (row-major) */
static int lsq_on_device(m3d_ctx *ctx, const double *d_s, const double *d_d, uint32_t n, int nb, int with_scaling, double *part) {
    double *mean = part + 10 * (size_t)nb, *dT = mean + 8;
    lsq_mean_kernel<<<nb, 256, 0, ctx->stream>>>(d_s, d_d, n, part);
    M3D_LAUNCHED(ctx);
    lsq_mean_final_kernel<<<1, 32, 0, ctx->stream>>>(part, nb, n, mean);
    M3D_LAUNCHED(ctx);
    lsq_cov_kernel<<<nb, 256, 0, ctx->stream>>>(d_s, d_d, n, mean, part);
    M3D_LAUNCHED(ctx);
    lsq_final_kernel<<<1, 32, 0, ctx->stream>>>(part, nb, n, mean, with_scaling, dT);
    M3D_LAUNCHED(ctx);
    return M3D_OK;
}

/* est_k update on improvement (Open3D): (int)ceil of a non-finite value is UB in the reference --
 * emulated as INT_MIN, which is what x86-64 yields */
static int reg_update_limit(uint64_t good, size_t m, double confidence, int est_k) {
    const double ratio = (double)good / (double)m;
    const double est = std::log(1.0 - confidence) / std::log(1.0 - std::pow(ratio, 3));
    if (est < (double)est_k) {
        const double c = std::ceil(est);
        if (c >= -2147483648.0 && c <= 2147483647.0) return (int)c;
        return INT_MIN;
    }
    return est_k;
}

struct RegSmall {
    unsigned long long mn[6], mx[6];
    RegMeta meta;
    uint32_t list_count, queue_count;
    double T[16];
    double eval[2];
};

}  // namespace m3d

using namespace m3d;

extern "C" {

int m3d_ransac_registration(m3d_ctx *ctx, const double *src_xyz, size_t ns, const double *dst_xyz, size_t nd,
                            const size_t *c0, const size_t *c1, size_t m, double threshold, int max_iter,
                            double edge_thr, double confidence, uint32_t seed, double *T_out, m3d_reg_stats *stats) {
    if (!ctx || !T_out || (m && (!c0 || !c1)) || (ns && !src_xyz) || (nd && !dst_xyz)) return M3D_ERR_INVALID_ARG;
    static const double I4[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};
    memcpy(T_out, I4, sizeof I4);
    m3d_reg_stats st{};
    st.stop_index = max_iter > 0 ? (uint64_t)max_iter : 0;
    if (stats) *stats = st;
    if (ns < 3 || nd < 3) /* transform_estimation.cpp:130-133 throws */
        return ctx->fail(M3D_ERR_TOO_FEW_POINTS, "There must be at least 3 points to solve the transformation");
    if (m < 3 || !(threshold > 0.0) || max_iter <= 0) return 0; /* Open3D returns its default result */
    if (m >= (1ull << 31) || ns >= (1ull << 32) || nd >= (1ull << 32))
        return ctx->fail(M3D_ERR_INVALID_ARG, "too many points / correspondences");
    for (size_t i = 0; i < m; ++i)
        if (c0[i] >= ns || c1[i] >= nd) return ctx->fail(M3D_ERR_INVALID_ARG, "correspondence %zu out of range", i);
    M3D_CUDA(ctx, cudaSetDevice(ctx->device));

    /* device buffers */
    constexpr uint32_t kWaveMax = 1u << 16, kQueueCap = 1u << 22;
    M3D_CUDA(ctx, ctx->d_tmp0.reserve(sizeof(double) * 3 * ns));
    M3D_CUDA(ctx, ctx->d_tmp1.reserve(sizeof(double) * 3 * nd));
    M3D_CUDA(ctx, ctx->d_tmp2.reserve(sizeof(uint32_t) * 2 * m));
    M3D_CUDA(ctx, ctx->d_tmp3.reserve(sizeof(float4) * 2 * m));
    M3D_CUDA(ctx, ctx->d_tmp4.reserve((sizeof(double) * 17 + 1 + 4 + 4) * (size_t)kWaveMax + 64));
    M3D_CUDA(ctx, ctx->d_samples.reserve(sizeof(uint32_t) * 3 * (size_t)kWaveMax));
    M3D_CUDA(ctx, ctx->h_samples.reserve(sizeof(uint32_t) * 3 * (size_t)kWaveMax));
    M3D_CUDA(ctx, ctx->d_queue.reserve(sizeof(uint2) * (size_t)kQueueCap + 16));
    M3D_CUDA(ctx, ctx->d_small.reserve(sizeof(RegSmall) + 4096));
    M3D_CUDA(ctx, ctx->h_small.reserve(sizeof(RegSmall) + 4096));
    M3D_CUDA(ctx, ctx->h_counts.reserve((8 + 4 + 1) * (size_t)kWaveMax));
    M3D_CUDA(ctx, ctx->d_part.reserve(sizeof(double) * 2 * 1024));
    RegSmall *ds = ctx->d_small.as<RegSmall>();
    RegSmall *hs = ctx->h_small.as<RegSmall>();

    std::vector<uint32_t> hc(2 * m);
    for (size_t i = 0; i < m; ++i) {
        hc[i] = (uint32_t)c0[i];
        hc[m + i] = (uint32_t)c1[i];
    }
    M3D_CUDA(ctx, cudaEventRecord(ctx->ev[0], ctx->stream));
    M3D_CUDA(ctx, cudaMemcpyAsync(ctx->d_tmp0.p, src_xyz, sizeof(double) * 3 * ns, cudaMemcpyHostToDevice, ctx->stream));
    M3D_CUDA(ctx, cudaMemcpyAsync(ctx->d_tmp1.p, dst_xyz, sizeof(double) * 3 * nd, cudaMemcpyHostToDevice, ctx->stream));
    M3D_CUDA(ctx, cudaMemcpyAsync(ctx->d_tmp2.p, hc.data(), sizeof(uint32_t) * 2 * m, cudaMemcpyHostToDevice, ctx->stream));
    M3D_CUDA(ctx, cudaMemsetAsync(ds, 0, sizeof(RegSmall), ctx->stream));
    M3D_CUDA(ctx, cudaMemsetAsync(ds->mn, 0xff, sizeof(ds->mn), ctx->stream));

    RegArgs a{};
    a.src = ctx->d_tmp0.as<double>();
    a.dst = ctx->d_tmp1.as<double>();
    a.c0 = ctx->d_tmp2.as<uint32_t>();
    a.c1 = a.c0 + m;
    a.m = (uint32_t)m;
    a.pq = ctx->d_tmp3.as<float4>();
    a.meta = &ds->meta;
    a.picks = ctx->d_samples.as<uint32_t>();
    a.thr = threshold;
    a.thr2 = threshold * threshold; /* max_dis2 */
    a.edge_thr = edge_thr;
    char *w = ctx->d_tmp4.as<char>();
    a.T = reinterpret_cast<double *>(w);
    a.err2 = reinterpret_cast<double *>(w + sizeof(double) * 16 * (size_t)kWaveMax);
    a.counts = reinterpret_cast<uint32_t *>(w + sizeof(double) * 17 * (size_t)kWaveMax);
    a.list = a.counts + kWaveMax;
    a.pass = reinterpret_cast<uint8_t *>(a.list + kWaveMax);
    a.list_count = &ds->list_count;
    a.queue_count = &ds->queue_count;
    a.queue = reinterpret_cast<uint2 *>(ctx->d_queue.as<char>() + 16);
    a.queue_cap = kQueueCap;

    const int gb = ctx->sm_count * 4;
    reg_bbox_kernel<<<gb, 256, 0, ctx->stream>>>(a, ds->mn, ds->mx);
    M3D_LAUNCHED(ctx);
    reg_meta_kernel<<<1, 32, 0, ctx->stream>>>(ds->mn, ds->mx, &ds->meta);
    M3D_LAUNCHED(ctx);
    reg_gather_kernel<<<gb, 256, 0, ctx->stream>>>(a, ctx->d_tmp3.as<float4>());
    M3D_LAUNCHED(ctx);

    const size_t smem = (size_t)kRegStages * kRegTile * 2 * sizeof(float4) + kRegStages * sizeof(uint64_t);
    M3D_CUDA(ctx, cudaFuncSetAttribute(reg_score_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const uint32_t ntiles = ((uint32_t)m + kRegTile - 1) / kRegTile;

    /* (count, rmse) of one hypothesis for tie-breaks: T of row j of the current wave or the best so far */
    std::vector<double> bestT(16, 0.0);
    auto eval_T = [&](const double *hT, bool exact, double *rmse, uint64_t expect) -> int {
        M3D_CUDA(ctx, cudaMemcpyAsync(ds->T, hT, sizeof(double) * 16, cudaMemcpyHostToDevice, ctx->stream));
        if (exact) {
            reg_seq_eval_kernel<<<1, 32, 0, ctx->stream>>>(a, ds->T, ds->eval);
            M3D_LAUNCHED(ctx);
        } else {
            reg_eval_kernel<<<256, 256, 0, ctx->stream>>>(a, ds->T, ctx->d_part.as<double>());
            M3D_LAUNCHED(ctx);
            reg_eval_final_kernel<<<1, 32, 0, ctx->stream>>>(ctx->d_part.as<double>(), 256, ds->eval);
            M3D_LAUNCHED(ctx);
        }
        M3D_CUDA(ctx, cudaMemcpyAsync(hs->eval, ds->eval, sizeof(double) * 2, cudaMemcpyDeviceToHost, ctx->stream));
        M3D_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        if ((uint64_t)hs->eval[0] != expect)
            return ctx->fail(M3D_ERR_INTERNAL, "registration: scoring kernel counted %llu, fp64 pass %llu",
                             (unsigned long long)expect, (unsigned long long)hs->eval[0]);
        *rmse = expect ? std::sqrt(hs->eval[1] / (double)expect) : 0.0;
        return 0;
    };

    std::mt19937 rng(seed);
    std::uniform_int_distribution<int> dist(0, (int)m - 1);
    double best_fit = 0, best_rmse = 0;
    bool best_rmse_known = true, best_rmse_accurate = true, best_rmse_exact = true, found = false;
    int est_k = max_iter;
    bool stopped = false;
    uint32_t wave = (confidence >= 1.0) ? kWaveMax : 1024;
    float score_ms = 0;
    std::vector<double> waveT;
    int done = 0;
    while (done < max_iter && !stopped) {
        const uint32_t rows = (uint32_t)std::min<int64_t>(wave, (int64_t)max_iter - done);
        uint32_t *hp = ctx->h_samples.as<uint32_t>();
        for (uint32_t r = 0; r < 3 * rows; ++r) hp[r] = (uint32_t)dist(rng);
        M3D_CUDA(ctx, cudaMemcpyAsync(ctx->d_samples.p, hp, sizeof(uint32_t) * 3 * rows, cudaMemcpyHostToDevice, ctx->stream));
        M3D_CUDA(ctx, cudaMemsetAsync(&ds->list_count, 0, 2 * sizeof(uint32_t), ctx->stream));
        a.rows = rows;
        reg_solve_kernel<<<(rows + 127) / 128, 128, 0, ctx->stream>>>(a);
        M3D_LAUNCHED(ctx);
        /* survivors are unknown on the host: size the grid for `rows`, idle CTAs exit at once */
        const uint32_t hb = (rows + kRegThreads - 1) / kRegThreads;
        uint32_t chunks = std::max<uint32_t>(1, std::min<uint32_t>(ntiles, (uint32_t)(8 * ctx->sm_count * 4) / std::max(1u, hb / 3 + 1)));
        a.chunk_tiles = (ntiles + chunks - 1) / chunks;
        chunks = (ntiles + a.chunk_tiles - 1) / a.chunk_tiles;
        M3D_CUDA(ctx, cudaEventRecord(ctx->ev[2], ctx->stream));
        reg_score_kernel<<<dim3(hb, chunks), kRegThreads, smem, ctx->stream>>>(a);
        M3D_LAUNCHED(ctx);
        reg_resolve_kernel<<<ctx->sm_count * 2, 256, 0, ctx->stream>>>(a);
        M3D_LAUNCHED(ctx);
        M3D_CUDA(ctx, cudaEventRecord(ctx->ev[3], ctx->stream));
        double *herr = ctx->h_counts.as<double>();
        uint32_t *hcnt = reinterpret_cast<uint32_t *>(herr + kWaveMax);
        uint8_t *hpass = reinterpret_cast<uint8_t *>(hcnt + kWaveMax);
        M3D_CUDA(ctx, cudaMemcpyAsync(herr, a.err2, sizeof(double) * rows, cudaMemcpyDeviceToHost, ctx->stream));
        M3D_CUDA(ctx, cudaMemcpyAsync(hcnt, a.counts, sizeof(uint32_t) * rows, cudaMemcpyDeviceToHost, ctx->stream));
        M3D_CUDA(ctx, cudaMemcpyAsync(hpass, a.pass, rows, cudaMemcpyDeviceToHost, ctx->stream));
        M3D_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        float ms = 0;
        cudaEventElapsedTime(&ms, ctx->ev[2], ctx->ev[3]);
        score_ms += ms;

        auto fetch_T = [&](uint32_t r, double *hT) -> int {
            M3D_CUDA(ctx, cudaMemcpyAsync(hT, a.T + (size_t)r * 16, sizeof(double) * 16, cudaMemcpyDeviceToHost, ctx->stream));
            M3D_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
            return 0;
        };
        for (uint32_t r = 0; r < rows; ++r) {
            const int itr = done + (int)r;
            if (!(itr < est_k)) { /* every later iteration is skipped as well (est_k never grows) */
                stopped = true;
                st.stop_index = (uint64_t)itr;
                break;
            }
            if (!hpass[r]) continue;
            st.evaluated++;
            const uint64_t good = hcnt[r];
            if (good == 0) continue; /* fitness 0, rmse 0: never better than (0, 0) or a found model */
            const double fitness = (double)good / (double)m;
            bool better = false;
            double mine = 0;
            bool mine_known = false, mine_exact = false, mine_accurate = false;
            if (fitness > best_fit) {
                better = true;
                mine = std::sqrt(std::max(herr[r], 0.0) / (double)good); /* kernel estimate, see below */
                mine_known = true;
            } else if (fitness == best_fit) {
                /* inlier-count tie (frequent here): rank by rmse.  The scoring kernel's sum of d^2 is
                 * good to ~2e-6 relative; only closer calls are re-evaluated in fp64 (parallel sum,
                 * then the reference's index-order sum if still too close to call) */
                mine = std::sqrt(std::max(herr[r], 0.0) / (double)good);
                mine_known = true;
                double theirs = best_rmse;
                if (std::fabs(mine - theirs) <= 1e-5 * std::max(mine, theirs) || !best_rmse_known) {
                    double hT[16];
                    if (int rc = fetch_T(r, hT)) return rc;
                    if (!best_rmse_accurate) {
                        if (int rc = eval_T(bestT.data(), false, &theirs, st.best_count)) return rc;
                        best_rmse_accurate = true;
                        best_rmse_exact = false;
                    }
                    if (int rc = eval_T(hT, false, &mine, good)) return rc;
                    mine_accurate = true;
                    if (std::fabs(mine - theirs) <= 1e-9 * std::max(mine, theirs)) {
                        if (!best_rmse_exact)
                            if (int rc = eval_T(bestT.data(), true, &theirs, st.best_count)) return rc;
                        if (int rc = eval_T(hT, true, &mine, good)) return rc;
                        best_rmse_exact = true;
                        mine_exact = true;
                    }
                    best_rmse = theirs;
                }
                better = mine < theirs;
            }
            if (better) {
                best_fit = fitness;
                if (int rc = fetch_T(r, bestT.data())) return rc;
                st.best_index = (uint64_t)itr;
                st.best_count = good;
                best_rmse_known = mine_known;
                best_rmse_accurate = mine_accurate;
                best_rmse_exact = mine_exact;
                if (mine_known) best_rmse = mine;
                found = true;
                est_k = reg_update_limit(good, m, confidence, est_k);
            }
        }
        done += (int)rows;
        if (wave < kWaveMax) wave = std::min<uint32_t>(kWaveMax, wave * 4);
    }
    if (found) {
        memcpy(T_out, bestT.data(), sizeof(double) * 16);
        if (!best_rmse_accurate) {
            double r = 0;
            if (int rc = eval_T(bestT.data(), false, &r, st.best_count)) return rc;
            best_rmse = r;
        }
        st.best_rmse = best_rmse;
    }
    M3D_CUDA(ctx, cudaEventRecord(ctx->ev[1], ctx->stream));
    M3D_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    cudaEventElapsedTime(&st.device_ms, ctx->ev[0], ctx->ev[1]);
    st.score_ms = score_ms;
    if (stats) *stats = st;
    return 1;
}

int m3d_least_squares_transform(m3d_ctx *ctx, const double *src_xyz, const double *dst_xyz, size_t n,
                                int with_scaling, double *T_out) {
    if (!ctx || !T_out) return M3D_ERR_INVALID_ARG;
    if (n < 3) return ctx->fail(M3D_ERR_TOO_FEW_POINTS, "The number of points pair is less than 3."); /* transform_estimation.cpp:29-31 */
    if (!src_xyz || !dst_xyz) return ctx->fail(M3D_ERR_INVALID_ARG, "null point array");
    if (n >= (1ull << 32)) return ctx->fail(M3D_ERR_INVALID_ARG, "too many points");
    M3D_CUDA(ctx, cudaSetDevice(ctx->device));
    const int nb = std::max(1, std::min<int>(ctx->sm_count * 4, (int)((n + 255) / 256)));
    M3D_CUDA(ctx, ctx->d_tmp0.reserve(sizeof(double) * 3 * n));
    M3D_CUDA(ctx, ctx->d_tmp1.reserve(sizeof(double) * 3 * n));
    M3D_CUDA(ctx, ctx->d_part.reserve(sizeof(double) * (10 * (size_t)nb + 32)));
    double *part = ctx->d_part.as<double>();
    double *mean = part + 10 * (size_t)nb, *dT = mean + 8;
    M3D_CUDA(ctx, cudaMemcpyAsync(ctx->d_tmp0.p, src_xyz, sizeof(double) * 3 * n, cudaMemcpyHostToDevice, ctx->stream));
    M3D_CUDA(ctx, cudaMemcpyAsync(ctx->d_tmp1.p, dst_xyz, sizeof(double) * 3 * n, cudaMemcpyHostToDevice, ctx->stream));
    if (int rc = lsq_on_device(ctx, ctx->d_tmp0.as<double>(), ctx->d_tmp1.as<double>(), (uint32_t)n, nb, with_scaling, part)) return rc;
    M3D_CUDA(ctx, cudaMemcpyAsync(T_out, dT, sizeof(double) * 16, cudaMemcpyDeviceToHost, ctx->stream));
    M3D_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return M3D_OK;
}

int m3d_registration_refit(m3d_ctx *ctx, const double *src_xyz, size_t ns, const double *dst_xyz, size_t nd,
                           const size_t *c0, const size_t *c1, size_t m, const double *T_in, double threshold,
                           int with_scaling, double *T_out, size_t *n_inliers) {
    if (!ctx || !T_in || !T_out) return M3D_ERR_INVALID_ARG;
    for (int i = 0; i < 16; ++i) T_out[i] = T_in[i];
    if (n_inliers) *n_inliers = 0;
    if (m == 0) return M3D_OK;
    if (!src_xyz || !dst_xyz || !c0 || !c1) return ctx->fail(M3D_ERR_INVALID_ARG, "null array");
    if (!(threshold > 0.0)) return ctx->fail(M3D_ERR_INVALID_ARG, "threshold must be positive");
    if (m >= (1ull << 31) || ns >= (1ull << 32) || nd >= (1ull << 32))
        return ctx->fail(M3D_ERR_INVALID_ARG, "too many points / correspondences");
    for (size_t i = 0; i < m; ++i)
        if (c0[i] >= ns || c1[i] >= nd) return ctx->fail(M3D_ERR_INVALID_ARG, "correspondence %zu out of range", i);
    M3D_CUDA(ctx, cudaSetDevice(ctx->device));
    const int nb = std::max(1, std::min<int>(ctx->sm_count * 4, (int)((m + 255) / 256)));
    M3D_CUDA(ctx, ctx->d_tmp0.reserve(sizeof(double) * 3 * m)); /* selected source points      */
    M3D_CUDA(ctx, ctx->d_tmp1.reserve(sizeof(double) * 3 * m)); /* selected destination points */
    M3D_CUDA(ctx, ctx->d_tmp2.reserve(sizeof(uint32_t) * 2 * m));
    M3D_CUDA(ctx, ctx->d_tmp3.reserve(sizeof(double) * 3 * ns));
    M3D_CUDA(ctx, ctx->d_tmp5.reserve(sizeof(double) * 3 * nd));
    M3D_CUDA(ctx, ctx->d_part.reserve(sizeof(double) * (10 * (size_t)nb + 64)));
    M3D_CUDA(ctx, ctx->h_small.reserve(64));
    double *part = ctx->d_part.as<double>();
    double *mean = part + 10 * (size_t)nb, *dT = mean + 8, *dTin = dT + 16;
    uint32_t *d_count = reinterpret_cast<uint32_t *>(dTin + 16);
    std::vector<uint32_t> hc(2 * m);
    for (size_t i = 0; i < m; ++i) {
        hc[i] = (uint32_t)c0[i];
        hc[m + i] = (uint32_t)c1[i];
    }
    M3D_CUDA(ctx, cudaMemcpyAsync(ctx->d_tmp3.p, src_xyz, sizeof(double) * 3 * ns, cudaMemcpyHostToDevice, ctx->stream));
    M3D_CUDA(ctx, cudaMemcpyAsync(ctx->d_tmp5.p, dst_xyz, sizeof(double) * 3 * nd, cudaMemcpyHostToDevice, ctx->stream));
    M3D_CUDA(ctx, cudaMemcpyAsync(ctx->d_tmp2.p, hc.data(), sizeof(uint32_t) * 2 * m, cudaMemcpyHostToDevice, ctx->stream));
    M3D_CUDA(ctx, cudaMemcpyAsync(dTin, T_in, sizeof(double) * 16, cudaMemcpyHostToDevice, ctx->stream));
    refit_select_kernel<<<1, 1024, 0, ctx->stream>>>(ctx->d_tmp3.as<double>(), ctx->d_tmp5.as<double>(),
                                                     ctx->d_tmp2.as<uint32_t>(), ctx->d_tmp2.as<uint32_t>() + m, (uint32_t)m,
                                                     dTin, threshold * threshold, ctx->d_tmp0.as<double>(),
                                                     ctx->d_tmp1.as<double>(), d_count);
    M3D_LAUNCHED(ctx);
    uint32_t *h_count = ctx->h_small.as<uint32_t>();
    M3D_CUDA(ctx, cudaMemcpyAsync(h_count, d_count, sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
    M3D_CUDA(ctx, cudaStreamSynchronize(ctx->stream)); /* also: hc may go out of scope now */
    const uint32_t k = *h_count;
    if (n_inliers) *n_inliers = k;
    if (k < 3) return M3D_OK; /* nothing to estimate from: the transformation is returned unchanged */
    const int nbk = std::max(1, std::min<int>(nb, (int)((k + 255) / 256)));
    if (int rc = lsq_on_device(ctx, ctx->d_tmp0.as<double>(), ctx->d_tmp1.as<double>(), k, nbk, with_scaling, part)) return rc;
    M3D_CUDA(ctx, cudaMemcpyAsync(T_out, part + 10 * (size_t)nbk + 8, sizeof(double) * 16, cudaMemcpyDeviceToHost, ctx->stream));
    M3D_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return M3D_OK;
}

} /* extern "C" */
