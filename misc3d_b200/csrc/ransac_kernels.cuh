/*
 * ransac_kernels.cuh -- sm_100a kernels of the RANSAC primitive-fitting path.
 *
 *   cloud preparation   bbox_kernel / bbox_final_kernel / convert_kernel
 *   hot kernel          score_kernel<KIND,THREADS,HPT>   (sample gather -> minimal solve ->
 *                       all-point inlier count, fp32 guard-banded + fp64 reference-order resolve)
 *   reference-order     score_exact_kernel<KIND>         (fp64 only; debug / non-finite clouds)
 *   RefineModel         refine_count / refine_scan / refine_write / refine_final
 *   helpers             minimal_fit_rows_kernel, seq_err_kernel
 *
 * Data layout in HBM: xyz  N x 3 f64 AoS (the reference's vector<Vector3d>),
 * pts32 N x float4 {x-cx, y-cy, z-cz, |p-c|^2} (16 B/pt, what the hot kernel streams through
 * shared memory with 1-D TMA bulk copies), samples rows x k u32, counts rows u32.
 */
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

#include "context.h"
#include "exact_math.cuh"
#include "scan.h"

namespace m3d {

constexpr int kTile = 1024;  /* points per TMA stage (16 KB)                      */
constexpr int kStages = 3;   /* ring depth                                        */
#ifndef M3D_UNROLL
#define M3D_UNROLL 32
#endif
constexpr int kUnroll = M3D_UNROLL;
constexpr int kSub = 32; /* points between two "any point inside the band?" checks = one warp-wide rescan */
constexpr uint32_t kInvalidBit = 0x80000000u;
constexpr double kU32 = 5.9604644775390625e-08; /* 2^-24 */
constexpr double kU64 = 1.1102230246251565e-16; /* 2^-53 */

/* ------------------------------------------------------------------ mbarrier / TMA (1-D bulk) */
__device__ __forceinline__ uint32_t smem_u32(const void *p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "LAB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra DONE;\n"
        "bra LAB_WAIT;\n"
        "DONE:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
/* the producer thread is whole tiles ahead of the consumers: it waits with a suspend-time hint, i.e. the
 * hardware parks the thread until the phase completes (or the hint expires) instead of letting it spin.
 * (Round 1 polled try_wait + __nanosleep(256): ncu showed that loop issuing 25 % of all warp instructions
 * of the kernel -- nanosleep returned almost immediately -- on the schedulers the consumer warps need.) */
__device__ __forceinline__ void mbar_wait_relaxed(uint64_t *bar, uint32_t parity) {
    const uint32_t b = smem_u32(bar);
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "LAB_WAIT_R:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1, %2;\n"
        "@P1 bra DONE_R;\n"
        "bra LAB_WAIT_R;\n"
        "DONE_R:\n"
        "}\n" ::"r"(b),
        "r"(parity), "r"(1000000u)
        : "memory");
}
/* cp.async.bulk (TMA, SASS UBLKCP): global -> shared, completion on an mbarrier */
__device__ __forceinline__ void tma_load_1d(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
    const uint32_t b = smem_u32(bar);
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(bytes) : "memory");
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
            smem_u32(dst)),
        "l"(src), "r"(bytes), "r"(b)
        : "memory");
}

/* ------------------------------------------------------------------------- cloud preparation */
struct BBoxPart {
    double mn[3], mx[3], mraw;
    int nonfinite;
    int pad;
};

__device__ __forceinline__ double warp_min(double v) {
    for (int o = 16; o; o >>= 1) v = fmin(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ double warp_max(double v) {
    for (int o = 16; o; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ double warp_sum(double v) {
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__global__ void __launch_bounds__(256) bbox_kernel(const double *__restrict__ xyz, uint32_t n,
                                                   BBoxPart *__restrict__ part) {
    double mn[3] = {INFINITY, INFINITY, INFINITY}, mx[3] = {-INFINITY, -INFINITY, -INFINITY};
    double mraw = 0;
    int bad = 0;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const double v = xyz[3 * (size_t)i + c];
            bad |= !isfinite(v);
            mn[c] = fmin(mn[c], v);
            mx[c] = fmax(mx[c], v);
            mraw = fmax(mraw, fabs(v));
        }
    }
    __shared__ BBoxPart sh[8];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        mn[c] = warp_min(mn[c]);
        mx[c] = warp_max(mx[c]);
    }
    mraw = warp_max(mraw);
    bad = __any_sync(0xffffffffu, bad);
    if (lane == 0) {
        for (int c = 0; c < 3; ++c) {
            sh[w].mn[c] = mn[c];
            sh[w].mx[c] = mx[c];
        }
        sh[w].mraw = mraw;
        sh[w].nonfinite = bad;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        BBoxPart r = sh[0];
        for (int k = 1; k < 8; ++k) {
            for (int c = 0; c < 3; ++c) {
                r.mn[c] = fmin(r.mn[c], sh[k].mn[c]);
                r.mx[c] = fmax(r.mx[c], sh[k].mx[c]);
            }
            r.mraw = fmax(r.mraw, sh[k].mraw);
            r.nonfinite |= sh[k].nonfinite;
        }
        part[blockIdx.x] = r;
    }
}

__global__ void __launch_bounds__(256) bbox_final_kernel(const BBoxPart *__restrict__ part, int nparts, CloudMeta *meta) {
    double mn[3] = {INFINITY, INFINITY, INFINITY}, mx[3] = {-INFINITY, -INFINITY, -INFINITY};
    double mraw = 0;
    int bad = 0;
    for (int k = threadIdx.x; k < nparts; k += blockDim.x) {
        const BBoxPart p = part[k];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            mn[c] = fmin(mn[c], p.mn[c]);
            mx[c] = fmax(mx[c], p.mx[c]);
        }
        mraw = fmax(mraw, p.mraw);
        bad |= p.nonfinite;
    }
    __shared__ BBoxPart sh[8];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        mn[c] = warp_min(mn[c]);
        mx[c] = warp_max(mx[c]);
    }
    mraw = warp_max(mraw);
    bad = __any_sync(0xffffffffu, bad);
    if (lane == 0) {
        for (int c = 0; c < 3; ++c) {
            sh[w].mn[c] = mn[c];
            sh[w].mx[c] = mx[c];
        }
        sh[w].mraw = mraw;
        sh[w].nonfinite = bad;
    }
    __syncthreads();
    if (threadIdx.x != 0) return;
    BBoxPart r = sh[0];
    for (int k = 1; k < 8; ++k) {
        for (int c = 0; c < 3; ++c) {
            r.mn[c] = fmin(r.mn[c], sh[k].mn[c]);
            r.mx[c] = fmax(r.mx[c], sh[k].mx[c]);
        }
        r.mraw = fmax(r.mraw, sh[k].mraw);
        r.nonfinite |= sh[k].nonfinite;
    }
    /* mc = max |centred coordinate| over the cloud: attained at a face of the bounding box (x - c is monotone in x,
     * and so is its rounding), hence computable here without another pass */
    double mc = 0;
    for (int c = 0; c < 3; ++c) {
        const double ctr0 = 0.5 * (r.mn[c] + r.mx[c]);
        const double ctr = isfinite(ctr0) ? ctr0 : 0.0;
        meta->center[c] = ctr;
        mc = fmax(mc, fmax(fabs(r.mn[c] - ctr), fabs(r.mx[c] - ctr)));
    }
    meta->mraw = r.mraw;
    meta->nonfinite = r.nonfinite;
    meta->mc = isfinite(mc) ? mc : 0.0;
}

/* pts32[i] = {x-cx, y-cy, z-cz, |p-c|^2} */
__global__ void __launch_bounds__(256) convert_kernel(const double *__restrict__ xyz, uint32_t n,
                                                      const CloudMeta *__restrict__ meta,
                                                      float4 *__restrict__ pts32) {
    const double cx = meta->center[0], cy = meta->center[1], cz = meta->center[2];
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const double x = xyz[3 * (size_t)i] - cx, y = xyz[3 * (size_t)i + 1] - cy,
                     z = xyz[3 * (size_t)i + 2] - cz;
        pts32[i] = make_float4((float)x, (float)y, (float)z, (float)(x * x + y * y + z * z));
    }
}

/* 32-bit cluster labels -> the size_t labels of the 64-bit entry point (0xFFFFFFFF -> UINT64_MAX) */
__global__ void widen_labels_kernel(const uint32_t *__restrict__ in, unsigned long long *__restrict__ out, uint32_t n) {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const uint32_t v = in[i];
        out[i] = v == 0xffffffffu ? ~0ull : (unsigned long long)v;
    }
}

__global__ void iota_kernel(uint32_t *out, uint32_t n) {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) out[i] = i;
}

/* --------------------------------------------------------------- sample gather + minimal fit */
/* RandomSampler row (draw order) -> SelectByIndex order (ascending, ransac.h:578 / Open3D mask
 * pass) -> MinimalFit.  Returns MinimalFit's bool. */
template <int KIND>
__device__ __forceinline__ bool fit_row(const double *__restrict__ xyz, const double *__restrict__ nrm,
                                        const uint32_t *__restrict__ samples, uint32_t row, double *m,
                                        const double *__restrict__ row_nrm = nullptr) {
    /* row_nrm (optional): the normals of exactly the sampled points, [row][draw position][3] -- what the
     * host-buffer entry point uploads instead of the whole normal array (only the cylinder's two sample
     * normals are ever read, ransac.h:376-383) */
    constexpr int K = sample_size(KIND);
    uint32_t s[K];
    int ord[K];
#pragma unroll
    for (int i = 0; i < K; ++i) {
        s[i] = samples[(size_t)row * K + i];
        ord[i] = i;
    }
#pragma unroll
    for (int i = 1; i < K; ++i) /* K <= 4: sorting network by insertion */
#pragma unroll
        for (int j = i; j > 0; --j)
            if (s[j - 1] > s[j]) {
                const uint32_t t = s[j];
                s[j] = s[j - 1];
                s[j - 1] = t;
                const int o = ord[j];
                ord[j] = ord[j - 1];
                ord[j - 1] = o;
            }
    double pts[3 * K];
    double nr[KIND == kCylinder ? 3 * K : 1];
#pragma unroll
    for (int i = 0; i < K; ++i)
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            pts[3 * i + c] = xyz[3 * (size_t)s[i] + c];
            if (KIND == kCylinder)
                nr[3 * i + c] = row_nrm ? row_nrm[((size_t)row * K + ord[i]) * 3 + c] : nrm[3 * (size_t)s[i] + c];
        }
    return ex::minimal_fit<KIND>(pts, nr, m);
}

template <int KIND>
__global__ void minimal_fit_rows_kernel(const double *__restrict__ xyz, const double *__restrict__ nrm,
                                        const uint32_t *__restrict__ samples, uint32_t rows,
                                        double *__restrict__ models, uint8_t *__restrict__ valid,
                                        const double *__restrict__ row_nrm) {
    const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= rows) return;
    double m[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    const bool ok = fit_row<KIND>(xyz, nrm, samples, r, m, row_nrm);
#pragma unroll
    for (int i = 0; i < 8; ++i) models[(size_t)r * 8 + i] = ok ? m[i] : 0.0;
    valid[r] = ok ? 1 : 0;
}

/* ------------------------------------------------------------------ fp32 guard-banded scoring */
/* Fast<KIND>: for a centred fp32 point p = {x,y,z,|p|^2} let v = |t(p)| - T.  Then
 *      v <= -band  =>  the reference predicate `distance < threshold` is certainly true
 *      v >=  band  =>  certainly false
 * and a point with |v| < band is decided by the fp64 reference-order distance.  The inner loop
 * counts sign(v) (provisional decision) and tracks min |v|; only when min |v| < band is a
 * sub-tile looked at again.
 *   plane     t = w.p + w3'                      (T = threshold*||w||)
 *   sphere    t = |p - c|^2 - mid                (T = half; [mid-half, mid+half] is the
 *   cylinder  t = |p - c|^2 - ((p-c).n)^2 - mid   interval of squared distances (r-thr)^2..(r+thr)^2)
 * Per point-hypothesis pair: 3 / 4 / 8 FFMA-pipe ops + FADD + LEA.HI + FMNMX.
 */
template <int KIND>
struct Fast {
    float c[KIND == kCylinder ? 8 : 4];
    float T, band;
};
/* v = |t| - T (one FADD with the |.| operand modifier) */
template <int KIND>
__device__ __forceinline__ float fast_v(const Fast<KIND> &f, const float4 p);

template <int KIND>
__device__ __forceinline__ float fast_eval(const Fast<KIND> &f, const float4 p) {
    if (KIND == kPlane) {
        return fmaf(f.c[0], p.x, fmaf(f.c[1], p.y, fmaf(f.c[2], p.z, f.c[3])));
    } else if (KIND == kSphere) {
        return fmaf(f.c[0], p.x, fmaf(f.c[1], p.y, fmaf(f.c[2], p.z, __fadd_rn(p.w, f.c[3]))));
    } else {
        const float a = fmaf(f.c[0], p.x, fmaf(f.c[1], p.y, fmaf(f.c[2], p.z, __fadd_rn(p.w, f.c[3]))));
        const float b = fmaf(f.c[4], p.x, fmaf(f.c[5], p.y, fmaf(f.c[6], p.z, f.c[7])));
        return fmaf(-b, b, a);
    }
}

template <int KIND>
__device__ __forceinline__ float fast_v(const Fast<KIND> &f, const float4 p) {
    return __fsub_rn(fabsf(fast_eval<KIND>(f, p)), f.T);
}

/* ---- packed fp32x2 form (Blackwell FFMA2 / FADD2): one thread evaluates TWO hypotheses at the
 * same point with one packed instruction per term.  `fma.rn.f32x2` / `add.rn.f32x2` round each half
 * exactly like the scalar fmaf / __fadd_rn in fast_eval, in the same order, so the value a lane
 * computes here is bit-identical to what rescan_warp recomputes with the scalar form.  ptxas folds
 * the point coordinate into the broadcast operand form (`R.F32`) and the negation into `-R.F32x2`,
 * so packing costs no extra instruction.  M3D_PACKED=0 keeps the scalar inner loop. */
#ifndef M3D_PACKED
#define M3D_PACKED 1
#endif
typedef unsigned long long f32x2_t;
__device__ __forceinline__ f32x2_t pack2(float lo, float hi) {
    f32x2_t r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void unpack2(f32x2_t r, float &lo, float &hi) {
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(r));
}
__device__ __forceinline__ f32x2_t ffma2(f32x2_t a, f32x2_t b, f32x2_t c) {
    f32x2_t d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
__device__ __forceinline__ f32x2_t fadd2(f32x2_t a, f32x2_t b) {
    f32x2_t d;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ f32x2_t fneg2(f32x2_t a) {
    float lo, hi;
    unpack2(a, lo, hi);
    return pack2(-lo, -hi);
}
template <int KIND>
struct Fast2 {
    f32x2_t c[KIND == kCylinder ? 8 : 4]; /* {hypothesis 2q, hypothesis 2q+1} */
};
template <int KIND>
__device__ __forceinline__ void pack_fast(const Fast<KIND> &f0, const Fast<KIND> &f1, Fast2<KIND> &g) {
    constexpr int NC = KIND == kCylinder ? 8 : 4;
#pragma unroll
    for (int i = 0; i < NC; ++i) g.c[i] = pack2(f0.c[i], f1.c[i]);
}
template <int KIND>
__device__ __forceinline__ void unpack_fast(const Fast2<KIND> &g, int half, float T, float band, Fast<KIND> &f) {
    constexpr int NC = KIND == kCylinder ? 8 : 4;
#pragma unroll
    for (int i = 0; i < NC; ++i) {
        float lo, hi;
        unpack2(g.c[i], lo, hi);
        f.c[i] = half ? hi : lo;
    }
    f.T = T;
    f.band = band;
}
/* t of both hypotheses of the pair; same operation order as fast_eval */
template <int KIND>
__device__ __forceinline__ void fast_eval2(const Fast2<KIND> &g, const float4 p, float &t0, float &t1) {
    const f32x2_t px = pack2(p.x, p.x), py = pack2(p.y, p.y), pz = pack2(p.z, p.z);
    f32x2_t t;
    if (KIND == kPlane) {
        t = ffma2(g.c[0], px, ffma2(g.c[1], py, ffma2(g.c[2], pz, g.c[3])));
    } else if (KIND == kSphere) {
        t = ffma2(g.c[0], px, ffma2(g.c[1], py, ffma2(g.c[2], pz, fadd2(pack2(p.w, p.w), g.c[3]))));
    } else {
        const f32x2_t a = ffma2(g.c[0], px, ffma2(g.c[1], py, ffma2(g.c[2], pz, fadd2(pack2(p.w, p.w), g.c[3]))));
        const f32x2_t b = ffma2(g.c[4], px, ffma2(g.c[5], py, ffma2(g.c[6], pz, g.c[7])));
        t = ffma2(fneg2(b), b, a);
    }
    unpack2(t, t0, t1);
}

/* inner-loop bookkeeping: provisional inlier count (sign bit of v) and min |v| */
__device__ __forceinline__ void accumulate_v(float v, uint32_t &clo, float &mn) {
    clo += __float_as_uint(v) >> 31;
    mn = fminf(mn, fabsf(v));
}

template <int KIND>
__device__ inline void make_fast(const double *m, bool ok, const CloudMeta &M, double thr, Fast<KIND> &f) {
    constexpr int NC = KIND == kCylinder ? 8 : 4;
#pragma unroll
    for (int i = 0; i < NC; ++i) f.c[i] = 0.f;
    f.T = -1.f;   /* v = |t| + 1 >= 1: never an inlier ...                                      */
    f.band = 0.f; /* ... and never inside the band (t is finite: p and the coefficients are)    */
    if (!ok || !(thr > 0)) return;
    bool fin = true;
#pragma unroll
    for (int i = 0; i < param_count(KIND); ++i) fin = fin && isfinite(m[i]);
    if (!fin) return; /* a NaN/inf model makes every reference distance NaN/inf: never an inlier */

    double c[NC], Tc, bd, smax = 0;
    const double cx = M.center[0], cy = M.center[1], cz = M.center[2];
    if (KIND == kPlane) {
        const double nrm = ex::plane_norm(m);
        const double thrn = thr * nrm;
        const double w3c = m[3] + (m[0] * cx + m[1] * cy + m[2] * cz);
        const double l1 = fabs(m[0]) + fabs(m[1]) + fabs(m[2]);
        const double S = l1 * M.mc + fabs(w3c);
        const double band = 16 * kU32 * S + 4 * kU32 * thrn + 16 * kU64 * (l1 * M.mraw + fabs(m[3]));
        smax = S;
        c[0] = m[0];
        c[1] = m[1];
        c[2] = m[2];
        c[3] = w3c;
        Tc = thrn;
        bd = band;
    } else {
        const double r = (KIND == kSphere) ? m[3] : m[6];
        const double Hi = (r + thr) * (r + thr);
        const double Lo = (r >= thr) ? (r - thr) * (r - thr) : -Hi;
        const double mid = 0.5 * (Lo + Hi), half = 0.5 * (Hi - Lo);
        const double rh = sqrt(Hi);
        double ax = m[0] - cx, ay = m[1] - cy, az = m[2] - cz; /* centre / axis point, centred */
        double band;
        if (KIND == kSphere) {
            const double c3 = (ax * ax + ay * ay + az * az) - mid;
            const double l1 = fabs(ax) + fabs(ay) + fabs(az);
            const double S = 3 * M.mc * M.mc + fabs(c3) + 2 * l1 * M.mc;
            band = 24 * kU32 * S + 4 * kU32 * half +
                   2 * rh * 16 * kU64 * (M.mraw + fabs(m[0]) + fabs(m[1]) + fabs(m[2]) + rh);
            smax = S;
            c[3] = c3;
        } else {
            /* the reference measures the distance to the line through `center` and
             * `center + dir` (ransac.h:438-442): direction = (center + dir) - center */
            const ex::V3 cen = {m[0], m[1], m[2]};
            const ex::V3 ref = {ex::add(m[0], m[3]), ex::add(m[1], m[4]), ex::add(m[2], m[5])};
            const ex::V3 ne = ex::sub3(ref, cen);
            const double L2 = ne.x * ne.x + ne.y * ne.y + ne.z * ne.z;
            if (!(L2 > 0) || !isfinite(L2)) return; /* 0/0 = NaN distance: never an inlier */
            const double L = sqrt(L2);
            const double nx = ne.x / L, ny = ne.y / L, nz = ne.z / L;
            const double cnorm = sqrt(ax * ax + ay * ay + az * az);
            /* re-anchor the axis point at the foot of the cloud centre (same line, no
             * cancellation between |c|^2 and (c.n)^2) */
            const double proj = ax * nx + ay * ny + az * nz;
            ax -= proj * nx;
            ay -= proj * ny;
            az -= proj * nz;
            const double c3 = (ax * ax + ay * ay + az * az) - mid;
            const double l1 = fabs(ax) + fabs(ay) + fabs(az);
            const double n3 = -(nx * ax + ny * ay + nz * az);
            const double Sa = 3 * M.mc * M.mc + fabs(c3) + 2 * l1 * M.mc;
            const double Sb = (fabs(nx) + fabs(ny) + fabs(nz)) * M.mc + fabs(n3);
            const double A = 1.7320508075688772 * M.mraw + sqrt(m[0] * m[0] + m[1] * m[1] + m[2] * m[2]) + L;
            band = 32 * kU32 * (Sa + Sb * Sb) + 4 * kU32 * half +
                   2 * rh * 16 * kU64 * (A * A / L + cnorm + M.mc);
            smax = Sa + Sb * Sb;
            c[3] = c3;
            c[4] = nx;
            c[5] = ny;
            c[6] = nz;
            c[7] = n3;
        }
        c[0] = -2 * ax;
        c[1] = -2 * ay;
        c[2] = -2 * az;
        Tc = half;
        bd = band;
    }
    /* smax bounds every intermediate of fast_eval: keep it far from fp32 overflow */
    bool okf = isfinite(Tc) && isfinite(bd) && (smax < 1e30);
#pragma unroll
    for (int i = 0; i < NC; ++i) {
        f.c[i] = (float)c[i];
        okf = okf && isfinite(f.c[i]);
    }
    f.T = (float)Tc;
    /* + the rounding of T itself and of the subtraction |t| - T */
    f.band = __double2float_ru(bd * (1.0 + 8 * kU32) + 2 * kU32 * fabs(Tc));
    okf = okf && isfinite(f.T) && isfinite(f.band);
    if (!okf) { /* fp32 cannot represent this model: decide every point by the fp64 path */
#pragma unroll
        for (int i = 0; i < NC; ++i) f.c[i] = 0.f;
        f.T = 0.f;
        f.band = INFINITY; /* |v| < inf for every (finite) point */
    }
}

struct ScoreArgs {
    const float4 *blob;   /* Morton-ordered tiles + bounding spheres (score_cull.cuh) or null */
    const uint32_t *perm; /* sorted position -> original index (with blob)                    */
    const float4 *pts32;
    const double *xyz;
    const double *nrm;
    const double *row_nrm;   /* optional: normals of the sampled points only, [row][k][3] (see fit_row) */
    const CloudMeta *meta;
    const uint32_t *samples; /* rows x k of this wave (device)                          */
    /* optional: the minimal models of ALL wave rows, computed beforehand ([wave row][8] + MinimalFit flags).  Set by
     * the chunked host-buffer fit: a launch then scores one chunk of the cloud (xyz / pts32 / blob / perm / meta / n
     * describe the chunk) while later chunks are still being uploaded, and the sample points may lie in any chunk. */
    const double *models_in;
    const uint8_t *valid_in;
    uint32_t *counts;        /* [rows]: inlier count (atomicAdd per chunk), bit31 = MinimalFit false */
    unsigned long long *resolves;
    double *models;        /* [rows][8] minimal models (written by the point-chunk-0 CTAs)  */
    uint2 *queue;          /* (local row, point) pairs inside the guard band                */
    uint32_t *queue_count; /* atomic cursor of `queue`                                      */
    uint32_t queue_cap;
    double thr;
    uint32_t n;
    uint32_t row_begin; /* first row of the wave buffer this launch scores             */
    const uint32_t *row_map; /* optional: launch-local row -> shard-local row (a subset of
                              * [row_begin, ...) in any order); null = row_begin + launch-local row     */
    uint32_t flags;          /* M3D_FLAG_* of the call (host side only)                               */
    uint32_t shard_world, shard_rank; /* block-cyclic hypothesis sharding (scan.h ShardMap); world <= 1: none */
    /* shard-local row -> row of the wave buffer (index into samples / row_nrm) */
    __device__ __forceinline__ uint32_t wave_row(uint32_t l) const {
        return shard_world <= 1 ? l : ((l / kShardBlock) * shard_world + shard_rank) * kShardBlock + l % kShardBlock;
    }
    /* wave-buffer row and index into `counts` (shard-local, relative to row_begin) of launch-local row r */
    __device__ __forceinline__ uint32_t src_row(uint32_t r) const { return wave_row(row_map ? row_map[r] : row_begin + r); }
    __device__ __forceinline__ uint32_t cnt_row(uint32_t r) const { return row_map ? row_map[r] - row_begin : r; }
    uint32_t rows;      /* number of rows this launch scores                           */
    uint32_t chunk_tiles;    /* score_exact_kernel only */
    uint32_t *tile_counter;  /* [hypothesis blocks] next unclaimed tile (zeroed before the launch) */
};

/* the rare path, first half: some lane's hypothesis saw a point of this 32-point sub-tile inside
 * its guard band.  The whole warp re-examines the sub-tile for that hypothesis -- lane L takes
 * point L with the owner's coefficients (shuffled) -- and queues the (hypothesis, point,
 * provisional decision) triples for resolve_queue_kernel.  All lanes stay converged. */
template <int KIND>
__device__ __forceinline__ void rescan_warp(const ScoreArgs &a, unsigned need, uint32_t row_local,
                                            const Fast<KIND> &f, const float4 *sp, uint32_t gbase, int cnt,
                                            uint32_t &nres) {
    constexpr int NC = KIND == kCylinder ? 8 : 4;
    const unsigned full = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const float4 p = sp[lane < cnt ? lane : 0];
    while (need) {
        const int src = __ffs(need) - 1;
        need &= need - 1;
        Fast<KIND> g;
#pragma unroll
        for (int i = 0; i < NC; ++i) g.c[i] = __shfl_sync(full, f.c[i], src);
        g.T = __shfl_sync(full, f.T, src);
        g.band = __shfl_sync(full, f.band, src);
        const uint32_t r = __shfl_sync(full, row_local, src);
        const float v = fast_v<KIND>(g, p);
        const bool amb = lane < cnt && fabsf(v) < g.band;
        const unsigned am = __ballot_sync(full, amb);
        if (am == 0) continue;
        uint32_t pos0 = 0;
        if (lane == 0) {
            pos0 = atomicAdd(a.queue_count, (uint32_t)__popc(am));
            nres += __popc(am);
        }
        pos0 = __shfl_sync(full, pos0, 0);
        if (amb) {
            const uint32_t prov = __float_as_uint(v) >> 31; /* what the inner loop counted */
            const uint32_t pos = pos0 + __popc(am & ((1u << lane) - 1));
            if (pos < a.queue_cap) {
                a.queue[pos] = make_uint2(r, (gbase + lane) | (prov << 31));
            } else { /* queue full: decide here with the reference arithmetic */
                double m[8];
                const bool ok = fit_row<KIND>(a.xyz, a.nrm, a.samples, a.src_row(r), m, a.row_nrm);
                uint32_t in = 0;
                if (ok) {
                    ex::Dist<KIND> dist;
                    dist.set(m);
                    in = dist(ex::ld3(a.xyz + 3 * (size_t)(gbase + lane))) < a.thr ? 1u : 0u;
                }
                if (in != prov) atomicAdd(&a.counts[a.cnt_row(r)], in - prov);
            }
        }
    }
}

/* second half: one thread per queued pair evaluates the reference's fp64 predicate with the
 * model the scoring kernel stored and corrects the provisional count */
template <int KIND>
__global__ void __launch_bounds__(256) resolve_queue_kernel(const ScoreArgs a) {
    const uint32_t total = min(*a.queue_count, a.queue_cap);
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const uint2 e = a.queue[i];
        const uint32_t prov = e.y >> 31;
        uint32_t pt = e.y & 0x7fffffffu;
        if (a.blob) pt = a.perm[pt]; /* the culling kernel queues positions in the Morton-ordered copy */
        double m[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) m[k] = a.models[(size_t)e.x * 8 + k];
        ex::Dist<KIND> dist;
        dist.set(m);
        const uint32_t in = (dist(ex::ld3(a.xyz + 3 * (size_t)pt)) < a.thr) ? 1u : 0u;
        if (in != prov) atomicAdd(&a.counts[a.cnt_row(e.x)], in - prov); /* +1 or -1 (mod 2^32) */
    }
}

/* Persistent CTAs.  A CTA = THREADS consumer threads (HPT hypotheses each) + one producer warp.
 * blockIdx.x selects the hypothesis block; all CTAs with the same blockIdx.x pull 1024-point tiles
 * from one atomic cursor (tile_counter[blockIdx.x]) until the cloud is exhausted, so the grid is one
 * resident wave with no tail, and the fp64 prologue runs once per CTA.  The producer warp feeds a
 * kStages-deep TMA ring (full[] / empty[] mbarriers): consumers never meet at a CTA-wide barrier
 * inside the main loop.  Counts are integer sums, so the dynamic tile order does not affect results. */
constexpr uint32_t kNoTile = 0xffffffffu;
template <int KIND, int THREADS, int HPT>
__global__ void __launch_bounds__(THREADS + 32) score_kernel(const ScoreArgs a) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float4 *tiles = reinterpret_cast<float4 *>(smem_raw);
    uint64_t *full = reinterpret_cast<uint64_t *>(smem_raw + (size_t)kStages * kTile * sizeof(float4));
    uint64_t *empty = full + kStages;
    volatile uint32_t *tile_id = reinterpret_cast<volatile uint32_t *>(empty + kStages);

    const int tid = threadIdx.x;
    const uint32_t ntiles = (a.n + kTile - 1) / kTile;

    if (tid == THREADS) {
#pragma unroll
        for (int s = 0; s < kStages; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], THREADS / 32);
        }
        mbar_fence_init();
    }
    __syncthreads(); /* the only CTA-wide barrier: ring barriers are initialised */

    if (tid >= THREADS) { /* ---------------- producer warp: one elected lane claims tiles and issues the bulk copies */
        if (tid == THREADS) {
            for (uint32_t k = 0;; ++k) {
                const int st = k % kStages;
                if (k >= kStages) mbar_wait_relaxed(&empty[st], ((k / kStages) - 1) & 1); /* slot released by all warps */
                const uint32_t t = atomicAdd(&a.tile_counter[blockIdx.x], 1u);
                if (t >= ntiles) {
                    tile_id[st] = kNoTile;
                    mbar_arrive(&full[st]); /* release: consumers see the end marker */
                    break;
                }
                tile_id[st] = t;
                const uint32_t base = t * kTile;
                const uint32_t npt = min((uint32_t)kTile, a.n - base);
                tma_load_1d(tiles + (size_t)st * kTile, a.pts32 + base, npt * (uint32_t)sizeof(float4), &full[st]);
            }
        }
        return;
    }

    /* ---------------- consumers.  Prologue: this thread's hypotheses -- gather, solve (fp64
     * reference order), fp32 coefficients */
    const CloudMeta M = *a.meta;
    uint32_t row[HPT];
    Fast<KIND> f[HPT];
    uint32_t clo[HPT];
    float mn[HPT];
    bool invalid[HPT];
#pragma unroll
    for (int h = 0; h < HPT; ++h) {
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
        make_fast<KIND>(m, ok, M, a.thr, f[h]);
        clo[h] = 0;
        mn[h] = INFINITY;
    }

    /* hypothesis pairs share packed coefficient registers (FFMA2); T and band stay per hypothesis */
    constexpr bool kPacked = (M3D_PACKED != 0) && (HPT % 2 == 0);
    constexpr int NP = kPacked ? HPT / 2 : 1;
    Fast2<KIND> f2[NP];
    if (kPacked) {
#pragma unroll
        for (int q = 0; q < NP; ++q) pack_fast<KIND>(f[2 * q], f[2 * q + 1], f2[q]);
    }
    auto eval_point = [&](const float4 p) {
        if (kPacked) {
#pragma unroll
            for (int q = 0; q < NP; ++q) {
                float t0, t1;
                fast_eval2<KIND>(f2[q], p, t0, t1);
                accumulate_v(__fsub_rn(fabsf(t0), f[2 * q].T), clo[2 * q], mn[2 * q]);
                accumulate_v(__fsub_rn(fabsf(t1), f[2 * q + 1].T), clo[2 * q + 1], mn[2 * q + 1]);
            }
        } else {
#pragma unroll
            for (int h = 0; h < HPT; ++h) accumulate_v(fast_v<KIND>(f[h], p), clo[h], mn[h]);
        }
    };

    uint32_t nres = 0;
    for (uint32_t k = 0;; ++k) {
        const int st = k % kStages;
        mbar_wait(&full[st], (k / kStages) & 1);
        const uint32_t t = tile_id[st];
        if (t == kNoTile) break;
        const float4 *sp = tiles + (size_t)st * kTile;
        const uint32_t base = t * kTile;
        const int npt = (int)min((uint32_t)kTile, a.n - base);
        for (int s0 = 0; s0 < npt; s0 += kSub) {
            const int cnt = min(kSub, npt - s0);
            if (cnt == kSub) {
#pragma unroll kUnroll
                for (int j = 0; j < kSub; ++j) eval_point(sp[s0 + j]);
            } else {
                for (int j = 0; j < cnt; ++j) eval_point(sp[s0 + j]);
            }
            bool any_flag = false;
#pragma unroll
            for (int h = 0; h < HPT; ++h) any_flag = any_flag || (mn[h] < f[h].band);
            if (__any_sync(0xffffffffu, any_flag)) { /* rare: some lane saw a point inside its band */
#pragma unroll
                for (int h = 0; h < HPT; ++h) {
                    const bool flag = mn[h] < f[h].band;
                    const unsigned need = __ballot_sync(0xffffffffu, flag && row[h] < a.rows);
                    if (need) {
                        if (kPacked) { /* scalar view of this hypothesis' half of the packed registers */
                            Fast<KIND> g;
                            unpack_fast<KIND>(f2[h / 2], h & 1, f[h].T, f[h].band, g);
                            rescan_warp<KIND>(a, need, row[h], g, sp + s0, base + s0, cnt, nres);
                        } else {
                            rescan_warp<KIND>(a, need, row[h], f[h], sp + s0, base + s0, cnt, nres);
                        }
                    }
                    if (flag) mn[h] = INFINITY;
                }
            }
        }
        __syncwarp();
        if ((tid & 31) == 0) mbar_arrive(&empty[st]); /* this warp is done with the slot */
    }

#pragma unroll
    for (int h = 0; h < HPT; ++h) {
        if (row[h] < a.rows) {
            const uint32_t ci = a.cnt_row(row[h]);
            if (clo[h]) atomicAdd(&a.counts[ci], clo[h]);
            if (invalid[h] && blockIdx.y == 0) atomicOr(&a.counts[ci], kInvalidBit);
        }
    }
    if (nres) atomicAdd(a.resolves, (unsigned long long)nres);
}

/* fp64 reference-order scoring of every point (no fp32 copy involved) */
template <int KIND>
__global__ void __launch_bounds__(128) score_exact_kernel(const ScoreArgs a) {
    const uint32_t row = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t ntiles = (a.n + kTile - 1) / kTile;
    const uint32_t p0 = min((uint64_t)a.n, (uint64_t)blockIdx.y * a.chunk_tiles * kTile);
    const uint32_t p1 = min((uint64_t)a.n, ((uint64_t)blockIdx.y + 1) * a.chunk_tiles * kTile);
    (void)ntiles;
    if (row >= a.rows) return;
    double m[8];
    const bool ok = fit_row<KIND>(a.xyz, a.nrm, a.samples, a.src_row(row), m, a.row_nrm);
    if (blockIdx.y == 0) {
#pragma unroll
        for (int i = 0; i < 8; ++i) a.models[(size_t)row * 8 + i] = (ok && i < param_count(KIND)) ? m[i] : 0.0;
    }
    if (!ok) {
        if (blockIdx.y == 0) atomicOr(&a.counts[a.cnt_row(row)], kInvalidBit);
        return;
    }
    ex::Dist<KIND> dist;
    dist.set(m);
    uint32_t c = 0;
    for (uint32_t i = p0; i < p1; ++i) c += (dist(ex::ld3(a.xyz + 3 * (size_t)i)) < a.thr) ? 1u : 0u;
    if (c) atomicAdd(&a.counts[a.cnt_row(row)], c);
    atomicAdd(a.resolves, (unsigned long long)(p1 - p0));
}

/* ------------------------------------------------------------------------------- RefineModel */
constexpr int kRB = 256;    /* threads per refine block                     */
constexpr int kRItems = 8;  /* points per thread; block = 2048 consecutive points */
constexpr int kRBlockPts = kRB * kRItems;

struct RefineMid { /* written by refine_scan */
    unsigned long long n_inl;
    double err;     /* sum of inlier distances (parallel fp64 sum, fixed order) */
    double mean[3]; /* inlier centroid */
};
struct RefineOut {
    double model[8];
    int ok;
    int pad;
};

/* pass 1 (ransac.h:536-543 predicate): per-block inlier count and partial sums */
template <int KIND>
__global__ void __launch_bounds__(kRB) refine_count_kernel(const double *__restrict__ xyz, uint32_t n,
                                                           const double *__restrict__ model, double thr,
                                                           uint32_t *__restrict__ blk_cnt,
                                                           double *__restrict__ blk_part /*[nblk][4]*/) {
    ex::Dist<KIND> dist;
    {
        double m[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) m[i] = model[i];
        dist.set(m);
    }
    uint32_t cnt = 0;
    double s[4] = {0, 0, 0, 0};
    const uint32_t base = blockIdx.x * kRBlockPts;
#pragma unroll
    for (int it = 0; it < kRItems; ++it) {
        const uint32_t i = base + it * kRB + threadIdx.x;
        if (i < n) {
            const ex::V3 q = ex::ld3(xyz + 3 * (size_t)i);
            const double d = dist(q);
            if (d < thr) {
                ++cnt;
                s[0] += d;
                s[1] += q.x;
                s[2] += q.y;
                s[3] += q.z;
            }
        }
    }
    __shared__ double sh[8][4];
    __shared__ uint32_t shc[8];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    for (int o = 16; o; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
#pragma unroll
    for (int k = 0; k < 4; ++k) s[k] = warp_sum(s[k]);
    if (lane == 0) {
        shc[w] = cnt;
#pragma unroll
        for (int k = 0; k < 4; ++k) sh[w][k] = s[k];
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t c = 0;
        double r[4] = {0, 0, 0, 0};
        for (int k = 0; k < 8; ++k) {
            c += shc[k];
            for (int q = 0; q < 4; ++q) r[q] += sh[k][q];
        }
        blk_cnt[blockIdx.x] = c;
        for (int q = 0; q < 4; ++q) blk_part[(size_t)blockIdx.x * 4 + q] = r[q];
    }
}

/* pass 2: exclusive scan of the block counts + fixed-order reduction of the partial sums */
__global__ void __launch_bounds__(1024) refine_scan_kernel(const uint32_t *__restrict__ blk_cnt,
                                                           const double *__restrict__ blk_part,
                                                           uint32_t nblk, uint32_t *__restrict__ blk_off,
                                                           RefineMid *__restrict__ mid) {
    __shared__ uint32_t wsum[32];
    __shared__ uint32_t carry_s;
    __shared__ double red[32][4];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    double s[4] = {0, 0, 0, 0};
    for (uint32_t b0 = 0; b0 < nblk; b0 += 1024) {
        const uint32_t b = b0 + threadIdx.x;
        const uint32_t v = b < nblk ? blk_cnt[b] : 0;
        if (b < nblk) {
#pragma unroll
            for (int k = 0; k < 4; ++k) s[k] += blk_part[(size_t)b * 4 + k];
        }
        uint32_t x = v;
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
            if (lane >= o) x += y;
        }
        if (lane == 31) wsum[w] = x;
        __syncthreads();
        if (w == 0) {
            uint32_t ws = wsum[lane];
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t y = __shfl_up_sync(0xffffffffu, ws, o);
                if (lane >= o) ws += y;
            }
            wsum[lane] = ws; /* inclusive */
        }
        __syncthreads();
        const uint32_t carry = carry_s;
        const uint32_t woff = w ? wsum[w - 1] : 0;
        if (b < nblk) blk_off[b] = carry + woff + x - v;
        __syncthreads();
        if (threadIdx.x == 0) carry_s = carry + wsum[31];
        __syncthreads();
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) s[k] = warp_sum(s[k]);
    if (lane == 0) {
#pragma unroll
        for (int k = 0; k < 4; ++k) red[w][k] = s[k];
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        double r[4] = {0, 0, 0, 0};
        for (int k = 0; k < 32; ++k)
            for (int q = 0; q < 4; ++q) r[q] += red[k][q];
        const unsigned long long ni = carry_s;
        mid->n_inl = ni;
        mid->err = r[0];
        const double inv = ni ? 1.0 / (double)ni : 0.0;
        mid->mean[0] = r[1] * inv;
        mid->mean[1] = r[2] * inv;
        mid->mean[2] = r[3] * inv;
    }
}

/* segmentation side outputs of pass 3 (all optional) */
struct SegArgs {
    const float4 *pts32_in;
    const uint32_t *orig_in; /* original index of every current point */
    double *xyz_out;
    float4 *pts32_out;
    uint32_t *orig_out;
    uint32_t *labels; /* [n_original], 0xFFFFFFFF = unassigned */
    uint32_t plane_id;
};

/* pass 3: ascending inlier indices (stable), centred moments for GeneralFit; with SEG also the
 * stable compaction of the remaining points (SelectByIndex(inliers, invert=true),
 * iterative_plane_segmentation.cpp:32-33) and the cluster labels */
template <int KIND, bool SEG>
__global__ void __launch_bounds__(kRB) refine_write_kernel(const double *__restrict__ xyz, uint32_t n,
                                                           const double *__restrict__ model, double thr,
                                                           const uint32_t *__restrict__ blk_off,
                                                           const RefineMid *__restrict__ mid,
                                                           unsigned long long *__restrict__ inl,
                                                           double *__restrict__ blk_mom /*[nblk][10]*/,
                                                           const SegArgs seg) {
    ex::Dist<KIND> dist;
    {
        double m[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) m[i] = model[i];
        dist.set(m);
    }
    const double mx = mid->mean[0], my = mid->mean[1], mz = mid->mean[2];
    __shared__ uint32_t wcnt[8];
    __shared__ double sh[8][10];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const uint32_t base = blockIdx.x * kRBlockPts;
    uint32_t running = blk_off[blockIdx.x];
    double mom[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll 1
    for (int it = 0; it < kRItems; ++it) {
        const uint32_t i = base + it * kRB + threadIdx.x;
        bool in = false;
        ex::V3 q = {0, 0, 0};
        if (i < n) {
            q = ex::ld3(xyz + 3 * (size_t)i);
            in = dist(q) < thr;
        }
        const uint32_t bal = __ballot_sync(0xffffffffu, in);
        if (lane == 0) wcnt[w] = __popc(bal);
        __syncthreads();
        uint32_t before = 0, total = 0;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const uint32_t c = wcnt[k];
            before += (k < w) ? c : 0;
            total += c;
        }
        const uint32_t rank = running + before + __popc(bal & ((1u << lane) - 1));
        if (in) {
            inl[rank] = i;
            const double ux = q.x - mx, uy = q.y - my, uz = q.z - mz;
            mom[0] += ux * ux;
            mom[1] += ux * uy;
            mom[2] += ux * uz;
            mom[3] += uy * uy;
            mom[4] += uy * uz;
            mom[5] += uz * uz;
            if (KIND == kSphere) {
                const double uu = ux * ux + uy * uy + uz * uz;
                mom[6] += ux * uu;
                mom[7] += uy * uu;
                mom[8] += uz * uu;
                mom[9] += uu;
            }
            if (SEG) seg.labels[seg.orig_in[i]] = seg.plane_id;
        } else if (SEG && i < n) {
            const uint32_t dst = i - rank; /* rank == number of inliers before i */
            seg.xyz_out[3 * (size_t)dst] = q.x;
            seg.xyz_out[3 * (size_t)dst + 1] = q.y;
            seg.xyz_out[3 * (size_t)dst + 2] = q.z;
            seg.pts32_out[dst] = seg.pts32_in[i];
            seg.orig_out[dst] = seg.orig_in[i];
        }
        running += total;
        __syncthreads();
    }
    constexpr int NM = (KIND == kSphere) ? 10 : 6;
#pragma unroll
    for (int k = 0; k < NM; ++k) mom[k] = warp_sum(mom[k]);
    if (lane == 0) {
#pragma unroll
        for (int k = 0; k < NM; ++k) sh[w][k] = mom[k];
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int q = 0; q < NM; ++q) {
            double r = 0;
            for (int k = 0; k < 8; ++k) r += sh[k][q];
            blk_mom[(size_t)blockIdx.x * 10 + q] = r;
        }
    }
}

/* pass 4: GeneralFit from the moments (PlaneEstimator::GeneralFit ransac.h:164-213;
 * SphereEstimator::GeneralFit ransac.h:296-330 as centred normal equations instead of the
 * reference's full-U BDCSVD; CylinderEstimator::GeneralFit ransac.h:427-433 is a no-op) */
template <int KIND>
__global__ void __launch_bounds__(256) refine_final_kernel(const double *__restrict__ blk_mom, uint32_t nblk,
                                                           const RefineMid *__restrict__ mid,
                                                           const double *__restrict__ model,
                                                           RefineOut *__restrict__ out) {
    __shared__ double sh[8][10];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    double mom[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    for (uint32_t b = threadIdx.x; b < nblk; b += blockDim.x)
#pragma unroll
        for (int k = 0; k < 10; ++k) mom[k] += blk_mom[(size_t)b * 10 + k];
#pragma unroll
    for (int k = 0; k < 10; ++k) mom[k] = warp_sum(mom[k]);
    if (lane == 0)
#pragma unroll
        for (int k = 0; k < 10; ++k) sh[w][k] = mom[k];
    __syncthreads();
    if (threadIdx.x != 0) return;
    for (int q = 0; q < 10; ++q) {
        double r = 0;
        for (int k = 0; k < 8; ++k) r += sh[k][q];
        mom[q] = r;
    }
    double m[8];
    for (int i = 0; i < 8; ++i) m[i] = model[i];
    int ok = 1;
    const unsigned long long ni = mid->n_inl;
    const double mx = mid->mean[0], my = mid->mean[1], mz = mid->mean[2];
    if (KIND == kPlane) {
        if (ni < 3) {
            ok = 0;
        } else {
            const double xx = mom[0], xy = mom[1], xz = mom[2], yy = mom[3], yz = mom[4], zz = mom[5];
            const double det_x = yy * zz - yz * yz, det_y = xx * zz - xz * xz, det_z = xx * yy - xy * xy;
            double a, b, c;
            if (det_x > det_y && det_x > det_z) { /* ransac.h:195-201 */
                a = det_x;
                b = xz * yz - xy * zz;
                c = xy * yz - xz * yy;
            } else if (det_y > det_z) {
                a = xz * yz - xy * zz;
                b = det_y;
                c = xy * xz - yz * xx;
            } else {
                a = xy * yz - xz * yy;
                b = xy * xz - yz * xx;
                c = det_z;
            }
            const double nrm = sqrt(a * a + b * b + c * c);
            if (nrm < kEps) {
                ok = 0;
            } else {
                a /= nrm;
                b /= nrm;
                c /= nrm;
                m[0] = a;
                m[1] = b;
                m[2] = c;
                m[3] = -(a * mx + b * my + c * mz);
            }
        }
    } else if (KIND == kSphere) {
        if (ni < 4) {
            ok = 0;
        } else {
            /* min sum (2 c'.u + w - |u|^2)^2 with sum u = 0:  2 (sum u u^T) c' = sum u |u|^2 ;
             * r^2 = |c'|^2 + mean |u|^2 */
            const double A00 = mom[0], A01 = mom[1], A02 = mom[2], A11 = mom[3], A12 = mom[4], A22 = mom[5];
            const double b0 = 0.5 * mom[6], b1 = 0.5 * mom[7], b2 = 0.5 * mom[8];
            const double c00 = A11 * A22 - A12 * A12, c01 = A02 * A12 - A01 * A22, c02 = A01 * A12 - A02 * A11;
            const double c11 = A00 * A22 - A02 * A02, c12 = A01 * A02 - A00 * A12, c22 = A00 * A11 - A01 * A01;
            const double det = A00 * c00 + A01 * c01 + A02 * c02;
            double x, y, z;
            const double tr = A00 + A11 + A22;
            if (fabs(det) > 1e-12 * tr * tr * tr && isfinite(det)) {
                x = (c00 * b0 + c01 * b1 + c02 * b2) / det;
                y = (c01 * b0 + c11 * b1 + c12 * b2) / det;
                z = (c02 * b0 + c12 * b1 + c22 * b2) / det;
            } else {
                /* coplanar / collinear inliers: the system is rank deficient.  The reference's bdcSvd().solve()
                 * returns a finite minimum-norm solution; here the minimum-norm solution of the CENTRED system
                 * (pseudo-inverse through a Jacobi eigen-decomposition, eigenvalues below 1e-12 of the largest
                 * dropped) -- finite and on the same sphere family, not bit-comparable (documented deviation). */
                double a[3][3] = {{A00, A01, A02}, {A01, A11, A12}, {A02, A12, A22}};
                double v[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
                for (int sweep = 0; sweep < 12; ++sweep)
                    for (int p = 0; p < 2; ++p)
                        for (int q = p + 1; q < 3; ++q) {
                            if (fabs(a[p][q]) < 1e-300) continue;
                            const double th = (a[q][q] - a[p][p]) / (2.0 * a[p][q]);
                            const double t = (th >= 0 ? 1.0 : -1.0) / (fabs(th) + sqrt(th * th + 1.0));
                            const double c = 1.0 / sqrt(t * t + 1.0), sn = t * c;
                            for (int k = 0; k < 3; ++k) { /* A <- A J */
                                const double akp = a[k][p], akq = a[k][q];
                                a[k][p] = c * akp - sn * akq;
                                a[k][q] = sn * akp + c * akq;
                            }
                            for (int k = 0; k < 3; ++k) { /* A <- J^T A */
                                const double apk = a[p][k], aqk = a[q][k];
                                a[p][k] = c * apk - sn * aqk;
                                a[q][k] = sn * apk + c * aqk;
                            }
                            for (int k = 0; k < 3; ++k) {
                                const double vkp = v[k][p], vkq = v[k][q];
                                v[k][p] = c * vkp - sn * vkq;
                                v[k][q] = sn * vkp + c * vkq;
                            }
                        }
                const double lmax = fmax(fabs(a[0][0]), fmax(fabs(a[1][1]), fabs(a[2][2])));
                x = y = z = 0;
                for (int i = 0; i < 3; ++i) {
                    const double lam = a[i][i];
                    if (!(lam > 1e-12 * lmax) || !(lmax > 0)) continue;
                    const double w = (v[0][i] * b0 + v[1][i] * b1 + v[2][i] * b2) / lam;
                    x += w * v[0][i];
                    y += w * v[1][i];
                    z += w * v[2][i];
                }
            }
            m[0] = x + mx;
            m[1] = y + my;
            m[2] = z + mz;
            m[3] = sqrt(x * x + y * y + z * z + mom[9] / (double)ni);
        }
    }
    for (int i = 0; i < 8; ++i) out->model[i] = m[i];
    out->ok = ok;
}

/* EvaluateModel's error exactly as the reference sums it (ransac.h:632-640): index order, one
 * accumulator.  Only used to break inlier-count ties whose parallel sums are too close to call. */
template <int KIND>
__global__ void seq_err_kernel(const double *__restrict__ xyz, const unsigned long long *__restrict__ inl,
                               unsigned long long n_inl, const double *__restrict__ model,
                               double *__restrict__ out) {
    ex::Dist<KIND> dist;
    {
        double m[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) m[i] = model[i];
        dist.set(m);
    }
    const int lane = threadIdx.x;
    double acc = 0;
    for (unsigned long long b = 0; b < n_inl; b += 32) {
        const unsigned long long j = b + lane;
        double d = 0;
        if (j < n_inl) d = dist(ex::ld3(xyz + 3 * (size_t)inl[j]));
        const int cnt = (int)min((unsigned long long)32, n_inl - b);
        for (int k = 0; k < cnt; ++k) acc = ex::add(acc, __shfl_sync(0xffffffffu, d, k));
    }
    if (lane == 0) *out = acc;
}

}  // namespace m3d
