/*
 * matching.cu -- descriptor correspondence matching on sm_100a.
 *
 * Replaces misc3d::registration::ANNMatcher::Match / NearestSearch
 * (src/correspondence_matching.cpp:13-84): bidirectional 1-NN in descriptor space followed by the
 * mutual cross-check (:67-78).  The reference searches with FLANN (exact kd-tree) or Annoy
 * (approximate); here both enum values run one exact brute-force search:
 *
 *   nn_top2_kernel   fp32 |a|^2 + |b|^2 - 2 a.b over 128 x 128 tiles (descriptors staged k-major
 *                    in shared memory by 1-D TMA bulk copies, 8 x 8 register micro-tiles), keeping
 *                    the best and second-best distance of every query row;
 *   nn_exact_kernel  rows whose best/second-best gap is inside the fp32 error bound are searched
 *                    again in fp64 with the reference metric's accumulation order (nanoflann
 *                    L2_Adaptor: groups of four) and the lowest-index tie rule;
 *   mutual_*         nn10[nn01[i]] == i, stable compaction in ascending source index.
 */
#include <algorithm>
#include <vector>

#include "context.h"
#include "exact_math.cuh"
#include "matching_tc.cuh"

namespace m3d {

constexpr int kMT = 128;      /* descriptors per tile */
constexpr int kMaxKP = 128;   /* largest (padded) dimension the fp32 tile kernel handles */
constexpr float kPadNorm = 1e38f;

__device__ __forceinline__ uint32_t m_smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void m_mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(m_smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void m_mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "LAB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra DONE;\n"
        "bra LAB_WAIT;\n"
        "DONE:\n"
        "}\n" ::"r"(m_smem_u32(bar)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void m_tma_load_1d(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
    const uint32_t b = m_smem_u32(bar);
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(bytes) : "memory");
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
            m_smem_u32(dst)),
        "l"(src), "r"(bytes), "r"(b)
        : "memory");
}

/* order-preserving map double -> uint64 (for atomicMin/atomicMax) */
__device__ __forceinline__ unsigned long long enc_f64(double x) {
    const unsigned long long b = (unsigned long long)__double_as_longlong(x);
    return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}
__device__ __forceinline__ double dec_f64(unsigned long long e) {
    const unsigned long long b = (e >> 63) ? (e & 0x7fffffffffffffffull) : ~e;
    return __longlong_as_double((long long)b);
}

/* per-dimension min / max over a descriptor set (F: dim x count column-major) */
__global__ void __launch_bounds__(256) feat_minmax_kernel(const double *__restrict__ F, size_t total, int dim,
                                                          unsigned long long *__restrict__ mn,
                                                          unsigned long long *__restrict__ mx) {
    for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (size_t)gridDim.x * blockDim.x) {
        const double v = F[e];
        if (v != v) continue;
        const int k = (int)(e % (size_t)dim);
        const unsigned long long c = enc_f64(v);
        if (c < mn[k]) atomicMin(&mn[k], c);
        if (c > mx[k]) atomicMax(&mx[k], c);
    }
}
__global__ void feat_center_kernel(const unsigned long long *mn, const unsigned long long *mx, int dim,
                                   double *center) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= dim) return;
    const double c = 0.5 * (dec_f64(mn[k]) + dec_f64(mx[k]));
    center[k] = isfinite(c) ? c : 0.0;
}

/* fp64 column-major descriptors -> centred fp32 tiles [(KP+1)][128] (k-major; row KP = |.|^2) */
__global__ void __launch_bounds__(128) feat_convert_kernel(const double *__restrict__ F, uint32_t count, int dim,
                                                           int KP, const double *__restrict__ center,
                                                           float *__restrict__ tiles,
                                                           uint32_t *__restrict__ maxnorm_bits) {
    const uint32_t tile = blockIdx.x, lane = threadIdx.x;
    const uint32_t j = tile * kMT + lane;
    float *t = tiles + (size_t)tile * (KP + 1) * kMT;
    double n2 = 0;
    if (j < count) {
        for (int k = 0; k < dim; ++k) {
            const double v = F[(size_t)j * dim + k] - center[k];
            n2 += v * v;
            t[(size_t)k * kMT + lane] = (float)v;
        }
        for (int k = dim; k < KP; ++k) t[(size_t)k * kMT + lane] = 0.f;
        const float nf = (float)n2;
        t[(size_t)KP * kMT + lane] = nf;
        if (nf == nf && nf < 3e38f) atomicMax(maxnorm_bits, __float_as_uint(nf));
    } else {
        for (int k = 0; k < KP; ++k) t[(size_t)k * kMT + lane] = 0.f;
        t[(size_t)KP * kMT + lane] = kPadNorm;
    }
}

/* best / second-best neighbour of every row of A among the columns B */
__global__ void __launch_bounds__(256, 2) nn_top2_kernel(const float *__restrict__ At, const float *__restrict__ Bt,
                                                         uint32_t na, uint32_t nb, int KP,
                                                         const uint32_t *__restrict__ maxnorm_bits,
                                                         uint32_t *__restrict__ nn, uint32_t *__restrict__ amb_list,
                                                         uint32_t *__restrict__ amb_count) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int tile_f = (KP + 1) * kMT;
    float *As = reinterpret_cast<float *>(smem_raw);
    float *Bs = As + tile_f;
    uint64_t *bars = reinterpret_cast<uint64_t *>(Bs + 2 * tile_f);
    const uint32_t tile_bytes = (uint32_t)tile_f * 4u;
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const uint32_t ntb = (nb + kMT - 1) / kMT;

    if (tid == 0) {
        m_mbar_init(&bars[0], 1);
        m_mbar_init(&bars[1], 1);
        m_mbar_init(&bars[2], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        m_tma_load_1d(As, At + (size_t)blockIdx.x * tile_f, tile_bytes, &bars[2]);
        m_tma_load_1d(Bs, Bt, tile_bytes, &bars[0]);
        if (ntb > 1) m_tma_load_1d(Bs + tile_f, Bt + (size_t)tile_f, tile_bytes, &bars[1]);
    }
    __syncthreads();
    m_mbar_wait(&bars[2], 0);

    float an[8], m1[8], m2[8];
    uint32_t i1[8];
#pragma unroll
    for (int r = 0; r < 8; ++r) {
        an[r] = As[KP * kMT + ty * 8 + r];
        m1[r] = INFINITY;
        m2[r] = INFINITY;
        i1[r] = 0;
    }
    const int c_lo = tx * 4, c_hi = 64 + tx * 4; /* this thread's columns: c_lo..+3 and c_hi..+3 */

    for (uint32_t t = 0; t < ntb; ++t) {
        const float *B = Bs + (size_t)(t & 1) * tile_f;
        m_mbar_wait(&bars[t & 1], (t >> 1) & 1);
        float acc[8][8];
#pragma unroll
        for (int r = 0; r < 8; ++r)
#pragma unroll
            for (int c = 0; c < 8; ++c) acc[r][c] = 0.f;
#pragma unroll 4
        for (int k = 0; k < KP; ++k) {
            const float4 a0 = *reinterpret_cast<const float4 *>(As + k * kMT + ty * 8);
            const float4 a1 = *reinterpret_cast<const float4 *>(As + k * kMT + ty * 8 + 4);
            const float4 b0 = *reinterpret_cast<const float4 *>(B + k * kMT + c_lo);
            const float4 b1 = *reinterpret_cast<const float4 *>(B + k * kMT + c_hi);
            const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
            for (int r = 0; r < 8; ++r)
#pragma unroll
                for (int c = 0; c < 8; ++c) acc[r][c] = fmaf(a[r], b[c], acc[r][c]);
        }
        const float4 n0 = *reinterpret_cast<const float4 *>(B + KP * kMT + c_lo);
        const float4 n1 = *reinterpret_cast<const float4 *>(B + KP * kMT + c_hi);
        const float bn[8] = {n0.x, n0.y, n0.z, n0.w, n1.x, n1.y, n1.z, n1.w};
        const uint32_t jbase = t * kMT;
        const bool last = (t == ntb - 1);
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            const uint32_t j = jbase + (c < 4 ? c_lo + c : c_hi + c - 4);
            if (last && j >= nb) continue;
#pragma unroll
            for (int r = 0; r < 8; ++r) {
                const float d = fmaf(-2.f, acc[r][c], an[r] + bn[c]);
                if (d < m2[r]) {
                    if (d < m1[r]) {
                        m2[r] = m1[r];
                        m1[r] = d;
                        i1[r] = j;
                    } else {
                        m2[r] = d;
                    }
                }
            }
        }
        __syncthreads(); /* everyone is done with this stage */
        if (tid == 0 && t + 2 < ntb)
            m_tma_load_1d(Bs + (size_t)(t & 1) * tile_f, Bt + (size_t)(t + 2) * tile_f, tile_bytes, &bars[t & 1]);
    }

    /* merge the 16 threads (tx) that share a row; columns of a thread ascend with tx within each
     * half, so ties are resolved on the index explicitly */
    const float bnmax = __uint_as_float(*maxnorm_bits);
#pragma unroll
    for (int r = 0; r < 8; ++r) {
        float a1 = m1[r], a2 = m2[r];
        uint32_t ai = i1[r];
#pragma unroll
        for (int o = 1; o < 16; o <<= 1) {
            const float b1 = __shfl_xor_sync(0xffffffffu, a1, o);
            const float b2 = __shfl_xor_sync(0xffffffffu, a2, o);
            const uint32_t bi = __shfl_xor_sync(0xffffffffu, ai, o);
            if (b1 < a1 || (b1 == a1 && bi < ai)) {
                a2 = fminf(a1, b2);
                a1 = b1;
                ai = bi;
            } else {
                a2 = fminf(a2, b1);
            }
        }
        const uint32_t row = blockIdx.x * kMT + ty * 8 + r;
        if (tx == 0 && row < na) {
            nn[row] = ai;
            /* |d32 - exact| <= (KP + 10) * 2^-24 * (|a|^2 + max|b|^2) */
            const float E = (float)(KP + 10) * 5.9604645e-08f * (an[r] + bnmax);
            if (!(a2 - a1 > 2.5f * E)) amb_list[atomicAdd(amb_count, 1u)] = row;
        }
    }
}

/* nanoflann L2_Adaptor::evalMetric order: groups of four, then the tail */
__device__ __forceinline__ double l2_groups4(const double *a, const double *b, int dim) {
    double result = 0;
    int d = 0;
    for (; d + 3 < dim; d += 4) {
        const double d0 = ex::sub(a[d], b[d]), d1 = ex::sub(a[d + 1], b[d + 1]), d2 = ex::sub(a[d + 2], b[d + 2]),
                     d3 = ex::sub(a[d + 3], b[d + 3]);
        result = ex::add(result, ex::add(ex::add(ex::add(ex::mul(d0, d0), ex::mul(d1, d1)), ex::mul(d2, d2)),
                                         ex::mul(d3, d3)));
    }
    for (; d < dim; ++d) {
        const double d0 = ex::sub(a[d], b[d]);
        result = ex::add(result, ex::mul(d0, d0));
    }
    return result;
}

/* exact 1-NN (strict <, lowest index on ties) for the listed rows.  blockIdx.x walks groups of kXR
 * rows (every database descriptor fetched from L2 serves kXR queries), blockIdx.y owns one of kXS
 * contiguous column ranges; partial winners go to (pd, pj)[slot][y] and nn_exact_merge_kernel
 * combines them in ascending column order.  list == nullptr: all rows */
constexpr int kXR = 16, kXS = 8;
__global__ void __launch_bounds__(256) nn_exact_kernel(const double *__restrict__ A, const double *__restrict__ B,
                                                       uint32_t na, uint32_t nb, int dim,
                                                       const uint32_t *__restrict__ list,
                                                       const uint32_t *__restrict__ list_count,
                                                       double *__restrict__ pd, uint32_t *__restrict__ pj) {
    extern __shared__ double qa[]; /* kXR x dim doubles (rows past the end: copies of row 0 of the group) */
    __shared__ double sd[kXR][8];
    __shared__ uint32_t sj[kXR][8];
    const uint32_t total = list ? *list_count : na;
    const uint32_t groups = (total + kXR - 1) / kXR;
    const uint32_t per = (nb + kXS - 1) / kXS;
    const uint32_t j0 = min(nb, blockIdx.y * per), j1 = min(nb, j0 + per);
    for (uint32_t g = blockIdx.x; g < groups; g += gridDim.x) {
        const uint32_t nr = min((uint32_t)kXR, total - g * kXR);
        __syncthreads();
        for (uint32_t e = threadIdx.x; e < (uint32_t)kXR * (uint32_t)dim; e += blockDim.x) {
            const uint32_t r = e / dim, k = e % dim;
            const uint32_t rr = r < nr ? r : 0u;
            const uint32_t row = list ? list[g * kXR + rr] : g * kXR + rr;
            qa[r * dim + k] = A[(size_t)row * dim + k];
        }
        __syncthreads();
        double best[kXR];
        uint32_t bj[kXR];
#pragma unroll
        for (int r = 0; r < kXR; ++r) {
            best[r] = INFINITY;
            bj[r] = 0xffffffffu;
        }
        /* thread = database column; the kXR query rows of the group share every loaded descriptor word (the kernel is
         * bound by the fp64 pipe, not by loads).  Per row the sum is formed exactly as l2_groups4 does. */
        for (uint32_t j = j0 + threadIdx.x; j < j1; j += blockDim.x) {
            const double *b = B + (size_t)j * dim;
            double acc[kXR];
#pragma unroll
            for (int r = 0; r < kXR; ++r) acc[r] = 0;
            int d = 0;
            for (; d + 3 < dim; d += 4) {
                const double b0 = b[d], b1 = b[d + 1], b2 = b[d + 2], b3 = b[d + 3];
#pragma unroll
                for (int r = 0; r < kXR; ++r) {
                    const double *a = qa + r * dim + d;
                    const double d0 = ex::sub(a[0], b0), d1 = ex::sub(a[1], b1), d2 = ex::sub(a[2], b2), d3 = ex::sub(a[3], b3);
                    acc[r] = ex::add(acc[r], ex::add(ex::add(ex::add(ex::mul(d0, d0), ex::mul(d1, d1)), ex::mul(d2, d2)),
                                                     ex::mul(d3, d3)));
                }
            }
            for (; d < dim; ++d) {
                const double b0 = b[d];
#pragma unroll
                for (int r = 0; r < kXR; ++r) {
                    const double d0 = ex::sub(qa[r * dim + d], b0);
                    acc[r] = ex::add(acc[r], ex::mul(d0, d0));
                }
            }
#pragma unroll
            for (int r = 0; r < kXR; ++r)
                if (acc[r] < best[r]) {
                    best[r] = acc[r];
                    bj[r] = j;
                }
        }
#pragma unroll
        for (int r = 0; r < kXR; ++r) {
            double bb = best[r];
            uint32_t jj = bj[r];
            for (int o = 16; o; o >>= 1) {
                const double ob = __shfl_xor_sync(0xffffffffu, bb, o);
                const uint32_t oj = __shfl_xor_sync(0xffffffffu, jj, o);
                if (ob < bb || (ob == bb && oj < jj)) {
                    bb = ob;
                    jj = oj;
                }
            }
            if ((threadIdx.x & 31) == 0) {
                sd[r][threadIdx.x >> 5] = bb;
                sj[r][threadIdx.x >> 5] = jj;
            }
        }
        __syncthreads();
        if (threadIdx.x < nr) {
            const int r = threadIdx.x;
            double bb = sd[r][0];
            uint32_t jj = sj[r][0];
            for (int w = 1; w < 8; ++w)
                if (sd[r][w] < bb || (sd[r][w] == bb && sj[r][w] < jj)) {
                    bb = sd[r][w];
                    jj = sj[r][w];
                }
            const size_t slot = (size_t)(g * kXR + r) * kXS + blockIdx.y;
            pd[slot] = bb;
            pj[slot] = jj;
        }
    }
}
__global__ void nn_exact_merge_kernel(uint32_t na, const uint32_t *__restrict__ list,
                                      const uint32_t *__restrict__ list_count, const double *__restrict__ pd,
                                      const uint32_t *__restrict__ pj, uint32_t *__restrict__ nn) {
    const uint32_t total = list ? *list_count : na;
    for (uint32_t it = blockIdx.x * blockDim.x + threadIdx.x; it < total; it += gridDim.x * blockDim.x) {
        double bb = INFINITY;
        uint32_t jj = 0xffffffffu;
        for (int y = 0; y < kXS; ++y) { /* ascending column ranges: strict < keeps the lowest index */
            const double d = pd[(size_t)it * kXS + y];
            if (d < bb) {
                bb = d;
                jj = pj[(size_t)it * kXS + y];
            }
        }
        nn[list ? list[it] : it] = (jj == 0xffffffffu) ? 0u : jj; /* nothing below +inf: the reference keeps 0 */
    }
}

/* exact decision among the GEMM pass's candidates of every ambiguous row (thread = slot): fp64
 * reference-order distance, lowest value then lowest index.  Rows with more candidates than the
 * list holds go to the full fp64 re-search (list2). */
__global__ void __launch_bounds__(128) cand_exact_kernel(const double *__restrict__ A, const double *__restrict__ B,
                                                         int dim, const uint32_t *__restrict__ amb_list,
                                                         const uint32_t *__restrict__ amb_count,
                                                         const uint32_t *__restrict__ cand,
                                                         const uint32_t *__restrict__ cand_count,
                                                         uint32_t *__restrict__ nn, uint32_t *__restrict__ list2,
                                                         uint32_t *__restrict__ list2_count) {
    const uint32_t n = *amb_count;
    for (uint32_t s = blockIdx.x * blockDim.x + threadIdx.x; s < n; s += gridDim.x * blockDim.x) {
        const uint32_t row = amb_list[s];
        const uint32_t c = cand_count[s];
        if (c == 0 || c > (uint32_t)tc::kCandCap) { /* c == 0 cannot happen for finite data (the best column qualifies) */
            list2[atomicAdd(list2_count, 1u)] = row;
            continue;
        }
        const double *a = A + (size_t)row * dim;
        double best = INFINITY;
        uint32_t bj = 0xffffffffu;
        for (uint32_t k = 0; k < c; ++k) {
            const uint32_t j = cand[(size_t)s * tc::kCandCap + k];
            const double d = l2_groups4(a, B + (size_t)j * dim, dim);
            if (d < best || (d == best && j < bj)) {
                best = d;
                bj = j;
            }
        }
        if (bj == 0xffffffffu) { /* all candidates NaN/inf: let the full search apply the reference's rule */
            list2[atomicAdd(list2_count, 1u)] = row;
            continue;
        }
        nn[row] = bj;
    }
}

/* ---- mutual check (correspondence_matching.cpp:67-78) + stable compaction */
constexpr int kMB = 256, kMItems = 8;
__device__ __forceinline__ bool mutual(const uint32_t *nn01, const uint32_t *nn10, uint32_t i, uint32_t nd) {
    const uint32_t j = nn01[i];
    return j < nd && nn10[j] == i;
}
__global__ void __launch_bounds__(kMB) mutual_count_kernel(const uint32_t *__restrict__ nn01,
                                                           const uint32_t *__restrict__ nn10, uint32_t ns,
                                                           uint32_t nd, uint32_t *__restrict__ blk_cnt) {
    uint32_t c = 0;
    const uint32_t base = blockIdx.x * kMB * kMItems;
    for (int it = 0; it < kMItems; ++it) {
        const uint32_t i = base + it * kMB + threadIdx.x;
        if (i < ns && mutual(nn01, nn10, i, nd)) ++c;
    }
    __shared__ uint32_t sh[8];
    for (int o = 16; o; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t s = 0;
        for (int k = 0; k < 8; ++k) s += sh[k];
        blk_cnt[blockIdx.x] = s;
    }
}
__global__ void mutual_scan_kernel(const uint32_t *__restrict__ blk_cnt, uint32_t nblk,
                                   uint32_t *__restrict__ blk_off, uint32_t *__restrict__ total) {
    if (threadIdx.x != 0) return; /* nblk <= ~1000: a serial scan is a few microseconds */
    uint32_t run = 0;
    for (uint32_t b = 0; b < nblk; ++b) {
        blk_off[b] = run;
        run += blk_cnt[b];
    }
    *total = run;
}
__global__ void __launch_bounds__(kMB) mutual_write_kernel(const uint32_t *__restrict__ nn01,
                                                           const uint32_t *__restrict__ nn10, uint32_t ns,
                                                           uint32_t nd, const uint32_t *__restrict__ blk_off,
                                                           unsigned long long *__restrict__ idx0,
                                                           unsigned long long *__restrict__ idx1) {
    __shared__ uint32_t wcnt[8];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    uint32_t running = blk_off[blockIdx.x];
    const uint32_t base = blockIdx.x * kMB * kMItems;
    for (int it = 0; it < kMItems; ++it) {
        const uint32_t i = base + it * kMB + threadIdx.x;
        const bool f = i < ns && mutual(nn01, nn10, i, nd);
        const uint32_t bal = __ballot_sync(0xffffffffu, f);
        if (lane == 0) wcnt[w] = __popc(bal);
        __syncthreads();
        uint32_t before = 0, tot = 0;
        for (int k = 0; k < 8; ++k) {
            before += (k < w) ? wcnt[k] : 0;
            tot += wcnt[k];
        }
        if (f) {
            const uint32_t pos = running + before + __popc(bal & ((1u << lane) - 1));
            idx0[pos] = i;
            idx1[pos] = nn01[i];
        }
        running += tot;
        __syncthreads();
    }
}

/* ------------------------------------------------------------------------------ host side */
struct FeatDev {
    double *f64 = nullptr; /* dim x count, column-major */
    float *tiles = nullptr;
    uint32_t count = 0, ntiles = 0;
    /* tensor-core path: bf16x3 split tiles in query / database form, fp32 squared norms */
    __nv_bfloat16 *tq = nullptr, *td = nullptr;
    float *norms = nullptr;
};

/* which search kernel: 2 = tcgen05 GEMM (dim <= 47), 1 = fp32 CUDA-core tiles (dim <= 128), 0 = fp64 only */
static int match_path(int dim) {
    const char *e = getenv("M3D_MATCH_PATH"); /* "fp32" / "fp64" force the other kernels (tests, A/B timing) */
    const int KP = (dim + 3) & ~3;
    if (e && !strcmp(e, "fp64")) return 0;
    if (e && !strcmp(e, "fp32")) return KP <= kMaxKP ? 1 : 0;
    if (tc::kprime(dim) <= tc::kMaxKPrime) return 2;
    return KP <= kMaxKP ? 1 : 0;
}

struct MatchScratch { /* layout of the small device block */
    unsigned long long mn[kMaxKP], mx[kMaxKP];
    double center[kMaxKP];
    uint32_t maxnorm[2]; /* per set */
    uint32_t amb_count[2];
    uint32_t total;
};

static int launch_exact(m3d_ctx *ctx, const FeatDev &A, const FeatDev &B, int dim, const uint32_t *d_list,
                        const uint32_t *d_count, uint32_t *d_nn) {
    const size_t xsmem = sizeof(double) * (size_t)dim * kXR;
    if (xsmem > 200 * 1024) return ctx->fail(M3D_ERR_INVALID_ARG, "descriptor dimension %d too large", dim);
    if (xsmem > 48 * 1024)
        M3D_CUDA(ctx, cudaFuncSetAttribute(nn_exact_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)xsmem));
    const size_t slots = ((size_t)A.count + kXR) * kXS;
    M3D_CUDA(ctx, ctx->d_part.reserve(slots * (sizeof(double) + sizeof(uint32_t)) + 64));
    double *pd = ctx->d_part.as<double>();
    uint32_t *pj = reinterpret_cast<uint32_t *>(pd + slots);
    const int gx = std::max(1, ctx->sm_count * 4 / kXS);
    nn_exact_kernel<<<dim3(gx, kXS), 256, xsmem, ctx->stream>>>(A.f64, B.f64, A.count, B.count, dim, d_list, d_count,
                                                               pd, pj);
    M3D_LAUNCHED(ctx);
    nn_exact_merge_kernel<<<ctx->sm_count, 256, 0, ctx->stream>>>(A.count, d_list, d_count, pd, pj, d_nn);
    M3D_LAUNCHED(ctx);
    return M3D_OK;
}

/* M3D_MATCH_TC=2 selects the two-CTA (cta_group::2) tcgen05 kernel.  It halves the L2 -> SM operand stream (ncu: 65 GB
 * instead of 130 GB per direction) and returns identical results, but measured SLOWER than the one-CTA kernel (15.6 ms
 * vs 13.9 ms per direction at C4): the per-tile hand-shake across the pair paces it (12.8 ms with the MMAs and the
 * operand traffic switched off).  Kept as an experiment; default: the one-CTA kernel. */
static bool tc_two_cta() {
    static const bool v = getenv("M3D_MATCH_TC") && atoi(getenv("M3D_MATCH_TC")) == 2;
    return v;
}

static int nn_direction(m3d_ctx *ctx, const FeatDev &A, const FeatDev &B, int dim, int KP, int path,
                        const uint32_t *d_maxnorm_B, uint32_t *d_nn, uint32_t *d_amb, uint32_t *d_amb_count) {
    if (path == 2) {
        tc::TcArgs ta{};
        ta.Aq = A.tq;
        ta.Bd = B.td;
        ta.a_norms = A.norms;
        ta.na = A.count;
        ta.nb = B.count;
        ta.KPr = tc::kprime(dim);
        ta.maxnorm_bits = d_maxnorm_B;
        ta.nn = d_nn;
        ta.amb_list = d_amb;
        ta.amb_count = d_amb_count;
        const size_t smem = (size_t)(tc::kRB + tc::kBStages) * tc::kRows * ta.KPr * 2 + 128;
        /* scratch of the candidate refinement */
        const bool two = tc_two_cta();
        const uint32_t atiles = two ? (A.ntiles + 1) / 2 * 2 : (A.ntiles + tc::kRB - 1) / tc::kRB * tc::kRB; /* whole pairs */
        const size_t smem2 = (size_t)tc::kRows * ta.KPr * 2 + (size_t)tc::kB2Stages * (tc::kRows / 2) * ta.KPr * 2 + 512;
        M3D_CUDA(ctx, ctx->d_models.reserve((size_t)atiles * tc::kRows * ta.KPr * 2));                 /* slot tiles */
        M3D_CUDA(ctx, ctx->d_queue.reserve(sizeof(uint32_t) * (size_t)A.count * tc::kCandCap + 64));  /* candidates */
        M3D_CUDA(ctx, ctx->d_counts.reserve(sizeof(float) * 3 * (size_t)A.count + 64));               /* cut, slot_cut, cand_count */
        M3D_CUDA(ctx, ctx->d_counts_all.reserve(sizeof(uint32_t) * ((size_t)A.count + 4)));           /* list2 + its counter */
        float *d_cut = ctx->d_counts.as<float>();
        float *d_slot_cut = d_cut + A.count;
        uint32_t *d_cand_count = reinterpret_cast<uint32_t *>(d_slot_cut + A.count);
        uint32_t *d_list2_count = ctx->d_counts_all.as<uint32_t>();
        uint32_t *d_list2 = d_list2_count + 4;
        ta.cut = d_cut;
        M3D_CUDA(ctx, cudaMemsetAsync(d_list2_count, 0, sizeof(uint32_t), ctx->stream));
        M3D_CUDA(ctx, cudaFuncSetAttribute(tc::nn_top2_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        M3D_CUDA(ctx, cudaFuncSetAttribute(tc::nn_top2_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        if (two) {
            M3D_CUDA(ctx, cudaFuncSetAttribute(tc::nn_top2_tc2_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2));
            M3D_CUDA(ctx, cudaFuncSetAttribute(tc::nn_top2_tc2_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2));
            tc::nn_top2_tc2_kernel<false><<<atiles, tc::kTc2Threads, smem2, ctx->stream>>>(ta);
        } else {
            tc::nn_top2_tc_kernel<false><<<atiles / tc::kRB, 192, smem, ctx->stream>>>(ta);
        }
        M3D_LAUNCHED(ctx);
        /* rows too close to call: second GEMM pass over just those rows collecting every column within
         * the error bound of the best key, then the fp64 reference-order decision among the candidates */
        tc::gather_slots_kernel<<<ctx->sm_count * 4, 256, 0, ctx->stream>>>(A.tq, ta.KPr, d_amb, d_amb_count, d_cut,
                                                                        ctx->d_models.as<__nv_bfloat16>(), d_slot_cut,
                                                                        d_cand_count, A.count);
        M3D_LAUNCHED(ctx);
        tc::TcArgs tb = ta;
        tb.Aq = ctx->d_models.as<__nv_bfloat16>();
        tb.slot_count = d_amb_count;
        tb.slot_cut = d_slot_cut;
        tb.cand = ctx->d_queue.as<uint32_t>();
        tb.cand_count = d_cand_count;
        static const int splits_env = getenv("M3D_MATCH_SPLITS") ? atoi(getenv("M3D_MATCH_SPLITS")) : 0;
        tb.col_splits = splits_env > 0 ? (uint32_t)splits_env : 16u; /* 2: 144 ms, 4: 128, 8: 122, 16: 118, 32: 117 (200k FPFH descriptors) */
        if (two)
            tc::nn_top2_tc2_kernel<true><<<dim3(atiles, tb.col_splits), tc::kTc2Threads, smem2, ctx->stream>>>(tb);
        else
            tc::nn_top2_tc_kernel<true><<<dim3(atiles / tc::kRB, tb.col_splits), 192, smem, ctx->stream>>>(tb);
        M3D_LAUNCHED(ctx);
        cand_exact_kernel<<<ctx->sm_count * 2, 128, 0, ctx->stream>>>(A.f64, B.f64, dim, d_amb, d_amb_count, tb.cand,
                                                                     d_cand_count, d_nn, d_list2, d_list2_count);
        M3D_LAUNCHED(ctx);
        if (getenv("M3D_MATCH_DEBUG")) { /* candidate statistics of the direction (tuning aid) */
            cudaStreamSynchronize(ctx->stream);
            uint32_t namb = 0, nl2 = 0;
            cudaMemcpy(&namb, d_amb_count, 4, cudaMemcpyDeviceToHost);
            cudaMemcpy(&nl2, d_list2_count, 4, cudaMemcpyDeviceToHost);
            std::vector<uint32_t> cc(namb);
            if (namb) cudaMemcpy(cc.data(), d_cand_count, 4 * (size_t)namb, cudaMemcpyDeviceToHost);
            std::sort(cc.begin(), cc.end());
            auto q = [&](double f) { return namb ? cc[(size_t)(f * (namb - 1))] : 0u; };
            fprintf(stderr, "[m3d match] rows %u ambiguous %u overflow %u; candidates per ambiguous row: median %u p90 %u p99 %u max %u\n",
                    A.count, namb, nl2, q(0.5), q(0.9), q(0.99), q(1.0));
        }
        if (int rc = launch_exact(ctx, A, B, dim, d_list2, d_list2_count, d_nn)) return rc;
    } else if (path == 1) {
        const size_t smem = (size_t)3 * (KP + 1) * kMT * sizeof(float) + 3 * sizeof(uint64_t);
        M3D_CUDA(ctx, cudaFuncSetAttribute(nn_top2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        nn_top2_kernel<<<A.ntiles, 256, smem, ctx->stream>>>(A.tiles, B.tiles, A.count, B.count, KP, d_maxnorm_B,
                                                            d_nn, d_amb, d_amb_count);
        M3D_LAUNCHED(ctx);
        if (int rc = launch_exact(ctx, A, B, dim, d_amb, d_amb_count, d_nn)) return rc;
    } else {
        if (int rc = launch_exact(ctx, A, B, dim, nullptr, nullptr, d_nn)) return rc;
    }
    return M3D_OK;
}

/* uploads both descriptor sets, builds the fp32 tiles; returns device handles in A, B */
static int match_impl(m3d_ctx *ctx, const double *src, size_t ns, const double *dst, size_t nd, int dim,
                      bool both_directions, size_t *nn_out, size_t *idx0, size_t *idx1, size_t *n_out,
                      float *device_ms, bool on_device = false) {
    if (!ctx || dim <= 0 || (ns && !src) || (nd && !dst)) return M3D_ERR_INVALID_ARG;
    if (ns >= (1ull << 31) || nd >= (1ull << 31)) return ctx->fail(M3D_ERR_INVALID_ARG, "more than 2^31 descriptors");
    M3D_CUDA(ctx, cudaSetDevice(ctx->device));
    if (n_out) *n_out = 0;
    if (device_ms) *device_ms = 0;
    if (ns == 0 || nd == 0) return M3D_OK;
    const int KP = (dim + 3) & ~3;
    const int path = match_path(dim);
    const bool fast = path != 0;
    FeatDev A, B;
    A.count = (uint32_t)ns;
    B.count = (uint32_t)nd;
    A.ntiles = (A.count + kMT - 1) / kMT;
    B.ntiles = (B.count + kMT - 1) / kMT;
    const size_t fa = sizeof(double) * (size_t)dim * ns, fb = sizeof(double) * (size_t)dim * nd;
    const int KPr = tc::kprime(dim);
    /* tensor-core path: per set one query-form and one database-form tile array (bf16) */
    const uint32_t at = (A.ntiles + 1) / 2 * 2, bt = (B.ntiles + 1) / 2 * 2; /* query-form arrays are padded to whole CTA pairs */
    const size_t ta = path == 2 ? (size_t)2 * at * tc::kRows * KPr * 2
                                : (path == 1 ? sizeof(float) * (size_t)A.ntiles * (KP + 1) * kMT : 16);
    const size_t tb = path == 2 ? (size_t)2 * bt * tc::kRows * KPr * 2
                                : (path == 1 ? sizeof(float) * (size_t)B.ntiles * (KP + 1) * kMT : 16);
    M3D_CUDA(ctx, ctx->d_tmp5.reserve(sizeof(float) * (ns + nd) + 64));
    if (!on_device) {
        M3D_CUDA(ctx, ctx->d_tmp0.reserve(fa));
        M3D_CUDA(ctx, ctx->d_tmp1.reserve(fb));
    }
    M3D_CUDA(ctx, ctx->d_tmp2.reserve(ta));
    M3D_CUDA(ctx, ctx->d_tmp3.reserve(tb));
    M3D_CUDA(ctx, ctx->d_tmp4.reserve(sizeof(uint32_t) * 2 * (ns + nd) + 64));
    M3D_CUDA(ctx, ctx->d_small.reserve(sizeof(MatchScratch) + 4096));
    M3D_CUDA(ctx, ctx->h_small.reserve(sizeof(MatchScratch) + 4096));
    /* descriptors already on the device (m3d_features) are read in place */
    A.f64 = on_device ? const_cast<double *>(src) : ctx->d_tmp0.as<double>();
    B.f64 = on_device ? const_cast<double *>(dst) : ctx->d_tmp1.as<double>();
    A.tiles = ctx->d_tmp2.as<float>();
    B.tiles = ctx->d_tmp3.as<float>();
    A.tq = ctx->d_tmp2.as<__nv_bfloat16>();
    A.td = A.tq + (size_t)at * tc::kRows * KPr;
    B.tq = ctx->d_tmp3.as<__nv_bfloat16>();
    B.td = B.tq + (size_t)bt * tc::kRows * KPr;
    A.norms = ctx->d_tmp5.as<float>();
    B.norms = A.norms + ns;
    uint32_t *d_nn01 = ctx->d_tmp4.as<uint32_t>();
    uint32_t *d_nn10 = d_nn01 + ns;
    uint32_t *d_amb01 = d_nn10 + nd;
    uint32_t *d_amb10 = d_amb01 + ns;
    MatchScratch *sc = ctx->d_small.as<MatchScratch>();

    M3D_CUDA(ctx, cudaEventRecord(ctx->ev[0], ctx->stream));
    /* pageable descriptor arrays (numpy / Eigen) are staged through pinned memory piece by piece (context.cu) */
    if (!on_device) {
        if (int rc = host_to_device(ctx, A.f64, src, fa, ctx->stream)) return rc;
        if (int rc = host_to_device(ctx, B.f64, dst, fb, ctx->stream)) return rc;
    }
    M3D_CUDA(ctx, cudaMemsetAsync(sc, 0, sizeof(MatchScratch), ctx->stream));
    if (fast) {
        M3D_CUDA(ctx, cudaMemsetAsync(sc->mn, 0xff, sizeof(sc->mn), ctx->stream));
        const int gb = ctx->sm_count * 8;
        feat_minmax_kernel<<<gb, 256, 0, ctx->stream>>>(A.f64, (size_t)dim * ns, dim, sc->mn, sc->mx);
        M3D_LAUNCHED(ctx);
        feat_minmax_kernel<<<gb, 256, 0, ctx->stream>>>(B.f64, (size_t)dim * nd, dim, sc->mn, sc->mx);
        M3D_LAUNCHED(ctx);
        feat_center_kernel<<<1, 128, 0, ctx->stream>>>(sc->mn, sc->mx, dim, sc->center);
        M3D_LAUNCHED(ctx);
        if (path == 2) {
            const int drows = tc_two_cta() ? 64 : 128; /* rows per database tile */
            tc::feat_split_kernel<<<at, tc::kRows, 0, ctx->stream>>>(A.f64, A.count, dim, KPr, sc->center, 0, A.tq,
                                                                        A.norms, &sc->maxnorm[0], 128);
            M3D_LAUNCHED(ctx);
            tc::feat_split_kernel<<<B.ntiles, tc::kRows, 0, ctx->stream>>>(B.f64, B.count, dim, KPr, sc->center, 1, B.td,
                                                                        B.norms, &sc->maxnorm[1], drows);
            M3D_LAUNCHED(ctx);
            if (both_directions) {
                tc::feat_split_kernel<<<bt, tc::kRows, 0, ctx->stream>>>(B.f64, B.count, dim, KPr, sc->center, 0,
                                                                            B.tq, B.norms, &sc->maxnorm[1], 128);
                M3D_LAUNCHED(ctx);
                tc::feat_split_kernel<<<A.ntiles, tc::kRows, 0, ctx->stream>>>(A.f64, A.count, dim, KPr, sc->center, 1,
                                                                            A.td, A.norms, &sc->maxnorm[0], drows);
                M3D_LAUNCHED(ctx);
            }
        } else {
            feat_convert_kernel<<<A.ntiles, kMT, 0, ctx->stream>>>(A.f64, A.count, dim, KP, sc->center, A.tiles,
                                                                  &sc->maxnorm[0]);
            M3D_LAUNCHED(ctx);
            feat_convert_kernel<<<B.ntiles, kMT, 0, ctx->stream>>>(B.f64, B.count, dim, KP, sc->center, B.tiles,
                                                                  &sc->maxnorm[1]);
            M3D_LAUNCHED(ctx);
        }
    }
    if (int rc = nn_direction(ctx, A, B, dim, KP, path, &sc->maxnorm[1], d_nn01, d_amb01, &sc->amb_count[0])) return rc;
    if (both_directions) {
        if (int rc = nn_direction(ctx, B, A, dim, KP, path, &sc->maxnorm[0], d_nn10, d_amb10, &sc->amb_count[1]))
            return rc;
        /* mutual check + stable compaction */
        const uint32_t nblk = ((uint32_t)ns + kMB * kMItems - 1) / (kMB * kMItems);
        M3D_CUDA(ctx, ctx->d_blk.reserve(sizeof(uint32_t) * 2 * nblk + 64));
        M3D_CUDA(ctx, ctx->d_inl.reserve(sizeof(unsigned long long) * 2 * ns));
        uint32_t *blk_cnt = ctx->d_blk.as<uint32_t>(), *blk_off = blk_cnt + nblk;
        unsigned long long *d_i0 = ctx->d_inl.as<unsigned long long>(), *d_i1 = d_i0 + ns;
        mutual_count_kernel<<<nblk, kMB, 0, ctx->stream>>>(d_nn01, d_nn10, (uint32_t)ns, (uint32_t)nd, blk_cnt);
        M3D_LAUNCHED(ctx);
        mutual_scan_kernel<<<1, 32, 0, ctx->stream>>>(blk_cnt, nblk, blk_off, &sc->total);
        M3D_LAUNCHED(ctx);
        mutual_write_kernel<<<nblk, kMB, 0, ctx->stream>>>(d_nn01, d_nn10, (uint32_t)ns, (uint32_t)nd, blk_off, d_i0,
                                                          d_i1);
        M3D_LAUNCHED(ctx);
        MatchScratch *hs = ctx->h_small.as<MatchScratch>();
        M3D_CUDA(ctx, cudaMemcpyAsync(hs, sc, sizeof(MatchScratch), cudaMemcpyDeviceToHost, ctx->stream));
        M3D_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        const size_t m = hs->total;
        static_assert(sizeof(size_t) == sizeof(unsigned long long), "size_t must be 64-bit");
        if (m) {
            M3D_CUDA(ctx, cudaMemcpyAsync(idx0, d_i0, sizeof(size_t) * m, cudaMemcpyDeviceToHost, ctx->stream));
            M3D_CUDA(ctx, cudaMemcpyAsync(idx1, d_i1, sizeof(size_t) * m, cudaMemcpyDeviceToHost, ctx->stream));
        }
        M3D_CUDA(ctx, cudaEventRecord(ctx->ev[1], ctx->stream));
        M3D_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        *n_out = m;
    } else {
        std::vector<uint32_t> h(ns);
        M3D_CUDA(ctx, cudaMemcpyAsync(h.data(), d_nn01, sizeof(uint32_t) * ns, cudaMemcpyDeviceToHost, ctx->stream));
        M3D_CUDA(ctx, cudaEventRecord(ctx->ev[1], ctx->stream));
        M3D_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        for (size_t i = 0; i < ns; ++i) nn_out[i] = h[i];
    }
    if (device_ms) cudaEventElapsedTime(device_ms, ctx->ev[0], ctx->ev[1]);
    return M3D_OK;
}

}  // namespace m3d

extern "C" {

int m3d_match_correspondence(m3d_ctx *ctx, const double *src, size_t ns, const double *dst, size_t nd, int dim,
                             int method, int n_trees, size_t *idx0, size_t *idx1, size_t *n_out, float *device_ms) {
    (void)n_trees; /* Annoy forest size: the exact search has no such parameter */
    if (!ctx || !n_out || (ns && (!idx0 || !idx1))) return M3D_ERR_INVALID_ARG;
    if (method != M3D_MATCH_FLANN && method != M3D_MATCH_ANNOY)
        return ctx->fail(M3D_ERR_INVALID_ARG, "unknown match method %d", method);
    return m3d::match_impl(ctx, src, ns, dst, nd, dim, true, nullptr, idx0, idx1, n_out, device_ms);
}

int m3d_match_features(m3d_ctx *ctx, const m3d_features *a, const m3d_features *b, size_t *idx0, size_t *idx1,
                       size_t *n_out, float *device_ms) {
    if (!ctx || !a || !b || !n_out || (a->n && (!idx0 || !idx1))) return M3D_ERR_INVALID_ARG;
    if (a->dim != b->dim) return ctx->fail(M3D_ERR_INVALID_ARG, "descriptor sets of dimension %d and %d", a->dim, b->dim);
    if (a->ctx != ctx || b->ctx != ctx) return ctx->fail(M3D_ERR_INVALID_ARG, "features belong to another context");
    return m3d::match_impl(ctx, a->data.as<double>(), a->n, b->data.as<double>(), b->n, a->dim, true, nullptr, idx0, idx1,
                           n_out, device_ms, true);
}

int m3d_nearest(m3d_ctx *ctx, const double *src, size_t ns, const double *dst, size_t nd, int dim, size_t *nn,
                float *device_ms) {
    if (!ctx || (ns && !nn)) return M3D_ERR_INVALID_ARG;
    return m3d::match_impl(ctx, src, ns, dst, nd, dim, false, nn, nullptr, nullptr, nullptr, device_ms);
}

} /* extern "C" */
