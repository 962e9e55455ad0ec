/*
 * matching_tc.cuh -- tensor-core (tcgen05 / TMEM / TMA) variant of the brute-force nearest-neighbour
 * search of matching.cu.
 *
 * The squared distance is recast as a dense bf16 GEMM with fp32 accumulation in tensor memory:
 * every centred fp32 component x is split into three bf16 terms x = h + m + l (24 mantissa bits),
 * and the six significant cross products of a.b are laid out along K:
 *     query form    Q(a) = [ h, h, m, h, l, m | 1, 1, 1 | 0.. ]
 *     database form D(b) = [-2h,-2m,-2h,-2l,-2h,-2m | n_h, n_m, n_l | 0.. ]      (n = |b|^2 split the same way)
 * so that  Q(a).D(b) = |b|^2 - 2 a.b  up to ~2^-24 relative to |a||b| -- the ranking key of row a
 * (|a|^2 is constant per row).  K' = 6*dim + 3 padded to a multiple of 16 (dim 33 -> 208 = 13 MMAs of
 * K = 16).  Rows whose best / second-best gap is inside the error bound are re-searched in fp64 by
 * nn_exact_kernel exactly as for the fp32 kernel, so the result stays bit-identical to the oracle.
 *
 * One CTA = one 128-row block of queries x all database tiles:
 *   warp 0      TMA producer  (1-D cp.async.bulk of pre-tiled operands, 3-stage ring)
 *   warp 1      TMEM allocator + single-thread tcgen05.mma issuer (M128 x N128 x K16, kind::f16, bf16 in, f32 out)
 *   warps 2-5   epilogue: tcgen05.ld 32 lanes x 32 columns -> registers, running (best, index, second best)
 * Operands sit in shared memory in the canonical no-swizzle K-major UMMA layout
 *   element (row r, k) at  (k/8)*2048 + r*16 + (k%8)*2  bytes  (LBO = 2048 B, SBO = 128 B),
 * which is also the global tile layout, so one bulk copy moves a whole 128 x K' tile.
 */
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <cstdint>


namespace m3d {
namespace tc {

constexpr int kRows = 128;          /* rows of a tile (UMMA M and N)            */
constexpr int kChunkBytes = 2048;   /* one 8-wide K chunk of 128 rows           */
constexpr int kRB = 1;              /* query row blocks per CTA                  */
constexpr int kBStages = 3;         /* database-tile ring depth                  */
constexpr int kMaxKPrime = 208;     /* (kRB + kBStages) tiles of 128 x K' bf16 must fit in shared memory */

__host__ __device__ inline int kprime(int dim) { return ((6 * dim + 3 + 15) / 16) * 16; }

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
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
__device__ __forceinline__ void tma_load_1d(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
    const uint32_t b = smem_u32(bar);
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(bytes) : "memory");
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
            smem_u32(dst)),
        "l"(src), "r"(bytes), "r"(b)
        : "memory");
}

/* shared-memory matrix descriptor, K-major, no swizzle (cute::UMMA::SmemDescriptor bit layout):
 * [0,14) start>>4 | [16,30) leading byte offset>>4 | [32,46) stride byte offset>>4 | [46,48) version=1 */
__device__ __forceinline__ uint64_t umma_desc(uint32_t smem_addr, uint32_t lbo_bytes = kChunkBytes) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3fff);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3fff) << 16;   /* LBO: next 8-wide K chunk (rows of the tile x 16 B) */
    d |= (uint64_t)((128 >> 4) & 0x3fff) << 32;         /* SBO: next group of 8 rows         */
    d |= (uint64_t)1 << 46;                             /* descriptor version (Blackwell)    */
    return d;
}
/* instruction descriptor (cute::UMMA::InstrDescriptor): D f32, A/B bf16, both K-major, M x N */
__host__ __device__ constexpr uint32_t umma_idesc(int M, int N) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                          uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
/* 32 lanes x 32 columns of fp32 accumulators -> 32 registers per thread (thread = TMEM lane = row) */
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]); /* valid after tmem_ld_wait() */
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

/* fp64 column-major descriptors -> bf16x3 split tiles in the UMMA layout.  role 0 = query form,
 * 1 = database form.  Also writes the fp32 squared norms and their max. */
__global__ void __launch_bounds__(128) feat_split_kernel(const double *__restrict__ F, uint32_t count, int dim,
                                                         int KPr, const double *__restrict__ center, int role,
                                                         __nv_bfloat16 *__restrict__ tiles,
                                                         float *__restrict__ norms,
                                                         uint32_t *__restrict__ maxnorm_bits, int tile_rows) {
    /* tile_rows = rows per tile of the output layout: 128, or 64 for the database form of the two-CTA kernel (each CTA
     * of a pair holds half of a 128-column database tile) */
    const uint32_t j = blockIdx.x * kRows + threadIdx.x;
    const uint32_t tile = j / (uint32_t)tile_rows, r = j % (uint32_t)tile_rows;
    __nv_bfloat16 *t = tiles + (size_t)tile * tile_rows * KPr;
    const size_t chunk = (size_t)tile_rows * 8; /* elements per 8-wide K chunk of a tile */
    auto put = [&](int kk, float v) { t[(size_t)(kk >> 3) * chunk + r * 8 + (kk & 7)] = __float2bfloat16_rn(v); };
    const float sc = role ? -2.f : 1.f;
    double n2 = 0;
    if (j < count) {
        for (int k = 0; k < dim; ++k) {
            const double xd = F[(size_t)j * dim + k] - center[k];
            n2 += xd * xd;
            const float x = (float)xd;
            const float h = __bfloat162float(__float2bfloat16_rn(x));
            const float r1 = x - h;
            const float m = __bfloat162float(__float2bfloat16_rn(r1));
            const float l = r1 - m; /* rounded to bf16 by put() */
            if (role == 0) {
                put(0 * dim + k, h);
                put(1 * dim + k, h);
                put(2 * dim + k, m);
                put(3 * dim + k, h);
                put(4 * dim + k, l);
                put(5 * dim + k, m);
            } else {
                put(0 * dim + k, sc * h);
                put(1 * dim + k, sc * m);
                put(2 * dim + k, sc * h);
                put(3 * dim + k, sc * l);
                put(4 * dim + k, sc * h);
                put(5 * dim + k, sc * m);
            }
        }
        const float nf = (float)n2;
        if (role == 0) {
            put(6 * dim + 0, 1.f);
            put(6 * dim + 1, 1.f);
            put(6 * dim + 2, 1.f);
        } else {
            const float h = __bfloat162float(__float2bfloat16_rn(nf));
            const float r1 = nf - h;
            const float m = __bfloat162float(__float2bfloat16_rn(r1));
            put(6 * dim + 0, h);
            put(6 * dim + 1, m);
            put(6 * dim + 2, r1 - m);
        }
        for (int kk = 6 * dim + 3; kk < KPr; ++kk) put(kk, 0.f);
        norms[j] = nf;
        if (nf == nf && nf < 3e38f) atomicMax(maxnorm_bits, __float_as_uint(nf));
    } else { /* padding rows: zero vector; as a database column its "norm" is huge so it never wins */
        for (int kk = 0; kk < KPr; ++kk) put(kk, 0.f);
        if (role == 1) put(6 * dim + 0, 1e30f);
    }
}

struct TcArgs {
    const __nv_bfloat16 *Aq; /* query-form tiles of the rows   */
    const __nv_bfloat16 *Bd; /* database-form tiles of the columns */
    const float *a_norms;    /* |a|^2 per row (fp32)           */
    uint32_t na, nb;
    int KPr;                 /* padded K'                      */
    const uint32_t *maxnorm_bits;
    uint32_t *nn, *amb_list, *amb_count;
    float *cut;              /* [na] first pass: best key + 2.5 E of every row (candidate cut-off)      */
    /* candidate pass (CAND): rows are the compacted ambiguous rows ("slots") */
    const uint32_t *slot_count; /* number of slots (device)                                            */
    const float *slot_cut;      /* [slots] cut-off of the slot's row                                   */
    uint32_t *cand;             /* [slots][kCandCap] database columns with key <= cut                   */
    uint32_t *cand_count;       /* [slots]                                                              */
    uint32_t col_splits;        /* gridDim.y: each CTA scans 1/col_splits of the database tiles         */
};
constexpr int kCandCap = 256; /* rows with more columns inside the error bound go to the full fp64 search: real FPFH
                                  * descriptors cluster (near-identical patches), 32 sent most rows there */

/* one tile of the top-2 scan: 128 fp32 keys of this thread's row (4 x 32 TMEM columns) */
template <int G>
__device__ __forceinline__ void scan_tile(const float (&v)[G][32], uint32_t jbase, uint32_t ncol, float &m1, float &m2,
                                          uint32_t &i1, uint32_t col0 = 0) {
#pragma unroll
    for (int g = 0; g < G; ++g) {
        const uint32_t c0 = col0 + 32u * g;
        /* cheap screen: min of the 32 values (FMNMX3 tree); the running (best, second best) changes
         * O(log n) times per row, so the update below is the rare path */
        float lo = fminf(fminf(v[g][0], v[g][1]), v[g][2]);
#pragma unroll
        for (int i = 3; i + 1 < 32; i += 2) lo = fminf(fminf(lo, v[g][i]), v[g][i + 1]);
        lo = fminf(lo, v[g][31]);
        if (c0 + 32 <= ncol && !(lo < m2)) continue;
#pragma unroll
        for (int i = 0; i < 32; ++i) {
            const float d = v[g][i];
            if (c0 + i < ncol && d < m2) {
                if (d < m1) {
                    m2 = m1;
                    m1 = d;
                    i1 = jbase + c0 + i;
                } else {
                    m2 = d;
                }
            }
        }
    }
}

/* candidate pass: every column whose key is within the cut-off of this row */
template <int G>
__device__ __forceinline__ void collect_tile(const float (&v)[G][32], uint32_t jbase, uint32_t ncol, float cut,
                                             uint32_t *cand, uint32_t *cand_count, uint32_t col0 = 0) {
#pragma unroll
    for (int g = 0; g < G; ++g) {
        const uint32_t c0 = col0 + 32u * g;
        float lo = fminf(fminf(v[g][0], v[g][1]), v[g][2]);
#pragma unroll
        for (int i = 3; i + 1 < 32; i += 2) lo = fminf(fminf(lo, v[g][i]), v[g][i + 1]);
        lo = fminf(lo, v[g][31]);
        if (!(lo <= cut)) continue;
#pragma unroll
        for (int i = 0; i < 32; ++i) {
            if (c0 + i < ncol && v[g][i] <= cut) {
                const uint32_t k = atomicAdd(cand_count, 1u); /* column ranges of one row run in several CTAs */
                if (k < (uint32_t)kCandCap) cand[k] = jbase + c0 + i;
            }
        }
    }
}

/* One CTA = kRB query row blocks (2 x 128 rows) x all database tiles: every database tile fetched
 * from L2 feeds 2 x 13 MMAs, which halves the L2 -> SM operand traffic (the limiter with one row
 * block per CTA).  TMEM: (2 buffers) x (kRB row blocks) x 128 columns = all 512 columns. */
template <bool CAND>
__global__ void __launch_bounds__(192, 1) nn_top2_tc_kernel(const TcArgs a) {
    if (CAND && blockIdx.x * kRB * kRows >= *a.slot_count) return; /* fewer ambiguous rows than the grid was sized for */
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const uint32_t tile_bytes = (uint32_t)kRows * a.KPr * 2;
    unsigned char *As = smem_raw;
    unsigned char *Bs = smem_raw + kRB * (size_t)tile_bytes;
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem_raw + (kRB + kBStages) * (size_t)tile_bytes);
    uint64_t *a_full = bars, *b_full = bars + 1, *b_empty = b_full + kBStages, *t_full = b_empty + kBStages,
             *t_empty = t_full + 2;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(t_empty + 2);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t ntb_all = (a.nb + kRows - 1) / kRows;
    /* database tiles of this CTA: all of them, or one of col_splits contiguous ranges (CAND) */
    const uint32_t per = CAND ? (ntb_all + a.col_splits - 1) / a.col_splits : ntb_all;
    const uint32_t tb0 = CAND ? min(ntb_all, blockIdx.y * per) : 0u;
    const uint32_t ntb = min(ntb_all, tb0 + per) - tb0;
    const int nk = a.KPr / 16;

    if (tid == 0) {
        mbar_init(a_full, 1);
        for (int s = 0; s < kBStages; ++s) {
            mbar_init(&b_full[s], 1);
            mbar_init(&b_empty[s], 1);
        }
        for (int s = 0; s < 2; ++s) {
            mbar_init(&t_full[s], 1);
            mbar_init(&t_empty[s], 4); /* one arrive per epilogue warp */
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                     "r"(512)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) { /* ---------------- TMA producer (the query tile array is padded to a multiple of kRB tiles) */
            tma_load_1d(As, a.Aq + (size_t)blockIdx.x * kRB * kRows * a.KPr, kRB * tile_bytes, a_full);
            for (uint32_t t = 0; t < ntb; ++t) {
                const int st = t % kBStages;
                mbar_wait(&b_empty[st], ((t / kBStages) & 1) ^ 1);
                tma_load_1d(Bs + (size_t)st * tile_bytes, a.Bd + (size_t)(tb0 + t) * kRows * a.KPr, tile_bytes, &b_full[st]);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) { /* ---------------- MMA issuer */
            const uint32_t idesc = umma_idesc(128, 128);
            mbar_wait(a_full, 0);
            for (uint32_t t = 0; t < ntb; ++t) {
                const int st = t % kBStages, buf = t & 1;
                mbar_wait(&b_full[st], (t / kBStages) & 1);
                mbar_wait(&t_empty[buf], ((t >> 1) & 1) ^ 1);
                tc_fence_after();
                const uint32_t b0 = smem_u32(Bs + (size_t)st * tile_bytes);
#pragma unroll
                for (int rb = 0; rb < kRB; ++rb) {
                    const uint32_t a0 = smem_u32(As + (size_t)rb * tile_bytes);
                    const uint32_t d0 = tmem_base + (uint32_t)(buf * kRB + rb) * 128u;
                    for (int ks = 0; ks < nk; ++ks)
                        umma_bf16(d0, umma_desc(a0 + ks * 2 * kChunkBytes), umma_desc(b0 + ks * 2 * kChunkBytes), idesc,
                                  ks > 0 ? 1u : 0u);
                }
                umma_commit(&b_empty[st]); /* smem slot reusable once these MMAs have read it */
                umma_commit(&t_full[buf]); /* both accumulators ready for the epilogue       */
            }
        }
    } else { /* ---------------- epilogue warps 2..5: TMEM lane quarter = warp % 4; one row per row block per thread */
        const uint32_t q = (uint32_t)warp & 3u;
        float m1[kRB], m2[kRB], cutv[kRB];
        uint32_t i1[kRB];
        const uint32_t nslots = CAND ? *a.slot_count : 0u;
#pragma unroll
        for (int rb = 0; rb < kRB; ++rb) {
            m1[rb] = INFINITY;
            m2[rb] = INFINITY;
            i1[rb] = 0;
            cutv[rb] = -INFINITY;
            if (CAND) {
                const uint32_t slot = (blockIdx.x * kRB + rb) * kRows + q * 32 + lane;
                if (slot < nslots) cutv[rb] = a.slot_cut[slot];
            }
        }
        for (uint32_t t = 0; t < ntb; ++t) {
            const int buf = t & 1;
            mbar_wait(&t_full[buf], (t >> 1) & 1);
            tc_fence_after();
            const uint32_t jbase = (tb0 + t) * kRows;
            const uint32_t ncol = min((uint32_t)kRows, a.nb - jbase);
#pragma unroll
            for (int rb = 0; rb < kRB; ++rb) {
                float v[4][32];
                const uint32_t tbase = tmem_base + ((q * 32u) << 16) + (uint32_t)(buf * kRB + rb) * 128u;
#pragma unroll
                for (int g = 0; g < 4; ++g) tmem_ld32(tbase + 32u * g, v[g]); /* four loads in flight */
                tmem_ld_wait();
                if (rb == kRB - 1) { /* the buffer pair is free as soon as the values sit in registers */
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&t_empty[buf]);
                }
                if (CAND) {
                    const uint32_t slot = (blockIdx.x * kRB + rb) * kRows + q * 32 + lane;
                    if (slot < nslots) collect_tile<4>(v, jbase, ncol, cutv[rb], a.cand + (size_t)slot * kCandCap,
                                                    a.cand_count + slot);
                } else {
                    scan_tile<4>(v, jbase, ncol, m1[rb], m2[rb], i1[rb]);
                }
            }
        }
        if (!CAND) {
            const float bnmax = __uint_as_float(*a.maxnorm_bits);
#pragma unroll
            for (int rb = 0; rb < kRB; ++rb) {
                const uint32_t row = (blockIdx.x * kRB + rb) * kRows + q * 32 + lane;
                if (row < a.na) {
                    a.nn[row] = i1[rb];
                    /* |GEMM key - exact key| <= E: fp32 input rounding, dropped split
                     * terms, and 2 K' truncating fp32 accumulations of partial sums <= 2(|a|^2 + |b|^2) */
                    const float E = (float)(2 * a.KPr + 64) * 1.1920929e-07f * (a.a_norms[row] + bnmax);
                    a.cut[row] = m1[rb] + 2.5f * E;
                    if (!(m2[rb] - m1[rb] > 2.5f * E)) a.amb_list[atomicAdd(a.amb_count, 1u)] = row;
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
    }
}


/* ------------------------------------------------------------------------------------------------------------------
 * The two-CTA form (tcgen05 cta_group::2): a CLUSTER OF TWO CTAs (one SM pair) computes a 256 x 128 block of keys per
 * database tile.  Each CTA keeps its own 128 query rows (A) and loads only HALF of every database tile -- 64 of the 128
 * columns (B) -- and the pair's tensor cores read both halves: the L2 -> SM operand stream per SM is halved, which is
 * what bounded the one-CTA kernel (53 KB per 128 x 128 tile, ~9.3 TB/s chip-wide, tensor pipe 56 %).
 *   both CTAs   warp 0: TMA producer of the CTA's A tile and B halves (6-stage ring); warps 2-5: epilogue on the CTA's
 *               own accumulator (its 128 rows x 128 columns in its own TMEM)
 *   leader      warp 1: single-thread tcgen05.mma.cta_group::2 issuer (M 256 x N 128 x K 16); tcgen05.commit multicast
 *               to both CTAs frees the ring slot / publishes the accumulator in both
 *   peer        warp 1: relay -- forwards "my half has landed" to the leader's barriers (remote mbarrier arrive)
 * The peer's epilogue warps release the accumulator buffer on the LEADER's barrier (the leader issues the MMAs that
 * overwrite it).  Database tiles are stored as 64-row tiles (feat_split_kernel, tile_rows = 64). */
constexpr int kB2Stages = 6;
constexpr int kTc2Threads = 320; /* producer warp, MMA / relay warp, 8 epilogue warps */
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
/* arrive on the mbarrier at this CTA-local address in CTA `rank` of the cluster */
__device__ __forceinline__ void mbar_arrive_remote(uint64_t *bar, uint32_t rank) {
    uint32_t remote;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(smem_u32(bar)), "r"(rank));
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(remote) : "memory");
}
__device__ __forceinline__ void mbar_wait_cluster(uint64_t *bar, uint32_t parity) { /* acquire at cluster scope */
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "LAB_WAIT_C:\n"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra DONE_C;\n"
        "bra LAB_WAIT_C;\n"
        "DONE_C:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void umma2_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma2_commit(uint64_t *bar, uint16_t cta_mask) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                     smem_u32(bar)),
                 "h"(cta_mask)
                 : "memory");
}

template <bool CAND>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kTc2Threads, 1) nn_top2_tc2_kernel(const TcArgs a) {
    /* fewer ambiguous rows than the grid was sized for: decided per pair (a CTA that left could not be waited for) */
    if (CAND && (blockIdx.x / 2) * 2 * kRows >= *a.slot_count) return;
    constexpr int kAcc = 4; /* accumulator buffers: all 512 TMEM columns */
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const uint32_t crank = cluster_ctarank(); /* 0 = leader */
    const uint32_t tile_bytes = (uint32_t)kRows * a.KPr * 2;
    const uint32_t half_bytes = tile_bytes / 2; /* 64 database rows */
    unsigned char *As = smem_raw;
    unsigned char *Bs = smem_raw + tile_bytes;
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem_raw + tile_bytes + (size_t)kB2Stages * half_bytes);
    uint64_t *a_full = bars, *peer_a = bars + 1, *b_full = bars + 2, *peer_b = b_full + kB2Stages,
             *b_empty = peer_b + kB2Stages, *t_full = b_empty + kB2Stages, *t_empty = t_full + kAcc;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(t_empty + kAcc);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t ntb_all = (a.nb + kRows - 1) / kRows;
    const uint32_t per = CAND ? (ntb_all + a.col_splits - 1) / a.col_splits : ntb_all;
    const uint32_t tb0 = CAND ? min(ntb_all, blockIdx.y * per) : 0u;
    const uint32_t ntb = min(ntb_all, tb0 + per) - tb0;
    const int nk = a.KPr / 16;
    /* FOUR accumulator buffers: the release of a buffer crosses the pair (the peer's epilogue arrives on the leader's
     * barrier), and with two buffers that round trip sat on the MMA issuer's critical path */

    if (tid == 0) {
        mbar_init(a_full, 1);
        mbar_init(peer_a, 1);
        for (int s = 0; s < kB2Stages; ++s) {
            mbar_init(&b_full[s], 1);
            mbar_init(&peer_b[s], 1);
            mbar_init(&b_empty[s], 1); /* one multicast commit of the leader per use */
        }
        for (int s = 0; s < kAcc; ++s) {
            mbar_init(&t_full[s], 1);
            mbar_init(&t_empty[s], 16); /* 8 epilogue warps of each CTA (used in the leader only) */
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(kAcc * 128)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all(); /* both CTAs' barriers and accumulators exist before anybody signals across */
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) { /* ---------------- TMA producer: own query rows, own half of every database tile */
            tma_load_1d(As, a.Aq + (size_t)blockIdx.x * kRows * a.KPr, tile_bytes, a_full);
            for (uint32_t t = 0; t < ntb; ++t) {
                const int st = t % kB2Stages;
                mbar_wait(&b_empty[st], ((t / kB2Stages) & 1) ^ 1);
                tma_load_1d(Bs + (size_t)st * half_bytes,
                            reinterpret_cast<const unsigned char *>(a.Bd) + ((size_t)(tb0 + t) * 2 + crank) * half_bytes,
                            half_bytes, &b_full[st]);
            }
        }
    } else if (warp == 1) {
        if (lane == 0 && crank == 0) { /* ---------------- MMA issuer (leader) */
            const uint32_t idesc = umma_idesc(256, 128);
            mbar_wait(a_full, 0);
            mbar_wait(peer_a, 0);
            for (uint32_t t = 0; t < ntb; ++t) {
                const int st = t % kB2Stages, buf = t % kAcc;
                mbar_wait(&b_full[st], (t / kB2Stages) & 1);
                mbar_wait(&peer_b[st], (t / kB2Stages) & 1);
                mbar_wait(&t_empty[buf], ((t / kAcc) & 1) ^ 1);
                tc_fence_after();
                const uint32_t b0 = smem_u32(Bs + (size_t)st * half_bytes);
                const uint32_t a0 = smem_u32(As);
                const uint32_t d0 = tmem_base + (uint32_t)buf * 128u;
                for (int ks = 0; ks < nk; ++ks)
                    umma2_bf16(d0, umma_desc(a0 + ks * 2 * kChunkBytes, kChunkBytes),
                               umma_desc(b0 + ks * 2 * (kChunkBytes / 2), kChunkBytes / 2), idesc, ks > 0 ? 1u : 0u);
                umma2_commit(&b_empty[st], 0x3); /* the slot is free in BOTH CTAs once these MMAs have read it */
                umma2_commit(&t_full[buf], 0x3); /* both CTAs' accumulators are ready                          */
            }
        } else if (lane == 0) { /* ---------------- relay (peer): tell the leader when my operands have landed */
            mbar_wait(a_full, 0);
            mbar_arrive_remote(peer_a, 0);
            for (uint32_t t = 0; t < ntb; ++t) {
                const int st = t % kB2Stages;
                mbar_wait(&b_full[st], (t / kB2Stages) & 1);
                mbar_arrive_remote(&peer_b[st], 0);
            }
        }
    } else { /* ---------------- epilogue warps 2..9: TMEM lane quarter = warp % 4 (a warp can only read its quarter),
              * column half = (warp - 2) / 4: two warps share a quarter and scan 64 of the 128 columns each.  (With four
              * warps the scan of 128 keys per row and tile paced the whole kernel: 12.8 ms with the MMAs and the
              * operand traffic switched off.) */
        const uint32_t q = (uint32_t)warp & 3u, half = (uint32_t)(warp - 2) >> 2;
        float m1 = INFINITY, m2 = INFINITY, cutv = -INFINITY;
        uint32_t i1 = 0;
        const uint32_t nslots = CAND ? *a.slot_count : 0u;
        const uint32_t my = blockIdx.x * kRows + q * 32 + lane; /* row (or slot) of this thread */
        if (CAND && my < nslots) cutv = a.slot_cut[my];
        for (uint32_t t = 0; t < ntb; ++t) {
            const int buf = t % kAcc;
            mbar_wait(&t_full[buf], (t / kAcc) & 1);
            tc_fence_after();
            const uint32_t jbase = (tb0 + t) * kRows;
            const uint32_t ncol = min((uint32_t)kRows, a.nb - jbase);
            float v[2][32];
            const uint32_t tbase = tmem_base + ((q * 32u) << 16) + (uint32_t)buf * 128u + half * 64u;
#pragma unroll
            for (int g = 0; g < 2; ++g) tmem_ld32(tbase + 32u * g, v[g]);
            tmem_ld_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) { /* the buffer is free as soon as the values sit in registers: tell the leader */
                if (crank == 0) mbar_arrive(&t_empty[buf]);
                else mbar_arrive_remote(&t_empty[buf], 0);
            }
            if (CAND) {
                if (my < nslots) collect_tile<2>(v, jbase, ncol, cutv, a.cand + (size_t)my * kCandCap, a.cand_count + my, half * 64u);
            } else {
                scan_tile<2>(v, jbase, ncol, m1, m2, i1, half * 64u);
            }
        }
        if (!CAND) { /* merge the two column halves of every row (through the idle operand ring) */
            float4 *xch = reinterpret_cast<float4 *>(Bs);
            __syncwarp();
            asm volatile("bar.sync 3, 256;" ::: "memory"); /* all epilogue warps are done with the last tile: ring is idle */
            if (half == 1) xch[q * 32 + lane] = make_float4(m1, m2, __uint_as_float(i1), 0.f);
            asm volatile("bar.sync 3, 256;" ::: "memory");
            if (half == 0 && my < a.na) {
                const float4 o = xch[q * 32 + lane];
                /* the two smallest of {m1, m2, o.x, o.y}; an exact tie of the minima leaves a gap of 0 = ambiguous */
                if (o.x < m1) {
                    m2 = fminf(m1, o.y);
                    m1 = o.x;
                    i1 = __float_as_uint(o.z);
                } else {
                    m2 = fminf(m2, o.x);
                }
                const float bnmax = __uint_as_float(*a.maxnorm_bits);
                a.nn[my] = i1;
                const float E = (float)(2 * a.KPr + 64) * 1.1920929e-07f * (a.a_norms[my] + bnmax);
                a.cut[my] = m1 + 2.5f * E;
                if (!(m2 - m1 > 2.5f * E)) a.amb_list[atomicAdd(a.amb_count, 1u)] = my;
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all(); /* nobody leaves while the pair may still signal it or read its shared memory */
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kAcc * 128) : "memory");
    }
}

/* compacts the query-form vectors of the ambiguous rows into fresh tiles ("slots") for the
 * candidate pass; thread = (slot, 8-wide K chunk) */
__global__ void __launch_bounds__(256) gather_slots_kernel(const __nv_bfloat16 *__restrict__ Aq, int KPr,
                                                           const uint32_t *__restrict__ amb_list,
                                                           const uint32_t *__restrict__ amb_count,
                                                           const float *__restrict__ cut,
                                                           __nv_bfloat16 *__restrict__ Aq2, float *__restrict__ slot_cut,
                                                           uint32_t *__restrict__ cand_count, uint32_t max_slots) {
    const uint32_t n = min(*amb_count, max_slots);
    const uint32_t npad = (n + kRows - 1) / kRows * kRows;
    const int KC = KPr / 8;
    const size_t total = (size_t)npad * KC;
    for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (size_t)gridDim.x * blockDim.x) {
        const uint32_t slot = (uint32_t)(e / KC), kc = (uint32_t)(e % KC);
        uint4 val = make_uint4(0, 0, 0, 0);
        if (slot < n) {
            const uint32_t row = amb_list[slot];
            val = *reinterpret_cast<const uint4 *>(Aq + (size_t)(row / kRows) * kRows * KPr + (size_t)kc * (kChunkBytes / 2) +
                                                   (row % kRows) * 8);
            if (kc == 0) {
                slot_cut[slot] = cut[row];
                cand_count[slot] = 0;
            }
        }
        *reinterpret_cast<uint4 *>(Aq2 + (size_t)(slot / kRows) * kRows * KPr + (size_t)kc * (kChunkBytes / 2) +
                                   (slot % kRows) * 8) = val;
    }
}

}  // namespace tc
}  // namespace m3d
