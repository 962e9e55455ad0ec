/*
 * score_cell.cuh -- the default hot kernel: hierarchical inlier scoring with per-cell hypothesis lists
 * (reference: RANSAC<>::EvaluateModel, include/misc3d/common/ransac.h:626-654, once per hypothesis).
 *
 * Same cloud layout and the same conservative bounding-sphere tests as score_cull.cuh (Morton-ordered
 * tiles of 32 cells x 32 points, one TMA bulk copy per tile), same fp32 guard-banded point test and
 * fp64 resolve queue, hence the same bit-identical counts.  What changes is who shares a load.
 * score_cull_kernel walks the surviving cells of ONE hypothesis with lane = point: every surviving
 * (hypothesis, cell) pair costs one LDS.128 per lane, and that shared-memory traffic bound it (ncu,
 * profiles/r01_ncu_score_cull_run14.md: 51-72 % of the wavefront peak at 27 % occupancy).  Here each
 * tile is processed in two phases by the CTA's consumer warps:
 *
 *   phase 1  lane = hypothesis against the tile sphere, then lane = cell for every surviving
 *            hypothesis; a lane whose cell survives APPENDS the hypothesis to that cell's list
 *            (shared-memory atomic cursor per cell) -- the (hypothesis, cell) incidence is transposed
 *            for free, because the lane that finds the survivor owns the list;
 *   phase 2  lane = hypothesis: a warp takes a cell and up to 64 hypotheses of its list (two per lane,
 *            packed FFMA2 arithmetic), gathers their coefficients once, and streams the cell's 32 points
 *            as BROADCAST loads -- the inner loop of the dense kernel (one LDS.128 wavefront per 32-64
 *            point-hypothesis pairs instead of four per 32), run only on the surviving pairs.
 *
 * A hypothesis that passes through most of the cloud simply sits in most lists and is evaluated at
 * dense-kernel cost, so no pre-classification of the hypotheses (cull_classify_kernel) is needed.
 * Two consumer-only named barriers per tile separate the phases (see the pipeline at the end of the kernel);
 * the producer warp keeps a 2-3 stage TMA ring filled (a tile takes ~10 us of CTA time, far above the copy
 * latency) and parks on its mbarrier with a suspend-time hint instead of spinning.
 */
#pragma once
#include "score_cull.cuh"

namespace m3d {

/* counters of the statistics build of the kernel (STATS = true): what the bench reports as the work the
 * kernel actually did.  [0] (hypothesis, tile) tests, [1] (hypothesis, tile) survivors, [2] (hypothesis,
 * cell) list entries, [3] one-per-lane passes, [4] two-per-lane passes, [5] guard-band re-scans */
__device__ unsigned long long g_cell_stats[8];

template <int THREADS>
struct CellCfg {
    static constexpr int kStages = THREADS >= 384 ? 3 : 2;     /* TMA ring depth                              */
    static constexpr int kMinBlocks = THREADS >= 384 ? 1 : (THREADS >= 192 ? 2 : 4);
};

template <int KIND, int THREADS, int NH>
struct CellSmem {
    static constexpr int kStages = CellCfg<THREADS>::kStages;
    static constexpr int kPlanes = KIND == kPlane ? 2 : (KIND == kSphere ? 3 : 4); /* float4 planes of hypothesis parameters */
    static constexpr int kListStride = NH + 2; /* u16 entries per cell list; +2 shifts consecutive cells by one bank */
    static constexpr size_t kListBytes = ((size_t)kTileCells * kListStride * sizeof(uint16_t) + 15) & ~(size_t)15;
    static constexpr size_t kSurvBytes = ((size_t)NH * sizeof(uint16_t) + 15) & ~(size_t)15;
    static constexpr size_t bytes() {
        return (size_t)kStages * kBlobF4 * sizeof(float4)             /* tile ring                    */
               + (size_t)kPlanes * (NH + 1) * sizeof(float4)          /* parameters (+ one dummy row) */
               + kListBytes                                            /* per-cell hypothesis lists    */
               + kSurvBytes                                            /* hypotheses surviving the tile test */
               + (size_t)(THREADS / 32) * kQBuf * sizeof(uint2)       /* guard-band staging           */
               + 2 * kStages * sizeof(uint64_t)                       /* full / empty barriers        */
               + (size_t)NH * sizeof(uint32_t)                        /* per-hypothesis counts        */
               + (kStages + 2 * kTileCells + 6) * sizeof(uint32_t)    /* tile ids, list lengths x2, cursors */
               + 32;
    }
};

/* parameters of hypothesis h from the SoA planes: {c0..c3} [{c4..c7}] {T, band, cull.a, cull.b} [{cull.c, cull.d}] */
template <int KIND, int NH>
__device__ __forceinline__ void load_fast(const float4 *hyp, uint32_t h, Fast<KIND> &g) {
    const float4 q0 = hyp[h];
    g.c[0] = q0.x, g.c[1] = q0.y, g.c[2] = q0.z, g.c[3] = q0.w;
    if (KIND == kCylinder) {
        const float4 q1 = hyp[(NH + 1) + h];
        g.c[4] = q1.x, g.c[5] = q1.y, g.c[6] = q1.z, g.c[7] = q1.w;
    }
    const float4 q2 = hyp[(KIND == kCylinder ? 2 : 1) * (NH + 1) + h];
    g.T = q2.x, g.band = q2.y;
}
template <int KIND, int NH>
__device__ __forceinline__ void load_fast_cull(const float4 *hyp, uint32_t h, Fast<KIND> &g, CullP &gk) {
    const float4 q0 = hyp[h];
    g.c[0] = q0.x, g.c[1] = q0.y, g.c[2] = q0.z, g.c[3] = q0.w;
    if (KIND == kCylinder) {
        const float4 q1 = hyp[(NH + 1) + h];
        g.c[4] = q1.x, g.c[5] = q1.y, g.c[6] = q1.z, g.c[7] = q1.w;
    }
    constexpr int PB = KIND == kCylinder ? 2 : 1;
    const float4 q2 = hyp[PB * (NH + 1) + h];
    g.T = q2.x, g.band = q2.y, gk.a = q2.z, gk.b = q2.w;
    gk.c = 0.f, gk.d = 0.f;
    if (KIND != kPlane) {
        const float4 q3 = hyp[(PB + 1) * (NH + 1) + h];
        gk.c = q3.x, gk.d = q3.y;
    }
}

/* PRE = the minimal models were solved beforehand (ScoreArgs::models_in; the host-buffer fit whose upload is still in
 * flight): a separate instantiation, so that the branch does not perturb the register allocation of the common one */
template <int KIND, int THREADS, int NH, bool STATS, bool PRE = false>
__global__ void __launch_bounds__(THREADS + 32, CellCfg<THREADS>::kMinBlocks) score_cell_kernel(const ScoreArgs a) {
    constexpr int HPT = (NH + THREADS - 1) / THREADS; /* hypotheses a consumer thread prepares in the prologue */
    constexpr int NC = KIND == kCylinder ? 8 : 4;
    constexpr int PB = KIND == kCylinder ? 2 : 1;
    constexpr int WARPS = THREADS / 32;
    constexpr int S = CellCfg<THREADS>::kStages;
    using L = CellSmem<KIND, THREADS, NH>;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float4 *tiles = reinterpret_cast<float4 *>(smem_raw);
    float4 *hyp = tiles + (size_t)S * kBlobF4;
    uint16_t *lists = reinterpret_cast<uint16_t *>(hyp + (size_t)L::kPlanes * (NH + 1));
    uint16_t *surv = reinterpret_cast<uint16_t *>(reinterpret_cast<unsigned char *>(lists) + L::kListBytes);
    uint2 *qbufs = reinterpret_cast<uint2 *>(reinterpret_cast<unsigned char *>(surv) + L::kSurvBytes);
    uint64_t *full = reinterpret_cast<uint64_t *>(qbufs + WARPS * kQBuf);
    uint64_t *empty = full + S;
    uint32_t *scnt = reinterpret_cast<uint32_t *>(empty + S);  /* inlier counts of the CTA's hypotheses      */
    volatile uint32_t *tile_id = scnt + NH;                    /* [S] tile in each stage                     */
    uint32_t *ccnt = const_cast<uint32_t *>(tile_id) + S;      /* [2][32] list lengths, by tile parity       */
    uint32_t *unit_next = ccnt + 2 * kTileCells;               /* [2] next unclaimed evaluation unit         */
    uint32_t *surv_cnt = unit_next + 2;                        /* [2] survivors of the tile test             */
    uint32_t *surv_next = surv_cnt + 2;                        /* [2] next unclaimed survivor                */

    const int tid = threadIdx.x;
    const uint32_t ntiles = (a.n + kTile - 1) / kTile;

    if (tid == THREADS) {
#pragma unroll
        for (int s = 0; s < S; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], 1);
        }
        mbar_fence_init();
    }
    if (tid < 2 * kTileCells + 6) ccnt[tid] = 0; /* list lengths and all cursors */
    __syncthreads();

    if (tid >= THREADS) { /* ---------------- producer: claims tiles, one bulk copy per tile */
        if (tid == THREADS) {
            for (uint32_t k = 0;; ++k) {
                const int st = k % S;
                if (k >= (uint32_t)S) mbar_wait_relaxed(&empty[st], ((k / S) - 1) & 1);
                const uint32_t t = atomicAdd(&a.tile_counter[blockIdx.x], 1u);
                if (t >= ntiles) {
                    tile_id[st] = kNoTile;
                    mbar_arrive(&full[st]);
                    break;
                }
                tile_id[st] = t;
                tma_load_1d(tiles + (size_t)st * kBlobF4, a.blob + (size_t)t * kBlobF4,
                            (uint32_t)(kBlobF4 * sizeof(float4)), &full[st]);
            }
        }
        return;
    }

    /* ---------------- consumers.  Prologue as in score_kernel: gather, MinimalFit (fp64), fp32 form */
    const CloudMeta M = *a.meta;
    const unsigned fullmask = 0xffffffffu;
    const int lane = tid & 31, warp = tid >> 5;
    const unsigned ltmask = (1u << lane) - 1;
    uint2 *qbuf = qbufs + warp * kQBuf;
    uint32_t qn = 0; /* entries staged in qbuf (warp-uniform) */
    unsigned long long st_tests = 0, st_tiles = 0, st_cells = 0, st_p1 = 0, st_p2 = 0, st_rescan = 0;
    /* staged guard-band pairs -> the global queue: one atomicAdd per flush */
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
                    bool ok;
                    if (PRE) {
                        const uint32_t sr = a.src_row(e.x);
                        ok = a.valid_in[sr] != 0;
                        for (int i = 0; i < 8; ++i) m[i] = a.models_in[(size_t)sr * 8 + i];
                    } else {
                        ok = fit_row<KIND>(a.xyz, a.nrm, a.samples, a.src_row(e.x), m, a.row_nrm);
                    }
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
        const uint32_t hl = h * THREADS + tid; /* CTA-local hypothesis index */
        const bool mine = hl < (uint32_t)NH;
        row[h] = mine ? blockIdx.x * NH + hl : 0xffffffffu;
        double m[8];
        bool ok = false;
        if (row[h] < a.rows) {
            if (PRE) {
                const uint32_t sr = a.src_row(row[h]);
                ok = a.valid_in[sr] != 0;
#pragma unroll
                for (int i = 0; i < 8; ++i) m[i] = a.models_in[(size_t)sr * 8 + i];
            } else {
                ok = fit_row<KIND>(a.xyz, a.nrm, a.samples, a.src_row(row[h]), m, a.row_nrm);
            }
        }
        invalid[h] = (row[h] < a.rows) && !ok;
        if (blockIdx.y == 0 && row[h] < a.rows) {
#pragma unroll
            for (int i = 0; i < 8; ++i)
                a.models[(size_t)row[h] * 8 + i] = (ok && i < param_count(KIND)) ? m[i] : 0.0;
        }
        if (!mine) continue;
        Fast<KIND> f;
        CullP ck;
        make_fast<KIND>(m, ok, M, a.thr, f);
        make_cull<KIND>(f, m, M, a.thr, ck);
        hyp[hl] = make_float4(f.c[0], f.c[1], f.c[2], f.c[3]);
        if (KIND == kCylinder) hyp[(NH + 1) + hl] = make_float4(f.c[4], f.c[5], f.c[6], f.c[7]);
        hyp[PB * (NH + 1) + hl] = make_float4(f.T, f.band, ck.a, ck.b);
        if (KIND != kPlane) hyp[(PB + 1) * (NH + 1) + hl] = make_float4(ck.c, ck.d, 0.f, 0.f);
        scnt[hl] = 0;
    }
    if (tid == 0) { /* the dummy hypothesis (index NH) pads incomplete passes / chunks: never an inlier, never in the
                     * band, culled by every sphere test (the encoding make_cull uses for a failed MinimalFit) */
        hyp[NH] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (KIND == kCylinder) hyp[(NH + 1) + NH] = make_float4(0.f, 0.f, 0.f, 0.f);
        hyp[PB * (NH + 1) + NH] = KIND == kPlane ? make_float4(-1.f, 0.f, -1.f, 0.f) : make_float4(-1.f, 0.f, 0.f, 0.f);
        if (KIND != kPlane) hyp[(PB + 1) * (NH + 1) + NH] = make_float4(-INFINITY, -1.f, 0.f, 0.f);
    }
    asm volatile("bar.sync 1, %0;" ::"n"(THREADS) : "memory");

    constexpr uint32_t kGroups = NH / 32;
    const uint32_t cta_row0 = blockIdx.x * NH; /* local index + cta_row0 = row of the launch */
    uint32_t nres = 0;

    /* the rare path: lane's hypothesis `hown` saw a point of the cell inside its guard band.  The whole warp
     * re-examines the cell for every flagged hypothesis (lane = point) and stages (row, sorted position,
     * provisional decision) for resolve_queue_kernel. */
    auto rescan = [&](unsigned need, uint32_t hown, const float4 *cell, uint32_t gbase) {
        const float4 p = cell[lane];
        while (need) {
            const int src = __ffs(need) - 1;
            need &= need - 1;
            const uint32_t h = __shfl_sync(fullmask, hown, src);
            Fast<KIND> g;
            load_fast<KIND, NH>(hyp, h, g);
            const float v = fast_v<KIND>(g, p);
            const bool amb = fabsf(v) < g.band; /* false for the NaN padding */
            const unsigned am = __ballot_sync(fullmask, amb);
            if (STATS) st_rescan++;
            if (am == 0) continue;
            const uint32_t na = __popc(am);
            if (qn + na > (uint32_t)kQBuf) flush_queue();
            if (amb) {
                const uint32_t prov = __float_as_uint(v) >> 31;
                qbuf[qn + __popc(am & ltmask)] = make_uint2(cta_row0 + h, (gbase + lane) | (prov << 31));
            }
            qn += na;
            nres += (lane == 0) ? na : 0u;
        }
    };

    /* step A of tile k: lane = hypothesis against the tile's bounding sphere; the survivors go to `surv`.
     * Groups are dealt to the warps statically: the step is short and balanced. */
    auto tile_tests = [&](uint32_t k) -> bool {
        const int st = k % S;
        mbar_wait(&full[st], (k / S) & 1);
        if (tile_id[st] == kNoTile) return false;
        const float4 tb = tiles[(size_t)st * kBlobF4 + kTile + kTileCells];
        for (uint32_t grp = warp; grp < kGroups; grp += WARPS) { /* kGroups / WARPS groups per warp (+-1) */
            Fast<KIND> g;
            CullP gk;
            load_fast_cull<KIND, NH>(hyp, grp * 32 + lane, g, gk);
            const bool alive = !cull_test<KIND>(g.c, gk, tb);
            const unsigned live = __ballot_sync(fullmask, alive);
            if (STATS) st_tests += 32, st_tiles += __popc(live);
            if (live) {
                uint32_t pos = 0;
                if (lane == 0) pos = atomicAdd(&surv_cnt[k & 1], (uint32_t)__popc(live));
                pos = __shfl_sync(fullmask, pos, 0);
                if (alive) surv[pos + __popc(live & ltmask)] = (uint16_t)(grp * 32 + lane);
            }
        }
        return true;
    };
    /* step B of tile k: lane = cell; every surviving hypothesis is tested against the 32 cell spheres and the
     * lane whose cell survives appends it to that cell's list.  Survivors are claimed four at a time. */
    auto cell_tests = [&](uint32_t k) {
        const float4 *sp = tiles + (size_t)(k % S) * kBlobF4;
        const float4 cb = sp[kTile + lane];
        uint32_t *cc = ccnt + (k & 1) * kTileCells;
        uint16_t *mylist = lists + lane * L::kListStride;
        const uint32_t total = surv_cnt[k & 1];
        if (warp == 0 && lane < 2) { /* the other parity's survivor cursors: idle until the next tile's step A */
            if (lane == 0) surv_cnt[(k + 1) & 1] = 0;
            else surv_next[(k + 1) & 1] = 0;
        }
        /* four survivors per trip, dealt to the warps statically: every trip costs the same 4 x 32 cell tests, and a
         * claim through a shared-memory cursor would put an atomic round trip in front of ~70 instructions of work */
        static_assert(NH % 4 == 0, "survivor quads are read with one 64-bit load");
        uint2 ids = make_uint2(0u, 0u);
        if (4u * warp < total) ids = *reinterpret_cast<const uint2 *>(surv + 4u * warp);
        for (uint32_t i0 = 4u * warp; i0 < total; i0 += 4u * WARPS) {
            const uint32_t e[4] = {ids.x & 0xffffu, ids.x >> 16, ids.y & 0xffffu, ids.y >> 16};
            if (i0 + 4u * WARPS < total) ids = *reinterpret_cast<const uint2 *>(surv + i0 + 4u * WARPS); /* next trip's */
#pragma unroll
            for (int half = 0; half < 2; ++half) { /* two hypotheses per trip for instruction-level parallelism */
                const uint32_t i = i0 + 2 * half;
                if (i >= total) break;
                const uint32_t h0 = e[2 * half];
                const uint32_t h1 = (i + 1 < total) ? e[2 * half + 1] : (uint32_t)NH;
                Fast<KIND> g0, g1;
                CullP k0, k1;
                load_fast_cull<KIND, NH>(hyp, h0, g0, k0);
                load_fast_cull<KIND, NH>(hyp, h1, g1, k1);
                const bool keep0 = !cull_test<KIND>(g0.c, k0, cb);
                const bool keep1 = !cull_test<KIND>(g1.c, k1, cb); /* the dummy is culled everywhere */
                if (keep0 || keep1) {
                    const uint32_t pos = atomicAdd(&cc[lane], (keep0 ? 1u : 0u) + (keep1 ? 1u : 0u));
                    if (keep0) mylist[pos] = (uint16_t)h0;
                    if (keep1) mylist[pos + (keep0 ? 1u : 0u)] = (uint16_t)h1;
                }
                if (STATS) st_cells += __popc(__ballot_sync(fullmask, keep0)) + __popc(__ballot_sync(fullmask, keep1));
            }
        }
    };
    /* phase 2 of tile k: lane = hypothesis.  A unit = up to 64 entries of one cell's list; the warp gathers their
     * coefficients once and streams the cell's 32 points as broadcast loads (the dense kernel's inner loop). */
    auto evaluate = [&](uint32_t k) {
        const float4 *sp = tiles + (size_t)(k % S) * kBlobF4;
        const uint32_t base = tile_id[k % S] * kTile;
        const uint32_t *cc = ccnt + (k & 1) * kTileCells;
        /* units are numbered cell by cell (inclusive scan of the cells' pass counts, in registers): no claim is empty */
        const uint32_t mycnt = cc[lane];
        const uint32_t mypass = (mycnt + 63) >> 6;
        uint32_t inc = mypass;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t tt = __shfl_up_sync(fullmask, inc, o);
            if (lane >= o) inc += tt;
        }
        const uint32_t units = __shfl_sync(fullmask, inc, 31);
        if (warp == 0) { /* reset the other parity's list lengths / unit cursor for the next tile (idle in this phase) */
            ccnt[((k + 1) & 1) * kTileCells + lane] = 0;
            if (lane == 0) unit_next[(k + 1) & 1] = 0;
        }
        /* units are claimed one ahead: the cursor's atomic round trip for the next unit is in flight while this one
         * is evaluated (every warp ends up claiming one unit past the end; the cursor is reset per tile) */
        uint32_t u = 0;
        if (lane == 0) u = atomicAdd(&unit_next[k & 1], 1u);
        u = __shfl_sync(fullmask, u, 0);
        while (u < units) {
            uint32_t u_next = 0;
            if (lane == 0) u_next = atomicAdd(&unit_next[k & 1], 1u);
            const uint32_t c = __popc(__ballot_sync(fullmask, inc <= u)); /* first cell whose units reach past u */
            const uint32_t off = (u - __shfl_sync(fullmask, inc - mypass, c)) * 64;
            const uint32_t rem = __shfl_sync(fullmask, mycnt, c) - off;
            const uint16_t *lst = lists + c * L::kListStride + off;
            const float4 *cell = sp + c * kCellPts;
            if (rem > 32) { /* two hypotheses per lane: packed fp32x2 arithmetic, bit-identical halves */
                const uint32_t h0 = lst[lane];
                const uint32_t h1 = (32u + lane < rem) ? (uint32_t)lst[32 + lane] : (uint32_t)NH;
                Fast<KIND> f0, f1;
                load_fast<KIND, NH>(hyp, h0, f0);
                load_fast<KIND, NH>(hyp, h1, f1);
                /* packed ONCE per pass.  The halves arrive in LDS.128 register quads, so ptxas has to move them
                 * into aligned pairs -- and, left alone, re-materialises those moves in front of every FFMA2 of the
                 * point loop (6 MOV per point).  Routing each pair through x * 1 + (-0) (exact for every x, -0
                 * included) makes the pair the result of a real instruction: it is built once and stays. */
                Fast2<KIND> f2;
                {
                    const f32x2_t one2 = pack2(1.f, 1.f), nz2 = pack2(-0.f, -0.f);
#pragma unroll
                    for (int i = 0; i < NC; ++i)
                        asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(f2.c[i]) : "l"(pack2(f0.c[i], f1.c[i])), "l"(one2), "l"(nz2));
                }
                uint32_t c0 = 0, c1 = 0;
                float m0 = INFINITY, m1 = INFINITY;
#pragma unroll
                for (int j = 0; j < kCellPts; ++j) {
                    float t0, t1;
                    fast_eval2<KIND>(f2, cell[j], t0, t1);
                    accumulate_v(__fsub_rn(fabsf(t0), f0.T), c0, m0);
                    accumulate_v(__fsub_rn(fabsf(t1), f1.T), c1, m1);
                }
                const bool fl0 = m0 < f0.band, fl1 = m1 < f1.band;
                if (__any_sync(fullmask, fl0 || fl1)) {
                    rescan(__ballot_sync(fullmask, fl0), h0, cell, base + c * kCellPts);
                    rescan(__ballot_sync(fullmask, fl1), h1, cell, base + c * kCellPts);
                }
                if (c0) atomicAdd(&scnt[h0], c0);
                if (c1) atomicAdd(&scnt[h1], c1); /* h1 == NH (dummy) never counts */
                if (STATS) st_p2++;
            } else {
                const uint32_t h0 = (lane < rem) ? (uint32_t)lst[lane] : (uint32_t)NH;
                Fast<KIND> f0;
                load_fast<KIND, NH>(hyp, h0, f0);
                uint32_t c0 = 0;
                float m0 = INFINITY;
#pragma unroll
                for (int j = 0; j < kCellPts; ++j) accumulate_v(fast_v<KIND>(f0, cell[j]), c0, m0);
                const bool fl0 = m0 < f0.band;
                if (__any_sync(fullmask, fl0)) rescan(__ballot_sync(fullmask, fl0), h0, cell, base + c * kCellPts);
                if (c0) atomicAdd(&scnt[h0], c0);
                if (STATS) st_p1++;
            }
            u = __shfl_sync(fullmask, u_next, 0);
        }
    };

    /* Software pipeline over the tiles, two consumer-only barriers per tile:
     *     [evaluate(k) ; tile_tests(k+1)]  X  [cell_tests(k+1)]  Y  [evaluate(k+1) ; tile_tests(k+2)]  X ...
     * X: every warp is done with tile k (its stage goes back to the producer) and the survivor list of tile k+1
     * is complete; Y: the cell lists of tile k+1 are complete.  The short, statically balanced tile_tests step
     * rides on the long evaluate phase instead of needing a barrier of its own. */
    bool more = tile_tests(0);
    asm volatile("bar.sync 1, %0;" ::"n"(THREADS) : "memory");
    if (more) {
        cell_tests(0);
        asm volatile("bar.sync 1, %0;" ::"n"(THREADS) : "memory");
        for (uint32_t k = 0;; ++k) {
            evaluate(k);
            more = tile_tests(k + 1);
            asm volatile("bar.sync 1, %0;" ::"n"(THREADS) : "memory"); /* X */
            if (tid == 0) mbar_arrive(&empty[k % S]);
            if (!more) break;
            cell_tests(k + 1);
            asm volatile("bar.sync 1, %0;" ::"n"(THREADS) : "memory"); /* Y */
        }
    }

    flush_queue();
    asm volatile("bar.sync 1, %0;" ::"n"(THREADS) : "memory"); /* all shared-memory counts are final */
#pragma unroll
    for (int h = 0; h < HPT; ++h) {
        if (row[h] < a.rows) {
            const uint32_t c = scnt[h * THREADS + tid]; /* row[h] < a.rows implies a CTA-local index < NH */
            const uint32_t ci = a.cnt_row(row[h]);
            if (c) atomicAdd(&a.counts[ci], c);
            if (invalid[h] && blockIdx.y == 0) atomicOr(&a.counts[ci], kInvalidBit);
        }
    }
    if (nres) atomicAdd(a.resolves, (unsigned long long)nres);
    if (STATS && lane == 0) {
        atomicAdd(&g_cell_stats[0], st_tests);
        atomicAdd(&g_cell_stats[1], st_tiles);
        atomicAdd(&g_cell_stats[2], st_cells);
        atomicAdd(&g_cell_stats[3], st_p1);
        atomicAdd(&g_cell_stats[4], st_p2);
        atomicAdd(&g_cell_stats[5], st_rescan);
    }
}

}  // namespace m3d
