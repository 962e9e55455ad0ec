/*
 * loop_kernels.cuh -- the bookkeeping of the RANSAC loop on the device.
 *
 *   sample draw    mt_stream_kernel / dup_positions_kernel / row_breaks_kernel / build_rows_kernel:
 *                  the reference's RandomSampler (include/misc3d/utils.h:74-97: std::mt19937,
 *                  `rng_() % size_`, duplicates inside a row rejected) reproduced bit for bit on the
 *                  GPU, so the sample table of a wave never exists on the host;
 *   arg-best       wave_best_kernel / best_merge_kernel: the outcome of the sequential best-update of
 *                  ransac.h:592-613 for a wave without early exit (probability == 1): the largest
 *                  inlier count, the first row that reaches it, the rows tied with it (the host breaks
 *                  ties with inlier_rmse exactly as before -- rare), the number of successful
 *                  MinimalFits.  With R ranks one 64-byte record per rank is exchanged instead of all
 *                  counts.
 */
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

#include "scan.h"
#include "mt_jump_table.h"

namespace m3d {

/* ------------------------------------------------------------------------------ sample stream */
struct MtInit { /* std::mt19937 state after seeding (624 words, passed as a kernel parameter) */
    uint32_t mt[624];
};

__device__ __forceinline__ uint32_t mt_twist(uint32_t hi, uint32_t lo, uint32_t far_) {
    const uint32_t y = (hi & 0x80000000u) | (lo & 0x7fffffffu);
    return far_ ^ (y >> 1) ^ ((0u - (y & 1u)) & 0x9908b0dfu);
}
__device__ __forceinline__ uint32_t mt_temper(uint32_t y) {
    y ^= y >> 11;
    y ^= (y << 7) & 0x9d2c5680u;
    y ^= (y << 15) & 0xefc60000u;
    y ^= y >> 18;
    return y;
}
/* x % size for 32-bit x (Lemire's exact fastmod, magic = ceil(2^64 / size)) */
__device__ __forceinline__ uint32_t mt_reduce(uint32_t x, uint32_t size, uint64_t magic) {
    return size <= 1 ? 0u : (uint32_t)__umul64hi(magic * (uint64_t)x, (uint64_t)size);
}

/* One CTA, ONE barrier per block of 624 words.  The recurrence x[n+624] = f(x[n], x[n+1], x[n+397]) is sequential
 * over blocks but 227-wide inside one: thread t owns elements t, 227+t and 454+t of every block.  The "far" operand
 * of element 227+t is element t of the same block and that of 454+t is element 227+t -- both produced by the SAME
 * thread one step earlier, so they stay in registers; only the neighbour operand x[n+1] (and the far operand of the
 * first 227 elements) comes from other threads, out of the previous block, through a double-buffered copy of the state
 * in shared memory.  The one element that needs a value of its own block from another thread -- 623 = f(.., new[0],
 * new[396]) -- is deferred to the next iteration, where every thread recomputes it from three broadcast loads.
 * out[b*624 + i] = the RAW (b*624+i)-th state word; tempering and `% size` are left to stream_finish_kernel (grid-wide):
 * nothing but the recurrence sits on the sequential path.  Measured 0.31 us per block (0.12 ms for the 392 blocks of the
 * 80k-row table of an 8-GPU fit; the first version, three barriers per block with tempering and the 64-bit fastmod
 * inline, took 0.45 us).  A variant that kept the global stores away from the barrier (shared-memory ring + copy
 * warps) measured the same 0.31 us: the loop is bound by its dependent LDS -> twist x 3 -> STS -> barrier chain. */
struct MtSmem {
    uint32_t st[2][624];
    uint32_t s623[2];
};
/* the block loop: sm.st[0] and (a, b, c) hold the 624 words in front of the first block to produce */
__device__ __forceinline__ void mt_run(MtSmem &sm, uint32_t a, uint32_t b, uint32_t c, uint32_t nblocks,
                                       uint32_t *__restrict__ out) {
    const int t = threadIdx.x;
    const bool live = t < 227; /* threads 227..255 only keep the barriers whole */
    for (uint32_t blk = 0; blk < nblocks; ++blk) {
        const int par = blk & 1;
        if (live) {
            const uint32_t *s = sm.st[par];
            uint32_t *o = out + (size_t)blk * 624;
            /* the common case first: nothing here waits for element 623 of the previous block */
            uint32_t nA = mt_twist(a, s[t + 1], s[t + 397]); /* t == 226 reads the stale s[623]: redone below */
            uint32_t e623 = 0;
            const bool special = (t >> 5) == 5 || (t >> 5) == 7; /* the warps of threads 168, 169 and 226: the users of
                                                                  * element 623 of the previous block (whole warps, so
                                                                  * that the branch does not diverge) */
            if (special) {
                e623 = s[623];
                if (blk != 0) { /* (for blk == 0 the given state is complete) */
                    e623 = mt_twist(sm.s623[par], s[0], s[396]);
                    if (t == 169) out[(size_t)(blk - 1) * 624 + 623] = e623;
                }
                if (t == 169) sm.s623[par ^ 1] = e623; /* "old[623]" of the next iteration */
                if (t == 226) nA = mt_twist(a, s[t + 1], e623);
            }
            const uint32_t nB = mt_twist(b, s[228 + t], nA);
            uint32_t nC = 0;
            if (t < 169) nC = mt_twist(c, (special && t == 168) ? e623 : s[455 + t], nB);
            uint32_t *w = sm.st[par ^ 1];
            w[t] = nA, w[227 + t] = nB;
            o[t] = nA, o[227 + t] = nB;
            if (t < 169) w[454 + t] = nC, o[454 + t] = nC;
            a = nA, b = nB, c = nC;
        }
        __syncthreads();
    }
    if (nblocks && t == 169) { /* the deferred last element */
        const int par = nblocks & 1;
        out[(size_t)(nblocks - 1) * 624 + 623] = mt_twist(sm.s623[par], sm.st[par][0], sm.st[par][396]);
    }
}

/* blocks [0, nblocks) of the stream from the seed state.  `zero_windows` > 1: the windows (first blocks) of segments
 * 1 .. zero_windows-1 are cleared for mt_jump_kernel's XOR accumulation. */
__global__ void __launch_bounds__(256) mt_stream_kernel(const MtInit init, uint32_t nblocks, uint32_t *__restrict__ out,
                                                        uint32_t zero_windows, uint32_t seg_blocks) {
    __shared__ MtSmem sm;
    const int t = threadIdx.x;
    for (uint32_t p = 1; p < zero_windows; ++p)
        for (int i = t; i < 624; i += 256) out[(size_t)p * seg_blocks * 624 + i] = 0u;
    uint32_t a = 0, b = 0, c = 0;
    if (t < 227) {
        a = init.mt[t], b = init.mt[227 + t], c = t < 170 ? init.mt[454 + t] : 0u;
        sm.st[0][t] = a;
        sm.st[0][227 + t] = b;
        if (t < 170) sm.st[0][454 + t] = c;
    }
    __syncthreads();
    mt_run(sm, a, b, c, nblocks, out);
}

/* ---- jump-ahead.  The raw words obey one linear recurrence over GF(2) (characteristic polynomial of degree 19937), so
 * the 624-word window at stream position J is  y[J + j] = XOR_{i : g_i = 1} y[i + j]  with g = x^J mod phi -- a binary
 * convolution of the first 19937 + 623 words (kMtPrefixBlocks blocks) with a precomputed polynomial
 * (mt_jump_table.h, tools/gen_mt_jump.py).  Segment p of the stream starts at block p * kMtSegBlocks; with its window
 * known, one CTA per segment walks the recurrence, all segments at once.
 * grid (kMtJumpChunks, segments - 1), 640 threads: CTA (c, p-1) adds the terms i in [416 c, 416 c + 416) to window p. */
constexpr int kMtPrefixBlocks = 33;
constexpr int kMtJumpChunks = 48;
constexpr int kMtJumpBits = 416; /* 48 x 416 = 19968 = 624 x 32 */
__global__ void __launch_bounds__(640) mt_jump_kernel(uint32_t *__restrict__ stream, const uint32_t *__restrict__ polys,
                                                      uint32_t seg_blocks) {
    __shared__ uint32_t ys[kMtJumpBits + 624];
    __shared__ uint16_t terms[kMtJumpBits];
    __shared__ uint32_t n_terms;
    const int t = threadIdx.x, c = blockIdx.x, p = blockIdx.y + 1;
    if (t == 0) n_terms = 0;
    for (int i = t; i < kMtJumpBits + 623; i += 640) ys[i] = stream[c * kMtJumpBits + i];
    __syncthreads();
    if (t < kMtJumpBits) {
        const int bit = c * kMtJumpBits + t;
        if ((polys[(size_t)(p - 1) * 624 + (bit >> 5)] >> (bit & 31)) & 1u) terms[atomicAdd(&n_terms, 1u)] = (uint16_t)t;
    }
    __syncthreads();
    if (t < 624) {
        const uint32_t cnt = n_terms;
        uint32_t acc = 0;
        for (uint32_t i = 0; i < cnt; ++i) acc ^= ys[terms[i] + t];
        if (cnt) atomicXor(stream + (size_t)p * seg_blocks * 624 + t, acc);
    }
}

/* one CTA per segment: the window of segment p (block p * seg_blocks; for p == 0 the last prefix block) is the state,
 * the CTA produces the blocks up to the next segment's window (the last segment: up to total_blocks) */
__global__ void __launch_bounds__(256) mt_segments_kernel(uint32_t *__restrict__ stream, uint32_t seg_blocks,
                                                          uint32_t nseg, uint32_t total_blocks) {
    __shared__ MtSmem sm;
    const int t = threadIdx.x;
    const uint32_t p = blockIdx.x;
    const uint32_t sb = p == 0 ? (uint32_t)kMtPrefixBlocks - 1 : p * seg_blocks;
    const uint32_t end = (p + 1 == nseg) ? total_blocks : (p + 1) * seg_blocks;
    const uint32_t *state = stream + (size_t)sb * 624;
    uint32_t a = 0, b = 0, c = 0;
    if (t < 227) {
        a = state[t], b = state[227 + t], c = t < 170 ? state[454 + t] : 0u;
        sm.st[0][t] = a;
        sm.st[0][227 + t] = b;
        if (t < 170) sm.st[0][454 + t] = c;
    }
    __syncthreads();
    if (end > sb + 1) mt_run(sm, a, b, c, end - sb - 1, stream + (size_t)(sb + 1) * 624);
}

/* raw state words -> `rng() % size`: tempering and the exact modulo, grid-wide (kept off the sequential kernel) */
__global__ void __launch_bounds__(256) stream_finish_kernel(uint32_t *__restrict__ stream, uint32_t len, uint32_t size,
                                                            uint64_t magic) {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < len; i += gridDim.x * blockDim.x)
        stream[i] = mt_reduce(mt_temper(stream[i]), size, magic);
}

/* A row that starts at stream position s takes exactly k draws unless two of them coincide
 * (utils.h:88-94 rejects an index already in the row).  Such positions are rare (~k^2/2n of them):
 * they are listed here, everything else is regular. */
struct RowBreaks {
    uint32_t n_dup;     /* positions listed by dup_positions_kernel (may exceed the capacity)  */
    uint32_t n_breaks;  /* entries of `brk`                                                    */
    uint32_t status;    /* 0 ok, 1 = too many duplicates / stream too short: draw on the host  */
    uint32_t pad;
};
constexpr uint32_t kDupCap = 2048;

__global__ void __launch_bounds__(256) dup_positions_kernel(const uint32_t *__restrict__ stream, uint32_t len, int k,
                                                            RowBreaks *__restrict__ rb, uint32_t *__restrict__ list) {
    for (uint32_t s = blockIdx.x * blockDim.x + threadIdx.x; s + k <= len; s += gridDim.x * blockDim.x) {
        uint32_t v[4];
        bool dup = false;
        for (int i = 0; i < k; ++i) {
            v[i] = stream[s + i];
            for (int j = 0; j < i; ++j) dup = dup || (v[i] == v[j]);
        }
        if (dup) {
            const uint32_t pos = atomicAdd(&rb->n_dup, 1u);
            if (pos < kDupCap) list[pos] = s;
        }
    }
}

/* One CTA: sorts the listed positions and walks them once.  brk[j] = (first row, its stream position)
 * of the j-th run of regular rows; row r of a run starts at position + (r - first row) * k. */
__global__ void __launch_bounds__(1024) row_breaks_kernel(const uint32_t *__restrict__ stream, uint32_t len, int k,
                                                          uint32_t rows, RowBreaks *__restrict__ rb,
                                                          const uint32_t *__restrict__ list, uint2 *__restrict__ brk) {
    __shared__ uint32_t sl[kDupCap];
    const uint32_t nd = rb->n_dup;
    if (nd > kDupCap) {
        if (threadIdx.x == 0) rb->status = 1;
        return;
    }
    uint32_t cap = 2; /* sort the next power of two >= nd entries (typically a handful) */
    while (cap < nd) cap <<= 1;
    for (uint32_t i = threadIdx.x; i < cap; i += blockDim.x) sl[i] = i < nd ? list[i] : 0xffffffffu;
    __syncthreads();
    for (uint32_t size = 2; size <= cap; size <<= 1) /* bitonic sort, ascending */
        for (uint32_t stride = size >> 1; stride; stride >>= 1) {
            for (uint32_t i = threadIdx.x; i < cap / 2; i += blockDim.x) {
                const uint32_t lo = 2 * i - (i & (stride - 1)), hi = lo + stride;
                const bool up = (lo & size) == 0;
                const uint32_t a = sl[lo], b = sl[hi];
                if ((a > b) == up) {
                    sl[lo] = b;
                    sl[hi] = a;
                }
            }
            __syncthreads();
        }
    if (threadIdx.x != 0) return;
    uint32_t s = 0, r = 0, nb = 0, status = 0;
    brk[nb++] = make_uint2(0u, 0u);
    for (uint32_t j = 0; j < nd; ++j) {
        const uint32_t e = sl[j];
        if (e < s || (e - s) % (uint32_t)k) continue; /* no row starts there */
        const uint32_t re = r + (e - s) / (uint32_t)k;
        if (re >= rows) break;
        /* the row at e: the reference's rejection loop */
        uint32_t v[4], have = 0, p = e;
        while (have < (uint32_t)k && p < len) {
            const uint32_t x = stream[p++];
            bool dup = false;
            for (uint32_t q = 0; q < have; ++q) dup = dup || (v[q] == x);
            if (!dup) v[have++] = x;
        }
        if (have < (uint32_t)k) {
            status = 1;
            break;
        }
        s = p;
        r = re + 1;
        brk[nb++] = make_uint2(r, s);
    }
    if ((uint64_t)s + (uint64_t)(rows - min(r, rows)) * (uint32_t)k > (uint64_t)len) status = 1;
    rb->n_breaks = nb;
    rb->status = status;
}

__global__ void __launch_bounds__(256) build_rows_kernel(const uint32_t *__restrict__ stream, uint32_t len, int k,
                                                         uint32_t rows, const RowBreaks *__restrict__ rb,
                                                         const uint2 *__restrict__ brk, uint32_t *__restrict__ table) {
    if (rb->status) return;
    const uint32_t nb = rb->n_breaks;
    for (uint32_t r = blockIdx.x * blockDim.x + threadIdx.x; r < rows; r += gridDim.x * blockDim.x) {
        uint32_t lo = 0, hi = nb; /* last run whose first row is <= r */
        while (hi - lo > 1) {
            const uint32_t mid = (lo + hi) >> 1;
            if (brk[mid].x <= r)
                lo = mid;
            else
                hi = mid;
        }
        const uint2 b = brk[lo];
        uint32_t p = b.y + (r - b.x) * (uint32_t)k;
        uint32_t v[4], have = 0;
        while (have < (uint32_t)k && p < len) {
            const uint32_t x = stream[p++];
            bool dup = false;
            for (uint32_t q = 0; q < have; ++q) dup = dup || (v[q] == x);
            if (!dup) v[have++] = x;
        }
        for (int i = 0; i < k; ++i) table[(size_t)r * k + i] = v[i];
    }
}

/* ---------------------------------------------------------------------------------- arg-best */
constexpr uint32_t kNoRow = 0xffffffffu;
constexpr int kTiedCap = 8;
struct BestRec { /* 64 bytes: what one rank knows about its shard of a wave */
    uint32_t max_count;  /* largest inlier count over rows whose MinimalFit succeeded (0: none)        */
    uint32_t first_row;  /* smallest wave row that reaches it (kNoRow: none)                           */
    uint32_t n_tied;     /* rows that reach it                                                         */
    uint32_t n_valid;    /* successful MinimalFits (`count++` of ransac.h:583-586)                     */
    uint32_t first_full; /* smallest wave row with count == n points (fitness 1 stops the loop,
                            ransac.h:607-609), else kNoRow                                             */
    uint32_t status;     /* != 0: the device-side sample draw gave up (host draw needed)               */
    uint32_t pad[2];
    uint32_t tied[kTiedCap]; /* the first kTiedCap tied rows, ascending                                */
};
static_assert(sizeof(BestRec) == 64, "BestRec is exchanged as 64 bytes");

/* counts: the rank's shard-local rows [0, mine) of the wave; bit 31 = MinimalFit failed */
__global__ void __launch_bounds__(1024) wave_best_kernel(const uint32_t *__restrict__ counts, uint32_t mine,
                                                         uint32_t world, uint32_t rank, uint32_t n_points,
                                                         const RowBreaks *__restrict__ draw_status,
                                                         BestRec *__restrict__ out) {
    __shared__ unsigned long long skey[32];
    __shared__ uint32_t sfull[32], svalid[32];
    __shared__ unsigned long long best_key;
    __shared__ uint32_t n_tied, tied[kTiedCap];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    unsigned long long key = 0; /* (count << 32) | ~row: the maximum is the largest count at the smallest row */
    uint32_t full = kNoRow, nvalid = 0;
    for (uint32_t l = threadIdx.x; l < mine; l += blockDim.x) {
        const uint32_t raw = counts[l];
        if (raw & 0x80000000u) continue;
        ++nvalid;
        if (raw == 0) continue;
        const uint32_t g = ShardMap::wave_row_of(l, world, rank);
        const unsigned long long kk = ((unsigned long long)raw << 32) | (unsigned long long)(~g);
        key = kk > key ? kk : key;
        if (raw == n_points) full = min(full, g);
    }
    for (int o = 16; o; o >>= 1) {
        const unsigned long long k2 = __shfl_xor_sync(0xffffffffu, key, o);
        key = k2 > key ? k2 : key;
        full = min(full, __shfl_xor_sync(0xffffffffu, full, o));
        nvalid += __shfl_xor_sync(0xffffffffu, nvalid, o);
    }
    if (lane == 0) {
        skey[w] = key;
        sfull[w] = full;
        svalid[w] = nvalid;
    }
    if (threadIdx.x == 0) n_tied = 0;
    __syncthreads();
    if (w == 0) {
        key = skey[lane];
        full = sfull[lane];
        nvalid = svalid[lane];
        for (int o = 16; o; o >>= 1) {
            const unsigned long long k2 = __shfl_xor_sync(0xffffffffu, key, o);
            key = k2 > key ? k2 : key;
            full = min(full, __shfl_xor_sync(0xffffffffu, full, o));
            nvalid += __shfl_xor_sync(0xffffffffu, nvalid, o);
        }
        if (lane == 0) {
            best_key = key;
            sfull[0] = full;
            svalid[0] = nvalid;
        }
    }
    __syncthreads();
    const uint32_t best = (uint32_t)(best_key >> 32);
    if (best) {
        for (uint32_t l = threadIdx.x; l < mine; l += blockDim.x)
            if (counts[l] == best) { /* bit 31 clear by construction */
                const uint32_t slot = atomicAdd(&n_tied, 1u);
                if (slot < (uint32_t)kTiedCap) tied[slot] = ShardMap::wave_row_of(l, world, rank);
            }
    }
    __syncthreads();
    if (threadIdx.x != 0) return;
    BestRec r;
    r.max_count = best;
    r.first_row = best ? ~(uint32_t)(best_key & 0xffffffffu) : kNoRow;
    r.n_tied = n_tied;
    r.n_valid = svalid[0];
    r.first_full = sfull[0];
    r.status = draw_status ? draw_status->status : 0u;
    r.pad[0] = r.pad[1] = 0;
    const uint32_t nt = min(n_tied, (uint32_t)kTiedCap);
    for (uint32_t i = 1; i < nt; ++i) { /* ascending */
        const uint32_t v = tied[i];
        uint32_t j = i;
        for (; j > 0 && tied[j - 1] > v; --j) tied[j] = tied[j - 1];
        tied[j] = v;
    }
    for (int i = 0; i < kTiedCap; ++i) r.tied[i] = (uint32_t)i < nt ? tied[i] : kNoRow;
    *out = r;
}

/* One warp: merges the records of all ranks (identical on every rank) and stages the sample row (and,
 * in host-normals mode, the normals) of the provisional winner for the refit kernels.
 * n_tied > kTiedCap in the result = "more ties than listed": the host replays the wave's counts. */
__global__ void best_merge_kernel(const BestRec *__restrict__ recs, uint32_t world,
                                  const uint32_t *__restrict__ table, int k,
                                  const double *__restrict__ row_nrm, BestRec *__restrict__ merged,
                                  uint32_t *__restrict__ sample_out, double *__restrict__ nrm_out) {
    if (threadIdx.x != 0) return;
    BestRec m;
    m.max_count = 0, m.first_row = kNoRow, m.n_tied = 0, m.n_valid = 0, m.first_full = kNoRow, m.status = 0;
    m.pad[0] = m.pad[1] = 0;
    for (int i = 0; i < kTiedCap; ++i) m.tied[i] = kNoRow;
    for (uint32_t r = 0; r < world; ++r) m.max_count = max(m.max_count, recs[r].max_count);
    bool overflow = false;
    for (uint32_t r = 0; r < world; ++r) {
        const BestRec &q = recs[r];
        m.n_valid += q.n_valid;
        m.first_full = min(m.first_full, q.first_full);
        m.status |= q.status;
        if (m.max_count == 0 || q.max_count != m.max_count) continue;
        m.first_row = min(m.first_row, q.first_row);
        if (q.n_tied > (uint32_t)kTiedCap) overflow = true;
        m.n_tied += q.n_tied;
    }
    /* the kTiedCap smallest tied rows over all ranks (every rank's list is ascending, rows are distinct) */
    long long last = -1;
    for (int slot = 0; slot < kTiedCap && m.max_count; ++slot) {
        uint32_t best = kNoRow;
        for (uint32_t r = 0; r < world; ++r) {
            if (recs[r].max_count != m.max_count) continue;
            for (int i = 0; i < kTiedCap; ++i) {
                const uint32_t v = recs[r].tied[i];
                if (v != kNoRow && (long long)v > last && v < best) best = v;
            }
        }
        if (best == kNoRow) break;
        m.tied[slot] = best;
        last = best;
    }
    if (overflow && m.n_tied <= (uint32_t)kTiedCap) m.n_tied = kTiedCap + 1;
    *merged = m;
    const uint32_t row = m.first_row == kNoRow ? 0u : m.first_row;
    for (int j = 0; j < k; ++j) sample_out[j] = table[(size_t)row * k + j];
    if (row_nrm)
        for (int j = 0; j < 3 * k; ++j) nrm_out[j] = row_nrm[(size_t)row * k * 3 + j];
}

}  // namespace m3d
