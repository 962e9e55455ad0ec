/*
 * scan.h -- host side of the RANSAC loop: the reference's sample stream and the sequential
 * semantics of FitModelParallel replayed over batched GPU results ("ordered scan").
 */
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <limits>
#include <random>

#include "../../include/m3d_capi.h"

namespace m3d {

/* RandomSampler<size_t> (utils.h:72-97) with an injected seed: std::mt19937, idx = rng() % size,
 * duplicates rejected, k accepted draws per call, draw order kept (SelectByIndex re-orders later). */
/* Hypothesis sharding over `world` ranks: a wave of `rows` hypotheses is cut into blocks of
 * kShardBlock rows dealt out cyclically, block b to rank b % world.  Block-cyclic rather than one contiguous
 * slice per rank because the sample rows come from ONE sequential mt19937 stream that every rank has to
 * replay from the start: with cyclic blocks every rank owns rows near the beginning of the wave and can
 * launch its first part after drawing a fraction of the table, drawing the rest while its GPU works.
 * Shard-local rows are numbered in block order; `padded` (identical on all ranks) is the all-gather stride. */
constexpr uint32_t kShardBlock = 256;
struct ShardMap {
    uint32_t rows, world, rank;
#ifdef __CUDACC__
    __host__ __device__
#endif
    static inline uint32_t wave_row_of(uint32_t l, uint32_t world, uint32_t rank) {
        return world <= 1 ? l : ((l / kShardBlock) * world + rank) * kShardBlock + l % kShardBlock;
    }
    uint32_t blocks() const { return (rows + kShardBlock - 1) / kShardBlock; }
    uint32_t padded() const { return world <= 1 ? rows : ((blocks() + world - 1) / world) * kShardBlock; }
    uint32_t local_rows_of(uint32_t r) const { /* rows of the wave that rank r owns */
        if (world <= 1) return rows;
        uint32_t n = 0;
        for (uint32_t b = r; b < blocks(); b += world) n += std::min<uint32_t>(kShardBlock, rows - b * kShardBlock);
        return n;
    }
    uint32_t local_rows() const { return local_rows_of(rank); }
    uint32_t wave_row(uint32_t l) const { return wave_row_of(l, world, rank); }
    uint32_t rank_of(uint32_t g) const { return world <= 1 ? 0 : (g / kShardBlock) % world; }
    uint32_t local_of(uint32_t g) const {
        return world <= 1 ? g : (g / (kShardBlock * world)) * kShardBlock + g % kShardBlock;
    }
    /* position of wave row g in the rank-major all-gathered buffer */
    size_t gathered_index(uint32_t g) const { return (size_t)rank_of(g) * padded() + local_of(g); }
};

/* The sample stream of the reference (utils.h:81-97): idx = std::mt19937(seed)() % size with size_t
 * arithmetic, duplicates within a row rejected.  The table of a wave is drawn on the host while the GPU
 * waits for (part of) it, and with R ranks it is R times longer, so the generator is on the critical path:
 * the raw stream is produced 624 numbers at a time and reduced modulo `size` in the same pass (Lemire's
 * exact fastmod), with an AVX2 body selected at run time (sampler.cpp); the values are those of
 * std::mt19937 + `%` bit for bit (tests/test_host.py compares with the oracle and the compiled reference). */
struct SampleStream {
    uint32_t mt[624];
    uint32_t idx[624]; /* tempered outputs of the current block, already reduced modulo size */
    int pos = 624;
    uint32_t size;
    uint64_t magic; /* ceil(2^64 / size): x % size = ((magic * x mod 2^64) * size) >> 64 for 32-bit x */
    SampleStream(uint32_t seed, size_t n);
    void refill(); /* sampler.cpp */
    inline uint32_t next() {
        if (pos == 624) refill();
        return idx[pos++];
    }
    /* `rows` consecutive rows of k distinct indices each (draw order), out[rows][k] (sampler.cpp) */
    void draw_rows(int k, size_t rows, uint32_t *out);
    /* one row: k distinct indices in draw order */
    inline void draw(int k, uint32_t *out) {
        int have = 0;
        while (have < k) {
            const uint32_t v = next();
            bool dup = false;
            for (int j = 0; j < have; ++j) dup = dup || (out[j] == v);
            if (!dup) out[have++] = v;
        }
    }
};

/* ransac.h:601-610: size_t current_iteration = min(log(1-p)/log(1-fitness^k), max_it); the
 * implicit double -> size_t conversion of -inf / huge values is what x86-64 gcc produces
 * (2^63 = "no limit"; NaN -> 0), SURVEY Appendix A.3. */
inline size_t adaptive_limit(double fitness, int k, double prob, uint64_t max_it) {
    if (!(fitness < 1.0)) return 0; /* ransac.h:607-609 */
    const double v = std::min(std::log(1 - prob) / std::log(1 - std::pow(fitness, k)), (double)max_it);
    if (v != v) return 0;
    if (v < 0 || v >= 18446744073709551616.0) return (size_t)1 << 63;
    return (size_t)v;
}

/* Sequential replay of ransac.h:572-613 for i = 0,1,2,...  step() must be called in loop order.
 * get_rmse(j, exact, &rmse) is only invoked to break inlier-count ties (ransac.h:595-596):
 * exact = false asks for error/sqrt(count) from a parallel sum, exact = true for the reference's
 * index-order sum. */
struct OrderedScan {
    size_t n_points;
    int k;
    double prob;
    uint64_t max_it;
    double best_fit = 0, best_rmse = 0; /* ransac.h:459-460, 519-522 */
    bool best_rmse_known = true, best_rmse_exact = true;
    bool found = false, stopped = false;
    uint64_t best_index = 0, best_count = 0, stop_index = 0;
    size_t count = 0; /* successful MinimalFits so far  */
    size_t cur = std::numeric_limits<size_t>::max();
    int error = 0;

    OrderedScan(size_t n, int k_, double p, uint64_t mi) : n_points(n), k(k_), prob(p), max_it(mi) {
        stop_index = mi;
    }

    template <class F>
    void step(uint64_t i, bool valid, uint64_t cnt, F &&get_rmse) {
        if (stopped) return;
        if (count > cur) { /* ransac.h:573-575: every later iteration is skipped too */
            stopped = true;
            stop_index = i;
            return;
        }
        if (!valid) return; /* MinimalFit false: no count++ (ransac.h:583-586) */
        if (cnt < best_count) { /* fitness = cnt / n is strictly monotone in cnt: cannot be better (the common case) */
            count++;
            return;
        }
        bool better = false;
        if (cnt != 0) {
            const double fitness = (double)cnt / (double)n_points;
            if (fitness > best_fit) {
                better = true;
                best_rmse_known = false;
            } else if (fitness == best_fit) {
                double mine = 0, theirs = best_rmse;
                bool mine_exact = false;
                if (!best_rmse_known) {
                    error |= get_rmse(best_index, false, &theirs);
                    best_rmse_exact = false;
                }
                error |= get_rmse(i, false, &mine);
                const double tol = 1e-9 * std::max(std::fabs(mine), std::fabs(theirs));
                if (std::fabs(mine - theirs) <= tol) { /* too close for a parallel sum: go exact */
                    if (!best_rmse_exact) error |= get_rmse(best_index, true, &theirs);
                    error |= get_rmse(i, true, &mine);
                    best_rmse_exact = true;
                    mine_exact = true;
                }
                best_rmse = theirs;
                best_rmse_known = true;
                if (mine < theirs) {
                    better = true;
                    best_rmse = mine;
                    best_rmse_exact = mine_exact;
                }
            }
            if (better) {
                best_fit = fitness;
                best_index = i;
                best_count = cnt;
                found = true;
                cur = adaptive_limit(best_fit, k, prob, max_it);
            }
        }
        /* cnt == 0: fitness 0 / rmse 1e10 never beats the initial (0, 0) nor any found model */
        count++;
    }
    void fill(m3d_ransac_stats *st) const {
        st->best_index = best_index;
        st->best_count = best_count;
        st->iterations_run = count;
        st->stop_index = stopped ? stop_index : max_it;
        st->found = found ? 1 : 0;
        if (best_rmse_known && found) st->best_rmse = best_rmse;
    }
};

}  // namespace m3d
