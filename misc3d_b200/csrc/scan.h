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
/* std::mt19937's output stream, produced 624 numbers at a time (vectorisable twist + tempering loops)
 * instead of one call at a time: the table of a 10k-hypothesis wave is drawn on the host while the GPU
 * waits for it, so the generator's speed is on the critical path of a fit. */
struct Mt19937Bulk {
    uint32_t mt[624];
    uint32_t out[624];
    int pos = 624;
    explicit Mt19937Bulk(uint32_t seed) {
        mt[0] = seed;
        for (uint32_t i = 1; i < 624; ++i) mt[i] = 1812433253u * (mt[i - 1] ^ (mt[i - 1] >> 30)) + i;
    }
    static inline uint32_t twist(uint32_t hi, uint32_t lo, uint32_t far) {
        const uint32_t y = (hi & 0x80000000u) | (lo & 0x7fffffffu);
        return far ^ (y >> 1) ^ ((0u - (y & 1u)) & 0x9908b0dfu);
    }
    void refill() {
        for (int i = 0; i < 227; ++i) mt[i] = twist(mt[i], mt[i + 1], mt[i + 397]);
        for (int i = 227; i < 623; ++i) mt[i] = twist(mt[i], mt[i + 1], mt[i - 227]);
        mt[623] = twist(mt[623], mt[0], mt[396]);
        for (int i = 0; i < 624; ++i) {
            uint32_t y = mt[i];
            y ^= y >> 11;
            y ^= (y << 7) & 0x9d2c5680u;
            y ^= (y << 15) & 0xefc60000u;
            y ^= y >> 18;
            out[i] = y;
        }
        pos = 0;
    }
    inline uint32_t next() {
        if (pos == 624) refill();
        return out[pos++];
    }
};

struct SampleStream {
    Mt19937Bulk rng;
    uint32_t size;
    uint64_t magic; /* exact x % size for 32-bit x by two multiplications (Lemire's fastmod) */
    SampleStream(uint32_t seed, size_t n) : rng(seed), size((uint32_t)n), magic(n ? UINT64_MAX / (uint32_t)n + 1 : 0) {}
    inline uint32_t mod(uint32_t x) const {
        if (size == 1) return 0;
        const uint64_t low = magic * x;
        return (uint32_t)(((unsigned __int128)low * size) >> 64);
    }
    /* utils.h:81-97: idx = rng() % size (size_t arithmetic on a 32-bit draw), reject duplicates */
    void draw(int k, uint32_t *out) {
        int have = 0;
        while (have < k) {
            const uint32_t idx = mod(rng.next());
            bool dup = false;
            for (int j = 0; j < have; ++j) dup = dup || (out[j] == idx);
            if (!dup) out[have++] = idx;
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
