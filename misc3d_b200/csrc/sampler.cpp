/*
 * sampler.cpp -- block generator behind SampleStream (scan.h): std::mt19937's twist + tempering over 624
 * words at a time, fused with the exact reduction modulo the cloud size.  Plain C++ (built by the host
 * compiler, not nvcc) so that the AVX2 body can be selected at run time with a function-level target.
 * Reference semantics: include/misc3d/utils.h:74-97 (std::mt19937, `rng_() % size_`).
 */
#include <cstdint>

#include "scan.h"

#if defined(__x86_64__) && defined(__GNUC__)
#include <immintrin.h>
#define M3D_HAVE_AVX2_DISPATCH 1
#endif

namespace m3d {

SampleStream::SampleStream(uint32_t seed, size_t n)
    : size((uint32_t)n), magic(n > 1 ? UINT64_MAX / (uint32_t)n + 1 : 0) {
    mt[0] = seed;
    for (uint32_t i = 1; i < 624; ++i) mt[i] = 1812433253u * (mt[i - 1] ^ (mt[i - 1] >> 30)) + i;
}

namespace {

inline uint32_t twist(uint32_t hi, uint32_t lo, uint32_t far) {
    const uint32_t y = (hi & 0x80000000u) | (lo & 0x7fffffffu);
    return far ^ (y >> 1) ^ ((0u - (y & 1u)) & 0x9908b0dfu);
}
inline void twist_all(uint32_t *mt) {
    for (int i = 0; i < 227; ++i) mt[i] = twist(mt[i], mt[i + 1], mt[i + 397]);
    for (int i = 227; i < 623; ++i) mt[i] = twist(mt[i], mt[i + 1], mt[i - 227]);
    mt[623] = twist(mt[623], mt[0], mt[396]);
}
inline uint32_t temper(uint32_t y) {
    y ^= y >> 11;
    y ^= (y << 7) & 0x9d2c5680u;
    y ^= (y << 15) & 0xefc60000u;
    y ^= y >> 18;
    return y;
}

void refill_generic(SampleStream &s) {
    twist_all(s.mt);
    if (s.size <= 1) {
        for (int i = 0; i < 624; ++i) s.idx[i] = 0;
        return;
    }
    for (int i = 0; i < 624; ++i) {
        const uint64_t low = s.magic * temper(s.mt[i]);
        s.idx[i] = (uint32_t)(((unsigned __int128)low * s.size) >> 64);
    }
}

#ifdef M3D_HAVE_AVX2_DISPATCH
__attribute__((target("avx2"))) void refill_avx2(SampleStream &s) {
    uint32_t *mt = s.mt;
    /* the twist: same recurrences, written so that the compiler vectorises them under this target */
    for (int i = 0; i < 227; ++i) mt[i] = twist(mt[i], mt[i + 1], mt[i + 397]);
    for (int i = 227; i < 623; ++i) mt[i] = twist(mt[i], mt[i + 1], mt[i - 227]);
    mt[623] = twist(mt[623], mt[0], mt[396]);
    if (s.size <= 1) {
        for (int i = 0; i < 624; ++i) s.idx[i] = 0;
        return;
    }
    const __m256i m7 = _mm256_set1_epi32((int)0x9d2c5680u), m15 = _mm256_set1_epi32((int)0xefc60000u);
    const __m256i ml = _mm256_set1_epi64x((long long)(s.magic & 0xffffffffull));
    const __m256i mh = _mm256_set1_epi64x((long long)(s.magic >> 32));
    const __m256i d = _mm256_set1_epi64x((long long)s.size);
    for (int i = 0; i < 624; i += 8) {
        __m256i y = _mm256_loadu_si256(reinterpret_cast<const __m256i *>(mt + i));
        y = _mm256_xor_si256(y, _mm256_srli_epi32(y, 11));
        y = _mm256_xor_si256(y, _mm256_and_si256(_mm256_slli_epi32(y, 7), m7));
        y = _mm256_xor_si256(y, _mm256_and_si256(_mm256_slli_epi32(y, 15), m15));
        y = _mm256_xor_si256(y, _mm256_srli_epi32(y, 18));
        __m256i r[2];
        for (int h = 0; h < 2; ++h) { /* four 64-bit lanes at a time, x in the low half of each lane */
            const __m256i x = _mm256_cvtepu32_epi64(h ? _mm256_extracti128_si256(y, 1) : _mm256_castsi256_si128(y));
            /* low = magic * x mod 2^64 */
            const __m256i low = _mm256_add_epi64(_mm256_mul_epu32(ml, x), _mm256_slli_epi64(_mm256_mul_epu32(mh, x), 32));
            /* (low * size) >> 64 = (hi32(low)*size + (lo32(low)*size >> 32)) >> 32 */
            const __m256i a = _mm256_mul_epu32(low, d);
            const __m256i b = _mm256_mul_epu32(_mm256_srli_epi64(low, 32), d);
            r[h] = _mm256_srli_epi64(_mm256_add_epi64(b, _mm256_srli_epi64(a, 32)), 32);
        }
        /* pack the eight 64-bit results (values < 2^32) back to eight 32-bit lanes, in order */
        const __m256i idx8 = _mm256_set_epi32(6, 4, 2, 0, 6, 4, 2, 0);
        const __m128i lo = _mm256_castsi256_si128(_mm256_permutevar8x32_epi32(r[0], idx8));
        const __m128i hi = _mm256_castsi256_si128(_mm256_permutevar8x32_epi32(r[1], idx8));
        _mm256_storeu_si256(reinterpret_cast<__m256i *>(s.idx + i), _mm256_set_m128i(hi, lo));
    }
}
#endif

}  // namespace

void SampleStream::refill() {
#ifdef M3D_HAVE_AVX2_DISPATCH
    static const bool avx2 = __builtin_cpu_supports("avx2");
    if (avx2)
        refill_avx2(*this);
    else
#endif
        refill_generic(*this);
    pos = 0;
}

/* Bulk form of draw(): the block position lives in a register and a row whose K next indices are already
 * distinct (all but ~K^2/size of them) is copied without the rejection loop. */
namespace {
template <int K>
void draw_rows_k(SampleStream &s, size_t rows, uint32_t *__restrict__ out) {
    int p = s.pos;
    const uint32_t *__restrict__ idx = s.idx;
    for (size_t r = 0; r < rows; ++r, out += K) {
        if (p + K <= 624) {
            uint32_t v[K];
            for (int i = 0; i < K; ++i) v[i] = idx[p + i];
            bool distinct = true;
            for (int i = 1; i < K; ++i)
                for (int j = 0; j < i; ++j) distinct &= (v[i] != v[j]);
            if (distinct) {
                for (int i = 0; i < K; ++i) out[i] = v[i];
                p += K;
                continue;
            }
        }
        s.pos = p; /* block boundary or a duplicate: the one-at-a-time loop (same semantics) */
        s.draw(K, out);
        p = s.pos;
    }
    s.pos = p;
}
}  // namespace

void SampleStream::draw_rows(int k, size_t rows, uint32_t *out) {
    switch (k) {
        case 2: draw_rows_k<2>(*this, rows, out); break;
        case 3: draw_rows_k<3>(*this, rows, out); break;
        case 4: draw_rows_k<4>(*this, rows, out); break;
        default:
            for (size_t r = 0; r < rows; ++r) draw(k, out + r * (size_t)k);
    }
}

/* for the tests: the generic and the dispatched body must agree */
extern "C" int m3d_sampler_selfcheck(uint32_t seed, size_t n, int blocks) {
    SampleStream a(seed, n), b(seed, n);
    for (int k = 0; k < blocks; ++k) {
        a.refill();
        refill_generic(b);
        for (int i = 0; i < 624; ++i)
            if (a.idx[i] != b.idx[i]) return 0;
    }
    return 1;
}

}  // namespace m3d
