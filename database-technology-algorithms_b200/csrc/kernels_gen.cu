// kernels_gen.cu -- synthetic block images generated directly in HBM ("G_syn").
//
// Same schema as the reference generator (main.cpp:41-77: 100 live rows per block, recid = row
// index, 5-letter strings, "Hola" at row 1 of every block, valid=1, block dummy=100) with
// counter-based randomness so that any sub-range is reproducible.  The arithmetic is identical,
// bit for bit, to orc_gen_syn in oracle/dbt_oracle.c (integer-only on purpose).
#include "dbt_internal.cuh"
#include <algorithm>
#include <cstring>

namespace dbt {

__device__ __forceinline__ uint64_t mix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}
__device__ __forceinline__ uint32_t bij32(uint32_t x, uint32_t seed) {
    x ^= seed;
    x *= 0x9E3779B1u; x ^= x >> 15;
    x *= 0x85EBCA77u; x ^= x >> 13;
    x *= 0xC2B2AE3Du; x ^= x >> 16;
    return x;
}
// ---- kind 4: exact Zipf(s = 1.1) ranks by rejection-inversion (Hoermann & Derflinger 1996) -----------------
// P(rank = k) ~ k^-1.1 over k in [1, U]; key = rho(rank - 1) with rho(j) = (j * A + C) mod U a bijection of [0, U)
// (A prime > 2^31, so gcd(A, U) = 1), which scatters the hot ranks over the key domain.  The sampler needs log and
// exp; to stay bit-identical with the CPU restatement (oracle/dbt_oracle.c orc_zipf_rank) they are built here from
// correctly rounded IEEE double +, *, / only (explicit round-to-nearest intrinsics on the device: no FMA contraction),
// in a fixed order of operations.
#ifdef __CUDA_ARCH__
#define ZADD(a, b) __dadd_rn((a), (b))
#define ZMUL(a, b) __dmul_rn((a), (b))
#define ZDIV(a, b) __ddiv_rn((a), (b))
#else
#define ZADD(a, b) ((a) + (b))
#define ZMUL(a, b) ((a) * (b))
#define ZDIV(a, b) ((a) / (b))
#endif
__host__ __device__ inline double z_bits2d(uint64_t b) {
    double d;
    memcpy(&d, &b, 8);
    return d;
}
__host__ __device__ inline uint64_t z_d2bits(double d) {
    uint64_t b;
    memcpy(&b, &d, 8);
    return b;
}
__host__ __device__ inline double z_log(double x) { // x > 0, normal
    uint64_t b = z_d2bits(x);
    int e = (int)((b >> 52) & 0x7FF) - 1023;
    double m = z_bits2d((b & 0x000FFFFFFFFFFFFFull) | 0x3FF0000000000000ull); // [1, 2)
    if (m > 1.4142135623730951) {
        m = ZMUL(m, 0.5);
        e += 1;
    }
    const double f = ZDIV(ZADD(m, -1.0), ZADD(m, 1.0)), f2 = ZMUL(f, f);
    double p = 1.0 / 23.0; // ln m = 2 f (1 + f2/3 + f2^2/5 + ... + f2^11/23)
    for (int k = 21; k >= 1; k -= 2) p = ZADD(ZMUL(p, f2), ZDIV(1.0, (double)k));
    return ZADD(ZMUL((double)e, 0.6931471805599453), ZMUL(ZMUL(2.0, f), p));
}
__host__ __device__ inline double z_exp(double y) { // |y| < 700
    const double kf = ZMUL(y, 1.4426950408889634);
    const long long k = (long long)(kf < 0 ? ZADD(kf, -0.5) : ZADD(kf, 0.5));
    const double r = ZADD(ZADD(y, -ZMUL((double)k, 0.693147180369123816490)), -ZMUL((double)k, 1.90821492927058770002e-10));
    double p = 1.0 / 6227020800.0; // Taylor to degree 13, |r| <= 0.35
    const double inv_fact[13] = {1.0 / 479001600.0, 1.0 / 39916800.0, 1.0 / 3628800.0, 1.0 / 362880.0, 1.0 / 40320.0,
                                 1.0 / 5040.0,      1.0 / 720.0,      1.0 / 120.0,     1.0 / 24.0,     1.0 / 6.0,
                                 0.5,               1.0,              1.0};
    for (int i = 0; i < 13; ++i) p = ZADD(ZMUL(p, r), inv_fact[i]);
    return ZMUL(p, z_bits2d((uint64_t)(k + 1023) << 52));
}
constexpr double kZipfS = 1.1;
__host__ __device__ inline double z_h(double x) { return z_exp(ZMUL(-kZipfS, z_log(x))); }                       // x^-s
__host__ __device__ inline double z_H(double x) {                                                                // (x^(1-s) - 1) / (1-s)
    return ZDIV(ZADD(z_exp(ZMUL(1.0 - kZipfS, z_log(x))), -1.0), 1.0 - kZipfS);
}
__host__ __device__ inline double z_Hinv(double u) { // inverse of z_H
    double t = ZADD(1.0, ZMUL(1.0 - kZipfS, u));
    if (t < 1e-300) t = 1e-300;
    return z_exp(ZDIV(z_log(t), 1.0 - kZipfS));
}
__host__ __device__ inline uint64_t z_mix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}
struct ZipfConst {
    double h_x1, h_n, s_cut; // H(1.5) - 1, H(U + 0.5), 2 - Hinv(H(2.5) - h(2))
};
__host__ __device__ inline ZipfConst zipf_const(uint64_t U) {
    ZipfConst c;
    c.h_x1 = ZADD(z_H(1.5), -1.0);
    c.h_n = z_H(ZADD((double)U, 0.5));
    c.s_cut = ZADD(2.0, -z_Hinv(ZADD(z_H(2.5), -z_h(2.0))));
    return c;
}
__host__ __device__ inline uint64_t zipf_rank(uint64_t seed, uint64_t U, uint64_t r, const ZipfConst &c) {
    for (uint64_t it = 0;; ++it) {
        const uint64_t hsh = z_mix64(z_mix64(seed * 0x100000001B3ull + r) + it * 0xD6E8FEB86659FD93ull);
        const double u01 = ZMUL((double)(hsh >> 11), 1.1102230246251565e-16); // [0, 1)
        const double u = ZADD(c.h_n, ZMUL(u01, ZADD(c.h_x1, -c.h_n)));
        const double x = z_Hinv(u);
        long long k = (long long)ZADD(x, 0.5);
        if (k < 1) k = 1;
        if ((uint64_t)k > U) k = (long long)U;
        if (ZADD((double)k, -x) <= c.s_cut || u >= ZADD(z_H(ZADD((double)k, 0.5)), -z_h((double)k)) || it >= 63) return (uint64_t)k;
    }
}

__device__ __forceinline__ uint32_t syn_num(uint64_t seed, uint64_t n, uint64_t U, int kind, uint64_t r, const ZipfConst &zc) {
    const uint64_t A = 2654435761ull, C = 40503ull;
    if (kind == 4) {
        const uint64_t rank = zipf_rank(seed, U, r, zc); // 1..U, P(rank) ~ rank^-1.1
        return (uint32_t)(((rank - 1) * A + C) % U);
    }
    if (kind == 0) {
        uint64_t j = (r * A + C) % n; // r < 2^32, A < 2^32: no overflow
        return bij32((uint32_t)(j % U), (uint32_t)seed);
    } else if (kind == 1) {
        return (uint32_t)(mix64(seed * 0x100000001B3ull + r) % U);
    } else {
        uint64_t h = mix64(seed * 0x100000001B3ull + r);
        uint64_t p = h >> 32;
        for (int i = 0; i < 3; ++i) p = (p * p) >> 32; // u^8 in 0.32 fixed point
        uint64_t rank = (p * U) >> 32;
        return (uint32_t)(bij32((uint32_t)rank, (uint32_t)seed) % U);
    }
}

__global__ void __launch_bounds__(128)
gen_syn_kernel(uint64_t seed, uint64_t n_total, uint64_t U, int kind, uint64_t row0, uint64_t nrows, uint32_t recid0,
               uint32_t *__restrict__ img, uint64_t nblocks, ZipfConst zc) {
    uint64_t b = blockIdx.x;
    for (; b < nblocks; b += gridDim.x) {
        uint32_t *blk = img + b * kBlockWords;
        uint64_t lo = b * kRpb;
        uint32_t live = (uint32_t)min((uint64_t)kRpb, nrows - lo);
        if (threadIdx.x == 0) {
            blk[0] = (uint32_t)(row0 / kRpb + b);
            blk[1] = live;
            blk[kTrailerWord] = 1;
            blk[kTrailerWord + 1] = live;
        }
        if (threadIdx.x < kRpb) {
            uint32_t e = threadIdx.x;
            uint32_t *rec = blk + kEntriesWord + e * kRecWords;
            if (e < live) {
                uint64_t r = row0 + lo + e;
                uint32_t w[kRecWords];
#pragma unroll
                for (int i = 0; i < (int)kRecWords; ++i) w[i] = 0;
                w[0] = recid0 + (uint32_t)r;
                // kind 3: with probability 1/2 this row copies the (num, str) of a pseudo-random row of the PARTNER
                // relation (generated with kind 1 and seed ^ 0x5EED), so that about half of the composite keys match
                uint64_t kseed = seed, krow = r;
                int kkind = kind;
                if (kind == 3) {
                    const uint64_t h3 = mix64(seed * 0x100000001B3ull + r + 0x777ull);
                    kkind = 1;
                    if (h3 & 1) {
                        kseed = seed ^ 0x5EEDull;
                        krow = (h3 >> 1) % n_total;
                    }
                }
                w[1] = syn_num(kseed, n_total, U, kkind, krow, zc);
                r = krow; // the string below follows the same (seed, row)
                const uint64_t seed_s = kseed;
                if (r % kRpb == 1) {
                    w[2] = 0x616C6F48u; // "Hola" little-endian
                } else {
                    uint64_t h = mix64((seed_s ^ 0x5bd1e995ull) * 0x100000001B3ull + r);
                    uint32_t c0 = 'a' + (uint32_t)(h % 26); h /= 26;
                    uint32_t c1 = 'a' + (uint32_t)(h % 26); h /= 26;
                    uint32_t c2 = 'a' + (uint32_t)(h % 26); h /= 26;
                    uint32_t c3 = 'a' + (uint32_t)(h % 26); h /= 26;
                    uint32_t c4 = 'a' + (uint32_t)(h % 26);
                    w[2] = c0 | (c1 << 8) | (c2 << 16) | (c3 << 24);
                    w[3] = c4;
                }
                w[32] = 1; // valid
#pragma unroll
                for (int i = 0; i < (int)kRecWords; ++i) rec[i] = w[i];
            } else {
                for (int i = 0; i < (int)kRecWords; ++i) rec[i] = 0;
            }
        }
    }
}

int gen_syn(uint64_t seed, uint64_t n_total, uint64_t U, int kind, uint64_t row0, uint64_t nrows, uint32_t recid0,
            void *d_image, cudaStream_t st) {
    if (nrows == 0) return 0;
    if (U == 0 || n_total == 0 || n_total >= (1ull << 32)) {
        set_error("gen_syn: bad sizes");
        return DBT_ERR_ARG;
    }
    uint64_t nb = (nrows + kRpb - 1) / kRpb;
    int grid = (int)std::min<uint64_t>(nb, 148 * 64);
    if (kind < 0 || kind > 4 || (kind == 4 && U % 2654435761ull == 0)) {
        set_error("gen_syn: bad kind");
        return DBT_ERR_ARG;
    }
    ZipfConst zc{0, 0, 0};
    if (kind == 4) zc = zipf_const(U); // host arithmetic identical to the device's (IEEE +, *, / only)
    gen_syn_kernel<<<grid, 128, 0, st>>>(seed, n_total, U, kind, row0, nrows, recid0, (uint32_t *)d_image, nb, zc);
    count_launch();
    DBT_KERNEL_CHECK();
    return 0;
}

} // namespace dbt

extern "C" int dbt_gen_syn(uint64_t seed, uint64_t n_total, uint64_t U, int kind, uint64_t row0, uint64_t nrows,
                           uint32_t recid0, void *d_image, void *stream) {
    if (!d_image) {
        dbt::set_error("dbt_gen_syn: NULL image");
        return DBT_ERR_ARG;
    }
    return dbt::gen_syn(seed, n_total, U, kind, row0, nrows, recid0, d_image, (cudaStream_t)stream);
}
