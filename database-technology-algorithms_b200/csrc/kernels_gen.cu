// kernels_gen.cu -- synthetic block images generated directly in HBM ("G_syn").
//
// Same schema as the reference generator (main.cpp:41-77: 100 live rows per block, recid = row
// index, 5-letter strings, "Hola" at row 1 of every block, valid=1, block dummy=100) with
// counter-based randomness so that any sub-range is reproducible.  The arithmetic is identical,
// bit for bit, to orc_gen_syn in oracle/dbt_oracle.c (integer-only on purpose).
#include "dbt_internal.cuh"

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
__device__ __forceinline__ uint32_t syn_num(uint64_t seed, uint64_t n, uint64_t U, int kind, uint64_t r) {
    const uint64_t A = 2654435761ull, C = 40503ull;
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
               uint32_t *__restrict__ img, uint64_t nblocks) {
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
                w[1] = syn_num(kseed, n_total, U, kkind, krow);
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
    gen_syn_kernel<<<grid, 128, 0, st>>>(seed, n_total, U, kind, row0, nrows, recid0, (uint32_t *)d_image, nb);
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
