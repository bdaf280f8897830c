// kernels_sort.cu -- LSD "onesweep" radix sort of (u32 key, u32 row) pairs, hand-written for sm_100a.
//
// Replaces the reference's run generation (qsort of 140-byte records, DatabaseProject.cpp:207-214)
// and its priority_queue k-way merge (DatabaseProject.cpp:245-369): the sort happens on 8-byte
// (key, row) pairs, records move once at the end (kernels_gather.cu).
//
// One pass = one kernel: every CTA takes a tile (dynamic tile id), ranks its keys on an 8-bit digit
// with warp match-any, publishes its per-digit counts in a decoupled look-back chain, stages the
// tile sorted-by-digit in shared memory and writes digit runs coalesced.  Per pass the DRAM
// traffic is 8 B read + 8 B written per pair (keys+rows) -- the kernel is HBM-bound.
#include "dbt_internal.cuh"
#include "sort_common.cuh"
#include <cstdlib>
#include <cstring>
#include <vector>

namespace dbt {

// Decoupled look-back for one digit column: sum the aggregates of the predecessor tiles until one
// with an inclusive prefix is met.  Four predecessors are fetched per round trip (speculatively).
template <typename S>
__device__ __forceinline__ uint32_t lookback_exclusive(const S *state, uint32_t tile, uint32_t col) {
    using TS = TileState<S>;
    uint32_t excl = 0; // positions are < 2^32 whatever the width of the state words
    int p = (int)tile - 1;
    while (p >= 0) {
        S s[4];
#pragma unroll
        for (int q = 0; q < 4; ++q)
            s[q] = (p - q >= 0) ? TS::ld(&state[(size_t)(p - q) * kRadix + col]) : TS::kInc;
        int q = 0;
        bool done = false;
#pragma unroll
        for (; q < 4; ++q) {
            if ((s[q] >> TS::kFlagShift) == 0) break; // not published yet: poll again from here
            excl += (uint32_t)(s[q] & TS::kMask);
            if (s[q] & TS::kInc) {
                done = true;
                break;
            }
        }
        if (done) break;
        p -= q;
    }
    return excl;
}

// ---------------------------------------------------------------------------------------------
// OR / AND reduction of a word column: the bits where OR and AND differ are the only bits a sort
// has to look at (constant bits cannot change the order).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) or_and_kernel(const uint32_t *__restrict__ w, uint64_t n, uint32_t *out) {
    uint32_t o = 0, a = 0xFFFFFFFFu;
    uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        uint32_t v = w[i];
        o |= v;
        a &= v;
    }
    o = __reduce_or_sync(0xFFFFFFFFu, o);
    a = __reduce_and_sync(0xFFFFFFFFu, a);
    if ((threadIdx.x & 31) == 0) {
        atomicOr(&out[0], o);
        atomicAnd(&out[1], a);
    }
}
__global__ void init_or_and_kernel(uint32_t *out) {
    out[0] = 0;
    out[1] = 0xFFFFFFFFu;
}
int or_and_reduce(const uint32_t *d_words, uint64_t n, uint32_t *d_or_and, cudaStream_t st) {
    init_or_and_kernel<<<1, 1, 0, st>>>(d_or_and);
    count_launch();
    if (n) {
        int grid = (int)std::min<uint64_t>((n + 255) / 256, 148 * 8);
        or_and_kernel<<<grid, 256, 0, st>>>(d_words, n, d_or_and);
        count_launch();
    }
    DBT_KERNEL_CHECK();
    return 0;
}

// ---------------------------------------------------------------------------------------------
// Histogram of up to 4 digit positions in one read of the keys, then per-digit exclusive scan.
// ---------------------------------------------------------------------------------------------
struct DigitPlan {
    int npass;
    int shift[4];
};

__global__ void __launch_bounds__(512) hist_kernel(const uint32_t *__restrict__ keys, uint64_t n, DigitPlan plan,
                                                   uint32_t *__restrict__ ghist /*[4][256]*/) {
    __shared__ uint32_t sh[4][kRadix];
    for (int i = threadIdx.x; i < 4 * kRadix; i += blockDim.x) (&sh[0][0])[i] = 0;
    __syncthreads();
    // the caller's buffer may be only 4-byte aligned: scalar head up to the first 16-byte boundary, vector body, scalar tail
    const uint64_t head = min(n, (uint64_t)(((16u - (uint32_t)((uintptr_t)keys & 15u)) & 15u) >> 2));
    const uint64_t nvec = (n - head) / 4;
    const uint4 *kv = reinterpret_cast<const uint4 *>(keys + head);
    uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += stride) {
        uint4 v = kv[i];
#pragma unroll
        for (int p = 0; p < 4; ++p) {
            if (p < plan.npass) {
                int s = plan.shift[p];
                atomicAdd(&sh[p][(v.x >> s) & 0xFF], 1u);
                atomicAdd(&sh[p][(v.y >> s) & 0xFF], 1u);
                atomicAdd(&sh[p][(v.z >> s) & 0xFF], 1u);
                atomicAdd(&sh[p][(v.w >> s) & 0xFF], 1u);
            }
        }
    }
    // head and tail elements (at most 3 + 3) by the first threads of block 0
    if (blockIdx.x == 0 && threadIdx.x < 8) {
        const uint64_t ntail = (n - head) & 3;
        uint64_t idx = n; // none
        if (threadIdx.x < head) idx = threadIdx.x;
        else if (threadIdx.x >= 4 && threadIdx.x - 4 < ntail) idx = head + nvec * 4 + (threadIdx.x - 4);
        if (idx < n) {
            uint32_t k = keys[idx];
            for (int p = 0; p < plan.npass; ++p) atomicAdd(&sh[p][(k >> plan.shift[p]) & 0xFF], 1u);
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < plan.npass * kRadix; i += blockDim.x) {
        uint32_t c = (&sh[0][0])[i];
        if (c) atomicAdd(&ghist[i], c);
    }
}

// OR / AND of the keys and the histograms of their four bytes in ONE read (the pair-sort entry point: the digit
// plan is only known after the OR/AND, but when it turns out byte-aligned -- every full-range sort -- the byte
// histograms are exactly the planned digits' and the separate histogram read is skipped).  16-byte aligned keys.
__global__ void __launch_bounds__(512) or_and_hist_kernel(const uint32_t *__restrict__ keys, uint64_t n,
                                                          uint32_t *__restrict__ or_and /*2, initialised*/,
                                                          uint32_t *__restrict__ ghist /*[4][256], zeroed*/) {
    __shared__ uint32_t sh[4][kRadix];
    for (int i = threadIdx.x; i < 4 * kRadix; i += blockDim.x) (&sh[0][0])[i] = 0;
    __syncthreads();
    uint32_t o = 0, a = 0xFFFFFFFFu;
    auto one = [&](uint32_t k) {
        o |= k;
        a &= k;
        atomicAdd(&sh[0][k & 0xFF], 1u);
        atomicAdd(&sh[1][(k >> 8) & 0xFF], 1u);
        atomicAdd(&sh[2][(k >> 16) & 0xFF], 1u);
        atomicAdd(&sh[3][k >> 24], 1u);
    };
    const uint64_t nvec = n / 4;
    const uint4 *kv = reinterpret_cast<const uint4 *>(keys);
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += stride) {
        const uint4 v = kv[i];
        one(v.x);
        one(v.y);
        one(v.z);
        one(v.w);
    }
    if (blockIdx.x == 0 && threadIdx.x < (n & 3)) one(keys[nvec * 4 + threadIdx.x]);
    o = __reduce_or_sync(0xFFFFFFFFu, o);
    a = __reduce_and_sync(0xFFFFFFFFu, a);
    if ((threadIdx.x & 31) == 0) {
        atomicOr(&or_and[0], o);
        atomicAnd(&or_and[1], a);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 4 * kRadix; i += blockDim.x) {
        const uint32_t c = (&sh[0][0])[i];
        if (c) atomicAdd(&ghist[i], c);
    }
}

// in-place exclusive scan of each 256-bin histogram (one CTA per digit position)
__global__ void __launch_bounds__(kRadix) hist_scan_kernel(uint32_t *ghist) {
    __shared__ uint32_t wsum[8];
    uint32_t *h = ghist + blockIdx.x * kRadix;
    uint32_t v = h[threadIdx.x], x = v;
    int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t t = __shfl_up_sync(0xFFFFFFFFu, x, o);
        if (lane >= o) x += t;
    }
    if (lane == 31) wsum[warp] = x;
    __syncthreads();
    uint32_t pre = 0;
    for (int w = 0; w < warp; ++w) pre += wsum[w];
    h[threadIdx.x] = pre + x - v;
}

// ---------------------------------------------------------------------------------------------
// The onesweep pass.
// ---------------------------------------------------------------------------------------------
template <int THREADS, int ITEMS>
struct OnesweepSmem {
    static constexpr int WARPS = THREADS / 32;
    static constexpr int TILE = THREADS * ITEMS;
    uint32_t hist[WARPS][kRadix]; // per-warp digit counters, later exclusive-over-warps offsets
    uint32_t excl[kRadix];        // tile-local exclusive digit offsets
    uint32_t goff[kRadix];        // global offset of the digit run minus excl[d]
    uint32_t wsum[8];
    uint32_t tile;
    uint32_t keys[TILE];
    uint32_t vals[TILE];
};

template <int THREADS, int ITEMS, bool HAS_VALS, bool IOTA_VALS, typename S>
__global__ void __launch_bounds__(THREADS)
onesweep_kernel(const uint32_t *__restrict__ kin, uint32_t *__restrict__ kout, const uint32_t *__restrict__ vin,
                uint32_t *__restrict__ vout, uint32_t n, int shift, const uint32_t *__restrict__ digit_base,
                S *state /*[ntiles][256], zeroed*/, uint32_t *tile_ctr /*zeroed*/) {
    using TS = TileState<S>;
    using Smem = OnesweepSmem<THREADS, ITEMS>;
    constexpr int WARPS = Smem::WARPS;
    constexpr int TILE = Smem::TILE;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    Smem &sm = *reinterpret_cast<Smem *>(smem_raw);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) sm.tile = atomicAdd(tile_ctr, 1u); // dynamic tile id => every predecessor tile is resident or done
    for (int i = tid; i < WARPS * kRadix; i += THREADS) (&sm.hist[0][0])[i] = 0;
    __syncthreads();
    const uint32_t tile = sm.tile;
    const uint32_t base = tile * (uint32_t)TILE;
    const bool full = base + (uint32_t)TILE <= n;
    const uint32_t idx0 = base + warp * (ITEMS * 32) + lane; // warp-striped arrangement

    uint32_t key[ITEMS];
#pragma unroll
    for (int j = 0; j < ITEMS; ++j) {
        uint32_t idx = idx0 + j * 32;
        key[j] = (full || idx < n) ? kin[idx] : 0xFFFFFFFFu;
    }

    // ---- rank inside the warp.  The lanes holding my digit ("peers") come from 8 warp ballots, one per
    // digit bit (VOTE + LOP3 on the ALU path).  match.any would give the same mask in one instruction
    // but runs on the ADU pipe at ~2 cycles per distinct value in the warp (ncu: 64% ADU, the limiter
    // of the first version, profiles/r01_notes.md).  The group leader bumps the warp's private counter
    // once for the whole group, so there are no shared-memory atomics.
    uint32_t rank[ITEMS];
    const uint32_t lt = lanemask_lt();
#pragma unroll
    for (int j = 0; j < ITEMS; ++j) {
        const uint32_t d = (key[j] >> shift) & 0xFFu;
        uint32_t peers = 0xFFFFFFFFu;
#pragma unroll
        for (int b = 0; b < 8; ++b) {
            const bool bit = (d >> b) & 1u;
            const uint32_t m = __ballot_sync(0xFFFFFFFFu, bit);
            peers &= bit ? m : ~m;
        }
        bool valid = true;
        if (!full) {
            valid = idx0 + j * 32 < n;
            const uint32_t vm = __ballot_sync(0xFFFFFFFFu, valid);
            peers &= valid ? vm : ~vm;
        }
        const int leader = __ffs(peers) - 1;
        uint32_t old = 0;
        if (lane == leader && valid) {
            old = sm.hist[warp][d];
            sm.hist[warp][d] = old + __popc(peers);
        }
        old = __shfl_sync(0xFFFFFFFFu, old, leader);
        rank[j] = old + __popc(peers & lt);
        __syncwarp();
    }
    __syncthreads();

    // ---- per digit: exclusive over warps, tile count, publish the aggregate at once
    uint32_t count_d = 0, incl = 0;
    if (tid < kRadix) {
        uint32_t acc = 0;
#pragma unroll
        for (int w = 0; w < WARPS; ++w) {
            uint32_t t = sm.hist[w][tid];
            sm.hist[w][tid] = acc;
            acc += t;
        }
        count_d = acc;
        TS::st(&state[(size_t)tile * kRadix + tid], (tile == 0 ? TS::kInc : TS::kAgg) | count_d);
        // exclusive scan of the 256 counts (8 warps)
        incl = count_d;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            uint32_t t = __shfl_up_sync(0xFFFFFFFFu, incl, o);
            if (lane >= o) incl += t;
        }
        if (lane == 31) sm.wsum[warp] = incl;
    }
    __syncthreads();
    if (tid < kRadix) {
        uint32_t pre = 0;
        for (int w = 0; w < warp; ++w) pre += sm.wsum[w];
        sm.excl[tid] = pre + incl - count_d;
    }
    __syncthreads();

    // ---- stage the tile in shared memory ordered by digit (look-back latency hides behind this)
    uint32_t pos[ITEMS];
#pragma unroll
    for (int j = 0; j < ITEMS; ++j) {
        const uint32_t d = (key[j] >> shift) & 0xFFu;
        pos[j] = sm.excl[d] + sm.hist[warp][d] + rank[j];
        if (full || (idx0 + j * 32 < n)) sm.keys[pos[j]] = key[j];
    }
    if (HAS_VALS) {
#pragma unroll
        for (int j = 0; j < ITEMS; ++j) {
            uint32_t idx = idx0 + j * 32;
            if (full || idx < n) sm.vals[pos[j]] = IOTA_VALS ? idx : vin[idx];
        }
    }

    // ---- decoupled look-back, one thread per digit
    if (tid < kRadix) {
        uint32_t excl_prefix = 0;
        if (tile > 0) {
            excl_prefix = lookback_exclusive(state, tile, tid);
            TS::st(&state[(size_t)tile * kRadix + tid], TS::kInc | (S)(excl_prefix + count_d));
        }
        sm.goff[tid] = digit_base[tid] + excl_prefix - sm.excl[tid];
    }
    __syncthreads();

    // ---- coalesced writes of the digit runs
    const uint32_t nvalid = full ? (uint32_t)TILE : n - base;
#pragma unroll
    for (int j = 0; j < ITEMS; ++j) {
        uint32_t i = tid + j * THREADS;
        if (i < nvalid) {
            uint32_t k = sm.keys[i];
            uint32_t dst = sm.goff[(k >> shift) & 0xFFu] + i;
            kout[dst] = k;
            if (HAS_VALS) vout[dst] = sm.vals[i];
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Onesweep pass, version 2: persistent CTAs with a double-buffered TMA bulk-copy pipeline.
//
// Why (ncu, profiles/r01_notes.md): version 1 is latency-bound -- the loads of a tile are only issued
// when its CTA starts, ranking waits on them, and MATCH.ANY (ADU pipe) alone caps the pass at ~36% of
// the HBM roofline.  Here every CTA keeps the NEXT tile's keys and rows in flight with
// cp.async.bulk (global -> shared, completion on an mbarrier) while it ranks the current tile out of
// shared memory, ranks with warp ballots (ALU pipe), optionally mixing in match.any (ADU pipe) so
// both pipes work, and reuses the stage buffer for the digit-ordered staging of the scatter.
// ---------------------------------------------------------------------------------------------
template <int THREADS, int ITEMS>
struct Os2Smem {
    static constexpr int WARPS = THREADS / 32;
    static constexpr int TILE = THREADS * ITEMS;
    alignas(128) uint32_t keys[2][TILE];
    alignas(128) uint32_t vals[2][TILE];
    uint32_t hist[WARPS][kRadix]; // per-warp digit counters -> (tile-exclusive + warp-exclusive) offsets
    uint32_t goff[kRadix];        // global offset of the digit run minus its tile-local offset
    uint32_t wsum[8];
    alignas(8) uint64_t mbar[2];
    uint32_t next_tile[2];
};

template <int THREADS, int ITEMS, bool HAS_VALS, bool IOTA_VALS, int RANK, bool FULL, typename S>
__device__ __forceinline__ void os2_process_tile(Os2Smem<THREADS, ITEMS> &sm, int stg, uint32_t tile,
                                             const uint32_t *__restrict__ kin, uint32_t *__restrict__ kout,
                                             const uint32_t *__restrict__ vin, uint32_t *__restrict__ vout,
                                             uint32_t n, int shift, const uint32_t *__restrict__ digit_base,
                                             S *state) {
    using TS = TileState<S>;
    using Smem = Os2Smem<THREADS, ITEMS>;
    constexpr int WARPS = Smem::WARPS;
    constexpr int TILE = Smem::TILE;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t lt = lanemask_lt();
    const uint32_t base = tile * (uint32_t)TILE;
    const uint32_t nvalid = FULL ? (uint32_t)TILE : n - base;
    const uint32_t nbulk = FULL ? (uint32_t)TILE : (((nvalid * 4u) & ~15u) >> 2);
    uint32_t *skeys = sm.keys[stg];
    uint32_t *svals = sm.vals[stg];
    const uint32_t li0 = warp * (ITEMS * 32) + lane; // warp-striped arrangement inside the tile

    uint32_t key[ITEMS];
#pragma unroll
    for (int j = 0; j < ITEMS; ++j) {
        const uint32_t li = li0 + j * 32;
        if (FULL) key[j] = skeys[li];
        else key[j] = (li < nbulk) ? skeys[li] : ((li < nvalid) ? kin[base + li] : 0xFFFFFFFFu);
    }

    // ---- rank inside the warp (positions fit 16 bits: two per register)
    uint32_t rank2[(ITEMS + 1) / 2];
#pragma unroll
    for (int j = 0; j < (ITEMS + 1) / 2; ++j) rank2[j] = 0;
#pragma unroll
    for (int j = 0; j < ITEMS; ++j) {
        const uint32_t d = (key[j] >> shift) & 0xFFu;
        bool valid = true;
        uint32_t peers;
        if (RANK == 0) {
            if (!FULL) valid = li0 + j * 32 < nvalid;
            peers = __match_any_sync(0xFFFFFFFFu, valid ? d : 0x100u);
        } else if (RANK == 23) {
            // the hybrid below with the ballots written so that ptxas sets the bit predicates with one R2P and spends
            // VOTE + predicated NOT + OR per bit (3 instructions instead of the 6 it makes of `bit ? m : ~m`):
            // differing lanes are collected per bit and removed from the match mask at the end
            constexpr int K = 3; // (2: 0.496 ms per pass, 3: 0.498, 4: 0.570)
            if (!FULL) valid = li0 + j * 32 < nvalid;
            peers = __match_any_sync(0xFFFFFFFFu, valid ? (d & ((1u << K) - 1u)) : (1u << K));
            uint32_t mm[8 - K]; // lanes whose bit b differs from mine
#pragma unroll
            for (int b = K; b < 8; ++b) {
                uint32_t m, x;
                asm("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %2, 0;\n\tvote.sync.ballot.b32 %0, p, 0xffffffff;\n\tselp.u32 %1, 0xffffffff, 0, p;\n\t}"
                    : "=r"(m), "=r"(x)
                    : "r"(d & (1u << b)));
                mm[b - K] = m ^ x;
            }
            // peers = match & ~(mm[0] | ... ): three-input LOP3s (ptxas keeps two-input ORs otherwise)
            uint32_t diff = mm[0];
#pragma unroll
            for (int t = 1; t + 1 < 8 - K; t += 2) asm("lop3.b32 %0, %0, %1, %2, 0xFE;" : "+r"(diff) : "r"(mm[t]), "r"(mm[t + 1]));
            if ((8 - K) % 2 == 0) asm("lop3.b32 %0, %0, %1, %2, 0x10;" : "+r"(peers) : "r"(diff), "r"(mm[8 - K - 1])); // a & ~(b | c)
            else peers &= ~diff;
        } else if (RANK == 13) {
            // hybrid: match.any on the low K bits (cost ~ number of distinct values: <= 2^K groups, ADU pipe)
            // and one ballot per remaining bit (ALU pipe): every item loads both pipes lightly
            constexpr int K = 3;
            if (!FULL) valid = li0 + j * 32 < nvalid;
            peers = __match_any_sync(0xFFFFFFFFu, valid ? (d & ((1u << K) - 1u)) : (1u << K));
#pragma unroll
            for (int b = K; b < 8; ++b) {
                const bool bit = (d >> b) & 1u;
                const uint32_t m = __ballot_sync(0xFFFFFFFFu, bit);
                peers &= bit ? m : ~m;
            }
        } else {
            peers = 0xFFFFFFFFu;
#pragma unroll
            for (int b = 0; b < 8; ++b) {
                const bool bit = (d >> b) & 1u;
                const uint32_t m = __ballot_sync(0xFFFFFFFFu, bit);
                peers &= bit ? m : ~m;
            }
            if (!FULL) {
                valid = li0 + j * 32 < nvalid;
                const uint32_t vm = __ballot_sync(0xFFFFFFFFu, valid);
                peers &= valid ? vm : ~vm;
            }
        }
        if (RANK == 23) {
            // every lane reads the group's counter itself (one broadcast LDS) before the group's first lane bumps it: no
            // leader election (BREV + FLO), no shuffle
            const uint32_t before = peers & lt;
            uint32_t *cnt = &sm.hist[warp][0] + d;
            const uint32_t old = *cnt;
            const uint32_t ahead = __popc(before);
            if (ahead == 0 && valid) *cnt = old + __popc(peers);
            rank2[j >> 1] += (old + ahead) << ((j & 1) * 16); // (disjoint fields: + is |, and folds into one IMAD)
        } else {
            const int leader = __ffs(peers) - 1;
            uint32_t old = 0;
            if (lane == leader && valid) {
                old = sm.hist[warp][d];
                sm.hist[warp][d] = old + __popc(peers);
            }
            old = __shfl_sync(0xFFFFFFFFu, old, leader);
            rank2[j >> 1] |= (old + __popc(peers & lt)) << ((j & 1) * 16);
        }
        __syncwarp();
    }
    uint32_t val[HAS_VALS ? ITEMS : 1];
    if (HAS_VALS) {
#pragma unroll
        for (int j = 0; j < ITEMS; ++j) {
            const uint32_t li = li0 + j * 32;
            if (IOTA_VALS) val[j] = base + li;
            else if (FULL) val[j] = svals[li];
            else val[j] = (li < nbulk) ? svals[li] : ((li < nvalid) ? vin[base + li] : 0u);
        }
    }
    __syncthreads(); // every key/row of the tile is in registers; warp histograms are complete

    // ---- per digit: count, publish the aggregate at once, tile-local exclusive offsets
    uint32_t count_d = 0, incl = 0;
    if (tid < kRadix) {
#pragma unroll
        for (int w = 0; w < WARPS; ++w) count_d += sm.hist[w][tid];
        TS::st(&state[(size_t)tile * kRadix + tid], (tile == 0 ? TS::kInc : TS::kAgg) | count_d);
        incl = count_d;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, incl, o);
            if (lane >= o) incl += t;
        }
        if (lane == 31) sm.wsum[warp] = incl;
    }
    __syncthreads();
    uint32_t excl_d = 0;
    if (tid < kRadix) {
        uint32_t pre = 0;
        for (int w = 0; w < warp; ++w) pre += sm.wsum[w];
        excl_d = pre + incl - count_d;
        uint32_t acc = excl_d; // hist[w][d] := tile-local offset of warp w's first key with digit d
#pragma unroll
        for (int w = 0; w < WARPS; ++w) {
            const uint32_t t = sm.hist[w][tid];
            sm.hist[w][tid] = acc;
            acc += t;
        }
    }
    __syncthreads();

    // ---- stage the tile ordered by digit, in place (the look-back latency hides behind this)
#pragma unroll
    for (int j = 0; j < ITEMS; ++j) {
        const uint32_t d = (key[j] >> shift) & 0xFFu;
        const uint32_t pos = sm.hist[warp][d] + ((rank2[j >> 1] >> ((j & 1) * 16)) & 0xFFFFu);
        if (FULL || (li0 + j * 32 < nvalid)) {
            skeys[pos] = key[j];
            if (HAS_VALS) svals[pos] = val[j];
        }
    }
    if (tid < kRadix) { // decoupled look-back, one thread per digit
        uint32_t excl_prefix = 0;
        if (tile > 0) {
            excl_prefix = lookback_exclusive(state, tile, tid);
            TS::st(&state[(size_t)tile * kRadix + tid], TS::kInc | (S)(excl_prefix + count_d));
        }
        sm.goff[tid] = digit_base[tid] + excl_prefix - excl_d;
    }
    __syncthreads();

    // ---- coalesced writes of the digit runs
#pragma unroll
    for (int j = 0; j < ITEMS; ++j) {
        const uint32_t i = tid + j * THREADS;
        if (FULL || i < nvalid) {
            const uint32_t k = skeys[i];
            const uint32_t dst = sm.goff[(k >> shift) & 0xFFu] + i;
            kout[dst] = k;
            if (HAS_VALS) vout[dst] = svals[i];
        }
    }
}

// RANK: 23 = match.any on the 3 low digit bits + 5 hand-scheduled ballots (the default); kept for comparison runs
// (DBT_ONESWEEP_RANK): 0 = match.any on the whole digit, 1 = 8 ballots, 13 = 23 as the compiler schedules `bit ? m : ~m`
template <int THREADS, int ITEMS, bool HAS_VALS, bool IOTA_VALS, int RANK, typename S>
__global__ void __launch_bounds__(THREADS, (THREADS * ITEMS <= 4096) ? 3 : ((THREADS * ITEMS <= 6144) ? 2 : 1))
onesweep2_kernel(const uint32_t *__restrict__ kin, uint32_t *__restrict__ kout, const uint32_t *__restrict__ vin,
                 uint32_t *__restrict__ vout, uint32_t n, int shift, const uint32_t *__restrict__ digit_base,
                 S *state /*[ntiles][256], zeroed*/, uint32_t *tile_ctr /*zeroed*/) {
    using Smem = Os2Smem<THREADS, ITEMS>;
    constexpr int WARPS = Smem::WARPS;
    constexpr int TILE = Smem::TILE;
    constexpr bool LOAD_VALS = HAS_VALS && !IOTA_VALS;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    Smem &sm = *reinterpret_cast<Smem *>(smem_raw);
    const int tid = threadIdx.x;
    const uint32_t ntiles = (n + TILE - 1) / TILE;

    auto issue = [&](uint32_t t, int stg) { // thread 0: start the bulk copies of tile t into stage stg
        const uint32_t base = t * (uint32_t)TILE;
        const uint32_t valid = min((uint32_t)TILE, n - base);
        const uint32_t bytes = (valid * 4u) & ~15u;
        fence_proxy_async(); // earlier generic-proxy writes to this stage are ordered before the async writes
        mbar_expect_tx(&sm.mbar[stg], LOAD_VALS ? 2 * bytes : bytes);
        if (bytes) {
            bulk_g2s(sm.keys[stg], kin + base, bytes, &sm.mbar[stg]);
            if (LOAD_VALS) bulk_g2s(sm.vals[stg], vin + base, bytes, &sm.mbar[stg]);
        }
    };

    if (tid == 0) {
        mbar_init(&sm.mbar[0], 1);
        mbar_init(&sm.mbar[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        const uint32_t t = atomicAdd(tile_ctr, 1u);
        sm.next_tile[0] = t;
        if (t < ntiles) issue(t, 0);
    }
    __syncthreads();
    uint32_t tile = sm.next_tile[0];
    int stg = 0;
    uint32_t parity = 0; // bit s = parity to wait for on stage s

    while (tile < ntiles) {
        if (tid == 0) { // grab the next tile now and keep its data in flight while this one is processed
            const uint32_t t = atomicAdd(tile_ctr, 1u);
            sm.next_tile[stg ^ 1] = t;
            if (t < ntiles) issue(t, stg ^ 1);
        }
        for (int i = tid; i < WARPS * kRadix; i += THREADS) (&sm.hist[0][0])[i] = 0;
        while (!mbar_try_wait(&sm.mbar[stg], (parity >> stg) & 1u)) {
        }
        parity ^= 1u << stg;
        __syncthreads();
        if (tile * (uint32_t)TILE + (uint32_t)TILE <= n)
            os2_process_tile<THREADS, ITEMS, HAS_VALS, IOTA_VALS, RANK, true, S>(sm, stg, tile, kin, kout, vin, vout, n, shift,
                                                                              digit_base, state);
        else
            os2_process_tile<THREADS, ITEMS, HAS_VALS, IOTA_VALS, RANK, false, S>(sm, stg, tile, kin, kout, vin, vout, n, shift,
                                                                               digit_base, state);
        // every thread has written this stage through the generic proxy (the in-place digit staging); the next bulk copy
        // into it goes through the async proxy: the writers fence before the barrier, then thread 0 may issue the copy
        fence_proxy_async();
        __syncthreads(); // the stage buffers are free again; next_tile[stg^1] was written long ago
        tile = sm.next_tile[stg ^ 1];
        stg ^= 1;
    }
}

// ---- launch plumbing ----------------------------------------------------------------------
struct OnesweepCfg {
    int threads, items;
};
static OnesweepCfg current_cfg() {
    static OnesweepCfg cfg = [] {
        OnesweepCfg c{256, 24}; // tuned on B200: profiles/r01_notes.md
        if (const char *e = getenv("DBT_ONESWEEP_CFG")) { // tuning hook: "<threads>x<items>"
            int t = 0, i = 0;
            if (sscanf(e, "%dx%d", &t, &i) == 2) c = OnesweepCfg{t, i};
        }
        return c;
    }();
    return cfg;
}

template <int THREADS, int ITEMS, typename S = uint32_t>
static int launch_onesweep_t(const uint32_t *kin, uint32_t *kout, const uint32_t *vin, uint32_t *vout, uint32_t n,
                             int shift, const uint32_t *digit_base, void *state_v, uint32_t *ctr, bool has_vals,
                             bool iota, cudaStream_t st) {
    S *state = (S *)state_v;
    using Smem = OnesweepSmem<THREADS, ITEMS>;
    size_t smem = sizeof(Smem);
    uint32_t ntiles = (n + Smem::TILE - 1) / Smem::TILE;
#define DBT_LAUNCH_OS(HV, IO)                                                                                     \
    do {                                                                                                          \
        auto kfn = onesweep_kernel<THREADS, ITEMS, HV, IO, S>;                                                       \
        if (first_use_on_device((const void *)kfn))                                                               \
            DBT_CUDA(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));          \
        kfn<<<ntiles, THREADS, smem, st>>>(kin, kout, vin, vout, n, shift, digit_base, state, ctr);               \
    } while (0)
    if (!has_vals) DBT_LAUNCH_OS(false, false);
    else if (iota) DBT_LAUNCH_OS(true, true);
    else DBT_LAUNCH_OS(true, false);
#undef DBT_LAUNCH_OS
    count_launch();
    DBT_KERNEL_CHECK();
    return 0;
}

struct Os2Cfg {
    int impl;  // 1 = version 1 (one tile per CTA), 2 = persistent pipelined
    int rank;  // 23 (default), or 0 / 1 / 13 for comparison runs
    int ctas_per_sm;
};
static Os2Cfg os2_cfg() {
    static Os2Cfg cfg = [] {
        Os2Cfg c{2, 23, 2}; // persistent pipelined kernel, match(3 low bits)+5 ballots (hand-scheduled, 23), 2 CTAs/SM
        if (const char *e = getenv("DBT_ONESWEEP_IMPL")) c.impl = atoi(e);
        if (const char *e = getenv("DBT_ONESWEEP_RANK")) c.rank = atoi(e);
        if (const char *e = getenv("DBT_ONESWEEP_CTAS")) c.ctas_per_sm = atoi(e);
        return c;
    }();
    return cfg;
}

template <int THREADS, int ITEMS>
static int launch_onesweep2_t(const uint32_t *kin, uint32_t *kout, const uint32_t *vin, uint32_t *vout, uint32_t n,
                              int shift, const uint32_t *digit_base, void *state_v, uint32_t *ctr, bool iota, int rank,
                              cudaStream_t st) {
    uint32_t *state = (uint32_t *)state_v;
    using Smem = Os2Smem<THREADS, ITEMS>;
    size_t smem = sizeof(Smem) + 128;
    uint32_t ntiles = (n + Smem::TILE - 1) / Smem::TILE;
    int grid = (int)std::min<uint32_t>(ntiles, 148u * (uint32_t)os2_cfg().ctas_per_sm);
#define DBT_LAUNCH_OS2(IO, RK)                                                                                   \
    do {                                                                                                         \
        auto kfn = onesweep2_kernel<THREADS, ITEMS, true, IO, RK, uint32_t>;                                               \
        if (first_use_on_device((const void *)kfn))                                                              \
            DBT_CUDA(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));         \
        kfn<<<grid, THREADS, smem, st>>>(kin, kout, vin, vout, n, shift, digit_base, state, ctr);                \
    } while (0)
    if (iota) {
        if (rank == 0) DBT_LAUNCH_OS2(true, 0);
        else if (rank == 1) DBT_LAUNCH_OS2(true, 1);
        else if (rank == 13) DBT_LAUNCH_OS2(true, 13);
        else DBT_LAUNCH_OS2(true, 23);
    } else {
        if (rank == 0) DBT_LAUNCH_OS2(false, 0);
        else if (rank == 1) DBT_LAUNCH_OS2(false, 1);
        else if (rank == 13) DBT_LAUNCH_OS2(false, 13);
        else DBT_LAUNCH_OS2(false, 23);
    }
#undef DBT_LAUNCH_OS2
    count_launch();
    DBT_KERNEL_CHECK();
    return 0;
}

static int tile_items() { // the smallest tile any kernel this build may pick uses (sizes the tile-state array)
    OnesweepCfg c = current_cfg();
    int t = c.threads * c.items;
    return t;
}

// n >= 2^30: 64-bit tile states, one configuration (the tuned default) per kernel version
static int launch_onesweep_wide(const uint32_t *kin, uint32_t *kout, const uint32_t *vin, uint32_t *vout, uint32_t n, int shift,
                                const uint32_t *digit_base, uint64_t *state, uint32_t *ctr, bool has_vals, bool iota, cudaStream_t st) {
    const bool aligned = (((uintptr_t)kin | (uintptr_t)vin) & 15) == 0;
    if (!(has_vals && aligned)) return launch_onesweep_t<256, 24, uint64_t>(kin, kout, vin, vout, n, shift, digit_base, state, ctr, has_vals, iota, st);
    using Smem = Os2Smem<256, 24>;
    const size_t smem = sizeof(Smem) + 128;
    const uint32_t ntiles = (uint32_t)(((uint64_t)n + Smem::TILE - 1) / Smem::TILE);
    const int grid = (int)std::min<uint32_t>(ntiles, 148u * 2u);
#define DBT_LAUNCH_OS2W(IO)                                                                                      \
    do {                                                                                                         \
        auto kfn = onesweep2_kernel<256, 24, true, IO, 23, uint64_t>;                                            \
        if (first_use_on_device((const void *)kfn))                                                              \
            DBT_CUDA(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));         \
        kfn<<<grid, 256, smem, st>>>(kin, kout, vin, vout, n, shift, digit_base, state, ctr);                    \
    } while (0)
    if (iota) DBT_LAUNCH_OS2W(true);
    else DBT_LAUNCH_OS2W(false);
#undef DBT_LAUNCH_OS2W
    count_launch();
    DBT_KERNEL_CHECK();
    return 0;
}

static int launch_onesweep(const uint32_t *kin, uint32_t *kout, const uint32_t *vin, uint32_t *vout, uint32_t n,
                           int shift, const uint32_t *digit_base, uint32_t *state, uint32_t *ctr, bool has_vals,
                           bool iota, cudaStream_t st) {
    OnesweepCfg c = current_cfg();
    const bool aligned = (((uintptr_t)kin | (uintptr_t)vin) & 15) == 0; // cp.async.bulk needs 16-byte aligned sources
    if (os2_cfg().impl == 2 && has_vals && aligned) {
        int rk = os2_cfg().rank;
        if (c.threads == 256 && c.items == 16) // (three CTAs per SM: 0.61 ms per pass against 0.50; kept as the comparison point)
            return launch_onesweep2_t<256, 16>(kin, kout, vin, vout, n, shift, digit_base, state, ctr, iota, rk, st);
        if (c.threads == 256 && c.items == 24)
            return launch_onesweep2_t<256, 24>(kin, kout, vin, vout, n, shift, digit_base, state, ctr, iota, rk, st);
    }
#define DBT_CFG(T, I)                \
    if (c.threads == T && c.items == I) \
        return launch_onesweep_t<T, I>(kin, kout, vin, vout, n, shift, digit_base, state, ctr, has_vals, iota, st);
    DBT_CFG(256, 24)
    DBT_CFG(256, 16)
#undef DBT_CFG
    set_error("unsupported DBT_ONESWEEP_CFG");
    return DBT_ERR_ARG;
}

constexpr uint64_t kMaxSortElems = (1ull << 32) - (1ull << 16); // u32 row ids, and tile arithmetic that stays inside 32 bits

size_t sort_ws_bytes(uint64_t n) {
    uint64_t ntiles = (n + 2048 - 1) / 2048 + 1; // bound for the smallest tile any config uses (256x16 = 4096)
    if (n >= (1ull << 30)) return pad256(((n + 6143) / 6144 + 1) * kRadix * 8) + pad256(4 * kRadix * 4) + 4 * 256 + 4096; // 64-bit states
    return pad256(ntiles * kRadix * 4) + pad256(4 * kRadix * 4) + 4 * 256 + 4096;
}

static DigitPlan plan_digits(uint32_t varying) {
    DigitPlan p{};
    uint32_t m = varying;
    while (m) {
        int s = __builtin_ctz(m);
        p.shift[p.npass++] = s;
        uint64_t window = ((uint64_t)0xFF) << s;
        m &= ~(uint32_t)window;
    }
    return p;
}

// Sort pairs over the varying bits only.  keys/vals: ping buffers (input), *_alt: pong buffers.
// iota_vals: the input values are implicitly the element indices (content of `vals` ignored; the
// first pass synthesises them, saving one read and one write of the row column).
// On return keys/vals point at the buffers holding the result (swapped as needed).
int sort_pairs_masked(uint32_t *&keys, uint32_t *&keys_alt, uint32_t *&vals, uint32_t *&vals_alt, uint64_t n,
                      uint32_t varying_mask, bool iota_vals, Arena &ws, cudaStream_t st, const uint32_t *byte_hist) {
    if (n > kMaxSortElems) {
        set_error("sort_pairs: n must be at most 2^32 - 2^16 per call (row ids are 32 bits)");
        return DBT_ERR_UNSUPPORTED;
    }
    DigitPlan plan = plan_digits(varying_mask);
    if (n == 0 || plan.npass == 0) {
        if (iota_vals && n) DBT_TRY(iota_u32(vals, n, st)); // nothing to sort but the caller expects a row list
        return 0;
    }
    const bool wide = n >= (1ull << 30); // the 30-bit values of the 32-bit tile states no longer hold a position
    size_t m0 = ws.mark();
    const uint32_t tile = wide ? 256u * 24u : (uint32_t)tile_items();
    uint32_t ntiles = (uint32_t)((n + tile - 1) / tile);
    uint32_t *ghist = ws.take<uint32_t>(4 * kRadix);
    uint32_t *ctr = ws.take<uint32_t>(64);
    const size_t state_bytes = (size_t)ntiles * kRadix * (wide ? 8 : 4);
    void *state = ws.take<unsigned char>(state_bytes);
    if (!ghist || !ctr || !state) {
        set_error("sort_pairs: workspace too small");
        return DBT_ERR_WORKSPACE;
    }
    bool byte_aligned = byte_hist != nullptr;
    for (int p = 0; p < plan.npass; ++p) byte_aligned = byte_aligned && (plan.shift[p] % 8 == 0);
    {
        StageScope sc(ST_HIST, st);
        if (byte_aligned) { // the extraction already counted the key bytes: pick the planned digit positions, no key read
            for (int p = 0; p < plan.npass; ++p)
                DBT_CUDA(cudaMemcpyAsync(ghist + p * kRadix, byte_hist + (plan.shift[p] / 8) * kRadix, kRadix * 4,
                                         cudaMemcpyDeviceToDevice, st));
        } else {
            DBT_CUDA(cudaMemsetAsync(ghist, 0, 4 * kRadix * 4, st));
            int grid = (int)std::min<uint64_t>((n / 4 + 511) / 512 + 1, 148 * 4);
            hist_kernel<<<grid, 512, 0, st>>>(keys, n, plan, ghist);
            count_launch();
        }
        hist_scan_kernel<<<plan.npass, kRadix, 0, st>>>(ghist);
        count_launch();
        DBT_KERNEL_CHECK();
    }
    for (int p = 0; p < plan.npass; ++p) {
        StageScope sc(ST_ONESWEEP, st);
        DBT_CUDA(cudaMemsetAsync(state, 0, state_bytes, st));
        DBT_CUDA(cudaMemsetAsync(ctr, 0, 4, st));
        if (wide)
            DBT_TRY(launch_onesweep_wide(keys, keys_alt, vals, vals_alt, (uint32_t)n, plan.shift[p], ghist + p * kRadix,
                                         (uint64_t *)state, ctr, true, iota_vals && p == 0, st));
        else
            DBT_TRY(launch_onesweep(keys, keys_alt, vals, vals_alt, (uint32_t)n, plan.shift[p], ghist + p * kRadix,
                                    (uint32_t *)state, ctr, true, iota_vals && p == 0, st));
        std::swap(keys, keys_alt);
        std::swap(vals, vals_alt);
    }
    ws.release(m0);
    return 0;
}

__global__ void iota_kernel(uint32_t *d, uint64_t n) {
    uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) d[i] = (uint32_t)i;
}
int iota_u32(uint32_t *d, uint64_t n, cudaStream_t st) {
    if (!n) return 0;
    int grid = (int)std::min<uint64_t>((n + 255) / 256, 148 * 16);
    iota_kernel<<<grid, 256, 0, st>>>(d, n);
    count_launch();
    DBT_KERNEL_CHECK();
    return 0;
}

// curkey[i] = src[perm[i] * stride + word]  (perm == nullptr => identity)
__global__ void __launch_bounds__(256)
gather_word_kernel(const uint32_t *__restrict__ src, uint32_t stride, uint32_t word, const uint32_t *__restrict__ perm,
                   uint32_t *__restrict__ out, uint64_t n) {
    uint64_t gstride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gstride) {
        uint64_t r = perm ? perm[i] : i;
        out[i] = perm ? ld_sparse(src + r * stride + word) : src[r * stride + word];
    }
}
int gather_word(const uint32_t *d_src, uint32_t stride, uint32_t word, const uint32_t *d_perm, uint32_t *d_out,
                uint64_t n, cudaStream_t st) {
    if (!n) return 0;
    int grid = (int)std::min<uint64_t>((n + 255) / 256, 148 * 16);
    gather_word_kernel<<<grid, 256, 0, st>>>(d_src, stride, word, d_perm, d_out, n);
    count_launch();
    DBT_KERNEL_CHECK();
    return 0;
}

} // namespace dbt

using namespace dbt;

extern "C" size_t dbt_sort_pairs_ws_bytes(uint64_t n) { return sort_ws_bytes(n) + 256 + 4 * 256 * 4 + 512; } // + OR/AND words, byte histograms

extern "C" int dbt_sort_pairs_u32(uint32_t *d_keys, uint32_t *d_keys_alt, uint32_t *d_vals, uint32_t *d_vals_alt,
                                  uint64_t n, int begin_bit, int end_bit, void *d_ws, size_t ws_bytes, void *stream,
                                  int *result_in_alt) {
    if (!d_keys || !d_keys_alt || !d_vals || !d_vals_alt || begin_bit < 0 || end_bit > 32 || begin_bit > end_bit) {
        set_error("dbt_sort_pairs_u32: bad arguments");
        return DBT_ERR_ARG;
    }
    cudaStream_t st = (cudaStream_t)stream;
    Arena ws(d_ws, ws_bytes);
    uint32_t *oa = ws.take<uint32_t>(64);
    uint32_t *byte_hist = ws.take<uint32_t>(4 * kRadix);
    if (!oa || !byte_hist) {
        set_error("dbt_sort_pairs_u32: workspace too small");
        return DBT_ERR_WORKSPACE;
    }
    uint32_t h[2] = {0, 0};
    const bool fused = (((uintptr_t)d_keys) & 15) == 0; // one read for OR/AND and the byte histograms
    if (n) {
        StageScope sc(ST_MISC, st);
        if (fused) {
            init_or_and_kernel<<<1, 1, 0, st>>>(oa);
            DBT_CUDA(cudaMemsetAsync(byte_hist, 0, 4 * kRadix * 4, st));
            const int grid = (int)std::min<uint64_t>((n / 4 + 511) / 512 + 1, 148 * 4);
            or_and_hist_kernel<<<grid, 512, 0, st>>>(d_keys, n, oa, byte_hist);
            count_launch(2);
            DBT_KERNEL_CHECK();
        } else {
            DBT_TRY(or_and_reduce(d_keys, n, oa, st));
        }
        DBT_CUDA(cudaMemcpyAsync(h, oa, 8, cudaMemcpyDeviceToHost, st));
        DBT_CUDA(cudaStreamSynchronize(st));
    }
    uint32_t range = (end_bit - begin_bit >= 32) ? 0xFFFFFFFFu
                                                 : (uint32_t)((((uint64_t)1 << (end_bit - begin_bit)) - 1) << begin_bit);
    uint32_t varying = (h[0] ^ h[1]) & range;
    uint32_t *k = d_keys, *ka = d_keys_alt, *v = d_vals, *va = d_vals_alt;
    DBT_TRY(sort_pairs_masked(k, ka, v, va, n, varying, false, ws, st, (n && fused) ? byte_hist : nullptr));
    DBT_CUDA(cudaStreamSynchronize(st));
    stage_resolve();
    if (result_in_alt) *result_in_alt = (k == d_keys_alt) ? 1 : 0;
    return 0;
}
