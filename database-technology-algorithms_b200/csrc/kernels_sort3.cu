// kernels_sort3.cu -- version 3 of the onesweep pass (see kernels_sort.cu for the sort driver and versions 1/2).
#include "sort_common.cuh"
#include <algorithm>
#include <cstdlib>

namespace dbt {

// ---------------------------------------------------------------------------------------------
// Onesweep pass, version 3: no warp-synchronous ranking at all.
//
// Versions 1/2 rank an 8-bit digit with one match.any / five ballots, a leader update of a warp counter and a
// broadcast shuffle PER ITEM -- ~52 of the ~90 warp instructions spent per 32 keys, all of them serialised inside the
// warp (ncu: issue slots 44 % busy, no pipe saturated).  Here a tile is ordered by its digit with TWO local 4-bit
// counting passes entirely in shared memory (an LSD sort inside the tile: low nibble, then high nibble; both stable, so
// the pair is a stable 8-bit partition):
//   * every thread owns ITEMS consecutive elements (blocked arrangement) and 16 private 16-bit counters packed in 8
//     words of shared memory: counting an element is LDS / IADD / STS on a private word, the previous count is the
//     element's rank among the thread's own elements with that nibble -- no atomics, no shuffles, no votes;
//   * one block-wide exclusive scan of the 8 x 256 packed words in (nibble, thread) order (each warp rakes one word
//     lane; the low halves' total is carried into the high halves) gives every thread the tile position of its first
//     element of each nibble;
//   * the elements move to those positions as 64-bit (key, row) pairs (one STS.64 each).
// After the second pass the tile is in digit order; run boundaries found while the write-out operands are loaded give
// the 256 digit counts (no histogram atomics, nothing data dependent), the look-back chain and the coalesced write-out
// are those of version 2.  The next tile's bulk copy is issued as soon as the write-out operands are in registers.
// ---------------------------------------------------------------------------------------------
template <int ITEMS>
struct Os3Smem {
    static constexpr int THREADS = 256;
    static constexpr int TILE = THREADS * ITEMS;
    alignas(128) uint2 pairs[TILE]; // arrives as keys[TILE] | rows[TILE] (bulk copies), then (key, row) pairs in place
    alignas(16) uint32_t cnt[16][THREADS]; // cnt[b][t]: thread t's private counter of nibble b
    uint32_t start[kRadix], end[kRadix];   // tile-local digit runs [start, end)
    uint32_t goff[kRadix];
    uint32_t wtot[8];
    alignas(8) uint64_t mbar;
    uint32_t next_tile;
};

// Stable counting pass on the 4-bit digit at `sh4` for blocked elements, then the move: element j goes to its tile
// position as a (key, row) pair.  Local ranks are kept four to a register.  Leaves the thread's counters zeroed for
// the next pass.  Three CTA barriers; the first one also ends every read of the stage by other threads.
template <int ITEMS>
__device__ __forceinline__ void os3_pass(Os3Smem<ITEMS> &sm, const uint32_t (&key)[ITEMS], const uint32_t (&val)[ITEMS], int sh4) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    uint32_t *mycnt = &sm.cnt[0][tid];
    uint32_t lr[(ITEMS + 3) / 4];
#pragma unroll
    for (int j = 0; j < (ITEMS + 3) / 4; ++j) lr[j] = 0;
    // ptxas would hoist the nibble / address arithmetic of ALL items above this serial chain of private read-modify-writes
    // and spill; a (never set) bit of a counter read in the previous group of four enters the shift amount of the next
    // group, which bounds the look-ahead
    uint32_t shg = sh4;
#pragma unroll
    for (int j = 0; j < ITEMS; ++j) {
        const uint32_t nib = (key[j] >> shg) & 15u;
        uint32_t *c = mycnt + nib * 256;
        const uint32_t w = *c; // < ITEMS
        lr[j >> 2] += w << ((j & 3) * 8);
        *c = w + 1;
        if ((j & 3) == 3) shg = sh4 + (w >> 31);
    }
    __syncthreads();
    { // exclusive scan of the 4096 counters in (nibble, thread) order: thread i rakes words [16 i, 16 i + 16)
        uint4 *row = reinterpret_cast<uint4 *>(&sm.cnt[0][0]) + 4 * tid;
        uint4 q[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) q[k] = row[k];
        uint32_t run = 0;
#pragma unroll
        for (int k = 0; k < 4; ++k) { // q := exclusive prefix inside the thread
            uint32_t t;
            t = q[k].x, q[k].x = run, run += t;
            t = q[k].y, q[k].y = run, run += t;
            t = q[k].z, q[k].z = run, run += t;
            t = q[k].w, q[k].w = run, run += t;
        }
        uint32_t incl = run;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, incl, o);
            if (lane >= o) incl += t;
        }
        if (lane == 31) sm.wtot[warp] = incl;
        __syncthreads();
        uint32_t e = incl - run;
#pragma unroll
        for (int w = 0; w < 7; ++w)
            if (w < warp) e += sm.wtot[w];
#pragma unroll
        for (int k = 0; k < 4; ++k) row[k] = make_uint4(q[k].x + e, q[k].y + e, q[k].z + e, q[k].w + e);
    }
    __syncthreads();
    shg = sh4;
#pragma unroll
    for (int j = 0; j < ITEMS; ++j) {
        const uint32_t nib = (key[j] >> shg) & 15u;
        const uint32_t w = mycnt[nib * 256];
        const uint32_t pos = w + ((lr[j >> 2] >> ((j & 3) * 8)) & 0xFFu);
        sm.pairs[pos] = make_uint2(key[j], val[j]);
        if ((j & 3) == 3) shg = sh4 + (w >> 31);
    }
#pragma unroll
    for (int b = 0; b < 16; ++b) mycnt[b * 256] = 0;
}

#ifndef DBT_OS3_WINDOW
#define DBT_OS3_WINDOW 4
#endif

// look-back for one digit column: sleep on the immediate predecessor (tiles are ranked at about the same time, so its
// aggregate is usually still on its way: polling it from 256 threads would only burn issue slots), then walk back
template <int WINDOW>
__device__ __forceinline__ uint32_t os3_lookback(const uint32_t *state, uint32_t tile, uint32_t col) {
    const uint32_t *pstate = state + (size_t)(tile - 1) * kRadix + col;
    uint32_t s0 = ld_volatile(pstate);
    while ((s0 >> 30) == 0) {
        __nanosleep(40);
        s0 = ld_volatile(pstate);
    }
    uint32_t excl = s0 & kValMask;
    if (s0 & kFlagInc) return excl;
    int p = (int)tile - 2;
    while (p >= 0) {
        uint32_t s[WINDOW];
#pragma unroll
        for (int q = 0; q < WINDOW; ++q) s[q] = (p - q >= 0) ? ld_volatile(pstate - (size_t)(q + 1) * kRadix) : kFlagInc;
        int q = 0;
        bool done = false;
#pragma unroll
        for (; q < WINDOW; ++q) {
            if ((s[q] >> 30) == 0) break;
            excl += s[q] & kValMask;
            if (s[q] & kFlagInc) {
                done = true;
                break;
            }
        }
        if (done) break;
        p -= q;
        pstate -= (size_t)q * kRadix;
    }
    return excl;
}

template <int ITEMS, bool IOTA_VALS, bool TWO_NIBBLES, int MINB, int WINDOW>
__global__ void __launch_bounds__(256, MINB)
onesweep3_kernel(const uint32_t *__restrict__ kin, uint32_t *__restrict__ kout, const uint32_t *__restrict__ vin,
                 uint32_t *__restrict__ vout, uint32_t n, int shift, const uint32_t *__restrict__ digit_base,
                 uint32_t *state /*[ntiles][256], zeroed*/, uint32_t *tile_ctr /*zeroed*/) {
    using Smem = Os3Smem<ITEMS>;
    constexpr int THREADS = Smem::THREADS;
    constexpr int TILE = Smem::TILE;
    static_assert(ITEMS % 2 == 0, "blocked loads are 8- or 16-byte vectors");
    extern __shared__ __align__(128) unsigned char smem_raw[];
    Smem &sm = *reinterpret_cast<Smem *>(smem_raw);
    const int tid = threadIdx.x;
    const uint32_t ntiles = (n + TILE - 1) / TILE;
    uint32_t *skeys = reinterpret_cast<uint32_t *>(sm.pairs);
    uint32_t *svals = skeys + TILE;

    auto issue = [&](uint32_t t) { // thread 0: bulk copies of a FULL tile (the last, partial tile is loaded directly)
        const uint32_t base = t * (uint32_t)TILE;
        if (t >= ntiles || base + (uint32_t)TILE > n) return;
        mbar_expect_tx(&sm.mbar, IOTA_VALS ? TILE * 4u : TILE * 8u);
        bulk_g2s(skeys, kin + base, TILE * 4u, &sm.mbar);
        if (!IOTA_VALS) bulk_g2s(svals, vin + base, TILE * 4u, &sm.mbar);
    };

    if (tid == 0) {
        mbar_init(&sm.mbar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        const uint32_t t = atomicAdd(tile_ctr, 1u);
        sm.next_tile = t;
        issue(t);
    }
#pragma unroll
    for (int b = 0; b < 16; ++b) sm.cnt[b][tid] = 0;
    __syncthreads();
    uint32_t tile = sm.next_tile;
    uint32_t parity = 0;

    while (tile < ntiles) {
        const uint32_t base = tile * (uint32_t)TILE;
        const bool full = base + (uint32_t)TILE <= n;
        const uint32_t nvalid = full ? (uint32_t)TILE : n - base;
        sm.start[tid] = 0;
        sm.end[tid] = 0;

        // ---- blocked load: thread t owns elements [t * ITEMS, (t + 1) * ITEMS)
        uint32_t key[ITEMS], val[ITEMS];
        if (full) {
            while (!mbar_try_wait(&sm.mbar, parity)) {
            }
            parity ^= 1u;
            if (ITEMS % 4 == 0) {
#pragma unroll
                for (int j = 0; j < ITEMS / 4 * 4; j += 4) {
                    const uint4 k4 = *reinterpret_cast<const uint4 *>(skeys + tid * ITEMS + j);
                    key[j] = k4.x, key[j + 1] = k4.y, key[j + 2] = k4.z, key[j + 3] = k4.w;
                    if (!IOTA_VALS) {
                        const uint4 v4 = *reinterpret_cast<const uint4 *>(svals + tid * ITEMS + j);
                        val[j] = v4.x, val[j + 1] = v4.y, val[j + 2] = v4.z, val[j + 3] = v4.w;
                    }
                }
            } else { // ITEMS = 2 (mod 4): 8-byte loads are the conflict-free ones
#pragma unroll
                for (int j = 0; j < ITEMS; j += 2) {
                    const uint2 k2 = *reinterpret_cast<const uint2 *>(skeys + tid * ITEMS + j);
                    key[j] = k2.x, key[j + 1] = k2.y;
                    if (!IOTA_VALS) {
                        const uint2 v2 = *reinterpret_cast<const uint2 *>(svals + tid * ITEMS + j);
                        val[j] = v2.x, val[j + 1] = v2.y;
                    }
                }
            }
        } else {
#pragma unroll
            for (int j = 0; j < ITEMS; ++j) {
                const uint32_t li = tid * ITEMS + j;
                key[j] = li < nvalid ? kin[base + li] : 0xFFFFFFFFu; // padding sorts behind every real element (stable)
                if (!IOTA_VALS) val[j] = li < nvalid ? vin[base + li] : 0u;
            }
        }
        if (IOTA_VALS) {
#pragma unroll
            for (int j = 0; j < ITEMS; ++j) val[j] = base + tid * ITEMS + j;
        }

        // ---- low nibble: count, scan, move (the first barrier inside os3_pass also ends every read of the SoA stage)
        os3_pass<ITEMS>(sm, key, val, shift);
        if (TWO_NIBBLES) {
            __syncthreads();
#pragma unroll
            for (int j = 0; j < ITEMS; j += 2) {
                const uint4 p2 = *reinterpret_cast<const uint4 *>(&sm.pairs[tid * ITEMS + j]);
                key[j] = p2.x, val[j] = p2.y, key[j + 1] = p2.z, val[j + 1] = p2.w;
            }
            os3_pass<ITEMS>(sm, key, val, shift + 4);
        }
        __syncthreads();

        // ---- write-out operands (striped) and the digit runs of the tile; the next tile's ticket is taken now so that
        // its round trip hides behind this phase (taken any earlier, tickets and ranking order drift apart and the
        // look-back chain waits on tiles that are still parked)
        if (tid == 0) sm.next_tile = atomicAdd(tile_ctr, 1u);
#pragma unroll
        for (int j = 0; j < ITEMS; ++j) {
            const uint32_t i = tid + j * THREADS;
            const uint2 p = sm.pairs[i];
            key[j] = p.x, val[j] = p.y;
            const uint32_t kprev = i ? skeys[2 * i - 2] : ~p.x;
            if ((((p.x ^ kprev) >> shift) & 0xFFu) != 0 && i < nvalid) { // rare: one run per digit
                const uint32_t d = (p.x >> shift) & 0xFFu;
                sm.start[d] = i;
                if (i) sm.end[(kprev >> shift) & 0xFFu] = i;
            }
        }
        if (tid == 0) { // the run that ends the tile
            const uint32_t klast = skeys[2 * (nvalid - 1)];
            sm.end[(klast >> shift) & 0xFFu] = nvalid;
        }
        fence_proxy_async(); // generic-proxy writes to the stage are ordered before the next bulk copy into it
        __syncthreads();
        if (tid == 0) issue(sm.next_tile); // the stage is free: fetch the next tile while this one is resolved and written
        {
            const uint32_t st_d = sm.start[tid], count_d = sm.end[tid] - st_d;
            st_volatile(&state[(size_t)tile * kRadix + tid], (tile == 0 ? kFlagInc : kFlagAgg) | count_d);
            uint32_t excl_prefix = 0;
            if (tile > 0) {
                excl_prefix = os3_lookback<WINDOW>(state, tile, tid);
                st_volatile(&state[(size_t)tile * kRadix + tid], kFlagInc | (excl_prefix + count_d));
            }
            sm.goff[tid] = digit_base[tid] + excl_prefix - st_d;
        }
        __syncthreads();
        if (full) {
#pragma unroll
            for (int j = 0; j < ITEMS; ++j) {
                const uint32_t dst = sm.goff[(key[j] >> shift) & 0xFFu] + tid + j * THREADS;
                kout[dst] = key[j];
                vout[dst] = val[j];
            }
        } else {
#pragma unroll
            for (int j = 0; j < ITEMS; ++j) {
                const uint32_t i = tid + j * THREADS;
                if (i < nvalid) {
                    const uint32_t dst = sm.goff[(key[j] >> shift) & 0xFFu] + i;
                    kout[dst] = key[j];
                    vout[dst] = val[j];
                }
            }
        }
        const uint32_t nt = sm.next_tile;
        __syncthreads(); // goff / start / end / next_tile are reused by the next tile
        tile = nt;
    }
}

static int os3_items() {
    static int items = [] {
        int i = 22;
        if (const char *e = getenv("DBT_OS3_ITEMS")) i = atoi(e);
        return i;
    }();
    return items;
}
static int os3_ctas() { // CTAs per SM (0 = as many as fit)
    static int v = [] {
        const char *e = getenv("DBT_OS3_CTAS");
        return e ? atoi(e) : 0;
    }();
    return v;
}

template <int ITEMS, int MINB, int WINDOW = DBT_OS3_WINDOW>
static int launch_onesweep3_t(const uint32_t *kin, uint32_t *kout, const uint32_t *vin, uint32_t *vout, uint32_t n,
                              int shift, const uint32_t *digit_base, uint32_t *state, uint32_t *ctr, bool iota,
                              bool two_nibbles, cudaStream_t st) {
    using Smem = Os3Smem<ITEMS>;
    const size_t smem = sizeof(Smem) + 128;
    const uint32_t ntiles = (n + Smem::TILE - 1) / Smem::TILE;
    static int resident[16] = {0}; // CTAs per SM from the occupancy query, per device
    int dev = 0;
    DBT_CUDA(cudaGetDevice(&dev));
#define DBT_LAUNCH_OS3(IO, TN)                                                                                      \
    do {                                                                                                            \
        auto kfn = onesweep3_kernel<ITEMS, IO, TN, MINB, WINDOW>;                                                           \
        if (first_use_on_device((const void *)kfn))                                                                 \
            DBT_CUDA(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));            \
        if (!resident[dev & 15]) {                                                                                  \
            int r = 0;                                                                                              \
            DBT_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&r, kfn, 256, smem));                            \
            resident[dev & 15] = r > 0 ? r : 1;                                                                     \
        }                                                                                                           \
        int per_sm = resident[dev & 15];                                                                            \
        if (os3_ctas() > 0 && os3_ctas() < per_sm) per_sm = os3_ctas();                                             \
        const int grid = (int)std::min<uint32_t>(ntiles, 148u * (uint32_t)per_sm);                                  \
        kfn<<<grid, 256, smem, st>>>(kin, kout, vin, vout, n, shift, digit_base, state, ctr);                       \
    } while (0)
    if (iota) {
        if (two_nibbles) DBT_LAUNCH_OS3(true, true);
        else DBT_LAUNCH_OS3(true, false);
    } else {
        if (two_nibbles) DBT_LAUNCH_OS3(false, true);
        else DBT_LAUNCH_OS3(false, false);
    }
#undef DBT_LAUNCH_OS3
    count_launch();
    DBT_KERNEL_CHECK();
    return 0;
}


int onesweep3_tile_items() { return 256 * os3_items(); }

int launch_onesweep3(const uint32_t *kin, uint32_t *kout, const uint32_t *vin, uint32_t *vout, uint32_t n, int shift,
                     const uint32_t *digit_base, uint32_t *state, uint32_t *ctr, bool iota, bool two_nibbles,
                     cudaStream_t st) {
    static int win = [] { const char *e = getenv("DBT_OS3_WIN"); return e ? atoi(e) : 0; }(); // experiment hook
    if (win == 8 && os3_items() == 22) return launch_onesweep3_t<22, 3, 8>(kin, kout, vin, vout, n, shift, digit_base, state, ctr, iota, two_nibbles, st);
    if (win == 16 && os3_items() == 22) return launch_onesweep3_t<22, 3, 16>(kin, kout, vin, vout, n, shift, digit_base, state, ctr, iota, two_nibbles, st);
    if (win == 2 && os3_items() == 22) return launch_onesweep3_t<22, 3, 2>(kin, kout, vin, vout, n, shift, digit_base, state, ctr, iota, two_nibbles, st);
    switch (os3_items()) {
    case 18: return launch_onesweep3_t<18, 4>(kin, kout, vin, vout, n, shift, digit_base, state, ctr, iota, two_nibbles, st);
    case 20: return launch_onesweep3_t<20, 3>(kin, kout, vin, vout, n, shift, digit_base, state, ctr, iota, two_nibbles, st);
    case 22: return launch_onesweep3_t<22, 3>(kin, kout, vin, vout, n, shift, digit_base, state, ctr, iota, two_nibbles, st);
    case 26: return launch_onesweep3_t<26, 3>(kin, kout, vin, vout, n, shift, digit_base, state, ctr, iota, two_nibbles, st);
    case 28: return launch_onesweep3_t<28, 2>(kin, kout, vin, vout, n, shift, digit_base, state, ctr, iota, two_nibbles, st);
    default: set_error("unsupported DBT_OS3_ITEMS"); return DBT_ERR_ARG;
    }
}

} // namespace dbt
