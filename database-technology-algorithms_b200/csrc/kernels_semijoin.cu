// kernels_semijoin.cu -- HashJoin's probe phase as ONE streaming pass over the S image (fields '0' and '1').
//
// The reference's probe (DatabaseProject.cpp:561-640) is itself a single pass over S: read a chunk of blocks, look
// every record's key up in the table built from R, append the matching records to the output block, write it when it
// is full.  The first CUDA version split that into extraction (a full read of S for 8 bytes per row), a probe over
// the key column, a compaction and a random gather of the matching records (a second, scattered read of S).  Here S is
// read once, sequentially:
//
//   persistent CTAs take tiles of two consecutive S blocks in file order (a ticket per tile, taken only when the CTA
//   is ready for it) and bring them into shared memory with one cp.async.bulk (mbarrier completion); seven resident
//   CTAs per SM keep ~200 KB per SM in flight, which is what hides the DRAM latency;
//   every live row tests its key against the direct-address bitmap of keys(R) (L2-resident for reference-like key
//   ranges); the tile's match count goes into a decoupled look-back chain (64-bit tile states, the whole CTA looks
//   back 512 predecessors per round trip) that yields the output row of the block's first match;
//   the matching 140-byte records are copied from shared memory straight to their final place in the packed output
//   image (consecutive matches are contiguous there, so the 4-byte stores of a warp coalesce), with the CANON block
//   headers written by whoever emits a block's first row.
//
// DRAM traffic: 140 B read per S row + 140 B written per match -- the algorithmic minimum for an operator that must
// look at every S record and emit the matching ones (the split version moved ~141 + 8 + 8 + 4 + 267 s + 140 s).
#include "dbt_internal.cuh"
#include <algorithm>
#include <cstdlib>

namespace dbt {

constexpr int kSjThreads = 256;
constexpr int kSjTileBlocks = 2;                  // S blocks per tile: one row per thread (200 of the 256 threads)
constexpr int kSjTileRows = kSjTileBlocks * kRpb;
constexpr int kSjLbPerThread = 2;                 // look-back window = 2 x 256 predecessors per round trip
constexpr int kSjLbSegs = kSjLbPerThread * (kSjThreads / 32);
constexpr uint64_t kSjAgg = 1ull << 62, kSjInc = 2ull << 62, kSjMask = (1ull << 62) - 1;

__device__ __forceinline__ uint32_t sj_smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t sj_ld(const uint64_t *p) {
    uint64_t v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void sj_st(uint64_t *p, uint64_t v) {
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

struct SjSmem {
    alignas(128) uint32_t stage[kSjTileBlocks * kBlockWords];
    alignas(8) uint64_t mbar;
    uint64_t lb_sum[kSjLbSegs];
    uint32_t lb_inc[kSjLbSegs], lb_ok[kSjLbSegs];
    uint32_t tile;
    uint32_t src[kSjTileRows]; // word offset (inside the stage) of the k-th match's record
    uint32_t dst[kSjTileRows]; // word offset of the k-th match's output record, relative to the first one's
    uint32_t wcnt[kSjThreads / 32];
};

// Exclusive prefix (output rows before this tile) by decoupled look-back; the whole CTA takes part: thread t examines
// predecessors t and t + 256 of the current window, so one round trip to L2 covers 512 tiles.  Uniform result.
// A tile takes its ticket only when it is about to be processed (no tile is claimed ahead of time), so the aggregates
// of the predecessors are published within a DRAM latency of their tickets and the chain never waits on a parked tile.
__device__ __forceinline__ uint64_t sj_lookback(SjSmem &sm, uint64_t *state, uint32_t tile, uint32_t m, uint32_t *err) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tile == 0) {
        if (tid == 0) sj_st(&state[0], kSjInc | (uint64_t)m);
        return 0;
    }
    if (tid == 0) sj_st(&state[tile], kSjAgg | (uint64_t)m);
    uint64_t excl = 0;
    int64_t p = (int64_t)tile - 1;
    uint32_t spins = 0;
    while (true) {
#pragma unroll
        for (int h = 0; h < kSjLbPerThread; ++h) {
            const int64_t idx = p - (int64_t)(h * kSjThreads + tid);
            const uint64_t sv = (idx >= 0) ? sj_ld(&state[idx]) : kSjInc;
            const uint32_t ready = __ballot_sync(0xFFFFFFFFu, (sv >> 62) != 0);
            const uint32_t inc = __ballot_sync(0xFFFFFFFFu, (sv & kSjInc) != 0);
            const int first_inc = inc ? (__ffs(inc) - 1) : 32;
            const uint32_t need = (first_inc >= 31) ? 0xFFFFFFFFu : ((2u << first_inc) - 1u);
            uint64_t v = (lane <= first_inc) ? (sv & kSjMask) : 0ull;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, o);
            if (lane == 0) {
                const int s = h * (kSjThreads / 32) + warp; // segments in order of growing distance
                sm.lb_sum[s] = v;
                sm.lb_inc[s] = first_inc < 32;
                sm.lb_ok[s] = (ready & need) == need;
            }
        }
        __syncthreads();
        uint64_t acc = 0;
        bool found = false, fail = false;
        int s = 0;
        for (; s < kSjLbSegs; ++s) {
            if (!sm.lb_ok[s]) {
                fail = true;
                break;
            }
            acc += sm.lb_sum[s];
            if (sm.lb_inc[s]) {
                found = true;
                break;
            }
        }
        __syncthreads(); // the scratch is rewritten by the next round
        excl += acc;
        if (found) break;
        if (fail) { // segments before s were complete and held no inclusive prefix: keep them, poll again from s
            p -= 32 * s;
            if (++spins > (1u << 22)) { // a predecessor never published: report instead of hanging the device
                if (tid == 0) atomicExch(err, 1u);
                break;
            }
            __nanosleep(100);
            continue;
        }
        p -= kSjLbPerThread * kSjThreads;
    }
    if (tid == 0) sj_st(&state[tile], kSjInc | (excl + (uint64_t)m));
    return excl;
}

template <int FIELD> // 0 = recid, 1 = num
__global__ void __launch_bounds__(kSjThreads)
semijoin_stream_kernel(const uint32_t *__restrict__ img, uint32_t nblocks, const uint32_t *__restrict__ bm, uint32_t base,
                       uint32_t span, uint32_t *__restrict__ out, uint64_t cap_rows, uint64_t *state /*[ntiles] zeroed*/,
                       uint32_t *ctr /*[0] ticket counter, [1] error flag; zeroed*/, unsigned long long *total_out) {
    __shared__ SjSmem sm;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t lt = (1u << lane) - 1u;
    const uint32_t ntiles = (nblocks + kSjTileBlocks - 1) / kSjTileBlocks;
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(sj_smem_u32(&sm.mbar)), "r"(1));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    uint32_t parity = 0;
    const uint32_t my_blk = (uint32_t)tid / kRpb, my_e = (uint32_t)tid - my_blk * kRpb; // the row this thread tests
    while (true) {
        __syncthreads(); // everyone is done with the stage, src/dst and sm.tile of the previous tile
        if (tid == 0) {  // take a ticket and start the bulk copy of the tile's blocks (consecutive in the image)
            const uint32_t t = atomicAdd(&ctr[0], 1u);
            sm.tile = t;
            if (t < ntiles) {
                const uint32_t b0 = t * kSjTileBlocks;
                const uint32_t bytes = min((uint32_t)kSjTileBlocks, nblocks - b0) * (uint32_t)DBT_BLOCK_BYTES;
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(sj_smem_u32(&sm.mbar)), "r"(bytes) : "memory");
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                                 sj_smem_u32(sm.stage)),
                             "l"(img + (uint64_t)b0 * kBlockWords), "r"(bytes), "r"(sj_smem_u32(&sm.mbar))
                             : "memory");
            }
        }
        __syncthreads();
        const uint32_t tile = sm.tile;
        if (tile >= ntiles) break;
        uint32_t ok = 0;
        while (!ok)
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                         : "=r"(ok)
                         : "r"(sj_smem_u32(&sm.mbar)), "r"(parity)
                         : "memory");
        parity ^= 1u;
        const uint32_t nb_tile = min((uint32_t)kSjTileBlocks, nblocks - tile * kSjTileBlocks);
        bool match = false;
        uint32_t my_word = 0;
        if (my_blk < nb_tile) {
            const uint32_t *blk = sm.stage + my_blk * kBlockWords;
            const uint32_t nres = min(blk[1], kRpb);
            if (my_e < nres) {
                my_word = my_blk * kBlockWords + kEntriesWord + my_e * kRecWords;
                const uint32_t key = sm.stage[my_word + (FIELD == 0 ? 0 : 1)];
                const uint32_t v = key - base;
                match = (v <= span) && ((__ldg(bm + (v >> 5)) >> (v & 31)) & 1u);
            }
        }
        const uint32_t ballot = __ballot_sync(0xFFFFFFFFu, match);
        if (lane == 0) sm.wcnt[warp] = __popc(ballot);
        __syncthreads();
        uint32_t wpre = 0, m = 0;
#pragma unroll
        for (int w = 0; w < kSjThreads / 32; ++w) {
            const uint32_t c = sm.wcnt[w];
            if (w < warp) wpre += c;
            m += c;
        }
        if (match) sm.src[wpre + __popc(ballot & lt)] = my_word;
        const uint64_t excl = sj_lookback(sm, state, tile, m, &ctr[1]);
        const uint64_t off0 = slot_word(excl);
        if (tid < (int)m) {
            const uint64_t g = excl + (uint64_t)tid;
            const uint64_t b = g / kRpb;
            const uint32_t e = (uint32_t)(g - b * kRpb);
            sm.dst[tid] = (uint32_t)(b * kBlockWords + kEntriesWord + (uint64_t)e * kRecWords - off0);
            if (e == 0 && g < cap_rows) { // first row of an output block: its header (the last block's is fixed afterwards)
                uint32_t *ob = out + b * kBlockWords;
                ob[0] = (uint32_t)b;
                ob[1] = kRpb;
                ob[kTrailerWord] = 1;
                ob[kTrailerWord + 1] = kRpb;
            }
        }
        __syncthreads();
        const uint32_t nwords = m * kRecWords;
        for (uint32_t idx = tid; idx < nwords; idx += kSjThreads) {
            const uint32_t rec = idx / kRecWords, w = idx - rec * kRecWords;
            if (excl + rec < cap_rows) out[off0 + sm.dst[rec] + w] = sm.stage[sm.src[rec] + w];
        }
        if (tid == 0 && tile + 1 == ntiles) *total_out = excl + m;
    }
}

// the partly filled last output block: true header, unused slots zero (CANON)
__global__ void __launch_bounds__(256)
semijoin_finish_kernel(uint32_t *__restrict__ out, const unsigned long long *__restrict__ total, uint64_t cap_rows) {
    const uint64_t T = *total;
    if (T == 0 || T > cap_rows) return;
    const uint32_t cnt = (uint32_t)(T % kRpb);
    if (cnt == 0) return;
    uint32_t *blk = out + (T / kRpb) * kBlockWords;
    if (threadIdx.x == 0) {
        blk[0] = (uint32_t)(T / kRpb);
        blk[1] = cnt;
        blk[kTrailerWord] = 1;
        blk[kTrailerWord + 1] = cnt;
    }
    for (uint32_t i = cnt * kRecWords + threadIdx.x; i < kRpb * kRecWords; i += blockDim.x) blk[kEntriesWord + i] = 0;
}

// S rows (file order) whose key bit is set in the bitmap over [base, base + span]; *d_total receives the match count
// (rows beyond cap_rows are counted, not written).  d_err != 0 afterwards means the look-back chain broke.
int semijoin_stream(const void *d_s_img, uint64_t nblocks_s, int field, const uint32_t *d_bitmap, uint32_t base, uint32_t span,
                    void *d_out, uint64_t cap_rows, uint64_t *d_total, Arena &ws, cudaStream_t st) {
    if (nblocks_s == 0) {
        DBT_CUDA(cudaMemsetAsync(d_total, 0, 16, st));
        return 0;
    }
    if (nblocks_s >= (1ull << 32)) {
        set_error("semijoin: S must have fewer than 2^32 blocks");
        return DBT_ERR_UNSUPPORTED;
    }
    StageScope sc(ST_HASH_PROBE, st);
    const size_t m0 = ws.mark();
    const uint64_t ntiles = (nblocks_s + kSjTileBlocks - 1) / kSjTileBlocks;
    uint64_t *state = ws.take<uint64_t>(ntiles);
    uint32_t *ctr = ws.take<uint32_t>(64);
    if (!state || !ctr) {
        set_error("semijoin: workspace too small");
        return DBT_ERR_WORKSPACE;
    }
    DBT_CUDA(cudaMemsetAsync(state, 0, ntiles * 8, st));
    DBT_CUDA(cudaMemsetAsync(ctr, 0, 8, st));
    DBT_CUDA(cudaMemsetAsync(d_total, 0, 16, st));
    static int per_sm[2] = {0, 0};
    const int f = field == '0' ? 0 : 1;
    if (!per_sm[f]) {
        int occ = 0;
        if (f == 0) DBT_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, semijoin_stream_kernel<0>, kSjThreads, 0));
        else DBT_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, semijoin_stream_kernel<1>, kSjThreads, 0));
        if (const char *e = getenv("DBT_SEMIJOIN_CTAS")) occ = std::min(occ, atoi(e));
        per_sm[f] = std::max(occ, 1);
    }
    int nsm = 148;
    const int grid = (int)std::min<uint64_t>(ntiles, (uint64_t)nsm * per_sm[f]);
    if (f == 0)
        semijoin_stream_kernel<0><<<grid, kSjThreads, 0, st>>>((const uint32_t *)d_s_img, (uint32_t)nblocks_s, d_bitmap, base, span,
                                                               (uint32_t *)d_out, cap_rows, state, ctr, (unsigned long long *)d_total);
    else
        semijoin_stream_kernel<1><<<grid, kSjThreads, 0, st>>>((const uint32_t *)d_s_img, (uint32_t)nblocks_s, d_bitmap, base, span,
                                                               (uint32_t *)d_out, cap_rows, state, ctr, (unsigned long long *)d_total);
    semijoin_finish_kernel<<<1, 256, 0, st>>>((uint32_t *)d_out, (const unsigned long long *)d_total, cap_rows);
    count_launch(2);
    DBT_KERNEL_CHECK();
    // the error flag travels with the total: [1] of the same 16-byte result area is set by the caller's read
    DBT_CUDA(cudaMemcpyAsync(d_total + 1, ctr + 1, 4, cudaMemcpyDeviceToDevice, st));
    ws.release(m0);
    return 0;
}

// =====================================================================================================
// Two passes (the default): no chain, nothing to wait for.
//
//   pass 1  tests every live row's key against the bitmap and writes, per S block, the 100-bit match mask and the match
//           count (20 bytes per 14 KB block).  It reads 8 bytes per row through 64-byte-fill loads (~68 B of DRAM traffic
//           per row); round 2's first version streamed the whole image for it (140 B per row: 9.1 vs 4.8 ms per 400M rows);
//   scan    exclusive prefix of the per-block counts (the device-wide scan of kernels_gather.cu) = the output row of
//           every block's first match, and the total -- so a result that does not fit is refused before a byte moves;
//   pass 2  streams S (cp.async.bulk, three blocks in flight per CTA) and copies every matching 140-byte record from shared
//           memory to its final place in the packed output image (one warp per record: 35 consecutive words, consecutive
//           matches are contiguous).
//
// DRAM traffic 68 + 140 B per S row + 140 B per match, against ~68 + 8 + 8 + 4 + (204 + 140) per match for the column path
// (extraction, probe, compaction, random gather): S's records are read sequentially, once, instead of as a gather.
// =====================================================================================================
constexpr int kTpThreads = 128;
constexpr int kTpStages = 3;

struct TpPipe { // per-CTA bulk-copy pipeline over the blocks b = first + k * step
    uint32_t (*stage)[kBlockWords];
    uint64_t *mbar;
};
__device__ __forceinline__ void tp_issue(const TpPipe &p, const uint32_t *img, uint64_t block, int sidx) {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(sj_smem_u32(&p.mbar[sidx])), "r"((uint32_t)DBT_BLOCK_BYTES) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(sj_smem_u32(p.stage[sidx])),
                 "l"(img + block * kBlockWords), "r"((uint32_t)DBT_BLOCK_BYTES), "r"(sj_smem_u32(&p.mbar[sidx]))
                 : "memory");
}
__device__ __forceinline__ void tp_wait(const TpPipe &p, int sidx, uint32_t parity) {
    uint32_t ok = 0;
    while (!ok)
        asm volatile("{\n\t.reg .pred q;\n\tmbarrier.try_wait.parity.shared::cta.b64 q, [%1], %2;\n\tselp.u32 %0, 1, 0, q;\n\t}"
                     : "=r"(ok)
                     : "r"(sj_smem_u32(&p.mbar[sidx])), "r"(parity)
                     : "memory");
}

// Pass 1 does not need the image: the key is in the record's first 8 bytes, and ld.global.nc.L2::64B (ld_sparse) fills 64 bytes
// per L2 miss instead of a 128-byte line, so testing the keys costs ~68 bytes of DRAM traffic per S row instead of the 140 a
// streaming pass reads (round 2's first version: 9.1 ms per 400M rows; this one 4.8) (the nreserved word of the block header shares its 64 bytes with the first record).  Four blocks
// (4 x 100 independent loads) per CTA step.
constexpr int kCsBlocks = 4;
template <int FIELD>
__global__ void __launch_bounds__(kTpThreads)
semijoin_count_sparse_kernel(const uint32_t *__restrict__ img, uint64_t nblocks, const uint32_t *__restrict__ bm, uint32_t base,
                             uint32_t span, uint4 *__restrict__ masks, uint32_t *__restrict__ counts) {
    __shared__ uint32_t s_mask[kCsBlocks][4];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (uint64_t b0 = (uint64_t)blockIdx.x * kCsBlocks; b0 < nblocks; b0 += (uint64_t)gridDim.x * kCsBlocks) {
        uint32_t key[kCsBlocks], nres[kCsBlocks];
#pragma unroll
        for (int u = 0; u < kCsBlocks; ++u) {
            const uint64_t b = b0 + u;
            key[u] = nres[u] = 0;
            if (b < nblocks) {
                const uint32_t *blk = img + b * kBlockWords;
                nres[u] = ld_sparse(blk + 1);
                if (tid < (int)kRpb) key[u] = ld_sparse(blk + kEntriesWord + tid * kRecWords + (FIELD == 0 ? 0 : 1));
            }
        }
#pragma unroll
        for (int u = 0; u < kCsBlocks; ++u) {
            bool match = false;
            if (tid < (int)min(nres[u], kRpb)) {
                const uint32_t v = key[u] - base;
                match = (v <= span) && ((__ldg(bm + (v >> 5)) >> (v & 31)) & 1u);
            }
            const uint32_t ballot = __ballot_sync(0xFFFFFFFFu, match);
            if (lane == 0) s_mask[u][warp] = ballot;
        }
        __syncthreads();
        if (tid < kCsBlocks && b0 + tid < nblocks) {
            const uint4 m = make_uint4(s_mask[tid][0], s_mask[tid][1], s_mask[tid][2], s_mask[tid][3]);
            masks[b0 + tid] = m;
            counts[b0 + tid] = __popc(m.x) + __popc(m.y) + __popc(m.z) + __popc(m.w);
        }
        __syncthreads(); // s_mask is rewritten by the next step
    }
}

__global__ void __launch_bounds__(kTpThreads)
semijoin_copy_kernel(const uint32_t *__restrict__ img, uint64_t nblocks, const uint4 *__restrict__ masks, const uint32_t *__restrict__ counts,
                     const uint32_t *__restrict__ offs /*[nblocks] output row of the block's first match*/, uint32_t *__restrict__ out) {
    extern __shared__ __align__(128) unsigned char tp_raw[];
    TpPipe pipe{reinterpret_cast<uint32_t(*)[kBlockWords]>(tp_raw),
                reinterpret_cast<uint64_t *>(tp_raw + sizeof(uint32_t) * kBlockWords * kTpStages)};
    __shared__ uint32_t s_src[kRpb], s_dst[kRpb]; // word offset of the k-th match inside the stage / relative output word offset
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint64_t first = blockIdx.x, step = gridDim.x;
    const uint64_t mine = first < nblocks ? (nblocks - first + step - 1) / step : 0;
    if (tid == 0) {
        for (int i = 0; i < kTpStages; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(sj_smem_u32(&pipe.mbar[i])), "r"(1));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        for (uint64_t k = 0; k < (uint64_t)(kTpStages - 1) && k < mine; ++k) tp_issue(pipe, img, first + k * step, (int)k);
    }
    __syncthreads();
    for (uint64_t k = 0; k < mine; ++k) {
        if (tid == 0 && k + kTpStages - 1 < mine) tp_issue(pipe, img, first + (k + kTpStages - 1) * step, (int)((k + kTpStages - 1) % kTpStages));
        const int sidx = (int)(k % kTpStages);
        tp_wait(pipe, sidx, (uint32_t)((k / kTpStages) & 1));
        const uint64_t b = first + k * step;
        const uint32_t m = counts[b];
        const uint32_t *blk = pipe.stage[sidx];
        const uint64_t g0 = offs[b];
        const uint64_t off0 = slot_word(g0);
        if (tid < (int)kRpb) {
            const uint4 mk = masks[b];
            const uint32_t w = tid >> 5, bit = tid & 31;
            const uint32_t word = w == 0 ? mk.x : (w == 1 ? mk.y : (w == 2 ? mk.z : mk.w));
            if ((word >> bit) & 1u) {
                uint32_t r = __popc(word & ((1u << bit) - 1u));
                if (w > 0) r += __popc(mk.x);
                if (w > 1) r += __popc(mk.y);
                if (w > 2) r += __popc(mk.z);
                const uint64_t g = g0 + r, ob = g / kRpb;
                const uint32_t e = (uint32_t)(g - ob * kRpb);
                s_src[r] = kEntriesWord + (uint32_t)tid * kRecWords;
                s_dst[r] = (uint32_t)(ob * kBlockWords + kEntriesWord + (uint64_t)e * kRecWords - off0);
                if (e == 0) { // first row of an output block: its header (the last block's is fixed afterwards)
                    uint32_t *o = out + ob * kBlockWords;
                    o[0] = (uint32_t)ob;
                    o[1] = kRpb;
                    o[kTrailerWord] = 1;
                    o[kTrailerWord + 1] = kRpb;
                }
            }
        }
        __syncthreads();
        for (uint32_t r = warp; r < m; r += kTpThreads / 32) { // one warp per record: 32 + 3 consecutive words
            const uint32_t *src = blk + s_src[r];
            uint32_t *dst = out + off0 + s_dst[r];
            dst[lane] = src[lane];
            if (lane < 3) dst[32 + lane] = src[32 + lane];
        }
        __syncthreads(); // stage and lists are free again
    }
}

int semijoin_two_pass(const void *d_s_img, uint64_t nblocks_s, int field, const uint32_t *d_bitmap, uint32_t base, uint32_t span,
                      void *d_out, uint64_t cap_rows, uint64_t *d_total, uint64_t *h_total, Arena &ws, cudaStream_t st) {
    *h_total = 0;
    DBT_CUDA(cudaMemsetAsync(d_total, 0, 16, st));
    if (nblocks_s == 0) return 0;
    if (nblocks_s >= (1ull << 32)) {
        set_error("semijoin: S must have fewer than 2^32 blocks");
        return DBT_ERR_UNSUPPORTED;
    }
    const size_t m0 = ws.mark();
    uint4 *masks = ws.take<uint4>(nblocks_s);
    uint32_t *counts = ws.take<uint32_t>(nblocks_s), *offs = ws.take<uint32_t>(nblocks_s);
    if (!masks || !counts || !offs) {
        set_error("semijoin: workspace too small");
        return DBT_ERR_WORKSPACE;
    }
    const size_t smem = sizeof(uint32_t) * kBlockWords * kTpStages + 8 * kTpStages + 128;
    static int per_sm[3] = {0, 0, 0};
    const int f = field == '0' ? 0 : 1;
    auto occupancy = [&](const void *fn, int idx) -> int {
        if (first_use_on_device(fn)) DBT_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        if (!per_sm[idx]) {
            int occ = 0;
            DBT_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, fn, kTpThreads, smem));
            per_sm[idx] = std::max(occ, 1);
        }
        return 0;
    };
    DBT_TRY(occupancy((const void *)semijoin_copy_kernel, 2));
    {
        StageScope sc(ST_HASH_PROBE, st);
        const int g = (int)std::min<uint64_t>((nblocks_s + kCsBlocks - 1) / kCsBlocks, (uint64_t)148 * 16);
        if (f == 0) semijoin_count_sparse_kernel<0><<<g, kTpThreads, 0, st>>>((const uint32_t *)d_s_img, nblocks_s, d_bitmap, base, span, masks, counts);
        else semijoin_count_sparse_kernel<1><<<g, kTpThreads, 0, st>>>((const uint32_t *)d_s_img, nblocks_s, d_bitmap, base, span, masks, counts);
        count_launch();
        DBT_KERNEL_CHECK();
    }
    {
        StageScope sc(ST_COMPACT, st);
        DBT_TRY(exclusive_offsets(counts, nblocks_s, offs, d_total, ws, st));
    }
    DBT_CUDA(cudaMemcpyAsync(h_total, d_total, 8, cudaMemcpyDeviceToHost, st));
    DBT_CUDA(cudaStreamSynchronize(st));
    if (*h_total > cap_rows || *h_total == 0) { // nothing to move, or it would not fit: the caller reports the size
        ws.release(m0);
        return 0;
    }
    {
        StageScope sc(ST_GATHER, st);
        const int grid = (int)std::min<uint64_t>(nblocks_s, (uint64_t)148 * per_sm[2]);
        semijoin_copy_kernel<<<grid, kTpThreads, smem, st>>>((const uint32_t *)d_s_img, nblocks_s, masks, counts, offs, (uint32_t *)d_out);
        semijoin_finish_kernel<<<1, 256, 0, st>>>((uint32_t *)d_out, (const unsigned long long *)d_total, cap_rows);
        count_launch(2);
        DBT_KERNEL_CHECK();
    }
    ws.release(m0);
    return 0;
}

} // namespace dbt
