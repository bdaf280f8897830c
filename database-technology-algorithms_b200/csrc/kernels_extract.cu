// kernels_extract.cu -- block headers and AoS -> columnar key extraction.
//
// The reference walks entries[0..nreserved) of every block and compares records through
// compareID/NUM/STR/NUMSTR (DatabaseProject.cpp:44-92,198-205).  Here the keys are pulled out
// once into order-preserving u32 words: recid / num as they are, str as NUL-normalised
// big-endian words (strcmp order == unsigned word order), so every later stage works on
// integers only.
#include "dbt_internal.cuh"
#include <algorithm>
#include <cstdlib>

namespace dbt {

// ---------------------------------------------------------------------------------------------
// Headers: live rows per block.
// ---------------------------------------------------------------------------------------------
struct HeaderStats {
    unsigned long long nrows;
    uint32_t ragged; // some block before the last is not full
    uint32_t pad;
};

__global__ void __launch_bounds__(256) header_kernel(const uint32_t *__restrict__ img, uint64_t nblocks,
                                                     uint32_t *__restrict__ nres_out, HeaderStats *stats) {
    uint64_t b = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t r = 0, ragged = 0;
    if (b < nblocks) {
        r = ld_sparse(img + b * kBlockWords + 1);
        if (r > kRpb) r = kRpb;
        if (b + 1 < nblocks && r != kRpb) ragged = 1;
        if (nres_out) nres_out[b] = r;
    }
    uint32_t sum = __reduce_add_sync(0xFFFFFFFFu, r);
    uint32_t rg = __reduce_or_sync(0xFFFFFFFFu, ragged);
    if ((threadIdx.x & 31) == 0) {
        if (sum) atomicAdd(&stats->nrows, (unsigned long long)sum);
        if (rg) atomicOr(&stats->ragged, 1u);
    }
}

// ragged images (partial blocks in the middle: what a rank receives from the all-to-all, or a user file
// with holes): row offsets of the blocks come from the device-wide scan in kernels_gather.cu, then one
// thread per (block, entry) writes the slot list with coalesced stores.
__global__ void __launch_bounds__(128)
ragged_fill_kernel(const uint32_t *__restrict__ nres, const uint32_t *__restrict__ row_off, uint64_t nblocks,
                   uint32_t *__restrict__ row_slot) {
    for (uint64_t b = blockIdx.x; b < nblocks; b += gridDim.x) {
        const uint32_t e = threadIdx.x;
        if (e < nres[b]) row_slot[row_off[b] + e] = (uint32_t)(b * kRpb + e);
    }
}

int image_info(const void *d_image, uint64_t nblocks, uint32_t **d_row_slot_out, Arena &ws, cudaStream_t st,
               ImageInfo *info) {
    StageScope sc(ST_HEADERS, st);
    *d_row_slot_out = nullptr;
    info->blk_nres = nullptr;
    info->blk_row_off = nullptr;
    info->nrows = 0;
    info->prefix_full = 1;
    if (nblocks == 0) return 0;
    HeaderStats *d_stats = ws.take<HeaderStats>(1);
    uint32_t *d_nres = ws.take<uint32_t>(nblocks);
    if (!d_stats || !d_nres) {
        set_error("image_info: workspace too small");
        return DBT_ERR_WORKSPACE;
    }
    DBT_CUDA(cudaMemsetAsync(d_stats, 0, sizeof(HeaderStats), st));
    int grid = (int)((nblocks + 255) / 256);
    header_kernel<<<grid, 256, 0, st>>>((const uint32_t *)d_image, nblocks, d_nres, d_stats);
    count_launch();
    DBT_KERNEL_CHECK();
    HeaderStats h;
    DBT_CUDA(cudaMemcpyAsync(&h, d_stats, sizeof h, cudaMemcpyDeviceToHost, st));
    DBT_CUDA(cudaStreamSynchronize(st));
    info->nrows = h.nrows;
    info->prefix_full = h.ragged ? 0 : 1;
    if (h.nrows >= (1ull << 30)) {
        set_error("image has >= 2^30 live rows; shard it across GPUs");
        return DBT_ERR_UNSUPPORTED;
    }
    if (h.ragged) {
        uint32_t *slots = ws.take<uint32_t>(h.nrows ? h.nrows : 1);
        if (!slots) {
            set_error("image_info: workspace too small (ragged slot list)");
            return DBT_ERR_WORKSPACE;
        }
        uint32_t *row_off = ws.take<uint32_t>(nblocks);
        uint64_t *d_total = ws.take<uint64_t>(8);
        if (!row_off || !d_total) {
            set_error("image_info: workspace too small (ragged offsets)");
            return DBT_ERR_WORKSPACE;
        }
        DBT_TRY(exclusive_offsets(d_nres, nblocks, row_off, d_total, ws, st));
        int grid = (int)std::min<uint64_t>(nblocks, 148 * 32);
        ragged_fill_kernel<<<grid, 128, 0, st>>>(d_nres, row_off, nblocks, slots);
        count_launch();
        DBT_KERNEL_CHECK();
        *d_row_slot_out = slots;
        info->blk_nres = d_nres;
        info->blk_row_off = row_off;
    }
    return 0;
}

// ---------------------------------------------------------------------------------------------
// Key extraction.
// ---------------------------------------------------------------------------------------------
__global__ void init_stats_kernel(ExtractStats *s) {
    int t = threadIdx.x;
    uint32_t *w = reinterpret_cast<uint32_t *>(s);
    if (t < (int)(sizeof(ExtractStats) / 4)) w[t] = 0;
    __syncthreads();
    if (t == 0) {
        s->and_w0 = 0xFFFFFFFFu;
        s->and_recid = 0xFFFFFFFFu;
    }
    if (t < 32) s->str_and[t] = 0xFFFFFFFFu;
}

// NUL-normalise one little-endian word of a C string and make it big-endian-comparable.
// `ended` carries "a NUL was seen in an earlier word".  strcmp semantics of the reference
// comparators (DatabaseProject.cpp:57-68): bytes after the first NUL never matter.
__device__ __forceinline__ uint32_t norm_word(uint32_t w, bool &ended) {
    if (ended) return 0;
    uint32_t z = (w - 0x01010101u) & ~w & 0x80808080u; // lowest flagged byte is a true zero byte
    if (z) {
        int pos = (__ffs(z) - 1) >> 3; // index of the first NUL byte (byte 0 = lowest address)
        w &= (pos == 0) ? 0u : (0xFFFFFFFFu >> (32 - 8 * pos));
        ended = true;
    }
    return __byte_perm(w, 0, 0x0123);
}

template <int FIELD> // 0,1,2,3 (numeric, not ASCII)
__global__ void __launch_bounds__(256)
extract_kernel(const uint32_t *__restrict__ img, uint64_t nrows, const uint32_t *__restrict__ row_slot, uint32_t kw,
               uint32_t *__restrict__ out_w0, uint32_t *__restrict__ out_str, uint32_t *__restrict__ out_recid,
               ExtractStats *stats) {
    __shared__ uint32_t s_or[34], s_and[34]; // [0]=w0 [1]=recid [2..]=str words
    __shared__ uint32_t s_flags[2];
    const bool HAS_STR = (FIELD >= 2);
    for (int i = threadIdx.x; i < 34; i += blockDim.x) {
        s_or[i] = 0;
        s_and[i] = 0xFFFFFFFFu;
    }
    if (threadIdx.x < 2) s_flags[threadIdx.x] = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31;
    uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool live = r < nrows;
    uint32_t recid = 0, w0 = 0, unsorted = 0, overflow = 0;
    uint64_t base = 0;
    if (live) {
        uint64_t slot = row_slot ? row_slot[r] : r;
        base = slot_word(slot);
        recid = img[base];
        w0 = (FIELD == 0) ? recid : ((FIELD == 2) ? 0u : img[base + 1]);
        out_recid[r] = recid;
        if (FIELD != 2) out_w0[r] = w0;
    }
    // recid monotone in file order?  (then a stable sort on the key alone already breaks ties by recid)
    uint32_t prev = __shfl_up_sync(0xFFFFFFFFu, recid, 1);
    if (live && r > 0) {
        if (lane == 0) {
            uint64_t pslot = row_slot ? row_slot[r - 1] : r - 1;
            prev = img[slot_word(pslot)];
        }
        unsorted = recid < prev;
    }
    uint32_t o = __reduce_or_sync(0xFFFFFFFFu, live ? w0 : 0u);
    uint32_t a = __reduce_and_sync(0xFFFFFFFFu, live ? w0 : 0xFFFFFFFFu);
    uint32_t o2 = __reduce_or_sync(0xFFFFFFFFu, live ? recid : 0u);
    uint32_t a2 = __reduce_and_sync(0xFFFFFFFFu, live ? recid : 0xFFFFFFFFu);
    uint32_t un = __reduce_or_sync(0xFFFFFFFFu, unsorted);
    if (lane == 0) {
        atomicOr(&s_or[0], o);
        atomicAnd(&s_and[0], a);
        atomicOr(&s_or[1], o2);
        atomicAnd(&s_and[1], a2);
        if (un) s_flags[0] = 1;
    }
    if (HAS_STR) {
        bool ended = false;
        for (uint32_t j = 0; j < kw; ++j) {
            uint32_t w = 0;
            if (live) {
                w = norm_word(img[base + kStrWord + j], ended);
                out_str[r * kw + j] = w;
            }
            uint32_t so = __reduce_or_sync(0xFFFFFFFFu, live ? w : 0u);
            uint32_t sa = __reduce_and_sync(0xFFFFFFFFu, live ? w : 0xFFFFFFFFu);
            if (lane == 0) {
                atomicOr(&s_or[2 + j], so);
                atomicAnd(&s_and[2 + j], sa);
            }
        }
        overflow = live && !ended && kw < kStrWords;
        if (__any_sync(0xFFFFFFFFu, overflow) && lane == 0) s_flags[1] = 1;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        atomicOr(&stats->or_w0, s_or[0]);
        atomicAnd(&stats->and_w0, s_and[0]);
        atomicOr(&stats->or_recid, s_or[1]);
        atomicAnd(&stats->and_recid, s_and[1]);
        if (s_flags[0]) atomicOr(&stats->recid_unsorted, 1u);
        if (s_flags[1]) atomicOr(&stats->str_overflow, 1u);
    }
    if (HAS_STR && threadIdx.x < kw) {
        atomicOr(&stats->str_or[threadIdx.x], s_or[2 + threadIdx.x]);
        atomicAnd(&stats->str_and[threadIdx.x], s_and[2 + threadIdx.x]);
    }
}

// ---------------------------------------------------------------------------------------------
// Streaming extraction for the str / num+str keys (fields '2', '3'; fields '0', '1' use the sparse kernel further down).
// ncu (profiles/r01_notes.md): with the default 128-byte fills and rows 140 bytes apart, the strided kernel above pulls
// ~the whole image from DRAM (132 B per row) as scattered sector requests at 3 TB/s.  This kernel streams whole
// 14016-byte blocks into shared memory with cp.async.bulk (TMA bulk copy, one instruction per block, 3 blocks in flight
// per CTA) and picks the key words out of shared memory: same DRAM bytes, sequential, and almost no LSU work.  (40 of a
// row's 140 bytes are key: 64-byte fills would read ~100 B per row at the lower rate of sparse reads -- about even.)
// ---------------------------------------------------------------------------------------------
constexpr int kExThreads = 128;
constexpr int kExStages = 3;

__device__ __forceinline__ uint32_t ex_smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int FIELD>
__global__ void __launch_bounds__(kExThreads)
extract_stream_kernel(const uint32_t *__restrict__ img, uint64_t nblocks, uint64_t nrows, uint32_t kw,
                      const uint32_t *__restrict__ blk_nres, const uint32_t *__restrict__ blk_row_off,
                      uint32_t *__restrict__ out_w0, uint32_t *__restrict__ out_str, uint32_t *__restrict__ out_recid,
                      ExtractStats *stats, uint32_t *__restrict__ byte_hist /*[4][256] or null: histogram of w0's bytes*/) {
    extern __shared__ __align__(128) unsigned char ex_raw[];
    uint32_t(*stage)[kBlockWords] = reinterpret_cast<uint32_t(*)[kBlockWords]>(ex_raw);
    uint64_t *mbar = reinterpret_cast<uint64_t *>(ex_raw + sizeof(uint32_t) * kBlockWords * kExStages);
    __shared__ uint32_t s_or[34], s_and[34], s_flags[2];
    __shared__ uint32_t s_hist[4][256]; // bytes of w0 (the sort's digits when its plan is byte-aligned)
    const bool HAS_STR = (FIELD >= 2);
    const bool HIST = (FIELD == 0 || FIELD == 1) && byte_hist != nullptr;
    const int tid = threadIdx.x, lane = tid & 31;
    if (HIST)
        for (int i = tid; i < 4 * 256; i += kExThreads) (&s_hist[0][0])[i] = 0;
    for (int i = tid; i < 34; i += kExThreads) {
        s_or[i] = 0;
        s_and[i] = 0xFFFFFFFFu;
    }
    if (tid < 2) s_flags[tid] = 0;
    const uint64_t first = blockIdx.x, step = gridDim.x;
    const uint64_t my_blocks = first < nblocks ? (nblocks - first + step - 1) / step : 0;
    auto issue = [&](uint64_t k) { // thread 0: bulk-copy my k-th block into stage k % kExStages
        const int sidx = (int)(k % kExStages);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(ex_smem_u32(&mbar[sidx])),
                     "r"((uint32_t)DBT_BLOCK_BYTES)
                     : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                         ex_smem_u32(stage[sidx])),
                     "l"(img + (first + k * step) * kBlockWords), "r"((uint32_t)DBT_BLOCK_BYTES),
                     "r"(ex_smem_u32(&mbar[sidx]))
                     : "memory");
    };
    if (tid == 0) {
        for (int i = 0; i < kExStages; ++i)
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(ex_smem_u32(&mbar[i])), "r"(1));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        for (uint64_t k = 0; k < (uint64_t)(kExStages - 1) && k < my_blocks; ++k) issue(k);
    }
    __syncthreads();
    uint32_t o_w0 = 0, a_w0 = 0xFFFFFFFFu, o_id = 0, a_id = 0xFFFFFFFFu, unsorted = 0, overflow = 0;
    uint32_t o_s[HAS_STR ? 8 : 1], a_s[HAS_STR ? 8 : 1]; // kw == 8: per-thread OR/AND of the key words
#pragma unroll
    for (int j = 0; j < (HAS_STR ? 8 : 1); ++j) {
        o_s[j] = 0;
        a_s[j] = 0xFFFFFFFFu;
    }
    for (uint64_t k = 0; k < my_blocks; ++k) {
        if (tid == 0 && k + kExStages - 1 < my_blocks) issue(k + kExStages - 1);
        const int sidx = (int)(k % kExStages);
        const uint32_t parity = (uint32_t)((k / kExStages) & 1);
        uint32_t ok = 0;
        while (!ok)
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                         : "=r"(ok)
                         : "r"(ex_smem_u32(&mbar[sidx])), "r"(parity)
                         : "memory");
        const uint64_t b = first + k * step;
        const uint32_t *blk = stage[sidx];
        // block-dense images: row = 100 b + entry; ragged images: the block's row offset comes from the header scan
        const uint64_t row = (blk_row_off ? (uint64_t)blk_row_off[b] : b * kRpb) + tid;
        const bool live = blk_nres ? (tid < (int)blk_nres[b]) : (tid < (int)kRpb && row < nrows);
        const uint32_t *rec = blk + kEntriesWord + (live ? tid : 0) * kRecWords;
        if (live) {
            const uint32_t recid = rec[0];
            const uint32_t w0 = (FIELD == 0) ? recid : ((FIELD == 2) ? 0u : rec[1]);
            out_recid[row] = recid;
            if (FIELD != 2) out_w0[row] = w0;
            if (HIST) { // this kernel waits on DRAM: four shared-memory atomics per row are free here
                atomicAdd(&s_hist[0][w0 & 0xFF], 1u);
                atomicAdd(&s_hist[1][(w0 >> 8) & 0xFF], 1u);
                atomicAdd(&s_hist[2][(w0 >> 16) & 0xFF], 1u);
                atomicAdd(&s_hist[3][w0 >> 24], 1u);
            }
            o_w0 |= w0;
            a_w0 &= w0;
            o_id |= recid;
            a_id &= recid;
            uint32_t prev = recid;
            if (tid > 0) prev = rec[-(int)kRecWords];
            else if (b > 0) { // last live row of the nearest earlier non-empty block
                uint64_t pb = b - 1;
                uint32_t pn = blk_nres ? blk_nres[pb] : kRpb;
                while (pn == 0 && pb > 0) pn = blk_nres[--pb];
                if (pn) prev = img[pb * kBlockWords + kEntriesWord + (pn - 1) * kRecWords];
            }
            unsorted |= recid < prev;
        }
        if (HAS_STR && kw == 8) {
            // the common width: the row's eight key words are built in registers and leave as two 16-byte stores (a warp
            // writes 1 KB contiguous in two instructions instead of eight strided 4-byte stores); OR/AND stay in
            // registers until the CTA has seen all of its blocks
            if (live) {
                bool ended = false;
                uint32_t w[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    w[j] = norm_word(rec[kStrWord + j], ended);
                    o_s[j] |= w[j];
                    a_s[j] &= w[j];
                }
                uint4 *dst = reinterpret_cast<uint4 *>(out_str + row * 8);
                dst[0] = make_uint4(w[0], w[1], w[2], w[3]);
                dst[1] = make_uint4(w[4], w[5], w[6], w[7]);
                overflow |= !ended;
            }
        } else if (HAS_STR) { // whole warps take part: the per-word OR/AND go through warp reductions
            bool ended = false;
            uint32_t *dst = out_str + row * kw;
            for (uint32_t j = 0; j < kw; ++j) {
                uint32_t w = 0;
                if (live) {
                    w = norm_word(rec[kStrWord + j], ended);
                    dst[j] = w;
                }
                const uint32_t so = __reduce_or_sync(0xFFFFFFFFu, live ? w : 0u);
                const uint32_t sa = __reduce_and_sync(0xFFFFFFFFu, live ? w : 0xFFFFFFFFu);
                if (lane == 0) {
                    atomicOr(&s_or[2 + j], so);
                    atomicAnd(&s_and[2 + j], sa);
                }
            }
            overflow |= (live && !ended && kw < kStrWords);
        }
        __syncthreads(); // everyone is done with this stage before it is refilled
    }
    o_w0 = __reduce_or_sync(0xFFFFFFFFu, o_w0);
    a_w0 = __reduce_and_sync(0xFFFFFFFFu, a_w0);
    o_id = __reduce_or_sync(0xFFFFFFFFu, o_id);
    a_id = __reduce_and_sync(0xFFFFFFFFu, a_id);
    unsorted = __reduce_or_sync(0xFFFFFFFFu, unsorted);
    overflow = __reduce_or_sync(0xFFFFFFFFu, overflow);
    if (HAS_STR && kw == 8) {
#pragma unroll
        for (int j = 0; j < (HAS_STR ? 8 : 1); ++j) {
            const uint32_t so = __reduce_or_sync(0xFFFFFFFFu, o_s[j]);
            const uint32_t sa = __reduce_and_sync(0xFFFFFFFFu, a_s[j]);
            if (lane == 0) {
                atomicOr(&s_or[2 + j], so);
                atomicAnd(&s_and[2 + j], sa);
            }
        }
    }
    if (lane == 0) {
        atomicOr(&s_or[0], o_w0);
        atomicAnd(&s_and[0], a_w0);
        atomicOr(&s_or[1], o_id);
        atomicAnd(&s_and[1], a_id);
        if (unsorted) s_flags[0] = 1;
        if (overflow) s_flags[1] = 1;
    }
    __syncthreads();
    if (my_blocks == 0) return;
    if (HIST)
        for (int i = tid; i < 4 * 256; i += kExThreads) {
            const uint32_t c = (&s_hist[0][0])[i];
            if (c) atomicAdd(&byte_hist[i], c);
        }
    if (tid == 0) {
        atomicOr(&stats->or_w0, s_or[0]);
        atomicAnd(&stats->and_w0, s_and[0]);
        atomicOr(&stats->or_recid, s_or[1]);
        atomicAnd(&stats->and_recid, s_and[1]);
        if (s_flags[0]) atomicOr(&stats->recid_unsorted, 1u);
        if (s_flags[1]) atomicOr(&stats->str_overflow, 1u);
    }
    if (HAS_STR && tid < (int)kw) {
        atomicOr(&stats->str_or[tid], s_or[2 + tid]);
        atomicAnd(&stats->str_and[tid], s_and[2 + tid]);
    }
}

// ---------------------------------------------------------------------------------------------
// Sparse extraction for fields '0' / '1' (round 2).  The key of those fields is the record's first 8 bytes (recid, num).
// ld.global.nc.L2::64B makes an L2 miss fill 64 bytes instead of a 128-byte line (dbt_internal.cuh: ld_sparse), so a
// strided read of 8 bytes per 140-byte row costs ~68 bytes of DRAM traffic per row instead of the whole image (140):
// the streaming kernel above reads 14.0 GB per 100M rows, this one ~6.9 GB.  Persistent CTAs, 4 rows (8 independent
// loads) per thread per step, OR/AND in registers, byte histograms of the key in shared memory as in the streaming kernel.
// Serves dense and ragged images alike (row_slot).
// ---------------------------------------------------------------------------------------------
constexpr int kSpThreads = 256;
constexpr int kSpRows = 4;

template <int FIELD>
__global__ void __launch_bounds__(kSpThreads)
extract_sparse_kernel(const uint32_t *__restrict__ img, uint64_t nrows, const uint32_t *__restrict__ row_slot,
                      uint32_t *__restrict__ out_w0, uint32_t *__restrict__ out_recid, ExtractStats *stats,
                      uint32_t *__restrict__ byte_hist /*[4][256] or null*/) {
    static_assert(FIELD == 0 || FIELD == 1, "numeric fields only");
    __shared__ uint32_t s_hist[4][256];
    __shared__ uint32_t s_or[2], s_and[2], s_flag;
    const int tid = threadIdx.x, lane = tid & 31;
    const bool HIST = byte_hist != nullptr;
    if (HIST)
        for (int i = tid; i < 4 * 256; i += kSpThreads) (&s_hist[0][0])[i] = 0;
    if (tid < 2) {
        s_or[tid] = 0;
        s_and[tid] = 0xFFFFFFFFu;
    }
    if (tid == 0) s_flag = 0;
    __syncthreads();
    uint32_t o_w0 = 0, a_w0 = 0xFFFFFFFFu, o_id = 0, a_id = 0xFFFFFFFFu, unsorted = 0;
    constexpr uint64_t kTile = (uint64_t)kSpThreads * kSpRows;
    for (uint64_t t0 = (uint64_t)blockIdx.x * kTile; t0 < nrows; t0 += (uint64_t)gridDim.x * kTile) {
        uint32_t recid[kSpRows], w0[kSpRows], prev0[kSpRows];
#pragma unroll
        for (int u = 0; u < kSpRows; ++u) { // every load of the step is issued before the first use
            const uint64_t r = t0 + (uint64_t)u * kSpThreads + tid;
            recid[u] = w0[u] = prev0[u] = 0;
            if (r < nrows) {
                const uint64_t base = slot_word(row_slot ? row_slot[r] : r);
                recid[u] = ld_sparse(img + base);
                w0[u] = (FIELD == 0) ? recid[u] : ld_sparse(img + base + 1);
                if (lane == 0 && r > 0) prev0[u] = ld_sparse(img + slot_word(row_slot ? row_slot[r - 1] : r - 1));
            }
        }
#pragma unroll
        for (int u = 0; u < kSpRows; ++u) {
            const uint64_t r = t0 + (uint64_t)u * kSpThreads + tid;
            const bool live = r < nrows;
            uint32_t prev = __shfl_up_sync(0xFFFFFFFFu, recid[u], 1);
            if (lane == 0) prev = prev0[u];
            if (live) {
                out_recid[r] = recid[u];
                out_w0[r] = w0[u];
                if (HIST) {
                    atomicAdd(&s_hist[0][w0[u] & 0xFF], 1u);
                    atomicAdd(&s_hist[1][(w0[u] >> 8) & 0xFF], 1u);
                    atomicAdd(&s_hist[2][(w0[u] >> 16) & 0xFF], 1u);
                    atomicAdd(&s_hist[3][w0[u] >> 24], 1u);
                }
                o_w0 |= w0[u];
                a_w0 &= w0[u];
                o_id |= recid[u];
                a_id &= recid[u];
                unsorted |= (r > 0 && recid[u] < prev) ? 1u : 0u; // recid monotone in file order? (see extract_kernel)
            }
        }
    }
    o_w0 = __reduce_or_sync(0xFFFFFFFFu, o_w0);
    a_w0 = __reduce_and_sync(0xFFFFFFFFu, a_w0);
    o_id = __reduce_or_sync(0xFFFFFFFFu, o_id);
    a_id = __reduce_and_sync(0xFFFFFFFFu, a_id);
    unsorted = __reduce_or_sync(0xFFFFFFFFu, unsorted);
    if (lane == 0) {
        atomicOr(&s_or[0], o_w0);
        atomicAnd(&s_and[0], a_w0);
        atomicOr(&s_or[1], o_id);
        atomicAnd(&s_and[1], a_id);
        if (unsorted) s_flag = 1;
    }
    __syncthreads();
    if ((uint64_t)blockIdx.x * kTile >= nrows) return;
    if (HIST)
        for (int i = tid; i < 4 * 256; i += kSpThreads) {
            const uint32_t c = (&s_hist[0][0])[i];
            if (c) atomicAdd(&byte_hist[i], c);
        }
    if (tid == 0) {
        atomicOr(&stats->or_w0, s_or[0]);
        atomicAnd(&stats->and_w0, s_and[0]);
        atomicOr(&stats->or_recid, s_or[1]);
        atomicAnd(&stats->and_recid, s_and[1]);
        if (s_flag) atomicOr(&stats->recid_unsorted, 1u);
    }
}

template <int FIELD>
static int launch_extract_stream(const uint32_t *img, uint64_t nblocks, uint64_t nrows, uint32_t kw,
                                 const uint32_t *blk_nres, const uint32_t *blk_row_off, uint32_t *d_w0,
                                 uint32_t *d_str, uint32_t *d_recid, ExtractStats *d_stats, uint32_t *d_byte_hist,
                                 cudaStream_t st) {
    size_t smem = sizeof(uint32_t) * kBlockWords * kExStages + 8 * kExStages + 128;
    auto kfn = extract_stream_kernel<FIELD>;
    static int per_sm = 0; // resident CTAs per SM: the grid is exactly one wave of them
    if (first_use_on_device((const void *)kfn)) DBT_CUDA(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    if (!per_sm) {
        int occ = 0;
        DBT_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kfn, kExThreads, smem));
        if (const char *e = getenv("DBT_EXTRACT_CTAS")) occ = std::min(occ, atoi(e));
        per_sm = std::max(occ, 1);
    }
    int grid = (int)std::min<uint64_t>(nblocks, (uint64_t)148 * per_sm);
    kfn<<<grid, kExThreads, smem, st>>>(img, nblocks, nrows, kw, blk_nres, blk_row_off, d_w0, d_str, d_recid, d_stats, d_byte_hist);
    return 0;
}

int extract_keys(const void *d_image, uint64_t nblocks_img, uint64_t nrows, const uint32_t *d_row_slot,
                 const uint32_t *d_blk_nres, const uint32_t *d_blk_row_off, int field, uint32_t kw, uint32_t *d_w0,
                 uint32_t *d_str, uint32_t *d_recid, ExtractStats *d_stats, cudaStream_t st, uint32_t *d_byte_hist,
                 int *hist_done) {
    StageScope sc(ST_EXTRACT, st);
    if (hist_done) *hist_done = 0;
    init_stats_kernel<<<1, 128, 0, st>>>(d_stats);
    count_launch();
    const bool ragged = d_row_slot != nullptr;
    const bool stream_ok = (!ragged || (d_blk_nres && d_blk_row_off)) && ((uintptr_t)d_image % 16 == 0) &&
                           getenv("DBT_EXTRACT_STRIDED") == nullptr;
    const bool sparse = field == '0' || field == '1'; // the key is the record's first 8 bytes: strided 64-byte-fill reads
    if (d_byte_hist && sparse && nrows)
        DBT_CUDA(cudaMemsetAsync(d_byte_hist, 0, 4 * 256 * 4, st));
    if (nrows && sparse) {
        const uint32_t *img = (const uint32_t *)d_image;
        const int grid = (int)std::min<uint64_t>((nrows + kSpThreads * kSpRows - 1) / (kSpThreads * kSpRows), 148 * 8);
        if (field == '0') extract_sparse_kernel<0><<<grid, kSpThreads, 0, st>>>(img, nrows, d_row_slot, d_w0, d_recid, d_stats, d_byte_hist);
        else extract_sparse_kernel<1><<<grid, kSpThreads, 0, st>>>(img, nrows, d_row_slot, d_w0, d_recid, d_stats, d_byte_hist);
        count_launch();
        if (hist_done && d_byte_hist) *hist_done = 1;
    } else if (nrows && stream_ok) {
        const uint32_t *img = (const uint32_t *)d_image;
        const uint64_t nblocks = ragged ? nblocks_img : (nrows + kRpb - 1) / kRpb;
        const uint32_t *bn = ragged ? d_blk_nres : nullptr, *bo = ragged ? d_blk_row_off : nullptr;
        switch (field) {
        case '2': DBT_TRY(launch_extract_stream<2>(img, nblocks, nrows, kw, bn, bo, d_w0, d_str, d_recid, d_stats, d_byte_hist, st)); break;
        case '3': DBT_TRY(launch_extract_stream<3>(img, nblocks, nrows, kw, bn, bo, d_w0, d_str, d_recid, d_stats, d_byte_hist, st)); break;
        default: set_error("bad field"); return DBT_ERR_ARG;
        }
        count_launch();
    } else if (nrows) {
        int grid = (int)((nrows + 255) / 256);
        const uint32_t *img = (const uint32_t *)d_image;
        switch (field) {
        case '2': extract_kernel<2><<<grid, 256, 0, st>>>(img, nrows, d_row_slot, kw, d_w0, d_str, d_recid, d_stats); break;
        case '3': extract_kernel<3><<<grid, 256, 0, st>>>(img, nrows, d_row_slot, kw, d_w0, d_str, d_recid, d_stats); break;
        default: set_error("bad field"); return DBT_ERR_ARG;
        }
        count_launch();
    }
    DBT_KERNEL_CHECK();
    return 0;
}

} // namespace dbt
