// kernels_gather.cu -- adjacent-difference unique, count compaction (single-pass decoupled
// look-back scans) and the final record gather / block packer.
//
//  * unique_rows   replaces the reference's previous-record compare loop in EliminateDuplicates
//                  (DatabaseProject.cpp:121-162);
//  * compact_by_count turns per-row match counts into the list of rows to emit, in file order
//                  (the reference appends matches one by one, DatabaseProject.cpp:584-629);
//  * gather_records is the only place 140-byte records move (the reference copies every record
//                  in every pass, DatabaseProject.cpp:202,223,303,338).
#include "dbt_internal.cuh"
#include <algorithm>
#include <cstdlib>
#include <cstring>

namespace dbt {

// ---------------------------------------------------------------------------------------------
// Single-pass scan skeleton: each CTA takes a tile (dynamic id), sums its counts, resolves its
// exclusive prefix by decoupled look-back over 64-bit tile states, then emits.
// ---------------------------------------------------------------------------------------------
constexpr uint64_t kSfAgg = 1ull << 62, kSfInc = 2ull << 62, kSvMask = (1ull << 62) - 1;

__device__ __forceinline__ uint64_t ld_volatile64(const uint64_t *p) {
    uint64_t v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_volatile64(uint64_t *p, uint64_t v) {
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

struct ScanWs {
    uint64_t *state; // [ntiles] zeroed
    uint32_t *ctr;   // zeroed
};

#ifndef DBT_SCAN_THREADS
#define DBT_SCAN_THREADS 384 // x 3 CTAs per SM: 0.450 ms per 100M rows (512 x 2: 0.483, 256 x 4: 0.486, 1024 x 1: 0.495)
#endif
#ifndef DBT_SCAN_MINB
#define DBT_SCAN_MINB 3
#endif
constexpr int kScanThreads = DBT_SCAN_THREADS;
constexpr int kScanItems = 16; // two groups of 8 consecutive rows per thread
constexpr int kScanTile = kScanThreads * kScanItems;

// CountFn: void load(uint64_t i0, uint64_t n, uint32_t c[8], uint32_t pay[8]) const
//            -- for the 8 consecutive rows i0..i0+7: how many outputs each produces, and a payload word
//               (vector loads: a thread's 8 rows are 32 contiguous bytes of every input column)
// EmitFn : kTwo (second output column?), v1(i, pay) / v2(i, pay) = the values row i emits (c times), out1/out2/cap.
//          A tile's outputs are first compacted in shared memory and then written coalesced; a tile that emits
//          more than it holds (field-'3' multiplicities) writes directly.
template <class CountFn, class EmitFn>
__global__ void __launch_bounds__(kScanThreads, DBT_SCAN_MINB) // <= 56 registers: three 384-thread CTAs per SM
scan_emit_kernel(uint64_t n, CountFn cnt, EmitFn emit, ScanWs ws, unsigned long long *total_out) {
    __shared__ uint64_t s_wsum[kScanThreads / 32];
    __shared__ uint64_t s_prefix;
    __shared__ uint32_t s_tile;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) s_tile = atomicAdd(ws.ctr, 1u);
    __syncthreads();
    const uint32_t tile = s_tile;
    const uint64_t i0 = (uint64_t)tile * kScanTile + (uint64_t)tid * kScanItems; // blocked arrangement
    uint32_t c[kScanItems], pay[kScanItems];
    uint64_t local = 0;
    cnt.load(i0, n, c, pay);
    cnt.load(i0 + 8, n, c + 8, pay + 8);
#pragma unroll
    for (int k = 0; k < kScanItems; ++k) local += c[k];
    uint64_t x = local;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint64_t t = __shfl_up_sync(0xFFFFFFFFu, x, o);
        if (lane >= o) x += t;
    }
    if (lane == 31) s_wsum[warp] = x;
    __syncthreads();
    uint64_t pre = 0, agg = 0;
#pragma unroll
    for (int w = 0; w < kScanThreads / 32; ++w) {
        uint64_t t = s_wsum[w];
        if (w < warp) pre += t;
        agg += t;
    }
    if (warp == 0) { // warp-parallel decoupled look-back: 32 predecessors per round trip
        uint64_t excl = 0;
        if (tile == 0) {
            if (lane == 0) st_volatile64(&ws.state[0], kSfInc | agg);
        } else {
            if (lane == 0) st_volatile64(&ws.state[tile], kSfAgg | agg);
            int64_t p = (int64_t)tile - 1;
            while (true) {
                const int64_t idx = p - lane;
                const uint64_t sv = (idx >= 0) ? ld_volatile64(&ws.state[idx]) : kSfInc;
                const uint32_t ready = __ballot_sync(0xFFFFFFFFu, (sv >> 62) != 0);
                const uint32_t inc = __ballot_sync(0xFFFFFFFFu, (sv & kSfInc) != 0);
                const int first_inc = inc ? (__ffs(inc) - 1) : 32;
                const uint32_t need = (first_inc >= 31) ? 0xFFFFFFFFu : ((2u << first_inc) - 1u);
                if ((ready & need) != need) continue; // a predecessor before the first inclusive one is not published yet
                uint64_t v = (lane <= first_inc) ? (sv & kSvMask) : 0ull;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, o);
                excl += v;
                if (first_inc < 32) break;
                p -= 32;
            }
            if (lane == 0) st_volatile64(&ws.state[tile], kSfInc | (excl + agg));
        }
        if (lane == 0) {
            s_prefix = excl;
            if ((uint64_t)(tile + 1) * kScanTile >= n) *total_out = excl + agg; // last tile
        }
    }
    __syncthreads();
    extern __shared__ uint32_t s_stage[]; // [kScanTile] (+ [kScanTile] when EmitFn::kTwo)
    const uint64_t tile_prefix = s_prefix;
    if constexpr (EmitFn::kCustom) { // the functor writes its own (variable-length) output for row i at [off, off+c)
        uint64_t off = tile_prefix + pre + x - local;
#pragma unroll 1
        for (int k = 0; k < kScanItems; ++k) {
            if (c[k]) emit.expand(i0 + k, off, c[k], pay[k]);
            off += c[k];
        }
        return;
    } else if constexpr (EmitFn::kIndexed) { // plain exclusive scan: every input row records its offset
        uint64_t off = tile_prefix + pre + x - local;
#pragma unroll
        for (int k = 0; k < kScanItems; ++k) {
            if (i0 + k < n) emit.at(i0 + k, off);
            off += c[k];
        }
        return;
    } else if (agg <= (uint64_t)kScanTile) {
        uint32_t lo = (uint32_t)(pre + x - local); // tile-local offset of this thread's first output
#pragma unroll
        for (int k = 0; k < kScanItems; ++k) {
            if (c[k]) {
                const uint32_t a = emit.v1(i0 + k, pay[k]);
                const uint32_t b = EmitFn::kTwo ? emit.v2(i0 + k, pay[k]) : 0u;
                for (uint32_t t = 0; t < c[k]; ++t) {
                    s_stage[lo + t] = a;
                    if (EmitFn::kTwo) s_stage[kScanTile + lo + t] = b;
                }
                lo += c[k];
            }
        }
        __syncthreads();
        for (uint32_t i = tid; i < (uint32_t)agg; i += kScanThreads) {
            const uint64_t g = tile_prefix + i;
            if (g < emit.cap) {
                emit.out1[g] = s_stage[i];
                if (EmitFn::kTwo) emit.out2[g] = s_stage[kScanTile + i];
            }
        }
    } else { // rare (a tile emitting more than it holds): keep this path small, it only costs registers
        uint64_t off = tile_prefix + pre + x - local;
#pragma unroll 1
        for (int k = 0; k < kScanItems; ++k) {
            if (c[k]) {
                const uint32_t a = emit.v1(i0 + k, pay[k]);
                const uint32_t b = EmitFn::kTwo ? emit.v2(i0 + k, pay[k]) : 0u;
                for (uint32_t t = 0; t < c[k]; ++t)
                    if (off + t < emit.cap) {
                        emit.out1[off + t] = a;
                        if (EmitFn::kTwo) emit.out2[off + t] = b;
                    }
            }
            off += c[k];
        }
    }
}

template <class CountFn, class EmitFn>
static int run_scan_emit(uint64_t n, CountFn cnt, EmitFn emit, uint64_t *d_total, Arena &ws, cudaStream_t st) {
    if (n == 0) {
        DBT_CUDA(cudaMemsetAsync(d_total, 0, 8, st));
        return 0;
    }
    size_t m = ws.mark();
    uint64_t ntiles = (n + kScanTile - 1) / kScanTile;
    ScanWs sw;
    sw.state = ws.take<uint64_t>(ntiles);
    sw.ctr = ws.take<uint32_t>(64);
    if (!sw.state || !sw.ctr) {
        set_error("scan: workspace too small");
        return DBT_ERR_WORKSPACE;
    }
    DBT_CUDA(cudaMemsetAsync(sw.state, 0, ntiles * 8, st));
    DBT_CUDA(cudaMemsetAsync(sw.ctr, 0, 4, st));
    const size_t smem = (size_t)kScanTile * 4 * (EmitFn::kTwo ? 2 : 1);
    auto kfn = scan_emit_kernel<CountFn, EmitFn>;
    if (first_use_on_device((const void *)kfn))
        DBT_CUDA(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kfn<<<(unsigned)ntiles, kScanThreads, smem, st>>>(n, cnt, emit, sw, (unsigned long long *)d_total);
    count_launch();
    DBT_KERNEL_CHECK();
    ws.release(m);
    return 0;
}

// ---------------------------------------------------------------------------------------------
// unique: keep sorted position i iff its key differs from position i-1's key.
// ---------------------------------------------------------------------------------------------
struct RowKeyEq { // full-key equality of two rows through the key columns
    const uint32_t *w0;
    const uint32_t *str;
    uint32_t kw;
    __device__ bool operator()(uint32_t a, uint32_t b) const {
        if (w0 && w0[a] != w0[b]) return false;
        if (str) {
            const uint32_t *pa = str + (uint64_t)a * kw, *pb = str + (uint64_t)b * kw;
            for (uint32_t j = 0; j < kw; ++j)
                if (pa[j] != pb[j]) return false;
        }
        return true;
    }
};

// 8 consecutive words of a column (32 contiguous, 32-byte aligned bytes when the tile is inside the array)
__device__ __forceinline__ void load8(const uint32_t *col, uint64_t i0, uint64_t n, uint32_t v[8]) {
    if (i0 + 8 <= n) {
        const uint4 a = *reinterpret_cast<const uint4 *>(col + i0);
        const uint4 b = *reinterpret_cast<const uint4 *>(col + i0 + 4);
        v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
        v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
    } else {
#pragma unroll
        for (int k = 0; k < 8; ++k) v[k] = (i0 + k < n) ? col[i0 + k] : 0u;
    }
}

struct UniqueCountSorted { // 1-word keys: the sorted key column is at hand; payload = the row (perm[i])
    const uint32_t *sorted;
    const uint32_t *perm;
    __device__ void load(uint64_t i0, uint64_t n, uint32_t c[8], uint32_t pay[8]) const {
        uint32_t key[8];
        load8(sorted, i0, n, key);
        load8(perm, i0, n, pay);
        uint32_t prev = (i0 > 0 && i0 < n) ? sorted[i0 - 1] : 0u;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            c[k] = (i0 + k < n) ? ((i0 + k == 0 || key[k] != prev) ? 1u : 0u) : 0u;
            prev = key[k];
        }
    }
};
struct UniqueCountRows {
    const uint32_t *perm;
    RowKeyEq eq;
    __device__ void load(uint64_t i0, uint64_t n, uint32_t c[8], uint32_t pay[8]) const {
        load8(perm, i0, n, pay);
        if (eq.str && eq.kw == 8) {
            // 32-byte string keys: fetch each row's key once (two 16-byte loads) and carry it in registers as the
            // "previous key" of the next position -- the generic path below would fetch every key twice
            uint4 pa, pb;
            uint32_t pw = 0;
            const uint32_t prow = (i0 > 0 && i0 < n) ? perm[i0 - 1] : 0u;
            {
                const uint4 *kp = reinterpret_cast<const uint4 *>(eq.str + (uint64_t)prow * 8);
                pa = kp[0];
                pb = kp[1];
                if (eq.w0) pw = eq.w0[prow];
            }
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                if (i0 + k < n) {
                    const uint4 *kp = reinterpret_cast<const uint4 *>(eq.str + (uint64_t)pay[k] * 8);
                    const uint4 a = kp[0], b = kp[1];
                    const uint32_t w = eq.w0 ? eq.w0[pay[k]] : 0u;
                    const bool same = (i0 + k > 0) && w == pw && a.x == pa.x && a.y == pa.y && a.z == pa.z && a.w == pa.w &&
                                      b.x == pb.x && b.y == pb.y && b.z == pb.z && b.w == pb.w;
                    c[k] = same ? 0u : 1u;
                    pa = a;
                    pb = b;
                    pw = w;
                } else {
                    c[k] = 0u;
                }
            }
            return;
        }
        uint32_t prev = (i0 > 0 && i0 < n) ? perm[i0 - 1] : 0u;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            c[k] = (i0 + k < n) ? ((i0 + k == 0 || !eq(prev, pay[k])) ? 1u : 0u) : 0u;
            prev = pay[k];
        }
    }
};
template <bool TWO>
struct EmitPerm { // out1 = the row (payload); out2 = its key from the sorted column (optional)
    static constexpr bool kTwo = TWO;
    static constexpr bool kIndexed = false;
    static constexpr bool kCustom = false;
    uint32_t *out1;
    uint32_t *out2;
    const uint32_t *sorted;
    uint64_t cap;
    __device__ void at(uint64_t, uint64_t) const {}
    __device__ void expand(uint64_t, uint64_t, uint32_t, uint32_t) const {}
    __device__ uint32_t v1(uint64_t, uint32_t row) const { return row; }
    __device__ uint32_t v2(uint64_t i, uint32_t) const { return sorted[i]; }
};

struct UniqueCountSorted2 { // two-word keys, both columns in sorted order; payload = the row (perm[i])
    const uint32_t *hi, *lo, *perm;
    __device__ void load(uint64_t i0, uint64_t n, uint32_t c[8], uint32_t pay[8]) const {
        uint32_t h[8], l[8];
        load8(hi, i0, n, h);
        load8(lo, i0, n, l);
        load8(perm, i0, n, pay);
        uint32_t ph = (i0 > 0 && i0 < n) ? hi[i0 - 1] : 0u, pl = (i0 > 0 && i0 < n) ? lo[i0 - 1] : 0u;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            c[k] = (i0 + k < n) ? ((i0 + k == 0 || h[k] != ph || l[k] != pl) ? 1u : 0u) : 0u;
            ph = h[k];
            pl = l[k];
        }
    }
};
template <bool TWO>
struct EmitPermPos { // out1 = the row (payload); out2 = its sorted position (optional)
    static constexpr bool kTwo = TWO;
    static constexpr bool kIndexed = false;
    static constexpr bool kCustom = false;
    uint32_t *out1;
    uint32_t *out2;
    uint64_t cap;
    __device__ void at(uint64_t, uint64_t) const {}
    __device__ void expand(uint64_t, uint64_t, uint32_t, uint32_t) const {}
    __device__ uint32_t v1(uint64_t, uint32_t row) const { return row; }
    __device__ uint32_t v2(uint64_t i, uint32_t) const { return (uint32_t)i; }
};
int unique_rows_sorted2(const uint32_t *d_hi, const uint32_t *d_lo, const uint32_t *d_perm, uint64_t n, uint32_t *d_uperm,
                        uint32_t *d_upos, uint64_t *d_count, Arena &ws, cudaStream_t st) {
    StageScope sc(ST_UNIQUE, st);
    if (d_upos) return run_scan_emit(n, UniqueCountSorted2{d_hi, d_lo, d_perm}, EmitPermPos<true>{d_uperm, d_upos, n}, d_count, ws, st);
    return run_scan_emit(n, UniqueCountSorted2{d_hi, d_lo, d_perm}, EmitPermPos<false>{d_uperm, nullptr, n}, d_count, ws, st);
}
__global__ void __launch_bounds__(256)
take_pairs_kernel(const uint32_t *__restrict__ hi, const uint32_t *__restrict__ lo, const uint32_t *__restrict__ pos, uint64_t n,
                  uint2 *__restrict__ out) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const uint32_t q = pos[i];
        out[i] = make_uint2(hi[q], lo[q]);
    }
}
int take_pairs(const uint32_t *d_hi, const uint32_t *d_lo, const uint32_t *d_pos, uint64_t n, uint32_t *d_out, cudaStream_t st) {
    if (!n) return 0;
    StageScope sc(ST_UNIQUE, st);
    const int grid = (int)std::min<uint64_t>((n + 255) / 256, 148 * 16);
    take_pairs_kernel<<<grid, 256, 0, st>>>(d_hi, d_lo, d_pos, n, (uint2 *)d_out);
    count_launch();
    DBT_KERNEL_CHECK();
    return 0;
}

int unique_rows(const KeyCols &k, int field, const uint32_t *d_perm, const uint32_t *d_sorted_w0, uint64_t n,
                uint32_t *d_uperm, uint32_t *d_ukeys, uint64_t *d_count, Arena &ws, cudaStream_t st) {
    StageScope sc(ST_UNIQUE, st);
    if ((field == '0' || field == '1') && d_sorted_w0) {
        if (d_ukeys)
            return run_scan_emit(n, UniqueCountSorted{d_sorted_w0, d_perm}, EmitPerm<true>{d_uperm, d_ukeys, d_sorted_w0, n}, d_count, ws, st);
        return run_scan_emit(n, UniqueCountSorted{d_sorted_w0, d_perm}, EmitPerm<false>{d_uperm, nullptr, nullptr, n}, d_count, ws, st);
    }
    RowKeyEq eq{(field == '2') ? nullptr : k.w0, (field >= '2') ? k.str : nullptr, k.kw};
    return run_scan_emit(n, UniqueCountRows{d_perm, eq}, EmitPerm<false>{d_uperm, nullptr, nullptr, n}, d_count, ws, st);
}

// ---------------------------------------------------------------------------------------------
// compaction of per-row counts into a row list (row i repeated counts[i] times, ascending i)
// ---------------------------------------------------------------------------------------------
struct CountFromArray {
    const uint32_t *c;
    const uint32_t *values; // optional payload column (else the row index itself)
    __device__ void load(uint64_t i0, uint64_t n, uint32_t cnt[8], uint32_t pay[8]) const {
        load8(c, i0, n, cnt);
        if (values) load8(values, i0, n, pay);
        else {
#pragma unroll
            for (int k = 0; k < 8; ++k) pay[k] = (uint32_t)(i0 + k);
        }
    }
};
struct EmitRowRepeated {
    static constexpr bool kTwo = false;
    static constexpr bool kIndexed = false;
    static constexpr bool kCustom = false;
    uint32_t *out1;
    uint32_t *out2;
    uint64_t cap;
    __device__ void at(uint64_t, uint64_t) const {}
    __device__ void expand(uint64_t, uint64_t, uint32_t, uint32_t) const {}
    __device__ uint32_t v1(uint64_t, uint32_t v) const { return v; }
    __device__ uint32_t v2(uint64_t, uint32_t) const { return 0u; }
};
int compact_select(const uint32_t *d_counts, const uint32_t *d_values, uint64_t n, uint32_t *d_out, uint64_t out_cap,
                   uint64_t *d_total, Arena &ws, cudaStream_t st) {
    StageScope sc(ST_COMPACT, st);
    return run_scan_emit(n, CountFromArray{d_counts, d_values}, EmitRowRepeated{d_out, nullptr, out_cap}, d_total, ws, st);
}

// inner-join pair expansion: S row i with `c` matches starting at sorted-R position `first` emits
// (recid_R[rperm[first+t]], recid_S[i]) for t < c
struct EmitPairs {
    static constexpr bool kTwo = false;
    static constexpr bool kIndexed = false;
    static constexpr bool kCustom = true;
    const uint32_t *rperm, *r_recid, *s_recid;
    uint32_t *pairs;
    uint32_t *out1 = nullptr, *out2 = nullptr;
    uint64_t cap;
    __device__ void at(uint64_t, uint64_t) const {}
    __device__ uint32_t v1(uint64_t, uint32_t) const { return 0u; }
    __device__ uint32_t v2(uint64_t, uint32_t) const { return 0u; }
    __device__ void expand(uint64_t i, uint64_t off, uint32_t c, uint32_t first) const {
        const uint32_t sid = s_recid[i];
        for (uint32_t t = 0; t < c; ++t)
            if (off + t < cap) {
                pairs[2 * (off + t)] = r_recid[rperm[first + t]];
                pairs[2 * (off + t) + 1] = sid;
            }
    }
};
int expand_pairs(const uint32_t *d_count, const uint32_t *d_first, uint64_t ns, const uint32_t *d_rperm,
                 const uint32_t *d_r_recid, const uint32_t *d_s_recid, uint32_t *d_pairs, uint64_t cap, uint64_t *d_total,
                 Arena &ws, cudaStream_t st) {
    StageScope sc(ST_COMPACT, st);
    EmitPairs e;
    e.rperm = d_rperm;
    e.r_recid = d_r_recid;
    e.s_recid = d_s_recid;
    e.pairs = d_pairs;
    e.cap = cap;
    return run_scan_emit(ns, CountFromArray{d_count, d_first}, e, d_total, ws, st);
}

// exclusive prefix sums of a u32 array (offsets < 2^32): out[i] = sum(counts[0..i))
struct EmitOffset {
    static constexpr bool kTwo = false;
    static constexpr bool kIndexed = true;
    static constexpr bool kCustom = false;
    uint32_t *out;
    uint32_t *out1 = nullptr, *out2 = nullptr; // unused by the indexed mode
    uint64_t cap = 0;
    __device__ void at(uint64_t i, uint64_t off) const { out[i] = (uint32_t)off; }
    __device__ void expand(uint64_t, uint64_t, uint32_t, uint32_t) const {}
    __device__ uint32_t v1(uint64_t, uint32_t) const { return 0u; }
    __device__ uint32_t v2(uint64_t, uint32_t) const { return 0u; }
};
int exclusive_offsets(const uint32_t *d_counts, uint64_t n, uint32_t *d_off, uint64_t *d_total, Arena &ws,
                      cudaStream_t st) {
    EmitOffset e;
    e.out = d_off;
    return run_scan_emit(n, CountFromArray{d_counts, nullptr}, e, d_total, ws, st);
}

// ---------------------------------------------------------------------------------------------
// Record gather + block packer.  One CTA builds one 14016-byte output block in shared memory
// from 4-byte gathered words (records are only 4-byte aligned) and stores it with
// 16-byte vectors; 100 independent 140-byte random reads per block are in flight per CTA.
// ---------------------------------------------------------------------------------------------
constexpr int kGatherThreads = 256;

__global__ void __launch_bounds__(256) blockid_add_kernel(uint32_t *__restrict__ img, uint64_t nblocks, uint32_t blockid0) {
    const uint64_t b = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (b < nblocks) img[b * kBlockWords] += blockid0;
}

// The loop is software-pipelined (round 2).  With 64-byte fills (ld_sparse: DRAM read 267 -> 202 B per record) the first
// version -- offsets, loads, barrier, stage, barrier, store, barrier per block -- no longer saturated DRAM (ncu: 65 % busy,
// long-scoreboard and barrier stalls; a CTA's record loads were in flight ~40 % of its time).  Here the loads of block b+1
// are issued BEFORE block b is stored and the rows -> slot -> word-offset chain of block b+2 is fetched under them: record
// loads are in flight nearly all the time, two barriers per block.  6.03 -> 5.84 (64-byte fills) -> 5.63-5.66 ms per 90M rows.
// Two other designs end at the same 5.6 ms = 5.5 TB/s of actual traffic, which is therefore the memory system's rate for
// this mix of 64-byte random reads and streaming writes (profiles/r02_notes.md): aligned 16-byte cp.async.cg chunks into
// 160-byte slots, three blocks deep, no registers under the loads (5.64 ms); 4-byte cp.async.ca copies were 35 % slower (7.86).
__device__ __forceinline__ uint32_t ld_sparse_pinned(const uint32_t *p) { // stays where it is written (before the block store)
    uint32_t v;
#if DBT_SPARSE_LD
    asm volatile("ld.global.nc.L2::64B.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
#else
    asm volatile("ld.global.nc.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
#endif
    return v;
}

__global__ void __launch_bounds__(kGatherThreads, 5) // 48 registers, five CTAs per SM (four: 5.69 ms, six: 6.04 ms per 90M rows)
gather_kernel(const uint32_t *__restrict__ in, const uint32_t *__restrict__ rows, const uint32_t *__restrict__ row_slot,
              uint64_t nrows_out, uint4 *__restrict__ out, uint64_t nblocks_out) {
    __shared__ __align__(16) uint32_t stage[kBlockWords];
    __shared__ uint64_t src[2][kRpb];
    const int tid = threadIdx.x;
    constexpr int kPerThread = (kRpb * kRecWords + kGatherThreads - 1) / kGatherThreads;
    auto count_of = [&](uint64_t ob) { return (uint32_t)min((uint64_t)kRpb, nrows_out - ob * kRpb); };
    auto fetch_src = [&](uint64_t ob) -> uint64_t { // source word offset of record `tid` of block ob
        if (ob < nblocks_out && tid < (int)count_of(ob)) {
            const uint64_t r = ob * kRpb + tid;
            const uint32_t row = rows ? rows[r] : (uint32_t)r;
            const uint64_t slot = row_slot ? row_slot[row] : row;
            return slot_word(slot);
        }
        return 0;
    };
    uint32_t v[kPerThread];
    auto issue = [&](uint64_t ob, int s) { // all of the thread's loads for block ob (offsets in src[s])
        const uint32_t nwords = count_of(ob) * kRecWords;
#pragma unroll
        for (int k = 0; k < kPerThread; ++k) {
            const uint32_t idx = tid + k * kGatherThreads;
            const uint32_t rec = idx / kRecWords;
            const uint32_t w = idx - rec * kRecWords;
            v[k] = (idx < nwords) ? ld_sparse_pinned(in + src[s][rec] + w) : 0u;
        }
    };
    auto fill_stage = [&](uint64_t ob) {
#pragma unroll
        for (int k = 0; k < kPerThread; ++k) {
            const uint32_t idx = tid + k * kGatherThreads;
            if (idx < kRpb * kRecWords) stage[kEntriesWord + idx] = v[k];
        }
        if (tid == 0) {
            const uint32_t cnt = count_of(ob);
            stage[0] = (uint32_t)ob;       // blockid
            stage[1] = cnt;                // nreserved
            stage[kTrailerWord] = 1;       // valid=1, misc=0, padding 0
            stage[kTrailerWord + 1] = cnt; // dummy
        }
    };
    uint64_t ob = blockIdx.x;
    if (ob >= nblocks_out) return;
    {
        const uint64_t w0 = fetch_src(ob);
        const uint64_t w1 = fetch_src(ob + gridDim.x);
        if (tid < (int)kRpb) {
            src[0][tid] = w0;
            src[1][tid] = w1;
        }
    }
    __syncthreads();
    issue(ob, 0);
    fill_stage(ob);
    __syncthreads();
    for (int s = 0;; s ^= 1) {
        const uint64_t nxt = ob + gridDim.x;
        const bool more = nxt < nblocks_out;
        if (more) issue(nxt, s ^ 1);                        // in flight while block ob is stored
        const uint64_t w2 = fetch_src(nxt + gridDim.x);     // (and so is the offset chain of the block after that)
        const uint4 *sv = reinterpret_cast<const uint4 *>(stage);
        uint4 *dst = out + ob * kBlockVec4;
        for (uint32_t i = tid; i < kBlockVec4; i += kGatherThreads) dst[i] = sv[i];
        __syncthreads(); // the stage is free; src[s] was last read when block ob's loads were issued
        if (!more) break;
        fill_stage(nxt);
        if (tid < (int)kRpb) src[s][tid] = w2;
        __syncthreads();
        ob = nxt;
    }
}

int gather_records(const void *d_in, const uint32_t *d_rows, const uint32_t *d_row_slot, uint64_t nrows_out, void *d_out,
                   cudaStream_t st, int max_ctas, uint32_t blockid0) {
    StageScope sc(ST_GATHER, st);
    if (nrows_out == 0) return 0;
    uint64_t nb = (nrows_out + kRpb - 1) / kRpb;
    int grid = (int)std::min<uint64_t>(nb, max_ctas > 0 ? (uint64_t)max_ctas : 148 * 5); // persistent: the resident CTAs
    gather_kernel<<<grid, kGatherThreads, 0, st>>>((const uint32_t *)d_in, d_rows, d_row_slot, nrows_out, (uint4 *)d_out, nb);
    if (blockid0) { // a chunk of a larger image (out-of-core): renumber in a separate tiny pass -- an extra parameter in
                    // the round-1 gather_kernel itself changed its schedule and cost 6 % (profiles/r01_notes.md)
        blockid_add_kernel<<<(unsigned)((nb + 255) / 256), 256, 0, st>>>((uint32_t *)d_out, nb, blockid0);
        count_launch();
    }
    count_launch();
    DBT_KERNEL_CHECK();
    return 0;
}

} // namespace dbt

extern "C" int dbt_gather_records(const void *d_in_image, const uint32_t *d_rows, const uint32_t *d_row_slot,
                                  uint64_t nrows_out, void *d_out_image, void *stream) {
    if (!d_in_image || !d_out_image) {
        dbt::set_error("dbt_gather_records: NULL image");
        return DBT_ERR_ARG;
    }
    if (((uintptr_t)d_in_image & 3) || ((uintptr_t)d_out_image & 15)) {
        dbt::set_error("dbt_gather_records: the output image must be 16-byte aligned (the input 4-byte)");
        return DBT_ERR_ARG;
    }
    return dbt::gather_records(d_in_image, d_rows, d_row_slot, nrows_out, d_out_image, (cudaStream_t)stream, 0);
}
