// kernels_join.cu -- hash build/probe and sorted-list intersection.
//
//  * hash_join_counts replaces the reference's unordered_map / unordered_multimap build and
//    find/equal_range probe (DatabaseProject.cpp:510-547, 584-629): a linear-probing
//    open-addressing table in HBM keyed by the u32 key (fields '0'/'1') or by row id with full-key
//    compare (fields '2'/'3').  The result is, per S row, how many times the reference would emit
//    it: 0/1 for the set semantics of fields '0'..'2', the number of matching R rows for '3'.
//  * intersect_sorted replaces MergeJoin's two-pointer walk over the two deduplicated files
//    (DatabaseProject.cpp:414-482) and reproduces the number of block reads that walk performs.
#include "dbt_internal.cuh"
#include <algorithm>
#include <cstdlib>

namespace dbt {

constexpr uint32_t kEmptyKey = 0xFFFFFFFFu;

__device__ __forceinline__ uint32_t hash_u32(uint32_t k) {
    k ^= k >> 16;
    k *= 0x7FEB352Du;
    k ^= k >> 15;
    k *= 0x846CA68Bu;
    k ^= k >> 16;
    return k;
}

struct KeyView { // device view of KeyCols
    const uint32_t *w0;
    const uint32_t *str;
    uint32_t kw;
};
__device__ __forceinline__ uint32_t hash_row(const KeyView &k, uint32_t row) {
    uint32_t h = k.w0 ? hash_u32(k.w0[row]) : 0x9E3779B9u;
    if (k.str) {
        const uint32_t *p = k.str + (uint64_t)row * k.kw;
        for (uint32_t j = 0; j < k.kw; ++j) h = hash_u32(h ^ p[j]) + j;
    }
    return h;
}
__device__ __forceinline__ bool rows_equal(const KeyView &a, uint32_t ra, const KeyView &b, uint32_t rb) {
    if (a.w0 && a.w0[ra] != b.w0[rb]) return false;
    if (a.str) {
        const uint32_t *pa = a.str + (uint64_t)ra * a.kw, *pb = b.str + (uint64_t)rb * b.kw;
        for (uint32_t j = 0; j < a.kw; ++j)
            if (pa[j] != pb[j]) return false;
    }
    return true;
}
// three-way compare of row ra of a against row rb of b (both key sets have the same shape)
__device__ __forceinline__ int rows_cmp(const KeyView &a, uint32_t ra, const KeyView &b, uint32_t rb) {
    if (a.w0) {
        uint32_t x = a.w0[ra], y = b.w0[rb];
        if (x != y) return x < y ? -1 : 1;
    }
    if (a.str) {
        const uint32_t *pa = a.str + (uint64_t)ra * a.kw, *pb = b.str + (uint64_t)rb * b.kw;
        for (uint32_t j = 0; j < a.kw; ++j)
            if (pa[j] != pb[j]) return pa[j] < pb[j] ? -1 : 1;
    }
    return 0;
}

// ---- u32 keys: the table stores the keys themselves ---------------------------------------
__global__ void __launch_bounds__(256)
build_u32_kernel(const uint32_t *__restrict__ keys, uint64_t n, uint32_t *table, uint32_t mask, uint32_t *has_max) {
    uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        uint32_t k = keys[i];
        if (k == kEmptyKey) { // the sentinel value itself: remembered on the side
            *has_max = 1;
            continue;
        }
        uint32_t h = hash_u32(k) & mask;
        while (true) {
            uint32_t cur = table[h];
            if (cur == k) break;
            if (cur == kEmptyKey) {
                uint32_t old = atomicCAS(&table[h], kEmptyKey, k);
                if (old == kEmptyKey || old == k) break;
            }
            h = (h + 1) & mask;
        }
    }
}
__global__ void __launch_bounds__(256)
probe_u32_kernel(const uint32_t *__restrict__ keys, uint64_t n, const uint32_t *__restrict__ table, uint32_t mask,
                 const uint32_t *__restrict__ has_max, uint32_t *__restrict__ counts) {
    uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    const uint32_t hm = *has_max;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        uint32_t k = keys[i];
        uint32_t found = 0;
        if (k == kEmptyKey) {
            found = hm;
        } else {
            uint32_t h = hash_u32(k) & mask;
            while (true) {
                uint32_t cur = table[h];
                if (cur == k) {
                    found = 1;
                    break;
                }
                if (cur == kEmptyKey) break;
                h = (h + 1) & mask;
            }
        }
        counts[i] = found;
    }
}

// ---- multi-word keys: the table stores (R row + 1); equality goes through the key columns ---
template <bool MULTI>
__global__ void __launch_bounds__(256)
build_rows_kernel(KeyView r, uint64_t n, uint32_t *table, uint32_t *mult, uint32_t mask) {
    uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        uint32_t row = (uint32_t)i;
        uint32_t h = hash_row(r, row) & mask;
        while (true) {
            uint32_t cur = table[h];
            if (cur == 0) {
                cur = atomicCAS(&table[h], 0u, row + 1);
                if (cur == 0) cur = row + 1;
            }
            if (cur == row + 1 || rows_equal(r, cur - 1, r, row)) {
                if (MULTI) atomicAdd(&mult[h], 1u);
                break;
            }
            h = (h + 1) & mask;
        }
    }
}
template <bool MULTI>
__global__ void __launch_bounds__(256)
probe_rows_kernel(KeyView r, KeyView s, uint64_t n, const uint32_t *__restrict__ table,
                  const uint32_t *__restrict__ mult, uint32_t mask, uint32_t *__restrict__ counts) {
    uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        uint32_t row = (uint32_t)i;
        uint32_t h = hash_row(s, row) & mask;
        uint32_t c = 0;
        while (true) {
            uint32_t cur = table[h];
            if (cur == 0) break;
            if (rows_equal(r, cur - 1, s, row)) {
                c = MULTI ? mult[h] : 1u;
                break;
            }
            h = (h + 1) & mask;
        }
        counts[i] = c;
    }
}

// ---- direct-address bitmap (u32 keys, set semantics) ---------------------------------------
// When the build side's keys span a range whose bitmap fits in L2 (<= 2^29 keys => 64 MB of the
// 126 MB L2), membership is one L2-resident bit test: the probe becomes a pure stream over S's key
// column instead of one random 128-byte DRAM line per row (ncu: the hash probe moves ~136 B per row).
// Exact, order-free, skew-proof; wider key ranges fall back to the hash table below.
__global__ void __launch_bounds__(256) minmax_kernel(const uint32_t *__restrict__ k, uint64_t n, uint32_t *out /*min,max*/) {
    uint32_t lo = 0xFFFFFFFFu, hi = 0;
    uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        uint32_t v = k[i];
        lo = min(lo, v);
        hi = max(hi, v);
    }
    lo = __reduce_min_sync(0xFFFFFFFFu, lo);
    hi = __reduce_max_sync(0xFFFFFFFFu, hi);
    if ((threadIdx.x & 31) == 0) {
        atomicMin(&out[0], lo);
        atomicMax(&out[1], hi);
    }
}
__global__ void __launch_bounds__(256)
bitmap_build_kernel(const uint32_t *__restrict__ k, uint64_t n, uint32_t base, uint32_t *__restrict__ bm) {
    uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        uint32_t v = k[i] - base;
        atomicOr(&bm[v >> 5], 1u << (v & 31));
    }
}
__global__ void __launch_bounds__(256)
bitmap_probe_kernel(const uint32_t *__restrict__ k, uint64_t n, uint32_t base, uint32_t span,
                    const uint32_t *__restrict__ bm, uint32_t *__restrict__ counts) {
    uint64_t nvec = n / 4;
    uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    const uint4 *kv = reinterpret_cast<const uint4 *>(k);
    uint4 *cv = reinterpret_cast<uint4 *>(counts);
    auto test = [&](uint32_t key) -> uint32_t {
        uint32_t v = key - base;
        return (v <= span) ? ((__ldg(&bm[v >> 5]) >> (v & 31)) & 1u) : 0u;
    };
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += stride) {
        uint4 a = kv[i];
        cv[i] = make_uint4(test(a.x), test(a.y), test(a.z), test(a.w));
    }
    if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {
        uint64_t i = nvec * 4 + threadIdx.x;
        counts[i] = test(k[i]);
    }
}
constexpr uint64_t kBitmapMaxSpan = 1ull << 29; // 64 MB bitmap: stays L2-resident next to the streamed columns

// Wider key ranges: the range is cut into slices of 2^29 keys; for each slice one launch streams S's key column
// and tests only the keys that fall into the slice, so the 64 MB of bitmap being hit is L2-resident during that
// launch.  Every S row belongs to exactly one slice and is written exactly once.  The bitmap of the full 32-bit
// range is 512 MB in HBM.
__global__ void __launch_bounds__(256)
bitmap_probe_slice_kernel(const uint32_t *__restrict__ k, uint64_t n, uint32_t base, uint32_t slice,
                          const uint32_t *__restrict__ bm, uint32_t *__restrict__ counts) {
    uint64_t nvec = n / 4;
    uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    const uint4 *kv = reinterpret_cast<const uint4 *>(k);
    auto probe = [&](uint32_t key, uint64_t i) {
        uint32_t v = key - base; // keys below base wrap to huge values: they land in a slice beyond the bitmap and are handled there
        if ((v >> 29) == slice) counts[i] = (__ldg(&bm[v >> 5]) >> (v & 31)) & 1u;
    };
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += stride) {
        uint4 a = kv[i];
        probe(a.x, 4 * i);
        probe(a.y, 4 * i + 1);
        probe(a.z, 4 * i + 2);
        probe(a.w, 4 * i + 3);
    }
    if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {
        uint64_t i = nvec * 4 + threadIdx.x;
        probe(k[i], i);
    }
}

// Direct-address bitmap of a u32 key column over [min, max] (any span: up to 512 MB for the full 32-bit range), taken
// from the workspace.  *bm == nullptr afterwards means the workspace had no room for it (the caller falls back).
int build_key_bitmap(const uint32_t *d_keys, uint64_t n, Arena &ws, cudaStream_t st, uint32_t **bm, uint32_t *base, uint32_t *span) {
    *bm = nullptr;
    *base = *span = 0;
    if (!n) return 0;
    uint32_t *aux = ws.take<uint32_t>(64);
    if (!aux) return 0;
    uint32_t h_mm[2] = {0xFFFFFFFFu, 0};
    StageScope sc(ST_HASH_BUILD, st);
    DBT_CUDA(cudaMemcpyAsync(aux, h_mm, 8, cudaMemcpyHostToDevice, st));
    const int g = (int)std::min<uint64_t>((n + 255) / 256, 148 * 8);
    minmax_kernel<<<g, 256, 0, st>>>(d_keys, n, aux);
    count_launch();
    DBT_CUDA(cudaMemcpyAsync(h_mm, aux, 8, cudaMemcpyDeviceToHost, st));
    DBT_CUDA(cudaStreamSynchronize(st));
    const uint64_t words = ((uint64_t)h_mm[1] - h_mm[0]) / 32 + 1;
    uint32_t *b = ws.take<uint32_t>(words);
    if (!b) return 0;
    DBT_CUDA(cudaMemsetAsync(b, 0, words * 4, st));
    const int gb = (int)std::min<uint64_t>((n + 255) / 256 + 1, 148 * 16);
    bitmap_build_kernel<<<gb, 256, 0, st>>>(d_keys, n, h_mm[0], b);
    count_launch();
    DBT_KERNEL_CHECK();
    *bm = b;
    *base = h_mm[0];
    *span = h_mm[1] - h_mm[0];
    return 0;
}

size_t hash_table_slots(uint64_t nr) {
    uint64_t want = nr * 2 + 64, cap = 1024;
    while (cap < want) cap <<= 1;
    return cap;
}

int hash_join_counts(const KeyCols &r, const KeyCols &s, int field, uint32_t *d_counts, Arena &ws, cudaStream_t st) {
    if (s.n == 0) return 0;
    size_t slots = hash_table_slots(r.n);
    if (slots > (1ull << 32)) {
        set_error("hash table too large");
        return DBT_ERR_UNSUPPORTED;
    }
    uint32_t mask = (uint32_t)(slots - 1);
    uint32_t *table = ws.take<uint32_t>(slots);
    uint32_t *aux = ws.take<uint32_t>(64);
    if (!table || !aux) {
        set_error("hash join: workspace too small");
        return DBT_ERR_WORKSPACE;
    }
    int gb = (int)std::min<uint64_t>((r.n + 255) / 256 + 1, 148 * 16);
    int gp = (int)std::min<uint64_t>((s.n + 255) / 256 + 1, 148 * 16);
    if ((field == '0' || field == '1') && r.n && getenv("DBT_JOIN_NO_BITMAP") == nullptr) {
        uint32_t h_mm[2] = {0xFFFFFFFFu, 0};
        {
            StageScope sc(ST_HASH_BUILD, st);
            DBT_CUDA(cudaMemcpyAsync(aux + 8, h_mm, 8, cudaMemcpyHostToDevice, st));
            int g = (int)std::min<uint64_t>((r.n + 255) / 256, 148 * 8);
            minmax_kernel<<<g, 256, 0, st>>>(r.w0, r.n, aux + 8);
            count_launch();
            DBT_CUDA(cudaMemcpyAsync(h_mm, aux + 8, 8, cudaMemcpyDeviceToHost, st));
            DBT_CUDA(cudaStreamSynchronize(st));
        }
        const uint64_t span = (uint64_t)h_mm[1] - h_mm[0];
        if (span >= kBitmapMaxSpan && getenv("DBT_JOIN_NO_SLICES") == nullptr) {
            // sliced bitmap over [min, min + 2^32): one bit per possible key, cleared, filled from R, probed slice by slice
            const uint64_t words = (1ull << 32) / 32;
            uint32_t *bm = ws.take<uint32_t>(words);
            if (bm) {
                {
                    StageScope sc(ST_HASH_BUILD, st);
                    DBT_CUDA(cudaMemsetAsync(bm, 0, words * 4, st));
                    bitmap_build_kernel<<<gb, 256, 0, st>>>(r.w0, r.n, h_mm[0], bm);
                    count_launch();
                    DBT_KERNEL_CHECK();
                }
                StageScope sc(ST_HASH_PROBE, st);
                for (uint32_t slice = 0; slice < 8; ++slice) { // keys outside [min, max] hit zero bits: no match, as it must be
                    bitmap_probe_slice_kernel<<<gp, 256, 0, st>>>(s.w0, s.n, h_mm[0], slice, bm, d_counts);
                    count_launch();
                }
                DBT_KERNEL_CHECK();
                return 0;
            }
        }
        if (span < kBitmapMaxSpan && (span / 32 + 1) <= slots) { // the table allocation doubles as the bitmap
            const uint64_t words = span / 32 + 1;
            {
                StageScope sc(ST_HASH_BUILD, st);
                DBT_CUDA(cudaMemsetAsync(table, 0, words * 4, st));
                bitmap_build_kernel<<<gb, 256, 0, st>>>(r.w0, r.n, h_mm[0], table);
                count_launch();
                DBT_KERNEL_CHECK();
            }
            StageScope sc(ST_HASH_PROBE, st);
            bitmap_probe_kernel<<<gp, 256, 0, st>>>(s.w0, s.n, h_mm[0], (uint32_t)span, table, d_counts);
            count_launch();
            DBT_KERNEL_CHECK();
            return 0;
        }
    }
    if (field == '0' || field == '1') {
        {
            StageScope sc(ST_HASH_BUILD, st);
            DBT_CUDA(cudaMemsetAsync(table, 0xFF, slots * 4, st));
            DBT_CUDA(cudaMemsetAsync(aux, 0, 4, st));
            if (r.n) {
                build_u32_kernel<<<gb, 256, 0, st>>>(r.w0, r.n, table, mask, aux);
                count_launch();
            }
            DBT_KERNEL_CHECK();
        }
        StageScope sc(ST_HASH_PROBE, st);
        probe_u32_kernel<<<gp, 256, 0, st>>>(s.w0, s.n, table, mask, aux, d_counts);
        count_launch();
        DBT_KERNEL_CHECK();
        return 0;
    }
    KeyView rv{field == '2' ? nullptr : r.w0, r.str, r.kw};
    KeyView sv{field == '2' ? nullptr : s.w0, s.str, s.kw};
    const bool multi = (field == '3');
    uint32_t *mult = nullptr;
    if (multi) {
        mult = ws.take<uint32_t>(slots);
        if (!mult) {
            set_error("hash join: workspace too small");
            return DBT_ERR_WORKSPACE;
        }
    }
    {
        StageScope sc(ST_HASH_BUILD, st);
        DBT_CUDA(cudaMemsetAsync(table, 0, slots * 4, st));
        if (multi) DBT_CUDA(cudaMemsetAsync(mult, 0, slots * 4, st));
        if (r.n) {
            if (multi) build_rows_kernel<true><<<gb, 256, 0, st>>>(rv, r.n, table, mult, mask);
            else build_rows_kernel<false><<<gb, 256, 0, st>>>(rv, r.n, table, mult, mask);
            count_launch();
        }
        DBT_KERNEL_CHECK();
    }
    StageScope sc(ST_HASH_PROBE, st);
    if (multi) probe_rows_kernel<true><<<gp, 256, 0, st>>>(rv, sv, s.n, table, mult, mask, d_counts);
    else probe_rows_kernel<false><<<gp, 256, 0, st>>>(rv, sv, s.n, table, mult, mask, d_counts);
    count_launch();
    DBT_KERNEL_CHECK();
    return 0;
}

// ---------------------------------------------------------------------------------------------
// Radix-partitioned build / probe with per-partition open-addressing tables in SHARED memory
// (north_star; the table's shape -- linear probing, duplicate-rejecting insert -- is the reference's
// HashTable.cpp:32-91, its use DatabaseProject.cpp:537-544 (build) and :604-629 (probe, multiplicity for field '3')).
//
// For keys that are exact 64-bit values (str / num+str keys whose varying bits over R and S together fit 64 bits; the
// order-preserving compaction of device_ops.cu makes them): both sides are grouped by the top B bits of a mixing hash
// of the key -- one or two onesweep passes over (partition id, row) pairs, the same scatter the sort uses -- with B
// chosen so that a partition of R holds ~2,000 rows.  One CTA per partition then builds the partition's table in
// shared memory (key + multiplicity counter, atomicCAS on the 64-bit key), streams the partition's S rows through it
// and writes, per S row, how often the reference would emit it.  Every access that can miss is a 4-byte column read
// through the grouped row list (one sector); the table itself never leaves the SM.  The linear-probing table in HBM
// above stays as the fallback: keys wider than 64 varying bits, or a partition with more distinct keys than a table holds.
// ---------------------------------------------------------------------------------------------
constexpr uint32_t kPartSlots = 8192;             // 64 KB of keys + 32 KB of counters: two CTAs per SM
constexpr uint32_t kPartTargetRows = 2048;        // R rows per partition the partition count aims at
constexpr unsigned long long kPartEmpty = ~0ull;  // (the all-ones key itself is counted on the side)

// A BIJECTION of the 64-bit key (the finalizer of MurmurHash3: xor-shifts and odd multiplications, each invertible):
// equal mixed keys <=> equal keys, so the mixed key can stand in for the key everywhere, and its top bits are the
// partition id.  That is what lets the key travel THROUGH the partition passes as the (key, value) pair the onesweep
// kernel moves anyway -- the join kernel then reads a partition's keys sequentially instead of fetching them through
// a row list (one random 128-byte DRAM access per 4-byte word: the first version spent 5.8 of its 8.7 ms there).
__device__ __forceinline__ uint64_t mix_key64(uint32_t hi, uint32_t lo) {
    uint64_t x = ((uint64_t)hi << 32) | lo;
    x ^= x >> 33;
    x *= 0xFF51AFD7ED558CCDull;
    x ^= x >> 33;
    x *= 0xC4CEB9FE1A85EC53ull;
    x ^= x >> 33;
    return x;
}
__global__ void __launch_bounds__(256)
mix_keys_kernel(const uint32_t *__restrict__ hi, const uint32_t *__restrict__ lo, uint64_t n, uint32_t *__restrict__ mhi,
                uint32_t *__restrict__ mlo, uint32_t *__restrict__ mhi2 /*second copy of mhi (or null)*/) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const uint64_t m = mix_key64(hi ? hi[i] : 0u, lo[i]);
        mhi[i] = (uint32_t)(m >> 32);
        mlo[i] = (uint32_t)m;
        if (mhi2) mhi2[i] = (uint32_t)(m >> 32);
    }
}
// start[p] = first position of partition p (= top `bits` bits of the grouped mixed-hi column); n for trailing empty ones
__global__ void __launch_bounds__(256)
part_bounds_kernel(const uint32_t *__restrict__ grouped_mhi, uint64_t n, uint32_t bits, uint32_t *__restrict__ start /*[2^bits + 1], preset to 0xFFFFFFFF*/) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const uint32_t p = bits ? (grouped_mhi[i] >> (32 - bits)) : 0u;
        if (i == 0 || (bits ? (grouped_mhi[i - 1] >> (32 - bits)) : 0u) != p) start[p] = (uint32_t)i;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) start[1u << bits] = (uint32_t)n;
}
__global__ void __launch_bounds__(256) part_fill_kernel(uint32_t *start, uint32_t nparts) {
    // empty partitions start where the next non-empty one does: every thread looks ahead from its own entry (runs of empty
    // partitions are short: the partition ids are hash bits)
    const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= nparts || start[p] != 0xFFFFFFFFu) return;
    uint32_t q = p + 1;
    while (start[q] == 0xFFFFFFFFu) ++q; // start[nparts] = n is always set
    __syncwarp(__activemask());
    start[p] = start[q];
}

template <bool MULTI>
__global__ void __launch_bounds__(256, 2)
smem_join_kernel(const uint32_t *__restrict__ r_mhi, const uint32_t *__restrict__ r_mlo, const uint32_t *__restrict__ r_start,
                 const uint32_t *__restrict__ s_mhi, const uint32_t *__restrict__ s_mlo, const uint32_t *__restrict__ s_rows,
                 const uint32_t *__restrict__ s_start, uint32_t nparts, uint32_t *__restrict__ counts /*[ns] by S row*/,
                 uint32_t *__restrict__ overflow) {
    extern __shared__ __align__(16) unsigned char pj_raw[];
    unsigned long long *keys = reinterpret_cast<unsigned long long *>(pj_raw);
    uint32_t *cnt = reinterpret_cast<uint32_t *>(pj_raw + sizeof(unsigned long long) * kPartSlots);
    __shared__ uint32_t s_ones, s_over;
    const int tid = threadIdx.x;
    for (uint32_t p = blockIdx.x; p < nparts; p += gridDim.x) {
        const uint32_t rs = r_start[p], re = r_start[p + 1], ss = s_start[p], se = s_start[p + 1];
        if (ss == se) continue; // nobody probes this partition
        for (uint32_t i = tid; i < kPartSlots; i += blockDim.x) {
            keys[i] = kPartEmpty;
            cnt[i] = 0;
        }
        if (tid == 0) s_ones = s_over = 0;
        __syncthreads();
        for (uint32_t i = rs + tid; i < re; i += blockDim.x) { // build: duplicate keys share a slot and bump its counter
            const unsigned long long k = ((unsigned long long)r_mhi[i] << 32) | r_mlo[i];
            if (k == kPartEmpty) {
                atomicAdd(&s_ones, 1u);
                continue;
            }
            uint32_t h = (uint32_t)(k >> 7) & (kPartSlots - 1); // (the top bits are the partition id: use middle ones)
            uint32_t steps = 0;
            while (true) {
                const unsigned long long cur = atomicCAS(&keys[h], kPartEmpty, k);
                if (cur == kPartEmpty || cur == k) {
                    atomicAdd(&cnt[h], 1u);
                    break;
                }
                h = (h + 1) & (kPartSlots - 1);
                if (++steps >= kPartSlots) { // more distinct keys than slots: the caller reruns with the table in HBM
                    s_over = 1;
                    break;
                }
            }
        }
        __syncthreads();
        if (s_over) {
            if (tid == 0) atomicExch(overflow, 1u);
            __syncthreads();
            continue;
        }
        const uint32_t ones = s_ones;
        for (uint32_t j = ss + tid; j < se; j += blockDim.x) { // probe
            const unsigned long long k = ((unsigned long long)s_mhi[j] << 32) | s_mlo[j];
            uint32_t c = 0;
            if (k == kPartEmpty) {
                c = ones;
            } else {
                uint32_t h = (uint32_t)(k >> 7) & (kPartSlots - 1);
                for (uint32_t steps = 0; steps < kPartSlots; ++steps) {
                    const unsigned long long cur = keys[h];
                    if (cur == k) {
                        c = cnt[h];
                        break;
                    }
                    if (cur == kPartEmpty) break;
                    h = (h + 1) & (kPartSlots - 1);
                }
            }
            counts[s_rows[j]] = MULTI ? c : (c ? 1u : 0u);
        }
        __syncthreads(); // the table is rebuilt for the next partition
    }
}

// per S row: how often the reference emits it (multi: once per equal R row, else 0/1); keys = exact (hi, lo) words of both
// sides (hi may be null).  *overflowed is set when some partition had more distinct keys than a shared-memory table holds
// (d_counts is then incomplete and the caller falls back).
int radix_join_counts(const uint32_t *r_hi, const uint32_t *r_lo, uint64_t nr, const uint32_t *s_hi, const uint32_t *s_lo, uint64_t ns,
                      bool multi, uint32_t *d_counts, bool *overflowed, Arena &ws, cudaStream_t st) {
    *overflowed = false;
    if (!ns) return 0;
    if (nr >= (1ull << 32) || ns >= (1ull << 32)) {
        set_error("hash join: fewer than 2^32 rows per side");
        return DBT_ERR_UNSUPPORTED;
    }
    const size_t m0 = ws.mark();
    uint32_t bits = 0;
    while (bits < 20 && (nr >> bits) > kPartTargetRows) ++bits;
    const uint32_t nparts = 1u << bits;
    const uint32_t mask = bits ? (0xFFFFFFFFu << (32 - bits)) : 0u; // the partition id = the top bits of the mixed key
    const uint64_t nrr = std::max<uint64_t>(nr, 1);
    uint32_t *r_a = ws.take<uint32_t>(nrr), *r_b = ws.take<uint32_t>(nrr), *r_c = ws.take<uint32_t>(nrr), *r_d = ws.take<uint32_t>(nrr);
    uint32_t *s_a = ws.take<uint32_t>(ns), *s_b = ws.take<uint32_t>(ns), *s_c = ws.take<uint32_t>(ns), *s_d = ws.take<uint32_t>(ns);
    uint32_t *s_e = ws.take<uint32_t>(ns), *s_f = ws.take<uint32_t>(ns), *s_g = ws.take<uint32_t>(ns), *s_h = ws.take<uint32_t>(ns);
    uint32_t *r_start = ws.take<uint32_t>(nparts + 1), *s_start = ws.take<uint32_t>(nparts + 1), *d_over = ws.take<uint32_t>(64);
    if (!r_a || !r_b || !r_c || !r_d || !s_a || !s_b || !s_c || !s_d || !s_e || !s_f || !s_g || !s_h || !r_start || !s_start || !d_over) {
        set_error("hash join: workspace too small");
        return DBT_ERR_WORKSPACE;
    }
    {
        StageScope sc(ST_HASH_BUILD, st);
        if (nr) {
            const int g = (int)std::min<uint64_t>((nr + 255) / 256, 148 * 16);
            mix_keys_kernel<<<g, 256, 0, st>>>(r_hi, r_lo, nr, r_a, r_c, nullptr);
            count_launch();
        }
        const int g = (int)std::min<uint64_t>((ns + 255) / 256, 148 * 16);
        mix_keys_kernel<<<g, 256, 0, st>>>(s_hi, s_lo, ns, s_a, s_c, s_e);
        count_launch();
        DBT_KERNEL_CHECK();
    }
    // the partition passes: stable LSD over the partition bits; the pair that moves is (mixed hi, mixed lo) = the key itself.
    // S needs its row as well: a second sort of (mixed hi, row) with the same digits gives the same (stable) order.
    uint32_t *rk = r_a, *rka = r_b, *rv = r_c, *rva = r_d;
    if (nr) DBT_TRY(sort_pairs_masked(rk, rka, rv, rva, nr, mask, false, ws, st));
    uint32_t *sk = s_a, *ska = s_b, *sv = s_c, *sva = s_d;
    DBT_TRY(sort_pairs_masked(sk, ska, sv, sva, ns, mask, false, ws, st));
    uint32_t *sk2 = s_e, *sk2a = s_f, *srow = s_g, *srowa = s_h;
    DBT_TRY(sort_pairs_masked(sk2, sk2a, srow, srowa, ns, mask, true, ws, st));
    if (!mask) DBT_TRY(iota_u32(srow, ns, st)); // (a single partition: nothing was sorted, the row list is the identity)
    {
        StageScope sc(ST_HASH_BUILD, st);
        DBT_CUDA(cudaMemsetAsync(r_start, 0xFF, (size_t)(nparts + 1) * 4, st));
        DBT_CUDA(cudaMemsetAsync(s_start, 0xFF, (size_t)(nparts + 1) * 4, st));
        const int gr = (int)std::min<uint64_t>((nrr + 255) / 256, 148 * 16), gs = (int)std::min<uint64_t>((ns + 255) / 256, 148 * 16);
        part_bounds_kernel<<<gr, 256, 0, st>>>(rk, nr, bits, r_start);
        part_bounds_kernel<<<gs, 256, 0, st>>>(sk, ns, bits, s_start);
        part_fill_kernel<<<(nparts + 255) / 256, 256, 0, st>>>(r_start, nparts);
        part_fill_kernel<<<(nparts + 255) / 256, 256, 0, st>>>(s_start, nparts);
        count_launch(4);
        DBT_KERNEL_CHECK();
    }
    const size_t smem = (size_t)kPartSlots * 12;
    {
        StageScope sc(ST_HASH_PROBE, st);
        DBT_CUDA(cudaMemsetAsync(d_over, 0, 4, st));
        const int grid = (int)std::min<uint32_t>(nparts, 148u * 2u * 8u);
        if (multi) {
            if (first_use_on_device((const void *)smem_join_kernel<true>))
                DBT_CUDA(cudaFuncSetAttribute(smem_join_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            smem_join_kernel<true><<<grid, 256, smem, st>>>(rk, rv, r_start, sk, sv, srow, s_start, nparts, d_counts, d_over);
        } else {
            if (first_use_on_device((const void *)smem_join_kernel<false>))
                DBT_CUDA(cudaFuncSetAttribute(smem_join_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            smem_join_kernel<false><<<grid, 256, smem, st>>>(rk, rv, r_start, sk, sv, srow, s_start, nparts, d_counts, d_over);
        }
        count_launch();
        DBT_KERNEL_CHECK();
    }
    uint32_t h_over = 0;
    DBT_CUDA(cudaMemcpyAsync(&h_over, d_over, 4, cudaMemcpyDeviceToHost, st));
    DBT_CUDA(cudaStreamSynchronize(st));
    *overflowed = h_over != 0;
    ws.release(m0);
    return 0;
}

// ---------------------------------------------------------------------------------------------
// Intersection of two sorted unique lists.
// ---------------------------------------------------------------------------------------------
struct SortedList { // i-th smallest key of a relation: either a sorted u32 column or rows + key columns
    const uint32_t *keys; // sorted unique u32 keys (1-word keys) or nullptr
    const uint32_t *rows; // sorted unique row ids
    KeyView kv;
    uint64_t n;
};
__device__ __forceinline__ int list_cmp(const SortedList &a, uint64_t i, const SortedList &b, uint64_t j) {
    if (a.keys) {
        uint32_t x = a.keys[i], y = b.keys[j];
        return x < y ? -1 : (x > y ? 1 : 0);
    }
    return rows_cmp(a.kv, a.rows[i], b.kv, b.rows[j]);
}
// first j in [0, b.n) with b[j] >= a[i]
__device__ __forceinline__ uint64_t lower_bound(const SortedList &a, uint64_t i, const SortedList &b) {
    uint64_t lo = 0, hi = b.n;
    while (lo < hi) {
        uint64_t mid = (lo + hi) >> 1;
        if (list_cmp(b, mid, a, i) < 0) lo = mid + 1;
        else hi = mid;
    }
    return lo;
}

__global__ void __launch_bounds__(256) intersect_kernel(SortedList a, SortedList b, uint32_t *__restrict__ flags) {
    uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < a.n; i += stride) {
        uint64_t j = lower_bound(a, i, b);
        flags[i] = (j < b.n && list_cmp(a, i, b, j) == 0) ? 1u : 0u;
    }
}

// Number of block reads after the first two that the reference's two-pointer walk performs
// (DatabaseProject.cpp:414-482): the walk stops at the first exhausted list; R is tested before S.
// Closed form derived in DESIGN.md ("MergeJoin nios").
__global__ void walk_reads_kernel(SortedList r, SortedList s, unsigned long long *out) {
    if (threadIdx.x || blockIdx.x) return;
    uint64_t nr = r.n, ns = s.n, reads = 0;
    if (nr && ns) {
        int c = list_cmp(r, nr - 1, s, ns - 1);
        uint64_t a, b;
        bool inc_b;
        if (c <= 0) { // R runs out first (or both together): the empty R read ends the walk
            a = nr;
            uint64_t lb = lower_bound(r, nr - 1, s);
            bool match = lb < ns && list_cmp(r, nr - 1, s, lb) == 0;
            b = lb + (match ? 1 : 0);
            inc_b = match;
            uint64_t ymax = b - (inc_b ? 1 : 0);
            reads = (nr + kRpb - 1) / kRpb + ymax / kRpb;
        } else { // S runs out first
            b = ns;
            uint64_t lb = lower_bound(s, ns - 1, r);
            bool match = lb < nr && list_cmp(s, ns - 1, r, lb) == 0;
            a = lb + (match ? 1 : 0);
            reads = a / kRpb + (ns + kRpb - 1) / kRpb;
        }
    }
    *out = reads;
}

// ---------------------------------------------------------------------------------------------
// Merge-path intersection of two sorted unique key arrays A[na][kw], B[nb][kw] (contiguous).
// The merge path (A before B on ties) is cut into tiles of kMpTile merged elements by one binary
// search per tile; a CTA then holds its tile's B range (+1 element) in shared memory and every A
// element of the tile finds its equal, if any, by a short search there.  Both inputs are read once,
// sequentially: (kw*4) bytes per element instead of ~10 random 128-byte lines per binary search.
// ---------------------------------------------------------------------------------------------
constexpr int kMpTile = 1024;
constexpr int kMpThreads = 256;

__device__ __forceinline__ int key_cmp(const uint32_t *a, const uint32_t *b, uint32_t kw) {
    for (uint32_t j = 0; j < kw; ++j)
        if (a[j] != b[j]) return a[j] < b[j] ? -1 : 1;
    return 0;
}

__global__ void __launch_bounds__(256)
mp_partition_kernel(const uint32_t *__restrict__ A, uint64_t na, const uint32_t *__restrict__ B, uint64_t nb, uint32_t kw,
                    uint64_t ntiles, unsigned long long *__restrict__ a_start /*[ntiles+1]*/) {
    uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t > ntiles) return;
    uint64_t d = min(t * (uint64_t)kMpTile, na + nb);
    uint64_t lo = d > nb ? d - nb : 0, hi = min(d, na);
    while (lo < hi) {
        uint64_t mid = (lo + hi) >> 1; // candidate: mid elements of A, d-mid of B
        if (key_cmp(A + mid * kw, B + (d - 1 - mid) * kw, kw) <= 0) lo = mid + 1; // A[mid] precedes B[d-1-mid]: take more of A
        else hi = mid;
    }
    a_start[t] = lo;
}

__global__ void __launch_bounds__(kMpThreads)
mp_intersect_kernel(const uint32_t *__restrict__ A, uint64_t na, const uint32_t *__restrict__ B, uint64_t nb, uint32_t kw,
                    const unsigned long long *__restrict__ a_start, uint32_t *__restrict__ flags) {
    extern __shared__ uint32_t sb[]; // (kMpTile + 1) * kw words
    const uint64_t t = blockIdx.x;
    const uint64_t a0 = a_start[t], a1 = a_start[t + 1];
    const uint64_t d0 = min(t * (uint64_t)kMpTile, na + nb), d1 = min((t + 1) * (uint64_t)kMpTile, na + nb);
    const uint64_t b0 = d0 - a0, b1 = min(d1 - a1 + 1, nb); // one extra B element: the equal of the tile's last A may be there
    const uint32_t nbw = (uint32_t)((b1 > b0 ? b1 - b0 : 0) * kw);
    for (uint32_t i = threadIdx.x; i < nbw; i += kMpThreads) sb[i] = B[b0 * kw + i];
    __syncthreads();
    const uint32_t cntb = nbw / kw;
    for (uint64_t i = a0 + threadIdx.x; i < a1; i += kMpThreads) {
        const uint32_t *ka = A + i * kw;
        uint32_t lo = 0, hi = cntb;
        while (lo < hi) {
            uint32_t mid = (lo + hi) >> 1;
            if (key_cmp(sb + mid * kw, ka, kw) < 0) lo = mid + 1;
            else hi = mid;
        }
        flags[i] = (lo < cntb && key_cmp(sb + lo * kw, ka, kw) == 0) ? 1u : 0u;
    }
}

// contiguous copy of the keys of a sorted row list: out[i] = (w0[row], str[row][0..kw))
__global__ void __launch_bounds__(256)
gather_keys_kernel(KeyView kv, const uint32_t *__restrict__ rows, uint64_t n, uint32_t kwt, uint32_t *__restrict__ out) {
    uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const uint32_t row = rows[i];
        uint32_t *o = out + i * kwt;
        uint32_t j = 0;
        if (kv.w0) o[j++] = kv.w0[row];
        if (kv.str) {
            const uint32_t *p = kv.str + (uint64_t)row * kv.kw;
            for (uint32_t q = 0; q < kv.kw; ++q) o[j + q] = p[q];
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Inner-join pairs: for every S row the range of equal keys in the sorted R list.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
match_range_kernel(SortedList r, KeyView sv, uint64_t ns, uint32_t *__restrict__ first, uint32_t *__restrict__ count) {
    uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t q = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; q < ns; q += stride) {
        // lower and upper bound of S row q's key in the sorted R list
        auto cmp_at = [&](uint64_t i) -> int { // key(R_sorted[i]) vs key(S[q])
            if (r.keys) {
                uint32_t x = r.keys[i], y = sv.w0[q];
                return x < y ? -1 : (x > y ? 1 : 0);
            }
            return rows_cmp(r.kv, r.rows[i], sv, (uint32_t)q);
        };
        uint64_t lo = 0, hi = r.n;
        while (lo < hi) {
            uint64_t mid = (lo + hi) >> 1;
            if (cmp_at(mid) < 0) lo = mid + 1;
            else hi = mid;
        }
        uint64_t lb = lo;
        hi = r.n;
        while (lo < hi) {
            uint64_t mid = (lo + hi) >> 1;
            if (cmp_at(mid) <= 0) lo = mid + 1;
            else hi = mid;
        }
        first[q] = (uint32_t)lb;
        count[q] = (uint32_t)(lo - lb);
    }
}
int match_ranges(const KeyCols &r, const uint32_t *d_rperm, const uint32_t *d_rsorted_w0, const KeyCols &s, int field,
                 uint32_t *d_first, uint32_t *d_count, cudaStream_t st) {
    if (!s.n) return 0;
    StageScope sc(ST_HASH_PROBE, st);
    const bool one = (field == '0' || field == '1');
    SortedList rl{one ? d_rsorted_w0 : nullptr, d_rperm, KeyView{field == '2' ? nullptr : r.w0, r.str, r.kw}, r.n};
    KeyView sv{field == '2' ? nullptr : s.w0, s.str, s.kw};
    int grid = (int)std::min<uint64_t>((s.n + 255) / 256, 148 * 16);
    match_range_kernel<<<grid, 256, 0, st>>>(rl, sv, s.n, d_first, d_count);
    count_launch();
    DBT_KERNEL_CHECK();
    return 0;
}

int intersect_sorted(const KeyCols &r, const uint32_t *d_ur, const uint32_t *d_urkeys, uint64_t nur, const KeyCols &s,
                     const uint32_t *d_us, const uint32_t *d_uskeys, uint64_t nus, int field, uint32_t *d_flags,
                     uint64_t *d_later_reads, Arena &ws, cudaStream_t st, bool keys_are_contiguous) {
    StageScope sc(ST_INTERSECT, st);
    const bool one = (field == '0' || field == '1');
    SortedList a{one ? d_urkeys : nullptr, d_ur, KeyView{field == '2' ? nullptr : r.w0, r.str, r.kw}, nur};
    SortedList b{one ? d_uskeys : nullptr, d_us, KeyView{field == '2' ? nullptr : s.w0, s.str, s.kw}, nus};
    if (nur && nus == 0) {
        DBT_CUDA(cudaMemsetAsync(d_flags, 0, 4 * nur, st));
    } else if (nur) {
        const uint32_t kwt = one ? 1u : ((field == '2' ? 0u : 1u) + r.kw);
        const uint32_t *A = d_urkeys, *B = d_uskeys;
        size_t m0 = ws.mark();
        bool ok = true;
        if (!one && !keys_are_contiguous) { // multi-word keys: make the two sorted key lists contiguous first (one random 32-byte read per key)
            uint32_t *ca = ws.take<uint32_t>(nur * kwt), *cb = ws.take<uint32_t>(nus * kwt);
            if (!ca || !cb) ok = false;
            else {
                int ga = (int)std::min<uint64_t>((nur + 255) / 256, 148 * 16), gb2 = (int)std::min<uint64_t>((nus + 255) / 256, 148 * 16);
                gather_keys_kernel<<<ga, 256, 0, st>>>(a.kv, d_ur, nur, kwt, ca);
                gather_keys_kernel<<<gb2, 256, 0, st>>>(b.kv, d_us, nus, kwt, cb);
                count_launch(2);
                A = ca;
                B = cb;
            }
        }
        const uint64_t ntiles = (nur + nus + kMpTile - 1) / kMpTile;
        unsigned long long *a_start = ok ? ws.take<unsigned long long>(ntiles + 1) : nullptr;
        const size_t smem = (size_t)(kMpTile + 1) * kwt * 4;
        if (ok && a_start && smem <= 200 * 1024) {
            mp_partition_kernel<<<(unsigned)((ntiles + 1 + 255) / 256), 256, 0, st>>>(A, nur, B, nus, kwt, ntiles, a_start);
            if (first_use_on_device((const void *)mp_intersect_kernel))
                DBT_CUDA(cudaFuncSetAttribute(mp_intersect_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
            mp_intersect_kernel<<<(unsigned)ntiles, kMpThreads, smem, st>>>(A, nur, B, nus, kwt, a_start, d_flags);
            count_launch(2);
        } else { // not enough workspace for the contiguous copies: binary search through the row lists
            int grid = (int)std::min<uint64_t>((nur + 255) / 256, 148 * 16);
            intersect_kernel<<<grid, 256, 0, st>>>(a, b, d_flags);
            count_launch();
        }
        DBT_KERNEL_CHECK();
        ws.release(m0);
    }
    walk_reads_kernel<<<1, 1, 0, st>>>(a, b, (unsigned long long *)d_later_reads);
    count_launch();
    DBT_KERNEL_CHECK();
    return 0;
}

} // namespace dbt
