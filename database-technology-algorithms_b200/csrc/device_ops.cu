// device_ops.cu -- device-scope operators: block image in HBM -> block image in HBM.
//
// These are the GPU bodies of the four dbtproj.h operators (reference DatabaseProject.cpp:94-647).
// Pipeline shared by all of them:
//   headers -> key extraction (AoS -> columns) -> [LSD radix sort of (key word, row) pairs, one
//   word at a time from least to most significant, constant digits skipped] -> [unique | hash
//   build/probe | intersection] -> compaction -> ONE record gather into the output image.
#include "dbt_internal.cuh"
#include <algorithm>
#include <cstring>
#include <vector>

namespace dbt {

static inline bool is_aligned16(const void *p) { return ((uintptr_t)p & 15) == 0; }

static bool field_ok(int f) { return f >= '0' && f <= '3'; }

int prepare(const void *d_img, uint64_t nblocks, int field, Arena &ws, cudaStream_t st, Prepared *out, uint32_t force_kw) {
    memset(out, 0, sizeof *out);
    device_setup();
    DBT_TRY(image_info(d_img, nblocks, &out->row_slot, ws, st, &out->info));
    const uint64_t n = out->info.nrows;
    KeyCols &k = out->keys;
    k.n = n;
    k.kw = force_kw ? force_kw : 8;
    if (n == 0) return 0;
    const bool has_w0 = field != '2', has_str = field >= '2';
    ExtractStats *d_stats = ws.take<ExtractStats>(1);
    k.recid = ws.take<uint32_t>(n);
    k.w0 = has_w0 ? ws.take<uint32_t>(n) : nullptr;
    k.str = has_str ? ws.take<uint32_t>(n * k.kw) : nullptr;
    if (!d_stats || !k.recid || (has_w0 && !k.w0) || (has_str && !k.str)) {
        set_error("prepare: workspace too small");
        return (has_str && k.kw > 8) ? DBT_ERR_NEED_WIDE_KEYS : DBT_ERR_WORKSPACE;
    }
    ExtractStats h;
    uint32_t *d_byte_hist = has_w0 && !has_str ? ws.take<uint32_t>(4 * 256) : nullptr;
    int hist_done = 0;
    DBT_TRY(extract_keys(d_img, nblocks, n, out->row_slot, out->info.blk_nres, out->info.blk_row_off, field, k.kw, k.w0, k.str,
                         k.recid, d_stats, st, d_byte_hist, &hist_done));
    k.w0_byte_hist = hist_done ? d_byte_hist : nullptr;
    DBT_CUDA(cudaMemcpyAsync(&h, d_stats, sizeof h, cudaMemcpyDeviceToHost, st));
    DBT_CUDA(cudaStreamSynchronize(st));
    if (has_str && h.str_overflow) {
        // some str has no NUL in its first 32 bytes: fall back to the full 120-byte key
        k.kw = kStrWords;
        k.str = ws.take<uint32_t>(n * k.kw);
        if (!k.str) {
            set_error("prepare: workspace too small for 120-byte string keys (size it with dbt_dev_ws_bytes_kw(..., 30))");
            return DBT_ERR_NEED_WIDE_KEYS;
        }
        DBT_TRY(extract_keys(d_img, nblocks, n, out->row_slot, out->info.blk_nres, out->info.blk_row_off, field, k.kw, k.w0, k.str,
                         k.recid, d_stats, st));
        DBT_CUDA(cudaMemcpyAsync(&h, d_stats, sizeof h, cudaMemcpyDeviceToHost, st));
        DBT_CUDA(cudaStreamSynchronize(st));
    }
    k.vary_w0 = has_w0 ? (h.or_w0 ^ h.and_w0) : 0;
    k.vary_recid = h.or_recid ^ h.and_recid;
    for (uint32_t j = 0; j < 30; ++j) k.vary_str[j] = (has_str && j < k.kw) ? (h.str_or[j] ^ h.str_and[j]) : 0;
    k.recid_unsorted = (int)h.recid_unsorted;
    k.or_w0 = h.or_w0;
    k.and_w0 = h.and_w0;
    k.or_recid = h.or_recid;
    k.and_recid = h.and_recid;
    for (uint32_t j = 0; j < 30; ++j) {
        k.or_str[j] = h.str_or[j];
        k.and_str[j] = h.str_and[j];
    }
    return 0;
}

// ---- order-preserving key compaction --------------------------------------------------------------
// Bits that are constant over the whole relation cannot decide a comparison, so a multi-word key (str, num+str)
// is equivalent, for ordering and for equality WITHIN the relation, to the concatenation of its varying bits,
// most significant first.  Reference-like strings (5 letters) carry 31 varying bits in their 32 key bytes: one
// u32 instead of a multi-word LSD with a random word gather per extra word.  Used when at most 64 bits vary.
struct CompactPlan {
    uint32_t nruns;
    uint8_t word[64];  // 255 = w0, else the index of the str word
    uint8_t shift[64]; // lowest bit of the run inside its word
    uint8_t width[64]; // 1..32
    uint8_t out[64];   // lowest bit of the run inside the 64-bit compact key
};
__global__ void __launch_bounds__(256)
compact_keys_kernel(const uint32_t *__restrict__ w0, const uint32_t *__restrict__ str, uint32_t kw, uint64_t n, CompactPlan p,
                    uint32_t *__restrict__ lo, uint32_t *__restrict__ hi) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        unsigned long long key = 0;
        for (uint32_t r = 0; r < p.nruns; ++r) {
            const uint32_t w = (p.word[r] == 255) ? w0[i] : str[i * kw + p.word[r]];
            const uint32_t m = (p.width[r] >= 32) ? 0xFFFFFFFFu : ((1u << p.width[r]) - 1u);
            key |= (unsigned long long)((w >> p.shift[r]) & m) << p.out[r];
        }
        lo[i] = (uint32_t)key;
        if (hi) hi[i] = (uint32_t)(key >> 32);
    }
}
// Plan the runs for field '2' (str words) or '3' (w0 then str words); returns the number of varying bits (0 = do not compact).
static int plan_compaction(const KeyCols &k, int field, CompactPlan *p) {
    memset(p, 0, sizeof *p);
    struct Src {
        uint8_t word;
        uint32_t mask;
    };
    std::vector<Src> srcs; // most significant first
    if (field == '3' && k.w0) srcs.push_back({255, k.vary_w0});
    for (uint32_t j = 0; j < k.kw && j < 30; ++j) srcs.push_back({(uint8_t)j, k.vary_str[j]});
    int total = 0;
    for (const Src &s : srcs) total += __builtin_popcount(s.mask);
    if (total == 0 || total > 64) return 0;
    int out = total;
    for (const Src &s : srcs) {
        uint32_t m = s.mask;
        while (m) { // runs of consecutive varying bits, from the top
            const int hi = 31 - __builtin_clz(m);
            int lo = hi;
            while (lo > 0 && ((m >> (lo - 1)) & 1u)) --lo;
            const int width = hi - lo + 1;
            if (p->nruns >= 64) return 0;
            out -= width;
            p->word[p->nruns] = s.word;
            p->shift[p->nruns] = (uint8_t)lo;
            p->width[p->nruns] = (uint8_t)width;
            p->out[p->nruns] = (uint8_t)out;
            ++p->nruns;
            m &= (lo == 0) ? 0u : ((1u << lo) - 1u);
        }
    }
    return total;
}

// vary masks of the union of two relations (join sides), and the compaction plan they give (0 = do not compact)
static int joint_compaction(const KeyCols &a, const KeyCols &b, int field, CompactPlan *plan, KeyCols *masks) {
    *masks = a;
    masks->vary_w0 = (a.or_w0 | b.or_w0) ^ (a.and_w0 & b.and_w0);
    for (uint32_t w = 0; w < 30; ++w)
        masks->vary_str[w] = (w < a.kw) ? ((a.or_str[w] | b.or_str[w]) ^ (a.and_str[w] & b.and_str[w])) : 0u;
    if (getenv("DBT_NO_KEY_COMPACTION") != nullptr) return 0;
    return plan_compaction(*masks, field, plan);
}
static int launch_compaction(const KeyCols &k, int field, const CompactPlan &plan, uint32_t *lo, uint32_t *hi, cudaStream_t st) {
    StageScope sc(ST_WORD_GATHER, st);
    const int grid = (int)std::min<uint64_t>((k.n + 255) / 256, 148 * 16);
    compact_keys_kernel<<<grid, 256, 0, st>>>(field == '3' ? k.w0 : nullptr, k.str, k.kw, k.n, plan, lo, hi);
    count_launch();
    DBT_KERNEL_CHECK();
    return 0;
}

// Order the rows by (key(field), recid); ties beyond that keep file order (LSD passes are stable).
// Returns the row permutation and, for 1-word keys, the sorted key column.
int sort_rows_by_key(KeyCols &k, int field, Arena &ws, cudaStream_t st, uint32_t **perm_out, uint32_t **sorted_w0_out,
                     KeyCols *compact) {
    const uint64_t n = k.n;
    *perm_out = nullptr;
    *sorted_w0_out = nullptr;
    if (compact) memset(compact, 0, sizeof *compact);
    if (n == 0) return 0;
    struct Word {
        const uint32_t *src;
        uint32_t stride, idx, vary;
    };
    std::vector<Word> words; // least significant first
    if (field != '0' && k.recid_unsorted && k.vary_recid) words.push_back({k.recid, 1, 0, k.vary_recid});
    // single-relation callers (sort, dedup): a multi-word key whose varying bits fit 64 bits is sorted as those bits
    uint32_t *ck_lo = nullptr, *ck_hi = nullptr;
    int cbits = 0;
    if (compact && field >= '2' && getenv("DBT_NO_KEY_COMPACTION") == nullptr) {
        CompactPlan plan;
        cbits = plan_compaction(k, field, &plan);
        if (cbits) {
            ck_lo = ws.take<uint32_t>(n);
            ck_hi = cbits > 32 ? ws.take<uint32_t>(n) : nullptr;
            if (!ck_lo || (cbits > 32 && !ck_hi)) {
                set_error("sort: workspace too small");
                return DBT_ERR_WORKSPACE;
            }
            DBT_TRY(launch_compaction(k, field, plan, ck_lo, ck_hi, st));
            const int lo_bits = std::min(cbits, 32);
            words.push_back({ck_lo, 1, 0, lo_bits >= 32 ? 0xFFFFFFFFu : ((1u << lo_bits) - 1u)});
            if (ck_hi) words.push_back({ck_hi, 1, 0, (cbits - 32 >= 32) ? 0xFFFFFFFFu : ((1u << (cbits - 32)) - 1u)});
            // the view the caller's unique scan can use: equal compact keys <=> equal keys (within this relation)
            compact->n = n;
            compact->kw = 1;
            compact->w0 = ck_hi ? ck_hi : ck_lo;
            compact->str = ck_hi ? ck_lo : nullptr;
        }
    }
    if (field >= '2' && !cbits)
        for (int j = (int)k.kw - 1; j >= 0; --j)
            if (k.vary_str[j]) words.push_back({k.str, k.kw, (uint32_t)j, k.vary_str[j]});
    const bool w0_is_key = field != '2' && !cbits; // (field '3': num is inside the compact key)
    // field '0' with recids already ascending in file order: the file is sorted (stable) as it is
    const bool w0_sorted_already = (field == '0' && !k.recid_unsorted);
    if (w0_is_key && k.vary_w0 && !w0_sorted_already) words.push_back({k.w0, 1, 0, k.vary_w0});

    uint32_t *ka = nullptr, *kb = ws.take<uint32_t>(n), *va = ws.take<uint32_t>(n), *vb = ws.take<uint32_t>(n);
    // a single one-word key: sort the column in place (clobbers it)
    // (field '3' without a compact key keeps w0 intact: its callers compare full keys through the row permutation)
    const bool single_col = (words.size() == 1 && ((words[0].src == k.w0 && (field == '0' || field == '1')) || words[0].src == ck_lo));
    if (!single_col) ka = ws.take<uint32_t>(n);
    else ka = const_cast<uint32_t *>(words[0].src);
    if (!ka || !kb || !va || !vb) {
        set_error("sort: workspace too small");
        return DBT_ERR_WORKSPACE;
    }
    uint32_t *kk = ka, *kk_alt = kb, *vv = va, *vv_alt = vb;
    bool first = true;
    for (const Word &w : words) {
        if (!(first && single_col)) {
            StageScope sc(ST_WORD_GATHER, st);
            DBT_TRY(gather_word(w.src, w.stride, w.idx, first ? nullptr : vv, kk, n, st));
        }
        // the byte histogram made during extraction describes w0 in any order (a histogram is order-free)
        DBT_TRY(sort_pairs_masked(kk, kk_alt, vv, vv_alt, n, w.vary, first, ws, st, (w.src == k.w0) ? k.w0_byte_hist : nullptr));
        first = false;
    }
    if (first) DBT_TRY(iota_u32(vv, n, st)); // nothing varied: identity order
    *perm_out = vv;
    if (field == '0' || field == '1') {
        if (!words.empty() && words.back().src == k.w0) *sorted_w0_out = kk;
        else *sorted_w0_out = k.w0; // constant column, or already ascending: the column is its own sorted copy
    } else if (cbits && !ck_hi) {
        *sorted_w0_out = kk; // the sorted one-word compact key (the last word sorted)
        compact->sorted_lo = kk;
    } else if (cbits) { // two words: hi is the last word sorted; lo follows on demand (sorted_low_word)
        compact->sorted_hi = kk;
    }
    return 0;
}

// two-word compact keys: bring the low word into sorted order (one gather through the permutation) for the scans
static int sorted_low_word(KeyCols &view, const uint32_t *d_perm, Arena &ws, cudaStream_t st) {
    if (!view.n || !view.str || view.sorted_lo) return 0;
    uint32_t *slo = ws.take<uint32_t>(view.n);
    if (!slo) {
        set_error("sort: workspace too small");
        return DBT_ERR_WORKSPACE;
    }
    StageScope sc(ST_WORD_GATHER, st);
    DBT_TRY(gather_word(view.str, 1, 0, d_perm, slo, view.n, st));
    view.sorted_lo = slo;
    return 0;
}

int unique_sorted(const KeyCols &k, const KeyCols &compact, int field, const uint32_t *d_perm, const uint32_t *d_sorted_w0,
                  uint64_t n, uint32_t *d_uperm, uint64_t *d_count, Arena &ws, cudaStream_t st) {
    if (compact.n && !compact.str) // one-word compact key: its sorted column is at hand
        return unique_rows(compact, '1', d_perm, d_sorted_w0, n, d_uperm, nullptr, d_count, ws, st);
    if (compact.n) { // two words: both columns in sorted order
        KeyCols v = compact;
        DBT_TRY(sorted_low_word(v, d_perm, ws, st));
        return unique_rows_sorted2(v.sorted_hi, v.sorted_lo, d_perm, n, d_uperm, nullptr, d_count, ws, st);
    }
    return unique_rows(k, field, d_perm, d_sorted_w0, n, d_uperm, nullptr, d_count, ws, st);
}

// Both sides of a join: prepare R and S; when only one side has strings of 32+ bytes, the other side's keys are
// widened to the full 120 bytes as well -- its first (8-word) columns are released before it is prepared again.
static int prepare_pair(const void *d_r, uint64_t nbr, const void *d_s, uint64_t nbs, int field, Arena &ws, cudaStream_t st,
                        Prepared *pr, Prepared *ps) {
    const size_t m0 = ws.mark();
    DBT_TRY(prepare(d_r, nbr, field, ws, st, pr));
    const size_t m1 = ws.mark();
    DBT_TRY(prepare(d_s, nbs, field, ws, st, ps));
    if (field >= '2' && pr->keys.kw != ps->keys.kw && pr->info.nrows && ps->info.nrows) {
        const uint32_t kw = std::max(pr->keys.kw, ps->keys.kw);
        if (ps->keys.kw < kw) { // S is the narrow side: it was prepared last, redo it in place
            ws.release(m1);
            DBT_TRY(prepare(d_s, nbs, field, ws, st, ps, kw));
        } else { // R is the narrow side: redo both, wide from the start
            ws.release(m0);
            DBT_TRY(prepare(d_r, nbr, field, ws, st, pr, kw));
            DBT_TRY(prepare(d_s, nbs, field, ws, st, ps, kw));
        }
    }
    return 0;
}

static int read_u64(const uint64_t *d, uint64_t *h, int count, cudaStream_t st);

// The ordered (and optionally duplicate-free) row list of one image, without moving a record: what MergeSort /
// EliminateDuplicates gather.  The multi-GPU layer appends the lists of consecutive key sub-ranges to one output image.
int sorted_rows(const void *d_img, uint64_t nblocks, int field, bool dedup, Arena &ws, cudaStream_t st, uint32_t **rows,
                uint32_t **row_slot, uint64_t *n_in, uint64_t *n_out) {
    Prepared p;
    DBT_TRY(prepare(d_img, nblocks, field, ws, st, &p));
    const uint64_t n = p.info.nrows;
    *n_in = n;
    *n_out = 0;
    *rows = nullptr;
    *row_slot = p.row_slot;
    if (!n) return 0;
    uint32_t *perm, *sorted;
    KeyCols compact;
    DBT_TRY(sort_rows_by_key(p.keys, field, ws, st, &perm, &sorted, &compact));
    if (!dedup) {
        *rows = perm;
        *n_out = n;
        return 0;
    }
    uint32_t *uperm = ws.take<uint32_t>(n);
    uint64_t *d_cnt = ws.take<uint64_t>(8);
    if (!uperm || !d_cnt) {
        set_error("dedup: workspace too small");
        return DBT_ERR_WORKSPACE;
    }
    DBT_TRY(unique_sorted(p.keys, compact, field, perm, sorted, n, uperm, d_cnt, ws, st));
    DBT_TRY(read_u64(d_cnt, n_out, 1, st));
    *rows = uperm;
    return 0;
}

static int read_u64(const uint64_t *d, uint64_t *h, int count, cudaStream_t st) {
    DBT_CUDA(cudaMemcpyAsync(h, d, 8 * count, cudaMemcpyDeviceToHost, st));
    DBT_CUDA(cudaStreamSynchronize(st));
    return 0;
}

static int finish(cudaStream_t st) {
    DBT_CUDA(cudaStreamSynchronize(st));
    stage_resolve();
    return 0;
}

// ---- workspace bound ----------------------------------------------------------------------
static size_t rel_bytes(uint64_t nb, int field, uint32_t kw, bool sorted) {
    uint64_t n = nb * kRpb;
    size_t b = 0;
    b += 2 * pad256(4 * nb) + pad256(8 * (nb / 2048 + 2)) + 2048; // nreserved, block row offsets, scan state, stats
    b += pad256(4 * n);                              // ragged slot list
    b += 512 + pad256(4 * n);                        // stats + recid
    if (field != '2') b += pad256(4 * n);            // w0
    if (field >= '2') b += pad256(4 * n * 8) + (kw > 8 ? pad256(4 * n * kw) : 0) + (sorted ? 6 * pad256(4 * n) : 0); // + compact key words, lo in sorted order, positions, (hi, lo) pairs
    if (sorted) b += 4 * pad256(4 * n) + sort_ws_bytes(n) + 4096; // ping/pong keys+rows, look-back state
    return b;
}

} // namespace dbt

using namespace dbt;

extern "C" size_t dbt_dev_ws_bytes_kw(int op, uint64_t nbr, uint64_t nbs, int field, uint32_t kw) {
    uint64_t nr = nbr * kRpb, ns = nbs * kRpb;
    size_t scan = pad256(8 * ((std::max(nr, ns) + 2047) / 2048 + 1)) + 1024; // >= the scan's tile-state array
    size_t b = 1 << 20;
    switch (op) {
    case DBT_OP_SORT: b += rel_bytes(nbr, field, kw, true); break;
    case DBT_OP_DEDUP: b += rel_bytes(nbr, field, kw, true) + 2 * pad256(4 * nr) + scan + 512; break;
    case DBT_OP_MERGEJOIN:
        b += rel_bytes(nbr, field, kw, true) + rel_bytes(nbs, field, kw, true) + 2 * pad256(4 * nr) + 2 * pad256(4 * ns) +
             2 * pad256(4 * nr) + scan + 4096;
        if (field >= '2') b += pad256(4 * (kw + 1) * nr) + pad256(4 * (kw + 1) * ns); // contiguous sorted keys for the merge path
        b += pad256(8 * ((nr + ns) / 1024 + 2));
        break;
    case DBT_OP_HASHJOIN:
        b += rel_bytes(nbr, field, kw, false) + rel_bytes(nbs, field, kw, false) + 2 * pad256(4 * hash_table_slots(nr)) +
             2 * pad256(4 * ns) + scan + 4096;
        if (field != '3') b += (1ull << 29) + 4096 + pad256(4 * nr) + pad256(4 * ns); // the full-range key bitmap (512 MB); compact str keys
        if (field >= '2') b += 6 * pad256(4 * nr) + 10 * pad256(4 * ns) + 2 * sort_ws_bytes(std::max(nr, ns)) + (16u << 20); // radix partitioning: key words, mixed keys (ping/pong), rows, partition starts
        break;
    default: return 0;
    }
    return b + b / 16;
}
extern "C" size_t dbt_dev_hashjoin_ws_bytes(uint64_t nbr, uint64_t nbs, int field, uint32_t kw, uint64_t out_capacity_blocks) {
    size_t b = dbt_dev_ws_bytes_kw(DBT_OP_HASHJOIN, nbr, nbs, field, kw);
    if (field == '3' && out_capacity_blocks > nbs) b += pad256(4 * (out_capacity_blocks - nbs) * kRpb) + 4096; // the emitted-row list
    return b;
}
extern "C" size_t dbt_dev_ws_bytes(int op, uint64_t nbr, uint64_t nbs, int field) {
    return dbt_dev_ws_bytes_kw(op, nbr, nbs, field, 8);
}

static inline bool aligned16(const void *p) { return ((uintptr_t)p & 15) == 0; }
#define DBT_CHECK_ALIGNED(...)                                                                        \
    do {                                                                                              \
        const void *ptrs__[] = {__VA_ARGS__};                                                         \
        for (const void *q__ : ptrs__)                                                                \
            if (!aligned16(q__)) {                                                                    \
                set_error("device buffers (images, columns, workspace) must be 16-byte aligned");    \
                return DBT_ERR_ARG;                                                                   \
            }                                                                                         \
    } while (0)

#define DBT_CHECK_ARGS(cond, msg)  \
    do {                           \
        if (!(cond)) {             \
            set_error(msg);        \
            return DBT_ERR_ARG;    \
        }                          \
    } while (0)

// ---- multi-GPU building blocks ---------------------------------------------------------------
namespace dbt {
__device__ __forceinline__ uint32_t part_hash(uint32_t k) { // must not correlate with the join table's hash
    k ^= k >> 15;
    k *= 0x2C1B3C6Du;
    k ^= k >> 12;
    k *= 0x297A2D39u;
    k ^= k >> 15;
    return k;
}
struct Splitters {
    uint32_t v[64];
};
__global__ void __launch_bounds__(256)
dest_kernel(const uint32_t *__restrict__ keys, uint64_t n, int mode, Splitters sp, uint32_t nparts,
            uint32_t *__restrict__ dest) {
    uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        uint32_t k = keys[i], d = 0;
        if (mode == 0) {
            for (uint32_t j = 0; j + 1 < nparts; ++j) d += (sp.v[j] <= k) ? 1u : 0u; // <= 63 compares, branch-free
        } else {
            d = part_hash(k) % nparts;
        }
        dest[i] = d;
    }
}
__global__ void __launch_bounds__(256) dest_count_kernel(const uint32_t *__restrict__ dest, uint64_t n,
                                                         unsigned long long *counts /*[64]*/) {
    __shared__ uint32_t sh[64];
    if (threadIdx.x < 64) sh[threadIdx.x] = 0;
    __syncthreads();
    uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        uint32_t d = dest[i];
        uint32_t peers = __match_any_sync(__activemask(), d); // few distinct values per warp: cheap here
        if ((int)(threadIdx.x & 31) == __ffs(peers) - 1) atomicAdd(&sh[d], __popc(peers));
    }
    __syncthreads();
    if (threadIdx.x < 64 && sh[threadIdx.x]) atomicAdd(&counts[threadIdx.x], (unsigned long long)sh[threadIdx.x]);
}
} // namespace dbt

extern "C" size_t dbt_dev_partition_ws_bytes(uint64_t nblocks) {
    uint64_t n = nblocks * kRpb;
    return 4 * pad256(4 * n) + sort_ws_bytes(n) + (1 << 20);
}

extern "C" int dbt_dev_extract_keys_u32(const void *d_in, uint64_t nblocks, int field, uint32_t *d_keys, void *d_ws,
                                        size_t ws_bytes, void *stream, uint64_t *nrows) {
    // The routing word: the most significant word of the key.  Equal keys share it, so routing on it (range or
    // hash) brings equal keys to the same GPU for every field: recid ('0'), num ('1' and '3'), the first four
    // bytes of str ('2').
    if (!field_ok(field)) {
        set_error("Wrong field! Please give a field between 0 and 3!");
        return DBT_ERR_ARG;
    }
    cudaStream_t st = (cudaStream_t)stream;
    Arena ws(d_ws, ws_bytes);
    Prepared p;
    DBT_TRY(prepare(d_in, nblocks, field, ws, st, &p));
    if (p.info.nrows) {
        if (field == '2') DBT_TRY(gather_word(p.keys.str, p.keys.kw, 0, nullptr, d_keys, p.info.nrows, st));
        else DBT_CUDA(cudaMemcpyAsync(d_keys, p.keys.w0, 4 * p.info.nrows, cudaMemcpyDeviceToDevice, st));
    }
    if (nrows) *nrows = p.info.nrows;
    return finish(st);
}

extern "C" int dbt_dev_partition_rows(const uint32_t *d_keys, uint64_t n, int mode, const uint32_t *h_splitters,
                                      uint32_t nparts, uint32_t *d_rows_grouped, uint64_t *h_counts, void *d_ws,
                                      size_t ws_bytes, void *stream) {
    if (nparts == 0 || nparts > 64 || (mode != 0 && mode != 1) || !h_counts) {
        set_error("dbt_dev_partition_rows: bad arguments (1 <= nparts <= 64)");
        return DBT_ERR_ARG;
    }
    if (!is_aligned16(d_keys) || !is_aligned16(d_rows_grouped) || !is_aligned16(d_ws)) {
        set_error("dbt_dev_partition_rows: buffers must be 16-byte aligned");
        return DBT_ERR_ARG;
    }
    cudaStream_t st = (cudaStream_t)stream;
    Arena ws(d_ws, ws_bytes);
    for (uint32_t i = 0; i < nparts; ++i) h_counts[i] = 0;
    if (n == 0) return 0;
    uint32_t *dest = ws.take<uint32_t>(n), *dest_alt = ws.take<uint32_t>(n), *rows = ws.take<uint32_t>(n);
    unsigned long long *d_counts = ws.take<unsigned long long>(64);
    if (!dest || !dest_alt || !rows || !d_counts) {
        set_error("dbt_dev_partition_rows: workspace too small");
        return DBT_ERR_WORKSPACE;
    }
    Splitters sp;
    memset(&sp, 0, sizeof sp);
    if (mode == 0)
        for (uint32_t j = 0; j + 1 < nparts; ++j) sp.v[j] = h_splitters[j];
    {
        StageScope sc(ST_MISC, st);
        int grid = (int)std::min<uint64_t>((n + 255) / 256, 148 * 16);
        DBT_CUDA(cudaMemsetAsync(d_counts, 0, 64 * 8, st));
        dest_kernel<<<grid, 256, 0, st>>>(d_keys, n, mode, sp, nparts, dest);
        dest_count_kernel<<<grid, 256, 0, st>>>(dest, n, d_counts);
        count_launch(2);
        DBT_KERNEL_CHECK();
    }
    uint32_t mask = 0;
    while ((1u << __builtin_popcount(mask)) < nparts) mask = (mask << 1) | 1u; // bits needed for nparts-1
    uint32_t *k = dest, *ka = dest_alt, *v = rows, *va = d_rows_grouped;
    DBT_TRY(sort_pairs_masked(k, ka, v, va, n, mask, true, ws, st)); // one stable pass on the destination id
    if (v != d_rows_grouped) DBT_CUDA(cudaMemcpyAsync(d_rows_grouped, v, 4 * n, cudaMemcpyDeviceToDevice, st));
    unsigned long long hc[64];
    DBT_CUDA(cudaMemcpyAsync(hc, d_counts, sizeof hc, cudaMemcpyDeviceToHost, st));
    DBT_TRY(finish(st));
    for (uint32_t i = 0; i < nparts; ++i) h_counts[i] = hc[i];
    return 0;
}

extern "C" int dbt_dev_mergesort(const void *d_in, uint64_t nblocks, int field, void *d_out, void *d_ws, size_t ws_bytes,
                                 void *stream, uint64_t *nrows) {
    DBT_CHECK_ARGS(field_ok(field), "Wrong field! Please give a field between 0 and 3!");
    DBT_CHECK_ARGS((d_in && d_out && d_ws) || nblocks == 0, "dbt_dev_mergesort: NULL buffer");
    DBT_CHECK_ALIGNED(d_in, d_out, d_ws);
    cudaStream_t st = (cudaStream_t)stream;
    Arena ws(d_ws, ws_bytes);
    Prepared p;
    DBT_TRY(prepare(d_in, nblocks, field, ws, st, &p));
    uint32_t *perm, *sorted;
    KeyCols compact;
    DBT_TRY(sort_rows_by_key(p.keys, field, ws, st, &perm, &sorted, &compact));
    DBT_TRY(gather_records(d_in, perm, p.row_slot, p.info.nrows, d_out, st));
    if (nrows) *nrows = p.info.nrows;
    return finish(st);
}

extern "C" int dbt_dev_dedup(const void *d_in, uint64_t nblocks, int field, void *d_out, void *d_ws, size_t ws_bytes,
                             void *stream, uint64_t *nrows, uint64_t *nunique) {
    DBT_CHECK_ARGS(field_ok(field), "Wrong field! Please give a field between 0 and 3!");
    DBT_CHECK_ARGS((d_in && d_out && d_ws) || nblocks == 0, "dbt_dev_dedup: NULL buffer");
    DBT_CHECK_ALIGNED(d_in, d_out, d_ws);
    cudaStream_t st = (cudaStream_t)stream;
    Arena ws(d_ws, ws_bytes);
    Prepared p;
    DBT_TRY(prepare(d_in, nblocks, field, ws, st, &p));
    const uint64_t n = p.info.nrows;
    uint64_t u = 0;
    if (n) {
        uint32_t *perm, *sorted;
        KeyCols compact;
        DBT_TRY(sort_rows_by_key(p.keys, field, ws, st, &perm, &sorted, &compact));
        uint32_t *uperm = ws.take<uint32_t>(n);
        uint64_t *d_cnt = ws.take<uint64_t>(8);
        if (!uperm || !d_cnt) {
            set_error("dedup: workspace too small");
            return DBT_ERR_WORKSPACE;
        }
        DBT_TRY(unique_sorted(p.keys, compact, field, perm, sorted, n, uperm, d_cnt, ws, st));
        DBT_TRY(read_u64(d_cnt, &u, 1, st));
        DBT_TRY(gather_records(d_in, uperm, p.row_slot, u, d_out, st));
    }
    if (nrows) *nrows = n;
    if (nunique) *nunique = u;
    return finish(st);
}

// sort + unique of one prepared relation; leaves the unique row list (and the unique key column for one-word keys).
// compact_ok: the caller has given both join sides the same (joint) vary masks, so both compact their keys with the
// same plan and `view` (n != 0) describes keys that are comparable across the two relations.
static int dedup_prepared(Prepared *p, int field, Arena &ws, cudaStream_t st, bool compact_ok, uint32_t **urows,
                          uint32_t **ukeys, uint64_t *nu, KeyCols *view) {
    const uint64_t n = p->info.nrows;
    *urows = *ukeys = nullptr;
    *nu = 0;
    memset(view, 0, sizeof *view);
    if (!n) return 0;
    uint32_t *perm, *sorted;
    DBT_TRY(sort_rows_by_key(p->keys, field, ws, st, &perm, &sorted, compact_ok ? view : nullptr));
    *urows = ws.take<uint32_t>(n);
    const bool one = (field == '0' || field == '1') || (view->n && !view->str);
    const bool two = view->n && view->str; // two-word compact key: unique keys come out as contiguous (hi, lo) pairs
    *ukeys = one ? ws.take<uint32_t>(n) : (two ? ws.take<uint32_t>(2 * n) : nullptr);
    uint32_t *upos = two ? ws.take<uint32_t>(n) : nullptr;
    uint64_t *d_cnt = ws.take<uint64_t>(8);
    if (!*urows || ((one || two) && !*ukeys) || (two && !upos) || !d_cnt) {
        set_error("mergejoin: workspace too small");
        return DBT_ERR_WORKSPACE;
    }
    if (two) {
        DBT_TRY(sorted_low_word(*view, perm, ws, st));
        DBT_TRY(unique_rows_sorted2(view->sorted_hi, view->sorted_lo, perm, n, *urows, upos, d_cnt, ws, st));
    }
    else if (view->n) DBT_TRY(unique_rows(*view, '1', perm, sorted, n, *urows, *ukeys, d_cnt, ws, st));
    else DBT_TRY(unique_rows(p->keys, field, perm, sorted, n, *urows, *ukeys, d_cnt, ws, st));
    DBT_TRY(read_u64(d_cnt, nu, 1, st));
    if (two) DBT_TRY(take_pairs(view->sorted_hi, view->sorted_lo, upos, *nu, *ukeys, st));
    return 0;
}

extern "C" int dbt_dev_mergejoin(const void *d_in_r, uint64_t nbr, const void *d_in_s, uint64_t nbs, int field,
                                 void *d_out_ur, void *d_out_us, void *d_out, void *d_ws, size_t ws_bytes, void *stream,
                                 uint64_t *res) {
    DBT_CHECK_ARGS(field_ok(field), "Wrong field! Please give a field between 0 and 3!");
    DBT_CHECK_ARGS(d_ws && res, "dbt_dev_mergejoin: NULL buffer");
    DBT_CHECK_ALIGNED(d_in_r, d_in_s, d_out_ur, d_out_us, d_out, d_ws);
    cudaStream_t st = (cudaStream_t)stream;
    Arena ws(d_ws, ws_bytes);
    Prepared pr, ps;
    uint32_t *ur, *urk, *us, *usk;
    uint64_t nur, nus;
    DBT_TRY(prepare_pair(d_in_r, nbr, d_in_s, nbs, field, ws, st, &pr, &ps));
    // Key compaction across the two relations: with the vary masks of their UNION, both sides drop the same constant
    // bits, so the compact keys keep the joint order and compare across R and S.
    bool joint = false;
    if (field >= '2' && pr.info.nrows && ps.info.nrows) {
        KeyCols j;
        CompactPlan plan;
        if (joint_compaction(pr.keys, ps.keys, field, &plan, &j)) {
            joint = true;
            for (KeyCols *k : {&pr.keys, &ps.keys}) {
                k->vary_w0 = j.vary_w0;
                for (uint32_t w = 0; w < 30; ++w) k->vary_str[w] = j.vary_str[w];
            }
        }
    }
    KeyCols vr, vs;
    DBT_TRY(dedup_prepared(&pr, field, ws, st, joint, &ur, &urk, &nur, &vr));
    DBT_TRY(dedup_prepared(&ps, field, ws, st, joint, &us, &usk, &nus, &vs));
    const bool cj = vr.n && vs.n; // both sides made compact keys (same plan)
    // the side files "1outfile.bin" / "2outfile.bin" (DatabaseProject.cpp:385-394)
    if (d_out_ur) DBT_TRY(gather_records(d_in_r, ur, pr.row_slot, nur, d_out_ur, st));
    if (d_out_us) DBT_TRY(gather_records(d_in_s, us, ps.row_slot, nus, d_out_us, st));
    uint64_t nres = 0, reads = 0;
    if (nur && nus) {
        uint32_t *flags = ws.take<uint32_t>(nur);
        uint32_t *mrows = ws.take<uint32_t>(nur);
        uint64_t *d_res = ws.take<uint64_t>(8);
        if (!flags || !mrows || !d_res) {
            set_error("mergejoin: workspace too small");
            return DBT_ERR_WORKSPACE;
        }
        DBT_TRY(intersect_sorted(cj ? vr : pr.keys, ur, urk, nur, cj ? vs : ps.keys, us, usk, nus,
                                 cj ? (vr.str ? '3' : '1') : field, flags, d_res + 1, ws, st, cj && vr.str));
        DBT_TRY(compact_select(flags, ur, nur, mrows, nur, d_res, ws, st));
        uint64_t h[2];
        DBT_TRY(read_u64(d_res, h, 2, st));
        nres = h[0];
        reads = h[1];
        if (d_out) DBT_TRY(gather_records(d_in_r, mrows, pr.row_slot, nres, d_out, st));
    }
    res[0] = nres;
    res[1] = nur;
    res[2] = nus;
    res[3] = reads;
    return finish(st);
}

// Fields '0'/'1' with a key bitmap at hand: the probe phase as streaming passes over the S image -- no extraction of S,
// no row lists, no random gather (kernels_semijoin.cu).  DBT_JOIN_FUSED: 2 (default) = two streaming passes (count, then
// copy), 1 = one chained pass (exact, but latency-bound: profiles/r02_notes.md), 0 = the column path (extract, probe,
// compact, gather).  *done = false when the bitmap did not fit the workspace (the caller then runs the column path).
static int fused_semijoin(const uint32_t *d_rkeys, uint64_t nr, const void *d_in_s, uint64_t nbs, int field, void *d_out,
                          uint64_t cap_blocks, Arena &ws, cudaStream_t st, uint64_t *nres, bool *done) {
    *done = false;
    const int mode = getenv("DBT_JOIN_FUSED") ? atoi(getenv("DBT_JOIN_FUSED")) : 2;
    if (mode == 0) return 0;
    if (!nr || !nbs || nbs >= (1ull << 32)) return 0;
    const size_t m0 = ws.mark();
    uint32_t *bm, base, span;
    DBT_TRY(build_key_bitmap(d_rkeys, nr, ws, st, &bm, &base, &span));
    uint64_t *d_total = ws.take<uint64_t>(8);
    if (!bm || !d_total) {
        ws.release(m0);
        return 0;
    }
    uint64_t h[2] = {0, 0};
    if (mode == 1) {
        DBT_TRY(semijoin_stream(d_in_s, nbs, field, bm, base, span, d_out, cap_blocks * kRpb, d_total, ws, st));
        DBT_TRY(read_u64(d_total, h, 2, st));
    } else {
        DBT_TRY(semijoin_two_pass(d_in_s, nbs, field, bm, base, span, d_out, cap_blocks * kRpb, d_total, &h[0], ws, st));
    }
    *done = true;
    *nres = h[0];
    if ((uint32_t)h[1]) {
        set_error("semijoin: the look-back chain timed out (internal error)");
        return DBT_ERR_CUDA;
    }
    if (h[0] > cap_blocks * kRpb) {
        set_error("hashjoin: output capacity too small (nres returned)");
        return DBT_ERR_WORKSPACE;
    }
    return 0;
}

extern "C" int dbt_dev_semijoin_keys(const uint32_t *d_rkeys, uint64_t nr, const void *d_in_s, uint64_t nbs, int field,
                                     void *d_out, uint64_t out_capacity_blocks, void *d_ws, size_t ws_bytes, void *stream,
                                     uint64_t *nres) {
    DBT_CHECK_ARGS(field == '0' || field == '1', "dbt_dev_semijoin_keys: u32 keys only (fields '0' and '1')");
    DBT_CHECK_ARGS(d_ws && nres && (d_rkeys || nr == 0), "dbt_dev_semijoin_keys: NULL buffer");
    DBT_CHECK_ALIGNED(d_rkeys, d_in_s, d_out, d_ws);
    cudaStream_t st = (cudaStream_t)stream;
    Arena ws(d_ws, ws_bytes);
    *nres = 0;
    {
        bool done = false;
        DBT_TRY(fused_semijoin(d_rkeys, nr, d_in_s, nbs, field, d_out, out_capacity_blocks, ws, st, nres, &done));
        if (done) return finish(st);
    }
    Prepared ps;
    DBT_TRY(prepare(d_in_s, nbs, field, ws, st, &ps));
    const uint64_t ns = ps.info.nrows;
    if (ns && nr) {
        KeyCols rk;
        memset(&rk, 0, sizeof rk);
        rk.w0 = const_cast<uint32_t *>(d_rkeys);
        rk.n = nr;
        rk.kw = 8;
        uint32_t *counts = ws.take<uint32_t>(ns);
        const uint64_t cap = out_capacity_blocks * kRpb;
        const uint64_t rows_cap = std::max<uint64_t>(std::min<uint64_t>(cap, ns), 1);
        uint32_t *rows = ws.take<uint32_t>(rows_cap);
        uint64_t *d_total = ws.take<uint64_t>(8);
        if (!counts || !rows || !d_total) {
            set_error("semijoin: workspace too small");
            return DBT_ERR_WORKSPACE;
        }
        DBT_TRY(hash_join_counts(rk, ps.keys, field, counts, ws, st));
        DBT_TRY(compact_select(counts, nullptr, ns, rows, rows_cap, d_total, ws, st));
        uint64_t total = 0;
        DBT_TRY(read_u64(d_total, &total, 1, st));
        *nres = total;
        if (total > cap) {
            set_error("semijoin: output capacity too small (nres returned)");
            return DBT_ERR_WORKSPACE;
        }
        DBT_TRY(gather_records(d_in_s, rows, ps.row_slot, total, d_out, st));
    }
    return finish(st);
}

extern "C" int dbt_dev_innerjoin_pairs(const void *d_in_r, uint64_t nbr, const void *d_in_s, uint64_t nbs, int field,
                                       uint32_t *d_pairs, uint64_t pairs_capacity, void *d_ws, size_t ws_bytes, void *stream,
                                       uint64_t *npairs) {
    DBT_CHECK_ARGS(field_ok(field), "Wrong field! Please give a field between 0 and 3!");
    DBT_CHECK_ARGS(d_ws && npairs, "dbt_dev_innerjoin_pairs: NULL buffer");
    DBT_CHECK_ALIGNED(d_in_r, d_in_s, d_pairs, d_ws);
    cudaStream_t st = (cudaStream_t)stream;
    Arena ws(d_ws, ws_bytes);
    Prepared pr, ps;
    DBT_TRY(prepare_pair(d_in_r, nbr, d_in_s, nbs, field, ws, st, &pr, &ps));
    *npairs = 0;
    const uint64_t nr = pr.info.nrows, ns = ps.info.nrows;
    if (nr && ns) {
        // the sort may clobber R's key column (1-word keys): keep the recid column, it is what the pairs carry
        uint32_t *rperm, *rsorted;
        DBT_TRY(sort_rows_by_key(pr.keys, field, ws, st, &rperm, &rsorted));
        uint32_t *first = ws.take<uint32_t>(ns), *count = ws.take<uint32_t>(ns);
        uint64_t *d_total = ws.take<uint64_t>(8);
        if (!first || !count || !d_total) {
            set_error("innerjoin: workspace too small");
            return DBT_ERR_WORKSPACE;
        }
        DBT_TRY(match_ranges(pr.keys, rperm, rsorted, ps.keys, field, first, count, st));
        DBT_TRY(expand_pairs(count, first, ns, rperm, pr.keys.recid, ps.keys.recid, d_pairs, d_pairs ? pairs_capacity : 0, d_total,
                             ws, st));
        uint64_t total = 0;
        DBT_TRY(read_u64(d_total, &total, 1, st));
        *npairs = total;
        if (total > pairs_capacity) {
            set_error("innerjoin: pair capacity too small (npairs returned)");
            DBT_TRY(finish(st));
            return DBT_ERR_WORKSPACE;
        }
    }
    return finish(st);
}

extern "C" int dbt_dev_hashjoin(const void *d_in_r, uint64_t nbr, const void *d_in_s, uint64_t nbs, int field,
                                void *d_out, uint64_t out_capacity_blocks, void *d_ws, size_t ws_bytes, void *stream,
                                uint64_t *nres) {
    DBT_CHECK_ARGS(field_ok(field), "Wrong field! Please give a field between 0 and 3!");
    DBT_CHECK_ARGS(d_ws && nres, "dbt_dev_hashjoin: NULL buffer");
    DBT_CHECK_ALIGNED(d_in_r, d_in_s, d_out, d_ws);
    cudaStream_t st = (cudaStream_t)stream;
    Arena ws(d_ws, ws_bytes);
    Prepared pr, ps;
    *nres = 0;
    if (field == '0' || field == '1') { // u32 keys: R's key column -> bitmap, then one streaming pass over S
        const size_t m0 = ws.mark();
        DBT_TRY(prepare(d_in_r, nbr, field, ws, st, &pr));
        bool done = false;
        DBT_TRY(fused_semijoin(pr.keys.w0, pr.info.nrows, d_in_s, nbs, field, d_out, out_capacity_blocks, ws, st, nres, &done));
        if (done || !pr.info.nrows) return finish(st);
        ws.release(m0);
    }
    DBT_TRY(prepare_pair(d_in_r, nbr, d_in_s, nbs, field, ws, st, &pr, &ps));
    const uint64_t ns = ps.info.nrows;
    if (ns && pr.info.nrows) {
        uint32_t *counts = ws.take<uint32_t>(ns);
        uint64_t cap = out_capacity_blocks * kRpb;
        uint32_t *rows = ws.take<uint32_t>(std::max<uint64_t>(std::min<uint64_t>(cap, (field == '3') ? cap : ns), 1));
        uint64_t *d_total = ws.take<uint64_t>(8);
        if (!counts || !rows || !d_total) {
            set_error("hashjoin: workspace too small");
            return DBT_ERR_WORKSPACE;
        }
        uint64_t rows_cap = std::max<uint64_t>(std::min<uint64_t>(cap, (field == '3') ? cap : ns), 1);
        // str keys whose varying bits (over R and S together) fit one word are joined as that word: the set-semantics
        // u32 paths (direct-address bitmap, sliced bitmap, u32 table) replace the 32-byte-key hash table
        KeyCols rk = pr.keys, sk = ps.keys;
        int jf = field;
        bool have_counts = false;
        if (field >= '2') {
            // Keys whose varying bits (over R and S together) fit 64 bits are joined as those bits -- an exact key of one
            // or two words.  One word and set semantics (field '2'): the u32 paths (direct-address bitmap).  Otherwise:
            // radix-partitioned build / probe with per-partition tables in shared memory (field '3' carries a multiplicity
            // counter per key).  Wider keys, or a partition that overflows its table, use the linear-probing table in HBM.
            KeyCols j;
            CompactPlan plan;
            const int bits = joint_compaction(pr.keys, ps.keys, field, &plan, &j);
            if (field == '2' && bits && bits <= 32) {
                uint32_t *lo_r = ws.take<uint32_t>(pr.info.nrows), *lo_s = ws.take<uint32_t>(ns);
                if (lo_r && lo_s) {
                    DBT_TRY(launch_compaction(pr.keys, field, plan, lo_r, nullptr, st));
                    DBT_TRY(launch_compaction(ps.keys, field, plan, lo_s, nullptr, st));
                    rk.w0 = lo_r;
                    sk.w0 = lo_s;
                    jf = '1';
                }
            } else if (bits && bits <= 64 && getenv("DBT_JOIN_NO_RADIX") == nullptr) {
                const size_t m1 = ws.mark();
                const uint64_t nr = pr.info.nrows;
                uint32_t *lo_r = ws.take<uint32_t>(nr), *lo_s = ws.take<uint32_t>(ns);
                uint32_t *hi_r = bits > 32 ? ws.take<uint32_t>(nr) : nullptr, *hi_s = bits > 32 ? ws.take<uint32_t>(ns) : nullptr;
                if (lo_r && lo_s && (bits <= 32 || (hi_r && hi_s))) {
                    DBT_TRY(launch_compaction(pr.keys, field, plan, lo_r, hi_r, st));
                    DBT_TRY(launch_compaction(ps.keys, field, plan, lo_s, hi_s, st));
                    bool over = false;
                    int rc = radix_join_counts(hi_r, lo_r, nr, hi_s, lo_s, ns, field == '3', counts, &over, ws, st);
                    if (rc == 0 && !over) have_counts = true;
                    else if (rc != DBT_ERR_WORKSPACE && rc != 0) return rc;
                }
                ws.release(m1);
            }
        }
        if (!have_counts) DBT_TRY(hash_join_counts(rk, sk, jf, counts, ws, st));
        DBT_TRY(compact_select(counts, nullptr, ns, rows, rows_cap, d_total, ws, st));
        uint64_t total = 0;
        DBT_TRY(read_u64(d_total, &total, 1, st));
        *nres = total;
        if (total > cap || total > rows_cap) {
            set_error("hashjoin: output capacity too small (nres returned)");
            return DBT_ERR_WORKSPACE;
        }
        DBT_TRY(gather_records(d_in_s, rows, ps.row_slot, total, d_out, st));
    }
    return finish(st);
}
