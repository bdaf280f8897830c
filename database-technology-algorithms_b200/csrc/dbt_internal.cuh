// dbt_internal.cuh -- shared internals of libdbt_b200 (sm_100a only, no other back end).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stddef.h>
#include <string>
#include "../../include/dbt_b200.h"

namespace dbt {

// ---- on-disk layout in 32-bit words (reference: dbtproj.h:16-38) -------------
constexpr uint32_t kBlockWords = 3504;   // 14016 / 4
constexpr uint32_t kRecWords = 35;       // 140 / 4
constexpr uint32_t kRpb = 100;           // MAX_RECORDS_PER_BLOCK
constexpr uint32_t kEntriesWord = 2;     // entries start at byte 8
constexpr uint32_t kStrWord = 2;         // str starts at byte 8 of a record
constexpr uint32_t kStrWords = 30;       // 120 / 4
constexpr uint32_t kTrailerWord = 3502;  // valid/misc/pad word, then dummy at 3503
constexpr uint32_t kBlockVec4 = 876;     // 14016 / 16

// word offset of slot s (slot = block*100 + entry) inside an image
// Read-only 4-byte load for RANDOM accesses (record gathers, words fetched through a permutation, table probes).  By default
// an L2 miss fills a whole 128-byte line from DRAM; the .L2::64B qualifier makes it a 64-byte fill: ncu dram__bytes_read for
// 50M random 140-byte records drops from 267 to 202 bytes per record, for random 8-byte reads from 132 to 68
// (profiles/micro/l2gran2.cu, profiles/r02_notes.md; cudaLimitMaxL2FetchGranularity changes nothing).  DBT_SPARSE_LD=0 at
// build time restores the plain load for A/B runs.
#ifndef DBT_SPARSE_LD
#define DBT_SPARSE_LD 1
#endif
#ifdef __CUDACC__
__device__ __forceinline__ uint32_t ld_sparse(const uint32_t *p) {
#if DBT_SPARSE_LD
    uint32_t v;
    asm("ld.global.nc.L2::64B.u32 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
#else
    return __ldg(p);
#endif
}
#endif

__host__ __device__ __forceinline__ uint64_t slot_word(uint64_t slot) {
    uint64_t b = slot / kRpb;
    uint32_t e = (uint32_t)(slot - b * kRpb);
    return b * kBlockWords + kEntriesWord + (uint64_t)e * kRecWords;
}

// ---- error plumbing ---------------------------------------------------------------------
void set_error(const std::string &msg);
int cuda_fail(cudaError_t e, const char *what, const char *file, int line);

#define DBT_CUDA(expr)                                                         \
    do {                                                                       \
        cudaError_t e__ = (expr);                                              \
        if (e__ != cudaSuccess) return ::dbt::cuda_fail(e__, #expr, __FILE__, __LINE__); \
    } while (0)
#define DBT_TRY(expr)                  \
    do {                               \
        int rc__ = (expr);             \
        if (rc__ != 0) return rc__;    \
    } while (0)
#define DBT_KERNEL_CHECK() DBT_CUDA(cudaGetLastError())

// ---- stage timing ---------------------------------------------------------------------------
enum Stage {
    ST_HEADERS = 0,
    ST_EXTRACT,
    ST_HIST,
    ST_ONESWEEP,
    ST_WORD_GATHER,
    ST_UNIQUE,
    ST_GATHER,
    ST_HASH_BUILD,
    ST_HASH_PROBE,
    ST_COMPACT,
    ST_INTERSECT,
    ST_MISC,
    ST_H2D,
    ST_D2H,
    ST_COUNT
};
void count_launch(int n = 1);
struct StageScope { // records events around a stage when timing is enabled
    StageScope(int stage, cudaStream_t s);
    ~StageScope();
    int stage;
    cudaStream_t stream;
    int slot;
};
int device_setup();   // one-time per-process device limits
// true exactly once per (current device, key): function attributes such as the dynamic shared-memory limit are per
// device, and one process may drive several GPUs (host_multi.cu), so "set once per process" is not enough
bool first_use_on_device(const void *key);
void stage_resolve(); // call after a stream sync: folds pending event pairs into the totals

// ---- workspace bump allocator (no hidden device allocation on the device-scope path) -------
struct Arena {
    char *base;
    size_t cap, off;
    Arena(void *p, size_t bytes) : base((char *)p), cap(bytes), off(0) {}
    template <typename T> T *take(size_t count) {
        size_t bytes = (count * sizeof(T) + 255) & ~(size_t)255;
        if (off + bytes > cap) return nullptr;
        T *r = (T *)(base + off);
        off += bytes;
        return r;
    }
    size_t mark() const { return off; }
    void release(size_t m) { off = m; }
};
inline size_t pad256(size_t b) { return (b + 255) & ~(size_t)255; }

// ---- key description ------------------------------------------------------------------------
// A key is a few big-endian-comparable u32 words per row, most significant first:
//   field '0': recid                   -> w0                (column array)
//   field '1': num                     -> w0
//   field '2': str, NUL-normalised     -> str[kw]           (AoS array, kw = 8 or 30 words)
//   field '3': num then str            -> w0, str[kw]
struct KeyCols {
    uint32_t *w0;     // [n] most significant word for fields 0/1/3 (nullptr for field 2)
    uint32_t *str;    // [n][kw] AoS big-endian words (nullptr for fields 0/1)
    uint32_t kw;      // words per str key
    uint32_t *recid;  // [n] recid column (tie-break word when recids are not monotone in file order)
    uint64_t n;
    // host copies of the OR/AND statistics gathered during extraction
    uint32_t vary_w0, vary_recid, vary_str[30];
    int recid_unsorted;
    const uint32_t *w0_byte_hist; // [4][256] histogram of w0's bytes made during extraction (or nullptr)
    // the raw OR / AND words behind the vary_* masks (out-of-core: combined over all runs before the global sort)
    uint32_t or_w0, and_w0, or_recid, and_recid, or_str[30], and_str[30];
    // compact-key views made by sort_rows_by_key only: the key words in sorted order (two-word keys: both columns)
    const uint32_t *sorted_hi, *sorted_lo;
};

// ---- launchers implemented in the .cu files --------------------------------------------------
// headers / extraction (kernels_extract.cu)
struct ImageInfo {
    uint64_t nrows;      // live rows
    int prefix_full;     // every block except the last is full => slot == row
    const uint32_t *blk_nres;    // ragged images only: live rows per block ...
    const uint32_t *blk_row_off; // ... and the row index of each block's first live row
};
int image_info(const void *d_image, uint64_t nblocks, uint32_t **d_row_slot_out, Arena &ws, cudaStream_t st,
               ImageInfo *info);
struct ExtractStats { // device-side
    uint32_t or_w0, and_w0, or_recid, and_recid, recid_unsorted, str_overflow, pad[2];
    uint32_t str_or[32], str_and[32];
};
int extract_keys(const void *d_image, uint64_t nblocks_img, uint64_t nrows, const uint32_t *d_row_slot,
                 const uint32_t *d_blk_nres, const uint32_t *d_blk_row_off, int field, uint32_t kw, uint32_t *d_w0,
                 uint32_t *d_str, uint32_t *d_recid, ExtractStats *d_stats, cudaStream_t st,
                 uint32_t *d_byte_hist = nullptr, int *hist_done = nullptr);

// radix sort (kernels_sort.cu)
size_t sort_ws_bytes(uint64_t n);
// sorts pairs over the key bits set in `varying_mask` (digits covering only constant bits are skipped)
int sort_pairs_masked(uint32_t *&keys, uint32_t *&keys_alt, uint32_t *&vals, uint32_t *&vals_alt, uint64_t n,
                      uint32_t varying_mask, bool iota_vals, Arena &ws, cudaStream_t st,
                      const uint32_t *byte_hist = nullptr /* [4][256] counts of the keys' bytes, if already known */);
int or_and_reduce(const uint32_t *d_words, uint64_t n, uint32_t *d_or_and /*2 words*/, cudaStream_t st);
int gather_word(const uint32_t *d_src, uint32_t stride, uint32_t word, const uint32_t *d_perm, uint32_t *d_out,
                uint64_t n, cudaStream_t st);
int iota_u32(uint32_t *d, uint64_t n, cudaStream_t st);

// unique / compaction / gather (kernels_gather.cu)
int unique_rows(const KeyCols &k, int field, const uint32_t *d_perm, const uint32_t *d_sorted_w0, uint64_t n,
                uint32_t *d_uperm, uint32_t *d_ukeys, uint64_t *d_count /*device u64*/, Arena &ws, cudaStream_t st);
// two-word keys whose two columns are at hand in sorted order: first position of every group of equal (hi, lo);
// d_upos (optional) receives the sorted position of every unique row
int unique_rows_sorted2(const uint32_t *d_hi, const uint32_t *d_lo, const uint32_t *d_perm, uint64_t n, uint32_t *d_uperm,
                        uint32_t *d_upos, uint64_t *d_count, Arena &ws, cudaStream_t st);
// out[i] = {hi[pos[i]], lo[pos[i]]}
int take_pairs(const uint32_t *d_hi, const uint32_t *d_lo, const uint32_t *d_pos, uint64_t n, uint32_t *d_out, cudaStream_t st);
// out = values[i] (or i when values == nullptr) repeated counts[i] times, ascending i; *d_total = sum(counts)
int compact_select(const uint32_t *d_counts, const uint32_t *d_values, uint64_t n, uint32_t *d_out, uint64_t out_cap,
                   uint64_t *d_total, Arena &ws, cudaStream_t st);
int exclusive_offsets(const uint32_t *d_counts, uint64_t n, uint32_t *d_off, uint64_t *d_total, Arena &ws,
                      cudaStream_t st);
// blockid0: blockid of the first output block (a chunk of a larger image continues the numbering)
int gather_records(const void *d_in, const uint32_t *d_rows, const uint32_t *d_row_slot, uint64_t nrows_out,
                   void *d_out, cudaStream_t st, int max_ctas = 0, uint32_t blockid0 = 0);

// device-scope pipeline steps shared with the out-of-core host operators (device_ops.cu)
struct Prepared {
    ImageInfo info;
    uint32_t *row_slot; // nullptr when slot == row
    KeyCols keys;
};
// headers + key extraction of one image into columns taken from `ws`; force_kw: words per str key (0 = 8, widened
// to 30 automatically when a string has no NUL in its first 32 bytes)
int prepare(const void *d_img, uint64_t nblocks, int field, Arena &ws, cudaStream_t st, Prepared *out,
            uint32_t force_kw = 0);
// rows ordered by (key(field), recid), remaining ties in file order: the row permutation and, for one-word keys,
// the sorted key column
// compact != nullptr (single-relation callers only): a str / num+str key with at most 64 varying bits is sorted as the
// concatenation of those bits; *compact then describes that key (n != 0; w0 = its most significant word, str = the
// second word with kw = 1 when it needs two) and equal compact keys mean equal keys, so the caller's unique scan can
// run on it -- with the returned sorted column (field '1' semantics) when it is one word, through the view
// (field '3' semantics) when it is two.
int sort_rows_by_key(KeyCols &k, int field, Arena &ws, cudaStream_t st, uint32_t **perm_out, uint32_t **sorted_w0_out,
                     KeyCols *compact = nullptr);
int sorted_rows(const void *d_img, uint64_t nblocks, int field, bool dedup, Arena &ws, cudaStream_t st, uint32_t **rows,
                uint32_t **row_slot, uint64_t *n_in, uint64_t *n_out);
// first row of every group of equal keys in sorted order, using the compact key when sort_rows_by_key made one
int unique_sorted(const KeyCols &k, const KeyCols &compact, int field, const uint32_t *d_perm, const uint32_t *d_sorted_w0,
                  uint64_t n, uint32_t *d_uperm, uint64_t *d_count, Arena &ws, cudaStream_t st);

// joins (kernels_join.cu)
size_t hash_table_slots(uint64_t nr);
// radix-partitioned build / probe with per-partition shared-memory tables on exact 64-bit keys (hi may be null)
int radix_join_counts(const uint32_t *r_hi, const uint32_t *r_lo, uint64_t nr, const uint32_t *s_hi, const uint32_t *s_lo, uint64_t ns,
                      bool multi, uint32_t *d_counts, bool *overflowed, Arena &ws, cudaStream_t st);
int build_key_bitmap(const uint32_t *d_keys, uint64_t n, Arena &ws, cudaStream_t st, uint32_t **bm, uint32_t *base,
                     uint32_t *span);
// one streaming pass over the S image (kernels_semijoin.cu): matching records go straight to the packed output image;
// d_total: 16 bytes on the device {u64 matches, u32 look-back error flag}
int semijoin_stream(const void *d_s_img, uint64_t nblocks_s, int field, const uint32_t *d_bitmap, uint32_t base, uint32_t span,
                    void *d_out, uint64_t cap_rows, uint64_t *d_total, Arena &ws, cudaStream_t st);
// two streaming passes (count per block, then copy to the scanned offsets); *h_total = matches (nothing is written when
// they exceed cap_rows); synchronises the stream once (the total decides whether the copy pass runs)
int semijoin_two_pass(const void *d_s_img, uint64_t nblocks_s, int field, const uint32_t *d_bitmap, uint32_t base, uint32_t span,
                      void *d_out, uint64_t cap_rows, uint64_t *d_total, uint64_t *h_total, Arena &ws, cudaStream_t st);
int hash_join_counts(const KeyCols &r, const KeyCols &s, int field, uint32_t *d_counts /*[s.n]*/, Arena &ws,
                     cudaStream_t st);
// sorted unique row lists of R and S -> per-R-unique-row 0/1 match flags, and the reference walk's read count
int intersect_sorted(const KeyCols &r, const uint32_t *d_ur, const uint32_t *d_urkeys, uint64_t nur, const KeyCols &s,
                     const uint32_t *d_us, const uint32_t *d_uskeys, uint64_t nus, int field, uint32_t *d_flags,
                     uint64_t *d_later_reads, Arena &ws, cudaStream_t st, bool keys_are_contiguous = false);
// keys_are_contiguous: multi-word keys only -- d_urkeys / d_uskeys already hold the sorted unique keys, all words of a
// key side by side (w0 first), so the merge path reads them directly instead of collecting them through the row lists

int match_ranges(const KeyCols &r, const uint32_t *d_rperm, const uint32_t *d_rsorted_w0, const KeyCols &s, int field,
                 uint32_t *d_first, uint32_t *d_count, cudaStream_t st);
int expand_pairs(const uint32_t *d_count, const uint32_t *d_first, uint64_t ns, const uint32_t *d_rperm,
                 const uint32_t *d_r_recid, const uint32_t *d_s_recid, uint32_t *d_pairs, uint64_t cap, uint64_t *d_total,
                 Arena &ws, cudaStream_t st);

// generator (kernels_gen.cu)
int gen_syn(uint64_t seed, uint64_t n_total, uint64_t U, int kind, uint64_t row0, uint64_t nrows, uint32_t recid0,
            void *d_image, cudaStream_t st);

} // namespace dbt
