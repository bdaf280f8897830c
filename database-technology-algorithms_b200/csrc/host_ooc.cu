// host_ooc.cu -- out-of-core forms of the host-scope operators.
//
// The reference bounds its memory with `nmem_blocks`: MergeSort sorts runs of nmem_blocks blocks, writes
// them to segment files and merges them (DatabaseProject.cpp:191-233, 237-369); HashJoin reads R and S in
// chunks of nmem_blocks-1 blocks (:521-522, :564).  The GPU analogue, for images that do not fit in HBM:
//
//   sort / dedup   phase 1  the image goes through the device in chunks ("runs"): each is sorted (or sorted and
//                           deduplicated) by the in-core operator and comes back to a host scratch area, while
//                           its key columns (key words + recid, 8..36 bytes per row) stay on the device;
//                  phase 2  one global LSD sort of the resident columns gives the output order as a list of
//                           (run, row-in-run) positions.  Because every run is sorted, the rows of any output
//                           chunk are a contiguous slice of every run: the slices are staged side by side, one
//                           gather packs the chunk, and it goes to its place in the caller's output image.
//                  The records cross PCIe four times (the reference's two-phase sort reads and writes every
//                  block once per phase as well), the keys are sorted on the device only.
//   hash join      R's key columns are extracted chunk by chunk and stay resident; S streams through in chunks,
//                  each probed against R and its matching rows appended to the output (a partly filled last
//                  block is carried into the next chunk's gather so the output image stays packed).
//
// All kernels on this path are the in-core ones plus the two small index kernels below.
#include "host_ctx.cuh"
#include <cstdlib>
#include <vector>

namespace dbt {

static uint64_t g_chunk_override = 0;
constexpr uint32_t kMaxRuns = 2048;
// what the last out-of-core call did: runs (or R chunks), output chunks, chunk shrinks, key-width restarts,
// blocks staged for the gathers, S chunks
static uint64_t g_stats[8] = {0, 0, 0, 0, 0, 0, 0, 0};
enum { OS_RUNS, OS_CHUNKS, OS_SHRINKS, OS_WIDENED, OS_STAGED, OS_SCHUNKS };

uint64_t ooc_chunk_blocks(int op, uint64_t nbr, uint64_t nbs, int field, size_t cached_bytes) {
    uint64_t c = g_chunk_override;
    if (!c)
        if (const char *e = getenv("DBT_OOC_CHUNK_BLOCKS")) c = strtoull(e, nullptr, 10);
    if (!c) {
        size_t free_b = 0, total_b = 0;
        if (cudaMemGetInfo(&free_b, &total_b) != cudaSuccess) {
            cudaGetLastError();
            return 0;
        }
        // in-core: inputs, outputs and the workspace are resident together
        const double img = (double)DBT_BLOCK_BYTES;
        double need = (double)dbt_dev_ws_bytes_kw(op, nbr, nbs, field, 8);
        if (op == DBT_OP_HASHJOIN) need += img * (double)(nbr + 2 * nbs);
        else if (op == DBT_OP_MERGEJOIN) need += img * (double)(2 * nbr + 2 * nbs + std::min(nbr, nbs));
        else need += 2.0 * img * (double)nbr;
        // what this job can have: the memory that is free now plus what the context already holds and will reuse
        // (another job slot or a host framework sharing the GPU shrinks it: chunk instead of failing in cudaMalloc)
        const double avail = std::min((double)total_b, (double)free_b + (double)cached_bytes);
        if (need <= 0.92 * avail) return 0;
        c = (uint64_t)(0.13 * avail / img); // staged chunk + packed chunk + the in-core sort of one run
    }
    const uint64_t big = (op == DBT_OP_HASHJOIN || op == DBT_OP_MERGEJOIN) ? std::max(nbr, nbs) : nbr;
    return big > c ? c : 0;
}

// ---- small index kernels ------------------------------------------------------------------------
// L[q] is a position in the concatenation of all runs' rows; run r owns [off[r], off[r+1]).
__device__ __forceinline__ uint32_t run_of(const uint32_t *s_off, uint32_t R, uint32_t g) {
    uint32_t lo = 0, hi = R; // largest r with s_off[r] <= g
    while (hi - lo > 1) {
        const uint32_t mid = (lo + hi) >> 1;
        if (s_off[mid] <= g) lo = mid;
        else hi = mid;
    }
    return lo;
}

// per run: the smallest and largest row-in-run among L[q0, q1)
__global__ void __launch_bounds__(256)
slice_bounds_kernel(const uint32_t *__restrict__ L, uint64_t q0, uint64_t q1, const uint32_t *__restrict__ off, uint32_t R,
                    uint32_t *__restrict__ lo, uint32_t *__restrict__ hi) {
    extern __shared__ uint32_t sm[];
    uint32_t *s_off = sm, *s_lo = sm + R + 1, *s_hi = s_lo + R;
    for (uint32_t i = threadIdx.x; i <= R; i += blockDim.x) s_off[i] = off[i];
    for (uint32_t i = threadIdx.x; i < R; i += blockDim.x) {
        s_lo[i] = 0xFFFFFFFFu;
        s_hi[i] = 0;
    }
    __syncthreads();
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t q = q0 + (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; q < q1; q += stride) {
        const uint32_t g = L[q];
        const uint32_t r = run_of(s_off, R, g);
        const uint32_t i = g - s_off[r];
        atomicMin(&s_lo[r], i);
        atomicMax(&s_hi[r], i);
    }
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < R; i += blockDim.x)
        if (s_lo[i] != 0xFFFFFFFFu) {
            atomicMin(&lo[i], s_lo[i]);
            atomicMax(&hi[i], s_hi[i]);
        }
}

// position in the concatenation -> row slot in the staging image (base[r] is pre-wrapped modulo 2^32)
__global__ void __launch_bounds__(256)
slice_map_kernel(const uint32_t *__restrict__ L, uint64_t q0, uint64_t q1, const uint32_t *__restrict__ off, uint32_t R,
                 const uint32_t *__restrict__ base, uint32_t *__restrict__ idx) {
    extern __shared__ uint32_t sm[];
    uint32_t *s_off = sm, *s_base = sm + R + 1;
    for (uint32_t i = threadIdx.x; i <= R; i += blockDim.x) s_off[i] = off[i];
    for (uint32_t i = threadIdx.x; i < R; i += blockDim.x) s_base[i] = base[i];
    __syncthreads();
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t q = q0 + (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; q < q1; q += stride) {
        const uint32_t g = L[q];
        const uint32_t r = run_of(s_off, R, g);
        idx[q - q0] = s_base[r] + (g - s_off[r]);
    }
}

// hash join: carried rows are slots 0..carry-1 of staging block 0, the chunk's image starts at block 1
__global__ void __launch_bounds__(256)
carry_slots_kernel(const uint32_t *__restrict__ rows, const uint32_t *__restrict__ row_slot, uint32_t carry, uint64_t total,
                   uint32_t *__restrict__ slots) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < carry + total; i += stride) {
        if (i < carry) {
            slots[i] = (uint32_t)i;
        } else {
            const uint32_t row = rows[i - carry];
            slots[i] = (row_slot ? row_slot[row] : row) + kRpb;
        }
    }
}

// ---- resident key columns of all runs -------------------------------------------------------------
struct Columns {
    uint32_t *recid = nullptr, *w0 = nullptr, *str = nullptr;
    uint32_t kw = 8;
    uint64_t cap_rows = 0, n = 0;
    uint32_t or_w0 = 0, and_w0 = 0xFFFFFFFFu, or_recid = 0, and_recid = 0xFFFFFFFFu, or_str[30], and_str[30];
    int layout(Buf &buf, uint64_t rows, int field, uint32_t kw_) {
        kw = kw_;
        cap_rows = rows;
        n = 0;
        or_w0 = or_recid = 0;
        and_w0 = and_recid = 0xFFFFFFFFu;
        for (int j = 0; j < 30; ++j) {
            or_str[j] = 0;
            and_str[j] = 0xFFFFFFFFu;
        }
        const bool has_w0 = field != '2', has_str = field >= '2';
        size_t b = pad256(4 * rows) * (1 + (has_w0 ? 1 : 0)) + (has_str ? pad256(4 * rows * kw) : 0) + 1024;
        DBT_TRY(buf.ensure(b));
        char *p = (char *)buf.p;
        recid = (uint32_t *)p;
        p += pad256(4 * rows);
        w0 = has_w0 ? (uint32_t *)p : nullptr;
        if (has_w0) p += pad256(4 * rows);
        str = has_str ? (uint32_t *)p : nullptr;
        return 0;
    }
    // append the columns of one prepared image
    int append(const Prepared &p, cudaStream_t st) {
        const uint64_t m = p.info.nrows;
        if (!m) return 0;
        if (n + m > cap_rows || p.keys.kw != kw) {
            set_error("out-of-core: column store overflow");
            return DBT_ERR_WORKSPACE;
        }
        const KeyCols &k = p.keys;
        DBT_CUDA(cudaMemcpyAsync(recid + n, k.recid, 4 * m, cudaMemcpyDeviceToDevice, st));
        if (w0) DBT_CUDA(cudaMemcpyAsync(w0 + n, k.w0, 4 * m, cudaMemcpyDeviceToDevice, st));
        if (str) DBT_CUDA(cudaMemcpyAsync(str + n * kw, k.str, 4 * m * kw, cudaMemcpyDeviceToDevice, st));
        or_w0 |= k.or_w0;
        and_w0 &= k.and_w0;
        or_recid |= k.or_recid;
        and_recid &= k.and_recid;
        for (uint32_t j = 0; j < kw && j < 30; ++j) {
            or_str[j] |= k.or_str[j];
            and_str[j] &= k.and_str[j];
        }
        n += m;
        return 0;
    }
    KeyCols keycols(int field) const {
        KeyCols k;
        memset(&k, 0, sizeof k);
        k.w0 = w0;
        k.str = str;
        k.recid = recid;
        k.kw = kw;
        k.n = n;
        k.vary_w0 = w0 ? (or_w0 ^ and_w0) : 0;
        k.vary_recid = or_recid ^ and_recid;
        for (uint32_t j = 0; j < 30; ++j) k.vary_str[j] = (str && j < kw) ? (or_str[j] ^ and_str[j]) : 0;
        k.recid_unsorted = 1; // runs interleave: recid order across the concatenation is not file order
        (void)field;
        return k;
    }
};

// host scratch for the runs: pinned when the host can spare it, pageable (staged copies) otherwise
static int runs_scratch(HostCtx &c, size_t bytes, void **p) {
    if (c.runs.cap >= bytes) {
        *p = c.runs.p;
        return 0;
    }
    c.runs.release();
    void *q = nullptr;
    if (cudaHostAlloc(&q, bytes, cudaHostAllocDefault) == cudaSuccess) {
        c.runs.pinned = true;
    } else {
        cudaGetLastError();
        q = malloc(bytes);
        c.runs.pinned = false;
        c.runs.malloced = true;
        if (!q) {
            set_error("out-of-core: no host memory for the sorted runs");
            return DBT_ERR_WORKSPACE;
        }
    }
    c.runs.p = q;
    c.runs.cap = bytes;
    *p = q;
    return 0;
}

static int sync(cudaStream_t st) {
    DBT_CUDA(cudaStreamSynchronize(st));
    return 0;
}
// on every exit path (errors included) nothing may still be copying into the caller's buffers
struct DrainOnExit {
    cudaStream_t a, b;
    ~DrainOnExit() {
        cudaStreamSynchronize(a);
        cudaStreamSynchronize(b);
    }
};
static int sync_event(cudaEvent_t e) {
    DBT_CUDA(cudaEventSynchronize(e));
    return 0;
}

// Upload the image [h, h + nb blocks), prepare it at key width `kw` and append its columns.  *widened is set
// (and nothing appended) when the image has strings that need the full 120-byte key while kw is 8.
static int collect_columns(HostCtx &c, const void *d_img, uint64_t nb, int field, Columns &cols, bool *widened) {
    Arena ws(c.ws.p, c.ws.cap);
    Prepared p;
    DBT_TRY(prepare(d_img, nb, field, ws, c.st, &p, cols.kw == 8 ? 0 : cols.kw));
    if (p.info.nrows && field >= '2' && p.keys.kw != cols.kw) {
        *widened = true;
        return 0;
    }
    DBT_TRY(cols.append(p, c.st));
    return sync(c.st);
}

int ooc_sort(HostCtx &c, const void *h_in, uint64_t nblocks, int field, void *h_out, bool dedup, uint64_t C, uint64_t *nrows,
             uint64_t *nunique) {
    const uint64_t R = (nblocks + C - 1) / C;
    const uint64_t nmax = nblocks * kRpb;
    if (R > kMaxRuns || C < 3 * R + 8) {
        set_error("out-of-core sort: the chunk is too small for this image (too many runs)");
        return DBT_ERR_UNSUPPORTED;
    }
    if (nmax >= (1ull << 30)) {
        set_error("out-of-core sort: fewer than 2^30 rows per call");
        return DBT_ERR_UNSUPPORTED;
    }
    cudaStream_t st = c.st;
    DrainOnExit drain{c.st, c.st2};
    memset(g_stats, 0, sizeof g_stats);
    g_stats[OS_RUNS] = R;
    const int op = dedup ? DBT_OP_DEDUP : DBT_OP_SORT;
    const uint64_t S = C + 2 * R + 2; // staging blocks: a chunk's rows plus two partial blocks per run
    // double buffering: while chunk k's result goes home on st2, chunk k+1 comes in and is processed on st
    Buf *in[2] = {&c.in_r, &c.in_s}, *out[2] = {&c.out0, &c.out1};
    for (int k = 0; k < 2; ++k) {
        DBT_TRY(in[k]->ensure(S * DBT_BLOCK_BYTES));
        DBT_TRY(out[k]->ensure(C * DBT_BLOCK_BYTES));
    }
    cudaStream_t st2 = c.st2;
    void *runs = nullptr;
    DBT_TRY(runs_scratch(c, (size_t)nblocks * DBT_BLOCK_BYTES, &runs));
    Columns cols;
    DBT_TRY(cols.layout(c.cols, nmax, field, 8));
    std::vector<uint64_t> run_rows(R, 0);
    uint64_t n_in = 0;
    bool widened = false;

    // ---- phase 1: runs ----------------------------------------------------------------------------
    for (uint64_t r = 0; r < R; ++r) {
        const int k = (int)(r & 1);
        const uint64_t nb = std::min<uint64_t>(C, nblocks - r * C);
        if (r >= 2) DBT_CUDA(cudaStreamWaitEvent(st, c.landed[k], 0)); // out[k] still holds run r-2 until it is home
        DBT_TRY(upload(c, (const char *)h_in + r * C * DBT_BLOCK_BYTES, in[k]->p, nb * DBT_BLOCK_BYTES));
        uint64_t n = 0, u = 0;
        DBT_TRY(with_workspace(c, op, nb, 0, field, [&](void *ws, size_t wb) {
            return dedup ? dbt_dev_dedup(in[k]->p, nb, field, out[k]->p, ws, wb, st, &n, &u)
                         : dbt_dev_mergesort(in[k]->p, nb, field, out[k]->p, ws, wb, st, &n);
        }));
        n_in += n;
        run_rows[r] = dedup ? u : n;
        if (!widened) DBT_TRY(collect_columns(c, out[k]->p, blocks_for(run_rows[r]), field, cols, &widened));
        DBT_CUDA(cudaEventRecord(c.packed[k], st));
        DBT_CUDA(cudaStreamWaitEvent(st2, c.packed[k], 0));
        DBT_TRY(download(c, out[k]->p, (char *)runs + r * C * DBT_BLOCK_BYTES, blocks_for(run_rows[r]) * DBT_BLOCK_BYTES, st2));
        DBT_CUDA(cudaEventRecord(c.landed[k], st2));
    }
    DBT_TRY(sync(st2));
    DBT_TRY(sync(st));
    if (widened) { // some run has strings without a NUL in 32 bytes: collect every run's columns again at 120 bytes
        g_stats[OS_WIDENED] = 1;
        DBT_TRY(cols.layout(c.cols, nmax, field, kStrWords));
        DBT_TRY(c.ws.ensure(dbt_dev_ws_bytes_kw(op, C, 0, field, kStrWords)));
        for (uint64_t r = 0; r < R; ++r) {
            const uint64_t nb = blocks_for(run_rows[r]);
            DBT_TRY(upload(c, (const char *)runs + r * C * DBT_BLOCK_BYTES, in[0]->p, nb * DBT_BLOCK_BYTES));
            bool again = false;
            DBT_TRY(collect_columns(c, in[0]->p, nb, field, cols, &again));
        }
    }
    const uint64_t n = cols.n;
    if (nrows) *nrows = n_in;
    uint64_t m = n; // output rows

    // ---- phase 2: global order of the resident columns ----------------------------------------------
    std::vector<uint32_t> h_off(R + 1, 0);
    for (uint64_t r = 0; r < R; ++r) h_off[r + 1] = h_off[r] + (uint32_t)run_rows[r];
    const size_t ws2 = 10 * pad256(4 * n) + sort_ws_bytes(n) + pad256(8 * (n / 2048 + 2)) + pad256(4 * C * kRpb) + (4 << 20);
    DBT_TRY(c.ws.ensure(ws2));
    Arena ws(c.ws.p, c.ws.cap);
    uint32_t *d_off = ws.take<uint32_t>(R + 1), *d_lo = ws.take<uint32_t>(R), *d_hi = ws.take<uint32_t>(R),
             *d_base = ws.take<uint32_t>(R), *d_idx = ws.take<uint32_t>(C * kRpb);
    uint64_t *d_cnt = ws.take<uint64_t>(8);
    const uint32_t *L = nullptr;
    if (n) {
        KeyCols G = cols.keycols(field);
        uint32_t *perm, *sorted;
        KeyCols compact;
        DBT_TRY(sort_rows_by_key(G, field, ws, st, &perm, &sorted, &compact));
        L = perm;
        if (dedup) {
            uint32_t *urows = ws.take<uint32_t>(n);
            if (!urows || !d_cnt) {
                set_error("out-of-core dedup: workspace too small");
                return DBT_ERR_WORKSPACE;
            }
            DBT_TRY(unique_sorted(G, compact, field, perm, sorted, n, urows, d_cnt, ws, st));
            DBT_CUDA(cudaMemcpyAsync(&m, d_cnt, 8, cudaMemcpyDeviceToHost, st));
            DBT_TRY(sync(st));
            L = urows;
        }
    }
    if (nunique) *nunique = m;
    if (!d_off || !d_lo || !d_hi || !d_base || !d_idx) {
        set_error("out-of-core sort: workspace too small");
        return DBT_ERR_WORKSPACE;
    }
    DBT_CUDA(cudaMemcpyAsync(d_off, h_off.data(), 4 * (R + 1), cudaMemcpyHostToDevice, st));

    // ---- phase 2b: output chunks ---------------------------------------------------------------------
    std::vector<uint32_t> h_lo(R), h_hi(R), h_base(R);
    const size_t smem = (3 * R + 2) * 4;
    uint64_t want = C * kRpb; // rows per output chunk; shrinks when a dedup chunk's slices would not fit the staging
    uint64_t chunk_no = 0;
    for (uint64_t q0 = 0; q0 < m; ++chunk_no) {
        const int k = (int)(chunk_no & 1);
        uint64_t q1 = std::min<uint64_t>(m, q0 + want);
        uint64_t staged = 0;
        for (;;) {
            std::fill(h_lo.begin(), h_lo.end(), 0xFFFFFFFFu);
            std::fill(h_hi.begin(), h_hi.end(), 0u);
            DBT_CUDA(cudaMemcpyAsync(d_lo, h_lo.data(), 4 * R, cudaMemcpyHostToDevice, st));
            DBT_CUDA(cudaMemcpyAsync(d_hi, h_hi.data(), 4 * R, cudaMemcpyHostToDevice, st));
            const int grid = (int)std::min<uint64_t>((q1 - q0 + 255) / 256, 148 * 8);
            slice_bounds_kernel<<<grid, 256, smem, st>>>(L, q0, q1, d_off, (uint32_t)R, d_lo, d_hi);
            count_launch();
            DBT_KERNEL_CHECK();
            DBT_CUDA(cudaMemcpyAsync(h_lo.data(), d_lo, 4 * R, cudaMemcpyDeviceToHost, st));
            DBT_CUDA(cudaMemcpyAsync(h_hi.data(), d_hi, 4 * R, cudaMemcpyDeviceToHost, st));
            DBT_TRY(sync(st));
            staged = 0;
            for (uint64_t r = 0; r < R; ++r)
                if (h_lo[r] <= h_hi[r]) staged += h_hi[r] / kRpb - h_lo[r] / kRpb + 1;
            if (staged <= S) break;
            if (q1 - q0 <= kRpb) {
                set_error("out-of-core sort: internal error (a 100-row chunk does not fit the staging)");
                return DBT_ERR_UNSUPPORTED;
            }
            want = std::max<uint64_t>(kRpb, (q1 - q0) / 2 / kRpb * kRpb);
            q1 = q0 + want; // (q1 < m here, so the chunk stays a whole number of blocks)
            ++g_stats[OS_SHRINKS];
        }
        ++g_stats[OS_CHUNKS];
        g_stats[OS_STAGED] += staged;
        uint64_t cursor = 0;
        for (uint64_t r = 0; r < R; ++r) {
            h_base[r] = 0;
            if (h_lo[r] > h_hi[r]) continue;
            const uint64_t b0 = h_lo[r] / kRpb, nb = h_hi[r] / kRpb - b0 + 1;
            DBT_TRY(upload(c, (const char *)runs + (r * C + b0) * DBT_BLOCK_BYTES, (char *)in[k]->p + cursor * DBT_BLOCK_BYTES,
                           nb * DBT_BLOCK_BYTES));
            h_base[r] = (uint32_t)(cursor * kRpb - b0 * kRpb); // modulo 2^32; the sum with the row-in-run is in range
            cursor += nb;
        }
        DBT_CUDA(cudaMemcpyAsync(d_base, h_base.data(), 4 * R, cudaMemcpyHostToDevice, st));
        const int grid = (int)std::min<uint64_t>((q1 - q0 + 255) / 256, 148 * 8);
        slice_map_kernel<<<grid, 256, (2 * R + 1) * 4, st>>>(L, q0, q1, d_off, (uint32_t)R, d_base, d_idx);
        count_launch();
        DBT_KERNEL_CHECK();
        if (chunk_no >= 2) DBT_CUDA(cudaStreamWaitEvent(st, c.landed[k], 0)); // out[k]: chunk_no-2 must be home first
        DBT_TRY(gather_records(in[k]->p, d_idx, nullptr, q1 - q0, out[k]->p, st, 0, (uint32_t)(q0 / kRpb)));
        DBT_CUDA(cudaEventRecord(c.packed[k], st));
        DBT_CUDA(cudaStreamWaitEvent(st2, c.packed[k], 0));
        DBT_TRY(download(c, out[k]->p, (char *)h_out + (q0 / kRpb) * DBT_BLOCK_BYTES, blocks_for(q1 - q0) * DBT_BLOCK_BYTES, st2));
        DBT_CUDA(cudaEventRecord(c.landed[k], st2));
        q0 = q1;
        if (want < C * kRpb) want = std::min<uint64_t>(C * kRpb, want * 2);
    }
    DBT_TRY(sync(st2));
    DBT_TRY(sync(st));
    return 0;
}

int ooc_hashjoin(HostCtx &c, const void *h_in_r, uint64_t nbr, const void *h_in_s, uint64_t nbs, int field, void *h_out,
                 uint64_t out_capacity_blocks, uint64_t C, uint64_t *nres) {
    cudaStream_t st = c.st;
    DrainOnExit drain{c.st, c.st2};
    const uint64_t cap_rows = out_capacity_blocks * kRpb;
    if (nbr * kRpb >= (1ull << 32) || C * kRpb >= (1ull << 31)) {
        set_error("out-of-core hash join: R must have fewer than 2^32 rows");
        return DBT_ERR_UNSUPPORTED;
    }
    DBT_TRY(c.in_r.ensure(C * DBT_BLOCK_BYTES));
    DBT_TRY(c.in_s.ensure((C + 1) * DBT_BLOCK_BYTES)); // block 0: rows carried from the previous chunk
    Buf *out[2] = {&c.out0, &c.out1}; // chunk k's matches go home on st2 while chunk k+1 comes in on st
    for (int k = 0; k < 2; ++k) DBT_TRY(out[k]->ensure((C + 1) * DBT_BLOCK_BYTES));
    cudaStream_t st2 = c.st2;
    char *stage = (char *)c.in_s.p;
    Columns cols;
    uint32_t kw = 8;
    memset(g_stats, 0, sizeof g_stats);
    uint64_t total_out = 0; // rows matched so far (keeps counting past the capacity so the caller learns the size)
    for (int attempt = 0; attempt < 2; ++attempt) {
        // ---- R: key columns, chunk by chunk ------------------------------------------------------------
        DBT_TRY(cols.layout(c.cols, nbr * kRpb, field, kw));
        // one S chunk's columns + R's table (or bitmap): R's own key columns live in c.cols, not in the workspace
        DBT_TRY(c.ws.ensure(dbt_dev_ws_bytes_kw(DBT_OP_HASHJOIN, 0, C, field, kw) + 3 * pad256(4 * hash_table_slots(nbr * kRpb)) +
                            dbt_dev_ws_bytes_kw(DBT_OP_SORT, std::min<uint64_t>(C, nbr), 0, field, kw)));
        bool widened = false;
        g_stats[OS_RUNS] = g_stats[OS_SCHUNKS] = 0;
        for (uint64_t b = 0; b < nbr && !widened; b += C) {
            const uint64_t nb = std::min<uint64_t>(C, nbr - b);
            ++g_stats[OS_RUNS];
            DBT_TRY(upload(c, (const char *)h_in_r + b * DBT_BLOCK_BYTES, c.in_r.p, nb * DBT_BLOCK_BYTES));
            DBT_TRY(collect_columns(c, c.in_r.p, nb, field, cols, &widened));
        }
        if (widened) {
            kw = kStrWords;
            g_stats[OS_WIDENED] = 1;
            continue;
        }
        KeyCols rk = cols.keycols(field);
        // ---- S: stream, probe, append ------------------------------------------------------------------
        uint64_t out_block = 0; // full blocks already in h_out
        uint32_t carry = 0;     // rows waiting in staging block 0
        uint64_t emitted_chunks = 0;
        total_out = 0;
        for (uint64_t b = 0; b < nbs && !widened; b += C) {
            const uint64_t nb = std::min<uint64_t>(C, nbs - b);
            ++g_stats[OS_SCHUNKS];
            DBT_TRY(upload(c, (const char *)h_in_s + b * DBT_BLOCK_BYTES, stage + DBT_BLOCK_BYTES, nb * DBT_BLOCK_BYTES));
            Arena ws(c.ws.p, c.ws.cap);
            Prepared ps;
            DBT_TRY(prepare(stage + DBT_BLOCK_BYTES, nb, field, ws, st, &ps, kw == 8 ? 0 : kw));
            const uint64_t ns = ps.info.nrows;
            if (ns && field >= '2' && ps.keys.kw != kw) {
                widened = true;
                break;
            }
            if (!ns || !rk.n) continue;
            uint32_t *counts = ws.take<uint32_t>(ns);
            uint64_t *d_total = ws.take<uint64_t>(8);
            if (!counts || !d_total) {
                set_error("out-of-core hash join: workspace too small");
                return DBT_ERR_WORKSPACE;
            }
            DBT_TRY(hash_join_counts(rk, ps.keys, field, counts, ws, st));
            uint64_t rows_cap = std::max<uint64_t>(ns, kRpb);
            uint64_t total = 0;
            for (int pass = 0; pass < 2; ++pass) { // field '3' can emit a row several times: size, then emit
                DBT_TRY(c.out2.ensure(8 * (rows_cap + kRpb)));
                uint32_t *rows = (uint32_t *)c.out2.p;
                DBT_TRY(compact_select(counts, nullptr, ns, rows, rows_cap, d_total, ws, st));
                DBT_CUDA(cudaMemcpyAsync(&total, d_total, 8, cudaMemcpyDeviceToHost, st));
                DBT_TRY(sync(st));
                if (total <= rows_cap) break;
                rows_cap = total;
            }
            total_out += total;
            if (total_out > cap_rows) continue; // over capacity: keep counting only
            uint32_t *rows = (uint32_t *)c.out2.p, *slots = rows + rows_cap + kRpb;
            const uint64_t emit = carry + total;
            const int k = (int)(emitted_chunks & 1);
            if (emitted_chunks >= 2) DBT_TRY(sync_event(c.landed[k])); // out[k] may be regrown below: it must be idle
            DBT_TRY(out[k]->ensure((blocks_for(emit) + 1) * DBT_BLOCK_BYTES));
            if (emit) {
                const int grid = (int)std::min<uint64_t>((emit + 255) / 256, 148 * 8);
                carry_slots_kernel<<<grid, 256, 0, st>>>(rows, ps.row_slot, carry, total, slots);
                count_launch();
                DBT_KERNEL_CHECK();
                DBT_TRY(gather_records(stage, slots, nullptr, emit, out[k]->p, st, 0, (uint32_t)out_block));
            }
            const uint64_t full = emit / kRpb;
            carry = (uint32_t)(emit - full * kRpb);
            if (carry) // the partly filled last block waits in staging block 0 for the next chunk's rows
                DBT_CUDA(cudaMemcpyAsync(stage, (char *)out[k]->p + full * DBT_BLOCK_BYTES, DBT_BLOCK_BYTES, cudaMemcpyDeviceToDevice, st));
            DBT_CUDA(cudaEventRecord(c.packed[k], st));
            DBT_CUDA(cudaStreamWaitEvent(st2, c.packed[k], 0));
            DBT_TRY(download(c, out[k]->p, (char *)h_out + out_block * DBT_BLOCK_BYTES, full * DBT_BLOCK_BYTES, st2));
            DBT_CUDA(cudaEventRecord(c.landed[k], st2));
            out_block += full;
            ++emitted_chunks;
            DBT_TRY(sync(st)); // the row lists (out2) and the workspace are reused by the next chunk
        }
        DBT_TRY(sync(st2));
        if (widened) {
            kw = kStrWords;
            g_stats[OS_WIDENED] = 1;
            continue;
        }
        if (carry && total_out <= cap_rows) {
            DBT_TRY(download(c, stage, (char *)h_out + out_block * DBT_BLOCK_BYTES, DBT_BLOCK_BYTES));
            DBT_TRY(sync(st));
        }
        break;
    }
    if (nres) *nres = total_out;
    if (total_out > cap_rows) {
        set_error("hashjoin: output capacity too small (nres returned)");
        return DBT_ERR_WORKSPACE;
    }
    return 0;
}

// ---- merge join out of core = dedup(R), dedup(S) (each out of core when it has to be), then a streamed semi-join
// of dedup(R) against the keys of dedup(S).  dedup(R) is in ascending key order and holds the min-recid row of every
// key, so the rows that survive are exactly the reference's intersection output (DatabaseProject.cpp:414-460, emit
// R's record), in the same order; building on dedup(S) keeps field '3' at one emission per key.
static int host_dedup_any(HostCtx &c, const void *h_in, uint64_t nb, int field, void *h_out, uint64_t C, uint64_t *u) {
    uint64_t n = 0;
    if (nb > C) return ooc_sort(c, h_in, nb, field, h_out, true, C, &n, u);
    const size_t bytes = (size_t)nb * DBT_BLOCK_BYTES;
    DBT_TRY(c.in_r.ensure(bytes));
    DBT_TRY(c.out0.ensure(bytes));
    DBT_TRY(upload(c, h_in, c.in_r.p, bytes));
    DBT_TRY(with_workspace(c, DBT_OP_DEDUP, nb, 0, field, [&](void *ws, size_t wb) {
        return dbt_dev_dedup(c.in_r.p, nb, field, c.out0.p, ws, wb, c.st, &n, u);
    }));
    DBT_TRY(download(c, c.out0.p, h_out, blocks_for(*u) * DBT_BLOCK_BYTES));
    return sync(c.st);
}

// host-side key order of two records (DatabaseProject.cpp:44-92), for the walk's read counter only
static const unsigned char *host_row(const void *img, uint64_t i) {
    return (const unsigned char *)img + (i / kRpb) * DBT_BLOCK_BYTES + 8 + (i % kRpb) * DBT_RECORD_BYTES;
}
static int host_key_cmp(const unsigned char *a, const unsigned char *b, int field) {
    auto u32 = [](const unsigned char *p) {
        uint32_t v;
        memcpy(&v, p, 4);
        return v;
    };
    auto cmp_str = [](const unsigned char *x, const unsigned char *y) {
        for (int i = 0; i < 120; ++i) {
            if (x[i] != y[i]) return x[i] < y[i] ? -1 : 1;
            if (!x[i]) return 0;
        }
        return 0;
    };
    if (field == '0' || field == '1' || field == '3') {
        const uint32_t x = u32(a + (field == '0' ? 0 : 4)), y = u32(b + (field == '0' ? 0 : 4));
        if (x != y) return x < y ? -1 : 1;
        if (field != '3') return 0;
    }
    return cmp_str(a + 8, b + 8);
}
// block reads after the first two of the reference's two-pointer walk over two sorted unique images
// (the closed form of kernels_join.cu: walk_reads_kernel, evaluated on the host images)
uint64_t host_walk_reads(const void *ur, uint64_t nr, const void *us, uint64_t ns, int field) {
    if (!nr || !ns) return 0;
    auto lower_bound = [&](const void *img, uint64_t n, const unsigned char *key) {
        uint64_t lo = 0, hi = n;
        while (lo < hi) {
            const uint64_t mid = (lo + hi) / 2;
            if (host_key_cmp(host_row(img, mid), key, field) < 0) lo = mid + 1;
            else hi = mid;
        }
        return lo;
    };
    const unsigned char *lr = host_row(ur, nr - 1), *ls = host_row(us, ns - 1);
    if (host_key_cmp(lr, ls, field) <= 0) { // R runs out first (or both together)
        const uint64_t lb = lower_bound(us, ns, lr);
        return (nr + kRpb - 1) / kRpb + lb / kRpb;
    }
    const uint64_t lb = lower_bound(ur, nr, ls);
    const bool match = lb < nr && host_key_cmp(ls, host_row(ur, lb), field) == 0;
    return (lb + (match ? 1 : 0)) / kRpb + (ns + kRpb - 1) / kRpb;
}

static int host_scratch(Buf &b, size_t bytes, void **p) { // pinned if the host can spare it, pageable otherwise
    if (b.cap >= bytes && b.p) {
        *p = b.p;
        return 0;
    }
    b.release();
    void *q = nullptr;
    if (cudaHostAlloc(&q, bytes ? bytes : 1, cudaHostAllocDefault) == cudaSuccess) {
        b.pinned = true;
    } else {
        cudaGetLastError();
        q = malloc(bytes ? bytes : 1);
        if (!q) {
            set_error("out-of-core: no host memory for a side image");
            return DBT_ERR_WORKSPACE;
        }
        b.pinned = false;
        b.malloced = true;
    }
    b.p = q;
    b.cap = bytes;
    *p = q;
    return 0;
}

int ooc_mergejoin(HostCtx &c, const void *h_in_r, uint64_t nbr, const void *h_in_s, uint64_t nbs, int field, void *h_out_ur,
                  void *h_out_us, void *h_out, uint64_t C, uint64_t res[4]) {
    void *ur = h_out_ur, *us = h_out_us;
    if (!ur) DBT_TRY(host_scratch(c.side_r, (size_t)nbr * DBT_BLOCK_BYTES, &ur));
    if (!us) DBT_TRY(host_scratch(c.side_s, (size_t)nbs * DBT_BLOCK_BYTES, &us));
    uint64_t nur = 0, nus = 0, nres = 0;
    DBT_TRY(host_dedup_any(c, h_in_r, nbr, field, ur, C, &nur));
    DBT_TRY(host_dedup_any(c, h_in_s, nbs, field, us, C, &nus));
    const uint64_t bur = blocks_for(nur), bus = blocks_for(nus);
    void *out = h_out;
    if (!out) DBT_TRY(host_scratch(c.side_o, (size_t)std::min(bur, bus) * DBT_BLOCK_BYTES, &out));
    if (nur && nus) {
        // semi-join: probe side = dedup(R) (its rows are emitted, in its order), build side = the keys of dedup(S)
        if (std::max(bur, bus) > C) {
            DBT_TRY(ooc_hashjoin(c, us, bus, ur, bur, field, out, std::min(bur, bus), C, &nres));
        } else {
            DBT_TRY(c.in_r.ensure(bus * DBT_BLOCK_BYTES));
            DBT_TRY(c.in_s.ensure(bur * DBT_BLOCK_BYTES));
            DBT_TRY(c.out0.ensure(std::min(bur, bus) * DBT_BLOCK_BYTES));
            DBT_TRY(upload(c, us, c.in_r.p, bus * DBT_BLOCK_BYTES));
            DBT_TRY(upload(c, ur, c.in_s.p, bur * DBT_BLOCK_BYTES));
            DBT_TRY(with_workspace(c, DBT_OP_HASHJOIN, bus, bur, field, [&](void *ws, size_t wb) {
                return dbt_dev_hashjoin(c.in_r.p, bus, c.in_s.p, bur, field, c.out0.p, std::min(bur, bus), ws, wb, c.st, &nres);
            }));
            DBT_TRY(download(c, c.out0.p, out, blocks_for(nres) * DBT_BLOCK_BYTES));
            DBT_TRY(sync(c.st));
        }
    }
    res[0] = nres;
    res[1] = nur;
    res[2] = nus;
    res[3] = host_walk_reads(ur, nur, us, nus, field);
    return 0;
}

} // namespace dbt

extern "C" int dbt_host_ooc_stats(uint64_t out[6]) {
    if (!out) return DBT_ERR_ARG;
    for (int i = 0; i < 6; ++i) out[i] = dbt::g_stats[i];
    return 0;
}

extern "C" int dbt_host_set_chunk_blocks(uint64_t blocks) {
    dbt::g_chunk_override = blocks;
    return 0;
}
