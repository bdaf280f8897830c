// dist.cu -- the multi-GPU host layer (C++, no torch, no NCCL on the data path): dbt_dist_* of include/dbt_b200.h.
//
// SURVEY.md 8e / north_star: sort and dedup shard by sample-sort key-range splitters, joins by key (hash or range), one
// exchange over NVLink each.  One *rank* per GPU.  Ranks are either processes (one per GPU, e.g. under torchrun) that
// rendezvous through a POSIX shared-memory control block and map each other's buffers with CUDA IPC, or threads of one
// process (dbt_dist_init_local: the file entry points use every visible GPU that way) that share the control block on the
// heap and reach each other's buffers through cudaDeviceEnablePeerAccess.  Everything else is the same code.
//
//  control block   host barriers and small all-gathers (splitter samples, P x buckets count matrices, IPC handles);
//  data            records and key columns cross NVLink as plain stores of our own kernels into the owner's staging
//                  buffer (kernels_dist.cu: gather_push_kernel = record gather + all-to-all in one kernel);
//  completion      stream-ordered epoch flags in peer memory (signal / wait kernels): no host round trip between a
//                  sender's last store and the owner's first load.
//
// Distributed MergeSort / EliminateDuplicates (the reference's DatabaseProject.cpp:172-381 and :94-170 on P GPUs):
//   every rank cuts the key space into P x Q sub-ranges with splitters chosen from a global sample (split on the key
//   only, so equal keys meet), groups its rows by sub-range and pushes them sub-range by sub-range; an owner runs the
//   ordinary single-GPU operator on sub-range q as soon as its P segments have landed, while the pushes of q+1.. are
//   still crossing NVLink, and appends the result to its output image (rows that do not fill a block wait for the next
//   sub-range).  The concatenation of the ranks' outputs is the globally sorted (duplicate-free) file.
#include "host_ctx.cuh"
#include "dist_internal.cuh"
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <fcntl.h>
#include <string>
#include <sys/mman.h>
#include <sys/stat.h>
#include <thread>
#include <unistd.h>
#include <vector>

namespace dbt {

constexpr size_t kMailBytes = 40 * 1024;
constexpr uint32_t kCtlMagic = 0xDB7D157Au;
constexpr uint32_t kSamplesPerRank = 4096;
constexpr double kHostTimeoutS = 120.0, kDevTimeoutS = 30.0;

struct Ctl { // lives in POSIX shared memory (processes) or on the heap (threads of one process)
    std::atomic<uint32_t> magic;
    uint32_t world;
    std::atomic<uint32_t> attached, detached;
    alignas(64) std::atomic<uint32_t> bar_count;
    alignas(64) std::atomic<uint32_t> bar_gen;
    alignas(64) std::atomic<uint32_t> abort_flag; // a rank that fails sets it: peers waiting in a barrier give up at once
    alignas(64) unsigned char mail[kMaxRanks][kMailBytes];
};

struct SharedBuf { // a device buffer every rank can store into
    void *own = nullptr;
    size_t cap = 0;
    void *peer[kMaxRanks] = {};
};

struct Layout {
    uint32_t P = 0, Q = 0;
    uint64_t cnt[kMaxRanks][64]; // cnt[src][bucket], bucket = owner * Q + sub-range
    // capacity mode (streaming scatter): the regions of an owner start at offsets fixed BEFORE the rows were counted
    // (sums of the senders' capacities), and the owner's own segment comes first in every region -- the sender wrote it
    // in place while it scattered; the other segments follow packed, in rank order
    bool by_capacity = false;
    uint64_t cap_blocks[kMaxRanks][64]; // cap_blocks[src][bucket]
    uint64_t seg_blocks(uint32_t src, uint32_t owner, uint32_t q) const { return (cnt[src][owner * Q + q] + kRpb - 1) / kRpb; }
    uint64_t region_blocks(uint32_t owner, uint32_t q) const {
        uint64_t b = 0;
        for (uint32_t s = 0; s < P; ++s) b += seg_blocks(s, owner, q);
        return b;
    }
    uint64_t region_cap(uint32_t owner, uint32_t q) const {
        uint64_t b = 0;
        for (uint32_t s = 0; s < P; ++s) b += cap_blocks[s][owner * Q + q];
        return b;
    }
    uint64_t region_blk0(uint32_t owner, uint32_t q) const {
        uint64_t b = 0;
        for (uint32_t k = 0; k < q; ++k) b += by_capacity ? region_cap(owner, k) : region_blocks(owner, k);
        return b;
    }
    uint64_t seg_blk0(uint32_t src, uint32_t owner, uint32_t q) const {
        uint64_t b = region_blk0(owner, q);
        if (by_capacity) {
            if (src == owner) return b;
            b += seg_blocks(owner, owner, q);
            for (uint32_t s = 0; s < src; ++s)
                if (s != owner) b += seg_blocks(s, owner, q);
            return b;
        }
        for (uint32_t s = 0; s < src; ++s) b += seg_blocks(s, owner, q);
        return b;
    }
    uint64_t total_blocks(uint32_t owner) const { return region_blk0(owner, Q); }
    uint64_t region_rows(uint32_t owner, uint32_t q) const {
        uint64_t r = 0;
        for (uint32_t s = 0; s < P; ++s) r += cnt[s][owner * Q + q];
        return r;
    }
};

} // namespace dbt

using namespace dbt;

struct dbt_dist {
    int rank = 0, world = 1, device = 0;
    bool local = false; // ranks are threads of this process
    Ctl *ctl = nullptr;
    size_t ctl_bytes = 0;
    cudaStream_t side = nullptr; // pushes (lowest priority: link-bound, must not keep SMs from the owner-side work)
    cudaStream_t work = nullptr; // owner-side operators on what has landed (highest priority)
    cudaEvent_t ev_c = nullptr, ev_q[kMaxSub] = {};
    Buf send[2]; // per-owner block images waiting for the copy engines
    cudaEvent_t ev_a = nullptr, ev_b = nullptr;
    SharedBuf stag[2], keys, flags;
    Buf ws, lists, carry;
    uint32_t *d_err = nullptr;
    uint32_t epoch = 0;
    uint32_t nsub = 0; // key sub-ranges per owner; 0 = automatic
    bool peers_enabled = false;
    double stats[16] = {0};
};

namespace dbt {

// ---- control block -------------------------------------------------------------------------------------
static int ctl_fail(dbt_dist *d, int rc) {
    if (rc && d->ctl) d->ctl->abort_flag.store(1, std::memory_order_release);
    return rc;
}
#define DIST_TRY(expr)                         \
    do {                                       \
        int rc__ = (expr);                     \
        if (rc__ != 0) return ctl_fail(d, rc__); \
    } while (0)
#define DIST_CUDA(expr)                                                                      \
    do {                                                                                     \
        cudaError_t e__ = (expr);                                                            \
        if (e__ != cudaSuccess) return ctl_fail(d, ::dbt::cuda_fail(e__, #expr, __FILE__, __LINE__)); \
    } while (0)

static int host_barrier(dbt_dist *d) {
    Ctl *c = d->ctl;
    const uint32_t W = (uint32_t)d->world;
    if (W == 1) return 0;
    const uint32_t gen = c->bar_gen.load(std::memory_order_acquire);
    if (c->bar_count.fetch_add(1, std::memory_order_acq_rel) + 1 == W) {
        c->bar_count.store(0, std::memory_order_relaxed);
        c->bar_gen.fetch_add(1, std::memory_order_release);
        return 0;
    }
    const auto t0 = std::chrono::steady_clock::now();
    uint32_t spins = 0;
    while (c->bar_gen.load(std::memory_order_acquire) == gen) {
        if (c->abort_flag.load(std::memory_order_acquire)) {
            set_error("dist: a peer rank failed");
            return DBT_ERR_TIMEOUT;
        }
        if (++spins > 2000) {
            std::this_thread::yield();
            if ((spins & 0xFFF) == 0 &&
                std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() > kHostTimeoutS) {
                set_error("dist: host barrier timed out (a rank is missing)");
                c->abort_flag.store(1, std::memory_order_release);
                return DBT_ERR_TIMEOUT;
            }
        }
    }
    return 0;
}

// out[r * bytes ..] = rank r's `mine` (bytes <= kMailBytes)
static int host_allgather(dbt_dist *d, const void *mine, size_t bytes, void *out) {
    if (bytes > kMailBytes) {
        set_error("dist: all-gather message too large");
        return DBT_ERR_ARG;
    }
    memcpy(d->ctl->mail[d->rank], mine, bytes);
    DBT_TRY(host_barrier(d));
    for (int r = 0; r < d->world; ++r) memcpy((char *)out + (size_t)r * bytes, d->ctl->mail[r], bytes);
    return host_barrier(d);
}

static int enable_peers(dbt_dist *d) {
    if (d->peers_enabled || !d->local) return 0;
    int devs[kMaxRanks];
    int mine = d->device;
    DBT_TRY(host_allgather(d, &mine, sizeof mine, devs));
    for (int r = 0; r < d->world; ++r) {
        if (devs[r] == d->device) continue;
        cudaError_t e = cudaDeviceEnablePeerAccess(devs[r], 0);
        if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) return cuda_fail(e, "cudaDeviceEnablePeerAccess", __FILE__, __LINE__);
        cudaGetLastError();
    }
    d->peers_enabled = true;
    return 0;
}

// (Re)allocate a buffer every rank can store into, collectively, when some rank needs more than it holds.
static int ensure_shared(dbt_dist *d, SharedBuf &b, size_t need_bytes, cudaStream_t main) {
    uint64_t want = need_bytes, all[kMaxRanks];
    DBT_TRY(host_allgather(d, &want, 8, all));
    uint64_t mx = 0;
    for (int r = 0; r < d->world; ++r) mx = std::max<uint64_t>(mx, all[r]);
    if (b.own && b.cap >= mx) return 0; // the same decision on every rank: capacities are always set from mx
    DBT_CUDA(cudaStreamSynchronize(main));
    DBT_CUDA(cudaStreamSynchronize(d->work));
    DBT_CUDA(cudaStreamSynchronize(d->side));
    DBT_TRY(host_barrier(d)); // nobody stores into the old buffers any more
    for (int r = 0; r < d->world; ++r) {
        if (r != d->rank && b.peer[r] && !d->local) cudaIpcCloseMemHandle(b.peer[r]);
        b.peer[r] = nullptr;
    }
    DBT_TRY(host_barrier(d)); // every mapping of my old buffer is closed before it is freed
    if (b.own) cudaFree(b.own);
    b.own = nullptr;
    b.cap = (size_t)(mx + mx / 4 + (1 << 20));
    DBT_CUDA(cudaMalloc(&b.own, b.cap));
    struct Exch {
        cudaIpcMemHandle_t h;
        uint64_t raw;
    } me, every[kMaxRanks];
    memset(&me, 0, sizeof me);
    me.raw = (uint64_t)(uintptr_t)b.own;
    if (!d->local && d->world > 1) DBT_CUDA(cudaIpcGetMemHandle(&me.h, b.own));
    DBT_TRY(host_allgather(d, &me, sizeof me, every));
    for (int r = 0; r < d->world; ++r) {
        if (r == d->rank) b.peer[r] = b.own;
        else if (d->local) b.peer[r] = (void *)(uintptr_t)every[r].raw;
        else DBT_CUDA(cudaIpcOpenMemHandle(&b.peer[r], every[r].h, cudaIpcMemLazyEnablePeerAccess));
    }
    return 0;
}

static int ensure_flags(dbt_dist *d, cudaStream_t main) {
    if (d->flags.own) return 0;
    DBT_TRY(ensure_shared(d, d->flags, kFlagWords * 4, main));
    DBT_CUDA(cudaMemsetAsync(d->flags.own, 0, d->flags.cap, main));
    DBT_CUDA(cudaStreamSynchronize(main));
    return host_barrier(d); // every rank's flags are zero before the first signal
}

static FlagPtrs flag_ptrs(const dbt_dist *d) {
    FlagPtrs f;
    memset(&f, 0, sizeof f);
    for (int r = 0; r < d->world; ++r) f.p[r] = (uint32_t *)d->flags.peer[r];
    return f;
}

// ---- one relation prepared for routing ----------------------------------------------------------------
struct Routed {
    Prepared p;
    const uint32_t *keys = nullptr; // routing word per row (recid | num | first four bytes of str): equal keys share it
    uint32_t *rows = nullptr;       // rows grouped by bucket
    uint64_t n = 0;
    uint64_t counts[64];
};

static int route_prepare(dbt_dist *d, Arena &A, const void *d_in, uint64_t nblocks, int field, cudaStream_t main, Routed *r) {
    DBT_TRY(prepare(d_in, nblocks, field, A, main, &r->p, field >= '2' ? 8u : 0u)); // (routing only needs the first word: 8-word keys always do)
    r->n = r->p.info.nrows;
    if (field == '2') {
        uint32_t *k = A.take<uint32_t>(std::max<uint64_t>(r->n, 1));
        if (!k) {
            set_error("dist: list space too small");
            return DBT_ERR_WORKSPACE;
        }
        DBT_TRY(gather_word(r->p.keys.str, r->p.keys.kw, 0, nullptr, k, r->n, main));
        r->keys = k;
    } else {
        r->keys = r->p.keys.w0;
    }
    r->rows = A.take<uint32_t>(std::max<uint64_t>(r->n, 1));
    if (!r->rows) {
        set_error("dist: list space too small");
        return DBT_ERR_WORKSPACE;
    }
    return 0;
}

static size_t route_bytes(uint64_t nblocks, int field) {
    const uint64_t n = nblocks * kRpb;
    return dbt_dev_ws_bytes_kw(DBT_OP_HASHJOIN, nblocks, 0, field, 8) / 2 + 3 * pad256(4 * n) + (field >= '2' ? pad256(32 * n) : 0) + (4 << 20);
}

// evenly spaced samples of the routing keys -> host
static int take_samples(dbt_dist *d, Arena &A, const Routed &r, uint32_t nsamp, cudaStream_t main, std::vector<uint32_t> *out) {
    out->clear();
    if (!r.n || !nsamp) return 0;
    uint32_t *ds = A.take<uint32_t>(nsamp);
    if (!ds) {
        set_error("dist: list space too small");
        return DBT_ERR_WORKSPACE;
    }
    DBT_TRY(launch_sample(r.keys, r.n, nsamp, ds, main));
    out->resize(nsamp);
    DBT_CUDA(cudaMemcpyAsync(out->data(), ds, 4 * (size_t)nsamp, cudaMemcpyDeviceToHost, main));
    DBT_CUDA(cudaStreamSynchronize(main));
    return 0;
}

// nbuckets-1 ascending splitters cutting the union of every rank's samples into equal parts (identical on all ranks)
static int choose_splitters(dbt_dist *d, const std::vector<uint32_t> &mine, uint32_t nbuckets, uint32_t *splitters) {
    struct Msg {
        uint32_t n;
        uint32_t v[2 * kSamplesPerRank];
    };
    static_assert(sizeof(Msg) <= kMailBytes, "sample message must fit a mail slot");
    std::vector<Msg> all(d->world);
    Msg me;
    me.n = (uint32_t)std::min<size_t>(mine.size(), 2 * kSamplesPerRank);
    memcpy(me.v, mine.data(), 4 * (size_t)me.n);
    DBT_TRY(host_allgather(d, &me, sizeof(Msg), all.data()));
    std::vector<uint32_t> s;
    for (int r = 0; r < d->world; ++r) s.insert(s.end(), all[r].v, all[r].v + all[r].n);
    std::sort(s.begin(), s.end());
    for (uint32_t j = 0; j + 1 < nbuckets; ++j) splitters[j] = s.empty() ? 0u : s[std::min<size_t>(s.size() - 1, (s.size() * (size_t)(j + 1)) / nbuckets)];
    return 0;
}

// Group the rows by bucket, agree on the layout, grow the buffers if needed, and move the records, sub-range by
// sub-range: the SMs gather a sub-range's rows into one contiguous block image per owner (own rows straight into the own
// staging buffer, the others into a send buffer), and the COPY ENGINES carry those images over NVLink on the side
// stream while the SMs go on with the next sub-range -- and, after the last one, with the owner-side work on what has
// already landed.  Measured on B200 (profiles/micro/p2p_scatter.cu): copy engine 780 GB/s per direction, 16-byte block
// stores from SMs 716, 140-byte record stores 335, remote record reads 390; and a push kernel that shares the SMs with
// the owner-side sort only alternates with it (profiles/r02_notes.md).  A flag to every owner follows each sub-range.
static int exchange_push(dbt_dist *d, int slot, const void *d_in, Routed &r, int mode, uint32_t Q, const uint32_t *splitters,
                         cudaStream_t main, Layout *lay) {
    const uint32_t P = (uint32_t)d->world, nb = P * Q;
    {
        const size_t wsb = dbt_dev_partition_ws_bytes((r.n + kRpb - 1) / kRpb + 1);
        DIST_TRY(d->ws.ensure(wsb));
        DIST_TRY(dbt_dev_partition_rows(r.keys, r.n, mode, splitters, nb, r.rows, r.counts, d->ws.p, d->ws.cap, main));
    }
    lay->P = P;
    lay->Q = Q;
    uint64_t mine[64], every[kMaxRanks][64];
    for (uint32_t b = 0; b < 64; ++b) mine[b] = b < nb ? r.counts[b] : 0;
    DIST_TRY(host_allgather(d, mine, sizeof mine, every));
    for (uint32_t s = 0; s < P; ++s)
        for (uint32_t b = 0; b < 64; ++b) lay->cnt[s][b] = every[s][b];
    DIST_TRY(ensure_shared(d, d->stag[slot], (size_t)lay->total_blocks(d->rank) * DBT_BLOCK_BYTES + 256, main));
    uint64_t send_blocks = 0;
    for (uint32_t q = 0; q < Q && Q > 1; ++q)
        for (uint32_t o = 0; o < P; ++o)
            if ((int)o != d->rank) send_blocks += lay->seg_blocks(d->rank, o, q);
    DIST_TRY(d->send[slot].ensure((size_t)send_blocks * DBT_BLOCK_BYTES + 256));
    const FlagPtrs fp = flag_ptrs(d);
    uint64_t row_off[64];
    row_off[0] = 0;
    for (uint32_t b = 1; b < nb; ++b) row_off[b] = row_off[b - 1] + r.counts[b - 1];
    uint64_t remote = 0, total = 0, send_off = 0;
    bool first_copy = true;
    const bool first_exchange = d->stats[2] == 0; // (joins exchange two relations: the NVLink phase starts with the first)
    for (uint32_t q = 0; q < Q; ++q) {
        PushPlan plan;
        memset(&plan, 0, sizeof plan);
        plan.nseg = P;
        struct Copy {
            void *dst;
            const void *src;
            size_t bytes;
            uint32_t owner;
        } copies[kMaxRanks];
        uint32_t ncopies = 0;
        bool direct_remote = false;
        for (uint32_t k = 0; k < P; ++k) {
            const uint32_t owner = (d->rank + 1 + k) % P; // rotated by rank: CTA i of every rank starts on a different owner
            const uint32_t b = owner * Q + q;
            const size_t bytes = (size_t)lay->seg_blocks(d->rank, owner, q) * DBT_BLOCK_BYTES;
            char *at_owner = (char *)d->stag[slot].peer[owner] + lay->seg_blk0(d->rank, owner, q) * DBT_BLOCK_BYTES;
            plan.seg[k].rows = r.rows + row_off[b];
            plan.seg[k].nrows = r.counts[b];
            total += bytes;
            if ((int)owner == d->rank || Q == 1) {
                // my own rows need no copy; and with a single region nothing can overlap the transfer, so the gather
                // kernel stores straight into the owner's staging buffer over NVLink (gather + exchange in one kernel:
                // 10.3 ms instead of 6.7 + 9.0 for 7 GB per direction at P = 2)
                plan.seg[k].out = (uint4 *)at_owner;
                if ((int)owner != d->rank) {
                    remote += bytes;
                    direct_remote = true;
                }
            } else {
                plan.seg[k].out = (uint4 *)((char *)d->send[slot].p + send_off);
                copies[ncopies++] = Copy{at_owner, (char *)d->send[slot].p + send_off, bytes, owner};
                send_off += bytes;
                remote += bytes;
            }
        }
        if (Q == 1 && first_copy && first_exchange) DIST_CUDA(cudaEventRecord(d->ev_a, main));
        DIST_TRY(launch_gather_push(d_in, r.p.row_slot, plan, main));
        DIST_CUDA(cudaEventRecord(d->ev_q[q], main));
        const uint32_t flag_idx = (slot ? kFlagSlot1 : 0) + q * kMaxRanks + d->rank;
        if (Q == 1) { // everything was stored by the gather kernel itself: one flag to everybody
            (void)direct_remote;
            DIST_CUDA(cudaStreamWaitEvent(d->side, d->ev_q[q], 0));
            DIST_TRY(launch_signal(fp, P, flag_idx, d->epoch, d->side));
        } else {
            // the copy engines carry this sub-range's images, one owner after the other on ONE stream, then one flag to
            // everybody.  Tried and measured slower (profiles/r02_notes.md): one copy stream per destination -- the
            // engines drain the streams unevenly, so the LAST segment of the first sub-range arrived near the end of the
            // whole transfer (first sub-range complete at 30 ms instead of 15 at P = 8; 42.8 vs 36.8 ms per step).
            DIST_CUDA(cudaStreamWaitEvent(d->side, d->ev_q[q], 0));
            if (first_copy && first_exchange) DIST_CUDA(cudaEventRecord(d->ev_a, d->side));
            for (uint32_t c = 0; c < ncopies; ++c)
                if (copies[c].bytes) DIST_CUDA(cudaMemcpyAsync(copies[c].dst, copies[c].src, copies[c].bytes, cudaMemcpyDeviceToDevice, d->side));
            DIST_TRY(launch_signal(fp, P, flag_idx, d->epoch, d->side));
        }
        first_copy = false;
    }
    DIST_CUDA(cudaEventRecord(d->ev_b, d->side));
    d->stats[1] += (double)remote;
    d->stats[2] += (double)total;
    return 0;
}


// The same exchange for a routing key that is a record word (fields '0', '1', '3'), without key columns, row lists or
// gathers on the sender (kernels_dist.cu: route_scatter_kernel): sample the image, agree on splitters and on CAPACITIES
// (estimated bucket sizes plus slack: the exact counts are only known after the pass), scatter in one streaming pass --
// own rows straight into the own staging buffer, the others into per-bucket images of the send buffer -- then agree on
// the exact counts, write the block headers and let the copy engines carry the images, sub-range by sub-range.
// *fell_back = true: a bucket outgrew its capacity on some rank (every rank sees it) and nothing was sent.
static int exchange_stream(dbt_dist *d, const void *d_in, uint64_t nblocks, int field, uint32_t Q, cudaStream_t main, Layout *lay,
                           bool *fell_back) {
    const uint32_t P = (uint32_t)d->world, nb = P * Q;
    const uint32_t word = field == '0' ? 0u : 1u;
    constexpr uint32_t kEst = 32768; // samples for the capacity estimate (a few hundred per bucket)
    *fell_back = false;
    // ---- samples -> splitters and local bucket estimates
    DIST_TRY(d->lists.ensure((size_t)kEst * 8 + 64 * 8 + 4096));
    Arena A(d->lists.p, d->lists.cap);
    uint32_t *d_samp = A.take<uint32_t>(kEst), *d_ok = A.take<uint32_t>(kEst);
    unsigned long long *d_cursor = A.take<unsigned long long>(64);
    uint32_t *d_over = A.take<uint32_t>(64);
    const uint32_t nsamp = nblocks ? kEst : 0;
    std::vector<uint32_t> hs(nsamp), hok(nsamp), samp;
    if (nsamp) {
        DIST_TRY(launch_sample_image(d_in, nblocks, word, nsamp, d_samp, d_ok, main));
        DIST_CUDA(cudaMemcpyAsync(hs.data(), d_samp, 4 * (size_t)nsamp, cudaMemcpyDeviceToHost, main));
        DIST_CUDA(cudaMemcpyAsync(hok.data(), d_ok, 4 * (size_t)nsamp, cudaMemcpyDeviceToHost, main));
    }
    DIST_CUDA(cudaMemsetAsync(d_cursor, 0, 64 * 8, main));
    DIST_CUDA(cudaMemsetAsync(d_over, 0, 4, main));
    DIST_CUDA(cudaStreamSynchronize(main));
    for (uint32_t i = 0; i < nsamp; ++i)
        if (hok[i]) samp.push_back(hs[i]);
    std::vector<uint32_t> few; // the all-gathered message holds 2 * kSamplesPerRank keys
    const size_t lim = std::max<uint32_t>(1024u, 2 * kSamplesPerRank / P);
    for (size_t i = 0; i < lim && !samp.empty(); ++i) few.push_back(samp[i * samp.size() / lim]);
    uint32_t splitters[64];
    DIST_TRY(choose_splitters(d, few, nb, splitters));
    uint64_t est[64] = {0};
    for (uint32_t k : samp) {
        uint32_t b = 0;
        for (uint32_t j = 0; j + 1 < nb; ++j) b += splitters[j] <= k ? 1u : 0u;
        ++est[b];
    }
    lay->P = P;
    lay->Q = Q;
    lay->by_capacity = true;
    uint64_t mycap[64], everycap[kMaxRanks][64];
    for (uint32_t b = 0; b < 64; ++b) {
        // rows of this shard expected in bucket b, + 25 % + 3 sigma of the sampling error + a few blocks
        const double frac = samp.empty() ? 0.0 : (double)est[b] / (double)samp.size();
        const double rows = frac * (double)nblocks * kRpb;
        const double sigma = samp.empty() ? 0.0 : std::sqrt(frac * (1.0 - frac) / (double)samp.size()) * (double)nblocks * kRpb;
        mycap[b] = b < nb && nblocks ? (uint64_t)((rows * 1.25 + 3.0 * sigma) / kRpb) + 8 : 0;
    }
    DIST_TRY(host_allgather(d, mycap, sizeof mycap, everycap));
    for (uint32_t s = 0; s < P; ++s)
        for (uint32_t b = 0; b < 64; ++b) {
            lay->cap_blocks[s][b] = everycap[s][b];
            lay->cnt[s][b] = 0;
        }
    DIST_TRY(ensure_shared(d, d->stag[0], (size_t)lay->total_blocks(d->rank) * DBT_BLOCK_BYTES + 256, main));
    uint64_t send_blocks = 0;
    for (uint32_t b = 0; b < nb; ++b)
        if ((int)(b / Q) != d->rank) send_blocks += mycap[b];
    DIST_TRY(d->send[0].ensure((size_t)send_blocks * DBT_BLOCK_BYTES + 256));
    // ---- scatter
    RoutePlan plan;
    memset(&plan, 0, sizeof plan);
    plan.nb = nb;
    plan.word = word;
    for (uint32_t j = 0; j + 1 < nb; ++j) plan.split[j] = splitters[j];
    uint64_t send_off[64], off = 0;
    for (uint32_t b = 0; b < nb; ++b) {
        const uint32_t owner = b / Q, q = b % Q;
        if ((int)owner == d->rank) {
            plan.dst[b] = (uint32_t *)((char *)d->stag[0].own + lay->region_blk0(owner, q) * DBT_BLOCK_BYTES);
            send_off[b] = 0;
        } else {
            plan.dst[b] = (uint32_t *)((char *)d->send[0].p + off * DBT_BLOCK_BYTES);
            send_off[b] = off;
            off += mycap[b];
        }
        plan.cap[b] = (uint32_t)std::min<uint64_t>(mycap[b] * kRpb, 0xFFFFFFFFull);
    }
    DIST_TRY(launch_route_scatter(d_in, nblocks, plan, d_cursor, d_over, main));
    struct Counts {
        uint64_t c[64];
        uint64_t overflow;
    } mine, every[kMaxRanks];
    uint32_t over = 0;
    DIST_CUDA(cudaMemcpyAsync(mine.c, d_cursor, 64 * 8, cudaMemcpyDeviceToHost, main));
    DIST_CUDA(cudaMemcpyAsync(&over, d_over, 4, cudaMemcpyDeviceToHost, main));
    DIST_CUDA(cudaStreamSynchronize(main));
    mine.overflow = over;
    DIST_TRY(host_allgather(d, &mine, sizeof mine, every));
    for (uint32_t s = 0; s < P; ++s)
        if (every[s].overflow) *fell_back = true;
    if (*fell_back) return 0;
    for (uint32_t s = 0; s < P; ++s)
        for (uint32_t b = 0; b < 64; ++b) lay->cnt[s][b] = every[s].c[b];
    // ---- headers of everything I filled, then the copies, sub-range by sub-range, one flag to every owner behind each
    SegFillPlan fill;
    memset(&fill, 0, sizeof fill);
    for (uint32_t b = 0; b < nb; ++b)
        if (mine.c[b]) fill.seg[fill.nseg++] = SegFill{plan.dst[b], mine.c[b]};
    DIST_TRY(launch_seg_headers(fill, main));
    DIST_CUDA(cudaEventRecord(d->ev_q[0], main));
    DIST_CUDA(cudaStreamWaitEvent(d->side, d->ev_q[0], 0));
    DIST_CUDA(cudaEventRecord(d->ev_a, d->side));
    const FlagPtrs fp = flag_ptrs(d);
    uint64_t remote = 0, total = 0;
    for (uint32_t q = 0; q < Q; ++q) {
        for (uint32_t k = 0; k < P; ++k) {
            const uint32_t owner = (d->rank + 1 + k) % P;
            const size_t bytes = (size_t)lay->seg_blocks(d->rank, owner, q) * DBT_BLOCK_BYTES;
            total += bytes;
            if ((int)owner == d->rank || !bytes) continue;
            char *at_owner = (char *)d->stag[0].peer[owner] + lay->seg_blk0(d->rank, owner, q) * DBT_BLOCK_BYTES;
            DIST_CUDA(cudaMemcpyAsync(at_owner, (char *)d->send[0].p + send_off[owner * Q + q] * DBT_BLOCK_BYTES, bytes, cudaMemcpyDeviceToDevice, d->side));
            remote += bytes;
        }
        DIST_TRY(launch_signal(fp, P, q * kMaxRanks + d->rank, d->epoch, d->side));
    }
    DIST_CUDA(cudaEventRecord(d->ev_b, d->side));
    d->stats[1] += (double)remote;
    d->stats[2] += (double)total;
    return 0;
}

static int wait_region(dbt_dist *d, int slot, uint32_t q, cudaStream_t main) {
    return launch_wait((const uint32_t *)d->flags.own, (uint32_t)d->world, (slot ? kFlagSlot1 : 0) + q * kMaxRanks, 1, d->epoch,
                       kDevTimeoutS, d->d_err, main);
}

static int finish_op(dbt_dist *d, cudaStream_t main) {
    uint32_t err = 0;
    DIST_CUDA(cudaMemcpyAsync(&err, d->d_err, 4, cudaMemcpyDeviceToHost, main));
    DIST_CUDA(cudaStreamSynchronize(main));
    DIST_CUDA(cudaStreamSynchronize(d->work));
    DIST_CUDA(cudaStreamSynchronize(d->side));
    stage_resolve();
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, d->ev_a, d->ev_b) == cudaSuccess) d->stats[0] = ms;
    else cudaGetLastError();
    if (err) {
        set_error("dist: a peer's records never arrived (flag wait timed out)");
        cudaMemsetAsync(d->d_err, 0, 4, main);
        return ctl_fail(d, DBT_ERR_TIMEOUT);
    }
    return 0;
}

static int begin_op(dbt_dist *d, cudaStream_t main) {
    DBT_CUDA(cudaSetDevice(d->device));
    if (d->ctl->abort_flag.load(std::memory_order_acquire)) {
        set_error("dist: the group was aborted by an earlier failure");
        return DBT_ERR_TIMEOUT;
    }
    DIST_TRY(enable_peers(d));
    DIST_TRY(ensure_flags(d, main));
    ++d->epoch;
    for (double &s : d->stats) s = 0;
    return 0;
}

// run `call(ws, bytes)` on the context's workspace; retry once with room for 120-byte string keys if asked to
template <class F> static int dist_with_ws(dbt_dist *d, int op, uint64_t nbr, uint64_t nbs, int field, size_t extra, F call) {
    DBT_TRY(d->ws.ensure(dbt_dev_ws_bytes_kw(op, nbr, nbs, field, 8) + extra));
    int rc = call(d->ws.p, d->ws.cap);
    if (rc == DBT_ERR_NEED_WIDE_KEYS) {
        DBT_TRY(d->ws.ensure(dbt_dev_ws_bytes_kw(op, nbr, nbs, field, 30) + extra));
        rc = call(d->ws.p, d->ws.cap);
    }
    return rc;
}

static double ms_since(std::chrono::steady_clock::time_point t0) {
    return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
}

static uint32_t pick_sub_ranges(dbt_dist *d, uint64_t max_blocks) {
    uint32_t Q = d->nsub; // 0 = automatic
    if (!Q)
        if (const char *e = getenv("DBT_DIST_SUBRANGES")) Q = (uint32_t)atoi(e);
    // small shards: the pipeline's per-sub-range launches would dominate (8 sub-ranges measured slower than 4 at P = 2, 4, 8)
    if (!Q) Q = (max_blocks < 20000) ? 1 : 4;
    Q = std::max<uint32_t>(1, std::min<uint32_t>(Q, std::min<uint32_t>(kMaxSub, 64u / (uint32_t)d->world)));
    return Q;
}

} // namespace dbt

// =====================================================================================================
extern "C" {

int dbt_dist_init(const char *session, int rank, int world, int device, dbt_dist **out) {
    if (!session || !out || world < 1 || world > (int)kMaxRanks || rank < 0 || rank >= world) {
        set_error("dbt_dist_init: bad arguments (1 <= world <= 16)");
        return DBT_ERR_ARG;
    }
    DBT_CUDA(cudaSetDevice(device));
    const std::string name = std::string("/dbt_") + session;
    Ctl *c = nullptr;
    const auto t0 = std::chrono::steady_clock::now();
    if (rank == 0) {
        shm_unlink(name.c_str());
        int fd = shm_open(name.c_str(), O_CREAT | O_EXCL | O_RDWR, 0600);
        if (fd < 0 || ftruncate(fd, sizeof(Ctl)) != 0) {
            set_error("dbt_dist_init: cannot create the shared control block " + name);
            return DBT_ERR_IO;
        }
        c = (Ctl *)mmap(nullptr, sizeof(Ctl), PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
        close(fd);
        if (c == MAP_FAILED) {
            set_error("dbt_dist_init: mmap failed");
            return DBT_ERR_IO;
        }
        memset((void *)c, 0, sizeof(Ctl));
        c->world = (uint32_t)world;
        c->magic.store(kCtlMagic, std::memory_order_release);
    } else {
        while (true) {
            int fd = shm_open(name.c_str(), O_RDWR, 0600);
            if (fd >= 0) {
                struct stat sb;
                if (fstat(fd, &sb) == 0 && (size_t)sb.st_size >= sizeof(Ctl)) {
                    c = (Ctl *)mmap(nullptr, sizeof(Ctl), PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
                    close(fd);
                    if (c != MAP_FAILED && c->magic.load(std::memory_order_acquire) == kCtlMagic) break;
                    if (c != MAP_FAILED) munmap((void *)c, sizeof(Ctl));
                } else {
                    close(fd);
                }
            }
            if (std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() > kHostTimeoutS) {
                set_error("dbt_dist_init: rank 0's control block never appeared");
                return DBT_ERR_TIMEOUT;
            }
            std::this_thread::sleep_for(std::chrono::milliseconds(2));
        }
        if ((int)c->world != world) {
            set_error("dbt_dist_init: world size differs from rank 0's");
            return DBT_ERR_ARG;
        }
    }
    dbt_dist *d = new dbt_dist();
    d->rank = rank;
    d->world = world;
    d->device = device;
    d->local = false;
    d->ctl = c;
    d->ctl_bytes = sizeof(Ctl);
    c->attached.fetch_add(1, std::memory_order_acq_rel);
    int rc = host_barrier(d);
    if (rank == 0) shm_unlink(name.c_str()); // every rank has it mapped: the name can go (nothing is left behind)
    int lo = 0, hi = 0;
    if (!rc && cudaDeviceGetStreamPriorityRange(&lo, &hi) != cudaSuccess) lo = hi = 0;
    if (const char *e = getenv("DBT_DIST_SIDE_PRIO")) lo = atoi(e) > 0 ? hi : 0; // experiment hook: 1 = highest, 0 = default
    if (!rc && cudaStreamCreateWithPriority(&d->side, cudaStreamNonBlocking, lo) != cudaSuccess) rc = DBT_ERR_CUDA;
    if (!rc && cudaStreamCreateWithPriority(&d->work, cudaStreamNonBlocking, hi) != cudaSuccess) rc = DBT_ERR_CUDA;
    if (!rc && (cudaEventCreate(&d->ev_a) != cudaSuccess || cudaEventCreate(&d->ev_b) != cudaSuccess ||
                cudaEventCreateWithFlags(&d->ev_c, cudaEventDisableTiming) != cudaSuccess))
        rc = DBT_ERR_CUDA;
    for (uint32_t q = 0; q < kMaxSub && !rc; ++q)
        if (cudaEventCreateWithFlags(&d->ev_q[q], cudaEventDisableTiming) != cudaSuccess) rc = DBT_ERR_CUDA;
    if (!rc && cudaMalloc((void **)&d->d_err, 256) != cudaSuccess) rc = DBT_ERR_CUDA;
    if (!rc && cudaMemset(d->d_err, 0, 256) != cudaSuccess) rc = DBT_ERR_CUDA;
    if (rc) {
        if (rc == DBT_ERR_CUDA) set_error("dbt_dist_init: CUDA resource creation failed");
        return rc;
    }
    *out = d;
    return 0;
}

int dbt_dist_init_local(int world, const int *devices, dbt_dist **out) {
    if (!devices || !out || world < 1 || world > (int)kMaxRanks) {
        set_error("dbt_dist_init_local: bad arguments (1 <= world <= 16)");
        return DBT_ERR_ARG;
    }
    Ctl *c = new Ctl();
    memset((void *)c, 0, sizeof(Ctl));
    c->world = (uint32_t)world;
    c->magic.store(kCtlMagic);
    for (int r = 0; r < world; ++r) {
        DBT_CUDA(cudaSetDevice(devices[r]));
        dbt_dist *d = new dbt_dist();
        d->rank = r;
        d->world = world;
        d->device = devices[r];
        d->local = true;
        d->ctl = c;
        d->ctl_bytes = 0;
        c->attached.fetch_add(1);
        int lo = 0, hi = 0;
        cudaDeviceGetStreamPriorityRange(&lo, &hi);
        DBT_CUDA(cudaStreamCreateWithPriority(&d->side, cudaStreamNonBlocking, lo));
        DBT_CUDA(cudaStreamCreateWithPriority(&d->work, cudaStreamNonBlocking, hi));
        DBT_CUDA(cudaEventCreateWithFlags(&d->ev_c, cudaEventDisableTiming));
        for (uint32_t q = 0; q < kMaxSub; ++q) DBT_CUDA(cudaEventCreateWithFlags(&d->ev_q[q], cudaEventDisableTiming));
        DBT_CUDA(cudaEventCreate(&d->ev_a));
        DBT_CUDA(cudaEventCreate(&d->ev_b));
        DBT_CUDA(cudaMalloc((void **)&d->d_err, 256));
        DBT_CUDA(cudaMemset(d->d_err, 0, 256));
        out[r] = d;
    }
    return 0;
}

int dbt_dist_destroy(dbt_dist *d) {
    if (!d) return 0;
    cudaSetDevice(d->device);
    cudaDeviceSynchronize();
    for (SharedBuf *b : {&d->stag[0], &d->stag[1], &d->keys, &d->flags}) {
        for (int r = 0; r < d->world; ++r)
            if (r != d->rank && b->peer[r] && !d->local) cudaIpcCloseMemHandle(b->peer[r]);
    }
    // peers may still have my buffers mapped: free them only after everybody has closed (best effort, bounded wait)
    if (d->ctl && !d->ctl->abort_flag.load()) host_barrier(d);
    for (SharedBuf *b : {&d->stag[0], &d->stag[1], &d->keys, &d->flags})
        if (b->own) cudaFree(b->own);
    d->ws.release();
    d->lists.release();
    if (d->d_err) cudaFree(d->d_err);
    if (d->side) cudaStreamDestroy(d->side);
    if (d->work) cudaStreamDestroy(d->work);
    if (d->ev_c) cudaEventDestroy(d->ev_c);
    for (cudaEvent_t e : d->ev_q)
        if (e) cudaEventDestroy(e);
    d->send[0].release();
    d->send[1].release();
    if (d->ev_a) cudaEventDestroy(d->ev_a);
    if (d->ev_b) cudaEventDestroy(d->ev_b);
    if (d->ctl) {
        const uint32_t gone = d->ctl->detached.fetch_add(1) + 1;
        if (d->ctl_bytes) munmap((void *)d->ctl, d->ctl_bytes);
        else if (gone == (uint32_t)d->world) delete d->ctl;
    }
    delete d;
    return 0;
}

int dbt_dist_rank(const dbt_dist *d) { return d ? d->rank : -1; }
int dbt_dist_world(const dbt_dist *d) { return d ? d->world : 0; }
int dbt_dist_barrier(dbt_dist *d) { return d ? host_barrier(d) : DBT_ERR_ARG; }
int dbt_dist_set_sub_ranges(dbt_dist *d, uint32_t q) {
    if (!d || q > kMaxSub) return DBT_ERR_ARG;
    d->nsub = q; // 0 = automatic
    return 0;
}
// Collective: release the staging, send, list and workspace buffers (they are grown again on demand).  Callers that go
// from one large workload to a differently shaped one use it to get the memory back.
int dbt_dist_trim(dbt_dist *d) {
    if (!d) return DBT_ERR_ARG;
    DBT_CUDA(cudaSetDevice(d->device));
    DBT_CUDA(cudaDeviceSynchronize());
    DBT_TRY(host_barrier(d)); // nobody stores into my buffers any more
    for (SharedBuf *b : {&d->stag[0], &d->stag[1], &d->keys}) {
        for (int r = 0; r < d->world; ++r) {
            if (r != d->rank && b->peer[r] && !d->local) cudaIpcCloseMemHandle(b->peer[r]);
            b->peer[r] = nullptr;
        }
    }
    DBT_TRY(host_barrier(d)); // every mapping of my buffers is closed
    for (SharedBuf *b : {&d->stag[0], &d->stag[1], &d->keys}) {
        if (b->own) cudaFree(b->own);
        b->own = nullptr;
        b->cap = 0;
    }
    d->send[0].release();
    d->send[1].release();
    d->ws.release();
    d->lists.release();
    d->carry.release();
    return host_barrier(d);
}
int dbt_dist_stats(const dbt_dist *d, double out[16]) {
    if (!d || !out) return DBT_ERR_ARG;
    memcpy(out, d->stats, sizeof d->stats);
    return 0;
}
// every rank contributes `bytes` (<= 32 KB); out receives world * bytes in rank order (host memory)
int dbt_dist_allgather_host(dbt_dist *d, const void *mine, size_t bytes, void *out) {
    if (!d || !mine || !out) return DBT_ERR_ARG;
    return host_allgather(d, mine, bytes, out);
}

int dbt_dist_sort(dbt_dist *d, const void *d_in, uint64_t nblocks, int field, int dedup, void *d_out, uint64_t out_capacity_blocks,
                  void *stream, uint64_t *out_rows, uint64_t *rows_received) {
    if (!d || (!d_in && nblocks) || !d_out || field < '0' || field > '3') {
        set_error("dbt_dist_sort: bad arguments");
        return DBT_ERR_ARG;
    }
    cudaStream_t main = (cudaStream_t)stream;
    DBT_TRY(begin_op(d, main));
    const auto t_begin = std::chrono::steady_clock::now();
    const uint32_t P = (uint32_t)d->world;
    uint64_t nbs[kMaxRanks], mxb = 0;
    DIST_TRY(host_allgather(d, &nblocks, 8, nbs));
    for (uint32_t r = 0; r < P; ++r) mxb = std::max(mxb, nbs[r]);
    const uint32_t Q = pick_sub_ranges(d, mxb);
    Layout lay;
    // The streaming scatter (DBT_DIST_STREAM=1; needs a record word as the routing key and shards large enough for the
    // pipeline) is exact (tests/dist_check.py passes with it) but opt-in: it halves the sender's HBM traffic (28 instead
    // of 58 GB per 14 GB shard) yet the rows of a bucket arrive in arbitrary block order, so the owner has to sort the
    // recid word as well (4 more onesweep passes per sub-range) -- measured 27.1 vs 24.8 ms per step at P = 2
    // (profiles/r02_notes.md).
    static const bool stream_ok = [] { const char *e = getenv("DBT_DIST_STREAM"); return e && atoi(e) != 0; }();
    bool streamed = false;
    if (stream_ok && field != '2' && Q > 1 && P > 1) {
        bool fell_back = false;
        DIST_TRY(exchange_stream(d, d_in, nblocks, field, Q, main, &lay, &fell_back));
        streamed = !fell_back;
        d->stats[9] = streamed ? 1 : -1;
        d->stats[4] = d->stats[5] = ms_since(t_begin);
    }
    if (!streamed) {
        lay = Layout();
        DIST_TRY(d->lists.ensure(route_bytes(nblocks, field) + (64 << 10)));
        Arena A(d->lists.p, d->lists.cap);
        Routed r;
        DIST_TRY(route_prepare(d, A, d_in, nblocks, field, main, &r));
        d->stats[4] = ms_since(t_begin);
        std::vector<uint32_t> samp;
        DIST_TRY(take_samples(d, A, r, (uint32_t)std::min<uint64_t>(std::max<uint32_t>(1024u, 2 * kSamplesPerRank / P), r.n), main, &samp));
        uint32_t splitters[64];
        DIST_TRY(choose_splitters(d, samp, P * Q, splitters));
        d->stats[5] = ms_since(t_begin);
        DIST_TRY(exchange_push(d, 0, d_in, r, 0, Q, splitters, main, &lay));
    }
    DIST_TRY(d->carry.ensure(1024));
    uint32_t *d_carry = (uint32_t *)d->carry.p;
    d->stats[3] = Q;

    // ---- owner side: sub-range q is processed as soon as its P segments have landed (the copy engines keep moving the
    // later sub-ranges meanwhile; the SMs are all ours).  Tried and measured slower: ordering the sub-ranges as they land
    // but deferring the HBM-heavy record gather to ONE gather after the last transfer (to keep HBM free for the copy
    // engines): they only went from ~480 to ~510 GB/s and the 6 ms gather is exposed (P=4: 34.5 vs 30 ms, P=2: 28.3 vs 25.0).
    d->stats[6] = ms_since(t_begin);
    uint64_t carry = 0, out_blk = 0, n_in_total = 0;
    const uint64_t cap_rows = out_capacity_blocks * kRpb;
    for (uint32_t q = 0; q < Q; ++q) {
        DIST_TRY(wait_region(d, 0, q, main));
        const uint64_t nb = lay.region_blocks(d->rank, q), blk0 = lay.region_blk0(d->rank, q);
        if (!nb) continue;
        const char *img = (const char *)d->stag[0].own + blk0 * DBT_BLOCK_BYTES;
        const int op = dedup ? DBT_OP_DEDUP : DBT_OP_SORT;
        uint64_t n_in = 0, n_out = 0;
        int rc = dist_with_ws(d, op, nb, 0, field, pad256(4 * (nb * kRpb + 256)) + 4096, [&](void *wsp, size_t wsb) -> int {
            Arena W(wsp, wsb);
            uint32_t *rows = nullptr, *row_slot = nullptr;
            DBT_TRY(sorted_rows(img, nb, field, dedup != 0, W, main, &rows, &row_slot, &n_in, &n_out));
            const uint64_t avail = carry + n_out, emit = avail / kRpb * kRpb;
            if (out_blk * kRpb + avail > cap_rows) {
                set_error("dbt_dist_sort: output capacity too small");
                return DBT_ERR_WORKSPACE;
            }
            uint32_t *comb = W.take<uint32_t>(avail + 1);
            if (!comb) {
                set_error("dbt_dist_sort: workspace too small");
                return DBT_ERR_WORKSPACE;
            }
            DBT_TRY(launch_concat_slots(d_carry, (uint32_t)carry, rows, row_slot, n_out, (uint32_t)(blk0 * kRpb), comb, main));
            if (emit) DBT_TRY(gather_records(d->stag[0].own, comb, nullptr, emit, (char *)d_out + out_blk * DBT_BLOCK_BYTES, main, 0, (uint32_t)out_blk));
            if (avail - emit) DBT_CUDA(cudaMemcpyAsync(d_carry, comb + emit, 4 * (avail - emit), cudaMemcpyDeviceToDevice, main));
            carry = avail - emit;
            out_blk += emit / kRpb;
            return 0;
        });
        DIST_TRY(rc);
        n_in_total += n_in;
        if (q == 0) d->stats[7] = ms_since(t_begin);
    }
    if (carry) DIST_TRY(gather_records(d->stag[0].own, d_carry, nullptr, carry, (char *)d_out + out_blk * DBT_BLOCK_BYTES, main, 0, (uint32_t)out_blk));
    if (out_rows) *out_rows = out_blk * kRpb + carry;
    if (rows_received) *rows_received = n_in_total;
    const int rc_fin = finish_op(d, (cudaStream_t)stream);
    d->stats[8] = ms_since(t_begin);
    return rc_fin;
}

int dbt_dist_hashjoin(dbt_dist *d, const void *d_r, uint64_t nbr, const void *d_s, uint64_t nbs, int field, void *d_out,
                      uint64_t out_capacity_blocks, void *stream, uint64_t *nres) {
    if (!d || !nres || field < '0' || field > '3') {
        set_error("dbt_dist_hashjoin: bad arguments");
        return DBT_ERR_ARG;
    }
    cudaStream_t main = (cudaStream_t)stream;
    DBT_TRY(begin_op(d, main));
    const uint32_t P = (uint32_t)d->world;
    *nres = 0;
    if (field == '0' || field == '1') {
        // semi-join on u32 keys: replicate R's KEYS (4 bytes per R row) and probe every S shard where it lies -- no S
        // record crosses the fabric, key skew cannot unbalance the ranks, and the ranks' outputs concatenate in S file
        // order like the reference's (DatabaseProject.cpp:561-640 walks S in file order)
        DIST_TRY(d->lists.ensure(route_bytes(nbr, field)));
        Arena A(d->lists.p, d->lists.cap);
        Prepared pr;
        DIST_TRY(prepare(d_r, nbr, field, A, main, &pr));
        uint64_t nr = pr.info.nrows, all[kMaxRanks], off = 0, N = 0;
        DIST_TRY(host_allgather(d, &nr, 8, all));
        for (uint32_t r = 0; r < P; ++r) {
            if ((int)r < d->rank) off += all[r];
            N += all[r];
        }
        DIST_TRY(ensure_shared(d, d->keys, 4 * N + 256, main));
        KeyDst dst;
        memset(&dst, 0, sizeof dst);
        for (uint32_t r = 0; r < P; ++r) dst.p[r] = (uint32_t *)d->keys.peer[r] + off;
        DIST_CUDA(cudaEventRecord(d->ev_a, main));
        DIST_TRY(launch_broadcast_keys(pr.keys.w0, nr, dst, P, main));
        DIST_TRY(launch_signal(flag_ptrs(d), P, kFlagKeys + d->rank, d->epoch, main));
        DIST_TRY(launch_wait((const uint32_t *)d->flags.own, P, kFlagKeys, 1, d->epoch, kDevTimeoutS, d->d_err, main));
        DIST_CUDA(cudaEventRecord(d->ev_b, main));
        d->stats[1] = 4.0 * (double)nr * (P - 1);
        d->stats[2] = 4.0 * (double)nr * P;
        const uint64_t nkeys_blocks = (N + kRpb - 1) / kRpb + 1;
        int rc = dist_with_ws(d, DBT_OP_HASHJOIN, nkeys_blocks, nbs, field, 0, [&](void *wsp, size_t wsb) {
            return dbt_dev_semijoin_keys((const uint32_t *)d->keys.own, N, d_s, nbs, field, d_out, out_capacity_blocks, wsp, wsb, main, nres);
        });
        DIST_TRY(rc);
        return finish_op(d, main);
    }
    // str / composite keys: both relations are hash-partitioned on the key's first word, then the ordinary operator runs
    DIST_TRY(d->lists.ensure(route_bytes(nbr, field) + route_bytes(nbs, field)));
    Arena A(d->lists.p, d->lists.cap);
    Routed rr, rs;
    Layout lr, ls;
    DIST_TRY(route_prepare(d, A, d_r, nbr, field, main, &rr));
    DIST_TRY(route_prepare(d, A, d_s, nbs, field, main, &rs));
    DIST_TRY(exchange_push(d, 0, d_r, rr, 1, 1, nullptr, main, &lr));
    DIST_TRY(exchange_push(d, 1, d_s, rs, 1, 1, nullptr, main, &ls));
    DIST_TRY(wait_region(d, 0, 0, main));
    DIST_TRY(wait_region(d, 1, 0, main));
    const uint64_t b_r = lr.total_blocks(d->rank), b_s = ls.total_blocks(d->rank);
    const size_t extra = dbt_dev_hashjoin_ws_bytes(b_r, b_s, field, 8, out_capacity_blocks) - dbt_dev_ws_bytes_kw(DBT_OP_HASHJOIN, b_r, b_s, field, 8);
    int rc = dist_with_ws(d, DBT_OP_HASHJOIN, b_r, b_s, field, extra, [&](void *wsp, size_t wsb) {
        return dbt_dev_hashjoin(d->stag[0].own, b_r, d->stag[1].own, b_s, field, d_out, out_capacity_blocks, wsp, wsb, main, nres);
    });
    DIST_TRY(rc);
    return finish_op(d, main);
}

int dbt_dist_mergejoin(dbt_dist *d, const void *d_r, uint64_t nbr, const void *d_s, uint64_t nbs, int field, void *d_out,
                       uint64_t out_capacity_blocks, void *stream, uint64_t *res) {
    if (!d || !res || field < '0' || field > '3') {
        set_error("dbt_dist_mergejoin: bad arguments");
        return DBT_ERR_ARG;
    }
    cudaStream_t main = (cudaStream_t)stream;
    DBT_TRY(begin_op(d, main));
    const uint32_t P = (uint32_t)d->world;
    DIST_TRY(d->lists.ensure(route_bytes(nbr, field) + route_bytes(nbs, field)));
    Arena A(d->lists.p, d->lists.cap);
    Routed rr, rs;
    Layout lr, ls;
    DIST_TRY(route_prepare(d, A, d_r, nbr, field, main, &rr));
    DIST_TRY(route_prepare(d, A, d_s, nbs, field, main, &rs));
    // both relations use the SAME key-range splitters (from a joint sample), so equal keys of R and S meet on one rank
    std::vector<uint32_t> s1, s2;
    DIST_TRY(take_samples(d, A, rr, (uint32_t)std::min<uint64_t>(kSamplesPerRank, rr.n), main, &s1));
    DIST_TRY(take_samples(d, A, rs, (uint32_t)std::min<uint64_t>(kSamplesPerRank, rs.n), main, &s2));
    s1.insert(s1.end(), s2.begin(), s2.end());
    uint32_t splitters[64];
    DIST_TRY(choose_splitters(d, s1, P, splitters));
    DIST_TRY(exchange_push(d, 0, d_r, rr, 0, 1, splitters, main, &lr));
    DIST_TRY(exchange_push(d, 1, d_s, rs, 0, 1, splitters, main, &ls));
    DIST_TRY(wait_region(d, 0, 0, main));
    DIST_TRY(wait_region(d, 1, 0, main));
    const uint64_t b_r = lr.total_blocks(d->rank), b_s = ls.total_blocks(d->rank);
    if (std::min(b_r, b_s) > out_capacity_blocks) {
        set_error("dbt_dist_mergejoin: output capacity too small");
        return ctl_fail(d, DBT_ERR_WORKSPACE);
    }
    int rc = dist_with_ws(d, DBT_OP_MERGEJOIN, b_r, b_s, field, 0, [&](void *wsp, size_t wsb) {
        // no side images here: "1outfile.bin" / "2outfile.bin" belong to the file API (NULL skips their two gathers)
        return dbt_dev_mergejoin(d->stag[0].own, b_r, d->stag[1].own, b_s, field, nullptr, nullptr, d_out, wsp, wsb, main, res);
    });
    DIST_TRY(rc);
    return finish_op(d, main);
}

// Host-side self test of the control block (no CUDA): rendezvous, barriers, all-gathers, splitter choice and layout
// arithmetic must agree on every rank.  Returns 0 and a checksum that is identical on all ranks.
int dbt_dist_selftest_host(const char *session, int rank, int world, uint64_t *checksum) {
    if (!session || world < 1 || world > (int)kMaxRanks || rank < 0 || rank >= world) return DBT_ERR_ARG;
    const std::string name = std::string("/dbt_") + session;
    Ctl *c = nullptr;
    const auto t0 = std::chrono::steady_clock::now();
    if (rank == 0) {
        shm_unlink(name.c_str());
        int fd = shm_open(name.c_str(), O_CREAT | O_EXCL | O_RDWR, 0600);
        if (fd < 0 || ftruncate(fd, sizeof(Ctl)) != 0) return DBT_ERR_IO;
        c = (Ctl *)mmap(nullptr, sizeof(Ctl), PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
        close(fd);
        if (c == MAP_FAILED) return DBT_ERR_IO;
        memset((void *)c, 0, sizeof(Ctl));
        c->world = (uint32_t)world;
        c->magic.store(kCtlMagic, std::memory_order_release);
    } else {
        while (true) {
            int fd = shm_open(name.c_str(), O_RDWR, 0600);
            if (fd >= 0) {
                struct stat sb;
                if (fstat(fd, &sb) == 0 && (size_t)sb.st_size >= sizeof(Ctl)) {
                    c = (Ctl *)mmap(nullptr, sizeof(Ctl), PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
                    close(fd);
                    if (c != MAP_FAILED && c->magic.load(std::memory_order_acquire) == kCtlMagic) break;
                    if (c != MAP_FAILED) munmap((void *)c, sizeof(Ctl));
                } else {
                    close(fd);
                }
            }
            if (std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() > 30.0) return DBT_ERR_TIMEOUT;
            std::this_thread::sleep_for(std::chrono::milliseconds(1));
        }
    }
    dbt_dist dd;
    dbt_dist *d = &dd;
    d->rank = rank;
    d->world = world;
    d->ctl = c;
    int rc = host_barrier(d);
    if (rank == 0) shm_unlink(name.c_str());
    uint64_t sum = 0;
    for (int it = 0; it < 200 && !rc; ++it) { // barriers + all-gathers in a tight loop: every rank must see every message of every round
        uint64_t mine[4] = {(uint64_t)rank, (uint64_t)it, (uint64_t)rank * 1000003ull + it, 7}, all[kMaxRanks][4];
        rc = host_allgather(d, mine, sizeof mine, all);
        for (int r = 0; r < world && !rc; ++r) {
            if (all[r][0] != (uint64_t)r || all[r][1] != (uint64_t)it || all[r][2] != (uint64_t)r * 1000003ull + it) rc = DBT_ERR_IO;
            sum = sum * 1099511628211ull + all[r][2];
        }
    }
    if (!rc) { // splitters from per-rank samples: identical on all ranks, ascending
        std::vector<uint32_t> samp(512);
        for (uint32_t i = 0; i < 512; ++i) samp[i] = (uint32_t)((i * 2654435761u) ^ (rank * 40503u));
        uint32_t sp[64];
        const uint32_t nbk = (uint32_t)std::min(64, world * 4);
        rc = choose_splitters(d, samp, nbk, sp);
        for (uint32_t j = 0; j + 1 < nbk && !rc; ++j) {
            if (j && sp[j] < sp[j - 1]) rc = DBT_ERR_IO;
            sum = sum * 1099511628211ull + sp[j];
        }
    }
    if (!rc) { // layout arithmetic: what a sender computes for an owner equals what the owner computes for itself
        Layout lay;
        lay.P = (uint32_t)world;
        lay.Q = 3;
        for (int s = 0; s < world; ++s)
            for (uint32_t b = 0; b < 64; ++b) lay.cnt[s][b] = (b < (uint32_t)world * 3) ? (uint64_t)((s * 131 + b * 17) % 977) : 0;
        for (int o = 0; o < world && !rc; ++o) {
            uint64_t run = 0;
            for (uint32_t q = 0; q < 3; ++q)
                for (int s = 0; s < world; ++s) {
                    if (lay.seg_blk0((uint32_t)s, (uint32_t)o, q) != run) rc = DBT_ERR_IO;
                    run += lay.seg_blocks((uint32_t)s, (uint32_t)o, q);
                }
            if (lay.total_blocks((uint32_t)o) != run) rc = DBT_ERR_IO;
            sum = sum * 1099511628211ull + run;
        }
    }
    int rc2 = host_barrier(d);
    munmap((void *)c, sizeof(Ctl));
    d->ctl = nullptr;
    if (checksum) *checksum = sum;
    return rc ? rc : rc2;
}

} // extern "C"
