// host_multi.cu -- the dbtproj.h operators on SEVERAL GPUs of one box, single process: what MergeSort(),
// EliminateDuplicates(), MergeJoin() and HashJoin() (host_ops.cu) run when DBT_DEVICES names more than one device.
//
// One host thread per GPU drives one rank of a dbt_dist_init_local group (csrc/dist.cu: peer access, no IPC); rank r takes
// the r-th contiguous range of the input file's blocks, the ranks run the collective operator, and their outputs --
// each a packed image with rank-local block numbering -- are assembled into ONE packed image in rank order: a rank's
// rows that complete the previous rank's last block and its own last partial block are "loose" (fewer than 200 per
// rank, placed on the host), everything in between is re-blocked on the device with the shift that makes it start on a
// block boundary of the global file and goes home with one copy.  The result is byte-identical to the single-GPU one.
#include "host_ctx.cuh"
#include <mutex>
#include <string>
#include <thread>
#include <vector>

struct dbt_dist;
extern "C" {
int dbt_dist_init_local(int world, const int *devices, dbt_dist **out);
int dbt_dist_sort(dbt_dist *d, const void *d_in, uint64_t nblocks, int field, int dedup, void *d_out, uint64_t out_capacity_blocks,
                  void *stream, uint64_t *out_rows, uint64_t *rows_received);
int dbt_dist_hashjoin(dbt_dist *d, const void *d_r, uint64_t nbr, const void *d_s, uint64_t nbs, int field, void *d_out,
                      uint64_t out_capacity_blocks, void *stream, uint64_t *nres);
int dbt_dist_mergejoin(dbt_dist *d, const void *d_r, uint64_t nbr, const void *d_s, uint64_t nbs, int field, void *d_out,
                       uint64_t out_capacity_blocks, void *stream, uint64_t *res);
}

namespace dbt {

uint64_t host_walk_reads(const void *ur, uint64_t nr, const void *us, uint64_t ns, int field); // host_ooc.cu

struct RankBufs {
    Buf in_r, in_s, out, tmp, list;
    cudaStream_t st = nullptr;
};
struct MultiCtx {
    std::vector<int> devs;
    std::vector<dbt_dist *> ranks;
    std::vector<RankBufs> bufs;
};
static MultiCtx g_multi;
static std::mutex g_multi_mu;

// DBT_DEVICES = "all" | "<count>" | "i,j,k" ; unset, empty or a single device => the single-GPU path
int multi_devices(std::vector<int> *devs) {
    devs->clear();
    const char *e = getenv("DBT_DEVICES");
    if (!e || !*e) return 0;
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    const std::string s(e);
    if (s == "all") {
        for (int i = 0; i < n; ++i) devs->push_back(i);
    } else if (s.find(',') == std::string::npos) {
        const int k = atoi(e);
        for (int i = 0; i < k && i < n; ++i) devs->push_back(i);
    } else {
        size_t p = 0;
        while (p < s.size()) {
            const size_t q = s.find(',', p);
            const int v = atoi(s.substr(p, q == std::string::npos ? std::string::npos : q - p).c_str());
            if (v >= 0 && v < n) devs->push_back(v);
            if (q == std::string::npos) break;
            p = q + 1;
        }
    }
    if (devs->size() > 16) devs->resize(16);
    if (devs->size() < 2) devs->clear();
    return (int)devs->size();
}

static int multi_init(const std::vector<int> &devs) {
    if (g_multi.devs == devs && !g_multi.ranks.empty()) return 0;
    if (!g_multi.ranks.empty()) {
        set_error("DBT_DEVICES changed inside one process: the multi-GPU group is bound to the first device list");
        return DBT_ERR_UNSUPPORTED;
    }
    g_multi.ranks.assign(devs.size(), nullptr);
    DBT_TRY(dbt_dist_init_local((int)devs.size(), devs.data(), g_multi.ranks.data()));
    g_multi.bufs.resize(devs.size());
    for (size_t r = 0; r < devs.size(); ++r) {
        DBT_CUDA(cudaSetDevice(devs[r]));
        DBT_CUDA(cudaStreamCreateWithFlags(&g_multi.bufs[r].st, cudaStreamNonBlocking));
    }
    g_multi.devs = devs;
    return 0;
}

__global__ void __launch_bounds__(256) iota_offset_kernel(uint32_t *out, uint64_t n, uint32_t off) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) out[i] = (uint32_t)i + off;
}

struct Loose { // rows placed on the host: (global row, 140 bytes)
    uint64_t grow;
    unsigned char rec[DBT_RECORD_BYTES];
};

// Rank r's packed output image (c rows, rank-local numbering) -> its place in the global packed image h_out, whose rows
// [G, G + c) it owns.  Interior blocks are re-blocked on the device; head / tail rows come back as loose rows.
static int place_rank_output(RankBufs &b, const char *d_img, uint64_t c, uint64_t G, char *h_out, std::vector<Loose> *loose) {
    cudaStream_t st = b.st;
    const uint64_t head = std::min<uint64_t>((kRpb - G % kRpb) % kRpb, c);
    const uint64_t k = (c - head) / kRpb, tail = c - head - k * kRpb;
    auto fetch_rows = [&](uint64_t row0, uint64_t cnt) -> int { // local rows [row0, row0 + cnt) -> loose (global rows G + row0 ...)
        for (uint64_t i = 0; i < cnt;) {
            const uint64_t row = row0 + i, blk = row / kRpb, e = row % kRpb, run = std::min<uint64_t>(cnt - i, kRpb - e);
            const size_t at = loose->size();
            loose->resize(at + run);
            std::vector<unsigned char> raw(run * DBT_RECORD_BYTES);
            DBT_CUDA(cudaMemcpyAsync(raw.data(), d_img + blk * DBT_BLOCK_BYTES + 8 + e * DBT_RECORD_BYTES, raw.size(), cudaMemcpyDeviceToHost, st));
            DBT_CUDA(cudaStreamSynchronize(st));
            for (uint64_t j = 0; j < run; ++j) {
                (*loose)[at + j].grow = G + row + j;
                memcpy((*loose)[at + j].rec, raw.data() + j * DBT_RECORD_BYTES, DBT_RECORD_BYTES);
            }
            i += run;
        }
        return 0;
    };
    DBT_TRY(fetch_rows(0, head));
    DBT_TRY(fetch_rows(head + k * kRpb, tail));
    if (k) {
        const uint64_t gblk = (G + head) / kRpb;
        const char *src = d_img;
        if (head) { // shift by `head` rows: one gather through an offset row list
            DBT_TRY(b.tmp.ensure(k * DBT_BLOCK_BYTES));
            DBT_TRY(b.list.ensure(4 * k * kRpb));
            const int grid = (int)std::min<uint64_t>((k * kRpb + 255) / 256, 148 * 8);
            iota_offset_kernel<<<grid, 256, 0, st>>>((uint32_t *)b.list.p, k * kRpb, (uint32_t)head);
            count_launch();
            DBT_KERNEL_CHECK();
            DBT_TRY(gather_records(d_img, (const uint32_t *)b.list.p, nullptr, k * kRpb, b.tmp.p, st, 0, (uint32_t)gblk));
            src = (const char *)b.tmp.p;
        }
        DBT_CUDA(cudaMemcpyAsync(h_out + gblk * DBT_BLOCK_BYTES, src, k * DBT_BLOCK_BYTES, cudaMemcpyDeviceToHost, st));
        DBT_CUDA(cudaStreamSynchronize(st));
        if (!head && gblk) // already aligned: only the block numbers change
            for (uint64_t j = 0; j < k; ++j) {
                const uint32_t id = (uint32_t)(gblk + j);
                memcpy(h_out + (gblk + j) * DBT_BLOCK_BYTES, &id, 4);
            }
    }
    return 0;
}

// loose rows -> their blocks of the global image (headers of those blocks included)
static void place_loose(const std::vector<Loose> &loose, uint64_t total_rows, char *h_out) {
    uint64_t last_blk = ~0ull;
    for (const Loose &l : loose) {
        const uint64_t blk = l.grow / kRpb, e = l.grow % kRpb;
        char *bp = h_out + blk * DBT_BLOCK_BYTES;
        if (blk != last_blk) {
            const uint32_t live = (uint32_t)std::min<uint64_t>(kRpb, total_rows - blk * kRpb);
            uint32_t hdr[2] = {(uint32_t)blk, live}, trl[2] = {1u, live};
            memcpy(bp, hdr, 8);
            memcpy(bp + 14008, trl, 8);
            if (live < kRpb) memset(bp + 8 + (size_t)live * DBT_RECORD_BYTES, 0, (size_t)(kRpb - live) * DBT_RECORD_BYTES);
            last_blk = blk;
        }
        memcpy(bp + 8 + e * DBT_RECORD_BYTES, l.rec, DBT_RECORD_BYTES);
    }
}

// run fn(rank) on one thread per rank; the first failure's code and message win
template <class F> static int on_all_ranks(F fn) {
    const size_t P = g_multi.ranks.size();
    std::vector<int> rc(P, 0);
    std::vector<std::string> msg(P);
    std::vector<std::thread> th;
    for (size_t r = 0; r < P; ++r)
        th.emplace_back([&, r] {
            if (cudaSetDevice(g_multi.devs[r]) != cudaSuccess) {
                rc[r] = DBT_ERR_CUDA;
                msg[r] = "cudaSetDevice failed";
                return;
            }
            rc[r] = fn((int)r);
            if (rc[r]) msg[r] = dbt_last_error();
        });
    for (auto &t : th) t.join();
    // report the rank that failed first-hand (the others only learn "a peer rank failed" from the control block)
    int first = -1;
    for (size_t r = 0; r < P; ++r)
        if (rc[r] && (first < 0 || (msg[first].find("a peer rank failed") != std::string::npos && msg[r].find("a peer rank failed") == std::string::npos)))
            first = (int)r;
    if (first >= 0) {
        set_error("rank " + std::to_string(first) + ": " + msg[first]);
        return rc[first];
    }
    return 0;
}

static void shard_of(uint64_t nblocks, size_t P, size_t r, uint64_t *b0, uint64_t *nb) {
    const uint64_t per = (nblocks + P - 1) / P;
    *b0 = std::min<uint64_t>(nblocks, per * r);
    *nb = std::min<uint64_t>(per, nblocks - *b0);
}

// assemble the ranks' outputs (device images in bufs[r].out, counts[r] rows) into h_out
static int assemble(const std::vector<uint64_t> &counts, char *h_out, uint64_t *total_rows) {
    const size_t P = counts.size();
    std::vector<uint64_t> G(P + 1, 0);
    for (size_t r = 0; r < P; ++r) G[r + 1] = G[r] + counts[r];
    std::vector<std::vector<Loose>> loose(P);
    DBT_TRY(on_all_ranks([&](int r) { return place_rank_output(g_multi.bufs[r], (const char *)g_multi.bufs[r].out.p, counts[r], G[r], h_out, &loose[r]); }));
    std::vector<Loose> all;
    for (auto &l : loose) all.insert(all.end(), l.begin(), l.end());
    place_loose(all, G[P], h_out);
    *total_rows = G[P];
    return 0;
}

// ---- operators on host images (pinned), all GPUs of the group ---------------------------------------------
int multi_sort(const std::vector<int> &devs, const void *h_in, uint64_t nblocks, int field, bool dedup, void *h_out, uint64_t *nrows_in,
               uint64_t *nrows_out) {
    std::lock_guard<std::mutex> lk(g_multi_mu);
    DBT_TRY(multi_init(devs));
    const size_t P = devs.size();
    std::vector<uint64_t> counts(P, 0), recv(P, 0);
    DBT_TRY(on_all_ranks([&](int r) {
        RankBufs &b = g_multi.bufs[r];
        uint64_t b0, nb;
        shard_of(nblocks, P, r, &b0, &nb);
        const uint64_t cap = nblocks / P + nblocks / (2 * P) + 1024; // a rank owns about 1/P of the rows (sampled splitters: headroom)
        DBT_TRY(b.in_r.ensure(std::max<uint64_t>(nb, 1) * DBT_BLOCK_BYTES));
        DBT_TRY(b.out.ensure(cap * DBT_BLOCK_BYTES));
        DBT_CUDA(cudaMemcpyAsync(b.in_r.p, (const char *)h_in + b0 * DBT_BLOCK_BYTES, nb * DBT_BLOCK_BYTES, cudaMemcpyHostToDevice, b.st));
        return dbt_dist_sort(g_multi.ranks[r], b.in_r.p, nb, field, dedup ? 1 : 0, b.out.p, cap, b.st, &counts[r], &recv[r]);
    }));
    uint64_t total = 0, received = 0;
    DBT_TRY(assemble(counts, (char *)h_out, &total));
    for (uint64_t x : recv) received += x;
    if (nrows_in) *nrows_in = received;
    if (nrows_out) *nrows_out = total;
    return 0;
}

// fields '0' and '1' only (replicated build keys: the ranks' outputs concatenate in S file order, as the reference emits)
int multi_hashjoin(const std::vector<int> &devs, const void *h_r, uint64_t nbr, const void *h_s, uint64_t nbs, int field, void *h_out,
                   uint64_t out_capacity_blocks, uint64_t *nres) {
    std::lock_guard<std::mutex> lk(g_multi_mu);
    DBT_TRY(multi_init(devs));
    const size_t P = devs.size();
    std::vector<uint64_t> counts(P, 0);
    DBT_TRY(on_all_ranks([&](int r) {
        RankBufs &b = g_multi.bufs[r];
        uint64_t r0, rn, s0, sn;
        shard_of(nbr, P, r, &r0, &rn);
        shard_of(nbs, P, r, &s0, &sn);
        DBT_TRY(b.in_r.ensure(std::max<uint64_t>(rn, 1) * DBT_BLOCK_BYTES));
        DBT_TRY(b.in_s.ensure(std::max<uint64_t>(sn, 1) * DBT_BLOCK_BYTES));
        DBT_TRY(b.out.ensure(std::max<uint64_t>(sn, 1) * DBT_BLOCK_BYTES)); // set semantics: at most every S row of the shard
        DBT_CUDA(cudaMemcpyAsync(b.in_r.p, (const char *)h_r + r0 * DBT_BLOCK_BYTES, rn * DBT_BLOCK_BYTES, cudaMemcpyHostToDevice, b.st));
        DBT_CUDA(cudaMemcpyAsync(b.in_s.p, (const char *)h_s + s0 * DBT_BLOCK_BYTES, sn * DBT_BLOCK_BYTES, cudaMemcpyHostToDevice, b.st));
        return dbt_dist_hashjoin(g_multi.ranks[r], b.in_r.p, rn, b.in_s.p, sn, field, b.out.p, std::max<uint64_t>(sn, 1), b.st, &counts[r]);
    }));
    uint64_t total = 0;
    for (uint64_t x : counts) total += x;
    if (nres) *nres = total;
    if (total > out_capacity_blocks * kRpb) {
        set_error("hashjoin: output capacity too small (nres returned)");
        return DBT_ERR_WORKSPACE;
    }
    return assemble(counts, (char *)h_out, &total);
}

int multi_mergejoin(const std::vector<int> &devs, const void *h_r, uint64_t nbr, const void *h_s, uint64_t nbs, int field, void *h_out_ur,
                    void *h_out_us, void *h_out, uint64_t res[4]) {
    // the side files "1outfile.bin" / "2outfile.bin" (DatabaseProject.cpp:385-394) are two distributed dedups
    uint64_t nin = 0, nur = 0, nus = 0;
    DBT_TRY(multi_sort(devs, h_r, nbr, field, true, h_out_ur, &nin, &nur));
    DBT_TRY(multi_sort(devs, h_s, nbs, field, true, h_out_us, &nin, &nus));
    std::lock_guard<std::mutex> lk(g_multi_mu);
    const size_t P = devs.size();
    std::vector<uint64_t> counts(P, 0);
    DBT_TRY(on_all_ranks([&](int r) {
        RankBufs &b = g_multi.bufs[r];
        uint64_t r0, rn, s0, sn;
        shard_of(nbr, P, r, &r0, &rn);
        shard_of(nbs, P, r, &s0, &sn);
        const uint64_t cap = std::max(nbr, nbs) / P + std::max(nbr, nbs) / (2 * P) + 1024;
        DBT_TRY(b.in_r.ensure(std::max<uint64_t>(rn, 1) * DBT_BLOCK_BYTES));
        DBT_TRY(b.in_s.ensure(std::max<uint64_t>(sn, 1) * DBT_BLOCK_BYTES));
        DBT_TRY(b.out.ensure(cap * DBT_BLOCK_BYTES));
        DBT_CUDA(cudaMemcpyAsync(b.in_r.p, (const char *)h_r + r0 * DBT_BLOCK_BYTES, rn * DBT_BLOCK_BYTES, cudaMemcpyHostToDevice, b.st));
        DBT_CUDA(cudaMemcpyAsync(b.in_s.p, (const char *)h_s + s0 * DBT_BLOCK_BYTES, sn * DBT_BLOCK_BYTES, cudaMemcpyHostToDevice, b.st));
        uint64_t rr[4] = {0, 0, 0, 0};
        DBT_TRY(dbt_dist_mergejoin(g_multi.ranks[r], b.in_r.p, rn, b.in_s.p, sn, field, b.out.p, cap, b.st, rr));
        counts[r] = rr[0];
        return 0;
    }));
    uint64_t total = 0;
    DBT_TRY(assemble(counts, (char *)h_out, &total));
    res[0] = total;
    res[1] = nur;
    res[2] = nus;
    res[3] = host_walk_reads(h_out_ur, nur, h_out_us, nus, field); // the reference walk's later block reads, from the two side images
    return 0;
}

} // namespace dbt
