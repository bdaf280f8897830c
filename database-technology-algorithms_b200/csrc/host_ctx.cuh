// host_ctx.cuh -- state shared by the host-scope operators (host_ops.cu) and their out-of-core forms (host_ooc.cu).
#pragma once
#include "dbt_internal.cuh"
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <vector>

namespace dbt {

// ---- cached device / pinned buffers (grown on demand, reused across calls) ------------------
struct Buf {
    void *p = nullptr;
    size_t cap = 0;
    bool pinned = false;
    bool malloced = false; // plain host memory (the out-of-core runs when pinned memory is short)
    int ensure(size_t bytes) {
        if (bytes <= cap) return 0;
        release();
        size_t want = bytes + bytes / 8 + (1 << 20);
        if (pinned) DBT_CUDA(cudaHostAlloc(&p, want, cudaHostAllocDefault));
        else DBT_CUDA(cudaMalloc(&p, want));
        cap = want;
        return 0;
    }
    void release() {
        if (p) {
            if (malloced) free(p);
            else if (pinned) cudaFreeHost(p);
            else cudaFree(p);
        }
        p = nullptr;
        cap = 0;
        malloced = false;
    }
};
struct HostCtx {
    Buf in_r, in_s, out0, out1, out2, ws;
    Buf cols, runs; // out-of-core: resident key columns (device), sorted runs (pinned host)
    Buf side_r, side_s, side_o; // out-of-core merge join: dedup(R), dedup(S), result when the caller passes no buffer (host)
    Buf stage[2];
    cudaStream_t st = nullptr;
    cudaStream_t st2 = nullptr;            // out-of-core: results go home on this stream while the next chunk comes in
    cudaEvent_t ev[2] = {nullptr, nullptr};
    cudaEvent_t packed[2] = {nullptr, nullptr}, landed[2] = {nullptr, nullptr}; // out-of-core double buffering
    int device = -1;
    size_t cached_device_bytes() const { // device memory this context already holds and would reuse for the next job
        return in_r.cap + in_s.cap + out0.cap + out1.cap + out2.cap + ws.cap + cols.cap;
    }
    int init(int dev) {
        int n = 0;
        cudaError_t e = cudaGetDeviceCount(&n);
        if (e != cudaSuccess || n == 0) {
            cudaGetLastError();
            set_error("no CUDA device visible: libdbt_b200 has no CPU fallback");
            return DBT_ERR_CUDA;
        }
        if (dev < 0 || dev >= n) {
            set_error("bad device index");
            return DBT_ERR_ARG;
        }
        DBT_CUDA(cudaSetDevice(dev));
        if (device != dev) {
            if (device >= 0) { // moving to another GPU: drop the old device's buffers
                set_error("host-scope operators are bound to the first device they were used on");
                return DBT_ERR_UNSUPPORTED;
            }
            device = dev;
            DBT_CUDA(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
            DBT_CUDA(cudaStreamCreateWithFlags(&st2, cudaStreamNonBlocking));
            for (int i = 0; i < 2; ++i) {
                DBT_CUDA(cudaEventCreateWithFlags(&packed[i], cudaEventDisableTiming));
                DBT_CUDA(cudaEventCreateWithFlags(&landed[i], cudaEventDisableTiming));
            }
            DBT_CUDA(cudaEventCreateWithFlags(&ev[0], cudaEventDisableTiming));
            DBT_CUDA(cudaEventCreateWithFlags(&ev[1], cudaEventDisableTiming));
            stage[0].pinned = stage[1].pinned = true;
        }
        return 0;
    }
};

// `on`: the stream to copy on (default: the context's main stream)
int upload(HostCtx &c, const void *h, void *d, size_t bytes, cudaStream_t on = nullptr);
int download(HostCtx &c, const void *d, void *h, size_t bytes, cudaStream_t on = nullptr);
inline size_t blocks_for(uint64_t rows) { return (size_t)((rows + kRpb - 1) / kRpb); }

// out-of-core forms (host_ooc.cu).  chunk_blocks: the largest image that is processed in one piece.
// 0 = the job fits (in the device's FREE memory plus what the context already caches), run it in-core
uint64_t ooc_chunk_blocks(int op, uint64_t nbr, uint64_t nbs, int field, size_t cached_bytes = 0);
int ooc_sort(HostCtx &c, const void *h_in, uint64_t nblocks, int field, void *h_out, bool dedup, uint64_t chunk_blocks,
             uint64_t *nrows, uint64_t *nunique);
int ooc_hashjoin(HostCtx &c, const void *h_in_r, uint64_t nbr, const void *h_in_s, uint64_t nbs, int field, void *h_out,
                 uint64_t out_capacity_blocks, uint64_t chunk_blocks, uint64_t *nres);
int ooc_mergejoin(HostCtx &c, const void *h_in_r, uint64_t nbr, const void *h_in_s, uint64_t nbs, int field, void *h_out_ur,
                  void *h_out_us, void *h_out, uint64_t chunk_blocks, uint64_t res[4]);

// several GPUs of one box, single process (host_multi.cu); DBT_DEVICES = "all" | "<count>" | "i,j,k"
int multi_devices(std::vector<int> *devs); // number of devices to use (0 = the single-GPU path)
int multi_sort(const std::vector<int> &devs, const void *h_in, uint64_t nblocks, int field, bool dedup, void *h_out, uint64_t *nrows_in,
               uint64_t *nrows_out);
int multi_hashjoin(const std::vector<int> &devs, const void *h_r, uint64_t nbr, const void *h_s, uint64_t nbs, int field, void *h_out,
                   uint64_t out_capacity_blocks, uint64_t *nres);
int multi_mergejoin(const std::vector<int> &devs, const void *h_r, uint64_t nbr, const void *h_s, uint64_t nbs, int field, void *h_out_ur,
                    void *h_out_us, void *h_out, uint64_t res[4]);

} // namespace dbt

extern "C" size_t dbt_dev_ws_bytes_kw(int op, uint64_t nbr, uint64_t nbs, int field, uint32_t kw);

// run `call(ws_ptr, ws_bytes)`; when it reports that 120-byte string keys are needed, retry once with a larger workspace.
// extra_bytes: on top of the operator's bound (HashJoin field '3' with an output larger than S)
template <class F> static int with_workspace(dbt::HostCtx &c, int op, uint64_t nbr, uint64_t nbs, int field, F call,
                                             size_t extra_bytes = 0) {
    DBT_TRY(c.ws.ensure(dbt_dev_ws_bytes_kw(op, nbr, nbs, field, 8) + extra_bytes));
    int rc = call(c.ws.p, c.ws.cap);
    if (rc == DBT_ERR_NEED_WIDE_KEYS) {
        DBT_TRY(c.ws.ensure(dbt_dev_ws_bytes_kw(op, nbr, nbs, field, 30) + extra_bytes));
        rc = call(c.ws.p, c.ws.cap);
    }
    return rc;
}
