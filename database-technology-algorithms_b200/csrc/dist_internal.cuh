// dist_internal.cuh -- shared between the multi-GPU host layer (dist.cu) and its kernels (kernels_dist.cu).
#pragma once
#include "dbt_internal.cuh"

namespace dbt {

constexpr uint32_t kMaxRanks = 16;   // GPUs of one box
constexpr uint32_t kMaxSub = 16;     // key sub-ranges per owner (the pipeline depth of the distributed sort)
constexpr uint32_t kFlagWords = 1024; // per rank: [0, 256) record regions (sub-range * kMaxRanks + src), [256, 272) key columns, ...
constexpr uint32_t kFlagKeys = 256;
constexpr uint32_t kFlagSlot1 = 512; // the second staging buffer's regions (joins exchange two relations)

struct PushSeg {
    const uint32_t *rows; // row ids (file order rows of the local image) going to this owner, in file order
    uint64_t nrows;
    uint4 *out;           // where the owner wants this segment's block image (peer-mapped)
};
struct PushPlan {
    PushSeg seg[kMaxRanks];
    uint32_t nseg;
};
struct FlagPtrs {
    uint32_t *p[kMaxRanks];
};
struct KeyDst {
    uint32_t *p[kMaxRanks];
};

// ---- streaming scatter (the sender side of the pipelined sort when the routing key is a record word) ----------
struct RoutePlan {
    uint32_t nb;         // buckets (owner * Q + sub-range)
    uint32_t word;       // record word holding the routing key: 0 recid, 1 num
    uint32_t split[63];  // bucket of a key = number of splitters <= key (as dest_kernel, device_ops.cu)
    uint32_t pad_;
    uint32_t *dst[64];   // block image of bucket b (send buffer, or the own staging buffer)
    uint32_t cap[64];    // its capacity in rows
};
struct SegFill { // one block image whose headers are written once its row count is known
    uint32_t *img;
    uint64_t nrows;
};
struct SegFillPlan {
    SegFill seg[64];
    uint32_t nseg;
};
// one streaming pass over the image: every live row goes to the next free slot of its bucket's image
// (d_cursor[b]: rows handed out so far, may exceed cap[b] -- then *d_overflow is set and the surplus rows are dropped)
int launch_route_scatter(const void *d_in, uint64_t nblocks, const RoutePlan &plan, unsigned long long *d_cursor /*[64], zeroed*/,
                         uint32_t *d_overflow /*zeroed*/, cudaStream_t st);
int launch_seg_headers(const SegFillPlan &plan, cudaStream_t st);
// nsamp routing keys from evenly spaced live rows of the image (d_ok[i] = 0 where the probed block was empty)
int launch_sample_image(const void *d_in, uint64_t nblocks, uint32_t word, uint32_t nsamp, uint32_t *d_keys, uint32_t *d_ok,
                        cudaStream_t st);

int launch_gather_push(const void *d_in, const uint32_t *d_row_slot, const PushPlan &plan, cudaStream_t st);
int launch_signal(const FlagPtrs &peers, uint32_t nranks, uint32_t idx, uint32_t epoch, cudaStream_t st);
int launch_wait(const uint32_t *flags, uint32_t nranks, uint32_t idx0, uint32_t stride, uint32_t epoch, double timeout_s,
                uint32_t *d_err, cudaStream_t st);
int launch_broadcast_keys(const uint32_t *d_src, uint64_t n, const KeyDst &dst, uint32_t nranks, cudaStream_t st);
int launch_sample(const uint32_t *d_keys, uint64_t n, uint32_t nsamples, uint32_t *d_out, cudaStream_t st);
int launch_concat_slots(const uint32_t *d_carry, uint32_t ncarry, const uint32_t *d_rows, const uint32_t *d_row_slot, uint64_t n,
                        uint32_t slot_base, uint32_t *d_out, cudaStream_t st);

} // namespace dbt
