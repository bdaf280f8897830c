// kernels_dist.cu -- device side of the multi-GPU exchange: records and key columns travel over
// NVLink as plain stores into the owner's memory (peer-mapped), and "it has landed" travels the same way as a flag.
//
//  * gather_push_kernel  the record gather / block packer of kernels_gather.cu with one destination image per owner:
//                        CTA i builds block i / P of the image for owner i % P in shared memory from gathered 4-byte
//                        words and stores it with 16-byte vectors straight into that owner's staging buffer -- gather
//                        and all-to-all are one kernel, with no send buffer, and every launch spreads its stores over
//                        all P peers at once.  CTAs are short-lived (one block each) so that a higher-priority stream
//                        (the owner's sort of what has already landed) gets SMs as soon as it asks.
//  * signal / wait       stream-ordered flags in peer memory: the sender's signal kernel runs after its push kernel
//                        on the same stream (a kernel boundary orders the stores system-wide) and writes the step's
//                        epoch into every owner's flag word; the owner's wait kernel spins on its own flag words.
//                        No host round trip, no collective call on the data path.
//  * broadcast_keys      HashJoin on u32 keys replicates the build side's key column instead of moving S.
#include "dbt_internal.cuh"
#include "dist_internal.cuh"
#include <algorithm>

namespace dbt {

constexpr int kPushThreads = 256;

__global__ void __launch_bounds__(kPushThreads)
gather_push_kernel(const uint32_t *__restrict__ in, const uint32_t *__restrict__ row_slot, PushPlan plan) {
    __shared__ __align__(16) uint32_t stage[kBlockWords];
    __shared__ uint64_t src[kRpb];
    const int tid = threadIdx.x;
    const uint32_t s = blockIdx.x % plan.nseg;
    const uint64_t lb = blockIdx.x / plan.nseg; // block of segment s (segment-local numbering)
    const uint64_t nrows = plan.seg[s].nrows;
    if (lb * kRpb >= nrows) return;
    const uint32_t *rows = plan.seg[s].rows + lb * kRpb;
    const uint32_t cnt = (uint32_t)min((uint64_t)kRpb, nrows - lb * kRpb);
    if (tid < (int)cnt) {
        const uint32_t row = rows[tid];
        const uint64_t slot = row_slot ? row_slot[row] : row;
        src[tid] = slot_word(slot);
    }
    if (tid == 0) {
        stage[0] = (uint32_t)lb;
        stage[1] = cnt;
        stage[kTrailerWord] = 1;
        stage[kTrailerWord + 1] = cnt;
    }
    __syncthreads();
    const uint32_t nwords = cnt * kRecWords;
    constexpr int kPerThread = (kRpb * kRecWords + kPushThreads - 1) / kPushThreads;
    uint32_t v[kPerThread];
#pragma unroll
    for (int k = 0; k < kPerThread; ++k) { // every load of the thread in flight before the first shared-memory store
        const uint32_t idx = tid + k * kPushThreads;
        const uint32_t rec = idx / kRecWords;
        const uint32_t w = idx - rec * kRecWords;
        v[k] = (idx < nwords) ? __ldg(in + src[rec] + w) : 0u;
    }
#pragma unroll
    for (int k = 0; k < kPerThread; ++k) {
        const uint32_t idx = tid + k * kPushThreads;
        if (idx < kRpb * kRecWords) stage[kEntriesWord + idx] = v[k];
    }
    __syncthreads();
    const uint4 *sv = reinterpret_cast<const uint4 *>(stage);
    uint4 *dst = plan.seg[s].out + lb * kBlockVec4;
    for (uint32_t i = tid; i < kBlockVec4; i += kPushThreads) dst[i] = sv[i];
}

int launch_gather_push(const void *d_in, const uint32_t *d_row_slot, const PushPlan &plan, cudaStream_t st) {
    uint64_t max_nb = 0;
    for (uint32_t s = 0; s < plan.nseg; ++s) max_nb = std::max<uint64_t>(max_nb, (plan.seg[s].nrows + kRpb - 1) / kRpb);
    if (!max_nb) return 0;
    const uint64_t grid = max_nb * plan.nseg;
    if (grid >= (1ull << 31)) {
        set_error("push: too many blocks for one launch");
        return DBT_ERR_UNSUPPORTED;
    }
    StageScope sc(ST_GATHER, st);
    gather_push_kernel<<<(unsigned)grid, kPushThreads, 0, st>>>((const uint32_t *)d_in, d_row_slot, plan);
    count_launch();
    DBT_KERNEL_CHECK();
    return 0;
}

// ---- flags ------------------------------------------------------------------------------------------
__global__ void signal_kernel(FlagPtrs peers, uint32_t nranks, uint32_t idx, uint32_t epoch) {
    const uint32_t d = threadIdx.x;
    if (d < nranks) {
        __threadfence_system();
        asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(peers.p[d] + idx), "r"(epoch) : "memory");
    }
}
// lane s waits until flags[idx0 + s * stride] has reached `epoch` (wrap-safe); a flag that never arrives sets *err
// after ~timeout_ns instead of hanging the device
__global__ void wait_kernel(const uint32_t *flags, uint32_t nranks, uint32_t idx0, uint32_t stride, uint32_t epoch,
                            unsigned long long timeout_ns, uint32_t *err) {
    const uint32_t s = threadIdx.x;
    if (s >= nranks) return;
    const uint32_t *f = flags + idx0 + (uint64_t)s * stride;
    unsigned long long t0;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    uint32_t spins = 0;
    while (true) {
        uint32_t v;
        asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(f) : "memory");
        if ((int32_t)(v - epoch) >= 0) break;
        if ((++spins & 0x3FF) == 0) {
            unsigned long long t1;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
            if (t1 - t0 > timeout_ns) {
                atomicExch(err, 1u + s);
                break;
            }
        }
        __nanosleep(200);
    }
}

int launch_signal(const FlagPtrs &peers, uint32_t nranks, uint32_t idx, uint32_t epoch, cudaStream_t st) {
    signal_kernel<<<1, 32, 0, st>>>(peers, nranks, idx, epoch);
    count_launch();
    DBT_KERNEL_CHECK();
    return 0;
}
int launch_wait(const uint32_t *flags, uint32_t nranks, uint32_t idx0, uint32_t stride, uint32_t epoch, double timeout_s,
                uint32_t *d_err, cudaStream_t st) {
    wait_kernel<<<1, 32, 0, st>>>(flags, nranks, idx0, stride, epoch, (unsigned long long)(timeout_s * 1e9), d_err);
    count_launch();
    DBT_KERNEL_CHECK();
    return 0;
}

// ---- key column -> every peer's key buffer (16-byte vectors when the destination offset allows, else words) ----
__global__ void __launch_bounds__(256)
broadcast_keys_kernel(const uint32_t *__restrict__ src, uint64_t n, KeyDst dst, uint32_t nranks) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const uint32_t v = src[i];
        for (uint32_t d = 0; d < nranks; ++d) dst.p[d][i] = v;
    }
}
int launch_broadcast_keys(const uint32_t *d_src, uint64_t n, const KeyDst &dst, uint32_t nranks, cudaStream_t st) {
    if (!n) return 0;
    const int grid = (int)std::min<uint64_t>((n + 255) / 256, 148 * 8);
    broadcast_keys_kernel<<<grid, 256, 0, st>>>(d_src, n, dst, nranks);
    count_launch();
    DBT_KERNEL_CHECK();
    return 0;
}

// ---- small list kernels of the pipelined sort -------------------------------------------------------------
// evenly spaced samples of a key column
__global__ void sample_kernel(const uint32_t *__restrict__ keys, uint64_t n, uint32_t nsamples, uint32_t *__restrict__ out) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < nsamples) out[i] = keys[(uint64_t)i * n / nsamples];
}
int launch_sample(const uint32_t *d_keys, uint64_t n, uint32_t nsamples, uint32_t *d_out, cudaStream_t st) {
    sample_kernel<<<(nsamples + 255) / 256, 256, 0, st>>>(d_keys, n, nsamples, d_out);
    count_launch();
    DBT_KERNEL_CHECK();
    return 0;
}
// out = carry[0..ncarry) ++ (row_slot ? row_slot[rows[i]] : rows[i]) + slot_base: an ordered row list of one landed
// region, as slots of the whole staging image, behind the rows still waiting for their output block to fill up
__global__ void __launch_bounds__(256)
concat_slots_kernel(const uint32_t *__restrict__ carry, uint32_t ncarry, const uint32_t *__restrict__ rows,
                    const uint32_t *__restrict__ row_slot, uint64_t n, uint32_t slot_base, uint32_t *__restrict__ out) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < ncarry + n; i += stride) {
        if (i < ncarry) out[i] = carry[i];
        else {
            const uint32_t r = rows[i - ncarry];
            out[i] = (row_slot ? row_slot[r] : r) + slot_base;
        }
    }
}
int launch_concat_slots(const uint32_t *d_carry, uint32_t ncarry, const uint32_t *d_rows, const uint32_t *d_row_slot, uint64_t n,
                        uint32_t slot_base, uint32_t *d_out, cudaStream_t st) {
    if (!(ncarry + n)) return 0;
    const int grid = (int)std::min<uint64_t>((ncarry + n + 255) / 256, 148 * 8);
    concat_slots_kernel<<<grid, 256, 0, st>>>(d_carry, ncarry, d_rows, d_row_slot, n, slot_base, d_out);
    count_launch();
    DBT_KERNEL_CHECK();
    return 0;
}

} // namespace dbt
