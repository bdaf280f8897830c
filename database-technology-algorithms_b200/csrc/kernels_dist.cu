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
        v[k] = (idx < nwords) ? ld_sparse(in + src[rec] + w) : 0u;
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


// ---- streaming scatter ---------------------------------------------------------------------------------
// The sender side of the pipelined sort used to be: key extraction (one full read of the image for 8 bytes per row),
// a partition pass over the row ids, and one random 140-byte gather per row into per-owner block images (267 bytes
// read per row with 128-byte fills, 204 with the 64-byte fills of round 2, kernels_gather.cu) -- 58 GB (now ~45 GB) of HBM
// traffic for a 14 GB shard.  When the routing key is a record
// word (recid / num), ONE streaming pass does it: blocks arrive in shared memory by cp.async.bulk (as in the
// streaming semi-join), every live row finds its bucket among the splitters, takes the next free slot of the bucket's
// image (shared-memory ranks inside the block, one global atomic per block and bucket) and is copied there by one warp.
// Rows of a block that share a bucket are contiguous in the destination; the order inside a bucket is arbitrary (the
// owner sorts).  28 GB instead of 58, and no list space.
constexpr int kRtThreads = 128;
constexpr int kRtStages = 3;

__device__ __forceinline__ uint32_t rt_smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void __launch_bounds__(kRtThreads)
route_scatter_kernel(const uint32_t *__restrict__ img, uint64_t nblocks, const __grid_constant__ RoutePlan plan,
                     unsigned long long *__restrict__ cursor, uint32_t *__restrict__ overflow) {
    extern __shared__ __align__(128) unsigned char rt_raw[];
    uint32_t(*stage)[kBlockWords] = reinterpret_cast<uint32_t(*)[kBlockWords]>(rt_raw);
    uint64_t *mbar = reinterpret_cast<uint64_t *>(rt_raw + sizeof(uint32_t) * kBlockWords * kRtStages);
    __shared__ uint32_t s_split[64], s_cnt[64];
    __shared__ unsigned long long s_base[64];
    __shared__ uint32_t *s_dst[kRpb];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint64_t first = blockIdx.x, step = gridDim.x;
    const uint64_t mine = first < nblocks ? (nblocks - first + step - 1) / step : 0;
    auto issue = [&](uint64_t block, int sidx) {
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(rt_smem_u32(&mbar[sidx])), "r"((uint32_t)DBT_BLOCK_BYTES) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(rt_smem_u32(stage[sidx])),
                     "l"(img + block * kBlockWords), "r"((uint32_t)DBT_BLOCK_BYTES), "r"(rt_smem_u32(&mbar[sidx]))
                     : "memory");
    };
    if (tid == 0) {
        for (int i = 0; i < kRtStages; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(rt_smem_u32(&mbar[i])), "r"(1));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        for (uint64_t k = 0; k < (uint64_t)(kRtStages - 1) && k < mine; ++k) issue(first + k * step, (int)k);
    }
    if (tid < 64) {
        s_split[tid] = tid + 1 < (int)plan.nb ? plan.split[tid] : 0xFFFFFFFFu;
        s_cnt[tid] = 0;
    }
    __syncthreads();
    const uint32_t nsplit = plan.nb - 1;
    for (uint64_t k = 0; k < mine; ++k) {
        if (tid == 0 && k + kRtStages - 1 < mine) issue(first + (k + kRtStages - 1) * step, (int)((k + kRtStages - 1) % kRtStages));
        const int sidx = (int)(k % kRtStages);
        {
            const uint32_t parity = (uint32_t)((k / kRtStages) & 1);
            uint32_t ok = 0;
            while (!ok)
                asm volatile("{\n\t.reg .pred q;\n\tmbarrier.try_wait.parity.shared::cta.b64 q, [%1], %2;\n\tselp.u32 %0, 1, 0, q;\n\t}"
                             : "=r"(ok)
                             : "r"(rt_smem_u32(&mbar[sidx])), "r"(parity)
                             : "memory");
        }
        const uint32_t *blk = stage[sidx];
        const uint32_t nres = min(blk[1], kRpb);
        uint32_t b = 0, r = 0;
        if (tid < (int)nres) {
            const uint32_t key = blk[kEntriesWord + tid * kRecWords + plan.word];
            for (uint32_t j = 0; j < nsplit; ++j) b += (s_split[j] <= key) ? 1u : 0u;
            r = atomicAdd(&s_cnt[b], 1u);
        }
        __syncthreads();
        if (tid < (int)plan.nb) {
            const uint32_t c = s_cnt[tid];
            unsigned long long base = 0;
            if (c) {
                base = atomicAdd(&cursor[tid], (unsigned long long)c);
                if (base + c > plan.cap[tid]) atomicExch(overflow, 1u);
            }
            s_base[tid] = base;
            s_cnt[tid] = 0;
        }
        __syncthreads();
        if (tid < (int)nres) {
            const unsigned long long slot = s_base[b] + r;
            s_dst[tid] = slot < plan.cap[b] ? plan.dst[b] + slot_word(slot) : nullptr;
        }
        __syncthreads();
        for (uint32_t e = warp; e < nres; e += kRtThreads / 32) { // one warp per record: 32 + 3 consecutive words
            const uint32_t *src = blk + kEntriesWord + e * kRecWords;
            uint32_t *dst = s_dst[e];
            if (dst) {
                dst[lane] = src[lane];
                if (lane < 3) dst[32 + lane] = src[32 + lane];
            }
        }
        __syncthreads(); // the stage and the lists are free again
    }
}

int launch_route_scatter(const void *d_in, uint64_t nblocks, const RoutePlan &plan, unsigned long long *d_cursor, uint32_t *d_overflow,
                         cudaStream_t st) {
    if (!nblocks) return 0;
    const size_t smem = sizeof(uint32_t) * kBlockWords * kRtStages + 8 * kRtStages + 128;
    static int per_sm = 0;
    if (first_use_on_device((const void *)route_scatter_kernel))
        DBT_CUDA(cudaFuncSetAttribute(route_scatter_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    if (!per_sm) {
        int occ = 0;
        DBT_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, route_scatter_kernel, kRtThreads, smem));
        per_sm = std::max(occ, 1);
    }
    StageScope sc(ST_GATHER, st);
    const int grid = (int)std::min<uint64_t>(nblocks, (uint64_t)148 * per_sm);
    route_scatter_kernel<<<grid, kRtThreads, smem, st>>>((const uint32_t *)d_in, nblocks, plan, d_cursor, d_overflow);
    count_launch();
    DBT_KERNEL_CHECK();
    return 0;
}

// headers (and the unused tail of the last block) of the images the scatter filled: blockid = block number inside the
// image, nreserved = live rows, valid = 1, dummy = nreserved -- what gather_push_kernel writes
__global__ void __launch_bounds__(256) seg_headers_kernel(const __grid_constant__ SegFillPlan plan) {
    const SegFill sg = plan.seg[blockIdx.y];
    const uint64_t nb = (sg.nrows + kRpb - 1) / kRpb;
    for (uint64_t b = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; b < nb; b += (uint64_t)gridDim.x * blockDim.x) {
        const uint32_t cnt = (uint32_t)min((uint64_t)kRpb, sg.nrows - b * kRpb);
        uint32_t *o = sg.img + b * kBlockWords;
        o[0] = (uint32_t)b;
        o[1] = cnt;
        o[kTrailerWord] = 1;
        o[kTrailerWord + 1] = cnt;
        for (uint32_t w = kEntriesWord + cnt * kRecWords; w < kTrailerWord; ++w) o[w] = 0; // only the last block has a tail
    }
}
int launch_seg_headers(const SegFillPlan &plan, cudaStream_t st) {
    if (!plan.nseg) return 0;
    uint64_t mx = 0;
    for (uint32_t i = 0; i < plan.nseg; ++i) mx = std::max<uint64_t>(mx, (plan.seg[i].nrows + kRpb - 1) / kRpb);
    if (!mx) return 0;
    const dim3 grid((unsigned)std::min<uint64_t>((mx + 255) / 256, 148 * 4), plan.nseg);
    seg_headers_kernel<<<grid, 256, 0, st>>>(plan);
    count_launch();
    DBT_KERNEL_CHECK();
    return 0;
}

__global__ void sample_image_kernel(const uint32_t *__restrict__ img, uint64_t nblocks, uint32_t word, uint32_t nsamp,
                                    uint32_t *__restrict__ keys, uint32_t *__restrict__ ok) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nsamp) return;
    const uint64_t b = (uint64_t)i * nblocks / nsamp;
    const uint32_t *blk = img + b * kBlockWords;
    const uint32_t nres = min(blk[1], kRpb);
    if (!nres) {
        ok[i] = 0;
        keys[i] = 0;
        return;
    }
    const uint32_t e = (i * 37u) % nres;
    keys[i] = blk[kEntriesWord + e * kRecWords + word];
    ok[i] = 1;
}
int launch_sample_image(const void *d_in, uint64_t nblocks, uint32_t word, uint32_t nsamp, uint32_t *d_keys, uint32_t *d_ok,
                        cudaStream_t st) {
    if (!nsamp || !nblocks) return 0;
    sample_image_kernel<<<(nsamp + 255) / 256, 256, 0, st>>>((const uint32_t *)d_in, nblocks, word, nsamp, d_keys, d_ok);
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
