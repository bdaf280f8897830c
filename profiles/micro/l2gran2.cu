// Microbenchmark (round 2): does ANY load flavour fetch less than a 128-byte line from DRAM for a random record read?
// Reads `words` 4-byte words of 50M pseudo-random 140-byte records (7 GB image, far beyond L2) with different instructions.
// Run under `ncu --metrics dram__bytes_read.sum,gpu__time_duration.sum` to see the bytes; plain run prints times.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 l2gran2.cu -o _bin/l2gran2
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

enum { LD_PLAIN, LD_NC, LD_CG, LD_CV, LD_NOALLOC_EF, LD_RELAXED, LD_CPASYNC, LD_BULK, LD_L2_64, LD_L2_256, N_FLAVOURS };
static const char *kNames[] = {"ld.global", "ld.global.nc", "ld.global.cg", "ld.global.cv", "ld.nc.L1::no_allocate.L2::evict_first",
                               "ld.relaxed.gpu", "cp.async.ca 4B", "cp.async.bulk 16B", "ld.global.L2::64B", "ld.global.L2::256B"};

template <int F>
__device__ __forceinline__ uint32_t load(const uint32_t *p, uint32_t *sm_slot, uint64_t *mbar, uint32_t &phase) {
    uint32_t v = 0;
    if (F == LD_PLAIN) asm volatile("ld.global.u32 %0, [%1];" : "=r"(v) : "l"(p));
    if (F == LD_NC) asm volatile("ld.global.nc.u32 %0, [%1];" : "=r"(v) : "l"(p));
    if (F == LD_CG) asm volatile("ld.global.cg.u32 %0, [%1];" : "=r"(v) : "l"(p));
    if (F == LD_CV) asm volatile("ld.global.cv.u32 %0, [%1];" : "=r"(v) : "l"(p));
    if (F == LD_NOALLOC_EF) {
        uint64_t pol;
        asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
        asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.u32 %0, [%1], %2;" : "=r"(v) : "l"(p), "l"(pol));
    }
    if (F == LD_RELAXED) asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p));
    if (F == LD_L2_64) asm volatile("ld.global.L2::64B.u32 %0, [%1];" : "=r"(v) : "l"(p));
    if (F == LD_L2_256) asm volatile("ld.global.L2::256B.u32 %0, [%1];" : "=r"(v) : "l"(p));
    if (F == LD_CPASYNC) {
        const uint32_t s = (uint32_t)__cvta_generic_to_shared(sm_slot);
        asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(s), "l"(p));
        asm volatile("cp.async.wait_all;" ::: "memory");
        v = *(volatile uint32_t *)sm_slot;
    }
    return v;
}

// every thread: one record per iteration, `words` consecutive words of it
template <int F>
__global__ void __launch_bounds__(256) rnd(const uint32_t *__restrict__ in, uint64_t nrec, uint32_t *out, int words) {
    __shared__ uint32_t slots[256 * 4];
    __shared__ uint64_t mbar;
    uint32_t phase = 0;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    uint32_t acc = 0;
    if (F == LD_BULK) {
        // one elected lane per warp issues 32 bulk copies of 16 bytes (one per lane's record), all lanes then read shared memory
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
        __shared__ uint64_t wbar[8];
        __shared__ __align__(16) uint32_t wslots[8][32][4];
        if (lane == 0) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"((uint32_t)__cvta_generic_to_shared(&wbar[warp])));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        __syncwarp();
        for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nrec; i += stride) {
            const uint64_t src = (i * 2654435761ull + 12345) % nrec;
            const uint32_t *p = in + src * 35;
            const uint64_t pa = (uint64_t)p & ~15ull; // the aligned 16 bytes holding the record's first word
            const uint32_t bar = (uint32_t)__cvta_generic_to_shared(&wbar[warp]);
            if (lane == 0) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(32 * 16));
            __syncwarp();
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], 16, [%2];" ::"r"(
                             (uint32_t)__cvta_generic_to_shared(&wslots[warp][lane][0])),
                         "l"(pa), "r"(bar)
                         : "memory");
            uint32_t ok = 0;
            while (!ok)
                asm volatile("{.reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p;}"
                             : "=r"(ok)
                             : "r"(bar), "r"(phase)
                             : "memory");
            phase ^= 1;
            acc += wslots[warp][lane][((uint64_t)p >> 2) & 3];
            __syncwarp();
        }
    } else {
        for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nrec; i += stride) {
            const uint64_t src = (i * 2654435761ull + 12345) % nrec;
            const uint32_t *p = in + src * 35;
            for (int w = 0; w < words; ++w) acc += load<F>(p + w, &slots[threadIdx.x * 4 + (w & 3)], &mbar, phase);
        }
    }
    if (acc == 0x12345678) out[0] = acc;
}

template <int F>
static void run(const uint32_t *in, uint64_t nrec, uint32_t *out, int words) {
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    float ms = 0;
    for (int it = 0; it < 3; ++it) {
        cudaEventRecord(a);
        rnd<F><<<148 * 8, 256>>>(in, nrec, out, words);
        cudaEventRecord(b);
        cudaEventSynchronize(b);
        cudaEventElapsedTime(&ms, a, b);
    }
    cudaError_t e = cudaGetLastError();
    printf("%-42s words=%2d  %.3f ms  useful %.1f GB/s  (128-B lines: %.1f GB/s, 32-B sectors: %.1f GB/s)%s\n", kNames[F], words, ms,
           nrec * words * 4 / ms / 1e6, nrec * (words <= 2 ? 1.06 : 2.09) * 128 / ms / 1e6,
           nrec * (words <= 2 ? 1.19 : 5.3) * 32 / ms / 1e6, e ? cudaGetErrorString(e) : "");
}

int main(int argc, char **argv) {
    const uint64_t nrec = 50000000ull;
    uint32_t *in, *out;
    cudaMalloc(&in, nrec * 140);
    cudaMalloc(&out, 4);
    cudaMemset(in, 1, nrec * 140);
    for (int words : {2, 35}) {
        run<LD_PLAIN>(in, nrec, out, words);
        run<LD_NC>(in, nrec, out, words);
        run<LD_CG>(in, nrec, out, words);
        run<LD_CV>(in, nrec, out, words);
        run<LD_NOALLOC_EF>(in, nrec, out, words);
        run<LD_RELAXED>(in, nrec, out, words);
        run<LD_L2_64>(in, nrec, out, words);
        run<LD_L2_256>(in, nrec, out, words);
        if (words == 2) {
            run<LD_CPASYNC>(in, nrec, out, words);
            run<LD_BULK>(in, nrec, out, 1);
        }
    }
    return 0;
}
