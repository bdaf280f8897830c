// p2p_scatter.cu -- NVLink store efficiency by write pattern, GPU0 -> GPU1 (single process, peer access).
// Patterns: (a) whole 14016-byte blocks, 16-byte vectors (what gather_push_kernel does);
//           (b) 140-byte records (one warp per record, 4-byte lanes) to every k-th record slot in ascending order
//               (k = 1, 2, 8: what a sender would write if rows went straight to their final positions among P ranks);
//           (c) 140-byte records to random slots;
//           (d) remote READS of 140-byte records from every k-th slot (the pull direction), for comparison.
// build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a p2p_scatter.cu -o _bin/p2p_scatter
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include <vector>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)
constexpr unsigned kBlockWords = 3504, kRecWords = 35, kRpb = 100;
__host__ __device__ inline unsigned long long slot_word(unsigned long long slot) {
    unsigned long long b = slot / kRpb;
    return b * kBlockWords + 2 + (slot - b * kRpb) * kRecWords;
}
__global__ void block_copy(const uint4 *src, uint4 *dst, size_t nvec) {
    size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += stride) dst[i] = src[i];
}
// warp w copies record w of the local image to slot pos(w) of the remote image (or the other way round)
__global__ void rec_scatter(const unsigned *src, unsigned *dst, unsigned long long nrec, unsigned k, const unsigned *perm, int remote_is_src) {
    const unsigned lane = threadIdx.x & 31;
    unsigned long long w = ((unsigned long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const unsigned long long nw = ((unsigned long long)gridDim.x * blockDim.x) >> 5;
    for (; w < nrec; w += nw) {
        const unsigned long long far = perm ? perm[w] : w * k;
        const unsigned long long a = slot_word(remote_is_src ? far : w), b = slot_word(remote_is_src ? w : far);
        dst[b + lane] = src[a + lane];
        if (lane < 3) dst[b + 32 + lane] = src[a + 32 + lane];
    }
}
int main(int argc, char **argv) {
    const unsigned long long nrec = argc > 1 ? strtoull(argv[1], 0, 10) : 20000000ull; // records moved per test
    int n = 0;
    CK(cudaGetDeviceCount(&n));
    if (n < 2) { printf("needs 2 GPUs\n"); return 0; }
    const unsigned long long far_slots = nrec * 8 + 100, far_bytes = (far_slots / kRpb + 1) * kBlockWords * 4ull, near_bytes = (nrec / kRpb + 1) * kBlockWords * 4ull;
    unsigned *d0, *d1, *perm;
    CK(cudaSetDevice(1)); CK(cudaMalloc(&d1, far_bytes)); CK(cudaMemset(d1, 1, far_bytes));
    CK(cudaSetDevice(0)); CK(cudaDeviceEnablePeerAccess(1, 0)); CK(cudaMalloc(&d0, near_bytes)); CK(cudaMemset(d0, 2, near_bytes));
    std::vector<unsigned> h(nrec);
    unsigned long long x = 88172645463325252ull;
    for (unsigned long long i = 0; i < nrec; ++i) h[i] = (unsigned)i;
    for (unsigned long long i = nrec - 1; i > 0; --i) { x ^= x << 13; x ^= x >> 7; x ^= x << 17; unsigned long long j = x % (i + 1); unsigned t = h[i]; h[i] = h[j]; h[j] = t; }
    CK(cudaMalloc(&perm, nrec * 4)); CK(cudaMemcpy(perm, h.data(), nrec * 4, cudaMemcpyHostToDevice));
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    auto timeit = [&](const char *name, double bytes, auto launch) {
        launch(); CK(cudaDeviceSynchronize());
        float best = 1e9f;
        for (int it = 0; it < 3; ++it) { CK(cudaEventRecord(e0)); launch(); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms; }
        printf("{\"pattern\": \"%s\", \"ms\": %.3f, \"gbs\": %.1f}\n", name, best, bytes / best / 1e6); fflush(stdout);
    };
    const double rec_bytes = (double)nrec * 140.0;
    const size_t nvec = near_bytes / 16;
    for (int ctas : {148 * 2, 148 * 8}) {
        printf("-- %d CTAs x 256 threads\n", ctas);
        timeit("blocks, 16-byte vectors, local -> remote", (double)near_bytes, [&] { block_copy<<<ctas, 256>>>((const uint4 *)d0, (uint4 *)d1, nvec); });
        timeit("memcpy peer (copy engine)", (double)near_bytes, [&] { cudaMemcpyPeerAsync(d1, 1, d0, 0, near_bytes); });
        for (unsigned k : {1u, 2u, 8u}) {
            char nm[128]; snprintf(nm, sizeof nm, "records -> every %u-th remote slot, ascending (write)", k);
            timeit(nm, rec_bytes, [&] { rec_scatter<<<ctas, 256>>>(d0, d1, nrec, k, nullptr, 0); });
        }
        timeit("records -> random remote slots (write)", rec_bytes, [&] { rec_scatter<<<ctas, 256>>>(d0, d1, nrec, 1, perm, 0); });
        for (unsigned k : {1u, 2u, 8u}) {
            char nm[128]; snprintf(nm, sizeof nm, "records <- every %u-th remote slot, ascending (read)", k);
            timeit(nm, rec_bytes, [&] { rec_scatter<<<ctas, 256>>>(d1, d0, nrec, k, nullptr, 1); });
        }
        timeit("records <- random remote slots (read)", rec_bytes, [&] { rec_scatter<<<ctas, 256>>>(d1, d0, nrec, 1, perm, 1); });
    }
    return 0;
}
