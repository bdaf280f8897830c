// p2p_bidir.cu -- copy-engine and SM-store NVLink bandwidth with BOTH directions busy (GPU0 <-> GPU1), one process.
// build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a p2p_bidir.cu -o _bin/p2p_bidir
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)
__global__ void block_copy(const uint4 *src, uint4 *dst, size_t nvec) {
    size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += stride) dst[i] = src[i];
}
int main() {
    const size_t bytes = 4ull << 30;
    int n = 0; CK(cudaGetDeviceCount(&n)); if (n < 2) { printf("needs 2 GPUs\n"); return 0; }
    char *a[2], *b[2]; cudaStream_t st[2]; cudaEvent_t e0[2], e1[2];
    for (int d = 0; d < 2; ++d) {
        CK(cudaSetDevice(d)); CK(cudaDeviceEnablePeerAccess(1 - d, 0));
        CK(cudaMalloc(&a[d], bytes)); CK(cudaMalloc(&b[d], bytes)); CK(cudaMemset(a[d], d, bytes));
        CK(cudaStreamCreateWithFlags(&st[d], cudaStreamNonBlocking)); CK(cudaEventCreate(&e0[d])); CK(cudaEventCreate(&e1[d]));
    }
    auto run = [&](const char *name, int mode, int both) {
        float best[2] = {1e9f, 1e9f};
        for (int it = 0; it < 4; ++it) {
            for (int d = 0; d < 2; ++d) { CK(cudaSetDevice(d)); CK(cudaDeviceSynchronize()); }
            for (int d = 0; d < (both ? 2 : 1); ++d) {
                CK(cudaSetDevice(d)); CK(cudaEventRecord(e0[d], st[d]));
                if (mode == 0) CK(cudaMemcpyPeerAsync(b[1 - d], 1 - d, a[d], d, bytes, st[d]));
                else if (mode == 1) CK(cudaMemcpyAsync(b[1 - d], a[d], bytes, cudaMemcpyDeviceToDevice, st[d]));
                else block_copy<<<148 * 8, 256, 0, st[d]>>>((const uint4 *)a[d], (uint4 *)b[1 - d], bytes / 16);
                CK(cudaEventRecord(e1[d], st[d]));
            }
            for (int d = 0; d < (both ? 2 : 1); ++d) { CK(cudaSetDevice(d)); CK(cudaEventSynchronize(e1[d])); float ms; CK(cudaEventElapsedTime(&ms, e0[d], e1[d])); if (it && ms < best[d]) best[d] = ms; }
        }
        printf("{\"pattern\": \"%s\", \"directions\": %d, \"gbs_0to1\": %.1f, \"gbs_1to0\": %.1f}\n", name, both ? 2 : 1, bytes / best[0] / 1e6, both ? bytes / best[1] / 1e6 : 0.0);
        fflush(stdout);
    };
    run("cudaMemcpyPeerAsync", 0, 0); run("cudaMemcpyPeerAsync", 0, 1);
    run("cudaMemcpyAsync D2D (UVA peer pointer)", 1, 0); run("cudaMemcpyAsync D2D (UVA peer pointer)", 1, 1);
    run("SM 16-byte stores to peer", 2, 0); run("SM 16-byte stores to peer", 2, 1);
    return 0;
}
