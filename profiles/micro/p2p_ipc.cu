// p2p_ipc.cu -- copy-engine NVLink bandwidth between TWO PROCESSES (one per GPU, destination mapped with CUDA IPC), both
// directions busy, the copy split over 1/2/4/8 streams.  The distributed step (csrc/dist.cu under torchrun) moves its
// records this way; the one-process benchmark (p2p_bidir.cu) reaches 779 GB/s per direction.
// build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a p2p_ipc.cu -o _bin/p2p_ipc
#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cuda_runtime.h>
#include <sys/mman.h>
#include <sys/wait.h>
#include <unistd.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

struct Shared {
    std::atomic<int> arrive[256];
    cudaIpcMemHandle_t h[2];
    float ms[2];
};
static void barrier(Shared *s, int &phase) {
    s->arrive[phase].fetch_add(1);
    while (s->arrive[phase].load() < 2) usleep(50);
    ++phase;
}
__global__ void hbm_load(const uint4 *src, uint4 *dst, size_t nvec, int reps) { // local copy traffic on all SMs
    size_t stride = (size_t)gridDim.x * blockDim.x;
    for (int r = 0; r < reps; ++r)
        for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += stride) dst[i] = src[i];
}

int main() {
    const size_t bytes = 4ull << 30;
    Shared *s = (Shared *)mmap(nullptr, sizeof(Shared), PROT_READ | PROT_WRITE, MAP_SHARED | MAP_ANONYMOUS, -1, 0);
    memset(s, 0, sizeof(Shared));
    pid_t pid = fork();
    const int rank = pid == 0 ? 1 : 0;
    int phase = 0;
    int n = 0;
    CK(cudaGetDeviceCount(&n));
    if (n < 2) { if (rank == 0) printf("needs 2 GPUs\n"); return 0; }
    CK(cudaSetDevice(rank));
    char *a, *b, *c, *peer;
    CK(cudaMalloc(&a, bytes)); CK(cudaMalloc(&b, bytes)); CK(cudaMalloc(&c, bytes));
    CK(cudaMemset(a, rank + 1, bytes));
    CK(cudaIpcGetMemHandle(&s->h[rank], b));
    barrier(s, phase);
    CK(cudaIpcOpenMemHandle((void **)&peer, s->h[1 - rank], cudaIpcMemLazyEnablePeerAccess));
    cudaStream_t st[8], load;
    for (int i = 0; i < 8; ++i) CK(cudaStreamCreateWithFlags(&st[i], cudaStreamNonBlocking));
    CK(cudaStreamCreateWithFlags(&load, cudaStreamNonBlocking));
    cudaEvent_t e0, e1[8];
    CK(cudaEventCreate(&e0));
    for (int i = 0; i < 8; ++i) CK(cudaEventCreate(&e1[i]));
    for (int busy = 0; busy < 2; ++busy)
        for (int both = 0; both < 2; ++both)
            for (int k = 1; k <= 8; k *= 2) {
                float best = 1e9f;
                for (int it = 0; it < 3; ++it) {
                    CK(cudaDeviceSynchronize());
                    barrier(s, phase);
                    if (busy) hbm_load<<<148 * 4, 256, 0, load>>>((const uint4 *)a, (uint4 *)c, bytes / 16, 6); // ~25 ms of local HBM traffic
                    if (both || rank == 0) {
                        CK(cudaEventRecord(e0, st[0]));
                        for (int i = 1; i < k; ++i) CK(cudaStreamWaitEvent(st[i], e0, 0));
                        const size_t part = bytes / k;
                        for (int i = 0; i < k; ++i) {
                            CK(cudaMemcpyAsync(peer + i * part, a + i * part, part, cudaMemcpyDeviceToDevice, st[i]));
                            CK(cudaEventRecord(e1[i], st[i]));
                        }
                        float ms = 0;
                        for (int i = 0; i < k; ++i) { CK(cudaEventSynchronize(e1[i])); float m; CK(cudaEventElapsedTime(&m, e0, e1[i])); if (m > ms) ms = m; }
                        if (it && ms < best) best = ms;
                    }
                    CK(cudaDeviceSynchronize());
                }
                s->ms[rank] = best;
                barrier(s, phase);
                if (rank == 0) {
                    printf("{\"pattern\": \"IPC push, copy engines\", \"local_hbm_load\": %d, \"directions\": %d, \"streams\": %d, \"gbs_0to1\": %.1f, \"gbs_1to0\": %.1f}\n",
                           busy, both ? 2 : 1, k, bytes / s->ms[0] / 1e6, both ? bytes / s->ms[1] / 1e6 : 0.0);
                    fflush(stdout);
                }
                barrier(s, phase);
            }
    CK(cudaIpcCloseMemHandle(peer));
    if (rank == 0) { int st_; waitpid(pid, &st_, 0); }
    return 0;
}
