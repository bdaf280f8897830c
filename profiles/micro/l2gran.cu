// Microbenchmark: does cudaLimitMaxL2FetchGranularity change the DRAM cost of 140-byte random reads?
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 l2gran.cu -o l2gran ; run: ./l2gran <32|64|128>
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
__global__ void rnd(const unsigned *in, unsigned long long nrec, unsigned *out, int words) {
    unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    unsigned acc = 0;
    for (; i < nrec * words; i += stride) {
        unsigned long long r = i / words, w = i % words;
        unsigned long long src = (r * 2654435761ull + 12345) % nrec; // pseudo-random record
        acc += in[src * 35 + w];
    }
    if (acc == 0x12345678) out[0] = acc;
}
int main(int argc, char **argv) {
    size_t gran = argc > 1 ? atoi(argv[1]) : 0;
    if (gran) printf("set limit -> %d\n", (int)cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, gran));
    size_t g2 = 0; cudaDeviceGetLimit(&g2, cudaLimitMaxL2FetchGranularity); printf("limit now %zu\n", g2);
    unsigned long long nrec = 50000000ull; unsigned *in, *out;
    cudaMalloc(&in, nrec * 140); cudaMalloc(&out, 4); cudaMemset(in, 1, nrec * 140);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    for (int words : {2, 35}) {
        for (int it = 0; it < 3; ++it) {
            cudaEventRecord(a); rnd<<<148 * 16, 256>>>(in, nrec, out, words); cudaEventRecord(b); cudaEventSynchronize(b);
            float ms; cudaEventElapsedTime(&ms, a, b);
            if (it == 2) printf("words=%d  %.3f ms  useful %.1f GB/s\n", words, ms, nrec * words * 4 / ms / 1e6);
        }
    }
    return 0;
}
