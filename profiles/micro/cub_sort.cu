// Library reference point for the pair sort: cub::DeviceRadixSort::SortPairs (CUDA 12.9's CUB, its own sm_100 tuning)
// on the same inputs as bench.py's "pair_sort_1B_u32" and the 100M-pair case.  Not part of the product.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a profiles/micro/cub_sort.cu -o gpurun_out/cub_sort && gpurun_out/cub_sort
#include <cub/device/device_radix_sort.cuh>
#include <cstdio>
#include <cstdint>
#include <vector>

__global__ void fill(uint32_t *k, uint32_t *v, uint64_t n, uint64_t seed) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (uint64_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) {
        uint64_t x = (i + seed) * 0x9E3779B97F4A7C15ull;
        x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
        x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
        k[i] = (uint32_t)(x >> 32);
        v[i] = (uint32_t)i;
    }
}

int main() {
    for (uint64_t n : {100000000ull, 1000000000ull}) {
        uint32_t *k1, *k2, *v1, *v2;
        cudaMalloc(&k1, 4 * n); cudaMalloc(&k2, 4 * n); cudaMalloc(&v1, 4 * n); cudaMalloc(&v2, 4 * n);
        size_t tb = 0;
        cub::DoubleBuffer<uint32_t> dk(k1, k2), dv(v1, v2);
        cub::DeviceRadixSort::SortPairs(nullptr, tb, dk, dv, (int64_t)n, 0, 32);
        void *tmp; cudaMalloc(&tmp, tb);
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        float best = 1e9f;
        for (int it = 0; it < 4; ++it) {
            fill<<<148 * 8, 256>>>(k1, v1, n, 1);
            cub::DoubleBuffer<uint32_t> a(k1, k2), b(v1, v2);
            cudaDeviceSynchronize();
            cudaEventRecord(e0);
            cub::DeviceRadixSort::SortPairs(tmp, tb, a, b, (int64_t)n, 0, 32);
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            if (it && ms < best) best = ms;
        }
        printf("{\"cub_sort_pairs_u32\": {\"n\": %llu, \"ms\": %.3f, \"pairs_per_s\": %.4g, \"err\": \"%s\"}}\n",
               (unsigned long long)n, best, n / (best * 1e-3), cudaGetErrorString(cudaGetLastError()));
        cudaFree(k1); cudaFree(k2); cudaFree(v1); cudaFree(v2); cudaFree(tmp);
    }
    return 0;
}
