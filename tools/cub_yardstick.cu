// Yardstick only (never linked into libmss_b200.so): cub::DeviceRadixSort on the same box, same sizes,
// so the repo's own onesweep sort has an external anchor.   nvcc -O3 -arch=sm_100a tools/cub_yardstick.cu -o tools/cub_yardstick
#include <cub/cub.cuh>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <algorithm>
#include <random>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

__global__ void fill(uint32_t *k, uint8_t *v, size_t n, uint32_t seed) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        uint32_t x = (uint32_t)i * 2654435761u + seed;
        x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
        // float-like key distribution: a normal-ish score mapped through the order-preserving transform
        float f = ((x & 0xffff) + ((x >> 16) & 0xffff)) * (1.0f / 65536.0f) - 1.0f;   // triangular in (-1, 1)
        f *= 8.0f;
        uint32_t u = __float_as_uint(f);
        k[i] = ~((u >> 31) ? ~u : (u | 0x80000000u));
        v[i] = (x % 20u) == 0;
    }
}

int main(int argc, char **argv) {
    std::vector<size_t> sizes;
    for (int i = 1; i < argc; i++) sizes.push_back((size_t)atoll(argv[i]));
    if (sizes.empty()) sizes = {1u << 21, 1u << 23, 1u << 25, 1u << 27};
    for (size_t n : sizes) {
        uint32_t *k0, *k1; uint8_t *v0, *v1;
        CK(cudaMalloc(&k0, n * 4)); CK(cudaMalloc(&k1, n * 4)); CK(cudaMalloc(&v0, n)); CK(cudaMalloc(&v1, n));
        size_t tb_pairs = 0, tb_keys = 0;
        cub::DoubleBuffer<uint32_t> dk(k0, k1); cub::DoubleBuffer<uint8_t> dv(v0, v1);
        CK(cub::DeviceRadixSort::SortPairs(nullptr, tb_pairs, dk, dv, (int64_t)n));
        CK(cub::DeviceRadixSort::SortKeys(nullptr, tb_keys, dk, (int64_t)n));
        void *tmp; CK(cudaMalloc(&tmp, std::max(tb_pairs, tb_keys)));
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        float best_pairs = 1e30f, best_keys = 1e30f;
        for (int rep = 0; rep < 6; rep++) {
            fill<<<1184, 256>>>(k0, v0, n, 17u + rep);
            cub::DoubleBuffer<uint32_t> a(k0, k1); cub::DoubleBuffer<uint8_t> b(v0, v1);
            CK(cudaDeviceSynchronize());
            cudaEventRecord(e0);
            CK(cub::DeviceRadixSort::SortPairs(tmp, tb_pairs, a, b, (int64_t)n));
            cudaEventRecord(e1); CK(cudaDeviceSynchronize());
            float ms; cudaEventElapsedTime(&ms, e0, e1); if (rep) best_pairs = std::min(best_pairs, ms);
            fill<<<1184, 256>>>(k0, v0, n, 99u + rep);
            cub::DoubleBuffer<uint32_t> c(k0, k1);
            CK(cudaDeviceSynchronize());
            cudaEventRecord(e0);
            CK(cub::DeviceRadixSort::SortKeys(tmp, tb_keys, c, (int64_t)n));
            cudaEventRecord(e1); CK(cudaDeviceSynchronize());
            cudaEventElapsedTime(&ms, e0, e1); if (rep) best_keys = std::min(best_keys, ms);
        }
        printf("{\"n\": %zu, \"cub_sort_pairs_u32_u8_ms\": %.4f, \"cub_pairs_gpairs_s\": %.2f, \"cub_sort_keys_u32_ms\": %.4f, \"cub_keys_gkeys_s\": %.2f}\n",
               n, best_pairs, n / best_pairs / 1e6, best_keys, n / best_keys / 1e6);
        cudaFree(k0); cudaFree(k1); cudaFree(v0); cudaFree(v1); cudaFree(tmp);
    }
    return 0;
}
