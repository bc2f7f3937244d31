// fft_tables.cu — cached device twiddle tables for fft_smem.cuh.
#include <map>
#include <math.h>
#include <mutex>
#include <utility>

#include "common.cuh"
#include "fft_smem.cuh"

const float2 *fft_twiddles(int N) {
    static std::mutex mu;
    static std::map<std::pair<int, int>, float2 *> cache;
    std::lock_guard<std::mutex> lk(mu);
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return nullptr;
    auto key = std::make_pair(dev, N);
    auto it = cache.find(key);
    if (it != cache.end()) return it->second;
    std::vector<float2> h(N);
    for (int j = 0; j < N; ++j) {
        const double a = -2.0 * 3.14159265358979323846 * (double)j / (double)N;
        h[j] = make_float2((float)cos(a), (float)sin(a));
    }
    float2 *d = nullptr;
    if (cudaMalloc(&d, sizeof(float2) * N) != cudaSuccess) return nullptr;
    if (cudaMemcpy(d, h.data(), sizeof(float2) * N, cudaMemcpyHostToDevice) != cudaSuccess) { cudaFree(d); return nullptr; }
    cache[key] = d;
    return d;
}
