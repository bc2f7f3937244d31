// FFMA2 (fma.rn.f32x2) microbenchmark for the K1 inner loop (scratch)
#include <cstdio>
#include <cuda_runtime.h>
template <int M>
__global__ void __launch_bounds__(384, 1) k(float *out, const float2 *taps_g, const float2 *x_g, int iters) {
    float2 tap[4][11];
    for (int r = 0; r < 4; ++r) for (int j = 0; j < 11; ++j) tap[r][j] = taps_g[(r * 11 + j) * 32 + (threadIdx.x & 31)];
    float2 A[M][4], B[M][4];
    for (int s = 0; s < M; ++s) for (int r = 0; r < 4; ++r) { A[s][r] = make_float2(0.f, 0.f); B[s][r] = make_float2(0.f, 0.f); }
    extern __shared__ float2 xs[];
    for (int i = threadIdx.x; i < 4096; i += blockDim.x) xs[i] = x_g[i];
    __syncthreads();
    const int lane = threadIdx.x & 31;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int s = 0; s < M; ++s) {
#pragma unroll
            for (int j = 0; j < 11; ++j) {
                const float2 xv = xs[(it * 7 + s * 500 + 400 - lane - 32 * j) & 4095];
                const float2 xrr = make_float2(xv.x, xv.x), xii = make_float2(xv.y, xv.y);
#pragma unroll
                for (int r = 0; r < 4; ++r) A[s][r] = __ffma2_rn(tap[r][j], xrr, A[s][r]);
#pragma unroll
                for (int r = 0; r < 4; ++r) B[s][r] = __ffma2_rn(tap[r][j], xii, B[s][r]);
            }
        }
    }
    float sum = 0.f;
    for (int s = 0; s < M; ++s) for (int r = 0; r < 4; ++r) sum += A[s][r].x - B[s][r].y + B[s][r].x + A[s][r].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = sum;
}
int main() {
    float *out; float2 *taps, *x;
    cudaMalloc(&out, 148 * 384 * 4); cudaMalloc(&taps, 44 * 32 * 8); cudaMalloc(&x, 4096 * 8);
    cudaMemset(taps, 0, 44 * 32 * 8); cudaMemset(x, 0, 4096 * 8);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 20000;
    for (int mode = 0; mode < 2; ++mode) {
        for (int rep = 0; rep < 2; ++rep) {
            cudaEventRecord(e0);
            if (mode == 0) k<2><<<148, 384, 4096 * 8>>>(out, taps, x, iters);
            else k<4><<<148, 384, 4096 * 8>>>(out, taps, x, iters);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            const int M = mode == 0 ? 2 : 4;
            double fma_equiv = 148.0 * 12 * (176.0 * M) * iters;      // warp-level scalar-FFMA equivalents
            printf("FFMA2 M=%d: %.3f ms -> %.2f warp-FFMA-equiv/clk/SM at 1.965 GHz (scalar peak 4), %.1f TFLOP/s\n", M, ms,
                   fma_equiv / (ms * 1e-3) / 148 / 1.965e9, fma_equiv * 64 / (ms * 1e-3) / 1e12);
        }
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
