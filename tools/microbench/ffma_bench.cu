// FFMA issue-rate microbenchmark for the K1 operand pattern (scratch, not part of the library)
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE>
__global__ void __launch_bounds__(384, 1) k(float *out, const float2 *taps_g, const float2 *x_g, int iters) {
    float2 tap[4][11];
    for (int r = 0; r < 4; ++r) for (int j = 0; j < 11; ++j) tap[r][j] = taps_g[(r * 11 + j) * 32 + (threadIdx.x & 31)];
    float acc[32];
    for (int v = 0; v < 32; ++v) acc[v] = 0.f;
    extern __shared__ float2 xs[];
    for (int i = threadIdx.x; i < 4096; i += blockDim.x) xs[i] = x_g[i];
    __syncthreads();
    const int lane = threadIdx.x & 31;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int s = 0; s < 4; ++s) {
            float2 xv[11];
#pragma unroll
            for (int j = 0; j < 11; ++j) xv[j] = xs[(it * 7 + s * 500 + 400 - lane - 32 * j) & 4095];
#pragma unroll
            for (int j = 0; j < 11; ++j) {
#pragma unroll
                for (int r = 0; r < 4; ++r) {
                    const int slot = s * 8 + r * 2;
                    if (MODE == 0) {           // K1 pattern: complex MAC
                        acc[slot] = fmaf(tap[r][j].x, xv[j].x, acc[slot]);
                        acc[slot + 1] = fmaf(tap[r][j].x, xv[j].y, acc[slot + 1]);
                    } else {                   // same count, but independent of x (pure register FMA, max reuse)
                        acc[slot] = fmaf(tap[r][j].x, tap[r][j].y, acc[slot]);
                        acc[slot + 1] = fmaf(tap[r][j].y, tap[r][j].x, acc[slot + 1]);
                    }
                }
#pragma unroll
                for (int r = 0; r < 4; ++r) {
                    const int slot = s * 8 + r * 2;
                    if (MODE == 0) {
                        acc[slot] = fmaf(-tap[r][j].y, xv[j].y, acc[slot]);
                        acc[slot + 1] = fmaf(tap[r][j].y, xv[j].x, acc[slot + 1]);
                    } else {
                        acc[slot] = fmaf(tap[r][j].y, tap[r][j].y, acc[slot]);
                        acc[slot + 1] = fmaf(tap[r][j].x, tap[r][j].x, acc[slot + 1]);
                    }
                }
            }
        }
    }
    float sum = 0.f;
    for (int v = 0; v < 32; ++v) sum += acc[v];
    out[blockIdx.x * blockDim.x + threadIdx.x] = sum;
}

template <int MODE>
__global__ void __launch_bounds__(384, 1) k2(float *out, const float2 *taps_g, const float2 *x_g, int iters) {
    float2 tap[4][11];
    for (int r = 0; r < 4; ++r) for (int j = 0; j < 11; ++j) tap[r][j] = taps_g[(r * 11 + j) * 32 + (threadIdx.x & 31)];
    float acc[32];
    for (int v = 0; v < 32; ++v) acc[v] = 0.f;
    extern __shared__ float2 xs[];
    for (int i = threadIdx.x; i < 4096; i += blockDim.x) xs[i] = x_g[i];
    __syncthreads();
    const int lane = threadIdx.x & 31;
    for (int it = 0; it < iters; ++it) {
        if (MODE == 2) {                 // x-stationary: each x component feeds 8 consecutive FMAs
#pragma unroll
            for (int s = 0; s < 4; ++s) {
                float2 xv[11];
#pragma unroll
                for (int j = 0; j < 11; ++j) xv[j] = xs[(it * 7 + s * 500 + 400 - lane - 32 * j) & 4095];
#pragma unroll
                for (int j = 0; j < 11; ++j) {
#pragma unroll
                    for (int r = 0; r < 4; ++r) acc[s * 8 + r * 2] = fmaf(tap[r][j].x, xv[j].x, acc[s * 8 + r * 2]);
#pragma unroll
                    for (int r = 0; r < 4; ++r) acc[s * 8 + r * 2 + 1] = fmaf(tap[r][j].y, xv[j].x, acc[s * 8 + r * 2 + 1]);
#pragma unroll
                    for (int r = 0; r < 4; ++r) acc[s * 8 + r * 2 + 1] = fmaf(tap[r][j].x, xv[j].y, acc[s * 8 + r * 2 + 1]);
#pragma unroll
                    for (int r = 0; r < 4; ++r) acc[s * 8 + r * 2] = fmaf(-tap[r][j].y, xv[j].y, acc[s * 8 + r * 2]);
                }
            }
        } else {                         // tap-stationary across the 4 outputs s: each tap component feeds 8 consecutive FMAs
#pragma unroll
            for (int j0 = 0; j0 < 11; j0 += 3) {
                float2 xv[4][3];
#pragma unroll
                for (int s = 0; s < 4; ++s)
#pragma unroll
                    for (int jj = 0; jj < 3; ++jj)
                        if (j0 + jj < 11) xv[s][jj] = xs[(it * 7 + s * 500 + 400 - lane - 32 * (j0 + jj)) & 4095];
#pragma unroll
                for (int jj = 0; jj < 3; ++jj) {
                    if (j0 + jj >= 11) continue;
#pragma unroll
                    for (int r = 0; r < 4; ++r) {
#pragma unroll
                        for (int s = 0; s < 4; ++s) {
                            acc[s * 8 + r * 2] = fmaf(tap[r][j0 + jj].x, xv[s][jj].x, acc[s * 8 + r * 2]);
                            acc[s * 8 + r * 2 + 1] = fmaf(tap[r][j0 + jj].x, xv[s][jj].y, acc[s * 8 + r * 2 + 1]);
                        }
#pragma unroll
                        for (int s = 0; s < 4; ++s) {
                            acc[s * 8 + r * 2] = fmaf(-tap[r][j0 + jj].y, xv[s][jj].y, acc[s * 8 + r * 2]);
                            acc[s * 8 + r * 2 + 1] = fmaf(tap[r][j0 + jj].y, xv[s][jj].x, acc[s * 8 + r * 2 + 1]);
                        }
                    }
                }
            }
        }
    }
    float sum = 0.f;
    for (int v = 0; v < 32; ++v) sum += acc[v];
    out[blockIdx.x * blockDim.x + threadIdx.x] = sum;
}

template <int MODE>
__global__ void __launch_bounds__(384, 1) k3(float *out, const float *taps_g, const float2 *x_g, int iters) {
    float tre[4][11], tim[4][11];
    for (int r = 0; r < 4; ++r) for (int j = 0; j < 11; ++j) {
        tre[r][j] = taps_g[(r * 11 + j) * 64 + (threadIdx.x & 31)];
        tim[r][j] = taps_g[(r * 11 + j) * 64 + 32 + (threadIdx.x & 31)];
    }
    float are[16], aim[16];
    for (int v = 0; v < 16; ++v) { are[v] = 0.f; aim[v] = 0.f; }
    extern __shared__ float2 xs[];
    for (int i = threadIdx.x; i < 4096; i += blockDim.x) xs[i] = x_g[i];
    __syncthreads();
    const int lane = threadIdx.x & 31;
    for (int it = 0; it < iters; ++it) {
        if (MODE == 4) {
#pragma unroll
            for (int s = 0; s < 4; ++s) {
                float2 xv[11];
#pragma unroll
                for (int j = 0; j < 11; ++j) xv[j] = xs[(it * 7 + s * 500 + 400 - lane - 32 * j) & 4095];
#pragma unroll
                for (int j = 0; j < 11; ++j) {
#pragma unroll
                    for (int r = 0; r < 4; ++r) are[s * 4 + r] = fmaf(tre[r][j], xv[j].x, are[s * 4 + r]);
#pragma unroll
                    for (int r = 0; r < 4; ++r) aim[s * 4 + r] = fmaf(tim[r][j], xv[j].x, aim[s * 4 + r]);
#pragma unroll
                    for (int r = 0; r < 4; ++r) aim[s * 4 + r] = fmaf(tre[r][j], xv[j].y, aim[s * 4 + r]);
#pragma unroll
                    for (int r = 0; r < 4; ++r) are[s * 4 + r] = fmaf(-tim[r][j], xv[j].y, are[s * 4 + r]);
                }
            }
        } else {                           // two outputs interleaved: 16 FMAs per x pair
#pragma unroll
            for (int s = 0; s < 4; s += 2) {
                float2 xa[11], xb[11];
#pragma unroll
                for (int j = 0; j < 11; ++j) {
                    xa[j] = xs[(it * 7 + s * 500 + 400 - lane - 32 * j) & 4095];
                    xb[j] = xs[(it * 7 + (s + 1) * 500 + 400 - lane - 32 * j) & 4095];
                }
#pragma unroll
                for (int j = 0; j < 11; ++j) {
#pragma unroll
                    for (int r = 0; r < 4; ++r) are[s * 4 + r] = fmaf(tre[r][j], xa[j].x, are[s * 4 + r]);
#pragma unroll
                    for (int r = 0; r < 4; ++r) are[s * 4 + 4 + r] = fmaf(tre[r][j], xb[j].x, are[s * 4 + 4 + r]);
#pragma unroll
                    for (int r = 0; r < 4; ++r) aim[s * 4 + r] = fmaf(tim[r][j], xa[j].x, aim[s * 4 + r]);
#pragma unroll
                    for (int r = 0; r < 4; ++r) aim[s * 4 + 4 + r] = fmaf(tim[r][j], xb[j].x, aim[s * 4 + 4 + r]);
#pragma unroll
                    for (int r = 0; r < 4; ++r) aim[s * 4 + r] = fmaf(tre[r][j], xa[j].y, aim[s * 4 + r]);
#pragma unroll
                    for (int r = 0; r < 4; ++r) aim[s * 4 + 4 + r] = fmaf(tre[r][j], xb[j].y, aim[s * 4 + 4 + r]);
#pragma unroll
                    for (int r = 0; r < 4; ++r) are[s * 4 + r] = fmaf(-tim[r][j], xa[j].y, are[s * 4 + r]);
#pragma unroll
                    for (int r = 0; r < 4; ++r) are[s * 4 + 4 + r] = fmaf(-tim[r][j], xb[j].y, are[s * 4 + 4 + r]);
                }
            }
        }
    }
    float sum = 0.f;
    for (int v = 0; v < 16; ++v) sum += are[v] + aim[v];
    out[blockIdx.x * blockDim.x + threadIdx.x] = sum;
}
int main() {
    float *out; float2 *taps, *x;
    cudaMalloc(&out, 148 * 384 * 4); cudaMalloc(&taps, 44 * 32 * 8); cudaMalloc(&x, 4096 * 8);
    cudaMemset(taps, 0, 44 * 32 * 8); cudaMemset(x, 0, 4096 * 8);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 20000;
    for (int mode = 2; mode < 6; ++mode) {
        for (int rep = 0; rep < 2; ++rep) {
            cudaEventRecord(e0);
            if (mode == 0) k<0><<<148, 384, 4096 * 8>>>(out, taps, x, iters);
            else if (mode == 1) k<1><<<148, 384, 4096 * 8>>>(out, taps, x, iters);
            else if (mode == 2) k2<2><<<148, 384, 4096 * 8>>>(out, taps, x, iters);
            else if (mode == 3) k2<3><<<148, 384, 4096 * 8>>>(out, taps, x, iters);
            else if (mode == 4) k3<4><<<148, 384, 4096 * 8>>>(out, (const float*)taps, x, iters);
            else k3<5><<<148, 384, 4096 * 8>>>(out, (const float*)taps, x, iters);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            double ffma = 148.0 * 12 * 704.0 * iters;      // warp-level FFMAs
            printf("mode %d: %.3f ms  -> %.2f warp-FFMA/clk/SM at 1.965 GHz (peak 4), %.1f TFLOP/s\n", mode, ms,
                   ffma / (ms * 1e-3) / 148 / 1.965e9, ffma * 64 / (ms * 1e-3) / 1e12);
        }
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
