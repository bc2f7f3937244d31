// psd_fast.cu — K3 fast path: the periodogram frames of dsp.spectrum.periodogram / psd_est (reference Plotting.py:376-377,462;
// window -> zero-pad -> FFT -> |X|^2 -> mean over the line's frames) for power-of-two NFFT >= 512 with plain hop stepping.
//
// Differences from the generic psd_frames_kernel (psd.cu), which stays for small NFFT, sub-step frame starts (RTTY) and short
// captures:
//   * N = 16^A x RF (8192 = 16^3 x 2, 4096 = 16^2 x 16, ...).  The FIRST radix-16 pass runs on the registers the global loads
//     land in (thread t loads x[t + i N/16], i = 0..15: exactly one first-stage butterfly), the LAST radix-RF pass is fused with
//     the |X|^2 accumulation — shared memory sees 2 (A = 2) or 3 (A = 3) store passes and as many load passes instead of 5 + 5.
//   * Twiddles are per-thread constants across frames (the butterfly a thread owns in a pass does not change): tables
//     W[k][m] computed in float64 on the host, copied into shared memory once per CTA, one conflict-free LDS.64 per
//     multiplication — no sincospif and no 14-multiply power tree per butterfly.
//   * persistent CTAs over lines (the tables are loaded once), packed FP32x2 arithmetic (fft_smem.cuh, FFT_PACKED = 1).
// Spectra stay in POSITION order; psd_fast_pos_to_freq gives the bin of a position for the finalize kernel.
#include "common.cuh"
#define FFT_PACKED 1
#include "fft_smem.cuh"
#include "psd_fast.cuh"

template <int N>
struct FastPlan {
    static constexpr int LOG2 = FftPlan<N>::LOG2;
    static constexpr int RF = (LOG2 % 4 == 0) ? 16 : (1 << (LOG2 % 4));      // final radix, fused with the accumulation
    static constexpr int A = (LOG2 % 4 == 0) ? LOG2 / 4 - 1 : LOG2 / 4;      // radix-16 passes before it
    static constexpr int T = N / 16;
    static constexpr int S2 = N / 256;                                        // butterfly stride of the second pass
    static constexpr int S3 = N / 4096;                                       // ... of the third (A = 3)
    static_assert(A == 2 || A == 3, "psd_fast: N = 16^2 or 16^3 times 2..16");
};

int psd_fast_supported(int nfft) { return nfft == 512 || nfft == 1024 || nfft == 2048 || nfft == 4096 || nfft == 8192; }

int psd_fast_pos_to_freq(int nfft, int p) {
    int lg = 0;
    while ((1 << lg) < nfft) ++lg;
    const int rf = (lg % 4 == 0) ? 16 : (1 << (lg % 4));
    const int a = (lg % 4 == 0) ? lg / 4 - 1 : lg / 4;
    int k = 0, mul = 1, len = nfft;
    for (int i = 0; i < a; ++i) { len /= 16; k += (p / len) * mul; p %= len; mul *= 16; }
    (void)rf;
    return k + p * mul;                                                       // p < RF now
}

// number of float2 twiddles: [15][T] + [15][S2] (+ [15][S3])
size_t psd_fast_table_elems(int nfft) { return 15 * ((size_t)nfft / 16 + (size_t)nfft / 256 + (nfft >= 4096 * 2 ? (size_t)nfft / 4096 : 0)); }

void psd_fast_fill_table(int nfft, float2 *t) {
    const int T = nfft / 16, S2 = nfft / 256, S3 = nfft / 4096;
    size_t o = 0;
    auto put = [&](int ni, int stride) {                                      // W_ni^(k m), k = 1..15, m < stride
        for (int k = 1; k < 16; ++k)
            for (int m = 0; m < stride; ++m) {
                const double ang = -2.0 * M_PI * (double)((long long)k * m % ni) / (double)ni;
                t[o++] = make_float2((float)cos(ang), (float)sin(ang));
            }
    };
    put(nfft, T);
    put(nfft / 16, S2);
    if (nfft >= 8192) put(nfft / 256, S3);
}

template <int N, bool CPLX>
__global__ void __launch_bounds__(N / 16, 1) psd_frames_fast_kernel(const void *__restrict__ xv, const float *__restrict__ win, int chunk, int hop,
                                                                    int navg, i64 n_lines, const float2 *__restrict__ tw, float *__restrict__ part) {
    using P = FastPlan<N>;
    constexpr int T = P::T, RF = P::RF;
    extern __shared__ __align__(16) float2 sm_all[];
    float2 *s = sm_all;                                                       // FFT_SMEM_ELEMS(N)
    float2 *tA = s + FFT_SMEM_ELEMS(N);                                       // [15][T]
    float2 *tB = tA + 15 * T;                                                 // [15][S2]
    float2 *tC = tB + 15 * P::S2;                                             // [15][S3] (A = 3)
    const int tid = threadIdx.x;
    {
        constexpr int NT = 15 * (T + P::S2 + (P::A == 3 ? P::S3 : 0));
        for (int i = tid; i < NT; i += T) tA[i] = tw[i];
    }
    float wr[16];                                                             // the window taps this thread applies, every frame
#pragma unroll
    for (int i = 0; i < 16; ++i) { const int e = tid + i * T; wr[i] = e < chunk ? __ldg(win + e) : 0.f; }
    __syncthreads();

    for (i64 line = blockIdx.x; line < n_lines; line += gridDim.x) {
        float acc[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) acc[i] = 0.f;
        float2 v[16];
        auto load_frame = [&](i64 fr) {                                       // raw samples of frame fr -> v (window applied later)
            const i64 start = fr * (i64)hop;
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                const int e = tid + i * T;
                v[i] = make_float2(0.f, 0.f);
                if (e < chunk) {
                    if (CPLX) v[i] = ((const float2 *)xv)[start + e];
                    else v[i].x = ((const float *)xv)[start + e];
                }
            }
        };
        load_frame(line * navg);
        for (int f = 0; f < navg; ++f) {
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] = make_float2(v[i].x * wr[i], v[i].y * wr[i]);
            // pass 1: the butterfly of stride T this thread has just loaded
            dft_reg<16, false>(v);
#pragma unroll
            for (int k = 1; k < 16; ++k) v[k] = cmul(v[k], tA[(k - 1) * T + tid]);
#pragma unroll
            for (int q = 0; q < 16; ++q) s[FFT_PAD(tid + q * T)] = v[q];
            __syncthreads();
            // pass 2: blocks of N/16, stride S2
            {
                constexpr int NI = N / 16, ST = P::S2;
                const int blk = tid / ST, m = tid - blk * ST, base = blk * NI + m;
#pragma unroll
                for (int q = 0; q < 16; ++q) v[q] = s[FFT_PAD(base + q * ST)];
                dft_reg<16, false>(v);
#pragma unroll
                for (int k = 1; k < 16; ++k) v[k] = cmul(v[k], tB[(k - 1) * ST + m]);
#pragma unroll
                for (int q = 0; q < 16; ++q) s[FFT_PAD(base + q * ST)] = v[q];
                __syncthreads();
            }
            if constexpr (P::A == 3) {                                        // pass 3: blocks of N/256, stride S3 = RF
                constexpr int NI = N / 256, ST = P::S3;
                const int blk = tid / ST, m = tid - blk * ST, base = blk * NI + m;
#pragma unroll
                for (int q = 0; q < 16; ++q) v[q] = s[FFT_PAD(base + q * ST)];
                dft_reg<16, false>(v);
#pragma unroll
                for (int k = 1; k < 16; ++k) v[k] = cmul(v[k], tC[(k - 1) * ST + m]);
#pragma unroll
                for (int q = 0; q < 16; ++q) s[FFT_PAD(base + q * ST)] = v[q];
                __syncthreads();
            }
            // the next frame's samples travel while the last pass and the barrier run (v is dead until the next iteration)
            if (f + 1 < navg) load_frame(line * navg + f + 1);
            // final radix-RF butterflies on RF consecutive points, fused with |X|^2
#pragma unroll
            for (int j = 0; j < 16 / RF; ++j) {
                const int base = (tid + j * T) * RF;
                float2 w[RF];
#pragma unroll
                for (int q = 0; q < RF; ++q) w[q] = s[FFT_PAD(base + q)];
                dft_reg<RF, false>(w);
#pragma unroll
                for (int q = 0; q < RF; ++q) acc[j * RF + q] = fmaf(w[q].x, w[q].x, fmaf(w[q].y, w[q].y, acc[j * RF + q]));
            }
            __syncthreads();
        }
        float *pp = part + (size_t)line * N;
#pragma unroll
        for (int j = 0; j < 16 / RF; ++j)
#pragma unroll
            for (int q = 0; q < RF; ++q) pp[(tid + j * T) * RF + q] = acc[j * RF + q];
    }
}

template <int N>
static int launch_n(const void *d_x, int is_complex, const float *d_win, int chunk, int hop, int navg, i64 n_lines, const float2 *d_tw,
                    float *d_part, cudaStream_t st) {
    using P = FastPlan<N>;
    const size_t smem = sizeof(float2) * (FFT_SMEM_ELEMS(N) + 15 * (P::T + P::S2 + (P::A == 3 ? P::S3 : 0)));
    static unsigned long long attr_done = 0ull;
    const unsigned long long dev_bit = 1ull << (pysdr_device() & 63);
    if (!(attr_done & dev_bit)) {
        CUDA_TRY(cudaFuncSetAttribute(psd_frames_fast_kernel<N, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        CUDA_TRY(cudaFuncSetAttribute(psd_frames_fast_kernel<N, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_done |= dev_bit;
    }
    int per_sm = 1;
    CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, psd_frames_fast_kernel<N, true>, N / 16, smem));
    if (per_sm < 1) per_sm = 1;
    i64 grid = (i64)pysdr_sm_count() * per_sm;
    if (grid > n_lines) grid = n_lines;
    if (is_complex)
        psd_frames_fast_kernel<N, true><<<(unsigned)grid, N / 16, smem, st>>>(d_x, d_win, chunk, hop, navg, n_lines, d_tw, d_part);
    else
        psd_frames_fast_kernel<N, false><<<(unsigned)grid, N / 16, smem, st>>>(d_x, d_win, chunk, hop, navg, n_lines, d_tw, d_part);
    LAUNCH_CHECK();
    return PYSDR_OK;
}

int psd_fast_launch(int nfft, const void *d_x, int is_complex, const float *d_win, int chunk, int hop, int navg, i64 n_lines,
                    const float2 *d_tw, float *d_part, cudaStream_t st) {
    switch (nfft) {
        case 512: return launch_n<512>(d_x, is_complex, d_win, chunk, hop, navg, n_lines, d_tw, d_part, st);
        case 1024: return launch_n<1024>(d_x, is_complex, d_win, chunk, hop, navg, n_lines, d_tw, d_part, st);
        case 2048: return launch_n<2048>(d_x, is_complex, d_win, chunk, hop, navg, n_lines, d_tw, d_part, st);
        case 4096: return launch_n<4096>(d_x, is_complex, d_win, chunk, hop, navg, n_lines, d_tw, d_part, st);
        case 8192: return launch_n<8192>(d_x, is_complex, d_win, chunk, hop, navg, n_lines, d_tw, d_part, st);
    }
    pysdr_set_error("psd_fast: unsupported nfft %d", nfft);
    return PYSDR_ERR_ARG;
}
