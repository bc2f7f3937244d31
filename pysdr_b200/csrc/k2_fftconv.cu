// K2 (fast variant): audio-rate detection + AF FIR as overlap-save fast convolution in shared memory.
//
// Same arithmetic contract as the direct-form af_fir_kernel in bank.cu:
//     a[o] = sum_j g[j] * src[o + (L-1) - j]          (src = detected / complex baseband memory + new samples)
// evaluated per block of V = N - (L-1) outputs as  IFFT( H_pos * FFT(u) ), u = N consecutive src samples.
// Detection is fused into the load (AM |.|, NFM discriminator of reference sigs/nfm.m:123-127), the BFO
// re-insertion for CW into the store.  Stands behind dsp.Receiver.demod_data's demod stage
// (reference receiver.py:235) and mirrors the reference's own FFT convolver (dsp.convolver.convolve_fast,
// receiver.py:862).  One launch serves all receivers (blockIdx.y = receiver).
#include "common.cuh"
#include "fft_smem.cuh"

template <int N>
__global__ void __launch_bounds__(FftPlan<N>::THREADS)
taps_fft_kernel(const float2 *__restrict__ taps, int L, float2 *__restrict__ Hpos, const float2 *__restrict__ tw) {
    extern __shared__ __align__(16) float2 s[];
    constexpr int T = FftPlan<N>::THREADS;
    const int tid = threadIdx.x;
    for (int e = tid; e < N; e += T) s[FFT_PAD(e)] = (e < L) ? taps[e] : make_float2(0.f, 0.f);
    __syncthreads();
    fft_smem<N, false>(s, tid, tw);
    const float sc = 1.0f / (float)N;                                     // fold the inverse transform's 1/N
    for (int p = tid; p < N; p += T) {
        const float2 v = s[FFT_PAD(p)];
        Hpos[p] = make_float2(v.x * sc, v.y * sc);
    }
}

template <int N>
__global__ void __launch_bounds__(FftPlan<N>::THREADS)
af_fftconv_kernel(const FftConvArgs a, const float2 *__restrict__ tw) {
    extern __shared__ __align__(16) float2 s[];
    constexpr int T = FftPlan<N>::THREADS;
    const int tid = threadIdx.x;
    const int rx = blockIdx.y;
    const int mode = a.mode[rx];
    if (mode == PYSDR_MODE_RAW) return;                                   // no demod filter for this receiver (whole CTA)
    const int L = a.L;
    const int V = N - (L - 1);
    const i64 k0 = (i64)blockIdx.x * V;                                   // first src index of this block
    const i64 avail = (i64)(L - 1) + a.n_out;                             // valid src samples
    const float2 *C = a.C + (size_t)rx * a.c_stride;                      // C[k+2] <-> src[k] for complex modes

    // ---- load + fused detection (fully unrolled: all global loads of a thread are in flight together) -------
    constexpr int PER = N / T;
    {
        float2 c0[PER], c1[PER], c2[PER];
#pragma unroll
        for (int i = 0; i < PER; ++i) {
            const i64 k = k0 + tid + i * T;
            c0[i] = c1[i] = c2[i] = make_float2(0.f, 0.f);
            if (k < avail) {
                c2[i] = C[k + 2];
                if (mode == PYSDR_MODE_NFM) { c0[i] = C[k]; c1[i] = C[k + 1]; }
            }
        }
#pragma unroll
        for (int i = 0; i < PER; ++i) {
            float2 u = c2[i];
            if (mode == PYSDR_MODE_AMSYNC) {
                u = make_float2(c2[i].x, 0.f);                           // in-phase arm of the PLL-de-rotated memory
            } else if (mode == PYSDR_MODE_AM) {
                u = make_float2(sqrtf(c2[i].x * c2[i].x + c2[i].y * c2[i].y), 0.f);
            } else if (mode == PYSDR_MODE_NFM) {
                const float dr = c2[i].x - c0[i].x, di = c2[i].y - c0[i].y;
                u = make_float2(c1[i].x * di - c1[i].y * dr, 0.f);        // nfm.m:126
            }
            s[FFT_PAD(tid + i * T)] = u;
        }
    }
    __syncthreads();
    fft_smem<N, false>(s, tid, tw);
    {
        const float2 *H = a.H + (size_t)rx * N;
        float2 h[PER];
#pragma unroll
        for (int i = 0; i < PER; ++i) h[i] = __ldg(H + tid + i * T);
#pragma unroll
        for (int i = 0; i < PER; ++i) {
            const int p = tid + i * T;
            s[FFT_PAD(p)] = cmul(s[FFT_PAD(p)], h[i]);
        }
    }
    __syncthreads();
    fft_smem<N, true>(s, tid, tw);

    // ---- store the V valid outputs -----------------------------------------------------------------------
    float *out = a.out + (size_t)rx * 2 * a.a_stride;
    for (int e = (L - 1) + tid; e < N; e += T) {
        const i64 o = k0 + e - (L - 1);
        if (o >= a.n_out) break;
        const float2 c = s[FFT_PAD(e)];
        if (mode == PYSDR_MODE_IQ) {
            ((float2 *)out)[o] = c;
        } else if (mode == PYSDR_MODE_CW) {
            const float2 cs = nco_cs(a.bfo_inc[rx] * (u64)(a.m0 + o));
            out[o] = c.x * cs.x - c.y * cs.y;                             // Re{ z * e^{+j th} }
        } else {
            out[o] = c.x;
        }
    }
}

int fftconv_n_for(int L) {
    if (L <= 2049) return 4096;                // V = N-(L-1) >= 2048 valid outputs per block
    if (L <= 4097) return 8192;
    return 0;
}

int fftconv_supported(int L) { return fftconv_n_for(L) != 0 && L >= 2; }

template <int N>
static int taps_fft_launch(const float2 *d_taps, int L, float2 *d_H, cudaStream_t st) {
    const size_t smem = sizeof(float2) * FFT_SMEM_ELEMS(N);
    if (smem > 48 * 1024) CUDA_TRY(cudaFuncSetAttribute(taps_fft_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const float2 *tw = fft_twiddles(N);
    if (!tw) { pysdr_set_error("fft twiddle table allocation failed"); return PYSDR_ERR_CUDA; }
    taps_fft_kernel<N><<<1, FftPlan<N>::THREADS, smem, st>>>(d_taps, L, d_H, tw);
    LAUNCH_CHECK();
    return PYSDR_OK;
}

int fftconv_prepare_taps(const float2 *d_taps, int L, float2 *d_H, cudaStream_t st) {
    const int N = fftconv_n_for(L);
    if (N == 4096) return taps_fft_launch<4096>(d_taps, L, d_H, st);
    if (N == 8192) return taps_fft_launch<8192>(d_taps, L, d_H, st);
    pysdr_set_error("fftconv: unsupported AF filter length %d", L);
    return PYSDR_ERR_ARG;
}

template <int N>
static int fftconv_launch_n(const FftConvArgs &a, int n_rx, cudaStream_t st) {
    const size_t smem = sizeof(float2) * FFT_SMEM_ELEMS(N);
    if (smem > 48 * 1024) CUDA_TRY(cudaFuncSetAttribute(af_fftconv_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int V = N - (a.L - 1);
    dim3 grid((unsigned)((a.n_out + V - 1) / V), (unsigned)n_rx);
    const float2 *tw = fft_twiddles(N);
    if (!tw) { pysdr_set_error("fft twiddle table allocation failed"); return PYSDR_ERR_CUDA; }
    af_fftconv_kernel<N><<<grid, FftPlan<N>::THREADS, smem, st>>>(a, tw);
    LAUNCH_CHECK();
    return PYSDR_OK;
}

int fftconv_launch(const FftConvArgs &a, int n_rx, cudaStream_t st) {
    if (a.n_out <= 0) return PYSDR_OK;
    const int N = fftconv_n_for(a.L);
    if (N == 4096) return fftconv_launch_n<4096>(a, n_rx, st);
    if (N == 8192) return fftconv_launch_n<8192>(a, n_rx, st);
    pysdr_set_error("fftconv: unsupported AF filter length %d", a.L);
    return PYSDR_ERR_ARG;
}
