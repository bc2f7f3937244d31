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
#include <cuda_pipeline.h>

template <int N>
__global__ void __launch_bounds__(FftPlan<N>::THREADS)
taps_fft_kernel(const float2 *__restrict__ taps, int L, float2 *__restrict__ Hpos) {
    extern __shared__ __align__(16) float2 s[];
    constexpr int T = FftPlan<N>::THREADS;
    const int tid = threadIdx.x;
    for (int e = tid; e < N; e += T) s[FFT_PAD(e)] = (e < L) ? taps[e] : make_float2(0.f, 0.f);
    __syncthreads();
    fft_smem<N, false>(s, tid);
    const float sc = 1.0f / (float)N;                                     // fold the inverse transform's 1/N
    for (int p = tid; p < N; p += T) {
        const float2 v = s[FFT_PAD(p)];
        Hpos[p] = make_float2(v.x * sc, v.y * sc);
    }
}

// Stage raw complex memory C[k0 .. k0+N+4) of one receiver into LINEAR shared memory with cp.async (16-byte copies, no
// registers held, every copy of the CTA in flight at once); samples at or beyond `valid` are zero.  k0 must be EVEN
// (16-byte aligned source): a block whose first sample is odd — every other block when the AF filter length is even,
// V = N-(L-1) odd — stages from k0-1 and the readers skip one element (N+4 covers N+2 samples plus the shift).
template <int N, int T>
__device__ __forceinline__ void k2_stage_raw(float2 *raw, const float2 *__restrict__ C, i64 k0, i64 valid, int tid) {
    constexpr int CH = (N + 4) / 2;                                       // 16-byte chunks
    if (k0 + N + 4 <= valid) {                                            // interior block: no bounds checks
        const float2 *src = C + k0;
#pragma unroll
        for (int i = 0; i < CH / T; ++i) __pipeline_memcpy_async(raw + 2 * (tid + i * T), src + 2 * (tid + i * T), 16);
        if (tid < CH - (CH / T) * T) __pipeline_memcpy_async(raw + 2 * (tid + (CH / T) * T), src + 2 * (tid + (CH / T) * T), 16);
        return;
    }
    for (int c = tid; c < CH; c += T) {
        const i64 k = k0 + 2 * c;
        if (k + 1 < valid) {
            __pipeline_memcpy_async(raw + 2 * c, C + k, 16);
        } else {
            raw[2 * c] = (k < valid) ? C[k] : make_float2(0.f, 0.f);
            raw[2 * c + 1] = make_float2(0.f, 0.f);
        }
    }
}

// detector output for src index k0+e from the staged raw row (raw[e+2] <-> src[k0+e])
__device__ __forceinline__ float2 k2_detect(const float2 *raw, int e, int mode, bool ok) {
    if (!ok) return make_float2(0.f, 0.f);
    const float2 c2 = raw[e + 2];
    if (mode == PYSDR_MODE_AMSYNC) return make_float2(c2.x, 0.f);              // in-phase arm of the PLL-de-rotated memory
    if (mode == PYSDR_MODE_AM) return make_float2(sqrtf(c2.x * c2.x + c2.y * c2.y), 0.f);
    if (mode == PYSDR_MODE_NFM) {
        const float2 c0 = raw[e], c1 = raw[e + 1];
        const float dr = c2.x - c0.x, di = c2.y - c0.y;
        return make_float2(c1.x * di - c1.y * dr, 0.f);                         // nfm.m:126
    }
    return c2;
}

template <int N>
__global__ void __launch_bounds__(FftPlan<N>::THREADS, (N == 4096 ? 3 : 1))
af_fftconv_kernel(const FftConvArgs a) {
    extern __shared__ __align__(16) float2 s[];
    constexpr int T = FftPlan<N>::THREADS;
    constexpr int PER = N / T;
    constexpr int HALF = PER / 2;
    const int tid = threadIdx.x;
    const int rx = a.ua[blockIdx.y], rxb = a.ub[blockIdx.y];
    const bool pair = (N == 4096) && rxb >= 0;
    const int mode = a.mode[rx];
    const int L = a.L;
    const int V = N - (L - 1);
    const i64 k0 = (i64)blockIdx.x * V;                                   // first src index of this block
    const i64 avail = (i64)(L - 1) + a.n_out;                             // valid src samples; C holds avail + 2
    const float2 *C = a.C + (size_t)rx * a.c_stride;                      // C[k+2] <-> src[k]
    float *out = a.out + (size_t)rx * 2 * a.a_stride;
    float2 *s2 = s + FFT_SMEM_ELEMS(N);                                   // second buffer (N = 4096 launches only)
    __shared__ float s_max[2][FftPlan<N>::THREADS / 32];
    int pair_exp = 0;                                                     // b rides the transform scaled by 2^pair_exp
    pdl_trigger();
    pdl_wait();                                                           // K1's complex memory

    // ---- stage raw samples (async), detect from shared memory, lay out for the FFT ---------------------------------
    const int sh = (int)(k0 & 1);                                         // cp.async sources must be 16-byte aligned
    k2_stage_raw<N, T>(s, C, k0 - sh, avail + 2, tid);
    if (pair) k2_stage_raw<N, T>(s2, a.C + (size_t)rxb * a.c_stride, k0 - sh, avail + 2, tid);
    __pipeline_commit();
    __pipeline_wait_prior(0);
    __syncthreads();
    {
        float2 u[PER];
        const int modeb = pair ? a.mode[rxb] : 0;
#pragma unroll
        for (int i = 0; i < PER; ++i) {
            const int e = tid + i * T;
            const bool ok = k0 + e < avail;
            u[i] = k2_detect(s + sh, e, mode, ok);
            if (pair) u[i].y = k2_detect(s2 + sh, e, modeb, ok).x;             // two real detector outputs: u = det_a + j det_b
        }
        if (pair) {
            // The two signals share one transform, so its rounding error (~1e-7 of the LARGER one) lands on both.  Bring
            // them to the same binade first: b is scaled by an exact power of two (undone at the store), which keeps
            // e.g. the discriminator output of a carrier-less NFM channel (1e-6) accurate next to an AM envelope (1e-2).
            float ma = 0.f, mb = 0.f;
#pragma unroll
            for (int i = 0; i < PER; ++i) { ma = fmaxf(ma, fabsf(u[i].x)); mb = fmaxf(mb, fabsf(u[i].y)); }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                ma = fmaxf(ma, __shfl_xor_sync(0xffffffffu, ma, o));
                mb = fmaxf(mb, __shfl_xor_sync(0xffffffffu, mb, o));
            }
            if ((tid & 31) == 0) { s_max[0][tid >> 5] = ma; s_max[1][tid >> 5] = mb; }
        }
        __syncthreads();
        if (pair) {
            float ma = 0.f, mb = 0.f;
#pragma unroll
            for (int w = 0; w < T / 32; ++w) { ma = fmaxf(ma, s_max[0][w]); mb = fmaxf(mb, s_max[1][w]); }
            if (ma > 0.f && mb > 0.f && isfinite(ma) && isfinite(mb)) {
                int e = ilogbf(ma) - ilogbf(mb);
                e = e < -60 ? -60 : (e > 60 ? 60 : e);
                pair_exp = e;
                const float sb = scalbnf(1.0f, e);
#pragma unroll
                for (int i = 0; i < PER; ++i) u[i].y *= sb;
            }
        }
#pragma unroll
        for (int i = 0; i < PER; ++i) s[FFT_PAD(tid + i * T)] = u[i];
    }
    __syncthreads();
    fft_smem<N, false>(s, tid);

    float2 *sw = s;                                                       // buffer the inverse transform runs in
    if (pair) {
        // Z[k] -> Xa[k] = (Z[k] + conj Z[N-k]) / 2,  Xb[k] = (Z[k] - conj Z[N-k]) / 2j;  W = Ha Xa + j Hb Xb  -> s2
        const float2 *Ha = a.H + (size_t)rx * N, *Hb = a.H + (size_t)rxb * N;
        sw = s2;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            float2 ha[HALF], hb[HALF];
#pragma unroll
            for (int i = 0; i < HALF; ++i) {
                ha[i] = __ldg(Ha + tid + (h * HALF + i) * T);
                hb[i] = __ldg(Hb + tid + (h * HALF + i) * T);
            }
#pragma unroll
            for (int i = 0; i < HALF; ++i) {
                const int p = tid + (h * HALF + i) * T;
                const int kp = (N - fft_pos_to_freq<N>(p)) & (N - 1);
                const float2 zz = s[FFT_PAD(p)];
                const float2 zq = s[FFT_PAD(fft_pos_to_freq<N>(kp))];     // digit reversal is an involution for N = 16^3
                const float2 xa = make_float2(0.5f * (zz.x + zq.x), 0.5f * (zz.y - zq.y));
                const float2 xb = make_float2(0.5f * (zz.y + zq.y), -0.5f * (zz.x - zq.x));
                const float2 ya = cmul(xa, ha[i]), yb = cmul(xb, hb[i]);
                s2[FFT_PAD(p)] = make_float2(ya.x - yb.y, ya.y + yb.x);
            }
        }
    } else {
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
    fft_smem<N, true>(sw, tid);

    // ---- store the valid outputs of this block (32-bit loop, pointers hoisted) -----------------------------------
    const i64 left = a.n_out - k0;
    const int lim = left < (i64)V ? (int)left : V;
    const float2 *sv = sw;
    float *o0 = out + k0;
    if (pair) {
        float *o1 = a.out + (size_t)rxb * 2 * a.a_stride + k0;
        const float unscale = scalbnf(1.0f, -pair_exp);
#pragma unroll 4
        for (int j = tid; j < lim; j += T) {
            const float2 c = sv[FFT_PAD(j + L - 1)];
            o0[j] = c.x;
            o1[j] = c.y * unscale;
        }
    } else if (mode == PYSDR_MODE_IQ) {
        float2 *oc = (float2 *)out + k0;
#pragma unroll 4
        for (int j = tid; j < lim; j += T) oc[j] = sv[FFT_PAD(j + L - 1)];
    } else if (mode == PYSDR_MODE_CW) {
        const u64 inc = a.bfo_inc[rx];
        const u64 ph0 = inc * (u64)(a.m0 + k0);
#pragma unroll 4
        for (int j = tid; j < lim; j += T) {
            const float2 c = sv[FFT_PAD(j + L - 1)];
            const float2 cs = nco_cs(ph0 + inc * (u64)j);
            o0[j] = c.x * cs.x - c.y * cs.y;                              // Re{ z * e^{+j th} }
        }
    } else {
#pragma unroll 4
        for (int j = tid; j < lim; j += T) o0[j] = sv[FFT_PAD(j + L - 1)].x;
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
    taps_fft_kernel<N><<<1, FftPlan<N>::THREADS, smem, st>>>(d_taps, L, d_H);
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
static int fftconv_launch_n(const FftConvArgs &a0, int n_rx, cudaStream_t st) {
    const size_t smem = sizeof(float2) * FFT_SMEM_ELEMS(N) * (N == 4096 ? 2 : 1);     // paired units use a second buffer
    if (smem > 48 * 1024) CUDA_TRY(cudaFuncSetAttribute(af_fftconv_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    // work units: RAW rows have no demod filter; real-detector rows (AM / NFM / AM-Synch) pair up when the transform's
    // digit reversal is an involution (N = 16^3), everything else runs alone
    FftConvArgs a = a0;
    int n_units = 0, open_unit = -1;
    for (int r = 0; r < n_rx; ++r) {
        const int m = a.mode[r];
        if (m == PYSDR_MODE_RAW) continue;
        const bool real_det = (m == PYSDR_MODE_AM || m == PYSDR_MODE_NFM || m == PYSDR_MODE_AMSYNC);
        if (real_det && N == 4096 && open_unit >= 0) {
            a.ub[open_unit] = r;
            open_unit = -1;
            continue;
        }
        a.ua[n_units] = r;
        a.ub[n_units] = -1;
        if (real_det && N == 4096) open_unit = n_units;
        ++n_units;
    }
    if (n_units == 0) return PYSDR_OK;
    const int V = N - (a.L - 1);
    dim3 grid((unsigned)((a.n_out + V - 1) / V), (unsigned)n_units);
    CUDA_TRY(launch_pdl(af_fftconv_kernel<N>, grid, dim3(FftPlan<N>::THREADS), smem, st, a));
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

// ------------------------------------------------------------------------------------------------------------------------
// WFM video stage at the RF rate: LO + 1001-tap video FIR + FM discriminator in ONE kernel ("BCB FM is wideband so we need
// to demodulate first before resampling", reference gui.py:1703,1759-1762).  r01 ran the FIR in direct form through K1
// (8 008 flop per input sample, 31 TFLOP/s, 0.45 % of the HBM roofline); here it is an overlap-save fast convolution on
// the same shared-memory FFT as the AF filter:
//     y[n]  = e^{-j th(n)} * sum_j G[j] x[n-j],   G[j] = h[j] e^{+j w j}   (LO folded into the taps, like K1)
//     fm[n] = Re y[n-1] * Im d - Im y[n-1] * Re d,  d = y[n] - y[n-2]      (reference sigs/nfm.m:123-127)
// A CTA transforms N raw samples and emits S = N - (L-1) - 2 discriminator outputs: the two extra outputs of overlap
// make y[n-1], y[n-2] available in the block.  Carried state: the last L+1 RAW samples, and the last two y of the previous
// call (prev2) for the first two outputs — they are NOT recomputed from the raw memory, because after a retune the memory
// is seen through the new LO while y[-1], y[-2] were produced with the old one (what the reference's chunk-wise chain
// does).  prev2 is double buffered (slot `par` is read, slot par^1 written by the CTA that owns the last sample).
// Output is complex64 (fm, 0): the resampler bank reads complex.
struct WfmVidArgs {
    const float2 *x;          // n_in new samples
    const float2 *hist;       // L+1 samples preceding x
    const float2 *H;          // spectrum of G in position order, 1/N folded in
    float2 *fm;               // [n_in]
    float2 *prev2;            // [2][2]: y[-2], y[-1] of the previous call in slot par; the new pair goes to slot par ^ 1
    int par;
    i64 n_in;
    int L;
    u64 acc0, inc;            // LO phase at x[0], per-sample increment
};

template <int N>
__global__ void __launch_bounds__(FftPlan<N>::THREADS, (N == 4096 ? 3 : 1)) wfm_video_disc_kernel(const WfmVidArgs a) {
    extern __shared__ __align__(16) float2 s[];
    constexpr int T = FftPlan<N>::THREADS;
    constexpr int PER = N / T;
    const int tid = threadIdx.x;
    const int L = a.L, Hn = L + 1;
    const int S = N - (L - 1) - 2;
    const i64 g0 = (i64)blockIdx.x * S - Hn;                               // index (rel. x[0]) of this block's first sample
    pdl_trigger();
    pdl_wait();
    {
        float2 v[PER];
#pragma unroll
        for (int i = 0; i < PER; ++i) {
            const i64 g = g0 + tid + i * T;
            v[i] = g < 0 ? a.hist[g + Hn] : (g < a.n_in ? a.x[g] : make_float2(0.f, 0.f));
        }
#pragma unroll
        for (int i = 0; i < PER; ++i) s[FFT_PAD(tid + i * T)] = v[i];
    }
    __syncthreads();
    fft_smem<N, false>(s, tid);
    {
        float2 h[PER];
#pragma unroll
        for (int i = 0; i < PER; ++i) h[i] = __ldg(a.H + tid + i * T);
#pragma unroll
        for (int i = 0; i < PER; ++i) {
            const int p = tid + i * T;
            s[FFT_PAD(p)] = cmul(s[FFT_PAD(p)], h[i]);
        }
    }
    __syncthreads();
    fft_smem<N, true>(s, tid);
    // de-rotate the valid outputs (positions L-1 .. N-1) by the exact LO phase of their sample
    for (int j = L - 1 + tid; j < N; j += T) {
        const float2 cs = nco_cs(a.acc0 + a.inc * (u64)(g0 + j));
        const float2 z = s[FFT_PAD(j)];
        s[FFT_PAD(j)] = make_float2(z.x * cs.x + z.y * cs.y, z.y * cs.x - z.x * cs.y);
    }
    __syncthreads();
    const float2 *pin = a.prev2 + 2 * a.par;
    for (int j = L + 1 + tid; j < N; j += T) {
        const i64 n = g0 + j;
        if (n >= a.n_in) break;
        const float2 c2 = s[FFT_PAD(j)];
        const float2 c1 = n >= 1 ? s[FFT_PAD(j - 1)] : pin[1];
        const float2 c0 = n >= 2 ? s[FFT_PAD(j - 2)] : pin[n];             // n = 0 -> y[-2], n = 1 -> y[-1]
        const float dr = c2.x - c0.x, di = c2.y - c0.y;
        a.fm[n] = make_float2(c1.x * di - c1.y * dr, 0.f);
    }
    if (blockIdx.x == gridDim.x - 1 && tid == 0) {                         // this CTA owns the last sample of the call
        const int jl = (int)(a.n_in - 1 - g0);
        float2 *pout = a.prev2 + 2 * (a.par ^ 1);
        pout[1] = s[FFT_PAD(jl)];
        pout[0] = a.n_in >= 2 ? s[FFT_PAD(jl - 1)] : pin[1];
    }
}

// dst[0..n_keep) <- the last n_keep samples of [hist | x]
__global__ void wfm_hist_kernel(float2 *hist, const float2 *__restrict__ x, int n_keep, i64 n_in) {
    float2 tmp[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const int e = threadIdx.x + k * 1024;
        if (e < n_keep) {
            const i64 idx = n_in - n_keep + e;
            tmp[k] = idx >= 0 ? x[idx] : hist[n_keep + idx];
        }
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const int e = threadIdx.x + k * 1024;
        if (e < n_keep) hist[e] = tmp[k];
    }
}

extern "C" int pysdr_fir_spectrum(const void *d_taps_c64, int L, void *d_H, void *stream) {
    if (!d_taps_c64 || !d_H || !fftconv_supported(L)) { pysdr_set_error("fir_spectrum: unsupported filter length %d", L); return PYSDR_ERR_ARG; }
    return fftconv_prepare_taps((const float2 *)d_taps_c64, L, (float2 *)d_H, (cudaStream_t)stream);
}

extern "C" int pysdr_wfm_video_disc(const void *d_x, int64_t n_in, void *d_hist, void *d_prev2, int prev2_slot, const void *d_H,
                                    int L, uint64_t acc0, uint64_t inc, void *d_fm, void *stream) {
    if (!d_x || !d_hist || !d_prev2 || !d_H || !d_fm || n_in < 0 || L < 2 || (prev2_slot & ~1)) { pysdr_set_error("wfm_video_disc: bad arguments"); return PYSDR_ERR_ARG; }
    if (n_in == 0) return PYSDR_OK;
    const int N = fftconv_n_for(L + 2);
    if (N == 0 || L + 1 > 8192) { pysdr_set_error("wfm_video_disc: video filter length %d unsupported", L); return PYSDR_ERR_ARG; }
    cudaStream_t st = (cudaStream_t)stream;
    WfmVidArgs a;
    a.x = (const float2 *)d_x; a.hist = (const float2 *)d_hist; a.H = (const float2 *)d_H; a.fm = (float2 *)d_fm;
    a.n_in = n_in; a.L = L; a.acc0 = acc0; a.inc = inc;
    a.prev2 = (float2 *)d_prev2; a.par = prev2_slot;
    if (fftconv_n_for(L) != N) { pysdr_set_error("wfm_video_disc: video filter length %d unsupported", L); return PYSDR_ERR_ARG; }
    const int S = N - (L - 1) - 2;
    const unsigned grid = (unsigned)((n_in + S - 1) / S);
    const size_t smem = sizeof(float2) * FFT_SMEM_ELEMS(N);
    if (N == 4096) {
        if (smem > 48 * 1024) CUDA_TRY(cudaFuncSetAttribute(wfm_video_disc_kernel<4096>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        CUDA_TRY(launch_pdl(wfm_video_disc_kernel<4096>, dim3(grid), dim3(FftPlan<4096>::THREADS), smem, st, a));
    } else {
        if (smem > 48 * 1024) CUDA_TRY(cudaFuncSetAttribute(wfm_video_disc_kernel<8192>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        CUDA_TRY(launch_pdl(wfm_video_disc_kernel<8192>, dim3(grid), dim3(FftPlan<8192>::THREADS), smem, st, a));
    }
    wfm_hist_kernel<<<1, 1024, 0, st>>>((float2 *)d_hist, (const float2 *)d_x, L + 1, n_in);
    LAUNCH_CHECK();
    return PYSDR_OK;
}
