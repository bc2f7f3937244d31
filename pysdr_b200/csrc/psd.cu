// psd.cu — K3: windowed, zero-padded FFT periodogram with segment averaging, dB mapping and fftshift
// (dsp.spectrum.periodogram / psd_est; reference Plotting.py:376-377,462, sigs/iq.py:75-79; FFT idiom
// rtty.py:839-841) and the three_box_plot waterfall compute (reference Plotting.py:536-548,583-587,
// 618-626,689-695).
//
// One CTA owns one (line, split) work item: it loops over its frames, loading chunk_size samples times
// the window into shared memory (zero padded to nfft), runs the shared-memory radix-16 FFT (fft_smem.cuh)
// and accumulates |X|^2 in registers in POSITION order; only the final line is permuted to bins and
// fftshifted.  cuFFT is not used anywhere (tests cross-check against numpy).
#include "common.cuh"
#define FFT_PACKED 0                     /* see fft_smem.cuh: this kernel sits at its register limit */
#include "fft_smem.cuh"
#include "psd_fast.cuh"

struct pysdr_psd {
    int chunk, nfft, hop;
    double wsum2;
    float *d_win;
    float *d_part;         // partial sums workspace [lines][split][nfft] (position order)
    size_t part_cap;
    i64 launches;
    int sub;               // frame f starts at (f / sub) * hop + sub_off[f % sub]   (sub = 1, sub_off = {0}: plain hop)
    int sub_off[8];
    int flags;             // PYSDR_PSD_RAW | PYSDR_PSD_FLIP
    float2 *d_tw;          // twiddle tables of the fast path (psd_fast.cu), or null
};

struct PsdSteps { int sub; int off[8]; };

// grid (n_split, n_lines). Frames of line l: [l*navg, (l+1)*navg); split s takes frames s, s+n_split, ...
template <int N, bool CPLX>
__global__ void __launch_bounds__(FftPlan<N>::THREADS)
psd_frames_kernel(const void *__restrict__ xv, const float *__restrict__ win, int chunk, int hop, const PsdSteps steps, int navg,
                  int n_split, float *__restrict__ part) {
    extern __shared__ __align__(16) float2 s[];
    constexpr int T = FftPlan<N>::THREADS;
    constexpr int PER = (N + T - 1) / T;
    const int tid = threadIdx.x;
    const int line = blockIdx.y, split = blockIdx.x;
    float acc[PER];
#pragma unroll
    for (int i = 0; i < PER; ++i) acc[i] = 0.f;
    for (int f = split; f < navg; f += n_split) {
        const i64 fr = (i64)line * navg + f;
        const i64 start = (steps.sub == 1) ? fr * hop : (fr / steps.sub) * hop + steps.off[fr % steps.sub];
        {                                                        // unrolled: all of a thread's loads in flight together
            float2 xr[PER];
            float wr[PER];
#pragma unroll
            for (int i = 0; i < PER; ++i) {
                const int e = tid + i * T;
                xr[i] = make_float2(0.f, 0.f);
                wr[i] = 0.f;
                if (e < chunk && e < N) {
                    wr[i] = __ldg(win + e);
                    if (CPLX) xr[i] = ((const float2 *)xv)[start + e];
                    else xr[i].x = ((const float *)xv)[start + e];
                }
            }
#pragma unroll
            for (int i = 0; i < PER; ++i) {
                const int e = tid + i * T;
                if (e < N) s[FFT_PAD(e)] = make_float2(xr[i].x * wr[i], xr[i].y * wr[i]);
            }
        }
        __syncthreads();
        fft_smem<N, false>(s, tid);
#pragma unroll
        for (int i = 0; i < PER; ++i) {
            const int p = tid + i * T;
            if (p < N) {
                const float2 X = s[FFT_PAD(p)];
                acc[i] = fmaf(X.x, X.x, fmaf(X.y, X.y, acc[i]));
            }
        }
        __syncthreads();
    }
    float *pp = part + ((size_t)line * n_split + split) * N;
#pragma unroll
    for (int i = 0; i < PER; ++i) {
        const int p = tid + i * T;
        if (p < N) pp[p] = acc[i];
    }
}

// position -> bin of the fast path's pass order (radix-16 passes first, the odd radix last: psd_fast.cu)
template <int N>
__device__ __forceinline__ int fast_pos_to_freq(int p) {
    constexpr int LG = FftPlan<N>::LOG2;
    constexpr int A = (LG % 4 == 0) ? LG / 4 - 1 : LG / 4;
    int k = 0, mul = 1, len = N;
#pragma unroll
    for (int i = 0; i < A; ++i) { len /= 16; k += (p / len) * mul; p %= len; mul *= 16; }
    return k + p * mul;
}

// out[line][fftshift(bin(p))] = dB( sum_split part[p] / (navg * wsum2) )
template <int N, bool FAST = false>
__global__ void psd_finalize_kernel(const float *__restrict__ part, int n_split, float scale, int dB, int flags,
                                    float *__restrict__ out) {
    const int line = blockIdx.y;
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= N) return;
    const float *q = part + (size_t)line * n_split * N + p;
    float sum = 0.f;
    for (int sp = 0; sp < n_split; ++sp) sum += q[(size_t)sp * N];
    float v = sum * scale;
    if (dB) v = 10.f * log10f((flags & PYSDR_PSD_RAW) ? v : fmaxf(v, 1.0e-30f));
    const int k = FAST ? fast_pos_to_freq<N>(p) : fft_pos_to_freq<N>(p);
    int col = (k + N / 2) % N;                                           // fftshift
    if (flags & PYSDR_PSD_FLIP) col = N - 1 - col;                       // np.flipud of the shifted line (rtty.py:843)
    out[(size_t)line * N + col] = v;
}

extern "C" int pysdr_psd_create(int32_t chunk, int32_t nfft, int32_t hop, const float *window, pysdr_psd **out) {
    if (!out || !window || chunk < 1 || nfft < chunk || hop < 1 || (nfft & (nfft - 1)) != 0 || nfft > 16384 || nfft < 64) {
        pysdr_set_error("psd_create: need 1 <= chunk <= nfft, nfft a power of two in [64,16384] (got chunk=%d nfft=%d hop=%d)", chunk,
                        nfft, hop);
        return PYSDR_ERR_ARG;
    }
    pysdr_psd *p = new pysdr_psd();
    p->chunk = chunk; p->nfft = nfft; p->hop = hop;
    p->wsum2 = 0.0;
    for (int i = 0; i < chunk; ++i) p->wsum2 += (double)window[i] * (double)window[i];
    p->d_part = nullptr; p->part_cap = 0; p->launches = 0;
    p->sub = 1; p->flags = 0;
    p->d_tw = nullptr;
    for (int i = 0; i < 8; ++i) p->sub_off[i] = 0;
    if (cudaMalloc(&p->d_win, sizeof(float) * chunk) != cudaSuccess) {
        pysdr_set_error("psd_create: cudaMalloc failed");
        delete p;
        return PYSDR_ERR_CUDA;
    }
    CUDA_TRY(cudaMemcpy(p->d_win, window, sizeof(float) * chunk, cudaMemcpyHostToDevice));
    if (psd_fast_supported(nfft)) {
        std::vector<float2> tw(psd_fast_table_elems(nfft));
        psd_fast_fill_table(nfft, tw.data());
        CUDA_TRY(cudaMalloc(&p->d_tw, sizeof(float2) * tw.size()));
        CUDA_TRY(cudaMemcpy(p->d_tw, tw.data(), sizeof(float2) * tw.size(), cudaMemcpyHostToDevice));
    }
    *out = p;
    return PYSDR_OK;
}

extern "C" int pysdr_psd_destroy(pysdr_psd *p) {
    if (!p) return PYSDR_OK;
    cudaFree(p->d_win); cudaFree(p->d_part); cudaFree(p->d_tw);
    delete p;
    return PYSDR_OK;
}

extern "C" int pysdr_psd_configure(pysdr_psd *p, int32_t sub, const int32_t *sub_off, int32_t flags) {
    if (!p || sub < 1 || sub > 8 || (sub > 1 && !sub_off)) { pysdr_set_error("psd_configure: need 1 <= sub <= 8"); return PYSDR_ERR_ARG; }
    for (int i = 0; i < sub; ++i) {
        const int o = sub > 1 ? sub_off[i] : 0;
        if (o < 0 || o >= p->hop || (i > 0 && o < p->sub_off[i - 1])) {
            pysdr_set_error("psd_configure: sub-step offsets must be ascending within [0, hop)");
            return PYSDR_ERR_ARG;
        }
        p->sub_off[i] = o;
    }
    p->sub = sub; p->flags = flags;
    return PYSDR_OK;
}

extern "C" int64_t pysdr_psd_launch_count(const pysdr_psd *p) { return p ? p->launches : -1; }

template <int N>
static int psd_launch(pysdr_psd *p, const void *d_x, int is_complex, int navg, int n_split, i64 n_lines, int dB, float *d_out,
                      cudaStream_t st) {
    const size_t smem = sizeof(float2) * FFT_SMEM_ELEMS(N);
    if (smem > 48 * 1024) {
        CUDA_TRY(cudaFuncSetAttribute(psd_frames_kernel<N, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        CUDA_TRY(cudaFuncSetAttribute(psd_frames_kernel<N, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    }
    if (n_split == 1 && p->d_tw && p->sub == 1 && N >= 512 && N <= 8192 && !getenv("PYSDR_PSD_GENERIC")) {
        // fast path (psd_fast.cu): enough lines to fill the GPU with one CTA per line, plain hop stepping
        int rc = psd_fast_launch(N, d_x, is_complex, p->d_win, p->chunk, p->hop, navg, n_lines, p->d_tw, p->d_part, st);
        if (rc) return rc;
        const float scale = (p->flags & PYSDR_PSD_RAW) ? 1.0f / (float)navg : (float)(1.0 / ((double)navg * p->wsum2));
        dim3 g2((unsigned)((N + 255) / 256), (unsigned)n_lines);
        psd_finalize_kernel<N, true><<<g2, 256, 0, st>>>(p->d_part, 1, scale, dB, p->flags, d_out);
        LAUNCH_CHECK();
        p->launches += 2;
        return PYSDR_OK;
    }
    dim3 grid((unsigned)n_split, (unsigned)n_lines);
    PsdSteps steps;
    steps.sub = p->sub;
    for (int i = 0; i < 8; ++i) steps.off[i] = p->sub_off[i];
    if (is_complex)
        psd_frames_kernel<N, true><<<grid, FftPlan<N>::THREADS, smem, st>>>(d_x, p->d_win, p->chunk, p->hop, steps, navg, n_split, p->d_part);
    else
        psd_frames_kernel<N, false><<<grid, FftPlan<N>::THREADS, smem, st>>>(d_x, p->d_win, p->chunk, p->hop, steps, navg, n_split, p->d_part);
    LAUNCH_CHECK();
    const float scale = (p->flags & PYSDR_PSD_RAW) ? 1.0f / (float)navg : (float)(1.0 / ((double)navg * p->wsum2));
    dim3 g2((unsigned)((N + 255) / 256), (unsigned)n_lines);
    psd_finalize_kernel<N><<<g2, 256, 0, st>>>(p->d_part, n_split, scale, dB, p->flags, d_out);
    LAUNCH_CHECK();
    p->launches += 2;
    return PYSDR_OK;
}

static int psd_lines_batch(pysdr_psd *p, const void *d_x, int is_complex, int navg, i64 n_lines, int dB, float *d_out, cudaStream_t st) {
    int n_split = 1;
    if (n_lines < 296) {
        n_split = (int)((296 + n_lines - 1) / n_lines);
        if (n_split > navg) n_split = navg;
    }
    const size_t need = (size_t)n_lines * n_split * p->nfft;
    if (need > p->part_cap) {
        if (p->d_part) CUDA_TRY(cudaFree(p->d_part));
        CUDA_TRY(cudaMalloc(&p->d_part, sizeof(float) * need));
        p->part_cap = need;
    }
    switch (p->nfft) {
        case 64: return psd_launch<64>(p, d_x, is_complex, navg, n_split, n_lines, dB, d_out, st);
        case 128: return psd_launch<128>(p, d_x, is_complex, navg, n_split, n_lines, dB, d_out, st);
        case 256: return psd_launch<256>(p, d_x, is_complex, navg, n_split, n_lines, dB, d_out, st);
        case 512: return psd_launch<512>(p, d_x, is_complex, navg, n_split, n_lines, dB, d_out, st);
        case 1024: return psd_launch<1024>(p, d_x, is_complex, navg, n_split, n_lines, dB, d_out, st);
        case 2048: return psd_launch<2048>(p, d_x, is_complex, navg, n_split, n_lines, dB, d_out, st);
        case 4096: return psd_launch<4096>(p, d_x, is_complex, navg, n_split, n_lines, dB, d_out, st);
        case 8192: return psd_launch<8192>(p, d_x, is_complex, navg, n_split, n_lines, dB, d_out, st);
        case 16384: return psd_launch<16384>(p, d_x, is_complex, navg, n_split, n_lines, dB, d_out, st);
    }
    pysdr_set_error("psd_lines: unsupported nfft %d", p->nfft);
    return PYSDR_ERR_ARG;
}

extern "C" int pysdr_psd_lines(pysdr_psd *p, const void *d_x, int64_t n, int is_complex, int32_t navg, int dB, float *d_out,
                               int64_t *n_lines_p, void *stream) {
    if (!p || !d_x || !d_out || navg < 1) { pysdr_set_error("psd_lines: bad arguments"); return PYSDR_ERR_ARG; }
    cudaStream_t st = (cudaStream_t)stream;
    i64 n_frames = 0;
    if (p->sub == 1) {
        n_frames = n < p->chunk ? 0 : 1 + (n - p->chunk) / p->hop;
    } else if (n >= p->chunk) {                                    // frames with start(f) + chunk <= n (offsets ascending, < hop)
        const i64 g = (n - p->chunk) / p->hop;
        n_frames = g * p->sub;
        for (int i = 0; i < p->sub; ++i)
            if (g * p->hop + p->sub_off[i] + p->chunk <= n) n_frames = g * p->sub + i + 1;
    }
    const i64 n_lines = n_frames / navg;
    if (n_lines_p) *n_lines_p = n_lines;
    if (n_lines == 0) return PYSDR_OK;
    // a launch covers at most 65535 lines (grid.y of the finalize kernel); a whole capture may have more (e.g. NFFT 64 lines
    // of a 60 s capture): batches of whole lines, whose first frame starts at a multiple of navg frames.  Sub-step plans keep
    // the single-launch limit (their frame starts are not a plain multiple of the hop).
    const i64 BATCH = 65535;
    if (p->sub != 1 && n_lines > BATCH) { pysdr_set_error("psd_lines: more than 65535 lines per call with sub-step frame starts"); return PYSDR_ERR_CAPACITY; }
    const size_t elem = is_complex ? sizeof(float2) : sizeof(float);
    for (i64 l0 = 0; l0 < n_lines; l0 += BATCH) {
        const i64 nl = n_lines - l0 < BATCH ? n_lines - l0 : BATCH;
        const char *xb = (const char *)d_x + (size_t)(l0 * navg) * (size_t)p->hop * elem;
        int rc = psd_lines_batch(p, xb, is_complex, navg, nl, dB, d_out + (size_t)l0 * p->nfft, st);
        if (rc) return rc;
    }
    return PYSDR_OK;
}

// ---- waterfall (reference Plotting.py) ---------------------------------------------------------------
// wf is [nfft][ncols] row-major like the reference's numpy array.
__global__ void wf_shift_kernel(const float *__restrict__ wf_in, float *__restrict__ wf_out, int nfft, int ncols,
                                const float *__restrict__ line, int npsd, int roll) {
    // out[r][c] = c < ncols-1 ? in[(r+roll) mod nfft][c+1] : (r<npsd ? line[r] : -1e38)
    const int r = blockIdx.x;
    int src = (r + roll) % nfft;
    if (src < 0) src += nfft;
    for (int c = threadIdx.x; c < ncols; c += blockDim.x) {
        float v;
        if (c < ncols - 1) v = wf_in[(size_t)src * ncols + c + 1];
        else v = (r < npsd) ? line[r] : -1e38f;
        wf_out[(size_t)r * ncols + c] = v;
    }
}

__global__ void wf_rowmean_kernel(const float *__restrict__ wf, int nfft, int ncols, int cnt, float *__restrict__ mean) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= nfft) return;
    double sum = 0.0;
    for (int c = ncols - cnt; c < ncols; ++c) sum += (double)wf[(size_t)r * ncols + c];
    mean[r] = (float)(sum / cnt);
}

// median by rank counting (nfft <= 16384: O(n^2/threads) is fine at display rate); also max of wf[0:npsd]
__global__ void wf_median_kernel(const float *__restrict__ mean, int nfft, float *__restrict__ bk) {
    // np.median: average of the two middle order statistics for even n
    __shared__ float lo_s, hi_s;
    const int k_lo = (nfft - 1) / 2, k_hi = nfft / 2;
    for (int i = threadIdx.x; i < nfft; i += blockDim.x) {
        const float v = mean[i];
        int less = 0, eq = 0;
        for (int j = 0; j < nfft; ++j) {
            const float u = mean[j];
            less += (u < v);
            eq += (u == v);
        }
        if (less <= k_lo && k_lo < less + eq) lo_s = v;
        if (less <= k_hi && k_hi < less + eq) hi_s = v;
    }
    __syncthreads();
    if (threadIdx.x == 0) bk[0] = 0.5f * (lo_s + hi_s);
}

__global__ void wf_max_kernel(const float *__restrict__ wf, i64 n, float *__restrict__ mx) {
    __shared__ float sm[32];
    float m = -3.0e38f;
    for (i64 i = threadIdx.x; i < n; i += blockDim.x) m = fmaxf(m, wf[i]);
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < (int)(blockDim.x >> 5); ++w) m = fmaxf(m, sm[w]);
        mx[0] = m;
    }
}

__global__ void wf_image_kernel(const float *__restrict__ wf, i64 n, const float *__restrict__ bk, const float *__restrict__ mx,
                                float pan_dr, float *__restrict__ img) {
    const float b = bk[0];
    const float floor_v = (mx[0] - b) - pan_dr;
    i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    const i64 stride = (i64)gridDim.x * blockDim.x;
    for (; i < n; i += stride) img[i] = fmaxf(wf[i] - b, floor_v);
}

// image -> RGBA through a 256-entry lookup table (reference Plotting.py:139-142 `img.setLookupTable(lut)`; levels are
// the image's own [lo, hi] = [max - PAN_DR, max] after the clip of Plotting.py:618-626).
__global__ void wf_rgba_kernel(const float *__restrict__ img, i64 n, const float *__restrict__ bk, const float *__restrict__ mx,
                               float pan_dr, const uchar4 *__restrict__ lut, uchar4 *__restrict__ rgba) {
    __shared__ uchar4 s_lut[256];
    for (int i = threadIdx.x; i < 256; i += blockDim.x) s_lut[i] = lut[i];
    __syncthreads();
    const float hi = mx[0] - bk[0], lo = hi - pan_dr;
    const float sc = pan_dr > 0.f ? 255.0f / pan_dr : 0.f;
    i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    const i64 stride = (i64)gridDim.x * blockDim.x;
    for (; i < n; i += stride) {
        int k = (int)floorf((img[i] - lo) * sc + 0.5f);
        k = k < 0 ? 0 : (k > 255 ? 255 : k);
        rgba[i] = s_lut[k];
    }
}

// ---- peak picker (reference Plotting.py:594: scipy.signal.find_peaks(PSD2, distance=dist, height=bkgnd+10)) ------------------
// One CTA.  (1) strict local maxima, a plateau reported at the middle of its flat top (scipy _local_maxima_1d);
// (2) height >= min_height; (3) distance: from the highest peak down, a kept peak removes every candidate closer than
// ceil(distance) samples (scipy _select_by_peak_distance; equal heights: the one at the higher index wins, which is what a
// stable ascending argsort walked from the top gives).  Candidates are sorted by (height, index) with a bitonic sort in
// shared memory; the greedy sweep is the only sequential part (one thread, a few hundred candidates at display rate).
#define PK_MAX 4096
__global__ void __launch_bounds__(1024) find_peaks_kernel(const float *__restrict__ x, int n, const float *__restrict__ bk, float height_above_bk,
                                                          float min_height_abs, int use_bk, int distance, int *__restrict__ peaks,
                                                          int *__restrict__ count) {
    __shared__ float s_h[PK_MAX];
    __shared__ int s_i[PK_MAX];
    __shared__ unsigned char s_keep[PK_MAX];
    __shared__ int s_n;
    const int tid = threadIdx.x;
    const float hmin = use_bk ? bk[0] + height_above_bk : min_height_abs;
    if (tid == 0) s_n = 0;
    __syncthreads();
    for (int i = 1 + tid; i < n - 1; i += blockDim.x) {
        const float v = x[i];
        if (!(x[i - 1] < v)) continue;
        int ahead = i + 1;
        while (ahead < n - 1 && x[ahead] == v) ++ahead;
        if (x[ahead] < v) {
            const int mid = (i + ahead - 1) / 2;
            if (v >= hmin) {
                const int k = atomicAdd(&s_n, 1);
                if (k < PK_MAX) { s_h[k] = v; s_i[k] = mid; }
            }
        }
    }
    __syncthreads();
    const int m = s_n < PK_MAX ? s_n : PK_MAX;
    int m2 = 1;
    while (m2 < m) m2 <<= 1;
    for (int k = m + tid; k < m2; k += blockDim.x) { s_h[k] = 3.0e38f; s_i[k] = 0x7fffffff; }      // padding sorts to the end
    __syncthreads();
    // ascending by index first (the candidates were appended in arbitrary order), used for the neighbour walks
    for (int size = 2; size <= m2; size <<= 1)
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            for (int t = tid; t < m2; t += blockDim.x) {
                const int partner = t ^ stride;
                if (partner > t) {
                    const bool up = (t & size) == 0;
                    const bool gt = s_i[t] > s_i[partner];
                    if (gt == up) {
                        const float th = s_h[t]; s_h[t] = s_h[partner]; s_h[partner] = th;
                        const int ti = s_i[t]; s_i[t] = s_i[partner]; s_i[partner] = ti;
                    }
                }
            }
            __syncthreads();
        }
    for (int k = tid; k < m; k += blockDim.x) s_keep[k] = 1;
    __syncthreads();
    if (tid == 0) {
        // greedy by priority: repeatedly take the highest remaining candidate.  m is small; a selection loop keeps the
        // index-sorted order intact (no second sort, no permutation array).
        if (distance > 1) {
            for (int round = 0; round < m; ++round) {
                int best = -1;
                float bh = -3.0e38f;
                for (int k = 0; k < m; ++k)
                    if (s_keep[k] == 1 && s_h[k] >= bh) { bh = s_h[k]; best = k; }      // >= : ties go to the higher index
                if (best < 0) break;
                s_keep[best] = 2;                                                       // kept for good
                for (int k = best - 1; k >= 0 && s_i[best] - s_i[k] < distance; --k) if (s_keep[k] == 1) s_keep[k] = 0;
                for (int k = best + 1; k < m && s_i[k] - s_i[best] < distance; ++k) if (s_keep[k] == 1) s_keep[k] = 0;
            }
        }
        int c = 0;
        for (int k = 0; k < m; ++k)
            if (s_keep[k]) peaks[c++] = s_i[k];
        count[0] = c;
    }
}

extern "C" int pysdr_find_peaks(const float *d_x, int32_t n, const float *d_bkgnd, float height_above_bkgnd, float min_height,
                                float distance, int32_t *d_peaks, int32_t *d_count, void *stream) {
    if (!d_x || !d_peaks || !d_count || n < 3) { pysdr_set_error("find_peaks: bad arguments"); return PYSDR_ERR_ARG; }
    int dist = distance > 1.0f ? (int)ceilf(distance) : 1;
    find_peaks_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(d_x, n, d_bkgnd, height_above_bkgnd, min_height, d_bkgnd ? 1 : 0, dist, d_peaks,
                                                            d_count);
    LAUNCH_CHECK();
    return PYSDR_OK;
}

extern "C" int pysdr_waterfall_rgba(const float *d_img, int64_t n, const float *d_bkgnd, const float *d_scratch, int32_t nfft,
                                    int32_t ncols, float pan_dr, const void *d_lut256, void *d_rgba, void *stream) {
    if (!d_img || !d_bkgnd || !d_scratch || !d_lut256 || !d_rgba || n < 0) { pysdr_set_error("waterfall_rgba: bad arguments"); return PYSDR_ERR_ARG; }
    if (n == 0) return PYSDR_OK;
    const float *mx = d_scratch + (size_t)nfft * ncols + nfft;           // where pysdr_waterfall_push left max(wf)
    wf_rgba_kernel<<<pysdr_sm_count(), 256, 0, (cudaStream_t)stream>>>(d_img, n, d_bkgnd, mx, pan_dr, (const uchar4 *)d_lut256, (uchar4 *)d_rgba);
    LAUNCH_CHECK();
    return PYSDR_OK;
}

extern "C" int pysdr_waterfall_push(float *d_wf, int32_t nfft, int32_t ncols, int32_t cnt, const float *d_line, int32_t npsd,
                                    int32_t roll_bins, float pan_dr, float *d_img, float *d_bkgnd, float *d_scratch, void *stream) {
    // d_scratch: nfft*ncols (shifted copy) + nfft (row means) + 2 floats
    if (!d_wf || !d_line || !d_img || !d_bkgnd || !d_scratch || nfft < 1 || ncols < 2 || cnt < 1 || cnt > ncols || npsd > nfft) {
        pysdr_set_error("waterfall_push: bad arguments");
        return PYSDR_ERR_ARG;
    }
    cudaStream_t st = (cudaStream_t)stream;
    float *tmp = d_scratch, *mean = d_scratch + (size_t)nfft * ncols, *mx = mean + nfft;
    wf_shift_kernel<<<nfft, 128, 0, st>>>(d_wf, tmp, nfft, ncols, d_line, npsd, roll_bins);
    LAUNCH_CHECK();
    CUDA_TRY(cudaMemcpyAsync(d_wf, tmp, sizeof(float) * (size_t)nfft * ncols, cudaMemcpyDeviceToDevice, st));
    wf_rowmean_kernel<<<(nfft + 255) / 256, 256, 0, st>>>(d_wf, nfft, ncols, cnt, mean);
    LAUNCH_CHECK();
    wf_median_kernel<<<1, 1024, 0, st>>>(mean, nfft, d_bkgnd);
    LAUNCH_CHECK();
    wf_max_kernel<<<1, 1024, 0, st>>>(d_wf, (i64)npsd * ncols, mx);
    LAUNCH_CHECK();
    wf_image_kernel<<<pysdr_sm_count(), 256, 0, st>>>(d_wf, (i64)npsd * ncols, d_bkgnd, mx, pan_dr, d_img);
    LAUNCH_CHECK();
    return PYSDR_OK;
}
