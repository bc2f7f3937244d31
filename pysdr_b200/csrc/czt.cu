// czt.cu — K3b: periodogram lines for transform lengths that are NOT a power of two, by Bluestein's chirp-z identity.
//
// The reference's RF panel does exactly that: for chunk sizes above 65536 it sets chunk_size = int(65636/2) = 32818 and
// NFFT = 65636 (reference Plotting.py:370-375) and numpy happily transforms 65636 = 4 * 61 * 269 points.
//
//   X[k] = conj(b[k]) * sum_n (x[n] w[n] conj(b[n])) * b[k-n],   b[m] = exp(+j pi m^2 / NFFT)
//
// so |X[k]|^2 = |(a (*) b)[k]|^2 with a[n] = x[n] w[n] conj(b[n]): one linear convolution, evaluated as a circular one of
// length M = 2^17 >= NFFT + chunk - 1 through a four-step FFT (M = 512 x 256) built from the shared-memory radix-16
// transforms of fft_smem.cuh:
//
//   czt_cols_fwd  : load + window + chirp, 512-point FFTs down 16 columns at a time, twiddle W_M^(n2 k1)  -> T[p][n2]
//   czt_rows      : 256-point FFT along each row, multiply by the chirp spectrum (same layout, 1/M folded in),
//                   inverse 256-point FFT                                                               -> U[p][n2]
//   czt_cols_inv  : twiddle W_M^(-n2 k1), inverse 512-point FFTs down the columns, |y[n]|^2              -> P[f][k]
//   czt_finalize  : mean over the frames of a line, 1/sum(w^2), dB, fftshift
//
// Spectra are only ever touched in the transforms' "position" order (host code permutes the chirp spectrum once), so
// there is no reordering pass.  host side: pysdr_b200/sig_proc.py (class spectrum), tables from numpy in float64.
#include "common.cuh"
#define FFT_PACKED 0                     /* see fft_smem.cuh: this kernel sits at its register limit */
#include "fft_smem.cuh"

#define CZT_N1 512
#define CZT_N2 256
#define CZT_M (CZT_N1 * CZT_N2)
#define CZT_COLS 16                          /* columns / rows per CTA */
#define CZT_THREADS (32 * CZT_COLS)

struct pysdr_czt {
    int chunk, nfft, hop;
    double wsum2;
    float2 *d_wc;          // [chunk]  w[n] * conj(b[n])
    float2 *d_bspec;       // [512][256] chirp spectrum, rows = k1 position, cols = k2 position, times 1/M
    float2 *d_T, *d_U;     // [batch][M]
    float *d_P;            // [batch][nfft]
    int batch;
    i64 launches;
};

extern "C" int pysdr_fft_pos_to_freq(int nfft, int pos) {
    switch (nfft) {
        case 256: return fft_pos_to_freq<256>(pos);
        case 512: return fft_pos_to_freq<512>(pos);
        case 1024: return fft_pos_to_freq<1024>(pos);
        case 2048: return fft_pos_to_freq<2048>(pos);
        case 4096: return fft_pos_to_freq<4096>(pos);
        case 8192: return fft_pos_to_freq<8192>(pos);
        case 16384: return fft_pos_to_freq<16384>(pos);
    }
    return -1;
}

// grid (CZT_N2 / CZT_COLS, frames)
template <bool CPLX>
__global__ void __launch_bounds__(CZT_THREADS) czt_cols_fwd(const void *__restrict__ xv, i64 first_start, int hop, int chunk,
                                                            const float2 *__restrict__ wc, float2 *__restrict__ T) {
    extern __shared__ __align__(16) float2 s[];
    constexpr int PITCH = FFT_SMEM_ELEMS(CZT_N1);
    const int tid = threadIdx.x, c = tid & (CZT_COLS - 1), r0 = tid / CZT_COLS;
    const int n2 = blockIdx.x * CZT_COLS + c;
    const i64 start = first_start + (i64)blockIdx.y * hop;
    for (int i = 0; i < CZT_N1 / (CZT_THREADS / CZT_COLS); ++i) {
        const int n1 = r0 + i * (CZT_THREADS / CZT_COLS);
        const int n = n1 * CZT_N2 + n2;
        float2 v = make_float2(0.f, 0.f);
        if (n < chunk) {
            const float2 w = wc[n];
            if (CPLX) v = cmul(((const float2 *)xv)[start + n], w);
            else { const float xr = ((const float *)xv)[start + n]; v = make_float2(xr * w.x, xr * w.y); }
        }
        s[c * PITCH + FFT_PAD(n1)] = v;
    }
    __syncthreads();
    fft_smem<CZT_N1, false>(s + (tid >> 5) * PITCH, tid & 31);             // warp f owns column f of the tile
    float2 *Tf = T + (size_t)blockIdx.y * CZT_M;
    for (int i = 0; i < CZT_N1 / (CZT_THREADS / CZT_COLS); ++i) {
        const int p = r0 + i * (CZT_THREADS / CZT_COLS);
        const int k1 = fft_pos_to_freq<CZT_N1>(p);
        float sn, cs;
        sincospif(-2.0f * (float)(n2 * k1) / (float)CZT_M, &sn, &cs);               // n2*k1 < 2^17: exact in float
        Tf[(size_t)p * CZT_N2 + n2] = cmul(s[c * PITCH + FFT_PAD(p)], make_float2(cs, sn));
    }
}

// grid (CZT_N1 / CZT_COLS, frames): rows p = blockIdx.x*16 + warp
__global__ void __launch_bounds__(CZT_THREADS) czt_rows(const float2 *__restrict__ T, const float2 *__restrict__ bspec,
                                                        float2 *__restrict__ U) {
    extern __shared__ __align__(16) float2 s[];
    constexpr int PITCH = FFT_SMEM_ELEMS(CZT_N2);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int p = blockIdx.x * CZT_COLS + warp;
    const float2 *row = T + (size_t)blockIdx.y * CZT_M + (size_t)p * CZT_N2;
    float2 *sw = s + warp * PITCH;
#pragma unroll
    for (int i = 0; i < CZT_N2 / 32; ++i) sw[FFT_PAD(lane + 32 * i)] = row[lane + 32 * i];
    __syncthreads();
    fft_smem<CZT_N2, false>(sw, lane);
    const float2 *bp = bspec + (size_t)p * CZT_N2;
#pragma unroll
    for (int i = 0; i < CZT_N2 / 32; ++i) {
        const int q = lane + 32 * i;
        sw[FFT_PAD(q)] = cmul(sw[FFT_PAD(q)], __ldg(bp + q));
    }
    __syncthreads();
    fft_smem<CZT_N2, true>(sw, lane);
    float2 *orow = U + (size_t)blockIdx.y * CZT_M + (size_t)p * CZT_N2;
#pragma unroll
    for (int i = 0; i < CZT_N2 / 32; ++i) orow[lane + 32 * i] = sw[FFT_PAD(lane + 32 * i)];
}

// grid (CZT_N2 / CZT_COLS, frames)
__global__ void __launch_bounds__(CZT_THREADS) czt_cols_inv(const float2 *__restrict__ U, int nfft, float *__restrict__ P) {
    extern __shared__ __align__(16) float2 s[];
    constexpr int PITCH = FFT_SMEM_ELEMS(CZT_N1);
    const int tid = threadIdx.x, c = tid & (CZT_COLS - 1), r0 = tid / CZT_COLS;
    const int n2 = blockIdx.x * CZT_COLS + c;
    const float2 *Uf = U + (size_t)blockIdx.y * CZT_M;
    for (int i = 0; i < CZT_N1 / (CZT_THREADS / CZT_COLS); ++i) {
        const int p = r0 + i * (CZT_THREADS / CZT_COLS);
        const int k1 = fft_pos_to_freq<CZT_N1>(p);
        float sn, cs;
        sincospif(2.0f * (float)(n2 * k1) / (float)CZT_M, &sn, &cs);
        s[c * PITCH + FFT_PAD(p)] = cmul(Uf[(size_t)p * CZT_N2 + n2], make_float2(cs, sn));
    }
    __syncthreads();
    fft_smem<CZT_N1, true>(s + (tid >> 5) * PITCH, tid & 31);
    float *Pf = P + (size_t)blockIdx.y * nfft;
    for (int i = 0; i < CZT_N1 / (CZT_THREADS / CZT_COLS); ++i) {
        const int n1 = r0 + i * (CZT_THREADS / CZT_COLS);
        const int k = n1 * CZT_N2 + n2;
        if (k < nfft) {
            const float2 y = s[c * PITCH + FFT_PAD(n1)];
            Pf[k] = y.x * y.x + y.y * y.y;                                          // |conj(b[k]) y|^2 = |y|^2
        }
    }
}

// out[line][fftshift(k)] = dB( mean_f P[line*navg + f][k] / sum(w^2) );  grid (ceil(nfft/256), lines)
__global__ void czt_finalize(const float *__restrict__ P, int nfft, int navg, float scale, int dB, float *__restrict__ out) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= nfft) return;
    const float *q = P + (size_t)blockIdx.y * navg * nfft + k;
    float sum = 0.f;
    for (int f = 0; f < navg; ++f) sum += q[(size_t)f * nfft];
    float v = sum * scale;
    if (dB) v = 10.f * log10f(fmaxf(v, 1.0e-30f));
    const int half = nfft / 2;                                                       // np.fft.fftshift: k -> (k + n//2) % n
    out[(size_t)blockIdx.y * nfft + (k + half) % nfft] = v;
}

extern "C" int pysdr_czt_create(int32_t chunk, int32_t nfft, int32_t hop, const float *window, const float *wc_host,
                                const float *bspec_host, pysdr_czt **out) {
    if (!out || !window || !wc_host || !bspec_host || chunk < 1 || nfft < chunk || hop < 1 || (i64)nfft + chunk - 1 > CZT_M) {
        pysdr_set_error("czt_create: need 1 <= chunk <= nfft and nfft + chunk - 1 <= %d (got chunk=%d nfft=%d hop=%d)", CZT_M, chunk,
                        nfft, hop);
        return PYSDR_ERR_ARG;
    }
    pysdr_czt *p = new pysdr_czt();
    p->chunk = chunk; p->nfft = nfft; p->hop = hop;
    p->wsum2 = 0.0;
    for (int i = 0; i < chunk; ++i) p->wsum2 += (double)window[i] * (double)window[i];
    p->d_T = p->d_U = nullptr; p->d_P = nullptr; p->batch = 0; p->launches = 0;
    p->d_wc = p->d_bspec = nullptr;
    if (cudaMalloc(&p->d_wc, sizeof(float2) * chunk) != cudaSuccess || cudaMalloc(&p->d_bspec, sizeof(float2) * CZT_M) != cudaSuccess) {
        pysdr_set_error("czt_create: cudaMalloc failed");
        cudaFree(p->d_wc);
        delete p;
        return PYSDR_ERR_CUDA;
    }
    CUDA_TRY(cudaMemcpy(p->d_wc, wc_host, sizeof(float2) * chunk, cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMemcpy(p->d_bspec, bspec_host, sizeof(float2) * CZT_M, cudaMemcpyHostToDevice));
    *out = p;
    return PYSDR_OK;
}

extern "C" int pysdr_czt_destroy(pysdr_czt *p) {
    if (!p) return PYSDR_OK;
    cudaFree(p->d_wc); cudaFree(p->d_bspec); cudaFree(p->d_T); cudaFree(p->d_U); cudaFree(p->d_P);
    delete p;
    return PYSDR_OK;
}

extern "C" int64_t pysdr_czt_launch_count(const pysdr_czt *p) { return p ? p->launches : -1; }

extern "C" int pysdr_czt_lines(pysdr_czt *p, const void *d_x, int64_t n, int is_complex, int32_t navg, int dB, float *d_out,
                               int64_t *n_lines_p, void *stream) {
    if (!p || !d_x || !d_out || navg < 1) { pysdr_set_error("czt_lines: bad arguments"); return PYSDR_ERR_ARG; }
    cudaStream_t st = (cudaStream_t)stream;
    const i64 n_frames = n < p->chunk ? 0 : 1 + (n - p->chunk) / p->hop;
    const i64 n_lines = n_frames / navg;
    if (n_lines_p) *n_lines_p = n_lines;
    if (n_lines == 0) return PYSDR_OK;
    // frames are processed in batches of whole lines; the work buffers hold one batch
    i64 lines_per_batch = 64 / navg;
    if (lines_per_batch < 1) lines_per_batch = 1;
    if (lines_per_batch > n_lines) lines_per_batch = n_lines;
    const int need = (int)(lines_per_batch * navg);
    if (need > 4096) { pysdr_set_error("czt_lines: navg above 4096 frames per line is not supported"); return PYSDR_ERR_CAPACITY; }
    if (need > p->batch) {
        cudaFree(p->d_T); cudaFree(p->d_U); cudaFree(p->d_P);
        p->d_T = p->d_U = nullptr; p->d_P = nullptr; p->batch = 0;
        CUDA_TRY(cudaMalloc(&p->d_T, sizeof(float2) * (size_t)need * CZT_M));
        CUDA_TRY(cudaMalloc(&p->d_U, sizeof(float2) * (size_t)need * CZT_M));
        CUDA_TRY(cudaMalloc(&p->d_P, sizeof(float) * (size_t)need * p->nfft));
        p->batch = need;
    }
    const size_t smem_c = sizeof(float2) * CZT_COLS * FFT_SMEM_ELEMS(CZT_N1);
    const size_t smem_r = sizeof(float2) * CZT_COLS * FFT_SMEM_ELEMS(CZT_N2);
    CUDA_TRY(cudaFuncSetAttribute(czt_cols_fwd<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_c));
    CUDA_TRY(cudaFuncSetAttribute(czt_cols_fwd<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_c));
    CUDA_TRY(cudaFuncSetAttribute(czt_cols_inv, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_c));
    const float scale = (float)(1.0 / ((double)navg * p->wsum2));
    for (i64 l0 = 0; l0 < n_lines; l0 += lines_per_batch) {
        const i64 nl = (n_lines - l0 < lines_per_batch) ? n_lines - l0 : lines_per_batch;
        const unsigned nf = (unsigned)(nl * navg);
        const i64 first = l0 * navg * p->hop;
        dim3 gc(CZT_N2 / CZT_COLS, nf), gr(CZT_N1 / CZT_COLS, nf);
        if (is_complex) czt_cols_fwd<true><<<gc, CZT_THREADS, smem_c, st>>>(d_x, first, p->hop, p->chunk, p->d_wc, p->d_T);
        else czt_cols_fwd<false><<<gc, CZT_THREADS, smem_c, st>>>(d_x, first, p->hop, p->chunk, p->d_wc, p->d_T);
        LAUNCH_CHECK();
        czt_rows<<<gr, CZT_THREADS, smem_r, st>>>(p->d_T, p->d_bspec, p->d_U);
        LAUNCH_CHECK();
        czt_cols_inv<<<gc, CZT_THREADS, smem_c, st>>>(p->d_U, p->nfft, p->d_P);
        LAUNCH_CHECK();
        dim3 gf((unsigned)((p->nfft + 255) / 256), (unsigned)nl);
        czt_finalize<<<gf, 256, 0, st>>>(p->d_P, p->nfft, navg, scale, dB, d_out + (size_t)l0 * p->nfft);
        LAUNCH_CHECK();
        p->launches += 4;
    }
    return PYSDR_OK;
}
