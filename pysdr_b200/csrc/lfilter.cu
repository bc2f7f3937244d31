// lfilter.cu — scipy.signal.lfilter(b, a, x, zi=z) with carried state as a block-parallel linear scan.
//
// Reference: sigs/iir.py:90-105 (y1,z1 = lfilter(b,a,x1,zi=zi); y2,z2 = lfilter(b,a,x2,zi=z1) == one shot),
// sigs/squelch.m:125-128 (one-pole envelope smoother), sigs/agc.m:6-12.
//
// Transposed direct form II in float64 (what scipy evaluates):
//     y[n]   = b0 x[n] + z0[n-1]
//     z_i[n] = b_{i+1} x[n] + z_{i+1}[n-1] - a_{i+1} y[n]          (z_K = 0)
// The recursion is linear in (z, x), so a sequence of B samples maps z -> Phi^B z + s, with s the state
// reached from z = 0.  Three passes:
//   P1  every block (B samples) runs the recursion from zero state, keeps s_b           (parallel)
//   P2  one warp per channel chains z_{b+1} = Phi^B z_b + s_b over the blocks           (serial, tiny)
//   P3  every block re-runs the recursion from its true entry state and writes y        (parallel)
// Phi^B is formed on the host in float64 by repeated squaring of the KxK companion matrix.
#include "common.cuh"

#define LF_B 512            /* samples per scan block */
#define LF_ROWS 32          /* blocks handled per CTA (one per thread) */

// One DF2T step in scipy's own operation order without FMA contraction (scipy/signal/_lfilter.c.in:
//   y = z0 + x*b0 ; z_i = (z_{i+1} + x*b_{i+1}) - y*a_{i+1} ; z_last = x*b_last - y*a_last).
template <int K>
__device__ __forceinline__ double lf_step(const double *b, const double *a, double *z, double xv) {
    const double yv = __dadd_rn(z[0], __dmul_rn(xv, b[0]));
#pragma unroll
    for (int i = 0; i < K - 1; ++i)
        z[i] = __dsub_rn(__dadd_rn(z[i + 1], __dmul_rn(xv, b[i + 1])), __dmul_rn(yv, a[i + 1]));
    z[K - 1] = __dsub_rn(__dmul_rn(xv, b[K]), __dmul_rn(yv, a[K]));
    return yv;
}

template <int K>
struct LfCoef {
    double b[K + 1];
    double a[K + 1];       // a[0] == 1 after normalisation
};

// One CTA = 32 threads = 32 consecutive blocks of one channel; tiles are staged through shared memory so
// global accesses stay coalesced although each thread walks its own block serially.
template <int K, bool WRITE_Y>
__global__ void __launch_bounds__(LF_ROWS)
lf_block_kernel(LfCoef<K> c, const float *__restrict__ x, float *__restrict__ y, i64 n, i64 stride,
                const double *__restrict__ z_in /* [n_ch][nblk][K] or null (zero) */, double *__restrict__ s_out /* [n_ch][nblk][K] */,
                i64 nblk) {
    __shared__ float tile[LF_ROWS][33];
    const int t = threadIdx.x;
    const int ch = blockIdx.y;
    const i64 b0 = (i64)blockIdx.x * LF_ROWS;
    const float *xc = x + (size_t)ch * stride;
    float *yc = WRITE_Y ? y + (size_t)ch * stride : nullptr;
    const i64 blk = b0 + t;
    double z[K];
#pragma unroll
    for (int i = 0; i < K; ++i) z[i] = (z_in && blk < nblk) ? z_in[((size_t)ch * nblk + blk) * K + i] : 0.0;

    for (int j0 = 0; j0 < LF_B; j0 += 32) {
#pragma unroll 4
        for (int r = 0; r < LF_ROWS; ++r) {
            const i64 idx = (b0 + r) * LF_B + j0 + t;
            tile[r][t] = (idx < n) ? xc[idx] : 0.f;
        }
        __syncwarp();
#pragma unroll 4
        for (int j = 0; j < 32; ++j) {
            const double yv = lf_step<K>(c.b, c.a, z, (double)tile[t][j]);
            if (WRITE_Y) tile[t][j] = (float)yv;
        }
        __syncwarp();
        if (WRITE_Y) {
#pragma unroll 4
            for (int r = 0; r < LF_ROWS; ++r) {
                const i64 idx = (b0 + r) * LF_B + j0 + t;
                if (idx < n) yc[idx] = tile[r][t];
            }
            __syncwarp();
        }
    }
    if (!WRITE_Y && blk < nblk) {
#pragma unroll
        for (int i = 0; i < K; ++i) s_out[((size_t)ch * nblk + blk) * K + i] = z[i];
    }
}

// P2: z_entry[b+1] = PhiB z_entry[b] + s[b]; lane i owns row i.  The last partial block is handled by P3
// itself for y, and for the carried state by stepping the tail exactly (host passes n_tail) via phi_tail.
template <int K>
__global__ void lf_chain_kernel(const double *__restrict__ phiB /* [K][K] */, const double *__restrict__ s /* [n_ch][nblk][K] */,
                                double *__restrict__ z_entry /* [n_ch][nblk][K] */, double *__restrict__ zi /* [n_ch][K] in: entry of block 0 */,
                                i64 nblk) {
    const int ch = blockIdx.x;
    const int lane = threadIdx.x;
    double row[K];
#pragma unroll
    for (int j = 0; j < K; ++j) row[j] = (lane < K) ? phiB[lane * K + j] : 0.0;
    double zl = (lane < K) ? zi[(size_t)ch * K + lane] : 0.0;
    for (i64 b = 0; b < nblk; ++b) {
        if (lane < K) z_entry[((size_t)ch * nblk + b) * K + lane] = zl;
        double acc = (lane < K) ? s[((size_t)ch * nblk + b) * K + lane] : 0.0;
#pragma unroll
        for (int j = 0; j < K; ++j) {
            const double zj = __shfl_sync(0xffffffffu, zl, j);
            acc = fma(row[j], zj, acc);
        }
        zl = acc;
    }
    // zl is now the state after nblk FULL blocks; only meaningful when n is a multiple of LF_B (see host)
    if (lane < K) zi[(size_t)ch * K + lane] = zl;
}

// exact tail: final carried state when n is not a multiple of LF_B (one thread per channel re-runs the
// last partial block from its entry state)
template <int K>
__global__ void lf_tail_kernel(LfCoef<K> c, const float *__restrict__ x, i64 n, i64 stride,
                               const double *__restrict__ z_entry, double *__restrict__ zi, i64 nblk) {
    const int ch = blockIdx.x;
    const i64 b = nblk - 1;
    double z[K];
#pragma unroll
    for (int i = 0; i < K; ++i) z[i] = z_entry[((size_t)ch * nblk + b) * K + i];
    const float *xc = x + (size_t)ch * stride;
    for (i64 idx = b * LF_B; idx < n; ++idx) lf_step<K>(c.b, c.a, z, (double)xc[idx]);
#pragma unroll
    for (int i = 0; i < K; ++i) zi[(size_t)ch * K + i] = z[i];
}

// Sequential evaluation (one thread per channel): used when chaining blocks would be ill-conditioned.
// Same operation order as scipy, so it tracks scipy's own float64 rounding.
template <int K>
__global__ void lf_seq_kernel(LfCoef<K> c, const float *__restrict__ x, float *__restrict__ y, i64 n, i64 stride,
                              double *__restrict__ zi) {
    const int ch = blockIdx.x;
    double z[K];
#pragma unroll
    for (int i = 0; i < K; ++i) z[i] = zi[(size_t)ch * K + i];
    const float *xc = x + (size_t)ch * stride;
    float *yc = y + (size_t)ch * stride;
    for (i64 i = 0; i < n; ++i) yc[i] = (float)lf_step<K>(c.b, c.a, z, (double)xc[i]);
#pragma unroll
    for (int i = 0; i < K; ++i) zi[(size_t)ch * K + i] = z[i];
}

template <int K>
static int lfilter_run(const double *bn, const double *an, const float *d_x, float *d_y, i64 n, int n_ch, i64 stride,
                       double *d_zi, cudaStream_t st, int force_mode) {
    LfCoef<K> c;
    for (int i = 0; i <= K; ++i) { c.b[i] = bn[i]; c.a[i] = an[i]; }
    const i64 nblk = (n + LF_B - 1) / LF_B;
    // Phi^B column j = state reached after B zero-input steps from the unit state e_j (the recursion itself is
    // the numerically stable way to form it; powering the companion matrix is not).
    std::vector<double> phiB((size_t)K * K);
    double pmax = 0.0;
    for (int j = 0; j < K; ++j) {
        long double z[K];
        for (int i = 0; i < K; ++i) z[i] = (i == j) ? 1.0L : 0.0L;
        for (int stp = 0; stp < LF_B; ++stp) {
            const long double yv = z[0];
            for (int i = 0; i < K - 1; ++i) z[i] = z[i + 1] - (long double)an[i + 1] * yv;
            z[K - 1] = -(long double)an[K] * yv;
        }
        for (int i = 0; i < K; ++i) {
            phiB[(size_t)i * K + j] = (double)z[i];
            const double m = fabs((double)z[i]);
            if (!(m <= pmax)) pmax = m;             // also catches NaN
        }
    }
    // Conditioning gate: chaining multiplies state round-off by |Phi^B|.  Narrow-band high-order direct forms
    // (e.g. reference sigs/iir.py ellip-7 @200/8000 Hz, cheby2-15 band) exceed it and run sequentially.
    if (force_mode == 2 || (force_mode == 0 && !(pmax <= 1.0e3))) {
        lf_seq_kernel<K><<<n_ch, 1, 0, st>>>(c, d_x, d_y, n, stride, d_zi);
        LAUNCH_CHECK();
        return PYSDR_OK;
    }

    double *d_ws = nullptr;      // [phiB K*K][s n_ch*nblk*K][z_entry n_ch*nblk*K]
    const size_t per = (size_t)n_ch * nblk * K;
    CUDA_TRY(cudaMallocAsync(&d_ws, sizeof(double) * ((size_t)K * K + 2 * per), st));
    double *d_phi = d_ws, *d_s = d_ws + (size_t)K * K, *d_ze = d_s + per;
    CUDA_TRY(cudaMemcpyAsync(d_phi, phiB.data(), sizeof(double) * K * K, cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaStreamSynchronize(st));          // phiB is host-stack owned
    dim3 grid((unsigned)((nblk + LF_ROWS - 1) / LF_ROWS), (unsigned)n_ch);
    lf_block_kernel<K, false><<<grid, LF_ROWS, 0, st>>>(c, d_x, nullptr, n, stride, nullptr, d_s, nblk);
    LAUNCH_CHECK();
    lf_chain_kernel<K><<<n_ch, 32, 0, st>>>(d_phi, d_s, d_ze, d_zi, nblk);
    LAUNCH_CHECK();
    lf_block_kernel<K, true><<<grid, LF_ROWS, 0, st>>>(c, d_x, d_y, n, stride, d_ze, nullptr, nblk);
    LAUNCH_CHECK();
    if (n % LF_B != 0) {
        lf_tail_kernel<K><<<n_ch, 1, 0, st>>>(c, d_x, n, stride, d_ze, d_zi, nblk);
        LAUNCH_CHECK();
    }
    CUDA_TRY(cudaFreeAsync(d_ws, st));
    return PYSDR_OK;
}

static int g_lf_mode = 0;       // 0 auto, 1 force block scan, 2 force sequential
extern "C" int pysdr_lfilter_set_mode(int mode) { g_lf_mode = mode; return PYSDR_OK; }

// ---- elementwise helpers of the squelch detector (reference sigs/squelch.m:125-128, 141) ------------------
__global__ void abs_f32_kernel(const float *__restrict__ x, float *__restrict__ y, i64 n) {
    i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    const i64 stride = (i64)gridDim.x * blockDim.x;
    for (; i < n; i += stride) y[i] = fabsf(x[i]);
}
__global__ void ratio_f32_kernel(const float *__restrict__ a, const float *__restrict__ b, float *__restrict__ r, float floor_v,
                                 i64 n) {
    i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    const i64 stride = (i64)gridDim.x * blockDim.x;
    for (; i < n; i += stride) r[i] = a[i] / fmaxf(b[i], floor_v);
}
extern "C" int pysdr_abs_f32(const float *d_x, float *d_y, int64_t n, void *stream) {
    if (!d_x || !d_y || n < 0) { pysdr_set_error("abs_f32: bad arguments"); return PYSDR_ERR_ARG; }
    if (n == 0) return PYSDR_OK;
    i64 blocks = (n + 255) / 256;
    if (blocks > (i64)pysdr_sm_count() * 8) blocks = (i64)pysdr_sm_count() * 8;
    abs_f32_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(d_x, d_y, n);
    LAUNCH_CHECK();
    return PYSDR_OK;
}
extern "C" int pysdr_ratio_f32(const float *d_a, const float *d_b, float *d_r, float floor_v, int64_t n, void *stream) {
    if (!d_a || !d_b || !d_r || n < 0) { pysdr_set_error("ratio_f32: bad arguments"); return PYSDR_ERR_ARG; }
    if (n == 0) return PYSDR_OK;
    i64 blocks = (n + 255) / 256;
    if (blocks > (i64)pysdr_sm_count() * 8) blocks = (i64)pysdr_sm_count() * 8;
    ratio_f32_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(d_a, d_b, d_r, floor_v, n);
    LAUNCH_CHECK();
    return PYSDR_OK;
}

extern "C" int pysdr_lfilter(const double *b, int nb, const double *a, int na, const float *d_x, float *d_y,
                             int64_t n, int n_ch, int64_t stride, double *d_zi, void *stream) {
    if (!b || !a || nb < 1 || na < 1 || a[0] == 0.0 || !d_x || !d_y || n < 0 || n_ch < 1 || !d_zi) {
        pysdr_set_error("lfilter: bad arguments");
        return PYSDR_ERR_ARG;
    }
    if (n == 0) return PYSDR_OK;
    const int order = (na > nb ? na : nb) - 1;
    if (order < 1 || order > 32) {
        pysdr_set_error("lfilter: order %d unsupported (1..32)", order);
        return PYSDR_ERR_ARG;
    }
    int K = 1;
    while (K < order) K *= 2;
    double bn[33], an[33];
    for (int i = 0; i <= 32; ++i) { bn[i] = 0.0; an[i] = 0.0; }
    for (int i = 0; i < nb; ++i) bn[i] = b[i] / a[0];
    for (int i = 0; i < na; ++i) an[i] = a[i] / a[0];
    cudaStream_t st = (cudaStream_t)stream;
    // d_zi holds `order` doubles per channel; the kernels use K >= order: stage through a padded buffer
    double *d_zp = nullptr;
    CUDA_TRY(cudaMallocAsync(&d_zp, sizeof(double) * (size_t)n_ch * K, st));
    CUDA_TRY(cudaMemsetAsync(d_zp, 0, sizeof(double) * (size_t)n_ch * K, st));
    CUDA_TRY(cudaMemcpy2DAsync(d_zp, sizeof(double) * K, d_zi, sizeof(double) * order, sizeof(double) * order, n_ch,
                               cudaMemcpyDeviceToDevice, st));
    int rc;
    switch (K) {
        case 1: rc = lfilter_run<1>(bn, an, d_x, d_y, n, n_ch, stride, d_zp, st, g_lf_mode); break;
        case 2: rc = lfilter_run<2>(bn, an, d_x, d_y, n, n_ch, stride, d_zp, st, g_lf_mode); break;
        case 4: rc = lfilter_run<4>(bn, an, d_x, d_y, n, n_ch, stride, d_zp, st, g_lf_mode); break;
        case 8: rc = lfilter_run<8>(bn, an, d_x, d_y, n, n_ch, stride, d_zp, st, g_lf_mode); break;
        case 16: rc = lfilter_run<16>(bn, an, d_x, d_y, n, n_ch, stride, d_zp, st, g_lf_mode); break;
        default: rc = lfilter_run<32>(bn, an, d_x, d_y, n, n_ch, stride, d_zp, st, g_lf_mode); break;
    }
    if (rc) { cudaFreeAsync(d_zp, st); return rc; }
    CUDA_TRY(cudaMemcpy2DAsync(d_zi, sizeof(double) * order, d_zp, sizeof(double) * K, sizeof(double) * order, n_ch,
                               cudaMemcpyDeviceToDevice, st));
    CUDA_TRY(cudaFreeAsync(d_zp, st));
    return PYSDR_OK;
}
