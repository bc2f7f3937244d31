// lfilter.cu — scipy.signal.lfilter(b, a, x, zi=z) with carried state as a block-parallel linear scan.
//
// Reference: sigs/iir.py:90-105 (y1,z1 = lfilter(b,a,x1,zi=zi); y2,z2 = lfilter(b,a,x2,zi=z1) == one shot),
// sigs/squelch.m:125-128 (one-pole envelope smoother), sigs/agc.m:6-12.
//
// Transposed direct form II in float64 (what scipy evaluates):
//     y[n]   = b0 x[n] + z0[n-1]
//     z_i[n] = b_{i+1} x[n] + z_{i+1}[n-1] - a_{i+1} y[n]          (z_K = 0)
// The recursion is linear in (z, x), so a sequence of B samples maps z -> Phi^B z + s, with s the state
// reached from z = 0.  A hierarchical scan (float64 throughout):
//   P1   every block of B = 32 samples runs the recursion from zero state, keeps s_b                     (parallel)
//   UP   groups of 32 units chain z <- Phi^unit z + s_u from zero state, keep the group's s; repeated with Phi^(32 unit)
//        until at most 32 units are left                                                                  (parallel per level)
//   TOP  one warp per channel chains the remaining units from the carried state zi                        (<= 32 steps)
//   DOWN every group re-chains its units from its true entry state, level by level, down to the blocks    (parallel per level)
//   P3   every block re-runs the recursion from its true entry state and writes y                         (parallel)
// (r01 chained all 512-sample blocks serially in one warp and gave every thread 512 dependent float64 steps: 242 us for a
// one-pole filter over 10 s of 48 kHz audio; the de-emphasis of the stereo FM decoder spent more time there than the RF-rate
// stages.)  Phi^32 is formed on the host by running the recursion from the unit states, the higher powers by squaring.
#include "common.cuh"

#define LF_B 32             /* samples per scan block */
#define LF_G 32             /* units per group in the up / down passes */
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

// Chain over units: grid (n_groups, n_ch), one warp each.  Group g covers units [g*gsz, min((g+1)*gsz, n_units)).
//   z <- phi z + s[u]  from  z = z_in[g] (or 0);  z_entry[u] (optional) = state ENTERING unit u;  s_out[g] (optional) =
//   state leaving the group.  Lane i owns row i of phi (K <= 32).
template <int K>
__global__ void __launch_bounds__(32)
lf_chain_kernel(const double *__restrict__ phi /* [K][K] */, const double *__restrict__ s /* [n_ch][n_units][K] */, i64 n_units,
                i64 gsz, const double *__restrict__ z_in /* [n_ch][n_groups][K] or null */, i64 z_in_stride,
                double *__restrict__ z_entry /* [n_ch][n_units][K] or null */, double *__restrict__ s_out /* [n_ch][n_groups][K] or null */) {
    const int ch = blockIdx.y;
    const i64 g = blockIdx.x, n_groups = gridDim.x;
    const int lane = threadIdx.x;
    double row[K];
#pragma unroll
    for (int j = 0; j < K; ++j) row[j] = (lane < K) ? phi[lane * K + j] : 0.0;
    double zl = (z_in && lane < K) ? z_in[((size_t)ch * z_in_stride + g) * K + lane] : 0.0;
    const i64 u0 = g * gsz, u1 = (u0 + gsz < n_units) ? u0 + gsz : n_units;
    const double *sp = s + ((size_t)ch * n_units + u0) * K;
    double nxt = (lane < K && u0 < u1) ? sp[lane] : 0.0;                 // one unit of look-ahead on the s loads
    for (i64 u = u0; u < u1; ++u) {
        if (z_entry && lane < K) z_entry[((size_t)ch * n_units + u) * K + lane] = zl;
        double acc = nxt;
        if (lane < K && u + 1 < u1) nxt = sp[(u + 1 - u0) * K + lane];
#pragma unroll
        for (int j = 0; j < K; ++j) {
            const double zj = __shfl_sync(0xffffffffu, zl, j);
            acc = fma(row[j], zj, acc);
        }
        zl = acc;
    }
    if (s_out && lane < K) s_out[((size_t)ch * n_groups + g) * K + lane] = zl;
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

    // levels: level 0 = blocks; level l+1 = groups of LF_G level-l units, until <= LF_G units remain
    std::vector<i64> cnt;
    cnt.push_back(nblk);
    while (cnt.back() > LF_G) cnt.push_back((cnt.back() + LF_G - 1) / LF_G);
    const int n_lev = (int)cnt.size();
    // transition matrices per level: phi[0] = Phi^B, phi[l+1] = phi[l]^LF_G (5 squarings, long double)
    std::vector<double> phis((size_t)n_lev * K * K);
    {
        std::vector<long double> m((size_t)K * K), t((size_t)K * K);
        for (size_t i = 0; i < (size_t)K * K; ++i) { m[i] = phiB[i]; phis[i] = phiB[i]; }
        for (int l = 1; l < n_lev; ++l) {
            for (int sq = 0; sq < 5; ++sq) {                              // LF_G = 32 = 2^5
                for (int i = 0; i < K; ++i)
                    for (int j = 0; j < K; ++j) {
                        long double acc = 0.0L;
                        for (int k = 0; k < K; ++k) acc += m[(size_t)i * K + k] * m[(size_t)k * K + j];
                        t[(size_t)i * K + j] = acc;
                    }
                m.swap(t);
            }
            for (size_t i = 0; i < (size_t)K * K; ++i) {
                phis[(size_t)l * K * K + i] = (double)m[i];
                if (!(fabs((double)m[i]) <= 1.0e3)) {                     // growing powers: fall back to the sequential form
                    lf_seq_kernel<K><<<n_ch, 1, 0, st>>>(c, d_x, d_y, n, stride, d_zi);
                    LAUNCH_CHECK();
                    return PYSDR_OK;
                }
            }
        }
    }
    static_assert(LF_G == 32, "the level matrices are formed by five squarings");
    // workspace: [phis][per level: s (n_ch * cnt[l] * K) and z_entry (same)]
    size_t tot = (size_t)n_lev * K * K;
    std::vector<size_t> off_s(n_lev), off_z(n_lev);
    for (int l = 0; l < n_lev; ++l) {
        off_s[l] = tot; tot += (size_t)n_ch * cnt[l] * K;
        off_z[l] = tot; tot += (size_t)n_ch * cnt[l] * K;
    }
    double *d_ws = nullptr;
    CUDA_TRY(cudaMallocAsync(&d_ws, sizeof(double) * tot, st));
    CUDA_TRY(cudaMemcpyAsync(d_ws, phis.data(), sizeof(double) * phis.size(), cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaStreamSynchronize(st));          // phis is host-stack owned
    dim3 grid((unsigned)((nblk + LF_ROWS - 1) / LF_ROWS), (unsigned)n_ch);
    lf_block_kernel<K, false><<<grid, LF_ROWS, 0, st>>>(c, d_x, nullptr, n, stride, nullptr, d_ws + off_s[0], nblk);
    LAUNCH_CHECK();
    for (int l = 0; l + 1 < n_lev; ++l) {         // UP: zero-state exit of every group of level-l units
        dim3 g((unsigned)cnt[l + 1], (unsigned)n_ch);
        lf_chain_kernel<K><<<g, 32, 0, st>>>(d_ws + (size_t)l * K * K, d_ws + off_s[l], cnt[l], LF_G, nullptr, 0, nullptr,
                                              d_ws + off_s[l + 1]);
        LAUNCH_CHECK();
    }
    {                                              // TOP: the <= LF_G remaining units from the carried state
        const int l = n_lev - 1;
        dim3 g(1, (unsigned)n_ch);
        lf_chain_kernel<K><<<g, 32, 0, st>>>(d_ws + (size_t)l * K * K, d_ws + off_s[l], cnt[l], cnt[l], d_zi, 1, d_ws + off_z[l], nullptr);
        LAUNCH_CHECK();
    }
    for (int l = n_lev - 2; l >= 0; --l) {         // DOWN: entry states of the level-l units from their group's entry state
        dim3 g((unsigned)cnt[l + 1], (unsigned)n_ch);
        lf_chain_kernel<K><<<g, 32, 0, st>>>(d_ws + (size_t)l * K * K, d_ws + off_s[l], cnt[l], LF_G, d_ws + off_z[l + 1], cnt[l + 1],
                                              d_ws + off_z[l], nullptr);
        LAUNCH_CHECK();
    }
    lf_block_kernel<K, true><<<grid, LF_ROWS, 0, st>>>(c, d_x, d_y, n, stride, d_ws + off_z[0], nullptr, nblk);
    LAUNCH_CHECK();
    // carried state after all n samples: the last block re-run from its entry state (exact also for a partial block)
    lf_tail_kernel<K><<<n_ch, 1, 0, st>>>(c, d_x, n, stride, d_ws + off_z[0], d_zi, nblk);
    LAUNCH_CHECK();
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
