// K1 (generic variant): fused NCO-mix + polyphase FIR decimate, any UP/DOWN/taps/n_rx.
//
// Replaces rx.lo.quad_mixer + rx.dec resampling inside dsp.Receiver.demod_data (reference
// receiver.py:235; NCO receiver.py:822; resampler taps receiver.py:866 `dec.filter_bank`).
//
//   y_r[m] = e^{-j th_r(n_m)} * sum_j G_r[p_m][j] * x[n_m - j],   G_r[p][j] = h_r[p + j*UP] e^{+j 2pi f_r j}
//   p_m = (m*DOWN) % UP,  n_m = (m*DOWN) / UP                                    (bit-exact indexing)
//
// The per-sample LO is folded into per-receiver complex taps (host, float64) so the only sin/cos is
// one per OUTPUT sample, evaluated from the exact u64 phase accumulator.  One warp per output time m,
// lanes stride the taps, all receivers share the x loads.  This variant is the correctness baseline
// for every geometry; k1_fast.cu is the tuned tap-stationary kernel for the common ones.
#include "common.cuh"

#define K1G_RX 8             /* receivers per launch (register-resident accumulators); banks with more launch in groups */
__global__ void __launch_bounds__(256) k1_generic_kernel(const K1Args a, const int rx0) {
    const int lane = threadIdx.x & 31;
    const i64 warp = ((i64)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const i64 nwarps = ((i64)gridDim.x * blockDim.x) >> 5;
    const size_t rx_pitch = (size_t)a.up * a.lp_pad;

    for (i64 i = warp; i < a.n_out; i += nwarps) {
        const i64 m = a.m0 + i;
        const i64 t = m * a.down;
        const i64 nm = t / a.up;
        const int ph = (int)(t - nm * a.up);
        const i64 r0 = nm - a.n0;                       // newest input of this output, relative to x[0]
        float sr[K1G_RX], si[K1G_RX];
#pragma unroll
        for (int r = 0; r < K1G_RX; ++r) { sr[r] = 0.f; si[r] = 0.f; }
        const float2 *gp = a.g + (size_t)ph * a.lp_pad;
        for (int j = lane; j < a.lp; j += 32) {
            const i64 idx = r0 - j;
            float2 xv;
            if (idx >= 0) xv = __ldg(a.x + idx);
            else if (idx >= -(i64)a.need && a.hist) xv = a.hist[a.need + idx];
            else xv = make_float2(0.f, 0.f);
#pragma unroll
            for (int r = 0; r < K1G_RX; ++r) {
                if (rx0 + r < a.n_rx) {
                    const float2 gv = __ldg(gp + (rx0 + r) * rx_pitch + j);
                    sr[r] = fmaf(gv.x, xv.x, sr[r]);
                    sr[r] = fmaf(-gv.y, xv.y, sr[r]);
                    si[r] = fmaf(gv.x, xv.y, si[r]);
                    si[r] = fmaf(gv.y, xv.x, si[r]);
                }
            }
        }
#pragma unroll
        for (int r = 0; r < K1G_RX; ++r) {
            if (rx0 + r < a.n_rx) {
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    sr[r] += __shfl_xor_sync(0xffffffffu, sr[r], o);
                    si[r] += __shfl_xor_sync(0xffffffffu, si[r], o);
                }
            }
        }
#pragma unroll
        for (int r = 0; r < K1G_RX; ++r) {
            if (rx0 + r < a.n_rx && lane == r) {
                const int rr = rx0 + r;
                const float2 cs = nco_cs(a.acc[rr] + a.inc[rr] * (u64)r0);
                float2 y;
                y.x = sr[r] * cs.x + si[r] * cs.y;      // (sr + j si) * (cos - j sin)
                y.y = si[r] * cs.x - sr[r] * cs.y;
                a.c_out[(size_t)rr * a.c_stride + a.hc + i] = y;
                if (a.bb_out) a.bb_out[(size_t)rr * a.bb_stride + i] = y;
            }
        }
    }
}

int k1_launch_generic(const K1Args &a, cudaStream_t st) {
    if (a.n_out <= 0) return PYSDR_OK;
    i64 blocks = (a.n_out + 7) / 8;
    const i64 cap = (i64)pysdr_sm_count() * 16;
    if (blocks > cap) blocks = cap;
    for (int rx0 = 0; rx0 < a.n_rx; rx0 += K1G_RX) {
        k1_generic_kernel<<<(unsigned)blocks, 256, 0, st>>>(a, rx0);
        LAUNCH_CHECK();
    }
    return PYSDR_OK;
}
