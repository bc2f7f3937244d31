// wola.cu — K1c: many channels on a uniform raster from ONE windowing pass and ONE inverse DFT per output instant
// (BASELINE config 5: 1024 channel receivers, 9.6 kHz apart, on a 10 MS/s stream, 3/625).
//
// The per-channel fused mix + polyphase FIR that K1 evaluates (same indexing contract, same folded taps)
//
//     Y[c, m] = e^{-j th_c(n_m)} * sum_j h[p_m + UP*j] * e^{+j w_c j} * x[n_m - j],      w_c = 2 pi (f0 + c*df) / fs
//
// shares everything but the last factor between channels when df/fs = a/3125 in lowest terms (9.6 kHz / 10 MS/s = 3/3125):
//
//     v_m[j]  = G0[p_m][j] * x[n_m - j],   G0[p][j] = h[p + UP*j] * e^{+j w_0 j}          (lp <= 3125 products)
//     Y[c, m] = e^{-j th_c(n_m)} * Z_m[(a*c) mod 3125],   Z_m[k] = sum_j v_m[j] e^{+j 2 pi j k / 3125}
//
// Z_m is a 3125 = 5^5 point inverse DFT: radix-5 decimation in frequency in shared memory, 625 threads = one butterfly
// each per stage, the first stage pruned (only the first fifth of the input can be non-zero when lp <= 625), output
// picked in digit-reversed position order (host supplies pos[c] = digitrev5(a*c mod 3125)).  A CTA produces WOLA_MB
// consecutive output instants for all channels and writes them transposed (64 B per channel).
// Identity checked in float64 by tools/wola_prototype.py; parity with ChannelBank's K1 in tests/test_gpu_edges.py.
#include "common.cuh"

#define WOLA_ND 3125
#define WOLA_T 625
#define WOLA_MB 8
#define WOLA_MAX_CH 1024

struct WolaArgs {
    const float2 *x;          // samples, x[0] = absolute sample n0
    const float2 *hist;       // the n_before samples preceding x[0] (hist[n_before-1] = sample n0-1); may point at x - n_before
    i64 n0, n_before, n_in;   // anything earlier than the history (or >= n_in) reads as zero
    i64 m0, n_out;            // absolute index of the first output, number of outputs
    int up, down, lp;
    const float2 *g0;         // [up][lp] folded taps of channel 0
    int n_ch;
    const int *pos;           // [n_ch] position of channel c's bin in the transform's output order
    const u64 *inc;           // [n_ch] LO phase increment per sample (phase 0 at absolute sample 0)
    float2 *out;              // [n_ch][out_stride]
    i64 out_stride;
};

__device__ __forceinline__ float2 wcmul(float2 a, float2 b) {
    return make_float2(fmaf(a.x, b.x, -a.y * b.y), fmaf(a.x, b.y, a.y * b.x));
}

// inverse-sign 5-point DFT in registers
__device__ __forceinline__ void idft5(float2 *v) {
    const float c1 = 0.30901699437494742f, c2 = -0.80901699437494742f, s1 = 0.95105651629515357f, s2 = 0.58778525229247313f;
    const float2 t1 = make_float2(v[1].x + v[4].x, v[1].y + v[4].y), t2 = make_float2(v[2].x + v[3].x, v[2].y + v[3].y);
    const float2 t3 = make_float2(v[1].x - v[4].x, v[1].y - v[4].y), t4 = make_float2(v[2].x - v[3].x, v[2].y - v[3].y);
    const float2 x0 = v[0];
    const float2 a1 = make_float2(fmaf(c1, t1.x, fmaf(c2, t2.x, x0.x)), fmaf(c1, t1.y, fmaf(c2, t2.y, x0.y)));
    const float2 a2 = make_float2(fmaf(c2, t1.x, fmaf(c1, t2.x, x0.x)), fmaf(c2, t1.y, fmaf(c1, t2.y, x0.y)));
    const float2 b1 = make_float2(fmaf(s1, t3.x, s2 * t4.x), fmaf(s1, t3.y, s2 * t4.y));
    const float2 b2 = make_float2(fmaf(s2, t3.x, -s1 * t4.x), fmaf(s2, t3.y, -s1 * t4.y));
    v[0] = make_float2(x0.x + t1.x + t2.x, x0.y + t1.y + t2.y);
    v[1] = make_float2(a1.x - b1.y, a1.y + b1.x);            // a1 + j b1
    v[4] = make_float2(a1.x + b1.y, a1.y - b1.x);            // a1 - j b1
    v[2] = make_float2(a2.x - b2.y, a2.y + b2.x);
    v[3] = make_float2(a2.x + b2.y, a2.y - b2.x);
}

template <int NI>
__device__ __forceinline__ void wola_stage(float2 *s, int u) {
    constexpr int STRIDE = NI / 5;
    const int blk = u / STRIDE, m = u - blk * STRIDE;
    const int base = blk * NI + m;
    float2 v[5];
#pragma unroll
    for (int q = 0; q < 5; ++q) v[q] = s[base + q * STRIDE];
    idft5(v);
    if (STRIDE > 1) {
        float sn, cs;
        sincospif(2.0f * (float)m / (float)NI, &sn, &cs);     // W^{+m}
        const float2 w1 = make_float2(cs, sn), w2 = wcmul(w1, w1), w3 = wcmul(w2, w1), w4 = wcmul(w2, w2);
        v[1] = wcmul(v[1], w1); v[2] = wcmul(v[2], w2); v[3] = wcmul(v[3], w3); v[4] = wcmul(v[4], w4);
    }
#pragma unroll
    for (int q = 0; q < 5; ++q) s[base + q * STRIDE] = v[q];
}

__global__ void __launch_bounds__(WOLA_T) wola_kernel(const WolaArgs a) {
    extern __shared__ __align__(16) float2 smem[];
    float2 *s = smem;                                    // [3125] transform buffer
    float2 *stg = smem + WOLA_ND + 3;                    // [n_ch][WOLA_MB] staging for the transposed store
    const int tid = threadIdx.x;
    const i64 mi0 = (i64)blockIdx.x * WOLA_MB;
    const int nm_here = (int)((a.n_out - mi0 < WOLA_MB) ? (a.n_out - mi0) : WOLA_MB);
    for (int mi = 0; mi < nm_here; ++mi) {
        const i64 m = a.m0 + mi0 + mi;
        const i64 t = m * a.down;
        const i64 nm = t / a.up;                         // newest input sample of this output (absolute)
        const int pm = (int)(t - nm * a.up);
        // ---- windowing pass fused with the pruned first stage: s[j + 625 r] = v[j] W_3125^{+r j} --------------------
        {
            float2 v = make_float2(0.f, 0.f);
            if (tid < a.lp) {
                const i64 rel = nm - tid - a.n0;         // index into x
                float2 xv = make_float2(0.f, 0.f);
                if (rel >= 0) { if (rel < a.n_in) xv = a.x[rel]; }
                else if (rel >= -a.n_before) xv = a.hist[a.n_before + rel];
                v = wcmul(__ldg(a.g0 + (size_t)pm * a.lp + tid), xv);
            }
            float sn, cs;
            sincospif(2.0f * (float)tid / (float)WOLA_ND, &sn, &cs);
            const float2 w1 = make_float2(cs, sn), w2 = wcmul(w1, w1), w3 = wcmul(w2, w1), w4 = wcmul(w2, w2);
            s[tid] = v;
            s[tid + 625] = wcmul(v, w1);
            s[tid + 1250] = wcmul(v, w2);
            s[tid + 1875] = wcmul(v, w3);
            s[tid + 2500] = wcmul(v, w4);
        }
        __syncthreads();
        wola_stage<625>(s, tid); __syncthreads();
        wola_stage<125>(s, tid); __syncthreads();
        wola_stage<25>(s, tid); __syncthreads();
        wola_stage<5>(s, tid); __syncthreads();
        // ---- pick each channel's bin, de-rotate by its exact LO phase at n_m ----------------------------------------
        for (int c = tid; c < a.n_ch; c += WOLA_T) {
            const float2 z = s[a.pos[c]];
            const float2 cs = nco_cs(a.inc[c] * (u64)nm);
            stg[c * WOLA_MB + mi] = make_float2(fmaf(z.x, cs.x, z.y * cs.y), fmaf(z.y, cs.x, -z.x * cs.y));   // z e^{-j th}
        }
        __syncthreads();
    }
    // ---- transposed store: WOLA_MB consecutive outputs per channel ---------------------------------------------------
    for (int e = tid; e < a.n_ch * WOLA_MB; e += WOLA_T) {
        const int c = e / WOLA_MB, mi = e - c * WOLA_MB;
        if (mi < nm_here) a.out[(size_t)c * a.out_stride + mi0 + mi] = stg[e];
    }
}

extern "C" int pysdr_wola_channelize(const void *d_x, const void *d_hist, int64_t n0, int64_t n_before, int64_t n_in, int64_t m0,
                                     int64_t n_out,
                                     int32_t up, int32_t down, int32_t lp, const void *d_g0, int32_t n_ch, const int32_t *d_pos,
                                     const uint64_t *d_inc, void *d_out, int64_t out_stride, void *stream) {
    if (!d_x || !d_g0 || !d_pos || !d_inc || !d_out || n_out < 0 || up < 1 || down < 1 || lp < 1 || lp > WOLA_T || n_ch < 1 ||
        n_ch > WOLA_MAX_CH || out_stride < n_out) {
        pysdr_set_error("wola_channelize: need 1 <= lp <= %d taps per phase and 1 <= n_ch <= %d (got lp=%d n_ch=%d)", WOLA_T,
                        WOLA_MAX_CH, lp, n_ch);
        return PYSDR_ERR_ARG;
    }
    if (n_out == 0) return PYSDR_OK;
    WolaArgs a;
    a.x = (const float2 *)d_x; a.hist = d_hist ? (const float2 *)d_hist : (const float2 *)d_x - n_before; a.n0 = n0; a.n_before = n_before; a.n_in = n_in; a.m0 = m0; a.n_out = n_out;
    a.up = up; a.down = down; a.lp = lp; a.g0 = (const float2 *)d_g0; a.n_ch = n_ch; a.pos = d_pos;
    a.inc = (const u64 *)d_inc; a.out = (float2 *)d_out; a.out_stride = out_stride;
    const size_t smem = sizeof(float2) * (WOLA_ND + 3 + (size_t)n_ch * WOLA_MB);
    CUDA_TRY(cudaFuncSetAttribute(wola_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const i64 blocks = (n_out + WOLA_MB - 1) / WOLA_MB;
    wola_kernel<<<(unsigned)blocks, WOLA_T, smem, (cudaStream_t)stream>>>(a);
    LAUNCH_CHECK();
    return PYSDR_OK;
}
