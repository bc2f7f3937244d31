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
#define WOLA_BUF (WOLA_ND + 3)      /* float2 per transform buffer */
#define WOLA_SP (WOLA_MB + 1)      /* staging row pitch (float2): an odd pitch spreads the per-instant column writes over the banks */

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

// W^{+m q}, q = 1..4, of a stage of length NI for the butterfly this thread owns (m = its offset within the block): the same
// butterfly in every output instant, so the four twiddles live in registers for the CTA's lifetime (r01 recomputed one
// sincospif and three products per stage per instant: ~40 % of the kernel's instructions)
struct WolaTw { float2 w1, w2, w3, w4; };
__device__ __forceinline__ WolaTw wola_twiddles(int m, int ni) {
    float sn, cs;
    sincospif(2.0f * (float)m / (float)ni, &sn, &cs);
    WolaTw t;
    t.w1 = make_float2(cs, sn);
    t.w2 = wcmul(t.w1, t.w1); t.w3 = wcmul(t.w2, t.w1); t.w4 = wcmul(t.w2, t.w2);
    return t;
}

template <int NI>
__device__ __forceinline__ void wola_stage(float2 *s, int u, const WolaTw &tw) {
    constexpr int STRIDE = NI / 5;
    const int blk = u / STRIDE, m = u - blk * STRIDE;
    const int base = blk * NI + m;
    float2 v[5];
#pragma unroll
    for (int q = 0; q < 5; ++q) v[q] = s[base + q * STRIDE];
    idft5(v);
    if (STRIDE > 1) {
        v[1] = wcmul(v[1], tw.w1); v[2] = wcmul(v[2], tw.w2); v[3] = wcmul(v[3], tw.w3); v[4] = wcmul(v[4], tw.w4);
    }
#pragma unroll
    for (int q = 0; q < 5; ++q) s[base + q * STRIDE] = v[q];
}

__global__ void __launch_bounds__(WOLA_T) wola_kernel(const WolaArgs a) {
    extern __shared__ __align__(16) float2 smem[];
    float2 *s = smem;                                    // 2 x [3125] transform buffers
    float2 *stg = smem + 2 * WOLA_BUF;                   // [n_ch][WOLA_SP] staging for the transposed store
    const int tid = threadIdx.x;
    const i64 mi0 = (i64)blockIdx.x * WOLA_MB;
    const int nm_here = (int)((a.n_out - mi0 < WOLA_MB) ? (a.n_out - mi0) : WOLA_MB);
    const WolaTw tw0 = wola_twiddles(tid, WOLA_ND);                          // pruned first stage: s[j + 625 r] = v[j] W_3125^{r j}
    const WolaTw tw1 = wola_twiddles(tid % 125, 625), tw2 = wola_twiddles(tid % 25, 125), tw3 = wola_twiddles(tid % 5, 25);
    // the (at most two) channels this thread picks in every instant: bin position and LO increment stay in registers
    int pos_c[2];
    u64 inc_c[2];
#pragma unroll
    for (int k = 0; k < 2; ++k) {
        const int c = tid + k * WOLA_T;
        pos_c[k] = c < a.n_ch ? __ldg(a.pos + c) : 0;
        inc_c[k] = c < a.n_ch ? a.inc[c] : 0ull;
    }
    // Two output instants per pass through the stages (two transform buffers): the five barriers of a transform are shared by
    // the pair and each thread has two independent butterflies in flight between them.
    for (int mi = 0; mi < nm_here; mi += 2) {
        i64 nm2[2];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            float2 *sb = s + h * WOLA_BUF;
            const i64 m = a.m0 + mi0 + mi + h;
            const i64 t = m * a.down;
            const i64 nm = t / a.up;                     // newest input sample of this output (absolute)
            const int pm = (int)(t - nm * a.up);
            nm2[h] = nm;
            // ---- windowing pass fused with the pruned first stage: s[j + 625 r] = v[j] W_3125^{+r j} ----------------
            float2 v = make_float2(0.f, 0.f);
            if (tid < a.lp && mi + h < nm_here) {
                const i64 rel = nm - tid - a.n0;         // index into x
                float2 xv = make_float2(0.f, 0.f);
                if (rel >= 0) { if (rel < a.n_in) xv = a.x[rel]; }
                else if (rel >= -a.n_before) xv = a.hist[a.n_before + rel];
                v = wcmul(__ldg(a.g0 + (size_t)pm * a.lp + tid), xv);
            }
            sb[tid] = v;
            sb[tid + 625] = wcmul(v, tw0.w1);
            sb[tid + 1250] = wcmul(v, tw0.w2);
            sb[tid + 1875] = wcmul(v, tw0.w3);
            sb[tid + 2500] = wcmul(v, tw0.w4);
        }
        __syncthreads();
        wola_stage<625>(s, tid, tw1); wola_stage<625>(s + WOLA_BUF, tid, tw1); __syncthreads();
        wola_stage<125>(s, tid, tw2); wola_stage<125>(s + WOLA_BUF, tid, tw2); __syncthreads();
        wola_stage<25>(s, tid, tw3); wola_stage<25>(s + WOLA_BUF, tid, tw3); __syncthreads();
        wola_stage<5>(s, tid, tw3); wola_stage<5>(s + WOLA_BUF, tid, tw3); __syncthreads();      // stride 1: no twiddles
        // ---- pick each channel's bin, de-rotate by its exact LO phase at n_m ----------------------------------------
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            if (mi + h >= nm_here) break;
#pragma unroll
            for (int k = 0; k < 2; ++k) {
                const int c = tid + k * WOLA_T;
                if (c < a.n_ch) {
                    const float2 z = s[h * WOLA_BUF + pos_c[k]];
                    const u64 ph = inc_c[k] * (u64)nm2[h];
                    float sn, cs;                                            // the same phase -> angle law as K1's de-rotation
                    __sincosf((float)(int)(ph >> 32) * 1.4629180792671596e-09f, &sn, &cs);
                    stg[c * WOLA_SP + mi + h] = make_float2(fmaf(z.x, cs, z.y * sn), fmaf(z.y, cs, -z.x * sn));   // z e^{-j th}
                }
            }
        }
        __syncthreads();
    }
    // ---- transposed store: WOLA_MB consecutive outputs per channel ---------------------------------------------------
    for (int e = tid; e < a.n_ch * WOLA_MB; e += WOLA_T) {
        const int c = e / WOLA_MB, mi = e - c * WOLA_MB;
        if (mi < nm_here) a.out[(size_t)c * a.out_stride + mi0 + mi] = stg[c * WOLA_SP + mi];
    }
}

extern "C" int pysdr_wola_channelize(const void *d_x, const void *d_hist, int64_t n0, int64_t n_before, int64_t n_in, int64_t m0,
                                     int64_t n_out,
                                     int32_t up, int32_t down, int32_t lp, const void *d_g0, int32_t n_ch, const int32_t *d_pos,
                                     const uint64_t *d_inc, void *d_out, int64_t out_stride, void *stream) {
    if (!d_x || !d_g0 || !d_pos || !d_inc || !d_out || n_out < 0 || up < 1 || down < 1 || lp < 1 || lp > WOLA_T || n_ch < 1 ||
        n_ch > WOLA_MAX_CH || n_ch > 2 * WOLA_T || out_stride < n_out) {
        pysdr_set_error("wola_channelize: need 1 <= lp <= %d taps per phase and 1 <= n_ch <= %d (got lp=%d n_ch=%d)", WOLA_T,
                        WOLA_MAX_CH, lp, n_ch);
        return PYSDR_ERR_ARG;
    }
    if (n_out == 0) return PYSDR_OK;
    WolaArgs a;
    a.x = (const float2 *)d_x; a.hist = d_hist ? (const float2 *)d_hist : (const float2 *)d_x - n_before; a.n0 = n0; a.n_before = n_before; a.n_in = n_in; a.m0 = m0; a.n_out = n_out;
    a.up = up; a.down = down; a.lp = lp; a.g0 = (const float2 *)d_g0; a.n_ch = n_ch; a.pos = d_pos;
    a.inc = (const u64 *)d_inc; a.out = (float2 *)d_out; a.out_stride = out_stride;
    const size_t smem = sizeof(float2) * (2 * WOLA_BUF + (size_t)n_ch * WOLA_SP);
    CUDA_TRY(cudaFuncSetAttribute(wola_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const i64 blocks = (n_out + WOLA_MB - 1) / WOLA_MB;
    wola_kernel<<<(unsigned)blocks, WOLA_T, smem, (cudaStream_t)stream>>>(a);
    LAUNCH_CHECK();
    return PYSDR_OK;
}
