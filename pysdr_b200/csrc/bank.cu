// bank.cu — receiver bank: K2 audio-rate stages + the C ABI of pysdr_bank_* (include/pysdr_b200.h).
//
// Stage order of one process() (mirrors dsp.Receiver.demod_data, reference receiver.py:235):
//   K1  fused mix + polyphase decimate            -> C[rx] = [hc history | n_out new]   (k1_*.cu)
//   K2a detect (AM |.| , NFM discriminator nfm.m:123-127) -> R[rx]
//   K2b AF FIR (real / complex->real / complex), CW BFO re-insertion -> a[rx] (pre-AGC audio)
//   K2c per-block peak of |a|                      -> peaks[rx][block]
//   K2d AGC recursion over blocks (agc.m loop filter)   -> gains[rx][block]
//   K2e gain + per-block DC removal (receiver.py:250-252) -> am, am_dc
//   K2f roll the complex memory
#include <math.h>
#include <stdarg.h>
#include <stdlib.h>

#include "common.cuh"

// ------------------------------------------------------------------------------------------------
static thread_local char g_err[512] = "";
void pysdr_set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
extern "C" const char *pysdr_last_error(void) { return g_err; }

bool pysdr_pdl_enabled(void) {
    static int on = -1;
    if (on < 0) { const char *e = getenv("PYSDR_NO_PDL"); on = (e && e[0] == '1') ? 0 : 1; }
    return on != 0;
}
int pysdr_device(void) {
    int d = 0;
    if (cudaGetDevice(&d) != cudaSuccess || d < 0) d = 0;
    return d;
}
int pysdr_sm_count(void) {
    static int cached[64] = {0};
    const int d = pysdr_device() & 63;
    if (cached[d] <= 0) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, d) != cudaSuccess || n <= 0) n = 148;
        cached[d] = n;
    }
    return cached[d];
}
extern "C" int pysdr_version(void) { return 100; }

extern "C" uint64_t pysdr_freq_to_phase_inc(double f, double fs) {
    double r = f / fs;
    r = r - floor(r);
    return (uint64_t)(r * 18446744073709551616.0);
}
extern "C" double pysdr_phase_inc_to_freq(uint64_t inc, double fs) {
    return (double)(int64_t)inc / 18446744073709551616.0 * fs;
}

// ------------------------------------------------------------------------------------------------
__global__ void quad_mixer_kernel(const float2 *__restrict__ x, float2 *__restrict__ y, i64 n, u64 acc0,
                                  u64 inc) {
    i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    const i64 stride = (i64)gridDim.x * blockDim.x;
    for (; i < n; i += stride) {
        const float2 cs = nco_cs(acc0 + inc * (u64)i);
        const float2 v = x[i];
        y[i] = make_float2(v.x * cs.x + v.y * cs.y, v.y * cs.x - v.x * cs.y);
    }
}

extern "C" int pysdr_quad_mixer(const void *d_x, void *d_y, int64_t n, uint64_t acc0, uint64_t inc,
                                void *stream) {
    if (n <= 0) return PYSDR_OK;
    i64 blocks = (n + 255) / 256;
    if (blocks > (i64)pysdr_sm_count() * 32) blocks = (i64)pysdr_sm_count() * 32;
    quad_mixer_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>((const float2 *)d_x, (float2 *)d_y,
                                                                        n, acc0, inc);
    LAUNCH_CHECK();
    return PYSDR_OK;
}

// CS16 -> CF32 (reference receiver.py:614-617: xxx = sc*xx[0::2] + 1j*sc*xx[1::2], sc = 1/2048).  Done on the device
// so that int16 sources cross PCIe at 4 bytes per sample instead of 8.  One 16-byte load = 4 complex samples.
__global__ void cs16_to_cf32_kernel(const short *__restrict__ in, float2 *__restrict__ out, i64 n, float scale) {
    const i64 n4 = n >> 2;
    i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    const i64 stride = (i64)gridDim.x * blockDim.x;
    const bool aligned = (((unsigned long long)in & 15ull) == 0) && (((unsigned long long)out & 15ull) == 0);
    if (aligned) {
        for (; i < n4; i += stride) {
            const int4 v = ((const int4 *)in)[i];
            const int w[4] = {v.x, v.y, v.z, v.w};
            float4 o[2];
            float *of = (float *)o;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                of[2 * k] = (float)(short)(w[k] & 0xffff) * scale;
                of[2 * k + 1] = (float)(short)(w[k] >> 16) * scale;
            }
            ((float4 *)out)[2 * i] = o[0];
            ((float4 *)out)[2 * i + 1] = o[1];
        }
        i = (n4 << 2) + (i64)blockIdx.x * blockDim.x + threadIdx.x;
    } else {
        i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    }
    for (; i < n; i += stride) out[i] = make_float2((float)in[2 * i] * scale, (float)in[2 * i + 1] * scale);
}

extern "C" int pysdr_cs16_to_cf32(const void *d_in, void *d_out, int64_t n, double scale, void *stream) {
    if (n <= 0) return PYSDR_OK;
    if (!d_in || !d_out) { pysdr_set_error("cs16_to_cf32: null pointer"); return PYSDR_ERR_ARG; }
    i64 blocks = ((n >> 2) + 255) / 256 + 1;
    if (blocks > (i64)pysdr_sm_count() * 16) blocks = (i64)pysdr_sm_count() * 16;
    cs16_to_cf32_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>((const short *)d_in, (float2 *)d_out, n, (float)scale);
    LAUNCH_CHECK();
    return PYSDR_OK;
}

// mean |x|^2 of a chunk (auto-mute detector); single CTA, deterministic tree
__global__ void __launch_bounds__(1024) mean_power_kernel(const float2 *__restrict__ x, i64 n, float *__restrict__ out) {
    __shared__ double sm[32];
    double acc = 0.0;
    for (i64 i = threadIdx.x; i < n; i += blockDim.x) {
        const float2 v = x[i];
        acc += (double)v.x * v.x + (double)v.y * v.y;
    }
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int w = 0; w < 32; ++w) t += sm[w];
        out[0] = (float)(t / (double)n);
    }
}

extern "C" int pysdr_mean_power(const void *d_x, int64_t n, float *d_out, void *stream) {
    if (!d_x || !d_out || n < 1) { pysdr_set_error("mean_power: bad arguments"); return PYSDR_ERR_ARG; }
    mean_power_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>((const float2 *)d_x, n, d_out);
    LAUNCH_CHECK();
    return PYSDR_OK;
}

// ------------------------------------------------------------------------------------------------
// K2a: detection over the whole complex memory + new samples.  R[k] for k in [0, L-1+n_out):
//   AM : |C[k+2]|          NFM: Re(C[k+1])*Im(d) - Im(C[k+1])*Re(d), d = C[k+2]-C[k]   (nfm.m:124-126)
__global__ void detect_kernel(const float2 *__restrict__ C, float *__restrict__ R, i64 n, int nfm, int sync) {
    i64 k = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    const i64 stride = (i64)gridDim.x * blockDim.x;
    for (; k < n; k += stride) {
        const float2 c2 = C[k + 2];
        float v;
        if (nfm) {
            const float2 c0 = C[k], c1 = C[k + 1];
            const float dr = c2.x - c0.x, di = c2.y - c0.y;
            v = c1.x * di - c1.y * dr;
        } else if (sync) {
            v = c2.x;                                         // AM-Synch: in-phase arm of the de-rotated memory
        } else {
            v = sqrtf(c2.x * c2.x + c2.y * c2.y);
        }
        R[k] = v;
    }
}

// K2b: AF FIR.  out[i] = sum_j g[j] * src[i + (L-1) - j],  i in [0, n_out).
// KIND 0: real src, real taps -> real.  KIND 1: complex src, complex taps -> real part.
// KIND 2: complex src, real taps -> complex, then optional BFO rotation -> real (CW) or complex (IQ).
#define FIR_THREADS 256
#define FIR_PER_THREAD 4
#define FIR_TILE (FIR_THREADS * FIR_PER_THREAD)

template <int KIND>
__global__ void __launch_bounds__(FIR_THREADS)
af_fir_kernel(const void *__restrict__ src_v, const float2 *__restrict__ taps, int L, i64 n_out,
              float *__restrict__ out, int cw, u64 bfo_inc, i64 m0) {
    extern __shared__ float4 smem_raw[];
    typedef typename std::conditional<KIND == 0, float, float2>::type src_t;
    typedef typename std::conditional<KIND == 1, float2, float>::type tap_t;
    src_t *s_src = (src_t *)smem_raw;
    tap_t *s_tap = (tap_t *)(s_src + FIR_TILE + L - 1 + 1);
    const src_t *src = (const src_t *)src_v;
    const i64 o0 = (i64)blockIdx.x * FIR_TILE;
    const int tid = threadIdx.x;
    const i64 avail = (L - 1) + n_out;                   // valid src elements
    for (int e = tid; e < FIR_TILE + L - 1; e += FIR_THREADS) {
        const i64 g = o0 + e;
        src_t v;
        if (g < avail) v = src[g];
        else memset(&v, 0, sizeof(v));
        s_src[e] = v;
    }
    for (int j = tid; j < L; j += FIR_THREADS) {
        if (KIND == 1) ((float2 *)s_tap)[j] = taps[j];
        else ((float *)s_tap)[j] = taps[j].x;
    }
    __syncthreads();

    float ar[FIR_PER_THREAD], ai[FIR_PER_THREAD];
#pragma unroll
    for (int i = 0; i < FIR_PER_THREAD; ++i) { ar[i] = 0.f; ai[i] = 0.f; }
    const int base = tid + (L - 1);
#pragma unroll 4
    for (int j = 0; j < L; ++j) {
        if (KIND == 0) {
            const float g = ((const float *)s_tap)[j];
#pragma unroll
            for (int i = 0; i < FIR_PER_THREAD; ++i)
                ar[i] = fmaf(g, ((const float *)s_src)[base + i * FIR_THREADS - j], ar[i]);
        } else if (KIND == 1) {
            const float2 g = ((const float2 *)s_tap)[j];
#pragma unroll
            for (int i = 0; i < FIR_PER_THREAD; ++i) {
                const float2 v = ((const float2 *)s_src)[base + i * FIR_THREADS - j];
                ar[i] = fmaf(g.x, v.x, ar[i]);
                ar[i] = fmaf(-g.y, v.y, ar[i]);
            }
        } else {
            const float g = ((const float *)s_tap)[j];
#pragma unroll
            for (int i = 0; i < FIR_PER_THREAD; ++i) {
                const float2 v = ((const float2 *)s_src)[base + i * FIR_THREADS - j];
                ar[i] = fmaf(g, v.x, ar[i]);
                ai[i] = fmaf(g, v.y, ai[i]);
            }
        }
    }
#pragma unroll
    for (int i = 0; i < FIR_PER_THREAD; ++i) {
        const i64 o = o0 + tid + i * FIR_THREADS;
        if (o >= n_out) continue;
        if (KIND == 2) {
            if (cw) {
                const float2 cs = nco_cs(bfo_inc * (u64)(m0 + o));
                out[o] = ar[i] * cs.x - ai[i] * cs.y;     // Re{ z * e^{+j th} }
            } else {
                ((float2 *)out)[o] = make_float2(ar[i], ai[i]);
            }
        } else {
            out[o] = ar[i];
        }
    }
}

// Output index range [lo,hi) (relative to this call) of AGC/DC block b (relative to this call), warp-uniform: lanes 0/1
// do the two 64-bit divisions, the rest of the warp takes the result by shuffle.
__device__ __forceinline__ void block_range_warp(i64 b, i64 B0, i64 in_chunk, int up, int down, i64 m0, i64 n_out,
                                                 i64 &lo, i64 &hi) {
    const int lane = threadIdx.x & 31;
    i64 v = 0;
    if (lane < 2) v = ((i64)up * (B0 + b + lane) * in_chunk + down - 1) / down - m0;
    const i64 s = __shfl_sync(0xffffffffu, v, 0), e = __shfl_sync(0xffffffffu, v, 1);
    lo = s < 0 ? 0 : s;
    hi = e > n_out ? n_out : e;
}

// End-of-call state update: work item r < n_rx rolls complex memory C[r][0..hc) <- C[r][n_out..n_out+hc), work item n_rx
// moves the raw input memory on (the last `need` samples of [hist | x]).  Any block size; memories are <= 4096 samples.
struct StateArgs {
    float2 *C; i64 c_stride, n_out; int hc, n_rx;
    float2 *hist; const float2 *hist_src, *x; int need; i64 n_in;
    int enabled;
};
__device__ __forceinline__ void state_update_item(const StateArgs &sa, int item) {
    float2 tmp[16];
    const int T = blockDim.x;
    if (item < sa.n_rx) {
        float2 *row = sa.C + (size_t)item * sa.c_stride;
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            const int e = threadIdx.x + k * T;
            if (e < sa.hc) tmp[k] = row[sa.n_out + e];
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            const int e = threadIdx.x + k * T;
            if (e < sa.hc) row[e] = tmp[k];
        }
        return;
    }
#pragma unroll
    for (int k = 0; k < 16; ++k) {
        const int e = threadIdx.x + k * T;
        if (e < sa.need) {
            const i64 idx = sa.n_in - sa.need + e;           // relative to x[0]
            tmp[k] = idx >= 0 ? sa.x[idx] : (sa.hist_src ? sa.hist_src[sa.need + idx] : make_float2(0.f, 0.f));
        }
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 16; ++k) {
        const int e = threadIdx.x + k * T;
        if (e < sa.need) sa.hist[e] = tmp[k];
    }
}

// K2c: peak of |a| per block.  One warp per block (a block is ~OUT_CHUNK_SIZE = 1024 samples), 8 blocks per CTA,
// 8 loads in flight per lane.  grid (ceil(n_blocks/8), n_rx).
#define BLK_WARPS 8
// peak of |a| over AGC block `blk` of one receiver row (warp-collective); lane 0 stores it
__device__ __forceinline__ void block_peak_warp(const float *__restrict__ a, float *__restrict__ peak_out, i64 blk, i64 B0,
                                                i64 in_chunk, int up, int down, i64 m0, i64 n_out) {
    const int lane = threadIdx.x & 31;
    i64 lo, hi;
    block_range_warp(blk, B0, in_chunk, up, down, m0, n_out, lo, hi);
    float mx = 0.f;
    // scalar head up to the first 16-byte boundary, 16-byte body (8 loads = 128 B in flight per lane), scalar tail
    const i64 head = ((4 - (i64)(((unsigned long long)(a + lo) >> 2) & 3ull)) & 3);
    const i64 b0 = (lo + head < hi) ? lo + head : hi;
    const i64 n4 = (hi - b0) >> 2;
    if (lo + lane < b0) mx = fabsf(a[lo + lane]);
    const float4 *a4 = (const float4 *)(a + b0);
    for (i64 i = lane; i < n4; i += 32 * 8) {
        float4 v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) v[u] = (i + 32 * u < n4) ? a4[i + 32 * u] : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int u = 0; u < 8; ++u)
            mx = fmaxf(fmaxf(mx, fmaxf(fabsf(v[u].x), fabsf(v[u].y))), fmaxf(fabsf(v[u].z), fabsf(v[u].w)));
    }
    {
        const i64 t = b0 + (n4 << 2) + lane;
        if (t < hi) mx = fmaxf(mx, fabsf(a[t]));
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if (lane == 0) *peak_out = mx;
}

__global__ void __launch_bounds__(32 * BLK_WARPS)
block_peak_kernel(const float *__restrict__ a, i64 a_row, float *__restrict__ peaks, i64 peaks_row, i64 n_blocks, i64 B0,
                  i64 in_chunk, int up, int down, i64 m0, i64 n_out, int n_rows, const StateArgs sa) {
    if ((int)blockIdx.y == n_rows) {         // extra grid row: the end-of-call state update rides this launch
        if ((int)blockIdx.x <= sa.n_rx) state_update_item(sa, blockIdx.x);
        return;
    }
    const i64 blk = (i64)blockIdx.x * BLK_WARPS + (threadIdx.x >> 5);
    if (blk >= n_blocks) return;
    block_peak_warp(a + (size_t)blockIdx.y * a_row, peaks + (size_t)blockIdx.y * peaks_row + blk, blk, B0, in_chunk, up, down,
                    m0, n_out);
}

// K2d: AGC recursion, one thread per receiver.  Law documented in oracle/sig_proc_oracle.py (class agc);
// loop filter = reference sigs/agc.m:6-12.
__device__ __forceinline__ float agc_update(AgcState &s, double pk) {
    s.ring[s.k % PYSDR_AGC_NB] = pk;
    s.k += 1;
    double mb = 0.0;
#pragma unroll
    for (int i = 0; i < PYSDR_AGC_NB; ++i) mb = fmax(mb, s.ring[i]);
    s.maxbuf = mb;
    double want = s.ref / fmax(mb, 1.0e-9);
    want = fmin(want, 1.0e4);
    s.err = want - s.gain;
    if (want < s.gain) s.gain = want;
    else s.gain = s.beta * want + (1.0 - s.beta) * s.gain;
    return (float)s.gain;
}

struct AgcScanArgs {
    AgcState *state;                 // [n_rx]
    const float *peaks;              // [n_rx][peaks_stride]
    i64 peaks_stride;
    const float *prev_peaks;         // [n_rx][n_prev] or null
    i64 n_prev;
    float *gains;                    // [n_rx][gains_stride]
    i64 gains_stride;
    i64 n_blocks;
    i64 skip;                        // leading blocks of this call that are filter warm-up: not part of the recursion
    int n_rx;
    int enabled[PYSDR_MAX_RX];
};

// The per-block update is  gain <- f_b(gain) = min(w_b, beta*w_b + (1-beta)*gain)  (attack when w_b < gain,
// agc.m loop filter otherwise), w_b = min(ref / max(maxbuf_b, 1e-9), 1e4), maxbuf_b = max of the last 8 peaks.
// maxbuf/w are data-parallel.  f_b(g) = min(A, C + D*g) is closed under composition
//     (f2 o f1)(g) = min( min(A2, C2 + D2*A1),  (C2 + D2*C1) + (D2*D1)*g )
// so the recurrence over blocks — this call's own blocks and, in time-sharded runs, the replay of all EARLIER
// shards' blocks — is an ordered parallel scan in float64 (re-association changes gains at the 1e-16 level,
// far below the float32 gain that is applied).  One CTA per receiver.
#define AGC_TILE 3072                 /* one tile covers the 2812 blocks of the 60 s capture (36.9 KB of shared memory) */
#define AGC_THREADS 256

struct AgcFn { double A, C, D; };
__device__ __forceinline__ AgcFn agc_compose(const AgcFn &f1, const AgcFn &f2) {       // f2 after f1
    AgcFn r;
    r.A = fmin(f2.A, fma(f2.D, f1.A, f2.C));
    r.C = fma(f2.D, f1.C, f2.C);
    r.D = f2.D * f1.D;
    return r;
}

// the scan of ONE receiver by one CTA of AGC_THREADS threads (every thread of the CTA must call it)
__device__ void agc_scan_rx(const AgcScanArgs &p, const int rx) {
    const int tid = threadIdx.x;
    float *gains = p.gains + (size_t)rx * p.gains_stride;
    if (!p.enabled[rx]) {
        for (i64 b = tid; b < p.n_blocks; b += AGC_THREADS) gains[b] = 1.f;
        return;
    }
    __shared__ double s_w[AGC_TILE];
    __shared__ float s_pk[AGC_TILE + 8];
    __shared__ AgcFn s_fn[AGC_THREADS / 32];
    __shared__ AgcState st;
    __shared__ double s_g, s_err, s_mb;
    const float *prev = p.prev_peaks ? p.prev_peaks + (size_t)rx * p.n_prev : nullptr;
    const float *own = p.peaks + (size_t)rx * p.peaks_stride + p.skip;
    gains += p.skip;
    for (i64 b = tid; b < p.skip; b += AGC_THREADS) gains[b - p.skip] = 1.f;
    const i64 n_own = p.n_blocks - p.skip;
    const i64 n_prev = prev ? p.n_prev : 0;
    if (tid == 0) {
        st = p.state[rx];
        if (prev) {                                  // time shard: replay from the reset state
            for (int i = 0; i < PYSDR_AGC_NB; ++i) st.ring[i] = 0.0;
            st.k = 0; st.gain = 1.0; st.maxbuf = 0.0; st.err = 0.0;
        }
        s_g = st.gain; s_err = st.err; s_mb = st.maxbuf;
    }
    __syncthreads();
    const double ref = st.ref, beta = st.beta, D = 1.0 - st.beta;
    const i64 k0 = st.k;
    if (tid < 7) s_pk[6 - tid] = (float)st.ring[(int)(((k0 - 1 - tid) % 8 + 8) % 8)];   // 7 peaks before element 0
    const i64 n_total = n_prev + n_own;
    for (i64 t0 = 0; t0 < n_total;) {
        // a tile never straddles the prev/own boundary
        const bool in_prev = t0 < n_prev;
        const i64 lim = in_prev ? n_prev : n_total;
        const int len = (int)((lim - t0 < AGC_TILE) ? (lim - t0) : AGC_TILE);
        __syncthreads();
        for (int i = tid; i < len; i += AGC_THREADS) {
            const i64 e = t0 + i;
            s_pk[7 + i] = e < n_prev ? __ldcg(prev + e) : __ldcg(own + (e - n_prev));     // L2: written by other SMs in the fused kernel
        }
        __syncthreads();
        for (int i = tid; i < len; i += AGC_THREADS) {
            float mb = s_pk[i];
#pragma unroll
            for (int j = 1; j < 8; ++j) mb = fmaxf(mb, s_pk[i + j]);
            s_w[i] = fmin(ref / fmax((double)mb, 1.0e-9), 1.0e4);
            if (i == len - 1) s_mb = (double)mb;
        }
        __syncthreads();
        {
            // ordered parallel scan of the tile's block functions; thread t owns `per` consecutive blocks
            const double g_in = s_g;
            const int per = (len + AGC_THREADS - 1) / AGC_THREADS;
            const int i0 = tid * per;
            AgcFn f; f.A = 1.0e300; f.C = 0.0; f.D = 1.0;                  // identity
            for (int j = 0; j < per; ++j) {
                const int i = i0 + j;
                if (i < len) { AgcFn g; g.A = s_w[i]; g.C = beta * s_w[i]; g.D = D; f = agc_compose(f, g); }
            }
            AgcFn incl = f;
            for (int o = 1; o < 32; o <<= 1) {                             // inclusive ordered scan inside the warp
                AgcFn lo;
                lo.A = __shfl_up_sync(0xffffffffu, incl.A, o);
                lo.C = __shfl_up_sync(0xffffffffu, incl.C, o);
                lo.D = __shfl_up_sync(0xffffffffu, incl.D, o);
                if ((tid & 31) >= o) incl = agc_compose(lo, incl);
            }
            if ((tid & 31) == 31) s_fn[tid >> 5] = incl;
            AgcFn ex;                                                      // lanes before this one in the warp
            ex.A = __shfl_up_sync(0xffffffffu, incl.A, 1);
            ex.C = __shfl_up_sync(0xffffffffu, incl.C, 1);
            ex.D = __shfl_up_sync(0xffffffffu, incl.D, 1);
            __syncthreads();
            AgcFn pre; pre.A = 1.0e300; pre.C = 0.0; pre.D = 1.0;
            for (int w = 0; w < (tid >> 5); ++w) pre = agc_compose(pre, s_fn[w]);
            if ((tid & 31) > 0) pre = agc_compose(pre, ex);
            double g = fmin(pre.A, fma(pre.D, g_in, pre.C));               // gain entering this thread's blocks
            if (!in_prev) {
                const i64 b0 = t0 - n_prev;
                for (int j = 0; j < per; ++j) {
                    const int i = i0 + j;
                    if (i < len) {
                        const double w = s_w[i];
                        if (i == len - 1) s_err = w - g;
                        g = fmin(w, fma(D, g, beta * w));
                        gains[b0 + i] = (float)g;
                    }
                }
            }
            __syncthreads();                                               // every thread has read s_g
            if (tid == AGC_THREADS - 1) {
                const AgcFn tot = agc_compose(pre, f);
                s_g = fmin(tot.A, fma(tot.D, g_in, tot.C));
            }
        }
        __syncthreads();
        float ctxv = 0.f;                                                  // context for the next tile:
        if (tid < 7) ctxv = s_pk[len + tid];                               // the last 7 entries of [ctx | tile]
        __syncthreads();
        if (tid < 7) s_pk[tid] = ctxv;
        t0 += len;
    }
    __syncthreads();
    if (tid < 8) {                                                         // ring <- the last 8 peaks seen (one lane each)
        const i64 e = n_total - 1 - tid;
        if (e >= 0) st.ring[(int)((k0 + e) & 7)] = (double)(e < n_prev ? __ldcg(prev + e) : __ldcg(own + (e - n_prev)));
    }
    if (tid == 8) { st.gain = s_g; st.err = s_err; st.maxbuf = s_mb; st.k = k0 + n_total; }
    __syncthreads();
    static_assert(sizeof(AgcState) % 8 == 0, "AgcState is copied in 8-byte words");
    if (tid < (int)(sizeof(AgcState) / 8))
        ((unsigned long long *)&p.state[rx])[tid] = ((const unsigned long long *)&st)[tid];
}

__global__ void __launch_bounds__(AGC_THREADS) agc_scan_kernel(AgcScanArgs p) { agc_scan_rx(p, blockIdx.x); }

// ---- O(1) AGC carry between time shards (include/pysdr_b200.h: pysdr_bank_agc_summary / _enter) -----------------------
// One CTA per receiver: ordered reduction (composition is associative, not commutative) of the block functions of
// blocks 7 .. n-1; every w_b there depends on the shard's own peaks only.
#define SUM_THREADS 256
// summary of receiver rx into o[LEN] (o may be shared or global memory); block-wide, ends with a barrier
__device__ void agc_summary_rx(const AgcState *__restrict__ state, const float *__restrict__ peaks, i64 peaks_stride, i64 skip,
                               i64 n_blocks, const int rx, double *o) {
    const int tid = threadIdx.x;
    const float *own = peaks + (size_t)rx * peaks_stride + skip;
    const i64 n = n_blocks - skip;
    const double ref = state[rx].ref, beta = state[rx].beta, D = 1.0 - beta;
    __shared__ AgcFn s_fn[SUM_THREADS];
    // thread t owns the contiguous run [lo, hi) of blocks 7 .. n-1
    const i64 m = n > 7 ? n - 7 : 0;
    const i64 per = (m + SUM_THREADS - 1) / SUM_THREADS;
    const i64 lo = 7 + (i64)tid * per, hi = (lo + per < n) ? lo + per : n;
    AgcFn f; f.A = 1.0e300; f.C = 0.0; f.D = 1.0;
    for (i64 b = lo; b < hi; ++b) {
        float mb = __ldcg(own + b);
#pragma unroll
        for (int j = 1; j < 8; ++j) mb = fmaxf(mb, __ldcg(own + b - j));
        const double w = fmin(ref / fmax((double)mb, 1.0e-9), 1.0e4);
        AgcFn g; g.A = w; g.C = beta * w; g.D = D;
        f = agc_compose(f, g);
    }
    s_fn[tid] = f;
    __syncthreads();
    for (int o2 = 1; o2 < SUM_THREADS; o2 <<= 1) {          // ordered tree: element t absorbs its RIGHT neighbour run
        if ((tid & (2 * o2 - 1)) == 0) s_fn[tid] = agc_compose(s_fn[tid], s_fn[tid + o2]);
        __syncthreads();
    }
    if (tid == 0) { o[0] = s_fn[0].A; o[1] = s_fn[0].C; o[2] = s_fn[0].D; o[18] = (double)n; }
    if (tid < 7) o[3 + tid] = tid < n ? (double)__ldcg(own + tid) : 0.0;
    if (tid < 8) { const i64 e = n - 8 + tid; o[10 + tid] = e >= 0 ? (double)__ldcg(own + e) : 0.0; }
    __syncthreads();
}

__global__ void __launch_bounds__(SUM_THREADS) agc_summary_kernel(const AgcState *__restrict__ state, const float *__restrict__ peaks,
                                                                  i64 peaks_stride, i64 skip, i64 n_blocks, double *__restrict__ out) {
    agc_summary_rx(state, peaks, peaks_stride, skip, n_blocks, blockIdx.x, out + (size_t)blockIdx.x * PYSDR_AGC_SUMMARY_LEN);
}

// ---- the carry exchange over NVLink peer memory (no NCCL on the data path) ------------------------------------------------
// Every rank owns one symmetric buffer (torch.distributed._symmetric_memory: the same allocation mapped into every peer):
//     slots  double [DEPTH][world][n_rx][LEN]     summaries, slot = seq % DEPTH, row = the WRITING rank
//     flags  u64    [DEPTH][world]                flags[slot][q] = seq once rank q's summaries of step seq have landed here
//     acks   u64    [world]                       acks[q] = last step whose summaries rank q has finished reading
//     count  u64    [2]                           local: CTAs of the push kernel / readers of the back kernel that are done
// agc_summary_push_kernel computes this rank's summaries and STORES them straight into the slots of the LATER ranks (only
// they read them), fences system-wide, and the last CTA to finish raises the rank's flag in those peers.  The consumer is the
// fused back kernel: the scanner CTA of a receiver spins on the flags of the ranks before it, reads their summaries from its
// own copy and — last reader of the step — stores its ack into the earlier ranks' buffers.  A rank only ever waits for
// EARLIER ranks' data and for LATER ranks' acks of step seq - DEPTH (slot reuse), so the waits cannot form a cycle.
#define XCHG_DEPTH PYSDR_XCHG_DEPTH
struct XchgPeers { double *base[PYSDR_XCHG_MAX_WORLD]; };
__device__ __forceinline__ void st_release_sys_u64(unsigned long long *p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys_u64(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long global_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
// bounded spin: a peer that never shows up (dead rank) must not hang this GPU — trap after 10 s, the context reports an error
__device__ __forceinline__ void xchg_wait_ge(const unsigned long long *p, unsigned long long v) {
    if (ld_acquire_sys_u64(p) >= v) return;
    const unsigned long long t0 = global_ns();
    while (ld_acquire_sys_u64(p) < v) {
        __nanosleep(100);
        if (global_ns() - t0 > 10000000000ull) __trap();
    }
}
__host__ __device__ __forceinline__ size_t xchg_slot_doubles(int world, int n_rx) { return (size_t)world * n_rx * PYSDR_AGC_SUMMARY_LEN; }
__host__ __device__ __forceinline__ size_t xchg_flags_off(int world, int n_rx) { return (size_t)XCHG_DEPTH * xchg_slot_doubles(world, n_rx); }
__device__ __forceinline__ unsigned long long *xchg_flags(double *base, int world, int n_rx) {
    return (unsigned long long *)(base + xchg_flags_off(world, n_rx));
}
__device__ __forceinline__ unsigned long long *xchg_acks(double *base, int world, int n_rx) {
    return xchg_flags(base, world, n_rx) + (size_t)XCHG_DEPTH * world;
}
__device__ __forceinline__ unsigned long long *xchg_counts(double *base, int world, int n_rx) {
    return xchg_acks(base, world, n_rx) + world;
}

// block-wide (SUM_THREADS threads): summary of receiver rx -> the later ranks' slots; the last receiver's CTA raises the flags
__device__ void agc_summary_push_rx(const AgcState *__restrict__ state, const float *__restrict__ peaks, i64 peaks_stride, i64 skip,
                                    i64 n_blocks, const XchgPeers &peers, int world, int rank, int n_rx, unsigned long long seq, int rx) {
    __shared__ double s_o[PYSDR_AGC_SUMMARY_LEN];
    const int tid = threadIdx.x;
    agc_summary_rx(state, peaks, peaks_stride, skip, n_blocks, rx, s_o);        // block-wide; s_o valid after the barrier inside
    const int slot = (int)(seq % XCHG_DEPTH);
    const size_t off = (size_t)slot * xchg_slot_doubles(world, n_rx) + ((size_t)rank * n_rx + rx) * PYSDR_AGC_SUMMARY_LEN;
    const int n_later = world - 1 - rank;
    if (seq > XCHG_DEPTH && tid < n_later)                                      // slot reuse: the later ranks are done with step seq - DEPTH
        xchg_wait_ge(xchg_acks(peers.base[rank], world, n_rx) + rank + 1 + tid, seq - XCHG_DEPTH);
    __syncthreads();
    for (int e = tid; e < n_later * PYSDR_AGC_SUMMARY_LEN; e += SUM_THREADS) {
        const int q = rank + 1 + e / PYSDR_AGC_SUMMARY_LEN, i = e % PYSDR_AGC_SUMMARY_LEN;
        peers.base[q][off + i] = s_o[i];                                       // NVLink store into peer q's memory
    }
    __syncthreads();
    if (tid == 0) {
        __threadfence_system();                                                // this CTA's peer stores before the count
        const unsigned long long prev = atomicAdd(xchg_counts(peers.base[rank], world, n_rx), 1ull);
        if (prev + 1 == seq * (unsigned long long)n_rx) {                      // every receiver's summary has been stored everywhere
            __threadfence_system();
            for (int q = rank + 1; q < world; ++q)
                st_release_sys_u64(xchg_flags(peers.base[q], world, n_rx) + (size_t)slot * world + rank, seq);
        }
    }
    __syncthreads();
}

__global__ void __launch_bounds__(SUM_THREADS) agc_summary_push_kernel(const AgcState *__restrict__ state, const float *__restrict__ peaks,
                                                                       i64 peaks_stride, i64 skip, i64 n_blocks, const XchgPeers peers,
                                                                       int world, int rank, int n_rx, unsigned long long seq) {
    agc_summary_push_rx(state, peaks, peaks_stride, skip, n_blocks, peers, world, rank, n_rx, seq, blockIdx.x);
}

// consumer side, one thread per receiver: wait for the earlier ranks' flags of step seq in the LOCAL buffer
struct XchgWait { XchgPeers peers; int world, rank; unsigned long long seq; };
__device__ __forceinline__ const double *xchg_wait_summaries(const XchgWait &x, int n_rx) {
    const int slot = (int)(x.seq % XCHG_DEPTH);
    double *base = x.peers.base[x.rank];
    const unsigned long long *fl = xchg_flags(base, x.world, n_rx) + (size_t)slot * x.world;
    for (int q = 0; q < x.rank; ++q) xchg_wait_ge(fl + q, x.seq);
    return base + (size_t)slot * xchg_slot_doubles(x.world, n_rx);
}
// ... and once its reads are done: the last of the n_rx readers of the step tells the earlier ranks the slot is free
__device__ __forceinline__ void xchg_ack(const XchgWait &x, int n_rx) {
    __threadfence_system();
    const unsigned long long prev = atomicAdd(xchg_counts(x.peers.base[x.rank], x.world, n_rx) + 1, 1ull);
    if (prev + 1 == x.seq * (unsigned long long)n_rx)
        for (int q = 0; q < x.rank; ++q) st_release_sys_u64(xchg_acks(x.peers.base[q], x.world, n_rx) + x.rank, x.seq);
}

// SYS: the summaries were written into this GPU's memory by PEER GPUs over NVLink (pysdr_bank_agc_summary_push): read them
// with volatile (ld.volatile -> system-coherent) loads.
template <bool SYS>
__device__ __forceinline__ double ld_sum(const double *p) { return SYS ? *(const volatile double *)p : __ldcg(p); }
template <bool SYS = false>
__device__ void agc_enter_rx(AgcState *state, const double *sums, int n_before, int n_rx, int rx) {
    AgcState s = state[rx];
    for (int i = 0; i < PYSDR_AGC_NB; ++i) s.ring[i] = 0.0;
    s.k = 0; s.gain = 1.0; s.maxbuf = 0.0; s.err = 0.0;
    for (int q = 0; q < n_before; ++q) {
        const double *o = sums + ((size_t)q * n_rx + rx) * PYSDR_AGC_SUMMARY_LEN;
        const i64 n = (i64)ld_sum<SYS>(o + 18);
        const int head = n < 7 ? (int)n : 7;
        for (int j = 0; j < head; ++j) agc_update(s, ld_sum<SYS>(o + 3 + j));
        if (n > 7) {
            const double g_in = s.gain;
            s.gain = fmin(ld_sum<SYS>(o), fma(ld_sum<SYS>(o + 2), g_in, ld_sum<SYS>(o + 1)));
            for (int t = 0; t < 8; ++t) s.ring[(int)((s.k + (n - 7) - 1 - t) & 7)] = ld_sum<SYS>(o + 17 - t);
            s.k += n - 7;
            double mb = 0.0;
            for (int i = 0; i < PYSDR_AGC_NB; ++i) mb = fmax(mb, s.ring[i]);
            s.maxbuf = mb;
            s.err = fmin(s.ref / fmax(mb, 1.0e-9), 1.0e4) - s.gain;
        }
    }
    state[rx] = s;
}
__global__ void agc_enter_xchg_kernel(AgcState *state, int n_rx, const XchgWait x) {
    if ((int)threadIdx.x >= n_rx) return;
    const double *sums = xchg_wait_summaries(x, n_rx);
    agc_enter_rx<true>(state, sums, x.rank, n_rx, threadIdx.x);
    xchg_ack(x, n_rx);
}
__global__ void agc_enter_kernel(AgcState *__restrict__ state, const double *__restrict__ sums, int n_before, int n_rx) {
    if ((int)threadIdx.x < n_rx) agc_enter_rx<false>(state, sums, n_before, n_rx, threadIdx.x);
}

// K2e: am = a*gain ; am_dc = am - mean_block(am) for AM/USB.  grid (n_blocks, n_rx); IQ rows are skipped.
struct ApplyKinds { int kind[PYSDR_MAX_RX]; };     // 0 skip, 1 real, 2 real + per-block DC removal, 3 complex copy (IQ / RTTY rows)
// gain (and optional block-mean removal) on AGC block `blk` of one receiver row, warp-collective; pointers are row bases
__device__ __forceinline__ void agc_apply_warp(const float *__restrict__ a, const float g, float *__restrict__ am,
                                               float *__restrict__ am_dc, const int kind, i64 blk, i64 B0, i64 in_chunk, int up,
                                               int down, i64 m0, i64 n_out) {
    const int lane = threadIdx.x & 31;
    const int dc_remove = kind == 2;
    i64 lo, hi;
    block_range_warp(blk, B0, in_chunk, up, down, m0, n_out, lo, hi);
    if (!am_dc && ((((unsigned long long)(a + lo)) ^ ((unsigned long long)(am + lo))) & 15ull) == 0) {
        // audio only, source and destination equally aligned: scalar head, 16-byte body, scalar tail
        const i64 head = ((4 - (i64)(((unsigned long long)(a + lo) >> 2) & 3ull)) & 3);
        const i64 b0 = (lo + head < hi) ? lo + head : hi;
        const i64 n4 = (hi - b0) >> 2;
        if (lo + lane < b0) am[lo + lane] = a[lo + lane] * g;
        const float4 *a4 = (const float4 *)(a + b0);
        float4 *o4 = (float4 *)(am + b0);
        for (i64 i = lane; i < n4; i += 32 * 8) {
            float4 v[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) if (i + 32 * u < n4) v[u] = a4[i + 32 * u];
#pragma unroll
            for (int u = 0; u < 8; ++u)
                if (i + 32 * u < n4) o4[i + 32 * u] = make_float4(v[u].x * g, v[u].y * g, v[u].z * g, v[u].w * g);
        }
        const i64 t = b0 + (n4 << 2) + lane;
        if (t < hi) am[t] = a[t] * g;
        return;
    }
    double sum = 0.0;                                   // block mean in float64 (the subtraction below cancels the carrier)
    for (i64 i = lo + lane; i < hi; i += 32 * 8) {
        float v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) v[u] = (i + 32 * u < hi) ? a[i + 32 * u] * g : 0.f;
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            if (i + 32 * u < hi) am[i + 32 * u] = v[u];
            sum += (double)v[u];
        }
    }
    if (!am_dc) return;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    const float mean = (hi > lo && dc_remove) ? (float)(sum / (double)(hi - lo)) : 0.f;
    for (i64 i = lo + lane; i < hi; i += 32) am_dc[i] = a[i] * g - mean;
}

__global__ void __launch_bounds__(32 * BLK_WARPS)
agc_apply_kernel(const float *__restrict__ a, i64 a_row, const float *__restrict__ gains, i64 g_row, float *__restrict__ am,
                 float *__restrict__ am_dc, i64 am_row, ApplyKinds kinds, i64 n_blocks, i64 B0, i64 in_chunk, int up, int down,
                 i64 m0, i64 n_out) {
    // one warp per block, 8 blocks per CTA: grid (ceil(n_blocks/8), n_rx)
    const int kind = kinds.kind[blockIdx.y];
    if (kind != 1 && kind != 2) return;
    const i64 blk = (i64)blockIdx.x * BLK_WARPS + (threadIdx.x >> 5);
    if (blk >= n_blocks) return;
    agc_apply_warp(a + (size_t)blockIdx.y * a_row, gains[(size_t)blockIdx.y * g_row + blk], am + (size_t)blockIdx.y * am_row,
                   am_dc ? am_dc + (size_t)blockIdx.y * am_row : nullptr, kind, blk, B0, in_chunk, up, down, m0, n_out);
}

// ---- fused "back": block peaks -> AGC scan -> gain / DC removal in ONE launch -------------------------------------------
// The three stages are separated by grid-wide dependencies (every block peak of a receiver before its scan, every gain
// before it is applied), which used to cost two extra launches and their drain/fill gaps for ~90 MB of traffic.  Here a
// co-resident grid (sized from the occupancy calculator, so that every CTA is running) walks the stages with two
// grid barriers on a monotonically increasing device counter.  Stage bodies are the same device functions the
// stand-alone kernels use (kept for the time-sharded split form and for stereo WFM2).
__device__ __forceinline__ unsigned long long ld_acquire_u64(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void grid_barrier(unsigned long long *ctr, unsigned long long target) {
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        atomicAdd(ctr, 1ull);
        while (ld_acquire_u64(ctr) < target) { }
        __threadfence();
    }
    __syncthreads();
}

struct BackArgs {
    int do_peaks, do_enter, n_before, n_rx;
    const float *a; i64 a_row;                       // pre-AGC audio rows (floats)
    float *peaks; i64 peaks_row;
    i64 n_blocks, B0, in_chunk, m0, n_out;
    int up, down;
    StateArgs sa;
    AgcScanArgs scan;
    const double *sums;
    const float *gains; i64 g_row;
    float *am, *am_dc; i64 am_row;
    ApplyKinds kinds;
    unsigned long long *bar; unsigned long long bar_base;
    XchgWait x;                                      // do_enter == 2: the peer-memory exchange buffers and the step number
    int do_push; i64 push_skip;                      // also compute this shard's summaries (blocks >= push_skip) and push them first
};

#define BACK_THREADS AGC_THREADS
__global__ void __launch_bounds__(BACK_THREADS) agc_back_fused_kernel(const BackArgs p) {
    const int warp = threadIdx.x >> 5, wpc = BACK_THREADS / 32;
    const i64 gw = (i64)blockIdx.x * wpc + warp, gstride = (i64)gridDim.x * wpc;
    const i64 total = (i64)p.n_rx * p.n_blocks;
    pdl_trigger();
    pdl_wait();                                      // the AF filter kernel's audio (and, in split runs, the block peaks)
    if (p.do_peaks) {
        // the end-of-call state update rides on the LAST CTAs (the first n_rx run the scans)
        const int item = (int)gridDim.x - 1 - (int)blockIdx.x;
        if (p.sa.enabled && item <= p.sa.n_rx) state_update_item(p.sa, item);
        for (i64 t = gw; t < total; t += gstride) {
            const int rx = (int)(t / p.n_blocks);
            const i64 blk = t - (i64)rx * p.n_blocks;
            block_peak_warp(p.a + (size_t)rx * p.a_row, p.peaks + (size_t)rx * p.peaks_row + blk, blk, p.B0, p.in_chunk, p.up,
                            p.down, p.m0, p.n_out);
        }
        grid_barrier(p.bar, p.bar_base + gridDim.x);
    }
    // The summaries for the later ranks are pushed by CTAs n_rx .. 2 n_rx - 1 while CTAs 0 .. n_rx - 1 enter and scan: the
    // NVLink stores and their system-scope fences (~15 us) stay off this rank's critical path.  (A grid too small for that —
    // a handful of blocks — pushes from the scanner CTAs first.)
    const bool push_here = p.do_push && p.x.rank < p.x.world - 1;
    const bool push_apart = push_here && (int)gridDim.x >= 2 * p.n_rx;
    if (push_apart && (int)blockIdx.x >= p.n_rx && (int)blockIdx.x < 2 * p.n_rx)
        agc_summary_push_rx(p.scan.state, p.peaks, p.peaks_row, p.push_skip, p.n_blocks, p.x.peers, p.x.world, p.x.rank, p.n_rx,
                            p.x.seq, (int)blockIdx.x - p.n_rx);
    if ((int)blockIdx.x < p.n_rx) {
        if (push_here && !push_apart)
            agc_summary_push_rx(p.scan.state, p.peaks, p.peaks_row, p.push_skip, p.n_blocks, p.x.peers, p.x.world, p.x.rank, p.n_rx,
                                p.x.seq, blockIdx.x);
        if (p.do_enter) {
            if (threadIdx.x == 0) {
                if (p.do_enter == 2) {               // summaries pushed by the earlier ranks over NVLink: wait for their flags
                    const double *sums = xchg_wait_summaries(p.x, p.n_rx);
                    agc_enter_rx<true>(p.scan.state, sums, p.x.rank, p.n_rx, blockIdx.x);
                    xchg_ack(p.x, p.n_rx);
                } else {
                    agc_enter_rx<false>(p.scan.state, p.sums, p.n_before, p.n_rx, blockIdx.x);
                }
            }
            __syncthreads();
        }
        agc_scan_rx(p.scan, blockIdx.x);
    }
    grid_barrier(p.bar, p.bar_base + (p.do_peaks ? 2ull : 1ull) * gridDim.x);
    for (i64 t = gw; t < total; t += gstride) {
        const int rx = (int)(t / p.n_blocks);
        const int kind = p.kinds.kind[rx];
        if (kind != 1 && kind != 2) continue;
        const i64 blk = t - (i64)rx * p.n_blocks;
        agc_apply_warp(p.a + (size_t)rx * p.a_row, __ldcg(p.gains + (size_t)rx * p.g_row + blk), p.am + (size_t)rx * p.am_row,
                       p.am_dc ? p.am_dc + (size_t)rx * p.am_row : nullptr, kind, blk, p.B0, p.in_chunk, p.up, p.down, p.m0,
                       p.n_out);
    }
    for (int rx = 0; rx < p.n_rx; ++rx) {            // complex rows (IQ / RTTY): plain copy of 2*n_out floats
        if (p.kinds.kind[rx] != 3) continue;
        const float *src = p.a + (size_t)rx * p.a_row;
        float *d0 = p.am + (size_t)rx * p.am_row, *d1 = p.am_dc ? p.am_dc + (size_t)rx * p.am_row : nullptr;
        for (i64 i = (i64)blockIdx.x * BACK_THREADS + threadIdx.x; i < 2 * p.n_out; i += (i64)gridDim.x * BACK_THREADS) {
            const float v = src[i];
            d0[i] = v;
            if (d1) d1[i] = v;
        }
    }
}

__global__ void copy_f32_kernel(const float *__restrict__ s, float *__restrict__ d, i64 n) {
    i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    const i64 stride = (i64)gridDim.x * blockDim.x;
    for (; i < n; i += stride) d[i] = s[i];
}

__global__ void hist_update_kernel(float2 *hist, const float2 *__restrict__ hist_src, const float2 *__restrict__ x,
                                   int need, i64 n_in) {
    // single CTA; read everything first (hist may alias hist_src)
    float2 tmp[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int e = threadIdx.x + k * 1024;
        if (e < need) {
            const i64 idx = n_in - need + e;                 // relative to x[0]
            tmp[k] = idx >= 0 ? x[idx] : (hist_src ? hist_src[need + idx] : make_float2(0.f, 0.f));
        }
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int e = threadIdx.x + k * 1024;
        if (e < need) hist[e] = tmp[k];
    }
}

// seek(): every carried state back to "nothing before this sample" in ONE launch: blocks 0..n_rx-1 clear the complex
// memories, block n_rx clears the raw input memory and the PLL states and (at the stream origin) resets the AGCs.
__global__ void __launch_bounds__(256) seek_reset_kernel(float2 *C, i64 c_stride, int hc, int n_rx, float2 *hist, int n_hist,
                                                         double2 *pll, AgcState *agc, int reset_agc) {
    const float2 z = make_float2(0.f, 0.f);
    if ((int)blockIdx.x < n_rx) {
        float2 *row = C + (size_t)blockIdx.x * c_stride;
        for (int e = threadIdx.x; e < hc; e += blockDim.x) row[e] = z;
        return;
    }
    for (int e = threadIdx.x; e < n_hist; e += blockDim.x) hist[e] = z;
    if (threadIdx.x < PYSDR_MAX_RX) {
        const int r = threadIdx.x;
        pll[r] = make_double2(0.0, 0.0);
        if (reset_agc) {
            for (int i = 0; i < PYSDR_AGC_NB; ++i) agc[r].ring[i] = 0.0;
            agc[r].k = 0; agc[r].gain = 1.0; agc[r].maxbuf = 0.0; agc[r].err = 0.0;
        }
    }
}

// ------------------------------------------------------------------------------------------------
#define K1_CHAN_MIN_ROWS 2048                /* the same for the many-channel kernel (k1_chan.cu) */
#define K1_MMA_MIN_ROWS 8192                 /* interior super-periods from which the tensor-core K1 takes a call (mode 1) */
struct pysdr_bank {
    pysdr_bank_config cfg;
    int lp, lp_pad, need, hc;
    i64 max_out, max_blocks, c_stride, r_stride, a_stride;
    i64 n0;                                  // absolute input index of the next sample
    u64 inc[PYSDR_MAX_RX], acc0[PYSDR_MAX_RX];   // LO: phase(n) = acc0 + inc*n
    std::vector<float> dec_h[PYSDR_MAX_RX];
    bool g_dirty;
    int mode[PYSDR_MAX_RX];
    int af_cplx[PYSDR_MAX_RX];
    u64 bfo_inc[PYSDR_MAX_RX];
    bool demod_set[PYSDR_MAX_RX];
    // device
    bool h_dirty[PYSDR_MAX_RX];             // AF taps changed: their position-order FFT must be refreshed
    bool force_direct_fir;                  // testing: use the direct-form AF FIR instead of the FFT path
    bool k1_only;                           // WFM video stage: stop after K1
    bool real_input;                        // the caller promises Im x == 0 (pysdr_bank_set_real_input)
    bool k1_external;                       // the caller fills C[r][hc .. hc+n_out) itself (raster channelizer, wola.cu)
    bool c_external;                        // d_C was adopted from the caller (not freed here)
    float2 *d_H;                            // [n_rx][Nfft] FFT of AF taps (position order, 1/N folded in)
    float2 *d_hist, *d_g, *d_C, *d_af;
    float2 *d_a;                            // pre-AGC audio, rows of a_stride float2 (IQ mode fills complex)
    float *d_R, *d_peaks, *d_gains;
    AgcState *d_agc;
    double2 *d_pll;            // AM-Synch loop state (phi, w) per receiver
    bool stereo;               // WFM2 resampler bank: rows 0/1/2 = sum / difference / pilot -> L, R
    float pilot_min;
    double pll_k1, pll_k2;
    // pending front->back
    i64 pend_n_out, pend_m0, pend_B0, pend_blocks;
    const float *pend_peaks;
    bool pending;
    bool force_generic;
    K1MmaPlan *mma;                          // tensor-core K1 (k1_mma.cu); null when the geometry does not fit
    K1ChanPlan *chan;                        // many-channel tensor-core K1 (k1_chan.cu); null below 16 receivers
    int mma_mode;                            // 0 never, 1 calls of at least K1_MMA_MIN_ROWS interior super-periods, 2 whenever possible
    int k1_last;                             // kernel of the last call: 0 generic, 1 tap-stationary, 2 tensor-core interior + edges
    // fused back (agc_back_fused_kernel): block peaks deferred from front into the back launch; grid barrier counter
    bool force_unfused;                      // testing: the stand-alone tail kernels
    bool defer_peaks, peaks_deferred;
    StateArgs pend_sa;
    unsigned long long *d_bar;               // grid barrier counter of the fused back kernel (monotonic)
    unsigned long long bar_count;            // its value when the next launch starts
    int back_grid;
    // host-chunk executive (pysdr_bank_process_host): pinned staging + own device buffers, allocated on first use
    float2 *hc_h_in, *hc_d_in, *hc_d_iq;
    float *hc_d_am, *hc_d_dc;                 // device result rows — or, when the device can address page-locked host memory
    bool hc_zero_copy;                       // directly, the device view of hc_h_out: the back kernel stores its results there
    char *hc_h_out;                          // pinned: am [n_rx][2 max_out] f32 | iq [n_rx][max_out] c64 | am_dc [n_rx][2 max_out] f32
    // seek() folded into the next process_front / process_back (no launch of its own)
    bool lazy_seek, lazy_reset_agc;
    i64 launches;
    // optional on-stream stage timing (bench.py roofline): events e0 |K1| e1 |rest of front| e2 ... e3 |back| e4
    bool timing;
    std::vector<cudaEvent_t> evs;
};

static int flush_seek(pysdr_bank *b, cudaStream_t st);
static int launch_block_peaks(pysdr_bank *b, float *d_peaks, const StateArgs &sa, i64 n_blocks, i64 B0, i64 m0, i64 n_out,
                              cudaStream_t st);

static int bank_alloc(pysdr_bank *b) {
    const pysdr_bank_config &c = b->cfg;
    CUDA_TRY(cudaMalloc(&b->d_hist, sizeof(float2) * (size_t)(b->need + 8)));
    CUDA_TRY(cudaMalloc(&b->d_g, sizeof(float2) * (size_t)c.n_rx * c.up * b->lp_pad));
    CUDA_TRY(cudaMalloc(&b->d_C, sizeof(float2) * (size_t)c.n_rx * b->c_stride));
    CUDA_TRY(cudaMalloc(&b->d_af, sizeof(float2) * (size_t)c.n_rx * c.af_len));
    b->d_H = nullptr;
    if (fftconv_supported(c.af_len))
        CUDA_TRY(cudaMalloc(&b->d_H, sizeof(float2) * (size_t)c.n_rx * fftconv_n_for(c.af_len)));
    CUDA_TRY(cudaMalloc(&b->d_R, sizeof(float) * (size_t)c.n_rx * b->r_stride));
    CUDA_TRY(cudaMalloc(&b->d_a, sizeof(float2) * (size_t)c.n_rx * b->a_stride));
    CUDA_TRY(cudaMalloc(&b->d_peaks, sizeof(float) * (size_t)c.n_rx * b->max_blocks));
    CUDA_TRY(cudaMalloc(&b->d_gains, sizeof(float) * (size_t)c.n_rx * b->max_blocks));
    CUDA_TRY(cudaMalloc(&b->d_agc, sizeof(AgcState) * PYSDR_MAX_RX));
    CUDA_TRY(cudaMalloc(&b->d_pll, sizeof(double2) * PYSDR_MAX_RX));
    CUDA_TRY(cudaMalloc(&b->d_bar, sizeof(unsigned long long) * (2 * PYSDR_MAX_RX + 2)));
    CUDA_TRY(cudaMemset(b->d_bar, 0, sizeof(unsigned long long) * (2 * PYSDR_MAX_RX + 2)));
    return PYSDR_OK;
}

static void agc_host_reset(AgcState &s, double ref, double beta) {
    memset(&s, 0, sizeof(s));
    s.gain = 1.0;
    s.ref = ref;
    s.beta = beta;
}

extern "C" int pysdr_bank_reset(pysdr_bank *b) {
    if (!b) { pysdr_set_error("null bank"); return PYSDR_ERR_ARG; }
    b->n0 = 0;
    b->pending = false;
    b->lazy_seek = false; b->lazy_reset_agc = false; b->peaks_deferred = false;
    CUDA_TRY(cudaMemset(b->d_hist, 0, sizeof(float2) * (size_t)(b->need + 8)));
    CUDA_TRY(cudaMemset(b->d_C, 0, sizeof(float2) * (size_t)b->cfg.n_rx * b->c_stride));
    CUDA_TRY(cudaMemset(b->d_pll, 0, sizeof(double2) * PYSDR_MAX_RX));
    AgcState st[PYSDR_MAX_RX];
    AgcState cur[PYSDR_MAX_RX];
    CUDA_TRY(cudaMemcpy(cur, b->d_agc, sizeof(cur), cudaMemcpyDeviceToHost));
    for (int r = 0; r < PYSDR_MAX_RX; ++r) agc_host_reset(st[r], cur[r].ref, cur[r].beta);
    CUDA_TRY(cudaMemcpy(b->d_agc, st, sizeof(st), cudaMemcpyHostToDevice));
    return PYSDR_OK;
}

extern "C" int pysdr_bank_create(const pysdr_bank_config *cfg, pysdr_bank **out) {
    if (!cfg || !out) { pysdr_set_error("null argument"); return PYSDR_ERR_ARG; }
    if (cfg->n_rx < 1 || cfg->n_rx > PYSDR_MAX_RX || cfg->up < 1 || cfg->down < 1 || cfg->filt_len < 1 ||
        cfg->af_len < 1 || cfg->in_chunk < 1 || cfg->max_in < 1) {
        pysdr_set_error("bad bank config (n_rx=%d up=%d down=%d filt_len=%d af_len=%d)", cfg->n_rx, cfg->up,
                        cfg->down, cfg->filt_len, cfg->af_len);
        return PYSDR_ERR_ARG;
    }
    pysdr_bank *b = new pysdr_bank();
    b->cfg = *cfg;
    b->lp = (cfg->filt_len + cfg->up - 1) / cfg->up;
    b->lp_pad = k1_fast_lp_pad(b->lp);
    b->need = b->lp - 1;
    b->hc = cfg->af_len + 1;
    {   // AM-Synch loop gains: noise bandwidth 50 Hz, damping 1/sqrt(2) at the audio rate
        const double fs_out = cfg->srate * cfg->up / cfg->down, zeta = 0.70710678118654752440, bn = 50.0;
        const double th = bn / fs_out / (zeta + 1.0 / (4.0 * zeta)), d = 1.0 + 2.0 * zeta * th + th * th;
        b->pll_k1 = 4.0 * zeta * th / d;
        b->pll_k2 = 4.0 * th * th / d;
    }
    if (b->need > 4096 || b->hc > 4096) {
        pysdr_set_error("filter memories above 4096 samples are not supported (need=%d hc=%d)", b->need, b->hc);
        delete b;
        return PYSDR_ERR_ARG;
    }
    b->max_out = ((i64)cfg->up * cfg->max_in) / cfg->down + 2;
    b->max_blocks = (cfg->max_in + cfg->in_chunk - 1) / cfg->in_chunk + 1;
    b->c_stride = (b->hc + b->max_out + 3) / 2 * 2;
    b->r_stride = (cfg->af_len - 1 + b->max_out + 3) / 4 * 4;
    b->a_stride = (b->max_out + 1) / 2 * 2;
    b->g_dirty = true;
    b->force_generic = false;
    b->mma = k1_fast_supported(cfg->up, cfg->down, b->lp, cfg->n_rx) ? k1_mma_plan_create(cfg->up, cfg->down, b->lp, cfg->n_rx) : nullptr;
    b->chan = k1_chan_plan_create(cfg->up, cfg->down, b->lp, cfg->n_rx);
    b->mma_mode = 1; b->k1_last = -1;
    if (const char *e = getenv("PYSDR_K1_MMA")) b->mma_mode = atoi(e);
    b->force_direct_fir = false;
    b->k1_only = false;
    b->k1_external = false; b->c_external = false; b->real_input = false;
    b->stereo = false; b->pilot_min = 0.f;
    b->timing = false;
    b->launches = 0;
    b->pending = false;
    b->force_unfused = false; b->defer_peaks = false; b->peaks_deferred = false;
    b->bar_count = 0; b->back_grid = 0;
    b->lazy_seek = false; b->lazy_reset_agc = false;
    b->hc_h_in = nullptr; b->hc_d_in = nullptr; b->hc_d_iq = nullptr; b->hc_d_am = nullptr; b->hc_d_dc = nullptr; b->hc_h_out = nullptr; b->hc_zero_copy = false;
    for (int r = 0; r < PYSDR_MAX_RX; ++r) {
        b->inc[r] = 0; b->acc0[r] = 0; b->mode[r] = PYSDR_MODE_IQ; b->af_cplx[r] = 0; b->bfo_inc[r] = 0;
        b->demod_set[r] = false;
        b->h_dirty[r] = true;
    }
    int rc = bank_alloc(b);
    if (rc) { delete b; return rc; }
    AgcState st[PYSDR_MAX_RX];
    for (int r = 0; r < PYSDR_MAX_RX; ++r) agc_host_reset(st[r], 0.25, 0.1);
    if (cudaMemcpy(b->d_agc, st, sizeof(st), cudaMemcpyHostToDevice) != cudaSuccess) {
        pysdr_set_error("agc init copy failed");
        delete b;
        return PYSDR_ERR_CUDA;
    }
    rc = pysdr_bank_reset(b);
    if (rc) { delete b; return rc; }
    *out = b;
    return PYSDR_OK;
}

extern "C" int pysdr_bank_destroy(pysdr_bank *b) {
    if (!b) return PYSDR_OK;
    cudaFree(b->d_H);
    if (b->hc_h_in) cudaFreeHost(b->hc_h_in);
    if (b->hc_h_out) cudaFreeHost(b->hc_h_out);
    cudaFree(b->hc_d_in); cudaFree(b->hc_d_iq);
    if (!b->hc_zero_copy) { cudaFree(b->hc_d_am); cudaFree(b->hc_d_dc); }
    if (!b->c_external) cudaFree(b->d_C);
    k1_mma_plan_destroy(b->mma);
    k1_chan_plan_destroy(b->chan);
    cudaFree(b->d_hist); cudaFree(b->d_g); cudaFree(b->d_af); cudaFree(b->d_R);
    cudaFree(b->d_a); cudaFree(b->d_peaks); cudaFree(b->d_gains); cudaFree(b->d_agc); cudaFree(b->d_pll); cudaFree(b->d_bar);
    delete b;
    return PYSDR_OK;
}

#define CHECK_RX(b, rx)                                                                    \
    if (!(b) || (rx) < 0 || (rx) >= (b)->cfg.n_rx) {                                       \
        pysdr_set_error("bad bank/receiver index %d", (int)(rx));                          \
        return PYSDR_ERR_ARG;                                                              \
    }

extern "C" int pysdr_bank_set_lo(pysdr_bank *b, int rx, uint64_t inc) {
    CHECK_RX(b, rx);
    // phase-continuous at the current position: acc0' + inc'*n0 == acc0 + inc*n0
    b->acc0[rx] = b->acc0[rx] + (b->inc[rx] - (u64)inc) * (u64)b->n0;
    b->inc[rx] = inc;
    b->g_dirty = true;
    return PYSDR_OK;
}

extern "C" int pysdr_bank_set_dec_taps(pysdr_bank *b, int rx, const float *h, int n) {
    CHECK_RX(b, rx);
    if (!h || n < 1 || n > b->cfg.filt_len) {
        pysdr_set_error("dec taps: n=%d exceeds FILT_LEN=%d", n, b->cfg.filt_len);
        return PYSDR_ERR_ARG;
    }
    b->dec_h[rx].assign(h, h + n);
    b->g_dirty = true;
    return PYSDR_OK;
}

extern "C" int pysdr_bank_set_demod(pysdr_bank *b, int rx, int mode, const float *taps, int n, int is_complex,
                                    uint64_t bfo_inc) {
    CHECK_RX(b, rx);
    if (mode == PYSDR_MODE_RAW) {                     // no demod filter: Re{resampler output} goes straight to the AGC
        b->mode[rx] = mode; b->af_cplx[rx] = 0; b->bfo_inc[rx] = 0; b->demod_set[rx] = true;
        return PYSDR_OK;
    }
    if (mode < PYSDR_MODE_AM || (mode > PYSDR_MODE_NFM && mode != PYSDR_MODE_AMSYNC) || !taps || n != b->cfg.af_len) {
        pysdr_set_error("set_demod: mode=%d n=%d (af_len=%d)", mode, n, b->cfg.af_len);
        return PYSDR_ERR_ARG;
    }
    const bool want_cplx = (mode == PYSDR_MODE_USB || mode == PYSDR_MODE_LSB);
    if (want_cplx != (is_complex != 0)) {
        pysdr_set_error("set_demod: mode %d needs %s taps", mode, want_cplx ? "complex" : "real");
        return PYSDR_ERR_ARG;
    }
    std::vector<float2> t(n);
    for (int j = 0; j < n; ++j) {
        if (is_complex) t[j] = make_float2(taps[2 * j], taps[2 * j + 1]);   // LSB: host passes conj(g)
        else t[j] = make_float2(taps[j], 0.f);
    }
    CUDA_TRY(cudaMemcpy(b->d_af + (size_t)rx * b->cfg.af_len, t.data(), sizeof(float2) * n, cudaMemcpyHostToDevice));
    if (mode == PYSDR_MODE_AMSYNC && b->demod_set[rx] && b->mode[rx] != PYSDR_MODE_AMSYNC)   // a real mode change, not set-up
        CUDA_TRY(cudaMemset(b->d_pll + rx, 0, sizeof(double2)));               // the loop starts from rest (receiver.py:649)
    b->mode[rx] = mode;
    b->af_cplx[rx] = is_complex;
    b->bfo_inc[rx] = bfo_inc;
    b->demod_set[rx] = true;
    b->h_dirty[rx] = true;
    return PYSDR_OK;
}

extern "C" int pysdr_bank_set_stereo(pysdr_bank *b, int on, double pilot_min) {
    if (!b) { pysdr_set_error("null bank"); return PYSDR_ERR_ARG; }
    if (on && b->cfg.n_rx != 3) { pysdr_set_error("set_stereo: the stereo resampler bank has exactly 3 rows (sum, difference, pilot)"); return PYSDR_ERR_ARG; }
    b->stereo = on != 0;
    b->pilot_min = (float)pilot_min;
    return PYSDR_OK;
}

extern "C" int pysdr_bank_pll_reset(pysdr_bank *b, int rx) {
    CHECK_RX(b, rx);
    { int rc = flush_seek(b, 0); if (rc) return rc; }
    CUDA_TRY(cudaMemset(b->d_pll + rx, 0, sizeof(double2)));
    return PYSDR_OK;
}

extern "C" int pysdr_bank_pll_get(pysdr_bank *b, int rx, double *out2, void *stream) {
    CHECK_RX(b, rx);
    if (!out2) { pysdr_set_error("pll_get: null output"); return PYSDR_ERR_ARG; }
    { int rc = flush_seek(b, (cudaStream_t)stream); if (rc) return rc; }
    CUDA_TRY(cudaStreamSynchronize((cudaStream_t)stream));
    CUDA_TRY(cudaMemcpy(out2, b->d_pll + rx, sizeof(double2), cudaMemcpyDeviceToHost));
    return PYSDR_OK;
}

extern "C" int pysdr_bank_agc_reset(pysdr_bank *b, int rx) {
    CHECK_RX(b, rx);
    { int rc = flush_seek(b, 0); if (rc) return rc; }
    AgcState s;
    CUDA_TRY(cudaMemcpy(&s, b->d_agc + rx, sizeof(s), cudaMemcpyDeviceToHost));
    agc_host_reset(s, s.ref, s.beta);
    CUDA_TRY(cudaMemcpy(b->d_agc + rx, &s, sizeof(s), cudaMemcpyHostToDevice));
    return PYSDR_OK;
}

extern "C" int pysdr_bank_agc_config(pysdr_bank *b, int rx, double ref, double beta) {
    CHECK_RX(b, rx);
    { int rc = flush_seek(b, 0); if (rc) return rc; }
    AgcState s;
    CUDA_TRY(cudaMemcpy(&s, b->d_agc + rx, sizeof(s), cudaMemcpyDeviceToHost));
    s.ref = ref;
    s.beta = beta;
    CUDA_TRY(cudaMemcpy(b->d_agc + rx, &s, sizeof(s), cudaMemcpyHostToDevice));
    return PYSDR_OK;
}

extern "C" int pysdr_bank_agc_get(pysdr_bank *b, int rx, double out5[5], void *stream) {
    CHECK_RX(b, rx);
    { int rc = flush_seek(b, (cudaStream_t)stream); if (rc) return rc; }
    AgcState s;
    CUDA_TRY(cudaStreamSynchronize((cudaStream_t)stream));
    CUDA_TRY(cudaMemcpy(&s, b->d_agc + rx, sizeof(s), cudaMemcpyDeviceToHost));
    // rx.agc.agc (reference watchdog.py:298) = the gain the loop WANTS for the current peak buffer (ref / maxbuf, clamped),
    // rx.agc.gain = the gain it applies after the attack / loop-filter law; err = agc - gain before the last update
    const double want = fmin(s.ref / fmax(s.maxbuf, 1.0e-9), 1.0e4);
    out5[0] = s.k > 0 ? want : s.gain; out5[1] = s.gain; out5[2] = s.maxbuf; out5[3] = s.ref; out5[4] = s.err;
    return PYSDR_OK;
}

extern "C" int pysdr_bank_agc_trace(pysdr_bank *b, float *host_peaks, float *host_gains, int64_t capacity, int64_t *n_blocks,
                                    void *stream) {
    if (!b || !host_peaks || !host_gains || !n_blocks) { pysdr_set_error("agc_trace: bad arguments"); return PYSDR_ERR_ARG; }
    const i64 nb = b->pend_blocks;
    if (b->pending || nb <= 0 || !b->pend_peaks) { pysdr_set_error("agc_trace: no completed process call"); return PYSDR_ERR_STATE; }
    if (nb > capacity) { pysdr_set_error("agc_trace: %lld blocks exceed capacity %lld", nb, (i64)capacity); return PYSDR_ERR_CAPACITY; }
    CUDA_TRY(cudaStreamSynchronize((cudaStream_t)stream));
    CUDA_TRY(cudaMemcpy2D(host_peaks, sizeof(float) * (size_t)nb, b->pend_peaks, sizeof(float) * (size_t)nb,
                          sizeof(float) * (size_t)nb, b->cfg.n_rx, cudaMemcpyDeviceToHost));
    CUDA_TRY(cudaMemcpy2D(host_gains, sizeof(float) * (size_t)nb, b->d_gains, sizeof(float) * (size_t)b->max_blocks,
                          sizeof(float) * (size_t)nb, b->cfg.n_rx, cudaMemcpyDeviceToHost));
    *n_blocks = nb;
    return PYSDR_OK;
}

extern "C" int64_t pysdr_bank_n_out(const pysdr_bank *b, int64_t n_in) {
    if (!b) return -1;
    return n_out_total(b->n0 + n_in, b->cfg.up, b->cfg.down) - n_out_total(b->n0, b->cfg.up, b->cfg.down);
}
extern "C" int64_t pysdr_bank_position(const pysdr_bank *b) { return b ? b->n0 : -1; }
extern "C" int64_t pysdr_bank_n_blocks(const pysdr_bank *b, int64_t n_in) {
    if (!b) return -1;
    return (n_in + b->cfg.in_chunk - 1) / b->cfg.in_chunk;
}
extern "C" int pysdr_bank_force_direct_fir(pysdr_bank *b, int on) {
    if (!b) return PYSDR_ERR_ARG;
    b->force_direct_fir = on != 0;
    return PYSDR_OK;
}
extern "C" int pysdr_bank_adopt_c_memory(pysdr_bank *b, void *d_ptr, int64_t row_stride) {
    if (!b || !d_ptr || row_stride < b->c_stride || (row_stride & 1)) {
        pysdr_set_error("adopt_c_memory: need an even row stride >= %lld complex64 elements", b ? (long long)b->c_stride : 0LL);
        return PYSDR_ERR_ARG;
    }
    if (!b->c_external) cudaFree(b->d_C);
    b->d_C = (float2 *)d_ptr;
    b->c_stride = row_stride;
    b->c_external = true;
    return PYSDR_OK;
}

extern "C" int pysdr_bank_set_k1_external(pysdr_bank *b, int on) {
    if (!b) { pysdr_set_error("null bank"); return PYSDR_ERR_ARG; }
    if (b->k1_external && !on) b->g_dirty = true;     // back to the bank's own K1: (re)build the tap images on the next call
    b->k1_external = on != 0;
    return PYSDR_OK;
}

extern "C" int pysdr_bank_set_real_input(pysdr_bank *b, int on) {
    if (!b) return PYSDR_ERR_ARG;
    b->real_input = on != 0;
    return PYSDR_OK;
}

extern "C" int pysdr_bank_set_k1_only(pysdr_bank *b, int on) {
    if (!b) return PYSDR_ERR_ARG;
    b->k1_only = on != 0;
    return PYSDR_OK;
}

// 3-point FM discriminator (reference sigs/nfm.m:123-127) at any rate, two carried samples.
__global__ void fm_disc_kernel(const float2 *__restrict__ y, i64 n, const float2 *__restrict__ prev2, float2 *__restrict__ out) {
    i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    const i64 stride = (i64)gridDim.x * blockDim.x;
    for (; i < n; i += stride) {
        const float2 c2 = y[i];
        const float2 c1 = i >= 1 ? y[i - 1] : prev2[1];
        const float2 c0 = i >= 2 ? y[i - 2] : prev2[i];       // i==0 -> prev2[0], i==1 -> prev2[1]
        const float dr = c2.x - c0.x, di = c2.y - c0.y;
        out[i] = make_float2(c1.x * di - c1.y * dr, 0.f);
    }
}
__global__ void fm_disc_tail_kernel(const float2 *__restrict__ y, i64 n, float2 *prev2) {
    // new carried samples = the last two of [prev2 | y]
    const float2 a = n >= 2 ? y[n - 2] : (n == 1 ? prev2[1] : prev2[0]);
    const float2 b = n >= 1 ? y[n - 1] : prev2[1];
    prev2[0] = a;
    prev2[1] = b;
}
extern "C" int pysdr_fm_disc(const void *d_y, int64_t n, void *d_prev2, void *d_out, void *stream) {
    if (!d_y || !d_prev2 || !d_out || n < 0) { pysdr_set_error("fm_disc: bad arguments"); return PYSDR_ERR_ARG; }
    if (n == 0) return PYSDR_OK;
    cudaStream_t st = (cudaStream_t)stream;
    i64 blocks = (n + 255) / 256;
    if (blocks > (i64)pysdr_sm_count() * 16) blocks = (i64)pysdr_sm_count() * 16;
    fm_disc_kernel<<<(unsigned)blocks, 256, 0, st>>>((const float2 *)d_y, n, (const float2 *)d_prev2, (float2 *)d_out);
    LAUNCH_CHECK();
    fm_disc_tail_kernel<<<1, 1, 0, st>>>((const float2 *)d_y, n, (float2 *)d_prev2);
    LAUNCH_CHECK();
    return PYSDR_OK;
}

// AM-Synch carrier PLL (open choice): second-order loop, atan2 phase detector.
//     v = z e^{-j phi};  e = atan2(Im v, Re v);  w += K2 e;  phi += w + K1 e  (wrapped to [-pi, pi))
// The new samples of C are replaced by v IN PLACE, so the carried memory holds de-rotated samples and the AF FIR
// reads Re v.  Inherently serial at the audio rate: one warp per receiver, lane 0 runs the recurrence on a
// shared-memory tile the whole warp loads and stores.
struct PllModes { int mode[PYSDR_MAX_RX]; };
#define PLL_TILE 512
__global__ void __launch_bounds__(32) am_pll_kernel(float2 *__restrict__ Cbase, i64 c_stride, int hc, i64 n, const PllModes pm,
                                                    double2 *__restrict__ state, double k1, double k2) {
    const int rx = blockIdx.x;
    if (pm.mode[rx] != PYSDR_MODE_AMSYNC) return;
    __shared__ float2 t[PLL_TILE];
    float2 *C = Cbase + (size_t)rx * c_stride + hc;
    double phi = state[rx].x, w = state[rx].y;
    const double PI = 3.14159265358979323846, TWO_PI = 6.28318530717958647692;
    for (i64 base = 0; base < n; base += PLL_TILE) {
        const int cnt = (int)((n - base < PLL_TILE) ? (n - base) : PLL_TILE);
        for (int i = threadIdx.x; i < cnt; i += 32) t[i] = C[base + i];
        __syncwarp();
        if (threadIdx.x == 0) {
            for (int i = 0; i < cnt; ++i) {
                float sn, cs;
                sincosf((float)phi, &sn, &cs);
                const float2 z = t[i];
                const float vr = z.x * cs + z.y * sn, vi = z.y * cs - z.x * sn;
                t[i] = make_float2(vr, vi);
                const double e = (double)atan2f(vi, vr);
                w += k2 * e;
                phi += w + k1 * e;
                if (phi >= PI) phi -= TWO_PI;
                else if (phi < -PI) phi += TWO_PI;
            }
        }
        __syncwarp();
        for (int i = threadIdx.x; i < cnt; i += 32) C[base + i] = t[i];
        __syncwarp();
    }
    if (threadIdx.x == 0) state[rx] = make_double2(phi, w);
}

// WFM2 stereo matrix at the audio rate (open choice).  Rows of the resampler bank fed with the
// real FM multiplex: z0 = sum channel (LO 0), z1 = difference channel (LO 38 kHz), z2 = pilot (LO 19 kHz, narrow AF
// low-pass).  u = z2/|z2|;  S = Re z0;  D = 2 Re{ z1 conj(u)^2 }  (0 when |z2| <= pilot_min);  L = S + D, R = S - D.
__global__ void stereo_matrix_kernel(const float2 *__restrict__ z0, const float2 *__restrict__ z1, const float2 *__restrict__ z2,
                                     i64 n, float pilot_min, float *__restrict__ L, float *__restrict__ R) {
    i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    const i64 stride = (i64)gridDim.x * blockDim.x;
    for (; i < n; i += stride) {
        const float2 p = z2[i], d = z1[i];
        const float S = z0[i].x;
        const float m2 = p.x * p.x + p.y * p.y;
        float D = 0.f;
        if (m2 > pilot_min * pilot_min && m2 > 0.f) {
            const float inv = 1.0f / m2;                                  // u^2 = p^2 / |p|^2
            const float u2r = (p.x * p.x - p.y * p.y) * inv, u2i = 2.f * p.x * p.y * inv;
            D = 2.f * (d.x * u2r + d.y * u2i);                            // 2 Re{ d conj(u^2) }
        }
        L[i] = S + D;
        R[i] = S - D;
    }
}

__global__ void peak_link_kernel(float *__restrict__ p0, float *__restrict__ p1, i64 n) {   // one AGC for the stereo pair
    const i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        const float m = fmaxf(p0[i], p1[i]);
        p0[i] = m;
        p1[i] = m;
    }
}

__global__ void real_part_kernel(const float2 *__restrict__ c, float *__restrict__ a, i64 n) {
    i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    const i64 stride = (i64)gridDim.x * blockDim.x;
    for (; i < n; i += stride) a[i] = c[i].x;
}

extern "C" int pysdr_bank_force_generic(pysdr_bank *b, int on) {
    if (!b) return PYSDR_ERR_ARG;
    b->force_generic = on != 0;
    return PYSDR_OK;
}
extern "C" int pysdr_bank_set_k1_mma(pysdr_bank *b, int mode) {
    if (!b || mode < 0 || mode > 2) { pysdr_set_error("set_k1_mma: mode 0 (never), 1 (large calls) or 2 (whenever possible)"); return PYSDR_ERR_ARG; }
    b->mma_mode = mode;
    return PYSDR_OK;
}
extern "C" int pysdr_bank_k1_mma_available(const pysdr_bank *b) { return (b && b->mma) ? 1 : 0; }
extern "C" int pysdr_bank_k1_last(const pysdr_bank *b) { return b ? b->k1_last : -1; }
extern "C" int pysdr_bank_k1_variant(const pysdr_bank *b) {
    if (!b) return -1;
    return (!b->force_generic && k1_fast_supported(b->cfg.up, b->cfg.down, b->lp, b->cfg.n_rx)) ? 1 : 0;
}
extern "C" int64_t pysdr_bank_launch_count(const pysdr_bank *b) { return b ? b->launches : -1; }

extern "C" int pysdr_bank_c_memory(pysdr_bank *b, void **d_ptr, int64_t *row_stride, int32_t *hist_len) {
    if (!b || !d_ptr || !row_stride || !hist_len) { pysdr_set_error("c_memory: bad arguments"); return PYSDR_ERR_ARG; }
    *d_ptr = b->d_C;
    *row_stride = b->c_stride;
    *hist_len = b->hc;
    return PYSDR_OK;
}

// folded taps: G[rx][p][j] = h[p + j*up] * exp(+j*2*pi*frac(inc*j / 2^64)), zero padded to lp_pad
static int upload_folded_taps(pysdr_bank *b, cudaStream_t st) {
    const pysdr_bank_config &c = b->cfg;
    std::vector<float2> g((size_t)c.n_rx * c.up * b->lp_pad, make_float2(0.f, 0.f));
    for (int r = 0; r < c.n_rx; ++r) {
        const std::vector<float> &h = b->dec_h[r];
        if (h.empty()) { pysdr_set_error("receiver %d has no resampler taps (set_dec_taps)", r); return PYSDR_ERR_STATE; }
        for (int p = 0; p < c.up; ++p)
            for (int j = 0; j < b->lp; ++j) {
                const size_t k = (size_t)p + (size_t)j * c.up;
                if (k >= h.size()) continue;
                const u64 ph = b->inc[r] * (u64)j;
                const double cyc = (double)(int64_t)ph / 18446744073709551616.0;
                const double ang = 2.0 * M_PI * cyc;
                g[((size_t)r * c.up + p) * b->lp_pad + j] = make_float2((float)(h[k] * cos(ang)), (float)(h[k] * sin(ang)));
            }
    }
    CUDA_TRY(cudaMemcpyAsync(b->d_g, g.data(), sizeof(float2) * g.size(), cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaStreamSynchronize(st));      // g is a stack-owned staging vector
    if (b->mma) { int rc = k1_mma_upload_taps(b->mma, g.data(), b->lp_pad, st); if (rc) return rc; }
    if (b->chan && !b->k1_external) {        // (a bank fed by the raster channelizer never runs its own K1: no 8 MB tap image for it)
        int rc = k1_chan_upload_taps(b->chan, g.data(), b->lp_pad, st);
        if (rc) return rc;
    }
    b->g_dirty = false;
    return PYSDR_OK;
}

template <int KIND>
static int launch_fir(pysdr_bank *b, int rx, const void *src, i64 n_out, float *out, int cw, i64 m0, cudaStream_t st) {
    const int L = b->cfg.af_len;
    const size_t src_sz = (KIND == 0 ? sizeof(float) : sizeof(float2)) * (size_t)(FIR_TILE + L);
    const size_t tap_sz = (KIND == 1 ? sizeof(float2) : sizeof(float)) * (size_t)L;
    const size_t smem = src_sz + tap_sz + 16;
    if (smem > 48 * 1024) {
        CUDA_TRY(cudaFuncSetAttribute(af_fir_kernel<KIND>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    }
    const unsigned grid = (unsigned)((n_out + FIR_TILE - 1) / FIR_TILE);
    af_fir_kernel<KIND><<<grid, FIR_THREADS, smem, st>>>(src, b->d_af + (size_t)rx * L, L, n_out, out, cw,
                                                         b->bfo_inc[rx], m0);
    LAUNCH_CHECK();
    b->launches++;
    return PYSDR_OK;
}

// Standalone streaming-FIR building block (dsp.convolver.convolve_fast, reference receiver.py:862,216):
// out[i] = sum_j taps[j] * src[i + (L-1) - j], src holds L-1 history samples followed by n new ones.
extern "C" int pysdr_fir_valid(const void *d_src, int src_is_complex, const float *taps_host, int L, int64_t n,
                               void *d_out, void *stream) {
    if (!d_src || !taps_host || !d_out || L < 1 || L > 8192 || n < 0) { pysdr_set_error("fir_valid: bad arguments"); return PYSDR_ERR_ARG; }
    if (n == 0) return PYSDR_OK;
    cudaStream_t st = (cudaStream_t)stream;
    std::vector<float2> t(L);
    for (int j = 0; j < L; ++j) t[j] = make_float2(taps_host[j], 0.f);
    float2 *d_t = nullptr;
    CUDA_TRY(cudaMallocAsync(&d_t, sizeof(float2) * L, st));
    CUDA_TRY(cudaMemcpyAsync(d_t, t.data(), sizeof(float2) * L, cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    const size_t smem = (src_is_complex ? sizeof(float2) : sizeof(float)) * (size_t)(FIR_TILE + L) + sizeof(float) * (size_t)L + 16;
    const unsigned grid = (unsigned)((n + FIR_TILE - 1) / FIR_TILE);
    if (src_is_complex) {
        if (smem > 48 * 1024) CUDA_TRY(cudaFuncSetAttribute(af_fir_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        af_fir_kernel<2><<<grid, FIR_THREADS, smem, st>>>(d_src, d_t, L, n, (float *)d_out, 0, 0ull, 0);
    } else {
        if (smem > 48 * 1024) CUDA_TRY(cudaFuncSetAttribute(af_fir_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        af_fir_kernel<0><<<grid, FIR_THREADS, smem, st>>>(d_src, d_t, L, n, (float *)d_out, 0, 0ull, 0);
    }
    LAUNCH_CHECK();
    CUDA_TRY(cudaFreeAsync(d_t, st));
    return PYSDR_OK;
}

extern "C" int pysdr_bank_process_front(pysdr_bank *b, const void *d_iq, int64_t n_in, int halo_in_place,
                                        void *d_iq_bb, int64_t out_stride, float *d_peaks, int64_t *n_out_p,
                                        void *stream) {
    if (!b || !d_iq || n_in < 1 || !d_peaks) { pysdr_set_error("process: bad arguments"); return PYSDR_ERR_ARG; }
    const pysdr_bank_config &c = b->cfg;
    cudaStream_t st = (cudaStream_t)stream;
    if (n_in > c.max_in) { pysdr_set_error("process: n_in=%lld exceeds max_in=%lld", (i64)n_in, (i64)c.max_in); return PYSDR_ERR_CAPACITY; }
    if (!b->k1_only && b->n0 % c.in_chunk != 0) {        // a K1-only bank has no per-block stages: any position
        pysdr_set_error("process: stream position %lld is not on an IN_CHUNK_SIZE=%lld boundary", b->n0, (i64)c.in_chunk);
        return PYSDR_ERR_ALIGN;
    }
    for (int r = 0; r < c.n_rx; ++r)
        if (!b->demod_set[r] && !b->k1_only) { pysdr_set_error("receiver %d has no demodulator (set_demod)", r); return PYSDR_ERR_STATE; }
    if (b->g_dirty) { int rc = upload_folded_taps(b, st); if (rc) return rc; }

    const i64 m0 = n_out_total(b->n0, c.up, c.down);
    const i64 n_out = n_out_total(b->n0 + n_in, c.up, c.down) - m0;
    if (n_out > b->max_out || (d_iq_bb && n_out > out_stride)) {
        pysdr_set_error("process: n_out=%lld exceeds capacity/out_stride", n_out);
        return PYSDR_ERR_CAPACITY;
    }
    const i64 B0 = b->n0 / c.in_chunk;
    const i64 n_blocks = (n_in + c.in_chunk - 1) / c.in_chunk;

    // a pending seek() rides this call when K1 can absorb it (tap-stationary variant, audio stages present, no carrier PLL
    // state to clear): K1 reads a zero history and clears the carried complex memory; the AGC restart rides the back kernel
    bool any_sync = false;
    for (int r = 0; r < c.n_rx; ++r) any_sync = any_sync || b->mode[r] == PYSDR_MODE_AMSYNC;
    const bool k1_is_fast = !b->force_generic && k1_fast_supported(c.up, c.down, b->lp, c.n_rx);
    const bool fold_seek = b->lazy_seek && k1_is_fast && !b->k1_external && !b->k1_only && !any_sync && n_out > 0 && !b->force_unfused;
    if (b->lazy_seek && !fold_seek) { int rc0 = flush_seek(b, st); if (rc0) return rc0; }
    if (fold_seek) b->lazy_seek = false;

    K1Args a;
    a.x = (const float2 *)d_iq;
    a.hist = halo_in_place ? a.x - b->need : (fold_seek ? nullptr : b->d_hist);
    a.zero_c_hist = fold_seek ? 1 : 0;
    a.real_input = b->real_input ? 1 : 0;
    a.need = b->need;
    a.n0 = b->n0; a.n_in = n_in; a.m0 = m0; a.n_out = n_out;
    a.up = c.up; a.down = c.down; a.lp = b->lp; a.lp_pad = b->lp_pad; a.n_rx = c.n_rx;
    a.g = b->d_g;
    for (int r = 0; r < PYSDR_MAX_RX; ++r) { a.acc[r] = b->acc0[r] + b->inc[r] * (u64)b->n0; a.inc[r] = b->inc[r]; }
    a.c_out = b->d_C; a.c_stride = b->c_stride; a.hc = b->hc;
    a.bb_out = (float2 *)d_iq_bb; a.bb_stride = out_stride;
    int rc;
    auto mark = [&](void) -> int {
        if (!b->timing) return PYSDR_OK;
        cudaEvent_t e;
        CUDA_TRY(cudaEventCreate(&e));
        CUDA_TRY(cudaEventRecord(e, st));
        b->evs.push_back(e);
        return PYSDR_OK;
    };
    if ((rc = mark())) return rc;
    int chan_used = 0;
    if (!b->k1_external && !b->force_generic && b->chan && b->mma_mode > 0) {      // a bank of many channels: dense tensor-core contraction
        int nl = 0;
        rc = k1_launch_chan(b->chan, a, b->mma_mode >= 2 ? 1 : K1_CHAN_MIN_ROWS, st, &chan_used, &nl);
        if (rc) return rc;
        b->launches += nl;
    }
    if (b->k1_external) {
        rc = PYSDR_OK;                       // C[r][hc .. hc+n_out) was written by the caller on this stream
    } else if (chan_used) {
        rc = PYSDR_OK;
        b->k1_last = 3;
    } else if (!b->force_generic && k1_fast_supported(c.up, c.down, b->lp, c.n_rx)) {
        int used = 0, nl = 0;
        rc = PYSDR_OK;
        if (b->mma && b->mma_mode > 0) {
            rc = k1_launch_mma(b->mma, a, b->mma_mode >= 2 ? 1 : K1_MMA_MIN_ROWS, st, &used, &nl);
            b->launches += nl;
        }
        if (!rc && !used) {
            rc = k1_launch_fast(a, st);
            b->launches += k1_fast_groups(b->lp, c.n_rx);
        }
        b->k1_last = used ? 2 : 1;
    } else {
        rc = k1_launch_generic(a, st);
        b->launches++;
        b->k1_last = 0;
    }
    if (rc) return rc;
    if ((rc = mark())) return rc;

    if (n_out == 0 || b->k1_only) {
        if (b->need > 0) {                   // input memory for the next call (the audio-rate path does it in state_update)
            hist_update_kernel<<<1, 1024, 0, st>>>(b->d_hist, a.hist, a.x, b->need, n_in);
            LAUNCH_CHECK();
            b->launches++;
        }          // ragged tail too short to emit a sample (only the input memory moves),
                                             // or a K1-only bank (WFM video stage): no audio-rate stages
        b->pend_n_out = 0; b->pend_m0 = m0; b->pend_B0 = B0; b->pend_blocks = n_blocks;
        b->pend_peaks = d_peaks;
        b->pending = true;
        b->n0 += n_in;
        if (n_out_p) *n_out_p = n_out;
        if ((rc = mark())) return rc;
        return PYSDR_OK;
    }
    const int L = c.af_len;
    const bool use_fft = b->d_H && !b->force_direct_fir;
    {
        PllModes pm;
        bool any = false;
        for (int r = 0; r < PYSDR_MAX_RX; ++r) { pm.mode[r] = b->mode[r]; any = any || (r < c.n_rx && b->mode[r] == PYSDR_MODE_AMSYNC); }
        if (any) {
            am_pll_kernel<<<c.n_rx, 32, 0, st>>>(b->d_C, (i64)b->c_stride, b->hc, n_out, pm, b->d_pll, b->pll_k1, b->pll_k2);
            LAUNCH_CHECK();
            b->launches++;
        }
    }
    if (use_fft) {
        // K2 fast path: fused detection + overlap-save AF filter for all receivers in one launch
        const int nfft = fftconv_n_for(L);
        for (int r = 0; r < c.n_rx; ++r) {
            if (!b->h_dirty[r]) continue;
            rc = fftconv_prepare_taps(b->d_af + (size_t)r * L, L, b->d_H + (size_t)r * nfft, st);
            if (rc) return rc;
            b->launches++;
            b->h_dirty[r] = false;
        }
        FftConvArgs f;
        f.C = b->d_C; f.c_stride = b->c_stride; f.H = b->d_H; f.L = L; f.n_out = n_out; f.m0 = m0;
        f.out = (float *)b->d_a; f.a_stride = b->a_stride;
        for (int r = 0; r < PYSDR_MAX_RX; ++r) { f.mode[r] = b->mode[r]; f.bfo_inc[r] = b->bfo_inc[r]; }
        rc = fftconv_launch(f, c.n_rx, st);
        if (rc) return rc;
        b->launches++;
    }
    for (int r = 0; r < c.n_rx; ++r) {
        float2 *C = b->d_C + (size_t)r * b->c_stride;
        float *R = b->d_R + (size_t)r * b->r_stride;
        float *aout = (float *)(b->d_a + (size_t)r * b->a_stride);
        const int mode = b->mode[r];
        if (mode == PYSDR_MODE_RAW) {                 // WFM second stage: a = Re{resampler output}
            i64 blocks = (n_out + 255) / 256;
            if (blocks > (i64)pysdr_sm_count() * 8) blocks = (i64)pysdr_sm_count() * 8;
            real_part_kernel<<<(unsigned)blocks, 256, 0, st>>>(C + b->hc, aout, n_out);
            LAUNCH_CHECK();
            b->launches++;
            rc = PYSDR_OK;
        } else if (use_fft) {
            rc = PYSDR_OK;
        } else if (mode == PYSDR_MODE_AM || mode == PYSDR_MODE_NFM || mode == PYSDR_MODE_AMSYNC) {
            const i64 nr = (L - 1) + n_out;
            i64 blocks = (nr + 255) / 256;
            if (blocks > (i64)pysdr_sm_count() * 8) blocks = (i64)pysdr_sm_count() * 8;
            detect_kernel<<<(unsigned)blocks, 256, 0, st>>>(C, R, nr, mode == PYSDR_MODE_NFM, mode == PYSDR_MODE_AMSYNC);
            LAUNCH_CHECK();
            b->launches++;
            rc = launch_fir<0>(b, r, R, n_out, aout, 0, m0, st);
        } else if (mode == PYSDR_MODE_USB || mode == PYSDR_MODE_LSB) {
            rc = launch_fir<1>(b, r, C + 2, n_out, aout, 0, m0, st);
        } else {
            rc = launch_fir<2>(b, r, C + 2, n_out, aout, mode == PYSDR_MODE_CW, m0, st);
        }
        if (rc) return rc;
        (void)aout;
    }
    if (b->stereo) {
        for (int r = 0; r < 3; ++r)
            if (b->mode[r] != PYSDR_MODE_IQ) { pysdr_set_error("stereo bank: all three rows must be in IQ mode"); return PYSDR_ERR_STATE; }
        i64 blocks = (n_out + 255) / 256;
        if (blocks > (i64)pysdr_sm_count() * 8) blocks = (i64)pysdr_sm_count() * 8;
        float *Ls = b->d_R, *Rs = b->d_R + b->r_stride;
        stereo_matrix_kernel<<<(unsigned)blocks, 256, 0, st>>>(b->d_a, b->d_a + b->a_stride, b->d_a + 2 * b->a_stride, n_out,
                                                               b->pilot_min, Ls, Rs);
        LAUNCH_CHECK();
        copy_f32_kernel<<<(unsigned)blocks, 256, 0, st>>>(Ls, (float *)b->d_a, n_out);
        LAUNCH_CHECK();
        copy_f32_kernel<<<(unsigned)blocks, 256, 0, st>>>(Rs, (float *)(b->d_a + b->a_stride), n_out);
        LAUNCH_CHECK();
        b->launches += 3;
    }
    {
        // block peaks (IQ-mode rows produce unused values) + the end-of-call state update: complex memory
        // C[0..hc) <- C[n_out .. n_out+hc) and the raw input memory, for the next call.  When back follows at once
        // (pysdr_bank_process) both ride the fused back kernel instead of a launch of their own.
        StateArgs sa;
        sa.C = b->d_C; sa.c_stride = b->c_stride; sa.n_out = n_out; sa.hc = b->hc; sa.n_rx = c.n_rx;
        sa.hist = b->d_hist; sa.hist_src = a.hist; sa.x = a.x; sa.need = b->need; sa.n_in = n_in; sa.enabled = 1;
        if (b->defer_peaks && !b->stereo && !b->force_unfused) {
            b->pend_sa = sa;
            b->peaks_deferred = true;
        } else {
            b->peaks_deferred = false;
            rc = launch_block_peaks(b, d_peaks, sa, n_blocks, B0, m0, n_out, st);
            if (rc) return rc;
        }
    }
    if (b->stereo) {
        peak_link_kernel<<<(unsigned)((n_blocks + 255) / 256), 256, 0, st>>>(d_peaks, d_peaks + n_blocks, n_blocks);
        LAUNCH_CHECK();
        b->launches++;
    }

    if ((rc = mark())) return rc;
    b->pend_n_out = n_out; b->pend_m0 = m0; b->pend_B0 = B0; b->pend_blocks = n_blocks;
    b->pend_peaks = d_peaks;
    b->pending = true;
    b->n0 += n_in;
    if (n_out_p) *n_out_p = n_out;
    return PYSDR_OK;
}

static int launch_block_peaks(pysdr_bank *b, float *d_peaks, const StateArgs &sa, i64 n_blocks, i64 B0, i64 m0, i64 n_out,
                              cudaStream_t st) {
    const pysdr_bank_config &c = b->cfg;
    unsigned gx = (unsigned)((n_blocks + BLK_WARPS - 1) / BLK_WARPS);
    if (gx < (unsigned)c.n_rx + 1) gx = (unsigned)c.n_rx + 1;
    dim3 grid(gx, (unsigned)c.n_rx + 1);
    block_peak_kernel<<<grid, 32 * BLK_WARPS, 0, st>>>((const float *)b->d_a, 2 * b->a_stride, d_peaks, n_blocks, n_blocks, B0,
                                                        c.in_chunk, c.up, c.down, m0, n_out, c.n_rx, sa);
    LAUNCH_CHECK();
    b->launches++;
    return PYSDR_OK;
}

// back = AGC entry state (optional) -> [deferred block peaks] -> scan -> gain / DC removal.  One fused launch unless the
// bank is the stereo WFM2 resampler (its L/R peaks are linked between the stages) or the stand-alone kernels are forced.
static int back_impl(pysdr_bank *b, const float *d_prev_peaks, int64_t n_prev, int64_t skip_blocks, const double *d_sums,
                     int n_before, bool enter, float *d_am, float *d_am_dc, int64_t out_stride, cudaStream_t st,
                     const XchgWait *xw = nullptr, bool push = false) {
    if (!b || !b->pending) { pysdr_set_error("process_back without process_front"); return PYSDR_ERR_STATE; }
    if (!d_am) { pysdr_set_error("process_back: d_am is null"); return PYSDR_ERR_ARG; }
    const pysdr_bank_config &c = b->cfg;
    const i64 n_out = b->pend_n_out, n_blocks = b->pend_blocks;
    if (n_out > out_stride) { pysdr_set_error("process_back: out_stride too small"); return PYSDR_ERR_CAPACITY; }
    if (enter) b->lazy_reset_agc = false;          // an explicit entry state replaces the restart a seek(0) asked for
    if (n_out == 0) {                        // nothing was emitted: the AGC does not advance (like the reference's empty am)
        if (enter) {
            if (xw) agc_enter_xchg_kernel<<<1, PYSDR_MAX_RX, 0, st>>>(b->d_agc, c.n_rx, *xw);
            else agc_enter_kernel<<<1, PYSDR_MAX_RX, 0, st>>>(b->d_agc, d_sums, n_before, c.n_rx);
            LAUNCH_CHECK();
            b->launches++;
        }
        if (b->timing) {
            for (int k = 0; k < 2; ++k) {
                cudaEvent_t e;
                CUDA_TRY(cudaEventCreate(&e));
                CUDA_TRY(cudaEventRecord(e, st));
                b->evs.push_back(e);
            }
        }
        b->pending = false;
        b->peaks_deferred = false;
        return PYSDR_OK;
    }
    if (skip_blocks < 0 || skip_blocks >= n_blocks) { pysdr_set_error("process_back: bad skip_blocks"); return PYSDR_ERR_ARG; }
    if (!enter && b->lazy_reset_agc) {       // seek(0) folded into this call: the AGCs restart (the peak replay restarts them itself)
        enter = n_prev <= 0;
        d_sums = nullptr;
        n_before = 0;
        b->lazy_reset_agc = false;
    }
    AgcScanArgs s;
    s.state = b->d_agc;
    s.peaks = b->pend_peaks; s.peaks_stride = n_blocks;
    s.prev_peaks = (n_prev > 0) ? d_prev_peaks : nullptr; s.n_prev = n_prev;
    s.gains = b->d_gains; s.gains_stride = b->max_blocks;
    s.n_blocks = n_blocks; s.n_rx = c.n_rx; s.skip = skip_blocks;
    for (int r = 0; r < PYSDR_MAX_RX; ++r) s.enabled[r] = (r < c.n_rx && (b->mode[r] != PYSDR_MODE_IQ || (b->stereo && r < 2))) ? 1 : 0;
    ApplyKinds kinds;
    bool any_real = false;
    for (int r = 0; r < PYSDR_MAX_RX; ++r) {
        kinds.kind[r] = 0;
        if (r >= c.n_rx) continue;
        if (b->mode[r] == PYSDR_MODE_IQ && !(b->stereo && r < 2)) kinds.kind[r] = 3;
        else {
            kinds.kind[r] = (b->mode[r] == PYSDR_MODE_AM || b->mode[r] == PYSDR_MODE_USB) ? 2 : 1;
            any_real = true;
        }
    }
    if (b->timing) {
        cudaEvent_t e;
        CUDA_TRY(cudaEventCreate(&e));
        CUDA_TRY(cudaEventRecord(e, st));
        b->evs.push_back(e);
    }
    const bool fuse = !b->force_unfused && !b->stereo;
    if (fuse) {
        if (b->back_grid <= 0) {
            int occ = 0;
            CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, agc_back_fused_kernel, BACK_THREADS, 0));
            if (occ < 1) { pysdr_set_error("fused back kernel does not fit on an SM"); return PYSDR_ERR_CUDA; }
            b->back_grid = pysdr_sm_count() * (occ > 4 ? 4 : occ);         // every CTA resident at once: the grid barrier needs it
        }
        BackArgs p;
        memset(&p, 0, sizeof(p));
        p.do_peaks = b->peaks_deferred ? 1 : 0;
        p.do_enter = enter ? (xw ? 2 : 1) : 0; p.n_before = n_before; p.sums = d_sums; p.n_rx = c.n_rx;
        if (xw) p.x = *xw;
        p.do_push = (push && xw) ? 1 : 0; p.push_skip = skip_blocks;
        p.a = (const float *)b->d_a; p.a_row = 2 * b->a_stride;
        p.peaks = (float *)b->pend_peaks; p.peaks_row = n_blocks;
        p.n_blocks = n_blocks; p.B0 = b->pend_B0; p.in_chunk = c.in_chunk; p.m0 = b->pend_m0; p.n_out = n_out;
        p.up = c.up; p.down = c.down;
        if (b->peaks_deferred) p.sa = b->pend_sa;
        p.scan = s;
        p.gains = b->d_gains; p.g_row = b->max_blocks;
        p.am = d_am; p.am_dc = d_am_dc; p.am_row = 2 * out_stride;
        p.kinds = kinds;
        p.bar = b->d_bar; p.bar_base = b->bar_count;
        i64 want = ((i64)c.n_rx * n_blocks + BACK_THREADS / 32 - 1) / (BACK_THREADS / 32);
        if (want < c.n_rx + 2) want = c.n_rx + 2;                          // scans + the state-update items
        const int grid = (int)(want < b->back_grid ? want : b->back_grid);
        CUDA_TRY(launch_pdl(agc_back_fused_kernel, dim3(grid), dim3(BACK_THREADS), 0, st, p));
        b->bar_count += (unsigned long long)grid * (p.do_peaks ? 2ull : 1ull);
        b->launches++;
        b->peaks_deferred = false;
    } else {
        if (b->peaks_deferred) {
            int rc = launch_block_peaks(b, (float *)b->pend_peaks, b->pend_sa, n_blocks, b->pend_B0, b->pend_m0, n_out, st);
            if (rc) return rc;
            b->peaks_deferred = false;
        }
        if (enter) {
            if (xw) agc_enter_xchg_kernel<<<1, PYSDR_MAX_RX, 0, st>>>(b->d_agc, c.n_rx, *xw);
            else agc_enter_kernel<<<1, PYSDR_MAX_RX, 0, st>>>(b->d_agc, d_sums, n_before, c.n_rx);
            LAUNCH_CHECK();
            b->launches++;
        }
        agc_scan_kernel<<<c.n_rx, AGC_THREADS, 0, st>>>(s);
        LAUNCH_CHECK();
        b->launches++;
        for (int r = 0; r < c.n_rx; ++r) {
            if (kinds.kind[r] != 3) continue;
            const float *aout = (const float *)(b->d_a + (size_t)r * b->a_stride);
            float *am = d_am + (size_t)r * 2 * out_stride;
            float *amdc = d_am_dc ? d_am_dc + (size_t)r * 2 * out_stride : nullptr;
            i64 n = 2 * n_out, blocks = (n + 255) / 256;
            if (blocks > (i64)pysdr_sm_count() * 8) blocks = (i64)pysdr_sm_count() * 8;
            copy_f32_kernel<<<(unsigned)blocks, 256, 0, st>>>(aout, am, n);
            LAUNCH_CHECK();
            b->launches++;
            if (amdc) {
                copy_f32_kernel<<<(unsigned)blocks, 256, 0, st>>>(aout, amdc, n);
                LAUNCH_CHECK();
                b->launches++;
            }
        }
        if (any_real) {
            dim3 grid((unsigned)((n_blocks + BLK_WARPS - 1) / BLK_WARPS), (unsigned)c.n_rx);
            agc_apply_kernel<<<grid, 32 * BLK_WARPS, 0, st>>>((const float *)b->d_a, 2 * b->a_stride, b->d_gains, b->max_blocks, d_am,
                                                               d_am_dc, 2 * out_stride, kinds, n_blocks, b->pend_B0, c.in_chunk, c.up,
                                                               c.down, b->pend_m0, n_out);
            LAUNCH_CHECK();
            b->launches++;
        }
    }
    if (b->timing) {
        cudaEvent_t e;
        CUDA_TRY(cudaEventCreate(&e));
        CUDA_TRY(cudaEventRecord(e, st));
        b->evs.push_back(e);
    }
    b->pending = false;
    return PYSDR_OK;
}

extern "C" int pysdr_bank_process_back(pysdr_bank *b, const float *d_prev_peaks, int64_t n_prev, int64_t skip_blocks,
                                       float *d_am, float *d_am_dc, int64_t out_stride, void *stream) {
    return back_impl(b, d_prev_peaks, n_prev, skip_blocks, nullptr, 0, false, d_am, d_am_dc, out_stride, (cudaStream_t)stream);
}

extern "C" int pysdr_bank_process_back_carry(pysdr_bank *b, const double *d_summaries, int n_before, int64_t skip_blocks,
                                             float *d_am, float *d_am_dc, int64_t out_stride, void *stream) {
    if (n_before < 0 || (n_before > 0 && !d_summaries)) { pysdr_set_error("process_back_carry: bad arguments"); return PYSDR_ERR_ARG; }
    return back_impl(b, nullptr, 0, skip_blocks, d_summaries, n_before, true, d_am, d_am_dc, out_stride, (cudaStream_t)stream);
}

extern "C" int64_t pysdr_xchg_bytes(int world, int n_rx) {
    if (world < 1 || world > PYSDR_XCHG_MAX_WORLD || n_rx < 1 || n_rx > PYSDR_MAX_RX) return -1;
    return (int64_t)(sizeof(double) * xchg_flags_off(world, n_rx) + sizeof(unsigned long long) * ((size_t)XCHG_DEPTH * world + world + 8));
}

static int xchg_peers(const uint64_t *peer_bases, int world, int rank, uint64_t seq, XchgPeers *out, const char *who) {
    if (!peer_bases || world < 1 || world > PYSDR_XCHG_MAX_WORLD || rank < 0 || rank >= world || seq == 0) {
        pysdr_set_error("%s: bad arguments", who);
        return PYSDR_ERR_ARG;
    }
    for (int q = 0; q < PYSDR_XCHG_MAX_WORLD; ++q) out->base[q] = q < world ? (double *)(uintptr_t)peer_bases[q] : nullptr;
    return PYSDR_OK;
}

extern "C" int pysdr_bank_agc_summary_push(pysdr_bank *b, int64_t skip_blocks, const uint64_t *peer_bases, int world, int rank,
                                           uint64_t seq, void *stream) {
    XchgPeers peers;
    if (!b) return PYSDR_ERR_ARG;
    if (int rc = xchg_peers(peer_bases, world, rank, seq, &peers, "agc_summary_push")) return rc;
    if (!b->pending || !b->pend_peaks) { pysdr_set_error("agc_summary_push: call process_front first"); return PYSDR_ERR_STATE; }
    if (skip_blocks < 0 || skip_blocks >= b->pend_blocks) { pysdr_set_error("agc_summary_push: bad skip_blocks"); return PYSDR_ERR_ARG; }
    agc_summary_push_kernel<<<b->cfg.n_rx, SUM_THREADS, 0, (cudaStream_t)stream>>>(b->d_agc, b->pend_peaks, b->pend_blocks, skip_blocks,
                                                                                   b->pend_blocks, peers, world, rank, b->cfg.n_rx, seq);
    LAUNCH_CHECK();
    b->launches++;
    return PYSDR_OK;
}

extern "C" int pysdr_bank_process_back_xchg(pysdr_bank *b, const uint64_t *peer_bases, int world, int rank, uint64_t seq,
                                            int64_t skip_blocks, float *d_am, float *d_am_dc, int64_t out_stride, void *stream) {
    XchgWait xw;
    if (!b) return PYSDR_ERR_ARG;
    if (int rc = xchg_peers(peer_bases, world, rank, seq, &xw.peers, "process_back_xchg")) return rc;
    xw.world = world; xw.rank = rank; xw.seq = seq;
    return back_impl(b, nullptr, 0, skip_blocks, nullptr, rank, true, d_am, d_am_dc, out_stride, (cudaStream_t)stream, &xw);
}

// One time shard in three launches (K1, AF filter, fused back): the back kernel computes the block peaks, pushes this rank's
// AGC summaries to the later ranks, waits for the earlier ranks' summaries, scans and applies the gains.
extern "C" int pysdr_bank_process_shard_xchg(pysdr_bank *b, const void *d_iq, int64_t n_in, int halo_in_place, void *d_iq_bb,
                                             const uint64_t *peer_bases, int world, int rank, uint64_t seq, int64_t skip_blocks,
                                             float *d_am, float *d_am_dc, int64_t out_stride, int64_t *n_out, void *stream) {
    XchgWait xw;
    if (!b) { pysdr_set_error("null bank"); return PYSDR_ERR_ARG; }
    if (int rc = xchg_peers(peer_bases, world, rank, seq, &xw.peers, "process_shard_xchg")) return rc;
    xw.world = world; xw.rank = rank; xw.seq = seq;
    if (b->stereo || b->force_unfused) { pysdr_set_error("process_shard_xchg needs the fused back kernel"); return PYSDR_ERR_STATE; }
    b->defer_peaks = true;
    int rc = pysdr_bank_process_front(b, d_iq, n_in, halo_in_place, d_iq_bb, out_stride, b->d_peaks, n_out, stream);
    b->defer_peaks = false;
    if (rc) return rc;
    if (b->pend_n_out == 0) { pysdr_set_error("process_shard_xchg: the shard produced no output"); return PYSDR_ERR_ARG; }
    return back_impl(b, nullptr, 0, skip_blocks, nullptr, rank, true, d_am, d_am_dc, out_stride, (cudaStream_t)stream, &xw, true);
}

extern "C" int pysdr_bank_force_unfused(pysdr_bank *b, int on) {
    if (!b) return PYSDR_ERR_ARG;
    b->force_unfused = on != 0;
    return PYSDR_OK;
}

extern "C" int pysdr_bank_agc_summary(pysdr_bank *b, int64_t skip_blocks, double *d_summary, void *stream) {
    if (!b || !d_summary) { pysdr_set_error("agc_summary: bad arguments"); return PYSDR_ERR_ARG; }
    if (!b->pending || !b->pend_peaks) { pysdr_set_error("agc_summary: call process_front first"); return PYSDR_ERR_STATE; }
    if (skip_blocks < 0 || skip_blocks >= b->pend_blocks) { pysdr_set_error("agc_summary: bad skip_blocks"); return PYSDR_ERR_ARG; }
    agc_summary_kernel<<<b->cfg.n_rx, SUM_THREADS, 0, (cudaStream_t)stream>>>(b->d_agc, b->pend_peaks, b->pend_blocks, skip_blocks,
                                                                              b->pend_blocks, d_summary);
    LAUNCH_CHECK();
    b->launches++;
    return PYSDR_OK;
}

extern "C" int pysdr_bank_agc_enter(pysdr_bank *b, const double *d_summaries, int n_before, void *stream) {
    if (!b || n_before < 0 || (n_before > 0 && !d_summaries)) { pysdr_set_error("agc_enter: bad arguments"); return PYSDR_ERR_ARG; }
    b->lazy_reset_agc = false;               // this entry state replaces the restart a seek(0) asked for ...
    { int rc = flush_seek(b, (cudaStream_t)stream); if (rc) return rc; }     // ... and a pending seek must not undo it afterwards
    agc_enter_kernel<<<1, PYSDR_MAX_RX, 0, (cudaStream_t)stream>>>(b->d_agc, d_summaries, n_before, b->cfg.n_rx);
    LAUNCH_CHECK();
    b->launches++;
    return PYSDR_OK;
}

extern "C" int pysdr_bank_set_timing(pysdr_bank *b, int on) {
    if (!b) return PYSDR_ERR_ARG;
    for (cudaEvent_t e : b->evs) cudaEventDestroy(e);
    b->evs.clear();
    b->timing = on != 0;
    return PYSDR_OK;
}

// out4 = { sum K1 ms, sum rest-of-front ms, sum back ms, # of process calls }; clears the record.
extern "C" int pysdr_bank_get_timing(pysdr_bank *b, double out4[4], void *stream) {
    if (!b || !out4) return PYSDR_ERR_ARG;
    CUDA_TRY(cudaStreamSynchronize((cudaStream_t)stream));
    out4[0] = out4[1] = out4[2] = out4[3] = 0.0;
    const size_t n = b->evs.size() / 5;
    for (size_t i = 0; i < n; ++i) {
        float k1 = 0.f, fr = 0.f, bk = 0.f;
        CUDA_TRY(cudaEventElapsedTime(&k1, b->evs[5 * i], b->evs[5 * i + 1]));
        CUDA_TRY(cudaEventElapsedTime(&fr, b->evs[5 * i + 1], b->evs[5 * i + 2]));
        CUDA_TRY(cudaEventElapsedTime(&bk, b->evs[5 * i + 3], b->evs[5 * i + 4]));
        out4[0] += k1; out4[1] += fr; out4[2] += bk; out4[3] += 1.0;
        if (getenv("PYSDR_TIMING_TRACE")) fprintf(stderr, "timing step %zu: k1 %.4f front_rest %.4f back %.4f ms\n", i, k1, fr, bk);
    }
    for (cudaEvent_t e : b->evs) cudaEventDestroy(e);
    b->evs.clear();
    return PYSDR_OK;
}

extern "C" int pysdr_bank_process(pysdr_bank *b, const void *d_iq, int64_t n_in, int halo_in_place, void *d_iq_bb,
                                  float *d_am, float *d_am_dc, int64_t out_stride, int64_t *n_out, void *stream) {
    if (!b) { pysdr_set_error("null bank"); return PYSDR_ERR_ARG; }
    b->defer_peaks = true;                   // back follows at once: block peaks + state update ride its fused launch
    int rc = pysdr_bank_process_front(b, d_iq, n_in, halo_in_place, d_iq_bb, out_stride, b->d_peaks, n_out, stream);
    b->defer_peaks = false;
    if (rc) return rc;
    return pysdr_bank_process_back(b, nullptr, 0, 0, d_am, d_am_dc, out_stride, stream);
}

// ---- host-chunk executive -------------------------------------------------------------------------------------------------
// The call the reference's loop makes per chunk (receiver.py:724-725 -> demodulate_data -> rx.demod_data(x)): a HOST
// chunk in, HOST results out, for every receiver of the bank at once.  Everything between the two host buffers happens in
// this one C call on `stream`: upload (directly from the caller's buffer when it is page-locked, else through a pinned
// staging copy), the three step kernels, the result rows downloaded into ONE pinned block, one synchronisation.
// r01 did this from Python with ~14 tensor operations per chunk (0.33 ms per 21 ms chunk for 4 receivers).
extern "C" int pysdr_bank_process_host(pysdr_bank *b, const void *h_iq, int64_t n_in, int want_iq, int want_dc, float **h_am,
                                       void **h_iq_bb, float **h_am_dc, int64_t *row_floats, int64_t *n_out_p, void *stream) {
    if (!b || !h_iq || n_in < 1 || !h_am || !row_floats || !n_out_p) { pysdr_set_error("process_host: bad arguments"); return PYSDR_ERR_ARG; }
    const pysdr_bank_config &c = b->cfg;
    if (n_in > c.max_in) { pysdr_set_error("process_host: n_in=%lld exceeds max_in=%lld", (i64)n_in, (i64)c.max_in); return PYSDR_ERR_CAPACITY; }
    cudaStream_t st = (cudaStream_t)stream;
    const size_t row = 2 * (size_t)b->max_out;                                  // floats per result row
    const size_t am_bytes = sizeof(float) * row * c.n_rx;
    if (!b->hc_h_in) {
        CUDA_TRY(cudaHostAlloc(&b->hc_h_in, sizeof(float2) * (size_t)c.max_in, cudaHostAllocDefault));
        CUDA_TRY(cudaHostAlloc(&b->hc_h_out, 3 * am_bytes, cudaHostAllocDefault));
        CUDA_TRY(cudaMalloc(&b->hc_d_in, sizeof(float2) * (size_t)c.max_in));
        CUDA_TRY(cudaMalloc(&b->hc_d_iq, am_bytes));
        // the audio rows are a few KB per receiver and written once, by the last kernel of the step: let it store them
        // straight into the pinned result block over PCIe instead of paying two more copy operations per chunk
        int dev = 0, can = 0;
        void *dp = nullptr;
        b->hc_zero_copy = !getenv("PYSDR_NO_ZERO_COPY") && cudaGetDevice(&dev) == cudaSuccess &&
                          cudaDeviceGetAttribute(&can, cudaDevAttrCanUseHostPointerForRegisteredMem, dev) == cudaSuccess && can &&
                          cudaHostGetDevicePointer(&dp, b->hc_h_out, 0) == cudaSuccess && dp;
        cudaGetLastError();
        if (b->hc_zero_copy) {
            b->hc_d_am = (float *)dp;
            b->hc_d_dc = (float *)((char *)dp + 2 * am_bytes);
        } else {
            CUDA_TRY(cudaMalloc(&b->hc_d_am, am_bytes));
            CUDA_TRY(cudaMalloc(&b->hc_d_dc, am_bytes));
        }
    }
    const void *src = h_iq;
    cudaPointerAttributes pa;
    if (cudaPointerGetAttributes(&pa, h_iq) != cudaSuccess || pa.type != cudaMemoryTypeHost) {
        cudaGetLastError();                                                      // pageable memory: stage through the pinned block
        memcpy(b->hc_h_in, h_iq, sizeof(float2) * (size_t)n_in);
        src = b->hc_h_in;
    }
    CUDA_TRY(cudaMemcpyAsync(b->hc_d_in, src, sizeof(float2) * (size_t)n_in, cudaMemcpyHostToDevice, st));
    bool any_sync = false;
    for (int r = 0; r < c.n_rx; ++r) any_sync = any_sync || b->mode[r] == PYSDR_MODE_AMSYNC;
    int64_t n_out = 0;
    int rc = pysdr_bank_process(b, b->hc_d_in, n_in, 0, any_sync ? (void *)b->hc_d_iq : nullptr, b->hc_d_am, want_dc ? b->hc_d_dc : nullptr,
                                (int64_t)b->max_out, &n_out, stream);
    if (rc) return rc;
    float *o_am = (float *)b->hc_h_out, *o_dc = (float *)(b->hc_h_out + 2 * am_bytes);
    float2 *o_iq = (float2 *)(b->hc_h_out + am_bytes);
    if (n_out > 0) {
        const size_t w = sizeof(float) * 2 * (size_t)n_out;                      // complex rows use all of it, real rows the first half
        if (!b->hc_zero_copy) {
            CUDA_TRY(cudaMemcpy2DAsync(o_am, sizeof(float) * row, b->hc_d_am, sizeof(float) * row, w, c.n_rx, cudaMemcpyDeviceToHost, st));
            if (want_dc)
                CUDA_TRY(cudaMemcpy2DAsync(o_dc, sizeof(float) * row, b->hc_d_dc, sizeof(float) * row, w, c.n_rx, cudaMemcpyDeviceToHost, st));
        }
        if (want_iq) {
            if (any_sync)
                CUDA_TRY(cudaMemcpy2DAsync(o_iq, sizeof(float) * row, b->hc_d_iq, sizeof(float2) * (size_t)b->max_out, w, c.n_rx,
                                           cudaMemcpyDeviceToHost, st));
            else
                CUDA_TRY(cudaMemcpy2DAsync(o_iq, sizeof(float) * row, b->d_C + b->hc, sizeof(float2) * (size_t)b->c_stride, w, c.n_rx,
                                           cudaMemcpyDeviceToHost, st));
        }
    }
    CUDA_TRY(cudaStreamSynchronize(st));
    *h_am = o_am;
    if (h_iq_bb) *h_iq_bb = o_iq;
    if (h_am_dc) *h_am_dc = o_dc;
    *row_floats = (int64_t)row;
    *n_out_p = n_out;
    return PYSDR_OK;
}

extern "C" int pysdr_bank_host_chunk_ptr(pysdr_bank *b, void **d_in) {
    if (!b || !d_in || !b->hc_d_in) { pysdr_set_error("host_chunk_ptr: no host chunk has been processed yet"); return PYSDR_ERR_STATE; }
    *d_in = b->hc_d_in;
    return PYSDR_OK;
}

// A pending seek() is materialised either inside the next process call (K1 clears the carried complex memory and reads a
// zero history, the fused back kernel restarts the AGC: no launch of its own) or, for every other API that looks at the
// carried state, by this flush.
static int flush_seek(pysdr_bank *b, cudaStream_t st) {
    if (b->lazy_seek) {
        seek_reset_kernel<<<b->cfg.n_rx + 1, 256, 0, st>>>(b->d_C, b->c_stride, b->hc, b->cfg.n_rx, b->d_hist, b->need + 8, b->d_pll,
                                                           b->d_agc, b->lazy_reset_agc ? 1 : 0);
        LAUNCH_CHECK();
        b->launches++;
        b->lazy_seek = false;
        b->lazy_reset_agc = false;
    } else if (b->lazy_reset_agc) {              // the seek itself was folded into a front call; the AGC restart is still due
        agc_enter_kernel<<<1, PYSDR_MAX_RX, 0, st>>>(b->d_agc, nullptr, 0, b->cfg.n_rx);
        LAUNCH_CHECK();
        b->launches++;
        b->lazy_reset_agc = false;
    }
    return PYSDR_OK;
}

extern "C" int pysdr_bank_seek(pysdr_bank *b, int64_t n0_abs, void *stream) {
    if (!b || n0_abs < 0 || n0_abs % b->cfg.in_chunk != 0) {
        pysdr_set_error("seek: position must be a non-negative multiple of IN_CHUNK_SIZE");
        return PYSDR_ERR_ALIGN;
    }
    (void)stream;
    b->n0 = n0_abs;
    b->pending = false;
    b->peaks_deferred = false;
    // every carried state goes back to "nothing before this sample"; at the stream origin (n0 = 0) the AGCs restart too
    b->lazy_seek = true;
    b->lazy_reset_agc = b->lazy_reset_agc || n0_abs == 0;
    return PYSDR_OK;
}

// ---- checkpoint ------------------------------------------------------------------------------------
struct StateHeader {
    uint32_t magic, version;
    int32_t n_rx, need, hc, pad;
    i64 n0;
    u64 inc[PYSDR_MAX_RX], acc0[PYSDR_MAX_RX];
    double pll[PYSDR_MAX_RX][2];
};

extern "C" int64_t pysdr_bank_state_size(const pysdr_bank *b) {
    if (!b) return -1;
    return (int64_t)(sizeof(StateHeader) + sizeof(AgcState) * PYSDR_MAX_RX + sizeof(float2) * (size_t)b->need +
                     sizeof(float2) * (size_t)b->cfg.n_rx * b->hc);
}

extern "C" int pysdr_bank_get_state(pysdr_bank *b, void *blob, int64_t size, void *stream) {
    if (!b || !blob || size < pysdr_bank_state_size(b)) { pysdr_set_error("get_state: bad blob"); return PYSDR_ERR_ARG; }
    { int rc = flush_seek(b, (cudaStream_t)stream); if (rc) return rc; }
    CUDA_TRY(cudaStreamSynchronize((cudaStream_t)stream));
    char *p = (char *)blob;
    StateHeader h;
    memset(&h, 0, sizeof(h));
    h.magic = 0x50534452u; h.version = 2; h.n_rx = b->cfg.n_rx; h.need = b->need; h.hc = b->hc; h.n0 = b->n0;
    memcpy(h.inc, b->inc, sizeof(h.inc));
    memcpy(h.acc0, b->acc0, sizeof(h.acc0));
    CUDA_TRY(cudaMemcpy(h.pll, b->d_pll, sizeof(h.pll), cudaMemcpyDeviceToHost));
    memcpy(p, &h, sizeof(h)); p += sizeof(h);
    CUDA_TRY(cudaMemcpy(p, b->d_agc, sizeof(AgcState) * PYSDR_MAX_RX, cudaMemcpyDeviceToHost)); p += sizeof(AgcState) * PYSDR_MAX_RX;
    if (b->need) CUDA_TRY(cudaMemcpy(p, b->d_hist, sizeof(float2) * (size_t)b->need, cudaMemcpyDeviceToHost));
    p += sizeof(float2) * (size_t)b->need;
    for (int r = 0; r < b->cfg.n_rx; ++r) {
        CUDA_TRY(cudaMemcpy(p, b->d_C + (size_t)r * b->c_stride, sizeof(float2) * (size_t)b->hc, cudaMemcpyDeviceToHost));
        p += sizeof(float2) * (size_t)b->hc;
    }
    return PYSDR_OK;
}

extern "C" int pysdr_bank_set_state(pysdr_bank *b, const void *blob, int64_t size, void *stream) {
    if (!b || !blob || size < pysdr_bank_state_size(b)) { pysdr_set_error("set_state: bad blob"); return PYSDR_ERR_ARG; }
    b->lazy_seek = false; b->lazy_reset_agc = false; b->peaks_deferred = false;      // the blob replaces every carried state
    CUDA_TRY(cudaStreamSynchronize((cudaStream_t)stream));
    const char *p = (const char *)blob;
    StateHeader h;
    memcpy(&h, p, sizeof(h)); p += sizeof(h);
    if (h.magic != 0x50534452u || h.version != 2 || h.n_rx != b->cfg.n_rx || h.need != b->need || h.hc != b->hc) {
        pysdr_set_error("set_state: blob does not match this bank's geometry");
        return PYSDR_ERR_STATE;
    }
    b->n0 = h.n0;
    memcpy(b->inc, h.inc, sizeof(h.inc));
    memcpy(b->acc0, h.acc0, sizeof(h.acc0));
    CUDA_TRY(cudaMemcpy(b->d_pll, h.pll, sizeof(h.pll), cudaMemcpyHostToDevice));
    b->g_dirty = true;
    b->pending = false;
    CUDA_TRY(cudaMemcpy(b->d_agc, p, sizeof(AgcState) * PYSDR_MAX_RX, cudaMemcpyHostToDevice)); p += sizeof(AgcState) * PYSDR_MAX_RX;
    if (b->need) CUDA_TRY(cudaMemcpy(b->d_hist, p, sizeof(float2) * (size_t)b->need, cudaMemcpyHostToDevice));
    p += sizeof(float2) * (size_t)b->need;
    for (int r = 0; r < b->cfg.n_rx; ++r) {
        CUDA_TRY(cudaMemcpy(b->d_C + (size_t)r * b->c_stride, p, sizeof(float2) * (size_t)b->hc, cudaMemcpyHostToDevice));
        p += sizeof(float2) * (size_t)b->hc;
    }
    return PYSDR_OK;
}
