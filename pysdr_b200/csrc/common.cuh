// common.cuh — shared declarations for libpysdr_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <string>
#include <type_traits>
#include <vector>

#include "../../include/pysdr_b200.h"

typedef unsigned long long u64;
typedef long long i64;

void pysdr_set_error(const char *fmt, ...);
// multiprocessor count of the CURRENT device (cudaDevAttrMultiProcessorCount, cached per device; 148 on B200)
int pysdr_sm_count(void);
// current device ordinal, or 0 when it cannot be read (used to key per-device one-time set-up)
int pysdr_device(void);

#define CUDA_TRY(expr)                                                                         \
    do {                                                                                       \
        cudaError_t _e = (expr);                                                               \
        if (_e != cudaSuccess) {                                                               \
            pysdr_set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
            return PYSDR_ERR_CUDA;                                                             \
        }                                                                                      \
    } while (0)

#define LAUNCH_CHECK()                                                                         \
    do {                                                                                       \
        cudaError_t _e = cudaGetLastError();                                                   \
        if (_e != cudaSuccess) {                                                               \
            pysdr_set_error("%s:%d: kernel launch -> %s", __FILE__, __LINE__, cudaGetErrorString(_e)); \
            return PYSDR_ERR_CUDA;                                                             \
        }                                                                                      \
    } while (0)

// ---- programmatic dependent launch (PDL) --------------------------------------------------------------------------------
// The three kernels of a step (K1 -> AF filter -> fused back) run back to back on one stream.  Launched with
// programmaticStreamSerializationAllowed, a kernel's grid may be set up while its predecessor drains; it must not touch
// anything the predecessor writes before pdl_wait() (griddepcontrol.wait), and the predecessor lets the launch proceed as
// soon as all of its CTAs have passed pdl_trigger() (griddepcontrol.launch_dependents) — the dependent still cannot take
// an SM slot from a predecessor CTA that has not started.  PYSDR_NO_PDL=1 in the environment turns the attribute off.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
bool pysdr_pdl_enabled(void);
template <typename... KArgs, typename... Args>
static inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pysdr_pdl_enabled() ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, args...);
}

static inline i64 ceil_div_i64(i64 a, i64 b) { return -((-a) / b) ; }   // b>0, a>=0 in our uses
static inline i64 n_out_total(i64 n_in, int up, int down) { return (up * n_in + down - 1) / down; }

// NCO: phase accumulator is an exact u64 fraction of a cycle; only the top 32 bits feed sin/cos.
// Returns (cos, sin) of 2*pi*phase.
__device__ __forceinline__ float2 nco_cs(u64 phase) {
    int top = (int)(phase >> 32);
    float v = (float)top * 4.656612873077393e-10f;      // 2^-31: half-cycles in [-1,1)
    float s, c;
    sincospif(v, &s, &c);
    return make_float2(c, s);
}

struct AgcState {
    double ring[PYSDR_AGC_NB];
    i64 k;
    double gain, maxbuf, err, ref, beta;
};

// ---- K1 arguments (by value) ------------------------------------------------------------------
struct K1Args {
    const float2 *x;          // chunk, n_in samples (absolute index n0..)
    const float2 *hist;       // `need` samples preceding x (absolute n0-need..n0-1); NULL = all zero (stream start / after seek)
    int real_input;           // != 0: Im x == 0 for every sample (caller's promise): K1 skips the Im-x half of the FMAs
    int zero_c_hist;          // != 0: also clear the carried part C[rx][0..hc) of every complex memory row (seek folded into K1)
    int need;                 // lp-1
    i64 n0, n_in, m0, n_out;
    int up, down, lp, lp_pad, n_rx;
    const float2 *g;          // folded taps [n_rx][up][lp_pad]
    u64 acc[PYSDR_MAX_RX];    // LO phase at absolute sample n0
    u64 inc[PYSDR_MAX_RX];
    float2 *c_out;            // C[rx*c_stride + hc + i]
    i64 c_stride;
    int hc;
    float2 *bb_out;           // optional rx.iq copy [rx*bb_stride + i]
    i64 bb_stride;
};

int k1_launch_generic(const K1Args &a, cudaStream_t st);
// returns 1 if the tap-stationary fast path supports this geometry
int k1_fast_supported(int up, int down, int lp, int n_rx);
int k1_launch_fast(const K1Args &a, cudaStream_t st);
// taps-per-phase padding the fast path wants (multiple of 32)
int k1_fast_lp_pad(int lp);
// # of kernel launches the fast path needs for n_rx receivers (receiver groups of 1/2/4)
int k1_fast_groups(int lp, int n_rx);

// ---- K1 tensor-core variant (k1_mma.cu): tcgen05 TF32 split GEMM, interior of large calls ------------------------------
struct K1MmaPlan;
int k1_mma_supported(int up, int down, int lp, int n_rx);
K1MmaPlan *k1_mma_plan_create(int up, int down, int lp, int n_rx);
void k1_mma_plan_destroy(K1MmaPlan *p);
// g_host: the folded taps [n_rx][up][lp_pad] exactly as uploaded for k1_fast
int k1_mma_upload_taps(K1MmaPlan *p, const float2 *g_host, int lp_pad, cudaStream_t st);
// Runs the call through the tensor-core kernel.  *used = 0 and nothing launched when the call is too small or its
// geometry/alignment does not fit; the caller then takes the tap-stationary path.
int k1_launch_mma(K1MmaPlan *p, const K1Args &a, i64 min_rows, cudaStream_t st, int *used, int *launches);

// ---- K1 many-channel tensor-core variant (k1_chan.cu): banks of >= 16 receivers, any geometry with up * (1 or 2) <= 8 classes ----
struct K1ChanPlan;
int k1_chan_supported(int up, int down, int lp, int n_rx);
K1ChanPlan *k1_chan_plan_create(int up, int down, int lp, int n_rx);
void k1_chan_plan_destroy(K1ChanPlan *p);
int k1_chan_upload_taps(K1ChanPlan *p, const float2 *g_host, int lp_pad, cudaStream_t st);
// *used = 0 and nothing launched when the call has fewer than min_rows interior super-periods or its alignment does not fit
int k1_launch_chan(K1ChanPlan *p, const K1Args &a, i64 min_rows, cudaStream_t st, int *used, int *launches);

// ---- K2 fast path (k2_fftconv.cu) ---------------------------------------------------------------------
struct FftConvArgs {
    const float2 *C;          // complex memory + new samples, rows of c_stride: C[rx][0..hc+n_out)
    i64 c_stride;
    const float2 *H;          // per receiver: FFT of the AF taps in position order, scaled 1/N
    int L;                    // AF FIR length
    i64 n_out, m0;
    float *out;               // pre-AGC audio rows of 2*a_stride floats
    i64 a_stride;
    int mode[PYSDR_MAX_RX];
    u64 bfo_inc[PYSDR_MAX_RX];
    int ua[PYSDR_MAX_RX];     // work units (filled by fftconv_launch): receiver ua[u], paired with ub[u] (or -1) when both
    int ub[PYSDR_MAX_RX];     // have a real detector output and real taps: two real convolutions ride one complex FFT
};
int fftconv_supported(int L);
int fftconv_n_for(int L);
int fftconv_prepare_taps(const float2 *d_taps, int L, float2 *d_H, cudaStream_t st);
int fftconv_launch(const FftConvArgs &a, int n_rx, cudaStream_t st);
