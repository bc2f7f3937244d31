// K1 for MANY channels (BASELINE config 5, any set of offsets): the NCO mix + polyphase decimation of a whole bank of channel
// receivers as a dense split-TF32 contraction on tcgen05, sm_100a.  This is the case north_star reserves the tensor cores for:
// with 1024 channels the filter bank costs 13 kFLOP per input sample (474 flop/byte) and the FP32 kernels run it at 23 TFLOP/s.
//
// Same arithmetic contract as k1_generic.cu / k1_fast.cu / k1_mma.cu (reference: dsp.Receiver.demod_data's lo/dec stages,
// receiver.py:235,822,866; params.py:405 UP/DOWN; Tables.py:41-42 taps):  y_c[m] = e^{-j theta_c(n_m)} sum_j G_c[p_m][j] x[n_m-j].
//
// GEMM view (cfg5: 10 MS/s -> 48 kHz, 3/625, 334 taps per phase):
//     rows    = output instants of ONE class = (output phase i, parity s of the super-period when DOWN is odd): within a class the
//               rows are S*DOWN input samples apart (S = 1 or 2, so that the pitch is a multiple of 16 bytes), all rows use the same
//               taps and have the same 16-byte alignment                                                            -> M = 128 per tile
//     K       = the row's window of lp (+1 when the row has to start one sample early to be 16-byte aligned) complex samples,
//               re/im interleaved as they lie in the capture                                                        -> K = 672 at cfg5
//     columns = channels x (Re y, Im y), up to 96 channels per tile                                                 -> N <= 192
// A (the samples) is never materialised: one 2-D TMA map per class over the RAW capture (box {32 floats, 128 rows},
// SWIZZLE_128B) feeds a 4-stage shared-memory ring; converter warps split every fp32 sample into hi = x & 0xFFFFE000 (what the
// tensor core keeps of an fp32 operand) and lo = x - hi and tcgen05.st both into TENSOR MEMORY as the A operand (k1_mma.cu's
// scheme).  B (the channels' folded taps, hi = round-to-nearest TF32 and lo = remainder, host-built per class image and channel
// group, canonical K-major core matrices) does not fit shared memory for a whole bank, so it STREAMS from L2 through a 3-stage
// ring of 1-D bulk copies (4 K-steps = 48 KB per stage at N = 192).  Per K = 8 step three MMAs accumulate into the same fp32
// accumulator, smallest terms first: x_lo*g_hi, x_hi*g_lo, x_hi*g_hi (the dropped x_lo*g_lo term is 2^-23 of the product); one
// issuing thread, so the summation order is fixed and results are bit-reproducible.  D is double-buffered in tensor memory
// (2 x 192 columns) so the epilogue (tcgen05.ld, NCO de-rotation from the exact u64 phase, stores into the bank's complex
// memory) overlaps the next tile's MMAs.  Tiles are ordered class-fastest so that the CTAs running side by side write the
// interleaved outputs of the same super-periods (the 8-byte stores merge in L2).
//
// Measured: 0.333 ms per bank of 128 channels on a 4 s block of 10 MS/s, 2.67 ms for 1024 channels = 197 TFLOP/s
// of FP32-equivalent filtering (the FP32 kernels: 23); the kernel sits at the practical L2 -> SM bandwidth of its tile shape
// (9.3 TB/s of taps + samples), the MMAs complete at 127 clocks where a bare stream of them needs 74 (tools/microbench/umma_rate.cu).
//
// Warp roles (512 threads, 1 CTA/SM, persistent):  warp 0 TMA producer (samples) | warp 1 MMA issuer (+ TMEM allocation) |
// warp 2 bulk-copy producer (taps) | warp 3 stream edges (outputs whose window reaches before x[0] or past its end: plain FP32
// dot products from global memory) | warps 4-11 converters (two groups, alternate chunks) | warps 12-15 epilogue.
#include "umma.cuh"
#include <algorithm>

#define KC_THREADS 512
#define KC_NSA 4                   /* sample ring: stages of one 16 KB box = 32 floats (4 K-steps) of 128 rows */
#define KC_BOX_BYTES 16384
#define KC_NSB 3                   /* tap ring: stages of 4 K-steps x (hi slab + lo slab) */
#define KC_ROWS 128
#define KC_MAX_N 192               /* accumulator columns per buffer: 96 channels x (re, im) */
#define KC_MAX_CLS 8
#define KC_COL_A 0                 /* tensor memory: 2 A stages x (32 hi + 32 lo columns) */
#define KC_COL_D 128               /*                2 accumulator buffers x KC_MAX_N columns */
#define KC_MIN_RX 16

struct KcMaps { CUtensorMap m[KC_MAX_CLS]; };

struct KcGeom {
    int n_chunks;                  // K-steps / 4
    int ncls, ngroups, nch, N;     // nch channels per group, N = 2 nch accumulator columns (multiple of 16)
    int up, down, S;
    int cls_i[KC_MAX_CLS], cls_s[KC_MAX_CLS], cls_o[KC_MAX_CLS], cls_img[KC_MAX_CLS];
    i64 cls_rows[KC_MAX_CLS];
    i64 n_tiles;
    i64 q_a;                       // absolute super-period of row 0 of the parity-0 classes
    i64 out_lo, out_hi;            // outputs [out_lo, out_hi) of the call come from the tensor cores, the rest from warp 3
    const unsigned char *img;      // [image = phase*2 + shift][group][K-step][hi slab | lo slab]
    unsigned slab;                 // N * 32 bytes: one K = 8 step of N columns
    unsigned idesc;
    unsigned long long bdesc0;     // LBO | SBO | version; the start address is added per step
};

#define KC_OFF_B (KC_NSA * KC_BOX_BYTES)

__global__ void __launch_bounds__(KC_THREADS, 1) k1_chan_kernel(const __grid_constant__ KcMaps maps, const K1Args a, const KcGeom g) {
    extern __shared__ __align__(1024) unsigned char kc_raw[];
    unsigned char *sm = (unsigned char *)(((uintptr_t)kc_raw + 1023) & ~(uintptr_t)1023);
    const unsigned stage_b = 8u * g.slab;                                      // 4 K-steps x (hi + lo)
    unsigned long long *bars = (unsigned long long *)(sm + KC_OFF_B + KC_NSB * stage_b);
    unsigned *tmem_slot = (unsigned *)(bars + 24);
    // x_full/x_empty: sample ring (TMA -> converters; a stage is always read by the same converter group since KC_NSA is even);
    // a_full/a_empty: the two tensor-memory A stages (converter group -> issuer); b_full/b_empty: tap ring (bulk copies ->
    // issuer); d_full/d_empty: the two accumulator buffers (issuer -> epilogue)
    const unsigned x_full = km_smem(bars + 0), x_empty = km_smem(bars + 4), a_full = km_smem(bars + 8), a_empty = km_smem(bars + 10);
    const unsigned b_full = km_smem(bars + 12), b_empty = km_smem(bars + 15), d_full = km_smem(bars + 18), d_empty = km_smem(bars + 20);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) {
        for (int s = 0; s < KC_NSA; ++s) { km_mbar_init(x_full + 8 * s, 1); km_mbar_init(x_empty + 8 * s, 4); }
        for (int s = 0; s < KC_NSB; ++s) { km_mbar_init(b_full + 8 * s, 1); km_mbar_init(b_empty + 8 * s, 1); }
        for (int s = 0; s < 2; ++s) {
            km_mbar_init(a_full + 8 * s, 4); km_mbar_init(a_empty + 8 * s, 1);
            km_mbar_init(d_full + 8 * s, 1); km_mbar_init(d_empty + 8 * s, 4);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(km_smem(tmem_slot)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    pdl_trigger();
    pdl_wait();                                  // x may come from a conversion kernel; C is read by the previous call's kernels
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const unsigned tmem = *tmem_slot;

    const i64 T0 = blockIdx.x, Tstep = gridDim.x;
    const i64 my_tiles = T0 < g.n_tiles ? (g.n_tiles - 1 - T0) / Tstep + 1 : 0;
    const int NCH = g.n_chunks;
    const int per_rt = g.ncls * g.ngroups;       // tile T = (row tile * ngroups + group) * ncls + class

    if (warp == 0) {
        // ===== TMA producer: the class's sample rows =====
        if (lane == 0) {
            i64 it = 0;
            for (i64 t = 0; t < my_tiles; ++t) {
                const i64 T = T0 + t * Tstep;
                const int cls = (int)(T % g.ncls);
                const int row0 = (int)((T / per_rt) * KC_ROWS);
                const CUtensorMap *tm = &maps.m[cls];
                for (int c = 0; c < NCH; ++c, ++it) {
                    const int s = (int)(it & (KC_NSA - 1));
                    km_mbar_wait(x_empty + 8 * s, (unsigned)(((it / KC_NSA) & 1) ^ 1));      // first pass over the ring: free
                    km_mbar_expect_tx(x_full + 8 * s, KC_BOX_BYTES);
                    km_tma_box(km_smem(sm + s * KC_BOX_BYTES), tm, c * 32, row0, x_full + 8 * s);
                }
            }
        }
    } else if (warp == 2) {
        // ===== tap producer: the (class image, channel group)'s K-steps, one stage = 4 K-steps =====
        if (lane == 0) {
            int sb = 0;
            unsigned ph = 1;                                                 // first pass over the ring: the stages are free
            for (i64 t = 0; t < my_tiles; ++t) {
                const i64 T = T0 + t * Tstep;
                const int cls = (int)(T % g.ncls), grp = (int)((T / g.ncls) % g.ngroups);
                const unsigned char *src = g.img + ((size_t)g.cls_img[cls] * g.ngroups + grp) * ((size_t)NCH * stage_b);
                for (int c = 0; c < NCH; ++c) {
                    km_mbar_wait(b_empty + 8 * sb, ph);
                    km_mbar_expect_tx(b_full + 8 * sb, stage_b);
                    const unsigned dst = km_smem(sm + KC_OFF_B + sb * stage_b);
#pragma unroll
                    for (int kk = 0; kk < 4; ++kk)
                        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst + kk * 2 * g.slab),
                                     "l"(src + (size_t)c * stage_b + (size_t)kk * 2 * g.slab), "r"(2 * g.slab), "r"(b_full + 8 * sb)
                                     : "memory");
                    if (++sb == KC_NSB) { sb = 0; ph ^= 1u; }
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer: per K = 8 step x_lo*g_hi, x_hi*g_lo, x_hi*g_hi into the same accumulator.  The WHOLE warp runs the
        // loop (barrier waits included) and one elected lane issues: with warp-uniform control flow the descriptors and tensor-
        // memory addresses live in uniform registers; inside an `if (lane == 0)` region every MMA cost two register-file ->
        // uniform-register moves and the issue loop (137 clocks per MMA) was the kernel's bottleneck. =====
        const unsigned b_ring = km_smem(sm + KC_OFF_B);
        const unsigned long long dstep = (unsigned long long)((2u * g.slab) >> 4), dlo = (unsigned long long)(g.slab >> 4);
        const unsigned idesc = g.idesc;
        i64 it = 0;
        int sb = 0;
        unsigned bph = 0;
        for (i64 t = 0; t < my_tiles; ++t) {
            const int buf = (int)(t & 1);
            km_mbar_wait(d_empty + 8 * buf, (unsigned)(((t >> 1) & 1) ^ 1));                // first use of a buffer: free
            const unsigned dcol = tmem + KC_COL_D + buf * KC_MAX_N;
            for (int c = 0; c < NCH; ++c, ++it) {
                const int sa = (int)(it & 1);
                km_mbar_wait(a_full + 8 * sa, (unsigned)((it >> 1) & 1));
                km_mbar_wait(b_full + 8 * sb, bph);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const unsigned a_hi = tmem + KC_COL_A + sa * 64, a_lo = a_hi + 32;
                const unsigned long long bd = g.bdesc0 | (unsigned long long)((b_ring + sb * stage_b) >> 4);   // < 2^14: no carry
                if (km_elect()) {
#pragma unroll
                    for (int kk = 0; kk < 4; ++kk) {
                        const unsigned long long bhi = bd + kk * dstep, blo = bhi + dlo;
                        km_mma_acc(dcol, a_lo + kk * 8, bhi, idesc, (c | kk) != 0 ? 1u : 0u);
                        km_mma_acc(dcol, a_hi + kk * 8, blo, idesc, 1u);
                        km_mma_acc(dcol, a_hi + kk * 8, bhi, idesc, 1u);
                    }
                    km_commit(a_empty + 8 * sa);
                    km_commit(b_empty + 8 * sb);
                }
                __syncwarp();
                if (++sb == KC_NSB) { sb = 0; bph ^= 1u; }
            }
            if (km_elect()) km_commit(d_full + 8 * buf);
            __syncwarp();
        }
    } else if (warp == 3) {
        // ===== stream edges (and the pending seek's clearing of the carried complex memory) =====
        if (a.zero_c_hist) {
            for (int r = blockIdx.x; r < a.n_rx; r += gridDim.x) {
                float2 *crow = a.c_out + (size_t)r * a.c_stride;
                for (int e = lane; e < a.hc; e += 32) crow[e] = make_float2(0.f, 0.f);
            }
        }
        const size_t rx_pitch = (size_t)a.up * a.lp_pad;
        const i64 n_edge = g.out_lo + (a.n_out - g.out_hi);
        const int subs = (a.n_rx + 3) >> 2;
        for (i64 e = blockIdx.x; e < n_edge * subs; e += gridDim.x) {
            const i64 ei = e / subs;
            const int rx0 = (int)(e - ei * subs) * 4;
            const i64 i = ei < g.out_lo ? ei : g.out_hi + (ei - g.out_lo);
            const i64 tt = (a.m0 + i) * a.down;
            const i64 nm = tt / a.up;
            const int ph = (int)(tt - nm * a.up);
            const i64 r0 = nm - a.n0;                                           // newest input of this output, relative to x[0]
            float sr[4], si[4];
#pragma unroll
            for (int r = 0; r < 4; ++r) { sr[r] = 0.f; si[r] = 0.f; }
            const float2 *gp = a.g + (size_t)ph * a.lp_pad;
            for (int j = lane; j < a.lp; j += 32) {
                const i64 idx = r0 - j;
                float2 xv = make_float2(0.f, 0.f);
                if (idx >= 0) { if (idx < a.n_in) xv = __ldg(a.x + idx); }
                else if (idx >= -(i64)a.need && a.hist) xv = a.hist[a.need + idx];
#pragma unroll
                for (int r = 0; r < 4; ++r) {
                    if (rx0 + r < a.n_rx) {
                        const float2 gv = __ldg(gp + (size_t)(rx0 + r) * rx_pitch + j);
                        sr[r] = fmaf(gv.x, xv.x, sr[r]);
                        sr[r] = fmaf(-gv.y, xv.y, sr[r]);
                        si[r] = fmaf(gv.x, xv.y, si[r]);
                        si[r] = fmaf(gv.y, xv.x, si[r]);
                    }
                }
            }
#pragma unroll
            for (int r = 0; r < 4; ++r) {
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    sr[r] += __shfl_xor_sync(0xffffffffu, sr[r], o);
                    si[r] += __shfl_xor_sync(0xffffffffu, si[r], o);
                }
            }
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                const int rr = rx0 + r;
                if (rr < a.n_rx && lane == r) {
                    const u64 phs = a.acc[rr] + a.inc[rr] * (u64)r0;
                    const float ang = (float)(int)(phs >> 32) * 1.4629180792671596e-09f;
                    float sn, cs;
                    __sincosf(ang, &sn, &cs);
                    float2 y;
                    y.x = fmaf(sr[r], cs, si[r] * sn);
                    y.y = fmaf(si[r], cs, -sr[r] * sn);
                    a.c_out[(size_t)rr * a.c_stride + a.hc + i] = y;
                    if (a.bb_out) a.bb_out[(size_t)rr * a.bb_stride + i] = y;
                }
            }
        }
    } else if (warp >= 4 && warp < 12) {
        // ===== converters: shared memory (swizzled rows) -> hi / lo -> tensor memory.  Group 0 (warps 4-7) takes the even
        // chunks into A stage 0, group 1 (warps 8-11) the odd chunks into A stage 1. =====
        const int grp = (warp - 4) >> 2, w4 = warp & 3;
        const int row = w4 * 32 + lane;
        const unsigned lane_addr = (unsigned)(w4 * 32) << 16;
        const i64 n_it = my_tiles * NCH;
        const unsigned at = tmem + lane_addr + KC_COL_A + grp * 64;
        for (i64 it = grp; it < n_it; it += 2) {
            const int s = (int)(it & (KC_NSA - 1));
            km_mbar_wait(x_full + 8 * s, (unsigned)((it / KC_NSA) & 1));
            const unsigned rowp = km_smem(sm + s * KC_BOX_BYTES + row * 128);
            unsigned hi[32], lo[32];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const uint4 v = km_lds128(rowp + ((j ^ (row & 7)) << 4));
                const unsigned u[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const unsigned hh = u[e] & 0xFFFFE000u;
                    hi[4 * j + e] = hh;
                    lo[4 * j + e] = __float_as_uint(__uint_as_float(u[e]) - __uint_as_float(hh));
                }
            }
            km_mbar_wait(a_empty + 8 * grp, (unsigned)(((it >> 1) & 1) ^ 1));   // the MMAs that read this A stage last time are done
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            KM_ST16(at, hi); KM_ST16(at + 16, hi + 16);
            KM_ST16(at + 32, lo); KM_ST16(at + 48, lo + 16);
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            // the stage is released only now: the tcgen05.st consumed every register the loads fill, so the loads have been
            // performed.  Released right after ISSUING the loads (the hi/lo arithmetic gets scheduled below the arrive), the next
            // TMA box overwrote rows not yet read: one wrong row in ~30 000 under full load (measured, 46 of 192 512 outputs
            // per channel in one 4 s block; none since)
            if (lane == 0) {
                km_mbar_arrive(x_empty + 8 * s);
                km_mbar_arrive(a_full + 8 * grp);
            }
        }
    } else if (warp >= 12) {
        // ===== epilogue: accumulator -> NCO de-rotation -> the channels' complex memory =====
        const int w4 = warp & 3;
        const int row = w4 * 32 + lane;
        const unsigned lane_addr = (unsigned)(w4 * 32) << 16;
        const int n_cb = g.N >> 4;
        for (i64 t = 0; t < my_tiles; ++t) {
            const int buf = (int)(t & 1);
            const i64 T = T0 + t * Tstep;
            const int cls = (int)(T % g.ncls), grp = (int)((T / g.ncls) % g.ngroups);
            const i64 u = (T / per_rt) * KC_ROWS + row;                          // row of the class
            const i64 q = g.q_a + g.cls_s[cls] + (i64)g.S * u;                   // absolute super-period
            const i64 idx = q * g.up + g.cls_i[cls] - a.m0;                      // output index within the call
            const u64 n_rel = (u64)(q * g.down + g.cls_o[cls] - a.n0);           // newest input sample, relative to x[0]
            const bool row_ok = u < g.cls_rows[cls] && idx >= 0 && idx < a.n_out;
            km_mbar_wait(d_full + 8 * buf, (unsigned)((t >> 1) & 1));
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const unsigned dA = tmem + lane_addr + KC_COL_D + buf * KC_MAX_N;
            for (int cb = 0; cb < n_cb; ++cb) {
                unsigned v[16];
                KM_LD16(dA + cb * 16, v);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                if (cb == n_cb - 1) {                                            // the buffer is in registers: the issuer may reuse it
                    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                    __syncwarp();
                    if (lane == 0) km_mbar_arrive(d_empty + 8 * buf);
                }
                if (row_ok) {
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const int cl = cb * 8 + j, rx = grp * g.nch + cl;
                    if (cl >= g.nch || rx >= a.n_rx) continue;
                    const u64 ph = a.acc[rx] + a.inc[rx] * n_rel;
                    const float ang = (float)(int)(ph >> 32) * 1.4629180792671596e-09f;          // 2*pi*2^-32
                    float sn, cs;
                    __sincosf(ang, &sn, &cs);
                    const float yr = __uint_as_float(v[2 * j]), yi = __uint_as_float(v[2 * j + 1]);
                    float2 out;
                    out.x = fmaf(yr, cs, yi * sn);                               // (re + j im)(cos - j sin)
                    out.y = fmaf(yi, cs, -yr * sn);
                    a.c_out[(size_t)rx * a.c_stride + a.hc + idx] = out;
                    if (a.bb_out) a.bb_out[(size_t)rx * a.bb_stride + idx] = out;
                }
                }
                __syncwarp();                                                    // the tcgen05.ld of the next pass is warp-wide
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem));
}

// ---- host side ----------------------------------------------------------------------------------------------------------
struct K1ChanPlan {
    int up, down, lp, n_rx;
    int S, ncls;                               // super-period parities per class set (2 when DOWN is odd), classes = up * S
    int n_steps;                               // K = 8 steps per row, a multiple of 4; covers lp + 1 samples
    int ngroups, nch, N;
    unsigned slab;
    unsigned char *d_img;                      // [up * 2 images][ngroups][n_steps][2 * slab]
    bool attr_done[64];
};

static int kc_steps(int lp) { return ((2 * (lp + 1) + 7) / 8 + 3) / 4 * 4; }

int k1_chan_supported(int up, int down, int lp, int n_rx) {
    if (up < 1 || down < 1 || lp < 2 || n_rx < KC_MIN_RX || n_rx > PYSDR_MAX_RX) return 0;
    const int S = (down & 1) ? 2 : 1;
    if (up * S > KC_MAX_CLS) return 0;
    if (kc_steps(lp) * 8 > 4096) return 0;     // row window (floats): keeps the tap images and the tensor map's inner extent modest
    if (!km_encode_fn()) return 0;
    return 1;
}

static void kc_plan_init(K1ChanPlan *p, int up, int down, int lp, int n_rx) {
    p->up = up; p->down = down; p->lp = lp; p->n_rx = n_rx;
    p->S = (down & 1) ? 2 : 1;
    p->ncls = up * p->S;
    p->n_steps = kc_steps(lp);
    p->ngroups = (n_rx + KC_MAX_N / 2 - 1) / (KC_MAX_N / 2);
    p->nch = ((n_rx + p->ngroups - 1) / p->ngroups + 7) / 8 * 8;               // equal groups, N a multiple of 16
    p->N = 2 * p->nch;
    p->slab = (unsigned)p->N * 32u;
    p->d_img = nullptr;
    for (int i = 0; i < 64; ++i) p->attr_done[i] = false;
}

K1ChanPlan *k1_chan_plan_create(int up, int down, int lp, int n_rx) {
    if (!k1_chan_supported(up, down, lp, n_rx)) return nullptr;
    K1ChanPlan *p = new K1ChanPlan();
    kc_plan_init(p, up, down, lp, n_rx);
    return p;
}

void k1_chan_plan_destroy(K1ChanPlan *p) {
    if (!p) return;
    cudaFree(p->d_img);
    delete p;
}

// g_host: folded taps [n_rx][up][lp_pad] (what k1_fast reads).  Image (phase i, shift sh): B[k][col], k = 2 t (+1 for Im x) over
// the row's samples t = 0 .. 4 n_steps - 1 (the row starts sh samples before the window), col = 2 c (+1 for the Im output):
//   Re y += g_re x_re - g_im x_im,  Im y += g_im x_re + g_re x_im,  tap index j = (lp - 1) + sh - t.
static void kc_build_image(const K1ChanPlan *p, const float2 *g_host, int lp_pad, std::vector<float> &img) {
    const int N = p->N, ns = p->n_steps;
    const size_t step_floats = (size_t)2 * N * 8;                               // hi slab + lo slab
    const size_t grp_floats = (size_t)ns * step_floats;
    img.assign((size_t)p->up * 2 * p->ngroups * grp_floats, 0.f);
    for (int i = 0; i < p->up; ++i) {
        const int tap_phase = (int)(((i64)i * p->down) % p->up);
        for (int sh = 0; sh < 2; ++sh)
            for (int grp = 0; grp < p->ngroups; ++grp) {
                float *base = img.data() + ((size_t)(i * 2 + sh) * p->ngroups + grp) * grp_floats;
                for (int cl = 0; cl < p->nch; ++cl) {
                    const int rx = grp * p->nch + cl;
                    if (rx >= p->n_rx) break;
                    const float2 *gr = g_host + ((size_t)rx * p->up + tap_phase) * lp_pad;
                    for (int t = 0; t < 4 * ns; ++t) {
                        const int j = (p->lp - 1) + sh - t;
                        if (j < 0 || j >= p->lp) continue;
                        const float2 gg = gr[j];
                        for (int im_x = 0; im_x < 2; ++im_x) {
                            const int k = 2 * t + im_x, s = k >> 3, kk = k & 7;
                            const float vals[2] = {im_x ? -gg.y : gg.x, im_x ? gg.x : gg.y};     // columns Re y, Im y
                            float *hi_slab = base + (size_t)s * step_floats, *lo_slab = hi_slab + (size_t)N * 8;
                            for (int c = 0; c < 2; ++c) {
                                const int n = 2 * cl + c;
                                const float h = km_tf32_hi(vals[c]);
                                // canonical K-major no-swizzle layout: [chunk = kk/4][n/8][n%8][kk%4]
                                const size_t e = (size_t)(kk >> 2) * (N * 4) + (size_t)(n >> 3) * 32 + (n & 7) * 4 + (kk & 3);
                                hi_slab[e] = h;
                                lo_slab[e] = vals[c] - h;
                            }
                        }
                    }
                }
            }
    }
}

int k1_chan_upload_taps(K1ChanPlan *p, const float2 *g_host, int lp_pad, cudaStream_t st) {
    std::vector<float> img;
    kc_build_image(p, g_host, lp_pad, img);
    if (p->d_img) { CUDA_TRY(cudaFree(p->d_img)); p->d_img = nullptr; }
    CUDA_TRY(cudaMalloc(&p->d_img, img.size() * sizeof(float)));
    CUDA_TRY(cudaMemcpyAsync(p->d_img, img.data(), img.size() * sizeof(float), cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    return PYSDR_OK;
}

// Which super-periods of a call the tensor cores take and how each class's rows lie in the capture (pure host arithmetic).
// r0[cls]: first sample of row 0 (relative to x[0], the alignment shift already applied).  Returns 0 when the call does not fit.
static int kc_geometry(const K1ChanPlan *p, i64 n0, i64 n_in, i64 m0, i64 n_out, uintptr_t x_addr, i64 min_rows, KcGeom *g, i64 *r0s) {
    const int S = p->S, ns = p->n_steps, up = p->up, down = p->down, lp = p->lp;
    const i64 span = 4 * (i64)ns;                                              // samples a row's boxes read
    // first tensor-core super-period: its rows (lp-1 samples of look-back, +1 for the alignment shift) start inside x
    i64 q_a = (n0 + lp + down - 1) / down;
    const i64 q_first = m0 / up;
    if (q_a < q_first) q_a = q_first;
    // last one: every class's row (offset < down, read `span` samples from at most one sample early) ends inside x
    const i64 num = n_in + n0 - span + lp - down;
    if (num < 0) return 0;
    const i64 q_end = num / down + 1;
    const i64 rows_total = q_end - q_a;
    if (rows_total < min_rows || rows_total < 2 * S) return 0;
    memset(g, 0, sizeof(*g));
    g->n_chunks = ns / 4; g->ncls = p->ncls; g->ngroups = p->ngroups; g->nch = p->nch; g->N = p->N;
    g->up = up; g->down = down; g->S = S;
    i64 max_rows = 0;
    for (int i = 0; i < up; ++i)
        for (int s = 0; s < S; ++s) {
            const int cls = i * S + s;
            const int o = (int)(((i64)i * down) / up);
            const i64 rows = (rows_total - s + S - 1) / S;
            const i64 r0 = (q_a + s) * down + o - (lp - 1) - n0;                // first sample of row 0's window, relative to x[0]
            const int sh = (int)(((x_addr + (uintptr_t)r0 * 8) >> 3) & 1);      // start one sample early when that is the aligned one
            if (((x_addr + (uintptr_t)(r0 - sh) * 8) & 15) != 0 || r0 - sh < 0) return 0;
            g->cls_i[cls] = i; g->cls_s[cls] = s; g->cls_o[cls] = o; g->cls_img[cls] = i * 2 + sh; g->cls_rows[cls] = rows;
            r0s[cls] = r0 - sh;
            max_rows = rows > max_rows ? rows : max_rows;
        }
    g->q_a = q_a;
    g->n_tiles = ((max_rows + KC_ROWS - 1) / KC_ROWS) * p->ngroups * p->ncls;
    const i64 lo = q_a * up - m0, hi = q_end * up - m0;
    g->out_lo = lo < 0 ? 0 : (lo > n_out ? n_out : lo);
    g->out_hi = hi < g->out_lo ? g->out_lo : (hi > n_out ? n_out : hi);
    g->slab = p->slab;
    const int N = p->N;
    // shared-memory descriptor (K-major, no swizzle): LBO = bytes between the two 16-byte K chunks of a step | SBO = bytes between
    // 8-column groups | version 1;  instruction descriptor: D f32, A/B tf32, K-major, N, M = 128
    g->bdesc0 = ((unsigned long long)((N * 16) >> 4) << 16) | ((unsigned long long)(128 >> 4) << 32) | (1ull << 46);
    g->idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((unsigned)(N >> 3) << 17) | ((128u >> 4) << 24);
    return 1;
}

int k1_launch_chan(K1ChanPlan *p, const K1Args &a, i64 min_rows, cudaStream_t st, int *used, int *launches) {
    *used = 0;
    if (!p || !p->d_img || a.real_input || a.n_out <= 0 || a.up != p->up || a.down != p->down || a.lp != p->lp || a.n_rx != p->n_rx) return PYSDR_OK;
    KmEncodeFn enc = km_encode_fn();
    if (!enc) return PYSDR_OK;
    KcMaps maps;
    KcGeom g;
    i64 r0s[KC_MAX_CLS];
    memset(&maps, 0, sizeof(maps));
    if (!kc_geometry(p, a.n0, a.n_in, a.m0, a.n_out, (uintptr_t)a.x, min_rows, &g, r0s)) return PYSDR_OK;
    for (int cls = 0; cls < p->ncls; ++cls) {
        cuuint64_t dims[2] = {(cuuint64_t)p->n_steps * 8, (cuuint64_t)g.cls_rows[cls]};
        cuuint64_t strides[1] = {(cuuint64_t)p->S * a.down * 8};
        cuuint32_t box[2] = {32, KC_ROWS};
        cuuint32_t estr[2] = {1, 1};
        const CUresult cr = enc(&maps.m[cls], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void *)(a.x + r0s[cls]), dims, strides, box, estr,
                                CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (cr != CUDA_SUCCESS) { pysdr_set_error("k1_chan: cuTensorMapEncodeTiled failed (%d)", (int)cr); return PYSDR_ERR_CUDA; }
    }
    g.img = p->d_img;
    const size_t smem = 1008 + KC_OFF_B + (size_t)KC_NSB * 8 * p->slab + 26 * 8;
    if (smem > 227 * 1024) return PYSDR_OK;
    const int dev = pysdr_device() & 63;
    if (!p->attr_done[dev]) {
        CUDA_TRY(cudaFuncSetAttribute(k1_chan_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        p->attr_done[dev] = true;
    }
    const int sms = pysdr_sm_count();
    const int grid = (int)(g.n_tiles < sms ? g.n_tiles : sms);
    CUDA_TRY(launch_pdl(k1_chan_kernel, dim3(grid), dim3(KC_THREADS), smem, st, maps, a, g));
    if (launches) ++*launches;
    *used = 1;
    return PYSDR_OK;
}

// Test hook (host arithmetic only, no device): the plan, the per-call geometry and the tap images exactly as the kernel gets
// them, so that the CPU suite can emulate the contraction in numpy (tests/test_host_logic.py).  out[0..15]: used, q_a, out_lo,
// out_hi, n_steps, ngroups, nch, N, ncls, S, n_tiles; out[16 + 8 cls ..]: i, s, o, image, rows, r0 (shift applied).
// Returns the number of floats of the image (copied into img when img_cap is large enough), or a negative error code.
extern "C" int64_t pysdr_k1chan_debug_plan(int up, int down, int lp, int n_rx, const float *g_host, int lp_pad, int64_t n0, int64_t n_in,
                                           int64_t m0, int64_t n_out, uint64_t x_addr, int64_t min_rows, int64_t *out, float *img,
                                           int64_t img_cap) {
    if (!out || !g_host || up < 1 || down < 1 || lp < 2 || n_rx < 1 || n_rx > PYSDR_MAX_RX || up * ((down & 1) ? 2 : 1) > KC_MAX_CLS) {
        pysdr_set_error("k1chan_debug_plan: bad arguments");
        return PYSDR_ERR_ARG;
    }
    K1ChanPlan p;
    kc_plan_init(&p, up, down, lp, n_rx);
    KcGeom g;
    i64 r0s[KC_MAX_CLS];
    memset(out, 0, sizeof(int64_t) * (16 + 8 * KC_MAX_CLS));
    const int used = kc_geometry(&p, n0, n_in, m0, n_out, (uintptr_t)x_addr, min_rows, &g, r0s);
    out[0] = used; out[4] = p.n_steps; out[5] = p.ngroups; out[6] = p.nch; out[7] = p.N; out[8] = p.ncls; out[9] = p.S;
    if (used) {
        out[1] = g.q_a; out[2] = g.out_lo; out[3] = g.out_hi; out[10] = g.n_tiles;
        for (int c = 0; c < p.ncls; ++c) {
            int64_t *o = out + 16 + 8 * c;
            o[0] = g.cls_i[c]; o[1] = g.cls_s[c]; o[2] = g.cls_o[c]; o[3] = g.cls_img[c]; o[4] = g.cls_rows[c]; o[5] = r0s[c];
        }
    }
    std::vector<float> v;
    kc_build_image(&p, (const float2 *)g_host, lp_pad, v);
    if (img && (int64_t)v.size() <= img_cap) memcpy(img, v.data(), v.size() * sizeof(float));
    return (int64_t)v.size();
}
