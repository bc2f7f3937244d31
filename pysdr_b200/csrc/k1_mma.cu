// K1 (tensor-core variant): fused NCO-mix + polyphase FIR decimate as a 3xTF32-style split GEMM on tcgen05, sm_100a.
//
// Same arithmetic contract as k1_generic.cu / k1_fast.cu (reference: dsp.Receiver.demod_data's lo/dec stages,
// receiver.py:235,822,866; params.py:405 UP/DOWN; Tables.py:41-42 taps):  y_r[m] = e^{-j theta_r(n_m)} sum_j G_r[p_m][j] x[n_m-j].
//
// Why tensor cores here: at 4 receivers x 334 complex taps per output the FP32 pipes need 32 FMA per input sample, which
// puts k1_fast at the FMA roofline (0.84 ms per 480 M samples, 70 % of the HBM rate).  The same contraction as a GEMM:
//
//     rows    = super-periods (DOWN input samples -> UP outputs), 128 per tile                       -> M = 128
//     K       = the 2*DOWN floats (re, im interleaved) of the row's input window; the row starts lp-1 samples before the
//               super-period so that output 0's window starts at k = 0
//     columns = "pieces": the UP output phases' windows cut at the row boundary (cfg2: 3 windows -> 4 pieces, the last
//               window continues in the NEXT row), x n_rx receivers x (re, im) x (hi, lo) tap halves              -> 16 per piece
//
// A (the samples) never exists as a matrix in memory: a 2-D TMA box {32 floats, 128 rows} with row pitch DOWN*8 bytes reads
// the raw capture (SWIZZLE_128B, conflict-free per-row LDS.128), converter warps split every fp32 sample into
// hi = x & 0xFFFFE000 (exactly what the tensor core keeps of an fp32 operand: it truncates, measured with
// tools/microbench/umma_probe.cu) and lo = x - hi (exact), and tcgen05.st both halves into TENSOR MEMORY as the A operand
// (lane = row, column = k) — A from TMEM costs no shared-memory bandwidth, which is what makes this formulation faster than
// the HBM stream (A from shared memory would need ~80 B of shared-memory traffic per sample).  B (the folded taps, hi and lo
// halves side by side in the N dimension, K-major, no swizzle) sits in shared memory for the whole kernel (128 KB at cfg2).
// Per K = 8 step only the pieces whose window covers the step are multiplied (N = 32 of 64 columns at cfg2): the D columns
// are ordered by window start so that the live pieces are contiguous.  D = (hi+lo) x (hi+lo) keeps all four partial products
// in one fp32 accumulator pair; measured error vs float64 ~1e-6 of peak, the same as the FP32 kernel.
//
// Warp roles (512 threads, 1 CTA/SM, persistent over tiles of 127 output rows — row 127 only feeds row 126's last piece):
//   warp 0     TMA producer (3-stage ring of 2 x 16 KB boxes)             warps 1, 2  MMA issuers (one thread each: x_hi / x_lo products)
//   warps 4-11 converters (two groups, alternate chunks)             warps 12-15 epilogue: tcgen05.ld D, hi+lo, join the
//                                                                                piece of the next row, NCO de-rotation, store
//   warp 3     stream edges: the few outputs whose rows reach before x[0] (carried history) or past its end, one plain
//              FP32 dot product per output straight from global memory, spread over the CTAs and hidden under the pipeline
#include "umma.cuh"
#include <algorithm>

#define KM_THREADS 512
#define KM_NST 3                   /* shared-memory stages of one chunk = two 16 KB boxes = 64 floats (8 K-steps) of 128 rows */
#define KM_BOX_BYTES 16384
#define KM_STAGE_BYTES (2 * KM_BOX_BYTES)
#define KM_ROWS 128
#define KM_ADV 127
#define KM_MAX_STEPS 128
#define KM_MAXG 4                  /* pieces (column groups) */
#define KM_GW 16                   /* columns per piece: 4 receivers x (re, im) x (hi, lo) */
#define KM_DW (KM_MAXG * KM_GW)
#define KM_TAB_BYTES (KM_MAX_STEPS * 16)
#define KM_B_MAX (128 * 1024)

struct KmStep { unsigned long long bdesc; unsigned idesc; unsigned dcol; };   // 16 bytes per K = 8 step

struct KmGeom {
    int n_steps, n_chunks, ng, up, down;
    int o[8];                      // newest-sample offset of output i within the super-period
    int g0[8], g1[8];              // piece of output i in its own row / in the next row (-1: none)
    i64 n_tiles, rows_out;         // tiles of 127 output rows; output rows in this launch
    i64 q_a;                       // absolute super-period index of row 0
    i64 out_lo, out_hi;            // outputs [out_lo, out_hi) of the call come from the tensor cores, the rest from warp 3
    int rx0;
    unsigned b_bytes;              // bytes of the B image (multiple of 16)
    const unsigned char *img;      // global: [KmStep table, KM_TAB_BYTES][B image]
};


// shared memory map (dynamic, 1024-byte aligned base)
#define KM_OFF_STAGES 0
#define KM_OFF_TAB (KM_NST * KM_STAGE_BYTES)
#define KM_OFF_B (KM_OFF_TAB + KM_TAB_BYTES)

__global__ void __launch_bounds__(KM_THREADS, 1) k1_mma_kernel(const __grid_constant__ CUtensorMap tmap, const K1Args a, const KmGeom g) {
    extern __shared__ __align__(1024) unsigned char km_raw[];
    unsigned char *sm = (unsigned char *)(((uintptr_t)km_raw + 1023) & ~(uintptr_t)1023);
    unsigned char *tail = sm + KM_OFF_B + g.b_bytes;                         // 16-byte aligned
    float *xch = (float *)tail;                                              // [2][4 warps][KM_MAXG][8]
    unsigned long long *bars = (unsigned long long *)(tail + 2 * 4 * KM_MAXG * 8 * 4);
    unsigned *tmem_slot = (unsigned *)(bars + 18);
    // x_full/x_empty: the shared-memory ring (TMA -> converters); a_full/a_empty: the two tensor-memory A stages (converters ->
    // MMA issuers, one stage per converter group); d_full/d_empty: the two accumulator buffers (issuers -> epilogue).
    // x_full is per (converter group, stage): a waiter must see EVERY phase of its barrier (parity waits alias after two
    // completions), and a group only reads every other fill of a stage.
    const unsigned x_full = km_smem(bars + 0), x_empty = km_smem(bars + 6), a_full = km_smem(bars + 9), a_empty = km_smem(bars + 11);
    const unsigned d_full = km_smem(bars + 13), d_empty = km_smem(bars + 15), b_ready = km_smem(bars + 17);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) {
        for (int s = 0; s < 2 * KM_NST; ++s) km_mbar_init(x_full + 8 * s, 1);
        for (int s = 0; s < KM_NST; ++s) km_mbar_init(x_empty + 8 * s, 4);
        for (int s = 0; s < 2; ++s) {
            km_mbar_init(a_full + 8 * s, 4); km_mbar_init(a_empty + 8 * s, 2);
            km_mbar_init(d_full + 8 * s, 2); km_mbar_init(d_empty + 8 * s, 4);
        }
        km_mbar_init(b_ready, 1);
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
    // tensor memory: 2 A stages x (64 hi + 64 lo columns) | D: 2 buffers x 2 issuers (x_hi products, x_lo products) x KM_DW
    const unsigned colA = 0, colD = 256;

    const i64 T0 = blockIdx.x, Tstep = gridDim.x;
    const i64 my_tiles = T0 < g.n_tiles ? (g.n_tiles - 1 - T0) / Tstep + 1 : 0;
    const int NCH = g.n_chunks;

    if (warp == 0) {
        // ===== TMA producer =====
        if (lane == 0) {
            // taps + step table: one bulk copy per 32 KB piece, all on b_ready
            const unsigned total = KM_TAB_BYTES + g.b_bytes;
            km_mbar_expect_tx(b_ready, total);
            for (unsigned off = 0; off < total; off += 32768) {
                const unsigned n = total - off < 32768 ? total - off : 32768;
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(km_smem(sm + KM_OFF_TAB + off)),
                             "l"(g.img + off), "r"(n), "r"(b_ready)
                             : "memory");
            }
            int s = 0, grp = 0;
            unsigned ph = 1;                                                 // first pass over the ring: the stages are free
            for (i64 t = 0; t < my_tiles; ++t) {
                const int row0 = (int)((T0 + t * Tstep) * KM_ADV);
                for (int c = 0; c < NCH; ++c) {
                    km_mbar_wait(x_empty + 8 * s, ph);
                    const unsigned full = x_full + 8 * (grp * KM_NST + s);   // the chunk's reader: converter group (chunk & 1)
                    const unsigned dst = km_smem(sm + KM_OFF_STAGES + s * KM_STAGE_BYTES);
                    km_mbar_expect_tx(full, KM_STAGE_BYTES);
                    km_tma_box(dst, &tmap, c * 64, row0, full);
                    km_tma_box(dst + KM_BOX_BYTES, &tmap, c * 64 + 32, row0, full);
                    if (++s == KM_NST) { s = 0; ph ^= 1u; }
                    grp ^= 1;
                }
            }
        }
    } else if (warp == 1 || warp == 2) {
        // ===== MMA issuers: one thread of warp 1 multiplies the x_hi halves, one thread of warp 2 the x_lo halves, each into
        // its own accumulator.  A single thread issues one of these small MMAs per ~52 clocks (measured,
        // tools/microbench/umma_probe.cu) and pays ~350 clocks per barrier hand-over, hence two issuers and 8 MMAs per chunk. =====
        if (lane == 0) {
            const int half = warp - 1;
            km_mbar_wait(b_ready, 0);
            const unsigned tab = km_smem(sm + KM_OFF_TAB);
            const unsigned long long b_base = (unsigned long long)(km_smem(sm + KM_OFF_B) >> 4);   // descriptors hold image offsets
            i64 it = 0;
            uint4 e[8];
#pragma unroll
            for (int kk = 0; kk < 8; ++kk) e[kk] = km_lds128(tab + kk * 16);
            for (i64 t = 0; t < my_tiles; ++t) {
                const int buf = (int)(t & 1);
                km_mbar_wait(d_empty + 8 * buf, (unsigned)((t >> 1) & 1));      // completion #u: u = 0 is the initial zeroing
                const unsigned dbase = tmem + colD + (buf * 2 + half) * KM_DW;
                for (int c = 0; c < NCH; ++c, ++it) {
                    const int sa = (int)(it & 1);
                    km_mbar_wait(a_full + 8 * sa, (unsigned)((it >> 1) & 1));
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const unsigned abase = tmem + colA + sa * 128 + half * 64;
#pragma unroll
                    for (int kk = 0; kk < 8; ++kk) {
                        if (c * 8 + kk < g.n_steps) {
                            const unsigned long long bdesc = (((unsigned long long)e[kk].y << 32) | e[kk].x) + b_base;
                            km_mma(dbase + e[kk].w, abase + kk * 8, bdesc, e[kk].z);
                        }
                    }
                    km_commit(a_empty + 8 * sa);
                    const int cn = (c + 1 == NCH) ? 0 : c + 1;                  // next chunk's table entries while the MMAs run
#pragma unroll
                    for (int kk = 0; kk < 8; ++kk) e[kk] = km_lds128(tab + (cn * 8 + kk) * 16);
                }
                km_commit(d_full + 8 * buf);
            }
        }
    } else if (warp == 3) {
        // ===== stream edges (and the pending seek's clearing of the carried complex memory) =====
        if (a.zero_c_hist && g.rx0 == 0) {
            for (int r = blockIdx.x; r < a.n_rx; r += gridDim.x) {
                float2 *crow = a.c_out + (size_t)r * a.c_stride;
                for (int e = lane; e < a.hc; e += 32) crow[e] = make_float2(0.f, 0.f);
            }
        }
        const size_t rx_pitch = (size_t)a.up * a.lp_pad;
        const i64 n_edge = g.out_lo + (a.n_out - g.out_hi);
        for (i64 e = blockIdx.x; e < n_edge; e += gridDim.x) {
            const i64 i = e < g.out_lo ? e : g.out_hi + (e - g.out_lo);
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
                    if (g.rx0 + r < a.n_rx) {
                        const float2 gv = __ldg(gp + (size_t)(g.rx0 + r) * rx_pitch + j);
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
                const int rr = g.rx0 + r;
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
        const unsigned at = tmem + lane_addr + colA + grp * 128;
        int s = grp, sn = 0;                                                 // s = it % KM_NST for it = grp, grp + 2, ...; sn counts
        unsigned xph = 0;                                                    // the group's visits: each (group, stage) every 3rd
        for (i64 it = grp; it < n_it; it += 2) {
            km_mbar_wait(x_full + 8 * (grp * KM_NST + s), xph);
            const unsigned rowp = km_smem(sm + KM_OFF_STAGES + s * KM_STAGE_BYTES + row * 128);
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                unsigned hi[32], lo[32];
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const uint4 v = km_lds128(rowp + h * KM_BOX_BYTES + ((j ^ (row & 7)) << 4));
                    const unsigned u[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const unsigned hh = u[e] & 0xFFFFE000u;
                        hi[4 * j + e] = hh;
                        lo[4 * j + e] = __float_as_uint(__uint_as_float(u[e]) - __uint_as_float(hh));
                    }
                }
                if (h == 0) {
                    km_mbar_wait(a_empty + 8 * grp, (unsigned)(((it >> 1) & 1) ^ 1));   // the MMAs that read this A stage last time are done
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                }
                KM_ST16(at + h * 32, hi); KM_ST16(at + h * 32 + 16, hi + 16);
                KM_ST16(at + 64 + h * 32, lo); KM_ST16(at + 64 + h * 32 + 16, lo + 16);
                if (h == 1) {
                    // the stage is released only after the tcgen05.st of box 1 have ISSUED: they read every register the loads
                    // fill, so the loads have been performed.  (Arriving right after issuing the loads — the hi/lo arithmetic is
                    // scheduled below the arrive — lets the next TMA box overwrite rows not yet read: k1_chan.cu measured one
                    // wrong row in ~30 000 with that order under full load.)
                    __syncwarp();
                    if (lane == 0) km_mbar_arrive(x_empty + 8 * s);
                }
            }
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) km_mbar_arrive(a_full + 8 * grp);
            s += 2;
            if (s >= KM_NST) s -= KM_NST;
            if (++sn == KM_NST) { sn = 0; xph ^= 1u; }
        }
    } else if (warp >= 12) {
        // ===== epilogue =====
        const int w4 = warp & 3;
        const int row = w4 * 32 + lane;
        const unsigned lane_addr = (unsigned)(w4 * 32) << 16;
        unsigned zero[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) zero[i] = 0u;
        // all four accumulators start at zero (every MMA accumulates)
#pragma unroll
        for (int c = 0; c < 4 * KM_DW; c += 16) KM_ST16(tmem + lane_addr + colD + c, zero);
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncwarp();
        if (lane == 0) { km_mbar_arrive(d_empty); km_mbar_arrive(d_empty + 8); }

        u64 inc[4], acc[4];
        float2 *c_row[4], *bb_row[4];
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const int rx = g.rx0 + r;
            const bool ok = rx < a.n_rx;
            inc[r] = ok ? a.inc[rx] : 0ull;
            acc[r] = ok ? a.acc[rx] : 0ull;
            c_row[r] = ok ? a.c_out + (size_t)rx * a.c_stride + a.hc : nullptr;
            bb_row[r] = (ok && a.bb_out) ? a.bb_out + (size_t)rx * a.bb_stride : nullptr;
        }
        for (i64 t = 0; t < my_tiles; ++t) {
            const int buf = (int)(t & 1);
            const i64 T = T0 + t * Tstep;
            km_mbar_wait(d_full + 8 * buf, (unsigned)((t >> 1) & 1));
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            float sums[KM_MAXG][8];
            const unsigned dA = tmem + lane_addr + colD + buf * 2 * KM_DW;     // x_hi products; the x_lo products follow at + KM_DW
#pragma unroll
            for (int q = 0; q < KM_MAXG; ++q) {
                unsigned v[16], w[16];
                KM_LD16(dA + q * KM_GW, v);
                KM_LD16(dA + KM_DW + q * KM_GW, w);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                for (int j = 0; j < 8; ++j)                                      // (x_hi + x_lo) * (g_hi + g_lo); small terms first
                    sums[q][j] = (__uint_as_float(w[8 + j]) + __uint_as_float(w[j]) + __uint_as_float(v[8 + j])) + __uint_as_float(v[j]);
            }
#pragma unroll
            for (int c = 0; c < 2 * KM_DW; c += 16) KM_ST16(dA + c, zero);
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) km_mbar_arrive(d_empty + 8 * buf);                   // the MMA issuers may start the tile after next

            // pieces that continue a window of the PREVIOUS row: hand lane 0's values to the warp below
            float *xw = xch + (size_t)((t & 1) * 4 + w4) * (KM_MAXG * 8);
            if (lane == 0) {
#pragma unroll
                for (int q = 0; q < KM_MAXG; ++q)
#pragma unroll
                    for (int j = 0; j < 8; ++j) xw[q * 8 + j] = sums[q][j];
            }
            asm volatile("bar.sync 1, 128;" ::: "memory");
            const float *xn = xch + (size_t)((t & 1) * 4 + ((w4 + 1) & 3)) * (KM_MAXG * 8);

            const i64 q_row = g.q_a + T * KM_ADV + row;                        // absolute super-period of this row
            const bool row_ok = row < KM_ADV && T * KM_ADV + row < g.rows_out;
            const i64 ob = q_row * g.up - a.m0;                                  // output index of the row's output 0
            const i64 n_rel = q_row * g.down - a.n0;                             // input index (rel. x[0]) of the super-period
            for (int i = 0; i < g.up; ++i) {
                const int ga = g.g0[i], gb = g.g1[i];
                float y[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    float va = 0.f, vb = 0.f, vx = 0.f;
#pragma unroll
                    for (int q = 0; q < KM_MAXG; ++q) {
                        va = (q == ga) ? sums[q][j] : va;
                        vb = (q == gb) ? sums[q][j] : vb;
                    }
                    if (gb >= 0) {
                        vx = __shfl_down_sync(0xffffffffu, vb, 1);
                        if (lane == 31) vx = xn[gb * 8 + j];
                    }
                    y[j] = va + vx;
                }
                const i64 idx = ob + i;
                if (row_ok && idx >= 0 && idx < a.n_out) {
#pragma unroll
                    for (int r = 0; r < 4; ++r) {
                        if (!c_row[r]) continue;
                        const u64 ph = acc[r] + inc[r] * (u64)(n_rel + g.o[i]);
                        const float ang = (float)(int)(ph >> 32) * 1.4629180792671596e-09f;      // 2*pi*2^-32
                        float sn, cs;
                        __sincosf(ang, &sn, &cs);
                        float2 out;
                        out.x = fmaf(y[2 * r], cs, y[2 * r + 1] * sn);           // (re + j im)(cos - j sin)
                        out.y = fmaf(y[2 * r + 1], cs, -y[2 * r] * sn);
                        c_row[r][idx] = out;
                        if (bb_row[r]) bb_row[r][idx] = out;
                    }
                }
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem));
}

// ---- host side ----------------------------------------------------------------------------------------------------------
struct K1MmaPlan {
    int up, down, lp, n_rx;
    int n_groups;                              // receiver groups of 4
    // per shift (0 / 1: the row starts one sample early so that its address is 16-byte aligned)
    int n_steps[2], ng[2];
    int g0[2][8], g1[2][8], o[8];
    unsigned b_bytes[2];
    unsigned char *d_img[2];                   // [n_groups][KM_TAB_BYTES + b_bytes]: step table + B image per receiver group
    bool attr_done[64];
};


struct KmPiece { int out, dr, t0, t1; };       // samples [t0, t1) of the row (dr = 1: the row AFTER the output's own)

// pieces of the UP windows for row shift sh, ordered so that the pieces alive at any k are contiguous (by start, then end)
static bool km_pieces(int up, int down, int lp, int sh, std::vector<KmPiece> &pc, int *n_steps) {
    pc.clear();
    for (int i = 0; i < up; ++i) {
        const int o = (int)(((i64)i * down) / up);
        const int t0 = o + sh, t1 = o + sh + lp;                             // window [t0, t1) in row coordinates
        if (t1 <= down + sh) {                                               // a shifted row is read one sample past its period
            pc.push_back({i, 0, t0, t1});
        } else {
            if (t1 - down > down) return false;                              // would span three rows
            pc.push_back({i, 0, t0, down});
            pc.push_back({i, 1, 0, t1 - down});
        }
    }
    std::sort(pc.begin(), pc.end(), [](const KmPiece &x, const KmPiece &y) { return x.t0 != y.t0 ? x.t0 < y.t0 : x.t1 < y.t1; });
    int kmax = 0;
    for (auto &p : pc) kmax = p.t1 > kmax ? p.t1 : kmax;
    *n_steps = (2 * kmax + 7) / 8;
    return (int)pc.size() <= KM_MAXG && *n_steps <= KM_MAX_STEPS;
}

int k1_mma_supported(int up, int down, int lp, int n_rx) {
    if (up < 1 || up > 8 || (down & 1) || n_rx < 1 || lp < 2) return 0;
    if (!km_encode_fn()) return 0;
    for (int sh = 0; sh < 2; ++sh) {
        std::vector<KmPiece> pc;
        int ns;
        if (!km_pieces(up, down, lp, sh, pc, &ns)) return 0;
        // B image: per step, the live pieces' columns x 8 floats
        size_t bytes = 0;
        for (int s = 0; s < ns; ++s) {
            int lo = 99, hi = -1;
            for (int q = 0; q < (int)pc.size(); ++q)
                if (2 * pc[q].t0 < 8 * (s + 1) && 2 * pc[q].t1 > 8 * s) { lo = q < lo ? q : lo; hi = q > hi ? q : hi; }
            if (hi < 0) { lo = 0; hi = 0; }
            bytes += (size_t)(hi - lo + 1) * KM_GW * 32;
        }
        if (bytes > KM_B_MAX) return 0;
    }
    return 1;
}


K1MmaPlan *k1_mma_plan_create(int up, int down, int lp, int n_rx) {
    if (!k1_mma_supported(up, down, lp, n_rx)) return nullptr;
    K1MmaPlan *p = new K1MmaPlan();
    p->up = up; p->down = down; p->lp = lp; p->n_rx = n_rx;
    p->n_groups = (n_rx + 3) / 4;
    p->d_img[0] = p->d_img[1] = nullptr;
    for (int i = 0; i < 64; ++i) p->attr_done[i] = false;
    return p;
}

void k1_mma_plan_destroy(K1MmaPlan *p) {
    if (!p) return;
    cudaFree(p->d_img[0]); cudaFree(p->d_img[1]);
    delete p;
}

// g_host: folded taps [n_rx][up][lp_pad] (what k1_fast reads).  Builds, for both row shifts and every receiver group, the step
// table and the B image: B[k][col], k = 2 t (+1 for Im x), col = piece*16 + half*8 + 2 r (+1 for the Im output):
//   Re y += g_re x_re - g_im x_im,  Im y += g_im x_re + g_re x_im,  tap index j = (t1 - 1) - t within the piece's window.
int k1_mma_upload_taps(K1MmaPlan *p, const float2 *g_host, int lp_pad, cudaStream_t st) {
    for (int sh = 0; sh < 2; ++sh) {
        std::vector<KmPiece> pc;
        int ns;
        if (!km_pieces(p->up, p->down, p->lp, sh, pc, &ns)) { pysdr_set_error("k1_mma: geometry not supported"); return PYSDR_ERR_ARG; }
        p->n_steps[sh] = ns; p->ng[sh] = (int)pc.size();
        for (int i = 0; i < 8; ++i) { p->g0[sh][i] = -1; p->g1[sh][i] = -1; }
        for (int i = 0; i < p->up; ++i) p->o[i] = (int)(((i64)i * p->down) / p->up);
        for (int q = 0; q < (int)pc.size(); ++q) (pc[q].dr ? p->g1[sh] : p->g0[sh])[pc[q].out] = q;
        // step table (smem offsets are relative to the CTA's B base: the kernel's smem base is 1024-aligned and fixed)
        std::vector<KmStep> tab(KM_MAX_STEPS);
        std::vector<int> s_lo(ns), s_n(ns);
        std::vector<size_t> s_off(ns);
        size_t bytes = 0;
        for (int s = 0; s < ns; ++s) {
            int lo = 99, hi = -1;
            for (int q = 0; q < (int)pc.size(); ++q)
                if (2 * pc[q].t0 < 8 * (s + 1) && 2 * pc[q].t1 > 8 * s) { lo = q < lo ? q : lo; hi = q > hi ? q : hi; }
            if (hi < 0) { lo = 0; hi = 0; }
            s_lo[s] = lo; s_n[s] = (hi - lo + 1) * KM_GW; s_off[s] = bytes;
            bytes += (size_t)s_n[s] * 32;
        }
        p->b_bytes[sh] = (unsigned)bytes;
        const size_t img_bytes = KM_TAB_BYTES + bytes;
        std::vector<unsigned char> img((size_t)p->n_groups * img_bytes, 0);
        for (int grp = 0; grp < p->n_groups; ++grp) {
            unsigned char *base = img.data() + (size_t)grp * img_bytes;
            KmStep *t = (KmStep *)base;
            float *B = (float *)(base + KM_TAB_BYTES);
            for (int s = 0; s < ns; ++s) {
                const int N = s_n[s];
                // shared-memory descriptor (K-major, no swizzle): start (16-byte units; here the offset within the image, the
                // kernel adds the image's shared-memory address) | LBO = bytes between the two 16-byte K chunks | SBO = bytes
                // between 8-column groups | version 1
                t[s].bdesc = ((unsigned long long)((N * 16) >> 4) << 16) | ((unsigned long long)(128 >> 4) << 32) | (1ull << 46);
                t[s].idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((unsigned)(N >> 3) << 17) | ((128u >> 4) << 24);
                t[s].dcol = (unsigned)(s_lo[s] * KM_GW);
                t[s].bdesc |= (unsigned long long)(s_off[s] >> 4);
                float *slab = B + s_off[s] / 4;
                for (int q = s_lo[s]; q < s_lo[s] + N / KM_GW; ++q) {
                    const KmPiece &P = pc[q];
                    const int tap_phase = (int)(((i64)P.out * p->down) % p->up);
                    // window in row coordinates of the output's own row: [w0, w0 + lp); the piece of the next row is shifted by -down
                    const int w0 = p->o[P.out] + sh - (P.dr ? p->down : 0);
                    for (int kk = 0; kk < 8; ++kk) {
                        const int k = 8 * s + kk, tt = k >> 1, im_x = k & 1;
                        if (tt < P.t0 || tt >= P.t1) continue;
                        const int j = (w0 + p->lp - 1) - tt;
                        if (j < 0 || j >= p->lp) continue;
                        for (int r = 0; r < 4; ++r) {
                            const int rx = grp * 4 + r;
                            if (rx >= p->n_rx) continue;
                            const float2 gg = g_host[((size_t)rx * p->up + tap_phase) * lp_pad + j];
                            const float v_re = im_x ? -gg.y : gg.x;            // column Re y
                            const float v_im = im_x ? gg.x : gg.y;             // column Im y
                            const float vals[2] = {v_re, v_im};
                            for (int c = 0; c < 2; ++c) {
                                const float h = km_tf32_hi(vals[c]), l = vals[c] - h;
                                const int n_hi = (q - s_lo[s]) * KM_GW + 2 * r + c, n_lo = n_hi + 8;
                                // canonical K-major no-swizzle layout: [chunk = kk/4][n/8][n%8][kk%4]
                                slab[(kk >> 2) * (N * 4) + (n_hi >> 3) * 32 + (n_hi & 7) * 4 + (kk & 3)] = h;
                                slab[(kk >> 2) * (N * 4) + (n_lo >> 3) * 32 + (n_lo & 7) * 4 + (kk & 3)] = l;
                            }
                        }
                    }
                }
            }
        }
        if (p->d_img[sh]) CUDA_TRY(cudaFree(p->d_img[sh]));
        CUDA_TRY(cudaMalloc(&p->d_img[sh], img.size()));
        CUDA_TRY(cudaMemcpyAsync(p->d_img[sh], img.data(), img.size(), cudaMemcpyHostToDevice, st));
        CUDA_TRY(cudaStreamSynchronize(st));
    }
    return PYSDR_OK;
}

int k1_launch_mma(K1MmaPlan *p, const K1Args &a, i64 min_rows, cudaStream_t st, int *used, int *launches) {
    *used = 0;
    if (!p || !p->d_img[0] || a.real_input || a.n_out <= 0 || a.up != p->up || a.down != p->down || a.lp != p->lp) return PYSDR_OK;
    const i64 q_first = a.m0 / a.up;
    // first tensor-core super-period: its row (lp-1 samples of look-back, +1 for the alignment shift) must start inside x
    i64 q_a = (a.n0 + a.lp + a.down - 1) / a.down;
    if (q_a < q_first) q_a = q_first;
    const i64 r0 = q_a * a.down - (a.lp - 1) - a.n0;                           // row 0's first sample, relative to x[0]
    const int sh = (int)((((uintptr_t)(a.x + r0)) >> 3) & 1);                  // start one sample early when that is the aligned one
    if ((((uintptr_t)(a.x + r0 - sh)) & 15) != 0) return PYSDR_OK;             // x itself is not 8-byte aligned: not ours
    const int n_steps = p->n_steps[sh], n_chunks = (n_steps + 7) / 8;
    const i64 span = (i64)n_chunks * 32;                                       // samples a row's boxes read
    const i64 avail = a.n_in - (r0 - sh) - span;
    if (avail < 0) return PYSDR_OK;
    const i64 rows_avail = avail / a.down + 1;
    const i64 rows_out = rows_avail - 1;                                       // the last row only completes its predecessor
    if (rows_out < min_rows || rows_out < 1) return PYSDR_OK;
    KmEncodeFn enc = km_encode_fn();
    if (!enc) return PYSDR_OK;

    CUtensorMap tmap;
    cuuint64_t dims[2] = {(cuuint64_t)n_chunks * 64, (cuuint64_t)rows_avail};
    cuuint64_t strides[1] = {(cuuint64_t)a.down * 8};
    cuuint32_t box[2] = {32, KM_ROWS};
    cuuint32_t estr[2] = {1, 1};
    const CUresult cr = enc(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void *)(a.x + r0 - sh), dims, strides, box, estr,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (cr != CUDA_SUCCESS) { pysdr_set_error("k1_mma: cuTensorMapEncodeTiled failed (%d)", (int)cr); return PYSDR_ERR_CUDA; }

    KmGeom g;
    memset(&g, 0, sizeof(g));
    g.n_steps = n_steps; g.n_chunks = n_chunks; g.ng = p->ng[sh]; g.up = a.up; g.down = a.down;
    for (int i = 0; i < 8; ++i) { g.o[i] = p->o[i]; g.g0[i] = p->g0[sh][i]; g.g1[i] = p->g1[sh][i]; }
    g.rows_out = rows_out;
    g.n_tiles = (rows_out + KM_ADV - 1) / KM_ADV;
    g.q_a = q_a;
    {
        i64 lo = q_a * a.up - a.m0, hi = (q_a + rows_out) * a.up - a.m0;
        g.out_lo = lo < 0 ? 0 : (lo > a.n_out ? a.n_out : lo);
        g.out_hi = hi < g.out_lo ? g.out_lo : (hi > a.n_out ? a.n_out : hi);
    }
    g.b_bytes = p->b_bytes[sh];
    const size_t img_bytes = KM_TAB_BYTES + (size_t)p->b_bytes[sh];
    const size_t smem = 1008 + KM_OFF_B + p->b_bytes[sh] + 2 * 4 * KM_MAXG * 8 * 4 + 20 * 8;   // the dynamic window is 16-byte aligned
    if (smem > 227 * 1024) return PYSDR_OK;
    const int dev = pysdr_device() & 63;
    if (!p->attr_done[dev]) {
        CUDA_TRY(cudaFuncSetAttribute(k1_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        p->attr_done[dev] = true;
    }
    const int sms = pysdr_sm_count();
    const int grid = (int)(g.n_tiles < sms ? g.n_tiles : sms);
    (void)q_first;
    for (int grp = 0; grp < p->n_groups; ++grp) {
        g.rx0 = grp * 4;
        g.img = p->d_img[sh] + (size_t)grp * img_bytes;
        CUDA_TRY(launch_pdl(k1_mma_kernel, dim3(grid), dim3(KM_THREADS), smem, st, tmap, a, g));
        if (launches) ++*launches;
    }
    *used = 1;
    return PYSDR_OK;
}
