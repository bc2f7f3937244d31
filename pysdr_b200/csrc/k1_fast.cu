// K1 (fast variant): fused NCO-mix + polyphase FIR decimate, tap-stationary, sm_100a.
//
// Same arithmetic and indexing contract as k1_generic.cu (reference: dsp.Receiver.demod_data's
// lo/dec stages, receiver.py:235,822,866).  Design:
//
//  * persistent grid, one 384-thread CTA per SM; input streamed once from HBM into a 3-stage shared
//    memory ring by 1-D bulk async copies (cp.async.bulk -> SASS UBLKCP) completing on mbarriers;
//  * a tile = S "super-periods" (DOWN inputs -> UP outputs each); a warp-task = one output phase i,
//    M consecutive super-periods, NRXP receivers.  The warp keeps the folded complex taps of its
//    phase for all NRXP receivers in REGISTERS (lane l owns taps l, l+32, ...), so the inner loop is
//    one LDS.64 of x per 4*NRXP FMAs;
//  * the 32 lanes' partial dot products (32 accumulators per lane) are summed with a butterfly whose
//    first steps need no selects: lane bit 4 swizzles which super-period a slot accumulates (half-warp
//    granular -> conflict-free LDS), lane bits below it swizzle which receiver's taps a slot holds;
//  * one sincospi per OUTPUT sample from the exact u64 phase accumulator de-rotates the result.
#include "common.cuh"

#define K1F_WARPS 12
#define K1F_THREADS (K1F_WARPS * 32)
#ifndef K1F_STAGES
#define K1F_STAGES 3
#endif
#ifndef K1F_STAGE_ELEMS
#define K1F_STAGE_ELEMS 9216          /* float2 per stage: 72 KB */
#endif

struct FastGeom {
    int S;                 // super-periods per tile
    i64 n_tiles;
    i64 q_first;           // absolute super-period index of tile 0
    int rx0;               // first receiver of this launch's group
    int need_pad;          // lp_pad - 1
};

__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(unsigned bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(unsigned bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE;\n"
        "bra WAIT_LOOP;\n"
        "DONE:\n"
        "}\n" ::"r"(bar), "r"(parity)
        : "memory");
}
__device__ __forceinline__ void bulk_g2s(unsigned dst, const void *src, unsigned bytes, unsigned bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}

template <int NRXP>
struct Bits {
    static constexpr int RB = (NRXP == 4) ? 2 : (NRXP == 2 ? 1 : 0);
    static constexpr int SB = 3 - RB;          // 8 accumulator slots = 2^SB outputs x NRXP receivers
    static constexpr int M = 1 << SB;
};

__device__ __forceinline__ void mbar_arrive(unsigned bar) {
    asm volatile("mbarrier.arrive.release.cta.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}

// REAL: the input samples are real (imaginary parts identically zero, e.g. the FM discriminator output that feeds the
// WFM resampler rows): the FFMA2 stream that multiplies the taps by Im x is skipped — half the FMAs.
template <int NRXP, int TPL, bool REAL = false>
__global__ void __launch_bounds__(K1F_THREADS, 1) k1_fast_kernel(const K1Args a, const FastGeom g) {
    constexpr int RB = Bits<NRXP>::RB;
    constexpr int SB = Bits<NRXP>::SB;
    constexpr int M = Bits<NRXP>::M;
    constexpr int SLOW = 1 << (SB - 1);          // # values of the non-swizzled low s bits

    extern __shared__ __align__(16) unsigned char smem[];
    unsigned long long *bars = (unsigned long long *)smem;          // K1F_STAGES "full" mbarriers
    int *done_cnt = (int *)(smem + 32);                              // K1F_STAGES consumer counters
    float2 *stage0 = (float2 *)(smem + 64);

    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int warp = tid >> 5;
    const int up = a.up, down = a.down;
    const int need_pad = g.need_pad;
    const int par = (int)(((unsigned long long)a.x >> 3) & 1ull);
    const int S = g.S;

    if (tid == 0) {
        for (int s = 0; s < K1F_STAGES; ++s) { mbar_init(smem_u32(&bars[s]), 1); done_cnt[s] = 0; }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    pdl_trigger();
    pdl_wait();                      // x may come from a conversion kernel; C is still read by the previous call's back kernel
    __syncthreads();
    if (a.zero_c_hist && g.rx0 == 0) {                                     // a pending seek(): nothing before this sample
        for (int r = blockIdx.x; r < a.n_rx; r += gridDim.x) {
            float2 *row = a.c_out + (size_t)r * a.c_stride;
            for (int e = tid; e < a.hc; e += K1F_THREADS) row[e] = make_float2(0.f, 0.f);
        }
    }

    // tile geometry ---------------------------------------------------------------------------------
    auto tile_a2 = [&](i64 T, i64 &a2, int &cnt2) {
        const i64 q0 = g.q_first + T * S;
        const i64 ar = q0 * down - need_pad - a.n0;              // rel index of first needed element
        a2 = ar - ((ar + par) & 1);                              // 16-byte aligned global address
        const i64 br = (q0 + S) * down - a.n0;
        int cnt = (int)(br - a2);
        cnt2 = cnt + (cnt & 1);
    };

    // Fill stage `st` with tile T.  Called by ONE warp (all 32 lanes).  Every tile completes exactly one phase
    // of bars[st]: interior tiles by the bulk copy's complete_tx, edge tiles (stream start/end, zero fill,
    // carried history) by an explicit arrive after plain stores.
    auto produce = [&](i64 T, int st) {
        if (T >= g.n_tiles) return;
        i64 a2; int cnt2;
        tile_a2(T, a2, cnt2);
        float2 *dst = stage0 + (size_t)st * K1F_STAGE_ELEMS;
        const unsigned bar = smem_u32(&bars[st]);
        // part of the tile that exists in x[0, n_in), shrunk to 16-byte aligned ends -> one bulk copy
        i64 blo = a2 > 0 ? a2 : 0, bhi = (a2 + cnt2 < a.n_in) ? a2 + cnt2 : a.n_in;
        blo += (blo + par) & 1;
        bhi -= (bhi + par) & 1;
        const bool bulk = bhi > blo;
        if (!bulk) { blo = a2; bhi = a2; }
        // everything else (carried history before x[0], zero fill past the end, odd edge elements): plain stores
        const int n_head = (int)(blo - a2), n_tail = (int)(a2 + cnt2 - bhi);
        for (int e = lane; e < n_head + n_tail; e += 32) {
            const int ee = e < n_head ? e : (int)(bhi - a2) + (e - n_head);
            const i64 rel = a2 + ee;
            float2 v = make_float2(0.f, 0.f);
            if (rel >= 0) { if (rel < a.n_in) v = a.x[rel]; }
            else if (rel >= -(i64)a.need && a.hist) v = a.hist[a.need + rel];
            dst[ee] = v;
        }
        __syncwarp();
        if (lane == 0) {
            if (bulk) {
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                const unsigned bytes = (unsigned)(bhi - blo) * 8u;
                mbar_expect_tx(bar, bytes);
                bulk_g2s(smem_u32(dst + (blo - a2)), a.x + blo, bytes, bar);
            } else {
                mbar_arrive(bar);
            }
        }
    };

    // per-lane constants ------------------------------------------------------------------------------
    const int sw_s = (lane >> 4) & 1;                                   // flips the top s bit of a slot
    const int sw_r = (NRXP > 1) ? ((lane >> (4 - RB)) & (NRXP - 1)) : 0;
    const int s_log = (((lane >> 4) & 1) << (SB - 1)) | ((lane >> 2) & (SLOW - 1));   // output this lane stores
    const int r_log = (NRXP > 1) ? ((lane >> (4 - RB)) & (NRXP - 1)) : 0;
    const int rx = g.rx0 + r_log;
    const bool rx_ok = rx < a.n_rx;
    const u64 my_inc = rx_ok ? a.inc[rx] : 0ull;
    const u64 my_acc = rx_ok ? a.acc[rx] : 0ull;
    float2 *my_out = nullptr;                                           // even lanes -> C memory, odd lanes -> rx.iq copy
    if (rx_ok) {
        if ((lane & 3) == 0) my_out = a.c_out + (size_t)rx * a.c_stride + a.hc;
        else if ((lane & 3) == 1 && a.bb_out) my_out = a.bb_out + (size_t)rx * a.bb_stride;
    }

    float2 tap[NRXP][TPL];                                              // (g_re, g_im) pairs, FFMA2 operands
    int cur_i = -1, o_i = 0;

    const i64 T0 = blockIdx.x;
    const i64 Tstep = gridDim.x;
    if (warp < K1F_STAGES) produce(T0 + (i64)warp * Tstep, warp);

    const int tasks_per_tile = up * (S / M);
    const int di = K1F_WARPS % up, ds = K1F_WARPS / up;
    const int i_first = warp % up, sblk_first = warp / up;
    // ---- software-pipelined reduction: state of the task whose FMAs have just finished ---------------------
    float p_re[8], p_im[8];
#pragma unroll
    for (int v = 0; v < 8; ++v) { p_re[v] = 0.f; p_im[v] = 0.f; }
    float2 *p_ptr = nullptr;
    float p_ang = 0.f;
    auto reduce_store = [&]() {
        float are[8], aim[8];
#pragma unroll
        for (int v = 0; v < 8; ++v) { are[v] = p_re[v]; aim[v] = p_im[v]; }
        // butterfly over lanes: bit 4 and the RB bits below it are swizzled (no selects) ...
#pragma unroll
        for (int step = 0; step < 1 + RB; ++step) {
            const int off = 16 >> step;
            const int H = 4 >> step;                                    // slots kept (per array)
#pragma unroll
            for (int v = 0; v < H; ++v) {
                are[v] += __shfl_xor_sync(0xffffffffu, are[H + v], off);
                aim[v] += __shfl_xor_sync(0xffffffffu, aim[H + v], off);
            }
        }
        // ... the remaining low s bits use selects
#pragma unroll
        for (int step = 1 + RB; step < 3; ++step) {
            const int off = 16 >> step;
            const int H = 4 >> step;
            const bool upper = (lane & off) != 0;
#pragma unroll
            for (int v = 0; v < H; ++v) {
                const float sre = upper ? are[v] : are[H + v];
                const float kre = upper ? are[H + v] : are[v];
                const float sim = upper ? aim[v] : aim[H + v];
                const float kim = upper ? aim[H + v] : aim[v];
                are[v] = kre + __shfl_xor_sync(0xffffffffu, sre, off);
                aim[v] = kim + __shfl_xor_sync(0xffffffffu, sim, off);
            }
        }
        // ... and lane bits 1:0 are plain sums (every lane of a quad ends with the full dot product)
        float sr = are[0], si = aim[0];
        sr += __shfl_xor_sync(0xffffffffu, sr, 2);
        si += __shfl_xor_sync(0xffffffffu, si, 2);
        sr += __shfl_xor_sync(0xffffffffu, sr, 1);
        si += __shfl_xor_sync(0xffffffffu, si, 1);
        // de-rotate by the exact LO phase of the output's newest input sample and store
        if (p_ptr) {
            float sn, cs;
            __sincosf(p_ang, &sn, &cs);
            float2 y;
            y.x = fmaf(sr, cs, si * sn);                                // (sr + j si)(cos - j sin)
            y.y = fmaf(si, cs, -sr * sn);
            *p_ptr = y;
        }
    };

    unsigned phase_bits = 0;
    // per-tile state advanced incrementally (no 64-bit multiplies / divisions in the loop)
    const i64 q_start = g.q_first + T0 * S;
    i64 relb = q_start * down - a.n0;                                    // input index (rel. x[0]) of the tile's first sample
    i64 ob = q_start * up - a.m0;                                        // output index of the tile's first output
    u64 pbase = my_acc + my_inc * (u64)relb;                             // LO phase at relb
    const i64 d_rel = Tstep * S * down, d_ob = Tstep * S * up;
    const u64 d_ph = my_inc * (u64)d_rel;
    const int tile_outs = S * up;
    const int lane_off = need_pad - lane;
    int st = 0;
    for (i64 T = T0; T < g.n_tiles; T += Tstep) {
        mbar_wait(smem_u32(&bars[st]), (phase_bits >> st) & 1u);
        phase_bits ^= (1u << st);
        const float2 *xs = stage0 + (size_t)st * K1F_STAGE_ELEMS;
        const int shift = (int)((relb - need_pad + par) & 1);            // tile stored from an even-aligned element
        const bool tile_full = ob >= 0 && ob + tile_outs <= a.n_out;     // every output of the tile is in range
        float2 *outp = my_out ? my_out + ob : nullptr;

        int i = i_first, sblk = sblk_first;
        for (int t = warp; t < tasks_per_tile; t += K1F_WARPS) {
            if (i != cur_i) {                                           // (re)load this phase's taps
                const int idn = i * down;
                o_i = idn / up;
                const int p_i = idn - o_i * up;
#pragma unroll
                for (int r = 0; r < NRXP; ++r) {
                    const int rr = g.rx0 + (r ^ sw_r);
                    const float2 *gp = a.g + ((size_t)rr * up + p_i) * a.lp_pad + lane;
#pragma unroll
                    for (int k = 0; k < TPL; ++k) tap[r][k] = (rr < a.n_rx) ? __ldg(gp + 32 * k) : make_float2(0.f, 0.f);
                }
                cur_i = i;
            }
            reduce_store();                                             // previous task's reduction + store
            // Packed FP32x2 FMAs (Blackwell FFMA2): A += (g_re,g_im)*(x_re,x_re), B += (g_re,g_im)*(x_im,x_im)
            //   => re = A.x - B.y, im = A.y + B.x.  One issue slot per two FMAs.
            float2 A[8], B[8];
#pragma unroll
            for (int v = 0; v < 8; ++v) { A[v] = make_float2(0.f, 0.f); B[v] = make_float2(0.f, 0.f); }

            // tap loop outermost, the M outputs of the task innermost: 2*M*NRXP independent FFMA2 chains per tap
            const float2 *xp[M];
#pragma unroll
            for (int stop = 0; stop < 2; ++stop) {
#pragma unroll
                for (int sl = 0; sl < SLOW; ++sl) {
                    const int s_eff = ((stop ^ sw_s) << (SB - 1)) | sl;
                    xp[stop * SLOW + sl] = xs + ((sblk * M + s_eff) * down + o_i + lane_off + shift);
                }
            }
#pragma unroll
            for (int k = 0; k < TPL; ++k) {
#pragma unroll
                for (int stop = 0; stop < 2; ++stop) {
#pragma unroll
                    for (int sl = 0; sl < SLOW; ++sl) {
                        const float2 xv = xp[stop * SLOW + sl][-32 * k];
                        const float2 xrr = make_float2(xv.x, xv.x), xii = make_float2(xv.y, xv.y);
#pragma unroll
                        for (int r = 0; r < NRXP; ++r) {
                            const int slot = (stop << 2) | (r << (2 - RB)) | sl;
                            A[slot] = __ffma2_rn(tap[r][k], xrr, A[slot]);
                        }
                        if constexpr (!REAL) {
#pragma unroll
                            for (int r = 0; r < NRXP; ++r) {
                                const int slot = (stop << 2) | (r << (2 - RB)) | sl;
                                B[slot] = __ffma2_rn(tap[r][k], xii, B[slot]);
                            }
                        }
                    }
                }
            }
            // hand this task's partial sums to the software pipeline: they are reduced across lanes, de-rotated and
            // stored while the NEXT task's FMAs issue (the reduction is latency-bound, the FMA phase pipe-bound)
#pragma unroll
            for (int v = 0; v < 8; ++v) { p_re[v] = A[v].x - B[v].y; p_im[v] = A[v].y + B[v].x; }
            {
                const int qrel = sblk * M + s_log;
                const int oidx = qrel * up + i;                          // output index within the tile
                const bool ok = outp && (tile_full || (ob + oidx >= 0 && ob + oidx < a.n_out));
                p_ptr = ok ? outp + oidx : nullptr;
                const u64 ph = pbase + my_inc * (u64)(unsigned)(qrel * down + o_i);
                p_ang = (float)(int)(ph >> 32) * 1.4629180792671596e-09f;                 // 2*pi*2^-32
            }
            i += di; sblk += ds;
            if (i >= up) { i -= up; sblk += 1; }
        }

        // release the stage: the LAST warp to finish this tile refills it with tile T + STAGES*Tstep
        __syncwarp();
        int last = 0;
        if (lane == 0) {
            __threadfence_block();
            last = (atomicAdd(&done_cnt[st], 1) == K1F_WARPS - 1);
            if (last) done_cnt[st] = 0;
        }
        last = __shfl_sync(0xffffffffu, last, 0);
        if (last) {
            // Interior tiles (almost all): the refill is one aligned bulk copy whose address follows from the running
            // tile offset — no 64-bit multiplies and no edge loop on the warp that is already the last one
            // (0.859 -> 0.842 ms at cfg2).  Measured and rejected here: an extra cp.async.bulk.prefetch.L2 one, two or
            // four tiles ahead (0.865 ms), consuming the arrival count one task later (0.965 ms), and an "empty" mbarrier
            // with a rotating refill warp instead of the atomic counter (0.864 ms): anything that delays the refill by
            // even a fraction of a tile costs more than the atomic's latency.
            const i64 Tn = T + (i64)K1F_STAGES * Tstep;
            if (Tn < g.n_tiles) {
                const i64 rel_n = relb + (i64)K1F_STAGES * d_rel;
                const i64 ar = rel_n - need_pad;
                const i64 a2 = ar - ((ar + par) & 1);
                const int cnt = (int)(rel_n + (i64)S * down - a2);
                const int cnt2 = cnt + (cnt & 1);
                if (a2 >= 0 && a2 + cnt2 <= a.n_in) {
                    if (lane == 0) {
                        const unsigned bar = smem_u32(&bars[st]);
                        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                        mbar_expect_tx(bar, (unsigned)cnt2 * 8u);
                        bulk_g2s(smem_u32(stage0 + (size_t)st * K1F_STAGE_ELEMS), a.x + a2, (unsigned)cnt2 * 8u, bar);
                    }
                } else {
                    produce(Tn, st);
                }
            }
        }
        relb += d_rel; ob += d_ob; pbase += d_ph;
        st = (st + 1 == K1F_STAGES) ? 0 : st + 1;
    }
    reduce_store();                                                     // drain the pipeline
}

// ---------------------------------------------------------------------------------------------------
static int pick_tpl(int lp) {
    const int need = (lp + 31) / 32;
    const int opts[] = {4, 8, 11, 16, 22, 32};
    for (int o : opts) if (need <= o) return o;
    return -1;
}
static int pick_nrxp(int tpl, int n_rx) {
    int nr = 1;
    while (nr * 2 <= 4 && nr < n_rx && tpl * nr * 2 <= 48) nr *= 2;
    return nr;
}
int k1_fast_lp_pad(int lp) {
    const int t = pick_tpl(lp);
    return t > 0 ? 32 * t : (lp + 31) / 32 * 32;
}
int k1_fast_groups(int lp, int n_rx) {
    const int tpl = pick_tpl(lp);
    if (tpl < 0) return 1;
    const int nrxp = pick_nrxp(tpl, n_rx);
    return (n_rx + nrxp - 1) / nrxp;
}
static int pick_S(int down, int lp_pad, int M) {
    i64 s = (K1F_STAGE_ELEMS - lp_pad - 4) / down;
    s = s / M * M;
    if (s > 4096) s = 4096 / M * M;
    return (int)s;
}

int k1_fast_supported(int up, int down, int lp, int n_rx) {
    const int tpl = pick_tpl(lp);
    if (tpl < 0) return 0;
    const int nrxp = pick_nrxp(tpl, n_rx);
    const int M = 8 / nrxp;
    if (pick_S(down, 32 * tpl, M) < M) return 0;
    if ((i64)up * down > (1 << 30)) return 0;
    return 1;
}

template <int NRXP, int TPL, bool REAL = false>
static int launch_one(const K1Args &a, const FastGeom &g, int grid, cudaStream_t st) {
    if constexpr (!REAL && NRXP == 1) {
        if (a.real_input) return launch_one<NRXP, TPL, true>(a, g, grid, st);
    }
    const size_t smem = 64 + sizeof(float2) * (size_t)K1F_STAGES * K1F_STAGE_ELEMS;
    static unsigned long long attr_done = 0ull;          // one bit per device: the attribute is per (function, device)
    const unsigned long long dev_bit = 1ull << (pysdr_device() & 63);
    if (!(attr_done & dev_bit)) {
        CUDA_TRY(cudaFuncSetAttribute(k1_fast_kernel<NRXP, TPL, REAL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_done |= dev_bit;
    }
    CUDA_TRY(launch_pdl(k1_fast_kernel<NRXP, TPL, REAL>, dim3(grid), dim3(K1F_THREADS), smem, st, a, g));
    return PYSDR_OK;
}

template <int NRXP>
static int launch_tpl(int tpl, const K1Args &a, const FastGeom &g, int grid, cudaStream_t st) {
    switch (tpl) {
        case 4: return launch_one<NRXP, 4>(a, g, grid, st);
        case 8: return launch_one<NRXP, 8>(a, g, grid, st);
        case 11: return launch_one<NRXP, 11>(a, g, grid, st);
        case 16: if (NRXP <= 2) return launch_one<(NRXP <= 2 ? NRXP : 1), 16>(a, g, grid, st); break;
        case 22: if (NRXP <= 2) return launch_one<(NRXP <= 2 ? NRXP : 1), 22>(a, g, grid, st); break;
        case 32: if (NRXP == 1) return launch_one<1, 32>(a, g, grid, st); break;
    }
    pysdr_set_error("k1_fast: no instantiation for NRXP=%d TPL=%d", NRXP, tpl);
    return PYSDR_ERR_ARG;
}

int k1_launch_fast(const K1Args &a, cudaStream_t st) {
    if (a.n_out <= 0) return PYSDR_OK;
    const int tpl = pick_tpl(a.lp);
    if (tpl < 0 || a.lp_pad != 32 * tpl) {
        pysdr_set_error("k1_fast: lp=%d lp_pad=%d not laid out for the fast path", a.lp, a.lp_pad);
        return PYSDR_ERR_ARG;
    }
    const int nrxp = pick_nrxp(tpl, a.n_rx);
    const int M = 8 / nrxp;
    FastGeom g;
    g.S = pick_S(a.down, a.lp_pad, M);
    g.need_pad = a.lp_pad - 1;
    g.q_first = a.m0 / a.up;
    const i64 q_last = (a.m0 + a.n_out - 1) / a.up;
    g.n_tiles = (q_last - g.q_first) / g.S + 1;
    const int sms = pysdr_sm_count();
    int grid = (int)(g.n_tiles < sms ? g.n_tiles : sms);
    for (int rx0 = 0; rx0 < a.n_rx; rx0 += nrxp) {
        g.rx0 = rx0;
        int rc;
        if (nrxp == 4) rc = launch_tpl<4>(tpl, a, g, grid, st);
        else if (nrxp == 2) rc = launch_tpl<2>(tpl, a, g, grid, st);
        else rc = launch_tpl<1>(tpl, a, g, grid, st);
        if (rc) return rc;
    }
    return PYSDR_OK;
}
