// fft_smem.cuh — complex FFT of N = 2^n points held in shared memory, radix-16 butterflies in registers.
//
// Building block of K2 (overlap-save AF filtering, the GPU counterpart of the reference's
// dsp.convolver.convolve_fast, receiver.py:862) and K3 (dsp.spectrum, Plotting.py:376-377,462).
//
//   forward  : decimation in frequency, natural-order input -> DIGIT-REVERSED output
//   inverse  : decimation in time, digit-reversed input -> natural-order output (unnormalised)
// so  inverse(H_pos * forward(x))  needs no reordering pass at all: spectra are only ever touched in
// "position" order (fft_pos_to_freq maps position -> frequency bin when a caller needs bins).
//
// Plan: first pass radix 2^(n mod 4) when n is not a multiple of 4, then radix-16 passes.  N/16 threads own
// one radix-16 butterfly per pass.  Shared memory index p is stored at p + (p >> 4) (one pad per 16) which
// keeps every pass's LDS.64/STS.64 conflict-free per half-warp.  Inter-pass twiddles W_Ni^(k*m): one
// sincospif per butterfly, powers by a depth-4 product tree.
#pragma once
#include <cuda_runtime.h>

#define FFT_PAD(p) ((p) + ((p) >> 4))
#define FFT_SMEM_ELEMS(N) ((N) + ((N) >> 4))

// Two interchangeable arithmetic cores.  FFT_PACKED = 1 (default): complex numbers in 64-bit register pairs, FADD2 / FMUL2 /
// FFMA2.  FFT_PACKED = 0: scalar FP32.  The packed core needs aligned register pairs: it wins where registers are capped
// low anyway (AF filter kernel, 80 registers: 104 -> 100 us) and loses where the kernel sits at the register limit (8192-point
// PSD frames, 128 registers: 4 -> 96 bytes of spills, 4.16 -> 4.43 ms), so psd.cu and czt.cu keep the scalar core.
#ifndef FFT_PACKED
#define FFT_PACKED 1
#endif
#if FFT_PACKED
// ---- packed complex arithmetic ------------------------------------------------------------------------------------------
// A complex number lives in one 64-bit register pair (re, im) and is handled with Blackwell's two-wide FP32 instructions
// (PTX add/sub/mul/fma .f32x2 -> SASS FADD2 / FMUL2 / FFMA2): one issue slot per complex add instead of two.  The
// instructions take a scalar-broadcast operand (.F32) and a half-swapped operand (.F32x2.LO_HI) for free, so
//     a * w           = (a.y, a.x) * (w.y, w.y) * (-1, 1) + a * (w.x, w.x)                       3 instructions (4 scalar)
//     e +- (-j) o     = (o.y, o.x) * (+-1, -+1) + e                                             1 instruction each
// The butterflies of these transforms are add-dominated (FADD was 29..34 % of all warp instructions of the AF filter and
// PSD kernels, which were issue-bound at 54..63 % issue utilisation), so halving their issue slots is the lever.
__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return __fadd2_rn(a, b); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) {
    unsigned long long r;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(*(unsigned long long *)&a), "l"(*(unsigned long long *)&b));
    return *(float2 *)&r;
}
__device__ __forceinline__ float2 cswap(float2 a) { return make_float2(a.y, a.x); }
__device__ __forceinline__ float2 cmul(float2 a, float2 b) {                  // a * b
    const float2 t = __fmul2_rn(a, make_float2(b.x, b.x));
    const float2 u = __fmul2_rn(cswap(a), make_float2(b.y, b.y));             // (a.y b.y, a.x b.y)
    return __ffma2_rn(u, make_float2(-1.f, 1.f), t);
}
__device__ __forceinline__ float2 cmul_conj(float2 a, float2 b) {             // a * conj(b)
    const float2 t = __fmul2_rn(a, make_float2(b.x, b.x));
    const float2 u = __fmul2_rn(cswap(a), make_float2(b.y, b.y));
    return __ffma2_rn(u, make_float2(1.f, -1.f), t);
}
// a * (c, s) for a CONSTANT (c, s): the pair (-s, c) is a constant too -> 2 instructions
__device__ __forceinline__ float2 cmul_const(float2 a, float c, float s) {
    return __ffma2_rn(make_float2(a.y, a.y), make_float2(-s, c), __fmul2_rn(make_float2(a.x, a.x), make_float2(c, s)));
}
// a * (-j) when INV is false, a * (+j) when it is true
template <bool INV>
__device__ __forceinline__ float2 cmul_mj(float2 a) {
    return INV ? make_float2(-a.y, a.x) : make_float2(a.y, -a.x);
}

// In-register DFT of R points (R in {2,4,8,16}); natural order in, natural order out.  INV = false: kernel exp(-2 pi i
// jk/R); INV = true: exp(+2 pi i jk/R) (the conjugate transform, unnormalised) — the internal constants change sign, no
// conjugation passes over the data are needed.
template <int R, bool INV>
__device__ __forceinline__ void dft_reg(float2 *v) {
    if constexpr (R == 2) {
        const float2 a = v[0], b = v[1];
        v[0] = cadd(a, b);
        v[1] = csub(a, b);
    } else {
        float2 e[R / 2], o[R / 2];
#pragma unroll
        for (int i = 0; i < R / 2; ++i) { e[i] = v[2 * i]; o[i] = v[2 * i + 1]; }
        dft_reg<R / 2, INV>(e);
        dft_reg<R / 2, INV>(o);
        const float c1 = 0.92387953251128674f, s1 = 0.38268343236508977f, r2 = 0.70710678118654752f;
        const float sg = INV ? 1.f : -1.f;                                      // sign of the imaginary part of W_16^k
#pragma unroll
        for (int k = 0; k < R / 2; ++k) {
            const int kk = k * (16 / R);                                        // t = o[k] * W_16^kk
            if (kk == 0) {
                v[k] = cadd(e[k], o[k]);
                v[k + R / 2] = csub(e[k], o[k]);
            } else if (kk == 4) {                                               // W = -+j : e +- (-+j) o, one FFMA2 each
                const float2 so = cswap(o[k]);
                v[k] = __ffma2_rn(so, make_float2(-sg, sg), e[k]);
                v[k + R / 2] = __ffma2_rn(so, make_float2(sg, -sg), e[k]);
            } else {
                float2 t;
                if (kk == 2) {                                                  // r2 (1 -+ j):  r2 * (o + (-+j) o)
                    t = __fmul2_rn(__ffma2_rn(cswap(o[k]), make_float2(-sg, sg), o[k]), make_float2(r2, r2));
                } else if (kk == 6) {                                           // r2 (-1 -+ j): -r2 * (o - (-+j) o)
                    t = __fmul2_rn(__ffma2_rn(cswap(o[k]), make_float2(sg, -sg), o[k]), make_float2(-r2, -r2));
                } else {
                    const float c = (kk == 1) ? c1 : (kk == 3) ? s1 : (kk == 5) ? -s1 : -c1;
                    const float sn = ((kk == 1) ? s1 : (kk == 3) ? c1 : (kk == 5) ? c1 : s1) * sg;
                    t = cmul_const(o[k], c, sn);
                }
                v[k] = cadd(e[k], t);
                v[k + R / 2] = csub(e[k], t);
            }
        }
    }
}

// powers w[k] = w1^k, k = 1..R-1, by a shallow product tree (error ~ log2(k) ulp instead of k ulp)
template <int R>
__device__ __forceinline__ void twiddle_powers(float2 w1, float2 *w) {
    w[1] = w1;
    if constexpr (R > 2) { w[2] = cmul(w1, w1); w[3] = cmul(w[2], w1); }
    if constexpr (R > 4) {
        w[4] = cmul(w[2], w[2]);
#pragma unroll
        for (int k = 5; k < 8; ++k) w[k] = cmul(w[4], w[k - 4]);
    }
    if constexpr (R > 8) {
        w[8] = cmul(w[4], w[4]);
#pragma unroll
        for (int k = 9; k < 16; ++k) w[k] = cmul(w[8], w[k - 8]);
    }
}

// One pass over all blocks of length NI (NI = remaining transform length at this pass), radix R.
template <int N, int NI, int R, int T, bool INV>
__device__ __forceinline__ void fft_pass(float2 *s, int tid) {
    constexpr int STRIDE = NI / R;
#pragma unroll
    for (int u0 = 0; u0 < N / R; u0 += T) {
        const int u = u0 + tid;
        if ((N / R) % T != 0 && u >= N / R) break;
        const int blk = u / STRIDE, m = u - blk * STRIDE;
        const int base = blk * NI + m;
        float2 v[R];
#pragma unroll
        for (int q = 0; q < R; ++q) v[q] = s[FFT_PAD(base + q * STRIDE)];
        float2 w[R];
        if constexpr (STRIDE > 1) {
            // W_NI^(k*m), k = 1..R-1: one accurate sincospi + a depth-4 product tree.  (A table of W_N^j gathered
            // with 15 uncoalesced loads per butterfly was measured 35 % slower on B200.)
            float sn, cs;
            sincospif(-2.0f * (float)m / (float)NI, &sn, &cs);
            twiddle_powers<R>(make_float2(cs, sn), w);
        }
        if constexpr (!INV) {
            dft_reg<R, false>(v);
            if constexpr (STRIDE > 1) {
#pragma unroll
                for (int k = 1; k < R; ++k) v[k] = cmul(v[k], w[k]);
            }
        } else {
            if constexpr (STRIDE > 1) {
#pragma unroll
                for (int k = 1; k < R; ++k) v[k] = cmul_conj(v[k], w[k]);
            }
            dft_reg<R, true>(v);
        }
#pragma unroll
        for (int q = 0; q < R; ++q) s[FFT_PAD(base + q * STRIDE)] = v[q];
    }
    __syncthreads();
}

#else
__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
    return make_float2(fmaf(a.x, b.x, -a.y * b.y), fmaf(a.x, b.y, a.y * b.x));
}
__device__ __forceinline__ float2 cmul_conj(float2 a, float2 b) {          // a * conj(b)
    return make_float2(fmaf(a.x, b.x, a.y * b.y), fmaf(a.y, b.x, -a.x * b.y));
}
__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }

// W_16^k = exp(-2*pi*i*k/16), k = 0..7
__device__ __forceinline__ float2 w16(int k) {
    const float c1 = 0.92387953251128674f, s1 = 0.38268343236508977f, r2 = 0.70710678118654752f;
    switch (k) {
        case 0: return make_float2(1.f, 0.f);
        case 1: return make_float2(c1, -s1);
        case 2: return make_float2(r2, -r2);
        case 3: return make_float2(s1, -c1);
        case 4: return make_float2(0.f, -1.f);
        case 5: return make_float2(-s1, -c1);
        case 6: return make_float2(-r2, -r2);
        default: return make_float2(-c1, -s1);
    }
}

// In-register forward DFT of R points (R in {2,4,8,16}); natural order in, natural order out.
template <int R>
__device__ __forceinline__ void dft_fwd_reg(float2 *v) {
    if constexpr (R == 2) {
        const float2 a = v[0], b = v[1];
        v[0] = cadd(a, b);
        v[1] = csub(a, b);
    } else {
        float2 e[R / 2], o[R / 2];
#pragma unroll
        for (int i = 0; i < R / 2; ++i) { e[i] = v[2 * i]; o[i] = v[2 * i + 1]; }
        dft_fwd_reg<R / 2>(e);
        dft_fwd_reg<R / 2>(o);
#pragma unroll
        for (int k = 0; k < R / 2; ++k) {
            float2 t;
            const int kk = k * (16 / R);
            if (kk == 0) t = o[k];
            else if (kk == 4) t = make_float2(o[k].y, -o[k].x);                 // * (-j)
            else t = cmul(o[k], w16(kk));
            v[k] = cadd(e[k], t);
            v[k + R / 2] = csub(e[k], t);
        }
    }
}

template <int R>
__device__ __forceinline__ void conj_all(float2 *v) {
#pragma unroll
    for (int i = 0; i < R; ++i) v[i].y = -v[i].y;
}

// powers w[k] = w1^k, k = 1..R-1, by a shallow product tree (error ~ log2(k) ulp instead of k ulp)
template <int R>
__device__ __forceinline__ void twiddle_powers(float2 w1, float2 *w) {
    w[1] = w1;
    if constexpr (R > 2) { w[2] = cmul(w1, w1); w[3] = cmul(w[2], w1); }
    if constexpr (R > 4) {
        w[4] = cmul(w[2], w[2]);
#pragma unroll
        for (int k = 5; k < 8; ++k) w[k] = cmul(w[4], w[k - 4]);
    }
    if constexpr (R > 8) {
        w[8] = cmul(w[4], w[4]);
#pragma unroll
        for (int k = 9; k < 16; ++k) w[k] = cmul(w[8], w[k - 8]);
    }
}

// One pass over all blocks of length NI (NI = remaining transform length at this pass), radix R.
template <int N, int NI, int R, int T, bool INV>
__device__ __forceinline__ void fft_pass(float2 *s, int tid) {
    constexpr int STRIDE = NI / R;
#pragma unroll
    for (int u0 = 0; u0 < N / R; u0 += T) {
        const int u = u0 + tid;
        if ((N / R) % T != 0 && u >= N / R) break;
        const int blk = u / STRIDE, m = u - blk * STRIDE;
        const int base = blk * NI + m;
        float2 v[R];
#pragma unroll
        for (int q = 0; q < R; ++q) v[q] = s[FFT_PAD(base + q * STRIDE)];
        float2 w[R];
        if constexpr (STRIDE > 1) {
            // W_NI^(k*m), k = 1..R-1: one accurate sincospi + a depth-4 product tree.  (A table of W_N^j gathered
            // with 15 uncoalesced loads per butterfly was measured 35 % slower on B200.)
            float sn, cs;
            sincospif(-2.0f * (float)m / (float)NI, &sn, &cs);
            twiddle_powers<R>(make_float2(cs, sn), w);
        }
        if constexpr (!INV) {
            dft_fwd_reg<R>(v);
            if constexpr (STRIDE > 1) {
#pragma unroll
                for (int k = 1; k < R; ++k) v[k] = cmul(v[k], w[k]);
            }
        } else {
            if constexpr (STRIDE > 1) {
#pragma unroll
                for (int k = 1; k < R; ++k) v[k] = cmul_conj(v[k], w[k]);
            }
            conj_all<R>(v);
            dft_fwd_reg<R>(v);
            conj_all<R>(v);
        }
#pragma unroll
        for (int q = 0; q < R; ++q) s[FFT_PAD(base + q * STRIDE)] = v[q];
    }
    __syncthreads();
}

#endif

template <int N>
struct FftPlan {
    static constexpr int LOG2 = (N == 16 ? 4 : N == 32 ? 5 : N == 64 ? 6 : N == 128 ? 7 : N == 256 ? 8 : N == 512 ? 9 : N == 1024 ? 10
                                 : N == 2048 ? 11 : N == 4096 ? 12 : N == 8192 ? 13 : N == 16384 ? 14 : -1);
    static_assert(LOG2 > 0, "unsupported FFT size");
    static constexpr int R0 = 1 << (LOG2 % 4);              // 1 (absent), 2, 4 or 8
    static constexpr int N16 = LOG2 / 4;                    // number of radix-16 passes
    static constexpr int THREADS = (N / 16 >= 32) ? N / 16 : 32;
};

// forward (INV=false) or inverse (INV=true) transform of the N points in s (padded layout), T = FftPlan<N>::THREADS
template <int N, bool INV>
__device__ __forceinline__ void fft_smem(float2 *s, int tid) {
    using P = FftPlan<N>;
    constexpr int T = P::THREADS;
    constexpr int N1 = N / P::R0;                           // length after the odd first pass
    if constexpr (!INV) {
        if constexpr (P::R0 > 1) fft_pass<N, N, P::R0, T, false>(s, tid);
        if constexpr (P::N16 >= 1) fft_pass<N, N1, 16, T, false>(s, tid);
        if constexpr (P::N16 >= 2) fft_pass<N, N1 / 16, 16, T, false>(s, tid);
        if constexpr (P::N16 >= 3) fft_pass<N, N1 / 256, 16, T, false>(s, tid);
    } else {
        if constexpr (P::N16 >= 3) fft_pass<N, N1 / 256, 16, T, true>(s, tid);
        if constexpr (P::N16 >= 2) fft_pass<N, N1 / 16, 16, T, true>(s, tid);
        if constexpr (P::N16 >= 1) fft_pass<N, N1, 16, T, true>(s, tid);
        if constexpr (P::R0 > 1) fft_pass<N, N, P::R0, T, true>(s, tid);
    }
}

// position (index in the forward transform's output order) -> frequency bin
template <int N>
__host__ __device__ __forceinline__ int fft_pos_to_freq(int p) {
    using P = FftPlan<N>;
    int k = 0, mul = 1, len = N;
    if (P::R0 > 1) {
        len = N / P::R0;
        k += (p / len) * mul;
        p %= len;
        mul *= P::R0;
    }
#pragma unroll
    for (int i = 0; i < P::N16; ++i) {
        len /= 16;
        k += (p / len) * mul;
        p %= len;
        mul *= 16;
    }
    return k;
}

// Device table of W_N^j = exp(-2*pi*i*j/N), j = 0..N-1 (lazily built per device, float64-exact, cached).
