// umma_probe.cu — stand-alone probe of the three hardware contracts the tensor-core K1 (k1_mma.cu) relies on:
//   (1) tcgen05.mma kind::tf32 with A in TMEM (lane = row, column = k) and B in shared memory, K-major, no swizzle,
//       described by a hand-built shared-memory descriptor + instruction descriptor;
//   (2) what the tensor core does with the 13 low mantissa bits of an fp32 operand (truncate or round);
//   (3) a 2-D TMA box load with SWIZZLE_128B from a [rows][1000] float view (row pitch 4000 B) incl. out-of-bounds zero fill.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -o scratch/umma_probe tools/microbench/umma_probe.cu
#include <cuda.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count)); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile("{\n.reg .pred p;\nW: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D;\nbra W;\nD:\n}\n" ::"r"(bar), "r"(parity) : "memory");
}

// N columns, one K = 8 step: B image in shared memory = [2 chunks][N/8 groups][8 rows][4 floats]
__global__ void __launch_bounds__(128) mma_probe(const float *A, const float *Bimg, int N, int n_steps, float *D) {
    __shared__ __align__(1024) float sB[2 * 256 * 8];
    __shared__ __align__(8) unsigned long long bar;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int i = tid; i < n_steps * N * 8; i += 128) sB[i] = Bimg[i];
    if (tid == 0) { mbar_init(smem_u32(&bar), 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_base_s)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t tmem = tmem_base_s;
    const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
    const uint32_t colA = 0, colD = 256;
    // A rows -> TMEM (n_steps * 8 columns per row), D zeroed
    for (int s = 0; s < n_steps; ++s) {
        uint32_t a[8];
        for (int k = 0; k < 8; ++k) a[k] = __float_as_uint(A[(size_t)tid * n_steps * 8 + s * 8 + k]);
        asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(tmem + lane_base + colA + s * 8),
                     "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(a[4]), "r"(a[5]), "r"(a[6]), "r"(a[7]));
    }
    for (int c = 0; c < N; c += 8) {
        const uint32_t z = 0;
        asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%1,%1,%1,%1,%1,%1,%1};" ::"r"(tmem + lane_base + colD + c), "r"(z));
    }
    asm volatile("tcgen05.wait::st.sync.aligned;");
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (tid == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;");
        // instruction descriptor: D f32 (1<<4), A tf32 (2<<7), B tf32 (2<<10), K-major both, N>>3 at bit 17, M>>4 at bit 24
        const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24);
        for (int s = 0; s < n_steps; ++s) {
            const uint32_t b_addr = smem_u32(sB) + s * N * 32;
            // shared-memory descriptor: start>>4 | LBO>>4 at bit 16 | SBO>>4 at bit 32 | version 1 at bit 46 | swizzle none
            const uint64_t bdesc = (uint64_t)((b_addr >> 4) & 0x3FFF) | ((uint64_t)((N * 16) >> 4) << 16) | ((uint64_t)(128 >> 4) << 32) | (1ull << 46);
            const uint32_t acc = 1;
            asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
                         "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, {%5, %5, %5, %5}, p;\n}\n" ::"r"(tmem + colD),
                         "r"(tmem + colA + s * 8), "l"(bdesc), "r"(idesc), "r"(acc), "r"(0u));
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    }
    mbar_wait(smem_u32(&bar), 0);
    asm volatile("tcgen05.fence::after_thread_sync;");
    for (int c = 0; c < N; c += 8) {
        uint32_t v[8];
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];" : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]),
                     "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]) : "r"(tmem + lane_base + colD + c));
        asm volatile("tcgen05.wait::ld.sync.aligned;");
        for (int k = 0; k < 8; ++k) D[(size_t)tid * N + c + k] = __uint_as_float(v[k]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem));
}

// latency of tcgen05.commit -> mbarrier completion as seen by the issuing thread, after n_mma MMAs (N columns each, K = 8);
// dep != 0: all MMAs accumulate into the same D columns, else into disjoint column ranges
__global__ void __launch_bounds__(128) commit_latency(int n_mma, int N, int dep, int reps, long long *out) {
    __shared__ __align__(1024) float sB[256 * 8];
    __shared__ __align__(8) unsigned long long bar;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < 256 * 8; i += 128) sB[i] = 1.0f;
    if (tid == 0) { mbar_init(smem_u32(&bar), 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_base_s)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t tmem = tmem_base_s;
    if (tid == 0) {
        const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24);
        const uint64_t bdesc = (uint64_t)((smem_u32(sB) >> 4) & 0x3FFF) | ((uint64_t)((N * 16) >> 4) << 16) | ((uint64_t)(128 >> 4) << 32) | (1ull << 46);
        long long tot = 0, tot_issue = 0;
        for (int r = 0; r < reps; ++r) {
            const long long t0 = clock64();
            for (int m = 0; m < n_mma; ++m) {
                const uint32_t d = tmem + 256 + (dep ? 0 : (m * N) % 256);
                asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
                             "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, {%5, %5, %5, %5}, p;\n}\n" ::"r"(d),
                             "r"(tmem + (m % 4) * 8), "l"(bdesc), "r"(idesc), "r"(1u), "r"(0u));
            }
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
            const long long t1 = clock64();
            mbar_wait(smem_u32(&bar), r & 1);
            const long long t2 = clock64();
            tot += t2 - t0; tot_issue += t1 - t0;
        }
        out[0] = tot / reps; out[1] = tot_issue / reps;
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem));
}

// issue rate of small MMAs from 1, 2 or 4 threads (one per warp) at once, each into its own D columns; one commit per thread
__global__ void __launch_bounds__(128) issue_rate(int n_mma, int N, int n_thr, long long *out) {
    __shared__ __align__(1024) float sB[256 * 8];
    __shared__ __align__(8) unsigned long long bar[4];
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int i = tid; i < 256 * 8; i += 128) sB[i] = 1.0f;
    if (tid == 0) { for (int i = 0; i < 4; ++i) mbar_init(smem_u32(&bar[i]), 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_base_s)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t tmem = tmem_base_s;
    if (lane == 0 && warp < n_thr) {
        const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24);
        const uint64_t bdesc = (uint64_t)((smem_u32(sB) >> 4) & 0x3FFF) | ((uint64_t)((N * 16) >> 4) << 16) | ((uint64_t)(128 >> 4) << 32) | (1ull << 46);
        const uint32_t d = tmem + 128 + warp * 64, a0 = tmem + warp * 32;
        const long long t0 = clock64();
#pragma unroll 8
        for (int m = 0; m < n_mma; ++m) {
            asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
                         "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, {%5, %5, %5, %5}, p;\n}\n" ::"r"(d),
                         "r"(a0 + (m & 3) * 8), "l"(bdesc), "r"(idesc), "r"(1u), "r"(0u));
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar[warp])) : "memory");
        const long long t1 = clock64();
        mbar_wait(smem_u32(&bar[warp]), 0);
        const long long t2 = clock64();
        out[2 * warp] = t2 - t0; out[2 * warp + 1] = t1 - t0;
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem));
}

__global__ void __launch_bounds__(128) tma_probe(const __grid_constant__ CUtensorMap tmap, int c0, int c1, float *out) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    float *tile = (float *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    __shared__ __align__(8) unsigned long long bar;
    if (threadIdx.x == 0) {
        mbar_init(smem_u32(&bar), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)), "r"(128 * 128) : "memory");
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(smem_u32(tile)),
                     "l"(&tmap), "r"(c0), "r"(c1), "r"(smem_u32(&bar)) : "memory");
    }
    __syncthreads();
    mbar_wait(smem_u32(&bar), 0);
    for (int i = threadIdx.x; i < 128 * 32; i += 128) out[i] = tile[i];
}

static float tf32_trunc(float x) { uint32_t u; memcpy(&u, &x, 4); u &= 0xFFFFE000u; memcpy(&x, &u, 4); return x; }
static float tf32_rn(float x) { uint32_t u; memcpy(&u, &x, 4); u += 0x00000FFFu + ((u >> 13) & 1u); u &= 0xFFFFE000u; memcpy(&x, &u, 4); return x; }

typedef CUresult (*EncodeFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                             const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main(int argc, char **argv) {
    const int dim0x = argc > 1 ? atoi(argv[1]) : 1032, c0x = argc > 2 ? atoi(argv[2]) : 2;
    srand(1);
    for (int N : {32, 48, 16, 64}) {
        for (int mode = 0; mode < 2; ++mode) {            // 0: B exactly representable in tf32; 1: B with low bits too
            const int n_steps = 3;
            std::vector<float> A(128 * 8 * n_steps), B(n_steps * N * 8), Bimg(n_steps * N * 8), D(128 * N);
            for (auto &v : A) v = (float)rand() / RAND_MAX * 2 - 1;
            for (auto &v : B) { v = (float)rand() / RAND_MAX * 2 - 1; if (mode == 0) v = tf32_rn(v); }
            // logical B[s][n][k] (k < 8) -> image [s][chunk = k/4][n/8][n%8][k%4]
            for (int s = 0; s < n_steps; ++s)
                for (int n = 0; n < N; ++n)
                    for (int k = 0; k < 8; ++k)
                        Bimg[(size_t)s * N * 8 + (k / 4) * (N * 4) + (n / 8) * 32 + (n % 8) * 4 + (k % 4)] = B[((size_t)s * N + n) * 8 + k];
            float *dA, *dB, *dD;
            CK(cudaMalloc(&dA, A.size() * 4)); CK(cudaMalloc(&dB, Bimg.size() * 4)); CK(cudaMalloc(&dD, D.size() * 4));
            CK(cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice));
            CK(cudaMemcpy(dB, Bimg.data(), Bimg.size() * 4, cudaMemcpyHostToDevice));
            mma_probe<<<1, 128>>>(dA, dB, N, n_steps, dD);
            CK(cudaDeviceSynchronize());
            CK(cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost));
            double e_tt = 0, e_rr = 0, e_tr = 0, e_rt = 0, e_full = 0;    // hypotheses: (A,B) truncated/rounded
            for (int m = 0; m < 128; ++m)
                for (int n = 0; n < N; ++n) {
                    double tt = 0, rr = 0, tr = 0, rt = 0, full = 0;
                    for (int s = 0; s < n_steps; ++s)
                        for (int k = 0; k < 8; ++k) {
                            const float a = A[(size_t)m * n_steps * 8 + s * 8 + k], b = B[((size_t)s * N + n) * 8 + k];
                            tt += (double)tf32_trunc(a) * tf32_trunc(b); rr += (double)tf32_rn(a) * tf32_rn(b);
                            tr += (double)tf32_trunc(a) * tf32_rn(b); rt += (double)tf32_rn(a) * tf32_trunc(b); full += (double)a * b;
                        }
                    const double d = D[(size_t)m * N + n];
                    e_tt = fmax(e_tt, fabs(d - tt)); e_rr = fmax(e_rr, fabs(d - rr)); e_tr = fmax(e_tr, fabs(d - tr));
                    e_rt = fmax(e_rt, fabs(d - rt)); e_full = fmax(e_full, fabs(d - full));
                }
            printf("mma N=%d mode=%d: max|D - ref| with (A,B) = (trunc,trunc) %.3e  (rn,rn) %.3e  (trunc,rn) %.3e  (rn,trunc) %.3e  exact %.3e   D[0..3]=%g %g %g %g\n",
                   N, mode, e_tt, e_rr, e_tr, e_rt, e_full, D[0], D[1], D[2], D[3]);
            cudaFree(dA); cudaFree(dB); cudaFree(dD);
        }
    }
    {
        long long *dl; CK(cudaMalloc(&dl, 16));
        for (int dep = 0; dep < 2; ++dep)
            for (int N : {16, 32, 64, 128, 256})
                for (int n_mma : {0, 1, 2, 8, 32}) {
                    commit_latency<<<1, 128>>>(n_mma, N, dep, 200, dl);
                    CK(cudaDeviceSynchronize());
                    long long h[2]; CK(cudaMemcpy(h, dl, 16, cudaMemcpyDeviceToHost));
                    printf("commit latency: %2d MMAs N=%3d %s: issue->arrival %lld clk (issue loop %lld clk)\n", n_mma, N, dep ? "dependent  " : "independent", h[0], h[1]);
                }
    }
    {
        long long *dl; CK(cudaMalloc(&dl, 64));
        for (int N : {32, 64})
            for (int n_thr : {1, 2, 4}) {
                issue_rate<<<1, 128>>>(256, N, n_thr, dl);
                CK(cudaDeviceSynchronize());
                long long h[8]; CK(cudaMemcpy(h, dl, 64, cudaMemcpyDeviceToHost));
                printf("issue rate: %d thread(s) x 256 MMAs N=%d:", n_thr, N);
                for (int t = 0; t < n_thr; ++t) printf("  thr%d done %lld clk (issued %lld) = %.1f clk/MMA", t, h[2 * t], h[2 * t + 1], h[2 * t] / 256.0);
                printf("\n");
            }
    }
    // ---- TMA probe ----
    EncodeFn encode = nullptr;
    cudaDriverEntryPointQueryResult qres;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void **)&encode, cudaEnableDefault, &qres));
    if (!encode) { printf("no cuTensorMapEncodeTiled\n"); return 1; }
    const int rows = 300, rowlen = 1000;
    std::vector<float> X((size_t)rows * rowlen + 64);
    for (size_t i = 0; i < X.size(); ++i) X[i] = (float)i;
    float *dX, *dOut;
    CK(cudaMalloc(&dX, X.size() * 4)); CK(cudaMalloc(&dOut, 128 * 32 * 4));
    CK(cudaMemcpy(dX, X.data(), X.size() * 4, cudaMemcpyHostToDevice));
    for (int off = 0; off < 2; ++off) {                   // base pointer offset: 0 or 16 bytes
        CUtensorMap tmap;
        cuuint64_t dims[2] = {(cuuint64_t)rowlen, (cuuint64_t)rows};
        cuuint64_t strides[1] = {(cuuint64_t)rowlen * 4};
        cuuint32_t box[2] = {32, 128};
        cuuint32_t estr[2] = {1, 1};
        CUresult r = encode(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, dX + 4 * off, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        printf("encode (base + %d B): %d\n", 16 * off, (int)r);
        if (r != CUDA_SUCCESS) continue;
        for (int c0 : {0, 64, 992}) {
            const int c1 = 200;                           // rows 200..327: rows >= 300 are out of bounds
            tma_probe<<<1, 128, 128 * 128 + 1024>>>(tmap, c0, c1, dOut);
            CK(cudaDeviceSynchronize());
            std::vector<float> out(128 * 32);
            CK(cudaMemcpy(out.data(), dOut, out.size() * 4, cudaMemcpyDeviceToHost));
            int bad = 0;
            for (int rrow = 0; rrow < 128; ++rrow)
                for (int f = 0; f < 32; ++f) {
                    const int phys = rrow * 32 + (((f / 4) ^ (rrow & 7)) * 4) + (f % 4);
                    const bool inb = (c1 + rrow) < rows && (c0 + f) < rowlen;
                    const float want = inb ? (float)((size_t)(c1 + rrow) * rowlen + c0 + f + 4 * off) : 0.f;
                    if (out[phys] != want) { if (bad < 4) printf("  mismatch row %d f %d: got %g want %g\n", rrow, f, out[phys], want); ++bad; }
                }
            printf("tma c0=%d: %d mismatches against the SWIZZLE_128B hypothesis (chunk j of row r at j ^ (r & 7)), OOB -> 0\n", c0, bad);
        }
    }
    // ---- overlapping rows: dim0 * 4 B > row pitch, box start at an 8-byte (not 16-byte) aligned element ----
    {
        CUtensorMap tmap;
        cuuint64_t dims[2] = {(cuuint64_t)dim0x, (cuuint64_t)(rows - 2)};
        cuuint64_t strides[1] = {(cuuint64_t)rowlen * 4};
        cuuint32_t box[2] = {32, 128};
        cuuint32_t estr[2] = {1, 1};
        CUresult r = encode(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, dX, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        printf("encode rows (dim0 %d floats, pitch 1000): %d\n", dim0x, (int)r);
        if (r == CUDA_SUCCESS) {
            for (int c0 : {c0x}) {
                const int c1 = 100;
                tma_probe<<<1, 128, 128 * 128 + 1024>>>(tmap, c0, c1, dOut);
                CK(cudaDeviceSynchronize());
                std::vector<float> out(128 * 32);
                CK(cudaMemcpy(out.data(), dOut, out.size() * 4, cudaMemcpyDeviceToHost));
                int bad = 0;
                for (int rrow = 0; rrow < 128; ++rrow)
                    for (int f = 0; f < 32; ++f) {
                        const int phys = rrow * 32 + (((f / 4) ^ (rrow & 7)) * 4) + (f % 4);
                        const float want = (c0 + f) < dim0x ? (float)((size_t)(c1 + rrow) * rowlen + c0 + f) : 0.f;
                        if (out[phys] != want) { if (bad < 4) printf("  mismatch row %d f %d: got %g want %g\n", rrow, f, out[phys], want); ++bad; }
                    }
                printf("tma overlapping c0=%d: %d mismatches\n", c0, bad);
            }
        }
    }
    return 0;
}
