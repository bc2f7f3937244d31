// umma_rate.cu — how many clocks does one tcgen05.mma take, by operand form?  (sm_100a; timing only: operands are whatever
// lies in shared / tensor memory.)  One CTA, one issuing thread: t0, `reps` back-to-back MMAs into one accumulator, tcgen05.commit,
// mbarrier wait, t1.  Variants: A from tensor memory (TS) or shared memory (SS); B K-major with no swizzle (canonical 8 x 16 B core
// matrices, what k1_mma.cu / k1_chan.cu use) or SWIZZLE_128B; kind::tf32 (K = 8) or kind::f16 with bf16 operands (K = 16).
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o scratch/umma_rate tools/microbench/umma_rate.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

struct Variant { int ts, swz, bf16, N, two_acc; };

__global__ void __launch_bounds__(128) rate_kernel(Variant v, int reps, long long *out) {
    extern __shared__ __align__(1024) unsigned char raw[];
    unsigned char *sm = (unsigned char *)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
    __shared__ __align__(8) unsigned long long bar;
    __shared__ uint32_t tmem_s;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < 48 * 1024 / 4; i += 128) ((float *)sm)[i] = 1e-3f * (float)(i & 63);
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_s)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t tmem = tmem_s;
    // zero the A columns (0..15) and the accumulators (128..)
    {
        const uint32_t lane_base = (uint32_t)(warp * 32) << 16, z = 0;
        for (int c = 0; c < 512; c += 8)
            asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%1,%1,%1,%1,%1,%1,%1};" ::"r"(tmem + lane_base + c), "r"(z));
        asm volatile("tcgen05.wait::st.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (warp == 0) {          // the whole warp runs the loop, one elected lane issues: operands stay in uniform registers
        asm volatile("tcgen05.fence::after_thread_sync;");
        const int N = v.N;
        const uint32_t fmt = v.bf16 ? 1u : 2u;                                // kind::f16: 1 = bf16; kind::tf32: 2 = tf32
        const uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24);
        unsigned long long bdesc, adesc;                                      // B at sm + 16 KB, A (SS form) at sm
        if (v.swz) {          // SWIZZLE_128B, K-major: 8-row groups 1024 B apart
            bdesc = (unsigned long long)((smem_u32(sm + 16384) >> 4) & 0x3FFF) | (1ull << 16) | ((unsigned long long)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
            adesc = (unsigned long long)((smem_u32(sm) >> 4) & 0x3FFF) | (1ull << 16) | ((unsigned long long)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
        } else {              // no swizzle: the two 16-byte K chunks N*16 (M*16) bytes apart, 8-row groups 128 B apart
            bdesc = (unsigned long long)((smem_u32(sm + 16384) >> 4) & 0x3FFF) | ((unsigned long long)((N * 16) >> 4) << 16) | ((unsigned long long)(128 >> 4) << 32) | (1ull << 46);
            adesc = (unsigned long long)((smem_u32(sm) >> 4) & 0x3FFF) | ((unsigned long long)((128 * 16) >> 4) << 16) | ((unsigned long long)(128 >> 4) << 32) | (1ull << 46);
        }
        const uint32_t d0 = tmem + 128, d1 = v.two_acc ? tmem + 128 + 192 : d0, a_t = tmem;
        const long long t0 = clock64();
        for (int r = 0; r < reps; r += 8) {
            unsigned el;
            asm volatile("{\n.reg .pred p;\nelect.sync _|p, 0xffffffff;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(el));
            if (el) {
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    const uint32_t d = (q & 1) ? d1 : d0;
                    if (v.ts) {
                        if (v.bf16)
                            asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, {%5,%5,%5,%5}, p;\n}\n" ::"r"(d), "r"(a_t), "l"(bdesc), "r"(idesc), "r"(1u), "r"(0u) : "memory");
                        else
                            asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, {%5,%5,%5,%5}, p;\n}\n" ::"r"(d), "r"(a_t), "l"(bdesc), "r"(idesc), "r"(1u), "r"(0u) : "memory");
                    } else {
                        if (v.bf16)
                            asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, {%5,%5,%5,%5}, p;\n}\n" ::"r"(d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(1u), "r"(0u) : "memory");
                        else
                            asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, {%5,%5,%5,%5}, p;\n}\n" ::"r"(d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(1u), "r"(0u) : "memory");
                    }
                }
            }
            __syncwarp();
        }
        if (tid == 0) asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
        __syncwarp();
        asm volatile("{\n.reg .pred p;\nW: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n@p bra D;\nbra W;\nD:\n}\n" ::"r"(smem_u32(&bar)) : "memory");
        const long long t1 = clock64();
        if (tid == 0) out[0] = t1 - t0;
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem));
}

int main() {
    long long *d_out, h_out;
    CK(cudaMalloc(&d_out, 8));
    CK(cudaFuncSetAttribute(rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
    const int reps = 4000;
    const Variant vs[] = {{1, 0, 0, 128, 0}, {1, 0, 0, 192, 0}, {1, 0, 0, 64, 0}, {1, 0, 0, 32, 0}, {1, 0, 0, 128, 1}, {1, 1, 0, 128, 0}, {1, 1, 0, 192, 0},
                          {0, 0, 0, 128, 0}, {0, 1, 0, 128, 0}, {0, 1, 0, 256, 0}, {1, 0, 1, 128, 0}, {1, 1, 1, 128, 0}, {0, 1, 1, 128, 0}, {0, 1, 1, 256, 0}};
    for (const Variant &v : vs) {
        for (int rep = 0; rep < 2; ++rep) {
            rate_kernel<<<1, 128, 64 * 1024>>>(v, reps, d_out);
            CK(cudaDeviceSynchronize());
        }
        CK(cudaMemcpy(&h_out, d_out, 8, cudaMemcpyDeviceToHost));
        const double clk = (double)h_out / reps;
        const double macs = 128.0 * v.N * (v.bf16 ? 16 : 8);
        printf("A %s  B %-10s %s  N %3d  %s : %7.1f clocks per MMA = %6.0f MAC/clk\n", v.ts ? "TMEM" : "smem", v.swz ? "swizzle128" : "no-swizzle",
               v.bf16 ? "bf16 K16" : "tf32 K8 ", v.N, v.two_acc ? "2 accumulators" : "1 accumulator ", clk, macs / clk);
    }
    return 0;
}
