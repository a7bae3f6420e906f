// Development probe (not product code): cycles per tcgen05.mma (kind::f16, M=128, K=16) as a function of N,
// of the shared-memory layout of the A operand, and of whether consecutive MMAs hit the same accumulator.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_probe tools/mma_probe.cu && ./mma_probe
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tWAIT_LOOP:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra DONE;\n\tbra WAIT_LOOP;\n\tDONE:\n\t}"
        ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_mma(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void tc_mma_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
                 ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
// whole warp executes; one elected lane issues (no C++-level branch -> operands can live in uniform registers)
__device__ __forceinline__ void tc_mma_elect(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred pe, pa;\n\telect.sync _|pe, 0xffffffff;\n\tsetp.ne.b32 pa, %4, 0;\n\t"
                 "@pe tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, pa;\n\t}"
                 ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo, uint32_t layout) {
    return (uint64_t)((saddr >> 4) & 0x3fff) | ((uint64_t)((lbo >> 4) & 0x3fff) << 16) |
           ((uint64_t)((sbo >> 4) & 0x3fff) << 32) | (1ull << 46) | ((uint64_t)layout << 61);
}
__device__ __forceinline__ bool elect_one() {
    uint32_t pred = 0;
    asm volatile("{\n\t.reg .pred px;\n\telect.sync _|px, 0xffffffff;\n\tselp.u32 %0, 1, 0, px;\n\t}" : "=r"(pred));
    return pred != 0;
}
__host__ __device__ constexpr uint32_t make_idesc(int m, int n) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

// mode: 0 = A no-swizzle, SBO 128   1 = A no-swizzle, SBO 160 (halo rows)   2 = A 128B swizzle   3 = A from TMEM
// nwarps issuing warps, each with its own accumulator and its own completion barrier.
__global__ void __launch_bounds__(128, 1) probe(int m, int n, int mode, int nwarps, int iters, long long* out) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar[4];
    __shared__ uint32_t tmem_slot;
    for (int i = threadIdx.x; i < 96 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
    if (threadIdx.x == 0) { for (int i = 0; i < 4; ++i) mbar_init(&bar[i], 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = __shfl_sync(0xffffffffu, tmem_slot, 0);
    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
    long long t0 = clock64();
    if (warp < nwarps) {
        const uint32_t a_base = smem_u32(smem), b_base = a_base + 64 * 1024;
        uint64_t ad;
        if (mode == 0) ad = make_desc(a_base, 128 * (m / 8), 128, 0);
        else if (mode == 1) ad = make_desc(a_base, 17280, 160, 0);
        else ad = make_desc(a_base, 16, 1024, 2);
        const uint64_t bd = make_desc(b_base, 128 * (n / 8), 128, 0);
        const uint32_t idesc = make_idesc(m, n);
        const uint32_t d = tmem + (uint32_t)(warp * n) % 256u;
        for (int i = 0; i < iters; i += 4) {
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                if (mode == 3) { if (elect_one()) tc_mma_ts(d, tmem + 480, bd, idesc, 1); }
                else tc_mma_elect(d, ad + (uint64_t)(2 * u), bd, idesc, 1);
            }
        }
        if (elect_one()) tc_commit(&bar[warp]);
        __syncwarp();
        mbar_wait(&bar[warp], 0);
    }
    __syncthreads();
    long long t1 = clock64();
    if (blockIdx.x == 0 && threadIdx.x == 0) out[0] = t1 - t0;
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
}

int main() {
    long long* out;
    cudaMalloc(&out, 8);
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
    const int iters = 4096;
    const char* names[4] = {"SS nosw sbo128", "SS nosw sbo160", "SS sw128", "TS (A in TMEM)"};
    for (int mode = 0; mode < 4; ++mode)
        for (int n = 16; n <= 256; n *= 2)
            for (int nwarps = 1; nwarps <= 4; nwarps *= 2) {
                probe<<<148, 128, 96 * 1024>>>(128, n, mode, nwarps, iters, out);
                cudaError_t e = cudaDeviceSynchronize();
                long long cyc = 0;
                cudaMemcpy(&cyc, out, 8, cudaMemcpyDeviceToHost);
                printf("M=128 N=%3d %-16s issuing warps=%d : %7.1f cycles/MMA %s\n", n, names[mode], nwarps,
                       (double)cyc / iters / nwarps, e == cudaSuccess ? "" : cudaGetErrorString(e));
                if (e != cudaSuccess) return 1;
            }
    return 0;
}
