// Development probe (not product code) for the "kd-in-N" convolution planned in DESIGN.md section 10: three hardware
// questions that decide its TMEM layout, answered by exact integer-valued MMAs (M = 128, K = 16, bf16 -> fp32):
//   Q1  may an MMA's D operand start at a TMEM column that is a multiple of 8 but not of 16 (N = 32)?
//   Q2  do accumulating MMAs issued by DIFFERENT warps into the SAME accumulator columns add up without lost updates?
//   Q3  does tcgen05.st zero an accumulator so that accumulate=1 MMAs can start from it?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o kdn_probe tools/kdn_probe.cu && ./kdn_probe
#include <cuda_bf16.h>
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
__device__ __forceinline__ void tc_commit_elect(uint64_t* bar) {
    asm volatile("{\n\t.reg .pred pe;\n\telect.sync _|pe, 0xffffffff;\n\t"
                 "@pe tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_mma_elect(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred pe, pa;\n\telect.sync _|pe, 0xffffffff;\n\tsetp.ne.b32 pa, %4, 0;\n\t"
                 "@pe tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, pa;\n\t}"
                 ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
    return (uint64_t)((saddr >> 4) & 0x3fff) | ((uint64_t)((lbo >> 4) & 0x3fff) << 16) |
           ((uint64_t)((sbo >> 4) & 0x3fff) << 32) | (1ull << 46);
}
__host__ __device__ constexpr uint32_t make_idesc(int m, int n) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&v)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]) : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_st8_zero(uint32_t taddr) {
    const uint32_t z = 0u;
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %1, %1, %1, %1, %1, %1, %1};" ::"r"(taddr), "r"(z) : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// integer-valued operands: A[r][k] = (r % 5) - 2 + (k % 3), B[n][k] = (n % 7) - 3 + (k % 2)   (exact in bf16 / fp32)
__host__ __device__ inline float a_val(int r, int k) { return (float)((r % 5) - 2 + (k % 3)); }
__host__ __device__ inline float b_val(int n, int k) { return (float)((n % 7) - 3 + (k % 2)); }

constexpr int M = 128, N = 32, K = 16, NCOLS = 128;

// test 1: zero columns [0, NCOLS) with tcgen05.st, ONE accumulating MMA at column offset `col0` (multiple of 8)
// test 2: zero, then 4 warps x `reps` accumulating MMAs into the SAME columns (col0)
__global__ void __launch_bounds__(128, 1) probe(int test, int col0, int reps, float* out /* [M][NCOLS] */) {
    __shared__ __align__(1024) uint8_t sa[M * K * 2];
    __shared__ __align__(1024) uint8_t sb[N * K * 2];
    __shared__ uint64_t bar[4];
    __shared__ uint32_t tmem_slot;
    // K-major, no swizzle: [kc 2][8-row group][8 rows][8 elements]
    for (int i = threadIdx.x; i < M * K; i += 128) {
        const int r = i / K, k = i % K;
        reinterpret_cast<__nv_bfloat16*>(sa)[(k / 8) * (M * 8) + (r / 8) * 64 + (r % 8) * 8 + (k % 8)] = __float2bfloat16(a_val(r, k));
    }
    for (int i = threadIdx.x; i < N * K; i += 128) {
        const int n = i / K, k = i % K;
        reinterpret_cast<__nv_bfloat16*>(sb)[(k / 8) * (N * 8) + (n / 8) * 64 + (n % 8) * 8 + (k % 8)] = __float2bfloat16(b_val(n, k));
    }
    if (threadIdx.x == 0) { for (int i = 0; i < 4; ++i) mbar_init(&bar[i], 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(NCOLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_slot;
    const uint32_t lane_base = tmem + ((uint32_t)(warp * 32) << 16);
    // Q3: zero every column of this warp's lane quadrant
    for (int c = 0; c < NCOLS; c += 8) tmem_st8_zero(lane_base + c);
    tmem_st_wait();
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint64_t ad = make_desc(smem_u32(sa), M * 16, 128);       // LBO = K-chunk stride, SBO = 8-row group stride
    const uint64_t bd = make_desc(smem_u32(sb), N * 16, 128);
    const uint32_t idesc = make_idesc(M, N);
    const int issuers = test == 1 ? 1 : 4;
    if (warp < issuers) {
        for (int i = 0; i < (test == 1 ? 1 : reps); ++i) tc_mma_elect(tmem + col0, ad, bd, idesc, 1u);      // always accumulate
        tc_commit_elect(&bar[warp]);
    }
    __syncwarp();
    for (int w = 0; w < issuers; ++w) mbar_wait(&bar[w], 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    for (int c = 0; c < NCOLS; c += 8) {
        uint32_t v[8];
        tmem_ld8(lane_base + c, v);
        for (int k = 0; k < 8; ++k) out[(warp * 32 + lane) * NCOLS + c + k] = __uint_as_float(v[k]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(NCOLS) : "memory");
}

static int check(const float* h, int col0, float scale, const char* what) {
    int bad = 0;
    for (int r = 0; r < M; ++r)
        for (int c = 0; c < NCOLS; ++c) {
            float want = 0.f;
            if (c >= col0 && c < col0 + N) {
                for (int k = 0; k < K; ++k) want += a_val(r, k) * b_val(c - col0, k);
                want *= scale;
            }
            if (h[r * NCOLS + c] != want) {
                if (bad < 4) printf("  %s: row %d col %d got %g want %g\n", what, r, c, h[r * NCOLS + c], want);
                ++bad;
            }
        }
    printf("%-64s %s (%d mismatches)\n", what, bad ? "FAIL" : "ok", bad);
    return bad;
}

int main() {
    float* d;
    cudaMalloc(&d, sizeof(float) * M * NCOLS);
    float* h = (float*)malloc(sizeof(float) * M * NCOLS);
    int fails = 0;
    const int cols[4] = {0, 16, 8, 40};
    for (int i = 0; i < 4; ++i) {
        probe<<<1, 128>>>(1, cols[i], 1, d);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("test1 col0=%d: %s\n", cols[i], cudaGetErrorString(e)); return 1; }
        cudaMemcpy(h, d, sizeof(float) * M * NCOLS, cudaMemcpyDeviceToHost);
        char what[96];
        snprintf(what, sizeof(what), "Q1/Q3 zeroed TMEM + one accumulating MMA, N=32 at column %d", cols[i]);
        fails += check(h, cols[i], 1.f, what) != 0;
    }
    for (int reps = 1; reps <= 256; reps *= 16) {
        probe<<<1, 128>>>(2, 8, reps, d);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("test2 reps=%d: %s\n", reps, cudaGetErrorString(e)); return 1; }
        cudaMemcpy(h, d, sizeof(float) * M * NCOLS, cudaMemcpyDeviceToHost);
        char what[96];
        snprintf(what, sizeof(what), "Q2 four warps x %d accumulating MMAs into the same columns", reps);
        fails += check(h, 8, 4.f * reps, what) != 0;
    }
    printf(fails ? "SOME PROBES FAILED\n" : "all probes ok\n");
    return fails ? 2 : 0;
}
