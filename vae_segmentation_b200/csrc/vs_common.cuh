// Shared helpers for the vaeseg_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include "../../include/vaeseg_b200.h"

typedef __nv_bfloat16 bf16;

// ---- error plumbing ---------------------------------------------------------------------
void vs_set_error(const char* fmt, ...);
#define VS_FAIL(code, ...) do { vs_set_error(__VA_ARGS__); return (code); } while (0)
#define VS_CHECK_LAUNCH(name) do { cudaError_t e_ = cudaGetLastError(); \
    if (e_ != cudaSuccess) VS_FAIL(VS_ERR_CUDA, "%s: %s", name, cudaGetErrorString(e_)); } while (0)
#define VS_CUDA(call, name) do { cudaError_t e_ = (call); \
    if (e_ != cudaSuccess) VS_FAIL(VS_ERR_CUDA, "%s: %s", name, cudaGetErrorString(e_)); } while (0)
#define VS_REQUIRE(cond, code, ...) do { if (!(cond)) VS_FAIL(code, __VA_ARGS__); } while (0)

// ---- programmatic dependent launch (PDL) ------------------------------------------------------
// The step is a chain of ~500 short dependent kernels.  Kernels launched through vs_launch() carry
// cudaLaunchAttributeProgrammaticStreamSerialization (when enabled, vs_set_pdl): the next kernel of the stream may be
// scheduled as soon as every CTA of this one has executed pdl_trigger() -- its launch latency, barrier init, TMEM
// allocation and shared-memory carve-up then overlap this kernel's tail -- and it blocks in pdl_wait() until this
// kernel has completed and flushed its memory.  RULE: a kernel launched through vs_launch() executes pdl_wait() in
// every thread before its first global-memory access (reads of activations AND writes: the predecessor may still be
// reading what this kernel overwrites).  Without the attribute both instructions are no-ops.
extern int g_vs_pdl;
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
template <typename... KArgs, typename... Args>
static inline cudaError_t vs_launch(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    // g_vs_pdl = largest grid (CTAs) launched programmatically: a dependent that is scheduled early holds its SM slots
    // while it waits, which on the big persistent grids starves the concurrent streams (teacher forward, weight
    // gradients) -- measured 334 vs 384 vol/s with every launch programmatic -- so only the small latency-bound
    // launches of the deep levels use it
    cfg.numAttrs = (g_vs_pdl > 0 && (long long)grid.x * grid.y * grid.z <= (long long)g_vs_pdl) ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

// Cooperative launch: every CTA of the grid is co-resident (the driver gang-schedules the grid, also against kernels of
// other streams and inside captured graphs -- tools/coop_probe.cu), which is what makes an in-kernel grid barrier
// (vs_grid_barrier) safe.  The grid must fit the device at once (<= SMs x resident CTAs per SM) or the launch fails.
template <typename... KArgs, typename... Args>
static inline cudaError_t vs_launch_coop(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeCooperative;
    attr[0].val.cooperative = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}
// One thread per CTA calls this after the CTA's own threads have fenced and synchronised: arrive on a zero-initialised
// counter and wait for the whole grid.  Only under vs_launch_coop.  Traps instead of hanging on a protocol bug.
__device__ __forceinline__ void vs_grid_barrier(unsigned* counter) {
    __threadfence();
    atomicAdd(counter, 1u);
    unsigned spins = 0;
    while (*reinterpret_cast<volatile unsigned*>(counter) < gridDim.x * gridDim.y * gridDim.z) {
        __nanosleep(64);
        if (++spins > (1u << 23)) __trap();
    }
    __threadfence();
}

static inline bool vs_aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }
int vs_sm_count();
__device__ __forceinline__ bool vs_aligned16_dev(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

// ---- storage type traits ----------------------------------------------------------------
template <typename T> struct Store;
template <> struct Store<float> {
    static __device__ __forceinline__ float ld(const float* p) { return *p; }
    static __device__ __forceinline__ void st(float* p, float v) { *p = v; }
    // 8 consecutive elements (32 B aligned for float)
    static __device__ __forceinline__ void ld8(const float* p, float (&v)[8]) {
        float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
        v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
    }
    static __device__ __forceinline__ void st8(float* p, const float (&v)[8]) {
        *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
        *reinterpret_cast<float4*>(p + 4) = make_float4(v[4], v[5], v[6], v[7]);
    }
    static __device__ __forceinline__ void st2(float* p, float a, float b) { *reinterpret_cast<float2*>(p) = make_float2(a, b); }
    static __device__ __forceinline__ void ld2(const float* p, float& a, float& b) { float2 v = *reinterpret_cast<const float2*>(p); a = v.x; b = v.y; }
};
template <> struct Store<bf16> {
    static __device__ __forceinline__ float ld(const bf16* p) { return __bfloat162float(*p); }
    static __device__ __forceinline__ void st(bf16* p, float v) { *p = __float2bfloat16_rn(v); }
    static __device__ __forceinline__ void ld8(const bf16* p, float (&v)[8]) {
        uint4 r = *reinterpret_cast<const uint4*>(p);
        const uint32_t u[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) { v[2 * i] = __uint_as_float(u[i] << 16); v[2 * i + 1] = __uint_as_float(u[i] & 0xffff0000u); }
    }
    static __device__ __forceinline__ void st8(bf16* p, const float (&v)[8]) {
        uint32_t u[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) { __nv_bfloat162 h = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]); u[i] = *reinterpret_cast<uint32_t*>(&h); }
        *reinterpret_cast<uint4*>(p) = make_uint4(u[0], u[1], u[2], u[3]);
    }
    static __device__ __forceinline__ void st2(bf16* p, float a, float b) { *reinterpret_cast<__nv_bfloat162*>(p) = __floats2bfloat162_rn(a, b); }
    static __device__ __forceinline__ void ld2(const bf16* p, float& a, float& b) { float2 v = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(p)); a = v.x; b = v.y; }
};

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// mean / rstd of InstanceNorm3d (biased variance, eps 1e-5) from (sum, sumsq) over s voxels.
// The sums are fp64: E[x^2]-E[x]^2 cancels catastrophically in fp32 when |mean| >> sigma
// (e.g. the VAE in_block on a mostly-constant mask), and B200 has full-rate-enough FP64.
// _cg: the sums were accumulated by atomics of THIS kernel (other CTAs): read them from L2
__device__ __forceinline__ void in_mean_rstd_cg(const double* st, double inv_s, float& mean, float& rstd) {
    const double s0 = __ldcg(st), s1 = __ldcg(st + 1);
    const double m = s0 * inv_s;
    const double var = fmax(s1 * inv_s - m * m, 0.0);
    mean = (float)m;
    rstd = (float)(1.0 / sqrt(var + 1e-5));
}
__device__ __forceinline__ void in_mean_rstd(const double* st, double inv_s, float& mean, float& rstd) {
    const double m = st[0] * inv_s;
    const double var = fmax(st[1] * inv_s - m * m, 0.0);
    mean = (float)m;
    rstd = (float)(1.0 / sqrt(var + 1e-5));
}

static inline int vs_ceil_div(long long a, long long b) { return (int)((a + b - 1) / b); }

#define VS_DISPATCH_DTYPE(dtype, T, ...) \
    do { if ((dtype) == VS_F32) { typedef float T; __VA_ARGS__; } \
         else if ((dtype) == VS_BF16) { typedef bf16 T; __VA_ARGS__; } \
         else VS_FAIL(VS_ERR_UNSUPPORTED, "unknown dtype %d", (int)(dtype)); } while (0)
