// PTX wrappers shared by the tcgen05 / TMA kernels (sm_100a only): mbarrier, TMA loads, TMEM allocation,
// tcgen05.mma / commit / ld, UMMA shared-memory descriptors and the cuTensorMapEncodeTiled entry point.
#pragma once
#include <cuda.h>
#include "vs_common.cuh"

namespace {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, 0x989680;\n\t"    // suspend-time hint: sleep in hardware, do not poll the shared-memory pipe the MMA operands use
        "@p bra DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "DONE:\n\t"
        "}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void tma_load_5d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2,
                                            int c3, int c4) {
    asm volatile(
        "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
        ::"r"(smem_u32(smem_dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
        : "memory");
}
// Pull a tensor map (a __grid_constant__ kernel parameter) into the TMA unit's descriptor cache ahead of its first use:
// the first cp.async.bulk.tensor of a kernel otherwise pays the descriptor fetch on top of the data latency.
__device__ __forceinline__ void prefetch_tensormap(const CUtensorMap* map) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(map)) : "memory");
}

__device__ __forceinline__ void tma_load_4d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2, int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
        ::"r"(smem_u32(smem_dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}
// TMA stores (shared -> global, bulk async-group completion)
__device__ __forceinline__ void tma_store_5d(const CUtensorMap* map, const void* smem_src, int c0, int c1, int c2, int c3, int c4) {
    asm volatile("cp.async.bulk.tensor.5d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5, %6}], [%1];"
                 ::"l"(map), "r"(smem_u32(smem_src)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4) : "memory");
}
__device__ __forceinline__ void tma_store_4d(const CUtensorMap* map, const void* smem_src, int c0, int c1, int c2, int c3) {
    asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
                 ::"l"(map), "r"(smem_u32(smem_src)), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void bulk_commit_group() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void bulk_wait_group_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void sts128(uint32_t saddr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(saddr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}

__device__ __forceinline__ void bulk_load(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(smem_dst)), "l"(gsrc), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_mma(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
// whole warp executes, one elected lane issues (no C++-level branch around the instruction)
__device__ __forceinline__ void tc_mma_elect(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred pe, pa;\n\t"
        "elect.sync _|pe, 0xffffffff;\n\t"
        "setp.ne.b32 pa, %4, 0;\n\t"
        "@pe tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, pa;\n\t"
        "}" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tc_commit_elect(uint64_t* bar) {
    asm volatile(
        "{\n\t"
        ".reg .pred pe;\n\t"
        "elect.sync _|pe, 0xffffffff;\n\t"
        "@pe tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t"
        "}" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr));
}
__device__ __forceinline__ float4 lds128(uint32_t saddr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(saddr));
    return v;
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// K-major, no-swizzle shared-memory matrix descriptor (cute::UMMA::SmemDescriptor, version 1 = sm_100)
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return (uint64_t)((saddr >> 4) & 0x3fff) | ((uint64_t)((lbo_bytes >> 4) & 0x3fff) << 16) |
           ((uint64_t)((sbo_bytes >> 4) & 0x3fff) << 32) | (1ull << 46);
}
// K-major swizzled descriptor (rows of 32 / 64 / 128 bytes written by TMA with the matching CU_TENSOR_MAP_SWIZZLE_*):
// LBO is unused, SBO = 8 rows, layout type 6 / 4 / 2 (cute::UMMA::LayoutType SWIZZLE_32B / 64B / 128B).  The tile base
// must be aligned to the swizzle repeat (8 rows); advancing K by 16 elements = +32 bytes on the start address.
__device__ __forceinline__ uint64_t make_desc_sw(uint32_t saddr, uint32_t row_bytes) {
    const uint64_t layout = row_bytes == 128 ? 2ull : (row_bytes == 64 ? 4ull : 6ull);
    return (uint64_t)((saddr >> 4) & 0x3fff) | ((uint64_t)(((row_bytes * 8u) >> 4) & 0x3fff) << 32) | (1ull << 46) | (layout << 61);
}
// kind::f16 instruction descriptor: D fp32, A/B bf16, both K-major, M = 128
__host__ __device__ constexpr uint32_t make_idesc(int n) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}


__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&v)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                 : "r"(taddr));
}

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
          "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
          "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr));
}

// Division by a run-time constant as multiply-high + shift (valid for 0 <= x < 2^31): every warp role of the persistent
// kernels decodes its work item, and a hardware-less 32-bit division costs ~30 instructions on a single-warp chain.
struct FastDiv {
    uint32_t mul, shr, d;
};
inline FastDiv make_fastdiv(int d) {
    FastDiv f;
    f.d = (uint32_t)d;
    if (d == 1) { f.mul = 0; f.shr = 0; return f; }
    int lg = 0;
    while ((1u << lg) < (uint32_t)d) ++lg;
    const int pw = 31 + lg;
    f.mul = (uint32_t)(((1ull << pw) + (uint32_t)d - 1) / (uint32_t)d);
    f.shr = (uint32_t)(pw - 32);
    return f;
}
__device__ __forceinline__ uint32_t fdiv(uint32_t x, const FastDiv& f) { return f.d == 1 ? x : (__umulhi(x, f.mul) >> f.shr); }

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (fn == nullptr) {
        void* ptr = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) != cudaSuccess ||
            qres != cudaDriverEntryPointSuccess)
            return nullptr;
        fn = reinterpret_cast<EncodeTiledFn>(ptr);
    }
    return fn;
}



// ---- fused InstanceNorm + ReLU (+ skip) phase of the convolution kernels (after their grid barrier) -------------------
// One WARP turns one output tile [npl d-planes][16 h][8 w] x [NCV 8-channel vectors] at (n, d0, h0, w0), channels from
// co0, of the raw conv output y into the activation a = relu((y - mean) * rstd) + skip.  tab = shared-memory table
// [N * C][2] of (mean, rstd).  UNROLL independent 16-byte loads are in flight per lane (the pass is pure memory
// latency: y was written by this kernel a few microseconds ago and is read back from L2).
template <int NCV, int UNROLL>
__device__ __forceinline__ void in_relu_apply_tile(const bf16* __restrict__ y, const bf16* __restrict__ skip, bf16* __restrict__ a,
                                                   const float* tab, int lane, int n, int d0, int h0, int w0, int co0, int npl,
                                                   int D, int H, int W, int C) {
    const int nvec = npl * 128 * NCV;
    const uint32_t tab_addr = smem_u32(tab + ((long long)n * C + co0) * 2);
    for (int v0 = lane; v0 < nvec; v0 += 32 * UNROLL) {
        uint4 raw[UNROLL], sk[UNROLL];
        long long off[UNROLL];
        bool ok[UNROLL];
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            const int v = v0 + 32 * u;
            const int c8 = v % NCV, r = (v / NCV) & 127, pl = v / (NCV * 128);
            const int gh = h0 + (r >> 3), gw = w0 + (r & 7);
            ok[u] = v < nvec && gh < H && gw < W && co0 + c8 * 8 < C;
            off[u] = ((((long long)n * D + d0 + pl) * H + gh) * (long long)W + gw) * C + co0 + c8 * 8;
            raw[u] = make_uint4(0u, 0u, 0u, 0u);
            sk[u] = make_uint4(0u, 0u, 0u, 0u);
            if (ok[u]) {
                raw[u] = __ldcg(reinterpret_cast<const uint4*>(y + off[u]));
                if (skip != nullptr) sk[u] = __ldg(reinterpret_cast<const uint4*>(skip + off[u]));
            }
        }
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            if (!ok[u]) continue;
            const int c8 = (v0 + 32 * u) % NCV;
            const uint32_t yu[4] = {raw[u].x, raw[u].y, raw[u].z, raw[u].w};
            const uint32_t su[4] = {sk[u].x, sk[u].y, sk[u].z, sk[u].w};
            float o[8];
#pragma unroll
            for (int i2 = 0; i2 < 4; ++i2) {
                // (mean, rstd) pairs of channels 2*i2 and 2*i2+1 of this vector
                const float4 mr = lds128(tab_addr + (uint32_t)(c8 * 8 + 2 * i2) * 8u);
                const float y0 = __uint_as_float(yu[i2] << 16), y1 = __uint_as_float(yu[i2] & 0xffff0000u);
                o[2 * i2] = fmaxf((y0 - mr.x) * mr.y, 0.f) + __uint_as_float(su[i2] << 16);
                o[2 * i2 + 1] = fmaxf((y1 - mr.z) * mr.w, 0.f) + __uint_as_float(su[i2] & 0xffff0000u);
            }
            Store<bf16>::st8(a + off[u], o);
        }
    }
}

}  // namespace
