// 3x3x3 convolution backward-filter on the 5th-gen tensor cores (sm_100a).
// Replaces cuDNN wgrad for joint_model.py:40-46,106 (the Conv3d layers with Cin = 8 or a multiple of 16).
//
//   dw[co][ci][kd,kh,kw] = sum over voxels v of  dy[v][co] * x[v + (kd,kh,kw) - 1][ci]
//
// GEMM view: the reduction (K) dimension is the VOXEL, 16 voxels per tcgen05.mma (kind::f16, bf16 -> fp32 in TMEM):
//     D[64 co, 24 = (kw, 8 ci)] += A[64 co, 16 vox] * B[24, 16 vox]^T        for every (kd, kh, 8-channel group of ci).
// Both operands are "MN-major" (the voxel is the slow dimension of NDHWC data), which tcgen05 takes directly from
// shared memory through the descriptor's major-ness bits -- nothing is transposed or im2col'ed:
//   * A = a 4x16x8-voxel tile of dy, staged by TMA as 8-channel planes [voxel][8 co] (16 B per voxel).  A core matrix
//     is 8 consecutive w-voxels x 16 B; the two K core matrices of one MMA are two consecutive h-rows (LBO = 128 B),
//     the eight M groups are the eight channel planes (SBO = plane stride).
//   * B = the 6x18x10 halo of x for the same tile, staged ONCE by TMA (out-of-bounds zero fill = the padding) as
//     8-channel planes.  For a filter tap (kd, kh) the operand is just a different start address in that halo; the
//     three kw taps are the three N groups (SBO = 16 B = the next voxel), the K core matrices are consecutive halo
//     rows (LBO = 160 B).  Every input voxel is fetched once per CTA and reused by all 27 taps.
// The 9 (or 18 for a 16-channel slice) accumulators D live in TMEM for the CTA's whole life (persistent over tiles);
// they are read back once and added to dw with fp32 atomics (dw is zeroed by the host unless accumulating).
//
// Roles: warp 0 = TMA producer, warps 1-4 = MMA issuers (each owns every fourth accumulator, so concurrent issue
// never targets the same TMEM columns) and, at the end, the read-back of their TMEM lane quadrant.
#include "tc_ptx.cuh"

namespace {

constexpr int TD = 4, TH = 16, TW = 8;
constexpr int HD = TD + 2, HH = TH + 2, HW = TW + 2;
constexpr int HV = HD * HH * HW;
constexpr int PLANE_BYTES = HV * 16;                    // 17280: one 8-channel plane of the x halo
constexpr int DYPLANE_BYTES = TD * TH * TW * 16;        // 8192: one 8-channel plane of the dy tile
constexpr int NTHREADS = 160;
constexpr int MROWS = 64;                               // MMA M (output channels per CTA)

struct WgParams {
    int n, d, h, w, cin, cout;
    int tiles_d, tiles_h, tiles_w, tiles_per_n;
    int tiles;              // n * tiles_per_n
    int kslices;            // cin / 16 (1 for cin == 8)
    int dy_merged;          // dy map is the 4-D (w*c merged) one (cout == 8)
    float* dw;
};

// kind::f16 instruction descriptor, D fp32, A/B bf16, BOTH operands MN-major, M = 64
__host__ __device__ constexpr uint32_t make_idesc_mn(int n) {
    return (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(n >> 3) << 17) |
           ((uint32_t)(MROWS >> 4) << 24);
}

// dst[0..3] += v as ONE 16-byte L2 atomic when dst is 16-byte aligned (sm_90+), else four scalar atomics.  The
// read-back of the accumulators is bound by the L2 atomic unit (every CTA adds its partial dw onto the same lines).
__device__ __forceinline__ void atomic_add4(float* dst, float4 v) {
    if (vs_aligned16_dev(dst)) {
        atomicAdd(reinterpret_cast<float4*>(dst), v);
    } else {
        atomicAdd(dst + 0, v.x); atomicAdd(dst + 1, v.y); atomicAdd(dst + 2, v.z); atomicAdd(dst + 3, v.w);
    }
}
constexpr int WROW = 216 + 4;      // staged (co, 8-ci group) block: [c8][27] floats, padded row

template <bool CIN8, int NCOG, int NSTAGE>
__global__ void __launch_bounds__(NTHREADS, 1) conv3_wgrad_tc_kernel(const __grid_constant__ CUtensorMap xmap,
                                                                     const __grid_constant__ CUtensorMap dymap, WgParams p) {
    constexpr int XP = CIN8 ? 1 : 2;                          // x planes per stage
    constexpr int NG = 9 * XP;                                // accumulators: (8-channel group, kd, kh)
    constexpr int DY_BYTES = NCOG * DYPLANE_BYTES;
    constexpr int STAGE_BYTES = DY_BYTES + XP * PLANE_BYTES;
    constexpr int TAIL = (8 * DYPLANE_BYTES > STAGE_BYTES) ? (8 * DYPLANE_BYTES - STAGE_BYTES) : 0;   // see launch_wg
    constexpr int TMEM_COLS = NG * 24 <= 256 ? 256 : 512;
    static_assert(STAGE_BYTES % 128 == 0, "stage alignment");

    extern __shared__ __align__(128) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~uintptr_t(127));
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + NSTAGE * STAGE_BYTES + TAIL);
    uint64_t* empty_bar = full_bar + NSTAGE;
    uint64_t* done_bar = empty_bar + NSTAGE;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(done_bar + 1);

    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
    const int lane = threadIdx.x & 31;
    const int ks = blockIdx.y % p.kslices;                    // 16-channel slice of ci
    const int mc = blockIdx.y / p.kslices;                    // 64-channel chunk of co

    if (threadIdx.x == 0) {
        prefetch_tensormap(&xmap);
        prefetch_tensormap(&dymap);
        for (int s = 0; s < NSTAGE; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 4); }
        mbar_init(done_bar, 4);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);
    const bool has_work = (int)blockIdx.x < p.tiles;

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            uint32_t stage = 0, phase = 0;
            for (int tile = blockIdx.x; tile < p.tiles; tile += gridDim.x) {
                const int n = tile / p.tiles_per_n;
                int r = tile % p.tiles_per_n;
                const int w0 = (r % p.tiles_w) * TW; r /= p.tiles_w;
                const int h0 = (r % p.tiles_h) * TH; r /= p.tiles_h;
                const int d0 = r * TD;
                mbar_wait(&empty_bar[stage], phase ^ 1);
                uint8_t* sdy = smem + stage * STAGE_BYTES;
                uint8_t* sx = sdy + DY_BYTES;
                mbar_expect_tx(&full_bar[stage], STAGE_BYTES);
                if (p.dy_merged) {
                    tma_load_4d(sdy, &dymap, &full_bar[stage], w0 * 8, h0, d0, n);
                } else {
#pragma unroll
                    for (int c = 0; c < NCOG; ++c)            // planes past Cout are zero-filled by TMA
                        tma_load_5d(sdy + c * DYPLANE_BYTES, &dymap, &full_bar[stage], (mc * 8 + c) * 8, w0, h0, d0, n);
                }
                if (CIN8) {
                    tma_load_4d(sx, &xmap, &full_bar[stage], (w0 - 1) * 8, h0 - 1, d0 - 1, n);
                } else {
                    tma_load_5d(sx, &xmap, &full_bar[stage], ks * 16, w0 - 1, h0 - 1, d0 - 1, n);
                    tma_load_5d(sx + PLANE_BYTES, &xmap, &full_bar[stage], ks * 16 + 8, w0 - 1, h0 - 1, d0 - 1, n);
                }
                if (++stage == NSTAGE) { stage = 0; phase ^= 1; }
            }
        }
    } else {
        // ===================== MMA issuers (warps 1-4), then TMEM read-back =====================
        const int j = warp - 1;
        constexpr uint32_t idesc = make_idesc_mn(24);
        // this warp's accumulators g = j, j+4, ...: start-address offset (16-byte units) of tap (cg, kd, kh) in the
        // halo and TMEM column, computed once
        constexpr int MAXG = (NG + 3) / 4;
        uint32_t goff[MAXG], gcol[MAXG];
#pragma unroll
        for (int i = 0; i < MAXG; ++i) {
            const int g = j + 4 * i;
            const int cg = g / 9, kd = (g % 9) / 3, kh = g % 3;
            goff[i] = (uint32_t)(cg * PLANE_BYTES + ((kd * HH + kh) * HW) * 16) >> 4;
            gcol[i] = tmem_base + g * 24;
        }
        uint32_t stage = 0, phase = 0;
        bool first = true;
        for (int tile = blockIdx.x; tile < p.tiles; tile += gridDim.x) {
            int r = tile % p.tiles_per_n;
            r /= p.tiles_w;
            const int h0 = (r % p.tiles_h) * TH;
            const int d0 = (r / p.tiles_h) * TD;
            const int jd_end = min(TD, p.d - d0), h2_end = min(TH / 2, (p.h - h0 + 1) / 2);   // rows past the volume are zero
            mbar_wait(&full_bar[stage], phase);
            tc_fence_after();
            const uint32_t sdy = smem_u32(smem + stage * STAGE_BYTES);
            const uint64_t ad0 = make_desc(sdy, TW * 16u, DYPLANE_BYTES);
            const uint64_t bd0 = make_desc(sdy + DY_BYTES, HW * 16u, 16u);
            // the first K-step of the first tile overwrites the accumulators; every tile has >= 1 valid K-step
            uint32_t acc = first ? 0u : 1u;
#pragma unroll 1
            for (int jd = 0; jd < jd_end; ++jd) {
#pragma unroll 1
                for (int h2 = 0; h2 < h2_end; ++h2) {        // 16 voxels = two h-rows of one d-plane
                    const uint64_t ad = ad0 + (uint64_t)((jd * TH * TW + h2 * 2 * TW));           // 16-byte units
                    const uint64_t bdk = bd0 + (uint64_t)((jd * HH + h2 * 2) * HW);
#pragma unroll
                    for (int i = 0; i < MAXG; ++i)
                        if (j + 4 * i < NG) tc_mma_elect(gcol[i], ad, bdk + goff[i], idesc, acc);
                    acc = 1u;
                }
            }
            tc_commit_elect(&empty_bar[stage]);
            first = false;
            if (++stage == NSTAGE) { stage = 0; phase ^= 1; }
        }
        tc_commit_elect(done_bar);
        if (has_work) {
            mbar_wait(done_bar, 0);
            tc_fence_after();
            // M = 64 accumulator rows live in lanes (row & 15) + 32 * (row >> 4): warp quadrant q holds rows 16q..16q+15
            const int q = warp & 3;
            const int co = mc * MROWS + q * 16 + lane;
            // dw[co][ci][27]: for one row and one 8-channel group of ci the 9 accumulators x 24 columns are 216
            // CONTIGUOUS floats.  Stage them in that order in shared memory (the operand stages are free now: every
            // MMA of the CTA has completed) and add them with 54 16-byte atomics per row instead of 216 scalar ones.
            float* stage_f = reinterpret_cast<float*>(smem) + (size_t)q * 16 * WROW;
#pragma unroll 1
            for (int cg = 0; cg < XP; ++cg) {
#pragma unroll 1
                for (int t9 = 0; t9 < 9; ++t9) {
                    uint32_t r0[8], r1[8], r2[8];
                    const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (cg * 9 + t9) * 24;
                    tmem_ld8(taddr, r0);
                    tmem_ld8(taddr + 8, r1);
                    tmem_ld8(taddr + 16, r2);
                    tmem_ld_wait();
                    if (lane < 16) {
                        float* dst = stage_f + lane * WROW + t9 * 3;
#pragma unroll
                        for (int c8 = 0; c8 < 8; ++c8) {
                            dst[c8 * 27 + 0] = __uint_as_float(r0[c8]);
                            dst[c8 * 27 + 1] = __uint_as_float(r1[c8]);
                            dst[c8 * 27 + 2] = __uint_as_float(r2[c8]);
                        }
                    }
                }
                __syncwarp();
                for (int i = lane; i < 16 * 54; i += 32) {
                    const int row = i / 54, f4 = i - row * 54;
                    const int cor = mc * MROWS + q * 16 + row;
                    if (cor < p.cout) {
                        const float4 v = *reinterpret_cast<const float4*>(stage_f + row * WROW + f4 * 4);
                        atomic_add4(p.dw + ((long long)cor * p.cin + ks * 16 + cg * 8) * 27 + f4 * 4, v);
                    }
                }
                __syncwarp();
            }
            (void)co;
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, TMEM_COLS);
    }
}

template <bool CIN8, int NCOG, int NSTAGE>
int launch_wg(const CUtensorMap& xmap, const CUtensorMap& dymap, const WgParams& p, int slices, cudaStream_t st) {
    constexpr int XP = CIN8 ? 1 : 2;
    constexpr int STAGE_BYTES = NCOG * DYPLANE_BYTES + XP * PLANE_BYTES;
    // The M = 64 operand always spans eight dy planes; with fewer real planes the MMA reads whatever follows
    // (rows of D that are never read back).  TAIL keeps that over-read of the LAST stage inside the allocation.
    constexpr int TAIL = (8 * DYPLANE_BYTES > STAGE_BYTES) ? (8 * DYPLANE_BYTES - STAGE_BYTES) : 0;
    constexpr int SMEM = NSTAGE * STAGE_BYTES + TAIL + 128 + 8 * (2 * NSTAGE + 1) + 16 + 64;
    static_assert(SMEM <= 227 * 1024, "shared memory budget");
    auto kern = conv3_wgrad_tc_kernel<CIN8, NCOG, NSTAGE>;
    static bool configured = false;
    if (!configured) {
        VS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM), "conv3_wgrad_tc smem attribute");
        configured = true;
    }
    int parts = vs_sm_count() / slices;
    if (parts < 1) parts = 1;
    if (parts > p.tiles) parts = p.tiles;
    dim3 grid((unsigned)parts, (unsigned)slices);
    kern<<<grid, NTHREADS, SMEM, st>>>(xmap, dymap, p);
    VS_CHECK_LAUNCH("conv3_wgrad_tc_kernel");
    return VS_OK;
}


// ---------------------------------------------------------------------------------------------
// Cout = 8 variant ("d-shift"): with one real 8-channel plane of dy, seven of the eight M groups of the M = 64
// operand above are wasted.  Here the M groups are THREE D-SHIFTED COPIES of the dy tile instead (M-group stride =
// one staged d-plane):   D[(m, co), (kw, ci)] += sum_v dy[v + m*e_d][co] * x[v + (1, kh-1, kw-1)][ci]
// which, with u = v + m*e_d, is the filter tap kd = 2 - m.  One MMA therefore produces the three kd taps at once:
// 3 MMAs per 16-voxel K-step (per 8-channel group of ci) instead of 9.  The reduction domain of v is extended to
// d in [-2, D) so that u covers the whole volume for every m; TMA's out-of-bounds zero fill supplies both the
// convolution padding of x and the zeros of dy outside the volume.  Staged per tile: dy planes d0 .. d0+5 (6 planes),
// x halo planes d0+1 .. d0+4 (4 planes; the kd halo moved into dy).
// Issue balance: 12 accumulators, three per issuing warp.  Cin = 8: warp j takes every fourth K-step (the four
// partial accumulators of a tap are summed by the atomics of the read-back); Cin = 16-slice: warp j takes ci group
// j >> 1 on every second K-step.
// ---------------------------------------------------------------------------------------------
constexpr int DSH_DY_BYTES = (TD + 2) * TH * TW * 16;    // 12288
constexpr int DSH_XPLANE_BYTES = TD * HH * HW * 16;      // 11520

template <bool CIN8, int NSTAGE>
__global__ void __launch_bounds__(NTHREADS, 1) conv3_wgrad_dsh_kernel(const __grid_constant__ CUtensorMap xmap,
                                                                      const __grid_constant__ CUtensorMap dymap, WgParams p) {
    constexpr int XP = CIN8 ? 1 : 2;
    constexpr int STAGE_BYTES = DSH_DY_BYTES + XP * DSH_XPLANE_BYTES;
    constexpr int NACC = 12;
    constexpr int TMEM_COLS = 512;                            // 12 x 24 columns
    static_assert(STAGE_BYTES % 128 == 0, "stage alignment");
    // the M = 64 operand spans eight plane-strided groups from its start address: the furthest read (K-step jd = 3,
    // h2 = 7, group 7) ends inside the stage, so no tail padding is needed
    static_assert(((TD - 1) * TH * TW + (TH - 2) * TW) * 16 + 7 * TH * TW * 16 + 2 * TW * 16 <= STAGE_BYTES, "A over-read");

    extern __shared__ __align__(128) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~uintptr_t(127));
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + NSTAGE * STAGE_BYTES);
    uint64_t* empty_bar = full_bar + NSTAGE;
    uint64_t* done_bar = empty_bar + NSTAGE;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(done_bar + 1);

    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
    const int lane = threadIdx.x & 31;
    const int ks = blockIdx.y;                                // 16-channel slice of ci (0 for Cin = 8)

    if (threadIdx.x == 0) {
        prefetch_tensormap(&xmap);
        prefetch_tensormap(&dymap);
        for (int s = 0; s < NSTAGE; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 4); }
        mbar_init(done_bar, 4);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);
    const bool has_work = (int)blockIdx.x < p.tiles;

    if (warp == 0) {
        if (lane == 0) {
            uint32_t stage = 0, phase = 0;
            for (int tile = blockIdx.x; tile < p.tiles; tile += gridDim.x) {
                const int n = tile / p.tiles_per_n;
                int r = tile % p.tiles_per_n;
                const int w0 = (r % p.tiles_w) * TW; r /= p.tiles_w;
                const int h0 = (r % p.tiles_h) * TH; r /= p.tiles_h;
                const int d0 = r * TD - 2;
                mbar_wait(&empty_bar[stage], phase ^ 1);
                uint8_t* sdy = smem + stage * STAGE_BYTES;
                uint8_t* sx = sdy + DSH_DY_BYTES;
                mbar_expect_tx(&full_bar[stage], STAGE_BYTES);
                tma_load_4d(sdy, &dymap, &full_bar[stage], w0 * 8, h0, d0, n);
                if (CIN8) {
                    tma_load_4d(sx, &xmap, &full_bar[stage], (w0 - 1) * 8, h0 - 1, d0 + 1, n);
                } else {
                    tma_load_5d(sx, &xmap, &full_bar[stage], ks * 16, w0 - 1, h0 - 1, d0 + 1, n);
                    tma_load_5d(sx + DSH_XPLANE_BYTES, &xmap, &full_bar[stage], ks * 16 + 8, w0 - 1, h0 - 1, d0 + 1, n);
                }
                if (++stage == NSTAGE) { stage = 0; phase ^= 1; }
            }
        }
    } else {
        const int j = warp - 1;
        constexpr uint32_t idesc = make_idesc_mn(24);
        const int kmod = CIN8 ? 4 : 2, kres = CIN8 ? j : (j & 1);        // this warp's K-steps: kcount % kmod == kres
        const int cg = CIN8 ? 0 : (j >> 1);
        const uint32_t col0 = tmem_base + (uint32_t)(j * 3) * 24;
        uint32_t stage = 0, phase = 0;
        uint32_t acc = 0u;                                               // first K-step of this warp overwrites
        for (int tile = blockIdx.x; tile < p.tiles; tile += gridDim.x) {
            int r = tile % p.tiles_per_n;
            r /= p.tiles_w;
            const int h0 = (r % p.tiles_h) * TH;
            const int h2_end = min(TH / 2, (p.h - h0 + 1) / 2);          // rows past the volume are zero
            mbar_wait(&full_bar[stage], phase);
            tc_fence_after();
            const uint32_t sdy = smem_u32(smem + stage * STAGE_BYTES);
            const uint64_t ad0 = make_desc(sdy, TW * 16u, TH * TW * 16u);
            const uint64_t bd0 = make_desc(sdy + DSH_DY_BYTES + cg * DSH_XPLANE_BYTES, HW * 16u, 16u);
            int kcount = 0;
#pragma unroll 1
            for (int jd = 0; jd < TD; ++jd) {
#pragma unroll 1
                for (int h2 = 0; h2 < h2_end; ++h2, ++kcount) {          // 16 voxels = two h-rows of one d-plane
                    if (kcount % kmod != kres) continue;
                    const uint64_t ad = ad0 + (uint64_t)((jd * TH * TW + h2 * 2 * TW));           // 16-byte units
                    const uint64_t bdk = bd0 + (uint64_t)((jd * HH + h2 * 2) * HW);
#pragma unroll
                    for (int kh = 0; kh < 3; ++kh) tc_mma_elect(col0 + kh * 24, ad, bdk + (uint64_t)(kh * HW), idesc, acc);
                    acc = 1u;
                }
            }
            tc_commit_elect(&empty_bar[stage]);
            if (++stage == NSTAGE) { stage = 0; phase ^= 1; }
        }
        tc_commit_elect(done_bar);
        if (has_work) {
            mbar_wait(done_bar, 0);
            tc_fence_after();
            // M = 64 accumulator rows live in lanes (row & 15) + 32 * (row >> 4); row = m * 8 + co, m = 0..2 valid
            const int q = warp & 3;
            const int row = q * 16 + lane;
            const int m = row >> 3, co = row & 7;
            const bool valid = lane < 16 && m <= 2;
            if (q < 2) {
                // The K-split partial accumulators of one (ci group, kh) are summed in registers, staged in shared
                // memory (free now) in dw's own order -- [co][ci group][c8][27] -- and added with 16-byte atomics:
                // 54 per (co, ci group) instead of 216 scalar ones per partial (the L2 atomic unit bounds the tail).
                constexpr int NSETS = CIN8 ? 4 : 2, NCG = CIN8 ? 1 : 2;
                float* stage_f = reinterpret_cast<float*>(smem);                  // [8 co][NCG][WROW]
#pragma unroll 1
                for (int gcg = 0; gcg < NCG; ++gcg) {
#pragma unroll 1
                    for (int kh = 0; kh < 3; ++kh) {
                        float s0[8], s1[8], s2[8];
#pragma unroll
                        for (int c8 = 0; c8 < 8; ++c8) { s0[c8] = 0.f; s1[c8] = 0.f; s2[c8] = 0.f; }
#pragma unroll
                        for (int set = 0; set < NSETS; ++set) {
                            const int jw = CIN8 ? set : (gcg * 2 + set);          // issuing warp that owns the partial
                            uint32_t r0[8], r1[8], r2[8];
                            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (jw * 3 + kh) * 24;
                            tmem_ld8(taddr, r0);
                            tmem_ld8(taddr + 8, r1);
                            tmem_ld8(taddr + 16, r2);
                            tmem_ld_wait();
#pragma unroll
                            for (int c8 = 0; c8 < 8; ++c8) {
                                s0[c8] += __uint_as_float(r0[c8]); s1[c8] += __uint_as_float(r1[c8]); s2[c8] += __uint_as_float(r2[c8]);
                            }
                        }
                        if (valid) {
                            float* dst = stage_f + (co * NCG + gcg) * WROW + (2 - m) * 9 + kh * 3;
#pragma unroll
                            for (int c8 = 0; c8 < 8; ++c8) {
                                dst[c8 * 27 + 0] = s0[c8];
                                dst[c8 * 27 + 1] = s1[c8];
                                dst[c8 * 27 + 2] = s2[c8];
                            }
                        }
                    }
                }
                asm volatile("bar.sync 2, 64;" ::: "memory");                     // the two read-back warps (q = 0, 1)
                const int t64 = q * 32 + lane;
                for (int i = t64; i < 8 * NCG * 54; i += 64) {
                    const int blk = i / 54, f4 = i - blk * 54;
                    const int cor = blk / NCG, gcg = blk % NCG;
                    const float4 v = *reinterpret_cast<const float4*>(stage_f + blk * WROW + f4 * 4);
                    atomic_add4(p.dw + ((long long)cor * p.cin + ks * 16 + gcg * 8) * 27 + f4 * 4, v);
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, TMEM_COLS);
    }
}

template <bool CIN8, int NSTAGE>
int launch_wg_dsh(const CUtensorMap& xmap, const CUtensorMap& dymap, const WgParams& p, int slices, cudaStream_t st) {
    constexpr int XP = CIN8 ? 1 : 2;
    constexpr int STAGE_BYTES = DSH_DY_BYTES + XP * DSH_XPLANE_BYTES;
    constexpr int SMEM = NSTAGE * STAGE_BYTES + 128 + 8 * (2 * NSTAGE + 1) + 16 + 64;
    static_assert(SMEM <= 227 * 1024, "shared memory budget");
    auto kern = conv3_wgrad_dsh_kernel<CIN8, NSTAGE>;
    static bool configured = false;
    if (!configured) {
        VS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM), "conv3_wgrad_dsh smem attribute");
        configured = true;
    }
    int parts = vs_sm_count() / slices;
    if (parts < 1) parts = 1;
    if (parts > p.tiles) parts = p.tiles;
    dim3 grid((unsigned)parts, (unsigned)slices);
    kern<<<grid, NTHREADS, SMEM, st>>>(xmap, dymap, p);
    VS_CHECK_LAUNCH("conv3_wgrad_dsh_kernel");
    return VS_OK;
}

template <bool CIN8>
int dispatch_wg(int ncog, const CUtensorMap& xmap, const CUtensorMap& dymap, const WgParams& p, int slices, cudaStream_t st) {
    if (ncog <= 1) return launch_wg<CIN8, 1, 4>(xmap, dymap, p, slices, st);
    if (ncog <= 2) return launch_wg<CIN8, 2, 4>(xmap, dymap, p, slices, st);
    if (ncog <= 4) return launch_wg<CIN8, 4, 3>(xmap, dymap, p, slices, st);
    return launch_wg<CIN8, 8, 2>(xmap, dymap, p, slices, st);
}

}  // namespace

static int g_wgrad_dsh = 1;
// development switch (A/B runs, tests): 0 disables the d-shift variant for Cout = 8
extern "C" void vs_debug_set_wgrad_dsh(int on) { g_wgrad_dsh = on; }

extern "C" int vs_conv3_wgrad_tc_eligible(int cin, int cout) {
    return (cin == 8 || (cin % 16 == 0 && cin >= 16)) && cout % 8 == 0 && cout >= 8;
}

// dw[Cout][Cin][27] += sum_v dy[v,co] * x[v+tap,ci]; x, dy bf16 NDHWC; dw fp32 (already zeroed or holding the value to
// accumulate onto).
extern "C" int vs_conv3x3x3_wgrad_tc(const void* x, const void* dy, float* dw, int n, int d, int h, int w, int cin, int cout,
                                     void* stream) {
    VS_REQUIRE(x && dy && dw, VS_ERR_SHAPE, "conv3_wgrad_tc: null pointer");
    VS_REQUIRE(vs_conv3_wgrad_tc_eligible(cin, cout), VS_ERR_UNSUPPORTED, "conv3_wgrad_tc: unsupported channels Cin=%d Cout=%d", cin, cout);
    VS_REQUIRE(vs_aligned16(x) && vs_aligned16(dy), VS_ERR_ALIGN, "conv3_wgrad_tc: pointers must be 16B aligned");
    EncodeTiledFn encode = get_encode_fn();
    VS_REQUIRE(encode != nullptr, VS_ERR_CUDA, "conv3_wgrad_tc: cuTensorMapEncodeTiled unavailable");
    cudaStream_t st = (cudaStream_t)stream;
    const cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    const bool dsh = cout == 8 && g_wgrad_dsh;                 // d-shift variant (see conv3_wgrad_dsh_kernel)
    const cuuint32_t xbd = dsh ? TD : HD, dybd = dsh ? TD + 2 : TD;
    CUtensorMap xmap, dymap;
    CUresult cr;
    if (cin == 8) {
        const cuuint64_t gdim[4] = {(cuuint64_t)w * 8, (cuuint64_t)h, (cuuint64_t)d, (cuuint64_t)n};
        const cuuint64_t gstr[3] = {(cuuint64_t)w * 16, (cuuint64_t)h * w * 16, (cuuint64_t)d * h * w * 16};
        const cuuint32_t box[4] = {HW * 8, HH, xbd, 1};
        cr = encode(&xmap, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(x), gdim, gstr, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    } else {
        const cuuint64_t gdim[5] = {(cuuint64_t)cin, (cuuint64_t)w, (cuuint64_t)h, (cuuint64_t)d, (cuuint64_t)n};
        const cuuint64_t gstr[4] = {(cuuint64_t)cin * 2, (cuuint64_t)w * cin * 2, (cuuint64_t)h * w * cin * 2,
                                    (cuuint64_t)d * h * w * cin * 2};
        const cuuint32_t box[5] = {8, HW, HH, xbd, 1};
        cr = encode(&xmap, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(x), gdim, gstr, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    }
    VS_REQUIRE(cr == CUDA_SUCCESS, VS_ERR_CUDA, "conv3_wgrad_tc: cuTensorMapEncodeTiled(x) failed (%d)", (int)cr);
    if (cout == 8) {
        const cuuint64_t gdim[4] = {(cuuint64_t)w * 8, (cuuint64_t)h, (cuuint64_t)d, (cuuint64_t)n};
        const cuuint64_t gstr[3] = {(cuuint64_t)w * 16, (cuuint64_t)h * w * 16, (cuuint64_t)d * h * w * 16};
        const cuuint32_t box[4] = {TW * 8, TH, dybd, 1};
        cr = encode(&dymap, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(dy), gdim, gstr, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    } else {
        const cuuint64_t gdim[5] = {(cuuint64_t)cout, (cuuint64_t)w, (cuuint64_t)h, (cuuint64_t)d, (cuuint64_t)n};
        const cuuint64_t gstr[4] = {(cuuint64_t)cout * 2, (cuuint64_t)w * cout * 2, (cuuint64_t)h * w * cout * 2,
                                    (cuuint64_t)d * h * w * cout * 2};
        const cuuint32_t box[5] = {8, TW, TH, TD, 1};
        cr = encode(&dymap, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(dy), gdim, gstr, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    }
    VS_REQUIRE(cr == CUDA_SUCCESS, VS_ERR_CUDA, "conv3_wgrad_tc: cuTensorMapEncodeTiled(dy) failed (%d)", (int)cr);

    WgParams p;
    p.n = n; p.d = d; p.h = h; p.w = w; p.cin = cin; p.cout = cout;
    p.tiles_d = ((dsh ? d + 2 : d) + TD - 1) / TD; p.tiles_h = (h + TH - 1) / TH; p.tiles_w = (w + TW - 1) / TW;
    p.tiles_per_n = p.tiles_d * p.tiles_h * p.tiles_w;
    const long long tiles = (long long)n * p.tiles_per_n;
    VS_REQUIRE(tiles < 2147483647LL, VS_ERR_SHAPE, "conv3_wgrad_tc: too many tiles");
    p.tiles = (int)tiles;
    p.kslices = cin == 8 ? 1 : cin / 16;
    p.dy_merged = cout == 8;
    p.dw = dw;
    const int mchunks = (cout + MROWS - 1) / MROWS;
    const int slices = p.kslices * mchunks;
    const int ncog = cout >= MROWS ? 8 : cout / 8;            // planes of the (possibly only) 64-channel chunk
    if (dsh) {
        if (cin == 8) return launch_wg_dsh<true, 4>(xmap, dymap, p, slices, st);
        return launch_wg_dsh<false, 4>(xmap, dymap, p, slices, st);
    }
    if (cin == 8) return dispatch_wg<true>(ncog, xmap, dymap, p, slices, st);
    return dispatch_wg<false>(ncog, xmap, dymap, p, slices, st);
}
