// Conv3d(C,C,2,stride 2) and ConvTranspose3d(C,C,2,stride 2) on the 5th-gen tensor cores (sm_100a): tcgen05.mma with
// TMEM accumulators, operands staged by TMA.  Replaces cuDNN for joint_model.py:130 (Down) and :118 (Up): forward and
// input gradient of both layers are the two kernels below (the same "gather" / "scatter" algebra as k2s2.cu, which
// stays for the fp32 check mode and the weight gradient).
//
// A 2x2x2 stride-2 convolution touches non-overlapping patches, so it is a plain GEMM over the coarse voxels:
//
//   gather  (Conv3d fprop, ConvTranspose3d dgrad):  coarse[o, a] = bias[a] + sum_{k,b} wt[a][b][k] * fine[2o+k, b]
//       D[128 coarse voxels, A] += X[128, K = (kd,kh,kw,b)] * Wg[A, K]^T                       K = 8 B
//   scatter (ConvTranspose3d fprop, Conv3d dgrad):  fine[2o+k, b] = bias[b] + sum_a wt[a][b][k] * coarse[o, a]
//       D[128 coarse voxels, N = (kd,kh,kw,b)] += X[128, K = a] * Ws[N, K]^T                   N = 8 B
//
// M tile = 16 x 8 (h, w) coarse voxels of one coarse d-plane (row r = h * 8 + w: eight consecutive w-voxels are one
// 8-row core-matrix group, the 16 h-rows are the 16 groups).  Operands are K-major:
//   * gather: in NDHWC the fine voxels (2w, 2w+1) of a coarse voxel are adjacent, so for a fixed (kd,kh) the 2B values
//     (kw, b) are ONE contiguous run: the fine tensor is viewed as [n*Df][Hc][kh 2][Wc][2B] and a TMA box
//     [16 h][8 w][run] lands as 128 operand rows of min(4B, 128) bytes in the matching 32 / 64 / 128-byte SWIZZLE
//     layout -- no im2col, every input byte is read once, four boxes per tile instead of one per 8 channels
//     (measured 26.9 -> 20.6 us at 2 x 96^3 -> 48^3, 16 channels; `vs_debug_set_k2_tc(0)` keeps the 8-channel planes);
//   * scatter: the A operand is the coarse tile itself (A/8 no-swizzle planes of [128 rows][16 B]); the 8B accumulator
//     columns of a row are the eight output voxels of that coarse voxel, stored as contiguous (kw,b) runs.
// Pipeline: warp 0 = TMA producer, warp 1 = MMA issuer (+ TMEM owner), then NEPI epilogue warpgroups; shared-memory
// stages ring over (tile, K slice); there are NEPI TMEM accumulator buffers and warpgroup e drains buffer e (tiles
// e, e + NEPI, ... of the CTA), so NEPI epilogues run concurrently with the loads and MMAs of later tiles.  The
// epilogue is the long pole of these kernels (ncu, profiles/r2_k2s2_tc_ncu.txt: ~1300 instructions per warp and tile
// on ONE warp per scheduler in the first version): work-item decode by multiply-high, one division per tile, 32-column
// TMEM loads, 128-bit shared loads of the bias, 32-bit address arithmetic.  Persistent grid = min(work items, SMs).
// These layers are HBM-bound (AI 7-114 flop/B): the point of the tensor cores here is that the 8 B
// (or 8 A) MACs per byte no longer cost CUDA-core issue slots, so the kernel runs at the speed of its loads / stores.
#include "tc_ptx.cuh"
#include "tc_pack.cuh"

namespace {

constexpr int K2_TH = 16, K2_TW = 8;
constexpr int K2_PLANE = 128 * 16;             // bytes of one 8-channel plane of a 128-row tile

struct K2TcParams {
    int n, dc, hc, wc, a, b;                   // coarse dims, coarse channels A, fine channels B
    int tiles_h, tiles_w, nchunks;             // N chunks of NC accumulator columns
    FastDiv div_chunks, div_tw, div_th, div_dc, div_b;
    int kstages;                               // K slices per tile
    int ks;                                    // K elements per slice (multiple of 16, <= 128)
    int planes;                                // TMA boxes ("planes") the producer loads per slice
    int wbytes;                                // bytes per operand row of a plane: 16 = 8-channel planes (no swizzle);
                                               // 32 / 64 / 128 = swizzled rows (gather: the whole (kw,b) run of a row)
    int work_items;
    const bf16* wpack;
    const float* bias;
    bf16* out;
};

__device__ __forceinline__ void k2_decode(int item, const K2TcParams& p, int& n, int& d, int& h0, int& w0, int& chunk) {
    uint32_t t = fdiv((uint32_t)item, p.div_chunks);
    chunk = item - (int)t * p.nchunks;
    uint32_t q = fdiv(t, p.div_tw);
    w0 = (int)(t - q * (uint32_t)p.tiles_w) * K2_TW; t = q;
    q = fdiv(t, p.div_th);
    h0 = (int)(t - q * (uint32_t)p.tiles_h) * K2_TH; t = q;
    q = fdiv(t, p.div_dc);
    d = (int)(t - q * (uint32_t)p.dc);
    n = (int)q;
}

// SCATTER = false: gather (A operand = fine tensor through the [2B][Wc][kh][Hc][n*Df] view, N = A channels)
// SCATTER = true : scatter (A operand = coarse tile, N = 8 B)
template <int NC, int NSTAGE, int NEPI, bool SCATTER>
__global__ void __launch_bounds__(128 + 128 * NEPI, 1) k2s2_tc_kernel(const __grid_constant__ CUtensorMap xmap, K2TcParams p) {
    constexpr int NBUF = NEPI;
    constexpr int TMEM_COLS = (NBUF * NC) <= 32 ? 32 : (NBUF * NC) <= 64 ? 64 : (NBUF * NC) <= 128 ? 128 : (NBUF * NC) <= 256 ? 256 : 512;
    static_assert(NBUF * NC <= 512, "TMEM budget");
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const int a_bytes = (p.ks / 8) * K2_PLANE;                 // = 128 rows x ks x 2 bytes whatever the plane width
    const int b_bytes = NC * p.ks * 2;
    const int stage_bytes = a_bytes + b_bytes;                 // multiple of 1024 (host-checked)
    const int plane_bytes = 128 * p.wbytes;
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + NSTAGE * stage_bytes);
    uint64_t* empty_bar = full_bar + NSTAGE;
    uint64_t* tfull_bar = empty_bar + NSTAGE;
    uint64_t* tempty_bar = tfull_bar + NBUF;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + NBUF);
    float* sbias = reinterpret_cast<float*>(tmem_slot + 4);              // [256], 16-byte aligned

    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
    const int lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        prefetch_tensormap(&xmap);
        for (int s = 0; s < NSTAGE; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        for (int b = 0; b < NBUF; ++b) { mbar_init(&tfull_bar[b], 1); mbar_init(&tempty_bar[b], 4); }
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, TMEM_COLS);
    pdl_trigger();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);
    pdl_wait();                          // see vs_common.cuh: no global-memory access above this line

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            uint32_t stage = 0, phase = 0;
            for (int item = blockIdx.x; item < p.work_items; item += gridDim.x) {
                int n, d, h0, w0, chunk;
                k2_decode(item, p, n, d, h0, w0, chunk);
                for (int s = 0; s < p.kstages; ++s) {
                    mbar_wait(&empty_bar[stage], phase ^ 1);
                    uint8_t* sa = smem + stage * stage_bytes;
                    mbar_expect_tx(&full_bar[stage], (uint32_t)(p.planes * plane_bytes + b_bytes));
                    for (int j = 0; j < p.planes; ++j) {
                        const int k0 = s * p.ks + j * (p.wbytes / 2);
                        if (SCATTER) {
                            tma_load_5d(sa + j * plane_bytes, &xmap, &full_bar[stage], k0, w0, h0, d, n);
                        } else {
                            const int row = k0 / (2 * p.b), col = k0 - row * (2 * p.b);      // row = (kd, kh)
                            tma_load_5d(sa + j * plane_bytes, &xmap, &full_bar[stage], col, w0, row & 1, h0,
                                        n * (2 * p.dc) + 2 * d + (row >> 1));
                        }
                    }
                    bulk_load(sa + a_bytes, reinterpret_cast<const uint8_t*>(p.wpack) +
                              ((long long)chunk * p.kstages + s) * b_bytes, (uint32_t)b_bytes, &full_bar[stage]);
                    if (++stage == NSTAGE) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        constexpr uint32_t idesc = make_idesc(NC);
        uint32_t stage = 0, phase = 0, buf = 0, bphase = 0;
        const int nmma = p.ks / 16;
        // A = 8 (scatter, K = 8): the second K chunk of the single MMA re-reads the first plane (LBO = 0) against
        // zero-packed weights
        const uint32_t lbo = (p.wbytes == 16 && p.planes * 8 < p.ks) ? 0u : (uint32_t)K2_PLANE;
        for (int item = blockIdx.x; item < p.work_items; item += gridDim.x) {
            mbar_wait(&tempty_bar[buf], bphase ^ 1);
            tc_fence_after();
            const uint32_t dcol = tmem_base + buf * NC;
            for (int s = 0; s < p.kstages; ++s) {
                mbar_wait(&full_bar[stage], phase);
                tc_fence_after();
                const uint32_t a_base = smem_u32(smem + stage * stage_bytes);
                const uint32_t b_base = a_base + a_bytes;
#pragma unroll 1
                for (int m = 0; m < nmma; ++m) {
                    uint64_t ad;
                    if (p.wbytes == 16) {
                        ad = make_desc(a_base + (uint32_t)m * 2u * K2_PLANE, lbo, 128u);
                    } else {
                        const uint32_t koff = (uint32_t)m * 32u;                          // byte offset of this K step in the row run
                        ad = make_desc_sw(a_base + (koff / (uint32_t)p.wbytes) * (uint32_t)plane_bytes + koff % (uint32_t)p.wbytes,
                                          (uint32_t)p.wbytes);
                    }
                    const uint64_t bd = make_desc(b_base + (uint32_t)m * (NC * 32), NC * 16, 128u);
                    tc_mma_elect(dcol, ad, bd, idesc, (uint32_t)((s | m) != 0));
                }
                tc_commit_elect(&empty_bar[stage]);
                if (++stage == NSTAGE) { stage = 0; phase ^= 1; }
            }
            tc_commit_elect(&tfull_bar[buf]);
            if (++buf == NBUF) { buf = 0; bphase ^= 1; }
        }
    } else if (warp >= 4) {
        // ===================== epilogue: warpgroup e = (warp - 4) / 4 drains TMEM buffer e =====================
        const int q = warp & 3;
        const int e = (warp - 4) >> 2;
        const int row = q * 32 + lane;
        const int lh = row >> 3, lw = row & 7;
        const uint32_t lane_base = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(e * NC);
        const uint32_t sbias_addr = smem_u32(sbias);
        {
            // the bias (a parameter) is read after pdl_wait like everything else; only the epilogue warps use it
            const int nbias = SCATTER ? p.b : p.a;
            for (int i = (int)threadIdx.x - 128; i < nbias; i += 128 * NEPI) sbias[i] = p.bias != nullptr ? p.bias[i] : 0.f;
            asm volatile("bar.sync 1, %0;" ::"n"(128 * NEPI) : "memory");
        }
        uint32_t bphase = 0;
        const int wf = 2 * p.wc, hwf = 4 * p.hc * p.wc;          // fine row / plane pitch in voxels (32-bit: host-checked)
        for (int item = blockIdx.x + e * (int)gridDim.x; item < p.work_items; item += NEPI * (int)gridDim.x) {
            int n, d, h0, w0, chunk;
            k2_decode(item, p, n, d, h0, w0, chunk);
            const int gh = h0 + lh, gw = w0 + lw;
            const bool ok = gh < p.hc && gw < p.wc;
            mbar_wait(&tfull_bar[e], bphase);
            tc_fence_after();
            if (!SCATTER) {
                bf16* po = p.out + ((((long long)n * p.dc + d) * p.hc + gh) * (long long)p.wc + gw) * p.a + chunk * NC;
#pragma unroll
                for (int c16 = 0; c16 < NC / 16; ++c16) {
                    uint32_t r[16];
                    tmem_ld16(lane_base + c16 * 16, r);
                    tmem_ld_wait();
                    const int ch0 = chunk * NC + c16 * 16;
#pragma unroll
                    for (int h8 = 0; h8 < 2; ++h8) {
                        if (ok && ch0 + h8 * 8 < p.a) {
                            const float4 b0v = lds128(sbias_addr + (uint32_t)(ch0 + h8 * 8) * 4u);
                            const float4 b1v = lds128(sbias_addr + (uint32_t)(ch0 + h8 * 8 + 4) * 4u);
                            float o[8];
                            o[0] = __uint_as_float(r[h8 * 8 + 0]) + b0v.x; o[1] = __uint_as_float(r[h8 * 8 + 1]) + b0v.y;
                            o[2] = __uint_as_float(r[h8 * 8 + 2]) + b0v.z; o[3] = __uint_as_float(r[h8 * 8 + 3]) + b0v.w;
                            o[4] = __uint_as_float(r[h8 * 8 + 4]) + b1v.x; o[5] = __uint_as_float(r[h8 * 8 + 5]) + b1v.y;
                            o[6] = __uint_as_float(r[h8 * 8 + 6]) + b1v.z; o[7] = __uint_as_float(r[h8 * 8 + 7]) + b1v.w;
                            Store<bf16>::st8(po + c16 * 16 + h8 * 8, o);
                        }
                    }
                }
            } else {
                // accumulator column = k * B + b with k = (kd,kh,kw); a 16-column group is 32 contiguous output bytes
                // ((kw,b) is contiguous in NDHWC; for B = 8 the group spans the voxel pair kw = 0, 1)
                bf16* rowp = p.out + ((((long long)n * 2 * p.dc + 2 * d) * (2 * p.hc) + 2 * gh) * (long long)wf + 2 * gw) * p.b;
                int k = (int)fdiv((uint32_t)(chunk * NC), p.div_b);
                int b0 = chunk * NC - k * p.b;
                const int kstep = p.b == 8 ? 2 : 1;
#pragma unroll
                for (int c32 = 0; c32 < NC / 32; ++c32) {
                    uint32_t r[32];
                    tmem_ld32(lane_base + c32 * 32, r);
                    tmem_ld_wait();
#pragma unroll
                    for (int half = 0; half < 2; ++half) {
                        if (ok && k < 8) {
                            const int off = (((k >> 2) * hwf + ((k >> 1) & 1) * wf + (k & 1)) * p.b) + b0;
                            uint32_t u[8];
#pragma unroll
                            for (int j4 = 0; j4 < 4; ++j4) {
                                int bb = b0 + j4 * 4;
                                if (bb >= p.b) bb -= p.b;                       // B = 8: the second voxel of the pair
                                const float4 bv = lds128(sbias_addr + (uint32_t)bb * 4u);
                                const __nv_bfloat162 lo = __floats2bfloat162_rn(__uint_as_float(r[half * 16 + j4 * 4 + 0]) + bv.x,
                                                                                __uint_as_float(r[half * 16 + j4 * 4 + 1]) + bv.y);
                                const __nv_bfloat162 hi = __floats2bfloat162_rn(__uint_as_float(r[half * 16 + j4 * 4 + 2]) + bv.z,
                                                                                __uint_as_float(r[half * 16 + j4 * 4 + 3]) + bv.w);
                                u[j4 * 2] = *reinterpret_cast<const uint32_t*>(&lo);
                                u[j4 * 2 + 1] = *reinterpret_cast<const uint32_t*>(&hi);
                            }
                            uint4* dst = reinterpret_cast<uint4*>(rowp + off);
                            dst[0] = make_uint4(u[0], u[1], u[2], u[3]);
                            dst[1] = make_uint4(u[4], u[5], u[6], u[7]);
                        }
                        b0 += 16;
                        if (b0 >= p.b) { b0 = 0; k += kstep; }
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty_bar[e]);
            bphase ^= 1;
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, TMEM_COLS);
    }
}

__global__ void pack_k2s2_kernel(const float* __restrict__ w, bf16* __restrict__ out, int a, int b, int scatter, long long total) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x)
        out[i] = __float2bfloat16_rn(pack_k2s2_elem(w, i, a, b, scatter));
}

template <int NC, int NSTAGE, int NEPI, bool SCATTER>
int launch_k2(const CUtensorMap& map, const K2TcParams& p, cudaStream_t st) {
    const int stage_bytes = (p.ks / 8) * K2_PLANE + NC * p.ks * 2;
    VS_REQUIRE(stage_bytes % 1024 == 0, VS_ERR_UNSUPPORTED, "k2s2_tc: stage size %d is not a multiple of 1024", stage_bytes);
    const int smem = NSTAGE * stage_bytes + 1024 + 8 * (2 * NSTAGE + 2 * NEPI) + 16 + 256 * 4 + 64;
    VS_REQUIRE(smem <= 227 * 1024, VS_ERR_UNSUPPORTED, "k2s2_tc: shared memory budget exceeded (%d bytes)", smem);
    auto kern = k2s2_tc_kernel<NC, NSTAGE, NEPI, SCATTER>;
    static int configured = 0;
    if (configured < smem) {
        VS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem), "k2s2_tc smem attribute");
        configured = smem;
    }
    const int grid = p.work_items < vs_sm_count() ? p.work_items : vs_sm_count();
    VS_CUDA(vs_launch(kern, dim3((unsigned)grid), dim3(128 + 128 * NEPI), (size_t)smem, st, map, p), "k2s2_tc_kernel launch");
    VS_CHECK_LAUNCH("k2s2_tc_kernel");
    return VS_OK;
}

// A/B switch (tools, tests): swizzled gather operand rows (1, default) or 8-channel no-swizzle planes (0)
int g_k2_swizzle = 1;

// [2B][Wc][kh 2][Hc][n*Df] view of a fine NDHWC tensor: a (kd,kh) row of a coarse voxel's patch is 2B contiguous values
// (kw, b).  box0 = columns per box; swz = CU_TENSOR_MAP_SWIZZLE_* matching box0 * 2 bytes (or NONE).
CUresult encode_fine_view(EncodeTiledFn encode, CUtensorMap* map, const void* fine, int n, int dc, int hc, int wc, int b,
                          int box0, CUtensorMapSwizzle swz) {
    const cuuint64_t gdim[5] = {(cuuint64_t)2 * b, (cuuint64_t)wc, 2, (cuuint64_t)hc, (cuuint64_t)n * 2 * dc};
    const cuuint64_t rowb = (cuuint64_t)2 * wc * b * 2;            // bytes of one fine h-row
    const cuuint64_t gstr[4] = {(cuuint64_t)2 * b * 2, rowb, 2 * rowb, (cuuint64_t)2 * hc * rowb};
    const cuuint32_t box[5] = {(cuuint32_t)box0, K2_TW, 1, K2_TH, 1};
    const cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    return encode(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(fine), gdim, gstr, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
}

int k2_tc_common(const char* who, const void* x, const void* wpack, void* out, int n, int dc, int hc, int wc, int a, int b) {
    VS_REQUIRE(n > 0 && dc > 0 && hc > 0 && wc > 0, VS_ERR_SHAPE, "%s: bad shape", who);
    VS_REQUIRE(a % 8 == 0 && b % 8 == 0 && a >= 8 && b >= 8 && a <= 256 && b <= 256 && (a == 8 || a % 16 == 0) && (b == 8 || b % 16 == 0),
               VS_ERR_UNSUPPORTED, "%s: channels must be 8 or a multiple of 16, <= 256 (A=%d B=%d)", who, a, b);
    VS_REQUIRE(x && wpack && out, VS_ERR_SHAPE, "%s: null pointer", who);
    VS_REQUIRE(vs_aligned16(x) && vs_aligned16(wpack) && vs_aligned16(out), VS_ERR_ALIGN, "%s: pointers must be 16B aligned", who);
    VS_REQUIRE(8LL * dc * hc * wc * (long long)b < 2147483647LL, VS_ERR_SHAPE, "%s: one sample of the fine tensor must stay below 2^31 elements", who);
    return VS_OK;
}

void k2_fill_common(K2TcParams& p, int n, int dc, int hc, int wc, int a, int b) {
    p.n = n; p.dc = dc; p.hc = hc; p.wc = wc; p.a = a; p.b = b;
    p.tiles_h = (hc + K2_TH - 1) / K2_TH; p.tiles_w = (wc + K2_TW - 1) / K2_TW;
    p.div_tw = make_fastdiv(p.tiles_w); p.div_th = make_fastdiv(p.tiles_h); p.div_dc = make_fastdiv(dc);
    p.div_b = make_fastdiv(b);
}

}  // namespace

extern "C" void vs_debug_set_k2_tc(int swizzle) { g_k2_swizzle = swizzle; }

// bytes of the bf16 UMMA B-operand pack of a k2s2 weight wt[A][B][8] (scatter = 0: gather GEMM, 1: scatter GEMM)
extern "C" size_t vs_k2s2_tc_pack_bytes(int a, int b, int scatter) {
    if (a < 8 || b < 8 || a > 256 || b > 256 || !(a == 8 || a % 16 == 0) || !(b == 8 || b % 16 == 0)) return 0;
    return (size_t)k2_pack_elems(a, b, scatter) * 2;
}

extern "C" int vs_pack_k2s2_weight_tc(const float* w, void* out, int a, int b, int scatter, void* stream) {
    const size_t bytes = vs_k2s2_tc_pack_bytes(a, b, scatter);
    VS_REQUIRE(w && out && bytes > 0, VS_ERR_UNSUPPORTED, "pack_k2s2_weight_tc: unsupported shape A=%d B=%d", a, b);
    const long long total = (long long)(bytes / 2);
    pack_k2s2_kernel<<<(unsigned)min(1024LL, (total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(w, (bf16*)out, a, b, scatter, total);
    VS_CHECK_LAUNCH("pack_k2s2_kernel");
    return VS_OK;
}

// coarse[n,dc,hc,wc,A] = bias + gather(fine[n,2dc,2hc,2wc,B], wpack); bf16 NDHWC in and out.
extern "C" int vs_k2s2_gather_tc(const void* fine, const void* wpack, const float* bias, void* coarse,
                                 int n, int dc, int hc, int wc, int a, int b, void* stream) {
    int rc = k2_tc_common("k2s2_gather_tc", fine, wpack, coarse, n, dc, hc, wc, a, b);
    if (rc) return rc;
    EncodeTiledFn encode = get_encode_fn();
    VS_REQUIRE(encode != nullptr, VS_ERR_CUDA, "k2s2_gather_tc: cuTensorMapEncodeTiled unavailable");
    const int wbytes = g_k2_swizzle ? (4 * b < 128 ? 4 * b : 128) : 16;      // the (kw,b) run of a row, at most 128 bytes
    CUtensorMap map;
    CUresult cr = encode_fine_view(encode, &map, fine, n, dc, hc, wc, b, wbytes / 2,
                                   wbytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : wbytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B :
                                   wbytes == 32 ? CU_TENSOR_MAP_SWIZZLE_32B : CU_TENSOR_MAP_SWIZZLE_NONE);
    VS_REQUIRE(cr == CUDA_SUCCESS, VS_ERR_CUDA, "k2s2_gather_tc: cuTensorMapEncodeTiled failed (%d)", (int)cr);
    K2TcParams p;
    k2_fill_common(p, n, dc, hc, wc, a, b);
    const int nc = k2_gather_nc(a);
    p.nchunks = (a + nc - 1) / nc;
    p.div_chunks = make_fastdiv(p.nchunks);
    p.ks = k2_gather_ks(b);
    p.kstages = 8 * b / p.ks;
    p.wbytes = wbytes;
    p.planes = p.ks * 2 / wbytes;
    const long long items = (long long)n * dc * p.tiles_h * p.tiles_w * p.nchunks;
    VS_REQUIRE(items < 2147483647LL, VS_ERR_SHAPE, "k2s2_gather_tc: too many work items");
    p.work_items = (int)items;
    p.wpack = (const bf16*)wpack; p.bias = bias; p.out = (bf16*)coarse;
    cudaStream_t st = (cudaStream_t)stream;
    if (nc == 16) return launch_k2<16, 4, 2, false>(map, p, st);
    if (nc == 32) return launch_k2<32, 4, 2, false>(map, p, st);
    return launch_k2<64, 4, 2, false>(map, p, st);
}

// fine[n,2dc,2hc,2wc,B] = bias + scatter(coarse[n,dc,hc,wc,A], wpack); bf16 NDHWC in and out.
extern "C" int vs_k2s2_scatter_tc(const void* coarse, const void* wpack, const float* bias, void* fine,
                                  int n, int dc, int hc, int wc, int a, int b, void* stream) {
    int rc = k2_tc_common("k2s2_scatter_tc", coarse, wpack, fine, n, dc, hc, wc, a, b);
    if (rc) return rc;
    EncodeTiledFn encode = get_encode_fn();
    VS_REQUIRE(encode != nullptr, VS_ERR_CUDA, "k2s2_scatter_tc: cuTensorMapEncodeTiled unavailable");
    CUtensorMap map;
    const cuuint64_t gdim[5] = {(cuuint64_t)a, (cuuint64_t)wc, (cuuint64_t)hc, (cuuint64_t)dc, (cuuint64_t)n};
    const cuuint64_t gstr[4] = {(cuuint64_t)a * 2, (cuuint64_t)wc * a * 2, (cuuint64_t)hc * wc * a * 2, (cuuint64_t)dc * hc * wc * a * 2};
    const cuuint32_t box[5] = {8, K2_TW, K2_TH, 1, 1};
    const cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    CUresult cr = encode(&map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(coarse), gdim, gstr, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    VS_REQUIRE(cr == CUDA_SUCCESS, VS_ERR_CUDA, "k2s2_scatter_tc: cuTensorMapEncodeTiled failed (%d)", (int)cr);
    K2TcParams p;
    k2_fill_common(p, n, dc, hc, wc, a, b);
    const int nc = k2_scatter_nc(b);
    p.nchunks = (8 * b + nc - 1) / nc;
    p.div_chunks = make_fastdiv(p.nchunks);
    p.ks = k2_scatter_ks(a);
    p.kstages = (a < 16 ? 16 : a) / p.ks;
    p.wbytes = 16;
    p.planes = a < 16 ? 1 : p.ks / 8;
    const long long items = (long long)n * dc * p.tiles_h * p.tiles_w * p.nchunks;
    VS_REQUIRE(items < 2147483647LL, VS_ERR_SHAPE, "k2s2_scatter_tc: too many work items");
    p.work_items = (int)items;
    p.wpack = (const bf16*)wpack; p.bias = bias; p.out = (bf16*)fine;
    cudaStream_t st = (cudaStream_t)stream;
    if (nc == 64) return launch_k2<64, 3, 4, true>(map, p, st);
    return launch_k2<128, 3, 4, true>(map, p, st);
}
