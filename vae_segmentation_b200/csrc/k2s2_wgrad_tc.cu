// Weight gradient of Conv3d(C,C,2,stride 2) / ConvTranspose3d(C,C,2,stride 2) on the 5th-gen tensor cores (sm_100a).
// Replaces cuDNN backward-filter for joint_model.py:118,130:
//
//     dwt[a][b][k] (+)= sum over coarse voxels o of  coarse[o, a] * fine[2o + k, b]          k = (kd,kh,kw)
//
// GEMM view (same idea as conv3_wgrad_tc.cu): the reduction dimension is the VOXEL, 16 voxels per tcgen05.mma, both
// operands "MN-major" straight from the NDHWC tensors:
//     D[128 rows = (k, b), N = A] += P[(k,b), 16 vox] * C[A, 16 vox]^T
//   * P = the fine tensor through the [2B][Wc][kh][Hc][n*Df] view of k2s2_tc.cu: for a tile of 16 x 8 coarse voxels the
//     8 channels b0..b0+7 of filter position k are ONE TMA box = a plane [128 voxels][16 B]; a core matrix is 8 consecutive
//     w-voxels x 16 B, the two K core matrices of an MMA are two consecutive h-rows (LBO = 128 B), the 16 M groups are
//     16 planes (SBO = plane stride): M = 128 rows = 128 consecutive (k,b) indices.
//   * C = the coarse tile as A/8 planes (N groups, SBO = plane stride).
// The accumulators live in TMEM for the CTA's whole life (persistent over tiles) and are added to dwt once at the end.
// grid = (tile partitions, chunks of (k,b) rows); every fine byte is read once, the small coarse tensor once per row
// chunk.  HBM-bound (8 A B MACs per coarse voxel against 2 (A + 8 B) bytes).
#include "tc_ptx.cuh"

namespace {

constexpr int WG_TH = 16, WG_TW = 8;
constexpr int WG_PLANE = 128 * 16;
constexpr int WG_THREADS = 256;

struct K2WgParams {
    int n, dc, hc, wc, a, b;
    int tiles_h, tiles_w, tiles;        // tiles = n * dc * tiles_h * tiles_w
    int nchunks;                        // chunks of 128 (k,b) rows
    int cpc;                            // chunks per CTA (TMEM / shared-memory budget)
    int npad;                           // MMA N = A rounded up to 16
    int aplanes;                        // coarse planes per tile (A / 8)
    int nstage;
    float* dwt;
};

__host__ __device__ constexpr uint32_t make_idesc_mn128(int n) {
    return (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}

__global__ void __launch_bounds__(WG_THREADS, 1) k2s2_wgrad_tc_kernel(const __grid_constant__ CUtensorMap fmap,
                                                                      const __grid_constant__ CUtensorMap cmap, K2WgParams p) {
    extern __shared__ __align__(128) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~uintptr_t(127));
    const int chunk0 = blockIdx.y * p.cpc;
    const int nck = min(p.cpc, p.nchunks - chunk0);                      // chunks this CTA owns
    const int rows_total = 8 * p.b;
    // planes per chunk actually loaded (B = 8: 64 rows = 8 planes; else 16).  Plane slots that TMA never writes (rows
    // 64..127 for B = 8, the second N group for A = 8) hold garbage: a row / column of D depends only on its own operand
    // row, and those rows / columns are never read back.
    const int ppc = rows_total >= 128 ? 16 : rows_total / 8;
    const int f_bytes = p.cpc * 16 * WG_PLANE;                           // fine planes region of a stage
    const int c_bytes = (p.npad / 8) * WG_PLANE;                         // coarse planes region
    const int stage_bytes = f_bytes + c_bytes;
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + p.nstage * stage_bytes);
    uint64_t* empty_bar = full_bar + 4;
    uint64_t* done_bar = empty_bar + 4;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(done_bar + 1);

    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
    const int lane = threadIdx.x & 31;
    const int tmem_cols = p.cpc * p.npad <= 32 ? 32 : p.cpc * p.npad <= 64 ? 64 : p.cpc * p.npad <= 128 ? 128 : p.cpc * p.npad <= 256 ? 256 : 512;

    if (threadIdx.x == 0) {
        prefetch_tensormap(&fmap);
        prefetch_tensormap(&cmap);
        for (int s = 0; s < p.nstage; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        mbar_init(done_bar, 1);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, (uint32_t)tmem_cols);
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
                int t = tile;
                const int w0 = (t % p.tiles_w) * WG_TW; t /= p.tiles_w;
                const int h0 = (t % p.tiles_h) * WG_TH; t /= p.tiles_h;
                const int d = t % p.dc, n = t / p.dc;
                mbar_wait(&empty_bar[stage], phase ^ 1);
                uint8_t* sf = smem + stage * stage_bytes;
                uint8_t* sc = sf + f_bytes;
                mbar_expect_tx(&full_bar[stage], (uint32_t)((nck * ppc + p.aplanes) * WG_PLANE));
                for (int c = 0; c < nck; ++c) {
                    for (int pl = 0; pl < ppc; ++pl) {
                        const int m0 = (chunk0 + c) * 128 + pl * 8;          // row (k, b0) of this plane
                        const int k = m0 / p.b, b0 = m0 - k * p.b;
                        tma_load_5d(sf + (c * 16 + pl) * WG_PLANE, &fmap, &full_bar[stage], (k & 1) * p.b + b0, w0, (k >> 1) & 1, h0,
                                    n * (2 * p.dc) + 2 * d + (k >> 2));
                    }
                }
                for (int pl = 0; pl < p.aplanes; ++pl)
                    tma_load_5d(sc + pl * WG_PLANE, &cmap, &full_bar[stage], pl * 8, w0, h0, d, n);
                if (++stage == (uint32_t)p.nstage) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        const uint32_t idesc = make_idesc_mn128(p.npad);
        uint32_t stage = 0, phase = 0;
        uint32_t acc = 0u;
        for (int tile = blockIdx.x; tile < p.tiles; tile += gridDim.x) {
            mbar_wait(&full_bar[stage], phase);
            tc_fence_after();
            const uint32_t sf = smem_u32(smem + stage * stage_bytes);
            const uint32_t sc = sf + f_bytes;
#pragma unroll 1
            for (int ks = 0; ks < WG_TH / 2; ++ks) {                      // 16 voxels = two h-rows of the tile
                const uint64_t bd = make_desc(sc + ks * 256u, 128u, WG_PLANE);
#pragma unroll 1
                for (int c = 0; c < nck; ++c) {
                    const uint64_t ad = make_desc(sf + (uint32_t)(c * 16) * WG_PLANE + ks * 256u, 128u, WG_PLANE);
                    tc_mma_elect(tmem_base + c * p.npad, ad, bd, idesc, acc);
                }
                acc = 1u;
            }
            tc_commit_elect(&empty_bar[stage]);
            if (++stage == (uint32_t)p.nstage) { stage = 0; phase ^= 1; }
        }
        tc_commit_elect(done_bar);
    } else if (warp >= 4 && has_work) {
        // ===================== read-back: dwt[a][b][k] += D[(k,b)][a] =====================
        mbar_wait(done_bar, 0);
        tc_fence_after();
        const int q = warp & 3;
        const int r = q * 32 + lane;
        for (int c = 0; c < nck; ++c) {
            const int m = (chunk0 + c) * 128 + r;
            const bool ok = m < rows_total;
            const int k = ok ? m / p.b : 0, bb = ok ? m - k * p.b : 0;
            float* base = p.dwt + (long long)bb * 8 + k;
            for (int c16 = 0; c16 < p.npad / 16; ++c16) {
                uint32_t v[16];
                tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + c * p.npad + c16 * 16, v);
                tmem_ld_wait();
                if (ok) {
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        const int aa = c16 * 16 + i;
                        if (aa < p.a) atomicAdd(base + (long long)aa * p.b * 8, __uint_as_float(v[i]));
                    }
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, (uint32_t)tmem_cols);
    }
}

}  // namespace

// dwt[A][B][8] (+)= sum_o coarse[o,a] * fine[2o+k,b]; bf16 NDHWC operands.  Returns VS_ERR_UNSUPPORTED for shapes the
// tensor-core path does not take (the caller then uses the CUDA-core kernel).
extern "C" int vs_k2s2_wgrad_tc(const void* coarse, const void* fine, float* dwt, int accumulate, int n, int dc, int hc,
                                int wc, int a, int b, void* stream) {
    VS_REQUIRE(coarse && fine && dwt && n > 0 && dc > 0 && hc > 0 && wc > 0, VS_ERR_SHAPE, "k2s2_wgrad_tc: bad arguments");
    VS_REQUIRE(a % 8 == 0 && b % 8 == 0 && a >= 8 && b >= 8 && a <= 256 && b <= 256 && (a == 8 || a % 16 == 0) && (b == 8 || b % 16 == 0),
               VS_ERR_UNSUPPORTED, "k2s2_wgrad_tc: channels must be 8 or a multiple of 16, <= 256 (A=%d B=%d)", a, b);
    VS_REQUIRE(vs_aligned16(coarse) && vs_aligned16(fine), VS_ERR_ALIGN, "k2s2_wgrad_tc: pointers must be 16B aligned");
    EncodeTiledFn encode = get_encode_fn();
    VS_REQUIRE(encode != nullptr, VS_ERR_CUDA, "k2s2_wgrad_tc: cuTensorMapEncodeTiled unavailable");
    cudaStream_t st = (cudaStream_t)stream;
    const cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    CUtensorMap fmap, cmap;
    {
        const cuuint64_t gdim[5] = {(cuuint64_t)2 * b, (cuuint64_t)wc, 2, (cuuint64_t)hc, (cuuint64_t)n * 2 * dc};
        const cuuint64_t rowb = (cuuint64_t)2 * wc * b * 2;
        const cuuint64_t gstr[4] = {(cuuint64_t)2 * b * 2, rowb, 2 * rowb, (cuuint64_t)2 * hc * rowb};
        const cuuint32_t box[5] = {8, WG_TW, 1, WG_TH, 1};
        CUresult cr = encode(&fmap, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(fine), gdim, gstr, box, estr,
                             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        VS_REQUIRE(cr == CUDA_SUCCESS, VS_ERR_CUDA, "k2s2_wgrad_tc: cuTensorMapEncodeTiled (fine) failed (%d)", (int)cr);
    }
    {
        const cuuint64_t gdim[5] = {(cuuint64_t)a, (cuuint64_t)wc, (cuuint64_t)hc, (cuuint64_t)dc, (cuuint64_t)n};
        const cuuint64_t gstr[4] = {(cuuint64_t)a * 2, (cuuint64_t)wc * a * 2, (cuuint64_t)hc * wc * a * 2, (cuuint64_t)dc * hc * wc * a * 2};
        const cuuint32_t box[5] = {8, WG_TW, WG_TH, 1, 1};
        CUresult cr = encode(&cmap, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(coarse), gdim, gstr, box, estr,
                             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        VS_REQUIRE(cr == CUDA_SUCCESS, VS_ERR_CUDA, "k2s2_wgrad_tc: cuTensorMapEncodeTiled (coarse) failed (%d)", (int)cr);
    }
    K2WgParams p;
    p.n = n; p.dc = dc; p.hc = hc; p.wc = wc; p.a = a; p.b = b;
    p.tiles_h = (hc + WG_TH - 1) / WG_TH; p.tiles_w = (wc + WG_TW - 1) / WG_TW;
    const long long tiles = (long long)n * dc * p.tiles_h * p.tiles_w;
    VS_REQUIRE(tiles < 2147483647LL, VS_ERR_SHAPE, "k2s2_wgrad_tc: too many tiles");
    p.tiles = (int)tiles;
    p.nchunks = (8 * b + 127) / 128;
    p.npad = a < 16 ? 16 : a;
    p.aplanes = a / 8;
    p.cpc = (a <= 64 && p.nchunks >= 2) ? 2 : 1;
    const int stage_bytes = p.cpc * 16 * WG_PLANE + (p.npad / 8) * WG_PLANE;
    p.nstage = (200 * 1024) / stage_bytes;
    if (p.nstage > 4) p.nstage = 4;
    VS_REQUIRE(p.nstage >= 2, VS_ERR_UNSUPPORTED, "k2s2_wgrad_tc: stage of %d bytes does not fit twice", stage_bytes);
    p.dwt = dwt;
    if (!accumulate) VS_CUDA(cudaMemsetAsync(dwt, 0, sizeof(float) * 8 * (size_t)a * b, st), "k2s2_wgrad_tc memset");
    const int smem = p.nstage * stage_bytes + 128 + 8 * 9 + 16 + 64;
    static int configured = 0;
    if (configured < smem) {
        VS_CUDA(cudaFuncSetAttribute(k2s2_wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem), "k2s2_wgrad_tc smem attribute");
        configured = smem;
    }
    const int gy = (p.nchunks + p.cpc - 1) / p.cpc;
    // every CTA adds its partial dwt with atomics: keep >= 8 tiles per CTA so that the adds stay a small share
    int gx = vs_sm_count() / gy;
    if (gx < 1) gx = 1;
    if ((long long)gx * 8 > tiles) gx = (int)((tiles + 7) / 8);
    if (gx < 1) gx = 1;
    k2s2_wgrad_tc_kernel<<<dim3((unsigned)gx, (unsigned)gy), WG_THREADS, smem, st>>>(fmap, cmap, p);
    VS_CHECK_LAUNCH("k2s2_wgrad_tc_kernel");
    return VS_OK;
}
