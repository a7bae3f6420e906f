// 3x3x3 convolution (padding 1) as an implicit GEMM on the 5th-gen tensor cores (sm_100a):
//   tcgen05.mma (kind::f16, bf16 x bf16 -> fp32 in TMEM), operands staged by TMA, warp-specialised,
//   persistent over output tiles.  Replaces cuDNN fprop / dgrad for joint_model.py:40-46,106,224,366.
//
// GEMM view:  D[128 voxels, NC couts] += A[128 voxels, 16 cin] * B[NC couts, 16 cin]^T   per filter tap.
//
// A operand = the input halo tile itself.  A CTA owns a 4x16x8 (d,h,w) block of output voxels and
// stages the 6x18x10 halo of 16 input channels in shared memory as two "planes" of 8 channels
// ([halo voxel][8 ch] = 16 B per voxel), each filled by ONE 5-D TMA box load whose out-of-bounds
// zero fill implements the padding.  In the K-major no-swizzle UMMA layout a core matrix is 8 rows x 16 B
// stored contiguously, so 8 consecutive w-voxels of one plane ARE a core matrix; the 16 h-rows of the
// tile are the 16 row groups (SBO = one halo row = 160 B) and the two planes are the two K chunks
// (LBO = plane stride).  A filter tap (kd,kh,kw) is then nothing but a different start address in the
// same halo: every input voxel is fetched from L2/HBM once per CTA and reused by all 27 taps x 4 planes.
// For Cin = 8 two taps are paired into one K=16 MMA: the second K chunk is the first one displaced by LBO (any
// constant number of halo voxels), so the 27 taps in (kd,kh,kw) order pair up as (0,1),(2,3),...,(26,zero): 14 MMAs.
//
// Pipeline: warp 0 = TMA producer, warps 1-4 = MMA issuers (one per d-plane of the tile; warp 1 also owns the
// TMEM allocation), warps 5-8 = epilogue (tcgen05.ld -> shift -> InstanceNorm statistics -> bf16 -> global).
// smem stages ring over (tile, 16-channel slice); TMEM accumulators are double buffered across tiles so the
// epilogue of tile i overlaps the MMAs of tile i+1.
//
// Why four issuing warps: with N = 16..64 one MMA occupies the tensor pipe for only ~40-48 cycles
// (tools/mma_probe.cu: SS-mode floor ~ 32 + N/4 cycles at M = 128), but a single thread cannot issue them
// faster than one per ~75-130 cycles (each UTCHMMA needs its descriptors moved into uniform registers).
// Four warps issuing to four independent accumulators hide that issue latency; all role branches are
// warp-uniform (shfl-derived warp index, elect.sync inside the asm) so no divergent-region waterfall is emitted.
#include "tc_ptx.cuh"
#include "tc_pack.cuh"

namespace {

constexpr int TD = 4, TH = 16, TW = 8;                 // output tile (d,h,w); 128 rows per d-plane
constexpr int HD = TD + 2, HH = TH + 2, HW = TW + 2;   // halo 6 x 18 x 10
constexpr int HV = HD * HH * HW;                       // 1080 halo voxels
constexpr int PLANE_BYTES = HV * 16;                   // 17280 (multiple of 128)
constexpr int PLANE_PAD = 256;                         // slack read by the zero-weighted third tap of the Cin=8 pairs
constexpr int NTHREADS = 288;                          // 1 producer + 4 MMA + 4 epilogue warps

struct TcParams {
    int n, d, h, w, cin, cout;
    int tiles_d, tiles_h, tiles_w, tiles_per_n;
    int bw, bh, bd;         // TMA box = staged halo extents (<= 10 x 18 x 6): clipped to the volume so that small
                            // volumes do not pay for rows that are pure zero fill (TMA cost is per 16-byte row)
    int ref_tile;           // tile (within a sample) that holds the shift's reference voxel
    int td;                 // d-planes per work item (1, 2 or 4): deep levels use fewer so that the grid fills the SMs
    int nsum;               // accumulators per d-plane (K split): with td < 4 the 4 / td issuer warps that share a plane
                            // take the 16-channel K slices round-robin into their own TMEM accumulators and the epilogue
                            // adds them -- a single thread issues one MMA per ~100 cycles, so at the deep levels
                            // (Cin >= 32, one or two planes per CTA) the issue chain, not the tensor pipe, set the time
    int nchunks;            // cout chunks of NC
    int kslices;            // cin / 16 (1 for cin == 8)
    long long work_items;   // n * tiles_per_n * nchunks
    const bf16* wpack;
    bf16* y;
    double* stats;          // [n][cout][2] or null
    float* shift;           // [n][cout] or null; zeroed by the host, published in-kernel (see epilogue)
    // dgrad only -- fused InstanceNorm+ReLU backward reduction of the PREVIOUS layer (whose output gradient this
    // launch produces): psums[n][cout][2] += (sum g*mask, sum g*mask*xhat) with xhat from yprev / pstats.
    const bf16* yprev;      // raw conv output of the previous layer, same shape as y, or null
    const double* pstats;   // its (sum, sumsq) statistics [n][cout][2]
    double* psums;          // zeroed by the caller
    double inv_s;           // 1 / (d*h*w)
    // planar fp32 output of the first two GEMM output channels instead of the bf16 NDHWC store (Cout padded to 8):
    // 1 = the 2-class head: probs = softmax(acc + bias) -> yplanar[n][2][d][h][w]   (joint_model.py:224-225,366-367)
    // 2 = plain values (gradient w.r.t. a 2-channel planar module input)
    int planar_mode;
    float* yplanar;
    const float* bias;      // [2] or null (mode 1)
    // fused InstanceNorm + ReLU (+ skip): a = relu((y - mean) * rstd) + skip written by the same launch after a grid
    // barrier (cooperative launch); needs stats; gbar = zero-initialised counter
    bf16* a_out;
    const bf16* skip;
    unsigned* gbar;
    long long* dbg;         // tools/tc_phase_probe.py: clock64 of CTA 0 at the phases of its first work item, or null
};
#define TC_DBG(slot) do { if (p.dbg != nullptr && blockIdx.x == 0) p.dbg[slot] = clock64(); } while (0)

// Column sums of a 32-lane x 16-value register tile in 31 shuffles: afterwards v[0] of lane L holds the
// sum over all lanes of column (L & 15).
__device__ __forceinline__ float transpose_reduce16(float (&v)[16], int lane) {
#pragma unroll
    for (int k = 0; k < 16; ++k) v[k] += __shfl_xor_sync(0xffffffffu, v[k], 16);
#pragma unroll
    for (int h = 8; h >= 1; h >>= 1) {
        const bool up = (lane & h) != 0;
#pragma unroll
        for (int k = 0; k < h; ++k) {
            const float send = up ? v[k] : v[k + h];
            const float keep = up ? v[k + h] : v[k];
            v[k] = keep + __shfl_xor_sync(0xffffffffu, send, h);
        }
    }
    return v[0];
}

// Work item -> (sample, tile origin, cout chunk).  The first n * nchunks items are the tiles that hold the reference
// voxel of the shift (tile index p.ref_tile of every sample): they are the first items of the lowest-numbered CTAs,
// which the hardware dispatches first and which never wait, so a CTA spinning on a shift value always waits on a CTA
// that is already running -- whatever else shares the GPU.  The remaining items follow in (n, tile, chunk) order.
// 32-bit arithmetic (the host rejects launches with >= 2^31 work items): every role of every tile decodes its work
// item, and 64-bit divisions by run-time values cost hundreds of cycles on the single-warp epilogue / issue chains.
__device__ __forceinline__ void decode_work(long long item64, const TcParams& p, int& n, int& d0, int& h0, int& w0, int& chunk) {
    const unsigned item = (unsigned)item64;
    const unsigned nprod = (unsigned)(p.n * p.nchunks);
    int r;
    if (item < nprod) {
        chunk = (int)(item % (unsigned)p.nchunks);
        n = (int)(item / (unsigned)p.nchunks);
        r = p.ref_tile;
    } else {
        const unsigned idx = item - nprod;
        chunk = (int)(idx % (unsigned)p.nchunks);
        const unsigned t = idx / (unsigned)p.nchunks;
        n = (int)(t / (unsigned)(p.tiles_per_n - 1));
        r = (int)(t % (unsigned)(p.tiles_per_n - 1));
        r += (r >= p.ref_tile);
    }
    w0 = (r % p.tiles_w) * TW; r /= p.tiles_w;
    h0 = (r % p.tiles_h) * TH; r /= p.tiles_h;
    d0 = r * p.td;
}

// ---------------------------------------------------------------------------------------------
// NC: output channels per CTA (MMA N); CIN8: single 8-channel plane with paired taps.
// Shared memory: NSTAGE x { A planes | B taps } + barriers; sized by the host.
// ---------------------------------------------------------------------------------------------
// H8 (NC = 16 only): the layer has 8 output channels -- the epilogue touches accumulator columns 0..7 only (the MMA
// still runs at N = 16, its minimum at M = 128; columns 8..15 hold the zero-padded weights' zeros).
// FEAT: compile-time feature set (planar epilogue, fused InstanceNorm pass, fused norm-backward reduction): the rarely used
// paths get their own instantiations so the default body stays small (see conv3_tc_kdn.cu: compiled into one body they
// cost the hot kernel 30 % through registers and instruction-cache footprint).
constexpr int TC_PLANAR = 1, TC_APPLY = 2, TC_REDUCE = 8;
template <int NC, bool CIN8, int NSTAGE, bool H8, int FEAT>
__global__ void __launch_bounds__(NTHREADS, 1) conv3_tc_kernel(const __grid_constant__ CUtensorMap xmap, TcParams p) {
    static_assert(!H8 || NC == 16, "H8 is a variant of the 16-column kernel");
    constexpr int NV = H8 ? 8 : 16;                           // accumulator columns the epilogue processes per 16-column group
    constexpr int NH8 = H8 ? 1 : 2;
    constexpr int A_BYTES = CIN8 ? (PLANE_BYTES + PLANE_PAD) : 2 * PLANE_BYTES;
    constexpr int NMMA = CIN8 ? 14 : 27;                      // MMAs per d-plane per stage
    constexpr int B_BYTES = NMMA * NC * 32;                   // [mma][kc 2][NC/8][8 rows][16 B]
    constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    constexpr int NBUF = 2;
    constexpr int TMEM_COLS = NBUF * TD * NC;                 // 128 / 256 / 512
    static_assert(STAGE_BYTES % 128 == 0, "stage alignment");

    extern __shared__ __align__(128) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~uintptr_t(127));
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + NSTAGE * STAGE_BYTES);
    uint64_t* empty_bar = full_bar + NSTAGE;
    uint64_t* tfull_bar = empty_bar + NSTAGE;
    uint64_t* tempty_bar = tfull_bar + NBUF;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + NBUF);
    float* sshift = reinterpret_cast<float*>(tmem_slot + 4);          // [NC]
    float* smean = sshift + NC;                                       // [NC]  (fused norm-backward reduction)
    float* srstd = smean + NC;                                        // [NC]

    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);      // warp-uniform for the compiler
    const int lane = threadIdx.x & 31;
    if (threadIdx.x == 0) TC_DBG(0);

    if (threadIdx.x == 0) {
        prefetch_tensormap(&xmap);
        for (int s = 0; s < NSTAGE; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], TD); }
        for (int b = 0; b < NBUF; ++b) { mbar_init(&tfull_bar[b], TD); mbar_init(&tempty_bar[b], 4); }
        fence_barrier_init();
    }
    if (CIN8) {
        // The zero-weighted third tap of the Cin=8 tap pairs reads one voxel past the row it belongs to: past the
        // end of the staged box that is memory TMA never writes, which must not hold NaN patterns.  Zero the pad
        // behind the full-size plane, or the whole A region when the box is clipped.
        const bool clipped = p.bw < HW || p.bh < HH || p.bd < HD;
        const int words = clipped ? A_BYTES / 4 : PLANE_PAD / 4, off = clipped ? 0 : PLANE_BYTES;
        for (int i = threadIdx.x; i < NSTAGE * words; i += NTHREADS) {
            int s = i / words, k = i % words;
            reinterpret_cast<uint32_t*>(smem + s * STAGE_BYTES + off)[k] = 0u;
        }
        fence_proxy_async();
    }
    if (warp == 1) tmem_alloc(tmem_slot, TMEM_COLS);
    pdl_trigger();                       // the next kernel of the chain may set itself up behind this one
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);
    pdl_wait();                          // everything above overlapped the predecessor; global memory from here on
    if (threadIdx.x == 0) TC_DBG(1);

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            uint32_t stage = 0, phase = 0;
            for (long long item = blockIdx.x; item < p.work_items; item += gridDim.x) {
                int n, d0, h0, w0, chunk;
                decode_work(item, p, n, d0, h0, w0, chunk);
                for (int ks = 0; ks < p.kslices; ++ks) {
                    mbar_wait(&empty_bar[stage], phase ^ 1);
                    uint8_t* sa = smem + stage * STAGE_BYTES;
                    mbar_expect_tx(&full_bar[stage], (uint32_t)((CIN8 ? 1 : 2) * p.bd * p.bh * p.bw * 16 + B_BYTES));
                    if (CIN8) tma_load_4d(sa, &xmap, &full_bar[stage], (w0 - 1) * 8, h0 - 1, d0 - 1, n);   // (w,c) merged: 160 B rows
                    else tma_load_5d(sa, &xmap, &full_bar[stage], ks * 16, w0 - 1, h0 - 1, d0 - 1, n);
                    if (!CIN8) tma_load_5d(sa + PLANE_BYTES, &xmap, &full_bar[stage], ks * 16 + 8, w0 - 1, h0 - 1, d0 - 1, n);
                    const uint8_t* wsrc = reinterpret_cast<const uint8_t*>(p.wpack) +
                                          ((long long)chunk * p.kslices + ks) * B_BYTES;
                    bulk_load(sa + A_BYTES, wsrc, B_BYTES, &full_bar[stage]);
                    if (item == blockIdx.x && ks == 0) TC_DBG(2);
                    if (item == blockIdx.x && ks == p.kslices - 1) TC_DBG(3);
                    if (++stage == NSTAGE) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp <= TD) {
        // ===================== MMA issuers: warp 1 + j owns d-plane j % td of every tile (and, when several warps
        // share a plane, the K slices ks with ks % nsplit == j / td, accumulated in its own TMEM accumulator j) ========
        const int j = warp - 1;
        const int pl = j % p.td, part = j / p.td, nsplit = TD / p.td;
        constexpr uint32_t idesc = make_idesc(NC);
        uint32_t stage = 0, phase = 0, buf = 0, bphase = 0;
        for (long long item = blockIdx.x; item < p.work_items; item += gridDim.x) {
            int n, d0, h0, w0, chunk;
            decode_work(item, p, n, d0, h0, w0, chunk);
            const bool active = pl < min(p.td, p.d - d0) && part < p.nsum;
            mbar_wait(&tempty_bar[buf], bphase ^ 1);
            tc_fence_after();
            const uint32_t dcol = tmem_base + (buf * TD + j) * NC;
            for (int ks = 0; ks < p.kslices; ++ks) {
                mbar_wait(&full_bar[stage], phase);
                tc_fence_after();
                if (j == 0 && lane == 0 && item == blockIdx.x && ks == 0) TC_DBG(4);
                const int ksl = p.nsum > 1 ? ks - part : ks;     // >= 0 and a multiple of nsplit on this warp's slices; 0 on its first
                if (active && (p.nsum == 1 || (ks % nsplit) == part)) {
                    const uint32_t a_base = smem_u32(smem + stage * STAGE_BYTES) + (uint32_t)(pl * p.bh * p.bw) * 16u;
                    const uint32_t b_base = smem_u32(smem + stage * STAGE_BYTES) + A_BYTES;
                    if (CIN8) {
                        // Cin = 8: the two K chunks of an MMA are two filter taps.  Any two taps work -- the second
                        // chunk is just the first one displaced by a constant number of halo voxels (LBO) -- so the 27
                        // taps in (kd,kh,kw) order are paired (0,1),(2,3),...,(24,25),(26,zero weights): 14 MMAs.
#pragma unroll
                        for (int m = 0; m < 14; ++m) {
                            const int t1 = 2 * m, t2 = (2 * m + 1 < 27) ? 2 * m + 1 : 26;
                            const int o1 = ((t1 / 9) * p.bh + (t1 / 3) % 3) * p.bw + t1 % 3;
                            const int o2 = ((t2 / 9) * p.bh + (t2 / 3) % 3) * p.bw + t2 % 3;
                            const uint32_t lbo = (2 * m + 1 < 27) ? (uint32_t)(o2 - o1) * 16u : 16u;
                            const uint64_t ad = make_desc(a_base + (uint32_t)o1 * 16u, lbo, (uint32_t)p.bw * 16u);
                            const uint64_t bd = make_desc(b_base + m * (NC * 32), NC * 16, 128u);
                            tc_mma_elect(dcol, ad, bd, idesc, (ksl | m) != 0);
                        }
                    } else {
                    int m = 0;
#pragma unroll 1
                    for (int kd = 0; kd < 3; ++kd) {
#pragma unroll
                        for (int kh = 0; kh < 3; ++kh) {
                            const uint32_t row = a_base + (uint32_t)((kd * p.bh + kh) * p.bw) * 16u;
#pragma unroll
                            for (int kx = 0; kx < 3; ++kx, ++m) {
                                const uint64_t ad = make_desc(row + kx * 16u, PLANE_BYTES, (uint32_t)p.bw * 16u);
                                const uint64_t bd = make_desc(b_base + m * (NC * 32), NC * 16, 128u);
                                tc_mma_elect(dcol, ad, bd, idesc, (ksl | m) != 0);
                            }
                        }
                    }
                    }
                }
                tc_commit_elect(&empty_bar[stage]);             // this warp's reads of the smem stage have retired
                if (++stage == NSTAGE) { stage = 0; phase ^= 1; }
            }
            tc_commit_elect(&tfull_bar[buf]);                   // this warp's accumulator plane is complete
            if (j == 0 && lane == 0 && item == blockIdx.x) TC_DBG(5);
            if (++buf == NBUF) { buf = 0; bphase ^= 1; }
        }
    } else {
        // ===================== epilogue (warps 5..8) =====================
        const int q = warp & 3;                                  // TMEM lane quadrant this warp may read
        const int et = threadIdx.x - 32 * (TD + 1);              // 0..127 within the epilogue group
        const int row = q * 32 + lane;                           // accumulator row = voxel in the d-plane
        const int lh = row >> 3, lw = row & 7;
        uint32_t buf = 0, bphase = 0;
        int stat_n = -1, stat_chunk = -1;
        // running statistics: every thread keeps fp32 (sum, sum of squares) of ITS accumulator row for each of the
        // NC columns across all tiles of a (n, chunk) group -- two FP ops per element in the hot loop; the cross-lane
        // reduction (31 shuffles per 16 columns) and the fp64 atomics run once per group, not once per tile.  The
        // values are deviations from the shift, so fp32 partial sums of a few hundred terms lose nothing.
        float rs[NC / 16][16], rq[NC / 16][16];
#pragma unroll
        for (int c = 0; c < NC / 16; ++c)
#pragma unroll
            for (int k = 0; k < 16; ++k) { rs[c][k] = 0.f; rq[c][k] = 0.f; }
        const bool fused = (FEAT & TC_REDUCE) && p.psums != nullptr;
        double* const sout = fused ? p.psums : p.stats;
        auto flush_stats = [&]() {
            if (sout != nullptr && stat_n >= 0) {
#pragma unroll
                for (int c = 0; c < NC / 16; ++c) {
                    const float s1 = transpose_reduce16(rs[c], lane);
                    const float s2 = transpose_reduce16(rq[c], lane);
                    const int co = stat_chunk * NC + c * 16 + lane;
                    if (lane < 16 && co < p.cout) {
                        atomicAdd(&sout[((long long)stat_n * p.cout + co) * 2], (double)s1);
                        atomicAdd(&sout[((long long)stat_n * p.cout + co) * 2 + 1], (double)s2);
                    }
#pragma unroll
                    for (int k = 0; k < 16; ++k) { rs[c][k] = 0.f; rq[c][k] = 0.f; }
                }
            }
        };
        if (et < NC) sshift[et] = 0.f;
        asm volatile("bar.sync 1, 128;" ::: "memory");
        // The per-(n,c) shift (see header: InstanceNorm-invariant, keeps bf16 precision when |mean| >> sigma) is the
        // fp32 accumulator of a reference voxel of the sample.  It is produced by the CTA that owns the tile holding
        // that voxel (tile 0 of the sample = the lowest work item of its (n, chunk) group), published through global
        // memory as (bits | 1) into the host-zeroed `shift` array, and every other CTA spins on "non-zero" before its
        // first tile of the group.  Progress: the producing tiles are the FIRST items of CTAs 0 .. n*nchunks-1
        // (decode_work), a producing item never waits, and the hardware dispatches CTAs in index order -- so whenever
        // a waiter is resident, the CTA it waits on already is, independent of what else occupies the GPU.
        float hb0 = 0.f, hb1 = 0.f;
        if ((FEAT & TC_PLANAR) && p.planar_mode == 1 && p.bias != nullptr) { hb0 = p.bias[0]; hb1 = p.bias[1]; }
        const long long vol = (long long)p.d * p.h * p.w;
        const int rd = min(1, p.d - 1), rh = min(1, p.h - 1), rw = min(1, p.w - 1);
        const bool has_shift = p.shift != nullptr;
        const bool has_stats = p.stats != nullptr;
        const uint32_t sshift_addr = smem_u32(sshift);
        const uint32_t smean_addr = smem_u32(smean), srstd_addr = smem_u32(srstd);
        const int ref_row = rh * TW + rw;
        const int ref_d0 = (rd / p.td) * p.td;                   // first plane of the work item that owns the reference voxel
        float shr[NC == 16 ? 16 : 1];
#pragma unroll
        for (int k = 0; k < (NC == 16 ? 16 : 1); ++k) shr[k] = 0.f;
        for (long long item = blockIdx.x; item < p.work_items; item += gridDim.x) {
            int n, d0, h0, w0, chunk;
            decode_work(item, p, n, d0, h0, w0, chunk);
            const int co0 = chunk * NC;
            if (n != stat_n || chunk != stat_chunk) {
                flush_stats();
                stat_n = n; stat_chunk = chunk;
                if (fused) {
                    asm volatile("bar.sync 1, 128;" ::: "memory");       // everyone is done with the previous group's constants
                    if (et < NC) {
                        float mu = 0.f, rs_ = 0.f;
                        if (co0 + et < p.cout) in_mean_rstd(p.pstats + ((long long)n * p.cout + co0 + et) * 2, p.inv_s, mu, rs_);
                        smean[et] = mu; srstd[et] = rs_;
                    }
                    asm volatile("bar.sync 1, 128;" ::: "memory");
                }
                if (p.shift != nullptr) {
                    asm volatile("bar.sync 1, 128;" ::: "memory");       // everyone is done with the previous group's sshift
                    if (d0 == ref_d0 && h0 == 0 && w0 == 0) {
                        mbar_wait(&tfull_bar[buf], bphase);
                        tc_fence_after();
                        if (q == (ref_row >> 5)) {
#pragma unroll
                            for (int c16 = 0; c16 < NC / 16; ++c16) {
                                uint32_t r[16];
                                tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (buf * TD + (rd - ref_d0)) * NC + c16 * 16, r);
                                tmem_ld_wait();
                                for (int part = 1; part < p.nsum; ++part) {          // K-split accumulators of the plane
                                    uint32_t r2[16];
                                    tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (buf * TD + (rd - ref_d0) + part * p.td) * NC + c16 * 16, r2);
                                    tmem_ld_wait();
#pragma unroll
                                    for (int k = 0; k < 16; ++k) r[k] = __float_as_uint(__uint_as_float(r[k]) + __uint_as_float(r2[k]));
                                }
                                if (lane == (ref_row & 31)) {
#pragma unroll
                                    for (int k = 0; k < 16; ++k) {
                                        const uint32_t bits = r[k] | 1u;
                                        sshift[c16 * 16 + k] = __uint_as_float(bits);
                                        if (co0 + c16 * 16 + k < p.cout)
                                            *reinterpret_cast<volatile uint32_t*>(p.shift + (long long)n * p.cout + co0 + c16 * 16 + k) = bits;
                                    }
                                }
                            }
                        }
                    } else if (et < NC) {
                        float sv = 0.f;
                        if (co0 + et < p.cout) {
                            const volatile uint32_t* flag = reinterpret_cast<const volatile uint32_t*>(p.shift + (long long)n * p.cout + co0 + et);
                            uint32_t bits, spins = 0;
                            while ((bits = *flag) == 0u) {
                                __nanosleep(64);
                                if (++spins > (1u << 23)) __trap();                // never hang the GPU on a protocol bug
                            }
                            sv = __uint_as_float(bits);
                        }
                        sshift[et] = sv;
                    }
                    asm volatile("bar.sync 1, 128;" ::: "memory");
                    if (NC == 16) {
#pragma unroll
                        for (int k4 = 0; k4 < 4; ++k4) {
                            const float4 sh = lds128(sshift_addr + k4 * 16);
                            shr[NC == 16 ? k4 * 4 + 0 : 0] = sh.x; shr[NC == 16 ? k4 * 4 + 1 : 0] = sh.y;
                            shr[NC == 16 ? k4 * 4 + 2 : 0] = sh.z; shr[NC == 16 ? k4 * 4 + 3 : 0] = sh.w;
                        }
                    }
                }
            }
            const int jmax = min(p.td, p.d - d0);
            const int gh = h0 + lh, gw = w0 + lw;
            const bool rc_ok = gh < p.h && gw < p.w;
            // fused norm-backward reduction: the previous layer's raw output at this thread's voxels is fetched BEFORE
            // waiting for the accumulators, so the global latency hides behind the MMAs of the tile (NC = 16 only: at
            // NC >= 32 the registers are needed for the running sums and the loads stay inline)
            uint4 ypre[NC == 16 ? TD : 1][2];
            if (NC == 16 && fused) {
#pragma unroll
                for (int j = 0; j < TD; ++j) {
#pragma unroll
                    for (int h8 = 0; h8 < NH8; ++h8) {
                        ypre[NC == 16 ? j : 0][h8] = make_uint4(0u, 0u, 0u, 0u);
                        if (j < jmax && rc_ok && co0 + h8 * 8 < p.cout)
                            ypre[NC == 16 ? j : 0][h8] = __ldg(reinterpret_cast<const uint4*>(
                                p.yprev + (((long long)n * p.d + d0 + j) * p.h + gh) * (long long)p.w * p.cout + (long long)gw * p.cout + co0 + h8 * 8));
                    }
                }
            }
            if (et == 0 && item == blockIdx.x) TC_DBG(6);
            mbar_wait(&tfull_bar[buf], bphase);
            tc_fence_after();
            if (et == 0 && item == blockIdx.x) TC_DBG(7);
            // The epilogue is ONE warp per scheduler running a dependent chain: its latencies are not hidden by other
            // warps, and at the full-resolution levels it -- not the MMAs -- sets the tile time.  NC = 16: the TMEM
            // load of plane j+1 is in flight while plane j is processed, the shift lives in registers.
            uint32_t rpipe[NC == 16 ? 2 : 1][NV];
            if constexpr (NC == 16) {
                if constexpr (H8) tmem_ld8(tmem_base + ((uint32_t)(q * 32) << 16) + (buf * TD) * NC, rpipe[0]);
                else tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (buf * TD) * NC, rpipe[0]);
            }
#pragma unroll
            for (int j = 0; j < TD; ++j) {
                if (j >= jmax) break;
                const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (buf * TD + j) * NC;
                if constexpr (NC == 16) {
                    tmem_ld_wait();
                    if (j + 1 < jmax) {
                        if constexpr (H8) tmem_ld8(taddr + NC, rpipe[(j + 1) & 1]);
                        else tmem_ld16(taddr + NC, rpipe[(j + 1) & 1]);
                    }
                }
                const long long yoff = (((long long)n * p.d + d0 + j) * p.h + gh) * (long long)p.w * p.cout + (long long)gw * p.cout + co0;
                bf16* py = p.y + yoff;
#pragma unroll
                for (int c16 = 0; c16 < NC / 16; ++c16) {
                    uint32_t r[NV];
                    if constexpr (NC == 16) {
#pragma unroll
                        for (int k = 0; k < NV; ++k) r[k] = rpipe[j & 1][k];
                    } else {
                        tmem_ld16(taddr + c16 * 16, r);
                        tmem_ld_wait();
                    }
                    for (int part = 1; part < p.nsum; ++part) {                  // K-split accumulators of the plane
                        uint32_t r2[NV];
                        if constexpr (H8) tmem_ld8(taddr + part * p.td * NC + c16 * 16, r2);
                        else tmem_ld16(taddr + part * p.td * NC + c16 * 16, r2);
                        tmem_ld_wait();
#pragma unroll
                        for (int k = 0; k < NV; ++k) r[k] = __float_as_uint(__uint_as_float(r[k]) + __uint_as_float(r2[k]));
                    }
                    float v[NV];
                    if (has_shift && NC == 16) {
#pragma unroll
                        for (int k = 0; k < NV; ++k) v[k] = __uint_as_float(r[k]) - shr[k];
                    } else if (has_shift) {
                        // explicit 128-bit shared loads (warp-broadcast, conflict-free); a generic LD through the
                        // reinterpret-cast pointer costs ~8 wavefronts each on the pipe the MMA operands share
#pragma unroll
                        for (int k4 = 0; k4 < NV / 4; ++k4) {
                            const float4 sh = lds128(sshift_addr + (c16 * 16 + k4 * 4) * 4);
                            v[k4 * 4 + 0] = __uint_as_float(r[k4 * 4 + 0]) - sh.x;
                            v[k4 * 4 + 1] = __uint_as_float(r[k4 * 4 + 1]) - sh.y;
                            v[k4 * 4 + 2] = __uint_as_float(r[k4 * 4 + 2]) - sh.z;
                            v[k4 * 4 + 3] = __uint_as_float(r[k4 * 4 + 3]) - sh.w;
                        }
                    } else {
#pragma unroll
                        for (int k = 0; k < NV; ++k) v[k] = __uint_as_float(r[k]);
                    }
                    if (!rc_ok) {
#pragma unroll
                        for (int k = 0; k < NV; ++k) v[k] = 0.f;
                    }
                    if ((FEAT & TC_PLANAR) && p.planar_mode != 0) {
                        if (rc_ok && c16 == 0 && co0 == 0) {
                            float o0 = v[0] + hb0, o1 = v[1] + hb1;
                            if (p.planar_mode == 1) {
                                const float mx = fmaxf(o0, o1);
                                const float e0 = expf(o0 - mx), e1 = expf(o1 - mx);
                                const float inv = 1.f / (e0 + e1);
                                o0 = e0 * inv; o1 = e1 * inv;
                            }
                            float* pp = p.yplanar + (long long)n * 2 * vol + ((long long)(d0 + j) * p.h + gh) * p.w + gw;
                            pp[0] = o0;
                            pp[vol] = o1;
                        }
                    } else if (rc_ok) {
#pragma unroll
                        for (int h8 = 0; h8 < NH8; ++h8) {
                            if (co0 + c16 * 16 + h8 * 8 < p.cout) {
                                float o[8];
#pragma unroll
                                for (int k = 0; k < 8; ++k) o[k] = v[h8 * 8 + k];
                                Store<bf16>::st8(py + c16 * 16 + h8 * 8, o);
                            }
                        }
                    }
                    if (has_stats) {
#pragma unroll
                        for (int k = 0; k < NV; ++k) { rs[c16][k] += v[k]; rq[c16][k] = fmaf(v[k], v[k], rq[c16][k]); }
                    }
                    if (fused) {
                        // g = this launch's output (as stored, i.e. rounded to bf16); mask / xhat from the previous
                        // layer's raw output at the same voxel -- exactly what inorm_relu_bwd_reduce_kernel computes
#pragma unroll
                        for (int h8 = 0; h8 < NH8; ++h8) {
                            float yv[8];
                            if (NC == 16) {
                                const uint4 raw = ypre[NC == 16 ? j : 0][h8];
                                const uint32_t u[4] = {raw.x, raw.y, raw.z, raw.w};
#pragma unroll
                                for (int i2 = 0; i2 < 4; ++i2) { yv[2 * i2] = __uint_as_float(u[i2] << 16); yv[2 * i2 + 1] = __uint_as_float(u[i2] & 0xffff0000u); }
                            } else {
#pragma unroll
                                for (int k = 0; k < 8; ++k) yv[k] = 0.f;
                                if (rc_ok && co0 + c16 * 16 + h8 * 8 < p.cout) Store<bf16>::ld8(p.yprev + yoff + c16 * 16 + h8 * 8, yv);
                            }
#pragma unroll
                            for (int k4 = 0; k4 < 2; ++k4) {
                                const float4 mu = lds128(smean_addr + (c16 * 16 + h8 * 8 + k4 * 4) * 4);
                                const float4 rr = lds128(srstd_addr + (c16 * 16 + h8 * 8 + k4 * 4) * 4);
                                const float mus[4] = {mu.x, mu.y, mu.z, mu.w}, rrs[4] = {rr.x, rr.y, rr.z, rr.w};
#pragma unroll
                                for (int k = 0; k < 4; ++k) {
                                    const int kk = h8 * 8 + k4 * 4 + k;
                                    const float xh = (yv[k4 * 4 + k] - mus[k]) * rrs[k];
                                    const float gq = __bfloat162float(__float2bfloat16_rn(v[kk]));
                                    const float gm = xh > 0.f ? gq : 0.f;
                                    rs[c16][kk] += gm;
                                    rq[c16][kk] = fmaf(gm, xh, rq[c16][kk]);
                                }
                            }
                        }
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty_bar[buf]);
            if (et == 0 && item == blockIdx.x) TC_DBG(8);
            if (++buf == NBUF) { buf = 0; bphase ^= 1; }
        }
        flush_stats();
        if (et == 0) TC_DBG(9);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, TMEM_COLS);
    }
    if ((FEAT & TC_APPLY) && p.a_out != nullptr) {
        // ---- fused InstanceNorm + ReLU (+ skip); see conv3_tc_kdn.cu ---------------------------------------------------
        __threadfence();
        __syncthreads();
        if (threadIdx.x == 0) vs_grid_barrier(p.gbar);
        __syncthreads();
        float* tab = reinterpret_cast<float*>(smem);                          // [n * cout][2] = (mean, rstd)
        for (int i = threadIdx.x; i < p.n * p.cout; i += NTHREADS) {
            float m1, r1;
            in_mean_rstd_cg(p.stats + (long long)i * 2, p.inv_s, m1, r1);
            tab[2 * i] = m1; tab[2 * i + 1] = r1;
        }
        __syncthreads();
        constexpr int NW = NTHREADS / 32;
        int it = 0;
        for (long long item = blockIdx.x; item < p.work_items; item += gridDim.x, ++it) {
            if (it % NW != warp) continue;
            int n, d0, h0, w0, chunk;
            decode_work(item, p, n, d0, h0, w0, chunk);
            in_relu_apply_tile<NC / 8, 4>(p.y, p.skip, p.a_out, tab, lane, n, d0, h0, w0, chunk * NC, min(p.td, p.d - d0),
                                          p.d, p.h, p.w, p.cout);
        }
    }
    if (threadIdx.x == 0) TC_DBG(10);
}

// ---------------------------------------------------------------------------------------------
// weight packing: fp32 master [Cout][Cin][27] -> bf16 UMMA B operand blocks
//   general: [chunk][kslice][tap 27][kc 2][NC/8][8 rows][8 ch]
//   cin8   : [chunk][1][pair 14][kc 2][NC/8][8 rows][8 ch]  (kc0 = tap 2*pair, kc1 = tap 2*pair+1 in (kd,kh,kw) order; the 28th is zero)
// dgrad = same contraction with (ci,co) swapped and taps flipped.
// ---------------------------------------------------------------------------------------------
__global__ void pack_tc_kernel(const float* __restrict__ w, bf16* __restrict__ out, int cin_l, int cout_l, int dgrad,
                               int nc, int cin8, long long total) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x)
        out[i] = __float2bfloat16_rn(pack_tc_elem(w, i, cin_l, cout_l, cout_l, dgrad, nc, cin8));
}

int nc_for(int gout) { return nc_for_dev(gout); }
long long* g_tc_dbg = nullptr;   // vs_debug_set_tc_phase_buffer: device buffer of >= 16 clock64 slots (tools/tc_phase_probe.py)
int g_conv3_ksplit = 1;          // A/B switch (vs_debug_set_conv3_ksplit): K split over the issuer warps that share a plane

template <int NC, bool CIN8, int NSTAGE, bool H8, int FEAT>
int launch_tc_feat(const CUtensorMap& map, const TcParams& p, cudaStream_t st) {
    constexpr int A_BYTES = CIN8 ? (PLANE_BYTES + PLANE_PAD) : 2 * PLANE_BYTES;
    constexpr int B_BYTES = (CIN8 ? 14 : 27) * NC * 32;
    constexpr int SMEM_NEED = NSTAGE * (A_BYTES + B_BYTES) + 128 /*align*/ + 8 * (2 * NSTAGE + 4) + 16 + NC * 4 + NC * 16 + 64;
    // one CTA per SM whatever the register count (see conv3_tc_kdn.cu: a second resident CTA that blocks in tcgen05.alloc
    // can dead-lock the shift hand-off)
    constexpr int SMEM = SMEM_NEED > 120 * 1024 ? SMEM_NEED : 120 * 1024;
    static_assert(SMEM <= 227 * 1024, "shared memory budget");
    auto kern = conv3_tc_kernel<NC, CIN8, NSTAGE, H8, FEAT>;
    static bool configured = false;
    if (!configured) {
        VS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM), "conv3_tc smem attribute");
        configured = true;
    }
    const long long grid = p.work_items < (long long)vs_sm_count() ? p.work_items : (long long)vs_sm_count();
    if (p.a_out != nullptr)
        VS_CUDA(vs_launch_coop(kern, dim3((unsigned)grid), dim3(NTHREADS), SMEM, st, map, p), "conv3_tc_kernel cooperative launch");
    else
        VS_CUDA(vs_launch(kern, dim3((unsigned)grid), dim3(NTHREADS), SMEM, st, map, p), "conv3_tc_kernel launch");
    VS_CHECK_LAUNCH("conv3_tc_kernel");
    return VS_OK;
}

// picks the instantiation for the features this launch needs (see TC_* above)
template <int NC, bool CIN8, int NSTAGE, bool H8 = false>
int launch_tc(const CUtensorMap& map, const TcParams& p, cudaStream_t st) {
    if (p.a_out != nullptr) return launch_tc_feat<NC, CIN8, NSTAGE, H8, TC_APPLY>(map, p, st);
    if (p.planar_mode != 0) {
        if constexpr (H8) return launch_tc_feat<NC, CIN8, NSTAGE, H8, TC_PLANAR>(map, p, st);      // planar output: Cout padded to 8
        else VS_FAIL(VS_ERR_UNSUPPORTED, "conv3_tc: planar output needs Cout padded to 8");
    }
    if (p.psums != nullptr) return launch_tc_feat<NC, CIN8, NSTAGE, H8, TC_REDUCE>(map, p, st);
    return launch_tc_feat<NC, CIN8, NSTAGE, H8, 0>(map, p, st);
}

}  // namespace

// bytes of the bf16 tensor-core weight pack for a layer (fprop: dgrad=0, dgrad: dgrad=1); 0 = unsupported shape
extern "C" size_t vs_conv3_tc_pack_bytes(int cin, int cout, int dgrad) {
    const int gin = dgrad ? cout : cin, gout = dgrad ? cin : cout;
    if (!(gin == 8 || (gin % 16 == 0 && gin >= 16)) || gout % 8 != 0 || gout < 8) return 0;
    const int nc = nc_for(gout);
    const int nchunks = (gout + nc - 1) / nc;
    const int kslices = gin == 8 ? 1 : gin / 16;
    return (size_t)nchunks * kslices * (gin == 8 ? 14 : 27) * nc * 32;
}

// Padded variants: the fp32 master weight is [cout][cin][27]; the pack is built for cout_pad >= cout output and
// cin_pad >= cin input channels (zeros), e.g. the 2-class head as an 8-channel layer (fprop and dgrad) and the VAE
// in-block's 2-channel input gradient as an 8-channel one.
extern "C" size_t vs_conv3_tc_pack_bytes(int cin, int cout, int dgrad);
__global__ void pack_tc_padded_kernel(const float* __restrict__ w, bf16* __restrict__ out, int cin, int cout, int cin_pad,
                                      int cout_pad, int dgrad, int nc, int cin8, long long total) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x)
        out[i] = __float2bfloat16_rn(pack_tc_elem(w, i, cin_pad, cout_pad, cout, dgrad, nc, cin8, cin));
}
extern "C" int vs_pack_conv3_weight_tc_padded(const float* w, void* out, int cin, int cout, int cin_pad, int cout_pad, int dgrad,
                                              void* stream) {
    const size_t bytes = vs_conv3_tc_pack_bytes(cin_pad, cout_pad, dgrad);
    VS_REQUIRE(w && out && bytes > 0 && cin_pad >= cin && cout_pad >= cout, VS_ERR_UNSUPPORTED,
               "pack_conv3_weight_tc_padded: unsupported shape Cin=%d(%d) Cout=%d(%d)", cin, cin_pad, cout, cout_pad);
    const int gin = dgrad ? cout_pad : cin_pad, gout = dgrad ? cin_pad : cout_pad;
    const long long total = (long long)(bytes / 2);
    pack_tc_padded_kernel<<<(unsigned)min(1024LL, (total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        w, (bf16*)out, cin, cout, cin_pad, cout_pad, dgrad, nc_for(gout), gin == 8, total);
    VS_CHECK_LAUNCH("pack_tc_padded_kernel");
    return VS_OK;
}

extern "C" int vs_pack_conv3_weight_tc(const float* w, void* out, int cin, int cout, int dgrad, void* stream) {
    const size_t bytes = vs_conv3_tc_pack_bytes(cin, cout, dgrad);
    VS_REQUIRE(w && out && bytes > 0, VS_ERR_UNSUPPORTED, "pack_conv3_weight_tc: unsupported shape Cin=%d Cout=%d", cin, cout);
    const int gin = dgrad ? cout : cin, gout = dgrad ? cin : cout;
    const long long total = (long long)(bytes / 2);
    pack_tc_kernel<<<(unsigned)min(1024LL, (total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        w, (bf16*)out, cin, cout, dgrad, nc_for(gout), gin == 8, total);
    VS_CHECK_LAUNCH("pack_tc_kernel");
    return VS_OK;
}

// All derived weight packs of a network in ONE launch (the fused optimiser changes every weight every step):
// blockIdx.y = job; see vs_pack_job in the header.
__global__ void __launch_bounds__(256) pack_batched_kernel(const vs_pack_job* __restrict__ jobs) {
    const vs_pack_job j = jobs[blockIdx.y];
    const float* w = (const float*)j.w;
    const int cin = j.cin, cout = j.cout, cpad = j.cout_pad;
    const int cinpad = j.cin_pad > cin ? j.cin_pad : cin;
    const long long stride = (long long)gridDim.x * blockDim.x, t0 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (j.kind == 1) {
        // 2x2x2 stride-2 layer: wt[A = cout][B = cin][8] -> gather / scatter packs (k2s2_tc.cu)
        if (j.tcf != nullptr) {
            bf16* o = (bf16*)j.tcf;
            for (long long i = t0; i < j.tcf_elems; i += stride) o[i] = __float2bfloat16_rn(pack_k2s2_elem(w, i, cout, cin, 0));
        }
        if (j.tcd != nullptr) {
            bf16* o = (bf16*)j.tcd;
            for (long long i = t0; i < j.tcd_elems; i += stride) o[i] = __float2bfloat16_rn(pack_k2s2_elem(w, i, cout, cin, 1));
        }
        return;
    }
    if (j.wf != nullptr || j.wd != nullptr) {
        float* wf = (float*)j.wf; float* wd = (float*)j.wd;
        const long long total = (long long)cout * cin * 27;
        for (long long i = t0; i < total; i += stride) {
            const int tap = (int)(i % 27), ci = (int)((i / 27) % cin), co = (int)(i / (27 * cin));
            const float v = w[i];
            if (wf != nullptr) wf[((long long)tap * cin + ci) * cout + co] = v;
            if (wd != nullptr) wd[((long long)(26 - tap) * cout + co) * cin + ci] = v;
        }
    }
    if (j.tcf != nullptr) {
        bf16* o = (bf16*)j.tcf;
        const int nc = nc_for_dev(cpad);
        for (long long i = t0; i < j.tcf_elems; i += stride)
            o[i] = __float2bfloat16_rn(pack_tc_elem(w, i, cin, cpad, cout, 0, nc, cin == 8));
    }
    if (j.tcd != nullptr) {
        bf16* o = (bf16*)j.tcd;
        const int nc = nc_for_dev(cinpad);
        for (long long i = t0; i < j.tcd_elems; i += stride)
            o[i] = __float2bfloat16_rn(pack_tc_elem(w, i, cinpad, cpad, cout, 1, nc, cpad == 8, cin));
    }
    if (j.kdn != nullptr) {
        bf16* o = (bf16*)j.kdn;
        for (long long i = t0; i < j.kdn_elems; i += stride) o[i] = __float2bfloat16_rn(pack_kdn_elem(w, i, cinpad, cpad, j.kdn_dgrad, cin, cout));
    }
}

// y[n,d,h,w,gout] = conv3(x[n,d,h,w,gin], wtc) on the tensor cores; bf16 NDHWC in and out.
static int run_conv3_tc(const void* x, const void* wtc, void* y, double* stats, float* shift, int prezeroed,
                        const void* yprev, const double* pstats, double* psums, int planar_mode,
                        float* yplanar, const float* bias, int n, int d,
                        int h, int w, int gin, int gout, void* stream, void* a_out, const void* skip, unsigned* gbar);
extern "C" int vs_conv3x3x3_tc(const void* x, const void* wtc, void* y, double* stats, float* shift, int prezeroed,
                               const void* yprev, const double* pstats, double* psums, int planar_mode,
                               float* yplanar, const float* bias, int n, int d,
                               int h, int w, int gin, int gout, void* stream) {
    return run_conv3_tc(x, wtc, y, stats, shift, prezeroed, yprev, pstats, psums, planar_mode, yplanar, bias, n, d, h, w, gin, gout,
                        stream, nullptr, nullptr, nullptr);
}
// conv3 + InstanceNorm3d(eps 1e-5, biased) + ReLU (+ skip add) in ONE cooperative launch (joint_model.py:40-46,106):
// y = raw conv output (shifted), a = relu((y - mean) * rstd) + skip.  stats / shift / gbar zeroed by the caller.
extern "C" int vs_conv3x3x3_tc_in_relu(const void* x, const void* wtc, void* y, void* a, const void* skip, double* stats,
                                       float* shift, unsigned* gbar, int n, int d, int h, int w, int gin, int gout, void* stream) {
    VS_REQUIRE(a && stats && gbar && vs_aligned16(a) && vs_aligned16(skip) && (long long)n * gout * 8 <= 64 * 1024, VS_ERR_SHAPE,
               "conv3_tc_in_relu: needs the activation output, pre-zeroed statistics and a barrier word");
    return run_conv3_tc(x, wtc, y, stats, shift, 1, nullptr, nullptr, nullptr, 0, nullptr, nullptr, n, d, h, w, gin, gout, stream,
                        a, skip, gbar);
}
static int run_conv3_tc(const void* x, const void* wtc, void* y, double* stats, float* shift, int prezeroed,
                        const void* yprev, const double* pstats, double* psums, int planar_mode,
                        float* yplanar, const float* bias, int n, int d,
                        int h, int w, int gin, int gout, void* stream, void* a_out, const void* skip, unsigned* gbar) {
    VS_REQUIRE(x && wtc && (y || planar_mode), VS_ERR_SHAPE, "conv3_tc: null pointer");
    if (planar_mode) VS_REQUIRE(yplanar && gout == 8 && !stats && !shift && !psums, VS_ERR_SHAPE, "conv3_tc: planar output needs Cout padded to 8 and no statistics");
    VS_REQUIRE((gin == 8 || (gin % 16 == 0 && gin >= 16)) && gout % 8 == 0 && gout >= 8, VS_ERR_UNSUPPORTED,
               "conv3_tc: unsupported channels Cin=%d Cout=%d", gin, gout);
    VS_REQUIRE(vs_aligned16(x) && vs_aligned16(y) && vs_aligned16(wtc), VS_ERR_ALIGN, "conv3_tc: pointers must be 16B aligned");
    EncodeTiledFn encode = get_encode_fn();
    VS_REQUIRE(encode != nullptr, VS_ERR_CUDA, "conv3_tc: cuTensorMapEncodeTiled unavailable");
    cudaStream_t st = (cudaStream_t)stream;

    CUtensorMap map;
    const cuuint64_t gdim[5] = {(cuuint64_t)gin, (cuuint64_t)w, (cuuint64_t)h, (cuuint64_t)d, (cuuint64_t)n};
    const cuuint64_t gstr[4] = {(cuuint64_t)gin * 2, (cuuint64_t)w * gin * 2, (cuuint64_t)h * w * gin * 2,
                                (cuuint64_t)d * h * w * gin * 2};
    // d-planes per work item and the staged box, both fitted to the volume (see TcParams)
    const int tiles_h = (h + TH - 1) / TH, tiles_w = (w + TW - 1) / TW;
    const int nc = nc_for(gout);
    const int nchunks = (gout + nc - 1) / nc;
    // 4 planes per item unless fewer still fit one wave of CTAs -- the deep levels are latency-bound on a handful of
    // CTAs that each serialise 4 planes x all k-slices on one tensor pipe.  (One item per CTA also keeps the shift
    // hand-off trivially deadlock-free: the producing item is never queued behind a waiting one.)
    int td = TD;
    while (td > 1 && (long long)n * ((d + td / 2 - 1) / (td / 2)) * tiles_h * tiles_w * nchunks <= (long long)vs_sm_count())
        td >>= 1;                                                   // stays a single wave: every CTA has one item
    const int bw = min(HW, w + 2), bh = min(HH, h + 2), bd = min(HD, min(td, d) + 2);
    const cuuint32_t box[5] = {8, (cuuint32_t)bw, (cuuint32_t)bh, (cuuint32_t)bd, 1};
    const cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    CUresult cr;
    if (gin == 8) {
        // Cin = 8: a w-row of the halo is 10 voxels x 16 B contiguous in memory; merging (w, c) into one
        // dimension makes it ONE 160 B TMA row instead of ten 16 B rows (TMA cost is per row)
        const cuuint64_t gdim4[4] = {(cuuint64_t)w * 8, (cuuint64_t)h, (cuuint64_t)d, (cuuint64_t)n};
        const cuuint64_t gstr4[3] = {(cuuint64_t)w * 16, (cuuint64_t)h * w * 16, (cuuint64_t)d * h * w * 16};
        const cuuint32_t box4[4] = {(cuuint32_t)bw * 8, (cuuint32_t)bh, (cuuint32_t)bd, 1};
        cr = encode(&map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(x), gdim4, gstr4, box4, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    } else
    cr = encode(&map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(x), gdim, gstr, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    VS_REQUIRE(cr == CUDA_SUCCESS, VS_ERR_CUDA, "conv3_tc: cuTensorMapEncodeTiled failed (%d)", (int)cr);

    TcParams p;
    p.n = n; p.d = d; p.h = h; p.w = w; p.cin = gin; p.cout = gout;
    p.tiles_h = tiles_h; p.tiles_w = tiles_w;
    p.nchunks = nchunks;
    p.td = td; p.bw = bw; p.bh = bh; p.bd = bd;
    {
        const int nsplit = TD / td, ksl = gin == 8 ? 1 : gin / 16;
        p.nsum = g_conv3_ksplit ? (nsplit < ksl ? nsplit : ksl) : 1;
    }
    p.tiles_d = (d + p.td - 1) / p.td;
    p.tiles_per_n = p.tiles_d * p.tiles_h * p.tiles_w;
    p.ref_tile = (min(1, d - 1) / td) * tiles_h * tiles_w;        // voxel (1,1,1): h- and w-tile 0, d-tile 1/td
    p.kslices = gin == 8 ? 1 : gin / 16;
    p.work_items = (long long)n * p.tiles_per_n * p.nchunks;
    VS_REQUIRE(p.work_items < 2147483647LL, VS_ERR_SHAPE, "conv3_tc: too many work items");
    p.wpack = (const bf16*)wtc; p.y = (bf16*)y; p.stats = stats; p.shift = shift;
    p.yprev = (const bf16*)yprev; p.pstats = pstats; p.psums = psums; p.inv_s = 1.0 / ((double)d * h * w);
    p.planar_mode = planar_mode; p.yplanar = yplanar; p.bias = bias;
    p.a_out = (bf16*)a_out; p.skip = (const bf16*)skip; p.gbar = gbar;
    p.dbg = g_tc_dbg;
    if (psums != nullptr) {
        VS_REQUIRE(yprev && pstats && stats == nullptr && shift == nullptr, VS_ERR_SHAPE,
                   "conv3_tc: the fused norm-backward reduction needs y_prev + stats_prev and no forward statistics");
        VS_REQUIRE(vs_aligned16(yprev), VS_ERR_ALIGN, "conv3_tc: y_prev must be 16B aligned");
    }
    // zero the statistics and the shift/flag words (one memset when the caller laid them out back to back)
    const size_t stat_bytes = sizeof(double) * 2 * (size_t)n * gout, shift_bytes = sizeof(float) * (size_t)n * gout;
    if (prezeroed) {
        // the caller zeroed statistics and shift words (one arena memset per network pass instead of one per layer)
    } else if (stats && shift && reinterpret_cast<char*>(shift) == reinterpret_cast<char*>(stats) + stat_bytes) {
        VS_CUDA(cudaMemsetAsync(stats, 0, stat_bytes + shift_bytes, st), "conv3_tc stats+shift memset");
    } else {
        if (stats) VS_CUDA(cudaMemsetAsync(stats, 0, stat_bytes, st), "conv3_tc stats memset");
        if (shift) VS_CUDA(cudaMemsetAsync(shift, 0, shift_bytes, st), "conv3_tc shift memset");
    }

    if (gin == 8) {
        if (nc == 16 && gout == 8) return launch_tc<16, true, 4, true>(map, p, st);
        if (nc == 16) return launch_tc<16, true, 4>(map, p, st);
        if (nc == 32) return launch_tc<32, true, 4>(map, p, st);
        return launch_tc<64, true, 3>(map, p, st);
    }
    if (nc == 16 && gout == 8) return launch_tc<16, false, 4, true>(map, p, st);
    if (nc == 16) return launch_tc<16, false, 4>(map, p, st);
    if (nc == 32) return launch_tc<32, false, 3>(map, p, st);
    return launch_tc<64, false, 2>(map, p, st);
}

extern "C" void vs_debug_set_conv3_ksplit(int on) { g_conv3_ksplit = on; }
extern "C" void vs_debug_set_tc_phase_buffer(void* dev_ptr) { g_tc_dbg = (long long*)dev_ptr; }

extern "C" int vs_pack_conv3_batched(const void* jobs_dev, int njobs, void* stream) {
    VS_REQUIRE(jobs_dev && njobs > 0, VS_ERR_SHAPE, "pack_conv3_batched: bad arguments");
    dim3 grid((unsigned)vs_sm_count(), (unsigned)njobs);    // grid-stride per job: the 128/256-channel layers carry most elements
    pack_batched_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>((const vs_pack_job*)jobs_dev);
    VS_CHECK_LAUNCH("pack_batched_kernel");
    return VS_OK;
}
