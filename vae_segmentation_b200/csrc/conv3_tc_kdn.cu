// "kd-in-N" variant of the tcgen05 3x3x3 convolution for the full-resolution layers (Cout = 8 or 16): the default kernel
// for their forward AND input-gradient passes (engine.USE_KDN).
//
// conv3_tc.cu issues, for each of the 4 output planes of a tile, all 27 taps: every input voxel's A operand is fetched
// 27 x from shared memory, which is what bounds those layers once the epilogue is out of the way (operand pipe 58 %).
// Here the MMAs are issued per INPUT plane q of the halo (6 of them) and in-plane tap (kh,kw):
//
//     D[128 voxels, (kd', co)] += A[plane q, shifted by (kh,kw)] * B'[(kd', co), ci]        N = 4 * Cout
//
// B' holds the three kd taps of (kh,kw) side by side (block 3 = zeros, N must be a multiple of 16).  Block kd' belongs
// to OUTPUT plane p = q - kd', so with the accumulator of plane p at TMEM columns (5 - p) * Cout the three blocks of
// one MMA land in consecutive "slots"; 9 slots (p = -3 .. 5; only p = 0..3 are real, the others collect partial sums
// of planes outside the tile and are never read) form one buffer, two buffers double-buffer the tile.
//   * every MMA accumulates (one MMA touches fresh and running slots at once): the epilogue zeroes a buffer with
//     tcgen05.st after reading it, and zeroes both buffers before the first tile;
//   * the four issuer warps share the accumulators and simply split the (plane, tap) list round-robin;
//   * Cin = 8 pairs two in-plane taps per K = 16 MMA (second K chunk = first one displaced by LBO): 5 MMAs per plane.
// Hardware facts this relies on were probed first (tools/kdn_probe.cu, profiles/r1_kdn_probe.txt): D may start at
// any column that is a multiple of 8 (N = 32), accumulating MMAs issued by different warps into the same columns add
// up exactly, tcgen05.st zeroes an accumulator.
// Per tile and 16-channel slice: 30 (Cin = 8) / 54 MMAs instead of 56 / 108, A-operand traffic halved.
//
// Feature set: bf16 NDHWC output; fprop: shift + fp64 InstanceNorm statistics; dgrad: optionally the PREVIOUS layer's
// InstanceNorm-backward sums fused into the epilogue (as conv3_tc.cu); tiles of 4 d-planes only (volumes with D >= 4);
// no planar (head) epilogue.
// What bounds it (profiles/r2_kdn_phase_probe.txt, clock64 stamps around a steady-state tile): the MMAs themselves.
// A tile of 512 voxels is 30 (Cin = 8) / 54 MMAs at N = 32, each paced by its 4 KB A-operand read from shared memory
// (~40 cycles): ~1250 cycles per tile, which the epilogue (one warpgroup, ~1100 cycles of stores + 850 of decode and
// prefetch, overlapped through 3-4 TMEM buffers) just keeps up with.  Tried and measured (same profile): a second
// epilogue warpgroup (VS_KDN_NEPI=2: 36.7 vs 39.1 us on the 8 -> 8 layer but 28.5 vs 23.2 on 16 -> 16, slower in total),
// descriptors precomputed once per kernel (45 -> 10 uniform instructions per MMA, no change: the issue slots were not
// the limit), 2 vs 4 accumulator buffers (no change).  Warp roles are ordered epilogue < producer < issuers because
// the scheduler favours the highest warp index of a sub-partition.
#include "tc_ptx.cuh"
#include "tc_pack.cuh"

namespace {

constexpr int TD = 4, TH = 16, TW = 8;
constexpr int HD = TD + 2, HH = TH + 2, HW = TW + 2;
constexpr int HV = HD * HH * HW;
constexpr int PLANE_BYTES = HV * 16;
constexpr int PLANE_PAD = 256;
#ifndef VS_KDN_NEPI
#define VS_KDN_NEPI 1
#endif
#ifndef VS_KDN_NBUF8
#define VS_KDN_NBUF8 4
#endif
constexpr int NEPI = VS_KDN_NEPI;                      // epilogue warpgroups
constexpr int NTHREADS = 160 + 128 * NEPI;             // NEPI x 4 epilogue warps + 1 producer + 4 MMA warps
constexpr int NSLOT = 9;

struct KdnParams {
    int n, d, h, w, cin, cout;
    int tiles_d, tiles_h, tiles_w, tiles_per_n;
    int bw, bh, bd;
    int ref_tile;
    int kslices;
    int work_items;
    const bf16* wpack;
    bf16* y;
    double* stats;
    float* shift;
    // dgrad only -- fused InstanceNorm+ReLU backward reduction of the PREVIOUS layer (see conv3_tc.cu TcParams)
    const bf16* yprev;
    const double* pstats;
    double* psums;
    double inv_s;
    // planar fp32 output of GEMM output channels 0..1 instead of the bf16 NDHWC store (NCO = 8; see conv3_tc.cu TcParams):
    // 1 = 2-class head, probs = softmax(acc + bias); 2 = plain values (2-channel planar input gradient)
    int planar_mode;
    float* yplanar;
    const float* bias;
    // fused InstanceNorm + ReLU (+ skip): a = relu((y - mean) * rstd) + skip written by the same launch after a grid
    // barrier (cooperative launch); needs stats; gbar = zero-initialised counter
    bf16* a_out;
    const bf16* skip;
    unsigned* gbar;
    int ordered;             // 1: one issuer warp issues the whole (plane, tap) list in a fixed order (bit-reproducible sums)
    long long* dbg;          // tools/kdn_phase_probe.py: clock64 stamps of CTA 0 around its 9th tile, or null
};
#define KDN_DBG(slot) do { if (p.dbg != nullptr && blockIdx.x == 0) p.dbg[slot] = clock64(); } while (0)
constexpr int KDN_DBG_TILE = 8;

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

__device__ __forceinline__ void tmem_st8_zero(uint32_t taddr) {
    const uint32_t z = 0u;
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %1, %1, %1, %1, %1, %1, %1};" ::"r"(taddr), "r"(z) : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// work item -> (sample, tile origin); the tiles holding the shift's reference voxel come first (see conv3_tc.cu)
__device__ __forceinline__ void decode_item(unsigned item, const KdnParams& p, int& n, int& d0, int& h0, int& w0) {
    int r;
    if (item < (unsigned)p.n) {
        n = (int)item;
        r = p.ref_tile;
    } else {
        const unsigned t = item - (unsigned)p.n;
        n = (int)(t / (unsigned)(p.tiles_per_n - 1));
        r = (int)(t % (unsigned)(p.tiles_per_n - 1));
        r += (r >= p.ref_tile);
    }
    w0 = (r % p.tiles_w) * TW; r /= p.tiles_w;
    h0 = (r % p.tiles_h) * TH; r /= p.tiles_h;
    d0 = r * TD;
}

// FEAT: compile-time feature set -- the rarely used paths live in their OWN instantiations so that the default kernel
// (FEAT = 0: bf16 NDHWC store, concurrent issue) stays as small as it was before they existed: with the planar epilogue,
// the ordered issue and the fused InstanceNorm pass compiled into one body the 2 x 96^3 8->8 launch went from 31 to 41 us
// under ncu (registers 79 -> 128, the warp-specialised roles no longer fit the instruction cache together).
constexpr int KDN_PLANAR = 1, KDN_APPLY = 2, KDN_ORDERED = 4, KDN_REDUCE = 8;
template <bool CIN8, int NCO, int NSTAGE, int FEAT>
__global__ void __launch_bounds__(NTHREADS, 1) conv3_tc_kdn_kernel(const __grid_constant__ CUtensorMap xmap, KdnParams p) {
    constexpr int N = 4 * NCO;                                // kd' blocks 0..2 + one zero block
    constexpr int NMP = CIN8 ? 5 : 9;                         // MMAs per input plane per 16-channel slice
    constexpr int A_BYTES = CIN8 ? (PLANE_BYTES + PLANE_PAD) : 2 * PLANE_BYTES;
    constexpr int B_BYTES = NMP * N * 32;                     // [mma][kc 2][N/8][8 rows][16 B]
    constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    // TMEM accumulator buffers: the MMAs of tile i wait for the epilogue of tile i - NBUF, and an epilogue is a LONG
    // dependent chain (tfull wake-up, TMEM loads, global stores, TMEM zeroing: ~2400 cycles against ~1200 of MMAs), so
    // with two buffers the tile period was (MMA + epilogue) / 2 whatever the number of epilogue warps
    constexpr int NBUF = NCO == 8 ? VS_KDN_NBUF8 : 3;
    constexpr int BUFC = NSLOT * NCO;                         // accumulator columns per buffer (72 / 144)
    constexpr int TMEM_COLS = (NBUF * BUFC <= 256) ? 256 : 512;
    static_assert(STAGE_BYTES % 128 == 0, "stage alignment");
    static_assert(NBUF * BUFC <= TMEM_COLS, "TMEM budget");

    extern __shared__ __align__(128) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~uintptr_t(127));
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + NSTAGE * STAGE_BYTES);
    uint64_t* empty_bar = full_bar + NSTAGE;
    uint64_t* tfull_bar = empty_bar + NSTAGE;
    uint64_t* tempty_bar = tfull_bar + NBUF;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + NBUF);
    float* sshift = reinterpret_cast<float*>(tmem_slot + 4);          // [2 groups][16]
    float* smean = sshift + 32;                                       // [2][16]  (fused norm-backward reduction)
    float* srstd = smean + 32;                                        // [2][16]
    uint64_t* tab_a = reinterpret_cast<uint64_t*>(srstd + 32);        // [HD * NMP] operand descriptors relative to stage 0 (ordered mode)
    uint64_t* tab_b = tab_a + HD * 9;
    uint32_t* tab_d = reinterpret_cast<uint32_t*>(tab_b + HD * 9);

    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
    const int lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        prefetch_tensormap(&xmap);
        for (int s = 0; s < NSTAGE; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 4); }
        for (int b = 0; b < NBUF; ++b) { mbar_init(&tfull_bar[b], 4); mbar_init(&tempty_bar[b], 4); }
        fence_barrier_init();
    }
    if (CIN8) {
        // the zero-weighted last tap of the pairs reads one voxel past the staged box (see conv3_tc.cu)
        const bool clipped = p.bw < HW || p.bh < HH || p.bd < HD;
        const int words = clipped ? A_BYTES / 4 : PLANE_PAD / 4, off = clipped ? 0 : PLANE_BYTES;
        for (int i = threadIdx.x; i < NSTAGE * words; i += NTHREADS) {
            int s = i / words, k = i % words;
            reinterpret_cast<uint32_t*>(smem + s * STAGE_BYTES + off)[k] = 0u;
        }
        fence_proxy_async();
    }
    if (warp == 4 * NEPI + 1) tmem_alloc(tmem_slot, TMEM_COLS);
    pdl_trigger();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);
    pdl_wait();                          // see vs_common.cuh: no global-memory access above this line

    // Warp roles, lowest to highest index: 2 x 4 epilogue warps, the TMA producer, 4 MMA issuers.  The scheduler favours the
    // highest warp index among the eligible warps of a sub-partition: the issuers (one MMA per ~100 issue cycles of
    // descriptor arithmetic) must not queue behind the two epilogue warps they share it with.
    if (warp == 4 * NEPI) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            uint32_t stage = 0, phase = 0;
            for (int item = blockIdx.x; item < p.work_items; item += gridDim.x) {
                int n, d0, h0, w0;
                decode_item((unsigned)item, p, n, d0, h0, w0);
                for (int ks = 0; ks < p.kslices; ++ks) {
                    mbar_wait(&empty_bar[stage], phase ^ 1);
                    uint8_t* sa = smem + stage * STAGE_BYTES;
                    mbar_expect_tx(&full_bar[stage], (uint32_t)((CIN8 ? 1 : 2) * p.bd * p.bh * p.bw * 16 + B_BYTES));
                    if (CIN8) tma_load_4d(sa, &xmap, &full_bar[stage], (w0 - 1) * 8, h0 - 1, d0 - 1, n);
                    else tma_load_5d(sa, &xmap, &full_bar[stage], ks * 16, w0 - 1, h0 - 1, d0 - 1, n);
                    if (!CIN8) tma_load_5d(sa + PLANE_BYTES, &xmap, &full_bar[stage], ks * 16 + 8, w0 - 1, h0 - 1, d0 - 1, n);
                    bulk_load(sa + A_BYTES, reinterpret_cast<const uint8_t*>(p.wpack) + (long long)ks * B_BYTES, B_BYTES, &full_bar[stage]);
                    if (item == (int)blockIdx.x + KDN_DBG_TILE * (int)gridDim.x) KDN_DBG(0);
                    if (item == (int)blockIdx.x + (KDN_DBG_TILE + 1) * (int)gridDim.x) KDN_DBG(1);
                    if (++stage == NSTAGE) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp > 4 * NEPI) {
        // ===================== MMA issuers: the (input plane, tap) list split round-robin over the 4 warps ==========
        const int j = warp - (4 * NEPI + 1);
        constexpr uint32_t idesc = make_idesc(N);
        // The operand descriptors of this warp's (plane, tap) list depend on the tile only through the stage base
        // address: they are built ONCE (relative to stage 0) and the issue loop adds the stage offset -- the ~45 uniform
        // instructions of descriptor arithmetic per MMA were what the issuer warps spent their issue slots on
        // (profiles/r2_fullres_ncu.txt: the four sub-partitions issued on 46 % of all cycles, shared with the epilogue).
        constexpr int MAXI = (HD * NMP + 3) / 4;
        uint64_t ad_rel[MAXI], bd_rel[MAXI];
        uint32_t dcol_rel[MAXI];
        uint32_t valid = 0u;
        {
            const uint32_t a0 = smem_u32(smem);
            const uint32_t b_base = a0 + A_BYTES;
#pragma unroll
            for (int k = 0; k < MAXI; ++k) {
                const int i = j + 4 * k;
                const int q = i / NMP, m = i - q * NMP;
                const uint32_t a_plane = a0 + (uint32_t)(q * p.bh * p.bw) * 16u;
                if (CIN8) {
                    const int t1 = 2 * m, t2 = (2 * m + 1 < 9) ? 2 * m + 1 : 8;
                    const int o1 = (t1 / 3) * p.bw + t1 % 3, o2 = (t2 / 3) * p.bw + t2 % 3;
                    const uint32_t lbo = (2 * m + 1 < 9) ? (uint32_t)(o2 - o1) * 16u : 16u;
                    ad_rel[k] = make_desc(a_plane + (uint32_t)o1 * 16u, lbo, (uint32_t)p.bw * 16u);
                } else {
                    const int kh = m / 3, kw = m - kh * 3;
                    ad_rel[k] = make_desc(a_plane + (uint32_t)(kh * p.bw + kw) * 16u, PLANE_BYTES, (uint32_t)p.bw * 16u);
                }
                bd_rel[k] = make_desc(b_base + m * (N * 32), N * 16, 128u);
                dcol_rel[k] = (uint32_t)((5 - q) * NCO);                                  // slots (5-q) .. (5-q)+3
                if (i < HD * NMP && q < p.bd) valid |= 1u << k;                          // q >= bd: plane not staged (never with D >= 4)
                if ((FEAT & KDN_ORDERED) && lane == 0 && i < HD * NMP) { tab_a[i] = ad_rel[k]; tab_b[i] = bd_rel[k]; tab_d[i] = dcol_rel[k]; }
            }
        }
        // Ordered mode: accumulating MMAs issued by different warps into the same TMEM columns add in ISSUE order, which
        // varies from run to run (fp32 rounding of y moves, a few bf16 roundings flip).  Here warp j = 0 alone issues the
        // whole list, in list order, from the descriptor table the four warps just filled; the other three only keep the
        // barrier protocol going (their commits track no MMAs and arrive at once).
        if (FEAT & KDN_ORDERED) asm volatile("bar.sync 8, 128;" ::: "memory");
        uint32_t stage = 0, phase = 0, buf = 0, bphase = 0;
        for (int item = blockIdx.x; item < p.work_items; item += gridDim.x) {
            const bool dbg_tile = j == 0 && lane == 0 && item == (int)blockIdx.x + KDN_DBG_TILE * (int)gridDim.x;
            const bool dbg_next = j == 0 && lane == 0 && item == (int)blockIdx.x + (KDN_DBG_TILE + 1) * (int)gridDim.x;
            if (dbg_tile) KDN_DBG(2);
            mbar_wait(&tempty_bar[buf], bphase);                 // zeroed and released by the epilogue (also initially)
            tc_fence_after();
            if (dbg_tile) KDN_DBG(3);
            if (dbg_next) KDN_DBG(10);
            const uint32_t dbuf = tmem_base + buf * BUFC;
            for (int ks = 0; ks < p.kslices; ++ks) {
                mbar_wait(&full_bar[stage], phase);
                tc_fence_after();
                if (dbg_tile && ks == 0) KDN_DBG(4);
                const uint64_t soff = (uint64_t)((stage * (uint32_t)STAGE_BYTES) >> 4);   // start-address field, 16-byte units
                if (FEAT & KDN_ORDERED) {
                    if (j == 0) {
#pragma unroll 6
                        for (int i = 0; i < HD * NMP; ++i)
                            if (i / NMP < p.bd) tc_mma_elect(dbuf + tab_d[i], tab_a[i] + soff, tab_b[i] + soff, idesc, 1u);
                    }
                } else {
#pragma unroll
                    for (int k = 0; k < MAXI; ++k)
                        if (valid & (1u << k)) tc_mma_elect(dbuf + dcol_rel[k], ad_rel[k] + soff, bd_rel[k] + soff, idesc, 1u);
                }
                tc_commit_elect(&empty_bar[stage]);
                if (++stage == NSTAGE) { stage = 0; phase ^= 1; }
            }
            tc_commit_elect(&tfull_bar[buf]);
            if (dbg_tile) KDN_DBG(5);
            if (++buf == NBUF) { buf = 0; bphase ^= 1; }
        }
    } else {
        // ===================== epilogue: warpgroup e = warp / 4 drains TMEM buffer e (tiles e, e+2, ...) ==========
        const int q4 = warp & 3;
        const int e = warp >> 2;                                 // < NEPI
        const int et = threadIdx.x - 128 * e;
        const int row = q4 * 32 + lane;
        const int lh = row >> 3, lw = row & 7;
        const uint32_t lane_q = tmem_base + ((uint32_t)(q4 * 32) << 16);
        const uint32_t bar_id = 1u + (uint32_t)e;
#define EPI_SYNC() asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory")
        // the buffers start zeroed and "empty" (with two warpgroups each owns one buffer)
        for (int b = 0; b < NBUF; ++b) {
            if (b % NEPI != e) continue;
            for (int c = 0; c < BUFC; c += 8) tmem_st8_zero(lane_q + b * BUFC + c);
            tmem_st_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty_bar[b]);
        }
        int jt = e;                                              // index of the tile among this CTA's tiles: buffer jt % NBUF, use jt / NBUF
        int stat_n = -1;
        float rs[16], rq[16], shr[16];
#pragma unroll
        for (int k = 0; k < 16; ++k) { rs[k] = 0.f; rq[k] = 0.f; shr[k] = 0.f; }
        const bool fused = (FEAT & KDN_REDUCE) && p.psums != nullptr;     // compile-time false outside the dgrad-with-reduction instantiation
        double* const sout = fused ? p.psums : p.stats;
        auto flush_stats = [&]() {
            if (sout != nullptr && stat_n >= 0) {
                const float s1 = transpose_reduce16(rs, lane);
                const float s2 = transpose_reduce16(rq, lane);
                if (lane < NCO) {
                    atomicAdd(&sout[((long long)stat_n * p.cout + lane) * 2], (double)s1);
                    atomicAdd(&sout[((long long)stat_n * p.cout + lane) * 2 + 1], (double)s2);
                }
#pragma unroll
                for (int k = 0; k < 16; ++k) { rs[k] = 0.f; rq[k] = 0.f; }
            }
        };
        float* const my_shift = sshift + 16 * e;
        if (et < 16) my_shift[et] = 0.f;
        EPI_SYNC();
        const int rd = min(1, p.d - 1), rh = min(1, p.h - 1), rw = min(1, p.w - 1);
        const bool has_shift = p.shift != nullptr;
        const bool has_stats = p.stats != nullptr;
        const uint32_t sshift_addr = smem_u32(my_shift);
        float hb0 = 0.f, hb1 = 0.f;
        if ((FEAT & KDN_PLANAR) && p.planar_mode == 1 && p.bias != nullptr) { hb0 = p.bias[0]; hb1 = p.bias[1]; }
        const long long vol = (long long)p.d * p.h * p.w;
        const uint32_t smean_addr = smem_u32(smean + 16 * e), srstd_addr = smem_u32(srstd + 16 * e);
        const int ref_row = rh * TW + rw;
        for (int item = blockIdx.x + e * (int)gridDim.x; item < p.work_items; item += NEPI * (int)gridDim.x, jt += NEPI) {
            if (et == 0 && item == (int)blockIdx.x + (KDN_DBG_TILE + NEPI) * (int)gridDim.x) KDN_DBG(11);
            if (et == 0 && item == (int)blockIdx.x + KDN_DBG_TILE * (int)gridDim.x) KDN_DBG(12);
            int n, d0, h0, w0;
            decode_item((unsigned)item, p, n, d0, h0, w0);
            const int buf = jt % NBUF;
            const uint32_t bphase = (uint32_t)(jt / NBUF) & 1u;
            const uint32_t lane_base = lane_q + (uint32_t)(buf * BUFC);
            if (n != stat_n) {
                flush_stats();
                stat_n = n;
                if (fused) {
                    EPI_SYNC();                                  // everyone is done with the previous sample's constants
                    if (et < NCO) {
                        float mu, rsd;
                        in_mean_rstd(p.pstats + ((long long)n * p.cout + et) * 2, p.inv_s, mu, rsd);
                        smean[16 * e + et] = mu; srstd[16 * e + et] = rsd;
                    }
                    EPI_SYNC();
                }
                if (has_shift) {
                    // shift hand-off exactly as in conv3_tc.cu: the CTA owning the reference voxel (tile 0 of the sample, the
                    // first item of CTA n) publishes the fp32 accumulator of voxel (1,1,1); everyone else spins on non-zero
                    EPI_SYNC();
                    if (d0 == 0 && h0 == 0 && w0 == 0) {
                        mbar_wait(&tfull_bar[buf], bphase);
                        tc_fence_after();
                        if (q4 == (ref_row >> 5)) {
                            uint32_t r[16];
#pragma unroll
                            for (int k = 0; k < 16; ++k) r[k] = 0u;
                            {
                                uint32_t r8[8];
                                tmem_ld8(lane_base + (5 - rd) * NCO, r8);
                                tmem_ld_wait();
#pragma unroll
                                for (int k = 0; k < 8; ++k) r[k] = r8[k];
                                if (NCO == 16) {
                                    tmem_ld8(lane_base + (5 - rd) * NCO + 8, r8);
                                    tmem_ld_wait();
#pragma unroll
                                    for (int k = 0; k < 8; ++k) r[8 + k] = r8[k];
                                }
                            }
                            if (lane == (ref_row & 31)) {
#pragma unroll
                                for (int k = 0; k < NCO; ++k) {
                                    const uint32_t bits = r[k] | 1u;
                                    my_shift[k] = __uint_as_float(bits);
                                    *reinterpret_cast<volatile uint32_t*>(p.shift + (long long)n * p.cout + k) = bits;
                                }
                            }
                        }
                    } else if (et < NCO) {
                        const volatile uint32_t* flag = reinterpret_cast<const volatile uint32_t*>(p.shift + (long long)n * p.cout + et);
                        uint32_t bits, spins = 0;
                        while ((bits = *flag) == 0u) {
                            __nanosleep(64);
                            if (++spins > (1u << 23)) __trap();
                        }
                        my_shift[et] = __uint_as_float(bits);
                    }
                    EPI_SYNC();
#pragma unroll
                    for (int k4 = 0; k4 < 4; ++k4) {
                        const float4 sh = lds128(sshift_addr + k4 * 16);
                        shr[k4 * 4 + 0] = sh.x; shr[k4 * 4 + 1] = sh.y; shr[k4 * 4 + 2] = sh.z; shr[k4 * 4 + 3] = sh.w;
                    }
                }
            }
            const int jmax = min(TD, p.d - d0);
            const int gh = h0 + lh, gw = w0 + lw;
            const bool rc_ok = gh < p.h && gw < p.w;
            // fused norm-backward reduction: the previous layer's raw output at this thread's voxels is fetched BEFORE
            // waiting for the accumulators, so the global latency hides behind the MMAs of the tile
            uint4 ypre[TD][NCO / 8];
            if (fused) {
#pragma unroll
                for (int pl = 0; pl < TD; ++pl)
#pragma unroll
                    for (int h8 = 0; h8 < NCO / 8; ++h8) {
                        ypre[pl][h8] = make_uint4(0u, 0u, 0u, 0u);
                        if (pl < jmax && rc_ok)
                            ypre[pl][h8] = __ldg(reinterpret_cast<const uint4*>(
                                p.yprev + ((((long long)n * p.d + d0 + pl) * p.h + gh) * (long long)p.w + gw) * p.cout + h8 * 8));
                    }
            }
            const bool dbg_e = et == 0 && item == (int)blockIdx.x + KDN_DBG_TILE * (int)gridDim.x;
            if (dbg_e) KDN_DBG(6);
            mbar_wait(&tfull_bar[buf], bphase);
            tc_fence_after();
            if (dbg_e) KDN_DBG(7);
#pragma unroll
            for (int pl = 0; pl < TD; ++pl) {
                if (pl >= jmax) break;
                const uint32_t taddr = lane_base + (5 - pl) * NCO;
                float v[NCO];
                {
                    uint32_t r8[8];
                    tmem_ld8(taddr, r8);
                    tmem_ld_wait();
#pragma unroll
                    for (int k = 0; k < 8; ++k) v[k] = __uint_as_float(r8[k]) - shr[k];
                    if (NCO == 16) {
                        tmem_ld8(taddr + 8, r8);
                        tmem_ld_wait();
#pragma unroll
                        for (int k = 0; k < 8; ++k) v[NCO == 16 ? 8 + k : k] = __uint_as_float(r8[k]) - shr[8 + k];
                    }
                }
                if ((FEAT & KDN_PLANAR) && p.planar_mode != 0) {
                    if (rc_ok) {
                        float o0 = v[0] + hb0, o1 = v[1] + hb1;
                        if (p.planar_mode == 1) {
                            const float mx = fmaxf(o0, o1);
                            const float e0 = expf(o0 - mx), e1 = expf(o1 - mx);
                            const float inv = 1.f / (e0 + e1);
                            o0 = e0 * inv; o1 = e1 * inv;
                        }
                        float* pp = p.yplanar + (long long)n * 2 * vol + ((long long)(d0 + pl) * p.h + gh) * p.w + gw;
                        pp[0] = o0;
                        pp[vol] = o1;
                    }
                } else if (rc_ok) {
                    bf16* py = p.y + ((((long long)n * p.d + d0 + pl) * p.h + gh) * (long long)p.w + gw) * p.cout;
#pragma unroll
                    for (int h8 = 0; h8 < NCO / 8; ++h8) {
                        float o[8];
#pragma unroll
                        for (int k = 0; k < 8; ++k) o[k] = v[h8 * 8 + k];
                        Store<bf16>::st8(py + h8 * 8, o);
                    }
                    if (has_stats) {
#pragma unroll
                        for (int k = 0; k < NCO; ++k) { rs[k] += v[k]; rq[k] = fmaf(v[k], v[k], rq[k]); }
                    }
                    if (fused) {
                        // g = this launch's output as stored (rounded to bf16); mask / xhat from the previous layer's raw
                        // output at the same voxel -- exactly what inorm_relu_bwd_reduce_kernel computes
#pragma unroll
                        for (int h8 = 0; h8 < NCO / 8; ++h8) {
                            const uint4 raw = ypre[pl][h8];
                            const uint32_t u[4] = {raw.x, raw.y, raw.z, raw.w};
                            float yv[8];
#pragma unroll
                            for (int i2 = 0; i2 < 4; ++i2) { yv[2 * i2] = __uint_as_float(u[i2] << 16); yv[2 * i2 + 1] = __uint_as_float(u[i2] & 0xffff0000u); }
#pragma unroll
                            for (int k4 = 0; k4 < 2; ++k4) {
                                const float4 mu = lds128(smean_addr + (h8 * 8 + k4 * 4) * 4);
                                const float4 rr = lds128(srstd_addr + (h8 * 8 + k4 * 4) * 4);
                                const float mus[4] = {mu.x, mu.y, mu.z, mu.w}, rrs[4] = {rr.x, rr.y, rr.z, rr.w};
#pragma unroll
                                for (int k = 0; k < 4; ++k) {
                                    const int kk = h8 * 8 + k4 * 4 + k;
                                    const float xh = (yv[k4 * 4 + k] - mus[k]) * rrs[k];
                                    const float gq = __bfloat162float(__float2bfloat16_rn(v[kk]));
                                    const float gm = xh > 0.f ? gq : 0.f;
                                    rs[kk] += gm;
                                    rq[kk] = fmaf(gm, xh, rq[kk]);
                                }
                            }
                        }
                    }
                }
            }
            if (dbg_e) KDN_DBG(8);
            // hand the buffer back ZEROED: every MMA of this kernel accumulates
            for (int c = 0; c < BUFC; c += 8) tmem_st8_zero(lane_base + c);
            tmem_st_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty_bar[buf]);
            if (dbg_e) KDN_DBG(9);
        }
        flush_stats();
#undef EPI_SYNC
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 4 * NEPI + 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, TMEM_COLS);
    }
    if ((FEAT & KDN_APPLY) && p.a_out != nullptr) {
        // ---- fused InstanceNorm + ReLU (+ skip) ----------------------------------------------------------------------
        // The statistics are complete once every CTA has flushed: grid barrier (the launch is cooperative: all CTAs are
        // co-resident), then ALL warps of the CTA -- the pipeline roles are over, the stage buffers are free -- turn the
        // tiles this CTA stored (read back from L2) into the activation.  The separate apply pass (a launch + a read of
        // y from HBM) disappears.
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
        for (int item = blockIdx.x; item < p.work_items; item += gridDim.x, ++it) {
            if (it % NW != warp) continue;
            int n, d0, h0, w0;
            decode_item((unsigned)item, p, n, d0, h0, w0);
            in_relu_apply_tile<NCO / 8, 8>(p.y, p.skip, p.a_out, tab, lane, n, d0, h0, w0, 0, min(TD, p.d - d0), p.d, p.h, p.w, p.cout);
        }
    }
}

__global__ void pack_kdn_kernel(const float* __restrict__ w, bf16* __restrict__ out, int cin, int cout, int dgrad, long long total,
                                int cin_real, int cout_real) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x)
        out[i] = __float2bfloat16_rn(pack_kdn_elem(w, i, cin, cout, dgrad, cin_real, cout_real));
}

template <bool CIN8, int NCO, int NSTAGE, int FEAT>
int launch_kdn_feat(const CUtensorMap& map, const KdnParams& p, cudaStream_t st) {
    constexpr int A_BYTES = CIN8 ? (PLANE_BYTES + PLANE_PAD) : 2 * PLANE_BYTES;
    constexpr int B_BYTES = (CIN8 ? 5 : 9) * 4 * NCO * 32;
    constexpr int SMEM_NEED = NSTAGE * (A_BYTES + B_BYTES) + 128 + 8 * (2 * NSTAGE + 8) + 16 + 3 * 32 * 4 + 64 + HD * 9 * 20 + 16;
    // ONE CTA per SM, enforced through the shared-memory request: a CTA allocates all 512 TMEM columns, so a second
    // resident CTA (possible by registers and by the 92 KB the Cin = 8 variant really needs) would block in tcgen05.alloc
    // while holding its slot -- and when the CTA it blocks behind is spinning on a shift that the blocked CTA has to
    // publish, that is a dead-lock (observed as the spin's trap under the three-branch graph once the default
    // instantiation dropped below 114 registers)
    constexpr int SMEM = SMEM_NEED > 120 * 1024 ? SMEM_NEED : 120 * 1024;
    static_assert(SMEM <= 227 * 1024, "shared memory budget");
    auto kern = conv3_tc_kdn_kernel<CIN8, NCO, NSTAGE, FEAT>;
    static bool configured = false;
    if (!configured) {
        VS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM), "conv3_tc_kdn smem attribute");
        configured = true;
    }
    const int grid = p.work_items < vs_sm_count() ? p.work_items : vs_sm_count();
    if (p.a_out != nullptr)
        VS_CUDA(vs_launch_coop(kern, dim3((unsigned)grid), dim3(NTHREADS), SMEM, st, map, p), "conv3_tc_kdn_kernel cooperative launch");
    else
        VS_CUDA(vs_launch(kern, dim3((unsigned)grid), dim3(NTHREADS), SMEM, st, map, p), "conv3_tc_kdn_kernel launch");
    VS_CHECK_LAUNCH("conv3_tc_kdn_kernel");
    return VS_OK;
}

// picks the instantiation for the features this launch needs (see KDN_* above)
template <bool CIN8, int NCO, int NSTAGE>
int launch_kdn(const CUtensorMap& map, const KdnParams& p, cudaStream_t st) {
    if (p.a_out != nullptr) {
        if (p.ordered) return launch_kdn_feat<CIN8, NCO, NSTAGE, KDN_APPLY | KDN_ORDERED>(map, p, st);
        return launch_kdn_feat<CIN8, NCO, NSTAGE, KDN_APPLY>(map, p, st);
    }
    if (p.planar_mode != 0) {
        if constexpr (NCO == 8) {
            if (p.ordered) return launch_kdn_feat<CIN8, NCO, NSTAGE, KDN_PLANAR | KDN_ORDERED>(map, p, st);
            return launch_kdn_feat<CIN8, NCO, NSTAGE, KDN_PLANAR>(map, p, st);
        } else {
            VS_FAIL(VS_ERR_UNSUPPORTED, "conv3_tc_kdn: planar output needs 8 GEMM output channels");
        }
    }
    if (p.psums != nullptr) {
        if (p.ordered) return launch_kdn_feat<CIN8, NCO, NSTAGE, KDN_REDUCE | KDN_ORDERED>(map, p, st);
        return launch_kdn_feat<CIN8, NCO, NSTAGE, KDN_REDUCE>(map, p, st);
    }
    if (p.ordered) return launch_kdn_feat<CIN8, NCO, NSTAGE, KDN_ORDERED>(map, p, st);
    return launch_kdn_feat<CIN8, NCO, NSTAGE, 0>(map, p, st);
}

}  // namespace

static long long* g_kdn_dbg = nullptr;
static int g_kdn_ordered = 0;
// 1 = bit-reproducible accumulation order in the kd-in-N kernel (one issuer warp); see the kernel's issuer section
extern "C" void vs_set_kdn_ordered(int on) { g_kdn_ordered = on ? 1 : 0; }
extern "C" void vs_debug_set_kdn_phase_buffer(void* dev_ptr) { g_kdn_dbg = (long long*)dev_ptr; }

extern "C" size_t vs_conv3_tc_kdn_pack_bytes(int cin, int cout, int dgrad) {
    const int gin = dgrad ? cout : cin, gout = dgrad ? cin : cout;
    if (!(gin == 8 || (gin % 16 == 0 && gin >= 16)) || !(gout == 8 || gout == 16)) return 0;
    const int kslices = gin == 8 ? 1 : gin / 16;
    return (size_t)kslices * (gin == 8 ? 5 : 9) * 4 * gout * 32;
}

// w[cout][cin][27] fp32 -> kd-in-N pack of a layer with cin_pad x cout_pad channels (the missing ones zero); the pack
// has vs_conv3_tc_kdn_pack_bytes(cin_pad, cout_pad, dgrad) bytes.
extern "C" int vs_pack_conv3_weight_tc_kdn_padded(const float* w, void* out, int cin, int cout, int cin_pad, int cout_pad,
                                                  int dgrad, void* stream) {
    VS_REQUIRE(cin_pad >= cin && cout_pad >= cout, VS_ERR_SHAPE, "pack_conv3_weight_tc_kdn: padded channel counts below the real ones");
    const size_t bytes = vs_conv3_tc_kdn_pack_bytes(cin_pad, cout_pad, dgrad);
    VS_REQUIRE(w && out && bytes > 0, VS_ERR_UNSUPPORTED, "pack_conv3_weight_tc_kdn: unsupported shape Cin=%d Cout=%d", cin_pad, cout_pad);
    const long long total = (long long)(bytes / 2);
    pack_kdn_kernel<<<(unsigned)min(1024LL, (total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(w, (bf16*)out, cin_pad, cout_pad,
                                                                                                  dgrad, total, cin, cout);
    VS_CHECK_LAUNCH("pack_kdn_kernel");
    return VS_OK;
}
extern "C" int vs_pack_conv3_weight_tc_kdn(const float* w, void* out, int cin, int cout, int dgrad, void* stream) {
    return vs_pack_conv3_weight_tc_kdn_padded(w, out, cin, cout, cin, cout, dgrad, stream);
}

// y[n,d,h,w,gout] = conv3(x[n,d,h,w,gin], wkdn); bf16 NDHWC in and out; gout in {8, 16}; D >= 4.
// stats / shift: as vs_conv3x3x3_fprop (zeroed here unless prezeroed).
extern "C" int vs_conv3x3x3_tc_kdn_ex(const void* x, const void* wkdn, void* y, double* stats, float* shift, int prezeroed,
                                      const void* yprev, const double* pstats, double* psums,
                                      int n, int d, int h, int w, int gin, int gout, void* stream);
extern "C" int vs_conv3x3x3_tc_kdn(const void* x, const void* wkdn, void* y, double* stats, float* shift, int prezeroed,
                                   int n, int d, int h, int w, int gin, int gout, void* stream) {
    return vs_conv3x3x3_tc_kdn_ex(x, wkdn, y, stats, shift, prezeroed, nullptr, nullptr, nullptr, n, d, h, w, gin, gout, stream);
}

// As vs_conv3x3x3_tc_kdn; psums != NULL (dgrad): also accumulates the previous layer's InstanceNorm-backward sums
// psums[n][gout][2] += (sum g*mask, sum g*mask*xhat) from yprev / pstats (zeroed by the caller; no stats / shift then).
static int run_kdn(const void* x, const void* wkdn, void* y, double* stats, float* shift, int prezeroed,
                   const void* yprev, const double* pstats, double* psums, int planar_mode, float* yplanar, const float* bias,
                   int n, int d, int h, int w, int gin, int gout, void* stream,
                   void* a_out = nullptr, const void* skip = nullptr, unsigned* gbar = nullptr) {
    VS_REQUIRE(x && wkdn && (y || planar_mode), VS_ERR_SHAPE, "conv3_tc_kdn: null pointer");
    if (planar_mode) VS_REQUIRE(yplanar && gout == 8 && !stats && !shift && !psums, VS_ERR_SHAPE,
                                "conv3_tc_kdn: planar output needs Cout padded to 8 and no statistics");
    VS_REQUIRE((gin == 8 || (gin % 16 == 0 && gin >= 16)) && (gout == 8 || gout == 16) && d >= TD, VS_ERR_UNSUPPORTED,
               "conv3_tc_kdn: needs Cin = 8 or a multiple of 16, Cout in {8,16}, D >= 4 (Cin=%d Cout=%d D=%d)", gin, gout, d);
    VS_REQUIRE(vs_aligned16(x) && vs_aligned16(y) && vs_aligned16(wkdn), VS_ERR_ALIGN, "conv3_tc_kdn: pointers must be 16B aligned");
    EncodeTiledFn encode = get_encode_fn();
    VS_REQUIRE(encode != nullptr, VS_ERR_CUDA, "conv3_tc_kdn: cuTensorMapEncodeTiled unavailable");
    cudaStream_t st = (cudaStream_t)stream;
    const int tiles_h = (h + TH - 1) / TH, tiles_w = (w + TW - 1) / TW;
    const int bw = min(HW, w + 2), bh = min(HH, h + 2), bd = HD;
    const cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    CUtensorMap map;
    CUresult cr;
    if (gin == 8) {
        const cuuint64_t gdim4[4] = {(cuuint64_t)w * 8, (cuuint64_t)h, (cuuint64_t)d, (cuuint64_t)n};
        const cuuint64_t gstr4[3] = {(cuuint64_t)w * 16, (cuuint64_t)h * w * 16, (cuuint64_t)d * h * w * 16};
        const cuuint32_t box4[4] = {(cuuint32_t)bw * 8, (cuuint32_t)bh, (cuuint32_t)bd, 1};
        cr = encode(&map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(x), gdim4, gstr4, box4, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    } else {
        const cuuint64_t gdim[5] = {(cuuint64_t)gin, (cuuint64_t)w, (cuuint64_t)h, (cuuint64_t)d, (cuuint64_t)n};
        const cuuint64_t gstr[4] = {(cuuint64_t)gin * 2, (cuuint64_t)w * gin * 2, (cuuint64_t)h * w * gin * 2,
                                    (cuuint64_t)d * h * w * gin * 2};
        const cuuint32_t box[5] = {8, (cuuint32_t)bw, (cuuint32_t)bh, (cuuint32_t)bd, 1};
        cr = encode(&map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(x), gdim, gstr, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    }
    VS_REQUIRE(cr == CUDA_SUCCESS, VS_ERR_CUDA, "conv3_tc_kdn: cuTensorMapEncodeTiled failed (%d)", (int)cr);

    KdnParams p;
    p.n = n; p.d = d; p.h = h; p.w = w; p.cin = gin; p.cout = gout;
    p.tiles_h = tiles_h; p.tiles_w = tiles_w;
    p.tiles_d = (d + TD - 1) / TD;
    p.tiles_per_n = p.tiles_d * tiles_h * tiles_w;
    p.bw = bw; p.bh = bh; p.bd = bd;
    p.ref_tile = 0;                                                  // voxel (1,1,1) lies in tile (0,0,0) when D >= 4
    p.kslices = gin == 8 ? 1 : gin / 16;
    const long long items = (long long)n * p.tiles_per_n;
    VS_REQUIRE(items < 2147483647LL, VS_ERR_SHAPE, "conv3_tc_kdn: too many work items");
    p.work_items = (int)items;
    p.wpack = (const bf16*)wkdn; p.y = (bf16*)y; p.stats = stats; p.shift = shift;
    p.yprev = (const bf16*)yprev; p.pstats = pstats; p.psums = psums; p.inv_s = 1.0 / ((double)d * h * w);
    p.planar_mode = planar_mode; p.yplanar = yplanar; p.bias = bias;
    p.a_out = (bf16*)a_out; p.skip = (const bf16*)skip; p.gbar = gbar;
    if (a_out != nullptr)
        VS_REQUIRE(stats && gbar && prezeroed && !planar_mode && !psums && vs_aligned16(a_out) && vs_aligned16(skip) &&
                   (long long)n * gout * 8 <= 64 * 1024, VS_ERR_SHAPE,
                   "conv3_tc_kdn: the fused InstanceNorm+ReLU needs pre-zeroed statistics and a barrier word");
    p.dbg = g_kdn_dbg;
    p.ordered = g_kdn_ordered;
    if (psums != nullptr) {
        VS_REQUIRE(yprev && pstats && stats == nullptr && shift == nullptr, VS_ERR_SHAPE,
                   "conv3_tc_kdn: the fused norm-backward reduction needs y_prev + stats_prev and no forward statistics");
        VS_REQUIRE(vs_aligned16(yprev), VS_ERR_ALIGN, "conv3_tc_kdn: y_prev must be 16B aligned");
    }
    if (!prezeroed) {
        if (stats) VS_CUDA(cudaMemsetAsync(stats, 0, sizeof(double) * 2 * (size_t)n * gout, st), "conv3_tc_kdn stats memset");
        if (shift) VS_CUDA(cudaMemsetAsync(shift, 0, sizeof(float) * (size_t)n * gout, st), "conv3_tc_kdn shift memset");
    }
    if (gin == 8) {
        if (gout == 8) return launch_kdn<true, 8, 4>(map, p, st);
        return launch_kdn<true, 16, 4>(map, p, st);
    }
    if (gout == 8) return launch_kdn<false, 8, 4>(map, p, st);
    return launch_kdn<false, 16, 4>(map, p, st);
}

extern "C" int vs_conv3x3x3_tc_kdn_ex(const void* x, const void* wkdn, void* y, double* stats, float* shift, int prezeroed,
                                      const void* yprev, const double* pstats, double* psums,
                                      int n, int d, int h, int w, int gin, int gout, void* stream) {
    return run_kdn(x, wkdn, y, stats, shift, prezeroed, yprev, pstats, psums, 0, nullptr, nullptr, n, d, h, w, gin, gout, stream);
}

// Planar fp32 output of GEMM output channels 0..1 of an 8-channel kd-in-N convolution (pack built with
// vs_pack_conv3_weight_tc_kdn_padded): mode 1 = the 2-class head, out[N][2][D][H][W] = softmax(conv + bias)
// (joint_model.py:224-225,366-367); mode 2 = plain values (the VAE in-block's 2-channel planar input gradient).
extern "C" int vs_conv3x3x3_tc_kdn_planar(const void* x, const void* wkdn8, float* out, const float* bias, int mode,
                                          int n, int d, int h, int w, int gin, void* stream) {
    VS_REQUIRE(mode == 1 || mode == 2, VS_ERR_SHAPE, "conv3_tc_kdn_planar: mode must be 1 (softmax head) or 2 (plain)");
    return run_kdn(x, wkdn8, nullptr, nullptr, nullptr, 1, nullptr, nullptr, nullptr, mode, out, mode == 1 ? bias : nullptr,
                   n, d, h, w, gin, 8, stream);
}

// conv3 + InstanceNorm3d(eps 1e-5, biased) + ReLU (+ skip add) in ONE cooperative launch of the kd-in-N kernel
// (joint_model.py:40-46,106: Conv3d -> InstanceNorm3d -> ReLU): y = raw conv output (shifted, as vs_conv3x3x3_tc_kdn),
// a = relu((y - mean) * rstd) + skip.  stats / shift / gbar (one 32-bit word) must be zeroed by the caller.
extern "C" int vs_conv3x3x3_tc_kdn_in_relu(const void* x, const void* wkdn, void* y, void* a, const void* skip, double* stats,
                                           float* shift, unsigned* gbar, int n, int d, int h, int w, int gin, int gout,
                                           void* stream) {
    VS_REQUIRE(a != nullptr, VS_ERR_SHAPE, "conv3_tc_kdn_in_relu: null activation output");
    return run_kdn(x, wkdn, y, stats, shift, 1, nullptr, nullptr, nullptr, 0, nullptr, nullptr, n, d, h, w, gin, gout, stream,
                   a, skip, gbar);
}
