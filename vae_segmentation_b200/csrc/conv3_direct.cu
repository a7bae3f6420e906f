// 3x3x3 convolution (padding 1) on CUDA cores: fp32-accumulate direct kernels.
//
// Role: (a) the fp32 "check mode" path (1e-4 parity bar needs fp32 operands);
//       (b) the layers the tensor-core path does not take (Cin in {1,2}, Cout = 2);
//       (c) the wgrad contraction.
// Replaces cuDNN fprop / dgrad / wgrad reached from joint_model.py:40,43,46,106,224,366.
//
// Layout: activations NDHWC (T = float | bf16), or planar fp32 at the module boundary.
// Tile: 4x8x8 output voxels per CTA, 128 threads, halo 6x10x10 staged in shared memory as
// fp32 channel-quads so one LDS.128 feeds four input channels.
#include "vs_common.cuh"

namespace {

constexpr int TD = 4, TH = 8, TW = 8;
constexpr int HD = TD + 2, HH = TH + 2, HW = TW + 2;
constexpr int HV = HD * HH * HW;          // 600 halo voxels
constexpr int CIB = 8;                    // input channels per shared-memory stage
constexpr int NT = 128;

struct ConvDims { int n, d, h, w, cin, cout, tiles_d, tiles_h, tiles_w; };

__device__ __forceinline__ void decode_tile(int tile, const ConvDims& p, int& n, int& d0, int& h0, int& w0) {
    int tw = tile % p.tiles_w; tile /= p.tiles_w;
    int th = tile % p.tiles_h; tile /= p.tiles_h;
    int td = tile % p.tiles_d; n = tile / p.tiles_d;
    d0 = td * TD; h0 = th * TH; w0 = tw * TW;
}

// Stage the halo of `cib` channels starting at c0 into xs[plane][voxel] (fp32, zero padded).
template <typename TI, bool IN_PLANAR>
__device__ __forceinline__ void load_halo(float4 (*xs)[HV], const TI* __restrict__ x, const ConvDims& p,
                                          int n, int d0, int h0, int w0, int c0, int cib) {
    const long long S = (long long)p.d * p.h * p.w;
    for (int i = threadIdx.x; i < HV; i += NT) {
        int hw = i % HW, hh = (i / HW) % HH, hd = i / (HW * HH);
        int gd = d0 + hd - 1, gh = h0 + hh - 1, gw = w0 + hw - 1;
        float v[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        if (gd >= 0 && gd < p.d && gh >= 0 && gh < p.h && gw >= 0 && gw < p.w) {
            long long vox = ((long long)gd * p.h + gh) * p.w + gw;
            if (IN_PLANAR) {
                const float* xp = reinterpret_cast<const float*>(x);
                for (int c = 0; c < cib; ++c) v[c] = xp[((long long)n * p.cin + c0 + c) * S + vox];
            } else {
                const TI* px = x + ((long long)n * S + vox) * p.cin + c0;
                if (cib == 8 && (p.cin & 7) == 0) Store<TI>::ld8(px, v);
                else for (int c = 0; c < cib; ++c) v[c] = Store<TI>::ld(px + c);
            }
        }
        xs[0][i] = make_float4(v[0], v[1], v[2], v[3]);
        xs[1][i] = make_float4(v[4], v[5], v[6], v[7]);
    }
}

// y = conv3(x, wpk) (+bias), optional per-(n,c) sum / sum-of-squares of the fp32 accumulators.
// NCI > 0: the layer has exactly NCI (1 or 2) input channels (the in-blocks): only those channels are multiplied
// instead of a zero-padded channel quad.
template <typename TI, typename TO, int COB, bool IN_PLANAR, bool OUT_PLANAR, int NCI = 0>
__global__ void __launch_bounds__(NT) conv3_direct_kernel(const TI* __restrict__ x, const float* __restrict__ wpk,
                                                          const float* __restrict__ bias, TO* __restrict__ y,
                                                          double* __restrict__ stats,
                                                          const float* __restrict__ shift, ConvDims p, int ntiles) {
    __shared__ float4 xs[2][HV];
    __shared__ float4 ws[27][CIB][COB / 4];
    __shared__ double sred[COB][2];

    const int co0 = blockIdx.y * COB;
    const int t = threadIdx.x;
    const int tw = t & 7, th = (t >> 3) & 7, tdp = t >> 6;
    // Persistent over tiles (grid.x <= ntiles): the statistics of all tiles of one sample that this CTA visits are
    // accumulated in shared memory and flushed with ONE set of fp64 atomics per (CTA, sample) -- thousands of CTAs
    // adding onto the same 2*Cout words made the L2 atomic unit the bound of the full-resolution in-blocks.  With a
    // single input-channel stage (Cin <= 8) the weights are staged once per CTA.
    const bool one_stage = p.cin <= CIB;
    int stat_n = -1;
    if (stats != nullptr && t < COB * 2) sred[t >> 1][t & 1] = 0.0;
    auto flush_stats = [&]() {          // called by all threads
        __syncthreads();
        if (t < COB * 2 && stat_n >= 0) {
            if (co0 + (t >> 1) < p.cout)
                atomicAdd(&stats[((long long)stat_n * p.cout + co0 + (t >> 1)) * 2 + (t & 1)], sred[t >> 1][t & 1]);
            sred[t >> 1][t & 1] = 0.0;
        }
        __syncthreads();
    };

  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    int n, d0, h0, w0;
    decode_tile(tile, p, n, d0, h0, w0);
    if (stats != nullptr && n != stat_n) { flush_stats(); stat_n = n; }

    float acc0[COB], acc1[COB];
#pragma unroll
    for (int j = 0; j < COB; ++j) { acc0[j] = 0.f; acc1[j] = 0.f; }

    for (int c0 = 0; c0 < p.cin; c0 += CIB) {
        const int cib = min(CIB, p.cin - c0);
        __syncthreads();
        load_halo<TI, IN_PLANAR>(xs, x, p, n, d0, h0, w0, c0, cib);
        if (!one_stage || tile == (int)blockIdx.x)
        for (int i = t; i < 27 * CIB * (COB / 4); i += NT) {
            int j = i % (COB / 4), cc = (i / (COB / 4)) % CIB, tap = i / ((COB / 4) * CIB);
            float wv[4] = {0.f, 0.f, 0.f, 0.f};
            if (cc < cib) {
                const float* pw = wpk + ((long long)tap * p.cin + c0 + cc) * p.cout + co0 + j * 4;
#pragma unroll
                for (int q = 0; q < 4; ++q) if (co0 + j * 4 + q < p.cout) wv[q] = pw[q];
            }
            ws[tap][cc][j] = make_float4(wv[0], wv[1], wv[2], wv[3]);
        }
        __syncthreads();
        const int nplane = NCI > 0 ? 1 : (cib > 4 ? 2 : 1);
#pragma unroll 1
        for (int kd = 0; kd < 3; ++kd) {
#pragma unroll 1
            for (int kh = 0; kh < 3; ++kh) {
#pragma unroll
                for (int kw = 0; kw < 3; ++kw) {
                    const int tap = (kd * 3 + kh) * 3 + kw;
                    const int i0 = ((tdp + kd) * HH + th + kh) * HW + tw + kw;
                    const int i1 = i0 + 2 * HH * HW;
                    for (int pl = 0; pl < nplane; ++pl) {
                        const float4 a0 = xs[pl][i0], a1 = xs[pl][i1];
                        const float a0v[4] = {a0.x, a0.y, a0.z, a0.w};
                        const float a1v[4] = {a1.x, a1.y, a1.z, a1.w};
#pragma unroll
                        for (int cc = 0; cc < (NCI > 0 ? NCI : 4); ++cc) {
#pragma unroll
                            for (int j = 0; j < COB / 4; ++j) {
                                const float4 wv = ws[tap][pl * 4 + cc][j];
                                acc0[4 * j + 0] = fmaf(a0v[cc], wv.x, acc0[4 * j + 0]);
                                acc0[4 * j + 1] = fmaf(a0v[cc], wv.y, acc0[4 * j + 1]);
                                acc0[4 * j + 2] = fmaf(a0v[cc], wv.z, acc0[4 * j + 2]);
                                acc0[4 * j + 3] = fmaf(a0v[cc], wv.w, acc0[4 * j + 3]);
                                acc1[4 * j + 0] = fmaf(a1v[cc], wv.x, acc1[4 * j + 0]);
                                acc1[4 * j + 1] = fmaf(a1v[cc], wv.y, acc1[4 * j + 1]);
                                acc1[4 * j + 2] = fmaf(a1v[cc], wv.z, acc1[4 * j + 2]);
                                acc1[4 * j + 3] = fmaf(a1v[cc], wv.w, acc1[4 * j + 3]);
                            }
                        }
                    }
                }
            }
        }
    }

    // ---- epilogue ----
    const int gh = h0 + th, gw = w0 + tw;
    const int gd0 = d0 + tdp, gd1 = d0 + tdp + 2;
    const bool v0 = gd0 < p.d && gh < p.h && gw < p.w;
    const bool v1 = gd1 < p.d && gh < p.h && gw < p.w;
    if (bias != nullptr) {
#pragma unroll
        for (int j = 0; j < COB; ++j) {
            float b = (co0 + j < p.cout) ? bias[co0 + j] : 0.f;
            acc0[j] += b; acc1[j] += b;
        }
    }
    if (shift != nullptr) {
#pragma unroll
        for (int j = 0; j < COB; ++j) {
            float k = (co0 + j < p.cout) ? shift[(long long)n * p.cout + co0 + j] : 0.f;
            acc0[j] -= k; acc1[j] -= k;
        }
    }
    if (stats != nullptr) {
#pragma unroll
        for (int j = 0; j < COB; ++j) {
            const double a = v0 ? (double)acc0[j] : 0.0, b = v1 ? (double)acc1[j] : 0.0;
            const double s = warp_sum(a + b);
            const double q = warp_sum(a * a + b * b);
            if ((t & 31) == 0) { atomicAdd(&sred[j][0], s); atomicAdd(&sred[j][1], q); }
        }
    }
    const long long S = (long long)p.d * p.h * p.w;
#pragma unroll
    for (int sel = 0; sel < 2; ++sel) {
        const bool valid = sel ? v1 : v0;
        if (!valid) continue;
        const float* acc = sel ? acc1 : acc0;
        const long long vox = ((long long)(sel ? gd1 : gd0) * p.h + gh) * p.w + gw;
        if (OUT_PLANAR) {
            float* yp = reinterpret_cast<float*>(y);
#pragma unroll
            for (int j = 0; j < COB; ++j)
                if (co0 + j < p.cout) yp[((long long)n * p.cout + co0 + j) * S + vox] = acc[j];
        } else {
            TO* py = y + ((long long)n * S + vox) * p.cout + co0;
            if constexpr (COB % 8 == 0) {
#pragma unroll
                for (int j8 = 0; j8 < COB / 8; ++j8) {
                    float v[8];
#pragma unroll
                    for (int q = 0; q < 8; ++q) v[q] = acc[j8 * 8 + q];
                    Store<TO>::st8(py + j8 * 8, v);
                }
            } else {
#pragma unroll
                for (int j = 0; j < COB; ++j) if (co0 + j < p.cout) Store<TO>::st(py + j, acc[j]);
            }
        }
    }
  }   // tile loop
    if (stats != nullptr) flush_stats();
}


// ---------------------------------------------------------------------------------------------
// In-block kernel: Conv3d(NCI -> 8) on the planar fp32 module input (NCI = 1: Segmentation, 2: VAE), NDHWC output.
// FMA-bound by construction: a thread owns 4 consecutive w-voxels x 8 output channels; per (kd, kh, ci) it reads the
// 6 input values it needs from a conflict-free shared halo (row pitch 37 words) and the 3 x 8 weights as broadcast
// float4s -- 12 shared loads per 96 FMAs.  Tile 4 x 8 x 32 voxels (w contiguous: coalesced planar loads, 64-byte
// stores per thread), persistent CTAs, statistics flushed once per (CTA, sample).
// ---------------------------------------------------------------------------------------------
constexpr int IB_TD = 4, IB_TH = 8, IB_TW = 32, IB_NT = 256;
constexpr int IB_HD = IB_TD + 2, IB_HH = IB_TH + 2, IB_HW = IB_TW + 2, IB_PITCH = 37;

template <typename TO, int NCI>
__global__ void __launch_bounds__(IB_NT, 2) conv3_inblock_kernel(const float* __restrict__ x, const float* __restrict__ wpk,
                                                              const float* __restrict__ bias, TO* __restrict__ y,
                                                              double* __restrict__ stats, const float* __restrict__ shift,
                                                              ConvDims p, int tiles_w32, int tiles_h8, int ntiles) {
    __shared__ float xs[NCI][IB_HD][IB_HH][IB_PITCH];
    __shared__ float4 ws[27][NCI][2];
    __shared__ double sred[8][2];
    __shared__ float sconst[8];                     // bias - shift of the current sample
    const int t = threadIdx.x;
    const int lw = (t & 7) * 4, lh = (t >> 3) & 7, ld = t >> 6;
    const long long S = (long long)p.d * p.h * p.w;
    for (int i = t; i < 27 * NCI * 2; i += IB_NT) {
        const int j = i & 1, ci = (i >> 1) % NCI, tap = (i >> 1) / NCI;
        const float* pw = wpk + ((long long)tap * NCI + ci) * 8 + j * 4;          // wpk[27][Cin][Cout = 8]
        ws[tap][ci][j] = make_float4(pw[0], pw[1], pw[2], pw[3]);
    }
    if (t < 16) sred[t >> 1][t & 1] = 0.0;
    int stat_n = -1;
    const int tiles_per_n = p.tiles_d * tiles_h8 * tiles_w32;
    // The halo of the NEXT tile is fetched into registers while the current tile is computed (all loads of a thread
    // in flight at once): with load-then-compute the global latency of a dozen dependent rounds per tile, not the
    // FMAs, set the time.
    constexpr int NEL = NCI * IB_HD * IB_HH * IB_HW, NIT = (NEL + IB_NT - 1) / IB_NT;
    float pre[NIT];
    auto prefetch = [&](int tile) {
        const int n = tile / tiles_per_n;
        int r = tile - n * tiles_per_n;
        const int w0 = (r % tiles_w32) * IB_TW; r /= tiles_w32;
        const int h0 = (r % tiles_h8) * IB_TH; r /= tiles_h8;
        const int d0 = r * IB_TD;
#pragma unroll
        for (int k = 0; k < NIT; ++k) {
            const int i = t + k * IB_NT;
            const int hw = i % IB_HW, hh = (i / IB_HW) % IB_HH, hd = (i / (IB_HW * IB_HH)) % IB_HD, ci = i / (IB_HW * IB_HH * IB_HD);
            const int gd = d0 + hd - 1, gh = h0 + hh - 1, gw = w0 + hw - 1;
            float v = 0.f;
            if (i < NEL && gd >= 0 && gd < p.d && gh >= 0 && gh < p.h && gw >= 0 && gw < p.w)
                v = __ldg(x + ((long long)n * NCI + ci) * S + ((long long)gd * p.h + gh) * p.w + gw);
            pre[k] = v;
        }
    };
    if ((int)blockIdx.x < ntiles) prefetch(blockIdx.x);
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int n = tile / tiles_per_n;
        int r = tile - n * tiles_per_n;
        const int w0 = (r % tiles_w32) * IB_TW; r /= tiles_w32;
        const int h0 = (r % tiles_h8) * IB_TH; r /= tiles_h8;
        const int d0 = r * IB_TD;
        __syncthreads();                            // previous tile's reads of xs / sconst are done
        if (n != stat_n) {
            if (stats != nullptr && stat_n >= 0 && t < 16) {
                atomicAdd(&stats[((long long)stat_n * 8 + (t >> 1)) * 2 + (t & 1)], sred[t >> 1][t & 1]);
                sred[t >> 1][t & 1] = 0.0;
            }
            if (t < 8) sconst[t] = (bias != nullptr ? bias[t] : 0.f) - (shift != nullptr ? shift[(long long)n * 8 + t] : 0.f);
            stat_n = n;
        }
#pragma unroll
        for (int k = 0; k < NIT; ++k) {
            const int i = t + k * IB_NT;
            const int hw = i % IB_HW, hh = (i / IB_HW) % IB_HH, hd = (i / (IB_HW * IB_HH)) % IB_HD, ci = i / (IB_HW * IB_HH * IB_HD);
            if (i < NEL) xs[ci][hd][hh][hw] = pre[k];
        }
        __syncthreads();
        if (tile + (int)gridDim.x < ntiles) prefetch(tile + gridDim.x);
        float acc[4][8];
#pragma unroll
        for (int v = 0; v < 4; ++v)
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[v][j] = 0.f;
#pragma unroll 1
        for (int kd = 0; kd < 3; ++kd) {
#pragma unroll
            for (int kh = 0; kh < 3; ++kh) {
#pragma unroll
                for (int ci = 0; ci < NCI; ++ci) {
                    const float* row = &xs[ci][ld + kd][lh + kh][lw];
                    float xv[6];
#pragma unroll
                    for (int k = 0; k < 6; ++k) xv[k] = row[k];
#pragma unroll
                    for (int kw = 0; kw < 3; ++kw) {
                        const float4 wa = ws[(kd * 3 + kh) * 3 + kw][ci][0], wb = ws[(kd * 3 + kh) * 3 + kw][ci][1];
                        const float wv[8] = {wa.x, wa.y, wa.z, wa.w, wb.x, wb.y, wb.z, wb.w};
#pragma unroll
                        for (int v = 0; v < 4; ++v)
#pragma unroll
                            for (int j = 0; j < 8; ++j) acc[v][j] = fmaf(xv[v + kw], wv[j], acc[v][j]);
                    }
                }
            }
        }
        // ---- epilogue: (+bias) - shift, statistics, store ----
        const int gd = d0 + ld, gh = h0 + lh;
        const bool row_ok = gd < p.d && gh < p.h;
        float ssum[8], ssq[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) { ssum[j] = 0.f; ssq[j] = 0.f; }
#pragma unroll
        for (int v = 0; v < 4; ++v) {
            const int gw = w0 + lw + v;
            const bool ok = row_ok && gw < p.w;
            float o[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                o[j] = acc[v][j] + sconst[j];
                if (ok) { ssum[j] += o[j]; ssq[j] = fmaf(o[j], o[j], ssq[j]); }
            }
            if (ok) Store<TO>::st8(y + ((long long)n * S + ((long long)gd * p.h + gh) * p.w + gw) * 8, o);
        }
        if (stats != nullptr) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const double s1 = warp_sum((double)ssum[j]), s2 = warp_sum((double)ssq[j]);
                if ((t & 31) == 0) { atomicAdd(&sred[j][0], s1); atomicAdd(&sred[j][1], s2); }
            }
        }
    }
    __syncthreads();
    if (stats != nullptr && stat_n >= 0 && t < 16)
        atomicAdd(&stats[((long long)stat_n * 8 + (t >> 1)) * 2 + (t & 1)], sred[t >> 1][t & 1]);
}

// shift[n][co] = conv3(x, w) at voxel (1,1,1).  One CTA per (n, 8 output channels): the 27*Cin input values of
// the reference neighbourhood are staged in shared memory, thread = (co, one of 32 term partitions) with
// four independent accumulators so the weight loads pipeline; deterministic tree reduction.
template <typename TI, bool IN_PLANAR>
__global__ void __launch_bounds__(256) conv3_shift_kernel(const TI* __restrict__ x, const float* __restrict__ wpk,
                                                          float* __restrict__ shift, ConvDims p) {
    extern __shared__ float xs_ref[];          // [27 * cin]
    __shared__ float red[32][9];
    const int n = blockIdx.y, t = threadIdx.x;
    const long long S = (long long)p.d * p.h * p.w;
    const int terms = 27 * p.cin;
    for (int i = t; i < terms; i += 256) {
        const int tap = i / p.cin, ci = i - tap * p.cin;
        const int gd = tap / 9, gh = (tap / 3) % 3, gw = tap % 3;       // voxel (1,1,1) + tap - 1
        float xv = 0.f;
        if (gd < p.d && gh < p.h && gw < p.w) {
            const long long vox = ((long long)gd * p.h + gh) * p.w + gw;
            if (IN_PLANAR) xv = reinterpret_cast<const float*>(x)[((long long)n * p.cin + ci) * S + vox];
            else xv = Store<TI>::ld(x + ((long long)n * S + vox) * p.cin + ci);
        }
        xs_ref[i] = xv;
    }
    __syncthreads();
    const int col = t & 7, part = t >> 3;
    const int co = blockIdx.x * 8 + col;
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
    if (co < p.cout) {
        const float* pw = wpk + co;
        int i = part;
        for (; i + 96 < terms; i += 128) {
            a0 = fmaf(xs_ref[i], pw[(long long)i * p.cout], a0);
            a1 = fmaf(xs_ref[i + 32], pw[(long long)(i + 32) * p.cout], a1);
            a2 = fmaf(xs_ref[i + 64], pw[(long long)(i + 64) * p.cout], a2);
            a3 = fmaf(xs_ref[i + 96], pw[(long long)(i + 96) * p.cout], a3);
        }
        for (; i < terms; i += 32) a0 = fmaf(xs_ref[i], pw[(long long)i * p.cout], a0);
    }
    red[part][col] = (a0 + a1) + (a2 + a3);
    __syncthreads();
    if (t < 8 && blockIdx.x * 8 + t < p.cout) {
        float s = 0.f;
#pragma unroll
        for (int k = 0; k < 32; ++k) s += red[k][t];
        shift[(long long)n * p.cout + blockIdx.x * 8 + t] = s;
    }
}

// ---- wgrad: dw[co][ci][tap] += sum_v dy[v,co] * x[v+tap,ci] ------------------------------
// Thread = one (ci,co) pair of the CTA's chunk, 27 tap accumulators in registers, sliding
// three-wide register window along w.  CTAs are persistent over voxel tiles so the final
// atomics are O(grid), not O(tiles).
template <typename T, int WCIB, int WCOB, bool IN_PLANAR>
__global__ void __launch_bounds__(NT) conv3_wgrad_kernel(const T* __restrict__ x, const T* __restrict__ dy,
                                                         float* __restrict__ dw, float* __restrict__ db,
                                                         ConvDims p, int total_tiles) {
    constexpr int PAIRS = WCIB * WCOB;
    constexpr int G = NT / PAIRS;             // row groups
    static_assert(NT % PAIRS == 0, "pairs must divide the block");
    __shared__ float xs[HV][WCIB];
    __shared__ float dys[TD * TH * TW][WCOB];

    const int t = threadIdx.x;
    const int pair = t % PAIRS, grp = t / PAIRS;
    const int ci_l = pair / WCOB, co_l = pair % WCOB;
    const int c0 = blockIdx.y * WCIB, co0 = blockIdx.z * WCOB;
    const long long S = (long long)p.d * p.h * p.w;

    float acc[27];
#pragma unroll
    for (int i = 0; i < 27; ++i) acc[i] = 0.f;
    float gsum = 0.f;

    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        int n, d0, h0, w0;
        decode_tile(tile, p, n, d0, h0, w0);
        __syncthreads();
        for (int i = t; i < HV * WCIB; i += NT) {
            int c = i % WCIB, hv = i / WCIB;
            int hw = hv % HW, hh = (hv / HW) % HH, hd = hv / (HW * HH);
            int gd = d0 + hd - 1, gh = h0 + hh - 1, gw = w0 + hw - 1;
            float v = 0.f;
            if (c0 + c < p.cin && gd >= 0 && gd < p.d && gh >= 0 && gh < p.h && gw >= 0 && gw < p.w) {
                long long vox = ((long long)gd * p.h + gh) * p.w + gw;
                if (IN_PLANAR) v = reinterpret_cast<const float*>(x)[((long long)n * p.cin + c0 + c) * S + vox];
                else v = Store<T>::ld(x + ((long long)n * S + vox) * p.cin + c0 + c);
            }
            xs[hv][c] = v;
        }
        for (int i = t; i < TD * TH * TW * WCOB; i += NT) {
            int c = i % WCOB, v = i / WCOB;
            int lw = v % TW, lh = (v / TW) % TH, ld = v / (TW * TH);
            int gd = d0 + ld, gh = h0 + lh, gw = w0 + lw;
            float g = 0.f;
            if (co0 + c < p.cout && gd < p.d && gh < p.h && gw < p.w)
                g = Store<T>::ld(dy + ((long long)n * S + ((long long)gd * p.h + gh) * p.w + gw) * p.cout + co0 + c);
            dys[v][c] = g;
        }
        __syncthreads();
#pragma unroll 1
        for (int row = grp; row < TD * TH; row += G) {
            const int ld = row / TH, lh = row % TH;
            float xw[9][3];
#pragma unroll
            for (int r = 0; r < 9; ++r) {
                const int base = ((ld + r / 3) * HH + lh + r % 3) * HW;
                xw[r][0] = xs[base][ci_l];
                xw[r][1] = xs[base + 1][ci_l];
            }
#pragma unroll
            for (int lw = 0; lw < TW; ++lw) {
                const float g = dys[row * TW + lw][co_l];
                gsum += g;
#pragma unroll
                for (int r = 0; r < 9; ++r) {
                    const int base = ((ld + r / 3) * HH + lh + r % 3) * HW;
                    xw[r][2] = xs[base + lw + 2][ci_l];
                    acc[r * 3 + 0] = fmaf(g, xw[r][0], acc[r * 3 + 0]);
                    acc[r * 3 + 1] = fmaf(g, xw[r][1], acc[r * 3 + 1]);
                    acc[r * 3 + 2] = fmaf(g, xw[r][2], acc[r * 3 + 2]);
                    xw[r][0] = xw[r][1];
                    xw[r][1] = xw[r][2];
                }
            }
        }
    }
    const int ci = c0 + ci_l, co = co0 + co_l;
    if (ci < p.cin && co < p.cout) {
        float* pd = dw + ((long long)co * p.cin + ci) * 27;
#pragma unroll
        for (int i = 0; i < 27; ++i) atomicAdd(pd + i, acc[i]);
        if (db != nullptr && ci == 0) atomicAdd(db + co, gsum);
    }
}

// w[Cout][Cin][27] -> wf[27][Cin][Cout], wd[26-tap][Cout][Cin]
__global__ void pack_conv3_weight_kernel(const float* __restrict__ w, float* __restrict__ wf, float* __restrict__ wd,
                                         int cin, int cout) {
    const int total = cout * cin * 27;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        int tap = i % 27, ci = (i / 27) % cin, co = i / (27 * cin);
        float v = w[i];
        wf[((long long)tap * cin + ci) * cout + co] = v;
        if (wd != nullptr) wd[((long long)(26 - tap) * cout + co) * cin + ci] = v;
    }
}

int g_inblock_kernel = 1;          // development switch: 0 routes the in-blocks through the generic direct kernel

ConvDims make_dims(int n, int d, int h, int w, int cin, int cout) {
    ConvDims p;
    p.n = n; p.d = d; p.h = h; p.w = w; p.cin = cin; p.cout = cout;
    p.tiles_d = (d + TD - 1) / TD; p.tiles_h = (h + TH - 1) / TH; p.tiles_w = (w + TW - 1) / TW;
    return p;
}

template <typename TI, typename TO, bool IN_PLANAR, bool OUT_PLANAR>
int launch_conv3(const void* x, const float* wpk, const float* bias, void* y, double* stats, float* shift,
                 const ConvDims& p, cudaStream_t st) {
    if (shift != nullptr) {
        dim3 sgrid((p.cout + 7) / 8, p.n);
        conv3_shift_kernel<TI, IN_PLANAR><<<sgrid, 256, 27 * p.cin * sizeof(float), st>>>((const TI*)x, wpk, shift, p);
        VS_CHECK_LAUNCH("conv3_shift_kernel");
    }
    // bf16 mode only: the fp32 check mode keeps the generic kernel (fp64 per-element statistics, the summation
    // order its 1e-4 / gradient bounds were calibrated with)
    if constexpr (IN_PLANAR && !OUT_PLANAR && sizeof(TO) == 2) {
        if (p.cout == 8 && (p.cin == 1 || p.cin == 2) && g_inblock_kernel) {
            const int tw32 = (p.w + IB_TW - 1) / IB_TW, th8 = (p.h + IB_TH - 1) / IB_TH;
            const long long nt = (long long)p.n * p.tiles_d * th8 * tw32;            // tiles_d: TD = IB_TD = 4
            VS_REQUIRE(nt < 2147483647LL, VS_ERR_SHAPE, "conv3: too many tiles");
            const unsigned grid = (unsigned)min(nt, (long long)vs_sm_count() * 2);   // 2 resident CTAs per SM (launch bounds)
            if (p.cin == 1)
                conv3_inblock_kernel<TO, 1><<<grid, IB_NT, 0, st>>>((const float*)x, wpk, bias, (TO*)y, stats, shift, p, tw32, th8, (int)nt);
            else
                conv3_inblock_kernel<TO, 2><<<grid, IB_NT, 0, st>>>((const float*)x, wpk, bias, (TO*)y, stats, shift, p, tw32, th8, (int)nt);
            VS_CHECK_LAUNCH("conv3_inblock_kernel");
            return VS_OK;
        }
    }
    const long long tiles = (long long)p.n * p.tiles_d * p.tiles_h * p.tiles_w;
    VS_REQUIRE(tiles < 2147483647LL, VS_ERR_SHAPE, "conv3: too many tiles");
    const int cob = p.cout >= 16 ? 16 : (p.cout >= 8 ? 8 : 4);
    const int ntiles = (int)tiles;
    // persistent CTAs (see the kernel): ~8 resident CTAs per SM by shared memory (26-33 KB each)
    dim3 grid((unsigned)min(tiles, (long long)vs_sm_count() * 8), (p.cout + cob - 1) / cob);
    if (cob == 16)
        conv3_direct_kernel<TI, TO, 16, IN_PLANAR, OUT_PLANAR><<<grid, NT, 0, st>>>((const TI*)x, wpk, bias, (TO*)y, stats, shift, p, ntiles);
    else if (cob == 8 && p.cin == 1)
        conv3_direct_kernel<TI, TO, 8, IN_PLANAR, OUT_PLANAR, 1><<<grid, NT, 0, st>>>((const TI*)x, wpk, bias, (TO*)y, stats, shift, p, ntiles);
    else if (cob == 8 && p.cin == 2)
        conv3_direct_kernel<TI, TO, 8, IN_PLANAR, OUT_PLANAR, 2><<<grid, NT, 0, st>>>((const TI*)x, wpk, bias, (TO*)y, stats, shift, p, ntiles);
    else if (cob == 8)
        conv3_direct_kernel<TI, TO, 8, IN_PLANAR, OUT_PLANAR><<<grid, NT, 0, st>>>((const TI*)x, wpk, bias, (TO*)y, stats, shift, p, ntiles);
    else
        conv3_direct_kernel<TI, TO, 4, IN_PLANAR, OUT_PLANAR><<<grid, NT, 0, st>>>((const TI*)x, wpk, bias, (TO*)y, stats, shift, p, ntiles);
    VS_CHECK_LAUNCH("conv3_direct_kernel");
    return VS_OK;
}

int conv3_common(int in_dtype, int out_dtype, int in_planar, int out_planar, const void* x, const float* wpk,
                 const float* bias, void* y, double* stats, float* shift, int n, int d, int h, int w, int cin, int cout,
                 void* stream) {
    const bool prezeroed = (out_planar & 2) != 0;       // internal: bit 1 of out_planar = statistics already zeroed
    out_planar &= 1;
    VS_REQUIRE(n > 0 && d > 0 && h > 0 && w > 0 && cin > 0 && cout > 0, VS_ERR_SHAPE, "conv3: bad shape");
    VS_REQUIRE(x && wpk && y, VS_ERR_SHAPE, "conv3: null pointer");
    VS_REQUIRE(vs_aligned16(x) && vs_aligned16(y) && vs_aligned16(wpk), VS_ERR_ALIGN, "conv3: pointers must be 16B aligned");
    if (!out_planar) VS_REQUIRE(cout % 8 == 0 || cout == 2, VS_ERR_UNSUPPORTED, "conv3: NDHWC output needs Cout %% 8 == 0 or Cout == 2 (got %d)", cout);
    if (in_planar) VS_REQUIRE(in_dtype == VS_F32, VS_ERR_UNSUPPORTED, "conv3: planar input is fp32 only");
    if (out_planar) VS_REQUIRE(out_dtype == VS_F32, VS_ERR_UNSUPPORTED, "conv3: planar output is fp32 only");
    cudaStream_t st = (cudaStream_t)stream;
    ConvDims p = make_dims(n, d, h, w, cin, cout);
    if (stats && !prezeroed) VS_CUDA(cudaMemsetAsync(stats, 0, sizeof(double) * 2 * n * cout, st), "conv3 stats memset");
    if (in_planar) {
        VS_REQUIRE(!out_planar, VS_ERR_UNSUPPORTED, "conv3: planar->planar unsupported");
        if (out_dtype == VS_F32) return launch_conv3<float, float, true, false>(x, wpk, bias, y, stats, shift, p, st);
        return launch_conv3<float, bf16, true, false>(x, wpk, bias, y, stats, shift, p, st);
    }
    if (out_planar) {
        if (in_dtype == VS_F32) return launch_conv3<float, float, false, true>(x, wpk, bias, y, stats, shift, p, st);
        return launch_conv3<bf16, float, false, true>(x, wpk, bias, y, stats, shift, p, st);
    }
    if (in_dtype == VS_F32 && out_dtype == VS_F32) return launch_conv3<float, float, false, false>(x, wpk, bias, y, stats, shift, p, st);
    if (in_dtype == VS_BF16 && out_dtype == VS_BF16) return launch_conv3<bf16, bf16, false, false>(x, wpk, bias, y, stats, shift, p, st);
    if (in_dtype == VS_BF16 && out_dtype == VS_F32) return launch_conv3<bf16, float, false, false>(x, wpk, bias, y, stats, shift, p, st);
    VS_FAIL(VS_ERR_UNSUPPORTED, "conv3: dtype combination in=%d out=%d unsupported", in_dtype, out_dtype);
}

template <typename T, bool IN_PLANAR>
int launch_wgrad(const void* x, const void* dy, float* dw, float* db, const ConvDims& p, cudaStream_t st) {
    const long long tiles = (long long)p.n * p.tiles_d * p.tiles_h * p.tiles_w;
    const int sms = vs_sm_count();
    if (p.cin >= 8 && p.cout >= 16) {
        dim3 grid(1, (p.cin + 7) / 8, (p.cout + 15) / 16);
        long long slots = (long long)sms * 6 / ((long long)grid.y * grid.z);
        grid.x = (unsigned)max(1LL, min(tiles, slots));
        conv3_wgrad_kernel<T, 8, 16, IN_PLANAR><<<grid, NT, 0, st>>>((const T*)x, (const T*)dy, dw, db, p, (int)tiles);
    } else if (p.cin >= 8 && p.cout >= 8) {
        dim3 grid(1, (p.cin + 7) / 8, (p.cout + 7) / 8);
        long long slots = (long long)sms * 6 / ((long long)grid.y * grid.z);
        grid.x = (unsigned)max(1LL, min(tiles, slots));
        conv3_wgrad_kernel<T, 8, 8, IN_PLANAR><<<grid, NT, 0, st>>>((const T*)x, (const T*)dy, dw, db, p, (int)tiles);
    } else if (p.cin >= 8) {                       // out_block: Cout = 2
        dim3 grid(1, (p.cin + 7) / 8, (p.cout + 1) / 2);
        grid.x = (unsigned)max(1LL, min(tiles, (long long)sms * 6));
        conv3_wgrad_kernel<T, 8, 2, IN_PLANAR><<<grid, NT, 0, st>>>((const T*)x, (const T*)dy, dw, db, p, (int)tiles);
    } else {                                       // in_blocks: Cin in {1,2}
        dim3 grid(1, (p.cin + 1) / 2, (p.cout + 7) / 8);
        grid.x = (unsigned)max(1LL, min(tiles, (long long)sms * 6));
        conv3_wgrad_kernel<T, 2, 8, IN_PLANAR><<<grid, NT, 0, st>>>((const T*)x, (const T*)dy, dw, db, p, (int)tiles);
    }
    VS_CHECK_LAUNCH("conv3_wgrad_kernel");
    return VS_OK;
}

}  // namespace

extern "C" int vs_pack_conv3_weight(const float* w, float* wf, float* wd, int cin, int cout, void* stream) {
    VS_REQUIRE(w && wf && cin > 0 && cout > 0, VS_ERR_SHAPE, "pack_conv3_weight: bad arguments");
    const int total = cin * cout * 27;
    pack_conv3_weight_kernel<<<min(1024, (total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(w, wf, wd, cin, cout);
    VS_CHECK_LAUNCH("pack_conv3_weight_kernel");
    return VS_OK;
}

extern "C" int vs_conv3x3x3_fprop_direct(int in_dtype, int out_dtype, int in_planar, int out_planar, const void* x,
                                         const float* wpk, const float* bias, void* y, double* stats, float* shift,
                                         int n, int d, int h, int w, int cin, int cout, void* stream) {
    return conv3_common(in_dtype, out_dtype, in_planar, out_planar, x, wpk, bias, y, stats, shift, n, d, h, w, cin, cout,
                        stream);
}

// internal: per-(n,co) reference-voxel value used as the InstanceNorm-invariant shift (see header)
extern "C" int vs_conv3_shift_internal(int in_dtype, int in_planar, const void* x, const float* wpk, float* shift, int n,
                                       int d, int h, int w, int cin, int cout, void* stream) {
    ConvDims p = make_dims(n, d, h, w, cin, cout);
    dim3 sgrid((cout + 7) / 8, n);
    cudaStream_t st = (cudaStream_t)stream;
    const size_t sm = 27 * (size_t)cin * sizeof(float);
    if (in_planar) conv3_shift_kernel<float, true><<<sgrid, 256, sm, st>>>((const float*)x, wpk, shift, p);
    else if (in_dtype == VS_F32) conv3_shift_kernel<float, false><<<sgrid, 256, sm, st>>>((const float*)x, wpk, shift, p);
    else conv3_shift_kernel<bf16, false><<<sgrid, 256, sm, st>>>((const bf16*)x, wpk, shift, p);
    VS_CHECK_LAUNCH("conv3_shift_kernel");
    return VS_OK;
}

extern "C" size_t vs_conv3_wgrad_workspace_bytes(int, int, int, int, int, int) { return 0; }

#ifdef VS_WITH_TCGEN05
extern "C" int vs_conv3_wgrad_tc_eligible(int cin, int cout);
extern "C" int vs_conv3x3x3_wgrad_tc(const void* x, const void* dy, float* dw, int n, int d, int h, int w, int cin, int cout,
                                     void* stream);
#endif
extern "C" void vs_debug_set_inblock_kernel(int on) { g_inblock_kernel = on; }
static int g_wgrad_tc = 1;
// development switch (tools/kbench.py A/B runs): 0 forces the CUDA-core wgrad kernel
extern "C" void vs_debug_set_wgrad_tc(int on) { g_wgrad_tc = on; }

extern "C" int vs_conv3x3x3_wgrad(int dtype, int in_planar, const void* x, const void* dy, float* dw, float* db,
                                  void* workspace, size_t ws_bytes, int accumulate, int n, int d, int h, int w,
                                  int cin, int cout, void* stream) {
    (void)workspace; (void)ws_bytes;
    VS_REQUIRE(n > 0 && d > 0 && h > 0 && w > 0 && cin > 0 && cout > 0, VS_ERR_SHAPE, "conv3 wgrad: bad shape");
    VS_REQUIRE(x && dy && dw, VS_ERR_SHAPE, "conv3 wgrad: null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    ConvDims p = make_dims(n, d, h, w, cin, cout);
    if (!accumulate) {
        VS_CUDA(cudaMemsetAsync(dw, 0, sizeof(float) * 27 * cin * cout, st), "wgrad memset");
        if (db) VS_CUDA(cudaMemsetAsync(db, 0, sizeof(float) * cout, st), "wgrad memset db");
    }
#ifdef VS_WITH_TCGEN05
    // bf16 NDHWC operands without a bias gradient -> tcgen05 kernel (conv3_wgrad_tc.cu)
    if (g_wgrad_tc && dtype == VS_BF16 && !in_planar && db == nullptr && vs_conv3_wgrad_tc_eligible(cin, cout))
        return vs_conv3x3x3_wgrad_tc(x, dy, dw, n, d, h, w, cin, cout, stream);
#endif
    if (in_planar) {
        if (dtype == VS_F32) return launch_wgrad<float, true>(x, dy, dw, db, p, st);
        return launch_wgrad<bf16, true>(x, dy, dw, db, p, st);
    }
    if (dtype == VS_F32) return launch_wgrad<float, false>(x, dy, dw, db, p, st);
    if (dtype == VS_BF16) return launch_wgrad<bf16, false>(x, dy, dw, db, p, st);
    VS_FAIL(VS_ERR_UNSUPPORTED, "conv3 wgrad: unknown dtype %d", dtype);
}
