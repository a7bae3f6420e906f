// Element functions of the derived bf16 weight packs (UMMA B operands) shared by the per-layer pack kernels and the
// batched re-pack that follows every optimiser step (conv3_tc.cu: pack_batched_kernel).
#pragma once
#include "vs_common.cuh"

namespace {

// Output channels per CTA (MMA N).  An MMA costs ~(128 + N)/4 cycles (operand delivery from shared memory), so a
// small N wastes tensor-pipe time in aggregate but shortens each CTA: layers with >= 64 output channels only occur at
// the deep levels (<= 12^3), whose grids leave most SMs idle, so they are split into more, shorter CTAs.
#ifndef VS_NC_WIDE
#define VS_NC_WIDE 16
#endif
__host__ __device__ constexpr int nc_for_dev(int gout) { return gout >= 64 ? VS_NC_WIDE : (gout >= 32 ? 32 : 16); }

// One element of the bf16 UMMA B-operand pack.  cout_real < cout_l zero-pads the output channels (the head's
// 2 -> 8 channel dgrad pack).
__device__ __forceinline__ float pack_tc_elem(const float* __restrict__ w, long long i, int cin_l, int cout_l, int cout_real,
                                              int dgrad, int nc, int cin8, int cin_real = -1) {
    if (cin_real < 0) cin_real = cin_l;         // cin_real < cin_l zero-pads the input channels (2-channel in-block dgrad)
    // GEMM-side channel counts
    const int gin = dgrad ? cout_l : cin_l, gout = dgrad ? cin_l : cout_l;
    const int kslices = cin8 ? 1 : gin / 16;
    const int nmma = cin8 ? 14 : 27;
    long long r = i;
    const int ch8 = (int)(r % 8); r /= 8;
    const int r8 = (int)(r % 8); r /= 8;
    const int ng = (int)(r % (nc / 8)); r /= (nc / 8);
    const int kc = (int)(r % 2); r /= 2;
    const int m = (int)(r % nmma); r /= nmma;
    const int ks = (int)(r % kslices); r /= kslices;
    const int chunk = (int)r;
    const int go = chunk * nc + ng * 8 + r8;               // GEMM output channel
    int gi, tap;
    if (cin8) {
        gi = ch8;
        tap = 2 * m + kc;                                      // taps paired in (kd,kh,kw) order; the 28th is zero
        if (tap > 26) tap = -1;
    } else {
        gi = ks * 16 + kc * 8 + ch8;
        tap = m;
    }
    float v = 0.f;
    if (go < gout && gi < gin && tap >= 0) {
        if (dgrad) { if (gi < cout_real && go < cin_real) v = w[((long long)gi * cin_real + go) * 27 + (26 - tap)]; }   // w[co=gi][ci=go][flipped tap]
        else if (go < cout_real && gi < cin_real) v = w[((long long)go * cin_real + gi) * 27 + tap];
    }
    return v;
}


// One element of the kd-in-N pack: [ks][m][kc 2][N/8][8 rows][8 ch], N = 4 * gout.
//   row group ng -> (kd' = ng / (gout/8), output channel group); kd' = 3 is the zero block
//   m  -> in-plane tap(s): Cin = 8: taps 2m (kc 0) and 2m+1 (kc 1) of the 9 (kh,kw) taps, the 10th is zero;
//                          else   : tap m = kh*3 + kw, kc selects channels 0-7 / 8-15 of the slice
// dgrad = same contraction with (ci,co) swapped and taps flipped.
// cin_real / cout_real < cin_l / cout_l: the master weight is [cout_real][cin_real][27] and the missing channels are zero
// (2-class head padded to 8 output channels, 2-channel in-block padded to 8 input channels).
__device__ __forceinline__ float pack_kdn_elem(const float* __restrict__ w, long long i, int cin_l, int cout_l, int dgrad,
                                               int cin_real = -1, int cout_real = -1) {
    if (cin_real < 0) cin_real = cin_l;
    if (cout_real < 0) cout_real = cout_l;
    const int gin = dgrad ? cout_l : cin_l, gout = dgrad ? cin_l : cout_l;
    const bool cin8 = gin == 8;
    const int nmp = cin8 ? 5 : 9, ngroups = 4 * gout / 8;
    long long r = i;
    const int ch8 = (int)(r % 8); r /= 8;
    const int r8 = (int)(r % 8); r /= 8;
    const int ng = (int)(r % ngroups); r /= ngroups;
    const int kc = (int)(r % 2); r /= 2;
    const int m = (int)(r % nmp); r /= nmp;
    const int ks = (int)r;
    const int kdp = ng / (gout / 8), go = (ng % (gout / 8)) * 8 + r8;
    int gi, t;
    if (cin8) { gi = ch8; t = 2 * m + kc; if (t > 8) t = -1; }
    else { gi = ks * 16 + kc * 8 + ch8; t = m; }
    if (kdp > 2 || t < 0 || gi >= gin) return 0.f;
    const int tap = kdp * 9 + t;                              // (kd', kh, kw) in the GEMM's (input-side) orientation
    if (dgrad) {                                              // w[co = gi][ci = go][flipped tap]
        if (gi >= cout_real || go >= cin_real) return 0.f;
        return w[((long long)gi * cin_real + go) * 27 + (26 - tap)];
    }
    if (go >= cout_real || gi >= cin_real) return 0.f;
    return w[((long long)go * cin_real + gi) * 27 + tap];
}


// ---- 2x2x2 stride-2 layers (k2s2_tc.cu) ------------------------------------------------------------------------
// master weight wt[A][B][8] (Conv3d: [Cout = A][Cin = B][kd,kh,kw]; ConvTranspose3d: [Cin = A][Cout = B][kd,kh,kw]).
// N chunk / K slice sizes depend on the channel counts only, so one pack serves every volume size.
__host__ __device__ constexpr int k2_gather_nc(int a) { return a <= 16 ? 16 : (a == 32 ? 32 : 64); }
__host__ __device__ constexpr int k2_gather_ks(int b) { return 8 * b < 128 ? 8 * b : 128; }
__host__ __device__ constexpr int k2_scatter_nc(int b) { return b == 8 ? 64 : 128; }
__host__ __device__ constexpr int k2_scatter_ks(int a) { return a < 16 ? 16 : (a < 128 ? a : 128); }
__host__ __device__ inline long long k2_pack_elems(int a, int b, int scatter) {
    if (scatter) {
        const int nc = k2_scatter_nc(b);
        return (long long)((8 * b + nc - 1) / nc) * nc * (a < 16 ? 16 : a);
    }
    const int nc = k2_gather_nc(a);
    return (long long)((a + nc - 1) / nc) * nc * 8 * b;
}
// layout [chunk][K slice][mma][kc 2][NC/8][8 rows][8 k]:
//   gather : row = output channel a, K = (kd,kh,kw,b) = k * B + b
//   scatter: row = accumulator column (kd,kh,kw,b) = k * B + b, K = coarse channel a (zero-padded to 16)
__device__ __forceinline__ float pack_k2s2_elem(const float* __restrict__ w, long long i, int a, int b, int scatter) {
    const int nc = scatter ? k2_scatter_nc(b) : k2_gather_nc(a);
    const int ks = scatter ? k2_scatter_ks(a) : k2_gather_ks(b);
    const int ktot = scatter ? (a < 16 ? 16 : a) : 8 * b;
    const int kstages = ktot / ks, nmma = ks / 16;
    long long r = i;
    const int ch8 = (int)(r % 8); r /= 8;
    const int r8 = (int)(r % 8); r /= 8;
    const int ng = (int)(r % (nc / 8)); r /= (nc / 8);
    const int kc = (int)(r % 2); r /= 2;
    const int m = (int)(r % nmma); r /= nmma;
    const int s = (int)(r % kstages); r /= kstages;
    const int chunk = (int)r;
    const int row = chunk * nc + ng * 8 + r8;
    const int kk = s * ks + m * 16 + kc * 8 + ch8;
    if (scatter) {
        const int k = row / b, bb = row - k * b;
        return (kk < a && k < 8) ? w[((long long)kk * b + bb) * 8 + k] : 0.f;
    }
    const int k = kk / b, bb = kk - k * b;
    return row < a ? w[((long long)row * b + bb) * 8 + k] : 0.f;
}

}  // namespace
