// 2x2x2 stride-2 convolution / transposed convolution as non-overlapping patch contractions.
// Replaces cuDNN for Conv3d(C,C,2,stride=2) (joint_model.py:130) and
// ConvTranspose3d(C,C,2,stride=2) (joint_model.py:118): fprop, dgrad and wgrad of both are the
// three kernels below (see include/vaeseg_b200.h for the gather/scatter algebra).
// These layers are HBM-bound (AI 7-114 flop/B): fp32-accumulate CUDA-core
// kernels with 16-byte channel vectors; weights staged in shared memory.
#include "vs_common.cuh"

namespace {

struct K2Dims { int n, dc, hc, wc, a, b; };

__device__ __forceinline__ void decode_coarse(long long o, const K2Dims& p, int& n, int& od, int& oh, int& ow) {
    ow = (int)(o % p.wc); o /= p.wc;
    oh = (int)(o % p.hc); o /= p.hc;
    od = (int)(o % p.dc); n = (int)(o / p.dc);
}
__device__ __forceinline__ long long fine_index(const K2Dims& p, int n, int od, int oh, int ow, int k) {
    const int fd = 2 * od + (k >> 2), fh = 2 * oh + ((k >> 1) & 1), fw = 2 * ow + (k & 1);
    return (((long long)n * (2 * p.dc) + fd) * (2 * p.hc) + fh) * (2 * p.wc) + fw;
}

// coarse[o,a] = bias[a] + sum_{k,b} wt[a][b][k] * fine[2o+k, b]
template <typename T, int AOB>
__global__ void __launch_bounds__(128) k2s2_gather_kernel(const T* __restrict__ fine, const float* __restrict__ wt,
                                                          const float* __restrict__ bias, T* __restrict__ coarse,
                                                          K2Dims p, long long total) {
    __shared__ float4 ws[8][8][AOB / 4];      // [k][bb][a/4]
    const int t = threadIdx.x;
    const long long o = (long long)blockIdx.x * 128 + t;
    const bool valid = o < total;
    const int a0 = blockIdx.y * AOB;
    int n = 0, od = 0, oh = 0, ow = 0;
    if (valid) decode_coarse(o, p, n, od, oh, ow);
    float acc[AOB];
#pragma unroll
    for (int j = 0; j < AOB; ++j) acc[j] = 0.f;
    for (int b0 = 0; b0 < p.b; b0 += 8) {
        __syncthreads();
        for (int i = t; i < 8 * 8 * AOB; i += 128) {
            int k = i % 8, bb = (i / 8) % 8, aa = i / 64;
            reinterpret_cast<float*>(&ws[k][bb][0])[aa] = wt[((long long)(a0 + aa) * p.b + b0 + bb) * 8 + k];
        }
        __syncthreads();
        if (valid) {
#pragma unroll 2
            for (int k = 0; k < 8; ++k) {
                float xv[8];
                Store<T>::ld8(fine + fine_index(p, n, od, oh, ow, k) * p.b + b0, xv);
#pragma unroll
                for (int bb = 0; bb < 8; ++bb) {
#pragma unroll
                    for (int j = 0; j < AOB / 4; ++j) {
                        const float4 wv = ws[k][bb][j];
                        acc[4 * j + 0] = fmaf(xv[bb], wv.x, acc[4 * j + 0]);
                        acc[4 * j + 1] = fmaf(xv[bb], wv.y, acc[4 * j + 1]);
                        acc[4 * j + 2] = fmaf(xv[bb], wv.z, acc[4 * j + 2]);
                        acc[4 * j + 3] = fmaf(xv[bb], wv.w, acc[4 * j + 3]);
                    }
                }
            }
        }
    }
    if (valid) {
#pragma unroll
        for (int j8 = 0; j8 < AOB / 8; ++j8) {
            float v[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) v[q] = acc[j8 * 8 + q] + (bias ? bias[a0 + j8 * 8 + q] : 0.f);
            Store<T>::st8(coarse + o * p.a + a0 + j8 * 8, v);
        }
    }
}

// fine[2o+k, b] = bias[b] + sum_a wt[a][b][k] * coarse[o, a]
// thread = (two coarse voxels, one k); BOB output channels.
template <typename T, int BOB>
__global__ void __launch_bounds__(256) k2s2_scatter_kernel(const T* __restrict__ coarse, const float* __restrict__ wt,
                                                           const float* __restrict__ bias, T* __restrict__ fine,
                                                           K2Dims p, long long total) {
    __shared__ float4 ws[8][BOB / 4][8];      // [aa][b/4][k]
    const int t = threadIdx.x;
    const int k = t & 7;
    const long long o0 = (long long)blockIdx.x * 64 + (t >> 3);
    const long long o1 = o0 + 32;
    const bool v0 = o0 < total, v1 = o1 < total;
    const int b0 = blockIdx.y * BOB;
    float acc0[BOB], acc1[BOB];
#pragma unroll
    for (int j = 0; j < BOB; ++j) { acc0[j] = 0.f; acc1[j] = 0.f; }
    for (int a0 = 0; a0 < p.a; a0 += 8) {
        __syncthreads();
        for (int i = t; i < 8 * BOB * 8; i += 256) {
            int kk = i % 8, bb = (i / 8) % BOB, aa = i / (8 * BOB);
            reinterpret_cast<float*>(&ws[aa][bb / 4][kk])[bb % 4] = wt[((long long)(a0 + aa) * p.b + b0 + bb) * 8 + kk];
        }
        __syncthreads();
        float x0[8], x1[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) { x0[q] = 0.f; x1[q] = 0.f; }
        if (v0) Store<T>::ld8(coarse + o0 * p.a + a0, x0);
        if (v1) Store<T>::ld8(coarse + o1 * p.a + a0, x1);
#pragma unroll
        for (int aa = 0; aa < 8; ++aa) {
#pragma unroll
            for (int j = 0; j < BOB / 4; ++j) {
                const float4 wv = ws[aa][j][k];
                acc0[4 * j + 0] = fmaf(x0[aa], wv.x, acc0[4 * j + 0]);
                acc0[4 * j + 1] = fmaf(x0[aa], wv.y, acc0[4 * j + 1]);
                acc0[4 * j + 2] = fmaf(x0[aa], wv.z, acc0[4 * j + 2]);
                acc0[4 * j + 3] = fmaf(x0[aa], wv.w, acc0[4 * j + 3]);
                acc1[4 * j + 0] = fmaf(x1[aa], wv.x, acc1[4 * j + 0]);
                acc1[4 * j + 1] = fmaf(x1[aa], wv.y, acc1[4 * j + 1]);
                acc1[4 * j + 2] = fmaf(x1[aa], wv.z, acc1[4 * j + 2]);
                acc1[4 * j + 3] = fmaf(x1[aa], wv.w, acc1[4 * j + 3]);
            }
        }
    }
#pragma unroll
    for (int sel = 0; sel < 2; ++sel) {
        if (!(sel ? v1 : v0)) continue;
        const float* acc = sel ? acc1 : acc0;
        int n, od, oh, ow;
        decode_coarse(sel ? o1 : o0, p, n, od, oh, ow);
        T* pf = fine + fine_index(p, n, od, oh, ow, k) * p.b + b0;
#pragma unroll
        for (int j8 = 0; j8 < BOB / 8; ++j8) {
            float v[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) v[q] = acc[j8 * 8 + q] + (bias ? bias[b0 + j8 * 8 + q] : 0.f);
            Store<T>::st8(pf + j8 * 8, v);
        }
    }
}

// ---- k-split variants: thread = (coarse voxel, filter position k) --------------------------------------------
// The per-voxel kernels above serialise the whole 8*B (or A) reduction in one thread and re-stage the weights every
// 8 channels; on the deep levels (<= 24^3, C >= 32) that is a ~20-70 us latency chain for microseconds of work.  Here
// the CTA stages ALL weights of its 8-channel output chunk once ([k][in][8 out] fp32, +4 words of row padding so the
// eight k-lanes of a voxel read conflict-free) and eight lanes share a voxel.
template <typename T>
__global__ void __launch_bounds__(128) k2s2_gather_ksplit_kernel(const T* __restrict__ fine, const float* __restrict__ wt,
                                                                 const float* __restrict__ bias, T* __restrict__ coarse,
                                                                 K2Dims p, long long total) {
    extern __shared__ float wsm[];                     // [8 k][B][8 a] (+4 pad per k)
    const int t = threadIdx.x, k = t & 7;
    const int a0 = blockIdx.y * 8;
    const int kstride = p.b * 8 + 4;
    for (int i = t; i < 8 * p.b * 8; i += 128) {       // global order [aa][b][k] (contiguous per aa)
        const int kk = i & 7, bb = (i >> 3) % p.b, aa = i / (8 * p.b);
        wsm[kk * kstride + bb * 8 + aa] = wt[((long long)(a0 + aa) * p.b + bb) * 8 + kk];
    }
    __syncthreads();
    const long long o = (long long)blockIdx.x * 16 + (t >> 3);
    const bool valid = o < total;
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    if (valid) {
        int n, od, oh, ow;
        decode_coarse(o, p, n, od, oh, ow);
        const T* px = fine + fine_index(p, n, od, oh, ow, k) * p.b;
        const float* wk = wsm + k * kstride;
        for (int b0 = 0; b0 < p.b; b0 += 8) {
            float xv[8];
            Store<T>::ld8(px + b0, xv);
#pragma unroll
            for (int bb = 0; bb < 8; ++bb) {
                const float4 w0 = *reinterpret_cast<const float4*>(wk + (b0 + bb) * 8);
                const float4 w1 = *reinterpret_cast<const float4*>(wk + (b0 + bb) * 8 + 4);
                acc[0] = fmaf(xv[bb], w0.x, acc[0]); acc[1] = fmaf(xv[bb], w0.y, acc[1]);
                acc[2] = fmaf(xv[bb], w0.z, acc[2]); acc[3] = fmaf(xv[bb], w0.w, acc[3]);
                acc[4] = fmaf(xv[bb], w1.x, acc[4]); acc[5] = fmaf(xv[bb], w1.y, acc[5]);
                acc[6] = fmaf(xv[bb], w1.z, acc[6]); acc[7] = fmaf(xv[bb], w1.w, acc[7]);
            }
        }
    }
#pragma unroll
    for (int q = 0; q < 8; ++q) {                      // sum over the eight k-lanes of the voxel
        acc[q] += __shfl_xor_sync(0xffffffffu, acc[q], 1);
        acc[q] += __shfl_xor_sync(0xffffffffu, acc[q], 2);
        acc[q] += __shfl_xor_sync(0xffffffffu, acc[q], 4);
    }
    if (valid && k == 0) {
        float v[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) v[q] = acc[q] + (bias ? bias[a0 + q] : 0.f);
        Store<T>::st8(coarse + o * p.a + a0, v);
    }
}

template <typename T>
__global__ void __launch_bounds__(128) k2s2_scatter_ksplit_kernel(const T* __restrict__ coarse, const float* __restrict__ wt,
                                                                  const float* __restrict__ bias, T* __restrict__ fine,
                                                                  K2Dims p, long long total) {
    extern __shared__ float wsm[];                     // [8 k][A][8 b] (+4 pad per k)
    const int t = threadIdx.x, k = t & 7;
    const int b0 = blockIdx.y * 8;
    const int kstride = p.a * 8 + 4;
    for (int i = t; i < p.a * 64; i += 128) {          // global order [a][bb][k] (64 contiguous floats per a)
        const int kk = i & 7, bb = (i >> 3) & 7, aa = i >> 6;
        wsm[kk * kstride + aa * 8 + bb] = wt[((long long)aa * p.b + b0 + bb) * 8 + kk];
    }
    __syncthreads();
    const long long o = (long long)blockIdx.x * 16 + (t >> 3);
    if (o >= total) return;
    int n, od, oh, ow;
    decode_coarse(o, p, n, od, oh, ow);
    const T* px = coarse + o * p.a;
    const float* wk = wsm + k * kstride;
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int a0 = 0; a0 < p.a; a0 += 8) {
        float xv[8];
        Store<T>::ld8(px + a0, xv);
#pragma unroll
        for (int aa = 0; aa < 8; ++aa) {
            const float4 w0 = *reinterpret_cast<const float4*>(wk + (a0 + aa) * 8);
            const float4 w1 = *reinterpret_cast<const float4*>(wk + (a0 + aa) * 8 + 4);
            acc[0] = fmaf(xv[aa], w0.x, acc[0]); acc[1] = fmaf(xv[aa], w0.y, acc[1]);
            acc[2] = fmaf(xv[aa], w0.z, acc[2]); acc[3] = fmaf(xv[aa], w0.w, acc[3]);
            acc[4] = fmaf(xv[aa], w1.x, acc[4]); acc[5] = fmaf(xv[aa], w1.y, acc[5]);
            acc[6] = fmaf(xv[aa], w1.z, acc[6]); acc[7] = fmaf(xv[aa], w1.w, acc[7]);
        }
    }
    float v[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) v[q] = acc[q] + (bias ? bias[b0 + q] : 0.f);
    Store<T>::st8(fine + fine_index(p, n, od, oh, ow, k) * p.b + b0, v);
}

// dwt[a][b][k] += sum_o coarse[o,a] * fine[2o+k,b].
// CTA = one 8x8 (a, b) block of the weight and a contiguous range of coarse voxels.  Warp w = filter position k,
// lane = voxel lane: per voxel a thread issues two 16-byte loads (8 coarse channels, 8 channels of fine voxel 2o+k)
// and 64 FMAs into an 8x8 register block -- no shared-memory staging, FMA-bound.  One shuffle reduction over the 32
// voxel lanes and 64 atomics per warp at the end.
template <typename T>
__global__ void __launch_bounds__(256, 2) k2s2_wgrad_kernel(const T* __restrict__ coarse, const T* __restrict__ fine,
                                                         float* __restrict__ dwt, K2Dims p, long long total,
                                                         long long per_cta) {
    const int lane = threadIdx.x & 31, k = threadIdx.x >> 5;
    const int a0 = blockIdx.y * 8, b0 = blockIdx.z * 8;
    float acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int jj = 0; jj < 8; ++jj) acc[i][jj] = 0.f;
    const long long begin = (long long)blockIdx.x * per_cta;
    const long long end = min(total, begin + per_cta);
    // two voxels per thread per iteration: four independent 16-byte loads in flight before the FMAs; two CTAs
    // per SM (<= 128 registers) so that 16 warps hide the DRAM latency
    for (long long o = begin + lane; o < end; o += 64) {
        float c[2][8], f[2][8];
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const long long ou = o + 32 * u;
            if (ou < end) {
                int n, od, oh, ow;
                decode_coarse(ou, p, n, od, oh, ow);
                Store<T>::ld8(coarse + ou * p.a + a0, c[u]);
                Store<T>::ld8(fine + fine_index(p, n, od, oh, ow, k) * p.b + b0, f[u]);
            } else {
#pragma unroll
                for (int q = 0; q < 8; ++q) { c[u][q] = 0.f; f[u][q] = 0.f; }
            }
        }
#pragma unroll
        for (int u = 0; u < 2; ++u)
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int jj = 0; jj < 8; ++jj) acc[i][jj] = fmaf(c[u][i], f[u][jj], acc[i][jj]);
    }
    // Epilogue: the 8 warps (= the 8 filter positions k) hold one 8x8 block each; dwt is [a][b][k] with k innermost,
    // so the CTA's 512 results are 128 aligned groups of four consecutive floats.  Transpose through shared memory
    // and issue 128 float4 atomics instead of 512 scalar ones: with hundreds of CTAs adding onto the same few
    // cache lines the L2 atomic unit, not the FMAs, bounds this kernel.
    __shared__ float red[8][64];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int jj = 0; jj < 8; ++jj) {
            const float v = warp_sum(acc[i][jj]);
            if (lane == ((i * 8 + jj) & 31)) red[k][i * 8 + jj] = v;
        }
    __syncthreads();
    if (threadIdx.x < 128) {
        const int idx = threadIdx.x >> 1, half = threadIdx.x & 1;
        const int i = idx >> 3, jj = idx & 7;
        float* dst = dwt + ((long long)(a0 + i) * p.b + b0 + jj) * 8 + half * 4;
        const float4 v = make_float4(red[half * 4 + 0][idx], red[half * 4 + 1][idx], red[half * 4 + 2][idx], red[half * 4 + 3][idx]);
        if (vs_aligned16_dev(dst)) {
            atomicAdd(reinterpret_cast<float4*>(dst), v);
        } else {
            atomicAdd(dst + 0, v.x); atomicAdd(dst + 1, v.y); atomicAdd(dst + 2, v.z); atomicAdd(dst + 3, v.w);
        }
    }
}

// out[c] += sum_rows t[row][c]   (C % 8 == 0)
template <typename T>
__global__ void __launch_bounds__(256) channel_sum_kernel(const T* __restrict__ x, float* __restrict__ out,
                                                          long long rows, int c) {
    __shared__ double red[256][8];            // fp64 partials: these bias gradients are residuals of cancelling sums
    const int groups = c / 8;                 // <= 32
    const int t = threadIdx.x;
    const int lanes = 256 / groups;
    const int g = t % groups, lane = t / groups;
    double acc[8] = {0., 0., 0., 0., 0., 0., 0., 0.};
    if (lane < lanes) {
        for (long long r = (long long)blockIdx.x * lanes + lane; r < rows; r += (long long)gridDim.x * lanes) {
            float v[8];
            Store<T>::ld8(x + r * c + g * 8, v);
#pragma unroll
            for (int q = 0; q < 8; ++q) acc[q] += (double)v[q];
        }
    }
#pragma unroll
    for (int q = 0; q < 8; ++q) red[t][q] = (lane < lanes) ? acc[q] : 0.;
    __syncthreads();
    if (t < c) {
        const int gg = t / 8, q = t % 8;
        double s = 0.;
        for (int l = 0; l < lanes; ++l) s += red[l * groups + gg][q];
        atomicAdd(out + t, (float)s);
    }
}

template <typename T>
int channel_sum(const T* x, float* out, long long rows, int c, cudaStream_t st) {
    VS_REQUIRE(c % 8 == 0 && c <= 256, VS_ERR_UNSUPPORTED, "channel_sum: C must be a multiple of 8 <= 256 (got %d)", c);
    const int lanes = 256 / (c / 8);
    int blocks = (int)min((long long)vs_sm_count() * 4, (rows + lanes - 1) / lanes);
    channel_sum_kernel<T><<<max(blocks, 1), 256, 0, st>>>(x, out, rows, c);
    VS_CHECK_LAUNCH("channel_sum_kernel");
    return VS_OK;
}

int g_k2_ksplit_max = 8000;      // coarse voxels up to which the k-split kernels are used (tools/kbench.py sweeps it)
bool use_ksplit(long long total, int reduce_channels) {
    return total <= g_k2_ksplit_max && reduce_channels <= 256;
}

int check_k2(const void* p0, const void* p1, const void* p2, int n, int dc, int hc, int wc, int a, int b, const char* who) {
    VS_REQUIRE(n > 0 && dc > 0 && hc > 0 && wc > 0, VS_ERR_SHAPE, "%s: bad shape", who);
    VS_REQUIRE(a % 8 == 0 && b % 8 == 0 && a >= 8 && b >= 8, VS_ERR_UNSUPPORTED, "%s: channels must be multiples of 8 (A=%d B=%d)", who, a, b);
    VS_REQUIRE(p0 && p1 && p2, VS_ERR_SHAPE, "%s: null pointer", who);
    VS_REQUIRE(vs_aligned16(p0) && vs_aligned16(p1) && vs_aligned16(p2), VS_ERR_ALIGN, "%s: pointers must be 16B aligned", who);
    return VS_OK;
}

}  // namespace

extern "C" void vs_debug_set_k2_ksplit_max(int v) { g_k2_ksplit_max = v; }
#ifdef VS_WITH_TCGEN05
extern "C" int vs_k2s2_wgrad_tc(const void* coarse, const void* fine, float* dwt, int accumulate, int n, int dc, int hc,
                                int wc, int a, int b, void* stream);
#endif
static int g_k2_wgrad_tc = 1;      // 0: CUDA-core kernel; 1: tensor cores for the large layers; 2: tensor cores always (tests)
extern "C" void vs_debug_set_k2_wgrad_tc(int on) { g_k2_wgrad_tc = on; }

extern "C" int vs_k2s2_gather(int dtype, const void* fine, const float* wt, const float* bias, void* coarse,
                              int n, int dc, int hc, int wc, int a, int b, void* stream) {
    int rc = check_k2(fine, wt, coarse, n, dc, hc, wc, a, b, "k2s2_gather");
    if (rc) return rc;
    K2Dims p = {n, dc, hc, wc, a, b};
    const long long total = (long long)n * dc * hc * wc;
    cudaStream_t st = (cudaStream_t)stream;
    VS_DISPATCH_DTYPE(dtype, T, {
        if (use_ksplit(total, b)) {
            const size_t sm = (size_t)(8 * (b * 8 + 4)) * sizeof(float);
            auto kern = k2s2_gather_ksplit_kernel<T>;
            if (sm > 48 * 1024) VS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm), "k2s2 smem attribute");
            dim3 grid(vs_ceil_div(total, 16), a / 8);
            kern<<<grid, 128, sm, st>>>((const T*)fine, wt, bias, (T*)coarse, p, total);
        } else if (a % 16 == 0) {
            dim3 grid(vs_ceil_div(total, 128), a / 16);
            k2s2_gather_kernel<T, 16><<<grid, 128, 0, st>>>((const T*)fine, wt, bias, (T*)coarse, p, total);
        } else {
            dim3 grid(vs_ceil_div(total, 128), a / 8);
            k2s2_gather_kernel<T, 8><<<grid, 128, 0, st>>>((const T*)fine, wt, bias, (T*)coarse, p, total);
        }
    });
    VS_CHECK_LAUNCH("k2s2_gather_kernel");
    return VS_OK;
}

extern "C" int vs_k2s2_scatter(int dtype, const void* coarse, const float* wt, const float* bias, void* fine,
                               int n, int dc, int hc, int wc, int a, int b, void* stream) {
    int rc = check_k2(coarse, wt, fine, n, dc, hc, wc, a, b, "k2s2_scatter");
    if (rc) return rc;
    K2Dims p = {n, dc, hc, wc, a, b};
    const long long total = (long long)n * dc * hc * wc;
    cudaStream_t st = (cudaStream_t)stream;
    VS_DISPATCH_DTYPE(dtype, T, {
        if (use_ksplit(total, a)) {
            const size_t sm = (size_t)(8 * (a * 8 + 4)) * sizeof(float);
            auto kern = k2s2_scatter_ksplit_kernel<T>;
            if (sm > 48 * 1024) VS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm), "k2s2 smem attribute");
            dim3 grid(vs_ceil_div(total, 16), b / 8);
            kern<<<grid, 128, sm, st>>>((const T*)coarse, wt, bias, (T*)fine, p, total);
        } else if (b % 16 == 0) {
            dim3 grid(vs_ceil_div(total, 64), b / 16);
            k2s2_scatter_kernel<T, 16><<<grid, 256, 0, st>>>((const T*)coarse, wt, bias, (T*)fine, p, total);
        } else {
            dim3 grid(vs_ceil_div(total, 64), b / 8);
            k2s2_scatter_kernel<T, 8><<<grid, 256, 0, st>>>((const T*)coarse, wt, bias, (T*)fine, p, total);
        }
    });
    VS_CHECK_LAUNCH("k2s2_scatter_kernel");
    return VS_OK;
}

extern "C" int vs_k2s2_wgrad(int dtype, const void* coarse, const void* fine, float* dwt, float* dbias_coarse,
                             float* dbias_fine, int accumulate, int n, int dc, int hc, int wc, int a, int b,
                             void* stream) {
    int rc = check_k2(coarse, fine, dwt, n, dc, hc, wc, a, b, "k2s2_wgrad");
    if (rc) return rc;
    K2Dims p = {n, dc, hc, wc, a, b};
    const long long total = (long long)n * dc * hc * wc;
    cudaStream_t st = (cudaStream_t)stream;
    bool tc = false;
#ifdef VS_WITH_TCGEN05
    // tensor cores from 24^3 coarse voxels per pair of samples up (measured, tools/k2bench.py: 108 -> 33 us at 2 x 48^3, 16
    // channels; 55 -> 41 us at 24^3, 32 channels); below that the launch is a handful of tiles and the per-CTA atomic
    // read-back of the TMEM accumulators costs more than the CUDA-core kernel's whole run (16 vs 40 us at 12^3)
    tc = g_k2_wgrad_tc && dtype == VS_BF16 && a <= 256 && b <= 256 && (a == 8 || a % 16 == 0) && (b == 8 || b % 16 == 0) &&
         (total >= 20000 || g_k2_wgrad_tc == 2);
#endif
    if (!accumulate) {
        if (!tc) VS_CUDA(cudaMemsetAsync(dwt, 0, sizeof(float) * 8 * a * b, st), "k2s2 wgrad memset");
        if (dbias_coarse) VS_CUDA(cudaMemsetAsync(dbias_coarse, 0, sizeof(float) * a, st), "k2s2 wgrad memset");
        if (dbias_fine) VS_CUDA(cudaMemsetAsync(dbias_fine, 0, sizeof(float) * b, st), "k2s2 wgrad memset");
    }
#ifdef VS_WITH_TCGEN05
    if (tc) {
        rc = vs_k2s2_wgrad_tc(coarse, fine, dwt, accumulate, n, dc, hc, wc, a, b, stream);
        if (rc) return rc;
        if (dbias_coarse) { rc = channel_sum<bf16>((const bf16*)coarse, dbias_coarse, total, a, st); if (rc) return rc; }
        if (dbias_fine) { rc = channel_sum<bf16>((const bf16*)fine, dbias_fine, total * 8, b, st); if (rc) return rc; }
        return VS_OK;
    }
#endif
    VS_DISPATCH_DTYPE(dtype, T, {
        dim3 grid(1, a / 8, b / 8);
        const long long blocks = (long long)grid.y * grid.z;
        long long parts = max(1LL, (long long)vs_sm_count() * 4 / blocks);        // ~2 waves of 2 CTAs per SM
        parts = min(parts, (total + 63) / 64);                                      // >= 64 voxels per CTA
        const long long per_cta = (total + parts - 1) / parts;
        grid.x = (unsigned)((total + per_cta - 1) / per_cta);
        k2s2_wgrad_kernel<T><<<grid, 256, 0, st>>>((const T*)coarse, (const T*)fine, dwt, p, total, per_cta);
        VS_CHECK_LAUNCH("k2s2_wgrad_kernel");
        if (dbias_coarse) { rc = channel_sum<T>((const T*)coarse, dbias_coarse, total, a, st); if (rc) return rc; }
        if (dbias_fine) { rc = channel_sum<T>((const T*)fine, dbias_fine, total * 8, b, st); if (rc) return rc; }
    });
    return VS_OK;
}
