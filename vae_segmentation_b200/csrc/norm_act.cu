// InstanceNorm3d(affine=False, eps=1e-5, biased variance) + ReLU, forward and backward, the
// additive-skip gradient fan-in, and the 2-class softmax.  All HBM-bound streaming kernels:
// 16-byte (bf16) / 32-byte (fp32) channel vectors, per-(n,c) constants staged in shared memory,
// warp-shuffle + shared-memory reductions, one atomic per CTA per statistic.
// Replaces ATen instance_norm / relu / softmax reached from joint_model.py:9-15,38,104,225,367.
#include "vs_common.cuh"

namespace {

constexpr int NT = 256;

template <typename T>
__global__ void __launch_bounds__(NT) inorm_relu_apply_kernel(const T* __restrict__ y, const double* __restrict__ stats,
                                                              const T* __restrict__ skip, T* __restrict__ a,
                                                              long long s, int c) {
    __shared__ float sm[256], sr[256];
    pdl_trigger();
    pdl_wait();
    const int n = blockIdx.y, t = threadIdx.x;
    const double inv_s = 1.0 / (double)s;
    for (int ch = t; ch < c; ch += NT) in_mean_rstd(stats + ((long long)n * c + ch) * 2, inv_s, sm[ch], sr[ch]);
    __syncthreads();
    const int groups = c / 8;
    const long long nvec = s * groups;
    const long long base = (long long)n * s * c;
    for (long long i = (long long)blockIdx.x * NT + t; i < nvec; i += (long long)gridDim.x * NT) {
        const int cg = (int)(i % groups) * 8;
        float v[8];
        Store<T>::ld8(y + base + i * 8, v);
#pragma unroll
        for (int q = 0; q < 8; ++q) v[q] = fmaxf((v[q] - sm[cg + q]) * sr[cg + q], 0.f);
        if (skip != nullptr) {
            float k[8];
            Store<T>::ld8(skip + base + i * 8, k);
#pragma unroll
            for (int q = 0; q < 8; ++q) v[q] += k[q];
        }
        Store<T>::st8(a + base + i * 8, v);
    }
}

// sums[n][c] = (sum g*mask, sum g*mask*xhat)
template <typename T>
__global__ void __launch_bounds__(NT) inorm_relu_bwd_reduce_kernel(const T* __restrict__ g, const T* __restrict__ y,
                                                                   const double* __restrict__ stats,
                                                                   double* __restrict__ sums, long long s, int c) {
    __shared__ float sm[256], sr[256];
    __shared__ double red[NT][17];
    pdl_trigger();
    pdl_wait();
    const int n = blockIdx.y, t = threadIdx.x;
    const double inv_s = 1.0 / (double)s;
    for (int ch = t; ch < c; ch += NT) in_mean_rstd(stats + ((long long)n * c + ch) * 2, inv_s, sm[ch], sr[ch]);
    __syncthreads();
    const int groups = c / 8, lanes = NT / groups;
    const int cg = (t % groups) * 8, lane = t / groups;
    const long long base = (long long)n * s * c;
    // fp64 accumulation: dy = rstd*(g - mean(g) - xhat*mean(g*xhat)) cancels heavily when g is
    // nearly constant, and ATen's CPU path (the oracle) also accumulates these sums in double
    double a0[8], a1[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) { a0[q] = 0.0; a1[q] = 0.0; }
    for (long long v = (long long)blockIdx.x * lanes + lane; v < s; v += (long long)gridDim.x * lanes) {
        float gv[8], yv[8];
        Store<T>::ld8(g + base + v * c + cg, gv);
        Store<T>::ld8(y + base + v * c + cg, yv);
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            const float xh = (yv[q] - sm[cg + q]) * sr[cg + q];
            const float gm = xh > 0.f ? gv[q] : 0.f;
            a0[q] += (double)gm;
            a1[q] += (double)gm * (double)xh;
        }
    }
#pragma unroll
    for (int q = 0; q < 8; ++q) { red[t][q] = a0[q]; red[t][8 + q] = a1[q]; }
    __syncthreads();
    for (int o = t; o < c * 2; o += NT) {
        const int ch = o >> 1, which = o & 1;
        const int gg = ch / 8, q = ch % 8;
        double sacc = 0.0;
        for (int l = 0; l < lanes; ++l) sacc += red[l * groups + gg][which * 8 + q];
        atomicAdd(sums + ((long long)n * c + ch) * 2 + which, sacc);
    }
}

template <typename T>
__global__ void __launch_bounds__(NT) inorm_relu_bwd_apply_kernel(const T* __restrict__ g, const T* __restrict__ y,
                                                                  const double* __restrict__ stats,
                                                                  const double* __restrict__ sums, T* __restrict__ dy,
                                                                  long long s, int c) {
    __shared__ float sm[256], sr[256], m0[256], m1[256];
    pdl_trigger();
    pdl_wait();
    const int n = blockIdx.y, t = threadIdx.x;
    const double inv_s = 1.0 / (double)s;
    for (int ch = t; ch < c; ch += NT) {
        in_mean_rstd(stats + ((long long)n * c + ch) * 2, inv_s, sm[ch], sr[ch]);
        m0[ch] = (float)(sums[((long long)n * c + ch) * 2] * inv_s);
        m1[ch] = (float)(sums[((long long)n * c + ch) * 2 + 1] * inv_s);
    }
    __syncthreads();
    const int groups = c / 8;
    const long long nvec = s * groups;
    const long long base = (long long)n * s * c;
    for (long long i = (long long)blockIdx.x * NT + t; i < nvec; i += (long long)gridDim.x * NT) {
        const int cg = (int)(i % groups) * 8;
        float gv[8], yv[8], o[8];
        Store<T>::ld8(g + base + i * 8, gv);
        Store<T>::ld8(y + base + i * 8, yv);
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            const float xh = (yv[q] - sm[cg + q]) * sr[cg + q];
            const float gm = xh > 0.f ? gv[q] : 0.f;
            o[q] = sr[cg + q] * (gm - m0[cg + q] - xh * m1[cg + q]);
        }
        Store<T>::st8(dy + base + i * 8, o);
    }
}

template <typename T>
__global__ void __launch_bounds__(NT) add_inplace_kernel(T* __restrict__ dst, const T* __restrict__ src, long long nvec) {
    pdl_trigger();
    pdl_wait();
    for (long long i = (long long)blockIdx.x * NT + threadIdx.x; i < nvec; i += (long long)gridDim.x * NT) {
        float a[8], b[8];
        Store<T>::ld8(dst + i * 8, a);
        Store<T>::ld8(src + i * 8, b);
#pragma unroll
        for (int q = 0; q < 8; ++q) a[q] += b[q];
        Store<T>::st8(dst + i * 8, a);
    }
}

__global__ void __launch_bounds__(NT) softmax2_fwd_kernel(const float* __restrict__ logits, float* __restrict__ probs,
                                                          long long s) {
    const int n = blockIdx.y;
    const float2* lg = reinterpret_cast<const float2*>(logits) + (long long)n * s;
    float* p0 = probs + (long long)n * 2 * s;
    float* p1 = p0 + s;
    for (long long v = (long long)blockIdx.x * NT + threadIdx.x; v < s; v += (long long)gridDim.x * NT) {
        const float2 l = lg[v];
        const float m = fmaxf(l.x, l.y);
        const float e0 = expf(l.x - m), e1 = expf(l.y - m);
        const float inv = 1.f / (e0 + e1);
        p0[v] = e0 * inv;
        p1[v] = e1 * inv;
    }
}

template <typename T>
__global__ void __launch_bounds__(NT) softmax2_bwd_kernel(const float* __restrict__ dprobs, const float* __restrict__ probs,
                                                          T* __restrict__ dlogits, long long s) {
    const int n = blockIdx.y;
    const float* g0 = dprobs + (long long)n * 2 * s; const float* g1 = g0 + s;
    const float* p0 = probs + (long long)n * 2 * s; const float* p1 = p0 + s;
    T* dl = dlogits + (long long)n * 2 * s;
    for (long long v = (long long)blockIdx.x * NT + threadIdx.x; v < s; v += (long long)gridDim.x * NT) {
        const float a = p0[v], b = p1[v], ga = g0[v], gb = g1[v];
        const float dot = ga * a + gb * b;
        Store<T>::st2(dl + v * 2, a * (ga - dot), b * (gb - dot));
    }
}

// 2-class softmax backward writing the logit gradient as an 8-channel bf16 NDHWC tensor (channels 2..7 zero) so that
// the head's wgrad / dgrad run on the tensor-core kernels (which take channel counts in multiples of 8), plus the
// head's bias gradient db[c] += sum_v dlogit[v][c] (fp64 block partials, one atomic pair per CTA).
__global__ void __launch_bounds__(NT) softmax2_bwd_pad8_kernel(const float* __restrict__ dprobs, const float* __restrict__ probs,
                                                               bf16* __restrict__ dlogits8, float* __restrict__ db, long long s) {
    __shared__ double red[NT / 32][2];
    const int n = blockIdx.y;
    const float* g0 = dprobs + (long long)n * 2 * s; const float* g1 = g0 + s;
    const float* p0 = probs + (long long)n * 2 * s; const float* p1 = p0 + s;
    uint4* dl = reinterpret_cast<uint4*>(dlogits8) + (long long)n * s;
    double s0 = 0.0, s1 = 0.0;
    for (long long v = (long long)blockIdx.x * NT + threadIdx.x; v < s; v += (long long)gridDim.x * NT) {
        const float a = p0[v], b = p1[v], ga = g0[v], gb = g1[v];
        const float dot = ga * a + gb * b;
        const __nv_bfloat162 h = __floats2bfloat162_rn(a * (ga - dot), b * (gb - dot));
        dl[v] = make_uint4(*reinterpret_cast<const uint32_t*>(&h), 0u, 0u, 0u);
        const float2 r = __bfloat1622float2(h);           // the bias gradient sums exactly what wgrad / dgrad consume
        s0 += (double)r.x; s1 += (double)r.y;
    }
    if (db != nullptr) {
        s0 = warp_sum(s0); s1 = warp_sum(s1);
        if ((threadIdx.x & 31) == 0) { red[threadIdx.x >> 5][0] = s0; red[threadIdx.x >> 5][1] = s1; }
        __syncthreads();
        if (threadIdx.x < 2) {
            double t = 0.0;
            for (int k = 0; k < NT / 32; ++k) t += red[k][threadIdx.x];
            atomicAdd(db + threadIdx.x, (float)t);
        }
    }
}

// planar fp32 [N][C][S] (C < 8) -> NDHWC bf16 [N][S][8] with zero channels C..7: the in-block input as a
// tensor-core wgrad operand.
__global__ void __launch_bounds__(NT) planar_to_ndhwc8_kernel(const float* __restrict__ x, bf16* __restrict__ out, int c, long long s) {
    const int n = blockIdx.y;
    const float* xp = x + (long long)n * c * s;
    bf16* op = out + (long long)n * s * 8;
    for (long long v = (long long)blockIdx.x * NT + threadIdx.x; v < s; v += (long long)gridDim.x * NT) {
        float f[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        for (int k = 0; k < c; ++k) f[k] = xp[(long long)k * s + v];
        Store<bf16>::st8(op + v * 8, f);
    }
}

int g_norm_vpt = 4;      // 16-byte vectors per thread that size the grids of the apply passes (vs_debug_set_norm_vpt)
int stream_grid(long long work_items) {
    long long blocks = (work_items + NT - 1) / NT;
    long long cap = (long long)vs_sm_count() * 8;
    return (int)max(1LL, min(blocks, cap));
}

int check_norm(const void* a, const void* b, int n, long long s, int c, const char* who) {
    VS_REQUIRE(n > 0 && s > 0 && c >= 8 && c <= 256 && (c & (c - 1)) == 0, VS_ERR_UNSUPPORTED, "%s: need C a power of two in [8,256] (C=%d)", who, c);
    VS_REQUIRE(a && b, VS_ERR_SHAPE, "%s: null pointer", who);
    VS_REQUIRE(vs_aligned16(a) && vs_aligned16(b), VS_ERR_ALIGN, "%s: pointers must be 16B aligned", who);
    return VS_OK;
}

}  // namespace

extern "C" int vs_inorm_relu_apply(int dtype, const void* y, const double* stats, const void* skip, void* a,
                                   int n, long long s, int c, void* stream) {
    int rc = check_norm(y, a, n, s, c, "inorm_relu_apply");
    if (rc) return rc;
    dim3 grid(stream_grid(s * (c / 8) / g_norm_vpt + 1), n);
    VS_DISPATCH_DTYPE(dtype, T, { VS_CUDA(vs_launch(inorm_relu_apply_kernel<T>, grid, dim3(NT), 0, (cudaStream_t)stream,
        (const T*)y, stats, (const T*)skip, (T*)a, s, c), "inorm_relu_apply launch"); });
    VS_CHECK_LAUNCH("inorm_relu_apply_kernel");
    return VS_OK;
}

extern "C" int vs_inorm_relu_bwd_reduce(int dtype, const void* g, const void* y, const double* stats, double* sums,
                                        int n, long long s, int c, int flags, void* stream) {
    int rc = check_norm(g, y, n, s, c, "inorm_relu_bwd_reduce");
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    if (!(flags & VS_FLAG_PREZEROED)) VS_CUDA(cudaMemsetAsync(sums, 0, sizeof(double) * 2 * n * c, st), "inorm bwd memset");
    const int lanes = NT / (c / 8);
    long long blocks = (s + (long long)lanes * 8 - 1) / ((long long)lanes * 8);
    dim3 grid((unsigned)max(1LL, min(blocks, (long long)vs_sm_count() * 4)), n);
    VS_DISPATCH_DTYPE(dtype, T, { VS_CUDA(vs_launch(inorm_relu_bwd_reduce_kernel<T>, grid, dim3(NT), 0, st,
        (const T*)g, (const T*)y, stats, sums, s, c), "inorm_relu_bwd_reduce launch"); });
    VS_CHECK_LAUNCH("inorm_relu_bwd_reduce_kernel");
    return VS_OK;
}

extern "C" int vs_inorm_relu_bwd_apply(int dtype, const void* g, const void* y, const double* stats, const double* sums,
                                       void* dy, int n, long long s, int c, void* stream) {
    int rc = check_norm(g, dy, n, s, c, "inorm_relu_bwd_apply");
    if (rc) return rc;
    dim3 grid(stream_grid(s * (c / 8) / g_norm_vpt + 1), n);
    VS_DISPATCH_DTYPE(dtype, T, { VS_CUDA(vs_launch(inorm_relu_bwd_apply_kernel<T>, grid, dim3(NT), 0, (cudaStream_t)stream,
        (const T*)g, (const T*)y, stats, sums, (T*)dy, s, c), "inorm_relu_bwd_apply launch"); });
    VS_CHECK_LAUNCH("inorm_relu_bwd_apply_kernel");
    return VS_OK;
}

extern "C" int vs_add_inplace(int dtype, void* dst, const void* src, long long count, void* stream) {
    VS_REQUIRE(dst && src && count > 0 && count % 8 == 0, VS_ERR_SHAPE, "add_inplace: count must be a positive multiple of 8");
    VS_REQUIRE(vs_aligned16(dst) && vs_aligned16(src), VS_ERR_ALIGN, "add_inplace: pointers must be 16B aligned");
    const long long nvec = count / 8;
    VS_DISPATCH_DTYPE(dtype, T, { VS_CUDA(vs_launch(add_inplace_kernel<T>, dim3(stream_grid(nvec / g_norm_vpt + 1)), dim3(NT), 0,
        (cudaStream_t)stream, (T*)dst, (const T*)src, nvec), "add_inplace launch"); });
    VS_CHECK_LAUNCH("add_inplace_kernel");
    return VS_OK;
}

extern "C" int vs_softmax2_fwd(const float* logits, float* probs, int n, long long s, void* stream) {
    VS_REQUIRE(logits && probs && n > 0 && s > 0, VS_ERR_SHAPE, "softmax2_fwd: bad arguments");
    dim3 grid(stream_grid(s / 4 + 1), n);
    softmax2_fwd_kernel<<<grid, NT, 0, (cudaStream_t)stream>>>(logits, probs, s);
    VS_CHECK_LAUNCH("softmax2_fwd_kernel");
    return VS_OK;
}

extern "C" int vs_softmax2_bwd(int dtype, const float* dprobs, const float* probs, void* dlogits, int n, long long s,
                               void* stream) {
    VS_REQUIRE(dprobs && probs && dlogits && n > 0 && s > 0, VS_ERR_SHAPE, "softmax2_bwd: bad arguments");
    dim3 grid(stream_grid(s / 4 + 1), n);
    VS_DISPATCH_DTYPE(dtype, T, { softmax2_bwd_kernel<T><<<grid, NT, 0, (cudaStream_t)stream>>>(dprobs, probs, (T*)dlogits, s); });
    VS_CHECK_LAUNCH("softmax2_bwd_kernel");
    return VS_OK;
}

extern "C" int vs_softmax2_bwd_pad8(const float* dprobs, const float* probs, void* dlogits8, float* db, int n, long long s,
                                    void* stream) {
    VS_REQUIRE(dprobs && probs && dlogits8 && n > 0 && s > 0, VS_ERR_SHAPE, "softmax2_bwd_pad8: bad arguments");
    VS_REQUIRE(vs_aligned16(dlogits8), VS_ERR_ALIGN, "softmax2_bwd_pad8: output must be 16B aligned");
    dim3 grid(stream_grid(s / 4 + 1), n);
    softmax2_bwd_pad8_kernel<<<grid, NT, 0, (cudaStream_t)stream>>>(dprobs, probs, (bf16*)dlogits8, db, s);
    VS_CHECK_LAUNCH("softmax2_bwd_pad8_kernel");
    return VS_OK;
}

extern "C" int vs_planar_to_ndhwc8(const float* x, void* out, int n, int c, long long s, void* stream) {
    VS_REQUIRE(x && out && n > 0 && s > 0 && c > 0 && c <= 8, VS_ERR_SHAPE, "planar_to_ndhwc8: bad arguments (C=%d)", c);
    VS_REQUIRE(vs_aligned16(out), VS_ERR_ALIGN, "planar_to_ndhwc8: output must be 16B aligned");
    dim3 grid(stream_grid(s / 2 + 1), n);
    planar_to_ndhwc8_kernel<<<grid, NT, 0, (cudaStream_t)stream>>>(x, (bf16*)out, c, s);
    VS_CHECK_LAUNCH("planar_to_ndhwc8_kernel");
    return VS_OK;
}

extern "C" void vs_debug_set_norm_vpt(int vpt) { g_norm_vpt = vpt > 0 ? vpt : 4; }
