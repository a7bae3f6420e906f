// Per-(n,c) affine + ReLU passes: a = relu(k * y + b) (+ skip) and its backward, for normalisation layers whose
// statistics are NOT per (n,c) and therefore cannot use the InstanceNorm kernels of norm_act.cu -- BatchNorm3d with
// affine parameters and running statistics (joint_model.py:12-13, `Normalization(norm_type=2)`).  The host side
// (engine.py) pools the per-(n,c) convolution statistics over the batch, folds gamma / beta / the stored shift into the
// coefficient tables, and these kernels do the three tensor passes:
//   apply       a  = relu(k y + b) + skip                                   kb   [N][C][2] fp32
//   bwd_reduce  sums[n][c] = (sum gm, sum gm * xhat), gm = g * [k y + b > 0], xhat = k2 y + b2   (fp64 accumulation)
//   bwd_apply   dy = c0 gm + c1 + c2 y                                      coef [N][C][3] fp32
// NDHWC storage, C a power of two in [8, 256], 16-byte channel vectors; HBM-bound like their InstanceNorm twins.
#include "vs_common.cuh"

namespace {

constexpr int NT = 256;

template <typename T>
__global__ void __launch_bounds__(NT) affine_relu_apply_kernel(const T* __restrict__ y, const float* __restrict__ kb,
                                                               const T* __restrict__ skip, T* __restrict__ a, long long s, int c) {
    __shared__ float sk[256], sb[256];
    const int n = blockIdx.y, t = threadIdx.x;
    for (int ch = t; ch < c; ch += NT) { sk[ch] = kb[((long long)n * c + ch) * 2]; sb[ch] = kb[((long long)n * c + ch) * 2 + 1]; }
    __syncthreads();
    const int groups = c / 8;
    const long long nvec = s * groups;
    const long long base = (long long)n * s * c;
    for (long long i = (long long)blockIdx.x * NT + t; i < nvec; i += (long long)gridDim.x * NT) {
        const int cg = (int)(i % groups) * 8;
        float v[8];
        Store<T>::ld8(y + base + i * 8, v);
#pragma unroll
        for (int q = 0; q < 8; ++q) v[q] = fmaxf(fmaf(sk[cg + q], v[q], sb[cg + q]), 0.f);
        if (skip != nullptr) {
            float k[8];
            Store<T>::ld8(skip + base + i * 8, k);
#pragma unroll
            for (int q = 0; q < 8; ++q) v[q] += k[q];
        }
        Store<T>::st8(a + base + i * 8, v);
    }
}

template <typename T>
__global__ void __launch_bounds__(NT) affine_relu_bwd_reduce_kernel(const T* __restrict__ g, const T* __restrict__ y,
                                                                    const float* __restrict__ kb, const float* __restrict__ k2b2,
                                                                    double* __restrict__ sums, long long s, int c) {
    __shared__ float sk[256], sb[256], sk2[256], sb2[256];
    __shared__ double red[NT][17];
    const int n = blockIdx.y, t = threadIdx.x;
    for (int ch = t; ch < c; ch += NT) {
        sk[ch] = kb[((long long)n * c + ch) * 2]; sb[ch] = kb[((long long)n * c + ch) * 2 + 1];
        sk2[ch] = k2b2[((long long)n * c + ch) * 2]; sb2[ch] = k2b2[((long long)n * c + ch) * 2 + 1];
    }
    __syncthreads();
    const int groups = c / 8, lanes = NT / groups;
    const int cg = (t % groups) * 8, lane = t / groups;
    const long long base = (long long)n * s * c;
    double a0[8], a1[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) { a0[q] = 0.0; a1[q] = 0.0; }
    for (long long v = (long long)blockIdx.x * lanes + lane; v < s; v += (long long)gridDim.x * lanes) {
        float gv[8], yv[8];
        Store<T>::ld8(g + base + v * c + cg, gv);
        Store<T>::ld8(y + base + v * c + cg, yv);
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            const float pre = fmaf(sk[cg + q], yv[q], sb[cg + q]);
            const float xh = fmaf(sk2[cg + q], yv[q], sb2[cg + q]);
            const float gm = pre > 0.f ? gv[q] : 0.f;
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
__global__ void __launch_bounds__(NT) affine_relu_bwd_apply_kernel(const T* __restrict__ g, const T* __restrict__ y,
                                                                   const float* __restrict__ kb, const float* __restrict__ coef,
                                                                   T* __restrict__ dy, long long s, int c) {
    __shared__ float sk[256], sb[256], c0[256], c1[256], c2[256];
    const int n = blockIdx.y, t = threadIdx.x;
    for (int ch = t; ch < c; ch += NT) {
        sk[ch] = kb[((long long)n * c + ch) * 2]; sb[ch] = kb[((long long)n * c + ch) * 2 + 1];
        c0[ch] = coef[((long long)n * c + ch) * 3]; c1[ch] = coef[((long long)n * c + ch) * 3 + 1]; c2[ch] = coef[((long long)n * c + ch) * 3 + 2];
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
            const float pre = fmaf(sk[cg + q], yv[q], sb[cg + q]);
            const float gm = pre > 0.f ? gv[q] : 0.f;
            o[q] = fmaf(c0[cg + q], gm, fmaf(c2[cg + q], yv[q], c1[cg + q]));
        }
        Store<T>::st8(dy + base + i * 8, o);
    }
}

int affine_grid(long long work_items) {
    long long blocks = (work_items + NT - 1) / NT;
    long long cap = (long long)vs_sm_count() * 8;
    return (int)max(1LL, min(blocks, cap));
}

int check_affine(const void* a, const void* b, const void* tab, int n, long long s, int c, const char* who) {
    VS_REQUIRE(n > 0 && s > 0 && c >= 8 && c <= 256 && (c & (c - 1)) == 0, VS_ERR_UNSUPPORTED, "%s: need C a power of two in [8,256] (C=%d)", who, c);
    VS_REQUIRE(a && b && tab, VS_ERR_SHAPE, "%s: null pointer", who);
    VS_REQUIRE(vs_aligned16(a) && vs_aligned16(b), VS_ERR_ALIGN, "%s: pointers must be 16B aligned", who);
    return VS_OK;
}

}  // namespace

extern "C" int vs_affine_relu_apply(int dtype, const void* y, const float* kb, const void* skip, void* a, int n, long long s,
                                    int c, void* stream) {
    int rc = check_affine(y, a, kb, n, s, c, "affine_relu_apply");
    if (rc) return rc;
    dim3 grid(affine_grid(s * (c / 8) / 4 + 1), n);
    VS_DISPATCH_DTYPE(dtype, T, { affine_relu_apply_kernel<T><<<grid, NT, 0, (cudaStream_t)stream>>>(
        (const T*)y, kb, (const T*)skip, (T*)a, s, c); });
    VS_CHECK_LAUNCH("affine_relu_apply_kernel");
    return VS_OK;
}

extern "C" int vs_affine_relu_bwd_reduce(int dtype, const void* g, const void* y, const float* kb, const float* k2b2,
                                         double* sums, int n, long long s, int c, void* stream) {
    int rc = check_affine(g, y, kb, n, s, c, "affine_relu_bwd_reduce");
    if (rc) return rc;
    VS_REQUIRE(k2b2 && sums, VS_ERR_SHAPE, "affine_relu_bwd_reduce: null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    VS_CUDA(cudaMemsetAsync(sums, 0, sizeof(double) * 2 * n * c, st), "affine bwd memset");
    const int lanes = NT / (c / 8);
    long long blocks = (s + (long long)lanes * 8 - 1) / ((long long)lanes * 8);
    dim3 grid((unsigned)max(1LL, min(blocks, (long long)vs_sm_count() * 4)), n);
    VS_DISPATCH_DTYPE(dtype, T, { affine_relu_bwd_reduce_kernel<T><<<grid, NT, 0, st>>>((const T*)g, (const T*)y, kb, k2b2, sums, s, c); });
    VS_CHECK_LAUNCH("affine_relu_bwd_reduce_kernel");
    return VS_OK;
}

extern "C" int vs_affine_relu_bwd_apply(int dtype, const void* g, const void* y, const float* kb, const float* coef, void* dy,
                                        int n, long long s, int c, void* stream) {
    int rc = check_affine(g, dy, kb, n, s, c, "affine_relu_bwd_apply");
    if (rc) return rc;
    VS_REQUIRE(y && coef && vs_aligned16(y), VS_ERR_SHAPE, "affine_relu_bwd_apply: null or misaligned pointer");
    dim3 grid(affine_grid(s * (c / 8) / 4 + 1), n);
    VS_DISPATCH_DTYPE(dtype, T, { affine_relu_bwd_apply_kernel<T><<<grid, NT, 0, (cudaStream_t)stream>>>(
        (const T*)g, (const T*)y, kb, coef, (T*)dy, s, c); });
    VS_CHECK_LAUNCH("affine_relu_bwd_apply_kernel");
    return VS_OK;
}
