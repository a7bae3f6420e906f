// Multi-tensor optimiser / EMA-teacher updates over flat fp32 arenas: one launch per step
// instead of torch.optim's per-tensor foreach loops (main_target.py:347-352,512-516,886-891).
#include "vs_common.cuh"

namespace {
constexpr int NT = 256;

__global__ void __launch_bounds__(NT) sgd_kernel(float* __restrict__ p, const float* __restrict__ g,
                                                 float* __restrict__ buf, long long count, float lr, float momentum,
                                                 int first, float gscale) {
    for (long long i = (long long)blockIdx.x * NT + threadIdx.x; i < count; i += (long long)gridDim.x * NT) {
        float gi = g[i] * gscale;
        if (momentum != 0.f) {
            const float b = first ? gi : fmaf(momentum, buf[i], gi);
            buf[i] = b;
            gi = b;
        }
        p[i] = p[i] - lr * gi;
    }
}

__global__ void __launch_bounds__(NT) adam_kernel(float* __restrict__ p, const float* __restrict__ g,
                                                  float* __restrict__ m, float* __restrict__ v, long long count,
                                                  float lr, float beta1, float beta2, float eps, float bc1, float bc2,
                                                  float gscale) {
    // torch.optim.Adam (no amsgrad, no wd): p -= lr/bc1 * m / (sqrt(v)/sqrt(bc2) + eps)
    const float step_size = lr / bc1, inv_sqrt_bc2 = rsqrtf(bc2);
    for (long long i = (long long)blockIdx.x * NT + threadIdx.x; i < count; i += (long long)gridDim.x * NT) {
        const float gi = g[i] * gscale;
        const float mi = beta1 * m[i] + (1.f - beta1) * gi;
        const float vi = beta2 * v[i] + (1.f - beta2) * gi * gi;
        m[i] = mi; v[i] = vi;
        p[i] = p[i] - step_size * mi / (sqrtf(vi) * inv_sqrt_bc2 + eps);
    }
}

__global__ void __launch_bounds__(NT) ema_kernel(float* __restrict__ teacher, const float* __restrict__ student,
                                                 long long count, float alpha) {
    for (long long i = (long long)blockIdx.x * NT + threadIdx.x; i < count; i += (long long)gridDim.x * NT)
        teacher[i] = alpha * teacher[i] + (1.f - alpha) * student[i];
}

// dst[r][i] += src[r * src_row_stride + i] with atomics: parameter-gradient fan-in from concurrent streams (per-sample
// backward chains add their partial weight gradients onto the same .grad)
__global__ void __launch_bounds__(NT) atomic_add_rows_kernel(float* __restrict__ dst, const float* __restrict__ src,
                                                             long long rows, long long row_len, long long src_row_stride) {
    const long long total = rows * row_len;
    for (long long i = (long long)blockIdx.x * NT + threadIdx.x; i < total; i += (long long)gridDim.x * NT) {
        const long long r = i / row_len, c = i - r * row_len;
        atomicAdd(dst + i, src[r * src_row_stride + c]);
    }
}

int ew_grid(long long count) { return (int)max(1LL, min((count + NT - 1) / NT, (long long)vs_sm_count() * 8)); }
}  // namespace

extern "C" int vs_sgd_step(float* p, const float* g, float* buf, long long count, float lr, float momentum, int first,
                           float gscale, void* stream) {
    VS_REQUIRE(p && g && count > 0 && (momentum == 0.f || buf), VS_ERR_SHAPE, "sgd_step: bad arguments");
    sgd_kernel<<<ew_grid(count), NT, 0, (cudaStream_t)stream>>>(p, g, buf, count, lr, momentum, first, gscale);
    VS_CHECK_LAUNCH("sgd_kernel");
    return VS_OK;
}

extern "C" int vs_adam_step(float* p, const float* g, float* m, float* v, long long count, float lr, float beta1,
                            float beta2, float eps, int step, float gscale, void* stream) {
    VS_REQUIRE(p && g && m && v && count > 0 && step >= 1, VS_ERR_SHAPE, "adam_step: bad arguments");
    // bias corrections in double on the host, like torch.optim.Adam (1 - 0.999f in fp32 carries ~5e-5 relative error)
    const float bc1 = (float)(1.0 - pow((double)beta1, (double)step)), bc2 = (float)(1.0 - pow((double)beta2, (double)step));
    adam_kernel<<<ew_grid(count), NT, 0, (cudaStream_t)stream>>>(p, g, m, v, count, lr, beta1, beta2, eps, bc1, bc2, gscale);
    VS_CHECK_LAUNCH("adam_kernel");
    return VS_OK;
}

extern "C" int vs_ema_update(float* teacher, const float* student, long long count, float alpha, void* stream) {
    VS_REQUIRE(teacher && student && count > 0, VS_ERR_SHAPE, "ema_update: bad arguments");
    ema_kernel<<<ew_grid(count), NT, 0, (cudaStream_t)stream>>>(teacher, student, count, alpha);
    VS_CHECK_LAUNCH("ema_kernel");
    return VS_OK;
}

extern "C" int vs_atomic_add_rows(float* dst, const float* src, long long rows, long long row_len, long long src_row_stride,
                                  void* stream) {
    VS_REQUIRE(dst && src && rows > 0 && row_len > 0 && src_row_stride >= row_len, VS_ERR_SHAPE, "atomic_add_rows: bad arguments");
    atomic_add_rows_kernel<<<ew_grid(rows * row_len), NT, 0, (cudaStream_t)stream>>>(dst, src, rows, row_len, src_row_stride);
    VS_CHECK_LAUNCH("atomic_add_rows_kernel");
    return VS_OK;
}
