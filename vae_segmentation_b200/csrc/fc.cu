// Linear layers of the shape VAE + reparameterisation (joint_model.py:216-218,241-250).
// M = batch is tiny, so these are weight-bandwidth-bound GEMV-like kernels: fc_mean and
// fc_std share one read of x, the reparameterisation z_lat = mean + z*std*scale is fused,
// and activations are read/written NDHWC while honouring the reference's NCDHW flatten
// order i = c*S3 + v.  Replaces cuBLAS addmm + ATen relu/elementwise.
#include "vs_common.cuh"

namespace {

constexpr int NT = 256;
constexpr int MAXB = 8;          // batch rows handled per pass

template <typename T>
__device__ __forceinline__ float load_flat(const T* __restrict__ x, int b, int i, int s3, int c) {
    const int ch = i / s3, v = i - ch * s3;
    return Store<T>::ld(x + ((long long)b * s3 + v) * c + ch);
}

// one CTA per latent index j
template <typename T>
__global__ void __launch_bounds__(NT) fc_encode_fwd_kernel(const T* __restrict__ x, const float* __restrict__ wm,
                                                           const float* __restrict__ bm, const float* __restrict__ ws,
                                                           const float* __restrict__ bs, const float* __restrict__ z,
                                                           float scale, int use_z, float* __restrict__ mean,
                                                           float* __restrict__ std, float* __restrict__ lat,
                                                           int batch, int s3, int c, int dim) {
    __shared__ float red[2][MAXB][NT / 32];
    const int j = blockIdx.x, t = threadIdx.x;
    const int flat = s3 * c;
    for (int b0 = 0; b0 < batch; b0 += MAXB) {
        const int nb = min(MAXB, batch - b0);
        float am[MAXB], as[MAXB];
#pragma unroll
        for (int b = 0; b < MAXB; ++b) { am[b] = 0.f; as[b] = 0.f; }
        for (int i = t; i < flat; i += NT) {
            const float a = wm[(long long)j * flat + i], s = ws[(long long)j * flat + i];
#pragma unroll
            for (int b = 0; b < MAXB; ++b) {
                if (b < nb) {
                    const float xv = load_flat(x, b0 + b, i, s3, c);
                    am[b] = fmaf(a, xv, am[b]);
                    as[b] = fmaf(s, xv, as[b]);
                }
            }
        }
        __syncthreads();
#pragma unroll
        for (int b = 0; b < MAXB; ++b) {
            const float m = warp_sum(am[b]), s = warp_sum(as[b]);
            if ((t & 31) == 0) { red[0][b][t >> 5] = m; red[1][b][t >> 5] = s; }
        }
        __syncthreads();
        if (t < nb) {
            float m = bm[j], s = bs[j];
            for (int w = 0; w < NT / 32; ++w) { m += red[0][t][w]; s += red[1][t][w]; }
            s = fmaxf(s, 0.f);
            const long long o = (long long)(b0 + t) * dim + j;
            mean[o] = m;
            std[o] = s;
            lat[o] = use_z ? m + z[o] * s * scale : m;
        }
    }
}

// one warp per flat row i: h[b][i] = b2[i] + sum_j lat[b][j] w2[i][j]
template <typename T>
__global__ void __launch_bounds__(NT) fc_decode_fwd_kernel(const float* __restrict__ lat, const float* __restrict__ w2,
                                                           const float* __restrict__ b2, T* __restrict__ h,
                                                           int batch, int s3, int c, int dim) {
    extern __shared__ float slat[];           // [batch][dim]
    for (int i = threadIdx.x; i < batch * dim; i += NT) slat[i] = lat[i];
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int flat = s3 * c;
    for (int i = blockIdx.x * (NT / 32) + warp; i < flat; i += gridDim.x * (NT / 32)) {
        const int ch = i / s3, v = i - ch * s3;
        for (int b0 = 0; b0 < batch; b0 += MAXB) {
            const int nb = min(MAXB, batch - b0);
            float acc[MAXB];
#pragma unroll
            for (int b = 0; b < MAXB; ++b) acc[b] = 0.f;
            for (int j = lane; j < dim; j += 32) {
                const float wv = w2[(long long)i * dim + j];
#pragma unroll
                for (int b = 0; b < MAXB; ++b) if (b < nb) acc[b] = fmaf(wv, slat[(b0 + b) * dim + j], acc[b]);
            }
#pragma unroll
            for (int b = 0; b < MAXB; ++b) {
                const float r = warp_sum(acc[b]);
                if (lane == 0 && b < nb) Store<T>::st(h + ((long long)(b0 + b) * s3 + v) * c + ch, r + b2[i]);
            }
        }
    }
}

// dlat[b][j] += sum_i dh[b][i] w2[i][j];  dw2[i][j] (+)= sum_b dh[b][i] lat[b][j];  db2[i] (+)= sum_b dh[b][i]
template <typename T>
__global__ void __launch_bounds__(NT) fc_decode_bwd_kernel(const T* __restrict__ dh, const float* __restrict__ lat,
                                                           const float* __restrict__ w2, float* __restrict__ dlat,
                                                           float* __restrict__ dw2, float* __restrict__ db2,
                                                           int accumulate, int batch, int s3, int c, int dim) {
    extern __shared__ float sm[];             // slat[batch][dim], sdl[batch][dim]
    float* slat = sm;
    float* sdl = sm + batch * dim;
    for (int i = threadIdx.x; i < batch * dim; i += NT) { slat[i] = lat[i]; sdl[i] = 0.f; }
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int flat = s3 * c;
    for (int i = blockIdx.x * (NT / 32) + warp; i < flat; i += gridDim.x * (NT / 32)) {
        const int ch = i / s3, v = i - ch * s3;
        float gsum = 0.f;
        for (int b = 0; b < batch; ++b) {
            const float g = Store<T>::ld(dh + ((long long)b * s3 + v) * c + ch);
            gsum += g;
            for (int j = lane; j < dim; j += 32) atomicAdd(&sdl[b * dim + j], g * w2[(long long)i * dim + j]);
        }
        if (dw2 != nullptr) {
            for (int j = lane; j < dim; j += 32) {
                float acc = 0.f;
                for (int b = 0; b < batch; ++b)
                    acc = fmaf(Store<T>::ld(dh + ((long long)b * s3 + v) * c + ch), slat[b * dim + j], acc);
                const long long o = (long long)i * dim + j;
                dw2[o] = accumulate ? dw2[o] + acc : acc;
            }
            if (lane == 0 && db2 != nullptr) db2[i] = accumulate ? db2[i] + gsum : gsum;
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < batch * dim; i += NT) atomicAdd(dlat + i, sdl[i]);
}

// gbuf[0][b][j] = gm, gbuf[1][b][j] = gs
__global__ void fc_encode_bwd_prep_kernel(const float* __restrict__ z, float scale, int use_z,
                                          const float* __restrict__ std, const float* __restrict__ dlat,
                                          const float* __restrict__ gmean_ext, const float* __restrict__ gstd_ext,
                                          float* __restrict__ gbuf, float* __restrict__ dbm, float* __restrict__ dbs,
                                          int accumulate, int batch, int dim) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= dim) return;
    float sm_ = 0.f, ss_ = 0.f;
    for (int b = 0; b < batch; ++b) {
        const int o = b * dim + j;
        const float dl = dlat ? dlat[o] : 0.f;
        float gm = dl + (gmean_ext ? gmean_ext[o] : 0.f);
        float gs = (use_z ? dl * z[o] * scale : 0.f) + (gstd_ext ? gstd_ext[o] : 0.f);
        gs = std[o] > 0.f ? gs : 0.f;
        gbuf[o] = gm;
        gbuf[batch * dim + o] = gs;
        sm_ += gm; ss_ += gs;
    }
    if (dbm) dbm[j] = accumulate ? dbm[j] + sm_ : sm_;
    if (dbs) dbs[j] = accumulate ? dbs[j] + ss_ : ss_;
}

// thread per flat index i: dx[b][i] = sum_j gm[b][j] wm[j][i] + gs[b][j] ws[j][i]
// and (optionally) dwm[j][i] (+)= sum_b gm[b][j] x[b][i], dws likewise.
template <typename T>
__global__ void __launch_bounds__(NT) fc_encode_bwd_kernel(const T* __restrict__ x, const float* __restrict__ wm,
                                                           const float* __restrict__ ws, const float* __restrict__ gbuf,
                                                           T* __restrict__ dx, float* __restrict__ dwm,
                                                           float* __restrict__ dws, int accumulate, int batch, int s3,
                                                           int c, int dim) {
    extern __shared__ float sg[];             // [2][batch][dim]
    for (int i = threadIdx.x; i < 2 * batch * dim; i += NT) sg[i] = gbuf[i];
    __syncthreads();
    const int flat = s3 * c;
    const int i = blockIdx.x * NT + threadIdx.x;
    if (i >= flat) return;
    const int ch = i / s3, v = i - ch * s3;
    const float* gm = sg;
    const float* gs = sg + batch * dim;
    for (int b0 = 0; b0 < batch; b0 += MAXB) {
        const int nb = min(MAXB, batch - b0);
        float acc[MAXB], xv[MAXB];
#pragma unroll
        for (int b = 0; b < MAXB; ++b) { acc[b] = 0.f; xv[b] = (b < nb && dwm) ? Store<T>::ld(x + ((long long)(b0 + b) * s3 + v) * c + ch) : 0.f; }
        for (int j = 0; j < dim; ++j) {
            const long long o = (long long)j * flat + i;
            const float a = wm[o], s = ws[o];
            float dwa = 0.f, dwb = 0.f;
#pragma unroll
            for (int b = 0; b < MAXB; ++b) {
                if (b < nb) {
                    const float g0 = gm[(b0 + b) * dim + j], g1 = gs[(b0 + b) * dim + j];
                    acc[b] = fmaf(g0, a, fmaf(g1, s, acc[b]));
                    dwa = fmaf(g0, xv[b], dwa);
                    dwb = fmaf(g1, xv[b], dwb);
                }
            }
            if (dwm != nullptr) {
                const bool add = accumulate || b0 > 0;
                dwm[o] = add ? dwm[o] + dwa : dwa;
                dws[o] = add ? dws[o] + dwb : dwb;
            }
        }
        if (dx != nullptr) {
#pragma unroll
            for (int b = 0; b < MAXB; ++b)
                if (b < nb) Store<T>::st(dx + ((long long)(b0 + b) * s3 + v) * c + ch, acc[b]);
        }
    }
}

int check_fc(int batch, int s3, int c, int dim, const char* who) {
    VS_REQUIRE(batch > 0 && s3 > 0 && c > 0 && dim > 0, VS_ERR_SHAPE, "%s: bad shape", who);
    VS_REQUIRE((size_t)2 * batch * dim * sizeof(float) <= 40 * 1024, VS_ERR_UNSUPPORTED, "%s: batch*dim too large for the shared-memory stage (batch=%d dim=%d)", who, batch, dim);
    return VS_OK;
}

}  // namespace

extern "C" int vs_fc_encode_fwd(int dtype, const void* x, const float* wm, const float* bm, const float* ws,
                                const float* bs, const float* z, float scale, int use_z, float* mean, float* std,
                                float* lat, int batch, int s3, int c, int dim, void* stream) {
    int rc = check_fc(batch, s3, c, dim, "fc_encode_fwd");
    if (rc) return rc;
    VS_REQUIRE(x && wm && bm && ws && bs && mean && std && lat && (!use_z || z), VS_ERR_SHAPE, "fc_encode_fwd: null pointer");
    VS_DISPATCH_DTYPE(dtype, T, { fc_encode_fwd_kernel<T><<<dim, NT, 0, (cudaStream_t)stream>>>(
        (const T*)x, wm, bm, ws, bs, z, scale, use_z, mean, std, lat, batch, s3, c, dim); });
    VS_CHECK_LAUNCH("fc_encode_fwd_kernel");
    return VS_OK;
}

extern "C" int vs_fc_decode_fwd(int dtype, const float* lat, const float* w2, const float* b2, void* h, int batch,
                                int s3, int c, int dim, void* stream) {
    int rc = check_fc(batch, s3, c, dim, "fc_decode_fwd");
    if (rc) return rc;
    VS_REQUIRE(lat && w2 && b2 && h, VS_ERR_SHAPE, "fc_decode_fwd: null pointer");
    const int flat = s3 * c;
    const int blocks = min(vs_ceil_div(flat, NT / 32), vs_sm_count() * 8);
    const size_t smem = sizeof(float) * batch * dim;
    VS_DISPATCH_DTYPE(dtype, T, { fc_decode_fwd_kernel<T><<<blocks, NT, smem, (cudaStream_t)stream>>>(
        lat, w2, b2, (T*)h, batch, s3, c, dim); });
    VS_CHECK_LAUNCH("fc_decode_fwd_kernel");
    return VS_OK;
}

extern "C" int vs_fc_decode_bwd(int dtype, const void* dh, const float* lat, const float* w2, float* dlat, float* dw2,
                                float* db2, int accumulate, int batch, int s3, int c, int dim, void* stream) {
    int rc = check_fc(batch, s3, c, dim, "fc_decode_bwd");
    if (rc) return rc;
    VS_REQUIRE(dh && lat && w2 && dlat, VS_ERR_SHAPE, "fc_decode_bwd: null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    VS_CUDA(cudaMemsetAsync(dlat, 0, sizeof(float) * batch * dim, st), "fc_decode_bwd memset");
    const int flat = s3 * c;
    const int blocks = min(vs_ceil_div(flat, NT / 32), vs_sm_count() * 2);
    const size_t smem = sizeof(float) * 2 * batch * dim;
    VS_DISPATCH_DTYPE(dtype, T, { fc_decode_bwd_kernel<T><<<blocks, NT, smem, st>>>(
        (const T*)dh, lat, w2, dlat, dw2, db2, accumulate, batch, s3, c, dim); });
    VS_CHECK_LAUNCH("fc_decode_bwd_kernel");
    return VS_OK;
}

extern "C" int vs_fc_encode_bwd(int dtype, const void* x, const float* wm, const float* ws, const float* z,
                                float scale, int use_z, const float* std, const float* dlat, const float* gmean_ext,
                                const float* gstd_ext, float* gbuf, void* dx, float* dwm, float* dbm, float* dws,
                                float* dbs, int accumulate, int batch, int s3, int c, int dim, void* stream) {
    int rc = check_fc(batch, s3, c, dim, "fc_encode_bwd");
    if (rc) return rc;
    VS_REQUIRE(wm && ws && std && gbuf && (!use_z || z), VS_ERR_SHAPE, "fc_encode_bwd: null pointer");
    VS_REQUIRE((dwm == nullptr) == (dws == nullptr), VS_ERR_SHAPE, "fc_encode_bwd: dwm and dws must be given together");
    VS_REQUIRE(dwm == nullptr || x != nullptr, VS_ERR_SHAPE, "fc_encode_bwd: weight gradients need x");
    cudaStream_t st = (cudaStream_t)stream;
    fc_encode_bwd_prep_kernel<<<vs_ceil_div(dim, 128), 128, 0, st>>>(z, scale, use_z, std, dlat, gmean_ext, gstd_ext,
                                                                     gbuf, dbm, dbs, accumulate, batch, dim);
    VS_CHECK_LAUNCH("fc_encode_bwd_prep_kernel");
    if (dx == nullptr && dwm == nullptr) return VS_OK;
    const int flat = s3 * c;
    const size_t smem = sizeof(float) * 2 * batch * dim;
    VS_DISPATCH_DTYPE(dtype, T, { fc_encode_bwd_kernel<T><<<vs_ceil_div(flat, NT), NT, smem, st>>>(
        (const T*)x, wm, ws, gbuf, (T*)dx, dwm, dws, accumulate, batch, s3, c, dim); });
    VS_CHECK_LAUNCH("fc_encode_bwd_kernel");
    return VS_OK;
}

// ---- generic Linear (+ activation) for the discriminator / encoder head (joint_model.py:287-304: fc1, fc2, fc_mean) --------
// y[b][o] = act(bias[o] + sum_i W[o][i] x[b][i]); act 0 = identity, 1 = ReLU, 2 = sigmoid.  Batch <= 8 rows per pass:
// weight-bandwidth-bound (W is read once), one warp per output row.
namespace {

__device__ __forceinline__ float act_fwd(float v, int act) {
    if (act == 1) return fmaxf(v, 0.f);
    if (act == 2) return 1.f / (1.f + expf(-v));
    return v;
}
// derivative expressed through the OUTPUT y (what the backward pass has at hand)
__device__ __forceinline__ float act_bwd(float y, int act) {
    if (act == 1) return y > 0.f ? 1.f : 0.f;
    if (act == 2) return y * (1.f - y);
    return 1.f;
}

__global__ void __launch_bounds__(NT) linear_fwd_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                        const float* __restrict__ bias, float* __restrict__ y, int batch,
                                                        int in_f, int out_f, int act) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int o = blockIdx.x * (NT / 32) + warp; o < out_f; o += gridDim.x * (NT / 32)) {
        const float* wr = w + (long long)o * in_f;
        for (int b0 = 0; b0 < batch; b0 += MAXB) {
            const int nb = min(MAXB, batch - b0);
            float acc[MAXB];
#pragma unroll
            for (int b = 0; b < MAXB; ++b) acc[b] = 0.f;
            for (int i = lane; i < in_f; i += 32) {
                const float wv = wr[i];
#pragma unroll
                for (int b = 0; b < MAXB; ++b)
                    if (b < nb) acc[b] = fmaf(wv, x[(long long)(b0 + b) * in_f + i], acc[b]);
            }
#pragma unroll
            for (int b = 0; b < MAXB; ++b) {
                const float s = warp_sum(acc[b]);
                if (lane == 0 && b < nb) y[(long long)(b0 + b) * out_f + o] = act_fwd(s + (bias ? bias[o] : 0.f), act);
            }
        }
    }
}

// g[b][o] = dy[b][o] * act'(y[b][o]);  db[o] (+)= sum_b g
__global__ void __launch_bounds__(NT) linear_bwd_prep_kernel(const float* __restrict__ dy, const float* __restrict__ y,
                                                             float* __restrict__ g, float* __restrict__ db, int accumulate,
                                                             int batch, int out_f, int act) {
    const int o = blockIdx.x * NT + threadIdx.x;
    if (o >= out_f) return;
    float s = 0.f;
    for (int b = 0; b < batch; ++b) {
        const float v = dy[(long long)b * out_f + o] * act_bwd(y[(long long)b * out_f + o], act);
        g[(long long)b * out_f + o] = v;
        s += v;
    }
    if (db != nullptr) db[o] = accumulate ? db[o] + s : s;
}

// dx[b][i] = sum_o g[b][o] W[o][i]  and  dW[o][i] (+)= sum_b g[b][o] x[b][i]; thread = input column i (coalesced over W rows)
__global__ void __launch_bounds__(NT) linear_bwd_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                        const float* __restrict__ g, float* __restrict__ dx,
                                                        float* __restrict__ dw, int accumulate, int batch, int in_f, int out_f) {
    const int i = blockIdx.x * NT + threadIdx.x;
    if (i >= in_f) return;
    for (int b0 = 0; b0 < batch; b0 += MAXB) {
        const int nb = min(MAXB, batch - b0);
        float xv[MAXB], acc[MAXB];
#pragma unroll
        for (int b = 0; b < MAXB; ++b) { xv[b] = (b < nb && x != nullptr) ? x[(long long)(b0 + b) * in_f + i] : 0.f; acc[b] = 0.f; }
        for (int o = 0; o < out_f; ++o) {
            const float wv = w[(long long)o * in_f + i];
            float dwv = 0.f;
#pragma unroll
            for (int b = 0; b < MAXB; ++b) {
                if (b < nb) {
                    const float gv = g[(long long)(b0 + b) * out_f + o];          // warp-uniform address: broadcast
                    acc[b] = fmaf(gv, wv, acc[b]);
                    dwv = fmaf(gv, xv[b], dwv);
                }
            }
            if (dw != nullptr) {
                float* p = dw + (long long)o * in_f + i;
                *p = (accumulate || b0 > 0) ? *p + dwv : dwv;
            }
        }
        if (dx != nullptr) {
#pragma unroll
            for (int b = 0; b < MAXB; ++b)
                if (b < nb) dx[(long long)(b0 + b) * in_f + i] = acc[b];
        }
    }
}

}  // namespace

extern "C" int vs_linear_fwd(const float* x, const float* w, const float* bias, float* y, int batch, int in_f, int out_f,
                             int act, void* stream) {
    VS_REQUIRE(x && w && y && batch > 0 && in_f > 0 && out_f > 0 && act >= 0 && act <= 2, VS_ERR_SHAPE, "linear_fwd: bad arguments");
    const int blocks = min(vs_ceil_div(out_f, NT / 32), vs_sm_count() * 8);
    linear_fwd_kernel<<<blocks, NT, 0, (cudaStream_t)stream>>>(x, w, bias, y, batch, in_f, out_f, act);
    VS_CHECK_LAUNCH("linear_fwd_kernel");
    return VS_OK;
}

extern "C" int vs_linear_bwd(const float* x, const float* w, const float* y, const float* dy, float* gbuf, float* dx, float* dw,
                             float* db, int accumulate, int batch, int in_f, int out_f, int act, void* stream) {
    VS_REQUIRE(w && y && dy && gbuf && batch > 0 && in_f > 0 && out_f > 0 && act >= 0 && act <= 2, VS_ERR_SHAPE, "linear_bwd: bad arguments");
    VS_REQUIRE(dw == nullptr || x != nullptr, VS_ERR_SHAPE, "linear_bwd: the weight gradient needs x");
    cudaStream_t st = (cudaStream_t)stream;
    linear_bwd_prep_kernel<<<vs_ceil_div(out_f, NT), NT, 0, st>>>(dy, y, gbuf, db, accumulate, batch, out_f, act);
    VS_CHECK_LAUNCH("linear_bwd_prep_kernel");
    if (dx == nullptr && dw == nullptr) return VS_OK;
    linear_bwd_kernel<<<vs_ceil_div(in_f, NT), NT, 0, st>>>(x, w, gbuf, dx, dw, accumulate, batch, in_f, out_f);
    VS_CHECK_LAUNCH("linear_bwd_kernel");
    return VS_OK;
}
