// Dice / KL / pseudo-label losses of utils/evaluation.py (and the eps=1e-4 twin in
// main_source.py:150-182) as fused single-pass reductions over planar fp32 [N][C][S] tensors.
// The reference spends 3 reduction passes + temporaries per avg_dsc call; here one read of
// (src, tgt) yields all three sums, with binarize / confident_binarize / one-hot / argmax
// applied on the fly.  HBM-bound: float4 loads, warp-shuffle + one atomic triple per CTA.
#include "vs_common.cuh"

namespace {

constexpr int NT = 256;

__device__ __forceinline__ float target_transform(float t, int mode) {
    if (mode == VS_TGT_BINARIZE) return t >= 0.5f ? 1.f : 0.f;            // utils/evaluation.py:9-10
    if (mode == VS_TGT_CONFIDENT) return t > 0.8f ? 1.f : (t < 0.2f ? 0.f : t);   // :12-18
    return t;
}

// Loads (s', t') for channel c at voxel v under `mode`.
__device__ __forceinline__ void load_pair(const float* __restrict__ src, const float* __restrict__ tgt, int mode,
                                          int c, long long s, long long v, float& sv, float& tv) {
    if (mode == VS_TGT_ARGMAX) {            // binary=True, utils/evaluation.py:58-64 (n_class = 2; ties -> class 0)
        const int as = src[s + v] > src[v] ? 1 : 0;
        const int at = tgt[s + v] > tgt[v] ? 1 : 0;
        sv = as == c ? 1.f : 0.f;
        tv = at == c ? 1.f : 0.f;
    } else if (mode == VS_TGT_LABEL) {      // one-hot of an integer label volume [N][1][S]
        sv = src[(long long)c * s + v];
        tv = ((int)tgt[v] == c) ? 1.f : 0.f;
    } else {
        sv = src[(long long)c * s + v];
        tv = target_transform(tgt[(long long)c * s + v], mode);
    }
}

__global__ void __launch_bounds__(NT) dice_sums_kernel(const float* __restrict__ src, const float* __restrict__ tgt,
                                                       int mode, float* __restrict__ sums, int c_total, long long s) {
    __shared__ float red[3][NT / 32];
    const int n = blockIdx.z, c = blockIdx.y, t = threadIdx.x;
    const float* ps = src + (long long)n * c_total * s;
    const float* pt = tgt + (long long)n * (mode == VS_TGT_LABEL ? 1 : c_total) * s;
    float I = 0.f, SS = 0.f, TT = 0.f;
    const bool vec = (mode == VS_TGT_TENSOR || mode == VS_TGT_BINARIZE || mode == VS_TGT_CONFIDENT) && (s % 4 == 0) &&
                     vs_aligned16_dev(ps) && vs_aligned16_dev(pt);
    if (vec) {
        const float4* s4 = reinterpret_cast<const float4*>(ps + (long long)c * s);
        const float4* t4 = reinterpret_cast<const float4*>(pt + (long long)c * s);
        for (long long i = (long long)blockIdx.x * NT + t; i < s / 4; i += (long long)gridDim.x * NT) {
            const float4 a = s4[i], b = t4[i];
            const float av[4] = {a.x, a.y, a.z, a.w};
            const float bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const float tv = target_transform(bv[q], mode);
                I = fmaf(av[q], tv, I); SS += av[q]; TT += tv;
            }
        }
    } else {
        for (long long v = (long long)blockIdx.x * NT + t; v < s; v += (long long)gridDim.x * NT) {
            float sv, tv;
            load_pair(ps, pt, mode, c, s, v, sv, tv);
            I = fmaf(sv, tv, I); SS += sv; TT += tv;
        }
    }
    I = warp_sum(I); SS = warp_sum(SS); TT = warp_sum(TT);
    if ((t & 31) == 0) { red[0][t >> 5] = I; red[1][t >> 5] = SS; red[2][t >> 5] = TT; }
    __syncthreads();
    if (t < 3) {
        float acc = 0.f;
        for (int w = 0; w < NT / 32; ++w) acc += red[t][w];
        atomicAdd(sums + ((long long)n * c_total + c) * 3 + t, acc);
    }
}

// gsrc = A*t' - B, gtgt = A*s - B with A = 2 g/D, B = 2 I g/D^2, D = S + T + eps
__global__ void __launch_bounds__(NT) dice_bwd_kernel(const float* __restrict__ src, const float* __restrict__ tgt,
                                                      int mode, const float* __restrict__ sums,
                                                      const float* __restrict__ gper, float eps,
                                                      float* __restrict__ gsrc, float* __restrict__ gtgt,
                                                      int accumulate, int c_total, long long s) {
    const int n = blockIdx.z, c = blockIdx.y;
    const float* sm = sums + ((long long)n * c_total + c) * 3;
    const float g = gper[(long long)n * c_total + c];
    const float D = sm[1] + sm[2] + eps;
    const float A = 2.f * g / D, B = 2.f * sm[0] * g / (D * D);
    const float* ps = src + (long long)n * c_total * s;
    const float* pt = tgt + (long long)n * (mode == VS_TGT_LABEL ? 1 : c_total) * s;
    const long long off = ((long long)n * c_total + c) * s;
    for (long long v = (long long)blockIdx.x * NT + threadIdx.x; v < s; v += (long long)gridDim.x * NT) {
        float sv, tv;
        load_pair(ps, pt, mode, c, s, v, sv, tv);
        if (gsrc != nullptr) {
            const float o = A * tv - B;
            gsrc[off + v] = accumulate ? gsrc[off + v] + o : o;
        }
        if (gtgt != nullptr) {
            const float o = A * sv - B;
            gtgt[off + v] = accumulate ? gtgt[off + v] + o : o;
        }
    }
}

__global__ void kl_fwd_kernel(const float* __restrict__ mean, const float* __restrict__ std, float* __restrict__ out,
                              int batch, int dim) {
    __shared__ float red[NT / 32];
    float acc = 0.f;
    for (int i = threadIdx.x; i < batch * dim; i += NT) {
        const float m = mean[i], sd = std[i];
        acc += sd * sd + m * m - 2.f * logf(sd + 0.00001f);
    }
    acc = warp_sum(acc);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        float tot = 0.f;
        for (int w = 0; w < NT / 32; ++w) tot += red[w];
        out[0] = 0.5f * tot / (float)batch;
    }
}

__global__ void kl_bwd_kernel(const float* __restrict__ mean, const float* __restrict__ std,
                              const float* __restrict__ gout, float* __restrict__ gmean, float* __restrict__ gstd,
                              int batch, int dim) {
    const float g = gout[0] / (float)batch;
    for (int i = blockIdx.x * NT + threadIdx.x; i < batch * dim; i += gridDim.x * NT) {
        gmean[i] = g * mean[i];
        gstd[i] = g * (std[i] - 1.f / (std[i] + 0.00001f));
    }
}

__global__ void __launch_bounds__(NT) binarize_kernel(const float* __restrict__ a, float* __restrict__ out, int mode,
                                                      long long count) {
    for (long long i = (long long)blockIdx.x * NT + threadIdx.x; i < count; i += (long long)gridDim.x * NT)
        out[i] = target_transform(a[i], mode);
}

// Intensity normalisation of the input volumes on the device: out = (clip(x, lo, hi) - sub) / div in IEEE fp32 -- the
// reference's Clip + CenterIntensities dataset transforms (utils/utils.py:508-533,572-618; main_target.py:223-224 uses
// Clip(-200, 400), CenterIntensities(100, 300)).  TI = float, or int16 raw Hounsfield units (half the H2D bytes).
template <typename TI>
__global__ void __launch_bounds__(NT) clip_center_kernel(const TI* __restrict__ x, float* __restrict__ out, long long count,
                                                         float lo, float hi, float sub, float div) {
    for (long long i = (long long)blockIdx.x * NT + threadIdx.x; i < count; i += (long long)gridDim.x * NT) {
        float v = (float)x[i];
        v = fminf(fmaxf(v, lo), hi);
        out[i] = __fdiv_rn(__fsub_rn(v, sub), div);
    }
}

__global__ void __launch_bounds__(NT) one_hot_kernel(const float* __restrict__ label, float* __restrict__ out, int c,
                                                     long long s) {
    const int n = blockIdx.y;
    for (long long v = (long long)blockIdx.x * NT + threadIdx.x; v < s; v += (long long)gridDim.x * NT) {
        const int l = (int)label[(long long)n * s + v];
        for (int k = 0; k < c; ++k) out[((long long)n * c + k) * s + v] = (k == l) ? 1.f : 0.f;
    }
}

// (recon_loss, dsc_loss_fake, klloss) -> final loss and its three scalar weights
__global__ void compose_target_loss_kernel(const float* __restrict__ terms, float lambda_vae, int loss_type, int use_kl,
                                           float* __restrict__ final_loss, float* __restrict__ weights) {
    const float recon = terms[0], fake = terms[1], kl = terms[2];
    float wr, wf, wk;
    if (loss_type == 8) {                       // main_target.py:550-560
        float cur;
        if (recon < 0.15f) cur = lambda_vae * 0.6f;
        else if (recon < 0.225f) cur = lambda_vae * 1.2f;
        else if (recon < 0.3f) cur = lambda_vae * 2.0f;
        else cur = lambda_vae * 3.0f;
        if (cur > 1.f) { wr = 1.f; wf = 1.f / cur; wk = use_kl ? 1.f : 0.f; }
        else { wr = cur; wf = 1.f; wk = use_kl ? cur : 0.f; }
    } else {                                    // main_target.py:588-590
        wr = lambda_vae; wf = 1.f; wk = use_kl ? 0.00002f * lambda_vae : 0.f;
    }
    weights[0] = wr; weights[1] = wf; weights[2] = wk;
    final_loss[0] = wr * recon + wf * fake + wk * kl;
}

// One thread: the three Dice evaluations of a teacher-student step -> scalar losses, the composed final loss and
// the gradient of the final loss w.r.t. every per-(n,c) Dice value (what vs_dice_bwd takes as gper).
//   loss_x = 1 - mean_{n, c in [bot,top)} 2 I/(S+T+eps)     main_target.py:543-546,588-590 (type 0), :550-560 (type 8)
__global__ void joint_target_finish_kernel(const float* __restrict__ sums_r, const float* __restrict__ sums_g,
                                           const float* __restrict__ sums_f, const float* __restrict__ kl, int n, int c,
                                           int bot, int top, float eps, float lambda_vae, int loss_type, int use_kl,
                                           int only_pseudo, float* __restrict__ out, float* __restrict__ gper) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    const float inv = 1.f / (float)(n * (top - bot));
    float loss[3];
    const float* srcs[3] = {sums_r, sums_g, sums_f};
    for (int k = 0; k < 3; ++k) {
        float acc = 0.f;
        if (srcs[k] != nullptr)
            for (int i = 0; i < n; ++i)
                for (int ch = bot; ch < top; ++ch) {
                    const float* sm = srcs[k] + ((long long)i * c + ch) * 3;
                    acc += 2.f * sm[0] / (sm[1] + sm[2] + eps);
                }
        loss[k] = 1.f - acc * inv;
    }
    const float recon = loss[0], fake = loss[2], klv = kl != nullptr ? kl[0] : 0.f;
    float wr, wf, wk;
    if (only_pseudo) { wr = 0.f; wf = 1.f; wk = 0.f; }
    else if (loss_type == 8) {                  // main_target.py:550-560
        float cur;
        if (recon < 0.15f) cur = lambda_vae * 0.6f;
        else if (recon < 0.225f) cur = lambda_vae * 1.2f;
        else if (recon < 0.3f) cur = lambda_vae * 2.0f;
        else cur = lambda_vae * 3.0f;
        if (cur > 1.f) { wr = 1.f; wf = 1.f / cur; wk = use_kl ? 1.f : 0.f; }
        else { wr = cur; wf = 1.f; wk = use_kl ? cur : 0.f; }
    } else {                                    // main_target.py:588-590
        wr = lambda_vae; wf = 1.f; wk = use_kl ? 0.00002f * lambda_vae : 0.f;
    }
    out[0] = wr * recon + wf * fake + wk * klv;
    out[1] = recon; out[2] = loss[1]; out[3] = fake; out[4] = klv;
    for (int i = 0; i < n; ++i)
        for (int ch = 0; ch < c; ++ch) {
            const float g = (ch >= bot && ch < top) ? -inv : 0.f;
            gper[(long long)i * c + ch] = wr * g;
            gper[(long long)(n + i) * c + ch] = wf * g;
        }
}

int red_grid(long long s) { return (int)max(1LL, min((s / 4 + NT - 1) / NT, (long long)vs_sm_count() * 2)); }

}  // namespace

extern "C" int vs_dice_sums(const float* src, const float* tgt, int mode, float* sums, int n, int c, long long s,
                            void* stream) {
    VS_REQUIRE(src && tgt && sums && n > 0 && c > 0 && s > 0, VS_ERR_SHAPE, "dice_sums: bad arguments");
    VS_REQUIRE(mode >= VS_TGT_TENSOR && mode <= VS_TGT_ARGMAX, VS_ERR_UNSUPPORTED, "dice_sums: unknown mode %d", mode);
    VS_REQUIRE(mode != VS_TGT_ARGMAX || c == 2, VS_ERR_UNSUPPORTED, "dice_sums: argmax mode supports n_class == 2 only");
    cudaStream_t st = (cudaStream_t)stream;
    VS_CUDA(cudaMemsetAsync(sums, 0, sizeof(float) * 3 * n * c, st), "dice_sums memset");
    dim3 grid(red_grid(s), c, n);
    dice_sums_kernel<<<grid, NT, 0, st>>>(src, tgt, mode, sums, c, s);
    VS_CHECK_LAUNCH("dice_sums_kernel");
    return VS_OK;
}

extern "C" int vs_dice_bwd(const float* src, const float* tgt, int mode, const float* sums, const float* gper,
                           float eps, float* gsrc, float* gtgt, int accumulate, int n, int c, long long s,
                           void* stream) {
    VS_REQUIRE(src && tgt && sums && gper && n > 0 && c > 0 && s > 0, VS_ERR_SHAPE, "dice_bwd: bad arguments");
    VS_REQUIRE(mode >= VS_TGT_TENSOR && mode <= VS_TGT_LABEL, VS_ERR_UNSUPPORTED, "dice_bwd: mode %d is not differentiable", mode);
    VS_REQUIRE(gtgt == nullptr || mode == VS_TGT_TENSOR, VS_ERR_UNSUPPORTED, "dice_bwd: target gradient only for plain tensor targets");
    dim3 grid((unsigned)max(1LL, min((s + NT - 1) / NT, (long long)vs_sm_count() * 4)), c, n);
    dice_bwd_kernel<<<grid, NT, 0, (cudaStream_t)stream>>>(src, tgt, mode, sums, gper, eps, gsrc, gtgt, accumulate, c, s);
    VS_CHECK_LAUNCH("dice_bwd_kernel");
    return VS_OK;
}

extern "C" int vs_kl_fwd(const float* mean, const float* std, float* out, int batch, int dim, void* stream) {
    VS_REQUIRE(mean && std && out && batch > 0 && dim > 0, VS_ERR_SHAPE, "kl_fwd: bad arguments");
    kl_fwd_kernel<<<1, NT, 0, (cudaStream_t)stream>>>(mean, std, out, batch, dim);
    VS_CHECK_LAUNCH("kl_fwd_kernel");
    return VS_OK;
}

extern "C" int vs_kl_bwd(const float* mean, const float* std, const float* gout, float* gmean, float* gstd, int batch,
                         int dim, void* stream) {
    VS_REQUIRE(mean && std && gout && gmean && gstd && batch > 0 && dim > 0, VS_ERR_SHAPE, "kl_bwd: bad arguments");
    kl_bwd_kernel<<<vs_ceil_div((long long)batch * dim, NT), NT, 0, (cudaStream_t)stream>>>(mean, std, gout, gmean, gstd, batch, dim);
    VS_CHECK_LAUNCH("kl_bwd_kernel");
    return VS_OK;
}

extern "C" int vs_binarize(const float* a, float* out, int mode, long long count, void* stream) {
    VS_REQUIRE(a && out && count > 0, VS_ERR_SHAPE, "binarize: bad arguments");
    VS_REQUIRE(mode == VS_TGT_BINARIZE || mode == VS_TGT_CONFIDENT, VS_ERR_UNSUPPORTED, "binarize: unknown mode %d", mode);
    binarize_kernel<<<(unsigned)max(1LL, min((count + NT - 1) / NT, (long long)vs_sm_count() * 8)), NT, 0, (cudaStream_t)stream>>>(a, out, mode, count);
    VS_CHECK_LAUNCH("binarize_kernel");
    return VS_OK;
}

extern "C" int vs_clip_center(int in_kind, const void* x, float* out, long long count, float lo, float hi, float sub,
                              float div, void* stream) {
    VS_REQUIRE(x && out && count > 0 && hi >= lo && div != 0.f, VS_ERR_SHAPE, "clip_center: bad arguments");
    VS_REQUIRE(in_kind == 0 || in_kind == 2, VS_ERR_UNSUPPORTED, "clip_center: input kind %d (0 = fp32, 2 = int16)", in_kind);
    const unsigned grid = (unsigned)max(1LL, min((count + NT - 1) / NT, (long long)vs_sm_count() * 8));
    if (in_kind == 0) clip_center_kernel<float><<<grid, NT, 0, (cudaStream_t)stream>>>((const float*)x, out, count, lo, hi, sub, div);
    else clip_center_kernel<short><<<grid, NT, 0, (cudaStream_t)stream>>>((const short*)x, out, count, lo, hi, sub, div);
    VS_CHECK_LAUNCH("clip_center_kernel");
    return VS_OK;
}

extern "C" int vs_one_hot(const float* label, float* out, int n, int c, long long s, void* stream) {
    VS_REQUIRE(label && out && n > 0 && c > 0 && s > 0, VS_ERR_SHAPE, "one_hot: bad arguments");
    dim3 grid((unsigned)max(1LL, min((s + NT - 1) / NT, (long long)vs_sm_count() * 8)), n);
    one_hot_kernel<<<grid, NT, 0, (cudaStream_t)stream>>>(label, out, c, s);
    VS_CHECK_LAUNCH("one_hot_kernel");
    return VS_OK;
}

extern "C" int vs_joint_target_finish(const float* sums_recon, const float* sums_gt, const float* sums_fake, const float* kl,
                                      int n, int c, int bot, int top, float eps, float lambda_vae, int loss_type,
                                      int use_kl, int only_pseudo, float* out5, float* gper2, void* stream) {
    VS_REQUIRE(sums_recon && sums_fake && out5 && gper2 && n > 0 && c > 0 && bot >= 0 && top > bot && top <= c, VS_ERR_SHAPE,
               "joint_target_finish: bad arguments");
    joint_target_finish_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(sums_recon, sums_gt, sums_fake, kl, n, c, bot, top, eps,
                                                                  lambda_vae, loss_type, use_kl, only_pseudo, out5, gper2);
    VS_CHECK_LAUNCH("joint_target_finish_kernel");
    return VS_OK;
}

extern "C" int vs_compose_target_loss(const float* terms, float lambda_vae, int loss_type, int use_kl,
                                      float* final_loss, float* weights, void* stream) {
    VS_REQUIRE(terms && final_loss && weights, VS_ERR_SHAPE, "compose_target_loss: null pointer");
    compose_target_loss_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(terms, lambda_vae, loss_type, use_kl, final_loss, weights);
    VS_CHECK_LAUNCH("compose_target_loss_kernel");
    return VS_OK;
}
