// CropResize on the device: the reference's per-sample geometric transform
//   utils/utils.py:220-293  crop a cube around the pancreas bounding box, zero-pad it back to a cube (centred by diff/2,
//   a reference quirk that is reproduced), then skimage.transform.resize(img, output_size)             [order 1, anti-aliased]
//                                              skimage.transform.resize(label, output_size, order=0, anti_aliasing=False)
// skimage 0.18.3 (requirements.txt:94) resizes n-D volumes as: optional Gaussian pre-filter with sigma = max(0, (s-1)/2) per
// axis (s = input/output extent; scipy.ndimage.gaussian_filter, mode 'mirror', truncate 4, float32 between the three passes),
// then scipy.ndimage.map_coordinates at x = s * (i + 0.5) - 0.5, order 1 (linear) or 0 (nearest = floor(x + 0.5)), mode
// 'mirror'.  Both are restated here in double-precision index / weight arithmetic (ndimage computes in double and stores the
// input dtype), one thread per output voxel; HBM-bound streaming kernels.  The cube is never materialised: `cube_at` maps a
// cube coordinate to the source volume (or to the zero padding).
#include "vs_common.cuh"

namespace {

constexpr int NT = 256;

struct CubeMap {
    const float* src;           // source volume [sd][sh][sw]
    int sd, sh, sw;
    int c0[3];                  // first source index of the crop per axis (already clamped to >= 0)
    int len[3];                 // crop length per axis (already clamped to the volume)
    int pad[3];                 // leading zero padding per axis (int(diff / 2))
    int L;                      // cube side
};

__device__ __forceinline__ float cube_at(const CubeMap& m, int i, int j, int k) {
    const int a = i - m.pad[0], b = j - m.pad[1], c = k - m.pad[2];
    if (a < 0 || a >= m.len[0] || b < 0 || b >= m.len[1] || c < 0 || c >= m.len[2]) return 0.f;
    return m.src[((long long)(m.c0[0] + a) * m.sh + (m.c0[1] + b)) * m.sw + (m.c0[2] + c)];
}

// scipy.ndimage 'mirror': d c b | a b c d | c b a  (reflection about the centre of the edge samples)
__device__ __forceinline__ int mirror(int i, int n) {
    if (n == 1) return 0;
    const int period = 2 * n - 2;
    i %= period;
    if (i < 0) i += period;
    return i < n ? i : period - i;
}

// one separable Gaussian pass along `axis` over an L^3 cube; in == nullptr reads the cube through `m`
__global__ void __launch_bounds__(NT) gauss_pass_kernel(CubeMap m, const float* __restrict__ in, float* __restrict__ out,
                                                        int axis, int radius, double sigma) {
    // normalised weights as scipy.ndimage._gaussian_kernel1d computes them: exp(-0.5 t^2 / sigma^2) / sum, in double
    __shared__ double w[129];
    if (threadIdx.x <= 2 * radius) w[threadIdx.x] = exp(-0.5 / (sigma * sigma) * (double)((int)threadIdx.x - radius) * ((int)threadIdx.x - radius));
    __syncthreads();
    double wsum = 0.0;
    for (int t = 0; t <= 2 * radius; ++t) wsum += w[t];
    const double winv = 1.0 / wsum;
    const long long total = (long long)m.L * m.L * m.L;
    for (long long idx = (long long)blockIdx.x * NT + threadIdx.x; idx < total; idx += (long long)gridDim.x * NT) {
        const int k = (int)(idx % m.L), j = (int)((idx / m.L) % m.L), i = (int)(idx / ((long long)m.L * m.L));
        double acc = 0.0;
        for (int t = -radius; t <= radius; ++t) {
            int ii = i, jj = j, kk = k;
            if (axis == 0) ii = mirror(i + t, m.L); else if (axis == 1) jj = mirror(j + t, m.L); else kk = mirror(k + t, m.L);
            const float v = in != nullptr ? in[((long long)ii * m.L + jj) * m.L + kk] : cube_at(m, ii, jj, kk);
            acc += (w[t + radius] * winv) * (double)v;
        }
        out[idx] = (float)acc;
    }
}

// out[od][oh][ow] = resample(cube) at x = s * (i + 0.5) - 0.5; order 1: trilinear, order 0: nearest
__global__ void __launch_bounds__(NT) resample_kernel(CubeMap m, const float* __restrict__ in, float* __restrict__ out,
                                                      int od, int oh, int ow, int order) {
    const long long total = (long long)od * oh * ow;
    const double s0 = (double)m.L / od, s1 = (double)m.L / oh, s2 = (double)m.L / ow;
    for (long long idx = (long long)blockIdx.x * NT + threadIdx.x; idx < total; idx += (long long)gridDim.x * NT) {
        const int k = (int)(idx % ow), j = (int)((idx / ow) % oh), i = (int)(idx / ((long long)ow * oh));
        const double x0 = s0 * ((double)i + 0.5) - 0.5, x1 = s1 * ((double)j + 0.5) - 0.5, x2 = s2 * ((double)k + 0.5) - 0.5;
        if (order == 0) {
            const int a = mirror((int)floor(x0 + 0.5), m.L), b = mirror((int)floor(x1 + 0.5), m.L), c = mirror((int)floor(x2 + 0.5), m.L);
            out[idx] = in != nullptr ? in[((long long)a * m.L + b) * m.L + c] : cube_at(m, a, b, c);
            continue;
        }
        const int a0 = (int)floor(x0), b0 = (int)floor(x1), c0 = (int)floor(x2);
        const double ta = x0 - a0, tb = x1 - b0, tc = x2 - c0;
        double acc = 0.0;
#pragma unroll
        for (int da = 0; da < 2; ++da)
#pragma unroll
            for (int db = 0; db < 2; ++db)
#pragma unroll
                for (int dc = 0; dc < 2; ++dc) {
                    const int a = mirror(a0 + da, m.L), b = mirror(b0 + db, m.L), c = mirror(c0 + dc, m.L);
                    const float v = in != nullptr ? in[((long long)a * m.L + b) * m.L + c] : cube_at(m, a, b, c);
                    acc += (da ? ta : 1.0 - ta) * (db ? tb : 1.0 - tb) * (dc ? tc : 1.0 - tc) * (double)v;
                }
        out[idx] = (float)acc;
    }
}

unsigned ew_grid(long long count) { return (unsigned)max(1LL, min((count + NT - 1) / NT, (long long)vs_sm_count() * 16)); }

}  // namespace

// Radius of scipy.ndimage.gaussian_filter1d for a given sigma (truncate = 4): int(4 * sigma + 0.5)
extern "C" int vs_gauss_radius(double sigma) { return sigma > 1e-15 ? (int)(4.0 * sigma + 0.5) : 0; }

extern "C" int vs_crop_resize(const float* src, int sd, int sh, int sw, const int* crop9_host, int side, float* out, int od,
                              int oh, int ow, int order, int anti_alias, float* tmp0, float* tmp1, void* stream) {
    VS_REQUIRE(src && out && crop9_host && sd > 0 && sh > 0 && sw > 0 && side > 0 && od > 0 && oh > 0 && ow > 0, VS_ERR_SHAPE,
               "crop_resize: bad arguments");
    VS_REQUIRE(order == 0 || order == 1, VS_ERR_UNSUPPORTED, "crop_resize: order %d (0 = nearest, 1 = linear)", order);
    CubeMap m;
    m.src = src; m.sd = sd; m.sh = sh; m.sw = sw; m.L = side;
    const int dims[3] = {sd, sh, sw};
    for (int a = 0; a < 3; ++a) {
        m.c0[a] = crop9_host[a]; m.len[a] = crop9_host[3 + a]; m.pad[a] = crop9_host[6 + a];
        VS_REQUIRE(m.c0[a] >= 0 && m.len[a] >= 0 && m.c0[a] + m.len[a] <= dims[a] && m.pad[a] >= 0 && m.pad[a] + m.len[a] <= side,
                   VS_ERR_SHAPE, "crop_resize: crop window out of range on axis %d", a);
    }
    cudaStream_t st = (cudaStream_t)stream;
    const long long cube = (long long)side * side * side;
    const float* cur = nullptr;                 // nullptr = read the cube through the crop map
    if (anti_alias) {
        const int outs[3] = {od, oh, ow};
        for (int a = 0; a < 3; ++a) {
            const double sigma = fmax(0.0, ((double)side / outs[a] - 1.0) / 2.0);
            const int radius = vs_gauss_radius(sigma);
            if (radius > 0 || sigma > 1e-15) {
                VS_REQUIRE(tmp0 && tmp1, VS_ERR_SHAPE, "crop_resize: the anti-aliasing filter needs the two cube workspaces");
                VS_REQUIRE(radius <= 64, VS_ERR_UNSUPPORTED, "crop_resize: filter radius %d too large", radius);
                float* dst = (cur == tmp0) ? tmp1 : tmp0;
                gauss_pass_kernel<<<ew_grid(cube), NT, 0, st>>>(m, cur, dst, a, radius, sigma);
                VS_CHECK_LAUNCH("gauss_pass_kernel");
                cur = dst;
            }
        }
    }
    resample_kernel<<<ew_grid((long long)od * oh * ow), NT, 0, st>>>(m, cur, out, od, oh, ow, order);
    VS_CHECK_LAUNCH("resample_kernel");
    return VS_OK;
}
