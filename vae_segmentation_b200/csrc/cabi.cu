// C-ABI glue: error string, device query, and the dispatch between the tcgen05 and the
// CUDA-core convolution paths.  No CPU fallback exists anywhere in this library.
#include <stdarg.h>
#include <string.h>
#include "vs_common.cuh"

static thread_local char g_err[512] = "";
int g_vs_pdl = 0;

void vs_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int vs_sm_count() {
    static int sms[64] = {0};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
    if (sms[dev] == 0) {
        int v = 0;
        if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v <= 0) v = 148;
        sms[dev] = v;
    }
    return sms[dev];
}

extern "C" int vs_conv3x3x3_fprop_direct(int in_dtype, int out_dtype, int in_planar, int out_planar, const void* x,
                                         const float* wpk, const float* bias, void* y, double* stats, float* shift,
                                         int n, int d, int h, int w, int cin, int cout, void* stream);

extern "C" int vs_conv3_shift_internal(int in_dtype, int in_planar, const void* x, const float* wpk, float* shift, int n,
                                       int d, int h, int w, int cin, int cout, void* stream);
#ifdef VS_WITH_TCGEN05
extern "C" int vs_conv3x3x3_tc(const void* x, const void* wtc, void* y, double* stats, float* shift, int prezeroed,
                               const void* yprev, const double* pstats, double* psums, int planar_mode,
                               float* yplanar, const float* bias, int n, int d,
                               int h, int w, int gin, int gout, void* stream);
extern "C" size_t vs_conv3_tc_pack_bytes(int cin, int cout, int dgrad);
#endif

static bool tc_eligible(int gin, int gout) {
    return (gin == 8 || (gin % 16 == 0 && gin >= 16)) && gout % 8 == 0 && gout >= 8;
}

extern "C" const char* vs_last_error_string(void) { return g_err; }
extern "C" int vs_version(void) { return 100; }
extern "C" int vs_set_pdl(int max_ctas) { const int old = g_vs_pdl; g_vs_pdl = max_ctas > 0 ? max_ctas : 0; return old; }
extern "C" int vs_has_tcgen05(void) {
#ifdef VS_WITH_TCGEN05
    return 1;
#else
    return 0;
#endif
}

extern "C" int vs_conv3x3x3_fprop(int in_dtype, int out_dtype, int in_planar, int out_planar, int flags, const void* x,
                                  const float* wpk, const void* wtc, const float* bias, void* y, double* stats,
                                  float* shift, int n, int d, int h, int w, int cin, int cout, void* stream) {
    const int prezeroed = (flags & VS_FLAG_PREZEROED) != 0;
#ifdef VS_WITH_TCGEN05
    if (wtc != nullptr && in_dtype == VS_BF16 && out_dtype == VS_BF16 && !in_planar && !out_planar && bias == nullptr &&
        tc_eligible(cin, cout)) {
        // the tensor-core kernel derives and publishes the shift itself (no separate launch)
        return vs_conv3x3x3_tc(x, wtc, y, stats, shift, prezeroed, nullptr, nullptr, nullptr, 0, nullptr, nullptr, n, d, h, w,
                               cin, cout, stream);
    }
#else
    (void)wtc;
#endif
    return vs_conv3x3x3_fprop_direct(in_dtype, out_dtype, in_planar, out_planar | (prezeroed ? 2 : 0), x, wpk, bias, y, stats,
                                     shift, n, d, h, w, cin, cout, stream);
}

extern "C" int vs_conv3x3x3_dgrad(int in_dtype, int out_dtype, int out_planar, const void* dy, const float* wdpk,
                                  const void* wdtc, void* dx, const void* y_prev, const double* stats_prev,
                                  double* sums_prev, int n, int d, int h, int w, int cin, int cout, void* stream) {
    // dx[.., cin] = conv3(dy[.., cout], wd[27][cout][cin]): the fprop contraction with channels swapped
#ifdef VS_WITH_TCGEN05
    if (wdtc != nullptr && in_dtype == VS_BF16 && out_dtype == VS_BF16 && !out_planar && tc_eligible(cout, cin))
        return vs_conv3x3x3_tc(dy, wdtc, dx, nullptr, nullptr, 0, y_prev, stats_prev, sums_prev, 0, nullptr, nullptr, n, d, h, w,
                               cout, cin, stream);
    // 2-channel planar fp32 input gradient (the VAE in-block, whose input is the module's NCDHW tensor): the pack is
    // built with the input channels zero-padded to 8 (vs_pack_conv3_weight_tc_padded) and the epilogue stores
    // channels 0..1 as planar fp32
    if (wdtc != nullptr && in_dtype == VS_BF16 && out_planar && cin == 2 && sums_prev == nullptr && tc_eligible(cout, 8))
        return vs_conv3x3x3_tc(dy, wdtc, nullptr, nullptr, nullptr, 0, nullptr, nullptr, nullptr, 2, (float*)dx, nullptr, n, d, h,
                               w, cout, 8, stream);
#else
    (void)wdtc;
#endif
    VS_REQUIRE(sums_prev == nullptr, VS_ERR_UNSUPPORTED,
               "conv3 dgrad: the fused norm-backward reduction exists on the tensor-core path only (bf16 NDHWC, Cout in "
               "{8, 16k}, Cin %% 8 == 0); call vs_inorm_relu_bwd_reduce instead");
    return vs_conv3x3x3_fprop_direct(in_dtype, out_dtype, 0, out_planar, dy, wdpk, nullptr, dx, nullptr, nullptr, n, d, h, w,
                                     cout, cin, stream);
}

// 2-class head: probs[N][2][D][H][W] = softmax(conv3(x, w) + bias) in one launch (tensor-core kernel with the output
// channels padded to 8; bias + softmax + planar fp32 store fused into the epilogue).  joint_model.py:224-225,366-367
extern "C" int vs_head_conv_softmax2_fwd(const void* x, const void* wtc8, const float* bias, float* probs, int n, int d,
                                         int h, int w, int cin, void* stream) {
#ifdef VS_WITH_TCGEN05
    VS_REQUIRE(x && wtc8 && probs, VS_ERR_SHAPE, "head_conv_softmax2_fwd: null pointer");
    VS_REQUIRE(tc_eligible(cin, 8), VS_ERR_UNSUPPORTED, "head_conv_softmax2_fwd: Cin must be 8 or a multiple of 16 (got %d)", cin);
    return vs_conv3x3x3_tc(x, wtc8, nullptr, nullptr, nullptr, 0, nullptr, nullptr, nullptr, 1, probs, bias, n, d, h, w, cin, 8,
                           stream);
#else
    VS_FAIL(VS_ERR_UNSUPPORTED, "library built without the tcgen05 kernels");
#endif
}

#ifndef VS_WITH_TCGEN05
extern "C" size_t vs_conv3_tc_kdn_pack_bytes(int, int, int) { return 0; }
extern "C" int vs_pack_conv3_weight_tc_kdn(const float*, void*, int, int, int, void*) {
    VS_FAIL(VS_ERR_UNSUPPORTED, "library built without the tcgen05 kernels");
}
extern "C" int vs_conv3x3x3_tc_kdn(const void*, const void*, void*, double*, float*, int, int, int, int, int, int, int, void*) {
    VS_FAIL(VS_ERR_UNSUPPORTED, "library built without the tcgen05 kernels");
}
extern "C" int vs_conv3x3x3_tc_in_relu(const void*, const void*, void*, void*, const void*, double*, float*, unsigned*, int, int,
                                       int, int, int, int, void*) {
    VS_FAIL(VS_ERR_UNSUPPORTED, "library built without the tcgen05 kernels");
}
extern "C" int vs_conv3x3x3_tc_kdn_in_relu(const void*, const void*, void*, void*, const void*, double*, float*, unsigned*, int,
                                           int, int, int, int, int, void*) {
    VS_FAIL(VS_ERR_UNSUPPORTED, "library built without the tcgen05 kernels");
}
extern "C" int vs_pack_conv3_weight_tc_kdn_padded(const float*, void*, int, int, int, int, int, void*) {
    VS_FAIL(VS_ERR_UNSUPPORTED, "library built without the tcgen05 kernels");
}
extern "C" int vs_conv3x3x3_tc_kdn_planar(const void*, const void*, float*, const float*, int, int, int, int, int, int, void*) {
    VS_FAIL(VS_ERR_UNSUPPORTED, "library built without the tcgen05 kernels");
}
extern "C" int vs_pack_conv3_weight_tc_padded(const float*, void*, int, int, int, int, int, void*) {
    VS_FAIL(VS_ERR_UNSUPPORTED, "library built without the tcgen05 kernels");
}
extern "C" size_t vs_conv3_tc_pack_bytes(int, int, int) { return 0; }
extern "C" int vs_pack_conv3_weight_tc(const float*, void*, int, int, int, void*) {
    VS_FAIL(VS_ERR_UNSUPPORTED, "library built without the tcgen05 kernels");
}
#endif
