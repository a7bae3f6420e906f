// C-ABI glue: error string, device query, and the dispatch between the tcgen05 and the
// CUDA-core convolution paths.  No CPU fallback exists anywhere in this library.
#include <stdarg.h>
#include <string.h>
#include "vs_common.cuh"

static thread_local char g_err[512] = "";

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

extern "C" const char* vs_last_error_string(void) { return g_err; }
extern "C" int vs_version(void) { return 100; }
extern "C" int vs_has_tcgen05(void) {
#ifdef VS_WITH_TCGEN05
    return 1;
#else
    return 0;
#endif
}

extern "C" int vs_conv3x3x3_fprop(int in_dtype, int out_dtype, int in_planar, int out_planar, const void* x,
                                  const float* wpk, const float* bias, void* y, double* stats, float* shift, int n,
                                  int d, int h, int w, int cin, int cout, void* stream) {
    return vs_conv3x3x3_fprop_direct(in_dtype, out_dtype, in_planar, out_planar, x, wpk, bias, y, stats, shift, n, d, h,
                                     w, cin, cout, stream);
}

extern "C" int vs_conv3x3x3_dgrad(int in_dtype, int out_dtype, int out_planar, const void* dy, const float* wdpk,
                                  void* dx, int n, int d, int h, int w, int cin, int cout, void* stream) {
    // dx[.., cin] = conv3(dy[.., cout], wd[27][cout][cin]): the fprop contraction with channels swapped
    return vs_conv3x3x3_fprop_direct(in_dtype, out_dtype, 0, out_planar, dy, wdpk, nullptr, dx, nullptr, nullptr, n, d, h, w,
                                     cout, cin, stream);
}
