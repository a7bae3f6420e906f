#!/usr/bin/env python
"""k2s2 (Conv3d / ConvTranspose3d k2 s2) kernels at the joint step's shapes: tcgen05 (csrc/k2s2_tc.cu, with and without
the swizzled operand rows / TMA-store epilogue) vs CUDA-core (csrc/k2s2.cu).  Each timing = CUDA events around 8
back-to-back launches over 8 DIFFERENT input/output sets (> L2 in total at the big shapes), queued behind a spin kernel so
the host's launch latency is not charged.  Diagnostic (numbers go to profiles/ by hand)."""
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from vae_segmentation_b200 import _cabi, ops  # noqa: E402

dev = "cuda"
PEAK = 6546.6
NSET = 8
handle = ctypes.CDLL(_cabi.LIB_PATH)


def timeit(fns):
    for f in fns[:2]:
        f()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(3):
        torch.cuda._sleep(int(2e-3 * 1.9e9))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for f in fns:
            f()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / len(fns) * 1e3)
    return best


P = int(sys.argv[1]) if len(sys.argv) > 1 else 96
N = 2
print("%-8s %6s %4s %9s %9s %9s %9s %8s   (wgrad: cuda = CUDA-core kernel, tc0 = tc = tensor-core kernel, no bias sums)" % ("op", "coarse", "C", "cuda us", "tc0 us", "tc us", "GB/s", "frac"))
shapes = []
side, c = P // 2, 8
while side >= 3 and c <= 256:                      # Down path: C at 2*side -> side
    shapes.append((side, c))
    side //= 2
    c *= 2
side, c = P // 2, 16
while side >= 3 and c <= 256:                      # Up path: C at side -> 2*side
    shapes.append((side, c))
    side //= 2
    c *= 2
for side, c in sorted(set(shapes), key=lambda t: (-t[0], t[1])):
    w = torch.randn(c, c, 2, 2, 2, device=dev) * 0.1
    b = torch.randn(c, device=dev)
    nset = NSET if side >= 24 else 2
    fines = [torch.randn(N, 2 * side, 2 * side, 2 * side, c, device=dev).bfloat16() for _ in range(nset)]
    coarses = [torch.randn(N, side, side, side, c, device=dev).bfloat16() for _ in range(nset)]
    pg = ops.pack_k2s2_weight_tc(w, c, c, False)
    ps = ops.pack_k2s2_weight_tc(w, c, c, True)
    dims = (N, side, side, side)
    byts = (fines[0].numel() + coarses[0].numel()) * 2
    for name in ("gather", "scatter", "wgrad"):
        if name == "gather":
            cuda = [lambda f=f: ops.k2s2_gather(f, w, b, dims, c, c) for f in fines]
            tc = [lambda f=f: ops.k2s2_gather(f, w, b, dims, c, c, wtc=pg) for f in fines]
        elif name == "scatter":
            cuda = [lambda x=x: ops.k2s2_scatter(x, w, b, dims, c, c) for x in coarses]
            tc = [lambda x=x: ops.k2s2_scatter(x, w, b, dims, c, c, wtc=ps) for x in coarses]
        else:
            dws = [torch.zeros(c, c, 2, 2, 2, device=dev) for _ in coarses]
            cuda = tc = [lambda x=x, f=f, d=d: ops.k2s2_wgrad(x, f, dims, c, c, dwt=d, accumulate=True) for x, f, d in zip(coarses, fines, dws)]
            handle.vs_debug_set_k2_wgrad_tc(0)
        t_cuda = timeit(cuda)
        handle.vs_debug_set_k2_wgrad_tc(1)
        handle.vs_debug_set_k2_tc(0)
        t_tc0 = timeit(tc)
        handle.vs_debug_set_k2_tc(1)
        t_tc = timeit(tc)
        print("%-8s %6d %4d %9.1f %9.1f %9.1f %9.0f %8.3f" % (name, side, c, t_cuda, t_tc0, t_tc, byts / t_tc / 1e3, byts / t_tc / 1e3 / PEAK), flush=True)
