#!/usr/bin/env python
"""k2s2 (Conv3d / ConvTranspose3d k2 s2) kernels at the joint step's shapes: tcgen05 (csrc/k2s2_tc.cu) vs CUDA-core
(csrc/k2s2.cu), CUDA events, L2 flushed between iterations.  Diagnostic (numbers go to profiles/ by hand)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from vae_segmentation_b200 import ops  # noqa: E402

dev = "cuda"
PEAK = 6546.6
flush = torch.empty(256 << 20, device=dev, dtype=torch.uint8)


def timeit(fn, iters=10):
    for _ in range(3):
        fn()
    tot = 0.0
    for _ in range(iters):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    return tot / iters * 1e3


P = int(sys.argv[1]) if len(sys.argv) > 1 else 96
N = 2
print("%-8s %5s %4s %10s %10s %8s %8s" % ("op", "coarse", "C", "cuda us", "tc us", "tc GB/s", "frac"))
side, c = P // 2, 8
while side >= 3 and c <= 256:
    if (2 * side) % 2 == 0:
        w = torch.randn(c, c, 2, 2, 2, device=dev) * 0.1
        b = torch.randn(c, device=dev)
        fine = torch.randn(N, 2 * side, 2 * side, 2 * side, c, device=dev).bfloat16()
        coarse = torch.randn(N, side, side, side, c, device=dev).bfloat16()
        pg = ops.pack_k2s2_weight_tc(w, c, c, False)
        ps = ops.pack_k2s2_weight_tc(w, c, c, True)
        dims = (N, side, side, side)
        byts = (fine.numel() + coarse.numel()) * 2
        for name, f_cuda, f_tc in (
                ("gather", lambda: ops.k2s2_gather(fine, w, b, dims, c, c), lambda: ops.k2s2_gather(fine, w, b, dims, c, c, wtc=pg)),
                ("scatter", lambda: ops.k2s2_scatter(coarse, w, b, dims, c, c), lambda: ops.k2s2_scatter(coarse, w, b, dims, c, c, wtc=ps))):
            t0, t1 = timeit(f_cuda), timeit(f_tc)
            print("%-8s %5d %4d %10.1f %10.1f %8.0f %8.3f" % (name, side, c, t0, t1, byts / t1 / 1e3, byts / t1 / 1e3 / PEAK), flush=True)
    side //= 2
    c *= 2
