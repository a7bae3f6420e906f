#!/usr/bin/env python
"""Per-kernel micro-benchmark at the joint step's layer shapes (B=2, 96^3 patch by default).

    python tools/kbench.py [--set conv,wgrad,k2s2,norm] [--check] [--iters 10] [--patch 96] [--batch 2]

CUDA-event timing (3 warm-up + `iters` timed launches on the current stream, L2 flushed by a
256 MB write between launches), one line per (op, shape): us, algorithmic GB/s and TFLOP/s.
--check also compares the tensor-core convolution with the CUDA-core direct kernel on the same
bf16 operands.  This is a development tool (gpurun), not part of the product path.
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from vae_segmentation_b200 import ops  # noqa: E402

DEV = "cuda"
BF = torch.bfloat16
_flush = None


def flush_l2():
    global _flush
    if _flush is None:
        _flush = torch.empty(64 * 1024 * 1024, device=DEV, dtype=torch.float32)
    _flush.add_(1.0)


def timeit(fn, iters, flush=True):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    tot = 0.0
    for _ in range(iters):
        if flush:
            flush_l2()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        e1.synchronize()
        tot += e0.elapsed_time(e1)
    return tot / iters * 1e3          # us


def report(name, shape, us, flops, bytes_):
    print("%-14s %-28s %9.1f us %8.1f GB/s %8.2f TF/s" % (name, shape, us, bytes_ / us / 1e3, flops / us / 1e6), flush=True)


def seg_conv_shapes(P, B):
    """(d, cin, cout) of every distinct 3x3x3 conv of Seg + VAE that the tensor-core path takes."""
    out = []
    for lvl, c in enumerate([8, 16, 32, 64, 128, 256]):
        s = P >> lvl
        if s < 1:
            continue
        if lvl > 0:
            out.append((s, c // 2, c))          # down.c0
        out.append((s, c, c))                   # c3 / c6
        if lvl < 5:
            out.append((s, 2 * c, c))           # up.c0
    return out


ONLY = None


def bench_conv(P, B, iters, check):
    for s, cin, cout in seg_conv_shapes(P, B):
        if ONLY is not None and (s, cin, cout) != ONLY:
            continue
        dims = (B, s, s, s)
        vox = B * s ** 3
        x = torch.randn(B, s, s, s, cin, device=DEV).to(BF)
        w = torch.randn(cout, cin, 3, 3, 3, device=DEV) * 0.1
        wf, wd = ops.pack_conv3_weight(w)
        wtc = ops.pack_conv3_weight_tc(w, dgrad=False)
        wdtc = ops.pack_conv3_weight_tc(w, dgrad=True)
        fl = 2.0 * 27 * cin * cout * vox
        by = vox * (cin + cout) * 2
        us = timeit(lambda: ops.conv3_fprop(x, wf, None, dims, cin, cout, BF, wtc=wtc), iters)
        report("conv3_fprop", "%d^3 %d->%d" % (s, cin, cout), us, fl, by)
        gy = torch.randn(B, s, s, s, cout, device=DEV).to(BF)
        us = timeit(lambda: ops.conv3_dgrad(gy, wd, dims, cin, cout, BF, wdtc=wdtc), iters)
        report("conv3_dgrad", "%d^3 %d<-%d" % (s, cin, cout), us, fl, by)
        if check:
            y, st = ops.conv3_fprop(x, wf, None, dims, cin, cout, BF, shifted=False, wtc=wtc)
            wq = w.to(BF).float().reshape(cout, cin, 27).permute(2, 1, 0).contiguous()
            y2, st2 = ops.conv3_fprop(x, wq, None, dims, cin, cout, BF, shifted=False, wtc=None)
            err = (y.float() - y2.float()).abs().max().item() / y2.float().abs().max().item()
            serr = ((st - st2).abs().max() / st2.abs().max()).item()
            dx = ops.conv3_dgrad(gy, wd, dims, cin, cout, BF, wdtc=wdtc)
            wdq = wd.to(BF).float()
            dx2 = ops.conv3_dgrad(gy, wdq, dims, cin, cout, BF, wdtc=None)
            derr = (dx.float() - dx2.float()).abs().max().item() / dx2.float().abs().max().item()
            flag = "" if max(err, serr, derr) < 1e-2 else "   <-- MISMATCH"
            print("   check: fprop rel-max %.2e stats %.2e dgrad %.2e%s" % (err, serr, derr, flag), flush=True)


def bench_wgrad(P, B, iters, check):
    shapes = [(P, 1, 8), (P, 8, 2)] + seg_conv_shapes(P, B)
    for s, cin, cout in shapes:
        if s < 3:
            continue
        dims = (B, s, s, s)
        vox = B * s ** 3
        x = torch.randn(B, s, s, s, cin, device=DEV).to(BF) if cin >= 8 else torch.randn(B, cin, s, s, s, device=DEV)
        gy = torch.randn(B, s, s, s, cout, device=DEV).to(BF)
        dw = torch.zeros(cout, cin, 3, 3, 3, device=DEV)
        fl = 2.0 * 27 * cin * cout * vox
        by = vox * (cin + cout) * 2
        us = timeit(lambda: ops.conv3_wgrad(x, gy, dims, cin, cout, dw=dw, in_planar=cin < 8), iters)
        report("conv3_wgrad", "%d^3 %d,%d" % (s, cin, cout), us, fl, by)


def bench_k2s2(P, B, iters, check):
    for lvl, c in enumerate([8, 16, 32, 64, 128]):
        sc = P >> (lvl + 1)
        if sc < 1:
            continue
        cd = (B, sc, sc, sc)
        vox = B * sc ** 3
        fine = torch.randn(B, 2 * sc, 2 * sc, 2 * sc, c, device=DEV).to(BF)
        coarse = torch.randn(B, sc, sc, sc, c, device=DEV).to(BF)
        wt = torch.randn(c, c, 2, 2, 2, device=DEV) * 0.1
        bias = torch.randn(c, device=DEV)
        fl = 2.0 * 8 * c * c * vox
        by = vox * 9 * c * 2
        us = timeit(lambda: ops.k2s2_gather(fine, wt, bias, cd, c, c), iters)
        report("k2s2_gather", "%d^3->%d^3 C=%d" % (2 * sc, sc, c), us, fl, by)
        us = timeit(lambda: ops.k2s2_scatter(coarse, wt, bias, cd, c, c), iters)
        report("k2s2_scatter", "%d^3->%d^3 C=%d" % (sc, 2 * sc, c), us, fl, by)
        dwt = torch.zeros(c, c, 2, 2, 2, device=DEV)
        db = torch.zeros(c, device=DEV)
        us = timeit(lambda: ops.k2s2_wgrad(coarse, fine, cd, c, c, dwt=dwt, dbias_coarse=db), iters)
        report("k2s2_wgrad", "%d^3 C=%d" % (sc, c), us, fl, by)


def bench_norm(P, B, iters, check):
    for lvl, c in enumerate([8, 16, 32, 64, 128]):
        s = P >> lvl
        y = torch.randn(B, s, s, s, c, device=DEV).to(BF)
        g = torch.randn(B, s, s, s, c, device=DEV).to(BF)
        stats = torch.stack([y.float().sum((1, 2, 3)).double(), (y.float() ** 2).sum((1, 2, 3)).double()], -1).contiguous()
        by = y.numel() * 2
        us = timeit(lambda: ops.inorm_relu_apply(y, stats), iters)
        report("inorm_apply", "%d^3 C=%d" % (s, c), us, 0, 2 * by)
        us = timeit(lambda: ops.inorm_relu_bwd(g, y, stats), iters)
        report("inorm_bwd(2k)", "%d^3 C=%d" % (s, c), us, 0, 5 * by)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--set", default="conv,wgrad,k2s2,norm")
    ap.add_argument("--check", action="store_true")
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--patch", type=int, default=96)
    ap.add_argument("--batch", type=int, default=2)
    ap.add_argument("--k2-ksplit-max", type=int, default=None, help="override the k-split threshold of the k2s2 kernels")
    ap.add_argument("--only", default=None, help="conv set: 'size,cin,cout' of the single shape to run")
    a = ap.parse_args()
    global ONLY
    if a.only:
        ONLY = tuple(int(v) for v in a.only.split(","))
    if a.k2_ksplit_max is not None:
        from vae_segmentation_b200 import _cabi
        _cabi.lib().vs_debug_set_k2_ksplit_max(a.k2_ksplit_max)
    torch.manual_seed(0)
    fns = {"conv": bench_conv, "wgrad": bench_wgrad, "k2s2": bench_k2s2, "norm": bench_norm}
    for k in a.set.split(","):
        fns[k](a.patch, a.batch, a.iters, a.check)


if __name__ == "__main__":
    main()
