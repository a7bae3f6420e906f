#!/usr/bin/env python
"""Full-resolution convolution kernels at the joint step's shapes (2 x 96^3 / 2 x 48^3): kd-in-N fprop / dgrad, the
tap-per-MMA kernel, the weight gradients, the InstanceNorm passes.  CUDA events around 6 back-to-back launches over
rotating buffers (> L2 in total), queued behind a spin kernel; best of 3.  Diagnostic (numbers go to profiles/ by hand)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from vae_segmentation_b200 import ops  # noqa: E402

dev = "cuda"
PEAK = 6546.6
NSET = 6


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


print("%-34s %9s %9s %7s" % ("kernel", "us", "GB/s", "frac"))
for (s, cin, cout) in [(96, 8, 8), (96, 16, 8), (48, 16, 16), (48, 32, 16), (48, 8, 16)]:
    n = 2
    dims = (n, s, s, s)
    xs = [torch.randn(n, s, s, s, cin, device=dev).bfloat16() for _ in range(NSET)]
    dys = [torch.randn(n, s, s, s, cout, device=dev).bfloat16() for _ in range(NSET)]
    w = torch.randn(cout, cin, 3, 3, 3, device=dev) * 0.1
    wf, wdg = ops.pack_conv3_weight(w)
    wtc, wdtc = ops.pack_conv3_weight_tc(w, dgrad=False), ops.pack_conv3_weight_tc(w, dgrad=True)
    wk, wkd = ops.pack_conv3_weight_tc_kdn(w, dgrad=False), ops.pack_conv3_weight_tc_kdn(w, dgrad=True)
    vox = n * s ** 3
    byts = vox * (cin + cout) * 2
    rows = []
    rows.append(("kdn fprop", [lambda x=x: ops.conv3_tc_kdn(x, wk, dims, cin, cout, want_stats=True) for x in xs]))
    rows.append(("tap-per-MMA fprop", [lambda x=x: ops.conv3_fprop(x, wf, None, dims, cin, cout, torch.bfloat16, wtc=wtc) for x in xs]))
    if wkd is not None:
        ys = [torch.randn(n, s, s, s, cin, device=dev).bfloat16() for _ in range(NSET)]
        st = torch.stack([torch.zeros(n, cin, dtype=torch.float64), torch.full((n, cin), float(s ** 3), dtype=torch.float64)], -1).to(dev)
        sums = torch.zeros(n, cin, 2, device=dev, dtype=torch.float64)
        rows.append(("kdn dgrad + fused reduce", [lambda d=d, y=y: ops.conv3_tc_kdn(d, wkd, dims, cout, cin, prev=(y, st, sums)) for d, y in zip(dys, ys)]))
        rows.append(("tap-per-MMA dgrad + fused reduce", [lambda d=d, y=y: ops.conv3_dgrad(d, wdg, dims, cin, cout, torch.bfloat16, wdtc=wdtc, prev=(y, st, sums)) for d, y in zip(dys, ys)]))
    rows.append(("wgrad", [lambda x=x, d=d: ops.conv3_wgrad(x, d, dims, cin, cout) for x, d in zip(xs, dys)]))
    stt = torch.stack([torch.zeros(n, cout, dtype=torch.float64), torch.full((n, cout), float(s ** 3), dtype=torch.float64)], -1).to(dev)
    rows.append(("inorm_relu_apply (cout)", [lambda d=d: ops.inorm_relu_apply(d, stt) for d in dys]))
    for name, fns in rows:
        t = timeit(fns)
        b = byts if "inorm" not in name else vox * cout * 4
        print("%3d^3 %2d->%2d %-22s %9.1f %9.0f %7.3f" % (s, cin, cout, name[:22], t, b / t / 1e3, b / t / 1e3 / PEAK), flush=True)
