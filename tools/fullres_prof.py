#!/usr/bin/env python
"""Launches the full-resolution 8-channel convolution kernels at 2 x 96^3 (for ncu): kd-in-N fprop, kd-in-N dgrad with the
fused norm-backward reduction, the 2-class head (kd-in-N, planar softmax epilogue), tap-per-MMA dgrad (round-2 baseline),
d-shift wgrad."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from vae_segmentation_b200 import ops  # noqa: E402

dev = "cuda"
n, s, c = 2, 96, 8
dims = (n, s, s, s)
x = torch.randn(n, s, s, s, c, device=dev).bfloat16()
dy = torch.randn(n, s, s, s, c, device=dev).bfloat16()
w = torch.randn(c, c, 3, 3, 3, device=dev) * 0.1
wk = ops.pack_conv3_weight_tc_kdn(w, dgrad=False)
wdtc = ops.pack_conv3_weight_tc(w, dgrad=True)
wkd = ops.pack_conv3_weight_tc_kdn(w, dgrad=True)
wh = torch.randn(2, c, 3, 3, 3, device=dev) * 0.1
wk8 = ops.pack_conv3_weight_tc_kdn_padded(wh, c, 8, dgrad=False)
bias = torch.zeros(2, device=dev)
for _ in range(2):
    y, stats = ops.conv3_tc_kdn(x, wk, dims, c, c, want_stats=True)
    sums = torch.zeros(n, c, 2, device=dev, dtype=torch.float64)
    ops.conv3_tc_kdn(dy, wkd, dims, c, c, prev=(y, stats, sums))
    ops.conv3_tc_kdn_planar(x, wk8, dims, c, 1, bias=bias)
    sums = torch.zeros(n, c, 2, device=dev, dtype=torch.float64)
    ops.conv3_dgrad(dy, None, dims, c, c, torch.bfloat16, wdtc=wdtc, prev=(y, stats, sums))
    ops.conv3_wgrad(x, dy, dims, c, c)
torch.cuda.synchronize()
