#!/usr/bin/env python
"""Runs one 3x3x3 wgrad shape a few times (ncu target).  python tools/wgrad_probe.py S CIN COUT [iters]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from vae_segmentation_b200 import ops  # noqa: E402

s, cin, cout = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
iters = int(sys.argv[4]) if len(sys.argv) > 4 else 3
B = 2
torch.manual_seed(0)
x = torch.randn(B, s, s, s, cin, device="cuda").bfloat16()
gy = torch.randn(B, s, s, s, cout, device="cuda").bfloat16()
dw = torch.zeros(cout, cin, 3, 3, 3, device="cuda")
for _ in range(iters):
    ops.conv3_wgrad(x, gy, (B, s, s, s), cin, cout, dw=dw)
torch.cuda.synchronize()
print("ok", dw.abs().mean().item())
