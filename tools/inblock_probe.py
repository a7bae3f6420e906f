#!/usr/bin/env python
"""Times the in-block (planar fp32 -> NDHWC bf16) and head convolutions at 2 x 96^3 (CUDA events, L2 flushed)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from tools.kbench import timeit  # noqa: E402
from vae_segmentation_b200 import ops  # noqa: E402

B, S = 2, int(sys.argv[1]) if len(sys.argv) > 1 else 96
dims = (B, S, S, S)
for cin in (1, 2):
    x = torch.randn(B, cin, S, S, S, device="cuda")
    w = torch.randn(8, cin, 3, 3, 3, device="cuda") * 0.2
    wf, _ = ops.pack_conv3_weight(w)
    us = timeit(lambda: ops.conv3_fprop(x, wf, None, dims, cin, 8, torch.bfloat16, in_planar=True), 10)
    print("in_block %d->8 (shift + conv + stats): %.1f us" % (cin, us))
a = torch.randn(B, S, S, S, 8, device="cuda").bfloat16()
w = torch.randn(2, 8, 3, 3, 3, device="cuda") * 0.2
b = torch.randn(2, device="cuda")
wf, _ = ops.pack_conv3_weight(w)
us = timeit(lambda: ops.conv3_fprop(a, wf, b, dims, 8, 2, torch.float32, want_stats=False), 10)
print("head 8->2 direct: %.1f us" % us)
