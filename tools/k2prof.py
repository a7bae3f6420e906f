#!/usr/bin/env python
"""Launches the k2s2 tensor-core kernels at the 48^3 <-> 96^3, 16-channel shape (for ncu): scatter with direct stores,
scatter with TMA stores, gather with swizzled rows, gather with 8-channel planes."""
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from vae_segmentation_b200 import _cabi, ops  # noqa: E402

dev = "cuda"
handle = ctypes.CDLL(_cabi.LIB_PATH)
side, c, N = 48, 16, 2
w = torch.randn(c, c, 2, 2, 2, device=dev) * 0.1
b = torch.randn(c, device=dev)
fine = torch.randn(N, 2 * side, 2 * side, 2 * side, c, device=dev).bfloat16()
coarse = torch.randn(N, side, side, side, c, device=dev).bfloat16()
pg = ops.pack_k2s2_weight_tc(w, c, c, False)
ps = ops.pack_k2s2_weight_tc(w, c, c, True)
dims = (N, side, side, side)
for variant in (0, 1):
    handle.vs_debug_set_k2_tc(variant)
    ops.k2s2_scatter(coarse, w, b, dims, c, c, wtc=ps)
    ops.k2s2_gather(fine, w, b, dims, c, c, wtc=pg)
torch.cuda.synchronize()
