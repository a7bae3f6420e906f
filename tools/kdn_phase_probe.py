#!/usr/bin/env python
"""Steady-state tile timeline of the kd-in-N convolution (clock64 stamps of CTA 0 around its 9th tile)."""
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from vae_segmentation_b200 import _cabi, ops  # noqa: E402

dev = "cuda"
h = ctypes.CDLL(_cabi.LIB_PATH)
h.vs_debug_set_kdn_phase_buffer.argtypes = [ctypes.c_void_p]
NAMES = ["TMA issued (tile t)", "TMA issued (t+1)", "MMA: at tempty wait", "MMA: tempty ok", "MMA: data landed", "MMA: issued+committed",
         "EPI: at tfull wait", "EPI: accumulators done", "EPI: stored", "EPI: zeroed+released", "MMA(t+1): tempty ok", "EPI: loop top (next tile of group)", "EPI: loop top (tile t)"]
for (s, cin, cout) in [(96, 8, 8), (96, 16, 8)]:
    n = 2
    x = torch.randn(n, s, s, s, cin, device=dev).bfloat16()
    w = torch.randn(cout, cin, 3, 3, 3, device=dev) * 0.1
    wk = ops.pack_conv3_weight_tc_kdn(w, dgrad=False)
    dbg = torch.zeros(16, device=dev, dtype=torch.int64)
    for it in range(3):
        if it == 2:
            h.vs_debug_set_kdn_phase_buffer(ctypes.c_void_p(dbg.data_ptr()))
        ops.conv3_tc_kdn(x, wk, (n, s, s, s), cin, cout, want_stats=True)
        torch.cuda.synchronize()
    h.vs_debug_set_kdn_phase_buffer(None)
    t = dbg.cpu().tolist()
    base = t[2]
    print("%d x %d^3 %d->%d | " % (n, s, cin, cout) + "  ".join("%s %+d" % (NAMES[i], t[i] - base) for i in range(13)), flush=True)
