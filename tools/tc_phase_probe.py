#!/usr/bin/env python
"""Where does a deep-level tensor-core convolution launch spend its time?  clock64 stamps of CTA 0 at the phases of its
first work item (csrc/conv3_tc.cu: TC_DBG) for the joint step's deep layer shapes.  Diagnostic only."""
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from vae_segmentation_b200 import _cabi, ops  # noqa: E402

dev = "cuda"
h = ctypes.CDLL(_cabi.LIB_PATH)
h.vs_debug_set_tc_phase_buffer.argtypes = [ctypes.c_void_p]
NAMES = ["entry", "setup done", "1st stage issued", "last stage issued", "1st stage landed", "MMAs issued",
         "epilogue ready", "accumulators done", "tile stored", "stats flushed", "exit"]
for (n, s, cin, cout) in [(2, 24, 32, 32), (2, 12, 64, 64), (2, 6, 128, 128), (2, 12, 128, 64), (2, 48, 16, 16)]:
    for ks in (1, 0):
        h.vs_debug_set_conv3_ksplit(ks)
        x = torch.randn(n, s, s, s, cin, device=dev).bfloat16()
        w = torch.randn(cout, cin, 3, 3, 3, device=dev) * 0.05
        wf, _ = ops.pack_conv3_weight(w)
        wtc = ops.pack_conv3_weight_tc(w, dgrad=False)
        dbg = torch.zeros(16, device=dev, dtype=torch.int64)
        for it in range(3):
            if it == 2:
                h.vs_debug_set_tc_phase_buffer(ctypes.c_void_p(dbg.data_ptr()))
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda._sleep(200000)
            e0.record()
            ops.conv3_fprop(x, wf, None, (n, s, s, s), cin, cout, torch.bfloat16, wtc=wtc)
            e1.record()
            torch.cuda.synchronize()
        h.vs_debug_set_tc_phase_buffer(None)
        t = dbg.cpu().tolist()
        base = t[0]
        print("%d x %d^3 %d->%d ksplit=%d  event %.1f us | " % (n, s, cin, cout, ks, e0.elapsed_time(e1) * 1e3) +
              "  ".join("%s %+d" % (NAMES[i], t[i] - base) for i in range(1, 11)), flush=True)
h.vs_debug_set_conv3_ksplit(1)
