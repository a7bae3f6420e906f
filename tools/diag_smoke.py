#!/usr/bin/env python
"""bf16 Seg forward at 32^3 / 64^3 against the CPU oracle: rel-L2, max-abs and argmax agreement of the probabilities,
for several seeds (what __graft_entry__.smoke() and tests/test_models_gpu.py bound)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from oracle import ref_torch as R  # noqa: E402
from vae_segmentation_b200 import _cabi  # noqa: E402
from vae_segmentation_b200 import joint_model as jm  # noqa: E402
from vae_segmentation_b200.synthetic import synth_image  # noqa: E402

dev = "cuda:0"
for patch in (32, 64):
    for seed in (0, 1, 2):
        torch.manual_seed(seed)
        sd = R.init_seg_state()
        img = synth_image(1, patch)
        ref = R.seg_forward(sd, img) if hasattr(R, "seg_forward") else None
        if ref is None:
            _, _, ref = R.seg_train_step(sd, img, (torch.rand(1, 1, patch, patch, patch) > 0.97).float())
        for inblk in (1, 0):
            _cabi.lib().vs_debug_set_inblock_kernel(inblk)
            seg = jm.Segmentation(1, 2, norm_type=1)
            seg.load_state_dict(sd)
            seg = seg.to(dev).set_precision("bf16")
            with torch.no_grad():
                pred = seg.predict(img.to(dev)).cpu()
            d = pred - ref
            print("P=%d seed=%d inblock=%d head_direct=%s: rel-L2 %.3e  max-abs %.3e  argmax agree %.4f" % (
                patch, seed, inblk, os.environ.get("VAESEG_HEAD_DIRECT", "0"), (d.norm() / ref.norm()).item(), d.abs().max().item(),
                (pred.argmax(1) == ref.argmax(1)).float().mean().item()), flush=True)
_cabi.lib().vs_debug_set_inblock_kernel(1)
