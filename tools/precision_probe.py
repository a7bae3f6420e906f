"""Measures bf16-mode deviation from the fp32 check mode (itself pinned to the oracle at 1e-4)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vae_segmentation_b200 import joint_model as jm
from vae_segmentation_b200.synthetic import synth_image

for patch in (32, 64, 96):
    torch.manual_seed(0)
    seg = jm.Segmentation(1, 2, norm_type=1).cuda()
    img = synth_image(2, patch).cuda()
    with torch.no_grad():
        p32 = seg.set_precision("fp32").predict(img)
        p16 = seg.set_precision("bf16").predict(img)
    d = (p32 - p16).abs()
    agree = (p32.argmax(1) == p16.argmax(1)).float().mean().item()
    rel = d / p32.clamp_min(1e-6)
    print("P=%d max|dp|=%.4f mean|dp|=%.5f p99.9|dp|=%.4f relL2=%.4e maxrel=%.3f argmax=%.5f" % (
        patch, d.max().item(), d.mean().item(), d.flatten().float().kthvalue(int(d.numel() * 0.999)).values.item(),
        ((p32 - p16).norm() / p32.norm()).item(), rel.max().item(), agree))
