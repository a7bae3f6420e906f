#!/usr/bin/env python
"""Whole-gradient cosine / rel-L2 of the bf16 path against the fp32 oracle at RANDOM init, InstanceNorm vs BatchNorm models
(same seeds, 2 x 32^3): is the BatchNorm path any less precise than the InstanceNorm one?  Diagnostic."""
import os
import sys
from collections import OrderedDict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from oracle import ref_torch as R  # noqa: E402
from vae_segmentation_b200 import evaluation as ev  # noqa: E402
from vae_segmentation_b200 import joint_model as jm  # noqa: E402

for norm in (1, 2):
    for seed in (31, 32, 33):
        torch.manual_seed(seed)
        sd = OrderedDict((k, v.clone()) for k, v in jm.Segmentation(1, 2, norm_type=norm).state_dict().items())
        img, label = torch.randn(2, 1, 32, 32, 32), (torch.rand(2, 1, 32, 32, 32) > 0.7).float()
        rsd = R._leafify(sd)
        pred_ref = R.seg_forward(rsd, img)
        (1 - R.avg_dsc(pred_ref, R.one_hot(label), botindex=1, topindex=2, eps=0.0001)).backward()
        for prec in ("fp32", "bf16"):
            seg = jm.Segmentation(1, 2, norm_type=norm)
            seg.load_state_dict(sd)
            seg.cuda().set_precision(prec)
            pred = seg({"img": img.cuda()}, "img", "pred")["pred"]
            loss = 1 - ev.avg_dsc({"p": pred, "t": ev.one_hot(label.cuda(), 2)}, "p", "t", botindex=1, topindex=2, eps=0.0001)
            loss.backward()
            keys = [k for k, p in seg.named_parameters() if p.grad is not None and rsd[k].grad is not None]
            a = torch.cat([dict(seg.named_parameters())[k].grad.reshape(-1).double().cpu() for k in keys])
            b = torch.cat([rsd[k].grad.reshape(-1).double() for k in keys])
            print("norm_type %d seed %d %s: probs rel-L2 %.3e  grad cosine %.4f  rel-L2 %.3e" % (
                norm, seed, prec, ((pred.cpu() - pred_ref.detach()).norm() / pred_ref.norm()).item(),
                (a @ b / (a.norm() * b.norm())).item(), ((a - b).norm() / b.norm()).item()))
