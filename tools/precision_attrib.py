#!/usr/bin/env python
"""Which bf16 storage point costs the gradient accuracy?  Runs the fp32 check mode on conditioned weights while
rounding selected tensor classes through bf16 (engine.SIMULATE_BF16: y = raw conv outputs, a = activations,
g = gradients w.r.t. activations, dy = gradients w.r.t. raw conv outputs, k2 = k2s2 outputs) and prints the
per-parameter gradient error against the fp32 oracle.  Diagnostic only."""
import os
import sys
from collections import OrderedDict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from oracle import conditioned as C  # noqa: E402
from oracle import ref_torch as R  # noqa: E402
from tools.precision_cond import grad_report, rel  # noqa: E402
from vae_segmentation_b200 import engine, evaluation as ev, joint_model as jm  # noqa: E402

P = int(sys.argv[1]) if len(sys.argv) > 1 else 64
seg_sd, _ = C.train_seg(60, patch=32, lr=0.1)
torch.manual_seed(1234)
img, label = C.blob_batch(1, P)
loss_ref, g_ref, pred_ref = R.seg_train_step(seg_sd, img, label, eps=0.0001)
for sim in ([], ["y"], ["a"], ["g"], ["dy"], ["k2"], ["y", "a", "dy", "k2"], ["y", "a", "g", "dy", "k2"]):
    engine.SIMULATE_BF16 = set(sim)
    seg = jm.Segmentation(1, 2, norm_type=1)
    seg.load_state_dict(seg_sd, strict=True)
    seg = seg.to("cuda").set_precision("fp32")
    pred = seg.predict(img.to("cuda"))
    loss = 1 - ev.avg_dsc_fused(pred, label.to("cuda"), "label", botindex=1, topindex=2, eps=0.0001)
    loss.backward()
    got = OrderedDict((k, p.grad.detach().cpu().clone()) for k, p in seg.named_parameters() if p.grad is not None)
    print("== simulate bf16 on %s: probs rel-L2 %.3e" % (sim, rel(pred, pred_ref)))
    grad_report("   ", got, g_ref)
engine.SIMULATE_BF16 = set()

# ---- VAE train step (main_source.py:389-406): no skip connections, every gradient passes through the 2^3 .. 4^3 levels
from vae_segmentation_b200.synthetic import synth_label  # noqa: E402
vae_sd, _ = C.train_vae(60, patch=P, lr=0.1)
torch.manual_seed(1235)
label = synth_label(2, P)
z = torch.randn(2, 128)
loss_ref, dsc_ref, kl_ref, g_ref, rec_ref = R.vae_train_step(vae_sd, label, scale=0.35, z=z, eps=0.0001)
for sim in ([], ["y"], ["a"], ["g"], ["dy"], ["k2"], ["y", "a", "g", "dy", "k2"]):
    engine.SIMULATE_BF16 = set(sim)
    vae = jm.VAE(2, 2, norm_type=1, dim=128, patch=P)
    vae.load_state_dict(vae_sd, strict=True)
    vae = vae.to("cuda").set_precision("fp32")
    oh = ev.one_hot(label.to("cuda"), 2)
    recon, mean, std = vae(oh, if_random=True, scale=0.35, z=z)
    d = {"recon": recon, "onehot": oh, "mean": mean, "std": std}
    (1 - ev.avg_dsc(d, source_key="recon", target_key="onehot", botindex=1, topindex=2, eps=0.0001) + 0.00002 * ev.KLloss(d)).backward()
    got = OrderedDict((k, p.grad.detach().cpu().clone()) for k, p in vae.named_parameters() if p.grad is not None)
    print("== VAE, simulate bf16 on %s: recon rel-L2 %.3e" % (sim, rel(recon, rec_ref)))
    grad_report("   ", got, g_ref)
engine.SIMULATE_BF16 = set()
