"""fp32 check mode, joint step: per-parameter gradient error vs float64 oracle (ours and the fp32 oracle)."""
import sys, os, re
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from collections import OrderedDict
from oracle import ref_torch as R
from vae_segmentation_b200 import joint_model as jm, train_step as ts
from vae_segmentation_b200.synthetic import synth_image, synth_label

def rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()

patch = 64
for seed in (31, 32):
    torch.manual_seed(seed)
    seg_sd = R.init_seg_state(); vae_sd = R.init_vae_state(2, 128, patch)
    teacher_sd = OrderedDict((k, v + 0.01 * torch.randn_like(v)) for k, v in seg_sd.items())
    img, label = synth_image(1, patch), synth_label(1, patch)
    _, g32 = R.joint_target_step(seg_sd, vae_sd, teacher_sd, img, label, lambda_vae=1.0, loss_type=0, kl=False)
    _, g64 = R.joint_target_step(seg_sd, vae_sd, teacher_sd, img, label, lambda_vae=1.0, loss_type=0, kl=False, dtype=torch.float64)
    def mk(sd):
        s = jm.Segmentation(1, 2, norm_type=1); s.load_state_dict(sd)
        v = jm.VAE(2, 2, norm_type=1, dim=128, patch=patch); v.load_state_dict(vae_sd)
        return jm.Joint([s.cuda().set_precision("fp32"), v.cuda().set_precision("fp32")])
    tr = ts.JointTrainer(mk(seg_sd), mk(teacher_sd), lambda_vae=1.0, loss_type=0)
    tr.step(img.cuda(), label.cuda())
    ours = {k: p.grad for k, p in tr.student.Seg.named_parameters() if p.grad is not None}
    keys = [k for k in ours if not re.search(r'(in_block\.conv\.0|conv\.1\.conv\.[036])\.bias$', k)]
    eo = sorted((rel(ours[k], g64[k]), rel(g32[k], g64[k]), k) for k in keys)
    tot_o = rel(torch.cat([ours[k].reshape(-1) for k in keys]), torch.cat([g64[k].reshape(-1) for k in keys]))
    tot_r = rel(torch.cat([g32[k].reshape(-1) for k in keys]), torch.cat([g64[k].reshape(-1) for k in keys]))
    print("seed %d: all-params rel-L2 ours %.3e ref32 %.3e; median ours %.3e ref %.3e" % (seed, tot_o, tot_r, eo[len(eo)//2][0], sorted(e[1] for e in eo)[len(eo)//2]))
    for e in eo[-6:]:
        print("   %-30s ours %.3e ref32 %.3e" % (e[2], e[0], e[1]))
