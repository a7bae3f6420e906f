"""fp32 check mode: per-parameter gradient error vs the float64 oracle, next to the fp32 oracle's own deviation."""
import sys, os, re
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import ref_torch as R
from vae_segmentation_b200 import joint_model as jm, evaluation as ev
from vae_segmentation_b200.synthetic import synth_image, synth_label

def rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()

torch.manual_seed(11)
sd = R.init_seg_state()
img, label = synth_image(2, 32), synth_label(2, 32)
_, g32, _ = R.seg_train_step(sd, img, label, eps=0.0001)
_, g64, _ = R.seg_train_step(sd, img, label, eps=0.0001, dtype=torch.float64)
for run in range(1):
    seg = jm.Segmentation(1, 2, norm_type=1); seg.load_state_dict(sd); seg = seg.cuda().set_precision("fp32")
    b = seg({"img": img.cuda()}, "img", "pred")
    b["onehot"] = ev.one_hot(label.cuda(), 2)
    loss = 1 - ev.avg_dsc(b, source_key="pred", target_key="onehot", botindex=1, topindex=2, eps=0.0001)
    loss.backward()
    rows = []
    for k, p in seg.named_parameters():
        if p.grad is None or re.search(r'(in_block\.conv\.0|conv\.1\.conv\.[036])\.bias$', k): continue
        rows.append((rel(p.grad, g64[k]), rel(g32[k], g64[k]), k, g64[k].norm().item()))
    rows.sort(reverse=True)
    print("run", run)
    for e, r, k, nrm in rows:
        a, b = dict(seg.named_parameters())[k].grad.double().cpu().reshape(-1), g64[k].reshape(-1)
        print("  %-28s ours %.3e  ref32 %.3e  |g64| %.3e  scale %.6f" % (k, e, r, nrm, (a @ b / (b @ b)).item()))
    print("pred err", (b["pred"].detach().cpu() - R.seg_forward(sd, img)).abs().max().item()) if False else None
