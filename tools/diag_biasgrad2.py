"""fp32 check mode: where (in backward order) does the gradient error vs the float64 oracle jump? several seeds."""
import sys, os, re
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import ref_torch as R
from vae_segmentation_b200 import joint_model as jm, evaluation as ev
from vae_segmentation_b200.synthetic import synth_image, synth_label

def rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()

for seed, patch in ((11, 32), (12, 32), (13, 32), (14, 48), (15, 64)):
    torch.manual_seed(seed)
    sd = R.init_seg_state()
    img, label = synth_image(2, patch), synth_label(2, patch)
    _, g32, _ = R.seg_train_step(sd, img, label, eps=0.0001)
    _, g64, _ = R.seg_train_step(sd, img, label, eps=0.0001, dtype=torch.float64)
    seg = jm.Segmentation(1, 2, norm_type=1); seg.load_state_dict(sd); seg = seg.cuda().set_precision("fp32")
    b = seg({"img": img.cuda()}, "img", "pred")
    b["onehot"] = ev.one_hot(label.cuda(), 2)
    loss = 1 - ev.avg_dsc(b, source_key="pred", target_key="onehot", botindex=1, topindex=2, eps=0.0001)
    loss.backward()
    names = [k for k, p in seg.named_parameters() if p.grad is not None and not re.search(r'(in_block\.conv\.0|conv\.1\.conv\.[036])\.bias$', k)]
    line = []
    for k in reversed(names):
        if k.endswith("weight"):
            line.append("%s %.1e/%.1e" % (k.replace(".conv.1.conv", ".c").replace(".conv.0", ".k2").replace(".weight", ""),
                                           rel(dict(seg.named_parameters())[k].grad, g64[k]), rel(g32[k], g64[k])))
    print("seed %d patch %d:" % (seed, patch), " | ".join(line))
