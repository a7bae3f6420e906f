"""Attributes bf16-mode gradient error to storage points by rounding tensor classes through bf16 inside fp32 mode."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vae_segmentation_b200 import joint_model as jm, evaluation as ev, engine
from vae_segmentation_b200.synthetic import synth_image, synth_label

def rel(a, b): return ((a - b).norm() / b.norm()).item()
patch = 64
torch.manual_seed(0)
seg = jm.Segmentation(1, 2, norm_type=1).cuda().set_precision("fp32")
img, lab = synth_image(2, patch).cuda(), synth_label(2, patch).cuda()
def run(sim):
    engine.SIMULATE_BF16 = set(sim)
    for p in seg.parameters(): p.grad = None
    pr = seg.predict(img)
    (1 - ev.avg_dsc_fused(pr, lab, "label", botindex=1, topindex=2)).backward()
    engine.SIMULATE_BF16 = set()
    return pr.detach(), torch.cat([p.grad.reshape(-1) for p in seg.parameters()])
p0, g0 = run([])
for sim in (["y"], ["a"], ["k2"], ["y", "a", "k2"], ["g"], ["dy"], ["g", "dy"], ["y", "a", "k2", "g", "dy"]):
    p1, g1 = run(sim)
    print("%-28s probs relL2 %.3e grads relL2 %.3e" % ("+".join(sim), rel(p1, p0), rel(g1, g0)))
