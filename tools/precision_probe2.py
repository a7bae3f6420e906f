"""bf16-mode vs fp32-check-mode deviation of gradients / VAE outputs at several patch sizes."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vae_segmentation_b200 import joint_model as jm, evaluation as ev, train_step as ts
from vae_segmentation_b200.synthetic import synth_image, synth_label

def flat_grads(m):
    return torch.cat([p.grad.reshape(-1) for p in m.parameters() if p.grad is not None and p.grad.abs().max() > 0])

def rel(a, b):
    return ((a - b).norm() / b.norm()).item()

for patch in (32, 64, 96):
    torch.manual_seed(0)
    seg = jm.Segmentation(1, 2, norm_type=1).cuda()
    img, lab = synth_image(2, patch).cuda(), synth_label(2, patch).cuda()
    out = {}
    for prec in ("fp32", "bf16"):
        seg.set_precision(prec)
        for p in seg.parameters(): p.grad = None
        pr = seg.predict(img)
        (1 - ev.avg_dsc_fused(pr, lab, "label", botindex=1, topindex=2)).backward()
        out[prec] = (pr.detach(), flat_grads(seg), {k: p.grad.clone() for k, p in seg.named_parameters()})
    per = sorted((rel(out["bf16"][2][k], g), k) for k, g in out["fp32"][2].items() if g.norm() > 0)
    print("SEG P=%d: probs relL2 %.3e grads total relL2 %.3e; median layer %.3e worst %s" % (
        patch, rel(out["bf16"][0], out["fp32"][0]), rel(out["bf16"][1], out["fp32"][1]), per[len(per)//2][0], per[-2:]))

for patch in (64, 96, 128):
    torch.manual_seed(1)
    vae = jm.VAE(2, 2, norm_type=1, dim=128, patch=patch).cuda()
    lab = synth_label(1, patch).cuda()
    oh = ev.one_hot(lab, 2)
    z = torch.randn(1, 128)
    out = {}
    for prec in ("fp32", "bf16"):
        vae.set_precision(prec)
        for p in vae.parameters(): p.grad = None
        recon, mean, std = vae(oh, if_random=True, scale=0.35, z=z)
        kl = ev.KLloss({"mean": mean, "std": std})
        (1 - ev.avg_dsc_fused(recon, lab, "label", botindex=1, topindex=2) + 2e-5 * kl).backward()
        out[prec] = (recon.detach(), flat_grads(vae), mean.detach(), kl.item())
    print("VAE P=%d: recon relL2 %.3e max %.3e mean relL2 %.3e kl %.2f/%.2f grads total relL2 %.3e" % (
        patch, rel(out["bf16"][0], out["fp32"][0]), (out["bf16"][0] - out["fp32"][0]).abs().max().item(),
        rel(out["bf16"][2], out["fp32"][2]), out["bf16"][3], out["fp32"][3], rel(out["bf16"][1], out["fp32"][1])))

for patch in (64, 96):
    res = {}
    for prec in ("fp32", "bf16"):
        torch.manual_seed(2)
        mk = lambda: jm.Joint([jm.Segmentation(1, 2, norm_type=1), jm.VAE(2, 2, norm_type=1, dim=128, patch=patch)]).cuda().set_precision(prec)
        student, teacher = mk(), mk()
        teacher.load_state_dict(student.state_dict())
        tr = ts.JointTrainer(student, teacher)
        torch.manual_seed(3)
        img, lab = synth_image(2, patch).cuda(), synth_label(2, patch).cuda()
        final, mon, batch = tr.losses(img, lab)
        final.backward()
        res[prec] = (tr.arena.grad.clone(), {k: v.item() for k, v in mon.items()}, batch["recon_pred"].detach())
    print("JOINT P=%d: grads total relL2 %.3e recon relL2 %.3e losses bf16 %s fp32 %s" % (
        patch, rel(res["bf16"][0], res["fp32"][0]), rel(res["bf16"][2], res["fp32"][2]), res["bf16"][1], res["fp32"][1]))
