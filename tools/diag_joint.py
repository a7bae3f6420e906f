import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from collections import OrderedDict
from oracle import ref_torch as R
from vae_segmentation_b200 import joint_model as jm, train_step as ts, evaluation as ev
from vae_segmentation_b200.synthetic import synth_image, synth_label

patch = 64
torch.manual_seed(31)
seg_sd = R.init_seg_state()
vae_sd = R.init_vae_state(2, 128, patch)
teacher_sd = OrderedDict((k, v + 0.01 * torch.randn_like(v)) for k, v in seg_sd.items())
img, label = synth_image(1, patch), synth_label(1, patch)

def build(sd, cls, **kw):
    m = cls(**kw); m.load_state_dict(sd); return m.cuda().set_precision("fp32")
seg = lambda sd: build(sd, lambda: jm.Segmentation(1, 2, norm_type=1))
vae = lambda: build(vae_sd, lambda: jm.VAE(2, 2, norm_type=1, dim=128, patch=patch))

# oracle pieces
sd = R._leafify(seg_sd); vsd = R._leafify(vae_sd, False); tsd = R._leafify(teacher_sd, False)
pred, recon, _, _ = R.joint_forward(sd, vsd, img, dropout=True)
with torch.no_grad():
    t_pred, _, t_mean, t_std = R.joint_forward(tsd, vsd, img, dropout=False)
print("oracle kl", R.kl_loss(t_mean, t_std).item(), "zeros", (t_std == 0).sum().item())

student = jm.Joint([seg(seg_sd), vae()]); teacher = jm.Joint([seg(teacher_sd), vae()])
x = img.cuda()
b = student({"img": x}, "img", "pred", "recon", dropout=True)
with torch.no_grad():
    tb = teacher({"img": x}, "img", "tp", "tr")
e = lambda a, r: (a.detach().cpu() - r.detach()).abs().max().item()
print("pred err %.3e recon err %.3e teacher pred err %.3e mean err %.3e std err %.3e" % (
    e(b["pred"], pred), e(b["recon"], recon), e(tb["tp"], t_pred), e(tb["mean"], t_mean), e(tb["std"], t_std)))
print("kl ours", ev.KLloss(tb).item(), "zeros", (tb["std"] == 0).sum().item(), "kl of ours on cpu", R.kl_loss(tb["mean"].cpu(), tb["std"].cpu()).item())
# term by term gradients
def grads_ours(loss):
    for p in student.parameters(): p.grad = None
    loss.backward(retain_graph=True)
    return OrderedDict((k, p.grad.detach().cpu().clone()) for k, p in student.Seg.named_parameters())
def grads_ref(loss):
    for v in sd.values(): v.grad = None
    loss.backward(retain_graph=True)
    return OrderedDict((k, v.grad.clone()) for k, v in sd.items())
def cmp(name, go, gr):
    w = sorted(((( go[k] - gr[k]).norm() / gr[k].norm().clamp_min(1e-30)).item(), k) for k in gr if gr[k].norm() > 1e-7)
    tot = (torch.cat([go[k].reshape(-1) for k in gr]) - torch.cat([gr[k].reshape(-1) for k in gr])).norm() / torch.cat([gr[k].reshape(-1) for k in gr]).norm()
    print(name, "total relL2 %.3e worst" % tot.item(), w[-3:])
pseudo = R.binarize(t_pred)
r_fake = 1 - R.avg_dsc(pred, pseudo, botindex=1, topindex=2)
o_fake = 1 - ev.avg_dsc_fused(b["pred"], tb["tp"], "binarize", botindex=1, topindex=2)
print("fake loss", r_fake.item(), o_fake.item())
cmp("fake-term grads", grads_ours(o_fake), grads_ref(r_fake))
r_rec = 1 - R.avg_dsc(pred, recon, botindex=1, topindex=2)
o_rec = 1 - ev.avg_dsc_fused(b["pred"], b["recon"], "tensor", botindex=1, topindex=2)
print("recon loss", r_rec.item(), o_rec.item())
cmp("recon-term grads", grads_ours(o_rec), grads_ref(r_rec))
# recon term with detached target (no VAE path) and with detached source (VAE path only)
cmp("recon-term, target detached", grads_ours(1 - ev.avg_dsc_fused(b["pred"], b["recon"].detach(), "tensor", botindex=1, topindex=2)),
    grads_ref(1 - R.avg_dsc(pred, recon.detach(), botindex=1, topindex=2)))
cmp("recon-term, source detached", grads_ours(1 - ev.avg_dsc_fused(b["pred"].detach(), b["recon"], "tensor", botindex=1, topindex=2)),
    grads_ref(1 - R.avg_dsc(pred.detach(), recon, botindex=1, topindex=2)))

# ---- conditioning check: the same recon-term gradient from the oracle in float64 ----
def to64(d, rg):
    return OrderedDict((k, v.detach().double().requires_grad_(rg)) for k, v in d.items())
sd64, vsd64 = to64(seg_sd, True), to64(vae_sd, False)
torch.manual_seed(0)
pred64, recon64, _, _ = R.joint_forward(sd64, vsd64, img.double(), dropout=True)
l64 = 1 - R.avg_dsc(pred64.detach(), recon64, botindex=1, topindex=2)
l64.backward()
g64 = OrderedDict((k, v.grad.float()) for k, v in sd64.items())
g32 = grads_ref(1 - R.avg_dsc(pred.detach(), recon, botindex=1, topindex=2))
go = grads_ours(1 - ev.avg_dsc_fused(b["pred"].detach(), b["recon"], "tensor", botindex=1, topindex=2))
cmp("VAE-path grads: oracle fp32 vs oracle fp64", g32, g64)
cmp("VAE-path grads: ours   fp32 vs oracle fp64", go, g64)
print("recon fwd: oracle32 vs 64 %.3e ; ours vs 64 %.3e" % ((recon.detach().double() - recon64.detach()).abs().max().item(),
      (b["recon"].detach().cpu().double() - recon64.detach()).abs().max().item()))
