"""Block-by-block comparison of the fp32 check mode against the oracle (each block is fed the
oracle's own input so errors do not compound), plus run-to-run determinism."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
from collections import OrderedDict
from oracle import ref_torch as R
from vae_segmentation_b200 import joint_model as jm, ops
from vae_segmentation_b200.synthetic import synth_image, synth_label

torch.manual_seed(21)
patch = 64
sd = R.init_vae_state(2, 128, patch)
label = synth_label(2, patch)
x = R.one_hot(label)
vae = jm.VAE(2, 2, norm_type=1, dim=128, patch=patch)
vae.load_state_dict(sd)
vae = vae.cuda().set_precision("fp32")

def sub(prefix):
    return OrderedDict((k, v) for k, v in sd.items() if k.startswith(prefix + "."))

def cmp(name, got, want):
    d = (got.cpu() - want).abs().max().item()
    print("%-10s max abs err %.3e (ref max %.3e) shape %s" % (name, d, want.abs().max().item(), tuple(want.shape)))

with torch.no_grad():
    h = R.conv_in_relu(sd, "in_block.conv.0", x)
    cmp("in_block", vae.in_block(x.cuda()), h)
    for i in range(1, 6):
        h2 = R.down(sd, "down%d" % i, h)
        cmp("down%d" % i, getattr(vae, "down%d" % i)(h.cuda()), h2)
        h = h2
    flat = h.reshape(h.size(0), -1)
    mean = F.linear(flat, sd["fc_mean.weight"], sd["fc_mean.bias"])
    std = F.relu(F.linear(flat, sd["fc_std.weight"], sd["fc_std.bias"]))
    hn = h.permute(0, 2, 3, 4, 1).contiguous().cuda()
    d = lambda k: sd[k].cuda()
    m, s, lat = ops.fc_encode_fwd(hn, d("fc_mean.weight"), d("fc_mean.bias"), d("fc_std.weight"), d("fc_std.bias"),
                                  None, 1.0, False, 2, 8, 256, 128)
    cmp("fc_mean", m, mean); cmp("fc_std", s, std)
    print("std zeros ref/got:", (std == 0).sum().item(), (s == 0).sum().item(), " kl ref %.3f got %.3f" % (
        R.kl_loss(mean, std).item(), R.kl_loss(m.cpu(), s.cpu()).item()))
    h = F.linear(mean, sd["fc2.weight"], sd["fc2.bias"]).view(2, 256, 2, 2, 2)
    hd = ops.fc_decode_fwd(mean.cuda(), d("fc2.weight"), d("fc2.bias"), 2, 8, 256, 128, torch.float32, 2)
    cmp("fc2", hd.permute(0, 4, 1, 2, 3), h)
    for i in range(1, 6):
        h2 = R.up(sd, "up%d" % i, h)
        cmp("up%d" % i, getattr(vae, "up%d" % i)(h.cuda()), h2)
        h = h2
    full_ref, mr, sr = R.vae_forward(sd, x)
    a, b, c = vae(x.cuda())
    cmp("vae full", a, full_ref); cmp("mean", b, mr); cmp("std", c, sr)
    a2, b2, c2 = vae(x.cuda())
    print("run-to-run recon diff %.3e mean diff %.3e" % ((a - a2).abs().max().item(), (b - b2).abs().max().item()))
    # condition of the 2^3 level: spread of per-(n,c) variance
    hh = R.conv_in_relu(sd, "in_block.conv.0", x)
    for i in range(1, 6):
        hh = R.down(sd, "down%d" % i, hh)
    print("down5 out: frac zero %.3f" % (hh == 0).float().mean().item())

# Seg fp32 determinism of grads
torch.manual_seed(11)
ssd = R.init_seg_state()
img, lab = synth_image(2, 32), synth_label(2, 32)
from vae_segmentation_b200 import evaluation as ev
res = []
for rep in range(3):
    seg = jm.Segmentation(1, 2, norm_type=1); seg.load_state_dict(ssd); seg = seg.cuda().set_precision("fp32")
    p = seg.predict(img.cuda())
    loss = 1 - ev.avg_dsc_fused(p, lab.cuda(), "label", botindex=1, topindex=2, eps=1e-4)
    loss.backward()
    res.append((p.detach().clone(), torch.cat([q.grad.reshape(-1) for q in seg.parameters()]).clone()))
for rep in (1, 2):
    print("seg fp32 run0 vs run%d: probs %.3e grads relL2 %.3e" % (rep, (res[0][0] - res[rep][0]).abs().max().item(),
          ((res[0][1] - res[rep][1]).norm() / res[0][1].norm()).item()))
lr, gr, pr = R.seg_train_step(ssd, img, lab, eps=1e-4)
gref = torch.cat([g.reshape(-1) for g in gr.values()])
print("seg fp32 vs oracle: probs %.3e grads relL2 %.3e" % ((res[0][0].cpu() - pr).abs().max().item(),
      ((res[0][1].cpu() - gref).norm() / gref.norm()).item()))
off = 0
worst = []
for k, g in gr.items():
    n = g.numel(); mine = res[0][1][off:off + n].cpu().reshape(g.shape); off += n
    if g.norm() > 0:
        worst.append((((mine - g).norm() / g.norm()).item(), k, g.norm().item()))
print(sorted(worst, reverse=True)[:8])
