import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from collections import OrderedDict
from oracle import ref_torch as R
from vae_segmentation_b200 import joint_model as jm, train_step as ts, evaluation as ev
from vae_segmentation_b200.synthetic import synth_image, synth_label
patch = 64
torch.manual_seed(31)
seg_sd = R.init_seg_state(); vae_sd = R.init_vae_state(2, 128, patch)
teacher_sd = OrderedDict((k, v + 0.01 * torch.randn_like(v)) for k, v in seg_sd.items())
img, label = synth_image(1, patch), synth_label(1, patch)
with torch.no_grad():
    t_pred, _, t_mean, t_std = R.joint_forward(teacher_sd, vae_sd, img, dropout=False)
print("oracle kl", R.kl_loss(t_mean, t_std).item())
def build(sd, mk):
    m = mk(); m.load_state_dict(sd); return m.cuda().set_precision("fp32")
mkseg = lambda: jm.Segmentation(1, 2, norm_type=1)
mkvae = lambda: jm.VAE(2, 2, norm_type=1, dim=128, patch=patch)
student = jm.Joint([build(seg_sd, mkseg), build(vae_sd, mkvae)])
teacher = jm.Joint([build(teacher_sd, mkseg), build(vae_sd, mkvae)])
x = img.cuda()
def probe(tag):
    with torch.no_grad():
        tb = teacher({"img": x}, "img", "tp", "tr")
    print("%-28s teacher pred err %.3e mean err %.3e kl %.3f" % (tag, (tb["tp"].cpu() - t_pred).abs().max().item(),
          (tb["mean"].cpu() - t_mean).abs().max().item(), ev.KLloss(tb).item()))
probe("fresh")
tr = ts.JointTrainer(student, teacher, lr=1e-2, momentum=0.9, lambda_vae=1.0, loss_type=0, kl=False)
probe("after trainer construction")
final, mon, batch = tr.losses(x, label.cuda())
print("losses() kl", mon["kl_loss"].item())
probe("after losses()")
