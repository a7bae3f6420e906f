import sys, os, traceback
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vae_segmentation_b200 import joint_model as jm, train_step as ts, evaluation as ev
from vae_segmentation_b200.synthetic import synth_image, synth_label
P, B = int(sys.argv[1]), int(sys.argv[2])
default_first = sys.argv[3] == "1"
torch.manual_seed(0)
mk = lambda: jm.Joint([jm.Segmentation(1, 2, norm_type=1), jm.VAE(2, 2, norm_type=1, dim=128, patch=P)]).cuda()
student, teacher = mk(), mk()
teacher.load_state_dict(student.state_dict())
jt = ts.JointTrainer(student, teacher)
img, lab = synth_image(B, P).cuda(), synth_label(B, P).cuda()
if default_first:
    for _ in range(2): jt.step(img, lab)
    torch.cuda.synchronize()
try:
    jt.capture(img, lab, warmup=1)
    for _ in range(3): mon = jt.step_graphed()
    torch.cuda.synchronize()
    print("OK", P, B, default_first, {k: round(v.item(), 4) for k, v in mon.items()})
except Exception as e:
    print("FAIL", P, B, default_first, str(e).splitlines()[0])
    traceback.print_exc(limit=12)
