import sys, os, traceback
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vae_segmentation_b200 import joint_model as jm, train_step as ts, evaluation as ev
from vae_segmentation_b200.synthetic import synth_image, synth_label
P = 32
torch.manual_seed(0)
seg = jm.Segmentation(1, 2, norm_type=1).cuda()
img, lab = synth_image(1, P).cuda(), synth_label(1, P).cuda()
tr = ts.SegTrainer(seg)
side = torch.cuda.Stream()
side.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(side):
    for _ in range(2): tr.step(img, lab)
torch.cuda.current_stream().wait_stream(side)
torch.cuda.synchronize()

def attempt(name, fn):
    g = torch.cuda.CUDAGraph()
    try:
        with torch.cuda.graph(g, stream=side):
            fn()
        g.replay(); torch.cuda.synchronize()
        print("OK  ", name)
    except Exception as e:
        print("FAIL", name, "->", str(e).splitlines()[0])
        traceback.print_exc(limit=6)
    torch.cuda.synchronize()

def fwd():
    with torch.no_grad():
        seg.predict(img)
attempt("forward no_grad", fwd)
def fwd_grad():
    seg.predict(img)
attempt("forward with tape", fwd_grad)
def fwd_loss():
    tr.loss(img, lab)
attempt("forward + loss", fwd_loss)
def full():
    tr.arena.zero_grad()
    loss, _ = tr.loss(img, lab)
    loss.backward()
attempt("forward + loss + backward", full)

P = 32 * 2
torch.manual_seed(0)
mk = lambda: jm.Joint([jm.Segmentation(1, 2, norm_type=1), jm.VAE(2, 2, norm_type=1, dim=128, patch=P)]).cuda()
student, teacher = mk(), mk()
teacher.load_state_dict(student.state_dict())
jt = ts.JointTrainer(student, teacher)
img, lab = synth_image(1, P).cuda(), synth_label(1, P).cuda()
side.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(side):
    for _ in range(2): jt.step(img, lab)
torch.cuda.current_stream().wait_stream(side)
torch.cuda.synchronize()
def j_student_fwd():
    student({"img": img}, "img", "pred", "recon", dropout=True)
attempt("joint: student fwd", j_student_fwd)
def j_teacher_fwd():
    with torch.no_grad():
        teacher({"img": img}, "img", "tp", "tr")
attempt("joint: teacher fwd", j_teacher_fwd)
def j_losses():
    jt.losses(img, lab)
attempt("joint: losses", j_losses)
def j_vae_bwd():
    b = student({"img": img}, "img", "pred", "recon", dropout=True)
    (1 - ev.avg_dsc_fused(b["pred"].detach(), b["recon"], "tensor", botindex=1, topindex=2)).backward()
attempt("joint: vae-path backward", j_vae_bwd)
def j_full():
    jt.arena.zero_grad()
    final, mon, _ = jt.losses(img, lab)
    final.backward()
attempt("joint: full", j_full)
