#!/usr/bin/env python
"""One eager joint step between cudaProfilerStart/Stop (ncu --profile-from-start off target).

    ncu --profile-from-start off --metrics ... --csv --log-file gpurun_out/step.csv python tools/one_step.py
Single stream (overlap off) so that every kernel is profiled alone, in program order.
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from vae_segmentation_b200 import joint_model as jm  # noqa: E402
from vae_segmentation_b200 import train_step as ts  # noqa: E402
from vae_segmentation_b200.synthetic import synth_image, synth_label  # noqa: E402

P = int(sys.argv[1]) if len(sys.argv) > 1 else 96
B = 2
dev = torch.device("cuda", 0)
torch.manual_seed(0)
mk = lambda: jm.Joint([jm.Segmentation(1, 2, norm_type=1), jm.VAE(2, 2, norm_type=1, dim=128, patch=P)])
student, teacher = mk(), mk()
teacher.load_state_dict(student.state_dict())
student.to(dev).set_precision("bf16")
teacher.to(dev).set_precision("bf16")
tr = ts.JointTrainer(student, teacher, overlap=False)
img, lab = synth_image(B, P).to(dev), synth_label(B, P).to(dev)
with torch.cuda.stream(tr.stream):
    for _ in range(3):
        tr.step(img, lab)
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStart()
    tr.step(img, lab)
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStop()
print("one step done")
