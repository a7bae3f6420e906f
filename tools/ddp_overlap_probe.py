#!/usr/bin/env python
"""2-rank probe of the bucketed (overlapped) gradient all-reduce: eager steps, then a CUDA-graph captured step.
torchrun --nproc-per-node 2 tools/ddp_overlap_probe.py ; every phase is guarded by a faulthandler watchdog that dumps
all Python stacks and exits instead of hanging the GPU box."""
import faulthandler
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["VAESEG_DDP_OVERLAP"] = "1"
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
log = open(os.path.join(ROOT, "gpurun_out", "ddp_probe_rank%d.log" % rank), "w")


def say(*a):
    print(*a, file=log, flush=True)


faulthandler.enable(file=log)
faulthandler.dump_traceback_later(45, exit=True, file=log)
torch.cuda.set_device(rank)
dev = torch.device("cuda", rank)
dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
from vae_segmentation_b200 import joint_model as jm  # noqa: E402
from vae_segmentation_b200 import train_step as ts  # noqa: E402
from vae_segmentation_b200.synthetic import synth_image, synth_label  # noqa: E402

P = 64
torch.manual_seed(3)
mk = lambda: jm.Joint([jm.Segmentation(1, 2, norm_type=1), jm.VAE(2, 2, norm_type=1, dim=128, patch=P)])
student, teacher = mk(), mk()
teacher.load_state_dict(student.state_dict())
student.to(dev).set_precision("bf16")
teacher.to(dev).set_precision("bf16")
tr = ts.JointTrainer(student, teacher, lr=0.0)
torch.manual_seed(99)
img, lab = synth_image(2, P), synth_label(2, P)
xi, xl = img[rank:rank + 1].to(dev), lab[rank:rank + 1].to(dev)
say("built; overlap", tr.ddp_overlap)
with torch.cuda.stream(tr.stream):
    for i in range(2):
        tr.step(xi, xl)
        torch.cuda.synchronize()
        say("eager step", i, "reduced", tr._grads_reduced, "grad sum %.6e" % tr.arena.grad.double().sum().item())
    eager = tr.arena.grad.clone()
    # reference: plain all-reduce path on the same data
    tr.ddp_overlap = False
    tr.step(xi, xl)
    torch.cuda.synchronize()
    plain = tr.arena.grad.clone()
    say("plain path grad sum %.6e  rel diff %.3e" % (plain.double().sum().item(), ((plain - eager).norm() / plain.norm()).item()))
    tr.ddp_overlap = True
    faulthandler.cancel_dump_traceback_later()
    faulthandler.dump_traceback_later(45, exit=True, file=log)
    say("capturing")
    tr.capture(xi, xl, warmup=1)
    say("captured")
    for i in range(3):
        tr.step_graphed()
    torch.cuda.synchronize()
    g = tr.arena.grad
    say("graph grad sum %.6e  rel diff vs eager %.3e" % (g.double().sum().item(), ((g - eager).norm() / eager.norm()).item()))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(20):
        tr.step_graphed()
    e1.record()
    torch.cuda.synchronize()
    say("graphed step %.3f ms" % (e0.elapsed_time(e1) / 20))
faulthandler.cancel_dump_traceback_later()
say("ok")
dist.destroy_process_group()
