#!/usr/bin/env python
"""Where does the bf16 path stand against the north-star tolerances on CONDITIONED weights (oracle/conditioned.py)?

Prints, for Seg / VAE / joint steps: probability rel-L2, argmax agreement, loss error, per-parameter gradient rel-L2
(worst and whole-gradient) of the GPU path (bf16 and fp32 check mode) against the CPU fp32 oracle, next to the fp32
oracle's own deviation from the float64 oracle.  Diagnostic only (tests/test_models_gpu.py holds the asserts).
"""
import argparse
import os
import re
import sys
from collections import OrderedDict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from oracle import conditioned as C  # noqa: E402
from oracle import ref_torch as R  # noqa: E402

_BIAS = re.compile(r"(in_block\.conv\.0|conv\.1\.conv\.[036])\.bias$")


def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return ((a - b).norm() / b.norm().clamp_min(1e-300)).item()


def grad_report(tag, got, want, truth=None):
    keys = [k for k in want if not _BIAS.search(k)]
    per = [(rel(got[k], want[k]), k) for k in keys]
    worst = max(per)
    cat = lambda d: torch.cat([d[k].reshape(-1).double().cpu() for k in keys])
    tot = rel(cat(got), cat(want))
    line = "%s grads: whole rel-L2 %.3e  worst %.3e (%s)  median %.3e" % (
        tag, tot, worst[0], worst[1], sorted(p[0] for p in per)[len(per) // 2])
    if truth is not None:
        line += "  | fp32 oracle vs fp64: whole %.3e worst %.3e" % (
            rel(cat(want), cat(truth)), max(rel(want[k], truth[k]) for k in keys))
    print(line, flush=True)
    bad = sorted((p for p in per if p[0] > 3e-2), reverse=True)
    if bad:
        print("     > 3e-2: " + ", ".join("%s %.3f (|g| %.2e of %.2e)" % (k, e, want[k].double().norm().item(),
                                                                       cat(want).norm().item()) for e, k in bad), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seg-steps", type=int, default=60)
    ap.add_argument("--vae-steps", type=int, default=60)
    ap.add_argument("--lr", type=float, default=0.1)
    ap.add_argument("--patch", type=int, default=64)
    ap.add_argument("--batch", type=int, default=1)
    ap.add_argument("--prepare", action="store_true", help="only train + cache the conditioned weights (CPU)")
    ap.add_argument("--modes", default="seg,vae,joint")
    ap.add_argument("--precisions", default="bf16,fp32")
    args = ap.parse_args()
    seg_sd, seg_losses = C.train_seg(args.seg_steps, patch=32, lr=args.lr)
    print("conditioned Seg: loss %s" % ([round(l, 3) for l in seg_losses[:1] + seg_losses[-1:]],))
    vae_sd, vae_losses = C.train_vae(args.vae_steps, patch=args.patch, lr=args.lr)
    print("conditioned VAE: loss %s" % ([round(l, 3) for l in vae_losses[:1] + vae_losses[-1:]],))
    if args.prepare:
        return
    from vae_segmentation_b200 import evaluation as ev
    from vae_segmentation_b200 import joint_model as jm
    from vae_segmentation_b200 import train_step as ts
    from vae_segmentation_b200.synthetic import synth_label
    dev = "cuda"
    P, B = args.patch, args.batch
    grads_of = lambda m: OrderedDict((k, p.grad.detach().cpu().clone()) for k, p in m.named_parameters() if p.grad is not None)

    def build_seg(sd, prec):
        m = jm.Segmentation(1, 2, norm_type=1)
        m.load_state_dict(sd, strict=True)
        return m.to(dev).set_precision(prec)

    def build_vae(sd, prec):
        m = jm.VAE(2, 2, norm_type=1, dim=128, patch=P)
        m.load_state_dict(sd, strict=True)
        return m.to(dev).set_precision(prec)

    modes = args.modes.split(",")
    if "seg" in modes:
        torch.manual_seed(1234)
        img, label = C.blob_batch(B, P)
        loss_ref, g_ref, pred_ref = R.seg_train_step(seg_sd, img, label, eps=0.0001)
        _, g64, _ = R.seg_train_step(seg_sd, img, label, eps=0.0001, dtype=torch.float64)
        for prec in args.precisions.split(","):
            seg = build_seg(seg_sd, prec)
            pred = seg.predict(img.to(dev))
            loss = 1 - ev.avg_dsc_fused(pred, label.to(dev), "label", botindex=1, topindex=2, eps=0.0001)
            loss.backward()
            agree = (pred.argmax(1).cpu() == pred_ref.argmax(1)).float().mean().item()
            print("[seg %s %d^3 B%d] probs rel-L2 %.3e max-abs %.3e argmax %.6f loss %.5f (ref %.5f)" % (
                prec, P, B, rel(pred, pred_ref), (pred.cpu() - pred_ref).abs().max().item(), agree, loss.item(), loss_ref.item()))
            grad_report("[seg %s]" % prec, grads_of(seg), g_ref, g64)
    if "vae" in modes:
        torch.manual_seed(1235)
        label = synth_label(B, P)
        z = torch.randn(B, 128)
        loss_ref, dsc_ref, kl_ref, g_ref, rec_ref = R.vae_train_step(vae_sd, label, scale=0.35, z=z, eps=0.0001)
        _, _, _, g64, _ = R.vae_train_step(vae_sd, label, scale=0.35, z=z, eps=0.0001, dtype=torch.float64)
        for prec in args.precisions.split(","):
            vae = build_vae(vae_sd, prec)
            oh = ev.one_hot(label.to(dev), 2)
            recon, mean, std = vae(oh, if_random=True, scale=0.35, z=z)
            d = {"recon": recon, "onehot": oh, "mean": mean, "std": std}
            kl = ev.KLloss(d)
            dsc = 1 - ev.avg_dsc(d, source_key="recon", target_key="onehot", botindex=1, topindex=2, eps=0.0001)
            (dsc + 0.00002 * kl).backward()
            agree = (recon.argmax(1).cpu() == rec_ref.argmax(1)).float().mean().item()
            print("[vae %s %d^3 B%d] recon rel-L2 %.3e max-abs %.3e argmax %.6f dsc %.5f (ref %.5f) kl %.3f (ref %.3f)" % (
                prec, P, B, rel(recon, rec_ref), (recon.cpu() - rec_ref).abs().max().item(), agree, dsc.item(),
                dsc_ref.item(), kl.item(), kl_ref.item()))
            grad_report("[vae %s]" % prec, grads_of(vae), g_ref, g64)
    if "joint" in modes:
        torch.manual_seed(1236)
        img, label = C.blob_batch(B, P)
        teacher_sd, _ = C.train_seg(max(args.seg_steps - 5, 0), patch=32, lr=args.lr)
        out_ref, g_ref = R.joint_target_step(seg_sd, vae_sd, teacher_sd, img, label, lambda_vae=1.0, loss_type=0)
        _, g64 = R.joint_target_step(seg_sd, vae_sd, teacher_sd, img, label, lambda_vae=1.0, loss_type=0, dtype=torch.float64)
        for prec in args.precisions.split(","):
            student = jm.Joint([build_seg(seg_sd, prec), build_vae(vae_sd, prec)])
            teacher = jm.Joint([build_seg(teacher_sd, prec), build_vae(vae_sd, prec)])
            tr = ts.JointTrainer(student, teacher, lambda_vae=1.0, loss_type=0)
            final, mon, batch = tr.losses(img.to(dev), label.to(dev))
            tr.arena.zero_grad()
            final.backward()
            torch.cuda.synchronize()
            print("[joint %s %d^3 B%d] pred rel-L2 %.3e recon rel-L2 %.3e argmax %.6f | final %.5f (ref %.5f) recon_loss %.5f (%.5f) fake %.5f (%.5f)" % (
                prec, P, B, rel(batch["pred"], out_ref["pred"]), rel(batch["recon_pred"], out_ref["recon"]),
                (batch["pred"].argmax(1).cpu() == out_ref["pred"].argmax(1)).float().mean().item(),
                mon["final_loss"].item(), out_ref["final"].item(), mon["recon_loss"].item(), out_ref["recon_loss"].item(),
                mon["dice_loss_fake"].item(), out_ref["dsc_loss_fake"].item()))
            grad_report("[joint %s]" % prec, grads_of(student.Seg), g_ref, g64)


if __name__ == "__main__":
    main()
