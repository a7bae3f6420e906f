#!/usr/bin/env python
"""Kernel timeline of ONE captured joint step (CUPTI through torch.profiler; there is no nsys in the image).

Writes gpurun_out/timeline_<tag>.csv (name, stream, start us, duration us) for one graph replay and prints, per
stream, the busy time, the span and the idle gaps, plus the kernels of the main (critical) stream grouped by name.
A number printed under the profiler is a diagnostic, never a bench value.
"""
import argparse
import collections
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
from torch.profiler import ProfilerActivity, profile  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--patch", type=int, default=96)
    ap.add_argument("--batch", type=int, default=2)
    ap.add_argument("--tag", default="r1")
    ap.add_argument("--no-overlap", action="store_true")
    args = ap.parse_args()
    from vae_segmentation_b200 import joint_model as jm
    from vae_segmentation_b200 import train_step as ts
    from vae_segmentation_b200.synthetic import synth_image, synth_label
    dev = torch.device("cuda", 0)
    P, B = args.patch, args.batch
    torch.manual_seed(0)
    mk = lambda: jm.Joint([jm.Segmentation(1, 2, norm_type=1), jm.VAE(2, 2, norm_type=1, dim=128, patch=P)])
    student, teacher = mk(), mk()
    teacher.load_state_dict(student.state_dict())
    student.to(dev).set_precision("bf16")
    teacher.to(dev).set_precision("bf16")
    tr = ts.JointTrainer(student, teacher, overlap=not args.no_overlap)
    img, lab = synth_image(B, P).to(dev), synth_label(B, P).to(dev)
    with torch.cuda.stream(tr.stream):
        for _ in range(3):
            tr.step(img, lab)
        tr.capture(img, lab, warmup=1)
        for _ in range(3):
            tr.step_graphed()
        torch.cuda.synchronize()
        with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
            for _ in range(3):
                tr.step_graphed()
            torch.cuda.synchronize()
    evs = []
    if True:
        # older/newer kineto: fall back to the raw trace
        import json
        import tempfile
        path = tempfile.mktemp(suffix=".json")
        prof.export_chrome_trace(path)
        tr_json = json.load(open(path))
        for e in tr_json["traceEvents"]:
            if e.get("cat") in ("kernel", "gpu_memset", "gpu_memcpy"):
                evs.append((e["ts"], e["dur"], e["name"], e.get("args", {}).get("stream", e.get("tid"))))
    evs.sort()
    # keep the middle replay: split on the sgd kernel (one per step, the last launch of a step)
    ends = [i for i, e in enumerate(evs) if "sgd_kernel" in e[2]]
    if len(ends) >= 2:
        evs = evs[ends[0] + 1:ends[1] + 1]
    t0 = evs[0][0]
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    out = os.path.join(ROOT, "gpurun_out", "timeline_%s.csv" % args.tag)
    with open(out, "w") as f:
        f.write("start_us,dur_us,stream,name\n")
        for s, d, nm, st in evs:
            f.write("%.2f,%.2f,%s,%s\n" % (s - t0, d, st, nm.replace(",", ";")[:100]))
    span = max(s + d for s, d, _, _ in evs) - t0
    print("one step: %d device activities, span %.1f us" % (len(evs), span))
    by_stream = collections.defaultdict(list)
    for e in evs:
        by_stream[e[3]].append(e)
    for st, lst in sorted(by_stream.items(), key=lambda kv: -len(kv[1])):
        busy = sum(d for _, d, _, _ in lst)
        gaps = []
        for a, b in zip(lst, lst[1:]):
            gaps.append(b[0] - (a[0] + a[1]))
        print("stream %s: %d activities, busy %.1f us, first %.1f last-end %.1f, median gap %.2f us, gap sum %.1f us" % (
            st, len(lst), busy, lst[0][0] - t0, lst[-1][0] + lst[-1][1] - t0,
            sorted(gaps)[len(gaps) // 2] if gaps else 0, sum(g for g in gaps if g > 0)))
    main_stream = max(by_stream.items(), key=lambda kv: len(kv[1]))[0]
    agg = collections.defaultdict(lambda: [0, 0.0])
    for s, d, nm, st in by_stream[main_stream]:
        a = agg[nm[:60]]
        a[0] += 1
        a[1] += d
    print("main stream %s by kernel:" % main_stream)
    for nm, (cnt, tot) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:30]:
        print("  %-62s %4d %9.1f us" % (nm, cnt, tot))


if __name__ == "__main__":
    main()
