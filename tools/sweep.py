#!/usr/bin/env python
"""BASELINE.json configs[4]: throughput sweep over patch size and per-GPU batch (joint step), one bench.py run per cell.

    python tools/sweep.py --gpus 1 --patches 64,96,128,160 --batches 1,2,4,8 --out profiles/r2_sweep_1gpu.md

Each cell is `bench.py --mode joint --patch P --batch B --no-roofline --no-cpu-baseline` (resident-input and end-to-end
volumes/s, CUDA-event timed, max over ranks); for N > 1 the cell is launched through torch.distributed.run like the
driver does.  The CPU-reference column comes from `bench.py --impl reference --patch P` (batch 1 per step)."""
import argparse
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run(cmd, timeout):
    try:
        out = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=timeout)
    except subprocess.TimeoutExpired:
        return None, "timeout"
    for line in out.stdout.splitlines():
        if line.startswith("{"):
            return json.loads(line), ""
    return None, (out.stderr or out.stdout)[-300:]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--patches", default="64,96,128,160")
    ap.add_argument("--batches", default="1,2,4,8")
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--cpu", action="store_true", help="also time the CPU reference arm per patch size")
    ap.add_argument("--out", default="")
    args = ap.parse_args()
    patches = [int(v) for v in args.patches.split(",")]
    batches = [int(v) for v in args.batches.split(",")]
    rows, port = [], 29700
    cpu = {}
    for P in patches:
        if args.cpu:
            line, err = run([sys.executable, "bench.py", "--impl", "reference", "--patch", str(P), "--steps", "2", "--warmup", "1"], 900)
            cpu[P] = line["value"] if line else None
        for B in batches:
            base = ["bench.py", "--gpus", str(args.gpus), "--mode", "joint", "--patch", str(P), "--batch", str(B), "--steps",
                    str(args.steps), "--warmup", "3", "--no-roofline", "--no-cpu-baseline"]
            if args.gpus > 1:
                port += 1
                cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(args.gpus),
                       "--master-addr", "127.0.0.1", "--master-port", str(port)] + base
            else:
                cmd = [sys.executable] + base
            line, err = run(cmd, 600)
            if line is None:
                rows.append((P, B, None, None, None, err.strip().splitlines()[-1] if err.strip() else "failed"))
            else:
                rows.append((P, B, line["value"], line["e2e"]["value"], line["ms_per_step"], ""))
            print(rows[-1], flush=True)
    lines = ["| patch | per-GPU batch | GPUs | vol/s (resident) | vol/s (e2e) | ms/step | CPU reference vol/s |", "|---|---|---|---|---|---|---|"]
    for P, B, v, e, ms, err in rows:
        c = cpu.get(P)
        lines.append("| %d^3 | %d | %d | %s | %s | %s | %s |" % (
            P, B, args.gpus, "%.1f" % v if v else "-- (%s)" % err, "%.1f" % e if e else "--", "%.3f" % ms if ms else "--",
            "%.3f" % c if c else "--"))
    text = "\n".join(lines) + "\n"
    print(text)
    if args.out:
        with open(os.path.join(ROOT, args.out), "w") as f:
            f.write("joint teacher-student step, bf16 tcgen05 path, CUDA-graph replay; tools/sweep.py\n\n" + text)


if __name__ == "__main__":
    main()
