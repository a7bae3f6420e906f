#!/usr/bin/env python
"""Headline benchmark: joint teacher-student train step (seg + frozen-VAE recon + pseudo loss)
at 96^3 patches -- BASELINE.json config[2] ("C3" in SURVEY.md section 8), volumes/sec.

  python bench.py --gpus N --steps K --warmup W            our arm (one process per GPU)
  python bench.py --impl reference --steps K --warmup W    reference arm: the CPU oracle port of
                                                           the reference's PyTorch path, host cores

One JSON line on stdout (rank 0).  `value` times the step with inputs resident in HBM; `e2e`
times the same step fed from pinned HOST buffers with the loss read back every step.  The
roofline object is for the kernel signature that takes the largest share of the step, timed
with CUDA events around its launches inside the timed region.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

PATCH = 96
PER_GPU_BATCH = 2
METRIC = "joint train-step volumes/sec at 96^3 patch"
UNIT = "volumes/s"
RIDGE = 249.0          # flop/byte, measured peaks

# BASELINE.json configs -> bench modes.  The default (no flags) is configs[2], the one the metric is quoted on.
MODES = {
    # mode: (default patch, metric, workload description)
    "joint": (96, "joint train-step volumes/sec at %d^3 patch",
              "joint teacher-student step (student Seg+frozen VAE fwd, teacher Joint fwd, recon + pseudo Dice, bwd "
              "through VAE into Seg, SGD m=.9), BASELINE.json config[2]"),
    "seg": (96, "seg train-step volumes/sec at %d^3 patch",
            "source-domain Segmentation train step (fwd, 1 - Dice_fg, bwd, SGD m=.9; main_source.py:415-437), "
            "BASELINE.json config[1]"),
    "vae": (64, "vae train-step volumes/sec at %d^3 patch",
            "shape-VAE train step on one-hot masks (if_random, scale .35, 1 - Dice_fg + 2e-5 KL, bwd, SGD m=.9; "
            "main_source.py:389-406), BASELINE.json config[0]"),
    "joint_ttt": (128, "joint train-step volumes/sec at %d^3 patch (dynamic lambda type 8) + test-time-training cases/sec",
                  "joint teacher-student step with dynamic lambda (domain_loss_type 8) + validation cases with "
                  "val_finetune=1 test-time training (scripts/target/domain_msd_dh_ft1.bash; main_target.py:550-560,"
                  "807-900), BASELINE.json config[3]"),
}


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(path):
        p = json.load(open(path))
        return {"hbm": p["hbm_gbs"], "tensor_burst": p["bf16_tflops"],
                "tensor": p.get("bf16_tflops_sustained", p["bf16_tflops"]), "source": "measured"}
    return {"hbm": 6650.0, "tensor_burst": 1590.0, "tensor": 1400.0, "source": "fallback"}


# ---------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the oracle port of the reference's PyTorch CPU path
# ---------------------------------------------------------------------------------------------
def cpu_steps(mode, steps, warmup, batch, patch, seed=0):
    """Times `steps` reference-faithful train steps of `mode` on the host cores with all threads (oracle port of the
    reference's torch CPU fp32 path: main_target.py:520-592,734-736 / main_source.py:389-437).  Seconds per step."""
    from oracle import ref_torch as R
    from vae_segmentation_b200.synthetic import synth_image, synth_label
    torch.set_num_threads(os.cpu_count() or 1)
    torch.manual_seed(seed)
    seg_sd = R.init_seg_state() if mode != "vae" else None
    vae_sd = R.init_vae_state(2, 128, patch) if mode != "seg" else None
    img, label = synth_image(batch, patch), synth_label(batch, patch)
    z = torch.randn(batch, 128)
    bufs = None
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        if mode == "seg":
            _, grads, _ = R.seg_train_step(seg_sd, img, label, eps=0.0001)
            seg_sd, bufs = R.sgd_step(seg_sd, grads, bufs, lr=1e-2, momentum=0.9)
        elif mode == "vae":
            _, _, _, grads, _ = R.vae_train_step(vae_sd, label, scale=0.35, z=z, eps=0.0001)
            vae_sd, bufs = R.sgd_step(vae_sd, grads, bufs, lr=1e-2, momentum=0.9)
        else:
            _, grads = R.joint_target_step(seg_sd, vae_sd, seg_sd, img, label, lambda_vae=1.0,
                                           loss_type=8 if mode == "joint_ttt" else 0)
            seg_sd, bufs = R.sgd_step(seg_sd, grads, bufs, lr=1e-2, momentum=0.9)
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
    return sum(times) / len(times)


def cpu_joint_steps(steps, warmup, batch, patch=PATCH, seed=0):
    return cpu_steps("joint", steps, warmup, batch, patch, seed)


def run_reference_arm(args, rank):
    if rank != 0:
        return
    mode = args.mode
    patch = args.patch or MODES[mode][0]
    batch = 1            # bounded sample: one volume per step (the CPU step is seconds long)
    sec = cpu_steps(mode, args.steps, args.warmup, batch, patch)
    value = batch / sec
    cores = torch.get_num_threads()
    sample = "%d steps x %d volume(s) of the %d^3 %s step (oracle port, torch CPU fp32)" % (args.steps, batch, patch, mode)
    line = {"impl": "reference", "metric": MODES[mode][1] % patch, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": MODES[mode][2], "mode": mode,
                       "patch": patch, "per_gpu_batch": args.batch, "global_batch": args.batch * max(args.gpus, 1),
                       "lambda_vae": 1.0, "loss_type": 8 if mode == "joint_ttt" else args.loss_type,
                       "parallelism": "host cores (torch CPU)", "sample_batch_per_step": batch},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------
# clocks
# ---------------------------------------------------------------------------------------------
class ClockSampler(object):
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu_index = gpu_index
        self.tmp = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "200"], stdout=self.tmp, stderr=subprocess.DEVNULL)
        except OSError:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        self.tmp.flush()
        rows = [r.strip().split(",") for r in open(self.tmp.name).read().splitlines() if r.strip()]
        os.unlink(self.tmp.name)
        mhz, reasons, mx = [], set(), None
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        for r in rows:
            try:
                if int(r[0]) != self.gpu_index:
                    continue
                mhz.append(float(r[1]))
                mx = float(r[2])
                for nm, v in zip(names, r[5:9]):
                    if v.strip().lower() == "active":
                        reasons.add(nm)
            except (ValueError, IndexError):
                continue
        if mhz:
            mhz.sort()
            out = {"sm_mhz": mhz[len(mhz) // 2], "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(mhz)}
        return out


# ---------------------------------------------------------------------------------------------
# per-kernel accounting
# ---------------------------------------------------------------------------------------------
def kernel_cost(name, key):
    """ALGORITHMIC (flops, bytes) of one launch, from its shape signature."""
    size = {0: 4, 1: 2}
    if name == "vs_conv3x3x3_fprop":
        idt, odt, _, _, _, n, d, h, w, cin, cout = key
        vox = n * d * h * w
        return 2.0 * 27 * cin * cout * vox, vox * (cin * size[idt] + cout * size[odt])
    if name == "vs_conv3x3x3_dgrad":
        idt, odt, _, n, d, h, w, cin, cout = key
        vox = n * d * h * w
        return 2.0 * 27 * cin * cout * vox, vox * (cout * size[idt] + cin * size[odt])
    if name == "vs_conv3x3x3_wgrad":
        dt, _, _, _, n, d, h, w, cin, cout = key
        vox = n * d * h * w
        return 2.0 * 27 * cin * cout * vox, vox * (cin + cout) * size[dt]
    if name in ("vs_conv3x3x3_tc_kdn", "vs_conv3x3x3_tc_kdn_ex"):
        # kd-in-N tensor-core convolution (fprop or dgrad): bf16 in and out, GEMM input gin / output gout channels
        n, d, h, w, gin, gout = key[-6:]
        vox = n * d * h * w
        return 2.0 * 27 * gin * gout * vox, vox * (gin + gout) * 2
    if name == "vs_conv3x3x3_tc_kdn_planar":
        # 8-output-channel kd-in-N convolution storing channels 0..1 as planar fp32 (2-class head / in-block dgrad)
        _, n, d, h, w, gin = key[-6:]
        vox = n * d * h * w
        return 2.0 * 27 * gin * 8 * vox, vox * (gin * 2 + 8)
    if name in ("vs_k2s2_gather", "vs_k2s2_scatter", "vs_k2s2_wgrad"):
        dt = key[0]
        n, dc, hc, wc, a, b = key[-6:]
        vox = n * dc * hc * wc
        return 2.0 * 8 * a * b * vox, vox * (a + 8 * b) * size[dt]
    if name in ("vs_k2s2_gather_tc", "vs_k2s2_scatter_tc"):
        n, dc, hc, wc, a, b = key[-6:]
        vox = n * dc * hc * wc
        return 2.0 * 8 * a * b * vox, vox * (a + 8 * b) * 2
    if name == "vs_inorm_relu_apply":
        dt, n, s, c = key
        return 0.0, 2.0 * n * s * c * size[dt]
    if name in ("vs_inorm_relu_bwd_reduce",):
        dt, n, s, c, _ = key
        return 0.0, 2.0 * n * s * c * size[dt]
    if name == "vs_inorm_relu_bwd_apply":
        dt, n, s, c = key
        return 0.0, 3.0 * n * s * c * size[dt]
    return 0.0, 0.0


class KernelTimer(object):
    """_cabi profiler hook: CUDA events (on the launching stream) around selected calls."""

    def __init__(self, only_name=None):
        self.only_name = only_name
        self.events = {}
        from vae_segmentation_b200 import _cabi
        self._key = _cabi.call_key

    def __call__(self, name, args, fn):
        if self.only_name is not None and name != self.only_name:
            return fn(*args)
        sig = (name, self._key(name, args))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        rc = fn(*args)
        e1.record()
        self.events.setdefault(sig, []).append((e0, e1))
        return rc

    def totals(self):
        torch.cuda.synchronize()
        return {sig: (sum(a.elapsed_time(b) for a, b in evs), len(evs)) for sig, evs in self.events.items()}


# ---------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--mode", default="joint", choices=sorted(MODES),
                    help="which BASELINE.json config: joint = configs[2] (default, the metric's), seg = configs[1], "
                         "vae = configs[0], joint_ttt = configs[3]")
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--patch", type=int, default=0, help="patch edge (default: the mode's)")
    ap.add_argument("--batch", type=int, default=PER_GPU_BATCH)
    ap.add_argument("--loss-type", type=int, default=0)
    ap.add_argument("--ttt-cases", type=int, default=8, help="joint_ttt: validation cases per rank in the TTT leg")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="eager launches instead of the CUDA-graph-captured step")
    ap.add_argument("--no-roofline", action="store_true", help="skip the per-kernel event pass (sweeps)")
    ap.add_argument("--kernel-table", action="store_true", help="also print the per-kernel time table to stderr")
    ap.add_argument("--lean-teacher", action="store_true",
                    help="joint modes: skip the frozen teacher's VAE forward, whose outputs only feed the MONITORED kl term "
                         "(the reference computes it, main_target.py:532; not the headline configuration)")
    ap.add_argument("--no-e2e-prefetch", action="store_true",
                    help="e2e leg: copy -> step -> read back serially on one stream instead of the default "
                         "double-buffered inputs (the host->device copy of step i+1 overlaps step i)")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference_arm(args, rank)
        return

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the hot path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    stdout_fd = None
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
            os.environ["NCCL_DEBUG"] = "WARN"
        # NCCL prints its version banner on stdout whatever the debug level: park fd 1 on stderr until the JSON line so
        # that stdout carries exactly that one line
        sys.stdout.flush()
        stdout_fd = os.dup(1)
        os.dup2(2, 1)
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)

    from vae_segmentation_b200 import _cabi
    from vae_segmentation_b200 import joint_model as jm
    from vae_segmentation_b200 import train_step as ts
    from vae_segmentation_b200.synthetic import synth_image, synth_label

    mode = args.mode
    P, B = args.patch or MODES[mode][0], args.batch
    loss_type = 8 if mode == "joint_ttt" else args.loss_type
    torch.manual_seed(0)                     # identical replicas on every rank
    mk = lambda: jm.Joint([jm.Segmentation(1, 2, norm_type=1), jm.VAE(2, 2, norm_type=1, dim=128, patch=P)])
    finetune = None
    if mode in ("joint", "joint_ttt"):
        student, teacher = mk(), mk()
        teacher.load_state_dict(student.state_dict())          # teacher = copy of student (main_target.py:428)
        student.to(dev).set_precision(args.precision)
        teacher.to(dev).set_precision(args.precision)
        trainer = ts.JointTrainer(student, teacher, lr=1e-2, momentum=0.9, lambda_vae=1.0, loss_type=loss_type,
                                  faithful_teacher=not args.lean_teacher)
        if mode == "joint_ttt":
            finetune = mk()
            finetune.load_state_dict(student.state_dict())
            finetune.to(dev).set_precision(args.precision)
    elif mode == "seg":
        trainer = ts.SegTrainer(jm.Segmentation(1, 2, norm_type=1).to(dev).set_precision(args.precision))
    else:
        trainer = ts.VAETrainer(jm.VAE(2, 2, norm_type=1, dim=128, patch=P).to(dev).set_precision(args.precision))

    torch.manual_seed(1000 + rank)           # per-rank shard of the global batch
    # host (pinned) inputs of one step and their static device twins, in the order the trainer's step takes them
    if mode == "vae":
        host = [synth_label(B, P).pin_memory(), torch.randn(B, 128).pin_memory()]      # label, z (CPU generator, F11)
    else:
        host = [synth_image(B, P).pin_memory(), synth_label(B, P).pin_memory()]
    devs = [t.to(dev) for t in host]
    h2d = sum(t.numel() * t.element_size() for t in host)

    stream_ctx = torch.cuda.stream(trainer.stream)
    stream_ctx.__enter__()          # every step (eager or captured) runs on the trainer's stream
    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return ms.item()

    def step_resident():
        return trainer.step(*devs)

    loss_host = torch.empty(1).pin_memory()

    # ---- warm-up + kernel table (picks the dominant kernel signature) -----------------------
    for _ in range(max(args.warmup, 3) - 1):
        step_resident()
    def gpu_lag(ms=120.0):
        # CUDA events time [record, record] on the stream: if the GPU is waiting for the host to enqueue the next
        # call, the host's launch gap is charged to the kernel.  Park the stream behind a spin kernel first so that
        # every launch of the profiled step is already queued when the GPU reaches it.
        torch.cuda._sleep(int(ms * 1e-3 * 1.9e9))

    dominant, by_entry, step_ms_profiled = None, {}, 0.0
    if not args.no_roofline:
        table = KernelTimer()
        _cabi.set_profiler(table)
        trainer.overlap = False          # per-kernel event timing: one stream, so a kernel's interval is its own
        gpu_lag()
        step_resident()
        trainer.overlap = True
        _cabi.set_profiler(None)
        totals = table.totals()
        # the dominant kernel = the C-ABI entry point (one kernel template, all layer shapes) with the largest share of
        # the step; no single layer shape carries more than ~5 % of this step
        for (nm, key), (ms, cnt) in totals.items():
            by_entry[nm] = by_entry.get(nm, 0.0) + ms
        dominant = max(by_entry.items(), key=lambda kv: kv[1])[0]
        step_ms_profiled = sum(v[0] for v in totals.values())
        if args.kernel_table and rank == 0:
            by_name = {}
            for (nm, key), (ms, cnt) in totals.items():
                a = by_name.setdefault(nm, [0.0, 0])
                a[0] += ms
                a[1] += cnt
            for nm, (ms, cnt) in sorted(by_name.items(), key=lambda kv: -kv[1][0]):
                print("%-28s %4d launches %9.3f ms" % (nm, cnt, ms), file=sys.stderr)
            for (nm, key), (ms, cnt) in sorted(totals.items(), key=lambda kv: -kv[1][0])[:40]:
                fl, by = kernel_cost(nm, key)
                print("  %-24s %-44s x%d %8.3f ms  %7.1f TF/s %7.1f GB/s" % (
                    nm, key, cnt, ms, fl * cnt / ms / 1e9 if ms else 0, by * cnt / ms / 1e6 if ms else 0), file=sys.stderr)

    calls0 = _cabi.N_CALLS
    step_resident()
    calls_per_step = _cabi.N_CALLS - calls0
    use_graph = not args.no_graph
    if use_graph:
        trainer.capture(*devs, warmup=1)
        run_step = trainer.step_graphed
    else:
        run_step = step_resident

    def step_e2e_serial():
        for d_, h_ in zip(devs, host):
            d_.copy_(h_, non_blocking=True)
        mon = run_step()
        loss_host.copy_(mon["final_loss"].reshape(1), non_blocking=False)      # device->host read of the loss

    # ---- timed: resident inputs ----------------------------------------------------------------
    clocks = ClockSampler(local_rank)
    clocks.start()
    run_step()
    ms_total = timed(run_step, args.steps)
    launches = calls_per_step * args.steps
    # ---- timed: end to end from pinned host buffers -----------------------------------------
    e2e_how = "serial copy -> step -> read-back on one stream"
    if use_graph and not args.no_e2e_prefetch:
        # two static input buffer sets and two captured graphs: while graph(slot) runs, the next batch is copied into
        # the other set on a copy stream.  Every timed step still pays one H2D copy of its inputs and one D2H read of
        # its loss; the copies are merely enqueued one step ahead.
        devs2 = [t.clone() for t in devs]                     # valid data: capture() warms up on them
        trainer.capture(*devs2, warmup=1, slot=1)
        bufs = (devs, devs2)
        copy_stream = torch.cuda.Stream()
        state = {"i": 0, "primed": False, "done": [None, None]}

        def enqueue_copy(slot):
            # the set may only be overwritten once the last step that READ it has finished (its own event), not after
            # everything enqueued so far -- the step that is running now reads the other set
            if state["done"][slot] is not None:
                copy_stream.wait_event(state["done"][slot])
            with torch.cuda.stream(copy_stream):
                for d_, h_ in zip(bufs[slot], host):
                    d_.copy_(h_, non_blocking=True)

        def step_e2e_prefetch():
            slot = state["i"] & 1
            cur = torch.cuda.current_stream()
            if not state["primed"]:
                enqueue_copy(slot)
                state["primed"] = True
            cur.wait_stream(copy_stream)                                # this step's inputs have arrived
            mon = trainer.step_graphed(slot=slot)
            ev = torch.cuda.Event()
            ev.record(cur)
            state["done"][slot] = ev
            enqueue_copy(slot ^ 1)                                      # next step's inputs, overlapping this step
            loss_host.copy_(mon["final_loss"].reshape(1), non_blocking=False)
            state["i"] += 1

        step_e2e_fn = step_e2e_prefetch
        e2e_how = "double-buffered inputs: H2D copy of step i+1 on a copy stream overlaps step i; loss read back every step"
    else:
        step_e2e_fn = step_e2e_serial
    step_e2e_fn()
    ms_e2e = timed(step_e2e_fn, args.steps)

    # ---- joint_ttt: the validation leg (per-case test-time training + inference + binary Dice) -------------------
    ttt = None
    if mode == "joint_ttt":
        torch.manual_seed(2000 + rank)
        ncase = max(args.ttt_cases, 1)
        cases_all = [(synth_image(1, P).to(dev), synth_label(1, P).to(dev)) for _ in range(ncase * world)]
        trainer.validate(cases_all[:world], finetune=finetune, val_finetune=1)          # warm-up (one case per rank)
        barrier()
        t0 = time.perf_counter()
        out = trainer.validate(cases_all, finetune=finetune, val_finetune=1)
        barrier()
        dt = torch.tensor([time.perf_counter() - t0], device=dev)
        if world > 1:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        ttt = {"cases": ncase * world, "cases_per_s": ncase * world / dt.item(), "s_per_case_per_gpu": dt.item() / ncase,
               "val_finetune": 1, "dsc": out["dsc"], "dsc_noft": out["dsc_noft"],
               "timing": "host wall clock around validate() incl. its final host sync, max over ranks (one captured graph replay per case; VAESEG_VAL_GRAPH=0 for eager launches)"}

    # ---- the dominant kernel, CUDA events around each of its launches over the same K steps (eager launches:
    #      events cannot be recorded inside a replayed graph) ---------------------------------------
    dom = {}
    if dominant is not None:
        only = KernelTimer(only_name=dominant)
        _cabi.set_profiler(only)
        trainer.overlap = False
        for _ in range(args.steps):
            gpu_lag(40.0)
            step_resident()
        trainer.overlap = True
        _cabi.set_profiler(None)
        dom = only.totals()
    clock_info = clocks.stop()

    stream_ctx.__exit__(None, None, None)
    global_batch = B * world
    value = global_batch * args.steps / (ms_total / 1e3)
    e2e_value = global_batch * args.steps / (ms_e2e / 1e3)

    peaks = measured_peaks()
    roof = None
    if dom:
        dom_ms = sum(ms for ms, _ in dom.values())
        dom_cnt = sum(cnt for _, cnt in dom.values())
        fl = sum(kernel_cost(nm, key)[0] * cnt for (nm, key), (_, cnt) in dom.items())
        by = sum(kernel_cost(nm, key)[1] * cnt for (nm, key), (_, cnt) in dom.items())
        dur_s = dom_ms / 1e3                       # all launches of the dominant entry point over the K profiled steps
        if by > 0 and fl / by >= RIDGE:
            roof = {"bound": "tensor", "achieved": fl / dur_s / 1e12, "peak": peaks["tensor"], "unit": "TFLOP/s"}
        else:
            roof = {"bound": "hbm", "achieved": by / dur_s / 1e9, "peak": peaks["hbm"], "unit": "GB/s"}
        roof["frac"] = roof["achieved"] / roof["peak"]
        roof["traffic"] = None
        roof["kernel"] = dominant
        roof["launches_timed"] = dom_cnt
        roof["avg_us"] = dom_ms / dom_cnt * 1e3
        roof["algorithmic_bytes_per_launch"] = by / dom_cnt
        roof["algorithmic_flops_per_launch"] = fl / dom_cnt
        roof["share_of_step"] = by_entry[dominant] / step_ms_profiled
        roof["peak_source"] = peaks["source"] + (" sustained" if roof["bound"] == "tensor" else "")
        # the three heaviest layer shapes of that entry point, each against its own bound (decided per layer)
        shapes = []
        for (nm, key), (ms, cnt) in sorted(dom.items(), key=lambda kv: -kv[1][0])[:3]:
            f1, b1 = kernel_cost(nm, key)
            tensor = b1 > 0 and f1 / b1 >= RIDGE
            ach = (f1 * cnt / (ms / 1e3) / 1e12) if tensor else (b1 * cnt / (ms / 1e3) / 1e9)
            shapes.append({"shape": list(key), "launches": cnt, "avg_us": ms / cnt * 1e3, "bound": "tensor" if tensor else "hbm",
                           "achieved": ach, "frac": ach / (peaks["tensor"] if tensor else peaks["hbm"])})
        roof["by_shape"] = shapes
        tr_path = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.isfile(tr_path):
            tr = json.load(open(tr_path)).get(dominant)
            if tr is not None:
                roof["traffic"] = tr.get("dram_bytes_per_launch")
                roof["traffic_note"] = tr.get("note")

    line = {"metric": MODES[mode][1] % P, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_total / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "bf16" if args.precision == "bf16" else "f32",
            "data": "synthetic",
            "config": {"workload": MODES[mode][2], "mode": mode,
                       "patch": P, "per_gpu_batch": B, "global_batch": global_batch, "lambda_vae": 1.0,
                       "loss_type": loss_type, "parallelism": "dp%d" % world,
                       "l2": "per-step working set (>1 GB activations) exceeds the 126 MB L2; no explicit flush",
                       "launch": "eager" if args.no_graph else "cuda-graph (fwd+bwd captured; all-reduce + optimiser eager)",
                       "e2e": e2e_how},
            "clocks": clock_info,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4,
                    "ms_per_step": ms_e2e / args.steps},
            "gpu_launches": launches}
    if roof is not None:
        line["roofline"] = roof
    if ttt is not None:
        line["config"]["ttt"] = ttt
    if args.lean_teacher:
        line["config"]["lean_teacher"] = "teacher VAE forward skipped (its outputs feed only the monitored kl term): NOT the reference's work"

    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        sec = cpu_steps(mode, 2, 1, 1, P)
        line["cpu_baseline"] = {"value": 1.0 / sec, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
                                "sample": "2 timed steps (1 warm-up) of the same %s step at batch 1, %d^3, oracle "
                                          "port of the reference's torch CPU fp32 path" % (mode, P)}
    if stdout_fd is not None:
        sys.stdout.flush()
        os.dup2(stdout_fd, 1)
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        # The captured graph may hold NCCL kernels (the all-reduce buckets overlapped with backward): release it, make
        # sure every rank is done, and leave without the process-group teardown, which does not return after NCCL work
        # has been captured into a CUDA graph on this stack (observed on B200 x2; the JSON line is already flushed).
        if use_graph:
            trainer.release_graph()
        dist.barrier()
        torch.cuda.synchronize()
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)


if __name__ == "__main__":
    main()
