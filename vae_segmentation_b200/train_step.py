"""Train-step harness: the hot-loop bodies of the reference drivers as reusable objects.

  SegTrainer    main_source.py:415-437,660-661   seg_train   (1 - Dice_fg, SGD m=.9)
  VAETrainer    main_source.py:389-406,660-661   vae_train   (1 - Dice_fg + 2e-5 KL, if_random, scale .35)
  JointTrainer  main_target.py:508-592,734-736   domain_adaptation teacher-student step incl.
                EMA teacher (:508-518), dynamic lambda type 8 (:550-560), KL option, confident
                pseudo labels, and the test-time-training inner loop (:807-900)

The reference's module-level script code is restated as methods; the arithmetic is the
drop-in modules + fused losses.  Differences that do not change results: parameters and
gradients live in flat fp32 arenas so the optimiser / EMA / gradient all-reduce are ONE
kernel / collective each; pseudo labels and one-hot targets are never materialised (the
Dice reduction thresholds / compares on the fly); the dynamic-lambda thresholds are
evaluated on the device (no .item() sync).  Data parallelism = one process per
GPU (torch.distributed, NCCL): per-rank batch shards, one gradient all-reduce per step and,
for type 8, a 1-float all-reduce of the recon Dice so every rank takes the same branch.
"""
import os

import torch
import torch.distributed as dist

from . import evaluation as ev
from . import ops


class FlatArena(object):
    """Re-homes a module's parameters (and their .grad) into contiguous fp32 arenas."""

    def __init__(self, module, with_grad=True):
        self.module = module
        self.params = [p for p in module.parameters()]
        n = sum(p.numel() for p in self.params)
        dev = self.params[0].device
        if dev.type != "cuda":
            raise RuntimeError("FlatArena: move the module to CUDA first")
        self.data = torch.empty(n, device=dev, dtype=torch.float32)
        self.grad = torch.zeros(n, device=dev, dtype=torch.float32) if with_grad else None
        off = 0
        for p in self.params:
            k = p.numel()
            self.data[off:off + k].copy_(p.data.reshape(-1))
            p.data = self.data[off:off + k].view(p.shape)
            if with_grad and p.requires_grad:
                p.grad = self.grad[off:off + k].view(p.shape)
            off += k
        self.numel = n
        for m in module.modules():                       # parameters were re-homed: forget every derived pack
            if hasattr(m, "_cache"):
                m._cache.drop()

    def zero_grad(self):
        if self.grad is not None:
            self.grad.zero_()


class FusedSGD(object):
    """torch.optim.SGD(lr, momentum, dampening 0, wd 0) over a FlatArena in one launch."""

    def __init__(self, arena, lr=1e-2, momentum=0.9):
        self.arena, self.lr, self.momentum = arena, lr, momentum
        self.buf = torch.zeros_like(arena.data) if momentum != 0 else None
        self.steps = 0

    def step(self, gscale=1.0):
        ops.sgd_step(self.arena.data, self.arena.grad, self.buf, self.lr, self.momentum, first=(self.steps == 0),
                     gscale=gscale)
        self.steps += 1
        self.arena.module.repack_packs()


class FusedAdam(object):
    def __init__(self, arena, lr=1e-3, betas=(0.9, 0.999), eps=1e-8):
        self.arena, self.lr, self.betas, self.eps = arena, lr, betas, eps
        self.m = torch.zeros_like(arena.data)
        self.v = torch.zeros_like(arena.data)
        self.steps = 0

    def step(self, gscale=1.0):
        self.steps += 1
        ops.adam_step(self.arena.data, self.arena.grad, self.m, self.v, self.lr, self.betas[0], self.betas[1],
                      self.eps, self.steps, gscale=gscale)
        self.arena.module.repack_packs()


def _world():
    return dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1


def allreduce_mean_(flat_grad):
    """Sum over ranks in place; returns the scale (1/world) the optimiser kernel applies, so
    the averaging costs no extra pass.  Matches DataParallel's loss-on-gathered-batch mean
    for equal per-rank batches."""
    world = _world()
    if world > 1:
        dist.all_reduce(flat_grad, op=dist.ReduceOp.SUM)
    return 1.0 / world


class BucketedAllReduce(object):
    """Gradient all-reduce overlapped with backward: the flat gradient arena is cut into `nbuckets`
    contiguous buckets on parameter boundaries.  Backward produces gradients from the LAST parameter towards the
    first, so the ready region grows downwards from the end of the arena; as soon as it covers a bucket that bucket's
    NCCL all-reduce (sum; the 1/world scale is folded into the optimiser kernel) is issued behind the streams that
    hold the producing kernels, while the remaining backward pass keeps running.  `finish()` issues whatever is left
    and makes the current stream wait for every bucket.  Capturable in a CUDA graph (NCCL collectives are)."""

    def __init__(self, arena, nbuckets=3):
        self.arena = arena
        self.offsets = {}
        off = 0
        for p in arena.params:
            self.offsets[id(p)] = (off, p.numel())
            off += p.numel()
        # bucket boundaries snapped to parameter starts, roughly equal in elements
        starts = sorted(o for o, _ in self.offsets.values())
        cuts = [0]
        for k in range(1, nbuckets):
            target = arena.numel * k // nbuckets
            cut = min(starts, key=lambda o: abs(o - target))
            if cut > cuts[-1]:
                cuts.append(cut)
        self.bounds = list(zip(cuts, cuts[1:] + [arena.numel]))          # ascending [lo, hi)
        self.gate = None
        self.reset()

    def reset(self):
        self.not_ready = dict(self.offsets)
        self.next = len(self.bounds) - 1                                   # buckets are launched from the last one down
        self.works = []

    def _launch(self, k, streams):
        lo, hi = self.bounds[k]
        if not self.arena.grad.is_cuda:          # host-logic tests (gloo): no streams to order
            self.works.append(dist.all_reduce(self.arena.grad[lo:hi], op=dist.ReduceOp.SUM, async_op=True))
            return
        if self.gate is None:
            self.gate = torch.cuda.Stream()
        # the gradient kernels of the bucket were enqueued on the current stream and on the weight-gradient streams:
        # a gate stream waits for those, and NCCL's stream is ordered behind the gate -- the backward chain itself
        # never waits for a weight gradient
        self.gate.wait_stream(torch.cuda.current_stream())
        for ws in streams:
            self.gate.wait_stream(ws)
        with torch.cuda.stream(self.gate):
            self.works.append(dist.all_reduce(self.arena.grad[lo:hi], op=dist.ReduceOp.SUM, async_op=True))

    def on_ready(self, p):
        self.not_ready.pop(id(p), None)
        frontier = max((o + n for o, n in self.not_ready.values()), default=0)
        from . import engine
        while self.next >= 0 and self.bounds[self.next][0] >= frontier:
            self._launch(self.next, engine._wgrad_streams())
            self.next -= 1

    def finish(self):
        from . import engine
        while self.next >= 0:
            self._launch(self.next, engine._wgrad_streams())
            self.next -= 1
        for w in self.works:
            w.wait()
        self.works = []


def sync_first_term_(terms):
    """Dynamic lambda (type 8, main_target.py:550-560) thresholds recon_loss on the host of the single
    DataParallel process, i.e. on the GLOBAL batch.  Under process-per-GPU every rank must take the same
    branch: replace terms[0] by its mean over ranks (equal per-rank batches), one 1-float all-reduce."""
    world = _world()
    if world == 1:
        return terms
    g = terms[0:1].clone()
    dist.all_reduce(g, op=dist.ReduceOp.SUM)
    return torch.cat([g / world, terms[1:]])


class _GraphedTrainer(object):
    """Shared plumbing of the single-network trainers: side stream for the weight gradients, CUDA-graph capture of
    zero_grad + forward + backward on static input buffers, eager all-reduce + fused optimiser after the replay."""

    def _init_streams(self):
        hp = -1 if os.environ.get("VAESEG_STREAM_PRIORITY", "1") == "1" else 0
        self.stream = torch.cuda.Stream(priority=hp)
        self.wgrad_stream = [torch.cuda.Stream()]
        self.overlap = True
        self._graphs = {}

    def _backward(self, loss):
        from . import engine
        engine.WGRAD_STREAM = self.wgrad_stream if self.overlap else None
        try:
            loss.backward()
            engine.join_wgrad_stream()
        finally:
            engine.WGRAD_STREAM = None

    def capture(self, *static_inputs, warmup=2, slot=0):
        """Same contract as JointTrainer.capture: run every step of this trainer under
        `with torch.cuda.stream(trainer.stream)` if it will be captured."""
        side = self.stream
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(max(warmup, 1)):          # forward + backward only: no optimiser state is touched
                self.forward_backward(*static_inputs)
            self.arena.zero_grad()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph, stream=side):
            mon = self.forward_backward(*static_inputs)
        self._graphs[slot] = (graph, mon)
        return self

    def step_graphed(self, slot=0):
        graph, mon = self._graphs[slot]
        graph.replay()
        self.opt.step(allreduce_mean_(self.arena.grad))
        return mon

    def release_graph(self):
        self._graphs = {}
        torch.cuda.synchronize()


class SegTrainer(_GraphedTrainer):
    def __init__(self, seg, lr=1e-2, momentum=0.9, eps=0.0001):
        self.seg = seg
        self.arena = FlatArena(seg)
        self.opt = FusedSGD(self.arena, lr, momentum)
        self.eps = eps                       # main_source.py:150-182 uses 1e-4
        self._init_streams()

    def loss(self, img, label):
        pred = self.seg.predict(img)
        return 1 - ev.avg_dsc_fused(pred, label, "label", botindex=1, topindex=2, eps=self.eps), pred

    def forward_backward(self, img, label):
        self.arena.zero_grad()
        loss, _ = self.loss(img, label)
        self._backward(loss)
        return {"dice_loss": loss.detach(), "final_loss": loss.detach()}

    def step(self, img, label):
        mon = self.forward_backward(img, label)
        self.opt.step(allreduce_mean_(self.arena.grad))
        return mon


class VAETrainer(_GraphedTrainer):
    def __init__(self, vae, lr=1e-2, momentum=0.9, scale=0.35, eps=0.0001):
        self.vae = vae
        self.arena = FlatArena(vae)
        self.opt = FusedSGD(self.arena, lr, momentum)
        self.scale, self.eps = scale, eps
        self._init_streams()

    def loss(self, label, z=None):
        onehot = ev.one_hot(label, 2)
        recon, mean, std = self.vae(onehot, if_random=True, scale=self.scale, z=z)
        kl = ev.KLloss({"mean": mean, "std": std})
        dsc = 1 - ev.avg_dsc_fused(recon, label, "label", botindex=1, topindex=2, eps=self.eps)
        return dsc + 0.00002 * kl, dsc, kl, recon

    def forward_backward(self, label, z=None):
        """z: the reparameterisation noise; pass a static CUDA buffer when the step is captured (the reference draws it
        from the CPU generator every call, joint_model.py:246 -- the caller then refreshes the buffer per step)."""
        self.arena.zero_grad()
        loss, dsc, kl, _ = self.loss(label, z)
        self._backward(loss)
        return {"final_loss": loss.detach(), "dice_loss": dsc.detach(), "kl_loss": kl.detach()}

    def step(self, label, z=None):
        mon = self.forward_backward(label, z)
        self.opt.step(allreduce_mean_(self.arena.grad))
        return mon


class JointTrainer(object):
    """Teacher-student target-domain step.  `student` / `teacher` are Joint modules sharing
    the frozen VAE architecture; only student.Seg trains (main_target.py:396-433)."""

    def __init__(self, student, teacher, lr=1e-2, momentum=0.9, lambda_vae=1.0, loss_type=0, kl=False,
                 confident=False, only_pseudo=False, alpha=0.995, adam=False, faithful_teacher=True, overlap=True):
        self.student, self.teacher = student, teacher
        for p in student.Vae.parameters():
            p.requires_grad = False
        for p in teacher.parameters():
            p.requires_grad = False
        self.arena = FlatArena(student.Seg)
        self.teacher_arena = FlatArena(teacher.Seg, with_grad=False)
        self.opt = FusedAdam(self.arena, lr) if adam else FusedSGD(self.arena, lr, momentum)
        self.lambda_vae, self.loss_type, self.kl = lambda_vae, loss_type, kl
        self.confident, self.only_pseudo, self.alpha = confident, only_pseudo, alpha
        self.faithful_teacher = faithful_teacher
        self.fused_loss = os.environ.get("VAESEG_NO_FUSED_LOSS", "0") != "1"     # ev.joint_target_loss (one autograd node)
        # Per-sample chains: every op of the step is per-sample (InstanceNorm per (n,c), Dice per sample then batch
        # mean), so the B samples of the per-GPU batch run as B independent forward/backward chains on their own
        # streams (parallel branches of the captured graph).  Half of a chain's time is launch-latency-bound deep-level
        # kernels on a handful of SMs: two chains overlap there, and serialise only on the full-resolution kernels
        # that fill the GPU anyway.  Weight gradients of all chains accumulate onto the same .grad with atomics.
        # Not used with the dynamic lambda (type 8 thresholds the BATCH reconstruction loss before composing).
        # MEASURED (B200, 2 x 96^3): 279 vol/s with chains vs 313 without -- the full-resolution kernels are persistent
        # 148-CTA grids that serialise across chains while their fixed cost doubles -- so it is opt-in
        # (VAESEG_SAMPLE_PARALLEL=1); parity-tested (tests/test_models_gpu.py runs the joint step both ways).
        self.sample_parallel = os.environ.get("VAESEG_SAMPLE_PARALLEL", "0") == "1"
        self._chain_streams = {}
        # data parallel.  The 9.1 MB gradient arena is all-reduced in `ddp_buckets` buckets, each issued as soon as its
        # parameters' gradients are complete and overlapped with the rest of backward (BucketedAllReduce; captured
        # inside the CUDA graph of the step).  VAESEG_DDP_OVERLAP=0 (or trainer.ddp_overlap = False before the first
        # step) selects ONE all-reduce after backward.  Round 1, 2 x B200: no difference outside the noise (657.6 / 648.6
        # plain vs 648.7 overlapped: the payload costs ~40 us on NVLink, < 1 % of the step).
        # NOTE for callers: destroy the captured graph (trainer.release_graph()) before tearing the process group down --
        # destroy_process_group() does not return while a live CUDA graph still holds NCCL kernels (observed, B200 x2).
        self.ddp_buckets = int(os.environ.get("VAESEG_DDP_BUCKETS", "3"))
        # Round 2, 8 x B200: 2894 vol/s overlapped vs 2852 plain (2 GPUs: 741 vs 729) -- the overlapped path is the default
        # under data parallelism; VAESEG_DDP_OVERLAP=0 selects the single all-reduce after backward.
        self.ddp_overlap = os.environ.get("VAESEG_DDP_OVERLAP", "1") == "1"
        self._bucketer = None
        self._grads_reduced = False
        self._graphs = {}                        # slot -> (captured CUDA graph, monitored terms), see capture()
        self._val_graphs = {}                    # validation: (shape, TTT setting) -> captured per-case graph
        self._graph = self._graph_mon = None
        # the step's own (critical) chain runs at high priority; the teacher forward and the weight gradients, which
        # only have to finish by the loss / the optimiser step, fill in behind it at the default priority
        hp = -1 if os.environ.get("VAESEG_STREAM_PRIORITY", "1") == "1" else 0
        self.stream = torch.cuda.Stream(priority=hp)        # see capture()
        # Concurrency inside the step (parallel branches of the captured graph): the frozen teacher's forward is
        # independent of the student's until the losses, and the weight-gradient kernels are leaves of the backward
        # chain.  Both are dominated by deep-level kernels that occupy 8-48 of the 148 SMs, so they overlap well.
        self.overlap = overlap
        self.teacher_stream = torch.cuda.Stream()
        self.wgrad_stream = [torch.cuda.Stream() for _ in range(max(1, int(os.environ.get("VAESEG_WGRAD_STREAMS", "1"))))]

    def ema_teacher(self):
        # main_target.py:512-516 on the Seg state_dict: parameters through the flat arenas, buffers (the running
        # statistics of a norm_type=2 model; none with InstanceNorm) one by one
        ops.ema_update(self.teacher_arena.data, self.arena.data, self.alpha)
        with torch.no_grad():
            for tb, sb in zip(self.teacher.Seg.buffers(), self.student.Seg.buffers()):
                if tb.is_floating_point():
                    tb.mul_(self.alpha).add_(sb, alpha=1.0 - self.alpha)
                else:                                   # num_batches_tracked: the reference's float result is cast back on load
                    tb.copy_((self.alpha * tb.double() + (1.0 - self.alpha) * sb.double()).to(tb.dtype))
        self.teacher.repack_packs()

    def losses(self, img, label, student=None, sync=True):
        """Forward of one step; returns (final_loss, dict of monitored terms).  sync=False: never issue a collective
        (test-time training / validation: cases are sharded over ranks and the reference thresholds the dynamic lambda
        on each case's OWN recon loss, main_target.py:838-847)."""
        student = student or self.student
        cur = torch.cuda.current_stream()

        def run_teacher():
            with torch.no_grad():                                                     # :532 (frozen teacher)
                if self.faithful_teacher or self.kl:
                    return self.teacher({"img": img}, "img", "only_fake", "unused_recon")
                out = self.teacher.Seg({"img": img}, "img", "only_fake")
                out["mean"] = out["std"] = None
                return out

        if self.overlap:
            self.teacher_stream.wait_stream(cur)
            with torch.cuda.stream(self.teacher_stream):
                tb = run_teacher()
        batch = {"img": img}
        batch = student(batch, "img", "pred", "recon_pred", dropout=True)            # main_target.py:531
        if self.overlap:
            cur.wait_stream(self.teacher_stream)
            for v in tb.values():
                if torch.is_tensor(v) and v.is_cuda:
                    v.record_stream(cur)
        else:
            tb = run_teacher()
        pred = batch["pred"]
        if self.fused_loss and not (sync and self.loss_type == 8 and _world() > 1):
            klv = ev.KLloss(tb) if tb.get("mean") is not None else None                                   # :544 (teacher's, F8)
            final, m5 = ev.joint_target_loss(pred, batch["recon_pred"], label, tb["only_fake"], kl=klv,
                                             lambda_vae=self.lambda_vae, loss_type=self.loss_type, use_kl=self.kl,
                                             only_pseudo=self.only_pseudo, confident=self.confident)
            mon = {"final_loss": m5[0], "recon_loss": m5[1], "dice_loss": m5[2], "dice_loss_fake": m5[3], "kl_loss": m5[4]}
            return final, mon, batch
        recon_loss = 1 - ev.avg_dsc_fused(pred, batch["recon_pred"], "tensor", botindex=1, topindex=2)      # :543
        dsc_loss = 1 - ev.avg_dsc_fused(pred.detach(), label, "label", botindex=1, topindex=2)             # :545 (monitor)
        dsc_loss_fake = 1 - ev.avg_dsc_fused(pred, tb["only_fake"], "confident" if self.confident else "binarize",
                                             botindex=1, topindex=2)                                        # :534-537,546
        klloss = ev.KLloss(tb) if tb.get("mean") is not None else torch.zeros((), device=img.device)       # :544 (teacher's, F8)
        if self.only_pseudo:
            final = dsc_loss_fake
        else:
            terms = torch.stack([recon_loss.detach(), dsc_loss_fake.detach(), klloss.detach()])
            if self.loss_type == 8 and sync:
                terms = sync_first_term_(terms)
            _, wts = ops.compose_target_loss(terms, self.lambda_vae, self.loss_type, self.kl)
            final = wts[0] * recon_loss + wts[1] * dsc_loss_fake + (wts[2] * klloss if self.kl else 0)
        mon = {"final_loss": final.detach(), "recon_loss": recon_loss.detach(), "dice_loss": dsc_loss.detach(),
               "dice_loss_fake": dsc_loss_fake.detach(), "kl_loss": klloss.detach()}
        return final, mon, batch

    def _use_chains(self, img):
        return self.sample_parallel and self.overlap and self.fused_loss and self.loss_type != 8 and img.shape[0] > 1

    def _chains(self, b):
        st = self._chain_streams.get(b)
        if st is None:
            st = self._chain_streams[b] = (torch.cuda.Stream(), torch.cuda.Stream(), torch.cuda.Stream())
        return st            # (student chain, teacher, weight gradients) of sample b

    def forward_backward(self, img, label):
        """zero_grad + forward + backward of one step on the current stream; returns the monitored terms."""
        from . import engine
        self.arena.zero_grad()
        self._grads_reduced = False
        if _world() > 1 and self.ddp_overlap and not self._use_chains(img):
            if self._bucketer is None:
                self._bucketer = BucketedAllReduce(self.arena, self.ddp_buckets)
            self._bucketer.reset()
            final, mon, _ = self.losses(img, label)
            engine.GRAD_READY_HOOK = self._bucketer.on_ready
            try:
                self._backward(final, join=False)
                self._bucketer.finish()                    # remaining buckets + wait for all of them
                engine.join_wgrad_stream()
            finally:
                engine.GRAD_READY_HOOK = None
                engine.WGRAD_STREAM = None
            self._grads_reduced = True
            return mon
        if not self._use_chains(img):
            final, mon, _ = self.losses(img, label)
            self._backward(final)
            return mon
        cur = torch.cuda.current_stream()
        B = img.shape[0]
        mons = []
        saved_teacher, saved_overlap = self.teacher_stream, self.overlap
        try:
            for b in range(B):
                s_main, s_teacher, s_wgrad = self._chains(b)
                s_main.wait_stream(cur)
                self.teacher_stream = s_teacher
                with torch.cuda.stream(s_main):
                    final, mon, _ = self.losses(img[b:b + 1], label[b:b + 1])
                    engine.WGRAD_STREAM = s_wgrad
                    try:
                        (final / B).backward()
                        engine.join_wgrad_stream()
                    finally:
                        engine.WGRAD_STREAM = None
                mons.append(mon)
            for b in range(B):
                cur.wait_stream(self._chains(b)[0])
        finally:
            self.teacher_stream, self.overlap = saved_teacher, saved_overlap
        out = {}
        for k in mons[0]:
            vals = torch.stack([m[k].reshape(()) for m in mons])
            for v in (m[k] for m in mons):
                if torch.is_tensor(v) and v.is_cuda:
                    v.record_stream(cur)
            out[k] = vals.mean()
        return out

    def step(self, img, label, update_teacher=False):
        if update_teacher:
            self.ema_teacher()
        mon = self.forward_backward(img, label)
        self.opt.step(self._grad_scale())
        return mon

    def _grad_scale(self):
        """1/world for the optimiser kernel; all-reduces the gradient arena here unless backward already did."""
        if self._grads_reduced:
            return 1.0 / _world()
        return allreduce_mean_(self.arena.grad)

    def _backward(self, final, join=True):
        from . import engine
        engine.WGRAD_STREAM = self.wgrad_stream if self.overlap else None
        if not join:                 # the caller joins (and resets WGRAD_STREAM) after it has used the side streams
            final.backward()
            return
        try:
            final.backward()
            engine.join_wgrad_stream()
        finally:
            engine.WGRAD_STREAM = None

    def capture(self, img_static, label_static, warmup=2, slot=0):
        """CUDA-graph capture of zero_grad + forward (student, teacher, losses) + backward on the given static
        input buffers; the gradient all-reduce and the fused optimiser kernel stay eager (two launches).  At
        ~600 kernel launches per step the Python/launch path, not the GPU, bounds an eager step.
        Autograd pins each parameter's gradient-accumulation node to the stream of its first backward, and a
        capture may not depend on the legacy default stream: run EVERY step of this trainer under
        `with torch.cuda.stream(trainer.stream)` if it will be captured later.
        `slot`: several graphs over different static input buffers may coexist (e.g. two, so that the host->device copy
        of the next batch overlaps the current step: bench.py --e2e-prefetch); `step_graphed(slot=...)` replays one."""
        side = self.stream
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):            # autograd's stream bookkeeping must see the capture stream
            # forward + backward only: warming up must not take optimiser steps (weights, momentum and the Adam step
            # counter stay exactly where the training schedule left them); also runs every lazy one-time init
            for _ in range(max(warmup, 1)):
                self.forward_backward(img_static, label_static)
            self.arena.zero_grad()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph, stream=side):
            mon = self.forward_backward(img_static, label_static)
        if not hasattr(self, "_graphs"):
            self._graphs = {}
        self._graphs[slot] = (graph, mon)
        if slot == 0:
            self._graph, self._graph_mon = graph, mon
        return self

    def release_graph(self):
        """Drops the captured graph (and the NCCL kernels it holds)."""
        self._graph = None
        self._graph_mon = None
        self._graphs = {}
        self._val_graphs = {}
        torch.cuda.synchronize()

    def step_graphed(self, update_teacher=False, slot=0):
        if update_teacher:
            self.ema_teacher()
        graph, mon = self._graphs[slot]
        graph.replay()
        self.opt.step(self._grad_scale())
        return mon

    def test_time_train(self, finetune, img, label, iters=1, lr_finetune=1e-2):
        """main_target.py:807-900: per validation case, `finetune` (a Joint) starts from the
        student's weights and takes `iters` plain-SGD (momentum 0, fresh optimiser) steps on
        the same loss; cases are independent, so under data parallelism they are simply
        distributed over ranks with no collective.  Returns (student pred, finetuned pred)."""
        ft = getattr(self, "_ft_arena", None)
        if ft is None or ft.module is not finetune.Seg:
            for p in finetune.Vae.parameters():
                p.requires_grad = False
            ft = self._ft_arena = FlatArena(finetune.Seg)
        ft.data.copy_(self.arena.data)                                                # :810 load_state_dict (Seg part)
        with torch.no_grad():                                                         # ... and the Vae part: the whole
            for dst, src in zip(finetune.Vae.parameters(), self.student.Vae.parameters()):   # Joint state is loaded
                if dst.data_ptr() != src.data_ptr():
                    dst.copy_(src)
            for dst, src in zip(finetune.buffers(), self.student.buffers()):         # BatchNorm running statistics, if any
                dst.copy_(src)
        finetune.repack_packs()
        for _ in range(iters):
            ft.zero_grad()
            final, _, _ = self.losses(img, label, student=finetune, sync=False)      # per-case: no collective
            final.backward()
            ops.sgd_step(ft.data, ft.grad, None, lr_finetune, 0.0, first=True)       # :886 SGD(lr_finetune, momentum 0)
            finetune.repack_packs()
        with torch.no_grad():                                                         # :902-914
            p0 = self.student.Seg.predict(img)
            p1 = finetune.Seg.predict(img)
        return p0, p1

    def _case_scores(self, finetune, img, label, val_finetune, lr_finetune):
        """One validation case: optional TTT, inference, (finetuned, student) binary Dice as a 2-vector on the device."""
        if val_finetune and finetune is not None:
            p0, p1 = self.test_time_train(finetune, img, label, iters=val_finetune, lr_finetune=lr_finetune)
        else:
            with torch.no_grad():
                p0 = self.student.Seg.predict(img)
            p1 = p0
        onehot = ev.one_hot(label, p0.shape[1])
        top = p0.shape[1]
        s0 = ev.avg_dsc({"p": p0, "t": onehot}, "p", "t", binary=True, botindex=1, topindex=top).reshape(())
        s1 = s0 if p1 is p0 else ev.avg_dsc({"p": p1, "t": onehot}, "p", "t", binary=True, botindex=1, topindex=top).reshape(())
        return torch.stack([s1, s0])

    def _case_graph(self, finetune, img, label, val_finetune, lr_finetune):
        """The per-case validation work captured ONCE per (case shape, TTT setting) on static input buffers: every case
        of a validation pass has the same patch shape, and a case is ~600 launches (weight copy, teacher + finetune
        forward, backward, SGD, re-pack, two inferences, Dice) that the host cannot issue as fast as the GPU runs them
        (eager: 25 cases/s at 128^3).  Replay: copy the case into the static buffers, launch the graph."""
        key = (id(finetune), tuple(img.shape), str(img.dtype), int(val_finetune), float(lr_finetune), self.loss_type,
               float(self.lambda_vae) if not torch.is_tensor(self.lambda_vae) else -1.0)
        ent = self._val_graphs.get(key)
        if ent is None:
            s_img, s_label = img.clone(), label.clone()
            cur = torch.cuda.current_stream()
            side = self.stream
            side.wait_stream(cur)
            with torch.cuda.stream(side):
                for _ in range(2):                         # allocations, pack-cache entries, the finetune arena
                    self._case_scores(finetune, s_img, s_label, val_finetune, lr_finetune)
            cur.wait_stream(side)
            torch.cuda.synchronize()
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph, stream=side):
                out = self._case_scores(finetune, s_img, s_label, val_finetune, lr_finetune)
            self._val_graphs = {key: (graph, s_img, s_label, out)}      # one at a time: a graph pins GBs of activations
            ent = self._val_graphs[key]
        return ent

    def validate(self, cases, finetune=None, val_finetune=0, lr_finetune=1e-2, graphed=None):
        """The validation pass of main_target.py:795-960 (method 'domain_adaptation') over `cases`, an iterable of
        (img [1,1,D,H,W], label [1,1,D,H,W]) CUDA tensors: per case optionally `val_finetune` test-time-training
        iterations on a private copy of the student (:807-900), then no-grad inference and the binary (argmax) Dice of
        the foreground class against the ground truth, `avg_dsc(binary=True, botindex=1, topindex=n_class)` (:941-945),
        for the student ("noft") and the finetuned model.  Returns {'dsc', 'dsc_noft', 'scores', 'scores_noft'} with
        dsc = mean over cases (the reference's `dsc_pancreas / len(val_loader)`).  Scores stay on the device until the
        end (one host sync per pass instead of one `.item()` per case); under data parallelism every rank takes
        cases[rank::world] and the two sums are all-reduced -- cases are independent, nothing else is exchanged."""
        world = _world()
        rank = dist.get_rank() if world > 1 else 0
        if graphed is None:
            graphed = os.environ.get("VAESEG_VAL_GRAPH", "1") == "1"
        scores, scores_noft = [], []
        for idx, (img, label) in enumerate(cases):
            if idx % world != rank:
                continue
            if graphed:
                graph, s_img, s_label, out = self._case_graph(finetune, img, label, val_finetune, lr_finetune)
                s_img.copy_(img, non_blocking=True)
                s_label.copy_(label, non_blocking=True)
                graph.replay()
                both = out.clone()
            else:
                both = self._case_scores(finetune, img, label, val_finetune, lr_finetune)
            scores.append(both[0])
            scores_noft.append(both[1])
        dev = self.arena.data.device
        local = torch.stack([torch.stack(scores).sum() if scores else torch.zeros((), device=dev),
                             torch.stack(scores_noft).sum() if scores_noft else torch.zeros((), device=dev),
                             torch.tensor(float(len(scores)), device=dev)])
        if world > 1:
            dist.all_reduce(local, op=dist.ReduceOp.SUM)
        total = local.cpu()
        n = max(int(total[2].item()), 1)
        return {"dsc": total[0].item() / n, "dsc_noft": total[1].item() / n,
                "scores": [s.item() for s in scores], "scores_noft": [s.item() for s in scores_noft]}
