"""Data-parallel host logic on CPU: world_size 2, gloo backend.

The CUDA kernels cannot run here, so the per-rank gradients come from the oracle; what is under
test is the PRODUCT's collective logic (`train_step.allreduce_mean_`, `train_step.sync_first_term_`):
sum over ranks + 1/world scale must equal the reference's DataParallel semantics, i.e. the gradient
of the loss on the gathered (concatenated) batch (main_target.py:436-438,734-736)."""
import os
import socket
import tempfile

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import ref_torch as R
from vae_segmentation_b200 import train_step as ts
from vae_segmentation_b200.synthetic import synth_image, synth_label

PATCH = 32


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, outdir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.set_num_threads(2)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.manual_seed(7)                              # identical replicas on every rank
        sd = R.init_seg_state()
        torch.manual_seed(100)                            # the global batch, identical on every rank ...
        img, label = synth_image(world, PATCH), synth_label(world, PATCH)
        loss, grads, _ = R.seg_train_step(sd, img[rank:rank + 1], label[rank:rank + 1], eps=0.0001,
                                          dtype=torch.float64)                            # ... sharded (fp64:
        # fp32 gradients of this net carry ~1e-2 conditioning noise)
        flat = torch.cat([g.reshape(-1) for g in grads.values()])
        scale = ts.allreduce_mean_(flat)                  # product code: SUM all-reduce, returns 1/world
        terms = torch.tensor([0.1 + 0.2 * rank, 1.0 + rank, 2.0 + rank], dtype=torch.float64)
        synced = ts.sync_first_term_(terms)
        torch.save({"flat": flat * scale, "scale": scale, "terms": synced, "loss": loss.detach()},
                   os.path.join(outdir, "rank%d.pt" % rank))
    finally:
        dist.destroy_process_group()


def test_world2_gradient_allreduce_equals_gathered_batch():
    world = 2
    with tempfile.TemporaryDirectory() as outdir:
        mp.spawn(_worker, args=(world, _free_port(), outdir), nprocs=world, join=True)
        outs = [torch.load(os.path.join(outdir, "rank%d.pt" % r)) for r in range(world)]
    assert all(o["scale"] == 0.5 for o in outs)
    assert torch.equal(outs[0]["flat"], outs[1]["flat"])                    # every rank holds the same averaged gradient
    # dynamic lambda: terms[0] replaced by the global mean on every rank, the rest untouched
    assert torch.allclose(outs[0]["terms"], torch.tensor([0.2, 1.0, 2.0], dtype=torch.float64))
    assert torch.allclose(outs[1]["terms"], torch.tensor([0.2, 2.0, 3.0], dtype=torch.float64))
    # single-process result on the gathered batch: Dice is per-sample then batch-mean, InstanceNorm is per-(n,c),
    # so mean-of-shard-gradients == gradient of the gathered-batch loss
    torch.manual_seed(7)
    sd = R.init_seg_state()
    torch.manual_seed(100)
    img, label = synth_image(world, PATCH), synth_label(world, PATCH)
    loss, grads, _ = R.seg_train_step(sd, img, label, eps=0.0001, dtype=torch.float64)
    want = torch.cat([g.reshape(-1) for g in grads.values()])
    got = outs[0]["flat"]
    assert ((got - want).norm() / want.norm()).item() < 1e-9
    assert abs(0.5 * (outs[0]["loss"] + outs[1]["loss"]).item() - loss.item()) < 1e-9


def test_single_process_is_identity():
    g = torch.arange(6.0)
    assert ts.allreduce_mean_(g) == 1.0 and torch.equal(g, torch.arange(6.0))
    t = torch.tensor([0.3, 1.0, 2.0])
    assert ts.sync_first_term_(t) is t


class _FakeArena(object):
    """What BucketedAllReduce needs of a FlatArena: params (offsets by order), numel, grad."""

    def __init__(self, sizes, rank):
        self.params = [torch.nn.Parameter(torch.zeros(n)) for n in sizes]
        self.numel = sum(sizes)
        self.grad = torch.arange(self.numel, dtype=torch.float64) * (rank + 1)


def _bucket_worker(rank, world, port, outdir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        sizes = [5, 3, 40, 8, 64, 16, 2]
        arena = _FakeArena(sizes, rank)
        b = ts.BucketedAllReduce(arena, nbuckets=3)
        launched = []
        for p in reversed(arena.params):                 # backward order: last parameter first
            b.on_ready(p)
            launched.append(len(b.works))
        b.finish()
        torch.save({"grad": arena.grad, "bounds": b.bounds, "launched": launched}, os.path.join(outdir, "b%d.pt" % rank))
    finally:
        dist.destroy_process_group()


def test_bucketed_allreduce_covers_every_element_once():
    """train_step.BucketedAllReduce (gradient all-reduce overlapped with backward): buckets tile the arena, are
    launched as soon as the ready region (growing down from the last parameter) covers them, and the result equals
    one SUM all-reduce of the whole arena."""
    world = 2
    with tempfile.TemporaryDirectory() as outdir:
        mp.spawn(_bucket_worker, args=(world, _free_port(), outdir), nprocs=world, join=True)
        outs = [torch.load(os.path.join(outdir, "b%d.pt" % r)) for r in range(world)]
    n = 138
    want = torch.arange(n, dtype=torch.float64) * 3            # rank0 (x1) + rank1 (x2): every element reduced exactly once
    assert torch.equal(outs[0]["grad"], want) and torch.equal(outs[1]["grad"], want)
    bounds = outs[0]["bounds"]
    assert bounds[0][0] == 0 and bounds[-1][1] == n and all(a[1] == b[0] for a, b in zip(bounds, bounds[1:]))
    assert len(bounds) >= 2
    # the last bucket went out before the first parameters were ready (overlap), everything by the end
    assert outs[0]["launched"][0] <= outs[0]["launched"][-1] and outs[0]["launched"][-2] >= 1


def _validate_worker(rank, world, port, outdir):
    """JointTrainer.validate's host logic (sharding cases[rank::world], ONE all-reduce of the two sums and the count at
    the end, nothing per case) with the per-case GPU work stubbed out: an odd number of cases over two ranks."""
    import types
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        tr = object.__new__(ts.JointTrainer)                       # no CUDA modules: only validate() itself is under test
        tr.arena = types.SimpleNamespace(data=torch.zeros(1))
        tr._val_graphs = {}
        calls = []

        def fake_case_scores(finetune, img, label, val_finetune, lr_finetune):
            calls.append(float(img))
            return torch.stack([img * 0.5, img * 0.25])             # (finetuned, student) "scores" of case `img`
        tr._case_scores = fake_case_scores
        cases = [(torch.tensor(float(i + 1)), torch.zeros(())) for i in range(3)]
        out = tr.validate(cases, finetune=None, val_finetune=1, graphed=False)
        torch.save({"out": out, "calls": calls}, os.path.join(outdir, "val%d.pt" % rank))
    finally:
        dist.destroy_process_group()


def test_world2_validation_shards_an_odd_case_count_without_hanging():
    world = 2
    with tempfile.TemporaryDirectory() as outdir:
        mp.spawn(_validate_worker, args=(world, _free_port(), outdir), nprocs=world, join=True)
        outs = [torch.load(os.path.join(outdir, "val%d.pt" % r)) for r in range(world)]
    assert outs[0]["calls"] == [1.0, 3.0] and outs[1]["calls"] == [2.0]          # cases[rank::world]
    for o in outs:                                                               # both ranks report the GLOBAL means
        assert abs(o["out"]["dsc"] - 0.5 * (1 + 2 + 3) / 3) < 1e-6
        assert abs(o["out"]["dsc_noft"] - 0.25 * (1 + 2 + 3) / 3) < 1e-6
    assert [round(x, 6) for x in outs[0]["out"]["scores"]] == [0.5, 1.5] and [round(x, 6) for x in outs[1]["out"]["scores"]] == [1.0]
