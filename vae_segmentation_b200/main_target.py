"""Driver shim: the reference's `main_target.py` command line on the drop-in modules.

    python -m vae_segmentation_b200.main_target domain_msd_dh_ft1 -G 0 --method domain_adaptation \\
        --load_prefix seg_nih --load_prefix_vae vae_nih --train_list MSD_train --val_list MSD_val \\
        --data_root <MSD numpy dir> --val_data_root <MSD numpy dir> --data_path data/Multi_all.json --pan_index 10 \\
        --lambda_vae 1.0 --domain_loss_type 8 --val_finetune 1 --eval_epoch 2 --save_epoch 100 --max_epoch 50

i.e. the argument vector of scripts/target/*.bash (main_target.py:28-82) is accepted verbatim.  What runs:

  * method domain_adaptation (the teacher-student path, main_target.py:505-760): JointTrainer -- EMA teacher update
    (:508-518, `--pseudo_save_epoch`, `--update_every_iteration`, `--alpha`, `--tag`), dynamic lambda (`--domain_loss_type`
    0 / 8), `--kl`, `--only_pseudo`, `--use_confident_binarize`, `--adam`, `--from_scratch`; validation every
    `--eval_epoch` epochs with `--val_finetune` test-time-training iterations (:795-960); `--test_only`; `--resume`;
    checkpoints in the reference's {'epoch','model_state_dict','optimizer_state_dict'} format under
    checkpoints/<prefix>/ (:1049-1062), loaded from checkpoints/<load_prefix>/<checkpoint_name> like the reference.
  * data: the reference's `merge.npy` volumes ([...,0] image, [...,1] label index) listed in the `--data_path` JSON
    (utils/utils.py:326-384), remapped with `--pan_index`, then ON THE DEVICE CropResize -> Clip(-200,400) ->
    CenterIntensities(100,300) (transforms.py; main_target.py:203-225).  `--synthetic N` replaces the files by N
    synthetic volumes (no dataset ships with the reference).  The batchgenerators spatial augmentation
    (MySpatialTransform, :207-218) is not implemented: the run behaves as with `--no_aug` and says so.
  * multi-GPU: launch under torchrun (one process per GPU, `-G` is then ignored); batches are sharded over ranks.

Everything else of main_target.py (tensorboard Saver, figure dumps, the discriminator / embed methods) is out of scope.
"""
import argparse
import json
import os
import re
import sys

import torch
import torch.distributed as dist


def build_parser():
    p = argparse.ArgumentParser(description=__doc__.split("\n")[0])
    p.add_argument("prefix", help="prefix")
    p.add_argument("-P", "--target_phase", default="arterial")
    p.add_argument("-G", "--GPU", default="0,1,2,3")
    p.add_argument("-b", "--batch_size", type=int, default=4)
    p.add_argument("-E", "--max_epoch", type=int, default=1600)
    p.add_argument("--save_epoch", type=int, default=50)
    p.add_argument("--eval_epoch", type=int, default=50)
    p.add_argument("--turn_epoch", type=int, default=-1)
    p.add_argument("-S", "--softrelu", type=int, default=0)
    p.add_argument("-M", "--method", default="vae_train")
    p.add_argument("--data_root", default="../nih_data/numpy_data/")
    p.add_argument("--val_data_root", default="../nih_data/numpy_data/")
    p.add_argument("--pseudo_data_root", default="../nih_data/numpy_data/")
    p.add_argument("-l", "--data_path", default="Multi_all.json")
    p.add_argument("--train_list", default="NIH_train")
    p.add_argument("--val_list", default="NIH_val")
    p.add_argument("--pseudo_list", default=None)
    p.add_argument("--load_prefix", default=None)
    p.add_argument("--checkpoint_name", default="best_model.ckpt")
    p.add_argument("--load_prefix_vae", default=None)
    p.add_argument("--load_prefix_encoder", default=None)
    p.add_argument("--load_prefix_joint", default=None)
    p.add_argument("--pan_index", default="1")
    p.add_argument("--pseudo_pan_index", default="1")
    p.add_argument("--lambda_vae", type=float, default=0.1)
    p.add_argument("--lambda_vae_warmup", type=int, default=0)
    p.add_argument("--lr_seg", type=float, default=1e-2)
    p.add_argument("--lr_vae", type=float, default=0)
    for flag in ("test_only", "resume", "save_more_reference", "save_eval_result", "no_aug", "only_pseudo", "fix_layer",
                 "use_confident_binarize", "tag", "from_scratch", "adam", "kl", "update_every_iteration",
                 "generate_bounding_boxes"):
        p.add_argument("--" + flag, action="store_true")
    p.add_argument("--analysis_figure_name", default=None)
    p.add_argument("--pseudo_save_epoch", type=int, default=0)
    p.add_argument("--domain_loss_type", type=int, default=0)
    p.add_argument("--vae_mont_number", type=int, default=1)
    p.add_argument("--vae_forward_scale", type=float, default=0.0)
    p.add_argument("--vae_decoder_dropout", type=float, default=0.0)
    p.add_argument("--seg_dropout", type=float, default=0.0)
    p.add_argument("--val_finetune", type=int, default=0)
    p.add_argument("--lr_finetune", type=float, default=1e-2)
    p.add_argument("--alpha", type=float, default=0.995)
    p.add_argument("--shift", type=int, default=0)
    # ---- additions of the shim (not in the reference) ----
    p.add_argument("--synthetic", type=int, default=0, help="use N synthetic training volumes (and N//4+1 validation cases)")
    p.add_argument("--patch", type=int, default=128, help="patch edge (the reference hard-codes 128)")
    p.add_argument("--precision", default="bf16", choices=["bf16", "fp32"])
    p.add_argument("--save_root", default="checkpoints", help="the reference's save_root_path")
    p.add_argument("--max_iters", type=int, default=0, help="stop after this many training iterations (smoke runs)")
    p.add_argument("--no_graph", action="store_true", help="eager launches instead of CUDA-graph replay of the train step")
    return p


def mask_index_from(pan_index):
    """main_target.py:121-124."""
    if pan_index != "10":
        return [[0, 0]] + [[int(f), idx + 1] for idx, f in enumerate(pan_index.split(","))]
    return [[0, 0], [[1, 2], 1]]


def filedict_from_json(json_path, key, epoch=1):
    with open(json_path, "r") as f:
        listdict = json.load(f).get(key, [])
    return listdict * epoch


class DeviceCases(object):
    """merge.npy volumes -> device patches through the device transforms (CropResize, Clip + CenterIntensities)."""

    def __init__(self, root, names, mask_index, patch, shift, device):
        from .transforms import ClipCenter, CropResize
        self.root, self.names, self.mask_index, self.device = root, list(names), mask_index, device
        self.crop = CropResize(["venous"], [patch] * 3, shift=shift)
        self.norm = ClipCenter(["venous"], -200.0, 400.0, 100.0, 300.0)

    def __len__(self):
        return len(self.names)

    def load(self, name):
        import numpy as np
        merge = np.load(os.path.join(self.root, name))
        img = torch.from_numpy(merge[..., 0].astype(np.float32)).to(self.device)
        raw = torch.from_numpy(merge[..., 1].astype(np.float32)).to(self.device)
        label = torch.zeros_like(raw)
        for lab in self.mask_index:                                       # utils/utils.py:369-374
            for v in (lab[0] if isinstance(lab[0], list) else [lab[0]]):
                label[raw == v] = float(lab[1])
        d = {"venous": img, "venous_pancreas": label, "id": "".join(re.findall(r"\d+", name))}
        d = self.norm(self.crop(d))
        return d["venous"][None, None], d["venous_pancreas"][None, None]


class SyntheticCases(object):
    def __init__(self, count, patch, device, seed):
        from .synthetic import synth_image, synth_label
        g = torch.random.get_rng_state()
        torch.manual_seed(seed)
        self.items = []
        for _ in range(count):
            label = synth_label(1, patch)
            img = (0.5 * synth_image(1, patch) + label - 0.25).clamp_(-1.0, 1.0)
            self.items.append((img.to(device), label.to(device)))
        torch.random.set_rng_state(g)
        self.names = ["synthetic_%03d" % i for i in range(count)]

    def __len__(self):
        return len(self.items)

    def load(self, name):
        return self.items[self.names.index(name)]


def main(argv=None):
    args = build_parser().parse_args(argv)
    if args.method != "domain_adaptation":
        raise SystemExit("vae_segmentation_b200.main_target: --method %s is not wired here (domain_adaptation is; the source-domain "
                         "steps are train_step.SegTrainer / VAETrainer)" % args.method)
    if args.kl:
        assert args.domain_loss_type in (0, 8)
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", args.GPU.split(",")[0] if world == 1 else "0"))
    if not torch.cuda.is_available():
        raise SystemExit("vae_segmentation_b200.main_target: no CUDA device -- the hot path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    from . import checkpoint as ck
    from . import joint_model as models
    from . import train_step as ts

    say = (lambda *a: print(*a, flush=True)) if rank == 0 else (lambda *a: None)
    P = args.patch
    mk = lambda: models.Joint([models.Segmentation(n_channels=1, n_class=2, norm_type=1),
                               models.VAE(n_channels=2, n_class=2, norm_type=1, dim=128, patch=P)],
                              vae_forward_scale=args.vae_forward_scale, vae_decoder_dropout=args.vae_decoder_dropout,
                              seg_dropout=args.seg_dropout)                                # main_target.py:314-333
    torch.manual_seed(0)
    model, model_fix = mk(), mk()
    model_finetune = mk() if args.val_finetune != 0 else None
    save_dir = os.path.join(args.save_root, args.prefix)
    if args.load_prefix:                                                                  # :355-364
        path = os.path.join(args.save_root, args.load_prefix, args.checkpoint_name)
        ck.load_checkpoint(path, model_fix.Seg if args.from_scratch else model.Seg)
    if args.load_prefix_vae:                                                              # :366-371
        path = os.path.join(args.save_root, args.load_prefix_vae, "best_model.ckpt")
        if args.from_scratch:
            ck.load_checkpoint(path, model_fix.Vae)
        ck.load_checkpoint(path, model.Vae)
    if args.load_prefix_joint:
        ck.load_checkpoint(os.path.join(args.save_root, args.load_prefix_joint, "best_model.ckpt"), model)
    if not args.from_scratch:
        model_fix.load_state_dict(model.state_dict())                                     # :428 teacher = student
    for m in (model, model_fix, model_finetune):
        if m is not None:
            m.to(dev).set_precision(args.precision)
    # main_target.py:336,399,433 (no effect on the InstanceNorm models the driver builds; kept for the same module state)
    model.Vae.eval()
    model_fix.eval()
    if model_finetune is not None:
        model_finetune.Vae.eval()
    trainer = ts.JointTrainer(model, model_fix, lr=args.lr_seg, momentum=0.9, lambda_vae=args.lambda_vae,
                              loss_type=args.domain_loss_type, kl=args.kl, confident=args.use_confident_binarize,
                              only_pseudo=args.only_pseudo, alpha=args.alpha, adam=args.adam)
    start_epoch = 0
    if args.resume:
        start_epoch = ck.load_checkpoint(os.path.join(save_dir, "model_latest.ckpt"), model.Seg, trainer=trainer)
        say("resumed from epoch", start_epoch)

    # ---- data -------------------------------------------------------------------------------------------------------
    if not args.no_aug:
        say("note: the batchgenerators spatial augmentation (main_target.py:207-218) is not implemented; running as --no_aug")
    if args.synthetic:
        train_set = SyntheticCases(args.synthetic, P, dev, seed=1)
        val_set = SyntheticCases(args.synthetic // 4 + 1, P, dev, seed=2)
        train_names, val_names = train_set.names * args.eval_epoch, val_set.names
    else:
        mi = mask_index_from(args.pan_index)
        train_names = filedict_from_json(args.data_path, args.train_list, args.eval_epoch)
        val_names = filedict_from_json(args.data_path, args.val_list, 1)
        train_set = DeviceCases(args.data_root, train_names, mi, P, args.shift, dev)
        val_set = DeviceCases(args.val_data_root, val_names, mi, P, args.shift, dev)
    B = max(args.batch_size // world, 1)                                                  # global batch sharded over ranks
    steps_per_epoch = len(train_names) // (B * world)                                     # drop_last=True
    say("train volumes per epoch %d, per-GPU batch %d x %d GPU(s), %d validation cases" % (len(train_names), B, world, len(val_names)))

    best, iters, lambda_vae = 0.0, 0, args.lambda_vae
    static = None                      # (image, label) buffers of the captured train step + the lambda it was captured with
    with torch.cuda.stream(trainer.stream):
        for epoch in range(start_epoch, args.max_epoch // args.eval_epoch):
            if not args.test_only and epoch > 0:                                          # :506 `if epoch == 0: continue`
                shuffle = torch.Generator()                  # own generator: the order must not depend on how many draws
                shuffle.manual_seed(1000 + epoch)            # the model code made (graph capture warms the step up twice)
                perm = torch.randperm(len(train_names), generator=shuffle).tolist()
                for idx in range(steps_per_epoch):
                    update = False
                    if args.pseudo_save_epoch != 0 and epoch % max(args.pseudo_save_epoch // args.eval_epoch, 1) == 0:
                        update = args.update_every_iteration or idx % max(steps_per_epoch // args.eval_epoch, 1) == 0
                    if update and args.tag:
                        lambda_vae = args.alpha * lambda_vae                              # :517
                        trainer.lambda_vae = lambda_vae
                    pick = perm[(idx * world + rank) * B:(idx * world + rank + 1) * B]
                    pairs = [train_set.load(train_names[i]) for i in pick]
                    img = torch.cat([p[0] for p in pairs]).contiguous()
                    label = torch.cat([p[1] for p in pairs]).contiguous()
                    if args.no_graph:
                        mon = trainer.step(img, label, update_teacher=update)
                    else:
                        # ~600 launches per step: the host cannot issue them as fast as the GPU runs them, so the step
                        # (zero_grad + forwards + losses + backward) is captured once per (batch shape, lambda) and replayed
                        if static is None or static[0].shape != img.shape or static[2] != lambda_vae:
                            trainer.release_graph()
                            static = (img.clone(), label.clone(), lambda_vae)
                            trainer.capture(static[0], static[1])
                        static[0].copy_(img)
                        static[1].copy_(label)
                        mon = trainer.step_graphed(update_teacher=update)
                    iters += 1
                    if idx % 10 == 0:
                        say("epoch %d iter %d: " % (epoch, idx) + " ".join("%s %.4f" % (k, v.item()) for k, v in mon.items()))
                    if args.max_iters and iters >= args.max_iters:
                        break
            # ---- validation (:795-960) ----
            cases = (val_set.load(n) for n in val_names)
            model.eval()                                                                  # :756
            out = trainer.validate(cases, finetune=model_finetune, val_finetune=args.val_finetune, lr_finetune=args.lr_finetune)
            model.train()                                                                 # :1041-1043
            model.Vae.eval()
            say("epoch %d validation result: %f, best result %f." % (epoch + 1, out["dsc"], best) +
                (" (no finetune: %f)" % out["dsc_noft"] if args.val_finetune else ""))
            if args.test_only or (args.max_iters and iters >= args.max_iters):
                break
            if rank == 0:
                os.makedirs(save_dir, exist_ok=True)
                if out["dsc"] > best:
                    best = out["dsc"]
                    ck.save_checkpoint(os.path.join(save_dir, "best_model.ckpt"), model.Seg, epoch=(epoch + 1) * args.eval_epoch, trainer=trainer)
                ck.save_checkpoint(os.path.join(save_dir, "model_latest.ckpt"), model.Seg, epoch=epoch + 1, trainer=trainer)
                if ((epoch + 1) * args.eval_epoch) % args.save_epoch == 0:
                    ck.save_checkpoint(os.path.join(save_dir, "model_epoch%d.ckpt" % ((epoch + 1) * args.eval_epoch)), model.Seg,
                                       epoch=(epoch + 1) * args.eval_epoch, trainer=trainer)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
