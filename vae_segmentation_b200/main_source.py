"""Driver shim: the reference's `main_source.py` command line on the drop-in modules, for the two source-domain
methods the shipped presets run (scripts/source/seg_nih.bash, vae_nih.bash; BASELINE.json configs[0] and [1]).

    python -m vae_segmentation_b200.main_source seg_nih -G 0 --method seg_train --train_list NIH_train --val_list NIH_val \\
        --data_root <NIH numpy dir> --val_data_root <NIH numpy dir> --data_path data/Multi_all.json \\
        --eval_epoch 20 --save_epoch 800 --max_epoch 2400

i.e. the argument vector of scripts/source/*.bash (main_source.py:26-57) is accepted verbatim.  What runs:

  * `--method vae_train` (main_source.py:389-413): VAETrainer -- one-hot of the label, `vae(onehot, if_random=True,
    scale=0.35)`, loss (1 - Dice_fg) + 2e-5 KL, SGD(lr_seg, momentum 0.9) (:350-351); validation (:632-647):
    reconstruction with if_random=False, binary Dice of the foreground.
  * `--method seg_train` (:415-447): SegTrainer -- 1 - Dice_fg, SGD(lr_seg, momentum 0.9); the first block of
    `--eval_epoch` epochs only validates (`if epoch == 0: continue`, :416); validation (:649-760): binary Dice of the
    prediction.  `--load_prefix` loads checkpoints/<load_prefix>/<checkpoint_name> into the model (:300-305).
  * every `--save_epoch` epochs `model_epoch<N>.ckpt`, and `best_model.ckpt` when the validation Dice improved
    (:826-844), in the reference's {'epoch','model_state_dict','optimizer_state_dict'} format; `--test_only`.
  * the train step is replayed from a captured CUDA graph (`--no_graph`: eager launches); data, `--synthetic N`,
    multi-GPU (torchrun) exactly as in vae_segmentation_b200.main_target.

The other methods of main_source.py (joint_train, sep_joint_train, embed_train, refine_vae) are not wired: no shipped
preset uses them; `domain_adaptation` lives in vae_segmentation_b200.main_target.
"""
import argparse
import os
import sys

import torch
import torch.distributed as dist

from .main_target import DeviceCases, SyntheticCases, filedict_from_json, mask_index_from


def build_parser():
    p = argparse.ArgumentParser(description=__doc__.split("\n")[0])
    # main_source.py:26-57, same names / short options / defaults
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
    p.add_argument("-R", "--data_root", default="../nih_data/numpy_data/")
    p.add_argument("-V", "--val_data_root", default="../nih_data/numpy_data/")
    p.add_argument("-l", "--data_path", default="Multi_all.json")
    p.add_argument("-t", "--train_list", default="NIH_train")
    p.add_argument("-v", "--val_list", default="NIH_val")
    p.add_argument("--load_prefix", default=None)
    p.add_argument("--checkpoint_name", default="best_model.ckpt")
    p.add_argument("--load_prefix_vae", default=None)
    p.add_argument("--load_prefix_joint", default=None)
    p.add_argument("--pan_index", default="1")
    p.add_argument("--lambda_vae", type=float, default=0.1)
    p.add_argument("--lambda_vae_warmup", type=int, default=0)
    p.add_argument("--lr_seg", type=float, default=1e-2)
    p.add_argument("--lr_vae", type=float, default=0)
    for flag in ("test_only", "resume", "save_more_reference", "save_eval_result", "no_aug", "adam"):
        p.add_argument("--" + flag, action="store_true")
    p.add_argument("--mode", type=int, default=0)
    # additions of this shim
    p.add_argument("--synthetic", type=int, default=0, help="use N synthetic training volumes (and N//4+1 validation cases)")
    p.add_argument("--patch", type=int, default=128, help="patch edge (the reference hard-codes 128)")
    p.add_argument("--precision", default="bf16", choices=["bf16", "fp32"])
    p.add_argument("--save_root", default="checkpoints", help="the reference's save_root_path")
    p.add_argument("--max_iters", type=int, default=0, help="stop after this many training iterations (smoke runs)")
    p.add_argument("--no_graph", action="store_true", help="eager launches instead of CUDA-graph replay of the train step")
    return p


def validate(method, model, cases, names, world, rank, dev):
    """main_source.py:632-647 (vae_train) / :649-760 (seg_train): mean over the cases of the binary (argmax) Dice of
    the foreground class; cases sharded over ranks, two sums all-reduced."""
    from . import evaluation as ev
    scores = []
    with torch.no_grad():
        for idx, name in enumerate(names):
            if idx % world != rank:
                continue
            img, label = cases.load(name)
            onehot = ev.one_hot(label, 2)
            if method == "vae_train":
                out, _, _ = model(onehot, if_random=False)
            else:
                out = model.predict(img)
            scores.append(ev.avg_dsc({"p": out, "t": onehot}, "p", "t", binary=True, botindex=1, topindex=2).reshape(()))
    local = torch.stack([torch.stack(scores).sum() if scores else torch.zeros((), device=dev),
                         torch.tensor(float(len(scores)), device=dev)])
    if world > 1:
        dist.all_reduce(local, op=dist.ReduceOp.SUM)
    total = local.cpu()
    return total[0].item() / max(int(total[1].item()), 1)


def main(argv=None):
    args = build_parser().parse_args(argv)
    if args.method not in ("vae_train", "seg_train"):
        raise SystemExit("vae_segmentation_b200.main_source: --method %s is not wired here (vae_train and seg_train are; "
                         "domain_adaptation is vae_segmentation_b200.main_target)" % args.method)
    assert args.save_epoch % args.eval_epoch == 0                                         # main_source.py:89
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", args.GPU.split(",")[0] if world == 1 else "0"))
    if not torch.cuda.is_available():
        raise SystemExit("vae_segmentation_b200.main_source: no CUDA device -- the hot path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    from . import checkpoint as ck
    from . import joint_model as models
    from . import train_step as ts

    say = (lambda *a: print(*a, flush=True)) if rank == 0 else (lambda *a: None)
    P, method = args.patch, args.method
    torch.manual_seed(0)
    if method == "vae_train":                                                              # main_source.py:249-250
        model = models.VAE(n_channels=2, n_class=2, norm_type=1, dim=128, patch=P)
    else:                                                                                  # :251-253
        model = models.Segmentation(n_channels=1, n_class=2, norm_type=1)
    save_dir = os.path.join(args.save_root, args.prefix)
    if args.load_prefix:                                                                   # :300-305
        ck.load_checkpoint(os.path.join(args.save_root, args.load_prefix, args.checkpoint_name), model)
    model.to(dev).set_precision(args.precision)
    if method == "vae_train":
        trainer = ts.VAETrainer(model, lr=args.lr_seg, momentum=0.9, scale=0.35)          # :350-351 SGD(lr1, momentum .9)
    else:
        trainer = ts.SegTrainer(model, lr=args.lr_seg, momentum=0.9)
    if args.load_prefix_vae and method == "seg_train":
        say("note: the frozen reference VAE of seg_train (main_source.py:307-313) only feeds a monitored recon_loss; not computed")

    # ---- data -------------------------------------------------------------------------------------------------------
    if not args.no_aug:
        say("note: the batchgenerators spatial augmentation (main_source.py:196-207) is not implemented; running as --no_aug")
    if args.synthetic:
        train_set = SyntheticCases(args.synthetic, P, dev, seed=1)
        val_set = SyntheticCases(args.synthetic // 4 + 1, P, dev, seed=2)
        train_names, val_names = train_set.names * args.eval_epoch, val_set.names
    else:
        mi = mask_index_from(args.pan_index)
        train_names = filedict_from_json(args.data_path, args.train_list, args.eval_epoch)
        val_names = filedict_from_json(args.data_path, args.val_list, 1)
        train_set = DeviceCases(args.data_root, train_names, mi, P, 0, dev)
        val_set = DeviceCases(args.val_data_root, val_names, mi, P, 0, dev)
    B = max(args.batch_size // world, 1)                                                   # global batch sharded over ranks
    steps_per_epoch = len(train_names) // (B * world)                                      # drop_last=True
    say("train volumes per epoch %d, per-GPU batch %d x %d GPU(s), %d validation cases" % (len(train_names), B, world, len(val_names)))

    best, iters = 0.0, 0
    static = None
    with torch.cuda.stream(trainer.stream):
        for epoch in range(args.max_epoch // args.eval_epoch):
            skip_train = args.test_only or (method == "seg_train" and epoch == 0)          # :416 `if epoch == 0: continue`
            if not skip_train:
                shuffle = torch.Generator()
                shuffle.manual_seed(1000 + epoch)
                perm = torch.randperm(len(train_names), generator=shuffle).tolist()        # DataLoader(shuffle=True), :280
                for idx in range(steps_per_epoch):
                    pick = perm[(idx * world + rank) * B:(idx * world + rank + 1) * B]
                    pairs = [train_set.load(train_names[i]) for i in pick]
                    img = torch.cat([p[0] for p in pairs]).contiguous()
                    label = torch.cat([p[1] for p in pairs]).contiguous()
                    z = torch.randn(B, 128) if method == "vae_train" else None            # CPU generator, joint_model.py:246
                    if args.no_graph:
                        mon = trainer.step(label, z) if method == "vae_train" else trainer.step(img, label)
                    else:
                        if static is None or static[1].shape != label.shape:
                            trainer.release_graph()
                            static = (img.clone(), label.clone(), torch.zeros(B, 128, device=dev))
                            if method == "vae_train":
                                trainer.capture(static[1], static[2])
                            else:
                                trainer.capture(static[0], static[1])
                        static[0].copy_(img)
                        static[1].copy_(label)
                        if z is not None:
                            static[2].copy_(z, non_blocking=True)
                        mon = trainer.step_graphed()
                    iters += 1
                    if idx % 10 == 0:
                        say("[%3d, %3d] loss: " % ((epoch + 1) * args.eval_epoch, idx + 1) +
                            ", ".join("%.4f" % mon[k].item() for k in (("dice_loss", "kl_loss") if method == "vae_train" else ("dice_loss",))))
                    if args.max_iters and iters >= args.max_iters:
                        break
            # ---- validation ----
            model.eval()                                                                   # main_source.py:690
            dsc = validate(method, model, val_set, val_names, world, rank, dev)
            model.train()                                                                  # :822
            say("epoch %d validation result: %f, best result %f." % (epoch + 1, dsc, best))
            if args.test_only or (args.max_iters and iters >= args.max_iters):
                break
            if (epoch + 1) % (args.save_epoch // args.eval_epoch) == 0 and rank == 0:      # :826-844
                os.makedirs(save_dir, exist_ok=True)
                ck.save_checkpoint(os.path.join(save_dir, "model_epoch%d.ckpt" % ((epoch + 1) * args.eval_epoch)), model,
                                   epoch=(epoch + 1) * args.eval_epoch, trainer=trainer)
                if dsc > best:
                    best = dsc
                    ck.save_checkpoint(os.path.join(save_dir, "best_model.ckpt"), model, epoch=(epoch + 1) * args.eval_epoch,
                                       trainer=trainer)
    trainer.release_graph()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
