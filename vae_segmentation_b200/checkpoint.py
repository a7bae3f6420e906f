"""Checkpoint interop with the reference drivers.

The reference saves `{'epoch', 'model_state_dict', 'optimizer_state_dict'}` with `torch.save`
(main_target.py:1049-1062, main_source.py likewise) -- `model.state_dict()` of the UNWRAPPED module (Segmentation,
VAE or Joint with `Seg.` / `Vae.` prefixes) and the `state_dict()` of a `torch.optim.SGD` / `Adam` over
`model.parameters()` -- and loads whole or partial models from it (`main_target.py:358-394`: `model.load_state_dict`,
`model.Seg.load_state_dict`, `model.Vae.load_state_dict`).  The drop-in modules keep the reference's parameter names,
shapes and fp32 dtype, so the model part is a plain `load_state_dict(strict=True)`; this module adds the file format
and the translation between torch's per-parameter optimiser state and the flat arenas of the fused optimisers
(`train_step.FusedSGD` / `FusedAdam`), so a run can be resumed on either side.
"""
import collections

import torch


def _strip_module_prefix(sd):
    """`nn.DataParallel(model).state_dict()` prefixes every key with 'module.'; the reference saves the unwrapped
    model, but checkpoints written from the wrapper exist in the wild."""
    if sd and all(k.startswith("module.") for k in sd):
        return collections.OrderedDict((k[len("module."):], v) for k, v in sd.items())
    return sd


def load_model_state(model, path_or_ckpt, part=None, strict=True, map_location="cpu"):
    """Loads `ckpt['model_state_dict']` into `model` (or into `getattr(model, part)`, e.g. part='Seg' for
    main_target.py:363 `model.Seg.load_state_dict(...)`).  Returns the checkpoint's epoch (or None)."""
    ckpt = torch.load(path_or_ckpt, map_location=map_location) if not isinstance(path_or_ckpt, dict) else path_or_ckpt
    sd = _strip_module_prefix(ckpt["model_state_dict"] if "model_state_dict" in ckpt else ckpt)
    target = getattr(model, part) if part else model
    target.load_state_dict(sd, strict=strict)
    for m in model.modules():                    # derived weight packs of the drop-in modules are now stale
        cache = getattr(m, "_cache", None)
        if cache is not None:
            cache.invalidate()
    return ckpt.get("epoch") if isinstance(ckpt, dict) else None


def flat_to_param_state(params, flat, key):
    """Splits a flat fp32 buffer laid out in `params` order into torch.optim's per-parameter state entries."""
    out, off = {}, 0
    for i, p in enumerate(params):
        n = p.numel()
        out[i] = {key: flat[off:off + n].detach().reshape(p.shape).clone()}
        off += n
    return out


def sgd_state_dict(all_params, trained, momentum_flat, lr, momentum, steps):
    """`torch.optim.SGD(all_params, lr, momentum).state_dict()` equivalent.  all_params: `list(model.parameters())` of
    the model the reference optimiser was built over; trained: the (ordered) subset whose momentum lives in
    `momentum_flat` (the fused optimiser's arena order).  Parameters that never received a gradient (the frozen VAE)
    carry no state, exactly as in torch."""
    index = {id(p): i for i, p in enumerate(all_params)}
    state = {}
    if momentum_flat is not None and momentum != 0 and steps > 0:
        per = flat_to_param_state(trained, momentum_flat, "momentum_buffer")
        for j, p in enumerate(trained):
            state[index[id(p)]] = per[j]
    group = {"lr": lr, "momentum": momentum, "dampening": 0, "weight_decay": 0, "nesterov": False, "maximize": False,
             "foreach": None, "differentiable": False, "fused": None, "params": list(range(len(all_params)))}
    return {"state": state, "param_groups": [group]}


def adam_state_dict(all_params, trained, m_flat, v_flat, lr, betas, eps, steps):
    index = {id(p): i for i, p in enumerate(all_params)}
    state = {}
    if steps > 0:
        pm = flat_to_param_state(trained, m_flat, "exp_avg")
        pv = flat_to_param_state(trained, v_flat, "exp_avg_sq")
        for j, p in enumerate(trained):
            state[index[id(p)]] = {"step": torch.tensor(float(steps)), "exp_avg": pm[j]["exp_avg"],
                                   "exp_avg_sq": pv[j]["exp_avg_sq"]}
    group = {"lr": lr, "betas": tuple(betas), "eps": eps, "weight_decay": 0, "amsgrad": False, "maximize": False,
             "foreach": None, "capturable": False, "differentiable": False, "fused": None,
             "params": list(range(len(all_params)))}
    return {"state": state, "param_groups": [group]}


def param_state_to_flat(all_params, trained, osd, key, out):
    """Inverse of `flat_to_param_state`: copies state entry `key` of every trained parameter into the flat buffer
    `out` (arena order); parameters without state (never stepped) are zero-filled.  Returns #parameters restored."""
    index = {id(p): i for i, p in enumerate(all_params)}
    state = osd.get("state", {})
    off = restored = 0
    for p in trained:
        n = p.numel()
        ent = state.get(index[id(p)])
        if ent is not None and key in ent:
            out[off:off + n].copy_(ent[key].reshape(-1).to(out.device, out.dtype))
            restored += 1
        else:
            out[off:off + n].zero_()
        off += n
    return restored


def save_checkpoint(path, model, epoch, trainer=None, optimizer=None):
    """Writes the reference's checkpoint format.  optimizer: a torch.optim optimiser (its own state_dict is used) --
    or trainer: a train_step.SegTrainer / VAETrainer / JointTrainer whose fused optimiser state is translated."""
    if optimizer is not None:
        osd = optimizer.state_dict()
    elif trainer is not None:
        opt = trainer.opt
        all_params = list(model.parameters())
        trained = opt.arena.params
        if hasattr(opt, "momentum"):
            osd = sgd_state_dict(all_params, trained, opt.buf, opt.lr, opt.momentum, opt.steps)
        else:
            osd = adam_state_dict(all_params, trained, opt.m, opt.v, opt.lr, opt.betas, opt.eps, opt.steps)
    else:
        osd = {}
    msd = collections.OrderedDict((k, v.detach().cpu().clone()) for k, v in model.state_dict().items())
    torch.save({"epoch": epoch, "model_state_dict": msd, "optimizer_state_dict": osd}, path)


def load_checkpoint(path, model, trainer=None, strict=True, map_location="cpu"):
    """Loads a reference-format checkpoint into `model` and, when given, restores the fused optimiser of `trainer`
    (momentum / Adam moments) from `optimizer_state_dict`.  Returns the epoch."""
    ckpt = torch.load(path, map_location=map_location)
    epoch = load_model_state(model, ckpt, strict=strict)
    osd = ckpt.get("optimizer_state_dict") or {}
    if trainer is not None:
        # the trainer re-homed the parameters into its arena: load_state_dict copied INTO those views, so the arena
        # is current; only the derived packs and the optimiser moments remain
        trainer.opt.arena.module.repack_packs()
        if hasattr(trainer, "release_graph"):
            trainer.release_graph()                  # captured graphs may hold lazily re-packed state of the old weights
        if osd.get("state"):
            opt = trainer.opt
            all_params, trained = list(model.parameters()), opt.arena.params
            if hasattr(opt, "momentum"):
                if opt.buf is not None:
                    n = param_state_to_flat(all_params, trained, osd, "momentum_buffer", opt.buf)
                    opt.steps = max(opt.steps, 1 if n else 0)         # buffers are live: not the "first" step any more
            else:
                param_state_to_flat(all_params, trained, osd, "exp_avg", opt.m)
                param_state_to_flat(all_params, trained, osd, "exp_avg_sq", opt.v)
                steps = [int(e["step"]) for e in osd["state"].values() if "step" in e]
                opt.steps = max(steps) if steps else opt.steps
    return epoch
