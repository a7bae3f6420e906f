"""TEST INFRASTRUCTURE ONLY -- "conditioned" parity fixtures: weights after K SGD steps of the reference path.

At PyTorch's default initialisation the pre-activations of this network sit on the ReLU knife-edge and the logit
margin is ~N(.46,.51): a single mask flip moves every upstream gradient by a per-cent, so a reduced-precision
implementation cannot be held to tight gradient / argmax bounds there (the fp32 reference itself misses them against
float64).  The parity tests therefore ALSO run on weights a few optimiser steps into training, produced by running the
reference's own train step (the CPU oracle, pinned torch.equal to the real reference modules by tests/test_oracle.py;
`tests/golden/conditioned.npz` holds the loss trajectory and weight checksums of the same recipe run through the REAL
reference, written by oracle/make_golden.py) on a synthetic task where the label is visible in the image:

    image = clamp(0.5 * randn + label - 0.25, -1, 1)        label = pancreas-like ellipsoid blob (synthetic.synth_label)

  seg:  main_source.py:415-437,660-661  (1 - Dice_fg, SGD momentum .9)
  vae:  main_source.py:389-406,660-661  (1 - Dice_fg + 2e-5 KL, if_random, scale .35)

The parity tests use lr 0.1 (the reference's default 1e-2 needs hundreds of steps to leave the random-init regime):
60 steps bring the Seg Dice loss from 0.97 to ~0.18 and the logit margin from N(.46,.51) to |margin| ~ 6.

Everything is drawn from the torch CPU generator, so the same seeds give the same tensors on the authoring container
and on the GPU box; results are cached per (kind, arguments) for the life of the process.
"""
import os
from collections import OrderedDict

import torch

from oracle import ref_torch as R
from vae_segmentation_b200.synthetic import synth_image, synth_label

_CACHE = {}
EPS = 0.000001          # utils/evaluation.py:48 avg_dsc (the importable twin of main_source.py's eps=1e-4 copy)
# optional on-disk cache (git-ignored; travels to the GPU box with the gpurun snapshot so the box does not re-train)
_DISK = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_cache")


def _disk_get(key):
    path = os.path.join(_DISK, "_".join(str(k) for k in key) + ".pt")
    if os.path.isfile(path):
        try:
            return torch.load(path)
        except Exception:
            return None
    return None


def _disk_put(key, value):
    if key[1] < 20:          # short recipes (the golden check of the recipe itself) are cheaper to redo than to ship
        return
    try:
        os.makedirs(_DISK, exist_ok=True)
        torch.save(value, os.path.join(_DISK, "_".join(str(k) for k in key) + ".pt"))
    except OSError:
        pass


def blob_batch(batch, patch):
    """(image, label): the label blob is visible in the image (contrast 1.0 on noise sigma 0.5)."""
    label = synth_label(batch, patch)
    img = (0.5 * synth_image(batch, patch) + label - 0.25).clamp_(-1.0, 1.0)
    return img, label


def train_seg(steps, patch=32, batch=2, seed=7, lr=1e-2, momentum=0.9, step_fn=None):
    """K reference Seg train steps from default init.  Returns (state_dict, losses list).  `step_fn(sd, img, label)
    -> (loss, grads)` lets oracle/make_golden.py run the same recipe through the real reference modules."""
    key = ("seg", steps, patch, batch, seed, lr, momentum, EPS, step_fn is None)
    if key in _CACHE:
        return _CACHE[key]
    hit = _disk_get(key) if step_fn is None else None
    if hit is not None:
        _CACHE[key] = hit
        return hit
    torch.manual_seed(seed)
    sd = R.init_seg_state()
    bufs, losses = None, []
    for _ in range(steps):
        img, label = blob_batch(batch, patch)
        if step_fn is None:
            loss, grads, _ = R.seg_train_step(sd, img, label, eps=EPS)
        else:
            loss, grads = step_fn(sd, img, label)
        sd, bufs = R.sgd_step(sd, grads, bufs, lr=lr, momentum=momentum)
        losses.append(float(loss))
    out = (OrderedDict((k, v.detach().clone()) for k, v in sd.items()), losses)
    if step_fn is None:
        _CACHE[key] = out
        _disk_put(key, out)
    return out


def train_vae(steps, patch=64, batch=2, seed=8, lr=1e-2, momentum=0.9, scale=0.35, step_fn=None):
    """K reference VAE train steps (main_source.py:389-406) from default init on blob masks."""
    key = ("vae", steps, patch, batch, seed, lr, momentum, scale, EPS, step_fn is None)
    if key in _CACHE:
        return _CACHE[key]
    hit = _disk_get(key) if step_fn is None else None
    if hit is not None:
        _CACHE[key] = hit
        return hit
    torch.manual_seed(seed)
    sd = R.init_vae_state(2, 128, patch)
    bufs, losses = None, []
    for _ in range(steps):
        label = synth_label(batch, patch)
        z = torch.randn(batch, 128)
        if step_fn is None:
            loss, _, _, grads, _ = R.vae_train_step(sd, label, scale=scale, z=z, eps=EPS)
        else:
            loss, grads = step_fn(sd, label, z)
        sd, bufs = R.sgd_step(sd, grads, bufs, lr=lr, momentum=momentum)
        losses.append(float(loss))
    out = (OrderedDict((k, v.detach().clone()) for k, v in sd.items()), losses)
    if step_fn is None:
        _CACHE[key] = out
        _disk_put(key, out)
    return out


def checksum(sd):
    """[sum, l2] per parameter in float64: compact fingerprint of a state_dict."""
    return torch.stack([torch.stack([v.double().sum(), v.double().norm()]) for v in sd.values()]).numpy()
