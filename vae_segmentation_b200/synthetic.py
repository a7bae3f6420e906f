"""Synthetic CT-like volumes and pancreas-like label masks.

The reference ships no data; benchmarks and parity tests use these.
Everything is drawn from the torch CPU default generator so a seed reproduces the same
tensors on the authoring container and on the GPU box.
"""
import torch


def synth_image(batch, patch, channels=1):
    """randn clipped to [-1, 1]: mimics Clip(-200,400) -> (x-100)/300
    (reference main_target.py:223-224)."""
    return torch.randn(batch, channels, patch, patch, patch).clamp_(-1.0, 1.0)


def synth_label(batch, patch):
    """Binary ellipsoid blob [B,1,P,P,P] float {0,1}, roughly 1-6 % foreground."""
    ax = torch.arange(patch, dtype=torch.float32)
    zz, yy, xx = torch.meshgrid(ax, ax, ax, indexing="ij")
    out = torch.zeros(batch, 1, patch, patch, patch)
    for b in range(batch):
        c = (0.35 + 0.3 * torch.rand(3)) * patch
        r = (0.12 + 0.12 * torch.rand(3)) * patch
        d = ((zz - c[0]) / r[0]) ** 2 + ((yy - c[1]) / r[1]) ** 2 + ((xx - c[2]) / r[2]) ** 2
        out[b, 0] = (d <= 1.0).float()
    return out


def synth_batch(seed, batch, patch):
    torch.manual_seed(seed)
    return synth_image(batch, patch), synth_label(batch, patch)
