"""Checkpoint interop with the reference's `{'epoch','model_state_dict','optimizer_state_dict'}` files
(main_target.py:1049-1062, :358-394).  CPU only: the file format, strict key/shape compatibility in BOTH directions
against the real reference modules (when /root/reference is present) and the optimiser-state translation against
torch.optim itself."""
import os

import pytest
import torch

from vae_segmentation_b200 import checkpoint as ck
from vae_segmentation_b200 import joint_model as jm

HAVE_REF = os.path.isdir("/root/reference")


def _ours():
    return jm.Joint([jm.Segmentation(1, 2, norm_type=1), jm.VAE(2, 2, norm_type=1, dim=128)])


@pytest.mark.skipif(not HAVE_REF, reason="the real reference is only mounted in the authoring container")
def test_reference_checkpoint_round_trip(tmp_path):
    from oracle import reference_shim
    ref, _ = reference_shim.load()
    torch.manual_seed(0)
    rj = ref.Joint(models=[ref.Segmentation(n_channels=1, n_class=2, norm_type=1),
                           ref.VAE(n_channels=2, n_class=2, norm_type=1, dim=128)])
    ropt = torch.optim.SGD(rj.parameters(), lr=1e-2, weight_decay=0, momentum=0.9)
    path = str(tmp_path / "best_model.ckpt")
    torch.save({"epoch": 40, "model_state_dict": rj.state_dict(), "optimizer_state_dict": ropt.state_dict()}, path)   # the reference's writer
    ours = _ours()
    assert ck.load_checkpoint(path, ours, strict=True) == 40
    for (k, a), (k2, b) in zip(rj.state_dict().items(), ours.state_dict().items()):
        assert k == k2 and torch.equal(a, b)
    # partial loads as the reference does them (main_target.py:363,372)
    seg_path = str(tmp_path / "seg.ckpt")
    torch.save({"epoch": 1, "model_state_dict": rj.Seg.state_dict(), "optimizer_state_dict": {}}, seg_path)
    ours2 = _ours()
    ck.load_model_state(ours2, seg_path, part="Seg")
    assert torch.equal(ours2.Seg.state_dict()["up5.conv.0.weight"], rj.Seg.state_dict()["up5.conv.0.weight"])
    # and back: a checkpoint written here loads into the reference with strict=True
    out = str(tmp_path / "ours.ckpt")
    ck.save_checkpoint(out, ours, epoch=41, optimizer=torch.optim.SGD(ours.parameters(), lr=1e-2, momentum=0.9))
    back = torch.load(out)
    assert back["epoch"] == 41 and set(back) == {"epoch", "model_state_dict", "optimizer_state_dict"}
    rj.load_state_dict(back["model_state_dict"], strict=True)
    ropt.load_state_dict(back["optimizer_state_dict"])


def test_dataparallel_prefix_and_file_format(tmp_path):
    ours = _ours()
    sd = {"module." + k: v for k, v in ours.state_dict().items()}
    other = _ours()
    ck.load_model_state(other, {"epoch": 3, "model_state_dict": sd})
    assert all(torch.equal(a, b) for a, b in zip(ours.state_dict().values(), other.state_dict().values()))


def test_fused_optimizer_state_translates_to_torch_optim():
    """The flat momentum arena <-> torch.optim.SGD state_dict: load the translated state into a real torch optimiser,
    take one step on both sides, compare."""
    torch.manual_seed(1)
    model = _ours()
    all_params = list(model.parameters())
    trained = list(model.Seg.parameters())                      # JointTrainer trains Seg only (the VAE is frozen)
    n = sum(p.numel() for p in trained)
    mom = torch.randn(n)
    osd = ck.sgd_state_dict(all_params, trained, mom, lr=1e-2, momentum=0.9, steps=7)
    opt = torch.optim.SGD(all_params, lr=1e-2, momentum=0.9)
    opt.load_state_dict(osd)                                    # torch accepts the translated dict
    grads = torch.randn(n)
    off = 0
    for p in trained:
        p.grad = grads[off:off + p.numel()].reshape(p.shape).clone()
        off += p.numel()
    before = torch.cat([p.detach().reshape(-1) for p in trained]).clone()
    opt.step()
    after = torch.cat([p.detach().reshape(-1) for p in trained])
    want = before - 1e-2 * (0.9 * mom + grads)                  # what the fused kernel computes (optim.cu sgd_kernel)
    assert torch.allclose(after, want, rtol=1e-6, atol=1e-7)
    # and back into a flat buffer
    flat = torch.empty(n)
    assert ck.param_state_to_flat(all_params, trained, opt.state_dict(), "momentum_buffer", flat) == len(trained)
    assert torch.allclose(flat, 0.9 * mom + grads, rtol=1e-6, atol=1e-7)
    # Adam moments
    m, v = torch.randn(n), torch.rand(n)
    aosd = ck.adam_state_dict(all_params, trained, m, v, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, steps=5)
    aopt = torch.optim.Adam(all_params, lr=1e-3)
    aopt.load_state_dict(aosd)
    f2 = torch.empty(n)
    ck.param_state_to_flat(all_params, trained, aopt.state_dict(), "exp_avg_sq", f2)
    assert torch.equal(f2, v)
