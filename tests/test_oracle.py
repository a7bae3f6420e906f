"""Pins the CPU oracle (oracle/ref_torch.py): against the golden fixtures generated from the
REAL reference (tests/golden/, oracle/make_golden.py) and, when /root/reference is present
(authoring container), against the real reference modules executed live."""
import os
from collections import OrderedDict

import numpy as np
import pytest
import torch

from oracle import make_golden as MG
from oracle import ref_torch as R
from oracle import reference_shim


def _summary(grads):
    return MG.grad_summary(grads)


def _close(a, b, rtol, atol):
    np.testing.assert_allclose(np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64), rtol=rtol, atol=atol)


def test_losses_match_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, "losses.npz"))
    a, t = torch.from_numpy(g["a"]), torch.from_numpy(g["t"])
    mean, std = torch.from_numpy(g["mean"]), torch.from_numpy(g["std"])
    _close(R.avg_dsc(a, t), g["dsc_full"], 1e-6, 0)
    _close(R.avg_dsc(a, t, botindex=1, topindex=2), g["dsc_fg"], 1e-6, 0)
    _close(R.avg_dsc(a, t, botindex=1, topindex=2, return_mean=False), g["dsc_fg_vec"], 1e-6, 0)
    _close(R.avg_dsc(a, t, binary=True, botindex=1, topindex=2), g["dsc_binary"], 1e-6, 0)
    _close(R.kl_loss(mean, std), g["kl"], 1e-6, 0)
    _close(R.dice(a, t), g["dice"], 1e-6, 0)
    assert np.array_equal(R.binarize(a).numpy(), g["binarize"])
    assert np.array_equal(R.confident_binarize(a).numpy(), g["confident"])


def test_seg_train_step_matches_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, "seg_p32.npz"))
    seg_sd, _, img, label = MG.case_inputs(MG.SEG_CASE, seg=True)
    loss, grads, pred = R.seg_train_step(seg_sd, img, label, eps=0.000001)
    _close(loss, g["loss"], 1e-5, 0)
    _close(MG.sample(pred), g["probs_sample"], 1e-4, 1e-6)
    _close(pred.sum((2, 3, 4)), g["probs_sum"], 1e-5, 0)
    assert list(grads.keys()) == list(g["grad_names"])
    _close(_summary(grads), g["grad_summary"], 2e-3, 2e-6)


@pytest.mark.timeout(600)
def test_vae_train_step_matches_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, "vae_p128.npz"))
    _, vae_sd, _, label = MG.case_inputs(MG.VAE_CASE, vae=True)
    torch.manual_seed(MG.VAE_CASE["seed"] + 1000)
    loss, dsc, kl, grads, recon = R.vae_train_step(vae_sd, label, scale=0.35, eps=0.000001)
    _close(loss, g["loss"], 1e-5, 0)
    _close(kl, g["kl"], 1e-5, 0)
    _close(dsc, g["dsc"], 1e-5, 0)
    _close(MG.sample(recon, 5), g["recon_sample"], 1e-4, 1e-6)
    _close(_summary(grads), g["grad_summary"], 5e-3, 5e-6)


@pytest.mark.timeout(600)
def test_joint_step_matches_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, "joint_p128.npz"))
    seg_sd, vae_sd, img, label = MG.case_inputs(MG.JOINT_CASE, vae=True, seg=True)
    out, grads = R.joint_target_step(seg_sd, vae_sd, seg_sd, img, label, lambda_vae=1.0, loss_type=0)
    for k in ("final", "recon_loss", "dsc_loss", "dsc_loss_fake", "klloss"):
        _close(out[k], g[k], 2e-5, 0)
    _close(MG.sample(out["pred"], 5), g["pred_sample"], 1e-4, 1e-6)
    _close(MG.sample(out["recon"], 5), g["recon_sample"], 1e-4, 1e-6)
    _close(_summary(grads), g["grad_summary"], 5e-3, 5e-6)


def test_dynamic_lambda_composition():
    # main_target.py:550-560
    one = torch.tensor
    for recon, lam, want in ((0.10, 1.0, 0.6 * 0.10 + 0.5), (0.20, 1.0, 0.20 + 0.5 / 1.2),
                             (0.25, 1.0, 0.25 + 0.5 / 2.0), (0.40, 0.2, 0.6 * 0.40 + 0.5)):
        got = R.compose_target_loss(one(recon), one(0.5), one(3.0), lambda_vae=lam, loss_type=8)
        assert abs(float(got) - want) < 1e-6
    got = R.compose_target_loss(one(0.2), one(0.5), one(3.0), lambda_vae=2.0, loss_type=0, kl=True)
    assert abs(float(got) - (2.0 * 0.2 + 0.5 + 0.00002 * 2.0 * 3.0)) < 1e-6


def test_sgd_and_ema_match_torch():
    torch.manual_seed(0)
    p = {"a": torch.randn(5), "b": torch.randn(3, 2)}
    params = [torch.nn.Parameter(v.clone()) for v in p.values()]
    opt = torch.optim.SGD(params, lr=1e-2, momentum=0.9)
    sd, bufs = dict(p), None
    for _ in range(3):
        grads = {k: torch.randn_like(v) for k, v in sd.items()}
        for q, g in zip(params, grads.values()):
            q.grad = g.clone()
        opt.step()
        sd, bufs = R.sgd_step(sd, grads, bufs, lr=1e-2, momentum=0.9)
    for q, v in zip(params, sd.values()):
        assert torch.allclose(q.detach(), v, atol=1e-7)
    e = R.ema_update({"a": torch.ones(2)}, {"a": torch.zeros(2)}, alpha=0.995)
    assert torch.allclose(e["a"], torch.full((2,), 0.995))


needs_ref = pytest.mark.skipif(not reference_shim.available(), reason="reference tree not present")


@needs_ref
def test_oracle_equals_real_reference_seg_and_init():
    jm, ev = reference_shim.load()
    torch.manual_seed(7)
    seg = jm.Segmentation(1, 2, norm_type=1)
    torch.manual_seed(7)
    sd = R.init_seg_state()
    ref_sd = seg.state_dict()
    assert list(ref_sd.keys()) == list(sd.keys())
    assert all(torch.equal(ref_sd[k], sd[k]) for k in sd)
    x = torch.randn(1, 1, 32, 32, 32).clamp(-1, 1)
    with torch.no_grad():
        ref = seg({"i": x}, "i", "o")["o"]
        mine = R.seg_forward(sd, x)
    assert torch.equal(ref, mine)


@needs_ref
@pytest.mark.timeout(600)
def test_oracle_equals_real_reference_vae_and_losses():
    jm, ev = reference_shim.load()
    torch.manual_seed(8)
    vae = jm.VAE(2, 2, norm_type=1, dim=128)
    torch.manual_seed(8)
    sd = R.init_vae_state(2, 128, 128)
    assert all(torch.equal(v, sd[k]) for k, v in vae.state_dict().items())
    x = R.one_hot((torch.rand(1, 1, 128, 128, 128) > 0.9).float())
    with torch.no_grad():
        torch.manual_seed(9)
        ref, rm, rs = vae(x, if_random=True, scale=0.35)
        torch.manual_seed(9)
        mine, mm, ms = R.vae_forward(sd, x, if_random=True, scale=0.35)
    assert torch.equal(ref, mine) and torch.equal(rm, mm) and torch.equal(rs, ms)
    d = {"a": ref, "b": x, "mean": rm, "std": rs}
    assert torch.equal(ev.avg_dsc(d, "a", "b", botindex=1, topindex=2), R.avg_dsc(ref, x, botindex=1, topindex=2))
    assert torch.equal(ev.KLloss(d), R.kl_loss(rm, rs))
    # generalised flat dim: P=64 runs through the restatement (the real VAE cannot)
    sd64 = R.init_vae_state(2, 128, 64)
    out, _, _ = R.vae_forward(sd64, R.one_hot((torch.rand(1, 1, 64, 64, 64) > 0.9).float()))
    assert out.shape == (1, 2, 64, 64, 64)


def test_clip_center_oracle_matches_reference_statements():
    """oracle clip_center == the reference's two transforms written out (utils/utils.py:508-533,572-618)."""
    import numpy as np
    from oracle import ref_torch as R
    rng = np.random.RandomState(0)
    x = (rng.randn(3, 4, 5) * 500).astype(np.float32)
    val = np.clip(x.copy(), -200, 400)            # Clip.__call__
    val = val.reshape((val.shape[0], -1))         # CenterIntensities.__call__
    val -= 100
    val /= 300
    assert np.array_equal(R.clip_center(x), val.reshape(x.shape))
    assert R.clip_center(np.array([-5000, 100, 5000], dtype=np.int16)).tolist() == [-1.0, 0.0, 1.0]


def test_conditioned_recipe_matches_real_reference_golden(golden_dir):
    """oracle/conditioned.py (weights after K SGD steps of the reference's train step, the fixture of the bf16 parity
    tests) against the same recipe run through the REAL reference modules (tests/golden/conditioned.npz, written by
    oracle/make_golden.py): loss trajectory and per-parameter (sum, l2) checksums of the trained weights."""
    import numpy as np
    from oracle import conditioned as C
    from oracle.make_golden import COND_CASE
    g = np.load(os.path.join(golden_dir, "conditioned.npz"))
    K, lr = COND_CASE["steps"], COND_CASE["lr"]
    sd, losses = C.train_seg(K, patch=COND_CASE["patch_seg"], lr=lr)
    np.testing.assert_allclose(losses, g["seg_losses"], rtol=2e-4)
    np.testing.assert_allclose(C.checksum(sd)[:, 1], g["seg_checksum"][:, 1], rtol=1e-4)
    vd, vlosses = C.train_vae(K, patch=COND_CASE["patch_vae"], lr=lr)
    np.testing.assert_allclose(vlosses, g["vae_losses"], rtol=2e-4)
    np.testing.assert_allclose(C.checksum(vd)[:, 1], g["vae_checksum"][:, 1], rtol=1e-4)


def test_resize_restatement_basic_properties():
    """oracle.ref_resize.resize (skimage 0.18.3's n-D path over scipy.ndimage): same-size resize is the identity, constants
    stay constant, 2x linear upsampling of a ramp interpolates at the half-sample positions, nearest picks floor(x + .5)."""
    import numpy as np
    from oracle import ref_resize as RR
    rng = np.random.RandomState(0)
    v = rng.randn(6, 7, 5).astype(np.float32)
    assert np.allclose(RR.resize(v, v.shape), v, atol=1e-6)
    assert np.allclose(RR.resize(np.full((9, 9, 9), 3.5, np.float32), (4, 6, 13)), 3.5, atol=1e-6)
    ramp = np.tile(np.arange(4, dtype=np.float32)[:, None, None], (1, 2, 2))
    up = RR.resize(ramp, (8, 2, 2))[:, 0, 0]
    assert np.allclose(up, [0.25, 0.25, 0.75, 1.25, 1.75, 2.25, 2.75, 2.75], atol=1e-6)      # x = .5 (i + .5) - .5, mirror edges
    lab = (rng.rand(5, 5, 5) > 0.5).astype(np.float32)
    near = RR.resize(lab, (10, 10, 10), order=0, anti_aliasing=False)
    assert set(np.unique(near)) <= {0.0, 1.0} and np.array_equal(near[::2, ::2, ::2], lab)     # x = .5 i - .25 -> floor(.5 i + .25)


@needs_ref
@pytest.mark.timeout(600)
def test_oracle_equals_real_reference_encoder_and_fusion():
    """The remaining model variants (joint_model.py:274-305 Encoder, :392-436 Fusion): functional restatements against
    the real modules on the same state_dict."""
    jm, ev = reference_shim.load()
    torch.manual_seed(12)
    enc = jm.Encoder(1, 1, norm_type=1)
    x = torch.rand(1, 1, 128, 128, 128)
    with torch.no_grad():
        assert torch.equal(enc(x), R.encoder_forward(enc.state_dict(), x))
    fus = jm.Fusion(1, 2, 2, norm_type=1)
    img, mask = torch.randn(1, 1, 32, 32, 32), R.one_hot((torch.rand(1, 1, 32, 32, 32) > 0.9).float())
    with torch.no_grad():
        assert torch.equal(fus({"i": img, "m": mask}, "i", "m", "o")["o"], R.fusion_forward(fus.state_dict(), img, mask))


@needs_ref
def test_oracle_batchnorm_path_equals_real_reference():
    """norm_type=2: the oracle's F.batch_norm branch against the reference's nn.BatchNorm3d-based Segmentation, training
    mode (forward, every gradient, running-statistics update) and eval mode (forward)."""
    jm, ev = reference_shim.load()
    torch.manual_seed(5)
    m = jm.Segmentation(1, 2, norm_type=2)
    with torch.no_grad():
        for k, p in m.named_parameters():
            if k.rsplit(".", 2)[-2] in ("1", "4", "7") and "conv" in k:          # BatchNorm affine parameters
                p.copy_(0.5 + torch.rand_like(p) if k.endswith("weight") else 0.2 * torch.randn_like(p))
    sd = OrderedDict((k, v.clone()) for k, v in m.state_dict().items())
    x, lab = torch.randn(2, 1, 32, 32, 32), (torch.rand(2, 1, 32, 32, 32) > 0.7).float()
    m.train()
    pred = m({"img": x}, "img", "pred")["pred"]
    onehot = R.one_hot(lab)
    loss = 1 - ev.avg_dsc({"p": pred, "t": onehot}, source_key="p", target_key="t", botindex=1, topindex=2)
    loss.backward()
    rsd = R._leafify(sd)
    pred_o = R.seg_forward(rsd, x)
    loss_o = 1 - R.avg_dsc(pred_o, onehot, botindex=1, topindex=2)
    loss_o.backward()
    assert torch.equal(pred, pred_o) and torch.equal(loss, loss_o)
    for k, p in m.named_parameters():
        assert torch.allclose(p.grad, rsd[k].grad, rtol=1e-6, atol=1e-9), k
    new = m.state_dict()
    for k in sd:
        if k.endswith("running_mean") or k.endswith("running_var"):
            assert torch.equal(new[k], rsd[k].detach()), k
    m.eval()
    R.BN_EVAL = True
    try:
        with torch.no_grad():
            assert torch.equal(m({"img": x}, "img", "pred")["pred"], R.seg_forward(R._leafify(new, requires_grad=False), x))
    finally:
        R.BN_EVAL = False
