"""CPU-side checks of the drop-in module surface: state_dict layout, initialisation stream,
layer programs (no kernels run here)."""
import pytest
import torch

from oracle import ref_torch as R
from vae_segmentation_b200 import engine, joint_model


def test_segmentation_state_dict_matches_reference_layout_and_init():
    torch.manual_seed(3)
    seg = joint_model.Segmentation(1, 2, norm_type=1)
    torch.manual_seed(3)
    sd = R.init_seg_state()
    msd = seg.state_dict()
    assert list(msd.keys()) == list(sd.keys()) and len(msd) == 68
    assert all(torch.equal(msd[k], sd[k]) for k in sd)
    assert sum(p.numel() for p in seg.parameters()) == 2276018
    seg.load_state_dict(sd, strict=True)


def test_vae_state_dict_matches_reference_layout_and_init():
    torch.manual_seed(4)
    vae = joint_model.VAE(2, 2, norm_type=1, dim=128)
    torch.manual_seed(4)
    sd = R.init_vae_state(2, 128, 128)
    msd = vae.state_dict()
    assert list(msd.keys()) == list(sd.keys()) and len(msd) == 90
    assert all(torch.equal(msd[k], sd[k]) for k in sd)
    assert sum(p.numel() for p in vae.parameters()) == 15434378
    assert msd["fc_mean.weight"].shape == (128, 16384)
    assert joint_model.VAE(2, 2, norm_type=1, dim=128, patch=96).fc2.weight.shape == (6912, 128)


def test_joint_wrapper_attribute_names():
    j = joint_model.Joint([joint_model.Segmentation(1, 2, norm_type=1), joint_model.VAE(2, 2, norm_type=1, dim=128)])
    assert len(j.state_dict()) == 158
    assert all(k.startswith(("Seg.", "Vae.")) for k in j.state_dict())
    names = [n for n, _ in j.Vae.named_parameters()]
    assert any(n.startswith("fc_mean") for n in names) and any(n.startswith("down5") for n in names)
    assert hasattr(j.Seg, "up5") and hasattr(j.Seg, "out_block")


def test_layer_programs():
    seg = joint_model.Segmentation(1, 2, norm_type=1)
    layers, params = seg._prog()
    kinds = [l.kind for l in layers]
    assert kinds.count(engine.C3IN) == 25 and kinds.count(engine.K2DOWN) == 4
    assert kinds.count(engine.K2UP) == 4 and kinds[-1] == engine.HEAD
    assert [l.save_as for l in layers if l.save_as] == ["x2", "x3"]
    assert [l.skip_from for l in layers if l.skip_from] == ["x3", "x2"]
    for l in layers:
        w = params[l.wi]
        if l.kind in (engine.C3IN, engine.HEAD):
            assert tuple(w.shape) == (l.cout, l.cin, 3, 3, 3)
        else:
            assert tuple(w.shape) == (l.cin, l.cout, 2, 2, 2) and l.cin == l.cout
    vae = joint_model.VAE(2, 2, norm_type=1, dim=128, patch=64)
    (enc, dec, fc), vparams = vae._prog()
    assert len(enc) == 1 + 5 * 4 and len(dec) == 5 * 4 + 1
    assert tuple(vparams[fc[0]].shape) == (128, 2048)


def test_variant_modules_state_dict_matches_the_real_reference():
    """Encoder / Fusion / Joint2 / Embed (joint_model.py:274-305,392-501): same state_dict keys and shapes as the real
    modules, so reference checkpoints load strictly."""
    import pytest
    from oracle import reference_shim
    if not reference_shim.available():
        pytest.skip("reference tree not present")
    rjm, _ = reference_shim.load()
    pairs = [(joint_model.Encoder(1, 1, norm_type=1), rjm.Encoder(1, 1, norm_type=1)),
             (joint_model.Fusion(1, 2, 2, norm_type=1), rjm.Fusion(1, 2, 2, norm_type=1))]
    for ours, ref in pairs:
        a, b = ours.state_dict(), ref.state_dict()
        assert list(a.keys()) == list(b.keys())
        assert all(a[k].shape == b[k].shape for k in a)
        ours.load_state_dict(b, strict=True)
    j2 = joint_model.Joint2([joint_model.Segmentation(1, 2, norm_type=1), joint_model.Encoder(1, 1, norm_type=1)])
    r2 = rjm.Joint2([rjm.Segmentation(1, 2, norm_type=1), rjm.Encoder(1, 1, norm_type=1)])
    assert list(j2.state_dict().keys()) == list(r2.state_dict().keys())
    em = joint_model.Embed([joint_model.Encoder(1, 128, norm_type=1), joint_model.VAE(2, 2, norm_type=1, dim=128),
                            joint_model.Fusion(1, 2, 2, norm_type=1)])
    rm = rjm.Embed([rjm.Encoder(1, 128, norm_type=1), rjm.VAE(2, 2, norm_type=1, dim=128), rjm.Fusion(1, 2, 2, norm_type=1)])
    assert list(em.state_dict().keys()) == list(rm.state_dict().keys())


def test_batchnorm_variant_state_dict_matches_reference():
    """norm_type=2 (the constructors' default): same keys, shapes, dtypes and initial values as the reference's
    nn.BatchNorm3d-based modules, and reference state loads strictly."""
    import os
    import pytest
    if not os.path.isdir("/root/reference"):
        pytest.skip("reference tree not present")
    from oracle import reference_shim
    rjm, _ = reference_shim.load()
    torch.manual_seed(3)
    for ours, ref in ((joint_model.Segmentation(1, 2), rjm.Segmentation(1, 2)),
                      (joint_model.VAE(2, 2, dim=128), rjm.VAE(2, 2, dim=128))):
        a, b = ours.state_dict(), ref.state_dict()
        assert list(a.keys()) == list(b.keys())
        for k in a:
            assert a[k].shape == b[k].shape and a[k].dtype == b[k].dtype, k
            if k.rsplit(".", 1)[-1] in ("running_mean", "running_var", "num_batches_tracked") or \
                    (k.rsplit(".", 2)[-2] in ("1", "4", "7") and "conv" in k):
                assert torch.equal(a[k], b[k]), k                                  # BatchNorm init: ones / zeros / 0
        ours.load_state_dict(b, strict=True)


@pytest.mark.parametrize("training", [True, False])
def test_batchnorm_host_tables_reproduce_torch_batch_norm(training):
    """engine._bn_forward_tables / _bn_backward_tables (the host-side part of the norm_type=2 path: pooling of the
    per-(n,c) convolution statistics, running-statistics update, folding of gamma / beta / shift into the coefficient
    tables of csrc/affine_act.cu) against torch's batch_norm + ReLU with autograd.  The three affine kernels are
    restated in torch here (they are elementwise); everything else is the product code, run on CPU tensors."""
    import torch.nn.functional as F
    from vae_segmentation_b200 import engine
    from vae_segmentation_b200.joint_model import _BatchNormParams
    torch.manual_seed(17 + int(training))
    n, c, S = 3, 8, 5 * 6 * 7
    bias = torch.randn(c, dtype=torch.float64)                       # conv bias: y_true = y_nobias + bias
    y_nobias = (torch.randn(n, c, S, dtype=torch.float64) * 2.0 + 0.7)
    shift = y_nobias[:, :, 1].clone()                                # stored output = y_nobias - shift[n,c]
    y_s = y_nobias - shift[..., None]
    stats = torch.stack([y_s.sum(-1), (y_s ** 2).sum(-1)], -1)      # what the convolution kernels emit
    bn = _BatchNormParams(c).double()
    with torch.no_grad():
        bn.weight.copy_(0.5 + torch.rand(c)); bn.weight[0] = -0.7
        bn.bias.copy_(0.3 * torch.randn(c))
        bn.running_mean.copy_(0.2 * torch.randn(c)); bn.running_var.copy_(0.5 + torch.rand(c))
    bn.train(training)
    rm0, rv0 = bn.running_mean.clone(), bn.running_var.clone()
    L = engine.Layer(engine.C3IN, "t", c, c, 0, 1, bn=bn, gi=2, bti=3)
    tensors = [None, bias, bn.weight, bn.bias]

    # reference: torch batch_norm on the TRUE conv output (with bias), ReLU, upstream gradient g
    x = (y_nobias + bias[None, :, None]).clone().requires_grad_(True)
    gam, bet = bn.weight.detach().clone().requires_grad_(True), bn.bias.detach().clone().requires_grad_(True)
    rm, rv = rm0.clone(), rv0.clone()
    out_ref = F.relu(F.batch_norm(x, rm, rv, gam, bet, training=training, momentum=0.1, eps=1e-5))
    g = torch.randn_like(out_ref)
    out_ref.backward(g)

    kb, k2b2, aux = engine._bn_forward_tables(L, tensors, stats, shift.float(), float(S))
    kb, k2b2 = kb.double(), k2b2.double()
    pre = kb[..., 0, None] * y_s + kb[..., 1, None]                  # vs_affine_relu_apply
    assert torch.allclose(pre.clamp_min(0), out_ref.detach(), rtol=1e-5, atol=1e-5)
    assert torch.allclose(bn.running_mean, rm, rtol=1e-6, atol=1e-6) and torch.allclose(bn.running_var, rv, rtol=1e-6, atol=1e-6)
    assert int(bn.num_batches_tracked) == (1 if training else 0)
    gm = g * (pre > 0)                                               # vs_affine_relu_bwd_reduce
    xhat = k2b2[..., 0, None] * y_s + k2b2[..., 1, None]
    sums = torch.stack([gm.sum(-1), (gm * xhat).sum(-1)], -1)
    coef, dgamma, dbeta, dbias = engine._bn_backward_tables(L, sums, kb.float(), k2b2.float(), aux, float(S), training)
    coef = coef.double()
    dy = coef[..., 0, None] * gm + coef[..., 1, None] + coef[..., 2, None] * y_s      # vs_affine_relu_bwd_apply
    assert torch.allclose(dy, x.grad, rtol=1e-4, atol=1e-6)
    assert torch.allclose(dgamma.double(), gam.grad, rtol=1e-5, atol=1e-6) and torch.allclose(dbeta.double(), bet.grad, rtol=1e-5, atol=1e-6)
    # conv bias: x = y_nobias + bias, so dbias = sum over (n, voxels) of dx -- exactly zero under batch statistics
    assert torch.allclose(dbias.double(), x.grad.sum((0, 2)), rtol=1e-4, atol=1e-5)
