"""CPU-side checks of the drop-in module surface: state_dict layout, initialisation stream,
layer programs (no kernels run here)."""
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
