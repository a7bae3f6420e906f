"""tcgen05/TMEM/TMA implicit-GEMM convolution against torch.nn.functional (CPU fp32) and against the
CUDA-core direct kernel on identical bf16 operands: fprop (+ shifted output + fp64 statistics) and dgrad."""
import pytest
import torch
import torch.nn.functional as F

from vae_segmentation_b200 import ops

pytestmark = pytest.mark.gpu
DEV = "cuda"

# (n, d, h, w, cin, cout): ragged extents exercise partial tiles (tile = 4 x 16 x 8) and TMA zero fill
CASES = [(1, 4, 16, 8, 16, 16), (2, 5, 17, 9, 16, 16), (1, 7, 20, 19, 8, 8), (2, 6, 9, 11, 8, 16), (1, 8, 18, 10, 16, 8),
         (1, 9, 12, 12, 32, 32), (1, 6, 6, 6, 64, 64), (2, 3, 3, 3, 128, 64), (1, 6, 6, 6, 64, 128), (1, 5, 7, 6, 32, 16),
         (1, 3, 3, 3, 256, 256), (1, 12, 24, 24, 16, 32)]


def to_ndhwc(x):
    return x.permute(0, 2, 3, 4, 1).contiguous().to(DEV, torch.bfloat16)


def from_ndhwc(x):
    return x.float().cpu().permute(0, 4, 1, 2, 3).contiguous()


@pytest.mark.parametrize("case", CASES)
def test_tc_fprop_and_dgrad(case):
    if not ops.has_tcgen05():
        pytest.fail("library built without the tcgen05 kernels")
    n, d, h, w, cin, cout = case
    torch.manual_seed(sum(case))
    x = torch.randn(n, cin, d, h, w).bfloat16().float()
    wt = (torch.randn(cout, cin, 3, 3, 3) * 0.1)
    wq = wt.bfloat16().float()                       # the tensor-core path rounds weights to bf16
    y_ref = F.conv3d(x, wq, None, padding=1)
    wd = wt.to(DEV)
    wf, wdg = ops.pack_conv3_weight(wd)
    wtc = ops.pack_conv3_weight_tc(wd, dgrad=False)
    wdtc = ops.pack_conv3_weight_tc(wd, dgrad=True)
    assert wtc is not None and wdtc is not None
    xd = to_ndhwc(x)
    dims = (n, d, h, w)
    y, stats = ops.conv3_fprop(xd, wf, None, dims, cin, cout, torch.bfloat16, shifted=False, wtc=wtc)
    torch.cuda.synchronize()
    got = from_ndhwc(y)
    scale = y_ref.abs().max().item()
    assert (got - y_ref).abs().max().item() < 1e-2 * scale, "tc fprop max err %.3e (scale %.3e)" % ((got - y_ref).abs().max().item(), scale)
    s_ref = torch.stack([y_ref.double().sum((2, 3, 4)), (y_ref.double() ** 2).sum((2, 3, 4))], -1)
    assert torch.allclose(stats.cpu(), s_ref, rtol=2e-3, atol=2e-3 * s_ref.abs().max().item())
    # shifted variant (what the InstanceNorm layers use)
    ys, st2 = ops.conv3_fprop(xd, wf, None, dims, cin, cout, torch.bfloat16, shifted=True, wtc=wtc)
    ys_ref = y_ref - F.conv3d(x, wt, None, padding=1)[:, :, 1:2, 1:2, 1:2]
    assert (from_ndhwc(ys) - ys_ref).abs().max().item() < 1.5e-2 * scale
    # dgrad
    gy = torch.randn_like(y_ref).bfloat16().float()
    dx_ref = F.conv_transpose3d(gy, wq, None, padding=1)
    dx = ops.conv3_dgrad(to_ndhwc(gy), wdg, dims, cin, cout, torch.bfloat16, wdtc=wdtc)
    torch.cuda.synchronize()
    gscale = dx_ref.abs().max().item()
    assert (from_ndhwc(dx) - dx_ref).abs().max().item() < 1e-2 * gscale
    # same operands through the CUDA-core kernel agree to bf16 output rounding
    y2, _ = ops.conv3_fprop(xd, wq.to(DEV).reshape(cout, cin, 27).permute(2, 1, 0).contiguous(), None, dims, cin, cout,
                            torch.bfloat16, shifted=False, wtc=None)
    assert (y2.float() - y.float()).abs().max().item() < 1e-2 * scale


@pytest.mark.parametrize("case", CASES + [(2, 8, 32, 16, 8, 8), (1, 4, 16, 8, 16, 128), (1, 5, 6, 7, 128, 256),
                                          (1, 1, 5, 5, 8, 8), (1, 3, 3, 3, 8, 8), (2, 2, 17, 9, 32, 8), (1, 13, 33, 17, 16, 8),
                                          (1, 24, 24, 24, 8, 8)])
def test_tc_wgrad(case):
    """tcgen05 backward-filter (voxel = K, MN-major operands) against torch autograd on the same bf16 operands;
    ragged extents exercise the TMA zero fill of BOTH operands, Cout > 64 the 64-row chunks, accumulate=True the
    add-onto semantics."""
    n, d, h, w, cin, cout = case
    torch.manual_seed(sum(case) + 1)
    x = torch.randn(n, cin, d, h, w).bfloat16().float()
    gy = torch.randn(n, cout, d, h, w).bfloat16().float()
    wt = torch.zeros(cout, cin, 3, 3, 3, requires_grad=True)
    F.conv3d(x, wt, None, padding=1).backward(gy)
    want = wt.grad
    dims = (n, d, h, w)
    xd, gyd = to_ndhwc(x), to_ndhwc(gy)
    dw, _ = ops.conv3_wgrad(xd, gyd, dims, cin, cout)
    torch.cuda.synchronize()
    scale = want.abs().max().item()
    err = (dw.cpu() - want).abs().max().item()
    assert err < 2e-3 * scale, "tc wgrad max err %.3e (scale %.3e)" % (err, scale)
    dw2, _ = ops.conv3_wgrad(xd, gyd, dims, cin, cout, dw=dw.clone(), accumulate=True)
    assert (dw2.cpu() - 2 * want).abs().max().item() < 4e-3 * scale


@pytest.mark.parametrize("case", CASES)
def test_tc_dgrad_fused_norm_backward_reduction(case):
    """dgrad epilogue that also accumulates the previous layer's InstanceNorm+ReLU backward sums
    (sum g*mask, sum g*mask*xhat) against the standalone reduction kernel run on the dgrad's own output,
    and the resulting dy against torch autograd of relu(instance_norm(y_prev))."""
    n, d, h, w, cin, cout = case
    torch.manual_seed(sum(case) + 2)
    dims = (n, d, h, w)
    wt = torch.randn(cout, cin, 3, 3, 3) * 0.1
    wd = wt.to(DEV)
    _, wdg = ops.pack_conv3_weight(wd)
    wdtc = ops.pack_conv3_weight_tc(wd, dgrad=True)
    gy = torch.randn(n, cout, d, h, w).bfloat16().float()
    yprev = (torch.randn(n, cin, d, h, w) * 1.5 + 0.3).bfloat16().float()       # previous layer's raw output
    stats = torch.stack([yprev.double().sum((2, 3, 4)), (yprev.double() ** 2).sum((2, 3, 4))], -1).to(DEV)
    ypd, gyd = to_ndhwc(yprev), to_ndhwc(gy)
    assert ops.dgrad_can_fuse_reduce(gyd, cin, cout, torch.bfloat16, False, wdtc)
    arena = ops.StatsArena(n * cin * 2, DEV)
    sums = arena.take(n * cin * 2).view(n, cin, 2)
    dx = ops.conv3_dgrad(gyd, wdg, dims, cin, cout, torch.bfloat16, wdtc=wdtc, prev=(ypd, stats, sums))
    dx_plain = ops.conv3_dgrad(gyd, wdg, dims, cin, cout, torch.bfloat16, wdtc=wdtc)
    torch.cuda.synchronize()
    assert torch.equal(dx, dx_plain), "the fused epilogue must not change the dgrad output"
    # reference sums from the standalone kernel on the same dx
    import vae_segmentation_b200._cabi as cabi
    sums_ref = torch.empty(n, cin, 2, device=DEV, dtype=torch.float64)
    cabi.call("vs_inorm_relu_bwd_reduce", cabi.VS_BF16, dx.data_ptr(), ypd.data_ptr(), stats.data_ptr(), sums_ref.data_ptr(),
              n, d * h * w, cin, 0, torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    scale = sums_ref.abs().max().item() + 1e-6
    # fp32 per-thread partial sums (a few hundred terms) vs the standalone kernel's fp64 accumulation
    assert (sums - sums_ref).abs().max().item() < 2e-5 * scale + 1e-4, \
        "fused sums differ: %.3e (scale %.3e)" % ((sums - sums_ref).abs().max().item(), scale)
    dy_fused = ops.inorm_relu_bwd(dx, ypd, stats, sums=sums, reduced=True)
    dy_split = ops.inorm_relu_bwd(dx, ypd, stats)
    yr = yprev.clone().requires_grad_()
    F.relu(F.instance_norm(yr, eps=1e-5)).backward(from_ndhwc(dx))
    ref = yr.grad
    gscale = ref.abs().max().item()
    assert (from_ndhwc(dy_fused) - ref).abs().max().item() < 2e-2 * gscale
    assert (dy_fused.float() - dy_split.float()).abs().max().item() < 1e-2 * gscale


@pytest.mark.parametrize("case", [(2, 5, 17, 9, 8), (1, 8, 16, 8, 16), (1, 7, 20, 19, 8), (2, 3, 3, 3, 32)])
def test_tc_head_conv_bias_softmax(case):
    """The 2-class head in one tensor-core launch (output channels padded to 8, bias + softmax + planar fp32 store in
    the epilogue) against softmax(conv3d + bias) of torch on the same bf16 operands."""
    n, d, h, w, cin = case
    torch.manual_seed(sum(case) + 3)
    x = torch.randn(n, cin, d, h, w).bfloat16().float()
    wt = torch.randn(2, cin, 3, 3, 3) * 0.2
    b = torch.randn(2)
    ref = F.softmax(F.conv3d(x, wt.bfloat16().float(), b, padding=1), dim=1)
    wtc8 = ops.pack_conv3_weight_tc_padded(wt.to(DEV), cin, 8, dgrad=False)
    probs = ops.head_conv_softmax2(to_ndhwc(x), wtc8, b.to(DEV), (n, d, h, w), cin)
    torch.cuda.synchronize()
    assert probs.shape == ref.shape and probs.dtype == torch.float32
    assert (probs.cpu() - ref).abs().max().item() < 2e-3
    assert (probs.cpu().argmax(1) == ref.argmax(1)).float().mean().item() > 0.999


@pytest.mark.parametrize("case", [(2, 5, 17, 9), (1, 8, 16, 8), (1, 3, 3, 3)])
def test_tc_inblock2_planar_dgrad(case):
    """Planar fp32 2-channel input gradient of an in-block (2 -> 8) on the tensor cores: the dgrad pack has its input
    channels zero-padded to 8, the epilogue stores channels 0..1 as planar fp32."""
    n, d, h, w = case
    torch.manual_seed(sum(case) + 4)
    wt = torch.randn(8, 2, 3, 3, 3) * 0.2
    gy = torch.randn(n, 8, d, h, w).bfloat16().float()
    ref = F.conv_transpose3d(gy, wt.bfloat16().float(), None, padding=1)
    wdtc = ops.pack_conv3_weight_tc_padded(wt.to(DEV), 8, 8, dgrad=True)
    _, wdg = ops.pack_conv3_weight(wt.to(DEV))
    dx = ops.conv3_dgrad(to_ndhwc(gy), wdg, (n, d, h, w), 2, 8, torch.bfloat16, out_planar=True, wdtc=wdtc)
    torch.cuda.synchronize()
    assert dx.shape == ref.shape and dx.dtype == torch.float32
    assert (dx.cpu() - ref).abs().max().item() < 2e-3 * ref.abs().max().item()


@pytest.mark.parametrize("case", [(2, 5, 17, 9, 8), (1, 8, 16, 8, 16), (1, 7, 20, 19, 8), (1, 4, 33, 12, 16)])
def test_tc_kdn_planar_head_and_inblock_dgrad(case):
    """kd-in-N kernel with the planar fp32 epilogue: mode 1 = 2-class head (weights zero-padded to 8 output channels,
    bias + softmax fused), mode 2 = the 2-channel planar input gradient of an in-block (input channels padded to 8);
    plus the head's 8-channel dgrad through the padded dgrad pack.  All against torch on the same bf16 operands."""
    n, d, h, w, cin = case
    torch.manual_seed(sum(case) + 11)
    x = torch.randn(n, cin, d, h, w).bfloat16().float()
    wt = torch.randn(2, cin, 3, 3, 3) * 0.2
    b = torch.randn(2)
    ref = F.softmax(F.conv3d(x, wt.bfloat16().float(), b, padding=1), dim=1)
    wk8 = ops.pack_conv3_weight_tc_kdn_padded(wt.to(DEV), cin, 8, dgrad=False)
    assert wk8 is not None
    probs = ops.conv3_tc_kdn_planar(to_ndhwc(x), wk8, (n, d, h, w), cin, 1, bias=b.to(DEV))
    torch.cuda.synchronize()
    assert probs.shape == ref.shape and probs.dtype == torch.float32
    assert (probs.cpu() - ref).abs().max().item() < 2e-3
    assert (probs.cpu().argmax(1) == ref.argmax(1)).float().mean().item() > 0.999
    # head dgrad: 8-channel padded logit gradient (channels 2..7 arbitrary: their weights are zero) -> Cin channels
    gl = torch.randn(n, 8, d, h, w).bfloat16().float()
    dref = F.conv_transpose3d(gl[:, :2], wt.bfloat16().float(), None, padding=1)
    wkd8 = ops.pack_conv3_weight_tc_kdn_padded(wt.to(DEV), cin, 8, dgrad=True)
    assert wkd8 is not None
    dx, _ = ops.conv3_tc_kdn(to_ndhwc(gl), wkd8, (n, d, h, w), 8, cin)
    torch.cuda.synchronize()
    assert (from_ndhwc(dx) - dref).abs().max().item() < 1e-2 * dref.abs().max().item()
    if cin == 8:
        # in-block 2 -> 8: planar 2-channel input gradient
        wi = torch.randn(8, 2, 3, 3, 3) * 0.2
        gy = torch.randn(n, 8, d, h, w).bfloat16().float()
        iref = F.conv_transpose3d(gy, wi.bfloat16().float(), None, padding=1)
        wki = ops.pack_conv3_weight_tc_kdn_padded(wi.to(DEV), 8, 8, dgrad=True)
        dxi = ops.conv3_tc_kdn_planar(to_ndhwc(gy), wki, (n, d, h, w), 8, 2)
        torch.cuda.synchronize()
        assert dxi.shape == iref.shape and dxi.dtype == torch.float32
        assert (dxi.cpu() - iref).abs().max().item() < 2e-3 * iref.abs().max().item()


def test_batched_repack_matches_per_layer_packs():
    """engine.PackCache.repack_all (one launch for every derived pack, incl. the padded kd-in-N packs of the head and the
    2-channel in-block) writes exactly what the per-layer pack entry points write."""
    from vae_segmentation_b200 import engine
    torch.manual_seed(5)
    cache = engine.PackCache()
    w_head = (torch.randn(2, 8, 3, 3, 3) * 0.3).to(DEV)
    w_in = (torch.randn(8, 2, 3, 3, 3) * 0.3).to(DEV)
    w_c = (torch.randn(16, 8, 3, 3, 3) * 0.3).to(DEV)
    first = {"hk": cache.head_kdn(w_head), "ht": cache.head_tc(w_head), "ik": cache.inblock2_dgrad_kdn(w_in),
             "it": cache.inblock2_dgrad_tc(w_in), "ck": cache.conv3_kdn(w_c), "ckd": cache.conv3_kdn(w_c, dgrad=True)}
    flat = lambda d: [t for v in d.values() for t in (v if isinstance(v, tuple) else (v,))]
    assert all(t is not None for t in flat(first))
    before = [t.clone() for t in flat(first)]
    for wt in (w_head, w_in, w_c):
        wt.mul_(-1.5)                                  # new weights, same storage
    cache.repack_all()
    torch.cuda.synchronize()
    fresh = engine.PackCache()
    want = {"hk": fresh.head_kdn(w_head), "ht": fresh.head_tc(w_head), "ik": fresh.inblock2_dgrad_kdn(w_in),
            "it": fresh.inblock2_dgrad_tc(w_in), "ck": fresh.conv3_kdn(w_c), "ckd": fresh.conv3_kdn(w_c, dgrad=True)}
    for got, exp, old in zip(flat(first), flat(want), before):
        assert torch.equal(got, exp)
        assert not torch.equal(got, old)


FUSED_CASES = [(2, 5, 17, 9, 8, 8, True), (1, 8, 18, 10, 16, 8, True), (2, 6, 9, 11, 16, 16, True), (1, 12, 24, 24, 32, 16, True),
               (2, 5, 17, 9, 8, 8, False), (2, 6, 6, 6, 128, 128, False), (1, 12, 12, 12, 64, 32, False), (2, 3, 3, 3, 256, 64, False),
               (1, 24, 24, 24, 32, 32, False), (1, 48, 40, 56, 8, 8, True)]


@pytest.mark.parametrize("with_skip", [False, True])
@pytest.mark.parametrize("case", FUSED_CASES)
def test_tc_conv_instance_norm_relu_in_one_launch(case, with_skip):
    """Conv3d -> InstanceNorm3d -> ReLU (+ skip) as ONE cooperative launch with an in-kernel grid barrier
    (vs_conv3x3x3_tc[_kdn]_in_relu) against torch on the same bf16 operands and against conv + separate apply pass."""
    n, d, h, w, cin, cout, kdn = case
    torch.manual_seed(sum(case[:6]) + 23)
    x = torch.randn(n, cin, d, h, w).bfloat16().float()
    wt = torch.randn(cout, cin, 3, 3, 3) * 0.1
    skip = torch.randn(n, cout, d, h, w).bfloat16().float() if with_skip else None
    ref = F.relu(F.instance_norm(F.conv3d(x, wt.bfloat16().float(), None, padding=1), eps=1e-5))
    if skip is not None:
        ref = ref + skip
    wd = wt.to(DEV)
    dims = (n, d, h, w)
    wpack = ops.pack_conv3_weight_tc_kdn(wd, dgrad=False) if kdn else ops.pack_conv3_weight_tc(wd, dgrad=False)
    assert wpack is not None
    skd = to_ndhwc(skip) if skip is not None else None
    arena = ops.StatsArena(ops.stats_words(n, cout), DEV)
    y, stats, a = ops.conv3_in_relu(to_ndhwc(x), wpack, dims, cin, cout, arena, skip=skd, kdn=kdn)
    torch.cuda.synchronize()
    assert (from_ndhwc(a) - ref).abs().max().item() < 3e-2 * max(1.0, ref.abs().max().item())
    # the two-launch path on the same operands: same raw output and statistics (up to summation order), same activation
    arena2 = ops.StatsArena(ops.stats_words(n, cout), DEV)
    if kdn:
        y2, stats2 = ops.conv3_tc_kdn(to_ndhwc(x), wpack, dims, cin, cout, want_stats=True, arena=arena2)
    else:
        wf, _ = ops.pack_conv3_weight(wd)
        y2, stats2 = ops.conv3_fprop(to_ndhwc(x), wf, None, dims, cin, cout, torch.bfloat16, wtc=wpack, arena=arena2)
    a2 = ops.inorm_relu_apply(y2, stats2, skd)
    torch.cuda.synchronize()
    scale = y2.float().abs().max().item()
    assert (y.float() - y2.float()).abs().max().item() <= 1e-2 * scale           # kd-in-N issue order: a bf16 ulp at most
    assert torch.allclose(stats, stats2, rtol=1e-3, atol=1e-3 * stats2.abs().max().item())
    assert (a.float() - a2.float()).abs().max().item() <= 2e-2 * max(1.0, a2.float().abs().max().item())
    # a second launch reusing nothing: the barrier word is consumed per launch
    arena3 = ops.StatsArena(ops.stats_words(n, cout), DEV)
    _, _, a3 = ops.conv3_in_relu(to_ndhwc(x), wpack, dims, cin, cout, arena3, skip=skd, kdn=kdn)
    torch.cuda.synchronize()
    assert (a3.float() - a.float()).abs().max().item() <= 2e-2 * max(1.0, a2.float().abs().max().item())


KDN_CASES = [(1, 4, 16, 8, 8, 8), (2, 5, 17, 9, 8, 8), (1, 8, 18, 10, 16, 8), (1, 7, 20, 19, 8, 16), (2, 6, 9, 11, 16, 16),
             (1, 12, 24, 24, 32, 16)]


@pytest.mark.parametrize("ordered", [0, 1])
@pytest.mark.parametrize("case", KDN_CASES)
def test_tc_kdn_fprop_and_dgrad(case, ordered):
    """kd-in-N convolution against torch (same bf16 operands) and against the tap-per-MMA kernel; ordered = 1 is the
    single-issuer mode (vs_set_kdn_ordered), which must also give identical bits on a second run."""
    import os
    import vae_segmentation_b200._cabi as cabi
    cabi.lib().vs_set_kdn_ordered(ordered)
    try:
        _tc_kdn_case(case, ordered)
    finally:
        cabi.lib().vs_set_kdn_ordered(int(os.environ.get("VAESEG_KDN_ORDERED", cabi.KDN_ORDERED_DEFAULT)))


def _tc_kdn_case(case, ordered):
    n, d, h, w, cin, cout = case
    torch.manual_seed(sum(case) + 7)
    x = torch.randn(n, cin, d, h, w).bfloat16().float()
    wt = torch.randn(cout, cin, 3, 3, 3) * 0.1
    wq = wt.bfloat16().float()
    y_ref = F.conv3d(x, wq, None, padding=1)
    wd = wt.to(DEV)
    dims = (n, d, h, w)
    wk = ops.pack_conv3_weight_tc_kdn(wd, dgrad=False)
    assert wk is not None
    y, stats = ops.conv3_tc_kdn(to_ndhwc(x), wk, dims, cin, cout, want_stats=True)
    torch.cuda.synchronize()
    scale = y_ref.abs().max().item()
    ys_ref = y_ref - F.conv3d(x, wt, None, padding=1)[:, :, 1:2, 1:2, 1:2]          # shifted output
    assert (from_ndhwc(y) - ys_ref).abs().max().item() < 1.5e-2 * scale
    s_ref = torch.stack([ys_ref.double().sum((2, 3, 4)), (ys_ref.double() ** 2).sum((2, 3, 4))], -1)
    assert torch.allclose(stats.cpu(), s_ref, rtol=5e-3, atol=5e-3 * s_ref.abs().max().item())
    # plain (no shift / statistics) == the production kernel on the same operands
    y2, _ = ops.conv3_tc_kdn(to_ndhwc(x), wk, dims, cin, cout, want_stats=False)
    if ordered:
        y2b, _ = ops.conv3_tc_kdn(to_ndhwc(x), wk, dims, cin, cout, want_stats=False)
        assert torch.equal(y2, y2b)
    wf, _ = ops.pack_conv3_weight(wd)
    y3, _ = ops.conv3_fprop(to_ndhwc(x), wf, None, dims, cin, cout, torch.bfloat16, shifted=False, want_stats=False,
                            wtc=ops.pack_conv3_weight_tc(wd, dgrad=False))
    assert (y2.float() - y3.float()).abs().max().item() < 1e-2 * scale
    # dgrad through the same kernel (GEMM input = Cout of the layer must be 8 or 16k, output = Cin in {8,16})
    wkd = ops.pack_conv3_weight_tc_kdn(wd, dgrad=True)
    if wkd is not None:
        gy = torch.randn_like(y_ref).bfloat16().float()
        dx_ref = F.conv_transpose3d(gy, wq, None, padding=1)
        dx, _ = ops.conv3_tc_kdn(to_ndhwc(gy), wkd, dims, cout, cin, want_stats=False)
        torch.cuda.synchronize()
        assert (from_ndhwc(dx) - dx_ref).abs().max().item() < 1e-2 * dx_ref.abs().max().item()
        # dgrad with the previous layer's InstanceNorm-backward sums fused into the epilogue == the standalone reduction
        import vae_segmentation_b200._cabi as cabi
        yprev = (torch.randn(n, cin, d, h, w) * 1.5 + 0.3).bfloat16().float()
        pst = torch.stack([yprev.double().sum((2, 3, 4)), (yprev.double() ** 2).sum((2, 3, 4))], -1).to(DEV)
        ypd = to_ndhwc(yprev)
        sums = torch.zeros(n, cin, 2, device=DEV, dtype=torch.float64)
        dx2, _ = ops.conv3_tc_kdn(to_ndhwc(gy), wkd, dims, cout, cin, prev=(ypd, pst, sums))
        sums_ref = torch.empty(n, cin, 2, device=DEV, dtype=torch.float64)
        cabi.call("vs_inorm_relu_bwd_reduce", cabi.VS_BF16, dx2.data_ptr(), ypd.data_ptr(), pst.data_ptr(), sums_ref.data_ptr(),
                  n, d * h * w, cin, 0, torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        assert (dx2.float() - dx.float()).abs().max().item() < 1e-3 * dx_ref.abs().max().item()     # issue-order rounding only
        sc = sums_ref.abs().max().item() + 1e-6
        assert (sums - sums_ref).abs().max().item() < 2e-5 * sc + 1e-4, (sums - sums_ref).abs().max().item()


K2_TC_CASES = [(2, 3, 4, 5, 8), (1, 2, 17, 9, 16), (1, 1, 2, 2, 64), (1, 3, 3, 3, 32), (2, 2, 6, 6, 128), (1, 2, 3, 3, 256),
               (1, 5, 20, 11, 8), (1, 4, 33, 8, 16)]


@pytest.mark.parametrize("variant", [1, 0])
@pytest.mark.parametrize("case", K2_TC_CASES)
def test_k2s2_tensor_core_gather_and_scatter(case, variant):
    import ctypes
    from vae_segmentation_b200 import _cabi
    try:
        _k2s2_tc_case(case, variant)
    finally:
        ctypes.CDLL(_cabi.LIB_PATH).vs_debug_set_k2_tc(1)
        ctypes.CDLL(_cabi.LIB_PATH).vs_debug_set_k2_wgrad_tc(1)


def _k2s2_tc_case(case, variant):
    """Conv3d(C,C,2,stride 2) / ConvTranspose3d(C,C,2,stride 2) forward and input gradient through the tcgen05 kernels
    (csrc/k2s2_tc.cu) against torch fp32 on the SAME bf16-rounded operands: only accumulation order and the bf16
    rounding of the output differ.  Ragged tiles (H, W not multiples of 16 / 8) exercise the TMA zero fill."""
    n, dc, hc, wc, c = case
    # variant 1 = swizzled gather operand rows (default), 0 = 8-channel no-swizzle planes
    import ctypes
    from vae_segmentation_b200 import _cabi
    ctypes.CDLL(_cabi.LIB_PATH).vs_debug_set_k2_tc(variant)
    torch.manual_seed(sum(case) + 3)
    wt = (torch.randn(c, c, 2, 2, 2) * (0.5 / c ** 0.5))
    wq = wt.bfloat16().float()
    b = torch.randn(c)
    wd = wt.to(DEV)
    pg = ops.pack_k2s2_weight_tc(wd, c, c, scatter=False)
    ps = ops.pack_k2s2_weight_tc(wd, c, c, scatter=True)
    assert pg is not None and ps is not None
    dims = (n, dc, hc, wc)

    def close(got, ref, what):
        err = (got - ref).abs().max().item()
        assert err < 1e-2 * ref.abs().max().item(), "%s: max abs err %.3e (scale %.3e)" % (what, err, ref.abs().max().item())

    fine = torch.randn(n, c, 2 * dc, 2 * hc, 2 * wc).bfloat16().float()
    coarse = torch.randn(n, c, dc, hc, wc).bfloat16().float()
    # Conv3d fprop = gather (+bias); its dgrad = scatter (no bias)
    y = ops.k2s2_gather(to_ndhwc(fine), wd, b.to(DEV), dims, c, c, wtc=pg)
    close(from_ndhwc(y), F.conv3d(fine, wq, b, stride=2), "conv k2s2 fprop")
    dx = ops.k2s2_scatter(to_ndhwc(coarse), wd, None, dims, c, c, wtc=ps)
    close(from_ndhwc(dx), F.conv_transpose3d(coarse, wq, None, stride=2), "conv k2s2 dgrad")
    # ConvTranspose3d fprop = scatter (+bias); its dgrad = gather (no bias).  Weight layout [Cin = A][Cout = B][8].
    y2 = ops.k2s2_scatter(to_ndhwc(coarse), wd, b.to(DEV), dims, c, c, wtc=ps)
    close(from_ndhwc(y2), F.conv_transpose3d(coarse, wq, b, stride=2), "convT fprop")
    dx2 = ops.k2s2_gather(to_ndhwc(fine), wd, None, dims, c, c, wtc=pg)
    close(from_ndhwc(dx2), F.conv3d(fine, wq, None, stride=2), "convT dgrad")
    # weight gradient on the tensor cores (csrc/k2s2_wgrad_tc.cu): dwt[a][b][k] = sum_o coarse[o,a] fine[2o+k,b], i.e.
    # the Conv3d weight gradient for dy = coarse, x = fine; accumulate = True adds onto the existing buffer
    ctypes.CDLL(_cabi.LIB_PATH).vs_debug_set_k2_wgrad_tc(2)          # force the tensor-core kernel at these small shapes
    fr = fine.clone().requires_grad_(False)
    wr = torch.zeros(c, c, 2, 2, 2, requires_grad=True)
    F.conv3d(fr, wr, None, stride=2).backward(coarse)
    dw = ops.k2s2_wgrad(to_ndhwc(coarse), to_ndhwc(fine), dims, c, c)
    scale = wr.grad.abs().max().item()
    assert (dw.cpu() - wr.grad).abs().max().item() < 2e-3 * scale + 1e-4, "k2s2 wgrad: %.3e (scale %.3e)" % ((dw.cpu() - wr.grad).abs().max().item(), scale)
    dw2 = ops.k2s2_wgrad(to_ndhwc(coarse), to_ndhwc(fine), dims, c, c, dwt=dw.clone(), accumulate=True)
    assert (dw2.cpu() - 2 * wr.grad).abs().max().item() < 4e-3 * scale + 2e-4
    # and against the CUDA-core kernels on the same operands
    y3 = ops.k2s2_gather(to_ndhwc(fine), wq.to(DEV), b.to(DEV), dims, c, c)
    assert (y.float() - y3.float()).abs().max().item() < 1e-2 * y3.float().abs().max().item()
    torch.cuda.synchronize()
