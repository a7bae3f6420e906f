"""Per-kernel known-answer tests: every C-ABI entry point against torch.nn.functional on CPU
fp32 (the same arithmetic the reference reaches through torch.nn).  fp32 check mode is held
to 1e-4; bf16 storage to 2e-2 (outputs) -- the tolerances BASELINE.json's north_star states."""
import pytest
import torch
import torch.nn.functional as F

from oracle import ref_torch as R
from vae_segmentation_b200 import _cabi, ops

pytestmark = pytest.mark.gpu
DEV = "cuda"
DTYPES = [torch.float32, torch.bfloat16]


def tol(dtype):
    return (1e-4, 1e-5) if dtype == torch.float32 else (2e-2, 2e-2)


def to_ndhwc(x, dtype):
    return x.permute(0, 2, 3, 4, 1).contiguous().to(DEV, dtype)


def from_ndhwc(x):
    return x.float().cpu().permute(0, 4, 1, 2, 3).contiguous()


def check(got, want, dtype, scale=1.0, what=""):
    rtol, atol = tol(dtype)
    got, want = got.float().cpu(), want.float().cpu()
    err = (got - want).abs().max().item()
    ref = want.abs().max().item()
    assert torch.allclose(got, want, rtol=rtol, atol=atol * max(scale, ref)), \
        "%s: max abs err %.3e (ref max %.3e)" % (what, err, ref)


def q(x, dtype):
    """Round test inputs to the storage dtype so both sides see identical operands."""
    return x.to(dtype).float()


CONV_CASES = [  # (n, d, h, w, cin, cout, planar_in)
    (2, 5, 9, 10, 1, 8, True), (1, 4, 8, 8, 2, 8, True), (2, 6, 7, 9, 8, 16, False), (1, 9, 8, 17, 16, 8, False),
    (1, 4, 4, 4, 32, 32, False), (1, 3, 3, 3, 64, 16, False), (2, 4, 5, 6, 8, 2, False),
    (2, 9, 19, 70, 1, 8, True), (2, 6, 17, 33, 2, 8, True), (1, 3, 3, 3, 2, 8, True)]


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("case", CONV_CASES)
def test_conv3_fprop_dgrad_wgrad(case, dtype):
    n, d, h, w, cin, cout, planar = case
    torch.manual_seed(sum(case))
    in_dtype = torch.float32 if planar else dtype
    x = q(torch.randn(n, cin, d, h, w), in_dtype)
    wt = torch.randn(cout, cin, 3, 3, 3) * 0.2
    b = torch.randn(cout)
    xr, wr, br = x.clone().requires_grad_(), wt.clone().requires_grad_(), b.clone().requires_grad_()
    y_ref = F.conv3d(xr, wr, br, padding=1)
    wf, wd = ops.pack_conv3_weight(wt.to(DEV))
    assert torch.equal(wf.cpu(), wt.reshape(cout, cin, 27).permute(2, 1, 0))
    xin = x.to(DEV) if planar else to_ndhwc(x, dtype)
    y, stats = ops.conv3_fprop(xin, wf, b.to(DEV), (n, d, h, w), cin, cout, dtype, in_planar=planar, shifted=False)
    check(from_ndhwc(y), y_ref, dtype, what="fprop")
    s_ref = torch.stack([y_ref.double().sum((2, 3, 4)), (y_ref.double() ** 2).sum((2, 3, 4))], -1).detach()
    assert stats.dtype == torch.float64
    check(stats, s_ref, torch.float32 if dtype == torch.float32 else dtype, what="stats")
    # shifted variant: every (n, co) channel minus its value at voxel (1,1,1)
    ys, stats_s = ops.conv3_fprop(xin, wf, None, (n, d, h, w), cin, cout, dtype, in_planar=planar, shifted=True)
    ys_ref = (y_ref - y_ref[:, :, 1:2, 1:2, 1:2]).detach()
    check(from_ndhwc(ys), ys_ref, dtype, scale=y_ref.abs().max().item(), what="shifted fprop")
    s_ref = torch.stack([ys_ref.double().sum((2, 3, 4)), (ys_ref.double() ** 2).sum((2, 3, 4))], -1)
    check(stats_s, s_ref, torch.float32 if dtype == torch.float32 else dtype, scale=s_ref.abs().max().item(), what="shifted stats")
    # backward
    gy = q(torch.randn_like(y_ref), dtype)
    y_ref.backward(gy)
    gyd = to_ndhwc(gy, dtype)
    dx = ops.conv3_dgrad(gyd, wd, (n, d, h, w), cin, cout, dtype, out_planar=planar)
    check(dx if planar else from_ndhwc(dx), xr.grad, dtype, what="dgrad")
    db = torch.zeros(cout, device=DEV)
    dw, _ = ops.conv3_wgrad(xin, gyd, (n, d, h, w), cin, cout, db=db, in_planar=planar)
    check(dw, wr.grad, dtype, what="wgrad")
    check(db, br.grad, dtype, what="bias grad")
    dw2, _ = ops.conv3_wgrad(xin, gyd, (n, d, h, w), cin, cout, dw=dw.clone(), in_planar=planar, accumulate=True)
    check(dw2, 2 * wr.grad, dtype, what="wgrad accumulate")


@pytest.mark.parametrize("dtype", DTYPES)
def test_conv3_head_fp32_logits(dtype):
    """out_block: NDHWC activations in, fp32 logits out (kept fp32 for argmax parity)."""
    torch.manual_seed(5)
    n, d, h, w, cin, cout = 1, 5, 6, 7, 8, 2
    x = q(torch.randn(n, cin, d, h, w), dtype)
    wt, b = torch.randn(cout, cin, 3, 3, 3) * 0.2, torch.randn(cout)
    wf, _ = ops.pack_conv3_weight(wt.to(DEV))
    y, _ = ops.conv3_fprop(to_ndhwc(x, dtype), wf, b.to(DEV), (n, d, h, w), cin, cout, torch.float32, want_stats=False)
    assert y.dtype == torch.float32
    check(from_ndhwc(y), F.conv3d(x, wt, b, padding=1), torch.float32, what="head logits")


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("shape", [(2, 3, 4, 5, 8), (1, 2, 2, 3, 16), (1, 1, 2, 2, 64), (1, 3, 3, 3, 32)])
def test_k2s2_conv_and_transpose(shape, dtype):
    n, dc, hc, wc, c = shape
    torch.manual_seed(sum(shape))
    # Conv3d(C,C,2,stride 2): fprop = gather, dgrad = scatter
    x = q(torch.randn(n, c, 2 * dc, 2 * hc, 2 * wc), dtype)
    wt, b = torch.randn(c, c, 2, 2, 2) * 0.3, torch.randn(c)
    xr, wr, br = x.clone().requires_grad_(), wt.clone().requires_grad_(), b.clone().requires_grad_()
    y_ref = F.conv3d(xr, wr, br, stride=2)
    xd = to_ndhwc(x, dtype)
    y = ops.k2s2_gather(xd, wt.to(DEV), b.to(DEV), (n, dc, hc, wc), c, c)
    check(from_ndhwc(y), y_ref, dtype, what="k2s2 conv fprop")
    gy = q(torch.randn_like(y_ref), dtype)
    y_ref.backward(gy)
    gyd = to_ndhwc(gy, dtype)
    dx = ops.k2s2_scatter(gyd, wt.to(DEV), None, (n, dc, hc, wc), c, c)
    check(from_ndhwc(dx), xr.grad, dtype, what="k2s2 conv dgrad")
    db = torch.empty(c, device=DEV)
    dw = ops.k2s2_wgrad(gyd, xd, (n, dc, hc, wc), c, c, dbias_coarse=db)
    check(dw, wr.grad, dtype, what="k2s2 conv wgrad")
    check(db, br.grad, dtype, what="k2s2 conv bias grad")
    # ConvTranspose3d(C,C,2,stride 2): fprop = scatter, dgrad = gather
    x = q(torch.randn(n, c, dc, hc, wc), dtype)
    xr, wr, br = x.clone().requires_grad_(), wt.clone().requires_grad_(), b.clone().requires_grad_()
    y_ref = F.conv_transpose3d(xr, wr, br, stride=2)
    xd = to_ndhwc(x, dtype)
    y = ops.k2s2_scatter(xd, wt.to(DEV), b.to(DEV), (n, dc, hc, wc), c, c)
    check(from_ndhwc(y), y_ref, dtype, what="convT fprop")
    gy = q(torch.randn_like(y_ref), dtype)
    y_ref.backward(gy)
    gyd = to_ndhwc(gy, dtype)
    dx = ops.k2s2_gather(gyd, wt.to(DEV), None, (n, dc, hc, wc), c, c)
    check(from_ndhwc(dx), xr.grad, dtype, what="convT dgrad")
    db = torch.empty(c, device=DEV)
    dw = ops.k2s2_wgrad(xd, gyd, (n, dc, hc, wc), c, c, dbias_fine=db)
    check(dw, wr.grad, dtype, what="convT wgrad")
    check(db, br.grad, dtype, what="convT bias grad")


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("shape", [(2, 8, 5, 6, 7), (1, 16, 4, 4, 4), (1, 256, 2, 2, 2), (2, 32, 3, 5, 9)])
def test_instance_norm_relu_fwd_bwd(shape, dtype):
    n, c, d, h, w = shape
    torch.manual_seed(sum(shape))
    y = q(torch.randn(n, c, d, h, w) * 2 + 0.5, dtype)
    skip = q(torch.randn(n, c, d, h, w), dtype)
    yr = y.clone().requires_grad_()
    a_ref = F.relu(F.instance_norm(yr, eps=1e-5)) + skip
    stats = torch.stack([y.double().sum((2, 3, 4)), (y.double() ** 2).sum((2, 3, 4))], -1).to(DEV)
    yd = to_ndhwc(y, dtype)
    a = ops.inorm_relu_apply(yd, stats, to_ndhwc(skip, dtype))
    check(from_ndhwc(a), a_ref, dtype, what="inorm+relu+skip")
    a2 = ops.inorm_relu_apply(yd, stats, None)
    check(from_ndhwc(a2), a_ref - skip, dtype, what="inorm+relu")
    g = q(torch.randn_like(y), dtype)
    a_ref.backward(g)
    dy = ops.inorm_relu_bwd(to_ndhwc(g, dtype), yd, stats)
    check(from_ndhwc(dy), yr.grad, dtype, scale=yr.grad.abs().max().item(), what="inorm+relu backward")
    acc = to_ndhwc(g, dtype)
    ops.add_inplace(acc, to_ndhwc(skip, dtype))
    check(from_ndhwc(acc), g + skip, dtype, what="add_inplace")


@pytest.mark.parametrize("dtype", DTYPES)
def test_softmax2(dtype):
    torch.manual_seed(0)
    n, d, h, w = 2, 3, 5, 7
    logits = torch.randn(n, 2, d, h, w) * 3
    lr = logits.clone().requires_grad_()
    p_ref = F.softmax(lr, dim=1)
    probs = ops.softmax2_fwd(to_ndhwc(logits, torch.float32), (n, d, h, w))
    check(probs, p_ref, torch.float32, what="softmax")
    assert torch.equal(probs.argmax(1).cpu(), p_ref.argmax(1))
    g = torch.randn_like(logits)
    p_ref.backward(g)
    dl = ops.softmax2_bwd(g.to(DEV), probs, (n, d, h, w), dtype)
    check(from_ndhwc(dl), lr.grad, dtype, what="softmax backward")


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("use_z", [False, True])
def test_fc_encode_decode(dtype, use_z):
    torch.manual_seed(11)
    batch, side, c, dim = 3, 2, 16, 24
    s3, flat = side ** 3, 16 * side ** 3
    x = q(torch.randn(batch, c, side, side, side), dtype)
    wm, bm = torch.randn(dim, flat) * 0.1, torch.randn(dim) * 0.1
    ws, bs = torch.randn(dim, flat) * 0.1, torch.randn(dim) * 0.1
    w2, b2 = torch.randn(flat, dim) * 0.1, torch.randn(flat) * 0.1
    z = torch.randn(batch, dim)
    scale = 0.35
    leaves = [t.clone().requires_grad_() for t in (x, wm, bm, ws, bs, w2, b2)]
    xr, wmr, bmr, wsr, bsr, w2r, b2r = leaves
    xf = xr.reshape(batch, -1)
    mean_ref = F.linear(xf, wmr, bmr)
    std_ref = F.relu(F.linear(xf, wsr, bsr))
    lat_ref = mean_ref + z * std_ref * scale if use_z else mean_ref
    h_ref = F.linear(lat_ref, w2r, b2r).view(batch, c, side, side, side)
    d = lambda t: t.to(DEV)
    xd = to_ndhwc(x, dtype)
    mean, std, lat = ops.fc_encode_fwd(xd, d(wm), d(bm), d(ws), d(bs), d(z), scale, use_z, batch, s3, c, dim)
    check(mean, mean_ref, torch.float32, what="mean")
    check(std, std_ref, torch.float32, what="std")
    check(lat, lat_ref, torch.float32, what="lat")
    h = ops.fc_decode_fwd(lat, d(w2), d(b2), batch, s3, c, dim, dtype, side)
    check(from_ndhwc(h), h_ref, dtype, what="fc2")
    gh = q(torch.randn_like(h_ref), dtype)
    gm_ext, gs_ext = torch.randn(batch, dim), torch.randn(batch, dim)
    (h_ref * gh).sum().add((mean_ref * gm_ext).sum()).add((std_ref * gs_ext).sum()).backward()
    dw2, db2 = torch.empty_like(d(w2)), torch.empty_like(d(b2))
    dlat = ops.fc_decode_bwd(to_ndhwc(gh, dtype), lat, d(w2), batch, s3, c, dim, dw2=dw2, db2=db2)
    check(dw2, w2r.grad, dtype, what="dw2")
    check(db2, b2r.grad, dtype, what="db2")
    dwm, dbm, dws, dbs = [torch.empty_like(d(t)) for t in (wm, bm, ws, bs)]
    dx = ops.fc_encode_bwd(xd, d(wm), d(ws), d(z), scale, use_z, std, dlat, d(gm_ext), d(gs_ext), batch, s3, c, dim,
                           want_dx=True, dwm=dwm, dbm=dbm, dws=dws, dbs=dbs)
    check(from_ndhwc(dx), xr.grad, dtype, what="fc dx")
    check(dwm, wmr.grad, dtype, what="dwm")
    check(dbm, bmr.grad, dtype, what="dbm")
    check(dws, wsr.grad, dtype, what="dws")
    check(dbs, bsr.grad, dtype, what="dbs")


def _tgt_ref(t, mode):
    if mode == _cabi.TGT_BINARIZE:
        return (t >= 0.5).float()
    if mode == _cabi.TGT_CONFIDENT:
        b = t.clone()
        b[b > 0.8] = 1
        b[b < 0.2] = 0
        return b
    return t


@pytest.mark.parametrize("shape", [(2, 2, 4, 6, 8), (3, 2, 5, 5, 5), (1, 1, 3, 3, 7)])
def test_dice_sums_and_backward(shape):
    torch.manual_seed(sum(shape))
    n, c = shape[0], shape[1]
    src = torch.rand(*shape)
    tgt = torch.rand(*shape)
    for mode in (_cabi.TGT_TENSOR, _cabi.TGT_BINARIZE, _cabi.TGT_CONFIDENT):
        sr, tr = src.clone().requires_grad_(), tgt.clone().requires_grad_()
        tt = _tgt_ref(tr, mode)
        i_ref, s_ref, t_ref = (sr * tt).sum((2, 3, 4)), sr.sum((2, 3, 4)), tt.sum((2, 3, 4))
        sums = ops.dice_sums(src.to(DEV), tgt.to(DEV), mode)
        check(sums, torch.stack([i_ref, s_ref, t_ref], -1), torch.float32, what="dice sums mode %d" % mode)
        per = 2 * i_ref / (s_ref + t_ref + 1e-6)
        gper = torch.randn(n, c)
        per.backward(gper)
        gs, gt = ops.dice_bwd(src.to(DEV), tgt.to(DEV), mode, sums, gper.to(DEV), 1e-6,
                              want_src=True, want_tgt=(mode == _cabi.TGT_TENSOR))
        check(gs, sr.grad, torch.float32, scale=sr.grad.abs().max().item(), what="dice grad src mode %d" % mode)
        if mode == _cabi.TGT_TENSOR:
            check(gt, tr.grad, torch.float32, scale=tr.grad.abs().max().item(), what="dice grad tgt")
    if c == 2:
        label = (torch.rand(n, 1, *shape[2:]) > 0.7).float()
        oh = torch.zeros(*shape).scatter_(1, label.long(), 1)
        sums = ops.dice_sums(src.to(DEV), label.to(DEV), _cabi.TGT_LABEL)
        want = torch.stack([(src * oh).sum((2, 3, 4)), src.sum((2, 3, 4)), oh.sum((2, 3, 4))], -1)
        check(sums, want, torch.float32, what="dice sums label")
        assert torch.equal(ops.one_hot(label.to(DEV), 2).cpu(), oh)
        so = torch.zeros(*shape).scatter_(1, src.argmax(1, keepdim=True), 1)
        to = torch.zeros(*shape).scatter_(1, tgt.argmax(1, keepdim=True), 1)
        sums = ops.dice_sums(src.to(DEV), tgt.to(DEV), _cabi.TGT_ARGMAX)
        want = torch.stack([(so * to).sum((2, 3, 4)), so.sum((2, 3, 4)), to.sum((2, 3, 4))], -1)
        check(sums, want, torch.float32, what="dice sums argmax")


def test_kl_binarize_compose():
    torch.manual_seed(2)
    mean, std = torch.randn(3, 128), torch.rand(3, 128)
    std[0, :5] = 0.0                       # relu'd std can be exactly zero
    mr, sr = mean.clone().requires_grad_(), std.clone().requires_grad_()
    ref = torch.mean(0.5 * (torch.sum(sr ** 2, 1) + torch.sum(mr ** 2, 1) - 2 * torch.sum(torch.log(sr + 0.00001), 1)))
    out = ops.kl_fwd(mean.to(DEV), std.to(DEV))
    check(out, ref.reshape(1), torch.float32, what="kl")
    ref.backward()
    gm, gs = ops.kl_bwd(mean.to(DEV), std.to(DEV), torch.ones(1, device=DEV))
    check(gm, mr.grad, torch.float32, what="kl gmean")
    check(gs, sr.grad, torch.float32, scale=1.0, what="kl gstd")
    a = torch.rand(2, 2, 3, 4, 5)
    a[0, 0, 0, 0, :3] = torch.tensor([0.5, 0.8, 0.2])
    assert torch.equal(ops.binarize(a.to(DEV), _cabi.TGT_BINARIZE).cpu(), (a >= 0.5).float())
    assert torch.equal(ops.binarize(a.to(DEV), _cabi.TGT_CONFIDENT).cpu(), _tgt_ref(a, _cabi.TGT_CONFIDENT))
    from oracle import ref_torch as R
    for recon, lam, lt, kl in ((0.10, 1.0, 8, False), (0.20, 1.0, 8, True), (0.25, 1.0, 8, False), (0.40, 0.2, 8, True),
                               (0.33, 2.0, 0, True), (0.33, 1.0, 0, False)):
        terms = torch.tensor([recon, 0.5, 3.0])
        final, wts = ops.compose_target_loss(terms.to(DEV), lam, lt, kl)
        want = R.compose_target_loss(terms[0], terms[1], terms[2], lambda_vae=lam, loss_type=lt, kl=kl)
        assert abs(final.item() - float(want)) < 1e-6
        assert abs((wts.cpu() * terms).sum().item() - float(want)) < 1e-6


def test_sgd_adam_ema():
    torch.manual_seed(3)
    n = 1000
    p0 = torch.randn(n)
    pr = torch.nn.Parameter(p0.clone())
    opt = torch.optim.SGD([pr], lr=1e-2, momentum=0.9)
    p, buf = p0.to(DEV), torch.zeros(n, device=DEV)
    for step in range(3):
        g = torch.randn(n)
        pr.grad = g.clone()
        opt.step()
        ops.sgd_step(p, g.to(DEV), buf, 1e-2, 0.9, first=(step == 0))
    check(p, pr.detach(), torch.float32, what="sgd momentum")
    pr = torch.nn.Parameter(p0.clone())
    opt = torch.optim.SGD([pr], lr=1e-2, momentum=0.0)
    p = p0.to(DEV)
    g = torch.randn(n)
    pr.grad = g.clone()
    opt.step()
    ops.sgd_step(p, g.to(DEV), None, 1e-2, 0.0, first=True)
    check(p, pr.detach(), torch.float32, what="sgd plain")
    pr = torch.nn.Parameter(p0.clone())
    opt = torch.optim.Adam([pr], lr=1e-3)
    p, m, v = p0.to(DEV), torch.zeros(n, device=DEV), torch.zeros(n, device=DEV)
    for step in range(3):
        g = torch.randn(n)
        pr.grad = g.clone()
        opt.step()
        ops.adam_step(p, g.to(DEV), m, v, 1e-3, 0.9, 0.999, 1e-8, step + 1)
    check(p, pr.detach(), torch.float32, what="adam")
    t, s = torch.randn(n), torch.randn(n)
    td = t.to(DEV)
    ops.ema_update(td, s.to(DEV), 0.995)
    check(td, 0.995 * t + 0.005 * s, torch.float32, what="ema")


@pytest.mark.parametrize("cfg", [(1.0, 0, False, False, False), (0.7, 8, True, False, True), (1.0, 0, False, True, False),
                                 (2.0, 8, False, False, False)])
def test_joint_target_loss_fused_vs_composed(cfg):
    """ev.joint_target_loss (one autograd node, one scalar kernel) against the same loss composed from the
    reference-named pieces (avg_dsc / KLloss + main_target.py:550-560,588-590 arithmetic in torch)."""
    from vae_segmentation_b200 import evaluation as ev
    lam, loss_type, use_kl, only_pseudo, confident = cfg
    torch.manual_seed(5)
    n, d = 2, 12
    mk = lambda: torch.softmax(torch.randn(n, 2, d, d, d, device=DEV) * 2, 1)
    pred0, recon0, tpred = mk(), mk(), mk()
    label = (torch.rand(n, 1, d, d, d, device=DEV) > 0.7).float()
    kl = torch.tensor(3.25, device=DEV)
    outs = []
    for fused in (True, False):
        pred, recon = pred0.clone().requires_grad_(), recon0.clone().requires_grad_()
        if fused:
            final, mon = ev.joint_target_loss(pred, recon, label, tpred, kl=kl, lambda_vae=lam, loss_type=loss_type,
                                              use_kl=use_kl, only_pseudo=only_pseudo, confident=confident)
            mon = mon.cpu()
        else:
            r = 1 - ev.avg_dsc_fused(pred, recon, "tensor", botindex=1, topindex=2)
            g = 1 - ev.avg_dsc_fused(pred.detach(), label, "label", botindex=1, topindex=2)
            f = 1 - ev.avg_dsc_fused(pred, tpred, "confident" if confident else "binarize", botindex=1, topindex=2)
            if only_pseudo:
                final = f
            elif loss_type == 8:
                rv = r.item()
                cur = lam * (0.6 if rv < 0.15 else 1.2 if rv < 0.225 else 2.0 if rv < 0.3 else 3.0)
                final = (r + f / cur + (kl if use_kl else 0)) if cur > 1 else (cur * r + f + (cur * kl if use_kl else 0))
            else:
                final = lam * r + f + (0.00002 * lam * kl if use_kl else 0)
            mon = torch.stack([final.detach(), r.detach(), g.detach(), f.detach(), kl]).cpu()
        final.backward()
        outs.append((final.detach().cpu(), mon, pred.grad.cpu(), recon.grad.cpu() if recon.grad is not None else None))
    (fa, ma, ga, ra), (fb, mb, gb, rb) = outs
    assert torch.allclose(fa, fb, rtol=1e-5, atol=1e-6)
    assert torch.allclose(ma, mb, rtol=1e-5, atol=1e-6)
    assert torch.allclose(ga, gb, rtol=1e-4, atol=1e-9)
    if rb is None:
        assert ra is None or ra.abs().max().item() == 0.0
    else:
        assert torch.allclose(ra, rb, rtol=1e-4, atol=1e-9)


@pytest.mark.parametrize("kind", ["float32", "int16"])
def test_clip_center_bit_exact(kind):
    """Fused Clip + CenterIntensities (utils/utils.py:508-533,572-618 as used at main_target.py:223-224) against the
    numpy restatement: IEEE fp32 clip / subtract / divide, so the comparison is bit-exact."""
    from vae_segmentation_b200 import transforms
    torch.manual_seed(8)
    if kind == "int16":
        x = torch.randint(-1024, 3000, (2, 1, 9, 10, 11), dtype=torch.int16)
    else:
        x = (torch.randn(2, 1, 9, 10, 11) * 400 + 50)
    want = torch.from_numpy(R.clip_center(x.numpy()))
    out = transforms.ClipCenter(["venous"])({"venous": x.to(DEV), "other": None})["venous"]
    assert out.dtype == torch.float32 and out.shape == x.shape
    assert torch.equal(out.cpu(), want)
    assert want.min().item() >= -1.0 and want.max().item() <= 1.0


@pytest.mark.parametrize("case", [(40, 44, 36, (14, 12, 10), (20, 20, 18), 32, 0), (64, 64, 64, (30, 26, 28), (32, 30, 34), 32, 0),
                                  (48, 40, 56, (10, 8, 12), (3, 36, 50), 24, 2), (36, 36, 36, None, None, 16, 0)])
def test_crop_resize_matches_skimage_restatement(case):
    """Device CropResize (csrc/resize.cu) against the restatement of utils/utils.py:220-293 + skimage 0.18.3 resize over
    scipy.ndimage (oracle/ref_resize.py): image within 1e-5 of its range (float32 storage of double arithmetic on both
    sides), label bit-exact.  Cases: downsampling (anti-aliasing filter active), upsampling, a blob touching the volume
    border (clamped crop + the reference's centred re-padding) with a shift, and an empty label (the reference's fallback
    cube)."""
    import numpy as np
    from oracle import ref_resize as RR
    from vae_segmentation_b200.transforms import CropResize
    d, h, w, radii, centre, out, shift = case
    rng = np.random.RandomState(d + h + w)
    img = (rng.randn(d, h, w) * 300 + 50).astype(np.float32)
    zz, yy, xx = np.meshgrid(np.arange(d), np.arange(h), np.arange(w), indexing="ij")
    if radii is None:
        label = np.zeros((d, h, w), np.float32)
        label_big = np.zeros((128, 128, 128), np.float32)          # the fallback cube is centred at (64,64,64): needs a big volume
        img_big = (rng.randn(128, 128, 128) * 300).astype(np.float32)
        img, label = img_big, label_big
    else:
        label = ((((zz - centre[0]) / radii[0]) ** 2 + ((yy - centre[1]) / radii[1]) ** 2 + ((xx - centre[2]) / radii[2]) ** 2) <= 1).astype(np.float32)
    want_img, want_lab, want_shape = RR.crop_resize(img, label, [out] * 3, shift=shift)
    dd = {"venous": torch.from_numpy(img).to(DEV), "venous_pancreas": torch.from_numpy(label).to(DEV)}
    dd = CropResize(["venous"], [out] * 3, shift=shift)(dd)
    torch.cuda.synchronize()
    got_img, got_lab = dd["venous"].cpu().numpy(), dd["venous_pancreas"].cpu().numpy()
    assert got_img.shape == want_img.shape == (out, out, out)
    assert np.array_equal(got_lab, want_lab), "label patch differs in %d voxels" % int((got_lab != want_lab).sum())
    scale = float(np.abs(want_img).max())
    assert np.abs(got_img - want_img).max() < 1e-5 * scale, np.abs(got_img - want_img).max()
    assert dd["ori_shape"].tolist() == want_shape.tolist()
