"""Tensor-level wrappers over the C ABI (one function per entry point of include/vaeseg_b200.h).

PyTorch supplies device memory and the current CUDA stream; every op launches hand-written
kernels from libvaeseg_b200.so.  CPU tensors are rejected: there is no fallback path.
"""
import torch

from . import _cabi
from ._cabi import VS_BF16, VS_F32, VS_FLAG_PREZEROED

_DT = {torch.float32: VS_F32, torch.bfloat16: VS_BF16}


def _dt(t):
    try:
        return _DT[t.dtype]
    except KeyError:
        raise RuntimeError("vaeseg_b200: unsupported activation dtype %s" % t.dtype)


def _p(t):
    if t is None:
        return None
    if not t.is_cuda:
        raise RuntimeError("vaeseg_b200: tensor is on %s; the hot path runs on CUDA only (no CPU fallback)" % t.device)
    if not t.is_contiguous():
        raise RuntimeError("vaeseg_b200: tensor must be contiguous")
    return t.data_ptr()


def _f32(t, what):
    if t is not None and t.dtype != torch.float32:
        raise RuntimeError("vaeseg_b200: %s must be float32, got %s" % (what, t.dtype))
    return t


def _stream():
    return torch.cuda.current_stream().cuda_stream


def has_tcgen05():
    return bool(_cabi.lib().vs_has_tcgen05())


# ---- 3x3x3 convolution ---------------------------------------------------------------
def pack_conv3_weight(w, want_dgrad=True, out=None):
    """[Cout,Cin,3,3,3] fp32 -> (wf [27,Cin,Cout], wd [27,Cout,Cin] or None); `out=(wf, wd)` re-packs in place."""
    cout, cin = w.shape[0], w.shape[1]
    _f32(w, "weight")
    if out is not None:
        wf, wd = out
    else:
        wf = torch.empty(27, cin, cout, device=w.device, dtype=torch.float32)
        wd = torch.empty(27, cout, cin, device=w.device, dtype=torch.float32) if want_dgrad else None
    _cabi.call("vs_pack_conv3_weight", _p(w), _p(wf), _p(wd), cin, cout, _stream())
    return wf, wd


def pack_conv3_weight_tc(w, dgrad=False, out=None):
    """bf16 tensor-core (UMMA B operand) pack of a [Cout,Cin,3,3,3] weight, or None if the tcgen05
    path does not take this shape / the library was built without it.  `out` re-packs in place."""
    cout, cin = w.shape[0], w.shape[1]
    nbytes = _cabi.lib().vs_conv3_tc_pack_bytes(cin, cout, int(dgrad))
    if nbytes == 0:
        return None
    if out is None:
        out = torch.empty(nbytes // 2, device=w.device, dtype=torch.bfloat16)
    _cabi.call("vs_pack_conv3_weight_tc", _p(_f32(w, "weight")), _p(out), cin, cout, int(dgrad), _stream())
    return out


def pack_conv3_weight_tc_padded(w, cin_pad, cout_pad, dgrad=False, out=None):
    """bf16 tensor-core pack of a [Cout,Cin,3,3,3] weight zero-padded to cin_pad / cout_pad channels."""
    cout, cin = w.shape[0], w.shape[1]
    nbytes = _cabi.lib().vs_conv3_tc_pack_bytes(cin_pad, cout_pad, int(dgrad))
    if nbytes == 0:
        return None
    if out is None:
        out = torch.empty(nbytes // 2, device=w.device, dtype=torch.bfloat16)
    _cabi.call("vs_pack_conv3_weight_tc_padded", _p(_f32(w, "weight")), _p(out), cin, cout, cin_pad, cout_pad, int(dgrad), _stream())
    return out


def head_conv_softmax2(x, wtc8, bias, dims, cin):
    """probs [N,2,D,H,W] fp32 = softmax(conv3(x, w) + bias): the 2-class head in one tensor-core launch."""
    n, d, h, w = dims
    if x.dtype != torch.bfloat16:
        raise RuntimeError("vaeseg_b200: head_conv_softmax2 takes bf16 NDHWC activations")
    probs = torch.empty(n, 2, d, h, w, device=x.device, dtype=torch.float32)
    _cabi.call("vs_head_conv_softmax2_fwd", _p(x), _p(wtc8), _p(_f32(bias, "bias")), _p(probs), n, d, h, w, cin, _stream())
    return probs


def pack_conv3_batched(jobs_dev, njobs):
    """One launch re-packing every layer described by the device job table (engine.PackCache.repack_all)."""
    _cabi.call("vs_pack_conv3_batched", _p(jobs_dev), int(njobs), _stream())


def stats_words(n, cout):
    """fp64 words of one layer's statistics + shift block (see conv3_fprop / StatsArena) + one word whose first 32 bits
    are the grid-barrier counter of the fused conv + InstanceNorm + ReLU launch (conv3_in_relu)."""
    return n * cout * 2 + (n * cout + 1) // 2 + 1


class StatsArena(object):
    """One zero-filled fp64 buffer carved into per-layer (statistics, shift) or (sums) blocks: a network pass zeroes
    it with ONE launch instead of one memset node per layer on the critical path."""

    def __init__(self, words, device):
        self.buf = torch.zeros(max(int(words), 1), device=device, dtype=torch.float64)
        self.off = 0

    def take(self, words):
        if self.off + words > self.buf.numel():
            raise RuntimeError("vaeseg_b200: StatsArena exhausted")
        out = self.buf[self.off:self.off + words]
        self.off += words
        return out


def conv3_fprop(x, wf, bias, dims, cin, cout, out_dtype, in_planar=False, out_planar=False, want_stats=True,
                shifted=None, wtc=None, arena=None, return_shift=False):
    """dims = (N, D, H, W).  Returns (y, stats).  `shifted` (default: same as want_stats) subtracts the
    per-(n,co) reference-voxel value from the output (InstanceNorm-invariant, see the header).  `arena`: a
    zero-filled StatsArena that supplies the statistics / shift words (the call then skips its memset)."""
    n, d, h, w = dims
    dev = x.device
    if out_planar:
        y = torch.empty(n, cout, d, h, w, device=dev, dtype=torch.float32)
    else:
        y = torch.empty(n, d, h, w, cout, device=dev, dtype=out_dtype)
    if shifted is None:
        shifted = want_stats
    stats = shift = None
    flags = 0
    if want_stats and shifted:
        # one allocation, statistics first and the shift words right behind them: the library zeroes both with a
        # single memset (the tensor-core kernel publishes the shift through those words, see conv3_tc.cu)
        if arena is not None:
            buf = arena.take(stats_words(n, cout))
            flags = VS_FLAG_PREZEROED
        else:
            buf = torch.empty(stats_words(n, cout), device=dev, dtype=torch.float64)
        stats = buf[:n * cout * 2].view(n, cout, 2)
        shift = buf[n * cout * 2:].view(torch.float32)[:n * cout].view(n, cout)
    elif want_stats:
        stats = torch.empty(n, cout, 2, device=dev, dtype=torch.float64)
    elif shifted:
        shift = torch.empty(n, cout, device=dev, dtype=torch.float32)
    _cabi.call("vs_conv3x3x3_fprop", _dt(x), _dt(y), int(in_planar), int(out_planar), flags, _p(x), _p(_f32(wf, "wf")), _p(wtc),
               _p(_f32(bias, "bias")), _p(y), _p(stats), _p(shift), n, d, h, w, cin, cout, _stream())
    if return_shift:
        return y, stats, shift
    return y, stats


def dgrad_can_fuse_reduce(dy, cin, cout, out_dtype, out_planar, wdtc, dims=None):
    """Whether conv3_dgrad takes the tensor-core path for this call, i.e. may fuse the previous layer's
    InstanceNorm-backward reduction into its epilogue -- and whether that pays: the 16-column kernel variant prefetches
    the previous layer's output behind the MMAs (fusion is free at every size); the 32-column variant (32 <= Cin < 64)
    loads it inline, which only beats a separate reduction launch on small volumes (measured: <= 16^3 per sample)."""
    gin, gout = cout, cin
    ok = (wdtc is not None and dy.dtype == torch.bfloat16 and out_dtype == torch.bfloat16 and not out_planar
          and (gin == 8 or (gin >= 16 and gin % 16 == 0)) and gout >= 8 and gout % 8 == 0)
    if ok and dims is not None and 32 <= gout < 64:
        ok = dims[1] * dims[2] * dims[3] <= 4096
    return ok


def conv3_dgrad(dy, wd, dims, cin, cout, out_dtype, out_planar=False, wdtc=None, prev=None):
    """dy: [N,D,H,W,Cout] -> dx [N,D,H,W,Cin] (or planar fp32 [N,Cin,D,H,W]).  prev = (y_prev, stats_prev,
    sums_prev): also accumulate the previous layer's norm-backward sums (tensor-core path only; sums pre-zeroed)."""
    n, d, h, w = dims
    yp, sp, sums = prev if prev is not None else (None, None, None)
    if out_planar:
        dx = torch.empty(n, cin, d, h, w, device=dy.device, dtype=torch.float32)
    else:
        dx = torch.empty(n, d, h, w, cin, device=dy.device, dtype=out_dtype)
    _cabi.call("vs_conv3x3x3_dgrad", _dt(dy), _dt(dx), int(out_planar), _p(dy), _p(_f32(wd, "wd")), _p(wdtc), _p(dx),
               _p(yp), _p(sp), _p(sums), n, d, h, w, cin, cout, _stream())
    return dx


def conv3_wgrad(x, dy, dims, cin, cout, dw=None, db=None, in_planar=False, accumulate=False):
    """dw [Cout,Cin,3,3,3] fp32 (+)= ; db [Cout] optional.  Returns (dw, db)."""
    n, d, h, w = dims
    if dw is None:
        dw = torch.empty(cout, cin, 3, 3, 3, device=dy.device, dtype=torch.float32)
        accumulate = False
    _cabi.call("vs_conv3x3x3_wgrad", _dt(dy), int(in_planar), _p(x), _p(dy), _p(_f32(dw, "dw")), _p(_f32(db, "db")),
               None, 0, int(accumulate), n, d, h, w, cin, cout, _stream())
    return dw, db


# ---- k2s2 ------------------------------------------------------------------------------
def pack_k2s2_weight_tc(wt, a, b, scatter, out=None):
    """bf16 tensor-core (UMMA B operand) pack of a k2s2 weight viewed as wt[A][B][8], or None when the tcgen05 path
    does not take the channel counts.  `out` re-packs in place."""
    nbytes = _cabi.lib().vs_k2s2_tc_pack_bytes(a, b, int(scatter))
    if nbytes == 0:
        return None
    if out is None:
        out = torch.empty(nbytes // 2, device=wt.device, dtype=torch.bfloat16)
    _cabi.call("vs_pack_k2s2_weight_tc", _p(_f32(wt, "wt")), _p(out), a, b, int(scatter), _stream())
    return out


def k2s2_gather(fine, wt, bias, cdims, a, b, wtc=None):
    """fine [N,2dc,2hc,2wc,B] -> coarse [N,dc,hc,wc,A]; cdims = (N,dc,hc,wc).  wtc: tensor-core gather pack (bf16
    activations only) -> tcgen05 kernel, else the CUDA-core kernel on the fp32 weight."""
    n, dc, hc, wc = cdims
    coarse = torch.empty(n, dc, hc, wc, a, device=fine.device, dtype=fine.dtype)
    if wtc is not None and fine.dtype == torch.bfloat16:
        _cabi.call("vs_k2s2_gather_tc", _p(fine), _p(wtc), _p(_f32(bias, "bias")), _p(coarse), n, dc, hc, wc, a, b, _stream())
        return coarse
    _cabi.call("vs_k2s2_gather", _dt(fine), _p(fine), _p(_f32(wt, "wt")), _p(_f32(bias, "bias")), _p(coarse),
               n, dc, hc, wc, a, b, _stream())
    return coarse


def k2s2_scatter(coarse, wt, bias, cdims, a, b, wtc=None):
    n, dc, hc, wc = cdims
    fine = torch.empty(n, 2 * dc, 2 * hc, 2 * wc, b, device=coarse.device, dtype=coarse.dtype)
    if wtc is not None and coarse.dtype == torch.bfloat16:
        _cabi.call("vs_k2s2_scatter_tc", _p(coarse), _p(wtc), _p(_f32(bias, "bias")), _p(fine), n, dc, hc, wc, a, b, _stream())
        return fine
    _cabi.call("vs_k2s2_scatter", _dt(coarse), _p(coarse), _p(_f32(wt, "wt")), _p(_f32(bias, "bias")), _p(fine),
               n, dc, hc, wc, a, b, _stream())
    return fine


def k2s2_wgrad(coarse, fine, cdims, a, b, dwt=None, dbias_coarse=None, dbias_fine=None, accumulate=False):
    n, dc, hc, wc = cdims
    if dwt is None:
        dwt = torch.empty(a, b, 2, 2, 2, device=coarse.device, dtype=torch.float32)
        accumulate = False
    _cabi.call("vs_k2s2_wgrad", _dt(coarse), _p(coarse), _p(fine), _p(_f32(dwt, "dwt")), _p(_f32(dbias_coarse, "db")),
               _p(_f32(dbias_fine, "db")), int(accumulate), n, dc, hc, wc, a, b, _stream())
    return dwt


# ---- InstanceNorm + ReLU -------------------------------------------------------------
def inorm_relu_apply(y, stats, skip=None):
    n, c = y.shape[0], y.shape[-1]
    s = y.numel() // (n * c)
    a = torch.empty_like(y)
    _cabi.call("vs_inorm_relu_apply", _dt(y), _p(y), _p(stats), _p(skip), _p(a), n, s, c, _stream())
    return a


def inorm_relu_bwd(g, y, stats, sums=None, reduced=False):
    """Returns dy (gradient w.r.t. the raw conv output).  `sums`: pre-zeroed fp64 [N,C,2] block (StatsArena);
    reduced=True: `sums` already holds the reduction (fused into the producing dgrad), only the apply pass runs."""
    n, c = y.shape[0], y.shape[-1]
    s = y.numel() // (n * c)
    if not reduced:
        flags = VS_FLAG_PREZEROED if sums is not None else 0
        if sums is None:
            sums = torch.empty(n, c, 2, device=y.device, dtype=torch.float64)
        _cabi.call("vs_inorm_relu_bwd_reduce", _dt(y), _p(g), _p(y), _p(stats), _p(sums), n, s, c, flags, _stream())
    dy = torch.empty_like(y)
    _cabi.call("vs_inorm_relu_bwd_apply", _dt(y), _p(g), _p(y), _p(stats), _p(sums), _p(dy), n, s, c, _stream())
    return dy


# ---- per-(n,c) affine + ReLU (BatchNorm path) -------------------------------------------------------------------------
def affine_relu_apply(y, kb, skip=None):
    """a = relu(k*y + b) (+ skip); kb [N,C,2] fp32."""
    n, c = y.shape[0], y.shape[-1]
    s = y.numel() // (n * c)
    a = torch.empty_like(y)
    _cabi.call("vs_affine_relu_apply", _dt(y), _p(y), _p(_f32(kb, "kb")), _p(skip), _p(a), n, s, c, _stream())
    return a


def affine_relu_bwd_reduce(g, y, kb, k2b2):
    """sums [N,C,2] fp64 = (sum gm, sum gm*xhat) with gm = g*[k*y+b > 0], xhat = k2*y + b2."""
    n, c = y.shape[0], y.shape[-1]
    s = y.numel() // (n * c)
    sums = torch.empty(n, c, 2, device=y.device, dtype=torch.float64)
    _cabi.call("vs_affine_relu_bwd_reduce", _dt(y), _p(g), _p(y), _p(_f32(kb, "kb")), _p(_f32(k2b2, "k2b2")), _p(sums), n, s, c,
               _stream())
    return sums


def affine_relu_bwd_apply(g, y, kb, coef):
    """dy = c0*gm + c1 + c2*y; coef [N,C,3] fp32."""
    n, c = y.shape[0], y.shape[-1]
    s = y.numel() // (n * c)
    dy = torch.empty_like(y)
    _cabi.call("vs_affine_relu_bwd_apply", _dt(y), _p(g), _p(y), _p(_f32(kb, "kb")), _p(_f32(coef, "coef")), _p(dy), n, s, c,
               _stream())
    return dy


def add_inplace(dst, src):
    assert dst.dtype == src.dtype and dst.numel() == src.numel()
    _cabi.call("vs_add_inplace", _dt(dst), _p(dst), _p(src), dst.numel(), _stream())
    return dst


# ---- softmax over two classes --------------------------------------------------------
def softmax2_fwd(logits, dims):
    n, d, h, w = dims
    probs = torch.empty(n, 2, d, h, w, device=logits.device, dtype=torch.float32)
    _cabi.call("vs_softmax2_fwd", _p(_f32(logits, "logits")), _p(probs), n, d * h * w, _stream())
    return probs


def softmax2_bwd(dprobs, probs, dims, out_dtype):
    n, d, h, w = dims
    dlogits = torch.empty(n, d, h, w, 2, device=probs.device, dtype=out_dtype)
    _cabi.call("vs_softmax2_bwd", _dt(dlogits), _p(_f32(dprobs, "dprobs")), _p(_f32(probs, "probs")), _p(dlogits),
               n, d * h * w, _stream())
    return dlogits


def softmax2_bwd_pad8(dprobs, probs, dims, db=None):
    """Logit gradient as bf16 NDHWC [N,D,H,W,8] (channels 2..7 zero); db[2] (+)= the bias gradient."""
    n, d, h, w = dims
    dlogits8 = torch.empty(n, d, h, w, 8, device=probs.device, dtype=torch.bfloat16)
    _cabi.call("vs_softmax2_bwd_pad8", _p(_f32(dprobs, "dprobs")), _p(_f32(probs, "probs")), _p(dlogits8),
               _p(_f32(db, "db")), n, d * h * w, _stream())
    return dlogits8


def planar_to_ndhwc8(x):
    """[N,C,D,H,W] fp32 (C <= 8) -> [N,D,H,W,8] bf16, zero padded channels."""
    n, c = x.shape[0], x.shape[1]
    out = torch.empty(n, x.shape[2], x.shape[3], x.shape[4], 8, device=x.device, dtype=torch.bfloat16)
    _cabi.call("vs_planar_to_ndhwc8", _p(_f32(x, "x")), _p(out), n, c, x.numel() // (n * c), _stream())
    return out


# ---- VAE linear layers -----------------------------------------------------------------
def fc_encode_fwd(x, wm, bm, ws, bs, z, scale, use_z, batch, s3, c, dim):
    dev = x.device
    mean = torch.empty(batch, dim, device=dev, dtype=torch.float32)
    std = torch.empty_like(mean)
    lat = torch.empty_like(mean)
    _cabi.call("vs_fc_encode_fwd", _dt(x), _p(x), _p(wm), _p(bm), _p(ws), _p(bs), _p(z), float(scale), int(use_z),
               _p(mean), _p(std), _p(lat), batch, s3, c, dim, _stream())
    return mean, std, lat


def fc_decode_fwd(lat, w2, b2, batch, s3, c, dim, out_dtype, side):
    h = torch.empty(batch, side, side, side, c, device=lat.device, dtype=out_dtype)
    _cabi.call("vs_fc_decode_fwd", _dt(h), _p(_f32(lat, "lat")), _p(w2), _p(b2), _p(h), batch, s3, c, dim, _stream())
    return h


def fc_decode_bwd(dh, lat, w2, batch, s3, c, dim, dw2=None, db2=None, accumulate=False):
    dlat = torch.empty(batch, dim, device=dh.device, dtype=torch.float32)
    _cabi.call("vs_fc_decode_bwd", _dt(dh), _p(dh), _p(lat), _p(w2), _p(dlat), _p(dw2), _p(db2), int(accumulate),
               batch, s3, c, dim, _stream())
    return dlat


def fc_encode_bwd(x, wm, ws, z, scale, use_z, std, dlat, gmean_ext, gstd_ext, batch, s3, c, dim, want_dx=True,
                  dwm=None, dbm=None, dws=None, dbs=None, accumulate=False):
    dev = std.device
    gbuf = torch.empty(2, batch, dim, device=dev, dtype=torch.float32)
    dx = torch.empty_like(x) if want_dx else None
    _cabi.call("vs_fc_encode_bwd", _dt(x), _p(x), _p(wm), _p(ws), _p(z), float(scale), int(use_z), _p(std), _p(dlat),
               _p(gmean_ext), _p(gstd_ext), _p(gbuf), _p(dx), _p(dwm), _p(dbm), _p(dws), _p(dbs), int(accumulate),
               batch, s3, c, dim, _stream())
    return dx


ACT_NONE, ACT_RELU, ACT_SIGMOID = 0, 1, 2


def linear_fwd(x, w, bias, act=ACT_NONE):
    """y = act(x @ w.T + bias); x [B,in] fp32, w [out,in] fp32."""
    y = torch.empty(x.shape[0], w.shape[0], device=x.device, dtype=torch.float32)
    _cabi.call("vs_linear_fwd", _p(_f32(x, "x")), _p(_f32(w, "w")), _p(_f32(bias, "bias")), _p(y), x.shape[0], w.shape[1],
               w.shape[0], int(act), _stream())
    return y


def linear_bwd(x, w, y, dy, act=ACT_NONE, want_dx=True, dw=None, db=None, accumulate=False):
    """Returns dx (or None); dw / db (+)= in place when given."""
    gbuf = torch.empty_like(y)
    dx = torch.empty_like(x) if want_dx else None
    _cabi.call("vs_linear_bwd", _p(x), _p(w), _p(y), _p(_f32(dy, "dy")), _p(gbuf), _p(dx), _p(dw), _p(db), int(accumulate),
               x.shape[0], w.shape[1], w.shape[0], int(act), _stream())
    return dx


# ---- losses ----------------------------------------------------------------------------
def dice_sums(src, tgt, mode):
    n, c = src.shape[0], src.shape[1]
    s = src.numel() // (n * c)
    sums = torch.empty(n, c, 3, device=src.device, dtype=torch.float32)
    _cabi.call("vs_dice_sums", _p(_f32(src, "src")), _p(_f32(tgt, "tgt")), int(mode), _p(sums), n, c, s, _stream())
    return sums


def dice_bwd(src, tgt, mode, sums, gper, eps, want_src=True, want_tgt=False, gsrc=None, gtgt=None, accumulate=False):
    n, c = src.shape[0], src.shape[1]
    s = src.numel() // (n * c)
    if want_src and gsrc is None:
        gsrc = torch.empty_like(src)
    if want_tgt and gtgt is None:
        gtgt = torch.empty_like(src)
    _cabi.call("vs_dice_bwd", _p(src), _p(tgt), int(mode), _p(sums), _p(_f32(gper, "gper")), float(eps), _p(gsrc),
               _p(gtgt), int(accumulate), n, c, s, _stream())
    return gsrc, gtgt


def kl_fwd(mean, std):
    out = torch.empty(1, device=mean.device, dtype=torch.float32)
    _cabi.call("vs_kl_fwd", _p(_f32(mean, "mean")), _p(_f32(std, "std")), _p(out), mean.shape[0], mean.shape[1], _stream())
    return out


def kl_bwd(mean, std, gout):
    gm, gs = torch.empty_like(mean), torch.empty_like(std)
    _cabi.call("vs_kl_bwd", _p(mean), _p(std), _p(_f32(gout, "gout")), _p(gm), _p(gs), mean.shape[0], mean.shape[1], _stream())
    return gm, gs


def binarize(a, mode):
    out = torch.empty_like(a)
    _cabi.call("vs_binarize", _p(_f32(a, "a")), _p(out), int(mode), a.numel(), _stream())
    return out


def one_hot(label, n_class):
    n = label.shape[0]
    s = label.numel() // n
    out = torch.empty(n, n_class, *label.shape[2:], device=label.device, dtype=torch.float32)
    _cabi.call("vs_one_hot", _p(_f32(label, "label")), _p(out), n, n_class, s, _stream())
    return out


# ---- optimiser -------------------------------------------------------------------------
def sgd_step(p, g, buf, lr, momentum, first, gscale=1.0):
    _cabi.call("vs_sgd_step", _p(_f32(p, "p")), _p(_f32(g, "g")), _p(buf), p.numel(), float(lr), float(momentum),
               int(first), float(gscale), _stream())


def adam_step(p, g, m, v, lr, beta1, beta2, eps, step, gscale=1.0):
    _cabi.call("vs_adam_step", _p(_f32(p, "p")), _p(_f32(g, "g")), _p(m), _p(v), p.numel(), float(lr), float(beta1),
               float(beta2), float(eps), int(step), float(gscale), _stream())


def ema_update(teacher, student, alpha):
    _cabi.call("vs_ema_update", _p(_f32(teacher, "teacher")), _p(_f32(student, "student")), teacher.numel(),
               float(alpha), _stream())


def compose_target_loss(terms, lambda_vae, loss_type, use_kl):
    final = torch.empty(1, device=terms.device, dtype=torch.float32)
    weights = torch.empty(3, device=terms.device, dtype=torch.float32)
    _cabi.call("vs_compose_target_loss", _p(_f32(terms, "terms")), float(lambda_vae), int(loss_type), int(use_kl),
               _p(final), _p(weights), _stream())
    return final, weights


def joint_target_finish(sums_r, sums_g, sums_f, kl, bot, top, eps, lambda_vae, loss_type, use_kl, only_pseudo):
    """(out5, gper2): see vs_joint_target_finish."""
    n, c = sums_r.shape[0], sums_r.shape[1]
    out5 = torch.empty(5, device=sums_r.device, dtype=torch.float32)
    gper2 = torch.empty(2, n, c, device=sums_r.device, dtype=torch.float32)
    _cabi.call("vs_joint_target_finish", _p(sums_r), _p(sums_g), _p(sums_f), _p(_f32(kl, "kl")), n, c, int(bot), int(top),
               float(eps), float(lambda_vae), int(loss_type), int(use_kl), int(only_pseudo), _p(out5), _p(gper2), _stream())
    return out5, gper2


def atomic_add_rows(dst, src, rows, row_len, src_row_stride):
    """dst[r, :row_len] += src[r*src_row_stride : +row_len] with atomics (dst contiguous fp32; src a base pointer)."""
    _cabi.call("vs_atomic_add_rows", _p(_f32(dst, "dst")), src.data_ptr(), int(rows), int(row_len), int(src_row_stride), _stream())


def clip_center(x, lo, hi, sub, div):
    """(clip(x, lo, hi) - sub) / div as fp32; x: CUDA fp32 or int16 tensor (any shape, contiguous)."""
    kind = {torch.float32: 0, torch.int16: 2}.get(x.dtype)
    if kind is None:
        raise RuntimeError("vaeseg_b200: clip_center takes float32 or int16 volumes, got %s" % x.dtype)
    out = torch.empty(x.shape, device=x.device, dtype=torch.float32)
    _cabi.call("vs_clip_center", kind, _p(x), _p(out), x.numel(), float(lo), float(hi), float(sub), float(div), _stream())
    return out


def crop_resize(src, crop9, side, out_size, order=1, anti_alias=True):
    """src [D,H,W] fp32 CUDA volume -> [od,oh,ow]: crop (start[3], length[3], leading pad[3]) inside a zero cube of side
    `side`, then skimage-style resize (see vs_crop_resize).  Data-pipeline op (allocates its workspaces per call)."""
    import ctypes
    if src.dim() != 3:
        raise RuntimeError("vaeseg_b200: crop_resize takes one [D,H,W] volume")
    od, oh, ow = [int(v) for v in out_size]
    out = torch.empty(od, oh, ow, device=src.device, dtype=torch.float32)
    need_filter = anti_alias and any(side > o for o in (od, oh, ow))
    tmp0 = torch.empty(side ** 3, device=src.device, dtype=torch.float32) if need_filter else None
    tmp1 = torch.empty_like(tmp0) if need_filter else None
    arr = (ctypes.c_int * 9)(*[int(v) for v in crop9])
    _cabi.call("vs_crop_resize", _p(_f32(src, "volume")), src.shape[0], src.shape[1], src.shape[2], arr, int(side), _p(out),
               od, oh, ow, int(order), int(bool(anti_alias)), _p(tmp0), _p(tmp1), _stream())
    return out


# ---- kd-in-N tensor-core convolution (full-resolution forward layers, csrc/conv3_tc_kdn.cu) ----------------------
def pack_conv3_weight_tc_kdn(w, dgrad=False, out=None):
    cout, cin = w.shape[0], w.shape[1]
    nbytes = _cabi.lib().vs_conv3_tc_kdn_pack_bytes(cin, cout, int(dgrad))
    if nbytes == 0:
        return None
    if out is None:
        out = torch.empty(nbytes // 2, device=w.device, dtype=torch.bfloat16)
    _cabi.call("vs_pack_conv3_weight_tc_kdn", _p(_f32(w, "weight")), _p(out), cin, cout, int(dgrad), _stream())
    return out


def pack_conv3_weight_tc_kdn_padded(w, cin_pad, cout_pad, dgrad=False, out=None):
    """kd-in-N pack of w [Cout,Cin,3,3,3] with the channels zero-padded to cin_pad x cout_pad (2-class head, 2-channel
    in-block input gradient), or None when the kernel does not take the padded shape.  `out` re-packs in place."""
    cout, cin = w.shape[0], w.shape[1]
    nbytes = _cabi.lib().vs_conv3_tc_kdn_pack_bytes(cin_pad, cout_pad, int(dgrad))
    if nbytes == 0:
        return None
    if out is None:
        out = torch.empty(nbytes // 2, device=w.device, dtype=torch.bfloat16)
    _cabi.call("vs_pack_conv3_weight_tc_kdn_padded", _p(_f32(w, "weight")), _p(out), cin, cout, cin_pad, cout_pad, int(dgrad),
               _stream())
    return out


def conv3_tc_kdn_planar(x, wkdn8, dims, gin, mode, bias=None):
    """out [N,2,D,H,W] fp32 = channels 0..1 of an 8-output-channel kd-in-N convolution: mode 1 softmax(conv + bias) (the
    2-class head), mode 2 plain values (planar 2-channel input gradient)."""
    n, d, h, w = dims
    if x.dtype != torch.bfloat16:
        raise RuntimeError("vaeseg_b200: conv3_tc_kdn_planar takes bf16 NDHWC activations")
    out = torch.empty(n, 2, d, h, w, device=x.device, dtype=torch.float32)
    _cabi.call("vs_conv3x3x3_tc_kdn_planar", _p(x), _p(wkdn8), _p(out), _p(_f32(bias, "bias") if bias is not None else None),
               int(mode), n, d, h, w, gin, _stream())
    return out


def conv3_in_relu(x, wpack, dims, gin, gout, arena, skip=None, kdn=False):
    """(y, stats, a): Conv3d(3, padding 1) -> InstanceNorm3d -> ReLU (+ skip) in ONE cooperative tensor-core launch
    (vs_conv3x3x3_tc[_kdn]_in_relu): y = raw conv output (kept for backward), a = the activation.  bf16 NDHWC.
    `arena`: zero-filled StatsArena supplying statistics, shift and the barrier word."""
    n, d, h, w = dims
    y = torch.empty(n, d, h, w, gout, device=x.device, dtype=torch.bfloat16)
    a = torch.empty_like(y)
    buf = arena.take(stats_words(n, gout))
    stats = buf[:n * gout * 2].view(n, gout, 2)
    shift = buf[n * gout * 2:].view(torch.float32)[:n * gout].view(n, gout)
    gbar = buf[-1:]
    _cabi.call("vs_conv3x3x3_tc_kdn_in_relu" if kdn else "vs_conv3x3x3_tc_in_relu", _p(x), _p(wpack), _p(y), _p(a), _p(skip),
               _p(stats), _p(shift), _p(gbar), n, d, h, w, gin, gout, _stream())
    return y, stats, a


def conv3_tc_kdn(x, wkdn, dims, gin, gout, want_stats=False, arena=None, prev=None):
    """y (bf16 NDHWC) = conv3(x, wkdn) through the kd-in-N kernel; returns (y, stats or None).
    arena: zero-filled StatsArena supplying the statistics / shift words (as conv3_fprop).  prev = (y_prev, stats_prev,
    sums_prev) (dgrad): also accumulate the previous layer's norm-backward sums (sums pre-zeroed), as conv3_dgrad."""
    n, d, h, w = dims
    y = torch.empty(n, d, h, w, gout, device=x.device, dtype=torch.bfloat16)
    stats = shift = None
    prezeroed = 0
    if want_stats:
        if arena is not None:
            buf = arena.take(stats_words(n, gout))
            prezeroed = 1
        else:
            buf = torch.empty(stats_words(n, gout), device=x.device, dtype=torch.float64)
        stats = buf[:n * gout * 2].view(n, gout, 2)
        shift = buf[n * gout * 2:].view(torch.float32)[:n * gout].view(n, gout)
    yp, sp, sums = prev if prev is not None else (None, None, None)
    _cabi.call("vs_conv3x3x3_tc_kdn_ex", _p(x), _p(wkdn), _p(y), _p(stats), _p(shift), prezeroed, _p(yp), _p(sp), _p(sums),
               n, d, h, w, gin, gout, _stream())
    return y, stats
