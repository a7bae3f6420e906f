"""Drop-in replacements for the reference's utils/evaluation.py losses.

Same names, signatures and values as /root/reference/utils/evaluation.py (dice :6-7,
binarize :9-10, confident_binarize :12-18, avg_ce :29-39, KLloss :42-45, avg_dsc :48-80);
`eps` selects the main_source.py:150-182 twin (1e-4).  Each Dice evaluation is ONE fused
reduction pass over (source, target) plus one fused elementwise backward pass, instead of
the reference's three reductions and temporaries; binarize / one-hot / argmax variants are
applied on the fly (`avg_dsc_fused`).
"""
import torch

from . import ops
from ._cabi import TGT_ARGMAX, TGT_BINARIZE, TGT_CONFIDENT, TGT_LABEL, TGT_TENSOR


def _prep(t, what):
    if not torch.is_tensor(t) or not t.is_cuda:
        raise RuntimeError("vaeseg_b200.%s: tensors must live on CUDA (no CPU fallback)" % what)
    if t.dtype != torch.float32:
        t = t.float()
    return t.contiguous()


class _DiceFn(torch.autograd.Function):
    """per[n,c] = 2*sum(s*t') / (sum(s) + sum(t') + eps) over the spatial dims."""

    @staticmethod
    def forward(ctx, src, tgt, mode, eps):
        sums = ops.dice_sums(src, tgt, mode)
        ctx.save_for_backward(src, tgt, sums)
        ctx.mode, ctx.eps = mode, eps
        return 2.0 * sums[..., 0] / (sums[..., 1] + sums[..., 2] + eps)

    @staticmethod
    def backward(ctx, gper):
        src, tgt, sums = ctx.saved_tensors
        want_src = ctx.needs_input_grad[0]
        want_tgt = ctx.needs_input_grad[1]
        if ctx.mode == TGT_ARGMAX:
            return None, None, None, None           # argmax is piecewise constant
        if want_tgt and ctx.mode != TGT_TENSOR:
            want_tgt = False                        # thresholded / label targets are constants
        gsrc, gtgt = ops.dice_bwd(src, tgt, ctx.mode, sums, gper.contiguous().float(), ctx.eps,
                                  want_src=want_src, want_tgt=want_tgt)
        return gsrc, gtgt, None, None


def _reduce(per, n_channels, botindex, topindex, return_mean):
    if n_channels > 1:
        per = per[:, botindex:topindex]
    return torch.mean(per) if return_mean else torch.mean(per, 1)


def avg_dsc(data_dict, source_key='align_lung', target_key='source_lung', binary=False, topindex=2, botindex=0,
            pad=[0, 0, 0], return_mean=True, detach=False, eps=0.000001):
    source = _prep(data_dict[source_key], "avg_dsc")
    target = _prep(data_dict[target_key], "avg_dsc")
    if detach:
        target = target.detach()
    per = _DiceFn.apply(source, target, TGT_ARGMAX if binary else TGT_TENSOR, eps)
    return _reduce(per, source.shape[1], botindex, topindex, return_mean)


def avg_dsc_fused(source, target, target_mode="tensor", topindex=2, botindex=0, return_mean=True, eps=0.000001):
    """avg_dsc with the target transform fused into the reduction: target_mode in
    {'tensor', 'binarize', 'confident', 'label'}; 'label' takes the integer label volume
    [B,1,D,H,W] (float) instead of a materialised one-hot tensor."""
    mode = {"tensor": TGT_TENSOR, "binarize": TGT_BINARIZE, "confident": TGT_CONFIDENT, "label": TGT_LABEL}[target_mode]
    source = _prep(source, "avg_dsc_fused")
    target = _prep(target, "avg_dsc_fused")
    per = _DiceFn.apply(source, target.detach() if mode != TGT_TENSOR else target, mode, eps)
    return _reduce(per, source.shape[1], botindex, topindex, return_mean)


def dice(A, B):
    a = _prep(A, "dice").reshape(1, 1, -1)
    b = _prep(B, "dice").reshape(1, 1, -1)
    return _DiceFn.apply(a, b, TGT_TENSOR, 0.000001).reshape(())


def binarize(A):
    return ops.binarize(_prep(A, "binarize").detach(), TGT_BINARIZE)


def confident_binarize(A, max=0.8, min=0.2):
    if max != 0.8 or min != 0.2:
        raise NotImplementedError("confident_binarize: only the reference thresholds (0.8, 0.2) are implemented")
    return ops.binarize(_prep(A, "confident_binarize").detach(), TGT_CONFIDENT)


def one_hot(label, n_class=2):
    """zeros.scatter_(1, label.long(), 1) of main_target.py:520-522 for a float label volume."""
    return ops.one_hot(_prep(label, "one_hot"), n_class)


class _KLFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, mean, std):
        ctx.save_for_backward(mean, std)
        return ops.kl_fwd(mean, std).reshape(())

    @staticmethod
    def backward(ctx, g):
        mean, std = ctx.saved_tensors
        gm, gs = ops.kl_bwd(mean, std, g.reshape(1).contiguous().float())
        return gm, gs


def KLloss(data_dict, mean_key='mean', std_key='std'):
    return _KLFn.apply(_prep(data_dict[mean_key], "KLloss"), _prep(data_dict[std_key], "KLloss"))


def avg_ce(data_dict, source_key='align_lung', target_key='source_lung'):
    raise NotImplementedError("avg_ce (BCE) is not on the training hot path: no shipped preset calls it "
                              "(SURVEY.md section 2); not implemented in vaeseg_b200")
