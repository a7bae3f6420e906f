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


class _JointTargetLossFn(torch.autograd.Function):
    """The loss tail of a teacher-student step as ONE autograd node: three fused Dice reductions, one scalar kernel
    (losses, dynamic-lambda composition, per-Dice gradients), and in backward two fused elementwise passes -- instead
    of ~45 tiny ATen kernels.  main_target.py:543-546 (terms), :550-560 / :588-590 (composition)."""

    @staticmethod
    def forward(ctx, pred, recon, label, teacher_pred, kl, cfg):
        lambda_vae, loss_type, use_kl, only_pseudo, fake_mode, eps = cfg
        sums_r = ops.dice_sums(pred, recon, TGT_TENSOR)
        sums_g = ops.dice_sums(pred, label, TGT_LABEL) if label is not None else None
        sums_f = ops.dice_sums(pred, teacher_pred, fake_mode)
        out5, gper2 = ops.joint_target_finish(sums_r, sums_g, sums_f, kl, 1, 2, eps, lambda_vae, loss_type, use_kl, only_pseudo)
        ctx.save_for_backward(pred, recon, teacher_pred, sums_r, sums_f, gper2)
        ctx.cfg = (fake_mode, eps)
        mon = out5.detach()
        ctx.mark_non_differentiable(mon)
        return out5[0], mon

    @staticmethod
    def backward(ctx, g, _gmon):
        pred, recon, teacher_pred, sums_r, sums_f, gper2 = ctx.saved_tensors
        fake_mode, eps = ctx.cfg
        gper = gper2 * g
        want_pred, want_recon = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
        gpred, grecon = ops.dice_bwd(pred, recon, TGT_TENSOR, sums_r, gper[0], eps, want_src=want_pred, want_tgt=want_recon)
        if want_pred:
            ops.dice_bwd(pred, teacher_pred, fake_mode, sums_f, gper[1], eps, want_src=True, gsrc=gpred, accumulate=True)
        return gpred, grecon, None, None, None, None


def joint_target_loss(pred, recon, label, teacher_pred, kl=None, lambda_vae=1.0, loss_type=0, use_kl=False,
                      only_pseudo=False, confident=False, eps=0.000001):
    """(final, mon) with mon = [final, recon_loss, dice_loss, dice_loss_fake, kl] (detached).  Single-process form of
    the loss tail of main_target.py:543-603; with loss_type 8 under data parallelism use the composed path of
    train_step.JointTrainer (the recon term is all-reduced between the reductions and the composition)."""
    pred = _prep(pred, "joint_target_loss")
    recon = _prep(recon, "joint_target_loss")
    teacher_pred = _prep(teacher_pred, "joint_target_loss").detach()
    label = _prep(label, "joint_target_loss").detach() if label is not None else None
    kl = kl.detach().reshape(1).float().contiguous() if kl is not None else None
    cfg = (float(lambda_vae), int(loss_type), bool(use_kl), bool(only_pseudo), TGT_CONFIDENT if confident else TGT_BINARIZE, eps)
    return _JointTargetLossFn.apply(pred, recon, label, teacher_pred, kl, cfg)


def avg_ce(data_dict, source_key='align_lung', target_key='source_lung'):
    raise NotImplementedError("avg_ce (BCE) is not on the training hot path: no shipped preset calls it "
                              "(SURVEY.md section 2); not implemented in vaeseg_b200")
