"""TEST INFRASTRUCTURE ONLY -- CPU fp32 restatement of the reference hot path.

Functional (state_dict in, tensors out) restatement of the reference's training hot
path, written against torch.nn.functional on CPU fp32.  Every function cites the
reference file:line it follows (paths relative to /root/reference).

Parity pinning: the reference ships no tests / golden vectors, so this
restatement is pinned against the real reference modules executed in the authoring
container (tests/test_oracle_vs_reference.py, skipped when /root/reference is absent)
and against fixtures generated from the real reference (tests/golden/, produced by
oracle/make_golden.py).  The only deliberate deviation from the reference is the
generalised VAE flat dimension 256*(P/32)^3 (joint_model.py:216-218,241,253 hard-code
16384 = 128^3 patches); at P=128 it is identical.

Nothing in the shipped package imports this file.
"""
import math
from collections import OrderedDict

import torch
import torch.nn.functional as F

N_FMAPS = [8, 16, 32, 64, 128, 256]          # joint_model.py:207,352
IN_EPS = 1e-5                                 # nn.InstanceNorm3d default, joint_model.py:11


# ----------------------------------------------------------------------------------
# parameter initialisation in the reference's construction order (same RNG stream)
# ----------------------------------------------------------------------------------
def _init_conv(sd, name, w_shape, transposed=False):
    """torch.nn.Conv3d / ConvTranspose3d / Linear default reset_parameters():
    kaiming_uniform_(a=sqrt(5)) on the weight, U(+-1/sqrt(fan_in)) on the bias, where
    fan_in = weight.size(1) * receptive field (so Cout*8 for ConvTranspose3d)."""
    w = torch.empty(w_shape)
    torch.nn.init.kaiming_uniform_(w, a=math.sqrt(5))
    fan_in = w_shape[1]
    for k in w_shape[2:]:
        fan_in *= k
    bound = 1.0 / math.sqrt(fan_in)
    n_bias = w_shape[1] if transposed else w_shape[0]
    b = torch.empty(n_bias).uniform_(-bound, bound)
    sd[name + ".weight"] = w
    sd[name + ".bias"] = b


def _init_double_conv(sd, prefix, cin, cout):
    # joint_model.py:35-52: three Conv3d(3, padding=1) at Sequential indices 0, 3, 6
    _init_conv(sd, prefix + ".conv.0", (cout, cin, 3, 3, 3))
    _init_conv(sd, prefix + ".conv.3", (cout, cout, 3, 3, 3))
    _init_conv(sd, prefix + ".conv.6", (cout, cout, 3, 3, 3))


def _init_down(sd, name, cin, cout):
    # joint_model.py:126-136
    _init_conv(sd, name + ".conv.0", (cin, cin, 2, 2, 2))
    _init_double_conv(sd, name + ".conv.1", cin, cout)


def _init_up(sd, name, cin, cout):
    # joint_model.py:114-124; ConvTranspose3d weight is [Cin, Cout, 2, 2, 2]
    _init_conv(sd, name + ".conv.0", (cin, cin, 2, 2, 2), transposed=True)
    _init_double_conv(sd, name + ".conv.1", cin, cout)


def init_seg_state(n_channels=1, n_class=2):
    """state_dict of Segmentation(n_channels, n_class, norm_type=1) in construction
    order (joint_model.py:352-366).  Uses the global CPU generator like nn.Module init."""
    f = N_FMAPS
    sd = OrderedDict()
    _init_conv(sd, "in_block.conv.0", (f[0], n_channels, 3, 3, 3))
    for i in range(4):
        _init_down(sd, "down%d" % (i + 1), f[i], f[i + 1])
    for i, lvl in enumerate([4, 3, 2, 1]):
        _init_up(sd, "up%d" % (i + 2), f[lvl], f[lvl - 1])
    _init_conv(sd, "out_block", (n_class, f[0], 3, 3, 3))
    return sd


def vae_flat_dim(patch):
    assert patch % 32 == 0, "VAE needs five stride-2 levels"
    return N_FMAPS[5] * (patch // 32) ** 3


def init_vae_state(n_class=2, dim=128, patch=128):
    """state_dict of VAE(n_channels, n_class, norm_type=1, dim=dim) in construction
    order (joint_model.py:210-224); flat dim generalised (16384 at patch=128)."""
    f = N_FMAPS
    flat = vae_flat_dim(patch)
    sd = OrderedDict()
    _init_conv(sd, "in_block.conv.0", (f[0], n_class, 3, 3, 3))
    for i in range(5):
        _init_down(sd, "down%d" % (i + 1), f[i], f[i + 1])
    _init_conv(sd, "fc_mean", (dim, flat))
    _init_conv(sd, "fc_std", (dim, flat))
    _init_conv(sd, "fc2", (flat, dim))
    for i, lvl in enumerate([5, 4, 3, 2, 1]):
        _init_up(sd, "up%d" % (i + 1), f[lvl], f[lvl - 1])
    _init_conv(sd, "out_block", (n_class, f[0], 3, 3, 3))
    return sd


# BatchNorm layers (norm_type 2) use running statistics when True (module.eval()); see conv_in_relu
BN_EVAL = False

# ----------------------------------------------------------------------------------
# blocks
# ----------------------------------------------------------------------------------
def conv_in_relu(sd, name, x):
    """Conv3d(3,p=1) -> InstanceNorm3d(affine=False, eps=1e-5, biased var) -> ReLU
    (joint_model.py:101-112 and each third of :35-52)."""
    y = F.conv3d(x, sd[name + ".weight"], sd[name + ".bias"], padding=1)
    head, _, last = name.rpartition(".")
    norm = "%s.%d" % (head, int(last) + 1)
    if norm + ".running_mean" in sd:
        # norm_type 2: nn.BatchNorm3d(C, momentum=0.1) at the next Sequential index (joint_model.py:12-13).  Training mode
        # (batch statistics; the running buffers of `sd` are updated in place like the module's) unless BN_EVAL is set.
        y = F.batch_norm(y, sd[norm + ".running_mean"], sd[norm + ".running_var"], sd[norm + ".weight"], sd[norm + ".bias"],
                         training=not BN_EVAL, momentum=0.1, eps=IN_EPS)
    else:
        y = F.instance_norm(y, eps=IN_EPS)
    return F.relu(y)


def double_conv(sd, prefix, x):
    # joint_model.py:35-52 -- three conv+IN+ReLU despite the name
    for idx in (0, 3, 6):
        x = conv_in_relu(sd, "%s.conv.%d" % (prefix, idx), x)
    return x


def down(sd, name, x):
    # joint_model.py:126-136: Conv3d(C,C,2,stride 2) (no norm/act) -> DoubleConv
    x = F.conv3d(x, sd[name + ".conv.0.weight"], sd[name + ".conv.0.bias"], stride=2)
    return double_conv(sd, name + ".conv.1", x)


def up(sd, name, x):
    # joint_model.py:114-124: ConvTranspose3d(C,C,2,stride 2) (no norm/act) -> DoubleConv
    x = F.conv_transpose3d(x, sd[name + ".conv.0.weight"], sd[name + ".conv.0.bias"], stride=2)
    return double_conv(sd, name + ".conv.1", x)


# ----------------------------------------------------------------------------------
# models
# ----------------------------------------------------------------------------------
def seg_forward(sd, x, return_logits=False):
    """Segmentation.forward with dropout=0 (joint_model.py:369-390).  Skips are additive
    and only at two levels: up3(.)+x3, up4(.)+x2."""
    x1 = conv_in_relu(sd, "in_block.conv.0", x)
    x2 = down(sd, "down1", x1)
    x3 = down(sd, "down2", x2)
    x4 = down(sd, "down3", x3)
    x5 = down(sd, "down4", x4)
    y = up(sd, "up2", x5)
    y = up(sd, "up3", y) + x3
    y = up(sd, "up4", y) + x2
    y = up(sd, "up5", y)
    logits = F.conv3d(y, sd["out_block.weight"], sd["out_block.bias"], padding=1)
    probs = F.softmax(logits, dim=1)
    return (probs, logits) if return_logits else probs


def vae_encode(sd, x):
    # joint_model.py:233-243 (flat dim generalised)
    h = conv_in_relu(sd, "in_block.conv.0", x)
    for i in range(1, 6):
        h = down(sd, "down%d" % i, h)
    h = h.reshape(h.size(0), -1)
    mean = F.linear(h, sd["fc_mean.weight"], sd["fc_mean.bias"])
    std = F.relu(F.linear(h, sd["fc_std.weight"], sd["fc_std.bias"]))
    return mean, std


def vae_decode(sd, lat, return_logits=False):
    # joint_model.py:250-266 (view generalised to (256, P/32, P/32, P/32)); dropout = 0
    h = F.linear(lat, sd["fc2.weight"], sd["fc2.bias"])
    s = round((h.size(1) // N_FMAPS[5]) ** (1.0 / 3.0))
    h = h.view(h.size(0), N_FMAPS[5], s, s, s)
    for i in range(1, 6):
        h = up(sd, "up%d" % i, h)
    logits = F.conv3d(h, sd["out_block.weight"], sd["out_block.bias"], padding=1)
    probs = F.softmax(logits, dim=1)
    return (probs, logits) if return_logits else probs


def vae_forward(sd, x, if_random=False, scale=1, mid_input=False, z=None):
    """VAE.forward (joint_model.py:227-272).  z is drawn from the CPU default generator
    on every non-mid_input call, used or not (joint_model.py:246)."""
    if mid_input:
        return vae_decode(sd, x)
    mean, std = vae_encode(sd, x)
    if z is None:
        z = torch.randn(mean.size(0), mean.size(1)).to(mean.dtype)
    lat = mean + z * std * scale if if_random else mean
    return vae_decode(sd, lat), mean, std


def encoder_forward(sd, x):
    """Encoder.forward (joint_model.py:290-305): conv trunk -> fc1 -> ReLU -> fc2 -> ReLU -> fc_mean -> sigmoid; the flat
    dimension is generalised like the VAE's (16384 at 128^3 patches)."""
    h = conv_in_relu(sd, "in_block.conv.0", x)
    for i in range(1, 6):
        h = down(sd, "down%d" % i, h)
    h = h.reshape(h.size(0), -1)
    h = F.relu(F.linear(h, sd["fc1.weight"], sd["fc1.bias"]))
    h = F.relu(F.linear(h, sd["fc2.weight"], sd["fc2.bias"]))
    return torch.sigmoid(F.linear(h, sd["fc_mean.weight"], sd["fc_mean.bias"]))


def fusion_forward(sd, img, mask):
    """Fusion.forward (joint_model.py:415-436)."""
    x2 = down(sd, "down1", conv_in_relu(sd, "in_block.conv.0", img)) + \
        down(sd, "down1_mask", conv_in_relu(sd, "in_block_mask.conv.0", mask))
    x2 = conv_in_relu(sd, "merge.conv.0", x2)
    x3 = down(sd, "down2", x2)
    x5 = down(sd, "down4", down(sd, "down3", x3))
    y = up(sd, "up2", x5)
    y = up(sd, "up3", y) + x3
    y = up(sd, "up4", y) + x2
    y = up(sd, "up5", y)
    return F.softmax(F.conv3d(y, sd["out_block.weight"], sd["out_block.bias"], padding=1), dim=1)


def joint_forward(seg_sd, vae_sd, x, dropout=False, vae_forward_scale=0.0):
    """Joint.forward (joint_model.py:447-452).  With dropout=True the student's mean/std
    are discarded; decoder/seg dropout probabilities are 0 in every shipped
    script so the dropout branches are no-ops here."""
    pred = seg_forward(seg_sd, x)
    recon, mean, std = vae_forward(vae_sd, pred, if_random=False, scale=vae_forward_scale)
    if dropout:
        return pred, recon, None, None
    return pred, recon, mean, std


# ----------------------------------------------------------------------------------
# losses (utils/evaluation.py; main_source.py:133-182 is the eps=1e-4 twin)
# ----------------------------------------------------------------------------------
def dice(a, b):
    # utils/evaluation.py:6-7
    return 2.0 * torch.sum(a * b) / (torch.sum(a) + torch.sum(b) + 0.000001)


def binarize(a):
    # utils/evaluation.py:9-10
    return (a >= 0.5).float()


def confident_binarize(a, hi=0.8, lo=0.2):
    # utils/evaluation.py:12-18
    b = a.clone()
    b[b > hi] = 1
    b[b < lo] = 0
    return b


def one_hot(label, n_class=2):
    # main_target.py:520-522 / main_source.py:390-392: zeros.scatter_(1, label.long(), 1)
    lab = label.long()
    out = torch.zeros(lab.size(0), n_class, *lab.shape[2:], dtype=label.dtype if label.is_floating_point() else torch.float32)
    return out.scatter_(1, lab, 1)


def kl_loss(mean, std):
    # utils/evaluation.py:42-45
    return torch.mean(0.5 * (torch.sum(std ** 2, 1) + torch.sum(mean ** 2, 1)
                             - 2 * torch.sum(torch.log(std + 0.00001), 1)))


def avg_dsc(source, target, binary=False, topindex=2, botindex=0, return_mean=True,
            detach=False, eps=0.000001):
    """utils/evaluation.py:48-80 (eps 1e-6) / main_source.py:150-182 (eps 1e-4)."""
    if detach:
        target = target.detach()
    if binary:
        source = one_hot(torch.argmax(source, dim=1, keepdim=True), source.size(1))
        target = one_hot(torch.argmax(target, dim=1, keepdim=True), target.size(1))
    per = 2 * torch.sum(source * target, (2, 3, 4)) / (
        torch.sum(source, (2, 3, 4)) + torch.sum(target, (2, 3, 4)) + eps)
    if source.shape[1] > 1:
        per = per[:, botindex:topindex]
    return torch.mean(per) if return_mean else torch.mean(per, 1)


def dynamic_lambda(recon_loss, lambda_vae):
    # main_target.py:550-554 (type 8 thresholds on the host value of recon_loss)
    r = float(recon_loss)
    if r < 0.15:
        return lambda_vae * 0.6
    if r < 0.225:
        return lambda_vae * 1.2
    if r < 0.3:
        return lambda_vae * 2.0
    return lambda_vae * 3.0


def compose_target_loss(recon_loss, dsc_loss_fake, klloss, lambda_vae=1.0, loss_type=0,
                        kl=False, only_pseudo=False):
    """main_target.py:548-592 for the shipped presets: only_pseudo, type 8 (dynamic
    lambda) and the default branch with epoch >= warm-up (type 0)."""
    if only_pseudo:
        return dsc_loss_fake
    if loss_type == 8:
        cur = dynamic_lambda(recon_loss, lambda_vae)
        if cur > 1:
            if kl:
                return recon_loss + klloss + 1 / cur * dsc_loss_fake
            return recon_loss + 1 / cur * dsc_loss_fake
        if kl:
            return cur * (recon_loss + klloss) + dsc_loss_fake
        return cur * recon_loss + dsc_loss_fake
    loss = lambda_vae * recon_loss + dsc_loss_fake
    if kl:
        loss = loss + 0.00002 * lambda_vae * klloss
    return loss


# ----------------------------------------------------------------------------------
# train steps
# ----------------------------------------------------------------------------------
_BN_BUFFERS = ("running_mean", "running_var", "num_batches_tracked")


def _leafify(sd, requires_grad=True, dtype=None):
    out = OrderedDict()
    for k, v in sd.items():
        if k.rsplit(".", 1)[-1] in _BN_BUFFERS:          # BatchNorm buffers (norm_type 2): no gradient, updated in place
            out[k] = v.detach().clone().to(dtype if (dtype is not None and v.is_floating_point()) else v.dtype)
        else:
            out[k] = v.detach().clone().to(dtype or v.dtype).requires_grad_(requires_grad)
    return out


def sgd_step(sd, grads, bufs, lr=1e-2, momentum=0.9):
    """torch.optim.SGD(lr, momentum, dampening 0, wd 0, nesterov False) as used at
    main_target.py:347-352 / main_source.py:279-280.  bufs: dict or None (first step)."""
    new_sd, new_bufs = OrderedDict(), OrderedDict()
    for k, p in sd.items():
        g = grads[k]
        if momentum != 0:
            b = g.clone() if (bufs is None or k not in bufs) else momentum * bufs[k] + g
            new_bufs[k] = b
            g = b
        new_sd[k] = p.detach() - lr * g
    return new_sd, new_bufs


def ema_update(teacher_sd, student_sd, alpha=0.995):
    # main_target.py:512-516
    return OrderedDict((k, alpha * teacher_sd[k] + (1 - alpha) * student_sd[k]) for k in student_sd)


def seg_train_step(seg_sd, img, label, eps=0.0001, dtype=None):
    """main_source.py:415-437: loss = 1 - avg_dsc(pred, onehot)[fg]; returns loss, grads, pred.
    dtype=torch.float64 gives the high-precision "truth" used to calibrate fp32 noise."""
    sd = _leafify(seg_sd, dtype=dtype)
    if dtype is not None:
        img, label = img.to(dtype), label.to(dtype)
    pred = seg_forward(sd, img)
    loss = 1 - avg_dsc(pred, one_hot(label), botindex=1, topindex=2, eps=eps)
    loss.backward()
    return loss.detach(), OrderedDict((k, v.grad) for k, v in sd.items()), pred.detach()


def vae_train_step(vae_sd, label, scale=0.35, z=None, eps=0.0001, dtype=None):
    """main_source.py:389-406: recon = VAE(onehot, if_random=True, scale=.35);
    loss = 1 - avg_dsc(recon, onehot)[fg] + 2e-5 * KL."""
    sd = _leafify(vae_sd, dtype=dtype)
    if dtype is not None:
        label = label.to(dtype)
        z = z.to(dtype) if z is not None else None
    oh = one_hot(label)
    recon, mean, std = vae_forward(sd, oh, if_random=True, scale=scale, z=z)
    kl = kl_loss(mean, std)
    dsc = 1 - avg_dsc(recon, oh, botindex=1, topindex=2, eps=eps)
    loss = dsc + 0.00002 * kl
    loss.backward()
    grads = OrderedDict((k, v.grad) for k, v in sd.items())
    return loss.detach(), dsc.detach(), kl.detach(), grads, recon.detach()


def joint_target_step(student_seg_sd, vae_sd, teacher_seg_sd, img, label, lambda_vae=1.0,
                      loss_type=0, kl=False, confident=False, only_pseudo=False, dtype=None):
    """main_target.py:520-592,734-736: student Joint fwd (dropout=True), teacher Joint
    fwd (sets mean/std, F8), pseudo = binarize(teacher pred), recon / pseudo Dice,
    backward through the frozen VAE into Seg (F9)."""
    sd = _leafify(student_seg_sd, dtype=dtype)
    vsd = _leafify(vae_sd, requires_grad=False, dtype=dtype)
    tsd = _leafify(teacher_seg_sd, requires_grad=False, dtype=dtype)
    if dtype is not None:
        img, label = img.to(dtype), label.to(dtype)
    pred, recon, _, _ = joint_forward(sd, vsd, img, dropout=True)
    with torch.no_grad():
        t_pred, _, t_mean, t_std = joint_forward(tsd, vsd, img, dropout=False)
    pseudo = confident_binarize(t_pred) if confident else binarize(t_pred)
    recon_loss = 1 - avg_dsc(pred, recon, botindex=1, topindex=2)
    klloss = kl_loss(t_mean, t_std)
    dsc_loss = 1 - avg_dsc(pred, one_hot(label), botindex=1, topindex=2)
    dsc_loss_fake = 1 - avg_dsc(pred, pseudo, botindex=1, topindex=2)
    final = compose_target_loss(recon_loss, dsc_loss_fake, klloss, lambda_vae, loss_type, kl, only_pseudo)
    final.backward()
    grads = OrderedDict((k, v.grad) for k, v in sd.items())
    out = dict(final=final.detach(), recon_loss=recon_loss.detach(), dsc_loss=dsc_loss.detach(),
               dsc_loss_fake=dsc_loss_fake.detach(), klloss=klloss.detach(),
               pred=pred.detach(), recon=recon.detach(), pseudo=pseudo)
    return out, grads


def clip_center(x, new_min=-200.0, new_max=400.0, subtrahend=100.0, divisor=300.0):
    """numpy restatement of the reference's Clip -> CenterIntensities dataset transforms on a float32 volume
    (utils/utils.py:508-533: np.clip(val, new_min, new_max); :572-618: val -= subtrahend; val /= divisor, in place on the
    float32 array the loader produced with .astype(np.float32)).  Test infrastructure only."""
    import numpy as np
    val = np.asarray(x).astype(np.float32)
    val = np.clip(val, new_min, new_max)
    val -= subtrahend
    val /= divisor
    return val
