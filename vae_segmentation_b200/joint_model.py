"""Drop-in replacements for the reference's joint_model.py hot-path classes.

Same class names, constructor / forward signatures, attribute names and state_dict keys,
shapes and dtypes as /root/reference/joint_model.py (Normalization :9-15, DoubleConv :35-52,
Conv :101-112, Up :114-124, Down :126-136, VAE :204-272, Segmentation :349-390, Joint
:438-452), so reference checkpoints load with strict=True and the drivers' calls
(`model(batch, in_key, out_key, recon_key, dropout=True)`, `vae(x, if_random=..., scale=...)`)
work unchanged.  Parameters stay fp32 in PyTorch layouts; all arithmetic runs in the
hand-written sm_100a kernels of libvaeseg_b200.so through `engine`.  There is no torch.nn
compute and no CPU fallback: CPU tensors raise.

Differences, all explicit:
  * norm_type=1 (InstanceNorm3d, the configuration every shipped script uses) and norm_type=2
    (BatchNorm3d, the constructors' default) are implemented; 3 (GSNorm3d) raises
    NotImplementedError.
  * VAE takes an extra keyword `patch` (default 128): the reference hard-codes the 128^3
    flat dimension 16384 (joint_model.py:216-218,241,253); flat = 256*(patch/32)^3 here.
  * dropout probabilities other than 0 raise (every shipped preset uses 0).
  * `set_precision('bf16'|'fp32')` selects bf16 storage (default) or the fp32 check mode.
"""
import math
import os

import torch
import torch.nn as nn

from . import engine
from .engine import C3IN, HEAD, K2DOWN, K2UP, Layer

_DEFAULT_FMAPS = [8, 16, 32, 64, 128, 256]
_PRECISIONS = {"bf16": torch.bfloat16, "fp32": torch.float32}


def default_precision():
    return os.environ.get("VAESEG_PRECISION", "bf16")


# ---- parameter holders (PyTorch default initialisation, same RNG stream as nn.ConvNd) ----
class _ConvParams(nn.Module):
    """weight/bias of a Conv3d ([Cout,Cin,k,k,k]) or ConvTranspose3d ([Cin,Cout,k,k,k])."""

    def __init__(self, in_ch, out_ch, k, transposed=False):
        super().__init__()
        self.in_ch, self.out_ch, self.k, self.transposed = in_ch, out_ch, k, transposed
        shape = (in_ch, out_ch, k, k, k) if transposed else (out_ch, in_ch, k, k, k)
        self.weight = nn.Parameter(torch.empty(shape))
        self.bias = nn.Parameter(torch.empty(out_ch))
        self.reset_parameters()

    def reset_parameters(self):
        nn.init.kaiming_uniform_(self.weight, a=math.sqrt(5))
        fan_in = self.weight.size(1) * self.k ** 3
        bound = 1.0 / math.sqrt(fan_in)
        nn.init.uniform_(self.bias, -bound, bound)

    def extra_repr(self):
        return "%d, %d, kernel_size=%d%s" % (self.in_ch, self.out_ch, self.k, ", transposed" if self.transposed else "")


class _LinearParams(nn.Module):
    def __init__(self, in_features, out_features):
        super().__init__()
        self.in_features, self.out_features = in_features, out_features
        self.weight = nn.Parameter(torch.empty(out_features, in_features))
        self.bias = nn.Parameter(torch.empty(out_features))
        nn.init.kaiming_uniform_(self.weight, a=math.sqrt(5))
        bound = 1.0 / math.sqrt(in_features)
        nn.init.uniform_(self.bias, -bound, bound)

    def extra_repr(self):
        return "in_features=%d, out_features=%d" % (self.in_features, self.out_features)


class _Fused(nn.Module):
    """Placeholder keeping the reference's nn.Sequential indices (norm / activation slots);
    the operation itself is fused into the neighbouring kernels."""

    def __init__(self, what):
        super().__init__()
        self.what = what

    def extra_repr(self):
        return self.what + " (fused)"


class _BatchNormParams(nn.Module):
    """Parameters and running statistics of nn.BatchNorm3d(C, momentum=0.1) (joint_model.py:12-13), same state_dict keys
    (weight, bias, running_mean, running_var, num_batches_tracked); the arithmetic is fused into the neighbouring kernels
    (engine._bn_forward_tables + csrc/affine_act.cu).  .train() / .eval() select batch / running statistics."""

    def __init__(self, num_features, eps=1e-5, momentum=0.1):
        super().__init__()
        self.num_features, self.eps, self.momentum = num_features, eps, momentum
        self.weight = nn.Parameter(torch.ones(num_features))
        self.bias = nn.Parameter(torch.zeros(num_features))
        self.register_buffer("running_mean", torch.zeros(num_features))
        self.register_buffer("running_var", torch.ones(num_features))
        self.register_buffer("num_batches_tracked", torch.tensor(0, dtype=torch.long))

    def extra_repr(self):
        return "%d, eps=%g, momentum=%g, affine=True, track_running_stats=True (fused)" % (self.num_features, self.eps, self.momentum)


def Normalization(norm_type, out_channels, num_group=1):
    if norm_type == 1:
        return _Fused("InstanceNorm3d(%d, eps=1e-05, affine=False)" % out_channels)
    if norm_type == 2:
        return _BatchNormParams(out_channels)
    raise NotImplementedError("vaeseg_b200 implements norm_type=1 (InstanceNorm3d) and 2 (BatchNorm3d); got norm_type=%r "
                              "(GSNorm3d is not used by any shipped script)" % (norm_type,))


# ---- program building --------------------------------------------------------------------
def _index(module):
    names, params = [], []
    for k, p in module.named_parameters():
        names.append(k)
        params.append(p)
    return {k: i for i, k in enumerate(names)}, params


def _j(prefix, name):
    return prefix + "." + name if prefix else name


def _conv_layers(idx, prefix, cin, cout, in_planar=False, **kw):
    # the normalisation sits at the next nn.Sequential index (joint_model.py:40-46,106); a _BatchNormParams there
    # switches the layer to the BatchNorm path of the engine
    head, _, last = prefix.rpartition(".")
    norm_name = (head + "." if head else "") + str(int(last) + 1)
    bn = idx.get("__modules__", {}).get(norm_name)
    if isinstance(bn, _BatchNormParams):
        kw = dict(kw, bn=bn, gi=idx[norm_name + ".weight"], bti=idx[norm_name + ".bias"])
    return [Layer(C3IN, prefix, cin, cout, idx[prefix + ".weight"], idx[prefix + ".bias"], in_planar=in_planar, **kw)]


def _double_conv_layers(idx, prefix, cin, cout, in_planar=False, save_as=None, skip_from=None):
    ls = _conv_layers(idx, _j(prefix, "conv.0"), cin, cout, in_planar=in_planar)
    ls += _conv_layers(idx, _j(prefix, "conv.3"), cout, cout)
    ls += _conv_layers(idx, _j(prefix, "conv.6"), cout, cout, save_as=save_as, skip_from=skip_from)
    return ls


def _down_layers(idx, prefix, cin, cout, **kw):
    p = _j(prefix, "conv.0")
    return [Layer(K2DOWN, p, cin, cin, idx[p + ".weight"], idx[p + ".bias"])] + \
        _double_conv_layers(idx, _j(prefix, "conv.1"), cin, cout, **kw)


def _up_layers(idx, prefix, cin, cout, **kw):
    p = _j(prefix, "conv.0")
    return [Layer(K2UP, p, cin, cin, idx[p + ".weight"], idx[p + ".bias"])] + \
        _double_conv_layers(idx, _j(prefix, "conv.1"), cin, cout, **kw)


class _Engineered(nn.Module):
    """Mixin: precision selection + weight-pack cache + lazily built layer program."""

    def _engine_init(self):
        self._precision = default_precision()
        self._cache = engine.PackCache()
        self._program = None

    def set_precision(self, precision):
        if precision not in _PRECISIONS:
            raise ValueError("precision must be one of %s" % sorted(_PRECISIONS))
        for m in self.modules():
            if isinstance(m, _Engineered):
                m._precision = precision
        return self

    @property
    def compute_dtype(self):
        return _PRECISIONS[self._precision]

    def invalidate_packs(self):
        """Call after updating parameters through raw pointers (fused optimiser, EMA): the derived weight packs are
        re-made lazily, layer by layer, on next use."""
        for m in self.modules():
            if isinstance(m, _Engineered):
                m._cache.invalidate()

    def repack_packs(self):
        """Same contract as invalidate_packs(), but refreshes every pack now, in place, with one launch per
        network -- what the fused optimiser calls every step (and what keeps captured CUDA graphs valid)."""
        for m in self.modules():
            if isinstance(m, _Engineered):
                m._cache.repack_all()

    def _apply(self, fn, *a, **k):
        out = super()._apply(fn, *a, **k)
        self._cache.drop()
        self._program = None
        return out

    def _prog(self):
        if self._program is None:
            idx, params = _index(self)
            idx["__modules__"] = dict(self.named_modules())
            self._program = (self._build(idx), params)
        return self._program

    # standalone block execution: NCDHW fp32 in / out, converted with torch (not the hot path)
    def _run_block(self, x):
        layers, params = self._prog()
        dtype = self.compute_dtype
        xin = x.permute(0, 2, 3, 4, 1).contiguous().to(dtype)
        out = engine.ProgramFn.apply((layers, dtype, self._cache, params, torch.is_grad_enabled()), xin, *params)
        return out.permute(0, 4, 1, 2, 3).float()


def _check_input(x, what):
    if not torch.is_tensor(x) or not x.is_cuda:
        raise RuntimeError("vaeseg_b200.%s: input must be a CUDA tensor (no CPU fallback)" % what)
    if x.dtype != torch.float32:
        raise RuntimeError("vaeseg_b200.%s: input must be float32 NCDHW, got %s" % (what, x.dtype))


class DoubleConv(_Engineered):
    # joint_model.py:35-52 -- three Conv3d(3,p1)+IN+ReLU; params at Sequential indices 0,3,6
    def __init__(self, in_ch, out_ch, norm_type=2, soft=False):
        super().__init__()
        if soft:
            raise NotImplementedError("Softplus activation (soft=True) is not implemented; every reference block is built with soft=False")
        self.in_ch, self.out_ch = in_ch, out_ch
        self.conv = nn.Sequential(
            _ConvParams(in_ch, out_ch, 3), Normalization(norm_type, out_ch), _Fused("ReLU"),
            _ConvParams(out_ch, out_ch, 3), Normalization(norm_type, out_ch), _Fused("ReLU"),
            _ConvParams(out_ch, out_ch, 3), Normalization(norm_type, out_ch), _Fused("ReLU"))
        self._engine_init()

    def _build(self, idx):
        return _double_conv_layers(idx, "", self.in_ch, self.out_ch)

    def forward(self, x):
        _check_input(x, "DoubleConv")
        return self._run_block(x)


class Conv(_Engineered):
    # joint_model.py:101-112
    def __init__(self, in_ch, out_ch, norm_type=2, num_group=1, activation=True, norm=True, soft=False):
        super().__init__()
        if soft:
            raise NotImplementedError("Softplus activation (soft=True) is not implemented")
        self.in_ch, self.out_ch = in_ch, out_ch
        self.conv = nn.Sequential(_ConvParams(in_ch, out_ch, 3), Normalization(norm_type, out_ch), _Fused("ReLU"))
        self._engine_init()

    def _build(self, idx):
        return _conv_layers(idx, "conv.0", self.in_ch, self.out_ch)

    def forward(self, x):
        _check_input(x, "Conv")
        return self._run_block(x)


class Up(_Engineered):
    # joint_model.py:114-124
    def __init__(self, in_ch, out_ch, norm_type=2, kernal_size=(2, 2, 2), stride=(2, 2, 2), soft=False):
        super().__init__()
        if tuple(kernal_size) != (2, 2, 2) or tuple(stride) != (2, 2, 2):
            raise NotImplementedError("Up supports kernel 2, stride 2 only")
        self.in_ch, self.out_ch = in_ch, out_ch
        self.conv = nn.Sequential(_ConvParams(in_ch, in_ch, 2, transposed=True), DoubleConv(in_ch, out_ch, norm_type, soft=False))
        self._engine_init()

    def _build(self, idx):
        return _up_layers(idx, "", self.in_ch, self.out_ch)

    def forward(self, x):
        _check_input(x, "Up")
        return self._run_block(x)


class Down(_Engineered):
    # joint_model.py:126-136
    def __init__(self, in_ch, out_ch, norm_type=2, kernal_size=(2, 2, 2), stride=(2, 2, 2), soft=False):
        super().__init__()
        if tuple(kernal_size) != (2, 2, 2) or tuple(stride) != (2, 2, 2):
            raise NotImplementedError("Down supports kernel 2, stride 2 only")
        self.in_ch, self.out_ch = in_ch, out_ch
        self.conv = nn.Sequential(_ConvParams(in_ch, in_ch, 2), DoubleConv(in_ch, out_ch, norm_type, soft=False))
        self._engine_init()

    def _build(self, idx):
        return _down_layers(idx, "", self.in_ch, self.out_ch)

    def forward(self, x):
        _check_input(x, "Down")
        return self._run_block(x)


def _no_dropout(p, who):
    if p:
        raise NotImplementedError("vaeseg_b200.%s: dropout=%r is not implemented (every shipped preset uses 0; "
                                  "F.dropout draws from the device generator and is not reproducible against the "
                                  "reference anyway)" % (who, p))


class Segmentation(_Engineered):
    # joint_model.py:349-390
    def __init__(self, n_channels, n_class, norm_type=2, n_fmaps=[8, 16, 32, 64, 128, 256]):
        super().__init__()
        if n_class != 2:
            raise NotImplementedError("vaeseg_b200 implements the 2-class softmax head only (n_class=%r)" % (n_class,))
        f = list(n_fmaps)
        self.n_fmaps = f
        self.n_channels = n_channels
        self.in_block = Conv(n_channels, f[0], norm_type=norm_type, soft=False)
        self.down1 = Down(f[0], f[1], norm_type=norm_type, soft=False)
        self.down2 = Down(f[1], f[2], norm_type=norm_type, soft=False)
        self.down3 = Down(f[2], f[3], norm_type=norm_type, soft=False)
        self.down4 = Down(f[3], f[4], norm_type=norm_type, soft=False)
        self.up2 = Up(f[4], f[3], norm_type=norm_type, soft=False)
        self.up3 = Up(f[3], f[2], norm_type=norm_type, soft=False)
        self.up4 = Up(f[2], f[1], norm_type=norm_type, soft=False)
        self.up5 = Up(f[1], f[0], norm_type=norm_type, soft=False)
        self.out_block = _ConvParams(f[0], n_class, 3)
        self.final = _Fused("Softmax(dim=1)")
        self.n_class = n_class
        self._engine_init()

    def _build(self, idx):
        f = self.n_fmaps
        ls = _conv_layers(idx, "in_block.conv.0", self.n_channels, f[0], in_planar=True)
        ls += _down_layers(idx, "down1", f[0], f[1], save_as="x2")
        ls += _down_layers(idx, "down2", f[1], f[2], save_as="x3")
        ls += _down_layers(idx, "down3", f[2], f[3])
        ls += _down_layers(idx, "down4", f[3], f[4])
        ls += _up_layers(idx, "up2", f[4], f[3])
        ls += _up_layers(idx, "up3", f[3], f[2], skip_from="x3")     # up3(x) + x3, joint_model.py:381
        ls += _up_layers(idx, "up4", f[2], f[1], skip_from="x2")     # up4(x) + x2, joint_model.py:383
        ls += _up_layers(idx, "up5", f[1], f[0])
        ls += [Layer(HEAD, "out_block", f[0], self.n_class, idx["out_block.weight"], idx["out_block.bias"])]
        return ls

    def predict(self, x):
        """x: [B, n_channels, P, P, P] fp32 CUDA -> softmax probabilities [B, n_class, P, P, P]."""
        _check_input(x, "Segmentation")
        if x.shape[2] % 16 or x.shape[3] % 16 or x.shape[4] % 16:
            raise RuntimeError("Segmentation: spatial dims must be divisible by 16 (four stride-2 levels), got %s" % (tuple(x.shape[2:]),))
        layers, params = self._prog()
        return engine.ProgramFn.apply((layers, self.compute_dtype, self._cache, params, torch.is_grad_enabled()), x, *params)

    def forward(self, data_dict, in_key, out_key, dropout=0.0):
        _no_dropout(dropout, "Segmentation")
        data_dict[out_key] = self.predict(data_dict[in_key])
        return data_dict


class VAE(_Engineered):
    # joint_model.py:204-272
    def __init__(self, n_channels, n_class, norm_type=2, n_fmaps=[8, 16, 32, 64, 128, 256], dim=1024, soft=False,
                 patch=128):
        super().__init__()
        if n_class != 2:
            raise NotImplementedError("vaeseg_b200 implements the 2-class softmax head only (n_class=%r)" % (n_class,))
        if patch % 32:
            raise ValueError("VAE patch size must be divisible by 32 (five stride-2 levels)")
        f = list(n_fmaps)
        self.n_fmaps, self.dim, self.patch = f, dim, patch
        self.side = patch // 32
        flat = f[5] * self.side ** 3            # 16384 at the reference's 128^3
        self.in_block = Conv(n_class, f[0], norm_type=norm_type, soft=False)
        self.down1 = Down(f[0], f[1], norm_type=norm_type, soft=False)
        self.down2 = Down(f[1], f[2], norm_type=norm_type, soft=False)
        self.down3 = Down(f[2], f[3], norm_type=norm_type, soft=False)
        self.down4 = Down(f[3], f[4], norm_type=norm_type, soft=False)
        self.down5 = Down(f[4], f[5], norm_type=norm_type, soft=False)
        self.fc_mean = _LinearParams(flat, dim)
        self.fc_std = _LinearParams(flat, dim)
        self.fc2 = _LinearParams(dim, flat)
        self.up1 = Up(f[5], f[4], norm_type=norm_type, soft=False)
        self.up2 = Up(f[4], f[3], norm_type=norm_type, soft=False)
        self.up3 = Up(f[3], f[2], norm_type=norm_type, soft=False)
        self.up4 = Up(f[2], f[1], norm_type=norm_type, soft=False)
        self.up5 = Up(f[1], f[0], norm_type=norm_type, soft=False)
        self.out_block = _ConvParams(f[0], n_class, 3)
        self.final = _Fused("Softmax(dim=1)")
        self.n_class = n_class
        self._engine_init()

    def _build(self, idx):
        f = self.n_fmaps
        enc = _conv_layers(idx, "in_block.conv.0", self.n_class, f[0], in_planar=True)
        for i in range(1, 6):
            enc += _down_layers(idx, "down%d" % i, f[i - 1], f[i])
        dec = []
        for i in range(1, 6):
            dec += _up_layers(idx, "up%d" % i, f[6 - i], f[5 - i])
        dec += [Layer(HEAD, "out_block", f[0], self.n_class, idx["out_block.weight"], idx["out_block.bias"])]
        fc = tuple(idx[k] for k in ("fc_mean.weight", "fc_mean.bias", "fc_std.weight", "fc_std.bias",
                                    "fc2.weight", "fc2.bias"))
        return enc, dec, fc

    def forward(self, x, if_random=False, scale=1, mid_input=False, dropout=0.0, z=None):
        _no_dropout(dropout, "VAE")
        (enc, dec, fc), params = self._prog()
        dtype = self.compute_dtype
        if mid_input:
            if not x.is_cuda:
                raise RuntimeError("vaeseg_b200.VAE: input must be a CUDA tensor (no CPU fallback)")
            return engine.DecodeFn.apply((dec, fc, dtype, self._cache, params, self.dim, self.side, torch.is_grad_enabled()), x, *params)
        _check_input(x, "VAE")
        if x.shape[2] != self.patch or x.shape[3] != self.patch or x.shape[4] != self.patch:
            raise RuntimeError("VAE built for %d^3 patches got input %s" % (self.patch, tuple(x.shape)))
        if z is None:
            # drawn from the CPU default generator on every call, used or not (joint_model.py:246)
            z = torch.randn(x.size(0), self.dim)
        zdev = None
        if if_random:
            zdev = z if z.is_cuda else z.pin_memory().to(x.device, non_blocking=True)
            zdev = zdev.float().contiguous()
        return engine.VAEFn.apply((enc, dec, fc, dtype, self._cache, params, self.dim, torch.is_grad_enabled()), x, zdev, float(scale),
                                  bool(if_random), *params)


class Encoder(_Engineered):
    # joint_model.py:274-305 -- conv trunk (in_block + five Down) -> fc1 -> ReLU -> fc2 -> ReLU -> fc_mean -> sigmoid;
    # the discriminator of `Joint2` and the latent predictor of `Embed`.  `patch` generalises the hard-coded 16384
    # (= 256 * 4^3 at 128^3 patches) like VAE.patch.
    def __init__(self, n_channels, dim, norm_type=2, n_fmaps=[8, 16, 32, 64, 128, 256], soft=False, patch=128):
        super().__init__()
        if patch % 32:
            raise ValueError("Encoder patch size must be divisible by 32 (five stride-2 levels)")
        f = list(n_fmaps)
        self.n_fmaps, self.n_channels, self.patch, self.side = f, n_channels, patch, patch // 32
        flat = f[5] * self.side ** 3
        self.in_block = Conv(n_channels, f[0], norm_type=norm_type, soft=False)
        self.down1 = Down(f[0], f[1], norm_type=norm_type, soft=False)
        self.down2 = Down(f[1], f[2], norm_type=norm_type, soft=False)
        self.down3 = Down(f[2], f[3], norm_type=norm_type, soft=False)
        self.down4 = Down(f[3], f[4], norm_type=norm_type, soft=False)
        self.down5 = Down(f[4], f[5], norm_type=norm_type, soft=False)
        self.fc1 = _LinearParams(flat, 1024)
        self.fc2 = _LinearParams(1024, 128)
        self.fc_mean = _LinearParams(128, dim)
        self._engine_init()

    def _build(self, idx):
        f = self.n_fmaps
        ls = _conv_layers(idx, "in_block.conv.0", self.n_channels, f[0], in_planar=True)
        for i in range(1, 6):
            ls += _down_layers(idx, "down%d" % i, f[i - 1], f[i])
        return ls

    def forward(self, x):
        _check_input(x, "Encoder")
        if self.n_channels > 2:
            raise NotImplementedError("Encoder: the planar in-block kernel takes 1 or 2 input channels")
        layers, params = self._prog()
        h = engine.ProgramFn.apply((layers, self.compute_dtype, self._cache, params, torch.is_grad_enabled()), x.contiguous(), *params)
        # internal NDHWC -> the reference's NCDHW flatten order (x.view(B, 16384)); a [B, 256 * side^3] tensor: not hot
        h = h.permute(0, 4, 1, 2, 3).reshape(h.shape[0], -1).float()
        h = engine.LinearFn.apply(h, self.fc1.weight, self.fc1.bias, engine.ops.ACT_RELU)
        h = engine.LinearFn.apply(h, self.fc2.weight, self.fc2.bias, engine.ops.ACT_RELU)
        return engine.LinearFn.apply(h, self.fc_mean.weight, self.fc_mean.bias, engine.ops.ACT_SIGMOID)


class Fusion(_Engineered):
    # joint_model.py:392-436 -- two in-block + Down branches (image, mask) summed, a merge Conv, then the Segmentation
    # trunk from down2 on with the same two additive skips
    def __init__(self, n_channels_img, n_channels_mask, n_class, norm_type=2, n_fmaps=[8, 16, 32, 64, 128, 256]):
        super().__init__()
        if n_class != 2:
            raise NotImplementedError("vaeseg_b200 implements the 2-class softmax head only (n_class=%r)" % (n_class,))
        f = list(n_fmaps)
        self.n_fmaps, self.n_class = f, n_class
        self.n_channels_img, self.n_channels_mask = n_channels_img, n_channels_mask
        self.in_block = Conv(n_channels_img, f[0], norm_type=norm_type, soft=False)
        self.down1 = Down(f[0], f[1], norm_type=norm_type, soft=False)
        self.in_block_mask = Conv(n_channels_mask, f[0], norm_type=norm_type, soft=False)
        self.down1_mask = Down(f[0], f[1], norm_type=norm_type, soft=False)
        self.merge = Conv(f[1], f[1], norm_type=norm_type, soft=False)
        self.down2 = Down(f[1], f[2], norm_type=norm_type, soft=False)
        self.down3 = Down(f[2], f[3], norm_type=norm_type, soft=False)
        self.down4 = Down(f[3], f[4], norm_type=norm_type, soft=False)
        self.up2 = Up(f[4], f[3], norm_type=norm_type, soft=False)
        self.up3 = Up(f[3], f[2], norm_type=norm_type, soft=False)
        self.up4 = Up(f[2], f[1], norm_type=norm_type, soft=False)
        self.up5 = Up(f[1], f[0], norm_type=norm_type, soft=False)
        self.out_block = _ConvParams(f[0], n_class, 3)
        self.final = _Fused("Softmax(dim=1)")
        self._engine_init()

    def _build(self, idx):
        f = self.n_fmaps
        img = _conv_layers(idx, "in_block.conv.0", self.n_channels_img, f[0], in_planar=True) + _down_layers(idx, "down1", f[0], f[1])
        mask = _conv_layers(idx, "in_block_mask.conv.0", self.n_channels_mask, f[0], in_planar=True) + \
            _down_layers(idx, "down1_mask", f[0], f[1])
        trunk = _conv_layers(idx, "merge.conv.0", f[1], f[1], save_as="x2")
        trunk += _down_layers(idx, "down2", f[1], f[2], save_as="x3")
        trunk += _down_layers(idx, "down3", f[2], f[3])
        trunk += _down_layers(idx, "down4", f[3], f[4])
        trunk += _up_layers(idx, "up2", f[4], f[3])
        trunk += _up_layers(idx, "up3", f[3], f[2], skip_from="x3")
        trunk += _up_layers(idx, "up4", f[2], f[1], skip_from="x2")
        trunk += _up_layers(idx, "up5", f[1], f[0])
        trunk += [Layer(HEAD, "out_block", f[0], self.n_class, idx["out_block.weight"], idx["out_block.bias"])]
        return img, mask, trunk

    def forward(self, data_dict, in_key_img, in_key_mask, out_key):
        x_img, x_mask = data_dict[in_key_img], data_dict[in_key_mask]
        _check_input(x_img, "Fusion")
        _check_input(x_mask, "Fusion")
        (img, mask, trunk), params = self._prog()
        spec = lambda ls: (ls, self.compute_dtype, self._cache, params, torch.is_grad_enabled())
        a = engine.ProgramFn.apply(spec(img), x_img.contiguous(), *params)
        b = engine.ProgramFn.apply(spec(mask), x_mask.contiguous(), *params)
        data_dict[out_key] = engine.ProgramFn.apply(spec(trunk), a + b, *params)      # x2_img + x2_mask, joint_model.py:424
        return data_dict


class Joint2(_Engineered):
    # joint_model.py:454-465 -- segmentation + discriminator on the foreground probability
    def __init__(self, models, seg_dropout=0.0):
        super().__init__()
        self.Seg = models[0]
        self.Dis = models[1]
        self.seg_dropout = seg_dropout
        self._engine_init()

    def forward(self, data_dict, in_key, out_key, score_key, dropout=False):
        if dropout:
            data_dict = self.Seg(data_dict, in_key, out_key, dropout=self.seg_dropout)
        else:
            data_dict = self.Seg(data_dict, in_key, out_key)
        data_dict[score_key] = self.Dis(data_dict[out_key][:, 1:2, :, :, :].contiguous())
        return data_dict


class Embed(_Engineered):
    # joint_model.py:468-501
    def __init__(self, models):
        super().__init__()
        self.Encoder = models[0]
        self.Vae = models[1]
        self.Fusion = models[2]
        self._engine_init()

    def forward(self, data_dict, in_key, out_key, test_mode=False, loop_input=None, seg_input=None, latent_input=None):
        if latent_input:
            data_dict["latent_code"] = data_dict[latent_input]
        else:
            data_dict["latent_code"] = self.Encoder(data_dict[in_key])
        data_dict["gt_recon"], data_dict["latent_code_gt"], data_dict["latent_code_std"] = self.Vae(
            data_dict["venous_pancreas_only"], if_random=True, scale=0.5, mid_input=False)
        if loop_input:
            data_dict[loop_input], data_dict["latent_code_loop"], _ = self.Vae(data_dict[loop_input], if_random=False, scale=0,
                                                                               mid_input=False)
        if seg_input:
            data_dict["init_seg"] = data_dict[seg_input]
        else:
            data_dict["init_seg"] = self.Vae(data_dict["latent_code"], if_random=False, scale=0, mid_input=True)
        if loop_input:
            data_dict = self.Fusion(data_dict, in_key, loop_input, out_key)
        elif test_mode:
            data_dict = self.Fusion(data_dict, in_key, "init_seg", out_key)
        else:
            data_dict = self.Fusion(data_dict, in_key, "gt_recon", out_key)
        data_dict["seg_recon"], _, _ = self.Vae(data_dict["init_seg"].detach(), if_random=False, scale=0, mid_input=False)
        return data_dict


class Joint(_Engineered):
    # joint_model.py:438-452
    def __init__(self, models, vae_forward_scale=0.0, vae_decoder_dropout=0.0, seg_dropout=0.0):
        super().__init__()
        self.Seg = models[0]
        self.Vae = models[1]
        self.vae_forward_scale = vae_forward_scale
        self.vae_decoder_dropout = vae_decoder_dropout
        self.seg_dropout = seg_dropout
        self._engine_init()

    def forward(self, data_dict, in_key, out_key, out_key_recon, dropout=False):
        if dropout:
            data_dict = self.Seg(data_dict, in_key, out_key, dropout=self.seg_dropout)
        else:
            data_dict = self.Seg(data_dict, in_key, out_key)
        if dropout:
            # the student's mean/std are discarded in this branch
            data_dict[out_key_recon], _, _ = self.Vae(data_dict[out_key], if_random=False,
                                                      scale=self.vae_forward_scale, dropout=self.vae_decoder_dropout)
        else:
            data_dict[out_key_recon], data_dict["mean"], data_dict["std"] = self.Vae(
                data_dict[out_key], if_random=False, scale=self.vae_forward_scale)
        return data_dict
