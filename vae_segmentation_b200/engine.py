"""Layer-program executor behind the drop-in modules.

A network (or a single block) is described as a flat list of `Layer`s; `program_forward`
launches the C-ABI kernels layer by layer and records a tape, `program_backward` replays the
tape in reverse with hand-written dgrad / wgrad / norm-backward kernels.  torch.autograd only
sees one node per network (`ProgramFn`, `VAEFn`), so the backward schedule, buffer reuse and
gradient accumulation are ours.

Internal activation layout is NDHWC in the compute dtype (bf16, or fp32 in check mode); the
nn.Module boundary stays NCDHW fp32 ("planar"): first layers read planar input directly and
the softmax head writes planar probabilities, so no layout-conversion kernels run.

Reference semantics implemented here: joint_model.py:35-52,101-136 (blocks), :369-390
(Segmentation.forward incl. the two additive skips), :227-272 (VAE.forward).
"""
import ctypes
import os

import torch

from . import ops

C3IN, K2DOWN, K2UP, HEAD = "c3in", "k2down", "k2up", "head"

# bf16 mode routes eligible 3x3x3 convolutions (fprop + dgrad) to the tcgen05/TMEM/TMA kernel; the CUDA-core
# direct kernel remains for the fp32 check mode and the Cin in {1,2} / Cout = 2 layers.  VAESEG_NO_TC=1 forces
# the direct kernel everywhere (A/B measurements).
USE_TENSOR_CORES = os.environ.get("VAESEG_NO_TC", "0") != "1"
# VAESEG_NO_FUSE_REDUCE=1 keeps the InstanceNorm-backward reduction a separate pass (A/B measurements, parity tests)
FUSE_BWD_REDUCE = os.environ.get("VAESEG_NO_FUSE_REDUCE", "0") != "1"
FUSE_BWD_REDUCE_MAX_VOX = int(os.environ.get("VAESEG_FUSE_REDUCE_MAX_VOX", "0"))      # 0 = no limit (voxels per sample)
HEAD_DIRECT = os.environ.get("VAESEG_HEAD_DIRECT", "0") == "1"
# The forward 3x3x3 convolutions with 8 / 16 output channels at >= 48^3 run through the kd-in-N kernel
# (csrc/conv3_tc_kdn.cu: the three kd taps folded into the MMA's N, halving the A-operand shared-memory traffic).
# Measured on B200, joint step 2 x 96^3: 345.3 vs 327.2 vol/s.  VAESEG_KDN=0 falls back to the tap-per-MMA kernel.
USE_KDN = os.environ.get("VAESEG_KDN", "1") == "1"
# Full-resolution 2-channel in-block (the VAE's: planar fp32 probabilities in): convert the input once to 8 zero-padded
# bf16 channels and run the kd-in-N kernel instead of the CUDA-core in-block kernel (91 us -> 14 + 40 us at 2 x 96^3); the
# converted input is reused by the weight gradient.  Not for the 1-channel image in-block of Segmentation: rounding the
# IMAGE to bf16 moved the conditioned-fixture gradients of the 12^3 level from 4.x e-2 to 5.05e-2 of the floored norm
# (north-star bound 5e-2) for 11 us.  VAESEG_INBLOCK_TC=0 keeps the CUDA-core kernel everywhere.
INBLOCK_TC = os.environ.get("VAESEG_INBLOCK_TC", "1") == "1"
# Conv -> InstanceNorm -> ReLU (+ skip) of the tensor-core layers as ONE cooperative launch with an in-kernel grid barrier
# (ops.conv3_in_relu) instead of conv + a separate apply pass.  Correct (tests/test_conv3_tc_gpu.py) and measured SLOWER:
# joint step 316 vs 385 vol/s, seg 826 vs 864, vae 971 vs 1006 (profiles/r2_fused_apply_ab.txt).  A cooperative grid is
# gang-scheduled -- it starts only when ALL of its CTAs fit at once, so it cannot slide into the SMs a kernel of another
# stream (teacher forward, weight gradients) frees one by one, and the in-kernel pass (fence + grid barrier + L2 re-read
# on one CTA per SM) is no faster than the stand-alone apply kernel at full occupancy (16 vs 14 us at 2 x 96^3).  Opt-in.
FUSE_APPLY = os.environ.get("VAESEG_FUSE_APPLY", "0") == "1"
FUSE_APPLY_MAX_VOX = int(os.environ.get("VAESEG_FUSE_APPLY_MAX_VOX", "0"))        # 0 = every size (voxels per sample)
# 2x2x2 stride-2 convolutions / transposed convolutions (forward and input gradient) on the tensor cores
# (csrc/k2s2_tc.cu); VAESEG_K2_TC=0 keeps the CUDA-core kernels of csrc/k2s2.cu (A/B measurements).
USE_K2_TC = os.environ.get("VAESEG_K2_TC", "1") == "1"

# tools/precision_attrib.py and the calibration of tests/test_models_gpu.py only: names of tensor classes to round through bf16 while running the fp32
# check mode ("y", "a", "g", "dy", "k2"), to attribute bf16-mode error to a storage point.  Empty in production.
SIMULATE_BF16 = set()


def _sim(t, what):
    if what in SIMULATE_BF16 and t is not None and t.dtype == torch.float32:
        return t.bfloat16().float()
    return t


class Layer(object):
    __slots__ = ("kind", "name", "cin", "cout", "wi", "bi", "save_as", "skip_from", "in_planar", "bn", "gi", "bti")

    def __init__(self, kind, name, cin, cout, wi, bi, save_as=None, skip_from=None, in_planar=False, bn=None, gi=None, bti=None):
        self.kind, self.name, self.cin, self.cout = kind, name, cin, cout
        self.wi, self.bi = wi, bi
        self.save_as, self.skip_from, self.in_planar = save_as, skip_from, in_planar
        # BatchNorm3d instead of InstanceNorm3d behind the conv (norm_type 2): the parameter-holding module (running
        # statistics, momentum, eps, .training) and the indices of its weight (gamma) / bias (beta) in `tensors`
        self.bn, self.gi, self.bti = bn, gi, bti


class _PackJob(ctypes.Structure):
    # mirrors vs_pack_job (include/vaeseg_b200.h)
    _fields_ = [("w", ctypes.c_void_p), ("wf", ctypes.c_void_p), ("wd", ctypes.c_void_p), ("tcf", ctypes.c_void_p),
                ("tcd", ctypes.c_void_p), ("tcf_elems", ctypes.c_longlong), ("tcd_elems", ctypes.c_longlong),
                ("cin", ctypes.c_int), ("cout", ctypes.c_int), ("cout_pad", ctypes.c_int), ("cin_pad", ctypes.c_int),
                ("kdn", ctypes.c_void_p), ("kdn_elems", ctypes.c_longlong), ("kind", ctypes.c_int), ("kdn_dgrad", ctypes.c_int)]


class PackCache(object):
    """Derived caches of the fp32 master weights: fp32 [27][Cin][Cout] fprop / dgrad packs and the bf16 tensor-core
    (UMMA B operand) packs.  Buffers are allocated once per weight and re-packed IN PLACE, so CUDA graphs that captured
    their addresses stay valid.  An entry is current while (storage pointer, tensor version, epoch) match; the fused
    optimiser / EMA update (raw-pointer writes that torch's version counters do not see) call `repack_all()`, which
    refreshes every known entry with ONE batched launch, or `invalidate()` (lazy per-layer re-pack on next use)."""

    def __init__(self):
        self.epoch = 0
        self._store = {}          # (kind, data_ptr) -> entry dict
        self._jobs = None         # (device table, n) of the batched re-pack, rebuilt when the entry set changes

    def invalidate(self):
        self.epoch += 1

    def drop(self):
        """Forget everything (the parameters moved: .to(device), load into new storage)."""
        self.epoch += 1
        self._store.clear()
        self._jobs = None

    def _entry(self, kind, w, tc):
        key = (w.data_ptr(), w._version, self.epoch)
        ent = self._store.get((kind, w.data_ptr()))
        fresh = ent is None or ent["shape"] != tuple(w.shape) or (tc and not ent["tc"])
        if fresh:
            ent = {"w": w, "shape": tuple(w.shape), "tc": False, "key": None, "wf": None, "wd": None, "tcf": None,
                   "tcd": None, "kdn": None, "kind": kind}
            self._store[(kind, w.data_ptr())] = ent
            self._jobs = None
        return ent, key

    def conv3(self, w, tc=False):
        """Returns (wf, wd, wtc_fprop, wtc_dgrad); the tensor-core packs are None unless tc=True (bf16
        mode) and the tcgen05 path takes the shape."""
        tc = bool(tc and USE_TENSOR_CORES)
        ent, key = self._entry("conv3", w, tc)
        stale_fp32 = ent["wf"] is None and not (tc and ent["tcf"] is not None and ent["tcd"] is not None)
        if ent["key"] != key or (tc and not ent["tc"]) or stale_fp32:
            wdet = w.detach()
            cout, cin = wdet.shape[0], wdet.shape[1]
            if tc:
                ent["tcf"] = ops.pack_conv3_weight_tc(wdet, dgrad=False, out=ent["tcf"])
                ent["tcd"] = ops.pack_conv3_weight_tc(wdet, dgrad=True, out=ent["tcd"])
                if not ent["tc"]:
                    self._jobs = None
                ent["tc"] = True
            # the fp32 [27][Cin][Cout] packs feed the CUDA-core direct kernels only: skipped for layers whose fprop AND
            # dgrad both run on the tensor cores (most of the re-pack traffic after every optimiser step)
            if not (tc and ent["tcf"] is not None and ent["tcd"] is not None):
                if ent["wf"] is None:
                    ent["wf"] = torch.empty(27, cin, cout, device=wdet.device, dtype=torch.float32)
                    ent["wd"] = torch.empty(27, cout, cin, device=wdet.device, dtype=torch.float32)
                    self._jobs = None
                ops.pack_conv3_weight(wdet, out=(ent["wf"], ent["wd"]))
            ent["key"] = key
        return ent["wf"], ent["wd"], ent["tcf"], ent["tcd"]

    def head_tc(self, w):
        """bf16 tensor-core (fprop, dgrad) packs of the head weight [n_class,Cin,3,3,3] zero-padded to 8 output
        channels: the head runs as an 8-channel layer whose epilogue stores softmax probabilities (forward) and whose
        logit gradient is produced as an 8-channel tensor (ops.softmax2_bwd_pad8).  (None, None) without tcgen05."""
        if not USE_TENSOR_CORES:
            return None, None
        ent, key = self._entry("head8", w, True)
        if ent["key"] != key:
            wdet = w.detach()
            cin = wdet.shape[1]
            ent["tcf"] = ops.pack_conv3_weight_tc_padded(wdet, cin, 8, dgrad=False, out=ent["tcf"])
            ent["tcd"] = ops.pack_conv3_weight_tc_padded(wdet, cin, 8, dgrad=True, out=ent["tcd"])
            if not ent["tc"]:
                self._jobs = None
            ent["tc"] = True
            ent["key"] = key
        return ent["tcf"], ent["tcd"]

    def inblock2_dgrad_tc(self, w):
        """bf16 tensor-core dgrad pack of a 2-input-channel in-block weight [Cout,2,3,3,3] with the input channels
        zero-padded to 8 (the planar 2-channel input gradient of the VAE, cabi vs_conv3x3x3_dgrad), or None."""
        if not USE_TENSOR_CORES:
            return None
        ent, key = self._entry("inblk8", w, True)
        if ent["key"] != key:
            wdet = w.detach()
            ent["tcd"] = ops.pack_conv3_weight_tc_padded(wdet, 8, wdet.shape[0], dgrad=True, out=ent["tcd"])
            if not ent["tc"]:
                self._jobs = None
            ent["tc"] = True
            ent["key"] = key
        return ent["tcd"]

    def _kdn_padded(self, kind, w, cin_pad, cout_pad, dgrad):
        ent, key = self._entry(kind, w, True)
        if ent["key"] != key:
            had = ent["kdn"] is not None
            ent["kdn"] = ops.pack_conv3_weight_tc_kdn_padded(w.detach(), cin_pad, cout_pad, dgrad=dgrad, out=ent["kdn"])
            if not had:
                self._jobs = None
            ent["tc"] = True
            ent["key"] = key
        return ent["kdn"]

    def head_kdn(self, w):
        """kd-in-N (fprop, dgrad) packs of the head weight [n_class,Cin,3,3,3] zero-padded to 8 output channels (the
        full-resolution head through the kd-in-N kernel), entries None where that kernel does not take the shape."""
        if not (USE_TENSOR_CORES and USE_KDN):
            return None, None
        cin = w.shape[1]
        return (self._kdn_padded("head8kdn", w, cin, 8, False), self._kdn_padded("head8kdnd", w, cin, 8, True))

    def inblock_kdn(self, w):
        """kd-in-N fprop pack of an in-block weight [Cout,Cin<=2,3,3,3] with the input channels zero-padded to 8, or None."""
        if not (USE_TENSOR_CORES and USE_KDN):
            return None
        return self._kdn_padded("inblk8kdn", w, 8, w.shape[0], False)

    def inblock2_dgrad_kdn(self, w):
        """kd-in-N dgrad pack of a 2-input-channel in-block weight with the input channels zero-padded to 8, or None."""
        if not (USE_TENSOR_CORES and USE_KDN):
            return None
        return self._kdn_padded("inblk8kdnd", w, 8, w.shape[0], True)

    def conv3_kdn(self, w, dgrad=False):
        """kd-in-N fprop (or dgrad) pack of a 3x3x3 weight (re-packed in place, part of the batched re-pack), or None."""
        ent, key = self._entry("kdnd" if dgrad else "kdn", w, True)
        if ent["key"] != key:
            had = ent["kdn"] is not None
            ent["kdn"] = ops.pack_conv3_weight_tc_kdn(w.detach(), dgrad=dgrad, out=ent["kdn"])
            if not had:
                self._jobs = None
            ent["tc"] = True
            ent["key"] = key
        return ent["kdn"]

    def k2s2(self, w, a, b):
        """(gather pack, scatter pack) of a 2x2x2 stride-2 weight viewed as wt[A][B][8] for the tcgen05 kernels, or
        (None, None) when the tensor-core path does not take the channel counts."""
        if not USE_TENSOR_CORES or not USE_K2_TC:
            return None, None
        ent, key = self._entry("k2s2", w, True)
        if ent["key"] != key:
            had = ent["tcf"] is not None
            wdet = w.detach()
            ent["tcf"] = ops.pack_k2s2_weight_tc(wdet, a, b, scatter=False, out=ent["tcf"])
            ent["tcd"] = ops.pack_k2s2_weight_tc(wdet, a, b, scatter=True, out=ent["tcd"])
            if not had:
                self._jobs = None
            ent["tc"] = True
            ent["key"] = key
        return ent["tcf"], ent["tcd"]

    def repack_all(self):
        """Re-packs every known entry in place with one launch; entries become current for the present weights."""
        ents = [e for e in self._store.values() if e["key"] is not None]
        if not ents:
            self.invalidate()
            return
        if self._jobs is None or self._jobs[1] != len(ents):
            arr = (_PackJob * len(ents))()
            for j, e in zip(arr, ents):
                w = e["w"].detach()
                j.w = w.data_ptr()
                j.cout, j.cin = w.shape[0], w.shape[1]
                j.cout_pad = 8 if e["kind"] in ("head8", "head8kdn", "head8kdnd") else w.shape[0]
                j.cin_pad = 8 if e["kind"] in ("inblk8", "inblk8kdn", "inblk8kdnd") else 0
                j.wf = e["wf"].data_ptr() if e["wf"] is not None else None
                j.wd = e["wd"].data_ptr() if e["wd"] is not None else None
                j.tcf = e["tcf"].data_ptr() if e["tcf"] is not None else None
                j.tcd = e["tcd"].data_ptr() if e["tcd"] is not None else None
                j.tcf_elems = e["tcf"].numel() if e["tcf"] is not None else 0
                j.tcd_elems = e["tcd"].numel() if e["tcd"] is not None else 0
                j.kdn = e["kdn"].data_ptr() if e["kdn"] is not None else None
                j.kdn_elems = e["kdn"].numel() if e["kdn"] is not None else 0
                j.kind = 1 if e["kind"] == "k2s2" else 0         # k2s2: w = wt[A = shape[0]][B = shape[1]][8]
                j.kdn_dgrad = 1 if e["kind"] in ("kdnd", "head8kdnd", "inblk8kdnd") else 0
            host = torch.frombuffer(bytearray(bytes(arr)), dtype=torch.uint8)
            self._jobs = (host.to(ents[0]["w"].device), len(ents), ents)
        ops.pack_conv3_batched(self._jobs[0], self._jobs[1])
        self.epoch += 1
        for e in self._jobs[2]:
            e["key"] = (e["w"].data_ptr(), e["w"]._version, self.epoch)


# Optional side stream for the weight-gradient kernels (set by the trainers): wgrad(L) only needs dy(L) and the saved
# input of L, while the backward chain continues with dgrad(L) -> norm-backward(L-1) -> ...; on a second stream (a
# parallel branch of the captured CUDA graph) the wgrads fill the SMs that the small deep-level kernels leave idle.
# Only used when the gradient is accumulated in place into an existing .grad (the trainers' flat arenas): the
# consumer (optimiser) then waits on the side stream once per step (`join_wgrad_stream`).
WGRAD_STREAM = None            # a stream, or a list of streams used round-robin (independent layers overlap)
_wgrad_rr = [0]


def _wgrad_streams():
    ws = WGRAD_STREAM
    if ws is None:
        return []
    return list(ws) if isinstance(ws, (list, tuple)) else [ws]


def _wgrad_async(fn, inplace, *inputs):
    streams = _wgrad_streams()
    if not streams or not inplace:
        return fn()
    ws = streams[_wgrad_rr[0] % len(streams)]
    _wgrad_rr[0] += 1
    cur = torch.cuda.current_stream()
    ws.wait_stream(cur)
    with torch.cuda.stream(ws):
        out = fn()
    for t in inputs:
        if t is not None:
            t.record_stream(ws)
    return out


def join_wgrad_stream():
    for ws in _wgrad_streams():
        torch.cuda.current_stream().wait_stream(ws)
    _wgrad_rr[0] = 0


# Optional callable(param) invoked right after the kernels that produce a parameter's gradient have been ENQUEUED (on
# the current stream or a weight-gradient side stream): the data-parallel trainers use it to start the all-reduce of a
# gradient bucket while the rest of the backward pass is still running (train_step.BucketedAllReduce).
GRAD_READY_HOOK = None


def _grad_ready(*params):
    hook = GRAD_READY_HOOK
    if hook is not None:
        for p in params:
            if p is not None:
                hook(p)


def _tc_channels(c):
    return c == 8 or (c >= 16 and c % 16 == 0)


def _grad_target(p_ref, need):
    """Where a parameter gradient goes: into an existing .grad (accumulate in place, return
    None to autograd) or into a fresh tensor returned to autograd."""
    if not need:
        return None, False
    g = p_ref.grad if p_ref is not None else None
    if g is not None and g.is_cuda and g.dtype == torch.float32 and g.is_contiguous():
        return g, True
    return None, False


def _pair_targets(param_refs, need, wi, bi, wshape, bshape, device):
    """Resolves where the (weight, bias) gradients of one layer go.  Returns (dw, db, acc):
    buffers handed to the kernel (db None when the bias needs no gradient) and whether the
    kernel accumulates into them (existing .grad) or overwrites fresh tensors."""
    tw, aw = _grad_target(param_refs[wi], need[wi])
    tb, ab = _grad_target(param_refs[bi], need[bi])
    if need[wi] and need[bi] and aw != ab:
        raise RuntimeError("weight and bias of one layer must both have (or both lack) a .grad buffer")
    acc = aw if need[wi] else ab
    if tw is None:                       # fresh gradient, or scratch when only the bias trains
        tw = torch.empty(wshape, device=device, dtype=torch.float32)
    if need[bi] and tb is None:
        tb = torch.empty(bshape, device=device, dtype=torch.float32)
    return tw, tb, acc


def _hand_back(grads, need, wi, bi, tw, tb, acc):
    grads[wi] = tw if (need[wi] and not acc) else None
    grads[bi] = tb if (need[bi] and not acc) else None


def _bn_forward_tables(L, tensors, stats, shift, S):
    """BatchNorm3d (joint_model.py:12-13: momentum 0.1, eps 1e-5, affine, running statistics) on top of the per-(n,c)
    outputs of the convolution kernel: stats [N,C,2] = (sum, sum of squares) of the STORED output y_s = conv - shift[n,c]
    (no bias: it cancels in training mode and is folded in below in eval mode).  Returns the coefficient tables of the
    affine kernels -- kb: a = relu(k*y_s + b); k2b2: xhat = k2*y_s + b2 -- and (k [C], rstd [C]) in fp64.  A few hundred
    numbers: plain torch ops, fp64."""
    bn = L.bn
    n, c = stats.shape[0], stats.shape[1]
    gamma = tensors[L.gi].detach().double()
    beta = tensors[L.bti].detach().double()
    bias = tensors[L.bi].detach().double()
    sh = shift.double()
    if bn.training:
        m_n = sh + stats[..., 0] / S                                  # per-sample mean of the bias-free conv output
        v_n = stats[..., 1] / S - (stats[..., 0] / S) ** 2            # per-sample biased variance
        mean = m_n.mean(0)
        var = ((v_n + m_n ** 2).mean(0) - mean ** 2).clamp_min(0.0)   # biased variance over (N, D, H, W)
        with torch.no_grad():
            cnt = float(n * S)
            mom = bn.momentum
            bn.running_mean.mul_(1.0 - mom).add_((mean + bias).float(), alpha=mom)        # BN sees conv + bias
            bn.running_var.mul_(1.0 - mom).add_((var * (cnt / max(cnt - 1.0, 1.0))).float(), alpha=mom)
            bn.num_batches_tracked.add_(1)
    else:
        mean = bn.running_mean.double() - bias
        var = bn.running_var.double()
    rstd = torch.rsqrt(var + bn.eps)
    k = gamma * rstd
    b = beta[None] - k[None] * (mean[None] - sh)
    kb = torch.stack([k[None].expand(n, c), b], -1).float().contiguous()
    k2b2 = torch.stack([rstd[None].expand(n, c), (sh - mean[None]) * rstd[None]], -1).float().contiguous()
    return kb, k2b2, (k, rstd)


def _bn_backward_tables(L, sums, kb, k2b2, aux, S, training):
    """From the per-(n,c) sums (sum gm, sum gm*xhat): dgamma, dbeta, the conv-bias gradient and the coefficients of
    dy = c0*gm + c1 + c2*y_s (BatchNorm backward in training mode; a plain scale in eval mode)."""
    k, rstd = aux
    n, c = sums.shape[0], sums.shape[1]
    g0, g1 = sums[..., 0].sum(0), sums[..., 1].sum(0)                 # [C]
    dgamma, dbeta = g1.float(), g0.float()
    k2, b2 = k2b2[..., 0].double(), k2b2[..., 1].double()
    if training:
        cnt = float(n * S)
        m0, m1 = g0 / cnt, g1 / cnt
        c0 = k[None].expand(n, c)
        c1 = -k[None] * m0[None] - k[None] * m1[None] * b2
        c2 = -k[None] * m1[None] * k2
        dbias = torch.zeros(c, device=sums.device, dtype=torch.float32)   # cancels in the batch statistics
    else:
        c0 = k[None].expand(n, c)
        c1 = torch.zeros(n, c, device=sums.device, dtype=torch.float64)
        c2 = c1
        dbias = (k * g0).float()
    coef = torch.stack([c0, c1, c2], -1).float().contiguous()
    return coef, dgamma, dbeta, dbias


def _accumulate_small(param_refs, need, grads, i, value):
    """Hands a small parameter gradient (BatchNorm gamma / beta, conv bias) to its .grad buffer or back to autograd."""
    if i is None or not need[i]:
        return
    tgt, acc = _grad_target(param_refs[i], True)
    if acc:
        tgt.add_(value.reshape(tgt.shape))
        grads[i] = None
    else:
        grads[i] = value.reshape(param_refs[i].shape).clone()


def program_forward(layers, tensors, x, dims, dtype, cache, record=True):
    """x: NDHWC `dtype` tensor (planar fp32 when layers[0].in_planar).  dims=(N,D,H,W) of x.
    Returns (out, out_dims, tape)."""
    n, d, h, w = dims
    tape = []
    slots = {}
    cur = x
    # statistics + shift words of every Conv+InstanceNorm layer, zeroed by ONE launch (not one memset per layer)
    arena = ops.StatsArena(sum(ops.stats_words(n, L.cout) for L in layers if L.kind == C3IN), x.device)
    for L in layers:
        if L.kind == C3IN and L.bn is not None:
            # Conv3d -> BatchNorm3d -> ReLU (norm_type 2): same convolution kernels (tap-per-MMA / CUDA-core), batch
            # statistics pooled on the host side of the launch stream, affine + ReLU kernels of affine_act.cu
            wf, wd, wtc, wdtc = cache.conv3(tensors[L.wi], tc=(dtype == torch.bfloat16))
            y, stats, shift = ops.conv3_fprop(cur, wf, None, (n, d, h, w), L.cin, L.cout, dtype, in_planar=L.in_planar,
                                              wtc=None if L.in_planar else wtc, return_shift=True)
            skip = slots[L.skip_from] if L.skip_from is not None else None
            kb, k2b2, aux = _bn_forward_tables(L, tensors, stats, shift, float(d * h * w))
            a = ops.affine_relu_apply(y, kb, skip)
            if record:
                tape.append((L, cur, y, (kb, k2b2, aux, bool(L.bn.training)), (n, d, h, w), (wd, wdtc)))
            cur = a
        elif L.kind == C3IN:
            wf, wd, wtc, wdtc = cache.conv3(tensors[L.wi], tc=(dtype == torch.bfloat16))
            wkd_in = None
            if L.in_planar and L.cin == 2 and record and dtype == torch.bfloat16 and L.cout % 8 == 0:
                wdtc = cache.inblock2_dgrad_tc(tensors[L.wi])      # planar 2-channel input gradient on the tensor cores
                if L.cout == 8 and d >= 4 and d * h * w >= 48 ** 3:
                    wkd_in = cache.inblock2_dgrad_kdn(tensors[L.wi])       # ... through the kd-in-N kernel at full resolution
            # the conv bias is a no-op ahead of InstanceNorm(affine=False): skipped
            wk = None
            x8 = None
            if (USE_KDN and USE_TENSOR_CORES and dtype == torch.bfloat16 and not L.in_planar and L.cout in (8, 16)
                    and _tc_channels(L.cin) and d >= 4 and d * h * w >= 48 ** 3):
                wk = cache.conv3_kdn(tensors[L.wi])
            elif (USE_KDN and USE_TENSOR_CORES and INBLOCK_TC and dtype == torch.bfloat16 and L.in_planar and L.cin == 2
                    and L.cout in (8, 16) and d >= 4 and d * h * w >= 48 ** 3 and not SIMULATE_BF16):
                wk = cache.inblock_kdn(tensors[L.wi])
                if wk is not None:
                    x8 = ops.planar_to_ndhwc8(cur)
            skip = slots[L.skip_from] if L.skip_from is not None else None
            fuse = (FUSE_APPLY and USE_TENSOR_CORES and dtype == torch.bfloat16 and arena is not None and not SIMULATE_BF16
                    and (FUSE_APPLY_MAX_VOX == 0 or d * h * w <= FUSE_APPLY_MAX_VOX))
            a = None
            if x8 is not None and fuse:
                y, stats, a = ops.conv3_in_relu(x8, wk, (n, d, h, w), 8, L.cout, arena, skip=skip, kdn=True)
            elif x8 is not None:
                y, stats = ops.conv3_tc_kdn(x8, wk, (n, d, h, w), 8, L.cout, want_stats=True, arena=arena)
            elif wk is not None and fuse:
                y, stats, a = ops.conv3_in_relu(cur, wk, (n, d, h, w), L.cin, L.cout, arena, skip=skip, kdn=True)
            elif wk is not None:
                y, stats = ops.conv3_tc_kdn(cur, wk, (n, d, h, w), L.cin, L.cout, want_stats=True, arena=arena)
            elif fuse and not L.in_planar and wtc is not None and _tc_channels(L.cin) and L.cout % 8 == 0:
                y, stats, a = ops.conv3_in_relu(cur, wtc, (n, d, h, w), L.cin, L.cout, arena, skip=skip, kdn=False)
            else:
                y, stats = ops.conv3_fprop(cur, wf, None, (n, d, h, w), L.cin, L.cout, dtype, in_planar=L.in_planar,
                                           wtc=None if L.in_planar else wtc, arena=arena)
            y = _sim(y, "y")
            if a is None:
                a = _sim(ops.inorm_relu_apply(y, stats, skip), "a")
            if record:
                wkd = None
                if (USE_KDN and USE_TENSOR_CORES and dtype == torch.bfloat16 and not L.in_planar and L.cin in (8, 16)
                        and _tc_channels(L.cout) and d >= 4 and d * h * w >= 48 ** 3):
                    wkd = cache.conv3_kdn(tensors[L.wi], dgrad=True)       # input gradient through the kd-in-N kernel too
                tape.append((L, cur, y, stats, (n, d, h, w), (wd, wdtc, wkd, wkd_in, x8)))
            cur = a
        elif L.kind == K2DOWN:
            d, h, w = d // 2, h // 2, w // 2
            k2g, k2s = cache.k2s2(tensors[L.wi], L.cout, L.cin) if dtype == torch.bfloat16 else (None, None)
            out = _sim(ops.k2s2_gather(cur, tensors[L.wi].detach(), tensors[L.bi].detach(), (n, d, h, w), L.cout, L.cin, wtc=k2g), "k2")
            if record:
                tape.append((L, cur, None, None, (n, d, h, w), (k2g, k2s)))
            cur = out
        elif L.kind == K2UP:
            k2g, k2s = cache.k2s2(tensors[L.wi], L.cin, L.cout) if dtype == torch.bfloat16 else (None, None)
            out = _sim(ops.k2s2_scatter(cur, tensors[L.wi].detach(), tensors[L.bi].detach(), (n, d, h, w), L.cin, L.cout, wtc=k2s), "k2")
            if record:
                tape.append((L, cur, None, None, (n, d, h, w), (k2g, k2s)))
            d, h, w = d * 2, h * 2, w * 2
            cur = out
        elif L.kind == HEAD:
            wtc8 = wdtc = None
            if dtype == torch.bfloat16 and L.cout == 2 and _tc_channels(L.cin) and not L.in_planar:
                wtc8, wdtc = cache.head_tc(tensors[L.wi])
                if HEAD_DIRECT:          # A/B switch: fp32-weight CUDA-core head forward (backward stays on the tensor cores)
                    wtc8 = None
            wk8 = wkd8 = None
            if wtc8 is not None and USE_KDN and d >= 4 and d * h * w >= 48 ** 3:
                wk8, wkd8 = cache.head_kdn(tensors[L.wi])
            if wk8 is not None:
                # conv + bias + softmax + planar store in ONE kd-in-N launch
                wd = None
                probs = ops.conv3_tc_kdn_planar(cur, wk8, (n, d, h, w), L.cin, 1, bias=tensors[L.bi].detach())
            elif wtc8 is not None:
                # conv + bias + softmax + planar store in ONE tensor-core launch
                wd = None
                probs = ops.head_conv_softmax2(cur, wtc8, tensors[L.bi].detach(), (n, d, h, w), L.cin)
            else:
                wf, wd, _, _ = cache.conv3(tensors[L.wi], tc=False)
                logits, _ = ops.conv3_fprop(cur, wf, tensors[L.bi].detach(), (n, d, h, w), L.cin, L.cout, torch.float32,
                                            in_planar=L.in_planar, want_stats=False)
                probs = ops.softmax2_fwd(logits, (n, d, h, w))
            if record:
                tape.append((L, cur, probs, None, (n, d, h, w), (wd, wdtc, wkd8)))
            cur = probs
        else:
            raise RuntimeError("unknown layer kind %r" % (L.kind,))
        if L.save_as is not None:
            slots[L.save_as] = cur
    return cur, (n, d, h, w), tape


def program_backward(tape, g, dtype, need, grads, param_refs, need_input_grad):
    """Replays `tape` in reverse.  need[i]: whether tensor i wants a gradient; grads[i] receives
    the tensor handed back to autograd (None when accumulated in place into .grad).
    Returns the gradient w.r.t. the program input (or None)."""
    pending = {}                     # slot name -> gradient arriving through an additive skip
    # InstanceNorm-backward sums of every Conv+InstanceNorm layer: one zero-filled arena; where the gradient of a
    # layer's activation is produced by a tensor-core dgrad (and no skip gradient is added to it afterwards) the
    # reduction is fused into that dgrad's epilogue and only the apply pass remains.
    arena = ops.StatsArena(sum(e[4][0] * e[0].cout * 2 for e in tape if e[0].kind == C3IN), g.device)
    sums_of = {}
    fused = set()

    def fuse_prev(idx, dy_like, cin, cout, wdtc):
        """(y_prev, stats_prev, sums_prev) if the dgrad of tape[idx] may also reduce for tape[idx-1], else None."""
        if idx == 0 or SIMULATE_BF16 or not FUSE_BWD_REDUCE:
            return None
        Lp = tape[idx - 1][0]
        if Lp.kind != C3IN or Lp.bn is not None or (Lp.save_as is not None and Lp.save_as in pending):
            return None
        dd = tape[idx][4]
        if FUSE_BWD_REDUCE_MAX_VOX and dd[1] * dd[2] * dd[3] > FUSE_BWD_REDUCE_MAX_VOX:
            return None
        if not ops.dgrad_can_fuse_reduce(dy_like, cin, cout, dtype, False, wdtc, tape[idx][4]):
            return None
        fused.add(idx - 1)
        return tape[idx - 1][2], tape[idx - 1][3], sums_of[idx - 1]

    for idx, e in enumerate(tape):
        if e[0].kind == C3IN:
            sums_of[idx] = arena.take(e[4][0] * e[0].cout * 2).view(e[4][0], e[0].cout, 2)
    for idx in range(len(tape) - 1, -1, -1):
        L, x_in, y, stats, dims, wd, wt = tape[idx]
        first = idx == 0
        want_dx = (not first) or need_input_grad
        if L.save_as is not None and L.save_as in pending:
            g = ops.add_inplace(g, pending.pop(L.save_as))
        if L.kind == C3IN and L.bn is not None:
            if L.skip_from is not None:
                pending[L.skip_from] = g
            kb, k2b2, aux, was_training = stats
            S = float(dims[1] * dims[2] * dims[3])
            sums = ops.affine_relu_bwd_reduce(g, y, kb, k2b2)
            coef, dgamma, dbeta, dbias = _bn_backward_tables(L, sums, kb, k2b2, aux, S, was_training)
            dy = ops.affine_relu_bwd_apply(g, y, kb, coef)
            _accumulate_small(param_refs, need, grads, L.gi, dgamma)
            _accumulate_small(param_refs, need, grads, L.bti, dbeta)
            _accumulate_small(param_refs, need, grads, L.bi, dbias)
            if need[L.wi]:
                tgt, acc = _grad_target(param_refs[L.wi], True)
                dw = ops.conv3_wgrad(x_in, dy, dims, L.cin, L.cout, dw=tgt, in_planar=L.in_planar, accumulate=acc)[0]
                grads[L.wi] = None if acc else dw
            _grad_ready(param_refs[L.wi], param_refs[L.bi], param_refs[L.gi], param_refs[L.bti])
            g = ops.conv3_dgrad(dy, wd[0], dims, L.cin, L.cout, dtype, out_planar=L.in_planar, wdtc=wd[1]) if want_dx else None
        elif L.kind == C3IN:
            if L.skip_from is not None:
                pending[L.skip_from] = g
            dy = _sim(ops.inorm_relu_bwd(g, y, stats, sums=sums_of[idx], reduced=idx in fused), "dy")
            if need[L.wi]:
                tgt, acc = _grad_target(param_refs[L.wi], True)
                def run_wgrad(L=L, x_in=x_in, dy=dy, dims=dims, tgt=tgt, acc=acc, wd=wd):
                    if L.in_planar and dtype == torch.bfloat16 and USE_TENSOR_CORES and L.cin < 8 and L.cout % 8 == 0:
                        # in-block (Cin 1 or 2, planar fp32 input): pad the input to 8 bf16 channels and take the
                        # tensor-core wgrad; the padded input channels give zero rows that are dropped
                        x8 = wd[4] if len(wd) > 4 and wd[4] is not None else ops.planar_to_ndhwc8(x_in)
                        dw8, _ = ops.conv3_wgrad(x8, dy, dims, 8, L.cout)
                        if acc:
                            ops.atomic_add_rows(tgt, dw8, L.cout, L.cin * 27, 8 * 27)      # += dw8[:, :cin]; atomics:
                            # other per-sample backward chains may be adding onto the same .grad concurrently
                            return tgt
                        return dw8[:, :L.cin].contiguous()
                    return ops.conv3_wgrad(x_in, dy, dims, L.cin, L.cout, dw=tgt, in_planar=L.in_planar, accumulate=acc)[0]
                dw = _wgrad_async(run_wgrad, acc, x_in, dy, wd[4] if len(wd) > 4 else None)
                grads[L.wi] = None if acc else dw
            if need[L.bi]:
                # exactly zero: the bias cancels in InstanceNorm
                tgt, acc = _grad_target(param_refs[L.bi], True)
                grads[L.bi] = None if acc else torch.zeros(L.cout, device=dy.device, dtype=torch.float32)
            _grad_ready(param_refs[L.wi], param_refs[L.bi])
            if want_dx and len(wd) > 2 and wd[2] is not None:
                # full-resolution layers: kd-in-N kernel (GEMM input = the layer's Cout, output = its Cin in {8, 16})
                g, _ = ops.conv3_tc_kdn(dy, wd[2], dims, L.cout, L.cin, prev=fuse_prev(idx, dy, L.cin, L.cout, wd[1]))
            elif want_dx and len(wd) > 3 and wd[3] is not None and not SIMULATE_BF16:
                g = ops.conv3_tc_kdn_planar(dy, wd[3], dims, L.cout, 2)        # planar 2-channel input gradient (VAE in-block)
            else:
                g = _sim(ops.conv3_dgrad(dy, wd[0], dims, L.cin, L.cout, dtype, out_planar=L.in_planar, wdtc=wd[1],
                                         prev=None if L.in_planar else fuse_prev(idx, dy, L.cin, L.cout, wd[1])), "g") \
                    if want_dx else None
        elif L.kind == K2DOWN:
            # dims are the coarse (output) dims; g is the coarse gradient
            if need[L.wi] or need[L.bi]:
                tw, tb, acc = _pair_targets(param_refs, need, L.wi, L.bi, (L.cout, L.cin, 2, 2, 2), (L.cout,), g.device)
                _wgrad_async(lambda: ops.k2s2_wgrad(g, x_in, dims, L.cout, L.cin, dwt=tw, dbias_coarse=tb, accumulate=acc),
                             acc, g, x_in)
                _hand_back(grads, need, L.wi, L.bi, tw, tb, acc)
            _grad_ready(param_refs[L.wi], param_refs[L.bi])
            g = _sim(ops.k2s2_scatter(g, wt, None, dims, L.cout, L.cin, wtc=wd[1] if wd else None), "g") if want_dx else None
        elif L.kind == K2UP:
            # dims are the coarse (input) dims; g is the fine gradient
            if need[L.wi] or need[L.bi]:
                tw, tb, acc = _pair_targets(param_refs, need, L.wi, L.bi, (L.cin, L.cout, 2, 2, 2), (L.cout,), g.device)
                _wgrad_async(lambda: ops.k2s2_wgrad(x_in, g, dims, L.cin, L.cout, dwt=tw, dbias_fine=tb, accumulate=acc),
                             acc, g, x_in)
                _hand_back(grads, need, L.wi, L.bi, tw, tb, acc)
            _grad_ready(param_refs[L.wi], param_refs[L.bi])
            g = _sim(ops.k2s2_gather(g, wt, None, dims, L.cin, L.cout, wtc=wd[0] if wd else None), "g") if want_dx else None
        elif L.kind == HEAD:
            probs = y
            if wd[1] is not None:
                # bf16 mode: 8-channel padded logit gradient -> tensor-core wgrad / dgrad; bias gradient fused
                tw = tb = None
                acc = False
                if need[L.wi] or need[L.bi]:
                    tw, tb, acc = _pair_targets(param_refs, need, L.wi, L.bi, (L.cout, L.cin, 3, 3, 3), (L.cout,), g.device)
                    if tb is not None and not acc:
                        tb.zero_()
                dl8 = ops.softmax2_bwd_pad8(g, probs, dims, db=tb)
                if need[L.wi]:
                    def run_head_wgrad(x_in=x_in, dl8=dl8, dims=dims, tw=tw, acc=acc, L=L):
                        dw8, _ = ops.conv3_wgrad(x_in, dl8, dims, L.cin, 8)
                        if acc:
                            ops.atomic_add_rows(tw, dw8, 1, L.cout * L.cin * 27, dw8.numel())  # += dw8[:cout] (atomics)
                        else:
                            tw.copy_(dw8[:L.cout])
                    _wgrad_async(run_head_wgrad, acc, x_in, dl8)
                if need[L.wi] or need[L.bi]:
                    _hand_back(grads, need, L.wi, L.bi, tw, tb, acc)
                _grad_ready(param_refs[L.wi], param_refs[L.bi])
                if want_dx and len(wd) > 2 and wd[2] is not None and L.cin in (8, 16):
                    g, _ = ops.conv3_tc_kdn(dl8, wd[2], dims, 8, L.cin, prev=fuse_prev(idx, dl8, L.cin, 8, wd[1]))
                else:
                    g = ops.conv3_dgrad(dl8, None, dims, L.cin, 8, dtype, wdtc=wd[1],
                                        prev=fuse_prev(idx, dl8, L.cin, 8, wd[1])) if want_dx else None
                continue
            dlogits = _sim(ops.softmax2_bwd(g, probs, dims, dtype), "dy")
            if need[L.wi] or need[L.bi]:
                tw, tb, acc = _pair_targets(param_refs, need, L.wi, L.bi, (L.cout, L.cin, 3, 3, 3), (L.cout,), g.device)
                ops.conv3_wgrad(x_in, dlogits, dims, L.cin, L.cout, dw=tw, db=tb, in_planar=L.in_planar, accumulate=acc)
                _hand_back(grads, need, L.wi, L.bi, tw, tb, acc)
            _grad_ready(param_refs[L.wi], param_refs[L.bi])
            g = _sim(ops.conv3_dgrad(dlogits, wd[0], dims, L.cin, L.cout, dtype, out_planar=L.in_planar, wdtc=wd[1]), "g") \
                if want_dx else None
    return g


def _record_weights(tape, tensors):
    """Appends the (detached) k2s2 weight to the tape entries that need it in backward."""
    out = []
    for entry in tape:
        L = entry[0]
        wt = tensors[L.wi].detach() if L.kind in (K2DOWN, K2UP) else None
        out.append(entry + (wt,))
    return out


class ProgramFn(torch.autograd.Function):
    """One autograd node for a whole conv program (Segmentation, or a single block)."""

    @staticmethod
    def forward(ctx, spec, x, *tensors):
        layers, dtype, cache, param_refs = spec[:4]
        grad_on = spec[4] if len(spec) > 4 else True        # torch.is_grad_enabled() at apply time (always off in here)
        planar = layers[0].in_planar
        if planar:
            n, d, h, w = x.shape[0], x.shape[2], x.shape[3], x.shape[4]
        else:
            n, d, h, w = x.shape[0], x.shape[1], x.shape[2], x.shape[3]
        record = grad_on and any(ctx.needs_input_grad)       # needs_input_grad ignores torch.no_grad(): no tape for inference
        out, _, tape = program_forward(layers, tensors, x.contiguous(), (n, d, h, w), dtype, cache, record=record)
        ctx.tape = _record_weights(tape, tensors) if record else None
        ctx.dtype = dtype
        ctx.param_refs = param_refs
        ctx.ntensors = len(tensors)
        return out

    @staticmethod
    def backward(ctx, g):
        need = list(ctx.needs_input_grad[2:])
        grads = [None] * ctx.ntensors
        gx = program_backward(ctx.tape, g.contiguous(), ctx.dtype, need, grads, ctx.param_refs,
                              ctx.needs_input_grad[1])
        return (None, gx) + tuple(grads)


class VAEFn(torch.autograd.Function):
    """Shape VAE as one autograd node: encoder program -> fc_mean/fc_std/reparam -> fc2 ->
    decoder program -> softmax head.  Returns (recon, mean, std).  joint_model.py:227-272."""

    @staticmethod
    def forward(ctx, spec, x, z, scale, use_z, *tensors):
        enc, dec, fc, dtype, cache, param_refs, dim = spec[:7]
        grad_on = spec[7] if len(spec) > 7 else True
        n, d, h, w = x.shape[0], x.shape[2], x.shape[3], x.shape[4]
        record = grad_on and any(ctx.needs_input_grad)
        ctx.set_materialize_grads(False)
        hcur, (n, d2, h2, w2), tape_e = program_forward(enc, tensors, x.contiguous(), (n, d, h, w), dtype, cache, record)
        c = enc[-1].cout
        s3 = d2 * h2 * w2
        wm, bm, ws, bs, w2_, b2_ = [tensors[i].detach() for i in fc]
        if wm.shape[1] != s3 * c:
            raise RuntimeError("VAE: fc_mean expects flat dim %d but the encoder produced %d (patch size vs. "
                               "constructor `patch` mismatch)" % (wm.shape[1], s3 * c))
        mean, std, lat = ops.fc_encode_fwd(hcur, wm, bm, ws, bs, z, scale, use_z, n, s3, c, dim)
        hdec = ops.fc_decode_fwd(lat, w2_, b2_, n, s3, c, dim, dtype, d2)
        recon, _, tape_d = program_forward(dec, tensors, hdec, (n, d2, h2, w2), dtype, cache, record)
        if record:
            ctx.tape_e = _record_weights(tape_e, tensors)
            ctx.tape_d = _record_weights(tape_d, tensors)
            ctx.fc_saved = (hcur, z, scale, use_z, std, lat, (n, s3, c, dim), [tensors[i].detach() for i in fc])
        ctx.spec_small = (fc, dtype, param_refs)
        ctx.ntensors = len(tensors)
        return recon, mean, std

    @staticmethod
    def backward(ctx, g_recon, g_mean, g_std):
        fc, dtype, param_refs = ctx.spec_small
        need = list(ctx.needs_input_grad[5:])
        grads = [None] * ctx.ntensors
        hcur, z, scale, use_z, std, lat, (n, s3, c, dim), fcw = ctx.fc_saved
        wm, bm, ws, bs, w2_, b2_ = fcw
        i_wm, i_bm, i_ws, i_bs, i_w2, i_b2 = fc
        dlat = None
        if g_recon is not None:
            dh = program_backward(ctx.tape_d, g_recon.contiguous(), dtype, need, grads, param_refs, True)
            if need[i_w2] or need[i_b2]:
                tw, tb, acc = _pair_targets(param_refs, need, i_w2, i_b2, tuple(w2_.shape), tuple(b2_.shape), dh.device)
                dlat = ops.fc_decode_bwd(dh, lat, w2_, n, s3, c, dim, dw2=tw, db2=tb, accumulate=acc)
                _hand_back(grads, need, i_w2, i_b2, tw, tb, acc)
            else:
                dlat = ops.fc_decode_bwd(dh, lat, w2_, n, s3, c, dim)
        enc_params = any(need[i] for i in (i_wm, i_bm, i_ws, i_bs))
        # anything upstream of the latent?
        upstream = ctx.needs_input_grad[1] or enc_params or any(need[e[0].wi] for e in ctx.tape_e)
        gx = None
        if upstream and (dlat is not None or g_mean is not None or g_std is not None):
            tgts = {}
            acc_flags = set()
            for i, like in ((i_wm, wm), (i_bm, bm), (i_ws, ws), (i_bs, bs)):
                t, a = _grad_target(param_refs[i], need[i])
                if need[i] and t is None:
                    t = torch.empty_like(like)
                    a = False
                tgts[i] = t
                if need[i]:
                    acc_flags.add(a)
            if enc_params and (len(acc_flags) != 1 or not all(need[i] for i in (i_wm, i_bm, i_ws, i_bs))):
                raise RuntimeError("fc_mean / fc_std parameters must share requires_grad and .grad allocation state")
            acc = acc_flags.pop() if acc_flags else False
            dhe = ops.fc_encode_bwd(hcur, wm, ws, z, scale, use_z, std, dlat,
                                    g_mean.contiguous() if g_mean is not None else None,
                                    g_std.contiguous() if g_std is not None else None,
                                    n, s3, c, dim, want_dx=True, dwm=tgts[i_wm], dbm=tgts[i_bm], dws=tgts[i_ws],
                                    dbs=tgts[i_bs], accumulate=acc)
            for i in (i_wm, i_bm, i_ws, i_bs):
                grads[i] = None if (acc or not need[i]) else tgts[i]
            gx = program_backward(ctx.tape_e, dhe, dtype, need, grads, param_refs, ctx.needs_input_grad[1])
        return (None, gx, None, None, None) + tuple(grads)


class DecodeFn(torch.autograd.Function):
    """VAE decoder only (mid_input=True, joint_model.py:251-271): latent -> recon."""

    @staticmethod
    def forward(ctx, spec, lat, *tensors):
        dec, fc, dtype, cache, param_refs, dim, side = spec[:7]
        grad_on = spec[7] if len(spec) > 7 else True
        w2_, b2_ = tensors[fc[4]].detach(), tensors[fc[5]].detach()
        n = lat.shape[0]
        c = dec[0].cin
        s3 = side ** 3
        record = grad_on and any(ctx.needs_input_grad)
        lat = lat.contiguous().float()
        hdec = ops.fc_decode_fwd(lat, w2_, b2_, n, s3, c, dim, dtype, side)
        recon, _, tape = program_forward(dec, tensors, hdec, (n, side, side, side), dtype, cache, record)
        if record:
            ctx.tape = _record_weights(tape, tensors)
            ctx.saved = (lat, w2_, b2_, (n, s3, c, dim))
        ctx.small = (fc, dtype, param_refs)
        ctx.ntensors = len(tensors)
        return recon

    @staticmethod
    def backward(ctx, g):
        fc, dtype, param_refs = ctx.small
        need = list(ctx.needs_input_grad[2:])
        grads = [None] * ctx.ntensors
        lat, w2_, b2_, (n, s3, c, dim) = ctx.saved
        dh = program_backward(ctx.tape, g.contiguous(), dtype, need, grads, param_refs, True)
        i_w2, i_b2 = fc[4], fc[5]
        if need[i_w2] or need[i_b2]:
            tw, tb, acc = _pair_targets(param_refs, need, i_w2, i_b2, tuple(w2_.shape), tuple(b2_.shape), dh.device)
            dlat = ops.fc_decode_bwd(dh, lat, w2_, n, s3, c, dim, dw2=tw, db2=tb, accumulate=acc)
            _hand_back(grads, need, i_w2, i_b2, tw, tb, acc)
        else:
            dlat = ops.fc_decode_bwd(dh, lat, w2_, n, s3, c, dim)
        return (None, dlat if ctx.needs_input_grad[1] else None) + tuple(grads)


class LinearFn(torch.autograd.Function):
    """y = act(x @ W.T + b) through vs_linear_fwd / vs_linear_bwd (the encoder / discriminator head,
    joint_model.py:287-304)."""

    @staticmethod
    def forward(ctx, x, w, b, act):
        x = x.contiguous().float()
        y = ops.linear_fwd(x, w.detach(), b.detach(), act)
        ctx.save_for_backward(x, w.detach(), y)
        ctx.act = act
        ctx.refs = (w, b)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, w, y = ctx.saved_tensors
        wref, bref = ctx.refs
        need_x, need_w, need_b = ctx.needs_input_grad[0], ctx.needs_input_grad[1], ctx.needs_input_grad[2]
        tw, aw = _grad_target(wref, need_w)
        tb, ab = _grad_target(bref, need_b)
        if need_w and tw is None:
            tw = torch.empty_like(w)
        if need_b and tb is None:
            tb = torch.empty_like(y[0])
        acc = bool(aw and (ab or not need_b))
        if need_w and need_b and aw != ab:
            raise RuntimeError("Linear weight and bias must both have (or both lack) a .grad buffer")
        dx = ops.linear_bwd(x, w, y, dy.contiguous(), ctx.act, want_dx=need_x, dw=tw if need_w else None,
                            db=tb if need_b else None, accumulate=acc)
        return dx, (None if (acc or not need_w) else tw), (None if (acc or not need_b) else tb), None
