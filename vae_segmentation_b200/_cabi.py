"""ctypes binding of libvaeseg_b200.so (declared in include/vaeseg_b200.h).

The shared library is built in-tree by csrc/build.sh (see __graft_entry__.build).  There is
no fallback: if the library is missing or a call fails, a RuntimeError is raised.
"""
import ctypes
import os
from ctypes import c_char_p, c_float, c_int, c_longlong, c_size_t, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("VAESEG_LIB") or os.path.join(_HERE, "csrc", "libvaeseg_b200.so")      # VAESEG_LIB: A/B builds (tools)

VS_F32, VS_BF16 = 0, 1
VS_FLAG_PREZEROED = 1
TGT_TENSOR, TGT_BINARIZE, TGT_CONFIDENT, TGT_LABEL, TGT_ARGMAX = 0, 1, 2, 3, 4

_P, _I, _L, _F, _Z = c_void_p, c_int, c_longlong, c_float, c_size_t

# name -> argtypes (restype is int unless listed in _RESTYPES)
_SIGNATURES = {
    "vs_last_error_string": [],
    "vs_version": [],
    "vs_has_tcgen05": [],
    "vs_set_pdl": [_I],
    "vs_set_kdn_ordered": [_I],
    "vs_pack_conv3_weight": [_P, _P, _P, _I, _I, _P],
    "vs_conv3_tc_pack_bytes": [_I, _I, _I],
    "vs_pack_conv3_weight_tc": [_P, _P, _I, _I, _I, _P],
    "vs_pack_conv3_batched": [_P, _I, _P],
    "vs_pack_conv3_weight_tc_padded": [_P, _P, _I, _I, _I, _I, _I, _P],
    "vs_head_conv_softmax2_fwd": [_P, _P, _P, _P, _I, _I, _I, _I, _I, _P],
    "vs_conv3_tc_kdn_pack_bytes": [_I, _I, _I],
    "vs_pack_conv3_weight_tc_kdn": [_P, _P, _I, _I, _I, _P],
    "vs_pack_conv3_weight_tc_kdn_padded": [_P, _P, _I, _I, _I, _I, _I, _P],
    "vs_conv3x3x3_tc_kdn_planar": [_P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _P],
    "vs_conv3x3x3_tc_in_relu": [_P, _P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _P],
    "vs_conv3x3x3_tc_kdn_in_relu": [_P, _P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _P],
    "vs_conv3x3x3_tc_kdn": [_P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _P],
    "vs_conv3x3x3_tc_kdn_ex": [_P, _P, _P, _P, _P, _I, _P, _P, _P, _I, _I, _I, _I, _I, _I, _P],
    "vs_conv3x3x3_fprop": [_I, _I, _I, _I, _I, _P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _P],
    "vs_conv3x3x3_dgrad": [_I, _I, _I, _P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _P],
    "vs_conv3_wgrad_workspace_bytes": [_I, _I, _I, _I, _I, _I],
    "vs_conv3x3x3_wgrad": [_I, _I, _P, _P, _P, _P, _P, _Z, _I, _I, _I, _I, _I, _I, _I, _P],
    "vs_k2s2_gather": [_I, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _P],
    "vs_k2s2_scatter": [_I, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _P],
    "vs_k2s2_tc_pack_bytes": [_I, _I, _I],
    "vs_pack_k2s2_weight_tc": [_P, _P, _I, _I, _I, _P],
    "vs_k2s2_gather_tc": [_P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _P],
    "vs_k2s2_scatter_tc": [_P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _P],
    "vs_k2s2_wgrad": [_I, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _P],
    "vs_inorm_relu_apply": [_I, _P, _P, _P, _P, _I, _L, _I, _P],
    "vs_inorm_relu_bwd_reduce": [_I, _P, _P, _P, _P, _I, _L, _I, _I, _P],
    "vs_inorm_relu_bwd_apply": [_I, _P, _P, _P, _P, _P, _I, _L, _I, _P],
    "vs_add_inplace": [_I, _P, _P, _L, _P],
    "vs_affine_relu_apply": [_I, _P, _P, _P, _P, _I, _L, _I, _P],
    "vs_affine_relu_bwd_reduce": [_I, _P, _P, _P, _P, _P, _I, _L, _I, _P],
    "vs_affine_relu_bwd_apply": [_I, _P, _P, _P, _P, _P, _I, _L, _I, _P],
    "vs_softmax2_fwd": [_P, _P, _I, _L, _P],
    "vs_softmax2_bwd": [_I, _P, _P, _P, _I, _L, _P],
    "vs_softmax2_bwd_pad8": [_P, _P, _P, _P, _I, _L, _P],
    "vs_planar_to_ndhwc8": [_P, _P, _I, _I, _L, _P],
    "vs_fc_encode_fwd": [_I, _P, _P, _P, _P, _P, _P, _F, _I, _P, _P, _P, _I, _I, _I, _I, _P],
    "vs_fc_decode_fwd": [_I, _P, _P, _P, _P, _I, _I, _I, _I, _P],
    "vs_fc_decode_bwd": [_I, _P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _P],
    "vs_fc_encode_bwd": [_I, _P, _P, _P, _P, _F, _I, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _P],
    "vs_linear_fwd": [_P, _P, _P, _P, _I, _I, _I, _I, _P],
    "vs_linear_bwd": [_P, _P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _P],
    "vs_dice_sums": [_P, _P, _I, _P, _I, _I, _L, _P],
    "vs_dice_bwd": [_P, _P, _I, _P, _P, _F, _P, _P, _I, _I, _I, _L, _P],
    "vs_kl_fwd": [_P, _P, _P, _I, _I, _P],
    "vs_kl_bwd": [_P, _P, _P, _P, _P, _I, _I, _P],
    "vs_binarize": [_P, _P, _I, _L, _P],
    "vs_one_hot": [_P, _P, _I, _I, _L, _P],
    "vs_clip_center": [_I, _P, _P, _L, _F, _F, _F, _F, _P],
    "vs_crop_resize": [_P, _I, _I, _I, _P, _I, _P, _I, _I, _I, _I, _I, _P, _P, _P],
    "vs_gauss_radius": [ctypes.c_double],
    "vs_sgd_step": [_P, _P, _P, _L, _F, _F, _I, _F, _P],
    "vs_adam_step": [_P, _P, _P, _P, _L, _F, _F, _F, _F, _I, _F, _P],
    "vs_ema_update": [_P, _P, _L, _F, _P],
    "vs_compose_target_loss": [_P, _F, _I, _I, _P, _P, _P],
    "vs_atomic_add_rows": [_P, _P, _L, _L, _L, _P],
    "vs_joint_target_finish": [_P, _P, _P, _P, _I, _I, _I, _I, _F, _F, _I, _I, _I, _P, _P, _P],
}
_RESTYPES = {"vs_last_error_string": c_char_p, "vs_conv3_wgrad_workspace_bytes": c_size_t,
             "vs_conv3_tc_pack_bytes": c_size_t, "vs_conv3_tc_kdn_pack_bytes": c_size_t, "vs_k2s2_tc_pack_bytes": c_size_t,
             "vs_set_kdn_ordered": None}

EXPORTED_SYMBOLS = tuple(_SIGNATURES)

_lib = None
PDL_DEFAULT = "0"
KDN_ORDERED_DEFAULT = "0"


def lib():
    """Loads the shared library once; raises if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.isfile(LIB_PATH):
            raise RuntimeError(
                "vaeseg_b200: %s is missing -- run `python -c 'import __graft_entry__ as g; g.build()'` "
                "(or vae_segmentation_b200/csrc/build.sh); there is no CPU fallback" % LIB_PATH)
        handle = ctypes.CDLL(LIB_PATH)
        for name, argtypes in _SIGNATURES.items():
            fn = getattr(handle, name)
            fn.argtypes = argtypes
            fn.restype = _RESTYPES.get(name, c_int)
        # programmatic dependent launch for launches of at most VAESEG_PDL CTAs (include/vaeseg_b200.h: vs_set_pdl); 0 disables
        handle.vs_set_pdl(int(os.environ.get("VAESEG_PDL", PDL_DEFAULT)))
        handle.vs_set_kdn_ordered(int(os.environ.get("VAESEG_KDN_ORDERED", KDN_ORDERED_DEFAULT)))
        if os.environ.get("VAESEG_NORM_VPT"):                            # A/B switch (tools): grid sizing of the apply passes
            handle.vs_debug_set_norm_vpt(int(os.environ["VAESEG_NORM_VPT"]))
        if os.environ.get("VAESEG_CONV3_KSPLIT", "1") == "0":          # A/B switch (tools): K split over the conv issuer warps
            handle.vs_debug_set_conv3_ksplit(0)
        _lib = handle
    return _lib


def last_error():
    msg = lib().vs_last_error_string()
    return msg.decode("utf-8", "replace") if msg else ""


N_CALLS = 0            # C-ABI calls made so far (each launches >= 1 kernel); bench.py reports the delta
_profiler = None       # optional callable(name, key, launch) installed by bench.py / tools


def call_key(name, args):
    """Shape signature of a call: its non-pointer arguments."""
    return tuple(a for a, t in zip(args, _SIGNATURES[name]) if t is not _P)


def set_profiler(fn):
    global _profiler
    _profiler = fn


def call(name, *args):
    """Calls an int-returning entry point; a negative status raises RuntimeError."""
    global N_CALLS
    N_CALLS += 1
    fn = getattr(lib(), name)
    if _profiler is not None:
        rc = _profiler(name, args, fn)
    else:
        rc = fn(*args)
    if rc != 0:
        raise RuntimeError("vaeseg_b200.%s failed (status %d): %s" % (name, rc, last_error()))
