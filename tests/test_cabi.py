"""CPU-side checks of the C-ABI boundary: the library loads, exports every symbol the header
declares, and the Python host layer refuses CPU tensors (no fallback)."""
import ctypes
import os
import re

import pytest
import torch

from vae_segmentation_b200 import _cabi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    text = open(os.path.join(ROOT, "include", "vaeseg_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(vs_[a-z0-9_]+)\s*\(", text)))


def test_library_built_and_loads():
    assert os.path.isfile(_cabi.LIB_PATH), "run __graft_entry__.build() first"
    lib = _cabi.lib()
    assert lib.vs_version() >= 100


def test_every_declared_symbol_is_exported_and_bound():
    handle = ctypes.CDLL(_cabi.LIB_PATH)
    declared = _header_symbols()
    assert len(declared) >= 30
    for name in declared:
        assert hasattr(handle, name), "header declares %s but the library does not export it" % name
        assert name in _cabi.EXPORTED_SYMBOLS, "%s has no ctypes signature in _cabi.py" % name
    for name in _cabi.EXPORTED_SYMBOLS:
        assert name in declared, "%s bound in _cabi.py but not declared in include/vaeseg_b200.h" % name


def test_error_reporting_without_gpu():
    # argument validation happens before any CUDA call, so this is safe on a CPU-only box
    rc = _cabi.lib().vs_dice_sums(None, None, 0, None, 0, 0, 0, None)
    assert rc < 0
    assert "dice_sums" in _cabi.last_error()
    with pytest.raises(RuntimeError, match="dice_sums"):
        _cabi.call("vs_dice_sums", None, None, 0, None, 0, 0, 0, None)


def test_cpu_tensors_are_rejected():
    from vae_segmentation_b200 import evaluation, joint_model
    seg = joint_model.Segmentation(1, 2, norm_type=1)
    with pytest.raises(RuntimeError, match="CUDA"):
        seg({"img": torch.zeros(1, 1, 16, 16, 16)}, "img", "pred")
    with pytest.raises(RuntimeError, match="CUDA"):
        evaluation.avg_dsc({"a": torch.zeros(1, 2, 4, 4, 4), "b": torch.zeros(1, 2, 4, 4, 4)}, "a", "b")
    with pytest.raises(NotImplementedError):
        joint_model.Segmentation(1, 2, norm_type=3)          # GSNorm3d: the one normalisation that is not implemented
