"""TEST INFRASTRUCTURE ONLY -- imports the real reference (authoring container only).

/root/reference is read-only and does not exist on the GPU box, so this module is used
only (a) by oracle/make_golden.py to produce tests/golden/ fixtures and (b) by the
`not gpu` tests that pin oracle/ref_torch.py against the real code when the tree is
present.  VAE.forward and the losses hard-require CUDA types (joint_model.py:246,
utils/evaluation.py:52-66); the three-line shim below maps them to CPU.
"""
import importlib
import os
import sys

import torch

REFERENCE_ROOT = os.environ.get("VAESEG_REFERENCE_ROOT", "/root/reference")


def available():
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "joint_model.py"))


def load():
    """Returns (joint_model module, utils.evaluation module) of the real reference."""
    if not available():
        raise RuntimeError("reference tree not present at %s" % REFERENCE_ROOT)
    torch.cuda.FloatTensor = torch.FloatTensor
    torch.cuda.LongTensor = torch.LongTensor
    torch.Tensor.cuda = lambda self, *a, **k: self
    saved = {k: sys.modules.pop(k) for k in ("joint_model", "utils", "utils.evaluation") if k in sys.modules}
    sys.path.insert(0, REFERENCE_ROOT)
    try:
        jm = importlib.import_module("joint_model")
        ev = importlib.import_module("utils.evaluation")
    finally:
        sys.path.remove(REFERENCE_ROOT)
        for k in ("joint_model", "utils", "utils.evaluation"):
            sys.modules.pop(k, None)
        sys.modules.update(saved)
    return jm, ev
