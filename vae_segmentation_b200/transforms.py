"""Device-side versions of the reference's per-sample input transforms.

At > 300 volumes/s per GPU the reference's CPU DataLoader (numpy transforms in 16 worker processes) cannot feed the
step; the elementwise part of that pipeline runs here as one fused kernel on the batch already resident in HBM:

    Clip(fields, new_min, new_max)  ->  CenterIntensities(fields, subtrahend, divisor)      utils/utils.py:508-533,572-618
    (main_target.py:223-224, main_source.py:211-212:  Clip(-200, 400), CenterIntensities(100, 300))

`ClipCenter` keeps the reference's dict-transform calling convention (`transform(data_dict) -> data_dict`) for CUDA
tensors; int16 inputs (raw Hounsfield units) are accepted so that the host->device copy moves 2 bytes per voxel.
The geometric transforms (CropResize: skimage.transform.resize) stay on the host: skimage is not in this image, so
their arithmetic cannot be pinned against the reference.
"""
from . import ops


class ClipCenter(object):
    def __init__(self, fields, new_min=-200.0, new_max=400.0, subtrahend=100.0, divisor=300.0):
        self.fields = list(fields)
        self.new_min, self.new_max = float(new_min), float(new_max)
        self.subtrahend, self.divisor = float(subtrahend), float(divisor)

    def __call__(self, data_dict):
        for f in self.fields:
            x = data_dict.get(f)
            if x is not None:
                data_dict[f] = ops.clip_center(x.contiguous(), self.new_min, self.new_max, self.subtrahend, self.divisor)
        return data_dict
