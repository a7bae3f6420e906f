"""Device-side versions of the reference's per-sample input transforms.

At > 300 volumes/s per GPU the reference's CPU DataLoader (numpy transforms in 16 worker processes) cannot feed the
step; the elementwise part of that pipeline runs here as one fused kernel on the batch already resident in HBM:

    Clip(fields, new_min, new_max)  ->  CenterIntensities(fields, subtrahend, divisor)      utils/utils.py:508-533,572-618
    (main_target.py:223-224, main_source.py:211-212:  Clip(-200, 400), CenterIntensities(100, 300))

`ClipCenter` keeps the reference's dict-transform calling convention (`transform(data_dict) -> data_dict`) for CUDA
tensors; int16 inputs (raw Hounsfield units) are accepted so that the host->device copy moves 2 bytes per voxel.
`CropResize` is the reference's geometric transform (utils/utils.py:220-293: bounding-box cube around the pancreas label,
zero padding, skimage.transform.resize to the patch size -- linear + anti-aliased for the image, nearest for the label) as
device kernels (csrc/resize.cu): the bounding box is reduced with torch on the device (one small host read of six
integers per volume), the cube is never materialised, and the resampled patch lands in HBM in the in-block's layout.
"""
import torch

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


class CropResize(object):
    """utils/utils.py:220-293 for CUDA tensors: `data_dict[f]` ([D,H,W] fp32 image) and `data_dict[f + '_pancreas']`
    (label) are replaced by their `output_size` patches, `data_dict['ori_shape']` is set as the reference does.  The
    `_pancreas_pred` branch of the reference (:236-251, an inference-time path none of the shipped scripts reach) is
    not implemented."""

    def __init__(self, fields, output_size, pad=32, shift=0):
        self.fields = list(fields)
        self.output_size = [int(v) for v in output_size]
        self.pad, self.shift = pad, int(shift)

    @staticmethod
    def window(label, shift):
        """(start[3], stop[3], side) of the bounding-box cube before clamping (utils/utils.py:253-266)."""
        nz = torch.nonzero(label > 0)
        if nz.shape[0] > 0:
            bmax, bmin = nz.max(0).values, nz.min(0).values
            center = torch.div(bmax + bmin, 2, rounding_mode="floor")
            L = int((bmax - bmin).max().item())
            center = [int(v) for v in center.tolist()]
        else:
            center, L = [64, 64, 64], 32
        pw = int(L * 0.1)
        start = [c - L // 2 - pw + shift for c in center]
        stop = [c + L // 2 + pw + shift for c in center]
        return start, stop, L + 2 * pw

    def __call__(self, data_dict):
        for f in self.fields:
            img = data_dict.get(f)
            if img is None:
                continue
            if isinstance(data_dict.get(f + "_pancreas_pred"), torch.Tensor):
                raise NotImplementedError("CropResize: the '_pancreas_pred' branch (utils/utils.py:236-251) is not implemented")
            label = data_dict[f + "_pancreas"]
            if not (img.is_cuda and label.is_cuda):
                raise RuntimeError("vaeseg_b200.CropResize: volumes must be CUDA tensors (no CPU fallback)")
            start, stop, side = self.window(label, self.shift)
            crop9, cropped = [], []
            for a in range(3):
                lo, hi = max(start[a], 0), min(stop[a], label.shape[a])
                crop9.append(lo)
                cropped.append(hi - lo)
            crop9 += cropped + [int((side - c) / 2) for c in cropped]
            data_dict["ori_shape"] = torch.tensor(list(label.shape) + [side, side, side])
            data_dict[f] = ops.crop_resize(img.float().contiguous(), crop9, side, self.output_size, order=1, anti_alias=True)
            data_dict[f + "_pancreas"] = ops.crop_resize(label.float().contiguous(), crop9, side, self.output_size, order=0,
                                                        anti_alias=False)
        return data_dict
