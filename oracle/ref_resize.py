"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference's CropResize transform (utils/utils.py:220-293).

The reference calls `skimage.transform.resize` (scikit-image 0.18.3, requirements.txt:94), which is NOT in this image;
its n-D code path is a thin layer over `scipy.ndimage` (present here, so the interpolation arithmetic itself is the
third-party original, not a re-derivation):

    factors = input_shape / output_shape
    if anti_aliasing (default True for non-bool images; the label call passes anti_aliasing=False):
        image = ndi.gaussian_filter(image, sigma=max(0, (factors - 1) / 2), cval=0, mode='mirror')
    coords[i] = factors[i] * (arange(out_i) + 0.5) - 0.5
    out = ndi.map_coordinates(image, meshgrid(coords, indexing='ij'), order=order, mode='mirror', cval=0)
    (clip to the input range when order > 0: a no-op for linear interpolation)

with skimage's mode 'reflect' (the default) mapped to ndimage's 'mirror' and float32 images kept float32.
Parity pinning: skimage itself is absent, so this restatement is pinned only against scipy.ndimage's own semantics
(tests/test_oracle.py: identity resize, constant volumes, known 1-D cases); the crop / pad index arithmetic follows the
reference line by line.  Nothing in the shipped package imports this file.
"""
import numpy as np
from scipy import ndimage as ndi


def resize(image, output_shape, order=1, anti_aliasing=True):
    """skimage.transform.resize(image, output_shape, order=order, anti_aliasing=anti_aliasing) for n-D float volumes
    (skimage 0.18.3 transform/_warps.py:resize, n-dimensional branch)."""
    image = np.asarray(image)
    input_shape = image.shape
    factors = np.asarray(input_shape, dtype=float) / np.asarray(output_shape, dtype=float)
    if anti_aliasing:
        sigma = np.maximum(0, (factors - 1) / 2)
        image = ndi.gaussian_filter(image, sigma, cval=0, mode="mirror")
    coord_arrays = [factors[i] * (np.arange(d) + 0.5) - 0.5 for i, d in enumerate(output_shape)]
    coord_map = np.array(np.meshgrid(*coord_arrays, sparse=False, indexing="ij"))
    out = ndi.map_coordinates(image, coord_map, order=order, mode="mirror", cval=0)
    if order != 0:
        out = np.clip(out, image.min(), image.max())
    return out


def crop_window(label, shift=0):
    """Bounding-box cube of utils/utils.py:253-276: returns (start[3], stop[3], side) BEFORE clamping to the volume."""
    index = np.array(np.where(label > 0)).T
    if index.shape[0] > 0:
        bbox_max, bbox_min = np.max(index, 0), np.min(index, 0)
        center = (bbox_max + bbox_min) // 2
        L = int(np.max(bbox_max - bbox_min))
    else:
        center, L = np.array([64, 64, 64]), 32
    pad_width = int(L * 0.1)
    start = center - L // 2 - pad_width + shift
    stop = center + L // 2 + pad_width + shift
    return start, stop, L + pad_width * 2


def crop_resize(img, label, output_size, shift=0):
    """CropResize.__call__ for one field (utils/utils.py:253-291): returns (image, label, ori_shape)."""
    start, stop, side = crop_window(label, shift)
    ori_shape = list(label.shape)

    def cut(vol):
        sl = tuple(slice(max(int(start[a]), 0), min(int(stop[a]), vol.shape[a])) for a in range(3))
        vol = vol[sl]
        diff = list(side - np.array(vol.shape))
        return np.pad(vol, [(int(d / 2), d - int(d / 2)) for d in diff])

    lab_c, img_c = cut(label), cut(img)
    ori_shape += list(lab_c.shape)
    return (resize(img_c, output_size), resize(lab_c, output_size, order=0, anti_aliasing=False), np.array(ori_shape))
