"""Image loading / preprocessing with the signatures of starster/image.py (the input format of the hot path:
`Scene.add_images(model, imgs)` takes the tensors these functions produce).  Host-side, not a hot path.

process_image (image.py:43-76): resize the longest edge to `size` (bicubic, antialiased, as torchvision's
tensor resize does), centre-crop so that each HALF extent is a multiple of 8 (so H and W end up multiples of 16: the
MASt3R patch size; a 1080p frame becomes 1920x1072), normalise to [-1, 1]."""
from pathlib import Path

import numpy as np
import torch

__all__ = ("make_pair_indices", "process_image", "load_image", "load_images", "prepare_images_for_mast3r")


def make_pair_indices(n: int, symmetric: bool = True):
    """image.py:25-40: (i, j) for j < i, then the mirrored pairs."""
    pairs = [(i, j) for i in range(n) for j in range(i)]
    if symmetric:
        pairs += [(j, i) for i, j in pairs]
    return pairs


def process_image(img, size: int) -> torch.Tensor:
    """img: (C, H, W) float tensor in [0, 1] (or array).  Returns (C, H', W') float32 in [-1, 1]."""
    img = torch.as_tensor(img)
    h, w = img.shape[1:]
    new_h, new_w = (int(x * size / max(h, w)) for x in (h, w))
    work = img[None].float()
    work = torch.nn.functional.interpolate(work, size=(new_h, new_w), mode="bicubic", align_corners=False, antialias=True)
    img = work[0].to(img.dtype if img.is_floating_point() else torch.float32)
    cy, cx = img.shape[1] // 2, img.shape[2] // 2
    hh, wh = (cy // 8) * 8, (cx // 8) * 8
    img = img[..., cy - hh:cy + hh, cx - wh:cx + wh]
    return (img - 0.5) / 0.5


def load_image(path, size: int = 224) -> torch.Tensor:
    """image.py:79-101: EXIF-transposed RGB file -> process_image."""
    from PIL import Image
    from PIL.ImageOps import exif_transpose
    img = exif_transpose(Image.open(Path(path))).convert("RGB")
    arr = torch.from_numpy(np.array(img)).permute(2, 0, 1).float() / 255.0      # torchvision ToTensor on uint8 HWC
    return process_image(arr, size)


def load_images(paths, size: int = 224):
    return [load_image(p, size) for p in paths]


def prepare_images_for_mast3r(imgs):
    """image.py:112-139: the dict format MASt3R's legacy code expects."""
    return [dict(img=im[None], true_shape=np.int32([im.shape[-2:]]), idx=i, instance=str(i)) for i, im in enumerate(imgs)]
