"""Host-side image preprocessing (starst3r_b200/image.py) vs the behaviour of starster/image.py:43-139."""
import numpy as np
import pytest
import torch

from starst3r_b200 import image


def test_process_image_shapes_and_range():
    g = torch.Generator().manual_seed(0)
    img = torch.rand(3, 1080, 1920, generator=g)
    out = image.process_image(img, 1920)
    assert out.shape == (3, 1072, 1920)                 # half extents are multiples of 8 (SURVEY: 1080p -> 1920x1072)
    assert out.dtype == torch.float32 and -1.0001 <= float(out.min()) and float(out.max()) <= 1.0001
    out = image.process_image(torch.rand(3, 300, 400, generator=g), 224)
    assert out.shape == (3, 160, 224) and out.shape[1] % 16 == 0 and out.shape[2] % 16 == 0
    # an already conforming image is only normalised
    x = torch.rand(3, 64, 96, generator=g)
    assert torch.allclose(image.process_image(x, 96), x * 2 - 1, atol=1e-5)


def test_process_image_matches_torchvision_pipeline():
    tv = pytest.importorskip("torchvision.transforms")
    g = torch.Generator().manual_seed(1)
    img = torch.rand(3, 333, 517, generator=g)
    size = 256
    new_size = [int(x * size / max(img.shape[1:])) for x in img.shape[1:]]
    ref = tv.functional.resize(img, new_size, tv.InterpolationMode.BICUBIC)          # what image.py:62 calls
    cx, cy = ref.shape[2] // 2, ref.shape[1] // 2
    wh, hh = (cx // 8) * 8, (cy // 8) * 8
    ref = tv.Normalize(mean=(0.5,) * 3, std=(0.5,) * 3)(ref[..., cy - hh:cy + hh, cx - wh:cx + wh])
    assert torch.allclose(image.process_image(img, size), ref, atol=1e-5)


def test_load_image_and_mast3r_format(tmp_path):
    from PIL import Image
    arr = (np.random.default_rng(0).random((120, 200, 3)) * 255).astype(np.uint8)
    Image.fromarray(arr).save(tmp_path / "a.png")
    imgs = image.load_images([tmp_path / "a.png", str(tmp_path / "a.png")], size=128)
    assert imgs[0].shape == (3, 64, 128) and torch.equal(imgs[0], imgs[1])
    d = image.prepare_images_for_mast3r(imgs)
    assert d[1]["img"].shape == (1, 3, 64, 128) and d[1]["idx"] == 1 and d[1]["instance"] == "1"
    assert d[0]["true_shape"].tolist() == [[64, 128]] and d[0]["true_shape"].dtype == np.int32
    assert image.make_pair_indices(3) == [(1, 0), (2, 0), (2, 1), (0, 1), (0, 2), (1, 2)]
