"""Pins tests/data_ref.py (the checker of the GPU input pipeline) against torchvision's own to_tensor / normalize, which is what the
reference's datasets/transforms.py:233-250 call, and against the padding rule of util/collate_fn.py:24-41."""
import numpy as np
import torch
import torchvision.transforms.functional as F
from PIL import Image

from data_ref import reference_collate


def test_reference_collate_is_torchvision_plus_padding():
    g = torch.Generator().manual_seed(0)
    sizes = [(37, 53), (64, 40), (5, 64)]
    images = [torch.randint(0, 256, (h, w, 3), dtype=torch.uint8, generator=g) for h, w in sizes]
    got = reference_collate(images)
    H, W = 64, 64
    assert got.tensors.shape == (3, 3, H, W) and got.mask.shape == (3, H, W)
    for b, im in enumerate(images):
        t = F.normalize(F.to_tensor(Image.fromarray(np.asarray(im))), mean=[0.485, 0.456, 0.406], std=[0.229, 0.224, 0.225])
        h, w = im.shape[:2]
        assert torch.equal(got.tensors[b, :, :h, :w], t)
        assert got.tensors[b, :, h:, :].abs().sum() == 0 and got.tensors[b, :, :, w:].abs().sum() == 0   # zero padding (collate_fn.py:33)
        assert not got.mask[b, :h, :w].any() and got.mask[b, h:, :].all() and got.mask[b, :, w:].all()   # True = padding (:34-37)
