"""TEST INFRASTRUCTURE: what the reference computes on the CPU for raw uint8 images -- torchvision ``to_tensor`` + ``Normalize``
(datasets/transforms.py:233-250) and ``nested_tensor_from_tensor_list`` (util/collate_fn.py:24-41) -- restated in torch; the checker
of reftr_b200/data.py (tests/test_data_gpu.py).  Never imported by the product package."""
import torch

from reftr_b200.data import IMAGENET_MEAN, IMAGENET_STD
from reftr_b200.synthetic import ImageList


def reference_collate(images, mean=IMAGENET_MEAN, std=IMAGENET_STD):
    B = len(images)
    H, W = max(im.shape[0] for im in images), max(im.shape[1] for im in images)
    m, s = torch.tensor(mean).view(3, 1, 1), torch.tensor(std).view(3, 1, 1)
    out = torch.zeros(B, 3, H, W)
    mask = torch.ones(B, H, W, dtype=torch.bool)
    for b, im in enumerate(images):
        t = im.permute(2, 0, 1).to(torch.float32).div(255)
        t = t.sub(m).div(s)
        out[b, :, :im.shape[0], :im.shape[1]] = t
        mask[b, :im.shape[0], :im.shape[1]] = False
    return ImageList(out, mask)
