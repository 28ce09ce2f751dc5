"""Input pipeline on the GPU.

The reference normalises every image on the CPU (datasets/transforms.py:233-250: ``to_tensor`` + ``Normalize``), pads the batch
on the CPU (``nested_tensor_from_tensor_list``, util/collate_fn.py:24-41) and ships fp32 to the device from a side stream
(``data_prefetcher``, engine_vg.py:234-291).  ``collate_images_u8`` takes the RAW uint8 HWC images instead: they are packed back
to back into one pinned buffer (a quarter of the bytes of the normalised fp32 batch, and no padding), copied with one asynchronous
H2D, and one kernel (``rb_collate_u8``) writes the normalised, padded [B,3,H,W] batch and the padding mask [B,H,W] -- the same
values, bit for bit, as the reference's CPU path.
"""
import torch

from . import ops
from .synthetic import ImageList

IMAGENET_MEAN = (0.485, 0.456, 0.406)
IMAGENET_STD = (0.229, 0.224, 0.225)


class DeviceCollator:
    """Reusable pinned staging buffers + device buffers for one data loader (one instance per process)."""

    def __init__(self, device, mean=IMAGENET_MEAN, std=IMAGENET_STD):
        self.device = torch.device(device)
        self.mean, self.std = tuple(mean), tuple(std)
        self._pinned = self._table = None

    def __call__(self, images, stream=None):
        """images: list of uint8 tensors [h_i, w_i, 3] on the host.  Returns ImageList(tensors fp32 [B,3,H,W], mask bool [B,H,W])
        on the device (the NestedTensor contract of util/misc.py:308-332).  The copies and the kernel are enqueued on ``stream``
        (default: the current stream), nothing synchronises."""
        B = len(images)
        if B == 0 or any(im.dtype != torch.uint8 or im.dim() != 3 or im.shape[2] != 3 or not im.device.type == "cpu" for im in images):
            raise ValueError("collate_images_u8 expects a non-empty list of host uint8 [h, w, 3] tensors")
        H, W = max(im.shape[0] for im in images), max(im.shape[1] for im in images)
        total = sum(im.numel() for im in images)
        if self._pinned is None or self._pinned.numel() < total:
            self._pinned = torch.empty(max(total, 1 << 20), dtype=torch.uint8).pin_memory() if torch.cuda.is_available() else torch.empty(total, dtype=torch.uint8)
        if self._table is None or self._table.shape[0] < B:
            self._table = torch.empty(B, 3, dtype=torch.int64)
            if torch.cuda.is_available():
                self._table = self._table.pin_memory()
        off = 0
        for b, im in enumerate(images):
            n = im.numel()
            self._pinned[off:off + n].copy_(im.reshape(-1))
            self._table[b, 0], self._table[b, 1], self._table[b, 2] = off, im.shape[0], im.shape[1]
            off += n
        ctx = torch.cuda.stream(stream) if stream is not None else _null()
        with ctx:
            packed = self._pinned[:total].to(self.device, non_blocking=True)
            table = self._table[:B].to(self.device, non_blocking=True)
            out = torch.empty(B, 3, H, W, dtype=torch.float32, device=self.device)
            mask = torch.empty(B, H, W, dtype=torch.bool, device=self.device)
            ops.require_device(out)
            ops.collate_u8(packed, table, B, H, W, self.mean, self.std, out, mask)
        return ImageList(out, mask)


class _null:
    def __enter__(self):
        return None

    def __exit__(self, *a):
        return False


def collate_images_u8(images, device, mean=IMAGENET_MEAN, std=IMAGENET_STD, stream=None):
    return DeviceCollator(device, mean, std)(images, stream)
