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


def resized_size(w, h, size, max_size=None):
    """(oh, ow) of datasets/transforms.py:84-101 (``get_size_with_aspect_ratio``): the SHORTER side becomes ``size`` unless the longer
    one would exceed ``max_size``; int() truncation and round() exactly as there."""
    if max_size is not None:
        mn, mx = float(min(w, h)), float(max(w, h))
        if mx / mn * size > max_size:
            size = int(round(max_size * mn / mx))
    if (w <= h and w == size) or (h <= w and h == size):
        return h, w
    if w < h:
        return int(size * h / w), size
    return size, int(size * w / h)


_PIL_PREC = 22  # Pillow: PRECISION_BITS = 32 - 8 - 2


def pil_bilinear_coeffs(in_size, out_size):
    """Pillow's ``precompute_coeffs`` + ``normalize_coeffs_8bpc`` (src/libImaging/Resample.c) for the bilinear filter (support 1) over
    the whole axis: (bounds int32 [out, 2] = (first input index, count), kk int32 [out, ksize]).  The double arithmetic, the (int)
    truncations and the +-0.5 roundings are Pillow's, so the two-pass uint8 resample built on them (``rb_resize_u8``) is bit-exact
    with ``PIL.Image.resize(..., BILINEAR)``, i.e. with torchvision's ``F.resize`` of a PIL image (datasets/transforms.py:111)."""
    import math
    scale = filterscale = in_size / out_size
    if filterscale < 1.0:
        filterscale = 1.0
    support = 1.0 * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    ss = 1.0 / filterscale
    bounds = torch.zeros(out_size, 2, dtype=torch.int32)
    kk = torch.zeros(out_size, ksize, dtype=torch.int32)
    one = float(1 << _PIL_PREC)
    for xx in range(out_size):
        center = 0.0 + (xx + 0.5) * scale
        xmin = int(center - support + 0.5)
        if xmin < 0:
            xmin = 0
        xmax = int(center + support + 0.5)
        if xmax > in_size:
            xmax = in_size
        xmax -= xmin
        w, ww = [], 0.0
        for x in range(xmax):
            a = (x + xmin - center + 0.5) * ss
            if a < 0.0:
                a = -a
            v = 1.0 - a if a < 1.0 else 0.0
            w.append(v)
            ww += v
        for x in range(xmax):
            v = w[x] / ww if ww != 0.0 else w[x]
            kk[xx, x] = int(-0.5 + v * one) if v < 0 else int(0.5 + v * one)
        bounds[xx, 0], bounds[xx, 1] = xmin, xmax
    return bounds, kk


class DeviceCollator:
    """Reusable pinned staging buffers + device buffers for one data loader (one instance per process)."""

    N_SLOTS = 2  # pinned staging slots: batch i+1 is packed on the host while the DMA of batch i may still be in flight

    def __init__(self, device, mean=IMAGENET_MEAN, std=IMAGENET_STD, resize=None):
        """``resize`` = (size, max_size): the images are resized ON THE DEVICE after the H2D copy of their raw bytes, like
        ``T.RandomResize([size], max_size=max_size)`` of the reference's datasets (refer_multiphrase.py:42, transforms.py:81-111;
        Pillow's bilinear resample, bit-exact) -- the host ships the ORIGINAL pixels and never resizes.  ``resized_size`` gives the
        output size of an image (the caller scales its boxes by the same ratios, transforms.py:116-122)."""
        self.device = torch.device(device)
        self.mean, self.std = tuple(mean), tuple(std)
        self.resize = tuple(resize) if resize is not None else None
        self._coef = {}  # (in, out) -> device (bounds, kk)
        self._slots = [dict(pinned=None, table=None, event=None) for _ in range(self.N_SLOTS)]
        self._next = 0

    @staticmethod
    def pack(images):
        """Host side of the collation, for the data-loader worker: the raw uint8 HWC images packed back to back into ONE pinned
        buffer plus the (offset, h, w) table -- what ``upload`` ships.  Returns a ``PackedBatch``."""
        B = len(images)
        if B == 0 or any(im.dtype != torch.uint8 or im.dim() != 3 or im.shape[2] != 3 or not im.device.type == "cpu" for im in images):
            raise ValueError("DeviceCollator.pack expects a non-empty list of host uint8 [h, w, 3] tensors")
        total = sum(im.numel() for im in images)
        cuda = torch.cuda.is_available()
        buf = torch.empty(total, dtype=torch.uint8)
        tab = torch.empty(B, 3, dtype=torch.int64)
        if cuda:
            buf, tab = buf.pin_memory(), tab.pin_memory()
        off = 0
        for b, im in enumerate(images):
            n = im.numel()
            buf[off:off + n].copy_(im.reshape(-1))
            tab[b, 0], tab[b, 1], tab[b, 2] = off, im.shape[0], im.shape[1]
            off += n
        return PackedBatch(buf, tab, max(im.shape[0] for im in images), max(im.shape[1] for im in images))

    def upload(self, packed, stream=None, consumer_stream=None):
        """Device side: one H2D of the packed bytes (+ the table) and rb_collate_u8, on ``stream``.  ``packed`` (a ``PackedBatch``) owns
        its pinned memory, so nothing here is reused across calls; the caller keeps ``packed`` alive until the returned
        ``ImageList.ready`` event has passed (the bench keeps one batch for the whole run)."""
        cuda = torch.cuda.is_available()
        B = packed.table.shape[0]
        consumer = consumer_stream if consumer_stream is not None else (torch.cuda.current_stream(self.device) if cuda else None)
        ctx = torch.cuda.stream(stream) if stream is not None else _null()
        with ctx:
            dbuf = packed.buf.to(self.device, non_blocking=True)
            H, W = packed.H, packed.W
            if self.resize is not None:
                dbuf, table, H, W = self._resize_packed(dbuf, packed.table)
                dtab = table.to(self.device, non_blocking=True)
            else:
                dtab = packed.table.to(self.device, non_blocking=True)
            out = torch.empty(B, 3, H, W, dtype=torch.float32, device=self.device)
            mask = torch.empty(B, H, W, dtype=torch.bool, device=self.device)
            ops.require_device(out)
            ops.collate_u8(dbuf, dtab, B, H, W, self.mean, self.std, out, mask)
            ready = None
            if cuda:
                ready = torch.cuda.Event()
                ready.record()
        if cuda and stream is not None and consumer is not None and consumer != stream:
            for t in (out, mask):
                t.record_stream(consumer)
        res = ImageList(out, mask)
        res.ready = ready
        return res

    def _tables(self, n_in, n_out):
        key = (int(n_in), int(n_out))
        t = self._coef.get(key)
        if t is None:
            b, k = pil_bilinear_coeffs(*key)
            t = self._coef[key] = (b.to(self.device), k.to(self.device))
        return t

    def _resize_packed(self, dbuf, table):
        """dbuf: the raw images back to back on the device, table: their host (offset, h, w) rows.  Returns the resized images packed
        the same way + their host table (pinned when CUDA is there) + the batch's padded size."""
        size, max_size = self.resize
        rows = [(int(o), int(h), int(w)) + resized_size(int(w), int(h), size, max_size) for o, h, w in table.tolist()]
        total = sum(oh * ow * 3 for _, _, _, oh, ow in rows)
        dst = torch.empty(total, dtype=torch.uint8, device=self.device)
        tmp = torch.empty(max(h * ow * 3 for _, h, _, _, ow in rows), dtype=torch.uint8, device=self.device)
        new = torch.empty(len(rows), 3, dtype=torch.int64)
        if torch.cuda.is_available():
            new = new.pin_memory()
        off = 0
        for b, (o, h, w, oh, ow) in enumerate(rows):
            ops.resize_u8(dbuf[o:o + h * w * 3], h, w, dst[off:off + oh * ow * 3], oh, ow, self._tables(w, ow) if ow != w else None,
                          self._tables(h, oh) if oh != h else None, tmp)
            new[b, 0], new[b, 1], new[b, 2] = off, oh, ow
            off += oh * ow * 3
        return dst, new, max(r[3] for r in rows), max(r[4] for r in rows)

    def __call__(self, images, stream=None, consumer_stream=None):
        """images: list of uint8 tensors [h_i, w_i, 3] on the host.  Returns ImageList(tensors fp32 [B,3,H,W], mask bool [B,H,W])
        on the device (the NestedTensor contract of util/misc.py:308-332).  The copies and the kernel are enqueued on ``stream``
        (default: the current stream).

        Hazards handled here (the host runs far ahead of the GPU under CUDA-graph replay): (1) a pinned staging slot is rewritten
        only after the H2D copies that last read it have EXECUTED -- each slot carries a CUDA event recorded after its copies, and
        ``event.synchronize()`` is called before the slot is reused (with two slots that wait is normally already over); (2) the
        returned tensors are allocated on ``stream`` but consumed elsewhere: ``record_stream`` is called for ``consumer_stream``
        (default: the stream that is current when the collator is called), as engine_vg.data_prefetcher does (engine_vg.py:271-283);
        the caller still has to make the consumer wait for ``stream`` (``wait_stream`` / the returned ``ImageList.ready`` event)."""
        B = len(images)
        if B == 0 or any(im.dtype != torch.uint8 or im.dim() != 3 or im.shape[2] != 3 or not im.device.type == "cpu" for im in images):
            raise ValueError("collate_images_u8 expects a non-empty list of host uint8 [h, w, 3] tensors")
        H, W = max(im.shape[0] for im in images), max(im.shape[1] for im in images)
        total = sum(im.numel() for im in images)
        cuda = torch.cuda.is_available()
        slot = self._slots[self._next]
        self._next = (self._next + 1) % self.N_SLOTS
        if slot["event"] is not None:
            slot["event"].synchronize()  # the DMA that last read this slot has finished
        if slot["pinned"] is None or slot["pinned"].numel() < total:
            slot["pinned"] = torch.empty(max(total, 1 << 20), dtype=torch.uint8).pin_memory() if cuda else torch.empty(total, dtype=torch.uint8)
        if slot["table"] is None or slot["table"].shape[0] < B:
            slot["table"] = torch.empty(B, 3, dtype=torch.int64).pin_memory() if cuda else torch.empty(B, 3, dtype=torch.int64)
        pinned, tab = slot["pinned"], slot["table"]
        off = 0
        for b, im in enumerate(images):
            n = im.numel()
            pinned[off:off + n].copy_(im.reshape(-1))
            tab[b, 0], tab[b, 1], tab[b, 2] = off, im.shape[0], im.shape[1]
            off += n
        consumer = consumer_stream if consumer_stream is not None else (torch.cuda.current_stream(self.device) if cuda else None)
        ctx = torch.cuda.stream(stream) if stream is not None else _null()
        with ctx:
            packed = pinned[:total].to(self.device, non_blocking=True)
            table = tab[:B].to(self.device, non_blocking=True)
            if cuda:
                slot["event"] = torch.cuda.Event()
                slot["event"].record()  # on `stream`: after both copies
            out = torch.empty(B, 3, H, W, dtype=torch.float32, device=self.device)
            mask = torch.empty(B, H, W, dtype=torch.bool, device=self.device)
            ops.require_device(out)
            ops.collate_u8(packed, table, B, H, W, self.mean, self.std, out, mask)
            ready = None
            if cuda:
                ready = torch.cuda.Event()
                ready.record()
        if cuda and stream is not None and consumer is not None and consumer != stream:
            for t in (out, mask):
                t.record_stream(consumer)
        res = ImageList(out, mask)
        res.ready = ready  # consumer: torch.cuda.current_stream().wait_event(res.ready)
        return res


class PackedBatch:
    """Pinned host buffers of one collated batch: ``buf`` uint8 (all images back to back), ``table`` int64 [B, 3] = (offset, h, w)."""

    def __init__(self, buf, table, H, W):
        self.buf, self.table, self.H, self.W = buf, table, int(H), int(W)

    def nbytes(self):
        return self.buf.numel() + self.table.numel() * 8


class _null:
    def __enter__(self):
        return None

    def __exit__(self, *a):
        return False


def collate_images_u8(images, device, mean=IMAGENET_MEAN, std=IMAGENET_STD, stream=None):
    return DeviceCollator(device, mean, std)(images, stream)
