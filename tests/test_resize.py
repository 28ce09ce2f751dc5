"""On-device resize of the input pipeline (datasets/transforms.py:81-111, RandomResize :196-204).  The reference
resizes a PIL image with torchvision's F.resize = Pillow's bilinear ImagingResample.  CPU: the restated coefficient computation
(reftr_b200/data.py:pil_bilinear_coeffs) + the two-pass uint8 arithmetic (tests/emu_ops.py:resize_u8) reproduce PIL bit for bit, and
the output-size rule reproduces the reference's.  GPU: rb_resize_u8 and DeviceCollator(resize=...) against PIL + the reference collate."""
import os
import sys

import pytest
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
PIL = pytest.importorskip("PIL")
from PIL import Image  # noqa: E402

from reftr_b200.data import pil_bilinear_coeffs, resized_size  # noqa: E402

SIZES = [(480, 640, 480, 640), (375, 500, 480, 640), (500, 375, 640, 480), (1080, 1920, 360, 640), (333, 517, 640, 413), (64, 48, 17, 200),
         (427, 640, 427, 640), (768, 1024, 480, 640), (33, 57, 640, 1105)]


def _pil(img, oh, ow):
    import numpy as np
    return torch.from_numpy(np.asarray(Image.fromarray(img.numpy()).resize((ow, oh), Image.BILINEAR)).copy())


@pytest.mark.parametrize("h,w,oh,ow", SIZES)
def test_restated_resample_is_bit_exact_with_pillow(h, w, oh, ow):
    import emu_ops
    g = torch.Generator().manual_seed(h * 7 + w)
    img = torch.randint(0, 256, (h, w, 3), generator=g, dtype=torch.uint8)
    dst = torch.empty(oh * ow * 3, dtype=torch.uint8)
    emu_ops.resize_u8(img.reshape(-1), h, w, dst, oh, ow, pil_bilinear_coeffs(w, ow) if ow != w else None,
                      pil_bilinear_coeffs(h, oh) if oh != h else None, None)
    assert torch.equal(dst.view(oh, ow, 3), _pil(img, oh, ow))


def test_coefficients_sum_to_one_in_fixed_point():
    for n_in, n_out in [(640, 480), (480, 640), (1920, 640), (57, 1105)]:
        b, k = pil_bilinear_coeffs(n_in, n_out)
        assert b.shape == (n_out, 2) and k.shape[0] == n_out
        assert (b[:, 0] >= 0).all() and (b[:, 0] + b[:, 1] <= n_in).all()
        assert ((k.sum(1) - (1 << 22)).abs() <= k.shape[1]).all()       # each weight is rounded separately
        assert (k >= 0).all()


def test_output_size_rule_matches_the_reference():
    """get_size_with_aspect_ratio (datasets/transforms.py:84-101): checked against the reference's own function when its sources are
    present (this container), and against values worked out from it otherwise."""
    cases = [((640, 480), 640, 640), ((480, 640), 640, 640), ((500, 375), 640, 640), ((1920, 1080), 640, 640), ((333, 517), 800, 1333),
             ((640, 640), 640, 640), ((1000, 200), 640, 640), ((375, 500), 512, 600)]
    expect = [(480, 640), (640, 480), (480, 640), (360, 640), (1242, 800), (640, 640), (128, 640), (600, 450)]
    for ((w, h), size, max_size), e in zip(cases, expect):
        assert resized_size(w, h, size, max_size) == e, ((w, h), size, max_size)
    ref_dir = "/root/reference"
    if os.path.isdir(ref_dir):
        import importlib.util
        import types
        src = open(os.path.join(ref_dir, "datasets", "transforms.py")).read()
        start = src.index("    def get_size_with_aspect_ratio(image_size, size, max_size=None):")
        end = src.index("    def get_size(image_size, size, max_size=None):")
        ns = {}
        exec("\n".join(l[4:] for l in src[start:end].split("\n")), ns)   # the function's own text, de-indented
        for (w, h), size, max_size in cases + [((w, h), 640, 640) for w in range(300, 900, 37) for h in range(300, 900, 41)]:
            assert resized_size(w, h, size, max_size) == tuple(ns["get_size_with_aspect_ratio"]((w, h), size, max_size))


@pytest.mark.gpu
@pytest.mark.parametrize("h,w,oh,ow", SIZES)
def test_rb_resize_u8_is_bit_exact_with_pillow(h, w, oh, ow):
    from reftr_b200 import ops
    g = torch.Generator().manual_seed(h * 7 + w)
    img = torch.randint(0, 256, (h, w, 3), generator=g, dtype=torch.uint8)
    src = img.reshape(-1).cuda()
    dst = torch.full((oh * ow * 3,), 7, dtype=torch.uint8, device="cuda")
    tmp = torch.empty(h * ow * 3, dtype=torch.uint8, device="cuda")
    th = tuple(t.cuda() for t in pil_bilinear_coeffs(w, ow)) if ow != w else None
    tv = tuple(t.cuda() for t in pil_bilinear_coeffs(h, oh)) if oh != h else None
    ops.resize_u8(src, h, w, dst, oh, ow, th, tv, tmp)
    assert torch.equal(dst.view(oh, ow, 3).cpu(), _pil(img, oh, ow))


@pytest.mark.gpu
def test_device_collator_with_resize_matches_the_reference_pipeline():
    """RandomResize([640], max_size=640) + ToTensor + Normalize + nested_tensor_from_tensor_list of the reference on the CPU (PIL,
    tests/data_ref.py) against ONE upload of the original pixels + resize + collate on the device: bit-exact, ragged sizes."""
    from data_ref import reference_collate
    from reftr_b200.data import DeviceCollator
    g = torch.Generator().manual_seed(5)
    shapes = [(375, 500), (500, 375), (480, 640), (1080, 1920), (333, 517), (640, 640)]
    images = [torch.randint(0, 256, (h, w, 3), generator=g, dtype=torch.uint8) for h, w in shapes]
    resized = [_pil(im, *resized_size(im.shape[1], im.shape[0], 640, 640)) for im in images]
    ref = reference_collate(resized)
    col = DeviceCollator("cuda", resize=(640, 640))
    out = col.upload(DeviceCollator.pack(images))
    torch.cuda.synchronize()
    assert out.tensors.shape == ref.tensors.shape
    assert torch.equal(out.tensors.cpu(), ref.tensors) and torch.equal(out.mask.cpu(), ref.mask)
