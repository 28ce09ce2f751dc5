"""GPU input pipeline (reftr_b200/data.py) against the reference's CPU path restated in torch (to_tensor + Normalize,
datasets/transforms.py:233-250; nested_tensor_from_tensor_list, util/collate_fn.py:24-41): bit-exact, ragged sizes included."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("sizes", [[(640, 640)] * 4, [(480, 640), (640, 427), (333, 500), (17, 9), (640, 640)], [(1, 1)], [(224, 224), (200, 300)]])
def test_collate_u8_matches_reference_cpu_path(sizes):
    from data_ref import reference_collate
    from reftr_b200.data import DeviceCollator
    g = torch.Generator().manual_seed(len(sizes))
    images = [torch.randint(0, 256, (h, w, 3), dtype=torch.uint8, generator=g) for h, w in sizes]
    ref = reference_collate(images)
    col = DeviceCollator("cuda")
    for _ in range(2):  # second call reuses the staging buffers
        got = col(images)
        torch.cuda.synchronize()
        assert got.tensors.shape == ref.tensors.shape and got.mask.dtype == torch.bool
        assert torch.equal(got.mask.cpu(), ref.mask)
        assert torch.equal(got.tensors.cpu(), ref.tensors)  # same fp32 operations in the same order: bit-exact
    side = torch.cuda.Stream()
    got = col(images, stream=side)
    side.synchronize()
    assert torch.equal(got.tensors.cpu(), ref.tensors)


def test_pack_then_upload_matches_reference_cpu_path():
    """The split form: DeviceCollator.pack in the loader worker (pinned host memory), DeviceCollator.upload on the copy stream."""
    from data_ref import reference_collate
    from reftr_b200.data import DeviceCollator
    g = torch.Generator().manual_seed(11)
    images = [torch.randint(0, 256, (h, w, 3), dtype=torch.uint8, generator=g) for h, w in [(480, 640), (640, 427), (64, 64)]]
    ref = reference_collate(images)
    packed = DeviceCollator.pack(images)
    assert packed.buf.is_pinned() and packed.nbytes() == sum(im.numel() for im in images) + 3 * 3 * 8
    side = torch.cuda.Stream()
    col = DeviceCollator("cuda")
    for _ in range(2):
        got = col.upload(packed, stream=side)
        torch.cuda.current_stream().wait_event(got.ready)
        assert torch.equal(got.tensors.cpu(), ref.tensors) and torch.equal(got.mask.cpu(), ref.mask)


def test_collator_staging_is_not_overwritten_while_a_copy_is_in_flight():
    """The host runs ahead of the GPU: batch i+1 (and i+2, which reuses batch i's pinned slot) is packed while the stream is still
    busy with earlier work.  Every batch must arrive intact."""
    from data_ref import reference_collate
    from reftr_b200.data import DeviceCollator
    g = torch.Generator().manual_seed(3)
    batches = [[torch.randint(0, 256, (320, 320, 3), dtype=torch.uint8, generator=g) for _ in range(4)] for _ in range(6)]
    refs = [reference_collate(b) for b in batches]
    col = DeviceCollator("cuda")
    side = torch.cuda.Stream()
    busy = torch.empty(64 << 20, device="cuda")
    outs = []
    for b in batches:
        with torch.cuda.stream(side):
            for _ in range(20):
                busy.add_(1.0)          # keep the copy stream backed up so the host gets ahead of the DMA
        outs.append(col(b, stream=side))
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    for got, ref in zip(outs, refs):
        assert torch.equal(got.tensors.cpu(), ref.tensors) and torch.equal(got.mask.cpu(), ref.mask)


def test_collated_batch_feeds_the_model():
    from oracle.cases import CASES
    from reftr_b200.data import DeviceCollator
    from reftr_b200.synthetic import synthetic_samples
    from util_build import build_candidate
    case = CASES["pad_box"]
    g = torch.Generator().manual_seed(0)
    images = [torch.randint(0, 256, (192, 256, 3), dtype=torch.uint8, generator=g), torch.randint(0, 256, (192, 200, 3), dtype=torch.uint8, generator=g)]
    s = synthetic_samples(**case["inputs"], device="cuda")
    s["img"] = DeviceCollator("cuda")(images)
    model = build_candidate(case, device="cuda")
    out = model(s)
    assert torch.isfinite(out["pred_boxes"]).all() and out["pred_boxes"].shape[0] == 2
