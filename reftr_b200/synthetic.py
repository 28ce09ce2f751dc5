"""Seeded synthetic weights and inputs.

There is no network for checkpoints or datasets, so every parity test and bench run uses
  * ``synthetic_weights(module, seed)``: re-initialises every parameter/buffer of a module as a pure
    function of (seed, state_dict key, shape).  Because the reference, the oracle and the CUDA module
    share the reference's state_dict layout, the same call gives all three identical weights without
    shipping a 600 MB checkpoint.  It also randomises ``bbox_embed.layers[-1]`` (zero-init in the
    reference, reftr_transformer.py:131-132, which would make ``pred_boxes == 0.5``) and gives
    FrozenBN non-trivial statistics so the BN fold is actually exercised.
  * ``synthetic_samples(...)``: the input dict of ``RefTR.forward`` (reftr_transformer.py:159-248).
"""
import zlib

import torch


class ImageList:
    """Minimal stand-in for util.misc.NestedTensor (util/misc.py:308-332): ``.tensors``, ``.mask``, ``.decompose()``."""

    def __init__(self, tensors, mask):
        self.tensors, self.mask = tensors, mask

    def decompose(self):
        return self.tensors, self.mask

    def to(self, device, non_blocking=False):
        return ImageList(self.tensors.to(device, non_blocking=non_blocking), self.mask.to(device, non_blocking=non_blocking))


def _gen(seed, name):
    g = torch.Generator()
    g.manual_seed((seed * 1000003 + zlib.crc32(name.encode())) & 0x7FFFFFFF)
    return g


@torch.no_grad()
def synthetic_weights(module, seed=0):
    sd = module.state_dict()
    for name, t in sd.items():
        if not t.is_floating_point():
            continue
        g = _gen(seed, name)
        shape = tuple(t.shape)
        leaf = name.rsplit(".", 1)[-1]
        is_norm = any(k in name for k in (".bn", "downsample.1", "norm", "LayerNorm", "gn")) or \
            (t.dim() == 1 and leaf == "weight")
        if leaf == "running_var":
            v = 0.5 + torch.rand(shape, generator=g)
        elif leaf == "running_mean":
            v = 0.1 * torch.randn(shape, generator=g)
        elif name.endswith("bn3.weight"):
            # last BatchNorm of a bottleneck: a small gain keeps the residual stream bounded through 16 / 33 blocks, as in any
            # trained ResNet (C5 activations of O(10..100)); gain 1 with He-initialised convolutions doubles the variance per
            # block (ResNet-101: 1e8 at C5), which no 16-bit activation format represents
            v = 0.3 + 0.05 * torch.randn(shape, generator=g)
        elif t.dim() == 1 and leaf == "weight" and is_norm:
            v = 1.0 + 0.1 * torch.randn(shape, generator=g)
        elif t.dim() <= 1 or leaf in ("bias", "in_proj_bias"):
            v = 0.02 * torch.randn(shape, generator=g)
        elif any(k in name for k in ("embeddings.", "query_embed", "level_embed")):
            v = 0.5 * torch.randn(shape, generator=g)
        else:
            fan_in = t[0].numel()
            # He-like gain keeps activations O(1) through the ReLU stacks of the backbone
            gain = 2.0 if ("conv" in name or "downsample.0" in name or "lay" in name or "adapter" in name) else 1.0
            v = torch.randn(shape, generator=g) * (gain / fan_in) ** 0.5
        t.copy_(v.to(t.dtype))
    return module


def synthetic_samples(B, H, W, L, n_valid=None, seed=1, pad_frac=0.0, n_ph=0, device="cpu", vocab=30000):
    """Inputs of RefTR.forward.  ``pad_frac`` > 0 masks the right part of every other image (padding)."""
    g = torch.Generator().manual_seed(seed)
    nv = L if n_valid is None else n_valid
    assert 3 <= nv <= L
    img = torch.randn(B, 3, H, W, generator=g)
    mask = torch.zeros(B, H, W, dtype=torch.bool)
    if pad_frac > 0:
        w0 = int(W * (1 - pad_frac))
        mask[1::2, :, w0:] = True
        img[1::2, :, :, w0:] = 0
    sent = torch.randint(1000, vocab, (B, L), generator=g)
    sent[:, 0] = 101
    sent[:, nv - 1] = 102
    sent[:, nv:] = 0
    smask = torch.zeros(B, L, dtype=torch.long)
    smask[:, :nv] = 1
    s = {"img": ImageList(img.to(device), mask.to(device)), "sentence": sent.to(device), "sentence_mask": smask.to(device)}
    if n_ph > 0:
        Lp = 22
        ph = torch.zeros(B, n_ph, Lp, dtype=torch.long)
        ph[:, :, 0] = 101
        ph[:, :, 1:4] = torch.randint(1000, vocab, (B, n_ph, 3), generator=g)
        ph[:, :, 4] = 102
        pm = torch.zeros(B, n_ph, Lp, dtype=torch.long)
        pm[:, :, :5] = 1
        if n_ph > 1:  # last phrase of every sample is an empty "[CLS] [SEP]" pad phrase (refer_dataset padding)
            ph[:, -1, 1] = 102
            ph[:, -1, 2:] = 0
            pm[:, -1, 2:] = 0
        s["phrase"] = ph.to(device)
        s["phrase_mask"] = pm.to(device)
        s["phrase_pos_l"] = torch.ones(B, n_ph, dtype=torch.long, device=device)
        s["phrase_pos_r"] = torch.full((B, n_ph), 4, dtype=torch.long, device=device)
    return s


def synthetic_targets(B, n_ph=1, seed=2, device="cpu"):
    """cxcywh boxes, centres in [0.25,0.75], sizes in [0.1,0.4]."""
    g = torch.Generator().manual_seed(seed)
    c = 0.25 + 0.5 * torch.rand(B, n_ph, 2, generator=g)
    wh = 0.1 + 0.3 * torch.rand(B, n_ph, 2, generator=g)
    return torch.cat([c, wh], -1).to(device)


def synthetic_mask_targets(B, H, W, seed=4, device="cpu"):
    """bool [B, 1, H, W] random rectangles (targets["masks"] of the segmentation criterion)."""
    g = torch.Generator().manual_seed(seed)
    m = torch.zeros(B, 1, H, W, dtype=torch.bool)
    for b in range(B):
        y0, x0 = int(torch.randint(0, H // 2, (1,), generator=g)), int(torch.randint(0, W // 2, (1,), generator=g))
        hh, ww = int(torch.randint(H // 8, H // 2, (1,), generator=g)), int(torch.randint(W // 8, W // 2, (1,), generator=g))
        m[b, 0, y0:y0 + hh, x0:x0 + ww] = True
    return m.to(device)


def criterion_targets(boxes, phrase_mask=None, masks=None, sizes=None):
    """The reference's target format (list of per-sample dicts, refer_dataset.py:186-194 / criterion.py:113-130) from the dense
    synthetic targets: ``boxes`` [B, n_ph, 4]; ``phrase_mask`` bool [B, n_ph(*k)] selects the valid phrases of a multi-phrase batch."""
    B, n_ph = boxes.shape[:2]
    out = []
    for b in range(B):
        if phrase_mask is not None:
            pm = phrase_mask[b].view(n_ph, -1)[:, 0]
            bx = boxes[b][pm]
        else:
            bx = boxes[b]
        t = {"boxes": bx, "labels": torch.zeros(bx.shape[0], dtype=torch.long, device=boxes.device)}
        if masks is not None:
            t["masks"] = masks[b]
        if sizes is not None:
            t["orig_size"] = t["size"] = torch.tensor(sizes, device=boxes.device)
        out.append(t)
    return out
