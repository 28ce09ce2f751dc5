"""TEST INFRASTRUCTURE: torch restatement of the counter-based dropout generator of the CUDA path (rb_dropout,
include/reftr_b200.h) and the hooks that make the ORACLE draw the same masks, so that train-mode parity can be checked
mask for mask (the reference draws its masks from torch's generator, which no other implementation can reproduce).

    key  = mix32(lo32(seed) ^ mix32(hi32(seed) + site * 0x9E3779B9))
    word = mix32((row * ((cols + 1) // 2) + col // 2) * 0x9E3779B1 + key)
    keep = ((word >> 16 * (col & 1)) & 0xFFFF) >= round(p * 65536)
"""
import zlib

import torch

M32 = 0xFFFFFFFF


def mix32(x):
    """x: int64 tensor (or python int) holding uint32 values."""
    x = x & M32
    x = x ^ (x >> 16)
    x = (x * 0x21F0AAAD) & M32
    x = x ^ (x >> 15)
    x = (x * 0x735A2D97) & M32
    x = x ^ (x >> 15)
    return x


def site_id(name):
    return zlib.crc32(name.encode()) & 0x7FFFFFFF


def site_key(seed, site):
    seed &= 0xFFFFFFFFFFFFFFFF
    lo, hi = seed & M32, (seed >> 32) & M32
    return mix32(lo ^ mix32((hi + site * 0x9E3779B9) & M32))


def thr_scale(p):
    thr = min(int(p * 65536.0 + 0.5), 65535)
    return thr, 65536.0 / (65536 - thr)


def keep_mask(seed, site, rows, cols, p, device="cpu"):
    """bool [rows, cols]: True where the element is kept."""
    thr, _ = thr_scale(p)
    key = site_key(int(seed), int(site))
    wpr = (cols + 1) // 2
    r = torch.arange(rows, dtype=torch.int64, device=device).view(-1, 1)
    c = torch.arange(cols, dtype=torch.int64, device=device).view(1, -1)
    ctr = (r * wpr + (c >> 1)) & M32
    w = mix32((ctr * 0x9E3779B1 + key) & M32)  # int64 wrap-around keeps the low 32 bits exact
    lane = (w >> ((c & 1) * 16)) & 0xFFFF
    return lane >= thr


def mask_scale(seed, name, rows, cols, p, device="cpu"):
    """fp32 [rows, cols]: 0 where dropped, 1/(1-p') where kept."""
    _, sc = thr_scale(p)
    return keep_mask(seed, site_id(name), rows, cols, p, device).to(torch.float32) * sc


class OracleDropoutHook:
    """oracle.reftr_oracle.DROPOUT_HOOK: drops x with the mask the CUDA path draws for site `tag` under `seed`."""

    def __init__(self, seed, heads=8):
        self.seed, self.heads = int(seed), heads
        self.seen = []

    def __call__(self, x, p, tag, kind):
        self.seen.append(tag)
        if kind == "rows":
            cols = x.shape[-1]
            m = mask_scale(self.seed, tag, x.numel() // cols, cols, p).view(x.shape)
        elif kind == "seq":  # [S, B, d] in the oracle, rows b*S + s in the engine
            S, B, d = x.shape
            m = mask_scale(self.seed, tag, B * S, d, p).view(B, S, d).transpose(0, 1)
        elif kind in ("attn", "attn_bert"):  # [B*h, T, S]
            BH, T, S = x.shape
            if kind == "attn" and T == 1 and S == 1:  # single-query self-attention: the engine drops whole heads in a GEMM epilogue, site tensor [B, h]
                m = mask_scale(self.seed, tag, BH // self.heads, self.heads, p).view(BH, 1, 1)
            else:
                m = mask_scale(self.seed, tag, BH * T, S, p).view(BH, T, S)
        else:
            raise ValueError(kind)
        return x * m.to(x.device)


def hook_hf_bert(bert, hook, ctx):
    """Makes a HuggingFace BertModel draw its dropout masks through `hook` (site names as in reftr_b200/bert.py):
    embeddings.dropout, every BertSelfOutput / BertOutput dropout, and the attention-probability dropout (eager attention).
    `ctx` is oracle.reftr_oracle.DROP_CTX (which BERT invocation is running).  Returns an undo function."""
    import transformers.models.bert.modeling_bert as mb
    from torch import nn

    class _D(nn.Module):
        def __init__(self, p, fmt):
            super().__init__()
            self.p, self.fmt = p, fmt

        def forward(self, x):
            if not self.training or self.p <= 0:
                return x
            return hook(x, self.p, self.fmt.format(ctx["bert"]), "rows")

    bert.config._attn_implementation = "eager"
    bert.embeddings.dropout = _D(bert.embeddings.dropout.p, "bert.{}.emb")
    for li, lay in enumerate(bert.encoder.layer):
        lay.attention.output.dropout = _D(lay.attention.output.dropout.p, "bert.{}.%d.drop1" % li)
        lay.output.dropout = _D(lay.output.dropout.p, "bert.{}.%d.drop2" % li)
    orig = mb.eager_attention_forward

    def hooked(module, query, key, value, attention_mask, scaling=None, dropout=0.0, **kw):
        if scaling is None:
            scaling = query.size(-1) ** -0.5
        w = torch.matmul(query, key.transpose(2, 3)) * scaling
        if attention_mask is not None:
            w = w + attention_mask
        w = torch.softmax(w, dim=-1)
        if module.training and dropout > 0:
            B, H, S, S2 = w.shape
            w = hook(w.reshape(B * H, S, S2), dropout, "bert.%s.%d.attn" % (ctx["bert"], module.layer_idx), "attn_bert").view(B, H, S, S2)
        out = torch.matmul(w, value).transpose(1, 2).contiguous()
        return out, w

    mb.eager_attention_forward = hooked

    def undo():
        mb.eager_attention_forward = orig
    return undo
