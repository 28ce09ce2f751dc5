"""BERT-base (HuggingFace ``BertModel``) forward / backward on the C-ABI kernels -- SURVEY.md 8(f) row N1.

The reference calls ``self.lang_backbone(sentence, token_type_ids=None, attention_mask=sentence_mask)`` (reftr_transformer.py:200)
and, for multi-phrase inputs, once more on the phrases (:217), taking ``[0]`` (sequence output) and ``[1]`` (pooler output).
The HF module stays the PARAMETER CONTAINER (state_dict keys ``lang_backbone.*`` unchanged); this class restates its math:
BertEmbeddings (word + position + token-type-0, LayerNorm eps 1e-12) -> N x [self-attention (12 heads x 64, additive key mask),
dense + residual + LayerNorm, dense + exact GELU, dense + residual + LayerNorm] -> pooler (dense on token 0, tanh).
Residual stream and LayerNorm statistics are fp32; GEMM operands bf16 (as in the rest of the hot path).
"""
import torch

from . import ops
from .pack import PackedLinear, PackedStack, lin_taps


class _L:
    pass


class BertEngine:
    @staticmethod
    def eligible(bert):
        cfg = getattr(bert, "config", None)
        if cfg is None or type(bert).__name__ != "BertModel":
            return False
        ok = cfg.hidden_act == "gelu" and cfg.hidden_size in (768, 1024) and cfg.hidden_size // cfg.num_attention_heads == 64
        ok = ok and getattr(cfg, "position_embedding_type", None) in (None, "absolute") and getattr(bert, "pooler", None) is not None
        flags = {p.requires_grad for p in bert.parameters()}
        return ok and len(flags) == 1  # all trainable or all frozen

    def __init__(self, eng, bert):
        self.eng, self.bert = eng, bert
        cfg = bert.config
        self.D, self.H, self.FF, self.eps = cfg.hidden_size, cfg.num_attention_heads, cfg.intermediate_size, cfg.layer_norm_eps
        self.trainable = all(p.requires_grad for p in bert.parameters())
        self.layers = []
        self.packs = []
        for lay in bert.encoder.layer:
            a = lay.attention
            l = _L()
            l.mod = lay
            from .engine import HILO
            hl = "bert" in HILO  # forward GEMMs multiply by (hi | residual) weight pairs, see engine.py ("bert_qkv" .. : one layer type)
            l.qkv = PackedStack([a.self.query.weight, a.self.key.weight, a.self.value.weight],
                                [a.self.query.bias, a.self.key.bias, a.self.value.bias], 0, self.D, hilo=hl or "bert_qkv" in HILO)
            l.o = PackedLinear(a.output.dense.weight, a.output.dense.bias, hilo=hl or "bert_o" in HILO)
            l.i = PackedLinear(lay.intermediate.dense.weight, lay.intermediate.dense.bias, hilo=hl or "bert_ffn1" in HILO)
            l.o2 = PackedLinear(lay.output.dense.weight, lay.output.dense.bias, hilo=hl or "bert_ffn2" in HILO)
            self.layers.append(l)
            self.packs += [l.qkv, l.o, l.i, l.o2]
        self.p_hidden, self.p_attn = float(cfg.hidden_dropout_prob), float(cfg.attention_probs_dropout_prob)
        self.pool = PackedLinear(bert.pooler.dense.weight, bert.pooler.dense.bias)
        self.packs.append(self.pool)
        self.saved = {}
        # BERT's chains run BESIDE the conv backbone (engine._branch): what they cost the step is the SM time their launches hold, not
        # their latency -- a 144-CTA launch that lives 8 us for 1 us of tensor-core work (M = 320 rows) blocks the whole chip.  The
        # backward's GEMMs are therefore launched narrow (rb_gemm_args.sm_limit): input gradients on <= 72 SMs, weight gradients on <= 36 (measured: tools/gpu_bertsm.sh).
        import os
        self.lim_fwd = int(os.environ.get("REFTR_B200_BERT_SMS_FWD", "0"))
        self.lim_bwd = int(os.environ.get("REFTR_B200_BERT_SMS_BWD", "72"))
        self.lim_wgrad = int(os.environ.get("REFTR_B200_BERT_SMS_WGRAD", "36"))

    # ------------------------------------------------------------------------------------------------------------
    def forward(self, tag, ids, mask_u8, Bn, L):
        with ops.sm_limit_scope(self.lim_fwd):
            return self._forward(tag, ids, mask_u8, Bn, L)

    def _forward(self, tag, ids, mask_u8, Bn, L):
        """ids int64 [Bn, L], mask_u8 [Bn, L] (1 = padding key).  Returns (seq fp32 [Bn*L, D], seq bf16, pooled fp32 [Bn, D])."""
        eng, ws, bert = self.eng, self.eng.ws, self.bert
        D, H, FF = self.D, self.H, self.FF
        if L > bert.config.max_position_embeddings:
            raise ValueError(f"{L} tokens exceed BERT's {bert.config.max_position_embeddings} positions")
        rows = Bn * L
        f32 = torch.float32
        emb = bert.embeddings
        e32 = ws.get(f"bert.{tag}.emb", [rows, D], f32)
        ops.bert_embed_fwd(ids, L, emb.word_embeddings.weight, emb.position_embeddings.weight, emb.token_type_embeddings.weight[0], e32)
        x32, xb = ws.get(f"bert.{tag}.x0", [rows, D], f32), ws.get(f"bert.{tag}.x0b", [rows, D])
        me, re_ = ws.get(f"bert.{tag}.me", [rows], f32), ws.get(f"bert.{tag}.re", [rows], f32)
        # train mode: BertEmbeddings.dropout, attention_probs dropout, BertSelfOutput / BertOutput dropouts (HF modeling_bert)
        ops.ln_wide_fwd(e32, emb.LayerNorm.weight, emb.LayerNorm.bias, rows, y32=x32, yb=xb, mean=me, rstd=re_, eps=self.eps,
                        drop=eng.drop(f"bert.{tag}.emb", self.p_hidden))
        scale = 64 ** -0.5
        per_layer = []
        for li, l in enumerate(self.layers):
            k = f"bert.{tag}.{li}"
            lay = l.mod
            qkv = ws.get(k + ".qkv", [rows, 3 * D])
            wq, tq = lin_taps(l.qkv)
            ops.gemm(xb, wq, rows, 3 * D, D, taps=tq, bias=l.qkv.bias, out=qkv)
            ctx = ws.get(k + ".ctx", [rows, D])
            P = ws.get(k + ".P", [Bn, H, L, L], f32)
            ops.attn_small_fwd(qkv[:, :D], qkv[:, D:2 * D], qkv[:, 2 * D:], mask_u8, ctx, P, Bn, H, L, scale, drop=eng.drop(k + ".attn", self.p_attn))
            y1 = ws.get(k + ".y1", [rows, D], f32)
            wo, to = lin_taps(l.o)
            ops.gemm(ctx, wo, rows, D, D, taps=to, bias=l.o.bias, res32=x32, out32=y1, drop=eng.drop(k + ".drop1", self.p_hidden))
            x1, x1b = ws.get(k + ".x1", [rows, D], f32), ws.get(k + ".x1b", [rows, D])
            m1, r1 = ws.get(k + ".m1", [rows], f32), ws.get(k + ".r1", [rows], f32)
            ln1 = lay.attention.output.LayerNorm
            ops.ln_wide_fwd(y1, ln1.weight, ln1.bias, rows, y32=x1, yb=x1b, mean=m1, rstd=r1, eps=self.eps)
            hpre, h = ws.get(k + ".hpre", [rows, FF]), ws.get(k + ".h", [rows, FF])
            wi, ti = lin_taps(l.i)
            ops.gemm(x1b, wi, rows, FF, D, taps=ti, bias=l.i.bias, out=hpre)
            ops.gelu_fwd(hpre, h)
            y2 = ws.get(k + ".y2", [rows, D], f32)
            wo2, to2 = lin_taps(l.o2)
            ops.gemm(h, wo2, rows, D, FF, taps=to2, bias=l.o2.bias, res32=x1, out32=y2, drop=eng.drop(k + ".drop2", self.p_hidden))
            xo, xob = ws.get(k + ".xo", [rows, D], f32), ws.get(k + ".xob", [rows, D])
            m2, r2 = ws.get(k + ".m2", [rows], f32), ws.get(k + ".r2", [rows], f32)
            ln2 = lay.output.LayerNorm
            ops.ln_wide_fwd(y2, ln2.weight, ln2.bias, rows, y32=xo, yb=xob, mean=m2, rstd=r2, eps=self.eps)
            per_layer.append((xb, qkv, ctx, P, y1, m1, r1, x1b, hpre, h, y2, m2, r2))
            x32, xb = xo, xob
        cls_b = xb.view(Bn, L * D)[:, :D]  # token 0 of every sequence (row pitch L*D)
        ppre = ws.get(f"bert.{tag}.ppre", [Bn, D], f32)
        ops.gemm(cls_b, self.pool.wb, Bn, D, D, bias=self.pool.bias, out32=ppre)
        pooled = ws.get(f"bert.{tag}.pooled", [Bn, D], f32)
        ops.tanh_fwd(ppre, pooled)
        self.saved[tag] = (ids, mask_u8, Bn, L, e32, me, re_, per_layer, cls_b, pooled)
        return x32, xb, pooled

    # ------------------------------------------------------------------------------------------------------------
    def backward(self, tag, d_seq, d_pooled, layers=None):
        """``layers`` = (hi, lo): process encoder layers hi-1 .. lo only (the split backward under data parallelism runs BERT's backward
        as two graphs so that the upper layers' gradient slice is exchanged while the lower layers still compute); the pooler part
        belongs to the call with hi == n_layers, the embeddings to the call with lo == 0.  None = everything."""
        with ops.sm_limit_scope(self.lim_bwd):
            return self._backward(tag, d_seq, d_pooled, layers)

    def _backward(self, tag, d_seq, d_pooled, layers=None):
        """d_seq fp32 [Bn*L, D] or None, d_pooled fp32 [Bn, D] or None; parameter gradients are accumulated into the engine's
        flat gradient buffer (the sentence and the phrase invocation share the weights)."""
        if not self.trainable:
            return
        eng, ws, bert = self.eng, self.eng.ws, self.bert
        G = eng.G
        wl = self.lim_wgrad if self.lim_wgrad > 0 else None
        D, H, FF = self.D, self.H, self.FF
        ids, mask_u8, Bn, L, e32, me, re_, per_layer, cls_b, pooled = self.saved[tag]
        rows = Bn * L
        f32 = torch.float32
        gbuf = [ws.get(f"bertb.{tag}.gA", [rows, D], f32), ws.get(f"bertb.{tag}.gB", [rows, D], f32)]
        nL = len(self.layers)
        hi, lo = layers if layers is not None else (nL, 0)
        g = gbuf[0] if hi == nL else gbuf[(nL - hi) & 1]   # the buffers alternate per layer: after nL - hi layers the gradient sits here
        if hi == nL:
            if d_seq is None:
                g.zero_()
            else:
                g.copy_(d_seq.reshape(rows, D))
        if hi == nL and d_pooled is not None:
            dpre = ws.get(f"bertb.{tag}.dpre", [Bn, D], f32)
            dpreb = ws.get(f"bertb.{tag}.dpreb", [Bn, D])
            ops.tanh_bwd(d_pooled.reshape(Bn, D), pooled, dx=dpre, dxb=dpreb)
            eng.wgrad_linear(dpreb, cls_b, G(bert.pooler.dense.weight), D, D, Bn, bias=G(bert.pooler.dense.bias), sm_limit=wl)
            dcls = ws.get(f"bertb.{tag}.dcls", [Bn, D], f32)
            ops.gemm(dpreb, self.pool.wt, Bn, D, D, out32=dcls)
            ops.rows_scatter_add(dcls, g, Bn, D, map_dst=(1, L, 0, 0))
        scale = 64 ** -0.5
        for li in reversed(range(lo, hi)):
            l = self.layers[li]
            lay = l.mod
            a = lay.attention
            xb, qkv, ctx, P, y1, m1, r1, x1b, hpre, h, y2, m2, r2 = per_layer[li]
            ln1, ln2 = a.output.LayerNorm, lay.output.LayerNorm
            dy2, dy2b = ws.get(f"bertb.{tag}.{li}.dy2", [rows, D], f32), ws.get(f"bertb.{tag}.{li}.dy2b", [rows, D])
            k = f"bert.{tag}.{li}"
            dr1, dr2 = eng.drop(k + ".drop1", self.p_hidden), eng.drop(k + ".drop2", self.p_hidden)
            ops.ln_wide_bwd(g, y2, ln2.weight, m2, r2, rows, dx32=dy2, dxb=dy2b, dgamma=G(ln2.weight), dbeta=G(ln2.bias), dxb_drop=dr2)
            eng.wgrad_linear(dy2b, h, G(lay.output.dense.weight), D, FF, rows, bias=G(lay.output.dense.bias), sm_limit=wl)
            dh = ws.get(f"bertb.{tag}.{li}.dh", [rows, FF])
            ops.gemm(dy2b, l.o2.wt, rows, FF, D, out=dh)
            dhp = ws.get(f"bertb.{tag}.{li}.dhp", [rows, FF])
            ops.gelu_bwd(dh, hpre, dhp)
            eng.wgrad_linear(dhp, x1b, G(lay.intermediate.dense.weight), FF, D, rows, bias=G(lay.intermediate.dense.bias), sm_limit=wl)
            g1 = ws.get(f"bertb.{tag}.{li}.g1", [rows, D], f32)
            ops.gemm(dhp, l.i.wt, rows, D, FF, res32=dy2, out32=g1)
            dy1, dy1b = ws.get(f"bertb.{tag}.{li}.dy1", [rows, D], f32), ws.get(f"bertb.{tag}.{li}.dy1b", [rows, D])
            ops.ln_wide_bwd(g1, y1, ln1.weight, m1, r1, rows, dx32=dy1, dxb=dy1b, dgamma=G(ln1.weight), dbeta=G(ln1.bias), dxb_drop=dr1)
            eng.wgrad_linear(dy1b, ctx, G(a.output.dense.weight), D, D, rows, bias=G(a.output.dense.bias), sm_limit=wl)
            dctx = ws.get(f"bertb.{tag}.{li}.dctx", [rows, D])
            ops.gemm(dy1b, l.o.wt, rows, D, D, out=dctx)
            dqkv = ws.get(f"bertb.{tag}.{li}.dqkv", [rows, 3 * D])
            ops.attn_small_bwd(qkv[:, :D], qkv[:, D:2 * D], qkv[:, 2 * D:], dctx, P, dqkv[:, :D], dqkv[:, D:2 * D], dqkv[:, 2 * D:], Bn, H, L, scale,
                               drop=eng.drop(k + ".attn", self.p_attn))
            for j, lin in enumerate((a.self.query, a.self.key, a.self.value)):
                sl = dqkv[:, j * D:(j + 1) * D]
                eng.wgrad_linear(sl, xb, G(lin.weight), D, D, rows, bias=G(lin.bias), sm_limit=wl)
            g_in = gbuf[1] if g is gbuf[0] else gbuf[0]
            ops.gemm(dqkv, l.qkv.wt, rows, D, 3 * D, res32=dy1, out32=g_in)
            g = g_in
        if lo != 0:
            return
        emb = bert.embeddings
        de = ws.get(f"bertb.{tag}.de", [rows, D], f32)
        ops.ln_wide_bwd(g, e32, emb.LayerNorm.weight, me, re_, rows, dx32=de, dgamma=G(emb.LayerNorm.weight), dbeta=G(emb.LayerNorm.bias),
                        dy_drop=eng.drop(f"bert.{tag}.emb", self.p_hidden))
        with eng._off():
            ops.bert_embed_bwd(de, ids, L, G(emb.word_embeddings.weight), G(emb.position_embeddings.weight), G(emb.token_type_embeddings.weight)[0])
