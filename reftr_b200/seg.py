"""Segmentation head of RefTRSeg on the C-ABI kernels: MHAttentionMap (reftr_segmentation.py:178-207) and
MaskHeadSmallConv (reftr_segmentation.py:210-280), forward and backward, wired into RefTREngine (engine.py).

Reference call stack restated (reftr_segmentation.py:152-175): visual rows of the encoder memory -> [B,256,h,w] (:166);
``bbox_attention(hs_last, memory_visual, mask)`` (:171) -> [B,1,8,h,w]; ``mask_head(cat[src_proj, memory_visual], bbox_mask,
[C4, C3, C2])`` (:172): cat -> 520 channels -> (3x3 conv + GroupNorm(8) + ReLU) x 5 with 1x1 adapters on C4/C3/C2 added after
a nearest upsample (:256-274) -> 3x3 conv to one channel at stride 4 (:276).

All convolutions run on rb_gemm over padded NHWC grids (9 row-shifted taps); GroupNorm reads the fp32 conv output and
writes the bf16 grid the next conv consumes.  The one-channel output conv is padded to 8 output channels (TMA pitch).
"""
import torch

from . import ops
from .pack import PackedConv, PackedLinear, _ver

D = 256
NH = 8


class PackedConvPadded:
    """A conv whose Cout is zero-padded to ``cout_pad`` (the 1-channel ``out_lay``): padded fp32 master copies are packed."""

    def __init__(self, conv, cout_pad=8):
        self.conv = conv
        self.Cout, self.Cin, self.kh, self.kw = conv.weight.shape
        self.Cp = cout_pad
        self.taps = self.kh * self.kw
        self._key = None
        self.wf = self.wd = self.scale = self.bias = self.w_pad = self.b_pad = None

    def current_key(self):
        return _ver(self.conv.weight, self.conv.bias)

    def refresh(self):
        w, b = self.conv.weight, self.conv.bias
        key = _ver(w, b)
        if key == self._key:
            return
        dev = w.device
        if self.wf is None or self.wf.device != dev:
            self.w_pad = torch.zeros(self.Cp, self.Cin, self.kh, self.kw, dtype=torch.float32, device=dev)
            self.b_pad = torch.zeros(self.Cp, dtype=torch.float32, device=dev)
            self.wf = torch.zeros(self.Cp, self.taps * self.Cin, dtype=ops.t16(), device=dev)
            self.wd = torch.zeros(self.Cin, self.taps * self.Cp, dtype=ops.t16(), device=dev)
            self.scale = torch.empty(self.Cp, dtype=torch.float32, device=dev)
            self.bias = torch.empty(self.Cp, dtype=torch.float32, device=dev)
        self.w_pad[:self.Cout].copy_(w.detach())
        self.b_pad[:self.Cout].copy_(b.detach())
        ops.pack_conv(self.w_pad, None, self.b_pad, self.wf, self.taps * self.Cin, self.wd, self.scale, self.bias)
        self._key = key


class SegHead:
    def __init__(self, eng, model):
        self.eng, self.model = eng, model
        mh, ba = model.mask_head, model.bbox_attention
        self.q = PackedLinear(ba.q_linear.weight, ba.q_linear.bias)
        self.k = PackedLinear(ba.k_linear.weight, ba.k_linear.bias)
        self.lays = [PackedConv(getattr(mh, f"lay{i}"), None, need_dgrad=True) for i in range(1, 6)]
        self.gns = [getattr(mh, f"gn{i}") for i in range(1, 6)]
        self.adapters = [PackedConv(getattr(mh, f"adapter{i}"), None, need_dgrad=True) for i in range(1, 4)]
        self.out = PackedConvPadded(mh.out_lay)
        self.packs = [self.q, self.k, self.out] + self.lays + self.adapters
        self.out_masks = self.out_att = None
        self.saved = None

    def reserve_scratch(self, scratch, off):
        for i, pc in enumerate(self.lays):
            n = pc.Cout * 9 * pc.Cin
            scratch[f"seg.lay{i + 1}"] = (off, n)
            off += (n + 63) // 64 * 64
        n = self.out.Cp * 9 * self.out.Cin
        scratch["seg.out"] = (off, n)
        off += (n + 63) // 64 * 64
        return off

    # ------------------------------------------------------------------------------------------------------------
    def _conv_gn(self, i, x, g, B):
        """lay{i+1} (3x3 + bias) -> GroupNorm(8) -> ReLU on grid g; returns (conv fp32 out, bf16 activation, mean, rstd)."""
        ws = self.eng.ws
        pc, gn = self.lays[i], self.gns[i]
        c32 = ws.get(f"seg.c{i}", [g.R, pc.Cout], torch.float32)
        ops.gemm(x, pc.wf, g.R, pc.Cout, pc.Cin, taps=[(s, t * pc.Cin) for t, s in enumerate(g.shifts())], bias=pc.bias, out32=c32, geom=g.geom)
        y = ws.get(f"seg.y{i}", [g.R, pc.Cout])
        mean, rstd = ws.get(f"seg.m{i}", [B * gn.num_groups], torch.float32), ws.get(f"seg.r{i}", [B * gn.num_groups], torch.float32)
        ops.groupnorm_nhwc_fwd(c32, gn.weight, gn.bias, B, g.H, g.W, pc.Cout, gn.num_groups, y, mean, rstd, relu=True, eps=gn.eps)
        return c32, y, mean, rstd

    def forward(self, feats, proj32, mem32, memb, hs32, hsb, kpm, B, h, w, L, S, T):
        eng, ws = self.eng, self.eng.ws
        if T != 1:
            raise NotImplementedError("segmentation is built for one query per image (reftr_segmentation.py:97)")
        rows, hw = B * S, h * w
        nl = len(eng.dec)
        g5 = feats[4][1]
        # ---- MHAttentionMap (:196-207) ----------------------------------------------------------------------------------
        hs_last = hsb[(nl - 1) * B:nl * B]
        q32 = ws.get("seg.q32", [B, D], torch.float32)
        ops.gemm(hs_last, self.q.wb, B, D, D, bias=self.q.bias, out32=q32)
        k32 = ws.get("seg.k32", [rows, D], torch.float32)
        ops.gemm(memb, self.k.wb, rows, D, D, bias=self.k.bias, out32=k32)
        C0 = 2 * D + NH
        grid0 = ws.get("seg.grid0", [g5.R, C0], zero=True)  # border stays zero: only interior pixels are ever written
        att = ws.get("seg.att", [B, NH, hw], torch.float32)
        scale = float(D / NH) ** -0.5
        ops.attn_map_fwd(q32, k32, kpm, B, S, L, hw, w, scale, att, grid0, 2 * D)
        # ---- cat([src_proj, memory_visual, bbox_mask]) (:166, :172, :244) ---------------------------------------------
        src32 = ws.get("enc.x0", [rows, D], torch.float32)  # input_proj + GroupNorm output, token layout (engine.forward)
        ops.tokens_to_grid(src32, B, S, L, h, w, D, grid0, 0)
        ops.tokens_to_grid(mem32, B, S, L, h, w, D, grid0, D)
        # ---- mask head (:243-280) -------------------------------------------------------------------------------------------
        sv = {}
        c0, y0, m0, r0 = self._conv_gn(0, grid0, g5, B)
        c1, y1, m1, r1 = self._conv_gn(1, y0, g5, B)
        sv[0], sv[1] = (grid0, g5, c0, y0, m0, r0), (y0, g5, c1, y1, m1, r1)
        x, gx = y1, g5
        for j, layer in enumerate((3, 2, 1)):  # FPN levels C4, C3, C2
            f, g = feats[layer]
            ad = self.adapters[j]
            cur = ws.get(f"seg.cur{j}", [g.R, ad.Cout])
            ops.gemm(f, ad.wf, g.R, ad.Cout, ad.Cin, bias=ad.bias, out=cur, geom=g.geom)
            xin = ws.get(f"seg.xin{j}", [g.R, ad.Cout])
            ops.upsample_add(x, cur, xin, B, gx.H, gx.W, g.H, g.W, ad.Cout)
            c, y, m, r = self._conv_gn(2 + j, xin, g, B)
            sv[2 + j] = (xin, g, c, y, m, r, gx, f)
            x, gx = y, g
        logits = ws.get("seg.logits", [gx.R, self.out.Cp], torch.float32)
        ops.gemm(x, self.out.wf, gx.R, self.out.Cp, self.out.Cin, taps=[(s, t * self.out.Cin) for t, s in enumerate(gx.shifts())],
                 bias=self.out.bias, out32=logits, geom=gx.geom)
        self.out_masks = ws.get("seg.pred_masks", [B, 1, gx.H, gx.W], torch.float32)
        self.out_masks.copy_(logits.view(B, gx.Hp, gx.Wp, self.out.Cp)[:, 1:-1, 1:-1, 0].unsqueeze(1))
        self.out_att = att.view(B, NH, h, w)
        self.saved = (sv, q32, k32, att, hs_last, memb, gx, B, h, w, L, S)
        return [self.out_masks, self.out_att]

    # ------------------------------------------------------------------------------------------------------------
    def backward(self, g_masks, g_att, d_hs):
        """g_masks fp32 [B,1,H/4,W/4], g_att fp32 [B,8,h,w]; adds the query path into d_hs (last decoder layer rows).
        Returns (g_mem_seg fp32 [B*S,256], g_src fp32 [B*S,256], {layer: bf16 gradient at that backbone layer's output})."""
        eng, ws, m = self.eng, self.eng.ws, self.model
        mh = m.mask_head
        sv, q32, k32, att, hs_last, memb, g2, B, h, w, L, S = self.saved
        rows, hw = B * S, h * w
        nl = len(eng.dec)
        G = eng.G
        # ---- out_lay -------------------------------------------------------------------------------------------------------
        Cp = self.out.Cp
        dl = ws.get("segb.dl", [g2.R, Cp], zero=True)
        dl.view(B, g2.Hp, g2.Wp, Cp)[:, 1:-1, 1:-1, 0].copy_(g_masks[:, 0])
        y4 = sv[4][3]
        b8 = ws.get("segb.b8", [Cp], torch.float32)
        b8.zero_()
        ops.colsum(dl, b8)  # (tiny; stays on the main stream: b8 is consumed right below)
        G(mh.out_lay.bias).add_(b8[:1])
        sc = eng.scratch_view("seg.out").view(Cp, 9 * self.out.Cin)
        sh = g2.shifts()
        ops.gemm(dl, y4, Cp, self.out.Cin, g2.R, mode=1, taps=[(0, o) for o in sh], out32=sc, atomic=True,
                 splits=0, out32_z_stride=self.out.Cin)
        w8 = ws.get("segb.w8", [Cp, self.out.Cin, 3, 3], torch.float32)
        ops.unpack_conv_grad(sc, None, w8, Cp, self.out.Cin, 9)
        G(mh.out_lay.weight).add_(w8[:1])
        dy = ws.get("segb.dy4", [g2.R, self.out.Cin])
        ops.gemm(dl, self.out.wd, g2.R, self.out.Cin, Cp, taps=[(s, t * Cp) for t, s in enumerate(sh)], out=dy, geom=g2.geom)
        g_fpn = {}
        # ---- levels C2, C3, C4 (reverse), then lay2, lay1 ----------------------------------------------------------
        for i in (4, 3, 2, 1, 0):
            pc, gn = self.lays[i], self.gns[i]
            lay = getattr(mh, f"lay{i + 1}")
            xin, g, c32, y, mean, rstd = sv[i][:6]
            dc = ws.get(f"segb.dc{i}", [g.R, pc.Cout])
            ops.groupnorm_nhwc_bwd(dy, y, c32, gn.weight, mean, rstd, B, g.H, g.W, pc.Cout, gn.num_groups, dc, G(gn.weight), G(gn.bias), relu=True)
            shg = g.shifts()
            eng.wgrad_conv(pc, dc, xin, g.R, key=f"seg.lay{i + 1}", b_offsets=shg, bias=G(lay.bias))
            dxin = ws.get(f"segb.dxin{i}", [g.R, pc.Cin])
            ops.gemm(dc, pc.wd, g.R, pc.Cin, pc.Cout, taps=[(s, t * pc.Cout) for t, s in enumerate(shg)], out=dxin, geom=g.geom)
            if i >= 2:
                j = i - 2
                ad = self.adapters[j]
                adm = getattr(mh, f"adapter{j + 1}")
                glo, f = sv[i][6], sv[i][7]
                layer = (3, 2, 1)[j]
                eng.wgrad_conv(ad, dxin, f, g.R, bias=G(adm.bias))
                first_trainable_layer = min(b.layer for b in eng.blocks if b.trainable) if any(b.trainable for b in eng.blocks) else 99
                if layer >= first_trainable_layer:  # gradient into the backbone feature (C2 = frozen layer1 output: none)
                    gf = ws.get(f"segb.gf{layer}", [g.R, ad.Cin])
                    ops.gemm(dxin, ad.wd, g.R, ad.Cin, ad.Cout, out=gf)
                    g_fpn[layer] = gf
                dy = ws.get(f"segb.dyl{i}", [glo.R, pc.Cin])
                ops.upsample_bwd(dxin, dy, B, glo.H, glo.W, g.H, g.W, pc.Cin)
            else:
                dy = dxin
        # ---- split d(cat) -> src_proj, memory_visual, bbox_mask -------------------------------------------------------
        dgrid0 = dy  # [g5.R, 520]
        g_src = ws.get("segb.g_src", [rows, D], torch.float32, zero=True)   # language rows stay zero
        g_mem = ws.get("segb.g_mem", [rows, D], torch.float32, zero=True)
        ops.grid_to_tokens(dgrid0, 0, B, S, L, h, w, D, g_src)
        ops.grid_to_tokens(dgrid0, D, B, S, L, h, w, D, g_mem)
        dq = ws.get("segb.dq", [B, D], torch.float32)
        dk = ws.get("segb.dk", [rows, D], torch.float32)
        scale = float(D / NH) ** -0.5
        ops.attn_map_bwd(g_att.reshape(B, NH, hw) if g_att is not None else None, dgrid0, 2 * D, att, q32, k32, B, S, L, hw, w, scale, dq, dk)
        ba = m.bbox_attention
        dkb = ws.get("segb.dkb", [rows, D])
        ops.cast_bf16(dk, dkb)
        eng.wgrad_linear(dkb, memb, G(ba.k_linear.weight), D, D, rows, bias=G(ba.k_linear.bias))
        ops.gemm(dkb, self.k.wt, rows, D, D, res32=g_mem, out32=g_mem)
        dqb = ws.get("segb.dqb", [B, D])
        ops.cast_bf16(dq, dqb)
        eng.wgrad_linear(dqb, hs_last, G(ba.q_linear.weight), D, D, B, bias=G(ba.q_linear.bias))
        d_last = d_hs[(nl - 1) * B:nl * B]
        ops.gemm(dqb, self.q.wt, B, D, D, res32=d_last, out32=d_last)
        return g_mem, g_src, g_fpn
