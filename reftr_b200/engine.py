"""Forward / backward orchestration of the RefTR hot path on the C-ABI kernels (include/reftr_b200.h).

``RefTREngine`` owns, per model: the packed bf16 weights (pack.py), a workspace of static device buffers (so the whole
forward and the whole backward are CUDA-graph capturable: no allocation, no host sync, fixed addresses), and one flat
fp32 gradient buffer.  ``HotPathFunction`` is the single autograd node through which the reference's nn.Module surface
(modules.py) reaches it; BERT stays a HuggingFace module outside (third party in the reference).

Reference call stack restated here: Joiner/Backbone (backbone.py:101-145) -> input_proj + GroupNorm
(reftr_transformer.py:172-175) -> map_sentence / map_phrase (:201, :250) -> VLTransformer.encode (reftr.py:99-120) ->
TransformerEncoderLayer.forward_post x N (transformer.py:168-181) -> QueryEncoder (reftr_transformer.py:41-66) ->
TransformerDecoderLayer.forward_post x N (transformer.py:231-252, decoder.norm on every layer :131-138) -> bbox_embed
(:287); segmentation adds MHAttentionMap + MaskHeadSmallConv (reftr_segmentation.py:152-175, :196-280).

Layouts: backbone activations are "padded NHWC" bf16 matrices [B*(H+2)*(W+2), C] with an exactly-zero border, so every
convolution is a GEMM over row-shifted copies of the same matrix; tokens are batch-major fp32/bf16 matrices
[B*S, 256] (row b*S + s; language tokens first).  The residual stream, LayerNorm/GroupNorm statistics and softmax are
fp32; bf16 appears only as tensor-core operands and as saved activations / activation gradients.
"""
import math
import os

import torch

from . import ops
from .pack import lin_taps, PackedStem, PackedConv, PackedLinear, PackedStack

NH = 8      # heads
DH = 32     # head dim
D = 256     # model dim


def _cdiv(a, b):
    return (a + b - 1) // b


class Grid:
    """Geometry of a padded NHWC activation [B, H+2, W+2, C] viewed as a matrix with R rows."""

    def __init__(self, B, H, W):
        self.B, self.H, self.W = B, H, W
        self.Hp, self.Wp = H + 2, W + 2
        self.R = B * self.Hp * self.Wp
        self.geom = ops.make_geom(1, self.Wp, self.Hp * self.Wp, H, W, 0)

    def half(self):
        return Grid(self.B, (self.H + 1) // 2, (self.W + 1) // 2)

    def shifts(self):
        """Row shift of tap (r, s) of a 3x3 / stride-1 / pad-1 convolution, r-major."""
        return [(r - 1) * self.Wp + (s - 1) for r in range(3) for s in range(3)]

    def s2_offsets(self):
        """Row offset into the 4 parity planes (built on THIS grid = the stride-2 output grid) of tap (r, s)."""
        offs = []
        for r in range(3):
            for s in range(3):
                plane = 2 * (r & 1) + (s & 1)
                du = (1 if r == 2 else 0) - 1
                dv = (1 if s == 2 else 0) - 1
                offs.append((plane, du * self.Wp + dv))
        return offs


HILO = set(os.environ.get("REFTR_B200_HILO", "enc,bert").split(","))  # layer groups whose forward uses (hi | residual) weight pairs


class Workspace:
    """Named static device buffers (allocated on first use, reused by every later step with the same shapes)."""

    def __init__(self, device):
        self.device = device
        self.bufs = {}

    def get(self, name, shape, dtype=None, zero=False):
        dtype = dtype or ops.t16()
        shape = tuple(int(s) for s in shape)
        t = self.bufs.get(name)
        if t is None or tuple(t.shape) != shape or t.dtype != dtype:
            if torch.cuda.is_available() and torch.cuda.is_current_stream_capturing():
                raise RuntimeError(f"workspace buffer {name!r} {shape} would be allocated during CUDA-graph capture")
            t = (torch.zeros if zero else torch.empty)(shape, dtype=dtype, device=self.device)
            self.bufs[name] = t
        return t

    def nbytes(self):
        return sum(t.numel() * t.element_size() for t in self.bufs.values())


class _Block:
    pass


class RefTREngine:
    def __init__(self, model):
        self.model = model
        self.seg = hasattr(model, "mask_head")
        body = model.img_backbone[0].body
        self.return_interm = model.img_backbone[0].return_interm_layers
        # ---- packed weights -------------------------------------------------------------------------------------
        self.stem = PackedStem(body.conv1, body.bn1, need_dgrad=False, ldk=160)
        self._stem_fused = os.environ.get("REFTR_B200_STEM_FUSED", "1") != "0"
        self.blocks = []
        first_trainable = True
        for li in range(1, 5):
            for bi, bp in enumerate(getattr(body, f"layer{li}")):
                b = _Block()
                b.key = f"layer{li}.{bi}"
                b.pname = f"img_backbone.0.body.layer{li}.{bi}"
                b.stride, b.cin, b.width = bp.stride, bp.cin, bp.width
                b.trainable = bp.conv1.weight.requires_grad
                b.c1 = PackedConv(bp.conv1, bp.bn1, need_dgrad=b.trainable)
                b.c2 = PackedConv(bp.conv2, bp.bn2, need_dgrad=b.trainable)
                b.c3 = PackedConv(bp.conv3, bp.bn3, need_dgrad=b.trainable)
                b.ds = PackedConv(bp.downsample[0], bp.downsample[1], need_dgrad=b.trainable) if bp.downsample is not None else None
                b.need_gx = b.trainable and not first_trainable  # gradient stops at the first trainable block's input
                if b.trainable:
                    first_trainable = False
                b.layer_end = bi == len(getattr(body, f"layer{li}")) - 1
                b.layer = li
                self.blocks.append(b)
        self.iproj = PackedConv(model.input_proj[0][0], None, need_dgrad=True)
        vt = model.vl_transformer
        self.enc = []
        for lay in vt.encoder.layers:
            e = _Block()
            # forward GEMMs of the encoder multiply by (hi | residual) weight pairs: the 16-bit rounding of the WEIGHTS of these
            # latency-bound layers (and of BERT's) carried most of the box error (tools/err_attrib.py)
            hl = "enc" in HILO
            e.inp = PackedLinear(lay.self_attn.in_proj_weight, lay.self_attn.in_proj_bias, hilo=hl)
            e.out = PackedLinear(lay.self_attn.out_proj.weight, lay.self_attn.out_proj.bias, hilo=hl)
            e.l1 = PackedLinear(lay.linear1.weight, lay.linear1.bias, hilo=hl)
            e.l2 = PackedLinear(lay.linear2.weight, lay.linear2.bias, hilo=hl)
            e.mod = lay
            self.enc.append(e)
        self.dec = []
        for lay in vt.decoder.layers:
            d = _Block()
            d.sa = PackedLinear(lay.self_attn.in_proj_weight, lay.self_attn.in_proj_bias)
            d.sa_out = PackedLinear(lay.self_attn.out_proj.weight, lay.self_attn.out_proj.bias)
            d.ca = PackedLinear(lay.multihead_attn.in_proj_weight, lay.multihead_attn.in_proj_bias)
            d.ca_out = PackedLinear(lay.multihead_attn.out_proj.weight, lay.multihead_attn.out_proj.bias)
            d.l1 = PackedLinear(lay.linear1.weight, lay.linear1.bias)
            d.l2 = PackedLinear(lay.linear2.weight, lay.linear2.bias)
            d.mod = lay
            self.dec.append(d)
        caw = [lay.multihead_attn.in_proj_weight for lay in vt.decoder.layers]
        cab = [lay.multihead_attn.in_proj_bias for lay in vt.decoder.layers]
        self.kstack = PackedStack(caw, cab, D, D)
        self.vstack = PackedStack(caw, cab, 2 * D, D)

        def mlp(seq):
            return [PackedLinear(seq[0].weight, seq[0].bias), PackedLinear(seq[4].weight, seq[4].bias)]
        self.map_sentence = mlp(model.map_sentence)
        self.map_phrase = mlp(model.map_phrase)
        qe = model.query_encoder
        self.qe_lin = [PackedLinear(m.weight, m.bias) for m in (qe.linear1, qe.linear2, qe.linear3)]
        self.qe_fuse = mlp(qe.fuse_encoder_query)
        self.qe_cout = PackedLinear(qe.context_out[0].weight, qe.context_out[0].bias)
        bl = model.bbox_embed.layers
        self.bbox = [PackedLinear(bl[0].weight, bl[0].bias), PackedLinear(bl[1].weight, bl[1].bias),
                     PackedLinear(bl[2].weight, bl[2].bias, pad_to=64)]
        self.packs = [self.stem, self.iproj, self.kstack, self.vstack] + self.map_sentence + self.map_phrase + self.qe_lin + \
            self.qe_fuse + [self.qe_cout] + self.bbox
        for b in self.blocks:
            self.packs += [p for p in (b.c1, b.c2, b.c3, b.ds) if p is not None]
        for e in self.enc:
            self.packs += [e.inp, e.out, e.l1, e.l2]
        for d in self.dec:
            self.packs += [d.sa, d.sa_out, d.ca, d.ca_out, d.l1, d.l2]
        if self.seg:
            from .seg import SegHead
            self.seghead = SegHead(self, model)
            self.packs += self.seghead.packs
        # ---- language backbone: BERT on the same kernels when it is a plain HF BertModel, else it stays PyTorch
        from .bert import BertEngine
        self.bert = None
        if os.environ.get("REFTR_B200_NATIVE_BERT", "1") != "0" and BertEngine.eligible(model.lang_backbone):
            self.bert = BertEngine(self, model.lang_backbone)
            self.packs += self.bert.packs
        # ---- gradient store -------------------------------------------------------------------------------------
        self.named = [(n, p) for n, p in model.named_parameters()
                      if p.requires_grad and (self.bert is not None or not n.startswith("lang_backbone."))]
        self.pnames = {id(p): n for n, p in self.named}
        self.slots = {}
        off = 0
        for n, p in self.named:
            self.slots[n] = (off, tuple(p.shape))
            off += _cdiv(p.numel(), 64) * 64
        self.n_grad = off
        self.scratch = {}
        for b in self.blocks:
            if b.trainable:
                n = b.c2.Cout * 9 * b.c2.Cin
                self.scratch[b.key + ".c2"] = (off, n)
                off += _cdiv(n, 64) * 64
        if self.seg:
            off = self.seghead.reserve_scratch(self.scratch, off)
        self.n_flat = off
        # the language backbone's slots are contiguous in named_parameters() order: [b0, b1) of the flat buffer
        bert_offs = [self.slots[n][0] for n, _ in self.named if n.startswith("lang_backbone.")]
        after = [self.slots[n][0] for n, _ in self.named if not n.startswith("lang_backbone.") and bert_offs and self.slots[n][0] > bert_offs[0]]
        self._bert_slice = (min(bert_offs), min(after) if after else self.n_grad) if bert_offs else (0, 0)
        self._split_stream = None
        self._comm_stream = None
        self._bb_gy = None
        self._bw = None
        self.gflat = None
        self.ws = None
        self.saved = {}
        self._dev = None
        self._sig = None
        self._states = {}
        self._cur = None
        self.step_id = 0
        self.use_graphs = os.environ.get("REFTR_B200_GRAPHS", "1") != "0"
        self.use_side = os.environ.get("REFTR_B200_SIDE_STREAM", "1") != "0"
        # one side stream per CATEGORY of off-critical-path work ("t": transformer / heads, "bert", "bb": conv backbone): within a
        # category the launches form one chain in issue order, categories are independent of each other (they touch disjoint
        # parameters).  With ONE chain for everything the conv backbone's weight gradients queued behind all of BERT's -- whose chain
        # is issued first -- and ran as a 1.7 ms tail after the input-gradient chain instead of next to it.
        self._sides, self._side_cat = {}, "t"
        self._side_cats = os.environ.get("REFTR_B200_SIDE_CATEGORIES", "1") != "0"
        self._side_sms_t = int(os.environ.get("REFTR_B200_SIDE_SMS_T", "32"))  # SM limit of the transformer's weight-gradient launches
        self._side2, self._side2_used = None, False
        self._side2u, self._side2u_used = None, False
        self._prio_branch = os.environ.get("REFTR_B200_BRANCH_PRIORITY", "1") != "0"
        # Priority of the stream the main chain is captured on (-1: above the side streams).  Under data parallelism it stays 0:
        # ProcessGroupNCCL's kernels run on a normal-priority stream and must not queue behind this rank's own compute
        # (measured at 2 GPUs: 10.89 ms with 0, 11.00 ms with -1; profiles/r02_branch_scheduling.log item 10)
        self._main_prio = int(os.environ.get("REFTR_B200_MAIN_PRIORITY", "-1"))
        if self._main_prio < 0 and torch.distributed.is_available() and torch.distributed.is_initialized() and torch.distributed.get_world_size() > 1 \
                and os.environ.get("REFTR_B200_MAIN_PRIORITY_DDP", "0") == "0":
            self._main_prio = 0
        self._cap_stream = None
        self._forks = os.environ.get("REFTR_B200_FORKS", "1") != "0" and self._side_cats
        self._tracked = [t for t in list(model.parameters()) + list(model.buffers())]
        self._rg_sig = tuple(p.requires_grad for p in model.parameters())
        self._synced = False
        # overflow sentinel of the 16-bit backward (see run_backward): [0] = number of non-finite steps seen so far (device counter)
        self.overflow_dev = None
        self.skip_nonfinite = os.environ.get("REFTR_B200_SKIP_NONFINITE", "1") != "0"
        self._vsig, self._pack_dev = None, None
        self._repack_graphs, self._repack_seen = {}, None
        self.force_eager = False  # bench.py: run the next steps launch by launch on the graphed workspace (profiling)
        self.launches = 0
        # ---- dropout (train mode): counter-based, nothing stored; ONE device seed rewritten before every forward --------
        # static loss scale of the backward pass: activation gradients are stored in 16 bits (IEEE half by default), so the incoming
        # gradient is multiplied by 2^k on entry and the parameter gradients by 2^-k on exit (exact); fp32 paths are unaffected
        self.grad_scale = float(os.environ.get("REFTR_B200_GRAD_SCALE", "1024"))
        # dynamic adjustment (GradScaler-style, but lazy): every SCALE_CHECK_EVERY backward passes the overflow counter is read (ONE host
        # synchronisation per interval -- the reference's own loop synchronises every step, engine_vg.py:53); new overflows halve the
        # scale, SCALE_GROWTH_INTERVAL clean steps double it (up to 2^16).  REFTR_B200_DYNAMIC_SCALE=0 keeps the scale fixed.
        self.dynamic_scale = os.environ.get("REFTR_B200_DYNAMIC_SCALE", "1") != "0"
        self._scale_seen, self._scale_clean, self._bwd_calls = 0, 0, 0
        self.p_drop = float(vt.dropout)
        self.train_mode = False
        self.seed_dev = None
        self.next_seed = None  # tests: force the seed of the next forward
        self.last_seed = None
        self._drops = {}

    # ------------------------------------------------------------------------------------------------------------
    def param_list(self):
        return [p for _, p in self.named]

    def drop(self, name, p=None):
        """Handle of the dropout site `name` (rb_dropout), or None when dropout is inactive (eval mode / p == 0)."""
        p = self.p_drop if p is None else float(p)
        if not self.train_mode or p <= 0.0:
            return None
        d = self._drops.get(name)
        if d is None or d.p != p or d.seed is not self.seed_dev:
            d = self._drops[name] = ops.Drop(self.seed_dev, name, p)
        return d

    def _new_seed(self):
        """One 63-bit seed per forward from torch's CPU generator (so torch.manual_seed governs the masks; ranks differ when
        their torch seeds differ, main_vg.py:173-177).  Written to the device by a stream-ordered fill: no host sync, and a
        replayed CUDA graph reads the new value."""
        if self.next_seed is not None:
            seed, self.next_seed = int(self.next_seed), None
        else:
            seed = int(torch.empty((), dtype=torch.int64).random_().item()) & 0x7FFFFFFFFFFFFFFF
        self.last_seed = seed
        self.seed_dev.fill_(seed)

    def _prepare(self, device):
        """Re-packs weights whose master copy changed (optimizer step / load_state_dict) and drops captured graphs when
        any parameter storage moved (graphs hold raw device addresses)."""
        if tuple(p.requires_grad for p in self.model.parameters()) != self._rg_sig:
            raise RuntimeError("reftr_b200: requires_grad of a parameter changed after the engine was built (the gradient slots, the DDP "
                               "ignore list and the frozen-layer plan are fixed at construction); freeze / unfreeze parameters BEFORE the "
                               "first forward, or call model.reset_engine() after changing them")
        self._sync_initial_state()
        sig = tuple(p.data_ptr() for _, p in self.named)
        if self._dev != device or sig != self._sig:
            self._dev, self._sig = device, sig
            self._vsig = None
            self._repack_graphs, self._repack_seen = {}, None  # graphs hold raw parameter addresses
            self._states = {}
            self.gflat = torch.zeros(self.n_flat, dtype=torch.float32, device=device)
            self._gflat_clean = True
            self.seed_dev = torch.zeros(1, dtype=torch.int64, device=device)
            self._drops = {}
        # cheap change detection first (one tuple compare); the per-pack refresh walk only runs when something moved
        vsig = tuple(t._version for t in self._tracked)
        if vsig != self._vsig or device != self._pack_dev:
            self._refresh_packs(device)
            self._vsig, self._pack_dev = tuple(t._version for t in self._tracked), device

    def _sync_initial_state(self):
        """DistributedDataParallel broadcasts rank 0's parameters and buffers at construction (main_vg.py:293-296) -- but it SKIPS
        what ``_ddp_params_and_buffers_to_ignore`` names, i.e. everything this engine owns (modules.RefTR._mark_ddp_ignored).  The
        reference seeds every rank with ``seed + rank`` before ``build_reftr`` (main_vg.py:173-177), so without this the replicas would
        start from different decoder / query-encoder / head weights and never converge to each other (only gradients are averaged).
        Once, at the first forward after a process group exists: ONE flat broadcast of every ignored floating-point tensor from
        rank 0 (a no-op for tensors that already agree)."""
        if self._synced or self._dist_world() <= 1:
            return
        self._synced = True
        ignored = set(getattr(self.model, "_ddp_params_and_buffers_to_ignore", []) or [])
        ts = [t.data for n, t in list(self.model.named_parameters()) + list(self.model.named_buffers()) if n in ignored and t.is_floating_point()]
        if not ts:
            return
        with torch.no_grad():
            flat = torch.cat([t.reshape(-1).to(torch.float32) for t in ts])
            torch.distributed.broadcast(flat, 0)
            off = 0
            for t in ts:
                t.copy_(flat[off:off + t.numel()].view_as(t))
                off += t.numel()

    def _refresh_packs(self, device):
        """Re-packs the 16-bit kernel-layout copies of every weight whose master changed.  After an optimizer step that is ALL
        trainable weights (~150 small launches every iteration), so the second time the same set is stale the launches are captured
        into one CUDA graph (sources and destinations are fixed addresses) and replayed from then on."""
        stale = [p for p in self.packs if getattr(p, "_key", None) != p.current_key()] if all(hasattr(p, "current_key") for p in self.packs) else None
        if stale is None or device.type != "cuda" or not self.use_graphs or device != self._pack_dev or not stale:
            for p in self.packs:
                p.refresh()
            return
        ids = tuple(id(p) for p in stale)
        g = self._repack_graphs.get(ids)
        if g is None and self._repack_seen == ids and len(stale) > 8:
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, **self._capture_kw()):
                for p in stale:
                    p.refresh()
            self._repack_graphs = {ids: g}
        self._repack_seen = ids
        if g is None:
            for p in stale:
                p.refresh()
        else:
            g.replay()
            for p in stale:
                p._key = p.current_key()

    # ------------------------------------------------------------------------------------------------------------
    # step driver: eager on the first step of a shape, CUDA-graph capture on the second, replay afterwards
    # ------------------------------------------------------------------------------------------------------------
    MAX_GRAPHED_SHAPES = 3

    def _state(self, key):
        st = self._states.get(key)
        if st is None:
            graphed = self.use_graphs and self._dev.type == "cuda" and \
                sum(1 for s in self._states.values() if s["graphed"]) < self.MAX_GRAPHED_SHAPES
            if not graphed and "eager" in self._states:
                st = self._states["eager"]  # all un-graphed shapes share one workspace
            else:
                st = dict(ws=Workspace(self._dev), graphed=graphed, fwd=None, bwd=None, nf=0, nb=0, fl=0, bl=0, saved=None, dims=None,
                          outs=None, bouts=None)
            self._states[key if graphed else "eager"] = st
        return st

    def run_forward(self, img, img_mask, sent_mask, mask_context, query_mask, n_ph, want_seg, sent_feat, pooled, sent_ids=None, ph_ids=None,
                    ph_mask=None):
        """Language input is either BERT's outputs (``sent_feat`` [B,L,768], ``pooled`` [B*n_ph,768]; BERT ran in PyTorch) or,
        when BERT runs on our kernels (``self.bert``), the token ids: ``sent_ids`` [B,L] and optionally ``ph_ids`` / ``ph_mask``
        [B,n_ph,Lp] of the multi-phrase configs."""
        ops.require_device(img)
        if next(self.model.parameters()).device != img.device:
            raise RuntimeError("reftr_b200: model parameters and inputs are on different devices")
        self._prepare(img.device)
        self.step_id += 1
        native = self.bert is not None
        B, L = sent_mask.shape[:2]
        T = n_ph * self.model.num_queries_per_phrase
        Lp = ph_ids.shape[-1] if (native and ph_ids is not None) else 0
        self.train_mode = bool(self.model.training)
        if self.train_mode:
            self._new_seed()
        key = (tuple(img.shape), L, n_ph, bool(want_seg), native, Lp, self.train_mode)
        st = self._state(key)
        self._cur = st
        self.ws = ws = st["ws"]
        args = [ws.get("in.img", img.shape, torch.float32), ws.get("in.imask", img_mask.shape, torch.bool),
                ws.get("in.smask", [B, L], torch.int64), ws.get("in.mctx", [B, n_ph, L], torch.uint8), ws.get("in.qmask", [B, T], torch.uint8)]
        srcs = [img, img_mask, sent_mask, mask_context, query_mask]
        if native:
            args += [ws.get("in.sids", [B, L], torch.int64), ws.get("in.lmask", [B, L], torch.uint8)]
            srcs += [sent_ids, sent_mask == 0]
            if Lp:
                args += [ws.get("in.pids", [B * n_ph, Lp], torch.int64), ws.get("in.pmask", [B * n_ph, Lp], torch.uint8)]
                srcs += [ph_ids, ph_mask == 0]
            lang = ("ids",) + tuple(args[5:])
        else:
            args += [ws.get("in.sf", sent_feat.shape, torch.float32), ws.get("in.pl", [B * n_ph, pooled.shape[-1]], torch.float32)]
            srcs += [sent_feat, pooled]
            lang = ("feat",) + tuple(args[5:])
        for dst, src in zip(args, srcs):
            dst.copy_(src.reshape(dst.shape))
        a = (args[0], args[1], args[2], args[3], args[4], n_ph, want_seg, lang)
        if self.force_eager:
            st["outs"] = self.forward(*a)
            if st["fwd"] is None:
                st["saved"], st["dims"] = self.saved, self.dims
        elif st["fwd"] is not None:
            self.saved, self.dims = st["saved"], st["dims"]
            st["fwd"].replay()
        elif st["graphed"] and st["nf"] >= 1:
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, **self._capture_kw()):
                st["outs"] = self.forward(*a)
            st["fwd"], st["saved"], st["dims"] = g, self.saved, self.dims
            g.replay()
        else:
            n0 = ops.launch_count()
            st["outs"] = self.forward(*a)
            st["fl"] = ops.launch_count() - n0
            st["saved"], st["dims"] = self.saved, self.dims
        st["nf"] += 1
        self.launches += st["fl"]
        return st["outs"]

    def run_backward(self, g_logits, g_masks, g_att):
        st = self._cur
        self.ws = ws = st["ws"]
        self.saved, self.dims = st["saved"], st["dims"]
        self._maybe_adjust_scale()
        S = self.grad_scale
        gl = ws.get("in.g_logits", g_logits.shape, torch.float32)
        torch.mul(g_logits, S, out=gl)
        gm = ga = None
        if g_masks is not None:
            gm = ws.get("in.g_masks", g_masks.shape, torch.float32)
            torch.mul(g_masks, S, out=gm)
            ga = ws.get("in.g_att", self.seghead.out_att.shape, torch.float32)
            if g_att is None:
                ga.zero_()
            else:
                torch.mul(g_att, S, out=ga)
        self._ensure_gflat_clean()
        split = self._split_wanted()
        if split and not self.force_eager and (st.get("bwd3") is not None or (st["graphed"] and st["fwd"] is not None and st["nb"] >= 1)):
            return self._run_backward_split(st, gl, gm, ga)
        if self.force_eager:
            st["bouts"] = self.backward(gl, gm, ga)
        elif st["bwd"] is not None:
            st["bwd"].replay()
        elif st["graphed"] and st["fwd"] is not None and st["nb"] >= 1:
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, **self._capture_kw()):
                st["bouts"] = self.backward(gl, gm, ga)
            st["bwd"] = g
            g.replay()
        else:
            n0 = ops.launch_count()
            st["bouts"] = self.backward(gl, gm, ga)
            st["bl"] = ops.launch_count() - n0
        st["nb"] += 1
        self.launches += st["bl"]
        d_sent, d_pooled = st["bouts"]
        import time as _t
        _t0 = _t.perf_counter()
        # fresh storage per step: autograd may keep these as .grad of leaf tensors (the copy also undoes the loss scale)
        flat = torch.empty(self.n_grad, dtype=torch.float32, device=self.gflat.device)
        self._handover(0, self.n_grad, flat)
        _t1 = _t.perf_counter()
        if getattr(self.model, "engine_allreduce", False) and torch.distributed.is_available() and torch.distributed.is_initialized():
            world = torch.distributed.get_world_size()
            if world > 1:
                # the data-parallel exchange of the path: ONE all-reduce over the flat gradient buffer, then the mean
                if torch.distributed.get_backend() == "nccl":
                    torch.distributed.all_reduce(flat, op=torch.distributed.ReduceOp.AVG)  # mean inside the collective
                else:
                    torch.distributed.all_reduce(flat)
                    flat.mul_(1.0 / world)
        self._finish_guard(flat)
        self._clear_gflat()
        self.host_ms = {"clone": (_t1 - _t0) * 1e3, "allreduce": (_t.perf_counter() - _t1) * 1e3}
        if self.bert is not None:
            return None, None, flat
        return d_sent * (1.0 / S), d_pooled * (1.0 / S), flat

    def _handover(self, lo, hi, flat):
        """flat[lo:hi] = gflat[lo:hi] / grad_scale into fresh storage, and the overflow sentinel of the 16-bit backward in the same
        pass.  Activation gradients are IEEE half under a static loss scale, so an overflow (real checkpoints, exploding loss) would
        put inf / NaN into the flat gradient and from there into clip_grad_norm_ and the AdamW moments.  rb_scale_copy_check raises
        a device flag when it copies a non-finite value (no extra traffic); ``_finish_guard`` then zeroes the whole gradient of the
        step (the update becomes a no-op apart from weight decay, like a GradScaler-skipped step) and counts it -- all on the
        device, no host synchronisation.  REFTR_B200_SKIP_NONFINITE=0 only counts (the reference's behaviour: engine_vg.py:55-58
        checks the loss, not the gradients)."""
        if self.overflow_dev is None or self.overflow_dev.device != flat.device:
            self.overflow_dev = torch.zeros(1, dtype=torch.float32, device=flat.device)
            self._flag = torch.zeros(1, dtype=torch.int32, device=flat.device)
        ops.scale_copy_check(self.gflat[lo:hi], flat[lo:hi], 1.0 / self.grad_scale, self._flag)

    def _finish_guard(self, flat):
        """After the last hand-over (and, under data parallelism, after the all-reduces: a rank that overflowed has already spread its
        inf / NaN to every replica's sum, so the flags are combined first and every rank takes the same decision)."""
        if self._dist_world() > 1:
            torch.distributed.all_reduce(self._flag, op=torch.distributed.ReduceOp.MAX)
        if self.skip_nonfinite:
            ops.zero_if(flat, self._flag, self.overflow_dev)
        else:
            self.overflow_dev.add_(self._flag.to(torch.float32))
        self._flag.zero_()

    SCALE_CHECK_EVERY = 50
    SCALE_GROWTH_INTERVAL = 2000
    SCALE_MAX, SCALE_MIN = 65536.0, 1.0

    def _maybe_adjust_scale(self):
        """Called at the start of every backward (before the incoming gradient is scaled)."""
        self._bwd_calls += 1
        if not self.dynamic_scale or self._bwd_calls % self.SCALE_CHECK_EVERY or self.overflow_dev is None:
            return
        n = self.overflow_steps()          # the one synchronisation of the interval
        if n > self._scale_seen:
            self.grad_scale = max(self.SCALE_MIN, self.grad_scale * 0.5 ** min(n - self._scale_seen, 4))
            self._scale_seen, self._scale_clean = n, 0
        else:
            self._scale_clean += self.SCALE_CHECK_EVERY
            if self._scale_clean >= self.SCALE_GROWTH_INTERVAL and self.grad_scale < self.SCALE_MAX:
                self.grad_scale *= 2.0
                self._scale_clean = 0

    def _ensure_gflat_clean(self):
        """The weight-gradient GEMMs ACCUMULATE (split-K atomics) into ``gflat``, so it must be zero when a backward starts.  Normally
        ``_clear_gflat`` has already done that behind the previous step; after an interrupted backward (exception) it is done here."""
        if not getattr(self, "_gflat_clean", False):
            self.gflat.zero_()
        ev = getattr(self, "_gflat_ev", None)
        if ev is not None:
            torch.cuda.current_stream().wait_event(ev)
            self._gflat_ev = None
        self._gflat_clean = False

    def _clear_gflat(self):
        """Zero-fill of the accumulation buffer for the NEXT backward, on a side stream ordered after this step's hand-over: 607 MB of
        writes (0.1 ms) leave the backward's critical path and overlap the next forward, which never touches the buffer."""
        if self._dev is None or self._dev.type != "cuda":
            self.gflat.zero_()
            self._gflat_clean = True
            return
        if getattr(self, "_zero_stream", None) is None:
            self._zero_stream = torch.cuda.Stream(device=self._dev)
        zs = self._zero_stream
        zs.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(zs):
            self.gflat.zero_()
            self._gflat_ev = torch.cuda.Event()
            self._gflat_ev.record(zs)
        self._gflat_clean = True

    def overflow_steps(self):
        """Number of steps whose gradient was non-finite so far (synchronises; for logging)."""
        return 0 if self.overflow_dev is None else int(self.overflow_dev.item())

    def _dist_world(self):
        if getattr(self.model, "engine_allreduce", False) and torch.distributed.is_available() and torch.distributed.is_initialized():
            return torch.distributed.get_world_size()
        return 1

    def _split_wanted(self):
        """Three-graph backward (see ``_run_backward_split``): when the engine owns the gradient exchange of a multi-GPU run and
        BERT runs on our kernels; REFTR_B200_SPLIT_BWD=1 forces it on one GPU (tests), =0 disables it."""
        env = os.environ.get("REFTR_B200_SPLIT_BWD", "")
        if env == "0" or self.bert is None or not self.bert.trainable:
            return False
        return env == "1" or self._dist_world() > 1

    def _allreduce_(self, t):
        world = self._dist_world()
        if world > 1:
            if torch.distributed.get_backend() == "nccl":
                torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.AVG)  # mean inside the collective
            else:
                torch.distributed.all_reduce(t)
                t.mul_(1.0 / world)

    def _split_plan(self):
        """Parts of the split backward and the slices of the flat gradient buffer each one COMPLETES, in the order their all-reduces are
        issued (identical on every rank).  The slots follow named_parameters(): img_backbone (layer2, layer3, layer4 -- conv1 / layer1 are
        frozen and have no slot), lang_backbone, vl_transformer .. query_encoder, input_proj."""
        def spans(pred):
            """maximal runs of consecutive slots whose parameter name satisfies pred"""
            runs = []
            for n, p in self.named:
                lo = self.slots[n][0]
                hi = lo + _cdiv(p.numel(), 64) * 64
                if pred(n):
                    if runs and runs[-1][1] == lo:
                        runs[-1] = (runs[-1][0], hi)
                    else:
                        runs.append((lo, hi))
            return runs
        b0, b1 = self._bert_slice
        bert_parts = [("bert", [(b0, b1)] if b1 > b0 else [])]
        nL = len(self.bert.layers) if self.bert is not None else 0
        if b1 > b0 and nL >= 2 and os.environ.get("REFTR_B200_SPLIT_BERT_HALVES", "1") != "0":
            # two graphs: upper half of the encoder layers (+ pooler), lower half (+ embeddings); named_parameters() order is
            # embeddings, layer 0 .. nL-1, pooler, so each half completes one or two contiguous runs of slots
            half = nL // 2
            def upper(n):
                if not n.startswith("lang_backbone."):
                    return False
                if ".encoder.layer." in n:
                    return int(n.split(".encoder.layer.")[1].split(".")[0]) >= half
                return ".pooler." in n
            up = spans(upper)
            lw = spans(lambda n: n.startswith("lang_backbone.") and not upper(n))
            if up and lw:
                bert_parts = [(f"bert:{nL}:{half}", up), (f"bert:{half}:0", lw)]
        iproj = spans(lambda n: n.startswith("input_proj."))
        rest = spans(lambda n: not n.startswith(("img_backbone.", "lang_backbone.", "input_proj.")))
        plan = [("heads", rest)]
        layers = sorted({b.layer for b in self.blocks if b.trainable}, reverse=True)
        first = True
        for li in layers:
            sl = spans(lambda n, li=li: n.startswith(f"img_backbone.0.body.layer{li}."))
            plan.append((f"bb:{li}", (iproj if first else []) + sl))
            if bert_parts:  # exchange order: layer4, BERT's upper half, layer3, BERT's lower half, layer2 (roughly the order of completion)
                plan.append(bert_parts.pop(0))
            first = False
        if not layers:  # frozen backbone: input_proj alone, then BERT
            plan.append(("bb:0", iproj))
        plan.extend(bert_parts)
        covered = sorted(sl for _, sls in plan for sl in sls)
        pos = 0
        for lo, hi in covered:  # every slot is handed over exactly once
            assert lo == pos, (plan, pos)
            pos = hi
        assert pos == self.n_grad, (plan, pos, self.n_grad)
        return plan

    def _run_backward_split(self, st, gl, gm, ga):
        """Backward as SEVERAL CUDA graphs so that the data-parallel exchange overlaps compute (one exchange step, 607 MB
        of fp32 gradients).  Parts: heads + decoder + encoder | ResNet layer4 (+ input_proj) | BERT (on a second stream, next to the
        ResNet) | layer3 | layer2.  As soon as a part has finished, its slice of the flat gradient buffer is handed over (scaled copy +
        overflow check) and all-reduced on a communication stream while the following parts still compute; only layer2's 5 MB slice is
        exchanged after the last kernel.  (Round 1 exchanged the whole non-BERT remainder, 169 MB, after the backbone: 1.85 ms exposed
        per step at 8 GPUs.)  The all-reduces are issued in the same fixed order on every rank."""
        main = torch.cuda.current_stream()
        if self._split_stream is None:
            # BERT's part runs on its own HIGH-PRIORITY stream (captured there, so its kernel nodes carry the priority): a chain of
            # short, narrow launches that must finish early -- its 440 MB slice is the largest all-reduce of the step
            self._split_stream = torch.cuda.Stream(device=self._dev, priority=self._main_prio - 1 if self._prio_branch else 0)
            self._comm_stream = torch.cuda.Stream(device=self._dev, priority=self._main_prio - 2)  # hand-over copies before anything else
        br, cs = self._split_stream, self._comm_stream
        if st.get("bwd3") is None:
            plan = self._split_plan()
            graphs = {}
            for part, _ in plan:
                g = torch.cuda.CUDAGraph()
                if part.startswith("bert"):
                    br.wait_stream(main)
                    with torch.cuda.graph(g, stream=br), self._side_category("bert"):
                        self.backward(gl, gm, ga, part=part)
                    main.wait_stream(br)
                else:
                    with torch.cuda.graph(g, **self._capture_kw()):
                        self.backward(gl, gm, ga, part=part)
                graphs[part] = g
            st["bwd3"] = (plan, graphs)
        plan, graphs = st["bwd3"]
        flat = torch.empty(self.n_grad, dtype=torch.float32, device=self._dev)
        flat.record_stream(br)
        flat.record_stream(cs)

        def exchange(event, slices):
            cs.wait_event(event)
            with torch.cuda.stream(cs):
                for lo, hi in slices:
                    self._handover(lo, hi, flat)   # fresh storage + undo the loss scale + overflow sentinel (as in run_backward)
                    self._allreduce_(flat[lo:hi])

        def done_event(stream):
            ev = torch.cuda.Event()
            ev.record(stream)
            return ev

        # Launch order: heads, then BERT's parts back to back on their own stream (they only need the heads part) next to layer4 /
        # layer3 / layer2 on the main stream.  Exchange order (the same on every rank) = the plan's order: heads, layer4, BERT's upper
        # half, layer3, BERT's lower half, layer2 -- roughly the order of completion, so no slice queues behind one that is not ready.
        early_bert = os.environ.get("REFTR_B200_SPLIT_BERT_EARLY", "1") != "0"
        bert_done = {}
        for part, slices in plan:
            if part.startswith("bert"):
                if part not in bert_done:  # (late mode: launched at its position in the plan)
                    br.wait_event(done_event(main))
                    with torch.cuda.stream(br):
                        graphs[part].replay()
                    bert_done[part] = done_event(br)
                exchange(bert_done[part], slices)
                continue
            graphs[part].replay()
            exchange(done_event(main), slices)
            if part == "heads" and early_bert:
                br.wait_event(done_event(main))   # the heads part produced BERT's incoming gradient
                with torch.cuda.stream(br):
                    for bp, _ in plan:
                        if bp.startswith("bert"):
                            graphs[bp].replay()
                            bert_done[bp] = done_event(br)
        main.wait_stream(br)
        main.wait_stream(cs)
        self._finish_guard(flat)
        self._clear_gflat()
        st["nb"] += 1
        self.launches += st["bl"]
        self.host_ms = {"clone": 0.0, "allreduce": 0.0}
        return None, None, flat

    def G(self, name_or_param):
        n = name_or_param if isinstance(name_or_param, str) else self.pnames[id(name_or_param)]
        off, shape = self.slots[n]
        numel = 1
        for s in shape:
            numel *= s
        return self.gflat[off:off + numel].view(shape)

    def scratch_view(self, key):
        off, n = self.scratch[key]
        return self.gflat[off:off + n]

    # ------------------------------------------------------------------------------------------------------------
    # generic helpers
    # ------------------------------------------------------------------------------------------------------------
    # ------------------------------------------------------------------------------------------------------------
    # Off-critical-path work.  In the backward pass the chain of INPUT gradients is the critical path; weight gradients, bias
    # gradients (column sums) and embedding scatters only feed the flat gradient buffer.  They are launched on a side stream
    # (forked after the kernel that produced their operand, joined at the end of the backward), so the hundreds of small,
    # latency-bound launches overlap the main chain -- inside the captured CUDA graph they become parallel branches.
    # Backward scratch buffers are per layer for that reason (a later layer must not overwrite an operand still being read).
    # ------------------------------------------------------------------------------------------------------------
    def _off(self):
        import contextlib
        if not self.use_side or self._dev is None or self._dev.type != "cuda":
            return contextlib.nullcontext()
        ent = self._sides.get(self._side_cat if self._side_cats else "t")
        if ent is None:
            hi = self._side_cats and self._prio_branch and self._side_cat == "bert"
            prio = self._main_prio - 1 if hi else (self._main_prio if self._side_cat == "fk" else 0)
            ent = self._sides[self._side_cat if self._side_cats else "t"] = [torch.cuda.Stream(device=self._dev, priority=prio), False]
        ev = torch.cuda.Event()
        ev.record()  # on the main (current) stream: everything launched so far is visible to the side stream
        ent[0].wait_event(ev)
        ent[1] = True
        return torch.cuda.stream(ent[0])

    def _fork(self):
        """Context manager: launches inside go to the FORK stream (main-chain priority) -- for an independent piece of the critical
        chain itself; forks from the current point of the main stream, ``_join_side()`` joins it back."""
        import contextlib
        if not self._forks:
            return contextlib.nullcontext()

        @contextlib.contextmanager
        def scope():
            with self._side_category("fk"), self._off():
                yield
        return scope()

    def _capture_kw(self):
        """Graph-capture stream: REFTR_B200_MAIN_PRIORITY < 0 captures the main chain on a stream of that priority (kernel nodes keep
        it), so critical-chain kernels are placed before pending side-stream CTAs; BERT's urgent branch then sits one level above."""
        if self._main_prio >= 0 or self._dev is None or self._dev.type != "cuda":
            return {}
        if self._cap_stream is None:
            self._cap_stream = torch.cuda.Stream(device=self._dev, priority=self._main_prio)
        return {"stream": self._cap_stream}

    def _side_category(self, cat):
        """Context manager: off-critical-path launches issued inside go to the side stream of category ``cat``."""
        import contextlib

        @contextlib.contextmanager
        def scope():
            prev, self._side_cat = self._side_cat, cat
            try:
                yield
            finally:
                self._side_cat = prev
        return scope()

    def _branch(self, urgent=False):
        """A second side stream for a whole independent CHAIN (BERT forward next to the conv backbone, BERT backward next to the
        backbone backward); ``_join_branch`` must be called on the main stream before the chain's results are consumed.
        ``urgent``: a high-priority stream (the priority is kept by graph capture as a kernel-node attribute) -- for a chain of
        many short, narrow launches next to wide persistent kernels: without it every one of its launches waits for a whole wide
        kernel to drain (BERT's backward stretched from 1.2 ms to 3.9 ms and ended after the conv backbone's)."""
        import contextlib
        if not self.use_side or self._dev is None or self._dev.type != "cuda":
            return contextlib.nullcontext()
        if urgent and self._prio_branch:
            if self._side2u is None:
                self._side2u = torch.cuda.Stream(device=self._dev, priority=self._main_prio - 1)
            ev = torch.cuda.Event()
            ev.record()
            self._side2u.wait_event(ev)
            self._side2u_used = True
            return torch.cuda.stream(self._side2u)
        if self._side2 is None:
            self._side2 = torch.cuda.Stream(device=self._dev)
        ev = torch.cuda.Event()
        ev.record()
        self._side2.wait_event(ev)
        self._side2_used = True
        return torch.cuda.stream(self._side2)

    def _join_branch(self):
        if self._side2_used:
            torch.cuda.current_stream().wait_stream(self._side2)
            self._side2_used = False
        if self._side2u_used:
            torch.cuda.current_stream().wait_stream(self._side2u)
            self._side2u_used = False

    def _join_side(self):
        if not self._sides:
            return
        cur = torch.cuda.current_stream()
        for ent in self._sides.values():
            if ent[1]:
                cur.wait_stream(ent[0])
                ent[1] = False

    def colsum(self, x, out, rows=None, N=None):
        """Bias gradient: out[N] += column sums of x (off the critical path)."""
        with self._off():
            ops.colsum(x, out, rows=rows, N=N)

    def wgrad_linear(self, dY, X, gview, M, N, K, bias=None, sm_limit=None):
        """gview[M, N] += dY[:K, :M]^T X[:K, :N]   (TN GEMM, split-K, fp32 atomics; off the critical path).
        ``bias`` [M]: the layer's bias gradient, bias[m] += sum_r dY[r, m], produced by the SAME launch (an extra N=16 MMA against
        ones in the tiles of the first column block) instead of a separate column-sum pass over dY."""
        if sm_limit is None and self._side_cat == "t" and self._side_sms_t > 0:
            sm_limit = self._side_sms_t
        with self._off():
            ops.gemm(dY, X, M, N, K, mode=1, out32=gview, atomic=True, splits=0, bias_grad=bias, sm_limit=sm_limit)  # splits 0: picked by rb_gemm

    def wgrad_conv(self, pc, dY, X, K, key=None, b_offsets=None, bias=None):
        """Folded-layout weight gradient of a convolution; b_offsets = row offset of X per tap (None: 1x1, no shift);
        ``bias``: the convolution's bias gradient from the same launch (see ``wgrad_linear``)."""
        with self._off():
            self._wgrad_conv(pc, dY, X, K, key, b_offsets, bias)

    def _wgrad_conv(self, pc, dY, X, K, key=None, b_offsets=None, bias=None):
        g = self.G(pc.conv.weight)
        Cout, Cin = pc.Cout, pc.Cin
        if b_offsets is None or len(b_offsets) == 1:
            # 1x1: the parameter layout [Cout, Cin, 1, 1] IS the GEMM's [M, N]; the FrozenBN fold (x scale[co]) rides in the epilogue
            gv = g.view(Cout, Cin)
            ops.gemm(dY, X, Cout, Cin, K, mode=1, taps=[(0, b_offsets[0] if b_offsets else 0)], out32=gv, atomic=True,
                     splits=0, row_scale=pc.scale if pc.bn is not None else None, bias_grad=bias)
        else:
            taps = len(b_offsets)
            sc = self.scratch_view(key).view(Cout, taps * Cin)
            ops.gemm(dY, X, Cout, Cin, K, mode=1, taps=[(0, o) for o in b_offsets], out32=sc, atomic=True,
                     splits=0, out32_z_stride=Cin, bias_grad=bias)
            ops.unpack_conv_grad(sc, pc.scale, g, Cout, Cin, taps)

    # ------------------------------------------------------------------------------------------------------------
    # backbone
    # ------------------------------------------------------------------------------------------------------------
    def _block_fwd(self, b, x, g):
        ws = self.ws
        k, w, cin, cout = b.key, b.width, b.cin, 4 * b.width
        a1 = ws.get(k + ".a1", [g.R, w])
        ops.gemm(x, b.c1.wf, g.R, w, cin, bias=b.c1.bias, relu=True, out=a1, geom=g.geom)
        a1s = xs = None
        if b.stride == 1:
            go = g
            a2 = ws.get(k + ".a2", [go.R, w])
            ops.gemm(a1, b.c2.wf, go.R, w, w, taps=[(s, t * w) for t, s in enumerate(g.shifts())], bias=b.c2.bias, relu=True,
                     out=a2, geom=go.geom)
        else:
            go = g.half()
            a1s = ws.get(k + ".a1s", [4 * go.R, w])
            ops.parity_split(a1, a1s, g.B, g.H, g.W, w, go.H, go.W)
            a2 = ws.get(k + ".a2", [go.R, w])
            ops.gemm(a1s, b.c2.wf, go.R, w, w, taps=[(p * go.R + o, t * w) for t, (p, o) in enumerate(go.s2_offsets())],
                     bias=b.c2.bias, relu=True, out=a2, geom=go.geom)
        if b.ds is None:
            idn = x
        elif b.stride == 1:
            idn = ws.get(k + ".idn", [go.R, cout])
            ops.gemm(x, b.ds.wf, go.R, cout, cin, bias=b.ds.bias, out=idn, geom=go.geom)
        else:
            # the 1x1 / stride-2 downsample reads input pixels (2i+1, 2j+1) of the padded grid only = parity plane 3
            xs = ws.get(k + ".xs3", [go.R, cin])
            ops.parity_split_plane(x, xs, g.B, g.H, g.W, cin, go.H, go.W, 3)
            idn = ws.get(k + ".idn", [go.R, cout])
            ops.gemm(xs, b.ds.wf, go.R, cout, cin, taps=[(-go.Wp - 1, 0)], bias=b.ds.bias, out=idn, geom=go.geom)
        y = ws.get(k + ".y", [go.R, cout])
        ops.gemm(a2, b.c3.wf, go.R, cout, w, bias=b.c3.bias, res=idn, relu=True, out=y, geom=go.geom)
        self.saved[k] = (x, g, go, a1, a1s, a2, xs, y)
        return y, go

    def _block_bwd(self, b, gy, g_extra=None):
        """gy: bf16 [go.R, cout], gradient w.r.t. the block's pre-ReLU output (already masked by y > 0).
        Returns the gradient w.r.t. the pre-ReLU output of the previous block (or None when not needed)."""
        ws = self.ws
        k, w, cin, cout = b.key, b.width, b.cin, 4 * b.width
        x, g, go, a1, a1s, a2, xs, _ = self.saved[k]
        # conv3 (1x1)
        self.wgrad_conv(b.c3, gy, a2, go.R)
        d_a2 = ws.get(k + ".d_a2", [go.R, w])
        ops.gemm(gy, b.c3.wd, go.R, w, cout, mask_src=a2, out=d_a2)
        # conv2 (3x3)
        d_a1 = ws.get(k + ".d_a1", [g.R, w])
        if b.stride == 1:
            sh = g.shifts()
            self.wgrad_conv(b.c2, d_a2, a1, g.R, key=k + ".c2", b_offsets=sh)
            ops.gemm(d_a2, b.c2.wd, g.R, w, w, taps=[(s, t * w) for t, s in enumerate(sh)], mask_src=a1, out=d_a1)
        else:
            so = go.s2_offsets()
            self.wgrad_conv(b.c2, d_a2, a1s, go.R, key=k + ".c2", b_offsets=[p * go.R + o for p, o in so])
            dxs = ws.get(k + ".dxs", [4 * go.R, w])
            for P in range(4):
                taps = [(-o, (8 - t) * w) for t, (p, o) in enumerate(so) if p == P]
                ops.gemm(d_a2, b.c2.wd, go.R, w, w, taps=taps, out=dxs[P * go.R:(P + 1) * go.R])
            ops.parity_merge(dxs, None, a1, d_a1, g.B, g.H, g.W, w, go.H, go.W)
        # conv1 (1x1) and the identity path
        self.wgrad_conv(b.c1, d_a1, x, g.R)
        if b.ds is not None and b.stride == 2:
            self.wgrad_conv(b.ds, gy, xs, go.R, b_offsets=[-go.Wp - 1])
        elif b.ds is not None:
            self.wgrad_conv(b.ds, gy, x, go.R)
        if not b.need_gx:
            return None
        gx = ws.get(k + ".gx", [g.R, cin])
        if b.ds is None:
            ops.gemm(d_a1, b.c1.wd, g.R, cin, w, res=gy, mask_src=x, out=gx)
        elif b.stride == 1:
            t = ws.get(k + ".t", [g.R, cin])
            ops.gemm(gy, b.ds.wd, g.R, cin, cout, res=g_extra, out=t)
            ops.gemm(d_a1, b.c1.wd, g.R, cin, w, res=t, mask_src=x, out=gx)
        else:
            t = ws.get(k + ".t", [g.R, cin])
            ops.gemm(d_a1, b.c1.wd, g.R, cin, w, res=g_extra, out=t)
            dxs2 = ws.get(k + ".dxs_ds3", [go.R, cin])  # gradient w.r.t. parity plane 3 (the only plane the 1x1/s2 conv reads)
            ops.gemm(gy, b.ds.wd, go.R, cin, cout, taps=[(go.Wp + 1, 0)], out=dxs2)
            ops.parity_merge_plane(dxs2, 3, t, x, gx, g.B, g.H, g.W, cin, go.H, go.W)
        return gx

    def _backbone_fwd(self, img):
        ws = self.ws
        B, _, H, W = img.shape
        H1, W1 = (H + 6 - 7) // 2 + 1, (W + 6 - 7) // 2 + 1
        H2, W2 = (H1 + 2 - 3) // 2 + 1, (W1 + 2 - 3) // 2 + 1
        g = Grid(B, H2, W2)
        x = ws.get("stem.pool", [g.R, 64])
        if self._stem_fused and W % 2 == 0 and body_is_7x7(self.stem):
            # conv1 + bn1 + relu + maxpool in one pass (the stride-2 map never reaches HBM); the batch is first rewritten as 16-bit HWC4
            hwc4 = ws.get("stem.hwc4", [B * H * (W + 2), 4])
            ops.stem_pool(img, self.stem.wrow, self.stem.bias, hwc4, x, B, H, W, H1, W1, H2, W2)
        else:
            c1 = ws.get("stem.out", [B * H1 * W1, 64])
            ops.stem_conv(img, self.stem.wf, self.stem.bias, c1, B, H, W, H1, W1)  # conv1 + bn1 + relu, im2col in shared memory
            ops.maxpool_3x3s2(c1, x, B, H1, W1, 64, H2, W2)
        feats = {}
        for b in self.blocks:
            x, g = self._block_fwd(b, x, g)
            if b.layer_end:
                feats[b.layer] = (x, g)
        return feats

    def _backbone_bwd(self, gy, g_fpn, only_layer=None):
        """gy: masked gradient at the output of the last block to process; g_fpn: {layer: unmasked bf16 gradient added at that layer's
        output}.  ``only_layer``: process that ResNet layer's blocks only (split backward); returns the gradient for the layer below."""
        for b in reversed(self.blocks):
            if not b.trainable:
                break
            if only_layer is not None and b.layer != only_layer:
                continue
            # a gradient arriving at a layer's output from the FPN adapters is injected where the NEXT layer's first
            # block forms its input gradient; that block is processed before this one, so look it up by layer index
            g_extra = g_fpn.get(b.layer - 1) if (b.ds is not None and b.need_gx) else None
            gy = self._block_bwd(b, gy, g_extra)
            if gy is None:
                break
        return gy

    # ------------------------------------------------------------------------------------------------------------
    # small fused pieces
    # ------------------------------------------------------------------------------------------------------------
    def _mlp_map_fwd(self, key, packs, seq, A, rows, K, *, y32, yb=None, ypb=None, pos32=None, rowmap=(0, 0, 0)):
        """mlp_mapping (reftr_transformer.py:14-23): Linear -> LN -> ReLU -> [Dropout] -> Linear -> LN -> ReLU."""
        ws = self.ws
        y0 = ws.get(key + ".y0", [rows, D], torch.float32)
        ops.gemm(A, packs[0].wb, rows, D, K, bias=packs[0].bias, out32=y0)
        a1 = ws.get(key + ".a1", [rows, D], torch.float32)
        a1b = ws.get(key + ".a1b", [rows, D])
        m0, r0 = ws.get(key + ".m0", [rows], torch.float32), ws.get(key + ".r0", [rows], torch.float32)
        dr = self.drop(key + ".drop", seq[3].p)  # nn.Dropout(0.1) after the first ReLU (reftr_transformer.py:19)
        ops.layernorm_fwd(y0, seq[1].weight, seq[1].bias, rows, y32=a1, yb=a1b, relu=True, mean=m0, rstd=r0, eps=seq[1].eps, drop=dr)
        y1 = ws.get(key + ".y1", [rows, D], torch.float32)
        ops.gemm(a1b, packs[1].wb, rows, D, D, bias=packs[1].bias, out32=y1)
        m1, r1 = ws.get(key + ".m1", [rows], torch.float32), ws.get(key + ".r1", [rows], torch.float32)
        ops.layernorm_fwd(y1, seq[5].weight, seq[5].bias, rows, y32=y32, yb=yb, pos32=pos32, ypb=ypb, relu=True, mean=m1, rstd=r1,
                          rowmap=rowmap, eps=seq[5].eps)
        self.saved[key] = (A, rows, K, y0, a1, a1b, m0, r0, y1, m1, r1, y32, rowmap)

    def _mlp_map_bwd(self, key, packs, seq, dy, *, need_dA=True):
        """dy: fp32 gradient w.r.t. the MLP output (read at the rows the forward wrote).  Returns fp32 dA [rows, K]."""
        ws = self.ws
        A, rows, K, y0, a1, a1b, m0, r0, y1, m1, r1, y32, rowmap = self.saved[key]
        d1 = ws.get(key + ".d1", [rows, D], torch.float32)
        d1b = ws.get(key + ".d1b", [rows, D])
        ops.layernorm_bwd(dy, y1, seq[5].weight, m1, r1, rows, y_relu=y32, dx32=d1, dxb=d1b, dgamma=self.G(seq[5].weight),
                          dbeta=self.G(seq[5].bias), rowmap=rowmap)
        self.wgrad_linear(d1b, a1b, self.G(seq[4].weight), D, D, rows, bias=self.G(seq[4].bias))
        da1 = ws.get(key + ".da1", [rows, D], torch.float32)
        ops.gemm(d1b, packs[1].wt, rows, D, D, out32=da1)
        d0 = ws.get(key + ".d0", [rows, D], torch.float32)
        d0b = ws.get(key + ".d0b", [rows, D])
        dr = self.drop(key + ".drop", seq[3].p)  # a1 > 0 <=> ReLU passed AND kept; kept gradients are scaled by 1/(1-p)
        ops.layernorm_bwd(da1, y0, seq[1].weight, m0, r0, rows, y_relu=a1, relu_scale=dr.scale if dr else 1.0, dx32=d0, dxb=d0b,
                          dgamma=self.G(seq[1].weight), dbeta=self.G(seq[1].bias))
        self.wgrad_linear(d0b, A, self.G(seq[0].weight), D, K, rows, bias=self.G(seq[0].bias))
        if not need_dA:
            return None
        dA = ws.get(key + ".dA", [rows, K], torch.float32)
        ops.gemm(d0b, packs[0].wt, rows, K, D, out32=dA)
        return dA

    # ------------------------------------------------------------------------------------------------------------
    # encoder layer (transformer.py:168-181)
    # ------------------------------------------------------------------------------------------------------------
    def _enc_fwd(self, l, e, x32, xb, xpb, kpm, pos32, B, S):
        ws = self.ws
        rows = B * S
        k = f"enc{l}"
        lay = e.mod
        qkv = ws.get(k + ".qkv", [rows, 3 * D])
        wv, tv = lin_taps(e.inp, slice(2 * D, 3 * D))
        wqk, tqk = lin_taps(e.inp, slice(0, 2 * D))
        with self._fork():  # the value projection (of x) beside the query / key projection (of x + pos)
            ops.gemm(xb, wv, rows, D, D, taps=tv, bias=e.inp.bias[2 * D:], out=qkv[:, 2 * D:])
        ops.gemm(xpb, wqk, rows, 2 * D, D, taps=tqk, bias=e.inp.bias[:2 * D], out=qkv[:, :2 * D])
        self._join_side()
        o = ws.get(k + ".o", [rows, D])
        lse = ws.get(k + ".lse", [B, NH, S], torch.float32)
        ops.attn_fwd(qkv[:, :D], qkv[:, D:2 * D], qkv[:, 2 * D:], kpm, o, lse, B, NH, S, S, DH ** -0.5, drop=self.drop(k + ".attn"))
        y1 = ws.get(k + ".y1", [rows, D], torch.float32)
        wo, to = lin_taps(e.out)
        ops.gemm(o, wo, rows, D, D, taps=to, bias=e.out.bias, res32=x32, out32=y1, drop=self.drop(k + ".drop1"))
        x1 = ws.get(k + ".x1", [rows, D], torch.float32)
        x1b = ws.get(k + ".x1b", [rows, D])
        m1, r1 = ws.get(k + ".m1", [rows], torch.float32), ws.get(k + ".r1", [rows], torch.float32)
        ops.layernorm_fwd(y1, lay.norm1.weight, lay.norm1.bias, rows, y32=x1, yb=x1b, mean=m1, rstd=r1, eps=lay.norm1.eps)
        dff = e.l1.N
        h = ws.get(k + ".h", [rows, dff])
        w1, t1 = lin_taps(e.l1)
        ops.gemm(x1b, w1, rows, dff, D, taps=t1, bias=e.l1.bias, relu=True, out=h, drop=self.drop(k + ".ffn"))
        y2 = ws.get(k + ".y2", [rows, D], torch.float32)
        w2, t2 = lin_taps(e.l2)
        ops.gemm(h, w2, rows, D, dff, taps=t2, bias=e.l2.bias, res32=x1, out32=y2, drop=self.drop(k + ".drop2"))
        xo = ws.get(k + ".xo", [rows, D], torch.float32)
        xob = ws.get(k + ".xob", [rows, D])
        xopb = ws.get(k + ".xopb", [rows, D])
        m2, r2 = ws.get(k + ".m2", [rows], torch.float32), ws.get(k + ".r2", [rows], torch.float32)
        ops.layernorm_fwd(y2, lay.norm2.weight, lay.norm2.bias, rows, y32=xo, yb=xob, pos32=pos32, ypb=xopb, mean=m2, rstd=r2,
                          eps=lay.norm2.eps)
        self.saved[k] = (xb, xpb, qkv, o, lse, y1, m1, r1, x1b, h, y2, m2, r2)
        return xo, xob, xopb

    def _ffn_bwd(self, key, lin1, lin2, mod, dy32, dyb, x_in_b, h, rows, out32, drop_out=None, drop_h=None):
        """Shared by encoder and decoder: given dy (= gradient at the FFN's residual sum, fp32 + bf16), accumulates the
        weight/bias gradients of linear1/linear2 and writes out32 = dy + d(FFN input).  In train mode ``dyb`` already carries the
        mask of the dropout on linear2's output (``drop_out``; dy32 stays the undropped residual gradient) and ``h`` is the
        dropped hidden activation, so h > 0 <=> ReLU passed and kept (``drop_h`` gives the 1/(1-p) scale)."""
        ws = self.ws
        dff = lin1.N
        self.wgrad_linear(dyb, h, self.G(mod.linear2.weight), D, dff, rows, bias=self.G(mod.linear2.bias))
        dh = ws.get(key + ".dh", [rows, dff])
        ops.gemm(dyb, lin2.wt, rows, dff, D, mask_src=h, out=dh, mask_scale=drop_h.scale if drop_h is not None else 1.0)
        self.wgrad_linear(dh, x_in_b, self.G(mod.linear1.weight), dff, D, rows, bias=self.G(mod.linear1.bias))
        ops.gemm(dh, lin1.wt, rows, D, dff, res32=dy32, out32=out32)

    def _enc_bwd(self, l, e, g, g_out, kpm, dpos, B, S):
        """g: fp32 gradient w.r.t. the layer output; writes g_out = gradient w.r.t. the layer input; dpos += d(pos)."""
        ws = self.ws
        rows = B * S
        k = f"enc{l}"
        lay = e.mod
        xb, xpb, qkv, o, lse, y1, m1, r1, x1b, h, y2, m2, r2 = self.saved[k]
        dy2 = ws.get(f"encb{l}.dy2", [rows, D], torch.float32)
        dy2b = ws.get(f"encb{l}.dy2b", [rows, D])
        dr1, dr2 = self.drop(k + ".drop1"), self.drop(k + ".drop2")
        ops.layernorm_bwd(g, y2, lay.norm2.weight, m2, r2, rows, dx32=dy2, dxb=dy2b, dgamma=self.G(lay.norm2.weight),
                          dbeta=self.G(lay.norm2.bias), dxb_drop=dr2)
        g1 = ws.get(f"encb{l}.g1", [rows, D], torch.float32)
        self._ffn_bwd(f"encb{l}", e.l1, e.l2, lay, dy2, dy2b, x1b, h, rows, g1, drop_out=dr2, drop_h=self.drop(k + ".ffn"))
        dy1 = ws.get(f"encb{l}.dy1", [rows, D], torch.float32)
        dy1b = ws.get(f"encb{l}.dy1b", [rows, D])
        ops.layernorm_bwd(g1, y1, lay.norm1.weight, m1, r1, rows, dx32=dy1, dxb=dy1b, dgamma=self.G(lay.norm1.weight),
                          dbeta=self.G(lay.norm1.bias), dxb_drop=dr1)
        self.wgrad_linear(dy1b, o, self.G(lay.self_attn.out_proj.weight), D, D, rows, bias=self.G(lay.self_attn.out_proj.bias))
        do = ws.get(f"encb{l}.do", [rows, D])
        ops.gemm(dy1b, e.out.wt, rows, D, D, out=do)
        dqkv = ws.get(f"encb{l}.dqkv", [rows, 3 * D])
        dbuf = ws.get(f"encb{l}.dbuf", [B, NH, S], torch.float32)
        ops.attn_bwd(qkv[:, :D], qkv[:, D:2 * D], qkv[:, 2 * D:], kpm, o, do, lse, dqkv[:, :D], dqkv[:, D:2 * D], dqkv[:, 2 * D:], dbuf,
                     B, NH, S, S, DH ** -0.5, drop=self.drop(k + ".attn"))
        gw, gbi = self.G(lay.self_attn.in_proj_weight), self.G(lay.self_attn.in_proj_bias)
        self.wgrad_linear(dqkv[:, :2 * D], xpb, gw[:2 * D], 2 * D, D, rows, bias=gbi[:2 * D])
        self.wgrad_linear(dqkv[:, 2 * D:], xb, gw[2 * D:], D, D, rows, bias=gbi[2 * D:])
        ops.gemm(dqkv, e.inp.wt, rows, D, 3 * D, res32=dy1, out32=g_out)
        with self._off():  # d(pos) of the encoder layers only feeds the embedding gradients at the very end: off the critical path
            ops.gemm(dqkv[:, :2 * D], e.inp.wt[:, :2 * D], rows, D, 2 * D, res32=dpos, out32=dpos)

    # ------------------------------------------------------------------------------------------------------------
    # query encoder (reftr_transformer.py:41-66)
    # ------------------------------------------------------------------------------------------------------------
    def _qenc_fwd(self, mem32, ph32, mctx, B, S, L, n_ph, n_q):
        ws = self.ws
        qe = self.model.query_encoder
        rl, rp, T = B * L, B * n_ph, n_ph * n_q
        ctx32 = ws.get("qe.ctx32", [rl, D], torch.float32)
        ctxb = ws.get("qe.ctxb", [rl, D])
        ops.rows_add(mem32, None, rl, D, y32=ctx32, yb=ctxb, map_a=(L, S, 1, 0))
        cls_b = ctxb.view(B, L * D)[:, :D]  # post-encoder CLS token of every sample
        k32 = ws.get("qe.k32", [B, D], torch.float32)
        q32 = ws.get("qe.q32", [rl, D], torch.float32)
        v32 = ws.get("qe.v32", [rl, D], torch.float32)
        ops.gemm(cls_b, self.qe_lin[0].wb, B, D, D, bias=self.qe_lin[0].bias, out32=k32)
        ops.gemm(ctxb, self.qe_lin[1].wb, rl, D, D, bias=self.qe_lin[1].bias, out32=q32)
        ops.gemm(ctxb, self.qe_lin[2].wb, rl, D, D, bias=self.qe_lin[2].bias, out32=v32)
        att = ws.get("qe.att", [B, n_ph, L], torch.float32)
        c32 = ws.get("qe.c32", [rp, D], torch.float32)
        ops.qenc_pool_fwd(k32, q32, v32, mctx, B, L, n_ph, att, c32)
        cb = ws.get("qe.cb", [rp, D])
        ops.cast_bf16(c32, cb)
        co32 = ws.get("qe.co32", [rp, D], torch.float32)
        ops.gemm(cb, self.qe_cout.wb, rp, D, D, bias=self.qe_cout.bias, out32=co32)
        cn32 = ws.get("qe.cn32", [rp, D], torch.float32)
        mc, rc = ws.get("qe.mc", [rp], torch.float32), ws.get("qe.rc", [rp], torch.float32)
        ln = qe.context_out[1]
        ops.layernorm_fwd(co32, ln.weight, ln.bias, rp, y32=cn32, mean=mc, rstd=rc, eps=ln.eps)
        fin = ws.get("qe.fin", [rp, 2 * D])
        ops.rows_add(cn32, mem32, rp, D, yb=fin[:, :D], map_b=(n_ph, S, 0, 0))   # + memory CLS token (residual, :58)
        ops.rows_add(ph32, None, rp, D, yb=fin[:, D:])                            # cat([context, phrase]) (:60)
        f32 = ws.get("qe.f32", [rp, D], torch.float32)
        self._mlp_map_fwd("qe.fuse", self.qe_fuse, qe.fuse_encoder_query, fin, rp, 2 * D, y32=f32)
        rt = B * T
        tgt32 = ws.get("qe.tgt32", [rt, D], torch.float32)
        tgtb = ws.get("qe.tgtb", [rt, D])
        qpos32 = ws.get("qe.qpos32", [rt, D], torch.float32)
        tqb = ws.get("qe.tqb", [rt, D])
        qw = qe.query_embed.weight
        ops.rows_add(f32, qw[:, :D], rt, D, y32=tgt32, yb=tgtb, map_a=(n_q, 1, 0, 0), map_b=(n_q, 0, 1, 0))
        ops.rows_add(f32, qw[:, D:], rt, D, y32=qpos32, map_a=(n_q, 1, 0, 0), map_b=(n_q, 0, 1, 0))
        ops.rows_add(tgt32, qpos32, rt, D, yb=tqb)
        self.saved["qe"] = (ctxb, cls_b, k32, q32, v32, att, cb, co32, mc, rc, f32)
        return tgt32, tgtb, qpos32, tqb

    def _qenc_bwd(self, d_tgt, d_qpos, g_mem, mctx, B, S, L, n_ph, n_q):
        """Adds the query encoder's gradient into g_mem (language rows) and returns fp32 d(phrase feats) [B*n_ph, 256]."""
        ws = self.ws
        qe = self.model.query_encoder
        rl, rp, T = B * L, B * n_ph, n_ph * n_q
        rt = B * T
        ctxb, cls_b, k32, q32, v32, att, cb, co32, mc, rc, f32 = self.saved["qe"]
        d_f = ws.get("qeb.d_f", [rp, D], torch.float32)
        d_f.zero_()
        ops.rows_scatter_add(d_tgt, d_f, rt, D, map_dst=(n_q, 1, 0, 0))
        ops.rows_scatter_add(d_qpos, d_f, rt, D, map_dst=(n_q, 1, 0, 0))
        gq = self.G(qe.query_embed.weight)
        with self._off():
            ops.rows_scatter_add(d_tgt, gq[:, :D], rt, D, map_dst=(n_q, 0, 1, 0))
            ops.rows_scatter_add(d_qpos, gq[:, D:], rt, D, map_dst=(n_q, 0, 1, 0))
        d_fin = self._mlp_map_bwd("qe.fuse", self.qe_fuse, qe.fuse_encoder_query, d_f)  # fp32 [rp, 512]
        d_left = ws.get("qeb.d_left", [rp, D], torch.float32)
        d_ph = ws.get("qeb.d_ph", [rp, D], torch.float32)
        ops.rows_add(d_fin[:, :D], None, rp, D, y32=d_left)
        ops.rows_add(d_fin[:, D:], None, rp, D, y32=d_ph)
        ops.rows_scatter_add(d_left, g_mem, rp, D, map_dst=(n_ph, S, 0, 0))  # residual to the memory CLS token
        ln = qe.context_out[1]
        d_co = ws.get("qeb.d_co", [rp, D], torch.float32)
        d_cob = ws.get("qeb.d_cob", [rp, D])
        ops.layernorm_bwd(d_left, co32, ln.weight, mc, rc, rp, dx32=d_co, dxb=d_cob, dgamma=self.G(ln.weight), dbeta=self.G(ln.bias))
        self.wgrad_linear(d_cob, cb, self.G(qe.context_out[0].weight), D, D, rp, bias=self.G(qe.context_out[0].bias))
        d_c = ws.get("qeb.d_c", [rp, D], torch.float32)
        ops.gemm(d_cob, self.qe_cout.wt, rp, D, D, out32=d_c)
        dk = ws.get("qeb.dk", [B, D], torch.float32)
        dq = ws.get("qeb.dq", [rl, D], torch.float32)
        dv = ws.get("qeb.dv", [rl, D], torch.float32)
        ops.qenc_pool_bwd(d_c, k32, q32, v32, att, B, L, n_ph, dk, dq, dv)
        dkb, dqb, dvb = ws.get("qeb.dkb", [B, D]), ws.get("qeb.dqb", [rl, D]), ws.get("qeb.dvb", [rl, D])
        ops.cast_bf16(dk, dkb)
        ops.cast_bf16(dq, dqb)
        ops.cast_bf16(dv, dvb)
        for lin, mod, d32, db, x, n in ((self.qe_lin[0], qe.linear1, dk, dkb, cls_b, B), (self.qe_lin[1], qe.linear2, dq, dqb, ctxb, rl),
                                        (self.qe_lin[2], qe.linear3, dv, dvb, ctxb, rl)):
            self.wgrad_linear(db, x, self.G(mod.weight), D, D, n, bias=self.G(mod.bias))
        d_ctx = ws.get("qeb.d_ctx", [rl, D], torch.float32)
        ops.gemm(dqb, self.qe_lin[1].wt, rl, D, D, out32=d_ctx)
        ops.gemm(dvb, self.qe_lin[2].wt, rl, D, D, res32=d_ctx, out32=d_ctx)
        d_cls = ws.get("qeb.d_cls", [B, D], torch.float32)
        ops.gemm(dkb, self.qe_lin[0].wt, B, D, D, out32=d_cls)
        ops.rows_scatter_add(d_ctx, g_mem, rl, D, map_dst=(L, S, 1, 0))
        ops.rows_scatter_add(d_cls, g_mem, B, D, map_dst=(1, S, 0, 0))
        return d_ph

    # ------------------------------------------------------------------------------------------------------------
    # decoder (transformer.py:114-143, :231-252)
    # ------------------------------------------------------------------------------------------------------------
    def _dec_kv_fwd(self, memb, mempb, rows):
        """The cross-attention key / value projections of ALL decoder layers (two N = 6 x 256 GEMMs over the memory); they only need
        the encoder's output, so they run on the fork stream beside the query encoder's chain of small launches."""
        nl = len(self.dec)
        kall = self.ws.get("dec.kall", [rows, nl * D])
        vall = self.ws.get("dec.vall", [rows, nl * D])
        ops.gemm(mempb, self.kstack.wb, rows, nl * D, D, bias=self.kstack.bias, out=kall)
        ops.gemm(memb, self.vstack.wb, rows, nl * D, D, bias=self.vstack.bias, out=vall)

    def _dec_fwd(self, tgt32, tgtb, qpos32, tqb, memb, mempb, kpm, qmask, B, S, T):
        ws = self.ws
        rows, rt = B * S, B * T
        nl = len(self.dec)
        vt = self.model.vl_transformer
        kall = ws.get("dec.kall", [rows, nl * D])   # written by _dec_kv_fwd on the fork stream
        vall = ws.get("dec.vall", [rows, nl * D])
        self._join_side()
        hs32 = ws.get("dec.hs32", [nl * rt, D], torch.float32)
        hsb = ws.get("dec.hsb", [nl * rt, D])
        mh, rh = ws.get("dec.mh", [nl * rt], torch.float32), ws.get("dec.rh", [nl * rt], torch.float32)
        scale = DH ** -0.5
        for l, d in enumerate(self.dec):
            k = f"dec{l}"
            lay = d.mod
            # self attention over the T queries of a sample (key padding = query_mask)
            qkv = ws.get(k + ".qkv", [rt, 3 * D])
            lse_s = ws.get(k + ".lse_s", [B, NH, T], torch.float32)
            if T == 1:
                # one query per sample (every RES/REC config): the self-attention softmax runs over a single key, so its output is
                # exactly the value projection; q / k projections, the attention kernel and their (exactly zero) gradients are skipped
                # (train mode: the dropout on that single attention probability zeroes whole heads -- epilogue dropout over [rt, 8])
                o_s = qkv[:, 2 * D:]
                ops.gemm(tgtb, d.sa.wb[2 * D:], rt, D, D, bias=d.sa.bias[2 * D:], out=o_s, drop=self.drop(k + ".sa"), drop_gshift=5)
            else:
                ops.gemm(tqb, d.sa.wb[:2 * D], rt, 2 * D, D, bias=d.sa.bias[:2 * D], out=qkv[:, :2 * D])
                ops.gemm(tgtb, d.sa.wb[2 * D:], rt, D, D, bias=d.sa.bias[2 * D:], out=qkv[:, 2 * D:])
                o_s = ws.get(k + ".o_s", [rt, D])
                ops.attn_fwd(qkv[:, :D], qkv[:, D:2 * D], qkv[:, 2 * D:], qmask, o_s, lse_s, B, NH, T, T, scale, drop=self.drop(k + ".sa"))
            y1 = ws.get(k + ".y1", [rt, D], torch.float32)
            ops.gemm(o_s, d.sa_out.wb, rt, D, D, bias=d.sa_out.bias, res32=tgt32, out32=y1, drop=self.drop(k + ".drop1"))
            t1 = ws.get(k + ".t1", [rt, D], torch.float32)
            t1qb = ws.get(k + ".t1qb", [rt, D])
            m1, r1 = ws.get(k + ".m1", [rt], torch.float32), ws.get(k + ".r1", [rt], torch.float32)
            ops.layernorm_fwd(y1, lay.norm1.weight, lay.norm1.bias, rt, y32=t1, pos32=qpos32, ypb=t1qb, mean=m1, rstd=r1, eps=lay.norm1.eps)
            # cross attention: q = t1 + query_pos, k = memory + pos, v = memory
            qc = ws.get(k + ".qc", [rt, D])
            ops.gemm(t1qb, d.ca.wb[:D], rt, D, D, bias=d.ca.bias[:D], out=qc)
            kl, vl = kall[:, l * D:(l + 1) * D], vall[:, l * D:(l + 1) * D]
            o_c = ws.get(k + ".o_c", [rt, D])
            lse_c = ws.get(k + ".lse_c", [B, NH, T], torch.float32)
            ops.attn_fwd(qc, kl, vl, kpm, o_c, lse_c, B, NH, T, S, scale, drop=self.drop(k + ".ca"))
            y2 = ws.get(k + ".y2", [rt, D], torch.float32)
            ops.gemm(o_c, d.ca_out.wb, rt, D, D, bias=d.ca_out.bias, res32=t1, out32=y2, drop=self.drop(k + ".drop2"))
            t2 = ws.get(k + ".t2", [rt, D], torch.float32)
            t2b = ws.get(k + ".t2b", [rt, D])
            m2, r2 = ws.get(k + ".m2", [rt], torch.float32), ws.get(k + ".r2", [rt], torch.float32)
            ops.layernorm_fwd(y2, lay.norm2.weight, lay.norm2.bias, rt, y32=t2, yb=t2b, mean=m2, rstd=r2, eps=lay.norm2.eps)
            dff = d.l1.N
            h = ws.get(k + ".h", [rt, dff])
            ops.gemm(t2b, d.l1.wb, rt, dff, D, bias=d.l1.bias, relu=True, out=h, drop=self.drop(k + ".ffn"))
            y3 = ws.get(k + ".y3", [rt, D], torch.float32)
            ops.gemm(h, d.l2.wb, rt, D, dff, bias=d.l2.bias, res32=t2, out32=y3, drop=self.drop(k + ".drop3"))
            t3 = ws.get(k + ".t3", [rt, D], torch.float32)
            t3b = ws.get(k + ".t3b", [rt, D])
            t3qb = ws.get(k + ".t3qb", [rt, D])
            m3, r3 = ws.get(k + ".m3", [rt], torch.float32), ws.get(k + ".r3", [rt], torch.float32)
            ops.layernorm_fwd(y3, lay.norm3.weight, lay.norm3.bias, rt, y32=t3, yb=t3b, pos32=qpos32, ypb=t3qb, mean=m3, rstd=r3,
                              eps=lay.norm3.eps)
            # shared final LayerNorm on every layer's output (return_intermediate, transformer.py:131-138)
            sl = slice(l * rt, (l + 1) * rt)
            ops.layernorm_fwd(t3, vt.decoder.norm.weight, vt.decoder.norm.bias, rt, y32=hs32[sl], yb=hsb[sl], mean=mh[sl], rstd=rh[sl],
                              eps=vt.decoder.norm.eps)
            self.saved[k] = (tgtb, tqb, qkv, o_s, lse_s, y1, m1, r1, t1qb, qc, o_c, lse_c, y2, m2, r2, t2b, h, y3, m3, r3, t3)
            tgt32, tgtb, tqb = t3, t3b, t3qb
        self.saved["dec"] = (kall, vall, mh, rh)
        return hs32, hsb

    def _dec_bwd(self, d_hs, memb, mempb, kpm, qmask, g_mem, dpos, B, S, T):
        """d_hs: fp32 [nl*B*T, 256].  Writes g_mem (= d memory from the cross-attention V path + K path) and dpos
        (= K path only); returns (d_tgt0, d_qpos) fp32 [B*T, 256]."""
        ws = self.ws
        rows, rt = B * S, B * T
        nl = len(self.dec)
        vt = self.model.vl_transformer
        kall, vall, mh, rh = self.saved["dec"]
        dkall = ws.get("decb.dkall", [rows, nl * D])
        dvall = ws.get("decb.dvall", [rows, nl * D])
        dqpos = ws.get("decb.dqpos", [rt, D], torch.float32)
        dqpos.zero_()
        gbuf = [ws.get("decb.gA", [rt, D], torch.float32), ws.get("decb.gB", [rt, D], torch.float32)]
        g_next = None
        scale = DH ** -0.5
        for l in reversed(range(nl)):
            d = self.dec[l]
            k = f"dec{l}"
            lay = d.mod
            tgtb, tqb, qkv, o_s, lse_s, y1, m1, r1, t1qb, qc, o_c, lse_c, y2, m2, r2, t2b, h, y3, m3, r3, t3 = self.saved[k]
            sl = slice(l * rt, (l + 1) * rt)
            gh = ws.get(f"decb{l}.gh", [rt, D], torch.float32)
            ops.layernorm_bwd(d_hs[sl], t3, vt.decoder.norm.weight, mh[sl], rh[sl], rt, dx32=gh, dgamma=self.G(vt.decoder.norm.weight),
                              dbeta=self.G(vt.decoder.norm.bias))
            dy3 = ws.get(f"decb{l}.dy3", [rt, D], torch.float32)
            dy3b = ws.get(f"decb{l}.dy3b", [rt, D])
            dr1, dr2, dr3 = self.drop(k + ".drop1"), self.drop(k + ".drop2"), self.drop(k + ".drop3")
            ops.layernorm_bwd(gh, y3, lay.norm3.weight, m3, r3, rt, dy2=g_next, dx32=dy3, dxb=dy3b, dgamma=self.G(lay.norm3.weight),
                              dbeta=self.G(lay.norm3.bias), dxb_drop=dr3)
            g2 = ws.get(f"decb{l}.g2", [rt, D], torch.float32)
            self._ffn_bwd(f"decb{l}", d.l1, d.l2, lay, dy3, dy3b, t2b, h, rt, g2, drop_out=dr3, drop_h=self.drop(k + ".ffn"))
            dy2 = ws.get(f"decb{l}.dy2", [rt, D], torch.float32)
            dy2b = ws.get(f"decb{l}.dy2b", [rt, D])
            ops.layernorm_bwd(g2, y2, lay.norm2.weight, m2, r2, rt, dx32=dy2, dxb=dy2b, dgamma=self.G(lay.norm2.weight),
                              dbeta=self.G(lay.norm2.bias), dxb_drop=dr2)
            # cross attention
            self.wgrad_linear(dy2b, o_c, self.G(lay.multihead_attn.out_proj.weight), D, D, rt, bias=self.G(lay.multihead_attn.out_proj.bias))
            do_c = ws.get(f"decb{l}.do_c", [rt, D])
            ops.gemm(dy2b, d.ca_out.wt, rt, D, D, out=do_c)
            dqc = ws.get(f"decb{l}.dqc", [rt, D])
            dbuf = ws.get(f"decb{l}.dbuf", [B, NH, T], torch.float32)
            cs = slice(l * D, (l + 1) * D)
            ops.attn_bwd(qc, kall[:, cs], vall[:, cs], kpm, o_c, do_c, lse_c, dqc, dkall[:, cs], dvall[:, cs], dbuf, B, NH, T, S, scale,
                         drop=self.drop(k + ".ca"))
            gw = self.G(lay.multihead_attn.in_proj_weight)
            gb = self.G(lay.multihead_attn.in_proj_bias)
            self.wgrad_linear(dqc, t1qb, gw[:D], D, D, rt, bias=gb[:D])
            self.wgrad_linear(dkall[:, cs], mempb, gw[D:2 * D], D, D, rows, bias=gb[D:2 * D])
            self.wgrad_linear(dvall[:, cs], memb, gw[2 * D:], D, D, rows, bias=gb[2 * D:])
            g1 = ws.get(f"decb{l}.g1", [rt, D], torch.float32)
            ops.gemm(dqc, d.ca.wt[:, :D], rt, D, D, res32=dy2, out32=g1)
            ops.gemm(dqc, d.ca.wt[:, :D], rt, D, D, res32=dqpos, out32=dqpos)
            dy1 = ws.get(f"decb{l}.dy1", [rt, D], torch.float32)
            dy1b = ws.get(f"decb{l}.dy1b", [rt, D])
            ops.layernorm_bwd(g1, y1, lay.norm1.weight, m1, r1, rt, dx32=dy1, dxb=dy1b, dgamma=self.G(lay.norm1.weight),
                              dbeta=self.G(lay.norm1.bias), dxb_drop=dr1)
            # self attention
            self.wgrad_linear(dy1b, o_s, self.G(lay.self_attn.out_proj.weight), D, D, rt, bias=self.G(lay.self_attn.out_proj.bias))
            do_s = ws.get(f"decb{l}.do_s", [rt, D])
            if T == 1:  # the head dropout of the forward shortcut, applied to d(attention output) = d(value projection)
                ops.gemm(dy1b, d.sa_out.wt, rt, D, D, out=do_s, drop=self.drop(k + ".sa"), drop_gshift=5)
            else:
                ops.gemm(dy1b, d.sa_out.wt, rt, D, D, out=do_s)
            g_prev = gbuf[l & 1]
            gws = self.G(lay.self_attn.in_proj_weight)
            if T == 1:  # d(value projection) = d(attention output); q / k receive no gradient
                self.wgrad_linear(do_s, tgtb, gws[2 * D:], D, D, rt, bias=self.G(lay.self_attn.in_proj_bias)[2 * D:])
                ops.gemm(do_s, d.sa.wt[:, 2 * D:], rt, D, D, res32=dy1, out32=g_prev)
            else:
                dqkv = ws.get(f"decb{l}.dqkv", [rt, 3 * D])
                ops.attn_bwd(qkv[:, :D], qkv[:, D:2 * D], qkv[:, 2 * D:], qmask, o_s, do_s, lse_s, dqkv[:, :D], dqkv[:, D:2 * D], dqkv[:, 2 * D:],
                             dbuf, B, NH, T, T, scale, drop=self.drop(k + ".sa"))
                gbs = self.G(lay.self_attn.in_proj_bias)
                self.wgrad_linear(dqkv[:, :2 * D], tqb, gws[:2 * D], 2 * D, D, rt, bias=gbs[:2 * D])
                self.wgrad_linear(dqkv[:, 2 * D:], tgtb, gws[2 * D:], D, D, rt, bias=gbs[2 * D:])
                ops.gemm(dqkv, d.sa.wt, rt, D, 3 * D, res32=dy1, out32=g_prev)
                ops.gemm(dqkv[:, :2 * D], d.sa.wt[:, :2 * D], rt, D, 2 * D, res32=dqpos, out32=dqpos)
            g_next = g_prev
        # memory gradient of all layers' K / V projections in two GEMMs (K = nl*256)
        ops.gemm(dkall, self.kstack.wt, rows, D, nl * D, out32=dpos)
        ops.gemm(dvall, self.vstack.wt, rows, D, nl * D, res32=dpos, out32=g_mem)
        return g_next, dqpos

    # ------------------------------------------------------------------------------------------------------------
    # top level
    # ------------------------------------------------------------------------------------------------------------
    def forward(self, img, img_mask, sent_mask, mask_context, query_mask, n_ph, want_seg, lang):
        """All tensor arguments are static workspace buffers staged by ``run_forward`` (fp32 image, bool image mask,
        int64 sentence mask, u8 context / query masks; ``lang`` = ("feat", BERT features, pooled) or ("ids", token ids, key mask,
        [phrase ids, phrase key mask]))."""
        m = self.model
        ws = self.ws
        self.saved = {}
        vt = m.vl_transformer
        B, _, H, W = img.shape
        L = sent_mask.shape[1]
        n_q = m.num_queries_per_phrase
        T = n_ph * n_q
        if L > vt.max_lang_seq:
            raise ValueError(f"sentence length {L} exceeds max_lang_seq {vt.max_lang_seq} (reftr.py:81)")
        # ---- language backbone first, on a branch stream: it is independent of the conv backbone and latency-bound ---------
        if lang[0] == "ids":  # BERT on our kernels (reftr_transformer.py:200, :215-217)
            with self._branch(urgent=os.environ.get("REFTR_B200_BRANCH_PRIORITY_FWD", "1") == "1"):
                _, sfb, pooled = self.bert.forward("s", lang[1], lang[2], B, L)
                if len(lang) > 3:
                    _, _, pooled = self.bert.forward("p", lang[3], lang[4], B * n_ph, lang[3].shape[1])
        # ---- backbone (backbone.py:101-109) ---------------------------------------------------------------------
        feats = self._backbone_fwd(img)
        c5, g5 = feats[4]
        h, w = g5.H, g5.W
        S = L + h * w
        rows = B * S
        # ---- positions + key-padding mask (position_encoding.py:36-56, reftr.py:57-97) ---------------------------
        pos32 = ws.get("pos32", [rows, D], torch.float32)
        kpm = ws.get("kpm", [B, S], torch.uint8)
        ops.build_pos_mask(img_mask, B, H, W, h, w, sent_mask, L, vt.lang_pos_embeddings.weight, vt.token_type_embeddings.weight,
                           vt.level_embed, pos32, kpm)
        mctx, qmask = mask_context, query_mask
        # ---- input_proj + GroupNorm -> visual token rows (reftr_transformer.py:172-175, reftr.py:57) --------------
        proj32 = ws.get("iproj.out", [g5.R, D], torch.float32)
        ops.gemm(c5, self.iproj.wf, g5.R, D, 2048, bias=self.iproj.bias, out32=proj32)
        x32 = ws.get("enc.x0", [rows, D], torch.float32)
        xb = ws.get("enc.x0b", [rows, D])
        xpb = ws.get("enc.x0pb", [rows, D])
        gn = m.input_proj[0][1]
        gmean, grstd = ws.get("iproj.mean", [B * 32], torch.float32), ws.get("iproj.rstd", [B * 32], torch.float32)
        ops.groupnorm_tokens_fwd(proj32, gn.weight, gn.bias, B, h, w, S, L, x32, xb, pos32, xpb, gmean, grstd, eps=gn.eps)
        # ---- language features -> language token rows (reftr_transformer.py:201, reftr.py:79-97) -------------------
        if lang[0] == "ids":
            self._join_branch()
        else:
            sent_feat, pooled = lang[1], lang[2]
            sfb = ws.get("lang.sfb", [B * L, sent_feat.shape[-1]])
            ops.cast_bf16(sent_feat.view(B * L, -1), sfb)
        self._mlp_map_fwd("map_sentence", self.map_sentence, m.map_sentence, sfb, B * L, sfb.shape[1], y32=x32, yb=xb, ypb=xpb,
                          pos32=pos32, rowmap=(L, S, 0))
        plb = ws.get("lang.plb", [B * n_ph, pooled.shape[-1]])
        ops.cast_bf16(pooled.view(B * n_ph, -1), plb)
        ph32 = ws.get("lang.ph32", [B * n_ph, D], torch.float32)
        self._mlp_map_fwd("map_phrase", self.map_phrase, m.map_phrase, plb, B * n_ph, plb.shape[1], y32=ph32)
        # ---- encoder ------------------------------------------------------------------------------------------------
        for l, e in enumerate(self.enc):
            x32, xb, xpb = self._enc_fwd(l, e, x32, xb, xpb, kpm, pos32, B, S)
        mem32, memb, mempb = x32, xb, xpb
        # ---- query encoder + decoder + box head -------------------------------------------------------------------------
        with self._fork():
            self._dec_kv_fwd(memb, mempb, B * S)
        tgt32, tgtb, qpos32, tqb = self._qenc_fwd(mem32, ph32, mctx, B, S, L, n_ph, n_q)
        hs32, hsb = self._dec_fwd(tgt32, tgtb, qpos32, tqb, memb, mempb, kpm, qmask, B, S, T)
        nl = len(self.dec)
        rh = nl * B * T
        z0 = ws.get("bbox.z0", [rh, D])
        z1 = ws.get("bbox.z1", [rh, D])
        logits = ws.get("bbox.logits", [rh, 64], torch.float32)
        ops.gemm(hsb, self.bbox[0].wb, rh, D, D, bias=self.bbox[0].bias, relu=True, out=z0)
        ops.gemm(z0, self.bbox[1].wb, rh, D, D, bias=self.bbox[1].bias, relu=True, out=z1)
        ops.gemm(z1, self.bbox[2].wb, rh, 64, D, bias=self.bbox[2].bias, out32=logits)
        self.dims = (B, H, W, h, w, L, S, T, n_ph, n_q, lang[0] == "ids", len(lang) > 3)
        self.saved["top"] = (feats, c5, g5, pos32, kpm, mctx, qmask, proj32, gmean, grstd, mem32, memb, mempb, hs32, hsb, z0, z1)
        outs = [logits[:, :4].reshape(nl, B, n_ph, n_q, 4)]
        if want_seg:
            outs += self.seghead.forward(feats, proj32, mem32, memb, hs32, hsb, kpm, B, h, w, L, S, T)
        return outs

    def backward(self, g_logits, g_masks=None, g_att=None, part=None):
        """Returns (d_sent_feat [B, L, 768], d_pooled [B*n_ph, 768]); parameter gradients are left in ``self.gflat``.
        ``part`` runs one third of the pass only (``_run_backward_split``: three CUDA graphs, so that the gradient all-reduce of
        the language backbone overlaps the conv backbone's backward): "heads" = box / mask heads, decoder, query encoder, encoder,
        map_sentence; "bert" = the language backbone; "backbone" = input_proj + ResNet.  None = everything, BERT on a branch."""
        if part is not None and part != "heads":
            return self._backward_tail(part)
        m = self.model
        ws = self.ws
        vt = m.vl_transformer
        B, H, W, h, w, L, S, T, n_ph, n_q, native_bert, has_phrases = self.dims
        feats, c5, g5, pos32, kpm, mctx, qmask, proj32, gmean, grstd, mem32, memb, mempb, hs32, hsb, z0, z1 = self.saved["top"]
        rows, rt = B * S, B * T
        nl = len(self.dec)
        rh = nl * rt
        # (the 607 MB accumulation buffer is NOT zero-filled here, on the backward's critical path: run_backward clears it on a side
        # stream right after the hand-over of the previous step, where it overlaps the next forward -- see _clear_gflat)
        # ---- box head (backbone.py:35-38) ---------------------------------------------------------------------------------
        bl = m.bbox_embed.layers
        gl = ws.get("bboxb.gl", [rh, 64], torch.float32, zero=True)
        gl[:, :4].copy_(g_logits.reshape(rh, 4))
        glb = ws.get("bboxb.glb", [rh, 64])
        ops.cast_bf16(gl, glb)
        self.wgrad_linear(glb, z1, self.G(bl[2].weight), 4, D, rh, bias=self.G(bl[2].bias))
        dz1 = ws.get("bboxb.dz1", [rh, D])
        ops.gemm(glb, self.bbox[2].wt, rh, D, 64, mask_src=z1, out=dz1)
        self.wgrad_linear(dz1, z0, self.G(bl[1].weight), D, D, rh, bias=self.G(bl[1].bias))
        dz0 = ws.get("bboxb.dz0", [rh, D])
        ops.gemm(dz1, self.bbox[1].wt, rh, D, D, mask_src=z0, out=dz0)
        self.wgrad_linear(dz0, hsb, self.G(bl[0].weight), D, D, rh, bias=self.G(bl[0].bias))
        d_hs = ws.get("bboxb.d_hs", [rh, D], torch.float32)
        ops.gemm(dz0, self.bbox[0].wt, rh, D, D, out32=d_hs)
        g_mem = ws.get("bwd.g_mem", [rows, D], torch.float32)
        dpos = ws.get("bwd.dpos", [rows, D], torch.float32)
        g_fpn = {}
        g_src = None
        if g_masks is not None:
            # segmentation head: adds into d_hs (last layer), returns d(memory visual rows), d(input_proj out), FPN grads
            g_mem_seg, g_src, g_fpn = self.seghead.backward(g_masks, g_att, d_hs)
        # ---- decoder + query encoder ----------------------------------------------------------------------------------------
        d_tgt, d_qpos = self._dec_bwd(d_hs, memb, mempb, kpm, qmask, g_mem, dpos, B, S, T)
        if g_masks is not None:
            ops.rows_add(g_mem, g_mem_seg, rows, D, y32=g_mem)
        d_ph = self._qenc_bwd(d_tgt, d_qpos, g_mem, mctx, B, S, L, n_ph, n_q)
        d_pooled = self._mlp_map_bwd("map_phrase", self.map_phrase, m.map_phrase, d_ph)
        # ---- encoder --------------------------------------------------------------------------------------------------------
        gbuf = [ws.get("bwd.gA", [rows, D], torch.float32), ws.get("bwd.gB", [rows, D], torch.float32)]
        g = g_mem
        for l in reversed(range(len(self.enc))):
            g_out = gbuf[l & 1]
            self._enc_bwd(l, self.enc[l], g, g_out, kpm, dpos, B, S)
            g = g_out
        with self._off():
            ops.embed_grad(dpos, B, S, L, self.G(vt.lang_pos_embeddings.weight), self.G(vt.token_type_embeddings.weight), self.G(vt.level_embed))
        # ---- language rows -> map_sentence; visual rows -> GroupNorm -> input_proj -> backbone ---------------------------------------
        d_sent = self._mlp_map_bwd("map_sentence", self.map_sentence, m.map_sentence, g)
        self._bw = (g, g_src, g_fpn, d_sent, d_pooled)
        if part == "heads":
            self._join_side()
            return None
        return self._backward_tail(None)

    def _backward_tail(self, part):
        """part None: everything after the heads (BERT on a branch stream); "bert": the language backbone only; "bb:<k>": the conv
        backbone's layer k (the first such part also runs GroupNorm / input_proj backward and forms the gradient at C5)."""
        m = self.model
        ws = self.ws
        B, H, W, h, w, L, S, T, n_ph, n_q, native_bert, has_phrases = self.dims
        feats, c5, g5, pos32, kpm, mctx, qmask, proj32, gmean, grstd, mem32, memb, mempb, hs32, hsb, z0, z1 = self.saved["top"]
        g, g_src, g_fpn, d_sent, d_pooled = self._bw
        if part is not None and part.startswith("bb:"):
            layer = int(part[3:])
            layers = sorted({b.layer for b in self.blocks if b.trainable}, reverse=True)
            with self._side_category("bb"):
                if not layers or layer == layers[0]:
                    self._bb_gy = self._iproj_bwd(g, g_src, g_fpn)
                if layers and self._bb_gy is not None:
                    self._bb_gy = self._backbone_bwd(self._bb_gy, g_fpn, only_layer=layer)
            self._join_side()
            return None
        is_bert = part is not None and part.startswith("bert")
        if native_bert and (part is None or is_bert):  # BERT's backward chain runs next to the backbone backward (independent of it)
            import contextlib
            layers = None
            if is_bert and ":" in part:  # "bert:<hi>:<lo>": encoder layers hi-1 .. lo only (split backward)
                layers = tuple(int(v) for v in part.split(":")[1:])
            with (self._branch(urgent=True) if part is None else contextlib.nullcontext()), self._side_category("bert"):
                if has_phrases:
                    self.bert.backward("p", None, d_pooled, layers)
                    self.bert.backward("s", d_sent, None, layers)
                else:
                    self.bert.backward("s", d_sent, d_pooled, layers)
        if is_bert:
            self._join_side()
            return None
        with self._side_category("bb"):
            g5y = self._iproj_bwd(g, g_src, g_fpn)
            if g5y is not None:
                self._backbone_bwd(g5y, g_fpn)
        self._join_branch()
        self._join_side()
        return d_sent.view(B, L, -1), d_pooled

    def _iproj_bwd(self, g, g_src, g_fpn):
        """GroupNorm + input_proj backward (reftr_transformer.py:172-175); returns the masked gradient at the C5 output, or None when
        the backbone is frozen."""
        m = self.model
        ws = self.ws
        B, H, W, h, w, L, S, T, n_ph, n_q, native_bert, has_phrases = self.dims
        feats, c5, g5, pos32, kpm, mctx, qmask, proj32, gmean, grstd = self.saved["top"][:10]
        gn = m.input_proj[0][1]
        dproj = ws.get("iproj.dx", [g5.R, D], zero=True)
        ops.groupnorm_tokens_bwd(g, g_src, proj32, gn.weight, gmean, grstd, B, h, w, S, L, dproj, self.G(gn.weight), self.G(gn.bias))
        self.wgrad_conv(self.iproj, dproj, c5, g5.R, bias=self.G(m.input_proj[0][0].bias))
        if not self.blocks[-1].trainable:
            return None
        g5y = ws.get("iproj.gc5", [g5.R, 2048])
        ops.gemm(dproj, self.iproj.wd, g5.R, 2048, D, res=g_fpn.get(4), mask_src=c5, out=g5y)
        return g5y


def body_is_7x7(stem):
    return stem.kh == 7 and stem.kw == 7 and stem.Cin == 3 and stem.Cout == 64


class HotPathFunction(torch.autograd.Function):
    """The single autograd node of the hot path.  Inputs that carry gradient: BERT's sentence features and pooled
    phrase features (when BERT runs as the HuggingFace module), and every trainable hot-path parameter (so DistributedDataParallel's
    hooks fire as usual)."""

    @staticmethod
    def forward(ctx, eng, want_seg, sent_feat, pooled, *params):
        """The GPU work of the forward was ALREADY enqueued by ``RefTREngine.run_forward`` (RefTR._hot_path calls it first): autograd's
        bookkeeping for ~470 parameter inputs costs 0.8 ms of host time per call, which now overlaps the forward graph on the GPU
        instead of delaying its launch (it matters whenever the training loop synchronises every step, as engine_vg.py:53 does)."""
        ctx.eng = eng
        ctx.n_params = len(params)
        ctx.want_seg = want_seg
        ctx.step = eng.step_id
        return tuple(o.clone() for o in eng._cur["outs"])

    @staticmethod
    def backward(ctx, *gouts):
        eng = ctx.eng
        if ctx.step != eng.step_id:
            raise RuntimeError("reftr_b200: backward() of a stale forward (the engine keeps activations of the LAST forward only)")
        g_logits = gouts[0]
        g_masks = gouts[1] if ctx.want_seg else None
        g_att = gouts[2] if ctx.want_seg else None
        if ctx.want_seg and g_masks is None:
            g_masks = torch.zeros_like(eng.seghead.out_masks)
        import time as _t
        _t0 = _t.perf_counter()
        d_sent, d_pooled, flat = eng.run_backward(g_logits, g_masks, g_att)
        _t1 = _t.perf_counter()
        grads = []
        for n, p in eng.named:
            off, shape = eng.slots[n]
            grads.append(flat[off:off + p.numel()].view(shape))
        eng.host_ms.update({"run_backward": (_t1 - _t0) * 1e3, "views": (_t.perf_counter() - _t1) * 1e3})
        return (None, None, d_sent, d_pooled, *grads)
