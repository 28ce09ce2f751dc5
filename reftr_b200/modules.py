"""nn.Module surface of RefTR / RefTRSeg with the reference's exact parameter names and shapes, so that
``state_dict`` / ``load_state_dict`` / the LR groups of main_vg.py:223-262 / DistributedDataParallel work unchanged.

The sub-modules below are PARAMETER CONTAINERS: their own ``forward`` is never called.  ``RefTR.forward`` hands the
parameters to the CUDA engine (reftr_b200/engine.py) through one autograd.Function; BERT (third-party in the reference,
reftr_transformer.py:8) stays a HuggingFace module in the autograd graph.
"""
import torch
from torch import nn

from .engine import HotPathFunction, RefTREngine

RESNET_BLOCKS = {"resnet50": (3, 4, 6, 3), "resnet101": (3, 4, 23, 3)}


class FrozenBatchNorm2d(nn.Module):
    """Buffers only (backbone.py:43-68); folded into the packed conv weights, never executed on its own."""

    def __init__(self, n, eps=1e-5):
        super().__init__()
        self.register_buffer("weight", torch.ones(n))
        self.register_buffer("bias", torch.zeros(n))
        self.register_buffer("running_mean", torch.zeros(n))
        self.register_buffer("running_var", torch.ones(n))
        self.eps = eps

    def _load_from_state_dict(self, state_dict, prefix, *args):
        state_dict.pop(prefix + "num_batches_tracked", None)
        super()._load_from_state_dict(state_dict, prefix, *args)


class BottleneckParams(nn.Module):
    def __init__(self, cin, width, stride, has_ds):
        super().__init__()
        self.stride, self.cin, self.width = stride, cin, width
        self.conv1 = nn.Conv2d(cin, width, 1, bias=False)
        self.bn1 = FrozenBatchNorm2d(width)
        self.conv2 = nn.Conv2d(width, width, 3, stride=stride, padding=1, bias=False)
        self.bn2 = FrozenBatchNorm2d(width)
        self.conv3 = nn.Conv2d(width, width * 4, 1, bias=False)
        self.bn3 = FrozenBatchNorm2d(width * 4)
        if has_ds:
            self.downsample = nn.Sequential(nn.Conv2d(cin, width * 4, 1, stride=stride, bias=False), FrozenBatchNorm2d(width * 4))
        else:
            self.downsample = None


class ResNetParams(nn.Module):
    """torchvision ResNet-50/101 parameter tree (names: conv1, bn1, layer{1..4}.{i}.{conv,bn}{1..3}, downsample.{0,1})."""

    def __init__(self, name):
        super().__init__()
        if name not in RESNET_BLOCKS:
            raise NotImplementedError(f"backbone {name!r}: the reference asserts resnet50/101 (backbone.py:122); only those are built")
        self.conv1 = nn.Conv2d(3, 64, 7, stride=2, padding=3, bias=False)
        self.bn1 = FrozenBatchNorm2d(64)
        cin = 64
        for li, (width, nb) in enumerate(zip((64, 128, 256, 512), RESNET_BLOCKS[name])):
            blocks = []
            for bi in range(nb):
                stride = 2 if (bi == 0 and li > 0) else 1
                blocks.append(BottleneckParams(cin, width, stride, has_ds=(bi == 0)))
                cin = width * 4
            setattr(self, f"layer{li + 1}", nn.Sequential(*blocks))
        for m in self.modules():  # torchvision's default init
            if isinstance(m, nn.Conv2d):
                nn.init.kaiming_normal_(m.weight, mode="fan_out", nonlinearity="relu")


class BackboneParams(nn.Module):
    def __init__(self, name, train_backbone, return_interm_layers):
        super().__init__()
        self.body = ResNetParams(name)
        for pname, p in self.body.named_parameters():  # backbone.py:87-89
            if not train_backbone or ("layer2" not in pname and "layer3" not in pname and "layer4" not in pname):
                p.requires_grad_(False)
        self.return_interm_layers = return_interm_layers
        self.strides = [4, 8, 16, 32] if return_interm_layers else [32]
        self.num_channels = [256, 512, 1024, 2048] if return_interm_layers else [2048]


class PositionEmbeddingSine(nn.Module):
    """Parameter-free (position_encoding.py:20-56); computed inside rb_build_pos_mask."""

    def __init__(self, num_pos_feats):
        super().__init__()
        self.num_pos_feats = num_pos_feats


class Joiner(nn.Sequential):  # backbone.py:128-133
    def __init__(self, backbone, pos):
        super().__init__(backbone, pos)
        self.strides, self.num_channels = backbone.strides, backbone.num_channels


def mlp_mapping(i, o):  # same Sequential indices as reftr_transformer.py:14-23 (params at 0, 1, 4, 5)
    return nn.Sequential(nn.Linear(i, o), nn.LayerNorm(o), nn.ReLU(), nn.Dropout(0.1), nn.Linear(o, o), nn.LayerNorm(o), nn.ReLU())


class MLPParams(nn.Module):  # backbone.py:26-33
    def __init__(self, i, h, o, n):
        super().__init__()
        self.num_layers = n
        dims = [i] + [h] * (n - 1) + [o]
        self.layers = nn.ModuleList(nn.Linear(a, b) for a, b in zip(dims[:-1], dims[1:]))


class EncoderLayerParams(nn.Module):
    def __init__(self, d, h, dff, dropout):
        super().__init__()
        self.self_attn = nn.MultiheadAttention(d, h, dropout=dropout)
        self.linear1 = nn.Linear(d, dff)
        self.linear2 = nn.Linear(dff, d)
        self.norm1 = nn.LayerNorm(d)
        self.norm2 = nn.LayerNorm(d)


class DecoderLayerParams(nn.Module):
    def __init__(self, d, h, dff, dropout):
        super().__init__()
        self.self_attn = nn.MultiheadAttention(d, h, dropout=dropout)
        self.multihead_attn = nn.MultiheadAttention(d, h, dropout=dropout)
        self.linear1 = nn.Linear(d, dff)
        self.linear2 = nn.Linear(dff, d)
        self.norm1 = nn.LayerNorm(d)
        self.norm2 = nn.LayerNorm(d)
        self.norm3 = nn.LayerNorm(d)


class _Stack(nn.Module):
    def __init__(self, layers, norm=None):
        super().__init__()
        self.layers = nn.ModuleList(layers)
        if norm is not None:
            self.norm = norm


class VLTransformerParams(nn.Module):
    """reftr.py:10-49."""

    def __init__(self, d, h, enc, dec, dff, dropout, num_feature_levels, max_lang_seq):
        super().__init__()
        self.d_model, self.nhead, self.max_lang_seq, self.dropout = d, h, max_lang_seq, dropout
        self.lang_pos_embeddings = nn.Embedding(max_lang_seq, d)
        self.token_type_embeddings = nn.Embedding(2, d)
        self.level_embed = nn.Parameter(torch.Tensor(num_feature_levels, d))
        self.encoder = _Stack([EncoderLayerParams(d, h, dff, dropout) for _ in range(enc)])
        self.use_decoder = dec > 0
        if self.use_decoder:
            self.decoder = _Stack([DecoderLayerParams(d, h, dff, dropout) for _ in range(dec)], nn.LayerNorm(d))
        for p in self.parameters():
            if p.dim() > 1:
                nn.init.xavier_uniform_(p)
        nn.init.normal_(self.level_embed)


class QueryEncoderParams(nn.Module):  # reftr_transformer.py:26-39
    def __init__(self, n_q, d):
        super().__init__()
        self.hidden_dim = d
        self.query_embed = nn.Embedding(n_q, d * 2)
        self.linear1 = nn.Linear(d, d)
        self.linear2 = nn.Linear(d, d)
        self.linear3 = nn.Linear(d, d)
        self.fuse_encoder_query = mlp_mapping(d * 2, d)
        self.context_out = nn.Sequential(nn.Linear(d, d), nn.LayerNorm(d))


class MHAttentionMapParams(nn.Module):  # reftr_segmentation.py:178-194
    def __init__(self, d, h):
        super().__init__()
        self.num_heads, self.hidden_dim = h, d
        self.q_linear = nn.Linear(d, d)
        self.k_linear = nn.Linear(d, d)
        nn.init.zeros_(self.k_linear.bias)
        nn.init.zeros_(self.q_linear.bias)
        nn.init.xavier_uniform_(self.k_linear.weight)
        nn.init.xavier_uniform_(self.q_linear.weight)


class MaskHeadParams(nn.Module):  # reftr_segmentation.py:216-241
    def __init__(self, dim, fpn_dims, ctx):
        super().__init__()
        inter = [dim, ctx // 2, ctx // 4, ctx // 8, ctx // 16]
        self.lay1 = nn.Conv2d(dim, dim, 3, padding=1)
        self.gn1 = nn.GroupNorm(8, dim)
        self.lay2 = nn.Conv2d(dim, inter[1], 3, padding=1)
        self.gn2 = nn.GroupNorm(8, inter[1])
        self.lay3 = nn.Conv2d(inter[1], inter[2], 3, padding=1)
        self.gn3 = nn.GroupNorm(8, inter[2])
        self.lay4 = nn.Conv2d(inter[2], inter[3], 3, padding=1)
        self.gn4 = nn.GroupNorm(8, inter[3])
        self.lay5 = nn.Conv2d(inter[3], inter[4], 3, padding=1)
        self.gn5 = nn.GroupNorm(8, inter[4])
        self.out_lay = nn.Conv2d(inter[4], 1, 3, padding=1)
        self.adapter1 = nn.Conv2d(fpn_dims[0], inter[1], 1)
        self.adapter2 = nn.Conv2d(fpn_dims[1], inter[2], 1)
        self.adapter3 = nn.Conv2d(fpn_dims[2], inter[3], 1)
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                nn.init.kaiming_uniform_(m.weight, a=1)
                nn.init.constant_(m.bias, 0)


class RefTR(nn.Module):
    """Drop-in for models.reftr_transformer.RefTR (reftr_transformer.py:69-304)."""

    def __init__(self, img_backbone, lang_backbone, vl_transformer, num_feature_levels=1, num_queries_per_phrase=1,
                 freeze_lang_backbone=False, aux_loss=False, ablation="none"):
        super().__init__()
        if num_feature_levels != 1:
            raise NotImplementedError("reftr_b200 accelerates --num_feature_levels 1 (every shipped config, SURVEY.md 0.4)")
        if not vl_transformer.use_decoder:
            raise NotImplementedError("--no_decoder is not built")
        self.img_backbone = img_backbone
        self.lang_backbone = lang_backbone
        self.vl_transformer = vl_transformer
        self.num_feature_levels = num_feature_levels
        self.num_queries_per_phrase = num_queries_per_phrase
        self.hidden_dim = d = vl_transformer.d_model
        if d != 256 or d // vl_transformer.nhead != 32:
            raise NotImplementedError("kernels are built for hidden_dim 256 / head_dim 32 (all shipped configs)")
        self.bbox_embed = MLPParams(d, d, 4, 3)
        self.lang_hidden_dim = lang_backbone.config.hidden_size
        self.map_sentence = mlp_mapping(self.lang_hidden_dim, d)
        self.use_decoder = True
        self.map_phrase = mlp_mapping(self.lang_hidden_dim, d)
        self.query_encoder = QueryEncoderParams(num_queries_per_phrase, d)
        assert img_backbone.num_channels[-1] == 2048
        self.input_proj = nn.ModuleList([nn.Sequential(nn.Conv2d(2048, d, kernel_size=1), nn.GroupNorm(32, d))])
        self.aux_loss = aux_loss
        self.freeze_lang_backbone = freeze_lang_backbone
        nn.init.constant_(self.bbox_embed.layers[-1].weight.data, 0)  # reftr_transformer.py:131-135
        nn.init.constant_(self.bbox_embed.layers[-1].bias.data, 0)
        nn.init.xavier_uniform_(self.input_proj[0][0].weight, gain=1)
        nn.init.constant_(self.input_proj[0][0].bias, 0)
        self.tf32_bert = True
        self._engine = None
        self._mark_ddp_ignored()

    def _mark_ddp_ignored(self):
        """Data-parallel gradients: the engine produces every hot-path gradient in ONE flat fp32 buffer, so it all-reduces that
        buffer itself (one NCCL call, no bucket copies; engine.run_backward) instead of letting DistributedDataParallel re-bucket
        ~700 tensors.  DDP (main_vg.py:293-296) is told to skip those parameters through its documented-by-use
        ``_ddp_params_and_buffers_to_ignore`` attribute; one tiny parameter stays with DDP because it refuses a module without
        any (its second averaging of an already-averaged gradient is the identity)."""
        import os
        from .bert import BertEngine
        if os.environ.get("REFTR_B200_GRAD_ALLREDUCE", "engine") != "engine":
            self._ddp_params_and_buffers_to_ignore = []
            self.engine_allreduce = False
            return
        native_bert = os.environ.get("REFTR_B200_NATIVE_BERT", "1") != "0" and BertEngine.eligible(self.lang_backbone)
        keep = "vl_transformer.level_embed"
        self._ddp_params_and_buffers_to_ignore = [n for n, p in self.named_parameters()
                                                  if p.requires_grad and n != keep and (native_bert or not n.startswith("lang_backbone."))]
        # FrozenBatchNorm statistics never change (backbone.py:43-80) and are identical on every rank (same checkpoint): keep DDP's
        # per-forward buffer broadcast away from them -- it would bump their version counters and force a re-pack of every folded
        # convolution weight on every step.
        # (the same holds for the only other buffers of the model, BERT's constant position / token-type id tables)
        self._ddp_params_and_buffers_to_ignore += [n for n, _ in self.named_buffers()]
        self.engine_allreduce = True

    # -- checkpoint helpers of the reference -----------------------------------------------------------------
    def init_from_pretrained_detr(self, state_dict):  # reftr_transformer.py:137-146
        backbone = {k.split(".", 1)[1]: v for k, v in state_dict.items() if k.split(".", 1)[0] == "backbone"}
        encoder = {k.split(".", 2)[2]: v for k, v in state_dict.items() if "transformer.encoder" in k}
        self.img_backbone.load_state_dict(backbone)
        self.vl_transformer.encoder.load_state_dict(encoder)

    # -- forward ------------------------------------------------------------------------------------------------
    def engine(self):
        if self._engine is None:
            self._engine = RefTREngine(self)
        return self._engine

    def reset_engine(self):
        """Rebuilds the engine and the DDP ignore list after ``requires_grad`` flags were changed (e.g. freezing BERT by hand after
        construction).  Call it BEFORE wrapping the module in DistributedDataParallel."""
        self._engine = None
        self._mark_ddp_ignored()

    def _language(self, samples):
        """BERT + the phrase / context masks of reftr_transformer.py:197-248, vectorised (no host syncs)."""
        sentence, sentence_mask = samples["sentence"], samples["sentence_mask"]
        B, L = sentence.shape
        n_q = self.num_queries_per_phrase
        if self.engine().bert is not None:  # BERT runs inside the engine on the C-ABI kernels; only the masks are built here
            return self._language_masks(samples) + (None, None)
        if self.tf32_bert and not torch.backends.cuda.matmul.allow_tf32:
            # BERT is a third-party PyTorch module (reftr_transformer.py:8); its fp32 GEMMs run on the tensor cores in TF32,
            # forward AND backward (autograd runs the backward outside any scope, so the switch is process-wide).
            torch.backends.cuda.matmul.allow_tf32 = True
        lo = self.lang_backbone(sentence, token_type_ids=None, attention_mask=sentence_mask)
        sent_feat, pooled = lo[0], lo[1]
        if "phrase" in samples:
            ph, pm = samples["phrase"], samples["phrase_mask"]
            n_ph = ph.size(1)
            pooled = self.lang_backbone(ph.reshape(B * n_ph, -1), token_type_ids=None, attention_mask=pm.reshape(B * n_ph, -1))[1]
        return self._language_masks(samples) + (sent_feat, pooled)

    def _language_masks(self, samples):
        """phrase / context masks of reftr_transformer.py:206-248, vectorised (no host syncs)."""
        sentence, sentence_mask = samples["sentence"], samples["sentence_mask"]
        B, L = sentence.shape
        n_q = self.num_queries_per_phrase
        if "phrase" in samples:
            pm = samples["phrase_mask"]
            n_ph = samples["phrase"].size(1)
            ar = torch.arange(L, device=sentence.device).view(1, 1, L)
            inside = (ar >= samples["phrase_pos_l"].unsqueeze(-1)) & (ar < samples["phrase_pos_r"].unsqueeze(-1))
            mask_context = ~inside
            query_mask = pm.view(B, n_ph, -1)[:, :, 2:3].logical_not().expand(-1, -1, n_q).reshape(B, n_ph * n_q)
        else:
            n_ph = 1
            slen = sentence_mask.to(torch.int64).sum(-1)
            ar = torch.arange(L, device=sentence.device).view(1, L)
            mask_context = (sentence_mask.to(torch.bool).logical_not() | (ar == 0) | (ar == (slen - 1).view(B, 1))).view(B, 1, L)
            query_mask = torch.zeros((B, 1), dtype=torch.bool, device=sentence.device)
        return mask_context, query_mask, n_ph

    def _hot_path(self, samples, want_seg=False):
        img = samples["img"]
        if not hasattr(img, "decompose"):
            raise TypeError("samples['img'] must be a NestedTensor-like object with .tensors / .mask (util/misc.py:308)")
        tensors, mask = img.decompose()
        mask_context, query_mask, n_ph, sent_feat, pooled = self._language(samples)
        eng = self.engine()
        params = eng.param_list()
        ph_ids = samples.get("phrase") if isinstance(samples, dict) else None
        ph_mask = samples.get("phrase_mask") if ph_ids is not None else None
        # enqueue the GPU work first, then let autograd do its (host-side) bookkeeping while the GPU is already running
        eng.run_forward(tensors, mask, samples["sentence_mask"], mask_context, query_mask, n_ph, want_seg,
                        sent_feat.detach() if sent_feat is not None else None, pooled.detach() if pooled is not None else None,
                        samples["sentence"], ph_ids, ph_mask)
        outs = HotPathFunction.apply(eng, want_seg, sent_feat, pooled, *params)
        return outs, query_mask, n_ph

    def forward(self, samples):
        outs, query_mask, n_ph = self._hot_path(samples)
        logits = outs[0]  # [n_layers, B, n_ph, n_q, 4]
        coord = logits.sigmoid()
        pm = query_mask.logical_not()
        out = {"pred_boxes": coord[-1], "phrase_mask": pm}
        if self.aux_loss:
            out["aux_outputs"] = [{"pred_boxes": b, "phrase_mask": pm} for b in coord[:-1]]
            # (every layer's boxes are slices of ONE tensor: reftr_b200's criterion finds it through Tensor._base and computes all box
            # losses in one kernel; the dict holds exactly the reference's keys, reftr_transformer.py:293-304)
        return out


class RefTRSeg(RefTR):
    """Drop-in for models.reftr_segmentation.RefTRSeg (reftr_segmentation.py:44-175)."""

    def __init__(self, img_backbone, lang_backbone, vl_transformer, num_feature_levels=1, num_queries_per_phrase=1,
                 freeze_reftr=False, cem_loss=False):
        super().__init__(img_backbone, lang_backbone, vl_transformer, num_feature_levels, num_queries_per_phrase,
                         freeze_lang_backbone=False, aux_loss=False)
        if freeze_reftr:
            for p in self.parameters():
                p.requires_grad = False
        if cem_loss:
            raise NotImplementedError("--ablation cem_loss is not built")
        if num_queries_per_phrase != 1:
            raise NotImplementedError("segmentation is built for one query per image (reftr_segmentation.py:97)")
        d, h = self.hidden_dim, vl_transformer.nhead
        self.bbox_attention = MHAttentionMapParams(d, h)
        self.mask_head = MaskHeadParams(d * 2 + h, [1024, 512, 256], d)
        self.cem_loss = False
        self._mark_ddp_ignored()

    def init_from_pretrained(self, pretrained_state_dict):  # reftr_segmentation.py:66-74
        missing, unexpected = self.load_state_dict(pretrained_state_dict, strict=False)
        print("Unexpected keys: ", unexpected)
        print("Missing keys: ", missing)

    def forward(self, samples):
        assert "phrase" not in samples
        outs, query_mask, _ = self._hot_path(samples, want_seg=True)
        logits, pred_masks, mask_att = outs
        out = {"pred_boxes": logits[-1].sigmoid(), "phrase_mask": query_mask.logical_not()}
        out["pred_masks"] = pred_masks
        out["mask_att"] = mask_att
        return out
