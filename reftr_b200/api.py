"""Drop-in for ``models.build_reftr`` (models/__init__.py:4-11 -> reftr_transformer.py:307-347 / reftr_segmentation.py:343-384).

``build_reftr(args) -> (model, criterion, postprocessors)`` takes the argparse namespace of main_vg.py:26-164 unchanged.
The model is the B200 module (modules.py); criterion and post-processors are the reference's own classes when the
reference package is importable (the drop-in case: main_vg.py runs with the reference on sys.path), otherwise the
restatements in reftr_b200/criterion.py (stand-alone use: bench.py, tests).
"""
import os

import torch

from .modules import BackboneParams, Joiner, PositionEmbeddingSine, RefTR, RefTRSeg, VLTransformerParams


def build_backbone(args):  # backbone.py:148-154, position_encoding.py:87-97
    if args.position_embedding not in ("v2", "sine"):
        raise NotImplementedError("reftr_b200 builds the sine position embedding (every shipped config); got "
                                  f"--position_embedding {args.position_embedding}")
    if getattr(args, "dilation", False):
        raise NotImplementedError("--dilation is not built (no shipped config uses it)")
    return_interm = bool(args.masks) or args.num_feature_levels > 1
    backbone = BackboneParams(args.backbone, args.lr_backbone > 0, return_interm)
    return Joiner(backbone, PositionEmbeddingSine(args.hidden_dim // 2))


def build_vl_transformer(args):  # reftr.py:140-152
    return VLTransformerParams(args.hidden_dim, args.nheads, args.enc_layers, args.dec_layers, args.dim_feedforward, args.dropout,
                               args.num_feature_levels, args.max_lang_seq)


def load_lang_backbone(name):
    """reftr_transformer.py:315-318.  Offline boxes have no HF cache: REFTR_B200_RANDOM_BERT=1 builds a random-init
    BERT-base of the same architecture instead (synthetic-weight benchmarking only)."""
    from transformers import BertConfig, BertModel, RobertaModel
    cls = RobertaModel if name.split("-")[0] == "roberta" else BertModel
    if os.environ.get("REFTR_B200_RANDOM_BERT") == "1":
        torch.manual_seed(1234)
        layers = int(os.environ.get("REFTR_B200_RANDOM_BERT_LAYERS", "12"))  # (tests shrink the random BERT to keep CPU runs short)
        return BertModel(BertConfig(num_hidden_layers=layers))
    return cls.from_pretrained(name)


def _weight_dict(args, keys):
    wd = dict(keys)
    if args.aux_loss:  # reftr_transformer.py:324-329
        aux = {}
        for i in range(args.dec_layers - 1):
            aux.update({k + f"_{i}": v for k, v in wd.items()})
        aux.update({k + "_enc": v for k, v in wd.items()})
        wd.update(aux)
    return wd


def _reference_classes():
    """Criterion / post-processor classes.  Default: the restatements in reftr_b200/criterion.py (same constructor, weight_dict,
    loss keys and values as the reference's; sync-free and with all box losses in one kernel).  REFTR_B200_REF_CRITERION=1 uses
    the reference's own classes when its ``models`` package is importable."""
    if os.environ.get("REFTR_B200_REF_CRITERION") != "1":
        from .criterion import CriterionVGMultiPhrase, CriterionVGOnePhraseSeg, PostProcessSegm, PostProcessVGMultiPhrase
        return CriterionVGMultiPhrase, PostProcessVGMultiPhrase, CriterionVGOnePhraseSeg, PostProcessSegm
    try:
        from models.criterion import CriterionVGMultiPhrase
        from models.post_process import PostProcessVGMultiPhrase
        from models.reftr_segmentation import CriterionVGOnePhraseSeg, PostProcessSegm
        return CriterionVGMultiPhrase, PostProcessVGMultiPhrase, CriterionVGOnePhraseSeg, PostProcessSegm
    except Exception:
        from .criterion import CriterionVGMultiPhrase, CriterionVGOnePhraseSeg, PostProcessSegm, PostProcessVGMultiPhrase
        return CriterionVGMultiPhrase, PostProcessVGMultiPhrase, CriterionVGOnePhraseSeg, PostProcessSegm


def build_reftr(args):
    if not args.reftr_type.startswith("transformer"):
        raise NotImplementedError  # models/__init__.py:10-11
    device = torch.device(args.device)
    if getattr(args, "no_decoder", False):
        args.dec_layers = 0
    Crit, Post, CritSeg, PostSeg = _reference_classes()
    img_backbone = build_backbone(args)
    vl_transformer = build_vl_transformer(args)
    if args.masks:
        if args.reftr_type != "transformer_single_phrase":
            raise NotImplementedError  # reftr_segmentation.py:380-381
        wd = _weight_dict(args, {"loss_giou": args.giou_loss_coef, "loss_bbox": args.bbox_loss_coef, "loss_dice": args.dice_loss_coef,
                                 "loss_mask": args.mask_loss_coef, "loss_cem": 1.0})
        model = RefTRSeg(img_backbone, load_lang_backbone(args.bert_model), vl_transformer, num_feature_levels=args.num_feature_levels,
                         num_queries_per_phrase=args.num_queries_per_phrase, freeze_reftr=False, cem_loss=args.ablation == "cem_loss")
        criterion = CritSeg(wd, losses=["masks", "boxes"])
        post = {"bbox": Post(), "segm": PostSeg()}
    else:
        wd = _weight_dict(args, {"loss_giou": args.giou_loss_coef, "loss_bbox": args.bbox_loss_coef})
        model = RefTR(img_backbone, load_lang_backbone(args.bert_model), vl_transformer, num_feature_levels=args.num_feature_levels,
                      num_queries_per_phrase=args.num_queries_per_phrase, freeze_lang_backbone=args.freeze_bert, aux_loss=args.aux_loss,
                      ablation=args.ablation)
        criterion = Crit(wd, losses=["boxes"])
        post = {"bbox": Post()}
    criterion.to(device)
    return model, criterion, post
