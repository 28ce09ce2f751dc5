"""The flags of the reference's launcher that shape the hot path (main_vg.py:26-164), with the reference's names and
defaults, so that ``build_reftr(parse(flags))`` accepts the flag lists of configs/*/*.sh.  Flags that belong to the parts
of the reference outside the hot path (optimizer, data, checkpoints) are accepted and ignored via parse_known_args."""
import argparse

_FLAGS = [
    # name, default, type
    ("lr_backbone", 1e-5, float), ("reftr_type", "transformer_single_phrase", str), ("ablation", "none", str),
    ("backbone", "resnet50", str), ("position_embedding", "sine", str), ("num_feature_levels", 4, int),
    ("enc_layers", 6, int), ("dec_layers", 6, int), ("dim_feedforward", 2048, int), ("hidden_dim", 256, int),
    ("dropout", 0.1, float), ("nheads", 8, int), ("bert_model", "bert-base-uncased", str), ("max_lang_seq", 128, int),
    ("num_queries_per_phrase", 1, int), ("mask_loss_coef", 1.0, float), ("dice_loss_coef", 1.0, float),
    ("bbox_loss_coef", 1.0, float), ("giou_loss_coef", 1.0, float), ("device", "cuda", str), ("batch_size", 8, int),
]
_SWITCHES = ["no_decoder", "dilation", "masks", "freeze_reftr", "freeze_bert", "aux_loss"]


def get_args_parser():
    p = argparse.ArgumentParser("reftr_b200 (hot-path flags of RefTR's main_vg.py)", add_help=False)
    for name, default, typ in _FLAGS:
        p.add_argument("--" + name, default=default, type=typ)
    for name in _SWITCHES:
        p.add_argument("--" + name, action="store_true")
    return p


def parse(flags=()):
    args, _ = get_args_parser().parse_known_args(list(flags))
    return args


# BASELINE.json configs, as flag lists in the style of configs/refcoco/*.sh
CONFIG_FLAGS = {
    "cfg2_box_r50": ["--num_feature_levels", "1", "--dec_layers", "6", "--aux_loss", "--batch_size", "16"],
    "cfg3_seg_r50": ["--num_feature_levels", "1", "--dec_layers", "6", "--masks", "--batch_size", "8"],
    "cfg4_flickr": ["--num_feature_levels", "1", "--dec_layers", "6", "--aux_loss", "--reftr_type", "transformer", "--batch_size", "32"],
    "cfg5_r101": ["--num_feature_levels", "1", "--dec_layers", "6", "--aux_loss", "--backbone", "resnet101", "--batch_size", "64"],
}
