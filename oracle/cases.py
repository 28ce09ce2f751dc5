"""Parity cases shared by oracle/make_golden.py (reference side) and tests/ (oracle + CUDA side).

TEST INFRASTRUCTURE (see oracle/reftr_oracle.py header).  ``flags`` are the reference's own argparse
flags (main_vg.py:26-164); ``oracle_kw`` are the equivalent constructor arguments of the oracle.
BERT is third-party in the reference; the small cases use a 2-layer BertConfig to keep the CPU suite fast
(identical object on both sides, so parity is unaffected).
"""
from transformers import BertConfig

_COMMON = ["--num_feature_levels", "1", "--aux_loss", "--dropout", "0.0"]


def _gf(name):  # gradients stored in the fixture: a spread of tensors from every stage
    keys = ("layer2.0.conv1.weight", "layer2.0.downsample.0.weight", "layer3.1.conv2.weight", "layer4.2.conv3.weight",
            "input_proj.0.0.weight", "input_proj.0.1.weight", "encoder.layers.0.self_attn.in_proj_weight",
            "encoder.layers.0.linear1.weight", "encoder.layers.0.norm2.bias", "decoder.layers.0.multihead_attn.in_proj_weight",
            "decoder.layers.0.self_attn.out_proj.weight", "decoder.norm.weight", "level_embed", "lang_pos_embeddings.weight",
            "token_type_embeddings.weight", "query_encoder.query_embed.weight", "query_encoder.linear2.weight",
            "query_encoder.fuse_encoder_query.0.weight", "map_sentence.0.weight", "map_phrase.4.weight",
            "bbox_embed.layers.0.weight", "bbox_embed.layers.2.bias", "lang_backbone.embeddings.word_embeddings.weight",
            "lang_backbone.encoder.layer.0.attention.self.query.weight", "lang_backbone.pooler.dense.weight",
            "bbox_attention.q_linear.weight", "bbox_attention.k_linear.weight", "mask_head.lay1.weight",
            "mask_head.gn3.weight", "mask_head.adapter2.weight", "mask_head.out_lay.weight")
    return any(name.endswith(k) for k in keys)


CASES = {
    # BASELINE.json configs[0] analogue: the reference refuses ResNet-18 (backbone.py:122), so R50 is used.
    "cfg1_box": dict(
        flags=_COMMON + ["--enc_layers", "1", "--dec_layers", "1"],
        oracle_kw=dict(enc=1, dec=1, dropout=0.0, aux_loss=True), seg=False, bert_layers=2, wseed=0,
        inputs=dict(B=2, H=224, W=224, L=8), grad_filter=_gf),
    # padding in the image (key-padding mask + position encoding) and in the sentence
    "pad_box": dict(
        flags=_COMMON + ["--enc_layers", "2", "--dec_layers", "2"],
        oracle_kw=dict(enc=2, dec=2, dropout=0.0, aux_loss=True), seg=False, bert_layers=2, wseed=3,
        inputs=dict(B=2, H=192, W=256, L=12, n_valid=7, pad_frac=0.25), grad_filter=_gf),
    # Flickr-style multi-phrase input (reftr_transformer.py:206-238), last phrase empty
    "multi_phrase": dict(
        flags=_COMMON + ["--enc_layers", "1", "--dec_layers", "2", "--reftr_type", "transformer"],
        oracle_kw=dict(enc=1, dec=2, dropout=0.0, aux_loss=True), seg=False, bert_layers=2, wseed=5,
        inputs=dict(B=2, H=160, W=160, L=16, n_valid=12, n_ph=3), grad_filter=_gf),
    # BASELINE.json configs[4] architecture: ResNet-101 backbone (23 layer3 blocks), longer phrase
    "r101_box": dict(
        flags=_COMMON + ["--enc_layers", "1", "--dec_layers", "1", "--backbone", "resnet101"],
        oracle_kw=dict(enc=1, dec=1, dropout=0.0, aux_loss=True, backbone="resnet101"), seg=False, bert_layers=2, wseed=9,
        inputs=dict(B=1, H=160, W=128, L=40, n_valid=33), grad_filter=_gf),
    # segmentation model (reftr_segmentation.py)
    "seg": dict(
        flags=_COMMON + ["--enc_layers", "1", "--dec_layers", "1", "--masks"],
        oracle_kw=dict(enc=1, dec=1, dropout=0.0), seg=True, bert_layers=2, wseed=7,
        inputs=dict(B=2, H=160, W=192, L=8, n_valid=6), grad_filter=_gf),
}


def bert_config(case):
    return BertConfig(num_hidden_layers=case["bert_layers"])


def build_oracle(case):
    """Oracle instance with the case's by-name synthetic weights, in eval mode (dropout inactive)."""
    import torch
    from transformers import BertModel
    from oracle.reftr_oracle import RefTROracle, RefTRSegOracle
    from reftr_b200.synthetic import synthetic_weights
    torch.manual_seed(1234)
    bert = BertModel(bert_config(case))
    cls = RefTRSegOracle if case["seg"] else RefTROracle
    model = cls(bert, **case["oracle_kw"])
    synthetic_weights(model, seed=case["wseed"])
    return model.eval()
