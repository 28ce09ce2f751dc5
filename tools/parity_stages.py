"""Where does the forward error come from?  Full-size cfg2 (B=4) on the GPU: rel-L2 of the candidate's intermediates (C5, encoder
memory, every decoder layer's normalised output, boxes) against the fp32 oracle (TF32 off)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
from transformers import BertConfig, BertModel
from oracle.reftr_oracle import RefTROracle
from reftr_b200.synthetic import synthetic_weights, synthetic_samples
from reftr_b200.modules import BackboneParams, Joiner, PositionEmbeddingSine, RefTR, VLTransformerParams

torch.backends.cuda.matmul.allow_tf32 = False
torch.backends.cudnn.allow_tf32 = False
B = int(os.environ.get("PB", "4"))
torch.manual_seed(1234)
oracle = RefTROracle(BertModel(BertConfig()), enc=6, dec=6, dropout=0.1, aux_loss=True)
synthetic_weights(oracle, seed=0)
oracle = oracle.cuda().eval()
torch.manual_seed(1234)
cand = RefTR(Joiner(BackboneParams("resnet50", True, False), PositionEmbeddingSine(128)), BertModel(BertConfig()),
             VLTransformerParams(256, 8, 6, 6, 2048, 0.1, 1, 128), aux_loss=True)
synthetic_weights(cand, seed=0)
cand = cand.cuda().eval()
s = synthetic_samples(B=B, H=640, W=640, L=20, device="cuda")
cap = {}
oracle.vl_transformer.decoder.register_forward_hook(lambda m, i, o: cap.__setitem__("hs", o.detach()))
oracle.img_backbone[0].register_forward_hook(lambda m, i, o: cap.__setitem__("bb", o))
for l, lay in enumerate(oracle.vl_transformer.encoder.layers):
    lay.register_forward_hook(lambda m, i, o, l=l: cap.__setitem__(f"enc{l}", o.detach()))
with torch.no_grad():
    out_o = oracle(s)
    os.environ["REFTR_B200_GRAPHS"] = "0"
    out_c = cand(s)
torch.cuda.synchronize()
eng = cand.engine()
feats, c5, g5, pos32, kpm, mctx, qmask, proj32, gmean, grstd, mem32, memb, mempb, hs32, hsb, z0, z1 = eng.saved["top"]
rel = lambda a, b: ((a.float() - b.float()).norm() / b.float().norm()).item()
bb = cap["bb"]
c5_o = (bb[0] if isinstance(bb, (tuple, list)) else bb)
c5_o = c5_o[-1] if isinstance(c5_o, (tuple, list)) else c5_o
c5_c = c5.view(B, g5.Hp, g5.Wp, 2048)[:, 1:-1, 1:-1, :].permute(0, 3, 1, 2)
print("C5 rel-L2", rel(c5_c, c5_o), "max|C5|", c5_o.abs().max().item())
S = mem32.shape[0] // B
for l in range(6):
    k = f"enc{l}"
    sv = eng.saved[k]
print("memory rel-L2", rel(mem32.view(B, S, 256).transpose(0, 1), out_o["_memory"]))
hs_o = cap["hs"]  # [nl, T, B, 256]
hs_c = hs32.view(6, B, -1, 256).transpose(1, 2)
for l in range(6):
    print(f"decoder layer {l}: hs rel-L2 {rel(hs_c[l], hs_o[l]):.3e}")
lay_c = [a["pred_boxes"] for a in out_c["aux_outputs"]] + [out_c["pred_boxes"]]
lay_o = [a["pred_boxes"] for a in out_o["aux_outputs"]] + [out_o["pred_boxes"]]
for l in range(6):
    print(f"boxes layer {l}: rel-L2 {rel(lay_c[l], lay_o[l]):.3e} max abs {(lay_c[l]-lay_o[l]).abs().max().item():.3e}")
# how much of the decoder error is the decoder's own 16-bit arithmetic?  run the ORACLE's decoder on the CANDIDATE's memory
# how much of the box error is the box head's own 16-bit arithmetic?  The ORACLE's fp32 head on the CANDIDATE's fp32 decoder outputs
with torch.no_grad():
    hb = oracle.bbox_embed(hs32.view(6, B, -1, 256)).sigmoid()
for l in range(6):
    print(f"boxes layer {l} with an fp32 head on the candidate's hs: rel-L2 {rel(hb[l].reshape(B, -1, 4), lay_o[l].reshape(B, -1, 4)):.3e}")
