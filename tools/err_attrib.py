"""Which 16-bit buffers of the forward carry the box error?  CPU, emulated kernels (tests/emu_ops.py rounds wherever the engine stores
a 16-bit workspace buffer): the cfg2 model at a reduced size, with the buffers whose name matches a pattern kept in fp32 instead.
Prints memory / decoder-output / box rel-L2 against the fp32 oracle for each pattern.  A diagnostic; not part of the product."""
import fnmatch
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
os.environ.setdefault("REFTR_B200_RANDOM_BERT", "1")
import torch
import emu_ops
import reftr_b200.engine as E, reftr_b200.bert as BM, reftr_b200.pack as PK, reftr_b200.modules as M, reftr_b200.criterion as C
for mod in (E, BM, PK, M, C):
    if hasattr(mod, "ops"):
        mod.ops = emu_ops
try:
    import reftr_b200.seg as SG
    SG.ops = emu_ops
except ImportError:
    pass
from oracle.cases import CASES, build_oracle
from reftr_b200.synthetic import synthetic_samples
from util_build import build_candidate, rel_l2

case = dict(CASES["cfg1_box"])
case["oracle_kw"] = dict(enc=6, dec=6, dropout=0.0, aux_loss=True)
case["bert_layers"] = 12
inp = dict(B=int(os.environ.get("PB", "2")), H=int(os.environ.get("PH", "320")), W=int(os.environ.get("PH", "320")), L=20)
torch.set_num_threads(os.cpu_count())
oracle = build_oracle(case)
s = synthetic_samples(**inp)
with torch.no_grad():
    out_o = oracle(s)
lay_o = [a["pred_boxes"] for a in out_o["aux_outputs"]] + [out_o["pred_boxes"]]
PAT = [""]
_get0 = E.Workspace.get


def _get(self, name, shape, dtype=None, zero=False):
    if dtype is None and PAT[0] and any(fnmatch.fnmatch(name, p) for p in PAT[0].split(",")):
        dtype = torch.float32
    return _get0(self, name, shape, dtype, zero)


E.Workspace.get = _get
F32W = set()
_t16 = emu_ops.t16


def _wrap_refresh(cls):
    r0 = cls.refresh

    def refresh(self):
        if id(self) in F32W:
            emu_ops.t16 = lambda: torch.float32
            try:
                return r0(self)
            finally:
                emu_ops.t16 = _t16
        return r0(self)
    cls.refresh = refresh


for _c in (PK.PackedLinear, PK.PackedConv, PK.PackedStack):
    _wrap_refresh(_c)
patterns = sys.argv[1:] or ["", "*"]
for pat in patterns:
    cand = build_candidate(case)
    eng = cand.engine()
    wsel = ""
    if pat.startswith("W:"):      # fp32 packed WEIGHTS for a group of layers instead of fp32 activation buffers
        wsel, pat_b = pat[2:], ""
    else:
        pat_b = pat
    PAT[0] = pat_b
    groups = {"enc": [p_ for e in eng.enc for p_ in (e.inp, e.out, e.l1, e.l2)],
              "dec": [p_ for d in eng.dec for p_ in vars(d).values() if isinstance(p_, PK.PackedLinear)] + [eng.kstack, eng.vstack],
              "head": list(eng.bbox), "iproj": [eng.iproj],
              "bert": list(eng.bert.packs) if eng.bert is not None else [],
              "conv": [p_ for b in eng.blocks for p_ in (b.c1, b.c2, b.c3, getattr(b, "ds", None)) if p_ is not None] + [eng.stem],
              "map": list(eng.map_sentence) + list(eng.map_phrase)}
    chosen = set()
    for g in wsel.split("+"):
        if g == "all":
            chosen |= {id(p_) for p_ in eng.packs}
        elif g:
            chosen |= {id(p_) for p_ in groups[g]}
    F32W.clear(); F32W.update(chosen)
    with torch.no_grad():
        out_c = cand(s)
    lay_c = [a["pred_boxes"] for a in out_c["aux_outputs"]] + [out_c["pred_boxes"]]
    mem32 = eng.saved["top"][10]
    B = inp["B"]
    S = mem32.shape[0] // B
    mem_err = rel_l2(mem32.view(B, S, 256).transpose(0, 1), out_o["_memory"])
    print(f"fp32 buffers {pat or '(none)':40s} memory {mem_err:.3e}  boxes per layer " + " ".join(f"{rel_l2(a, b):.2e}" for a, b in zip(lay_c, lay_o)), flush=True)
