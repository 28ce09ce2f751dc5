"""TEST INFRASTRUCTURE (run as a subprocess by tests/test_dropin_reference.py; needs /root/reference, so build container only).

Runs the REFERENCE's own, unmodified ``main_vg.main(args)`` (main_vg.py:167-430) -- argument parser, seeding, optimizer with the
four LR groups (:223-268), StepLR, ``engine_vg.train_one_epoch`` (engine_vg.py:22-78) with its ``data_prefetcher``,
``engine_vg.evaluate`` (:82-225), checkpoint save (:372-412) and ``--resume`` (:298-339) -- with ``models.build_reftr`` supplied by
``shim/models`` (= reftr_b200.build_reftr).  The data layer is out of scope (no dataset files exist): ``build_refer_dataset`` is
replaced by a synthetic dataset that yields items in the reference's own format (refer_dataset.py:186-194), which go through the
reference's own ``collate_fn_vg``.  There is no GPU here, so the C-ABI kernels are replaced by tests/emu_ops.py (as in
tests/test_engine_emulated.py) and the CUDA stream calls of ``data_prefetcher`` by no-ops.
Prints one JSON line with what the test asserts on."""
import argparse
import contextlib
import importlib.util
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("REFTR_REF", "/root/reference")
sys.path[:0] = [os.path.join(ROOT, "shim"), ROOT, os.path.join(ROOT, "tests"), REF]
os.environ["REFTR_B200_RANDOM_BERT"] = "1"
os.environ.setdefault("REFTR_B200_RANDOM_BERT_LAYERS", "2")
os.environ["REFTR_B200_GRAPHS"] = "0"

import reftr_compat  # noqa: E402

reftr_compat.install()
import torch  # noqa: E402

torch.set_num_threads(os.cpu_count())

# ---- the kernels: CPU emulation (test infrastructure) -----------------------------------------------------------------------
import emu_ops  # noqa: E402
import reftr_b200.bert  # noqa: E402
import reftr_b200.engine  # noqa: E402
import reftr_b200.pack  # noqa: E402
import reftr_b200.seg  # noqa: E402

for _m in (reftr_b200.engine, reftr_b200.pack, reftr_b200.bert, reftr_b200.seg):
    _m.ops = emu_ops

# ---- CUDA stream API used by engine_vg.data_prefetcher (engine_vg.py:240-283): no-ops on a CPU-only machine -------------------
if not torch.cuda.is_available():
    class _NoStream:
        def __init__(self, *a, **k):
            pass

        def wait_stream(self, other):
            pass
    torch.cuda.Stream = _NoStream
    torch.cuda.current_stream = lambda *a, **k: _NoStream()
    torch.cuda.stream = lambda s: contextlib.nullcontext()
    torch.Tensor.record_stream = lambda self, s: None


class SyntheticReferDataset(torch.utils.data.Dataset):
    """Items in the format of datasets/grounding_datasets/refer_dataset.py:186-194 (single-phrase RES/REC configs)."""

    def __init__(self, split, n, H, W, L, seg=False):
        self.split, self.n, self.H, self.W, self.L, self.seg = split, n, H, W, L, seg

    def __len__(self):
        return self.n

    def __getitem__(self, i):
        g = torch.Generator().manual_seed(1000 + i)
        nv = 5 + (i % 3)
        sent = torch.zeros(self.L, dtype=torch.long)
        sent[:nv] = torch.randint(1000, 30000, (nv,), generator=g)
        sent[0], sent[nv - 1] = 101, 102
        smask = torch.zeros(self.L, dtype=torch.long)
        smask[:nv] = 1
        sample = {"img": torch.randn(3, self.H, self.W, generator=g), "sentence": sent, "sentence_mask": smask}
        c = 0.25 + 0.5 * torch.rand(1, 2, generator=g)
        wh = 0.1 + 0.3 * torch.rand(1, 2, generator=g)
        target = {"boxes": torch.cat([c, wh], -1), "labels": torch.zeros(1, dtype=torch.long),
                  "orig_size": torch.tensor([self.H, self.W]), "size": torch.tensor([self.H, self.W]),
                  "dataset_id": torch.tensor(i), "image_id": torch.tensor(i)}
        if self.seg:
            m = torch.zeros(1, self.H, self.W, dtype=torch.bool)
            m[:, self.H // 4: self.H // 2 + i, self.W // 4: self.W // 2 + 2 * i] = True
            target["masks"] = m
        return sample, target


def load_reference_main():
    spec = importlib.util.spec_from_file_location("reference_main_vg", os.path.join(REF, "main_vg.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def main():
    out_dir, masks = sys.argv[1], len(sys.argv) > 2 and sys.argv[2] == "masks"
    ref = load_reference_main()
    import models
    assert models.build_reftr.__module__ == "reftr_b200.api", models.build_reftr.__module__
    assert ref.build_reftr is models.build_reftr
    assert ref.train_one_epoch.__code__.co_filename == os.path.join(REF, "engine_vg.py")
    H = W = 64
    ref.build_refer_dataset = lambda split, args: SyntheticReferDataset(split, 4 if split == "trainval" else 2, H, W, 8, seg=bool(args.masks))
    parser = argparse.ArgumentParser(parents=[ref.get_args_parser()])
    flags = ["--device", "cpu", "--batch_size", "2", "--num_workers", "0", "--num_feature_levels", "1", "--enc_layers", "1",
             "--dec_layers", "2", "--epochs", "1", "--lr_drop", "1", "--ckpt_cycle", "1", "--output_dir", out_dir,
             "--lr", "1e-3", "--lr_backbone", "1e-4"] + (["--masks"] if masks else ["--aux_loss"])
    args = parser.parse_args(flags)
    # ---- run 1: one epoch of the reference's training loop + evaluation + checkpoint --------------------------------------------
    captured = {}
    real_build = models.build_reftr

    def spy_build(a):
        r = real_build(a)
        captured["model"] = r[0]
        captured["init"] = {k: v.clone() for k, v in r[0].state_dict().items()}
        return r
    ref.build_reftr = spy_build
    evals = []
    real_eval = ref.evaluate

    def spy_eval(*a, **k):
        r = real_eval(*a, **k)
        evals.append(r[0])
        return r
    ref.evaluate = spy_eval
    ref.main(args)
    model = captured["model"]
    after = {k: v.clone() for k, v in model.state_dict().items()}
    changed = sum(1 for k, v in after.items() if v.is_floating_point() and not torch.equal(v, captured["init"][k]))
    frozen_ok = all(torch.equal(after[k], captured["init"][k]) for k in after if ".layer1." in k or k.endswith("body.conv1.weight"))
    ckpt = torch.load(os.path.join(out_dir, "checkpoint.pth"), map_location="cpu", weights_only=False)
    log = [json.loads(l) for l in open(os.path.join(out_dir, "log.txt"))]
    groups = [len(g["params"]) for g in ckpt["optimizer"]["param_groups"]]
    res = {"changed": changed, "frozen_ok": frozen_ok, "ckpt_keys": sorted(ckpt.keys()), "n_state": len(ckpt["model"]),
           "ckpt_matches_model": all(torch.equal(ckpt["model"][k], after[k]) for k in after), "log": log, "opt_groups": groups,
           "engine_launches": model.engine().launches}
    # ---- run 2: --resume (model + optimizer + scheduler), evaluation, one more epoch ----------------------------------------------
    args2 = parser.parse_args(flags + ["--resume", os.path.join(out_dir, "checkpoint.pth")])
    args2.epochs = 2
    captured.clear()
    ref.main(args2)
    model2 = captured["model"]
    log2 = [json.loads(l) for l in open(os.path.join(out_dir, "log.txt"))]
    ckpt2 = torch.load(os.path.join(out_dir, "checkpoint.pth"), map_location="cpu", weights_only=False)
    res.update({"resume_epoch": ckpt2["epoch"], "log2": log2[len(log):], "evals": evals,
                "resume_moved": sum(1 for k, v in model2.state_dict().items() if v.is_floating_point() and not torch.equal(v, after[k]))})
    print("DROPIN_RESULT " + json.dumps(res))


if __name__ == "__main__":
    main()
