"""The drop-in contract of SURVEY.md 8(b), exercised by the REFERENCE'S OWN driver: tests/dropin_driver.py runs the unmodified
``main_vg.main(args)`` of /root/reference (argument parser, the four LR groups + AdamW of main_vg.py:223-268, StepLR,
``engine_vg.train_one_epoch`` with its data_prefetcher, ``engine_vg.evaluate``, checkpoint save and ``--resume``) with
``models.build_reftr`` resolved through ``shim/models`` to this package.  CPU only (kernels emulated by tests/emu_ops.py); the GPU
box has no /root/reference, so the test skips there."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("REFTR_REF", "/root/reference")

pytestmark = pytest.mark.skipif(not os.path.isfile(os.path.join(REF, "main_vg.py")), reason="needs the reference checkout (build container only)")


def _run(tmp_path, *extra, env_extra=None):
    env = dict(os.environ)
    env.pop("PYTHONPATH", None)
    env.update(env_extra or {})
    p = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "dropin_driver.py"), str(tmp_path), *extra], capture_output=True, text=True,
                       timeout=900, env=env, cwd=str(tmp_path))
    assert p.returncode == 0, p.stdout[-3000:] + p.stderr[-3000:]
    line = [l for l in p.stdout.splitlines() if l.startswith("DROPIN_RESULT ")][-1]
    return json.loads(line[len("DROPIN_RESULT "):]), p.stdout


@pytest.mark.parametrize("criterion", ["repo", "reference"])
def test_reference_main_vg_trains_evaluates_and_resumes_with_this_model(tmp_path, criterion):
    """criterion = "reference": build_reftr hands main_vg the reference's own CriterionVGMultiPhrase / PostProcessVGMultiPhrase
    (REFTR_B200_REF_CRITERION=1; models/criterion.py prints "Using multi phrase loss" on construction)."""
    r, out = _run(tmp_path, env_extra={"REFTR_B200_REF_CRITERION": "1" if criterion == "reference" else "0"})
    assert ("Using multi phrase loss" in out) == (criterion == "reference")
    # epoch 0 of engine_vg.train_one_epoch: 2 iterations, finite losses for every key of the reference's weight_dict, parameters moved
    log0 = r["log"][0]
    for k in ("train_loss", "train_loss_bbox", "train_loss_giou", "train_loss_bbox_0", "train_loss_giou_0", "train_grad_norm", "test_accuracy_iou0.5", "test_miou"):
        assert k in log0 and log0[k] == log0[k] and abs(log0[k]) < 1e4, (k, log0)
    assert r["changed"] > 150 and r["frozen_ok"]          # conv1 / layer1 stay frozen (backbone.py:87-89), everything trainable moved
    assert r["opt_groups"][0] > 0 and r["opt_groups"][1] > 0 and r["opt_groups"][2] > 0   # default / img_backbone.0 / lang_backbone LR groups
    assert r["engine_launches"] > 500                      # the engine (not some fallback) did the work
    # checkpoint written by main_vg.py:372-385 holds exactly the module's state_dict
    assert r["ckpt_keys"] == ["args", "best_val_acc", "epoch", "lr_scheduler", "model", "optimizer"]
    assert r["ckpt_matches_model"]
    # --resume: the first evaluation of the resumed run reproduces the last evaluation of the first run bit for bit
    ev = r["evals"]
    assert len(ev) == 3
    assert ev[1]["loss"] == ev[0]["loss"] and ev[1]["miou"] == ev[0]["miou"]
    assert r["resume_epoch"] == 1 and r["resume_moved"] > 150
    assert r["log2"][0]["epoch"] == 1


def test_reference_main_vg_segmentation_config(tmp_path):
    r, _ = _run(tmp_path, "masks")
    log0 = r["log"][0]
    for k in ("train_loss_mask", "train_loss_dice", "train_loss_bbox", "train_loss_giou", "test_seg_miou"):
        assert k in log0 and log0[k] == log0[k], (k, log0)
    assert r["opt_groups"][3] > 0                           # bbox_attention / mask_head LR group (main_vg.py:31)
    assert r["evals"][1]["loss"] == r["evals"][0]["loss"]
