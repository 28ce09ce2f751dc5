"""bench.py -- samples/s of the RefTR forward+backward hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference|stock-gpu]

A "step" is one forward + criterion + backward pass in TRAIN mode (every dropout of the reference active; no optimizer step, as
in SURVEY.md 8(d)) over one synthetic batch of
BASELINE.json configs[1]: ResNet-50 + 6+6-layer RefTR, 640x640 images, 20-token phrases, batch 16 PER GPU (the
reference's --batch_size is per process, main_vg.py:208-209; weak scaling), random-init weights of that architecture.
Under torchrun (N > 1) the model is wrapped in DistributedDataParallel (NCCL gradient all-reduce, as main_vg.py:293-296).

  value : inputs already resident in HBM when the timed region starts
  e2e   : the same step through the public API (build_reftr -> model(samples) -> criterion -> backward) with the batch in
          pinned HOST memory: H2D copy of the inputs (prefetched on a side stream, as engine_vg.py's data_prefetcher does) and
          D2H read of the loss inside the timed region, every step
  roofline : the tcgen05 GEMM / implicit-conv kernel (dominant kernel), timed live with CUDA events launch by launch in an
          extra eager pass after the timed region: algorithmic FLOPs of those launches / their summed duration
  cpu_baseline : the fp32 oracle (oracle/reftr_oracle.py, a restatement of the reference) on the host cores, rank 0, N=1
  --impl reference : the oracle on the host cores with all threads (the reference is pure Python and cannot travel to the
          GPU box; the oracle is pinned against it by tests/golden/), each step a bounded sample of the same workload
  --impl stock-gpu : the oracle on the GPU in stock fp32 PyTorch (cuDNN/cuBLAS) -- the "5x" denominator of the north star
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

GF_FWD_BWD = 207.83  # algorithmic GFLOP per sample, fwd+bwd, config 2 (FlopCounterMode on the reference)
NCU_SUMMARY = os.path.join(ROOT, "profiles", "ncu_summary.json")  # written by tools/ncu_to_json.py from committed ncu captures


def ncu_summary():
    """Profiler-derived figures (DRAM traffic of the GEMM family, tensor-pipe % of the encoder attention kernel) are NOT measured by
    this script (a number taken under a profiler is never a bench value, and ncu cannot run inside the timed process): they are
    read from profiles/ncu_summary.json, which tools/ncu_to_json.py writes from the committed ncu CSV captures.  The line carries
    the file's SHA-256 and the hash of the kernel SOURCES the capture was taken with, next to the hash of the sources the loaded
    library is built from, so a stale capture is visible instead of silently quoted (the binary's own hash is not reproducible from
    build to build)."""
    import hashlib
    try:
        raw = open(NCU_SUMMARY, "rb").read()
        d = json.loads(raw)
        d["file"] = "profiles/ncu_summary.json"
        d["file_sha256"] = hashlib.sha256(raw).hexdigest()[:16]
    except Exception:
        return None
    try:
        from reftr_b200 import _lib
        d["kernel_src_sha256_now"] = _lib.kernel_source_hash()
        d["stale"] = d.get("kernel_src_sha256") != d["kernel_src_sha256_now"]
    except Exception:
        pass
    return d
GF_BY_WORKLOAD = {"cfg2": 207.83, "cfg3": 220.77, "cfg4": 304.78, "cfg5": 616.74}
WORKLOAD = dict(B=16, H=640, W=640, L=20)
METRIC = "samples/sec fwd+bwd (640x640, 20-tok phrase, bs16/GPU)"


CONFIG = {"workload": "cfg2: ResNet-50 + 6+6-layer RefTR box model, 640x640, 20-token phrase, bs16 per GPU, aux_loss",
          "mode": "model.train(): dropout 0.1 active at every site of the reference (MHA probabilities, residual / FFN dropouts, "
                  "mlp_mapping, BERT), fwd+criterion+bwd, no optimizer"}


def peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))), "measured"
    except Exception:
        return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons of one GPU while the timed region runs (NVML, 50 ms period)."""

    BITS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.stop_flag, self.sm, self.reasons, self.max_mhz = index, False, [], set(), None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        while not self.stop_flag and self.nv is not None:
            try:
                self.sm.append(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM))
                r = self.nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in self.BITS.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.05)

    def summary(self):
        self.stop_flag = True
        if not self.sm:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["unavailable"]}
        s = sorted(self.sm)
        return {"sm_mhz": s[len(s) // 2], "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(s)}


WORKLOADS = {  # name -> (flag list key, synthetic input shape)
    "cfg2": ("cfg2_box_r50", dict(B=16, H=640, W=640, L=20)),
    "cfg3": ("cfg3_seg_r50", dict(B=8, H=640, W=640, L=20)),
    "cfg4": ("cfg4_flickr", dict(B=32, H=640, W=640, L=90, n_valid=30, n_ph=5)),
    "cfg5": ("cfg5_r101", dict(B=16, H=800, W=800, L=40)),
}


def build_ours(device, workload="cfg2"):
    os.environ.setdefault("REFTR_B200_RANDOM_BERT", "1")  # no network / HF cache on the box: random-init BERT-base
    from reftr_b200 import build_reftr
    from reftr_b200.args import CONFIG_FLAGS, parse
    from reftr_b200.synthetic import synthetic_weights
    extra = os.environ.get("REFTR_B200_BENCH_FLAGS", "").split()  # diagnostics only (e.g. --freeze_bert): not the metric's configuration
    args = parse(CONFIG_FLAGS[WORKLOADS[workload][0]] + ["--device", str(device)] + extra)
    torch.manual_seed(0)
    model, criterion, _ = build_reftr(args)
    synthetic_weights(model, seed=0)
    model.to(device)
    return model, criterion, args


def build_oracle_model(device):
    from transformers import BertConfig, BertModel
    from oracle.reftr_oracle import RefTROracle
    from reftr_b200.synthetic import synthetic_weights
    torch.manual_seed(1234)
    model = RefTROracle(BertModel(BertConfig()), enc=6, dec=6, dropout=0.1, aux_loss=True)
    synthetic_weights(model, seed=0)
    return model.to(device)


def host_batch(B, pinned, shape=None):
    from reftr_b200.synthetic import ImageList, synthetic_samples, synthetic_targets
    shape = dict(shape or WORKLOAD)
    shape["B"] = B
    s = synthetic_samples(**shape)
    tgt = synthetic_targets(B, max(shape.get("n_ph", 0), 1))
    if pinned:
        s = {k: (ImageList(v.tensors.pin_memory(), v.mask.pin_memory()) if k == "img" else v.pin_memory()) for k, v in s.items()}
        tgt = tgt.pin_memory()
    return s, tgt


def to_device(s, tgt, device):
    d = {k: v.to(device, non_blocking=True) for k, v in s.items()}
    t = tgt.to(device, non_blocking=True)
    return d, t


def targets_list(tgt, masks=False):
    # multi-phrase synthetic samples end with one empty pad phrase (reftr_b200/synthetic.py): it has no target box
    out = [{"boxes": (b if b.shape[0] == 1 else b[:-1]), "labels": [0] * (b.shape[0] if b.shape[0] == 1 else b.shape[0] - 1)} for b in tgt]
    if masks:  # cfg3: one random-rectangle mask per sample at image resolution
        for i, t in enumerate(out):
            m = torch.zeros(1, 640, 640, dtype=torch.bool, device=tgt.device)
            m[:, 100 + 10 * i:400, 150:500 - 10 * i] = True
            t["masks"] = m
    return out


def nbytes(s, tgt):
    n = tgt.numel() * tgt.element_size()
    for k, v in s.items():
        for t in ((v.tensors, v.mask) if k == "img" else (v,)):
            n += t.numel() * t.element_size()
    return n


def oracle_cpu_rate(B_sample, steps, warmup):
    """fp32 oracle fwd + loss + bwd on the host cores; returns (samples/s, threads)."""
    from oracle.reftr_oracle import total_box_loss
    from reftr_b200.synthetic import synthetic_samples, synthetic_targets
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    model = build_oracle_model("cpu").train()  # train mode: dropout active, as in our arm (see config.mode)
    times = []
    for i in range(warmup + steps):
        s = synthetic_samples(B_sample, WORKLOAD["H"], WORKLOAD["W"], WORKLOAD["L"], seed=1 + i)
        tgt = synthetic_targets(B_sample)
        t0 = time.perf_counter()
        model.zero_grad(set_to_none=True)
        total_box_loss(model(s), tgt).backward()
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
    return B_sample * len(times) / sum(times), threads, sum(times) / len(times)


def stock_and_parity(model, s_dev, t_dev, device, value):
    """Two things the north star asks to sit next to the number, measured in the same process on the same GPU with the oracle
    (the fp32 restatement of the reference, pinned to it by tests/golden/) used as the CHECKER / the stock baseline, never as the
    thing measured above:
      stock_gpu_baseline : the reference's stock PyTorch-CUDA path (fp32, cuDNN TF32 convolutions, fp32 matmul: PyTorch defaults),
                           train mode, fwd + criterion + bwd, same batch -- the ">= 5x" denominator
      parity             : eval-mode forward of THIS package against the oracle (TF32 off) on the metric's batch: rel-L2 of pred_boxes
                           for every decoder layer (north star tolerance 1e-3) and the discrete outputs"""
    from oracle.reftr_oracle import total_box_loss
    res = {}
    try:
        oracle = build_oracle_model(device)
        # ---- parity (eval mode, TF32 off) ------------------------------------------------------------------------------
        tf32 = (torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32)
        torch.backends.cuda.matmul.allow_tf32 = False
        torch.backends.cudnn.allow_tf32 = False
        was_training = model.training
        try:
            oracle.eval()
            model.eval()
            with torch.no_grad():
                oo = oracle(s_dev)
                oc = model(s_dev)
            lo = [x["pred_boxes"] for x in oo["aux_outputs"]] + [oo["pred_boxes"]]
            lc = [x["pred_boxes"] for x in oc["aux_outputs"]] + [oc["pred_boxes"]]
            rels = [((c.float() - o).norm() / o.norm()).item() for c, o in zip(lc, lo)]
            res["parity"] = {"pred_boxes_rel_l2_per_decoder_layer": [round(r, 6) for r in rels], "max": max(rels), "tolerance": 1e-3,
                             "within_tolerance": max(rels) < 1e-3, "phrase_mask_bit_exact": bool(torch.equal(oc["phrase_mask"], oo["phrase_mask"])),
                             "max_abs_box_error": max((c.float() - o).abs().max().item() for c, o in zip(lc, lo)),
                             "checker": "fp32 oracle (oracle/reftr_oracle.py, TF32 off) on the same GPU, same weights, the metric's bs16 batch, eval mode; "
                                        "pred_masks / mask_att of cfg3 are held to 5e-3 (tests/test_full_size_gpu.py; SURVEY.md 0.9)"}
        finally:
            torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = tf32
            model.train(was_training)
        # ---- stock PyTorch-CUDA baseline (train mode, PyTorch default math modes) -------------------------------------------------
        oracle.train()

        def stock_step():
            oracle.zero_grad(set_to_none=True)
            total_box_loss(oracle(s_dev), t_dev).backward()
        for _ in range(3):
            stock_step()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n = 10
        e0.record()
        for _ in range(n):
            stock_step()
        e1.record()
        torch.cuda.synchronize()
        B = s_dev["sentence"].shape[0]
        rate = B * n / e0.elapsed_time(e1) * 1e3
        res["stock_gpu_baseline"] = {"value": rate, "unit": "samples/s", "steps": n, "ratio_ours_over_stock": value / rate,
                                     "what": "the oracle (restatement of the reference, pinned by tests/golden) in stock PyTorch on this GPU: fp32, cuDNN TF32 "
                                             "convolutions, fp32 matmul, model.train(), fwd + box loss + bwd, same bs16 batch; north star target >= 5x"}
        del oracle
        torch.cuda.empty_cache()
    except Exception as e:  # the bench line must not die with the diagnostic
        res["stock_gpu_baseline_error"] = repr(e)[:300]
    return res


def verify_data_parallel(model, net, crit, device, world, rank, workload):
    """Hardware check of the multi-GPU path (SURVEY section 4 "distributed" tier): the gradient that rank 0 holds after the split
    backward + sliced NCCL all-reduce of a data-parallel step (every rank on its own shard) against the gradient of ONE process
    running the global batch.  Eval mode (dropout off, so both runs are deterministic); every rank runs both passes (the criterion's
    num_boxes all-reduce is collective).  rel-L2 over the whole flat gradient and the worst per-tensor rel-L2."""
    import contextlib
    from reftr_b200.synthetic import ImageList, synthetic_samples, synthetic_targets
    shape = dict(WORKLOADS[workload][1])
    bv = 4
    shape["B"] = bv * world
    s_all = synthetic_samples(**shape, seed=123, device=device)
    n_ph = max(shape.get("n_ph", 0), 1)
    t_all = synthetic_targets(bv * world, n_ph, seed=321, device=device)

    def shard(lo, hi):
        return {k: (ImageList(v.tensors[lo:hi], v.mask[lo:hi]) if k == "img" else v[lo:hi]) for k, v in s_all.items()}, t_all[lo:hi]

    def run(module, s, t):
        model.zero_grad(set_to_none=True)
        ld = crit(module(s), targets_list(t, masks=False))
        sum(ld[k] * crit.weight_dict[k] for k in ld if k in crit.weight_dict).backward()
        torch.cuda.synchronize()
        return {n: p.grad.detach().clone() for n, p in model.named_parameters() if p.grad is not None}

    was_training = model.training
    model.eval()
    try:
        s_loc, t_loc = shard(rank * bv, (rank + 1) * bv)
        for _ in range(3):                       # eager, capture, replay: the replayed split-graph path is the one compared
            g_dp = run(net, s_loc, t_loc)
        model.engine_allreduce = False
        ctx = net.no_sync() if hasattr(net, "no_sync") else contextlib.nullcontext()
        with ctx:
            # (a) the same arithmetic without any exchange: this process runs EVERY rank's shard itself and averages the gradients
            g_ref = None
            for r in range(world):
                s_r, t_r = shard(r * bv, (r + 1) * bv)
                for _ in range(2):
                    g_r = run(model, s_r, t_r)
                g_ref = g_r if g_ref is None else {n: g_ref[n] + g_r[n] for n in g_ref}
            g_ref = {n: g / world for n, g in g_ref.items()}
            # (b) one process on the global batch (different batch size => 1 / num_boxes scales the 16-bit activation gradients differently)
            for _ in range(2):
                g_one = run(model, s_all, t_all)
        model.engine_allreduce = True
    finally:
        model.train(was_training)

    def cmp(ga, gb):
        num = sum(((ga[n].double() - gb[n].double()) ** 2).sum() for n in gb).sqrt().item()
        den = sum((gb[n].double() ** 2).sum() for n in gb).sqrt().item()
        big = max(g.norm().item() for g in gb.values())
        per = {n: ((ga[n] - gb[n]).norm() / (gb[n].norm() + 1e-20)).item() for n in gb if gb[n].norm().item() > 1e-4 * big}
        worst = max(per.items(), key=lambda kv: kv[1])
        return num / max(den, 1e-30), worst, len(per)
    rel, worst, n = cmp(g_dp, g_ref)
    rel_g, worst_g, _ = cmp(g_dp, g_one)
    return {"world": world, "samples_per_rank": bv, "rel_l2_flat_gradient": rel, "worst_tensor": worst[0], "worst_tensor_rel_l2": worst[1],
            "tensors_compared": n, "tolerance": 1e-4, "ok": rel < 1e-4,
            "vs_one_process_on_the_global_batch": {"rel_l2_flat_gradient": rel_g, "worst_tensor": worst_g[0], "worst_tensor_rel_l2": worst_g[1],
                                                   "note": "not the same arithmetic: the global batch halves 1/num_boxes, i.e. the magnitude of every 16-bit activation gradient"},
            "what": "rank 0's gradient after the split backward + sliced NCCL all-reduces (each rank on its shard) vs the average of the per-shard "
                    "gradients computed by ONE process without any exchange (same kernels, same batch shapes; only fp32 atomic order differs); eval mode"}


def optimizer_diag(model, step_fn, device):
    """SURVEY 8(f) row N3 (NOT part of the metric's timed region): the reference's per-iteration clip_grad_norm_(0.1) + AdamW.step()
    (engine_vg.py:62-67, main_vg.py:234-268) as torch runs it vs reftr_b200.optim on flat buffers, on the gradients of one step; and
    a whole training iteration (fwd + bwd + clip + step, which also re-packs every folded 16-bit weight copy)."""
    from reftr_b200 import optim as ro

    def groups(named):
        bb = [p for n, p in named if "img_backbone.0" in n]
        bert = [p for n, p in named if "lang_backbone" in n]
        rest = [p for n, p in named if "img_backbone.0" not in n and "lang_backbone" not in n]
        return [{"params": rest, "lr": 1e-4}, {"params": bb, "lr": 1e-5}, {"params": bert, "lr": 1e-5}]

    def timed_ms(fn, n):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n

    named = [(n, p) for n, p in model.named_parameters() if p.requires_grad]
    step_fn()
    twin = [(n, torch.nn.Parameter(p.detach().clone())) for n, p in named]
    for (_, t), (_, p) in zip(twin, named):
        t.grad = p.grad.detach().clone()
    opt_t = torch.optim.AdamW(groups(twin), lr=1e-4, weight_decay=1e-4)
    tp = [t for _, t in twin]

    def torch_step():
        torch.nn.utils.clip_grad_norm_(tp, 0.1)
        opt_t.step()
    torch_step()
    ms_torch = timed_ms(torch_step, 5)
    del opt_t, twin, tp
    opt = ro.FusedAdamW.for_model(model, groups(named), lr=1e-6, weight_decay=1e-4)
    params = [p for _, p in named]

    def fused_step():
        ro.clip_grad_norm_(params, 0.1)
        opt.step()
    step_fn()
    fused_step()
    zero_copy = opt.flat_g is None
    ms_fused = timed_ms(fused_step, 5)

    def train_iter():
        step_fn()
        fused_step()
    for _ in range(3):
        train_iter()
    ms_iter = timed_ms(train_iter, 5)
    return {"torch_clip_adamw_ms": ms_torch, "fused_clip_adamw_ms": ms_fused, "params": sum(p.numel() for p in params),
            "zero_copy_gradients": zero_copy, "train_iteration_ms": ms_iter,
            "note": "train_iteration = fwd + criterion + bwd + clip + AdamW, incl. the re-pack of all folded 16-bit weights after each update"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference", "stock-gpu"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--workload", default="cfg2", choices=["cfg2", "cfg3", "cfg4", "cfg5"],
                    help="cfg2 (default, the metric's configuration); cfg3 = +mask head bs8; cfg4 = Flickr multi-phrase L=90 n_ph=5 bs32; "
                         "cfg5 = ResNet-101 800x800 L=40 (bs16 here)")
    ap.add_argument("--windows", type=int, default=5, help="number of timed windows of --steps steps each; value = the median window")
    ap.add_argument("--batch", type=int, default=0, help="override the per-GPU batch of --workload (e.g. cfg5 at its stated bs64)")
    ap.add_argument("--verify", action="store_true", help="N > 1: check the all-reduced gradient against a 1-GPU run of the global batch")
    ap.add_argument("--diag", default="", help="diagnostics only: 'noddp' = no DDP wrapper (engine all-reduce only); 'replicas' = no exchange at all")
    a = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    pk, pk_src = peaks()
    base = {"metric": METRIC, "unit": "samples/s", "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "data": "synthetic"}

    # ------------------------------------------------------------------------------------------------ reference arm
    if a.impl == "reference":
        if rank != 0:
            return
        bs = 2
        rate, threads, sec = oracle_cpu_rate(bs, a.steps, a.warmup)
        sample = f"fp32 oracle fwd+loss+bwd on host CPU, {bs} samples of the cfg2 workload per step (640x640, L=20), {threads} threads"
        out = dict(base, impl="reference", value=rate, ms_per_step=sec * 1e3, dtype="f32",
                   config=dict(CONFIG, global_batch=bs, parallelism="cpu", sample=sample),
                   cpu_baseline={"value": rate, "unit": "samples/s", "cores": threads, "kind": "port", "sample": sample},
                   e2e={"value": rate, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, gpu_launches=0)
        print(json.dumps(out))
        return

    if world > 1:
        torch.distributed.init_process_group("nccl", device_id=torch.device("cuda", local))
    device = torch.device("cuda", local)
    torch.cuda.set_device(device)
    B = WORKLOAD["B"]

    if a.impl == "stock-gpu":
        from oracle.reftr_oracle import total_box_loss
        model = build_oracle_model(device).train()
        crit = None
    else:
        model, crit, _ = build_ours(device, a.workload)
        for pref in os.environ.get("REFTR_B200_BENCH_FREEZE", "").split():  # diagnostics only: what a branch's backward costs
            for n_, p_ in model.named_parameters():
                if n_.startswith(pref):
                    p_.requires_grad_(False)
        model.train() if os.environ.get("REFTR_B200_BENCH_EVAL") != "1" else model.eval()  # train mode: every dropout of the reference active
    net = model
    if a.diag == "replicas":
        model.engine_allreduce = False
    if world > 1 and not a.diag:
        net = torch.nn.parallel.DistributedDataParallel(model, device_ids=[local])

    wl_shape = dict(WORKLOADS[a.workload][1])
    if a.batch:
        wl_shape["B"] = a.batch
    B = wl_shape["B"]
    s_host, t_host = host_batch(B, pinned=True, shape=wl_shape)
    s_dev, t_dev = to_device(s_host, t_host, device)
    torch.cuda.synchronize()

    def loss_of(out, tgt):
        if crit is None:
            return total_box_loss(out, tgt)
        ld = crit(out, targets_list(tgt, masks=a.workload == "cfg3"))
        return sum(ld[k] * crit.weight_dict[k] for k in ld if k in crit.weight_dict)

    def step_resident():
        net.zero_grad(set_to_none=True)
        loss = loss_of(net(s_dev), t_dev)
        loss.backward()
        return loss

    # e2e: the batch lives in pinned host memory; every step copies ITS inputs host->device and reads its loss back.  The copy of
    # step i+1 is issued on a side stream while step i computes, exactly what the reference's data_prefetcher does
    # (engine_vg.py:234-291: side CUDA stream + record_stream).
    copy_stream = torch.cuda.Stream(device=device)
    pending = {}
    # the images travel as RAW uint8 HWC (what the data-loader workers hold before to_tensor + Normalize,
    # datasets/transforms.py:233-250) -- a quarter of the bytes of the normalised fp32 batch -- and reftr_b200.data.DeviceCollator
    # normalises / pads / builds the mask on the device (bit-exact with the reference's CPU path, tests/test_data_gpu.py).
    # REFTR_B200_E2E_FP32=1 ships the reference's normalised fp32 batch instead (round-1 behaviour).
    use_u8 = a.impl == "ours" and os.environ.get("REFTR_B200_E2E_FP32") != "1"
    if use_u8:
        from reftr_b200.data import DeviceCollator
        collator = DeviceCollator(device)
        g8 = torch.Generator().manual_seed(7)
        imgs_u8 = [torch.randint(0, 256, (wl_shape["H"], wl_shape["W"], 3), dtype=torch.uint8, generator=g8) for _ in range(B)]
        packed = DeviceCollator.pack(imgs_u8)   # the data-loader worker's part (pinned host memory), outside the step like the fp32 collate
        s_rest = {k: v for k, v in s_host.items() if k != "img"}
        h2d_bytes = packed.nbytes() + sum(v.numel() * v.element_size() for v in s_rest.values()) + t_host.numel() * t_host.element_size()
    else:
        h2d_bytes = nbytes(s_host, t_host)

    def prefetch():
        with torch.cuda.stream(copy_stream):
            if use_u8:
                s = {k: v.to(device, non_blocking=True) for k, v in s_rest.items()}
                s["img"] = collator.upload(packed, stream=copy_stream)
                t = t_host.to(device, non_blocking=True)
            else:
                s, t = to_device(s_host, t_host, device)
            ev = torch.cuda.Event()
            ev.record(copy_stream)
        pending["next"] = (s, t, ev)

    def step_e2e():
        if "next" not in pending:
            prefetch()
        s, t, ev = pending.pop("next")
        torch.cuda.current_stream().wait_event(ev)
        for v in s.values():
            for x in ((v.tensors, v.mask) if hasattr(v, "tensors") else (v,)):
                x.record_stream(torch.cuda.current_stream())
        t.record_stream(torch.cuda.current_stream())
        loss = loss_of(net(s), t)
        prefetch()  # next step's inputs: the copy overlaps this step's compute
        loss.backward()
        net.zero_grad(set_to_none=True)  # (where optimizer.zero_grad() sits in a training loop: after the update, before logging)
        # D2H read of the step's result, every step.  REFTR_B200_E2E_LAG=0: blocking read right here (engine_vg.py:53 style).  Default:
        # the loss goes to pinned host memory asynchronously and is READ one step later, after the next step's forward has been
        # enqueued (a logging lag of one iteration; the last one is read before the timed region closes) -- the GPU is not left idle
        # while the host prepares the next step.
        if lag_read:
            prev = pending.pop("loss", None)
            buf = loss_bufs[step_no[0] & 1]
            buf.copy_(loss.detach(), non_blocking=True)
            ev = torch.cuda.Event()
            ev.record()
            pending["loss"] = (buf, ev)
            step_no[0] += 1
            if prev is not None:
                prev[1].synchronize()
                return float(prev[0])
            return None
        return loss.item()

    lag_read = os.environ.get("REFTR_B200_E2E_LAG", "1") != "0"
    loss_bufs = [torch.zeros((), dtype=torch.float32).pin_memory() for _ in range(2)]
    step_no = [0]

    def e2e_flush():
        prev = pending.pop("loss", None)
        if prev is not None:
            prev[1].synchronize()
            return float(prev[0])

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=device)
        if world > 1:
            torch.distributed.all_reduce(ms, op=torch.distributed.ReduceOp.MAX)
        return ms.item()

    for _ in range(max(a.warmup, 3)):
        step_resident()
    eng = model.engine() if a.impl == "ours" else None
    sampler = ClockSampler(local)
    sampler.start()
    l0 = eng.launches if eng else 0
    # EXACTLY --steps steps per window, barrier + synchronize on both sides, max over ranks; several windows, the MEDIAN one is reported
    # (a single 0.25 s window is noise-dominated at N = 2..8: one late NCCL call moves it by several percent)
    win = sorted(timed(step_resident, a.steps) for _ in range(max(1, a.windows)))
    ms = win[len(win) // 2]
    launches = ((eng.launches - l0) // max(1, a.windows)) if eng else 0
    clocks = sampler.summary()
    for _ in range(2):
        step_e2e()
    def timed_e2e():
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(a.steps):
            step_e2e()
        e2e_flush()   # the last step's loss is read inside the timed region too
        e1.record()
        barrier()
        t_ms = torch.tensor([e0.elapsed_time(e1)], device=device)
        if world > 1:
            torch.distributed.all_reduce(t_ms, op=torch.distributed.ReduceOp.MAX)
        return t_ms.item()
    e2e_flush()
    win_e2e = sorted(timed_e2e() for _ in range(max(1, a.windows)))
    ms_e2e = win_e2e[len(win_e2e) // 2]

    # host-side cost of one step (diagnostic): time to enqueue a whole step without any sync, and time until the forward is enqueued
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    net.zero_grad(set_to_none=True)
    o_ = net(s_dev)
    t1 = time.perf_counter()
    loss_of(o_, t_dev).backward()
    t2 = time.perf_counter()
    torch.cuda.synchronize()
    host_ms = {"enqueue_fwd_ms": (t1 - t0) * 1e3, "enqueue_step_ms": (t2 - t0) * 1e3}
    if eng is not None:
        host_ms.update({k: round(v, 3) for k, v in getattr(eng, "host_ms", {}).items()})

    value = world * B * a.steps / ms * 1e3
    e2e = world * B * a.steps / ms_e2e * 1e3
    out = dict(base, value=value, ms_per_step=ms / a.steps, dtype="f16",
               config={"workload": CONFIG["workload"] if a.workload == "cfg2" else f"{a.workload} (NOT the metric's configuration): {wl_shape}", "global_batch": world * B, "parallelism": f"dp{world}", "mode": CONFIG["mode"] if model.training else "eval+grad (dropout inactive; diagnostic)",
                       "bert": "BERT-base on the same C-ABI kernels (IEEE-half operands, fp32 accumulate / residual / LN), inside the step graphs", "operands": "IEEE half (fp16) tensor-core operands and saved activations, fp32 accumulation, residual stream, normalisation and softmax; static 2^10 loss scale in the backward",
                       "l2": "per-step working set (saved activations, several GB) far exceeds the 126 MB L2; no explicit flush",
                       "timing": f"{max(1, a.windows)} windows of exactly {a.steps} steps (barrier + synchronize on both sides, CUDA events, max over ranks); value = median window"},
               e2e={"value": e2e, "unit": "samples/s", "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": 4,
                    "ms_per_step": ms_e2e / a.steps, "windows_ms_per_step": [round(w / a.steps, 4) for w in win_e2e],
                    "input": ("raw uint8 HWC images in pinned host memory -> one H2D -> rb_collate_u8 (normalise + pad + mask on the device), "
                              "token ids / masks / targets as int64 / fp32") if use_u8 else "normalised fp32 batch in pinned host memory",
                    "loss_read": "every step's loss is copied to pinned host memory and read one step later (REFTR_B200_E2E_LAG=0: blocking read every step)" if lag_read else "blocking .item() every step"},
               windows_ms_per_step=[round(w / a.steps, 4) for w in win],
               gpu_launches=launches, clocks=clocks, host=host_ms)
    if os.environ.get("REFTR_B200_BENCH_FLAGS") or os.environ.get("REFTR_B200_BENCH_FREEZE"):
        out["config"]["diagnostic_flags"] = (os.environ.get("REFTR_B200_BENCH_FLAGS", "") + " frozen: " + os.environ.get("REFTR_B200_BENCH_FREEZE", "")
                                             + " (NOT the metric's configuration)")
    if a.impl == "stock-gpu":
        out["impl"] = "stock-gpu"
        out["dtype"] = "f32 (cuDNN TF32 conv, fp32 matmul: PyTorch defaults)"
    # whole-step roofline: algorithmic FLOPs of the reference graph (BERT included) over the sustained tensor peak
    gf = GF_BY_WORKLOAD[a.workload]
    out["step_roofline"] = {"bound": "tensor", "achieved": value * gf / 1e3 / world, "peak": pk["bf16_tflops_sustained"],
                            "unit": "TFLOP/s", "frac": value * gf / 1e3 / world / pk["bf16_tflops_sustained"],
                            "peak_source": pk_src + " (sustained)", "gflop_per_sample": gf}

    if a.impl == "ours":
        # ---- dominant kernel (tcgen05 GEMM / implicit conv) ---------------------------------------------------------------
        # One eager step records every rb_gemm launch descriptor; the launches are then re-issued group by group (same shapes
        # and epilogue flags) between two CUDA events on the launching stream, three passes per group, each pass one CUDA graph of
        # the group's launches, so the per-launch duration is a device time free of host launch gaps.  achieved = sum of algorithmic FLOPs / sum of those durations.
        from reftr_b200 import ops
        eng.force_eager = True
        ops.PROFILE = []
        step_resident()
        torch.cuda.synchronize()
        rec, ops.PROFILE = ops.PROFILE, None
        eng.force_eager = False
        groups = {}
        for args, fl, sig in rec:
            groups.setdefault(sig, []).append((args, fl))
        g_ms = g_fl = 0.0
        top = []
        for sig, members in groups.items():
            for args, _ in members:  # warm-up pass
                ops.relaunch_gemm(args)
            # the group's launches are replayed as ONE CUDA graph: a python / ctypes call costs ~7 us, more than the small launches
            # themselves, and would be billed to the kernels (eager re-issue is the fallback if the capture fails)
            graph = None
            try:
                torch.cuda.synchronize()
                gph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(gph):
                    for args, _ in members:
                        ops.relaunch_gemm(args)
                gph.replay()
                torch.cuda.synchronize()
                graph = gph
            except Exception:
                graph = None
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(3):
                if graph is not None:
                    graph.replay()
                else:
                    for args, _ in members:
                        ops.relaunch_gemm(args)
            e1.record()
            torch.cuda.synchronize()
            ms_g = e0.elapsed_time(e1) / 3
            fl_g = sum(f for _, f in members)
            g_ms += ms_g
            g_fl += fl_g
            top.append((ms_g, len(members), sig, fl_g))
        top.sort(reverse=True)
        if os.environ.get("REFTR_B200_BENCH_GROUPS"):  # diagnostics: every launch group, not only the six largest
            with open(os.environ["REFTR_B200_BENCH_GROUPS"], "w") as fh:
                for m_, n_, sg, f_ in top:
                    fh.write(f"{m_ * 1e3:9.1f} us  x{n_:3d}  {m_ * 1e3 / n_:7.1f} us each  {f_ / (m_ * 1e-3) / 1e12:7.1f} TFLOP/s  {sg}\n")
        ach = g_fl / (g_ms * 1e-3) / 1e12 if g_ms > 0 else 0.0
        out["roofline"] = {"bound": "tensor", "achieved": ach, "peak": pk["bf16_tflops_sustained"], "unit": "TFLOP/s",
                           "frac": ach / pk["bf16_tflops_sustained"], "traffic": None,
                           "kernel": "rb_gemm: umma_gemm_kernel (tcgen05; every conv fwd+dgrad+wgrad and linear layer of one step) + gemm_skinny_kernel "
                                     "(mma.sync, the M <= 128 decoder / head layers, 0.3 % of the FLOPs)",
                           "launches": len(rec), "avg_launch_us": g_ms * 1e3 / max(len(rec), 1), "gemm_ms_per_step": g_ms,
                           "gemm_gflop_per_step": g_fl / 1e9, "peak_source": pk_src + " (sustained)",
                           "note": "the path is mixed: conv1/layer1/layer2 GEMMs are HBM-bound (SURVEY 8(d)); see DESIGN.md section 5",
                           "top_groups": [{"ms": round(m_, 4), "launches": n_, "mode_M_N_K_taps": list(sg[:5]),
                                           "tflops": round(f_ / (m_ * 1e-3) / 1e12, 1)} for m_, n_, sg, f_ in top[:6]]}
        ns = ncu_summary()
        if ns is not None:
            gm = ns.get("gemm_family") or {}
            if gm.get("dram_bytes_per_launch"):
                out["roofline"]["traffic"] = gm["dram_bytes_per_launch"]
                out["roofline"]["traffic_note"] = gm.get("note", "") + f" [{ns['file']} sha256 {ns['file_sha256']}, stale={ns.get('stale')}]"
            if ns.get("encoder_mha"):
                out["encoder_mha"] = dict(ns["encoder_mha"], source=f"{ns['file']} sha256 {ns['file_sha256']}", kernel_src_sha256_capture=ns.get("kernel_src_sha256"),
                                          kernel_src_sha256_now=ns.get("kernel_src_sha256_now"), stale=ns.get("stale"))
        if world == 1 and a.workload == "cfg2" and os.environ.get("REFTR_B200_BENCH_STOCK", "1") == "1":
            out.update(stock_and_parity(model, s_dev, t_dev, device, value))
        if world == 1 and a.workload == "cfg2" and os.environ.get("REFTR_B200_BENCH_OPTIM", "1") == "1":
            out["next_rows"] = {"N3_optimizer_step": optimizer_diag(model, step_resident, device)}
        if rank == 0 and a.gpus == 1 and not a.no_cpu_baseline:
            bs = 8
            rate, threads, sec = oracle_cpu_rate(bs, 1, 1)
            out["cpu_baseline"] = {"value": rate, "unit": "samples/s", "cores": threads, "kind": "port",
                                   "sample": f"fp32 oracle fwd+loss+bwd, one {bs}-sample batch of the same workload after one warm-up batch, {threads} threads, {sec:.1f} s"}
    if a.verify and world > 1 and a.impl == "ours":
        out["verify"] = verify_data_parallel(model, net, crit, device, world, rank, a.workload)
    if rank == 0:
        print(json.dumps(out))
    if world > 1:
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
