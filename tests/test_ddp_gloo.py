"""Data-parallel host logic on CPU (gloo, world_size 2): the module wrapped in DistributedDataParallel, exactly as
main_vg.py:293-296 wraps the reference, must produce the full-batch gradients when every rank runs half of the batch.
The CUDA kernels are replaced by their torch emulation (tests/emu_ops.py, EXACT fp32 mode) so that the check isolates
the wiring: one autograd node handing every parameter gradient to DDP's hooks, bucketed all-reduce, averaging."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))


def _patch_emulated():
    sys.path.insert(0, HERE)
    sys.path.insert(0, os.path.dirname(HERE))
    import emu_ops
    import reftr_b200.bert as bert
    import reftr_b200.engine as engine
    import reftr_b200.pack as pack
    import reftr_b200.seg as seg

    class Fp32Proxy:
        def __getattr__(self, name):
            return torch.float32 if name == "bfloat16" else getattr(torch, name)

    for m in (engine, pack, bert, seg):
        m.ops = emu_ops
        m.torch = Fp32Proxy()
    emu_ops.EXACT[0] = True


def _slice_samples(s, lo, hi):
    from reftr_b200.synthetic import ImageList
    out = {}
    for k, v in s.items():
        out[k] = ImageList(v.tensors[lo:hi], v.mask[lo:hi]) if k == "img" else v[lo:hi]
    return out


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.set_num_threads(max(1, (os.cpu_count() or 2) // world))
    _patch_emulated()
    from oracle.cases import CASES
    from oracle.reftr_oracle import total_box_loss
    from reftr_b200.synthetic import synthetic_samples, synthetic_targets
    from util_build import build_candidate
    dist.init_process_group("gloo", rank=rank, world_size=world)
    case = CASES["cfg1_box"]
    model = build_candidate(case)
    ddp = torch.nn.parallel.DistributedDataParallel(model)
    B = case["inputs"]["B"]
    per = B // world
    s = _slice_samples(synthetic_samples(**case["inputs"]), rank * per, (rank + 1) * per)
    tgt = synthetic_targets(B)[rank * per:(rank + 1) * per]
    total_box_loss(ddp(s), tgt).backward()
    names = ["bbox_embed.layers.0.weight", "vl_transformer.encoder.layers.0.linear1.weight", "img_backbone.0.body.layer3.1.conv2.weight",
             "input_proj.0.0.weight", "lang_backbone.encoder.layer.0.attention.self.query.weight", "vl_transformer.level_embed"]
    params = dict(model.named_parameters())
    if rank == 0:
        q.put({n: params[n].grad.detach().numpy().copy() for n in names})
    dist.barrier()
    dist.destroy_process_group()


def _worker_seeds(rank, world, port, q):
    """main_vg.py:173-177 seeds every rank with seed + rank BEFORE build_reftr: the replicas are born different.  DDP's constructor
    broadcast skips everything in ``_ddp_params_and_buffers_to_ignore`` (= everything the engine owns), so the engine must bring the
    replicas together itself (RefTREngine._sync_initial_state)."""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.set_num_threads(max(1, (os.cpu_count() or 2) // world))
    _patch_emulated()
    from oracle.cases import CASES
    from oracle.reftr_oracle import total_box_loss
    from reftr_b200.synthetic import synthetic_samples, synthetic_targets, synthetic_weights
    from util_build import build_candidate
    dist.init_process_group("gloo", rank=rank, world_size=world)
    case = CASES["cfg1_box"]
    model = build_candidate(case)
    synthetic_weights(model, seed=100 + rank)          # different weights (and FrozenBN statistics) on every rank
    names = ["bbox_embed.layers.0.weight", "vl_transformer.decoder.layers.0.linear1.weight", "query_encoder.linear2.weight",
             "input_proj.0.0.weight", "map_sentence.0.weight", "vl_transformer.lang_pos_embeddings.weight", "vl_transformer.level_embed",
             "img_backbone.0.body.layer1.0.conv1.weight", "img_backbone.0.body.layer3.0.bn2.running_var",
             "lang_backbone.encoder.layer.1.output.dense.weight"]
    sd = model.state_dict()
    before = {n: sd[n].detach().clone() for n in names}
    ddp = torch.nn.parallel.DistributedDataParallel(model)  # (broadcasts the one parameter it keeps, level_embed, right here)
    B = case["inputs"]["B"]
    per = B // world
    s = _slice_samples(synthetic_samples(**case["inputs"]), rank * per, (rank + 1) * per)
    tgt = synthetic_targets(B)[rank * per:(rank + 1) * per]
    opt = torch.optim.SGD([p for p in model.parameters() if p.requires_grad], lr=1e-2)
    total_box_loss(ddp(s), tgt).backward()
    sd = model.state_dict()
    after_fwd = {n: sd[n].detach().clone() for n in names}
    opt.step()
    sd = model.state_dict()
    after_step = {n: sd[n].detach().clone() for n in names}
    q.put((rank, {n: (before[n].numpy(), after_fwd[n].numpy(), after_step[n].numpy()) for n in names}))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(900)
def test_ddp_replicas_with_different_seeds_are_synchronised():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker_seeds, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=800) for _ in range(2))
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    import numpy as np
    for n in got[0]:
        b0, f0, s0 = got[0][n]
        b1, f1, s1 = got[1][n]
        assert not np.array_equal(b0, b1), n            # the replicas really started different
        assert np.array_equal(f0, b0), n                # rank 0 is the source
        assert np.array_equal(f1, f0), n                # after the first forward every rank holds rank 0's state
        assert np.array_equal(s1, s0), n                # and they stay together after an optimizer step (averaged gradients)
    assert not np.array_equal(got[0]["bbox_embed.layers.0.weight"][2], got[0]["bbox_embed.layers.0.weight"][1])  # the step did move


@pytest.mark.timeout(900)
def test_ddp_world2_matches_full_batch():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = q.get(timeout=800)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    # single process, full batch
    _patch_emulated()
    from oracle.cases import CASES
    from oracle.reftr_oracle import total_box_loss
    from reftr_b200.synthetic import synthetic_samples, synthetic_targets
    from util_build import build_candidate
    import emu_ops
    try:
        case = CASES["cfg1_box"]
        model = build_candidate(case)
        total_box_loss(model(synthetic_samples(**case["inputs"])), synthetic_targets(case["inputs"]["B"])).backward()
        params = dict(model.named_parameters())
        for n, g in got.items():
            ref = params[n].grad
            g = torch.from_numpy(g)
            err = ((g - ref).norm() / (ref.norm() + 1e-12)).item()
            assert err < 1e-3, (n, err)
    finally:
        emu_ops.EXACT[0] = False
        import importlib
        import reftr_b200.bert as bert
        import reftr_b200.engine as engine
        import reftr_b200.pack as pack
        import reftr_b200.seg as seg
        from reftr_b200 import ops as real_ops
        for m in (engine, pack, bert, seg):
            m.ops = real_ops
            m.torch = torch
