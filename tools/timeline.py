"""Kernel timeline of one replayed step (the nsys substitute: nsys is not installed; torch.profiler's CUPTI activity records carry
name, start, duration and stream of every kernel of a CUDA-graph replay, including the ctypes-launched ones).

    python tools/timeline.py [--workload cfg2] [--out gpurun_out/timeline] [--world N under torchrun]

Writes <out>_kernels.csv (start_us, dur_us, stream, name) and <out>_summary.txt: the step's span, busy time per stream, the list of
gaps on the union of all streams, time per kernel family, and -- what decides where to optimise -- the CRITICAL CHAIN: walking back
from the last kernel, at each point the kernel that finished last before the current one started."""
import argparse
import json
import os
import sys
from collections import defaultdict

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402


def short(name):
    n = name.replace("void ", "").replace("rb::", "")
    for k in ("umma_gemm_kernel<0, 0>", "umma_gemm_kernel<0, 1>", "umma_gemm_kernel<0, 2>", "umma_gemm_kernel<1, 2>", "umma_gemm_kernel<1, 0>"):
        if k in n:
            return k
    n = n.split("(")[0]
    if "at::" in n or "vectorized_elementwise" in n:
        for k in ("MulFunctor", "FillFunctor", "CUDAFunctor_add", "direct_copy", "sigmoid", "where"):
            if k in name:
                return "at::" + k
        return "at::other"
    return n[:60]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="cfg2")
    ap.add_argument("--out", default="gpurun_out/timeline")
    ap.add_argument("--steps", type=int, default=3)
    a = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        torch.distributed.init_process_group("nccl", device_id=torch.device("cuda", local))
    device = torch.device("cuda", local)
    torch.cuda.set_device(device)
    model, crit, _ = bench.build_ours(device, a.workload)
    model.train()
    net = torch.nn.parallel.DistributedDataParallel(model, device_ids=[local]) if world > 1 else model
    shape = bench.WORKLOADS[a.workload][1]
    s_host, t_host = bench.host_batch(shape["B"], pinned=True, shape=shape)
    s, t = bench.to_device(s_host, t_host, device)

    def step():
        net.zero_grad(set_to_none=True)
        ld = crit(net(s), bench.targets_list(t, masks=a.workload == "cfg3"))
        sum(ld[k] * crit.weight_dict[k] for k in ld if k in crit.weight_dict).backward()

    for _ in range(6):
        step()
    torch.cuda.synchronize()
    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        step()                      # ONE step between two synchronisations: the span below is that step's GPU timeline
        torch.cuda.synchronize()
    if rank != 0:
        return
    trace = a.out + "_trace.json"
    os.makedirs(os.path.dirname(a.out) or ".", exist_ok=True)
    prof.export_chrome_trace(trace)
    ev = json.load(open(trace))["traceEvents"]
    ks = [e for e in ev if e.get("cat") in ("kernel", "gpu_memset", "gpu_memcpy") and "dur" in e]
    ks.sort(key=lambda e: e["ts"])
    os.remove(trace)
    step_ks = ks
    t0 = step_ks[0]["ts"]
    rows = [(e["ts"] - t0, e["dur"], e.get("args", {}).get("stream", e.get("tid")), e["name"]) for e in step_ks]
    with open(a.out + "_kernels.csv", "w") as f:
        f.write("start_us,dur_us,stream,name\n")
        for r in rows:
            f.write(f"{r[0]:.2f},{r[1]:.2f},{r[2]},\"{short(r[3])}\"\n")
    span = max(r[0] + r[1] for r in rows)
    lines = [f"workload {a.workload}  world {world}  kernels {len(rows)}  span {span:.1f} us  sum of durations {sum(r[1] for r in rows):.1f} us"]
    by_stream = defaultdict(list)
    for r in rows:
        by_stream[r[2]].append(r)
    for sid, rs in sorted(by_stream.items(), key=lambda kv: -sum(r[1] for r in kv[1])):
        lines.append(f"  stream {sid}: {len(rs)} kernels, busy {sum(r[1] for r in rs):.1f} us, first {rs[0][0]:.1f}, last end {max(r[0] + r[1] for r in rs):.1f}")
    # union busy / idle
    iv = sorted((r[0], r[0] + r[1]) for r in rows)
    busy, cur_s, cur_e, gaps = 0.0, iv[0][0], iv[0][1], []
    for s_, e_ in iv[1:]:
        if s_ > cur_e:
            busy += cur_e - cur_s
            gaps.append((s_ - cur_e, cur_e))
            cur_s, cur_e = s_, e_
        else:
            cur_e = max(cur_e, e_)
    busy += cur_e - cur_s
    lines.append(f"GPU busy (union of streams) {busy:.1f} us, idle {span - busy:.1f} us in {len(gaps)} gaps; gaps > 3 us: {sum(1 for g in gaps if g[0] > 3)} "
                 f"totalling {sum(g[0] for g in gaps if g[0] > 3):.1f} us")
    fam = defaultdict(lambda: [0.0, 0])
    for r in rows:
        fam[short(r[3])][0] += r[1]
        fam[short(r[3])][1] += 1
    lines.append("time per kernel family (sum of durations, overlapped or not):")
    for k, (d, n) in sorted(fam.items(), key=lambda kv: -kv[1][0])[:40]:
        lines.append(f"  {d:9.1f} us  x{n:4d}  {k}")
    # critical chain: from the last-ending kernel walk back to the kernel (any stream) that ended last before this one started
    chain = []
    cur = max(rows, key=lambda r: r[0] + r[1])
    by_end = sorted(rows, key=lambda r: r[0] + r[1])
    import bisect
    end_keys = [r[0] + r[1] for r in by_end]
    while True:
        chain.append(cur)
        i = bisect.bisect_right(end_keys, cur[0] + 1e-6) - 1
        if i < 0:
            break
        nxt = by_end[i]
        if nxt is cur:
            break
        cur = nxt
    chain.reverse()
    cfam = defaultdict(lambda: [0.0, 0.0, 0])
    prev_end = 0.0
    for r in chain:
        f = cfam[short(r[3])]
        f[0] += r[1]
        f[1] += max(0.0, r[0] - prev_end)
        f[2] += 1
        prev_end = r[0] + r[1]
    lines.append(f"critical chain: {len(chain)} kernels, kernel time {sum(r[1] for r in chain):.1f} us, gaps before them {sum(v[1] for v in cfam.values()):.1f} us")
    for k, (d, g, n) in sorted(cfam.items(), key=lambda kv: -(kv[1][0] + kv[1][1]))[:30]:
        lines.append(f"  {d:9.1f} us kernel + {g:7.1f} us gap  x{n:4d}  {k}")
    # phases along the chain: cumulative time at a few markers
    open(a.out + "_summary.txt", "w").write("\n".join(lines) + "\n")
    print("\n".join(lines))
    if world > 1:
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
