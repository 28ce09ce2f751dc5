"""all-reduce timing probe (run under torchrun): sizes around the flat gradient buffer, fp32 and bf16."""
import os, torch, torch.distributed as dist
local = int(os.environ["LOCAL_RANK"]); torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
w = dist.get_world_size()
for dt in (torch.float32, torch.bfloat16):
    for n in (16 << 20, 64 << 20, 152 << 20):
        x = torch.ones(n, dtype=dt, device="cuda")
        for _ in range(3): dist.all_reduce(x)
        torch.cuda.synchronize(); dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5): dist.all_reduce(x)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        by = n * x.element_size()
        if dist.get_rank() == 0:
            print(f"{dt} {by/1e6:.0f} MB: {ms:.3f} ms  algbw {by/ms/1e6:.0f} GB/s  busbw {by/ms/1e6*2*(w-1)/w:.0f} GB/s", flush=True)
dist.destroy_process_group()
