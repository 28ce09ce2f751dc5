"""Host-side cost of the pieces of a data-parallel step (run under torchrun, 2 ranks): how long does the CPU spend enqueueing
(a) a 600 MB all-reduce while the GPU is busy, (b) a 1-element all-reduce, (c) both with async_op=True."""
import os, time, torch, torch.distributed as dist
local = int(os.environ["LOCAL_RANK"]); torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
x = torch.ones(152 << 20, device="cuda"); s = torch.ones(1, device="cuda")
a = torch.randn(8192, 8192, device="cuda"); 
def busy():
    for _ in range(6): torch.mm(a, a)
for _ in range(3):
    dist.all_reduce(x); dist.all_reduce(s)
torch.cuda.synchronize(); dist.barrier()
for name, fn in (("big sync-api", lambda: dist.all_reduce(x)), ("small sync-api", lambda: dist.all_reduce(s)),
                 ("big async_op", lambda: dist.all_reduce(x, async_op=True)), ("clone 600MB", lambda: x.clone())):
    ts = []
    for _ in range(5):
        torch.cuda.synchronize(); busy()
        t0 = time.perf_counter(); r = fn(); ts.append((time.perf_counter() - t0) * 1e3)
        torch.cuda.synchronize()
    if dist.get_rank() == 0: print(f"{name}: host ms {['%.3f' % t for t in ts]}", flush=True)
dist.destroy_process_group()
