"""Latency of a tiny NCCL all-reduce between steps (2+ ranks): python -m torch.distributed.run --nproc-per-node 2 nccl_small.py"""
import os, time, torch, torch.distributed as dist
local = int(os.environ["LOCAL_RANK"]); torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
buf = torch.zeros(8, device="cuda")
big = torch.randn(8192, 8192, device="cuda", dtype=torch.bfloat16)
for name, work in (("allreduce only", lambda: None), ("after 8k matmul", lambda: big @ big)):
    for _ in range(3):
        work(); dist.all_reduce(buf)
    torch.cuda.synchronize(); dist.barrier()
    ts = []
    for _ in range(10):
        work()
        torch.cuda.synchronize()
        t0 = time.perf_counter(); dist.all_reduce(buf); torch.cuda.synchronize(); ts.append((time.perf_counter() - t0) * 1e3)
    if dist.get_rank() == 0:
        print(name, "all_reduce(8 floats) ms:", " ".join("%.3f" % t for t in ts))
dist.destroy_process_group()
