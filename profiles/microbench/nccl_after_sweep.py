"""Where do the ~85 ms go when a tiny NCCL all-reduce follows the K4 sweep?  (2 ranks)"""
import os, sys, time, torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
local = int(os.environ["LOCAL_RANK"]); torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
from hiertcn_b200.args import make_args
from hiertcn_b200.model_hier import HierTCN
from hiertcn_b200.data_loader import synthetic_batch
N, B = 200_000, 512
model = HierTCN(make_args(["--item_num", str(N)]), None, precision="bf16").build()
x, y, m = synthetic_batch(B, 10, 20, N, seed=1, lengths="dense", id_dist="uniform")
staged = model.stage(x, y, m, None)
buf = torch.zeros(8, device="cuda")
def step(mode):
    scores, st = model.forward(staged=staged)
    r = model.loss(scores, metrics=True)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    ev[0].record()
    t0 = time.perf_counter()
    if mode == "nccl": dist.all_reduce(buf)
    elif mode == "nccl_async": h = dist.all_reduce(buf, async_op=True)
    t1 = time.perf_counter()
    ev[1].record()
    buf.add_(1.0)
    ev[2].record()
    return ev, (t1 - t0) * 1e3
for mode in ("none", "nccl", "nccl_async", "none"):
    for _ in range(3): step(mode)
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    recs = [step(mode) for _ in range(5)]
    torch.cuda.synchronize()
    tot = (time.perf_counter() - t0) * 1e3 / 5
    if dist.get_rank() == 0:
        print("%-10s step %.2f ms | host time of the collective call: %s | gpu gap around it: %s" % (
            mode, tot, " ".join("%.2f" % r[1] for r in recs), " ".join("%.2f" % r[0][0].elapsed_time(r[0][1]) for r in recs)))
dist.destroy_process_group()
