"""python profiles/sharded_loss_order.py -- 1-GPU emulation of the bench's sharded parity check: CE partials of a 1M-item catalog swept as W shards x 3 splits
against the whole catalog x 4 splits -- how far do the fp32 loss rows move with the summation order?"""
import sys, torch
sys.path.insert(0, ".")
from hiertcn_b200 import _cabi as cabi
from hiertcn_b200.dist import CatalogTable, CudaScoreOps, shard_bounds
dev = torch.device("cuda", 0)
cabi.load()
N = 1_000_000
g = torch.Generator(device=dev).manual_seed(4242)
wt = torch.empty((N, cabi.WT_PITCH_BF16), dtype=torch.bfloat16, device=dev)
w = torch.randn((128, N), device=dev, generator=g) * 0.3
b = torch.randn(N, device=dev, generator=g) * 0.2
cabi.call("htcn_prepare_wout", w.data_ptr(), b.data_ptr(), N, wt.data_ptr(), cabi.HTCN_BF16, torch.cuda.current_stream(dev).cuda_stream)
full = CatalogTable(wt, None, "bf16", N)
for W in (2, 4, 8):
    worst = 0.0
    for rank in range(W):
        gq = torch.Generator(device=dev).manual_seed(900 + rank)
        hp = torch.randn((512, 128), device=dev, generator=gq).to(torch.bfloat16)
        yp = torch.randint(1, N, (512,), device=dev, generator=gq, dtype=torch.int32)
        ops = CudaScoreOps(full)
        zy = torch.zeros(512, dtype=torch.float32, device=dev)
        ops.target_logit(hp, yp, 0, N, zy)
        part = ops.sweep(hp, yp, zy, 0, N, 0, 4, True, True)
        ref = ops.finish(part["pm"], part["ps"], part["pc"], yp, zy)
        bnd = shard_bounds(N, W)
        pm, ps, pc = [], [], []
        for r in range(W):
            o = CudaScoreOps(full.rows(bnd[r], bnd[r + 1]))
            p = o.sweep(hp, yp, zy, bnd[r], bnd[r + 1], 0, 3, True, True)
            pm.append(p["pm"]); ps.append(p["ps"]); pc.append(p["pc"])
        got = ops.finish(torch.cat(pm), torch.cat(ps), torch.cat(pc), yp, zy)
        d = (got["loss_row"] - ref["loss_row"]).abs()
        worst = max(worst, float(d.max()))
        assert torch.equal(got["rank_row"], ref["rank_row"])
        # fp64 reference of the loss from the dumped logits of 8 rows
    print("W=%d max |loss_sharded - loss_whole| = %.3e (loss ~ %.2f)" % (W, worst, float(ref["loss_row"].mean())))
