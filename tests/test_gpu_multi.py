"""Multi-GPU (NCCL) test of catalog-sharded scoring: world = 2 ranks, each owning half of the catalog
(hiertcn_b200.dist.ShardedCatalogScorer + CudaScoreOps) must reproduce single-GPU scoring of the whole catalog --
loss rows, ranks and top-k lists.  Skipped on boxes with fewer than 2 GPUs (run: gpurun --gpus 2 -- pytest ...)."""
import os
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, precision, ret):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        from hiertcn_b200.args import make_args
        from hiertcn_b200.dist import CudaScoreOps, ShardedCatalogScorer, make_sharded_model
        from hiertcn_b200.model_hier import HierTCN
        from hiertcn_b200.weights import hier_weight_shapes, init_weights
        N, Ql, k = 50_000, 300, 100
        w = init_weights(hier_weight_shapes(N), seed=3, kernel_scale=2.0, bias_noise=0.2)
        a = make_args(["--item_num", str(N)])
        rng = np.random.default_rng(0)
        h_all = (rng.normal(size=(world * Ql, 128))).astype(np.float32)
        y_all = rng.integers(1, N, size=world * Ql).astype(np.int32)
        dt = torch.bfloat16 if precision == "bf16" else torch.float32
        h = torch.from_numpy(h_all[rank * Ql:(rank + 1) * Ql]).cuda().to(dt)
        y = torch.from_numpy(y_all[rank * Ql:(rank + 1) * Ql]).cuda()
        # sharded: this rank holds rows [n0, n1) of W_out^T
        m_sh, n0, n1 = make_sharded_model(a, w, rank, world, precision)
        sc = ShardedCatalogScorer(CudaScoreOps(m_sh), dist, rank, world, N, n_split=3)
        out = sc.score(h, y, k=k)
        # single GPU, whole catalog
        m_full = HierTCN(a, w, precision=precision).build()
        ops = CudaScoreOps(m_full)
        zy = torch.zeros(Ql, dtype=torch.float32, device="cuda")
        ops.target_logit(h, y, 0, N, zy)
        part = ops.sweep(h, y, zy, 0, N, k, 2, True, True)
        ref = ops.finish(part["pm"], part["ps"], part["pc"], y, zy)
        ref.update(ops.topk_merge(part["tv"], part["ti"], k))
        torch.cuda.synchronize()
        ok = (torch.equal(out["rank_row"], ref["rank_row"]) and torch.equal(out["topk_idx"], ref["topk_idx"])
              and torch.equal(out["topk_val"], ref["topk_val"])
              and torch.allclose(out["loss_row"], ref["loss_row"], rtol=2e-6, atol=2e-6))
        ret[rank] = bool(ok)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("precision", ["bf16", "f32"])
def test_sharded_catalog_scoring_nccl(precision):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    ret = ctx.Manager().dict()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, precision, ret)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(300)
        assert p.exitcode == 0
    assert all(ret.get(r) for r in range(2)), dict(ret)


def _peer_worker(rank, world, port, precision, ret):
    """the peer-memory exchange kernels (hiertcn_b200.peer) against the NCCL collectives: same inputs, same shard kernels,
    every output bit for bit, over several calls (the epoch flags) and with k = 0"""
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        from hiertcn_b200.args import make_args
        from hiertcn_b200.dist import CudaScoreOps, ShardedCatalogScorer, make_sharded_model
        from hiertcn_b200.peer import PeerShardedCatalogScorer
        from hiertcn_b200.weights import hier_weight_shapes, init_weights
        N, Ql, k = 50_000, 300, 100
        w = init_weights(hier_weight_shapes(N), seed=3, kernel_scale=2.0, bias_noise=0.2)
        a = make_args(["--item_num", str(N)])
        dt = torch.bfloat16 if precision == "bf16" else torch.float32
        m_sh, n0, n1 = make_sharded_model(a, w, rank, world, precision)
        ops = CudaScoreOps(m_sh)
        nccl = ShardedCatalogScorer(ops, dist, rank, world, N, n_split=3)
        peer = PeerShardedCatalogScorer(ops, dist, rank, world, N, n_split=3)
        ok = True
        for it, kk in enumerate((k, k, 0, k)):
            g = torch.Generator(device="cuda").manual_seed(100 * it + rank)
            h = torch.randn((Ql, 128), device="cuda", generator=g).to(dt)
            y = torch.randint(1, N, (Ql,), device="cuda", generator=g, dtype=torch.int32)
            want = nccl.score(h, y, k=kk)
            got = peer.score(h, y, k=kk)
            torch.cuda.synchronize()
            peer.check()
            assert set(got) == set(want), (sorted(got), sorted(want))
            for name in want:
                ok &= bool(torch.equal(got[name], want[name]))
        assert peer.buf is not None                      # the peer path ran (no silent fall-back to the collectives)
        peer.close()
        ret[rank] = bool(ok)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("precision", ["bf16", "f32"])
def test_sharded_catalog_scoring_peer_memory(precision):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    ret = ctx.Manager().dict()
    port = _free_port()
    procs = [ctx.Process(target=_peer_worker, args=(r, 2, port, precision, ret)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(300)
        assert p.exitcode == 0
    assert all(ret.get(r) for r in range(2)), dict(ret)


def _train_worker(rank, world, port, path, ret):
    """data-parallel training: each rank owns half of the users; after 3 Adam steps every rank must hold the weights a
    single process gets from the whole batch (gradients and the user count are all-reduced, hiertcn_b200.train) -- with the
    exchange as one peer-memory kernel fused with Adam (htcn_peer_allreduce_adam) or as NCCL all-reduces + htcn_adam_step."""
    import sys
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    if path == "nccl":
        os.environ["HTCN_TRAIN_NCCL"] = "1"
    else:
        os.environ.pop("HTCN_TRAIN_NCCL", None)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
        from helpers import small_case
        from hiertcn_b200.args import make_args
        from hiertcn_b200.data_loader import synthetic_batch
        from hiertcn_b200.model_hier import HierTCN
        from hiertcn_b200.train import HierTCNTrainer
        N, B = 503, 12
        _, _, _, s0, w = small_case(B=B, S=3, L=6, N=N, seed=11, kernel_scale=1.0)
        a = make_args(["--item_num", str(N)])
        batches = [synthetic_batch(B, 3, 6, N, seed=40 + i, lengths="ragged", id_dist="uniform", mask_keep=0.7) for i in range(3)]
        lo, hi = rank * B // world, (rank + 1) * B // world
        tr = HierTCNTrainer(HierTCN(a, w, precision="f32").build(), learning_rate=1e-2, dist=dist, world=world)
        assert (tr.peer is not None) == (path == "peer")           # no silent fall-back to the collectives
        ref = HierTCNTrainer(HierTCN(a, w, precision="f32").build(), learning_rate=1e-2)
        st_dp, st_ref, ok = s0[lo:hi], s0, True
        for x, y, m in batches:
            o = tr.train_step([v[lo:hi] for v in x], [v[lo:hi] for v in y], [v[lo:hi] for v in m], st_dp)
            r = ref.train_step(x, y, m, st_ref)
            st_dp, st_ref = o["state"], r["state"]
            ok &= abs(o["loss"] - r["loss"]) <= 1e-5 * abs(r["loss"]) and o["user_count"] == r["user_count"]
            ok &= bool(np.allclose(st_dp, st_ref[lo:hi], rtol=1e-4, atol=1e-5))
        wd, wr = tr.state_dict(), ref.state_dict()
        budget = 3 * 1e-2
        for k in wr:      # atomics reorder fp32 sums, and Adam turns a sign flip of a ~0 gradient into an lr-sized move
            d = np.abs(wd[k] - wr[k])
            ok &= float(np.mean(d)) <= 0.01 * budget and float(np.mean(d > 0.25 * budget)) < 0.005
        # the replicas hold the same bits: every rank applied the same reduced gradient
        cs = torch.stack([tr.params.double().sum(), tr.params.double().abs().sum()])
        lo_, hi_ = cs.clone(), cs.clone()
        dist.all_reduce(lo_, op=dist.ReduceOp.MIN)
        dist.all_reduce(hi_, op=dist.ReduceOp.MAX)
        ok &= bool(torch.equal(lo_, hi_))
        tr.check_peer()
        ret[rank] = bool(ok)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("path", ["peer", "nccl"])
def test_data_parallel_training_nccl(path):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    ret = ctx.Manager().dict()
    port = _free_port()
    procs = [ctx.Process(target=_train_worker, args=(r, 2, port, path, ret)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(300)
        assert p.exitcode == 0
    assert all(ret.get(r) for r in range(2)), dict(ret)
