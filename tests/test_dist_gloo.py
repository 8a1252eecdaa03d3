"""world_size-2 gloo tests (CPU) of the multi-GPU choreography in hiertcn_b200/dist.py: catalog-sharded
scoring (all-gather queries -> all-reduce target logits -> local sweep -> all-to-all partials -> merge) must
equal single-shard scoring, and the data-parallel scalar all-reduce must equal the global means.
The per-shard arithmetic is a numpy stand-in injected through the `ops` interface (the CUDA ops are covered by
tests/test_gpu_kernels.py::test_k4_sharded_catalog_equals_single)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from hiertcn_b200.dist import ShardedCatalogScorer, allreduce_scalars, shard_bounds
from oracle import hiertcn_oracle as O


class NumpyScoreOps:
    """CPU stand-in with the contract of dist.CudaScoreOps (torch CPU tensors in/out)."""

    def __init__(self, w_out_t, b_out):
        self.w, self.b = w_out_t, b_out          # full catalog; shards are sliced by [n0, n1)

    def all_gather_rows(self, d, t, world):
        out = torch.empty((world * t.shape[0],) + tuple(t.shape[1:]), dtype=t.dtype)
        d.all_gather_into_tensor(out, t.contiguous())
        return out

    def all_to_all_rows(self, d, part, world, Ql):
        ns, tail = part.shape[0], tuple(part.shape[2:])
        send = part.reshape((ns, world, Ql) + tail).transpose(0, 1).contiguous()
        recv = torch.empty_like(send)
        d.all_to_all_single(recv, send)
        return recv.reshape((world * ns, Ql) + tail)

    def zeros_f32(self, n):
        return torch.zeros(n, dtype=torch.float32)

    def slice_rows(self, t, start, n):
        return t[start:start + n].contiguous()

    def _z(self, h, n0, n1):
        # fp32-rounded logits everywhere, so a target logit compares EQUAL to itself in the sweep
        return (h.numpy().astype(np.float64) @ self.w[n0:n1].T.astype(np.float64) + self.b[n0:n1]).astype(np.float32)

    def target_logit(self, h_all, y_all, n0, n1, zy):
        y = y_all.numpy()
        own = (y >= n0) & (y < n1)
        z = self._z(h_all, n0, n1)
        zy.numpy()[own] = z[np.nonzero(own)[0], y[own] - n0].astype(np.float32)

    def sweep(self, h_all, y_all, zy, n0, n1, k, n_split, ce, rank):
        z = self._z(h_all, n0, n1)
        Q = z.shape[0]
        cols = np.array_split(np.arange(n1 - n0), n_split)
        out = dict(pm=None, ps=None, pc=None, tv=None, ti=None)
        if ce:
            out["pm"] = torch.tensor(np.stack([z[:, c].max(1) for c in cols]), dtype=torch.float32)
            out["ps"] = torch.tensor(np.stack([np.exp(z[:, c] - z[:, c].max(1, keepdims=True)).sum(1) for c in cols]), dtype=torch.float32)
        if rank:
            t = zy.numpy()[:, None]
            out["pc"] = torch.tensor(np.stack([(z[:, c] > t).sum(1) for c in cols]), dtype=torch.int32)
        if k:
            tv, ti = [], []
            for c in cols:
                v, i = O.top_k(z[:, c], min(k, len(c)))
                pad = k - v.shape[1]
                tv.append(np.pad(v, ((0, 0), (0, pad)), constant_values=-np.inf))
                ti.append(np.pad(i + n0 + c[0], ((0, 0), (0, pad)), constant_values=-1))
            out["tv"] = torch.tensor(np.stack(tv), dtype=torch.float32)
            out["ti"] = torch.tensor(np.stack(ti), dtype=torch.int32)
        return out

    def finish(self, pm, ps, pc, y_id, zy_local):
        out = {}
        if pm is not None:
            M = pm.max(0).values
            s = (ps * torch.exp(pm - M)).sum(0)
            out["loss_row"] = (M + torch.log(s)) - zy_local
        if pc is not None:
            out["rank_row"] = pc.sum(0).to(torch.float32)
        return out

    def topk_merge(self, tv, ti, k):
        v = tv.permute(1, 0, 2).reshape(tv.shape[1], -1).numpy().astype(np.float64)
        i = ti.permute(1, 0, 2).reshape(ti.shape[1], -1).numpy().astype(np.int64)
        ov, oi = np.empty((v.shape[0], k), np.float32), np.empty((v.shape[0], k), np.int32)
        for r in range(v.shape[0]):
            order = np.lexsort((np.where(i[r] < 0, 2 ** 40, i[r]), -v[r]))[:k]
            ov[r], oi[r] = v[r][order], i[r][order]
        return dict(topk_val=torch.tensor(ov), topk_idx=torch.tensor(oi))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(0)                      # same data on every rank
        N, Ql, k = 1500, 24, 20
        w = (rng.normal(size=(N, 128)) * 0.3).astype(np.float32)
        b = (rng.normal(size=N) * 0.2).astype(np.float32)
        h_all = rng.normal(size=(world * Ql, 128)).astype(np.float32)
        y_all = rng.integers(1, N, size=world * Ql).astype(np.int32)
        sc = ShardedCatalogScorer(NumpyScoreOps(w, b), dist, rank, world, N, n_split=3)
        h = torch.tensor(h_all[rank * Ql:(rank + 1) * Ql])
        y = torch.tensor(y_all[rank * Ql:(rank + 1) * Ql])
        out = sc.score(h, y, k=k)
        # single-shard truth for this rank's rows
        z = (h.numpy().astype(np.float64) @ w.T.astype(np.float64) + b).astype(np.float32)
        loss = O.softmax_cross_entropy_with_logits(y.numpy(), z.astype(np.float64))
        rank_ref = (z > z[np.arange(Ql), y.numpy()][:, None]).sum(1)
        v_ref, i_ref = O.top_k(z, k)
        ok = (np.allclose(out["loss_row"].numpy(), loss, rtol=1e-5, atol=1e-5)
              and np.array_equal(out["rank_row"].numpy(), rank_ref)
              and np.array_equal(out["topk_idx"].numpy(), i_ref)
              and np.allclose(out["topk_val"].numpy(), v_ref, rtol=1e-6))
        # data-parallel scalar reduction: per-rank means weighted by user_count
        s = torch.tensor([1.0 + rank, .1, .2, .3, .4, .5, 3.0 + rank, 10.0])
        g = allreduce_scalars(s, dist, world)
        exp_loss = sum((1.0 + r) * (3.0 + r) for r in range(world)) / sum(3.0 + r for r in range(world))
        ok = ok and abs(float(g[0]) - exp_loss) < 1e-6 and float(g[6]) == sum(3.0 + r for r in range(world))
        # a rank whose batch holds no scored user reports NaN means (0/0) with user_count 0: it must not poison the result
        nan = float("nan")
        s0 = torch.tensor([nan] * 6 + [0.0, 0.0]) if rank == 0 else torch.tensor([2.0, .1, .2, .3, .4, .5, 4.0, 9.0])
        g0 = allreduce_scalars(s0, dist, world)
        ok = ok and bool(torch.isfinite(g0).all()) and abs(float(g0[0]) - 2.0) < 1e-6 and float(g0[6]) == 4.0 * (world - 1)
        ret[rank] = bool(ok)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_scoring_choreography_gloo(world):
    ctx = mp.get_context("spawn")
    ret = ctx.Manager().dict()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, ret)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert all(ret.get(r) for r in range(world)), dict(ret)


def test_shard_bounds():
    b = shard_bounds(1_000_000, 8)
    assert b[0] == 0 and b[-1] == 1_000_000 and all(x % 256 == 0 for x in b[:-1]) and sorted(b) == b
    assert shard_bounds(100, 3) == [0, 100, 100, 100]
