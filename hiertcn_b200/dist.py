"""Multi-GPU plumbing for the hot path: one process per GPU, torch.distributed (NCCL on the box, gloo in
the CPU tests).  The reference has no distributed code at all; this is the B200-native
addition the north star asks for.

Two shardings:

* users (data-parallel): rows of the batch are independent users (reference data_loader.py:153-156), so each
  rank runs the whole forward + scoring on its own users with a replicated catalog; the only exchange is the
  all-reduce of the loss / metric partial sums (``allreduce_scalars``).
* catalog (config 4, 8M items): rank r owns rows [n0_r, n1_r) of W_out^T.  ``ShardedCatalogScorer.score``:
    1. all-gather the query rows (user embeddings Hout, 256 B/row in bf16) and their target ids,
    2. every rank fills the target logits of the ids it owns; all-reduce(SUM) completes the vector,
    3. local sweep over the shard -> per-row partials (max, sumexp, count) and a local top-k,
    4. all-to-all: each rank receives, for ITS OWN rows, the partials of every shard,
    5. merge: log-sum-exp across shards, rank = sum of counts, k-way top-k merge (score desc, global index asc).

The arithmetic of steps 2, 3 and 5 is delegated to an ``ops`` object: ``CudaScoreOps`` (the sm_100a kernels
through the C ABI) in production; the gloo tests inject a numpy implementation to exercise the choreography
on CPU.
"""
from __future__ import annotations

import numpy as np


def shard_bounds(n_items: int, world: int, align: int = 256):
    """Row ranges of the catalog shards: contiguous, multiples of ``align`` items (whole MMA tiles) except the last."""
    per = -(-n_items // world)
    per = -(-per // align) * align
    b = [min(r * per, n_items) for r in range(world + 1)]
    b[-1] = n_items
    return b


def choose_n_split(Q: int, n_items: int, sms: int = 148) -> int:
    """Catalog split count of a K4 sweep over Q query rows: the grid is ceil(Q/128) row tiles x n_split, one CTA per SM, so
    the sweep takes ceil(tiles * n_split / sms) waves of (catalog / n_split) items each.  Large Q (many waves anyway): 4
    splits, so the CTAs of a wave share a quarter of the catalog in L2 (measured +4 % at config 2).  Small Q (config 4:
    4096 queries = 32 row tiles): the split count that wastes the least of its last wave -- 10 splits are 320 CTAs = 2.16
    waves, i.e. three waves of a tenth each (0.30 of a full sweep per SM), 9 splits are 288 CTAs = two nearly full waves
    of a ninth (0.22): a quarter less time for the same work."""
    tiles_q = max(1, -(-Q // 128))
    tiles_q += tiles_q & 1                                   # CTA pairs (cta_group::2)
    cap = int(max(1, min(32, n_items // 256)))
    if tiles_q * 4 >= 8 * sms:
        return min(4, cap)
    best, best_cost = min(4, cap), None
    for ns in range(min(4, cap), cap + 1):
        waves = -(-tiles_q * ns // sms)
        cost = waves / ns * (1.0 + 0.004 * ns)               # a little per-split overhead: partial rows, pipeline fill
        if best_cost is None or cost < best_cost - 1e-12:
            best, best_cost = ns, cost
    return best


def allreduce_scalars(scalars, dist, world):
    """Global loss/metrics from per-rank ``scalars[8]`` = {loss, r@1, r@5, r@10, mrr, mrp, user_count, n_valid}:
    the per-rank values are means over that rank's users (model.py:116-117), so weight by user_count."""
    if world == 1:
        return scalars
    t = scalars.clone()
    # a rank whose batch holds no scored user reports 0/0 = NaN means with user_count 0: it must contribute nothing
    # (NaN * 0 would poison the global means)
    t[:6] = (scalars[:6] * scalars[6]).where(scalars[6] > 0, scalars.new_zeros(()))
    dist.all_reduce(t)
    t[:6] /= t[6]
    return t


class ShardedCatalogScorer:
    """Catalog-sharded scoring with CE / rank / top-k merged across ranks (see module docstring)."""

    def __init__(self, ops, dist, rank: int, world: int, n_items: int, n_split: int = 1):
        self.ops, self.dist, self.rank, self.world = ops, dist, rank, world
        self.bounds = shard_bounds(n_items, world)
        self.n0, self.n1 = self.bounds[rank], self.bounds[rank + 1]
        self.n_split = n_split
        self.phases = None          # set to [] to have every phase boundary recorded (ops.mark -> CUDA events; bench.py)

    def _mark(self, name):
        if self.phases is not None and hasattr(self.ops, "mark"):
            self.phases.append((name, self.ops.mark()))

    def score(self, hout, y_id, k: int = 0, ce: bool = True, rank_metric: bool = True):
        """hout [Q_local,128], y_id [Q_local] (global ids; may be None when only top-k is wanted).
        Every rank must pass the same Q_local.  Returns dict(loss_row, rank_row, topk_val, topk_idx) for the
        LOCAL rows."""
        o, d, W = self.ops, self.dist, self.world
        Ql = hout.shape[0]
        need_t = (ce or rank_metric) and y_id is not None
        self._mark("start")
        # 1. every shard must see every query row
        gather = (lambda t: t) if W == 1 else (lambda t: o.all_gather_rows(d, t, W))      # world 1: one shard, no exchange
        exchange = (lambda t: t) if W == 1 else (lambda t: o.all_to_all_rows(d, t, W, Ql))
        h_all = gather(hout)                                        # [W*Ql, 128]
        y_all = gather(y_id) if need_t else None
        self._mark("allgather_queries")
        Q = W * Ql
        # 2. target logits: owner fills, all-reduce completes
        zy = None
        if need_t:
            zy = o.zeros_f32(Q)
            o.target_logit(h_all, y_all, self.n0, self.n1, zy)
            self._mark("target_logit")
            if W > 1:
                d.all_reduce(zy)
            self._mark("allreduce_target")
        # 3. local sweep (+ exact redo of the rows whose target-referenced partial sum left fp32 range in THIS shard:
        #    the dominant logit of a row may live in another shard than its target)
        part = o.sweep(h_all, y_all, zy, self.n0, self.n1, k, self.n_split, ce and need_t, rank_metric and need_t)
        if ce and need_t and hasattr(o, "repair_parts"):
            o.repair_parts(h_all, self.n0, self.n1, part["pm"], part["ps"])
        self._mark("sweep")
        # 4. all-to-all of the partials: [n_split, W, Ql, ...] -> every rank gets its own rows from every shard
        out = {}
        merged = {}
        if need_t:
            for name in ("pm", "ps", "pc"):
                if part.get(name) is not None:
                    merged[name] = exchange(part[name])                          # [W*n_split, Ql]
        if k:
            tv = exchange(part["tv"])                                            # [W*n_split, Ql, k]
            ti = exchange(part["ti"])
        self._mark("alltoall_partials")
        # 5. merge
        if need_t:
            zy_local = o.slice_rows(zy, self.rank * Ql, Ql)
            out.update(o.finish(merged.get("pm"), merged.get("ps"), merged.get("pc"), y_id, zy_local))
            out["target_logit"] = zy_local
        if k:
            out.update(o.topk_merge(tv, ti, k))
        self._mark("merge")
        return out

    def phase_ms(self):
        """device milliseconds of every phase recorded since ``phases`` was last reset (call after a synchronize)"""
        out = {}
        ph = self.phases or []
        for (_, e0), (name, e1) in zip(ph[:-1], ph[1:]):
            if name != "start":
                out[name] = out.get(name, 0.0) + e0.elapsed_time(e1)
        return out


class CudaScoreOps:
    """The production ``ops``: device tensors + libhtcn kernels (through the C ABI)."""

    def __init__(self, model):
        import torch
        from . import _cabi as cabi
        self.torch, self.cabi, self.m = torch, cabi, model

    # ---- collectives on device tensors
    def all_gather_rows(self, dist, t, world):
        out = self.torch.empty((world * t.shape[0],) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
        dist.all_gather_into_tensor(out, t.contiguous())
        return out

    def all_to_all_rows(self, dist, part, world, Ql):
        """part [n_split, world*Ql, ...] -> [world*n_split, Ql, ...] holding this rank's rows from every shard"""
        ns = part.shape[0]
        tail = tuple(part.shape[2:])
        send = part.reshape((ns, world, Ql) + tail).transpose(0, 1).contiguous()     # [world, ns, Ql, ...]
        recv = self.torch.empty_like(send)
        dist.all_to_all_single(recv, send)
        return recv.reshape((world * ns, Ql) + tail)

    def zeros_f32(self, n):
        return self.torch.zeros(n, dtype=self.torch.float32, device=self.m.device)

    def mark(self):
        """CUDA event on the compute stream (torch's collectives make the compute stream wait for them, so the span
        between two marks covers the collectives enqueued in between)"""
        e = self.torch.cuda.Event(enable_timing=True)
        e.record(self.torch.cuda.current_stream(self.m.device))
        return e

    def repair_parts(self, h_all, n0, n1, pm, ps):
        m = self.m
        self.cabi.call("htcn_score_ce_repair_shard", h_all.data_ptr(), m.act_dtype, h_all.shape[0], m.wt.data_ptr(),
                       n1 - n0, pm.data_ptr(), ps.data_ptr(), pm.shape[0], None, m.stream_ptr())

    def slice_rows(self, t, start, n):
        return t[start:start + n].contiguous()

    # ---- kernels
    def target_logit(self, h_all, y_all, n0, n1, zy):
        m = self.m
        self.cabi.call("htcn_target_logit", h_all.data_ptr(), m.act_dtype, h_all.shape[0], m.wt.data_ptr(),
                       m.b_out.data_ptr() if m.b_out is not None else None, n1 - n0, n0, y_all.data_ptr(), zy.data_ptr(),
                       m.stream_ptr())

    def sweep(self, h_all, y_all, zy, n0, n1, k, n_split, ce, rank, out=None, sync_overflow=True):
        """CE / rank partials [n_split, Q] and the shard's sorted top-k list [1, Q, k].  ``out``: caller-owned result tensors
        (pm, ps, pc, tv, ti, ovf) to write into instead of fresh ones.  The two-pass top-k reports rows whose candidate list
        overflowed (mass ties at the threshold) in ``ovf``: with ``sync_overflow`` the flag is read here (a host sync) and the
        always-exact heap sweep redoes the list; without, ``out["ovf"]`` is left for the caller to check after its own sync."""
        torch, cabi, m = self.torch, self.cabi, self.m
        Q = h_all.shape[0]
        f32, i32 = torch.float32, torch.int32
        P = lambda t: t.data_ptr() if t is not None else None  # noqa: E731
        given = out or {}
        out = dict(pm=None, ps=None, pc=None, tv=None, ti=None, ovf=None)
        new = lambda name, shape, dt: given[name] if given.get(name) is not None else torch.empty(shape, dtype=dt, device=m.device)  # noqa: E731
        flags = (cabi.SCORE_CE if ce else 0) | (cabi.SCORE_RANK if rank else 0)
        fused = bool(k) and ce and rank and getattr(self, "fuse_topk", True)
        if flags:
            out["pm"] = new("pm", (n_split, Q), f32) if ce else None
            out["ps"] = new("ps", (n_split, Q), f32) if ce else None
            out["pc"] = new("pc", (n_split, Q), i32) if rank else None
        if k:
            out["tv"] = new("tv", (1, Q, k), f32)
            out["ti"] = new("ti", (1, Q, k), i32)
            if given.get("ovf") is not None:
                out["ovf"] = given["ovf"]
                out["ovf"].zero_()
            else:
                out["ovf"] = torch.zeros(1, dtype=i32, device=m.device)
        if fused:
            # loss + rank + top-k in two sweeps of the shard: the CE / rank sweep records the group maxima pass 1 of the
            # two-pass top-k would need a sweep of its own for
            nb = int(cabi.load().htcn_topk_workspace_bytes(m.act_dtype, Q, n1 - n0, k, n_split))
            ws = torch.empty(nb, dtype=torch.uint8, device=m.device)
            cabi.call("htcn_score_ce_rank_topk_fused", h_all.data_ptr(), m.act_dtype, Q, m.wt.data_ptr(), P(m.b_out), n1 - n0,
                      n0, y_all.data_ptr(), zy.data_ptr(), k, n_split, ws.data_ptr(), nb, P(out["pm"]), P(out["ps"]),
                      P(out["pc"]), out["tv"].data_ptr(), out["ti"].data_ptr(), out["ovf"].data_ptr(), m.stream_ptr())
            if sync_overflow and int(out["ovf"].item()):     # mass ties overflowed a candidate list: the always-exact heap sweep
                self._heap_topk(h_all, n0, n1, k, n_split, out)
            return out
        if flags:
            cabi.call("htcn_score_ce_rank_topk", h_all.data_ptr(), m.act_dtype, Q, m.wt.data_ptr(), P(m.b_out),
                      n1 - n0, n0, y_all.data_ptr(), zy.data_ptr(), 1, flags, 0, n_split, P(out["pm"]), P(out["ps"]),
                      P(out["pc"]), None, None, m.stream_ptr())
        if k:
            # exact local top-k of the shard in one call (two tensor-core sweeps on large bf16 shards, heap sweep
            # otherwise); one sorted list per row = one "part" for the cross-shard merge
            nb = int(cabi.load().htcn_topk_workspace_bytes(m.act_dtype, Q, n1 - n0, k, n_split))
            ws = torch.empty(nb, dtype=torch.uint8, device=m.device)
            cabi.call("htcn_score_topk", h_all.data_ptr(), m.act_dtype, Q, m.wt.data_ptr(), P(m.b_out), n1 - n0, n0, k,
                      n_split, ws.data_ptr(), nb, out["tv"].data_ptr(), out["ti"].data_ptr(), out["ovf"].data_ptr(), m.stream_ptr())
            if sync_overflow and int(out["ovf"].item()):
                self._heap_topk(h_all, n0, n1, k, n_split, out)
        return out

    def _heap_topk(self, h_all, n0, n1, k, n_split, out):
        """one k-entry heap per row in shared memory: slower, but exact whatever the ties"""
        torch, cabi, m = self.torch, self.cabi, self.m
        Q = h_all.shape[0]
        tv = torch.empty((n_split, Q, k), dtype=torch.float32, device=m.device)
        ti = torch.empty((n_split, Q, k), dtype=torch.int32, device=m.device)
        cabi.call("htcn_score_ce_rank_topk", h_all.data_ptr(), m.act_dtype, Q, m.wt.data_ptr(),
                  m.b_out.data_ptr() if m.b_out is not None else None, n1 - n0, n0, None, None, 1, cabi.SCORE_TOPK, k, n_split,
                  None, None, None, tv.data_ptr(), ti.data_ptr(), m.stream_ptr())
        cabi.call("htcn_topk_merge", tv.data_ptr(), ti.data_ptr(), n_split, Q, k, out["tv"].data_ptr(), out["ti"].data_ptr(),
                  m.stream_ptr())
        return out

    def finish(self, pm, ps, pc, y_id, zy_local):
        torch, cabi, m = self.torch, self.cabi, self.m
        Ql = zy_local.shape[0]
        n_part = (pm if pm is not None else pc).shape[0]
        P = lambda t: t.data_ptr() if t is not None else None  # noqa: E731
        loss_row = torch.empty(Ql, dtype=torch.float32, device=m.device) if pm is not None else None
        rank_row = torch.empty(Ql, dtype=torch.float32, device=m.device) if pc is not None else None
        cabi.call("htcn_score_finish", P(pm), P(ps), P(pc), n_part, Ql, y_id.data_ptr(), zy_local.data_ptr(),
                  P(loss_row), P(rank_row), m.stream_ptr())
        return dict(loss_row=loss_row, rank_row=rank_row)

    def topk_merge(self, tv, ti, k):
        torch, cabi, m = self.torch, self.cabi, self.m
        n_part, Ql = tv.shape[0], tv.shape[1]
        ov = torch.empty((Ql, k), dtype=torch.float32, device=m.device)
        oi = torch.empty((Ql, k), dtype=torch.int32, device=m.device)
        cabi.call("htcn_topk_merge", tv.data_ptr(), ti.data_ptr(), n_part, Ql, k, ov.data_ptr(), oi.data_ptr(),
                  m.stream_ptr())
        return dict(topk_val=ov, topk_idx=oi)


class CatalogTable:
    """A catalog (shard) resident in HBM for scoring-only use -- BASELINE config 4: user embeddings in, CE / rank / top-k
    out, no encoder.  Holds W_out^T in the scoring layout of ``htcn_prepare_wout`` ([n,144] bf16 with the bias folded in,
    or [n,128] f32 + b_out) and exposes what ``CudaScoreOps`` reads from a model."""

    def __init__(self, wt, b_out=None, precision="bf16", n_items=None):
        from . import _cabi as cabi
        import torch
        cabi.load()
        self.wt, self.b_out, self.precision = wt, b_out, precision
        self.act_dtype = cabi.HTCN_BF16 if precision == "bf16" else cabi.HTCN_F32
        self.device = wt.device
        self.n_out = int(wt.shape[0])
        self.N = int(n_items if n_items is not None else self.n_out)
        self._torch = torch

    def stream_ptr(self):
        return self._torch.cuda.current_stream(self.device).cuda_stream

    def rows(self, n0, n1):
        """view of rows [n0, n1) as its own table (a shard of a replicated catalog shares the memory)"""
        return CatalogTable(self.wt[n0:n1], None if self.b_out is None else self.b_out[n0:n1], self.precision, self.N)


def make_sharded_model(args, weights, rank, world, precision="bf16"):
    """HierTCN whose W_out^T / b_out hold only this rank's catalog shard (rows [n0, n1))."""
    from .model_hier import HierTCN
    n_items = int(args.item_num)
    b = shard_bounds(n_items, world)
    n0, n1 = b[rank], b[rank + 1]
    w = dict(weights)
    w["hier/tcn/dense/kernel"] = np.ascontiguousarray(weights["hier/tcn/dense/kernel"][:, n0:n1])
    w["hier/tcn/dense/bias"] = np.ascontiguousarray(weights["hier/tcn/dense/bias"][n0:n1])
    m = HierTCN(args, w, precision=precision)
    m.out_rows = n1 - n0
    return m.build(), n0, n1
