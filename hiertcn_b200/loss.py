"""The reference's loss.py surface on top of the fused kernels (reference loss.py:5-221).

The reference functions take the materialised ``pred [B,T,N]`` tensor; here ``pred`` is the lazy
``CatalogScores`` returned by ``model_hier`` / ``HierTCN.forward`` -- the logits are never written to HBM and every
function below is a view on ONE streaming sweep over the catalog (cached on the handle).

    calc_loss(pred, y)                          -> loss [B,T]   softmax cross-entropy (loss.py:20-21), masked
    calc_loss(pred, y, y_impression=neg_ids)    -> loss rows    sampled ranking loss selected by args.loss (loss.py:22-71)
    calc_score(pred, cand_ids)                  -> score [Q,k]  'l2' / 'inner_prod' (loss.py:76-105)
    calc_metric_fast(pred, ...)                 -> the reference's 7-tuple (loss.py:163-221)
    top_k(pred, k)                              -> (values, indices) [Q,k] in tf.nn.top_k order (loss.py:120)
"""
from __future__ import annotations

import numpy as np

from . import _cabi as cabi


def _torch():
    import torch
    return torch


def calc_loss(pred, y=None, y_impression=None):
    """reference loss.py:5.  ``y`` is implied by the handle (the targets the forward was staged with); pass
    ``y_impression`` = negative item ids [Q,k] to get the sampled ranking loss of ``args.loss`` per scored row."""
    m = pred.model
    if y_impression is not None:
        return m.sampled_loss(pred, y_impression)
    return m.loss(pred, metrics=False, per_position=True)["loss_bt"]


def calc_score(pred, y_impression, rank_metric=None):
    """reference loss.py:76: scores of the candidate rows ``y_impression`` [Q,k] (item ids) for every scored row."""
    torch = _torch()
    m = pred.model
    mode = {"l2": 0, "inner_prod": 1}[rank_metric or m.args.rank_metric]
    if m.wt_f32 is None:
        m.wt_f32 = m.wt[:, :128].float().contiguous()
    cand = y_impression if hasattr(y_impression, "data_ptr") else \
        torch.from_numpy(np.ascontiguousarray(y_impression, np.int32)).to(m.device)
    Q, k = cand.shape
    out = torch.empty((Q, k), dtype=torch.float32, device=m.device)
    cabi.call("htcn_calc_score", pred.hout.data_ptr(), m.act_dtype, Q, m.wt_f32.data_ptr(), cand.data_ptr(), k, mode,
              out.data_ptr(), m.stream_ptr())
    return out


def calc_metric_fast(pred, mask_y=None, activity_count=None, user_count=None, y=None):
    """reference loss.py:163: (recall@1, recall@5, recall@10, mrr, mrp, ranks_float [B,T], ranks [B,T]); the mask and the
    counts are recomputed from the staged targets exactly as model.py:62,112-114 does."""
    r = pred.model.loss(pred, metrics=True, per_position=True)
    sc = r["scalars"]
    return sc[1], sc[2], sc[3], sc[4], sc[5], r["ranks_float"], r["ranks"]


def top_k(pred, k):
    """tf.nn.top_k(score, k) over the catalog for every scored row (reference loss.py:120)."""
    t = pred.model.topk(pred.hout, pred.Q, k)
    return t["topk_val"], t["topk_idx"]
