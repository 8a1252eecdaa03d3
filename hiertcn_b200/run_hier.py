"""Evaluation loop of the hot path's caller, mirroring reference run_hier_xing.py:83-207 (``evaluate_hier``):
loop over a hierarchical loader until it reports ``done``, carry the user state across batches (on the device
here; the reference round-trips it through host numpy, run_hier_xing.py:241,291,301), accumulate loss / MRP /
MRR / recall@{1,5,10} per batch and report their means, plus the per-position and per-user rank analyses
(run_hier_xing.py:11-31, 59-79) computed from the ``ranks_float`` map the scoring kernel emits.

The training half of the same function (backward + Adam, run_hier_xing.py:257-307) is hiertcn_b200.train.run_hier.
"""
from __future__ import annotations

import time

import numpy as np


def ranks_analysis(ranks_float, mask_y, max_len):
    """Mean rank percentile per position inside the batch window (run_hier_xing.py:11-31)."""
    T = min(ranks_float.shape[1], max_len)
    s = (ranks_float[:, :T] * mask_y[:, :T]).sum(0)
    n = mask_y[:, :T].sum(0)
    return s, n


def ranks_user_analysis(ranks_float, mask_y):
    """Per-user mean rank percentile of the batch (run_hier_xing.py:59-79)."""
    n = mask_y.sum(1)
    with np.errstate(invalid="ignore", divide="ignore"):
        return np.where(n > 0, (ranks_float * mask_y).sum(1) / np.maximum(n, 1), np.nan)


def evaluate_hier(model, loader, max_batches=None, results_path=None, verbose=False):
    """Run ``model`` (hiertcn_b200.model_hier.HierTCN) over ``loader`` (get_batch() -> x_list, y_list, mask_list,
    info_list; ``.done`` flag) and return the epoch means the reference writes to results.txt
    (run_hier_xing.py:152-161,197-201)."""
    keys = ("loss", "mrp", "mrr", "recall1", "recall5", "recall10")
    acc = {k: 0.0 for k in keys}
    pos_sum = pos_cnt = None
    user_means = []
    state = None
    n_batches = 0
    t_load = t_run = 0.0
    while True:
        t0 = time.time()
        if hasattr(loader, "next"):                              # DeviceBatcher: the batch is assembled in HBM
            staged = loader.next(state)
            t1 = time.time()
            out = model.step(staged=staged, per_position=("ranks_float",), state_on_device=True)
            y_list = [staged["y_id"].cpu().numpy()]
        else:
            x_list, y_list, mask_list, _ = loader.get_batch()
            t1 = time.time()
            out = model.step(x_list, y_list, mask_list, state, per_position=("ranks_float",), state_on_device=True)
        state = out["state"]                                     # stays in HBM between batches
        t2 = time.time()
        t_load += t1 - t0
        t_run += t2 - t1
        for k in keys:
            acc[k] += float(out[k])
        y_id = np.concatenate([np.asarray(y) for y in y_list], 1)
        mask_y = (y_id > 0).astype(np.float32)
        s, n = ranks_analysis(out["ranks_float"], mask_y, out["ranks_float"].shape[1])
        if pos_sum is None:
            pos_sum, pos_cnt = np.zeros(0), np.zeros(0)
        if len(s) > len(pos_sum):                                # batches differ in T = sum of per-slot max lengths
            pos_sum = np.concatenate([pos_sum, np.zeros(len(s) - len(pos_sum))])
            pos_cnt = np.concatenate([pos_cnt, np.zeros(len(n) - len(pos_cnt))])
        pos_sum[:len(s)] += s
        pos_cnt[:len(n)] += n
        user_means.append(ranks_user_analysis(out["ranks_float"], mask_y))
        n_batches += 1
        if verbose:
            print("batch %d loss %.4f mrr %.4f" % (n_batches, out["loss"], out["mrr"]))
        if getattr(loader, "done", False) or (max_batches is not None and n_batches >= max_batches):
            break
    res = {k: acc[k] / n_batches for k in keys}
    res.update(batches=n_batches, time_load=t_load, time_run=t_run,
               rank_by_position=(pos_sum / np.maximum(pos_cnt, 1)).tolist(),
               user_rank_mean=float(np.nanmean(np.concatenate(user_means))))
    if results_path:
        with open(results_path, "a") as f:                      # run_hier_xing.py:197-201
            f.write("loss %.6f mrp %.6f mrr %.6f recall1 %.6f recall5 %.6f recall10 %.6f batches %d\n"
                    % (res["loss"], res["mrp"], res["mrr"], res["recall1"], res["recall5"], res["recall10"], n_batches))
    return res
