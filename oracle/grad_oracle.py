"""Gradient / optimiser oracle for the training step (TEST INFRASTRUCTURE ONLY -- see oracle/__init__.py).

The reference trains with ``tf.train.AdamOptimizer(lr).minimize(loss)`` (model.py:134-141) on the loss of
model.py:105-117; TensorFlow's autodiff is not restated anywhere in the reference sources, so the gradient
oracle is torch autograd (float64, CPU) over a restatement of the forward whose VALUE is pinned: ``loss_fp64``
below must agree with oracle/hiertcn_oracle.forward_loss_metrics (checked against the golden vectors that the
reference's own python produced, tests/test_oracle.py), in the literal form (one-hot x table matmuls, state
concat, S unrolled cell calls -- model.py:59-61, model_hier.py:39-94) and in the restructured form the kernels
follow.  Gradients of the two forms agree to fp64 round-off, and both are checked against central finite
differences in tests/test_oracle.py.  Parity status: values pinned, TF's Adam formula restated from its
published definition -- "parity unpinned" for the optimiser.

The carried user state enters through a placeholder (model.py:44, run_hier_xing.py:291), so no gradient flows
into ``state`` -- backpropagation through time is truncated at the batch boundary.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F

PARAM_ORDER = ("hier/emb/kernel", "hier/emb/bias", "hier/tcn/emb/kernel", "hier/tcn/dense/kernel", "hier/tcn/dense/bias")


def _params(w, dtype=torch.float64):
    return {k: torch.tensor(np.asarray(v), dtype=dtype, requires_grad=True) for k, v in w.items()}


def _gru(p, x, state, G, H):
    """customed_gru_cell.py:309-337 (cell), :1050-1073 (stacking, state_is_tuple=False)"""
    new, cur = [], x
    for g in range(G):
        pre = f"hier/multi_rnn_cell/cell_{g}/gru_cell/"
        h = state[:, g * H:(g + 1) * H]
        v = torch.sigmoid(torch.cat([cur, h], 1) @ p[pre + "gates/kernel"] + p[pre + "gates/bias"])
        r, u = v[:, :H], v[:, H:]
        c = torch.tanh(torch.cat([cur, r * h], 1) @ p[pre + "candidate/kernel"] + p[pre + "candidate/bias"])
        cur = u * h + (1 - u) * c
        new.append(cur)
    return torch.cat(new, 1)


def _tcn(p, h, n_levels, K, drop=None):
    """customized_tcn_cell.py:46-49,109-127,157-161: one conv per block, relu inside and after the residual.
    ``drop`` [n_levels, C]: dropout scales (0 or 1/keep) of this slot's TemporalBlocks -- tf.layers.Dropout with
    noise_shape [1,1,C] on relu(conv) before the residual add (customized_tcn_cell.py:100,119)"""
    for lvl in range(n_levels):
        pre = f"hier/tcn/temporal_conv_net/tblock_{lvl}/conv1/"
        d = 2 ** lvl
        xp = F.pad(h.transpose(1, 2), ((K - 1) * d, 0))
        a = F.conv1d(xp, p[pre + "kernel"].permute(2, 1, 0), p[pre + "bias"], dilation=d).transpose(1, 2)
        ds = f"hier/tcn/temporal_conv_net/tblock_{lvl}/dense/"
        res = h @ p[ds + "kernel"] + p[ds + "bias"] if (ds + "kernel") in p else h      # customized_tcn_cell.py:102-106,123-124
        a = torch.relu(a)
        if drop is not None:
            a = a * torch.as_tensor(np.asarray(drop[lvl])[:a.shape[-1]], dtype=a.dtype)
        h = torch.relu(a + res)
    return h


def _sampled_rows(hout, table, y_id, neg_bt, kind, delta=0.1, nce_weight=1.0, num_neg=20):
    """reference loss.py:22-71 with pred = the user embedding, y = the target's row of ``table`` and y_impression = the rows
    of the sampled negatives (id 0 -> zero row).  hout [B,T,d], y_id [B,T], neg_bt [B,T,k] -> loss [B,T]"""
    mask0 = lambda ids: (ids > 0).unsqueeze(-1).to(hout.dtype)  # noqa: E731
    y = table[y_id] * mask0(y_id)
    yi = table[neg_bt] * mask0(neg_bt)
    ss = (hout * hout).sum(-1, keepdim=True)
    ph = hout * torch.rsqrt(torch.clamp(ss, min=1e-12))                         # tf.nn.l2_normalize
    inner = (ph * y).sum(-1)
    ip = torch.einsum("btd,btkd->btk", ph, yi)
    ls = F.logsigmoid
    if kind == "nce":
        return -ls(inner) - ls(-ip).sum(2) / num_neg * nce_weight
    if kind == "hinge_sigmoid":
        return torch.relu(torch.sigmoid(ip) - torch.sigmoid(inner).unsqueeze(-1) + delta).mean(2)
    if kind == "hinge_logsigmoid":
        return torch.relu(ls(ip) - ls(inner).unsqueeze(-1) + delta).mean(2)
    if kind == "hinge_linear":
        return torch.relu(ip - inner.unsqueeze(-1) + delta).mean(2)
    if kind == "bpr":
        return -ls(torch.sigmoid(inner).unsqueeze(-1) - torch.sigmoid(ip)).mean(2)
    raise NotImplementedError(kind)


def loss_fp64(p, x_list, y_list, mask_list, state, num_layer=2, literal=False, dropout_scales=None, neg_ids=None,
              loss_kind="hinge_logsigmoid", hinge_delta=0.1, nce_weight=1.0, num_neg_sample=20):
    """Scalar training loss (model.py:105-117) as a torch expression of the parameter dict ``p``."""
    dt = p["hier/emb/kernel"].dtype
    E, be = p["hier/emb/kernel"], p["hier/emb/bias"]
    N, H = E.shape[0], p["hier/multi_rnn_cell/cell_0/gru_cell/candidate/kernel"].shape[1]
    n_levels = sum(1 for k in p if k.endswith("conv1/kernel"))
    K = p["hier/tcn/temporal_conv_net/tblock_0/conv1/kernel"].shape[0]
    state = torch.tensor(np.asarray(state), dtype=dt)
    houts = []
    for s in range(len(x_list)):
        x = torch.from_numpy(np.asarray(x_list[s]).astype(np.int64))
        y = torch.from_numpy(np.asarray(y_list[s]).astype(np.int64))
        if literal:
            ohx = F.one_hot(x, N).to(dt) * torch.sign(x).unsqueeze(-1).to(dt)
            ohy = F.one_hot(y, N).to(dt) * torch.sign(y).unsqueeze(-1).to(dt)
            xe = ohx @ E                                                        # model_hier.py:50
            cnt = torch.sign(ohy.abs().sum(2)).sum(1, keepdim=True)
            ys = (ohy.sum(1) / cnt) @ E + be                                    # model_hier.py:83-85
        else:
            xe = E[x] * (x > 0).unsqueeze(-1)
            ys = (E[y] * (y > 0).unsqueeze(-1)).sum(1) / (y > 0).sum(1, keepdim=True) + be
        feat = state.unsqueeze(1).expand(-1, xe.shape[1], -1)
        h0 = torch.cat([xe, feat], -1) @ p["hier/tcn/emb/kernel"]               # model_hier.py:54-55, model_tcn.py:35
        houts.append(_tcn(p, h0, n_levels, K, None if dropout_scales is None else dropout_scales[s]))
        state = _gru(p, ys, state, num_layer, H) * torch.tensor(np.asarray(mask_list[s]), dtype=dt).reshape(-1, 1)
    hout = torch.cat(houts, 1)
    y_id = torch.from_numpy(np.concatenate([np.asarray(v) for v in y_list], 1).astype(np.int64))
    mask = (y_id > 0).to(dt)
    if neg_ids is not None:          # sampled ranking loss against the rows of the output table (loss.py:22-71)
        neg = np.asarray(neg_ids)
        neg_bt = np.zeros(tuple(y_id.shape) + (neg.shape[1],), np.int64)
        neg_bt[y_id.numpy() > 0] = neg                                          # rows of neg_ids follow the scored positions
        loss_bt = _sampled_rows(hout, p["hier/tcn/dense/kernel"].t(), y_id, torch.from_numpy(neg_bt), loss_kind, hinge_delta,
                                nce_weight, num_neg_sample) * mask
        act = mask.sum(1)
        uc = torch.sign(act).sum()
        return ((loss_bt.sum(1) / (act + 1e-6)).sum() / uc), state
    z = (hout @ p["hier/tcn/dense/kernel"] + p["hier/tcn/dense/bias"]) * mask.unsqueeze(-1)     # model.py:105
    lse = torch.logsumexp(z, -1)
    zy = z.gather(2, y_id.unsqueeze(-1)).squeeze(-1)
    loss_bt = (lse - zy) * mask                                                  # loss.py:20-21, model.py:110
    act = mask.sum(1)
    uc = torch.sign(act).sum()
    return ((loss_bt.sum(1) / (act + 1e-6)).sum() / uc), state                   # model.py:111-117


def loss_and_grads(w, x_list, y_list, mask_list, state, num_layer=2, literal=False, dropout_scales=None, **sampled):
    """-> (loss float, dict name -> fp64 numpy gradient, new state).  ``sampled``: neg_ids / loss_kind / hinge_delta /
    nce_weight / num_neg_sample of the sampled ranking losses"""
    p = _params(w)
    loss, st = loss_fp64(p, x_list, y_list, mask_list, state, num_layer, literal, dropout_scales, **sampled)
    loss.backward()
    g = {k: (v.grad.numpy() if v.grad is not None else np.zeros(v.shape)) for k, v in p.items()}
    return float(loss.detach()), g, st.detach().numpy()


def adam_tf(w, g, m, v, t, lr=1e-2, beta1=0.9, beta2=0.999, eps=1e-8):
    """One step of TF-1.x AdamOptimizer: epsilon outside the bias-corrected root.
    ``t`` is the 1-based step number.  Updates the dicts in place (numpy, any float dtype)."""
    lr_t = lr * np.sqrt(1.0 - beta2 ** t) / (1.0 - beta1 ** t)
    for k in w:
        m[k] = beta1 * m[k] + (1 - beta1) * g[k]
        v[k] = beta2 * v[k] + (1 - beta2) * g[k] * g[k]
        w[k] = w[k] - lr_t * m[k] / (np.sqrt(v[k]) + eps)


def train_steps(w, batches, state0, lr=1e-2, num_layer=2):
    """Run len(batches) Adam steps in fp64 carrying the user state like run_hier_xing.py:291,301.
    Returns (list of losses, final weights fp64, final state)."""
    w = {k: np.asarray(a, np.float64).copy() for k, a in w.items()}
    m = {k: np.zeros_like(a) for k, a in w.items()}
    v = {k: np.zeros_like(a) for k, a in w.items()}
    state, losses = np.asarray(state0, np.float64), []
    for t, (x, y, mk) in enumerate(batches, 1):
        loss, g, state = loss_and_grads(w, x, y, mk, state, num_layer)
        losses.append(loss)
        adam_tf(w, g, m, v, t, lr)
    return losses, w, state
