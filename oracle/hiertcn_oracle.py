"""numpy restatement of the HierTCN hot path (TEST INFRASTRUCTURE -- see oracle/__init__.py).

Parity status: TF op semantics unpinned (no TensorFlow here, no golden vectors in the
reference); control flow pinned by running the reference's own python under
``oracle/tf_shim`` (``oracle/make_golden.py``).

Every function cites the reference file:line (relative to the reference checkout) it follows.
Two routes are provided and must agree (tests/test_oracle.py):

* ``model_hier_literal``      -- op-for-op mirror of the TF graph: one-hot x table matmuls,
                                 tile+concat of the state, interleaved TCN/GRU session loop,
                                 materialised ``[B,T,N]`` logits.
* ``model_hier_restructured`` -- the algebra the CUDA path relies on: gather instead of one-hot
                                 matmul, GRU chain hoisted out of the session loop (its input is
                                 teacher-forced), in-projection split into ``Xe.W_in[:D]`` plus a
                                 per-(user,session) bias ``state.W_in[D:]``.

``precision`` is one of ``"f32"``, ``"f64"`` or ``"bf16"`` (operands of the tensor-core GEMMs
rounded to bfloat16, fp32 accumulate -- mirrors the roundings of the sm_100a bf16 tier).
"""
from __future__ import annotations

import numpy as np

# --------------------------------------------------------------------------------------
# helpers
# --------------------------------------------------------------------------------------


def bf16_round(a: np.ndarray) -> np.ndarray:
    """Round fp32 -> bfloat16 (round-to-nearest-even), returned as fp32."""
    a = np.ascontiguousarray(a, dtype=np.float32)
    u = a.view(np.uint32).astype(np.uint64)
    lsb = (u >> 16) & 1
    r = ((u + 0x7FFF + lsb) >> 16) << 16
    out = (r & 0xFFFFFFFF).astype(np.uint32).view(np.float32)
    nan = np.isnan(a)
    if nan.any():
        out = out.copy()
        out[nan] = np.nan
    return out.reshape(a.shape)


def _dt(precision: str):
    return np.float64 if precision == "f64" else np.float32


def _q(a: np.ndarray, precision: str) -> np.ndarray:
    """Quantise a GEMM operand for the given tier."""
    if precision == "bf16":
        return bf16_round(a)
    return a.astype(_dt(precision), copy=False)


def sigmoid(x):
    return 1.0 / (1.0 + np.exp(-x))


def relu(x):
    return np.maximum(x, 0)


# --------------------------------------------------------------------------------------
# A.2 embedding (model.py:54-63 one-hot * sign(id); model_hier.py:50,83-85 dense 'emb')
# --------------------------------------------------------------------------------------


def one_hot_signed(ids: np.ndarray, depth: int, dtype=np.float32) -> np.ndarray:
    """model.py:56-61 -- tf.one_hot(id, N) * sign(id): id 0 gives an all-zero row."""
    ids = np.asarray(ids).astype(np.int64)
    oh = np.zeros(ids.shape + (depth,), dtype=dtype)
    np.put_along_axis(oh, ids[..., None], 1.0, axis=-1)
    oh *= np.sign(ids)[..., None].astype(dtype)
    return oh


def emb_gather(ids: np.ndarray, table: np.ndarray) -> np.ndarray:
    """Restructured form of one_hot_signed(ids) @ table: row gather, id 0 -> zeros (bit-exact copy)."""
    ids = np.asarray(ids).astype(np.int64)
    out = table[ids]
    out[ids == 0] = 0
    return out


def meanpool_emb(y_ids: np.ndarray, table: np.ndarray, bias: np.ndarray) -> np.ndarray:
    """model_hier.py:83-85 restructured: mean over valid positions of E[y] plus emb/bias.

    Sequential left-to-right fp accumulation over t (the order the CUDA kernel uses).
    n == 0 yields NaN exactly like the reference's 0/0.
    """
    y_ids = np.asarray(y_ids).astype(np.int64)
    B, L = y_ids.shape
    acc = np.zeros((B, table.shape[1]), dtype=table.dtype)
    for t in range(L):
        acc = acc + emb_gather(y_ids[:, t], table)
    n = (y_ids > 0).sum(1).astype(table.dtype)[:, None]
    with np.errstate(invalid="ignore", divide="ignore"):
        return acc / n + bias


# --------------------------------------------------------------------------------------
# dense / conv (customized_dense_layer.py:155-172, customized_convolution_layer.py:171-198,
# customized_tcn_cell.py:46-49)
# --------------------------------------------------------------------------------------


def dense(x: np.ndarray, kernel: np.ndarray, bias=None) -> np.ndarray:
    """customized_dense_layer.py:155-172 -- tensordot over the last axis, then bias_add."""
    out = x @ kernel
    if bias is not None:
        out = out + bias
    return out


def causal_conv1d(x: np.ndarray, kernel: np.ndarray, bias: np.ndarray, dilation: int,
                  precision: str = "f32") -> np.ndarray:
    """customized_tcn_cell.py:46-49 (left pad (K-1)*d) + customized_convolution_layer.py:171-198
    (VALID dilated cross-correlation, kernel [K,Cin,Cout], bias_add, relu activation).

    out[b,t,:] = relu(bias + sum_k x[b, t-(K-1-k)*d, :] @ W[k]),   x[b, tau<0] = 0
    """
    B, L, Cin = x.shape
    K = kernel.shape[0]
    xq = _q(x, precision)
    wq = _q(kernel, precision)
    acc_dt = np.float64 if precision == "f64" else np.float32
    pad = (K - 1) * dilation
    xp = np.zeros((B, L + pad, Cin), dtype=xq.dtype)
    xp[:, pad:, :] = xq
    out = np.zeros((B, L, kernel.shape[2]), dtype=acc_dt)
    for k in range(K):
        out += xp[:, k * dilation:k * dilation + L, :].astype(acc_dt) @ wq[k].astype(acc_dt)
    out = out + bias.astype(acc_dt)
    return relu(out)


def temporal_block(x, w, prefix: str, dilation: int, precision="f32"):
    """customized_tcn_cell.py:109-127 -- ONE conv per block (conv2 is built but never called),
    relu inside the conv and again after the residual add; dropout rate 0 = identity; no norm."""
    a = causal_conv1d(x, w[prefix + "/conv1/kernel"], w[prefix + "/conv1/bias"], dilation, precision)
    if (prefix + "/dense/kernel") in w:                      # :102-106,123-124 (Cin != Cout)
        res = dense(_q(x, precision), _q(w[prefix + "/dense/kernel"], precision)) + w[prefix + "/dense/bias"]
    else:
        res = x
    return relu(a + res)


def temporal_conv_net(x, w, scope: str, n_levels: int, precision="f32", round_between=False):
    """customized_tcn_cell.py:147-161 -- levels chained with dilation 2**level."""
    h = x
    for lvl in range(n_levels):
        h = temporal_block(h, w, f"{scope}/temporal_conv_net/tblock_{lvl}", 2 ** lvl, precision)
        if round_between and precision == "bf16":
            h = bf16_round(h)            # the bf16 tier keeps level activations as bf16 on chip
    return h


def n_tcn_levels(w, scope: str) -> int:
    n = 0
    while f"{scope}/temporal_conv_net/tblock_{n}/conv1/kernel" in w:
        n += 1
    return n


def model_tcn(x, w, scope="hier/tcn", precision="f32", with_head=True):
    """model_tcn.py:26-44 -- 'emb' in-projection (no bias) -> TemporalConvNet -> dense to N logits."""
    h0 = dense(_q(x, precision), _q(w[scope + "/emb/kernel"], precision))
    if precision == "bf16":
        h0 = bf16_round(h0)
    h = temporal_conv_net(h0, w, scope, n_tcn_levels(w, scope), precision, round_between=True)
    if not with_head:
        return h
    return dense(_q(h, precision), _q(w[scope + "/dense/kernel"], precision), w[scope + "/dense/bias"])


# --------------------------------------------------------------------------------------
# A.4 GRU (customed_gru_cell.py:309-337 GRUCell.call, :1050-1073 MultiRNNCell.call,
# :1187-1197 _Linear: concat(args) @ W + b)
# --------------------------------------------------------------------------------------


def gru_cell(inp, h, wg, bg, wc, bc):
    """customed_gru_cell.py:309-337: [r,u] = sigmoid([x,h]Wg+bg) (r first); c = tanh([x,r*h]Wc+bc);
    h' = u*h + (1-u)*c."""
    H = h.shape[1]
    value = sigmoid(np.concatenate([inp, h], 1) @ wg + bg)
    r, u = value[:, :H], value[:, H:]
    c = np.tanh(np.concatenate([inp, r * h], 1) @ wc + bc)
    return u * h + (1 - u) * c


def multi_rnn_cell(inp, state, w, scope="hier", num_layer=2):
    """customed_gru_cell.py:1050-1073 with state_is_tuple=False: state = concat of layer states,
    layer g+1 input = layer g new state."""
    H = state.shape[1] // num_layer
    cur = inp
    new_states = []
    for g in range(num_layer):
        p = f"{scope}/multi_rnn_cell/cell_{g}/gru_cell"
        h = state[:, g * H:(g + 1) * H]
        cur = gru_cell(cur, h, w[p + "/gates/kernel"], w[p + "/gates/bias"],
                       w[p + "/candidate/kernel"], w[p + "/candidate/bias"])
        new_states.append(cur)
    return cur, np.concatenate(new_states, 1)


# --------------------------------------------------------------------------------------
# model_hier (model_hier.py:21-94)
# --------------------------------------------------------------------------------------


def _cast_weights(w, precision):
    dt = _dt(precision)
    return {k: np.asarray(v).astype(dt) for k, v in w.items()}


def model_hier_literal(x_ids, y_ids, masks, state, w, num_layer=2, precision="f32", x_gap=None, gap_bandwidth=168.0,
                       l2_norm=False):
    """Literal mirror of model.py:59-61 + model_hier.py:39-94 (feasible for small N only).

    x_ids / y_ids: S-lists of int arrays [B, L_s]; masks: S-list of [B,1]; state [B, G*H].
    Returns (pred_all [B,T,N], state [B,G*H]).
    """
    assert precision in ("f32", "f64")
    w = _cast_weights(w, precision)
    dt = _dt(precision)
    state = np.asarray(state).astype(dt)
    N = w["hier/emb/kernel"].shape[0]
    preds = []
    for s in range(len(x_ids)):
        if x_gap is not None:                                                 # model_hier.py:40-47 (train_gap off)
            state = state * np.exp(-np.asarray(x_gap[s]).astype(dt) / dt(gap_bandwidth))
        x = one_hot_signed(x_ids[s], N, dt)                                   # model.py:59
        y = one_hot_signed(y_ids[s], N, dt)                                   # model.py:60
        x_slice = dense(x, w["hier/emb/kernel"])                              # model_hier.py:50
        feat = np.tile(state[:, None, :], (1, x_slice.shape[1], 1))           # :54
        x_slice = np.concatenate([x_slice, feat], -1)                         # :55
        p = model_tcn(x_slice, w, "hier/tcn", precision)                     # :63
        preds.append(l2_normalize(p) if l2_norm else p)                      # model_tcn.py:42-43
        cnt = np.sign(np.abs(y).sum(2)).sum(1, keepdims=True)                 # :83
        with np.errstate(invalid="ignore", divide="ignore"):
            y_slice = y.sum(1) / cnt                                          # :84
        y_slice = dense(y_slice, w["hier/emb/kernel"], w["hier/emb/bias"])    # :85
        _, state = multi_rnn_cell(y_slice, state, w, "hier", num_layer)       # :91
        state = state * np.asarray(masks[s]).astype(dt)                       # :93
    return np.concatenate(preds, 1), state                                   # :76-79,94


def gru_over_sessions(y_ids, masks, state0, w, num_layer=2, precision="f32", x_gap=None, gap_bandwidth=168.0):
    """Hoisted recurrence: returns (state_pre [S,B,GH] = state seen by session s's TCN,
    state_out [B,GH], Yp [S,B,D]).  Valid because the GRU input is teacher-forced
    (model_hier.py:83-91) and the TCN output never feeds it."""
    dt = _dt(precision)
    E, be = w["hier/emb/kernel"].astype(dt), w["hier/emb/bias"].astype(dt)
    state = np.asarray(state0).astype(dt)
    pre, yps = [], []
    wc = _cast_weights({k: v for k, v in w.items() if "multi_rnn_cell" in k}, precision)
    for s in range(len(y_ids)):
        if x_gap is not None:                                                 # model_hier.py:40-47: decay before the slot
            state = state * np.exp(-np.asarray(x_gap[s]).astype(dt) / dt(gap_bandwidth))
        pre.append(state)
        yp = meanpool_emb(y_ids[s], E, be)
        yps.append(yp)
        _, state = multi_rnn_cell(yp, state, wc, "hier", num_layer)
        state = state * np.asarray(masks[s]).astype(dt)
    return np.stack(pre), state, np.stack(yps)


def tcn_hidden_restructured(x_ids, state_pre, w, precision="f32"):
    """Per slot: h0 = E[x] @ W_in[:D] + state_pre[s] @ W_in[D:]; conv stack.  Returns the S-list of
    Hout_s [B,L_s,C] (input of the catalog-scoring GEMM) and the S-list of sbias [B,C]."""
    dt = _dt(precision)
    E = w["hier/emb/kernel"].astype(dt)
    D = E.shape[1]
    w_in = w["hier/tcn/emb/kernel"].astype(dt)
    wt = {k: v.astype(dt) for k, v in w.items() if k.startswith("hier/tcn/temporal_conv_net")}
    n_levels = n_tcn_levels(w, "hier/tcn")
    houts, sbiases = [], []
    for s in range(len(x_ids)):
        xe = emb_gather(x_ids[s], E)                                  # bit-exact gather
        sbias = state_pre[s].astype(dt) @ w_in[D:]                    # fp32/fp64 in every tier
        h0 = _q(xe, precision) @ _q(w_in[:D], precision) + sbias[:, None, :]
        if precision == "bf16":
            h0 = bf16_round(h0)
        h = temporal_conv_net(h0, wt, "hier/tcn", n_levels, precision, round_between=True)
        houts.append(h)
        sbiases.append(sbias)
    return houts, sbiases


def score_catalog(hout, w, precision="f32"):
    """model_tcn.py:41 -- Z = Hout @ W_out + b_out."""
    dt = _dt(precision)
    return dense(_q(hout, precision), _q(w["hier/tcn/dense/kernel"].astype(dt), precision),
                 w["hier/tcn/dense/bias"].astype(dt))


def model_hier_restructured(x_ids, y_ids, masks, state, w, num_layer=2, precision="f32",
                            return_hidden=False, x_gap=None, gap_bandwidth=168.0, l2_norm=False):
    """Restructured route; same outputs as model_hier_literal (up to fp summation order)."""
    state_pre, state_out, _ = gru_over_sessions(y_ids, masks, state, w, num_layer,
                                                "f64" if precision == "f64" else "f32", x_gap, gap_bandwidth)
    houts, _ = tcn_hidden_restructured(x_ids, state_pre, w, precision)
    hout = np.concatenate(houts, 1)
    if return_hidden:
        return hout, state_out
    z = score_catalog(hout, w, precision)
    return (l2_normalize(z) if l2_norm else z), state_out


# --------------------------------------------------------------------------------------
# A.5 loss + metrics (model.py:98-117, loss.py:20-21, loss.py:163-221, loss.py:120)
# --------------------------------------------------------------------------------------


def softmax_cross_entropy_with_logits(labels_onehot_or_ids, logits):
    """loss.py:21.  Accepts int ids (0 = all-zero label row, giving 0 loss contribution
    from the label term ... see below) or a dense label tensor."""
    z = logits
    m = z.max(-1, keepdims=True)
    lse = (m + np.log(np.exp(z - m).sum(-1, keepdims=True)))[..., 0]
    lab = np.asarray(labels_onehot_or_ids)
    if lab.ndim == z.ndim:                                       # dense labels: -sum(lab * logsoftmax)
        return (lab * (lse[..., None] - z)).sum(-1)
    ids = lab.astype(np.int64)
    zy = np.take_along_axis(z, ids[..., None], -1)[..., 0]
    # label row is one_hot(id)*sign(id): for id 0 the label vector is all zero -> loss term 0
    return np.where(ids > 0, lse - zy, 0.0).astype(z.dtype)


def hier_loss(pred_all, y_id, mask_warmstart=None):
    """model.py:62,105-117: mask logits, CE, mask loss, per-user mean (+1e-6), mean over users with
    >=1 valid position.  Returns (loss scalar, loss_bt [B,T] masked, mask_y, activity_count(+1e-6), user_count)."""
    y_id = np.asarray(y_id).astype(np.int64)
    dt = pred_all.dtype
    mask_y = np.sign(y_id).astype(dt)                                         # model.py:62
    if mask_warmstart is not None:
        mask_y = mask_y * np.asarray(mask_warmstart).astype(dt)               # model.py:102-103
    pred = pred_all * mask_y[..., None]                                       # :105
    loss_bt = softmax_cross_entropy_with_logits(y_id, pred) * mask_y          # :108,111
    activity_count = mask_y.sum(1)                                            # :112
    user_count = np.sign(activity_count).sum()                                # :113
    activity_count = activity_count + dt.type(1e-6)                           # :114
    loss = ((loss_bt.sum(1) / activity_count).sum() / user_count).astype(dt)  # :116-117
    return loss, loss_bt, mask_y, activity_count, user_count, pred


def calc_metric_fast(score, mask_y, activity_count, user_count, y_id, item_num=None):
    """loss.py:163-221 (non-'mv' branch): rank = #{j: score_j > score_y} (strict, over all N columns),
    ranks_float = rank/N, rr = 1/(1+rank), recall@{1,5,10}; each masked, per-user mean, user mean.
    ``score`` must already be masked (model.py:105)."""
    y_id = np.asarray(y_id).astype(np.int64)
    N = score.shape[-1] if item_num is None else item_num
    dt = score.dtype
    target = np.take_along_axis(score, y_id[..., None], -1)      # == reduce_sum(score*y_onehot) :179
    target = np.where(y_id[..., None] > 0, target, 0).astype(dt)  # id 0: score*0 summed = 0
    ranks = (score > target).sum(-1).astype(dt)
    ranks_float = ranks / dt.type(N)                              # :190
    rr = 1.0 / (1 + ranks)                                        # :191
    rec1 = (ranks <= 0).astype(dt)                                # :194-196
    rec5 = (ranks <= 4).astype(dt)
    rec10 = (ranks <= 9).astype(dt)
    ranks = ranks * mask_y                                        # :199-205
    ranks_float = ranks_float * mask_y
    rr, rec1, rec5, rec10 = rr * mask_y, rec1 * mask_y, rec5 * mask_y, rec10 * mask_y

    def user_mean(a):                                             # :208-219
        return (a.sum(1) / activity_count).sum() / user_count

    return (user_mean(rec1), user_mean(rec5), user_mean(rec10), user_mean(rr), user_mean(ranks_float),
            ranks_float, ranks)


def top_k(score, k):
    """tf.nn.top_k ordering (loss.py:120) [TF-sem]: descending score, ties -> lower index first."""
    N = score.shape[-1]
    idx = np.arange(N)
    flat = score.reshape(-1, N)
    out_i = np.empty((flat.shape[0], k), dtype=np.int64)
    for r in range(flat.shape[0]):
        order = np.lexsort((idx, -flat[r].astype(np.float64)))
        out_i[r] = order[:k]
    out_v = np.take_along_axis(flat, out_i, 1)
    return out_v.reshape(score.shape[:-1] + (k,)), out_i.reshape(score.shape[:-1] + (k,))


def rank_ambiguity(score, y_id, rel_eps):
    """Per row: number of columns whose score is within rel_eps*max(1,|score_y|) of the target
    (excluding the target itself).  |rank_a - rank_b| between two fp implementations is bounded by this."""
    y_id = np.asarray(y_id).astype(np.int64)
    target = np.take_along_axis(score, y_id[..., None], -1)
    tol = rel_eps * np.maximum(1.0, np.abs(target))
    return (np.abs(score - target) <= tol).sum(-1) - 1


def rank_sign_bit(z32, zy32):
    """Bit-level restatement of the rank count of the bf16 tier's fused CE+rank sweep (k4_score_bf16.cu: kSignRank), the
    device-side form of `rank = #{j: Z[j] > Z[y]}` (loss.py:179): a logit is counted when the SIGN BIT of
        tn = fma_rn(z, -c, T),   c = fl32(log2 e),   T = fl32_round_up(z_y * c)
    is set.  fma rounds the exact value T - z*c once, and rounding never changes the sign of a non-zero value, so the
    sign bit is set iff T - z*c < 0 exactly.  z32 [Q,N] float32 logits, zy32 [Q] float32 target logits."""
    c = np.float64(np.float32(1.4426950408889634))
    z = np.asarray(z32, np.float32).astype(np.float64)
    zy = np.asarray(zy32, np.float32).astype(np.float64)
    p = zy * c                                                    # exact: 24 x 24 significant bits fit in a double
    t = p.astype(np.float32)
    t = np.where(t.astype(np.float64) < p, np.nextafter(t, np.float32(np.inf)), t).astype(np.float64)   # round up
    e = t[..., None] - z * c     # exact where it matters (|e| <= |t|, Sterbenz); elsewhere the double rounding keeps the sign
    return (e < 0).sum(-1)


# --------------------------------------------------------------------------------------
# sampled ranking losses + calc_score (loss.py:22-71, 76-105)
# --------------------------------------------------------------------------------------


def l2_normalize(x, axis=-1, eps=1e-12):
    """tf.nn.l2_normalize [TF-sem]: x * rsqrt(max(sum(x^2), eps))."""
    ss = (x * x).sum(axis, keepdims=True)
    return x / np.sqrt(np.maximum(ss, eps))


def calc_loss_sampled(pred, y, y_impression, loss="hinge_logsigmoid", num_neg_sample=20,
                      nce_weight=1, hinge_delta=0.1):
    """loss.py:18-71.  pred [B,T,d], y [B,T,d], y_impression [B,T,k,d] -> loss [B,T]."""
    dt = pred.dtype
    if loss == "l2":                                                          # :18-19
        return ((pred - y) ** 2).sum(2)
    p = l2_normalize(pred)
    inner = (p * y).sum(-1)                                                   # [B,T]
    inner_prod = np.einsum("btd,btkd->btk", p, y_impression)                 # [B,T,k]
    logsig = lambda v: -np.logaddexp(0, -v)  # noqa: E731  log(sigmoid(v)), stable
    if loss == "nce":                                                         # :22-31
        part_1 = logsig(inner)
        part_2 = logsig(-inner_prod).sum(2)
        return (-part_1 - part_2 / num_neg_sample * nce_weight).astype(dt)
    if loss == "hinge_sigmoid":                                               # :32-40
        diff = sigmoid(inner_prod) - sigmoid(inner)[..., None] + hinge_delta
        return relu(diff).mean(2).astype(dt)
    if loss == "hinge_logsigmoid":                                            # :42-50
        diff = logsig(inner_prod) - logsig(inner)[..., None] + hinge_delta
        return relu(diff).mean(2).astype(dt)
    if loss == "hinge_linear":                                                # :52-60
        diff = inner_prod - inner[..., None] + hinge_delta
        return relu(diff).mean(2).astype(dt)
    if loss == "bpr":                                                         # :62-70
        diff = sigmoid(inner)[..., None] - sigmoid(inner_prod)
        return (-logsig(diff).mean(2)).astype(dt)
    raise NotImplementedError(loss)


def calc_score(pred, y_impression, rank_metric="l2"):
    """loss.py:76-105 ('l2' and 'inner_prod' modes)."""
    if rank_metric == "l2":
        return -((pred[:, :, None, :] - y_impression) ** 2).sum(-1)
    if rank_metric == "inner_prod":
        return np.einsum("btd,btkd->btk", pred, y_impression)
    raise NotImplementedError(rank_metric)


# --------------------------------------------------------------------------------------
# full forward + loss + metrics, the fetch list of run_hier_xing.py:145-149
# --------------------------------------------------------------------------------------


def forward_loss_metrics(x_ids, y_ids, masks, state, w, num_layer=2, precision="f32", literal=False, x_gap=None,
                         gap_bandwidth=168.0, l2_norm=False, mask_warmstart=None):
    fwd = model_hier_literal if literal else model_hier_restructured
    pred_all, state_out = fwd(x_ids, y_ids, masks, state, w, num_layer, precision, x_gap=x_gap,
                              gap_bandwidth=gap_bandwidth, l2_norm=l2_norm)
    y_id = np.concatenate([np.asarray(y) for y in y_ids], 1).astype(np.int64)   # run_hier_xing.py:278
    loss, loss_bt, mask_y, act, ucount, pred_masked = hier_loss(pred_all, y_id, mask_warmstart)
    # calc_metric_fast reads the target score through the one-hot label tensor y (loss.py:179), which the warm-start mask
    # does not touch: a position masked only by mask_warmstart compares its zeroed scores with 0 and is then masked out
    rec1, rec5, rec10, mrr, mrp, ranks_float, ranks = calc_metric_fast(pred_masked, mask_y, act, ucount, y_id)
    return dict(loss=loss, loss_bt=loss_bt, state=state_out, recall1=rec1, recall5=rec5, recall10=rec10,
                mrr=mrr, mrp=mrp, ranks_float=ranks_float, ranks=ranks, mask_y=mask_y, pred=pred_masked)
